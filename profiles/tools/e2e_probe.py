import sys, time, os
sys.path.insert(0, '/root/repo')
import numpy as np
from bench import make_workload, fresh_weights, HYPER, WEIGHTS
from rankfm_b200 import _rankfm
c = make_workload('cfg2')
X = c['X']; ui = _rankfm.UserItems.from_interactions(X, c['U_global'])
def step():
    ww = fresh_weights(c)
    t0 = time.perf_counter()
    _rankfm._fit(X, c["sw"], ui, c["x_uf"], c["x_if"], *[ww[k] for k in WEIGHTS], HYPER["alpha"], HYPER["beta"], HYPER["learning_rate"],
                 HYPER["learning_schedule"], HYPER["learning_exponent"], c["max_samples"], c['epochs'], False)
    return time.perf_counter() - t0
print([round(step()*1e3,1) for _ in range(6)])
import cProfile, pstats
cProfile.run('step()', '/tmp/prof.out')
pstats.Stats('/tmp/prof.out').sort_stats('cumtime').print_stats(12)
