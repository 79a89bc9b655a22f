"""Builds ``librankfm_b200.so`` (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

``python -m rankfm_b200.build`` or ``rankfm_b200.build.build()``.  The ``.so`` is git-ignored but travels to the
GPU box with the snapshot.  nvcc cross-compiles without a GPU.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "librankfm_b200.so")
SOURCES = ["rfm_api.cu", "rfm_comm.cu", "rfm_train.cu", "rfm_score.cu", "rfm_pack.cu", "rfm_gemm.cu", "rfm_prep.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
              "-diag-suppress", "63"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "rankfm_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force=False, verbose=False, out=None, extra_flags=()):
    """`out` / `extra_flags`: build a VARIANT of the library next to the product one (A/B experiments, e.g.
    extra_flags=["-DRFM_SAMPLER_ROUNDS"]); it is loaded with RANKFM_B200_LIB=<path>"""
    if out is None and not force and not needs_build():
        return LIB
    objdir = os.path.join(HERE, "build" if out is None else "build_" + os.path.splitext(os.path.basename(out))[0])
    os.makedirs(objdir, exist_ok=True)
    nvcc = _nvcc()
    objs, procs = [], []
    for src in SOURCES:
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        cmd = [nvcc] + NVCC_FLAGS + list(extra_flags) + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for src, pr in procs:
        log, _ = pr.communicate()
        if verbose or pr.returncode != 0:
            sys.stderr.write(log)
        if pr.returncode != 0:
            raise RuntimeError("nvcc failed on %s" % src)
    target = LIB if out is None else out
    subprocess.run([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", target] + objs + ["-ldl"], check=True)
    return target


if __name__ == "__main__":
    variant = [a for a in sys.argv[1:] if a.startswith("--out=")]
    flags = [a for a in sys.argv[1:] if a.startswith("-D")]
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, out=variant[0][6:] if variant else None, extra_flags=flags))
