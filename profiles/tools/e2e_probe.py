import sys, time, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', '..'))
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
n = int(os.environ.get('E2E_PROBE_STEPS', 6))
print([round(step()*1e3,1) for _ in range(n)])
if os.environ.get('E2E_PROBE_PROFILE'):
    import cProfile, pstats
    cProfile.run('step()', '/tmp/prof.out')
    pstats.Stats('/tmp/prof.out').sort_stats('cumtime').print_stats(12)
