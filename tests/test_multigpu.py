"""N>1 path.  CPU: world_size-2 gloo run of the host-side sharding + delta-combination logic against the oracle.
GPU (-m gpu, needs >= 2 devices): the real thing, one process per GPU -- the fused peer-memory exchange and the NCCL
fallback, resident session + cached communicator across `_fit` calls, hold-out hit-rate gate against one GPU.
RANKFM_TEST_WORLD (default 2) sets the number of GPUs (run with 4 / 8 through gpurun --gpus)."""
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def _launch(mode, world, port, extra_env=None):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(HERE, "multigpu_worker.py"), mode]
    env = dict(os.environ, OMP_NUM_THREADS="1", **(extra_env or {}))
    return subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)


def test_world2_gloo_sharding_and_delta_sum():
    out = _launch("gloo-oracle", 2, 29631)
    assert out.returncode == 0 and "gloo-oracle ok" in out.stdout, out.stdout[-2000:] + out.stderr[-4000:]


@pytest.mark.gpu
@pytest.mark.parametrize("exchange", ["p2p", "nccl"])
def test_multi_gpu_fit(exchange):
    from rankfm_b200 import _rankfm
    world = int(os.environ.get("RANKFM_TEST_WORLD", "2"))
    if _rankfm.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    out = _launch("gpu", world, 29632 if exchange == "p2p" else 29633, {"RANKFM_B200_EXCHANGE": exchange})
    assert out.returncode == 0 and "gpu ok" in out.stdout, out.stdout[-2000:] + out.stderr[-4000:]
    print(out.stdout[-1500:])
