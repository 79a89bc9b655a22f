"""N>1 path.  CPU: world_size-2 gloo run of the host-side sharding + delta-combination logic against the oracle.
GPU (-m gpu, needs >= 2 devices): the real NCCL path of rfm_fit, one process per GPU."""
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def _launch(mode, world, port):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(HERE, "multigpu_worker.py"), mode]
    env = dict(os.environ, OMP_NUM_THREADS="1")
    return subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)


def test_world2_gloo_sharding_and_delta_sum():
    out = _launch("gloo-oracle", 2, 29631)
    assert out.returncode == 0 and "gloo-oracle ok" in out.stdout, out.stdout[-2000:] + out.stderr[-4000:]


@pytest.mark.gpu
def test_world2_nccl_fit():
    from rankfm_b200 import _rankfm
    if _rankfm.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    out = _launch("nccl", 2, 29632)
    assert out.returncode == 0 and "nccl ok" in out.stdout, out.stdout[-2000:] + out.stderr[-4000:]
