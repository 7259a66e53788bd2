"""Multi-GPU parity (needs >= 2 GPUs on the box; skipped otherwise): torchrun, one rank per GPU,
NCCL halo exchange inside libbellman.so, each slab bit-compared with the oracle."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("halo", ["fused_p2p", "nccl_sendrecv"])
@pytest.mark.parametrize("kind", ["kirk", "kirk_odd", "attitude", "pos_att"])
def test_slab_partitioned_sweep_over_nccl(kind, halo):
    n = _ngpu()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    n = 2 if n < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
           "--master-addr", "127.0.0.1", "--master-port", "29541",
           os.path.join(ROOT, "scripts", "multi_gpu_check.py"), kind]
    env = dict(os.environ)
    if halo == "nccl_sendrecv":
        env["BELLMAN_NO_P2P"] = "1"
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert "MULTI_GPU_CHECK PASS" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
