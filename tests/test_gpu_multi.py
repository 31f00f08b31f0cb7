"""GPU (needs >= 2 devices): ONE scenario tree partitioned across GPUs (rapidnet_b200/partition.py, rn_dist_*) against the
same tree on one GPU.  tools/dist_check.py runs under torchrun, one process per GPU; the ranks exchange the chain heads
and the prox distances inside the persistent kernel over NVLink peer memory.  Tolerance: norm-wise 1e-5 up to 10
iterations, 1e-4 at 100 (the two runs differ only by fp32 rounding order in the zeta correction)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("workload,world,factors", [("C1", 2, "full"), ("C1r30", 2, "full"), ("C2", 2, "full"), ("C2", 2, "shared")])
def test_partitioned_tree_matches_single_gpu(workload, world, factors):
    if _ngpu() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tools", "dist_check.py"), "--workload", workload, "--iters", "1,10,100",
           "--factors", factors]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert "DIST_CHECK OK" in out.stdout, (out.stdout[-2000:], out.stderr[-2000:])
