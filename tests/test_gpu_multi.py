"""GPU (needs >= 2 devices): ONE scenario tree partitioned across GPUs (rapidnet_b200/partition.py, rn_dist_*) against the
same tree on one GPU.  tools/dist_check.py runs under torchrun, one process per GPU; the ranks exchange the per-parent
sums of the chain heads and the prox distances inside the persistent kernel over NVLink peer memory.  The cut is aligned
to the bottom-crown nodes, every exchanged sum is formed by one rank in the order of the single-GPU solve, so the iterates
are expected to be bit-identical; the gate is norm-wise 1e-6 at 1, 10, 100 and 500 iterations."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("workload,world,factors", [("C1r6", 2, "full"), ("C1r30", 2, "full"), ("C2", 2, "full"), ("C2", 2, "shared"),
                                                     ("C3", 4, "full"), ("C3", 8, "full")])
def test_partitioned_tree_matches_single_gpu(workload, world, factors):
    if _ngpu() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tools", "dist_check.py"), "--workload", workload, "--iters", "1,10,100,500",
           "--factors", factors]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert "DIST_CHECK OK" in out.stdout, (out.stdout[-2000:], out.stderr[-2000:])
