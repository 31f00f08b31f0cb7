"""GPU: tree shapes at the edges of what the persistent kernel handles, against the CPU oracle (same tolerance as
tests/test_gpu_parity.py), in the streamed and in the shared-factor formulation:
  * one scenario (no branching at all): no crown, the whole tree is a single chain of 24 columns, 2 grid barriers;
  * five binary branching stages: the deepest crown of the workloads (63 crown nodes above 32 chains of 19 stages);
  * the reference's minimal shape (branching at stage 1 only) with many children: 40 chains under one root."""
import pytest

from rapidnet_b200 import cabi
from rapidnet_b200.datagen import SEED0, barcelona_problem, make_tree
from test_gpu_parity import _check_u0, _compare_state, _setup, _setup64

pytestmark = pytest.mark.gpu

# (a non-branching stage ABOVE a branching one is outside the reference's domain: its final-branch node is the first
# stage with as many nodes as the next one, ScenarioTree.cu:149-169, and Omega / Theta alias from there on)
SHAPES = {"single_scenario": [], "deep_crown": [2, 2, 2, 2, 2], "wide_root": [40]}


@pytest.mark.parametrize("factors", [cabi.FACTORS_FULL, cabi.FACTORS_SHARED], ids=["full", "shared"])
@pytest.mark.parametrize("shape", sorted(SHAPES))
def test_edge_tree_matches_oracle(shape, factors):
    tree = make_tree(SHAPES[shape], 24, 88, 114, SEED0 + 40)
    prob = barcelona_problem(tree, seed=SEED0 + 41, max_iter=100)
    s, o = _setup(prob, cabi.SWEEP_PERSISTENT, factors)
    info = s.info()
    assert info.sweep_mode == cabi.SWEEP_PERSISTENT and info.factor_mode == factors
    o64 = _setup64(prob, s)
    for iters in (1, 10, 100):
        u0, _ = s.apg_solve(iters)
        o.apg(iters)
        o64.apg(iters)
        worst = _compare_state(s, o, f"{shape} it={iters}", o64)
        _check_u0(u0, o, o64, f"{shape} it={iters}")
        print(f"{shape} it={iters}: worst rel err {worst:.2e}")
    assert s.info().launches_per_iteration == 0      # the persistent kernel ran, not the per-stage fallback
    s.close(); o.close(); o64.close()
