"""Comparison rules of the reference's own tests, restated.

* engine_close : abs tol 1e-2                      (/root/reference/src/test/Testing.cu:33-76)
* smpc_close   : abs 0.1, or 0.1 % relative when |value| > 100
                 (/root/reference/src/test/TestSmpcController.cu:28-47)

Parity gates for whole solves.  The APG iteration amplifies rounding differences: after 500
iterations two correct fp32 implementations that merely sum in a different order (cuBLAS vs plain loops) differ by a
few 1e-4, and each is about that far from the same algorithm run in double.  Agreement with the reference build is
therefore bounded by the reference's OWN rounding uncertainty, and the gates are stated on accuracy:

  accuracy_gate : err(ours, f64) <= ACC_FACTOR * err(reference, f64) + ACC_ABS          per variable, and
                  median over the variables of a case of err(ours, f64) / err(reference, f64) <= ACC_MEDIAN
                  Both errors are single samples of a random quantity (rounding noise amplified by the iteration), so the
                  per-variable ratio scatters: measured on B200 over 6 configs x 4 iteration counts x 13 variables against the
                  reference build (profiles/r02_parity_table.md) the median ratio is 1.0, ours is closer to double than the
                  reference in half of the cells, and the largest ratio is 3.2 (toy, X, 500 iterations: 5.3e-4 against
                  1.7e-4; it was 2.4 one summation-order change earlier, and in the same solve ours is 1.5-2x CLOSER on all
                  eight dual variables and on u0).  ACC_FACTOR = 4 bounds the scatter of one variable, ACC_MEDIAN = 1.5 is
                  the statement that matters: no systematic loss of accuracy against the reference's build.
                  ACC_ABS = 1e-5 keeps differences at the rounding floor from being ranked (iteration 1: y = step (Hx - z)
                  cancels 40:1, 4.8e-6 against 1.1e-6).
  u0 / iterates : err(ours, reference) <= RTOL = 1e-4 (the north-star figure) wherever the reference's floor
                  err(reference, f64) allows it (<= RTOL / (1 + ACC_FACTOR)); otherwise the bound that follows from the
                  accuracy gate by the triangle inequality, (1 + ACC_FACTOR) * err(reference, f64), and the case is reported
                  as floor-limited.  `floor_tol` computes it.
Against the ORACLE PORT (plain sequential loops in the reference's operation order) the accuracy factor is ACC_FACTOR_PORT = 6:
the port is not the target, and on these cases it is up to 4x closer to the double trajectory than the reference's own
build (C3, X, 500 iterations: port 9.5e-4, reference build 3.7e-3, ours 3.7e-3).
tools/parity_table.py prints all three errors per config x iteration count x variable (profiles/r02_parity_table.md).
"""
import numpy as np

RTOL = 1e-4
ACC_FACTOR = 4.0
ACC_MEDIAN = 1.5
ACC_FACTOR_PORT = 6.0
ACC_ABS = 1e-5
KAPPA = 1.0 + ACC_FACTOR


def floor_tol(ref32, ref64, den=None, rtol=RTOL, kappa=KAPPA):
    """tolerance for comparing against `ref32`, given the same quantity from the double-precision oracle"""
    a = np.asarray(ref32, dtype=np.float64).reshape(-1)
    b = np.asarray(ref64, dtype=np.float64).reshape(-1)
    d = float(np.linalg.norm(b)) if den is None else float(den)
    floor = float(np.linalg.norm(a - b) / max(d, 1e-30))
    return max(rtol, kappa * floor), floor


def accuracy_gate(ours, ref32, ref64, den=None, factor=ACC_FACTOR, slack=ACC_ABS):
    """(ok, err(ours, f64), err(ref, f64)): ours is at most `factor` times as far from the double-precision trajectory as
    the reference is (norm-wise relative; `den` overrides the denominator for differences of nearly equal iterates)"""
    o = np.asarray(ours, dtype=np.float64).reshape(-1)
    a = np.asarray(ref32, dtype=np.float64).reshape(-1)
    b = np.asarray(ref64, dtype=np.float64).reshape(-1)
    d = max(float(np.linalg.norm(b)) if den is None else float(den), 1e-30)
    e_ours, e_ref = float(np.linalg.norm(o - b) / d), float(np.linalg.norm(a - b) / d)
    return e_ours <= factor * e_ref + slack, e_ours, e_ref


def median_ratio(pairs, slack=ACC_ABS):
    """median over (err(ours, f64), err(reference, f64)) pairs of ours / reference; pairs at the rounding floor
    (both below `slack`) count as 1"""
    r = [1.0 if max(a, b) <= slack else a / max(b, 1e-30) for a, b in pairs]
    return float(np.median(r)) if r else 1.0


def engine_close(got, want, tol=1e-2):
    got = np.asarray(got, dtype=np.float64).reshape(-1)
    want = np.asarray(want, dtype=np.float64).reshape(-1)
    assert got.shape == want.shape, (got.shape, want.shape)
    err = np.abs(got - want)
    bad = np.nonzero(~(err < tol))[0]
    assert bad.size == 0, f"{bad.size} mismatches, first idx {bad[:5]} got {got[bad[:5]]} want {want[bad[:5]]}"
    return float(err.max()) if err.size else 0.0


def smpc_close(got, want):
    got = np.asarray(got, dtype=np.float64).reshape(-1)
    want = np.asarray(want, dtype=np.float64).reshape(-1)
    assert got.shape == want.shape, (got.shape, want.shape)
    diff = got - want
    big = np.abs(got) > 1e2
    measure = np.where(big, diff / np.where(big, got, 1.0) * 100.0, diff)
    bad = np.nonzero(~(np.abs(measure) < 1e-1))[0]
    assert bad.size == 0, f"{bad.size} mismatches, first idx {bad[:5]} got {got[bad[:5]]} want {want[bad[:5]]}"
    return float(np.abs(measure).max()) if measure.size else 0.0


def rel_err(got, want):
    got = np.asarray(got, dtype=np.float64).reshape(-1)
    want = np.asarray(want, dtype=np.float64).reshape(-1)
    return float(np.linalg.norm(got - want) / max(np.linalg.norm(want), 1e-30))


def pinf_close(got, want, strict=100, frac=0.9):
    """vecPrimalInfs (SmpcController.cu:1487-1495) is the SIGNED residual at the arg-max-abs index, max over the
    xi / psi blocks: a discontinuous function of the iterates (two near-equal |residuals| of opposite sign swap on a
    rounding difference; the fp32 and fp64 oracles already disagree on a few entries of a 500-iteration log).  So:
    the first `strict` iterations must agree entry by entry (rtol 1e-3, atol 1e-2); later on at least `frac` (90 %;
    measured 94 % against the reference build on C1 at 500 iterations) of the entries must agree within rtol 1e-2,
    atol 1e-1."""
    got = np.asarray(got, dtype=np.float64).reshape(-1)
    want = np.asarray(want, dtype=np.float64).reshape(-1)[: got.size]
    n = min(strict, got.size)
    assert np.allclose(got[:n], want[:n], rtol=1e-3, atol=1e-2), (got[:n][~np.isclose(got[:n], want[:n], rtol=1e-3, atol=1e-2)][:5],
                                                               want[:n][~np.isclose(got[:n], want[:n], rtol=1e-3, atol=1e-2)][:5])
    if got.size > n:
        ok = np.isclose(got[n:], want[n:], rtol=1e-2, atol=1e-1)
        assert ok.mean() >= frac, f"only {ok.mean():.3f} of the late primal-infeasibility entries agree"
