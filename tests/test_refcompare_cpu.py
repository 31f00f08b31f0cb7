"""CPU: the parity gates of tests/refcompare.py behave as DESIGN.md section 5 states them."""
import numpy as np

from refcompare import ACC_ABS, ACC_FACTOR, ACC_MEDIAN, RTOL, accuracy_gate, floor_tol, median_ratio


def test_accuracy_gate_is_a_bar_on_accuracy_not_on_agreement():
    rng = np.random.default_rng(1)
    exact = rng.standard_normal(1000)
    ref = exact * (1 + 1e-4 * rng.standard_normal(1000))          # the reference: 1e-4 from exact arithmetic
    ours = exact * (1 + 1e-4 * rng.standard_normal(1000))         # ours: as accurate, but 1.4e-4 away from the reference
    ok, e_ours, e_ref = accuracy_gate(ours, ref, exact)
    assert ok and 0.5 < e_ours / e_ref < 2.0
    worse = exact * (1 + (ACC_FACTOR + 1.5) * 1e-4 * rng.standard_normal(1000))
    assert not accuracy_gate(worse, ref, exact)[0]
    # differences at the rounding floor are not ranked
    assert accuracy_gate(exact + 0.5 * ACC_ABS * np.abs(exact), exact, exact)[0]


def test_floor_tol_is_the_north_star_figure_unless_the_reference_floor_forbids_it():
    exact = np.ones(100)
    tol, floor = floor_tol(exact * (1 + 1e-6), exact)
    assert tol == RTOL and abs(floor - 1e-6) < 1e-9
    tol, floor = floor_tol(exact * (1 + 1e-3), exact)
    assert abs(tol - (1 + ACC_FACTOR) * 1e-3) < 1e-9


def test_median_ratio_ignores_cells_at_the_rounding_floor():
    assert median_ratio([(2e-4, 1e-4), (1e-4, 1e-4), (0.5e-4, 1e-4)]) == 1.0
    assert median_ratio([(4e-6, 1e-6), (4e-6, 1e-6), (1e-4, 1e-4)]) == 1.0      # the first two are below ACC_ABS
    assert median_ratio([(3e-4, 1e-4)] * 3) > ACC_MEDIAN
