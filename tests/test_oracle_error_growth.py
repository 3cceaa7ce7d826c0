"""Error-growth tracking of stored solutions (src/adaptive_spectral_integration.F90: need_new_solution, polyfit;
restore_solution's time / maxnorm bookkeeping, src/pprts.F90:3991-4049) in the oracle.

Restates tests/pprts_error_growth_tracking/pprts_error_growth_tracking.F90:49-128 (options: -max_solution_time 60,
-solar_diff_ksp_rtol 1e-4): a band that is solved again and again with unchanged inputs has to be recomputed in the
first three time steps (no history yet) and not afterwards."""
import numpy as np
import pytest

from oracle import binding as ob


def test_polyfit_known_answers():
    """polyfit (:174-257) is an ordinary least-squares fit through the normal equations"""
    x = np.array([0., 1., 2., 3.])
    c, ierr = ob.polyfit(x, 2 + 3 * x, 1)
    assert ierr == 0 and c == pytest.approx([2., 3.], abs=1e-12)
    c, ierr = ob.polyfit(x, 1 - x + 0.5 * x ** 2, 2)
    assert ierr == 0 and c == pytest.approx([1., -1., 0.5], abs=1e-11)
    y = np.array([0., 1., 0., 1.])                  # not on a line: least squares, same as numpy
    c, ierr = ob.polyfit(x, y, 1)
    assert c == pytest.approx(np.polyfit(x, y, 1)[::-1], abs=1e-12)
    c, ierr = ob.polyfit(np.array([0., 1., 1., 2.]), y, 1)       # equal abscissae: singular (:198-207)
    assert ierr == 1


def run_tracking(make, solve, need):
    kabs = np.full((3, 3, 3), 1. / 3)
    ksca = np.full((3, 3, 3), 1. / 3)
    g = np.zeros((3, 3, 3))
    s = make()
    needs = {}
    for it in range(1, 6):
        for k in (1, 2):
            needs[(it, k)] = need(s, k, float(it))
            solve(s, kabs, ksca, g, k, float(it))
    return s, needs


def test_error_growth_tracking(oracle_lut):
    """pprts_error_growth_tracking.F90:66-125"""
    def make():
        return ob.OracleSolver(oracle_lut, 3, 3, 3, 100., 100., ob.sundir_from_angles(0., 60.), np.full(3, 100.),
                               diff_solver=1, rtol=1e-4, max_solution_time=60.)

    def solve(s, kabs, ksca, g, k, t):
        s.set_optical_properties(0.1, kabs, ksca, g)
        assert s.solve(False, True, 1000., k, time=t) == 0
        s.get_result(k)

    s, needs = run_tracking(make, solve, lambda s, k, t: s.need_new_solution(k, t))
    for (it, k), lneed in needs.items():
        assert lneed == (it <= 3), (it, k)
    t, m = s.solution_times(1)
    assert list(t[:6]) == [5., 4., 3., 2., 1., -1.]
    assert m[4] > 1e-3 and np.all(m[:4] < 1e-3 * m[4])       # the first solve changed abso from zero, the others barely
    assert s.need_new_solution(1, None)                        # time not provided -> always recompute
    assert s.need_new_solution(1, 5. + 61.)                    # older than max_solution_time
    assert s.need_new_solution(7, 6.)                          # never solved
    s.destroy()


def test_error_estimates_disabled_by_default(oracle_lut):
    """-max_solution_time defaults to 0 (src/tenstream_options.F90:94): lenable_solutions_err_estimates is false"""
    s = ob.OracleSolver(oracle_lut, 3, 3, 3, 100., 100., ob.sundir_from_angles(0., 60.), np.full(3, 100.), diff_solver=1)
    kabs = np.full((3, 3, 3), 1. / 3)
    s.set_optical_properties(0.1, kabs, kabs, np.zeros((3, 3, 3)))
    for it in range(1, 6):
        assert s.need_new_solution(1, float(it))
        assert s.solve(False, True, 1000., 1, time=float(it)) == 0
    t, m = s.solution_times(1)
    assert list(t[:5]) == [5., 4., 3., 2., 1.] and np.all(m == 0)   # times are kept, norms are not computed
    s.destroy()


def test_growing_error_triggers_recompute(oracle_lut):
    """When the inputs drift, the fitted error curve extrapolates above -max_solution_err and the band is recomputed"""
    s = ob.OracleSolver(oracle_lut, 3, 3, 3, 100., 100., ob.sundir_from_angles(0., 60.), np.full(3, 100.),
                        diff_solver=1, rtol=1e-6, max_solution_time=600.)
    g = np.zeros((3, 3, 3))
    for it in range(1, 5):
        k = np.full((3, 3, 3), 2e-3 * (1. + 0.5 * it))
        s.set_optical_properties(0.1, k, k, g)
        assert s.solve(False, True, 1000., 1, time=float(it)) == 0
    assert s.need_new_solution(1, 5.)
    s.destroy()
