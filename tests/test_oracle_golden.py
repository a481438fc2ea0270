"""The C oracle (oracle/cf_oracle.c) against the golden vectors produced by the unmodified reference."""
import numpy as np
import pytest

from common import compare, golden_names, load_golden, problem_from_recipe
from oracle import cf_oracle as cfo


@pytest.mark.parametrize("name", golden_names())
def test_oracle_matches_reference_vector(name, fx):
    gold = load_golden(name)
    fl, cells, sp, g, tab, gla = problem_from_recipe(gold["recipe"], fx)
    assert list(sp["mcid"]) == list(gold["mcid"])                      # species order: integer bookkeeping, exact
    vah = gold["recipe"]["generator"] == "vah"
    dN, skipped, breakdown = cfo.smooth(fl, cells, sp, g, tab, gla, vah=vah)
    assert skipped == 0
    assert breakdown == int(gold["breakdown"])
    rep = compare(dN, gold["dN"], tol=1e-13)
    assert rep["ok"], rep
    if not vah:
        # the restatement keeps the reference's operation order: in practice it is bit-identical
        assert np.array_equal(dN, gold["dN"]), rep
    else:
        # alpha_L / Lambda are re-derived by the reference's own reader (explicit powers vs Horner: 1-ulp inputs)
        assert rep["max_rel"] < 1e-12, rep


def test_toy_cell_closed_form(fx):
    """Analytic known-answer test on the shipped one-cell surface."""
    gold = load_golden("toy_df1")
    hb = 0.197327053
    T = 0.786 * hb
    pT = fx["pT_tab"][:, 0]; y = fx["y_tab"][:, 0]
    a = gold["dN"].reshape(21, 24, 32, 3)
    for isp, (m, gdeg, theta) in enumerate(((0.138, 1.0, -1.0), (0.494, 1.0, -1.0), (0.938, 2.0, 1.0))):
        mT = np.sqrt(m * m + pT ** 2)
        x = mT[None, :] * np.cosh(y)[:, None] / T
        with np.errstate(over="ignore"):
            ref = gdeg * 1000.0 * mT[None, :] * np.cosh(y)[:, None] / ((2 * np.pi * hb) ** 3 * (np.exp(x) + theta))
        got = a[:, 0, :, isp]
        nz = ref > 0
        assert np.max(np.abs(got[nz] - ref[nz]) / ref[nz]) < 5e-15
        assert np.all(got[~nz] == 0)
    assert abs(a[10, 0, 0, 0] - 50.47391543445407) < 1e-12            # pi+ (y = 0, phi_0, pT_0)


def test_skipped_cells_contribute_zero(fx):
    """u.dsigma <= 0 cells: the intended semantics -- they add exactly nothing."""
    gold = load_golden("s3_df1")
    fl, cells, sp, g, tab, gla = problem_from_recipe(gold["recipe"], fx)
    bad = {k: np.concatenate([v, v[:5]]) for k, v in cells.items()}
    bad["dat"][-5:] *= -1.0; bad["dax"][-5:] *= -1.0; bad["day"][-5:] *= -1.0; bad["dan"][-5:] *= -1.0
    dN, skipped, _ = cfo.smooth(fl, bad, sp, g, tab, gla)
    assert skipped == 5
    assert np.array_equal(dN, gold["dN"])
