"""The algebra behind cf_shift.cu, checked in numpy on the CPU (no GPU, no product code): the shifted factorisation
e^{-x_jk} = e^{-(mT A_j - pT Bmax)} e^{-pT (Bmax - B_k)} keeps both factors in (0, 1], agrees with the direct exponential to
~1e-13 relative, and the "rare" criterion xm + pT (Bmax - Bmin) < 707.7 guarantees that neither factor nor the product leaves
the normal range -- on the cfg3 surface with the full species list and the reference's momentum tables."""
import numpy as np

from is3d_b200 import synthetic, tables

RARE_X = float(np.frombuffer(np.array([0x40861D9900000000], dtype=np.uint64).tobytes(), dtype=np.float64)[0])     # high word kRareHi
ALIVE_X = 709.782712893384           # ln(DBL_MAX): beyond it the reference's exp(x) overflows and the term is exactly 0
TINY = np.finfo(float).tiny


def _problem(n_cells=300):
    fx = tables.load_fixture()
    sp = tables.species(fx, 1, "chosen_urqmd"); g = tables.grid(fx)
    cells = synthetic.columns_to_cells(synthetic.surface_vh(n_cells, synthetic.SEEDS["cfg3"], three_d=True, viscous=True), 1)
    tau, eta, ux, uy, un, T = [np.asarray(cells[k]) for k in ("tau", "eta", "ux", "uy", "un", "T")]
    ut = np.sqrt(1.0 + ux * ux + uy * uy + tau * tau * un * un)
    y = np.asarray(g["y"]); phi = np.asarray(g["phi"]); pT = np.asarray(g["pT"])
    d = y[None, :] - eta[:, None]
    A = (np.cosh(d) * ut[:, None] - tau[:, None] * np.sinh(d) * un[:, None]) / T[:, None]            # [cell, slot]
    B = (np.cos(phi)[None, :] * ux[:, None] + np.sin(phi)[None, :] * uy[:, None]) / T[:, None]        # [cell, phi]
    mass = np.asarray(sp["mass"])
    return A, B, mass, pT


def test_shifted_factorisation_matches_direct_exponential():
    A, B, mass, pT = _problem()
    npt = 3
    Bt = B.reshape(B.shape[0], -1, npt)                      # [cell, phi tile, k]
    bmax = Bt.max(axis=2); bmin = Bt.min(axis=2)
    worst = 0.0; checked = 0
    for m in mass[::19]:                                     # 17 species across the mass range
        mT = np.sqrt(m * m + pT * pT)                        # [pT]
        for tile in (0, 3, 7):
            qm = pT[None, :] * bmax[:, tile, None]                                    # [cell, pT]
            xm = mT[None, None, :] * A[:, :, None] - qm[:, None, :]                   # [cell, slot, pT]: smallest argument of the group
            dB = pT[None, :] * (bmax[:, tile] - bmin[:, tile])[:, None]               # [cell, pT]
            fast = (xm <= ALIVE_X) & (xm + dB[:, None, :] < RARE_X)                   # groups the product form is used for
            eA = np.exp(-np.where(fast, xm, 0.0))
            for k in range(npt):
                dk = pT[None, :] * (bmax[:, tile] - Bt[:, tile, k])[:, None]          # [cell, pT] >= 0
                assert np.all(dk >= 0.0)
                eB = np.exp(-dk)
                xk = mT[None, None, :] * A[:, :, None] - (pT[None, :] * Bt[:, tile, k][:, None])[:, None, :]
                direct = np.exp(-np.where(fast, xk, 0.0))
                prod = eA * eB[:, None, :]
                # both factors and the product are normal numbers wherever the fast path is taken
                assert np.all(eA[fast] >= TINY) and np.all(np.broadcast_to(eB[:, None, :], fast.shape)[fast] >= TINY) and np.all(prod[fast] >= TINY)
                assert np.all(eA[fast] <= 1.0) and np.all(eB <= 1.0)
                rel = np.abs(prod[fast] - direct[fast]) / direct[fast]
                worst = max(worst, float(rel.max())); checked += int(fast.sum())
    print("checked %d evaluations, max relative deviation of the product form %.3g" % (checked, worst))
    assert checked > 1_000_000 and worst < 1e-12


def test_group_minimum_classifies_the_group():
    """xm = mT A - pT Bmax is the smallest argument of the group: a dead xm means every member is dead (exact zeros), and
    xm >= 12.5 / 37.5 puts every member in the dilute / ultra-dilute regime of the quantum-statistics factor."""
    A, B, mass, pT = _problem(120)
    Bt = B.reshape(B.shape[0], -1, 3)
    bmax = Bt.max(axis=2)
    m = mass[40]; mT = np.sqrt(m * m + pT * pT)
    for tile in range(Bt.shape[1]):
        xm = mT[None, None, :] * A[:, :, None] - (pT[None, :] * bmax[:, tile, None])[:, None, :]
        for k in range(3):
            xk = mT[None, None, :] * A[:, :, None] - (pT[None, :] * Bt[:, tile, k][:, None])[:, None, :]
            assert np.all(xk >= xm - 1e-9 * np.abs(xm))
    # the truncated occupation factors: |1/(1 + a) - (1 - a + a^2)| < a^3 for a < 2^-18, and 1/(1 + a) rounds to 1 for a < 2^-54
    a = np.exp(-np.linspace(12.5, 37.5, 1000))
    assert np.all(np.abs(1.0 / (1.0 + a) - (1.0 - a + a * a)) <= a ** 3 * 1.01 + 3.4e-16)     # + 1.5 ulp for the three roundings of this check itself
    a = np.exp(-np.linspace(37.5, 700.0, 1000))
    assert np.all(1.0 / (1.0 + a) == 1.0) and np.all(1.0 / (1.0 - a) == 1.0)
