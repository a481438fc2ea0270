"""NumPy model of the factored 14-moment kernel (cf_factored.cu) -- used on CPU to check the ALGEBRA against the oracle before
spending GPU time: e^{-x} = e^{-mT A_j} e^{+pT B_k}, the (u.p)^2 bulk term merged into the bilinear delta-f form, g = 1 + df clamped
to [0, 2], dilute / ultra-dilute occupation factors.  Development tool, not part of the product or the tests.

    python tools/model_factored.py [n_cells] [chosen]
"""
import sys
import os

import numpy as np
from scipy.interpolate import CubicSpline

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from is3d_b200 import synthetic, tables      # noqa: E402
from oracle import cf_oracle as cfo          # noqa: E402


def prepare(cells, tab, g):
    T = cells["T"]; P = cells["P"]; E = cells["E"]
    tau = cells["tau"]; tau2 = tau * tau
    ux, uy, un = cells["ux"], cells["uy"], cells["un"]
    ut = np.sqrt(1.0 + ux * ux + uy * uy + tau2 * un * un)
    utperp = np.sqrt(1.0 + ux * ux + uy * uy)
    pixx, pixy, pixn, piyy, piyn = (cells[k] for k in ("pixx", "pixy", "pixn", "piyy", "piyn"))
    pinn = (pixx * (ux * ux - ut * ut) + piyy * (uy * uy - ut * ut) + 2.0 * (pixy * ux * uy + tau2 * un * (pixn * ux + piyn * uy))) / (tau2 * utperp * utperp)
    pitn = (pixn * ux + piyn * uy + tau2 * pinn * un) / ut
    pity = (pixy * ux + piyy * uy + tau2 * piyn * un) / ut
    pitx = (pixx * ux + pixy * uy + tau2 * pixn * un) / ut
    pitt = (pitx * ux + pity * uy + tau2 * pitn * un) / ut
    T4 = T ** 4
    c0 = CubicSpline(tab["T"], tab["c0"], bc_type="natural")(T) / T4
    c2 = CubicSpline(tab["T"], tab["c2"], bc_type="natural")(T) / T4
    sc = 0.5 / (T * T * (E + P))
    K0 = cells["bulkPi"] * (c0 - c2)
    K2 = cells["bulkPi"] * (4.0 * c2 - c0) * T * T
    y = g["y"]
    d = y[None, :] - cells["eta"][:, None]
    ch = np.cosh(d); sh = np.sinh(d); tsh = tau[:, None] * sh
    A = (ch * ut[:, None] - tsh * un[:, None]) / T[:, None]
    Cp = ch * cells["dat"][:, None] + (sh / tau[:, None]) * cells["dan"][:, None]
    Qyy = sc[:, None] * (pitt[:, None] * ch * ch + pinn[:, None] * tsh * tsh - 2.0 * pitn[:, None] * tsh * ch)
    cs = np.cos(g["phi"]); sn = np.sin(g["phi"])
    B = (cs[None, :] * ux[:, None] + sn[None, :] * uy[:, None]) / T[:, None]
    D = cs[None, :] * cells["dax"][:, None] + sn[None, :] * cells["day"][:, None]
    Qpp = sc[:, None] * (pixx[:, None] * cs * cs + piyy[:, None] * sn * sn + 2.0 * pixy[:, None] * cs * sn)
    R1 = 2.0 * sc[:, None] * (pitx[:, None] * cs + pity[:, None] * sn)
    R2 = 2.0 * sc[:, None] * (pixn[:, None] * cs + piyn[:, None] * sn)
    return dict(A=A, Cp=Cp, Qyy=Qyy, U1=ch, U2=tsh, B=B, D=D, Qpp=Qpp, R1=R1, R2=R2, K0=K0, K2=K2)


def factored(cells, sp, g, tab, merged=True):
    r = prepare(cells, tab, g)
    n = len(cells["tau"])
    ns, npT, nphi, ny = len(sp["mass"]), len(g["pT"]), len(g["phi"]), len(g["y"])
    out = np.zeros((ny, nphi, npT, ns))
    pref = (2.0 * np.pi * 0.197327053) ** -3
    pT = g["pT"]
    # merged tables (cell, slot) / (cell, phi) / (cell, slot, phi)
    K2 = r["K2"]
    QyyM = r["Qyy"] + K2[:, None] * r["A"] ** 2
    QppM = r["Qpp"] + K2[:, None] * r["B"] ** 2
    pair = r["R2"][:, None, :] * r["U2"][:, :, None] - r["R1"][:, None, :] * r["U1"][:, :, None]
    pairM = pair - 2.0 * K2[:, None, None] * r["A"][:, :, None] * r["B"][:, None, :]
    for s in range(ns):
        m = sp["mass"][s]; sign = sp["sign"][s]
        mT = np.sqrt(m * m + pT * pT)                       # [pT]
        a = mT[None, None, :] * r["A"][:, :, None]          # [cell, y, pT]
        q = pT[None, None, :] * r["B"][:, :, None]          # [cell, phi, pT]
        ea = np.exp(-a); eq = np.exp(q)
        x = a[:, :, None, :] - q[:, None, :, :]             # [cell, y, phi, pT]
        av = ea[:, :, None, :] * eq[:, None, :, :]
        av = np.where(x > 709.782712893384, 0.0, av)
        # where the factors under/overflow individually use the direct form (the kernel's general path)
        direct = ~np.isfinite(av) | (ea[:, :, None, :] < 1e-280)
        av = np.where(direct, np.where(x > 709.782712893384, 0.0, np.exp(-np.minimum(x, 800.0))), av)
        fb = 1.0 / (1.0 + sign * av)
        feq = av * fb
        if merged:
            dfs = (mT * mT)[None, None, None, :] * QyyM[:, :, None, None] + ((pT * pT)[None, None, None, :] * QppM[:, None, :, None] + (r["K0"] * m * m)[:, None, None, None]) \
                + (mT * pT)[None, None, None, :] * pairM[:, :, :, None]
        else:
            s0 = (mT * mT)[None, None, None, :] * r["Qyy"][:, :, None, None] + ((pT * pT)[None, None, None, :] * r["Qpp"][:, None, :, None] + (r["K0"] * m * m)[:, None, None, None]) \
                + (mT * pT)[None, None, None, :] * pair[:, :, :, None]
            dfs = s0 + K2[:, None, None, None] * x * x
        gg = np.clip(1.0 + fb * dfs, 0.0, 2.0)
        f = feq * gg
        pv = mT[None, None, None, :] * r["Cp"][:, :, None, None] + pT[None, None, None, :] * r["D"][:, None, :, None]
        term = np.where(pv > 0.0, pv * f, 0.0)
        out[:, :, :, s] = pref * sp["degeneracy"][s] * term.sum(axis=0)
    return out.ravel()


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    chosen = sys.argv[2] if len(sys.argv) > 2 else "chosen_pikp"
    fx = tables.load_fixture()
    cells = synthetic.columns_to_cells(synthetic.surface_vh(n, 1003), 1)
    sp = tables.species(fx, 1, chosen); g = tables.grid(fx); tab = tables.df_tables(fx, 1)
    fl = tables.flags(df_mode=1, dimension=3)
    cond = np.zeros(cfo.n_bins(sp, g))
    ref, _, _ = cfo.smooth(fl, cells, sp, g, tab, None, conditioning=cond)
    for merged in (False, True):
        got = factored(cells, sp, g, tab, merged)
        nz = ref != 0
        rel = np.abs(got[nz] - ref[nz]) / np.abs(ref[nz])
        w = np.argmax(rel)
        print("merged=%d: max rel %.3e (cond %.1e) median %.2e  >1e-10: %d of %d; zeros match %s" % (
            merged, rel.max(), cond[nz][w] / abs(ref[nz][w]), np.median(rel), (rel > 1e-10).sum(), nz.sum(), np.all(got[~nz] == 0)))
        allow = 1e-10 * np.abs(ref[nz]) + 64 * np.finfo(float).eps * cond[nz]
        print("          with the conditioning allowance: worst ratio %.3f" % (np.abs(got[nz] - ref[nz]) / allow).max())


if __name__ == "__main__":
    main()
