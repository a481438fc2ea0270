"""Shared helpers for the test-suite: rebuild a golden case's inputs from its recipe, error metrics."""
import glob
import json
import os

import numpy as np

from is3d_b200 import synthetic, tables

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
REL_TOL = 1e-10     # north_star: 1e-10 relative per momentum bin


def golden_names():
    return sorted(os.path.basename(p)[len("golden_"):-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "golden_*.npz")))


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, "golden_%s.npz" % name))
    rec = {k: z[k] for k in z.files}
    rec["recipe"] = json.loads(str(rec["recipe"]))
    if "file_sha256" in rec:
        rec["file_sha256"] = json.loads(str(rec["file_sha256"]))
    return rec


def surface_columns(recipe, fx):
    if recipe["generator"] == "toy":
        return fx["toy_surface"]
    return synthetic.surface_vh(recipe["n_cells"], recipe["seed"], **recipe["kwargs"])


def problem_from_recipe(recipe, fx, jonah_with_oracle=True):
    """Inputs for the oracle / the C ABI: (flags, cells, species, grid, df tables, laguerre)."""
    p = recipe["params"]
    if recipe["generator"] == "vah":
        fl, cells, sp, g, _ = vah_problem(recipe, fx)
        return fl, cells, sp, g, None, None
    cols = surface_columns(recipe, fx)
    cells = synthetic.columns_to_cells(cols, 1)
    eos = p.get("hrg_eos", 1)
    chosen = recipe["chosen"]
    sp = tables.species(fx, eos, chosen)
    g = tables.grid(fx)
    tab = tables.df_tables(fx, eos)
    gla = tables.laguerre(fx)
    fl = tables.flags(df_mode=p["df_mode"], dimension=p["dimension"],
                      include_bulk=p.get("include_bulk_deltaf", 1), include_shear=p.get("include_shear_deltaf", 1),
                      regulate_deltaf=p.get("regulate_deltaf", 1), outflow=p.get("outflow", 1), deta_min=p.get("deta_min", 1e-5))
    if p["df_mode"] == 4 and jonah_with_oracle:
        tab.update(jonah_tables(cells, fx, eos, gla))
    return fl, cells, sp, g, tab, gla


def jonah_tables(cells, fx, eos, gla):
    """lambda/z tables at the surface-average temperature, through the 15-digit text round trip of the side file."""
    from oracle import cf_oracle as cfo
    avg = [float("%.15g" % v) for v in cfo.surface_averages(cells)]
    pdg = tables.pdg_table(fx, eos)
    return cfo.jonah_tables(pdg["mass"], pdg["gspin"].astype(float), pdg["sign"].astype(float), avg[0], gla)


def vah_cells(cols, fx):
    """Mode-2 columns -> SoA dict incl. alpha_L, Lambda (readindata.cpp:905-918) and per-cell c0..c4 from the vah tables
    (bilinear in (Lambda [fm^-1], alpha_L), / hbarC^3; only specification: reference src/cuda/deltafReader.cu:192-277)."""
    from oracle import cf_oracle as cfo
    hb = synthetic.HBARC
    cells = synthetic.columns_to_cells(cols, 2)
    a = np.asarray(cols)
    Tf, Pf, PLf = a[:, 13], a[:, 14], a[:, 15]
    lib = cfo.lib()
    aL = np.array([lib.cfo_aL_fit(float(x)) for x in PLf / Pf])
    Lam = np.array([t / (0.5 * al * lib.cfo_R200(float(al))) ** 0.25 for t, al in zip(Tf, aL)])
    cells["aL"] = aL; cells["Lambda"] = Lam * hb
    nL, naL = int(fx["df_vah/nL"]), int(fx["df_vah/naL"])
    Lg = fx["df_vah/L_col"][:nL]; ag = fx["df_vah/aL_col"][::nL]
    i1 = np.searchsorted(Lg, Lam, side="right"); i2 = np.searchsorted(ag, aL, side="right")
    assert np.all((i1 >= 1) & (i1 < nL) & (i2 >= 1) & (i2 < naL)), "cell outside the vah table"
    L1, L2, a1, a2 = Lg[i1 - 1], Lg[i1], ag[i2 - 1], ag[i2]
    for k in range(5):
        t = fx["df_vah/c%d" % k].reshape(naL, nL)          # file order: alpha_L outer, Lambda inner
        f11, f21, f12, f22 = t[i2 - 1, i1 - 1], t[i2 - 1, i1], t[i2, i1 - 1], t[i2, i1]
        v = ((f11 * (L2 - Lam) + f21 * (Lam - L1)) * (a2 - aL) + (f12 * (L2 - Lam) + f22 * (Lam - L1)) * (aL - a1)) / ((a2 - a1) * (L2 - L1))
        cells["c%d" % k] = v / (hb * hb * hb)
    return cells


def vah_columns(n_cells, seed, dimension):
    cols = synthetic.surface_vah(n_cells, seed)
    if dimension == 2:
        cols[:, 3] = 0.0; cols[:, 7] = 0.0; cols[:, 11] = 0.0       # eta, dsigma_eta, u^eta
    return cols


def vah_problem(recipe, fx):
    p = recipe["params"]
    cols = vah_columns(recipe["n_cells"], recipe["seed"], p["dimension"])
    cells = vah_cells(cols, fx)
    sp = tables.species(fx, p.get("hrg_eos", 1), recipe["chosen"])
    g = tables.grid(fx)
    fl = tables.flags(df_mode=p.get("df_mode", 1), dimension=p["dimension"], include_bulk=p.get("include_bulk_deltaf", 1),
                      include_shear=p.get("include_shear_deltaf", 1), regulate_deltaf=p.get("regulate_deltaf", 1), outflow=p.get("outflow", 1))
    fl["mode"] = 2
    return fl, cells, sp, g, cols


def compare(got, ref, tol=REL_TOL, conditioning=None):
    """Per-bin parity: |got - ref| <= tol |ref| where ref != 0, got == 0 where ref == 0.  Returns a report dict.

    conditioning (optional): the oracle's per-bin magnitude sum dN_abs (every delta-f term taken in absolute value).  A bin whose
    value is the nearly cancelling difference 1 + df ~ 0 of O(1) terms carries an absolute rounding noise of a few ulp of
    dN_abs in the *reference itself*; for those bins the allowance is tol |ref| + 64 eps dN_abs.  For well-conditioned bins
    (dN_abs ~ |ref|) the second term is 1.4e-14 |ref| and the bar stays tol = 1e-10."""
    got = np.asarray(got); ref = np.asarray(ref)
    assert got.shape == ref.shape
    nz = ref != 0
    rel = np.abs(got[nz] - ref[nz]) / np.abs(ref[nz])
    if conditioning is not None:
        allow = tol * np.abs(ref[nz]) + 64 * np.finfo(float).eps * np.asarray(conditioning)[nz]
        rel = rel * (tol * np.abs(ref[nz]) / allow)          # rescaled so that the same `<= tol` test applies
    worst = int(np.argmax(rel)) if rel.size else -1
    idx = np.flatnonzero(nz)[worst] if rel.size else -1
    return dict(max_rel=float(rel.max()) if rel.size else 0.0, median_rel=float(np.median(rel)) if rel.size else 0.0,
                zeros_match=bool(np.all(got[~nz] == 0)), worst_bin=int(idx), worst_ref=float(ref[idx]) if rel.size else 0.0,
                ok=bool((rel.size == 0 or rel.max() <= tol) and np.all(got[~nz] == 0)))


# ------------------------------------------------------------------------------------------ operation = 0 golden vectors
def spacetime_names():
    return sorted(os.path.basename(p)[len("spacetime_"):-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "spacetime_*.npz")))


def load_spacetime(name):
    z = np.load(os.path.join(GOLDEN_DIR, "spacetime_%s.npz" % name))
    out = {k: z[k] for k in z.files}
    out["recipe"] = json.loads(str(z["recipe"]))
    out["file_text"] = json.loads(str(z["file_text"]))
    return out


def spacetime_problem(recipe, fx):
    """(flags, cells incl. x, y, species, grid, df tables, laguerre, bins, surface columns) of a spacetime golden vector"""
    from is3d_b200 import synthetic, tables
    cols = synthetic.surface_vh(recipe["n_cells"], recipe["seed"], **recipe["kwargs"])
    cells = synthetic.columns_to_cells(cols, 1)
    prm = recipe["params"]
    sp = tables.species(fx, prm["hrg_eos"], "chosen_pikp"); g = tables.grid(fx); tab = tables.df_tables(fx, prm["hrg_eos"])
    gla = tables.laguerre(fx)
    if prm["df_mode"] == 4:
        tab.update(jonah_tables(cells, fx, prm["hrg_eos"], gla))
    fl = tables.flags(df_mode=prm["df_mode"], dimension=prm["dimension"])
    return fl, cells, sp, g, tab, gla, recipe["bins"], cols


def spacetime_close(got, gold, rel=2e-6):
    """7-significant-digit text precision of the reference files: compare against each histogram's largest entry"""
    for k in ("dN_tau", "dN_r", "dN_taur", "dN_dydeta"):
        a = np.asarray(got[k], dtype=np.float64).reshape(gold[k].shape)
        scale = np.abs(gold[k]).max()
        assert np.abs(a - gold[k]).max() <= rel * scale, k
    assert np.abs(np.asarray(got["dN_dy"]) - gold["dN_dy_printed"]).max() < 1e-6
