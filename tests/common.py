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


def compare(got, ref, tol=REL_TOL):
    """Per-bin parity: |got - ref| <= tol |ref| where ref != 0, got == 0 where ref == 0.  Returns a report dict."""
    got = np.asarray(got); ref = np.asarray(ref)
    assert got.shape == ref.shape
    nz = ref != 0
    rel = np.abs(got[nz] - ref[nz]) / np.abs(ref[nz])
    worst = int(np.argmax(rel)) if rel.size else -1
    idx = np.flatnonzero(nz)[worst] if rel.size else -1
    return dict(max_rel=float(rel.max()) if rel.size else 0.0, median_rel=float(np.median(rel)) if rel.size else 0.0,
                zeros_match=bool(np.all(got[~nz] == 0)), worst_bin=int(idx), worst_ref=float(ref[idx]) if rel.size else 0.0,
                ok=bool((rel.size == 0 or rel.max() <= tol) and np.all(got[~nz] == 0)))
