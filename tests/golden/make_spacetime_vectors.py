#!/usr/bin/env python
"""Generate tests/golden/spacetime_*.npz: operation = 0 (spacetime distributions) run by the UNMODIFIED reference
(oracle/_ref/is3d_ref full) on small seeded surfaces.

    python tests/golden/make_spacetime_vectors.py

The reference keeps these histograms in local arrays and only writes text with 7 significant digits
(emissionfunction_smooth_kernels.cpp:1404-1435), so each vector stores the parsed files per species (bin-volume normalisation
undone), the printed dN_dy values and -- for the first species -- the text of the four files, to pin the writers' format.
"""
import json
import os
import shutil
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from is3d_b200 import synthetic, tables, workdir  # noqa: E402
from oracle import cf_oracle as cfo  # noqa: E402

BINS = dict(tau_min=0.0, tau_max=12.0, tau_bins=12, r_min=0.0, r_max=12.0, r_bins=6)
CASES = {
    # name: (kwargs of synthetic.surface_vh, n_cells, seed, params)
    "dx3_df1": (dict(three_d=True), 80, 2003, dict(hrg_eos=1, dimension=3, df_mode=1)),
    "dx3_df2": (dict(three_d=True), 80, 2003, dict(hrg_eos=1, dimension=3, df_mode=2)),
    "dx3stress_df3": (dict(three_d=True, stress=True), 80, 2007, dict(hrg_eos=1, dimension=3, df_mode=3)),
    "dx3_df4": (dict(three_d=True), 80, 2003, dict(hrg_eos=1, dimension=3, df_mode=4)),
    "dx2_df1": (dict(three_d=False), 16, 2002, dict(hrg_eos=1, dimension=2, df_mode=1)),
    "dx2_df3": (dict(three_d=False), 16, 2002, dict(hrg_eos=1, dimension=2, df_mode=3)),
    "dx2stress_df4": (dict(three_d=False, stress=True), 16, 2008, dict(hrg_eos=1, dimension=2, df_mode=4)),
}
FILES = ("dN_taudtaudy_%d.dat", "dN_twopirdrdy_%d.dat", "dN_twopitaurdtaudrdy_%d.dat")


def main():
    fx = tables.load_fixture()
    only = sys.argv[1:]
    for name, (kw, n, seed, params) in CASES.items():
        if only and name not in only:
            continue
        wd = tempfile.mkdtemp(prefix="is3d_spacetime_")
        cols = synthetic.surface_vh(n, seed, **kw)
        workdir.materialize(wd, surface_columns=cols, chosen="chosen_pikp", fixture=fx, operation=0, mode=1, **params, **BINS)
        _, info = cfo.run_reference(wd, what="full")
        eta_pts = 241 if params["dimension"] == 2 else 1
        per = [cfo.read_spacetime_files(wd, int(m), BINS, eta_pts) for m in info["mcid"]]
        dndy = [float(l.split("=")[1]) for l in info["stdout"].splitlines() if l.startswith("dN_dy =")]
        d = os.path.join(wd, "results", "spacetime_distribution")
        m0 = int(info["mcid"][0])
        text = {f % m0: open(os.path.join(d, f % m0)).read() for f in FILES}
        rap = "dN_dydeta_%d_%dpt.dat" % (m0, eta_pts)
        text[rap] = open(os.path.join(d, rap)).read()
        rec = dict(mcid=np.array(info["mcid"], dtype=np.int64), dN_dy_printed=np.array(dndy),
                   recipe=np.array(json.dumps(dict(kwargs=kw, n_cells=n, seed=seed, params=params, bins=BINS))),
                   file_text=np.array(json.dumps(text)), eta_column=per[0]["eta_column"])
        for k in ("dN_tau", "dN_r", "dN_taur", "dN_dydeta"):
            rec[k] = np.stack([p[k] for p in per])
        out = os.path.join(ROOT, "tests", "golden", "spacetime_%s.npz" % name)
        np.savez_compressed(out, **rec)
        print("%-14s %4d cells  dN_dy %s -> %s (%d bytes)" % (name, n, dndy, os.path.basename(out), os.path.getsize(out)))
        shutil.rmtree(wd, ignore_errors=True)


if __name__ == "__main__":
    main()
