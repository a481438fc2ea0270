#!/usr/bin/env python
"""Generate tests/golden/yield_vectors.npz: the sampler's mean-yield estimate (calculate_total_yield) and the per-species densities
(compute_particle_densities) printed with 17 digits by the UNMODIFIED reference (oracle/_ref/is3d_ref yield) on small seeded
surfaces.

    python tests/golden/make_yield_vectors.py
"""
import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from is3d_b200 import synthetic, tables, workdir  # noqa: E402
from oracle import cf_oracle as cfo  # noqa: E402

CASES = [(3, 1, 300), (3, 2, 300), (3, 3, 300), (3, 4, 300), (2, 2, 50), (2, 4, 50)]      # (dimension, df_mode, cells)


def main():
    fx = tables.load_fixture()
    rec = {}
    for dimension, df_mode, n in CASES:
        cols = synthetic.surface_vh(n, 311 + df_mode, three_d=(dimension == 3), stress=(df_mode == 3))
        with tempfile.TemporaryDirectory() as wd:
            workdir.materialize(wd, surface_columns=cols, chosen="chosen_pikp", fixture=fx, operation=2, mode=1, hrg_eos=1,
                                dimension=dimension, df_mode=df_mode)
            _, info = cfo.run_reference(wd, what="yield")
        lines = info["stdout"].splitlines()
        N = [float(l.split()[1]) for l in lines if l.startswith("REF_YIELD")][0]
        dens = np.array([float(v) for l in lines if l.startswith("REF_DENSITIES") for v in l.split()[1:]]).reshape(-1, 2)
        key = "d%d_df%d" % (dimension, df_mode)
        rec[key + "/Ntot"] = np.array(N); rec[key + "/densities"] = dens
        rec[key + "/recipe"] = np.array(json.dumps(dict(dimension=dimension, df_mode=df_mode, n_cells=n, seed=311 + df_mode,
                                                        stress=(df_mode == 3), y_cut=5.0)))
        print(key, N, dens[:, 0])
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "yield_vectors.npz"), **rec)


if __name__ == "__main__":
    main()
