#!/usr/bin/env python
"""Generate tests/golden/golden_*.npz by running the UNMODIFIED reference (oracle/_ref/is3d_ref, built by
`make -C oracle ref` where /root/reference exists) on small seeded surfaces.

    python tests/golden/make_golden_vectors.py

Each vector stores the inputs' recipe (generator name, seed, size, parameters), the raw spectra array the reference
produced (`dN`, species fastest), the species order and -- for the `full` runs -- sha256 digests of the reference's own
results/*.dat text files, so that the writers can be checked byte for byte without committing megabytes of text.
All surfaces have u.dsigma > 0 and at most 10 000 cells (the reference's scratch array is never re-zeroed).
"""
import hashlib
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

HEAVY = [211, 2112, 12212, 42112, 2122, 13124, 3312, -2210, -2118, -3118, 10313, 337]   # all present in pdg-urqmd_v3.3+.dat

CASES = {
    # name: (generator, kwargs, n_cells, seed, chosen, params, what)
    "toy_df1": ("toy", {}, 1, 0, "chosen_pikp", dict(hrg_eos=2, dimension=3, df_mode=1), "full"),
    "s3_df1": ("vh", dict(three_d=True), 200, 1003, "chosen_pikp", dict(hrg_eos=1, dimension=3, df_mode=1), "kernel"),
    "s3_df2": ("vh", dict(three_d=True), 200, 1003, "chosen_pikp", dict(hrg_eos=1, dimension=3, df_mode=2), "kernel"),
    "s3_df3": ("vh", dict(three_d=True), 200, 1003, "chosen_pikp", dict(hrg_eos=1, dimension=3, df_mode=3), "kernel"),
    "s3_df4": ("vh", dict(three_d=True), 200, 1003, "chosen_pikp", dict(hrg_eos=1, dimension=3, df_mode=4), "kernel"),
    "s3stress_df3": ("vh", dict(three_d=True, stress=True), 100, 7, "chosen_pikp", dict(hrg_eos=1, dimension=3, df_mode=3), "kernel"),
    "s3stress_df2": ("vh", dict(three_d=True, stress=True), 100, 7, "chosen_pikp", dict(hrg_eos=1, dimension=3, df_mode=2), "kernel"),
    "s2_df1": ("vh", dict(three_d=False), 30, 1002, "chosen_pikp", dict(hrg_eos=1, dimension=2, df_mode=1), "full"),
    "s2_df3": ("vh", dict(three_d=False), 30, 1002, "chosen_pikp", dict(hrg_eos=1, dimension=2, df_mode=3), "kernel"),
    "s2_df2": ("vh", dict(three_d=False), 30, 1002, "chosen_pikp", dict(hrg_eos=1, dimension=2, df_mode=2), "kernel"),
    "s2_df4": ("vh", dict(three_d=False), 30, 1002, "chosen_pikp", dict(hrg_eos=1, dimension=2, df_mode=4), "kernel"),
    "s2stress_df3": ("vh", dict(three_d=False, stress=True), 24, 8, "chosen_pikp", dict(hrg_eos=1, dimension=2, df_mode=3), "kernel"),
    "s3stress_df4": ("vh", dict(three_d=True, stress=True), 100, 7, "chosen_pikp", dict(hrg_eos=2, dimension=3, df_mode=4), "kernel"),
    "s2_ideal": ("vh", dict(three_d=False, viscous=False), 30, 1002, "chosen_pikp",
                 dict(hrg_eos=1, dimension=2, df_mode=1, include_bulk_deltaf=0, include_shear_deltaf=0), "kernel"),
    "s3_heavy_df1": ("vh", dict(three_d=True), 60, 1003, HEAVY, dict(hrg_eos=1, dimension=3, df_mode=1), "kernel"),
    "vah_3d": ("vah", {}, 120, 1005, "chosen_pikp", dict(hrg_eos=1, dimension=3, df_mode=1, mode=2), "vah"),
    "vah_2d": ("vah", {}, 16, 1006, "chosen_pikp", dict(hrg_eos=1, dimension=2, df_mode=1, mode=2), "vah"),
    "s3_noreg_df1": ("vh", dict(three_d=True), 100, 1003, "chosen_pikp",
                     dict(hrg_eos=1, dimension=3, df_mode=1, regulate_deltaf=0, outflow=0), "kernel"),
}


def surface(case, fx):
    gen, kw, n, seed = CASES[case][:4]
    if gen == "toy":
        return fx["toy_surface"]
    return synthetic.surface_vh(n, seed, **kw)


def sha(path):
    return hashlib.sha256(open(path, "rb").read()).hexdigest()


def main():
    fx = tables.load_fixture()
    only = sys.argv[1:]
    for name, (gen, kw, n, seed, chosen, params, what) in CASES.items():
        if only and name not in only:
            continue
        wd = tempfile.mkdtemp(prefix="is3d_golden_")
        if gen == "vah":
            # the reference's anisotropic kernel is called directly; per-cell c0..c4 (no reader exists in src/cpp) come
            # from the test-suite's restatement of the only specification (src/cuda/deltafReader.cu:192-277)
            sys.path.insert(0, os.path.join(ROOT, "tests"))
            from common import vah_cells, vah_columns
            cols = vah_columns(n, seed, params["dimension"])
            cells = vah_cells(cols, fx)
            workdir.materialize(wd, surface_columns=cols, chosen=chosen, fixture=fx, operation=1, **params)
            np.concatenate([cells["c%d" % k] for k in range(5)]).tofile(os.path.join(wd, "input", "vah_coefficients.bin"))
        else:
            cols = surface(name, fx)
            workdir.materialize(wd, surface_columns=cols, chosen=chosen, fixture=fx, operation=1, mode=1, **params)
        dN, info = cfo.run_reference(wd, what=what)
        rec = dict(dN=dN, mcid=np.array(info["mcid"], dtype=np.int64), breakdown=np.array(info["breakdown"]),
                   recipe=np.array(json.dumps(dict(generator=gen, kwargs=kw, n_cells=n, seed=seed, params=params, what=what,
                                                   chosen=chosen if isinstance(chosen, str) else list(chosen)))))
        if what == "full":
            digests = {}
            for dirpath, _, files in os.walk(os.path.join(wd, "results")):
                for f in sorted(files):
                    if f.endswith(".dat"):
                        p = os.path.join(dirpath, f)
                        digests[os.path.relpath(p, wd)] = sha(p)
            rec["file_sha256"] = np.array(json.dumps(digests))
            rec["averages_file"] = np.array(open(os.path.join(wd, "average_thermodynamic_quantities.dat")).read())
        out = os.path.join(ROOT, "tests", "golden", "golden_%s.npz" % name)
        np.savez_compressed(out, **rec)
        print("%-14s %6d cells  %7d bins  nonzero %6d  breakdown %d  -> %s (%d bytes)"
              % (name, len(cols), dN.size, int((dN != 0).sum()), info["breakdown"], os.path.basename(out), os.path.getsize(out)))
        shutil.rmtree(wd, ignore_errors=True)


if __name__ == "__main__":
    main()
