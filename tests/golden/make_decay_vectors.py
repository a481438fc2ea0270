"""Golden vectors for the resonance-decay feed-down, from the REFERENCE's own routine
(EmissionFunctionArray::do_resonance_decays, emissionfunction_resonance_decays.cpp:124), which the reference snapshot disables with an
exit(-1) at entry and which oracle/_ref/is3d_ref_decays runs behind oracle/ref_decays_prefix.h without editing the source.
Label: the reference author flags the MTmax handling of the interpolation as unfinished.  Run here (needs /root/reference).

    python tests/golden/make_decay_vectors.py        -> tests/golden/decays_{2d,3d,3d_full}.npz
"""
import hashlib
import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from is3d_b200 import synthetic, tables, workdir      # noqa: E402
from oracle import cf_oracle as cfo                   # noqa: E402

CASES = {
    # light mesons / baryons + meson and (anti-)baryon resonances, parents later in the list than their daughters
    "2d": dict(dimension=2, n_cells=60, seed=77, strides=None,
               chosen=[211, -211, 111, 321, -321, 2212, -2212, 2112, -2112, 311, -311, 221, 3122, -3122, 113, 213, -213, 223, 323, -323, 313, 2224, -2224, 2214, 331, 333]),
    "3d": dict(dimension=3, n_cells=60, seed=78, strides=dict(pT=2, phi=3, y=4),
               chosen=[211, -211, 111, 321, -321, 2212, -2212, 2112, 311, -311, 221, 113, 213, -213, 223, 323, -323, 2224, -2224, 2214, 331, 333]),
    "3d_full": dict(dimension=3, n_cells=40, seed=79, strides=None, chosen=[211, -211, 111, 2212, -2212, 113, 223, 2224, -2224]),
}


def case_inputs(fx, rec):
    """(grid dict, table overrides, surface columns, thermal spectra from the oracle, particle list, chosen particle-list indices)"""
    tabs = None
    if rec["strides"]:
        s = rec["strides"]                                   # an int = stride, a list = explicit row indices
        pick = lambda t, v: t[::v] if isinstance(v, int) else t[list(v)]
        tabs = dict(pT=pick(fx["pT_tab"], s["pT"]), phi=pick(fx["phi_tab"], s["phi"]), y=pick(fx["y_tab"], s["y"]))
    g = tables.grid(fx, tabs)
    cols = synthetic.surface_vh(rec["n_cells"], rec["seed"], three_d=(rec["dimension"] == 3))
    cells = synthetic.columns_to_cells(cols, 1)
    sp = tables.species(fx, 1, rec["chosen"])
    fl = tables.flags(df_mode=1, dimension=rec["dimension"])
    dN, _, _ = cfo.smooth(fl, cells, sp, g, tables.df_tables(fx, 1), None)
    pdg = tables.pdg_decay_table(fx, 1)
    first = {}
    for n, m in enumerate(pdg["mcid"]):
        first.setdefault(int(m), n)
    return g, tabs, cols, dN, pdg, [first[m] for m in rec["chosen"]]


def main():
    fx = tables.load_fixture()
    for name, rec in CASES.items():
        g, tabs, cols, dN, pdg, chosen_idx = case_inputs(fx, rec)
        wd = tempfile.mkdtemp(prefix="decays_")
        workdir.materialize(wd, surface_columns=cols, chosen=rec["chosen"], fixture=fx, tables=tabs, operation=1, mode=1, hrg_eos=1,
                            dimension=rec["dimension"], df_mode=1, do_resonance_decays=1)
        out, info = cfo.run_reference_decays(wd, dN)
        sha = {}
        for f in ("dN_pTdpTdphidy_resonance_decays.dat", "dN_dpTdphidy_resonance_decays.dat"):
            sha["results/" + f] = hashlib.sha256(open(os.path.join(wd, "results", f), "rb").read()).hexdigest()
        ns = len(rec["chosen"]); y_pts = 1 if rec["dimension"] == 2 else len(g["y"])
        plane = ns * len(g["pT"]) * len(g["phi"]) * y_pts                       # 2+1D: only the y = 0 plane is used
        assert np.array_equal(out[plane:], dN[plane:])
        np.savez_compressed(os.path.join(ROOT, "tests", "golden", "decays_%s.npz" % name), recipe=json.dumps(rec), dN_in=dN[:plane], dN_out=out[:plane],
                            file_sha256=json.dumps(sha), mcid=np.array(info["mcid"]))
        print(name, "bins", plane, "reference seconds %.1f" % info["seconds"], "sum in %.6g out %.6g" % (dN.sum(), out.sum()), "nan", int(np.isnan(out).sum()))


if __name__ == "__main__":
    main()
