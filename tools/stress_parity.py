"""Development tool: GPU vs oracle on deliberately nasty surfaces (all df modes)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from common import compare, jonah_tables
from is3d_b200 import api, synthetic, tables
from oracle import cf_oracle as cfo
fx = tables.load_fixture(); api.init()
g = tables.grid(fx); gla = tables.laguerre(fx)
sp = tables.species(fx, 1, [211, 321, 2212, -3334, 337])

def variants():
    base = synthetic.surface_vh(48, 77)
    out = {"base": base}
    s = base.copy(); s[:, 10] = np.sinh(np.linspace(-2.0, 2.0, len(s))) / s[:, 0]; out["fast_longitudinal_flow"] = s
    s = base.copy(); s[:, 0] = np.linspace(0.05, 0.5, len(s)); s[:, 10] = 0.1 / s[:, 0]; out["small_tau"] = s
    s = base.copy(); s[:, 12] = np.linspace(0.1001, 0.1999, len(s)) / synthetic.HBARC; out["T_table_edges"] = s
    s = base.copy(); s[:, 8] = 0.0; s[:, 9] = 0.0; out["no_transverse_flow"] = s
    s = base.copy(); s[:, 19] = np.linspace(-1.5, 1.5, len(s)) * 0.05 / synthetic.HBARC; out["huge_bulk"] = s
    s = synthetic.surface_vh(48, 78, stress=True); out["stress"] = s
    s = base.copy(); s[:, 3] = np.linspace(-6, 6, len(s)); out["far_eta"] = s
    # fix u.dsigma > 0 by making dsigma_tau large
    for k, v in out.items():
        v[:, 4] = np.abs(v[:, 4]) * 5 + 0.5
    return out

for name, cols in variants().items():
    cells = synthetic.columns_to_cells(cols, 1)
    for dfm in (1, 2, 3, 4):
        for extra in ({}, dict(regulate_deltaf=0, outflow=0)):
            tab = tables.df_tables(fx, 1)
            fl = tables.flags(df_mode=dfm, dimension=3, **extra)
            try:
                if dfm == 4:
                    tab.update(jonah_tables(cells, fx, 1, gla))
                cond = np.zeros(len(sp["mass"]) * 32 * 24 * 21) if dfm in (1, 2) else None
                ref, sk, bd = cfo.smooth(fl, cells, sp, g, tab, gla, conditioning=cond)
            except RuntimeError as e:
                print("%-24s df%d %-8s oracle error %s" % (name, dfm, "noreg" if extra else "", e)); continue
            try:
                got, st = api.smooth_spectra(fl, cells, sp, g, tab, gla)
            except api.Is3dError as e:
                print("%-24s df%d %-8s gpu error %s" % (name, dfm, "noreg" if extra else "", e)); continue
            plain = compare(got, ref); c2 = compare(got, ref, conditioning=cond) if cond is not None else plain
            fin = np.isfinite(ref).all()
            print("%-24s df%d %-6s max %.2e (cond-aware %.2e) zeros %s bd %d/%d skipped %d/%d finite %s %s" % (
                name, dfm, "noreg" if extra else "", plain["max_rel"], c2["max_rel"], plain["zeros_match"], st["cells_feqmod_breakdown"], bd,
                st["cells_skipped_udsigma"], sk, fin, "" if c2["ok"] else "<<<<<<"), flush=True)
