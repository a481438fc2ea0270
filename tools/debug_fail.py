import os, sys, tempfile, ctypes as C
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from common import compare, load_golden, problem_from_recipe, surface_columns
from is3d_b200 import api, synthetic, tables, workdir
from oracle import cf_oracle as cfo
fx = tables.load_fixture()
api.init()

def show(tag, got, ref, sp_n=3):
    rep = compare(got, ref)
    i = rep["worst_bin"]
    isp = i % sp_n; r = i // sp_n; ipT = r % 32; r //= 32; iphi = r % 24; iy = r // 24
    print(tag, {k: rep[k] for k in ("max_rel", "median_rel", "zeros_match", "ok")}, "worst: sp %d pT %d phi %d y %d ref %.6e got %.6e" % (isp, ipT, iphi, iy, ref[i], got[i]), flush=True)
    nz = ref != 0
    rel = np.abs(got[nz] - ref[nz]) / np.abs(ref[nz])
    print("   n(rel>1e-10) =", int((rel > 1e-10).sum()), "of", rel.size, " n(rel>1e-12) =", int((rel > 1e-12).sum()))

# (c) vah_2d
gold = load_golden("vah_2d")
fl, cells, sp, g, tab, gla = problem_from_recipe(gold["recipe"], fx)
for v in (0, 1, 3):
    dN, st = api.smooth_spectra(fl, cells, sp, g, tab, gla, tile_variant=v)
    show("vah_2d v%d" % v, dN, gold["dN"])
# (a) 2+1D ideal 300 cells
cells = synthetic.columns_to_cells(synthetic.surface_vh(300, synthetic.SEEDS["cfg2"], three_d=False, viscous=False), 1)
sp = tables.species(fx, 1, "chosen_pikp"); g = tables.grid(fx); tab = tables.df_tables(fx, 1)
fl = tables.flags(df_mode=1, dimension=2, include_bulk=0, include_shear=0)
ref, _, _ = cfo.smooth(fl, cells, sp, g, tab, None)
for v in (0, 3):
    for ch in (0, 1, 30):
        dN, st = api.smooth_spectra(fl, cells, sp, g, tab, None, tile_variant=v, n_chunks=ch)
        show("2d ideal v%d chunks %d" % (v, st["n_chunks"]), dN, ref)
# (b) s2_df1 via API then via run_workdir
gold = load_golden("s2_df1")
fl, cells, sp, g, tab, gla = problem_from_recipe(gold["recipe"], fx)
dN, st = api.smooth_spectra(fl, cells, sp, g, tab, gla)
show("s2_df1 api", dN, gold["dN"])
lib = api.lib()
with tempfile.TemporaryDirectory() as wd:
    workdir.materialize(wd, surface_columns=surface_columns(gold["recipe"], fx), chosen=gold["recipe"]["chosen"], fixture=fx, operation=1, mode=1, **gold["recipe"]["params"])
    dN = np.zeros(gold["dN"].size); st = api.Stats()
    rc = lib.is3d_b200_run_workdir(wd.encode(), dN.ctypes.data_as(C.POINTER(C.c_double)), C.c_int64(dN.size), None, 0, C.byref(st))
    show("s2_df1 run_workdir rc=%d" % rc, dN, gold["dN"])
# full species 64 cells
cells = synthetic.columns_to_cells(synthetic.surface_vh(20000, synthetic.SEEDS["cfg3"]), 1)
sub = {k: v[:64] for k, v in cells.items()}
sp = tables.species(fx, 1, "chosen_urqmd"); tab = tables.df_tables(fx, 1)
fl = tables.flags(df_mode=1, dimension=3)
ref, _, _ = cfo.smooth(fl, sub, sp, g, tab, None)
dN, _ = api.smooth_spectra(fl, sub, sp, g, tab, None)
show("full species 64 cells", dN, ref, sp_n=305)
rep = compare(dN, ref); i = rep["worst_bin"]
isp = i % 305; r = i // 305; ipT = r % 32; r //= 32; iphi = r % 24; iy = r // 24
print("species mass", sp["mass"][isp], "mcid", sp["mcid"][isp], "pT", g["pT"][ipT], "y", g["y"][iy])
# per-cell contributions for that bin from the oracle
best = []
for c in range(64):
    one = {k: v[c:c + 1] for k, v in sub.items()}
    sp1 = {k: v[isp:isp + 1] for k, v in sp.items()}
    r1, _, _ = cfo.smooth(fl, one, sp1, g, tab, None)
    g1, _ = api.smooth_spectra(fl, one, sp1, g, tab, None)
    j = ipT + 32 * (iphi + 24 * iy)
    best.append((r1[j], g1[j], c))
best.sort(reverse=True)
for rv, gv, c in best[:4]:
    print("  cell %d ref %.6e got %.6e rel %.3e  eta %.3f" % (c, rv, gv, abs(gv - rv) / rv if rv else 0, sub["eta"][c]))
