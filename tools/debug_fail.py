import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from common import compare, load_golden, problem_from_recipe
from is3d_b200 import api, tables
from oracle import cf_oracle as cfo
fx = tables.load_fixture()
api.init()
name = sys.argv[1] if len(sys.argv) > 1 else "s3stress_df4"
gold = load_golden(name)
fl, cells, sp, g, tab, gla = problem_from_recipe(gold["recipe"], fx)
dN, st = api.smooth_spectra(fl, cells, sp, g, tab, gla)
rep = compare(dN, gold["dN"]); print(rep, st["cells_feqmod_breakdown"])
ref = gold["dN"]; nz = ref != 0
rel = np.abs(dN[nz] - ref[nz]) / np.abs(ref[nz])
print("n > 1e-10:", int((rel > 1e-10).sum()), "of", rel.size, "max", rel.max())
n = len(cells["tau"])
bad = []
for c in range(n):
    one = {k: v[c:c + 1] for k, v in cells.items()}
    r1, _, _ = cfo.smooth(fl, one, sp, g, tab, gla)
    g1, _ = api.smooth_spectra(fl, one, sp, g, tab, gla)
    m = r1 != 0
    e = np.max(np.abs(g1[m] - r1[m]) / np.abs(r1[m])) if m.any() else 0.0
    z = bool(np.all(g1[~m] == 0))
    if e > 1e-10 or not z:
        bad.append((c, e, z))
print("bad cells:", bad[:20], len(bad))
for c, e, z in bad[:5]:
    print("cell", c, {k: float(cells[k][c]) for k in ("tau", "eta", "T", "P", "bulkPi", "pixx", "pixy", "piyy", "ux", "uy", "un")})
