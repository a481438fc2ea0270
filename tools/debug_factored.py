"""Development aid (GPU): locate which cells make cf_factored_kernel differ from the oracle on the 45-species test problem."""
import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from is3d_b200 import api, synthetic, tables
from oracle import cf_oracle as cfo

api.init()
fx = tables.load_fixture()
ids = fx["chosen_urqmd"]; n = 45
sp = tables.species(fx, 1, list(ids[:: max(1, len(ids) // n)][:n])); g = tables.grid(fx); tab = tables.df_tables(fx, 1)
cells = synthetic.columns_to_cells(synthetic.surface_vh(150, 4242), 1)
fl = tables.flags(df_mode=1, dimension=3)
ref, _, _ = cfo.smooth(fl, cells, sp, g, tab, None)
for variant in (0, 10, 18, 19):
    dN, st = api.smooth_spectra(fl, cells, sp, g, tab, None, tile_variant=variant)
    nz = ref != 0
    rel = np.abs(dN[nz] - ref[nz]) / np.abs(ref[nz])
    bad = np.flatnonzero(nz)[rel > 1e-10]
    print("variant", variant, "max rel %.3e" % rel.max(), "bins > 1e-10:", len(bad), "zeros", np.all(dN[~nz] == 0), "chunks", st["n_chunks"])
    if variant == 0:
        badbins = bad
print("bad bins (ipart, ipT, iphi, iy):", [(b % 45, (b // 45) % 32, (b // 45 // 32) % 24, b // 45 // 32 // 24) for b in badbins[:20]])
b = badbins[np.argmax((np.abs(api.smooth_spectra(fl, cells, sp, g, tab, None)[0][badbins] - ref[badbins]) / ref[badbins]))] if len(badbins) else 347721
print("bin", b)
# per-cell
for i in range(150):
    sub = {k: v[i:i + 1] for k, v in cells.items()}
    r, _, _ = cfo.smooth(fl, sub, sp, g, tab, None)
    d, _ = api.smooth_spectra(fl, sub, sp, g, tab, None)
    o, _ = api.smooth_spectra(fl, sub, sp, g, tab, None, tile_variant=10)
    if r[b] != 0 and abs(d[b] - r[b]) > 1e-11 * abs(r[b]):
        print("cell", i, "ref %.17g fact %.17g old %.17g  rel %.3e  share %.3e" % (r[b], d[b], o[b], abs(d[b] - r[b]) / abs(r[b]), r[b] / ref[b]),
              "eta %.3f T %.4f" % (cells["eta"][i], cells["T"][i]))
    # whole-array check for this cell
    nzc = r != 0
    relc = np.abs(d[nzc] - r[nzc]) / np.abs(r[nzc])
    if relc.size and relc.max() > 1e-9:
        w = np.flatnonzero(nzc)[np.argmax(relc)]
        print("   cell", i, "worst bin", (w % 45, (w // 45) % 32, (w // 45 // 32) % 24, w // 45 // 32 // 24), "rel %.3e ref %.3e" % (relc.max(), r[w]))
