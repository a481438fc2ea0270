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
for flkw in (dict(), dict(regulate_deltaf=0), dict(outflow=0), dict(include_bulk=0), dict(include_shear=0)):
    fl = tables.flags(df_mode=1, dimension=3, **flkw)
    tot = 0
    for i in (18, 60, 9):
        sub = {k: v[i:i + 1] for k, v in cells.items()}
        r, _, _ = cfo.smooth(fl, sub, sp, g, tab, None)
        d, _ = api.smooth_spectra(fl, sub, sp, g, tab, None)
        nz = r != 0
        rel = np.zeros_like(r); rel[nz] = (d[nz] - r[nz]) / np.abs(r[nz])
        bad = np.flatnonzero(np.abs(rel) > 1e-9)
        tot += len(bad)
        if not flkw:
            print("cell", i, "bad bins", len(bad), "of", nz.sum())
            for b in bad[:40]:
                print("   sp %2d pT %2d phi %2d (k %d) y %2d (j %d)  rel %+.3e ref %.3e" % (b % 45, (b // 45) % 32, (b // 45 // 32) % 24, (b // 45 // 32) % 24 % 3, b // 45 // 32 // 24, (b // 45 // 32 // 24) % 7, rel[b], r[b]))
    print(flkw, "bad bins total", tot)
