import sys, time, numpy as np, torch
sys.path.insert(0, "/root/repo")
from is3d_b200 import api, synthetic, tables
fx = tables.load_fixture()
g = tables.grid(fx); gla = tables.laguerre(fx)
for chosen, n in (("chosen_pikp", 100000), ("chosen_urqmd", 20000)):
    sp = tables.species(fx, 1, chosen); tab = tables.df_tables(fx, 1)
    cells = synthetic.columns_to_cells(synthetic.surface_vh(n, 1002, three_d=False), 1)
    keys = ("tau","eta","dat","dax","day","dan","ux","uy","un","T","P","E","pixx","pixy","pixn","piyy","piyn","bulkPi")
    dev = {k: torch.from_numpy(np.ascontiguousarray(cells[k])).cuda() for k in keys}
    for dfm in (1, 2):
        fl = tables.flags(df_mode=dfm, dimension=2)
        best = None
        for v in [0] + list(range(9, 17)):
            api.smooth_spectra(fl, dev, sp, g, tab, gla, memory="device", tile_variant=v)
            _, st = api.smooth_spectra(fl, dev, sp, g, tab, gla, memory="device", tile_variant=v)
            r = st["evaluations"] / (st["kernel_ms"] * 1e-3)
            print("2D viscous %s df%d variant %2d: %.3e evals/s" % (chosen, dfm, v, r), flush=True)
