"""Scratch GPU driver used during development: parity vs the oracle on small cases + a tile-variant timing sweep.
Usage (on a GPU box): python tools/gpu_check.py [parity] [sweep] [--cells N] [--species K]"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from is3d_b200 import api, synthetic, tables  # noqa: E402


def relerr(got, ref):
    nz = ref != 0
    rel = np.abs(got[nz] - ref[nz]) / np.abs(ref[nz])
    i = int(np.argmax(rel)) if rel.size else -1
    return dict(max=float(rel.max()) if rel.size else 0.0, median=float(np.median(rel)) if rel.size else 0.0,
                zeros_equal=bool(np.array_equal(ref == 0, got == 0)), n_zero_ref=int((~nz).sum()),
                worst_ref=float(ref[nz][i]) if rel.size else 0.0)


def parity():
    from oracle import cf_oracle as cfo
    fx = tables.load_fixture()
    g = tables.grid(fx); gla = tables.laguerre(fx)
    cases = [
        ("toy_df1", fx["toy_surface"], 3, 1, 2, "chosen_pikp", {}),
        ("s3_df1", synthetic.surface_vh(200, 1003), 3, 1, 1, "chosen_pikp", {}),
        ("s3_df2", synthetic.surface_vh(200, 1003), 3, 2, 1, "chosen_pikp", {}),
        ("s3_df1_noreg_noout", synthetic.surface_vh(200, 1003), 3, 1, 1, "chosen_pikp", dict(regulate_deltaf=0, outflow=0)),
        ("s3_df1_ideal", synthetic.surface_vh(200, 1003, viscous=False), 3, 1, 1, "chosen_pikp", dict(include_bulk=0, include_shear=0)),
        ("s3stress_df2", synthetic.surface_vh(200, 7, stress=True), 3, 2, 1, "chosen_pikp", {}),
        ("s2_df1", synthetic.surface_vh(30, 1002, three_d=False), 2, 1, 1, "chosen_pikp", {}),
        ("s2_df2", synthetic.surface_vh(30, 1002, three_d=False), 2, 2, 1, "chosen_pikp", {}),
        ("s3_df1_full", synthetic.surface_vh(40, 1003), 3, 1, 1, "chosen_urqmd", {}),
    ]
    ok = True
    for name, cols, dim, dfm, eos, chosen, extra in cases:
        cells = synthetic.columns_to_cells(cols, 1)
        sp = tables.species(fx, eos, chosen); tab = tables.df_tables(fx, eos)
        fl = tables.flags(df_mode=dfm, dimension=dim, **extra)
        ref, sk, bd = cfo.smooth(fl, cells, sp, g, tab, gla)
        for variant in (0, 2, 5):
            got, st = api.smooth_spectra(fl, cells, sp, g, tab, gla, tile_variant=variant + 1)
            e = relerr(got, ref)
            good = e["max"] <= 1e-10 and e["zeros_equal"] and st["cells_skipped_udsigma"] == sk
            ok &= good
            print("PARITY %-20s v%d %s max %.3e median %.2e zeros_equal %s (%d zero bins) worst_ref %.3e kernel_ms %.3f chunks %d"
                  % (name, variant, "ok  " if good else "FAIL", e["max"], e["median"], e["zeros_equal"], e["n_zero_ref"],
                     e["worst_ref"], st["kernel_ms"], st["n_chunks"]), flush=True)
    print("PARITY_ALL", "ok" if ok else "FAIL")
    return ok


def sweep(n_cells, n_species, variants, dims=(3,), models=("lin14", "lince")):
    import torch
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
    fx = tables.load_fixture()
    g = tables.grid(fx)
    sp_all = tables.species(fx, 1, "chosen_urqmd")
    sp = {k: v[:n_species] for k, v in sp_all.items()}
    tab = tables.df_tables(fx, 1); gla = tables.laguerre(fx)
    peak, ms = api.measure_fp64_peak()
    print("FP64_PEAK %.2f TFLOP/s (%.2f ms)" % (peak, ms), flush=True)
    W = {"lin14": 85, "lince": 87, "mike": 140, "jonah": 140, "vah": 94}
    out = []
    for dim in dims:
        nc = n_cells if dim == 3 else max(n_cells // 10, 100)
        for model in models:
            if model == "vah":
                from common import vah_cells, vah_columns
                cells = vah_cells(vah_columns(nc, 1005, dim), fx)
                fl = tables.flags(df_mode=1, dimension=dim); fl["mode"] = 2
                t2 = None
            else:
                cols = synthetic.surface_vh(nc, 1003, three_d=(dim == 3))
                cells = synthetic.columns_to_cells(cols, 1)
                dfm = {"lin14": 1, "lince": 2, "mike": 3, "jonah": 4}[model]
                fl = tables.flags(df_mode=dfm, dimension=dim)
                t2 = dict(tab)
                if dfm == 4:
                    pdg = tables.pdg_table(fx, 1); avg = api.surface_averages(cells)
                    t2.update(api.jonah_tables(pdg["mass"], pdg["gspin"].astype(float), pdg["sign"].astype(float), avg[0], gla))
            dev = {k: torch.tensor(v, device="cuda") for k, v in cells.items() if v is not None}
            for v in variants:
                best = None
                for rep in range(2):
                    res, st = api.smooth_spectra(fl, dev, sp, g, t2, gla, memory="device", tile_variant=v + 1)
                    torch.cuda.synchronize()
                    if best is None or st["kernel_ms"] < best["kernel_ms"]:
                        best = st
                ev = best["evaluations"] / (best["kernel_ms"] * 1e-3)
                rec = dict(dim=dim, model=model, variant=v, n_cells=nc, n_species=len(sp["mass"]), kernel_ms=best["kernel_ms"],
                           prepare_ms=best["prepare_ms"], evals_per_s=ev, chunks=best["n_chunks"], frac=ev * W[model] / (peak * 1e12),
                           checksum=float(res.sum().item()))
                out.append(rec)
                print("SWEEP", json.dumps(rec), flush=True)
    return out


if __name__ == "__main__":
    args = sys.argv[1:]
    def opt(name, default):
        return int(args[args.index(name) + 1]) if name in args else default
    t0 = time.time()
    api.init()
    if "parity" in args or not args:
        parity()
    if "sweep" in args:
        vs = list(range(8)) if "--variants" not in args else [int(v) for v in args[args.index("--variants") + 1].split(",")]
        dims = (3, 2) if "--dim2" in args else ((2,) if "--only2d" in args else (3,))
        models = args[args.index("--models") + 1].split(",") if "--models" in args else ["lin14", "lince"]
        sweep(opt("--cells", 20000), opt("--species", 305), vs, dims, models)
    print("done in %.1fs" % (time.time() - t0))
