"""operation = 0 (spacetime distributions): CUDA path through the C ABI vs the oracle.

Tolerance: 1e-10 relative on every histogram bin that holds a non-negligible share of the species' yield (the bins are sums of
non-negative cell yields, so there is no cancellation across cells); bins below 1e-12 of the largest one are compared
absolutely against that scale.  Empty bins must be exactly zero."""
import ctypes as C
import os
import re
import tempfile

import numpy as np
import pytest

from common import jonah_tables, load_spacetime, spacetime_close, spacetime_names, spacetime_problem
from is3d_b200 import api, synthetic, tables, workdir
from oracle import cf_oracle as cfo

pytestmark = pytest.mark.gpu

BINS = dict(tau_min=0.0, tau_max=12.0, tau_bins=24, r_min=0.0, r_max=12.0, r_bins=12)


def check(got, ref, tol=1e-10):
    for k in ("dN_tau", "dN_r", "dN_taur", "dN_dydeta", "dN_dy"):
        a = np.asarray(got[k], dtype=np.float64).reshape(ref[k].shape); b = ref[k]
        assert np.array_equal(a == 0.0, b == 0.0), k
        scale = np.abs(b).max() if b.size else 0.0
        err = np.abs(a - b)
        bad = err > tol * np.abs(b) + 1e-12 * tol * scale
        assert not bad.any(), (k, float((err / np.maximum(np.abs(b), 1e-300)).max()))


def problem(fx, n_cells, dimension, df_mode, seed, chosen="chosen_pikp", stress=False, **flag_kw):
    cols = synthetic.surface_vh(n_cells, seed, three_d=(dimension == 3), stress=stress)
    cells = synthetic.columns_to_cells(cols, 1)
    sp = tables.species(fx, 1, chosen); g = tables.grid(fx); tab = tables.df_tables(fx, 1); gla = tables.laguerre(fx)
    if df_mode == 4:
        tab.update(jonah_tables(cells, fx, 1, gla))
    fl = tables.flags(df_mode=df_mode, dimension=dimension, **flag_kw)
    return fl, cells, sp, g, tab, gla


@pytest.mark.parametrize("dimension,df_mode,n_cells", [(3, 1, 700), (3, 2, 700), (2, 1, 40), (2, 2, 40)])
def test_spacetime_linear_df(fx, dimension, df_mode, n_cells):
    fl, cells, sp, g, tab, gla = problem(fx, n_cells, dimension, df_mode, 9100 + df_mode)
    ref, skipped = cfo.spacetime(fl, cells, sp, g, tab, BINS)
    got, st = api.spacetime_distributions(fl, cells, sp, g, tab, gla, BINS)
    assert st["cells_skipped_udsigma"] == skipped
    check(got, ref)


@pytest.mark.parametrize("dimension,df_mode,n_cells,stress", [(3, 3, 500, False), (3, 4, 500, False), (3, 3, 500, True), (3, 4, 500, True),
                                                              (2, 3, 40, False), (2, 4, 40, False), (2, 3, 40, True), (2, 4, 40, True)])
def test_spacetime_feqmod(fx, dimension, df_mode, n_cells, stress):
    """calculate_dN_dX_feqmod: incl. cells that break down (stress, df_mode 3) and 2+1D cells with detA > 1 (eta rescaled)"""
    fl, cells, sp, g, tab, gla = problem(fx, n_cells, dimension, df_mode, 9600 + df_mode, stress=stress)
    ref, skipped = cfo.spacetime(fl, cells, sp, g, tab, BINS, gla)
    got, st = api.spacetime_distributions(fl, cells, sp, g, tab, gla, BINS)
    assert st["cells_skipped_udsigma"] == skipped
    assert st["cells_feqmod_breakdown"] == ref["breakdown"]
    if stress and df_mode == 3:
        assert ref["breakdown"] > 0
    check(got, ref)


def test_spacetime_full_species_list_and_fine_bins(fx):
    fl, cells, sp, g, tab, gla = problem(fx, 300, 3, 1, 9200, chosen="chosen_urqmd")
    bins = dict(tau_min=0.5, tau_max=11.0, tau_bins=120, r_min=1.0, r_max=13.0, r_bins=60)     # cells below and above both ranges
    ref, _ = cfo.spacetime(fl, cells, sp, g, tab, bins)
    got, _ = api.spacetime_distributions(fl, cells, sp, g, tab, gla, bins)
    check(got, ref)


def test_spacetime_unregulated_no_outflow(fx):
    fl, cells, sp, g, tab, gla = problem(fx, 300, 3, 2, 9300, regulate_deltaf=0, outflow=0)
    ref, _ = cfo.spacetime(fl, cells, sp, g, tab, BINS)
    got, _ = api.spacetime_distributions(fl, cells, sp, g, tab, gla, BINS)
    for k in ("dN_tau", "dN_r", "dN_dy"):              # signed terms: compare against the yield scale
        scale = np.abs(ref[k]).max()
        assert np.abs(got[k] - ref[k]).max() <= 1e-10 * scale, k


def test_spacetime_skipped_cells_and_chunk_split(fx):
    fl, cells, sp, g, tab, gla = problem(fx, 500, 3, 1, 9400)
    cells["dat"][::7] *= -1.0; cells["dax"][::7] *= -1.0; cells["day"][::7] *= -1.0; cells["dan"][::7] *= -1.0    # u.dsigma < 0
    coarse = dict(tau_min=0.0, tau_max=12.0, tau_bins=2, r_min=0.0, r_max=15.0, r_bins=1)   # big categories -> split into several chunks
    ref, skipped = cfo.spacetime(fl, cells, sp, g, tab, coarse)
    assert skipped > 0
    got, st = api.spacetime_distributions(fl, cells, sp, g, tab, gla, coarse, n_chunks=7)
    assert st["cells_skipped_udsigma"] == skipped
    check(got, ref)


def test_spacetime_device_memory_and_empty_surface(fx):
    import torch
    fl, cells, sp, g, tab, gla = problem(fx, 200, 3, 1, 9500)
    ref, _ = cfo.spacetime(fl, cells, sp, g, tab, BINS)
    dev = {k: torch.from_numpy(np.ascontiguousarray(v)).cuda() for k, v in cells.items()}
    got, _ = api.spacetime_distributions(fl, dev, sp, g, tab, gla, BINS, memory="device")
    check(got, ref)
    empty = {k: v[:0] for k, v in cells.items()}
    got, _ = api.spacetime_distributions(fl, empty, sp, g, tab, gla, BINS)
    assert all(not np.any(v) for v in got.values())


# ---------------------------------------------------------------------------------------- reference vectors, file interface
@pytest.mark.parametrize("name", spacetime_names())
def test_spacetime_golden_vectors(name, fx):
    """CUDA path vs what the unmodified reference wrote (7 significant digits) and vs the oracle (1e-10)"""
    gold = load_spacetime(name)
    fl, cells, sp, g, tab, gla, bins, _ = spacetime_problem(gold["recipe"], fx)
    got, st = api.spacetime_distributions(fl, cells, sp, g, tab, gla, bins)
    spacetime_close(got, gold)
    ref, _ = cfo.spacetime(fl, cells, sp, g, tab, bins, gla)
    check(got, ref)


@pytest.mark.parametrize("name", ["dx3_df1", "dx2_df3", "dx3stress_df3"])
def test_run_workdir_operation_0(name, fx):
    """is3d_b200_run_workdir with operation = 0 writes results/spacetime_distribution/ like calculate_dN_dX{,_feqmod}"""
    gold = load_spacetime(name)
    rec = gold["recipe"]; bins = rec["bins"]
    fl, cells, sp, g, tab, gla, _, cols = spacetime_problem(rec, fx)
    nt, nr = bins["tau_bins"], bins["r_bins"]; ns = len(sp["mass"]); eta_pts = 241 if rec["params"]["dimension"] == 2 else 1
    n_raw = ns * (nt + nr + nt * nr + eta_pts + 1)
    with tempfile.TemporaryDirectory() as wd:
        workdir.materialize(wd, surface_columns=cols, chosen="chosen_pikp", fixture=fx, operation=0, mode=1, **rec["params"], **bins)
        raw = np.zeros(n_raw); mcid = np.zeros(8, dtype=np.int32); st = api.Stats()
        rc = api.lib().is3d_b200_run_workdir(wd.encode(), raw.ctypes.data_as(C.POINTER(C.c_double)), C.c_int64(raw.size),
                                             mcid.ctypes.data_as(C.POINTER(C.c_int32)), 8, C.byref(st))
        assert rc == 0
        assert list(mcid[:ns]) == list(gold["mcid"])
        o = np.cumsum([0, ns * nt, ns * nr, ns * nt * nr, ns * eta_pts, ns])
        got = dict(dN_tau=raw[o[0]:o[1]].reshape(ns, nt), dN_r=raw[o[1]:o[2]].reshape(ns, nr), dN_taur=raw[o[2]:o[3]].reshape(ns, nt, nr),
                   dN_dydeta=raw[o[3]:o[4]].reshape(ns, eta_pts), dN_dy=raw[o[4]:o[5]])
        spacetime_close(got, gold)
        # the files: same names, line counts and number format as the reference's; values to the text precision
        d = os.path.join(wd, "results", "spacetime_distribution")
        number = r"-?\d\.\d{6}e[+-]\d{2}"
        n_same = n_lines = 0
        for fname, ref_text in gold["file_text"].items():
            text = open(os.path.join(d, fname)).read()
            a, b = text.splitlines(), ref_text.splitlines()
            assert len(a) == len(b), fname
            for la, lb in zip(a, b):
                assert re.fullmatch(number + "(\t" + number + ")+", la), (fname, la)
                va, vb = np.array(la.split("\t"), dtype=float), np.array(lb.split("\t"), dtype=float)
                assert np.all(np.abs(va - vb) <= 2e-6 * np.abs(vb) + 1e-30), (fname, la, lb)
                n_same += (la == lb); n_lines += 1
        assert n_same >= 0.97 * n_lines          # last-digit flips only where a sum sits on a rounding boundary
        for m in gold["mcid"][1:]:
            assert os.path.getsize(os.path.join(d, "dN_taudtaudy_%d.dat" % m)) > 0
        files = cfo.read_spacetime_files(wd, int(gold["mcid"][1]), bins, eta_pts)
        assert np.abs(files["dN_tau"] - gold["dN_tau"][1]).max() <= 4e-6 * np.abs(gold["dN_tau"][1]).max()


def test_in_memory_surface_operation_0(fx):
    """is3d_b200_run_surface with operation = 0: the cells (incl. x, y) come from memory"""
    gold = load_spacetime("dx3_df4")
    rec = gold["recipe"]; bins = rec["bins"]
    fl, cells, sp, g, tab, gla, _, cols = spacetime_problem(rec, fx)
    nt, nr = bins["tau_bins"], bins["r_bins"]; ns = len(sp["mass"])
    with tempfile.TemporaryDirectory() as wd:
        workdir.materialize(wd, surface_columns=cols, chosen="chosen_pikp", fixture=fx, operation=0, mode=1, **rec["params"], **bins)
        os.remove(os.path.join(wd, "input", "surface.dat"))
        m = api._Marshal(False)
        sf = api.Surface(); sf.n_cells = len(cells["tau"])
        for k in api.SURFACE_FIELDS:
            if k in cells:
                setattr(sf, k, m.cells(cells[k]))
        raw = np.zeros(ns * (nt + nr + nt * nr + 1 + 1)); st = api.Stats()
        rc = api.lib().is3d_b200_run_surface(wd.encode(), C.byref(sf), raw.ctypes.data_as(C.POINTER(C.c_double)), C.c_int64(raw.size), None, 0, C.byref(st))
        assert rc == 0
        assert np.abs(raw[:ns * nt].reshape(ns, nt) - gold["dN_tau"]).max() <= 2e-6 * np.abs(gold["dN_tau"]).max()
        assert os.path.getsize(os.path.join(wd, "results", "spacetime_distribution", "dN_twopirdrdy_211.dat")) > 0


def test_spacetime_rejects_what_the_reference_cannot_do(fx):
    fl, cells, sp, g, tab, gla = problem(fx, 10, 3, 1, 9700)
    no_xy = {k: v for k, v in cells.items() if k not in ("x", "y")}
    with pytest.raises(KeyError):
        api.spacetime_distributions(fl, no_xy, sp, g, tab, gla, BINS)
    with pytest.raises(api.Is3dError):
        api.spacetime_distributions(dict(fl, mode=2), cells, sp, g, tab, gla, BINS)                 # no anisotropic dN_dX routine
    with pytest.raises(api.Is3dError):
        api.spacetime_distributions(fl, cells, sp, g, tab, gla, dict(BINS, tau_bins=0))


@pytest.mark.parametrize("dimension,df_mode", [(3, 1), (3, 3), (2, 2), (2, 4)])
def test_spacetime_nonstandard_grids(fx, dimension, df_mode):
    """table lengths that do not line up with warps or register tiles: 13 pT points (a warp spans several species, a species
    spans two warps), 5 phi points, 9 y points, 37 eta points; 11 species"""
    fl, cells, sp, g0, tab, gla = problem(fx, 120 if dimension == 3 else 24, dimension, df_mode, 9800 + df_mode,
                                          chosen=fx["chosen_urqmd"][:11])
    g = dict(g0)
    g["pT"] = g0["pT"][:13]; g["pT_weight"] = g0["pT_weight"][:13]
    g["phi"] = g0["phi"][:5]; g["phi_weight"] = g0["phi_weight"][:5]
    g["y"] = g0["y"][3:12]; g["eta"] = g0["eta"][100:137]; g["eta_weight"] = g0["eta_weight"][100:137]
    ref, skipped = cfo.spacetime(fl, cells, sp, g, tab, BINS, gla)
    for variant in (0, 1, 4):
        got, st = api.spacetime_distributions(fl, cells, sp, g, tab, gla, BINS, tile_variant=variant)
        assert st["cells_skipped_udsigma"] == skipped
        check(got, ref)
