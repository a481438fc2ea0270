"""operation = 0 (spacetime distributions): CUDA path through the C ABI vs the oracle.

Tolerance: 1e-10 relative on every histogram bin that holds a non-negligible share of the species' yield (the bins are sums of
non-negative cell yields, so there is no cancellation across cells); bins below 1e-12 of the largest one are compared
absolutely against that scale.  Empty bins must be exactly zero."""
import numpy as np
import pytest

from common import jonah_tables
from is3d_b200 import api, synthetic, tables
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
