"""Sampler mean yield: oracle and C++ host layer against vectors printed by the unmodified reference
(tests/golden/yield_vectors.npz, 17 significant digits), CUDA surface reduction against the oracle."""
import json
import os

import numpy as np
import pytest

from common import GOLDEN_DIR, jonah_tables
from is3d_b200 import api, synthetic, tables
from oracle import cf_oracle as cfo

VEC = np.load(os.path.join(GOLDEN_DIR, "yield_vectors.npz"))
CASES = sorted({k.split("/")[0] for k in VEC.files})


def setup(case, fx):
    rec = json.loads(str(VEC[case + "/recipe"]))
    cols = synthetic.surface_vh(rec["n_cells"], rec["seed"], three_d=(rec["dimension"] == 3), stress=rec["stress"])
    cells = synthetic.columns_to_cells(cols, 1)
    sp = tables.species(fx, 1, "chosen_pikp"); tab = tables.df_tables(fx, 1); gla = tables.laguerre(fx)
    if rec["df_mode"] == 4:
        tab.update(jonah_tables(cells, fx, 1, gla))
    fl = tables.flags(df_mode=rec["df_mode"], dimension=rec["dimension"])
    gla3 = dict(gla, root3=fx["gla_root"][3].copy(), weight3=fx["gla_weight"][3].copy())
    return rec, fl, cells, sp, tab, gla, gla3


@pytest.mark.parametrize("case", CASES)
def test_oracle_yield_is_bit_identical_to_the_reference(case, fx):
    rec, fl, cells, sp, tab, gla, gla3 = setup(case, fx)
    avg = cfo.surface_averages(cells)
    avg = np.array([float("%.15g" % v) for v in avg])              # the averages side file carries 15 digits
    neq, bulk, _ = cfo.particle_densities(sp, avg, rec["df_mode"], tab, gla, gla3["root3"], gla3["weight3"])
    assert np.array_equal(neq, VEC[case + "/densities"][:, 0]) and np.array_equal(bulk, VEC[case + "/densities"][:, 1])
    N, skipped = cfo.total_yield(fl, cells, neq, bulk, tab, rec["y_cut"])
    assert skipped == 0 and N == float(VEC[case + "/Ntot"])


@pytest.mark.parametrize("case", CASES)
def test_host_layer_densities(case, fx):
    """C++ host layer (no GPU work): same densities as the reference to a few ulp (libm pow/exp orderings)"""
    rec, fl, cells, sp, tab, gla, gla3 = setup(case, fx)
    avg = api.surface_averages(cells)
    neq, bulk, diff = api.particle_densities(sp, avg, rec["df_mode"], tab, gla3)
    ref = VEC[case + "/densities"]
    np.testing.assert_allclose(neq, ref[:, 0], rtol=1e-14)
    np.testing.assert_allclose(bulk, ref[:, 1], rtol=1e-13, atol=1e-300)


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES)
def test_gpu_mean_yield(case, fx):
    rec, fl, cells, sp, tab, gla, gla3 = setup(case, fx)
    ref = float(VEC[case + "/Ntot"])
    avg = api.surface_averages(cells)
    neq, bulk, _ = api.particle_densities(sp, avg, rec["df_mode"], tab, gla3)
    N, st = api.mean_yield(fl, cells, neq, bulk, tab, y_cut=rec["y_cut"])
    assert st["cells_skipped_udsigma"] == 0 and st["gpu_launches"] == 1
    assert abs(N - ref) <= 1e-12 * abs(ref)


@pytest.mark.gpu
def test_gpu_mean_yield_large_skips_and_device_memory(fx):
    import torch
    n = 300_000
    cells = synthetic.columns_to_cells(synthetic.surface_vh(n, 4711), 1)
    for k in ("dat", "dax", "day", "dan"):
        cells[k][::5] *= -1.0                                                  # u.dsigma < 0: skipped by the reference (:690)
    sp = tables.species(fx, 1, "chosen_urqmd"); tab = tables.df_tables(fx, 1); gla = tables.laguerre(fx)
    gla3 = dict(gla, root3=fx["gla_root"][3].copy(), weight3=fx["gla_weight"][3].copy())
    fl = tables.flags(df_mode=2, dimension=3)
    avg = api.surface_averages(cells)
    neq, bulk, _ = api.particle_densities(sp, avg, 2, tab, gla3)
    ref, skipped = cfo.total_yield(fl, cells, neq, bulk, tab, 5.0)
    assert skipped == n // 5
    N, st = api.mean_yield(fl, cells, neq, bulk, tab)
    assert st["cells_skipped_udsigma"] == skipped and abs(N - ref) <= 1e-12 * abs(ref)
    dev = {k: torch.from_numpy(np.ascontiguousarray(v)).cuda() for k, v in cells.items()}
    N2, _ = api.mean_yield(fl, dev, neq, bulk, tab, memory="device")
    assert N2 == N                                                              # same deterministic reduction
    empty = {k: v[:0] for k, v in cells.items()}
    assert api.mean_yield(fl, empty, neq, bulk, tab)[0] == 0.0
