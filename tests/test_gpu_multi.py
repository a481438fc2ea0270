"""One process, several GPUs behind the C ABI (is3d_b200_init_devices / *_multi): cells sharded over one host
thread + stream per device, spectra combined with one ncclAllReduce.  The single-device checks run on any GPU box; the others need
two or more GPUs (`gpurun --gpus 2`)."""
import os
import subprocess
import tempfile

import numpy as np
import pytest

from common import compare, jonah_tables, load_golden, problem_from_recipe, surface_columns
from is3d_b200 import api, synthetic, tables, workdir

pytestmark = pytest.mark.gpu


def _n_gpus():
    import torch
    return torch.cuda.device_count()


@pytest.fixture()
def one_device_after():
    yield
    api.init_devices(1)


def test_multi_entry_with_one_device_is_the_plain_call(fx, one_device_after):
    assert api.init_devices(1) == 1
    gold = load_golden("s3_df1")
    fl, cells, sp, g, tab, gla = problem_from_recipe(gold["recipe"], fx)
    a, sa = api.smooth_spectra(fl, cells, sp, g, tab, gla)
    b, sb = api.smooth_spectra(fl, cells, sp, g, tab, gla, multi=True)
    assert np.array_equal(a, b) and sb["n_gpus"] == 1
    assert compare(b, gold["dN"])["ok"]


def test_table_range_error_leaves_result_untouched(fx):
    """IS3D_ERR_TABLE_RANGE must be reported before anything is added into the caller's array (host and device memory)"""
    import torch
    gold = load_golden("s3_df1")
    fl, cells, sp, g, tab, gla = problem_from_recipe(gold["recipe"], fx)
    hot = dict(cells); hot["T"] = cells["T"].copy(); hot["T"][3] = 0.25
    out = np.full(gold["dN"].size, 7.0)
    with pytest.raises(api.Is3dError) as e:
        api.smooth_spectra(fl, hot, sp, g, tab, gla, out=out)
    assert e.value.code == 3 and np.all(out == 7.0)
    dev = {k: torch.tensor(v, device="cuda") for k, v in hot.items()}
    dout = torch.full((gold["dN"].size,), 7.0, dtype=torch.float64, device="cuda")
    with pytest.raises(api.Is3dError):
        api.smooth_spectra(fl, dev, sp, g, tab, gla, out=dout, memory="device")
    torch.cuda.synchronize()
    assert bool((dout == 7.0).all())


@pytest.mark.skipif("__import__('torch').cuda.device_count() < 2")
@pytest.mark.parametrize("df_mode", [1, 3, 4])
def test_two_or_more_gpus_match_one(fx, df_mode, one_device_after):
    n_dev = min(_n_gpus(), 4)
    sp = tables.species(fx, 1, "chosen_urqmd"); g = tables.grid(fx); tab = tables.df_tables(fx, 1); gla = tables.laguerre(fx)
    cells = synthetic.columns_to_cells(synthetic.surface_vh(6001, 515, stress=(df_mode == 3)), 1)       # ragged shards
    for k in ("dat", "dax", "day", "dan"):
        cells[k][100:117] *= -1.0                                                                        # skipped cells in shard 0
    if df_mode == 4:
        tab.update(jonah_tables(cells, fx, 1, gla))
    fl = tables.flags(df_mode=df_mode, dimension=3)
    api.init_devices(1)
    one, s1 = api.smooth_spectra(fl, cells, sp, g, tab, gla, multi=True)
    assert api.init_devices(n_dev) == n_dev
    many, sn = api.smooth_spectra(fl, cells, sp, g, tab, gla, multi=True)
    assert sn["n_gpus"] == n_dev and sn["allreduce_ms"] > 0.0
    assert sn["cells_skipped_udsigma"] == s1["cells_skipped_udsigma"] == 17
    assert sn["cells_feqmod_breakdown"] == s1["cells_feqmod_breakdown"]
    assert sn["evaluations"] == s1["evaluations"]
    nz = one != 0
    assert np.array_equal(many == 0, one == 0)
    assert np.max(np.abs(many[nz] - one[nz]) / np.abs(one[nz])) < 1e-12       # summation order only
    # result is ADDED into the caller's array
    acc = np.full(one.size, 1.0)
    api.smooth_spectra(fl, cells, sp, g, tab, gla, out=acc, multi=True)
    big = one > 1e-3
    assert np.max(np.abs((acc - 1.0)[big] - one[big]) / one[big]) < 1e-11


@pytest.mark.skipif("__import__('torch').cuda.device_count() < 2")
def test_spacetime_distributions_multi(fx, one_device_after):
    sp = tables.species(fx, 1, "chosen_pikp"); g = tables.grid(fx); tab = tables.df_tables(fx, 1); gla = tables.laguerre(fx)
    cells = synthetic.columns_to_cells(synthetic.surface_vh(3000, 99), 1)
    fl = tables.flags(df_mode=2, dimension=3)
    bins = dict(tau_min=0.0, tau_max=12.0, tau_bins=24, r_min=0.0, r_max=12.0, r_bins=12)
    api.init_devices(1)
    one, _ = api.spacetime_distributions(fl, cells, sp, g, tab, gla, bins, multi=True)
    api.init_devices(2)
    two, st = api.spacetime_distributions(fl, cells, sp, g, tab, gla, bins, multi=True)
    assert st["n_gpus"] == 2
    for k in one:
        assert np.array_equal(one[k] == 0, two[k] == 0)
        assert np.abs(one[k] - two[k]).max() <= 1e-12 * np.abs(one[k]).max()


@pytest.mark.skipif("__import__('torch').cuda.device_count() < 2")
def test_executable_uses_all_gpus_and_writes_identical_files(fx):
    """the drop-in executable (RuniS3D equivalent) on a work directory with IS3D_B200_GPUS = 1 and = all: the results/ files are
    byte-for-byte the same (9 significant digits)"""
    exe = os.path.join(os.path.dirname(api.LIB_PATH), "is3d_b200_run")
    gold = load_golden("s3_df1")
    outs = []
    for n in (1, _n_gpus()):
        with tempfile.TemporaryDirectory() as wd:
            workdir.materialize(wd, surface_columns=surface_columns(gold["recipe"], fx), chosen=gold["recipe"]["chosen"], fixture=fx,
                                operation=1, mode=1, **gold["recipe"]["params"])
            r = subprocess.run([exe], cwd=wd, capture_output=True, text=True, env=dict(os.environ, IS3D_B200_GPUS=str(n)))
            assert r.returncode == 0, r.stdout + r.stderr
            outs.append({f: open(os.path.join(wd, "results", f), "rb").read() for f in sorted(os.listdir(os.path.join(wd, "results")))
                         if os.path.isfile(os.path.join(wd, "results", f))})
    assert outs[0].keys() == outs[1].keys() and len(outs[0]) >= 4
    for f in outs[0]:
        assert outs[0][f] == outs[1][f], f
