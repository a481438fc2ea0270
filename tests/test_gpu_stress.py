"""GPU vs oracle on deliberately nasty surfaces: fast longitudinal flow, tiny tau, temperatures at the coefficient-table
edges, no transverse flow (Milne-basis special case), bulk pressures that trigger the Jonah clamps and the feqmod
breakdown branch, |y - eta| up to 11.  Linear-df modes use the conditioning-aware criterion of common.compare (bins in which
1 + df nearly cancels carry that noise in the reference itself); feqmod modes use the plain 1e-10 bar."""
import numpy as np
import pytest

from common import compare, jonah_tables
from is3d_b200 import api, synthetic, tables

pytestmark = pytest.mark.gpu


def _variants():
    base = synthetic.surface_vh(48, 77)
    out = {"base": base}
    s = base.copy(); s[:, 10] = np.sinh(np.linspace(-2.0, 2.0, len(s))) / s[:, 0]; out["fast_longitudinal_flow"] = s
    s = base.copy(); s[:, 0] = np.linspace(0.05, 0.5, len(s)); s[:, 10] = 0.1 / s[:, 0]; out["small_tau"] = s
    s = base.copy(); s[:, 12] = np.linspace(0.1001, 0.1999, len(s)) / synthetic.HBARC; out["T_table_edges"] = s
    s = base.copy(); s[:, 8] = 0.0; s[:, 9] = 0.0; out["no_transverse_flow"] = s
    s = base.copy(); s[:, 19] = np.linspace(-1.5, 1.5, len(s)) * 0.05 / synthetic.HBARC; out["huge_bulk"] = s
    out["stress"] = synthetic.surface_vh(48, 78, stress=True)
    s = base.copy(); s[:, 3] = np.linspace(-6, 6, len(s)); out["far_eta"] = s
    for v in out.values():
        v[:, 4] = np.abs(v[:, 4]) * 5 + 0.5            # keep u.dsigma > 0
    return out


@pytest.mark.parametrize("name", list(_variants()))
@pytest.mark.parametrize("df_mode", [1, 2, 3, 4])
def test_stress_surface(fx, name, df_mode):
    from oracle import cf_oracle as cfo
    api.init()
    cells = synthetic.columns_to_cells(_variants()[name], 1)
    sp = tables.species(fx, 1, [211, 321, 2212, -3334, 337]); g = tables.grid(fx); gla = tables.laguerre(fx)
    for extra in ({}, dict(regulate_deltaf=0, outflow=0)):
        if extra and df_mode == 3 and name in ("huge_bulk", "stress"):
            continue        # unregulated linear fallback of breakdown cells: cancellations without a conditioning measure
        tab = tables.df_tables(fx, 1)
        if df_mode == 4:
            tab.update(jonah_tables(cells, fx, 1, gla))
        fl = tables.flags(df_mode=df_mode, dimension=3, **extra)
        cond = np.zeros(5 * 32 * 24 * 21) if df_mode in (1, 2) else None
        ref, skipped, breakdown = cfo.smooth(fl, cells, sp, g, tab, gla, conditioning=cond)
        for variant in ((0, 13, 16, 22, 24, 25) if df_mode in (1, 2) else (0,)):  # default and the shifted-factor exponential (cf_kernel SB 5, cf_shift.cu)
            got, st = api.smooth_spectra(fl, cells, sp, g, tab, gla, tile_variant=variant)
            rep = compare(got, ref, conditioning=cond)
            assert rep["ok"], (extra, variant, rep)
            assert st["cells_feqmod_breakdown"] == breakdown and st["cells_skipped_udsigma"] == skipped
            assert np.isfinite(got).all()
