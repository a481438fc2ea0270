"""Golden vectors for the (Lambda, alpha_L) -> c0..c4 lookup of the anisotropic model: inputs drawn over
and slightly beyond the table, outputs from the reference's own reader (oracle/_ref/vah_ref, built from
/root/reference/src/cuda/deltafReader.cu by `make -C oracle ref`).  Run here (needs /root/reference); commits
tests/golden/vah_coefficients.npz."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from is3d_b200 import tables      # noqa: E402
from oracle import cf_oracle as cfo      # noqa: E402

fx = tables.load_fixture()
rng = np.random.default_rng(20261018)
hb = 0.197327053
nL = int(fx["df_vah/nL"]); naL = int(fx["df_vah/naL"])
Lg = fx["df_vah/L_col"][:nL]; ag = fx["df_vah/aL_col"][::nL]
n = 4000
Lam_fm = rng.uniform(Lg[0] - 0.02, Lg[-1] + 0.02, n)
aL = rng.uniform(ag[0] - 0.02, ag[-1] + 0.02, n)
# exact grid nodes and cell edges are the interesting cases of a "first index with value > x" search
Lam_fm[:40] = rng.choice(Lg, 40); aL[40:80] = rng.choice(ag, 40)
Lam_fm[80:90] = Lg[rng.integers(0, nL, 10)]; aL[80:90] = ag[rng.integers(0, naL, 10)]
Lam_GeV = Lam_fm * hb
out = cfo.run_vah_reference(aL, Lam_GeV, fx)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "vah_coefficients.npz"), aL=aL, Lambda_GeV=Lam_GeV, c=out)
print("cells", n, "inside table", int((out[:, 0] != -12345.0).sum()))
