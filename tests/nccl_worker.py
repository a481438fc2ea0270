"""torchrun worker for test_two_gpu_nccl: shard by cell, NCCL all-reduce, rank 0 also computes the single-GPU result."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from is3d_b200 import api, distributed, synthetic, tables  # noqa: E402

out = sys.argv[1]
rank = int(os.environ["RANK"]); local = int(os.environ.get("LOCAL_RANK", rank))
torch.cuda.set_device(local)
dist.init_process_group("nccl")
api.init()
fx = tables.load_fixture()
cells = synthetic.columns_to_cells(synthetic.surface_vh(3001, 11), 1)
sp = tables.species(fx, 1, "chosen_pikp"); g = tables.grid(fx); tab = tables.df_tables(fx, 1)
fl = tables.flags(df_mode=1, dimension=3)
dev = {k: torch.tensor(v, device="cuda") for k, v in cells.items()}
dN, st = distributed.smooth_spectra_sharded(fl, dev, sp, g, tab, None, memory="device")
torch.cuda.synchronize()
np.save(os.path.join(out, "dN_%d.npy" % rank), dN.cpu().numpy())
if rank == 0:
    single, _ = api.smooth_spectra(fl, dev, sp, g, tab, None, memory="device")
    np.save(os.path.join(out, "dN_single.npy"), single.cpu().numpy())
# operation = 0: the histograms are sums over cells, one all-reduce of the concatenated raw sums
bins = dict(tau_min=0.0, tau_max=12.0, tau_bins=24, r_min=0.0, r_max=12.0, r_bins=12)
res, _ = distributed.spacetime_distributions_sharded(fl, dev, sp, g, tab, None, bins, memory="device")
np.savez(os.path.join(out, "st_%d.npz" % rank), **res)
if rank == 0:
    single, _ = api.spacetime_distributions(fl, dev, sp, g, tab, None, bins, memory="device")
    np.savez(os.path.join(out, "st_single.npz"), **single)
dist.barrier()
dist.destroy_process_group()
