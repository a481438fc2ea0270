// cf_epilogue.cuh -- what a hot-kernel block does with its register tile once every cell of its chunk has been streamed:
// either the spectra bins of partial[chunk][...] (operation = 1) or the momentum-integrated numbers of operation = 0.
// Shared by cf_kernels.cu and cf_factored.cu.
#pragma once
#include "cf_internal.h"

namespace is3d {

// acc[(DIM2 ? 0 : j * NPT) + k]: slot j, phi point k of this thread's (species, pT) lane.
// scratch: block-shared, at least per_slot * blockDim.x doubles (operation = 0 only); every thread of the block must call.
template <int NYT, int NPT, bool DIM2>
__device__ __forceinline__ void hot_epilogue(const HotParams &hp, const double *acc, double *scratch, int chunk, int gb, int ty, int tp,
                                             bool lane_valid, int ipart, int ipT)
{
  const Layout &L = hp.L;
  // ---- operation = 0: integrate over (pT, phi) with the table weights (smooth_kernels.cpp:1284-1371), one number per
  //      (species, chunk) -- or per (species, slot) -- instead of the spectra bins
  if (hp.integ_mode) {
    constexpr int NJ = DIM2 ? 1 : NYT;
    const int per_slot = (hp.integ_mode == 2) ? NJ : 1;
    double *red = scratch;                              // [per_slot][blockDim.x]
    const double wl = lane_valid ? hp.pT_weight[ipT] * (hp.prefactor * hp.degeneracy[ipart]) : 0.0;
    double tot = 0.0;
#pragma unroll
    for (int j = 0; j < NJ; j++) {
      double sj = 0.0;
      const bool slot_ok = DIM2 || ty * NYT + j < L.n_slots;
#pragma unroll
      for (int k = 0; k < NPT; k++) {
        const int iphi = tp * NPT + k;
        if (slot_ok && iphi < L.n_phi) sj = fma(hp.phi_weight[iphi], acc[(DIM2 ? 0 : j * NPT) + k], sj);
      }
      if (hp.integ_mode == 2) red[j * blockDim.x + threadIdx.x] = wl * sj;
      tot += sj;
    }
    if (hp.integ_mode == 1) red[threadIdx.x] = wl * tot;
    __syncthreads();
    const int pair0 = gb * (int)blockDim.x, n_pairs = L.n_species * L.n_pT;
    const int s_first = pair0 / L.n_pT;
    for (int w = threadIdx.x; w < per_slot * hp.integ_sl; w += blockDim.x) {
      const int j = w / hp.integ_sl, sl = w - j * hp.integ_sl;
      const int sidx = s_first + sl;
      int lo = sidx * L.n_pT, hi = lo + L.n_pT;
      if (lo < pair0) lo = pair0;
      if (hi > pair0 + (int)blockDim.x) hi = pair0 + (int)blockDim.x;
      if (hi > n_pairs) hi = n_pairs;
      double v = 0.0;
      for (int q = lo; q < hi; q++) v += red[j * blockDim.x + (q - pair0)];
      const int64_t unit = (hp.integ_mode == 2) ? (int64_t)chunk * (L.n_ytiles * NJ) + ty * NJ + j : (int64_t)chunk * L.n_ytiles + ty;
      hp.integ[((unit * L.n_ptiles + tp) * hp.n_groupblocks + gb) * hp.integ_sl + sl] = v;
    }
    return;
  }

  // ---- operation = 1: partial[chunk][ipart + n_species (ipT + n_pT (iphi + n_phi iy))]
  if (lane_valid) {
    const double scale = hp.prefactor * hp.degeneracy[ipart];
    const int64_t n_bins = (int64_t)L.n_species * L.n_pT * L.n_phi * L.n_y_out;
    double *out = hp.partial + (int64_t)chunk * n_bins;
#pragma unroll
    for (int j = 0; j < (DIM2 ? 1 : NYT); j++) {
      const int iy = DIM2 ? 0 : ty * NYT + j;
      if (iy >= (DIM2 ? 1 : L.n_slots)) continue;
#pragma unroll
      for (int k = 0; k < NPT; k++) {
        const int iphi = tp * NPT + k;
        if (iphi >= L.n_phi) continue;
        const int64_t iS3D = (int64_t)ipart + (int64_t)L.n_species * ((int64_t)ipT + (int64_t)L.n_pT * ((int64_t)iphi + (int64_t)L.n_phi * iy));
        out[iS3D] = scale * acc[(DIM2 ? 0 : j * NPT) + k];
      }
    }
  }
}

// bytes of block-shared scratch hot_epilogue() needs for this launch
inline size_t hot_epilogue_scratch_bytes(const HotParams &hp, int nyt, bool dim2_summed, int threads)
{
  if (!hp.integ_mode) return 0;
  const int per_slot = (hp.integ_mode == 2) ? (dim2_summed ? 1 : nyt) : 1;
  return (size_t)per_slot * threads * sizeof(double);
}

}  // namespace is3d
