// cf_api.cu -- the extern "C" boundary (include/is3d_b200.h): workspace management, host<->device staging, kernel
// sequencing and CUDA-event timing.  There is no CPU fallback: without a CUDA device every entry point fails.
#include "cf_internal.h"
#include "host_math.h"
#include <cmath>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

namespace is3d {

static std::mutex g_mutex;
static std::string g_last_error;
static bool g_init = false;
static int g_sm_count = 0;

#define CU_CHECK(call)                                                                          \
  do {                                                                                          \
    cudaError_t e__ = (call);                                                                   \
    if (e__ != cudaSuccess) {                                                                   \
      char buf__[512];                                                                          \
      snprintf(buf__, sizeof(buf__), "%s:%d: %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
      g_last_error = buf__;                                                                     \
      return IS3D_ERR_CUDA;                                                                     \
    }                                                                                           \
  } while (0)

// a grow-only device buffer
struct DevBuf {
  void *p = nullptr; size_t cap = 0;
  cudaError_t reserve(size_t bytes)
  {
    if (bytes <= cap) return cudaSuccess;
    if (p) { cudaFree(p); p = nullptr; cap = 0; }
    size_t want = bytes + bytes / 8 + 256;
    cudaError_t e = cudaMalloc(&p, want);
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
  template <class T> T *as() { return reinterpret_cast<T *>(p); }
};

struct Workspace {
  DevBuf raw, small, Y, P, S, Y2, P2, S2, partial, dN, counters, extra;
  cudaEvent_t ev[8];
  bool events = false;
  void release()
  {
    raw.release(); small.release(); Y.release(); P.release(); S.release(); Y2.release(); P2.release(); S2.release();
    partial.release(); dN.release();
    counters.release(); extra.release();
    if (events) { for (auto &e : ev) cudaEventDestroy(e); events = false; }
  }
};
static Workspace g_ws;

// bump allocator over the `small` buffer for tables; all offsets 256-byte aligned
struct SmallArena {
  std::vector<unsigned char> host;
  size_t put(const double *src, size_t n)
  {
    size_t off = (host.size() + 255) & ~(size_t)255;
    host.resize(off + n * sizeof(double));
    if (n) memcpy(host.data() + off, src, n * sizeof(double));
    return off;
  }
};

static int fail(int code, const char *msg) { g_last_error = msg; return code; }

}  // namespace is3d

using namespace is3d;

extern "C" {

int is3d_b200_version(void) { return 100; }

const char *is3d_b200_strerror(int code)
{
  switch (code) {
    case IS3D_OK: return "ok";
    case IS3D_ERR_ARGUMENT: return "invalid argument";
    case IS3D_ERR_UNSUPPORTED: return "flag combination not supported by the smooth Cooper-Frye path";
    case IS3D_ERR_TABLE_RANGE: return "cell temperature or Pi/P outside the delta-f coefficient table";
    case IS3D_ERR_CUDA: return "CUDA runtime error";
    case IS3D_ERR_NO_DEVICE: return "no CUDA device (this library has no CPU fallback)";
    case IS3D_ERR_IO: return "input/output error in the host layer";
    default: return "unknown error";
  }
}

const char *is3d_b200_last_error(void) { return g_last_error.c_str(); }

int is3d_b200_init(void)
{
  std::lock_guard<std::mutex> lk(g_mutex);
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) return fail(IS3D_ERR_NO_DEVICE, "no CUDA device visible");
  int dev = 0;
  CU_CHECK(cudaGetDevice(&dev));
  CU_CHECK(cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev));
  if (!g_ws.events) {
    for (auto &e : g_ws.ev) CU_CHECK(cudaEventCreate(&e));
    g_ws.events = true;
  }
  g_init = true;
  return IS3D_OK;
}

int is3d_b200_shutdown(void)
{
  std::lock_guard<std::mutex> lk(g_mutex);
  g_ws.release();
  g_init = false;
  return IS3D_OK;
}

int is3d_b200_measure_fp64_peak(double *tflops, double *ms_out)
{
  int rc = is3d_b200_init();
  if (rc) return rc;
  std::lock_guard<std::mutex> lk(g_mutex);
  CU_CHECK(g_ws.counters.reserve(256));
  int blocks = 0, threads = 0; long long per_thread = 0;
  const int iters = 20000;
  CU_CHECK(launch_fp64_peak(g_ws.counters.as<double>(), 2000, 0, &blocks, &threads, &per_thread));   // warm-up
  float best = 1e30f;
  for (int rep = 0; rep < 3; rep++) {
    CU_CHECK(cudaEventRecord(g_ws.ev[0], 0));
    CU_CHECK(launch_fp64_peak(g_ws.counters.as<double>(), iters, 0, &blocks, &threads, &per_thread));
    CU_CHECK(cudaEventRecord(g_ws.ev[1], 0));
    CU_CHECK(cudaEventSynchronize(g_ws.ev[1]));
    float ms = 0;
    CU_CHECK(cudaEventElapsedTime(&ms, g_ws.ev[0], g_ws.ev[1]));
    if (ms < best) best = ms;
  }
  const double flops = 2.0 * (double)per_thread * blocks * threads;
  if (tflops) *tflops = flops / (best * 1e-3) * 1e-12;
  if (ms_out) *ms_out = best;
  return IS3D_OK;
}

int is3d_b200_measure_fp64_sustained(double seconds, double *tflops)
{
  int rc = is3d_b200_init();
  if (rc) return rc;
  std::lock_guard<std::mutex> lk(g_mutex);
  CU_CHECK(g_ws.counters.reserve(256));
  int blocks = 0, threads = 0; long long per_thread = 0;
  const int iters = 20000;                       // ~21 ms per launch at the burst clock
  CU_CHECK(launch_fp64_peak(g_ws.counters.as<double>(), 2000, 0, &blocks, &threads, &per_thread));
  CU_CHECK(cudaDeviceSynchronize());
  int launches = (int)(seconds / 0.021) + 1;
  CU_CHECK(cudaEventRecord(g_ws.ev[0], 0));
  for (int i = 0; i < launches; i++) CU_CHECK(launch_fp64_peak(g_ws.counters.as<double>(), iters, 0, &blocks, &threads, &per_thread));
  CU_CHECK(cudaEventRecord(g_ws.ev[1], 0));
  CU_CHECK(cudaEventSynchronize(g_ws.ev[1]));
  float ms = 0;
  CU_CHECK(cudaEventElapsedTime(&ms, g_ws.ev[0], g_ws.ev[1]));
  const double flops = 2.0 * (double)per_thread * blocks * threads * launches;
  if (tflops) *tflops = flops / (ms * 1e-3) * 1e-12;
  return IS3D_OK;
}

int is3d_b200_smooth_spectra(const is3d_flags *fl, const is3d_surface *sf, const is3d_species *sp, const is3d_grid *gr,
                             const is3d_df_tables *df, const is3d_laguerre *gla, const is3d_options *opt_in,
                             double *dN_out, is3d_stats *stats)
{
  if (!fl || !sf || !sp || !gr || !dN_out) return fail(IS3D_ERR_ARGUMENT, "NULL argument");
  if (!g_init) { int rc = is3d_b200_init(); if (rc) return rc; }
  std::lock_guard<std::mutex> lk(g_mutex);
  is3d_options opt; memset(&opt, 0, sizeof(opt));
  if (opt_in) opt = *opt_in;
  cudaStream_t st = (cudaStream_t)opt.stream;
  is3d_stats stt; memset(&stt, 0, sizeof(stt));

  // ---- validation / dispatch (emissionfunction.cpp:1503-1673)
  const bool vah = (fl->mode == 2);
  if (fl->dimension != 2 && fl->dimension != 3) return fail(IS3D_ERR_ARGUMENT, "dimension must be 2 or 3");
  if (sp->n <= 0 || gr->n_pT <= 0 || gr->n_phi <= 0 || gr->n_y <= 0) return fail(IS3D_ERR_ARGUMENT, "empty species list or momentum table");
  if (!sp->mass || !sp->sign || !sp->degeneracy || !gr->pT || !gr->phi || !gr->y) return fail(IS3D_ERR_ARGUMENT, "NULL species or momentum table");
  if (fl->dimension == 2 && (gr->n_eta <= 0 || !gr->eta || !gr->eta_weight)) return fail(IS3D_ERR_ARGUMENT, "dimension = 2 needs the eta table");
  if (fl->include_baryon) return fail(IS3D_ERR_UNSUPPORTED, "include_baryon = 1 (the reference's (T, muB) lookup reads out of bounds, SURVEY R8)");
  if (!vah && (fl->df_mode < 1 || fl->df_mode > 4)) return fail(IS3D_ERR_ARGUMENT, "df_mode must be 1..4");
  if (!vah && (!df || df->n_T < 3 || !df->T)) return fail(IS3D_ERR_ARGUMENT, "delta-f coefficient tables missing");
  const bool feqmod = !vah && (fl->df_mode == 3 || fl->df_mode == 4);
  const bool ideal = !vah && (fl->df_mode == 1 || fl->df_mode == 2) && !fl->include_shear_deltaf && !fl->include_bulk_deltaf;
  const int model = vah ? M_VAH : ideal ? M_IDEAL : (fl->df_mode == 1 ? M_LIN14 : fl->df_mode == 2 ? M_LINCE : M_FEQMOD);
  if (!vah) {
    if (fl->df_mode == 1 && (!df->c0 || !df->c2)) return fail(IS3D_ERR_ARGUMENT, "df_mode 1 needs the c0 and c2 tables");
    if ((fl->df_mode == 2 || fl->df_mode == 3) && (!df->F || !df->betabulk || !df->betapi)) return fail(IS3D_ERR_ARGUMENT, "df_mode 2/3 need the F, betabulk and betapi tables");
    if (fl->df_mode == 4 && (!df->betapi || df->n_jonah < 3 || !df->jonah_x || !df->jonah_lambda2 || !df->jonah_z))
      return fail(IS3D_ERR_ARGUMENT, "df_mode 4 needs betapi and the Jonah lambda/z tables");
    if (fl->df_mode == 3 && (!gla || gla->n_points <= 0 || !gla->root1 || !gla->weight1 || !gla->root2 || !gla->weight2))
      return fail(IS3D_ERR_ARGUMENT, "df_mode 3 needs the Gauss-Laguerre nodes");
  }
  const int64_t n_cells = sf->n_cells;
  if (n_cells < 0) return fail(IS3D_ERR_ARGUMENT, "negative cell count");
  const bool dim2 = (fl->dimension == 2);
  const int64_t n_bins = (int64_t)sp->n * gr->n_pT * gr->n_phi * gr->n_y;

  const bool dim2_early = (fl->dimension == 2);
  // ---- layout
  Layout L; memset(&L, 0, sizeof(L));
  L.n_species = sp->n; L.n_pT = gr->n_pT; L.n_phi = gr->n_phi; L.n_y_out = gr->n_y;
  L.dim2 = dim2 ? 1 : 0;
  L.n_slots = dim2 ? gr->n_eta : gr->n_y;
  L.rec_y = vah ? kRecVah : kRec;
  // tile_variant: 0 = model default (tuned on B200, see profiles/), k > 0 = table entry k - 1 (tuning / tests)
  int variant;
  if (opt.tile_variant >= 1 && opt.tile_variant <= 16) variant = opt.tile_variant - 1;
  else if (dim2_early) variant = (model == M_FEQMOD || model == M_VAH) ? 12 : (model == M_IDEAL ? 13 : 10);
  else variant = (model == M_FEQMOD || model == M_VAH) ? 12 : 9;
  int nyt, npt, ct;
  hot_variant_shape(variant, L.dim2, &nyt, &npt, &ct);
  L.nst = dim2 ? L.n_slots : nyt;
  L.n_ytiles = dim2 ? 1 : (L.n_slots + nyt - 1) / nyt;
  L.npt = npt; L.n_ptiles = (L.n_phi + npt - 1) / npt;
  L.ct = ct;
  L.n_cells = n_cells;
  L.n_tiles = (n_cells + ct - 1) / ct;
  L.n_cells_pad = L.n_tiles * ct;

  const int n_pairs = sp->n * gr->n_pT;
  const int n_groups = (n_pairs + 31) / 32;
  int n_warps = kMaxWarps;                                  // widest block whose padding wastes <= 4 % of the warps
  {
    int best = 1 << 30;
    for (int w = kMaxWarps; w >= 1; w--) {
      const int launched = ((n_groups + w - 1) / w) * w;
      if ((launched - n_groups) * 25 <= launched) { n_warps = w; break; }
      if (launched < best) { best = launched; n_warps = w; }
    }
  }
  const int n_groupblocks = (n_groups + n_warps - 1) / n_warps;
  const int64_t n_bintiles = (int64_t)n_groupblocks * L.n_ytiles * L.n_ptiles;
  int n_chunks = opt.n_chunks;
  if (n_chunks <= 0) {
    const int64_t target_blocks = (int64_t)g_sm_count * 96;             // >= 16 waves at 6 blocks/SM: small tail
    n_chunks = (int)((target_blocks + n_bintiles - 1) / n_bintiles);
    const int64_t max_partial_bytes = (int64_t)2 << 30;                 // keep the partial buffer <= 2 GiB
    int64_t cap = max_partial_bytes / (n_bins * 8 > 0 ? n_bins * 8 : 1);
    if (feqmod) cap /= 2;
    if (cap < 1) cap = 1;
    if (n_chunks > cap) n_chunks = (int)cap;
  }
  if ((int64_t)n_chunks > L.n_tiles) n_chunks = (int)(L.n_tiles > 0 ? L.n_tiles : 1);
  if (n_chunks < 1) n_chunks = 1;
  if (n_bintiles * n_chunks > 2147483647LL) return fail(IS3D_ERR_ARGUMENT, "grid too large");

  // ---- small tables -> one staging buffer
  SmallArena ar;
  std::vector<double> cosphi(gr->n_phi), sinphi(gr->n_phi);
  for (int k = 0; k < gr->n_phi; k++) { cosphi[k] = cos(gr->phi[k]); sinphi[k] = sin(gr->phi[k]); }   // smooth_kernels.cpp:43-48
  std::vector<double> zero_baryon(sp->n, 0.0);
  const size_t o_mass = ar.put(sp->mass, sp->n), o_sign = ar.put(sp->sign, sp->n), o_deg = ar.put(sp->degeneracy, sp->n);
  const size_t o_bar = ar.put(sp->baryon ? sp->baryon : zero_baryon.data(), sp->n);
  const size_t o_pT = ar.put(gr->pT, gr->n_pT), o_cos = ar.put(cosphi.data(), gr->n_phi), o_sin = ar.put(sinphi.data(), gr->n_phi);
  const size_t o_sloty = dim2 ? ar.put(gr->eta, gr->n_eta) : ar.put(gr->y, gr->n_y);
  const size_t o_slotw = dim2 ? ar.put(gr->eta_weight, gr->n_eta) : 0;
  struct SplineOff { size_t x, y, c; int n; };
  auto put_spline = [&](const double *x, const double *y, int n) {
    SplineOff o{0, 0, 0, 0};
    if (!x || !y || n < 3) return o;
    std::vector<double> c(n);
    host_spline_init(x, y, n, c.data());
    o.x = ar.put(x, n); o.y = ar.put(y, n); o.c = ar.put(c.data(), n); o.n = n;
    return o;
  };
  SplineOff s_c0{}, s_c2{}, s_F{}, s_bb{}, s_bp{}, s_l2{}, s_z{};
  size_t o_r1 = 0, o_w1 = 0, o_r2 = 0, o_w2 = 0;
  if (!vah) {
    s_c0 = put_spline(df->T, df->c0, df->n_T); s_c2 = put_spline(df->T, df->c2, df->n_T); s_F = put_spline(df->T, df->F, df->n_T);
    s_bb = put_spline(df->T, df->betabulk, df->n_T); s_bp = put_spline(df->T, df->betapi, df->n_T);
    if (fl->df_mode == 4) { s_l2 = put_spline(df->jonah_x, df->jonah_lambda2, df->n_jonah); s_z = put_spline(df->jonah_x, df->jonah_z, df->n_jonah); }
    if (fl->df_mode == 3) {
      o_r1 = ar.put(gla->root1, gla->n_points); o_w1 = ar.put(gla->weight1, gla->n_points);
      o_r2 = ar.put(gla->root2, gla->n_points); o_w2 = ar.put(gla->weight2, gla->n_points);
    }
  }

  // ---- raw surface arrays
  const bool sh = fl->include_shear_deltaf != 0, bk = fl->include_bulk_deltaf != 0;
  struct RawItem { const double *src; bool need; const double **dst; };
  RawCells rc; memset(&rc, 0, sizeof(rc));
  rc.n = n_cells;
  RawItem items[] = {
    {sf->tau, true, &rc.tau}, {sf->eta, true, &rc.eta}, {sf->dat, true, &rc.dat}, {sf->dax, true, &rc.dax}, {sf->day, true, &rc.day},
    {sf->dan, true, &rc.dan}, {sf->ux, true, &rc.ux}, {sf->uy, true, &rc.uy}, {sf->un, true, &rc.un},
    {sf->T, !vah, &rc.T}, {sf->P, !vah, &rc.P}, {sf->E, !vah, &rc.E},
    {sf->pixx, vah || sh, &rc.pixx}, {sf->pixy, vah || sh, &rc.pixy}, {sf->pixn, vah || sh, &rc.pixn}, {sf->piyy, vah || sh, &rc.piyy},
    {sf->piyn, vah || sh, &rc.piyn}, {sf->bulkPi, vah ? bk : bk, &rc.bulkPi},
    {sf->pitt, vah, &rc.pitt}, {sf->pitx, vah, &rc.pitx}, {sf->pity, vah, &rc.pity}, {sf->pitn, vah, &rc.pitn}, {sf->pinn, vah, &rc.pinn},
    {sf->Wx, vah, &rc.Wx}, {sf->Wy, vah, &rc.Wy}, {sf->Lambda, vah, &rc.Lambda}, {sf->aL, vah, &rc.aL},
    {sf->c0, vah, &rc.c0}, {sf->c1, vah, &rc.c1}, {sf->c2, vah, &rc.c2}, {sf->c3, vah, &rc.c3}, {sf->c4, vah, &rc.c4}};
  const int n_raw = (int)(sizeof(items) / sizeof(items[0]));
  for (int a = 0; a < n_raw; a++)
    if (items[a].need && !items[a].src && n_cells > 0) return fail(IS3D_ERR_ARGUMENT, "a required surface array is NULL");

  const size_t rec_Y = (size_t)L.n_ytiles * L.n_cells_pad * L.nst * L.rec_y * 8;
  const size_t rec_P = (size_t)L.n_ptiles * L.n_cells_pad * L.npt * kRec * 8;
  const size_t rec_S = (size_t)L.n_cells_pad * kScal * 8;
  const int partial_sets = feqmod ? 2 : 1;
  CU_CHECK(g_ws.small.reserve(ar.host.size() + 256));
  CU_CHECK(g_ws.Y.reserve(rec_Y + 256));
  CU_CHECK(g_ws.P.reserve(rec_P + 256));
  CU_CHECK(g_ws.S.reserve(rec_S + 256));
  if (feqmod) {
    CU_CHECK(g_ws.Y2.reserve(rec_Y + 256));
    CU_CHECK(g_ws.P2.reserve(rec_P + 256));
    CU_CHECK(g_ws.S2.reserve(rec_S + 256));
    if (fl->df_mode == 3) CU_CHECK(g_ws.extra.reserve(((size_t)L.n_species + 8) * L.n_cells_pad * 8 + 256));
  }
  CU_CHECK(g_ws.partial.reserve((size_t)partial_sets * n_chunks * n_bins * 8 + 256));
  CU_CHECK(g_ws.counters.reserve(256));
  const size_t cell_stride = ((size_t)n_cells * 8 + 255) & ~(size_t)255;
  if (opt.memory == 0) {
    CU_CHECK(g_ws.raw.reserve(cell_stride * n_raw + 256));
    CU_CHECK(g_ws.dN.reserve((size_t)n_bins * 8 + 256));
  }

  cudaEvent_t *ev = g_ws.ev;
  CU_CHECK(cudaEventRecord(ev[0], st));
  // ---- host -> device
  unsigned char *small_d = g_ws.small.as<unsigned char>();
  CU_CHECK(cudaMemcpyAsync(small_d, ar.host.data(), ar.host.size(), cudaMemcpyHostToDevice, st));
  for (int a = 0; a < n_raw; a++) {
    *items[a].dst = nullptr;
    if (!items[a].need || n_cells == 0) continue;
    if (opt.memory == 0) {
      double *dst = reinterpret_cast<double *>(g_ws.raw.as<unsigned char>() + cell_stride * a);
      CU_CHECK(cudaMemcpyAsync(dst, items[a].src, (size_t)n_cells * 8, cudaMemcpyHostToDevice, st));
      *items[a].dst = dst;
    } else *items[a].dst = items[a].src;
  }
  double *dN_dev = (opt.memory == 0) ? g_ws.dN.as<double>() : dN_out;
  if (opt.memory == 0) CU_CHECK(cudaMemsetAsync(dN_dev, 0, (size_t)n_bins * 8, st));
  CU_CHECK(cudaMemsetAsync(g_ws.counters.p, 0, sizeof(PrepCounters), st));
  CU_CHECK(cudaEventRecord(ev[1], st));

  // ---- prepare
  auto dptr = [&](size_t off) { return reinterpret_cast<const double *>(small_d + off); };
  PrepTables tab; memset(&tab, 0, sizeof(tab));
  auto mk = [&](const SplineOff &o) { Spline s; s.x = dptr(o.x); s.y = dptr(o.y); s.c = dptr(o.c); s.n = o.n; return s; };
  tab.c0 = mk(s_c0); tab.c2 = mk(s_c2); tab.F = mk(s_F); tab.betabulk = mk(s_bb); tab.betapi = mk(s_bp); tab.lam2 = mk(s_l2); tab.z = mk(s_z);
  tab.bulkPi_over_Peq_max = df ? df->bulkPi_over_Peq_max : 0.0;
  tab.cosphi = dptr(o_cos); tab.sinphi = dptr(o_sin); tab.slot_y = dptr(o_sloty); tab.slot_w = dim2 ? dptr(o_slotw) : nullptr;
  tab.gla_root1 = dptr(o_r1); tab.gla_w1 = dptr(o_w1); tab.gla_root2 = dptr(o_r2); tab.gla_w2 = dptr(o_w2); tab.gla_n = gla ? gla->n_points : 0;
  tab.deta_min = fl->deta_min; tab.mass_pion0 = fl->mass_pion0;
  tab.eta_delta = (dim2 && gr->n_eta > 1) ? gr->eta[1] - gr->eta[0] : 0.0;                    // smooth_kernels.cpp:2175
  PrepCounters *cnt_d = g_ws.counters.as<PrepCounters>();
  if (vah) {
    CU_CHECK(launch_prepare_vah(*fl, rc, tab, L, g_ws.Y.as<double>(), g_ws.P.as<double>(), g_ws.S.as<double>(), cnt_d, st));
    stt.gpu_launches++;
  } else if (feqmod) {
    CU_CHECK(launch_prepare_feqmod(*fl, rc, tab, L, g_ws.Y.as<double>(), g_ws.P.as<double>(), g_ws.S.as<double>(),
                                   g_ws.Y2.as<double>(), g_ws.P2.as<double>(), g_ws.S2.as<double>(),
                                   dptr(o_mass), dptr(o_sign), dptr(o_deg), dptr(o_bar),
                                   fl->df_mode == 3 ? g_ws.extra.as<double>() : nullptr, cnt_d, st));
    stt.gpu_launches += (fl->df_mode == 3) ? 2 : 1;
  } else {
    CU_CHECK(launch_prepare_vh(*fl, rc, tab, L, g_ws.Y.as<double>(), g_ws.P.as<double>(), g_ws.S.as<double>(), cnt_d, st));
    stt.gpu_launches++;
  }
  CU_CHECK(cudaEventRecord(ev[2], st));

  // ---- hot kernel(s)
  HotParams hp; memset(&hp, 0, sizeof(hp));
  hp.L = L; hp.Y = g_ws.Y.as<double>(); hp.P = g_ws.P.as<double>(); hp.S = g_ws.S.as<double>();
  hp.mass = dptr(o_mass); hp.sign = dptr(o_sign); hp.degeneracy = dptr(o_deg); hp.pT = dptr(o_pT);
  hp.partial = g_ws.partial.as<double>();
  hp.renorm = (feqmod && fl->df_mode == 3) ? g_ws.extra.as<double>() : nullptr;
  hp.n_chunks = n_chunks; hp.n_groupblocks = n_groupblocks; hp.n_warps = n_warps;
  hp.regulate_thr = fl->regulate_deltaf ? 0x3ff00000 : 0x7ff80000;
  hp.outflow_thr = (fl->outflow && !vah) ? 0LL : (long long)0x8000000000000000ULL;   // the anisotropic kernel has no Theta(p.dsigma)
  const double hbarC = 0.197327053;
  hp.prefactor = vah ? 1.0 / (8.0 * (M_PI * M_PI * M_PI)) / hbarC / hbarC / hbarC      // smooth_kernels.cpp:2146
                     : pow(2.0 * M_PI * hbarC, -3);                                      // :36, :400
  CU_CHECK(launch_hot(model, hp, variant, st, nullptr));
  stt.gpu_launches++;
  int reduce_sets = 1;
  PrepCounters cnt; memset(&cnt, 0, sizeof(cnt));
  if (feqmod) {
    // cells where feqmod breaks down (and narrow-rapidity slots) take the linear-df branch: second pass, only if any
    CU_CHECK(cudaMemcpyAsync(&cnt, cnt_d, sizeof(cnt), cudaMemcpyDeviceToHost, st));
    CU_CHECK(cudaStreamSynchronize(st));
    if (cnt.linear_items > 0) {
      HotParams hl = hp;
      hl.Y = g_ws.Y2.as<double>(); hl.P = g_ws.P2.as<double>(); hl.S = g_ws.S2.as<double>();
      hl.partial = hp.partial + (size_t)n_chunks * n_bins; hl.renorm = nullptr;
      CU_CHECK(launch_hot(fl->df_mode == 3 ? M_LINCE : M_JONAHLIN, hl, variant, st, nullptr));
      stt.gpu_launches++;
      reduce_sets = 2;
    }
  }
  CU_CHECK(cudaEventRecord(ev[3], st));

  // ---- reduce chunks, add into the result
  CU_CHECK(launch_reduce(hp.partial, n_chunks * reduce_sets, n_bins, dim2 ? n_bins / gr->n_y : n_bins, dN_dev, st));
  stt.gpu_launches++;
  CU_CHECK(cudaEventRecord(ev[4], st));

  // ---- device -> host
  std::vector<double> host_dN;
  if (opt.memory == 0) {
    host_dN.resize((size_t)n_bins);
    CU_CHECK(cudaMemcpyAsync(host_dN.data(), dN_dev, (size_t)n_bins * 8, cudaMemcpyDeviceToHost, st));
  }
  CU_CHECK(cudaMemcpyAsync(&cnt, cnt_d, sizeof(cnt), cudaMemcpyDeviceToHost, st));
  CU_CHECK(cudaEventRecord(ev[5], st));
  CU_CHECK(cudaEventSynchronize(ev[5]));
  if (opt.memory == 0)
    for (int64_t i = 0; i < n_bins; i++) dN_out[i] += host_dN[(size_t)i];

  float ms;
  cudaEventElapsedTime(&ms, ev[0], ev[1]); stt.h2d_ms = ms;
  cudaEventElapsedTime(&ms, ev[1], ev[2]); stt.prepare_ms = ms;
  cudaEventElapsedTime(&ms, ev[2], ev[3]); stt.kernel_ms = ms;
  cudaEventElapsedTime(&ms, ev[3], ev[4]); stt.reduce_ms = ms;
  cudaEventElapsedTime(&ms, ev[4], ev[5]); stt.d2h_ms = ms;
  cudaEventElapsedTime(&ms, ev[0], ev[5]); stt.total_ms = ms;
  stt.cells_skipped_udsigma = (int64_t)cnt.skipped;
  stt.cells_feqmod_breakdown = (int64_t)cnt.breakdown;
  stt.evaluations = n_cells * (int64_t)sp->n * gr->n_pT * gr->n_phi * (dim2 ? (int64_t)gr->n_eta : (int64_t)gr->n_y);
  stt.n_chunks = n_chunks; stt.tile_variant = variant;
  if (stats) *stats = stt;
  if (cnt.range_error) return fail(IS3D_ERR_TABLE_RANGE, "a cell's T or Pi/P lies outside the delta-f coefficient table (or T_mod <= 0)");
  return IS3D_OK;
}

}  // extern "C"
