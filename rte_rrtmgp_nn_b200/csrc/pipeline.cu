// Whole-path drivers: what one pass of the reference drivers' block loop does
// (examples/rfmip-clear-sky/rrtmgp_rfmip_lw.F90:368-446, rrtmgp_rfmip_sw.F90:356-465):
//   LW: gas_optics(neural_nets=) -> rte_lw          SW: gas_optics(neural_nets=) -> boundary conditions -> rte_sw
// for ALL columns of a call, in column chunks whose optical-property arrays live in the context's persistent
// workspace (no allocation in steady state).  The *_host variants take host pointers and overlap the H2D copy
// of chunk k+1 and the D2H copy of chunk k-1 with the kernels of chunk k on three streams.
#include "common.cuh"
#include <algorithm>
#include <functional>
#include <thread>

namespace rrnn {
bool lw_v5_supports(int G, int L);  // rte_solvers_tma.cu
// rte_solvers.cu: clouds folded into the packed solvers (-1 = shape not taken)
int cloud_rows_lw(rrnn_ctx_t* ctx, size_t nsmp, int nbnd, const float* tau_bnd_d, float* rows_d);
int cloud_rows_sw(rrnn_ctx_t* ctx, size_t nsmp, int nbnd, const float* tau_bnd_d, const float* ssa_bnd_d, const float* g_bnd_d, float* rows_d);
int cloud_rows_fused(rrnn_ctx_t* ctx, const rrnn_cloud_lut_t* lut, int ncol, int nlay, const float* clwp_d, const float* ciwp_d,
                     const float* reliq_d, const float* reice_d, bool two_stream, float* rows_d);   // api.cu; -1: not taken
int lw_solver_clouds(rrnn_ctx_t* ctx, const rrnn_kdist_t* kd, int nlay, int ncol, int top_at_1, int nmus, const float* Ds, const float* weights,
                     const float* inc_flux_d, const float* tau_d, const float* lay_d, const float* lev_d, const float* planck_lay_d,
                     const float* planck_lev_d, const float* sfc_emis_gpt_d, const float* sfc_source_d, const float* cld_rows_d,
                     float* flux_up_d, float* flux_dn_d);
int sw_solver_clouds(rrnn_ctx_t* ctx, const rrnn_kdist_t* kd, int nlay, int ncol, int top_at_1, const float* inc_flux_d, const float* inc_flux_dif_d,
                     const float* tau_d, const float* ssa_d, const float* cld_rows_d, const float* mu0_d, const float* alb_dir_d,
                     const float* alb_dif_d, float* flux_up_d, float* flux_dn_d, float* flux_dir_d);
}
bool rrnn_gas_optics_tc_can(const rrnn_ctx_t* ctx, int mode, const rrnn_kdist_t* kd, const rrnn_model_t* const* models, int nmodels, int nlay,
                            bool compact);  // gas_optics_tc.cu

namespace rrnn {

__global__ void bcast_col_kernel(int ngpt, int ncol, const float* __restrict__ percol, float* __restrict__ out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < (size_t)ngpt * ncol) out[i] = percol[i / ngpt];
}

// SW boundary conditions of the RFMIP driver (rrtmgp_rfmip_sw.F90:409-434): TSI renormalisation of toa_flux,
// per-g-point albedo, mu0 := 1 for night columns.
__global__ void sw_bc_kernel(int ngpt, int ncol, const float* __restrict__ solar, float def_tsi, const float* __restrict__ tsi,
                             const float* __restrict__ alb, const float* __restrict__ mu0_in, float* __restrict__ toa,
                             float* __restrict__ alb_gpt, float* __restrict__ mu0_eff) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)ngpt * ncol) return;
  const size_t c = i / ngpt;
  const int g = (int)(i - c * ngpt);
  float t = solar[g];
  if (tsi) t = t * tsi[c] / def_tsi;
  toa[i] = t;
  alb_gpt[i] = alb[c];
  if (g == 0) mu0_eff[c] = (mu0_in[c] > 0.0f) ? mu0_in[c] : 1.0f;
}

// zero flux_up / flux_dn of night columns (rrtmgp_rfmip_sw.F90:458-463)
__global__ void sw_night_kernel(int nlev, int ncol, const float* __restrict__ mu0_in, float* __restrict__ fup, float* __restrict__ fdn) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)nlev * ncol) return;
  if (!(mu0_in[i / nlev] > 0.0f)) { fup[i] = 0.0f; fdn[i] = 0.0f; }
}

static int ensure_ws(rrnn_ctx_t* ctx, size_t bytes) {
  if (ctx->ws_bytes >= bytes) return 0;
  if (ctx->ws) { RRNN_CUDA(cudaStreamSynchronize(ctx->stream)); RRNN_CUDA(cudaFree(ctx->ws)); ctx->ws = nullptr; ctx->ws_bytes = 0; }
  RRNN_CUDA(cudaMalloc(&ctx->ws, bytes));
  ctx->ws_bytes = bytes;
  return 0;
}

static size_t align256(size_t n) { return (n + 255) & ~(size_t)255; }

static size_t lw_ws_bytes(int G, int L, int nc, bool compact) {
  if (compact)  // tau, pfrac, planck_lay, planck_lev, sfc_source, sfc_source_Jac, sfc_emis_gpt
    return 4 * (align256((size_t)nc * L * G) * 2 + align256((size_t)nc * L * 16) + align256((size_t)nc * (L + 1) * 16) + 3 * align256((size_t)nc * G));
  return 4 * (align256((size_t)nc * L * G) * 2 + align256((size_t)nc * (L + 1) * G) + 3 * align256((size_t)nc * G));
}

// rrnn_lw_fluxes keeps the sources factored between its two kernels when both take that form (gas_optics_tc.cu,
// rte_solvers_v5.cu); the fluxes are bit-identical either way (tested).
static bool lw_compact(const rrnn_ctx_t* ctx, const rrnn_kdist_t* kd, const rrnn_model_t* const* models, int nmodels, int L) {
  if (!(ctx->lw_compact_source && ctx->solver_variant == 0 && lw_v5_supports(kd->ngpt, L))) return false;
  if (kd->ngpt & 1) return false;
  for (int g = 0; g + 1 < kd->ngpt; g += 2)   // the packed solver looks a band value up once per pair of g-points
    if (kd->gpt2band[g] != kd->gpt2band[g + 1]) return false;
  return rrnn_gas_optics_tc_can(ctx, 0, kd, models, nmodels, L, true);
}
static size_t sw_ws_bytes(int G, int L, int nc) {
  return 4 * (align256((size_t)nc * L * G) * 2 + 2 * align256((size_t)nc * G) + align256((size_t)nc));
}

// all-sky: the chunk's cloud physical properties on the device and where their optical properties come from
struct CloudIn {
  const rrnn_cloud_lut_t* lut = nullptr;   // LUT or Pade handle (rrnn_cloud_lut_create / rrnn_cloud_pade_create)
  const float *clwp = nullptr, *ciwp = nullptr, *reliq = nullptr, *reice = nullptr;   // (nlay,ncol)
};
// by-band cloud optical properties (1 LW / 3 SW arrays of (16,nlay,ncol) at most) + the table rows the solver reads
static size_t cld_ws_bytes(int L, int nc, bool sw) {
  return 4 * ((sw ? 3 : 1) * align256((size_t)nc * L * 16) + align256((size_t)nc * L * (sw ? 48 : 16)));
}

static int pick_chunk(rrnn_ctx_t* ctx, int ncol, size_t bytes_per_col) {
  if (ctx->chunk_columns > 0) return std::min(ncol, ctx->chunk_columns);
  // default: at most ~12 GiB of optical-property workspace and never more than half of what the device has free right
  // now (the library may be embedded in a host model or share the GPU with a framework); at least enough columns to
  // fill the GPU
  size_t budget = (size_t)12 << 30, free_b = 0, total_b = 0;
  if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess) budget = std::min(budget, (free_b + ctx->ws_bytes) / 2);
  else cudaGetLastError();
  long long c = (long long)(budget / bytes_per_col);
  c = std::max<long long>(c, 256);
  c = std::min<long long>(c, 32768);
  return (int)std::min<long long>(c, ncol);
}

// Workspace for `chunk` columns; when the allocation fails the chunk is halved and tried again (down to 256 columns):
// a busy device gets a smaller workspace and more passes instead of an error.  bytes_of(chunk) -> workspace bytes.
static int ensure_ws_retry(rrnn_ctx_t* ctx, int& chunk, const std::function<size_t(int)>& bytes_of) {
  for (;;) {
    const size_t need = bytes_of(chunk);
    if (ctx->ws_bytes >= need) return 0;
    if (ctx->ws) { RRNN_CUDA(cudaStreamSynchronize(ctx->stream)); RRNN_CUDA(cudaFree(ctx->ws)); ctx->ws = nullptr; ctx->ws_bytes = 0; }
    const cudaError_t e = cudaMalloc(&ctx->ws, need);
    if (e == cudaSuccess) { ctx->ws_bytes = need; return 0; }
    cudaGetLastError();
    ctx->ws = nullptr;
    if (ctx->chunk_columns > 0 || chunk <= 256)
      return fail(std::string("workspace allocation failed (") + cudaGetErrorString(e) + ") for " + std::to_string(chunk) + " columns per chunk");
    chunk = std::max(256, chunk / 2);
  }
}

static void offset_gases(const rrnn_gas_t* in, int ngas, size_t c0, int nlay, std::vector<rrnn_gas_t>& out) {
  out.assign(in, in + ngas);
  for (auto& g : out)
    if (g.ndims == 2 && g.conc) g.conc += c0 * nlay;
}

static int lw_chunk(rrnn_ctx_t* ctx, const rrnn_kdist_t* kd, const rrnn_model_t* const* models, int nmodels, int nc, int L,
                    int top_at_1, int nang, const float* play, const float* plev, const float* tlay, const float* tlev,
                    const float* tsfc, const float* emis, const rrnn_gas_t* gases, int ngas, float* fup, float* fdn, float* ws,
                    const CloudIn* cl = nullptr) {
  static const float gauss_Ds[4][4] = {{1.66f, 0.f, 0.f, 0.f},
                                       {1.18350343f, 2.81649655f, 0.f, 0.f},
                                       {1.09719858f, 1.69338507f, 4.70941630f, 0.f},
                                       {1.06056257f, 1.38282560f, 2.40148179f, 7.15513024f}};
  static const float gauss_wts[4][4] = {{0.5f, 0.f, 0.f, 0.f},
                                        {0.3180413817f, 0.1819586183f, 0.f, 0.f},
                                        {0.2009319137f, 0.2292411064f, 0.0698269799f, 0.f},
                                        {0.1355069134f, 0.2034645680f, 0.1298475476f, 0.0311809710f}};
  const int G = kd->ngpt;
  const size_t n = (size_t)G * nc;
  float* tau = ws;
  float* lay = tau + align256((size_t)nc * L * G);
  if (lw_compact(ctx, kd, models, nmodels, L)) {
    float* bl = lay + align256((size_t)nc * L * G);
    float* bv = bl + align256((size_t)nc * L * 16);
    float* ssrc = bv + align256((size_t)nc * (L + 1) * 16);
    float* sjac = ssrc + align256((size_t)nc * G);
    float* egpt = sjac + align256((size_t)nc * G);
    if (int rc = rrnn_gas_optics_lw_compact(ctx, kd, models, nmodels, nc, L, play, plev, tlay, tsfc, gases, ngas, tlev, tau, lay, bl, bv, ssrc, sjac))
      return rc;
    if (cl) {   // all-sky (rrtmgp_allsky.F90:369-400): cloud optics by band, increment folded into the solver
      float* ctau = egpt + align256((size_t)nc * G);
      float* rows = ctau + align256((size_t)nc * L * 16);
      if (int rf = cloud_rows_fused(ctx, cl->lut, nc, L, cl->clwp, cl->ciwp, cl->reliq, cl->reice, false, rows)) {
        if (rf > 0) return rf;
        if (int rc = rrnn_cloud_optics(ctx, cl->lut, nc, L, cl->clwp, cl->ciwp, cl->reliq, cl->reice, ctau, nullptr, nullptr)) return rc;
        if (int rc = cloud_rows_lw(ctx, (size_t)nc * L, kd->nbnd, ctau, rows)) return rc;
      }
      NvtxRange nvtx_rte("rte_lw");
      bcast_col_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(G, nc, emis, egpt);
      RRNN_LAUNCH_CHECK(ctx);
      const int rc = lw_solver_clouds(ctx, kd, L, nc, top_at_1, nang, gauss_Ds[nang - 1], gauss_wts[nang - 1], nullptr, tau, lay, nullptr, bl, bv,
                                      egpt, ssrc, rows, fup, fdn);
      return rc < 0 ? fail("rrnn_lw_fluxes_allsky: shape not taken by the packed solver") : rc;
    }
    NvtxRange nvtx_rte("rte_lw");
    bcast_col_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(G, nc, emis, egpt);
    RRNN_LAUNCH_CHECK(ctx);
    return rrnn_lw_solver_noscat_compact(ctx, kd, L, nc, top_at_1, nang, gauss_Ds[nang - 1], gauss_wts[nang - 1], tau, lay, bl, bv, egpt,
                                         ssrc, fup, fdn);
  }
  float* lev = lay + align256((size_t)nc * L * G);
  float* ssrc = lev + align256((size_t)nc * (L + 1) * G);
  float* sjac = ssrc + align256((size_t)nc * G);
  float* egpt = sjac + align256((size_t)nc * G);
  if (int rc = rrnn_gas_optics_lw(ctx, kd, models, nmodels, nc, L, play, plev, tlay, tsfc, gases, ngas, tlev, tau, lay, lev, ssrc, sjac))
    return rc;
  if (cl) {
    float* ctau = egpt + align256((size_t)nc * G);
    float* rows = ctau + align256((size_t)nc * L * 16);
    if (int rc = rrnn_cloud_optics(ctx, cl->lut, nc, L, cl->clwp, cl->ciwp, cl->reliq, cl->reice, ctau, nullptr, nullptr)) return rc;
    if (int rc = cloud_rows_lw(ctx, (size_t)nc * L, kd->nbnd, ctau, rows)) return rc;
    NvtxRange nvtx_rte("rte_lw");
    bcast_col_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(G, nc, emis, egpt);
    RRNN_LAUNCH_CHECK(ctx);
    int rc = lw_solver_clouds(ctx, kd, L, nc, top_at_1, nang, gauss_Ds[nang - 1], gauss_wts[nang - 1], nullptr, tau, lay, lev, nullptr, nullptr,
                              egpt, ssrc, rows, fup, fdn);
    if (rc >= 0) return rc;
    // shapes the packed solver does not take: the increment as a pass over tau, then the plain solver
    if ((rc = rrnn_increment_1scl_bybnd(ctx, kd, L, nc, tau, ctau))) return rc;
    return rrnn_lw_solver_noscat(ctx, G, L, nc, top_at_1, nang, gauss_Ds[nang - 1], gauss_wts[nang - 1], nullptr, tau, lay, lev, egpt, ssrc, fup, fdn);
  }
  NvtxRange nvtx_rte("rte_lw");
  bcast_col_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(G, nc, emis, egpt);
  RRNN_LAUNCH_CHECK(ctx);
  return rrnn_lw_solver_noscat(ctx, G, L, nc, top_at_1, nang, gauss_Ds[nang - 1], gauss_wts[nang - 1], nullptr, tau, lay, lev,
                               egpt, ssrc, fup, fdn);
}

static int sw_chunk(rrnn_ctx_t* ctx, const rrnn_kdist_t* kd, const rrnn_model_t* const* models, int nc, int L, int top_at_1,
                    const float* play, const float* plev, const float* tlay, const float* mu0, const float* alb,
                    const float* tsi, const rrnn_gas_t* gases, int ngas, float* fup, float* fdn, float* fdir, float* ws,
                    const CloudIn* cl = nullptr) {
  const int G = kd->ngpt;
  float* tau = ws;
  float* ssa = tau + align256((size_t)nc * L * G);
  float* toa = ssa + align256((size_t)nc * L * G);
  float* agpt = toa + align256((size_t)nc * G);
  float* mu0e = agpt + align256((size_t)nc * G);
  // g is identically zero on the NN path (mo_gas_optics_rrtmgp.F90:560-567): not materialised, the solver is told so
  if (int rc = rrnn_gas_optics_sw(ctx, kd, models, nc, L, play, plev, tlay, gases, ngas, tau, ssa, nullptr, nullptr)) return rc;
  float def_tsi = 0.0f;
  for (float v : kd->solar_source) def_tsi += v;  // :409-416, same for every column
  const size_t n = (size_t)G * nc;
  NvtxRange nvtx_rte("rte_sw");
  sw_bc_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(G, nc, kd->d_solar_source, def_tsi, tsi, alb, mu0, toa, agpt, mu0e);
  RRNN_LAUNCH_CHECK(ctx);
  if (cl) {   // all-sky (rrtmgp_allsky.F90:369-431): cloud optics by band, delta-scaling by band, increment folded into the solver
    float* ctau = mu0e + align256((size_t)nc);
    float* cssa = ctau + align256((size_t)nc * L * 16);
    float* cg = cssa + align256((size_t)nc * L * 16);
    float* rows = cg + align256((size_t)nc * L * 16);
    if (int rf = cloud_rows_fused(ctx, cl->lut, nc, L, cl->clwp, cl->ciwp, cl->reliq, cl->reice, true, rows)) {
      if (rf > 0) return rf;
      if (int rc = rrnn_cloud_optics(ctx, cl->lut, nc, L, cl->clwp, cl->ciwp, cl->reliq, cl->reice, ctau, cssa, cg)) return rc;
      if (int rc = rrnn_delta_scale_2str(ctx, (size_t)nc * L * kd->nbnd, ctau, cssa, cg)) return rc;
      if (int rc = cloud_rows_sw(ctx, (size_t)nc * L, kd->nbnd, ctau, cssa, cg, rows)) return rc;
    }
    const int rc = sw_solver_clouds(ctx, kd, L, nc, top_at_1, toa, nullptr, tau, ssa, rows, mu0e, agpt, agpt, fup, fdn, fdir);
    if (rc) return rc < 0 ? fail("rrnn_sw_fluxes_allsky: shape not taken by the packed solver") : rc;
  } else if (int rc = rrnn_sw_solver_2stream(ctx, G, L, nc, top_at_1, toa, nullptr, tau, ssa, nullptr, mu0e, agpt, agpt, fup, fdn, fdir)) {
    return rc;
  }
  const size_t nf = (size_t)(L + 1) * nc;
  sw_night_kernel<<<(unsigned)((nf + 255) / 256), 256, 0, ctx->stream>>>(L + 1, nc, mu0, fup, fdn);
  RRNN_LAUNCH_CHECK(ctx);
  return 0;
}

}  // namespace rrnn

using namespace rrnn;

static int lw_fluxes_dev(rrnn_ctx_t* ctx, const rrnn_kdist_t* kd, const rrnn_model_t* const* models, int nmodels, int ncol, int nlay,
                         int top_at_1, int n_gauss_angles, const float* play_d, const float* plev_d, const float* tlay_d, const float* tlev_d,
                         const float* tsfc_d, const float* sfc_emis_d, const rrnn_gas_t* gases, int ngas, float* flux_up_d, float* flux_dn_d,
                         const CloudIn* cl) {
  RRNN_CHECK(ctx && kd && models, "rrnn_lw_fluxes: null handle");
  RRNN_CHECK(n_gauss_angles >= 1 && n_gauss_angles <= 4, "rte_lw: n_gauss_angles must be in 1..4");
  if (ncol <= 0) return 0;
  RRNN_CUDA(cudaSetDevice(ctx->device));
  const int G = kd->ngpt, L = nlay;
  const bool compact = lw_compact(ctx, kd, models, nmodels, L);
  auto bytes_of = [&](int c) { return lw_ws_bytes(G, L, c, compact) + (cl ? cld_ws_bytes(L, c, false) : 0); };
  int chunk = pick_chunk(ctx, ncol, bytes_of(1));
  if (int rc = ensure_ws_retry(ctx, chunk, bytes_of)) return rc;
  std::vector<rrnn_gas_t> gs;
  for (int c0 = 0; c0 < ncol; c0 += chunk) {
    const int nc = std::min(chunk, ncol - c0);
    offset_gases(gases, ngas, (size_t)c0, L, gs);
    CloudIn c1;
    if (cl) { c1 = *cl; c1.clwp += (size_t)c0 * L; c1.ciwp += (size_t)c0 * L; c1.reliq += (size_t)c0 * L; c1.reice += (size_t)c0 * L; }
    if (int rc = lw_chunk(ctx, kd, models, nmodels, nc, L, top_at_1, n_gauss_angles, play_d + (size_t)c0 * L,
                          plev_d + (size_t)c0 * (L + 1), tlay_d + (size_t)c0 * L, tlev_d ? tlev_d + (size_t)c0 * (L + 1) : nullptr,
                          tsfc_d + c0, sfc_emis_d + c0, gs.data(), ngas, flux_up_d + (size_t)c0 * (L + 1),
                          flux_dn_d + (size_t)c0 * (L + 1), (float*)ctx->ws, cl ? &c1 : nullptr))
      return rc;
  }
  return 0;
}

extern "C" int rrnn_lw_fluxes(rrnn_ctx_t* ctx, const rrnn_kdist_t* kd, const rrnn_model_t* const* models, int nmodels,
                              int ncol, int nlay, int top_at_1, int n_gauss_angles, const float* play_d, const float* plev_d,
                              const float* tlay_d, const float* tlev_d, const float* tsfc_d, const float* sfc_emis_d,
                              const rrnn_gas_t* gases, int ngas, float* flux_up_d, float* flux_dn_d) {
  rrnn::NvtxRange nvtx_("clear_sky_total (LW)");
  return lw_fluxes_dev(ctx, kd, models, nmodels, ncol, nlay, top_at_1, n_gauss_angles, play_d, plev_d, tlay_d, tlev_d, tsfc_d, sfc_emis_d, gases,
                       ngas, flux_up_d, flux_dn_d, nullptr);
}

extern "C" int rrnn_lw_fluxes_allsky(rrnn_ctx_t* ctx, const rrnn_kdist_t* kd, const rrnn_model_t* const* models, int nmodels,
                                     const rrnn_cloud_lut_t* cloud_optics, int ncol, int nlay, int top_at_1, int n_gauss_angles,
                                     const float* play_d, const float* plev_d, const float* tlay_d, const float* tlev_d, const float* tsfc_d,
                                     const float* sfc_emis_d, const rrnn_gas_t* gases, int ngas, const float* clwp_d, const float* ciwp_d,
                                     const float* reliq_d, const float* reice_d, float* flux_up_d, float* flux_dn_d) {
  rrnn::NvtxRange nvtx_("cloudy_sky_total");
  RRNN_CHECK(cloud_optics && clwp_d && ciwp_d && reliq_d && reice_d, "cloud optics: no data has been initialized");
  CloudIn cl;
  cl.lut = cloud_optics; cl.clwp = clwp_d; cl.ciwp = ciwp_d; cl.reliq = reliq_d; cl.reice = reice_d;
  return lw_fluxes_dev(ctx, kd, models, nmodels, ncol, nlay, top_at_1, n_gauss_angles, play_d, plev_d, tlay_d, tlev_d, tsfc_d, sfc_emis_d, gases,
                       ngas, flux_up_d, flux_dn_d, &cl);
}

static int sw_fluxes_dev(rrnn_ctx_t* ctx, const rrnn_kdist_t* kd, const rrnn_model_t* const* models, int ncol, int nlay, int top_at_1,
                         const float* play_d, const float* plev_d, const float* tlay_d, const float* mu0_d, const float* sfc_alb_d,
                         const float* tsi_d, const rrnn_gas_t* gases, int ngas, float* flux_up_d, float* flux_dn_d, float* flux_dn_dir_d,
                         const CloudIn* cl) {
  RRNN_CHECK(ctx && kd && models, "rrnn_sw_fluxes: null handle");
  RRNN_CHECK(kd->d_solar_source, "rrnn_sw_fluxes: k-distribution has no solar source");
  if (ncol <= 0) return 0;
  RRNN_CUDA(cudaSetDevice(ctx->device));
  const int G = kd->ngpt, L = nlay;
  auto bytes_of = [&](int c) { return sw_ws_bytes(G, L, c) + (cl ? cld_ws_bytes(L, c, true) : 0); };
  int chunk = pick_chunk(ctx, ncol, bytes_of(1));
  if (int rc = ensure_ws_retry(ctx, chunk, bytes_of)) return rc;
  std::vector<rrnn_gas_t> gs;
  for (int c0 = 0; c0 < ncol; c0 += chunk) {
    const int nc = std::min(chunk, ncol - c0);
    offset_gases(gases, ngas, (size_t)c0, L, gs);
    CloudIn c1;
    if (cl) { c1 = *cl; c1.clwp += (size_t)c0 * L; c1.ciwp += (size_t)c0 * L; c1.reliq += (size_t)c0 * L; c1.reice += (size_t)c0 * L; }
    if (int rc = sw_chunk(ctx, kd, models, nc, L, top_at_1, play_d + (size_t)c0 * L, plev_d + (size_t)c0 * (L + 1),
                          tlay_d + (size_t)c0 * L, mu0_d + c0, sfc_alb_d + c0, tsi_d ? tsi_d + c0 : nullptr, gs.data(), ngas,
                          flux_up_d + (size_t)c0 * (L + 1), flux_dn_d + (size_t)c0 * (L + 1),
                          flux_dn_dir_d + (size_t)c0 * (L + 1), (float*)ctx->ws, cl ? &c1 : nullptr))
      return rc;
  }
  return 0;
}

extern "C" int rrnn_sw_fluxes(rrnn_ctx_t* ctx, const rrnn_kdist_t* kd, const rrnn_model_t* const* models, int ncol, int nlay,
                              int top_at_1, const float* play_d, const float* plev_d, const float* tlay_d, const float* mu0_d,
                              const float* sfc_alb_d, const float* tsi_d, const rrnn_gas_t* gases, int ngas,
                              float* flux_up_d, float* flux_dn_d, float* flux_dn_dir_d) {
  rrnn::NvtxRange nvtx_("clear_sky_total (SW)");
  return sw_fluxes_dev(ctx, kd, models, ncol, nlay, top_at_1, play_d, plev_d, tlay_d, mu0_d, sfc_alb_d, tsi_d, gases, ngas, flux_up_d, flux_dn_d,
                       flux_dn_dir_d, nullptr);
}

extern "C" int rrnn_sw_fluxes_allsky(rrnn_ctx_t* ctx, const rrnn_kdist_t* kd, const rrnn_model_t* const* models,
                                     const rrnn_cloud_lut_t* cloud_optics, int ncol, int nlay, int top_at_1, const float* play_d,
                                     const float* plev_d, const float* tlay_d, const float* mu0_d, const float* sfc_alb_d, const float* tsi_d,
                                     const rrnn_gas_t* gases, int ngas, const float* clwp_d, const float* ciwp_d, const float* reliq_d,
                                     const float* reice_d, float* flux_up_d, float* flux_dn_d, float* flux_dn_dir_d) {
  rrnn::NvtxRange nvtx_("cloudy_sky_total");
  RRNN_CHECK(cloud_optics && clwp_d && ciwp_d && reliq_d && reice_d, "cloud optics: no data has been initialized");
  CloudIn cl;
  cl.lut = cloud_optics; cl.clwp = clwp_d; cl.ciwp = ciwp_d; cl.reliq = reliq_d; cl.reice = reice_d;
  return sw_fluxes_dev(ctx, kd, models, ncol, nlay, top_at_1, play_d, plev_d, tlay_d, mu0_d, sfc_alb_d, tsi_d, gases, ngas, flux_up_d, flux_dn_d,
                       flux_dn_dir_d, &cl);
}

// ---- host-buffer drivers ---------------------------------------------------------------------------
namespace {

struct HostField {
  const float* host;   // host source (per column stride `per_col`), null = absent
  size_t per_col;      // floats per column
  float* dev[2];       // device staging (double-buffered)
  bool pageable;       // the caller's memory is not page-locked: it goes through the pinned bounce ring
  float* pin[2];       // this field's slots in the bounce ring
};

// Is this host pointer page-locked (cudaHostAlloc / cudaHostRegister)?  Anything else -- malloc, Fortran allocate, numpy --
// is pageable, and a cudaMemcpyAsync from it would be staged by the driver synchronously and without overlap.
bool is_pinned(const void* p) {
  cudaPointerAttributes a{};
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeHost;
}

// memcpy split over a few host threads (one thread moves ~10 GB/s; the PCIe / C2C link wants more)
void parallel_copy(void* dst, const void* src, size_t bytes, int nthreads) {
  if (bytes == 0) return;
  const size_t min_per_thread = (size_t)4 << 20;
  int nt = (int)std::min<size_t>((size_t)std::max(nthreads, 1), (bytes + min_per_thread - 1) / min_per_thread);
  if (nt <= 1) { memcpy(dst, src, bytes); return; }
  std::vector<std::thread> th;
  th.reserve(nt - 1);
  const size_t per = ((bytes / nt) + 4095) & ~(size_t)4095;
  for (int t = 1; t < nt; ++t) {
    const size_t o = std::min(bytes, per * t), e = std::min(bytes, per * (t + 1));
    if (e > o) th.emplace_back([=]() { memcpy((char*)dst + o, (const char*)src + o, e - o); });
  }
  memcpy(dst, src, std::min(bytes, per));
  for (auto& x : th) x.join();
}

int ensure_pinned(rrnn_ctx_t* ctx, size_t bytes) {
  if (ctx->pinned_bytes >= bytes) return 0;
  if (ctx->pinned) { RRNN_CUDA(cudaFreeHost(ctx->pinned)); ctx->pinned = nullptr; ctx->pinned_bytes = 0; }
  RRNN_CUDA(cudaHostAlloc(&ctx->pinned, bytes, cudaHostAllocDefault));
  ctx->pinned_bytes = bytes;
  return 0;
}

}  // namespace

// One pass over all columns with HOST buffers.  Three streams overlap the H2D copy of chunk k+1, the kernels of chunk k
// and the D2H copy of chunk k-1.  Page-locked caller memory is copied from / to directly; pageable caller memory (what a
// Fortran or C host normally passes) is staged through a persistent pinned bounce ring of the context, two slots per
// array, filled and drained by a few host threads while the GPU works on the neighbouring chunks.
static int run_host_pipeline(rrnn_ctx_t* ctx, bool lw, const rrnn_kdist_t* kd, const rrnn_model_t* const* models, int nmodels,
                             int ncol, int L, int top_at_1, int nang, std::vector<HostField>& fields,
                             const rrnn_gas_t* gases, int ngas, float* const* out_host, int nout,
                             const rrnn_cloud_lut_t* cloud_optics = nullptr) {   // all-sky: fields[6..9] = clwp, ciwp, reliq, reice
  const int G = kd->ngpt;
  // 2-D gas fields are appended to the staged fields
  std::vector<int> gas_field(ngas, -1);
  for (int g = 0; g < ngas; ++g)
    if (gases[g].ndims == 2) { gas_field[g] = (int)fields.size(); fields.push_back({gases[g].conc, (size_t)L, {nullptr, nullptr}, false, {nullptr, nullptr}}); }
  // 1-D gas profiles: uploaded once
  std::vector<float*> gas1d(ngas, nullptr);
  size_t in_per_col = 0, in_pageable_per_col = 0;
  for (auto& f : fields)
    if (f.host) {
      in_per_col += f.per_col;
      f.pageable = !is_pinned(f.host);
      if (f.pageable) in_pageable_per_col += f.per_col;
    }
  bool out_pageable[3] = {false, false, false};
  int nout_pageable = 0;
  for (int o = 0; o < nout; ++o) { out_pageable[o] = !is_pinned(out_host[o]); nout_pageable += out_pageable[o] ? 1 : 0; }
  const size_t out_per_col = (size_t)nout * (L + 1);
  const bool compact = lw && lw_compact(ctx, kd, models, nmodels, L);
  const size_t opt_per_col = (lw ? lw_ws_bytes(G, L, 1, compact) : sw_ws_bytes(G, L, 1)) + (cloud_optics ? cld_ws_bytes(L, 1, !lw) : 0);
  int chunk = pick_chunk(ctx, ncol, opt_per_col + 8 * (in_per_col + out_per_col));
  size_t gas1d_floats = 0;
  for (int g = 0; g < ngas; ++g) if (gases[g].ndims == 1) gas1d_floats += align256((size_t)L);
  auto stage_floats_of = [&](int c) {
    size_t n = 0;
    for (auto& f : fields) if (f.host) n += 2 * align256(f.per_col * c);
    n += 2 * (size_t)nout * align256((size_t)(L + 1) * c);
    return n;
  };
  auto opt_bytes_of = [&](int c) { return (lw ? lw_ws_bytes(G, L, c, compact) : sw_ws_bytes(G, L, c)) + (cloud_optics ? cld_ws_bytes(L, c, !lw) : 0); };
  if (int rc = ensure_ws_retry(ctx, chunk, [&](int c) { return opt_bytes_of(c) + 4 * (stage_floats_of(c) + gas1d_floats); })) return rc;
  const size_t opt_bytes = opt_bytes_of(chunk);
  float* p = (float*)((char*)ctx->ws + opt_bytes);
  for (auto& f : fields) if (f.host) for (int b = 0; b < 2; ++b) { f.dev[b] = p; p += align256(f.per_col * chunk); }
  float* outd[2][3] = {};
  for (int b = 0; b < 2; ++b) for (int o = 0; o < nout; ++o) { outd[b][o] = p; p += align256((size_t)(L + 1) * chunk); }
  // the pinned bounce ring (only what pageable arrays need)
  float* outpin[2][3] = {};
  {
    const size_t pin_floats = 2 * (align256(in_pageable_per_col * chunk) + (size_t)fields.size() * 256) +
                              2 * (size_t)nout_pageable * align256((size_t)(L + 1) * chunk);
    if (in_pageable_per_col || nout_pageable) {
      if (int rc = ensure_pinned(ctx, 4 * pin_floats)) return rc;
      float* q = (float*)ctx->pinned;
      for (auto& f : fields) if (f.host && f.pageable) for (int b = 0; b < 2; ++b) { f.pin[b] = q; q += align256(f.per_col * chunk); }
      for (int b = 0; b < 2; ++b) for (int o = 0; o < nout; ++o) if (out_pageable[o]) { outpin[b][o] = q; q += align256((size_t)(L + 1) * chunk); }
    }
  }
  const int nthreads = ctx->host_copy_threads > 0 ? ctx->host_copy_threads : (int)std::min(8u, std::max(1u, std::thread::hardware_concurrency()));
  cudaStream_t s_in = ctx->copy_stream, s_cmp = ctx->stream, s_out = ctx->out_stream;
  // every exit path leaves no copy in flight on the caller's (or the ring's) memory
  auto drain = [&]() { cudaStreamSynchronize(s_in); cudaStreamSynchronize(s_out); cudaStreamSynchronize(s_cmp); };
#define RRNN_PIPE(call)                                                                  \
  do {                                                                                   \
    cudaError_t _e = (call);                                                             \
    if (_e != cudaSuccess) { drain(); return fail(std::string(#call) + ": " + cudaGetErrorString(_e)); } \
  } while (0)
  for (int g = 0; g < ngas; ++g)
    if (gases[g].ndims == 1) {
      gas1d[g] = p; p += align256((size_t)L);
      RRNN_PIPE(cudaMemcpyAsync(gas1d[g], gases[g].conc, L * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    }
  // events: 0/1 = H2D done (per buffer), 2/3 = compute done (per buffer), 4/5 = D2H done (per buffer)
  std::vector<rrnn_gas_t> gs(gases, gases + ngas);
  auto unstage = [&](int kk) {  // pinned ring -> the caller's pageable flux arrays, chunk kk
    if (!nout_pageable) return cudaSuccess;
    const int bb = kk & 1;
    const cudaError_t e = cudaEventSynchronize(ctx->ev[4 + bb]);
    if (e != cudaSuccess) return e;
    const size_t cc0 = (size_t)kk * chunk;
    const int ncc = (int)std::min<size_t>(chunk, (size_t)ncol - cc0);
    for (int o = 0; o < nout; ++o)
      if (out_pageable[o]) parallel_copy(out_host[o] + cc0 * (L + 1), outpin[bb][o], (size_t)(L + 1) * ncc * sizeof(float), nthreads);
    return cudaSuccess;
  };
  int k = 0;
  for (int c0 = 0; c0 < ncol; c0 += chunk, ++k) {
    const int nc = std::min(chunk, ncol - c0);
    const int b = k & 1;
    // inputs of chunk k may overwrite staging buffer b once the kernels of chunk k-2 are done; ring slot b once its H2D is
    if (k >= 2) {
      RRNN_PIPE(cudaStreamWaitEvent(s_in, ctx->ev[2 + b], 0));
      if (in_pageable_per_col) RRNN_PIPE(cudaEventSynchronize(ctx->ev[b]));
    }
    for (auto& f : fields)
      if (f.host) {
        const float* src = f.host + (size_t)c0 * f.per_col;
        const size_t bytes = f.per_col * nc * sizeof(float);
        if (f.pageable) { parallel_copy(f.pin[b], src, bytes, nthreads); src = f.pin[b]; }
        RRNN_PIPE(cudaMemcpyAsync(f.dev[b], src, bytes, cudaMemcpyHostToDevice, s_in));
      }
    RRNN_PIPE(cudaEventRecord(ctx->ev[b], s_in));
    RRNN_PIPE(cudaStreamWaitEvent(s_cmp, ctx->ev[b], 0));
    // flux staging buffer b must have been drained (D2H of chunk k-2 on s_out)
    if (k >= 2) RRNN_PIPE(cudaStreamWaitEvent(s_cmp, ctx->ev[4 + b], 0));
    for (int g = 0; g < ngas; ++g) {
      if (gas_field[g] >= 0) gs[g].conc = fields[gas_field[g]].dev[b];
      else if (gases[g].ndims == 1) gs[g].conc = gas1d[g];
    }
    int rc;
    CloudIn cl;
    if (cloud_optics) { cl.lut = cloud_optics; cl.clwp = fields[6].dev[b]; cl.ciwp = fields[7].dev[b]; cl.reliq = fields[8].dev[b]; cl.reice = fields[9].dev[b]; }
    if (lw) {
      rc = lw_chunk(ctx, kd, models, nmodels, nc, L, top_at_1, nang, fields[0].dev[b], fields[1].dev[b], fields[2].dev[b],
                    fields[3].host ? fields[3].dev[b] : nullptr, fields[4].dev[b], fields[5].dev[b], gs.data(), ngas, outd[b][0],
                    outd[b][1], (float*)ctx->ws, cloud_optics ? &cl : nullptr);
    } else {
      rc = sw_chunk(ctx, kd, models, nc, L, top_at_1, fields[0].dev[b], fields[1].dev[b], fields[2].dev[b], fields[3].dev[b],
                    fields[4].dev[b], fields[5].host ? fields[5].dev[b] : nullptr, gs.data(), ngas, outd[b][0], outd[b][1],
                    outd[b][2], (float*)ctx->ws, cloud_optics ? &cl : nullptr);
    }
    if (rc) { drain(); return rc; }
    RRNN_PIPE(cudaEventRecord(ctx->ev[2 + b], s_cmp));
    // D2H of this chunk's fluxes on its own stream (overlaps the next chunk's H2D and kernels).  Ring slot b was emptied
    // by unstage(k-2) one iteration ago.
    RRNN_PIPE(cudaStreamWaitEvent(s_out, ctx->ev[2 + b], 0));
    for (int o = 0; o < nout; ++o) {
      float* dst = out_pageable[o] ? outpin[b][o] : out_host[o] + (size_t)c0 * (L + 1);
      RRNN_PIPE(cudaMemcpyAsync(dst, outd[b][o], (size_t)(L + 1) * nc * sizeof(float), cudaMemcpyDeviceToHost, s_out));
    }
    RRNN_PIPE(cudaEventRecord(ctx->ev[4 + b], s_out));
    // while the GPU works on chunk k: hand chunk k-1's fluxes to the caller
    if (k >= 1) RRNN_PIPE(unstage(k - 1));
  }
  if (k >= 1) RRNN_PIPE(unstage(k - 1));
  RRNN_PIPE(cudaStreamSynchronize(s_in));
  RRNN_PIPE(cudaStreamSynchronize(s_out));
  RRNN_PIPE(cudaStreamSynchronize(s_cmp));
#undef RRNN_PIPE
  return 0;
}

extern "C" int rrnn_lw_fluxes_host(rrnn_ctx_t* ctx, const rrnn_kdist_t* kd, const rrnn_model_t* const* models, int nmodels,
                                   int ncol, int nlay, int top_at_1, int n_gauss_angles, const float* play, const float* plev,
                                   const float* tlay, const float* tlev, const float* tsfc, const float* sfc_emis,
                                   const rrnn_gas_t* gases, int ngas, float* flux_up, float* flux_dn) {
  rrnn::NvtxRange nvtx_("clear_sky_total (LW)");
  RRNN_CHECK(ctx && kd && models && play && plev && tlay && tsfc && sfc_emis && flux_up && flux_dn, "rrnn_lw_fluxes_host: null argument");
  RRNN_CHECK(n_gauss_angles >= 1 && n_gauss_angles <= 4, "rte_lw: n_gauss_angles must be in 1..4");
  if (ncol <= 0) return 0;
  RRNN_CUDA(cudaSetDevice(ctx->device));
  const size_t L = nlay;
  std::vector<HostField> f = {{play, L, {}, false, {}}, {plev, L + 1, {}, false, {}}, {tlay, L, {}, false, {}}, {tlev, L + 1, {}, false, {}}, {tsfc, 1, {}, false, {}}, {sfc_emis, 1, {}, false, {}}};
  float* outs[2] = {flux_up, flux_dn};
  return run_host_pipeline(ctx, true, kd, models, nmodels, ncol, nlay, top_at_1, n_gauss_angles, f, gases, ngas, outs, 2);
}

extern "C" int rrnn_sw_fluxes_host(rrnn_ctx_t* ctx, const rrnn_kdist_t* kd, const rrnn_model_t* const* models, int ncol,
                                   int nlay, int top_at_1, const float* play, const float* plev, const float* tlay,
                                   const float* mu0, const float* sfc_alb, const float* tsi, const rrnn_gas_t* gases,
                                   int ngas, float* flux_up, float* flux_dn, float* flux_dn_dir) {
  rrnn::NvtxRange nvtx_("clear_sky_total (SW)");
  RRNN_CHECK(ctx && kd && models && play && plev && tlay && mu0 && sfc_alb && flux_up && flux_dn && flux_dn_dir, "rrnn_sw_fluxes_host: null argument");
  RRNN_CHECK(kd->d_solar_source, "rrnn_sw_fluxes_host: k-distribution has no solar source");
  if (ncol <= 0) return 0;
  RRNN_CUDA(cudaSetDevice(ctx->device));
  const size_t L = nlay;
  std::vector<HostField> f = {{play, L, {}, false, {}}, {plev, L + 1, {}, false, {}}, {tlay, L, {}, false, {}}, {mu0, 1, {}, false, {}}, {sfc_alb, 1, {}, false, {}}, {tsi, 1, {}, false, {}}};
  float* outs[3] = {flux_up, flux_dn, flux_dn_dir};
  return run_host_pipeline(ctx, false, kd, models, 2, ncol, nlay, top_at_1, 1, f, gases, ngas, outs, 3);
}

// The all-sky pass with HOST buffers (examples/all-sky/rrtmgp_allsky.F90:366-446 for all columns at once): the clear-sky pipeline
// above with four more staged fields -- cloud liquid / ice water path and effective radii, (nlay,ncol) -- cloud optics by band
// on the device and the increment folded into the solvers.
extern "C" int rrnn_lw_fluxes_allsky_host(rrnn_ctx_t* ctx, const rrnn_kdist_t* kd, const rrnn_model_t* const* models, int nmodels,
                                          const rrnn_cloud_lut_t* cloud_optics, int ncol, int nlay, int top_at_1, int n_gauss_angles,
                                          const float* play, const float* plev, const float* tlay, const float* tlev, const float* tsfc,
                                          const float* sfc_emis, const rrnn_gas_t* gases, int ngas, const float* clwp, const float* ciwp,
                                          const float* reliq, const float* reice, float* flux_up, float* flux_dn) {
  rrnn::NvtxRange nvtx_("cloudy_sky_total");
  RRNN_CHECK(ctx && kd && models && play && plev && tlay && tsfc && sfc_emis && flux_up && flux_dn, "rrnn_lw_fluxes_allsky_host: null argument");
  RRNN_CHECK(cloud_optics && clwp && ciwp && reliq && reice, "cloud optics: no data has been initialized");
  RRNN_CHECK(n_gauss_angles >= 1 && n_gauss_angles <= 4, "rte_lw: n_gauss_angles must be in 1..4");
  if (ncol <= 0) return 0;
  RRNN_CUDA(cudaSetDevice(ctx->device));
  const size_t L = nlay;
  std::vector<HostField> f = {{play, L, {}, false, {}}, {plev, L + 1, {}, false, {}}, {tlay, L, {}, false, {}}, {tlev, L + 1, {}, false, {}},
                              {tsfc, 1, {}, false, {}}, {sfc_emis, 1, {}, false, {}}, {clwp, L, {}, false, {}}, {ciwp, L, {}, false, {}},
                              {reliq, L, {}, false, {}}, {reice, L, {}, false, {}}};
  float* outs[2] = {flux_up, flux_dn};
  return run_host_pipeline(ctx, true, kd, models, nmodels, ncol, nlay, top_at_1, n_gauss_angles, f, gases, ngas, outs, 2, cloud_optics);
}

extern "C" int rrnn_sw_fluxes_allsky_host(rrnn_ctx_t* ctx, const rrnn_kdist_t* kd, const rrnn_model_t* const* models,
                                          const rrnn_cloud_lut_t* cloud_optics, int ncol, int nlay, int top_at_1, const float* play,
                                          const float* plev, const float* tlay, const float* mu0, const float* sfc_alb, const float* tsi,
                                          const rrnn_gas_t* gases, int ngas, const float* clwp, const float* ciwp, const float* reliq,
                                          const float* reice, float* flux_up, float* flux_dn, float* flux_dn_dir) {
  rrnn::NvtxRange nvtx_("cloudy_sky_total");
  RRNN_CHECK(ctx && kd && models && play && plev && tlay && mu0 && sfc_alb && flux_up && flux_dn && flux_dn_dir, "rrnn_sw_fluxes_allsky_host: null argument");
  RRNN_CHECK(cloud_optics && clwp && ciwp && reliq && reice, "cloud optics: no data has been initialized");
  RRNN_CHECK(kd->d_solar_source, "rrnn_sw_fluxes_allsky_host: k-distribution has no solar source");
  if (ncol <= 0) return 0;
  RRNN_CUDA(cudaSetDevice(ctx->device));
  const size_t L = nlay;
  std::vector<HostField> f = {{play, L, {}, false, {}}, {plev, L + 1, {}, false, {}}, {tlay, L, {}, false, {}}, {mu0, 1, {}, false, {}},
                              {sfc_alb, 1, {}, false, {}}, {tsi, 1, {}, false, {}}, {clwp, L, {}, false, {}}, {ciwp, L, {}, false, {}},
                              {reliq, L, {}, false, {}}, {reice, L, {}, false, {}}};
  float* outs[3] = {flux_up, flux_dn, flux_dn_dir};
  return run_host_pipeline(ctx, false, kd, models, 2, ncol, nlay, top_at_1, 1, f, gases, ngas, outs, 3, cloud_optics);
}
