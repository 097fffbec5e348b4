// Whole-path drivers: what one pass of the reference drivers' block loop does
// (examples/rfmip-clear-sky/rrtmgp_rfmip_lw.F90:368-446, rrtmgp_rfmip_sw.F90:356-465):
//   LW: gas_optics(neural_nets=) -> rte_lw          SW: gas_optics(neural_nets=) -> boundary conditions -> rte_sw
// for ALL columns of a call, in column chunks whose optical-property arrays live in the context's persistent
// workspace (no allocation in steady state).  The *_host variants take host pointers and overlap the H2D copy
// of chunk k+1 and the D2H copy of chunk k-1 with the kernels of chunk k on three streams.
#include "common.cuh"
#include <algorithm>

namespace rrnn {
bool lw_v5_supports(int G, int L);  // rte_solvers_v5.cu
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
  return ctx->lw_compact_source && ctx->solver_variant == 0 && lw_v5_supports(kd->ngpt, L) &&
         rrnn_gas_optics_tc_can(ctx, 0, kd, models, nmodels, L, true);
}
static size_t sw_ws_bytes(int G, int L, int nc) {
  return 4 * (align256((size_t)nc * L * G) * 2 + 2 * align256((size_t)nc * G) + align256((size_t)nc));
}

static int pick_chunk(rrnn_ctx_t* ctx, int ncol, size_t bytes_per_col) {
  if (ctx->chunk_columns > 0) return std::min(ncol, ctx->chunk_columns);
  // default: at most ~12 GiB of optical-property workspace, at least enough columns to fill the GPU
  const size_t budget = (size_t)12 << 30;
  long long c = (long long)(budget / bytes_per_col);
  c = std::max<long long>(c, 256);
  c = std::min<long long>(c, 32768);
  return (int)std::min<long long>(c, ncol);
}

static void offset_gases(const rrnn_gas_t* in, int ngas, size_t c0, int nlay, std::vector<rrnn_gas_t>& out) {
  out.assign(in, in + ngas);
  for (auto& g : out)
    if (g.ndims == 2 && g.conc) g.conc += c0 * nlay;
}

static int lw_chunk(rrnn_ctx_t* ctx, const rrnn_kdist_t* kd, const rrnn_model_t* const* models, int nmodels, int nc, int L,
                    int top_at_1, int nang, const float* play, const float* plev, const float* tlay, const float* tlev,
                    const float* tsfc, const float* emis, const rrnn_gas_t* gases, int ngas, float* fup, float* fdn, float* ws) {
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
  bcast_col_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(G, nc, emis, egpt);
  RRNN_LAUNCH_CHECK(ctx);
  return rrnn_lw_solver_noscat(ctx, G, L, nc, top_at_1, nang, gauss_Ds[nang - 1], gauss_wts[nang - 1], nullptr, tau, lay, lev,
                               egpt, ssrc, fup, fdn);
}

static int sw_chunk(rrnn_ctx_t* ctx, const rrnn_kdist_t* kd, const rrnn_model_t* const* models, int nc, int L, int top_at_1,
                    const float* play, const float* plev, const float* tlay, const float* mu0, const float* alb,
                    const float* tsi, const rrnn_gas_t* gases, int ngas, float* fup, float* fdn, float* fdir, float* ws) {
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
  sw_bc_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(G, nc, kd->d_solar_source, def_tsi, tsi, alb, mu0, toa, agpt, mu0e);
  RRNN_LAUNCH_CHECK(ctx);
  if (int rc = rrnn_sw_solver_2stream(ctx, G, L, nc, top_at_1, toa, nullptr, tau, ssa, nullptr, mu0e, agpt, agpt, fup, fdn, fdir)) return rc;
  const size_t nf = (size_t)(L + 1) * nc;
  sw_night_kernel<<<(unsigned)((nf + 255) / 256), 256, 0, ctx->stream>>>(L + 1, nc, mu0, fup, fdn);
  RRNN_LAUNCH_CHECK(ctx);
  return 0;
}

}  // namespace rrnn

using namespace rrnn;

extern "C" int rrnn_lw_fluxes(rrnn_ctx_t* ctx, const rrnn_kdist_t* kd, const rrnn_model_t* const* models, int nmodels,
                              int ncol, int nlay, int top_at_1, int n_gauss_angles, const float* play_d, const float* plev_d,
                              const float* tlay_d, const float* tlev_d, const float* tsfc_d, const float* sfc_emis_d,
                              const rrnn_gas_t* gases, int ngas, float* flux_up_d, float* flux_dn_d) {
  RRNN_CHECK(ctx && kd && models, "rrnn_lw_fluxes: null handle");
  RRNN_CHECK(n_gauss_angles >= 1 && n_gauss_angles <= 4, "rte_lw: n_gauss_angles must be in 1..4");
  if (ncol <= 0) return 0;
  RRNN_CUDA(cudaSetDevice(ctx->device));
  const int G = kd->ngpt, L = nlay;
  const bool compact = lw_compact(ctx, kd, models, nmodels, L);
  const int chunk = pick_chunk(ctx, ncol, lw_ws_bytes(G, L, 1, compact));
  if (int rc = ensure_ws(ctx, lw_ws_bytes(G, L, chunk, compact))) return rc;
  std::vector<rrnn_gas_t> gs;
  for (int c0 = 0; c0 < ncol; c0 += chunk) {
    const int nc = std::min(chunk, ncol - c0);
    offset_gases(gases, ngas, (size_t)c0, L, gs);
    if (int rc = lw_chunk(ctx, kd, models, nmodels, nc, L, top_at_1, n_gauss_angles, play_d + (size_t)c0 * L,
                          plev_d + (size_t)c0 * (L + 1), tlay_d + (size_t)c0 * L, tlev_d ? tlev_d + (size_t)c0 * (L + 1) : nullptr,
                          tsfc_d + c0, sfc_emis_d + c0, gs.data(), ngas, flux_up_d + (size_t)c0 * (L + 1),
                          flux_dn_d + (size_t)c0 * (L + 1), (float*)ctx->ws))
      return rc;
  }
  return 0;
}

extern "C" int rrnn_sw_fluxes(rrnn_ctx_t* ctx, const rrnn_kdist_t* kd, const rrnn_model_t* const* models, int ncol, int nlay,
                              int top_at_1, const float* play_d, const float* plev_d, const float* tlay_d, const float* mu0_d,
                              const float* sfc_alb_d, const float* tsi_d, const rrnn_gas_t* gases, int ngas,
                              float* flux_up_d, float* flux_dn_d, float* flux_dn_dir_d) {
  RRNN_CHECK(ctx && kd && models, "rrnn_sw_fluxes: null handle");
  RRNN_CHECK(kd->d_solar_source, "rrnn_sw_fluxes: k-distribution has no solar source");
  if (ncol <= 0) return 0;
  RRNN_CUDA(cudaSetDevice(ctx->device));
  const int G = kd->ngpt, L = nlay;
  const int chunk = pick_chunk(ctx, ncol, sw_ws_bytes(G, L, 1));
  if (int rc = ensure_ws(ctx, sw_ws_bytes(G, L, chunk))) return rc;
  std::vector<rrnn_gas_t> gs;
  for (int c0 = 0; c0 < ncol; c0 += chunk) {
    const int nc = std::min(chunk, ncol - c0);
    offset_gases(gases, ngas, (size_t)c0, L, gs);
    if (int rc = sw_chunk(ctx, kd, models, nc, L, top_at_1, play_d + (size_t)c0 * L, plev_d + (size_t)c0 * (L + 1),
                          tlay_d + (size_t)c0 * L, mu0_d + c0, sfc_alb_d + c0, tsi_d ? tsi_d + c0 : nullptr, gs.data(), ngas,
                          flux_up_d + (size_t)c0 * (L + 1), flux_dn_d + (size_t)c0 * (L + 1),
                          flux_dn_dir_d + (size_t)c0 * (L + 1), (float*)ctx->ws))
      return rc;
  }
  return 0;
}

// ---- host-buffer drivers ---------------------------------------------------------------------------
namespace {

struct HostField {
  const float* host;   // host source (per column stride `per_col`), null = absent
  size_t per_col;      // floats per column
  float* dev[2];       // device staging (double-buffered)
};

struct Staging {
  std::vector<HostField> in;
  float* out_dev[2][3];
  float* ws_opt;
  size_t total_bytes;
};

}  // namespace

static int run_host_pipeline(rrnn_ctx_t* ctx, bool lw, const rrnn_kdist_t* kd, const rrnn_model_t* const* models, int nmodels,
                             int ncol, int L, int top_at_1, int nang, std::vector<HostField>& fields,
                             const rrnn_gas_t* gases, int ngas, float* const* out_host, int nout) {
  const int G = kd->ngpt;
  // 2-D gas fields are appended to the staged fields
  std::vector<int> gas_field(ngas, -1);
  for (int g = 0; g < ngas; ++g)
    if (gases[g].ndims == 2) { gas_field[g] = (int)fields.size(); fields.push_back({gases[g].conc, (size_t)L, {nullptr, nullptr}}); }
  // 1-D gas profiles: uploaded once
  std::vector<float*> gas1d(ngas, nullptr);
  size_t in_per_col = 0;
  for (auto& f : fields) if (f.host) in_per_col += f.per_col;
  const size_t out_per_col = (size_t)nout * (L + 1);
  const bool compact = lw && lw_compact(ctx, kd, models, nmodels, L);
  const size_t opt_per_col = lw ? lw_ws_bytes(G, L, 1, compact) : sw_ws_bytes(G, L, 1);
  const int chunk = pick_chunk(ctx, ncol, opt_per_col + 8 * (in_per_col + out_per_col));
  const size_t opt_bytes = lw ? lw_ws_bytes(G, L, chunk, compact) : sw_ws_bytes(G, L, chunk);
  size_t stage_floats = 0;
  for (auto& f : fields) if (f.host) stage_floats += 2 * align256(f.per_col * chunk);
  stage_floats += 2 * (size_t)nout * align256((size_t)(L + 1) * chunk);
  size_t gas1d_floats = 0;
  for (int g = 0; g < ngas; ++g) if (gases[g].ndims == 1) gas1d_floats += align256((size_t)L);
  if (int rc = ensure_ws(ctx, opt_bytes + 4 * (stage_floats + gas1d_floats))) return rc;
  float* p = (float*)((char*)ctx->ws + opt_bytes);
  for (auto& f : fields) if (f.host) for (int b = 0; b < 2; ++b) { f.dev[b] = p; p += align256(f.per_col * chunk); }
  float* outd[2][3] = {};
  for (int b = 0; b < 2; ++b) for (int o = 0; o < nout; ++o) { outd[b][o] = p; p += align256((size_t)(L + 1) * chunk); }
  for (int g = 0; g < ngas; ++g)
    if (gases[g].ndims == 1) {
      gas1d[g] = p; p += align256((size_t)L);
      RRNN_CUDA(cudaMemcpyAsync(gas1d[g], gases[g].conc, L * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    }
  cudaStream_t s_in = ctx->copy_stream, s_cmp = ctx->stream, s_out = ctx->out_stream;
  // events: 0/1 = H2D done (per buffer), 2/3 = compute done (per buffer), 4/5 = D2H done (per buffer)
  std::vector<rrnn_gas_t> gs(gases, gases + ngas);
  int k = 0;
  for (int c0 = 0; c0 < ncol; c0 += chunk, ++k) {
    const int nc = std::min(chunk, ncol - c0);
    const int b = k & 1;
    // inputs of chunk k may overwrite staging buffer b once the kernels of chunk k-2 are done
    if (k >= 2) RRNN_CUDA(cudaStreamWaitEvent(s_in, ctx->ev[2 + b], 0));
    for (auto& f : fields)
      if (f.host)
        RRNN_CUDA(cudaMemcpyAsync(f.dev[b], f.host + (size_t)c0 * f.per_col, f.per_col * nc * sizeof(float), cudaMemcpyHostToDevice, s_in));
    RRNN_CUDA(cudaEventRecord(ctx->ev[b], s_in));
    RRNN_CUDA(cudaStreamWaitEvent(s_cmp, ctx->ev[b], 0));
    // flux staging buffer b must have been drained (D2H of chunk k-2 on s_out)
    if (k >= 2) RRNN_CUDA(cudaStreamWaitEvent(s_cmp, ctx->ev[4 + b], 0));
    for (int g = 0; g < ngas; ++g) {
      if (gas_field[g] >= 0) gs[g].conc = fields[gas_field[g]].dev[b];
      else if (gases[g].ndims == 1) gs[g].conc = gas1d[g];
    }
    int rc;
    if (lw) {
      rc = lw_chunk(ctx, kd, models, nmodels, nc, L, top_at_1, nang, fields[0].dev[b], fields[1].dev[b], fields[2].dev[b],
                    fields[3].host ? fields[3].dev[b] : nullptr, fields[4].dev[b], fields[5].dev[b], gs.data(), ngas, outd[b][0],
                    outd[b][1], (float*)ctx->ws);
    } else {
      rc = sw_chunk(ctx, kd, models, nc, L, top_at_1, fields[0].dev[b], fields[1].dev[b], fields[2].dev[b], fields[3].dev[b],
                    fields[4].dev[b], fields[5].host ? fields[5].dev[b] : nullptr, gs.data(), ngas, outd[b][0], outd[b][1],
                    outd[b][2], (float*)ctx->ws);
    }
    if (rc) return rc;
    RRNN_CUDA(cudaEventRecord(ctx->ev[2 + b], s_cmp));
    // D2H of this chunk's fluxes on its own stream (overlaps the next chunk's H2D and kernels)
    RRNN_CUDA(cudaStreamWaitEvent(s_out, ctx->ev[2 + b], 0));
    for (int o = 0; o < nout; ++o)
      RRNN_CUDA(cudaMemcpyAsync(out_host[o] + (size_t)c0 * (L + 1), outd[b][o], (size_t)(L + 1) * nc * sizeof(float), cudaMemcpyDeviceToHost, s_out));
    RRNN_CUDA(cudaEventRecord(ctx->ev[4 + b], s_out));
  }
  RRNN_CUDA(cudaStreamSynchronize(s_in));
  RRNN_CUDA(cudaStreamSynchronize(s_out));
  RRNN_CUDA(cudaStreamSynchronize(s_cmp));
  return 0;
}

extern "C" int rrnn_lw_fluxes_host(rrnn_ctx_t* ctx, const rrnn_kdist_t* kd, const rrnn_model_t* const* models, int nmodels,
                                   int ncol, int nlay, int top_at_1, int n_gauss_angles, const float* play, const float* plev,
                                   const float* tlay, const float* tlev, const float* tsfc, const float* sfc_emis,
                                   const rrnn_gas_t* gases, int ngas, float* flux_up, float* flux_dn) {
  RRNN_CHECK(ctx && kd && models && play && plev && tlay && tsfc && sfc_emis && flux_up && flux_dn, "rrnn_lw_fluxes_host: null argument");
  RRNN_CHECK(n_gauss_angles >= 1 && n_gauss_angles <= 4, "rte_lw: n_gauss_angles must be in 1..4");
  if (ncol <= 0) return 0;
  RRNN_CUDA(cudaSetDevice(ctx->device));
  const size_t L = nlay;
  std::vector<HostField> f = {{play, L, {}}, {plev, L + 1, {}}, {tlay, L, {}}, {tlev, L + 1, {}}, {tsfc, 1, {}}, {sfc_emis, 1, {}}};
  float* outs[2] = {flux_up, flux_dn};
  return run_host_pipeline(ctx, true, kd, models, nmodels, ncol, nlay, top_at_1, n_gauss_angles, f, gases, ngas, outs, 2);
}

extern "C" int rrnn_sw_fluxes_host(rrnn_ctx_t* ctx, const rrnn_kdist_t* kd, const rrnn_model_t* const* models, int ncol,
                                   int nlay, int top_at_1, const float* play, const float* plev, const float* tlay,
                                   const float* mu0, const float* sfc_alb, const float* tsi, const rrnn_gas_t* gases,
                                   int ngas, float* flux_up, float* flux_dn, float* flux_dn_dir) {
  RRNN_CHECK(ctx && kd && models && play && plev && tlay && mu0 && sfc_alb && flux_up && flux_dn && flux_dn_dir, "rrnn_sw_fluxes_host: null argument");
  RRNN_CHECK(kd->d_solar_source, "rrnn_sw_fluxes_host: k-distribution has no solar source");
  if (ncol <= 0) return 0;
  RRNN_CUDA(cudaSetDevice(ctx->device));
  const size_t L = nlay;
  std::vector<HostField> f = {{play, L, {}}, {plev, L + 1, {}}, {tlay, L, {}}, {mu0, 1, {}}, {sfc_alb, 1, {}}, {tsi, 1, {}}};
  float* outs[3] = {flux_up, flux_dn, flux_dn_dir};
  return run_host_pipeline(ctx, false, kd, models, 2, ncol, nlay, top_at_1, 1, f, gases, ngas, outs, 3);
}
