// Column-parallel reductions either side of the solvers (rows N2 / N4 of the scope table's "next" list):
//   * by-band fluxes  -- ty_fluxes_byband%reduce, extensions/mo_fluxes_byband.F90:41-131, kernels
//     sum_byband / net_byband_full / net_byband_precalc, extensions/mo_fluxes_byband_kernels.F90:33-86;
//   * net flux        -- net_broadband_precalc (ty_fluxes_broadband%reduce with flux_net associated), the same
//     elementwise difference as net_byband_precalc;
//   * compute_optimal_angles, rrtmgp/mo_gas_optics_rrtmgp.F90:1712-1758;
//   * set_solar_variability, rrtmgp/mo_gas_optics_rrtmgp.F90:1058-1095 (host arithmetic on ngpt numbers + upload).
// This fork's layout throughout: g-point (or band) fastest, then level, then column -- mo_fluxes_byband.F90 itself
// still declares the upstream (ncol,nlev,ngpt) order.  Sums run serially in g-point order, as the reference loops do,
// so the results are bit-identical to the restatement in oracle/.
#include "common.cuh"
#include <cmath>

namespace rrnn {

__device__ __forceinline__ float4 ld_stream4(const float* p) {
  float4 v;
  // volatile on purpose: hoisting the four loads of a band ahead of the serial sums measured slower (g256: 0.44 against 0.80 of
  // the copy peak)
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}

// one thread per (column-level row, band).  The 16 (or so) g-points of a band are 64 contiguous bytes: a warp covers
// 32 consecutive bands = 2 KB of the row-major g-point array, every fetched sector is used.
// VEC: every band starts at a multiple of 4 g-points and holds a multiple of 4 (checked on the host, with ngpt % 4 == 0 and
// 16-byte aligned arrays): 16-byte loads, a quarter of the load instructions and L1 wavefronts.  The additions keep the
// g-point order either way.
template <bool NET, bool VEC>
__global__ void __launch_bounds__(256) byband_kernel(size_t nrow, int ngpt, int nbnd, const int* __restrict__ band_lims,
                                                     const float* __restrict__ a, const float* __restrict__ b,
                                                     float* __restrict__ out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nrow * nbnd) return;
  const size_t row = i / nbnd;
  const int bnd = (int)(i - row * nbnd);
  const int g0 = band_lims[2 * bnd] - 1, g1 = band_lims[2 * bnd + 1] - 1;
  const float* pa = a + row * ngpt;
  const float* pb = NET ? b + row * ngpt : nullptr;  // a = down, b = up
  float acc;
  if (VEC) {
    float4 x = ld_stream4(pa + g0);
    if (NET) {
      float4 y = ld_stream4(pb + g0);
      acc = __fsub_rn(x.x, y.x);                       // (net + dn) - up, :70-72
      acc = __fsub_rn(__fadd_rn(acc, x.y), y.y); acc = __fsub_rn(__fadd_rn(acc, x.z), y.z); acc = __fsub_rn(__fadd_rn(acc, x.w), y.w);
      for (int g = g0 + 4; g <= g1; g += 4) {
        x = ld_stream4(pa + g); y = ld_stream4(pb + g);
        acc = __fsub_rn(__fadd_rn(acc, x.x), y.x); acc = __fsub_rn(__fadd_rn(acc, x.y), y.y);
        acc = __fsub_rn(__fadd_rn(acc, x.z), y.z); acc = __fsub_rn(__fadd_rn(acc, x.w), y.w);
      }
    } else {
      acc = __fadd_rn(__fadd_rn(__fadd_rn(x.x, x.y), x.z), x.w);
      for (int g = g0 + 4; g <= g1; g += 4) {
        x = ld_stream4(pa + g);
        acc = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(acc, x.x), x.y), x.z), x.w);
      }
    }
  } else if (NET) {
    acc = __fsub_rn(pa[g0], pb[g0]);
    for (int g = g0 + 1; g <= g1; ++g) acc = __fsub_rn(__fadd_rn(acc, pa[g]), pb[g]);
  } else {
    acc = pa[g0];
    for (int g = g0 + 1; g <= g1; ++g) acc = __fadd_rn(acc, pa[g]);
  }
  out[i] = acc;
}

__global__ void __launch_bounds__(256) net_flux_kernel(size_t n, const float* __restrict__ dn, const float* __restrict__ up,
                                                       float* __restrict__ net) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) net[i] = dn[i] - up[i];
}

__global__ void __launch_bounds__(256) net_flux4_kernel(size_t n4, const float* __restrict__ dn, const float* __restrict__ up,
                                                        float* __restrict__ net) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const float4 d = ld_stream4(dn + 4 * i), u = ld_stream4(up + 4 * i);
  reinterpret_cast<float4*>(net)[i] = make_float4(d.x - u.x, d.y - u.y, d.z - u.z, d.w - u.w);
}

// one thread per (column, g-point); consecutive lanes = consecutive g-points, so every layer is one coalesced row.
__global__ void __launch_bounds__(256) optimal_angles_kernel(int ncol, int nlay, int ngpt, const int* __restrict__ gpt2band,
                                                             const float* __restrict__ fit, const float* __restrict__ tau,
                                                             float* __restrict__ angles) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)ncol * ngpt) return;
  const size_t col = i / ngpt;
  const int g = (int)(i - col * ngpt);
  const float* p = tau + col * (size_t)nlay * ngpt + g;
  float t = 0.0f;
  for (int l = 0; l < nlay; ++l) t = __fadd_rn(t, ld_stream(p + (size_t)l * ngpt));
  const float trans_total = expf(-t);
  const int bnd = gpt2band[g];
  angles[i] = __fadd_rn(__fmul_rn(fit[2 * bnd], trans_total), fit[2 * bnd + 1]);
}

static inline unsigned nblk(size_t n, int t = 256) { return (unsigned)((n + t - 1) / t); }

static bool aligned16(const void* p) { return ((uintptr_t)p & 15) == 0; }

// 16-byte loads are possible when every band is a whole number of aligned 4-g-point groups
static bool bands_vectorise(const rrnn_kdist_t* kd) {
  if (kd->ngpt & 3) return false;
  for (int b = 0; b < kd->nbnd; ++b) {
    const int s0 = kd->band_lims_gpt[2 * b] - 1, n = kd->band_lims_gpt[2 * b + 1] - s0;
    if ((s0 & 3) || (n & 3)) return false;
  }
  return true;
}

}  // namespace rrnn
using namespace rrnn;

extern "C" int rrnn_sum_byband(rrnn_ctx_t* ctx, const rrnn_kdist_t* kd, int nlev, int ncol, const float* gpt_flux_d,
                               float* bnd_flux_d) {
  RRNN_CHECK(ctx && kd, "reduce: null handle");
  RRNN_CHECK(gpt_flux_d && bnd_flux_d, "reduce: null array");
  if (ncol <= 0 || nlev <= 0) return 0;
  RRNN_CUDA(cudaSetDevice(ctx->device));
  const size_t nrow = (size_t)ncol * nlev;
  if (bands_vectorise(kd) && aligned16(gpt_flux_d))
    byband_kernel<false, true><<<nblk(nrow * kd->nbnd), 256, 0, ctx->stream>>>(nrow, kd->ngpt, kd->nbnd, kd->d_band_lims_gpt,
                                                                              gpt_flux_d, nullptr, bnd_flux_d);
  else
    byband_kernel<false, false><<<nblk(nrow * kd->nbnd), 256, 0, ctx->stream>>>(nrow, kd->ngpt, kd->nbnd, kd->d_band_lims_gpt,
                                                                               gpt_flux_d, nullptr, bnd_flux_d);
  RRNN_LAUNCH_CHECK(ctx);
  return 0;
}

extern "C" int rrnn_net_byband(rrnn_ctx_t* ctx, const rrnn_kdist_t* kd, int nlev, int ncol, const float* gpt_flux_dn_d,
                               const float* gpt_flux_up_d, float* bnd_flux_net_d) {
  RRNN_CHECK(ctx && kd, "reduce: null handle");
  RRNN_CHECK(gpt_flux_dn_d && gpt_flux_up_d && bnd_flux_net_d, "reduce: null array");
  if (ncol <= 0 || nlev <= 0) return 0;
  RRNN_CUDA(cudaSetDevice(ctx->device));
  const size_t nrow = (size_t)ncol * nlev;
  if (bands_vectorise(kd) && aligned16(gpt_flux_dn_d) && aligned16(gpt_flux_up_d))
    byband_kernel<true, true><<<nblk(nrow * kd->nbnd), 256, 0, ctx->stream>>>(nrow, kd->ngpt, kd->nbnd, kd->d_band_lims_gpt,
                                                                             gpt_flux_dn_d, gpt_flux_up_d, bnd_flux_net_d);
  else
    byband_kernel<true, false><<<nblk(nrow * kd->nbnd), 256, 0, ctx->stream>>>(nrow, kd->ngpt, kd->nbnd, kd->d_band_lims_gpt,
                                                                              gpt_flux_dn_d, gpt_flux_up_d, bnd_flux_net_d);
  RRNN_LAUNCH_CHECK(ctx);
  return 0;
}

extern "C" int rrnn_net_flux(rrnn_ctx_t* ctx, size_t n, const float* flux_dn_d, const float* flux_up_d, float* flux_net_d) {
  RRNN_CHECK(ctx, "reduce: null context");
  RRNN_CHECK(flux_dn_d && flux_up_d && flux_net_d, "reduce: null array");
  if (n == 0) return 0;
  RRNN_CUDA(cudaSetDevice(ctx->device));
  if (!(n & 3) && aligned16(flux_dn_d) && aligned16(flux_up_d) && aligned16(flux_net_d))
    net_flux4_kernel<<<nblk(n / 4), 256, 0, ctx->stream>>>(n / 4, flux_dn_d, flux_up_d, flux_net_d);
  else
    net_flux_kernel<<<nblk(n), 256, 0, ctx->stream>>>(n, flux_dn_d, flux_up_d, flux_net_d);
  RRNN_LAUNCH_CHECK(ctx);
  return 0;
}

extern "C" int rrnn_kdist_set_optimal_angle_fit(rrnn_kdist_t* k, const float* fit) {
  RRNN_CHECK(k && fit, "set_optimal_angle_fit: null argument");
  k->optimal_angle_fit.assign(fit, fit + 2 * (size_t)k->nbnd);
  RRNN_CUDA(cudaSetDevice(k->device));
  if (!k->d_optimal_angle_fit) RRNN_CUDA(cudaMalloc((void**)&k->d_optimal_angle_fit, 2 * (size_t)k->nbnd * sizeof(float)));
  RRNN_CUDA(cudaMemcpy(k->d_optimal_angle_fit, fit, 2 * (size_t)k->nbnd * sizeof(float), cudaMemcpyHostToDevice));
  return 0;
}

extern "C" int rrnn_compute_optimal_angles(rrnn_ctx_t* ctx, const rrnn_kdist_t* kd, int nlay, int ncol, const float* tau_d,
                                           float* optimal_angles_d) {
  RRNN_CHECK(ctx && kd, "gas_optics%compute_optimal_angles: null handle");
  RRNN_CHECK(kd->d_optimal_angle_fit, "gas_optics%compute_optimal_angles: no optimal_angle_fit loaded");
  RRNN_CHECK(tau_d && optimal_angles_d, "gas_optics%compute_optimal_angles: null array");
  if (ncol <= 0) return 0;
  RRNN_CUDA(cudaSetDevice(ctx->device));
  optimal_angles_kernel<<<nblk((size_t)ncol * kd->ngpt), 256, 0, ctx->stream>>>(ncol, nlay, kd->ngpt, kd->d_gpt2band,
                                                                               kd->d_optimal_angle_fit, tau_d, optimal_angles_d);
  RRNN_LAUNCH_CHECK(ctx);
  return 0;
}

extern "C" int rrnn_kdist_set_solar_tables(rrnn_kdist_t* k, const float* quiet, const float* facular, const float* sunspot) {
  RRNN_CHECK(k && quiet && facular && sunspot, "set_solar_tables: null argument");
  RRNN_CHECK(!k->solar_source.empty(), "set_solar_tables: no solar source (not a shortwave k-distribution)");
  k->solar_quiet.assign(quiet, quiet + k->ngpt);
  k->solar_facular.assign(facular, facular + k->ngpt);
  k->solar_sunspot.assign(sunspot, sunspot + k->ngpt);
  return 0;
}

extern "C" int rrnn_kdist_set_solar_variability(rrnn_kdist_t* k, float mg_index, float sb_index, int have_tsi, float tsi) {
  RRNN_CHECK(k, "set_solar_variability: null handle");
  RRNN_CHECK(!k->solar_quiet.empty(), "set_solar_variability: no solar variability tables loaded");
  RRNN_CHECK(sb_index >= 0.f, "sb_index out of range");  // the later message wins, :1078-1079
  RRNN_CHECK(mg_index >= 0.f, "mg_index out of range");
  const float a_offset = 0.1495954f, b_offset = 0.00066696f;
  for (int g = 0; g < k->ngpt; ++g) {
    volatile float fac = (mg_index - a_offset) * k->solar_facular[g];  // volatile: no contraction into FMAs by the host compiler
    volatile float spot = (sb_index - b_offset) * k->solar_sunspot[g];
    volatile float s = k->solar_quiet[g] + fac;
    k->solar_source[g] = s + spot;
  }
  if (have_tsi) return rrnn_kdist_set_tsi(k, tsi);
  RRNN_CUDA(cudaSetDevice(k->device));
  RRNN_CUDA(cudaMemcpy(k->d_solar_source, k->solar_source.data(), k->solar_source.size() * sizeof(float), cudaMemcpyHostToDevice));
  return 0;
}

// solar_var_ind_interp, extensions/solar_variability/mo_solar_variability.F90:91-183 (host arithmetic, as in the reference)
extern "C" int rrnn_solar_var_ind_interp(const float* avgcyc_ind, int nsolarfrac, float solcycfrac, float* mg_index_out, float* sb_index_out) {
  RRNN_CHECK(avgcyc_ind && mg_index_out && sb_index_out && nsolarfrac >= 3, "solar_var_ind_interp: no index table loaded");
  RRNN_CHECK(solcycfrac >= 0.f && solcycfrac <= 1.f, "solar_var_ind_interp: solcycfrac out of range");
  auto mg = [&](int i) { return avgcyc_ind[2 * (i - 1)]; };       // avgcyc_ind(1, i), 1-based like the reference
  auto sb = [&](int i) { return avgcyc_ind[2 * (i - 1) + 1]; };   // avgcyc_ind(2, i)
  if (solcycfrac == 0.f) { *mg_index_out = mg(1); *sb_index_out = sb(1); return 0; }
  if (solcycfrac == 1.f) { *mg_index_out = mg(nsolarfrac); *sb_index_out = sb(nsolarfrac); return 0; }
  volatile float intrvl_len = 1.0f / (float)(nsolarfrac - 2);     // volatile: every step rounded to fp32, no contraction
  volatile float intrvl_len_hf = 0.5f * intrvl_len;
  int sfid = 1;
  volatile float fraclo = 0.f, frachi = intrvl_len_hf;            // the first half month of the cycle
  volatile float hi_edge = 1.0f - intrvl_len_hf;
  if (solcycfrac > intrvl_len_hf && solcycfrac < hi_edge) {       // month centres
    volatile float x = (solcycfrac - intrvl_len_hf) * (float)(nsolarfrac - 2);
    sfid = (int)std::floor(x) + 2;
    volatile float lo = (float)(sfid - 2) * intrvl_len;
    fraclo = lo + intrvl_len_hf;
    frachi = fraclo + intrvl_len;
  }
  if (solcycfrac >= hi_edge) {                                    // the last half month
    sfid = nsolarfrac - 1;
    fraclo = hi_edge;
    frachi = 1.0f;
  }
  volatile float num = solcycfrac - fraclo, den = frachi - fraclo;
  volatile float intfrac = num / den;
  volatile float dm = mg(sfid + 1) - mg(sfid), ds = sb(sfid + 1) - sb(sfid);
  volatile float pm = intfrac * dm, ps = intfrac * ds;
  *mg_index_out = mg(sfid) + pm;
  *sb_index_out = sb(sfid) + ps;
  return 0;
}

extern "C" int rrnn_kdist_get_solar_source(const rrnn_kdist_t* k, float* solar_source_out) {
  RRNN_CHECK(k && solar_source_out, "get_solar_source: null argument");
  RRNN_CHECK(!k->solar_source.empty(), "get_solar_source: no solar source");
  std::memcpy(solar_source_out, k->solar_source.data(), k->solar_source.size() * sizeof(float));
  return 0;
}
