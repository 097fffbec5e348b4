// RTE flux solvers for sm_100a: fused per-(column, g-point) layer scans with the broadband sum over
// g-points done by warp shuffles, so per-g-point radiances never touch HBM.
//
//  lw_solver_kernel  <- lw_solver_noscat + lw_source_noscat + lw_transport_noscat_dn/_up + inlined broadband
//                       sums (rte/kernels/mo_rte_solver_kernels.F90:119-330, 742-776, 950-1009, 301-314) and the
//                       angle loop of lw_solver_noscat_GaussQuad (:332-415).
//  sw_solver_kernel  <- sw_solver_2stream + sw_two_stream_source + adding (:541-692, 1366-1480, 1526-1637).
//
// Work decomposition: one warp owns one (column, 32-g-point chunk); lanes are consecutive g-points, so every
// global access is one fully coalesced 128-byte line per warp.  Each input element (tau, sources / ssa, g) is
// read from HBM exactly once: what the reverse sweep needs is parked in shared memory (8 B per element for LW:
// transmittance and upward source; 12 B for SW: the back-substitution coefficients), which is what bounds the
// number of resident warps.  Per-level partial sums of a warp are combined across the g-chunks of a column
// with one coalesced fp32 red.global per 32 levels.
//
// SW numerics: the reference does three sweeps (direct beam down, adding up, fluxes down).  Here the adding
// recurrences are written as the mirror-image elimination from the top (reflectance `alpha` of the atmosphere
// ABOVE a level to upwelling radiation and downwelling source `beta`), fused with the direct-beam sweep, followed
// by one back-substitution from the surface up.  It is the same linear two-point boundary-value problem solved
// in the other direction: identical in exact arithmetic, and within fp32 rounding (<< 0.01 W m-2) of the
// reference order -- checked against the oracle in tests/test_solvers_gpu.py.
#include "common.cuh"
#include <cooperative_groups.h>

namespace cg = cooperative_groups;

namespace rrnn {

// Combine the per-level partial sums of the g-point chunks of one column.
//  CLUSTER = true : the chunks of a column are the CTAs of one thread-block cluster; rank 0 reads the other
//                   ranks' partial sums through distributed shared memory and adds them in rank order, so the
//                   result is deterministic and the rounding is the same at every level (which is what keeps
//                   heating rates, i.e. differences of adjacent levels, clean).  No memset, no atomics.
//  CLUSTER = false: fallback for more than 8 chunks (ngpt > 256): fp32 atomics on zero-initialised arrays.
template <bool CLUSTER, int NARR>
__device__ __forceinline__ void combine_chunks(float* part /* [NARR][L+1] in this CTA's smem */, int L, int lane,
                                               float* const (&gout)[NARR]) {
  if (CLUSTER) {
    cg::cluster_group cluster = cg::this_cluster();
    cluster.sync();
    if (cluster.block_rank() == 0) {
      const unsigned nr = cluster.num_blocks();
      for (int i = lane; i < NARR * (L + 1); i += 32) {
        float s = part[i];
        for (unsigned r = 1; r < nr; ++r) s += *cluster.map_shared_rank(part + i, r);
        const int a = i / (L + 1);
        gout[a][i - a * (L + 1)] = s;
      }
    }
    cluster.sync();  // keep every rank's shared memory alive until rank 0 has read it
  } else {
    for (int i = lane; i < NARR * (L + 1); i += 32) {
      const int a = i / (L + 1);
      atomicAdd(gout[a] + (i - a * (L + 1)), part[i]);
    }
  }
}

struct LwParams {
  int ngpt, nlay, ncol, top_at_1, nmus, bug_compat, nchunks;
  float Ds[4], wts[4];
  const float* inc_flux;  // (ngpt,ncol) or null
  const float* tau;       // (ngpt,nlay,ncol)
  const float* lay_source;
  const float* lev_source;  // (ngpt,nlay+1,ncol)
  const float* sfc_emis;    // (ngpt,ncol)
  const float* sfc_source;  // (ngpt,ncol)
  float* flux_up;           // (nlay+1,ncol), zero-initialised
  float* flux_dn;
};

constexpr float kPi = 3.14159265358979323846f;
constexpr int kLwUnroll = 4;

template <bool FAST, bool CLUSTER>
__global__ void __launch_bounds__(64) lw_solver_kernel(const LwParams p) {
  extern __shared__ float smem[];
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  const int wpb = blockDim.x >> 5;  // 1 when CLUSTER
  const long long item = (long long)blockIdx.x * wpb + wib;
  if (item >= (long long)p.ncol * p.nchunks) return;  // never true when CLUSTER (grid == ncol * nchunks)
  const int col = (int)(item / p.nchunks);
  const int chunk = (int)(item % p.nchunks);
  const int g = chunk * 32 + lane;
  const bool act = g < p.ngpt;
  const int G = p.ngpt, L = p.nlay;

  // per-warp shared memory: [L][32] float2 (t, src_up), then flux partials [2][L+1]
  const size_t per_warp = (size_t)L * 64 + 2 * (size_t)(L + 1);
  float* wbase = smem + (size_t)wib * per_warp;
  float2* buf = reinterpret_cast<float2*>(wbase);
  float* fup = wbase + (size_t)L * 64;
  float* fdn = fup + (L + 1);
  for (int i = lane; i < 2 * (L + 1); i += 32) fup[i] = 0.0f;
  __syncwarp();

  const size_t gl_off = (size_t)col * L * G + g;         // + l*G
  const size_t gv_off = (size_t)col * (L + 1) * G + g;   // + lev*G
  const size_t gc_off = (size_t)col * G + g;
  const float* tau = p.tau + gl_off;
  const float* lay = p.lay_source + gl_off;
  const float* lev = p.lev_source + gv_off;
  const float tau_thresh = 3.4526698e-4f;  // sqrt(epsilon(1._sp)), mo_rte_solver_kernels.F90:754
  const float emis = act ? p.sfc_emis[gc_off] : 0.0f;
  const float ssrc = act ? p.sfc_source[gc_off] : 0.0f;
  const float inc = (act && p.inc_flux) ? p.inc_flux[gc_off] : 0.0f;

  // sweep direction in memory: top_at_1 -> down sweep walks l = 0..L-1, else L-1..0
  const int l0 = p.top_at_1 ? 0 : L - 1;
  const int dl = p.top_at_1 ? 1 : -1;
  // which level row feeds source_dn / source_up for layer l (array indices):
  //   reference (bug-compatible, Q1): dn <- lev[l+1], up <- lev[l] whatever the orientation.
  //   physical for top_at_1=false:     dn <- lev[l],   up <- lev[l+1].
  const bool swap_lev = (!p.top_at_1) && (!p.bug_compat);

  for (int imu = 0; imu < p.nmus; ++imu) {
    const float D = p.Ds[imu];
    const float fac = 2.0f * kPi * p.wts[imu];
    float I = inc / fac;  // radn_dn(top) = inc_flux/(2 pi w), :196-201
    {
      float s = warp_sum(fac * I);
      if (lane == 0) fdn[p.top_at_1 ? 0 : L] += s;
    }
    // ---------------- downward sweep ----------------
    for (int i0 = 0; i0 < L; i0 += kLwUnroll) {
      float vt[kLwUnroll], vlay[kLwUnroll], vlo[kLwUnroll], vhi[kLwUnroll];
#pragma unroll
      for (int u = 0; u < kLwUnroll; ++u) {
        const int i = i0 + u;
        if (i < L && act) {
          const int l = l0 + dl * i;
          vt[u] = ld_stream(tau + (size_t)l * G);
          vlay[u] = ld_stream(lay + (size_t)l * G);
          vlo[u] = ld_stream(lev + (size_t)l * G);
          vhi[u] = ld_stream(lev + (size_t)(l + 1) * G);
        } else {
          vt[u] = 0.f; vlay[u] = 0.f; vlo[u] = 0.f; vhi[u] = 0.f;
        }
      }
#pragma unroll
      for (int u = 0; u < kLwUnroll; ++u) {
        const int i = i0 + u;
        if (i < L) {
          const int l = l0 + dl * i;
          const float tl = vt[u] * D;
          float t, omt;
          if (FAST) { t = __expf(-tl); omt = 1.0f - t; }
          else exp_and_complement(tl, t, omt);
          float fact;
          if (tl > tau_thresh) fact = fdiv<FAST>(omt, tl) - t;
          else fact = tl * (0.5f - (1.0f / 3.0f) * tl);
          const float lev_dn = swap_lev ? vlo[u] : vhi[u];
          const float lev_up = swap_lev ? vhi[u] : vlo[u];
          const float src_dn = omt * lev_dn + 2.0f * fact * (vlay[u] - lev_dn);
          const float src_up = omt * lev_up + 2.0f * fact * (vlay[u] - lev_up);
          I = t * I + src_dn;
          buf[(size_t)l * 32 + lane] = make_float2(t, src_up);
          const float s = warp_sum(fac * I);
          if (lane == 0) fdn[p.top_at_1 ? l + 1 : l] += s;
        }
      }
    }
    // ---------------- surface ----------------
    float U = I * (1.0f - emis) + emis * ssrc;  // :269
    {
      const float s = warp_sum(fac * U);
      if (lane == 0) fup[p.top_at_1 ? L : 0] += s;
    }
    __syncwarp();
    // ---------------- upward sweep (reverse memory order) ----------------
    for (int i = L - 1; i >= 0; --i) {
      const int l = l0 + dl * i;
      const float2 b = buf[(size_t)l * 32 + lane];
      U = b.x * U + b.y;
      const float s = warp_sum(fac * U);
      if (lane == 0) fup[p.top_at_1 ? l : l + 1] += s;
    }
    __syncwarp();
  }
  // combine the g-chunks of this column
  float* const gout[2] = {p.flux_up + (size_t)col * (L + 1), p.flux_dn + (size_t)col * (L + 1)};
  combine_chunks<CLUSTER, 2>(fup, L, lane, gout);
}

// ---------------------------------------------------------------------------------------------------
struct SwParams {
  int ngpt, nlay, ncol, top_at_1, nchunks;
  const float* inc_flux;      // (ngpt,ncol)
  const float* inc_flux_dif;  // (ngpt,ncol) or null
  const float* tau;
  const float* ssa;
  const float* g;  // or null (g = 0)
  const float* mu0;
  const float* alb_dir;
  const float* alb_dif;
  float* flux_up;  // zero-initialised
  float* flux_dn;
  float* flux_dir;
};

constexpr int kSwUnroll = 2;

template <bool FAST, bool HAS_G, bool CLUSTER>
__global__ void __launch_bounds__(64) sw_solver_kernel(const SwParams p) {
  extern __shared__ float smem[];
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  const int wpb = blockDim.x >> 5;
  const long long item = (long long)blockIdx.x * wpb + wib;
  if (item >= (long long)p.ncol * p.nchunks) return;
  const int col = (int)(item / p.nchunks);
  const int chunk = (int)(item % p.nchunks);
  const int gp = chunk * 32 + lane;
  const bool act = gp < p.ngpt;
  const int G = p.ngpt, L = p.nlay;

  // per-warp shared memory: e[L][32], f[L][32], alpha_below[L][32], then flux partials [3][L+1]
  const size_t per_warp = (size_t)L * 96 + 3 * (size_t)(L + 1);
  float* wbase = smem + (size_t)wib * per_warp;
  float* be = wbase;
  float* bf = be + (size_t)L * 32;
  float* ba = bf + (size_t)L * 32;
  float* fup = ba + (size_t)L * 32;
  float* fdn = fup + (L + 1);
  float* fdr = fdn + (L + 1);
  for (int i = lane; i < 3 * (L + 1); i += 32) fup[i] = 0.0f;
  __syncwarp();

  const size_t gl_off = (size_t)col * L * G + gp;
  const size_t gc_off = (size_t)col * G + gp;
  const float* tau = p.tau + gl_off;
  const float* ssa = p.ssa + gl_off;
  const float* gas = HAS_G ? p.g + gl_off : nullptr;
  const float mu0 = p.mu0[col];
  const float mu0_inv = 1.0f / mu0;
  const float k_min = 1.e-4f;       // mo_rte_solver_kernels.F90:76-82 (single precision)
  const float eps = 1.1920929e-7f;  // epsilon(1._sp)

  const int l0 = p.top_at_1 ? 0 : L - 1;
  const int dl = p.top_at_1 ? 1 : -1;
  const int top_level = p.top_at_1 ? 0 : L;

  float dir = act ? p.inc_flux[gc_off] * mu0 : 0.0f;                      // :589
  float beta = (act && p.inc_flux_dif) ? p.inc_flux_dif[gc_off] : 0.0f;   // :590
  float alpha = 0.0f;
  {
    const float sd = warp_sum(dir), sb = warp_sum(beta + dir);
    if (lane == 0) { fdr[top_level] += sd; fdn[top_level] += sb; }
  }
  // ---------------- sweep 1: top -> surface ----------------
  for (int i0 = 0; i0 < L; i0 += kSwUnroll) {
    float vt[kSwUnroll], vw[kSwUnroll], vg[kSwUnroll];
#pragma unroll
    for (int u = 0; u < kSwUnroll; ++u) {
      const int i = i0 + u;
      if (i < L && act) {
        const int l = l0 + dl * i;
        vt[u] = ld_stream(tau + (size_t)l * G);
        vw[u] = ld_stream(ssa + (size_t)l * G);
        vg[u] = HAS_G ? ld_stream(gas + (size_t)l * G) : 0.0f;
      } else {
        vt[u] = 0.f; vw[u] = 0.f; vg[u] = 0.f;
      }
    }
#pragma unroll
    for (int u = 0; u < kSwUnroll; ++u) {
      const int i = i0 + u;
      if (i < L) {
        const int l = l0 + dl * i;
        const float tauv = vt[u], w0 = vw[u], gg = vg[u];
        // ---- sw_two_stream_source :1405-1475 ----
        const float Tnoscat = exp_neg<FAST>(-tauv * mu0_inv);
        const float gamma1 = (8.0f - w0 * (5.0f + 3.0f * gg)) * 0.25f;
        const float gamma2 = 3.0f * (w0 * (1.0f - gg)) * 0.25f;
        const float gamma3 = (2.0f - 3.0f * mu0 * gg) * 0.25f;
        const float gamma4 = 1.0f - gamma3;
        const float alpha1 = gamma1 * gamma4 + gamma2 * gamma3;
        const float alpha2 = gamma1 * gamma3 + gamma2 * gamma4;
        const float k = fsqrt<FAST>(fmaxf((gamma1 - gamma2) * (gamma1 + gamma2), k_min));
        const float ekt = exp_neg<FAST>(-tauv * k);
        const float e2kt = ekt * ekt;
        const float k2e = 2.0f * k * ekt;
        float RT = rcp<FAST>(k * (1.0f + e2kt) + gamma1 * (1.0f - e2kt));
        const float Rdif = RT * gamma2 * (1.0f - e2kt);
        const float Tdif = RT * 2.0f * k * ekt;
        const float k_mu = k * mu0;
        const float k_mu2 = k_mu * k_mu;
        const float k_gamma3 = k * gamma3;
        const float k_gamma4 = k * gamma4;
        const float om = 1.0f - k_mu2;
        const float dd = (fabsf(om) >= eps) ? om : eps;
        RT = fdiv<FAST>(w0 * RT, dd);
        float Rdir = RT * ((1.0f - k_mu) * (alpha2 + k_gamma3) - (1.0f + k_mu) * (alpha2 - k_gamma3) * e2kt -
                           k2e * (gamma3 - alpha2 * mu0) * Tnoscat);
        float Tdir = RT * (k2e * (gamma4 + alpha1 * mu0) -
                           Tnoscat * ((1.0f + k_mu) * (alpha1 + k_gamma4) - (1.0f - k_mu) * (alpha1 - k_gamma4) * e2kt));
        Rdir = fmaxf(0.0f, fminf(Rdir, 1.0f - Tnoscat));
        Tdir = fmaxf(0.0f, fminf(Tdir, 1.0f - Tnoscat - Rdir));
        const float s_up = Rdir * dir;
        const float s_dn = Tdir * dir;
        dir = Tnoscat * dir;
        // ---- adding, eliminated from the top (mirror image of :1560-1577) ----
        const float d = rcp<FAST>(1.0f - Rdif * alpha);
        const float e = d * Tdif;
        const float f = d * (Rdif * beta + s_up);
        beta = s_dn + e * (beta + alpha * s_up);
        alpha = Rdif + Tdif * e * alpha;
        be[(size_t)l * 32 + lane] = e;
        bf[(size_t)l * 32 + lane] = f;
        ba[(size_t)l * 32 + lane] = alpha;  // reflectance seen from the level BELOW layer l
        const float sd = warp_sum(dir), sb = warp_sum(beta + dir);
        const int lvl = p.top_at_1 ? l + 1 : l;
        if (lane == 0) { fdr[lvl] += sd; fdn[lvl] += sb; }
      }
    }
  }
  // ---------------- surface ----------------
  const float a_s = act ? p.alb_dif[gc_off] : 0.0f;
  const float S_s = act ? dir * p.alb_dir[gc_off] : 0.0f;  // source_sfc :1477
  float U = fdiv<FAST>(a_s * beta + S_s, 1.0f - a_s * alpha);
  {
    const int sfc = p.top_at_1 ? L : 0;
    const float su = warp_sum(U), sa = warp_sum(alpha * U);
    if (lane == 0) { fup[sfc] += su; fdn[sfc] += sa; }
  }
  __syncwarp();
  // ---------------- sweep 2: surface -> top (back substitution) ----------------
  for (int i = L - 1; i >= 0; --i) {
    const int l = l0 + dl * i;
    U = be[(size_t)l * 32 + lane] * U + bf[(size_t)l * 32 + lane];
    const int lvl = p.top_at_1 ? l : l + 1;  // level at the top of layer l
    const float su = warp_sum(U);
    float sa = 0.0f;
    if (i > 0) {
      const int labove = l - dl;
      sa = warp_sum(ba[(size_t)labove * 32 + lane] * U);
    }
    if (lane == 0) { fup[lvl] += su; fdn[lvl] += sa; }
  }
  __syncwarp();
  float* const gout[3] = {p.flux_up + (size_t)col * (L + 1), p.flux_dn + (size_t)col * (L + 1),
                          p.flux_dir + (size_t)col * (L + 1)};
  combine_chunks<CLUSTER, 3>(fup, L, lane, gout);
}

// expand (rte/mo_rte_lw.F90:429-447): band -> g-point
__global__ void expand_kernel(int nbnd, int ngpt, int ncol, const int* __restrict__ gpt2band,
                              const float* __restrict__ in, float* __restrict__ out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)ngpt * ncol) return;
  const int g = (int)(i % ngpt);
  const size_t c = i / ngpt;
  out[i] = in[c * nbnd + gpt2band[g]];
}

}  // namespace rrnn

using namespace rrnn;

static int pick_warps_per_block(size_t per_warp_bytes) {
  // two warps per CTA unless that does not fit
  return (2 * per_warp_bytes <= 200 * 1024) ? 2 : 1;
}

// Launch `kernel` with one 32-thread CTA per (column, chunk) and the chunks of a column forming one cluster.
template <typename P>
static cudaError_t launch_clustered(void (*kernel)(const P), const P& p, long long ncta, int cluster, size_t smem,
                                    cudaStream_t stream) {
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)ncta);
  cfg.blockDim = dim3(32);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)cluster;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, p);
}

extern "C" int rrnn_lw_solver_noscat(rrnn_ctx_t* ctx, int ngpt, int nlay, int ncol, int top_at_1, int nmus, const float* Ds,
                                     const float* weights, const float* inc_flux_d, const float* tau_d,
                                     const float* lay_source_d, const float* lev_source_d, const float* sfc_emis_gpt_d,
                                     const float* sfc_source_d, float* flux_up_d, float* flux_dn_d) {
  RRNN_CHECK(ctx, "rrnn_lw_solver_noscat: null context");
  RRNN_CHECK(ngpt > 0 && nlay > 0 && ncol >= 0, "rrnn_lw_solver_noscat: bad extents");
  RRNN_CHECK(nmus >= 1 && nmus <= 4, "rte_lw: have to ask for between 1 and 4 quadrature points for no-scattering calculation");
  if (ncol == 0) return 0;
  RRNN_CUDA(cudaSetDevice(ctx->device));
  LwParams p{};
  p.ngpt = ngpt; p.nlay = nlay; p.ncol = ncol; p.top_at_1 = top_at_1 ? 1 : 0; p.nmus = nmus;
  p.bug_compat = ctx->lw_source_bug_compat;
  p.nchunks = (ngpt + 31) / 32;
  for (int i = 0; i < nmus; ++i) { p.Ds[i] = Ds[i]; p.wts[i] = weights[i]; }
  p.inc_flux = inc_flux_d; p.tau = tau_d; p.lay_source = lay_source_d; p.lev_source = lev_source_d;
  p.sfc_emis = sfc_emis_gpt_d; p.sfc_source = sfc_source_d; p.flux_up = flux_up_d; p.flux_dn = flux_dn_d;
  const size_t per_warp = ((size_t)nlay * 64 + 2 * (size_t)(nlay + 1)) * sizeof(float);
  RRNN_CHECK(per_warp <= ctx->smem_optin, "rrnn_lw_solver_noscat: nlay too large for the on-chip layer buffer");
  const bool clustered = p.nchunks <= 8;
  const int wpb = clustered ? 1 : pick_warps_per_block(per_warp);
  const size_t smem = per_warp * wpb;
  const long long items = (long long)ncol * p.nchunks;
  const long long blocks = (items + wpb - 1) / wpb;
  RRNN_CHECK(blocks < 2147483647LL, "rrnn_lw_solver_noscat: too many columns for one launch");
  const size_t nflux = (size_t)ncol * (nlay + 1) * sizeof(float);
  if (!clustered) {
    RRNN_CUDA(cudaMemsetAsync(flux_up_d, 0, nflux, ctx->stream));
    RRNN_CUDA(cudaMemsetAsync(flux_dn_d, 0, nflux, ctx->stream));
  }
  const int ps = prof_begin(ctx, K_LW_SOLVER);
  if (clustered) {
    if (ctx->fast_math) RRNN_CUDA(launch_clustered(lw_solver_kernel<true, true>, p, blocks, p.nchunks, smem, ctx->stream));
    else RRNN_CUDA(launch_clustered(lw_solver_kernel<false, true>, p, blocks, p.nchunks, smem, ctx->stream));
  } else if (ctx->fast_math) {
    RRNN_CUDA(cudaFuncSetAttribute(lw_solver_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    lw_solver_kernel<true, false><<<(unsigned)blocks, wpb * 32, smem, ctx->stream>>>(p);
  } else {
    RRNN_CUDA(cudaFuncSetAttribute(lw_solver_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    lw_solver_kernel<false, false><<<(unsigned)blocks, wpb * 32, smem, ctx->stream>>>(p);
  }
  prof_end(ctx, K_LW_SOLVER, ps);
  RRNN_LAUNCH_CHECK(ctx);
  return 0;
}

extern "C" int rrnn_sw_solver_2stream(rrnn_ctx_t* ctx, int ngpt, int nlay, int ncol, int top_at_1, const float* inc_flux_d,
                                      const float* inc_flux_dif_d, const float* tau_d, const float* ssa_d, const float* g_d,
                                      const float* mu0_d, const float* sfc_alb_dir_d, const float* sfc_alb_dif_d,
                                      float* flux_up_d, float* flux_dn_d, float* flux_dir_d) {
  RRNN_CHECK(ctx, "rrnn_sw_solver_2stream: null context");
  RRNN_CHECK(ngpt > 0 && nlay > 0 && ncol >= 0, "rrnn_sw_solver_2stream: bad extents");
  if (ncol == 0) return 0;
  RRNN_CUDA(cudaSetDevice(ctx->device));
  SwParams p{};
  p.ngpt = ngpt; p.nlay = nlay; p.ncol = ncol; p.top_at_1 = top_at_1 ? 1 : 0;
  p.nchunks = (ngpt + 31) / 32;
  p.inc_flux = inc_flux_d; p.inc_flux_dif = inc_flux_dif_d; p.tau = tau_d; p.ssa = ssa_d; p.g = g_d; p.mu0 = mu0_d;
  p.alb_dir = sfc_alb_dir_d; p.alb_dif = sfc_alb_dif_d; p.flux_up = flux_up_d; p.flux_dn = flux_dn_d; p.flux_dir = flux_dir_d;
  const size_t per_warp = ((size_t)nlay * 96 + 3 * (size_t)(nlay + 1)) * sizeof(float);
  RRNN_CHECK(per_warp <= ctx->smem_optin, "rrnn_sw_solver_2stream: nlay too large for the on-chip layer buffer");
  const bool clustered = p.nchunks <= 8;
  const int wpb = clustered ? 1 : pick_warps_per_block(per_warp);
  const size_t smem = per_warp * wpb;
  const long long items = (long long)ncol * p.nchunks;
  const long long blocks = (items + wpb - 1) / wpb;
  RRNN_CHECK(blocks < 2147483647LL, "rrnn_sw_solver_2stream: too many columns for one launch");
  const size_t nflux = (size_t)ncol * (nlay + 1) * sizeof(float);
  if (!clustered) {
    RRNN_CUDA(cudaMemsetAsync(flux_up_d, 0, nflux, ctx->stream));
    RRNN_CUDA(cudaMemsetAsync(flux_dn_d, 0, nflux, ctx->stream));
    RRNN_CUDA(cudaMemsetAsync(flux_dir_d, 0, nflux, ctx->stream));
  }
#define SW_LAUNCH(F, HG)                                                                                                 \
  do {                                                                                                                   \
    if (clustered) {                                                                                                     \
      RRNN_CUDA(launch_clustered(sw_solver_kernel<F, HG, true>, p, blocks, p.nchunks, smem, ctx->stream));               \
    } else {                                                                                                             \
      RRNN_CUDA(cudaFuncSetAttribute(sw_solver_kernel<F, HG, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
      sw_solver_kernel<F, HG, false><<<(unsigned)blocks, wpb * 32, smem, ctx->stream>>>(p);                              \
    }                                                                                                                    \
  } while (0)
  const int ps = prof_begin(ctx, K_SW_SOLVER);
  if (ctx->fast_math) { if (g_d) SW_LAUNCH(true, true); else SW_LAUNCH(true, false); }
  else { if (g_d) SW_LAUNCH(false, true); else SW_LAUNCH(false, false); }
#undef SW_LAUNCH
  prof_end(ctx, K_SW_SOLVER, ps);
  RRNN_LAUNCH_CHECK(ctx);
  return 0;
}

// rte_lw (rte/mo_rte_lw.F90:60-424) for ty_optical_props_1scl
extern "C" int rrnn_rte_lw(rrnn_ctx_t* ctx, const rrnn_kdist_t* kd, int nlay, int ncol, int top_at_1, int n_gauss_angles,
                           const float* inc_flux_d, const float* tau_d, const float* lay_source_d,
                           const float* lev_source_d, const float* sfc_source_d, const float* sfc_emis_d,
                           float* flux_up_d, float* flux_dn_d) {
  RRNN_CHECK(ctx && kd, "rte_lw: null handle");
  // rte/mo_rte_lw.F90:113-125
  static const float gauss_Ds[4][4] = {{1.66f, 0.f, 0.f, 0.f},
                                       {1.18350343f, 2.81649655f, 0.f, 0.f},
                                       {1.09719858f, 1.69338507f, 4.70941630f, 0.f},
                                       {1.06056257f, 1.38282560f, 2.40148179f, 7.15513024f}};
  static const float gauss_wts[4][4] = {{0.5f, 0.f, 0.f, 0.f},
                                        {0.3180413817f, 0.1819586183f, 0.f, 0.f},
                                        {0.2009319137f, 0.2292411064f, 0.0698269799f, 0.f},
                                        {0.1355069134f, 0.2034645680f, 0.1298475476f, 0.0311809710f}};
  RRNN_CHECK(n_gauss_angles <= 4, "rte_lw: asking for too many quadrature points for no-scattering calculation");
  RRNN_CHECK(n_gauss_angles >= 1, "rte_lw: have to ask for at least one quadrature point for no-scattering calculation");
  if (ncol == 0) return 0;
  RRNN_CUDA(cudaSetDevice(ctx->device));
  const int ngpt = kd->ngpt;
  // sfc_emis_gpt lives in the context workspace tail (small: ngpt*ncol floats)
  float* emis_gpt = nullptr;
  RRNN_CUDA(cudaMallocAsync((void**)&emis_gpt, (size_t)ngpt * ncol * sizeof(float), ctx->stream));
  const size_t n = (size_t)ngpt * ncol;
  expand_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(kd->nbnd, ngpt, ncol, kd->d_gpt2band, sfc_emis_d, emis_gpt);
  RRNN_LAUNCH_CHECK(ctx);
  int rc = rrnn_lw_solver_noscat(ctx, ngpt, nlay, ncol, top_at_1, n_gauss_angles, gauss_Ds[n_gauss_angles - 1],
                                 gauss_wts[n_gauss_angles - 1], inc_flux_d, tau_d, lay_source_d, lev_source_d, emis_gpt,
                                 sfc_source_d, flux_up_d, flux_dn_d);
  cudaFreeAsync(emis_gpt, ctx->stream);
  return rc;
}

// rte_sw (rte/mo_rte_sw.F90:48-266) for ty_optical_props_2str
extern "C" int rrnn_rte_sw(rrnn_ctx_t* ctx, int ngpt, int nlay, int ncol, int top_at_1, const float* mu0_d,
                           const float* inc_flux_d, const float* sfc_alb_dir_gpt_d, const float* sfc_alb_dif_gpt_d,
                           const float* inc_flux_dif_d, const float* tau_d, const float* ssa_d, const float* g_d,
                           float* flux_up_d, float* flux_dn_d, float* flux_dn_dir_d) {
  RRNN_CHECK(flux_up_d && flux_dn_d && flux_dn_dir_d, "rte_sw: no space allocated for fluxes");
  return rrnn_sw_solver_2stream(ctx, ngpt, nlay, ncol, top_at_1, inc_flux_d, inc_flux_dif_d, tau_d, ssa_d, g_d, mu0_d,
                                sfc_alb_dir_gpt_d, sfc_alb_dif_gpt_d, flux_up_d, flux_dn_d, flux_dn_dir_d);
}
