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
constexpr int kLwU = 8;  // layers per software-pipelined group (LW)
constexpr int kSwU = 4;  // layers per group (SW)

// Sum N (power of two, <= 8) values per lane over the 32 lanes with N-1 + log2(32/N) shuffles instead of 5N:
// a butterfly that halves the number of live values at every step.  On return v[0] of lane `lane` holds the
// all-lane sum of the original v[multi_index(lane)].
template <int N>
__device__ __forceinline__ void multi_reduce(float (&v)[N], int lane) {
  int off = 16;
#pragma unroll
  for (int n = N; n > 1; n >>= 1) {
    const int half = n >> 1;
    const bool upper = (lane & off) != 0;
#pragma unroll
    for (int k = 0; k < half; ++k) {
      const float send = upper ? v[k] : v[k + half];
      const float keep = upper ? v[k + half] : v[k];
      v[k] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
    off >>= 1;
  }
#pragma unroll
  for (; off >= 1; off >>= 1) v[0] += __shfl_xor_sync(0xffffffffu, v[0], off);
}
template <int N>
__device__ __forceinline__ int multi_index(int lane) {
  int idx = 0, off = 16;
#pragma unroll
  for (int n = N; n > 1; n >>= 1) {
    if (lane & off) idx += n >> 1;
    off >>= 1;
  }
  return idx;
}
template <int N>
__device__ __forceinline__ bool multi_writer(int lane) {  // one lane per distinct index
  return (lane & ((32 / N) - 1)) == 0;
}

template <bool FAST, bool CLUSTER>
__global__ void __launch_bounds__(64) lw_solver_kernel(const LwParams p) {
  extern __shared__ float smem[];
  constexpr int U = kLwU;
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

  // inactive lanes (ngpt not a multiple of 32) read lane 0's g-point and contribute zero
  const int gs = act ? g : chunk * 32;
  const float* tau = p.tau + (size_t)col * L * G + gs;
  const float* lay = p.lay_source + (size_t)col * L * G + gs;
  const float* lev = p.lev_source + (size_t)col * (L + 1) * G + gs;
  const size_t gc_off = (size_t)col * G + gs;
  const float tau_thresh = 3.4526698e-4f;  // sqrt(epsilon(1._sp)), mo_rte_solver_kernels.F90:754
  const float emis = p.sfc_emis[gc_off];
  const float ssrc = p.sfc_source[gc_off];
  const float inc = p.inc_flux ? p.inc_flux[gc_off] : 0.0f;
  const float live = act ? 1.0f : 0.0f;

  // Sweep order i = 0..L-1 runs from the top of the atmosphere down: layer l(i) = l0 + dl*i.  In sweep order
  // layer i is bounded by level rows ent(i) (towards the top) and ext(i) = ent(i+1) (towards the surface).
  const int top = p.top_at_1;
  const int l0 = top ? 0 : L - 1;
  const int dl = top ? 1 : -1;
  // lw_source_noscat (:770-773) takes source_dn from lev(l+1) and source_up from lev(l) whatever the orientation
  // (quirk Q1).  In sweep terms: top_at_1 -> dn uses ext, up uses ent (physical).  Otherwise the reference uses
  // dn <- lev(l+1) = ent, up <- lev(l) = ext; the physical choice is again dn <- ext, up <- ent.
  const bool dn_uses_ext = top || !p.bug_compat;

  for (int imu = 0; imu < p.nmus; ++imu) {
    const float D = p.Ds[imu];
    const float fac = 2.0f * kPi * p.wts[imu] * live;
    float I = inc / (2.0f * kPi * p.wts[imu]);  // radn_dn(top) = inc_flux/(2 pi w), :196-201
    {
      const float s = warp_sum(fac * I);
      if (lane == 0) fdn[top ? 0 : L] += s;
    }
    // ---------------- downward sweep, software pipelined in groups of U layers ----------------
    float n_tau[U], n_lay[U], n_ext[U];
    float carry = ld_stream(lev + (size_t)(top ? 0 : L) * G);  // ent(0)
    auto load_group = [&](int i0) {
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int i = min(i0 + u, L - 1);
        const int l = l0 + dl * i;
        n_tau[u] = ld_stream(tau + (size_t)l * G);
        n_lay[u] = ld_stream(lay + (size_t)l * G);
        n_ext[u] = ld_stream(lev + (size_t)(top ? l + 1 : l) * G);
      }
    };
    load_group(0);
    for (int i0 = 0; i0 < L; i0 += U) {
      float c_tau[U], c_lay[U], c_ext[U];
#pragma unroll
      for (int u = 0; u < U; ++u) { c_tau[u] = n_tau[u]; c_lay[u] = n_lay[u]; c_ext[u] = n_ext[u]; }
      if (i0 + U < L) load_group(i0 + U);
      float tv[U], sdn[U], sup[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const float ent = (u == 0) ? carry : c_ext[u - 1];
        const float ext = c_ext[u];
        const float tl = c_tau[u] * D;
        float t, omt;
        if (FAST) { t = __expf(-tl); omt = 1.0f - t; }
        else exp_and_complement(tl, t, omt);
        float fact;
        if (tl > tau_thresh) fact = fdiv<FAST>(omt, tl) - t;
        else fact = tl * (0.5f - (1.0f / 3.0f) * tl);
        const float lev_dn = dn_uses_ext ? ext : ent;
        const float lev_up = dn_uses_ext ? ent : ext;
        tv[u] = t;
        sdn[u] = omt * lev_dn + 2.0f * fact * (c_lay[u] - lev_dn);
        sup[u] = omt * lev_up + 2.0f * fact * (c_lay[u] - lev_up);
      }
      carry = c_ext[U - 1];
      float red[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int i = i0 + u;
        if (i < L) {  // warp-uniform; the ragged tail of the last group is computed on clamped loads and dropped
          I = tv[u] * I + sdn[u];
          buf[(size_t)(l0 + dl * i) * 32 + lane] = make_float2(tv[u], sup[u]);
        }
        red[u] = fac * I;
      }
      multi_reduce<U>(red, lane);
      const int i = i0 + multi_index<U>(lane);
      if (multi_writer<U>(lane) && i < L) fdn[top ? (l0 + dl * i) + 1 : (l0 + dl * i)] += red[0];
    }
    // ---------------- surface ----------------
    float U0 = I;
    U0 = U0 * (1.0f - emis) + emis * ssrc;  // :269
    {
      const float s = warp_sum(fac * U0);
      if (lane == 0) fup[top ? L : 0] += s;
    }
    __syncwarp();
    // ---------------- upward sweep (reverse order) from the on-chip buffer ----------------
    float Uu = U0;
    for (int i1 = L - 1; i1 >= 0; i1 -= U) {
      float2 b[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int i = max(i1 - u, 0);
        b[u] = buf[(size_t)(l0 + dl * i) * 32 + lane];
      }
      float red[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (i1 - u >= 0) Uu = b[u].x * Uu + b[u].y;
        red[u] = fac * Uu;
      }
      multi_reduce<U>(red, lane);
      const int i = i1 - multi_index<U>(lane);
      if (multi_writer<U>(lane) && i >= 0) fup[top ? (l0 + dl * i) : (l0 + dl * i) + 1] += red[0];
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
  float* flux_up;
  float* flux_dn;
  float* flux_dir;
};

template <bool FAST, bool HAS_G, bool CLUSTER>
__global__ void __launch_bounds__(64) sw_solver_kernel(const SwParams p) {
  extern __shared__ float smem[];
  constexpr int U = kSwU;
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

  const int gs = act ? gp : chunk * 32;  // inactive lanes shadow lane 0 and contribute zero
  const float live = act ? 1.0f : 0.0f;
  const float* tau = p.tau + (size_t)col * L * G + gs;
  const float* ssa = p.ssa + (size_t)col * L * G + gs;
  const float* gas = HAS_G ? p.g + (size_t)col * L * G + gs : nullptr;
  const size_t gc_off = (size_t)col * G + gs;
  const float mu0 = p.mu0[col];
  const float mu0_inv = 1.0f / mu0;
  const float k_min = 1.e-4f;       // mo_rte_solver_kernels.F90:76-82 (single precision)
  const float eps = 1.1920929e-7f;  // epsilon(1._sp)

  const int top = p.top_at_1;
  const int l0 = top ? 0 : L - 1;
  const int dl = top ? 1 : -1;
  const int top_level = top ? 0 : L;

  float dir = live * p.inc_flux[gc_off] * mu0;                         // :589
  float beta = p.inc_flux_dif ? live * p.inc_flux_dif[gc_off] : 0.0f;  // :590
  float alpha = 0.0f;
  {
    const float sd = warp_sum(dir), sb = warp_sum(beta + dir);
    if (lane == 0) { fdr[top_level] += sd; fdn[top_level] += sb; }
  }
  // ---------------- sweep 1: top -> surface, software pipelined in groups of U layers ----------------
  float n_t[U], n_w[U], n_g[U];
  auto load_group = [&](int i0) {
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int l = l0 + dl * min(i0 + u, L - 1);
      n_t[u] = ld_stream(tau + (size_t)l * G);
      n_w[u] = ld_stream(ssa + (size_t)l * G);
      n_g[u] = HAS_G ? ld_stream(gas + (size_t)l * G) : 0.0f;
    }
  };
  load_group(0);
  for (int i0 = 0; i0 < L; i0 += U) {
    float c_t[U], c_w[U], c_g[U];
#pragma unroll
    for (int u = 0; u < U; ++u) { c_t[u] = n_t[u]; c_w[u] = n_w[u]; c_g[u] = n_g[u]; }
    if (i0 + U < L) load_group(i0 + U);
    // layer coefficients: independent across the U layers (instruction-level parallelism)
    float Rdif[U], Tdif[U], Rdir[U], Tdir[U], Tnos[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const float tauv = c_t[u], w0 = c_w[u], gg = c_g[u];
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
      Rdif[u] = RT * gamma2 * (1.0f - e2kt);
      Tdif[u] = RT * 2.0f * k * ekt;
      const float k_mu = k * mu0;
      const float k_mu2 = k_mu * k_mu;
      const float k_gamma3 = k * gamma3;
      const float k_gamma4 = k * gamma4;
      const float om = 1.0f - k_mu2;
      const float dd = (fabsf(om) >= eps) ? om : eps;
      RT = fdiv<FAST>(w0 * RT, dd);
      float rd = RT * ((1.0f - k_mu) * (alpha2 + k_gamma3) - (1.0f + k_mu) * (alpha2 - k_gamma3) * e2kt -
                       k2e * (gamma3 - alpha2 * mu0) * Tnoscat);
      float td = RT * (k2e * (gamma4 + alpha1 * mu0) -
                       Tnoscat * ((1.0f + k_mu) * (alpha1 + k_gamma4) - (1.0f - k_mu) * (alpha1 - k_gamma4) * e2kt));
      rd = fmaxf(0.0f, fminf(rd, 1.0f - Tnoscat));
      td = fmaxf(0.0f, fminf(td, 1.0f - Tnoscat - rd));
      Rdir[u] = rd; Tdir[u] = td; Tnos[u] = Tnoscat;
    }
    // the sequential part: direct beam and the adding recurrences, eliminated from the top
    float red[2 * U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int i = i0 + u;
      if (i < L) {  // warp-uniform
        const int l = l0 + dl * i;
        const float s_up = Rdir[u] * dir;
        const float s_dn = Tdir[u] * dir;
        dir = Tnos[u] * dir;
        const float d = rcp<FAST>(1.0f - Rdif[u] * alpha);
        const float e = d * Tdif[u];
        const float f = d * (Rdif[u] * beta + s_up);
        beta = s_dn + e * (beta + alpha * s_up);
        alpha = Rdif[u] + Tdif[u] * e * alpha;
        be[(size_t)l * 32 + lane] = e;
        bf[(size_t)l * 32 + lane] = f;
        ba[(size_t)l * 32 + lane] = alpha;  // reflectance seen from the level BELOW layer l
      }
      red[u] = dir;
      red[U + u] = beta + dir;
    }
    multi_reduce<2 * U>(red, lane);
    {
      const int idx = multi_index<2 * U>(lane);
      const int i = i0 + (idx & (U - 1));
      if (multi_writer<2 * U>(lane) && i < L) {
        const int l = l0 + dl * i;
        const int lvl = top ? l + 1 : l;
        if (idx < U) fdr[lvl] += red[0]; else fdn[lvl] += red[0];
      }
    }
  }
  // ---------------- surface ----------------
  const float a_s = p.alb_dif[gc_off];
  const float S_s = dir * p.alb_dir[gc_off];  // source_sfc :1477
  float Uu = fdiv<FAST>(a_s * beta + S_s, 1.0f - a_s * alpha) * live;
  {
    const int sfc = top ? L : 0;
    const float su = warp_sum(Uu), sa = warp_sum(alpha * Uu);
    if (lane == 0) { fup[sfc] += su; fdn[sfc] += sa; }
  }
  __syncwarp();
  // ---------------- sweep 2: surface -> top (back substitution) ----------------
  for (int i1 = L - 1; i1 >= 0; i1 -= U) {
    float ce[U], cf[U], ca[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int i = max(i1 - u, 0);
      const int l = l0 + dl * i;
      ce[u] = be[(size_t)l * 32 + lane];
      cf[u] = bf[(size_t)l * 32 + lane];
      // reflectance of the atmosphere above the level at the top of layer i (0 at the top of the domain)
      ca[u] = (i > 0) ? ba[(size_t)(l - dl) * 32 + lane] : 0.0f;
    }
    float red[2 * U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (i1 - u >= 0) Uu = ce[u] * Uu + cf[u];
      red[u] = Uu;
      red[U + u] = ca[u] * Uu;
    }
    multi_reduce<2 * U>(red, lane);
    {
      const int idx = multi_index<2 * U>(lane);
      const int i = i1 - (idx & (U - 1));
      if (multi_writer<2 * U>(lane) && i >= 0) {
        const int l = l0 + dl * i;
        const int lvl = top ? l : l + 1;  // level at the top of layer i
        if (idx < U) fup[lvl] += red[0]; else fdn[lvl] += red[0];
      }
    }
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
