// RTE flux solvers, packed variant: two g-points per lane, all arithmetic on fp32x2 register pairs.
//
//  lw_solver_v4  <- lw_solver_noscat + lw_source_noscat + lw_transport_noscat_dn/_up + inlined broadband sums
//                   (rte/kernels/mo_rte_solver_kernels.F90:119-330, 742-776, 950-1009, 301-314), angle loop :332-415
//  sw_solver_v4  <- sw_solver_2stream + sw_two_stream_source + adding (:541-692, 1366-1480, 1526-1637)
//
// Same algorithm and data flow as rte_solvers.cu (every input element read from HBM once, reverse-sweep coefficients
// parked on chip, deterministic cluster/DSMEM combination of the g-point chunks), re-cut for what actually bounds these
// kernels on B200 -- instruction issue, not HBM (ncu: ~100 / ~230 warp instructions per 32 (g-point, layer) elements
// for LW / SW, issue slots half used, DRAM 10-30 %):
//   * one lane owns TWO adjacent g-points: inputs arrive as 8-byte loads (256 B per warp per row), the two-stream /
//     source / recurrence arithmetic runs as FADD2 / FMUL2 / FFMA2 (f32x2.cuh), the per-level broadband sum starts with
//     an in-lane add, and one warp covers 64 g-points (4 warps per column instead of 8);
//   * exp() is 2 MUFU + 5 packed instructions per pair with the product's rounding error folded back in;
//     reciprocals / square roots are MUFU seeds with one packed Newton step;
//   * the reverse-sweep buffer (LW: t, source_up; SW: e, f, alpha -- 8 / 12 B per element) lives in an L2-resident
//     global scratch ring written and read back with evict_last policy, 16-byte accesses per lane.
// Numerically each component performs the same IEEE fp32 operations as the scalar kernels up to FMA contraction order.
#include "solver_common.cuh"
#include "f32x2.cuh"
#include <algorithm>

namespace rrnn {
namespace v4 {

constexpr int kU = 4;  // layers per software-pipelined group

__device__ __forceinline__ f2 ld_once2(const float* p, uint64_t pol) {
  f2 v;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.b64 %0, [%1], %2;" : "=l"(v.v) : "l"(p), "l"(pol));
  return v;
}
__device__ __forceinline__ f2 ld2(const float* p) {
  f2 v;
  asm volatile("ld.global.nc.b64 %0, [%1];" : "=l"(v.v) : "l"(p));
  return v;
}
// reverse-sweep buffer accessors (global scratch kept in L2)
__device__ __forceinline__ void buf_st22(void* p, f2 a, f2 b, uint64_t pol) {
  asm volatile("st.global.L1::no_allocate.L2::cache_hint.v2.b64 [%0], {%1,%2}, %3;" ::"l"(p), "l"(a.v), "l"(b.v), "l"(pol) : "memory");
}
__device__ __forceinline__ void buf_ld22(const void* p, f2& a, f2& b, uint64_t pol) {
  asm volatile("ld.global.L1::no_allocate.L2::cache_hint.v2.b64 {%0,%1}, [%2], %3;" : "=l"(a.v), "=l"(b.v) : "l"(p), "l"(pol) : "memory");
}
__device__ __forceinline__ void buf_st2(void* p, f2 a, uint64_t pol) {
  asm volatile("st.global.L1::no_allocate.L2::cache_hint.b64 [%0], %1, %2;" ::"l"(p), "l"(a.v), "l"(pol) : "memory");
}
__device__ __forceinline__ f2 buf_ld2(const void* p, uint64_t pol) {
  f2 a;
  asm volatile("ld.global.L1::no_allocate.L2::cache_hint.b64 %0, [%1], %2;" : "=l"(a.v) : "l"(p), "l"(pol) : "memory");
  return a;
}
// componentwise select: m ? a : b
__device__ __forceinline__ f2 sel2(bool mx, bool my, f2 a, f2 b) {
  float ax, ay, bx, by;
  unpack2(a, ax, ay);
  unpack2(b, bx, by);
  return mk2(mx ? ax : bx, my ? ay : by);
}

// exp(-x) and 1 - exp(-x) for x >= 0 without the cancellation of the literal 1 - exp(-x) at small x (see common.cuh,
// exp_and_complement): degree-7 Taylor polynomial of expm1 below 0.35, the literal form above.
template <bool FAST>
__device__ __forceinline__ void exp_and_complement2(f2 x, f2& t, f2& omt) {
  const f2 y = neg2(x);
  const f2 e = exp2x<FAST>(y);
  if (FAST) { t = e; omt = splat2(1.0f) - e; return; }
  f2 p = fma2(y, splat2(1.0f / 5040.0f), splat2(1.0f / 720.0f));
  p = fma2(p, y, splat2(1.0f / 120.0f));
  p = fma2(p, y, splat2(1.0f / 24.0f));
  p = fma2(p, y, splat2(1.0f / 6.0f));
  p = fma2(p, y, splat2(0.5f));
  p = fma2(p, y, splat2(1.0f));
  const f2 em1 = p * y;  // expm1(-x)
  float xx, xy;
  unpack2(x, xx, xy);
  const bool sx = xx < 0.35f, sy = xy < 0.35f;
  omt = sel2(sx, sy, neg2(em1), splat2(1.0f) - e);
  t = sel2(sx, sy, splat2(1.0f) + em1, e);
}

// ---------------------------------------------------------------------------------------------------- LW
template <bool FAST>
__global__ void __launch_bounds__(32) lw_solver_v4(const LwParams p) {
  extern __shared__ float smem[];
  constexpr int U = kU;
  const int lane = threadIdx.x;
  const int G = p.ngpt, L = p.nlay;
  const uint64_t pol_in = policy_evict_first();
  const uint64_t pol_buf = policy_evict_last();
  cg::cluster_group cluster = cg::this_cluster();
  const int chunk = (int)cluster.block_rank();
  const int csize = (int)cluster.num_blocks();
  float* fup = smem;            // per-CTA partial fluxes [2][L+1]
  float* fdn = fup + (L + 1);
  // reverse-sweep buffer of this CTA: [L][32 lanes] x (t.x, t.y, sup.x, sup.y)
  char* buf = reinterpret_cast<char*>(p.scratch) + (size_t)blockIdx.x * L * 32 * 16 + (size_t)lane * 16;

  const float tau_thresh = 3.4526698e-4f;  // sqrt(epsilon(1._sp)), mo_rte_solver_kernels.F90:754
  // Sweep order i = 0..L-1 runs from the top of the atmosphere down: layer l(i) = l0 + dl*i; in sweep order layer i is
  // bounded by level rows ent(i) (towards the top) and ext(i) = ent(i+1) (towards the surface).
  const int top = p.top_at_1;
  const int l0 = top ? 0 : L - 1;
  const int sG = (top ? 1 : -1) * G;
  // lw_source_noscat (:770-773) takes source_dn from lev(l+1) and source_up from lev(l) whatever the orientation
  // (quirk Q1): physical for top_at_1; otherwise swapped unless lw_source_bug_compat = 0.
  const bool dn_uses_ext = top || !p.bug_compat;

  const int g = chunk * 64 + 2 * lane;
  const bool act = g < G;                     // ngpt is even: a pair is live or not as a whole
  const int gs = act ? g : chunk * 64;        // idle lanes shadow the chunk's first pair and contribute zero
  const float live = act ? 1.0f : 0.0f;

  for (int col = blockIdx.x / csize; col < p.ncol; col += gridDim.x / csize) {
    for (int i = lane; i < 2 * (L + 1); i += 32) fup[i] = 0.0f;
    __syncwarp();
    const float* tau = p.tau + (size_t)col * L * G + (size_t)l0 * G + gs;  // layer i at tau[i*sG]
    const float* lay = p.lay_source + (size_t)col * L * G + (size_t)l0 * G + gs;
    const float* lev = p.lev_source + (size_t)col * (L + 1) * G + gs;
    const float* lext = lev + (size_t)(top ? 1 : L - 1) * G;               // ext(i) at lext[i*sG]
    const size_t gc_off = (size_t)col * G + gs;
    const f2 emis = ld2(p.sfc_emis + gc_off);
    const f2 ssrc = ld2(p.sfc_source + gc_off);
    const f2 inc = p.inc_flux ? ld2(p.inc_flux + gc_off) : splat2(0.0f);

    for (int imu = 0; imu < p.nmus; ++imu) {
      const f2 D = splat2(p.Ds[imu]);
      const f2 fac = splat2(2.0f * kPi * p.wts[imu] * live);
      const float rad_norm = 2.0f * kPi * p.wts[imu];
      f2 I = map2(inc, [&](float v) { return v / rad_norm; });  // radn_dn(top) = inc_flux/(2 pi w), :196-201
      {
        const float s = warp_sum(hsum2(fac * I));
        if (lane == 0) fdn[top ? 0 : L] += s;
      }
      // ---------------- downward sweep, software pipelined in groups of U layers ----------------
      f2 n_tau[U], n_lay[U], n_ext[U];
      f2 carry = ld_once2(lev + (size_t)(top ? 0 : L) * G, pol_in);  // ent(0)
      auto load_group = [&](int i0) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int o = min(i0 + u, L - 1) * sG;
          n_tau[u] = ld_once2(tau + o, pol_in);
          n_lay[u] = ld_once2(lay + o, pol_in);
          n_ext[u] = ld_once2(lext + o, pol_in);
        }
      };
      load_group(0);
      for (int i0 = 0; i0 < L; i0 += U) {
        f2 c_tau[U], c_lay[U], c_ext[U];
#pragma unroll
        for (int u = 0; u < U; ++u) { c_tau[u] = n_tau[u]; c_lay[u] = n_lay[u]; c_ext[u] = n_ext[u]; }
        if (i0 + U < L) load_group(i0 + U);
        f2 tv[U], sdn[U], sup[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const f2 ent = (u == 0) ? carry : c_ext[u - 1];
          const f2 ext = c_ext[u];
          const f2 tl = c_tau[u] * D;
          f2 t, omt;
          exp_and_complement2<FAST>(tl, t, omt);
          // fact = (1-t)/tau' - t, or its series where tau' is tiny (:757-768)
          const f2 fa = div2<true>(omt, tl) - t;
          const f2 fb = tl * fnma2(tl, splat2(1.0f / 3.0f), splat2(0.5f));
          float tx, ty;
          unpack2(tl, tx, ty);
          const f2 fact = sel2(tx > tau_thresh, ty > tau_thresh, fa, fb);
          const f2 f2x = fact + fact;
          const f2 lev_dn = dn_uses_ext ? ext : ent;
          const f2 lev_up = dn_uses_ext ? ent : ext;
          tv[u] = t;
          sdn[u] = fma2(f2x, c_lay[u] - lev_dn, omt * lev_dn);
          sup[u] = fma2(f2x, c_lay[u] - lev_up, omt * lev_up);
        }
        carry = c_ext[U - 1];
        float red[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int i = i0 + u;
          if (i < L) {  // warp-uniform; the ragged tail of the last group is computed on clamped loads and dropped
            I = fma2(tv[u], I, sdn[u]);
            buf_st22(buf + (size_t)i * (32 * 16), tv[u], sup[u], pol_buf);
          }
          red[u] = hsum2(fac * I);
        }
        multi_reduce<U>(red, lane);
        const int i = i0 + multi_index<U>(lane);
        if (multi_writer<U>(lane) && i < L) fdn[top ? i + 1 : L - 1 - i] += red[0];
      }
      // ---------------- surface ----------------
      f2 Uu = fma2(I, splat2(1.0f) - emis, emis * ssrc);  // :269
      {
        const float s = warp_sum(hsum2(fac * Uu));
        if (lane == 0) fup[top ? L : 0] += s;
      }
      // ---------------- upward sweep (reverse order) from the buffer, software pipelined ----------------
      f2 nt[U], ns[U];
      auto load_back = [&](int i1) {
#pragma unroll
        for (int u = 0; u < U; ++u) buf_ld22(buf + (size_t)max(i1 - u, 0) * (32 * 16), nt[u], ns[u], pol_buf);
      };
      load_back(L - 1);
      for (int i1 = L - 1; i1 >= 0; i1 -= U) {
        f2 bt[U], bs[U];
#pragma unroll
        for (int u = 0; u < U; ++u) { bt[u] = nt[u]; bs[u] = ns[u]; }
        if (i1 - U >= 0) load_back(i1 - U);
        float red[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          if (i1 - u >= 0) Uu = fma2(bt[u], Uu, bs[u]);
          red[u] = hsum2(fac * Uu);
        }
        multi_reduce<U>(red, lane);
        const int i = i1 - multi_index<U>(lane);
        if (multi_writer<U>(lane) && i >= 0) fup[top ? i : L - i] += red[0];
      }
      __syncwarp();
    }
    float* const gout[2] = {p.flux_up + (size_t)col * (L + 1), p.flux_dn + (size_t)col * (L + 1)};
    combine_chunks<true, 2>(fup, L, lane, gout);
  }
}

// ---------------------------------------------------------------------------------------------------- SW
// Two-stream coefficients of one layer for a pair of g-points (sw_two_stream_source :1405-1475; PIFM, Zdunkowski).
template <bool FAST, bool HAS_G>
__device__ __forceinline__ void two_stream2(f2 tau, f2 w0, f2 gg, float mu0, float mu0_inv, f2& Rdif, f2& Tdif, f2& Rdir, f2& Tdir,
                                            f2& Tnos) {
  const float k_min = 1.e-4f;       // mo_rte_solver_kernels.F90:76-82 (single precision)
  const float eps = 1.1920929e-7f;  // epsilon(1._sp)
  const f2 one = splat2(1.0f), quarter = splat2(0.25f);
  Tnos = exp2x<FAST>(tau * splat2(-mu0_inv));
  f2 gamma1, gamma2, gamma3, gamma4, alpha1, alpha2;
  if (HAS_G) {
    gamma1 = fnma2(w0, fma2(gg, splat2(3.0f), splat2(5.0f)), splat2(8.0f)) * quarter;
    gamma2 = (splat2(3.0f) * (w0 * (one - gg))) * quarter;
    gamma3 = fnma2(splat2(3.0f * mu0), gg, splat2(2.0f)) * quarter;
    gamma4 = one - gamma3;
    alpha1 = fma2(gamma1, gamma4, gamma2 * gamma3);
    alpha2 = fma2(gamma1, gamma3, gamma2 * gamma4);
  } else {
    // g = 0 (always, on the NN path): gamma3 = gamma4 = 1/2 exactly, alpha1 = alpha2 = (gamma1 + gamma2)/2
    gamma1 = fnma2(w0, splat2(5.0f), splat2(8.0f)) * quarter;
    gamma2 = (splat2(3.0f) * w0) * quarter;
    gamma3 = splat2(0.5f);
    gamma4 = gamma3;
    alpha1 = (gamma1 + gamma2) * gamma3;
    alpha2 = alpha1;
  }
  const f2 k = sqrt2<FAST>(max2((gamma1 - gamma2) * (gamma1 + gamma2), splat2(k_min)));
  const f2 ekt = exp2x<FAST>(neg2(tau) * k);
  const f2 e2kt = ekt * ekt;
  const f2 k2e = (k + k) * ekt;
  const f2 ome2 = one - e2kt;
  f2 RT = rcp2<FAST>(fma2(gamma1, ome2, k * (one + e2kt)));
  Rdif = (RT * gamma2) * ome2;
  Tdif = RT * k2e;
  const f2 k_mu = k * splat2(mu0);
  const f2 k_g3 = k * gamma3, k_g4 = k * gamma4;
  const f2 om = fnma2(k_mu, k_mu, one);
  float ox, oy;
  unpack2(om, ox, oy);
  const f2 dd = mk2(fabsf(ox) >= eps ? ox : eps, fabsf(oy) >= eps ? oy : eps);
  RT = div2<FAST>(w0 * RT, dd);
  const f2 a_m = one - k_mu, a_p = one + k_mu;
  f2 rd = RT * ((a_m * (alpha2 + k_g3) - (a_p * (alpha2 - k_g3)) * e2kt) - (k2e * fnma2(alpha2, splat2(mu0), gamma3)) * Tnos);
  f2 td = RT * ((k2e * fma2(alpha1, splat2(mu0), gamma4)) - Tnos * (a_p * (alpha1 + k_g4) - (a_m * (alpha1 - k_g4)) * e2kt));
  const f2 lim = one - Tnos;
  rd = max2(splat2(0.0f), min2(rd, lim));
  td = max2(splat2(0.0f), min2(td, lim - rd));
  Rdir = rd;
  Tdir = td;
}

template <bool FAST, bool HAS_G>
__global__ void __launch_bounds__(32) sw_solver_v4(const SwParams p) {
  extern __shared__ float smem[];
  constexpr int U = kU;
  const int lane = threadIdx.x;
  const int G = p.ngpt, L = p.nlay;
  const uint64_t pol_in = policy_evict_first();
  const uint64_t pol_buf = policy_evict_last();
  cg::cluster_group cluster = cg::this_cluster();
  const int chunk = (int)cluster.block_rank();
  const int csize = (int)cluster.num_blocks();
  float* fup = smem;  // per-CTA partial fluxes [3][L+1]
  float* fdn = fup + (L + 1);
  float* fdr = fdn + (L + 1);
  // reverse-sweep buffer of this CTA: [L][32 lanes] x (e, f) 16 B, then [L][32] x alpha 8 B
  char* buf_ef = reinterpret_cast<char*>(p.scratch) + (size_t)blockIdx.x * L * 32 * 24 + (size_t)lane * 16;
  char* buf_a = reinterpret_cast<char*>(p.scratch) + (size_t)blockIdx.x * L * 32 * 24 + (size_t)L * 32 * 16 + (size_t)lane * 8;

  const int top = p.top_at_1;
  const int l0 = top ? 0 : L - 1;
  const int sG = (top ? 1 : -1) * G;
  const int top_level = top ? 0 : L;
  const int g = chunk * 64 + 2 * lane;
  const bool act = g < G;
  const int gs = act ? g : chunk * 64;
  const f2 live = splat2(act ? 1.0f : 0.0f);

  for (int col = blockIdx.x / csize; col < p.ncol; col += gridDim.x / csize) {
    for (int i = lane; i < 3 * (L + 1); i += 32) fup[i] = 0.0f;
    __syncwarp();
    const float* tau = p.tau + (size_t)col * L * G + (size_t)l0 * G + gs;  // layer i (sweep order) at [i*sG]
    const float* ssa = p.ssa + (size_t)col * L * G + (size_t)l0 * G + gs;
    const float* gas = HAS_G ? p.g + (size_t)col * L * G + (size_t)l0 * G + gs : nullptr;
    const size_t gc_off = (size_t)col * G + gs;
    const float mu0 = __ldg(p.mu0 + col);
    const float mu0_inv = 1.0f / mu0;

    f2 dir = (live * ld2(p.inc_flux + gc_off)) * splat2(mu0);                        // :589
    f2 beta = p.inc_flux_dif ? live * ld2(p.inc_flux_dif + gc_off) : splat2(0.0f);   // :590
    f2 alpha = splat2(0.0f);
    {
      const float sd = warp_sum(hsum2(dir)), sb = warp_sum(hsum2(beta + dir));
      if (lane == 0) { fdr[top_level] += sd; fdn[top_level] += sb; }
    }
    // ---------------- sweep 1: top -> surface, software pipelined in groups of U layers ----------------
    f2 n_t[U], n_w[U], n_g[U];
    auto load_group = [&](int i0) {
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int o = min(i0 + u, L - 1) * sG;
        n_t[u] = ld_once2(tau + o, pol_in);
        n_w[u] = ld_once2(ssa + o, pol_in);
        if (HAS_G) n_g[u] = ld_once2(gas + o, pol_in);
      }
    };
    load_group(0);
    for (int i0 = 0; i0 < L; i0 += U) {
      f2 c_t[U], c_w[U], c_g[U];
#pragma unroll
      for (int u = 0; u < U; ++u) { c_t[u] = n_t[u]; c_w[u] = n_w[u]; if (HAS_G) c_g[u] = n_g[u]; }
      if (i0 + U < L) load_group(i0 + U);
      // layer coefficients: independent across the U layers (instruction-level parallelism)
      f2 Rdif[U], Tdif[U], Rdir[U], Tdir[U], Tnos[U];
#pragma unroll
      for (int u = 0; u < U; ++u)
        two_stream2<FAST, HAS_G>(c_t[u], c_w[u], HAS_G ? c_g[u] : splat2(0.0f), mu0, mu0_inv, Rdif[u], Tdif[u], Rdir[u], Tdir[u], Tnos[u]);
      // the sequential part: direct beam and the adding recurrences, eliminated from the top
      float red[2 * U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int i = i0 + u;
        if (i < L) {  // warp-uniform
          const f2 s_up = Rdir[u] * dir;
          const f2 s_dn = Tdir[u] * dir;
          dir = Tnos[u] * dir;
          const f2 d = rcp2<FAST>(fnma2(Rdif[u], alpha, splat2(1.0f)));
          const f2 e = d * Tdif[u];
          const f2 f = d * fma2(Rdif[u], beta, s_up);
          beta = fma2(e, fma2(alpha, s_up, beta), s_dn);
          alpha = fma2(Tdif[u] * e, alpha, Rdif[u]);
          buf_st22(buf_ef + (size_t)i * (32 * 16), e, f, pol_buf);
          buf_st2(buf_a + (size_t)i * (32 * 8), alpha, pol_buf);  // reflectance seen from the level BELOW layer i
        }
        red[u] = hsum2(dir);
        red[U + u] = hsum2(beta + dir);
      }
      multi_reduce<2 * U>(red, lane);
      {
        const int idx = multi_index<2 * U>(lane);
        const int i = i0 + (idx & (U - 1));
        if (multi_writer<2 * U>(lane) && i < L) {
          const int lvl = top ? i + 1 : L - 1 - i;
          if (idx < U) fdr[lvl] += red[0]; else fdn[lvl] += red[0];
        }
      }
    }
    // ---------------- surface ----------------
    const f2 a_s = ld2(p.alb_dif + gc_off);
    const f2 S_s = dir * ld2(p.alb_dir + gc_off);  // source_sfc :1477
    f2 Uu = div2<FAST>(fma2(a_s, beta, S_s), fnma2(a_s, alpha, splat2(1.0f))) * live;
    {
      const int sfc = top ? L : 0;
      const float su = warp_sum(hsum2(Uu)), sa = warp_sum(hsum2(alpha * Uu));
      if (lane == 0) { fup[sfc] += su; fdn[sfc] += sa; }
    }
    // ---------------- sweep 2: surface -> top (back substitution), software pipelined ----------------
    f2 ne[U], nf[U], na[U];
    auto load_back = [&](int i1) {
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int i = max(i1 - u, 0);
        buf_ld22(buf_ef + (size_t)i * (32 * 16), ne[u], nf[u], pol_buf);
        // reflectance of the atmosphere above the level at the top of layer i (0 at the top of the domain)
        na[u] = (i > 0) ? buf_ld2(buf_a + (size_t)(i - 1) * (32 * 8), pol_buf) : splat2(0.0f);
      }
    };
    load_back(L - 1);
    for (int i1 = L - 1; i1 >= 0; i1 -= U) {
      f2 ce[U], cf[U], ca[U];
#pragma unroll
      for (int u = 0; u < U; ++u) { ce[u] = ne[u]; cf[u] = nf[u]; ca[u] = na[u]; }
      if (i1 - U >= 0) load_back(i1 - U);
      float red[2 * U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (i1 - u >= 0) Uu = fma2(ce[u], Uu, cf[u]);
        red[u] = hsum2(Uu);
        red[U + u] = hsum2(ca[u] * Uu);
      }
      multi_reduce<2 * U>(red, lane);
      {
        const int idx = multi_index<2 * U>(lane);
        const int i = i1 - (idx & (U - 1));
        if (multi_writer<2 * U>(lane) && i >= 0) {
          const int lvl = top ? i : L - i;  // level at the top of layer i
          if (idx < U) fup[lvl] += red[0]; else fdn[lvl] += red[0];
        }
      }
    }
    __syncwarp();
    float* const gout[3] = {p.flux_up + (size_t)col * (L + 1), p.flux_dn + (size_t)col * (L + 1),
                            p.flux_dir + (size_t)col * (L + 1)};
    combine_chunks<true, 3>(fup, L, lane, gout);
  }
}

}  // namespace v4

// ---- launchers: return -1 when the shape does not fit the packed kernels (the caller uses rte_solvers.cu) ----
// Resident clusters are capped so that the reverse-sweep scratch of all of them stays L2-sized.
static int resident_clusters(const rrnn_ctx_t* ctx, int occ_clusters, int csize, size_t per_cta_bytes, int ncol) {
  const size_t budget = (size_t)(ctx->solver_scratch_mb > 0 ? ctx->solver_scratch_mb : 72) << 20;
  long long n = (long long)(budget / (per_cta_bytes * (size_t)csize));
  n = std::max<long long>(n, ctx->num_sms / 2);  // never starve the GPU outright
  n = std::min<long long>(n, occ_clusters);
  n = std::min<long long>(n, ncol);
  return (int)std::max<long long>(n, 1);
}

int launch_lw_v4(rrnn_ctx_t* ctx, LwParams& p) {
  const int G = p.ngpt, L = p.nlay;
  const int csize = (G + 63) / 64;
  if ((G & 1) || csize > 8) return -1;
  for (const void* q : {(const void*)p.tau, (const void*)p.lay_source, (const void*)p.lev_source, (const void*)p.sfc_emis,
                        (const void*)p.sfc_source, (const void*)p.inc_flux})
    if ((uintptr_t)q & 7) return -1;  // 8-byte loads of g-point pairs
  const size_t smem = 2 * (size_t)(L + 1) * sizeof(float);
  const size_t per_cta = (size_t)L * 32 * 16;
  cudaLaunchConfig_t cfg; cudaLaunchAttribute attr[1]; int ncta = 0;
  auto kernel = ctx->fast_math ? v4::lw_solver_v4<true> : v4::lw_solver_v4<false>;
  RRNN_CUDA(cluster_config(kernel, csize, smem, p.ncol, ctx->stream, cfg, attr, ncta));
  const int ncl = resident_clusters(ctx, ncta / csize, csize, per_cta, p.ncol);
  ncta = ncl * csize;
  cfg.gridDim = dim3((unsigned)ncta);
  if (int rc = ensure_scratch(ctx, (size_t)ncta * per_cta)) return rc;
  p.scratch = (float*)ctx->scratch;
  RRNN_CUDA(cudaLaunchKernelEx(&cfg, kernel, p));
  return 0;
}

int launch_sw_v4(rrnn_ctx_t* ctx, SwParams& p, bool fast) {
  const int G = p.ngpt, L = p.nlay;
  const int csize = (G + 63) / 64;
  if ((G & 1) || csize > 8) return -1;
  for (const void* q : {(const void*)p.tau, (const void*)p.ssa, (const void*)p.g, (const void*)p.inc_flux, (const void*)p.inc_flux_dif,
                        (const void*)p.alb_dir, (const void*)p.alb_dif})
    if ((uintptr_t)q & 7) return -1;
  const size_t smem = 3 * (size_t)(L + 1) * sizeof(float);
  const size_t per_cta = (size_t)L * 32 * 24;
  cudaLaunchConfig_t cfg; cudaLaunchAttribute attr[1]; int ncta = 0;
  void (*kernel)(const SwParams);
  if (p.g) kernel = fast ? v4::sw_solver_v4<true, true> : v4::sw_solver_v4<false, true>;
  else kernel = fast ? v4::sw_solver_v4<true, false> : v4::sw_solver_v4<false, false>;
  RRNN_CUDA(cluster_config(kernel, csize, smem, p.ncol, ctx->stream, cfg, attr, ncta));
  const int ncl = resident_clusters(ctx, ncta / csize, csize, per_cta, p.ncol);
  ncta = ncl * csize;
  cfg.gridDim = dim3((unsigned)ncta);
  if (int rc = ensure_scratch(ctx, (size_t)ncta * per_cta)) return rc;
  p.scratch = (float*)ctx->scratch;
  RRNN_CUDA(cudaLaunchKernelEx(&cfg, kernel, p));
  return 0;
}

}  // namespace rrnn
