// RTE flux solvers, TMA-staged packed kernels (the default): two g-points per lane, fp32x2 arithmetic, inputs staged by
// the TMA unit, reverse-sweep rows through an L2-resident scratch.
//
//  lw_solver_v6  <- lw_solver_noscat + lw_source_noscat + lw_transport_noscat_dn/_up + inlined broadband sums
//                   (rte/kernels/mo_rte_solver_kernels.F90:119-330, 742-776, 950-1009, 301-314), angle loop :332-415
//  sw_solver_v6  <- sw_solver_2stream + sw_two_stream_source + adding (:541-692, 1366-1480, 1526-1637)
//
//   * one elected lane issues ONE cp.async.bulk.tensor.2d per input array per group of 8 layers (box 64 g-points x 8 rows of
//     the [rows][ngpt] tensor) into a ring of shared-memory stages, one or two groups ahead, signalled by mbarriers; lanes
//     read their two g-points with LDS.64 at immediate offsets -- no per-lane address arithmetic, no prefetch registers;
//   * the reverse-sweep coefficients (LW: t, source_up -- 16 B per lane and layer; SW: e, f, alpha_above -- 24 B) go to an
//     L2-resident scratch with per-lane 8-byte stores (evict_last) and come back by one bulk copy per group into the idle
//     input ring (see the v6 note below);
//   * arithmetic in f32x2.cuh; orientation and the level-source convention are template parameters.
// One warp = one solver (a 64-g-point chunk of one column), `solver_warps` of them per CTA on adjacent columns, the
// ceil(ngpt/64) chunk-CTAs of a column form a cluster, partial fluxes are combined through DSMEM in rank order
// (deterministic), clusters are persistent over columns.
// History: v3 (rte_solvers.cu, one g-point per lane, the fallback for odd shapes) -> v4 (packed, per-lane loads) -> v5 (TMA
// inputs, staged bulk stores of the scratch) -> v6; v4 and v5 are gone from the tree (git history; profiles/r1*).
#include "solver_common.cuh"
#include <cstdio>
#include <cstdlib>
#include "f32x2.cuh"
#include <cuda.h>
#include <algorithm>
#include <type_traits>

namespace rrnn {
namespace v5 {

#ifndef RRNN_V5_DISCARD
#define RRNN_V5_DISCARD 1
#endif
constexpr int MAX_WARPS = 4;  // solvers (warps) per CTA
constexpr int LW_U = 8;       // layers per group (one TMA box)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
#ifndef RRNN_MBAR_SUSPEND_NS
#define RRNN_MBAR_SUSPEND_NS 1000   // suspend-time hint of mbarrier.try_wait: how long the hardware may park the thread per attempt
#endif
__device__ __forceinline__ uint32_t mbar_try(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(ok) : "r"(bar), "r"(parity), "r"((uint32_t)RRNN_MBAR_SUSPEND_NS) : "memory");
  return ok;
}
// the phase has not completed at the first attempt: poll (kept out of line: the hot loops only carry the first attempt)
__device__ __noinline__ void mbar_wait_slow(uint32_t bar, uint32_t parity) {
  unsigned spins = 0;
  while (!mbar_try(bar, parity))
    if (++spins > (1u << 26)) __trap();  // a wait of seconds is a protocol bug: fail the launch instead of hanging
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (!mbar_try(bar, parity)) mbar_wait_slow(bar, parity);
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, uint32_t bar, uint64_t pol) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3}], [%4], %5;" ::"r"(dst),
               "l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1), "r"(bar), "l"(pol)
               : "memory");
}
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint64_t pol) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(bar), "l"(pol)
               : "memory");
}
__device__ __forceinline__ void bulk_store(void* dst, uint32_t src, uint32_t bytes, uint64_t pol) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(dst), "r"(src), "r"(bytes), "l"(pol) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

__device__ __forceinline__ f2 lds2(const void* p) { f2 v; v.v = *reinterpret_cast<const unsigned long long*>(p); return v; }
__device__ __forceinline__ void lds22(const void* p, f2& a, f2& b) {
  const ulonglong2 v = *reinterpret_cast<const ulonglong2*>(p);
  a.v = v.x; b.v = v.y;
}
__device__ __forceinline__ void sts22(void* p, f2 a, f2 b) { *reinterpret_cast<ulonglong2*>(p) = make_ulonglong2(a.v, b.v); }
__device__ __forceinline__ void sts2(void* p, f2 a) { *reinterpret_cast<unsigned long long*>(p) = a.v; }
__device__ __forceinline__ f2 ldg2(const float* p) { f2 v; asm volatile("ld.global.nc.b64 %0, [%1];" : "=l"(v.v) : "l"(p)); return v; }
__device__ __forceinline__ f2 sel2(bool mx, bool my, f2 a, f2 b) {
  float ax, ay, bx, by;
  unpack2(a, ax, ay);
  unpack2(b, bx, by);
  return mk2(mx ? ax : bx, my ? ay : by);
}

// a * b as an instruction of its own.  ptxas contracts a single-use mul.rn.f32x2 into the add.rn / sub.rn.f32x2 that consumes it
// (seen in SASS: FMUL2 + FADD2 -> FFMA2; the explicit .rn does not stop it for the packed forms), which changes the last bit of
// the factored Planck sources pfrac * B(T) against the materialised arrays.  An FMA with a -0 addend is the same product, rounded
// once -- but ptxas folds a LITERAL -0 addend back into a multiply and contracts that (lw_solver_v7, bottom-up: 32 FADD2 became
// FFMA2).  So the -0 comes from the kernel parameters (LwV5Params::neg_zero, set by the launchers): a value the compiler cannot
// see through, an FFMA2 that stays one, no extra instruction.
__device__ __forceinline__ f2 mul_keep(f2 a, f2 b, f2 nz) { return fma2(a, b, nz); }

// exp(-x) and 1 - exp(-x) for x >= 0 without the cancellation of the literal 1 - exp(-x) at small x (common.cuh,
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
  t = e;   // (exp2x is within ~1 ulp everywhere; only the complement needs the series)
}

// The reverse-sweep scratch is dead once the upward sweep has pulled a group back into shared memory: discarding its L2
// lines (no write-back; the contents become undefined until the next column's downward sweep rewrites them in full) saves
// the DRAM write of every line that was still resident -- the L2 only writes dirty lines back when it evicts them.
__device__ __forceinline__ void discard_scratch(const uint8_t* base, uint32_t bytes, int lane) {
  if (RRNN_V5_DISCARD) {   // at most 8 rows of 768 B = 48 lines: two rounds of the 32 lanes, no loop
    const uint32_t o = (uint32_t)lane * 128u;
    if (o < bytes) asm volatile("discard.global.L2 [%0], 128;" ::"l"(base + o) : "memory");
    if (o + 4096u < bytes) asm volatile("discard.global.L2 [%0], 128;" ::"l"(base + o + 4096u) : "memory");
  }
}

// one lane of a converged warp (elect.sync): ptxas then issues the uniform-datapath TMA instructions straight, without
// the per-lane retry loop it wraps around them under an ordinary `lane == 0` branch
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(pred));
  return pred != 0;
}

// Rows of a group of U layers in a [rows][ngpt] tensor.  Sweep order i = k*U + u runs from the top of the atmosphere
// down; the tensor row of sweep layer i is r0 + i (TOP) or r0 - i (bottom-up arrays), r0 = row of sweep layer 0.
// The TMA box is [start, start + U): start = r0 + k*U (TOP) or r0 - k*U - (U-1); sweep layer k*U + u then sits in
// shared-memory row u (TOP) or U-1-u.  Two boxes can leave the tensor: the last box of the last column (TOP; rows past
// the end are zero-filled by the TMA unit and belong to layers >= nlay, never used) and the ragged last box of column 0
// bottom-up, whose start would be negative: that one is moved to row 0 and `shift` says by how much.
template <bool TOP, int U>
__device__ __forceinline__ int box_start(int r0, int k, int& shift) {
  if (TOP) { shift = 0; return r0 + k * U; }
  const int start = r0 - k * U - (U - 1);
  shift = min(start, 0);  // <= 0
  return start - shift;
}
template <bool TOP, int U>
__device__ __forceinline__ int box_row(int u, int shift) { return TOP ? u : max(U - 1 - u + shift, 0); }

struct LwV5Params {
  LwParams b;
  int ngroups;
  int warp_smem;  // bytes of shared memory per warp
  int* next_col;  // dynamic column assignment (see next_columns)
  float neg_zero = -0.0f;  // see mul_keep
};

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
  // the brackets cancel heavily near k*mu0 = 1: every difference is closed by an FMA (one rounding less per term)
  f2 rd = fnma2(a_p * (alpha2 - k_g3), e2kt, a_m * (alpha2 + k_g3));
  rd = RT * fnma2(k2e * fnma2(alpha2, splat2(mu0), gamma3), Tnos, rd);
  f2 td = fnma2(a_m * (alpha1 - k_g4), e2kt, a_p * (alpha1 + k_g4));
  td = RT * fnma2(Tnos, td, k2e * fma2(alpha1, splat2(mu0), gamma4));
  const f2 lim = one - Tnos;
  rd = max2(splat2(0.0f), min2(rd, lim));
  td = max2(splat2(0.0f), min2(td, lim - rd));
  Rdir = rd;
  Tdir = td;
}

// The same coefficients for U layers at once, written step by step across the layers: the U evaluations are independent,
// and presenting them to the compiler already interleaved is what lets one warp cover its own MUFU / FMA latencies
// (each warp runs alone on its scheduler slot most of the time).
// nz: -0.0 from the kernel parameters (see mul_keep): gamma2 = 0.75 w0 as ONE instruction that ptxas cannot contract into the sums
// gamma1 +- gamma2 (the reference rounds gamma2 before it adds).
template <bool FAST, bool HAS_G, int U>
__device__ __forceinline__ void two_stream2_batch(const f2 (&tau)[U], const f2 (&w0)[U], const f2 (&gg)[U], float mu0, float mu0_inv,
                                                  f2 (&Rdif)[U], f2 (&Tdif)[U], f2 (&Rdir)[U], f2 (&Tdir)[U], f2 (&Tnos)[U],
                                                  f2 nz = splat2(-0.0f)) {
  const float k_min = 1.e-4f;       // mo_rte_solver_kernels.F90:76-82 (single precision)
  const float eps = 1.1920929e-7f;  // epsilon(1._sp)
  const f2 one = splat2(1.0f), quarter = splat2(0.25f), half = splat2(0.5f);
  const f2 mu = splat2(mu0), nmi = splat2(-mu0_inv);
  f2 gamma1[U], gamma2[U], gamma3[U], gamma4[U], alpha1[U], alpha2[U], k[U], ekt[U];
#pragma unroll
  for (int u = 0; u < U; ++u) Tnos[u] = tau[u] * nmi;
#pragma unroll
  for (int u = 0; u < U; ++u) {
    if (HAS_G) {
      gamma1[u] = fnma2(w0[u], fma2(gg[u], splat2(3.0f), splat2(5.0f)), splat2(8.0f)) * quarter;
      gamma2[u] = (splat2(3.0f) * (w0[u] * (one - gg[u]))) * quarter;
      gamma3[u] = fnma2(splat2(3.0f * mu0), gg[u], splat2(2.0f)) * quarter;
      gamma4[u] = one - gamma3[u];
      alpha1[u] = fma2(gamma1[u], gamma4[u], gamma2[u] * gamma3[u]);
      alpha2[u] = fma2(gamma1[u], gamma3[u], gamma2[u] * gamma4[u]);
    } else {
      // g = 0 (always, on the NN path): gamma3 = gamma4 = 1/2 exactly, alpha1 = alpha2 = (gamma1 + gamma2)/2
      // (not w0 * 0.75: ptxas contracts a single multiply into the sums below and the bits leave the reference's; the trailing
      // exact * 0.25 is harmless when contracted)
      // one instruction each, same bits: RN(8 - 5 w) / 4 == RN(2 - 1.25 w) and RN(3 w) / 4 == RN(0.75 w) (a power of two
      // commutes with the rounding)
      gamma1[u] = fma2(w0[u], splat2(-1.25f), splat2(2.0f));
      gamma2[u] = fma2(w0[u], splat2(0.75f), nz);
      gamma3[u] = half;
      gamma4[u] = half;
      alpha1[u] = (gamma1[u] + gamma2[u]) * half;
      alpha2[u] = alpha1[u];
    }
  }
#pragma unroll
  for (int u = 0; u < U; ++u) Tnos[u] = exp2x<FAST>(Tnos[u]);
#pragma unroll
  for (int u = 0; u < U; ++u) k[u] = max2((gamma1[u] - gamma2[u]) * (gamma1[u] + gamma2[u]), splat2(k_min));
#pragma unroll
  for (int u = 0; u < U; ++u) k[u] = sqrt2<FAST>(k[u]);
#pragma unroll
  for (int u = 0; u < U; ++u) ekt[u] = exp2x<FAST>(neg2(tau[u]) * k[u]);
  f2 e2kt[U], k2e[U], ome2[U], RT[U];
#pragma unroll
  for (int u = 0; u < U; ++u) {
    e2kt[u] = ekt[u] * ekt[u];
    k2e[u] = (k[u] + k[u]) * ekt[u];
    ome2[u] = one - e2kt[u];
    RT[u] = fma2(gamma1[u], ome2[u], k[u] * (one + e2kt[u]));
  }
#pragma unroll
  for (int u = 0; u < U; ++u) RT[u] = rcp2<FAST>(RT[u]);
  f2 k_mu[U], dd[U];
#pragma unroll
  for (int u = 0; u < U; ++u) {
    Rdif[u] = (RT[u] * gamma2[u]) * ome2[u];
    Tdif[u] = RT[u] * k2e[u];
    k_mu[u] = k[u] * mu;
    const f2 om = fnma2(k_mu[u], k_mu[u], one);
    float ox, oy;
    unpack2(om, ox, oy);
    dd[u] = mk2(fabsf(ox) >= eps ? ox : eps, fabsf(oy) >= eps ? oy : eps);
  }
#pragma unroll
  for (int u = 0; u < U; ++u) RT[u] = div2<FAST>(w0[u] * RT[u], dd[u]);
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const f2 k_g3 = k[u] * gamma3[u], k_g4 = k[u] * gamma4[u];
    const f2 a_m = one - k_mu[u], a_p = one + k_mu[u];
    // the brackets cancel heavily near k*mu0 = 1: every difference is closed by an FMA (one rounding less per term)
    f2 rd = fnma2(a_p * (alpha2[u] - k_g3), e2kt[u], a_m * (alpha2[u] + k_g3));
    rd = RT[u] * fnma2(k2e[u] * fnma2(alpha2[u], mu, gamma3[u]), Tnos[u], rd);
    f2 td = fnma2(a_m * (alpha1[u] - k_g4), e2kt[u], a_p * (alpha1[u] + k_g4));
    td = RT[u] * fnma2(Tnos[u], td, k2e[u] * fma2(alpha1[u], mu, gamma4[u]));
    const f2 lim = one - Tnos[u];
    rd = max2(splat2(0.0f), min2(rd, lim));
    td = max2(splat2(0.0f), min2(td, lim - rd));
    Rdir[u] = rd;
    Tdir[u] = td;
  }
}

struct SwV5Params {
  SwParams b;
  int ngroups;
  int warp_smem;
  int* next_col;
  float neg_zero = -0.0f;  // see mul_keep / two_stream2_batch
};

// ==================================================================================================== v6
// Second generation of the TMA-staged packed solvers.  What ncu and three ablation builds of sw_solver_v5 showed
// (profiles/r2_sw_solver_ablation.md): the kernel is latency-bound -- its time is inversely proportional to the resident
// warps (7.7 / 11.5 warps per SM: 6.2 / 4.45 ms) and indifferent to the L2 spill of the scratch (60 against 137 layers at
// equal work: no faster) --, the butterfly reductions cost 12 % although they are 7 % of the instructions, and 2.4 of
// 4.45 ms remain with all arithmetic removed: the per-group machinery (staging tiles, proxy fences, MEMBAR, elect / bulk
// store sequences, mbarrier waits of the upward sweep) is what a column trip waits for.  So here
//   * the reverse-sweep rows go to the L2-resident scratch with plain per-lane 8-byte stores at immediate offsets
//     (three coalesced 256-byte segments per layer: e | f | alpha_above) and come back by per-lane loads issued one group
//     of 8 layers ahead into registers: no staging tiles, no fences, no bulk-copy issue, no mbarrier on the way up;
//   * the shared memory that frees (6 KB of 17.7 per solver) is what lets 8 instead of 6 CTAs = 16 instead of 11.5 solver
//     warps reside per SM (the kernels stay at <= 128 registers);
//   * the per-level broadband sums are a transposition through 2.3 KB of shared memory (each lane stores its 16 values,
//     16 lanes x 2 halves each add one value over 16 lanes with four 16-byte loads, one shuffle joins the halves):
//     42 instead of ~100 instructions per group and direction, and no selects;
//   * a ragged last group only computes the halves it needs (137 = 17 x 8 + 1 used to cost a whole group).
// Arithmetic, orientation handling, the cluster combine and every interface are those of v5.
__device__ __forceinline__ void stg_scr(uint8_t* p, f2 v, uint64_t pol) {
  asm volatile("st.global.L1::no_allocate.L2::cache_hint.b64 [%0], %1, %2;" ::"l"(p), "l"(v.v), "l"(pol) : "memory");
}
__device__ __forceinline__ f2 ldg_scr(const uint8_t* p, uint64_t pol) {
  f2 v;
  asm volatile("ld.global.L1::no_allocate.L2::cache_hint.b64 %0, [%1], %2;" : "=l"(v.v) : "l"(p), "l"(pol) : "memory");
  return v;
}
#ifndef RRNN_SW_CLD_SKIP
// all-sky SW, what a half group (4 layers x 64 g-points) that sees no cloud does: 0 = the increment like everybody else; 1 = skips it
// and takes the g == 0 coefficient formulas (a second copy of the two-stream batch: 9.4 ms per 100 000 x 60 columns of the all-sky
// benchmark -- the loop body outgrows the instruction cache); 2 = skips the increment only (7.9 ms; 8.2 ms for 0).  A/B on one B200.
#define RRNN_SW_CLD_SKIP 2
#endif
#ifndef RRNN_V6_SW_S
#define RRNN_V6_SW_S 3  // stages of the input ring (groups of 8 layers in flight ahead of the downward sweep)
#endif
#ifndef RRNN_V6_LW_MINB
#define RRNN_V6_LW_MINB 4  // 4 CTAs of 128 threads = 16 warps per SM: at most 128 registers per thread
#endif
#ifndef RRNN_V6_MINB
#define RRNN_V6_MINB 3  // 3 CTAs of 128 threads = 12 warps per SM: at most 168 registers per thread
#endif
constexpr int TR_PITCH = 36;  // floats per row of the transposition buffer (32 lanes + 4: conflict-free 16-byte reads)
// All-lane sums of N (8 or 16) values per lane through shared memory: on return lane l holds the sum of v[l % N].
// Fixed summation order (deterministic).  tr: [N][TR_PITCH] floats owned by the warp.
template <int N>
__device__ __forceinline__ float tr_reduce(const float (&v)[N], float* tr, int lane) {
  static_assert(N == 8 || N == 16, "tr_reduce: 8 or 16 values");
  constexpr int PER = 32 / (32 / N) / 1;  // lanes summed by one reader = N ... readers per value = 32 / N
  constexpr int NRD = 32 / N;             // readers per value (2 or 4), each over 32 / NRD = N lanes
  (void)PER;
  __syncwarp();  // the readers of the previous call are done with tr
#pragma unroll
  for (int i = 0; i < N; ++i) tr[i * TR_PITCH + lane] = v[i];
  __syncwarp();
  const int idx = lane & (N - 1), part = lane / N;
  const float4* src = reinterpret_cast<const float4*>(tr + idx * TR_PITCH + part * N);
  float4 q[N / 4];
#pragma unroll
  for (int j = 0; j < N / 4; ++j) q[j] = src[j];
  float s[N / 4];
#pragma unroll
  for (int j = 0; j < N / 4; ++j) s[j] = (q[j].x + q[j].y) + (q[j].z + q[j].w);
  float t = (N == 16) ? (s[0] + s[1]) + (s[2] + s[N / 4 - 1]) : s[0] + s[1];
  if (NRD == 4) t += __shfl_xor_sync(0xffffffffu, t, 8);
  t += __shfl_xor_sync(0xffffffffu, t, 16);
  return t;
}
// The same sums, and on the way the sums over the 8 lanes = 16 g-points = ONE BAND that each reader adds up anyway: lane l ends up
// with the total of v[l % N] (as tr_reduce, same association, same bits) and with the partial sums of the bands its quarter of the
// row covers -- N = 16: bands 2 (l / 16) and 2 (l / 16) + 1 of the warp's four; N = 8: band l / 8 (bB unused).
template <int N>
__device__ __forceinline__ float tr_reduce_bands(const float (&v)[N], float* tr, int lane, float& bA, float& bB) {
  static_assert(N == 8 || N == 16, "tr_reduce: 8 or 16 values");
  constexpr int NRD = 32 / N;
  __syncwarp();
#pragma unroll
  for (int i = 0; i < N; ++i) tr[i * TR_PITCH + lane] = v[i];
  __syncwarp();
  const int idx = lane & (N - 1), part = lane / N;
  const float4* src = reinterpret_cast<const float4*>(tr + idx * TR_PITCH + part * N);
  float4 q[N / 4];
#pragma unroll
  for (int j = 0; j < N / 4; ++j) q[j] = src[j];
  float s[N / 4];
#pragma unroll
  for (int j = 0; j < N / 4; ++j) s[j] = (q[j].x + q[j].y) + (q[j].z + q[j].w);
  bA = s[0] + s[1];
  bB = (N == 16) ? s[2] + s[N / 4 - 1] : 0.0f;
  float t = (N == 16) ? bA + bB : bA;
  if (NRD == 4) t += __shfl_xor_sync(0xffffffffu, t, 8);
  t += __shfl_xor_sync(0xffffffffu, t, 16);
  return t;
}
// tr_reduce<8> for the wide kernels (four g-points per lane): the 8 lanes a reader adds are 32 g-points = TWO bands, its two float4
// partial sums; same association as tr_reduce<8>, so the broadband sum keeps its bits
__device__ __forceinline__ float tr_reduce_bands_wide(const float (&v)[8], float* tr, int lane, float& bA, float& bB) {
  __syncwarp();
#pragma unroll
  for (int i = 0; i < 8; ++i) tr[i * TR_PITCH + lane] = v[i];
  __syncwarp();
  const int idx = lane & 7, part = lane >> 3;
  const float4* src = reinterpret_cast<const float4*>(tr + idx * TR_PITCH + part * 8);
  const float4 q0 = src[0], q1 = src[1];
  bA = (q0.x + q0.y) + (q0.z + q0.w);
  bB = (q1.x + q1.y) + (q1.z + q1.w);
  float t = bA + bB;
  t += __shfl_xor_sync(0xffffffffu, t, 8);
  t += __shfl_xor_sync(0xffffffffu, t, 16);
  return t;
}
// sum over each group of 8 lanes (one band): every lane of the group gets it
__device__ __forceinline__ float band_sum8(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  v += __shfl_xor_sync(0xffffffffu, v, 4);
  return v;
}

// Dynamic column assignment.  The clusters are persistent, but which columns a cluster solves is NOT fixed by its index: each
// takes its first block of `nwarps` adjacent columns by index and every further one from a global counter.  With a static stride a
// cluster that becomes resident late -- because another kernel holds some SMs when the solver is launched: the NCCL gather of the
// previous piece of the shard does (bench.py, N > 1) -- still owes its full share of the columns and finishes long after the
// others (measured at 8 GPUs: lw_solver 14.4 -> 18.1 ms per 125 000-column shard).  The leader thread of the cluster fetches the
// next block at the START of a column trip (the atomic's latency hides under the sweeps) and hands it to every rank through
// distributed shared memory just before the cluster barrier the trip ends with anyway: no extra rendezvous.  Which cluster
// solves a column does not change its fluxes (every column is reduced in rank order), so results stay bit-identical.
struct NextColumns {
  int* slot;   // [2] in this CTA's shared memory, double-buffered by trip parity
  int fetched;
  bool leader;
  __device__ __forceinline__ void begin(int* counter, int nwarps, int first_dynamic) {
    if (leader) fetched = atomicAdd(counter, nwarps) + first_dynamic;
  }
  __device__ __forceinline__ void publish(cg::cluster_group& cluster, int csize, int trip) {   // call right before cluster.sync()
    if (leader)
      for (int r = 0; r < csize; ++r) *cluster.map_shared_rank(slot + ((trip + 1) & 1), r) = fetched;
  }
  __device__ __forceinline__ int next(int trip) const { return slot[(trip + 1) & 1]; }   // after that cluster.sync()
};

// rows of the reverse-sweep scratch: three 256-byte segments (32 lanes x 8 B) per layer
constexpr int SW6_ROW = 768, SW6_F = 256, SW6_A = 512;

// GM: where the asymmetry parameter comes from.  0: g == 0 (the NN gas optics; nothing read), 1: a (ngpt,nlay,ncol) array,
// 2: CLOUDS -- by-band cloud properties whose increment has not been applied (SwParams::cld: t2 | s2 | sg2, one 192-byte row per
// layer): inc_2stream_by_2stream_bybnd (mo_optical_props_kernels.F90:453-485) on gas properties with g == 0 runs in registers,
//   tau = tau1 + t2,  ssa = (tau1 ssa1 + s2) / max(eps, tau),  g = sg2 / max(eps, tau1 ssa1 + s2),
// so the all-sky path neither rewrites tau / ssa nor materialises g.
// BND: by-band fluxes as well (SwParams::bnd_*): the per-level sums stop at a band on their way to the broadband sum.  A template
// parameter, not a run-time flag: the flag's tests and the second copy of the reduction cost the broadband-only kernel 5 %.
template <bool FAST, int GM, bool TOP, bool BND = false>
__global__ void __launch_bounds__(32 * MAX_WARPS, RRNN_V6_MINB) sw_solver_v6(const __grid_constant__ SwV5Params pp, const __grid_constant__ CUtensorMap tm_tau,
                                                   const __grid_constant__ CUtensorMap tm_ssa, const __grid_constant__ CUtensorMap tm_g) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  constexpr bool HAS_G = GM != 0;
  constexpr int NIN = GM == 1 ? 3 : 2;
  // GM == 2: tau | ssa | the cloud rows (8 x 192 B), padded so that two stages hold the upward sweep's ring like the GM == 1 layout
  constexpr int STAGE = GM == 2 ? 3 * 8 * 256 : NIN * 8 * 256;
  constexpr int STAGE_TX = NIN * 8 * 256 + (GM == 2 ? 8 * 192 : 0);   // bytes the TMA unit delivers per stage
  constexpr int U = 8, S = GM == 0 ? RRNN_V6_SW_S : 2, SB = 2, H = 4, NH = U / H;  // groups of 8 layers (one TMA box), coefficients in halves of 4
  static_assert(SB * U * SW6_ROW <= S * STAGE, "the upward sweep's stages live in the input ring");
  const SwParams& p = pp.b;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;  // every warp is its own solver
  const int G = p.ngpt, L = p.nlay;
  const uint64_t pol_in = policy_evict_first();
  const uint64_t pol_buf = policy_evict_last();
  cg::cluster_group cluster = cg::this_cluster();
  const int chunk = (int)cluster.block_rank();
  const int csize = (int)cluster.num_blocks();

  uint8_t* smem = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u) + (size_t)warp * pp.warp_smem;
  uint8_t* in_ring = smem;                                                   // [S][STAGE]: tau, ssa (, g | cloud rows), U rows each
  float* tr = reinterpret_cast<float*>(in_ring + S * STAGE);                 // [16][TR_PITCH]
  float* part = tr + 16 * TR_PITCH;                                          // [2 sets][3][L+1]
  const int part_set = 3 * (L + 1) + ((L + 1) & 1);                          // keeps the barriers 8-byte aligned
  uint64_t* bars = reinterpret_cast<uint64_t*>(part + 2 * part_set);
  const uint32_t bar_in = smem_u32(bars), bar_bb = smem_u32(bars + S);
  const uint32_t in_a = smem_u32(in_ring);
  if (lane == 0) {
    for (int s = 0; s < S + SB; ++s) mbar_init(bar_in + 8 * s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  uint32_t n_in = 0, n_bb = 0;

  const int g = chunk * 64 + 2 * lane;
  const bool act = g < G;
  const int gs = act ? g : chunk * 64;
  const f2 live = splat2(act ? 1.0f : 0.0f);
  const int NG = pp.ngroups;
  const int NGF = L / U;
  // this lane's 8-byte slot in the e-segment of scratch row 0 of this solver
  uint8_t* const srow = reinterpret_cast<uint8_t*>(p.scratch) + ((size_t)blockIdx.x * nwarps + warp) * L * SW6_ROW + (size_t)lane * 8u;
  const uint32_t lane_in = (uint32_t)lane * 8u;
  const int top_level = TOP ? 0 : L;
  // which of the 16 reduced values of a group this lane ends up with: lanes 0-7 / 16-23 quantity A of layer u = lane & 7,
  // lanes 8-15 / 24-31 quantity B; only lanes < 16 write
  const int ru = lane & 7;
  const bool rB = (lane & 8) != 0, rW = lane < 16;
  // GM == 2: byte offsets of the bands of this lane's two g-points in a 64-byte table segment
  uint32_t bo0 = 0;   // (both g-points of a lane lie in one band: SwParams::pairs_in_band, checked by the launcher)
  if (GM == 2) bo0 = 4u * (uint32_t)__ldg(p.gpt2band + gs);
  auto band_pair = [&](const uint8_t* seg) { return splat2(*reinterpret_cast<const float*>(seg + bo0)); };

  NextColumns nx;
  nx.slot = reinterpret_cast<int*>(smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u) + (size_t)nwarps * pp.warp_smem);
  nx.leader = chunk == 0 && threadIdx.x == 0;
  nx.fetched = 0;
  int ncols_done = 0;
  for (int cb = (blockIdx.x / csize) * nwarps; cb < p.ncol; ++ncols_done) {
    nx.begin(pp.next_col, nwarps, (int)(gridDim.x / csize) * nwarps);
    const bool owner = cb + warp < p.ncol;
    const int col = owner ? cb + warp : p.ncol - 1;
    float* fup = part + (ncols_done & 1) * part_set;  // this column's partial fluxes [3][L+1]
    float* fdn = fup + (L + 1);
    float* fdr = fdn + (L + 1);
    for (int i = lane; i < 3 * (L + 1); i += 32) fup[i] = 0.0f;
    const size_t gc_off = (size_t)col * G + gs;
    const float mu0 = __ldg(p.mu0 + col);
    const float mu0_inv = 1.0f / mu0;
    const int lay0 = col * L + (TOP ? 0 : L - 1);
    f2 dir = (live * ldg2(p.inc_flux + gc_off)) * splat2(mu0);                        // :589
    f2 beta = p.inc_flux_dif ? live * ldg2(p.inc_flux_dif + gc_off) : splat2(0.0f);   // :590
    f2 alpha = splat2(0.0f);
    const f2 a_s = ldg2(p.alb_dif + gc_off);
    const f2 a_d = ldg2(p.alb_dir + gc_off);
    __syncwarp();
    {
      const float sd = warp_sum(hsum2(dir)), sb = warp_sum(hsum2(beta + dir));
      if (lane == 0) { fdr[top_level] += sd; fdn[top_level] += sb; }
    }
    // by-band outputs (warp-uniform): this lane's band = chunk * 4 + lane / 8; rows (nbnd) of level `lev` of this column
    constexpr bool bands = BND;
    const int my_band = chunk * 4 + (lane >> 3);
    const size_t bcol = (size_t)col * (L + 1) * (size_t)p.nbnd;
    if (bands) {
      const float bd = band_sum8(hsum2(dir)), bb = band_sum8(hsum2(beta + dir));
      if (owner && (lane & 7) == 0 && my_band < p.nbnd) {
        p.bnd_dir[bcol + (size_t)top_level * p.nbnd + my_band] = bd;
        p.bnd_dn[bcol + (size_t)top_level * p.nbnd + my_band] = bb;
      }
    }
    auto issue_in = [&](int k) {
      if (k < NG) {
        const uint32_t st = (n_in + (uint32_t)k) % S;
        int sh;
        const int rl = box_start<TOP, U>(lay0, k, sh);
        if (elect_one()) {
          const uint32_t bar = bar_in + 8 * st;
          const uint32_t dst = in_a + st * STAGE;
          mbar_expect_tx(bar, STAGE_TX);
          tma_load_2d(dst, &tm_tau, chunk * 64, rl, bar, pol_in);
          tma_load_2d(dst + U * 256, &tm_ssa, chunk * 64, rl, bar, pol_in);
          if (GM == 1) tma_load_2d(dst + 2 * U * 256, &tm_g, chunk * 64, rl, bar, pol_in);
          if (GM == 2) tma_load_2d(dst + 2 * U * 256, &tm_g, 0, rl, bar, pol_in);   // (tm_g maps the (48,nlay,ncol) cloud rows)
        }
        __syncwarp();
      }
    };
#pragma unroll
    for (int k = 0; k < S - 1; ++k) issue_in(k);
    // per-level broadband sums of a group: reduced one group later (their latency then overlaps the next group's arithmetic)
    float pend[2 * U];
#pragma unroll
    for (int u = 0; u < 2 * U; ++u) pend[u] = 0.0f;
    int pend_k = -1;
    // by-band: lane l holds quantity (l & 8 ? B : A) of layer l & 7 for the bands 2 (l / 16), 2 (l / 16) + 1 of this chunk
    auto put_bands = [&](float* arr, int lev, float bA, float bB, bool add) {
      const int b0 = chunk * 4 + 2 * (lane >> 4);
      float* q = arr + bcol + (size_t)lev * p.nbnd + b0;
      if (b0 < p.nbnd) q[0] = add ? q[0] + bA : bA;
      if (b0 + 1 < p.nbnd) q[1] = add ? q[1] + bB : bB;
    };
    auto flush_fwd = [&]() {  // pend[u] = dir, pend[U + u] = diffuse + dir at the bottom of sweep layer pend_k * U + u
      float bA = 0.0f, bB = 0.0f;
      const float t = bands ? tr_reduce_bands<2 * U>(pend, tr, lane, bA, bB) : tr_reduce<2 * U>(pend, tr, lane);
      const int i = pend_k * U + ru;
      if (pend_k >= 0 && i < L) {
        const int lev = TOP ? i + 1 : L - 1 - i;
        if (rW) {
          float* dst = (rB ? fdn : fdr) + lev;
          *dst += t;
        }
        if (bands && owner) put_bands(rB ? p.bnd_dn : p.bnd_dir, lev, bA, bB, false);
      }
    };
    auto flush_bwd = [&]() {  // pend[u] = up, pend[U + u] = alpha_above * up at the top of sweep layer pend_k * U + (U - 1 - u)
      float bA = 0.0f, bB = 0.0f;
      const float t = bands ? tr_reduce_bands<2 * U>(pend, tr, lane, bA, bB) : tr_reduce<2 * U>(pend, tr, lane);
      const int i = pend_k * U + (U - 1 - ru);
      if (pend_k >= 0 && i < L) {
        const int lev = TOP ? i : L - i;
        if (rW) {
          float* dst = (rB ? fdn : fup) + lev;
          *dst += t;
        }
        // the downward by-band flux gets its second part here (alpha_above * U on top of the beta + dir of sweep 1, written by
        // another lane of this warp before the surface: ordered by the __syncwarp()s in between)
        if (bands && owner) put_bands(rB ? p.bnd_dn : p.bnd_up, lev, bA, bB, rB);
      }
    };
    // ---------------- sweep 1: top -> surface ----------------
    auto forward_group = [&](int k, auto tail_c) {
      constexpr bool TAIL = decltype(tail_c)::value;
      __syncwarp();
      issue_in(k + S - 1);
      const uint32_t nk = n_in + (uint32_t)k;
      const uint32_t st = nk % S;
      mbar_wait(bar_in + 8 * st, (nk / S) & 1u);
      const uint8_t* stg = in_ring + st * STAGE;
      const uint8_t* base = stg + lane_in;
      int shl = 0, nvalid = U;
      if (TAIL) {
        box_start<TOP, U>(lay0, k, shl);
        nvalid = min(U, L - k * U);
      }
      flush_fwd();
      uint8_t* const sg = srow + (size_t)k * (U * SW6_ROW);
      float red[2 * U];
#pragma unroll
      for (int h = 0; h < NH; ++h) {
        if (TAIL && h * H >= nvalid) {  // warp-uniform: nothing of this half belongs to the column
#pragma unroll
          for (int uu = 0; uu < H; ++uu) { red[h * H + uu] = 0.0f; red[U + h * H + uu] = 0.0f; }
          continue;
        }
        f2 tau[H], w0[H], gg[H], t2[H];
        bool cloud_here = false;
#pragma unroll
        for (int uu = 0; uu < H; ++uu) {
          const int u = h * H + uu;
          const int rl = TAIL ? box_row<TOP, U>(u, shl) : (TOP ? u : U - 1 - u);
          tau[uu] = lds2(base + rl * 256);
          w0[uu] = lds2(base + U * 256 + rl * 256);
          gg[uu] = GM == 1 ? lds2(base + 2 * U * 256 + rl * 256) : splat2(0.0f);
          if (GM == 2) {
            t2[uu] = band_pair(stg + 2 * U * 256 + rl * 192);
            cloud_here = cloud_here || lo2(t2[uu]) != 0.0f || hi2(t2[uu]) != 0.0f;
          }
        }
        f2 Rdif[H], Tdif[H], Rdir[H], Tdir[H], Tnos[H];
        // Clouds sit in a few contiguous layers of some columns: a half group none of whose 64 g-points x 4 layers sees cloud
        // (warp-uniform) keeps the gas properties as they are -- tau + 0, ssa = (tau ssa) / tau up to its last bit, g = 0
        const bool cloudy = GM == 2 && (RRNN_SW_CLD_SKIP == 0 || __any_sync(0xffffffffu, cloud_here));
        if (GM == 2 && RRNN_SW_CLD_SKIP == 1 && !cloudy) {
          two_stream2_batch<FAST, false, H>(tau, w0, gg, mu0, mu0_inv, Rdif, Tdif, Rdir, Tdir, Tnos, splat2(pp.neg_zero));
        } else {
          if (GM == 2 && cloudy) {   // inc_2stream_by_2stream_bybnd with g1 == 0, same operation order (products first, then the sum)
            const f2 eps = splat2(3.0f * 1.17549435e-38f);
#pragma unroll
            for (int uu = 0; uu < H; ++uu) {
              const int u = h * H + uu;
              const int rl = TAIL ? box_row<TOP, U>(u, shl) : (TOP ? u : U - 1 - u);
              const uint8_t* crow = stg + 2 * U * 256 + rl * 192;
              const f2 tau12 = tau[uu] + t2[uu];
              const f2 scat = tau[uu] * w0[uu] + band_pair(crow + 64);
              gg[uu] = div2<FAST>(band_pair(crow + 128), max2(eps, scat));
              w0[uu] = div2<FAST>(scat, max2(eps, tau12));
              tau[uu] = tau12;
            }
          }
          two_stream2_batch<FAST, HAS_G, H>(tau, w0, gg, mu0, mu0_inv, Rdif, Tdif, Rdir, Tdir, Tnos, splat2(pp.neg_zero));
        }
#pragma unroll
        for (int uu = 0; uu < H; ++uu) {
          const int u = h * H + uu;
          if (!TAIL || u < nvalid) {  // warp-uniform
            const f2 s_up = Rdir[uu] * dir;
            const f2 s_dn = Tdir[uu] * dir;
            dir = Tnos[uu] * dir;
            const f2 d = rcp2<FAST>(fnma2(Rdif[uu], alpha, splat2(1.0f)));
            const f2 e = d * Tdif[uu];
            const f2 f = d * fma2(Rdif[uu], beta, s_up);
            stg_scr(sg + u * SW6_ROW, e, pol_buf);
            stg_scr(sg + u * SW6_ROW + SW6_F, f, pol_buf);
            stg_scr(sg + u * SW6_ROW + SW6_A, alpha, pol_buf);  // reflectance of the atmosphere ABOVE this layer: what sweep 2 needs
            beta = fma2(e, fma2(alpha, s_up, beta), s_dn);
            alpha = fma2(Tdif[uu] * e, alpha, Rdif[uu]);
          }
          red[u] = hsum2(dir);
          red[U + u] = hsum2(beta + dir);
        }
      }
#pragma unroll
      for (int u = 0; u < 2 * U; ++u) pend[u] = red[u];
      pend_k = k;
    };
    {
      const int nfast = (TOP || col > 0) ? NGF : max(NGF - 1, 0);
      for (int k = 0; k < nfast; ++k) forward_group(k, std::false_type{});
      for (int k = nfast; k < NG; ++k) forward_group(k, std::true_type{});
    }
    flush_fwd();
    pend_k = -1;
    n_in += (uint32_t)NG;
    // ---------------- surface ----------------
    const f2 S_s = dir * a_d;  // source_sfc :1477
    f2 Uu = div2<FAST>(fma2(a_s, beta, S_s), fnma2(a_s, alpha, splat2(1.0f))) * live;
    {
      const int sfc = TOP ? L : 0;
      const float su = warp_sum(hsum2(Uu)), sa = warp_sum(hsum2(alpha * Uu));
      if (lane == 0) { fup[sfc] += su; fdn[sfc] += sa; }
      if (bands) {
        __syncwarp();   // (the forward sweep's last by-band rows, written by other lanes, are complete)
        const float bu = band_sum8(hsum2(Uu)), ba = band_sum8(hsum2(alpha * Uu));
        if (owner && (lane & 7) == 0 && my_band < p.nbnd) {
          p.bnd_up[bcol + (size_t)sfc * p.nbnd + my_band] = bu;
          p.bnd_dn[bcol + (size_t)sfc * p.nbnd + my_band] += ba;
        }
      }
    }
    // ---------------- sweep 2: surface -> top (back substitution) ----------------
    // The rows of a group come back with ONE bulk copy into the (now idle) input ring (SB stages), SB - 1 groups ahead of
    // their use; a group is pulled into registers in one go, which frees its stage for the copy of the group SB further up.
    asm volatile("fence.proxy.async.global;" ::: "memory");  // this lane's row stores (generic proxy) before the bulk loads (async proxy)
    __syncwarp();
    auto issue_bb = [&](int j) {  // j-th group of the upward sweep = forward group NG-1-j -> stage (n_bb + j) % SB
      if (j < NG) {
        const int k = NG - 1 - j;
        const uint32_t st = (n_bb + (uint32_t)j) % SB;
        const uint32_t bytes = (uint32_t)min(U, L - k * U) * SW6_ROW;
        if (elect_one()) {
          mbar_expect_tx(bar_bb + 8 * st, bytes);
          bulk_load(in_a + st * (U * SW6_ROW), srow - (size_t)lane * 8u + (size_t)k * (U * SW6_ROW), bytes, bar_bb + 8 * st, pol_buf);
        }
        __syncwarp();
      }
    };
#pragma unroll
    for (int j = 0; j < SB - 1; ++j) issue_bb(j);
    auto backward_group = [&](int j, auto tail_c) {
      constexpr bool TAIL = decltype(tail_c)::value;
      const int k = NG - 1 - j;
      const int nvalid = TAIL ? min(U, L - k * U) : U;
      __syncwarp();  // every lane has pulled the previous group into registers: its stage may be refilled
      issue_bb(j + SB - 1);
      const uint32_t nj = n_bb + (uint32_t)j;
      const uint32_t st = nj % SB;
      mbar_wait(bar_bb + 8 * st, (nj / SB) & 1u);
      f2 e[U], f[U], a[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const uint8_t* row = in_ring + st * (U * SW6_ROW) + (TAIL ? min(u, nvalid - 1) : u) * SW6_ROW + lane_in;
        e[u] = lds2(row);
        f[u] = lds2(row + SW6_F);
        a[u] = lds2(row + SW6_A);
      }
      flush_bwd();
      float red[2 * U];
#pragma unroll
      for (int u = 0; u < U; ++u) {  // sweep layers k*U + (U-1-u): upwards
        const int uu = U - 1 - u;
        if (!TAIL || uu < nvalid) Uu = fma2(e[uu], Uu, f[uu]);
        red[u] = hsum2(Uu);               // upward flux at the level on top of that layer
        red[U + u] = hsum2(a[uu] * Uu);   // diffuse downward flux there: alpha_above * U (+ beta, added in sweep 1)
      }
      // the rows are in registers: their L2 lines are dead (no write-back; the next column rewrites them in full)
      discard_scratch(srow - (size_t)lane * 8u + (size_t)k * (U * SW6_ROW), (uint32_t)nvalid * SW6_ROW, lane);
#pragma unroll
      for (int u = 0; u < 2 * U; ++u) pend[u] = red[u];
      pend_k = k;
    };
    {
      int j = 0;
      if (NG > NGF) backward_group(j++, std::true_type{});  // the ragged group comes first on the way up
      for (; j < NG; ++j) backward_group(j, std::false_type{});
    }
    n_bb += (uint32_t)NG;
    flush_bwd();
    pend_k = -1;
    __syncwarp();
    // ---- combine the chunks of this column (see lw_solver_v5)
    nx.publish(cluster, csize, ncols_done);
    cluster.sync();
    {
      float* const gout[3] = {p.flux_up + (size_t)col * (L + 1), p.flux_dn + (size_t)col * (L + 1), p.flux_dir + (size_t)col * (L + 1)};
      const int n = 3 * (L + 1), lo = chunk * n / csize, hi = (chunk + 1) * n / csize;
      for (int i = lo + lane; i < hi && owner; i += 32) {
        float sacc = 0.0f;
        for (int r = 0; r < csize; ++r) sacc += *cluster.map_shared_rank(fup + i, r);
        const int a = i / (L + 1);
        gout[a][i - a * (L + 1)] = sacc;
      }
    }
    cb = nx.next(ncols_done);
  }
  cluster.sync();
}

// ---------------------------------------------------------------------------------------------------- LW, v6
// lw_solver_v5 with the v6 machinery (see sw_solver_v6): reverse-sweep rows (t | source_up: two 256-byte segments per
// layer) stored per lane at immediate offsets, brought back by one bulk copy per group into the idle input ring and pulled
// into registers; per-level sums by transposition through shared memory; a ragged last group computes only its layers.
constexpr int LW6_ROW = 512, LW6_S = 256;

// CLD: by-band cloud optical depths whose increment has not been applied (LwParams::cld_tau) ride along as one more 64-byte
// table row per layer and are added to tau in registers -- the all-sky path never rewrites the (ngpt,nlay,ncol) array.
// BND: by-band fluxes as well (LwParams::bnd_*; a template parameter for the reason given at sw_solver_v6).
template <bool FAST, bool TOP, bool DN_EXT, bool COMPACT, bool CLD, bool BND = false>
__global__ void __launch_bounds__(32 * MAX_WARPS, RRNN_V6_LW_MINB) lw_solver_v6(const __grid_constant__ LwV5Params pp, const __grid_constant__ CUtensorMap tm_tau,
                                                   const __grid_constant__ CUtensorMap tm_lay, const __grid_constant__ CUtensorMap tm_lev,
                                                   const __grid_constant__ CUtensorMap tm_bl, const __grid_constant__ CUtensorMap tm_bv,
                                                   const __grid_constant__ CUtensorMap tm_cld) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  constexpr int U = 8, S = 2, SB = 2;
  // one stage of the input ring.  Materialised: tau | lay_source | lev_source(ext rows), U rows of 256 B each.
  // COMPACT: tau (U rows) | pfrac (PFR rows: top-down sweeps also need the next layer's) | B_lay | B_lev(ext) (U rows of 64 B)
  constexpr int PFR = COMPACT ? (TOP ? U + 1 : U) : U;
  constexpr int OFF_PF = U * 256, OFF_3 = OFF_PF + PFR * 256, OFF_BV = OFF_3 + U * 64;
  constexpr int OFF_CLD = COMPACT ? OFF_BV + U * 64 : 3 * U * 256;   // CLD: by-band cloud optical depth (U rows of 64 B)
  constexpr int STAGE = OFF_CLD + (CLD ? U * 64 : 0);
  static_assert(SB * U * LW6_ROW <= S * STAGE, "the upward sweep's stages live in the input ring");
  const LwParams& p = pp.b;
  const f2 NZ = splat2(pp.neg_zero);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;  // every warp is its own solver
  const int G = p.ngpt, L = p.nlay;
  const uint64_t pol_in = policy_evict_first();
  const uint64_t pol_buf = policy_evict_last();
  cg::cluster_group cluster = cg::this_cluster();
  const int chunk = (int)cluster.block_rank();
  const int csize = (int)cluster.num_blocks();

  uint8_t* smem = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u) + (size_t)warp * pp.warp_smem;
  uint8_t* in_ring = smem;                                            // [S][STAGE]
  float* tr = reinterpret_cast<float*>(in_ring + S * STAGE);          // [8][TR_PITCH]
  float* part = tr + 8 * TR_PITCH;                                    // [2 sets][2][L+1]
  const int part_set = 2 * (L + 1);
  uint64_t* bars = reinterpret_cast<uint64_t*>(part + 2 * part_set);
  const uint32_t bar_in = smem_u32(bars), bar_bb = smem_u32(bars + S);
  const uint32_t in_a = smem_u32(in_ring);
  if (lane == 0) {
    for (int s = 0; s < S + SB; ++s) mbar_init(bar_in + 8 * s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  uint32_t n_in = 0, n_bb = 0;  // groups consumed so far from each ring (stage = n % S, parity = (n / S) & 1)

  const float tau_thresh = 3.4526698e-4f;  // sqrt(epsilon(1._sp)), mo_rte_solver_kernels.F90:754
  const int g = chunk * 64 + 2 * lane;
  const bool act = g < G;                // ngpt is even: a pair is live or not as a whole
  const int gs = act ? g : chunk * 64;   // idle lanes shadow the chunk's first pair and contribute zero
  const float live = act ? 1.0f : 0.0f;
  const int NG = pp.ngroups;
  const int NGF = L / U;                 // full groups; the ragged one (if any) is group NGF
  // this lane's 8-byte slot in the t-segment of scratch row 0 of this solver
  uint8_t* const srow = reinterpret_cast<uint8_t*>(p.scratch) + ((size_t)blockIdx.x * nwarps + warp) * L * LW6_ROW + (size_t)lane * 8u;
  const uint32_t lane_in = (uint32_t)lane * 8u;   // byte offset of this lane's pair in a 256-byte row
  // COMPACT: byte offsets of the bands of this lane's two g-points in a 64-byte row of the Planck tables
  // (both g-points of a lane lie in one band -- LwParams::pairs_in_band, checked by the launcher -- so one look-up serves the pair)
  uint32_t bo0 = 0;
  if (COMPACT || CLD) bo0 = 4u * (uint32_t)__ldg(p.gpt2band + gs);
  auto band_pair = [&](const uint8_t* row) { return splat2(*reinterpret_cast<const float*>(row + bo0)); };
  const int ru = lane & 7;      // the reduced value (layer within its group) this lane ends up with; lanes < 8 write
  const bool rW = lane < 8;

  NextColumns nx;
  nx.slot = reinterpret_cast<int*>(smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u) + (size_t)nwarps * pp.warp_smem);
  nx.leader = chunk == 0 && threadIdx.x == 0;
  nx.fetched = 0;
  int ncols_done = 0;
  for (int cb = (blockIdx.x / csize) * nwarps; cb < p.ncol; ++ncols_done) {
    nx.begin(pp.next_col, nwarps, (int)(gridDim.x / csize) * nwarps);
    const bool owner = cb + warp < p.ncol;
    const int col = owner ? cb + warp : p.ncol - 1;
    float* fup = part + (ncols_done & 1) * part_set;  // this column's partial fluxes [2][L+1]
    float* fdn = fup + (L + 1);
    for (int i = lane; i < 2 * (L + 1); i += 32) fup[i] = 0.0f;
    const size_t gc_off = (size_t)col * G + gs;
    const f2 emis = ldg2(p.sfc_emis + gc_off);
    const f2 ssrc = ldg2(p.sfc_source + gc_off);
    const f2 inc = p.inc_flux ? ldg2(p.inc_flux + gc_off) : splat2(0.0f);
    // tensor rows of sweep layer 0: layers (tau, lay_source) and the level towards the surface (lev_source)
    const int lay0 = col * L + (TOP ? 0 : L - 1);
    const int ext0 = col * (L + 1) + (TOP ? 1 : L - 1);
    // lev_source at the level where the sweep enters the atmosphere
    f2 ent0;
    if (COMPACT) {
      const float* bv0 = p.planck_lev + ((size_t)col * (L + 1) + (TOP ? 0 : L)) * 16;
      ent0 = mul_keep(ldg2(p.lay_source + ((size_t)col * L + (TOP ? 0 : L - 1)) * G + gs), splat2(__ldg(bv0 + (bo0 >> 2))), NZ);
    } else {
      ent0 = ldg2(p.lev_source + ((size_t)col * (L + 1) + (TOP ? 0 : L)) * G + gs);
    }
    __syncwarp();

    for (int imu = 0; imu < p.nmus; ++imu) {
      const f2 D = splat2(p.Ds[imu]);
      // Idle lanes (ngpt not a multiple of 64: the ragged last chunk) read the zero fill of the TMA boxes beyond ngpt: tau = 0,
      // sources = 0, so their radiances stay exactly zero once the two places that feed them shadow data -- the incident flux and
      // the surface emission -- are masked.  No per-layer masking.
      const float rad_norm = 2.0f * kPi * p.wts[imu];
      f2 I = map2(inc, [&](float v) { return v / rad_norm; }) * splat2(live);  // radn_dn(top) = inc_flux/(2 pi w), :196-201
      {
        const float s = warp_sum(hsum2(I));
        if (lane == 0) fdn[TOP ? 0 : L] += rad_norm * s;
      }
      // by-band outputs (warp-uniform): this lane's band = chunk * 4 + lane / 8; later quadrature angles add to the first one's
      const bool bands = p.bnd_up != nullptr;
      const int my_band = chunk * 4 + (lane >> 3);
      const size_t bcol = (size_t)col * (L + 1) * (size_t)p.nbnd;
      // quirk Q3 (mo_rte_solver_kernels.F90:284-291): with ONE quadrature angle the reference leaves its g-point "fluxes" as radiances,
      // un-multiplied by 2 pi w (only the inlined broadband sum is scaled), and ty_fluxes_byband reduces exactly those
      const float band_norm = p.nmus == 1 ? 1.0f : rad_norm;
      auto put_band = [&](float* arr, int lev, float b) {
        if (owner && my_band < p.nbnd) {
          float* q = arr + bcol + (size_t)lev * p.nbnd + my_band;
          *q = imu == 0 ? band_norm * b : *q + band_norm * b;
        }
      };
      if (bands) {
        const float b = band_sum8(hsum2(I));
        if ((lane & 7) == 0) put_band(p.bnd_dn, TOP ? 0 : L, b);
      }
      // one elected lane feeds the input ring: group k -> stage (n_in + k) % S
      auto issue_in = [&](int k) {
        if (k < NG) {
          const uint32_t st = (n_in + (uint32_t)k) % S;
          int sh;
          const int rl = box_start<TOP, U>(lay0, k, sh), rv = box_start<TOP, U>(ext0, k, sh);
          if (elect_one()) {
            const uint32_t bar = bar_in + 8 * st;
            const uint32_t dst = in_a + st * STAGE;
            mbar_expect_tx(bar, STAGE);
            tma_load_2d(dst, &tm_tau, chunk * 64, rl, bar, pol_in);
            tma_load_2d(dst + OFF_PF, &tm_lay, chunk * 64, rl, bar, pol_in);
            if (COMPACT) {
              tma_load_2d(dst + OFF_3, &tm_bl, 0, rl, bar, pol_in);
              tma_load_2d(dst + OFF_BV, &tm_bv, 0, rv, bar, pol_in);
            } else {
              tma_load_2d(dst + OFF_3, &tm_lev, chunk * 64, rv, bar, pol_in);
            }
            if (CLD) tma_load_2d(dst + OFF_CLD, &tm_cld, 0, rl, bar, pol_in);
          }
          __syncwarp();
        }
      };
      f2 carry = ent0;  // ent(0)
      // (the upward sweep of the previous angle / column read its last rows from the ring: refill it only now)
      __syncwarp();
#pragma unroll
      for (int k = 0; k < S - 1; ++k) issue_in(k);
      // per-level broadband sums of a group: reduced one group later (their latency overlaps the next group's arithmetic)
      float pend[U];
#pragma unroll
      for (int u = 0; u < U; ++u) pend[u] = 0.0f;
      int pend_k = -1;
      // (the quadrature factor 2 pi w multiplies the REDUCED sums: one multiply per level instead of one per lane and layer)
      auto flush_dn = [&]() {   // by-band: lane l holds layer l & 7 of band l / 8 of this chunk
        float bA = 0.0f, bB = 0.0f;
        const float t = bands ? tr_reduce_bands<U>(pend, tr, lane, bA, bB) : tr_reduce<U>(pend, tr, lane);
        const int i = pend_k * U + ru;
        if (pend_k >= 0 && i < L) {
          if (rW) fdn[TOP ? i + 1 : L - 1 - i] += rad_norm * t;
          if (bands) put_band(p.bnd_dn, TOP ? i + 1 : L - 1 - i, bA);
        }
      };
      auto flush_up = [&]() {
        float bA = 0.0f, bB = 0.0f;
        const float t = bands ? tr_reduce_bands<U>(pend, tr, lane, bA, bB) : tr_reduce<U>(pend, tr, lane);
        const int i = pend_k * U + (U - 1 - ru);
        if (pend_k >= 0 && i < L) {
          if (rW) fup[TOP ? i : L - i] += rad_norm * t;
          if (bands) put_band(p.bnd_up, TOP ? i : L - i, bA);
        }
      };
      // ---------------- downward sweep: one group of U layers ----------------
      // TAIL = false: a full group whose boxes sit where box_start put them (immediate shared-memory offsets);
      // TAIL = true: the ragged last group (nvalid < U) and/or a box that was moved (column 0, bottom-up)
      auto forward_group = [&](int k, auto tail_c) {
        constexpr bool TAIL = decltype(tail_c)::value;
        __syncwarp();                 // every lane is done with the stage that group k+S-1 overwrites
        issue_in(k + S - 1);
        const uint32_t nk = n_in + (uint32_t)k;
        const uint32_t st = nk % S;
        mbar_wait(bar_in + 8 * st, (nk / S) & 1u);
        const uint8_t* stg = in_ring + st * STAGE;
        const uint8_t* base = stg + lane_in;
        int shl = 0, shv = 0, nvalid = U;
        if (TAIL) {
          box_start<TOP, U>(lay0, k, shl);
          box_start<TOP, U>(ext0, k, shv);
          nvalid = min(U, L - k * U);
        }
        f2 tau[U], lay[U], ext[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          if (TAIL && u >= nvalid) { tau[u] = splat2(1.0f); lay[u] = splat2(0.0f); ext[u] = splat2(0.0f); continue; }  // warp-uniform
          const int rl = TAIL ? box_row<TOP, U>(u, shl) : (TOP ? u : U - 1 - u);
          const int rv = TAIL ? box_row<TOP, U>(u, shv) : (TOP ? u : U - 1 - u);
          tau[u] = lds2(base + rl * 256);
          if (CLD) tau[u] = tau[u] + band_pair(stg + OFF_CLD + rl * 64);   // inc_1scalar_by_1scalar_bybnd
          if (COMPACT) {
            const f2 pf = lds2(base + OFF_PF + rl * 256);
            lay[u] = mul_keep(pf, band_pair(stg + OFF_3 + rl * 64), NZ);
            // the level below the bottom layer takes that layer's fraction (:667-669); bottom-up sweeps leave through
            // the layer's own level
            const f2 pfx = (TOP && k * U + u != L - 1) ? lds2(base + OFF_PF + (rl + 1) * 256) : pf;
            ext[u] = mul_keep(pfx, band_pair(stg + OFF_BV + rv * 64), NZ);
          } else {
            lay[u] = lds2(base + OFF_PF + rl * 256);
            ext[u] = lds2(base + OFF_3 + rv * 256);
          }
        }
        flush_dn();
        uint8_t* const sg = srow + (size_t)k * (U * LW6_ROW);
        float red[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          if (TAIL && u >= nvalid) { red[u] = 0.0f; continue; }  // warp-uniform: this layer is not part of the column
          const f2 ent = (u == 0) ? carry : ext[u - 1];
          const f2 tl = tau[u] * D;
          f2 t, omt;
          exp_and_complement2<FAST>(tl, t, omt);
          // fact = (1-t)/tau' - t, or its series where tau' is tiny (:757-768)
          const f2 fa = div2<true>(omt, tl) - t;
          const f2 fb = tl * fnma2(tl, splat2(1.0f / 3.0f), splat2(0.5f));
          float tx, ty;
          unpack2(tl, tx, ty);
          const f2 fact = sel2(tx > tau_thresh, ty > tau_thresh, fa, fb);
          const f2 f2x = fact + fact;
          // lw_source_noscat (:770-773): source_dn from lev(l+1), source_up from lev(l) whatever the orientation (quirk Q1)
          const f2 lev_dn = DN_EXT ? ext[u] : ent;
          const f2 lev_up = DN_EXT ? ent : ext[u];
          const f2 sdn = fma2(f2x, lay[u] - lev_dn, omt * lev_dn);
          const f2 sup = fma2(f2x, lay[u] - lev_up, omt * lev_up);
          stg_scr(sg + u * LW6_ROW, t, pol_buf);
          stg_scr(sg + u * LW6_ROW + LW6_S, sup, pol_buf);
          I = fma2(t, I, sdn);
          red[u] = hsum2(I);
        }
        carry = ext[U - 1];  // (a ragged group is the last one of its sweep: its carry is not used)
#pragma unroll
        for (int u = 0; u < U; ++u) pend[u] = red[u];
        pend_k = k;
      };
      {
        // bottom-up, column 0: the boxes of the last groups may have been moved -> generic path for those
        const int nfast = (TOP || col > 0) ? NGF : max(NGF - 1, 0);
        for (int k = 0; k < nfast; ++k) forward_group(k, std::false_type{});
        for (int k = nfast; k < NG; ++k) forward_group(k, std::true_type{});
      }
      flush_dn();
      pend_k = -1;
      n_in += (uint32_t)NG;
      // ---------------- surface ----------------
      f2 Uu = fma2(I, splat2(1.0f) - emis, emis * ssrc) * splat2(live);  // :269
      {
        const float s = warp_sum(hsum2(Uu));
        if (lane == 0) fup[TOP ? L : 0] += rad_norm * s;
        if (bands) {
          const float b = band_sum8(hsum2(Uu));
          if ((lane & 7) == 0) put_band(p.bnd_up, TOP ? L : 0, b);
        }
      }
      // ---------------- upward sweep (reverse order): rows back by bulk copies into the idle input ring ----------------
      asm volatile("fence.proxy.async.global;" ::: "memory");  // this lane's row stores (generic proxy) before the bulk loads (async proxy)
      __syncwarp();
      auto issue_bb = [&](int j) {  // j-th group of the upward sweep = forward group NG-1-j
        if (j < NG) {
          const int k = NG - 1 - j;
          const uint32_t st = (n_bb + (uint32_t)j) % SB;
          const uint32_t bytes = (uint32_t)min(U, L - k * U) * LW6_ROW;
          if (elect_one()) {
            mbar_expect_tx(bar_bb + 8 * st, bytes);
            bulk_load(in_a + st * (U * LW6_ROW), srow - (size_t)lane * 8u + (size_t)k * (U * LW6_ROW), bytes, bar_bb + 8 * st, pol_buf);
          }
          __syncwarp();
        }
      };
#pragma unroll
      for (int j = 0; j < SB - 1; ++j) issue_bb(j);
      auto backward_group = [&](int j, auto tail_c) {
        constexpr bool TAIL = decltype(tail_c)::value;
        const int k = NG - 1 - j;
        const int nvalid = TAIL ? min(U, L - k * U) : U;
        __syncwarp();  // every lane has pulled the previous group into registers: its stage may be refilled
        issue_bb(j + SB - 1);
        const uint32_t nj = n_bb + (uint32_t)j;
        const uint32_t st = nj % SB;
        mbar_wait(bar_bb + 8 * st, (nj / SB) & 1u);
        f2 t[U], s[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const uint8_t* row = in_ring + st * (U * LW6_ROW) + (TAIL ? min(u, nvalid - 1) : u) * LW6_ROW + lane_in;
          t[u] = lds2(row);
          s[u] = lds2(row + LW6_S);
        }
        flush_up();
        float red[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {  // sweep layers k*U + (U-1-u): upwards
          const int uu = U - 1 - u;
          if (!TAIL || uu < nvalid) Uu = fma2(t[uu], Uu, s[uu]);
          red[u] = hsum2(Uu);
        }
        // the rows are in registers: their L2 lines are dead (no write-back; the next sweep rewrites them in full)
        discard_scratch(srow - (size_t)lane * 8u + (size_t)k * (U * LW6_ROW), (uint32_t)nvalid * LW6_ROW, lane);
#pragma unroll
        for (int u = 0; u < U; ++u) pend[u] = red[u];
        pend_k = k;
      };
      {
        int j = 0;
        if (NG > NGF) backward_group(j++, std::true_type{});  // the ragged group comes first on the way up
        for (; j < NG; ++j) backward_group(j, std::false_type{});
      }
      flush_up();
      pend_k = -1;
      n_bb += (uint32_t)NG;
      __syncwarp();
    }
    // ---- combine the chunks of this column (see lw_solver_v5)
    nx.publish(cluster, csize, ncols_done);
    cluster.sync();
    {
      float* const gout[2] = {p.flux_up + (size_t)col * (L + 1), p.flux_dn + (size_t)col * (L + 1)};
      const int n = 2 * (L + 1), lo = chunk * n / csize, hi = (chunk + 1) * n / csize;
      for (int i = lo + lane; i < hi && owner; i += 32) {
        float sacc = 0.0f;
        for (int r = 0; r < csize; ++r) sacc += *cluster.map_shared_rank(fup + i, r);
        const int a = i / (L + 1);
        gout[a][i - a * (L + 1)] = sacc;
      }
    }
    cb = nx.next(ncols_done);
  }
  cluster.sync();  // nobody leaves while another rank may still read its shared memory
}

// ---------------------------------------------------------------------------------------------------- LW, v7 ("wide")
// lw_solver_v6 with FOUR ADJACENT g-points per lane: one warp solves 128 g-points of a column, the ceil(ngpt/128) chunk-CTAs of a
// column form the cluster.  Why: lw_solver_v6 issues ~100 warp instructions per (layer, 64 g-points) of which ~43 are the
// arithmetic (profiles/r2y_lw_solver_by_source_line.txt); the rest -- TMA issue, mbarrier waits, loop control, the transposition
// sums, shared-memory loads, scratch stores -- is per WARP and layer, not per g-point.  With four g-points per lane that
// machinery is spent once per 128 g-points, loads / stores are 16 bytes wide (LDS.128, STG.128: half the memory instructions),
// one band look-up serves four g-points, and every lane carries two independent packed chains (the latency hiding that
// v6 gets from a second resident warp).  The per-g-point arithmetic is v6's, instruction for instruction; only the summation
// order over g-points differs (in-lane (a + b) packed, then the same transposition; two cluster ranks instead of four).
// Plain clear-sky kernel: clouds and by-band outputs stay on v6.
struct f4 { f2 a, b; };
__device__ __forceinline__ f4 lds4(const void* p) { f4 v; lds22(p, v.a, v.b); return v; }
__device__ __forceinline__ f4 ldg4(const float* p) {
  f4 v;
  asm volatile("ld.global.nc.v2.b64 {%0, %1}, [%2];" : "=l"(v.a.v), "=l"(v.b.v) : "l"(p));
  return v;
}
__device__ __forceinline__ void stg_scr4(uint8_t* p, f4 v, uint64_t pol) {
  asm volatile("st.global.L1::no_allocate.L2::cache_hint.v2.b64 [%0], {%1, %2}, %3;" ::"l"(p), "l"(v.a.v), "l"(v.b.v), "l"(pol) : "memory");
}
__device__ __forceinline__ f4 splat4(float x) { f4 v; v.a = splat2(x); v.b = v.a; return v; }
__device__ __forceinline__ f4 mul4(f4 x, f2 s, f2 nz) { f4 v; v.a = mul_keep(x.a, s, nz); v.b = mul_keep(x.b, s, nz); return v; }   // (the factored sources: see mul_keep)
__device__ __forceinline__ float hsum4(f4 x) { return hsum2(x.a + x.b); }

// one layer of the downward sweep for a pair of g-points (lw_source_noscat + lw_transport_noscat_dn, :742-776, 950-965):
// the arithmetic of lw_solver_v6's forward loop, unchanged
template <bool FAST, bool DN_EXT>
__device__ __forceinline__ void lw_layer_dn2(f2 tau, f2 lay, f2 ext, f2 ent, f2 D, f2& I, f2& t, f2& sup) {
  const float tau_thresh = 3.4526698e-4f;  // sqrt(epsilon(1._sp)), mo_rte_solver_kernels.F90:754
  const f2 tl = tau * D;
  f2 omt;
  exp_and_complement2<FAST>(tl, t, omt);
  const f2 fa = div2<true>(omt, tl) - t;
  const f2 fb = tl * fnma2(tl, splat2(1.0f / 3.0f), splat2(0.5f));
  float tx, ty;
  unpack2(tl, tx, ty);
  const f2 fact = sel2(tx > tau_thresh, ty > tau_thresh, fa, fb);
  const f2 f2x = fact + fact;
  const f2 lev_dn = DN_EXT ? ext : ent;
  const f2 lev_up = DN_EXT ? ent : ext;
  const f2 sdn = fma2(f2x, lay - lev_dn, omt * lev_dn);
  sup = fma2(f2x, lay - lev_up, omt * lev_up);
  I = fma2(t, I, sdn);
}

constexpr int LW7_ROW = 1024, LW7_S = 512;   // reverse-sweep row of one layer: t | source_up, 32 lanes x 16 B each
#ifndef RRNN_V7_LW_S
#define RRNN_V7_LW_S 2      // stages of the input ring
#endif
#ifndef RRNN_V7_LW_SB
#define RRNN_V7_LW_SB 2     // stages of the upward sweep's ring (bulk copies of reverse-sweep rows in flight ahead of their use)
#endif
#ifndef RRNN_V7_LW_MINB
#define RRNN_V7_LW_MINB 2   // 2 CTAs of 128 threads: the shared memory (23 KB per solver) allows 8 - 9 solver warps per SM
#endif

// CLD / BND: as in lw_solver_v6 (pending by-band cloud optical depths added in registers; by-band fluxes on the way to the broadband
// sums -- here a reader of the transposition adds 8 lanes = 32 g-points = TWO bands, so it hands out both halves).
template <bool FAST, bool TOP, bool DN_EXT, bool COMPACT, bool CLD = false, bool BND = false>
__global__ void __launch_bounds__(32 * MAX_WARPS, RRNN_V7_LW_MINB) lw_solver_v7(const __grid_constant__ LwV5Params pp, const __grid_constant__ CUtensorMap tm_tau,
                                                   const __grid_constant__ CUtensorMap tm_lay, const __grid_constant__ CUtensorMap tm_lev,
                                                   const __grid_constant__ CUtensorMap tm_bl, const __grid_constant__ CUtensorMap tm_bv,
                                                   const __grid_constant__ CUtensorMap tm_cld) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  constexpr int U = 8, S = RRNN_V7_LW_S, SB = RRNN_V7_LW_SB, H = 4, NH = U / H, RB = 512;   // groups of 8 layers (one TMA box, rows of 128 g-points), computed in halves of 4
  constexpr int PFR = COMPACT ? (TOP ? U + 1 : U) : U;
  constexpr int OFF_PF = U * RB, OFF_3 = OFF_PF + PFR * RB, OFF_BV = OFF_3 + U * 64;
  constexpr int OFF_CLD = COMPACT ? OFF_BV + U * 64 : 3 * U * RB;   // CLD: by-band cloud optical depth (U rows of 64 B)
  constexpr int STAGE = OFF_CLD + (CLD ? U * 64 : 0);
  static_assert(SB * U * LW7_ROW <= S * STAGE, "the upward sweep's stages live in the input ring");
  const LwParams& p = pp.b;
  const f2 NZ = splat2(pp.neg_zero);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;  // every warp is its own solver
  const int G = p.ngpt, L = p.nlay;
  const uint64_t pol_in = policy_evict_first();
  const uint64_t pol_buf = policy_evict_last();
  cg::cluster_group cluster = cg::this_cluster();
  const int chunk = (int)cluster.block_rank();
  const int csize = (int)cluster.num_blocks();

  uint8_t* smem = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u) + (size_t)warp * pp.warp_smem;
  uint8_t* in_ring = smem;                                            // [S][STAGE]
  float* tr = reinterpret_cast<float*>(in_ring + S * STAGE);          // [8][TR_PITCH]
  float* part = tr + 8 * TR_PITCH;                                    // [2 sets][2][L+1]
  const int part_set = 2 * (L + 1);
  uint64_t* bars = reinterpret_cast<uint64_t*>(part + 2 * part_set);
  const uint32_t bar_in = smem_u32(bars), bar_bb = smem_u32(bars + S);
  const uint32_t in_a = smem_u32(in_ring);
  if (lane == 0) {
    for (int s = 0; s < S + SB; ++s) mbar_init(bar_in + 8 * s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  uint32_t n_in = 0, n_bb = 0;  // groups consumed so far from each ring (stage = n % S, parity = (n / S) & 1)

  const int g = chunk * 128 + 4 * lane;
  const bool act = g < G;                 // ngpt is a multiple of 4: a lane's four g-points are live or not as a whole
  const int gs = act ? g : chunk * 128;   // idle lanes shadow the chunk's first quad and contribute zero
  const float live = act ? 1.0f : 0.0f;
  const int NG = pp.ngroups;
  const int NGF = L / U;                  // full groups; the ragged one (if any) is group NGF
  // this lane's 16-byte slot in the t-segment of scratch row 0 of this solver
  uint8_t* const srow = reinterpret_cast<uint8_t*>(p.scratch) + ((size_t)blockIdx.x * nwarps + warp) * L * LW7_ROW + (size_t)lane * 16u;
  const uint32_t lane_in = (uint32_t)lane * 16u;   // byte offset of this lane's four g-points in a 512-byte row
  // COMPACT: byte offset of the band of this lane's g-points in a 64-byte row of the Planck tables (all four lie in one band:
  // LwParams::pairs_in_band == 2, checked by the launcher)
  uint32_t bo0 = 0;
  if (COMPACT || CLD) bo0 = 4u * (uint32_t)__ldg(p.gpt2band + gs);
  auto band_val = [&](const uint8_t* row) { return splat2(*reinterpret_cast<const float*>(row + bo0)); };
  const int ru = lane & 7;      // the reduced value (layer within its group) this lane ends up with; lanes < 8 write
  const bool rW = lane < 8;

  NextColumns nx;
  nx.slot = reinterpret_cast<int*>(smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u) + (size_t)nwarps * pp.warp_smem);
  nx.leader = chunk == 0 && threadIdx.x == 0;
  nx.fetched = 0;
  int ncols_done = 0;
  for (int cb = (blockIdx.x / csize) * nwarps; cb < p.ncol; ++ncols_done) {
    nx.begin(pp.next_col, nwarps, (int)(gridDim.x / csize) * nwarps);
    const bool owner = cb + warp < p.ncol;
    const int col = owner ? cb + warp : p.ncol - 1;
    float* fup = part + (ncols_done & 1) * part_set;  // this column's partial fluxes [2][L+1]
    float* fdn = fup + (L + 1);
    for (int i = lane; i < 2 * (L + 1); i += 32) fup[i] = 0.0f;
    const size_t gc_off = (size_t)col * G + gs;
    // (the per-column arrays need only be 8-byte aligned, as for lw_solver_v6: two 8-byte loads)
    auto ldg4u = [](const float* q) { f4 v; v.a = ldg2(q); v.b = ldg2(q + 2); return v; };
    const f4 emis = ldg4u(p.sfc_emis + gc_off);
    const f4 ssrc = ldg4u(p.sfc_source + gc_off);
    const f4 inc = p.inc_flux ? ldg4u(p.inc_flux + gc_off) : splat4(0.0f);
    // tensor rows of sweep layer 0: layers (tau, lay_source) and the level towards the surface (lev_source)
    const int lay0 = col * L + (TOP ? 0 : L - 1);
    const int ext0 = col * (L + 1) + (TOP ? 1 : L - 1);
    // lev_source at the level where the sweep enters the atmosphere
    f4 ent0;
    if (COMPACT) {
      const float* bv0 = p.planck_lev + ((size_t)col * (L + 1) + (TOP ? 0 : L)) * 16;
      ent0 = mul4(ldg4(p.lay_source + ((size_t)col * L + (TOP ? 0 : L - 1)) * G + gs), splat2(__ldg(bv0 + (bo0 >> 2))), NZ);
    } else {
      ent0 = ldg4(p.lev_source + ((size_t)col * (L + 1) + (TOP ? 0 : L)) * G + gs);
    }
    __syncwarp();

    for (int imu = 0; imu < p.nmus; ++imu) {
      const f2 D = splat2(p.Ds[imu]);
      // (idle lanes of a ragged chunk read the zero fill of the TMA boxes beyond ngpt and stay exactly zero: see lw_solver_v6)
      const float rad_norm = 2.0f * kPi * p.wts[imu];
      f4 I;   // radn_dn(top) = inc_flux/(2 pi w), :196-201
      I.a = map2(inc.a, [&](float v) { return v / rad_norm; }) * splat2(live);
      I.b = map2(inc.b, [&](float v) { return v / rad_norm; }) * splat2(live);
      {
        const float s = warp_sum(hsum4(I));
        if (lane == 0) fdn[TOP ? 0 : L] += rad_norm * s;
      }
      // by-band outputs: a band = 16 g-points = 4 lanes; later quadrature angles add to the first one's; quirk Q3 as in lw_solver_v6
      const size_t bcol = (size_t)col * (L + 1) * (size_t)p.nbnd;
      const float band_norm = p.nmus == 1 ? 1.0f : rad_norm;
      auto put_band = [&](float* arr, int lev, int band, float b) {
        if (owner && band < p.nbnd) {
          float* q = arr + bcol + (size_t)lev * p.nbnd + band;
          *q = imu == 0 ? band_norm * b : *q + band_norm * b;
        }
      };
      auto band_sum4 = [&](float v) { v += __shfl_xor_sync(0xffffffffu, v, 1); v += __shfl_xor_sync(0xffffffffu, v, 2); return v; };
      if (BND) {
        const float b = band_sum4(hsum4(I));
        if ((lane & 3) == 0) put_band(p.bnd_dn, TOP ? 0 : L, chunk * 8 + (lane >> 2), b);
      }
      // one elected lane feeds the input ring: group k -> stage (n_in + k) % S
      auto issue_in = [&](int k) {
        if (k < NG) {
          const uint32_t st = (n_in + (uint32_t)k) % S;
          int sh;
          const int rl = box_start<TOP, U>(lay0, k, sh), rv = box_start<TOP, U>(ext0, k, sh);
          if (elect_one()) {
            const uint32_t bar = bar_in + 8 * st;
            const uint32_t dst = in_a + st * STAGE;
            mbar_expect_tx(bar, STAGE);
            if (CLD) tma_load_2d(dst + OFF_CLD, &tm_cld, 0, rl, bar, pol_in);
            tma_load_2d(dst, &tm_tau, chunk * 128, rl, bar, pol_in);
            tma_load_2d(dst + OFF_PF, &tm_lay, chunk * 128, rl, bar, pol_in);
            if (COMPACT) {
              tma_load_2d(dst + OFF_3, &tm_bl, 0, rl, bar, pol_in);
              tma_load_2d(dst + OFF_BV, &tm_bv, 0, rv, bar, pol_in);
            } else {
              tma_load_2d(dst + OFF_3, &tm_lev, chunk * 128, rv, bar, pol_in);
            }
          }
          __syncwarp();
        }
      };
      f4 carry = ent0;  // ent(0)
      // (the upward sweep of the previous angle / column read its last rows from the ring: refill it only now)
      __syncwarp();
#pragma unroll
      for (int k = 0; k < S - 1; ++k) issue_in(k);
      // per-level broadband sums of a group: reduced one group later (their latency overlaps the next group's arithmetic)
      float pend[U];
#pragma unroll
      for (int u = 0; u < U; ++u) pend[u] = 0.0f;
      int pend_k = -1;
      // by-band: lane l holds layer l & 7 of the bands chunk * 8 + 2 (l / 8) and + 1
      auto flush_dn = [&]() {
        float bA = 0.0f, bB = 0.0f;
        const float t = BND ? tr_reduce_bands_wide(pend, tr, lane, bA, bB) : tr_reduce<U>(pend, tr, lane);
        const int i = pend_k * U + ru;
        if (pend_k >= 0 && i < L) {
          if (rW) fdn[TOP ? i + 1 : L - 1 - i] += rad_norm * t;
          if (BND) { put_band(p.bnd_dn, TOP ? i + 1 : L - 1 - i, chunk * 8 + 2 * (lane >> 3), bA); put_band(p.bnd_dn, TOP ? i + 1 : L - 1 - i, chunk * 8 + 2 * (lane >> 3) + 1, bB); }
        }
      };
      auto flush_up = [&]() {
        float bA = 0.0f, bB = 0.0f;
        const float t = BND ? tr_reduce_bands_wide(pend, tr, lane, bA, bB) : tr_reduce<U>(pend, tr, lane);
        const int i = pend_k * U + (U - 1 - ru);
        if (pend_k >= 0 && i < L) {
          if (rW) fup[TOP ? i : L - i] += rad_norm * t;
          if (BND) { put_band(p.bnd_up, TOP ? i : L - i, chunk * 8 + 2 * (lane >> 3), bA); put_band(p.bnd_up, TOP ? i : L - i, chunk * 8 + 2 * (lane >> 3) + 1, bB); }
        }
      };
      // ---------------- downward sweep: one group of U layers, in halves of H ----------------
      auto forward_group = [&](int k, auto tail_c) {
        constexpr bool TAIL = decltype(tail_c)::value;
        __syncwarp();                 // every lane is done with the stage that group k+S-1 overwrites
        issue_in(k + S - 1);
        const uint32_t nk = n_in + (uint32_t)k;
        const uint32_t st = nk % S;
        mbar_wait(bar_in + 8 * st, (nk / S) & 1u);
        const uint8_t* stg = in_ring + st * STAGE;
        const uint8_t* base = stg + lane_in;
        int shl = 0, shv = 0, nvalid = U;
        if (TAIL) {
          box_start<TOP, U>(lay0, k, shl);
          box_start<TOP, U>(ext0, k, shv);
          nvalid = min(U, L - k * U);
        }
        flush_dn();
        uint8_t* const sg = srow + (size_t)k * (U * LW7_ROW);
        float red[U];
#pragma unroll
        for (int h = 0; h < NH; ++h) {
          if (TAIL && h * H >= nvalid) {   // warp-uniform: nothing of this half belongs to the column
#pragma unroll
            for (int uu = 0; uu < H; ++uu) red[h * H + uu] = 0.0f;
            continue;
          }
          f4 tau[H], lay[H], ext[H];
#pragma unroll
          for (int uu = 0; uu < H; ++uu) {
            const int u = h * H + uu;
            if (TAIL && u >= nvalid) { tau[uu] = splat4(1.0f); lay[uu] = splat4(0.0f); ext[uu] = splat4(0.0f); continue; }  // warp-uniform
            const int rl = TAIL ? box_row<TOP, U>(u, shl) : (TOP ? u : U - 1 - u);
            const int rv = TAIL ? box_row<TOP, U>(u, shv) : (TOP ? u : U - 1 - u);
            tau[uu] = lds4(base + rl * RB);
            if (CLD) { const f2 tc = band_val(stg + OFF_CLD + rl * 64); tau[uu].a = tau[uu].a + tc; tau[uu].b = tau[uu].b + tc; }   // inc_1scalar_by_1scalar_bybnd
            if (COMPACT) {
              const f4 pf = lds4(base + OFF_PF + rl * RB);
              lay[uu] = mul4(pf, band_val(stg + OFF_3 + rl * 64), NZ);
              // the level below the bottom layer takes that layer's fraction (:667-669); bottom-up sweeps leave through
              // the layer's own level
              const f4 pfx = (TOP && k * U + u != L - 1) ? lds4(base + OFF_PF + (rl + 1) * RB) : pf;
              ext[uu] = mul4(pfx, band_val(stg + OFF_BV + rv * 64), NZ);
            } else {
              lay[uu] = lds4(base + OFF_PF + rl * RB);
              ext[uu] = lds4(base + OFF_3 + rv * RB);
            }
          }
#pragma unroll
          for (int uu = 0; uu < H; ++uu) {
            const int u = h * H + uu;
            if (TAIL && u >= nvalid) { red[u] = 0.0f; continue; }  // warp-uniform: this layer is not part of the column
            const f4 ent = (uu == 0) ? carry : ext[uu - 1];
            f4 t, sup;
            lw_layer_dn2<FAST, DN_EXT>(tau[uu].a, lay[uu].a, ext[uu].a, ent.a, D, I.a, t.a, sup.a);
            lw_layer_dn2<FAST, DN_EXT>(tau[uu].b, lay[uu].b, ext[uu].b, ent.b, D, I.b, t.b, sup.b);
            stg_scr4(sg + u * LW7_ROW, t, pol_buf);
            stg_scr4(sg + u * LW7_ROW + LW7_S, sup, pol_buf);
            red[u] = hsum4(I);
          }
          carry = ext[H - 1];  // (a ragged half is the last one of its sweep: its carry is not used)
        }
#pragma unroll
        for (int u = 0; u < U; ++u) pend[u] = red[u];
        pend_k = k;
      };
      {
        // bottom-up, column 0: the boxes of the last groups may have been moved -> generic path for those
        const int nfast = (TOP || col > 0) ? NGF : max(NGF - 1, 0);
        for (int k = 0; k < nfast; ++k) forward_group(k, std::false_type{});
        for (int k = nfast; k < NG; ++k) forward_group(k, std::true_type{});
      }
      flush_dn();
      pend_k = -1;
      n_in += (uint32_t)NG;
      // ---------------- surface ----------------
      f4 Uu;   // :269
      Uu.a = fma2(I.a, splat2(1.0f) - emis.a, emis.a * ssrc.a) * splat2(live);
      Uu.b = fma2(I.b, splat2(1.0f) - emis.b, emis.b * ssrc.b) * splat2(live);
      {
        const float s = warp_sum(hsum4(Uu));
        if (lane == 0) fup[TOP ? L : 0] += rad_norm * s;
        if (BND) {
          const float b = band_sum4(hsum4(Uu));
          if ((lane & 3) == 0) put_band(p.bnd_up, TOP ? L : 0, chunk * 8 + (lane >> 2), b);
        }
      }
      // ---------------- upward sweep (reverse order): rows back by bulk copies into the idle input ring ----------------
      asm volatile("fence.proxy.async.global;" ::: "memory");  // this lane's row stores (generic proxy) before the bulk loads (async proxy)
      __syncwarp();
      auto issue_bb = [&](int j) {  // j-th group of the upward sweep = forward group NG-1-j
        if (j < NG) {
          const int k = NG - 1 - j;
          const uint32_t st = (n_bb + (uint32_t)j) % SB;
          const uint32_t bytes = (uint32_t)min(U, L - k * U) * LW7_ROW;
          if (elect_one()) {
            mbar_expect_tx(bar_bb + 8 * st, bytes);
            bulk_load(in_a + st * (U * LW7_ROW), srow - (size_t)lane * 16u + (size_t)k * (U * LW7_ROW), bytes, bar_bb + 8 * st, pol_buf);
          }
          __syncwarp();
        }
      };
#pragma unroll
      for (int j = 0; j < SB - 1; ++j) issue_bb(j);
      auto backward_group = [&](int j, auto tail_c) {
        constexpr bool TAIL = decltype(tail_c)::value;
        const int k = NG - 1 - j;
        const int nvalid = TAIL ? min(U, L - k * U) : U;
        __syncwarp();  // every lane has pulled the previous group into registers: its stage may be refilled
        issue_bb(j + SB - 1);
        const uint32_t nj = n_bb + (uint32_t)j;
        const uint32_t st = nj % SB;
        mbar_wait(bar_bb + 8 * st, (nj / SB) & 1u);
        f4 t[U], s[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const uint8_t* row = in_ring + st * (U * LW7_ROW) + (TAIL ? min(u, nvalid - 1) : u) * LW7_ROW + lane_in;
          t[u] = lds4(row);
          s[u] = lds4(row + LW7_S);
        }
        flush_up();
        float red[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {  // sweep layers k*U + (U-1-u): upwards
          const int uu = U - 1 - u;
          if (!TAIL || uu < nvalid) { Uu.a = fma2(t[uu].a, Uu.a, s[uu].a); Uu.b = fma2(t[uu].b, Uu.b, s[uu].b); }
          red[u] = hsum4(Uu);
        }
        // the rows are in registers: their L2 lines are dead (no write-back; the next sweep rewrites them in full)
        discard_scratch(srow - (size_t)lane * 16u + (size_t)k * (U * LW7_ROW), (uint32_t)nvalid * LW7_ROW, lane);
#pragma unroll
        for (int u = 0; u < U; ++u) pend[u] = red[u];
        pend_k = k;
      };
      {
        int j = 0;
        if (NG > NGF) backward_group(j++, std::true_type{});  // the ragged group comes first on the way up
        for (; j < NG; ++j) backward_group(j, std::false_type{});
      }
      flush_up();
      pend_k = -1;
      n_bb += (uint32_t)NG;
      __syncwarp();
    }
    // ---- combine the chunks of this column (see lw_solver_v6)
    nx.publish(cluster, csize, ncols_done);
    cluster.sync();
    {
      float* const gout[2] = {p.flux_up + (size_t)col * (L + 1), p.flux_dn + (size_t)col * (L + 1)};
      const int n = 2 * (L + 1), lo = chunk * n / csize, hi = (chunk + 1) * n / csize;
      for (int i = lo + lane; i < hi && owner; i += 32) {
        float sacc = 0.0f;
        for (int r = 0; r < csize; ++r) sacc += *cluster.map_shared_rank(fup + i, r);
        const int a = i / (L + 1);
        gout[a][i - a * (L + 1)] = sacc;
      }
    }
    cb = nx.next(ncols_done);
  }
  cluster.sync();  // nobody leaves while another rank may still read its shared memory
}

// sw_solver_v7 (four g-points per lane, as lw_solver_v7): correct (2.3e-7 of the largest flux from sw_solver_v6) but SLOWER -- 4.20 (6 solver
// warps per SM) / 3.95 ms (7) against 3.78 ms per 30 000 x 137 x 224 on one box -- so it is an opt-in experiment:
// make EXTRA=-DRRNN_EXPERIMENT_SW_WIDE, context flag solver_wide_sw = 1, tools/check_wide_sw.py.
#ifdef RRNN_EXPERIMENT_SW_WIDE
#include "experiments/rte_solvers_sw_wide.cuh"
#endif

// ---------------------------------------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* f = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = reinterpret_cast<EncodeTiledFn>(f);
  return fn;
}
// [rows][ngpt] fp32 tensor, box 64 g-points x box_rows rows, no swizzle (rows of 256 B, read with 8-byte LDS per lane)
static int make_map(CUtensorMap* tm, const float* base, int G, long long rows, int box_rows, int box_cols = 64) {
  EncodeTiledFn enc = encode_fn();
  if (!enc) return fail("rte solvers: cuTensorMapEncodeTiled is not available from the driver");
  const cuuint64_t dims[2] = {(cuuint64_t)G, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)G * 4};
  const cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail("rte solvers: cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")");
  return 0;
}

}  // namespace v5

// Resident clusters are capped so that the reverse-sweep scratch of all of them stays L2-sized.
static int resident_clusters5(const rrnn_ctx_t* ctx, int occ_clusters, int csize, size_t per_cta_bytes, int ncol, int default_mb) {
  const size_t budget = (size_t)(ctx->solver_scratch_mb > 0 ? ctx->solver_scratch_mb : default_mb) << 20;
  long long n = (long long)(budget / (per_cta_bytes * (size_t)csize));
  n = std::max<long long>(n, ctx->num_sms / 2);  // never starve the GPU outright
  n = std::min<long long>(n, occ_clusters);
  n = std::min<long long>(n, ncol);
  return (int)std::max<long long>(n, 1);
}

// One warp is one solver (a 64-g-point chunk of one column); a CTA carries `solver_warps` of them on adjacent columns
// because a cluster launch keeps at most 8 CTAs per SM resident (cudaOccupancyMaxActiveClusters: 284 clusters of 4 on
// 148 SMs whatever the shared memory).  Measured at 137 layers: both sweeps keep gaining from resident warps up to the
// 11.5 per SM that their shared memory allows (2 per CTA, 6 CTAs), now that consumed scratch is discarded from the L2.
template <typename K, typename P, typename... Maps>
static int launch_clustered(rrnn_ctx_t* ctx, K kernel, int csize, size_t warp_smem, size_t warp_scratch, int default_mb, int default_warps, int ncol, P& pp,
                            float** scratch_slot, const Maps&... maps) {
  warp_smem = (warp_smem + 127) & ~(size_t)127;
  pp.warp_smem = (int)warp_smem;
  cudaLaunchConfig_t cfg{};
  cudaLaunchAttribute attr[1];
  cfg.stream = ctx->stream;
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)csize;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int W = std::min(std::max(ctx->solver_warps > 0 ? ctx->solver_warps : default_warps, 1), v5::MAX_WARPS);
  int occ = 0, ncl = 1;
  size_t smem = 0;
  for (;; --W) {
    smem = 128 + (size_t)W * warp_smem + 16;   // (+ the two slots of the dynamic column assignment)
    RRNN_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cfg.gridDim = dim3((unsigned)csize);
    cfg.blockDim = dim3(32u * (unsigned)W);
    cfg.dynamicSmemBytes = smem;
    RRNN_CUDA(cudaOccupancyMaxActiveClusters(&occ, kernel, &cfg));
    ncl = resident_clusters5(ctx, std::max(occ, 1), csize, warp_scratch * W, (ncol + W - 1) / W, default_mb);
    // few columns: spread them over more CTAs instead of stacking them in one
    if (W == 1 || ctx->solver_warps > 0 || (long long)ncl * W <= ncol) break;
  }
  const int ncta = ncl * csize;
  static const bool dbg = getenv("RRNN_SOLVER_DEBUG") != nullptr;
  if (dbg) fprintf(stderr, "[rrnn] solver v5: csize %d warps %d smem %zu occ_clusters %d -> resident %d (%.1f warps/SM, scratch %.1f MB)\n", csize, W, smem,
                   occ, ncl, (double)ncta * W / ctx->num_sms, (double)ncta * W * warp_scratch / 1048576.0);
  cfg.gridDim = dim3((unsigned)ncta);
  if (int rc = ensure_scratch(ctx, (size_t)ncta * W * warp_scratch)) return rc;
  *scratch_slot = (float*)ctx->scratch;
  if (!ctx->col_counter) RRNN_CUDA(cudaMalloc((void**)&ctx->col_counter, 256));
  RRNN_CUDA(cudaMemsetAsync(ctx->col_counter, 0, sizeof(int), ctx->stream));
  pp.next_col = ctx->col_counter;
  RRNN_CUDA(cudaLaunchKernelEx(&cfg, kernel, pp, maps...));
  return 0;
}

// The shapes the v5 kernels take (TMA: row pitch a multiple of 16 B; one cluster per column)
bool lw_v5_supports(int G, int L) { return !(G & 3) && (G + 63) / 64 <= 8 && L >= v5::LW_U; }

// v7: four g-points per lane, 128 per warp (see lw_solver_v7) -- the default where the shape fits; -1 otherwise
static int launch_lw_v7(rrnn_ctx_t* ctx, LwParams& p) {
  const int G = p.ngpt, L = p.nlay;
  const int csize = (G + 127) / 128;
  const bool compact = p.planck_lay != nullptr, cld = p.cld_tau != nullptr, bnd = p.bnd_up != nullptr;
  constexpr int U = 8, S = RRNN_V7_LW_S, SB = RRNN_V7_LW_SB;
  if ((G & 3) || G < 128 || csize > 8 || L < U) return -1;
  if ((compact || cld) && (p.pairs_in_band < 2 || !p.gpt2band)) return -1;   // a lane's four g-points share their band look-ups
  if (compact && !p.planck_lev) return -1;
  if (cld && ctx->fast_math) return -1;                                    // (the cloud / by-band variants are built for the default arithmetic only)
  if (bnd && (compact || cld || ctx->fast_math || !p.bnd_dn || p.nbnd * 16 != G)) return -1;
  for (const void* q : {(const void*)p.tau, (const void*)p.lay_source, (const void*)p.lev_source, (const void*)p.planck_lay, (const void*)p.planck_lev,
                        (const void*)p.cld_tau})
    if ((uintptr_t)q & 15) return -1;
  for (const void* q : {(const void*)p.sfc_emis, (const void*)p.sfc_source, (const void*)p.inc_flux})
    if ((uintptr_t)q & 7) return -1;
  v5::LwV5Params pp;
  pp.b = p;
  pp.ngroups = (L + U - 1) / U;
  const long long rows_lay = (long long)p.ncol * L, rows_lev = (long long)p.ncol * (L + 1);
  if (rows_lev >= (1LL << 31) - 8) return -1;
  const bool top = p.top_at_1 != 0, dn_ext = top || !p.bug_compat, fast = ctx->fast_math != 0;
  CUtensorMap tm_tau, tm_lay, tm_lev, tm_bl, tm_bv, tm_cld;
  if (int rc = v5::make_map(&tm_tau, p.tau, G, rows_lay, U, 128)) return rc;
  tm_cld = tm_tau;
  if (cld) { if (int rc = v5::make_map(&tm_cld, p.cld_tau, 16, rows_lay, U, 16)) return rc; }
  size_t stage;
  if (compact) {
    const int pfr = top ? U + 1 : U;
    if (int rc = v5::make_map(&tm_lay, p.lay_source, G, rows_lay, pfr, 128)) return rc;
    if (int rc = v5::make_map(&tm_bl, p.planck_lay, 16, rows_lay, U, 16)) return rc;
    if (int rc = v5::make_map(&tm_bv, p.planck_lev, 16, rows_lev, U, 16)) return rc;
    tm_lev = tm_tau;
    stage = (size_t)U * 512 + (size_t)pfr * 512 + 2 * U * 64;
  } else {
    if (int rc = v5::make_map(&tm_lay, p.lay_source, G, rows_lay, U, 128)) return rc;
    if (int rc = v5::make_map(&tm_lev, p.lev_source, G, rows_lev, U, 128)) return rc;
    tm_bl = tm_tau; tm_bv = tm_tau;
    stage = (size_t)3 * U * 512;
  }
  if (cld) stage += (size_t)U * 64;
  const size_t smem = (size_t)S * stage + 8 * v5::TR_PITCH * 4 + 4 * (size_t)(L + 1) * 4 + (S + SB) * 8;
  const size_t per_cta = (size_t)L * v5::LW7_ROW;
  // scratch budget: measured optimum at 137 layers (125 ... 145 MB = 6.3 ... 7.3 of the 8 solver warps per SM the shared memory allows:
  // 2.54 - 2.62 ms per 30 000 columns against 2.80 - 2.85 uncapped and 2.75 for lw_solver_v6)
  constexpr int kScratchMbDefault = 140;
  static const int kScratchMb = getenv("RRNN_LW_SCRATCH_MB") ? atoi(getenv("RRNN_LW_SCRATCH_MB")) : kScratchMbDefault;   // (A/B knob)
#define LW7(F, T, D, C, CL, B) launch_clustered(ctx, v5::lw_solver_v7<F, T, D, C, CL, B>, csize, smem, per_cta, kScratchMb, 2, p.ncol, pp, &pp.b.scratch, tm_tau, tm_lay, tm_lev, tm_bl, tm_bv, tm_cld)
#define LW7C(F, T, D, CL) (compact ? LW7(F, T, D, true, CL, false) : LW7(F, T, D, false, CL, false))
  if (bnd) {
    if (top) return LW7(false, true, true, false, false, true);
    return dn_ext ? LW7(false, false, true, false, false, true) : LW7(false, false, false, false, false, true);
  }
  if (cld) {
    if (top) return LW7C(false, true, true, true);
    return dn_ext ? LW7C(false, false, true, true) : LW7C(false, false, false, true);
  }
  if (fast) {
    if (top) return LW7C(true, true, true, false);
    return dn_ext ? LW7C(true, false, true, false) : LW7C(true, false, false, false);
  }
  if (top) return LW7C(false, true, true, false);
  return dn_ext ? LW7C(false, false, true, false) : LW7C(false, false, false, false);
#undef LW7C
#undef LW7
}

// v6: the default (see lw_solver_v6); returns -1 when the shape does not fit
int launch_lw_v6(rrnn_ctx_t* ctx, LwParams& p) {
  if (ctx->solver_wide) {
    const int rc = launch_lw_v7(ctx, p);
    if (rc >= 0) return rc;
  }
  const int G = p.ngpt, L = p.nlay;
  const int csize = (G + 63) / 64;
  const bool compact = p.planck_lay != nullptr;
  constexpr int U = 8, S = 2, SB = 2;
  if (!lw_v5_supports(G, L)) return -1;
  for (const void* q : {(const void*)p.tau, (const void*)p.lay_source, (const void*)p.lev_source, (const void*)p.planck_lay, (const void*)p.planck_lev,
                        (const void*)p.cld_tau})
    if ((uintptr_t)q & 15) return -1;
  const bool cld = p.cld_tau != nullptr;
  if (cld && (!p.gpt2band || ctx->fast_math)) return -1;   // (the cloud variants are built for the default arithmetic only)
  if ((cld || compact) && !p.pairs_in_band) return -1;     // a lane's two g-points share their band look-ups
  for (const void* q : {(const void*)p.sfc_emis, (const void*)p.sfc_source, (const void*)p.inc_flux})
    if ((uintptr_t)q & 7) return -1;
  v5::LwV5Params pp;
  pp.b = p;
  pp.ngroups = (L + U - 1) / U;
  const long long rows_lay = (long long)p.ncol * L, rows_lev = (long long)p.ncol * (L + 1);
  if (rows_lev >= (1LL << 31) - 8) return -1;
  const bool top = p.top_at_1 != 0, dn_ext = top || !p.bug_compat, fast = ctx->fast_math != 0;
  CUtensorMap tm_tau, tm_lay, tm_lev, tm_bl, tm_bv, tm_cld;
  if (int rc = v5::make_map(&tm_tau, p.tau, G, rows_lay, U)) return rc;
  tm_cld = tm_tau;
  if (cld) { if (int rc = v5::make_map(&tm_cld, p.cld_tau, 16, rows_lay, U, 16)) return rc; }
  size_t stage;
  if (compact) {
    if (!p.planck_lev || !p.gpt2band) return fail("lw_solver: incomplete compact source description");
    const int pfr = top ? U + 1 : U;
    if (int rc = v5::make_map(&tm_lay, p.lay_source, G, rows_lay, pfr)) return rc;
    if (int rc = v5::make_map(&tm_bl, p.planck_lay, 16, rows_lay, U, 16)) return rc;
    if (int rc = v5::make_map(&tm_bv, p.planck_lev, 16, rows_lev, U, 16)) return rc;
    tm_lev = tm_tau;
    stage = (size_t)U * 256 + (size_t)pfr * 256 + 2 * U * 64;
  } else {
    if (int rc = v5::make_map(&tm_lay, p.lay_source, G, rows_lay, U)) return rc;
    if (int rc = v5::make_map(&tm_lev, p.lev_source, G, rows_lev, U)) return rc;
    tm_bl = tm_tau; tm_bv = tm_tau;
    stage = (size_t)3 * U * 256;
  }
  if (cld) stage += (size_t)U * 64;
  const size_t smem = (size_t)S * stage + 8 * v5::TR_PITCH * 4 + 4 * (size_t)(L + 1) * 4 + (S + SB) * 8;
  const size_t per_cta = (size_t)L * v5::LW6_ROW;
#define LW6(F, T, D, C, CL) launch_clustered(ctx, v5::lw_solver_v6<F, T, D, C, CL>, csize, smem, per_cta, 400, 2, p.ncol, pp, &pp.b.scratch, tm_tau, tm_lay, tm_lev, tm_bl, tm_bv, tm_cld)
#define LW6C(F, T, D, CL) (compact ? LW6(F, T, D, true, CL) : LW6(F, T, D, false, CL))
  if (p.bnd_up) {   // by-band outputs: materialised sources, no clouds, default arithmetic (what rrnn_rte_lw_byband passes)
    if (compact || cld || fast || !p.bnd_dn || p.nbnd * 16 != G) return -1;
#define LW6B(T, D) launch_clustered(ctx, v5::lw_solver_v6<false, T, D, false, false, true>, csize, smem, per_cta, 400, 2, p.ncol, pp, &pp.b.scratch, tm_tau, tm_lay, tm_lev, tm_bl, tm_bv, tm_cld)
    if (top) return LW6B(true, true);
    return dn_ext ? LW6B(false, true) : LW6B(false, false);
#undef LW6B
  }
  if (cld) {
    if (top) return LW6C(false, true, true, true);
    return dn_ext ? LW6C(false, false, true, true) : LW6C(false, false, false, true);
  }
  if (fast) {
    if (top) return LW6C(true, true, true, false);
    return dn_ext ? LW6C(true, false, true, false) : LW6C(true, false, false, false);
  }
  if (top) return LW6C(false, true, true, false);
  return dn_ext ? LW6C(false, false, true, false) : LW6C(false, false, false, false);
#undef LW6C
#undef LW6
}

#ifdef RRNN_EXPERIMENT_SW_WIDE
// v7: four g-points per lane, 128 per warp (see sw_solver_v7): clear-sky broadband only; -1 where the shape does not fit
static int launch_sw_v7(rrnn_ctx_t* ctx, SwParams& p, bool fast) {
  const int G = p.ngpt, L = p.nlay;
  const int csize = (G + 127) / 128;
  constexpr int U = 8, S = RRNN_V7_SW_S, SB = 2;
  if (p.cld || p.g || p.bnd_up) return -1;
  if ((G & 3) || G < 112 || csize > 8 || L < U) return -1;
  for (const void* q : {(const void*)p.tau, (const void*)p.ssa})
    if ((uintptr_t)q & 15) return -1;
  for (const void* q : {(const void*)p.inc_flux, (const void*)p.inc_flux_dif, (const void*)p.alb_dir, (const void*)p.alb_dif})
    if ((uintptr_t)q & 7) return -1;
  v5::SwV5Params pp;
  pp.b = p;
  pp.ngroups = (L + U - 1) / U;
  const long long rows = (long long)p.ncol * L;
  if (rows >= (1LL << 31) - 8) return -1;
  CUtensorMap tm_tau, tm_ssa;
  if (int rc = v5::make_map(&tm_tau, p.tau, G, rows, U, 128)) return rc;
  if (int rc = v5::make_map(&tm_ssa, p.ssa, G, rows, U, 128)) return rc;
  const size_t stage = (size_t)2 * U * 512;
  const size_t smem = (size_t)S * stage + 16 * v5::TR_PITCH * 4 + 2 * (size_t)(3 * (L + 1) + 1) * 4 + (S + SB) * 8;
  const size_t per_cta = (size_t)L * v5::SW7_ROW;
  const int mb = ctx->solver_scratch_mb_sw_wide > 0 ? ctx->solver_scratch_mb_sw_wide : 400;
  const bool top = p.top_at_1 != 0;
#define SW7(F, T) launch_clustered(ctx, v5::sw_solver_v7<F, T>, csize, smem, per_cta, mb, 2, p.ncol, pp, &pp.b.scratch, tm_tau, tm_ssa)
  if (fast) return top ? SW7(true, true) : SW7(true, false);
  return top ? SW7(false, true) : SW7(false, false);
#undef SW7
}
#endif

// v6: the default (see sw_solver_v6)
int launch_sw_v6(rrnn_ctx_t* ctx, SwParams& p, bool fast) {
#ifdef RRNN_EXPERIMENT_SW_WIDE
  if (ctx->solver_wide_sw) {
    const int rc = launch_sw_v7(ctx, p, fast);
    if (rc >= 0) return rc;
  }
#endif
  const int G = p.ngpt, L = p.nlay;
  const int csize = (G + 63) / 64;
  constexpr int U = 8, SB = 2;
  const int gm = p.cld ? 2 : (p.g ? 1 : 0);
  if (p.cld && (p.g || !p.gpt2band || !p.pairs_in_band)) return -1;   // clouds are folded in on top of gas properties with g == 0 only
  const int S = gm == 0 ? RRNN_V6_SW_S : 2;  // (three input arrays / two + cloud rows: two stages already hold the upward sweep's ring)
  if ((G & 3) || csize > 8 || L < U) return -1;
  for (const void* q : {(const void*)p.tau, (const void*)p.ssa, (const void*)p.g, (const void*)p.cld})
    if ((uintptr_t)q & 15) return -1;
  for (const void* q : {(const void*)p.inc_flux, (const void*)p.inc_flux_dif, (const void*)p.alb_dir, (const void*)p.alb_dif})
    if ((uintptr_t)q & 7) return -1;
  v5::SwV5Params pp;
  pp.b = p;
  pp.ngroups = (L + U - 1) / U;
  const long long rows = (long long)p.ncol * L;
  if (rows >= (1LL << 31) - 8) return -1;
  CUtensorMap tm_tau, tm_ssa, tm_g;
  if (int rc = v5::make_map(&tm_tau, p.tau, G, rows, U)) return rc;
  if (int rc = v5::make_map(&tm_ssa, p.ssa, G, rows, U)) return rc;
  if (gm == 1) { if (int rc = v5::make_map(&tm_g, p.g, G, rows, U)) return rc; }
  else if (gm == 2) { if (int rc = v5::make_map(&tm_g, p.cld, 48, rows, U, 48)) return rc; }
  else tm_g = tm_ssa;
  const size_t stage = (size_t)(gm == 0 ? 2 : 3) * U * 256;
  const size_t smem = (size_t)S * stage + 16 * v5::TR_PITCH * 4 + 2 * (size_t)(3 * (L + 1) + 1) * 4 + (S + SB) * 8;
  const size_t per_cta = (size_t)L * v5::SW6_ROW;
  const bool top = p.top_at_1 != 0;
  static const int kSwScratchMb = getenv("RRNN_SW_SCRATCH_MB") ? atoi(getenv("RRNN_SW_SCRATCH_MB")) : 400;   // (A/B knob; 400 = no cap at 137 layers)
#define SW6(F, GMODE, T) launch_clustered(ctx, v5::sw_solver_v6<F, GMODE, T>, csize, smem, per_cta, kSwScratchMb, 2, p.ncol, pp, &pp.b.scratch, tm_tau, tm_ssa, tm_g)
#define SW6T(F, GMODE) (top ? SW6(F, GMODE, true) : SW6(F, GMODE, false))
  if (p.bnd_up) {   // by-band outputs (rrnn_rte_sw_byband): g == 0 or a g array, no pending clouds
    if (gm == 2 || !p.bnd_dn || !p.bnd_dir || p.nbnd * 16 != G) return -1;
#define SW6B(F, GMODE) (top ? launch_clustered(ctx, v5::sw_solver_v6<F, GMODE, true, true>, csize, smem, per_cta, 400, 2, p.ncol, pp, &pp.b.scratch, tm_tau, tm_ssa, tm_g) \
                            : launch_clustered(ctx, v5::sw_solver_v6<F, GMODE, false, true>, csize, smem, per_cta, 400, 2, p.ncol, pp, &pp.b.scratch, tm_tau, tm_ssa, tm_g))
    if (fast) return gm == 1 ? SW6B(true, 1) : SW6B(true, 0);
    return gm == 1 ? SW6B(false, 1) : SW6B(false, 0);
#undef SW6B
  }
  if (fast) return gm == 2 ? SW6T(true, 2) : (gm == 1 ? SW6T(true, 1) : SW6T(true, 0));
  return gm == 2 ? SW6T(false, 2) : (gm == 1 ? SW6T(false, 1) : SW6T(false, 0));
#undef SW6T
#undef SW6
}
}  // namespace rrnn
