// NN gas optics on the 5th-generation tensor cores (tcgen05 + TMEM + TMA), sm_100a only.
//
// Same fused path as gas_optics_nn.cu (compute_nn_inputs + get_col_dry -> MLP chain -> tau / Planck-source / ssa
// epilogues; every output float written exactly once, neural/mod_network_rrtmgp.F90:125-317,
// rrtmgp/kernels/mo_gas_optics_kernels.F90:615-683) as a warp-specialised, persistent, software-pipelined kernel:
//
//   one CTA per SM, 416 threads, walking tiles of 128 rows (= the 128 TMEM lanes);
//     warps 0-3  FRONT     thread = row: scaled inputs (log p, h2o^1/4, o3^1/4, min-max scaling) -> A operand of
//                          layer 1; hidden epilogues (tcgen05.ld, bias, softsign, fp16 hi/lo split) -> A operand of
//                          the next layer.  Runs about one tile ahead of the output epilogue.
//     warps 4-11 EPILOGUE  two groups of four warps on alternate jobs; thread = row: tcgen05.ld of 32 g-points of both networks at a time from a ring of TMEM
//                          slots; (.)^8 * N_dry | pfrac^2 * Planck(T_lay), Planck(T_lev) | tau_abs+tau_ray, ssa in
//                          registers; rows are staged in 128B-swizzled shared memory and written with TMA tensor stores
//                          (cp.async.bulk.tensor.3d): no load/store instruction is spent on the 3 KB/row of output,
//                          column boundaries inside a tile are handled by the tensor map's bounds (negative start
//                          coordinates and rows beyond nlay are clipped by the TMA unit).
//     warp 12    MMA       one elected thread issues every tcgen05.mma and commits to mbarriers.
//   LW rows run over (level, column) with nlay+1 rows per column: the extra row repeats the bottom layer and produces
//   lev_source(nlay+1) (and the surface source when the surface is at layer nlay), so lev_source needs no special case.
//   precision: every fp32 operand v is split v = hi + lo, hi = fp16(v), lo = fp16(v - hi) (22 mantissa bits) and each
//   GEMM is issued as hi*Whi + lo*Whi + hi*Wlo + lo*Wlo on kind::f16 with fp32 accumulation in TMEM.  The output scaling
//   ystd*(z+b)+ymean is folded into the last layer's weights (times 2^10 so that the fp16 lo parts stay normal; the
//   2^-80 comes back with N_dry) and, where the padded K has a spare column, the bias rides on a column of ones.
//
// Supported: every model generation the reference ships (neural/data): two networks (absorption + Planck fraction,
// absorption + Rayleigh) with 2 hidden layers each, widths up to 80 and different per network (58/16, 64/24, 72/24,
// 80/32, 32/16, 32/32 ...), or ONE longwave network with 2*ngpt outputs (the "both" models,
// rrtmgp/kernels/mo_gas_optics_kernels.F90:745-767: two output heads on one hidden stack); linear output, <= 32
// inputs, ngpt a multiple of 16 and <= 256 (256, 224, 128, 112: a ragged last job of 16 g-points is padded with zero
// weight rows and clipped by the TMA store), nbnd <= 16.  Anything else returns -1 and the caller uses the fp32 FFMA
// kernel.
#include "common.cuh"
#include "f32x2.cuh"
#include <cuda.h>
#include <cuda_fp16.h>
#include <algorithm>
#include <mutex>

namespace rrnn {
namespace tc {

constexpr int TM = 128;          // rows per tile (TMEM lanes)
constexpr int THREADS = 416;     // 4 front warps + 2 x 4 epilogue warps + 1 MMA warp
constexpr int MMA_WARP = 12;
constexpr int KIN_MAX = 32;      // padded number of network inputs
constexpr int NSLOT = 6;         // ring of output accumulators: 6 jobs of 32 g-points x 2 networks = 384 TMEM columns, laid out as
                                 // 3 pairs of 128 columns [net 0: 64 | net 1: 64] so that ONE N = 64 MMA per network and k-step
                                 // feeds two jobs (an SS-mode MMA re-reads its 4 KB A operand from shared memory whatever N is:
                                 // measured ~60 cycles per N = 32 MMA against a 16-cycle tensor floor)
constexpr int RING_COL0 = 128;   // TMEM columns 0..127: hidden accumulators (net 0 at column 0, net 1 at NetP::hid_col)
constexpr int HMAX = 80;         // widest hidden layer (padded to a multiple of 16)
constexpr int REC_F = 6;         // floats per row record
constexpr int STAGE_BYTES = 4096;  // one staged output tile: 32 rows x 32 g-points
constexpr float OUT_SCALE = 1024.0f;            // folded into the last layer of the tau-type networks
constexpr float OUT_UNSCALE = 8.271806125530277e-25f;  // 2^-80 = OUT_SCALE^-8

enum { BAR_AIN = 0, BAR_HID = 1, BAR_ACT = 3, BAR_ACTFREE = 5, BAR_SLOT_FULL = 7, BAR_SLOT_EMPTY = 7 + NSLOT, NBAR = 7 + 2 * NSLOT };

struct NetP {
  int H, K3, Hraw, fold, act0, act1, ntot;
  int hid_col;               // TMEM column of this network's hidden accumulator
  uint32_t w[3][2];          // shared-memory byte offsets of W{1,2,3}{hi,lo} (canonical K-major layout, fp16)
  uint32_t b[3];             // byte offsets of b1[H], b2[H], b3''[N] (fp32)
  uint32_t act_hi, act_lo;   // activation operand (A of layers 2 and 3)
};

struct GasIn {
  const float* ptr;
  int mode;  // 1 = per-layer profile, 2 = (nlay,ncol) field
};

struct Params {
  int ncol, nlay, ngpt, nx, kin, nbnd, ntemp, period, nchunks;
  int nstage;  // staging tiles per epilogue warp (2 when shared memory allows, else 1)
  int nrec;  // depth of the per-row record ring (how many tiles the front warps may run ahead of the output epilogue)
  int nhid;  // hidden stacks: 2 = two networks, 1 = one network with two output heads ("both" models; net[1] shares net[0]'s activations)
  int totplnk_global;  // 1: the Planck table is read from global memory (shared memory is short), 0: from the image in shared memory
  unsigned nrows;
  const float *play, *plev, *tlay, *tlev, *tsfc;
  GasIn gas[KIN_MAX];
  float xmin[KIN_MAX], xmax[KIN_MAX];
  float xconst[KIN_MAX];  // scaled value of inputs that do not vary per sample (scalar gases, missing gases, padding)
  int xvar[KIN_MAX];      // 1 = varies per sample (tlay, play, 1-D / 2-D gas fields)
  float h2o_const;        // h2o vmr when it is a scalar
  NetP net[2];
  const uint8_t* blob;    // weights, biases, band tables exactly as they sit in shared memory
  uint32_t blob_bytes;
  uint32_t off_band4, off_totplnk, off_ain_hi, off_ain_lo, off_stage, off_rec, off_bar;
  const float* totplnk;
  float temp_ref_min, totplnk_delta;
  float *sfc_source, *sfc_jac;
  float *planck_lay, *planck_lev;  // COMPACT only: band Planck functions, rows of 16 floats per layer / level
  unsigned* dbg;   // debug only (RRNN_TC_DEBUG): host-mapped progress words, else null
  int dbg_flags;   // 2 = issue no TMA store, 4 = first box only
};

// ------------------------------------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// shared-memory matrix descriptor: K-major, no swizzle (layout_type 0), version 1 (Blackwell)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
// instruction descriptor, kind::f16: A = B = fp16 (format 0), D = fp32 (c_format 1), both K-major, M x N
__device__ __forceinline__ uint32_t make_idesc(int M, int N) {
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc),
      "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
#ifndef RRNN_MBAR_SUSPEND_NS
#define RRNN_MBAR_SUSPEND_NS 1000   // suspend-time hint of mbarrier.try_wait: how long the hardware may park the thread per attempt
#endif
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity), "r"((uint32_t)RRNN_MBAR_SUSPEND_NS)
      : "memory");
  return ok != 0;
}
#ifndef RRNN_TC_TRACE
#define RRNN_TC_TRACE 0  // 1: compile the progress marks / timeline stamps into the hot loops (debug builds only)
#endif
__device__ unsigned* g_dbg = nullptr;
__device__ __forceinline__ void dbg_mark(unsigned* dbg, int role, unsigned code) {
#if RRNN_TC_TRACE
  if (dbg && blockIdx.x == 0) { volatile unsigned* d = dbg; d[role] = code; }
#endif
}
// debug timeline: SM clock at event `idx` of one steady-state tile of block 0
__device__ __forceinline__ void dbg_ts(unsigned* dbg, int it, int idx) {
#if RRNN_TC_TRACE
  if (dbg && blockIdx.x == 0 && it == 5) { volatile unsigned* d = dbg; d[32 + idx] = (unsigned)clock64(); }
#endif
}
// Spin on an mbarrier phase.  A wait that lasts ~seconds can only be a protocol bug: trap (the launch then fails with
// an error) rather than hang the device.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  unsigned spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 26)) {
      if (g_dbg) {
        volatile unsigned* d = g_dbg;
        d[16] = 0xdead0000u | (unsigned)threadIdx.x; d[17] = bar; d[18] = parity; d[19] = blockIdx.x;
        __threadfence_system();
      }
      __trap();
    }
  }
}
// one lane of a converged warp (elect.sync): lets ptxas issue the uniform-datapath tcgen05 instructions straight,
// without the per-lane retry loop it wraps around them under an ordinary `lane == 0` branch
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tma_store_2d(const CUtensorMap* tm, uint32_t smem, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(reinterpret_cast<uint64_t>(tm)),
               "r"(c0), "r"(c1), "r"(smem)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// packed fp32x2 arithmetic (sm_100): one instruction for two lanes of a register pair
__device__ __forceinline__ void mul2(float& a0, float& a1, float b0, float b1) {
  uint64_t a, b, r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "f"(a0), "f"(a1));
  asm("mov.b64 %0, {%1, %2};" : "=l"(b) : "f"(b0), "f"(b1));
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a0), "=f"(a1) : "l"(r));
}
__device__ __forceinline__ void add2(float& a0, float& a1, float b0, float b1) {  // (a0,a1) += (b0,b1), each lane IEEE like FADD
  uint64_t a, b, r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "f"(a0), "f"(a1));
  asm("mov.b64 %0, {%1, %2};" : "=l"(b) : "f"(b0), "f"(b1));
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a0), "=f"(a1) : "l"(r));
}
__device__ __forceinline__ void mul2to(float& d0, float& d1, float a0, float a1, float b) {  // (d0,d1) = (a0,a1) * b
  uint64_t a, bb, r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "f"(a0), "f"(a1));
  asm("mov.b64 %0, {%1, %2};" : "=l"(bb) : "f"(b), "f"(b));
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(bb));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(d0), "=f"(d1) : "l"(r));
}
__device__ __forceinline__ void prefetch_l2(const void* ptr) { asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr)); }

#define RRNN_R8(a, o) "=r"(a[o]), "=r"(a[o + 1]), "=r"(a[o + 2]), "=r"(a[o + 3]), "=r"(a[o + 4]), "=r"(a[o + 5]), "=r"(a[o + 6]), "=r"(a[o + 7])
// one TMEM lane (= row) x 64 columns per thread
__device__ __forceinline__ void tmem_ld64(uint32_t taddr, float (&v)[64]) {
  uint32_t r[64];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x64.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,"
      "%32,%33,%34,%35,%36,%37,%38,%39,%40,%41,%42,%43,%44,%45,%46,%47,%48,%49,%50,%51,%52,%53,%54,%55,%56,%57,%58,%59,%60,%61,%62,%63}, [%64];\n\t"
      "tcgen05.wait::ld.sync.aligned;"
      : RRNN_R8(r, 0), RRNN_R8(r, 8), RRNN_R8(r, 16), RRNN_R8(r, 24), RRNN_R8(r, 32), RRNN_R8(r, 40), RRNN_R8(r, 48), RRNN_R8(r, 56)
      : "r"(taddr)
      : "memory");
#pragma unroll
  for (int i = 0; i < 64; ++i) v[i] = __uint_as_float(r[i]);
}
// two runs of 32 columns of this thread's lane (net 0 -> v[0..31], net 1 -> v[32..63]), one wait for both
__device__ __forceinline__ void tmem_ld32x2(uint32_t taddr0, uint32_t taddr1, float (&v)[64]) {
  uint32_t r[64];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%64];\n\t"
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%32,%33,%34,%35,%36,%37,%38,%39,%40,%41,%42,%43,%44,%45,%46,%47,%48,%49,%50,%51,%52,%53,%54,%55,%56,%57,%58,%59,%60,%61,%62,%63}, [%65];\n\t"
      "tcgen05.wait::ld.sync.aligned;"
      : RRNN_R8(r, 0), RRNN_R8(r, 8), RRNN_R8(r, 16), RRNN_R8(r, 24), RRNN_R8(r, 32), RRNN_R8(r, 40), RRNN_R8(r, 48), RRNN_R8(r, 56)
      : "r"(taddr0), "r"(taddr1)
      : "memory");
#pragma unroll
  for (int i = 0; i < 64; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n\t"
      "tcgen05.wait::ld.sync.aligned;"
      : RRNN_R8(r, 0), RRNN_R8(r, 8)
      : "r"(taddr)
      : "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
#undef RRNN_R8

// canonical K-major no-swizzle layout of an operand with R rows: byte offset of the 16-byte unit (row r, k-unit ku)
__host__ __device__ __forceinline__ uint32_t unit_off(int R, int r, int ku) { return (uint32_t)ku * (R * 16) + (r >> 3) * 128 + (r & 7) * 16; }

// split 8 fp32 values into fp16 hi / lo (packed conversions) and store the two 16-byte units
__device__ __forceinline__ void store_split8(uint8_t* hi_base, uint8_t* lo_base, uint32_t off, const float* v) {
  __half2 h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    h[i] = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
    const float2 hf = __half22float2(h[i]);
    l[i] = __floats2half2_rn(v[2 * i] - hf.x, v[2 * i + 1] - hf.y);
  }
  *reinterpret_cast<uint4*>(hi_base + off) = *reinterpret_cast<uint4*>(h);
  *reinterpret_cast<uint4*>(lo_base + off) = *reinterpret_cast<uint4*>(l);
}
// softsign x/(|x|+1) with a Newton-refined reciprocal (branch-free, ~1 ulp)
__device__ __forceinline__ float softsign(float x) {
  const float d = fabsf(x) + 1.0f;
  float r = rcp_approx(d);
  r = fmaf(fmaf(-d, r, 1.0f), r, r);
  return x * r;
}
// packed variants for the hidden-layer epilogue: two columns per instruction
__device__ __forceinline__ f2 softsign2(f2 x) {
  float a, b;
  unpack2(x, a, b);
  const float da = fabsf(a) + 1.0f, db = fabsf(b) + 1.0f;
  const f2 d = mk2(da, db);
  f2 r = mk2(rcp_approx(da), rcp_approx(db));
  r = fma2(fnma2(d, r, splat2(1.0f)), r, r);
  return x * r;
}
// split 4 pairs (8 fp32 values) into fp16 hi / lo and store the two 16-byte units
__device__ __forceinline__ void store_split8p(uint8_t* hi_base, uint8_t* lo_base, uint32_t off, const f2* y) {
  __half2 h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float a, b;
    unpack2(y[i], a, b);
    h[i] = __floats2half2_rn(a, b);
    const float2 hf = __half22float2(h[i]);
    const f2 lo = y[i] - mk2(hf.x, hf.y);
    unpack2(lo, a, b);
    l[i] = __floats2half2_rn(a, b);
  }
  *reinterpret_cast<uint4*>(hi_base + off) = *reinterpret_cast<uint4*>(h);
  *reinterpret_cast<uint4*>(lo_base + off) = *reinterpret_cast<uint4*>(l);
}
// generic activations (neural/mod_activation.F90): kept out of line, the shipped models only use softsign
__device__ __noinline__ float act_apply(int code, float x) {
  switch (code) {
    case RRNN_ACT_SOFTSIGN: return softsign(x);
    case RRNN_ACT_RELU: return fmaxf(0.0f, x);
    case RRNN_ACT_SIGMOID: return 1.0f / (1.0f + expf(-x));
    case RRNN_ACT_HARD_SIGMOID: return fmaxf(0.0f, fminf(1.0f, 0.2f * x + 0.5f));
    default: return x;
  }
}

// interpolate1D of compute_Planck_source_nn (mo_gas_optics_kernels.F90:1024-1043): index clamped, fraction not (quirk Q4)
struct PlanckPos {
  int idx;
  float frac;
};
__device__ __forceinline__ PlanckPos planck_pos(float T, float tmin, float delta, int ntemp) {
  const float val0 = (T - tmin) / delta;
  const int iv = (int)val0;
  PlanckPos pp;
  pp.frac = val0 - (float)iv;
  pp.idx = min(ntemp - 1, max(1, iv + 1));
  return pp;
}
__device__ __forceinline__ float planck_at(const PlanckPos pp, const float* __restrict__ tab /* band row */) {
  const float t0 = __ldg(tab + pp.idx - 1);
  return t0 + pp.frac * (__ldg(tab + pp.idx) - t0);
}

// Issue D[128 x N] = A[128 x K] * W[N x K]^T as the four split products hi*hi + lo*hi + hi*lo + lo*lo.
// Operands are given as ready-made shared-memory descriptors of their first k-step (the address field is the low 14
// bits, in 16-byte units, so advancing an operand is one integer add): A advances 2 k-units x 2048 B = 256 units per
// k-step, W advances `wstep` units.  The issuing thread is a serial resource -- ~200 MMAs per tile -- so nothing but
// the adds and the MMA itself is left in these loops.
// Order matters: the tensor core truncates (does not round) when it adds a product block to the accumulator, one
// accumulator-ulp per tcgen05.mma, always in the same direction.  The three correction products are 2^-11 of the main
// one, so they are accumulated FIRST, while the accumulator is still small (their truncations are then negligible),
// and the hi*hi blocks last: K/16 full-size truncations instead of 4K/16 (measured: the systematic relative tau offset
// of the interleaved order, -4e-4 at the most amplified g-point, drops accordingly).
struct OperandDesc {
  uint64_t hi, lo;
};
__device__ __forceinline__ void issue_gemm(uint32_t tmem_d, const OperandDesc a, const OperandDesc w, uint32_t wstep, int ksteps,
                                           uint32_t idesc) {
  mma_f16(tmem_d, a.lo, w.lo, idesc, 0);
  mma_f16(tmem_d, a.lo, w.hi, idesc, 1);
  mma_f16(tmem_d, a.hi, w.lo, idesc, 1);
  for (int s = 1; s < ksteps; ++s) {
    const uint64_t ao = (uint64_t)(256u * s), wo = (uint64_t)(wstep * s);
    mma_f16(tmem_d, a.lo + ao, w.lo + wo, idesc, 1);
    mma_f16(tmem_d, a.lo + ao, w.hi + wo, idesc, 1);
    mma_f16(tmem_d, a.hi + ao, w.lo + wo, idesc, 1);
  }
  for (int s = 0; s < ksteps; ++s) mma_f16(tmem_d, a.hi + (uint64_t)(256u * s), w.hi + (uint64_t)(wstep * s), idesc, 1);
}

// get_col_dry for one layer (rrtmgp/mo_gas_optics_rrtmgp.F90:1697-1703)
__device__ __forceinline__ float col_dry_of(float h2o, float p0, float p1) {
  const float dp = fabsf(p0 - p1);
  const float fact = 1.0f / (1.0f + h2o);
  const float m_air = (0.028964f + 0.018016f * h2o) * fact;
  return 10.0f * dp * 6.02214076e23f * fact / (1000.0f * m_air * 100.0f * 9.80665f);
}

// MODE 0 = LW (net 0: absorption -> tau; net 1: Planck fraction -> lay_source, lev_source, sfc_source[_Jac])
// MODE 1 = SW (net 0: absorption, net 1: Rayleigh -> tau = abs + ray, ssa = ray / tau)
template <int MODE, bool COMPACT>
__global__ void __launch_bounds__(THREADS, 1)
gas_optics_tc_kernel(const __grid_constant__ Params p, const __grid_constant__ CUtensorMap tm0,
                     const __grid_constant__ CUtensorMap tm1, const __grid_constant__ CUtensorMap tm2,
                     const __grid_constant__ CUtensorMap tm0s, const __grid_constant__ CUtensorMap tm1s) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // the dynamic shared memory window is at least 16-byte aligned; the swizzled staging tiles want 1024
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int L = p.nlay, G = p.ngpt, P = p.period;

  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.off_bar);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + NBAR);
  const uint32_t bar0 = smem_u32(bars);
  auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };

  // ---------------------------------------------------------------------------------- one-time set-up
  {
    const uint4* src = reinterpret_cast<const uint4*>(p.blob);
    uint4* dst = reinterpret_cast<uint4*>(smem);
    for (uint32_t i = tid; i < p.blob_bytes / 16; i += THREADS) dst[i] = __ldg(src + i);
    // activation operands start from zero (padding k-units are never written again); a folded bias that lives in an
    // extension k-unit (Hraw == H) gets its column of ones here, once
    for (int n = 0; n < p.nhid; ++n) {
      const NetP& nt = p.net[n];
      uint4* a = reinterpret_cast<uint4*>(smem + nt.act_hi);
      const uint32_t units = (uint32_t)(2 * TM * nt.K3 * 2) / 16;  // hi and lo are contiguous
      for (uint32_t i = tid; i < units; i += THREADS) a[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    uint4* ain = reinterpret_cast<uint4*>(smem + p.off_ain_hi);
    for (uint32_t i = tid; i < (uint32_t)(2 * TM * p.kin * 2) / 16; i += THREADS) ain[i] = make_uint4(0u, 0u, 0u, 0u);
  }
  __syncthreads();
  for (int n = 0; n < p.nhid; ++n) {
    const NetP& nt = p.net[n];
    if (nt.fold && nt.Hraw >= nt.H && tid < TM) {
      __half* a = reinterpret_cast<__half*>(smem + nt.act_hi + unit_off(TM, tid, nt.Hraw >> 3));
      a[nt.Hraw & 7] = __float2half_rn(1.0f);
    }
  }
  if (tid == 0) {
    mbar_init(BAR(BAR_AIN), 128);
    for (int n = 0; n < 2; ++n) {
      mbar_init(BAR(BAR_HID + n), 1);
      mbar_init(BAR(BAR_ACT + n), 128);
      mbar_init(BAR(BAR_ACTFREE + n), 1);
    }
    for (int s = 0; s < NSLOT; ++s) {
      mbar_init(BAR(BAR_SLOT_FULL + s), 1);
      mbar_init(BAR(BAR_SLOT_EMPTY + s), 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_async_smem();
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const unsigned ntiles = (p.nrows + TM - 1) / TM;
  if (p.dbg && tid == 0) { g_dbg = p.dbg; dbg_mark(p.dbg, 15, 0x1000u + (bar0 & 0xffffu)); }

  if (warp < 4) {
    // =================================================================================== FRONT: thread = row
    const int r = tid;
    const uint32_t tmem_row = tmem_base + ((uint32_t)(32 * warp) << 16);
    uint8_t* ain_hi = smem + p.off_ain_hi;
    uint8_t* ain_lo = smem + p.off_ain_lo;
    uint32_t ph_hid[2] = {0u, 0u}, ph_free[2] = {0u, 0u};
    int sfc_lev = -1;
    if (MODE == 0) sfc_lev = (__ldg(p.play) > __ldg(p.play + L - 1)) ? 0 : L;  // merge(1,nlay,play(1,1) > play(nlay,1)); layer nlay <-> extra row
    // Raw per-row values of the NEXT tile are loaded while this tile's hidden layers run, so the prologue never waits on
    // global memory: T_lay, p_lay, the first two gases (h2o, o3), the two bounding p_lev, T_lev, T_sfc.
    struct RowIn {
      unsigned col;
      int lev, lay;
      size_t smp;
      bool valid;
      float x0, x1, x2, x3, p0, p1, tv, ts;
    };
    auto load_row = [&](unsigned tile) {
      RowIn q;
      const unsigned s = tile * TM + r;
      q.valid = tile < ntiles && s < p.nrows;
      q.col = q.valid ? s / (unsigned)P : 0u;
      q.lev = q.valid ? (int)(s - q.col * (unsigned)P) : 0;
      q.lay = min(q.lev, L - 1);
      q.smp = (size_t)q.col * L + q.lay;
      q.x0 = q.x1 = 200.0f; q.x2 = q.x3 = 0.0f; q.p0 = q.p1 = 0.0f; q.tv = q.ts = 200.0f;
      if (q.valid) {
        q.x0 = __ldg(p.tlay + q.smp);
        q.x1 = __ldg(p.play + q.smp);
        if (p.xvar[2]) q.x2 = (p.gas[2].mode == 2) ? __ldg(p.gas[2].ptr + q.smp) : __ldg(p.gas[2].ptr + q.lay);
        if (p.xvar[3]) q.x3 = (p.gas[3].mode == 2) ? __ldg(p.gas[3].ptr + q.smp) : __ldg(p.gas[3].ptr + q.lay);
        const float* pl = p.plev + (size_t)q.col * (L + 1) + q.lay;
        q.p0 = __ldg(pl); q.p1 = __ldg(pl + 1);
        if (MODE == 0) {
          q.tv = __ldg(p.tlev + (size_t)q.col * (L + 1) + q.lev);
          if (q.lev == sfc_lev) q.ts = __ldg(p.tsfc + q.col);
        }
      }
      return q;
    };
    RowIn nxt = load_row(blockIdx.x);
    int it = 0;
    for (unsigned tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
      // ---- compute_nn_inputs (mo_gas_optics_rrtmgp.F90:713-782) for this row
      if (tid == 0) dbg_ts(p.dbg, it, 0);
      const RowIn cur = nxt;
      const bool valid = cur.valid;
      const unsigned col = cur.col;
      const int lev = cur.lev, lay = cur.lay;
      const size_t smp = cur.smp;
#pragma unroll 1
      for (int ku = 0; ku < (p.kin >> 3); ++ku) {
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int k = 8 * ku + j;
          float x = p.xconst[k];
          if (p.xvar[k]) {
            float raw;
            if (k == 0) raw = cur.x0;
            else if (k == 1) raw = logf(cur.x1);
            else if (k == 2) raw = sqrtf(sqrtf(cur.x2));
            else if (k == 3) raw = sqrtf(sqrtf(cur.x3));
            else raw = (p.gas[k].mode == 2) ? __ldg(p.gas[k].ptr + smp) : __ldg(p.gas[k].ptr + lay);
            x = (raw - p.xmin[k]) / (p.xmax[k] - p.xmin[k]);
          }
          v[j] = valid ? x : 0.0f;
        }
        store_split8(ain_hi, ain_lo, unit_off(TM, r, ku), v);
      }
      // ---- per-row record for the output epilogue of this tile (a ring of nrec tiles; ordered by the
      //      AIN -> ... -> SLOT_FULL barrier chain): N_dry * 2^-80 and where T_lay, T_lev, T_sfc fall in the Planck table
      {
        float* rec = reinterpret_cast<float*>(smem + p.off_rec) + (it % p.nrec) * (REC_F * TM) + r;
        float cdp = 0.0f;
        if (valid) {
          const float h = p.xvar[2] ? cur.x2 : p.h2o_const;
          cdp = col_dry_of(h, cur.p0, cur.p1) * OUT_UNSCALE;
        }
        rec[0] = cdp;
        if (MODE == 0) {
          const PlanckPos pl_ = planck_pos(cur.x0, p.temp_ref_min, p.totplnk_delta, p.ntemp);
          const PlanckPos pv_ = planck_pos(cur.tv, p.temp_ref_min, p.totplnk_delta, p.ntemp);
          rec[1 * TM] = pl_.frac; rec[2 * TM] = __int_as_float(pl_.idx);
          rec[3 * TM] = pv_.frac; rec[4 * TM] = __int_as_float(pv_.idx);
          if (valid && lev == sfc_lev) rec[5 * TM] = cur.ts;
        }
      }
      fence_async_smem();
      mbar_arrive(BAR(BAR_AIN));
      if (tid == 0) { dbg_mark(p.dbg, 0, (it << 8) | 1); dbg_ts(p.dbg, it, 1); }
      // ---- the next tile's raw inputs start their way to registers now
      nxt = load_row(tile + gridDim.x);
      (void)col;
      // ---- hidden epilogues: layer 1 of net 0, net 1; layer 2 of net 0, net 1
#pragma unroll 1
      for (int l = 0; l < 2; ++l) {
#pragma unroll 1
        for (int n = 0; n < p.nhid; ++n) {
          const NetP& nt = p.net[n];
          uint8_t* act_hi = smem + nt.act_hi;
          uint8_t* act_lo = smem + nt.act_lo;
          // the previous tile's output-layer MMAs read this buffer: wait until they are done
          if (l == 0 && it > 0) { mbar_wait(BAR(BAR_ACTFREE + n), ph_free[n]); ph_free[n] ^= 1u; }
          if (tid == 0) dbg_ts(p.dbg, it, 2 + 3 * (2 * l + n));
          mbar_wait(BAR(BAR_HID + n), ph_hid[n]); ph_hid[n] ^= 1u;
          fence_after();
          if (tid == 0) dbg_ts(p.dbg, it, 3 + 3 * (2 * l + n));
          const float* bb = reinterpret_cast<const float*>(smem + nt.b[l]);
          const int act = l ? nt.act1 : nt.act0;
          const bool ones = (l == 1) && nt.fold && nt.Hraw < nt.H;
          for (int c0 = 0; c0 < nt.H; c0 += 16) {
            float v[16];
            tmem_ld16(tmem_row + (uint32_t)(nt.hid_col + c0), v);
            const bool has_one = ones && (nt.Hraw >> 4) == (c0 >> 4);
            if (act == RRNN_ACT_SOFTSIGN && !has_one) {
              // the common case, in packed arithmetic: bias, softsign, fp16 hi/lo split of two columns at a time
              f2 y[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float2 b = *reinterpret_cast<const float2*>(bb + c0 + 2 * j);
                y[j] = softsign2(mk2(v[2 * j], v[2 * j + 1]) + mk2(b.x, b.y));
              }
              store_split8p(act_hi, act_lo, unit_off(TM, r, c0 >> 3), y);
              store_split8p(act_hi, act_lo, unit_off(TM, r, (c0 >> 3) + 1), y + 4);
            } else {
              if (act == RRNN_ACT_SOFTSIGN) {
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = softsign(v[j] + bb[c0 + j]);
              } else {
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = act_apply(act, v[j] + bb[c0 + j]);
              }
              if (has_one) {
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = ((nt.Hraw & 15) == j) ? 1.0f : v[j];
              }
              store_split8(act_hi, act_lo, unit_off(TM, r, c0 >> 3), v);
              store_split8(act_hi, act_lo, unit_off(TM, r, (c0 >> 3) + 1), v + 8);
            }
          }
          fence_before();
          fence_async_smem();
          mbar_arrive(BAR(BAR_ACT + n));
          if (tid == 0) { dbg_mark(p.dbg, 0, (it << 8) | (0x10 + 2 * l + n)); dbg_ts(p.dbg, it, 4 + 3 * (2 * l + n)); }
        }
      }
    }
  } else if (warp == MMA_WARP) {
    // =================================================================================== MMA issuer
    // The whole warp walks the loop and waits on the barriers (it stays converged); one elected lane issues.
    {
      uint32_t ph_ain = 0u, ph_act[2] = {0u, 0u};
      unsigned jc = 0;
      const uint32_t sb = smem_u32(smem);
      // every operand descriptor is tile-invariant: build them once
      const OperandDesc d_ain = {make_desc(sb + p.off_ain_hi, TM * 16, 128), make_desc(sb + p.off_ain_lo, TM * 16, 128)};
      OperandDesc d_act[2], d_w[2][3];
      uint32_t wstep12[2], idesc_h[2];
#pragma unroll
      for (int n = 0; n < 2; ++n) {
        const NetP& nt = p.net[n];
        d_act[n] = {make_desc(sb + nt.act_hi, TM * 16, 128), make_desc(sb + nt.act_lo, TM * 16, 128)};
#pragma unroll
        for (int l = 0; l < 3; ++l) {
          const uint32_t lbo = (uint32_t)(l == 2 ? nt.ntot : nt.H) * 16u;
          d_w[n][l] = {make_desc(sb + nt.w[l][0], lbo, 128), make_desc(sb + nt.w[l][1], lbo, 128)};
        }
        wstep12[n] = 2u * (uint32_t)nt.H;  // two k-units of H rows x 16 B, in 16-byte units
        idesc_h[n] = make_idesc(TM, nt.H);
      }
      const uint32_t wstep3 = 2u * (uint32_t)p.net[0].ntot;  // both heads are packed with the same (padded) number of rows
      const uint32_t idesc_o = make_idesc(TM, 32), idesc_o2 = make_idesc(TM, 64);
      const int nch_pad = (p.nchunks + 1) & ~1;
      int mit = 0;
      for (unsigned tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++mit) {
        if (lane == 0) dbg_ts(p.dbg, mit, 20);
        mbar_wait(BAR(BAR_AIN), ph_ain); ph_ain ^= 1u;
        fence_after();
        if (lane == 0) dbg_ts(p.dbg, mit, 21);
        if (elect_one()) {
#pragma unroll
          for (int n = 0; n < 2; ++n) {
            if (n < p.nhid) {
              issue_gemm(tmem_base + (uint32_t)p.net[n].hid_col, d_ain, d_w[n][0], wstep12[n], p.kin >> 4, idesc_h[n]);
              mma_commit(BAR(BAR_HID + n));
            }
          }
        }
        __syncwarp();
        if (lane == 0) dbg_ts(p.dbg, mit, 22);
#pragma unroll
        for (int n = 0; n < 2; ++n) {
          if (n >= p.nhid) continue;  // warp-uniform
          mbar_wait(BAR(BAR_ACT + n), ph_act[n]); ph_act[n] ^= 1u;
          fence_after();
          if (lane == 0) dbg_ts(p.dbg, mit, 23 + 2 * n);
          if (elect_one()) {
            issue_gemm(tmem_base + (uint32_t)p.net[n].hid_col, d_act[n], d_w[n][1], wstep12[n], p.net[n].H >> 4, idesc_h[n]);
            mma_commit(BAR(BAR_HID + n));
          }
          __syncwarp();
          if (lane == 0) dbg_ts(p.dbg, mit, 24 + 2 * n);
        }
#pragma unroll
        for (int n = 0; n < 2; ++n) {
          if (n < p.nhid) { mbar_wait(BAR(BAR_ACT + n), ph_act[n]); ph_act[n] ^= 1u; }
        }
        fence_after();
        if (lane == 0) dbg_ts(p.dbg, mit, 27);
        // two jobs (64 g-points) per set of MMAs; an odd last job (ngpt = 224) is padded with an empty one so that job
        // pairs and slot pairs stay aligned across tiles
        for (int c = 0; c < nch_pad; c += 2, jc += 2) {
          const unsigned slot = jc % NSLOT, round = jc / NSLOT;  // slot is even: the pair (slot, slot + 1)
          if (round > 0) {
            mbar_wait(BAR(BAR_SLOT_EMPTY + slot), (round - 1u) & 1u);
            mbar_wait(BAR(BAR_SLOT_EMPTY + slot + 1), (round - 1u) & 1u);
            fence_after();
          }
          if (elect_one()) {
            const bool both = c + 1 < p.nchunks;
#pragma unroll
            for (int n = 0; n < 2; ++n) {
              // rows [32c, 32c+64) of W3: 8-row groups are 128 B apart -> 512 B = 32 descriptor units per 32 rows
              const OperandDesc w3 = {d_w[n][2].hi + (uint64_t)(32u * c), d_w[n][2].lo + (uint64_t)(32u * c)};
              issue_gemm(tmem_base + RING_COL0 + 128u * (slot >> 1) + 64u * n, d_act[n], w3, wstep3, p.net[n].K3 >> 4,
                         both ? idesc_o2 : idesc_o);
            }
            mma_commit(BAR(BAR_SLOT_FULL + slot));
            mma_commit(BAR(BAR_SLOT_FULL + slot + 1));
          }
          __syncwarp();
          if (lane == 0) dbg_ts(p.dbg, mit, 28 + c);
        }
        if (elect_one()) {
          mma_commit(BAR(BAR_ACTFREE + 0));
          if (p.nhid > 1) mma_commit(BAR(BAR_ACTFREE + 1));
        }
        __syncwarp();
      }
    }
  } else {
    // =================================================================================== EPILOGUE: thread = row
    // Two groups of four warps (warps 4-7 and 8-11): a warp may only touch the TMEM lanes of quarter warp % 4, so the
    // groups share the rows and take alternate jobs -- two instruction streams per scheduler to hide each other's stalls.
    const int q = warp & 3;            // TMEM lane quarter
    const int eg = (warp - 4) >> 2;    // epilogue group: handles the jobs with jc % 2 == eg
    const int r = 32 * q + lane;
    const uint32_t tmem_row = tmem_base + ((uint32_t)(32 * q) << 16);
    uint8_t* stage = smem + p.off_stage + (uint32_t)(warp - 4) * (uint32_t)(p.nstage * STAGE_BYTES);
    const uint32_t stage_a = smem_u32(stage);
    const int* band4_s = reinterpret_cast<const int*>(smem + p.off_band4);  // band of each group of 4 g-points
    // totplnk [band][ntemp]: in the shared-memory image, or (wide networks, shared memory short) in global memory / L1
    const float* tp_s = p.totplnk_global ? p.totplnk : reinterpret_cast<const float*>(smem + p.off_totplnk);
    const float* b3_0 = reinterpret_cast<const float*>(smem + p.net[0].b[2]);
    const float* b3_1 = reinterpret_cast<const float*>(smem + p.net[1].b[2]);
    const bool fold0 = p.net[0].fold != 0, fold1 = p.net[1].fold != 0;
    int sfc_lev = -1;
    if (MODE == 0) sfc_lev = (__ldg(p.play) > __ldg(p.play + L - 1)) ? 0 : L;
    int it = 0;
    unsigned jc = 0;
    const int nch_pad = (p.nchunks + 1) & ~1;  // jobs per tile incl. the padding job of an odd count (see the MMA warp)
    int sbuf = 0;
    const unsigned nrows_lay = (unsigned)p.ncol * (unsigned)L;  // rows of tau / lay_source / ssa

    // Stage one tile of up to 32 rows x 32 g-points and write it with one TMA store.  Every array is a 2-D tensor
    // [rows][ngpt] whose rows are contiguous across columns, so a tile is one box starting at row `dest`; `slot_row` is
    // this thread's row inside the staged tile (-1: this row is not part of the array) -- see the LW row numbering.
    auto stage_and_store = [&](const CUtensorMap* tm, const float (&o)[32], int g0, int slot_row, unsigned dest, unsigned nrows_arr) {
      if (p.nstage == 1) {
        // Shared memory holds ONE 4 KB staging tile per warp (the g256 LW networks): it is used as two half tiles of
        // 32 rows x 16 g-points (64-byte rows, 64-byte swizzle; the tensor maps are encoded to match), so that a half is
        // refilled while the TMA store of the other is still reading -- no blocking wait, twice the stores.
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          if (elect_one()) bulk_wait_read<1>();  // the store issued two halves ago has read this half
          __syncwarp();
          if (slot_row >= 0) {
            uint8_t* row = stage + sbuf * (STAGE_BYTES / 2) + slot_row * 64;
            const uint32_t swz = (uint32_t)((slot_row >> 1) & 3);
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4)
              *reinterpret_cast<float4*>(row + (((uint32_t)j4 ^ swz) << 4)) =
                  make_float4(o[16 * hf + 4 * j4], o[16 * hf + 4 * j4 + 1], o[16 * hf + 4 * j4 + 2], o[16 * hf + 4 * j4 + 3]);
          }
          fence_async_smem();
          __syncwarp();
          if (elect_one()) {   // (bulk async-groups are per thread: elect.sync picks the same lane every time, waits included)
            // (a half tile that starts past the last g-point -- ngpt = 112 -- is not stored; one that starts inside is clipped)
            if (dest < nrows_arr && g0 + 16 * hf < G && !(p.dbg_flags & 2)) tma_store_2d(tm, stage_a + sbuf * (STAGE_BYTES / 2), g0 + 16 * hf, (int)dest);
            bulk_commit();
          }
          sbuf ^= 1;
        }
        return;
      }
      if (elect_one()) bulk_wait_read<1>();  // the staging tile about to be overwritten has been read by its TMA store
      __syncwarp();
      if (slot_row >= 0) {
        uint8_t* row = stage + sbuf * STAGE_BYTES + slot_row * 128;
        const uint32_t swz = (uint32_t)(slot_row & 7);
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4)
          *reinterpret_cast<float4*>(row + (((uint32_t)j4 ^ swz) << 4)) = make_float4(o[4 * j4], o[4 * j4 + 1], o[4 * j4 + 2], o[4 * j4 + 3]);
      }
      fence_async_smem();
      __syncwarp();
      if (elect_one()) {
        // a box that starts inside the tensor may hang over its end: those rows are clipped by the TMA unit
        if (dest < nrows_arr && !(p.dbg_flags & 2)) tma_store_2d(tm, stage_a + sbuf * STAGE_BYTES, g0, (int)dest);
        bulk_commit();
      }
      sbuf ^= 1;
    };

    for (unsigned tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
      const unsigned s0 = tile * TM;
      const unsigned s = s0 + r;
      const bool valid = s < p.nrows;
      const unsigned col = valid ? s / (unsigned)P : 0u;
      const int lev = valid ? (int)(s - col * (unsigned)P) : 0;
      // Destination rows of this warp's 32 rows.  SW: row s of the tile is row s of tau / ssa.  LW: rows are numbered
      // over nlay+1 levels per column, which IS the row number of lev_source; tau and lay_source have nlay rows per
      // column, so the extra row (lev == nlay; at most one per warp since nlay >= 31) is squeezed out of the staged tile:
      // rows after it move up by one and the box is 31 rows high (tensor maps tm0s / tm1s).
      const unsigned w0 = s0 + 32u * q;
      const unsigned wcol = w0 / (unsigned)P;
      const int wlev = (int)(w0 - wcol * (unsigned)P);
      int slot_row = lane;
      bool squeezed = false;
      unsigned dest_lay = w0;
      if (MODE == 0) {
        const int ph = L - wlev;  // position of the extra row in this warp's 32 rows, if 0 <= ph < 32
        squeezed = ph < 32 && w0 + (unsigned)ph < p.nrows;
        if (squeezed) slot_row = (lane == ph) ? -1 : (lane > ph ? lane - 1 : lane);
        dest_lay = w0 - wcol;
      }
      const bool is_sfc = (MODE == 0) && valid && lev == sfc_lev;
      const float* rec = reinterpret_cast<const float*>(smem + p.off_rec) + (it % p.nrec) * (REC_F * TM) + r;
      float cdp = 0.0f, frac_l = 0.0f, frac_v = 0.0f;
      const float *tp_l = tp_s, *tp_v = tp_s;   // &totplnk[0][idx-1] for T_lay and T_lev of this row
      PlanckPos ps0{1, 0.0f}, ps1{1, 0.0f};      // T_sfc, T_sfc + 1 (surface row only)
      bool have_rec = false;
      for (int c = 0; c < nch_pad; ++c, ++jc) {
        if ((int)(jc & 1u) != eg) continue;
        const unsigned slot = jc % NSLOT, round = jc / NSLOT;
        const int g0 = 32 * c;
        if (tid == 128) dbg_ts(p.dbg, it, 40 + 2 * c);
        mbar_wait(BAR(BAR_SLOT_FULL + slot), round & 1u);
        fence_after();
        if (c >= p.nchunks) {  // the padding job of an odd job count: keep the slot protocol going, nothing to do
          __syncwarp();
          if (lane == 0) mbar_arrive(BAR(BAR_SLOT_EMPTY + slot));
          continue;
        }
        if (tid == 128) dbg_ts(p.dbg, it, 41 + 2 * c);
        if (!have_rec) {
          // the record of this tile was written by the front warps before the MMAs this barrier tracks were issued
          have_rec = true;
          cdp = rec[0];
          if (MODE == 0) {
            frac_l = rec[1 * TM]; tp_l = tp_s + __float_as_int(rec[2 * TM]) - 1;
            frac_v = rec[3 * TM]; tp_v = tp_s + __float_as_int(rec[4 * TM]) - 1;
            if (is_sfc) {
              const float Ts = rec[5 * TM];
              ps0 = planck_pos(Ts, p.temp_ref_min, p.totplnk_delta, p.ntemp);
              ps1 = planck_pos(Ts + 1.0f, p.temp_ref_min, p.totplnk_delta, p.ntemp);
            }
          }
        }
        if (tid == 128) dbg_mark(p.dbg, 2, (jc << 8) | 1);
        float z[64];
        {
          const uint32_t pair_col = tmem_row + RING_COL0 + 128u * (slot >> 1) + 32u * (slot & 1u);
          tmem_ld32x2(pair_col, pair_col + 64u, z);
        }
        fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(BAR(BAR_SLOT_EMPTY + slot));
        if (!fold0) {
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4) {
            const float4 b = *reinterpret_cast<const float4*>(b3_0 + g0 + 4 * j4);
            add2(z[4 * j4], z[4 * j4 + 1], b.x, b.y); add2(z[4 * j4 + 2], z[4 * j4 + 3], b.z, b.w);
          }
        }
        if (!fold1) {
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4) {
            const float4 b = *reinterpret_cast<const float4*>(b3_1 + g0 + 4 * j4);
            add2(z[32 + 4 * j4], z[32 + 4 * j4 + 1], b.x, b.y); add2(z[32 + 4 * j4 + 2], z[32 + 4 * j4 + 3], b.z, b.w);
          }
        }
        float o[32];
        if (MODE == 0) {
          // ---- tau = ((ystd*(z+b)+ymean)**8)*col_dry (mod_network_rrtmgp.F90:209-222); z here is 2^10 x that bracket
#pragma unroll
          for (int j = 0; j < 32; j += 2) {
            float t0 = z[j], t1 = z[j + 1];
            mul2(t0, t1, t0, t1); mul2(t0, t1, t0, t1); mul2(t0, t1, t0, t1);
            mul2(t0, t1, cdp, cdp);
            o[j] = t0; o[j + 1] = t1;
          }
          stage_and_store(squeezed ? &tm0s : &tm0, o, g0, slot_row, dest_lay, nrows_lay);
          // ---- Planck fraction (:309-312) -> lay_source, lev_source (compute_Planck_source_nn, interpolate1D); every
          //      group of 4 g-points lies in one band (checked on the host), bands change rarely along g
#pragma unroll
          for (int j = 0; j < 32; j += 2) mul2(z[32 + j], z[33 + j], z[32 + j], z[33 + j]);
          int b4[8];
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4) b4[j4] = band4_s[(g0 >> 2) + j4];
          if (COMPACT) {
            // The sources stay factored: the Planck fraction goes out as it is, and the band Planck functions of this
            // row (16 floats for T_lay, 16 for T_lev; written once per row, by the job of the first g-points) are
            // multiplied in by lw_solver_v5<COMPACT> -- the same single fp32 product, 8 instead of 12 bytes per g-point.
#pragma unroll
            for (int j = 0; j < 32; ++j) o[j] = z[32 + j];
            stage_and_store(squeezed ? &tm1s : &tm1, o, g0, slot_row, dest_lay, nrows_lay);
            if (c == 0 && valid) {
              float4* bv_out = reinterpret_cast<float4*>(p.planck_lev + (size_t)s * 16);
              float4* bl_out = reinterpret_cast<float4*>(p.planck_lay + ((size_t)col * L + lev) * 16);
#pragma unroll 1
              for (int q4 = 0; q4 < 4; ++q4) {
                float vl[4], vv[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const int b = min(4 * q4 + e, p.nbnd - 1);
                  const float* tl = tp_l + b * p.ntemp;
                  const float* tv = tp_v + b * p.ntemp;
                  const float l0 = tl[0], v0 = tv[0];
                  vl[e] = l0 + frac_l * (tl[1] - l0);
                  vv[e] = v0 + frac_v * (tv[1] - v0);
                }
                bv_out[q4] = make_float4(vv[0], vv[1], vv[2], vv[3]);
                if (lev < L) bl_out[q4] = make_float4(vl[0], vl[1], vl[2], vl[3]);
              }
            }
          } else {
          float bl = 0.0f, bv = 0.0f;
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4) {
            if (j4 == 0 || b4[j4] != b4[j4 - 1]) {  // warp-uniform
              const float* t = tp_l + b4[j4] * p.ntemp;
              const float t0 = t[0];
              bl = t0 + frac_l * (t[1] - t0);
            }
            mul2to(o[4 * j4], o[4 * j4 + 1], z[32 + 4 * j4], z[33 + 4 * j4], bl);
            mul2to(o[4 * j4 + 2], o[4 * j4 + 3], z[34 + 4 * j4], z[35 + 4 * j4], bl);
          }
          stage_and_store(squeezed ? &tm1s : &tm1, o, g0, slot_row, dest_lay, nrows_lay);
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4) {
            if (j4 == 0 || b4[j4] != b4[j4 - 1]) {
              const float* t = tp_v + b4[j4] * p.ntemp;
              const float t0 = t[0];
              bv = t0 + frac_v * (t[1] - t0);
            }
            mul2to(o[4 * j4], o[4 * j4 + 1], z[32 + 4 * j4], z[33 + 4 * j4], bv);
            mul2to(o[4 * j4 + 2], o[4 * j4 + 3], z[34 + 4 * j4], z[35 + 4 * j4], bv);
          }
          stage_and_store(&tm2, o, g0, lane, w0, p.nrows);
          }
          if (is_sfc) {  // surface source and its Jacobian: one row per column, written by the owning lane
            float* ss = p.sfc_source + (size_t)col * G + g0;
            float* sj = p.sfc_jac + (size_t)col * G + g0;
            float b0 = 0.0f, bj = 0.0f;
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4) {
              if (j4 == 0 || b4[j4] != b4[j4 - 1]) {
                const float* t = tp_s + b4[j4] * p.ntemp;
                const float u0 = t[ps0.idx - 1], u1 = t[ps1.idx - 1];
                b0 = u0 + ps0.frac * (t[ps0.idx] - u0);
                bj = (u1 + ps1.frac * (t[ps1.idx] - u1)) - b0;
              }
              if (g0 + 4 * j4 < G) {  // (the padding g-points of a ragged last job are not part of the arrays)
                *reinterpret_cast<float4*>(ss + 4 * j4) = make_float4(z[32 + 4 * j4] * b0, z[33 + 4 * j4] * b0, z[34 + 4 * j4] * b0, z[35 + 4 * j4] * b0);
                *reinterpret_cast<float4*>(sj + 4 * j4) = make_float4(z[32 + 4 * j4] * bj, z[33 + 4 * j4] * bj, z[34 + 4 * j4] * bj, z[35 + 4 * j4] * bj);
              }
            }
          }
        } else {
          // ---- SW: tau_abs (net 0), tau_ray (net 1): tau = abs + ray, ssa = ray / tau (mod_network_rrtmgp.F90:209-231)
          float w[32];
#pragma unroll
          for (int j = 0; j < 32; j += 2) {
            float a0 = z[j], a1 = z[j + 1], r0 = z[32 + j], r1 = z[33 + j];
            mul2(a0, a1, a0, a1); mul2(a0, a1, a0, a1); mul2(a0, a1, a0, a1); mul2(a0, a1, cdp, cdp);
            mul2(r0, r1, r0, r1); mul2(r0, r1, r0, r1); mul2(r0, r1, r0, r1); mul2(r0, r1, cdp, cdp);
            z[j] = a0; z[j + 1] = a1; z[32 + j] = r0; z[33 + j] = r1;
          }
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float ry = z[32 + j];
            const float tot = z[j] + ry;
            o[j] = tot;
            // ray / tot: reciprocal seed + two residual corrections (correctly rounded but for rare ties; no branches);
            // no zero guard, as the reference (quirk Q2): 0/0 gives NaN here too
            const float rc = rcp_approx(tot);
            float qd = ry * rc;
            qd = fmaf(fmaf(-tot, qd, ry), rc, qd);
            w[j] = fmaf(fmaf(-tot, qd, ry), rc, qd);
          }
          stage_and_store(&tm0, o, g0, lane, w0, nrows_lay);
          stage_and_store(&tm1, w, g0, lane, w0, nrows_lay);
        }
      }
    }
    __syncwarp();
    if (elect_one()) bulk_wait_all();
  }
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
}

// ------------------------------------------------------------------------------------------------ host side
// pack O output columns of one layer W (row-major [K][ldw] fp32, scaled per output) into the canonical K-major layout of
// the B operand: rows = outputs (padded to OP with zeros), K padded to KP; hi block then lo block.  bias_k >= 0 puts
// bias[o] on that k.
static void pack_layer(const float* W, int K, int O, int ldw, int KP, int OP, const double* oscale, int bias_k, const double* bias,
                       std::vector<uint8_t>& out, uint32_t& off_hi, uint32_t& off_lo) {
  const size_t base = out.size();
  out.resize(base + (size_t)2 * OP * KP * 2, 0);
  off_hi = (uint32_t)base;
  off_lo = (uint32_t)(base + (size_t)OP * KP * 2);
  __half* hi = reinterpret_cast<__half*>(out.data() + off_hi);
  __half* lo = reinterpret_cast<__half*>(out.data() + off_lo);
  for (int o = 0; o < OP; ++o)
    for (int k = 0; k < KP; ++k) {
      double wd = 0.0;
      if (o < O && k < K) wd = (double)W[(size_t)k * ldw + o] * (oscale ? oscale[o] : 1.0);
      else if (o < O && k == bias_k) wd = bias[o];
      const float w = (float)wd;
      const __half h = __float2half_rn(w);
      const __half l = __float2half_rn(w - __half2float(h));
      const size_t idx = ((size_t)(k >> 3) * (OP * 16) + (o >> 3) * 128 + (o & 7) * 16 + (k & 7) * 2) / 2;
      hi[idx] = h;
      lo[idx] = l;
    }
}

struct HostPlan {
  Params p;
  std::vector<uint8_t> blob;
  size_t smem = 0;
};

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
// [rows][ngpt] fp32 tensor, box 32 g-points x box_rows rows, 128-byte swizzle (half = true: 16 g-points, 64-byte swizzle)
static int make_map(CUtensorMap* tm, float* base, int G, unsigned long long rows, int box_rows, bool half = false) {
  EncodeTiledFn enc = encode_fn();
  if (!enc) return fail("gas_optics (tensor cores): cuTensorMapEncodeTiled is not available from the driver");
  const cuuint64_t dims[2] = {(cuuint64_t)G, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)G * 4};
  const cuuint32_t box[2] = {half ? 16u : 32u, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         half ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail("gas_optics (tensor cores): cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")");
  return 0;
}

}  // namespace tc
}  // namespace rrnn

using namespace rrnn;

static int pad16(int n) { return (n + 15) & ~15; }
static int pad32(int n) { return (n + 31) & ~31; }

// Is this set of networks supported by the tensor-core kernel?  nmodels = 2: two networks with ngpt outputs each;
// nmodels = 1 (longwave only): one network with 2*ngpt outputs (tau head, Planck-fraction head).
static bool tc_supported(const rrnn_model_t* const* models, int nmodels, const rrnn_kdist_t* kd, int mode) {
  const int ngpt = kd->ngpt;
  if (ngpt % 16 != 0 || ngpt > 256 || ngpt < 32) return false;
  if (mode == 0 && (kd->nbnd > 16 || kd->nbnd < 1)) return false;
  if (nmodels != 2 && !(nmodels == 1 && mode == 0)) return false;
  int hsum = 0;
  for (int n = 0; n < nmodels; ++n) {
    const rrnn_model_t* m = models[n];
    if (!m || m->nlayers != 3) return false;
    if (m->dims[0] > tc::KIN_MAX || m->dims[0] < 4 || m->dims[1] > tc::HMAX || m->dims[2] != m->dims[1]) return false;
    if (m->dims[3] != (nmodels == 1 ? 2 * ngpt : ngpt)) return false;
    if (m->act[2] != RRNN_ACT_LINEAR) return false;
    hsum += pad16(m->dims[1]);
  }
  if (hsum > tc::RING_COL0) return false;  // the hidden accumulators share TMEM columns 0..127
  if (nmodels == 2 && models[0]->dims[0] != models[1]->dims[0]) return false;
  if ((int)models[0]->ymean.size() < ngpt || (int)models[0]->ystd.size() < ngpt) return false;
  if (mode == 1 && (models[1]->ymean.empty() || models[1]->ystd.empty())) return false;
  return true;
}

// Build the shared-memory image (weights, biases, band tables) and the layout; cached per (model set, kdist, mode).
struct TcCache {
  unsigned long long m0 = 0, m1 = 0, kd = 0;  // uids (m1 = 0: one network with two heads)
  int mode = -1, device = -1;
  tc::Params p{};
  size_t smem = 0;
  uint8_t* d_blob = nullptr;
};
static std::vector<TcCache> g_tc_cache;
static std::mutex g_tc_mutex;

// returns 0 and a copy of the plan, -1 if the networks do not fit the kernel's shared-memory design, > 0 on error
static int tc_plan(rrnn_ctx_t* ctx, int mode, const rrnn_kdist_t* kd, const rrnn_model_t* const* models, int nmodels, TcCache* out) {
  std::lock_guard<std::mutex> lock(g_tc_mutex);
  const unsigned long long uid1 = nmodels == 2 ? models[1]->uid : 0ull;
  for (auto& c : g_tc_cache)
    if (c.m0 == models[0]->uid && c.m1 == uid1 && c.kd == kd->uid && c.mode == mode && c.device == ctx->device) { *out = c; return 0; }
  const int G = kd->ngpt, Gp = pad32(G);
  const int nx = models[0]->dims[0];
  const int kin = pad16(nx);
  const bool both = nmodels == 1;
  // Shared memory is what limits the wide networks.  Attempts, cheapest concession first:
  //   0: output bias folded into the last layer even if that costs one more k-step; Planck table in shared memory
  //   1: bias folded only where the padded K has a spare column (else added in the epilogue)
  //   2: like 1, and the Planck table is read from global memory (L1-resident: 12.5 KB for 16 bands)
  for (int attempt = 0; attempt < 3; ++attempt) {
    TcCache c;
    c.m0 = models[0]->uid; c.m1 = uid1; c.kd = kd->uid; c.mode = mode; c.device = ctx->device;
    tc::Params& p = c.p;
    p.nhid = both ? 1 : 2;
    p.totplnk_global = (attempt == 2 && mode == 0) ? 1 : 0;
    std::vector<uint8_t> blob;
    int hid_col = 0;
    for (int n = 0; n < 2; ++n) {
      const rrnn_model_t* m = models[both ? 0 : n];
      tc::NetP& nt = p.net[n];
      const int Hraw = m->dims[1];
      const int O = m->dims[3];            // outputs of the network: G, or 2G for a two-headed one
      const int o0 = (both && n == 1) ? G : 0;  // first output of this head
      nt.Hraw = Hraw; nt.H = pad16(Hraw); nt.ntot = Gp;
      nt.act0 = m->act[0]; nt.act1 = m->act[1];
      nt.hid_col = hid_col;
      if (!(both && n == 1)) hid_col += nt.H;
      // bias of the output layer on a column of ones: free when the padded K has a spare column, otherwise one more
      // k-step (first attempt) or not at all (later attempts, when shared memory is short)
      nt.fold = (Hraw < nt.H) ? 1 : (attempt == 0 ? 1 : 0);
      nt.K3 = (nt.fold && Hraw >= nt.H) ? nt.H + 16 : nt.H;
      const bool tau_type = (mode == 1) || (n == 0);
      std::vector<double> oscale(Gp, 1.0), bias(Gp, 0.0);
      for (int o = 0; o < G; ++o) {
        const double b3 = m->bpack[m->b_off[2] + o0 + o];
        if (tau_type) {
          oscale[o] = (double)m->ystd[o] * tc::OUT_SCALE;
          bias[o] = ((double)m->ystd[o] * b3 + (double)m->ymean[o]) * tc::OUT_SCALE;
        } else {
          bias[o] = b3;
        }
      }
      if (both && n == 1) {
        // second head: same hidden stack and activation operand as head 0
        nt.w[0][0] = p.net[0].w[0][0]; nt.w[0][1] = p.net[0].w[0][1];
        nt.w[1][0] = p.net[0].w[1][0]; nt.w[1][1] = p.net[0].w[1][1];
      } else {
        tc::pack_layer(m->wpack.data() + m->w_off[0], nx, Hraw, Hraw, kin, nt.H, nullptr, -1, nullptr, blob, nt.w[0][0], nt.w[0][1]);
        tc::pack_layer(m->wpack.data() + m->w_off[1], Hraw, Hraw, Hraw, nt.H, nt.H, nullptr, -1, nullptr, blob, nt.w[1][0], nt.w[1][1]);
      }
      // output layer of this head: columns o0 .. o0+G-1 of the (Hraw, O) matrix, rows padded to Gp with zeros
      tc::pack_layer(m->wpack.data() + m->w_off[2] + o0, Hraw, G, O, nt.K3, Gp, oscale.data(), nt.fold ? Hraw : -1, bias.data(), blob,
                     nt.w[2][0], nt.w[2][1]);
      // fp32 biases: b1[H], b2[H], b3''[Gp]
      const size_t fb = blob.size();
      blob.resize(fb + (size_t)(2 * nt.H + Gp) * 4, 0);
      float* f = reinterpret_cast<float*>(blob.data() + fb);
      for (int i = 0; i < Hraw; ++i) { f[i] = m->bpack[m->b_off[0] + i]; f[nt.H + i] = m->bpack[m->b_off[1] + i]; }
      for (int o = 0; o < G; ++o) f[2 * nt.H + o] = (float)bias[o];
      nt.b[0] = (uint32_t)fb; nt.b[1] = (uint32_t)(fb + nt.H * 4); nt.b[2] = (uint32_t)(fb + 2 * nt.H * 4);
    }
    // band of every group of 4 g-points (the kernel multiplies 4 g-points by one Planck value: mixed groups -> fallback);
    // the padding groups of a ragged last job repeat the last band
    p.off_band4 = (uint32_t)blob.size();
    blob.resize(blob.size() + 64 * 4, 0);
    {
      int* band4 = reinterpret_cast<int*>(blob.data() + p.off_band4);
      for (int g4 = 0; g4 < Gp / 4; ++g4) {
        const int gq = std::min(4 * g4, G - 4);
        const int b = kd->gpt2band.empty() ? 0 : kd->gpt2band[gq];
        if (mode == 0)
          for (int e = 1; e < 4; ++e)
            if (kd->gpt2band[gq + e] != b) return -1;
        band4[g4] = b;
      }
    }
    // LW: the Planck table totplnk [band][ntemp] rides in the image too (unless shared memory is short)
    p.off_totplnk = (uint32_t)blob.size();
    if (mode == 0 && !p.totplnk_global) {
      blob.resize(blob.size() + (((size_t)kd->nbnd * kd->ntemp * 4 + 15) & ~(size_t)15), 0);
      memcpy(blob.data() + p.off_totplnk, kd->totplnk.data(), (size_t)kd->nbnd * kd->ntemp * 4);
    }
    p.blob_bytes = (uint32_t)blob.size();
    uint32_t off = (p.blob_bytes + 127u) & ~127u;
    p.off_ain_hi = off; off += tc::TM * kin * 2;
    p.off_ain_lo = off; off += tc::TM * kin * 2;
    for (int n = 0; n < p.nhid; ++n) {
      p.net[n].act_hi = off; off += tc::TM * p.net[n].K3 * 2;
      p.net[n].act_lo = off; off += tc::TM * p.net[n].K3 * 2;
    }
    if (both) { p.net[1].act_hi = p.net[0].act_hi; p.net[1].act_lo = p.net[0].act_lo; }
    off = (off + 1023u) & ~1023u;
    p.off_stage = off;
    const uint32_t off_stage_end2 = off + 8 * 2 * tc::STAGE_BYTES;  // 8 epilogue warps x 2 staging tiles, if they fit
    p.nstage = 2; off = off_stage_end2;
    // Per-row records, a ring over tiles.  The front warps start tile t+2 only after every output-layer MMA of tile t
    // has been issued, i.e. when the epilogue (either group) is at most NSLOT jobs short of the end of tile t, which is
    // at most ceil(NSLOT/nchunks) tiles back.
    const int nchunks = Gp / 32;
    p.nrec = (tc::NSLOT + nchunks - 1) / nchunks + 2;  // (the padded job count of the kernel is >= nchunks: this stays an upper bound)
    p.off_rec = off; off += (uint32_t)p.nrec * tc::REC_F * tc::TM * 4;
    p.off_bar = off; off += tc::NBAR * 8 + 16;
    c.smem = (size_t)off + 1024;  // alignment slack
    if (c.smem > ctx->smem_optin) {  // one staging tile per warp instead of two
      const uint32_t shrink = 8 * tc::STAGE_BYTES;
      p.nstage = 1; p.off_rec -= shrink; p.off_bar -= shrink; c.smem -= shrink;
    }
    if (c.smem > ctx->smem_optin) continue;
    p.kin = kin; p.nx = nx; p.ngpt = G; p.nchunks = nchunks;
    p.nbnd = kd->nbnd; p.ntemp = kd->ntemp;
    RRNN_CUDA(cudaMalloc((void**)&c.d_blob, blob.size()));
    RRNN_CUDA(cudaMemcpy(c.d_blob, blob.data(), blob.size(), cudaMemcpyHostToDevice));
    p.blob = c.d_blob;
    if (g_tc_cache.size() >= 32) {  // bounded: drop the oldest plan
      cudaFree(g_tc_cache.front().d_blob);
      g_tc_cache.erase(g_tc_cache.begin());
    }
    g_tc_cache.push_back(c);
    *out = c;
    return 0;
  }
  return -1;
}

// Does the tensor-core kernel take this configuration?  (pipeline.cu decides the workspace layout with it)
bool rrnn_gas_optics_tc_can(const rrnn_ctx_t* ctx, int mode, const rrnn_kdist_t* kd, const rrnn_model_t* const* models, int nmodels, int nlay,
                            bool compact) {
  if (!ctx->nn_tensor_cores || nmodels < 1 || nmodels > 2 || !models[0] || (nmodels == 2 && !models[1])) return false;
  if (!tc_supported(models, nmodels, kd, mode)) return false;
  if (mode == 0 && nlay < 31) return false;
  if (compact && (mode != 0 || kd->nbnd > 16)) return false;
  return true;
}

// Launch the tensor-core gas optics; returns -1 if the configuration is not supported (caller falls back).
int rrnn_gas_optics_tc(rrnn_ctx_t* ctx, int mode, const rrnn_kdist_t* kd, const rrnn_model_t* const* models, int nmodels, int ncol, int nlay,
                       const float* play, const float* plev, const float* tlay, const float* tlev, const float* tsfc,
                       const rrnn_gas_t* gases, int ngas, float* out0, float* out1, float* out2, float* sfc_source,
                       float* sfc_jac, int prof_kind, float* planck_lay, float* planck_lev) {
  if (nmodels < 1 || nmodels > 2 || !tc_supported(models, nmodels, kd, mode)) return -1;
  const bool compact = planck_lay != nullptr;  // LW only: out1 receives the Planck fraction, out2 is not written
  if (compact && (mode != 0 || !planck_lev || kd->nbnd > 16)) return -1;
  const int period = (mode == 0) ? nlay + 1 : nlay;
  if (mode == 0 && nlay < 31) return -1;  // at most one extra (bottom-level) row per warp of 32 rows
  if ((long long)ncol * period >= (1LL << 31) - tc::TM) return -1;
  TcCache cache;
  {
    const int rc = tc_plan(ctx, mode, kd, models, nmodels, &cache);
    if (rc != 0) return rc;
  }
  tc::Params p = cache.p;
  const rrnn_model_t* m = models[0];
  const int nx = m->dims[0];
  // map ty_gas_concs entries onto the network's inputs BY NAME (compute_nn_inputs :708-760)
  for (int i = 0; i < tc::KIN_MAX; ++i) { p.gas[i].ptr = nullptr; p.gas[i].mode = 0; p.xmin[i] = 0.f; p.xmax[i] = 1.f; p.xconst[i] = 0.f; p.xvar[i] = 0; }
  p.h2o_const = 0.0f;
  for (int i = 0; i < nx; ++i) {
    p.xmin[i] = m->xmin[i]; p.xmax[i] = m->xmax[i];
    if (i < 2) { p.xvar[i] = 1; continue; }
    int found = -1;
    for (int g = 0; g < ngas; ++g) {
      std::string nm(gases[g].name, strnlen(gases[g].name, 32));
      while (!nm.empty() && (nm.back() == ' ' || nm.back() == '\0')) nm.pop_back();
      size_t b = 0;
      while (b < nm.size() && nm[b] == ' ') ++b;
      if (nm.substr(b) == m->input_names[i]) { found = g; break; }
    }
    if (found < 0) {
      if (i < 4) return fail(std::string("compute_nn_inputs: gas ") + m->input_names[i] + " is required but was not provided");
      p.xconst[i] = (0.0f - p.xmin[i]) / (p.xmax[i] - p.xmin[i]);  // missing gas: vmr 0 (:757-759)
      continue;
    }
    const rrnn_gas_t& gs = gases[found];
    if (gs.ndims < 0 || gs.ndims > 2) return fail("gas_concs: ndims must be 0, 1 or 2");
    if (gs.ndims > 0 && !gs.conc) return fail("gas_concs: null concentration pointer");
    if (gs.ndims == 0) {
      float raw = gs.value;
      if (i == 2) p.h2o_const = raw;
      if (i == 2 || i == 3) raw = sqrtf(sqrtf(raw));
      p.xconst[i] = (raw - p.xmin[i]) / (p.xmax[i] - p.xmin[i]);
    } else {
      p.xvar[i] = 1; p.gas[i].ptr = gs.conc; p.gas[i].mode = gs.ndims;
    }
  }
  p.ncol = ncol; p.nlay = nlay; p.period = period; p.nrows = (unsigned)((long long)ncol * period);
  p.play = play; p.plev = plev; p.tlay = tlay; p.tlev = tlev; p.tsfc = tsfc;
  p.totplnk = kd->d_totplnk; p.temp_ref_min = kd->temp_ref_min; p.totplnk_delta = kd->totplnk_delta;
  p.sfc_source = sfc_source; p.sfc_jac = sfc_jac;
  p.planck_lay = planck_lay; p.planck_lev = planck_lev;
  CUtensorMap tm0, tm1, tm2, tm0s, tm1s;
  const unsigned long long nrows_lay = (unsigned long long)ncol * nlay;
  const bool half = p.nstage == 1;  // one staging tile per warp: stored as two half tiles (see stage_and_store)
  if (int rc = tc::make_map(&tm0, out0, kd->ngpt, nrows_lay, 32, half)) return rc;
  if (int rc = tc::make_map(&tm1, out1, kd->ngpt, nrows_lay, 32, half)) return rc;
  if (mode == 0) {
    if (compact) tm2 = tm1;
    else if (int rc = tc::make_map(&tm2, out2, kd->ngpt, (unsigned long long)ncol * (nlay + 1), 32, half)) return rc;
    if (int rc = tc::make_map(&tm0s, out0, kd->ngpt, nrows_lay, 31, half)) return rc;
    if (int rc = tc::make_map(&tm1s, out1, kd->ngpt, nrows_lay, 31, half)) return rc;
  } else {
    tm2 = tm1; tm0s = tm0; tm1s = tm1;
  }
  if (mode == 1 && out2)  // g is identically zero on the NN path (mo_gas_optics_rrtmgp.F90:560-567)
    RRNN_CUDA(cudaMemsetAsync(out2, 0, (size_t)ncol * nlay * kd->ngpt * sizeof(float), ctx->stream));
  const unsigned ntiles = (p.nrows + tc::TM - 1) / tc::TM;
  const unsigned grid = std::min<unsigned>(ntiles, (unsigned)ctx->num_sms);
  static unsigned* dbg_host = nullptr;
  const char* dbg_env = getenv("RRNN_TC_DEBUG");
  const int dbg_flags = dbg_env ? atoi(dbg_env) : 0;
  p.dbg = nullptr; p.dbg_flags = dbg_flags;
  if (dbg_flags) {
    if (!dbg_host) RRNN_CUDA(cudaHostAlloc((void**)&dbg_host, 128 * sizeof(unsigned), cudaHostAllocMapped));
    memset(dbg_host, 0, 128 * sizeof(unsigned));
    unsigned* dptr = nullptr;
    RRNN_CUDA(cudaHostGetDevicePointer((void**)&dptr, dbg_host, 0));
    p.dbg = dptr;
  }
  const int ps = prof_begin(ctx, prof_kind);
  if (mode == 0 && compact) {
    RRNN_CUDA(cudaFuncSetAttribute(tc::gas_optics_tc_kernel<0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cache.smem));
    tc::gas_optics_tc_kernel<0, true><<<grid, tc::THREADS, cache.smem, ctx->stream>>>(p, tm0, tm1, tm2, tm0s, tm1s);
  } else if (mode == 0) {
    RRNN_CUDA(cudaFuncSetAttribute(tc::gas_optics_tc_kernel<0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cache.smem));
    tc::gas_optics_tc_kernel<0, false><<<grid, tc::THREADS, cache.smem, ctx->stream>>>(p, tm0, tm1, tm2, tm0s, tm1s);
  } else {
    RRNN_CUDA(cudaFuncSetAttribute(tc::gas_optics_tc_kernel<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cache.smem));
    tc::gas_optics_tc_kernel<1, false><<<grid, tc::THREADS, cache.smem, ctx->stream>>>(p, tm0, tm1, tm2, tm0s, tm1s);
  }
  prof_end(ctx, prof_kind, ps);
  RRNN_LAUNCH_CHECK(ctx);
  ctx->last_nn_kernel = RRNN_NN_KERNEL_TCGEN05;
  ctx->nn_tc_launches++;
  if (dbg_flags) {
    const cudaError_t e = cudaStreamSynchronize(ctx->stream);
    fprintf(stderr, "[tc debug] mode %d ncol %d nlay %d grid %u smem %zu -> %s | front %08x mma %08x epi %08x store %08x | wd %08x bar %08x par %u blk %u | init %08x\n",
            mode, ncol, nlay, grid, cache.smem, cudaGetErrorString(e), dbg_host[0], dbg_host[1], dbg_host[2], dbg_host[3], dbg_host[16],
            dbg_host[17], dbg_host[18], dbg_host[19], dbg_host[15]);
    if (dbg_flags & 8) {  // timeline of tile 5 of block 0, SM cycles relative to the front warps' start of that tile
      const unsigned t0 = dbg_host[32];
      static const char* names[60] = {"F prologue start", "F AIN arrive", "F L1n0 free ok", "F L1n0 hid ok", "F L1n0 act arrive", "F L1n1 free ok",
        "F L1n1 hid ok", "F L1n1 act arrive", "F L2n0 -", "F L2n0 hid ok", "F L2n0 act arrive", "F L2n1 -", "F L2n1 hid ok", "F L2n1 act arrive",
        "", "", "", "", "", "", "M loop top", "M AIN ok", "M L1 issued", "M act0 ok", "M L2n0 issued", "M act1 ok", "M L2n1 issued", "M act(2) ok",
        "M job0 issued", "M job1 issued", "M job2 issued", "M job3 issued", "M job4 issued", "M job5 issued", "M job6 issued", "M job7 issued",
        "", "", "", "", "E job0 wait", "E job0 go", "E job1 wait", "E job1 go", "E job2 wait", "E job2 go", "E job3 wait", "E job3 go", "E job4 wait",
        "E job4 go", "E job5 wait", "E job5 go", "E job6 wait", "E job6 go", "E job7 wait", "E job7 go", "", "", "", ""};
      for (int i = 0; i < 60; ++i)
        if (dbg_host[32 + i]) fprintf(stderr, "  [tl] %-20s %8d\n", names[i], (int)(dbg_host[32 + i] - t0));
    }
    if (e != cudaSuccess) return fail(std::string("gas_optics (tensor cores): ") + cudaGetErrorString(e));
  }
  return 0;
}
