// NN gas optics on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a only.
//
// Same fused path as gas_optics_nn.cu (inputs -> MLP chain -> tau / Planck-source / ssa epilogues, every output
// written once) with the three GEMMs of each network on tcgen05.mma:
//   * one CTA per SM (256 threads) walks tiles of 128 samples = the 128 TMEM lanes;
//   * operands live in shared memory in the canonical K-major no-swizzle UMMA layout (8 x 16-byte core matrices);
//     weights are staged once per CTA, activations are written by the epilogue threads of the previous layer;
//   * precision: every fp32 value v is split v = hi + lo with hi = fp16(v), lo = fp16(v - hi) (22 mantissa bits
//     together) and each GEMM is issued as hi*Whi + lo*Whi + hi*Wlo on kind::f16 with fp32 accumulation in TMEM --
//     fp32-class accuracy at fp16 tensor rate (the dropped lo*lo term is < 2^-22 relative);
//   * accumulators: 256 TMEM columns for the output layer, 64 for the hidden layers; tcgen05.ld (32x32b) brings one
//     sample row x 32 g-points per thread into registers; bias, softsign, (ystd*z+ymean)^8*N_dry, pfrac^2 * Planck,
//     tau_abs+tau_ray / ssa are applied there;
//   * stores: each warp transposes its 32 rows x 32 g-points through a private padded shared-memory tile so that
//     every st.global.v4 covers full 128-byte lines (4 rows x 128 B per instruction).
// One elected thread issues the MMAs and commits them to an mbarrier; the 8 warps are the epilogue (two warps per
// TMEM lane quarter, each taking half of the columns).
//
// Supported: 2 hidden layers, hidden width <= 64, <= 32 inputs, ngpt a multiple of 32 and <= 256 (all g256/g224/g128
// two-network models of the reference); anything else falls back to the fp32 FFMA kernel.
#include "common.cuh"
#include <cuda_fp16.h>

namespace rrnn {

int map_gases(const rrnn_model_t* m, const rrnn_gas_t* gases, int ngas, struct GoParams& p);  // gas_optics_nn.cu

namespace tc {

constexpr int TM = 128;  // samples per tile (TMEM lanes)
constexpr int THREADS = 256;
constexpr int KIN = 32;  // padded number of network inputs
constexpr int STAGE_LD = 36;  // floats per staging row (32 + 4 pad): conflict-free 16-byte accesses

struct Net {
  int H, N;            // padded hidden width (16..64, multiple of 16), outputs (multiple of 32)
  int act[3];
  int w_bytes;         // bytes of the packed weights (hi/lo blocks of the three layers)
  const uint8_t* w;    // device: [W1hi][W1lo][W2hi][W2lo][W3hi][W3lo], canonical layout, fp16
  const float* b;      // device: b1[H] b2[H] b3[N]
  const float* ymean;  // device, N (or null)
  const float* ystd;
};

struct GasIn {
  const float* ptr;
  float value;
  int mode;
};

struct Params {
  int mode;  // 0 = LW (tau net + Planck-fraction net), 1 = SW (absorption net + Rayleigh net)
  int ncol, nlay, ngpt, nx;
  long long nsamples;
  const float *play, *plev, *tlay, *tlev, *tsfc;
  GasIn gas[KIN];
  float xmin[KIN], xmax[KIN];
  float xconst[KIN];  // scaled value of inputs that do not vary per sample (scalar gases, missing gases, padding)
  int xvar[KIN];      // 1 = varies per sample (tlay, play, 1-D / 2-D gas fields)
  Net net[2];
  int nbnd, ntemp;
  const int* gpt2band;
  const float* totplnk;
  float temp_ref_min, totplnk_delta;
  float *out0, *out1, *out2, *sfc_source, *sfc_jac;
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
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {}
}
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n\t"
      "tcgen05.wait::ld.sync.aligned;"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n\t"
      "tcgen05.wait::ld.sync.aligned;"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// canonical K-major no-swizzle layout of an operand with R rows: byte offset of the 16-byte unit (row r, k-unit ku)
__device__ __forceinline__ uint32_t unit_off(int R, int r, int ku) { return (uint32_t)ku * (R * 16) + (r >> 3) * 128 + (r & 7) * 16; }

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
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(d));
  r = fmaf(fmaf(-d, r, 1.0f), r, r);
  return x * r;
}

__device__ __forceinline__ float act_apply(int code, float x) {
  switch (code) {
    case RRNN_ACT_SOFTSIGN: return x / (fabsf(x) + 1.0f);
    case RRNN_ACT_RELU: return fmaxf(0.0f, x);
    case RRNN_ACT_SIGMOID: return 1.0f / (1.0f + expf(-x));
    case RRNN_ACT_HARD_SIGMOID: return fmaxf(0.0f, fminf(1.0f, 0.2f * x + 0.5f));
    default: return x;
  }
}

__device__ __forceinline__ float planck_interp(float T, float tmin, float delta, const float* __restrict__ tab, int ntemp) {
  const float val0 = (T - tmin) / delta;
  const int iv = (int)val0;
  const float frac = val0 - (float)iv;
  const int idx = min(ntemp - 1, max(1, iv + 1));
  const float t0 = __ldg(tab + idx - 1);
  return t0 + frac * (__ldg(tab + idx) - t0);
}

// Issue D[128 x N] (+)= A[128 x K] * W[N x K]^T as the three split products; A hi/lo and W hi/lo in canonical layout.
__device__ __forceinline__ void issue_gemm(uint32_t tmem_d, uint32_t a_hi, uint32_t a_lo, uint32_t w_hi, uint32_t w_lo, int K, int N) {
  const uint32_t idesc = make_idesc(TM, N);
  const int ksteps = K >> 4;
  uint32_t acc = 0;
  for (int s = 0; s < ksteps; ++s) {
    // one k-step = 16 fp16 = two 16-byte k-units; operands advance by two units
    const uint32_t ao = (uint32_t)(2 * s) * (TM * 16), wo = (uint32_t)(2 * s) * (N * 16);
    const uint64_t dah = make_desc(a_hi + ao, TM * 16, 128), dal = make_desc(a_lo + ao, TM * 16, 128);
    const uint64_t dwh = make_desc(w_hi + wo, N * 16, 128), dwl = make_desc(w_lo + wo, N * 16, 128);
    mma_f16(tmem_d, dah, dwh, idesc, acc);
    acc = 1;
    mma_f16(tmem_d, dal, dwh, idesc, 1);
    mma_f16(tmem_d, dah, dwl, idesc, 1);
  }
}

struct NetPlan {
  uint32_t w[3][2];  // shared-memory byte offsets of W{1,2,3}{hi,lo}
  uint32_t b[3];     // float offsets of b1, b2, b3 in the bias area
  uint32_t ystd, ymean;
};

template <int MODE>
__global__ void __launch_bounds__(THREADS, 1) gas_optics_tc_kernel(const Params p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int quarter = warp & 3;    // TMEM lane quarter this warp may access
  const int half = warp >> 2;      // which half of the columns this warp handles
  const int row = 32 * quarter + lane;
  const int L = p.nlay, G = p.ngpt;

  // ---------------------------------------------------------------------------------- shared-memory carve-up
  uint32_t off = 0;
  NetPlan plan[2];
  uint32_t woff[2];
  for (int n = 0; n < 2; ++n) {
    const int H = p.net[n].H, N = p.net[n].N;
    woff[n] = off;
    uint32_t o = off;
    plan[n].w[0][0] = o; o += H * KIN * 2; plan[n].w[0][1] = o; o += H * KIN * 2;
    plan[n].w[1][0] = o; o += H * H * 2;   plan[n].w[1][1] = o; o += H * H * 2;
    plan[n].w[2][0] = o; o += N * H * 2;   plan[n].w[2][1] = o; o += N * H * 2;
    off = o;
  }
  float* fl = reinterpret_cast<float*>(smem + off);
  uint32_t fo = 0;
  for (int n = 0; n < 2; ++n) {
    const int H = p.net[n].H, N = p.net[n].N;
    plan[n].b[0] = fo; fo += H; plan[n].b[1] = fo; fo += H; plan[n].b[2] = fo; fo += N;
    plan[n].ystd = fo; fo += N; plan[n].ymean = fo; fo += N;
  }
  off += fo * 4;
  uint8_t* ain_hi = smem + off; off += TM * KIN * 2;
  uint8_t* ain_lo = smem + off; off += TM * KIN * 2;
  uint8_t* act_hi = smem + off; off += TM * 64 * 2;
  uint8_t* act_lo = smem + off; off += TM * 64 * 2;
  float* stage = reinterpret_cast<float*>(smem + off) + warp * (32 * STAGE_LD); off += 8 * 32 * STAGE_LD * 4;
  float* coldry_s = reinterpret_cast<float*>(smem + off); off += TM * 4;
  int* levrow_s = reinterpret_cast<int*>(smem + off); off += TM * 4;   // row index into lev_source; -1 = sample beyond the end
  int* flag_s = reinterpret_cast<int*>(smem + off); off += TM * 4;
  int* band_s = reinterpret_cast<int*>(smem + off); off += 256 * 4;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + off); off += 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + off); off += 8;

  // ---------------------------------------------------------------------------------- one-time set-up
  for (int n = 0; n < 2; ++n) {
    const uint4* src = reinterpret_cast<const uint4*>(p.net[n].w);
    uint4* dst = reinterpret_cast<uint4*>(smem + woff[n]);
    for (int i = tid; i < p.net[n].w_bytes / 16; i += THREADS) dst[i] = src[i];
    const int H = p.net[n].H, N = p.net[n].N;
    for (int i = tid; i < 2 * H + N; i += THREADS) fl[plan[n].b[0] + i] = p.net[n].b[i];
    for (int i = tid; i < N; i += THREADS) {
      fl[plan[n].ystd + i] = p.net[n].ystd ? p.net[n].ystd[i] : 0.0f;
      fl[plan[n].ymean + i] = p.net[n].ymean ? p.net[n].ymean[i] : 0.0f;
    }
  }
  for (int i = tid; i < 256; i += THREADS) band_s[i] = (MODE == 0 && i < G) ? p.gpt2band[i] : 0;
  if (tid == 0) mbar_init(smem_u32(bar), 1);
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  fence_async_smem();
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_row = tmem_base + ((uint32_t)(32 * quarter) << 16);  // this warp's lanes
  uint32_t phase = 0;
  const uint32_t bar_a = smem_u32(bar);

  int sfc_lay0 = -1;
  if (MODE == 0) sfc_lay0 = (p.play[0] > p.play[L - 1]) ? 0 : L - 1;  // merge(1,nlay,play(1,1) > play(nlay,1))

  const long long ntiles = (p.nsamples + TM - 1) / TM;
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long s0 = tile * TM;
    // ------------------------------------------------------------------------------ prologue: inputs, col_dry
    {
      const int r = tid & 127, part = tid >> 7;  // two threads per sample: inputs [0,16) and [16,32)
      const long long smp = s0 + r;
      const bool ok = smp < p.nsamples;
      const long long col = ok ? smp / L : 0;
      const int lay = ok ? (int)(smp - col * L) : 0;
      float v[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int k = 16 * part + j;
        float x = 0.0f;
        if (ok && k < p.nx) {
          float raw;
          if (k == 0) raw = p.tlay[smp];
          else if (k == 1) raw = logf(p.play[smp]);
          else {
            const GasIn gi = p.gas[k];
            raw = (gi.mode == 2) ? gi.ptr[smp] : (gi.mode == 1 ? gi.ptr[lay] : (gi.mode == 0 ? gi.value : 0.0f));
            if (k == 2 || k == 3) raw = sqrtf(sqrtf(raw));
          }
          x = (raw - p.xmin[k]) / (p.xmax[k] - p.xmin[k]);
        }
        v[j] = x;
      }
      store_split8(ain_hi, ain_lo, unit_off(TM, r, 2 * part), v);
      store_split8(ain_hi, ain_lo, unit_off(TM, r, 2 * part + 1), v + 8);
      if (part == 0) {
        float cd = 0.0f;
        int fl_ = 0, lr = -1;
        if (ok) {
          const GasIn gh = p.gas[2];
          const float h = (gh.mode == 2) ? gh.ptr[smp] : (gh.mode == 1 ? gh.ptr[lay] : gh.value);
          const float dp = fabsf(p.plev[col * (L + 1) + lay] - p.plev[col * (L + 1) + lay + 1]);
          const float fact = 1.0f / (1.0f + h);
          const float m_air = (0.028964f + 0.018016f * h) * fact;
          cd = 10.0f * dp * 6.02214076e23f * fact / (1000.0f * m_air * 100.0f * 9.80665f);
          if (lay == L - 1) fl_ |= 1;
          if (lay == sfc_lay0) fl_ |= 2;
          lr = (int)(col * (L + 1) + lay);
        }
        coldry_s[r] = cd; flag_s[r] = fl_; levrow_s[r] = lr;
      }
    }

    for (int n = 0; n < 2; ++n) {
      const Net& net = p.net[n];
      const int H = net.H, N = net.N;
      // TMEM columns: output-layer accumulator at 0 (LW, SW net 0) or 256 (SW net 1); hidden accumulator beside it
      const uint32_t col_out = (MODE == 1 && n == 1) ? 256u : 0u;
      const uint32_t col_hid = (MODE == 1 && n == 1) ? (uint32_t)p.net[0].N : 256u;
      // ---------------------------------------------------------------------------- two hidden layers
      for (int l = 0; l < 2; ++l) {
        fence_async_smem();
        fence_before();
        __syncthreads();
        if (tid == 0) {
          fence_after();
          if (l == 0) issue_gemm(tmem_base + col_hid, smem_u32(ain_hi), smem_u32(ain_lo), smem_u32(smem + plan[n].w[0][0]), smem_u32(smem + plan[n].w[0][1]), KIN, H);
          else issue_gemm(tmem_base + col_hid, smem_u32(act_hi), smem_u32(act_lo), smem_u32(smem + plan[n].w[1][0]), smem_u32(smem + plan[n].w[1][1]), H, H);
          mma_commit(bar_a);
        }
        mbar_wait(bar_a, phase); phase ^= 1;
        fence_after();
        // epilogue: bias + activation, split, write the next A operand (16 columns per step)
        const float* bb = fl + plan[n].b[l];
        const int cols_per_half = (H >= 32) ? H / 2 : H;
        if (H >= 32 || half == 0) {
          for (int c0 = half * cols_per_half; c0 < (half + 1) * cols_per_half && c0 < H; c0 += 16) {
            float v[16];
            tmem_ld16(tmem_row + col_hid + c0, v);
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = act_apply(net.act[l], v[j] + bb[c0 + j]);
            store_split8(act_hi, act_lo, unit_off(TM, row, c0 >> 3), v);
            store_split8(act_hi, act_lo, unit_off(TM, row, (c0 >> 3) + 1), v + 8);
          }
        }
      }
      // ---------------------------------------------------------------------------- output layer
      fence_async_smem();
      fence_before();
      __syncthreads();
      if (tid == 0) {
        fence_after();
        issue_gemm(tmem_base + col_out, smem_u32(act_hi), smem_u32(act_lo), smem_u32(smem + plan[n].w[2][0]), smem_u32(smem + plan[n].w[2][1]), H, N);
        mma_commit(bar_a);
      }
      mbar_wait(bar_a, phase); phase ^= 1;
      fence_after();
      if (MODE == 1 && n == 0) continue;  // SW: the absorption accumulator waits in TMEM for the Rayleigh network

      // ---------------------------------------------------------------------------- output epilogue
      const int nchunks = N >> 5;
      const int c_begin = half ? (nchunks + 1) / 2 : 0;
      const int c_end = half ? nchunks : (nchunks + 1) / 2;
      const long long smp = s0 + row;
      const bool ok = smp < p.nsamples;
      const float cd = coldry_s[row];
      const int fl_ = flag_s[row];
      const long long col = ok ? smp / L : 0;
      const int lay = ok ? (int)(smp - col * L) : 0;
      const float* b3 = fl + plan[n].b[2];
      for (int ch = c_begin; ch < c_end; ++ch) {
        const int g0 = 32 * ch;
        float z[32];
        tmem_ld32(tmem_row + col_out + g0, z);
        if (MODE == 1) {
          // ---- SW: tau_abs (net 0, TMEM columns 0..) and tau_ray (net 1): tau_tot, ssa (mod_network_rrtmgp.F90:209-231)
          float za[32];
          tmem_ld32(tmem_row + g0, za);
          const float* b3a = fl + plan[0].b[2];
          const float *ysa = fl + plan[0].ystd, *yma = fl + plan[0].ymean, *ysr = fl + plan[1].ystd, *ymr = fl + plan[1].ymean;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float a = ysa[g0 + j] * (za[j] + b3a[g0 + j]) + yma[g0 + j];
            a = a * a; a = a * a; a = a * a;
            a = a * cd;
            float r = ysr[g0 + j] * (z[j] + b3[g0 + j]) + ymr[g0 + j];
            r = r * r; r = r * r; r = r * r;
            r = r * cd;
            const float tot = a + r;
            za[j] = tot;
            z[j] = r / tot;
          }
          // tau_tot
#pragma unroll
          for (int j = 0; j < 8; ++j) *reinterpret_cast<float4*>(stage + lane * STAGE_LD + 4 * j) = make_float4(za[4 * j], za[4 * j + 1], za[4 * j + 2], za[4 * j + 3]);
          __syncwarp();
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int rr = 4 * j + (lane >> 3), cc = (lane & 7) * 4;
            const long long s2 = s0 + 32 * quarter + rr;
            if (s2 < p.nsamples) st_stream4(reinterpret_cast<float4*>(p.out0 + s2 * G + g0 + cc), *reinterpret_cast<const float4*>(stage + rr * STAGE_LD + cc));
          }
          __syncwarp();
          // ssa
#pragma unroll
          for (int j = 0; j < 8; ++j) *reinterpret_cast<float4*>(stage + lane * STAGE_LD + 4 * j) = make_float4(z[4 * j], z[4 * j + 1], z[4 * j + 2], z[4 * j + 3]);
          __syncwarp();
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int rr = 4 * j + (lane >> 3), cc = (lane & 7) * 4;
            const long long s2 = s0 + 32 * quarter + rr;
            if (s2 < p.nsamples) {
              st_stream4(reinterpret_cast<float4*>(p.out1 + s2 * G + g0 + cc), *reinterpret_cast<const float4*>(stage + rr * STAGE_LD + cc));
              if (p.out2) st_stream4(reinterpret_cast<float4*>(p.out2 + s2 * G + g0 + cc), make_float4(0.f, 0.f, 0.f, 0.f));
            }
          }
          __syncwarp();
        } else if (n == 0) {
          // ---- LW tau: ((ystd*(z+b)+ymean)**8)*col_dry
          const float *ys = fl + plan[0].ystd, *ym = fl + plan[0].ymean;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float t = ys[g0 + j] * (z[j] + b3[g0 + j]) + ym[g0 + j];
            t = t * t; t = t * t; t = t * t;
            z[j] = t * cd;
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) *reinterpret_cast<float4*>(stage + lane * STAGE_LD + 4 * j) = make_float4(z[4 * j], z[4 * j + 1], z[4 * j + 2], z[4 * j + 3]);
          __syncwarp();
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int rr = 4 * j + (lane >> 3), cc = (lane & 7) * 4;
            const long long s2 = s0 + 32 * quarter + rr;
            if (s2 < p.nsamples) st_stream4(reinterpret_cast<float4*>(p.out0 + s2 * G + g0 + cc), *reinterpret_cast<const float4*>(stage + rr * STAGE_LD + cc));
          }
          __syncwarp();
        } else {
          // ---- LW Planck fraction -> lay_source, lev_source (+ bottom level, surface source) : compute_Planck_source_nn
          const int lact = net.act[2];
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float zz = act_apply(lact, z[j] + b3[g0 + j]);
            z[j] = zz * zz;  // Planck fraction (:309-312)
          }
          // stage z[j] * B_band(T) for this thread's row; band changes are warp-uniform (same column for all lanes)
          auto stage_scaled = [&](float T, float dT) {
            int cur_band = -1;
            float bv = 0.0f;
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4) {
              float o[4];
#pragma unroll
              for (int c = 0; c < 4; ++c) {
                const int bnd = band_s[g0 + 4 * j4 + c];
                if (bnd != cur_band) {
                  cur_band = bnd;
                  const float* tab = p.totplnk + (size_t)bnd * p.ntemp;
                  bv = planck_interp(T, p.temp_ref_min, p.totplnk_delta, tab, p.ntemp);
                  if (dT != 0.0f) bv = planck_interp(T + dT, p.temp_ref_min, p.totplnk_delta, tab, p.ntemp) - bv;
                }
                o[c] = z[4 * j4 + c] * bv;
              }
              *reinterpret_cast<float4*>(stage + lane * STAGE_LD + 4 * j4) = make_float4(o[0], o[1], o[2], o[3]);
            }
          };
          auto copy_own_row = [&](float* dst) {  // rare rows (bottom level, surface): this thread's staged row -> dst
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4) *reinterpret_cast<float4*>(dst + 4 * j4) = *reinterpret_cast<const float4*>(stage + lane * STAGE_LD + 4 * j4);
          };
          const float t_lay = ok ? p.tlay[smp] : 200.0f;
          const float t_lev = ok ? p.tlev[col * (L + 1) + lay] : 200.0f;
          // lay_source
          stage_scaled(t_lay, 0.0f);
          __syncwarp();
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int rr = 4 * j + (lane >> 3), cc = (lane & 7) * 4;
            const long long s2 = s0 + 32 * quarter + rr;
            if (s2 < p.nsamples) st_stream4(reinterpret_cast<float4*>(p.out1 + s2 * G + g0 + cc), *reinterpret_cast<const float4*>(stage + rr * STAGE_LD + cc));
          }
          __syncwarp();
          // lev_source (rows col*(L+1)+lay)
          stage_scaled(t_lev, 0.0f);
          __syncwarp();
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int rr = 4 * j + (lane >> 3), cc = (lane & 7) * 4;
            const int lr = levrow_s[32 * quarter + rr];
            if (lr >= 0) st_stream4(reinterpret_cast<float4*>(p.out2 + (size_t)lr * G + g0 + cc), *reinterpret_cast<const float4*>(stage + rr * STAGE_LD + cc));
          }
          __syncwarp();
          // the bottom level and the surface terms belong to one sample per column: handled by the owning lane
          if (__any_sync(0xffffffffu, ok && fl_ != 0)) {
            if (ok && (fl_ & 1)) { stage_scaled(p.tlev[col * (L + 1) + L], 0.0f); copy_own_row(p.out2 + ((size_t)col * (L + 1) + L) * G + g0); }
            __syncwarp();
            if (ok && (fl_ & 2)) { stage_scaled(p.tsfc[col], 0.0f); copy_own_row(p.sfc_source + (size_t)col * G + g0); }
            __syncwarp();
            if (ok && (fl_ & 2)) { stage_scaled(p.tsfc[col], 1.0f); copy_own_row(p.sfc_jac + (size_t)col * G + g0); }
            __syncwarp();
          }
        }
      }
    }
    // all TMEM reads and operand buffers of this tile are done before the next tile's prologue overwrites them
    fence_before();
    __syncthreads();
    fence_after();
  }
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
}

// ------------------------------------------------------------------------------------------------ host side
// pack one layer W (row-major [K][O] fp32) into the canonical K-major layout of the B operand: rows = outputs
// (padded to OP), K padded to KP; hi block then lo block.
static void pack_layer(const float* W, int K, int O, int KP, int OP, std::vector<uint8_t>& out) {
  const size_t base = out.size();
  out.resize(base + (size_t)2 * OP * KP * 2, 0);
  __half* hi = reinterpret_cast<__half*>(out.data() + base);
  __half* lo = hi + (size_t)OP * KP;
  for (int o = 0; o < OP; ++o)
    for (int k = 0; k < KP; ++k) {
      const float w = (o < O && k < K) ? W[(size_t)k * O + o] : 0.0f;
      const __half h = __float2half_rn(w);
      const __half l = __float2half_rn(w - __half2float(h));
      const size_t idx = ((size_t)(k >> 3) * (OP * 16) + (o >> 3) * 128 + (o & 7) * 16 + (k & 7) * 2) / 2;
      hi[idx] = h;
      lo[idx] = l;
    }
}

}  // namespace tc
}  // namespace rrnn

using namespace rrnn;

// Is this pair of networks supported by the tensor-core kernel?
static bool tc_supported(const rrnn_model_t* const* models, int nmodels, int ngpt) {
  if (nmodels != 2) return false;
  if (ngpt % 32 != 0 || ngpt > 256) return false;
  for (int n = 0; n < 2; ++n) {
    const rrnn_model_t* m = models[n];
    if (!m || m->nlayers != 3) return false;
    if (m->dims[0] > tc::KIN || m->dims[1] > 64 || m->dims[2] != m->dims[1] || m->dims[3] != ngpt) return false;
  }
  return models[0]->dims[0] == models[1]->dims[0];
}

static int tc_prepare(rrnn_model_t* m) {
  if (m->d_tc_w) return 0;
  const int nx = m->dims[0], Hraw = m->dims[1], N = m->dims[3];
  const int H = Hraw <= 16 ? 16 : (Hraw <= 32 ? 32 : 64);
  std::vector<uint8_t> pack;
  tc::pack_layer(m->wpack.data() + m->w_off[0], nx, Hraw, tc::KIN, H, pack);
  tc::pack_layer(m->wpack.data() + m->w_off[1], Hraw, Hraw, H, H, pack);
  tc::pack_layer(m->wpack.data() + m->w_off[2], Hraw, N, H, N, pack);
  std::vector<float> b((size_t)2 * H + N, 0.0f);
  for (int i = 0; i < Hraw; ++i) { b[i] = m->bpack[m->b_off[0] + i]; b[H + i] = m->bpack[m->b_off[1] + i]; }
  for (int i = 0; i < N; ++i) b[2 * H + i] = m->bpack[m->b_off[2] + i];
  RRNN_CUDA(cudaMalloc((void**)&m->d_tc_w, pack.size()));
  RRNN_CUDA(cudaMemcpy(m->d_tc_w, pack.data(), pack.size(), cudaMemcpyHostToDevice));
  RRNN_CUDA(cudaMalloc((void**)&m->d_tc_b, b.size() * sizeof(float)));
  RRNN_CUDA(cudaMemcpy(m->d_tc_b, b.data(), b.size() * sizeof(float), cudaMemcpyHostToDevice));
  m->tc_w_bytes = (int)pack.size();
  m->tc_H = H;
  return 0;
}

// Launch the tensor-core gas optics; returns -1 if the configuration is not supported (caller falls back).
int rrnn_gas_optics_tc(rrnn_ctx_t* ctx, int mode, const rrnn_kdist_t* kd, const rrnn_model_t* const* models, int ncol, int nlay,
                       const float* play, const float* plev, const float* tlay, const float* tlev, const float* tsfc,
                       const rrnn_gas_t* gases, int ngas, float* out0, float* out1, float* out2, float* sfc_source,
                       float* sfc_jac, int prof_kind) {
  if (!tc_supported(models, 2, kd->ngpt)) return -1;
  if (mode == 1 && models[0]->dims[3] + (models[1]->dims[1] <= 16 ? 16 : (models[1]->dims[1] <= 32 ? 32 : 64)) > 256) return -1;
  tc::Params p{};
  {
    // reuse the by-name gas mapping of the FFMA path through a scratch GoParams-compatible view
    const rrnn_model_t* m = models[0];
    const int nx = m->dims[0];
    p.nx = nx;
    for (int i = 0; i < tc::KIN; ++i) { p.gas[i].ptr = nullptr; p.gas[i].value = 0.f; p.gas[i].mode = -1; p.xmin[i] = 0.f; p.xmax[i] = 1.f; }
    for (int i = 0; i < nx; ++i) {
      p.xmin[i] = m->xmin[i]; p.xmax[i] = m->xmax[i];
      if (i < 2) continue;
      for (int g = 0; g < ngas; ++g) {
        std::string nm(gases[g].name, strnlen(gases[g].name, 32));
        while (!nm.empty() && (nm.back() == ' ' || nm.back() == '\0')) nm.pop_back();
        if (nm == m->input_names[i]) {
          if (gases[g].ndims < 0 || gases[g].ndims > 2) return fail("gas_concs: ndims must be 0, 1 or 2");
          if (gases[g].ndims > 0 && !gases[g].conc) return fail("gas_concs: null concentration pointer");
          p.gas[i].ptr = gases[g].conc; p.gas[i].value = gases[g].value; p.gas[i].mode = gases[g].ndims;
          break;
        }
      }
      if (i < 4 && p.gas[i].mode < 0) return fail(std::string("compute_nn_inputs: gas ") + m->input_names[i] + " is required but was not provided");
    }
  }
  size_t smem = 0;
  for (int n = 0; n < 2; ++n) {
    rrnn_model_t* m = const_cast<rrnn_model_t*>(models[n]);
    if (int rc = tc_prepare(m)) return rc;
    p.net[n].H = m->tc_H; p.net[n].N = m->dims[3];
    for (int l = 0; l < 3; ++l) p.net[n].act[l] = m->act[l];
    p.net[n].w_bytes = m->tc_w_bytes; p.net[n].w = (const uint8_t*)m->d_tc_w; p.net[n].b = m->d_tc_b;
    p.net[n].ymean = m->d_ymean; p.net[n].ystd = m->d_ystd;
    smem += m->tc_w_bytes + (size_t)(2 * m->tc_H + 3 * m->dims[3]) * 4;
  }
  smem += 2 * tc::TM * tc::KIN * 2 + 2 * tc::TM * 64 * 2 + 8 * 32 * tc::STAGE_LD * 4 + 3 * tc::TM * 4 + 256 * 4 + 16;
  smem += 1024;  // alignment slack
  if (smem > ctx->smem_optin) return -1;
  p.mode = mode; p.ncol = ncol; p.nlay = nlay; p.ngpt = kd->ngpt; p.nsamples = (long long)ncol * nlay;
  p.play = play; p.plev = plev; p.tlay = tlay; p.tlev = tlev; p.tsfc = tsfc;
  p.nbnd = kd->nbnd; p.ntemp = kd->ntemp; p.gpt2band = kd->d_gpt2band; p.totplnk = kd->d_totplnk;
  p.temp_ref_min = kd->temp_ref_min; p.totplnk_delta = kd->totplnk_delta;
  p.out0 = out0; p.out1 = out1; p.out2 = out2; p.sfc_source = sfc_source; p.sfc_jac = sfc_jac;
  const long long ntiles = (p.nsamples + tc::TM - 1) / tc::TM;
  const unsigned grid = (unsigned)std::min<long long>(ntiles, ctx->num_sms);
  const int ps = prof_begin(ctx, prof_kind);
  if (mode == 0) {
    RRNN_CUDA(cudaFuncSetAttribute(tc::gas_optics_tc_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    tc::gas_optics_tc_kernel<0><<<grid, tc::THREADS, smem, ctx->stream>>>(p);
  } else {
    RRNN_CUDA(cudaFuncSetAttribute(tc::gas_optics_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    tc::gas_optics_tc_kernel<1><<<grid, tc::THREADS, smem, ctx->stream>>>(p);
  }
  prof_end(ctx, prof_kind, ps);
  RRNN_LAUNCH_CHECK(ctx);
  return 0;
}
