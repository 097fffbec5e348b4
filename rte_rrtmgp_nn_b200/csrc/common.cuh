// Shared definitions of librrnn_b200: context, error handling, small device helpers.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <nvtx3/nvToolsExt.h>   // header-only; resolves the profiler's injection library at run time (no link dependency)
#include "../../include/rrnn.h"

namespace rrnn {

void set_error(const std::string& msg);
int fail(const std::string& msg);

#define RRNN_CUDA(call)                                                                              \
  do {                                                                                               \
    cudaError_t _e = (call);                                                                         \
    if (_e != cudaSuccess)                                                                           \
      return ::rrnn::fail(std::string(#call) + ": " + cudaGetErrorString(_e));                       \
  } while (0)

#define RRNN_CHECK(cond, msg)                                                                        \
  do {                                                                                               \
    if (!(cond)) return ::rrnn::fail(msg);                                                           \
  } while (0)

#define RRNN_LAUNCH_CHECK(ctx)                                                                       \
  do {                                                                                               \
    (ctx)->launches++;                                                                               \
    cudaError_t _e = cudaGetLastError();                                                             \
    if (_e != cudaSuccess) return ::rrnn::fail(std::string("kernel launch: ") + cudaGetErrorString(_e)); \
  } while (0)

// NVTX range carrying the reference's GPTL timer name for the same stretch of work (gptlstart / gptlstop pairs in
// examples/rfmip-clear-sky/rrtmgp_rfmip_{lw,sw}.F90:360-446, examples/all-sky/rrtmgp_allsky.F90:361-440,
// rte/kernels/mo_rte_solver_kernels.F90:168-300, 616-640, rrtmgp/kernels/mo_gas_optics_kernels.F90:725): a timeline of this
// library reads like the reference's GPTL report.  A push / pop costs ~20 ns when no profiler is attached.
struct NvtxRange {
  explicit NvtxRange(const char* gptl_name) { nvtxRangePushA(gptl_name); }
  ~NvtxRange() { nvtxRangePop(); }
  NvtxRange(const NvtxRange&) = delete;
  NvtxRange& operator=(const NvtxRange&) = delete;
};

constexpr int MAX_NN_INPUTS = 32;
constexpr int MAX_LAYERS = 6;
constexpr int MAX_BANDS = 32;

}  // namespace rrnn

// ---- opaque handle definitions -------------------------------------------------------------------
struct rrnn_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  int num_sms = 148;
  size_t smem_optin = 0;
  long long launches = 0;
  int last_nn_kernel = 0;                        // RRNN_NN_KERNEL_* of the most recent NN gas-optics launch
  long long nn_tc_launches = 0, nn_ffma_launches = 0;
  // flags (rte/mo_rte_rrtmgp_config.F90:23-40 + this library's own)
  int lw_source_bug_compat = 1;
  int fast_math = 0;       // solver transcendental variant: 0 = IEEE-accurate libdevice, 1 = ex2/rcp/rsqrt approx
  int sw_fast_math = 1;    // the same for the SW solver only.  Default ON: measured (tools/sw_noise.py) the SW fluxes sit at the same
                           // distance from the fp64 evaluation with and without the Newton refinements (max 6.4e-2 vs 6.8e-2, rms
                           // 9.1e-3 both, strict fp32 oracle 6.4e-2 / 9.4e-3 at 200 x 137): the two-stream formulas' own conditioning
                           // sets the error, not 1-ulp differences of rcp / sqrt / exp; the solver is 11 % faster without them
  int solver_buffer = 0;   // reverse-sweep buffer: 0 auto, 1 shared memory, 2 L2-resident global scratch
  int solver_variant = 0;  // 0 = TMA-staged packed kernels (rte_solvers_tma.cu), 1 = one g-point per lane (rte_solvers.cu)
  int solver_scratch_mb = 0;  // L2 budget of the packed kernels' reverse-sweep scratch (0 = default)
  int solver_warps = 0;       // solvers per CTA in the v5 kernels (0 = default)
  int solver_wide = 1;        // 1: four g-points per lane in the LW solver where the shape fits (lw_solver_v7); 0: lw_solver_v6
  int solver_wide_sw = 0;     // (builds with -DRRNN_EXPERIMENT_SW_WIDE only) 1: four g-points per lane in the clear-sky SW solver too (sw_solver_v7, measured slower)
  int solver_scratch_mb_sw_wide = 0;   // scratch budget of sw_solver_v7 (0: its default)
  void* scratch = nullptr;
  size_t scratch_bytes = 0;
  int* col_counter = nullptr;  // the packed solvers' dynamic column assignment (one int, zeroed before every launch)
  int lw_compact_source = 1;  // fused LW path: sources stay factored between gas optics and solver (8 instead of 12 B per g-point and layer)
  int nn_tensor_cores = 1; // MLP variant: 1 = tcgen05 (fp16 hi/lo split operands, fp32 accumulation; default), 0 = fp32 FFMA
  int chunk_columns = 0;
  int host_copy_threads = 0;  // threads that copy pageable caller memory to / from the pinned bounce ring (0 = min(8, cores))
  int check_extents = 0, check_values = 0;  // rte/mo_rte_rrtmgp_config.F90:23-24, 52-53 (rte_config_checks); default .false. as there
  // persistent workspace for the whole-path drivers
  void* ws = nullptr;
  size_t ws_bytes = 0;
  void* pinned = nullptr;
  size_t pinned_bytes = 0;
  cudaStream_t copy_stream = nullptr;
  cudaStream_t out_stream = nullptr;
  cudaEvent_t ev[8] = {};
  // optional per-kernel timing (CUDA events around the four big kernels on the context's stream)
  int profile = 0;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof_ev[4];
  size_t prof_used[4] = {0, 0, 0, 0};
};

namespace rrnn {
enum { K_GAS_LW = 0, K_LW_SOLVER = 1, K_GAS_SW = 2, K_SW_SOLVER = 3 };
// record a start event (returns slot index or -1) / the matching stop event
inline int prof_begin(rrnn_ctx* c, int kind) {
  if (!c->profile) return -1;
  if (c->prof_used[kind] == c->prof_ev[kind].size()) {
    if (c->prof_ev[kind].size() >= 8192) return -1;
    cudaEvent_t a, b;
    if (cudaEventCreate(&a) != cudaSuccess || cudaEventCreate(&b) != cudaSuccess) return -1;
    c->prof_ev[kind].push_back({a, b});
  }
  const int slot = (int)c->prof_used[kind]++;
  cudaEventRecord(c->prof_ev[kind][slot].first, c->stream);
  return slot;
}
inline void prof_end(rrnn_ctx* c, int kind, int slot) {
  if (slot >= 0) cudaEventRecord(c->prof_ev[kind][slot].second, c->stream);
}
}  // namespace rrnn

namespace rrnn { unsigned long long next_uid(); }

struct rrnn_model {
  unsigned long long uid = rrnn::next_uid();  // never reused: keys caches safely across destroy / create
  int nlayers = 0;
  int dims[rrnn::MAX_LAYERS + 1] = {};
  int act[rrnn::MAX_LAYERS] = {};
  std::vector<float> wpack, bpack, xmin, xmax, ymean, ystd;  // host copies
  std::vector<std::string> input_names;
  int device = 0;
  // device copies
  float* d_wpack = nullptr;
  float* d_bpack = nullptr;
  float* d_ymean = nullptr;
  float* d_ystd = nullptr;
  size_t w_off[rrnn::MAX_LAYERS] = {};
  size_t b_off[rrnn::MAX_LAYERS] = {};
};

struct rrnn_kdist {
  unsigned long long uid = rrnn::next_uid();
  int nbnd = 0, ngpt = 0, ntemp = 0;
  float temp_ref_min = 0.f, totplnk_delta = 1.f;
  std::vector<int> band_lims_gpt;  // [nbnd][2], 1-based inclusive
  std::vector<int> gpt2band;       // [ngpt], 0-based band
  std::vector<float> totplnk, solar_source;
  std::vector<float> solar_quiet, solar_facular, solar_sunspot;  // set_solar_variability tables (optional)
  std::vector<float> optimal_angle_fit;                          // (2,nbnd) (optional)
  int device = 0;
  int* d_band_lims_gpt = nullptr;
  int* d_gpt2band = nullptr;
  float* d_totplnk = nullptr;
  float* d_solar_source = nullptr;
  float* d_optimal_angle_fit = nullptr;
};

struct rrnn_cloud_lut {
  int nbnd = 0, nsize_liq = 0, nsize_ice = 0;
  float radliq_lwr = 0, radice_lwr = 0, liq_step = 0, ice_step = 0;
  float* d_tables = nullptr;  // extliq, ssaliq, asyliq, extice, ssaice, asyice back to back
  size_t off[6] = {};
  int is_pade = 0;            // 1: the tables are Pade coefficients (ncoeff, 3, nbnd) and sizreg holds the 6 x 4 size-regime bounds
  float sizreg[24] = {};
};

// ---- device helpers ------------------------------------------------------------------------------
namespace rrnn {

__device__ __forceinline__ float warp_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 16);
  v += __shfl_xor_sync(0xffffffffu, v, 8);
  v += __shfl_xor_sync(0xffffffffu, v, 4);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v;
}

// streaming (read-once) global load: keep it out of L1
__device__ __forceinline__ float ld_stream(const float* p) {
  float v;
  asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ float4 ld_stream4(const float4* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p));
  return v;
}
// streaming (write-once) global store
__device__ __forceinline__ void st_stream4(float4* p, float4 v) {
  asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void st_stream(float* p, float v) {
  asm volatile("st.global.cs.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}

template <bool FAST>
__device__ __forceinline__ float exp_neg(float x) {  // exp(x), x <= 0 in practice
  if (FAST) return __expf(x);
  return expf(x);
}
// 1 - exp(-x) and exp(-x) for x >= 0 without the cancellation of the literal "1 - exp(-x)" at small x.
// The reference forms 1 - trans in working precision (rte/kernels/mo_rte_solver_kernels.F90:757-773); with
// CUDA's 1-2 ulp expf that literal form is 2-4x noisier than the reference built on a correctly rounded libm,
// which shows up in heating rates of thin layers.  Branch-free: degree-7 Taylor polynomial below 0.35,
// the literal form above (where it is harmless).
__device__ __forceinline__ void exp_and_complement(float x, float& t, float& omt) {
  const float y = -x;
  float p = fmaf(y, 1.0f / 5040.0f, 1.0f / 720.0f);
  p = fmaf(p, y, 1.0f / 120.0f);
  p = fmaf(p, y, 1.0f / 24.0f);
  p = fmaf(p, y, 1.0f / 6.0f);
  p = fmaf(p, y, 0.5f);
  p = fmaf(p, y, 1.0f);
  const float em1 = p * y;  // expm1(-x) for small x
  const float e = expf(y);
  const bool small = x < 0.35f;
  omt = small ? -em1 : 1.0f - e;
  t = small ? 1.0f + em1 : e;
}

// Division / reciprocal / square root.  FAST = raw MUFU approximations (1-2 ulp, denormals flushed).
// Otherwise: MUFU seed + one Newton-Raphson step in FMA arithmetic -- branch-free, within ~1 ulp of the IEEE result
// (the IEEE-exact CUDA sequences carry slow-path branches that split the unrolled layer groups into small basic
// blocks and stop the scheduler from interleaving independent layers).
__device__ __forceinline__ float rcp_approx(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
template <bool FAST>
__device__ __forceinline__ float rcp(float x) {
  const float r = rcp_approx(x);
  if (FAST) return r;
  return fmaf(fmaf(-x, r, 1.0f), r, r);
}
template <bool FAST>
__device__ __forceinline__ float fdiv(float a, float b) {
  if (FAST) return __fdividef(a, b);
  const float r = rcp<false>(b);
  const float q = a * r;
  return fmaf(fmaf(-b, q, a), r, q);
}
template <bool FAST>
__device__ __forceinline__ float fsqrt(float x) {  // x > 0
  if (FAST) {
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
  }
  float y;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  const float sq = x * y;
  return fmaf(fmaf(-sq, sq, x), 0.5f * y, sq);
}

}  // namespace rrnn
