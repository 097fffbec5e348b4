// Helpers shared by the RTE solver kernels (rte_solvers.cu: one g-point per lane; rte_solvers_v4.cu: packed, two
// g-points per lane): deterministic cluster / DSMEM combination of per-chunk partial fluxes, the butterfly multi-value
// warp reduction, L2 cache policies for the streamed inputs and the L2-resident reverse-sweep buffer.
#pragma once
#include "common.cuh"
#include <cooperative_groups.h>

namespace rrnn {

namespace cg = cooperative_groups;

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

constexpr float kPi = 3.14159265358979323846f;

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

// L2 cache policies: the optical-property arrays are read exactly once (evict first); the reverse-sweep buffer,
// when it lives in global memory, is written and read back in LIFO order within microseconds (evict last: it should
// never reach HBM).
__device__ __forceinline__ uint64_t policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ float ld_once(const float* p, uint64_t pol) {
  float v;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v) : "l"(p), "l"(pol));
  return v;
}
// reverse-sweep buffer accessors: shared memory or (GBUF) L2-resident global scratch
template <bool GBUF>
__device__ __forceinline__ void buf_st2(float2* p, float2 v, uint64_t pol) {
  if (GBUF) asm volatile("st.global.L1::no_allocate.L2::cache_hint.v2.f32 [%0], {%1,%2}, %3;" ::"l"(p), "f"(v.x), "f"(v.y), "l"(pol) : "memory");
  else *p = v;
}
template <bool GBUF>
__device__ __forceinline__ float2 buf_ld2(const float2* p, uint64_t pol) {
  if (GBUF) {
    float2 v;
    asm volatile("ld.global.L1::no_allocate.L2::cache_hint.v2.f32 {%0,%1}, [%2], %3;" : "=f"(v.x), "=f"(v.y) : "l"(p), "l"(pol) : "memory");
    return v;
  }
  return *p;
}
template <bool GBUF>
__device__ __forceinline__ void buf_st1(float* p, float v, uint64_t pol) {
  if (GBUF) asm volatile("st.global.L1::no_allocate.L2::cache_hint.f32 [%0], %1, %2;" ::"l"(p), "f"(v), "l"(pol) : "memory");
  else *p = v;
}
template <bool GBUF>
__device__ __forceinline__ float buf_ld1(const float* p, uint64_t pol) {
  if (GBUF) {
    float v;
    asm volatile("ld.global.L1::no_allocate.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v) : "l"(p), "l"(pol) : "memory");
    return v;
  }
  return *p;
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
  float* flux_up;           // (nlay+1,ncol)
  float* flux_dn;
  float* scratch;           // GBUF: nCTA * 2 * L * 32 floats
  // compact sources (lw_solver_v5 only; null otherwise): lay_source then holds the Planck fraction, lev_source is null
  const float* planck_lay = nullptr;  // (16,nlay,ncol)   band Planck function at T_lay
  const float* planck_lev = nullptr;  // (16,nlay+1,ncol) ... at T_lev
  const int* gpt2band = nullptr;      // (ngpt) 0-based band of each g-point
  int pairs_in_band = 0;              // every pair of g-points (2i, 2i+1) lies in ONE band (required by the table look-ups below)
  // clouds whose increment has not been applied to tau (packed kernel only; null otherwise): tau += cld_tau(band(g)) happens in
  // the solver's registers (inc_1scalar_by_1scalar_bybnd, rte/kernels/mo_optical_props_kernels.F90:358-378); needs gpt2band
  const float* cld_tau = nullptr;     // (16,nlay,ncol) by-band cloud optical depth, rows padded to 16 bands
  // by-band fluxes (nbnd,nlay+1,ncol) straight from the packed kernel (null otherwise): the per-level sums stop at a band (every
  // band = 16 consecutive g-points starting at a multiple of 16; checked by the caller) -- ty_fluxes_byband without g-point fluxes
  float* bnd_up = nullptr;
  float* bnd_dn = nullptr;
  int nbnd = 0;
};


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
  float* scratch;  // GBUF: nCTA * 3 * L * 32 floats
  // clouds whose increment has not been applied (packed kernel only; null otherwise; requires g == null, i.e. gas g = 0):
  // inc_2stream_by_2stream_bybnd (rte/kernels/mo_optical_props_kernels.F90:453-485) happens in the solver's registers from the
  // by-band products t2 = tau_c | s2 = tau_c ssa_c | sg2 = tau_c ssa_c g_c, three 16-band segments per (layer, column) row
  const float* cld = nullptr;         // (48,nlay,ncol)
  const int* gpt2band = nullptr;      // (ngpt) 0-based band of each g-point
  int pairs_in_band = 0;              // every pair of g-points (2i, 2i+1) lies in ONE band (required with cld)
  // by-band fluxes (nbnd,nlay+1,ncol) straight from the packed kernel (null otherwise; see LwParams)
  float* bnd_up = nullptr;
  float* bnd_dn = nullptr;
  float* bnd_dir = nullptr;
  int nbnd = 0;
};


}  // namespace rrnn

// every pair of g-points (2i, 2i+1) in one band?  (all of RRTMGP's k-distributions: bands of 16.)  The packed solvers carry a pair per
// lane and look a band value up ONCE per pair.  Returns 2 when every aligned group of FOUR g-points lies in one band as well (the
// wide kernels carry four per lane), 1 for pairs only, 0 otherwise.
inline int kd_pairs_in_band(const rrnn_kdist_t* kd) {
  if (kd->ngpt & 1) return 0;
  for (int g = 0; g + 1 < kd->ngpt; g += 2)
    if (kd->gpt2band[g] != kd->gpt2band[g + 1]) return 0;
  if (kd->ngpt & 3) return 1;
  for (int g = 0; g + 3 < kd->ngpt; g += 4)
    if (kd->gpt2band[g] != kd->gpt2band[g + 3]) return 1;
  return 2;
}

// ---- host-side launch helpers
inline int ensure_scratch(rrnn_ctx_t* ctx, size_t bytes) {
  if (ctx->scratch_bytes >= bytes) return 0;
  if (ctx->scratch) { RRNN_CUDA(cudaStreamSynchronize(ctx->stream)); RRNN_CUDA(cudaFree(ctx->scratch)); ctx->scratch = nullptr; ctx->scratch_bytes = 0; }
  RRNN_CUDA(cudaMalloc(&ctx->scratch, bytes));
  ctx->scratch_bytes = bytes;
  return 0;
}

// Persistent clustered launch: one 32-thread CTA per g-point chunk, the chunks of a column form one cluster, and as
// many clusters as can be co-resident loop over the columns.  Returns the grid size through ncta_out.
template <typename P>
inline cudaError_t cluster_config(void (*kernel)(const P), int cluster, size_t smem, int ncol, cudaStream_t stream,
                                  cudaLaunchConfig_t& cfg, cudaLaunchAttribute* attr, int& ncta_out) {
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  cfg = cudaLaunchConfig_t{};
  cfg.gridDim = dim3((unsigned)cluster);
  cfg.blockDim = dim3(32);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)cluster;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int nclusters = 0;
  e = cudaOccupancyMaxActiveClusters(&nclusters, kernel, &cfg);
  if (e != cudaSuccess) return e;
  if (nclusters < 1) nclusters = 1;
  if (nclusters > ncol) nclusters = ncol;
  ncta_out = nclusters * cluster;
  cfg.gridDim = dim3((unsigned)ncta_out);
  return cudaSuccess;
}


