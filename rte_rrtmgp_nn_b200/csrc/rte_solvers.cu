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
#include "solver_common.cuh"

namespace rrnn {

constexpr int kLwU = 8;  // layers per software-pipelined group (LW)
constexpr int kSwU = 4;  // layers per group (SW)

// Work distribution.  CLUSTER: persistent clusters, cluster c handles columns c, c+nclusters, ...; the CTA's rank in
// the cluster is its g-point chunk.  Otherwise one (column, chunk) item per warp, no loop.
template <bool CLUSTER>
struct ItemLoop {
  int col, chunk, ncol, step;
  __device__ ItemLoop(int ncol_, int nchunks) : ncol(ncol_) {
    if (CLUSTER) {
      cg::cluster_group cl = cg::this_cluster();
      const int csize = cl.num_blocks();
      col = blockIdx.x / csize;
      chunk = cl.block_rank();
      step = gridDim.x / csize;
    } else {
      const int wpb = blockDim.x >> 5;
      const long long item = (long long)blockIdx.x * wpb + (threadIdx.x >> 5);
      col = (item < (long long)ncol_ * nchunks) ? (int)(item / nchunks) : ncol_;
      chunk = (int)(item % nchunks);
      step = ncol_;  // single pass
    }
  }
  __device__ bool valid() const { return col < ncol; }
  __device__ void next() { col += step; }
};

template <bool FAST, bool CLUSTER, bool GBUF>
__global__ void __launch_bounds__(64) lw_solver_kernel(const LwParams p) {
  extern __shared__ float smem[];
  constexpr int U = kLwU;
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  const int G = p.ngpt, L = p.nlay;
  const uint64_t pol_in = policy_evict_first();
  const uint64_t pol_buf = policy_evict_last();

  // per-warp shared memory: [L][32] float2 (t, src_up) unless GBUF, then flux partials [2][L+1]
  const int buf_floats = GBUF ? 0 : L * 64;
  const int per_warp = buf_floats + 2 * (L + 1);
  float* wbase = smem + (size_t)wib * per_warp;
  float2* buf = GBUF ? reinterpret_cast<float2*>(p.scratch) + (size_t)blockIdx.x * L * 32 : reinterpret_cast<float2*>(wbase);
  float* fup = wbase + buf_floats;
  float* fdn = fup + (L + 1);

  const float tau_thresh = 3.4526698e-4f;  // sqrt(epsilon(1._sp)), mo_rte_solver_kernels.F90:754
  // Sweep order i = 0..L-1 runs from the top of the atmosphere down: layer l(i) = l0 + dl*i.  In sweep order
  // layer i is bounded by level rows ent(i) (towards the top) and ext(i) = ent(i+1) (towards the surface).
  const int top = p.top_at_1;
  const int l0 = top ? 0 : L - 1;
  const int dl = top ? 1 : -1;
  const int sG = dl * G;  // element stride between consecutive layers in sweep order
  // lw_source_noscat (:770-773) takes source_dn from lev(l+1) and source_up from lev(l) whatever the orientation
  // (quirk Q1).  In sweep terms: top_at_1 -> dn uses ext, up uses ent (physical).  Otherwise the reference uses
  // dn <- lev(l+1) = ent, up <- lev(l) = ext; the physical choice is again dn <- ext, up <- ent.
  const bool dn_uses_ext = top || !p.bug_compat;

  for (ItemLoop<CLUSTER> it(p.ncol, p.nchunks); it.valid(); it.next()) {
    const int col = it.col, chunk = it.chunk;
    const int g = chunk * 32 + lane;
    const bool act = g < G;
    for (int i = lane; i < 2 * (L + 1); i += 32) fup[i] = 0.0f;
    __syncwarp();
    // inactive lanes (ngpt not a multiple of 32) shadow the chunk's first g-point and contribute zero
    const int gs = act ? g : chunk * 32;
    const float* tau = p.tau + (size_t)col * L * G + (size_t)l0 * G + gs;          // layer i at tau[i*sG]
    const float* lay = p.lay_source + (size_t)col * L * G + (size_t)l0 * G + gs;
    const float* lev = p.lev_source + (size_t)col * (L + 1) * G + gs;
    const float* lext = lev + (size_t)(top ? 1 : L - 1) * G;                       // ext(i) at lext[i*sG]
    const size_t gc_off = (size_t)col * G + gs;
    const float emis = p.sfc_emis[gc_off];
    const float ssrc = p.sfc_source[gc_off];
    const float inc = p.inc_flux ? p.inc_flux[gc_off] : 0.0f;
    const float live = act ? 1.0f : 0.0f;

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
      float carry = ld_once(lev + (size_t)(top ? 0 : L) * G, pol_in);  // ent(0)
      auto load_group = [&](int i0) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int o = min(i0 + u, L - 1) * sG;
          n_tau[u] = ld_once(tau + o, pol_in);
          n_lay[u] = ld_once(lay + o, pol_in);
          n_ext[u] = ld_once(lext + o, pol_in);
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
          if (tl > tau_thresh) fact = __fdividef(omt, tl) - t;
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
            buf_st2<GBUF>(buf + i * 32 + lane, make_float2(tv[u], sup[u]), pol_buf);
          }
          red[u] = fac * I;
        }
        multi_reduce<U>(red, lane);
        const int i = i0 + multi_index<U>(lane);
        if (multi_writer<U>(lane) && i < L) fdn[top ? i + 1 : L - 1 - i] += red[0];
      }
      // ---------------- surface ----------------
      float Uu = I * (1.0f - emis) + emis * ssrc;  // :269
      {
        const float s = warp_sum(fac * Uu);
        if (lane == 0) fup[top ? L : 0] += s;
      }
      __syncwarp();
      // ---------------- upward sweep (reverse order) from the buffer, software pipelined ----------------
      float2 nb[U];
      auto load_back = [&](int i1) {
#pragma unroll
        for (int u = 0; u < U; ++u) nb[u] = buf_ld2<GBUF>(buf + max(i1 - u, 0) * 32 + lane, pol_buf);
      };
      load_back(L - 1);
      for (int i1 = L - 1; i1 >= 0; i1 -= U) {
        float2 b[U];
#pragma unroll
        for (int u = 0; u < U; ++u) b[u] = nb[u];
        if (i1 - U >= 0) load_back(i1 - U);
        float red[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          if (i1 - u >= 0) Uu = b[u].x * Uu + b[u].y;
          red[u] = fac * Uu;
        }
        multi_reduce<U>(red, lane);
        const int i = i1 - multi_index<U>(lane);
        if (multi_writer<U>(lane) && i >= 0) fup[top ? i : L - i] += red[0];
      }
      __syncwarp();
    }
    // combine the g-chunks of this column
    float* const gout[2] = {p.flux_up + (size_t)col * (L + 1), p.flux_dn + (size_t)col * (L + 1)};
    combine_chunks<CLUSTER, 2>(fup, L, lane, gout);
  }
}

// ---------------------------------------------------------------------------------------------------
template <bool FAST, bool HAS_G, bool CLUSTER, bool GBUF>
__global__ void __launch_bounds__(64) sw_solver_kernel(const SwParams p) {
  extern __shared__ float smem[];
  constexpr int U = kSwU;
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  const int G = p.ngpt, L = p.nlay;
  const uint64_t pol_in = policy_evict_first();
  const uint64_t pol_buf = policy_evict_last();

  // per-warp: e[L][32], f[L][32], alpha_below[L][32] (shared memory unless GBUF), then flux partials [3][L+1]
  const int buf_floats = GBUF ? 0 : L * 96;
  const int per_warp = buf_floats + 3 * (L + 1);
  float* wbase = smem + (size_t)wib * per_warp;
  float* be = GBUF ? p.scratch + (size_t)blockIdx.x * L * 96 : wbase;
  float* bf = be + L * 32;
  float* ba = bf + L * 32;
  float* fup = wbase + buf_floats;
  float* fdn = fup + (L + 1);
  float* fdr = fdn + (L + 1);

  const float k_min = 1.e-4f;       // mo_rte_solver_kernels.F90:76-82 (single precision)
  const float eps = 1.1920929e-7f;  // epsilon(1._sp)
  const int top = p.top_at_1;
  const int l0 = top ? 0 : L - 1;
  const int sG = (top ? 1 : -1) * G;
  const int top_level = top ? 0 : L;

  for (ItemLoop<CLUSTER> it(p.ncol, p.nchunks); it.valid(); it.next()) {
    const int col = it.col, chunk = it.chunk;
    const int gp = chunk * 32 + lane;
    const bool act = gp < G;
    for (int i = lane; i < 3 * (L + 1); i += 32) fup[i] = 0.0f;
    __syncwarp();
    const int gs = act ? gp : chunk * 32;  // inactive lanes shadow the chunk's first g-point and contribute zero
    const float live = act ? 1.0f : 0.0f;
    const float* tau = p.tau + (size_t)col * L * G + (size_t)l0 * G + gs;  // layer i (sweep order) at [i*sG]
    const float* ssa = p.ssa + (size_t)col * L * G + (size_t)l0 * G + gs;
    const float* gas = HAS_G ? p.g + (size_t)col * L * G + (size_t)l0 * G + gs : nullptr;
    const size_t gc_off = (size_t)col * G + gs;
    const float mu0 = p.mu0[col];
    const float mu0_inv = 1.0f / mu0;

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
        const int o = min(i0 + u, L - 1) * sG;
        n_t[u] = ld_once(tau + o, pol_in);
        n_w[u] = ld_once(ssa + o, pol_in);
        n_g[u] = HAS_G ? ld_once(gas + o, pol_in) : 0.0f;
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
          const float s_up = Rdir[u] * dir;
          const float s_dn = Tdir[u] * dir;
          dir = Tnos[u] * dir;
          const float d = rcp<FAST>(1.0f - Rdif[u] * alpha);
          const float e = d * Tdif[u];
          const float f = d * (Rdif[u] * beta + s_up);
          beta = s_dn + e * (beta + alpha * s_up);
          alpha = Rdif[u] + Tdif[u] * e * alpha;
          buf_st1<GBUF>(be + i * 32 + lane, e, pol_buf);
          buf_st1<GBUF>(bf + i * 32 + lane, f, pol_buf);
          buf_st1<GBUF>(ba + i * 32 + lane, alpha, pol_buf);  // reflectance seen from the level BELOW layer i
        }
        red[u] = dir;
        red[U + u] = beta + dir;
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
    const float a_s = p.alb_dif[gc_off];
    const float S_s = dir * p.alb_dir[gc_off];  // source_sfc :1477
    float Uu = fdiv<FAST>(a_s * beta + S_s, 1.0f - a_s * alpha) * live;
    {
      const int sfc = top ? L : 0;
      const float su = warp_sum(Uu), sa = warp_sum(alpha * Uu);
      if (lane == 0) { fup[sfc] += su; fdn[sfc] += sa; }
    }
    __syncwarp();
    // ---------------- sweep 2: surface -> top (back substitution), software pipelined ----------------
    float ne[U], nf[U], na[U];
    auto load_back = [&](int i1) {
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int i = max(i1 - u, 0);
        ne[u] = buf_ld1<GBUF>(be + i * 32 + lane, pol_buf);
        nf[u] = buf_ld1<GBUF>(bf + i * 32 + lane, pol_buf);
        // reflectance of the atmosphere above the level at the top of layer i (0 at the top of the domain)
        na[u] = (i > 0) ? buf_ld1<GBUF>(ba + (i - 1) * 32 + lane, pol_buf) : 0.0f;
      }
    };
    load_back(L - 1);
    for (int i1 = L - 1; i1 >= 0; i1 -= U) {
      float ce[U], cf[U], ca[U];
#pragma unroll
      for (int u = 0; u < U; ++u) { ce[u] = ne[u]; cf[u] = nf[u]; ca[u] = na[u]; }
      if (i1 - U >= 0) load_back(i1 - U);
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
          const int lvl = top ? i : L - i;  // level at the top of layer i
          if (idx < U) fup[lvl] += red[0]; else fdn[lvl] += red[0];
        }
      }
    }
    __syncwarp();
    float* const gout[3] = {p.flux_up + (size_t)col * (L + 1), p.flux_dn + (size_t)col * (L + 1),
                            p.flux_dir + (size_t)col * (L + 1)};
    combine_chunks<CLUSTER, 3>(fup, L, lane, gout);
  }
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

namespace rrnn {
int launch_lw_v6(rrnn_ctx_t* ctx, LwParams& p);           // rte_solvers_tma.cu (TMA-staged, packed); -1 = shape not supported
int launch_sw_v6(rrnn_ctx_t* ctx, SwParams& p, bool fast);
}

static int pick_warps_per_block(size_t per_warp_bytes) {
  // two warps per CTA unless that does not fit
  return (2 * per_warp_bytes <= 200 * 1024) ? 2 : 1;
}

// Where does the reverse-sweep buffer live?  Shared memory when enough warps fit per SM to hide latency, otherwise
// an L2-resident global scratch ring (ctx flag solver_buffer: 0 auto, 1 shared memory, 2 global).
static bool use_global_buffer(const rrnn_ctx_t* ctx, size_t smem_per_warp) {
  if (ctx->solver_buffer == 1) return false;
  if (ctx->solver_buffer == 2) return true;
  const size_t warps_by_smem = (size_t)220 * 1024 / smem_per_warp;
  return warps_by_smem < 12;
}

extern "C" int rrnn_lw_solver_noscat(rrnn_ctx_t* ctx, int ngpt, int nlay, int ncol, int top_at_1, int nmus, const float* Ds,
                                     const float* weights, const float* inc_flux_d, const float* tau_d,
                                     const float* lay_source_d, const float* lev_source_d, const float* sfc_emis_gpt_d,
                                     const float* sfc_source_d, float* flux_up_d, float* flux_dn_d) {
  rrnn::NvtxRange nvtx_("lw_solver_noscat");
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
  const size_t part = 2 * (size_t)(nlay + 1) * sizeof(float);
  const size_t bufb = (size_t)nlay * 64 * sizeof(float);
  const bool clustered = p.nchunks <= 8;
  const bool gbuf = clustered && use_global_buffer(ctx, bufb + part);
  const size_t per_warp = part + (gbuf ? 0 : bufb);
  RRNN_CHECK(per_warp <= ctx->smem_optin, "rrnn_lw_solver_noscat: nlay too large for the on-chip layer buffer");
  const int ps = prof_begin(ctx, K_LW_SOLVER);
  int rc4 = -1;
  if (ctx->solver_variant == 0) {
    rc4 = launch_lw_v6(ctx, p);
    if (rc4 > 0) { prof_end(ctx, K_LW_SOLVER, ps); return rc4; }
  }
  if (rc4 == 0) {
    // done by the packed kernel
  } else if (clustered) {
    cudaLaunchConfig_t cfg; cudaLaunchAttribute attr[1]; int ncta = 0;
#define LW_CL(F, GB)                                                                                              \
    do {                                                                                                          \
      RRNN_CUDA(cluster_config(lw_solver_kernel<F, true, GB>, p.nchunks, per_warp, ncol, ctx->stream, cfg, attr, ncta)); \
      if (GB) { if (int rc = ensure_scratch(ctx, (size_t)ncta * bufb)) return rc; p.scratch = (float*)ctx->scratch; } \
      RRNN_CUDA(cudaLaunchKernelEx(&cfg, lw_solver_kernel<F, true, GB>, p));                                      \
    } while (0)
    if (ctx->fast_math) { if (gbuf) LW_CL(true, true); else LW_CL(true, false); }
    else { if (gbuf) LW_CL(false, true); else LW_CL(false, false); }
#undef LW_CL
  } else {
    const int wpb = pick_warps_per_block(per_warp);
    const size_t smem = per_warp * wpb;
    const long long blocks = ((long long)ncol * p.nchunks + wpb - 1) / wpb;
    RRNN_CHECK(blocks < 2147483647LL, "rrnn_lw_solver_noscat: too many columns for one launch");
    const size_t nflux = (size_t)ncol * (nlay + 1) * sizeof(float);
    RRNN_CUDA(cudaMemsetAsync(flux_up_d, 0, nflux, ctx->stream));
    RRNN_CUDA(cudaMemsetAsync(flux_dn_d, 0, nflux, ctx->stream));
    if (ctx->fast_math) {
      RRNN_CUDA(cudaFuncSetAttribute(lw_solver_kernel<true, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      lw_solver_kernel<true, false, false><<<(unsigned)blocks, wpb * 32, smem, ctx->stream>>>(p);
    } else {
      RRNN_CUDA(cudaFuncSetAttribute(lw_solver_kernel<false, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      lw_solver_kernel<false, false, false><<<(unsigned)blocks, wpb * 32, smem, ctx->stream>>>(p);
    }
  }
  prof_end(ctx, K_LW_SOLVER, ps);
  RRNN_LAUNCH_CHECK(ctx);
  return 0;
}

namespace rrnn {
// rrnn_lw_solver_noscat on compact sources (pipeline.cu, rrnn_lw_fluxes): lay_source = pfrac * planck_lay and
// lev_source = pfrac * planck_lev are formed inside lw_solver_v5 (see there).  Only the default solver variant has it.
int lw_solver_noscat_compact(rrnn_ctx_t* ctx, int ngpt, int nlay, int ncol, int top_at_1, int nmus, const float* Ds, const float* weights,
                             const float* tau_d, const float* pfrac_d, const float* planck_lay_d, const float* planck_lev_d,
                             const int* gpt2band_d, int pairs_in_band, const float* sfc_emis_gpt_d, const float* sfc_source_d, float* flux_up_d,
                             float* flux_dn_d) {
  if (ncol == 0) return 0;
  RRNN_CUDA(cudaSetDevice(ctx->device));
  LwParams p{};
  p.ngpt = ngpt; p.nlay = nlay; p.ncol = ncol; p.top_at_1 = top_at_1 ? 1 : 0; p.nmus = nmus;
  p.bug_compat = ctx->lw_source_bug_compat;
  p.nchunks = (ngpt + 31) / 32;
  for (int i = 0; i < nmus; ++i) { p.Ds[i] = Ds[i]; p.wts[i] = weights[i]; }
  p.tau = tau_d; p.lay_source = pfrac_d; p.planck_lay = planck_lay_d; p.planck_lev = planck_lev_d; p.gpt2band = gpt2band_d;
  p.pairs_in_band = pairs_in_band;
  p.sfc_emis = sfc_emis_gpt_d; p.sfc_source = sfc_source_d; p.flux_up = flux_up_d; p.flux_dn = flux_dn_d;
  const int ps = prof_begin(ctx, K_LW_SOLVER);
  const int rc = launch_lw_v6(ctx, p);
  prof_end(ctx, K_LW_SOLVER, ps);  // (paired with prof_begin on every path)
  if (rc < 0) return fail("lw_solver (compact sources): shape not supported by the packed kernel");
  if (rc > 0) return rc;
  RRNN_LAUNCH_CHECK(ctx);
  return 0;
}
}  // namespace rrnn

extern "C" int rrnn_lw_solver_noscat_compact(rrnn_ctx_t* ctx, const rrnn_kdist_t* kd, int nlay, int ncol, int top_at_1, int nmus,
                                             const float* Ds, const float* weights, const float* tau_d, const float* pfrac_d,
                                             const float* planck_lay_d, const float* planck_lev_d, const float* sfc_emis_gpt_d,
                                             const float* sfc_source_d, float* flux_up_d, float* flux_dn_d) {
  rrnn::NvtxRange nvtx_("lw_solver_noscat");
  RRNN_CHECK(ctx && kd, "rrnn_lw_solver_noscat_compact: null handle");
  RRNN_CHECK(nlay > 0 && ncol >= 0, "rrnn_lw_solver_noscat_compact: bad extents");
  RRNN_CHECK(nmus >= 1 && nmus <= 4, "rte_lw: have to ask for between 1 and 4 quadrature points for no-scattering calculation");
  RRNN_CHECK(tau_d && pfrac_d && planck_lay_d && planck_lev_d && sfc_emis_gpt_d && sfc_source_d && flux_up_d && flux_dn_d,
             "rrnn_lw_solver_noscat_compact: null argument");
  return lw_solver_noscat_compact(ctx, kd->ngpt, nlay, ncol, top_at_1, nmus, Ds, weights, tau_d, pfrac_d, planck_lay_d, planck_lev_d,
                                  kd->d_gpt2band, kd_pairs_in_band(kd), sfc_emis_gpt_d, sfc_source_d, flux_up_d, flux_dn_d);
}

extern "C" int rrnn_sw_solver_2stream(rrnn_ctx_t* ctx, int ngpt, int nlay, int ncol, int top_at_1, const float* inc_flux_d,
                                      const float* inc_flux_dif_d, const float* tau_d, const float* ssa_d, const float* g_d,
                                      const float* mu0_d, const float* sfc_alb_dir_d, const float* sfc_alb_dif_d,
                                      float* flux_up_d, float* flux_dn_d, float* flux_dir_d) {
  rrnn::NvtxRange nvtx_("sw_two_stream_source + adding");
  RRNN_CHECK(ctx, "rrnn_sw_solver_2stream: null context");
  RRNN_CHECK(ngpt > 0 && nlay > 0 && ncol >= 0, "rrnn_sw_solver_2stream: bad extents");
  if (ncol == 0) return 0;
  RRNN_CUDA(cudaSetDevice(ctx->device));
  SwParams p{};
  p.ngpt = ngpt; p.nlay = nlay; p.ncol = ncol; p.top_at_1 = top_at_1 ? 1 : 0;
  p.nchunks = (ngpt + 31) / 32;
  p.inc_flux = inc_flux_d; p.inc_flux_dif = inc_flux_dif_d; p.tau = tau_d; p.ssa = ssa_d; p.g = g_d; p.mu0 = mu0_d;
  p.alb_dir = sfc_alb_dir_d; p.alb_dif = sfc_alb_dif_d; p.flux_up = flux_up_d; p.flux_dn = flux_dn_d; p.flux_dir = flux_dir_d;
  const size_t part = 3 * (size_t)(nlay + 1) * sizeof(float);
  const size_t bufb = (size_t)nlay * 96 * sizeof(float);
  const bool clustered = p.nchunks <= 8;
  const bool gbuf = clustered && use_global_buffer(ctx, bufb + part);
  const size_t per_warp = part + (gbuf ? 0 : bufb);
  RRNN_CHECK(per_warp <= ctx->smem_optin, "rrnn_sw_solver_2stream: nlay too large for the on-chip layer buffer");
  const bool fast = ctx->fast_math || ctx->sw_fast_math;
  const int ps = prof_begin(ctx, K_SW_SOLVER);
  int rc4 = -1;
  if (ctx->solver_variant == 0) {
    rc4 = launch_sw_v6(ctx, p, fast);
    if (rc4 > 0) { prof_end(ctx, K_SW_SOLVER, ps); return rc4; }
  }
  if (rc4 == 0) {
    // done by the packed kernel
  } else if (clustered) {
    cudaLaunchConfig_t cfg; cudaLaunchAttribute attr[1]; int ncta = 0;
#define SW_CL(F, HG, GB)                                                                                              \
    do {                                                                                                              \
      RRNN_CUDA(cluster_config(sw_solver_kernel<F, HG, true, GB>, p.nchunks, per_warp, ncol, ctx->stream, cfg, attr, ncta)); \
      if (GB) { if (int rc = ensure_scratch(ctx, (size_t)ncta * bufb)) return rc; p.scratch = (float*)ctx->scratch; } \
      RRNN_CUDA(cudaLaunchKernelEx(&cfg, sw_solver_kernel<F, HG, true, GB>, p));                                      \
    } while (0)
#define SW_CL2(F, HG) do { if (gbuf) SW_CL(F, HG, true); else SW_CL(F, HG, false); } while (0)
    if (fast) { if (g_d) SW_CL2(true, true); else SW_CL2(true, false); }
    else { if (g_d) SW_CL2(false, true); else SW_CL2(false, false); }
#undef SW_CL2
#undef SW_CL
  } else {
    const int wpb = pick_warps_per_block(per_warp);
    const size_t smem = per_warp * wpb;
    const long long blocks = ((long long)ncol * p.nchunks + wpb - 1) / wpb;
    RRNN_CHECK(blocks < 2147483647LL, "rrnn_sw_solver_2stream: too many columns for one launch");
    const size_t nflux = (size_t)ncol * (nlay + 1) * sizeof(float);
    RRNN_CUDA(cudaMemsetAsync(flux_up_d, 0, nflux, ctx->stream));
    RRNN_CUDA(cudaMemsetAsync(flux_dn_d, 0, nflux, ctx->stream));
    RRNN_CUDA(cudaMemsetAsync(flux_dir_d, 0, nflux, ctx->stream));
#define SW_PL(F, HG)                                                                                                       \
    do {                                                                                                                   \
      RRNN_CUDA(cudaFuncSetAttribute(sw_solver_kernel<F, HG, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
      sw_solver_kernel<F, HG, false, false><<<(unsigned)blocks, wpb * 32, smem, ctx->stream>>>(p);                         \
    } while (0)
    if (fast) { if (g_d) SW_PL(true, true); else SW_PL(true, false); }
    else { if (g_d) SW_PL(false, true); else SW_PL(false, false); }
#undef SW_PL
  }
  prof_end(ctx, K_SW_SOLVER, ps);
  RRNN_LAUNCH_CHECK(ctx);
  return 0;
}

// rte_lw (rte/mo_rte_lw.F90:60-424) for ty_optical_props_1scl
extern "C" int rrnn_rte_lw(rrnn_ctx_t* ctx, const rrnn_kdist_t* kd, int nlay, int ncol, int top_at_1, int n_gauss_angles,
                           const float* inc_flux_d, const float* tau_d, const float* lay_source_d,
                           const float* lev_source_d, const float* sfc_source_d, const float* sfc_emis_d,
                           float* flux_up_d, float* flux_dn_d) {
  rrnn::NvtxRange nvtx_("rte_lw");
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
  rrnn::NvtxRange nvtx_("rte_sw");
  RRNN_CHECK(flux_up_d && flux_dn_d && flux_dn_dir_d, "rte_sw: no space allocated for fluxes");
  return rrnn_sw_solver_2stream(ctx, ngpt, nlay, ncol, top_at_1, inc_flux_d, inc_flux_dif_d, tau_d, ssa_d, g_d, mu0_d,
                                sfc_alb_dir_gpt_d, sfc_alb_dif_gpt_d, flux_up_d, flux_dn_d, flux_dn_dir_d);
}

// ---------------------------------------------------------------------------------------------------- clouds folded into the solvers
// `clouds%increment(atmos)` (and, for the shortwave, the delta-scaled by-band properties it adds) is not
// a pass over the (ngpt,nlay,ncol) arrays here.  A small kernel turns the by-band cloud properties (16 times fewer numbers) into
// padded table rows, and the packed solvers add them to the gas optical properties in registers (lw_solver_v6<.., CLD>,
// sw_solver_v6<.., GM = 2>): the all-sky path reads what the clear-sky path reads plus 64 / 192 bytes per layer and column.
namespace rrnn {

// LW: tau_c (nbnd,nlay,ncol) -> rows of 16 floats (bands beyond nbnd zero)
__global__ void cloud_rows_lw_kernel(size_t nsmp, int nbnd, const float* __restrict__ tau, float* __restrict__ rows) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nsmp * 16) return;
  const size_t s = i >> 4;
  const int b = (int)(i & 15);
  rows[i] = b < nbnd ? tau[s * nbnd + b] : 0.0f;
}
// SW: (tau_c, ssa_c, g_c) (nbnd,nlay,ncol) -> rows of 48 floats: t2 = tau_c | s2 = tau_c*ssa_c | sg2 = tau_c*ssa_c*g_c, the operands
// of inc_2stream_by_2stream_bybnd (rte/kernels/mo_optical_props_kernels.F90:470-478) that do not depend on the g-point
__global__ void cloud_rows_sw_kernel(size_t nsmp, int nbnd, const float* __restrict__ tau, const float* __restrict__ ssa,
                                     const float* __restrict__ g, float* __restrict__ rows) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nsmp * 16) return;
  const size_t s = i >> 4;
  const int b = (int)(i & 15);
  float t2 = 0.0f, s2 = 0.0f, sg2 = 0.0f;
  if (b < nbnd) {
    t2 = tau[s * nbnd + b];
    s2 = __fmul_rn(t2, ssa[s * nbnd + b]);
    sg2 = __fmul_rn(s2, g[s * nbnd + b]);
  }
  rows[s * 48 + b] = t2;
  rows[s * 48 + 16 + b] = s2;
  rows[s * 48 + 32 + b] = sg2;
}

int cloud_rows_lw(rrnn_ctx_t* ctx, size_t nsmp, int nbnd, const float* tau_bnd_d, float* rows_d) {
  RRNN_CHECK(nbnd >= 1 && nbnd <= 16, "clouds: the packed solvers take at most 16 bands");
  cloud_rows_lw_kernel<<<(unsigned)((nsmp * 16 + 255) / 256), 256, 0, ctx->stream>>>(nsmp, nbnd, tau_bnd_d, rows_d);
  RRNN_LAUNCH_CHECK(ctx);
  return 0;
}
int cloud_rows_sw(rrnn_ctx_t* ctx, size_t nsmp, int nbnd, const float* tau_bnd_d, const float* ssa_bnd_d, const float* g_bnd_d, float* rows_d) {
  RRNN_CHECK(nbnd >= 1 && nbnd <= 16, "clouds: the packed solvers take at most 16 bands");
  cloud_rows_sw_kernel<<<(unsigned)((nsmp * 16 + 255) / 256), 256, 0, ctx->stream>>>(nsmp, nbnd, tau_bnd_d, ssa_bnd_d, g_bnd_d, rows_d);
  RRNN_LAUNCH_CHECK(ctx);
  return 0;
}

// lw_solver_noscat on (tau + clouds), sources materialised (planck_* null) or factored; -1 (no message) = the packed kernel does
// not take this call
int lw_solver_clouds(rrnn_ctx_t* ctx, const rrnn_kdist_t* kd, int nlay, int ncol, int top_at_1, int nmus, const float* Ds, const float* weights,
                     const float* inc_flux_d, const float* tau_d, const float* lay_d, const float* lev_d, const float* planck_lay_d,
                     const float* planck_lev_d, const float* sfc_emis_gpt_d, const float* sfc_source_d, const float* cld_rows_d,
                     float* flux_up_d, float* flux_dn_d) {
  if (ncol == 0) return 0;
  if (ctx->solver_variant != 0) return -1;
  RRNN_CUDA(cudaSetDevice(ctx->device));
  LwParams p{};
  p.ngpt = kd->ngpt; p.nlay = nlay; p.ncol = ncol; p.top_at_1 = top_at_1 ? 1 : 0; p.nmus = nmus;
  p.bug_compat = ctx->lw_source_bug_compat;
  p.nchunks = (kd->ngpt + 31) / 32;
  for (int i = 0; i < nmus; ++i) { p.Ds[i] = Ds[i]; p.wts[i] = weights[i]; }
  p.inc_flux = inc_flux_d; p.tau = tau_d; p.lay_source = lay_d; p.lev_source = lev_d; p.planck_lay = planck_lay_d; p.planck_lev = planck_lev_d;
  p.gpt2band = kd->d_gpt2band; p.cld_tau = cld_rows_d; p.pairs_in_band = kd_pairs_in_band(kd);
  p.sfc_emis = sfc_emis_gpt_d; p.sfc_source = sfc_source_d; p.flux_up = flux_up_d; p.flux_dn = flux_dn_d;
  const int ps = prof_begin(ctx, K_LW_SOLVER);
  const int rc = launch_lw_v6(ctx, p);
  prof_end(ctx, K_LW_SOLVER, ps);
  if (rc != 0) return rc;
  RRNN_LAUNCH_CHECK(ctx);
  return 0;
}

int sw_solver_clouds(rrnn_ctx_t* ctx, const rrnn_kdist_t* kd, int nlay, int ncol, int top_at_1, const float* inc_flux_d, const float* inc_flux_dif_d,
                     const float* tau_d, const float* ssa_d, const float* cld_rows_d, const float* mu0_d, const float* alb_dir_d,
                     const float* alb_dif_d, float* flux_up_d, float* flux_dn_d, float* flux_dir_d) {
  if (ncol == 0) return 0;
  if (ctx->solver_variant != 0) return -1;
  RRNN_CUDA(cudaSetDevice(ctx->device));
  SwParams p{};
  p.ngpt = kd->ngpt; p.nlay = nlay; p.ncol = ncol; p.top_at_1 = top_at_1 ? 1 : 0;
  p.nchunks = (kd->ngpt + 31) / 32;
  p.inc_flux = inc_flux_d; p.inc_flux_dif = inc_flux_dif_d; p.tau = tau_d; p.ssa = ssa_d; p.g = nullptr; p.mu0 = mu0_d;
  p.cld = cld_rows_d; p.gpt2band = kd->d_gpt2band; p.pairs_in_band = kd_pairs_in_band(kd);
  p.alb_dir = alb_dir_d; p.alb_dif = alb_dif_d; p.flux_up = flux_up_d; p.flux_dn = flux_dn_d; p.flux_dir = flux_dir_d;
  const int ps = prof_begin(ctx, K_SW_SOLVER);
  const int rc = launch_sw_v6(ctx, p, ctx->fast_math || ctx->sw_fast_math);
  prof_end(ctx, K_SW_SOLVER, ps);
  if (rc != 0) return rc;
  RRNN_LAUNCH_CHECK(ctx);
  return 0;
}

}  // namespace rrnn

static const char* kCloudShapeMsg =
    ": this shape is not taken by the packed solver (ngpt a multiple of 4 and <= 512, nlay >= 8, 16-byte aligned arrays, default "
    "solver_variant / fast_math); apply increment() and call the plain entry point";

// rte_lw(atmos, ...) with `clouds%increment(atmos)` still pending: cld_tau_bnd_d is the by-band cloud optical depth (nbnd,nlay,ncol)
extern "C" int rrnn_rte_lw_clouds(rrnn_ctx_t* ctx, const rrnn_kdist_t* kd, int nlay, int ncol, int top_at_1, int n_gauss_angles,
                                  const float* inc_flux_d, const float* tau_d, const float* lay_source_d, const float* lev_source_d,
                                  const float* sfc_source_d, const float* sfc_emis_d, const float* cld_tau_bnd_d, float* flux_up_d,
                                  float* flux_dn_d) {
  rrnn::NvtxRange nvtx_("rte_lw");
  RRNN_CHECK(ctx && kd && tau_d && lay_source_d && lev_source_d && sfc_source_d && sfc_emis_d && cld_tau_bnd_d && flux_up_d && flux_dn_d,
             "rte_lw: null argument");
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
  const size_t n = (size_t)ngpt * ncol, nsmp = (size_t)ncol * nlay;
  float* tmp = nullptr;   // sfc_emis by g-point, then the cloud rows
  RRNN_CUDA(cudaMallocAsync((void**)&tmp, (((n + 63) & ~(size_t)63) + nsmp * 16) * sizeof(float), ctx->stream));
  float* rows = tmp + ((n + 63) & ~(size_t)63);
  expand_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(kd->nbnd, ngpt, ncol, kd->d_gpt2band, sfc_emis_d, tmp);
  int rc = cloud_rows_lw(ctx, nsmp, kd->nbnd, cld_tau_bnd_d, rows);
  if (rc == 0)
    rc = lw_solver_clouds(ctx, kd, nlay, ncol, top_at_1, n_gauss_angles, gauss_Ds[n_gauss_angles - 1], gauss_wts[n_gauss_angles - 1], inc_flux_d,
                          tau_d, lay_source_d, lev_source_d, nullptr, nullptr, tmp, sfc_source_d, rows, flux_up_d, flux_dn_d);
  cudaFreeAsync(tmp, ctx->stream);
  if (rc < 0) return fail(std::string("rte_lw (clouds)") + kCloudShapeMsg);
  return rc;
}

// rte_sw(atmos, ...) with `clouds%increment(atmos)` still pending on gas optical properties whose g is 0 (the NN gas optics):
// cld_*_bnd_d are the (delta-scaled, if the caller wants that) by-band cloud properties (nbnd,nlay,ncol)
extern "C" int rrnn_rte_sw_clouds(rrnn_ctx_t* ctx, const rrnn_kdist_t* kd, int nlay, int ncol, int top_at_1, const float* mu0_d,
                                  const float* inc_flux_d, const float* sfc_alb_dir_gpt_d, const float* sfc_alb_dif_gpt_d,
                                  const float* inc_flux_dif_d, const float* tau_d, const float* ssa_d, const float* cld_tau_bnd_d,
                                  const float* cld_ssa_bnd_d, const float* cld_g_bnd_d, float* flux_up_d, float* flux_dn_d,
                                  float* flux_dn_dir_d) {
  rrnn::NvtxRange nvtx_("rte_sw");
  RRNN_CHECK(ctx && kd && mu0_d && inc_flux_d && sfc_alb_dir_gpt_d && sfc_alb_dif_gpt_d && tau_d && ssa_d && cld_tau_bnd_d && cld_ssa_bnd_d &&
                 cld_g_bnd_d, "rte_sw: null argument");
  RRNN_CHECK(flux_up_d && flux_dn_d && flux_dn_dir_d, "rte_sw: no space allocated for fluxes");
  if (ncol == 0) return 0;
  RRNN_CUDA(cudaSetDevice(ctx->device));
  const size_t nsmp = (size_t)ncol * nlay;
  float* rows = nullptr;
  RRNN_CUDA(cudaMallocAsync((void**)&rows, nsmp * 48 * sizeof(float), ctx->stream));
  int rc = cloud_rows_sw(ctx, nsmp, kd->nbnd, cld_tau_bnd_d, cld_ssa_bnd_d, cld_g_bnd_d, rows);
  if (rc == 0)
    rc = sw_solver_clouds(ctx, kd, nlay, ncol, top_at_1, inc_flux_d, inc_flux_dif_d, tau_d, ssa_d, rows, mu0_d, sfc_alb_dir_gpt_d,
                          sfc_alb_dif_gpt_d, flux_up_d, flux_dn_d, flux_dn_dir_d);
  cudaFreeAsync(rows, ctx->stream);
  if (rc < 0) return fail(std::string("rte_sw (clouds)") + kCloudShapeMsg);
  return rc;
}

// ---------------------------------------------------------------------------------------------------- by-band fluxes from the packed solvers
// ty_fluxes_byband (extensions/mo_fluxes_byband.F90:41-131) without g-point fluxes: the packed kernels' per-level sums pass through
// sums over 8 lanes = 16 g-points on their way to the broadband sum, which IS a band sum when every band is 16 consecutive g-points
// starting at a multiple of 16 (all of RRTMGP's k-distributions: 16 x 16 longwave, 14 x 16 shortwave).  The by-band arrays
// (nbnd,nlay+1,ncol) then cost a few stores per group of 8 layers; the general kernels + rrnn_sum_byband remain the path for
// anything else.
static bool bands_of_16(const rrnn_kdist_t* kd) {
  if (kd->nbnd < 1 || kd->nbnd > 32 || kd->ngpt != 16 * kd->nbnd) return false;
  for (int b = 0; b < kd->nbnd; ++b)
    if (kd->band_lims_gpt[2 * b] != 16 * b + 1 || kd->band_lims_gpt[2 * b + 1] != 16 * b + 16) return false;
  return true;
}
static const char* kBybandMsg =
    ": by-band fluxes from the packed solver need bands of 16 aligned g-points and a shape the packed solver takes (ngpt <= 512, "
    "nlay >= 8, 16-byte aligned arrays, default solver_variant); use the g-point fluxes and rrnn_sum_byband";

extern "C" int rrnn_rte_lw_byband(rrnn_ctx_t* ctx, const rrnn_kdist_t* kd, int nlay, int ncol, int top_at_1, int n_gauss_angles,
                                  const float* inc_flux_d, const float* tau_d, const float* lay_source_d, const float* lev_source_d,
                                  const float* sfc_source_d, const float* sfc_emis_d, float* flux_up_d, float* flux_dn_d,
                                  float* bnd_flux_up_d, float* bnd_flux_dn_d) {
  rrnn::NvtxRange nvtx_("rte_lw");
  RRNN_CHECK(ctx && kd && tau_d && lay_source_d && lev_source_d && sfc_source_d && sfc_emis_d && flux_up_d && flux_dn_d && bnd_flux_up_d &&
                 bnd_flux_dn_d, "rte_lw: null argument");
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
  if (!bands_of_16(kd) || ctx->solver_variant != 0) return fail(std::string("rte_lw (by band)") + kBybandMsg);
  RRNN_CUDA(cudaSetDevice(ctx->device));
  const int ngpt = kd->ngpt;
  const size_t n = (size_t)ngpt * ncol;
  float* emis_gpt = nullptr;
  RRNN_CUDA(cudaMallocAsync((void**)&emis_gpt, n * sizeof(float), ctx->stream));
  expand_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(kd->nbnd, ngpt, ncol, kd->d_gpt2band, sfc_emis_d, emis_gpt);
  LwParams p{};
  p.ngpt = ngpt; p.nlay = nlay; p.ncol = ncol; p.top_at_1 = top_at_1 ? 1 : 0; p.nmus = n_gauss_angles;
  p.bug_compat = ctx->lw_source_bug_compat;
  p.nchunks = (ngpt + 31) / 32;
  for (int i = 0; i < n_gauss_angles; ++i) { p.Ds[i] = gauss_Ds[n_gauss_angles - 1][i]; p.wts[i] = gauss_wts[n_gauss_angles - 1][i]; }
  p.inc_flux = inc_flux_d; p.tau = tau_d; p.lay_source = lay_source_d; p.lev_source = lev_source_d;
  p.sfc_emis = emis_gpt; p.sfc_source = sfc_source_d; p.flux_up = flux_up_d; p.flux_dn = flux_dn_d;
  p.bnd_up = bnd_flux_up_d; p.bnd_dn = bnd_flux_dn_d; p.nbnd = kd->nbnd;
  const int ps = prof_begin(ctx, K_LW_SOLVER);
  const int rc = launch_lw_v6(ctx, p);
  prof_end(ctx, K_LW_SOLVER, ps);
  cudaFreeAsync(emis_gpt, ctx->stream);
  if (rc < 0) return fail(std::string("rte_lw (by band)") + kBybandMsg);
  if (rc > 0) return rc;
  RRNN_LAUNCH_CHECK(ctx);
  return 0;
}

extern "C" int rrnn_rte_sw_byband(rrnn_ctx_t* ctx, const rrnn_kdist_t* kd, int nlay, int ncol, int top_at_1, const float* mu0_d,
                                  const float* inc_flux_d, const float* sfc_alb_dir_gpt_d, const float* sfc_alb_dif_gpt_d,
                                  const float* inc_flux_dif_d, const float* tau_d, const float* ssa_d, const float* g_d, float* flux_up_d,
                                  float* flux_dn_d, float* flux_dn_dir_d, float* bnd_flux_up_d, float* bnd_flux_dn_d,
                                  float* bnd_flux_dn_dir_d) {
  rrnn::NvtxRange nvtx_("rte_sw");
  RRNN_CHECK(ctx && kd && mu0_d && inc_flux_d && sfc_alb_dir_gpt_d && sfc_alb_dif_gpt_d && tau_d && ssa_d, "rte_sw: null argument");
  RRNN_CHECK(flux_up_d && flux_dn_d && flux_dn_dir_d && bnd_flux_up_d && bnd_flux_dn_d && bnd_flux_dn_dir_d, "rte_sw: no space allocated for fluxes");
  if (ncol == 0) return 0;
  if (!bands_of_16(kd) || ctx->solver_variant != 0) return fail(std::string("rte_sw (by band)") + kBybandMsg);
  RRNN_CUDA(cudaSetDevice(ctx->device));
  SwParams p{};
  p.ngpt = kd->ngpt; p.nlay = nlay; p.ncol = ncol; p.top_at_1 = top_at_1 ? 1 : 0;
  p.nchunks = (kd->ngpt + 31) / 32;
  p.inc_flux = inc_flux_d; p.inc_flux_dif = inc_flux_dif_d; p.tau = tau_d; p.ssa = ssa_d; p.g = g_d; p.mu0 = mu0_d;
  p.alb_dir = sfc_alb_dir_gpt_d; p.alb_dif = sfc_alb_dif_gpt_d; p.flux_up = flux_up_d; p.flux_dn = flux_dn_d; p.flux_dir = flux_dn_dir_d;
  p.bnd_up = bnd_flux_up_d; p.bnd_dn = bnd_flux_dn_d; p.bnd_dir = bnd_flux_dn_dir_d; p.nbnd = kd->nbnd;
  const int ps = prof_begin(ctx, K_SW_SOLVER);
  const int rc = launch_sw_v6(ctx, p, ctx->fast_math || ctx->sw_fast_math);
  prof_end(ctx, K_SW_SOLVER, ps);
  if (rc < 0) return fail(std::string("rte_sw (by band)") + kBybandMsg);
  if (rc > 0) return rc;
  RRNN_LAUNCH_CHECK(ctx);
  return 0;
}
