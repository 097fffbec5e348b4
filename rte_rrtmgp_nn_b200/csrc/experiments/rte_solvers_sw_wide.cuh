// sw_solver_v7: the "wide" shortwave solver (included by rte_solvers_tma.cu inside namespace rrnn::v5).
//
// sw_solver_v6 with FOUR ADJACENT g-points per lane (see lw_solver_v7 for the reasoning): one warp solves 128 g-points of a column,
// the ceil(ngpt/128) chunk-CTAs of a column form the cluster.  The per-warp machinery (TMA issue, mbarrier waits, transposition
// sums, loop control) is spent once per 128 g-points, loads / stores are 16 bytes wide, and the serial part of a layer -- the
// elimination recurrence d = 1/(1 - R alpha), e, f, beta', alpha' (sw_solver_v6; mo_rte_solver_kernels.F90:1526-1637 restated as
// one elimination + one back-substitution) -- runs as two independent packed chains per lane.  The two-stream coefficients
// (sw_two_stream_source, :1366-1480) of H layers x 2 pairs go through the same interleaved batch as sw_solver_v6's 4 layers x 1 pair.
// The per-g-point arithmetic is v6's, instruction for instruction; only the summation order over g-points differs.
// Clear-sky broadband kernel (g == 0, the NN path): clouds, a g array and by-band outputs stay on v6.
#pragma once

constexpr int SW7_ROW = 1536, SW7_F = 512, SW7_A = 1024;   // reverse-sweep row of one layer: e | f | alpha_above, 32 lanes x 16 B each
#ifndef RRNN_V7_SW_S
#define RRNN_V7_SW_S 3      // stages of the input ring (3 x 8 KB also hold the upward sweep's two stages of 12 KB)
#endif
#ifndef RRNN_V7_SW_MINB
#define RRNN_V7_SW_MINB 2   // 2 CTAs of 128 threads: at most 255 registers per thread (the shared memory admits 6 - 7 solver warps per SM)
#endif
#ifndef RRNN_V7_SW_H
#define RRNN_V7_SW_H 2      // layers per two-stream batch (x 2 pairs per lane = 4 interleaved evaluations, as in v6)
#endif

__device__ __forceinline__ void discard_scratch3(const uint8_t* base, uint32_t bytes, int lane) {
  if (RRNN_V5_DISCARD) {   // at most 8 rows of 1536 B = 96 lines: three rounds of the 32 lanes, no loop
    const uint32_t o = (uint32_t)lane * 128u;
    if (o < bytes) asm volatile("discard.global.L2 [%0], 128;" ::"l"(base + o) : "memory");
    if (o + 4096u < bytes) asm volatile("discard.global.L2 [%0], 128;" ::"l"(base + o + 4096u) : "memory");
    if (o + 8192u < bytes) asm volatile("discard.global.L2 [%0], 128;" ::"l"(base + o + 8192u) : "memory");
  }
}

// one layer of the elimination for a pair of g-points: sw_solver_v6's forward step, unchanged.  a_above = alpha on entry.
template <bool FAST>
__device__ __forceinline__ void sw_layer_step2(f2 Rdif, f2 Tdif, f2 Rdir, f2 Tdir, f2 Tnos, f2& dir, f2& beta, f2& alpha, f2& e, f2& f) {
  const f2 s_up = Rdir * dir;
  const f2 s_dn = Tdir * dir;
  dir = Tnos * dir;
  const f2 d = rcp2<FAST>(fnma2(Rdif, alpha, splat2(1.0f)));
  e = d * Tdif;
  f = d * fma2(Rdif, beta, s_up);
  beta = fma2(e, fma2(alpha, s_up, beta), s_dn);
  alpha = fma2(Tdif * e, alpha, Rdif);
}

template <bool FAST, bool TOP>
__global__ void __launch_bounds__(32 * MAX_WARPS, RRNN_V7_SW_MINB) sw_solver_v7(const __grid_constant__ SwV5Params pp, const __grid_constant__ CUtensorMap tm_tau,
                                                   const __grid_constant__ CUtensorMap tm_ssa) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  constexpr int U = 8, S = RRNN_V7_SW_S, SB = 2, H = RRNN_V7_SW_H, NH = U / H, RB = 512;
  constexpr int STAGE = 2 * U * RB;   // tau | ssa, U rows of 128 g-points each
  static_assert(SB * U * SW7_ROW <= S * STAGE, "the upward sweep's stages live in the input ring");
  const SwParams& p = pp.b;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;  // every warp is its own solver
  const int G = p.ngpt, L = p.nlay;
  const uint64_t pol_in = policy_evict_first();
  const uint64_t pol_buf = policy_evict_last();
  cg::cluster_group cluster = cg::this_cluster();
  const int chunk = (int)cluster.block_rank();
  const int csize = (int)cluster.num_blocks();

  uint8_t* smem = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u) + (size_t)warp * pp.warp_smem;
  uint8_t* in_ring = smem;                                                   // [S][STAGE]
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

  const int g = chunk * 128 + 4 * lane;
  const bool act = g < G;                 // ngpt is a multiple of 4: a lane's four g-points are live or not as a whole
  const int gs = act ? g : chunk * 128;   // idle lanes shadow the chunk's first quad and contribute zero
  const f2 live = splat2(act ? 1.0f : 0.0f);
  const int NG = pp.ngroups;
  const int NGF = L / U;
  // this lane's 16-byte slot in the e-segment of scratch row 0 of this solver
  uint8_t* const srow = reinterpret_cast<uint8_t*>(p.scratch) + ((size_t)blockIdx.x * nwarps + warp) * L * SW7_ROW + (size_t)lane * 16u;
  const uint32_t lane_in = (uint32_t)lane * 16u;
  const int top_level = TOP ? 0 : L;
  // which of the 16 reduced values of a group this lane ends up with (as in sw_solver_v6): lanes 0-7 / 16-23 quantity A of layer
  // lane & 7, lanes 8-15 / 24-31 quantity B; only lanes < 16 write
  const int ru = lane & 7;
  const bool rB = (lane & 8) != 0, rW = lane < 16;

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
    auto ldg4u = [](const float* q) { f4 v; v.a = ldg2(q); v.b = ldg2(q + 2); return v; };   // (the per-column arrays are 8-byte aligned)
    f4 dir, beta, alpha = splat4(0.0f);
    {
      const f4 inc = ldg4u(p.inc_flux + gc_off);
      dir.a = (live * inc.a) * splat2(mu0);                                            // :589
      dir.b = (live * inc.b) * splat2(mu0);
      beta = splat4(0.0f);
      if (p.inc_flux_dif) { const f4 d = ldg4u(p.inc_flux_dif + gc_off); beta.a = live * d.a; beta.b = live * d.b; }   // :590
    }
    const f4 a_s = ldg4u(p.alb_dif + gc_off);
    const f4 a_d = ldg4u(p.alb_dir + gc_off);
    __syncwarp();
    {
      const float sd = warp_sum(hsum4(dir)), sb = warp_sum(hsum2((beta.a + dir.a) + (beta.b + dir.b)));
      if (lane == 0) { fdr[top_level] += sd; fdn[top_level] += sb; }
    }
    auto issue_in = [&](int k) {
      if (k < NG) {
        const uint32_t st = (n_in + (uint32_t)k) % S;
        int sh;
        const int rl = box_start<TOP, U>(lay0, k, sh);
        if (elect_one()) {
          const uint32_t bar = bar_in + 8 * st;
          const uint32_t dst = in_a + st * STAGE;
          mbar_expect_tx(bar, STAGE);
          tma_load_2d(dst, &tm_tau, chunk * 128, rl, bar, pol_in);
          tma_load_2d(dst + U * RB, &tm_ssa, chunk * 128, rl, bar, pol_in);
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
    auto flush_fwd = [&]() {  // pend[u] = dir, pend[U + u] = diffuse + dir at the bottom of sweep layer pend_k * U + u
      const float t = tr_reduce<2 * U>(pend, tr, lane);
      const int i = pend_k * U + ru;
      if (pend_k >= 0 && i < L && rW) {
        float* dst = (rB ? fdn : fdr) + (TOP ? i + 1 : L - 1 - i);
        *dst += t;
      }
    };
    auto flush_bwd = [&]() {  // pend[u] = up, pend[U + u] = alpha_above * up at the top of sweep layer pend_k * U + (U - 1 - u)
      const float t = tr_reduce<2 * U>(pend, tr, lane);
      const int i = pend_k * U + (U - 1 - ru);
      if (pend_k >= 0 && i < L && rW) {
        float* dst = (rB ? fdn : fup) + (TOP ? i : L - i);
        *dst += t;
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
      const uint8_t* base = in_ring + st * STAGE + lane_in;
      int shl = 0, nvalid = U;
      if (TAIL) {
        box_start<TOP, U>(lay0, k, shl);
        nvalid = min(U, L - k * U);
      }
      flush_fwd();
      uint8_t* const sg = srow + (size_t)k * (U * SW7_ROW);
      float red[2 * U];
#pragma unroll
      for (int h = 0; h < NH; ++h) {
        if (TAIL && h * H >= nvalid) {  // warp-uniform: nothing of this part belongs to the column
#pragma unroll
          for (int uu = 0; uu < H; ++uu) { red[h * H + uu] = 0.0f; red[U + h * H + uu] = 0.0f; }
          continue;
        }
        f2 tau[2 * H], w0[2 * H], gg[2 * H];
#pragma unroll
        for (int uu = 0; uu < H; ++uu) {
          const int u = h * H + uu;
          const int rl = TAIL ? box_row<TOP, U>(u, shl) : (TOP ? u : U - 1 - u);
          lds22(base + rl * RB, tau[2 * uu], tau[2 * uu + 1]);
          lds22(base + U * RB + rl * RB, w0[2 * uu], w0[2 * uu + 1]);
          gg[2 * uu] = splat2(0.0f);
          gg[2 * uu + 1] = splat2(0.0f);
        }
        f2 Rdif[2 * H], Tdif[2 * H], Rdir[2 * H], Tdir[2 * H], Tnos[2 * H];
        two_stream2_batch<FAST, false, 2 * H>(tau, w0, gg, mu0, mu0_inv, Rdif, Tdif, Rdir, Tdir, Tnos, splat2(pp.neg_zero));
#pragma unroll
        for (int uu = 0; uu < H; ++uu) {
          const int u = h * H + uu;
          if (!TAIL || u < nvalid) {  // warp-uniform
            const f4 above = alpha;   // reflectance of the atmosphere ABOVE this layer: what sweep 2 needs
            f4 e, f;
            sw_layer_step2<FAST>(Rdif[2 * uu], Tdif[2 * uu], Rdir[2 * uu], Tdir[2 * uu], Tnos[2 * uu], dir.a, beta.a, alpha.a, e.a, f.a);
            sw_layer_step2<FAST>(Rdif[2 * uu + 1], Tdif[2 * uu + 1], Rdir[2 * uu + 1], Tdir[2 * uu + 1], Tnos[2 * uu + 1], dir.b, beta.b, alpha.b, e.b, f.b);
            stg_scr4(sg + u * SW7_ROW, e, pol_buf);
            stg_scr4(sg + u * SW7_ROW + SW7_F, f, pol_buf);
            stg_scr4(sg + u * SW7_ROW + SW7_A, above, pol_buf);
          }
          red[u] = hsum4(dir);
          red[U + u] = hsum2((beta.a + dir.a) + (beta.b + dir.b));
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
    f4 Uu;   // source_sfc :1477
    Uu.a = div2<FAST>(fma2(a_s.a, beta.a, dir.a * a_d.a), fnma2(a_s.a, alpha.a, splat2(1.0f))) * live;
    Uu.b = div2<FAST>(fma2(a_s.b, beta.b, dir.b * a_d.b), fnma2(a_s.b, alpha.b, splat2(1.0f))) * live;
    {
      const int sfc = TOP ? L : 0;
      const float su = warp_sum(hsum4(Uu)), sa = warp_sum(hsum2(alpha.a * Uu.a + alpha.b * Uu.b));
      if (lane == 0) { fup[sfc] += su; fdn[sfc] += sa; }
    }
    // ---------------- sweep 2: surface -> top (back substitution) ----------------
    asm volatile("fence.proxy.async.global;" ::: "memory");  // this lane's row stores (generic proxy) before the bulk loads (async proxy)
    __syncwarp();
    auto issue_bb = [&](int j) {  // j-th group of the upward sweep = forward group NG-1-j -> stage (n_bb + j) % SB
      if (j < NG) {
        const int k = NG - 1 - j;
        const uint32_t st = (n_bb + (uint32_t)j) % SB;
        const uint32_t bytes = (uint32_t)min(U, L - k * U) * SW7_ROW;
        if (elect_one()) {
          mbar_expect_tx(bar_bb + 8 * st, bytes);
          bulk_load(in_a + st * (U * SW7_ROW), srow - (size_t)lane * 16u + (size_t)k * (U * SW7_ROW), bytes, bar_bb + 8 * st, pol_buf);
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
      f4 e[U], f[U], a[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const uint8_t* row = in_ring + st * (U * SW7_ROW) + (TAIL ? min(u, nvalid - 1) : u) * SW7_ROW + lane_in;
        e[u] = lds4(row);
        f[u] = lds4(row + SW7_F);
        a[u] = lds4(row + SW7_A);
      }
      flush_bwd();
      float red[2 * U];
#pragma unroll
      for (int u = 0; u < U; ++u) {  // sweep layers k*U + (U-1-u): upwards
        const int uu = U - 1 - u;
        if (!TAIL || uu < nvalid) { Uu.a = fma2(e[uu].a, Uu.a, f[uu].a); Uu.b = fma2(e[uu].b, Uu.b, f[uu].b); }
        red[u] = hsum4(Uu);                                       // upward flux at the level on top of that layer
        red[U + u] = hsum2(a[uu].a * Uu.a + a[uu].b * Uu.b);      // diffuse downward flux there: alpha_above * U (+ beta, added in sweep 1)
      }
      // the rows are in registers: their L2 lines are dead (no write-back; the next column rewrites them in full)
      discard_scratch3(srow - (size_t)lane * 16u + (size_t)k * (U * SW7_ROW), (uint32_t)nvalid * SW7_ROW, lane);
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
    // ---- combine the chunks of this column (see sw_solver_v6)
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
