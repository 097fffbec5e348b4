// NN gas optics for sm_100a, fp32 path: ONE fused kernel per call.
//
// Replaces, for the neural_nets branch of ty_gas_optics_rrtmgp%gas_optics
// (rrtmgp/mo_gas_optics_rrtmgp.F90:368-411 LW, :529-573 SW):
//   compute_nn_inputs            rrtmgp/mo_gas_optics_rrtmgp.F90:618-798   (prologue, in shared memory)
//   get_col_dry                  rrtmgp/mo_gas_optics_rrtmgp.F90:1662-1707 (prologue)
//   output_sgemm_tau/_pfrac/_lw  neural/mod_network_rrtmgp.F90:125-409     (3 SGEMMs + bias/activation)
//   predict_nn_{lw,sw}_blas_sp   rrtmgp/kernels/mo_gas_optics_kernels.F90:690-774, 869-953
//   compute_Planck_source_nn     rrtmgp/kernels/mo_gas_optics_kernels.F90:615-683 (epilogue)
// so that nn_inputs, col_dry, the hidden activations and the Planck fraction never exist in HBM and every
// output float (tau, lay_source, lev_source | tau, ssa) is written exactly once, g-point-first, with 16-byte
// stores that cover 512 contiguous bytes per warp.
//
// Structure: a persistent CTA (256 threads, one per SM) keeps both networks' weights in shared memory and
// walks tiles of 64 samples (= 64 consecutive (layer, column) points).  Hidden layers: 4x4 register tiles.
// Output layer: warp w owns samples 8w..8w+7, lane owns g-points 4*lane..4*lane+3 (+128): an 8 x 8 register
// tile fed by two broadcast and two conflict-free 128-bit shared loads per k.
// fp32 FFMA throughout with the reference's summation order over the input index -- this is the parity path
// (tau relative error <= 1e-4); the tensor-core variant lives in gas_optics_nn_tc.cu.
#include "common.cuh"

namespace rrnn {

constexpr int S_TILE = 64;  // samples per tile
constexpr int GO_THREADS = 256;

enum Epi { EPI_LW2 = 0, EPI_LWBOTH = 1, EPI_SW = 2, EPI_TAU = 3, EPI_PFRAC = 4, EPI_RAW = 5 };

struct NetDev {
  int nlayers;
  int dims[MAX_LAYERS + 1];
  int act[MAX_LAYERS];
  const float* w[MAX_LAYERS];
  const float* b[MAX_LAYERS];
  const float* ymean;
  const float* ystd;
};

struct GasIn {
  const float* ptr;
  float value;
  int mode;  // 0 scalar, 1 per-layer, 2 (nlay,ncol), -1 missing (vmr = 0)
};

struct GoParams {
  int epi, ncol, nlay, ngpt, nx, fields;  // fields = 1: raw atmospheric fields; 0: precomputed x/coldry
  long long nsamples;
  const float *play, *plev, *tlay, *tlev, *tsfc;
  GasIn gas[MAX_NN_INPUTS];
  float xmin[MAX_NN_INPUTS], xmax[MAX_NN_INPUTS];
  const float* x;       // (nx, nsamples)
  const float* coldry;  // (nsamples)
  NetDev net[2];
  int nnets;
  int nbnd, ntemp;
  const int* gpt2band;
  const float* totplnk;
  float temp_ref_min, totplnk_delta;
  float *out0, *out1, *out2, *sfc_source, *sfc_jac;
  const float* toa_in;  // unused
};

__device__ __forceinline__ float act_apply(int code, float x) {
  switch (code) {
    case RRNN_ACT_SOFTSIGN: return x / (fabsf(x) + 1.0f);
    case RRNN_ACT_RELU: return fmaxf(0.0f, x);
    case RRNN_ACT_SIGMOID: return 1.0f / (1.0f + expf(-x));
    case RRNN_ACT_HARD_SIGMOID: return fmaxf(0.0f, fminf(1.0f, 0.2f * x + 0.5f));
    default: return x;
  }
}

__host__ __device__ inline int pad4(int n) { return (n + 3) & ~3; }

// shared-memory plan for one network: per layer [K][OP] weights then [OP] bias
struct NetSmem {
  int w_off[MAX_LAYERS];
  int b_off[MAX_LAYERS];
  int op[MAX_LAYERS];
  int total;
};

__host__ __device__ inline NetSmem plan_net(const NetDev& n, int np_last, int base) {
  NetSmem s;
  int off = base;
  for (int l = 0; l < n.nlayers; ++l) {
    const int K = n.dims[l];
    const int OP = (l == n.nlayers - 1) ? np_last : pad4(n.dims[l + 1]);
    s.op[l] = OP;
    s.w_off[l] = off;
    off += K * OP;
    s.b_off[l] = off;
    off += OP;
  }
  s.total = off - base;
  return s;
}

__device__ void load_net(const NetDev& n, const NetSmem& s, float* smem) {
  for (int l = 0; l < n.nlayers; ++l) {
    const int K = n.dims[l], O = n.dims[l + 1], OP = s.op[l];
    float* w = smem + s.w_off[l];
    for (int i = threadIdx.x; i < K * OP; i += blockDim.x) {
      const int k = i / OP, o = i - k * OP;
      w[i] = (o < O) ? n.w[l][(size_t)k * O + o] : 0.0f;
    }
    float* b = smem + s.b_off[l];
    for (int i = threadIdx.x; i < OP; i += blockDim.x) b[i] = (i < O) ? n.b[l][i] : 0.0f;
  }
}

// out[h][s] = act(sum_k in[k][s] * W[k][h] + b[h]); 4 samples x 4 neurons per work item
__device__ void hidden_layer(const float* __restrict__ W, const float* __restrict__ b, int K, int O, int OP, int act,
                             const float* __restrict__ in, float* __restrict__ out) {
  const int nhg = OP >> 2;
  const int items = (S_TILE / 4) * nhg;
  for (int it = threadIdx.x; it < items; it += blockDim.x) {
    const int hg = it % nhg, sg = it / nhg;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;
#pragma unroll 2
    for (int k = 0; k < K; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(in + k * S_TILE + 4 * sg);
      const float4 w = *reinterpret_cast<const float4*>(W + k * OP + 4 * hg);
      const float av[4] = {a.x, a.y, a.z, a.w};
      const float wv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(wv[j], av[i], acc[i][j]);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int h = 4 * hg + j;
      if (h < O) {
        const float bb = b[h];
        float4 r;
        r.x = act_apply(act, acc[0][j] + bb);
        r.y = act_apply(act, acc[1][j] + bb);
        r.z = act_apply(act, acc[2][j] + bb);
        r.w = act_apply(act, acc[3][j] + bb);
        *reinterpret_cast<float4*>(out + h * S_TILE + 4 * sg) = r;
      }
    }
  }
}

// acc[s][4*j+c] = sum_k in[k][8*warp+s] * W[k][128*j + 4*lane + c]
template <int NG>
__device__ __forceinline__ void last_layer(const float* __restrict__ W, int K, int NP, const float* __restrict__ in,
                                           int warp, int lane, float (&acc)[8][4 * NG]) {
#pragma unroll
  for (int s = 0; s < 8; ++s)
#pragma unroll
    for (int c = 0; c < 4 * NG; ++c) acc[s][c] = 0.0f;
  const float* ip = in + 8 * warp;
  const float* wp = W + 4 * lane;
#pragma unroll 2
  for (int k = 0; k < K; ++k) {
    const float4 a0 = *reinterpret_cast<const float4*>(ip + k * S_TILE);
    const float4 a1 = *reinterpret_cast<const float4*>(ip + k * S_TILE + 4);
    const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
    for (int j = 0; j < NG; ++j) {
      const float4 w = *reinterpret_cast<const float4*>(wp + k * NP + 128 * j);
      const float wv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
      for (int s = 0; s < 8; ++s)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[s][4 * j + c] = fmaf(wv[c], av[s], acc[s][4 * j + c]);
    }
  }
}

// interpolate1D, rrtmgp/kernels/mo_gas_optics_kernels.F90:1024-1043 for one band
__device__ __forceinline__ float planck_interp(float T, float tmin, float delta, const float* __restrict__ tab, int ntemp) {
  const float val0 = (T - tmin) / delta;
  const int iv = (int)val0;
  const float frac = val0 - (float)iv;
  int idx = iv + 1;
  idx = max(1, idx);
  idx = min(ntemp - 1, idx);
  const float t0 = tab[idx - 1];
  return t0 + frac * (tab[idx] - t0);
}

template <int EPIT, int NG>
__global__ void __launch_bounds__(GO_THREADS, 1) gas_optics_kernel(const GoParams p) {
  extern __shared__ __align__(16) float smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int NP = 128 * NG;
  const int nx = p.nx, L = p.nlay, N = p.ngpt;
  constexpr bool kPlanck = (EPIT == EPI_LW2 || EPIT == EPI_LWBOTH);
  const int nnets = p.nnets;

  // ---- shared memory carve-up (must match go_smem_floats on the host) ----
  int off = 0;
  NetSmem ns[2];
  int hmax = pad4(nx);
  for (int n = 0; n < nnets; ++n) {
    ns[n] = plan_net(p.net[n], NP, off);
    off += ns[n].total;
    for (int l = 1; l < p.net[n].nlayers; ++l) hmax = max(hmax, pad4(p.net[n].dims[l]));
  }
  float* xT = smem + off;      off += pad4(nx) * S_TILE;
  float* actA = smem + off;    off += hmax * S_TILE;
  float* actB = smem + off;    off += hmax * S_TILE;
  float* coldry_s = smem + off; off += S_TILE;
  int* flag_s = reinterpret_cast<int*>(smem + off); off += S_TILE;
  float* plk_lay = smem + off;
  float* plk_lev = plk_lay + (kPlanck ? S_TILE * p.nbnd : 0);
  float* plk_bot = plk_lev + (kPlanck ? S_TILE * p.nbnd : 0);
  float* plk_sfc = plk_bot + (kPlanck ? S_TILE * p.nbnd : 0);
  float* plk_jac = plk_sfc + (kPlanck ? S_TILE * p.nbnd : 0);

  for (int n = 0; n < nnets; ++n) load_net(p.net[n], ns[n], smem);

  // per-thread output-column constants
  int gidx[NG];
  bool gok[NG];
#pragma unroll
  for (int j = 0; j < NG; ++j) { gidx[j] = 128 * j + 4 * lane; }

  // surface layer (1-based) as the reference chooses it from column 1: merge(1,nlay,play(1,1) > play(nlay,1))
  int sfc_lay0 = -1;
  if (kPlanck && p.fields) sfc_lay0 = (p.play[0] > p.play[L - 1]) ? 0 : L - 1;

  const long long ntiles = (p.nsamples + S_TILE - 1) / S_TILE;
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long s0 = tile * S_TILE;
    __syncthreads();  // previous tile's readers are done (also orders load_net on the first pass)
    // ---------------- prologue: inputs, col_dry, Planck band values ----------------
    if (p.fields) {
      for (int it = tid; it < S_TILE * nx; it += GO_THREADS) {
        const int s = it % S_TILE, k = it / S_TILE;
        const long long smp = s0 + s;
        float v = 0.0f;
        if (smp < p.nsamples) {
          const int lay = (int)(smp % L);
          float raw;
          if (k == 0) raw = p.tlay[smp];
          else if (k == 1) raw = logf(p.play[smp]);
          else {
            const GasIn gi = p.gas[k];
            if (gi.mode == 2) raw = gi.ptr[smp];
            else if (gi.mode == 1) raw = gi.ptr[lay];
            else if (gi.mode == 0) raw = gi.value;
            else raw = 0.0f;
            if (k == 2 || k == 3) raw = sqrtf(sqrtf(raw));
          }
          v = (raw - p.xmin[k]) / (p.xmax[k] - p.xmin[k]);
        }
        xT[k * S_TILE + s] = v;
      }
      for (int s = tid; s < S_TILE; s += GO_THREADS) {
        const long long smp = s0 + s;
        float cd = 0.0f;
        int fl = 0;
        if (smp < p.nsamples) {
          const long long col = smp / L;
          const int lay = (int)(smp - col * L);
          // get_col_dry :1697-1703
          const float h = (p.gas[2].mode == 2) ? p.gas[2].ptr[smp] : (p.gas[2].mode == 1 ? p.gas[2].ptr[lay] : p.gas[2].value);
          const float dp = fabsf(p.plev[col * (L + 1) + lay] - p.plev[col * (L + 1) + lay + 1]);
          const float fact = 1.0f / (1.0f + h);
          const float m_air = (0.028964f + 0.018016f * h) * fact;
          cd = 10.0f * dp * 6.02214076e23f * fact / (1000.0f * m_air * 100.0f * 9.80665f);
          if (lay == L - 1) fl |= 1;
          if (lay == sfc_lay0) fl |= 2;
        }
        coldry_s[s] = cd;
        flag_s[s] = fl;
      }
      if (kPlanck) {
        for (int it = tid; it < S_TILE * p.nbnd; it += GO_THREADS) {
          const int b = it % p.nbnd, s = it / p.nbnd;
          const long long smp = s0 + s;
          if (smp < p.nsamples) {
            const long long col = smp / L;
            const int lay = (int)(smp - col * L);
            const float* tab = p.totplnk + (size_t)b * p.ntemp;
            plk_lay[it] = planck_interp(p.tlay[smp], p.temp_ref_min, p.totplnk_delta, tab, p.ntemp);
            plk_lev[it] = planck_interp(p.tlev[col * (L + 1) + lay], p.temp_ref_min, p.totplnk_delta, tab, p.ntemp);
            if (lay == L - 1)
              plk_bot[it] = planck_interp(p.tlev[col * (L + 1) + L], p.temp_ref_min, p.totplnk_delta, tab, p.ntemp);
            if (lay == sfc_lay0) {
              const float ts = p.tsfc[col];
              const float a = planck_interp(ts, p.temp_ref_min, p.totplnk_delta, tab, p.ntemp);
              plk_sfc[it] = a;
              plk_jac[it] = planck_interp(ts + 1.0f, p.temp_ref_min, p.totplnk_delta, tab, p.ntemp) - a;
            }
          }
        }
      }
    } else {
      for (int it = tid; it < S_TILE * nx; it += GO_THREADS) {
        const int k = it % nx, s = it / nx;
        const long long smp = s0 + s;
        xT[k * S_TILE + s] = (smp < p.nsamples) ? p.x[smp * nx + k] : 0.0f;
      }
      for (int s = tid; s < S_TILE; s += GO_THREADS) {
        const long long smp = s0 + s;
        coldry_s[s] = (p.coldry && smp < p.nsamples) ? p.coldry[smp] : 1.0f;
        flag_s[s] = 0;
      }
    }
    __syncthreads();

    float keep[(EPIT == EPI_SW) ? 8 : 1][(EPIT == EPI_SW) ? 4 * NG : 1];

    for (int n = 0; n < nnets; ++n) {
      const NetDev& net = p.net[n];
      const int nl = net.nlayers;
      // ---------------- hidden layers ----------------
      const float* in = xT;
      float* outb = actA;
      for (int l = 0; l < nl - 1; ++l) {
        hidden_layer(smem + ns[n].w_off[l], smem + ns[n].b_off[l], net.dims[l], net.dims[l + 1], ns[n].op[l], net.act[l],
                     in, outb);
        __syncthreads();
        in = outb;
        outb = (outb == actA) ? actB : actA;
      }
      // ---------------- output layer ----------------
      float acc[8][4 * NG];
      last_layer<NG>(smem + ns[n].w_off[nl - 1], net.dims[nl - 1], NP, in, warp, lane, acc);
      const float* b3 = smem + ns[n].b_off[nl - 1];
      const int Nout = net.dims[nl];
#pragma unroll
      for (int j = 0; j < NG; ++j) gok[j] = gidx[j] < Nout;

      // which post-processing does this network get?
      const bool is_tau = (EPIT == EPI_TAU) || (EPIT == EPI_SW) || (EPIT == EPI_LW2 && n == 0);
      const bool is_pfrac = (EPIT == EPI_PFRAC) || (EPIT == EPI_LW2 && n == 1);

      if (EPIT == EPI_LWBOTH) {
        // one network, 2*ngpt outputs (ngpt == 128): group 0 -> tau, group 1 -> Planck fraction
        // rrtmgp/kernels/mo_gas_optics_kernels.F90:745-767
#pragma unroll
        for (int s = 0; s < 8; ++s) {
          const int sl = 8 * warp + s;
          const long long smp = s0 + sl;
          if (smp >= p.nsamples) continue;
          const long long col = smp / L;
          const int lay = (int)(smp - col * L);
          const int g = 4 * lane;
          const float cd = coldry_s[sl];
          float tv[4], pv[4];
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            float t = net.ystd[g + c] * (acc[s][c] + b3[g + c]) + net.ymean[g + c];
            t = t * t; t = t * t; t = t * t;
            tv[c] = t * cd;
            const float z = acc[s][4 * (NG - 1) + c] + b3[128 * (NG - 1) + g + c];
            pv[c] = z * z;
          }
          st_stream4(reinterpret_cast<float4*>(p.out0 + smp * N + g), make_float4(tv[0], tv[1], tv[2], tv[3]));
          float lv[4], ly[4];
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const int b = p.gpt2band[g + c];
            ly[c] = pv[c] * plk_lay[sl * p.nbnd + b];
            lv[c] = pv[c] * plk_lev[sl * p.nbnd + b];
          }
          st_stream4(reinterpret_cast<float4*>(p.out1 + smp * N + g), make_float4(ly[0], ly[1], ly[2], ly[3]));
          st_stream4(reinterpret_cast<float4*>(p.out2 + (col * (L + 1) + lay) * N + g), make_float4(lv[0], lv[1], lv[2], lv[3]));
          const int fl = flag_s[sl];
          if (fl & 1) {
            float q[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) q[c] = pv[c] * plk_bot[sl * p.nbnd + p.gpt2band[g + c]];
            st_stream4(reinterpret_cast<float4*>(p.out2 + (col * (L + 1) + L) * N + g), make_float4(q[0], q[1], q[2], q[3]));
          }
          if (fl & 2) {
            float q[4], r[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              const int b = p.gpt2band[g + c];
              q[c] = pv[c] * plk_sfc[sl * p.nbnd + b];
              r[c] = pv[c] * plk_jac[sl * p.nbnd + b];
            }
            *reinterpret_cast<float4*>(p.sfc_source + col * N + g) = make_float4(q[0], q[1], q[2], q[3]);
            *reinterpret_cast<float4*>(p.sfc_jac + col * N + g) = make_float4(r[0], r[1], r[2], r[3]);
          }
        }
      } else if (is_tau) {
        // output_sgemm_tau epilogue, neural/mod_network_rrtmgp.F90:209-231
        float ysd[4 * NG], ymn[4 * NG], bb[4 * NG];
#pragma unroll
        for (int j = 0; j < NG; ++j)
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const int g = gidx[j] + c;
            ysd[4 * j + c] = gok[j] ? net.ystd[g] : 0.0f;
            ymn[4 * j + c] = gok[j] ? net.ymean[g] : 0.0f;
            bb[4 * j + c] = b3[g];
          }
#pragma unroll
        for (int s = 0; s < 8; ++s) {
          const int sl = 8 * warp + s;
          const long long smp = s0 + sl;
          if (smp >= p.nsamples) continue;
          const float cd = coldry_s[sl];
#pragma unroll
          for (int j = 0; j < NG; ++j) {
            if (!gok[j]) continue;
            float v[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              float t = ysd[4 * j + c] * (acc[s][4 * j + c] + bb[4 * j + c]) + ymn[4 * j + c];
              t = t * t; t = t * t; t = t * t;  // **8
              v[c] = t * cd;
            }
            if (EPIT == EPI_SW) {
              if (n == 0) {
                if (nnets == 1) {
                  st_stream4(reinterpret_cast<float4*>(p.out0 + smp * N + gidx[j]), make_float4(v[0], v[1], v[2], v[3]));
                } else {
#pragma unroll
                  for (int c = 0; c < 4; ++c) keep[s][4 * j + c] = v[c];
                }
              } else {
                // Rayleigh network: tau_tot = tau_abs + tau_ray; ssa = tau_ray / tau_tot (:224-229, no zero guard)
                float tt[4], sa[4];
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                  tt[c] = keep[s][4 * j + c] + v[c];
                  sa[c] = v[c] / tt[c];
                }
                st_stream4(reinterpret_cast<float4*>(p.out0 + smp * N + gidx[j]), make_float4(tt[0], tt[1], tt[2], tt[3]));
                st_stream4(reinterpret_cast<float4*>(p.out1 + smp * N + gidx[j]), make_float4(sa[0], sa[1], sa[2], sa[3]));
                if (p.out2)  // g = 0, rrtmgp/mo_gas_optics_rrtmgp.F90:560-567
                  st_stream4(reinterpret_cast<float4*>(p.out2 + smp * N + gidx[j]), make_float4(0.f, 0.f, 0.f, 0.f));
              }
            } else if (EPIT == EPI_TAU && p.out1) {
              // output2 present: out1 holds tau_abs in, tau_tot out; out0 receives ssa
              float4 a = *reinterpret_cast<const float4*>(p.out1 + smp * N + gidx[j]);
              float4 tt = make_float4(a.x + v[0], a.y + v[1], a.z + v[2], a.w + v[3]);
              *reinterpret_cast<float4*>(p.out1 + smp * N + gidx[j]) = tt;
              *reinterpret_cast<float4*>(p.out0 + smp * N + gidx[j]) = make_float4(v[0] / tt.x, v[1] / tt.y, v[2] / tt.z, v[3] / tt.w);
            } else {
              st_stream4(reinterpret_cast<float4*>(p.out0 + smp * N + gidx[j]), make_float4(v[0], v[1], v[2], v[3]));
            }
          }
        }
      } else if (is_pfrac) {
        // output_sgemm_pfrac epilogue (:309-312) [+ compute_Planck_source_nn when fused]
        const int lact = net.act[nl - 1];
#pragma unroll
        for (int s = 0; s < 8; ++s) {
          const int sl = 8 * warp + s;
          const long long smp = s0 + sl;
          if (smp >= p.nsamples) continue;
          const long long col = smp / L;
          const int lay = (int)(smp - col * L);
          const int fl = flag_s[sl];
#pragma unroll
          for (int j = 0; j < NG; ++j) {
            if (!gok[j]) continue;
            const int g = gidx[j];
            float pv[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              const float z = act_apply(lact, acc[s][4 * j + c] + b3[g + c]);
              pv[c] = z * z;
            }
            if (EPIT == EPI_PFRAC) {
              st_stream4(reinterpret_cast<float4*>(p.out0 + smp * N + g), make_float4(pv[0], pv[1], pv[2], pv[3]));
            } else {
              int bnd[4];
#pragma unroll
              for (int c = 0; c < 4; ++c) bnd[c] = p.gpt2band[g + c];
              float ly[4], lv[4];
#pragma unroll
              for (int c = 0; c < 4; ++c) {
                ly[c] = pv[c] * plk_lay[sl * p.nbnd + bnd[c]];
                lv[c] = pv[c] * plk_lev[sl * p.nbnd + bnd[c]];
              }
              st_stream4(reinterpret_cast<float4*>(p.out1 + smp * N + g), make_float4(ly[0], ly[1], ly[2], ly[3]));
              st_stream4(reinterpret_cast<float4*>(p.out2 + (col * (L + 1) + lay) * N + g), make_float4(lv[0], lv[1], lv[2], lv[3]));
              if (fl & 1) {
                float q[4];
#pragma unroll
                for (int c = 0; c < 4; ++c) q[c] = pv[c] * plk_bot[sl * p.nbnd + bnd[c]];
                st_stream4(reinterpret_cast<float4*>(p.out2 + (col * (L + 1) + L) * N + g), make_float4(q[0], q[1], q[2], q[3]));
              }
              if (fl & 2) {
                float q[4], r[4];
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                  q[c] = pv[c] * plk_sfc[sl * p.nbnd + bnd[c]];
                  r[c] = pv[c] * plk_jac[sl * p.nbnd + bnd[c]];
                }
                *reinterpret_cast<float4*>(p.sfc_source + col * N + g) = make_float4(q[0], q[1], q[2], q[3]);
                *reinterpret_cast<float4*>(p.sfc_jac + col * N + g) = make_float4(r[0], r[1], r[2], r[3]);
              }
            }
          }
        }
      } else {
        // EPI_RAW: output_sgemm_lw, bias only (:398-404)
#pragma unroll
        for (int s = 0; s < 8; ++s) {
          const int sl = 8 * warp + s;
          const long long smp = s0 + sl;
          if (smp >= p.nsamples) continue;
#pragma unroll
          for (int j = 0; j < NG; ++j) {
            if (!gok[j]) continue;
            const int g = gidx[j];
            st_stream4(reinterpret_cast<float4*>(p.out0 + smp * Nout + g),
                       make_float4(acc[s][4 * j] + b3[g], acc[s][4 * j + 1] + b3[g + 1], acc[s][4 * j + 2] + b3[g + 2],
                                   acc[s][4 * j + 3] + b3[g + 3]));
          }
        }
      }
      __syncthreads();  // activation buffers are reused by the next network
    }
  }
}

// toa_src(igpt,icol) = solar_source(igpt), rrtmgp/mo_gas_optics_rrtmgp.F90:594-599
__global__ void toa_src_kernel(int ngpt, int ncol, const float* __restrict__ solar, float* __restrict__ toa) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < (size_t)ngpt * ncol) toa[i] = solar[i % ngpt];
}

static size_t go_smem_floats(const GoParams& p, int NG, bool planck) {
  const int NP = 128 * NG;
  size_t off = 0;
  int hmax = pad4(p.nx);
  for (int n = 0; n < p.nnets; ++n) {
    NetSmem s = plan_net(p.net[n], NP, (int)off);
    off += s.total;
    for (int l = 1; l < p.net[n].nlayers; ++l) hmax = std::max(hmax, pad4(p.net[n].dims[l]));
  }
  off += (size_t)pad4(p.nx) * S_TILE + 2 * (size_t)hmax * S_TILE + 2 * S_TILE;
  if (planck) off += 5 * (size_t)S_TILE * p.nbnd;
  return off;
}

static bool on_device(const rrnn_model_t* m) { return m && m->d_wpack && m->d_bpack; }

static void fill_net(NetDev& d, const rrnn_model_t* m) {
  d.nlayers = m->nlayers;
  for (int i = 0; i <= m->nlayers; ++i) d.dims[i] = m->dims[i];
  for (int i = 0; i < m->nlayers; ++i) {
    d.act[i] = m->act[i];
    d.w[i] = m->d_wpack + m->w_off[i];
    d.b[i] = m->d_bpack + m->b_off[i];
  }
  d.ymean = m->d_ymean;
  d.ystd = m->d_ystd;
}

template <int EPIT>
static int launch_go(rrnn_ctx_t* ctx, const GoParams& p, int nout_max, int prof_kind = -1) {
  const int NG = (nout_max > 128) ? 2 : 1;
  const bool planck = (EPIT == EPI_LW2 || EPIT == EPI_LWBOTH);
  const size_t smem = go_smem_floats(p, NG, planck) * sizeof(float);
  if (smem > ctx->smem_optin) return fail("NN gas optics: networks too large for the shared-memory resident design");
  const long long ntiles = (p.nsamples + S_TILE - 1) / S_TILE;
  const unsigned grid = (unsigned)std::min<long long>(ntiles, ctx->num_sms);
  if (grid == 0) return 0;
  const int ps = prof_kind >= 0 ? prof_begin(ctx, prof_kind) : -1;
  if (NG == 2) {
    RRNN_CUDA(cudaFuncSetAttribute(gas_optics_kernel<EPIT, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    gas_optics_kernel<EPIT, 2><<<grid, GO_THREADS, smem, ctx->stream>>>(p);
  } else {
    RRNN_CUDA(cudaFuncSetAttribute(gas_optics_kernel<EPIT, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    gas_optics_kernel<EPIT, 1><<<grid, GO_THREADS, smem, ctx->stream>>>(p);
  }
  if (prof_kind >= 0) prof_end(ctx, prof_kind, ps);
  RRNN_LAUNCH_CHECK(ctx);
  if (prof_kind >= 0) {  // a gas_optics() call (the stand-alone output_sgemm_* entry points are FFMA by definition)
    ctx->last_nn_kernel = RRNN_NN_KERNEL_FFMA;
    ctx->nn_ffma_launches++;
  }
  return 0;
}

static std::string trim_name(const char* s) {
  std::string r(s, strnlen(s, 32));
  while (!r.empty() && (r.back() == ' ' || r.back() == '\0')) r.pop_back();
  size_t b = 0;
  while (b < r.size() && r[b] == ' ') ++b;
  return r.substr(b);
}

// Map ty_gas_concs entries onto the network's inputs BY NAME (compute_nn_inputs :708-760).
int map_gases(const rrnn_model_t* m, const rrnn_gas_t* gases, int ngas, GoParams& p) {
  const int nx = m->dims[0];
  if (nx > MAX_NN_INPUTS) return fail("compute_nn_inputs: too many NN inputs");
  if (nx < 4) return fail("compute_nn_inputs: the network must take tlay, play, h2o, o3 as its first inputs");
  p.nx = nx;
  for (int i = 0; i < nx; ++i) {
    p.xmin[i] = m->xmin[i];
    p.xmax[i] = m->xmax[i];
    p.gas[i].ptr = nullptr; p.gas[i].value = 0.0f; p.gas[i].mode = -1;
    if (i < 2) continue;
    for (int g = 0; g < ngas; ++g) {
      if (trim_name(gases[g].name) == m->input_names[i]) {
        p.gas[i].ptr = gases[g].conc;
        p.gas[i].value = gases[g].value;
        p.gas[i].mode = gases[g].ndims;
        if (gases[g].ndims < 0 || gases[g].ndims > 2) return fail("gas_concs: ndims must be 0, 1 or 2");
        if (gases[g].ndims > 0 && !gases[g].conc) return fail("gas_concs: null concentration pointer");
        break;
      }
    }
    if (i < 4 && p.gas[i].mode < 0)
      return fail(std::string("compute_nn_inputs: gas ") + m->input_names[i] + " is required but was not provided");
  }
  return 0;
}

}  // namespace rrnn

using namespace rrnn;

int rrnn_gas_optics_tc(rrnn_ctx_t* ctx, int mode, const rrnn_kdist_t* kd, const rrnn_model_t* const* models, int nmodels, int ncol, int nlay,
                       const float* play, const float* plev, const float* tlay, const float* tlev, const float* tsfc,
                       const rrnn_gas_t* gases, int ngas, float* out0, float* out1, float* out2, float* sfc_source,
                       float* sfc_jac, int prof_kind, float* planck_lay = nullptr,
                       float* planck_lev = nullptr);  // gas_optics_tc.cu; -1 = configuration not supported

extern "C" int rrnn_gas_optics_lw(rrnn_ctx_t* ctx, const rrnn_kdist_t* kd, const rrnn_model_t* const* models, int nmodels,
                                  int ncol, int nlay, const float* play_d, const float* plev_d, const float* tlay_d,
                                  const float* tsfc_d, const rrnn_gas_t* gases, int ngas, const float* tlev_d,
                                  float* tau_d, float* lay_source_d, float* lev_source_d, float* sfc_source_d,
                                  float* sfc_source_Jac_d) {
  rrnn::NvtxRange nvtx_("gas_optics (LW)");
  RRNN_CHECK(ctx && kd && models, "gas_optics(): null handle");
  RRNN_CHECK(nmodels == 1 || nmodels == 2, "gas_optics(): neural_nets must hold 1 or 2 networks for the longwave");
  RRNN_CHECK(kd->d_totplnk, "gas_optics(): k-distribution has no Planck table (not a longwave k-distribution)");
  RRNN_CHECK(nlay >= 2 && ncol >= 0, "gas_optics(): bad extents");
  for (int n = 0; n < nmodels; ++n) RRNN_CHECK(on_device(models[n]), "gas_optics(): network was loaded without a device context");
  if (ncol == 0) return 0;
  RRNN_CUDA(cudaSetDevice(ctx->device));
  GoParams p{};
  if (int rc = map_gases(models[0], gases, ngas, p)) return rc;
  p.fields = 1; p.ncol = ncol; p.nlay = nlay; p.ngpt = kd->ngpt; p.nsamples = (long long)ncol * nlay;
  p.play = play_d; p.plev = plev_d; p.tlay = tlay_d; p.tsfc = tsfc_d;
  float* tlev_tmp = nullptr;
  if (!tlev_d) {
    RRNN_CUDA(cudaMallocAsync((void**)&tlev_tmp, (size_t)ncol * (nlay + 1) * sizeof(float), ctx->stream));
    if (int rc = rrnn_interp_tlev(ctx, ncol, nlay, play_d, plev_d, tlay_d, tlev_tmp)) return rc;
    p.tlev = tlev_tmp;
  } else {
    p.tlev = tlev_d;
  }
  p.nnets = nmodels;
  for (int n = 0; n < nmodels; ++n) fill_net(p.net[n], models[n]);
  p.nbnd = kd->nbnd; p.ntemp = kd->ntemp; p.gpt2band = kd->d_gpt2band; p.totplnk = kd->d_totplnk;
  p.temp_ref_min = kd->temp_ref_min; p.totplnk_delta = kd->totplnk_delta;
  p.out0 = tau_d; p.out1 = lay_source_d; p.out2 = lev_source_d; p.sfc_source = sfc_source_d; p.sfc_jac = sfc_source_Jac_d;
  int rc = -1;
  if (ctx->nn_tensor_cores) {
    rc = rrnn_gas_optics_tc(ctx, 0, kd, models, nmodels, ncol, nlay, play_d, plev_d, tlay_d, p.tlev, tsfc_d, gases, ngas, tau_d,
                            lay_source_d, lev_source_d, sfc_source_d, sfc_source_Jac_d, K_GAS_LW);
  }
  if (rc >= 0) {
    // done on the tensor cores (or failed with an error)
  } else if (nmodels == 2) {
    RRNN_CHECK(models[0]->dims[models[0]->nlayers] == kd->ngpt && models[1]->dims[models[1]->nlayers] == kd->ngpt,
               "gas_optics(): network output size differs from the number of g-points");
    RRNN_CHECK(models[0]->d_ymean && models[0]->d_ystd, "output_sgemm_tau: NN output scaling coefficients missing");
    RRNN_CHECK(models[1]->dims[0] == models[0]->dims[0], "gas_optics(): the two networks take different inputs");
    RRNN_CHECK(kd->ngpt % 4 == 0 && kd->ngpt <= 256, "gas_optics(): ngpt must be a multiple of 4 and <= 256");
    rc = launch_go<EPI_LW2>(ctx, p, kd->ngpt, K_GAS_LW);
  } else {
    RRNN_CHECK(models[0]->dims[models[0]->nlayers] == 2 * kd->ngpt, "gas_optics(): 'both' network must have 2*ngpt outputs");
    RRNN_CHECK(kd->ngpt == 128, "gas_optics(): single-network longwave models are supported for ngpt = 128");
    RRNN_CHECK(models[0]->d_ymean && models[0]->d_ystd, "output_sgemm_tau: NN output scaling coefficients missing");
    rc = launch_go<EPI_LWBOTH>(ctx, p, 256, K_GAS_LW);
  }
  if (tlev_tmp) cudaFreeAsync(tlev_tmp, ctx->stream);
  return rc;
}

// Longwave gas optics with the sources left factored (see include/rrnn.h): tensor-core kernel only.
extern "C" int rrnn_gas_optics_lw_compact(rrnn_ctx_t* ctx, const rrnn_kdist_t* kd, const rrnn_model_t* const* models, int nmodels,
                                          int ncol, int nlay, const float* play_d, const float* plev_d, const float* tlay_d,
                                          const float* tsfc_d, const rrnn_gas_t* gases, int ngas, const float* tlev_d,
                                          float* tau_d, float* pfrac_d, float* planck_lay_d, float* planck_lev_d,
                                          float* sfc_source_d, float* sfc_source_Jac_d) {
  rrnn::NvtxRange nvtx_("gas_optics (LW)");
  RRNN_CHECK(ctx && kd && models, "gas_optics(): null handle");
  RRNN_CHECK(nmodels == 1 || nmodels == 2, "gas_optics(): neural_nets must hold 1 or 2 networks for the longwave");
  RRNN_CHECK(kd->d_totplnk, "gas_optics(): k-distribution has no Planck table (not a longwave k-distribution)");
  RRNN_CHECK(nlay >= 2 && ncol >= 0, "gas_optics(): bad extents");
  RRNN_CHECK(tau_d && pfrac_d && planck_lay_d && planck_lev_d && sfc_source_d && sfc_source_Jac_d, "gas_optics (compact sources): null output");
  for (int n = 0; n < nmodels; ++n) RRNN_CHECK(on_device(models[n]), "gas_optics(): network was loaded without a device context");
  if (ncol == 0) return 0;
  RRNN_CUDA(cudaSetDevice(ctx->device));
  float* tlev_tmp = nullptr;
  if (!tlev_d) {
    RRNN_CUDA(cudaMallocAsync((void**)&tlev_tmp, (size_t)ncol * (nlay + 1) * sizeof(float), ctx->stream));
    if (int rc = rrnn_interp_tlev(ctx, ncol, nlay, play_d, plev_d, tlay_d, tlev_tmp)) return rc;
  }
  const int rc = rrnn_gas_optics_tc(ctx, 0, kd, models, nmodels, ncol, nlay, play_d, plev_d, tlay_d, tlev_d ? tlev_d : tlev_tmp, tsfc_d, gases,
                                    ngas, tau_d, pfrac_d, nullptr, sfc_source_d, sfc_source_Jac_d, K_GAS_LW, planck_lay_d, planck_lev_d);
  if (tlev_tmp) RRNN_CUDA(cudaFreeAsync(tlev_tmp, ctx->stream));
  if (rc < 0) return fail("gas_optics (compact sources): configuration not supported by the tensor-core kernel");
  return rc;
}

extern "C" int rrnn_gas_optics_sw(rrnn_ctx_t* ctx, const rrnn_kdist_t* kd, const rrnn_model_t* const* models, int ncol,
                                  int nlay, const float* play_d, const float* plev_d, const float* tlay_d,
                                  const rrnn_gas_t* gases, int ngas, float* tau_d, float* ssa_d, float* g_d,
                                  float* toa_src_d) {
  rrnn::NvtxRange nvtx_("gas_optics (SW)");
  RRNN_CHECK(ctx && kd && models && models[0], "gas_optics(): null handle");
  RRNN_CHECK(nlay >= 1 && ncol >= 0, "gas_optics(): bad extents");
  RRNN_CHECK(on_device(models[0]) && (!ssa_d || on_device(models[1])), "gas_optics(): network was loaded without a device context");
  if (ncol == 0) return 0;
  RRNN_CUDA(cudaSetDevice(ctx->device));
  GoParams p{};
  if (int rc = map_gases(models[0], gases, ngas, p)) return rc;
  p.fields = 1; p.ncol = ncol; p.nlay = nlay; p.ngpt = kd->ngpt; p.nsamples = (long long)ncol * nlay;
  p.play = play_d; p.plev = plev_d; p.tlay = tlay_d;
  const bool two_stream = ssa_d != nullptr;
  p.nnets = two_stream ? 2 : 1;
  if (two_stream) RRNN_CHECK(models[1], "gas_optics(): Rayleigh network missing");
  for (int n = 0; n < p.nnets; ++n) {
    fill_net(p.net[n], models[n]);
    RRNN_CHECK(models[n]->dims[models[n]->nlayers] == kd->ngpt, "gas_optics(): network output size differs from the number of g-points");
    RRNN_CHECK(models[n]->d_ymean && models[n]->d_ystd, "output_sgemm_tau: NN output scaling coefficients missing");
    RRNN_CHECK(models[n]->dims[0] == models[0]->dims[0], "gas_optics(): the two networks take different inputs");
  }
  RRNN_CHECK(kd->ngpt % 4 == 0 && kd->ngpt <= 256, "gas_optics(): ngpt must be a multiple of 4 and <= 256");
  p.out0 = tau_d; p.out1 = ssa_d; p.out2 = g_d;
  int rc = -1;
  if (two_stream && ctx->nn_tensor_cores)
    rc = rrnn_gas_optics_tc(ctx, 1, kd, models, 2, ncol, nlay, play_d, plev_d, tlay_d, nullptr, nullptr, gases, ngas, tau_d, ssa_d,
                            g_d, nullptr, nullptr, K_GAS_SW);
  if (rc < 0) rc = launch_go<EPI_SW>(ctx, p, kd->ngpt, K_GAS_SW);
  if (rc) return rc;
  if (toa_src_d) {
    RRNN_CHECK(kd->d_solar_source, "gas_optics(): k-distribution has no solar source (not a shortwave k-distribution)");
    const size_t n = (size_t)kd->ngpt * ncol;
    toa_src_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(kd->ngpt, ncol, kd->d_solar_source, toa_src_d);
    RRNN_LAUNCH_CHECK(ctx);
  }
  return 0;
}

static int sgemm_common(rrnn_ctx_t* ctx, const rrnn_model_t* m, int nbatch, const float* x_d, const float* coldry_d, GoParams& p) {
  RRNN_CHECK(ctx && m, "output_sgemm: null handle");
  RRNN_CHECK(on_device(m), "output_sgemm: network was loaded without a device context");
  RRNN_CHECK(m->dims[0] <= MAX_NN_INPUTS, "output_sgemm: too many inputs");
  RRNN_CUDA(cudaSetDevice(ctx->device));
  p.fields = 0; p.nx = m->dims[0]; p.nsamples = nbatch; p.nlay = 1; p.ncol = nbatch; p.ngpt = m->dims[m->nlayers];
  p.x = x_d; p.coldry = coldry_d; p.nnets = 1;
  fill_net(p.net[0], m);
  RRNN_CHECK(p.ngpt % 4 == 0 && p.ngpt <= 256, "output_sgemm: output size must be a multiple of 4 and <= 256");
  return 0;
}

extern "C" int rrnn_output_sgemm_tau(rrnn_ctx_t* ctx, const rrnn_model_t* m, int nbatch, const float* x_d, const float* coldry_d,
                                     float* output_d, float* output2_d) {
  rrnn::NvtxRange nvtx_("compute_tau");
  GoParams p{};
  if (int rc = sgemm_common(ctx, m, nbatch, x_d, coldry_d, p)) return rc;
  RRNN_CHECK(m->d_ymean && m->d_ystd, "output_sgemm_tau: NN output scaling coefficients missing");
  if (nbatch == 0) return 0;
  p.out0 = output_d; p.out1 = output2_d;
  return launch_go<EPI_TAU>(ctx, p, p.ngpt);
}

extern "C" int rrnn_output_sgemm_pfrac(rrnn_ctx_t* ctx, const rrnn_model_t* m, int nbatch, const float* x_d, float* output_d) {
  rrnn::NvtxRange nvtx_("first_sgemm .. last_sgemm");
  GoParams p{};
  if (int rc = sgemm_common(ctx, m, nbatch, x_d, nullptr, p)) return rc;
  if (nbatch == 0) return 0;
  p.out0 = output_d;
  return launch_go<EPI_PFRAC>(ctx, p, p.ngpt);
}

extern "C" int rrnn_output_sgemm_lw(rrnn_ctx_t* ctx, const rrnn_model_t* m, int nbatch, const float* x_d, float* output_d) {
  rrnn::NvtxRange nvtx_("first_sgemm .. last_sgemm");
  GoParams p{};
  if (int rc = sgemm_common(ctx, m, nbatch, x_d, nullptr, p)) return rc;
  if (nbatch == 0) return 0;
  p.out0 = output_d;
  return launch_go<EPI_RAW>(ctx, p, p.ngpt);
}
