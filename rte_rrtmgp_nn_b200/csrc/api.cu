// Context / handle management and the small column-parallel kernels of librrnn_b200.
#include "common.cuh"
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cfloat>

namespace rrnn {

static thread_local std::string g_last_error;
void set_error(const std::string& msg) { g_last_error = msg; }
int fail(const std::string& msg) {
  g_last_error = msg;
  return 1;
}
unsigned long long next_uid() {
  static std::atomic<unsigned long long> counter{1};
  return counter.fetch_add(1);
}

// ---- small kernels -------------------------------------------------------------------------------
// get_col_dry, rrtmgp/mo_gas_optics_rrtmgp.F90:1697-1703
__global__ void col_dry_kernel(int ncol, int nlay, const float* __restrict__ h2o, const float* __restrict__ plev,
                               float* __restrict__ col_dry) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)ncol * nlay) return;
  const size_t col = i / nlay;
  const int lay = (int)(i - col * nlay);
  const float dp = fabsf(plev[col * (nlay + 1) + lay] - plev[col * (nlay + 1) + lay + 1]);
  const float h = h2o[i];
  const float fact = 1.0f / (1.0f + h);
  const float m_air = (0.028964f + 0.018016f * h) * fact;
  col_dry[i] = 10.0f * dp * 6.02214076e23f * fact / (1000.0f * m_air * 100.0f * 9.80665f);
}

// level temperatures, rrtmgp/mo_gas_optics_rrtmgp.F90:326-335
__global__ void interp_tlev_kernel(int ncol, int nlay, const float* __restrict__ play, const float* __restrict__ plev,
                                   const float* __restrict__ tlay, float* __restrict__ tlev) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)ncol * (nlay + 1)) return;
  const size_t col = i / (nlay + 1);
  const int l = (int)(i - col * (nlay + 1));
  const float* pl = play + col * nlay;
  const float* pv = plev + col * (nlay + 1);
  const float* tl = tlay + col * nlay;
  float r;
  if (l == 0) {
    r = tl[0] + (pv[0] - pl[0]) * (tl[1] - tl[0]) / (pl[1] - pl[0]);
  } else if (l == nlay) {
    r = tl[nlay - 1] + (pv[nlay] - pl[nlay - 1]) * (tl[nlay - 1] - tl[nlay - 2]) / (pl[nlay - 1] - pl[nlay - 2]);
  } else {
    r = (pl[l - 1] * tl[l - 1] * (pv[l] - pl[l]) + pl[l] * tl[l] * (pl[l - 1] - pv[l])) / (pv[l] * (pl[l - 1] - pl[l]));
  }
  tlev[i] = r;
}

struct NNInParams {
  int ncol, nlay, nx;
  const float *play, *tlay;
  const float* ptr[MAX_NN_INPUTS];
  float value[MAX_NN_INPUTS];
  int mode[MAX_NN_INPUTS];
  float xmin[MAX_NN_INPUTS], xmax[MAX_NN_INPUTS];
  float* out;
};

// compute_nn_inputs, rrtmgp/mo_gas_optics_rrtmgp.F90:713-782 (materialised form)
__global__ void nn_inputs_kernel(const NNInParams p) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t n = (size_t)p.ncol * p.nlay * p.nx;
  if (i >= n) return;
  const int k = (int)(i % p.nx);
  const size_t smp = i / p.nx;
  const int lay = (int)(smp % p.nlay);
  float raw;
  if (k == 0) raw = p.tlay[smp];
  else if (k == 1) raw = logf(p.play[smp]);
  else {
    if (p.mode[k] == 2) raw = p.ptr[k][smp];
    else if (p.mode[k] == 1) raw = p.ptr[k][lay];
    else if (p.mode[k] == 0) raw = p.value[k];
    else raw = 0.0f;
    if (k == 2 || k == 3) raw = sqrtf(sqrtf(raw));
  }
  p.out[i] = (raw - p.xmin[k]) / (p.xmax[k] - p.xmin[k]);
}

__device__ __forceinline__ float planck_interp1(float T, float tmin, float delta, const float* __restrict__ tab, int ntemp) {
  const float val0 = (T - tmin) / delta;
  const int iv = (int)val0;
  const float frac = val0 - (float)iv;
  int idx = min(ntemp - 1, max(1, iv + 1));
  const float t0 = tab[idx - 1];
  return t0 + frac * (tab[idx] - t0);
}

// compute_Planck_source_nn, rrtmgp/kernels/mo_gas_optics_kernels.F90:615-683 (stand-alone form):
// one thread per (g-point, level index 0..nlay, column); the level index nlay handles the bottom row and the
// surface terms.
__global__ void planck_source_kernel(int ncol, int nlay, int ngpt, int ntemp, const int* __restrict__ gpt2band,
                                     const float* __restrict__ totplnk, float tmin, float delta,
                                     const float* __restrict__ tlay, const float* __restrict__ tlev,
                                     const float* __restrict__ tsfc, int sfc_lay, float* __restrict__ sfc_source,
                                     float* __restrict__ sfc_jac, float* __restrict__ pfrac, float* __restrict__ lev_source) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t n = (size_t)ncol * (nlay + 1) * ngpt;
  if (i >= n) return;
  const int g = (int)(i % ngpt);
  const size_t r = i / ngpt;
  const size_t col = r / (nlay + 1);
  const int lev = (int)(r - col * (nlay + 1));
  const float* tab = totplnk + (size_t)gpt2band[g] * ntemp;
  if (lev == nlay) {
    // sources that read pfrac of other layers are handled by the thread of that layer (below) to keep
    // the in-place update race-free; nothing to do here.
    return;
  }
  const size_t e = (col * nlay + lev) * ngpt + g;
  const float pf = pfrac[e];
  lev_source[i] = pf * planck_interp1(tlev[col * (nlay + 1) + lev], tmin, delta, tab, ntemp);
  if (lev == nlay - 1)
    lev_source[(col * (nlay + 1) + nlay) * ngpt + g] = pf * planck_interp1(tlev[col * (nlay + 1) + nlay], tmin, delta, tab, ntemp);
  if (lev == sfc_lay - 1) {
    const float a = planck_interp1(tsfc[col], tmin, delta, tab, ntemp);
    sfc_source[col * ngpt + g] = pf * a;
    sfc_jac[col * ngpt + g] = pf * (planck_interp1(tsfc[col] + 1.0f, tmin, delta, tab, ntemp) - a);
  }
  pfrac[e] = pf * planck_interp1(tlay[col * nlay + lev], tmin, delta, tab, ntemp);
}

// compute_all_from_table + combine, extensions/cloud_optics/mo_cloud_optics.F90:603-645, 505-528
struct CloudParams {
  int ncol, nlay, nbnd, nliq, nice, two_stream;
  float liq_step, liq_off, ice_step, ice_off;
  const float *extliq, *ssaliq, *asyliq, *extice, *ssaice, *asyice;
  const float *clwp, *ciwp, *reliq, *reice;
  float *tau, *ssa, *g;
};

__device__ __forceinline__ void table3(float wp, float re, float off, float step, int nsteps, const float* __restrict__ e,
                                       const float* __restrict__ s, const float* __restrict__ a, bool mask, float& t,
                                       float& ts, float& tsg) {
  if (!mask) { t = 0.f; ts = 0.f; tsg = 0.f; return; }
  int index = min((int)floorf((re - off) / step) + 1, nsteps - 1);
  const float fint = (re - off) / step - (float)(index - 1);
  t = wp * (e[index - 1] + fint * (e[index] - e[index - 1]));
  ts = t * (s[index - 1] + fint * (s[index] - s[index - 1]));
  tsg = ts * (a[index - 1] + fint * (a[index] - a[index - 1]));
}

__global__ void cloud_optics_kernel(const CloudParams p) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t n = (size_t)p.ncol * p.nlay * p.nbnd;
  if (i >= n) return;
  const int b = (int)(i % p.nbnd);
  const size_t s = i / p.nbnd;
  float lt, lts, ltsg, it, its, itsg;
  const float lw = p.clwp[s], iw = p.ciwp[s];
  table3(lw, p.reliq[s], p.liq_off, p.liq_step, p.nliq, p.extliq + (size_t)b * p.nliq, p.ssaliq + (size_t)b * p.nliq,
         p.asyliq + (size_t)b * p.nliq, lw > 0.f, lt, lts, ltsg);
  table3(iw, p.reice[s], p.ice_off, p.ice_step, p.nice, p.extice + (size_t)b * p.nice, p.ssaice + (size_t)b * p.nice,
         p.asyice + (size_t)b * p.nice, iw > 0.f, it, its, itsg);
  if (!p.two_stream) {
    p.tau[i] = (lt - lts) + (it - its);
  } else {
    const float t = lt + it, ts = lts + its;
    p.g[i] = (ltsg + itsg) / fmaxf(FLT_EPSILON, ts);
    p.ssa[i] = ts / fmaxf(FLT_EPSILON, t);
    p.tau[i] = t;
  }
}

// The all-sky drivers' cloud preparation in ONE pass (LUT cloud optics -> [SW: delta_scale_2str_k] -> the packed solvers' table rows):
// compute_all_from_table + combine (above), rte/kernels/mo_optical_props_kernels.F90:72-93 and cloud_rows_{lw,sw}_kernel
// (rte_solvers.cu) with the by-band arrays kept in registers -- they are never written.  One thread per (sample, 4 bands): the size
// index and its fraction are per sample, the rows go out as 16-byte stores.  Same expressions in the same order as the three
// separate kernels (the stage API still runs those).
template <bool TWO_STREAM>
__global__ void cloud_rows_fused_kernel(const CloudParams p, float* __restrict__ rows) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t nsmp = (size_t)p.ncol * p.nlay;
  if (i >= nsmp * 4) return;
  const size_t s = i >> 2;
  const int q = (int)(i & 3);
  const float lw = p.clwp[s], iw = p.ciwp[s], rl = p.reliq[s], ri = p.reice[s];
  float t2[4], s2[4], sg2[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int b = 4 * q + j;
    t2[j] = 0.0f; s2[j] = 0.0f; sg2[j] = 0.0f;
    if (b >= p.nbnd) continue;
    float lt, lts, ltsg, it, its, itsg;
    table3(lw, rl, p.liq_off, p.liq_step, p.nliq, p.extliq + (size_t)b * p.nliq, p.ssaliq + (size_t)b * p.nliq, p.asyliq + (size_t)b * p.nliq,
           lw > 0.f, lt, lts, ltsg);
    table3(iw, ri, p.ice_off, p.ice_step, p.nice, p.extice + (size_t)b * p.nice, p.ssaice + (size_t)b * p.nice, p.asyice + (size_t)b * p.nice,
           iw > 0.f, it, its, itsg);
    if (!TWO_STREAM) {
      t2[j] = (lt - lts) + (it - its);
    } else {
      const float t = lt + it, ts = lts + its;
      const float gv = (ltsg + itsg) / fmaxf(FLT_EPSILON, ts);
      const float w = ts / fmaxf(FLT_EPSILON, t);
      // delta_scale_2str_k
      const float eps = 3.0f * FLT_MIN;
      const float f = gv * gv;
      const float wf = w * f;
      const float tau = (1.0f - wf) * t;
      const float ssa = (w - wf) / fmaxf(eps, 1.0f - wf);
      const float g = (gv - f) / fmaxf(eps, 1.0f - f);
      t2[j] = tau;
      s2[j] = __fmul_rn(tau, ssa);
      sg2[j] = __fmul_rn(s2[j], g);
    }
  }
  if (!TWO_STREAM) {
    reinterpret_cast<float4*>(rows)[s * 4 + q] = make_float4(t2[0], t2[1], t2[2], t2[3]);
  } else {
    float4* r = reinterpret_cast<float4*>(rows) + s * 12 + q;
    r[0] = make_float4(t2[0], t2[1], t2[2], t2[3]);
    r[4] = make_float4(s2[0], s2[1], s2[2], s2[3]);
    r[8] = make_float4(sg2[0], sg2[1], sg2[2], sg2[3]);
  }
}

// compute_all_from_pade + pade_eval_1 + combine, extensions/cloud_optics/mo_cloud_optics.F90:650-714, 757-781, 500-528:
// orders [2/3] (extinction) and [2/2] (co-albedo, asymmetry), three size regimes (:476-493)
struct PadeParams {
  int ncol, nlay, nbnd, two_stream;
  float sizreg[24];  // bounds of extliq, ssaliq, asyliq, extice, ssaice, asyice (4 each)
  const float *extliq, *ssaliq, *asyliq, *extice, *ssaice, *asyice;
  const float *clwp, *ciwp, *reliq, *reice;
  float *tau, *ssa, *g;
};

template <int M, int N>
__device__ __forceinline__ float pade_eval(int iband, int nbnd, const float* bounds, float re, const float* __restrict__ c) {
  // index into the size-regime table, as written (:683): works for exactly three regimes
  const int irad = min((int)floorf((re - bounds[1]) / bounds[2]) + 2, 3);
  const float* cc = c + (size_t)(irad - 1) * nbnd + iband;
  const size_t st = (size_t)3 * nbnd;  // stride between coefficients
  float denom = cc[(N + M) * st];
#pragma unroll
  for (int i = N - 1 + M; i >= 1 + M; --i) denom = cc[i * st] + re * denom;
  denom = 1.0f + re * denom;
  float numer = cc[M * st];
#pragma unroll
  for (int i = M - 1; i >= 1; --i) numer = cc[i * st] + re * numer;
  numer = cc[0] + re * numer;
  return numer / denom;
}

__device__ __forceinline__ void pade3(float wp, float re, bool mask, int b, int nbnd, const float* bounds, const float* e, const float* s,
                                      const float* a, float& t, float& ts, float& tsg) {
  if (!mask) { t = 0.f; ts = 0.f; tsg = 0.f; return; }
  t = wp * pade_eval<2, 3>(b, nbnd, bounds, re, e);
  ts = t * (1.0f - fmaxf(0.0f, pade_eval<2, 2>(b, nbnd, bounds + 4, re, s)));  // the co-albedo approximant can be negative
  tsg = ts * pade_eval<2, 2>(b, nbnd, bounds + 8, re, a);
}

__global__ void cloud_optics_pade_kernel(const PadeParams p) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t n = (size_t)p.ncol * p.nlay * p.nbnd;
  if (i >= n) return;
  const int b = (int)(i % p.nbnd);
  const size_t s = i / p.nbnd;
  float lt, lts, ltsg, it, its, itsg;
  const float lw = p.clwp[s], iw = p.ciwp[s];
  pade3(lw, p.reliq[s], lw > 0.f, b, p.nbnd, p.sizreg, p.extliq, p.ssaliq, p.asyliq, lt, lts, ltsg);
  pade3(iw, p.reice[s], iw > 0.f, b, p.nbnd, p.sizreg + 12, p.extice, p.ssaice, p.asyice, it, its, itsg);
  if (!p.two_stream) {
    p.tau[i] = (lt - lts) + (it - its);
  } else {
    const float t = lt + it, ts = lts + its;
    p.g[i] = (ltsg + itsg) / fmaxf(FLT_EPSILON, ts);
    p.ssa[i] = ts / fmaxf(FLT_EPSILON, t);
    p.tau[i] = t;
  }
}

// McICA sampling, extensions/cloud_optics/mo_cloud_sampling.F90:107-286 (sampled_mask_max_ran / sampled_mask_exp_ran): one
// thread per (column, g-point) walks the layers; adjacent g-points read adjacent random numbers (coalesced).  The module's
// stale (ncol,nlay,ngpt) declarations are replaced by this fork's (ngpt,nlay,ncol) layout; overlap == null: maximum-random.
__global__ void sampled_mask_kernel(int ngpt, int nlay, int ncol, const float* __restrict__ randoms, const float* __restrict__ cloud_frac,
                                    const float* __restrict__ overlap, unsigned char* __restrict__ mask) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)ngpt * ncol) return;
  const size_t c = i / ngpt;
  const int g = (int)(i - c * ngpt);
  const float* cf = cloud_frac + c * nlay;
  const float* rn = randoms + c * nlay * ngpt + g;
  unsigned char* m = mask + c * nlay * ngpt + g;
  float local = 0.0f;
  bool prev_cloudy = false, started = false;
  for (int l = 0; l < nlay; ++l) {
    const float f = cf[l];
    const bool cloudy = f > 0.0f;
    unsigned char out = 0;
    if (cloudy) {
      const float r = rn[(size_t)l * ngpt];
      if (!started || !prev_cloudy) local = r;                      // first cloudy layer, or the layer above is clear: new deviates
      else if (overlap) {                                           // exponential-random: correlated deviates (:267-274)
        const float rho = overlap[c * (nlay - 1) + (l - 1)];
        local = rho * (local - 0.5f) + sqrtf(1.0f - rho * rho) * (r - 0.5f) + 0.5f;
      }                                                             // maximum-random: keep the deviates (:158)
      out = local > (1.0f - f);
      started = true;
    }
    prev_cloudy = cloudy;
    m[(size_t)l * ngpt] = out;
  }
}

// draw_samples / apply_cloud_mask (:38-101, 292-308) for up to three fields at once
__global__ void apply_cloud_mask_kernel(size_t nsmp, int ngpt, int nbnd, const int* __restrict__ gpt2band, const unsigned char* __restrict__ mask,
                                        const float* __restrict__ i0, const float* __restrict__ i1, const float* __restrict__ i2,
                                        float* __restrict__ o0, float* __restrict__ o1, float* __restrict__ o2) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nsmp * ngpt) return;
  const size_t s = i / ngpt;
  const int b = gpt2band[i - s * ngpt];
  const bool on = mask[i] != 0;
  o0[i] = on ? i0[s * nbnd + b] : 0.0f;
  if (o1) o1[i] = on ? i1[s * nbnd + b] : 0.0f;
  if (o2) o2[i] = on ? i2[s * nbnd + b] : 0.0f;
}

// delta_scale_2str_k, rte/kernels/mo_optical_props_kernels.F90:72-93
__global__ void delta_scale_kernel(size_t n, float* __restrict__ tau, float* __restrict__ ssa, float* __restrict__ g) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float eps = 3.0f * FLT_MIN;
  const float gv = g[i], w = ssa[i];
  const float f = gv * gv;
  const float wf = w * f;
  tau[i] = (1.0f - wf) * tau[i];
  ssa[i] = (w - wf) / fmaxf(eps, 1.0f - wf);
  g[i] = (gv - f) / fmaxf(eps, 1.0f - f);
}

// inc_1scalar_by_1scalar_bybnd :358-378 (4 g-points per thread)
__global__ void inc_1scl_kernel(size_t nsmp, int ngpt, int nbnd, const int* __restrict__ gpt2band, float* __restrict__ tau1,
                                const float* __restrict__ tau2) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nsmp * ngpt) return;
  const int g = (int)(i % ngpt);
  const size_t s = i / ngpt;
  tau1[i] = tau1[i] + tau2[s * nbnd + gpt2band[g]];
}

// inc_2stream_by_2stream_bybnd :453-485; g1 may be null on input (un-materialised zero)
__global__ void inc_2str_kernel(size_t nsmp, int ngpt, int nbnd, const int* __restrict__ gpt2band, float* __restrict__ tau1,
                                float* __restrict__ ssa1, float* __restrict__ g1, int g1_is_zero, const float* __restrict__ tau2,
                                const float* __restrict__ ssa2, const float* __restrict__ g2) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nsmp * ngpt) return;
  const float eps = 3.0f * FLT_MIN;
  const int g = (int)(i % ngpt);
  const size_t s = i / ngpt;
  const size_t j = s * nbnd + gpt2band[g];
  const float t1 = tau1[i], w1 = ssa1[i], gg1 = g1_is_zero ? 0.0f : g1[i];
  const float t2 = tau2[j], w2 = ssa2[j], gg2 = g2[j];
  const float tau12 = t1 + t2;
  const float tauscat12 = t1 * w1 + t2 * w2;
  g1[i] = (t1 * w1 * gg1 + t2 * w2 * gg2) / fmaxf(eps, tauscat12);
  ssa1[i] = tauscat12 / fmaxf(eps, tau12);
  tau1[i] = tau12;
}

// heating rates: mode 0 = compute_heating_rate [K/s] (extensions/mo_heating_rates.F90:48-52),
//                mode 1 = calc_heating_rate [K/day] (rrtmgp_lw_eval_nn_rfmip.F90:624-653)
__global__ void heating_rate_kernel(int ncol, int nlay, int mode, const float* __restrict__ fup, const float* __restrict__ fdn,
                                    const float* __restrict__ plev, float* __restrict__ hr) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)ncol * nlay) return;
  const size_t col = i / nlay;
  const int l = (int)(i - col * nlay);
  const size_t a = col * (nlay + 1) + l;
  if (mode == 0) {
    hr[i] = (fup[a + 1] - fup[a] - fdn[a + 1] + fdn[a]) * 9.80665f / (1004.64f * (plev[a + 1] - plev[a]));
  } else {
    const float scaling = -(24.0f * 3600.0f * 9.80665f / 1004.0f);
    const float dF = (fdn[a + 1] - fup[a + 1]) - (fdn[a] - fup[a]);
    hr[i] = scaling * dF / (plev[a + 1] - plev[a]);
  }
}

static inline unsigned nblk(size_t n, int t = 256) { return (unsigned)((n + t - 1) / t); }

template <typename T>
static int to_device(T** dptr, const T* host, size_t n) {
  *dptr = nullptr;
  if (!host || n == 0) return 0;
  RRNN_CUDA(cudaMalloc((void**)dptr, n * sizeof(T)));
  RRNN_CUDA(cudaMemcpy(*dptr, host, n * sizeof(T), cudaMemcpyHostToDevice));
  return 0;
}

int map_gases_simple(const rrnn_model_t* m, const rrnn_gas_t* gases, int ngas, NNInParams& p);

}  // namespace rrnn

using namespace rrnn;

extern "C" const char* rrnn_last_error(void) { return g_last_error.c_str(); }
extern "C" int rrnn_version(void) { return 100; }
extern "C" int rrnn_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

extern "C" int rrnn_ctx_create(int device, void* stream, rrnn_ctx_t** out) {
  RRNN_CHECK(out, "rrnn_ctx_create: null output pointer");
  *out = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0)
    return fail(std::string("rrnn_ctx_create: no CUDA device available (this library has no CPU fallback): ") +
                (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0"));
  RRNN_CHECK(device >= 0 && device < n, "rrnn_ctx_create: device index out of range");
  RRNN_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  RRNN_CUDA(cudaGetDeviceProperties(&prop, device));
  RRNN_CHECK(prop.major >= 10, "rrnn_ctx_create: this library is built for sm_100a (Blackwell) only");
  rrnn_ctx_t* c = new rrnn_ctx_t();
  c->device = device;
  c->num_sms = prop.multiProcessorCount;
  c->smem_optin = prop.sharedMemPerBlockOptin;
  // NULL = the legacy default stream, which is also PyTorch's default stream: calls are then ordered with
  // whatever the caller enqueued through torch without extra synchronisation.
  c->stream = (cudaStream_t)stream;
  c->own_stream = false;
  if (cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaStreamCreateWithFlags(&c->out_stream, cudaStreamNonBlocking) != cudaSuccess) { delete c; return fail("rrnn_ctx_create: cannot create copy streams"); }
  for (auto& ev : c->ev) cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
  // defaults of the tuning flags can be overridden from the environment (used by the test matrix)
  if (const char* e = getenv("RRNN_FAST_MATH")) c->fast_math = atoi(e) ? 1 : 0;
  if (const char* e = getenv("RRNN_SW_FAST_MATH")) c->sw_fast_math = atoi(e) ? 1 : 0;
  if (const char* e = getenv("RRNN_SOLVER_BUFFER")) c->solver_buffer = atoi(e);
  if (const char* e = getenv("RRNN_NN_TENSOR_CORES")) c->nn_tensor_cores = atoi(e) ? 1 : 0;
  if (const char* e = getenv("RRNN_SOLVER_WIDE")) c->solver_wide = atoi(e) ? 1 : 0;
  if (const char* e = getenv("RRNN_SOLVER_WIDE_SW")) c->solver_wide_sw = atoi(e) ? 1 : 0;
  if (const char* e = getenv("RRNN_SW_WIDE_SCRATCH_MB")) c->solver_scratch_mb_sw_wide = atoi(e);
  *out = c;
  return 0;
}

extern "C" int rrnn_ctx_destroy(rrnn_ctx_t* c) {
  if (!c) return 0;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  if (c->ws) cudaFree(c->ws);
  if (c->scratch) cudaFree(c->scratch);
  if (c->col_counter) cudaFree(c->col_counter);
  if (c->pinned) cudaFreeHost(c->pinned);
  for (auto& ev : c->ev) if (ev) cudaEventDestroy(ev);
  for (auto& v : c->prof_ev) for (auto& pr : v) { cudaEventDestroy(pr.first); cudaEventDestroy(pr.second); }
  if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
  if (c->out_stream) cudaStreamDestroy(c->out_stream);
  if (c->own_stream) cudaStreamDestroy(c->stream);
  delete c;
  return 0;
}

extern "C" int rrnn_ctx_set_stream(rrnn_ctx_t* c, void* stream) {
  RRNN_CHECK(c, "null context");
  if (c->own_stream) { cudaStreamDestroy(c->stream); c->own_stream = false; }
  c->stream = (cudaStream_t)stream;
  return 0;
}
extern "C" void* rrnn_ctx_stream(rrnn_ctx_t* c) { return c ? (void*)c->stream : nullptr; }
extern "C" int rrnn_ctx_synchronize(rrnn_ctx_t* c) {
  RRNN_CHECK(c, "null context");
  RRNN_CUDA(cudaStreamSynchronize(c->stream));
  return 0;
}
extern "C" long long rrnn_ctx_launch_count(rrnn_ctx_t* c) { return c ? c->launches : 0; }
// Device memory for hosts that have no CUDA binding of their own (the Fortran veneer): stream-ordered on the context's stream.
extern "C" int rrnn_dev_malloc(rrnn_ctx_t* c, size_t bytes, void** out) {
  RRNN_CHECK(c && out, "rrnn_dev_malloc: null argument");
  RRNN_CUDA(cudaSetDevice(c->device));
  *out = nullptr;
  if (bytes == 0) return 0;
  RRNN_CUDA(cudaMalloc(out, bytes));
  return 0;
}
extern "C" int rrnn_dev_free(rrnn_ctx_t* c, void* p) {
  RRNN_CHECK(c, "rrnn_dev_free: null context");
  if (!p) return 0;
  RRNN_CUDA(cudaSetDevice(c->device));
  RRNN_CUDA(cudaStreamSynchronize(c->stream));
  RRNN_CUDA(cudaFree(p));
  return 0;
}
extern "C" int rrnn_memcpy_h2d(rrnn_ctx_t* c, void* dst_d, const void* src, size_t bytes) {
  RRNN_CHECK(c && (bytes == 0 || (dst_d && src)), "rrnn_memcpy_h2d: null argument");
  if (bytes == 0) return 0;
  RRNN_CUDA(cudaSetDevice(c->device));
  RRNN_CUDA(cudaMemcpyAsync(dst_d, src, bytes, cudaMemcpyHostToDevice, c->stream));
  RRNN_CUDA(cudaStreamSynchronize(c->stream));   // the caller may reuse src at once
  return 0;
}
extern "C" int rrnn_memcpy_d2h(rrnn_ctx_t* c, void* dst, const void* src_d, size_t bytes) {
  RRNN_CHECK(c && (bytes == 0 || (dst && src_d)), "rrnn_memcpy_d2h: null argument");
  if (bytes == 0) return 0;
  RRNN_CUDA(cudaSetDevice(c->device));
  RRNN_CUDA(cudaMemcpyAsync(dst, src_d, bytes, cudaMemcpyDeviceToHost, c->stream));
  RRNN_CUDA(cudaStreamSynchronize(c->stream));
  return 0;
}
extern "C" int rrnn_ctx_last_nn_kernel(rrnn_ctx_t* c) { return c ? c->last_nn_kernel : 0; }
extern "C" int rrnn_ctx_nn_kernel_counts(rrnn_ctx_t* c, long long* n_tc, long long* n_ffma) {
  RRNN_CHECK(c, "rrnn_ctx_nn_kernel_counts: null context");
  if (n_tc) *n_tc = c->nn_tc_launches;
  if (n_ffma) *n_ffma = c->nn_ffma_launches;
  return 0;
}
extern "C" int rrnn_ctx_set_chunk_columns(rrnn_ctx_t* c, int n) {
  RRNN_CHECK(c && n >= 0, "rrnn_ctx_set_chunk_columns: bad argument");
  c->chunk_columns = n;
  return 0;
}
extern "C" int rrnn_ctx_profile(rrnn_ctx_t* c, int enable) {
  RRNN_CHECK(c, "rrnn_ctx_profile: null context");
  c->profile = enable ? 1 : 0;
  for (auto& u : c->prof_used) u = 0;
  return 0;
}
extern "C" int rrnn_ctx_profile_read(rrnn_ctx_t* c, int kind, double* total_ms, int* nlaunches) {
  RRNN_CHECK(c && kind >= 0 && kind < 4 && total_ms && nlaunches, "rrnn_ctx_profile_read: bad argument");
  RRNN_CUDA(cudaStreamSynchronize(c->stream));
  double tot = 0.0;
  int n = 0;
  for (size_t i = 0; i < c->prof_used[kind]; ++i) {
    float ms = 0.f;
    // a launch that failed between its two records leaves the stop event unrecorded: skip that slot
    if (cudaEventElapsedTime(&ms, c->prof_ev[kind][i].first, c->prof_ev[kind][i].second) != cudaSuccess) { cudaGetLastError(); continue; }
    tot += ms;
    ++n;
  }
  *total_ms = tot;
  *nlaunches = n;
  return 0;
}

extern "C" int rrnn_ctx_set_flag(rrnn_ctx_t* c, const char* name, int value) {
  RRNN_CHECK(c && name, "rrnn_ctx_set_flag: null argument");
  const std::string s(name);
  if (s == "lw_source_bug_compat") c->lw_source_bug_compat = value ? 1 : 0;
  else if (s == "fast_math") c->fast_math = value ? 1 : 0;
  else if (s == "sw_fast_math") c->sw_fast_math = value ? 1 : 0;
  else if (s == "solver_buffer") c->solver_buffer = value;
  else if (s == "nn_tensor_cores") c->nn_tensor_cores = value ? 1 : 0;
  else if (s == "lw_compact_source") c->lw_compact_source = value ? 1 : 0;
  else if (s == "solver_variant") c->solver_variant = value;
  else if (s == "host_copy_threads") c->host_copy_threads = value;
  else if (s == "check_extents") c->check_extents = value ? 1 : 0;
  else if (s == "check_values") c->check_values = value ? 1 : 0;
  else if (s == "solver_scratch_mb") c->solver_scratch_mb = value;
  else if (s == "solver_warps") c->solver_warps = value;
  else if (s == "solver_wide") c->solver_wide = value;
  else if (s == "solver_wide_sw") c->solver_wide_sw = value;
  else if (s == "solver_scratch_mb_sw_wide") c->solver_scratch_mb_sw_wide = value;
  else return fail("rrnn_ctx_set_flag: unknown flag " + s);
  return 0;
}

// ---- models --------------------------------------------------------------------------------------
extern "C" int rrnn_model_create(rrnn_ctx_t* ctx, int nlayers, const int* dims, const float* wpack, const float* bpack,
                                 const int* activations, const float* xmin, const float* xmax, const float* ymean,
                                 const float* ystd, const char* input_names, rrnn_model_t** out) {
  RRNN_CHECK(out && dims && wpack && bpack && activations && xmin && xmax, "rrnn_model_create: null argument");
  RRNN_CHECK(nlayers >= 2 && nlayers <= MAX_LAYERS, "rrnn_model_create: number of layers must be between 2 and 6");
  *out = nullptr;
  // ctx == NULL builds a host-only model (file I/O and inspection without a GPU); compute entry points reject it.
  if (ctx) RRNN_CUDA(cudaSetDevice(ctx->device));
  rrnn_model_t* m = new rrnn_model_t();
  m->nlayers = nlayers;
  m->device = ctx ? ctx->device : -1;
  size_t nw = 0, nb = 0;
  for (int i = 0; i <= nlayers; ++i) {
    m->dims[i] = dims[i];
    if (dims[i] <= 0) { delete m; return fail("rrnn_model_create: non-positive layer size"); }
  }
  for (int l = 0; l < nlayers; ++l) {
    if (activations[l] < 0 || activations[l] > RRNN_ACT_HARD_SIGMOID) { delete m; return fail("rrnn_model_create: unknown activation code"); }
    m->act[l] = activations[l];
    m->w_off[l] = nw; m->b_off[l] = nb;
    nw += (size_t)dims[l] * dims[l + 1];
    nb += dims[l + 1];
  }
  m->wpack.assign(wpack, wpack + nw);
  m->bpack.assign(bpack, bpack + nb);
  const int nx = dims[0], ny = dims[nlayers];
  m->xmin.assign(xmin, xmin + nx);
  m->xmax.assign(xmax, xmax + nx);
  if (ymean) m->ymean.assign(ymean, ymean + ny);
  if (ystd) m->ystd.assign(ystd, ystd + ny);
  for (int i = 0; i < nx; ++i) {
    std::string s;
    if (input_names) {
      s.assign(input_names + 32 * i, strnlen(input_names + 32 * i, 32));
      while (!s.empty() && (s.back() == ' ' || s.back() == '\0')) s.pop_back();
      size_t b = 0;
      while (b < s.size() && s[b] == ' ') ++b;
      s = s.substr(b);
    }
    m->input_names.push_back(s);
  }
  int rc = 0;
  if (ctx) {
    rc = to_device(&m->d_wpack, m->wpack.data(), nw);
    if (!rc) rc = to_device(&m->d_bpack, m->bpack.data(), nb);
    if (!rc && ymean) rc = to_device(&m->d_ymean, m->ymean.data(), (size_t)ny);
    if (!rc && ystd) rc = to_device(&m->d_ystd, m->ystd.data(), (size_t)ny);
  }
  if (rc) { rrnn_model_destroy(m); return rc; }
  *out = m;
  return 0;
}

extern "C" int rrnn_model_destroy(rrnn_model_t* m) {
  if (!m) return 0;
  if (m->device >= 0) cudaSetDevice(m->device);
  cudaFree(m->d_wpack); cudaFree(m->d_bpack); cudaFree(m->d_ymean); cudaFree(m->d_ystd);
  delete m;
  return 0;
}
extern "C" int rrnn_model_nlayers(const rrnn_model_t* m) { return m ? m->nlayers : -1; }
extern "C" int rrnn_model_dims(const rrnn_model_t* m, int* dims_out) {
  RRNN_CHECK(m && dims_out, "rrnn_model_dims: null argument");
  for (int i = 0; i <= m->nlayers; ++i) dims_out[i] = m->dims[i];
  return 0;
}
extern "C" int rrnn_model_input_name(const rrnn_model_t* m, int i, char* buf32) {
  RRNN_CHECK(m && buf32 && i >= 0 && i < m->dims[0], "rrnn_model_input_name: bad argument");
  memset(buf32, 0, 32);
  strncpy(buf32, m->input_names[i].c_str(), 31);
  return 0;
}
extern "C" int rrnn_model_activation(const rrnn_model_t* m, int layer) {
  if (!m || layer < 0 || layer >= m->nlayers) return -1;
  return m->act[layer];
}
extern "C" int rrnn_model_get(const rrnn_model_t* m, int which, int layer, float* data_out, int* n_out) {
  RRNN_CHECK(m && n_out, "rrnn_model_get: null argument");
  const float* src = nullptr;
  size_t n = 0;
  switch (which) {
    case 0: RRNN_CHECK(layer >= 0 && layer < m->nlayers, "rrnn_model_get: bad layer");
      src = m->wpack.data() + m->w_off[layer]; n = (size_t)m->dims[layer] * m->dims[layer + 1]; break;
    case 1: RRNN_CHECK(layer >= 0 && layer < m->nlayers, "rrnn_model_get: bad layer");
      src = m->bpack.data() + m->b_off[layer]; n = m->dims[layer + 1]; break;
    case 2: src = m->xmin.data(); n = m->xmin.size(); break;
    case 3: src = m->xmax.data(); n = m->xmax.size(); break;
    case 4: src = m->ymean.data(); n = m->ymean.size(); break;
    case 5: src = m->ystd.data(); n = m->ystd.size(); break;
    default: return fail("rrnn_model_get: unknown selector");
  }
  *n_out = (int)n;
  if (data_out && n) memcpy(data_out, src, n * sizeof(float));
  return 0;
}

// ---- spectral tables -----------------------------------------------------------------------------
extern "C" int rrnn_kdist_create(rrnn_ctx_t* ctx, int nbnd, int ngpt, const int* band_lims_gpt, int ntemp,
                                 const float* totplnk, float temp_ref_min, float totplnk_delta,
                                 const float* solar_source, rrnn_kdist_t** out) {
  RRNN_CHECK(ctx && out && band_lims_gpt, "rrnn_kdist_create: null argument");
  RRNN_CHECK(nbnd > 0 && nbnd <= MAX_BANDS && ngpt > 0, "rrnn_kdist_create: bad extents");
  *out = nullptr;
  RRNN_CUDA(cudaSetDevice(ctx->device));
  rrnn_kdist_t* k = new rrnn_kdist_t();
  k->nbnd = nbnd; k->ngpt = ngpt; k->ntemp = ntemp; k->temp_ref_min = temp_ref_min; k->totplnk_delta = totplnk_delta;
  k->device = ctx->device;
  k->band_lims_gpt.assign(band_lims_gpt, band_lims_gpt + 2 * nbnd);
  k->gpt2band.assign(ngpt, -1);
  for (int b = 0; b < nbnd; ++b) {
    const int s = band_lims_gpt[2 * b], e = band_lims_gpt[2 * b + 1];
    if (s < 1 || e > ngpt || s > e) { delete k; return fail("rrnn_kdist_create: band_lims_gpt out of range"); }
    for (int g = s - 1; g <= e - 1; ++g) k->gpt2band[g] = b;
  }
  for (int g = 0; g < ngpt; ++g)
    if (k->gpt2band[g] < 0) { delete k; return fail("rrnn_kdist_create: g-points not covered by the bands"); }
  if (totplnk) {
    if (ntemp < 2) { delete k; return fail("rrnn_kdist_create: Planck table needs at least 2 temperatures"); }
    k->totplnk.assign(totplnk, totplnk + (size_t)nbnd * ntemp);
  }
  if (solar_source) k->solar_source.assign(solar_source, solar_source + ngpt);
  int rc = to_device(&k->d_band_lims_gpt, k->band_lims_gpt.data(), (size_t)2 * nbnd);
  if (!rc) rc = to_device(&k->d_gpt2band, k->gpt2band.data(), (size_t)ngpt);
  if (!rc && totplnk) rc = to_device(&k->d_totplnk, k->totplnk.data(), k->totplnk.size());
  if (!rc && solar_source) rc = to_device(&k->d_solar_source, k->solar_source.data(), (size_t)ngpt);
  if (rc) { rrnn_kdist_destroy(k); return rc; }
  *out = k;
  return 0;
}
extern "C" int rrnn_kdist_destroy(rrnn_kdist_t* k) {
  if (!k) return 0;
  cudaSetDevice(k->device);
  cudaFree(k->d_band_lims_gpt); cudaFree(k->d_gpt2band); cudaFree(k->d_totplnk); cudaFree(k->d_solar_source); cudaFree(k->d_optimal_angle_fit);
  delete k;
  return 0;
}
extern "C" int rrnn_kdist_set_tsi(rrnn_kdist_t* k, float tsi) {
  RRNN_CHECK(k, "set_tsi: null handle");
  RRNN_CHECK(tsi >= 0.f, "tsi out of range");
  RRNN_CHECK(!k->solar_source.empty(), "set_tsi: no solar source");
  float sum = 0.f;
  for (float v : k->solar_source) sum += v;
  const float norm = 1.0f / sum;
  for (float& v : k->solar_source) v = v * tsi * norm;
  RRNN_CUDA(cudaSetDevice(k->device));
  RRNN_CUDA(cudaMemcpy(k->d_solar_source, k->solar_source.data(), k->solar_source.size() * sizeof(float), cudaMemcpyHostToDevice));
  return 0;
}

// ---- building blocks -----------------------------------------------------------------------------
extern "C" int rrnn_get_col_dry(rrnn_ctx_t* ctx, int ncol, int nlay, const float* vmr_h2o_d, const float* plev_d, float* col_dry_d) {
  RRNN_CHECK(ctx, "get_col_dry: null context");
  if (ncol <= 0) return 0;
  RRNN_CUDA(cudaSetDevice(ctx->device));
  const size_t n = (size_t)ncol * nlay;
  col_dry_kernel<<<nblk(n), 256, 0, ctx->stream>>>(ncol, nlay, vmr_h2o_d, plev_d, col_dry_d);
  RRNN_LAUNCH_CHECK(ctx);
  return 0;
}

extern "C" int rrnn_interp_tlev(rrnn_ctx_t* ctx, int ncol, int nlay, const float* play_d, const float* plev_d,
                                const float* tlay_d, float* tlev_d) {
  RRNN_CHECK(ctx, "interp_tlev: null context");
  RRNN_CHECK(nlay >= 2, "gas_optics(): level-temperature interpolation needs at least two layers");
  if (ncol <= 0) return 0;
  RRNN_CUDA(cudaSetDevice(ctx->device));
  const size_t n = (size_t)ncol * (nlay + 1);
  interp_tlev_kernel<<<nblk(n), 256, 0, ctx->stream>>>(ncol, nlay, play_d, plev_d, tlay_d, tlev_d);
  RRNN_LAUNCH_CHECK(ctx);
  return 0;
}

extern "C" int rrnn_compute_nn_inputs(rrnn_ctx_t* ctx, const rrnn_model_t* m, int ncol, int nlay, const float* play_d,
                                      const float* tlay_d, const rrnn_gas_t* gases, int ngas, float* nn_inputs_d) {
  RRNN_CHECK(ctx && m, "compute_nn_inputs: null handle");
  const int nx = m->dims[0];
  RRNN_CHECK(nx <= MAX_NN_INPUTS && nx >= 4, "compute_nn_inputs: unsupported number of inputs");
  if (ncol <= 0) return 0;
  RRNN_CUDA(cudaSetDevice(ctx->device));
  NNInParams p{};
  p.ncol = ncol; p.nlay = nlay; p.nx = nx; p.play = play_d; p.tlay = tlay_d; p.out = nn_inputs_d;
  for (int i = 0; i < nx; ++i) {
    p.xmin[i] = m->xmin[i]; p.xmax[i] = m->xmax[i]; p.mode[i] = -1; p.ptr[i] = nullptr; p.value[i] = 0.f;
    if (i < 2) continue;
    for (int g = 0; g < ngas; ++g) {
      std::string nm(gases[g].name, strnlen(gases[g].name, 32));
      while (!nm.empty() && (nm.back() == ' ')) nm.pop_back();
      if (nm == m->input_names[i]) { p.mode[i] = gases[g].ndims; p.ptr[i] = gases[g].conc; p.value[i] = gases[g].value; break; }
    }
    if (i < 4 && p.mode[i] < 0) return fail("compute_nn_inputs: gas " + m->input_names[i] + " is required but was not provided");
  }
  const size_t n = (size_t)ncol * nlay * nx;
  nn_inputs_kernel<<<nblk(n), 256, 0, ctx->stream>>>(p);
  RRNN_LAUNCH_CHECK(ctx);
  return 0;
}

extern "C" int rrnn_planck_source_nn(rrnn_ctx_t* ctx, const rrnn_kdist_t* kd, int ncol, int nlay, const float* tlay_d,
                                     const float* tlev_d, const float* tsfc_d, int sfc_lay, float* sfc_source_d,
                                     float* sfc_source_Jac_d, float* pfrac_lay_source_d, float* lev_source_d) {
  rrnn::NvtxRange nvtx_("compute_Planck_source_nn");
  RRNN_CHECK(ctx && kd && kd->d_totplnk, "compute_Planck_source_nn: null handle or no Planck table");
  RRNN_CHECK(sfc_lay >= 1 && sfc_lay <= nlay, "compute_Planck_source_nn: sfc_lay out of range");
  if (ncol <= 0) return 0;
  RRNN_CUDA(cudaSetDevice(ctx->device));
  const size_t n = (size_t)ncol * (nlay + 1) * kd->ngpt;
  planck_source_kernel<<<nblk(n), 256, 0, ctx->stream>>>(ncol, nlay, kd->ngpt, kd->ntemp, kd->d_gpt2band, kd->d_totplnk,
                                                        kd->temp_ref_min, kd->totplnk_delta, tlay_d, tlev_d, tsfc_d, sfc_lay,
                                                        sfc_source_d, sfc_source_Jac_d, pfrac_lay_source_d, lev_source_d);
  RRNN_LAUNCH_CHECK(ctx);
  return 0;
}

// ---- cloud optics, increments, heating rates -------------------------------------------------------
extern "C" int rrnn_cloud_lut_create(rrnn_ctx_t* ctx, int nbnd, int nsize_liq, int nsize_ice, float radliq_lwr, float radliq_upr,
                                     float radice_lwr, float radice_upr, const float* lut_extliq, const float* lut_ssaliq,
                                     const float* lut_asyliq, const float* lut_extice, const float* lut_ssaice,
                                     const float* lut_asyice, rrnn_cloud_lut_t** out) {
  RRNN_CHECK(ctx && out && lut_extliq && lut_ssaliq && lut_asyliq && lut_extice && lut_ssaice && lut_asyice, "cloud_optics%init(): null argument");
  RRNN_CHECK(nbnd > 0 && nsize_liq > 1 && nsize_ice > 1, "cloud_optics%init(): bad table extents");
  *out = nullptr;
  RRNN_CUDA(cudaSetDevice(ctx->device));
  rrnn_cloud_lut_t* l = new rrnn_cloud_lut_t();
  l->nbnd = nbnd; l->nsize_liq = nsize_liq; l->nsize_ice = nsize_ice;
  l->radliq_lwr = radliq_lwr; l->radice_lwr = radice_lwr;
  l->liq_step = (radliq_upr - radliq_lwr) / (float)(nsize_liq - 1);  // mo_cloud_optics.F90:138-141
  l->ice_step = (radice_upr - radice_lwr) / (float)(nsize_ice - 1);
  const size_t nl = (size_t)nbnd * nsize_liq, ni = (size_t)nbnd * nsize_ice;
  std::vector<float> pack;
  const float* srcs[6] = {lut_extliq, lut_ssaliq, lut_asyliq, lut_extice, lut_ssaice, lut_asyice};
  for (int i = 0; i < 6; ++i) {
    l->off[i] = pack.size();
    pack.insert(pack.end(), srcs[i], srcs[i] + (i < 3 ? nl : ni));
  }
  int rc = to_device(&l->d_tables, pack.data(), pack.size());
  if (rc) { delete l; return rc; }
  *out = l;
  return 0;
}
// load_pade, extensions/cloud_optics/mo_cloud_optics.F90:178-262: coefficient arrays (ncoeff, nsizereg, nbnd) as in the files
// (ice: one roughness), six arrays of nbound = 4 size-regime bounds
extern "C" int rrnn_cloud_pade_create(rrnn_ctx_t* ctx, int nbnd, int nsizereg, int ncoeff_ext, int ncoeff_ssa_g, int nbound,
                                      const float* pade_extliq, const float* pade_ssaliq, const float* pade_asyliq,
                                      const float* pade_extice, const float* pade_ssaice, const float* pade_asyice,
                                      const float* sizreg_extliq, const float* sizreg_ssaliq, const float* sizreg_asyliq,
                                      const float* sizreg_extice, const float* sizreg_ssaice, const float* sizreg_asyice,
                                      rrnn_cloud_lut_t** out) {
  RRNN_CHECK(ctx && out && pade_extliq && pade_ssaliq && pade_asyliq && pade_extice && pade_ssaice && pade_asyice && sizreg_extliq &&
             sizreg_ssaliq && sizreg_asyliq && sizreg_extice && sizreg_ssaice && sizreg_asyice, "cloud_optics%init(): null argument");
  *out = nullptr;
  RRNN_CHECK(nbound == nsizereg + 1, "cloud_optics%init(): one or more Pade size regime arrays are inconsistently sized");
  RRNN_CHECK(nsizereg == 3, "cloud_optics%init(): Expecting precisely three size regimes for Pade approximants");
  RRNN_CHECK(ncoeff_ext == 6 && ncoeff_ssa_g == 5, "cloud optics: code assumes Pade orders [2/3] and [2/2] but data is otherwise");
  RRNN_CHECK(nbnd > 0, "cloud_optics%init(): number of bands inconsistent between lookup tables, spectral discretization");
  const float* sz[6] = {sizreg_extliq, sizreg_ssaliq, sizreg_asyliq, sizreg_extice, sizreg_ssaice, sizreg_asyice};
  // :247-256
  RRNN_CHECK(!(sz[1][0] < sz[0][0] || sz[2][0] < sz[0][0]), "cloud_optics%init(): one or more Pade size regimes have inconsistent lowest values");
  RRNN_CHECK(!(sz[4][0] < sz[3][0] || sz[5][0] < sz[3][0]), "cloud_optics%init(): one or more Pade size regimes have inconsistent lower values");
  RRNN_CHECK(!(sz[1][3] > sz[0][3] || sz[2][3] > sz[0][3]), "cloud_optics%init(): one or more Pade size regimes have lowest value less than radliq_upr");
  RRNN_CHECK(!(sz[4][3] > sz[3][3] || sz[5][3] > sz[3][3]), "cloud_optics%init(): one or more Pade size regimes have lowest value less than radice_upr");
  RRNN_CUDA(cudaSetDevice(ctx->device));
  rrnn_cloud_lut_t* l = new rrnn_cloud_lut_t();
  l->nbnd = nbnd; l->is_pade = 1;
  l->radliq_lwr = sz[0][0]; l->radice_lwr = sz[3][0];
  for (int a = 0; a < 6; ++a)
    for (int i = 0; i < 4; ++i) l->sizreg[4 * a + i] = sz[a][i];
  std::vector<float> pack;
  const float* srcs[6] = {pade_extliq, pade_ssaliq, pade_asyliq, pade_extice, pade_ssaice, pade_asyice};
  for (int i = 0; i < 6; ++i) {
    l->off[i] = pack.size();
    const size_t n = (size_t)((i % 3 == 0) ? ncoeff_ext : ncoeff_ssa_g) * nsizereg * nbnd;
    pack.insert(pack.end(), srcs[i], srcs[i] + n);
  }
  int rc = to_device(&l->d_tables, pack.data(), pack.size());
  if (rc) { delete l; return rc; }
  *out = l;
  return 0;
}
extern "C" int rrnn_cloud_lut_destroy(rrnn_cloud_lut_t* l) {
  if (!l) return 0;
  cudaFree(l->d_tables);
  delete l;
  return 0;
}

extern "C" int rrnn_cloud_optics(rrnn_ctx_t* ctx, const rrnn_cloud_lut_t* lut, int ncol, int nlay, const float* clwp_d,
                                 const float* ciwp_d, const float* reliq_d, const float* reice_d, float* tau_d, float* ssa_d,
                                 float* g_d) {
  rrnn::NvtxRange nvtx_("cloud_optics");
  RRNN_CHECK(ctx && lut, "cloud optics: no data has been initialized");
  RRNN_CHECK((ssa_d == nullptr) == (g_d == nullptr), "cloud optics: ssa and g must both be given or both be absent");
  if (ncol <= 0) return 0;
  RRNN_CUDA(cudaSetDevice(ctx->device));
  if (lut->is_pade) {
    PadeParams q{};
    q.ncol = ncol; q.nlay = nlay; q.nbnd = lut->nbnd; q.two_stream = ssa_d ? 1 : 0;
    for (int i = 0; i < 24; ++i) q.sizreg[i] = lut->sizreg[i];
    q.extliq = lut->d_tables + lut->off[0]; q.ssaliq = lut->d_tables + lut->off[1]; q.asyliq = lut->d_tables + lut->off[2];
    q.extice = lut->d_tables + lut->off[3]; q.ssaice = lut->d_tables + lut->off[4]; q.asyice = lut->d_tables + lut->off[5];
    q.clwp = clwp_d; q.ciwp = ciwp_d; q.reliq = reliq_d; q.reice = reice_d; q.tau = tau_d; q.ssa = ssa_d; q.g = g_d;
    const size_t nq = (size_t)ncol * nlay * lut->nbnd;
    cloud_optics_pade_kernel<<<nblk(nq), 256, 0, ctx->stream>>>(q);
    RRNN_LAUNCH_CHECK(ctx);
    return 0;
  }
  CloudParams p{};
  p.ncol = ncol; p.nlay = nlay; p.nbnd = lut->nbnd; p.nliq = lut->nsize_liq; p.nice = lut->nsize_ice;
  p.two_stream = ssa_d ? 1 : 0;
  p.liq_step = lut->liq_step; p.liq_off = lut->radliq_lwr; p.ice_step = lut->ice_step; p.ice_off = lut->radice_lwr;
  p.extliq = lut->d_tables + lut->off[0]; p.ssaliq = lut->d_tables + lut->off[1]; p.asyliq = lut->d_tables + lut->off[2];
  p.extice = lut->d_tables + lut->off[3]; p.ssaice = lut->d_tables + lut->off[4]; p.asyice = lut->d_tables + lut->off[5];
  p.clwp = clwp_d; p.ciwp = ciwp_d; p.reliq = reliq_d; p.reice = reice_d; p.tau = tau_d; p.ssa = ssa_d; p.g = g_d;
  const size_t n = (size_t)ncol * nlay * lut->nbnd;
  cloud_optics_kernel<<<nblk(n), 256, 0, ctx->stream>>>(p);
  RRNN_LAUNCH_CHECK(ctx);
  return 0;
}

// see cloud_rows_fused_kernel: rows_d holds 16 (LW: tau) or 48 (SW: t2 | s2 | sg2 after delta-scaling) floats per (layer, column);
// -1 (no message): Pade coefficients or more than 16 bands -- the caller runs the three separate kernels
namespace rrnn {
int cloud_rows_fused(rrnn_ctx_t* ctx, const rrnn_cloud_lut_t* lut, int ncol, int nlay, const float* clwp_d, const float* ciwp_d,
                     const float* reliq_d, const float* reice_d, bool two_stream, float* rows_d) {
  if (!lut || lut->is_pade || lut->nbnd > 16 || ((uintptr_t)rows_d & 15)) return -1;
  if (ncol <= 0) return 0;
  NvtxRange nvtx_("cloud_optics");
  CloudParams p{};
  p.ncol = ncol; p.nlay = nlay; p.nbnd = lut->nbnd; p.nliq = lut->nsize_liq; p.nice = lut->nsize_ice;
  p.two_stream = two_stream ? 1 : 0;
  p.liq_step = lut->liq_step; p.liq_off = lut->radliq_lwr; p.ice_step = lut->ice_step; p.ice_off = lut->radice_lwr;
  p.extliq = lut->d_tables + lut->off[0]; p.ssaliq = lut->d_tables + lut->off[1]; p.asyliq = lut->d_tables + lut->off[2];
  p.extice = lut->d_tables + lut->off[3]; p.ssaice = lut->d_tables + lut->off[4]; p.asyice = lut->d_tables + lut->off[5];
  p.clwp = clwp_d; p.ciwp = ciwp_d; p.reliq = reliq_d; p.reice = reice_d;
  const size_t n = (size_t)ncol * nlay * 4;
  if (two_stream) cloud_rows_fused_kernel<true><<<nblk(n), 256, 0, ctx->stream>>>(p, rows_d);
  else cloud_rows_fused_kernel<false><<<nblk(n), 256, 0, ctx->stream>>>(p, rows_d);
  RRNN_LAUNCH_CHECK(ctx);
  return 0;
}
}  // namespace rrnn

extern "C" int rrnn_sampled_mask(rrnn_ctx_t* ctx, int ngpt, int nlay, int ncol, const float* randoms_d, const float* cloud_frac_d,
                                 const float* overlap_param_d, unsigned char* cloud_mask_d) {
  RRNN_CHECK(ctx && randoms_d && cloud_frac_d && cloud_mask_d, "sampled_mask_max_ran: null argument");
  RRNN_CHECK(ngpt > 0 && nlay > 0 && ncol >= 0, "sampled_mask_max_ran: sizes of randoms(ngpt,nlay,ncol) and cloud_frac(ncol,nlay) are inconsistent");
  if (ncol == 0) return 0;
  RRNN_CUDA(cudaSetDevice(ctx->device));
  const size_t n = (size_t)ngpt * ncol;
  sampled_mask_kernel<<<nblk(n), 256, 0, ctx->stream>>>(ngpt, nlay, ncol, randoms_d, cloud_frac_d, overlap_param_d, cloud_mask_d);
  RRNN_LAUNCH_CHECK(ctx);
  return 0;
}

extern "C" int rrnn_draw_samples(rrnn_ctx_t* ctx, const rrnn_kdist_t* kd, int nlay, int ncol, const unsigned char* cloud_mask_d,
                                 const float* tau_bnd_d, const float* ssa_bnd_d, const float* g_bnd_d, float* tau_gpt_d, float* ssa_gpt_d,
                                 float* g_gpt_d) {
  RRNN_CHECK(ctx && kd, "draw_samples: cloud optical properties are not initialized");
  RRNN_CHECK(cloud_mask_d && tau_bnd_d && tau_gpt_d, "draw_samples: sampled cloud optical properties are not initialized");
  RRNN_CHECK((ssa_bnd_d == nullptr) == (ssa_gpt_d == nullptr) && (g_bnd_d == nullptr) == (g_gpt_d == nullptr) &&
                 (ssa_bnd_d == nullptr) == (g_bnd_d == nullptr),
             "draw_samples: by-band and sampled cloud properties need to be the same variable type");
  if (ncol <= 0) return 0;
  RRNN_CUDA(cudaSetDevice(ctx->device));
  const size_t nsmp = (size_t)ncol * nlay;
  apply_cloud_mask_kernel<<<nblk(nsmp * kd->ngpt), 256, 0, ctx->stream>>>(nsmp, kd->ngpt, kd->nbnd, kd->d_gpt2band, cloud_mask_d, tau_bnd_d,
                                                                         ssa_bnd_d, g_bnd_d, tau_gpt_d, ssa_gpt_d, g_gpt_d);
  RRNN_LAUNCH_CHECK(ctx);
  return 0;
}

extern "C" int rrnn_delta_scale_2str(rrnn_ctx_t* ctx, size_t n, float* tau_d, float* ssa_d, float* g_d) {
  rrnn::NvtxRange nvtx_("clouds_deltascale_increment");
  RRNN_CHECK(ctx, "delta_scale: null context");
  if (n == 0) return 0;
  RRNN_CUDA(cudaSetDevice(ctx->device));
  delta_scale_kernel<<<nblk(n), 256, 0, ctx->stream>>>(n, tau_d, ssa_d, g_d);
  RRNN_LAUNCH_CHECK(ctx);
  return 0;
}

extern "C" int rrnn_increment_1scl_bybnd(rrnn_ctx_t* ctx, const rrnn_kdist_t* kd, int nlay, int ncol, float* tau1_d, const float* tau2_d) {
  rrnn::NvtxRange nvtx_("clouds_increment");
  RRNN_CHECK(ctx && kd, "increment: null handle");
  if (ncol <= 0) return 0;
  RRNN_CUDA(cudaSetDevice(ctx->device));
  const size_t nsmp = (size_t)ncol * nlay;
  inc_1scl_kernel<<<nblk(nsmp * kd->ngpt), 256, 0, ctx->stream>>>(nsmp, kd->ngpt, kd->nbnd, kd->d_gpt2band, tau1_d, tau2_d);
  RRNN_LAUNCH_CHECK(ctx);
  return 0;
}

extern "C" int rrnn_increment_2str_bybnd(rrnn_ctx_t* ctx, const rrnn_kdist_t* kd, int nlay, int ncol, float* tau1_d, float* ssa1_d,
                                         float* g1_d, const float* tau2_d, const float* ssa2_d, const float* g2_d) {
  rrnn::NvtxRange nvtx_("clouds_deltascale_increment");
  RRNN_CHECK(ctx && kd && g1_d, "increment: null handle");
  if (ncol <= 0) return 0;
  RRNN_CUDA(cudaSetDevice(ctx->device));
  const size_t nsmp = (size_t)ncol * nlay;
  inc_2str_kernel<<<nblk(nsmp * kd->ngpt), 256, 0, ctx->stream>>>(nsmp, kd->ngpt, kd->nbnd, kd->d_gpt2band, tau1_d, ssa1_d, g1_d, 0,
                                                                 tau2_d, ssa2_d, g2_d);
  RRNN_LAUNCH_CHECK(ctx);
  return 0;
}

static int heating(rrnn_ctx_t* ctx, int mode, int ncol, int nlay, const float* fu, const float* fd, const float* plev, float* hr) {
  RRNN_CHECK(ctx, "heating_rate: null context");
  if (ncol <= 0) return 0;
  RRNN_CUDA(cudaSetDevice(ctx->device));
  heating_rate_kernel<<<nblk((size_t)ncol * nlay), 256, 0, ctx->stream>>>(ncol, nlay, mode, fu, fd, plev, hr);
  RRNN_LAUNCH_CHECK(ctx);
  return 0;
}
extern "C" int rrnn_heating_rate(rrnn_ctx_t* ctx, int ncol, int nlay, const float* flux_up_d, const float* flux_dn_d,
                                 const float* plev_d, float* heating_rate_d) {
  return heating(ctx, 0, ncol, nlay, flux_up_d, flux_dn_d, plev_d, heating_rate_d);
}
extern "C" int rrnn_calc_heating_rate(rrnn_ctx_t* ctx, int ncol, int nlay, const float* flux_up_d, const float* flux_dn_d,
                                      const float* plev_d, float* hr_K_day_d) {
  return heating(ctx, 1, ncol, nlay, flux_up_d, flux_dn_d, plev_d, hr_K_day_d);
}
