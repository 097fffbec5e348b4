/*
 * oracle.c -- CPU restatement of the RTE+RRTMGP-NN hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This file is the parity oracle (and the reported CPU baseline) for the CUDA path in
 * rte_rrtmgp_nn_b200/csrc.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load it; the product never does.
 *
 * PARITY PINNED IN PART.  Pinned -- against outputs of the reference's OWN PYTHON, imported unmodified and run in the build
 * container on the reference's RFMIP profiles (tools/make_ref_python_golden.py -> tests/golden/ref_python_golden.npz;
 * tests/test_oracle_cpu.py::test_oracle_pinned_by_the_references_own_python): orc_get_col_dry, orc_compute_nn_inputs, the
 * (ystd z + ymean)^8 N_dry output transform of orc_output_sgemm_tau (network outputs z from a float64 numpy evaluation of the
 * shipped weights: the reference evaluates them with Keras, absent here), orc_calc_heating_rate, and the scaling constants
 * read from the 2018 weight files.  PARITY UNPINNED for the rest -- Planck sources, rte_lw / rte_sw and their solvers, cloud
 * optics: that part of the reference exists only in Fortran, which cannot be compiled in this image (no Fortran compiler,
 * no netCDF, k-distribution files missing -- SURVEY.md section 0 F1/F2/F4), and the reference's own tests hold no
 * golden vector for it.  Every function below restates
 * the reference loops line by line in fp32 (wp = sp, rte/mo_rte_kind.F90:29-33) and cites the
 * file:line it follows under /root/reference.
 *
 * Array layout is the reference's: g-point fastest, then layer, then column --
 * Fortran (ngpt,nlay,ncol) == C [ncol][nlay][ngpt]; profiles (nlay,ncol) == C [ncol][nlay].
 * Layer/level indices below are 0-based; Fortran index i <-> C index i-1.
 */
#include <math.h>
#include <float.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_API __attribute__((visibility("default")))

/* -DORC_DOUBLE builds the SAME algorithm in double precision (liboracle_f64.so): the rounding-free yardstick
 * used by the tests to separate implementation differences from the fp32 rounding noise that the reference
 * arithmetic itself carries.  Algorithmic constants that the reference takes from the single-precision kind
 * (epsilon(1._sp), tiny(1._sp), k_min, tau_thresh) keep their fp32 values. */
#ifdef ORC_DOUBLE
#define float double
#define expf exp
#define logf log
#define sqrtf sqrt
#define fabsf fabs
#define fmaxf fmax
#define fminf fmin
#define floorf floor
#define acosf acos
#endif

/* rrtmgp/mo_rrtmgp_constants.F90:33-53 */
static const float M_H2O = 0.018016f;
static const float AVOGAD = 6.02214076e23f;
static const float M_DRY = 0.028964f;
static const float GRAV = 9.80665f;
static const float CP_DRY = 1004.64f;

/* activation codes (neural/mod_layer.F90:64-95) */
enum { ACT_LINEAR = 0, ACT_SOFTSIGN = 1, ACT_RELU = 2, ACT_SIGMOID = 3, ACT_HARD_SIGMOID = 4 };

ORC_API int orc_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* ------------------------------------------------------------------------------------------
 * get_col_dry: rrtmgp/mo_gas_optics_rrtmgp.F90:1662-1707 (latitude absent -> g0 = grav)
 * ------------------------------------------------------------------------------------------ */
ORC_API void orc_get_col_dry(int ncol, int nlay, const float* vmr_h2o, const float* plev, float* col_dry) {
  for (int icol = 0; icol < ncol; ++icol) {
    const float g0 = GRAV;
    for (int ilev = 0; ilev < nlay; ++ilev) {
      float delta_plev = fabsf(plev[icol * (nlay + 1) + ilev] - plev[icol * (nlay + 1) + ilev + 1]);
      float h = vmr_h2o[icol * nlay + ilev];
      float fact = 1.0f / (1.0f + h);
      float m_air = (M_DRY + M_H2O * h) * fact;
      col_dry[icol * nlay + ilev] = 10.0f * delta_plev * AVOGAD * fact / (1000.0f * m_air * 100.0f * g0);
    }
  }
}

/* ------------------------------------------------------------------------------------------
 * tlev interpolation when the caller gives no tlev: rrtmgp/mo_gas_optics_rrtmgp.F90:326-335
 * ------------------------------------------------------------------------------------------ */
ORC_API void orc_interp_tlev(int ncol, int nlay, const float* play, const float* plev, const float* tlay,
                             float* tlev) {
  for (int icol = 0; icol < ncol; ++icol) {
    const float* pl = play + (size_t)icol * nlay;
    const float* pv = plev + (size_t)icol * (nlay + 1);
    const float* tl = tlay + (size_t)icol * nlay;
    float* tv = tlev + (size_t)icol * (nlay + 1);
    tv[0] = tl[0] + (pv[0] - pl[0]) * (tl[1] - tl[0]) / (pl[1] - pl[0]);
    for (int i = 1; i < nlay; ++i) {
      tv[i] = (pl[i - 1] * tl[i - 1] * (pv[i] - pl[i]) + pl[i] * tl[i] * (pl[i - 1] - pv[i])) /
              (pv[i] * (pl[i - 1] - pl[i]));
    }
    tv[nlay] = tl[nlay - 1] + (pv[nlay] - pl[nlay - 1]) * (tl[nlay - 1] - tl[nlay - 2]) / (pl[nlay - 1] - pl[nlay - 2]);
  }
}

/* ------------------------------------------------------------------------------------------
 * compute_nn_inputs: rrtmgp/mo_gas_optics_rrtmgp.F90:618-798
 *   inputs 0,1 = tlay, log(play); 2,3 = h2o**(1/4), o3**(1/4) via sqrt(sqrt()) (:716-719);
 *   inputs 4.. = other gases, matched by name by the caller:
 *     gas_mode[i] = 0 scalar (:732-738), 1 per-layer profile (:739-745), 2 full (nlay,ncol) (:746-753),
 *                  -1 gas missing -> ref_vmr = 0 (nn_scenario_index = 0, :636, :757-759).
 *   gas_ptr[i] points at the concentration data (ignored for mode -1).  Entries 2 and 3 must be mode 2.
 * ------------------------------------------------------------------------------------------ */
ORC_API void orc_compute_nn_inputs(int ncol, int nlay, int nx, const float* play, const float* tlay,
                                   const float* const* gas_ptr, const int* gas_mode, const float* xmin,
                                   const float* xmax, float* nn_inputs) {
  for (int icol = 0; icol < ncol; ++icol) {
    for (int ilay = 0; ilay < nlay; ++ilay) {
      size_t s = (size_t)icol * nlay + ilay;
      float* x = nn_inputs + s * nx;
      x[0] = (tlay[s] - xmin[0]) / (xmax[0] - xmin[0]);
      x[1] = (logf(play[s]) - xmin[1]) / (xmax[1] - xmin[1]);
      x[2] = (sqrtf(sqrtf(gas_ptr[2][s])) - xmin[2]) / (xmax[2] - xmin[2]);
      x[3] = (sqrtf(sqrtf(gas_ptr[3][s])) - xmin[3]) / (xmax[3] - xmin[3]);
      for (int ig = 4; ig < nx; ++ig) {
        float v;
        switch (gas_mode[ig]) {
          case 0: v = gas_ptr[ig][0]; break;
          case 1: v = gas_ptr[ig][ilay]; break;
          case 2: v = gas_ptr[ig][s]; break;
          default: v = 0.0f; break;
        }
        x[ig] = (v - xmin[ig]) / (xmax[ig] - xmin[ig]);
      }
    }
  }
}

/* ------------------------------------------------------------------------------------------
 * The MLP.  Weights are given exactly as stored in the netCDF file: layer n is a row-major
 * (n_in, n_out) array == the reference's column-major w_transposed(n_out, n_in)
 * (neural/mod_network_rrtmgp.F90:95-97).  The reference calls BLAS sgemm
 * (mod_network_rrtmgp.F90:166,181,203); BLAS leaves the summation order unspecified, so the
 * oracle uses the plain sequential order over the input index in fp32.
 * bias_and_activation: neural/mod_activation.F90:50-184.
 * ------------------------------------------------------------------------------------------ */
static inline float act_apply(int code, float x) {
  switch (code) {
    case ACT_SOFTSIGN: return x / (fabsf(x) + 1.0f);                           /* :107-118 */
    case ACT_RELU: return fmaxf(0.0f, x);                                      /* :50-60   */
    case ACT_SIGMOID: return 1.0f / (1.0f + expf(-x));                         /* :79-89   */
    case ACT_HARD_SIGMOID: return fmaxf(0.0f, fminf(1.0f, 0.2f * x + 0.5f));   /* :135-145 */
    default: return x;                                                        /* :163-172 */
  }
}

/* one dense layer for one sample: out[o] = sum_i W[i][o]*in[i]; then (+b, activation) if requested */
static inline void dense(int n_in, int n_out, const float* W, const float* in, float* out) {
  for (int o = 0; o < n_out; ++o) out[o] = 0.0f;
  for (int i = 0; i < n_in; ++i) {
    const float xi = in[i];
    const float* w = W + (size_t)i * n_out;
    for (int o = 0; o < n_out; ++o) out[o] += w[o] * xi;
  }
}

typedef struct {
  int nlayers;        /* number of weight layers (hidden + output) */
  const int* dims;    /* nlayers+1 entries: nx, h1, ..., ny */
  const float* wpack; /* layer weights back to back, each row-major (n_in, n_out) */
  const float* bpack; /* biases back to back */
  const int* act;     /* nlayers activation codes */
} orc_net;

/* Hidden stack + last GEMM (no bias on the last layer): returns raw z in out[ny]. */
static void mlp_raw(const orc_net* net, const float* x, float* out, float* a, float* a_next,
                    const float** b_last) {
  const float* w = net->wpack;
  const float* b = net->bpack;
  const float* in = x;
  int L = net->nlayers;
  for (int n = 0; n < L - 1; ++n) {
    int n_in = net->dims[n], n_out = net->dims[n + 1];
    dense(n_in, n_out, w, in, a_next);
    for (int o = 0; o < n_out; ++o) a_next[o] = act_apply(net->act[n], a_next[o] + b[o]);
    w += (size_t)n_in * n_out;
    b += n_out;
    float* t = a; a = a_next; a_next = t;
    in = a;
  }
  dense(net->dims[L - 1], net->dims[L], w, in, out);
  *b_last = b;
}

static int max_dim(const orc_net* net) {
  int m = 0;
  for (int i = 0; i <= net->nlayers; ++i) if (net->dims[i] > m) m = net->dims[i];
  return m;
}

/* output_sgemm_tau: neural/mod_network_rrtmgp.F90:125-236.
 * output = ((ystd*(z+b) + ymean)**8) * coldry; if output2 given (SW Rayleigh net, :224-229):
 * output2 += output; output = output/output2  (tau_tot, ssa; no zero guard -- quirk Q2). */
ORC_API void orc_output_sgemm_tau(int nlayers, const int* dims, const float* wpack, const float* bpack,
                                  const int* act, const float* ymean, const float* ystd, int nbatch,
                                  const float* x, const float* coldry, float* output, float* output2) {
  orc_net net = {nlayers, dims, wpack, bpack, act};
  int nx = dims[0], ngpt = dims[nlayers], md = max_dim(&net);
#pragma omp parallel
  {
    float* a = (float*)malloc(sizeof(float) * md * 2);
#pragma omp for schedule(static)
    for (int j = 0; j < nbatch; ++j) {
      const float* bl;
      float* out = output + (size_t)j * ngpt;
      mlp_raw(&net, x + (size_t)j * nx, out, a, a + md, &bl);
      for (int i = 0; i < ngpt; ++i) {
        float o = out[i] + bl[i];
        o = ystd[i] * o + ymean[i];
        float o2 = o * o, o4 = o2 * o2;
        o = o4 * o4;                       /* **8 with an integer exponent = three squarings */
        o = o * coldry[j];
        if (output2) {
          float* p2 = output2 + (size_t)j * ngpt;
          p2[i] = p2[i] + o;
          o = o / p2[i];
        }
        out[i] = o;
      }
    }
    free(a);
  }
}

/* output_sgemm_pfrac: neural/mod_network_rrtmgp.F90:238-317: last layer bias_and_activation, then square. */
ORC_API void orc_output_sgemm_pfrac(int nlayers, const int* dims, const float* wpack, const float* bpack,
                                    const int* act, int nbatch, const float* x, float* output) {
  orc_net net = {nlayers, dims, wpack, bpack, act};
  int nx = dims[0], ny = dims[nlayers], md = max_dim(&net);
#pragma omp parallel
  {
    float* a = (float*)malloc(sizeof(float) * md * 2);
#pragma omp for schedule(static)
    for (int j = 0; j < nbatch; ++j) {
      const float* bl;
      float* out = output + (size_t)j * ny;
      mlp_raw(&net, x + (size_t)j * nx, out, a, a + md, &bl);
      for (int i = 0; i < ny; ++i) {
        float o = act_apply(act[nlayers - 1], out[i] + bl[i]);
        out[i] = o * o;
      }
    }
    free(a);
  }
}

/* output_sgemm_lw: neural/mod_network_rrtmgp.F90:319-409: raw output + bias (no post-processing). */
ORC_API void orc_output_sgemm_lw(int nlayers, const int* dims, const float* wpack, const float* bpack,
                                 const int* act, int nbatch, const float* x, float* output) {
  orc_net net = {nlayers, dims, wpack, bpack, act};
  int nx = dims[0], ny = dims[nlayers], md = max_dim(&net);
#pragma omp parallel
  {
    float* a = (float*)malloc(sizeof(float) * md * 2);
#pragma omp for schedule(static)
    for (int j = 0; j < nbatch; ++j) {
      const float* bl;
      float* out = output + (size_t)j * ny;
      mlp_raw(&net, x + (size_t)j * nx, out, a, a + md, &bl);
      for (int i = 0; i < ny; ++i) out[i] = out[i] + bl[i];
    }
    free(a);
  }
}

/* "both" network split loop: rrtmgp/kernels/mo_gas_optics_kernels.F90:745-767 */
ORC_API void orc_split_both(int ncol, int nlay, int ngpt, const float* outp_both, const float* ymean,
                            const float* ystd, const float* col_dry, float* tau, float* pfrac) {
  size_t nobs = (size_t)ncol * nlay;
  for (size_t s = 0; s < nobs; ++s) {
    const float* ob = outp_both + s * 2 * ngpt;
    for (int g = 0; g < ngpt; ++g) {
      float t = ystd[g] * ob[g] + ymean[g];
      float t2 = t * t, t4 = t2 * t2;
      t = t4 * t4;
      tau[s * ngpt + g] = t * col_dry[s];
      pfrac[s * ngpt + g] = ob[g + ngpt] * ob[g + ngpt];
    }
  }
}

/* ------------------------------------------------------------------------------------------
 * interpolate1D: rrtmgp/kernels/mo_gas_optics_kernels.F90:1024-1043 (index clamped, frac not: quirk Q4)
 * totplnk is Fortran (nPlanckTemp, nbnd) == C [nbnd][nPlanckTemp].
 * ------------------------------------------------------------------------------------------ */
static void interpolate1D(float val, float offset, float delta, const float* table, int ntemp, int nbnd,
                          float* res) {
  float val0 = (val - offset) / delta;
  float frac = val0 - (float)(int)val0;
  int index = (int)val0 + 1;          /* 1-based */
  if (index < 1) index = 1;
  if (index > ntemp - 1) index = ntemp - 1;
  for (int b = 0; b < nbnd; ++b) {
    const float* t = table + (size_t)b * ntemp;
    res[b] = t[index - 1] + frac * (t[index] - t[index - 1]);
  }
}

/* compute_Planck_source_nn: rrtmgp/kernels/mo_gas_optics_kernels.F90:615-683.
 * pfrac (in) becomes lay_source (out) in place.  sfc_lay is 1-based as in the reference
 * (merge(1,nlay,play(1,1) > play(nlay,1)), mo_gas_optics_rrtmgp.F90:402).
 * band_lims_gpt is Fortran (2,nbnd) == C [nbnd][2], 1-based inclusive g-point limits. */
ORC_API void orc_planck_source_nn(int ncol, int nlay, int nbnd, int ngpt, int ntemp, const float* tlay,
                                  const float* tlev, const float* tsfc, int sfc_lay,
                                  const int* band_lims_gpt, float temp_ref_min, float totplnk_delta,
                                  const float* totplnk, float* sfc_source, float* sfc_source_Jac,
                                  float* pfrac, float* lev_source) {
  const float delta_Tsurf = 1.0f;
#pragma omp parallel
  {
    float* pf_sfc = (float*)malloc(sizeof(float) * nbnd * 4);
    float* pf_jac = pf_sfc + nbnd;
    float* pf_lev = pf_jac + nbnd;
    float* pf_lay = pf_lev + nbnd;
#pragma omp for schedule(static)
    for (int icol = 0; icol < ncol; ++icol) {
      float* pfc = pfrac + (size_t)icol * nlay * ngpt;
      float* lev = lev_source + (size_t)icol * (nlay + 1) * ngpt;
      interpolate1D(tsfc[icol], temp_ref_min, totplnk_delta, totplnk, ntemp, nbnd, pf_sfc);
      interpolate1D(tsfc[icol] + delta_Tsurf, temp_ref_min, totplnk_delta, totplnk, ntemp, nbnd, pf_jac);
      interpolate1D(tlev[(size_t)icol * (nlay + 1) + nlay], temp_ref_min, totplnk_delta, totplnk, ntemp, nbnd,
                    pf_lev);
      for (int b = 0; b < nbnd; ++b) {
        int gS = band_lims_gpt[2 * b] - 1, gE = band_lims_gpt[2 * b + 1] - 1;
        for (int g = gS; g <= gE; ++g) {
          lev[(size_t)nlay * ngpt + g] = pfc[(size_t)(nlay - 1) * ngpt + g] * pf_lev[b];
          float p = pfc[(size_t)(sfc_lay - 1) * ngpt + g];
          sfc_source[(size_t)icol * ngpt + g] = p * pf_sfc[b];
          sfc_source_Jac[(size_t)icol * ngpt + g] = p * (pf_jac[b] - pf_sfc[b]);
        }
      }
      for (int ilay = 0; ilay < nlay; ++ilay) {
        interpolate1D(tlev[(size_t)icol * (nlay + 1) + ilay], temp_ref_min, totplnk_delta, totplnk, ntemp, nbnd,
                      pf_lev);
        interpolate1D(tlay[(size_t)icol * nlay + ilay], temp_ref_min, totplnk_delta, totplnk, ntemp, nbnd,
                      pf_lay);
        for (int b = 0; b < nbnd; ++b) {
          int gS = band_lims_gpt[2 * b] - 1, gE = band_lims_gpt[2 * b + 1] - 1;
          for (int g = gS; g <= gE; ++g) {
            float p = pfc[(size_t)ilay * ngpt + g];
            lev[(size_t)ilay * ngpt + g] = p * pf_lev[b];
            pfc[(size_t)ilay * ngpt + g] = p * pf_lay[b];
          }
        }
      }
    }
    free(pf_sfc);
  }
}

/* ------------------------------------------------------------------------------------------
 * expand: rte/mo_rte_lw.F90:429-447 (band -> g-point)
 * ------------------------------------------------------------------------------------------ */
ORC_API void orc_expand(int nband, int ngpt, int ncol, const int* band_limits, const float* arr_in,
                        float* arr_out) {
  for (int icol = 0; icol < ncol; ++icol)
    for (int b = 0; b < nband; ++b)
      for (int g = band_limits[2 * b] - 1; g <= band_limits[2 * b + 1] - 1; ++g)
        arr_out[(size_t)icol * ngpt + g] = arr_in[(size_t)icol * nband + b];
}

/* ------------------------------------------------------------------------------------------
 * lw_solver_noscat: rte/kernels/mo_rte_solver_kernels.F90:119-330 with
 *   lw_source_noscat :742-776 (ignores top_at_1: quirk Q1),
 *   lw_transport_noscat_dn :982-1009, lw_transport_noscat_up :950-980,
 *   do_rescaling (:179-181, :211-233, :275-277) with lw_transport_1rescl :1729-1795,
 *   the surface-temperature Jacobian (compute_Jac, a compile-time constant .false. in this fork,
 *   rte/mo_rte_rrtmgp_config.F90:29; :186, :270, :290, :319, :967, :975, :1761, :1782),
 *   broadband sums in four interleaved partial sums :301-314.
 * D is (ngpt,ncol).  When nmus != 1 the radiances are scaled by fac and left in
 * radn_up_out/radn_dn_out (ngpt,nlay+1,ncol) and flux_up/flux_dn are NOT written (:287-317).
 * When save_gpt && nmus == 1 the g-point arrays hold un-scaled radiances (quirk Q3).
 * ------------------------------------------------------------------------------------------ */
static void lw_solver_noscat_core(int ngpt, int nlay, int ncol, int top_at_1, int nmus, const float* D,
                                  float weight, const float* inc_flux, const float* tau,
                                  const float* lay_source, const float* lev_source, const float* sfc_emis,
                                  const float* sfc_source, float* flux_up, float* flux_dn, int save_gpt,
                                  float* radn_up_out, float* radn_dn_out, int compute_Jac,
                                  const float* sfc_source_Jac, float* flux_up_Jac, float* radn_up_Jac_out,
                                  int do_rescaling, const float* ssa, const float* gasym) {
  const float pi = acosf(-1.0f);
  const float tau_thresh = 3.4526698e-4f; /* sqrt(epsilon(1._sp)) */
  const int top_level = top_at_1 ? 0 : nlay;
  const int sfc_level = top_at_1 ? nlay : 0;
#pragma omp parallel
  {
    size_t nl = (size_t)ngpt * nlay;
    size_t nv = (size_t)ngpt * (nlay + 1);
    float* wk = (float*)malloc(sizeof(float) * (6 * nl + 3 * nv));
    float* tau_loc = wk;
    float* trans = tau_loc + nl;
    float* source_up = trans + nl;
    float* source_dn = source_up + nl;
    float* An = source_dn + nl;
    float* Cn = An + nl;
    float* radn_dn_arr = Cn + nl;
    float* radn_up_arr = radn_dn_arr + nv;
    float* radn_jac_arr = radn_up_arr + nv;
#pragma omp for schedule(static)
    for (int icol = 0; icol < ncol; ++icol) {
      float* radn_dn = save_gpt ? radn_dn_out + (size_t)icol * nv : radn_dn_arr;
      float* radn_up = save_gpt ? radn_up_out + (size_t)icol * nv : radn_up_arr;
      float* radn_jac = (save_gpt && radn_up_Jac_out) ? radn_up_Jac_out + (size_t)icol * nv : radn_jac_arr;
      const float* tauc = tau + (size_t)icol * nl;
      const float* layc = lay_source + (size_t)icol * nl;
      const float* levc = lev_source + (size_t)icol * nv;
      const float* Dc = D + (size_t)icol * ngpt;
      const float* emis = sfc_emis + (size_t)icol * ngpt;
      const float* ssrc = sfc_source + (size_t)icol * ngpt;
      /* boundary condition :196-201 */
      for (int g = 0; g < ngpt; ++g) {
        float v = inc_flux[(size_t)icol * ngpt + g];
        radn_dn[(size_t)top_level * ngpt + g] = v / (2.0f * pi * weight);
      }
      /* optical path and transmission :211-239 */
      if (do_rescaling) {
        const float* ssac = ssa + (size_t)icol * nl;
        const float* gc = gasym + (size_t)icol * nl;
        for (int l = 0; l < nlay; ++l)
          for (int g = 0; g < ngpt; ++g) {
            size_t i = (size_t)l * ngpt + g;
            float ssal = ssac[i];
            float wb = ssal * (1.0f - gc[i]) * 0.5f;
            float scaleTau = (1.0f - ssal + wb);
            Cn[i] = 0.4f * wb / scaleTau;
            float tl = tauc[i] * Dc[g] * scaleTau;
            tau_loc[i] = tl;
            trans[i] = expf(-tl);
            An[i] = (1.0f - trans[i] * trans[i]);
          }
      } else {
        for (int l = 0; l < nlay; ++l)
          for (int g = 0; g < ngpt; ++g) {
            float tl = tauc[(size_t)l * ngpt + g] * Dc[g];
            tau_loc[(size_t)l * ngpt + g] = tl;
            trans[(size_t)l * ngpt + g] = expf(-tl);
          }
      }
      /* lw_source_noscat :742-776 */
      for (int l = 0; l < nlay; ++l)
        for (int g = 0; g < ngpt; ++g) {
          size_t i = (size_t)l * ngpt + g;
          float t = tau_loc[i], tr = trans[i], fact;
          if (t > tau_thresh) fact = (1.0f - tr) / t - tr;
          else fact = t * (0.5f - 1.0f / 3.0f * t);
          float lev_dn = levc[(size_t)(l + 1) * ngpt + g], lev_up = levc[(size_t)l * ngpt + g];
          source_dn[i] = (1.0f - tr) * lev_dn + 2.0f * fact * (layc[i] - lev_dn);
          source_up[i] = (1.0f - tr) * lev_up + 2.0f * fact * (layc[i] - lev_up);
        }
      /* lw_transport_noscat_dn :982-1009 */
      if (top_at_1) {
        for (int lev = 1; lev <= nlay; ++lev)
          for (int g = 0; g < ngpt; ++g)
            radn_dn[(size_t)lev * ngpt + g] = trans[(size_t)(lev - 1) * ngpt + g] * radn_dn[(size_t)(lev - 1) * ngpt + g] +
                                              source_dn[(size_t)(lev - 1) * ngpt + g];
      } else {
        for (int lev = nlay - 1; lev >= 0; --lev)
          for (int g = 0; g < ngpt; ++g)
            radn_dn[(size_t)lev * ngpt + g] = trans[(size_t)lev * ngpt + g] * radn_dn[(size_t)(lev + 1) * ngpt + g] +
                                              source_dn[(size_t)lev * ngpt + g];
      }
      /* surface reflection and emission :269-270 */
      for (int g = 0; g < ngpt; ++g) {
        radn_up[(size_t)sfc_level * ngpt + g] =
            radn_dn[(size_t)sfc_level * ngpt + g] * (1 - emis[g]) + emis[g] * ssrc[g];
        if (compute_Jac) radn_jac[(size_t)sfc_level * ngpt + g] = emis[g] * sfc_source_Jac[(size_t)icol * ngpt + g];
      }
      if (do_rescaling) {
        /* lw_transport_1rescl :1729-1795: up with the adjustment, then down again */
        if (top_at_1) {
          for (int l = nlay - 1; l >= 0; --l)
            for (int g = 0; g < ngpt; ++g) {
              size_t i = (size_t)l * ngpt + g;
              float adj = Cn[i] * (An[i] * radn_dn[i] - trans[i] * source_dn[i] - source_up[i]);
              radn_up[i] = trans[i] * radn_up[i + ngpt] + source_up[i] + adj;
              if (compute_Jac) radn_jac[i] = trans[i] * radn_jac[i + ngpt];
            }
          for (int l = 0; l < nlay; ++l)
            for (int g = 0; g < ngpt; ++g) {
              size_t i = (size_t)l * ngpt + g;
              float adj = Cn[i] * (An[i] * radn_up[i] - trans[i] * source_up[i] - source_dn[i]);
              radn_dn[i + ngpt] = trans[i] * radn_dn[i] + source_dn[i] + adj;
            }
        } else {
          for (int l = 0; l < nlay; ++l)
            for (int g = 0; g < ngpt; ++g) {
              size_t i = (size_t)l * ngpt + g;
              float adj = Cn[i] * (An[i] * radn_dn[i + ngpt] - trans[i] * source_dn[i] - source_up[i]);
              radn_up[i + ngpt] = trans[i] * radn_up[i] + source_up[i] + adj;
              if (compute_Jac) radn_jac[i + ngpt] = trans[i] * radn_jac[i];
            }
          for (int l = nlay - 1; l >= 0; --l)
            for (int g = 0; g < ngpt; ++g) {
              size_t i = (size_t)l * ngpt + g;
              float adj = Cn[i] * (An[i] * radn_up[i] - trans[i] * source_up[i] - source_dn[i]);
              radn_dn[i] = trans[i] * radn_dn[i + ngpt] + source_dn[i] + adj;
            }
        }
      } else {
        /* lw_transport_noscat_up :950-980 */
        if (top_at_1) {
          for (int l = nlay - 1; l >= 0; --l)
            for (int g = 0; g < ngpt; ++g) {
              radn_up[(size_t)l * ngpt + g] = trans[(size_t)l * ngpt + g] * radn_up[(size_t)(l + 1) * ngpt + g] +
                                              source_up[(size_t)l * ngpt + g];
              if (compute_Jac) radn_jac[(size_t)l * ngpt + g] = trans[(size_t)l * ngpt + g] * radn_jac[(size_t)(l + 1) * ngpt + g];
            }
        } else {
          for (int lev = 1; lev <= nlay; ++lev)
            for (int g = 0; g < ngpt; ++g) {
              radn_up[(size_t)lev * ngpt + g] = trans[(size_t)(lev - 1) * ngpt + g] * radn_up[(size_t)(lev - 1) * ngpt + g] +
                                                source_up[(size_t)(lev - 1) * ngpt + g];
              if (compute_Jac)
                radn_jac[(size_t)lev * ngpt + g] = trans[(size_t)(lev - 1) * ngpt + g] * radn_jac[(size_t)(lev - 1) * ngpt + g];
            }
        }
      }
      float fac = 2.0f * pi * weight;
      if (nmus != 1) {
        for (size_t i = 0; i < nv; ++i) {
          radn_dn[i] = fac * radn_dn[i]; radn_up[i] = fac * radn_up[i];
          if (compute_Jac) radn_jac[i] = fac * radn_jac[i];
        }
      } else {
        if (ngpt % 4 == 0) {
          for (int lev = 0; lev <= nlay; ++lev) {
            float su[4] = {0, 0, 0, 0}, sd[4] = {0, 0, 0, 0};
            for (int g = 0; g < ngpt; g += 4)
              for (int j = 0; j < 4; ++j) {
                su[j] = su[j] + fac * radn_up[(size_t)lev * ngpt + g + j];
                sd[j] = sd[j] + fac * radn_dn[(size_t)lev * ngpt + g + j];
              }
            flux_up[(size_t)icol * (nlay + 1) + lev] = su[0] + su[1] + su[2] + su[3];
            flux_dn[(size_t)icol * (nlay + 1) + lev] = sd[0] + sd[1] + sd[2] + sd[3];
          }
        } else {
          /* :311-312 -- sum() of the un-scaled radiances, as written in the reference */
          for (int lev = 0; lev <= nlay; ++lev) {
            float su = 0, sd = 0;
            for (int g = 0; g < ngpt; ++g) { su += radn_up[(size_t)lev * ngpt + g]; sd += radn_dn[(size_t)lev * ngpt + g]; }
            flux_up[(size_t)icol * (nlay + 1) + lev] = su;
            flux_dn[(size_t)icol * (nlay + 1) + lev] = sd;
          }
        }
        if (compute_Jac) /* :319 -- sum() of the un-scaled Jacobian radiances, as written */
          for (int lev = 0; lev <= nlay; ++lev) {
            float sj = 0;
            for (int g = 0; g < ngpt; ++g) sj += radn_jac[(size_t)lev * ngpt + g];
            flux_up_Jac[(size_t)icol * (nlay + 1) + lev] = sj;
          }
      }
    }
    free(wk);
  }
}

ORC_API void orc_lw_solver_noscat(int ngpt, int nlay, int ncol, int top_at_1, int nmus, const float* D,
                                  float weight, const float* inc_flux, const float* tau,
                                  const float* lay_source, const float* lev_source, const float* sfc_emis,
                                  const float* sfc_source, float* flux_up, float* flux_dn, int save_gpt,
                                  float* radn_up_out, float* radn_dn_out) {
  lw_solver_noscat_core(ngpt, nlay, ncol, top_at_1, nmus, D, weight, inc_flux, tau, lay_source, lev_source, sfc_emis,
                        sfc_source, flux_up, flux_dn, save_gpt, radn_up_out, radn_dn_out, 0, NULL, NULL, NULL, 0, NULL, NULL);
}

/* sum_broadband: rte/kernels/mo_fluxes_broadband_kernels.F90:31-74 */
static void sum_broadband(int ngpt, int nlev, int ncol, const float* spectral, float* broadband) {
  for (int icol = 0; icol < ncol; ++icol)
    for (int lev = 0; lev < nlev; ++lev) {
      float s = 0.0f;
      const float* p = spectral + ((size_t)icol * nlev + lev) * ngpt;
      for (int g = 0; g < ngpt; ++g) s += p[g];
      broadband[(size_t)icol * nlev + lev] = s;
    }
}

/* lw_solver_noscat_GaussQuad: rte/kernels/mo_rte_solver_kernels.F90:332-415 */
ORC_API void orc_lw_solver_noscat_GaussQuad(int ngpt, int nlay, int ncol, int top_at_1, int nmus, const float* Ds,
                                            const float* weights, const float* inc_flux, const float* tau,
                                            const float* lay_source, const float* lev_source,
                                            const float* sfc_emis, const float* sfc_source, float* flux_up,
                                            float* flux_dn) {
  size_t ngc = (size_t)ngpt * ncol;
  float* Dg = (float*)malloc(sizeof(float) * ngc);
  for (size_t i = 0; i < ngc; ++i) Dg[i] = Ds[0];
  if (nmus == 1) {
    orc_lw_solver_noscat(ngpt, nlay, ncol, top_at_1, nmus, Dg, weights[0], inc_flux, tau, lay_source, lev_source,
                         sfc_emis, sfc_source, flux_up, flux_dn, 0, NULL, NULL);
  } else {
    size_t n3 = ngc * (nlay + 1);
    float* gup = (float*)malloc(sizeof(float) * n3 * 4);
    float* gdn = gup + n3;
    float* rup = gdn + n3;
    float* rdn = rup + n3;
    orc_lw_solver_noscat(ngpt, nlay, ncol, top_at_1, nmus, Dg, weights[0], inc_flux, tau, lay_source, lev_source,
                         sfc_emis, sfc_source, flux_up, flux_dn, 1, gup, gdn);
    for (int imu = 1; imu < nmus; ++imu) {
      for (size_t i = 0; i < ngc; ++i) Dg[i] = Ds[imu];
      orc_lw_solver_noscat(ngpt, nlay, ncol, top_at_1, nmus, Dg, weights[imu], inc_flux, tau, lay_source,
                           lev_source, sfc_emis, sfc_source, flux_up, flux_dn, 1, rup, rdn);
      for (size_t i = 0; i < n3; ++i) { gup[i] = gup[i] + rup[i]; gdn[i] = gdn[i] + rdn[i]; }
    }
    sum_broadband(ngpt, nlay + 1, ncol, gup, flux_up);
    sum_broadband(ngpt, nlay + 1, ncol, gdn, flux_dn);
    free(gup);
  }
  free(Dg);
}

/* lw_solver_noscat_GaussQuad with everything rte_lw can ask of it (rte/mo_rte_lw.F90:324-384, kernels :332-415):
 * per-g-point secants lw_Ds (one angle, :329-340; Ds_gpt (ngpt,ncol) or NULL), re-scaled scattering for _2str clouds
 * (ssa, g or NULL), the surface-temperature Jacobian (sfc_source_Jac, flux_up_Jac or NULL) and the g-point fluxes
 * (gpt_up, gpt_dn (ngpt,nlay+1,ncol) or NULL; quirk Q3 for one angle). */
ORC_API void orc_lw_solver_noscat_GaussQuad_ext(int ngpt, int nlay, int ncol, int top_at_1, int nmus, const float* Ds,
                                                const float* weights, const float* Ds_gpt, const float* inc_flux,
                                                const float* tau, const float* ssa, const float* g,
                                                const float* lay_source, const float* lev_source, const float* sfc_emis,
                                                const float* sfc_source, const float* sfc_source_Jac, float* flux_up,
                                                float* flux_dn, float* flux_up_Jac, float* gpt_up, float* gpt_dn) {
  size_t ngc = (size_t)ngpt * ncol;
  size_t n3 = ngc * (nlay + 1);
  const int jac = flux_up_Jac != NULL, resc = ssa != NULL, save = gpt_up != NULL;
  float* Dg = (float*)malloc(sizeof(float) * ngc);
  for (size_t i = 0; i < ngc; ++i) Dg[i] = Ds_gpt ? Ds_gpt[i] : Ds[0];
  if (nmus == 1) {
    float* gj = (save && jac) ? (float*)malloc(sizeof(float) * n3) : NULL;
    lw_solver_noscat_core(ngpt, nlay, ncol, top_at_1, nmus, Dg, weights[0], inc_flux, tau, lay_source, lev_source, sfc_emis,
                          sfc_source, flux_up, flux_dn, save, gpt_up, gpt_dn, jac, sfc_source_Jac, flux_up_Jac, gj, resc, ssa, g);
    free(gj);
  } else {
    float* own = save ? NULL : (float*)malloc(sizeof(float) * n3 * 2);
    float* gup = save ? gpt_up : own;
    float* gdn = save ? gpt_dn : own + n3;
    float* tmp = (float*)malloc(sizeof(float) * n3 * 4);
    float *rup = tmp, *rdn = tmp + n3, *gjac = tmp + 2 * n3, *rjac = tmp + 3 * n3;
    lw_solver_noscat_core(ngpt, nlay, ncol, top_at_1, nmus, Dg, weights[0], inc_flux, tau, lay_source, lev_source, sfc_emis,
                          sfc_source, flux_up, flux_dn, 1, gup, gdn, jac, sfc_source_Jac, flux_up_Jac, gjac, resc, ssa, g);
    for (int imu = 1; imu < nmus; ++imu) {
      for (size_t i = 0; i < ngc; ++i) Dg[i] = Ds[imu];
      lw_solver_noscat_core(ngpt, nlay, ncol, top_at_1, nmus, Dg, weights[imu], inc_flux, tau, lay_source, lev_source,
                            sfc_emis, sfc_source, flux_up, flux_dn, 1, rup, rdn, jac, sfc_source_Jac, flux_up_Jac, rjac, resc, ssa, g);
      for (size_t i = 0; i < n3; ++i) { gup[i] = gup[i] + rup[i]; gdn[i] = gdn[i] + rdn[i]; if (jac) gjac[i] = gjac[i] + rjac[i]; }
    }
    sum_broadband(ngpt, nlay + 1, ncol, gup, flux_up);
    sum_broadband(ngpt, nlay + 1, ncol, gdn, flux_dn);
    if (jac) sum_broadband(ngpt, nlay + 1, ncol, gjac, flux_up_Jac);
    free(tmp);
    free(own);
  }
  free(Dg);
}

/* ------------------------------------------------------------------------------------------
 * lw_solver_2stream: rte/kernels/mo_rte_solver_kernels.F90:426-486 with lw_two_stream :1018-1069 (LW_diff_sec = 1.66,
 * k_min), lw_source_2str :1112-1162 (lev_source used as it is; lay_source unused; tau <= 1e-8 -> no source),
 * adding :1526-1637 with sfc_albedo = 1 - sfc_emis and source_sfc = pi * sfc_emis * sfc_source, plain sums over
 * g-points (sum_broadband_nocol, mo_fluxes_broadband_kernels.F90:40-47).  gpt_up / gpt_dn (ngpt,nlay+1,ncol) optional.
 * ------------------------------------------------------------------------------------------ */
ORC_API void orc_lw_solver_2stream(int ngpt, int nlay, int ncol, int top_at_1, const float* inc_flux, const float* tau,
                                   const float* ssa, const float* gasym, const float* lev_source, const float* sfc_emis,
                                   const float* sfc_source, float* flux_up, float* flux_dn, float* gpt_up, float* gpt_dn) {
  const float k_min = 1.e-4f;
  const float LW_diff_sec = 1.66f;
  const float pi = acosf(-1.0f);
  const int top_level = top_at_1 ? 0 : nlay;
#pragma omp parallel
  {
    size_t nl = (size_t)ngpt * nlay, nv = (size_t)ngpt * (nlay + 1);
    float* wk = (float*)malloc(sizeof(float) * (5 * nl + 4 * nv));
    float* Rdif = wk;
    float* Tdif = Rdif + nl;
    float* source_up = Tdif + nl;
    float* source_dn = source_up + nl;
    float* denom = source_dn + nl;
    float* radn_up_arr = denom + nl;
    float* radn_dn_arr = radn_up_arr + nv;
    float* albedo = radn_dn_arr + nv;
    float* src = albedo + nv;
#pragma omp for schedule(static)
    for (int icol = 0; icol < ncol; ++icol) {
      float* radn_up = gpt_up ? gpt_up + (size_t)icol * nv : radn_up_arr;
      float* radn_dn = gpt_up ? gpt_dn + (size_t)icol * nv : radn_dn_arr;
      const float* tauc = tau + (size_t)icol * nl;
      const float* w0c = ssa + (size_t)icol * nl;
      const float* gc = gasym + (size_t)icol * nl;
      const float* levc = lev_source + (size_t)icol * nv;
      for (int g = 0; g < ngpt; ++g) radn_dn[(size_t)top_level * ngpt + g] = inc_flux ? inc_flux[(size_t)icol * ngpt + g] : 0.0f;
      for (int l = 0; l < nlay; ++l)
        for (int g = 0; g < ngpt; ++g) {
          size_t i = (size_t)l * ngpt + g;
          float gamma1 = LW_diff_sec * (1.0f - 0.5f * w0c[i] * (1.0f + gc[i]));
          float gamma2 = LW_diff_sec * 0.5f * w0c[i] * (1.0f - gc[i]);
          float k = sqrtf(fmaxf((gamma1 - gamma2) * (gamma1 + gamma2), k_min));
          float e1 = expf(-tauc[i] * k);
          float e2 = e1 * e1;
          float RT = 1.0f / (k * (1.0f + e2) + gamma1 * (1.0f - e2));
          Rdif[i] = RT * gamma2 * (1.0f - e2);
          Tdif[i] = RT * 2.0f * k * e1;
          float top = top_at_1 ? levc[i] : levc[i + ngpt], bot = top_at_1 ? levc[i + ngpt] : levc[i];
          if (tauc[i] > 1.0e-8f) {
            float Z = (bot - top) / (tauc[i] * (gamma1 + gamma2));
            float Zup_top = Z + top, Zup_bottom = Z + bot, Zdn_top = -Z + top, Zdn_bottom = -Z + bot;
            source_up[i] = pi * (Zup_top - Rdif[i] * Zdn_top - Tdif[i] * Zup_bottom);
            source_dn[i] = pi * (Zdn_bottom - Rdif[i] * Zup_bottom - Tdif[i] * Zdn_top);
          } else {
            source_up[i] = 0.0f;
            source_dn[i] = 0.0f;
          }
        }
      /* adding :1526-1637 */
      const float* em = sfc_emis + (size_t)icol * ngpt;
      const float* ss = sfc_source + (size_t)icol * ngpt;
      if (top_at_1) {
        for (int g = 0; g < ngpt; ++g) { albedo[(size_t)nlay * ngpt + g] = 1.0f - em[g]; src[(size_t)nlay * ngpt + g] = pi * em[g] * ss[g]; }
        for (int l = nlay - 1; l >= 0; --l)
          for (int g = 0; g < ngpt; ++g) {
            size_t i = (size_t)l * ngpt + g, ip = (size_t)(l + 1) * ngpt + g;
            denom[i] = 1.0f / (1.0f - Rdif[i] * albedo[ip]);
            albedo[i] = Rdif[i] + Tdif[i] * Tdif[i] * albedo[ip] * denom[i];
            src[i] = source_up[i] + Tdif[i] * denom[i] * (src[ip] + albedo[ip] * source_dn[i]);
          }
        for (int g = 0; g < ngpt; ++g) radn_up[g] = radn_dn[g] * albedo[g] + src[g];
        for (int lev = 1; lev <= nlay; ++lev)
          for (int g = 0; g < ngpt; ++g) {
            size_t i = (size_t)lev * ngpt + g, im = (size_t)(lev - 1) * ngpt + g;
            radn_dn[i] = (Tdif[im] * radn_dn[im] + Rdif[im] * src[i] + source_dn[im]) * denom[im];
            radn_up[i] = radn_dn[i] * albedo[i] + src[i];
          }
      } else {
        for (int g = 0; g < ngpt; ++g) { albedo[g] = 1.0f - em[g]; src[g] = pi * em[g] * ss[g]; }
        for (int l = 0; l < nlay; ++l)
          for (int g = 0; g < ngpt; ++g) {
            size_t i = (size_t)l * ngpt + g, ip = (size_t)(l + 1) * ngpt + g;
            denom[i] = 1.0f / (1.0f - Rdif[i] * albedo[i]);
            albedo[ip] = Rdif[i] + Tdif[i] * Tdif[i] * albedo[i] * denom[i];
            src[ip] = source_up[i] + Tdif[i] * denom[i] * (src[i] + albedo[i] * source_dn[i]);
          }
        {
          size_t t = (size_t)nlay * ngpt;
          for (int g = 0; g < ngpt; ++g) radn_up[t + g] = radn_dn[t + g] * albedo[t + g] + src[t + g];
        }
        for (int l = nlay - 1; l >= 0; --l)
          for (int g = 0; g < ngpt; ++g) {
            size_t i = (size_t)l * ngpt + g, ip = (size_t)(l + 1) * ngpt + g;
            radn_dn[i] = (Tdif[i] * radn_dn[ip] + Rdif[i] * src[i] + source_dn[i]) * denom[i];
            radn_up[i] = radn_dn[i] * albedo[i] + src[i];
          }
      }
      for (int lev = 0; lev <= nlay; ++lev) {
        float su = 0, sd = 0;
        for (int g = 0; g < ngpt; ++g) { su += radn_up[(size_t)lev * ngpt + g]; sd += radn_dn[(size_t)lev * ngpt + g]; }
        flux_up[(size_t)icol * (nlay + 1) + lev] = su;
        flux_dn[(size_t)icol * (nlay + 1) + lev] = sd;
      }
    }
    free(wk);
  }
}

/* ------------------------------------------------------------------------------------------
 * sw_solver_2stream: rte/kernels/mo_rte_solver_kernels.F90:541-692 with
 *   sw_two_stream_source :1366-1480, adding :1526-1637, k_min = 1e-4 (:76-82),
 *   broadband sums :643-680.
 * ------------------------------------------------------------------------------------------ */
static void sw_solver_2stream_core(int ngpt, int nlay, int ncol, int top_at_1, const float* inc_flux,
                                   const float* inc_flux_dif, const float* tau, const float* ssa,
                                   const float* gasym, const float* mu0v, const float* sfc_alb_dir,
                                   const float* sfc_alb_dif, float* flux_up, float* flux_dn, float* flux_dir,
                                   float* gpt_up, float* gpt_dn, float* gpt_dir) {
  /* gpt_*: optional g-point fluxes (ngpt,nlay+1,ncol), save_gpt_flux :557-587; flux_dn_gpt is the TOTAL downward flux :660-663 */
  const float k_min = 1.e-4f;
  const int top_level = top_at_1 ? 0 : nlay;
#pragma omp parallel
  {
    size_t nl = (size_t)ngpt * nlay, nv = (size_t)ngpt * (nlay + 1);
    float* wk = (float*)malloc(sizeof(float) * (5 * nl + 5 * nv + ngpt));
    float* Rdif = wk;
    float* Tdif = Rdif + nl;
    float* source_up = Tdif + nl;
    float* source_dn = source_up + nl;
    float* denom = source_dn + nl;
    float* radn_up_arr = denom + nl;
    float* radn_dn_arr = radn_up_arr + nv;
    float* radn_dir_arr = radn_dn_arr + nv;
    float* albedo = radn_dir_arr + nv;
    float* src = albedo + nv;
    float* source_sfc = src + nv;
#pragma omp for schedule(static)
    for (int icol = 0; icol < ncol; ++icol) {
      const float* tauc = tau + (size_t)icol * nl;
      const float* w0c = ssa + (size_t)icol * nl;
      const float* gc = gasym + (size_t)icol * nl;
      const float mu0 = mu0v[icol];
      float* radn_up = gpt_up ? gpt_up + (size_t)icol * nv : radn_up_arr;
      float* radn_dn = gpt_up ? gpt_dn + (size_t)icol * nv : radn_dn_arr;
      float* radn_dir = gpt_up ? gpt_dir + (size_t)icol * nv : radn_dir_arr;
      for (int g = 0; g < ngpt; ++g) {
        radn_dir[(size_t)top_level * ngpt + g] = inc_flux[(size_t)icol * ngpt + g] * mu0;
        radn_dn[(size_t)top_level * ngpt + g] = inc_flux_dif[(size_t)icol * ngpt + g];
      }
      /* ---- sw_two_stream_source :1366-1480 ---- */
      const float mu0_inv = 1.0f / mu0;
      const float* dir_flux_trans = NULL;
      for (int j = 0; j < nlay; ++j) {
        int ilev;
        float *dinc, *dtrans;
        if (top_at_1) { ilev = j; dinc = radn_dir + (size_t)ilev * ngpt; dtrans = radn_dir + (size_t)(ilev + 1) * ngpt; }
        else { ilev = nlay - j - 1; dinc = radn_dir + (size_t)(ilev + 1) * ngpt; dtrans = radn_dir + (size_t)ilev * ngpt; }
        for (int g = 0; g < ngpt; ++g) {
          size_t i = (size_t)ilev * ngpt + g;
          float Tnoscat = expf(-tauc[i] * mu0_inv);
          float w0 = w0c[i], gg = gc[i];
          float gamma1 = (8.0f - w0 * (5.0f + 3.0f * gg)) * .25f;
          float gamma2 = 3.0f * (w0 * (1.0f - gg)) * .25f;
          float gamma3 = (2.0f - 3.0f * mu0 * gg) * .25f;
          float gamma4 = 1.0f - gamma3;
          float alpha1 = gamma1 * gamma4 + gamma2 * gamma3;
          float alpha2 = gamma1 * gamma3 + gamma2 * gamma4;
          float k = sqrtf(fmaxf((gamma1 - gamma2) * (gamma1 + gamma2), k_min));
          float exp_minusktau = expf(-tauc[i] * k);
          float exp_minus2ktau = exp_minusktau * exp_minusktau;
          float k_2_exponential = 2.0f * k * exp_minusktau;
          float RT_term = 1.0f / (k * (1.0f + exp_minus2ktau) + gamma1 * (1.0f - exp_minus2ktau));
          Rdif[i] = RT_term * gamma2 * (1.0f - exp_minus2ktau);
          Tdif[i] = RT_term * 2.0f * k * exp_minusktau;
          float k_mu = k * mu0;
          float k_mu2 = k_mu * k_mu;
          float k_gamma3 = k * gamma3;
          float k_gamma4 = k * gamma4;
          float dd = (fabsf(1.0f - k_mu2) >= FLT_EPSILON) ? (1.0f - k_mu2) : FLT_EPSILON;
          RT_term = w0 * RT_term / dd;
          float Rdir = RT_term * ((1.0f - k_mu) * (alpha2 + k_gamma3) - (1.0f + k_mu) * (alpha2 - k_gamma3) * exp_minus2ktau -
                                  k_2_exponential * (gamma3 - alpha2 * mu0) * Tnoscat);
          float Tdir = RT_term * (k_2_exponential * (gamma4 + alpha1 * mu0) -
                                  Tnoscat * ((1.0f + k_mu) * (alpha1 + k_gamma4) - (1.0f - k_mu) * (alpha1 - k_gamma4) * exp_minus2ktau));
          Rdir = fmaxf(0.0f, fminf(Rdir, (1.0f - Tnoscat)));
          Tdir = fmaxf(0.0f, fminf(Tdir, (1.0f - Tnoscat - Rdir)));
          source_up[i] = Rdir * dinc[g];
          source_dn[i] = Tdir * dinc[g];
          dtrans[g] = Tnoscat * dinc[g];
        }
        dir_flux_trans = dtrans;
      }
      for (int g = 0; g < ngpt; ++g) source_sfc[g] = dir_flux_trans[g] * sfc_alb_dir[(size_t)icol * ngpt + g];
      /* ---- adding :1526-1637 ---- */
      const float* alb_sfc = sfc_alb_dif + (size_t)icol * ngpt;
      if (top_at_1) {
        for (int g = 0; g < ngpt; ++g) { albedo[(size_t)nlay * ngpt + g] = alb_sfc[g]; src[(size_t)nlay * ngpt + g] = source_sfc[g]; }
        for (int l = nlay - 1; l >= 0; --l)
          for (int g = 0; g < ngpt; ++g) {
            size_t i = (size_t)l * ngpt + g, ip = (size_t)(l + 1) * ngpt + g;
            denom[i] = 1.0f / (1.0f - Rdif[i] * albedo[ip]);
            albedo[i] = Rdif[i] + Tdif[i] * Tdif[i] * albedo[ip] * denom[i];
            src[i] = source_up[i] + Tdif[i] * denom[i] * (src[ip] + albedo[ip] * source_dn[i]);
          }
        for (int g = 0; g < ngpt; ++g) radn_up[g] = radn_dn[g] * albedo[g] + src[g];
        for (int lev = 1; lev <= nlay; ++lev)
          for (int g = 0; g < ngpt; ++g) {
            size_t i = (size_t)lev * ngpt + g, im = (size_t)(lev - 1) * ngpt + g;
            radn_dn[i] = (Tdif[im] * radn_dn[im] + Rdif[im] * src[i] + source_dn[im]) * denom[im];
            radn_up[i] = radn_dn[i] * albedo[i] + src[i];
          }
      } else {
        for (int g = 0; g < ngpt; ++g) { albedo[g] = alb_sfc[g]; src[g] = source_sfc[g]; }
        for (int l = 0; l < nlay; ++l)
          for (int g = 0; g < ngpt; ++g) {
            size_t i = (size_t)l * ngpt + g, ip = (size_t)(l + 1) * ngpt + g;
            denom[i] = 1.0f / (1.0f - Rdif[i] * albedo[i]);
            albedo[ip] = Rdif[i] + Tdif[i] * Tdif[i] * albedo[i] * denom[i];
            src[ip] = source_up[i] + Tdif[i] * denom[i] * (src[i] + albedo[i] * source_dn[i]);
          }
        {
          size_t t = (size_t)nlay * ngpt;
          for (int g = 0; g < ngpt; ++g) radn_up[t + g] = radn_dn[t + g] * albedo[t + g] + src[t + g];
        }
        for (int l = nlay - 1; l >= 0; --l)
          for (int g = 0; g < ngpt; ++g) {
            size_t i = (size_t)l * ngpt + g, ip = (size_t)(l + 1) * ngpt + g;
            radn_dn[i] = (Tdif[i] * radn_dn[ip] + Rdif[i] * src[i] + source_dn[i]) * denom[i];
            radn_up[i] = radn_dn[i] * albedo[i] + src[i];
          }
      }
      /* ---- broadband sums :643-680 ---- */
      if (ngpt % 4 == 0) {
        for (int lev = 0; lev <= nlay; ++lev) {
          float su[4] = {0, 0, 0, 0}, sd[4] = {0, 0, 0, 0}, sr[4] = {0, 0, 0, 0};
          for (int g = 0; g < ngpt; g += 4)
            for (int j = 0; j < 4; ++j) {
              size_t i = (size_t)lev * ngpt + g + j;
              su[j] = su[j] + radn_up[i];
              sr[j] = sr[j] + radn_dir[i];
              if (gpt_up) {
                radn_dn[i] = radn_dn[i] + radn_dir[i];
                sd[j] = sd[j] + radn_dn[i];
              } else {
                sd[j] = sd[j] + radn_dn[i] + radn_dir[i];
              }
            }
          flux_up[(size_t)icol * (nlay + 1) + lev] = su[0] + su[1] + su[2] + su[3];
          flux_dn[(size_t)icol * (nlay + 1) + lev] = sd[0] + sd[1] + sd[2] + sd[3];
          flux_dir[(size_t)icol * (nlay + 1) + lev] = sr[0] + sr[1] + sr[2] + sr[3];
        }
      } else {
        for (int lev = 0; lev <= nlay; ++lev) {
          float su = 0, sd = 0, sr = 0;
          for (int g = 0; g < ngpt; ++g) {
            size_t i = (size_t)lev * ngpt + g;
            if (gpt_up) radn_dn[i] = radn_dn[i] + radn_dir[i];
            su += radn_up[i]; sr += radn_dir[i]; sd += gpt_up ? radn_dn[i] : (radn_dn[i] + radn_dir[i]);
          }
          flux_up[(size_t)icol * (nlay + 1) + lev] = su;
          flux_dn[(size_t)icol * (nlay + 1) + lev] = sd;
          flux_dir[(size_t)icol * (nlay + 1) + lev] = sr;
        }
      }
    }
    free(wk);
  }
}

ORC_API void orc_sw_solver_2stream(int ngpt, int nlay, int ncol, int top_at_1, const float* inc_flux,
                                   const float* inc_flux_dif, const float* tau, const float* ssa,
                                   const float* gasym, const float* mu0v, const float* sfc_alb_dir,
                                   const float* sfc_alb_dif, float* flux_up, float* flux_dn, float* flux_dir) {
  sw_solver_2stream_core(ngpt, nlay, ncol, top_at_1, inc_flux, inc_flux_dif, tau, ssa, gasym, mu0v, sfc_alb_dir, sfc_alb_dif, flux_up,
                         flux_dn, flux_dir, NULL, NULL, NULL);
}

ORC_API void orc_sw_solver_2stream_gpt(int ngpt, int nlay, int ncol, int top_at_1, const float* inc_flux,
                                       const float* inc_flux_dif, const float* tau, const float* ssa,
                                       const float* gasym, const float* mu0v, const float* sfc_alb_dir,
                                       const float* sfc_alb_dif, float* flux_up, float* flux_dn, float* flux_dir,
                                       float* gpt_up, float* gpt_dn, float* gpt_dir) {
  sw_solver_2stream_core(ngpt, nlay, ncol, top_at_1, inc_flux, inc_flux_dif, tau, ssa, gasym, mu0v, sfc_alb_dir, sfc_alb_dif, flux_up,
                         flux_dn, flux_dir, gpt_up, gpt_dn, gpt_dir);
}

/* ------------------------------------------------------------------------------------------
 * Cloud optics from lookup tables: extensions/cloud_optics/mo_cloud_optics.F90:603-645 and the
 * liquid+ice combination :505-528.  Tables are Fortran (nsteps,nbnd) == C [nbnd][nsteps].
 * two_stream = 0 -> 1scl absorption optical depth; 1 -> tau, ssa, g by band.
 * ------------------------------------------------------------------------------------------ */
static void table_all(int nbnd, int nsteps, float step, float offset, float wp, float re, const float* ext,
                      const float* ssa, const float* asy, int mask, float* t, float* ts, float* tsg) {
  if (!mask) { for (int b = 0; b < nbnd; ++b) { t[b] = 0; ts[b] = 0; tsg[b] = 0; } return; }
  int index = (int)floorf((re - offset) / step) + 1;
  if (index > nsteps - 1) index = nsteps - 1;
  float fint = (re - offset) / step - (float)(index - 1);
  for (int b = 0; b < nbnd; ++b) {
    const float* e = ext + (size_t)b * nsteps; const float* s = ssa + (size_t)b * nsteps; const float* a = asy + (size_t)b * nsteps;
    float tt = wp * (e[index - 1] + fint * (e[index] - e[index - 1]));
    float tts = tt * (s[index - 1] + fint * (s[index] - s[index - 1]));
    tsg[b] = tts * (a[index - 1] + fint * (a[index] - a[index - 1]));
    ts[b] = tts;
    t[b] = tt;
  }
}

ORC_API void orc_cloud_optics_lut(int ncol, int nlay, int nbnd, const float* clwp, const float* ciwp,
                                  const float* reliq, const float* reice, int liq_nsteps, float liq_step,
                                  float radliq_lwr, const float* extliq, const float* ssaliq, const float* asyliq,
                                  int ice_nsteps, float ice_step, float radice_lwr, const float* extice,
                                  const float* ssaice, const float* asyice, int two_stream, float* tau,
                                  float* ssa, float* g) {
  float* wk = (float*)malloc(sizeof(float) * nbnd * 6);
  float *lt = wk, *lts = lt + nbnd, *ltsg = lts + nbnd, *it = ltsg + nbnd, *its = it + nbnd, *itsg = its + nbnd;
  for (size_t s = 0; s < (size_t)ncol * nlay; ++s) {
    table_all(nbnd, liq_nsteps, liq_step, radliq_lwr, clwp[s], reliq[s], extliq, ssaliq, asyliq, clwp[s] > 0.0f, lt, lts, ltsg);
    table_all(nbnd, ice_nsteps, ice_step, radice_lwr, ciwp[s], reice[s], extice, ssaice, asyice, ciwp[s] > 0.0f, it, its, itsg);
    for (int b = 0; b < nbnd; ++b) {
      if (!two_stream) {
        tau[s * nbnd + b] = (lt[b] - lts[b]) + (it[b] - its[b]);
      } else {
        float t = lt[b] + it[b];
        float ts = lts[b] + its[b];
        g[s * nbnd + b] = (ltsg[b] + itsg[b]) / fmaxf(FLT_EPSILON, ts);
        ssa[s * nbnd + b] = ts / fmaxf(FLT_EPSILON, t);
        tau[s * nbnd + b] = t;
      }
    }
  }
  free(wk);
}

/* delta_scale_2str_k: rte/kernels/mo_optical_props_kernels.F90:72-93; eps = 3*tiny (:36) */
ORC_API void orc_delta_scale_2str(size_t n, float* tau, float* ssa, float* g) {
  const float eps = 3.0f * FLT_MIN;
  for (size_t i = 0; i < n; ++i) {
    float f = g[i] * g[i];
    float wf = ssa[i] * f;
    tau[i] = (1.0f - wf) * tau[i];
    ssa[i] = (ssa[i] - wf) / fmaxf(eps, (1.0f - wf));
    g[i] = (g[i] - f) / fmaxf(eps, (1.0f - f));
  }
}

/* inc_1scalar_by_1scalar_bybnd: rte/kernels/mo_optical_props_kernels.F90:358-378 */
ORC_API void orc_inc_1scalar_by_1scalar_bybnd(int ngpt, int nlay, int ncol, float* tau1, const float* tau2,
                                              int nbnd, const int* gpt_lims) {
  for (size_t s = 0; s < (size_t)ncol * nlay; ++s)
    for (int b = 0; b < nbnd; ++b)
      for (int g = gpt_lims[2 * b] - 1; g <= gpt_lims[2 * b + 1] - 1; ++g)
        tau1[s * ngpt + g] = tau1[s * ngpt + g] + tau2[s * nbnd + b];
}

/* inc_2stream_by_2stream_bybnd: rte/kernels/mo_optical_props_kernels.F90:453-485 */
ORC_API void orc_inc_2stream_by_2stream_bybnd(int ngpt, int nlay, int ncol, float* tau1, float* ssa1, float* g1,
                                              const float* tau2, const float* ssa2, const float* g2, int nbnd,
                                              const int* gpt_lims) {
  const float eps = 3.0f * FLT_MIN;
  for (size_t s = 0; s < (size_t)ncol * nlay; ++s)
    for (int b = 0; b < nbnd; ++b)
      for (int g = gpt_lims[2 * b] - 1; g <= gpt_lims[2 * b + 1] - 1; ++g) {
        size_t i = s * ngpt + g, j = s * nbnd + b;
        float tau12 = tau1[i] + tau2[j];
        float tauscat12 = tau1[i] * ssa1[i] + tau2[j] * ssa2[j];
        g1[i] = (tau1[i] * ssa1[i] * g1[i] + tau2[j] * ssa2[j] * g2[j]) / fmaxf(eps, tauscat12);
        ssa1[i] = tauscat12 / fmaxf(eps, tau12);
        tau1[i] = tau12;
      }
}

/* ------------------------------------------------------------------------------------------
 * Cloud optics from Pade approximants: extensions/cloud_optics/mo_cloud_optics.F90:476-493 (orders [2/3] for the
 * extinction, [2/2] for the co-albedo and the asymmetry, three size regimes), compute_all_from_pade :650-714,
 * pade_eval_1 :757-781, combination of liquid and ice :500-528.  Coefficients as stored in the files:
 * (ncoeff, nsizereg, nbnd) in C order = the reference's (nbnd, nsizereg, 0:m+n); the ice arrays for ONE roughness.
 * ------------------------------------------------------------------------------------------ */
static float pade_eval_1(int iband, int nbnd, int nrads, int m, int n, int irad, float re, const float* c) {
#define PC(i) c[((size_t)(i) * nrads + (irad - 1)) * nbnd + iband]
  float denom = PC(n + m);
  for (int i = n - 1 + m; i >= 1 + m; --i) denom = PC(i) + re * denom;
  denom = 1.0f + re * denom;
  float numer = PC(m);
  for (int i = m - 1; i >= 1; --i) numer = PC(i) + re * numer;
  numer = PC(0) + re * numer;
#undef PC
  return numer / denom;
}

static void pade_all(int nbnd, float wp, float re, const float* b_ext, const float* c_ext, const float* b_ssa,
                     const float* c_ssa, const float* b_asy, const float* c_asy, float* t, float* ts, float* tsg) {
  for (int ib = 0; ib < nbnd; ++ib) {
    int irad = (int)floorf((re - b_ext[1]) / b_ext[2]) + 2;
    if (irad > 3) irad = 3;
    float tt = wp * pade_eval_1(ib, nbnd, 3, 2, 3, irad, re, c_ext);
    irad = (int)floorf((re - b_ssa[1]) / b_ssa[2]) + 2;
    if (irad > 3) irad = 3;
    float tts = tt * (1.0f - fmaxf(0.0f, pade_eval_1(ib, nbnd, 3, 2, 2, irad, re, c_ssa)));
    irad = (int)floorf((re - b_asy[1]) / b_asy[2]) + 2;
    if (irad > 3) irad = 3;
    tsg[ib] = tts * pade_eval_1(ib, nbnd, 3, 2, 2, irad, re, c_asy);
    ts[ib] = tts;
    t[ib] = tt;
  }
}

ORC_API void orc_cloud_optics_pade(int ncol, int nlay, int nbnd, const float* clwp, const float* ciwp, const float* reliq,
                                   const float* reice, const float* extliq, const float* ssaliq, const float* asyliq,
                                   const float* extice, const float* ssaice, const float* asyice, const float* sizreg,
                                   int two_stream, float* tau, float* ssa, float* g) {
  /* sizreg: the six bound arrays of 4 back to back: extliq, ssaliq, asyliq, extice, ssaice, asyice */
  const float eps = FLT_EPSILON;
  float lt[64], lts[64], ltsg[64], it[64], its[64], itsg[64];
  for (size_t s = 0; s < (size_t)ncol * nlay; ++s) {
    if (clwp[s] > 0.0f) pade_all(nbnd, clwp[s], reliq[s], sizreg, extliq, sizreg + 4, ssaliq, sizreg + 8, asyliq, lt, lts, ltsg);
    else for (int b = 0; b < nbnd; ++b) lt[b] = lts[b] = ltsg[b] = 0.0f;
    if (ciwp[s] > 0.0f) pade_all(nbnd, ciwp[s], reice[s], sizreg + 12, extice, sizreg + 16, ssaice, sizreg + 20, asyice, it, its, itsg);
    else for (int b = 0; b < nbnd; ++b) it[b] = its[b] = itsg[b] = 0.0f;
    for (int b = 0; b < nbnd; ++b) {
      size_t i = s * nbnd + b;
      if (!two_stream) {
        tau[i] = (lt[b] - lts[b]) + (it[b] - its[b]);
      } else {
        float t = lt[b] + it[b], ts = lts[b] + its[b];
        g[i] = (ltsg[b] + itsg[b]) / fmaxf(eps, ts);
        ssa[i] = ts / fmaxf(eps, t);
        tau[i] = t;
      }
    }
  }
}

/* ------------------------------------------------------------------------------------------
 * McICA sampling: extensions/cloud_optics/mo_cloud_sampling.F90
 *   sampled_mask_max_ran :107-170, sampled_mask_exp_ran :176-286, draw_samples / apply_cloud_mask :38-101, 292-308.
 * The module still declares its masks and fields in the pre-fork (ncol,nlay,ngpt) order; the restatement keeps the
 * algorithm and uses this fork's layout throughout: randoms, mask, sampled fields (ngpt,nlay,ncol) = C [ncol][nlay][ngpt],
 * cloud_frac (nlay,ncol) = C [ncol][nlay], overlap_param C [ncol][nlay-1], by-band fields C [ncol][nlay][nbnd].
 * overlap_param == NULL selects maximum-random overlap.
 * ------------------------------------------------------------------------------------------ */
ORC_API void orc_sampled_mask(int ngpt, int nlay, int ncol, const float* randoms, const float* cloud_frac,
                              const float* overlap_param, unsigned char* cloud_mask) {
  float* local = (float*)malloc(sizeof(float) * ngpt);
  for (int icol = 0; icol < ncol; ++icol) {
    const float* cf = cloud_frac + (size_t)icol * nlay;
    unsigned char* m = cloud_mask + (size_t)icol * nlay * ngpt;
    const float* rn = randoms + (size_t)icol * nlay * ngpt;
    int fst = -1, lst = -1;
    for (int l = 0; l < nlay; ++l)
      if (cf[l] > 0.0f) { if (fst < 0) fst = l; lst = l; }
    memset(m, 0, (size_t)nlay * ngpt);
    if (fst < 0) continue;
    for (int g = 0; g < ngpt; ++g) {
      local[g] = rn[(size_t)fst * ngpt + g];
      m[(size_t)fst * ngpt + g] = local[g] > (1.0f - cf[fst]);
    }
    for (int l = fst + 1; l <= lst; ++l) {
      if (!(cf[l] > 0.0f)) continue;
      if (cf[l - 1] > 0.0f) {
        if (overlap_param) { /* exponential-random: correlated deviates :267-274 */
          float rho = overlap_param[(size_t)icol * (nlay - 1) + (l - 1)];
          for (int g = 0; g < ngpt; ++g)
            local[g] = rho * (local[g] - 0.5f) + sqrtf(1.0f - rho * rho) * (rn[(size_t)l * ngpt + g] - 0.5f) + 0.5f;
        } /* maximum-random: the same deviates :158 */
      } else {
        for (int g = 0; g < ngpt; ++g) local[g] = rn[(size_t)l * ngpt + g];
      }
      for (int g = 0; g < ngpt; ++g) m[(size_t)l * ngpt + g] = local[g] > (1.0f - cf[l]);
    }
  }
  free(local);
}

/* apply_cloud_mask :292-308 for one field */
ORC_API void orc_apply_cloud_mask(int ngpt, int nlay, int ncol, int nbnd, const int* gpt_lims, const unsigned char* cloud_mask,
                                  const float* input_field, float* sampled_field) {
  for (size_t s = 0; s < (size_t)ncol * nlay; ++s)
    for (int b = 0; b < nbnd; ++b)
      for (int g = gpt_lims[2 * b] - 1; g <= gpt_lims[2 * b + 1] - 1; ++g)
        sampled_field[s * ngpt + g] = cloud_mask[s * ngpt + g] ? input_field[s * nbnd + b] : 0.0f;
}

/* ------------------------------------------------------------------------------------------
 * Heating rates.
 *  orc_heating_rate      : extensions/mo_heating_rates.F90:26-54 semantics [K/s], cp_dry = 1004.64,
 *                          restated in this fork's (nlay+1,ncol) layout.
 *  orc_calc_heating_rate : examples/rrtmgp-nn-training/rrtmgp_lw_eval_nn_rfmip.F90:624-653 [K/day].
 * ------------------------------------------------------------------------------------------ */
ORC_API void orc_heating_rate(int ncol, int nlay, const float* flux_up, const float* flux_dn, const float* plev,
                              float* hr) {
  for (int icol = 0; icol < ncol; ++icol)
    for (int l = 0; l < nlay; ++l) {
      size_t a = (size_t)icol * (nlay + 1) + l;
      hr[(size_t)icol * nlay + l] = (flux_up[a + 1] - flux_up[a] - flux_dn[a + 1] + flux_dn[a]) * GRAV /
                                   (CP_DRY * (plev[a + 1] - plev[a]));
    }
}

ORC_API void orc_calc_heating_rate(int ncol, int nlay, const float* flux_up, const float* flux_dn,
                                   const float* pressure_hl, float* hr_K_day) {
  const float scaling = -(24.0f * 3600.0f * GRAV / 1004.0f);
  for (int icol = 0; icol < ncol; ++icol)
    for (int l = 0; l < nlay; ++l) {
      size_t a = (size_t)icol * (nlay + 1) + l;
      float net1 = flux_dn[a + 1] - flux_up[a + 1];
      float net0 = flux_dn[a] - flux_up[a];
      float dF = net1 - net0;
      float dP = pressure_hl[a + 1] - pressure_hl[a];
      hr_K_day[(size_t)icol * nlay + l] = scaling * dF / dP;
    }
}

/* ------------------------------------------------------------------------------------------
 * By-band fluxes: extensions/mo_fluxes_byband_kernels.F90.  Restated in this fork's layout (g-point / band
 * fastest: gpt_flux [ncol][nlev][ngpt] -> bnd_flux [ncol][nlev][nbnd]); the module itself still declares the
 * upstream (ncol,nlev,ngpt) order.  band_lims is (2,nbnd), 1-based inclusive.
 *  orc_sum_byband      : sum_byband :33-51 (first g-point, then += in g-point order)
 *  orc_net_byband_full : net_byband_full :56-78 (net = net + dn - up, left to right)
 *  orc_net_flux        : net_byband_precalc :80-86 and net_broadband_precalc (flux_net = flux_dn - flux_up)
 * ------------------------------------------------------------------------------------------ */
ORC_API void orc_sum_byband(int ncol, int nlev, int ngpt, int nbnd, const int* band_lims, const float* spectral_flux,
                            float* byband_flux) {
  for (size_t r = 0; r < (size_t)ncol * nlev; ++r)
    for (int ibnd = 0; ibnd < nbnd; ++ibnd) {
      float acc = spectral_flux[r * ngpt + band_lims[2 * ibnd] - 1];
      for (int igpt = band_lims[2 * ibnd]; igpt <= band_lims[2 * ibnd + 1] - 1; ++igpt) acc = acc + spectral_flux[r * ngpt + igpt];
      byband_flux[r * nbnd + ibnd] = acc;
    }
}

ORC_API void orc_net_byband_full(int ncol, int nlev, int ngpt, int nbnd, const int* band_lims, const float* spectral_flux_dn,
                                 const float* spectral_flux_up, float* byband_flux_net) {
  for (size_t r = 0; r < (size_t)ncol * nlev; ++r)
    for (int ibnd = 0; ibnd < nbnd; ++ibnd) {
      int igpt = band_lims[2 * ibnd] - 1;
      float acc = spectral_flux_dn[r * ngpt + igpt] - spectral_flux_up[r * ngpt + igpt];
      for (igpt = band_lims[2 * ibnd]; igpt <= band_lims[2 * ibnd + 1] - 1; ++igpt)
        acc = acc + spectral_flux_dn[r * ngpt + igpt] - spectral_flux_up[r * ngpt + igpt];
      byband_flux_net[r * nbnd + ibnd] = acc;
    }
}

ORC_API void orc_net_flux(size_t n, const float* flux_dn, const float* flux_up, float* flux_net) {
  for (size_t i = 0; i < n; ++i) flux_net[i] = flux_dn[i] - flux_up[i];
}

/* ------------------------------------------------------------------------------------------
 * compute_optimal_angles: rrtmgp/mo_gas_optics_rrtmgp.F90:1712-1758.  tau [ncol][nlay][ngpt], gpt2band 0-based,
 * optimal_angle_fit (2,nbnd) == C [nbnd][2]; output [ncol][ngpt] (the layout rte_lw's lw_Ds takes in this fork).
 * ------------------------------------------------------------------------------------------ */
ORC_API void orc_compute_optimal_angles(int ncol, int nlay, int ngpt, const int* gpt2band, const float* optimal_angle_fit,
                                        const float* tau, float* optimal_angles) {
  for (int icol = 0; icol < ncol; ++icol)
    for (int igpt = 0; igpt < ngpt; ++igpt) {
      float t = 0.0f;
      for (int ilay = 0; ilay < nlay; ++ilay) t = t + tau[((size_t)icol * nlay + ilay) * ngpt + igpt];
      const float trans_total = expf(-t);
      const int bnd = gpt2band[igpt];
      optimal_angles[(size_t)icol * ngpt + igpt] = optimal_angle_fit[2 * bnd] * trans_total + optimal_angle_fit[2 * bnd + 1];
    }
}

/* ------------------------------------------------------------------------------------------
 * set_solar_variability: rrtmgp/mo_gas_optics_rrtmgp.F90:1058-1095, followed by set_tsi :1097-1120 when have_tsi.
 * ------------------------------------------------------------------------------------------ */
ORC_API void orc_set_solar_variability(int ngpt, const float* solar_quiet, const float* solar_facular, const float* solar_sunspot,
                                       float mg_index, float sb_index, int have_tsi, float tsi, float* solar_source) {
  const float a_offset = 0.1495954f, b_offset = 0.00066696f;
  for (int igpt = 0; igpt < ngpt; ++igpt)
    solar_source[igpt] = solar_quiet[igpt] + (mg_index - a_offset) * solar_facular[igpt] + (sb_index - b_offset) * solar_sunspot[igpt];
  if (have_tsi) {
    float sum = 0.0f;
    for (int igpt = 0; igpt < ngpt; ++igpt) sum += solar_source[igpt];
    const float norm = 1.0f / sum;
    for (int igpt = 0; igpt < ngpt; ++igpt) solar_source[igpt] = solar_source[igpt] * tsi * norm;
  }
}
