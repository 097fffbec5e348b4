// sw_solver_2stream with the optional g-point fluxes (rte/kernels/mo_rte_solver_kernels.F90:541-692, save_gpt_flux;
// ty_fluxes_flexible%gpt_flux_up / gpt_flux_dn / gpt_flux_dn_dir) -- row N2 of the scope table's "next" list.
// Unlike sw_solver_v5 (one fused elimination + back-substitution) this kernel keeps the reference's own three sweeps --
// sw_two_stream_source :1366-1480 from the top down, `adding` :1526-1637 up and down again -- in ARRAY order, one thread
// per g-point, intermediates in a coalesced global scratch [array][layer][g-point]; it therefore doubles as an
// independent cross-check of the reformulated production kernel (tests).  General, not tuned.
#include "solver_common.cuh"
#include <algorithm>
#include <cfloat>

namespace rrnn {

struct SwGenParams {
  int ngpt, nlay, ncol, top_at_1, gp;
  const float *inc_flux, *inc_flux_dif, *tau, *ssa, *g, *mu0, *alb_dir, *alb_dif;
  float *flux_up, *flux_dn, *flux_dir, *gpt_up, *gpt_dn, *gpt_dir;
  float* scratch;
  size_t scratch_per_block;  // floats
};

__global__ void __launch_bounds__(256) sw_general_kernel(const SwGenParams p) {
  const int G = p.ngpt, L = p.nlay, GP = p.gp;
  const bool top = p.top_at_1 != 0, save = p.gpt_up != nullptr;
  const float k_min = 1.e-4f;  // :76-82 (single precision)
  float* sc = p.scratch + (size_t)blockIdx.x * p.scratch_per_block;
  float* Rdif = sc;
  float* Tdif = Rdif + (size_t)L * GP;
  float* sup = Tdif + (size_t)L * GP;
  float* sdn = sup + (size_t)L * GP;
  float* den = sdn + (size_t)L * GP;
  float* alb = den + (size_t)L * GP;            // [L+1]
  float* src = alb + (size_t)(L + 1) * GP;      // [L+1]
  float* rup_s = src + (size_t)(L + 1) * GP;
  float* rdn_s = rup_s + (size_t)(L + 1) * GP;
  float* rdr_s = rdn_s + (size_t)(L + 1) * GP;
  const int top_level = top ? 0 : L;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;

  for (int col = blockIdx.x; col < p.ncol; col += gridDim.x) {
    const size_t nl = (size_t)col * L * G, nv = (size_t)col * (L + 1) * G;
    float* rup = save ? p.gpt_up + nv : rup_s;
    float* rdn = save ? p.gpt_dn + nv : rdn_s;
    float* rdr = save ? p.gpt_dir + nv : rdr_s;
    const int rp = save ? G : GP;
    const float mu0 = p.mu0[col];
    const float mu0_inv = 1.0f / mu0;
    for (int g = threadIdx.x; g < G; g += blockDim.x) {
      // boundary conditions :589-590
      float dinc = p.inc_flux[(size_t)col * G + g] * mu0;
      rdr[(size_t)top_level * rp + g] = dinc;
      rdn[(size_t)top_level * rp + g] = p.inc_flux_dif ? p.inc_flux_dif[(size_t)col * G + g] : 0.0f;
      // ---- sw_two_stream_source :1366-1480, from the top of the atmosphere down
      for (int j = 0; j < L; ++j) {
        const int l = top ? j : L - 1 - j;           // array layer
        const int lev_out = top ? l + 1 : l;         // level the direct beam leaves through
        const size_t i = nl + (size_t)l * G + g, k_ = (size_t)l * GP + g;
        const float tau = p.tau[i], w0 = p.ssa[i], gg = p.g ? p.g[i] : 0.0f;
        const float Tnoscat = expf(-tau * mu0_inv);
        const float gamma1 = (8.0f - w0 * (5.0f + 3.0f * gg)) * .25f;
        const float gamma2 = 3.0f * (w0 * (1.0f - gg)) * .25f;
        const float gamma3 = (2.0f - 3.0f * mu0 * gg) * .25f;
        const float gamma4 = 1.0f - gamma3;
        const float alpha1 = gamma1 * gamma4 + gamma2 * gamma3;
        const float alpha2 = gamma1 * gamma3 + gamma2 * gamma4;
        const float k = sqrtf(fmaxf((gamma1 - gamma2) * (gamma1 + gamma2), k_min));
        const float e1 = expf(-tau * k);
        const float e2 = e1 * e1;
        const float k2e = 2.0f * k * e1;
        float RT = 1.0f / (k * (1.0f + e2) + gamma1 * (1.0f - e2));
        Rdif[k_] = RT * gamma2 * (1.0f - e2);
        Tdif[k_] = RT * 2.0f * k * e1;
        const float k_mu = k * mu0;
        const float k_mu2 = k_mu * k_mu;
        const float k_g3 = k * gamma3, k_g4 = k * gamma4;
        const float dd = (fabsf(1.0f - k_mu2) >= FLT_EPSILON) ? (1.0f - k_mu2) : FLT_EPSILON;
        RT = w0 * RT / dd;
        float Rdir = RT * ((1.0f - k_mu) * (alpha2 + k_g3) - (1.0f + k_mu) * (alpha2 - k_g3) * e2 - k2e * (gamma3 - alpha2 * mu0) * Tnoscat);
        float Tdir = RT * (k2e * (gamma4 + alpha1 * mu0) - Tnoscat * ((1.0f + k_mu) * (alpha1 + k_g4) - (1.0f - k_mu) * (alpha1 - k_g4) * e2));
        Rdir = fmaxf(0.0f, fminf(Rdir, (1.0f - Tnoscat)));
        Tdir = fmaxf(0.0f, fminf(Tdir, (1.0f - Tnoscat - Rdir)));
        sup[k_] = Rdir * dinc;
        sdn[k_] = Tdir * dinc;
        dinc = Tnoscat * dinc;
        rdr[(size_t)lev_out * rp + g] = dinc;
      }
      const float source_sfc = dinc * p.alb_dir[(size_t)col * G + g];
      const float alb_sfc = p.alb_dif[(size_t)col * G + g];
      // ---- adding :1526-1637
      if (top) {
        float a = alb_sfc, s = source_sfc;
        alb[(size_t)L * GP + g] = a; src[(size_t)L * GP + g] = s;
        for (int l = L - 1; l >= 0; --l) {
          const size_t i = (size_t)l * GP + g;
          const float R = Rdif[i], T = Tdif[i];
          const float d = 1.0f / (1.0f - R * a);
          den[i] = d;
          const float a_new = R + T * T * a * d;
          s = sup[i] + T * d * (s + a * sdn[i]);
          a = a_new;
          alb[i] = a; src[i] = s;
        }
        float dn = rdn[g];
        rup[g] = dn * a + s;
        for (int lev = 1; lev <= L; ++lev) {
          const size_t im = (size_t)(lev - 1) * GP + g, i = (size_t)lev * GP + g;
          dn = (Tdif[im] * dn + Rdif[im] * src[i] + sdn[im]) * den[im];
          rdn[(size_t)lev * rp + g] = dn;
          rup[(size_t)lev * rp + g] = dn * alb[i] + src[i];
        }
      } else {
        float a = alb_sfc, s = source_sfc;
        alb[g] = a; src[g] = s;
        for (int l = 0; l < L; ++l) {
          const size_t i = (size_t)l * GP + g;
          const float R = Rdif[i], T = Tdif[i];
          const float d = 1.0f / (1.0f - R * a);
          den[i] = d;
          const float a_new = R + T * T * a * d;
          s = sup[i] + T * d * (s + a * sdn[i]);
          a = a_new;
          alb[i + GP] = a; src[i + GP] = s;
        }
        float dn = rdn[(size_t)L * rp + g];
        rup[(size_t)L * rp + g] = dn * a + s;
        for (int l = L - 1; l >= 0; --l) {
          const size_t i = (size_t)l * GP + g;
          dn = (Tdif[i] * dn + Rdif[i] * src[i] + sdn[i]) * den[i];
          rdn[(size_t)l * rp + g] = dn;
          rup[(size_t)l * rp + g] = dn * alb[i] + src[i];
        }
      }
      // adding computes only the diffuse flux; the saved flux_dn is the total (:660-663)
      if (save)
        for (int lev = 0; lev <= L; ++lev) rdn[(size_t)lev * rp + g] = rdn[(size_t)lev * rp + g] + rdr[(size_t)lev * rp + g];
    }
    __syncthreads();
    // ---- broadband sums :643-680: warp per level, deterministic
    for (int lev = warp; lev <= L; lev += nwarps) {
      float su = 0.0f, sd = 0.0f, sr = 0.0f;
      for (int g = lane; g < G; g += 32) {
        const size_t i = (size_t)lev * rp + g;
        su += rup[i];
        sr += rdr[i];
        sd += save ? rdn[i] : (rdn[i] + rdr[i]);
      }
      su = warp_sum(su); sd = warp_sum(sd); sr = warp_sum(sr);
      if (lane == 0) {
        p.flux_up[(size_t)col * (L + 1) + lev] = su;
        p.flux_dn[(size_t)col * (L + 1) + lev] = sd;
        p.flux_dir[(size_t)col * (L + 1) + lev] = sr;
      }
    }
    __syncthreads();
  }
}

}  // namespace rrnn

using namespace rrnn;

extern "C" int rrnn_sw_solver_2stream_ext(rrnn_ctx_t* ctx, int ngpt, int nlay, int ncol, int top_at_1, const float* inc_flux_d,
                                          const float* inc_flux_dif_d, const float* tau_d, const float* ssa_d, const float* g_d,
                                          const float* mu0_d, const float* sfc_alb_dir_d, const float* sfc_alb_dif_d, float* flux_up_d,
                                          float* flux_dn_d, float* flux_dir_d, float* gpt_flux_up_d, float* gpt_flux_dn_d,
                                          float* gpt_flux_dir_d) {
  rrnn::NvtxRange nvtx_("sw_two_stream_source + adding");
  RRNN_CHECK(ctx, "rrnn_sw_solver_2stream_ext: null context");
  RRNN_CHECK(ngpt > 0 && nlay > 0 && ncol >= 0, "rrnn_sw_solver_2stream_ext: bad extents");
  RRNN_CHECK(inc_flux_d && tau_d && ssa_d && mu0_d && sfc_alb_dir_d && sfc_alb_dif_d && flux_up_d && flux_dn_d && flux_dir_d,
             "rrnn_sw_solver_2stream_ext: null argument");
  const int ngp = (gpt_flux_up_d != nullptr) + (gpt_flux_dn_d != nullptr) + (gpt_flux_dir_d != nullptr);
  RRNN_CHECK(ngp == 0 || ngp == 3, "rrnn_sw_solver_2stream_ext: the three g-point flux arrays come together");
  if (ncol == 0) return 0;
  RRNN_CUDA(cudaSetDevice(ctx->device));
  SwGenParams p{};
  p.ngpt = ngpt; p.nlay = nlay; p.ncol = ncol; p.top_at_1 = top_at_1 ? 1 : 0;
  p.gp = (ngpt + 31) & ~31;
  p.inc_flux = inc_flux_d; p.inc_flux_dif = inc_flux_dif_d; p.tau = tau_d; p.ssa = ssa_d; p.g = g_d; p.mu0 = mu0_d;
  p.alb_dir = sfc_alb_dir_d; p.alb_dif = sfc_alb_dif_d;
  p.flux_up = flux_up_d; p.flux_dn = flux_dn_d; p.flux_dir = flux_dir_d;
  p.gpt_up = gpt_flux_up_d; p.gpt_dn = gpt_flux_dn_d; p.gpt_dir = gpt_flux_dir_d;
  p.scratch_per_block = ((size_t)5 * nlay + (size_t)5 * (nlay + 1)) * p.gp;
  const int threads = std::min(p.gp, 256);
  const int blocks = std::min(ncol, ctx->num_sms * 4);
  if (int rc = ensure_scratch(ctx, (size_t)blocks * p.scratch_per_block * sizeof(float))) return rc;
  p.scratch = (float*)ctx->scratch;
  const int ps = prof_begin(ctx, K_SW_SOLVER);
  sw_general_kernel<<<blocks, threads, 0, ctx->stream>>>(p);
  prof_end(ctx, K_SW_SOLVER, ps);
  RRNN_LAUNCH_CHECK(ctx);
  return 0;
}
