// The remaining branches of rte_lw's dispatch table (rte/mo_rte_lw.F90:324-384) behind lw_solver_noscat_GaussQuad
// (rte/kernels/mo_rte_solver_kernels.F90:332-415, 119-330):
//   * re-scaled scattering for ty_optical_props_2str (do_rescaling, :179-181, :211-233, lw_transport_1rescl :1729-1795),
//   * per-g-point secants lw_Ds (one angle, mo_rte_lw.F90:329-340),
//   * g-point fluxes (ty_fluxes_flexible%gpt_flux_up/dn; for one angle they hold un-scaled radiances, quirk Q3 :287-291),
//   * the surface-temperature Jacobian of the upward flux (compute_Jac -- a compile-time .false. in this fork,
//     rte/mo_rte_rrtmgp_config.F90:29; restated as written, including :319 which sums the UN-scaled Jacobian radiances
//     when there is one angle).
// These are the "next" rows N1 / N2 of the scope table, not the benchmark path (that is rte_solvers_v5.cu): this kernel
// is written for generality -- one thread per g-point marching through the layers in ARRAY order exactly as the
// reference loops do (both orientations spelled out, because the re-scaled transport is not symmetric in them), all
// per-layer intermediates in a global scratch laid out [array][layer][g-point] (coalesced), one block per column,
// blocks persistent over columns; the broadband sums are a deterministic second phase (warp per level).
#include "solver_common.cuh"
#include <algorithm>

namespace rrnn {

struct LwGenParams {
  int ngpt, nlay, ncol, top_at_1, nmus, swap_lev, gp;  // gp: g-point pitch of the scratch arrays (multiple of 32)
  float Ds[4], wts[4];
  const float* Ds_gpt;  // (ngpt,ncol) or null
  const float *inc_flux, *tau, *ssa, *g, *lay_source, *lev_source, *sfc_emis, *sfc_source, *sfc_source_Jac;
  float *flux_up, *flux_dn, *flux_up_Jac, *gpt_up, *gpt_dn;
  float* scratch;
  size_t scratch_per_block;  // floats
};

__global__ void __launch_bounds__(256) lw_general_kernel(const LwGenParams p) {
  const int G = p.ngpt, L = p.nlay, GP = p.gp;
  const bool top = p.top_at_1 != 0, resc = p.ssa != nullptr, jac = p.flux_up_Jac != nullptr;
  const float tau_thresh = 3.4526698e-4f;  // sqrt(epsilon(1._sp)), :754
  float* sc = p.scratch + (size_t)blockIdx.x * p.scratch_per_block;
  // per-layer arrays [L][GP], then per-level arrays [L+1][GP]
  float* trans = sc;
  float* sdn = trans + (size_t)L * GP;
  float* sup = sdn + (size_t)L * GP;
  float* Cn = sup + (size_t)L * GP;
  float* rdn = Cn + (size_t)L * GP;
  float* rup = rdn + (size_t)(L + 1) * GP;
  float* rjc = rup + (size_t)(L + 1) * GP;
  float* acc_dn_s = rjc + (size_t)(L + 1) * GP;
  float* acc_up_s = acc_dn_s + (size_t)(L + 1) * GP;
  float* acc_jc = acc_up_s + (size_t)(L + 1) * GP;
  const int top_level = top ? 0 : L, sfc_level = top ? L : 0;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;

  for (int col = blockIdx.x; col < p.ncol; col += gridDim.x) {
    const size_t nl = (size_t)col * L * G, nv = (size_t)col * (L + 1) * G;
    // the g-point fluxes, when asked for, are exactly the accumulators: (ngpt, nlay+1) per column
    float* acc_up = p.gpt_up ? p.gpt_up + nv : acc_up_s;
    float* acc_dn = p.gpt_dn ? p.gpt_dn + nv : acc_dn_s;
    const int ap = p.gpt_up ? G : GP;  // pitch of the accumulators
    for (int g = threadIdx.x; g < G; g += blockDim.x) {
      const float emis = p.sfc_emis[(size_t)col * G + g], ssrc = p.sfc_source[(size_t)col * G + g];
      const float inc = p.inc_flux ? p.inc_flux[(size_t)col * G + g] : 0.0f;
      for (int imu = 0; imu < p.nmus; ++imu) {
        const float D = p.Ds_gpt ? p.Ds_gpt[(size_t)col * G + g] : p.Ds[imu];
        const float weight = p.wts[imu];
        const float fac = 2.0f * kPi * weight;
        // ---- optical path, transmission (:211-239), lw_source_noscat (:742-776)
        for (int l = 0; l < L; ++l) {
          const size_t i = nl + (size_t)l * G + g;
          float tl = p.tau[i] * D;
          if (resc) {
            const float ssal = p.ssa[i];
            const float wb = ssal * (1.0f - p.g[i]) * 0.5f;
            const float scaleTau = (1.0f - ssal + wb);
            Cn[(size_t)l * GP + g] = 0.4f * wb / scaleTau;
            tl = tl * scaleTau;  // tau*D*scaleTau, left to right as written
          }
          const float tr = expf(-tl);
          // 1 - exp(-x) without the cancellation of the literal form just above the series threshold (the reference's
          // (1-t)/tau - t carries a relative error of ~1 there in fp32; cf. exp_and_complement in common.cuh)
          const float omt = -expm1f(-tl);
          float fact;
          if (tl > tau_thresh) fact = omt / tl - tr;
          else fact = tl * (0.5f - 1.0f / 3.0f * tl);
          const float lay = p.lay_source[i];
          const float la = p.lev_source[nv + (size_t)l * G + g], lb = p.lev_source[nv + (size_t)(l + 1) * G + g];
          const float lev_dn = p.swap_lev ? la : lb, lev_up = p.swap_lev ? lb : la;  // quirk Q1 unless swap_lev
          trans[(size_t)l * GP + g] = tr;
          sdn[(size_t)l * GP + g] = omt * lev_dn + 2.0f * fact * (lay - lev_dn);
          sup[(size_t)l * GP + g] = omt * lev_up + 2.0f * fact * (lay - lev_up);
        }
        // ---- lw_transport_noscat_dn (:982-1009)
        float rd = inc / (2.0f * kPi * weight);
        rdn[(size_t)top_level * GP + g] = rd;
        if (top) {
          for (int lev = 1; lev <= L; ++lev) {
            rd = trans[(size_t)(lev - 1) * GP + g] * rd + sdn[(size_t)(lev - 1) * GP + g];
            rdn[(size_t)lev * GP + g] = rd;
          }
        } else {
          for (int lev = L - 1; lev >= 0; --lev) {
            rd = trans[(size_t)lev * GP + g] * rd + sdn[(size_t)lev * GP + g];
            rdn[(size_t)lev * GP + g] = rd;
          }
        }
        // ---- surface (:269-270)
        float ru = rd * (1 - emis) + emis * ssrc;
        float rj = jac ? emis * p.sfc_source_Jac[(size_t)col * G + g] : 0.0f;
        rup[(size_t)sfc_level * GP + g] = ru;
        if (jac) rjc[(size_t)sfc_level * GP + g] = rj;
        // ---- up (lw_transport_noscat_up :950-980, or lw_transport_1rescl :1729-1795: up with the adjustment, down again)
        if (top) {
          for (int l = L - 1; l >= 0; --l) {
            const size_t i = (size_t)l * GP + g;
            const float tr = trans[i];
            float adj = 0.0f;
            if (resc) adj = Cn[i] * ((1.0f - tr * tr) * rdn[i] - tr * sdn[i] - sup[i]);
            ru = resc ? tr * ru + sup[i] + adj : tr * ru + sup[i];
            rup[i] = ru;
            if (jac) { rj = tr * rj; rjc[i] = rj; }
          }
          if (resc) {
            rd = rdn[g];
            for (int l = 0; l < L; ++l) {
              const size_t i = (size_t)l * GP + g;
              const float tr = trans[i];
              const float adj = Cn[i] * ((1.0f - tr * tr) * rup[i] - tr * sup[i] - sdn[i]);
              rd = tr * rd + sdn[i] + adj;
              rdn[i + GP] = rd;
            }
          }
        } else {
          for (int l = 0; l < L; ++l) {
            const size_t i = (size_t)l * GP + g;
            const float tr = trans[i];
            float adj = 0.0f;
            if (resc) adj = Cn[i] * ((1.0f - tr * tr) * rdn[i + GP] - tr * sdn[i] - sup[i]);
            ru = resc ? tr * ru + sup[i] + adj : tr * ru + sup[i];
            rup[i + GP] = ru;
            if (jac) { rj = tr * rj; rjc[i + GP] = rj; }
          }
          if (resc) {
            rd = rdn[(size_t)L * GP + g];
            for (int l = L - 1; l >= 0; --l) {
              const size_t i = (size_t)l * GP + g;
              const float tr = trans[i];
              const float adj = Cn[i] * ((1.0f - tr * tr) * rup[i] - tr * sup[i] - sdn[i]);
              rd = tr * rd + sdn[i] + adj;
              rdn[i] = rd;
            }
          }
        }
        // ---- one angle: the radiances themselves (:287-317); several: sum over angles of fac * radiance (:383-412)
        for (int lev = 0; lev <= L; ++lev) {
          const size_t i = (size_t)lev * GP + g, a = (size_t)lev * ap + g;
          if (p.nmus == 1) {
            acc_up[a] = rup[i]; acc_dn[a] = rdn[i];
            if (jac) acc_jc[i] = rjc[i];
          } else if (imu == 0) {
            acc_up[a] = fac * rup[i]; acc_dn[a] = fac * rdn[i];
            if (jac) acc_jc[i] = fac * rjc[i];
          } else {
            acc_up[a] = acc_up[a] + fac * rup[i]; acc_dn[a] = acc_dn[a] + fac * rdn[i];
            if (jac) acc_jc[i] = acc_jc[i] + fac * rjc[i];
          }
        }
      }
    }
    __syncthreads();
    // ---- broadband sums (:301-319, sum_broadband): warp per level, lanes stride the g-points, butterfly at the end
    const float fac1 = (p.nmus == 1) ? 2.0f * kPi * p.wts[0] : 1.0f;
    for (int lev = warp; lev <= L; lev += nwarps) {
      float su = 0.0f, sd = 0.0f, sj = 0.0f;
      for (int g = lane; g < G; g += 32) {
        su += fac1 * acc_up[(size_t)lev * ap + g];
        sd += fac1 * acc_dn[(size_t)lev * ap + g];
        if (jac) sj += acc_jc[(size_t)lev * GP + g];  // :319: not scaled, as written
      }
      su = warp_sum(su); sd = warp_sum(sd);
      if (jac) sj = warp_sum(sj);
      if (lane == 0) {
        p.flux_up[(size_t)col * (L + 1) + lev] = su;
        p.flux_dn[(size_t)col * (L + 1) + lev] = sd;
        if (jac) p.flux_up_Jac[(size_t)col * (L + 1) + lev] = sj;
      }
    }
    __syncthreads();
  }
}

// lw_solver_2stream (:426-486): lw_two_stream :1018-1069 (LW_diff_sec = 1.66), lw_source_2str :1112-1162, adding :1526-1637 with
// sfc_albedo = 1 - sfc_emis and source_sfc = pi * sfc_emis * sfc_source; plain sums over g-points.  Same layout of the work
// as lw_general_kernel (thread = g-point, array order, coalesced scratch); reuses LwGenParams (Ds / wts / Jacobian unused).
__global__ void __launch_bounds__(256) lw_2stream_kernel(const LwGenParams p) {
  const int G = p.ngpt, L = p.nlay, GP = p.gp;
  const bool top = p.top_at_1 != 0;
  const float k_min = 1.e-4f, LW_diff_sec = 1.66f;
  float* sc = p.scratch + (size_t)blockIdx.x * p.scratch_per_block;
  float* Rdif = sc;
  float* Tdif = Rdif + (size_t)L * GP;
  float* sup = Tdif + (size_t)L * GP;
  float* sdn = sup + (size_t)L * GP;
  float* den = sdn + (size_t)L * GP;
  float* alb = den + (size_t)L * GP;
  float* src = alb + (size_t)(L + 1) * GP;
  float* rup_s = src + (size_t)(L + 1) * GP;
  float* rdn_s = rup_s + (size_t)(L + 1) * GP;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  for (int col = blockIdx.x; col < p.ncol; col += gridDim.x) {
    const size_t nl = (size_t)col * L * G, nv = (size_t)col * (L + 1) * G;
    float* rup = p.gpt_up ? p.gpt_up + nv : rup_s;
    float* rdn = p.gpt_dn ? p.gpt_dn + nv : rdn_s;
    const int rp = p.gpt_up ? G : GP;
    for (int g = threadIdx.x; g < G; g += blockDim.x) {
      for (int l = 0; l < L; ++l) {
        const size_t i = nl + (size_t)l * G + g, k_ = (size_t)l * GP + g;
        const float tau = p.tau[i], w0 = p.ssa[i], gg = p.g[i];
        const float gamma1 = LW_diff_sec * (1.0f - 0.5f * w0 * (1.0f + gg));  // Fu et al. Eq 2.9
        const float gamma2 = LW_diff_sec * 0.5f * w0 * (1.0f - gg);           // Eq 2.10
        const float k = sqrtf(fmaxf((gamma1 - gamma2) * (gamma1 + gamma2), k_min));
        const float e1 = expf(-tau * k);
        const float e2 = e1 * e1;
        const float RT = 1.0f / (k * (1.0f + e2) + gamma1 * (1.0f - e2));
        const float R = RT * gamma2 * (1.0f - e2), T = RT * 2.0f * k * e1;
        Rdif[k_] = R; Tdif[k_] = T;
        const float la = p.lev_source[nv + (size_t)l * G + g], lb = p.lev_source[nv + (size_t)(l + 1) * G + g];
        const float ltop = top ? la : lb, lbot = top ? lb : la;
        float su = 0.0f, sd = 0.0f;
        if (tau > 1.0e-8f) {
          const float Z = (lbot - ltop) / (tau * (gamma1 + gamma2));
          const float Zup_top = Z + ltop, Zup_bottom = Z + lbot, Zdn_top = -Z + ltop, Zdn_bottom = -Z + lbot;
          su = kPi * (Zup_top - R * Zdn_top - T * Zup_bottom);
          sd = kPi * (Zdn_bottom - R * Zup_bottom - T * Zdn_top);
        }
        sup[k_] = su; sdn[k_] = sd;
      }
      const float em = p.sfc_emis[(size_t)col * G + g];
      float a = 1.0f - em, s = kPi * em * p.sfc_source[(size_t)col * G + g];
      const float inc = p.inc_flux ? p.inc_flux[(size_t)col * G + g] : 0.0f;
      if (top) {
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
        float dn = inc;
        rdn[g] = dn;
        rup[g] = dn * a + s;
        for (int lev = 1; lev <= L; ++lev) {
          const size_t im = (size_t)(lev - 1) * GP + g, i = (size_t)lev * GP + g;
          dn = (Tdif[im] * dn + Rdif[im] * src[i] + sdn[im]) * den[im];
          rdn[(size_t)lev * rp + g] = dn;
          rup[(size_t)lev * rp + g] = dn * alb[i] + src[i];
        }
      } else {
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
        float dn = inc;
        rdn[(size_t)L * rp + g] = dn;
        rup[(size_t)L * rp + g] = dn * a + s;
        for (int l = L - 1; l >= 0; --l) {
          const size_t i = (size_t)l * GP + g;
          dn = (Tdif[i] * dn + Rdif[i] * src[i] + sdn[i]) * den[i];
          rdn[(size_t)l * rp + g] = dn;
          rup[(size_t)l * rp + g] = dn * alb[i] + src[i];
        }
      }
    }
    __syncthreads();
    for (int lev = warp; lev <= L; lev += nwarps) {
      float su = 0.0f, sd = 0.0f;
      for (int g = lane; g < G; g += 32) { su += rup[(size_t)lev * rp + g]; sd += rdn[(size_t)lev * rp + g]; }
      su = warp_sum(su); sd = warp_sum(sd);
      if (lane == 0) { p.flux_up[(size_t)col * (L + 1) + lev] = su; p.flux_dn[(size_t)col * (L + 1) + lev] = sd; }
    }
    __syncthreads();
  }
}

__global__ void expand_emis_kernel(int nbnd, int ngpt, int ncol, const int* __restrict__ gpt2band, const float* __restrict__ in,
                                   float* __restrict__ out) {  // expand, rte/mo_rte_lw.F90:429-447
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)ngpt * ncol) return;
  const size_t c = i / ngpt;
  out[i] = in[c * nbnd + gpt2band[i - c * ngpt]];
}

}  // namespace rrnn

using namespace rrnn;

extern "C" int rrnn_lw_solver_noscat_ext(rrnn_ctx_t* ctx, int ngpt, int nlay, int ncol, int top_at_1, int nmus, const float* Ds,
                                         const float* weights, const float* lw_Ds_gpt_d, const float* inc_flux_d, const float* tau_d,
                                         const float* ssa_d, const float* g_d, const float* lay_source_d, const float* lev_source_d,
                                         const float* sfc_emis_gpt_d, const float* sfc_source_d, const float* sfc_source_Jac_d,
                                         float* flux_up_d, float* flux_dn_d, float* flux_up_Jac_d, float* gpt_flux_up_d,
                                         float* gpt_flux_dn_d) {
  rrnn::NvtxRange nvtx_("lw_solver_noscat");
  RRNN_CHECK(ctx, "rrnn_lw_solver_noscat_ext: null context");
  RRNN_CHECK(ngpt > 0 && nlay > 0 && ncol >= 0, "rrnn_lw_solver_noscat_ext: bad extents");
  RRNN_CHECK(nmus >= 1 && nmus <= 4, "rte_lw: have to ask for between 1 and 4 quadrature points for no-scattering calculation");
  RRNN_CHECK(tau_d && lay_source_d && lev_source_d && sfc_emis_gpt_d && sfc_source_d && flux_up_d && flux_dn_d,
             "rrnn_lw_solver_noscat_ext: null argument");
  RRNN_CHECK((ssa_d == nullptr) == (g_d == nullptr), "rrnn_lw_solver_noscat_ext: ssa and g come together");
  RRNN_CHECK(!lw_Ds_gpt_d || nmus == 1, "rte_lw: providing lw_Ds incompatible with specifying n_gauss_angles");
  RRNN_CHECK(!flux_up_Jac_d || sfc_source_Jac_d, "rte_lw: compute_Jac=true but Jacobian arrays not provided");
  RRNN_CHECK((gpt_flux_up_d == nullptr) == (gpt_flux_dn_d == nullptr), "rrnn_lw_solver_noscat_ext: gpt_flux_up and gpt_flux_dn come together");
  if (ncol == 0) return 0;
  RRNN_CUDA(cudaSetDevice(ctx->device));
  LwGenParams p{};
  p.ngpt = ngpt; p.nlay = nlay; p.ncol = ncol; p.top_at_1 = top_at_1 ? 1 : 0; p.nmus = nmus;
  p.swap_lev = (!top_at_1 && !ctx->lw_source_bug_compat) ? 1 : 0;
  p.gp = (ngpt + 31) & ~31;
  for (int i = 0; i < nmus; ++i) { p.Ds[i] = Ds[i]; p.wts[i] = weights[i]; }
  p.Ds_gpt = lw_Ds_gpt_d; p.inc_flux = inc_flux_d; p.tau = tau_d; p.ssa = ssa_d; p.g = g_d;
  p.lay_source = lay_source_d; p.lev_source = lev_source_d; p.sfc_emis = sfc_emis_gpt_d; p.sfc_source = sfc_source_d;
  p.sfc_source_Jac = sfc_source_Jac_d; p.flux_up = flux_up_d; p.flux_dn = flux_dn_d; p.flux_up_Jac = flux_up_Jac_d;
  p.gpt_up = gpt_flux_up_d; p.gpt_dn = gpt_flux_dn_d;
  p.scratch_per_block = ((size_t)4 * nlay + (size_t)6 * (nlay + 1)) * p.gp;
  const int threads = std::min(p.gp, 256);
  const int blocks = std::min(ncol, ctx->num_sms * 4);
  if (int rc = ensure_scratch(ctx, (size_t)blocks * p.scratch_per_block * sizeof(float))) return rc;
  p.scratch = (float*)ctx->scratch;
  const int ps = prof_begin(ctx, K_LW_SOLVER);
  lw_general_kernel<<<blocks, threads, 0, ctx->stream>>>(p);
  prof_end(ctx, K_LW_SOLVER, ps);
  RRNN_LAUNCH_CHECK(ctx);
  return 0;
}

// rte_lw for ty_optical_props_1scl (ssa_d = g_d = NULL) or ty_optical_props_2str (re-scaled, :363-384) with the optional
// arguments of rte/mo_rte_lw.F90:60-64: lw_Ds (ngpt,ncol), flux_up_Jac, g-point fluxes.
extern "C" int rrnn_rte_lw_ext(rrnn_ctx_t* ctx, const rrnn_kdist_t* kd, int nlay, int ncol, int top_at_1, int n_gauss_angles,
                               const float* inc_flux_d, const float* tau_d, const float* ssa_d, const float* g_d,
                               const float* lay_source_d, const float* lev_source_d, const float* sfc_source_d,
                               const float* sfc_emis_d, const float* lw_Ds_d, const float* sfc_source_Jac_d, float* flux_up_d,
                               float* flux_dn_d, float* flux_up_Jac_d, float* gpt_flux_up_d, float* gpt_flux_dn_d) {
  rrnn::NvtxRange nvtx_("rte_lw");
  RRNN_CHECK(ctx && kd, "rte_lw: null handle");
  static const float gauss_Ds[4][4] = {{1.66f, 0.f, 0.f, 0.f},  // rte/mo_rte_lw.F90:113-125
                                       {1.18350343f, 2.81649655f, 0.f, 0.f},
                                       {1.09719858f, 1.69338507f, 4.70941630f, 0.f},
                                       {1.06056257f, 1.38282560f, 2.40148179f, 7.15513024f}};
  static const float gauss_wts[4][4] = {{0.5f, 0.f, 0.f, 0.f},
                                        {0.3180413817f, 0.1819586183f, 0.f, 0.f},
                                        {0.2009319137f, 0.2292411064f, 0.0698269799f, 0.f},
                                        {0.1355069134f, 0.2034645680f, 0.1298475476f, 0.0311809710f}};
  RRNN_CHECK(n_gauss_angles <= 4, "rte_lw: asking for too many quadrature points for no-scattering calculation");
  RRNN_CHECK(n_gauss_angles >= 1, "rte_lw: have to ask for at least one quadrature point for no-scattering calculation");
  RRNN_CHECK(!(lw_Ds_d && ssa_d), "rte_lw: lw_Ds not valid input for _2str class");
  RRNN_CHECK(!(lw_Ds_d && n_gauss_angles != 1), "rte_lw: providing lw_Ds incompatible with specifying n_gauss_angles");
  RRNN_CHECK(flux_up_d && flux_dn_d, "rte_lw: no space allocated for fluxes");
  if (ncol == 0) return 0;
  RRNN_CUDA(cudaSetDevice(ctx->device));
  const int ngpt = kd->ngpt;
  float* emis_gpt = nullptr;
  RRNN_CUDA(cudaMallocAsync((void**)&emis_gpt, (size_t)ngpt * ncol * sizeof(float), ctx->stream));
  const size_t n = (size_t)ngpt * ncol;
  expand_emis_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(kd->nbnd, ngpt, ncol, kd->d_gpt2band, sfc_emis_d, emis_gpt);
  RRNN_LAUNCH_CHECK(ctx);
  const int rc = rrnn_lw_solver_noscat_ext(ctx, ngpt, nlay, ncol, top_at_1, n_gauss_angles, gauss_Ds[n_gauss_angles - 1],
                                           gauss_wts[n_gauss_angles - 1], lw_Ds_d, inc_flux_d, tau_d, ssa_d, g_d, lay_source_d,
                                           lev_source_d, emis_gpt, sfc_source_d, sfc_source_Jac_d, flux_up_d, flux_dn_d, flux_up_Jac_d,
                                           gpt_flux_up_d, gpt_flux_dn_d);
  cudaFreeAsync(emis_gpt, ctx->stream);
  return rc;
}

extern "C" int rrnn_lw_solver_2stream(rrnn_ctx_t* ctx, int ngpt, int nlay, int ncol, int top_at_1, const float* inc_flux_d,
                                      const float* tau_d, const float* ssa_d, const float* g_d, const float* lev_source_d,
                                      const float* sfc_emis_gpt_d, const float* sfc_source_d, float* flux_up_d, float* flux_dn_d,
                                      float* gpt_flux_up_d, float* gpt_flux_dn_d) {
  rrnn::NvtxRange nvtx_("lw_solver_2stream");
  RRNN_CHECK(ctx, "rrnn_lw_solver_2stream: null context");
  RRNN_CHECK(ngpt > 0 && nlay > 0 && ncol >= 0, "rrnn_lw_solver_2stream: bad extents");
  RRNN_CHECK(tau_d && ssa_d && g_d && lev_source_d && sfc_emis_gpt_d && sfc_source_d && flux_up_d && flux_dn_d,
             "rrnn_lw_solver_2stream: null argument");
  RRNN_CHECK((gpt_flux_up_d == nullptr) == (gpt_flux_dn_d == nullptr), "rrnn_lw_solver_2stream: gpt_flux_up and gpt_flux_dn come together");
  if (ncol == 0) return 0;
  RRNN_CUDA(cudaSetDevice(ctx->device));
  LwGenParams p{};
  p.ngpt = ngpt; p.nlay = nlay; p.ncol = ncol; p.top_at_1 = top_at_1 ? 1 : 0;
  p.gp = (ngpt + 31) & ~31;
  p.inc_flux = inc_flux_d; p.tau = tau_d; p.ssa = ssa_d; p.g = g_d; p.lev_source = lev_source_d;
  p.sfc_emis = sfc_emis_gpt_d; p.sfc_source = sfc_source_d; p.flux_up = flux_up_d; p.flux_dn = flux_dn_d;
  p.gpt_up = gpt_flux_up_d; p.gpt_dn = gpt_flux_dn_d;
  p.scratch_per_block = ((size_t)5 * nlay + (size_t)4 * (nlay + 1)) * p.gp;
  const int threads = std::min(p.gp, 256);
  const int blocks = std::min(ncol, ctx->num_sms * 4);
  if (int rc = ensure_scratch(ctx, (size_t)blocks * p.scratch_per_block * sizeof(float))) return rc;
  p.scratch = (float*)ctx->scratch;
  const int ps = prof_begin(ctx, K_LW_SOLVER);
  lw_2stream_kernel<<<blocks, threads, 0, ctx->stream>>>(p);
  prof_end(ctx, K_LW_SOLVER, ps);
  RRNN_LAUNCH_CHECK(ctx);
  return 0;
}

// rte_lw(..., use_2stream = .true.) for ty_optical_props_2str (rte/mo_rte_lw.F90:346-361); sfc_emis_d is (nbnd,ncol)
extern "C" int rrnn_rte_lw_2stream(rrnn_ctx_t* ctx, const rrnn_kdist_t* kd, int nlay, int ncol, int top_at_1, const float* inc_flux_d,
                                   const float* tau_d, const float* ssa_d, const float* g_d, const float* lev_source_d,
                                   const float* sfc_source_d, const float* sfc_emis_d, float* flux_up_d, float* flux_dn_d,
                                   float* gpt_flux_up_d, float* gpt_flux_dn_d) {
  rrnn::NvtxRange nvtx_("rte_lw");
  RRNN_CHECK(ctx && kd, "rte_lw: null handle");
  RRNN_CHECK(flux_up_d && flux_dn_d, "rte_lw: no space allocated for fluxes");
  if (ncol == 0) return 0;
  RRNN_CUDA(cudaSetDevice(ctx->device));
  const int ngpt = kd->ngpt;
  float* emis_gpt = nullptr;
  RRNN_CUDA(cudaMallocAsync((void**)&emis_gpt, (size_t)ngpt * ncol * sizeof(float), ctx->stream));
  const size_t n = (size_t)ngpt * ncol;
  expand_emis_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(kd->nbnd, ngpt, ncol, kd->d_gpt2band, sfc_emis_d, emis_gpt);
  RRNN_LAUNCH_CHECK(ctx);
  const int rc = rrnn_lw_solver_2stream(ctx, ngpt, nlay, ncol, top_at_1, inc_flux_d, tau_d, ssa_d, g_d, lev_source_d, emis_gpt, sfc_source_d,
                                        flux_up_d, flux_dn_d, gpt_flux_up_d, gpt_flux_dn_d);
  cudaFreeAsync(emis_gpt, ctx->stream);
  return rc;
}
