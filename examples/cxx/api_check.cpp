// Self-checking program for the parts of the C++ mirror (include/rrnn.hpp) the RFMIP driver does not use: ty_fluxes_byband%reduce,
// flux_net, compute_optimal_angles, set_solar_variability, and the error strings of the reference.  Exact expectations are
// chosen so that no oracle is needed: sums of small integers, a transparent column, the quiet-sun offsets.  Exit code 0 = all ok.
#include <cmath>
#include <cstdio>
#include <iostream>

#include "rrnn.hpp"

using namespace rrtmgp_nn;

static int failures = 0;
#define CHECK(cond)                                                        \
  do {                                                                     \
    if (!(cond)) { std::printf("FAILED %s:%d  %s\n", __FILE__, __LINE__, #cond); ++failures; } \
  } while (0)

int main() {
  context ctx(0);
  if (!ctx.error().empty()) { std::cerr << ctx.error() << std::endl; return 1; }
  const int nbnd = 3, ngpt = 12, ncol = 5, nlay = 7, nlev = nlay + 1;
  const int band_lims[6] = {1, 4, 5, 8, 9, 12};
  // ---- longwave k-distribution with an optimal-angle fit
  std::vector<float> totplnk(nbnd * 4, 1.0f);
  ty_gas_optics_rrtmgp k_lw;
  CHECK(k_lw.load(ctx, nbnd, ngpt, band_lims, 4, totplnk.data(), 160.f, 1.f, nullptr).empty());
  ty_optical_props_1scl op;
  CHECK(op.alloc_1scl(ncol, nlay, k_lw).empty());
  dev_array angles(static_cast<size_t>(ncol) * ngpt);
  CHECK(k_lw.compute_optimal_angles(op, angles).find("no optimal_angle_fit") != std::string::npos);
  const float fit[6] = {0.25f, 1.5f, 0.5f, 1.25f, 0.125f, 1.75f};   // (2,nbnd)
  CHECK(k_lw.load_optimal_angle_fit(fit).empty());
  std::vector<float> tau(static_cast<size_t>(ncol) * nlay * ngpt, 0.0f);   // transparent: exp(-0) = 1 -> fit(1) + fit(2), exactly
  op.tau.from_host(tau.data(), tau.size());
  CHECK(k_lw.compute_optimal_angles(op, angles).empty());
  std::vector<float> a(angles.size());
  angles.to_host(a.data());
  for (int c = 0; c < ncol; ++c)
    for (int g = 0; g < ngpt; ++g) CHECK(a[c * ngpt + g] == fit[2 * (g / 4)] + fit[2 * (g / 4) + 1]);
  dev_array wrong(7);
  CHECK(k_lw.compute_optimal_angles(op, wrong) == "gas_optics%compute_optimal_angles: optimal_angles different dimension (ncol)");
  // ---- by-band and net fluxes of small integers (exact in fp32)
  std::vector<float> up(static_cast<size_t>(ncol) * nlev * ngpt), dn(up.size());
  for (size_t i = 0; i < up.size(); ++i) { up[i] = static_cast<float>(i % 7); dn[i] = static_cast<float>((i * 3) % 11); }
  dev_array d_up, d_dn, b_up(static_cast<size_t>(ncol) * nlev * nbnd), b_dn(b_up.size()), b_net(b_up.size()), b_net2(b_up.size());
  d_up.from_host(up.data(), up.size()); d_dn.from_host(dn.data(), dn.size());
  ty_fluxes_byband fl;
  fl.bnd_flux_up = b_up.data(); fl.bnd_flux_dn = b_dn.data(); fl.bnd_flux_net = b_net.data();
  CHECK(fl.reduce(ctx, d_up.data(), d_dn.data(), k_lw, ncol, nlev).empty());
  ty_fluxes_byband only_net;
  only_net.bnd_flux_net = b_net2.data();
  CHECK(only_net.reduce(ctx, d_up.data(), d_dn.data(), k_lw, ncol, nlev).empty());
  ty_fluxes_byband bad;
  bad.bnd_flux_dn_dir = b_net2.data();
  CHECK(bad.reduce(ctx, d_up.data(), d_dn.data(), k_lw, ncol, nlev) == "reduce: requesting bnd_flux_dn_dir but direct flux hasn't been supplied");
  std::vector<float> hu(b_up.size()), hd(hu.size()), hn(hu.size()), hn2(hu.size());
  b_up.to_host(hu.data()); b_dn.to_host(hd.data()); b_net.to_host(hn.data()); b_net2.to_host(hn2.data());
  for (int r = 0; r < ncol * nlev; ++r)
    for (int b = 0; b < nbnd; ++b) {
      float su = 0, sd = 0;
      for (int g = 4 * b; g < 4 * b + 4; ++g) { su += up[static_cast<size_t>(r) * ngpt + g]; sd += dn[static_cast<size_t>(r) * ngpt + g]; }
      CHECK(hu[r * nbnd + b] == su); CHECK(hd[r * nbnd + b] == sd); CHECK(hn[r * nbnd + b] == sd - su); CHECK(hn2[r * nbnd + b] == sd - su);
    }
  // ---- solar variability: the offsets are the quiet sun; set_tsi fixes the integral; the reference's range errors
  std::vector<float> quiet(ngpt), fac(ngpt), spot(ngpt);
  for (int g = 0; g < ngpt; ++g) { quiet[g] = 1.0f + g; fac[g] = 0.5f; spot[g] = -0.25f; }
  ty_gas_optics_rrtmgp k_sw;
  CHECK(k_sw.load(ctx, nbnd, ngpt, band_lims, 0, nullptr, 0.f, 1.f, quiet.data()).empty());
  CHECK(k_sw.set_solar_variability(0.15f, 0.001f).find("no solar variability tables") != std::string::npos);
  CHECK(k_sw.load_solar_tables(quiet.data(), fac.data(), spot.data()).empty());
  CHECK(k_sw.set_solar_variability(0.1495954f, 0.00066696f).empty());
  CHECK(k_sw.get_solar_source() == quiet);
  CHECK(k_sw.set_solar_variability(0.1495954f + 0.5f, 0.00066696f, 100.0f).empty());   // + 0.25 per g-point, then scaled to 100 W m-2
  double sum = 0;
  for (float v : k_sw.get_solar_source()) sum += v;
  CHECK(std::fabs(sum - 100.0) < 1e-3);
  {  // ty_solar_var: a four-point mean cycle (end points + two month centres at 0.25 and 0.75) -- exact expectations
    const float tab[8] = {1.f, 10.f, 2.f, 20.f, 4.f, 40.f, 8.f, 80.f};   // [nsolarfrac][2]: mg, sb
    rrtmgp_nn::ty_solar_var sv;
    float mg = -1.f, sb = -1.f;
    CHECK(sv.solar_var_ind_interp(0.5f, mg, sb).empty() && mg == -1.f);   // no table: nothing computed, no message
    CHECK(sv.load(tab, 4).empty());
    CHECK(sv.solar_var_ind_interp(0.f, mg, sb).empty() && mg == 1.f && sb == 10.f);
    CHECK(sv.solar_var_ind_interp(1.f, mg, sb).empty() && mg == 8.f && sb == 80.f);
    CHECK(sv.solar_var_ind_interp(0.125f, mg, sb).empty() && mg == 1.5f && sb == 15.f);   // half way through the first half interval
    CHECK(sv.solar_var_ind_interp(0.5f, mg, sb).empty() && mg == 3.f && sb == 30.f);      // half way between the month centres
    CHECK(sv.solar_var_ind_interp(0.875f, mg, sb).empty() && mg == 6.f && sb == 60.f);
    CHECK(sv.solar_var_ind_interp(1.5f, mg, sb) == "solar_var_ind_interp: solcycfrac out of range");
    CHECK(sv.solar_var_ind_interp(0.5f, mg, sb).empty() && k_sw.set_solar_variability(mg, sb * 1e-4f).empty());
  }
  CHECK(k_sw.set_solar_variability(-1.f, 0.001f) == "mg_index out of range");
  CHECK(k_sw.set_solar_variability(-1.f, -0.001f) == "sb_index out of range");
  CHECK(k_sw.set_tsi(-5.f) == "tsi out of range");
  // ---- rte_lw / rte_sw argument checks
  ty_source_func_lw src;
  CHECK(src.alloc(ncol, nlay, k_lw).empty());
  ty_fluxes_broadband none;
  CHECK(rte_lw(ctx, op, true, src, nullptr, none) == "rte_lw: no space allocated for fluxes");
  ty_fluxes_broadband some;
  dev_array f1(static_cast<size_t>(ncol) * nlev), f2(f1.size());
  some.flux_up = f1.data(); some.flux_dn = f2.data();
  CHECK(rte_lw(ctx, op, true, src, nullptr, some, nullptr, 5) == "rte_lw: asking for too many quadrature points for no-scattering calculation");
  ctx.synchronize();
  std::printf(failures ? "api_check: %d check(s) FAILED\n" : "api_check: all checks passed\n", failures);
  return failures ? 1 : 0;
}
