// Compiled host code on the C++ mirror (include/rrnn.hpp): the block loops of the reference's RFMIP drivers,
//   examples/rfmip-clear-sky/rrtmgp_rfmip_lw.F90:368-446   gas_optics(neural_nets=) -> rte_lw   per block of columns
//   examples/rfmip-clear-sky/rrtmgp_rfmip_sw.F90:356-465   gas_optics(neural_nets=) -> TSI renormalisation :409-416 -> rte_sw
//                                                          -> night columns zeroed :458-463
// written the way the Fortran drivers are: derived types allocated once per block size, stop_on_err on every call.
//
//   rfmip_driver lw|sw <case dir> <net file 1> <net file 2>
// <case dir> holds raw little-endian arrays written by the caller (tests/test_cxx_host_gpu.py): meta.txt (key value lines),
// band_lims.i32, totplnk.f32 | solar_source.f32, play/plev/tlay/tlev/tsfc/sfc_emis/mu0/sfc_alb/tsi/usecol .f32, gas_<name>.f32
// (nlay,ncol) fields; scalar gases are "gas <name> <vmr>" lines of meta.txt.  Fluxes are written back as flux_up.f32, flux_dn.f32.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <map>
#include <sstream>

#include "rrnn.hpp"

using namespace rrtmgp_nn;

static void stop_on_err(const std::string& error_msg) {  // rrtmgp_rfmip_lw.F90:31-41
  if (!error_msg.empty()) {
    std::cerr << error_msg << "\nrrtmgp_rfmip stopping" << std::endl;
    std::exit(1);
  }
}

template <typename T>
static std::vector<T> read_raw(const std::string& path, size_t expect = 0) {
  std::ifstream f(path, std::ios::binary | std::ios::ate);
  if (!f) stop_on_err("cannot open " + path);
  const size_t bytes = static_cast<size_t>(f.tellg());
  std::vector<T> v(bytes / sizeof(T));
  f.seekg(0);
  f.read(reinterpret_cast<char*>(v.data()), static_cast<std::streamsize>(bytes));
  if (expect && v.size() != expect) stop_on_err(path + ": unexpected size");
  return v;
}

static void write_raw(const std::string& path, const std::vector<float>& v) {
  std::ofstream f(path, std::ios::binary);
  f.write(reinterpret_cast<const char*>(v.data()), static_cast<std::streamsize>(v.size() * sizeof(float)));
}

int main(int argc, char** argv) {
  if (argc < 5) { std::cerr << "usage: rfmip_driver lw|sw <case dir> <net 1> <net 2>\n"; return 2; }
  const bool lw = std::string(argv[1]) == "lw";
  const std::string dir = std::string(argv[2]) + "/";
  std::map<std::string, double> meta;
  std::vector<std::pair<std::string, float>> scalar_gases;
  std::vector<std::string> field_gases;
  {
    std::ifstream f(dir + "meta.txt");
    if (!f) stop_on_err("cannot open " + dir + "meta.txt");
    std::string line, key;
    while (std::getline(f, line)) {
      std::istringstream ss(line);
      ss >> key;
      if (key == "gas") { std::string n; float v; ss >> n >> v; scalar_gases.emplace_back(n, v); }
      else if (key == "gasfield") { std::string n; ss >> n; field_gases.push_back(n); }
      else { double v; ss >> v; meta[key] = v; }
    }
  }
  const int ncol = static_cast<int>(meta["ncol"]), nlay = static_cast<int>(meta["nlay"]), block_size = static_cast<int>(meta["block_size"]);
  const int nbnd = static_cast<int>(meta["nbnd"]), ngpt = static_cast<int>(meta["ngpt"]);
  const bool top_at_1 = meta["top_at_1"] != 0;
  const size_t nlev = static_cast<size_t>(nlay) + 1;

  context ctx(0);
  stop_on_err(ctx.error());
  // ---- load: spectral tables and the two networks (rrtmgp_rfmip_lw.F90:258-276, neural nets :277-283)
  ty_gas_optics_rrtmgp k_dist;
  const auto band_lims = read_raw<int>(dir + "band_lims.i32", 2 * static_cast<size_t>(nbnd));
  std::vector<float> totplnk, solar;
  if (lw) totplnk = read_raw<float>(dir + "totplnk.f32", static_cast<size_t>(nbnd) * static_cast<size_t>(meta["ntemp"]));
  else solar = read_raw<float>(dir + "solar_source.f32", static_cast<size_t>(ngpt));
  stop_on_err(k_dist.load(ctx, nbnd, ngpt, band_lims.data(), lw ? static_cast<int>(meta["ntemp"]) : 0, lw ? totplnk.data() : nullptr,
                          static_cast<float>(meta["temp_ref_min"]), static_cast<float>(meta["totplnk_delta"]), lw ? nullptr : solar.data()));
  rrtmgp_network_type net1, net2;
  stop_on_err(net1.load_netcdf(ctx, argv[3]));
  stop_on_err(net2.load_netcdf(ctx, argv[4]));
  const std::vector<const rrtmgp_network_type*> neural_nets = {&net1, &net2};

  // ---- inputs, (nlay,ncol) == [ncol][nlay]
  const size_t nl = static_cast<size_t>(nlay);
  const auto play = read_raw<float>(dir + "play.f32", ncol * nl), plev = read_raw<float>(dir + "plev.f32", ncol * nlev);
  const auto tlay = read_raw<float>(dir + "tlay.f32", ncol * nl);
  std::vector<float> tlev, tsfc, sfc_emis, mu0, sfc_alb, tsi, usecol;
  if (lw) { tlev = read_raw<float>(dir + "tlev.f32", ncol * nlev); tsfc = read_raw<float>(dir + "tsfc.f32", ncol); sfc_emis = read_raw<float>(dir + "sfc_emis.f32", ncol); }
  else { mu0 = read_raw<float>(dir + "mu0.f32", ncol); sfc_alb = read_raw<float>(dir + "sfc_alb.f32", ncol); tsi = read_raw<float>(dir + "tsi.f32", ncol);
         usecol = read_raw<float>(dir + "usecol.f32", ncol); }
  std::map<std::string, std::vector<float>> gas_fields;
  for (const auto& n : field_gases) gas_fields[n] = read_raw<float>(dir + "gas_" + n + ".f32", ncol * nl);

  std::vector<float> flux_up(ncol * nlev), flux_dn(ncol * nlev);
  float def_tsi = 0.f;
  for (float v : solar) def_tsi += v;  // sum(toa_flux(:,1)), rrtmgp_rfmip_sw.F90:411

  // ---- the block loop (:368-446 / :356-465): device arrays sized for one block, allocated once
  ty_optical_props_1scl optical_props_lw;
  ty_optical_props_2str optical_props_sw;
  ty_source_func_lw source;
  dev_array d_play, d_plev, d_tlay, d_tlev, d_tsfc, d_emis, d_mu0, d_alb, d_toa, d_up(block_size * nlev), d_dn(block_size * nlev), d_dir(block_size * nlev);
  std::map<std::string, dev_array> d_gas;
  for (int b0 = 0; b0 < ncol; b0 += block_size) {
    const int nb = std::min(block_size, ncol - b0);
    if (lw) { if (optical_props_lw.get_ncol() != nb) { stop_on_err(optical_props_lw.alloc_1scl(nb, nlay, k_dist)); stop_on_err(source.alloc(nb, nlay, k_dist)); } }
    else if (optical_props_sw.get_ncol() != nb) stop_on_err(optical_props_sw.alloc_2str(nb, nlay, k_dist));
    d_play.from_host(play.data() + b0 * nl, nb * nl); d_plev.from_host(plev.data() + b0 * nlev, nb * nlev); d_tlay.from_host(tlay.data() + b0 * nl, nb * nl);
    ty_gas_concs gas_concs;
    for (const auto& sg : scalar_gases) stop_on_err(gas_concs.set_vmr(sg.first, sg.second));
    for (const auto& gf : gas_fields) { d_gas[gf.first].from_host(gf.second.data() + b0 * nl, nb * nl); stop_on_err(gas_concs.set_vmr(gf.first, d_gas[gf.first].data())); }
    ty_fluxes_broadband fluxes;
    fluxes.flux_up = d_up.data(); fluxes.flux_dn = d_dn.data();
    if (lw) {
      d_tlev.from_host(tlev.data() + b0 * nlev, nb * nlev); d_tsfc.from_host(tsfc.data() + b0, nb);
      std::vector<float> emis_spec(static_cast<size_t>(nb) * nbnd);  // sfc_emis_spec(nbnd, block_size), :357-362
      for (int i = 0; i < nb; ++i) std::fill_n(emis_spec.begin() + static_cast<size_t>(i) * nbnd, nbnd, sfc_emis[b0 + i]);
      d_emis.from_host(emis_spec.data(), emis_spec.size());
      stop_on_err(k_dist.gas_optics(d_play.data(), d_plev.data(), d_tlay.data(), d_tsfc.data(), gas_concs, optical_props_lw, source, d_tlev.data(), neural_nets));
      stop_on_err(rte_lw(ctx, optical_props_lw, top_at_1, source, d_emis.data(), fluxes, nullptr, static_cast<int>(meta["n_quad_angles"])));
    } else {
      d_toa.resize(static_cast<size_t>(nb) * ngpt);
      stop_on_err(k_dist.gas_optics(d_play.data(), d_plev.data(), d_tlay.data(), gas_concs, optical_props_sw, d_toa, neural_nets));
      // TSI renormalisation of the driver (:409-416), on the host as the Fortran does it
      std::vector<float> toa(d_toa.size());
      d_toa.to_host(toa.data());
      for (int i = 0; i < nb; ++i)
        for (int g = 0; g < ngpt; ++g) toa[static_cast<size_t>(i) * ngpt + g] = toa[static_cast<size_t>(i) * ngpt + g] * tsi[b0 + i] / def_tsi;
      d_toa.from_host(toa.data(), toa.size());
      std::vector<float> alb_spec(static_cast<size_t>(nb) * ngpt);      // sfc_alb_spec(ngpt, block_size), :419-423
      for (int i = 0; i < nb; ++i) std::fill_n(alb_spec.begin() + static_cast<size_t>(i) * ngpt, ngpt, sfc_alb[b0 + i]);
      d_alb.from_host(alb_spec.data(), alb_spec.size());
      d_mu0.from_host(mu0.data() + b0, nb);
      fluxes.flux_dn_dir = d_dir.data();
      stop_on_err(rte_sw(ctx, optical_props_sw, top_at_1, d_mu0.data(), d_toa.data(), d_alb.data(), d_alb.data(), fluxes));
    }
    ctx.synchronize();
    cudaMemcpy(flux_up.data() + b0 * nlev, d_up.data(), nb * nlev * sizeof(float), cudaMemcpyDeviceToHost);
    cudaMemcpy(flux_dn.data() + b0 * nlev, d_dn.data(), nb * nlev * sizeof(float), cudaMemcpyDeviceToHost);
    if (!lw)
      for (int i = 0; i < nb; ++i)
        if (usecol[b0 + i] == 0.f) {  // zero out fluxes for which the original solar zenith angle is > 90 degrees, :458-463
          std::fill_n(flux_up.begin() + (b0 + i) * nlev, nlev, 0.f);
          std::fill_n(flux_dn.begin() + (b0 + i) * nlev, nlev, 0.f);
        }
  }
  write_raw(dir + "flux_up.f32", flux_up);
  write_raw(dir + "flux_dn.f32", flux_dn);
  double mean_dn = 0;
  for (float v : flux_dn) mean_dn += v;
  std::printf("mean of flux_down is: %.4f\n", mean_dn / static_cast<double>(flux_dn.size()));  // rrtmgp_rfmip_lw.F90:480
  return 0;
}
