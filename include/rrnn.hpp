// rrnn.hpp -- C++ host mirror of the reference's Fortran interface for the NN gas optics + RTE path, header-only, on
// top of the C ABI (rrnn.h).  The reference's host code is compiled Fortran and no Fortran compiler exists in the build
// image, so this is the compiled host side: same type and procedure names, same argument meaning, and the reference's
// error convention (every procedure returns the error message, empty on success).
//
//   rrtmgp_network_type        neural/mod_network_rrtmgp.F90:34-122          (load_netcdf)
//   ty_gas_concs               rrtmgp/mo_gas_concentrations.F90:50-250       (set_vmr scalar / (nlay,ncol))
//   ty_optical_props[_1scl|_2str]  rte/mo_optical_props.F90:62-190           (alloc_1scl / alloc_2str, get_ncol ...)
//   ty_source_func_lw          rte/mo_source_functions.F90:27-80             (alloc)
//   ty_fluxes_broadband        rte/mo_fluxes.F90:36-50                        (caller-associated outputs)
//   ty_fluxes_byband           extensions/mo_fluxes_byband.F90:31-131         (reduce: by-band sums of g-point fluxes)
//   ty_gas_optics_rrtmgp       rrtmgp/mo_gas_optics_rrtmgp.F90:239-243, 433-438, 1058-1120, 1712-1758 (gas_optics with neural_nets,
//                              set_tsi, set_solar_variability, compute_optimal_angles)
//   rte_lw / rte_sw            rte/mo_rte_lw.F90:60-64, rte/mo_rte_sw.F90:48-52
//   ty_solar_var               extensions/solar_variability/mo_solar_variability.F90:20-183 (load, solar_var_ind_interp)
//
// Arrays live on the device (dev_array: cudaMalloc'd float storage with host <-> device copies); layouts are the
// reference's (g-point fastest): tau (ngpt,nlay,ncol) == C [ncol][nlay][ngpt], profiles (nlay,ncol) == C [ncol][nlay].
#ifndef RRNN_HPP
#define RRNN_HPP

#include <cuda_runtime_api.h>

#include <cstring>
#include <memory>
#include <string>
#include <utility>
#include <vector>

#include "rrnn.h"

namespace rrtmgp_nn {

inline std::string err(int rc) { return rc == 0 ? std::string() : std::string(rrnn_last_error()); }

// float storage on the device
class dev_array {
 public:
  dev_array() = default;
  explicit dev_array(size_t n) { resize(n); }
  dev_array(const dev_array&) = delete;
  dev_array& operator=(const dev_array&) = delete;
  dev_array(dev_array&& o) noexcept : p_(o.p_), n_(o.n_) { o.p_ = nullptr; o.n_ = 0; }
  dev_array& operator=(dev_array&& o) noexcept { std::swap(p_, o.p_); std::swap(n_, o.n_); return *this; }
  ~dev_array() { if (p_) cudaFree(p_); }
  void resize(size_t n) {
    if (n == n_) return;
    if (p_) cudaFree(p_);
    p_ = nullptr; n_ = 0;
    if (n && cudaMalloc(reinterpret_cast<void**>(&p_), n * sizeof(float)) == cudaSuccess) n_ = n;
  }
  // blocking copies on the legacy default stream: ordered with the kernels of a default-stream context
  void from_host(const float* h, size_t n) { resize(n); if (n_) cudaMemcpy(p_, h, n * sizeof(float), cudaMemcpyHostToDevice); }
  void to_host(float* h) const { if (n_) cudaMemcpy(h, p_, n_ * sizeof(float), cudaMemcpyDeviceToHost); }
  float* data() { return p_; }
  const float* data() const { return p_; }
  size_t size() const { return n_; }

 private:
  float* p_ = nullptr;
  size_t n_ = 0;
};

class context {
 public:
  explicit context(int device = 0) { msg_ = err(rrnn_ctx_create(device, nullptr, &h_)); }
  context(const context&) = delete;
  ~context() { if (h_) rrnn_ctx_destroy(h_); }
  const std::string& error() const { return msg_; }  // "no CUDA device available (this library has no CPU fallback)" ...
  rrnn_ctx_t* h() const { return h_; }
  void synchronize() const { rrnn_ctx_synchronize(h_); }

 private:
  rrnn_ctx_t* h_ = nullptr;
  std::string msg_;
};

class rrtmgp_network_type {
 public:
  rrtmgp_network_type() = default;
  rrtmgp_network_type(const rrtmgp_network_type&) = delete;
  rrtmgp_network_type(rrtmgp_network_type&& o) noexcept : h_(o.h_) { o.h_ = nullptr; }
  ~rrtmgp_network_type() { if (h_) rrnn_model_destroy(h_); }
  std::string load_netcdf(const context& ctx, const std::string& filename) { return err(rrnn_model_load_netcdf(ctx.h(), filename.c_str(), &h_)); }
  const rrnn_model_t* h() const { return h_; }

 private:
  rrnn_model_t* h_ = nullptr;
};

class ty_gas_concs {
 public:
  std::string set_vmr(const std::string& gas, float w) {
    if (w < 0.f || w > 1.f) return "ty_gas_concs%set_vmr: concentrations should be >= 0, <= 1";
    rrnn_gas_t g{};
    std::strncpy(g.name, gas.c_str(), sizeof(g.name) - 1);
    g.value = w; g.ndims = 0;
    put(g);
    return "";
  }
  // w: device array (nlay,ncol), owned by the caller
  std::string set_vmr(const std::string& gas, const float* w_d) {
    rrnn_gas_t g{};
    std::strncpy(g.name, gas.c_str(), sizeof(g.name) - 1);
    g.conc = w_d; g.ndims = 2;
    put(g);
    return "";
  }
  const rrnn_gas_t* data() const { return gases_.data(); }
  int size() const { return static_cast<int>(gases_.size()); }

 private:
  void put(const rrnn_gas_t& g) {
    for (auto& x : gases_)
      if (!std::strncmp(x.name, g.name, sizeof(g.name))) { x = g; return; }
    gases_.push_back(g);
  }
  std::vector<rrnn_gas_t> gases_;
};

// spectral discretisation shared by the optical-property, source and gas-optics types
class ty_optical_props {
 public:
  int get_ngpt() const { return ngpt_; }
  int get_nband() const { return nbnd_; }
  const std::vector<int>& get_band_lims_gpoint() const { return band_lims_; }
  const rrnn_kdist_t* kd() const { return kd_.get(); }
  void init(const ty_optical_props& spectral) { ngpt_ = spectral.ngpt_; nbnd_ = spectral.nbnd_; band_lims_ = spectral.band_lims_; kd_ = spectral.kd_; }

 protected:
  int ngpt_ = 0, nbnd_ = 0;
  std::vector<int> band_lims_;  // (2,nbnd), 1-based inclusive
  std::shared_ptr<rrnn_kdist_t> kd_;
};

class ty_optical_props_1scl : public ty_optical_props {
 public:
  std::string alloc_1scl(int ncol, int nlay, const ty_optical_props& spectral) {
    if (ncol <= 0 || nlay <= 0) return "optical_props%alloc: must provide positive extents for ncol, nlay";
    init(spectral);
    ncol_ = ncol; nlay_ = nlay;
    tau.resize(static_cast<size_t>(ncol) * nlay * ngpt_);
    return tau.size() ? "" : "optical_props%alloc: device allocation failed";
  }
  int get_ncol() const { return ncol_; }
  int get_nlay() const { return nlay_; }
  dev_array tau;

 protected:
  int ncol_ = 0, nlay_ = 0;
};

class ty_optical_props_2str : public ty_optical_props_1scl {
 public:
  // g is identically 0 on the NN path (rrtmgp/mo_gas_optics_rrtmgp.F90:560-567) and is kept implicit unless asked for
  std::string alloc_2str(int ncol, int nlay, const ty_optical_props& spectral, bool with_g = false) {
    std::string e = alloc_1scl(ncol, nlay, spectral);
    if (!e.empty()) return e;
    ssa.resize(tau.size());
    if (with_g) g.resize(tau.size());
    return "";
  }
  dev_array ssa, g;
};

class ty_source_func_lw : public ty_optical_props {
 public:
  std::string alloc(int ncol, int nlay, const ty_optical_props& spectral) {
    if (ncol <= 0 || nlay <= 0) return "source_func_lw%alloc: must provide positive extents for ncol, nlay";
    init(spectral);
    ncol_ = ncol; nlay_ = nlay;
    const size_t G = static_cast<size_t>(ngpt_);
    lay_source.resize(G * nlay * ncol); lev_source.resize(G * (nlay + 1) * ncol);
    sfc_source.resize(G * ncol); sfc_source_Jac.resize(G * ncol);
    return "";
  }
  int get_ncol() const { return ncol_; }
  int get_nlay() const { return nlay_; }
  dev_array lay_source, lev_source, sfc_source, sfc_source_Jac;

 private:
  int ncol_ = 0, nlay_ = 0;
};

// outputs the caller associates: device pointers to (nlay+1,ncol) arrays, nullptr = not wanted
struct ty_fluxes_broadband {
  float* flux_up = nullptr;
  float* flux_dn = nullptr;
  float* flux_net = nullptr;
  float* flux_dn_dir = nullptr;
  bool are_desired() const { return flux_up || flux_dn || flux_net || flux_dn_dir; }
};

// ty_fluxes_byband (extensions/mo_fluxes_byband.F90:31-131): by-band outputs (nbnd,nlay+1,ncol) the caller associates
struct ty_fluxes_byband : ty_fluxes_broadband {
  float* bnd_flux_up = nullptr;
  float* bnd_flux_dn = nullptr;
  float* bnd_flux_net = nullptr;
  float* bnd_flux_dn_dir = nullptr;
  bool are_desired() const { return bnd_flux_up || bnd_flux_dn || bnd_flux_net || bnd_flux_dn_dir || ty_fluxes_broadband::are_desired(); }
  // reduce_byband, by-band part: g-point fluxes (ngpt,nlev,ncol) on the device -> the associated by-band arrays
  std::string reduce(const context& ctx, const float* gpt_flux_up_d, const float* gpt_flux_dn_d, const ty_optical_props& spectral_disc,
                     int ncol, int nlev, const float* gpt_flux_dn_dir_d = nullptr) const {
    if (bnd_flux_dn_dir && !gpt_flux_dn_dir_d) return "reduce: requesting bnd_flux_dn_dir but direct flux hasn't been supplied";
    const rrnn_kdist_t* kd = spectral_disc.kd();
    if (!kd) return "reduce: spectral discretization carries no band limits";
    int rc = 0;
    if (bnd_flux_up) rc = rrnn_sum_byband(ctx.h(), kd, nlev, ncol, gpt_flux_up_d, bnd_flux_up);
    if (!rc && bnd_flux_dn) rc = rrnn_sum_byband(ctx.h(), kd, nlev, ncol, gpt_flux_dn_d, bnd_flux_dn);
    if (!rc && bnd_flux_dn_dir) rc = rrnn_sum_byband(ctx.h(), kd, nlev, ncol, gpt_flux_dn_dir_d, bnd_flux_dn_dir);
    if (!rc && bnd_flux_net) {
      if (bnd_flux_dn && bnd_flux_up)  // net_byband_precalc
        rc = rrnn_net_flux(ctx.h(), static_cast<size_t>(ncol) * nlev * spectral_disc.get_nband(), bnd_flux_dn, bnd_flux_up, bnd_flux_net);
      else                             // net_byband_full
        rc = rrnn_net_byband(ctx.h(), kd, nlev, ncol, gpt_flux_dn_d, gpt_flux_up_d, bnd_flux_net);
    }
    return err(rc);
  }
};

// ty_solar_var, extensions/solar_variability/mo_solar_variability.F90:20-183: load(avgcyc_ind) keeps the mean-solar-cycle table of the
// facular and sunspot indices ((2, nsolarfrac) in the reference == [nsolarfrac][2] here), solar_var_ind_interp interpolates it to a cycle
// fraction in [0, 1]; the pair is what ty_gas_optics_rrtmgp::set_solar_variability takes.  Host-only, as in the reference.
class ty_solar_var {
 public:
  std::string load(const float* avgcyc_ind, int nsolarfrac) { avgcyc_ind_.assign(avgcyc_ind, avgcyc_ind + 2 * static_cast<size_t>(nsolarfrac)); return std::string(); }
  void finalize() { avgcyc_ind_.clear(); }
  std::string solar_var_ind_interp(float solcycfrac, float& mg_index, float& sb_index) const {
    if (solcycfrac < 0.f || solcycfrac > 1.f) return "solar_var_ind_interp: solcycfrac out of range";
    if (avgcyc_ind_.empty()) return std::string();      // as the reference: nothing is computed without a table
    return err(rrnn_solar_var_ind_interp(avgcyc_ind_.data(), static_cast<int>(avgcyc_ind_.size() / 2), solcycfrac, &mg_index, &sb_index));
  }

 private:
  std::vector<float> avgcyc_ind_;
};

class ty_gas_optics_rrtmgp : public ty_optical_props {
 public:
  // load: the spectral tables the NN path needs (rrtmgp/mo_gas_optics_rrtmgp.F90:1130-1326): band -> g-point limits,
  // totplnk (nPlanckTemp,nbnd) for a longwave, solar_source (ngpt) for a shortwave k-distribution (nullptr otherwise)
  std::string load(const context& ctx, int nbnd, int ngpt, const int* band_lims_gpt, int ntemp, const float* totplnk,
                   float temp_ref_min, float totplnk_delta, const float* solar_source) {
    ctx_ = &ctx;
    rrnn_kdist_t* k = nullptr;
    std::string e = err(rrnn_kdist_create(ctx.h(), nbnd, ngpt, band_lims_gpt, ntemp, totplnk, temp_ref_min, totplnk_delta, solar_source, &k));
    if (!e.empty()) return e;
    kd_.reset(k, [](rrnn_kdist_t* p) { rrnn_kdist_destroy(p); });
    nbnd_ = nbnd; ngpt_ = ngpt;
    band_lims_.assign(band_lims_gpt, band_lims_gpt + 2 * nbnd);
    internal_ = totplnk != nullptr;
    return "";
  }
  bool source_is_internal() const { return internal_; }
  bool source_is_external() const { return !internal_; }
  std::string set_tsi(float tsi) { return err(rrnn_kdist_set_tsi(kd_.get(), tsi)); }
  // the optional tables of load (:1163, 1210, 1317-1325): solar_source_quiet / _facular / _sunspot (ngpt), optimal_angle_fit (2,nbnd)
  std::string load_solar_tables(const float* solar_quiet, const float* solar_facular, const float* solar_sunspot) {
    return err(rrnn_kdist_set_solar_tables(kd_.get(), solar_quiet, solar_facular, solar_sunspot));
  }
  std::string load_optimal_angle_fit(const float* optimal_angle_fit) { return err(rrnn_kdist_set_optimal_angle_fit(kd_.get(), optimal_angle_fit)); }
  // set_solar_variability(mg_index, sb_index [, tsi]), :1058-1095
  std::string set_solar_variability(float mg_index, float sb_index) { return err(rrnn_kdist_set_solar_variability(kd_.get(), mg_index, sb_index, 0, 0.f)); }
  std::string set_solar_variability(float mg_index, float sb_index, float tsi) {
    return err(rrnn_kdist_set_solar_variability(kd_.get(), mg_index, sb_index, 1, tsi));
  }
  std::vector<float> get_solar_source() const {
    std::vector<float> v(static_cast<size_t>(ngpt_));
    if (rrnn_kdist_get_solar_source(kd_.get(), v.data())) v.clear();
    return v;
  }
  bool gpoints_are_equal(const ty_optical_props& other) const { return ngpt_ == other.get_ngpt() && band_lims_ == other.get_band_lims_gpoint(); }
  // compute_optimal_angles(optical_props, optimal_angles), :1712-1758; optimal_angles is (ngpt,ncol), what rte_lw's lw_Ds takes
  std::string compute_optimal_angles(const ty_optical_props_1scl& optical_props, dev_array& optimal_angles) const;

  // longwave: gas_optics(play, plev, tlay, tsfc, gas_desc, optical_props, sources, tlev=, neural_nets=)
  std::string gas_optics(const float* play_d, const float* plev_d, const float* tlay_d, const float* tsfc_d, const ty_gas_concs& gas_desc,
                         ty_optical_props_1scl& optical_props, ty_source_func_lw& sources, const float* tlev_d,
                         const std::vector<const rrtmgp_network_type*>& neural_nets) const {
    if (!internal_) return "gas_optics(): this k-distribution has no internal (Planck) source";
    if (neural_nets.empty()) return "gas_optics(): only the neural-network path (neural_nets=) is implemented; the LUT path is out of scope";
    const int ncol = optical_props.get_ncol(), nlay = optical_props.get_nlay();
    if (sources.get_ncol() != ncol || sources.get_nlay() != nlay || sources.get_ngpt() != ngpt_)
      return "gas_optics%gas_optics: source function arrays inconsistently sized";
    const rrnn_model_t* m[2] = {neural_nets[0]->h(), neural_nets.size() > 1 ? neural_nets[1]->h() : nullptr};
    return err(rrnn_gas_optics_lw(ctx_->h(), kd_.get(), m, static_cast<int>(neural_nets.size()), ncol, nlay, play_d, plev_d, tlay_d, tsfc_d,
                                  gas_desc.data(), gas_desc.size(), tlev_d, optical_props.tau.data(), sources.lay_source.data(),
                                  sources.lev_source.data(), sources.sfc_source.data(), sources.sfc_source_Jac.data()));
  }
  // shortwave: gas_optics(play, plev, tlay, gas_desc, optical_props, toa_src, neural_nets=)
  std::string gas_optics(const float* play_d, const float* plev_d, const float* tlay_d, const ty_gas_concs& gas_desc,
                         ty_optical_props_2str& optical_props, dev_array& toa_src,
                         const std::vector<const rrtmgp_network_type*>& neural_nets) const {
    if (internal_) return "gas_optics(): this k-distribution has no external (solar) source";
    if (neural_nets.size() != 2) return "gas_optics(): the shortwave path needs the absorption and the Rayleigh network";
    const int ncol = optical_props.get_ncol(), nlay = optical_props.get_nlay();
    if (toa_src.size() != static_cast<size_t>(ncol) * ngpt_) return "gas_optics(): array toa_src has wrong size";
    const rrnn_model_t* m[2] = {neural_nets[0]->h(), neural_nets[1]->h()};
    return err(rrnn_gas_optics_sw(ctx_->h(), kd_.get(), m, ncol, nlay, play_d, plev_d, tlay_d, gas_desc.data(), gas_desc.size(),
                                  optical_props.tau.data(), optical_props.ssa.data(), optical_props.g.size() ? optical_props.g.data() : nullptr,
                                  toa_src.data()));
  }
  const context& ctx() const { return *ctx_; }

 private:
  const context* ctx_ = nullptr;
  bool internal_ = false;
};

inline std::string ty_gas_optics_rrtmgp::compute_optimal_angles(const ty_optical_props_1scl& optical_props, dev_array& optimal_angles) const {
  if (!gpoints_are_equal(optical_props))
    return "gas_optics%compute_optimal_angles: optical_props has different spectral discretization than gas_optics";
  if (optimal_angles.size() != static_cast<size_t>(optical_props.get_ncol()) * ngpt_)
    return "gas_optics%compute_optimal_angles: optimal_angles different dimension (ncol)";
  return err(rrnn_compute_optimal_angles(ctx_->h(), kd_.get(), optical_props.get_nlay(), optical_props.get_ncol(), optical_props.tau.data(),
                                         optimal_angles.data()));
}

// rte_lw(optical_props, top_at_1, sources, sfc_emis, fluxes [, inc_flux] [, n_gauss_angles]); sfc_emis_d is (nbnd,ncol)
inline std::string rte_lw(const context& ctx, const ty_optical_props_1scl& optical_props, bool top_at_1, const ty_source_func_lw& sources,
                          const float* sfc_emis_d, const ty_fluxes_broadband& fluxes, const float* inc_flux_d = nullptr,
                          int n_gauss_angles = 1) {
  if (!fluxes.are_desired()) return "rte_lw: no space allocated for fluxes";
  if (n_gauss_angles > 4) return "rte_lw: asking for too many quadrature points for no-scattering calculation";
  if (n_gauss_angles < 1) return "rte_lw: have to ask for at least one quadrature point for no-scattering calculation";
  if (sources.get_ncol() != optical_props.get_ncol() || sources.get_nlay() != optical_props.get_nlay())
    return "rte_lw: sources and optical properties inconsistently sized";
  const int ncol = optical_props.get_ncol(), nlay = optical_props.get_nlay();
  std::string e = err(rrnn_rte_lw(ctx.h(), optical_props.kd(), nlay, ncol, top_at_1 ? 1 : 0, n_gauss_angles, inc_flux_d, optical_props.tau.data(),
                                  sources.lay_source.data(), sources.lev_source.data(), sources.sfc_source.data(), sfc_emis_d, fluxes.flux_up,
                                  fluxes.flux_dn));
  if (e.empty() && fluxes.flux_net)
    e = err(rrnn_net_flux(ctx.h(), static_cast<size_t>(ncol) * (nlay + 1), fluxes.flux_dn, fluxes.flux_up, fluxes.flux_net));
  return e;
}

// rte_sw(atmos, top_at_1, mu0, inc_flux, sfc_alb_dir, sfc_alb_dif, fluxes [, inc_flux_dif]); albedos per g-point (ngpt,ncol)
inline std::string rte_sw(const context& ctx, const ty_optical_props_2str& atmos, bool top_at_1, const float* mu0_d, const float* inc_flux_d,
                          const float* sfc_alb_dir_d, const float* sfc_alb_dif_d, const ty_fluxes_broadband& fluxes,
                          const float* inc_flux_dif_d = nullptr) {
  if (!fluxes.are_desired()) return "rte_sw: no space allocated for fluxes";
  const int ncol = atmos.get_ncol(), nlay = atmos.get_nlay();
  std::string e = err(rrnn_rte_sw(ctx.h(), atmos.get_ngpt(), nlay, ncol, top_at_1 ? 1 : 0, mu0_d, inc_flux_d, sfc_alb_dir_d, sfc_alb_dif_d,
                                  inc_flux_dif_d, atmos.tau.data(), atmos.ssa.data(), atmos.g.size() ? atmos.g.data() : nullptr, fluxes.flux_up,
                                  fluxes.flux_dn, fluxes.flux_dn_dir));
  if (e.empty() && fluxes.flux_net)
    e = err(rrnn_net_flux(ctx.h(), static_cast<size_t>(ncol) * (nlay + 1), fluxes.flux_dn, fluxes.flux_up, fluxes.flux_net));
  return e;
}

}  // namespace rrtmgp_nn
#endif  // RRNN_HPP
