// One process, N devices: the column-sharded whole-path drivers for hosts that are not launched one rank per GPU
// ("One process driving 8 devices is enough -- no MPI launcher").
//
// Every column is independent through gas optics and both solvers (the reference's only parallelism is an OpenMP parallel-do
// over column blocks, examples/rfmip-clear-sky/rrtmgp_rfmip_lw.F90:364-368), so a rrnn_multi_t holds one context per device
// with the spectral tables and the networks replicated on each, cuts the columns of a call into contiguous shards (sizes
// differing by at most one column), and runs the host-buffer pipeline of every shard (rrnn_{lw,sw}_fluxes_host: H2D / kernels /
// D2H overlapped per chunk) on its own host thread.  Every shard writes its fluxes straight into its slice of the caller's
// arrays: there is no exchange step, hence no collective (NCCL is only needed when the fluxes have to end up on the devices).
#include "common.cuh"
#include <thread>

struct rrnn_multi {
  std::vector<rrnn_ctx_t*> ctx;
  std::vector<std::vector<rrnn_kdist_t*>> kd;      // [id][device]
  std::vector<std::vector<rrnn_model_t*>> model;   // [id][device]
};

using namespace rrnn;

extern "C" int rrnn_multi_create(int ndev, const int* devices, rrnn_multi_t** out) {
  RRNN_CHECK(out && ndev >= 1, "rrnn_multi_create: bad argument");
  const int have = rrnn_device_count();
  RRNN_CHECK(have >= 1, "rrnn_multi_create: no CUDA device available (this library has no CPU fallback)");
  rrnn_multi* m = new rrnn_multi;
  for (int i = 0; i < ndev; ++i) {
    const int d = devices ? devices[i] : i;
    rrnn_ctx_t* c = nullptr;
    if (d < 0 || d >= have || rrnn_ctx_create(d, nullptr, &c) != 0) {
      const std::string why = (d < 0 || d >= have) ? "rrnn_multi_create: device " + std::to_string(d) + " does not exist" : std::string(rrnn_last_error());
      for (auto* x : m->ctx) rrnn_ctx_destroy(x);
      delete m;
      return fail(why);
    }
    // a context created with a null stream uses the legacy default stream, which serialises the devices' host threads
    // against each other less than it seems (it is per device) -- but give every shard its own stream anyway
    cudaStream_t s = nullptr;
    cudaSetDevice(d);
    if (cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking) == cudaSuccess) { rrnn_ctx_set_stream(c, s); c->own_stream = true; }
    m->ctx.push_back(c);
  }
  *out = m;
  return 0;
}

extern "C" int rrnn_multi_destroy(rrnn_multi_t* m) {
  if (!m) return 0;
  for (auto& v : m->kd) for (auto* k : v) rrnn_kdist_destroy(k);
  for (auto& v : m->model) for (auto* x : v) rrnn_model_destroy(x);
  for (auto* c : m->ctx) rrnn_ctx_destroy(c);
  delete m;
  return 0;
}

extern "C" int rrnn_multi_ndev(const rrnn_multi_t* m) { return m ? (int)m->ctx.size() : 0; }
extern "C" rrnn_ctx_t* rrnn_multi_ctx(rrnn_multi_t* m, int i) { return (m && i >= 0 && i < (int)m->ctx.size()) ? m->ctx[i] : nullptr; }

extern "C" int rrnn_multi_set_flag(rrnn_multi_t* m, const char* name, int value) {
  RRNN_CHECK(m, "rrnn_multi_set_flag: null handle");
  for (auto* c : m->ctx)
    if (int rc = rrnn_ctx_set_flag(c, name, value)) return rc;
  return 0;
}

extern "C" int rrnn_multi_model_load_netcdf(rrnn_multi_t* m, const char* filename, int* model_id) {
  RRNN_CHECK(m && filename && model_id, "rrnn_multi_model_load_netcdf: null argument");
  std::vector<rrnn_model_t*> v;
  for (auto* c : m->ctx) {
    rrnn_model_t* x = nullptr;
    if (int rc = rrnn_model_load_netcdf(c, filename, &x)) { for (auto* y : v) rrnn_model_destroy(y); return rc; }
    v.push_back(x);
  }
  m->model.push_back(v);
  *model_id = (int)m->model.size() - 1;
  return 0;
}

extern "C" int rrnn_multi_kdist_create(rrnn_multi_t* m, int nbnd, int ngpt, const int* band_lims_gpt, int ntemp, const float* totplnk,
                                       float temp_ref_min, float totplnk_delta, const float* solar_source, int* kdist_id) {
  RRNN_CHECK(m && kdist_id, "rrnn_multi_kdist_create: null argument");
  std::vector<rrnn_kdist_t*> v;
  for (auto* c : m->ctx) {
    rrnn_kdist_t* k = nullptr;
    if (int rc = rrnn_kdist_create(c, nbnd, ngpt, band_lims_gpt, ntemp, totplnk, temp_ref_min, totplnk_delta, solar_source, &k)) {
      for (auto* y : v) rrnn_kdist_destroy(y);
      return rc;
    }
    v.push_back(k);
  }
  m->kd.push_back(v);
  *kdist_id = (int)m->kd.size() - 1;
  return 0;
}

extern "C" int rrnn_multi_kdist_set_tsi(rrnn_multi_t* m, int kdist_id, float tsi) {
  RRNN_CHECK(m && kdist_id >= 0 && kdist_id < (int)m->kd.size(), "rrnn_multi_kdist_set_tsi: bad k-distribution id");
  for (auto* k : m->kd[kdist_id])
    if (int rc = rrnn_kdist_set_tsi(k, tsi)) return rc;
  return 0;
}

namespace {
// contiguous shard [c0, c1) of device i: sizes differ by at most one column (the rule of rte_rrtmgp_nn_b200/sharding.py)
inline void shard(long long ncol, int i, int n, long long& c0, long long& c1) { c0 = ncol * i / n; c1 = ncol * (i + 1) / n; }

void shard_gases(const rrnn_gas_t* in, int ngas, long long c0, int nlay, std::vector<rrnn_gas_t>& out) {
  out.assign(in, in + ngas);
  for (auto& g : out)
    if (g.ndims == 2 && g.conc) g.conc += (size_t)c0 * nlay;
}

// run fn(device index) on one host thread per device; the first error (by device order) is reported
template <typename F>
int on_all_devices(rrnn_multi_t* m, F fn) {
  const int n = (int)m->ctx.size();
  std::vector<int> rc(n, 0);
  std::vector<std::string> msg(n);
  std::vector<std::thread> th;
  for (int i = 0; i < n; ++i)
    th.emplace_back([&, i]() {
      rc[i] = fn(i);
      if (rc[i]) msg[i] = rrnn_last_error();   // the error string is thread-local: carry it over
    });
  for (auto& t : th) t.join();
  for (int i = 0; i < n; ++i)
    if (rc[i]) return fail("device " + std::to_string(m->ctx[i]->device) + ": " + msg[i]);
  return 0;
}
}  // namespace

extern "C" int rrnn_multi_lw_fluxes_host(rrnn_multi_t* m, int kdist_id, const int* model_ids, int nmodels, int ncol, int nlay, int top_at_1,
                                         int n_gauss_angles, const float* play, const float* plev, const float* tlay, const float* tlev,
                                         const float* tsfc, const float* sfc_emis, const rrnn_gas_t* gases, int ngas, float* flux_up,
                                         float* flux_dn) {
  RRNN_CHECK(m && model_ids && kdist_id >= 0 && kdist_id < (int)m->kd.size(), "rrnn_multi_lw_fluxes_host: bad handle or k-distribution id");
  RRNN_CHECK(nmodels == 1 || nmodels == 2, "gas_optics(): neural_nets must hold 1 or 2 networks for the longwave");
  for (int i = 0; i < nmodels; ++i) RRNN_CHECK(model_ids[i] >= 0 && model_ids[i] < (int)m->model.size(), "rrnn_multi_lw_fluxes_host: bad model id");
  RRNN_CHECK(play && plev && tlay && tsfc && sfc_emis && flux_up && flux_dn, "rrnn_multi_lw_fluxes_host: null argument");
  if (ncol <= 0) return 0;
  const int n = (int)m->ctx.size();
  const size_t L = nlay;
  return on_all_devices(m, [&](int i) {
    long long c0, c1;
    shard(ncol, i, n, c0, c1);
    if (c1 <= c0) return 0;
    std::vector<rrnn_gas_t> gs;
    shard_gases(gases, ngas, c0, nlay, gs);
    const rrnn_model_t* mods[2] = {m->model[model_ids[0]][i], nmodels > 1 ? m->model[model_ids[1]][i] : nullptr};
    return rrnn_lw_fluxes_host(m->ctx[i], m->kd[kdist_id][i], mods, nmodels, (int)(c1 - c0), nlay, top_at_1, n_gauss_angles, play + c0 * L,
                               plev + c0 * (L + 1), tlay + c0 * L, tlev ? tlev + c0 * (L + 1) : nullptr, tsfc + c0, sfc_emis + c0, gs.data(), ngas,
                               flux_up + c0 * (L + 1), flux_dn + c0 * (L + 1));
  });
}

extern "C" int rrnn_multi_sw_fluxes_host(rrnn_multi_t* m, int kdist_id, const int* model_ids, int ncol, int nlay, int top_at_1,
                                         const float* play, const float* plev, const float* tlay, const float* mu0, const float* sfc_alb,
                                         const float* tsi, const rrnn_gas_t* gases, int ngas, float* flux_up, float* flux_dn,
                                         float* flux_dn_dir) {
  RRNN_CHECK(m && model_ids && kdist_id >= 0 && kdist_id < (int)m->kd.size(), "rrnn_multi_sw_fluxes_host: bad handle or k-distribution id");
  for (int i = 0; i < 2; ++i) RRNN_CHECK(model_ids[i] >= 0 && model_ids[i] < (int)m->model.size(), "rrnn_multi_sw_fluxes_host: bad model id");
  RRNN_CHECK(play && plev && tlay && mu0 && sfc_alb && flux_up && flux_dn && flux_dn_dir, "rrnn_multi_sw_fluxes_host: null argument");
  if (ncol <= 0) return 0;
  const int n = (int)m->ctx.size();
  const size_t L = nlay;
  return on_all_devices(m, [&](int i) {
    long long c0, c1;
    shard(ncol, i, n, c0, c1);
    if (c1 <= c0) return 0;
    std::vector<rrnn_gas_t> gs;
    shard_gases(gases, ngas, c0, nlay, gs);
    const rrnn_model_t* mods[2] = {m->model[model_ids[0]][i], m->model[model_ids[1]][i]};
    return rrnn_sw_fluxes_host(m->ctx[i], m->kd[kdist_id][i], mods, (int)(c1 - c0), nlay, top_at_1, play + c0 * L, plev + c0 * (L + 1),
                               tlay + c0 * L, mu0 + c0, sfc_alb + c0, tsi ? tsi + c0 : nullptr, gs.data(), ngas, flux_up + c0 * (L + 1),
                               flux_dn + c0 * (L + 1), flux_dn_dir + c0 * (L + 1));
  });
}
