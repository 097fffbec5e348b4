// NN model file I/O (host side of rrtmgp_network_type):
//   rrnn_model_load_netcdf  <- load_netcdf, neural/mod_network_rrtmgp.F90:58-122
//   rrnn_model_load_ascii   <- network_type%load, neural/mod_network.F90:163-209 (+ sidecar scaling file)
//   rrnn_model_save_ascii   <- network_type%save, neural/mod_network.F90:495-510 (written consistently with load)
//
// The shipped weight files are netCDF-4 = HDF5.  Neither libnetcdf nor libhdf5 exists in the target image,
// so this is a self-contained reader for exactly the HDF5 subset those files use:
// version-2 object headers reached through dense link records ("<len><name><8-byte address>"), continuation
// blocks, dataspace / datatype / layout(v3) messages, contiguous or single-chunk little-endian data, no filters.
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>
#include "../../include/rrnn.h"
#include "nc4.hpp"

namespace rrnn {
int fail(const std::string& msg);

namespace {

using nc4::DsInfo;
using nc4::Nc4File;

int act_code(const std::string& s, bool& known) {
  known = true;
  if (s == "linear") return RRNN_ACT_LINEAR;
  if (s == "softsign") return RRNN_ACT_SOFTSIGN;
  if (s == "relu") return RRNN_ACT_RELU;
  if (s == "sigmoid") return RRNN_ACT_SIGMOID;
  if (s == "hard_sigmoid") return RRNN_ACT_HARD_SIGMOID;
  known = false;
  return RRNN_ACT_LINEAR;  // "failed to read activation function, setting to linear", mod_layer.F90:91-94
}
const char* act_name(int c) {
  switch (c) {
    case RRNN_ACT_SOFTSIGN: return "softsign";
    case RRNN_ACT_RELU: return "relu";
    case RRNN_ACT_SIGMOID: return "sigmoid";
    case RRNN_ACT_HARD_SIGMOID: return "hard_sigmoid";
    default: return "linear";
  }
}

}  // namespace
}  // namespace rrnn

using namespace rrnn;

extern "C" int rrnn_model_load_netcdf(rrnn_ctx_t* ctx, const char* filename, rrnn_model_t** out) {
  if (!filename || !out) return fail("load_netcdf: null argument");
  Nc4File f;
  std::string err;
  if (!f.open(filename, err)) return fail("mod_network_rrtmgp:load_netcdf: " + err);
  std::vector<float> v;
  std::vector<uint64_t> shp;
  if (!f.read_float("nn_dimsize", v, shp, err)) return fail("load_netcdf: " + err);
  const int nlayers = (int)v.size();
  if (nlayers < 2 || nlayers > 6) return fail("load_netcdf: unsupported number of layers");
  std::vector<float> xmin, xmax;
  if (!f.read_float("nn_input_coeffs_min", xmin, shp, err)) return fail("load_netcdf: " + err);
  if (!f.read_float("nn_input_coeffs_max", xmax, shp, err)) return fail("load_netcdf: " + err);
  const int nx = (int)xmin.size();
  if ((int)xmax.size() != nx) return fail("load_netcdf: inconsistent input scaling coefficients");
  std::vector<int> dims(nlayers + 1);
  dims[0] = nx;
  for (int i = 0; i < nlayers; ++i) dims[i + 1] = (int)v[i];
  std::vector<float> wpack, bpack;
  for (int n = 1; n <= nlayers; ++n) {
    std::vector<float> w, b;
    if (!f.read_float("nn_weights_" + std::to_string(n), w, shp, err)) return fail("load_netcdf: " + err);
    if (shp.size() != 2 || (int)shp[0] != dims[n - 1] || (int)shp[1] != dims[n]) return fail("load_netcdf: weight matrix has unexpected shape");
    if (!f.read_float("nn_bias_" + std::to_string(n), b, shp, err)) return fail("load_netcdf: " + err);
    if ((int)b.size() != dims[n]) return fail("load_netcdf: bias vector has unexpected length");
    wpack.insert(wpack.end(), w.begin(), w.end());
    bpack.insert(bpack.end(), b.begin(), b.end());
  }
  std::vector<std::string> acts, names;
  if (!f.read_strings("nn_activation_char", acts, err)) return fail("load_netcdf: " + err);
  if (!f.read_strings("nn_inputs_char", names, err)) return fail("load_netcdf: " + err);
  if ((int)acts.size() < nlayers || (int)names.size() < nx) return fail("load_netcdf: short activation or input-name list");
  std::vector<int> act(nlayers);
  for (int i = 0; i < nlayers; ++i) { bool known; act[i] = act_code(acts[i], known); }
  std::vector<char> nm((size_t)nx * 32, ' ');
  for (int i = 0; i < nx; ++i) memcpy(nm.data() + 32 * i, names[i].data(), std::min<size_t>(names[i].size(), 32));
  std::vector<float> ymean, ystd;
  // absent is fine (models without output scaling); present but unreadable is an error of its own, not "missing"
  const bool hm = f.has("nn_output_coeffs_mean"), hs = f.has("nn_output_coeffs_std");
  if (hm && !f.read_float("nn_output_coeffs_mean", ymean, shp, err)) return fail("mod_network_rrtmgp:load_netcdf: " + err);
  if (hs && !f.read_float("nn_output_coeffs_std", ystd, shp, err)) return fail("mod_network_rrtmgp:load_netcdf: " + err);
  if (hm && (int)ymean.size() != dims[nlayers]) return fail("load_netcdf: output mean has unexpected length");
  if (hs && (int)ystd.size() != dims[nlayers]) return fail("load_netcdf: output std has unexpected length");
  return rrnn_model_create(ctx, nlayers, dims.data(), wpack.data(), bpack.data(), act.data(), xmin.data(), xmax.data(),
                           hm ? ymean.data() : nullptr, hs ? ystd.data() : nullptr, nm.data(), out);
}

// ASCII model, neural/mod_network.F90:163-209: N (layers incl. input pseudo-layer); dims(1:N); N-1 bias records;
// N-1 weight records w_n(dims(n),dims(n+1)) in Fortran order (input index fastest); N-1 activation names.
// Sidecar (this library's own, the ASCII format has no scaling): "nx ny has_y", input names, xmin, xmax[, ymean, ystd].
extern "C" int rrnn_model_load_ascii(rrnn_ctx_t* ctx, const char* model_txt, const char* scaling_txt, rrnn_model_t** out) {
  if (!model_txt || !scaling_txt || !out) return fail("load: null argument");
  std::ifstream f(model_txt);
  if (!f) return fail(std::string("load: can't find file ") + model_txt);
  int N = 0;
  f >> N;
  if (N < 3 || N > 7) return fail("load: unsupported number of layers");
  std::vector<int> dims(N);
  for (int i = 0; i < N; ++i) f >> dims[i];
  const int nl = N - 1;
  std::vector<float> bpack, wpack;
  for (int n = 0; n < nl; ++n) for (int i = 0; i < dims[n + 1]; ++i) { float v; f >> v; bpack.push_back(v); }
  for (int n = 0; n < nl; ++n) {
    const int K = dims[n], O = dims[n + 1];
    std::vector<float> w((size_t)K * O);
    // file order: input index fastest -> w_f[o][k]; stored as row-major (K,O)
    for (int o = 0; o < O; ++o) for (int k = 0; k < K; ++k) { float v; f >> v; w[(size_t)k * O + o] = v; }
    wpack.insert(wpack.end(), w.begin(), w.end());
  }
  std::vector<int> act(nl);
  for (int n = 0; n < nl; ++n) { std::string s; f >> s; bool known; act[n] = act_code(s, known); }
  if (!f) return fail("load: truncated or malformed model file");
  std::ifstream g(scaling_txt);
  if (!g) return fail(std::string("load: can't find file ") + scaling_txt);
  int nx = 0, ny = 0, has_y = 0;
  g >> nx >> ny >> has_y;
  if (nx != dims[0] || ny != dims[nl]) return fail("load: scaling file does not match the model dimensions");
  std::vector<char> nm((size_t)nx * 32, ' ');
  for (int i = 0; i < nx; ++i) { std::string s; g >> s; memcpy(nm.data() + 32 * i, s.data(), std::min<size_t>(s.size(), 32)); }
  std::vector<float> xmin(nx), xmax(nx), ymean(ny), ystd(ny);
  for (auto& v : xmin) g >> v;
  for (auto& v : xmax) g >> v;
  if (has_y) { for (auto& v : ymean) g >> v; for (auto& v : ystd) g >> v; }
  if (!g) return fail("load: truncated or malformed scaling file");
  return rrnn_model_create(ctx, nl, dims.data(), wpack.data(), bpack.data(), act.data(), xmin.data(), xmax.data(),
                           has_y ? ymean.data() : nullptr, has_y ? ystd.data() : nullptr, nm.data(), out);
}

extern "C" int rrnn_model_save_ascii(const rrnn_model_t* m, const char* model_txt, const char* scaling_txt) {
  if (!m || !model_txt || !scaling_txt) return fail("save: null argument");
  const int nl = rrnn_model_nlayers(m);
  std::vector<int> dims(nl + 1);
  rrnn_model_dims(m, dims.data());
  FILE* f = fopen(model_txt, "w");
  if (!f) return fail(std::string("save: can't open ") + model_txt);
  fprintf(f, "%d\n", nl + 1);
  for (int i = 0; i <= nl; ++i) fprintf(f, "%d ", dims[i]);
  fprintf(f, "\n");
  std::vector<float> tmp;
  int n = 0;
  for (int l = 0; l < nl; ++l) {
    rrnn_model_get(m, 1, l, nullptr, &n); tmp.resize(n); rrnn_model_get(m, 1, l, tmp.data(), &n);
    for (int i = 0; i < n; ++i) fprintf(f, "%.9g ", tmp[i]);
    fprintf(f, "\n");
  }
  for (int l = 0; l < nl; ++l) {
    rrnn_model_get(m, 0, l, nullptr, &n); tmp.resize(n); rrnn_model_get(m, 0, l, tmp.data(), &n);
    const int K = dims[l], O = dims[l + 1];
    for (int o = 0; o < O; ++o) for (int k = 0; k < K; ++k) fprintf(f, "%.9g ", tmp[(size_t)k * O + o]);
    fprintf(f, "\n");
  }
  for (int l = 0; l < nl; ++l) fprintf(f, "%s\n", act_name(rrnn_model_activation(m, l)));
  fclose(f);
  FILE* g = fopen(scaling_txt, "w");
  if (!g) return fail(std::string("save: can't open ") + scaling_txt);
  int ny = 0;
  rrnn_model_get(m, 4, 0, nullptr, &ny);
  const int has_y = ny > 0;
  fprintf(g, "%d %d %d\n", dims[0], dims[nl], has_y);
  for (int i = 0; i < dims[0]; ++i) { char b[32]; rrnn_model_input_name(m, i, b); fprintf(g, "%s ", b[0] ? b : "_"); }
  fprintf(g, "\n");
  for (int which = 2; which <= (has_y ? 5 : 3); ++which) {
    rrnn_model_get(m, which, 0, nullptr, &n); tmp.resize(n); rrnn_model_get(m, which, 0, tmp.data(), &n);
    for (int i = 0; i < n; ++i) fprintf(g, "%.9g ", tmp[i]);
    fprintf(g, "\n");
  }
  fclose(g);
  return 0;
}
