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

namespace rrnn {
int fail(const std::string& msg);

namespace {

struct DsInfo {
  std::vector<uint64_t> shape;
  int cls = -1;      // 0 int, 1 float, 3 string
  uint32_t size = 0;
  bool big = false;
  int layout = -1;   // 0 compact, 1 contiguous, 2 chunked
  uint64_t addr = 0, nbytes = 0;
  std::vector<uint32_t> chunk;
  std::vector<uint8_t> compact;
  bool filtered = false;
};

class Nc4File {
 public:
  bool open(const std::string& path, std::string& err) {
    std::ifstream f(path, std::ios::binary);
    if (!f) { err = "can't find file " + path; return false; }
    buf_.assign(std::istreambuf_iterator<char>(f), std::istreambuf_iterator<char>());
    static const unsigned char magic[8] = {0x89, 'H', 'D', 'F', '\r', '\n', 0x1a, '\n'};
    if (buf_.size() < 64 || memcmp(buf_.data(), magic, 8) != 0) { err = path + ": not a netCDF-4/HDF5 file"; return false; }
    return true;
  }

  bool has(const std::string& name) const { return find_ohdr(name) != UINT64_MAX; }

  bool info(const std::string& name, DsInfo& d, std::string& err) const {
    const uint64_t addr = find_ohdr(name);
    if (addr == UINT64_MAX) { err = "variable " + name + " not found"; return false; }
    std::vector<std::pair<int, std::pair<uint64_t, uint32_t>>> msgs;
    if (!messages(addr, msgs, err)) return false;
    for (auto& m : msgs) {
      const uint8_t* b = buf_.data() + m.second.first;
      const uint32_t len = m.second.second;
      switch (m.first) {
        case 0x01: {
          if (len < 4) break;
          const int ver = b[0], rank = b[1];
          const int off = (ver == 1) ? 8 : 4;
          d.shape.clear();
          for (int k = 0; k < rank; ++k) d.shape.push_back(rd64(b + off + 8 * k));
          break;
        }
        case 0x03:
          d.cls = b[0] & 0x0F;
          d.big = (b[1] & 1) != 0;
          d.size = rd32(b + 4);
          break;
        case 0x08: {
          if (b[0] != 3) { err = name + ": unsupported data layout message version"; return false; }
          d.layout = b[1];
          if (d.layout == 1) { d.addr = rd64(b + 2); d.nbytes = rd64(b + 10); }
          else if (d.layout == 2) {
            const int nd = b[2];
            d.addr = rd64(b + 3);
            d.chunk.clear();
            for (int k = 0; k < nd; ++k) d.chunk.push_back(rd32(b + 11 + 4 * k));
          } else if (d.layout == 0) {
            const uint16_t sz = rd16(b + 2);
            d.compact.assign(b + 4, b + 4 + sz);
          }
          break;
        }
        case 0x0B: d.filtered = true; break;
        default: break;
      }
    }
    if (d.cls < 0 || d.layout < 0) { err = name + ": incomplete object header"; return false; }
    return true;
  }

  // read a numeric dataset as float (fp32 / fp64 / int32 / int64 sources)
  bool read_float(const std::string& name, std::vector<float>& out, std::vector<uint64_t>& shape, std::string& err) const {
    DsInfo d;
    std::vector<uint8_t> raw;
    if (!read_raw(name, d, raw, err)) return false;
    shape = d.shape;
    const size_t n = raw.size() / d.size;
    out.resize(n);
    for (size_t i = 0; i < n; ++i) {
      uint8_t tmp[8];
      memcpy(tmp, raw.data() + i * d.size, d.size);
      if (d.big) for (uint32_t k = 0; k < d.size / 2; ++k) std::swap(tmp[k], tmp[d.size - 1 - k]);
      if (d.cls == 1 && d.size == 4) { float v; memcpy(&v, tmp, 4); out[i] = v; }
      else if (d.cls == 1 && d.size == 8) { double v; memcpy(&v, tmp, 8); out[i] = (float)v; }
      else if (d.cls == 0 && d.size == 4) { int32_t v; memcpy(&v, tmp, 4); out[i] = (float)v; }
      else if (d.cls == 0 && d.size == 8) { int64_t v; memcpy(&v, tmp, 8); out[i] = (float)v; }
      else if (d.cls == 0 && d.size == 2) { int16_t v; memcpy(&v, tmp, 2); out[i] = (float)v; }
      else { err = name + ": unsupported datatype"; return false; }
    }
    return true;
  }

  // blank-padded character matrix (nrow, width) -> trimmed strings
  bool read_strings(const std::string& name, std::vector<std::string>& out, std::string& err) const {
    DsInfo d;
    std::vector<uint8_t> raw;
    if (!read_raw(name, d, raw, err)) return false;
    if (d.cls != 3) { err = name + ": not a character variable"; return false; }
    size_t nrow = d.shape.empty() ? 1 : d.shape[0];
    size_t width = raw.size() / (nrow ? nrow : 1);
    out.clear();
    for (size_t r = 0; r < nrow; ++r) {
      std::string s((const char*)raw.data() + r * width, width);
      size_t z = s.find('\0');
      if (z != std::string::npos) s.resize(z);
      while (!s.empty() && s.back() == ' ') s.pop_back();
      size_t b = 0;
      while (b < s.size() && s[b] == ' ') ++b;
      out.push_back(s.substr(b));
    }
    return true;
  }

 private:
  std::vector<uint8_t> buf_;

  static uint16_t rd16(const uint8_t* p) { uint16_t v; memcpy(&v, p, 2); return v; }
  static uint32_t rd32(const uint8_t* p) { uint32_t v; memcpy(&v, p, 4); return v; }
  static uint64_t rd64(const uint8_t* p) { uint64_t v; memcpy(&v, p, 8); return v; }

  uint64_t find_ohdr(const std::string& name) const {
    std::string key;
    key.push_back((char)name.size());
    key += name;
    const size_t n = buf_.size();
    for (size_t i = 0; i + key.size() + 8 <= n; ++i) {
      if (memcmp(buf_.data() + i, key.data(), key.size()) != 0) continue;
      const uint64_t addr = rd64(buf_.data() + i + key.size());
      if (addr + 6 <= n && memcmp(buf_.data() + addr, "OHDR", 4) == 0 && buf_[addr + 4] == 2) return addr;
    }
    return UINT64_MAX;
  }

  bool messages(uint64_t addr, std::vector<std::pair<int, std::pair<uint64_t, uint32_t>>>& out, std::string& err) const {
    const uint8_t flags = buf_[addr + 5];
    uint64_t p = addr + 6;
    if (flags & 0x20) p += 16;
    if (flags & 0x10) p += 4;
    const int w = 1 << (flags & 3);
    uint64_t size0 = 0;
    memcpy(&size0, buf_.data() + p, w);
    p += w;
    const bool track = (flags & 0x04) != 0;
    std::vector<std::pair<uint64_t, uint64_t>> blocks{{p, p + size0}};
    for (size_t bi = 0; bi < blocks.size(); ++bi) {
      uint64_t q = blocks[bi].first, end = blocks[bi].second;
      if (end > buf_.size()) { err = "corrupt object header"; return false; }
      while (q + 4 <= end) {
        const int mtype = buf_[q];
        const uint32_t msize = rd16(buf_.data() + q + 1);
        q += 4;
        if (track) q += 2;
        if (q + msize > end) break;
        if (mtype == 0x10) {
          const uint64_t off = rd64(buf_.data() + q), ln = rd64(buf_.data() + q + 8);
          if (off + ln > buf_.size() || memcmp(buf_.data() + off, "OCHK", 4) != 0) { err = "bad continuation block"; return false; }
          blocks.push_back({off + 4, off + ln - 4});
        } else {
          out.push_back({mtype, {q, msize}});
        }
        q += msize;
      }
    }
    return true;
  }

  bool read_raw(const std::string& name, DsInfo& d, std::vector<uint8_t>& raw, std::string& err) const {
    if (!info(name, d, err)) return false;
    if (d.filtered) { err = name + ": filtered (compressed) datasets are not supported"; return false; }
    uint64_t n = 1;
    for (uint64_t s : d.shape) n *= s;
    const uint64_t nbytes = n * d.size;
    if (d.layout == 1) {
      if (d.addr == UINT64_MAX || d.addr + nbytes > buf_.size()) { err = name + ": no data"; return false; }
      raw.assign(buf_.begin() + d.addr, buf_.begin() + d.addr + nbytes);
    } else if (d.layout == 0) {
      if (d.compact.size() < nbytes) { err = name + ": short compact data"; return false; }
      raw.assign(d.compact.begin(), d.compact.begin() + nbytes);
    } else {
      const size_t nd = d.chunk.size();
      if (nd != d.shape.size() + 1) { err = name + ": unexpected chunk rank"; return false; }
      for (size_t k = 0; k + 1 < nd; ++k)
        if (d.chunk[k] < d.shape[k]) { err = name + ": multi-chunk datasets are not supported"; return false; }
      const uint64_t bt = d.addr;
      if (bt + 8 > buf_.size() || memcmp(buf_.data() + bt, "TREE", 4) != 0 || buf_[bt + 4] != 1 || buf_[bt + 5] != 0 ||
          rd16(buf_.data() + bt + 6) != 1) { err = name + ": unsupported chunk index"; return false; }
      const uint64_t child = rd64(buf_.data() + bt + 8 + 16 + 8 + 8 * nd);
      // copy the [0:shape] corner of the single chunk
      raw.resize(nbytes);
      std::vector<uint64_t> idx(d.shape.size(), 0);
      const size_t rank = d.shape.size();
      const uint64_t row = (rank ? d.shape[rank - 1] : 1) * d.size;
      uint64_t nrows = 1;
      for (size_t k = 0; k + 1 < rank; ++k) nrows *= d.shape[k];
      for (uint64_t r = 0; r < nrows; ++r) {
        uint64_t src = 0, stride = 1, rem = r;
        // offset of row r inside the chunk
        std::vector<uint64_t> id(rank, 0);
        for (size_t k = rank - 1; k-- > 0;) { id[k] = rem % d.shape[k]; rem /= d.shape[k]; }
        stride = d.size;
        for (size_t k = rank; k-- > 0;) { src += id[k] * stride; stride *= d.chunk[k]; }
        if (child + src + row > buf_.size()) { err = name + ": chunk out of file"; return false; }
        memcpy(raw.data() + r * row, buf_.data() + child + src, row);
      }
    }
    return true;
  }
};

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
  const bool hm = f.has("nn_output_coeffs_mean") && f.read_float("nn_output_coeffs_mean", ymean, shp, err);
  const bool hs = f.has("nn_output_coeffs_std") && f.read_float("nn_output_coeffs_std", ystd, shp, err);
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
