// Self-contained reader for the HDF5 subset that netCDF-4 files use (no libnetcdf / libhdf5 in the target image; SURVEY.md
// Appendix A): superblock v0, version-2 object headers reached through link records ("<len><name><8-byte address>", dense or
// compact), continuation blocks, dataspace / datatype / layout(v3) messages, contiguous, compact or chunked (version-1 B-tree)
// unfiltered data of either byte order, version-1 attribute messages with string values.  Used by the NN model loader
// (model_io.cpp) and by the generic rrnn_nc_* entry points (nc4_io.cpp: RFMIP / Garand input files).
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <fstream>
#include <iterator>
#include <string>
#include <utility>
#include <vector>

namespace rrnn {
namespace nc4 {

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
          if ((uint32_t)(off + 8 * rank) > len) { err = name + ": truncated dataspace message"; return false; }
          d.shape.clear();
          for (int k = 0; k < rank; ++k) d.shape.push_back(rd64(b + off + 8 * k));
          break;
        }
        case 0x03:
          if (len < 8) { err = name + ": truncated datatype message"; return false; }
          d.cls = b[0] & 0x0F;
          d.big = (b[1] & 1) != 0;
          d.size = rd32(b + 4);
          break;
        case 0x08: {
          if (len < 4 || b[0] != 3) { err = name + ": unsupported data layout message version"; return false; }
          d.layout = b[1];
          if (d.layout == 1) {
            if (len < 18) { err = name + ": truncated layout message"; return false; }
            d.addr = rd64(b + 2); d.nbytes = rd64(b + 10);
          } else if (d.layout == 2) {
            const int nd = b[2];
            if ((uint32_t)(11 + 4 * nd) > len) { err = name + ": truncated layout message"; return false; }
            d.addr = rd64(b + 3);
            d.chunk.clear();
            for (int k = 0; k < nd; ++k) d.chunk.push_back(rd32(b + 11 + 4 * k));
          } else if (d.layout == 0) {
            const uint16_t sz = rd16(b + 2);
            if (4u + sz > len) { err = name + ": truncated compact dataset"; return false; }
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

  // string attribute of a variable (version-1 attribute message, e.g. the RFMIP 'units' scaling factors)
  bool attr_string(const std::string& var, const std::string& att, std::string& out, std::string& err) const {
    const uint64_t addr = find_ohdr(var);
    if (addr == UINT64_MAX) { err = "variable " + var + " not found"; return false; }
    std::vector<std::pair<int, std::pair<uint64_t, uint32_t>>> msgs;
    if (!messages(addr, msgs, err)) return false;
    auto pad8 = [](uint32_t v) { return (v + 7u) & ~7u; };
    for (auto& m : msgs) {
      if (m.first != 0x0C) continue;
      const uint8_t* b = buf_.data() + m.second.first;
      const uint32_t len = m.second.second;
      if (len < 8 || b[0] != 1) continue;
      const uint32_t nsz = rd16(b + 2), dsz = rd16(b + 4), ssz = rd16(b + 6);
      if (8 + pad8(nsz) + pad8(dsz) + pad8(ssz) > len || nsz == 0) continue;
      std::string nm((const char*)b + 8, nsz);
      nm.resize(strlen(nm.c_str()));
      if (nm != att) continue;
      const uint8_t* dt = b + 8 + pad8(nsz);
      if ((dt[0] & 0x0F) != 3) { err = var + ":" + att + " is not a string attribute"; return false; }
      const uint32_t strsize = rd32(dt + 4);
      const uint32_t p = 8 + pad8(nsz) + pad8(dsz) + pad8(ssz);
      if (p + strsize > len) { err = var + ":" + att + ": truncated attribute"; return false; }
      out.assign((const char*)b + p, strsize);
      out.resize(strlen(out.c_str()));
      return true;
    }
    err = "attribute " + att + " of " + var + " not found";
    return false;
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
    if (addr + 32 > buf_.size()) { err = "corrupt object header"; return false; }
    const uint8_t flags = buf_[addr + 5];
    uint64_t p = addr + 6;
    if (flags & 0x20) p += 16;
    if (flags & 0x10) p += 4;
    const int w = 1 << (flags & 3);
    uint64_t size0 = 0;
    if (p + w > buf_.size()) { err = "corrupt object header"; return false; }
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
          if (msize < 16) { err = "bad continuation message"; return false; }
          const uint64_t off = rd64(buf_.data() + q), ln = rd64(buf_.data() + q + 8);
          if (ln < 8 || off > buf_.size() || ln > buf_.size() - off || memcmp(buf_.data() + off, "OCHK", 4) != 0) { err = "bad continuation block"; return false; }
          blocks.push_back({off + 4, off + ln - 4});
        } else {
          out.push_back({mtype, {q, msize}});
        }
        q += msize;
      }
    }
    return true;
  }

  bool walk_chunks(const std::string& name, const DsInfo& d, uint64_t node, uint64_t chunk_bytes, std::vector<uint8_t>& raw, std::string& err,
                   int depth) const {
    const size_t rank = d.shape.size(), nd = rank + 1;
    if (depth > 8 || node + 24 > buf_.size() || memcmp(buf_.data() + node, "TREE", 4) != 0 || buf_[node + 4] != 1) {
      err = name + ": unsupported chunk index";
      return false;
    }
    const int level = buf_[node + 5];
    const int nent = rd16(buf_.data() + node + 6);
    uint64_t q = node + 8 + 16;   // left / right sibling addresses
    const uint64_t ksz = 8 + 8 * nd;
    for (int e = 0; e < nent; ++e) {
      if (q + ksz + 8 > buf_.size()) { err = name + ": corrupt chunk index"; return false; }
      const uint32_t csize = rd32(buf_.data() + q), mask = rd32(buf_.data() + q + 4);
      std::vector<uint64_t> off(nd);
      for (size_t k = 0; k < nd; ++k) off[k] = rd64(buf_.data() + q + 8 + 8 * k);
      const uint64_t child = rd64(buf_.data() + q + ksz);
      q += ksz + 8;
      if (level > 0) {
        if (!walk_chunks(name, d, child, chunk_bytes, raw, err, depth + 1)) return false;
        continue;
      }
      if (mask != 0 || csize != chunk_bytes) { err = name + ": filtered (compressed) chunks are not supported"; return false; }
      if (child + chunk_bytes > buf_.size()) { err = name + ": chunk out of file"; return false; }
      // copy the part of this chunk that lies inside the dataset, one innermost row at a time
      if (rank == 0) { memcpy(raw.data(), buf_.data() + child, d.size); continue; }
      const uint64_t inner = std::min<uint64_t>(d.chunk[rank - 1], d.shape[rank - 1] > off[rank - 1] ? d.shape[rank - 1] - off[rank - 1] : 0);
      if (inner == 0) continue;
      std::vector<uint64_t> id(rank, 0);   // index inside the chunk (innermost fixed at 0)
      for (;;) {
        bool inside = true;
        uint64_t src = 0, dst = 0, cstride = 1, dstride = 1;
        for (size_t k = rank; k-- > 0;) {
          const uint64_t g = off[k] + id[k];
          if (g >= d.shape[k]) inside = false;
          src += id[k] * cstride; cstride *= d.chunk[k];
          dst += g * dstride; dstride *= d.shape[k];
        }
        if (inside) memcpy(raw.data() + dst * d.size, buf_.data() + child + src * d.size, inner * d.size);
        size_t k = rank - 1;     // next row: increment from dimension rank-2 upwards
        for (;;) {
          if (k == 0) goto done;
          --k;
          if (++id[k] < d.chunk[k]) break;
          id[k] = 0;
        }
      }
    done:;
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
      // chunked: version-1 B-tree of raw-data chunks (leaves at level 0 point at chunks, inner nodes at nodes)
      const size_t nd = d.chunk.size();
      const size_t rank = d.shape.size();
      if (nd != rank + 1) { err = name + ": unexpected chunk rank"; return false; }
      uint64_t cn = 1;
      for (size_t k = 0; k < rank; ++k) cn *= d.chunk[k];
      raw.assign(nbytes, 0);
      if (!walk_chunks(name, d, d.addr, cn * d.size, raw, err, 0)) return false;
    }
    return true;
  }
};

}  // namespace nc4
}  // namespace rrnn
