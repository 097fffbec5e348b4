// netCDF-4 file access for the callers either side of the hot path: what the reference's drivers do
// through netCDF-Fortran in examples/mo_simple_netcdf.F90 (read_field :34-96, var_exists :308-318, create_dim :320-343,
// create_var :345-380, write_field :167-237), examples/rfmip-clear-sky/mo_rfmip_io.F90 (read_and_block_* :185-680,
// unblock_and_write :734-870) and examples/all-sky/mo_garand_atmos_io.F90 (read_atmos :41-88, write_*_fluxes :92-170).
// Neither libnetcdf nor libhdf5 exists in the target image:
//   * reading goes through the HDF5-subset reader of nc4.hpp (the reference's own input files parse with it);
//   * rrnn_nc_create .. rrnn_nc_close WRITE a netCDF-4 (= HDF5) file from scratch with the structures netCDF-C itself emits for
//     these shapes, modelled byte for byte on the reference's data files: superblock v0 whose root entry points at a version-2
//     object header, link messages with creation order, one dimension-scale dataset per dimension (CLASS / NAME /
//     _Netcdf4Dimid), float32 little-endian contiguous variables with a DIMENSION_LIST attribute (variable-length object
//     references in a global heap collection) and an optional `units` string, Jenkins lookup3 checksums (checked against the
//     checksums stored in the reference's files, tests/test_ncio_cpu.py).  What it does not do: append to an existing file (the
//     reference's drivers write into template files that ship with RFMIP) and REFERENCE_LIST back-pointers on the scales
//     (netCDF-C does not need them to read a file).
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>
#include "../../include/rrnn.h"
#include "nc4.hpp"

namespace rrnn {
int fail(const std::string& msg);
}
using rrnn::fail;

namespace {

// ---------------------------------------------------------------------------------------------- little-endian byte sink
struct Bytes {
  std::vector<uint8_t> b;
  void u8(uint32_t v) { b.push_back((uint8_t)v); }
  void u16(uint32_t v) { u8(v & 255); u8((v >> 8) & 255); }
  void u32(uint32_t v) { u16(v & 65535); u16(v >> 16); }
  void u64(uint64_t v) { u32((uint32_t)v); u32((uint32_t)(v >> 32)); }
  void raw(const void* p, size_t n) { const uint8_t* q = (const uint8_t*)p; b.insert(b.end(), q, q + n); }
  void str(const std::string& s) { raw(s.data(), s.size()); }
  void zeros(size_t n) { b.insert(b.end(), n, 0); }
  void pad8() { while (b.size() % 8) u8(0); }
  void append(const Bytes& o) { b.insert(b.end(), o.b.begin(), o.b.end()); }
  size_t size() const { return b.size(); }
};

// Bob Jenkins' lookup3 hashlittle() as the HDF5 library uses it for metadata checksums (H5_checksum_lookup3, initval 0)
inline uint32_t rot(uint32_t x, int k) { return (x << k) | (x >> (32 - k)); }
uint32_t lookup3(const uint8_t* k, size_t length) {
  uint32_t a, b, c;
  a = b = c = 0xdeadbeefu + (uint32_t)length;
  auto rd = [&](size_t i, size_t n) {
    uint32_t v = 0;
    for (size_t j = 0; j < 4 && i + j < n; ++j) v |= (uint32_t)k[i + j] << (8 * j);
    return v;
  };
  while (length > 12) {
    a += rd(0, 12); b += rd(4, 12); c += rd(8, 12);
    a -= c; a ^= rot(c, 4); c += b;
    b -= a; b ^= rot(a, 6); a += c;
    c -= b; c ^= rot(b, 8); b += a;
    a -= c; a ^= rot(c, 16); c += b;
    b -= a; b ^= rot(a, 19); a += c;
    c -= b; c ^= rot(b, 4); b += a;
    length -= 12; k += 12;
  }
  if (length == 0) return c;
  a += rd(0, length); b += rd(4, length); c += rd(8, length);
  c ^= b; c -= rot(b, 14);
  a ^= c; a -= rot(c, 11);
  b ^= a; b -= rot(a, 25);
  c ^= b; c -= rot(b, 16);
  a ^= c; a -= rot(c, 4);
  b ^= a; b -= rot(a, 14);
  c ^= b; c -= rot(b, 24);
  return c;
}

const uint64_t UNDEF = 0xFFFFFFFFFFFFFFFFull;

// ---------------------------------------------------------------------------------------------- HDF5 message bodies
Bytes dt_float32() {   // IEEE little-endian binary32: the 20 bytes the reference's files carry for their float variables
  Bytes m;
  const uint8_t v[20] = {0x11, 0x20, 0x1f, 0x00, 4, 0, 0, 0, 0, 0, 0x20, 0, 0x17, 0x08, 0x00, 0x17, 0x7f, 0, 0, 0};
  m.raw(v, 20);
  return m;
}
Bytes dt_int32() {
  Bytes m;
  const uint8_t v[12] = {0x10, 0x08, 0x00, 0x00, 4, 0, 0, 0, 0, 0, 0x20, 0};
  m.raw(v, 12);
  return m;
}
Bytes dt_string(uint32_t size) {   // fixed-length, null-terminated, ASCII
  Bytes m;
  m.u8(0x13); m.u8(0); m.u8(0); m.u8(0); m.u32(size);
  return m;
}
Bytes dt_vlen_objref() {           // variable-length sequence of object references (DIMENSION_LIST)
  Bytes m;
  m.u8(0x19); m.u8(0); m.u8(0); m.u8(0); m.u32(16);
  m.u8(0x17); m.u8(0); m.u8(0); m.u8(0); m.u32(8);
  return m;
}
Bytes dataspace(const std::vector<uint64_t>& dims, bool with_max) {   // version 1; rank 0 = scalar
  Bytes m;
  m.u8(1); m.u8((uint32_t)dims.size()); m.u8(with_max && !dims.empty() ? 1 : 0); m.u8(0); m.u32(0);
  for (uint64_t d : dims) m.u64(d);
  if (with_max) for (uint64_t d : dims) m.u64(d);
  return m;
}
// version-1 attribute message: name / datatype / dataspace each padded to 8 bytes
Bytes attribute(const std::string& name, const Bytes& dt, const Bytes& ds, const Bytes& data) {
  Bytes m;
  m.u8(1); m.u8(0); m.u16((uint32_t)name.size() + 1); m.u16((uint32_t)dt.size()); m.u16((uint32_t)ds.size());
  m.str(name); m.u8(0); m.pad8();
  m.append(dt); m.pad8();
  m.append(ds); m.pad8();
  m.append(data);
  return m;
}
Bytes attr_string(const std::string& name, const std::string& value) {
  Bytes d;
  d.str(value); d.u8(0);
  return attribute(name, dt_string((uint32_t)value.size() + 1), dataspace({}, false), d);
}

struct Msg { int type; int flags; Bytes body; };

// version-2 object header with the flags of the reference's files (0x0d: 2-byte chunk size, attribute creation order tracked and
// indexed), every message prefixed by type / size / flags / creation order, closed by its lookup3 checksum
Bytes object_header(const std::vector<Msg>& msgs) {
  Bytes h;
  h.str("OHDR"); h.u8(2); h.u8(0x0d);
  size_t total = 0;
  for (auto& m : msgs) total += 6 + m.body.size();
  h.u16((uint32_t)total);
  uint32_t order = 0;
  for (auto& m : msgs) {
    h.u8(m.type); h.u16((uint32_t)m.body.size()); h.u8(m.flags);
    h.u16(m.type == 0x0C ? order++ : 0);
    h.append(m.body);
  }
  h.u32(lookup3(h.b.data(), h.b.size()));
  return h;
}
Msg attr_info(int nattrs) {   // message 0x15: no dense storage
  Bytes m;
  m.u8(0); m.u8(3); m.u16(nattrs); m.u64(UNDEF); m.u64(UNDEF); m.u64(UNDEF);
  return {0x15, 0x04, m};
}

}  // namespace

// ================================================================================================== the handle
struct rrnn_ncfile {
  // reading
  rrnn::nc4::Nc4File in;
  bool writing = false;
  // writing
  std::string path;
  struct Dim { std::string name; uint64_t len; };
  struct Var { std::string name; std::vector<int> dimids; std::vector<float> data; std::string units; };
  std::vector<Dim> dims;
  std::vector<Var> vars;
};

namespace {

int write_file(const rrnn_ncfile& f) {
  const size_t nd = f.dims.size(), nv = f.vars.size(), nobj = nd + nv;
  // global heap: one 8-byte object reference per (variable, dimension) pair; object indices start at 1
  std::vector<std::vector<uint32_t>> heap_index(nv);
  uint32_t nrefs = 0;
  for (size_t v = 0; v < nv; ++v)
    for (size_t k = 0; k < f.vars[v].dimids.size(); ++k) heap_index[v].push_back(++nrefs);
  const uint64_t heap_size = std::max<uint64_t>(4096, (16 + (uint64_t)nrefs * 24 + 16 + 7) & ~7ull);

  // two passes: object-header sizes do not depend on the addresses they contain
  std::vector<uint64_t> ohdr_addr(nobj, 0), data_addr(nobj, 0);
  uint64_t heap_addr = 0, root_addr = 96, eof = 0;
  std::vector<Bytes> ohdr(nobj);
  Bytes root;
  for (int pass = 0; pass < 2; ++pass) {
    // root group: link info (creation order tracked + indexed, compact storage), group info, one hard link per object
    {
      std::vector<Msg> m;
      Bytes li;
      li.u8(0); li.u8(3); li.u64(nobj); li.u64(UNDEF); li.u64(UNDEF); li.u64(UNDEF);
      m.push_back({0x02, 0, li});
      Bytes gi;
      gi.u8(0); gi.u8(0);
      m.push_back({0x0A, 0x01, gi});
      m.push_back(attr_info(1));
      m.push_back({0x0C, 0, attr_string("_NCProperties", "version=2,rrnn_b200=1")});
      for (size_t i = 0; i < nobj; ++i) {
        const std::string& nm = i < nd ? f.dims[i].name : f.vars[i - nd].name;
        Bytes l;
        l.u8(1); l.u8(0x04); l.u64(i); l.u8((uint32_t)nm.size()); l.str(nm); l.u64(ohdr_addr[i]);
        m.push_back({0x06, 0, l});
      }
      root = object_header(m);
    }
    for (size_t i = 0; i < nobj; ++i) {
      std::vector<Msg> m;
      std::vector<uint64_t> shape;
      uint64_t nbytes;
      if (i < nd) {
        shape = {f.dims[i].len};
      } else {
        for (int d : f.vars[i - nd].dimids) shape.push_back(f.dims[d].len);
      }
      nbytes = 4;
      for (uint64_t s : shape) nbytes *= s;
      m.push_back({0x01, 0, dataspace(shape, true)});
      m.push_back({0x03, 0x01, dt_float32()});
      Bytes fv;   // fill value, version 2: allocation late, written if set, none defined
      fv.u8(2); fv.u8(2); fv.u8(2); fv.u8(0);
      m.push_back({0x05, 0x01, fv});
      Bytes lay;  // layout version 3, contiguous
      lay.u8(3); lay.u8(1); lay.u64(data_addr[i]); lay.u64(nbytes);
      m.push_back({0x08, 0, lay});
      if (i < nd) {
        char nm[96];
        snprintf(nm, sizeof nm, "This is a netCDF dimension but not a netCDF variable.%10d", (int)i);
        Bytes id;
        id.u32((uint32_t)i);
        m.push_back(attr_info(3));
        m.push_back({0x0C, 0, attr_string("CLASS", "DIMENSION_SCALE")});
        m.push_back({0x0C, 0, attr_string("NAME", nm)});
        m.push_back({0x0C, 0, attribute("_Netcdf4Dimid", dt_int32(), dataspace({}, false), id)});
      } else {
        const auto& v = f.vars[i - nd];
        m.push_back(attr_info(v.units.empty() ? 1 : 2));
        Bytes dl;   // one variable-length element per dimension: length 1, heap collection address, object index
        for (size_t k = 0; k < v.dimids.size(); ++k) { dl.u32(1); dl.u64(heap_addr); dl.u32(heap_index[i - nd][k]); }
        m.push_back({0x0C, 0, attribute("DIMENSION_LIST", dt_vlen_objref(), dataspace({(uint64_t)v.dimids.size()}, true), dl)});
        if (!v.units.empty()) m.push_back({0x0C, 0, attr_string("units", v.units)});
      }
      ohdr[i] = object_header(m);
    }
    // layout of the file
    uint64_t p = root_addr + root.size();
    for (size_t i = 0; i < nobj; ++i) { p = (p + 7) & ~7ull; ohdr_addr[i] = p; p += ohdr[i].size(); }
    p = (p + 7) & ~7ull;
    heap_addr = p;
    p += heap_size;
    for (size_t i = 0; i < nobj; ++i) {
      p = (p + 7) & ~7ull;
      data_addr[i] = p;
      uint64_t nbytes = 4;
      if (i < nd) nbytes *= f.dims[i].len;
      else nbytes = 4 * (uint64_t)f.vars[i - nd].data.size();
      p += nbytes;
    }
    eof = p;
  }

  std::vector<uint8_t> out(eof, 0);
  auto put = [&](uint64_t at, const Bytes& b) { memcpy(out.data() + at, b.b.data(), b.size()); };
  // superblock version 0 (the reference's files: 8-byte offsets and lengths, group leaf / internal node K = 4 / 16)
  {
    Bytes s;
    const uint8_t sig[8] = {0x89, 'H', 'D', 'F', '\r', '\n', 0x1a, '\n'};
    s.raw(sig, 8);
    s.u8(0); s.u8(0); s.u8(0); s.u8(0);     // superblock, free-space, root symbol-table versions, reserved
    s.u8(0); s.u8(8); s.u8(8); s.u8(0);     // shared-header version, size of offsets, size of lengths, reserved
    s.u16(4); s.u16(16); s.u32(0);          // group leaf node K, internal node K, consistency flags
    s.u64(0); s.u64(UNDEF); s.u64(eof); s.u64(UNDEF);   // base, free-space info, end of file, driver info
    s.u64(0); s.u64(root_addr); s.u32(0); s.u32(0); s.zeros(16);   // root symbol-table entry
    put(0, s);
  }
  put(root_addr, root);
  for (size_t i = 0; i < nobj; ++i) put(ohdr_addr[i], ohdr[i]);
  {
    Bytes g;   // global heap collection
    g.str("GCOL"); g.u8(1); g.zeros(3); g.u64(heap_size);
    for (size_t v = 0; v < nv; ++v)
      for (size_t k = 0; k < f.vars[v].dimids.size(); ++k) {
        g.u16(heap_index[v][k]); g.u16(0); g.u32(0); g.u64(8);
        g.u64(ohdr_addr[f.vars[v].dimids[k]]);   // the object reference: address of the dimension scale's header
      }
    const uint64_t left = heap_size - g.size();
    g.u16(0); g.u16(0); g.u32(0); g.u64(left);   // object 0: the free space (its size counts this 16-byte header)
    put(heap_addr, g);
  }
  for (size_t v = 0; v < nv; ++v)
    if (!f.vars[v].data.empty()) memcpy(out.data() + data_addr[nd + v], f.vars[v].data.data(), 4 * f.vars[v].data.size());
  FILE* fp = fopen(f.path.c_str(), "wb");
  if (!fp) return fail("rrnn_nc_close: cannot write " + f.path);
  const size_t n = fwrite(out.data(), 1, out.size(), fp);
  fclose(fp);
  if (n != out.size()) return fail("rrnn_nc_close: short write to " + f.path);
  return 0;
}

}  // namespace

// ================================================================================================== C ABI
extern "C" int rrnn_nc_open(const char* path, rrnn_ncfile_t** out) {
  if (!path || !out) return fail("rrnn_nc_open: null argument");
  rrnn_ncfile* f = new rrnn_ncfile;
  std::string err;
  if (!f->in.open(path, err)) { delete f; return fail(err); }
  *out = f;
  return 0;
}

extern "C" int rrnn_nc_create(const char* path, rrnn_ncfile_t** out) {
  if (!path || !out) return fail("rrnn_nc_create: null argument");
  rrnn_ncfile* f = new rrnn_ncfile;
  f->writing = true;
  f->path = path;
  *out = f;
  return 0;
}

extern "C" int rrnn_nc_close(rrnn_ncfile_t* f) {
  if (!f) return 0;
  int rc = 0;
  if (f->writing) rc = write_file(*f);
  delete f;
  return rc;
}

extern "C" int rrnn_nc_var_exists(const rrnn_ncfile_t* f, const char* name) {
  if (!f || !name) return 0;
  if (f->writing) {
    for (auto& v : f->vars) if (v.name == name) return 1;
    return 0;
  }
  return f->in.has(name) ? 1 : 0;
}

extern "C" int rrnn_nc_inq_var(const rrnn_ncfile_t* f, const char* name, int* ndims, long long* shape) {
  if (!f || !name || !ndims || !shape) return fail("rrnn_nc_inq_var: null argument");
  if (f->writing) return fail("rrnn_nc_inq_var: file is open for writing");
  rrnn::nc4::DsInfo d;
  std::string err;
  if (!f->in.info(name, d, err)) return fail(err);
  if (d.shape.size() > 8) return fail(std::string(name) + ": more than 8 dimensions");
  *ndims = (int)d.shape.size();
  for (size_t k = 0; k < d.shape.size(); ++k) shape[k] = (long long)d.shape[k];
  return 0;
}

extern "C" int rrnn_nc_get_var_float(const rrnn_ncfile_t* f, const char* name, float* out, size_t n) {
  if (!f || !name || !out) return fail("rrnn_nc_get_var_float: null argument");
  if (f->writing) return fail("rrnn_nc_get_var_float: file is open for writing");
  std::vector<float> v;
  std::vector<uint64_t> shape;
  std::string err;
  if (!f->in.read_float(name, v, shape, err)) return fail(err);
  if (v.size() != n) return fail(std::string(name) + ": has " + std::to_string(v.size()) + " elements, the caller expects " + std::to_string(n));
  memcpy(out, v.data(), 4 * n);
  return 0;
}

extern "C" int rrnn_nc_get_att_text(const rrnn_ncfile_t* f, const char* var, const char* att, char* buf, int nbuf) {
  if (!f || !var || !att || !buf || nbuf < 1) return fail("rrnn_nc_get_att_text: bad argument");
  if (f->writing) return fail("rrnn_nc_get_att_text: file is open for writing");
  std::string s, err;
  if (!f->in.attr_string(var, att, s, err)) return fail(err);
  if ((int)s.size() + 1 > nbuf) return fail(std::string(var) + ":" + att + ": buffer too small");
  memcpy(buf, s.c_str(), s.size() + 1);
  return 0;
}

extern "C" int rrnn_nc_def_dim(rrnn_ncfile_t* f, const char* name, long long len, int* dimid) {
  if (!f || !name || !dimid) return fail("rrnn_nc_def_dim: null argument");
  if (!f->writing) return fail("rrnn_nc_def_dim: file is open for reading");
  if (len < 1 || strlen(name) == 0 || strlen(name) > 200) return fail("rrnn_nc_def_dim: bad dimension");
  for (size_t i = 0; i < f->dims.size(); ++i)
    if (f->dims[i].name == name) {   // create_dim (mo_simple_netcdf.F90:320-343): an existing dimension must have the same length
      if ((long long)f->dims[i].len != len) return fail(std::string("dim ") + name + " is present but incorrectly sized");
      *dimid = (int)i;
      return 0;
    }
  f->dims.push_back({name, (uint64_t)len});
  *dimid = (int)f->dims.size() - 1;
  return 0;
}

extern "C" int rrnn_nc_put_var_float(rrnn_ncfile_t* f, const char* name, int ndims, const int* dimids, const float* data,
                                     const char* units) {
  if (!f || !name || !data || (ndims > 0 && !dimids)) return fail("rrnn_nc_put_var_float: null argument");
  if (!f->writing) return fail("rrnn_nc_put_var_float: file is open for reading");
  if (ndims < 1 || ndims > 8) return fail("rrnn_nc_put_var_float: 1 to 8 dimensions");
  size_t n = 1;
  rrnn_ncfile::Var v;
  v.name = name;
  for (int k = 0; k < ndims; ++k) {
    if (dimids[k] < 0 || dimids[k] >= (int)f->dims.size()) return fail("rrnn_nc_put_var_float: unknown dimension id");
    v.dimids.push_back(dimids[k]);
    n *= f->dims[dimids[k]].len;
  }
  for (auto& d : f->dims) if (d.name == name) return fail(std::string(name) + ": coordinate variables are not supported by this writer");
  for (auto& o : f->vars) if (o.name == name) return fail(std::string("variable ") + name + " exists");
  v.data.assign(data, data + n);
  if (units) v.units = units;
  f->vars.push_back(std::move(v));
  return 0;
}
