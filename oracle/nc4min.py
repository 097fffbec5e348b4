"""Minimal netCDF-4 (HDF5) dataset reader -- TEST/ORACLE INFRASTRUCTURE ONLY.

Independent (pure numpy) twin of the product C++ reader in
rte_rrtmgp_nn_b200/csrc/nc4_reader.cpp; used to cross-check it and to build fixtures.
Handles exactly what the reference's data files use: version-2
object headers, dense links in a fractal heap, contiguous or single-chunk unfiltered
little-endian datasets.  The reference reads the same files through netCDF-Fortran
(/root/reference/neural/mod_network_rrtmgp.F90:58-122, examples/mo_simple_netcdf.F90).
"""
import struct
import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF


class NC4File:
    def __init__(self, path):
        with open(path, "rb") as fh:
            self.buf = fh.read()
        if self.buf[:8] != b"\x89HDF\r\n\x1a\n":
            raise ValueError(f"{path}: not an HDF5/netCDF-4 file")

    # -- link lookup: "<len><name><8-byte object header address>" ------------------
    def _find_ohdr(self, name):
        key = bytes([len(name)]) + name.encode()
        buf, pos = self.buf, 0
        while True:
            i = buf.find(key, pos)
            if i < 0:
                return None
            a = i + len(key)
            if a + 8 <= len(buf):
                addr = struct.unpack_from("<Q", buf, a)[0]
                if addr + 6 <= len(buf) and buf[addr:addr + 4] == b"OHDR" and buf[addr + 4] == 2:
                    return addr
            pos = i + 1

    def has(self, name):
        return self._find_ohdr(name) is not None

    # -- object header v2 message iterator ----------------------------------------
    def _messages(self, addr):
        buf = self.buf
        flags = buf[addr + 5]
        p = addr + 6
        if flags & 0x20:
            p += 16
        if flags & 0x10:
            p += 4
        w = 1 << (flags & 3)
        size0 = int.from_bytes(buf[p:p + w], "little")
        p += w
        blocks = [(p, p + size0)]
        track = bool(flags & 0x04)
        out = []
        while blocks:
            p, end = blocks.pop(0)
            while p + 4 <= end:
                mtype = buf[p]
                msize = struct.unpack_from("<H", buf, p + 1)[0]
                p += 4
                if track:
                    p += 2
                body = buf[p:p + msize]
                if mtype == 0x10:  # continuation: offset, length -> "OCHK" ... checksum
                    off, ln = struct.unpack_from("<QQ", body, 0)
                    if buf[off:off + 4] != b"OCHK":
                        raise ValueError("bad continuation block")
                    blocks.append((off + 4, off + ln - 4))
                else:
                    out.append((mtype, body))
                p += msize
        return out

    def info(self, name):
        addr = self._find_ohdr(name)
        if addr is None:
            raise KeyError(name)
        shape = None; dt = None; layout = None; filtered = False
        for mtype, b in self._messages(addr):
            if mtype == 0x01:
                ver, rank = b[0], b[1]
                off = 8 if ver == 1 else 4
                shape = tuple(struct.unpack_from("<Q", b, off + 8 * k)[0] for k in range(rank))
            elif mtype == 0x03:
                cls = b[0] & 0x0F
                size = struct.unpack_from("<I", b, 4)[0]
                big = bool(b[1] & 1)
                signed = bool(b[1] & 0x08)
                dt = (cls, size, big, signed)
            elif mtype == 0x08:
                ver, lc = b[0], b[1]
                if ver != 3:
                    raise ValueError("layout version %d unsupported" % ver)
                if lc == 1:
                    a, s = struct.unpack_from("<QQ", b, 2)
                    layout = ("contig", a, s)
                elif lc == 2:
                    nd = b[2]
                    bt = struct.unpack_from("<Q", b, 3)[0]
                    cd = struct.unpack_from("<%dI" % nd, b, 11)
                    layout = ("chunk", bt, cd)
                elif lc == 0:
                    sz = struct.unpack_from("<H", b, 2)[0]
                    layout = ("compact", bytes(b[4:4 + sz]))
            elif mtype == 0x0B:
                filtered = True
        return shape, dt, layout, filtered

    def read(self, name):
        shape, dt, layout, filtered = self.info(name)
        if filtered:
            raise ValueError(f"{name}: filtered datasets unsupported")
        cls, size, big, signed = dt
        if cls == 0:
            npdt = np.dtype(("i" if signed else "u") + str(size))
        elif cls == 1:
            npdt = np.dtype("f" + str(size))
        elif cls == 3:
            npdt = np.dtype("S" + str(size))
        else:
            raise ValueError(f"{name}: datatype class {cls} unsupported")
        if cls != 3:
            npdt = npdt.newbyteorder(">" if big else "<")
        n = int(np.prod(shape)) if shape else 1
        nbytes = n * npdt.itemsize
        if layout[0] == "contig":
            a = layout[1]
            if a == UNDEF:
                raise ValueError(f"{name}: no data allocated")
            raw = self.buf[a:a + nbytes]
        elif layout[0] == "compact":
            raw = layout[1][:nbytes]
        else:
            bt, cd = layout[1], layout[2]
            cshape = tuple(cd[:-1])
            if len(cshape) != len(shape):
                raise ValueError(f"{name}: chunk rank differs from dataset rank")
            b = self.buf
            nd = len(cd)
            cn = int(np.prod(cshape))
            out = np.zeros(shape, dtype=npdt)

            def walk(addr):  # version-1 B-tree of raw-data chunks: leaves (level 0) point at chunks, inner nodes at nodes
                if b[addr:addr + 4] != b"TREE" or b[addr + 4] != 1:
                    raise ValueError("unsupported chunk b-tree")
                level = b[addr + 5]
                nent = struct.unpack_from("<H", b, addr + 6)[0]
                q = addr + 8 + 16  # siblings
                for _ in range(nent):
                    size, mask = struct.unpack_from("<II", b, q)
                    offs = struct.unpack_from("<%dQ" % nd, b, q + 8)
                    child = struct.unpack_from("<Q", b, q + 8 + 8 * nd)[0]
                    q += 8 + 8 * nd + 8
                    if level > 0:
                        walk(child)
                        continue
                    if mask != 0 or size != cn * npdt.itemsize:
                        raise ValueError(f"{name}: filtered chunks unsupported")
                    carr = np.frombuffer(b[child:child + size], dtype=npdt, count=cn).reshape(cshape)
                    dst = tuple(slice(o, min(o + c, s_)) for o, c, s_ in zip(offs, cshape, shape))
                    src = tuple(slice(0, d.stop - d.start) for d in dst)
                    out[dst] = carr[src]

            walk(bt)
            raw = np.ascontiguousarray(out).tobytes()
        arr = np.frombuffer(raw, dtype=npdt, count=n).reshape(shape)
        if cls != 3:
            arr = arr.astype(npdt.newbyteorder("="))
        return arr

    def attr_str(self, name, attr):
        """String attribute (v1 attribute message) of a variable, e.g. RFMIP 'units'."""
        addr = self._find_ohdr(name)
        for mtype, b in self._messages(addr):
            if mtype != 0x0C or b[0] != 1:
                continue
            nsz, dsz, ssz = struct.unpack_from("<HHH", b, 2)
            pad = lambda v: (v + 7) & ~7
            aname = bytes(b[8:8 + nsz]).rstrip(b"\x00").decode()
            if aname != attr:
                continue
            p = 8 + pad(nsz)
            strsize = struct.unpack_from("<I", b, p + 4)[0]
            p += pad(dsz) + pad(ssz)
            return bytes(b[p:p + strsize]).rstrip(b"\x00").decode()
        raise KeyError(attr)

    def read_strings(self, name):
        a = self.read(name)
        if a.dtype.kind == "S" and a.dtype.itemsize == 1:
            a = a.reshape(a.shape[0], -1)
            return [b"".join(r).decode().strip("\x00 ").strip() for r in a]
        return [x.decode().strip("\x00 ").strip() for x in a.ravel()]


def load_nn_model(path):
    """Return dict(dims, W[list of (n_in,n_out) row-major], b[list], activations, input_names,
    xmin, xmax, ymean, ystd) -- the content load_netcdf reads
    (/root/reference/neural/mod_network_rrtmgp.F90:58-122)."""
    f = NC4File(path)
    dimsize = f.read("nn_dimsize").astype(np.int64).ravel()
    nlayers = len(dimsize)
    xmin = f.read("nn_input_coeffs_min").astype(np.float32).ravel()
    xmax = f.read("nn_input_coeffs_max").astype(np.float32).ravel()
    nx = len(xmin)
    dims = [nx] + [int(d) for d in dimsize]
    W, B = [], []
    for n in range(1, nlayers + 1):
        w = f.read("nn_weights_%d" % n).astype(np.float32)
        b = f.read("nn_bias_%d" % n).astype(np.float32).ravel()
        assert w.shape == (dims[n - 1], dims[n]), (w.shape, dims)
        W.append(np.ascontiguousarray(w)); B.append(b)
    acts = f.read_strings("nn_activation_char")
    names = f.read_strings("nn_inputs_char")
    out = dict(dims=dims, W=W, b=B, activations=acts, input_names=names, xmin=xmin, xmax=xmax,
               ymean=None, ystd=None)
    if f.has("nn_output_coeffs_mean"):
        out["ymean"] = f.read("nn_output_coeffs_mean").astype(np.float32).ravel()
    if f.has("nn_output_coeffs_std"):
        out["ystd"] = f.read("nn_output_coeffs_std").astype(np.float32).ravel()
    return out
