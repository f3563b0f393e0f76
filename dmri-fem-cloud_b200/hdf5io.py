"""Minimal pure-Python HDF5 reader / writer for the solver's input container.

The reference reads its input with DOLFIN's HDF5File (GCloudDmriSolver.py:150-177: `mesh`, then the DG0 functions
`d00..d22`, `T2`, `phase`) and the pre-processing scripts write it (PreprocessingMultiCompt.py:148-152: `mesh, T2,
ic, phase, d00..d22`).  There is no libhdf5 / h5py in this image, so this module implements the subset of the HDF5
file format (HDF5 File Format Specification, version 3.0) those files use:

  read   superblock v0-v3; object headers v1 and v2 (continuation blocks); old-style groups (symbol table message ->
         v1 B-tree + local heap + symbol-table nodes) and compact new-style groups (link messages); datasets with
         contiguous, compact or chunked (v1 B-tree, optional deflate / shuffle filters) layout; fixed-point, IEEE
         float and fixed-length string datatypes; attributes (message v1-v3).
  write  superblock v0, v1 object headers, old-style groups, contiguous datasets, attributes -- what libhdf5 >= 1.8
         writes with its default (earliest) format, which is what DOLFIN produces.

and on top of it the DOLFIN layout (dolfin/io/HDF5File.cpp, third party, 2017.2-2019.1):
  /mesh/coordinates (nv, gdim) f64        /mesh/topology (nc, tdim+1) i64, attribute celltype = "tetrahedron" | ...
  /<f>/vector_0 (ndof,) f64   /<f>/cell_dofs   /<f>/x_cell_dofs (nc+1,)   /<f>/cells (nc,)   attribute <f>:signature
  DG0 value of mesh cell cells[i] = vector_0[cell_dofs[x_cell_dofs[i]]].

Not validated against libhdf5 itself (absent here): the writer follows the published layout field by field and the
reader is exercised on the writer's output and on hand-built variants (tests/test_hdf5io.py).
"""
import struct
import zlib

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF
SIGNATURE = b"\x89HDF\r\n\x1a\n"


class HDF5Error(RuntimeError):
    pass


# ===================================================================================================== reader

class _Datatype:
    def __init__(self, dtype, size, is_string=False):
        self.dtype, self.size, self.is_string = dtype, size, is_string


def _parse_datatype(buf, off):
    cv, b0, b1, b2, size = struct.unpack_from("<BBBBI", buf, off)
    cls = cv & 0x0F
    if cls == 0:      # fixed point
        order = ">" if (b0 & 1) else "<"
        signed = bool(b0 & 0x08)
        if size not in (1, 2, 4, 8):
            raise HDF5Error("unsupported integer size %d" % size)
        return _Datatype(np.dtype(order + ("i" if signed else "u") + str(size)), size)
    if cls == 1:      # floating point (IEEE layouts only)
        order = ">" if (b0 & 1) else "<"
        if size not in (4, 8):
            raise HDF5Error("unsupported float size %d" % size)
        return _Datatype(np.dtype(order + "f" + str(size)), size)
    if cls == 3:      # fixed-length string
        return _Datatype(np.dtype("S%d" % size), size, True)
    raise HDF5Error("unsupported datatype class %d" % cls)


def _parse_dataspace(buf, off):
    ver = buf[off]
    rank = buf[off + 1]
    if ver == 1:
        p = off + 8
    elif ver == 2:
        if buf[off + 3] == 2:
            return None      # null dataspace
        p = off + 4
    else:
        raise HDF5Error("unsupported dataspace version %d" % ver)
    return tuple(struct.unpack_from("<%dQ" % rank, buf, p)) if rank else ()


class _Object:
    """Parsed object header: messages as (type, flags, payload bytes)."""

    def __init__(self, f, addr):
        self.f, self.addr = f, addr
        self.msgs = []
        b = f.buf
        if b[addr:addr + 4] == b"OHDR":
            self._parse_v2(addr)
        else:
            self._parse_v1(addr)

    def _parse_v1(self, addr):
        b = self.f.buf
        ver, _, nmsg, _, hsize = struct.unpack_from("<BBHII", b, addr)
        if ver != 1:
            raise HDF5Error("bad object header version %d at %d" % (ver, addr))
        blocks = [(addr + 16, hsize)]
        count = 0
        while blocks and count < nmsg:
            p, size = blocks.pop(0)
            end = p + size
            while p + 8 <= end and count < nmsg:
                mtype, msize, mflags = struct.unpack_from("<HHB", b, p)
                data = b[p + 8:p + 8 + msize]
                p += 8 + msize
                count += 1
                if mtype == 0x10:
                    caddr, clen = struct.unpack_from("<QQ", data, 0)
                    blocks.append((self.f.base + caddr, clen))
                else:
                    self.msgs.append((mtype, mflags, data))

    def _parse_v2(self, addr):
        b = self.f.buf
        ver, flags = b[addr + 4], b[addr + 5]
        if ver != 2:
            raise HDF5Error("bad object header version")
        p = addr + 6
        if flags & 0x20:
            p += 16
        if flags & 0x10:
            p += 4
        nb = 1 << (flags & 3)
        size = int.from_bytes(b[p:p + nb], "little")
        p += nb
        blocks = [(p, size)]
        track = bool(flags & 0x04)
        while blocks:
            p, size = blocks.pop(0)
            end = p + size
            while p + 4 <= end:
                mtype = b[p]
                msize, mflags = struct.unpack_from("<HB", b, p + 1)
                p += 4 + (2 if track else 0)
                data = b[p:p + msize]
                p += msize
                if mtype == 0x10:
                    caddr, clen = struct.unpack_from("<QQ", data, 0)
                    blocks.append((self.f.base + caddr + 4, clen - 8))     # skip "OCHK", drop the checksum
                elif mtype != 0:
                    self.msgs.append((mtype, mflags, data))

    def find(self, mtype):
        return [m[2] for m in self.msgs if m[0] == mtype]

    # ---- attributes
    def attrs(self):
        out = {}
        for data in self.find(0x0C):
            ver = data[0]
            if ver == 1:
                nsz, tsz, ssz = struct.unpack_from("<HHH", data, 2)
                p = 8
                pad = lambda n: (n + 7) & ~7
                name = bytes(data[p:p + nsz]).split(b"\0")[0].decode()
                p += pad(nsz)
                dt = _parse_datatype(data, p)
                p += pad(tsz)
                shape = _parse_dataspace(data, p)
                p += pad(ssz)
            elif ver in (2, 3):
                nsz, tsz, ssz = struct.unpack_from("<HHH", data, 2)
                p = 8 + (1 if ver == 3 else 0)
                name = bytes(data[p:p + nsz]).split(b"\0")[0].decode()
                p += nsz
                dt = _parse_datatype(data, p)
                p += tsz
                shape = _parse_dataspace(data, p)
                p += ssz
            else:
                raise HDF5Error("unsupported attribute version %d" % ver)
            n = int(np.prod(shape)) if shape else 1
            arr = np.frombuffer(bytes(data[p:p + n * dt.size]), dtype=dt.dtype, count=n)
            if dt.is_string:
                vals = [v.split(b"\0")[0].decode(errors="replace") for v in arr.tolist()]
                out[name] = vals[0] if not shape else vals
            else:
                out[name] = arr[0] if not shape else arr.reshape(shape)
        return out

    # ---- groups
    def links(self):
        """name -> object header address of the members (None if this object is not a group)."""
        st = self.find(0x11)
        if st:
            btree, heap = struct.unpack_from("<QQ", st[0], 0)
            return self.f._symbol_table(self.f.base + btree, self.f.base + heap)
        ln = self.find(0x06)
        if ln or self.find(0x02):
            info = self.find(0x02)
            if info:
                d = info[0]
                p = 2 + (8 if d[1] & 1 else 0)
                fheap = struct.unpack_from("<Q", d, p)[0]
                if fheap != UNDEF:
                    raise HDF5Error("dense (fractal-heap) groups are not supported")
            out = {}
            for d in ln:
                fl = d[1]
                p = 2
                ltype = 0
                if fl & 0x08:
                    ltype = d[p]
                    p += 1
                if fl & 0x04:
                    p += 8
                if fl & 0x10:
                    p += 1
                nb = 1 << (fl & 3)
                nlen = int.from_bytes(d[p:p + nb], "little")
                p += nb
                name = bytes(d[p:p + nlen]).decode()
                p += nlen
                if ltype == 0:
                    out[name] = self.f.base + struct.unpack_from("<Q", d, p)[0]
            return out
        return None

    # ---- datasets
    def is_dataset(self):
        return bool(self.find(0x08))

    def read(self):
        b = self.f.buf
        dt = _parse_datatype(self.find(0x03)[0], 0)
        shape = _parse_dataspace(self.find(0x01)[0], 0)
        if shape is None:
            return np.zeros(0, dt.dtype)
        n = int(np.prod(shape)) if shape else 1
        lay = self.find(0x08)[0]
        ver = lay[0]
        if ver == 3:
            cls = lay[1]
            if cls == 1:
                addr, size = struct.unpack_from("<QQ", lay, 2)
                if addr == UNDEF:
                    arr = np.zeros(n, dt.dtype)
                else:
                    arr = np.frombuffer(b, dtype=dt.dtype, count=n, offset=self.f.base + addr)
            elif cls == 0:
                size = struct.unpack_from("<H", lay, 2)[0]
                arr = np.frombuffer(bytes(lay[4:4 + size]), dtype=dt.dtype, count=n)
            elif cls == 2:
                nd = lay[2]
                baddr = struct.unpack_from("<Q", lay, 3)[0]
                cdims = struct.unpack_from("<%dI" % nd, lay, 11)
                arr = self._read_chunked(dt, shape, self.f.base + baddr, cdims[:-1])
            else:
                raise HDF5Error("unsupported layout class %d" % cls)
        elif ver in (1, 2):
            nd, cls = lay[1], lay[2]
            p = 8
            addr = UNDEF
            if cls != 0:
                addr = struct.unpack_from("<Q", lay, p)[0]
                p += 8
            dims = struct.unpack_from("<%dI" % nd, lay, p)
            p += 4 * nd
            if cls == 1:
                arr = np.frombuffer(b, dtype=dt.dtype, count=n, offset=self.f.base + addr)
            elif cls == 2:
                arr = self._read_chunked(dt, shape, self.f.base + addr, dims[:-1])
            else:
                size = struct.unpack_from("<I", lay, p)[0]
                arr = np.frombuffer(bytes(lay[p + 4:p + 4 + size]), dtype=dt.dtype, count=n)
        else:
            raise HDF5Error("unsupported data layout version %d" % ver)
        arr = np.array(arr).reshape(shape)
        if dt.is_string:
            return np.array([v.split(b"\0")[0].decode(errors="replace") for v in arr.ravel().tolist()]).reshape(shape)
        return arr

    def _filters(self):
        out = []
        for d in self.find(0x0B):
            ver, nf = d[0], d[1]
            p = 8 if ver == 1 else 2
            for _ in range(nf):
                fid = struct.unpack_from("<H", d, p)[0]
                if ver == 1 or fid >= 256:
                    nlen = struct.unpack_from("<H", d, p + 2)[0]
                    p += 4
                else:
                    nlen = 0
                    p += 2
                _, ncv = struct.unpack_from("<HH", d, p)
                p += 4
                if nlen:
                    p += (nlen + 7) & ~7 if ver == 1 else nlen
                cvals = struct.unpack_from("<%dI" % ncv, d, p)
                p += 4 * ncv
                if ver == 1 and ncv % 2:
                    p += 4
                out.append((fid, cvals))
        return out

    def _read_chunked(self, dt, shape, baddr, cdims):
        b = self.f.buf
        nd = len(shape)
        out = np.zeros(shape, dt.dtype)
        filters = self._filters()

        def walk(addr):
            if b[addr:addr + 4] != b"TREE":
                raise HDF5Error("bad chunk B-tree node")
            _, level, used = struct.unpack_from("<BBH", b, addr + 4)
            p = addr + 24
            ksize = 8 + 8 * (nd + 1)
            for i in range(used):
                csize, fmask = struct.unpack_from("<II", b, p)
                offs = struct.unpack_from("<%dQ" % (nd + 1), b, p + 8)
                child = self.f.base + struct.unpack_from("<Q", b, p + ksize)[0]
                p += ksize + 8
                if level > 0:
                    walk(child)
                    continue
                raw = bytes(b[child:child + csize])
                for k, (fid, cv) in reversed(list(enumerate(filters))):
                    if fmask >> k & 1:
                        continue
                    if fid == 1:
                        raw = zlib.decompress(raw)
                    elif fid == 2:
                        es = cv[0] if cv else dt.size
                        raw = np.frombuffer(raw, np.uint8).reshape(es, -1).T.tobytes()
                    else:
                        raise HDF5Error("unsupported filter %d" % fid)
                chunk = np.frombuffer(raw, dtype=dt.dtype, count=int(np.prod(cdims))).reshape(cdims)
                sl = tuple(slice(o, min(o + c, s)) for o, c, s in zip(offs[:nd], cdims, shape))
                out[sl] = chunk[tuple(slice(0, s.stop - s.start) for s in sl)]

        if baddr != self.f.base + UNDEF and baddr < len(b):
            walk(baddr)
        return out


class File:
    """Read-only view of an HDF5 file: `f["mesh/coordinates"]` -> ndarray, `f.attrs("mesh/topology")` -> dict,
    `f.keys("")` -> member names of a group."""

    def __init__(self, path):
        with open(path, "rb") as fh:
            self.buf = memoryview(fh.read())
        b = self.buf
        start = 0
        while True:      # the superblock may sit at 0, 512, 1024, ...
            if start + 8 <= len(b) and bytes(b[start:start + 8]) == SIGNATURE:
                break
            start = 512 if start == 0 else start * 2
            if start >= len(b):
                raise HDF5Error("not an HDF5 file: " + path)
        ver = b[start + 8]
        if ver in (0, 1):
            if b[start + 13] != 8 or b[start + 14] != 8:
                raise HDF5Error("only 8-byte offsets and lengths are supported")
            p = start + 24 + (4 if ver == 1 else 0)
            self.base = struct.unpack_from("<Q", b, p)[0]
            root_entry = p + 32
            self.root = self.base + struct.unpack_from("<Q", b, root_entry + 8)[0]
        elif ver in (2, 3):
            if b[start + 9] != 8 or b[start + 10] != 8:
                raise HDF5Error("only 8-byte offsets and lengths are supported")
            self.base = struct.unpack_from("<Q", b, start + 12)[0]
            self.root = self.base + struct.unpack_from("<Q", b, start + 36)[0]
        else:
            raise HDF5Error("unsupported superblock version %d" % ver)

    def _symbol_table(self, btree, heap):
        b = self.buf
        if bytes(b[heap:heap + 4]) != b"HEAP":
            raise HDF5Error("bad local heap")
        data = self.base + struct.unpack_from("<Q", b, heap + 24)[0]
        out = {}

        def name_at(off):
            e = data + off
            end = e
            while b[end] != 0:
                end += 1
            return bytes(b[e:end]).decode()

        def walk(addr):
            sig = bytes(b[addr:addr + 4])
            if sig == b"TREE":
                _, level, used = struct.unpack_from("<BBH", b, addr + 4)
                p = addr + 24 + 8
                for _ in range(used):
                    walk(self.base + struct.unpack_from("<Q", b, p)[0])
                    p += 16
            elif sig == b"SNOD":
                nsym = struct.unpack_from("<H", b, addr + 6)[0]
                p = addr + 8
                for _ in range(nsym):
                    noff, oaddr = struct.unpack_from("<QQ", b, p)
                    out[name_at(noff)] = self.base + oaddr
                    p += 40
            else:
                raise HDF5Error("bad group node")

        walk(btree)
        return out

    def _resolve(self, path):
        obj = _Object(self, self.root)
        for part in [p for p in path.split("/") if p]:
            links = obj.links()
            if links is None or part not in links:
                raise KeyError(path)
            obj = _Object(self, links[part])
        return obj

    def __contains__(self, path):
        try:
            self._resolve(path)
            return True
        except KeyError:
            return False

    def __getitem__(self, path):
        obj = self._resolve(path)
        if not obj.is_dataset():
            raise KeyError(path + " is a group")
        return obj.read()

    def keys(self, path=""):
        links = self._resolve(path).links()
        return sorted(links) if links is not None else []

    def attrs(self, path=""):
        return self._resolve(path).attrs()


# ===================================================================================================== writer

def _pad8(b):
    return b + b"\0" * (-len(b) % 8)


def _dtype_message(dt):
    dt = np.dtype(dt)
    if dt.kind in "iu":
        bits = 0x08 if dt.kind == "i" else 0
        return struct.pack("<BBBBIHH", 0x10, bits, 0, 0, dt.itemsize, 0, 8 * dt.itemsize)
    if dt.kind == "f" and dt.itemsize == 8:
        return struct.pack("<BBBBIHHBBBBI", 0x11, 0x20, 63, 0, 8, 0, 64, 52, 11, 0, 52, 1023)
    if dt.kind == "f" and dt.itemsize == 4:
        return struct.pack("<BBBBIHHBBBBI", 0x11, 0x20, 31, 0, 4, 0, 32, 23, 8, 0, 23, 127)
    if dt.kind == "S":
        return struct.pack("<BBBBI", 0x13, 0x00, 0, 0, dt.itemsize)      # null-terminated ASCII
    raise HDF5Error("cannot write dtype %s" % dt)


def _dataspace_message(shape):
    return struct.pack("<BBBBI", 1, len(shape), 0, 0, 0) + b"".join(struct.pack("<Q", s) for s in shape)


def _message(mtype, data, flags=0):
    data = _pad8(data)
    return struct.pack("<HHBBBB", mtype, len(data), flags, 0, 0, 0) + data


def _attribute_message(name, value):
    if isinstance(value, str):
        raw = value.encode() + b"\0"
        dtm, dsm, data = _dtype_message(np.dtype("S%d" % len(raw))), _dataspace_message(()), raw
    else:
        arr = np.asarray(value)
        if arr.dtype.kind == "f":
            arr = arr.astype("<f8")
        elif arr.dtype.kind in "iu":
            arr = arr.astype("<i8" if arr.dtype.kind == "i" else "<u8")
        dtm, dsm, data = _dtype_message(arr.dtype), _dataspace_message(arr.shape), arr.tobytes()
    nm = name.encode() + b"\0"
    body = struct.pack("<BBHHH", 1, 0, len(nm), len(dtm), len(dsm)) + _pad8(nm) + _pad8(dtm) + _pad8(dsm) + data
    return _message(0x0C, body)


def _object_header(messages):
    body = b"".join(messages)
    return struct.pack("<BBHII", 1, 0, len(messages), 1, len(body)) + b"\0" * 4 + body


class Writer:
    """Builds an HDF5 file in memory: `w.dataset("mesh/coordinates", array, attrs={...})`,
    `w.group_attrs("T2", signature="...")`, then `w.save(path)`.  Intermediate groups are created on demand."""

    LEAF_K, INTERNAL_K = 4, 16

    def __init__(self):
        self.tree = {"children": {}, "attrs": {}}

    def _node(self, path, create=True):
        node = self.tree
        for part in [p for p in path.split("/") if p]:
            if part not in node["children"]:
                if not create:
                    raise KeyError(path)
                node["children"][part] = {"children": {}, "attrs": {}}
            node = node["children"][part]
        return node

    def dataset(self, path, array, attrs=None):
        parts = [p for p in path.split("/") if p]
        parent = self._node("/".join(parts[:-1]))
        arr = np.ascontiguousarray(array)
        if arr.dtype.byteorder == ">":
            arr = arr.astype(arr.dtype.newbyteorder("<"))
        parent["children"][parts[-1]] = {"data": arr, "attrs": dict(attrs or {})}

    def group_attrs(self, path, **attrs):
        self._node(path)["attrs"].update(attrs)

    def save(self, path):
        out = bytearray(96)      # superblock v0 is written last

        def alloc(data):
            while len(out) % 8:
                out.append(0)
            addr = len(out)
            out.extend(data)
            return addr

        def write_dataset(node):
            arr = node["data"]
            daddr = alloc(arr.tobytes()) if arr.size else UNDEF
            msgs = [_message(0x01, _dataspace_message(arr.shape)),
                    _message(0x03, _dtype_message(arr.dtype), flags=1),
                    _message(0x05, struct.pack("<BBBB", 2, 2, 2, 0)),                  # fill value: late alloc, undefined
                    _message(0x08, struct.pack("<BBQQ", 3, 1, daddr, arr.nbytes))]      # contiguous
            msgs += [_attribute_message(k, v) for k, v in node["attrs"].items()]
            return alloc(_object_header(msgs))

        def write_group(node):
            entries = []      # (name, object header address, cache type, scratch)
            for name in sorted(node["children"], key=lambda s: s.encode()):
                child = node["children"][name]
                if "data" in child:
                    entries.append((name, write_dataset(child), 0, b"\0" * 16))
                else:
                    oaddr, btree, heap = write_group(child)
                    entries.append((name, oaddr, 1, struct.pack("<QQ", btree, heap)))
            # local heap: the empty string at offset 0, then the member names
            heap_data = bytearray(8)
            offs = []
            for name, *_ in entries:
                offs.append(len(heap_data))
                heap_data.extend(_pad8(name.encode() + b"\0"))
            free_head = len(heap_data)
            heap_data.extend(struct.pack("<QQ", 1, 16))      # one free block closes the segment (1 = end of list)
            hdata_addr = alloc(bytes(heap_data))
            heap_addr = alloc(b"HEAP" + struct.pack("<BBBBQQQ", 0, 0, 0, 0, len(heap_data), free_head, hdata_addr))
            # symbol-table nodes of <= 2*LEAF_K entries under one level-0 B-tree node
            cap = 2 * self.LEAF_K
            nodes, keys = [], [0]
            for i in range(0, max(len(entries), 1), cap):
                part = list(zip(entries[i:i + cap], offs[i:i + cap]))
                body = b"SNOD" + struct.pack("<BBH", 1, 0, len(part))
                for (name, oaddr, ctype, scratch), noff in part:
                    body += struct.pack("<QQII", noff, oaddr, ctype, 0) + scratch
                body += b"\0" * (40 * (cap - len(part)))
                nodes.append(alloc(body))
                keys.append(part[-1][1] if part else 0)
            if len(nodes) > 2 * self.INTERNAL_K:
                raise HDF5Error("too many group members for this writer")
            bt = b"TREE" + struct.pack("<BBHQQ", 0, 0, len(nodes), UNDEF, UNDEF)
            for i, n in enumerate(nodes):
                bt += struct.pack("<QQ", keys[i], n)
            bt += struct.pack("<Q", keys[len(nodes)])
            bt += b"\0" * (24 + (2 * self.INTERNAL_K + 1) * 8 + 2 * self.INTERNAL_K * 8 - len(bt))
            btree_addr = alloc(bt)
            msgs = [_message(0x11, struct.pack("<QQ", btree_addr, heap_addr))]
            msgs += [_attribute_message(k, v) for k, v in node["attrs"].items()]
            return alloc(_object_header(msgs)), btree_addr, heap_addr

        root, btree, heap = write_group(self.tree)
        while len(out) % 8:
            out.append(0)
        sb = SIGNATURE + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, self.LEAF_K, self.INTERNAL_K, 0)
        sb += struct.pack("<QQQQ", 0, UNDEF, len(out), UNDEF)
        sb += struct.pack("<QQII", 0, root, 1, 0) + struct.pack("<QQ", btree, heap)
        assert len(sb) == 96
        out[0:96] = sb
        with open(path, "wb") as fh:
            fh.write(bytes(out))


# ===================================================================================================== DOLFIN layout

_CELLTYPES = {2: "interval", 3: "triangle", 4: "tetrahedron"}
DG0_SIGNATURE = {
    "tetrahedron": "FiniteElement('Discontinuous Lagrange', tetrahedron, 0)",
    "triangle": "FiniteElement('Discontinuous Lagrange', triangle, 0)",
    "interval": "FiniteElement('Discontinuous Lagrange', interval, 0)",
}


def write_dolfin_h5(path, xyz, cells, fields):
    """The container PreprocessingMultiCompt.py:148-152 writes: mesh + one DG0 function per entry of `fields`
    (name -> one value per cell)."""
    xyz = np.asarray(xyz, dtype="<f8")
    cells = np.asarray(cells)
    nc, nvc = cells.shape
    celltype = _CELLTYPES[nvc]
    w = Writer()
    w.dataset("mesh/coordinates", xyz)
    w.dataset("mesh/topology", cells.astype("<i8"),
              attrs={"celltype": celltype, "partition": np.array([0], dtype="<u8")})
    ident = np.arange(nc, dtype="<i8")
    for name, values in fields.items():
        v = np.asarray(values, dtype="<f8").reshape(nc)
        w.dataset(name + "/vector_0", v, attrs={"partition": np.array([0], dtype="<u8")})
        w.dataset(name + "/cell_dofs", ident)
        w.dataset(name + "/x_cell_dofs", np.arange(nc + 1, dtype="<u8"))
        w.dataset(name + "/cells", ident.astype("<u8"))
        w.group_attrs(name, signature=DG0_SIGNATURE[celltype])
    w.save(path)


def read_dolfin_h5(path):
    """-> {"xyz", "tets", <DG0 fields by name: T2, ic, phase, d00..d22 ...>} (what GCloudDmriSolver.py:150-177 reads)."""
    f = File(path)
    if "mesh/topology" not in f or "mesh/coordinates" not in f:
        raise HDF5Error("no /mesh/topology + /mesh/coordinates in " + path)
    xyz = np.asarray(f["mesh/coordinates"], dtype=float)
    cells = np.asarray(f["mesh/topology"]).astype(np.int32)
    out = {"xyz": xyz, "tets": cells}
    nc = len(cells)
    for name in f.keys(""):
        if name == "mesh":
            continue
        members = f.keys(name)
        vec = next((m for m in members if m.startswith("vector_")), None)
        if vec is None or "cell_dofs" not in members:
            continue
        values = np.asarray(f[name + "/" + vec], dtype=float)
        cell_dofs = np.asarray(f[name + "/cell_dofs"]).astype(np.int64)
        x = np.asarray(f[name + "/x_cell_dofs"]).astype(np.int64)
        cidx = np.asarray(f[name + "/cells"]).astype(np.int64)
        if len(x) != nc + 1 or not np.all(np.diff(x) == 1):
            continue      # not a DG0 function on this mesh
        field = np.empty(nc)
        field[cidx] = values[cell_dofs[x[:-1]]]
        out[name] = field
    return out
