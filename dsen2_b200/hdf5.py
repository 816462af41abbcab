"""Minimal pure-Python HDF5 reader/writer for the two file families on the DSen2 path.

The reference reads its fixtures and weights through h5py
(``testing/demoDSen2.py:14-28`` ``readh5``; Keras ``model.load_weights`` at
``testing/supres.py:63``).  h5py / libhdf5 are not part of this image, so this
module restates the subset of the HDF5 1.8 "classic" file format those files
use:

* MATLAB v7.3 ``.mat`` scenes: 512-byte user block, superblock v0, v1 object
  headers (+ continuation blocks), symbol-table groups (v1 B-tree + local
  heap), **chunked** datasets (layout v3, v1 chunk B-tree) with the deflate
  filter.
* Keras 2.x weight / full-model files: nested groups, **contiguous** (or
  compact) datasets, v1..v3 attribute messages holding fixed-length (or
  variable-length, via the global heap) string arrays.

`File` is read-only; `write_hdf5` emits the same classic layout (superblock v0,
symbol-table groups, contiguous datasets, v1 attributes) so that weights saved
by this package can be opened by h5py / Keras and so that tests can build
Keras-style fixtures without h5py.
"""
from __future__ import annotations

import struct
import zlib
from typing import Dict, Iterator, List, Optional, Tuple, Union

import numpy as np

_SIG = b"\x89HDF\r\n\x1a\n"
_UNDEF = 0xFFFFFFFFFFFFFFFF


class HDF5Error(OSError):
    """Raised for files this reader cannot parse (h5py raises OSError too)."""


# --------------------------------------------------------------------------- #
# reader
# --------------------------------------------------------------------------- #
class _Datatype:
    __slots__ = ("cls", "size", "dtype", "vlen_str", "base")

    def __init__(self, cls, size, dtype, vlen_str=False, base=None):
        self.cls, self.size, self.dtype, self.vlen_str, self.base = cls, size, dtype, vlen_str, base


def _parse_datatype(buf: bytes, off: int = 0) -> _Datatype:
    b0 = buf[off]
    cls, _ver = b0 & 0x0F, b0 >> 4
    bits = buf[off + 1] | (buf[off + 2] << 8) | (buf[off + 3] << 16)
    size = struct.unpack_from("<I", buf, off + 4)[0]
    if cls == 0:  # fixed point
        endian = ">" if bits & 1 else "<"
        signed = bool(bits & 0x08)
        return _Datatype(cls, size, np.dtype("%s%s%d" % (endian, "i" if signed else "u", size)))
    if cls == 1:  # floating point
        endian = ">" if bits & 1 else "<"
        return _Datatype(cls, size, np.dtype("%sf%d" % (endian, size)))
    if cls == 3:  # fixed-length string
        return _Datatype(cls, size, np.dtype("S%d" % size))
    if cls == 9:  # variable length
        vtype = bits & 0x0F
        base = _parse_datatype(buf, off + 8)
        return _Datatype(cls, size, np.dtype("O"), vlen_str=(vtype == 1), base=base)
    if cls == 7:  # reference (MATLAB cell arrays) -- returned as raw addresses
        return _Datatype(cls, size, np.dtype("<u%d" % size))
    if cls == 8:  # enum (MATLAB logical): base type follows
        base = _parse_datatype(buf, off + 8)
        return _Datatype(cls, size, base.dtype)
    raise HDF5Error("unsupported HDF5 datatype class %d" % cls)


def _parse_dataspace(buf: bytes) -> Tuple[int, ...]:
    ver, rank, flags = buf[0], buf[1], buf[2]
    if ver == 1:
        off = 8
    elif ver == 2:
        off = 4
        if buf[3] == 2:  # null dataspace
            return (0,)
    else:
        raise HDF5Error("unsupported dataspace version %d" % ver)
    return tuple(struct.unpack_from("<%dQ" % rank, buf, off)) if rank else ()


class _Object:
    """Parsed object header: list of (type, flags, payload) messages."""

    def __init__(self, f: "File", addr: int):
        self.f, self.addr = f, addr
        self.msgs: List[Tuple[int, int, bytes]] = []
        self._read_header(addr)

    def _read_header(self, addr: int) -> None:
        f = self.f
        head = f._read(addr, 16)
        if head[:4] == b"OHDR":
            self._read_header_v2(addr)
            return
        ver, _, nmsg, _refc, hsize = struct.unpack_from("<BBHII", head, 0)
        if ver != 1:
            raise HDF5Error("unsupported object header version %d" % ver)
        blocks = [(addr + 16, hsize)]
        while blocks and len(self.msgs) < nmsg:
            boff, blen = blocks.pop(0)
            data = f._read(boff, blen)
            p = 0
            while p + 8 <= blen and len(self.msgs) < nmsg:
                mtype, msize, mflags = struct.unpack_from("<HHB", data, p)
                payload = data[p + 8:p + 8 + msize]
                p += 8 + msize
                if mtype == 0x10:
                    coff, clen = struct.unpack_from("<QQ", payload, 0)
                    blocks.append((f._base + coff, clen))
                self.msgs.append((mtype, mflags, payload))

    def _read_header_v2(self, addr: int) -> None:
        f = self.f
        head = f._read(addr, 64)
        flags = head[5]
        p = 6
        if flags & 0x20:
            p += 16
        if flags & 0x10:
            p += 4
        szbytes = 1 << (flags & 3)
        chunk0 = int.from_bytes(head[p:p + szbytes], "little")
        p += szbytes
        track = bool(flags & 0x04)
        blocks = [(addr + p, chunk0)]
        while blocks:
            boff, blen = blocks.pop(0)
            data = f._read(boff, blen)
            q = 0
            while q + 4 <= blen:
                mtype, msize, mflags = data[q], struct.unpack_from("<H", data, q + 1)[0], data[q + 3]
                q += 4 + (2 if track else 0)
                payload = data[q:q + msize]
                q += msize
                if mtype == 0x10:
                    coff, clen = struct.unpack_from("<QQ", payload, 0)
                    # continuation chunks start with "OCHK" and end with a checksum
                    blocks.append((f._base + coff + 4, clen - 8))
                self.msgs.append((mtype, mflags, payload))

    def find(self, mtype: int) -> Optional[bytes]:
        for t, _fl, p in self.msgs:
            if t == mtype:
                return p
        return None

    def find_all(self, mtype: int) -> List[bytes]:
        return [p for t, _fl, p in self.msgs if t == mtype]


class AttributeManager:
    def __init__(self, obj: _Object):
        self._obj = obj
        self._cache: Optional[Dict[str, object]] = None

    def _load(self) -> Dict[str, object]:
        if self._cache is None:
            out: Dict[str, object] = {}
            for payload in self._obj.find_all(0x0C):
                name, val = self._obj.f._parse_attribute(payload)
                out[name] = val
            self._cache = out
        return self._cache

    def __getitem__(self, k):
        return self._load()[k]

    def __contains__(self, k):
        return k in self._load()

    def keys(self):
        return self._load().keys()

    def get(self, k, default=None):
        return self._load().get(k, default)


class Dataset:
    def __init__(self, f: "File", obj: _Object, name: str):
        self._f, self._obj, self.name = f, obj, name
        self._dt = _parse_datatype(obj.find(0x03))
        self.shape = _parse_dataspace(obj.find(0x01))
        self.dtype = self._dt.dtype
        self.attrs = AttributeManager(obj)

    def __getitem__(self, key):
        if key is not Ellipsis and key != ():
            return self._read()[key]
        return self._read()

    def __array__(self, dtype=None, copy=None):
        a = self._read()
        return a.astype(dtype) if dtype is not None else a

    def _read(self) -> np.ndarray:
        f = self._f
        lay = self._obj.find(0x08)
        ver = lay[0]
        if ver != 3:
            raise HDF5Error("unsupported data layout version %d" % ver)
        cls = lay[1]
        n = int(np.prod(self.shape)) if self.shape else 1
        esize = self._dt.size
        if cls == 0:  # compact
            size = struct.unpack_from("<H", lay, 2)[0]
            raw = lay[4:4 + size]
        elif cls == 1:  # contiguous
            addr, size = struct.unpack_from("<QQ", lay, 2)
            raw = b"\x00" * (n * esize) if addr == _UNDEF else f._read(f._base + addr, n * esize)
        elif cls == 2:  # chunked
            return self._read_chunked(lay)
        else:
            raise HDF5Error("unsupported layout class %d" % cls)
        return f._decode(raw, self._dt, self.shape)

    def _filters(self) -> List[int]:
        p = self._obj.find(0x0B)
        if p is None:
            return []
        ver, nf = p[0], p[1]
        ids = []
        q = 8 if ver == 1 else 2
        for _ in range(nf):
            fid, = struct.unpack_from("<H", p, q)
            if ver == 1 or fid >= 256:
                nlen, = struct.unpack_from("<H", p, q + 2)
                q += 4
            else:
                nlen = 0
                q += 2
            _fl, ncv = struct.unpack_from("<HH", p, q)
            q += 4
            q += (nlen + 7) // 8 * 8 if ver == 1 else nlen
            q += 4 * ncv
            if ver == 1 and ncv % 2:
                q += 4
            ids.append(fid)
        return ids

    def _read_chunked(self, lay: bytes) -> np.ndarray:
        f = self._f
        rank = lay[2]  # dataset rank + 1
        btree, = struct.unpack_from("<Q", lay, 3)
        cdims = struct.unpack_from("<%dI" % rank, lay, 11)
        chunk_shape = cdims[:-1]
        filters = self._filters()
        for fid in filters:
            if fid not in (1, 2):  # deflate, shuffle
                raise HDF5Error("unsupported HDF5 filter id %d" % fid)
        out = np.zeros(self.shape, dtype=self._dt.dtype.newbyteorder("="))
        if btree == _UNDEF:
            return out
        esize = self._dt.size
        for offs, csize, mask, caddr in f._iter_chunks(f._base + btree, rank):
            raw = f._read(f._base + caddr, csize)
            for i, fid in reversed(list(enumerate(filters))):
                if mask & (1 << i):
                    continue
                if fid == 1:
                    raw = zlib.decompress(raw)
                elif fid == 2:
                    a = np.frombuffer(raw, np.uint8).reshape(esize, -1)
                    raw = a.T.tobytes()
            chunk = np.frombuffer(raw, self._dt.dtype, count=int(np.prod(chunk_shape))).reshape(chunk_shape)
            sl_out, sl_in = [], []
            for o, c, s in zip(offs[:-1], chunk_shape, self.shape):
                e = min(o + c, s)
                sl_out.append(slice(o, e))
                sl_in.append(slice(0, e - o))
            out[tuple(sl_out)] = chunk[tuple(sl_in)]
        return out


class Group:
    def __init__(self, f: "File", obj: _Object, name: str):
        self._f, self._obj, self.name = f, obj, name
        self.attrs = AttributeManager(obj)
        self._links: Optional[Dict[str, int]] = None

    def _load(self) -> Dict[str, int]:
        if self._links is None:
            links: Dict[str, int] = {}
            st = self._obj.find(0x11)
            if st is not None:
                btree, heap = struct.unpack_from("<QQ", st, 0)
                links.update(self._f._read_symtab(self._f._base + btree, self._f._base + heap))
            for p in self._obj.find_all(0x06):  # new-style compact links
                nm, addr = self._f._parse_link(p)
                if addr is not None:
                    links[nm] = addr
            self._links = links
        return self._links

    def keys(self):
        return list(self._load().keys())

    def __iter__(self) -> Iterator[str]:
        return iter(self.keys())

    def __len__(self):
        return len(self._load())

    def __contains__(self, k):
        try:
            self[k]
            return True
        except KeyError:
            return False

    def __getitem__(self, path: str) -> Union["Group", Dataset]:
        node: Union[Group, Dataset] = self._f if path.startswith("/") else self
        for part in [p for p in path.split("/") if p]:
            if not isinstance(node, Group):
                raise KeyError(path)
            links = node._load()
            if part not in links:
                raise KeyError("%s (no member %r)" % (path, part))
            node = node._f._open(links[part], (node.name.rstrip("/") + "/" + part))
        return node

    def items(self):
        return [(k, self[k]) for k in self.keys()]

    def visit_datasets(self, prefix: str = "") -> Iterator[Tuple[str, Dataset]]:
        for k in self.keys():
            node = self[k]
            if isinstance(node, Group):
                yield from node.visit_datasets(prefix + k + "/")
            else:
                yield prefix + k, node


class File(Group):
    """Read-only HDF5 file (subset). ``File(path)['im10'][()]`` like h5py."""

    def __init__(self, path: str, mode: str = "r"):
        if mode != "r":
            raise ValueError("dsen2_b200.hdf5.File is read-only; use write_hdf5")
        with open(path, "rb") as fh:  # raises FileNotFoundError (an OSError) like h5py
            self._buf = fh.read()
        self.filename = path
        sb = -1
        off = 0
        while off + 8 <= len(self._buf):
            if self._buf[off:off + 8] == _SIG:
                sb = off
                break
            off = 512 if off == 0 else off * 2
        if sb < 0:
            raise HDF5Error("Unable to open file (file signature not found): %s" % path)
        ver = self._buf[sb + 8]
        if ver in (0, 1):
            so, sl = self._buf[sb + 13], self._buf[sb + 14]
            if (so, sl) != (8, 8):
                raise HDF5Error("only 8-byte offsets/lengths are supported")
            p = sb + 24 + (4 if ver == 1 else 0)
            base, _fs, _eof, _drv = struct.unpack_from("<QQQQ", self._buf, p)
            self._base = base
            _lno, root = struct.unpack_from("<QQ", self._buf, p + 32)
            root_addr = base + root
        elif ver in (2, 3):
            base, _ext, _eof, root = struct.unpack_from("<QQQQ", self._buf, sb + 12)
            self._base = base
            root_addr = base + root
        else:
            raise HDF5Error("unsupported superblock version %d" % ver)
        self._gheaps: Dict[int, Dict[int, bytes]] = {}
        Group.__init__(self, self, _Object(self, root_addr), "/")

    # context manager like h5py
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

    def close(self):
        pass

    # -- low-level helpers ------------------------------------------------ #
    def _read(self, off: int, n: int) -> bytes:
        if off < 0 or off + n > len(self._buf):
            raise HDF5Error("truncated HDF5 file (read %d bytes at %d)" % (n, off))
        return self._buf[off:off + n]

    def _open(self, addr: int, name: str) -> Union[Group, Dataset]:
        obj = _Object(self, self._base + addr)
        if obj.find(0x08) is not None and obj.find(0x03) is not None:
            return Dataset(self, obj, name)
        return Group(self, obj, name)

    def _heap_str(self, heap_addr: int, off: int) -> str:
        h = self._read(heap_addr, 32)
        if h[:4] != b"HEAP":
            raise HDF5Error("bad local heap signature")
        dseg, = struct.unpack_from("<Q", h, 24)
        start = self._base + dseg + off
        end = self._buf.index(b"\x00", start)
        return self._buf[start:end].decode("utf-8")

    def _read_symtab(self, btree: int, heap: int) -> Dict[str, int]:
        out: Dict[str, int] = {}
        node = self._read(btree, 24)
        if node[:4] == b"SNOD":
            nsym, = struct.unpack_from("<H", node, 6)
            ents = self._read(btree + 8, nsym * 40)
            for i in range(nsym):
                lno, oaddr = struct.unpack_from("<QQ", ents, i * 40)
                out[self._heap_str(heap, lno)] = oaddr
            return out
        if node[:4] != b"TREE":
            raise HDF5Error("bad group B-tree signature")
        ntype, level, nent = struct.unpack_from("<BBH", node, 4)
        if ntype != 0:
            raise HDF5Error("expected a group B-tree node")
        body = self._read(btree + 24, (2 * nent + 1) * 8)
        for i in range(nent):
            child, = struct.unpack_from("<Q", body, 8 + i * 16)
            out.update(self._read_symtab(self._base + child, heap))
        return out

    def _iter_chunks(self, addr: int, rank: int):
        node = self._read(addr, 24)
        if node[:4] != b"TREE":
            raise HDF5Error("bad chunk B-tree signature")
        ntype, level, nent = struct.unpack_from("<BBH", node, 4)
        if ntype != 1:
            raise HDF5Error("expected a chunk B-tree node")
        ksize = 8 + 8 * rank
        body = self._read(addr + 24, nent * (ksize + 8) + ksize)
        for i in range(nent):
            k = i * (ksize + 8)
            csize, mask = struct.unpack_from("<II", body, k)
            offs = struct.unpack_from("<%dQ" % rank, body, k + 8)
            child, = struct.unpack_from("<Q", body, k + ksize)
            if level == 0:
                yield offs, csize, mask, child
            else:
                yield from self._iter_chunks(self._base + child, rank)

    def _parse_link(self, p: bytes):
        ver, flags = p[0], p[1]
        q = 2
        ltype = 0
        if flags & 0x08:
            ltype = p[q]
            q += 1
        if flags & 0x04:
            q += 8
        if flags & 0x10:
            q += 1
        lsz = 1 << (flags & 3)
        nlen = int.from_bytes(p[q:q + lsz], "little")
        q += lsz
        name = p[q:q + nlen].decode("utf-8")
        q += nlen
        if ltype != 0:
            return name, None
        addr, = struct.unpack_from("<Q", p, q)
        return name, addr

    def _gheap_obj(self, addr: int, idx: int) -> bytes:
        if addr not in self._gheaps:
            head = self._read(self._base + addr, 16)
            if head[:4] != b"GCOL":
                raise HDF5Error("bad global heap signature")
            size, = struct.unpack_from("<Q", head, 8)
            data = self._read(self._base + addr, size)
            objs: Dict[int, bytes] = {}
            p = 16
            while p + 16 <= size:
                oidx, _rc, _r, osz = struct.unpack_from("<HHIQ", data, p)
                if oidx == 0:
                    break
                objs[oidx] = data[p + 16:p + 16 + osz]
                p += 16 + (osz + 7) // 8 * 8
            self._gheaps[addr] = objs
        return self._gheaps[addr][idx]

    def _decode(self, raw: bytes, dt: _Datatype, shape: Tuple[int, ...]):
        n = int(np.prod(shape)) if shape else 1
        if dt.cls == 9:
            vals = []
            for i in range(n):
                ln, addr, idx = struct.unpack_from("<IQI", raw, i * 16)
                if ln == 0 or addr == 0:
                    b = b""
                else:
                    b = self._gheap_obj(addr, idx)[:ln * dt.base.size]
                vals.append(b.decode("utf-8") if dt.vlen_str else np.frombuffer(b, dt.base.dtype))
            if not shape:
                return vals[0]
            out = np.empty(n, dtype=object)
            out[:] = vals
            return out.reshape(shape)
        arr = np.frombuffer(raw, dt.dtype, count=n)
        if dt.dtype.kind != "S" and dt.dtype.byteorder == ">":
            arr = arr.astype(dt.dtype.newbyteorder("="))
        arr = arr.reshape(shape).copy()
        if not shape:
            return arr[()]
        return arr

    def _parse_attribute(self, p: bytes):
        ver = p[0]
        if ver == 1:
            nsz, dsz, ssz = struct.unpack_from("<HHH", p, 2)
            q = 8
            pad = lambda x: (x + 7) // 8 * 8
            name = p[q:q + nsz].split(b"\x00")[0].decode("utf-8")
            q += pad(nsz)
            dt = _parse_datatype(p, q)
            q += pad(dsz)
            shape = _parse_dataspace(p[q:q + ssz])
            q += pad(ssz)
        elif ver in (2, 3):
            nsz, dsz, ssz = struct.unpack_from("<HHH", p, 2)
            q = 8 + (1 if ver == 3 else 0)
            name = p[q:q + nsz].split(b"\x00")[0].decode("utf-8")
            q += nsz
            dt = _parse_datatype(p, q)
            q += dsz
            shape = _parse_dataspace(p[q:q + ssz])
            q += ssz
        else:
            raise HDF5Error("unsupported attribute message version %d" % ver)
        return name, self._decode(p[q:], dt, shape)


# --------------------------------------------------------------------------- #
# writer (classic layout; enough for Keras-style weight files and .mat-like scenes)
# --------------------------------------------------------------------------- #
def _dt_message(a: np.ndarray) -> bytes:
    dt = a.dtype
    if dt.kind == "f":
        if dt.itemsize == 4:
            props = struct.pack("<HHBBBBI", 0, 32, 23, 8, 0, 23, 127)
            bits = (0x20, 31, 0)
        elif dt.itemsize == 8:
            props = struct.pack("<HHBBBBI", 0, 64, 52, 11, 0, 52, 1023)
            bits = (0x20, 63, 0)
        elif dt.itemsize == 2:
            props = struct.pack("<HHBBBBI", 0, 16, 10, 5, 0, 10, 15)
            bits = (0x20, 15, 0)
        else:
            raise ValueError("unsupported float size")
        return struct.pack("<BBBBI", 0x11, bits[0], bits[1], bits[2], dt.itemsize) + props
    if dt.kind in "iu":
        b0 = 0x08 if dt.kind == "i" else 0
        return struct.pack("<BBBBI", 0x10, b0, 0, 0, dt.itemsize) + struct.pack("<HH", 0, dt.itemsize * 8)
    if dt.kind == "S":
        return struct.pack("<BBBBI", 0x13, 0x00, 0, 0, dt.itemsize)  # null-terminated ASCII
    raise ValueError("unsupported dtype for HDF5 write: %r" % dt)


def _ds_message(shape) -> bytes:
    rank = len(shape)
    return struct.pack("<BBBB4x", 1, rank, 0, 0) + b"".join(struct.pack("<Q", s) for s in shape)


def _pad8(b: bytes) -> bytes:
    return b + b"\x00" * (-len(b) % 8)


def _msg(mtype: int, payload: bytes, flags: int = 0) -> bytes:
    payload = _pad8(payload)
    return struct.pack("<HHB3x", mtype, len(payload), flags) + payload


def _attr_message(name: str, value) -> bytes:
    a = np.asarray(value)
    if a.dtype.kind == "U":
        a = np.char.encode(a, "utf-8")
    if a.dtype.kind == "S":
        # h5py stores numpy bytes arrays as fixed-length strings
        a = a.astype("S%d" % max(1, a.dtype.itemsize))
    nm = name.encode("utf-8") + b"\x00"
    dtm, dsm = _dt_message(a), _ds_message(a.shape)
    body = struct.pack("<BxHHH", 1, len(nm), len(dtm), len(dsm)) + _pad8(nm) + _pad8(dtm) + _pad8(dsm)
    body += np.ascontiguousarray(a).tobytes()
    return _msg(0x0C, body)


class _Writer:
    def __init__(self):
        self.buf = bytearray()

    def alloc(self, data: bytes, align: int = 8) -> int:
        self.buf += b"\x00" * (-len(self.buf) % align)
        addr = len(self.buf)
        self.buf += data
        return addr

    def object_header(self, msgs: List[bytes]) -> int:
        body = b"".join(msgs)
        head = struct.pack("<BBHII4x", 1, 0, len(msgs), 1, len(body))
        return self.alloc(head + body)

    def dataset(self, arr: np.ndarray, attrs: Dict[str, object]) -> int:
        a = np.ascontiguousarray(arr)
        if a.dtype.kind == "f" and a.dtype.byteorder == ">":
            a = a.astype(a.dtype.newbyteorder("<"))
        data_addr = self.alloc(a.tobytes()) if a.size else _UNDEF
        msgs = [
            _msg(0x01, _ds_message(a.shape)),
            _msg(0x03, _dt_message(a), flags=1),
            _msg(0x05, struct.pack("<BBBB", 2, 2, 0, 0)),  # fill value v2: alloc late, undefined
            _msg(0x08, struct.pack("<BBQQ", 3, 1, data_addr, a.nbytes)),
        ]
        msgs += [_attr_message(k, v) for k, v in attrs.items()]
        return self.object_header(msgs)

    def group(self, members: Dict[str, int], attrs: Dict[str, object]) -> int:
        names = sorted(members.keys(), key=lambda s: s.encode("utf-8"))
        # local heap: offset 0 is the empty string
        heap = bytearray(b"\x00" * 8)
        offs = {}
        for nme in names:
            offs[nme] = len(heap)
            heap += _pad8(nme.encode("utf-8") + b"\x00")
        heap_data = bytes(heap) + b"\x00" * 16  # room for a free block
        free_off = len(heap)
        heap_data = bytearray(heap_data)
        struct.pack_into("<QQ", heap_data, free_off, 1, 16)  # free block: next=1 (none), size
        dseg = self.alloc(bytes(heap_data))
        heap_addr = self.alloc(b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap_data), free_off, dseg))
        # symbol nodes: split into leaves of at most 2K (K=4 -> 8) entries
        leaf_k = 4
        leaves: List[Tuple[int, str]] = []  # (address, last name)
        chunk = 2 * leaf_k
        groups_of = [names[i:i + chunk] for i in range(0, len(names), chunk)] or [[]]
        for part in groups_of:
            ents = b""
            for nme in part:
                ents += struct.pack("<QQII16x", offs[nme], members[nme], 0, 0)
            ents += b"\x00" * (40 * (chunk - len(part)))
            addr = self.alloc(b"SNOD" + struct.pack("<BBH", 1, 0, len(part)) + ents)
            leaves.append((addr, part[-1] if part else ""))
        # single-level B-tree (internal K=16 -> up to 32 children); nest if more
        def build(level_nodes: List[Tuple[int, str]], level: int) -> int:
            maxc = 32
            if len(level_nodes) <= maxc:
                body = struct.pack("<Q", 0)
                for addr, last in level_nodes:
                    body += struct.pack("<QQ", addr, offs.get(last, 0))
                body += b"\x00" * (16 * (maxc - len(level_nodes)))
                node = b"TREE" + struct.pack("<BBHQQ", 0, level, len(level_nodes), _UNDEF, _UNDEF) + body
                return self.alloc(node)
            parents = []
            for i in range(0, len(level_nodes), maxc):
                sub = level_nodes[i:i + maxc]
                parents.append((build(sub, level), sub[-1][1]))
            return build(parents, level + 1)

        btree = build(leaves, 0)
        msgs = [_msg(0x11, struct.pack("<QQ", btree, heap_addr))]
        msgs += [_attr_message(k, v) for k, v in attrs.items()]
        return self.object_header(msgs)


def write_hdf5(path: str, tree: Dict[str, object], attrs: Optional[Dict[str, Dict[str, object]]] = None) -> None:
    """Write ``tree`` (nested dicts of numpy arrays) as a classic-format HDF5 file.

    ``attrs`` maps an absolute group/dataset path (``"/"``, ``"/model_weights"``,
    ``"/model_weights/conv2d_1"`` ...) to a dict of attributes.
    """
    attrs = attrs or {}
    w = _Writer()
    w.buf += b"\x00" * 96  # superblock v0 with 8-byte offsets is 96 bytes

    def emit(node, path_: str) -> int:
        a = attrs.get(path_ or "/", {})
        if isinstance(node, dict):
            members = {k: emit(v, path_ + "/" + k) for k, v in node.items()}
            return w.group(members, a)
        return w.dataset(np.asarray(node), a)

    root = emit(tree, "")
    eof = len(w.buf)
    sb = _SIG + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, 4, 16, 0)
    sb += struct.pack("<QQQQ", 0, _UNDEF, eof, _UNDEF)
    sb += struct.pack("<QQII16x", 0, root, 0, 0)
    assert len(sb) == 96, len(sb)
    w.buf[:96] = sb
    with open(path, "wb") as fh:
        fh.write(bytes(w.buf))
