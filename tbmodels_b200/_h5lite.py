"""Minimal pure-Python HDF5 reader / writer for the files on the `tbmodels eigenvals` path.

The reference CLI (src/tbmodels/_cli.py:227-262) reads a model (`Model.from_hdf5`, src/tbmodels/_tb_model.py:1009-1041)
and a k-point list (`bands_inspect.io.load`) and writes `bands_inspect` `EigenvalsData`, all through h5py, which is not
available where this package runs.  This module implements exactly the subset of the HDF5 file format those files use
(as written by h5py / libhdf5 with default settings):

  reader: superblock v0/v1, v1 object headers (+ continuation blocks), symbol-table groups (v1 B-tree, local heap,
          SNOD), dataspace v1/v2, datatypes fixed-point / IEEE float / fixed and variable-length string / compound
          (complex = {r, i}) / enum (h5py bool), layouts compact / contiguous / chunked without filters (v1 chunk
          B-tree), global heap collections for variable-length strings.
  writer: superblock v0, symbol-table groups, contiguous datasets of float64 / int64 / complex128 / bool and
          variable-length UTF-8 scalar strings -- the same structures h5py emits, so that h5py-based tools (the
          reference, bands_inspect) can read the result.

It is host-side file plumbing for SURVEY.md section 8 row f1; no numerical work happens here.
"""
from __future__ import annotations

import struct
from typing import Any, Dict, Optional

import numpy as np

_SIG = b"\x89HDF\r\n\x1a\n"
_UNDEF = 0xFFFFFFFFFFFFFFFF


class H5Error(ValueError):
    """Raised for files outside the supported subset (or corrupt files)."""


# ======================================================================================================= reader
class _Reader:
    def __init__(self, data: bytes):
        self.b = data
        if data[:8] != _SIG:
            raise H5Error("not an HDF5 file (bad signature)")
        ver = data[8]
        if ver not in (0, 1):
            raise H5Error(f"superblock version {ver} is not supported (only the v0/v1 layout h5py writes by default)")
        self.O = data[13]
        self.L = data[14]
        if self.O != 8 or self.L != 8:
            raise H5Error("only 8-byte offsets / lengths are supported")
        p = 24 + (4 if ver == 1 else 0)
        self.base, _free, self.eof, _drv = struct.unpack_from("<QQQQ", data, p)
        p += 32
        # root group symbol table entry
        _name_off, self.root_header = struct.unpack_from("<QQ", data, p)

    # ---- object headers
    def messages(self, addr: int):
        b = self.b
        addr += self.base
        ver = b[addr]
        if ver != 1:
            raise H5Error(f"object header version {ver} is not supported")
        nmsg, = struct.unpack_from("<H", b, addr + 2)
        hsize, = struct.unpack_from("<I", b, addr + 8)
        blocks = [(addr + 16, hsize)]
        out = []
        while blocks and len(out) < nmsg:
            p, size = blocks.pop(0)
            end = p + size
            while p + 8 <= end and len(out) < nmsg:
                mtype, msize, _flags = struct.unpack_from("<HHB", b, p)
                body = p + 8
                if mtype == 0x0010:  # continuation
                    coff, clen = struct.unpack_from("<QQ", b, body)
                    blocks.append((coff + self.base, clen))
                out.append((mtype, body, msize))
                p = body + msize
        return out

    # ---- groups
    def _heap_data(self, heap_addr: int) -> int:
        b = self.b
        heap_addr += self.base
        if b[heap_addr:heap_addr + 4] != b"HEAP":
            raise H5Error("bad local heap signature")
        return struct.unpack_from("<Q", b, heap_addr + 24)[0] + self.base

    def _btree_group(self, addr: int, heap_data: int, out: Dict[str, int]):
        b = self.b
        addr += self.base
        if b[addr:addr + 4] != b"TREE":
            raise H5Error("bad B-tree signature")
        ntype, level, used = struct.unpack_from("<BBH", b, addr + 4)
        if ntype != 0:
            raise H5Error("expected a group B-tree node")
        p = addr + 24
        for i in range(used):
            child, = struct.unpack_from("<Q", b, p + 8 + 16 * i)
            if level > 0:
                self._btree_group(child, heap_data, out)
            else:
                self._snod(child, heap_data, out)

    def _snod(self, addr: int, heap_data: int, out: Dict[str, int]):
        b = self.b
        addr += self.base
        if b[addr:addr + 4] != b"SNOD":
            raise H5Error("bad symbol node signature")
        n, = struct.unpack_from("<H", b, addr + 6)
        p = addr + 8
        for _ in range(n):
            name_off, header = struct.unpack_from("<QQ", b, p)
            q = heap_data + name_off
            name = b[q:b.index(b"\0", q)].decode("utf-8")
            out[name] = header
            p += 40

    def links(self, msgs) -> Optional[Dict[str, int]]:
        for mtype, body, _ in msgs:
            if mtype == 0x0011:
                btree, heap = struct.unpack_from("<QQ", self.b, body)
                out: Dict[str, int] = {}
                self._btree_group(btree, self._heap_data(heap), out)
                return out
            if mtype in (0x0002, 0x0006):
                raise H5Error("new-style (link message) groups are not supported")
        return None

    # ---- datatypes
    def datatype(self, p: int):
        """-> (descriptor, bytes consumed).  descriptor: ('num', np.dtype) | ('str', size) | ('vstr',) | ('vlen', base)"""
        b = self.b
        cv = b[p]
        cls, ver = cv & 0x0F, cv >> 4
        bits = b[p + 1] | (b[p + 2] << 8) | (b[p + 3] << 16)
        size, = struct.unpack_from("<I", b, p + 4)
        q = p + 8
        order = ">" if (bits & 1) else "<"
        if cls == 0:
            signed = bool(bits & 0x08)
            return ("num", np.dtype(f"{order}{'i' if signed else 'u'}{size}")), 8 + 4
        if cls == 1:
            if size not in (2, 4, 8):
                raise H5Error(f"float size {size} unsupported")
            return ("num", np.dtype(f"{order}f{size}")), 8 + 12
        if cls == 3:
            return ("str", size), 8
        if cls == 6:
            nmemb = bits & 0xFFFF
            fields = []
            for _ in range(nmemb):
                end = b.index(b"\0", q)
                name = b[q:end].decode("utf-8")
                if ver < 3:
                    q += ((end - q) // 8 + 1) * 8
                    off, = struct.unpack_from("<I", b, q)
                    q += 4
                    if ver == 1:
                        q += 1 + 3 + 4 + 4 + 16
                else:
                    q = end + 1
                    nb = 1 if size < 256 else 2 if size < 65536 else 3 if size < 16777216 else 4
                    off = int.from_bytes(b[q:q + nb], "little")
                    q += nb
                (kind, *rest), used = self.datatype(q)
                if kind != "num":
                    raise H5Error("compound members must be numeric")
                q += used
                fields.append((name, rest[0], off))
            names = [f[0] for f in fields]
            if names == ["r", "i"] and fields[0][1] == fields[1][1] and fields[0][1].kind == "f":
                fs = fields[0][1].itemsize
                if fields[0][2] == 0 and fields[1][2] == fs and size == 2 * fs:
                    return ("num", np.dtype(f"{fields[0][1].byteorder.replace('=', '<').replace('|', '<')}c{2 * fs}")), q - p
            dt = np.dtype({"names": names, "formats": [f[1] for f in fields], "offsets": [f[2] for f in fields],
                           "itemsize": size})
            return ("num", dt), q - p
        if cls == 8:
            nmemb = bits & 0xFFFF
            (kind, base), used = self.datatype(q)
            q += used
            names = []
            for _ in range(nmemb):
                end = b.index(b"\0", q)
                names.append(b[q:end].decode("utf-8"))
                q = (q + ((end - q) // 8 + 1) * 8) if ver < 3 else end + 1
            vals = np.frombuffer(b, dtype=base, count=nmemb, offset=q)
            q += nmemb * base.itemsize
            if sorted(names) == ["FALSE", "TRUE"]:
                return ("bool", base, {int(v): n == "TRUE" for n, v in zip(names, vals)}), q - p
            return ("num", base), q - p
        if cls == 9:
            vtype = bits & 0x0F
            base, used = self.datatype(q)
            if vtype == 1:
                return ("vstr",), 8 + used
            return ("vlen", base), 8 + used
        raise H5Error(f"datatype class {cls} is not supported")

    # ---- raw data
    def _gheap_object(self, coll: int, index: int) -> bytes:
        b = self.b
        coll += self.base
        if b[coll:coll + 4] != b"GCOL":
            raise H5Error("bad global heap signature")
        csize, = struct.unpack_from("<Q", b, coll + 8)
        p = coll + 16
        end = coll + csize
        while p + 16 <= end:
            idx, _ref, _res, osize = struct.unpack_from("<HHIQ", b, p)
            if idx == 0:
                break
            if idx == index:
                return b[p + 16:p + 16 + osize]
            p += 16 + ((osize + 7) // 8) * 8
        raise H5Error("global heap object not found")

    def _chunks(self, addr: int, rank: int, out):
        b = self.b
        addr += self.base
        if b[addr:addr + 4] != b"TREE":
            raise H5Error("bad chunk B-tree signature")
        ntype, level, used = struct.unpack_from("<BBH", b, addr + 4)
        if ntype != 1:
            raise H5Error("expected a chunk B-tree node")
        ksize = 8 + 8 * (rank + 1)
        p = addr + 24
        for _ in range(used):
            csize, fmask = struct.unpack_from("<II", b, p)
            offs = struct.unpack_from(f"<{rank + 1}Q", b, p + 8)
            child, = struct.unpack_from("<Q", b, p + ksize)
            if level > 0:
                self._chunks(child, rank, out)
            else:
                if fmask != 0:
                    raise H5Error("filtered chunks are not supported")
                out.append((offs[:rank], child + self.base, csize))
            p += ksize + 8

    def dataset(self, msgs):
        b = self.b
        shape = dtype = layout = None
        for mtype, body, _msize in msgs:
            if mtype == 0x0001:
                ver, rank, flags = b[body], b[body + 1], b[body + 2]
                if ver == 1:
                    q = body + 8
                elif ver == 2:
                    q = body + 4
                    if b[body + 3] == 2:  # null dataspace
                        shape = None
                        continue
                else:
                    raise H5Error(f"dataspace version {ver} unsupported")
                shape = struct.unpack_from(f"<{rank}Q", b, q) if rank else ()
            elif mtype == 0x0003:
                dtype, _ = self.datatype(body)
            elif mtype == 0x0008:
                layout = body
            elif mtype == 0x000B:
                raise H5Error("filtered (compressed) datasets are not supported")
        if dtype is None or layout is None or shape is None:
            raise H5Error("object is not a (supported) dataset")
        count = int(np.prod(shape)) if shape else 1
        kind = dtype[0]
        if kind in ("num", "bool"):
            item = dtype[1].itemsize
        elif kind == "str":
            item = dtype[1]
        else:
            item = 16
        lver, lcls = b[layout], b[layout + 1]
        if lver != 3:
            raise H5Error(f"data layout message version {lver} unsupported")
        if lcls == 0:
            size, = struct.unpack_from("<H", b, layout + 2)
            raw = b[layout + 4:layout + 4 + size]
        elif lcls == 1:
            addr, size = struct.unpack_from("<QQ", b, layout + 2)
            raw = b"\0" * (count * item) if addr == _UNDEF else b[addr + self.base:addr + self.base + size]
        elif lcls == 2:
            rank1 = b[layout + 2]
            addr, = struct.unpack_from("<Q", b, layout + 3)
            cdims = struct.unpack_from(f"<{rank1}I", b, layout + 11)
            rank = rank1 - 1
            if kind not in ("num", "bool"):
                raise H5Error("chunked non-numeric datasets are not supported")
            arr = np.zeros(shape, dtype=dtype[1])
            if addr != _UNDEF:
                chunks = []
                self._chunks(addr, rank, chunks)
                for offs, caddr, csize in chunks:
                    blk = np.frombuffer(b, dtype=dtype[1], count=csize // item, offset=caddr).reshape(cdims[:rank])
                    sl = tuple(slice(o, min(o + c, s)) for o, c, s in zip(offs, cdims, shape))
                    arr[sl] = blk[tuple(slice(0, s.stop - s.start) for s in sl)]
            return self._finish(arr, dtype, shape)
        else:
            raise H5Error(f"layout class {lcls} unsupported")
        if kind in ("num", "bool"):
            arr = np.frombuffer(raw, dtype=dtype[1], count=count).reshape(shape)
            return self._finish(arr, dtype, shape)
        if kind == "str":
            vals = [raw[i * item:(i + 1) * item].split(b"\0", 1)[0].decode("utf-8") for i in range(count)]
        elif kind == "vstr":
            vals = []
            for i in range(count):
                _n, coll, idx = struct.unpack_from("<IQI", raw, 16 * i)
                vals.append(self._gheap_object(coll, idx).decode("utf-8"))
        else:
            raise H5Error("variable-length sequences are not supported")
        return vals[0] if shape == () else np.array(vals, dtype=object).reshape(shape)

    @staticmethod
    def _finish(arr, dtype, shape):
        if dtype[0] == "bool":
            arr = np.vectorize(lambda v: dtype[2][int(v)], otypes=[bool])(arr) if arr.size else arr.astype(bool)
        else:
            arr = arr.astype(arr.dtype.newbyteorder("="), copy=True)
        return arr[()] if shape == () else arr

    def node(self, addr: int, _seen=None) -> Any:
        seen = set() if _seen is None else _seen
        if addr in seen or addr + self.base + 16 > len(self.b):
            raise H5Error("corrupt file: object header address outside the file or visited twice on one path")
        msgs = self.messages(addr)
        links = self.links(msgs)
        if links is not None:
            return {name: self.node(a, seen | {addr}) for name, a in links.items()}
        return self.dataset(msgs)


def load(path: str) -> Dict[str, Any]:
    """Whole file as nested dicts (groups) of numpy arrays / scalars / str (datasets).  Attributes are ignored."""
    with open(path, "rb") as f:
        data = f.read()
    r = _Reader(data)
    try:
        tree = r.node(r.root_header)
    except H5Error:
        raise
    except (struct.error, IndexError, RecursionError, ValueError, UnicodeDecodeError) as exc:  # truncated / corrupt
        raise H5Error(f"'{path}': corrupt or truncated HDF5 structures ({exc})") from exc
    if not isinstance(tree, dict):
        raise H5Error("root object is not a group")
    return tree


# ======================================================================================================= writer
def _pad8(n: int) -> int:
    return (n + 7) & ~7


class _Writer:
    """Append-only image builder.  Everything is 8-byte aligned; addresses are absolute (base address 0)."""

    LEAF_K = 4       # symbol nodes hold up to 2 K entries
    INTERNAL_K = 16  # B-tree nodes hold up to 2 K children

    def __init__(self):
        self.buf = bytearray(96)  # superblock v0 (56 bytes) + root symbol table entry (40 bytes)
        self.gheap = []           # variable-length string payloads -> one global heap collection at the end
        self.gheap_fixups = []    # (position of the 8-byte collection address inside buf)

    def alloc(self, data: bytes) -> int:
        addr = len(self.buf)
        self.buf += data
        self.buf += b"\0" * (_pad8(len(self.buf)) - len(self.buf))
        return addr

    # ---- datatype / dataspace / layout messages
    @staticmethod
    def _dtype_msg(kind):
        if kind == "f8":
            return struct.pack("<BBBBI", 0x11, 0x20, 0x3F, 0x00, 8) + struct.pack("<HHBBBBI", 0, 64, 52, 11, 0, 52, 1023)
        if kind == "i8":
            return struct.pack("<BBBBI", 0x10, 0x08, 0x00, 0x00, 8) + struct.pack("<HH", 0, 64)
        if kind == "i1":
            return struct.pack("<BBBBI", 0x10, 0x08, 0x00, 0x00, 1) + struct.pack("<HH", 0, 8)
        if kind == "c16":  # h5py: compound {r: f8 @0, i: f8 @8}
            f8 = _Writer._dtype_msg("f8")
            body = b""
            for name, off in (("r", 0), ("i", 8)):
                body += name.encode() + b"\0" * (8 - len(name)) + struct.pack("<IB3xII16x", off, 0, 0, 0) + f8
            return struct.pack("<BBBBI", 0x16, 0x02, 0x00, 0x00, 16) + body
        if kind == "bool":  # h5py: enum over int8 {FALSE: 0, TRUE: 1}
            base = _Writer._dtype_msg("i1")
            names = b"FALSE\0\0\0" + b"TRUE\0\0\0\0"
            return struct.pack("<BBBBI", 0x18, 0x02, 0x00, 0x00, 1) + base + names + bytes([0, 1])
        if kind == "vstr":  # variable-length UTF-8 string, null-terminated padding
            base = struct.pack("<BBBBI", 0x10, 0x00, 0x00, 0x00, 1) + struct.pack("<HH", 0, 8)  # 1-byte char, as libhdf5
            return struct.pack("<BBBBI", 0x19, 0x01, 0x01, 0x00, 16) + base
        raise H5Error(f"cannot write dtype {kind}")

    def dataset(self, value) -> int:
        if isinstance(value, str):
            kind, shape = "vstr", ()
            self.gheap.append(value.encode("utf-8"))
            raw = struct.pack("<IQI", len(self.gheap[-1]), 0, len(self.gheap))
            fix = True
        else:
            arr = np.asarray(value)
            shape = arr.shape
            if arr.dtype == bool:
                kind, raw = "bool", arr.astype(np.int8).tobytes()
            elif arr.dtype.kind in "iu":
                kind, raw = "i8", arr.astype("<i8").tobytes()
            elif arr.dtype.kind == "f":
                kind, raw = "f8", arr.astype("<f8").tobytes()
            elif arr.dtype.kind == "c":
                kind, raw = "c16", arr.astype("<c16").tobytes()
            else:
                raise H5Error(f"cannot write array of dtype {arr.dtype}")
            fix = False
        data_addr = self.alloc(raw) if raw else _UNDEF
        if fix:
            self.gheap_fixups.append(data_addr + 4)
        rank = len(shape)
        dims = b"".join(struct.pack("<Q", s) for s in shape)
        space = struct.pack("<BBB5x", 1, rank, 1 if rank else 0) + dims + dims  # v1 simple dataspace, max dims = dims
        dtype = self._dtype_msg(kind)
        # fill value message v2 exactly as libhdf5 writes it: late allocation, write-if-set (never for vlen), size 0
        fill = struct.pack("<BBBBI", 2, 2, 0 if kind == "vstr" else 2, 1, 0)
        layout = struct.pack("<BBQQ", 3, 1, data_addr, len(raw)) + b"\0" * 6
        return self._object_header([(0x0001, space), (0x0003, dtype), (0x0005, fill), (0x0008, layout)])

    def _object_header(self, msgs) -> int:
        body = b""
        for mtype, data in msgs:
            data = data + b"\0" * (_pad8(len(data)) - len(data))
            flags = 1 if mtype in (0x0003, 0x0005, 0x0008) else 0  # "constant" messages, as libhdf5 marks them
            body += struct.pack("<HHB3x", mtype, len(data), flags) + data
        hdr = struct.pack("<BBHII4x", 1, 0, len(msgs), 1, len(body))
        return self.alloc(hdr + body)

    # ---- groups
    def group(self, entries: Dict[str, int]) -> tuple:
        """entries: name -> object header address.  Returns (header address, btree address, heap address)."""
        names = sorted(entries, key=lambda s: s.encode("utf-8"))
        heap = bytearray(b"\0" * 8)  # offset 0: the empty string (B-tree key 0)
        offs = {}
        for n in names:
            offs[n] = len(heap)
            e = n.encode("utf-8") + b"\0"
            heap += e + b"\0" * (_pad8(len(e)) - len(e))
        free_off = len(heap)
        heap += struct.pack("<QQ", 1, 16)  # one free block: next = 1 (none), size 16
        heap_data = self.alloc(bytes(heap))
        heap_addr = self.alloc(b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap), free_off, heap_data))
        cap = 2 * self.LEAF_K
        chunks = [names[i:i + cap] for i in range(0, len(names), cap)] or [[]]
        children = []  # (address, heap offset of the largest name below it)
        for ch in chunks:
            snod = b"SNOD" + struct.pack("<BBH", 1, 0, len(ch))
            for n in ch:
                snod += struct.pack("<QQII16x", offs[n], entries[n], 0, 0)
            snod += b"\0" * (40 * (cap - len(ch)))
            children.append((self.alloc(snod), offs[ch[-1]] if ch else 0))
        # v1 B-tree over the symbol nodes: as many levels as needed, 2 K children per node, siblings linked
        fan = 2 * self.INTERNAL_K
        level = 0
        while True:
            nodes = []
            prev_key = 0   # key 0 of the leftmost node: the empty string at heap offset 0
            prev_addr = None
            for i in range(0, len(children), fan):
                grp = children[i:i + fan]
                tree = b"TREE" + struct.pack("<BBHQQ", 0, level, len(grp), prev_addr if prev_addr is not None else _UNDEF,
                                             _UNDEF) + struct.pack("<Q", prev_key)
                for addr, last_key in grp:
                    tree += struct.pack("<QQ", addr, last_key)
                tree += b"\0" * (16 * (fan - len(grp)))
                addr = self.alloc(tree)
                if prev_addr is not None:
                    struct.pack_into("<Q", self.buf, prev_addr + 16, addr)  # right sibling of the previous node
                prev_addr, prev_key = addr, grp[-1][1]
                nodes.append((addr, prev_key))
            if len(nodes) == 1:
                tree_addr = nodes[0][0]
                break
            children = nodes
            level += 1
        header = self._object_header([(0x0011, struct.pack("<QQ", tree_addr, heap_addr))])
        return header, tree_addr, heap_addr

    def node(self, value) -> tuple:
        if isinstance(value, dict):
            entries = {}
            for k, v in value.items():
                entries[str(k)] = self.node(v)[0]
            return self.group(entries)
        return (self.dataset(value), None, None)

    def finish(self, root) -> bytes:
        header, tree, heap = root
        if self.gheap:
            objs = b""
            for i, payload in enumerate(self.gheap, 1):
                objs += struct.pack("<HHIQ", i, 0, 0, len(payload)) + payload + b"\0" * (_pad8(len(payload)) - len(payload))
            size = max(4096, _pad8(16 + len(objs) + 16))
            free = size - 16 - len(objs)
            coll = b"GCOL" + struct.pack("<B3xQ", 1, size) + objs + struct.pack("<HHIQ", 0, 0, 0, free)
            coll += b"\0" * (size - len(coll))
            addr = self.alloc(coll)
            for pos in self.gheap_fixups:
                struct.pack_into("<Q", self.buf, pos, addr)
        eof = len(self.buf)
        sb = _SIG + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, self.LEAF_K, self.INTERNAL_K, 0)
        sb += struct.pack("<QQQQ", 0, _UNDEF, eof, _UNDEF)
        sb += struct.pack("<QQII", 0, header, 1, 0) + struct.pack("<QQ", tree, heap)
        assert len(sb) == 96
        self.buf[:96] = sb
        return bytes(self.buf)


def save(tree: Dict[str, Any], path: str) -> None:
    """Write nested dicts of arrays / scalars / str as an HDF5 file (see the module docstring for the subset)."""
    w = _Writer()
    root = w.node(tree)
    with open(path, "wb") as f:
        f.write(w.finish(root))
