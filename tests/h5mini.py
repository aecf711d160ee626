"""TEST INFRASTRUCTURE: a minimal HDF5 *writer* (superblock 0, version-1 object headers, symbol-table groups, contiguous /
compact / chunked + deflate + shuffle datasets, version-1 attributes), enough to lay out MINC 2.0 volumes the way libminc
does, so that the product's reader (rminc_b200/csrc/minc2_reader.cpp) can be exercised without libhdf5 in the image.
Written from the HDF5 File Format Specification like the reader, so the two agreeing proves consistency, not
correctness: the reader's anchor to the real library is the libhdf5-written file of tests/test_minc2_reader.py."""
from __future__ import annotations

import struct
import zlib

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF


def _pad8(b: bytes) -> bytes:
    return b + b"\0" * (-len(b) % 8)


class Dataset:
    def __init__(self, data, attrs=None, chunks=None, compress=None, shuffle=False, big_endian=False, compact=False, fletcher=False):
        self.data = np.asarray(data)
        self.attrs = attrs or {}
        self.chunks = chunks
        self.compress = compress
        self.shuffle = shuffle
        self.big_endian = big_endian
        self.compact = compact
        self.fletcher = fletcher


class Group:
    def __init__(self, children=None, attrs=None):
        self.children = children or {}
        self.attrs = attrs or {}


class Writer:
    def __init__(self, user_block: int = 0, leaf_k: int = 64, chunk_k: int = 32, split_headers: bool = False, max_dims: bool = False):
        self.split_headers = split_headers
        self.max_dims = max_dims
        self.buf = bytearray(96)
        self.user_block = user_block
        self.leaf_k = leaf_k
        self.chunk_k = chunk_k

    # -- allocation
    def put(self, b: bytes) -> int:
        addr = len(self.buf)
        self.buf += _pad8(b)
        return addr

    # -- messages
    @staticmethod
    def datatype(dt: np.dtype, big_endian=False) -> bytes:
        dt = np.dtype(dt)
        if dt.kind in "iu":
            bits = (1 if big_endian else 0) | (8 if dt.kind == "i" else 0)
            return struct.pack("<BBBBI", 0x10, bits, 0, 0, dt.itemsize) + struct.pack("<HH", 0, 8 * dt.itemsize)
        if dt.kind == "f":
            if dt.itemsize == 8:
                props = struct.pack("<HHBBBBI", 0, 64, 52, 11, 0, 52, 1023)
                b1, b2 = 0x20, 63
            else:
                props = struct.pack("<HHBBBBI", 0, 32, 23, 8, 0, 23, 127)
                b1, b2 = 0x20, 31
            return struct.pack("<BBBBI", 0x11, b1 | (1 if big_endian else 0), b2, 0, dt.itemsize) + props
        if dt.kind == "S":
            return struct.pack("<BBBBI", 0x13, 0, 0, 0, dt.itemsize)
        raise TypeError(dt)

    def dataspace(self, shape) -> bytes:
        dims = b"".join(struct.pack("<Q", s) for s in shape)
        if self.max_dims and len(shape):                 # flag bit 0: the maximum dimensions follow the current ones
            return struct.pack("<BBBBI", 1, len(shape), 1, 0, 0) + dims + dims
        return struct.pack("<BBBBI", 1, len(shape), 0, 0, 0) + dims

    def attribute(self, name: str, value) -> bytes:
        if isinstance(value, str):
            value = np.bytes_(value.encode() + b"\0")          # libminc writes the terminator
            a = np.asarray(value)
        else:
            a = np.asarray(value)
            if a.dtype.kind == "U":
                a = a.astype("S")
        nm = name.encode() + b"\0"
        dt = self.datatype(a.dtype)
        ds = self.dataspace(a.shape)
        head = struct.pack("<BBHHH", 1, 0, len(nm), len(dt), len(ds))
        return head + _pad8(nm) + _pad8(dt) + _pad8(ds) + a.tobytes()

    def header(self, messages) -> int:
        """Version-1 object header.  With self.split_headers the messages after the third go to a continuation block
        (message 0x0010), the way libhdf5 grows the header of an object that collects attributes after its creation -- a MINC
        image variable does -- with a NIL message padding the first block."""
        def pack(msgs):
            body = b""
            for mtype, flags, data in msgs:
                d = _pad8(data)
                body += struct.pack("<HHBBBB", mtype, len(d), flags, 0, 0, 0) + d
            return body

        n = len(messages)
        if self.split_headers and n > 3:
            cont = pack(messages[3:])
            cont_addr = self.put(cont)
            first = messages[:3] + [(0x0000, 0, b"\0" * 8), (0x0010, 0, struct.pack("<QQ", cont_addr, len(cont)))]
            body = pack(first)
            n += 2
        else:
            body = pack(messages)
        return self.put(struct.pack("<BBHII", 1, 0, n, 1, len(body)) + b"\0" * 4 + body)

    # -- datasets
    def write_dataset(self, ds: Dataset) -> int:
        a = ds.data
        raw_dt = a.dtype.newbyteorder(">" if ds.big_endian else "<")
        es = a.dtype.itemsize
        msgs = [(0x01, 0, self.dataspace(a.shape)), (0x03, 1, self.datatype(a.dtype, ds.big_endian))]
        if ds.chunks is not None:
            chunks = tuple(int(c) for c in ds.chunks)
            assert len(chunks) == a.ndim
            filt, fl = b"", []
            if ds.shuffle:
                fl.append((2, [es]))
            if ds.compress is not None:
                fl.append((1, [int(ds.compress)]))
            if ds.fletcher:
                fl.append((3, []))
            if fl:
                filt = struct.pack("<BB6x", 1, len(fl))
                for fid, cd in fl:
                    filt += struct.pack("<HHHH", fid, 0, 1, len(cd)) + b"".join(struct.pack("<I", c) for c in cd)
                    if len(cd) & 1:
                        filt += b"\0" * 4
                msgs.append((0x0B, 0, filt))
            entries = []
            grid = [range(0, a.shape[d], chunks[d]) for d in range(a.ndim)]
            for off in np.ndindex(*[len(g) for g in grid]):
                o = tuple(grid[d][off[d]] for d in range(a.ndim))
                block = np.zeros(chunks, dtype=a.dtype)
                sl = tuple(slice(o[d], min(o[d] + chunks[d], a.shape[d])) for d in range(a.ndim))
                part = a[sl]
                block[tuple(slice(0, s) for s in part.shape)] = part
                raw = block.astype(raw_dt).tobytes()
                if ds.shuffle and es > 1:
                    raw = np.frombuffer(raw, np.uint8).reshape(-1, es).T.tobytes()
                if ds.compress is not None:
                    raw = zlib.compress(raw, int(ds.compress))
                if ds.fletcher:
                    raw += b"\0\0\0\0"          # the reader skips the checksum
                entries.append((o, len(raw), self.put(raw)))
            btree = self.chunk_btree(entries, a.ndim, es, a.shape, chunks)
            lay = struct.pack("<BBBQ", 3, 2, a.ndim + 1, btree) + b"".join(struct.pack("<I", c) for c in chunks + (es,))
        elif ds.compact:
            raw = a.astype(raw_dt).tobytes()
            lay = struct.pack("<BBH", 3, 0, len(raw)) + raw
        else:
            raw = a.astype(raw_dt).tobytes()
            addr = self.put(raw) if raw else UNDEF
            lay = struct.pack("<BBQQ", 3, 1, addr, len(raw))
        msgs.append((0x08, 1, lay))
        for k, v in ds.attrs.items():
            msgs.append((0x0C, 0, self.attribute(k, v)))
        return self.header(msgs)

    def chunk_btree(self, entries, rank, es, shape, chunks) -> int:
        def key(off, nbytes):
            return struct.pack("<II", nbytes, 0) + b"".join(struct.pack("<Q", o) for o in off) + struct.pack("<Q", 0)

        end_key = key(tuple(((s + c - 1) // c) * c for s, c in zip(shape, chunks)), 0)
        level, nodes = 0, [(o, nb, ad) for o, nb, ad in entries]
        cap = 2 * self.chunk_k
        while True:
            out = []
            for i in range(0, max(len(nodes), 1), cap):
                grp = nodes[i:i + cap]
                body = b"TREE" + struct.pack("<BBH", 1, level, len(grp)) + struct.pack("<QQ", UNDEF, UNDEF)
                for o, nb, ad in grp:
                    body += key(o, nb) + struct.pack("<Q", ad)
                body += end_key
                out.append((grp[0][0] if grp else (0,) * rank, grp[0][1] if grp else 0, self.put(body)))
            if len(out) == 1:
                return out[0][2]
            nodes, level = out, level + 1

    # -- groups
    def write_group(self, g: Group) -> int:
        names = sorted(g.children)
        addrs = {}
        for nm in names:
            c = g.children[nm]
            addrs[nm] = self.write_group(c) if isinstance(c, Group) else self.write_dataset(c)
        heap = bytearray(8)                      # offset 0: the empty name
        offs = {}
        for nm in names:
            offs[nm] = len(heap)
            heap += _pad8(nm.encode() + b"\0")
        heap_data = self.put(bytes(heap) + b"\0" * 16)
        heap_addr = self.put(b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap) + 16, len(heap), heap_data))
        snods, cap = [], 2 * self.leaf_k
        for i in range(0, max(len(names), 1), cap):
            grp = names[i:i + cap]
            body = b"SNOD" + struct.pack("<BBH", 1, 0, len(grp))
            for nm in grp:
                body += struct.pack("<QQII16x", offs[nm], addrs[nm], 0, 0)
            body += b"\0" * (40 * (cap - len(grp)))
            snods.append((offs[grp[-1]] if grp else 0, self.put(body)))
        tree = b"TREE" + struct.pack("<BBH", 0, 0, len(snods)) + struct.pack("<QQ", UNDEF, UNDEF) + struct.pack("<Q", 0)
        for last, ad in snods:
            tree += struct.pack("<QQ", ad, last)
        tree_addr = self.put(tree)
        msgs = [(0x11, 0, struct.pack("<QQ", tree_addr, heap_addr))]
        for k, v in g.attrs.items():
            msgs.append((0x0C, 0, self.attribute(k, v)))
        hdr = self.header(msgs)
        self._last = (tree_addr, heap_addr)
        return hdr

    def finish(self, root: Group) -> bytes:
        hdr = self.write_group(root)
        tree_addr, heap_addr = self._last
        sb = b"\x89HDF\r\n\x1a\n" + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, self.leaf_k, 16, 0)
        sb += struct.pack("<QQQQ", self.user_block, UNDEF, len(self.buf), UNDEF)
        sb += struct.pack("<QQII", 0, hdr, 1, 0) + struct.pack("<QQ", tree_addr, heap_addr)
        assert len(sb) == 96
        self.buf[:96] = sb
        return b"\0" * self.user_block + bytes(self.buf)


def write_h5(path, root: Group, **kw):
    data = Writer(**kw).finish(root)
    with open(path, "wb") as f:
        f.write(data)


def minc2_tree(stored, image_min=None, image_max=None, valid_range=None, steps=(1.0, 1.0, 1.0), starts=(0.0, 0.0, 0.0),
               dimorder=("zspace", "yspace", "xspace"), chunks=None, compress=4, big_endian=False, shuffle=False) -> Group:
    """The object tree libminc lays out for one volume (what RMINC's readers look at: the image, its ranges, the dimensions)."""
    stored = np.asarray(stored)
    dims = {}
    for d, nm in enumerate(dimorder):
        cos = [0.0, 0.0, 0.0]
        if nm in ("xspace", "yspace", "zspace"):
            cos["xyz".index(nm[0])] = 1.0
        dims[nm] = Dataset(np.int32(0), attrs={"varid": "MINC standard variable", "vartype": "dimension____", "version": "MINC Version    1.0",
                                               "class": "regular__", "length": np.int32(stored.shape[d]), "start": np.float64(starts[d]),
                                               "step": np.float64(steps[d]), "direction_cosines": np.asarray(cos), "units": "mm",
                                               "spacetype": "native____"})
    img_attrs = {"varid": "MINC standard variable", "vartype": "group________", "version": "MINC Version    1.0",
                 "complete": "true_", "dimorder": ",".join(dimorder)}
    if stored.dtype.kind in "iu":
        img_attrs["signtype"] = "signed__" if stored.dtype.kind == "i" else "unsigned"
    if valid_range is not None:
        img_attrs["valid_range"] = np.asarray(valid_range, dtype=np.float64)
    img0 = {"image": Dataset(stored, attrs=img_attrs, chunks=chunks, compress=compress if chunks is not None else None,
                             big_endian=big_endian, shuffle=shuffle)}
    if image_min is not None:
        lo, hi = np.asarray(image_min, dtype=np.float64), np.asarray(image_max, dtype=np.float64)
        rng_attrs = {"varid": "MINC standard variable", "vartype": "var_attribute", "version": "MINC Version    1.0"}
        if lo.ndim:
            rng_attrs["dimorder"] = ",".join(dimorder[:lo.ndim])
        img0["image-min"] = Dataset(lo, attrs=dict(rng_attrs))
        img0["image-max"] = Dataset(hi, attrs=dict(rng_attrs))
    return Group({"minc-2.0": Group({"dimensions": Group(dims), "image": Group({"0": Group(img0)}), "info": Group()},
                                    attrs={"ident": "rminc-b200 test file", "minc_version": "2.4.03", "history": "h5mini"})})


def write_minc2(path, stored, **kw):
    write_h5(path, minc2_tree(stored, **kw))
