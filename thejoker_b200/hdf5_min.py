"""A minimal, dependency-free, read-only HDF5 reader: just enough of the file format to
read the prior cache the reference writes where h5py is absent.

What the reference puts on disk (thejoker/samples.py:535-545, ``write_table_hdf5(...,
compression=False, serialize_meta=True, maxshape=(None,))`` through h5py with the library's
default "earliest" format bounds):

  * a version-0 superblock; old-style groups (symbol-table message -> version-1 B-tree of
    symbol-table nodes + a local heap with the link names);
  * dataset ``samples``: 1-D, compound datatype with one IEEE double member per column,
    *chunked* layout (the table is resizable) indexed by a version-1 B-tree, no filters;
  * dataset ``samples.__table_column_meta__``: 1-D array of fixed-length byte strings (the
    YAML header), contiguous layout.

Implemented from the public "HDF5 File Format Specification, Version 3.0": superblock
versions 0 and 1; version-1 object headers with continuation blocks; dataspace messages
(versions 1, 2); datatype classes fixed-point, floating-point, string and compound
(member encodings of datatype versions 1-3); data layout message versions 1-3 (compact,
contiguous, chunked); version-1 B-trees (group nodes and raw-data chunk nodes); local
heaps; symbol-table nodes.  Anything else -- filters / compression, new-style groups
(superblock 2/3, fractal heaps), variable-length data -- raises NotImplementedError with
the reason, never a wrong answer.

Interface: the slice of h5py's that ``cache.py`` uses --
``File(path)``, ``f[name]`` -> ``Dataset`` with ``shape``, ``dtype``, ``len()``,
``d[lo:hi]``, ``d[()]``, and the context-manager protocol.
"""
from __future__ import annotations

import mmap
import struct

import numpy as np

__all__ = ["File", "Dataset"]

_SIG = b"\x89HDF\r\n\x1a\n"
_UNDEF = 0xFFFFFFFFFFFFFFFF


class File:
    def __init__(self, filename, mode="r"):
        if mode != "r":
            raise ValueError("hdf5_min is read-only")
        self._fh = open(filename, "rb")
        try:
            self._buf = mmap.mmap(self._fh.fileno(), 0, access=mmap.ACCESS_READ)
        except ValueError:
            self._fh.close()
            raise OSError(f"{filename}: empty file")
        try:
            self._read_superblock()
            self._links = {}
            self._walk_group(self._root_btree, self._root_heap, "")
        except Exception:
            self.close()
            raise

    # -- context manager / h5py-like access ------------------------------------------
    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def close(self):
        if getattr(self, "_buf", None) is not None:
            self._buf.close()
            self._buf = None
        self._fh.close()

    def keys(self):
        return list(self._links)

    def __contains__(self, name):
        return name.strip("/") in self._links

    def __getitem__(self, name):
        name = name.strip("/")
        if name not in self._links:
            raise KeyError(name)
        return Dataset(self, self._links[name])

    # -- primitives ---------------------------------------------------------------------
    def _u(self, off, size):
        return int.from_bytes(self._buf[off:off + size], "little")

    def _addr(self, off):
        """An 'offset' field: file address relative to the base address."""
        a = self._u(off, self._so)
        return None if a == (1 << (8 * self._so)) - 1 else a + self._base

    def _read_superblock(self):
        buf, off = self._buf, 0
        while True:  # the superblock sits at 0 or at a power of two >= 512 (user block)
            if buf[off:off + 8] == _SIG:
                break
            off = 512 if off == 0 else off * 2
            if off + 8 > len(buf):
                raise OSError("not an HDF5 file (no superblock signature)")
        ver = buf[off + 8]
        if ver > 1:
            raise NotImplementedError(
                f"HDF5 superblock version {ver} (new-style groups): written with libver='latest'; "
                "the reference writes version 0 -- re-save the file with h5py's default libver")
        self._so, self._sl = buf[off + 13], buf[off + 14]
        p = off + 24 + (4 if ver == 1 else 0)
        self._base = 0
        base = self._u(p, self._so)
        # the base address is the superblock's own position when a user block precedes it
        self._base = base
        p += 4 * self._so  # base, free-space info, end of file, driver info
        # root group symbol-table entry
        obj = self._addr(p + self._so)
        cache_type = self._u(p + 2 * self._so, 4)
        scratch = p + 2 * self._so + 8
        if cache_type == 1:
            self._root_btree, self._root_heap = self._addr(scratch), self._addr(scratch + self._so)
        else:
            msgs = self._object_header(obj)
            st = [m for m in msgs if m[0] == 0x11]
            if not st:
                raise NotImplementedError("root group without a symbol table (new-style group)")
            self._root_btree, self._root_heap = self._addr(st[0][1]), self._addr(st[0][1] + self._so)

    # -- object headers -------------------------------------------------------------------
    def _object_header(self, addr):
        """[(type, data offset, data size, flags)] of a version-1 object header."""
        buf = self._buf
        if buf[addr:addr + 4] == b"OHDR":
            raise NotImplementedError("version-2 object header (file written with libver='latest')")
        if buf[addr] != 1:
            raise OSError(f"bad object header version {buf[addr]} at {addr}")
        n_msgs = self._u(addr + 2, 2)
        size = self._u(addr + 8, 4)
        blocks = [(addr + 16, size)]
        out = []
        while blocks and len(out) < n_msgs:
            p, left = blocks.pop(0)
            end = p + left
            while p + 8 <= end and len(out) < n_msgs:
                mtype, msize, flags = self._u(p, 2), self._u(p + 2, 2), buf[p + 4]
                data = p + 8
                if mtype == 0x10:  # continuation
                    blocks.append((self._addr(data), self._u(data + self._so, self._sl)))
                out.append((mtype, data, msize, flags))
                p = data + msize
        return out

    # -- groups -----------------------------------------------------------------------------
    def _heap_string(self, heap_addr, offset):
        buf = self._buf
        if buf[heap_addr:heap_addr + 4] != b"HEAP":
            raise OSError("bad local heap signature")
        data = self._addr(heap_addr + 8 + 2 * self._sl)
        start = data + offset
        end = buf.find(b"\x00", start)
        return buf[start:end].decode("utf-8")

    def _walk_group(self, btree, heap, prefix):
        buf = self._buf
        if buf[btree:btree + 4] != b"TREE" or buf[btree + 4] != 0:
            raise OSError("bad group B-tree node")
        level, used = buf[btree + 5], self._u(btree + 6, 2)
        p = btree + 8 + 2 * self._so
        for i in range(used):
            child = self._addr(p + self._sl + i * (self._sl + self._so))
            if level > 0:
                self._walk_group(child, heap, prefix)
                continue
            if buf[child:child + 4] != b"SNOD":
                raise OSError("bad symbol-table node")
            n_sym = self._u(child + 6, 2)
            e = child + 8
            for _ in range(n_sym):
                name = self._heap_string(heap, self._u(e, self._so))
                obj = self._addr(e + self._so)
                cache_type = self._u(e + 2 * self._so, 4)
                full = prefix + name
                if cache_type == 1:  # a sub-group with cached B-tree / heap addresses
                    sc = e + 2 * self._so + 8
                    self._walk_group(self._addr(sc), self._addr(sc + self._so), full + "/")
                else:
                    msgs = self._object_header(obj)
                    st = [m for m in msgs if m[0] == 0x11]
                    if st:
                        self._walk_group(self._addr(st[0][1]), self._addr(st[0][1] + self._so),
                                         full + "/")
                    else:
                        self._links[full] = obj
                e += 2 * self._so + 24

    # -- datatype -----------------------------------------------------------------------------
    def _datatype(self, p):
        """(numpy dtype, bytes consumed) of the datatype message at p."""
        buf = self._buf
        cls, ver = buf[p] & 0x0F, buf[p] >> 4
        bits = self._u(p + 1, 3)
        size = self._u(p + 4, 4)
        order = ">" if bits & 1 else "<"
        if cls == 0:  # fixed-point
            kind = "i" if bits & 0x08 else "u"
            return np.dtype(f"{order}{kind}{size}"), 8 + 4
        if cls == 1:  # floating-point (IEEE layouts only)
            if size not in (2, 4, 8):
                raise NotImplementedError(f"{size}-byte floating-point type")
            return np.dtype(f"{order}f{size}"), 8 + 12
        if cls == 3:  # fixed-length string
            return np.dtype(f"S{size}"), 8
        if cls == 6:  # compound
            n_mem = bits & 0xFFFF
            q = p + 8
            names, formats, offsets = [], [], []
            for _ in range(n_mem):
                end = buf.find(b"\x00", q)
                name = buf[q:end].decode("utf-8")
                if ver < 3:
                    q += ((end - q) // 8 + 1) * 8  # null-terminated, padded to a multiple of 8
                else:
                    q = end + 1
                if ver == 3:
                    nb = 1 if size < 256 else 2 if size < 65536 else 3 if size < (1 << 24) else 4
                    off = self._u(q, nb)
                    q += nb
                else:
                    off = self._u(q, 4)
                    q += 4
                if ver == 1:
                    rank = buf[q]
                    if rank:
                        raise NotImplementedError("array member in a version-1 compound datatype")
                    q += 1 + 3 + 4 + 4 + 16
                mt, used = self._datatype(q)
                q += used
                names.append(name)
                formats.append(mt)
                offsets.append(off)
            return np.dtype({"names": names, "formats": formats, "offsets": offsets,
                             "itemsize": size}), q - p
        if cls == 9:
            raise NotImplementedError("variable-length datatype (global heap)")
        raise NotImplementedError(f"HDF5 datatype class {cls}")


class Dataset:
    def __init__(self, f, obj_addr):
        self._f = f
        self.shape = self.dtype = None
        self._layout = None
        for mtype, data, size, flags in f._object_header(obj_addr):
            if mtype == 0x01:
                self.shape = self._dataspace(data)
            elif mtype == 0x03:
                self.dtype, _ = f._datatype(data)
            elif mtype == 0x08:
                self._layout = self._data_layout(data)
            elif mtype == 0x0B:
                nf = f._buf[data + 1]
                if nf:
                    raise NotImplementedError(
                        "filtered (compressed / shuffled) dataset: the reference writes its "
                        "prior cache with compression=False")
        if self.shape is None or self.dtype is None or self._layout is None:
            raise OSError("not a dataset (no dataspace / datatype / layout message)")

    def _dataspace(self, p):
        f, buf = self._f, self._f._buf
        ver, rank = buf[p], buf[p + 1]
        q = p + (8 if ver == 1 else 4)
        return tuple(f._u(q + i * f._sl, f._sl) for i in range(rank))

    def _data_layout(self, p):
        f, buf = self._f, self._f._buf
        ver = buf[p]
        if ver == 3:
            cls = buf[p + 1]
            if cls == 0:
                n = f._u(p + 2, 2)
                return ("compact", p + 4, n)
            if cls == 1:
                return ("contiguous", f._addr(p + 2), f._u(p + 2 + f._so, f._sl))
            if cls == 2:
                nd = buf[p + 2]
                bt = f._addr(p + 3)
                dims = [f._u(p + 3 + f._so + 4 * i, 4) for i in range(nd)]
                return ("chunked", bt, dims)
            raise NotImplementedError(f"data layout class {cls}")
        if ver in (1, 2):
            nd, cls = buf[p + 1], buf[p + 2]
            q = p + 8
            addr = None
            if cls != 0:
                addr = f._addr(q)
                q += f._so
            dims = [f._u(q + 4 * i, 4) for i in range(nd)]
            q += 4 * nd
            if cls == 0:
                return ("compact", q + 4, f._u(q, 4))
            if cls == 1:
                return ("contiguous", addr, None)
            return ("chunked", addr, dims)
        raise NotImplementedError(f"data layout message version {ver} (libver='latest')")

    # -- h5py-like access -------------------------------------------------------------------
    def __len__(self):
        if not self.shape:
            raise TypeError("scalar dataset")
        return self.shape[0]

    def __getitem__(self, key):
        if key == () or key is Ellipsis:
            if len(self.shape) > 1:  # whole N-D array: contiguous / compact storage only
                if self._layout[0] == "chunked":
                    raise NotImplementedError("chunked dataset of rank > 1")
                n = int(np.prod(self.shape))
                addr = self._layout[1]
                flat = np.zeros(n, dtype=self.dtype) if addr is None else np.array(
                    np.frombuffer(self._f._buf, dtype=self.dtype, count=n, offset=addr))
                return flat.reshape(self.shape)
            out = self._read(0, self.shape[0] if self.shape else 1)
            return out if self.shape else out[0]
        if isinstance(key, str):
            return self._read(0, len(self))[key]
        if isinstance(key, slice):
            lo, hi, step = key.indices(len(self))
            out = self._read(lo, max(lo, hi))
            return out[::step] if step != 1 else out
        raise TypeError("hdf5_min datasets take a slice, a field name or ()")

    def _read(self, lo, hi):
        if len(self.shape) > 1:
            raise NotImplementedError("only 1-D (and scalar) datasets")
        f, item = self._f, self.dtype.itemsize
        n = hi - lo
        out = np.empty(n, dtype=self.dtype)
        raw = out.view(np.uint8).reshape(-1)
        kind = self._layout[0]
        if kind in ("compact", "contiguous"):
            addr = self._layout[1]
            if addr is None:  # never written: fill value (zeros)
                raw[:] = 0
            else:
                raw[:] = np.frombuffer(f._buf, dtype=np.uint8, count=n * item, offset=addr + lo * item)
            return out
        _, btree, dims = self._layout
        rows = dims[0]
        if dims[-1] != item:
            raise OSError("chunk element size does not match the datatype")
        raw[:] = 0  # chunks that were never allocated read as the fill value
        if btree is not None:
            self._read_chunks(btree, lo, hi, rows, raw, item)
        return out

    def _read_chunks(self, node, lo, hi, rows, raw, item):
        f, buf = self._f, self._f._buf
        if buf[node:node + 4] != b"TREE" or buf[node + 4] != 1:
            raise OSError("bad chunk B-tree node")
        level, used = buf[node + 5], f._u(node + 6, 2)
        key_size = 8 + 8 * 2  # chunk size, filter mask, offsets of (rank 1 + element) dims
        p = node + 8 + 2 * f._so
        for i in range(used):
            k = p + i * (key_size + f._so)
            nbytes, mask, start = f._u(k, 4), f._u(k + 4, 4), f._u(k + 8, 8)
            child = f._addr(k + key_size)
            if level > 0:
                nxt = f._u(k + key_size + f._so + 8, 8)  # first row of the next key
                if (i + 1 < used and nxt <= lo) or start >= hi:
                    continue
                self._read_chunks(child, lo, hi, rows, raw, item)
                continue
            if start >= hi or start + rows <= lo:
                continue
            if mask or nbytes != rows * item:
                raise NotImplementedError("filtered chunk")
            a, b = max(lo, start), min(hi, start + rows)
            src = np.frombuffer(buf, dtype=np.uint8, count=(b - a) * item,
                                offset=child + (a - start) * item)
            raw[(a - lo) * item:(b - lo) * item] = src
