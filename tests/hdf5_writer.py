"""Test infrastructure: writes an HDF5 file with the on-disk structures the reference's
``JokerSamples.write`` produces through h5py / libhdf5 with default ("earliest") format
bounds -- built byte by byte from the HDF5 File Format Specification, without any HDF5
library (none is in the image):

  superblock v0 -> root group (object header v1 + symbol-table message, B-tree v1 group
  node, local heap, symbol-table node) ->
    "samples"                       1-D resizable dataset, compound datatype (version 1
                                    encoding, one IEEE double per column), chunked layout
                                    (layout message v3) indexed by a B-tree v1 (optionally
                                    two levels), filter pipeline absent, fill-value and
                                    modification-time messages present, as libhdf5 writes them
    "samples.__table_column_meta__" 1-D dataset of fixed-length strings, contiguous layout

Only tests/ use it: it exists to feed thejoker_b200/hdf5_min.py bytes that were not
produced by the reader's own author-side model of a *stand-in module*; the reader is also
run on a real libhdf5-written file (scipy's testhdf5_7.4_GLNX86.mat) in the same test.
"""
import struct

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF


def _pad8(b):
    return b + b"\x00" * (-len(b) % 8)


def _msg(mtype, data, flags=0):
    data = _pad8(data)
    return struct.pack("<HHB3x", mtype, len(data), flags) + data


def _object_header(msgs, extra_block=None):
    """Version-1 object header; ``extra_block`` = (address, messages) puts those messages
    in a continuation block (libhdf5 does that when a header outgrows its first block)."""
    body = b"".join(msgs)
    n = len(msgs)
    if extra_block is not None:
        addr, more = extra_block
        cont = b"".join(more)
        body += _msg(0x10, struct.pack("<QQ", addr, len(cont)))
        n += 1 + len(more)
    return struct.pack("<BxHII4x", 1, n, 1, len(body)) + body


def _dt_double():
    # class 1 (floating point) version 1; little-endian, IEEE: bit field 0x20 0x3f 0x00
    return struct.pack("<B3BI", 0x11, 0x20, 0x3F, 0x00, 8) + struct.pack("<HHBBBBI", 0, 64, 52, 11, 0, 52, 1023)


def _dt_compound(names):
    out = struct.pack("<B3BI", 0x16, len(names) & 0xFF, len(names) >> 8, 0, 8 * len(names))
    for i, nm in enumerate(names):
        b = nm.encode() + b"\x00"
        out += b + b"\x00" * (-len(b) % 8)
        out += struct.pack("<IB3xII16x", 8 * i, 0, 0, 0)  # offset, rank 0, perm, reserved, 4 dims
        out += _dt_double()
    return out


def _dt_string(size):
    return struct.pack("<B3BI", 0x13, 0x00, 0x00, 0x00, size)  # null-terminated ASCII


def _dataspace(n, unlimited):
    flags = 1 if unlimited else 0
    out = struct.pack("<BBB5x", 1, 1, flags) + struct.pack("<Q", n)
    if unlimited:
        out += struct.pack("<Q", UNDEF)
    return out


def write_reference_style_hdf5(path, rows, header_lines, chunk_rows=64, two_level=False,
                               user_block=0):
    """rows: structured array of float64 fields; header_lines: list of bytes."""
    names = list(rows.dtype.names)
    item = 8 * len(names)
    n = len(rows)
    raw = np.zeros(n, dtype=[(c, "<f8") for c in names])
    for c in names:
        raw[c] = rows[c]
    slen = max(len(h) for h in header_lines) + 1
    meta = np.array(header_lines, dtype=f"S{slen}")

    blob = bytearray()

    def alloc(b, align=8):
        blob.extend(b"\x00" * (-len(blob) % align))
        a = len(blob)
        blob.extend(b)
        return a

    # superblock v0 (96 bytes with the root symbol-table entry), patched at the end
    alloc(b"\x00" * 96)

    # ---- raw data: chunks of `chunk_rows` rows; the last chunk is allocated in full ----
    n_chunks = max(1, -(-n // chunk_rows))
    chunk_addr = []
    for k in range(n_chunks):
        part = raw[k * chunk_rows:(k + 1) * chunk_rows].tobytes()
        part += b"\x00" * (chunk_rows * item - len(part))
        chunk_addr.append(alloc(part))
    meta_addr = alloc(meta.tobytes())

    # ---- chunk B-tree (node type 1): key = size, filter mask, offsets (row, 0) ----------
    def chunk_key(row):
        return struct.pack("<IIQQ", chunk_rows * item, 0, row, 0)

    def chunk_node(level, entries, last_row):
        """entries: [(first row, child address)]"""
        body = struct.pack("<4sBBHQQ", b"TREE", 1, level, len(entries), UNDEF, UNDEF)
        for row, child in entries:
            body += chunk_key(row) + struct.pack("<Q", child)
        body += chunk_key(last_row)
        return alloc(body)

    leaves_in = [(k * chunk_rows, a) for k, a in enumerate(chunk_addr)]
    if two_level and n_chunks >= 4:
        half = n_chunks // 2
        l0 = chunk_node(0, leaves_in[:half], half * chunk_rows)
        l1 = chunk_node(0, leaves_in[half:], n_chunks * chunk_rows)
        btree = chunk_node(1, [(0, l0), (half * chunk_rows, l1)], n_chunks * chunk_rows)
    else:
        btree = chunk_node(0, leaves_in, n_chunks * chunk_rows)

    # ---- dataset object headers ------------------------------------------------------------
    layout_chunked = struct.pack("<BBB", 3, 2, 2) + struct.pack("<Q", btree) + struct.pack("<II", chunk_rows, item)
    fill = struct.pack("<BBBB", 2, 2, 0, 0)             # fill value v2: alloc late, never write, undefined
    mtime = struct.pack("<B3xI", 1, 1700000000)
    # the modification time goes into a continuation block, as in headers libhdf5 has grown
    cont_msgs = [_msg(0x12, mtime)]
    cont_addr = alloc(b"".join(cont_msgs))
    samples_hdr = alloc(_object_header(
        [_msg(0x01, _dataspace(n, True)), _msg(0x03, _dt_compound(names), flags=1), _msg(0x05, fill),
         _msg(0x08, layout_chunked)], extra_block=(cont_addr, cont_msgs)))
    layout_contig = struct.pack("<BB", 3, 1) + struct.pack("<QQ", meta_addr, meta.nbytes)
    meta_hdr = alloc(_object_header(
        [_msg(0x01, _dataspace(len(meta), False)), _msg(0x03, _dt_string(slen), flags=1),
         _msg(0x05, fill), _msg(0x08, layout_contig)]))

    # ---- root group: local heap, symbol-table node, B-tree, object header -------------------
    link_names = ["samples", "samples.__table_column_meta__"]   # sorted, as the B-tree requires
    heap_data = bytearray(b"\x00" * 8)                            # offset 0: the empty name
    name_off = []
    for nm in link_names:
        name_off.append(len(heap_data))
        heap_data.extend(_pad8(nm.encode() + b"\x00"))
    heap_data.extend(b"\x00" * 16)
    heap_data_addr = alloc(bytes(heap_data))
    heap = alloc(struct.pack("<4sB3xQQQ", b"HEAP", 0, len(heap_data), UNDEF, heap_data_addr))
    snod = struct.pack("<4sBxH", b"SNOD", 1, 2)
    for off, hdr in zip(name_off, (samples_hdr, meta_hdr)):
        snod += struct.pack("<QQII16x", off, hdr, 0, 0)
    snod += b"\x00" * (40 * (8 - 2))                              # 2 * leaf K = 8 entries per node
    snod_addr = alloc(snod)
    group_tree = alloc(struct.pack("<4sBBHQQ", b"TREE", 0, 0, 1, UNDEF, UNDEF)
                       + struct.pack("<Q", 0) + struct.pack("<Q", snod_addr)
                       + struct.pack("<Q", name_off[-1]))
    root_hdr = alloc(_object_header([_msg(0x11, struct.pack("<QQ", group_tree, heap))]))

    eof = len(blob)
    sb = b"\x89HDF\r\n\x1a\n" + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, 4, 16, 0)
    sb += struct.pack("<QQQQ", user_block, UNDEF, eof, UNDEF)
    sb += struct.pack("<QQII", 0, root_hdr, 1, 0) + struct.pack("<QQ", group_tree, heap)
    blob[0:len(sb)] = sb
    with open(path, "wb") as f:
        f.write(b"\x00" * user_block)   # addresses in the file are relative to the base address
        f.write(bytes(blob))
    return path
