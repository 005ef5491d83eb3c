"""
Stack ingestion, host side (no GPU): the self-contained HDF5 codec, the reference's read_h5 / save_h5 / read_image /
write_image contracts (io/h5.py:17-212, io/rw.py:64-189) and the block reader that feeds the pipeline.

h5py / libhdf5 are not installed here, so the codec is pinned against itself (writer <-> reader), against structures
assembled by hand in this file from the format specification, and structurally (addresses, signatures, sizes).
"""

import os
import struct
import zlib

import numpy as np
import pytest

from barc4dip_b200.io import h5 as h5io
from barc4dip_b200.io import hdf5, read_image, write_image
from barc4dip_b200.io.stream import H5StackSource, concat_results

PATH = "entry_0000/measurement/data"


def _stack(shape, dtype, seed=0):
    rng = np.random.default_rng(seed)
    if np.dtype(dtype).kind == "f":
        return rng.standard_normal(shape).astype(dtype)
    info = np.iinfo(dtype)
    return rng.integers(info.min, info.max, size=shape, endpoint=True, dtype=dtype)


@pytest.mark.parametrize("dtype", ["uint8", "uint16", "int16", "int32", "uint32", "int64", "float32", "float64", "float16"])
def test_round_trip_every_element_type(tmp_path, dtype):
    a = _stack((5, 37, 50), dtype)
    p = tmp_path / "s.h5"
    hdf5.write_stack(p, a)
    with hdf5.H5File(p) as f:
        assert PATH in f and "entry_0000/nothing" not in f
        d = f[PATH]
        assert d.shape == a.shape and d.dtype == a.dtype and d.chunks is not None
        assert d.filters == [(hdf5.FILTER_DEFLATE, [4])]
        np.testing.assert_array_equal(d.read(), a)
        np.testing.assert_array_equal(d[()], a)


@pytest.mark.parametrize("shape,chunks", [((7, 33, 21), (2, 8, 8)), ((7, 33, 21), (1, 33, 21)), ((3, 64, 64), (3, 64, 64)),
                                          ((9, 10, 11), (4, 3, 5)), ((40, 48), (7, 9)), ((100,), (13,))])
@pytest.mark.parametrize("shuffle", [False, True])
def test_ragged_chunks_ranges_and_shuffle(tmp_path, shape, chunks, shuffle):
    a = _stack(shape, "uint16", seed=3)
    p = tmp_path / "s.h5"
    hdf5.write_stack(p, a, chunks=chunks, shuffle=shuffle, compression=1)
    with hdf5.H5File(p) as f:
        d = f[PATH]
        assert d.chunks == chunks
        np.testing.assert_array_equal(d.read(), a)
        n = shape[0]
        for lo, hi in [(0, 1), (1, n), (n // 2, n // 2 + 1), (n - 1, n), (2, 2)]:
            np.testing.assert_array_equal(d.read(lo, hi, threads=1), a[lo:hi])
            np.testing.assert_array_equal(d[lo:hi], a[lo:hi])
        np.testing.assert_array_equal(d[-1], a[-1])
        out = np.full((2,) + shape[1:], 7, np.uint16)
        assert d.read(1, 3, out=out) is out
        np.testing.assert_array_equal(out, a[1:3])
        with pytest.raises(ValueError):
            d.read(0, 2, out=np.empty((2,) + shape[1:], np.float32))
        with pytest.raises(IndexError):
            d.read(0, n + 1)


def test_uncompressed_chunked_contiguous_and_empty(tmp_path):
    a = _stack((4, 20, 30), "float32")
    hdf5.write_stack(tmp_path / "c.h5", a, compression=None, chunks=None)
    hdf5.write_stack(tmp_path / "u.h5", a, compression=None, chunks=(1, 7, 30))
    hdf5.write_stack(tmp_path / "e.h5", a[:0])
    with hdf5.H5File(tmp_path / "c.h5") as f:
        d = f[PATH]
        assert d.chunks is None and d.filters == []
        np.testing.assert_array_equal(d.read(1, 3), a[1:3])
        np.testing.assert_array_equal(d[2], a[2])
    with hdf5.H5File(tmp_path / "u.h5") as f:
        d = f[PATH]
        assert d.filters == [] and d.chunks == (1, 7, 30)
        np.testing.assert_array_equal(d.read(), a)
    with hdf5.H5File(tmp_path / "e.h5") as f:
        assert f[PATH].shape == (0, 20, 30) and f[PATH].read().shape == (0, 20, 30)


@pytest.mark.parametrize("n_frames,chunks,levels", [(3, (1, 4, 8), 1), (40, (1, 4, 8), 2), (70, (1, 1, 1), 3)])
def test_chunk_index_depth_and_file_structure(tmp_path, n_frames, chunks, levels):
    """64 children per node: 6 / 80 / 4480 chunks need a 1-, 2- and 3-level v1 B-tree. Walk the file by hand."""
    a = _stack((n_frames, 8, 8), "uint16", seed=n_frames)
    p = tmp_path / "s.h5"
    hdf5.write_stack(p, a, chunks=chunks)
    raw = p.read_bytes()
    assert raw[:8] == hdf5.SIGNATURE and raw[8] == 0 and raw[13] == 8 and raw[14] == 8
    base, _, eof, _ = struct.unpack_from("<QQQQ", raw, 24)
    assert base == 0 and eof == len(raw)
    with hdf5.H5File(p) as f:
        d = f[PATH]
        root = d._layout["btree"]
        assert raw[root:root + 4] == b"TREE" and raw[root + 4] == 1 and raw[root + 5] == levels - 1
        idx = d._chunk_index()
        n_chunks = int(np.prod([-(-s // c) for s, c in zip(a.shape, chunks)]))
        assert len(idx) == n_chunks and [r[0] for r in idx] == sorted(r[0] for r in idx)
        # every chunk inflates to exactly one chunk of elements and lies inside the file
        for offs, addr, nbytes, mask in idx[:: max(1, n_chunks // 50)]:
            assert mask == 0 and addr + nbytes <= len(raw)
            assert len(zlib.decompress(raw[addr:addr + nbytes])) == int(np.prod(chunks)) * 2
        np.testing.assert_array_equal(d.read(), a)
        np.testing.assert_array_equal(d.read(n_frames // 2, n_frames), a[n_frames // 2:])

    # node invariants: keys ascend, the closing key of a node is the first key of its right sibling, siblings link up
    keysz = 8 + 8 * 4

    def check(node, level):
        assert raw[node:node + 4] == b"TREE" and raw[node + 5] == level
        n = struct.unpack_from("<H", raw, node + 6)[0]
        assert 1 <= n <= 64
        left, right = struct.unpack_from("<QQ", raw, node + 8)
        keys = [struct.unpack_from("<II4Q", raw, node + 24 + i * (keysz + 8))[2:] for i in range(n + 1)]
        assert keys == sorted(keys) and all(k[3] == 0 for k in keys)
        kids = [struct.unpack_from("<Q", raw, node + 24 + i * (keysz + 8) + keysz)[0] for i in range(n)]
        if right != hdf5.UNDEF:
            assert struct.unpack_from("<II4Q", raw, right + 24)[2:] == keys[-1]
            assert struct.unpack_from("<Q", raw, right + 8)[0] == node
        if level:
            for k, kid in zip(keys, kids):
                assert struct.unpack_from("<II4Q", raw, kid + 24)[2:] == k
                check(kid, level - 1)

    check(root, levels - 1)


def test_group_structures_follow_the_symbol_table_layout(tmp_path):
    """Root -> entry_0000 -> measurement -> data through B-tree / heap / SNOD triples; NX_class attributes in place."""
    p = tmp_path / "s.h5"
    h5io.save_h5(_stack((2, 16, 16), "uint16"), p)
    raw = p.read_bytes()
    with hdf5.H5File(p) as f:
        assert f.keys("/") == ["entry_0000"] and f.keys("entry_0000") == ["measurement"]
        assert f.keys("entry_0000/measurement") == ["data"]
        for grp, cls in (("entry_0000", b"NXentry"), ("entry_0000/measurement", b"NXcollection")):
            attrs = [m for m in f._messages(f._resolve(grp)) if m[0] == hdf5.MSG_ATTRIBUTE]
            assert len(attrs) == 1 and b"NX_class\x00" in bytes(attrs[0][2]) and cls + b"\x00" in bytes(attrs[0][2])
        # the superblock's root entry caches the root group's B-tree and heap (cache type 1)
        name_off, header, cache, _, tree, heap = struct.unpack_from("<QQIIQQ", raw, 56)
        assert cache == 1 and raw[tree:tree + 4] == b"TREE" and raw[heap:heap + 4] == b"HEAP" and header == f._root
        seg_size, free_off, seg = struct.unpack_from("<QQQ", raw, heap + 8)
        nxt, free_size = struct.unpack_from("<QQ", raw, seg + free_off)
        assert nxt == 1 and free_off + free_size == seg_size                      # one free block closing the heap


def _v2_file(a: np.ndarray, *, big_endian=False) -> bytes:
    """A 'latest'-flavoured file assembled by hand: superblock v2, v2 object headers (sizes in 2 bytes, with times), a
    hard link message per group, dataspace v2, contiguous layout v3; checksums are left zero (readers may skip them)."""
    O = 8

    def ohdr(msgs):
        body = b"".join(struct.pack("<BHB", t, len(d), 0) + d for t, d in msgs)
        return b"OHDR" + bytes([2, 0x21]) + b"\x00" * 16 + struct.pack("<H", len(body)) + body + b"\x00" * 4

    def link(name, addr):
        nm = name.encode()
        return struct.pack("<BBB", 1, 0x10, 1) + bytes([len(nm)]) + nm + struct.pack("<Q", addr)

    dt = a.dtype.newbyteorder(">") if big_endian else a.dtype
    payload = a.astype(dt).tobytes()
    sb_len = 12 + 4 * O + 4
    link_info = struct.pack("<BBQQ", 0, 0, hdf5.UNDEF, hdf5.UNDEF)
    sizes = [len(ohdr([(hdf5.MSG_LINK_INFO, link_info), (hdf5.MSG_LINK, link(n, 0))])) for n in ("entry_0000", "measurement", "data")]
    a_root = sb_len
    a_entry, a_meas = a_root + sizes[0], a_root + sizes[0] + sizes[1]
    a_dset = a_meas + sizes[2]
    kind = a.dtype.kind
    if kind == "f":
        dtm = struct.pack("<BBBBIHHBBBBI", 0x11, 0x20 | big_endian, 31, 0, 4, 0, 32, 23, 8, 0, 23, 127)
    else:
        dtm = struct.pack("<BBBBIHH", 0x10, (0x08 if kind == "i" else 0) | big_endian, 0, 0, a.dtype.itemsize, 0, 8 * a.dtype.itemsize)
    space = struct.pack("<BBBB", 2, a.ndim, 0, 1) + struct.pack(f"<{a.ndim}Q", *a.shape)
    dset = ohdr([(hdf5.MSG_DATASPACE, space), (hdf5.MSG_DATATYPE, dtm),
                 (hdf5.MSG_LAYOUT, struct.pack("<BBQQ", 3, 1, 0, len(payload)))])
    a_raw = a_dset + len(dset)
    dset = ohdr([(hdf5.MSG_DATASPACE, space), (hdf5.MSG_DATATYPE, dtm),
                 (hdf5.MSG_LAYOUT, struct.pack("<BBQQ", 3, 1, a_raw, len(payload)))])
    out = hdf5.SIGNATURE + bytes([2, O, 8, 0]) + struct.pack("<QQQQ", 0, hdf5.UNDEF, a_raw + len(payload), a_root) + b"\x00" * 4
    out += ohdr([(hdf5.MSG_LINK_INFO, link_info), (hdf5.MSG_LINK, link("entry_0000", a_entry))])
    out += ohdr([(hdf5.MSG_LINK_INFO, link_info), (hdf5.MSG_LINK, link("measurement", a_meas))])
    out += ohdr([(hdf5.MSG_LINK_INFO, link_info), (hdf5.MSG_LINK, link("data", a_dset))])
    return out + dset + payload


@pytest.mark.parametrize("dtype,big", [("uint16", False), ("int32", True), ("float32", True)])
def test_reader_on_hand_assembled_new_style_file(tmp_path, dtype, big):
    a = _stack((3, 9, 12), dtype, seed=11)
    p = tmp_path / "v2.h5"
    p.write_bytes(_v2_file(a, big_endian=big))
    with hdf5.H5File(p) as f:
        d = f[PATH]
        assert d.shape == a.shape and d.dtype.newbyteorder("=") == a.dtype and (d.dtype.byteorder == ">") == big
        got = d.read()
        assert got.dtype.isnative
        np.testing.assert_array_equal(got, a)
        np.testing.assert_array_equal(d.read(1, 2), a[1:2])
    np.testing.assert_array_equal(h5io.read_h5(str(p), image_number=-1), a[-1])


def test_hand_assembled_chunked_file_with_fletcher_and_filter_mask(tmp_path):
    """Writer-independent check of the chunked read path: a file patched by hand so that one chunk skips deflate
    (filter mask bit) and the pipeline carries a Fletcher-32 stage (4 trailing bytes per chunk)."""
    a = _stack((2, 8, 8), "uint16", seed=5)
    p = tmp_path / "s.h5"
    hdf5.write_stack(p, a, chunks=(1, 8, 8), compression=None)            # chunked, no filter: raw chunks on disk
    raw = bytearray(p.read_bytes())
    with hdf5.H5File(p) as f:
        d = f[PATH]
        recs = d._chunk_index()
        root = d._layout["btree"]
    # re-describe the dataset: pipeline (deflate, fletcher32) via a continuation block appended to the file; chunk 0 is
    # stored deflated + checksum, chunk 1 raw + checksum with deflate masked out (bit 0)
    c0 = zlib.compress(a[0].tobytes(), 6) + b"\x00" * 4
    c1 = a[1].tobytes() + b"\x00" * 4
    a0, a1 = len(raw), len(raw) + len(c0)
    raw += c0 + c1
    keysz = 8 + 8 * 4
    struct.pack_into("<II", raw, root + 24, len(c0), 0)
    struct.pack_into("<Q", raw, root + 24 + keysz, a0)
    struct.pack_into("<II", raw, root + 24 + keysz + 8, len(c1), 1)
    struct.pack_into("<Q", raw, root + 24 + 2 * keysz + 8, a1)
    body = struct.pack("<BB6x", 1, 2)
    body += struct.pack("<HHHH", 1, 8, 1, 1) + b"deflate\x00" + struct.pack("<I4x", 6)
    body += struct.pack("<HHHH", 3, 0, 0, 0)
    cont = hdf5._msg(hdf5.MSG_FILTERS, body)
    a_cont = len(raw)
    raw += cont
    # a fresh dataset header at the end of the file: the old messages minus the fill value, plus a continuation message
    # that leads to the pipeline message; the SNOD entry of "data" is re-pointed to it
    with hdf5.H5File(p) as f:
        hdr = f._resolve(PATH)
    nmsg = struct.unpack_from("<H", raw, hdr + 2)[0]
    msgs = []
    q = hdr + 16
    for _ in range(nmsg):
        t, s, fl = struct.unpack_from("<HHB", raw, q)
        if t != hdf5.MSG_FILL:
            msgs.append(bytes(raw[q:q + 8 + s]))
        q += 8 + s
    msgs.append(hdf5._msg(hdf5.MSG_CONTINUATION, struct.pack("<QQ", a_cont, len(cont))))
    new_hdr = bytearray(hdf5._object_header(msgs))
    struct.pack_into("<H", new_hdr, 2, len(msgs) + 1)                      # the count covers the continuation block too
    a_new = len(raw) + (-len(raw) % 8)
    raw += b"\x00" * (a_new - len(raw)) + new_hdr
    snod = raw.rfind(b"SNOD")
    assert struct.unpack_from("<Q", raw, snod + 16)[0] == hdr
    struct.pack_into("<Q", raw, snod + 16, a_new)
    struct.pack_into("<Q", raw, 40, len(raw))
    p2 = tmp_path / "patched.h5"
    p2.write_bytes(bytes(raw))
    with hdf5.H5File(p2) as f:
        d = f[PATH]
        assert d.filters == [(1, [6]), (3, [])]
        np.testing.assert_array_equal(d.read(), a)
    assert len(recs) == 2


def test_unsupported_and_damaged_files_fail_loudly(tmp_path):
    p = tmp_path / "junk.h5"
    p.write_bytes(b"not hdf5 at all" * 100)
    with pytest.raises(OSError):
        hdf5.H5File(p)
    (tmp_path / "empty.h5").write_bytes(b"")
    with pytest.raises(OSError):
        hdf5.H5File(tmp_path / "empty.h5")
    a = _stack((2, 8, 8), "uint16")
    q = tmp_path / "s.h5"
    hdf5.write_stack(q, a, chunks=(1, 8, 8))
    raw = bytearray(q.read_bytes())
    with hdf5.H5File(q) as f:
        hdr = f._resolve(PATH)
    # an unknown filter id in the pipeline message
    i = raw.find(b"deflate\x00", hdr) - 8
    struct.pack_into("<H", raw, i, 32008)
    (tmp_path / "f.h5").write_bytes(bytes(raw))
    with hdf5.H5File(tmp_path / "f.h5") as f:
        with pytest.raises(hdf5.H5Unsupported, match="32008"):
            f[PATH].read()
    with pytest.raises(OSError, match="Failed to read"):
        h5io.read_h5(str(tmp_path / "f.h5"))
    # truncated in the middle of the raw data
    (tmp_path / "t.h5").write_bytes(q.read_bytes()[:-40])
    with pytest.raises(OSError):
        h5io.read_h5(str(tmp_path / "t.h5"))
    with pytest.raises(FileExistsError):
        hdf5.write_stack(q, a)


def test_read_h5_save_h5_contract(tmp_path):
    """Argument checks, error types and stacking rules of io/h5.py:64-142, :172-212."""
    a = _stack((4, 12, 10), "uint16", seed=1)
    p = str(tmp_path / "a.h5")
    h5io.save_h5(a, p)
    np.testing.assert_array_equal(h5io.read_h5(p), a)
    np.testing.assert_array_equal(h5io.read_h5(p, image_number=2), a[2])
    np.testing.assert_array_equal(h5io.read_h5(p, image_number=-1), a[3])
    for bad in (4, -5):
        with pytest.raises(ValueError, match="out of bounds"):
            h5io.read_h5(p, image_number=bad)
    h5io.save_h5(a[0], tmp_path / "img.dat")                                # suffix replaced by .h5
    assert (tmp_path / "img.h5").exists()
    img = str(tmp_path / "img.h5")
    np.testing.assert_array_equal(h5io.read_h5(img), a[0])
    with pytest.raises(ValueError, match="only valid for 3D"):
        h5io.read_h5(img, image_number=0)
    h5io.save_h5(a[1], tmp_path / "img2.hdf5")
    np.testing.assert_array_equal(h5io.read_h5([img, str(tmp_path / "img2.hdf5")]), a[:2])
    h5io.save_h5(a[2:], tmp_path / "b.h5")
    np.testing.assert_array_equal(h5io.read_h5((p, str(tmp_path / "b.h5"))), np.concatenate([a, a[2:]]))
    with pytest.raises(ValueError, match="Mixed dataset dimensionality"):
        h5io.read_h5([p, img])
    h5io.save_h5(np.zeros((5, 5), np.uint16), tmp_path / "small.h5")
    with pytest.raises(ValueError, match="Inconsistent image shapes"):
        h5io.read_h5([img, str(tmp_path / "small.h5")])
    with pytest.raises(ValueError, match="only supported when image_path is a single file"):
        h5io.read_h5([p], image_number=0)
    with pytest.raises(ValueError, match="empty"):
        h5io.read_h5([])
    with pytest.raises(TypeError):
        h5io.read_h5(7)
    with pytest.raises(TypeError):
        h5io.read_h5([7])
    with pytest.raises(FileNotFoundError):
        h5io.read_h5(str(tmp_path / "missing.h5"))
    hdf5.write_stack(tmp_path / "other.h5", a, dataset_path="entry_0000/instrument/data")
    with pytest.raises(KeyError, match="Dataset not found"):
        h5io.read_h5(str(tmp_path / "other.h5"))
    hdf5.write_stack(tmp_path / "4d.h5", np.zeros((2, 2, 3, 3), np.uint8))
    with pytest.raises(ValueError, match="Expected 2D or 3D"):
        h5io.read_h5(str(tmp_path / "4d.h5"))
    with pytest.raises(TypeError):
        h5io.save_h5([[1, 2]], tmp_path / "x.h5")
    with pytest.raises(ValueError, match="2D or 3D"):
        h5io.save_h5(np.zeros(3), tmp_path / "x.h5")
    with pytest.raises(OSError, match="Refusing to overwrite"):
        h5io.save_h5(a, p)
    with pytest.raises(OSError, match="directory does not exist"):
        h5io.save_h5(a, tmp_path / "nope" / "x.h5")


def test_read_image_write_image_dispatch(tmp_path):
    """io/rw.py: extension tables, the order of the argument checks, mean=True."""
    a = _stack((3, 8, 8), "uint16", seed=2)
    write_image(a, tmp_path / "s.hdf5")
    np.testing.assert_array_equal(read_image(str(tmp_path / "s.hdf5")), a)
    np.testing.assert_array_equal(read_image(str(tmp_path / "s.hdf5"), image_number=1), a[1])
    np.testing.assert_array_equal(read_image(str(tmp_path / "s.hdf5"), mean=True), a.mean(axis=0))
    write_image(a[0], tmp_path / "noext", file_extension="h5")
    np.testing.assert_array_equal(read_image(str(tmp_path / "noext.h5")), a[0])
    with pytest.raises(ValueError, match="Cannot infer"):
        read_image(str(tmp_path / "noext"))
    with pytest.raises(ValueError, match="Unsupported read extension"):
        read_image("x.png")
    with pytest.raises(ValueError, match="Mixed file extensions"):
        read_image(["a.h5", "b.tif"])
    with pytest.raises(ValueError, match="image_number is only supported for HDF5"):
        read_image("x.tif", image_number=0)
    with pytest.raises(FileNotFoundError):
        read_image("x.edf")
    with pytest.raises(ValueError, match="TIFF output is not built"):
        write_image(a, tmp_path / "x.tif")
    with pytest.raises(ValueError, match="Writing EDF is not supported"):
        write_image(a, tmp_path / "x.edf")
    with pytest.raises(ValueError, match="Unsupported write extension"):
        write_image(a, tmp_path / "x.png")
    with pytest.raises(TypeError):
        write_image("nope", tmp_path / "x.h5")
    with pytest.raises(TypeError):
        read_image(3)


@pytest.mark.parametrize("dtype,staged", [("uint16", "uint16"), ("float32", "float32"), ("float64", "float32"), ("int64", "float32")])
def test_block_reader_delivers_every_frame_once(tmp_path, dtype, staged):
    a = _stack((11, 16, 24), dtype, seed=4)
    p = tmp_path / "s.h5"
    hdf5.write_stack(p, a, chunks=(2, 5, 24))
    with H5StackSource(p, block_frames=4, pinned=False, decode_threads=2) as src:
        assert len(src) == 11 and src.frame_shape == (16, 24) and src.dtype == np.dtype(staged)
        seen = [(i, blk.copy()) for i, blk in src]
    assert [i for i, _ in seen] == [0, 4, 8] and [len(b) for _, b in seen] == [4, 4, 3]
    np.testing.assert_array_equal(np.concatenate([b for _, b in seen]), a.astype(staged))
    with H5StackSource(p, frames=(3, 9), block_frames=32, pinned=False) as src:
        (i, blk), = list(src)
        assert i == 3
        np.testing.assert_array_equal(blk, a[3:9].astype(staged))
    with pytest.raises(ValueError):
        H5StackSource(p, frames=(3, 12), pinned=False)
    hdf5.write_stack(tmp_path / "img.h5", a[0])
    with pytest.raises(ValueError, match="stack"):
        H5StackSource(tmp_path / "img.h5", pinned=False)


def test_block_reader_surfaces_decode_errors_and_stops_early(tmp_path):
    a = _stack((6, 8, 8), "uint16")
    p = tmp_path / "s.h5"
    hdf5.write_stack(p, a, chunks=(1, 8, 8))
    raw = bytearray(p.read_bytes())
    with hdf5.H5File(p) as f:
        rec = f[PATH]._chunk_index()[4]
    raw[rec[1]:rec[1] + 4] = b"\xff\xff\xff\xff"                           # frame 4 no longer inflates
    (tmp_path / "bad.h5").write_bytes(bytes(raw))
    with H5StackSource(tmp_path / "bad.h5", block_frames=2, pinned=False) as src:
        with pytest.raises(OSError, match="does not inflate"):
            for _ in src:
                pass
    with H5StackSource(p, block_frames=1, pinned=False) as src:            # abandoning the iterator joins the reader
        for i, _ in src:
            if i == 1:
                break


def test_concat_results_joins_every_leaf_along_frames():
    parts = [{"stats": {"mean": np.arange(3.0)}, "table": np.zeros((3, 5))},
             {"stats": {"mean": np.arange(2.0)}, "table": np.ones((2, 5))}]
    out = concat_results(parts)
    assert out["stats"]["mean"].tolist() == [0, 1, 2, 0, 1] and out["table"].shape == (5, 5)
    assert concat_results(parts[:1]) is parts[0]


def test_read_tiff_single_files_and_sequences(tmp_path):
    """io/tiff.py:19-70 through PIL (the reference's decoder): dtype and shape as stored, sequences stacked."""
    Image = pytest.importorskip("PIL.Image")
    from barc4dip_b200.io.tiff import read_tiff
    a = _stack((3, 20, 30), "uint16", seed=8)
    paths = []
    for i, fr in enumerate(a):
        paths.append(str(tmp_path / f"f_{i:04d}.tif"))
        Image.fromarray(fr).save(paths[-1])
    one = read_image(paths[1])
    assert one.dtype == np.uint16
    np.testing.assert_array_equal(one, a[1])
    np.testing.assert_array_equal(read_image(paths), a)
    np.testing.assert_array_equal(read_tiff(tuple(paths)), a)
    np.testing.assert_array_equal(read_image(paths, mean=True), a.mean(axis=0))
    f32 = _stack((20, 30), "float32", seed=9)
    Image.fromarray(f32).save(tmp_path / "f.tiff")
    got = read_image(str(tmp_path / "f.tiff"))
    assert got.dtype == np.float32
    np.testing.assert_array_equal(got, f32)
    Image.fromarray(a[0][:10]).save(tmp_path / "short.tif")
    with pytest.raises(ValueError, match="Inconsistent image shapes"):
        read_tiff([paths[0], str(tmp_path / "short.tif")])
    with pytest.raises(ValueError, match="empty"):
        read_tiff([])
    with pytest.raises(TypeError):
        read_tiff([1])
    with pytest.raises(TypeError):
        read_tiff(1)


def test_device_path_packs_raw_deflate_streams_on_the_host(tmp_path):
    """The host half of the device ingestion path (io.stream.DeviceInflater._pack) needs no GPU: stored chunks are read
    into the staging buffer so that the raw deflate stream (zlib minus header / Adler-32) starts 16-byte aligned."""
    from concurrent.futures import ThreadPoolExecutor

    from barc4dip_b200.io.stream import DeviceInflater
    a = _stack((4, 64, 64), "uint16", seed=6)
    p = tmp_path / "s.h5"
    hdf5.write_stack(p, a, chunks=(1, 16, 64))
    with hdf5.H5File(p) as f:
        d = f[PATH]
        inf = DeviceInflater.__new__(DeviceInflater)                        # no device: only the packer is exercised
        inf.dset, inf._pack_threads, inf._pool = d, 3, ThreadPoolExecutor(3)
        recs = d._chunk_index()
        pin = np.zeros(sum(16 + r[2] + (-r[2] % 16) for r in recs), np.uint8)
        offs, sizes, total = inf._pack(recs, pin)
        assert total == len(pin) and not (offs % 16).any()
        raw = b"".join(zlib.decompress(pin[o:o + s].tobytes(), -15) for o, s in zip(offs, sizes))
        np.testing.assert_array_equal(np.frombuffer(raw, np.uint16).reshape(4, 4, 16, 64).reshape(4, 64, 64), a)
        # a chunk that is not a zlib stream is refused before anything reaches the decompression engine
        bad = bytearray(p.read_bytes())
        bad[recs[5][1]] = 0x00
        (tmp_path / "bad.h5").write_bytes(bytes(bad))
    with hdf5.H5File(tmp_path / "bad.h5") as f:
        inf.dset = f[PATH]
        with pytest.raises(OSError, match="not a zlib stream"):
            inf._pack(inf.dset._chunk_index(), pin)
    inf.close()


def test_read_edf_against_arrays_the_reference_read(tmp_path, golden):
    """Every EDF case of oracle/make_golden_io.py: the bytes the reference's reader was given, replayed through
    barc4dip_b200.io.edf.read_edf -- element types, byte orders, header block sizes, CRLF, several images per file."""
    from barc4dip_b200.io.edf import read_edf
    g = golden("edf")
    names = sorted({k.split("/")[0] for k in g.files} - {"sequence"})
    assert len(names) >= 20
    for name in names:
        p = str(tmp_path / f"{name}.edf")
        with open(p, "wb") as fh:
            fh.write(g[f"{name}/file"].tobytes())
        idx = int(g[f"{name}/index"])
        for key, dt in (("float32", np.float32), ("float64", np.float64)):
            got = read_edf(p, index=idx, dtype=dt)
            want = g[f"{name}/{key}"]
            assert got.dtype == want.dtype and got.shape == want.shape, name
            np.testing.assert_array_equal(got, want, err_msg=name)
        if idx == 0 and g[f"{name}/float32"].ndim == 2:
            np.testing.assert_array_equal(read_image(p), g[f"{name}/float32"])
    seq = [str(tmp_path / f"{n}.edf") for n in g["sequence/names"]]
    np.testing.assert_array_equal(read_edf(seq), g["sequence/float32"])
    np.testing.assert_array_equal(read_image(seq), g["sequence/float32"])
    # error behaviour observed on the reference (io/edf.py:51-52, io/uti_EdfFile.py)
    one = str(tmp_path / "le_uint16.edf")
    with pytest.raises(ValueError, match="index must be >= 0"):
        read_edf(one, index=-1)
    with pytest.raises(ValueError, match="Index out of limit"):
        read_edf(one, index=1)
    (tmp_path / "junk.edf").write_bytes(b"hello world\n" * 10)
    with pytest.raises(ValueError, match="Index out of limit"):
        read_edf(str(tmp_path / "junk.edf"))
    raw = g["le_uint16/file"].tobytes()
    (tmp_path / "trunc.edf").write_bytes(raw[:-5])
    with pytest.raises(ValueError):
        read_edf(str(tmp_path / "trunc.edf"))
    (tmp_path / "unk.edf").write_bytes(raw.replace(b"UnsignedShort", b"ComplexFloat "))
    with pytest.raises(TypeError, match="unknown EdfType"):
        read_edf(str(tmp_path / "unk.edf"))
    with pytest.raises(ValueError, match="Expected a 2D EDF image"):
        read_edf([str(tmp_path / "three_dims.edf")])
    with pytest.raises(ValueError, match="Inconsistent image shapes"):
        read_edf([one, str(tmp_path / "be_uint16.edf")])
    with pytest.raises(FileNotFoundError):
        read_edf(str(tmp_path / "missing.edf"))
    with pytest.raises(TypeError):
        read_edf([3])
    with pytest.raises(ValueError, match="empty"):
        read_edf([])


def test_read_tiff_against_arrays_the_reference_read(tmp_path, golden):
    pytest.importorskip("PIL.Image")
    from barc4dip_b200.io.tiff import read_tiff
    g = golden("tiff")
    names = sorted({k.split("/")[0] for k in g.files} - {"sequence"})
    for name in names:
        p = str(tmp_path / f"{name}.tif")
        with open(p, "wb") as fh:
            fh.write(g[f"{name}/file"].tobytes())
        got = read_tiff(p)
        assert got.dtype == g[f"{name}/array"].dtype
        np.testing.assert_array_equal(got, g[f"{name}/array"], err_msg=name)
    seq = [str(tmp_path / f"{n}.tif") for n in g["sequence/names"]]
    np.testing.assert_array_equal(read_image(seq), g["sequence/array"])


_SHARDED_WORKER = r'''
import os, sys
sys.path.insert(0, {root!r})
import numpy as np, torch.distributed as dist
from barc4dip_b200 import parallel
from barc4dip_b200.io import hdf5, stream
dist.init_process_group("gloo", rank=int(os.environ["RANK"]), world_size=int(os.environ["WORLD_SIZE"]))
rank, world = parallel.dist_info()

# the GPU analysis replaced by a host stand-in with the same contract (per-frame leaves for frames [lo, hi) of the FILE):
# what is under test is the host logic around it -- every rank reads its own range and the reference, tables are gathered
calls = []
def fake_analyze(path, *, reference=None, frames=None, **kw):
    calls.append(frames)
    with hdf5.H5File(path) as f:
        blk = f["entry_0000/measurement/data"].read(*frames).astype(np.float64)
    return {{"stats": {{"mean": blk.mean(axis=(1, 2))}}, "table": blk.reshape(len(blk), -1)[:, :3],
            "tracking": {{"dx": (blk - np.asarray(reference, np.float64)).sum(axis=(1, 2))}}}}
stream.analyze_h5_stack = fake_analyze

for T in (7, 1):
    path = {tmp!r} + f"/s{{T}}.h5"
    with hdf5.H5File(path) as f:
        full = f["entry_0000/measurement/data"].read().astype(np.float64)
    out = parallel.analyze_h5_stack_sharded(path, block_frames=2)
    lo, hi = parallel.frame_range(T, rank, world)
    assert out["frame_range"] == (lo, hi) and calls[-1] == ((lo, hi) if hi > lo else (0, 1))
    assert np.array_equal(out["stats"]["mean"], full.mean(axis=(1, 2)))
    assert np.array_equal(out["table"], full.reshape(T, -1)[:, :3])
    assert np.array_equal(out["tracking"]["dx"], (full - full[0]).sum(axis=(1, 2)))
dist.destroy_process_group()
print("ok", rank)
'''


def test_file_sharded_entry_on_two_gloo_ranks(tmp_path):
    """parallel.analyze_h5_stack_sharded: frame ranges per rank, the reference read by every rank, per-frame tables
    all-gathered in frame order, a rank without frames (T = 1 on two ranks) still joining the collectives."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for T in (7, 1):
        hdf5.write_stack(tmp_path / f"s{T}.h5", _stack((T, 12, 10), "uint16", seed=T), chunks=(2, 5, 10))
    script = tmp_path / "worker.py"
    script.write_text(_SHARDED_WORKER.format(root=root, tmp=str(tmp_path)))
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT="29573")
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    for p in procs:
        out, _ = p.communicate(timeout=120)
        assert p.returncode == 0 and "ok" in out, out


def _latest_file(a: np.ndarray, chunks, *, filtered: bool, page_bits: int, drop_last_page: bool = False) -> bytes:
    """A file as `libver="latest"` lays out a fixed-shape chunked dataset, assembled by hand: superblock v2, v2 object
    headers with link messages, dataspace v2, filter pipeline v2, layout v4 with a fixed-array chunk index (FAHD + FADB,
    paged when the chunk count exceeds 2**page_bits). Checksums are left zero."""
    rank, es = a.ndim, a.dtype.itemsize

    def ohdr(msgs):
        body = b"".join(struct.pack("<BHB", t, len(d), 0) + d for t, d in msgs)
        return b"OHDR" + bytes([2, 0x01]) + struct.pack("<H", len(body)) + body + b"\x00" * 4      # chunk-0 size in 2 bytes

    def link(name, addr):
        nm = name.encode()
        return struct.pack("<BB", 1, 0) + bytes([len(nm)]) + nm + struct.pack("<Q", addr)            # hard link, no flags

    grid = [-(-s // c) for s, c in zip(a.shape, chunks)]
    blobs = []
    for idx in np.ndindex(*grid):
        blk = np.zeros(chunks, a.dtype)
        sl = tuple(slice(i * c, min((i + 1) * c, s)) for i, c, s in zip(idx, chunks, a.shape))
        part = a[sl]
        blk[tuple(slice(0, n) for n in part.shape)] = part
        blobs.append(zlib.compress(blk.tobytes(), 4) if filtered else blk.tobytes())
    count = len(blobs)
    nlen = 3                                                         # bytes of the stored-size field of a filtered element
    esize = 8 + (nlen + 4 if filtered else 0)
    per_page = 1 << page_bits
    paged = count > per_page
    npages = -(-count // per_page) if paged else 0

    space = struct.pack("<BBBB", 2, rank, 0, 1) + struct.pack(f"<{rank}Q", *a.shape)
    kind = a.dtype.kind
    dtm = struct.pack("<BBBBIHHBBBBI", 0x11, 0x20, 31, 0, 4, 0, 32, 23, 8, 0, 23, 127) if kind == "f" else \
        struct.pack("<BBBBIHH", 0x10, 0x08 if kind == "i" else 0, 0, 0, es, 0, 8 * es)
    pipeline = struct.pack("<BB", 2, 1) + struct.pack("<HHHI", 1, 1, 1, 4)
    def layout(addr):
        return struct.pack("<BBBBB", 4, 2, 0, rank + 1, 4) + struct.pack(f"<{rank + 1}I", *chunks, es) + bytes([3, page_bits]) + struct.pack("<Q", addr)
    def dset_header(addr):
        msgs = [(hdf5.MSG_DATASPACE, space), (hdf5.MSG_DATATYPE, dtm)] + ([(hdf5.MSG_FILTERS, pipeline)] if filtered else [])
        return ohdr(msgs + [(hdf5.MSG_LAYOUT, layout(addr))])

    sb_len = 12 + 4 * 8 + 4
    names = ("entry_0000", "measurement", "data")
    sizes = [len(ohdr([(hdf5.MSG_LINK, link(n, 0))])) for n in names]
    a_root = sb_len
    a_entry, a_meas = a_root + sizes[0], a_root + sizes[0] + sizes[1]
    a_dset = a_meas + sizes[2]
    a_fahd = a_dset + len(dset_header(0))
    fahd_len = 8 + 8 + 8 + 4
    a_fadb = a_fahd + fahd_len
    prefix = 6 + 8 + ((npages + 7) // 8 if paged else count * esize) + 4
    pages_len = sum(per_page * esize + 4 for _ in range(npages))
    a_data = a_fadb + prefix + pages_len
    addrs, pos = [], a_data
    for b in blobs:
        addrs.append(pos)
        pos += len(b)

    def element(i):
        e = struct.pack("<Q", addrs[i])
        return e + (len(blobs[i]).to_bytes(nlen, "little") + struct.pack("<I", 0) if filtered else b"")

    fahd = b"FAHD" + bytes([0, int(filtered), esize, page_bits]) + struct.pack("<QQ", count, a_fadb) + b"\x00" * 4
    fadb = b"FADB" + bytes([0, int(filtered)]) + struct.pack("<Q", a_fahd)
    if not paged:
        fadb += b"".join(element(i) for i in range(count)) + b"\x00" * 4
    else:
        bits = bytearray((npages + 7) // 8)
        for pg in range(npages - (1 if drop_last_page else 0)):
            bits[pg // 8] |= 0x80 >> (pg % 8)
        fadb += bytes(bits) + b"\x00" * 4
        for pg in range(npages):
            els = b"".join(element(i) for i in range(pg * per_page, min(count, (pg + 1) * per_page)))
            if drop_last_page and pg == npages - 1:
                els = b"\xff" * len(els)                              # never written: whatever the file held there
            fadb += els + b"\x00" * (per_page * esize - len(els)) + b"\x00" * 4
    assert len(fahd) == fahd_len and len(fadb) == prefix + pages_len
    out = hdf5.SIGNATURE + bytes([2, 8, 8, 0]) + struct.pack("<QQQQ", 0, hdf5.UNDEF, pos, a_root) + b"\x00" * 4
    out += ohdr([(hdf5.MSG_LINK, link("entry_0000", a_entry))]) + ohdr([(hdf5.MSG_LINK, link("measurement", a_meas))])
    out += ohdr([(hdf5.MSG_LINK, link("data", a_dset))]) + dset_header(a_fahd) + fahd + fadb + b"".join(blobs)
    assert len(out) == pos
    return out


@pytest.mark.parametrize("filtered,page_bits", [(False, 10), (True, 10), (True, 2), (False, 1)])
def test_reader_on_hand_assembled_fixed_array_index(tmp_path, filtered, page_bits):
    """Layout v4 + fixed-array chunk index as libver='latest' writes it for fixed-shape datasets: plain and filtered
    elements, unpaged and paged data blocks (2**page_bits elements per page), ragged edge chunks."""
    a = _stack((5, 20, 18), "uint16", seed=14)
    chunks = (2, 8, 18)                                              # 3 x 3 x 1 = 9 chunks
    p = tmp_path / "latest.h5"
    p.write_bytes(_latest_file(a, chunks, filtered=filtered, page_bits=page_bits))
    with hdf5.H5File(p) as f:
        d = f[PATH]
        assert d.shape == a.shape and d.chunks == chunks and len(d._chunk_index()) == 9
        assert d.filters == ([(1, [4])] if filtered else [])
        np.testing.assert_array_equal(d.read(), a)
        np.testing.assert_array_equal(d.read(3, 5), a[3:5])
    np.testing.assert_array_equal(h5io.read_h5(str(p), image_number=2), a[2])
    if page_bits == 2:
        # a page whose bit is clear was never written: its chunks (the 9th of 9, frames 4.., rows 16..) read as zeros
        p.write_bytes(_latest_file(a, chunks, filtered=filtered, page_bits=page_bits, drop_last_page=True))
        want = a.copy()
        want[4:, 16:, :] = 0
        with hdf5.H5File(p) as f:
            assert len(f[PATH]._chunk_index()) == 8
            np.testing.assert_array_equal(f[PATH].read(), want)
