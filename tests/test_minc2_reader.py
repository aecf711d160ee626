"""The library's MINC 2 / HDF5 reader (rminc_b200/csrc/minc2_reader.cpp) on the CPU.

Pins: (1) a file written by the real HDF5 library -- scipy's MATLAB-7.3 fixture, a 512-byte user block + superblock 0 +
symbol-table root group + version-1 object header + version-2 contiguous layout + version-1 string attribute -- copied to
tests/golden/libhdf5_written_matlab73.mat (4 KB, BSD-licensed test data of scipy.io.matlab); (2) MINC-2.0-shaped files laid
out by the test-side writer tests/h5mini.py (every layout class, chunk B-trees of one and several levels, deflate,
shuffle, big-endian voxels, several symbol nodes per group).  What stays unpinned is stated in the reader's header:
libminc's stored -> real arithmetic (restated, like oracle/ does for the kernels)."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import h5mini  # noqa: E402

from rminc_b200 import _lib, api, minc2  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "libhdf5_written_matlab73.mat")


def h5_dataset(path, name, n_threads=2):
    lib = _lib.load()
    rank, es, tc = C.c_int32(), C.c_int32(), C.c_int32()
    dims = (C.c_int64 * 8)()
    err = _lib.errbuf()
    _lib.check(lib.rmg_h5_read_dataset(path.encode(), name.encode(), None, 0, C.byref(rank), dims, C.byref(es), C.byref(tc), 1,
                                       err, len(err)), err)
    shape = tuple(dims[i] for i in range(rank.value))
    buf = np.empty(int(np.prod(shape, dtype=np.int64)) * es.value, np.uint8)
    _lib.check(lib.rmg_h5_read_dataset(path.encode(), name.encode(), buf.ctypes.data, buf.nbytes, None, None, None, None,
                                       n_threads, err, len(err)), err)
    return shape, es.value, tc.value, buf


def h5_attribute(path, obj, name):
    lib = _lib.load()
    txt = C.create_string_buffer(256)
    nums = np.zeros(16)
    cnt = C.c_int64()
    err = _lib.errbuf()
    _lib.check(lib.rmg_h5_read_attribute(path.encode(), obj.encode(), name.encode(), nums.ctypes.data, nums.size, C.byref(cnt), txt,
                                         len(txt), err, len(err)), err)
    return txt.value.decode() if txt.value else nums[:cnt.value]


def test_reads_a_file_written_by_the_hdf5_library():
    shape, es, tc, buf = h5_dataset(GOLDEN, "testdouble")
    assert shape == (9, 1) and es == 8 and tc == 3
    assert np.array_equal(buf.view(np.float64), np.arange(9) * np.pi / 4)        # MATLAB's 0:pi/4:2*pi, bit for bit
    assert h5_attribute(GOLDEN, "testdouble", "MATLAB_class") == "double"
    assert minc2.is_minc2(GOLDEN)                                                 # an HDF5 file behind a user block ...
    with pytest.raises(_lib.RmincGpuError) as e:                                  # ... but not a MINC volume
        minc2.Minc2Volume(GOLDEN)
    assert e.value.code == _lib.RMG_ERR_IO and "not a MINC 2 volume" in str(e.value)


def expected_real(vol, lo, hi, vr):
    lo, hi = np.atleast_1d(lo), np.atleast_1d(hi)
    shape = (lo.size,) + (1,) * (vol.ndim - 1)
    scale = (hi - lo) / (vr[1] - vr[0])
    return ((vol.astype(np.float64) - vr[0]) * scale.reshape(shape) + lo.reshape(shape)).reshape(-1)


CASES = [
    # dtype, shape, chunks, kwargs of the writer
    (np.uint16, (7, 9, 11), (1, 4, 8), dict(compress=4)),                        # ragged edge chunks
    (np.uint16, (5, 16, 16), (1, 16, 16), dict(compress=9, shuffle=True)),       # libminc's slice chunks
    (np.uint16, (6, 10, 12), None, dict()),                                      # contiguous
    (np.uint16, (4, 6, 5), (2, 3, 5), dict(compress=None)),                      # chunked, no filter
    (np.uint16, (3, 8, 8), (1, 8, 8), dict(compress=1, big_endian=True)),
    (np.int16, (4, 7, 9), (2, 7, 9), dict(compress=4)),
    (np.uint8, (5, 6, 7), (5, 6, 7), dict(compress=4)),
    (np.int32, (2, 3, 4), None, dict()),
    (np.float32, (4, 5, 6), (1, 5, 6), dict(compress=4)),
    (np.float64, (3, 4, 5), (3, 2, 5), dict(compress=4)),
]


@pytest.mark.parametrize("dtype,shape,chunks,kw", CASES)
def test_round_trip_of_minc2_volumes(tmp_path, dtype, shape, chunks, kw):
    rng = np.random.default_rng(abs(hash((str(dtype), shape))) % 2**32)
    dt = np.dtype(dtype)
    if dt.kind == "f":
        vol = rng.normal(size=shape).astype(dt)
        vr = None
    else:
        info = np.iinfo(dt)
        vol = rng.integers(info.min, int(info.max) + 1, size=shape, dtype=dt)
        vr = (float(info.min), float(info.max))
    lo = rng.normal(size=shape[0])
    hi = lo + rng.uniform(0.5, 2.0, size=shape[0])
    path = str(tmp_path / "v.mnc")
    h5mini.write_minc2(path, vol, image_min=lo, image_max=hi, valid_range=vr, steps=(0.5, -0.25, 0.125), starts=(-1.0, 2.0, -3.0),
                       chunks=chunks, **kw)
    with minc2.Minc2Volume(path) as v:
        assert v.sizes == shape and v.dimnames == ("zspace", "yspace", "xspace")
        assert v.steps == (0.5, -0.25, 0.125) and v.starts == (-1.0, 2.0, -3.0)
        assert v.direction_cosines == ((0, 0, 1), (0, 1, 0), (1, 0, 0))
        assert v.dtype == dt and v.nslices == shape[0] and v.slice_len == shape[1] * shape[2]
        assert v.voxel_volume() == 0.5 * 0.25 * 0.125
        for nt in (1, 3):
            assert np.array_equal(v.stored(n_threads=nt), vol)
        a, b = v.ranges()
        assert np.array_equal(a, lo) and np.array_equal(b, hi)
        want = vol.astype(np.float64).reshape(-1) if dt.kind == "f" else expected_real(vol, lo, hi, vr)
        assert np.array_equal(v.real(), want)
    # what mincLm's default reader makes of it
    got = api.default_reader(path)
    if dt in (np.dtype(np.uint16), np.dtype(np.float32)):
        assert isinstance(got, api.StoredVolume) and np.array_equal(got.real(), want)
    else:
        assert np.array_equal(got, want)


def test_global_range_many_chunks_and_many_group_members(tmp_path):
    rng = np.random.default_rng(5)
    vol = rng.integers(0, 4096, size=(12, 20, 24), dtype=np.uint16)
    path = str(tmp_path / "g.mnc")
    tree = h5mini.minc2_tree(vol, image_min=-2.5, image_max=7.25, valid_range=(0, 4095), chunks=(1, 5, 6), compress=4)
    for k in range(9):                                  # > 2 * leaf_k members: several symbol nodes under one B-tree node
        tree.children["minc-2.0"].children["info"].children[f"note{k}"] = h5mini.Dataset(np.arange(k + 1, dtype=np.int32), compact=True)
    h5mini.write_h5(path, tree, leaf_k=2, chunk_k=2)   # 12 * 4 * 4 = 192 chunks in 4-entry nodes: a 4-level chunk B-tree
    with minc2.Minc2Volume(path) as v:
        assert v.nslices == 1 and v.slice_len == vol.size and v.valid_range == (0.0, 4095.0)
        assert np.array_equal(v.stored(n_threads=4), vol)
        assert np.array_equal(v.real(), expected_real(vol, -2.5, 7.25, (0.0, 4095.0)))
    for k in range(9):
        shape, es, tc, buf = h5_dataset(path, f"/minc-2.0/info/note{k}")
        assert shape == (k + 1,) and tc == 2 and np.array_equal(buf.view(np.int32), np.arange(k + 1))
    assert h5_attribute(path, "minc-2.0/image/0/image", "dimorder") == "zspace,yspace,xspace"
    assert np.array_equal(h5_attribute(path, "minc-2.0/image/0/image", "valid_range"), [0.0, 4095.0])
    assert np.array_equal(h5_attribute(path, "minc-2.0/dimensions/xspace", "length"), [24.0])


def test_header_continuation_blocks_user_block_and_max_dims(tmp_path):
    """What libhdf5 does to objects that grow: attributes added after creation live in an object-header continuation block
    (message 0x0010), dataspaces carry their maximum dimensions, and a file may start with a user block."""
    rng = np.random.default_rng(21)
    vol = rng.integers(0, 65536, size=(5, 8, 6), dtype=np.uint16)
    lo = rng.normal(size=5)
    tree = h5mini.minc2_tree(vol, image_min=lo, image_max=lo + 1.25, valid_range=(0, 65535), chunks=(1, 8, 6), steps=(2.0, 1.0, 0.5))
    path = str(tmp_path / "grown.mnc")
    h5mini.write_h5(path, tree, split_headers=True, max_dims=True, user_block=1024)
    assert minc2.is_minc2(path)
    with minc2.Minc2Volume(path) as v:
        assert v.sizes == (5, 8, 6) and v.steps == (2.0, 1.0, 0.5) and v.valid_range == (0.0, 65535.0) and v.nslices == 5
        assert np.array_equal(v.stored(), vol)
        assert np.array_equal(v.real(), expected_real(vol, lo, lo + 1.25, (0.0, 65535.0)))
    assert h5_attribute(path, "minc-2.0/image/0/image", "dimorder") == "zspace,yspace,xspace"       # lives in the continuation
    assert h5_attribute(path, "minc-2.0", "history") == "h5mini"


def test_read_table_decodes_a_study_in_parallel(tmp_path):
    rng = np.random.default_rng(9)
    shape, n = (6, 10, 9), 7
    paths, vols, los, his = [], [], [], []
    for j in range(n):
        vol = rng.integers(0, 65536, size=shape, dtype=np.uint16)
        lo = rng.normal(size=shape[0]) if j != 3 else np.float64(-1.0)          # one file with a single global range
        hi = lo + rng.uniform(0.5, 2.0, size=np.shape(lo))
        p = str(tmp_path / f"s{j}.mnc")
        h5mini.write_minc2(p, vol, image_min=lo, image_max=hi, valid_range=(0, 65535), chunks=(1, 10, 9))
        paths.append(p); vols.append(vol); los.append(np.broadcast_to(lo, (shape[0],))); his.append(np.broadcast_to(hi, (shape[0],)))
    # the single-range file comes first in neither order: nslices is taken from the first file
    U, lo, hi, vr, info = minc2.read_table(paths, n_threads=4)
    assert U.dtype == np.uint16 and U.flags.f_contiguous and U.shape == (vol.size, n) and vr == (0.0, 65535.0)
    for j in range(n):
        assert np.array_equal(U[:, j], vols[j].reshape(-1))
        assert np.array_equal(lo[:, j], los[j]) and np.array_equal(hi[:, j], his[j])
    Y, lo2, hi2, _, _ = minc2.read_table(paths, dtype=np.float64, n_threads=3)
    assert lo2 is None and Y.dtype == np.float64
    for j in range(n):
        assert np.array_equal(Y[:, j], expected_real(vols[j], los[j], his[j], (0.0, 65535.0)))
    # the staging mincLm does: the stored matrix and its ranges as scale / offset
    staged = api._minc2_stage(paths)
    assert staged[0] == "stored" and staged[5] == shape[1] * shape[2] and staged[4] == 0.0
    assert np.array_equal(api._expand_stored(*staged[1:]), Y)
    # errors: a volume of another size, another type, a missing file (the reference's message)
    odd = str(tmp_path / "odd.mnc")
    h5mini.write_minc2(odd, vols[0][:5], image_min=0.0, image_max=1.0, valid_range=(0, 65535))
    with pytest.raises(_lib.RmincGpuError, match="volume sizes differ"):
        minc2.read_table(paths + [odd])
    flt = str(tmp_path / "flt.mnc")
    h5mini.write_minc2(flt, vols[0].astype(np.float32))
    with pytest.raises(_lib.RmincGpuError, match="stored type differs"):
        minc2.read_table(paths + [flt])
    with pytest.raises(FileNotFoundError, match="Error opening input file: .*nowhere.mnc"):
        minc2.read_table(paths + [str(tmp_path / "nowhere.mnc")])
    with pytest.raises(FileNotFoundError, match="Error opening input file"):
        api.default_reader(str(tmp_path / "nowhere.mnc"))


def test_files_that_are_not_minc2(tmp_path):
    junk = tmp_path / "junk.mnc"
    junk.write_bytes(b"CDF\x01" + bytes(4096))          # a MINC 1 (NetCDF classic) header
    assert not minc2.is_minc2(junk)
    with pytest.raises(_lib.RmincGpuError, match="no HDF5 signature") as e:
        minc2.Minc2Volume(junk)
    assert e.value.code == _lib.RMG_ERR_IO
    with pytest.raises(ValueError, match="not a MINC 2"):
        api.default_reader(str(junk))
    tiny = tmp_path / "tiny.mnc"
    tiny.write_bytes(b"\x89HDF\r\n\x1a\n")
    with pytest.raises(_lib.RmincGpuError):
        minc2.Minc2Volume(tiny)


def test_damaged_files_fail_cleanly(tmp_path):
    """Truncations and random byte damage must end in an error code (or in data), never in a crash or a hang."""
    rng = np.random.default_rng(11)
    vol = rng.integers(0, 65536, size=(6, 12, 10), dtype=np.uint16)
    good = tmp_path / "good.mnc"
    h5mini.write_minc2(str(good), vol, image_min=np.zeros(6), image_max=np.ones(6), valid_range=(0, 65535), chunks=(1, 12, 10))
    blob = good.read_bytes()
    bad = tmp_path / "bad.mnc"
    outcomes = {"ok": 0, "error": 0}
    for cut in (100, 500, len(blob) // 2, len(blob) - 50):
        bad.write_bytes(blob[:cut])
        with pytest.raises(_lib.RmincGpuError):
            with minc2.Minc2Volume(bad) as v:
                v.stored()
    for trial in range(300):
        b = bytearray(blob)
        for _ in range(int(rng.integers(1, 6))):
            pos = int(rng.integers(0, len(b)))
            b[pos] = int(rng.integers(0, 256))
        bad.write_bytes(bytes(b))
        try:
            with minc2.Minc2Volume(bad) as v:
                if v.nvoxels <= 1 << 22:
                    v.stored(n_threads=2)
                    v.ranges()
            outcomes["ok"] += 1
        except (_lib.RmincGpuError, FileNotFoundError, MemoryError, ValueError, KeyError):
            outcomes["error"] += 1
    assert outcomes["ok"] + outcomes["error"] == 300 and outcomes["error"] > 0


def test_like_volume_and_mask_from_minc2_files(tmp_path):
    vol = np.zeros((4, 5, 6), dtype=np.uint8)
    vol[1:3, 1:4, 2:5] = 1
    path = str(tmp_path / "mask.mnc")
    h5mini.write_minc2(path, vol, valid_range=(0, 255), chunks=(4, 5, 6))        # a label volume: no real range
    assert api.volume_dims(path) == (4, 5, 6)
    m = api._stage_mask(path, vol.size, api.default_reader)
    assert m.dtype == np.float64 and np.array_equal(m, vol.reshape(-1).astype(np.float64))


def test_random_shapes_chunkings_and_filters(tmp_path):
    """Property check over the writer's parameter space: any shape, chunking (ragged edges included), filter set, byte order
    and B-tree fan-out reads back bit for bit, through one thread or several."""
    rng = np.random.default_rng(2026)
    dtypes = [np.uint8, np.int8, np.uint16, np.int16, np.uint32, np.int32, np.float32, np.float64]
    for trial in range(60):
        dt = np.dtype(dtypes[int(rng.integers(len(dtypes)))])
        shape = tuple(int(rng.integers(1, 14)) for _ in range(3))
        if dt.kind == "f":
            vol = rng.normal(size=shape).astype(dt)
            vr = None
        else:
            info = np.iinfo(dt)
            vol = rng.integers(info.min, int(info.max) + 1, size=shape, dtype=dt)
            vr = (float(info.min), float(info.max))
        mode = int(rng.integers(3))
        chunks = None if mode == 0 else tuple(int(rng.integers(1, s + 1)) for s in shape)
        kw = dict(compress=None if mode == 1 else int(rng.integers(1, 10)), shuffle=bool(rng.integers(2)) and mode == 2,
                  big_endian=bool(rng.integers(2)) and dt.itemsize > 1) if chunks else dict(big_endian=bool(rng.integers(2)) and dt.itemsize > 1)
        per_slice = bool(rng.integers(2))
        lo = rng.normal(size=shape[0]) if per_slice else np.float64(rng.normal())
        hi = lo + rng.uniform(0.5, 2.0, size=np.shape(lo))
        path = str(tmp_path / f"t{trial}.mnc")
        tree = h5mini.minc2_tree(vol, image_min=lo, image_max=hi, valid_range=vr, chunks=chunks, **kw)
        h5mini.write_h5(path, tree, leaf_k=int(rng.integers(1, 5)), chunk_k=int(rng.integers(1, 5)), split_headers=bool(rng.integers(2)),
                        max_dims=bool(rng.integers(2)), user_block=int(rng.choice([0, 512, 2048])))
        with minc2.Minc2Volume(path) as v:
            assert v.sizes == shape and v.dtype == dt, (trial, shape, dt)
            assert np.array_equal(v.stored(n_threads=int(rng.integers(1, 5))), vol), (trial, shape, dt, chunks, kw)
            a, b = v.ranges()
            assert np.array_equal(a, np.atleast_1d(lo)) and np.array_equal(b, np.atleast_1d(hi))
            if dt.kind != "f":
                assert np.array_equal(v.real(), expected_real(vol, lo, hi, vr)), (trial, shape, dt)
