"""The MINC 2 reader's own DEFLATE decoder (rminc_b200/csrc/inflate_fast.hpp, through rmg_zlib_uncompress) against zlib:
every stream zlib's compressor produces -- all levels, strategies (default, filtered, Huffman-only, RLE, fixed codes), window
sizes, multi-block streams with sync / full flushes and stored blocks, on noise, volume-like data, runs, text and empty input --
must decode to the same bytes; truncated and damaged streams must be rejected (or decode, never crash); an output buffer one
byte short is an error.  zlib is the oracle here: it ships with Python."""
import ctypes as C
import zlib

import numpy as np
import pytest

from rminc_b200 import _lib


def mine(comp, n, cap=None, use_zlib=0):
    cap = n if cap is None else cap
    out = np.empty(max(cap, 1), np.uint8)
    r = _lib.load().rmg_zlib_uncompress(C.c_char_p(comp), len(comp), out.ctypes.data, cap, use_zlib)
    return r, out[:max(r, 0)].tobytes()


def datasets():
    rng = np.random.default_rng(0)
    yield b""
    yield b"a"
    yield b"abc" * 1000
    yield bytes(70000)
    yield rng.integers(0, 256, 60000, dtype=np.uint8).tobytes()
    yield np.clip(np.rint(rng.normal(30000, 400, 80000)), 0, 65535).astype(np.uint16).tobytes()        # a uint16 volume
    yield np.cumsum(rng.normal(0, 3, 90000)).astype(np.int16).tobytes()                                 # a smooth one
    yield (np.arange(100000) % 251).astype(np.uint8).tobytes()
    yield open(__file__, "rb").read() * 5
    yield rng.integers(0, 4, 100000, dtype=np.uint8).tobytes()
    for n in (1, 2, 7, 8, 9, 255, 256, 257, 258, 259, 65535, 65536, 65537):
        yield rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        yield bytes([7]) * n


def test_every_stream_zlib_writes_decodes_to_the_same_bytes():
    count = 0
    for data in datasets():
        for level in (0, 1, 4, 6, 9):
            for strat in (zlib.Z_DEFAULT_STRATEGY, zlib.Z_FILTERED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE, zlib.Z_FIXED):
                for wbits in (15, 9):
                    co = zlib.compressobj(level, zlib.DEFLATED, wbits, 9 if level else 8, strat)
                    comp = co.compress(data) + co.flush()
                    r, out = mine(comp, len(data))
                    assert r == len(data) and out == data, (len(data), level, strat, wbits, r)
                    assert mine(comp, len(data), len(data) + 100)[0] == len(data)
                    assert mine(comp, len(data), use_zlib=1) == (r, out)                   # the twin agrees
                    if data:
                        assert mine(comp, len(data), len(data) - 1)[0] == -1               # output one byte short
                    count += 1
    assert count > 1500


def test_multi_block_streams_with_flushes():
    rng = np.random.default_rng(1)
    for trial in range(120):
        co = zlib.compressobj(int(rng.integers(0, 10)))
        parts, comp = [], b""
        for _ in range(int(rng.integers(1, 8))):
            d = rng.integers(0, int(rng.integers(2, 256)), int(rng.integers(0, 20000)), dtype=np.uint8).tobytes()
            parts.append(d)
            comp += co.compress(d)
            if rng.random() < 0.5:
                comp += co.flush(zlib.Z_SYNC_FLUSH if rng.random() < 0.5 else zlib.Z_FULL_FLUSH)
        comp += co.flush()
        data = b"".join(parts)
        r, out = mine(comp, len(data))
        assert r == len(data) and out == data, trial


def test_damaged_and_truncated_streams_are_rejected():
    rng = np.random.default_rng(2)
    data = np.clip(np.rint(rng.normal(30000, 400, 30000)), 0, 65535).astype(np.uint16).tobytes()
    comp = zlib.compress(data, 4)
    rejected = 0
    for trial in range(4000):
        b = bytearray(comp)
        if trial % 3 == 0:
            b = b[:int(rng.integers(0, len(b)))]
        else:
            for _ in range(int(rng.integers(1, 4))):
                b[int(rng.integers(0, len(b)))] = int(rng.integers(0, 256))
        r, out = mine(bytes(b), len(data))
        if r == -1:
            rejected += 1
        else:
            assert out == data            # the Adler-32 matched: then the bytes are the original ones
    assert rejected > 3900
    for junk in (b"", b"\x78", b"\x78\x9c", b"\x00" * 20, b"\x78\x9c" + b"\xff" * 40):
        assert mine(junk, 100)[0] == -1


@pytest.mark.parametrize("decoder", ["builtin", "zlib"])
def test_reader_gives_the_same_volume_with_either_decoder(tmp_path, decoder):
    """RMG_INFLATE is read once per process: each decoder reads the file in its own interpreter."""
    import os
    import subprocess
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import h5mini
    rng = np.random.default_rng(3)
    vol = np.clip(np.rint(rng.normal(30000, 400, (6, 40, 36))), 0, 65535).astype(np.uint16)
    path = str(tmp_path / "v.mnc")
    h5mini.write_minc2(path, vol, image_min=np.zeros(6), image_max=np.ones(6), valid_range=(0, 65535), chunks=(1, 40, 36), compress=6)
    np.save(tmp_path / "want.npy", vol)
    code = ("import numpy as np, sys\nfrom rminc_b200 import minc2\n"
            f"v = minc2.Minc2Volume({path!r})\nassert np.array_equal(v.stored(n_threads=3), np.load({str(tmp_path / 'want.npy')!r}))\n")
    env = dict(os.environ, RMG_INFLATE="zlib" if decoder == "zlib" else "")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-c", code], cwd=root, env=env, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
