"""The native (C++) index loaders validate everything before any GPU work: malformed files are reported as such on a
machine without a GPU; well-formed ones get past parsing and then fail loudly for lack of a device (no CPU path)."""
import ctypes
import shutil
import struct

import numpy as np
import pytest

from leann_b200 import capi
from leann_b200 import diskann_format as dfmt


def open_err(fn, *args):
    lib = capi.load()
    h = fn(*args)
    if h:
        lib.lb2_close(ctypes.c_void_p(h))
        return None
    return lib.lb2_last_error().decode()


def gpu_or_none(msg):
    """A well-formed index: either it opened (GPU box) or the only complaint is the missing device."""
    assert msg is None or "no CUDA device" in msg or "sm_100a" in msg, msg


def test_hnsw_loader_rejects_malformed_files(lib, golden_dir, tmp_path):
    good = (golden_dir / "hnsw_small_ip.index").read_bytes()
    f = tmp_path / "x.index"
    f.write_bytes(good)
    gpu_or_none(open_err(lib.lb2_open, str(f).encode(), 0))
    f.write_bytes(good[: len(good) // 2])
    assert "end of file" in open_err(lib.lb2_open, str(f).encode(), 0)
    # a corrupt element count must come back as an error string, never as bad_alloc across the C boundary
    import struct
    from leann_b200.csr import read_compact_index
    nnz = read_compact_index(str(golden_dir / "hnsw_small_ip.index")).neighbors.size
    pos = good.rfind(struct.pack("<Q", nnz))  # header of the last vec<> (compact_neighbors_data), the loosely bounded one
    assert pos > 0
    f.write_bytes(good[:pos] + struct.pack("<Q", (1 << 39) + 7) + good[pos + 8:])
    assert "end of file" in open_err(lib.lb2_open, str(f).encode(), 0)
    f.write_bytes(b"XXXX" + good[4:])
    msg = open_err(lib.lb2_open, str(f).encode(), 0)
    assert msg and "no CUDA device" not in msg
    msg = open_err(lib.lb2_open, str(tmp_path / "missing.index").encode(), 0)
    assert msg and "no CUDA device" not in msg


@pytest.fixture()
def da(golden_dir, tmp_path):
    for p in golden_dir.glob("vamana_small_l2*"):
        shutil.copy(p, tmp_path / p.name)
    return str(tmp_path / "vamana_small_l2")


def da_open(lib, prefix, part=b"", metric=0):
    return open_err(lib.lb2_diskann_open, prefix.encode(), part if isinstance(part, bytes) else part.encode(), metric, 0)


def test_diskann_loader_accepts_the_fixture_in_both_layouts(lib, da):
    gpu_or_none(da_open(lib, da))
    gpu_or_none(da_open(lib, da, da))


def test_diskann_loader_rejects_malformed_files(lib, da):
    assert "metric must be" in da_open(lib, da, metric=7)
    # neighbour id out of range in the standard layout
    raw = bytearray(open(da + "_disk.index", "rb").read())
    meta = np.frombuffer(bytes(raw[8:8 + 72]), np.uint64)
    ndims = int(meta[1])
    struct.pack_into("<I", raw, 4096 + ndims * 4 + 4, 10_000_000)  # first neighbour of node 0
    open(da + "_disk.index", "wb").write(raw)
    assert "neighbour id out of range" in da_open(lib, da)
    # degree larger than the row
    struct.pack_into("<I", raw, 4096 + ndims * 4, 1000)
    open(da + "_disk.index", "wb").write(raw)
    assert "degree exceeds" in da_open(lib, da)
    # the partition layout is independent of _disk.index and still loads
    gpu_or_none(da_open(lib, da, da))
    # id -> partition map inconsistent
    pb = bytearray(open(da + "_partition.bin", "rb").read())
    struct.pack_into("<I", pb, len(pb) - 4, 0)  # last node claims partition 0
    open(da + "_partition.bin", "wb").write(pb)
    assert "disagree" in da_open(lib, da, da)
    # PQ files: chunk count mismatch, truncated codes
    codes = dfmt.read_bin(da + "_pq_compressed.bin", np.uint8)
    dfmt.write_bin(da + "_pq_compressed.bin", codes[:, :-1])
    assert "chunk offsets do not match" in da_open(lib, da)
    open(da + "_pq_compressed.bin", "wb").write(open(da + "_pq_compressed.bin", "rb").read()[:100])
    assert "end of file" in da_open(lib, da)
    assert "cannot open" in da_open(lib, da + "_nope")


def test_diskann_loader_names_unsupported_index_flavours(lib, da):
    open(da + "_pq_pivots.bin_rotation_matrix.bin", "wb").write(b"\0" * 16)
    assert "rotation matrix" in da_open(lib, da)
    import os
    os.remove(da + "_pq_pivots.bin_rotation_matrix.bin")
    open(da + "_disk.index_labels.txt", "w").write("1\n")
    assert "filtered" in da_open(lib, da)
