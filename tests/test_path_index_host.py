"""The generate kernel's path-index maps (btrc_b200/csrc/device/path_index.cuh), compiled for the host: every map must be ONE-TO-ONE
onto the index set the renderer asks for -- the reference's image is the sum over exactly those (pixel, sample) paths (generate.cpp:
49-60), whatever the visiting order.  Covers index order, interleaved strips, interleaved 2-D tiles (pixel-major and warp-chunked),
the block-coherent order in its three in-block modes with ragged image edges, and renders with more than 2^32 paths (64-bit branch)
against an independent numpy restatement.  Runs without a GPU; the GPU tests check the same functions inside generate_kernel."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from .util import ROOT


@pytest.fixture(scope="module")
def hpi():
    src = os.path.join(ROOT, "tests", "host", "host_path_index.cpp")
    out = os.path.join(ROOT, "tests", "host", "_libhostpathindex.so")
    deps = [src, os.path.join(ROOT, "btrc_b200/csrc/device/path_index.cuh"), os.path.join(ROOT, "btrc_b200/csrc/common/types.h")]
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        subprocess.run(["g++", "-std=c++17", "-O2", "-fopenmp", "-fPIC", "-shared", "-I/usr/local/cuda/include", src, "-o", out], check=True)
    L = ctypes.CDLL(out)
    L.hpi_decode.argtypes = [ctypes.c_int] * 4 + [ctypes.c_longlong, ctypes.c_longlong] + [ctypes.c_int] * 9 + [ctypes.c_longlong, ctypes.c_void_p, ctypes.c_void_p]
    return L


def _decode(L, idx, W, H, spp, tile, end, stride=(0, 1, 0, 0, 0), order=(0, 0, 0, 0, 0)):
    idx = np.ascontiguousarray(idx, np.int64)
    out = np.empty_like(idx)
    L.hpi_decode(W, H, spp, int(tile), end, *stride, *order, idx.size, idx.ctypes.data, out.ctypes.data)
    return out


def _order_reference(v, W, H, ns, s0, bw, bh, mode):
    """Independent restatement (numpy int64) of the block-coherent order: rows of blocks, blocks left to right, inside a block by `mode`."""
    v = np.asarray(v, np.int64)
    per_row = W * bh * ns
    br = v // per_row
    r = v - br * per_row
    hr = np.minimum(bh, H - br * bh)
    if mode == 0:
        pq = r // ns; smp = s0 + r - pq * ns
        bc = pq // (bw * hr); within = pq - bc * bw * hr
    else:
        full = bw * hr * ns
        bc = r // full; ib = r - bc * full
        wc0 = np.minimum(bw, W - bc * bw); npb = wc0 * hr
        if mode == 2:
            ch = ib // (npb * 32); rem = ib - ch * npb * 32
            smp = s0 + ch * 32 + (rem & 31); within = rem >> 5
        else:
            pl = ib // npb; smp = s0 + pl; within = ib - pl * npb
    wc = np.minimum(bw, W - bc * bw)
    y = br * bh + within // wc; x = bc * bw + within % wc
    return smp * (W * H) + y * W + x


@pytest.mark.parametrize("W,H", [(96, 54), (70, 33), (64, 32), (33, 70)])
@pytest.mark.parametrize("mode", [0, 1, 2])
def test_block_coherent_order_is_a_permutation(hpi, W, H, mode):
    for (ns, s0) in ((32, 0), (64, 3), (5, 2)):
        if mode == 2 and ns % 32:
            continue
        for (bw, bh) in ((64, 32), (32, 32), (8, 8), (7, 5)):
            n = W * H * ns
            got = _decode(hpi, np.arange(n), W, H, s0 + ns, False, n, order=(ns, s0, bw, bh, mode))
            assert np.array_equal(np.sort(got), np.arange(s0 * W * H, (s0 + ns) * W * H)), (ns, s0, bw, bh)
            assert np.array_equal(got, _order_reference(np.arange(n), W, H, ns, s0, bw, bh, mode))
            if mode in (0, 2) and ns % 32 == 0:   # a warp (32 consecutive indices from a multiple of 32) starts in ONE pixel
                pix = (got % (W * H)).reshape(-1, 32)
                assert (pix == pix[:, :1]).all()


@pytest.mark.parametrize("spp", [5, 32, 96])
def test_tile_partitions_cover_the_index_space_once(hpi, spp):
    from btrc_b200 import partition

    for (W, H, N) in ((128, 48, 3), (192, 96, 8), (256, 32, 2), (150, 77, 3)):
        rows = partition.strip_rows(N, H)
        total = W * H * spp
        seen = []
        if partition.tiles_are_2d(N, W, H, rows):
            tw = partition.tile_width(N, W)
            block = tw * rows * spp
            tiles = (W // tw) * (H // rows)
            owner = partition.tile_owner(N, W, H).reshape(-1)
            for r in range(N):
                owned = (tiles - 1 - r) // N + 1 if r < tiles else 0
                got = _decode(hpi, np.arange(owned * block), W, H, spp, True, owned * block, stride=(block, N, r, tw, rows))
                assert (got >= 0).all() and (owner[got // spp] == r).all()       # exactly the pixels partition.tile_owner gives the rank
                if spp % 32 == 0:
                    pix = (got // spp).reshape(-1, 32)
                    assert (pix == pix[:, :1]).all()
                seen.append(got)
        else:
            block = rows * W * spp
            strips = (H + rows - 1) // rows
            for r in range(N):
                owned = (strips - 1 - r) // N + 1 if r < strips else 0
                if owned == 0:
                    continue
                owns_last = (strips - 1) % N == r
                n_virtual = (owned - 1) * block + (total - (strips - 1) * block) if owns_last else owned * block
                got = _decode(hpi, np.arange(n_virtual), W, H, spp, True, n_virtual, stride=(block, N, r, 0, 0))
                seen.append(got)
        allv = np.concatenate(seen)
        assert np.array_equal(np.sort(allv), np.arange(total)), (W, H, N, spp)


def test_index_order_and_ranges_beyond_32_bits(hpi):
    """Index order is the identity in both modes; renders with more than 2^32 paths take the 64-bit division and agree with numpy."""
    rng = np.random.default_rng(11)
    W, H, spp = 3840, 2160, 1024          # BASELINE config 5 on one GPU: 8.49e9 paths
    total = W * H * spp
    assert total > 2 ** 32
    idx = np.concatenate([rng.integers(0, total, 200000), np.arange(2 ** 32 - 1000, 2 ** 32 + 1000), [0, total - 1]])
    for tile in (False, True):
        assert np.array_equal(_decode(hpi, idx, W, H, spp, tile, total), idx)
    # block-coherent order over all 1024 planes: 64-bit branch of the row division, chunked and pixel-major blocks
    for (bw, bh, mode) in ((32, 32, 2), (64, 32, 0), (64, 32, 1)):
        got = _decode(hpi, idx, W, H, spp, False, total, order=(spp, 0, bw, bh, mode))
        assert np.array_equal(got, _order_reference(idx, W, H, spp, 0, bw, bh, mode))
    # and the same render split in two halves of 512 planes (each below 2^32 paths: the 32-bit branch) gives the same set per half
    half = W * H * 512
    sub = idx[idx < half]
    for s0 in (0, 512):
        got = _decode(hpi, sub, W, H, spp, False, half, order=(512, s0, 32, 32, 2))
        assert np.array_equal(got, _order_reference(sub, W, H, 512, s0, 32, 32, 2))
    # 2-D tiles of the 4K tile-mode space at 2 ranks: rank 0's share has 4.25e9 paths (just below 2^32), a single "rank" all 8.49e9
    for (N, r) in ((1, 0), (2, 1)):
        tw, rows = 256, 16
        block = tw * rows * spp
        tiles = (W // tw) * (H // rows)
        owned = (tiles - 1 - r) // N + 1
        v = rng.integers(0, owned * block, 200000)
        got = _decode(hpi, v, W, H, spp, True, owned * block, stride=(block, N, r, tw, rows))
        blk = (v // block) * N + r; within = v % block
        ch = within // (tw * rows * 32); rem = within - ch * tw * rows * 32
        pq = rem >> 5; smp = ch * 32 + (rem & 31)
        tx = W // tw
        y = (blk // tx) * rows + pq // tw; x = (blk % tx) * tw + pq % tw
        assert np.array_equal(got, (y * W + x) * spp + smp)
