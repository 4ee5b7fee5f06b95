"""Multi-GPU partition of one render: sample streams split across ranks, one sum-reduce of the film buffers.

Every camera path is fully determined by (pixel, sample index) through the sampler seed (builtin/sampler/independent.cpp:
11-15), so a render splits into independent ranges of the path index space (Uniform mode: index = sample * W*H + pixel,
generate.cpp:49-60).  Rank r of N renders the contiguous block of sample indices [r*spp/N, (r+1)*spp/N) for all pixels into
its own film; the films (radiance float4 + weight float [+ albedo, normal]) are summed onto rank 0 -- the path's only
collective -- and resolved there (value / weight, preview.cpp:28-80).  The reference is single-GPU; this is new work.
"""
from __future__ import annotations


def sample_block(rank: int, world: int, spp: int):
    """Sample indices [begin, end) of `rank`: blocks differ by at most one sample when world does not divide spp."""
    base, rem = divmod(spp, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def path_range(rank: int, world: int, spp: int, npix: int):
    """Camera-path index range [begin, end) of `rank` in Uniform mode (sample-major index space)."""
    b, e = sample_block(rank, world, spp)
    return b * npix, e * npix


def pixel_block(rank: int, world: int, npix: int):
    """Pixels [begin, end) of `rank` in row-major order (a band of image rows, up to one partial row at each end)."""
    base, rem = divmod(npix, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def tile_path_range(rank: int, world: int, spp: int, npix: int):
    """Camera-path index range [begin, end) of `rank` in Tile mode (pixel-major index space, index = pixel * spp + sample:
    generate.cpp:56-60; the scene's renderer must have "tile": true).  Image-tile partition of BASELINE config 5: every rank
    renders all samples of its own band of pixels; the films of the other bands stay zero, so the same sum-reduce doubles as
    the gather.  Bands are not load balanced (sky rows are cheap) -- the sample-stream partition is the default for that reason."""
    b, e = pixel_block(rank, world, npix)
    return b * spp, e * spp


def tile_width(world: int, width: int) -> int:
    """Width of the 2-D tiles of Renderer::render_part (renderer.cpp: kTileW): 64 pixels, or the first of 128 / 256 / 32 that makes
    the number of tiles per image row coprime to the number of ranks, so that the round-robin deal walks diagonally over the image."""
    from math import gcd

    for tw in (64, 128, 256, 32):
        if width % tw == 0 and gcd(width // tw, world) == 1:
            return tw
    return 64


def strip_rows(world: int, height: int) -> int:
    """Default tile / strip height of the interleaved tile partition (Renderer::render_part): at most 16 rows, at least 8 strips
    per rank, snapped down to a divisor of the height when one lies in [rows / 2, rows]."""
    rows = max(1, min(16, height // (8 * world)))
    for r in range(rows, max(1, rows // 2) - 1, -1):
        if height % r == 0:
            return r
    return rows


def tiles_are_2d(world: int, width: int, height: int, rows: int) -> bool:
    """Renderer::render_part deals tile_width x rows tiles when they divide the image, whole-row strips otherwise."""
    return width % tile_width(world, width) == 0 and height % rows == 0


def tile_owner(world: int, width: int, height: int, rows: int = 0):
    """Owning rank of every pixel, shape (height, width), under btrc_render_part's interleaved tile partition of a tile-mode
    render: tile (or strip) k in row-major tile order belongs to rank k % world.  What tests check the C ABI call against."""
    import numpy as np

    rows = rows or strip_rows(world, height)
    y, x = np.mgrid[0:height, 0:width]
    if tiles_are_2d(world, width, height, rows):
        tw = tile_width(world, width)
        k = (y // rows) * (width // tw) + x // tw
    else:
        k = y // rows
    return (k % world).astype(np.int32)


def tile_path_ranges(rank: int, world: int, spp: int, width: int, height: int, rows: int = 0):
    """Camera-path index ranges [(begin, end), ...] of `rank` in the INTERLEAVED tile partition (Tile mode, "tile": true): the
    image is cut into strips of `rows` pixel rows and strip k belongs to rank k % world, so sky and geometry rows are spread
    over all ranks (contiguous bands scaled 1.46x at 2 GPUs on C5, profiles/bench_c5_n2_tiles_r1h.json).  The C ABI renders the
    same set in one call (btrc_render_part) when the image does not divide into 2-D tiles (`tiles_are_2d`), or when
    BTRC_TILES_2D=0; this list form is what tests check the strip case against."""
    rows = rows or strip_rows(world, height)
    strips = (height + rows - 1) // rows
    out = []
    for k in range(rank, strips, world):
        y0, y1 = k * rows, min((k + 1) * rows, height)
        out.append((y0 * width * spp, y1 * width * spp))
    return out


def reduce_films(tensors, dst: int = 0):
    """Sum-reduce the film buffers onto `dst` (torch.distributed; NCCL on GPUs, gloo in the CPU tests)."""
    import torch.distributed as dist

    if not dist.is_initialized() or dist.get_world_size() == 1:
        return
    for t in tensors:
        if t is not None:
            dist.reduce(t, dst, op=dist.ReduceOp.SUM)
