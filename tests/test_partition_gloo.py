"""world_size-2 gloo test of the multi-GPU host logic (no GPU): sample-block partition, film sum-reduce, resolve on rank 0.

The per-rank renders run on the oracle here (the CUDA path needs a GPU; its own linearity in the sample range is covered by
test_gpu_parity.py::test_render_c1_full_size_properties): what this checks is that partition + reduce + resolve reproduce the
single-process image, including an spp that the world size does not divide.
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sample_blocks_cover_the_range():
    from btrc_b200 import partition

    for world in (1, 2, 3, 4, 8):
        for spp in (1, 7, 16, 1024):
            blocks = [partition.sample_block(r, world, spp) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == spp
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            sizes = [e - b for b, e in blocks]
            assert max(sizes) - min(sizes) <= 1
            assert partition.path_range(world - 1, world, spp, 100)[1] == spp * 100


def test_pixel_blocks_cover_the_image():
    from btrc_b200 import partition

    for world in (1, 2, 3, 8):
        for npix in (1, 97, 1920 * 1080):
            blocks = [partition.pixel_block(r, world, npix) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == npix
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            assert max(e - b for b, e in blocks) - min(e - b for b, e in blocks) <= 1
            for r in range(world):
                assert partition.tile_path_range(r, world, 16, npix) == (blocks[r][0] * 16, blocks[r][1] * 16)


def test_interleaved_strips_tile_the_image():
    """The tile partition of `btrc_render_part` (strips of pixel rows dealt round-robin): every pixel row belongs to exactly one
    rank, strips are whole rows, and no rank's share exceeds another's by more than one strip."""
    from btrc_b200 import partition

    for world in (1, 2, 3, 4, 8):
        for (w, h) in ((64, 48), (1920, 1080), (3840, 2160), (7, 5)):
            for spp in (1, 16):
                rows = partition.strip_rows(world, h)
                assert 1 <= rows <= 16
                owner = [-1] * h
                sizes = []
                for r in range(world):
                    n = 0
                    for b, e in partition.tile_path_ranges(r, world, spp, w, h):
                        assert b % (w * spp) == 0 and e % (w * spp) == 0 and e > b
                        for y in range(b // (w * spp), e // (w * spp)):
                            assert owner[y] == -1
                            owner[y] = r
                        n += e - b
                    sizes.append(n)
                assert all(o >= 0 for o in owner)
                assert sum(sizes) == w * h * spp
                assert max(sizes) - min(sizes) <= rows * w * spp


def test_tile_owner_map():
    """partition.tile_owner (the map GPU tests hold btrc_render_part to): every pixel has one owner; where the image does not divide
    into 64 x rows tiles it agrees with the strip list; where it does, tiles are tile_width x rows rectangles dealt round-robin in row-major
    tile order and no rank owns more than one tile more than another."""
    import numpy as np

    from btrc_b200 import partition

    for world in (1, 2, 3, 4, 8):
        for (w, h) in ((64, 48), (1920, 1080), (3840, 2160), (7, 5), (150, 77), (192, 96)):
            rows = partition.strip_rows(world, h)
            assert rows >= 1 and (h % rows == 0 or not any(h % r == 0 for r in range(max(1, rows // 2), rows + 1)))
            owner = partition.tile_owner(world, w, h)
            assert owner.shape == (h, w) and owner.min() >= 0 and owner.max() < world
            if partition.tiles_are_2d(world, w, h, rows):
                tw = partition.tile_width(world, w)
                tx, ty = w // tw, h // rows
                tiles = owner.reshape(ty, rows, tx, tw)
                assert (tiles == tiles[:, :1, :, :1]).all()                      # constant inside a tile
                k = np.arange(tx * ty).reshape(ty, tx)
                assert np.array_equal(tiles[:, 0, :, 0], k % world)
                counts = np.bincount(owner.reshape(-1), minlength=world)
                assert counts.max() - counts.min() <= tw * rows
            else:
                for r in range(world):
                    mine = np.zeros(h, bool)
                    for b, e in partition.tile_path_ranges(r, world, 1, w, h, rows):
                        mine[b // w:e // w] = True
                    assert np.array_equal((owner == r).all(axis=1), mine) and np.array_equal((owner == r).any(axis=1), mine)


def _worker(rank, world, path, port, out):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist

    from btrc_b200 import partition
    from oracle import front

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    o = front.OracleScene(path)
    b, e = partition.sample_block(rank, world, o.spp)
    r = o.render(b, e, raw=True)
    rad, wgt = torch.from_numpy(r["raw_radiance"].copy()), torch.from_numpy(r["raw_weight"].copy())
    partition.reduce_films([rad, wgt], 0)
    if rank == 0:
        color = np.zeros_like(rad.numpy())
        w = wgt.numpy()
        nz = w > 0
        color[nz, :3] = rad.numpy()[nz, :3] / w[nz, None]
        color[..., 3] = 1
        np.save(out, color)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_render_equals_single_process(tmp_path, built):
    import torch.multiprocessing as mp

    from btrc_b200 import scenes
    from oracle import front

    path = scenes.make_cornell(str(tmp_path), width=48, height=48, spp=5, state_count=5000)  # 5 spp over 2 ranks: 3 + 2
    out = str(tmp_path / "color.npy")
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, path, port, out), nprocs=2, join=True)
    got = np.load(out)
    ref = front.OracleScene(path).render()["color"]
    assert np.allclose(got, ref, rtol=1e-5, atol=1e-6)


def test_single_rank_comm_needs_no_nccl(built):
    """btrc_comm_create with one rank builds no communicator (a 1-GPU host needs no NCCL); bad ranks are refused."""
    from btrc_b200 import api

    c = api.Comm(1, 0, None)
    c.close()
    with pytest.raises(api.BtrcError):
        api.Comm(2, 5, bytes(128))
    with pytest.raises(api.BtrcError):
        api.Comm(2, 0, None)
