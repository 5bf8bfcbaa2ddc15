"""world_size-2/3 gloo tests (CPU) of the multi-GPU host logic: tile sharding, padded block gather, merge offsets.
The blocks are synthetic (there is no CPU render path); the GPU side of the same flow is exercised by bench.py --gpus N
and tools/check_multigpu.py on the B200 box."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from nori_b200 import multigpu as MG


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _synthetic_blocks(rank, world, W, H, border, n_max):
    E = 32 + 2 * border
    b = np.zeros((n_max, E, E, 4), dtype=np.float32)
    yy, xx = np.mgrid[0:E, 0:E]
    for k, (tid, ox, oy, sx, sy) in enumerate(MG.tiles_of(rank, world, W, H)):
        valid = (yy < sy + 2 * border) & (xx < sx + 2 * border)
        for c in range(4):
            b[k, ..., c] = np.where(valid, (tid + 1) * 0.5 + 0.01 * yy + 0.001 * xx + c, 0.0)
    return b


def _worker(rank, world, port, W, H, border, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n_max = MG.max_tiles(world, W, H)
    blocks = torch.from_numpy(_synthetic_blocks(rank, world, W, H, border, n_max))
    got = MG.gather_blocks(blocks, world, rank, dst=0)
    if rank == 0:
        film = MG.merge_blocks_numpy([g.numpy() for g in got], W, H, border)
        np.save(out_path, film)
    else:
        assert got is None
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,W,H,border", [(2, 100, 70, 2), (3, 96, 64, 1), (2, 31, 33, 0)])
def test_tile_shard_gather_merge(tmp_path, world, W, H, border):
    out = str(tmp_path / "film.npy")
    mp.spawn(_worker, args=(world, _free_port(), W, H, border, out), nprocs=world, join=True)
    film = np.load(out)
    # expected: single-rank assembly of the same synthetic blocks
    n1 = MG.max_tiles(1, W, H)
    all_blocks = np.zeros((n1, 32 + 2 * border, 32 + 2 * border, 4), np.float32)
    where1 = {(ox, oy): tid for tid, ox, oy, _, _ in MG.tiles_of(0, 1, W, H)}     # the numbering depends on the group size: go through the tile origin
    for r in range(world):
        b = _synthetic_blocks(r, world, W, H, border, MG.max_tiles(world, W, H))
        for k, (tid, ox, oy, _, _) in enumerate(MG.tiles_of(r, world, W, H)):
            all_blocks[where1[(ox, oy)]] = b[k]
    expect = MG.merge_blocks_numpy([all_blocks], W, H, border)
    assert np.allclose(film, expect, rtol=1e-6, atol=1e-6)


def test_tiles_partition_the_image():
    for world in (1, 2, 3, 8):
        for W, H in [(800, 600), (768, 768), (33, 1)]:
            seen = np.zeros((H, W), dtype=int)
            ids = []
            for r in range(world):
                for tid, ox, oy, sx, sy in MG.tiles_of(r, world, W, H):
                    assert tid % world == r
                    seen[oy:oy + sy, ox:ox + sx] += 1
                    ids.append(tid)
            assert np.all(seen == 1) and sorted(ids) == list(range(len(ids)))
            counts = [len(MG.tiles_of(r, world, W, H)) for r in range(world)]
            assert max(counts) - min(counts) <= 1
            # ownership follows the Latin pattern (bx + shift by) % world except for the few tiles moved to even out the counts
            shift = 3 if world % 3 else 5
            off = sum(1 for r in range(world) for _, ox, oy, _, _ in MG.tiles_of(r, world, W, H) if (ox // 32 + shift * (oy // 32)) % world != r)
            assert off <= 2 * world


@pytest.mark.parametrize("world", [2, 3, 4, 6, 8])
def test_oracle_shards_follow_the_same_tile_table(oracle, world):
    """The device path (nb_api.cu: build_tile_order), the oracle (oracle.c: orc_render) and this package (multigpu.tile_order)
    each build the group-size dependent tile numbering; the GPU tests compare the first with the other two, this one pins the
    last two against each other on the CPU: rank r's oracle frame has filter weight at the centre of exactly the tiles
    tiles_of(r, world) lists, and the ranks' frames add up to the unsharded frame."""
    from nori_b200 import scene as S
    W, H = 296, 200                                         # 10 x 7 tiles, ragged on both edges
    sc = S.Scene([S.ajax_standin(1)], S.Camera(S.lookat(**S._AJAX_CAM).astype(np.float32), 30.0, W, H), S.INT_NORMALS, 1)
    b = sc.border
    full, _ = oracle.OracleScene(sc).render(accel=1)
    total = np.zeros_like(full)
    for r in range(world):
        o = oracle.OracleScene(sc)
        o.set_tiles(r, world)
        part, st = o.render(accel=1)
        total += part
        mine = {(ox, oy) for _, ox, oy, _, _ in MG.tiles_of(r, world, W, H)}
        assert st.samples == sum(sx * sy for _, _, _, sx, sy in MG.tiles_of(r, world, W, H))
        for _, ox, oy, sx, sy in MG.tiles_of(0, 1, W, H):
            if sx < 8 or sy < 8:
                continue                                    # a sliver's centre lies inside the neighbour's filter footprint
            w = part[b + oy + sy // 2, b + ox + sx // 2, 3]
            assert (w > 0) == ((ox, oy) in mine), (r, ox, oy)
    assert np.allclose(total, full, rtol=1e-6, atol=1e-6)


def test_device_library_builds_the_same_tile_table():
    """nb_debug_tile_order runs nb_api.cu's build_tile_order on the host (no device): the numbering nb_set_tiles / nb_render
    shard by equals multigpu.tile_order for every group size and for ragged, single-row and single-tile grids."""
    from nori_b200 import abi
    for W, H in [(800, 600), (768, 768), (512, 512), (1920, 1080), (33, 1), (20, 20), (100, 700)]:
        for world in range(1, 17):
            assert abi.debug_tile_order(W, H, world) == MG.tile_order(W, H, world), (W, H, world)
