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


def _decode(plan, u):
    """render_kernel's decode of work unit u (nb_kernels.cuh, regeneration loop) -> (tile slot, patch, first sample, samples)."""
    if u < plan["split_units"]:
        chunk, nchunks, s0 = plan["chunk_a"], plan["nchunks_a"], 0
    else:
        u -= plan["split_units"]
        chunk, nchunks, s0 = plan["chunk"], plan["nchunks"], plan["split_sample"]
    patch, rest = u % 32, u // 32
    return rest // nchunks, patch, s0 + (rest % nchunks) * chunk, chunk


def test_work_unit_plan_covers_every_sample_once():
    """The guided schedule (nb_api.cu: plan_units, through nb_debug_unit_plan): whatever the options, the units of a launch
    cover every (tile, patch, sample) exactly once, coarse units come first and stay inside [0, split_sample)."""
    from nori_b200 import abi
    B200 = 148 * 10 * 4
    assert abi.debug_unit_plan(475, 64, B200) == dict(chunk=2, nchunks=8, split_sample=48, chunk_a=8, nchunks_a=6, split_units=475 * 32 * 6, n_units=475 * 32 * 14)
    p = abi.debug_unit_plan(119, 64, B200)                                     # a rank of the 4-GPU headline frame
    assert (p["chunk_a"], p["split_sample"], p["chunk"]) == (2, 48, 1)
    p = abi.debug_unit_plan(60, 64, B200)                                      # ... of the 8-GPU frame: too little work for coarse units
    assert (p["split_units"], p["chunk"], p["n_units"]) == (0, 1, 60 * 32 * 64)
    assert abi.debug_unit_plan(2040, 4096, B200)["chunk"] == 8                  # configs[4]: 8 samples everywhere
    with pytest.raises(abi.NoriError):
        abi.debug_unit_plan(1 << 20, 1 << 20, B200, chunk=1)                   # > 0xf0000000 units
    rng = np.random.default_rng(7)
    cases = [(5, 13, 4, 0, -1, 0), (7, 64, 8, 0, 75, 8), (3, 100, 2, 0, 100, 8), (4, 37, 1, 0, 40, 4), (6, 64, 8, 0, 0, 0), (2, 9, 4, 3, -1, 0),
             (1, 1, 4, 0, -1, 0), (0, 8, 4, 0, -1, 0), (9, 200, 16, 0, 75, 16), (3, 601, 3, 0, 1, 2)]
    cases += [(int(rng.integers(1, 12)), int(rng.integers(1, 300)), int(rng.integers(1, 12)), int(rng.choice([0, 0, 0, 0, 1, 5, 64])),
               int(rng.integers(-1, 101)), int(rng.choice([0, 2, 4, 8, 16]))) for _ in range(60)]
    split_seen = 0
    for n_tiles, spp, warps, chunk, guided, coarse in cases:
        p = abi.debug_unit_plan(n_tiles, spp, warps, chunk, guided, coarse)
        seen = np.zeros((max(n_tiles, 1), 32, spp), dtype=np.int32)
        for u in range(p["n_units"]):
            slot, patch, s0, ns = _decode(p, u)
            assert slot < n_tiles
            s1 = min(s0 + ns, spp)                                               # the kernel clips the last chunk to spp
            if u < p["split_units"]:
                assert s1 <= p["split_sample"]
            else:
                assert s0 >= p["split_sample"]
            seen[slot, patch, s0:s1] += 1
        assert p["n_units"] == 0 if n_tiles == 0 else np.all(seen == 1), (n_tiles, spp, warps, chunk, guided, coarse, p)
        split_seen += p["split_units"] > 0
        if chunk > 0:
            assert p["split_units"] == 0 and p["chunk"] == min(chunk, spp)
    assert split_seen >= 10                                                       # the coarse / fine split was exercised


def test_tile_table_properties_for_arbitrary_frames():
    """Property test (hypothesis): for any frame and group size the table is a permutation of the tile grid, the ranks' counts
    differ by at most one, rank r owns tiles r, r + N, ..., and no rank owns a whole column of a grid it could share."""
    from hypothesis import given, settings, strategies as st
    from nori_b200 import abi

    @settings(max_examples=150, deadline=None)
    @given(st.integers(1, 2200), st.integers(1, 1300), st.integers(1, 16))
    def check(W, H, world):
        order = abi.debug_tile_order(W, H, world)
        ntx, nty = MG.tile_grid(W, H)
        assert sorted(order) == [(bx, by) for bx in range(ntx) for by in range(nty)]
        assert order == MG.tile_order(W, H, world)
        counts = [len(range(r, len(order), world)) for r in range(world)]
        assert max(counts) - min(counts) <= 1
        if world > 1 and ntx >= world and nty >= world:
            for r in range(world):
                cols = {}
                for bx, by in order[r::world]:
                    cols[bx] = cols.get(bx, 0) + 1
                assert max(cols.values()) < nty, (W, H, world, r)
    check()
    with pytest.raises(abi.NoriError):
        abi.debug_tile_order(0, 10, 2)
