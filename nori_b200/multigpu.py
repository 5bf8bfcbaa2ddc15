"""Multi-GPU frame assembly: 32x32 image tiles are sharded tile_id % world across ranks (one process per GPU,
replacing BlockGenerator as the scheduler, ref: src/block.cpp:119-152); every rank renders its tiles into packed
ImageBlocks (32+2b edge, border included); ONE exchange at frame end gathers the finished blocks on rank 0, which
adds them into the full film exactly like ImageBlock::put(ImageBlock&) (ref: src/block.cpp:93-102).

The exchange is torch.distributed.gather -- NCCL over NVLink on GPUs (send/recv under the hood: NCCL has no native
gather), gloo on CPU for the host-logic tests.  There is no data-path collective inside the render.
"""
from __future__ import annotations

import numpy as np

BLOCK = 32


def tile_grid(W: int, H: int):
    return (W + BLOCK - 1) // BLOCK, (H + BLOCK - 1) // BLOCK


def tile_order(W: int, H: int, world: int):
    """[(bx, by)] in the tile numbering of the device path for `world` ranks (nb_api.cu: build_tile_order): rank t % world owns tile t,
    ownership follows the Latin pattern (bx + shift * by) % world (shift 3; 5 or 7 when 3 divides world), rank r's k-th tile (row by row) is tile k * world + r; the few tiles by
    which the pattern misses the implied counts move from the ranks with a surplus (their last tiles) to those with a deficit."""
    ntx, nty = tile_grid(W, H)
    total = ntx * nty
    lists = [[] for _ in range(world)]
    shift = 3 if world % 3 else 5 if world % 5 else 7
    for by in range(nty):
        for bx in range(ntx):
            lists[(bx + shift * by) % world].append((bx, by))
    target = [(total - r + world - 1) // world if total > r else 0 for r in range(world)]
    pool = []
    for r in range(world):
        while len(lists[r]) > target[r]:
            pool.append(lists[r].pop())
    for r in range(world):
        while len(lists[r]) < target[r]:
            lists[r].append(pool.pop(0))
    order = [None] * total
    for r in range(world):
        for k, t in enumerate(lists[r]):
            order[k * world + r] = t
    return order


def tiles_of(rank: int, world: int, W: int, H: int):
    """[(tile_id, ox, oy, sx, sy)] owned by `rank`, in the packed-block order of nb_render_blocks_device."""
    order = tile_order(W, H, world)
    out = []
    for tid in range(rank, len(order), world):
        bx, by = order[tid]
        ox, oy = bx * BLOCK, by * BLOCK
        out.append((tid, ox, oy, min(BLOCK, W - ox), min(BLOCK, H - oy)))
    return out


def max_tiles(world: int, W: int, H: int) -> int:
    return max(len(tiles_of(r, world, W, H)) for r in range(world))


def gather_blocks(blocks, world: int, rank: int, dst: int = 0):
    """blocks: torch tensor [n_max, E, E, 4] (padded to the same n_max on every rank).  Returns the list of all ranks'
    tensors on dst (None elsewhere).  One collective per frame."""
    import torch
    import torch.distributed as dist
    if world == 1:
        return [blocks]
    out = None
    if rank == dst:     # one contiguous [world, n_max, E, E, 4] buffer, so the merge is a single launch
        big = torch.empty((world,) + tuple(blocks.shape), dtype=blocks.dtype, device=blocks.device)
        out = [big[r] for r in range(world)]
    dist.gather(blocks, out, dst=dst)
    return out


def gathered_base(gathered):
    """The contiguous [world, n_max, E, E, 4] tensor behind gather_blocks' list (world > 1) or the single tensor."""
    return gathered[0]._base if getattr(gathered[0], "_base", None) is not None else gathered[0]


def merge_blocks_numpy(blocks_per_rank, W: int, H: int, border: int):
    """Host restatement of the merge (nb_merge_blocks_device) for tests: film += block at (offset - border)."""
    E = BLOCK + 2 * border
    world = len(blocks_per_rank)
    film = np.zeros((H + 2 * border, W + 2 * border, 4), dtype=np.float32)
    for r, blocks in enumerate(blocks_per_rank):
        blocks = np.asarray(blocks)
        assert blocks.shape[1:] == (E, E, 4)
        for k, (tid, ox, oy, sx, sy) in enumerate(tiles_of(r, world, W, H)):
            film[oy:oy + sy + 2 * border, ox:ox + sx + 2 * border] += blocks[k, :sy + 2 * border, :sx + 2 * border]
    return film
