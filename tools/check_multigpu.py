"""torchrun -N check: the N-GPU frame (tiles sharded, blocks gathered over NCCL, merged on rank 0) equals the 1-GPU frame."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nori_b200 import abi, multigpu as MG, scene as S  # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
sc = S.config_cbox(256, 192, 16, S.INT_PATH_MIS)
W, H, b = sc.camera.width, sc.camera.height, sc.border
ctx = abi.Context(local)
ctx.load(sc)
stream = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(stream)
ctx.set_tiles(rank, world)
n, e = ctx.tile_count(rank, world)
blocks = torch.zeros((MG.max_tiles(world, W, H), e, e, 4), dtype=torch.float32, device=dev)
ctx.render_blocks_device(blocks.data_ptr(), stream.cuda_stream)
got = MG.gather_blocks(blocks, world, rank)
if rank == 0:
    film = torch.zeros(sc.film_shape, dtype=torch.float32, device=dev)
    ctx.merge_all_blocks_device(MG.gathered_base(got).data_ptr(), world, MG.max_tiles(world, W, H), film.data_ptr(), stream.cuda_stream)
    stream.synchronize()
    ctx.set_tiles(0, 1)
    ref, _ = ctx.render()
    err = S.rel_l2(film.cpu().numpy(), ref)
    print(f"check_multigpu: world={world} rel-L2(N-GPU film, 1-GPU film) = {err:.3e}")
    assert err < 1e-6
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
