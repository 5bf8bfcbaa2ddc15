"""torchrun -N check of the process-per-GPU flavour of the C-side group render (nb_comm_init_rank + nb_render_gather):
the N-GPU frame (tiles sharded tile_id % N, finished blocks gathered with ONE grouped ncclSend/ncclRecv inside
libnori_b200.so, merged on rank 0) equals the 1-GPU frame.  torch.distributed only ships the 128-byte communicator id.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/check_multigpu.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nori_b200 import abi, scene as S  # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
ok = True
for sc in (S.config_cbox(256, 192, 16, S.INT_PATH_MIS), S.Scene([S.ajax_standin(2)], S.Camera(S.lookat(**S._AJAX_CAM).astype(np.float32), 30.0, 200, 136), S.INT_AO, 4)):
    ctx = abi.Context(local)
    if world > 1:
        uid = torch.tensor(list(abi.Context.comm_unique_id()) if rank == 0 else [0] * 128, dtype=torch.uint8, device=dev)
        dist.broadcast(uid, 0)
        ctx.comm_init_rank(bytes(uid.cpu().tolist()), rank, world)
    ctx.load(sc)            # rank 0 builds the hierarchy, the others receive the arrays over NVLink
    film = torch.zeros(sc.film_shape, dtype=torch.float32, device=dev) if rank == 0 else None
    st = ctx.render_gather(film.data_ptr() if rank == 0 else 0)
    ctx.upload()            # sharded re-upload: 1/N per PCIe link + one in-place ncclAllGather; the frame must not change
    st = ctx.render_gather(film.data_ptr() if rank == 0 else 0)
    rays = torch.tensor([st.rays], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(rays)
    if rank == 0:
        with abi.Context(local) as one:
            one.load(sc)
            ref, st1 = one.render()
        err = S.rel_l2(film.cpu().numpy(), ref)
        print(f"check_multigpu: {sc.name or 'ajax-ao'} world={world} rel-L2(N-GPU film, 1-GPU film) = {err:.3e}  rays {int(rays.item())} vs {st1.rays}")
        ok = ok and err < 1e-6 and int(rays.item()) == st1.rays
    ctx.close()
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
if rank == 0:
    assert ok
    print("check_multigpu: OK")
