"""N GPUs behind the C-ABI (VERDICT r1 "Next" item 5): nb_create_multi / `nori --gpus N` -- tiles sharded tile_id % N,
ONE NCCL gather of the finished ImageBlocks, ONE merge launch -- must return the 1-GPU film.

Needs a box with >= 2 GPUs (`gpurun --gpus 2`); on a 1-GPU box the tests skip (NCCL refuses two ranks on one device).
The process-per-GPU flavour (nb_comm_init_rank, what bench.py uses under torchrun) is checked by tools/check_multigpu.py.
"""
import os
import subprocess

import numpy as np
import pytest

from nori_b200 import abi, host
from nori_b200 import scene as S

pytestmark = pytest.mark.gpu


def _ngpus():
    import glob
    return len(glob.glob("/dev/nvidia[0-9]*"))


needs2 = pytest.mark.skipif(_ngpus() < 2, reason="needs >= 2 GPUs on the box (gpurun --gpus 2)")


def _scene(integrator=S.INT_AO, spp=4, w=200, h=136):
    cam = S.Camera(S.lookat(**S._AJAX_CAM).astype(np.float32), 30.0, w, h)
    return S.Scene([S.ajax_standin(2)], cam, integrator, spp, name="multi-ajax")


def test_group_of_one_is_the_plain_context(oracle):
    """nb_create_multi with one device is nb_create; nb_render_gather without a group is nb_render_device."""
    sc = _scene()
    with abi.Context([0]) as ctx:
        assert ctx.device_count == 1
        ctx.load(sc)
        film, st = ctx.render()
    ofilm, ost = oracle.OracleScene(sc).render(accel=1)
    assert st.rays == ost.rays and S.rel_l2(film, ofilm) <= 1e-4
    with pytest.raises(abi.NoriError, match="twice"):
        abi.Context([0, 0])


@needs2
@pytest.mark.parametrize("case", ["ao", "path_mis"])
def test_two_devices_behind_one_context(oracle, case):
    n = min(_ngpus(), 4)
    sc = _scene() if case == "ao" else S.config_cbox(96, 80, 8, S.INT_PATH_MIS)
    with abi.Context(0) as one:
        one.load(sc)
        ref, st1 = one.render()
    for devs in ([0, 1], list(range(n))):
        with abi.Context(devs) as ctx:
            assert ctx.device_count == len(devs)
            ctx.load(sc)                       # ONE host build, scene arrays replicated over NVLink (ncclBroadcast)
            film, st = ctx.render()
            assert st.samples == st1.samples and st.rays == st1.rays
            assert S.rel_l2(film, ref) <= 1e-6
            ctx.upload()                       # re-upload: PCIe once + broadcast
            film2, _ = ctx.render()
            assert S.rel_l2(film2, ref) <= 1e-6
            with pytest.raises(abi.NoriError, match="group owns the tile assignment"):
                ctx.set_tiles(0, 1)
            sc2 = _scene(S.INT_NORMALS, 2, 100, 70)      # setters reach every device; ragged tile grid, fewer tiles than usual
            ctx.load(sc2)
            f3, s3 = ctx.render()
            of, os_ = oracle.OracleScene(sc2).render(accel=1)
            assert s3.rays == os_.rays and S.rel_l2(f3, of) <= 1e-4


@needs2
def test_cli_gpus_2_matches_one_gpu(tmp_path):
    sc = S.config_cbox(96, 64, 8, S.INT_PATH_MIS)
    path = host.write_xml(sc, str(tmp_path), "cbox")
    hs = host.HostScene(path)
    f1, s1 = hs.render(0)
    f2, s2 = hs.render(0, gpus=2)
    assert s2.samples == s1.samples and s2.rays == s1.rays
    assert S.rel_l2(f2, f1) <= 1e-6
    if not os.path.exists(host.CLI_PATH):
        from nori_b200 import build as nb_build
        nb_build.build_host(force=True)
    r = subprocess.run([host.CLI_PATH, path, "--no-gui", "--gpus", "2"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "Rendering .. " in r.stdout and "done. (took" in r.stdout      # (NCCL may print its version line in between)
    exr = (tmp_path / "cbox.exr").read_bytes()
    W, H = 96, 64
    body = np.frombuffer(exr[-(H * (8 + 3 * W * 4)):], dtype=np.uint8).reshape(H, 8 + 3 * W * 4)[:, 8:]
    bgr = body.copy().view(np.float32).reshape(H, 3, W)
    rgb = np.stack([bgr[:, 2], bgr[:, 1], bgr[:, 0]], axis=-1)
    with abi.Context(0) as ctx:
        ctx.load(sc)
        ref = ctx.film_to_rgb(f1)
    assert S.rel_l2(rgb, ref) <= 1e-6
