"""GPU tests of the reference-facing host path: Nori XML -> C++ parser/plugins -> C-ABI -> film, and the multi-GPU
block API (render_blocks / merge) on one device."""
import os
import subprocess

import numpy as np
import pytest

from nori_b200 import abi, host, multigpu as MG
from nori_b200 import scene as S

pytestmark = pytest.mark.gpu


def test_xml_host_render_matches_python_path_and_oracle(tmp_path, oracle):
    sc = S.config_cbox(96, 64, 8, S.INT_PATH_MIS)
    sc.meshes[3] = S.with_(sc.meshes[3], S.microfacet((0.2, 0.2, 0.4), 0.28, 1.7))
    path = host.write_xml(sc, str(tmp_path), "cbox")
    hs = host.HostScene(path)
    film_h, st = hs.render(0)
    with abi.Context(0) as ctx:
        ctx.load(sc)
        film_p, st_p = ctx.render()
    ofilm, _ = oracle.OracleScene(sc).render(accel=1)
    assert st.samples == st_p.samples == 96 * 64 * 8
    assert S.rel_l2(film_h, ofilm) < 1e-4      # the C++ host computes its own sampleToCamera (double, rounded once)
    assert S.rel_l2(film_p, ofilm) < 1e-4


def test_reference_scene_bunny_through_cli(tmp_path, oracle):
    """BASELINE configs[0] end to end through the `nori` executable: XML in, EXR + PNG out."""
    sc = S.config_bunny()      # 768x768, 1 spp, normals, per-block seeding
    path = host.write_xml(sc, str(tmp_path), "bunny")
    if not os.path.exists(host.CLI_PATH):          # the executable is a g++-only artefact: rebuild it if the snapshot lost it
        from nori_b200 import build as nb_build
        nb_build.build_host(force=True)
    r = subprocess.run([host.CLI_PATH, path, "--no-gui", "--threads", "4"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "Rendering .. done." in r.stdout
    exr = tmp_path / "bunny.exr"; png = tmp_path / "bunny.png"
    assert exr.exists() and png.exists()
    raw = exr.read_bytes()
    assert raw[:4] == bytes([0x76, 0x2f, 0x31, 0x01])
    # uncompressed scanline EXR written by Bitmap::saveEXR: the pixel payload is the tail of the file
    W = H = 768
    body = np.frombuffer(raw[-(H * (8 + 3 * W * 4)):], dtype=np.uint8).reshape(H, 8 + 3 * W * 4)[:, 8:]
    bgr = body.copy().view(np.float32).reshape(H, 3, W)
    rgb = np.stack([bgr[:, 2], bgr[:, 1], bgr[:, 0]], axis=-1)
    ofilm, _ = oracle.OracleScene(sc).render(accel=1)
    assert S.rel_l2(rgb, oracle.film_to_rgb(ofilm, W, H, sc.border)) < 1e-4
    import cv2
    img = cv2.imread(str(png))
    assert img is not None and img.shape == (H, W, 3) and img.max() > 100
    bad = subprocess.run([host.CLI_PATH, str(tmp_path / "nope.xml")], capture_output=True, text=True)
    assert bad.returncode != 0 and "Fatal error" in bad.stderr


def test_unsupported_plugin_is_an_error_not_a_fallback(tmp_path):
    (tmp_path / "t.obj").write_text("v 0 0 0\nv 1 0 0\nv 0 1 0\nf 1 2 3\n")
    (tmp_path / "s.xml").write_text("""<scene><integrator type="ao"/><camera type="perspective"/>
      <sampler type="independent"><integer name="sampleCount" value="0"/></sampler>
      <mesh type="obj"><string name="filename" value="t.obj"/></mesh></scene>""")
    with pytest.raises(abi.NoriError) as e:
        host.HostScene(str(tmp_path / "s.xml")).render(0)
    assert "sampleCount" in str(e.value)


def test_block_api_shards_merge_to_the_full_frame():
    import torch
    sc = S.Scene([S.ajax_standin(2)], S.Camera(S.lookat(**S._AJAX_CAM).astype(np.float32), 30.0, 200, 136), S.INT_AO, 4)
    W, H, b = 200, 136, sc.border
    with abi.Context(0) as ctx:
        ctx.load(sc)
        full, _ = ctx.render()
        world = 3
        per_rank = []
        stream = torch.cuda.Stream()
        film = torch.zeros(sc.film_shape, dtype=torch.float32, device="cuda")
        with torch.cuda.stream(stream):
            for r in range(world):
                ctx.set_tiles(r, world)
                n, e = ctx.tile_count(r, world)
                assert n == len(MG.tiles_of(r, world, W, H)) and e == 32 + 2 * b
                blocks = torch.zeros((MG.max_tiles(world, W, H), e, e, 4), dtype=torch.float32, device="cuda")
                st = ctx.render_blocks_device(blocks.data_ptr(), stream.cuda_stream)
                assert st.samples == sum(sx * sy for _, _, _, sx, sy in MG.tiles_of(r, world, W, H)) * 4
                ctx.merge_blocks_device(blocks.data_ptr(), r, world, film.data_ptr(), stream.cuda_stream)
                per_rank.append(blocks.cpu().numpy())
            stream.synchronize()
        ctx.set_tiles(0, 1)
        film2 = torch.zeros(sc.film_shape, dtype=torch.float32, device="cuda")
        allb = torch.from_numpy(np.stack(per_rank)).cuda()
        ctx.merge_all_blocks_device(allb.data_ptr(), world, allb.shape[1], film2.data_ptr(), 0)
        torch.cuda.synchronize()
    assert S.rel_l2(film.cpu().numpy(), full) < 1e-6
    assert S.rel_l2(film2.cpu().numpy(), full) < 1e-6
    assert S.rel_l2(MG.merge_blocks_numpy(per_rank, W, H, b), full) < 1e-6


def test_device_tonemap_bytes_equal_host_loop_and_png(tmp_path, oracle):
    """normalise + sRGB tonemap + 8-bit pack run on the device (film_to_srgb8_kernel) and feed the PNG writer;
    the bytes equal the host loop's (Bitmap::toSRGB8: ref src/common.cpp:166-180, src/bitmap.cpp:100-110)."""
    sc = S.config_cbox(160, 120, 16, S.INT_PATH_MIS)      # dark corners (linear segment) and bright light (clamped) in one image
    path = host.write_xml(sc, str(tmp_path), "cbox")
    hs = host.HostScene(path)
    film, st, dev8, host8 = hs.render(0, srgb8=True)
    assert dev8.shape == (120, 160, 3) and dev8.max() == 255 and dev8.min() < 30
    assert np.array_equal(dev8, host8)
    with abi.Context(0) as ctx:
        ctx.load(sc)
        f2, _ = ctx.render()
        b2 = ctx.last_film_to_srgb8()                      # (this film comes from Python's camera matrices, the host's differ in the last ulp)
        assert b2.shape == dev8.shape and np.abs(b2.astype(int) - dev8.astype(int)).max() <= 1
        ctx.configure(sc)                                  # a new camera / filter invalidates the device film
        with pytest.raises(abi.NoriError, match="no film on the device"):
            ctx.last_film_to_srgb8()
    if not os.path.exists(host.CLI_PATH):
        from nori_b200 import build as nb_build
        nb_build.build_host(force=True)
    r = subprocess.run([host.CLI_PATH, path, "--no-gui"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    import cv2
    img = cv2.imread(str(tmp_path / "cbox.png"))[..., ::-1].astype(int)           # BGR -> RGB; another render: the film's float atomics
    diff = np.abs(img - dev8.astype(int))                                         # land in another order, a byte on a rounding edge may move
    assert diff.max() <= 1 and (diff != 0).mean() < 1e-3


def test_hierarchy_cache_on_the_render_path(tmp_path, oracle):
    """the built BVH is cached on disk (`nori --cache`, nb_set_accel_cache); a hit skips the build and
    renders the same film."""
    sc = S.Scene([S.ajax_standin(3)], S.Camera(S.lookat(**S._AJAX_CAM).astype(np.float32), 30.0, 160, 120), S.INT_AO, 4)
    cache = tmp_path / "scene.nbbvh"
    with abi.Context(0) as ctx:
        ctx.set_accel_cache(cache)
        ctx.load(sc)
        assert not ctx.accel_cache_hit and cache.exists()
        ref, st = ctx.render()
    with abi.Context(0) as ctx:
        ctx.set_accel_cache(cache)
        ctx.load(sc)
        assert ctx.accel_cache_hit and ctx.build_stats()["seconds"] == 0.0
        film, st2 = ctx.render()
        assert st2.rays == st.rays and S.rel_l2(film, ref) < 1e-6
        ctx.set_option("max_leaf", 2)                      # other build parameters: other key
        ctx.load(sc)
        assert not ctx.accel_cache_hit
        film, st3 = ctx.render()
        assert st3.rays == st.rays and S.rel_l2(film, ref) < 1e-6
    path = host.write_xml(sc, str(tmp_path), "ajax")
    hs = host.HostScene(path)
    f1, s1 = hs.render(0, accel_cache=tmp_path / "ajax.nbbvh")
    f2, s2 = hs.render(0, accel_cache=tmp_path / "ajax.nbbvh")
    assert (tmp_path / "ajax.nbbvh").exists() and s1.rays == s2.rays and S.rel_l2(f2, f1) < 1e-6


def test_progressive_frame_previews_and_final_film(tmp_path, oracle):
    """SURVEY 8f row 4 (progressive preview in place of NoriScreen, ref: src/gui.cpp:120-138): a frame rendered in passes;
    every preview is the film of the samples done so far, the last one is nb_render's frame."""
    sc = S.config_cbox(96, 80, 16, S.INT_PATH_MIS)
    with abi.Context(0) as ctx:
        ctx.load(sc)
        ref, st_ref = ctx.render()
        seen = []
        for done, film, rgb8, st in ctx.render_progressive(5, want_rgb8=True):      # passes of 5, 5, 5, 1 samples
            seen.append(done)
            sc_k = S.config_cbox(96, 80, done, S.INT_PATH_MIS)
            ofilm, ost = oracle.OracleScene(sc_k).render(accel=1)                    # the first `done` sample streams of every pixel
            assert st.rays == ost.rays and S.rel_l2(film, ofilm) <= 1e-4
            assert rgb8.max() == 255 and rgb8.shape == (80, 96, 3)
        assert seen == [5, 10, 15, 16]
        assert st.rays == st_ref.rays and S.rel_l2(film, ref) <= 1e-6
        ctx.set_option("engine", 2)                                                  # the wavefront engine accumulates passes the same way
        for done, film, _, st in ctx.render_progressive(8):
            pass
        ctx.set_option("engine", 0)
        assert st.rays == st_ref.rays and S.rel_l2(film, ref) <= 1e-6
        film2, st2 = ctx.render()                                                    # and a plain frame afterwards is unaffected
        assert st2.rays == st_ref.rays and S.rel_l2(film2, ref) <= 1e-6
    path = host.write_xml(sc, str(tmp_path), "cbox")
    if not os.path.exists(host.CLI_PATH):
        from nori_b200 import build as nb_build
        nb_build.build_host(force=True)
    r = subprocess.run([host.CLI_PATH, path, "--no-gui", "--preview", "6"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    import cv2
    a, b = cv2.imread(str(tmp_path / "cbox.png")), cv2.imread(str(tmp_path / "cbox_preview.png"))
    assert a is not None and b is not None and np.array_equal(a, b)                 # the last preview IS the final image
