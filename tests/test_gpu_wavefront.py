"""GPU parity of the wavefront engine (nori_b200/csrc/nb_wave.cuh, nb_set_option("engine", 2)): the same per-path arithmetic as
the fused kernel (shade<INTEG>, begin_path), rays streamed through device queues and traced by persistent warps with dynamic
fetch.  Same ray counts, same film (<= 1e-4 rel-L2 against the oracle; ~1e-7 in practice: only the order of the film atomics
differs), for every integrator, for pools far smaller than the frame (many iterations, slots reused), ragged image sizes
and sample counts that do not fill the last 8-sample chunk."""
import numpy as np
import pytest

from nori_b200 import abi
from nori_b200 import scene as S
from tests.test_gpu_entry_points import simple_scene

pytestmark = pytest.mark.gpu
TOL = 1e-4


def _case(name):
    if name in ("ao", "normals"):
        cam = S.Camera(S.lookat(**S._AJAX_CAM).astype(np.float32), 30.0, 170, 123)       # ragged: 170 = 5*32 + 10, 123 = 3*32 + 27
        return S.Scene([S.ajax_standin(2)], cam, S.INTEGRATORS[name], 13 if name == "ao" else 3, name=f"wave-{name}")
    if name == "simple":
        return simple_scene(160, 120, 8)
    if name == "microfacet":
        sc = S.config_cbox(64, 64, 16, S.INT_PATH_MIS)
        sc.meshes[3] = S.with_(sc.meshes[3], S.microfacet((0.2, 0.2, 0.4), 0.28, 1.7))
        sc.meshes[4] = S.with_(sc.meshes[4], S.dielectric())
        return sc
    if name == "specular":
        sc = S.config_cbox(64, 64, 16, S.INT_WHITTED)
        sc.meshes[3] = S.with_(sc.meshes[3], S.mirror())
        sc.meshes[4] = S.with_(sc.meshes[4], S.dielectric())
        return sc
    return S.config_cbox(72, 56, 11, S.INTEGRATORS[name])


@pytest.mark.parametrize("case", ["normals", "ao", "whitted", "path_mats", "path_ems", "path_mis", "simple", "microfacet", "specular"])
def test_wavefront_engine_parity(oracle, case):
    sc = _case(case)
    ofilm, ost = oracle.OracleScene(sc).render(accel=1)
    with abi.Context(0) as ctx:
        ctx.load(sc)
        film0, st0 = ctx.render()
        assert st0.rays == ost.rays and S.rel_l2(film0, ofilm) <= TOL
        ctx.set_option("engine", 2)
        for pool, tail, check in ((1 << 21, 20, 4), (4096, 8, 1), (640, 0, 3)):
            ctx.set_option("wf_pool", pool); ctx.set_option("occ_tail", tail); ctx.set_option("wf_check", check)
            film, st = ctx.render()
            assert st.samples == ost.samples and st.rays == ost.rays, (case, pool, st.rays, ost.rays)
            assert st.hits_shaded == st0.hits_shaded
            assert S.rel_l2(film, ofilm) <= TOL, (case, pool, tail)
        ctx.set_option("count", 1)                       # instrumented instantiation: same node / triangle totals as the fused kernel's
        ctx.set_option("engine", 0); _, c0 = ctx.render()
        ctx.set_option("engine", 2); _, c2 = ctx.render()
        ctx.set_option("count", 0)
        assert c2.rays == c0.rays and c2.tri_tests > 0 and c2.node_visits > 0
        ctx.set_option("engine", 0)


def test_wavefront_engine_tiles_and_block_seeding(oracle):
    """Tile shards (multi-GPU slices) through the wavefront engine, and the fall-back to the fused kernel for the reference's
    sequential per-block sampler streams."""
    sc = _case("ao")
    with abi.Context(0) as ctx:
        ctx.load(sc)
        full, st_full = ctx.render()
        ctx.set_option("engine", 2)
        acc = np.zeros_like(full); rays = 0
        for r in range(3):
            ctx.set_tiles(r, 3)
            part, st = ctx.render()
            acc += part; rays += st.rays
        ctx.set_tiles(0, 1)
        assert rays == st_full.rays and S.rel_l2(acc, full) < 1e-6
        sc.seed_mode = S.SEED_PER_BLOCK; sc.spp = 2
        ctx.configure(sc)
        film, st = ctx.render()
        ofilm, ost = oracle.OracleScene(sc).render(accel=1)
        assert st.rays == ost.rays and S.rel_l2(film, ofilm) <= TOL
