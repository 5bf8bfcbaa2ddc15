"""GPU parity of the `simple` point-light integrator (NB_INT_SIMPLE, ref: scenes/pa3/ajax-simple.xml:8-11).

Added after round 1's GPU minutes were spent: the device code compiles for sm_100a and mirrors oracle.c operation for
operation, but has NOT yet run on hardware.  The tests are therefore xfail(strict=False) -- they report XPASS when they
pass -- and live in the last-collected file so that nothing runs after them.  Remove the marks after their first green
run on a B200.
"""
import numpy as np
import pytest

from nori_b200 import abi
from nori_b200 import scene as S

pytestmark = [pytest.mark.gpu, pytest.mark.xfail(strict=False, reason="first hardware run pending (added without GPU budget)")]
TOL = 1e-4


def simple_scene(w=160, h=120, spp=8, levels=2):
    cam = S.Camera(S.lookat(**S._AJAX_CAM).astype(np.float32), 30.0, w, h)
    return S.Scene([S.ajax_standin(levels)], cam, S.INT_SIMPLE, spp, name="small-ajax-simple",
                   light_pos=(-20.0, 40.0, 20.0), light_energy=(3.76e4, 3.76e4, 3.76e4))   # ajax-simple.xml:9-10


def test_film_parity_simple(oracle):
    sc = simple_scene()
    with abi.Context(0) as ctx:
        ctx.load(sc)
        film, st = ctx.render()
    ofilm, ost = oracle.OracleScene(sc).render(accel=1)
    assert st.samples == ost.samples and st.rays == ost.rays, (st.rays, ost.rays)
    assert S.rel_l2(film, ofilm) <= TOL
    assert film[..., :3].max() > 0


def test_simple_block_seeding_and_missing_light(oracle):
    sc = simple_scene(100, 70, 2)
    sc.seed_mode = S.SEED_PER_BLOCK              # one device thread per block (render_block_mode_kernel<6>)
    with abi.Context(0) as ctx:
        ctx.load(sc)
        film, st = ctx.render()
        ofilm, ost = oracle.OracleScene(sc).render(accel=1)
        assert st.rays == ost.rays and S.rel_l2(film, ofilm) <= TOL
    with abi.Context(0) as ctx:                 # a fresh context has no point light: the C-ABI must refuse, not guess
        sc2 = simple_scene(64, 48, 1)
        sc2.light_pos = None
        ctx.load(sc2)
        with pytest.raises(abi.NoriError, match="nb_set_point_light"):
            ctx.render()
