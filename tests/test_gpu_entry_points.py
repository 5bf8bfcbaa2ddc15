"""GPU parity of the entry points added after round 1's GPU minutes were spent: the `simple` point-light integrator
(NB_INT_SIMPLE, ref: scenes/pa3/ajax-simple.xml:8-11), nb_li_samples and the `ttest` scene object built on it
(ref: src/ttest.cpp:140-176).

First hardware run: the driver's round-1 GPU tier (GPUTEST_r01.json, 13 XPASS).  They are plain tests since round 2.
"""
import os
import subprocess

import numpy as np
import pytest

from nori_b200 import abi, host
from nori_b200 import scene as S

pytestmark = pytest.mark.gpu
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


@pytest.mark.parametrize("integrator", ["normals", "ao", "whitted", "path_mats", "path_ems", "path_mis", "simple"])
def test_li_samples_bit_exact(oracle, integrator):
    """nb_li_samples against oracle.c:orc_li_samples -- per-path fp32 luminances, no atomics involved: bit-exact."""
    if integrator == "simple":
        sc = simple_scene(64, 48, 1)
    elif integrator in ("normals", "ao"):
        cam = S.Camera(S.lookat(**S._AJAX_CAM).astype(np.float32), 30.0, 64, 48)
        sc = S.Scene([S.ajax_standin(2)], cam, S.INTEGRATORS[integrator], 1)
    else:
        sc = S.config_cbox(64, 64, 1, S.INTEGRATORS[integrator])
    sc.seed = 7
    n = 20000
    with abi.Context(0) as ctx:
        ctx.load(sc)
        lum, st = ctx.li_samples(n)
    ref = oracle.OracleScene(sc).li_samples(n)
    assert st.samples == n and st.rays >= n
    assert np.array_equal(lum.view(np.uint32), ref.view(np.uint32))
    assert lum.max() > 0


def test_ttest_object_through_cli(tmp_path):
    """The reference's scenes/pa4/tests/test-mesh.xml, rebuilt from the committed golden meshes: five polygon-light
    scenes under <test type="ttest">, run by the `nori` executable -> 'Passed 5/5 tests.'; a wrong reference fails."""
    from tests import fixtures as FX
    parts = []
    for i in range(1, 6):
        path = host.write_xml(FX.polylum_scene(i, S.INT_PATH_MIS), str(tmp_path), f"poly{i}")
        parts.append(open(path).read().split("\n", 1)[1])           # drop the <?xml ...?> line
    def make_test_file(name, refs):
        p = tmp_path / name
        p.write_text('<test type="ttest">\n<string name="references" value="' + ", ".join("%.7g" % r for r in refs) + '"/>\n' + "".join(parts) + "</test>\n")
        return str(p)
    if not os.path.exists(host.CLI_PATH):
        from nori_b200 import build as nb_build
        nb_build.build_host(force=True)
    good = subprocess.run([host.CLI_PATH, make_test_file("good.xml", FX.POLYLUM_REFS)], capture_output=True, text=True, timeout=600)
    assert good.returncode == 0 and "Passed 5/5 tests." in good.stdout, good.stdout[-2000:] + good.stderr
    wrong = list(FX.POLYLUM_REFS); wrong[2] *= 1.2
    bad = subprocess.run([host.CLI_PATH, make_test_file("bad.xml", wrong)], capture_output=True, text=True, timeout=600)
    assert bad.returncode != 0 and "Passed 4/5 tests." in bad.stdout and "Some tests failed" in bad.stderr


def test_cuda_path_reproduces_golden_extras():
    import os as _os
    from tests.golden.make_oracle_vectors import golden_extras
    Z = np.load(_os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "golden", "oracle_vectors_extra.npz"))
    simple, li = golden_extras()
    with abi.Context(0) as ctx:
        ctx.load(simple)
        film, st = ctx.render()
        assert st.rays == int(Z["rays_simple"][0]) and S.rel_l2(film, Z["film_simple"]) < 1e-5
        ctx.load(li)
        lum, _ = ctx.li_samples(4096)
        assert lum.tobytes() == Z["li_path_mis"].tobytes()


def test_bsdf_batch_bit_exact(oracle):
    """nb_bsdf_sample / nb_bsdf_eval_pdf against the oracle's bsdf_sample / bsdf_eval / bsdf_pdf, all four BSDFs, random
    (wi, xi, wo) including the lower hemisphere: bit for bit."""
    import ctypes as C
    L = oracle.lib()
    rng = np.random.default_rng(17)
    n = 4000
    wi = rng.normal(size=(n, 3)); wi /= np.linalg.norm(wi, axis=1, keepdims=True); wi = wi.astype(np.float32)
    wo = rng.normal(size=(n, 3)); wo /= np.linalg.norm(wo, axis=1, keepdims=True); wo = wo.astype(np.float32)
    xi = rng.random((n, 2)).astype(np.float32)
    bsdfs = [S.diffuse((0.2, 0.5, 0.7)), S.mirror(), S.dielectric(), S.microfacet((0.1, 0.2, 0.15), 0.1, 1.5, 1.000277),
             S.microfacet((0.4, 0.2, 0.3), 0.6, 1.8, 1.3)]
    with abi.Context(0) as ctx:
        for b in bsdfs:
            ob = oracle.bsdf_struct(b)
            d = abi.BsdfDesc(); d.type = int(b.type)
            for k in range(3):
                d.albedo[k] = float(b.albedo[k])
            d.alpha, d.intIOR, d.extIOR, d.ks = ob.alpha, ob.intIOR, ob.extIOR, ob.ks
            got = ctx.bsdf_sample(d, wi, xi)
            ev = ctx.bsdf_eval_pdf(d, wi, wo)
            ref_wo = np.zeros(3, np.float32); ref_w = np.zeros(3, np.float32); eta = C.c_float(); meas = C.c_int()
            ref_ev = np.zeros((n, 4), np.float32)
            for k in range(n):
                L.orc_bsdf_sample(C.byref(ob), oracle._p(wi[k]), oracle._p(xi[k]), oracle._p(ref_wo), C.byref(eta), C.byref(meas), oracle._p(ref_w))
                assert got[k, 3:6].tobytes() == ref_w.tobytes(), (b.type, k)
                if np.any(ref_w != 0):
                    assert got[k, 0:3].tobytes() == ref_wo.tobytes() and int(got[k, 7]) == meas.value, (b.type, k)
                L.orc_bsdf_eval_pdf_batch(C.byref(ob), oracle._p(wi[k]), oracle._p(wo[k]), 1, oracle._p(ref_ev[k:k + 1]))
            assert ev.tobytes() == ref_ev.tobytes(), b.type
            one = ctx.bsdf_sample(d, wi[0], xi)                       # shared-wi form used by the test objects
            assert one[0].tobytes() == got[0].tobytes()


def test_reference_bsdf_fixtures_through_cli(tmp_path):
    """The reference's scenes/pa5/tests/ttest-microfacet.xml and chi2test-microfacet.xml, restated verbatim (they hold no
    meshes), run by the `nori` executable with BSDF::sample / pdf evaluated on the device."""
    from tests import fixtures as FX
    if not os.path.exists(host.CLI_PATH):
        from nori_b200 import build as nb_build
        nb_build.build_host(force=True)
    t = tmp_path / "ttest-microfacet.xml"
    t.write_text('<test type="ttest"><string name="angles" value="' + ", ".join(str(a) for a in FX.MICROFACET_ANGLES) + '"/>'
                 '<string name="references" value="' + ", ".join(str(r) for r in FX.MICROFACET_REFS) + '"/>'
                 '<bsdf type="microfacet"><float name="alpha" value="0.1"/><float name="intIOR" value="1.5"/>'
                 '<float name="extIOR" value="1.000277"/><color name="kd" value="0.1, 0.2, 0.15"/></bsdf></test>')
    r = subprocess.run([host.CLI_PATH, str(t)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "Passed 5/5 tests." in r.stdout, r.stdout[-1500:] + r.stderr
    c = tmp_path / "chi2test-microfacet.xml"
    c.write_text('<test type="chi2test">' + "".join(
        '<bsdf type="microfacet"><float name="alpha" value="%g"/><float name="intIOR" value="%g"/><float name="extIOR" value="%g"/>'
        '<color name="kd" value="%g, %g, %g"/></bsdf>' % (a, i, e, *kd)
        for a, i, e, kd in [(0.1, 1.33, 1.01, (0.0, 0.0, 0.0)), (0.3, 1.5, 1.01, (0.2, 0.1, 0.6)), (0.6, 1.8, 1.3, (0.4, 0.2, 0.3))]) + '</test>')
    r = subprocess.run([host.CLI_PATH, str(c)], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "Passed 15/15 tests." in r.stdout, r.stdout[-1500:] + r.stderr
