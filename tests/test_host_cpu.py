"""The C++ host mirror of Nori (NoriObject registry, PropertyList, XML parser, OBJ loader, plugins) -- CPU-only checks.
Error behaviour follows the reference (exceptions with the same messages, ref: include/nori/object.h:132-133,
src/parser.cpp:105-116,183-190, src/scene.cpp:64-76)."""
import os

import numpy as np
import pytest

from nori_b200 import abi, host
from nori_b200 import scene as S


def test_every_hot_path_plugin_is_registered():
    names = ["scene", "obj", "diffuse", "mirror", "dielectric", "microfacet", "area", "independent", "perspective",
             "gaussian", "mitchell", "tent", "box", "normals", "ao", "whitted", "path_mats", "path_ems", "path_mis", "simple", "ttest", "chi2test"]
    assert all(host.is_registered(n) for n in names)
    assert not host.is_registered("photonmapper")


def test_xml_pipeline_matches_python_description(tmp_path):
    sc = S.config_cbox(64, 48, 7, S.INT_PATH_MIS)
    sc.meshes[3] = S.with_(sc.meshes[3], S.microfacet((0.2, 0.2, 0.4), 0.28, 1.7))
    sc.meshes[4] = S.with_(sc.meshes[4], S.dielectric(1.33, 1.0))
    path = host.write_xml(sc, str(tmp_path), "cbox")
    hs = host.HostScene(path)
    info = hs.info()
    assert info == dict(width=64, height=48, border=2, spp=7, n_meshes=6, n_triangles=sc.n_tris, integrator=S.INT_PATH_MIS, seed_mode=0)
    cam = hs.camera()
    assert np.allclose(cam["s2c"], sc.camera.s2c, rtol=2e-6, atol=1e-9)
    assert np.array_equal(cam["c2w"], sc.camera.c2w)
    assert cam["filter_radius"] == 2.0 and np.allclose(cam["filter_table"], sc.filter_table, rtol=1e-6, atol=1e-9)
    for i, m in enumerate(sc.meshes):
        hm = hs.mesh(i)
        assert np.array_equal(hm["V"], m.V) and np.array_equal(hm["F"], m.F)
        assert (hm["N"] is None) == (m.N is None) and (hm["UV"] is None) == (m.UV is None)
        if m.N is not None:
            assert np.allclose(hm["N"], m.N, atol=1e-6)
        assert hm["bsdf"]["type"] == m.bsdf.type
        assert np.allclose(hm["bsdf"]["albedo"], np.float32(m.bsdf.albedo)) or m.bsdf.type in (S.BSDF_MIRROR, S.BSDF_DIELECTRIC)
        assert (hm["emitter"]["type"] == 1) == (m.radiance is not None)
    assert hs.mesh(3)["bsdf"]["ks"] == pytest.approx(0.6) and hs.mesh(3)["bsdf"]["alpha"] == pytest.approx(0.28)
    assert hs.mesh(4)["bsdf"]["intIOR"] == pytest.approx(1.33)
    assert hs.mesh(5)["emitter"]["radiance"] == (40.0, 40.0, 40.0)


def test_obj_loader_rules(tmp_path):
    """v/vt/vn/f, quad split (0,1,2),(3,0,2), dedup on the (p,uv,n) triple in first-use order, toWorld (ref: src/obj.cpp:43-112)."""
    (tmp_path / "q.obj").write_text("v 0 0 0\nv 1 0 0\nv 1 1 0\nv 0 1 0\nvt 0 0\nvt 1 0\nvt 1 1\nvt 0 1\nvn 0 0 1\n"
                                    "f 1/1/1 2/2/1 3/3/1 4/4/1\nf 1/2/1 2/2/1 3/3/1\n")
    (tmp_path / "s.xml").write_text("""<scene><integrator type="normals"/><camera type="perspective"/>
      <mesh type="obj"><string name="filename" value="q.obj"/>
        <transform name="toWorld"><scale value="2,2,2"/><translate value="1,0,0"/></transform></mesh></scene>""")
    hs = host.HostScene(str(tmp_path / "s.xml"))
    m = hs.mesh(0)
    ref = S.load_obj(str(tmp_path / "q.obj"), S.translate([1, 0, 0]) @ S.scale([2, 2, 2]))
    assert np.array_equal(m["F"], ref.F) and m["F"].tolist() == [[0, 1, 2], [3, 0, 2], [4, 1, 2]]
    assert np.allclose(m["V"], ref.V) and np.allclose(m["V"][2], [3, 2, 0])
    assert np.allclose(m["UV"], ref.UV) and np.allclose(m["N"], ref.N)
    assert m["bsdf"]["type"] == S.BSDF_DIFFUSE and m["bsdf"]["albedo"] == (0.5, 0.5, 0.5)    # default BSDF (ref: src/mesh.cpp:23-29)
    i = hs.info()
    assert (i["width"], i["height"], i["spp"]) == (1280, 720, 1)                          # camera / sampler defaults


def test_lookat_rotate_and_defaults(tmp_path):
    (tmp_path / "t.obj").write_text("v 0 0 0\nv 1 0 0\nv 0 1 0\nf 1 2 3\n")
    (tmp_path / "s.xml").write_text("""<?xml version="1.0"?><!-- c --><scene><integrator type="ao"/>
      <camera type="perspective"><transform name="toWorld"><rotate angle="90" axis="0,0,1"/>
        <lookat origin="1,2,3" target="0,0,0" up="0,1,0"/></transform>
        <rfilter type="tent"/></camera>
      <mesh type="obj"><string name="filename" value="t.obj"/></mesh></scene>""")
    cam = host.HostScene(str(tmp_path / "s.xml")).camera()
    c, s = np.cos(np.pi / 2), np.sin(np.pi / 2)
    R = np.array([[c, -s, 0, 0], [s, c, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1]])
    assert np.allclose(cam["c2w"], S.lookat([1, 2, 3], [0, 0, 0], [0, 1, 0]) @ R, atol=1e-6)
    assert cam["filter_radius"] == 1.0 and np.allclose(cam["filter_table"], S.tent_table()[0])


@pytest.mark.parametrize("xml,msg", [
    ('<scene><integrator type="photonmapper"/></scene>', 'A constructor for class "photonmapper" could not be found!'),
    ('<scene><integrator type="ao" foo="1"/></scene>', 'unexpected attribute "foo"'),
    ('<scene><integrator/></scene>', 'missing attribute "type"'),
    ('<scene><camera type="perspective"/></scene>', "No integrator was specified!"),
    ('<scene><integrator type="ao"/><camera type="perspective"/><camera type="perspective"/></scene>', "There can only be one camera per scene!"),
    ('<scene><integrator type="ao"/><camera type="ao"/></scene>', "Unexpectedly constructed an object of type <integrator>"),
    ('<scene><integrator type="ao"/><float name="x" value="1.0abc"/></scene>', 'Could not parse floating point value "1.0abc"'),
    ('<scene><integrator type="ao"/><translate value="1,2,3"/></scene>', "transform nodes can only contain transform operations"),
    ('<float name="x" value="1"/>', "must be a Nori object"),
    ('<scene><integrator type="ao"></scene>', "mismatched closing tag"),
    ('<scene><mesh type="obj"><string name="filename" value="missing.obj"/></mesh></scene>', "Unable to open OBJ file"),
    ('<scene><mesh type="obj"><emitter type="area"/></mesh></scene>', "Property 'radiance' is missing!"),
    ('<scene><integrator type="ao"/><camera type="perspective"><integer name="fov" value="30"/></camera></scene>', "Property 'fov' has the wrong type! (expected <float>)!"),
    ('<scene><integrator type="ao"/><camera type="perspective"><boolean name="x" value="maybe"/></camera></scene>', 'Could not parse boolean value "maybe"'),
    ('<scene><integrator type="ao"/><camera type="perspective"><transform name="toWorld"><matrix value="1 0 0"/></transform></camera></scene>', "Expected 16 values"),
    ('<scene><integrator type="ao"/><sampler type="independent"/><sampler type="independent"/></scene>', "There can only be one sampler per scene!"),
    ('<scene><integrator type="ao"/><camera type="perspective"><rfilter type="box"/><rfilter type="tent"/></camera></scene>', "tried to register multiple reconstruction filters"),
    ('<scene><integrator type="ao"/><bsdf type="diffuse"/></scene>', "Scene::addChild(<bsdf>) is not supported!"),
    ('<scene><integrator type="ao"/><sampler type="independent"><string name="seedMode" value="lattice"/></sampler></scene>', 'unknown seedMode "lattice"'),
])
def test_parser_errors(tmp_path, xml, msg):
    p = tmp_path / "bad.xml"
    p.write_text(xml)
    with pytest.raises(abi.NoriError) as e:
        host.HostScene(str(p))
    assert msg in str(e.value) and "Error while parsing" in str(e.value)


def test_block_generator_spiral_matches_oracle(oracle):
    """BlockGenerator order (ref: src/block.cpp:109-152): host C++ class == oracle restatement; sizes clip at the edges."""
    import ctypes as C
    for W, H in [(768, 768), (800, 600), (100, 70), (31, 33), (1, 1)]:
        hb = host.block_order(W, H)
        n = oracle.lib().orc_block_order(W, H, 32, None)
        xy = np.zeros((n, 2), dtype=np.int32)
        oracle.lib().orc_block_order(W, H, 32, xy.ctypes.data_as(C.c_void_p))
        assert hb.shape[0] == n == ((W + 31) // 32) * ((H + 31) // 32)
        assert np.array_equal(hb[:, :2], xy * 32)
        assert np.all(hb[:, 2] == np.minimum(32, W - hb[:, 0])) and np.all(hb[:, 3] == np.minimum(32, H - hb[:, 1]))
        assert len({(int(a), int(b)) for a, b in hb[:, :2]}) == n
    assert host.block_order(768, 768)[0, :2].tolist() == [384, 384]        # starts at the centre block


def test_obj_binary_cache_round_trip(tmp_path):
    """Opt-in mesh cache: second load comes from <obj>.nbcache and yields identical arrays; the cache
    is invalidated when the OBJ or the toWorld transform changes."""
    m = S.golden_mesh("cbox_sphere1")
    S.write_obj(str(tmp_path / "s.obj"), m)
    xml = """<scene><integrator type="normals"/><camera type="perspective"/>
      <mesh type="obj"><string name="filename" value="s.obj"/><boolean name="cache" value="true"/>%s</mesh></scene>"""
    (tmp_path / "a.xml").write_text(xml % "")
    a = host.HostScene(str(tmp_path / "a.xml")).mesh(0)
    assert (tmp_path / "s.obj.nbcache").exists()
    b = host.HostScene(str(tmp_path / "a.xml")).mesh(0)          # served by the cache
    for k in ("V", "F", "N"):
        assert np.array_equal(a[k], b[k])
    (tmp_path / "b.xml").write_text(xml % '<transform name="toWorld"><scale value="2,2,2"/></transform>')
    c = host.HostScene(str(tmp_path / "b.xml")).mesh(0)          # different transform -> cache rejected and rewritten
    assert np.allclose(c["V"], 2 * a["V"], rtol=1e-6)
    d = host.HostScene(str(tmp_path / "a.xml")).mesh(0)
    assert np.array_equal(d["V"], a["V"])


REF_SCENES = "/root/reference/scenes"
# reference scene files this mirror does not load, and why (everything else under scenes/ must parse)
REF_SCENE_GAPS = {
    "pa2/ajax-normals.xml": "ajax.obj", "pa3/ajax-ao.xml": "ajax.obj", "pa5/ajax/ajax-rough.xml": "ajax.obj",
    "pa5/ajax/ajax-smooth.xml": "ajax.obj", "pa3/ajax-simple.xml": "ajax.obj",   # mesh not shipped with the reference
    # `ttest` / `chi2test` parse, build their children and then need the device (nb_li_samples, nb_bsdf_*): no GPU in this container
    "pa4/tests/test-mesh.xml": "no CUDA device|is not a scene", "pa4/tests/test-mesh-furnace.xml": "no CUDA device|is not a scene",
    "pa5/tests/test-direct.xml": "no CUDA device|is not a scene", "pa5/tests/test-furnace.xml": "no CUDA device|is not a scene",
    "pa5/tests/ttest-microfacet.xml": "no CUDA device|is not a scene", "pa5/tests/chi2test-microfacet.xml": "no CUDA device|is not a scene",
}


@pytest.mark.skipif(not os.path.isdir(REF_SCENES), reason="reference checkout not present (GPU box)")
def test_every_shipped_reference_scene_loads():
    """Scene-file coverage of the XML pipeline: all 25 files under the reference's scenes/ either load (plugins,
    transforms, OBJ meshes, emitters resolved into C-ABI descriptors) or fail for the documented reason."""
    import glob
    seen = 0
    for p in sorted(glob.glob(os.path.join(REF_SCENES, "**", "*.xml"), recursive=True)):
        rel = os.path.relpath(p, REF_SCENES)
        seen += 1
        if rel in REF_SCENE_GAPS:
            with pytest.raises(abi.NoriError, match=REF_SCENE_GAPS[rel]):
                host.HostScene(p)
            continue
        h = host.HostScene(p)
        info = h.info()
        assert info["n_meshes"] >= 1 and info["n_triangles"] >= 2 and info["spp"] >= 1, rel
        h.close()
    assert seen == 25


def test_simple_integrator_xml(tmp_path):
    """ref: scenes/pa3/ajax-simple.xml:8-11 -- `simple` takes a point and a colour; both are mandatory."""
    obj = tmp_path / "tri.obj"
    obj.write_text("v 0 0 0\nv 1 0 0\nv 0 1 0\nf 1 2 3\n")
    body = '<mesh type="obj"><string name="filename" value="tri.obj"/></mesh><camera type="perspective"/>'
    ok = tmp_path / "ok.xml"
    ok.write_text('<scene><integrator type="simple"><point name="position" value="-20, 40, 20"/>'
                  '<color name="energy" value="3.76e4, 3.76e4, 3.76e4"/></integrator>' + body + '</scene>')
    h = host.HostScene(ok)
    assert h.info()["integrator"] == S.INT_SIMPLE
    h.close()
    bad = tmp_path / "bad.xml"
    bad.write_text('<scene><integrator type="simple"><point name="position" value="0,0,0"/></integrator>' + body + '</scene>')
    with pytest.raises(abi.NoriError, match="energy"):
        host.HostScene(bad)


def test_ttest_object_statistics_and_errors(tmp_path, capfd):
    """The `ttest` scene object (ref: src/ttest.cpp:47-189): p-value arithmetic against scipy (the reference gets it from
    the un-vendored `hypothesis` library) and the checks activate() makes before it needs the device."""
    import ctypes as C
    from scipy import stats
    from tests import fixtures as FX
    L = host.lib()
    L.nori_host_students_t_pvalue.restype = C.c_double
    L.nori_host_students_t_pvalue.argtypes = [C.c_double, C.c_double]
    L.nori_host_students_t_test.argtypes = [C.c_double, C.c_double, C.c_double, C.c_int, C.c_double, C.c_int, C.POINTER(C.c_double)]
    for t, dof in [(0.0, 99999), (0.5, 99999), (1.96, 99999), (2.7, 99999), (5.0, 99999), (1.3, 9), (0.2, 1), (12.0, 3)]:
        assert L.nori_host_students_t_pvalue(t, dof) == pytest.approx(2 * stats.t.sf(t, dof), rel=1e-9, abs=1e-300)
    rng = np.random.default_rng(3)
    for _ in range(50):
        mean, var, ref = rng.normal(1.0, 0.01), rng.uniform(0.5, 2.0), 1.0
        n, ntests = int(rng.integers(100, 200000)), int(rng.integers(1, 16))
        pv = C.c_double()
        ok = L.nori_host_students_t_test(mean, var, ref, n, FX.SIGNIFICANCE, ntests, C.byref(pv))
        pv_py = 2.0 * stats.t.sf(abs(mean - ref) * np.sqrt(n) / np.sqrt(var), n - 1)     # as tests/fixtures.py:t_test_pvalue
        assert pv.value == pytest.approx(pv_py, rel=1e-8, abs=1e-300)
        assert bool(ok) == bool(pv_py > FX.sidak(FX.SIGNIFICANCE, ntests))
    obj = tmp_path / "tri.obj"
    obj.write_text("v 0 0 0\nv 1 0 0\nv 0 1 0\nf 1 2 3\n")
    scene = '<scene><integrator type="path_mis"/><mesh type="obj"><string name="filename" value="tri.obj"/></mesh><camera type="perspective"/></scene>'
    bad = tmp_path / "count.xml"
    bad.write_text('<test type="ttest"><string name="references" value="0.5,\n\t 0.25"/>' + scene + '</test>')
    with pytest.raises(abi.NoriError, match="different number of scenes and reference values"):
        host.HostScene(bad)
    bsdf = tmp_path / "bsdf.xml"
    bsdf.write_text('<test type="ttest"><string name="angles" value="0, 10"/><string name="references" value="0.5"/><bsdf type="diffuse"/></test>')
    with pytest.raises(abi.NoriError, match="different number of angles and reference values"):
        host.HostScene(bsdf)
    both = tmp_path / "both.xml"
    both.write_text('<test type="ttest"><string name="angles" value="0"/><string name="references" value="0.5"/><bsdf type="diffuse"/>' + scene + '</test>')
    with pytest.raises(abi.NoriError, match="Cannot test BSDFs and scenes at the same time"):
        host.HostScene(both)
    chi = tmp_path / "chi.xml"
    chi.write_text('<test type="chi2test"><integer name="resolution" value="4"/>' + scene + '</test>')
    with pytest.raises(abi.NoriError, match=r"ChiSquareTest::addChild\(<scene>\) is not supported"):
        host.HostScene(chi)
    capfd.readouterr()


def test_chi2_object_statistics():
    """chi^2 arithmetic of the `chi2test` object (ref: src/chi2test.cpp:171-173 via hypothesis::chi2_test): p-values against
    scipy, pooling decisions against the restatement the oracle tests use (tests/fixtures.py:chi2_pvalue)."""
    import ctypes as C
    from scipy import stats
    from tests import fixtures as FX
    L = host.lib()
    L.nori_host_chi2_pvalue.restype = C.c_double
    L.nori_host_chi2_pvalue.argtypes = [C.c_double, C.c_int]
    dp = C.POINTER(C.c_double)
    L.nori_host_chi2_test.argtypes = [C.c_int, dp, dp, C.c_int, C.c_double, C.c_double, C.c_int, dp]
    for x, dof in [(0.0, 5), (3.2, 1), (10.0, 10), (180.0, 199), (260.0, 199), (400.0, 150), (1.0, 120), (5000.0, 199)]:
        assert L.nori_host_chi2_pvalue(x, dof) == pytest.approx(stats.chi2.sf(x, dof), rel=1e-9, abs=1e-300)
    rng = np.random.default_rng(9)
    for trial in range(40):
        cells, n = 200, 1_000_000
        p = rng.random(cells) ** 6                      # many low-expectation cells -> pooling
        p[rng.integers(0, cells, 10)] = 0.0             # and some empty ones
        p /= p.sum()
        exp = p * n
        obs = rng.multinomial(n, p if trial % 3 else np.roll(p, 1) * 0.999 + 0.001 / cells).astype(np.float64)
        pv = C.c_double()
        ok = L.nori_host_chi2_test(cells, obs.ctypes.data_as(dp), exp.ctypes.data_as(dp), n, 5.0, FX.SIGNIFICANCE, 15, C.byref(pv))
        ref = FX.chi2_pvalue(obs, exp, 5.0)
        assert pv.value == pytest.approx(ref, rel=1e-7, abs=1e-12)
        assert bool(ok) == bool(ref > FX.sidak(FX.SIGNIFICANCE, 15))


def test_write_xml_simple_round_trip(tmp_path):
    cam = S.Camera(S.lookat(origin=[0, 0, 5], target=[0, 0, 0], up=[0, 1, 0]).astype(np.float32), 30.0, 40, 30)
    sc = S.Scene([S.golden_mesh("bunny")], cam, S.INT_SIMPLE, 2, light_pos=(-20.0, 40.0, 20.0), light_energy=(3.76e4, 3.76e4, 3.76e4))
    h = host.HostScene(host.write_xml(sc, str(tmp_path), "simple"))
    i = h.info()
    assert (i["integrator"], i["width"], i["height"], i["spp"]) == (S.INT_SIMPLE, 40, 30, 2)
    h.close()


def test_parser_and_obj_loader_survive_mutated_input(tmp_path):
    """Robustness: randomly damaged scene / OBJ text must end in a NoriException (or a valid scene), never in a crash or
    a hang -- the reference gets this from pugixml and stream extraction; this mirror has its own readers."""
    import random
    rnd = random.Random(1234)
    obj = "v 0 0 0\nv 1 0 0\nv 0 1 0\nv 1 1 0\nvn 0 0 1\nvn 0 1 0\nvt 0 0\nvt 1 1\nf 1/1/1 2/2/1 3/1/2\nf 2/1/1 4/2/2 3/1/1\nf 1/1/1 2/2/2 3/1/1 4/2/2\n"
    xml = ('<?xml version="1.0"?><scene><integrator type="path_mis"/><sampler type="independent"><integer name="sampleCount" value="4"/></sampler>'
           '<camera type="perspective"><transform name="toWorld"><scale value="1,1,1"/><rotate angle="30" axis="0,1,0"/>'
           '<lookat target="0,0,0" origin="0,0,5" up="0,1,0"/><translate value="0, 0, 1"/></transform><float name="fov" value="30"/>'
           '<integer name="width" value="8"/><integer name="height" value="8"/><rfilter type="gaussian"/></camera>'
           '<mesh type="obj"><string name="filename" value="m.obj"/><bsdf type="microfacet"><color name="kd" value="0.2,0.2,0.4"/>'
           '<float name="alpha" value="0.3"/></bsdf><emitter type="area"><color name="radiance" value="1 1 1"/></emitter></mesh><!-- c --></scene>')
    tokens = ['<', '>', '"', "'", '/', '=', '&', ' ', '\n', 'x', '-', '1e99', '<!--', '&amp;', '<a>', '</scene>', '\x00', '//', 'nan', '99999999999']

    def mutate(text):
        s = list(text)
        for _ in range(rnd.randint(1, 5)):
            i = rnd.randrange(len(s)); op = rnd.random()
            if op < 0.35: del s[i:i + rnd.randint(1, 10)]
            elif op < 0.7: s.insert(i, rnd.choice(tokens))
            else: s[i] = rnd.choice('<>"\'/= &x0\n-.')
        return "".join(s)

    loaded = failed = 0
    for k in range(150):
        (tmp_path / "m.obj").write_text(mutate(obj) if k % 2 else obj)
        (tmp_path / "s.xml").write_text(xml if k % 2 else mutate(xml))
        try:
            host.HostScene(tmp_path / "s.xml").close(); loaded += 1
        except abi.NoriError:
            failed += 1
    assert loaded + failed == 150 and failed > 50
