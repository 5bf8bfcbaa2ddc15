"""Drives a sanitizer-instrumented build of libnori_host.so: every reference scene (if present), the block spiral, and
300 mutated scene files (tools/sanitizers/run.sh)."""
import sys, os, glob, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from nori_b200 import host
host.LIB_PATH = os.environ.get('NB_SAN_HOST', '/tmp/nb_san/libnori_host.so')
L = host.lib()
print("loaded", host.LIB_PATH)
n_ok = n_err = 0
for p in sorted(glob.glob('/root/reference/scenes/**/*.xml', recursive=True)):
    try:
        h = host.HostScene(p); h.info(); h.close(); n_ok += 1
    except Exception as e:
        n_err += 1
print("reference scenes", n_ok, n_err)
print(host.block_order(800, 600).shape, host.block_order(33, 1).shape)
# fuzz
random.seed(99)
d = '/tmp/nb_san/w'; os.makedirs(d, exist_ok=True)
open(d + '/tri.obj', 'w').write('v 0 0 0\nv 1 0 0\nv 0 1 0\nvn 0 0 1\nvt 0 0\nf 1/1/1 2/1/1 3/1/1\n')
base = '''<?xml version="1.0"?><scene><integrator type="path_mis"/><sampler type="independent"><integer name="sampleCount" value="4"/></sampler>
<camera type="perspective"><transform name="toWorld"><scale value="1,1,1"/><rotate angle="30" axis="0,1,0"/><lookat target="0,0,0" origin="0,0,5" up="0,1,0"/><translate value="0, 0, 1"/></transform><float name="fov" value="30"/><integer name="width" value="8"/><integer name="height" value="8"/><rfilter type="gaussian"/></camera>
<mesh type="obj"><string name="filename" value="tri.obj"/><bsdf type="microfacet"><color name="kd" value="0.2,0.2,0.4"/><float name="alpha" value="0.3"/></bsdf><emitter type="area"><color name="radiance" value="1 1 1"/></emitter></mesh><!-- c --></scene>'''
ok = err = 0
for it in range(300):
    s = list(base)
    for _ in range(random.randint(1, 6)):
        op = random.random(); i = random.randrange(len(s))
        if op < 0.3: del s[i:i + random.randint(1, 12)]
        elif op < 0.6: s.insert(i, random.choice(['<', '>', '"', "'", '/', '=', '&', ' ', '\n', 'x', '-', '1e99', '<!--', '&amp;', '<a>', '</scene>', '\x00']))
        elif op < 0.8: s[i] = random.choice('<>"\'/= &x0\n')
        else:
            j = random.randrange(len(s)); s[i:i] = s[j:j + random.randint(1, 40)]
    p = d + '/m.xml'; open(p, 'w').write(''.join(s))
    try:
        h = host.HostScene(p); h.close(); ok += 1
    except Exception as e:
        err += 1
print("fuzz", ok, err)
