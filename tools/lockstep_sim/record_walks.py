"""Host-side model of the lock-step kernel, step 1: walk lengths (node visits + 0.6 per triangle test) of camera rays and
of their ambient-occlusion rays on the Ajax stand-in, through the PRODUCT's SAH hierarchy (nb_debug_build_bvh) with the
device's traversal order, for 25 8x4 patches x 24 samples.  No GPU, no oracle.  Output: an .npy of (Lp, hit, La) per item."""
import sys, time, numpy as np
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__)))))
from nori_b200 import abi, scene as S
f32 = np.float32
m = S.ajax_standin(3)
nodes, tris, info = abi.debug_build_bvh(m.V, m.F, 3, 2048)
refs = nodes[:, 12:14].copy().view(np.int32)
N = nodes.astype(np.float64)
T = tris.astype(np.float64)
print(info)

def walk(o, d, maxt=np.inf, any_hit=False):
    inv = 1.0 / np.where(np.abs(d) > 1e-24, d, 1e-24); ood = o * inv
    best_t, best = maxt, -1
    stack = []; node = 0; steps = 0
    while True:
        if node >= 0:
            n = N[node]; steps += 1
            t0x, t1x = n[0]*inv[0]-ood[0], n[1]*inv[0]-ood[0]; t0y, t1y = n[2]*inv[1]-ood[1], n[3]*inv[1]-ood[1]; t0z, t1z = n[8]*inv[2]-ood[2], n[9]*inv[2]-ood[2]
            c0min = max(min(t0x,t1x), min(t0y,t1y), min(t0z,t1z), 1e-4); c0max = min(max(t0x,t1x), max(t0y,t1y), max(t0z,t1z), best_t)
            t0x, t1x = n[4]*inv[0]-ood[0], n[5]*inv[0]-ood[0]; t0y, t1y = n[6]*inv[1]-ood[1], n[7]*inv[1]-ood[1]; t0z, t1z = n[10]*inv[2]-ood[2], n[11]*inv[2]-ood[2]
            c1min = max(min(t0x,t1x), min(t0y,t1y), min(t0z,t1z), 1e-4); c1max = min(max(t0x,t1x), max(t0y,t1y), max(t0z,t1z), best_t)
            h0, h1 = c0min <= c0max, c1min <= c1max
            r0, r1 = int(refs[node,0]), int(refs[node,1])
            if h0 and h1:
                if c1min < c0min: node = r1; stack.append(r0)
                else: node = r0; stack.append(r1)
            elif h0 or h1: node = r0 if h0 else r1
            else:
                if not stack: break
                node = stack.pop()
            continue
        payload = (~node) & 0xffffffff; first, cnt = payload >> 3, (payload & 7) + 1
        steps += 0.6 * cnt              # a triangle test ~ 0.6 node steps of issue time
        for i in range(first, first+cnt):
            p0, p1, p2 = T[i,0:3], T[i,4:7], T[i,8:11]
            e1, e2 = p1-p0, p2-p0; pv = np.cross(d, e2); det = e1 @ pv
            if -1e-8 < det < 1e-8: continue
            tv = o - p0; u = (tv @ pv)/det
            if u < 0 or u > 1: continue
            q = np.cross(tv, e1); v = (d @ q)/det
            if v < 0 or u+v > 1: continue
            t = (e2 @ q)/det
            if t < 1e-4 or t > best_t: continue
            best_t, best = t, i
            if any_hit: return steps, best, best_t
        if not stack: break
        node = stack.pop()
    return steps, best, best_t

# camera of ajax-ao
sc = S.config_ajax_ao(800, 600, 1, 3)
cam = sc.camera
s2c = np.asarray(cam.s2c, np.float64).reshape(4,4); c2w = np.asarray(cam.c2w, np.float64).reshape(4,4)
def cam_ray(sx, sy):
    p = s2c @ np.array([sx/cam.width, sy/cam.height, 0, 1.0]); p = p[:3]/p[3]; d = p/np.linalg.norm(p)
    o = c2w[:3,3]; dw = c2w[:3,:3] @ d
    return o, dw
rng = np.random.default_rng(1)
def frame(n):
    a = np.array([0.0,1,0]) if abs(n[0])>0.9 else np.array([1.0,0,0]); s = np.cross(n,a); s/=np.linalg.norm(s); t=np.cross(n,s); return s,t
recs = []   # per item: (Lp, hit, La)
t0=time.time()
patches = [(x0,y0) for y0 in range(200, 420, 44) for x0 in range(280, 520, 48)]   # over the bust
for (x0,y0) in patches:
    for s in range(24):                     # 4 samples per pixel -> 128 items per patch
        for ly in range(4):
            for lx in range(8):
                o,d = cam_ray(x0+lx+rng.random(), y0+ly+rng.random())
                Lp, prim, t = walk(o,d)
                if prim < 0: recs.append((Lp, False, 0.0)); continue
                p = o + t*d
                p0,p1,p2 = T[prim,0:3],T[prim,4:7],T[prim,8:11]; n = np.cross(p1-p0,p2-p0); n/=np.linalg.norm(n)
                if n @ d > 0: n = -n
                u1,u2 = rng.random(2); r=np.sqrt(u1); ph=2*np.pi*u2; loc=np.array([r*np.cos(ph), r*np.sin(ph), np.sqrt(max(0,1-u1))])
                sx_,tx_ = frame(n); w = loc[0]*sx_+loc[1]*tx_+loc[2]*n
                La,_,_ = walk(p, w, any_hit=True)
                recs.append((Lp, True, La))
print("rays walked", len(recs), "time", round(time.time()-t0,1))
np.save(__import__('os').environ.get('NB_SIM_RECS', '/tmp/nb_lockstep_recs.npy'), np.array(recs, dtype=np.float64))
