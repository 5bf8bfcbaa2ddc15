"""Step 2a: replay the recorded walk lengths through a model of render_kernel's waves (mixed or phased, with a tail cut of
T lanes); cost unit = one node step of the warp; shading / regeneration cost a fixed number of units per wave."""
import numpy as np
recs = np.load(__import__('os').environ.get('NB_SIM_RECS', '/tmp/nb_lockstep_recs.npy'))
Lp, hit, La = recs[:,0], recs[:,1] > 0, recs[:,2]
print("items", len(recs), "hit frac", hit.mean(), "mean Lp %.1f La %.1f" % (Lp.mean(), La[hit].mean()), "p95 Lp %.0f La %.0f max %.0f %.0f" % (np.percentile(Lp,95), np.percentile(La[hit],95), Lp.max(), La[hit].max()))
# one warp processes items of a patch in order: 25 patches x (24 samples x 32 px) -> per patch a stream of 768 items
n_patch = 25; per = len(recs)//n_patch
def simulate(tail, c_shade_p=6.0, c_shade_a=2.0, c_regen=4.0, phased=True):
    """returns (total cost in node-step units, rays, lane-steps useful)"""
    total = 0.0; useful = 0.0; nrays = 0
    for pidx in range(n_patch):
        items = list(range(pidx*per, (pidx+1)*per))
        nxt = 0
        # lane state: kind 0 idle, 1 primary pending (rem), 2 ao pending (rem); item idx
        kind = np.zeros(32, int); rem = np.zeros(32); item = np.zeros(32, int); fresh = np.ones(32, bool)
        while True:
            # regen idle lanes
            for l in range(32):
                if kind[l] == 0 and nxt < len(items):
                    item[l] = items[nxt]; nxt += 1; kind[l] = 1; rem[l] = Lp[item[l]]; fresh[l] = True; nrays += 1
            if not (kind > 0).any(): break
            exhausted = nxt >= len(items)
            carry = (kind > 0) & ~fresh
            if phased:
                any_sh = ((kind == 2) & fresh).any()
                traced = carry | ((kind == 2) if any_sh else (kind == 1))
            else:
                traced = kind > 0
            r = np.where(traced, rem, 0.0)
            T = 0 if exhausted else tail
            srt = np.sort(r[traced])[::-1]
            # wave runs until number of unfinished lanes <= T: duration = (T+1)-th largest remaining (0 if fewer lanes)
            dur = srt[T] if len(srt) > T else (srt[-1] if False else 0.0)
            if T == 0: dur = srt[0]
            if dur <= 0: dur = srt[min(len(srt)-1, 0)] if T == 0 else min(srt[srt > 0]) if (srt > 0).any() else 0.0   # guarantee progress
            useful += np.minimum(r, dur).sum()
            done = traced & (rem <= dur + 1e-9)
            rem = np.where(traced, np.maximum(rem - dur, 0.0), rem)
            fresh[traced & ~done] = False
            shade_cost = 0.0
            if done.any():
                shade_cost = c_shade_p if (kind[done] == 1).any() else 0.0
                shade_cost = max(shade_cost, c_shade_a if (kind[done] == 2).any() else 0.0) if not ((kind[done]==1).any() and (kind[done]==2).any()) else c_shade_p + c_shade_a
            for l in np.flatnonzero(done):
                if kind[l] == 1 and hit[item[l]]:
                    kind[l] = 2; rem[l] = La[item[l]]; fresh[l] = True; nrays += 1
                else:
                    kind[l] = 0
            total += dur + shade_cost + c_regen
    return total, nrays, useful
base = None
for phased in (False, True):
    for tail in (0, 2, 4, 8, 12, 16, 20):
        tot, nr, useful = simulate(tail, phased=phased)
        if base is None: base = tot
        print("phased" if phased else "mixed ", "tail %2d" % tail, "cost %.0f" % tot, "rays/cost %.3f" % (nr/tot), "speedup vs mixed/0: %.3f" % (base/tot), "walk lane util %.3f" % (useful/(32*tot)))
