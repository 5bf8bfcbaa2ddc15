"""Step 2b: the same walk lengths through a persistent traversal kernel with dynamic fetch (refill once >= thresh lanes are
idle) against densely packed lock-step waves -- camera rays and occlusion rays separately."""
import numpy as np
recs = np.load(__import__('os').environ.get('NB_SIM_RECS', '/tmp/nb_lockstep_recs.npy'))
Lp, hit, La = recs[:,0], recs[:,1] > 0, recs[:,2]
def dynfetch(L, thresh, c_fetch, c_finish=1.0):
    """Aila-Laine persistent traversal with dynamic fetch: lanes walk independently (1 unit per node step for the warp
    as long as any lane is active); when >= thresh lanes are idle (or all), the warp spends c_fetch to refill them.
    c_finish: cost of the result-handling code executed when lanes finish (amortised per refill round)."""
    L = list(L); nxt = 0; rem = np.zeros(32); total = 0.0; useful = 0.0
    n = len(L)
    while True:
        idle = rem <= 0
        if nxt < n and (idle.sum() >= thresh or idle.all()):
            k = min(int(idle.sum()), n - nxt)
            idx = np.flatnonzero(idle)[:k]
            rem[idx] = L[nxt:nxt+k]; nxt += k
            total += c_fetch + c_finish
            continue
        if (rem > 0).sum() == 0: break
        # advance until next event: either a lane finishes such that idle count reaches thresh, or all finish
        act = rem[rem > 0]
        srt = np.sort(act)
        need = max(1, thresh - int((rem <= 0).sum())) if nxt < n else len(srt)
        need = min(need, len(srt))
        dur = srt[need-1]
        useful += np.minimum(rem[rem>0], dur).sum()
        rem = np.where(rem > 0, rem - dur, rem)
        total += dur
    return total, useful
for name, L, cf in (("primary", Lp, 4.0), ("ao", La[hit], 1.5)):
    lock = 0.0
    for i in range(0, len(L) - 31, 32): lock += L[i:i+32].max()
    print(name, "lock-step waves: cost %.0f util %.3f" % (lock, L[:len(L)//32*32].sum()/(32*lock)))
    for th in (1, 2, 4, 8, 12, 16):
        tot, useful = dynfetch(L, th, cf)
        print("   thresh %2d cost %.0f util %.3f speedup vs lockstep %.2f" % (th, tot, useful/(32*tot), lock/tot))
