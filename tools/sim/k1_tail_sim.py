"""Queue / tail simulator of the persistent K1 kernel (design aid, CPU only).

Model: 148 SMs x CTAS_PER_SM CTAs x WARPS warps x 32 lanes; a lane integrates one trajectory at a time, taking
`attempts[i]` loop iterations (per-trajectory attempt counts of the C2 ensemble measured with the C port); a warp's
iteration rate depends on how many warps are still active on its SM (FP64-pipe contention, from the measured
1 / 2 / 3 / 4 CTAs-per-SM times of DESIGN.md section 7).  Policies:
  none   -- today's kernel: refill from the global queue while it lasts, then every lane runs out alone
  cta    -- tail phase: warps of a CTA merge their surviving trajectories through a shared-memory pool
  global -- the same with one pool for the whole grid (upper bound for any cross-warp compaction)
Usage: python tools/sim/k1_tail_sim.py ATTEMPTS.npy B [policy] [threshold] [warps_per_cta]
"""
import sys
import numpy as np

SM = 148
GUARD = False


def rate_of(w, total_warps_per_sm=16):
    # per-warp speed relative to a fully loaded SM (16 warps): measured 13.9 / 8.87 / 7.59 / 7.21 ms at 4 / 8 / 12 / 16 warps
    xs = np.array([0, 1, 4, 8, 12, 16, 32])
    sm_throughput = np.array([0, 0.14, 7.21 / 13.9, 7.21 / 8.87, 7.21 / 7.59, 1.0, 1.0])  # relative SM throughput
    thr = np.interp(w, xs, sm_throughput)
    return np.where(w > 0, thr * 16.0 / np.maximum(w, 1), 0.0)


def simulate(att, B, policy="none", thr=16, warps_per_cta=4, ctas_per_sm=4, inner=8, refill_min=2, check_every=8):
    rng = np.random.default_rng(0)
    att = att[:B].astype(np.int64)
    n_cta = SM * ctas_per_sm
    W = n_cta * warps_per_cta
    cta_of = np.arange(W) // warps_per_cta
    sm_of = cta_of % SM  # round-robin CTA placement
    rem = np.zeros((W, 32), dtype=np.int64)
    alive = np.ones(W, dtype=bool)      # warp has not exited
    head = 0
    # initial fill: warps grab 32 each in launch order
    for w in range(W):
        k = min(32, B - head)
        if k > 0:
            rem[w, :k] = att[head:head + k]
            head += k
    progress = np.zeros(W)
    since_check = np.zeros(W, dtype=np.int64)
    pools = {}  # cta (or 0 for global) -> list of remaining counts
    live = np.full(n_cta, warps_per_cta)
    t = 0.0
    lane_iters_useful = att.sum()
    warp_iters = 0
    dt = 0.25
    while alive.any():
        active_w = alive & (rem > 0).any(axis=1)
        # warps that are alive but have nothing: handle exit/refill below
        nact_sm = np.bincount(sm_of[active_w], minlength=SM)
        r = rate_of(nact_sm)[sm_of]
        progress[active_w] += r[active_w] * dt
        t += dt
        step = np.nonzero(active_w & (progress >= 1.0))[0]
        if len(step):
            progress[step] -= 1.0
            rem[step] = np.maximum(rem[step] - 1, 0)
            since_check[step] += 1
            warp_iters += len(step)
        cand = np.nonzero(alive & ((since_check >= inner) | ~(rem > 0).any(axis=1)))[0]
        for w in cand:
            since_check[w] = 0
            idle = rem[w] == 0
            nidle = int(idle.sum())
            if nidle == 0:
                continue
            if head < B:
                if nidle >= refill_min or nidle == 32:
                    k = min(nidle, B - head)
                    slots = np.nonzero(idle)[0][:k]
                    rem[w, slots] = att[head:head + k]
                    head += k
                continue
            # tail phase
            if policy == "none":
                if nidle == 32:
                    alive[w] = False
                continue
            key = int(cta_of[w]) if policy == "cta" else 0
            pool = pools.setdefault(key, [])
            if pool and nidle:
                k = min(nidle, len(pool))
                slots = np.nonzero(idle)[0][:k]
                for sidx in slots:
                    rem[w, sidx] = pool.pop()
                nidle -= k
            a = 32 - nidle
            n_live = live[cta_of[w]] if policy == "cta" else int(alive.sum())
            if a == 0:
                if n_live > 1 or not pool:
                    alive[w] = False
                    live[cta_of[w]] -= 1
                continue
            if policy == "cta" and GUARD:
                # donate only when the other live warps of the CTA have room for these trajectories right now
                ws = np.arange(cta_of[w] * warps_per_cta, (cta_of[w] + 1) * warps_per_cta)
                room = sum(int((rem[x] == 0).sum()) for x in ws if x != w and alive[x]) - len(pool)
                if room < a:
                    continue
            if a <= thr and n_live > 1:
                pool.extend(int(x) for x in rem[w][rem[w] > 0])
                rem[w] = 0
                alive[w] = False
                live[cta_of[w]] -= 1
    return t, warp_iters, lane_iters_useful / (32.0 * warp_iters)


if __name__ == "__main__":
    att = np.load(sys.argv[1])
    B = int(sys.argv[2])
    policy = sys.argv[3] if len(sys.argv) > 3 else "none"
    thr = int(sys.argv[4]) if len(sys.argv) > 4 else 16
    wpc = int(sys.argv[5]) if len(sys.argv) > 5 else 4
    GUARD = len(sys.argv) > 6 and sys.argv[6] == "guard"
    t, wi, eff = simulate(att, B, policy, thr, wpc, ctas_per_sm=16 // wpc)
    ideal = att[:B].sum() / (SM * 16 * 32.0)
    print(f"B={B} policy={policy} thr={thr} warps/cta={wpc}: time {t:.0f} units, ideal {ideal:.0f}, ratio {ideal / t:.3f}, "
          f"lane efficiency {eff:.3f}")
