import sys, time, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'nav-stack-from-scratch_b200'))
import numpy as np
from nsfs import mapgen, capi
from oracle.oracle import Oracle, ASTAR, DJIKSTRA, THETASTAR

def check_map(H, W, p, seed, inflated=False):
    infl, lo = mapgen.synth_map(H, W, p, seed, inflated)
    P = capi.Planner(H, W); P.set_map(infl, lo)
    ij = np.stack(np.meshgrid(np.arange(H), np.arange(W), indexing='ij'), -1).reshape(-1, 2)
    tp = P.test_pos(ij).reshape(H, W)
    print(f"[pack {H}x{W}] testPos mismatches:", int((tp != mapgen.free_mask(infl, lo)).sum()))
    return infl, lo, P

def check_field(H, W, p, seed, inflated=False, oracle=True):
    infl, lo, P = check_map(H, W, p, seed, inflated)
    src = mapgen.field_source(infl, lo)
    t0 = time.time(); r = P.field(src); t1 = time.time()
    print(f"[field {H}x{W}] status {r['status']} reached {r['reached']} stats {r['stats']} wall {t1-t0:.4f}s")
    for _ in range(3):
        r2 = P.field_device(src); print("   gpu_ms", r2['stats']['gpu_ms'], 'rounds', r2['stats']['rounds'], 'visits', r2['stats']['tile_visits'], 'sweeps', r2['stats']['pops'])
    if oracle:
        O = Oracle(infl, lo)
        t0 = time.time(); o = O.field(src); t1 = time.time()
        go = o['g']; gg = r['g'].astype(np.float64)
        reach_o = go < 1e5; reach_g = gg < 1e5
        rel = np.abs(gg - go)[reach_o] / np.maximum(go[reach_o], 1)
        print(f"   oracle {t1-t0:.3f}s settled {o['settled']} reach mismatch {(reach_o != reach_g).sum()} max rel {rel.max():.3e} gmax {go[reach_o].max():.2f}")
        # path check
        goals = mapgen.sample_free_cells(infl, lo, 5, seed + 3)
        for g in goals:
            fp = P.field_path(g)
            print("   path", tuple(g), "n", fp['n_path'], "cost", fp['cost'], "oracle g", go[g[0], g[1]], "status", fp['status'])
    return P

def check_search(H, W, p, seed, nq=20, inflated=False):
    infl, lo, P = check_map(H, W, p, seed, inflated)
    O = Oracle(infl, lo)
    qs = mapgen.sample_queries(infl, lo, nq, seed)
    for algo, aname, modes in ((ASTAR, 'astar', ('f', 'fg')), (DJIKSTRA, 'djikstra', ('g',)), (THETASTAR, 'thetastar', ('f',))):
        for mode in modes:
            t0 = time.time(); b = P.plan_batch(aname, mode, qs, path_cap=4 * (H + W)); t1 = time.time()
            bad = 0
            for qi, q in enumerate(qs):
                o = O.plan(algo, mode, q[:2], q[2:])
                n = o['path'].shape[0]
                same = (b['status'][qi] == o['status'] and b['g_goal'][qi] == o['g_goal'] and b['path_len'][qi] == n
                        and b['settled'][qi] == o['settled'] and np.array_equal(b['paths'][qi][:n], o['path']))
                if not same and bad < 3:
                    print("   MISMATCH", aname, mode, q, b['status'][qi], o['status'], b['g_goal'][qi], o['g_goal'], b['path_len'][qi], n, b['settled'][qi], o['settled'])
                bad += not same
            print(f"[search {H}x{W} {aname}/{mode}] mismatches {bad}/{nq} batch wall {t1-t0:.3f}s gpu_ms {b['stats']['gpu_ms']:.2f} pops {b['stats']['pops']}")
    # single query
    r = P.plan('astar', 'f', qs[0][:2], qs[0][2:]); o = O.plan(ASTAR, 'f', qs[0][:2], qs[0][2:])
    print("   single plan equal:", np.array_equal(r['path'], o['path']), r['g_goal'] == o['g_goal'], r['stats']['gpu_ms'])
    # fallback
    bads = mapgen.sample_blocked_interior(infl, lo, 50, seed)
    fb = P.find_better_point_batch(bads)
    badc = 0
    for t, bcell in enumerate(bads):
        o = O.find_better_point(bcell)
        badc += not (tuple(fb['cells'][t]) == o['cell'] and fb['status'][t] == o['status'])
    print(f"[fallback] mismatches {badc}/50")
    # los
    rng = np.random.RandomState(seed)
    segs = np.stack([rng.randint(0, H, 500), rng.randint(0, W, 500), rng.randint(0, H, 500), rng.randint(0, W, 500)], 1).astype(np.int32)
    lg = P.los_batch(segs)
    lo_ = np.array([O.is_los(s[:2], s[2:]) for s in segs])
    print("[los] mismatches", int((lg != lo_).sum()), "of 500; frac los", (lo_ == 1).mean())
    return P

if __name__ == '__main__':
    what = sys.argv[1] if len(sys.argv) > 1 else 'all'
    if what in ('all', 'small'):
        check_field(96, 80, 0.10, 3)
        check_field(300, 320, 0.10, 4)
        check_field(384, 384, 0.005, 1, inflated=True)
        check_search(128, 128, 0.10, 7)
        check_search(96, 130, 0.005, 2, inflated=True)
    if what in ('all', 'big'):
        check_field(2048, 2048, 0.10, 5, oracle=True)
