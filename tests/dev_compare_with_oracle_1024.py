"""Development aid: step a 1,024-particle set and the oracle side by side and report where they first differ
(weights, matcher results, map cells).  Run on a GPU box: python tests/dev_compare_with_oracle_1024.py (a checker like the tests: the only places that may load oracle/)"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, oracle as O
from thesis_b200.particles import ParticleSet
G = np.load(os.path.join(ROOT, "tests/golden/ref_golden.npz"))
N, K, B = 1024, 30, 180
rng = np.random.default_rng(4)
ps = ParticleSet(N, B, pool_subtiles=60000)
f = O.Filter(N, B, K)
ang = G["intel_angles"]
for _ in range(2):
    ps.set_scan(G["intel_ranges"][0], ang); ps.integrate()
    f.set_scan(G["intel_ranges"][0], ang); f.integrate()
par = (0.002, 0.05, 0.01 * np.pi / 180, 0.05)
for step in (1, 2):
    u = (0.05, 0.0, -0.45)
    ps.motion(1, u, 1.0, par); f.motion(1, u, 1.0, par)
    z = rng.standard_normal((N, K, 3)); u01 = float(rng.random())
    r = G["intel_ranges"][step]
    pp = ps.poses.copy(); pc = ps.covs.copy()
    if step == 2:
        print(" prior pose diff", np.abs(pp - f.pose).max(), "prior cov diff", np.abs(pc - f.cov).max(), "rel", np.abs(pc - f.cov).max() / np.abs(f.cov).max())
    ps.set_scan(r, ang); ps.scan_match()
    mr = ps.match_result()
    ps.weight(z)
    pw = ps.weights.copy()
    ps.integrate(fallback_weights=True)
    if step == 2:
        s_ = O.Scan(r, ang)
        nb = 0
        for i in range(N):
            rx, ry = O.pose_range(f.cov[i])
            o = f.map(i).match(f.pose[i], s_, rx, ry)
            b = mr["best"][i]
            same = (int(b[0]), int(b[1]), int(b[2])) == o["best"] and int(b[3]) == o["M"] and mr["score"][i] == o["score"] \
                and np.array_equal(mr["pose"][i], o["pose"]) and (np.array_equal(mr["cov"][i], o["cov"]) or not o["valid"])
            if not same:
                nb += 1
                if nb < 4:
                    print("  match differs particle", i, b, o["best"], o["M"], mr["score"][i], o["score"], mr["cov"][i].ravel()[[0,1,4,8]], o["cov"].ravel()[[0,1,4,8]], "prior cov diag", np.diag(f.cov[i]), "gpu pose-f.pose", (pp[i]-f.pose[i]))
        print(" match mismatches", nb)
    f.set_scan(r, ang); f.map_update(z)
    gw, ow = ps.weights, np.array(f.weight)
    bad = np.flatnonzero(~np.isclose(gw, ow, rtol=1e-9))
    print("step", step, "bad weights", len(bad), bad[:10], "valid", mr["valid"].sum())
    print(" pose diff max", np.abs(ps.poses - f.pose).max())
    # maps
    nbad = 0
    for i in list(bad[:3]) + [0, 5, 100]:
        for (cx, cy), ref in f.map(int(i)).tiles().items():
            got = ps.export_tile(int(i), cx, cy)
            d = np.argwhere(np.rint(got * 10) != np.rint(ref * 10))
            if len(d):
                nbad += 1
                print("  particle", i, "tile", (cx, cy), "cells differ", len(d), d[:3], got[tuple(d[0])], ref[tuple(d[0])])
    print(" map mismatches", nbad)
    if len(bad):
        i = int(bad[0])
        # recompute the weight on the oracle for this particle from GPU's match result to localise
        s = O.Scan(r, ang)
        print("  particle", i, "gpu w", gw[i], "oracle w", ow[i], "valid", mr["valid"][i])
    did, anc = ps.resample(u01); odid, oanc = f.resample(u01)
    print(" ancestors equal", np.array_equal(anc, oanc))
