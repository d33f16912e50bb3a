"""Development aid (CPU, numpy + oracle): how many bitmap scoring passes would other
branch-and-bound schedules of the correlative matcher need?  Emulates the search of
k_match.cu on real searches of the synthetic workload -- exact scores from a dense
numpy evaluation, group bounds from dilated occupancy -- and counts passes for
  flat-8   : the kernel's schedule (8 seeds, 29 groups of 8 with dilation 5, members while the bound can win)
  two-level: groups of 16 (dilation 9) first, their two halves (dilation 5) only while the 16-bound can win
Run: python tests/dev_matcher_bound_study.py [n_searches]
Not collected by pytest.  A checker like the tests: the only places that may load oracle/."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracle as O  # noqa: E402
from thesis_b200 import synth  # noqa: E402

CS, NT, R = 0.05, 14, 260


def window(m, guess):
    """Dense occupancy (tenths > 10) around the guess cell, 3x3-dilated like the matcher's."""
    tiles = m.tiles()
    g0 = np.floor(np.array(guess[:2]) / CS + 1e-9).astype(int)             # good enough for a study
    occ = np.zeros((2 * R + 3, 2 * R + 3), dtype=bool)
    for (cx, cy), cells in tiles.items():
        ix, iy = np.nonzero(np.rint(cells * 10) > 10)
        gx = ix - 400 + int(round(cx / CS)) - g0[0] + R + 1
        gy = iy - 400 + int(round(cy / CS)) - g0[1] + R + 1
        ok = (gx >= 0) & (gx < occ.shape[0]) & (gy >= 0) & (gy < occ.shape[1])
        occ[gx[ok], gy[ok]] = True
    return occ, g0


def dilate(a, r):
    out = np.zeros_like(a)
    for dx in range(-r, r + 1):
        for dy in range(-r, r + 1):
            out |= np.roll(np.roll(a, dx, 0), dy, 1)
    return out


def scores(win, pts, rots, frac):
    """max over the 29x29 translations of the hit count, per rotation."""
    out = np.zeros(len(rots), dtype=int)
    sh = np.arange(-NT, NT + 1)
    for n, th in enumerate(rots):
        c, s = np.cos(th), np.sin(th)
        x = np.floor((c * pts[:, 0] - s * pts[:, 1] + frac[0]) / CS + 0.5).astype(int) + R + 1
        y = np.floor((s * pts[:, 0] + c * pts[:, 1] + frac[1]) / CS + 0.5).astype(int) + R + 1
        ok = (x > NT) & (x < win.shape[0] - NT - 1) & (y > NT) & (y < win.shape[1] - NT - 1)
        x, y = x[ok], y[ok]
        vol = win[(x[:, None, None] + sh[None, :, None]), (y[:, None, None] + sh[None, None, :])].sum(0)
        out[n] = vol.max()
    return out


def study(n_searches=6):
    w = synth.Workload(30, 360)
    f = O.Filter(1, 360, 30)
    rng = np.random.default_rng(3)
    f.set_scan(w.ranges[0], w.angles); f.integrate(); f.integrate()
    step, nk = O.rot_step(), O.rot_count()
    rots = np.arange(-nk, nk + 1) * step
    res = []
    for s in range(1, 30):
        f.motion(1, w.odom[s - 1], w.dt, w.par)
        f.set_scan(w.ranges[s], w.angles)
        if s >= 30 - n_searches:
            guess = f.pose[0].copy()
            m = f.map(0)
            pts = m.match_curr(guess, O.Scan(w.ranges[s], w.angles))
            occ, g0 = window(m, guess)
            frac = guess[:2] - g0 * CS
            b1, b5, b9 = dilate(occ, 1), dilate(occ, 6), dilate(occ, 10)
            exact = scores(b1, pts, rots, frac)
            best = exact.max()
            n = len(rots)
            # flat-8
            g8 = [(g, min(g + 8, n)) for g in range(0, n, 8)]
            ub8 = np.array([scores(b5, pts, rots[[min(a + 4, n - 1)]], frac)[0] for a, b in g8])
            seeds = set(range(nk - 4, nk + 4))
            run_best = exact[list(seeds)].max()
            passes8 = len(seeds) + len(g8)
            for gi in np.argsort(-ub8):
                if ub8[gi] < run_best:
                    break
                for k in range(*g8[gi]):
                    if k in seeds:
                        continue
                    passes8 += 1
                    run_best = max(run_best, exact[k])
            # two-level
            g16 = [(g, min(g + 16, n)) for g in range(0, n, 16)]
            ub16 = np.array([scores(b9, pts, rots[[min(a + 8, n - 1)]], frac)[0] for a, b in g16])
            run_best = exact[list(seeds)].max()
            passes2 = len(seeds) + len(g16)
            for gi in np.argsort(-ub16):
                if ub16[gi] < run_best:
                    break
                a, b = g16[gi]
                for h0 in (a, a + 8):
                    if h0 >= b:
                        continue
                    h1 = min(h0 + 8, b)
                    passes2 += 1
                    ub = scores(b5, pts, rots[[min(h0 + 4, n - 1)]], frac)[0]
                    if ub < run_best:
                        continue
                    for k in range(h0, h1):
                        if k in seeds:
                            continue
                        passes2 += 1
                        run_best = max(run_best, exact[k])
            assert run_best == best
            res.append((s, len(pts), best, passes8, passes2, int((ub8 >= best).sum()), int((ub16 >= best).sum())))
            print("scan %2d  M=%3d best=%3d  passes flat-8 %3d  two-level %3d   groups that survive: %d of %d (8), %d of %d (16)"
                  % (res[-1][:5] + (res[-1][5], len(g8), res[-1][6], len(g16))), flush=True)
        f.map_update(rng.standard_normal((1, 30, 3)))
        f.resample(float(rng.random()))
    a = np.array(res)
    print("mean passes: flat-8 %.1f, two-level %.1f" % (a[:, 3].mean(), a[:, 4].mean()))


if __name__ == "__main__":
    study(int(sys.argv[1]) if len(sys.argv) > 1 else 6)
