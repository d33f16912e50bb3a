"""Known-answer vectors of OUR restated NDT stage (oracle/rbpf_oracle.c ndt_refine) -- not reference
outputs: MathWorks matchScans is not available (parity unpinned).  They pin the restatement itself, so
that a later change of the oracle or of the CUDA stage that alters results is noticed.

    python tests/golden/make_ndt_regression.py      ->  tests/golden/ndt_oracle_kat.npz
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracle as O  # noqa: E402


def cases():
    G = np.load(os.path.join(HERE, "ref_golden.npz"))
    m = O.Map()
    for p, si in zip(G["upd_seed_poses"], G["upd_seed_scan_idx"]):
        m.update(p, O.Scan(G["intel_ranges"][int(si)], G["intel_angles"]))
    s = O.Scan(G["intel_ranges"][4], G["intel_angles"])
    rng = np.random.default_rng(2024)
    guesses = np.array([0.6, 0.15, 0.1]) + rng.normal(0, [0.15, 0.15, 0.08], (10, 3))
    windows = rng.uniform(0.25, 0.7, 10)
    return m, s, guesses, windows


def run():
    m, s, guesses, windows = cases()
    old = O.set_refine(True)
    try:
        out = [m.match(g, s, float(w), float(w)) for g, w in zip(guesses, windows)]
    finally:
        O.set_refine(old)
    return dict(guesses=guesses, windows=windows,
                pose=np.array([o["pose"] for o in out]), score=np.array([o["score"] for o in out]),
                valid=np.array([o["valid"] for o in out]), evals=np.array([o["ndt_evals"] for o in out]),
                accepted=np.array([o["ndt_accepted"] for o in out]), best=np.array([o["best"] for o in out]))


if __name__ == "__main__":
    r = run()
    np.savez_compressed(os.path.join(HERE, "ndt_oracle_kat.npz"), **r)
    print("valid", r["valid"].sum(), "accepted", r["accepted"].sum(), "evals", r["evals"])
