"""Generates tests/golden/pareto.npz by running the UNMODIFIED reference ``utils.simple_cull`` /
``utils.union_rectangles_fastest`` (imported read-only from /root/reference through oracle/ref_harness.py) on seeded
random point sets.  Run in the build container only:  python tests/golden/make_golden_pareto.py"""
import os
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
warnings.filterwarnings("ignore")

from oracle import ref_harness  # noqa: E402


def make_sets(rng, count, P):
    sets = []
    for k in range(count):
        n = rng.randint(1, P + 1)
        pts = rng.rand(n, 4)
        pts[:, 2:] *= 1.25                                   # some infeasible (con > 1)
        if k % 5 == 0:
            pts[:, :2] = np.sort(pts[:, :2], axis=0)         # a chain: front of one point
        if k % 7 == 0 and n > 2:
            pts[1] = pts[0]                                  # an exact duplicate
        if k % 4 == 1:                                       # an anti-chain: every point on the front
            pts[:, 0] = np.sort(pts[:, 0]); pts[:, 1] = np.sort(pts[:, 1])[::-1]
        pts[rng.randint(n), 2:] = 0.5                        # at least one feasible point
        sets.append(pts)
    return sets


def main():
    utils = ref_harness.load_utils("small_bridge")
    rng = np.random.RandomState(20)
    P = 50
    sets = make_sets(rng, 60, P)
    pts_all = np.full((len(sets), P, 4), np.nan)
    counts = np.zeros(len(sets), np.int32)
    flen = np.zeros(len(sets), np.int32)
    fronts = np.full((len(sets), P, 2), np.nan)
    stats = np.zeros((len(sets), 5))
    hv = np.zeros((len(sets), 2))
    refs = np.array([[1.0, 1.0], [0.9, 0.8]])
    for k, pts in enumerate(sets):
        counts[k] = len(pts); pts_all[k, :len(pts)] = pts
        front, max_d, dis_d, p_cd, sum_d, std_cd = utils.simple_cull([list(map(float, p)) for p in pts])
        flen[k] = len(front)
        fronts[k, :len(front)] = np.array(front)[:, :2]
        stats[k] = [max_d, dis_d, p_cd, sum_d, std_cd]
        for r, ref in enumerate(refs):
            hv[k, r] = utils.union_rectangles_fastest([list(f) for f in front], +1, -1, ref_point=list(ref))
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "pareto.npz"), points=pts_all,
                        counts=counts, front_len=flen, fronts=fronts, stats=stats, hv=hv, refs=refs)
    print("wrote pareto.npz:", len(sets), "sets")


if __name__ == "__main__":
    main()
