"""Host logic of the candidate-list matchers (no GPU): the Python FrameGrid mirror of Frame::AssignFeaturesToGrid /
GetFeaturesInArea produces the oracle's candidate lists (which are pinned to the unmodified reference functions,
tests/test_ref_frame_pin.py) in the same order."""
import numpy as np

from rumi_slam_b200.matcher import FrameGrid
from rumi_slam_b200.synth import synthetic_frame


def test_frame_grid_equals_oracle(oracle):
    from oracle import match_oracle as M
    M.build()
    rng = np.random.default_rng(1)
    for seed, (w, h) in enumerate([(640, 480), (752, 480), (1241, 376)]):
        k = oracle.extract(synthetic_frame(40 + seed, w, h), nfeatures=1500)[0]
        bounds = (0, 0, w, h) if seed != 1 else (-7, -5, w + 9, h + 4)          # undistorted image bounds can be off-image
        g = FrameGrid(k, bounds)
        for _ in range(200):
            x, y = rng.uniform(-30, w + 30), rng.uniform(-30, h + 30)
            r = float(rng.choice([0.5, 3, 7.5, 20, 100]))
            lv = int(rng.integers(-1, 8))
            lo, hi = (lv - 1, lv) if lv >= 0 else (-1, -1)
            assert np.array_equal(g.features_in_area(x, y, r, lo, hi), M.features_in_area(k, bounds, x, y, r, lo, hi))
        q = np.stack([k["x"], k["y"]], 1)[:300]
        off, idx = g.candidate_lists(q, 12.0, k["octave"][:300] - 1, k["octave"][:300])
        roff, ridx = M.candidate_lists(k, bounds, q, 12.0, k["octave"][:300] - 1, k["octave"][:300])
        assert np.array_equal(off, roff) and np.array_equal(idx, ridx)
