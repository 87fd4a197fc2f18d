"""GPU: the loop-closing / initialisation matchers ("next" row N1, second batch) through the C-ABI against the
committed goldens (independent Python restatement, oracle/pyref/match_py.py) and the C oracle."""
import numpy as np
import pytest

from conftest import golden_names, load_golden

pytestmark = pytest.mark.gpu
LOOPS = golden_names("loop_pair")


def _views(g):
    from psl_slam_b200 import FrameData
    b = tuple(g["bounds"])
    return FrameData(g["kps1"], g["desc1"], None, b), FrameData(g["kps2"], g["desc2"], None, b)


@pytest.mark.parametrize("name", LOOPS)
def test_search_by_bow_keyframes(orc, name):
    """SearchByBoW(pKF1, pKF2, vpMatches12) — ORBmatcher.cc:522-655."""
    from psl_slam_b200 import ORBmatcher
    g = load_golden(name)
    c1, c2 = (g["nodes1"], g["offs1"], g["idx1"]), (g["nodes2"], g["offs2"], g["idx2"])
    a1, a2 = g["kps1"]["angle"], g["kps2"]["angle"]
    m, n = ORBmatcher(0.75, True).SearchByBoWKeyFrames(g["desc1"], a1, g["valid1"], c1, g["desc2"], a2, g["valid2"], c2)
    assert np.array_equal(m, g["bow12"]) and n == int(g["nbow"])
    mm = ORBmatcher(0.9, False)
    mm.TH_LOW = 80
    m, n = mm.SearchByBoWKeyFrames(g["desc1"], a1, g["valid1"], c1, g["desc2"], a2, g["valid2"], c2)
    assert np.array_equal(m, g["bow12_noori"]) and n == int(g["nbow_noori"])
    # duplicated KF2 descriptors (distance ties, ratio test fails on exact duplicates) + everything valid: vs the oracle
    d2 = g["desc2"].copy()
    d2[1::2] = d2[::2][: len(d2[1::2])]
    ones1, ones2 = np.ones(len(g["desc1"]), np.uint8), np.ones(len(d2), np.uint8)
    m, n = ORBmatcher(0.75, True).SearchByBoWKeyFrames(g["desc1"], a1, ones1, c1, d2, a2, ones2, c2)
    wm, wn = orc.match_bow_kf(g["desc1"], a1, ones1, c1, d2, a2, ones2, c2, 0.75, 50, True)
    assert np.array_equal(m, wm) and n == wn
    # empty sides
    m, n = ORBmatcher().SearchByBoWKeyFrames(g["desc1"][:0], a1[:0], ones1[:0], (c1[0][:0], c1[1][:1], c1[2][:0]),
                                             g["desc2"], a2, g["valid2"], c2)
    assert n == 0 and len(m) == 0


@pytest.mark.parametrize("name", LOOPS)
def test_search_by_sim3(orc, name):
    """SearchBySim3 — ORBmatcher.cc:1102-1326."""
    from psl_slam_b200 import ORBmatcher
    g = load_golden(name)
    k1, k2 = _views(g)
    m, n = ORBmatcher().SearchBySim3(k1, k2, g["q12"], g["desc1"], g["q21"], g["desc2"])
    assert np.array_equal(m, g["sim12"]) and n == int(g["nsim"])
    # wider windows and a lower threshold: against the oracle
    q12, q21 = g["q12"].copy(), g["q21"].copy()
    q12["radius"] *= 2.5
    q21["radius"] *= 2.5
    mm = ORBmatcher()
    mm.TH_HIGH = 60
    m, n = mm.SearchBySim3(k1, k2, q12, g["desc1"], q21, g["desc2"])
    wm, wn = orc.match_sim3((g["kps1"], g["desc1"]), (g["kps2"], g["desc2"]), tuple(g["bounds"]), q12, g["desc1"], q21,
                            g["desc2"], 60)
    assert np.array_equal(m, wm) and n == wn and n > 50
    with pytest.raises(ValueError):
        ORBmatcher().SearchBySim3(k1, k2, g["q12"][:-1], g["desc1"][:-1], g["q21"], g["desc2"])


@pytest.mark.parametrize("name", LOOPS)
def test_fuse_and_projection_sim3_forms(orc, name):
    """Fuse(pKF, Scw, ...) — ORBmatcher.cc:983-1100; SearchByProjection(pKF, Scw, ...) — :290-403."""
    from psl_slam_b200 import ORBmatcher
    g = load_golden(name)
    _, k2 = _views(g)
    bi, bd = ORBmatcher().FuseSearchSim3(k2, g["qfuse"], g["desc1"])
    assert np.array_equal(bi, g["fuse_idx"]) and np.array_equal(bd, g["fuse_dist"])
    a, n = ORBmatcher().SearchByProjectionSim3(k2, g["qproj"], g["desc1"], g["held"])
    assert np.array_equal(a, g["proj_assign"]) and n == int(g["nproj"])


@pytest.mark.parametrize("name", LOOPS)
def test_search_for_initialization(orc, name):
    """SearchForInitialization — ORBmatcher.cc:405-520."""
    from psl_slam_b200 import ORBmatcher
    g = load_golden(name)
    _, f2 = _views(g)
    m, n, pm = ORBmatcher(0.9, True).SearchForInitialization(g["kps1"], g["desc1"], f2, g["prev_matched"], 100)
    assert np.array_equal(m, g["ini12"]) and n == int(g["nini"]) and np.array_equal(pm, g["prev_out"])
    mm = ORBmatcher(0.8, False)
    mm.TH_LOW = 70
    m, n, pm = mm.SearchForInitialization(g["kps1"], g["desc1"], f2, g["prev_matched"], 40)
    assert np.array_equal(m, g["ini12_b"]) and n == int(g["nini_b"]) and np.array_equal(pm, g["prev_out_b"])
    # crowded: every F1 keypoint at level 0 and near-duplicate descriptors, so that later keypoints take matches from
    # earlier ones (the un-matching path, :461-465) — against the oracle
    rng = np.random.default_rng(5)
    k1 = g["kps1"].copy()
    k1["octave"] = 0
    d1 = g["desc1"].copy()
    src = rng.integers(0, 40, len(d1))
    d1[:] = g["desc2"][src]
    for k in range(3):   # 0..3 flipped bits: a later, closer keypoint takes the match of an earlier one
        on = rng.integers(0, 4, len(d1)) > k
        d1[np.arange(len(d1)), 8 * k + rng.integers(0, 8, len(d1))] ^= ((1 << rng.integers(0, 8, len(d1))) * on).astype(np.uint8)
    k2 = g["kps2"].copy()
    k2["octave"][:200] = 0
    from psl_slam_b200 import FrameData
    f2b = FrameData(k2, g["desc2"], None, tuple(g["bounds"]))
    pmi = np.stack([k2["x"][src], k2["y"][src]], 1).astype(np.float32) + rng.normal(0, 3, (len(d1), 2)).astype(np.float32)
    m, n, pm = ORBmatcher(0.9, True).SearchForInitialization(k1, d1, f2b, pmi, 30)
    wm, wn, wpm = orc.match_initialization(k1, d1, pmi, (k2, g["desc2"]), tuple(g["bounds"]), 30, 0.9, 50, True)
    assert np.array_equal(m, wm) and n == wn and np.array_equal(pm, wpm) and 0 < n <= 40
    # nothing to do
    k1["octave"] = 3
    m, n, pm = ORBmatcher().SearchForInitialization(k1, d1, f2b, pmi, 30)
    assert n == 0 and (m == -1).all() and np.array_equal(pm, pmi)
