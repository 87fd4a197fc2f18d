"""GPU: the CUDA matchers through the C-ABI against the oracle and the committed goldens."""
import numpy as np
import pytest

from conftest import golden_names, load_golden

pytestmark = pytest.mark.gpu
PAIRS = golden_names("match_pair")


@pytest.fixture(scope="module")
def matcher():
    from psl_slam_b200 import ORBmatcher
    return ORBmatcher(0.9, True)


def _frame(g):
    from psl_slam_b200 import FrameData
    return FrameData(g["kps_cur"], g["desc_cur"], g["u_right_cur"], tuple(g["bounds"]))


def test_descriptor_distance(matcher, orc):
    rng = np.random.default_rng(0)
    a = rng.integers(0, 256, (500, 32), dtype=np.uint8)
    b = rng.integers(0, 256, (500, 32), dtype=np.uint8)
    b[:5] = a[:5]
    assert np.array_equal(matcher.DescriptorDistance(a, b), orc.descriptor_distance(a, b))
    assert matcher.DescriptorDistance(a[0], a[0]) == 0
    assert matcher.DescriptorDistance(np.zeros(32, np.uint8), np.full(32, 255, np.uint8)) == 256


@pytest.mark.parametrize("name", PAIRS)
def test_knn2_vs_golden_and_oracle(matcher, orc, name):
    from psl_slam_b200 import hamming_knn2
    g = load_golden(name)
    idx, dist = hamming_knn2(matcher.ctx, g["desc_last"][:200], g["desc_cur"][:200])
    assert np.array_equal(idx, g["knn_idx"]) and np.array_equal(dist, g["knn_dist"])
    # full size + forced ties (duplicated train rows): earliest index must win
    t = np.concatenate([g["desc_cur"], g["desc_cur"][:50]])
    idx, dist = hamming_knn2(matcher.ctx, g["desc_last"], t)
    oi, od = orc.hamming_knn2(g["desc_last"], t)
    assert np.array_equal(idx, oi) and np.array_equal(dist, od)
    idx, dist = hamming_knn2(matcher.ctx, g["desc_last"][:3], t[:1])
    assert idx[:, 1].tolist() == [-1, -1, -1] and idx[:, 0].tolist() == [0, 0, 0]


@pytest.mark.parametrize("name", PAIRS)
def test_search_by_projection_last_frame(name):
    from psl_slam_b200 import ORBmatcher
    g = load_golden(name)
    m = ORBmatcher(0.9, True)
    assign, nm = m.SearchByProjectionLastFrame(_frame(g), g["queries"], g["desc_last"])
    assert np.array_equal(assign, g["assign0"]) and nm == int(g["nmatches0"])


@pytest.mark.parametrize("name", PAIRS)
def test_search_by_projection_map_points(name):
    from psl_slam_b200 import ORBmatcher
    g = load_golden(name)
    m = ORBmatcher(0.8, True)
    assign, nm = m.SearchByProjectionMapPoints(_frame(g), g["queries1"], g["desc_last"], g["claimed1"])
    assert np.array_equal(assign, g["assign1"]) and nm == int(g["nmatches1"])


@pytest.mark.parametrize("name", PAIRS)
def test_search_by_bow(name):
    from psl_slam_b200 import ORBmatcher
    g = load_golden(name)
    m = ORBmatcher(0.7, True)
    match, nm = m.SearchByBoW(g["desc_last"], g["angle_last"], g["kf_valid"], (g["kf_nodes"], g["kf_offs"], g["kf_idx"]),
                              g["desc_cur"], g["kps_cur"]["angle"], (g["f_nodes"], g["f_offs"], g["f_idx"]))
    assert np.array_equal(match, g["bow_match"]) and nm == int(g["bow_nmatches"])


def test_projection_randomised_vs_oracle(orc):
    """Random windows / levels / claims, no orientation check and with it; many ties (few distinct descriptors)."""
    from psl_slam_b200 import FrameData, ORBmatcher
    from psl_slam_b200._lib import KP_DTYPE, QUERY_DTYPE
    rng = np.random.default_rng(9)
    for trial in range(6):
        n, nq = int(rng.integers(50, 1200)), int(rng.integers(1, 1500))
        kps = np.zeros(n, KP_DTYPE)
        kps["x"] = rng.uniform(0, 640, n).astype(np.float32)
        kps["y"] = rng.uniform(0, 480, n).astype(np.float32)
        kps["octave"] = rng.integers(0, 8, n)
        kps["angle"] = rng.uniform(0, 360, n).astype(np.float32)
        base = rng.integers(0, 256, (12, 32), dtype=np.uint8)
        desc = base[rng.integers(0, 12, n)] ^ (rng.random((n, 32)) < 0.02).astype(np.uint8)
        ur = np.where(rng.random(n) < 0.7, kps["x"] - rng.uniform(1, 30, n), -1).astype(np.float32)
        q = np.zeros(nq, QUERY_DTYPE)
        q["u"] = rng.uniform(-20, 660, nq)
        q["v"] = rng.uniform(-20, 500, nq)
        q["radius"] = rng.uniform(3, 60, nq)
        lv = rng.integers(0, 8, nq)
        q["min_level"] = np.where(rng.random(nq) < 0.3, -1, lv - 1)
        q["max_level"] = np.where(rng.random(nq) < 0.3, -1, lv + 1)
        q["u_right"] = q["u"] - rng.uniform(1, 30, nq)
        q["angle"] = rng.uniform(0, 360, nq)
        q["flags"] = (rng.random(nq) < 0.9) * 1 + (rng.random(nq) < 0.6) * 2
        qd = base[rng.integers(0, 12, nq)] ^ (rng.random((nq, 32)) < 0.02).astype(np.uint8)
        claimed = (rng.random(n) < 0.2).astype(np.uint8)
        bounds = (0.0, 0.0, 640.0, 480.0)
        fr = FrameData(kps, desc, ur if trial % 2 == 0 else None, bounds)
        for mode, ori in [(0, True), (0, False), (1, False)]:
            m = ORBmatcher(0.8, ori)
            fn = m.SearchByProjectionLastFrame if mode == 0 else m.SearchByProjectionMapPoints
            a, nm = fn(fr, q, qd, claimed)
            oa, onm = orc.match_projection(kps, fr.u_right, desc, bounds, q, qd, claimed, mode, 100, 0.8, ori)
            assert np.array_equal(a, oa) and nm == onm, (trial, mode, ori)


def test_empty_inputs(matcher):
    from psl_slam_b200 import FrameData
    from psl_slam_b200._lib import KP_DTYPE, QUERY_DTYPE
    g = load_golden(PAIRS[0])
    a, nm = matcher.SearchByProjectionLastFrame(_frame(g), g["queries"][:0], g["desc_last"][:0])
    assert nm == 0 and (a == -1).all()
    empty = FrameData(np.zeros(0, KP_DTYPE), np.zeros((0, 32), np.uint8), None, (0, 0, 640, 480))
    a, nm = matcher.SearchByProjectionLastFrame(empty, g["queries"], g["desc_last"])
    assert nm == 0 and len(a) == 0


@pytest.mark.parametrize("name", golden_names("triang_"))
def test_search_for_triangulation_vs_golden_and_oracle(orc, name):
    from psl_slam_b200 import ORBmatcher
    g = load_golden(name)
    kf1 = (g["kps1"], g["ur1"], g["desc1"], g["has_mp1"])
    kf2 = (g["kps2"], g["ur2"], g["desc2"], g["has_mp2"])
    fv1, fv2 = (g["nodes1"], g["offs1"], g["idx1"]), (g["nodes2"], g["offs2"], g["idx2"])
    m = ORBmatcher(0.6, True)
    pairs, m12, nm = m.SearchForTriangulation(kf1, fv1, kf2, fv2, g["F12"], (float(g["ex"]), float(g["ey"])), g["scale"],
                                              g["sigma2"], bool(g["only_stereo"]))
    assert np.array_equal(m12, g["matches12"]) and nm == int(g["nmatches"]) and len(pairs) == nm
    # without the rotation check, and with a shifted epipole: against the oracle
    m2 = ORBmatcher(0.6, False)
    for ex, ey, only in ((float(g["ex"]), float(g["ey"]), False), (320.0, 240.0, False), (100.0, 50.0, True)):
        want, wn = orc.match_triangulation(kf1, fv1, kf2, fv2, g["F12"], ex, ey, g["scale"], g["sigma2"], only, False, 50)
        _, got, gn = m2.SearchForTriangulation(kf1, fv1, kf2, fv2, g["F12"], (ex, ey), g["scale"], g["sigma2"], only)
        assert np.array_equal(got, want) and gn == wn


@pytest.mark.parametrize("name", golden_names("fuse_"))
def test_fuse_search_vs_golden_and_oracle(orc, name):
    from psl_slam_b200 import FrameData, ORBmatcher
    g = load_golden(name)
    m = ORBmatcher()
    bi, bd = m.FuseSearch(FrameData(g["kps"], g["desc"], g["u_right"], tuple(g["bounds"])), g["queries"], g["qdesc"],
                          g["inv_sigma2"])
    assert np.array_equal(bi, g["best_idx"]) and np.array_equal(bd, g["best_dist"])
    # mono keyframe (no right coordinates) and wider windows: against the oracle
    q = g["queries"].copy()
    q["radius"] *= 3
    want_i, want_d = orc.match_fuse(g["kps"], None, g["desc"], tuple(g["bounds"]), q, g["qdesc"], g["inv_sigma2"], 50)
    bi, bd = m.FuseSearch(FrameData(g["kps"], g["desc"], None, tuple(g["bounds"])), q, g["qdesc"], g["inv_sigma2"])
    assert np.array_equal(bi, want_i) and np.array_equal(bd, want_d)


@pytest.mark.parametrize("name", golden_names("reloc_"))
def test_search_by_projection_keyframe_vs_golden(name):
    """SearchByProjection(CurrentFrame, pKF, sAlreadyFound, th, ORBdist) — ORBmatcher.cc:1472-1599."""
    from psl_slam_b200 import FrameData, ORBmatcher
    g = load_golden(name)
    cur = FrameData(g["kps_cur"], g["desc_cur"], None, tuple(g["bounds"]))
    a, n = ORBmatcher(0.9, True).SearchByProjectionKeyFrame(cur, g["queries"], g["desc_kf"], g["held"], int(g["orb_dist"]))
    assert np.array_equal(a, g["assign"]) and n == int(g["nmatches"])
