"""CPU: the matcher oracle (oracle/c/orc_match.cpp) against the goldens written by the independent
Python restatement + cv2.BFMatcher (oracle/pyref/match_py.py)."""
import numpy as np
import pytest

from conftest import golden_names, load_golden

PAIRS = golden_names("match_pair")


def test_descriptor_distance_known_answers(orc):
    a = np.zeros((3, 32), np.uint8)
    b = np.zeros((3, 32), np.uint8)
    b[1] = 0xFF
    b[2, 5] = 0b10110000
    assert orc.descriptor_distance(a, b).tolist() == [0, 256, 3]


@pytest.mark.parametrize("name", PAIRS)
def test_knn2_vs_cv2(orc, name):
    g = load_golden(name)
    idx, dist = orc.hamming_knn2(g["desc_last"][:200], g["desc_cur"][:200])
    assert np.array_equal(idx, g["knn_idx"]) and np.array_equal(dist, g["knn_dist"])


def test_knn2_fewer_than_two_train_rows(orc):
    q = np.arange(64, dtype=np.uint8).reshape(2, 32)
    idx, dist = orc.hamming_knn2(q, q[:1])
    assert idx[:, 1].tolist() == [-1, -1] and idx[:, 0].tolist() == [0, 0] and dist[0, 0] == 0
    idx, dist = orc.hamming_knn2(q, q[:0])
    assert (idx == -1).all()


@pytest.mark.parametrize("name", PAIRS)
def test_search_by_projection_modes(orc, name):
    g = load_golden(name)
    a0, n0 = orc.match_projection(g["kps_cur"], g["u_right_cur"], g["desc_cur"], g["bounds"], g["queries"],
                                  g["desc_last"], None, 0, 100, 0.9, True)
    assert np.array_equal(a0, g["assign0"]) and n0 == int(g["nmatches0"])
    a1, n1 = orc.match_projection(g["kps_cur"], g["u_right_cur"], g["desc_cur"], g["bounds"], g["queries1"],
                                  g["desc_last"], g["claimed1"], 1, 100, 0.8, False)
    assert np.array_equal(a1, g["assign1"]) and n1 == int(g["nmatches1"])


@pytest.mark.parametrize("name", PAIRS)
def test_search_by_bow(orc, name):
    g = load_golden(name)
    m, n = orc.match_bow(g["desc_last"], g["angle_last"], g["kf_valid"], (g["kf_nodes"], g["kf_offs"], g["kf_idx"]),
                         g["desc_cur"], g["kps_cur"]["angle"], (g["f_nodes"], g["f_offs"], g["f_idx"]), 0.7, 50, True)
    assert np.array_equal(m, g["bow_match"]) and n == int(g["bow_nmatches"])


def test_empty_inputs(orc):
    g = load_golden(PAIRS[0])
    a, n = orc.match_projection(g["kps_cur"], None, g["desc_cur"], g["bounds"], g["queries"][:0], g["desc_last"][:0],
                                None, 0)
    assert n == 0 and (a == -1).all()
    a, n = orc.match_projection(g["kps_cur"][:0], None, g["desc_cur"][:0], g["bounds"], g["queries"], g["desc_last"],
                                None, 0)
    assert n == 0 and len(a) == 0


@pytest.mark.parametrize("name", golden_names("triang_"))
def test_search_for_triangulation(orc, name):
    g = load_golden(name)
    m12, nm = orc.match_triangulation((g["kps1"], g["ur1"], g["desc1"], g["has_mp1"]), (g["nodes1"], g["offs1"], g["idx1"]),
                                      (g["kps2"], g["ur2"], g["desc2"], g["has_mp2"]), (g["nodes2"], g["offs2"], g["idx2"]),
                                      g["F12"], float(g["ex"]), float(g["ey"]), g["scale"], g["sigma2"],
                                      bool(g["only_stereo"]), True, 50)
    assert np.array_equal(m12, g["matches12"]) and nm == int(g["nmatches"]) and nm > 20


@pytest.mark.parametrize("name", golden_names("fuse_"))
def test_fuse_search(orc, name):
    g = load_golden(name)
    bi, bd = orc.match_fuse(g["kps"], g["u_right"], g["desc"], tuple(g["bounds"]), g["queries"], g["qdesc"], g["inv_sigma2"], 50)
    assert np.array_equal(bi, g["best_idx"]) and np.array_equal(bd, g["best_dist"]) and (bi >= 0).sum() > 100


@pytest.mark.parametrize("name", golden_names("reloc_"))
def test_search_by_projection_keyframe(orc, name):
    """The relocalization overload (ORBmatcher.cc:1472-1599) is mode 0 of the projection matcher with no right-coordinate
    gate, every match claiming its keypoint, the pre-filled keypoints as claimed_in and ORBdist as threshold; the golden
    comes from a restatement written from that overload alone."""
    g = load_golden(name)
    q = g["queries"].copy()
    q["flags"] = np.where(q["flags"] & 1, 3, 0).astype(np.uint32)
    a, n = orc.match_projection(g["kps_cur"], None, g["desc_cur"], tuple(g["bounds"]), q, g["desc_kf"], g["held"], 0,
                                int(g["orb_dist"]), 0.9, True)
    assert np.array_equal(a, g["assign"]) and n == int(g["nmatches"]) and n > 100


def _kp_array(xy):
    from psl_slam_b200._lib import KP_DTYPE
    k = np.zeros(len(xy), KP_DTYPE)
    k["x"], k["y"] = xy[:, 0], xy[:, 1]
    k["size"], k["angle"], k["octave"], k["class_id"] = 31.0, 12.5, 2, -1
    return k


@pytest.mark.parametrize("cam", ["tum1", "tum2", "k1only", "strong"])
def test_undistort_keypoints_vs_cv2_golden(orc, cam):
    """Frame::UndistortKeyPoints / ComputeImageBounds against the real cv2.undistortPoints (bit-exact floats)."""
    from psl_slam_b200 import make_distortion
    g = load_golden("undistort")
    d = make_distortion(*[float(v) for v in g[f"cam_{cam}"]])
    for src, want in ((g["pts_kp"], g[f"kp_{cam}"]), (g["pts_rnd"], g[f"rnd_{cam}"])):
        k = _kp_array(src)
        out = orc.undistort_keypoints(k, d)
        assert np.array_equal(np.stack([out["x"], out["y"]], 1), want)
        for f in ("size", "angle", "response", "octave", "class_id"):
            assert np.array_equal(out[f], k[f])
    assert np.array_equal(np.array(orc.image_bounds(640, 480, d), np.float32), g[f"bounds_{cam}"])
    # k1 == 0: copied keypoints, image rectangle (Frame.cc:1064-1068, 1156-1162)
    z = make_distortion(500, 500, 320, 240, 0.0, 0.3, 0.01, 0.01, 0.1)
    k = _kp_array(g["pts_rnd"])
    assert orc.undistort_keypoints(k, z).tobytes() == k.tobytes()
    assert orc.image_bounds(640, 480, z) == (0.0, 0.0, 640.0, 480.0)


# ---- N1, second batch: loop closing + monocular initialisation (loop_pair* goldens, oracle/pyref/match_py.py) ----
LOOPS = golden_names("loop_pair")


def loop_proj_queries(g):
    """The qproj fuse-style queries of a loop golden as psl_proj_query records for mode 0 (see psl_frontend.h)."""
    from psl_slam_b200._lib import QUERY_DTYPE
    qp = g["qproj"]
    q = np.zeros(len(qp), QUERY_DTYPE)
    q["u"], q["v"], q["radius"] = qp["u"], qp["v"], qp["radius"]
    q["min_level"], q["max_level"] = qp["pred_level"] - 1, qp["pred_level"]
    q["flags"] = np.where(qp["flags"] & 1, 3, 0).astype(np.uint32)
    return q


@pytest.mark.parametrize("name", LOOPS)
def test_search_by_bow_keyframes(orc, name):
    g = load_golden(name)
    c1, c2 = (g["nodes1"], g["offs1"], g["idx1"]), (g["nodes2"], g["offs2"], g["idx2"])
    m, n = orc.match_bow_kf(g["desc1"], g["kps1"]["angle"], g["valid1"], c1, g["desc2"], g["kps2"]["angle"], g["valid2"],
                            c2, 0.75, 50, True)
    assert np.array_equal(m, g["bow12"]) and n == int(g["nbow"]) and n > 100
    m, n = orc.match_bow_kf(g["desc1"], g["kps1"]["angle"], g["valid1"], c1, g["desc2"], g["kps2"]["angle"], g["valid2"],
                            c2, 0.9, 80, False)
    assert np.array_equal(m, g["bow12_noori"]) and n == int(g["nbow_noori"])


@pytest.mark.parametrize("name", LOOPS)
def test_search_by_sim3(orc, name):
    g = load_golden(name)
    m, n = orc.match_sim3((g["kps1"], g["desc1"]), (g["kps2"], g["desc2"]), tuple(g["bounds"]), g["q12"], g["desc1"],
                          g["q21"], g["desc2"], 100)
    assert np.array_equal(m, g["sim12"]) and n == int(g["nsim"]) and n > 100


@pytest.mark.parametrize("name", LOOPS)
def test_fuse_and_projection_sim3_forms(orc, name):
    g = load_golden(name)
    bi, bd = orc.match_fuse(g["kps2"], None, g["desc2"], tuple(g["bounds"]), g["qfuse"], g["desc1"], None, 50)
    assert np.array_equal(bi, g["fuse_idx"]) and np.array_equal(bd, g["fuse_dist"]) and (bi >= 0).sum() > 300
    a, n = orc.match_projection(g["kps2"], None, g["desc2"], tuple(g["bounds"]), loop_proj_queries(g), g["desc1"],
                                g["held"], 0, 50, 0.9, False)
    assert np.array_equal(a, g["proj_assign"]) and n == int(g["nproj"]) and n > 200


@pytest.mark.parametrize("name", LOOPS)
def test_search_for_initialization(orc, name):
    g = load_golden(name)
    m, n, pm = orc.match_initialization(g["kps1"], g["desc1"], g["prev_matched"], (g["kps2"], g["desc2"]),
                                        tuple(g["bounds"]), 100, 0.9, 50, True)
    assert np.array_equal(m, g["ini12"]) and n == int(g["nini"]) and np.array_equal(pm, g["prev_out"]) and n > 100
    m, n, pm = orc.match_initialization(g["kps1"], g["desc1"], g["prev_matched"], (g["kps2"], g["desc2"]),
                                        tuple(g["bounds"]), 40, 0.8, 70, False)
    assert np.array_equal(m, g["ini12_b"]) and n == int(g["nini_b"]) and np.array_equal(pm, g["prev_out_b"])
    # no level-0 keypoint searches, nothing to search in
    k = g["kps1"].copy()
    k["octave"] = 1
    m, n, pm = orc.match_initialization(k, g["desc1"], g["prev_matched"], (g["kps2"], g["desc2"]), tuple(g["bounds"]))
    assert n == 0 and (m == -1).all() and np.array_equal(pm, g["prev_matched"])
    m, n, pm = orc.match_initialization(g["kps1"], g["desc1"], g["prev_matched"], (g["kps2"][:0], g["desc2"][:0]),
                                        tuple(g["bounds"]))
    assert n == 0 and (m == -1).all()
