"""CPU: the line-matcher oracle (oracle/c/orc_linematch.cpp) against goldens made by the independent Python
restatement over the real cv2.BFMatcher (oracle/pyref/linematch_py.py)."""
import numpy as np
import pytest

from conftest import golden_names, load_golden

PAIRS = golden_names("linematch_")


def view_of(g):
    from psl_slam_b200._lib import make_line_frame_view
    return make_line_frame_view(g["kl_cur"], g["desc_cur"], g["eq_cur"], g["lines3d_cur"], g["bounds"])


@pytest.mark.parametrize("name", PAIRS)
def test_descriptor_matchers(orc, name):
    g = load_golden(name)
    m, n = orc.line_match_nnr(g["desc_last"], g["desc_cur"], 0.95)
    assert np.array_equal(m, g["nnr12"]) and n == int(g["nnr_n"])
    a, n = orc.line_search_geom(g["kl_last"], g["desc_last"], g["has_ml"], g["kl_cur"], g["desc_cur"], g["bounds"], 0.95)
    assert np.array_equal(a, g["geom"]) and n == int(g["geom_n"])
    assert np.array_equal(orc.line_frame_bf_match(g["desc_last"], g["desc_cur"], 0.95, 50), g["bf"])
    m, n = orc.line_search_double(g["desc_last"], g["desc_cur"], 0.95, 50)
    assert np.array_equal(m, g["dbl"]) and n == int(g["dbl_n"])


@pytest.mark.parametrize("name", PAIRS)
def test_line_grid_and_projection(orc, name):
    g = load_golden(name)
    view, keep = view_of(g)
    q = g["queries0"]
    off = 0
    for i, ln in enumerate(g["area_len"]):
        got = orc.lines_in_area(view, q["x1"][i], q["y1"][i], q["x2"][i], q["y2"][i], 6.0, 0.96)
        assert np.array_equal(got, g["area_cat"][off:off + ln])
        off += ln
    a, n = orc.line_match_projection(view, q, g["desc_last"], g["claimed"], 0, 0.95)
    assert np.array_equal(a, g["proj0"]) and n == int(g["proj0_n"])
    a, n = orc.line_match_projection(view, g["queries1"], g["desc_last"], g["claimed"], 1, 0.8)
    assert np.array_equal(a, g["proj1"]) and n == int(g["proj1_n"])


def test_plane_assoc(orc):
    g = load_golden("plane_assoc")
    a, n = orc.plane_assoc(g["planes_cam"], g["pts"], g["Tcw"], g["map_planes"], g["map_bad"], 0.1, 0.86, 0)
    assert np.array_equal(a, g["assign0"]) and n == int(g["n0"])
    a, n = orc.plane_assoc(g["planes_cam"], g["pts"], g["Tcw"], g["map_planes"], None, 0.1, 0.86, 1)
    assert np.array_equal(a, g["assign1"]) and n == int(g["n1"])


def test_degenerate_inputs(orc):
    d = np.zeros((0, 32), np.uint8)
    one = np.full((1, 32), 7, np.uint8)
    three = np.arange(96, dtype=np.uint8).reshape(3, 32)
    assert orc.line_match_nnr(d, three, 0.95)[1] == 0
    m, n = orc.line_match_nnr(three, one, 0.95)  # < 2 train rows: the reference reads out of bounds; pinned: no match
    assert n == 0 and np.all(m == -1)
    assert np.all(orc.line_frame_bf_match(three, one, 0.95, 50) == -1)
    assert orc.line_search_double(d, three, 0.95, 50)[1] == 0
    a, n = orc.plane_assoc(np.zeros((0, 4)), np.zeros((0, 15)), np.eye(4), np.zeros((3, 4)), None, 0.1, 0.86, 0)
    assert n == 0 and len(a) == 0


@pytest.mark.parametrize("name", PAIRS)
def test_line_search_triangulation(orc, name):
    """SearchForTriangulation = the SearchDouble result (pinned above against cv2) minus pairs that already own a MapLine."""
    g = load_golden(name)
    rng = np.random.default_rng(5)
    ml1 = (rng.random(len(g["desc_last"])) < 0.3).astype(np.uint8)
    ml2 = (rng.random(len(g["desc_cur"])) < 0.3).astype(np.uint8)
    m, n = orc.line_search_triangulation(g["desc_last"], ml1, g["desc_cur"], ml2, 0.95, 50, True)
    want = g["dbl"].copy()
    for i, j in enumerate(want):
        if j >= 0 and (ml1[i] or ml2[j]):
            want[i] = -1
    assert np.array_equal(m, want) and n == int((want >= 0).sum())
    m2, n2 = orc.line_search_triangulation(g["desc_last"], ml1, g["desc_cur"], ml2, 0.95, 80, False)
    bf = orc.line_frame_bf_match(g["desc_last"], g["desc_cur"], 0.95, 80)
    want2 = np.where((bf >= 0) & (ml1 == 0) & (ml2[np.maximum(bf, 0)] == 0), bf, -1)
    assert np.array_equal(m2, want2) and n2 == int((want2 >= 0).sum())


@pytest.mark.parametrize("name", golden_names("linefuse_"))
def test_line_fuse(orc, name):
    """LSDmatcher::Fuse window search (KeyFrame::GetLinesInArea + level gate + min Hamming vs pKF->mDescriptors rows)."""
    g = load_golden(name)
    bi, bd, n = orc.line_fuse(g["kl"], g["kf_desc"], g["queries"], g["qdesc"], 0.998, 50)
    assert np.array_equal(bi, g["best_idx"]) and np.array_equal(bd, g["best_dist"]) and n == int((g["best_idx"] >= 0).sum())
    assert n > 5
    bi, bd, n = orc.line_fuse(g["kl"], g["kf_desc"], g["queries"], g["qdesc"], 0.9, 80)
    assert np.array_equal(bi, g["best_idx_loose"]) and np.array_equal(bd, g["best_dist_loose"])
    # edge cases: no lines, no queries, every query invalid
    q = g["queries"].copy()
    bi, bd, n = orc.line_fuse(g["kl"][:0], g["kf_desc"], q, g["qdesc"])
    assert n == 0 and (bi == -1).all() and (bd == 256).all()
    bi, bd, n = orc.line_fuse(g["kl"], g["kf_desc"], q[:0], g["qdesc"][:0])
    assert n == 0 and len(bi) == 0
    q["flags"] = 0
    bi, bd, n = orc.line_fuse(g["kl"], g["kf_desc"], q, g["qdesc"])
    assert n == 0 and (bi == -1).all()


def _same_planes(got, g):
    le, pl, nr, ow = got[:4]
    eq = lambda a, b: a.shape == b.shape and np.array_equal(a, b, equal_nan=True)  # noqa: E731  a degenerate pair yields NaNs
    assert eq(le, g["le_l"]) and eq(pl, g["planes"]) and eq(nr, g["normals"])
    assert np.array_equal(ow, g["junction_of"])


@pytest.mark.parametrize("name", golden_names("planes_"))
def test_plane_hypotheses(orc, name):
    """Frame::ExtractLSD plane hypotheses + OldPlane against the independent numpy restatement."""
    g = load_golden(name)
    got = orc.plane_hypotheses(g["kl"], g["line_eq"], g["lines3d"], g["junctions"])
    _same_planes(got, g)
    assert got[4] == len(g["planes"]) >= 1
    # capacity smaller than the result: the count is still reported, the first `cap` are stored
    le, pl, nr, ow, n = orc.plane_hypotheses(g["kl"], g["line_eq"], g["lines3d"], g["junctions"], cap=1)
    assert n == len(g["planes"]) and np.array_equal(pl, g["planes"][:1], equal_nan=True)
    le, pl, nr, ow, n = orc.plane_hypotheses(g["kl"], g["line_eq"], g["lines3d"], g["junctions"][:0])
    assert n == 0 and len(le) == 0


@pytest.mark.parametrize("name", golden_names("lines3d_"))
def test_lines_3d(orc, name):
    """Frame::isLineGood: the oracle (own Jacobi SVD) against the restatement over the real cv2.SVDecomp."""
    g = load_golden(name)
    cam = [float(v) for v in g["cam"]]
    l3, eq = orc.lines_3d(g["kl"], g["depth"], *cam, int(g["seed"]))
    assert np.array_equal(l3, g["lines3d"]) and np.array_equal(eq, g["line_eq"])
    assert (np.abs(l3).sum(1) > 0).sum() >= 20
    # another seed draws other pairs: end points may change, the set of lines with a fit hardly does
    l3b, _ = orc.lines_3d(g["kl"], g["depth"], *cam, int(g["seed"]) + 1)
    assert abs(int((np.abs(l3b).sum(1) > 0).sum()) - int((np.abs(l3).sum(1) > 0).sum())) <= 6
    # no depth at all / no lines
    l3z, eqz = orc.lines_3d(g["kl"], np.zeros_like(g["depth"]), *cam, 1)
    assert not l3z.any() and (eqz == -1).all()
    assert len(orc.lines_3d(g["kl"][:0], g["depth"], *cam, 1)[0]) == 0


@pytest.mark.parametrize("name", golden_names("junctions_"))
def test_line_junctions(orc, name):
    """CPartiallyRecoverConnectivity (PartiallyRecoverConnectivity.cpp:14-133): the fans are bit-exact against the
    restatement over the real cv2 primitives; the 3-D cross points of convertFansToKeyLines (Frame.cc:380-472) agree
    with the independent numpy.linalg solution to 1e-9 (the reference solves the 2x2 system with Eigen's QR)."""
    g = load_golden(name)
    w, h = (int(v) for v in g["size"])
    fans, js = orc.line_junctions(g["kl"], g["lines3d"], w, h, float(g["radius"]), float(g["fan_thr"]))
    assert np.array_equal(fans, g["fans"]) and len(fans) > 20
    want = g["junctions"]
    assert len(js) == len(want)
    for f in ("l1", "l2", "cross2d_x", "cross2d_y"):
        assert np.array_equal(js[f], want[f]), f
    assert np.allclose(js["cross3d"], want["cross3d"], rtol=1e-9, atol=1e-12)
    # fans only (no 3-D lines), no lines, a tiny output capacity is reported through the counts
    fans2, js2 = orc.line_junctions(g["kl"], None, w, h, float(g["radius"]), float(g["fan_thr"]))
    assert np.array_equal(fans2, fans) and len(js2) == 0
    fans3, js3 = orc.line_junctions(g["kl"][:0], g["lines3d"][:0], w, h)
    assert len(fans3) == 0 and len(js3) == 0


@pytest.mark.parametrize("name", golden_names("linetriangnew_"))
def test_line_search_triangulation_new(orc, name):
    """LSDmatcher::SearchForTriangulationNew (LSDmatcher.cpp:518-658, 783-824) against the independent Python restatement
    over cv2.BFMatcher."""
    g = load_golden(name)
    args = (g["kl1"], g["desc1"], g["func1"], g["ml1"], g["kl2"], g["desc2"], g["func2"], g["ml2"], g["F21"], g["F12"])
    m, n = orc.line_search_triangulation_new(*args, float(g["nn_ratio"]), 50, int(g["is_double"]))
    assert np.array_equal(m, g["pairs"]) and n == int(g["npairs"]) and n > 20
    m, n = orc.line_search_triangulation_new(*args, 0.99, 90, 1 - int(g["is_double"]))
    assert np.array_equal(m, g["pairs_b"]) and n == int(g["npairs_b"])
    # one line on the other side: knnMatch has no second neighbour, nothing matches; no lines at all
    one = (g["kl1"], g["desc1"], g["func1"], g["ml1"], g["kl2"][:1], g["desc2"][:1], g["func2"][:1], g["ml2"][:1], g["F21"], g["F12"])
    m, n = orc.line_search_triangulation_new(*one, 0.95, 50, 0)
    assert n == 0 and (m == -1).all()
