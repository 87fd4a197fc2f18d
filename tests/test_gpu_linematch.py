"""GPU: the CUDA line matchers through the C-ABI against the oracle and the committed goldens (bit-exact
indices and counts)."""
import numpy as np
import pytest

from conftest import golden_names, load_golden

pytestmark = pytest.mark.gpu
PAIRS = golden_names("linematch_")


def _frames(g):
    from psl_slam_b200 import LineFrameData
    b = tuple(g["bounds"])
    cur = LineFrameData(g["kl_cur"], g["desc_cur"], g["eq_cur"], g["lines3d_cur"], b)
    last = LineFrameData(g["kl_last"], g["desc_last"], None, None, b)
    return cur, last


@pytest.mark.parametrize("name", PAIRS)
def test_descriptor_matchers_vs_golden(name):
    from psl_slam_b200 import LSDmatcher
    g = load_golden(name)
    cur, last = _frames(g)
    m = LSDmatcher(0.95, True)
    m12, n = m.match(g["desc_last"], g["desc_cur"], 0.95)
    assert np.array_equal(m12, g["nnr12"]) and n == int(g["nnr_n"])
    a, n = m.SearchByGeomNApearance(cur, last, g["has_ml"], 0.95)
    assert np.array_equal(a, g["geom"]) and n == int(g["geom_n"])
    assert np.array_equal(m.FrameBFMatch(g["desc_last"], g["desc_cur"], 50), g["bf"])
    d, n = m.SearchDouble(g["desc_last"], g["desc_cur"])
    assert np.array_equal(d, g["dbl"]) and n == int(g["dbl_n"])


@pytest.mark.parametrize("name", PAIRS)
def test_projection_vs_golden(name):
    from psl_slam_b200 import LSDmatcher
    g = load_golden(name)
    cur, _ = _frames(g)
    a, n = LSDmatcher(0.95).SearchByProjectionLastFrame(cur, g["queries0"], g["desc_last"], g["claimed"])
    assert np.array_equal(a, g["proj0"]) and n == int(g["proj0_n"])
    a, n = LSDmatcher(0.8).SearchByProjectionMapLines(cur, g["queries1"], g["desc_last"], g["claimed"])
    assert np.array_equal(a, g["proj1"]) and n == int(g["proj1_n"])


def test_random_sweep_vs_oracle(orc):
    """Random descriptors / geometry at several sizes, including the degenerate ones."""
    from psl_slam_b200 import LineFrameData, LSDmatcher
    from psl_slam_b200._lib import KEYLINE_DTYPE, LINE_QUERY_DTYPE, make_line_frame_view
    rng = np.random.default_rng(3)
    m = LSDmatcher(0.9)
    for n1, n2 in [(0, 5), (5, 0), (1, 1), (3, 1), (2, 2), (40, 57), (200, 200), (333, 150)]:
        base = rng.integers(0, 256, (max(n1, n2, 1), 32), dtype=np.uint8)
        d1 = base[:n1] ^ (rng.random((n1, 32)) < 0.08).astype(np.uint8) * rng.integers(0, 256, (n1, 32), dtype=np.uint8)
        d2 = base[rng.permutation(len(base))[:n2]] if n2 else np.zeros((0, 32), np.uint8)
        want, wn = orc.line_match_nnr(d1, d2, 0.9)
        got, gn = m.match(d1, d2, 0.9)
        assert np.array_equal(got, want) and gn == wn
        assert np.array_equal(m.FrameBFMatch(d1, d2, 50), orc.line_frame_bf_match(d1, d2, 0.9, 50))
        want, wn = orc.line_search_double(d1, d2, 0.9, 50)
        got, gn = m.SearchDouble(d1, d2)
        assert np.array_equal(got, want) and gn == wn
        if n1 == 0 or n2 == 0:
            continue
        # random segments in a 640x480 image, queries near them
        def rand_kl(n):
            kl = np.zeros(n, KEYLINE_DTYPE)
            kl["start_x"], kl["start_y"] = rng.uniform(0, 639, n), rng.uniform(0, 479, n)
            ang, ln = rng.uniform(0, np.pi, n), rng.uniform(50, 300, n)
            kl["end_x"] = np.clip(kl["start_x"] + ln * np.cos(ang), 0, 639)
            kl["end_y"] = np.clip(kl["start_y"] + ln * np.sin(ang), 0, 479)
            for a, b in (("s_oct_x", "start_x"), ("s_oct_y", "start_y"), ("e_oct_x", "end_x"), ("e_oct_y", "end_y")):
                kl[a] = kl[b]
            kl["line_length"] = np.hypot(kl["end_x"] - kl["start_x"], kl["end_y"] - kl["start_y"])
            return kl
        klc = rand_kl(n2)
        eq = np.zeros((n2, 3))
        for i, k in enumerate(klc):
            l = np.cross([k["start_x"], k["start_y"], 1.0], [k["end_x"], k["end_y"], 1.0])
            eq[i] = l / np.hypot(l[0], l[1])
        l3d = rng.normal(0, 1, (n2, 6))
        q = np.zeros(n1, LINE_QUERY_DTYPE)
        src = klc[rng.integers(0, n2, n1)]
        q["x1"], q["y1"] = src["start_x"] + rng.normal(0, 2, n1), src["start_y"] + rng.normal(0, 2, n1)
        q["x2"], q["y2"] = src["end_x"] + rng.normal(0, 2, n1), src["end_y"] + rng.normal(0, 2, n1)
        q["radius"] = rng.choice([3.0, 6.0, 15.0, 24.0], n1)
        q["sx"], q["sy"], q["ex"], q["ey"] = q["x1"], q["y1"], q["x2"], q["y2"]
        q["length"] = src["line_length"] * rng.uniform(0.6, 1.2, n1)
        q["normal"] = rng.normal(0, 1, (n1, 3))
        q["flags"] = rng.integers(0, 4, n1)
        claimed = (rng.random(n2) < 0.1).astype(np.uint8)
        bounds = (0.0, 0.0, 640.0, 480.0)
        view, keep = make_line_frame_view(klc, d2, eq, l3d, bounds)
        fd = LineFrameData(klc, d2, eq, l3d, bounds)
        for mode in (0, 1):
            want, wn = orc.line_match_projection(view, q, d1, claimed, mode, 0.9)
            got, gn = (m.SearchByProjectionLastFrame if mode == 0 else m.SearchByProjectionMapLines)(fd, q, d1, claimed)
            assert np.array_equal(got, want) and gn == wn, (n1, n2, mode)
        kll = rand_kl(n1)
        has = (rng.random(n1) < 0.8).astype(np.uint8)
        want, wn = orc.line_search_geom(kll, d1, has, klc, d2, bounds, 0.95)
        got, gn = m.SearchByGeomNApearance(fd, LineFrameData(kll, d1, None, None, bounds), has, 0.95)
        assert np.array_equal(got, want) and gn == wn


def test_plane_assoc_vs_golden_and_oracle(orc):
    from psl_slam_b200 import InsectLineMatch
    g = load_golden("plane_assoc")
    m = InsectLineMatch(0.1, 0.86)
    a, n = m.SearchMapInsectline(g["planes_cam"], g["pts"], g["Tcw"], g["map_planes"], g["map_bad"])
    assert np.array_equal(a, g["assign0"]) and n == int(g["n0"])
    a, n = m.AssociatePlanesByBoundary(g["planes_cam"], g["pts"], g["Tcw"], g["map_planes"])
    assert np.array_equal(a, g["assign1"]) and n == int(g["n1"])
    rng = np.random.default_rng(8)
    for n_ljl, n_map in [(0, 4), (3, 0), (25, 300)]:
        pc = rng.normal(0, 1, (n_ljl, 4)).astype(np.float32)
        pts = rng.normal(0, 1, (n_ljl, 15))
        mp = rng.normal(0, 1, (n_map, 4)).astype(np.float32)
        mp[:, :3] /= np.linalg.norm(mp[:, :3], axis=1, keepdims=True) + 1e-9
        pc[:, :3] /= np.linalg.norm(pc[:, :3], axis=1, keepdims=True) + 1e-9
        for mode in (0, 1):
            want, wn = orc.plane_assoc(pc, pts, np.eye(4), mp, None, 1.5, 0.5, mode)
            got, gn = m.__class__(1.5, 0.5, ctx=m.ctx)._run(pc, pts, np.eye(4), mp, None, mode)
            assert np.array_equal(got, want) and gn == wn


@pytest.mark.parametrize("name", PAIRS)
def test_line_search_triangulation_vs_oracle(orc, name):
    from psl_slam_b200 import LSDmatcher
    g = load_golden(name)
    rng = np.random.default_rng(5)
    ml1 = (rng.random(len(g["desc_last"])) < 0.3).astype(np.uint8)
    ml2 = (rng.random(len(g["desc_cur"])) < 0.3).astype(np.uint8)
    m = LSDmatcher(0.95)
    for as_pairs, dbl, th in ((True, True, 50), (False, False, 80), (False, True, 80)):
        want, wn = orc.line_search_triangulation(g["desc_last"], ml1, g["desc_cur"], ml2, 0.95, th, as_pairs or dbl)
        got, gn = m.SearchForTriangulation(g["desc_last"], ml1, g["desc_cur"], ml2, as_pairs, dbl)
        assert np.array_equal(got, want) and gn == wn


@pytest.mark.parametrize("name", golden_names("linefuse_"))
def test_line_fuse_vs_golden_and_oracle(orc, name):
    from psl_slam_b200 import LSDmatcher, PslError
    g = load_golden(name)
    m = LSDmatcher(0.75)
    bi, bd = m.Fuse(g["kl"], g["kf_desc"], g["queries"], g["qdesc"])
    assert np.array_equal(bi, g["best_idx"]) and np.array_equal(bd, g["best_dist"])
    bi, bd = m.Fuse(g["kl"], g["kf_desc"], g["queries"], g["qdesc"], th_cos=0.9)
    want_i, want_d, _ = orc.line_fuse(g["kl"], g["kf_desc"], g["queries"], g["qdesc"], 0.9, 50)
    assert np.array_equal(bi, want_i) and np.array_equal(bd, want_d)
    # a larger synthetic case against the oracle: 1500 lines, 3000 queries, mixed levels, degenerate lines
    rng = np.random.default_rng(11)
    reps = 1500 // len(g["kl"]) + 1
    kl = np.tile(g["kl"], reps)[:1500].copy()
    for f in ("pt_x", "start_x", "end_x"):
        kl[f] += np.repeat(rng.uniform(-40, 40, reps), len(g["kl"]))[:1500].astype(np.float32)
    kl["end_x"][7], kl["end_y"][7] = kl["start_x"][7], kl["start_y"][7]
    kd = rng.integers(0, 256, (1500, 32), dtype=np.uint8)
    q = np.tile(g["queries"], 3000 // len(g["queries"]) + 1)[:3000].copy()
    q["radius"] = rng.uniform(5, 60, 3000).astype(np.float32)
    q["pred_level"] = rng.integers(-1, 4, 3000)
    qd = kd[rng.integers(0, 1500, 3000)] ^ (rng.integers(0, 256, (3000, 32), dtype=np.uint8) & rng.integers(0, 256, (3000, 32), dtype=np.uint8) & rng.integers(0, 256, (3000, 32), dtype=np.uint8))
    near, _, _ = orc.line_fuse(kl, kd, q, qd, 0.95, 256)   # any candidate at all: make half of those queries fusable
    hit = np.flatnonzero(near >= 0)[::2]
    qd[hit] = kd[near[hit]] ^ (qd[hit] & rng.integers(0, 256, (len(hit), 32), dtype=np.uint8) & 0x11)
    bi, bd = m.Fuse(kl, kd, q, qd, th_cos=0.95)
    want_i, want_d, wn = orc.line_fuse(kl, kd, q, qd, 0.95, 50)
    assert np.array_equal(bi, want_i) and np.array_equal(bd, want_d) and wn > 200
    # empty inputs and the descriptor-row contract
    bi, bd = m.Fuse(g["kl"][:0], g["kf_desc"], g["queries"], g["qdesc"])
    assert (bi == -1).all() and (bd == 256).all()
    bi, bd = m.Fuse(g["kl"], g["kf_desc"], g["queries"][:0], g["qdesc"][:0])
    assert len(bi) == 0
    with pytest.raises(PslError):
        m.Fuse(g["kl"], g["kf_desc"][: len(g["kl"]) - 1], g["queries"], g["qdesc"])


@pytest.mark.parametrize("name", golden_names("planes_"))
def test_plane_hypotheses_vs_golden_and_oracle(orc, name):
    from psl_slam_b200 import Context, PslError, default_config, plane_hypotheses
    g = load_golden(name)
    ctx = Context(default_config())
    le, pl, nr, ow = plane_hypotheses(ctx, g["kl"], g["line_eq"], g["lines3d"], g["junctions"])
    same = lambda a, b: a.shape == b.shape and np.array_equal(a, b, equal_nan=True)  # noqa: E731  a degenerate pair yields NaNs
    # (their payload bits are the only thing that may differ between the CPU and the GPU)
    assert same(le, g["le_l"]) and same(pl, g["planes"]) and same(nr, g["normals"]) and np.array_equal(ow, g["junction_of"])
    # many junctions (several 32-wide rounds, > 32 kept planes): tiled copies with every copy on its own plane offset
    reps = 12
    kl = np.tile(g["kl"], reps)
    eq = np.tile(g["line_eq"], (reps, 1))
    l3 = np.tile(g["lines3d"], (reps, 1)).copy()
    js = np.tile(g["junctions"], reps).copy()
    for r in range(reps):
        sl = slice(r * len(g["kl"]), (r + 1) * len(g["kl"]))
        l3[sl, 2::3] += 0.37 * r
        jsl = slice(r * len(g["junctions"]), (r + 1) * len(g["junctions"]))
        js["l1"][jsl] += r * len(g["kl"])
        js["l2"][jsl] += r * len(g["kl"])
        js["cross3d"][jsl, 2] += 0.37 * r
    want = orc.plane_hypotheses(kl, eq, l3, js)
    le, pl, nr, ow = plane_hypotheses(ctx, kl, eq, l3, js)
    assert same(le, want[0]) and same(pl, want[1]) and same(nr, want[2]) and np.array_equal(ow, want[3])
    if name != "planes_case2":
        assert len(pl) > 32
        with pytest.raises(PslError) as e:
            plane_hypotheses(ctx, kl, eq, l3, js, cap=3)
        assert e.value.code == -3
    le, pl, nr, ow = plane_hypotheses(ctx, g["kl"], g["line_eq"], g["lines3d"], g["junctions"][:0])
    assert len(pl) == 0 and len(le) == 0


@pytest.mark.parametrize("name", golden_names("lines3d_"))
def test_lines_3d_vs_golden_and_oracle(orc, name):
    """Frame::isLineGood on CUDA against the cv2-SVD golden and the oracle, host and batched device forms."""
    import ctypes as C

    import torch

    from psl_slam_b200 import KEYLINE_DTYPE, Context, _lib, default_config, lines_3d
    g = load_golden(name)
    cam = [float(v) for v in g["cam"]]
    ctx = Context(default_config())
    l3, eq = lines_3d(ctx, g["kl"], g["depth"], *cam, int(g["seed"]))
    assert np.array_equal(l3, g["lines3d"]) and np.array_equal(eq, g["line_eq"])
    for seed in (1, 2, 12345):
        want = orc.lines_3d(g["kl"], g["depth"], *cam, seed)
        got = lines_3d(ctx, g["kl"], g["depth"], *cam, seed)
        assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1])
    l3z, eqz = lines_3d(ctx, g["kl"], np.zeros_like(g["depth"]), *cam, 1)
    assert not l3z.any() and (eqz == -1).all()
    # batched device form: two frames (the second with half the lines and a shifted depth), strided depth rows
    kl = np.ascontiguousarray(g["kl"], KEYLINE_DTYPE)
    n, cap = len(kl), len(kl) + 3
    h, w = g["depth"].shape
    stride = w + 5
    dep2 = np.zeros((2, h, stride), np.float32)
    dep2[0, :, :w] = g["depth"]
    dep2[1, :, :w] = g["depth"] * np.float32(1.1)
    klb = np.zeros((2, cap), KEYLINE_DTYPE)
    klb[0, :n] = kl
    klb[1, : n // 2] = kl[: n // 2]
    ns = [n, n // 2]
    d_kl = torch.from_numpy(klb.view(np.uint8).reshape(2, cap, 68)).cuda()
    d_n = torch.tensor(ns, dtype=torch.int32, device="cuda")
    d_dep = torch.from_numpy(dep2).cuda()
    d_l3 = torch.full((2, cap, 6), 7.0, dtype=torch.float64, device="cuda")
    d_eq = torch.full((2, cap, 3), 7.0, dtype=torch.float32, device="cuda")
    ctx.check(_lib.lib().psl_lines_3d_dev(ctx.handle, d_kl.data_ptr(), d_n.data_ptr(), cap, 2, d_dep.data_ptr(), w, h, stride,
                                          h * stride, C.c_float(cam[0]), C.c_float(cam[1]), C.c_float(cam[2]),
                                          C.c_float(cam[3]), C.c_uint32(5), d_l3.data_ptr(), d_eq.data_ptr()))
    ctx.sync()
    for b in range(2):
        want = orc.lines_3d(klb[b, : ns[b]], np.ascontiguousarray(dep2[b, :, :w]), *cam, 5)
        assert np.array_equal(d_l3[b, : ns[b]].cpu().numpy(), want[0]) and np.array_equal(d_eq[b, : ns[b]].cpu().numpy(), want[1])
        assert (d_l3[b, ns[b]:].cpu().numpy() == 7.0).all()


@pytest.mark.parametrize("name", golden_names("junctions_"))
def test_line_junctions_vs_golden_and_oracle(orc, name):
    """The junction detection of Frame::ExtractLSD (PartiallyRecoverConnectivity.cpp:14-133, Frame.cc:380-472): fans
    bit-exact against the cv2-primitive golden; the junction records identical to the oracle's (same fp64 statements)
    and within 1e-9 of the independent numpy.linalg cross points of the golden."""
    import torch
    from psl_slam_b200 import Context, PslError, default_config, line_junctions
    from psl_slam_b200._lib import JUNCTION_DTYPE, KEYLINE_DTYPE, lib
    g = load_golden(name)
    ctx = Context(default_config())
    w, h = (int(v) for v in g["size"])
    r, t = float(g["radius"]), float(g["fan_thr"])
    fans, js = line_junctions(ctx, g["kl"], g["lines3d"], w, h, r, t)
    assert np.array_equal(fans, g["fans"])
    ofans, ojs = orc.line_junctions(g["kl"], g["lines3d"], w, h, r, t)
    assert np.array_equal(fans, ofans) and len(js) == len(ojs)
    for f in ("l1", "l2", "cross2d_x", "cross2d_y", "cross3d"):
        assert np.array_equal(js[f], ojs[f]), f
    want = g["junctions"]
    assert len(js) == len(want) and np.array_equal(js["l1"], want["l1"]) and np.array_equal(js["l2"], want["l2"])
    assert np.allclose(js["cross3d"], want["cross3d"], rtol=1e-9, atol=1e-12)
    # fans only; nothing; capacity
    f2, j2 = line_junctions(ctx, g["kl"], None, w, h, r, t)
    assert np.array_equal(f2, fans) and len(j2) == 0
    f3, j3 = line_junctions(ctx, g["kl"][:0], g["lines3d"][:0], w, h, r, t)
    assert len(f3) == 0 and len(j3) == 0
    with pytest.raises(PslError):
        line_junctions(ctx, g["kl"], g["lines3d"], w, h, r, t, cap=5)
    # the chain isLineGood -> junctions -> plane hypotheses runs on the junction records as they are
    from psl_slam_b200 import plane_hypotheses
    eq = np.where(np.abs(g["lines3d"]).sum(1, keepdims=True) > 0,
                  (g["lines3d"][:, 3:] - g["lines3d"][:, :3]) / np.maximum(
                      np.linalg.norm(g["lines3d"][:, 3:] - g["lines3d"][:, :3], axis=1, keepdims=True), 1e-12),
                  -1.0).astype(np.float32)
    le, pl, nr, ow = plane_hypotheses(ctx, g["kl"], eq, g["lines3d"], js)
    wle, wpl, wnr, wow, _ = orc.plane_hypotheses(g["kl"], eq, g["lines3d"], js)
    assert np.array_equal(pl, wpl, equal_nan=True) and np.array_equal(ow, wow)
    # batched device form: three frames (the golden's lines, none, the first half) in one launch
    n, cap, lc = len(g["kl"]), 512, len(g["kl"]) + 3
    kl = np.zeros((3, lc), KEYLINE_DTYPE)
    l3 = np.zeros((3, lc, 6))
    kl[0, :n], l3[0, :n] = g["kl"], g["lines3d"]
    kl[2, : n // 2], l3[2, : n // 2] = g["kl"][: n // 2], g["lines3d"][: n // 2]
    d_kl = torch.from_numpy(kl.view(np.uint8).reshape(-1)).cuda()
    d_l3 = torch.from_numpy(l3).cuda()
    d_n = torch.tensor([n, 0, n // 2], dtype=torch.int32, device="cuda")
    d_f = torch.zeros(3 * cap * 4, dtype=torch.float32, device="cuda")
    d_j = torch.zeros(3 * cap * JUNCTION_DTYPE.itemsize, dtype=torch.uint8, device="cuda")
    d_c = torch.zeros(6, dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    ctx.check(lib().psl_line_junctions_dev(ctx.handle, d_kl.data_ptr(), d_n.data_ptr(), lc, 3, d_l3.data_ptr(), w, h,
                                           r, t, d_f.data_ptr(), d_j.data_ptr(), cap, d_c.data_ptr(),
                                           d_c.data_ptr() + 12))
    ctx.sync()
    c = d_c.cpu().numpy()
    bf = d_f.cpu().numpy().reshape(3, cap, 4)
    bj = d_j.cpu().numpy().view(JUNCTION_DTYPE).reshape(3, cap)
    assert c[0] == len(fans) and c[3] == len(js) and c[1] == 0 and c[4] == 0
    assert np.array_equal(bf[0, : c[0]], fans) and np.array_equal(bj[0, : c[3]]["cross3d"], js["cross3d"])
    hf, hj = orc.line_junctions(g["kl"][: n // 2], g["lines3d"][: n // 2], w, h, r, t)
    assert c[2] == len(hf) and c[5] == len(hj) and np.array_equal(bf[2, : c[2]], hf)
    assert np.array_equal(bj[2, : c[5]]["l1"], hj["l1"]) and np.array_equal(bj[2, : c[5]]["cross3d"], hj["cross3d"])


def test_line_junctions_random_large(orc):
    """1100 random segments (more than 48 KB of per-frame line state in shared memory), with zero-length, vertical,
    horizontal and duplicated segments, and 3-D lines that are parallel / identical / missing: against the oracle."""
    from psl_slam_b200 import Context, default_config, line_junctions
    from psl_slam_b200._lib import KEYLINE_DTYPE
    rng = np.random.default_rng(77)
    n = 1100
    kl = np.zeros(n, KEYLINE_DTYPE)
    c = rng.uniform([20, 20], [1900, 1060], (n, 2))
    d = rng.normal(0, 60, (n, 2))
    kl["start_x"], kl["start_y"] = (c - d)[:, 0], (c - d)[:, 1]
    kl["end_x"], kl["end_y"] = (c + d)[:, 0], (c + d)[:, 1]
    kl["end_x"][:20] = kl["start_x"][:20]                       # vertical
    kl["end_y"][20:40] = kl["start_y"][20:40]                   # horizontal
    kl["end_x"][40:50], kl["end_y"][40:50] = kl["start_x"][40:50], kl["start_y"][40:50]   # zero length
    kl[50:60] = kl[60:70]                                       # duplicates
    for f in ("start_x", "start_y", "end_x", "end_y"):
        kl[f][100:200] = np.round(kl[f][100:200])               # integer coordinates: exact ties on the rectangle tests
    l3 = rng.normal(0, 1, (n, 6)) + np.array([0, 0, 3, 0, 0, 3])
    l3[:100] = 0                                                # no 3-D line
    l3[100:150, 3:] = l3[100:150, :3] + (l3[150:200, 3:] - l3[150:200, :3])   # parallel to another line
    l3[200:210] = l3[210:220]                                   # identical
    ctx = Context(default_config())
    fans, js = line_junctions(ctx, kl, l3, 1920, 1080, 20.0, float(np.float32(np.pi / 4)), cap=20000)
    wf, wj = orc.line_junctions(kl, l3, 1920, 1080, 20.0, float(np.float32(np.pi / 4)), cap=20000)
    assert len(fans) == len(wf) and len(fans) > 500
    assert np.array_equal(fans, wf, equal_nan=True)
    assert len(js) == len(wj) and np.array_equal(js["l1"], wj["l1"]) and np.array_equal(js["l2"], wj["l2"])
    assert np.array_equal(js["cross3d"], wj["cross3d"])


@pytest.mark.parametrize("name", golden_names("linetriangnew_"))
def test_line_search_triangulation_new(orc, name):
    """LSDmatcher::SearchForTriangulationNew (LSDmatcher.cpp:518-658, 783-824)."""
    from psl_slam_b200 import LSDmatcher
    g = load_golden(name)
    a = (g["kl1"], g["desc1"], g["func1"], g["ml1"])
    b = (g["kl2"], g["desc2"], g["func2"], g["ml2"])
    m, n = LSDmatcher(float(g["nn_ratio"])).SearchForTriangulationNew(a, b, g["F21"], g["F12"], bool(g["is_double"]))
    assert np.array_equal(m, g["pairs"]) and n == int(g["npairs"])
    mm = LSDmatcher(0.99)
    mm.TH_LOW = 90
    m, n = mm.SearchForTriangulationNew(a, b, g["F21"], g["F12"], not bool(g["is_double"]))
    assert np.array_equal(m, g["pairs_b"]) and n == int(g["npairs_b"])
    # perturbed geometry (other fundamental matrices, jittered end points): against the oracle
    rng = np.random.default_rng(9)
    k1 = g["kl1"].copy()
    for f in ("start_x", "start_y", "end_x", "end_y"):
        k1[f] += rng.normal(0, 4, len(k1)).astype(np.float32)
    F21 = (g["F21"] * rng.uniform(0.8, 1.2, (3, 3))).astype(np.float32)
    F12 = (g["F12"] * rng.uniform(0.8, 1.2, (3, 3))).astype(np.float32)
    m, n = LSDmatcher(0.95).SearchForTriangulationNew((k1, g["desc1"], g["func1"], g["ml1"]), b, F21, F12, True)
    wm, wn = orc.line_search_triangulation_new(k1, g["desc1"], g["func1"], g["ml1"], g["kl2"], g["desc2"], g["func2"], g["ml2"],
                                               F21, F12, 0.95, 50, 1)
    assert np.array_equal(m, wm) and n == wn
    # a single line on the other side / none
    one = (g["kl2"][:1], g["desc2"][:1], g["func2"][:1], g["ml2"][:1])
    m, n = LSDmatcher(0.95).SearchForTriangulationNew(a, one, g["F21"], g["F12"], False)
    assert n == 0 and (m == -1).all()
    none = (g["kl2"][:0], g["desc2"][:0], g["func2"][:0], g["ml2"][:0])
    m, n = LSDmatcher(0.95).SearchForTriangulationNew(a, none, g["F21"], g["F12"], False)
    assert n == 0 and (m == -1).all()
