"""GPU: the batched point front end (extract + RGB-D stereo + SearchByProjection against the previous
frame, cfg 2 / cfg 4) against the oracle chain, across chunk boundaries."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def oracle_chain(orc, gray, depth, T, K, th=15.0, nn_ratio=0.9, ori=True):
    scale, _, _, _ = orc.orb_tables(orc.params())
    cam = np.array([K["fx"], K["fy"], K["cx"], K["cy"], K["bf"]], np.float32)
    f = np.float32(1.0) / np.float32(K["depth_factor"])
    bounds = (0.0, 0.0, float(gray.shape[2]), float(gray.shape[1]))
    res, prev = [], None
    for b in range(len(gray)):
        kps, desc = orc.orb_extract(gray[b])
        ur, z = orc.stereo_from_rgbd(kps, depth[b], f, K["bf"])
        assign, nm = np.full(len(kps), -1, np.int32), 0
        if prev is not None:
            q = orc.queries_from_last_frame(prev[0], prev[3], T[b - 1], T[b], cam, scale, th, bounds)
            assign, nm = orc.match_projection(kps, ur, desc, bounds, q, prev[1], None, 0, 100, nn_ratio, ori)
        res.append((kps, desc, ur, z, assign, nm))
        prev = (kps, desc, ur, z)
    return res


def test_track_batch_vs_oracle_chain(orc):
    from psl_slam_b200 import ORBextractor, make_camera, make_track_params, synth, track_orb_batch
    K = synth.ICL
    gray, depth, T = synth.sequence(7, 5)
    T = T.astype(np.float32)
    T[1:, :3, 3] += np.random.default_rng(1).normal(0, 0.002, (4, 3)).astype(np.float32)  # imperfect pose prior
    ex = ORBextractor(chunk_frames=2)
    cam = make_camera(K["fx"], K["fy"], K["cx"], K["cy"], K["bf"], K["depth_factor"])
    out = track_orb_batch(ex, gray, depth, T, cam, make_track_params(15.0, 0.9, True))
    ref = oracle_chain(orc, gray, depth, T, K)
    assert out["nmatches"][0] == 0
    for b, (kps, desc, ur, z, assign, nm) in enumerate(ref):
        n = out["n"][b]
        assert n == len(kps)
        for fld in kps.dtype.names:
            assert np.array_equal(out["kps"][b, :n][fld], kps[fld]), (b, fld)
        assert np.array_equal(out["desc"][b, :n], desc)
        assert np.array_equal(out["u_right"][b, :n], ur) and np.array_equal(out["z"][b, :n], z)
        assert np.array_equal(out["assign"][b, :n], assign), b
        assert out["nmatches"][b] == nm
        if b:
            assert nm > 300  # the synthetic sequence really tracks


def test_track_wide_window_and_no_orientation(orc):
    from psl_slam_b200 import ORBextractor, make_camera, make_track_params, synth, track_orb_batch
    K = synth.ICL
    gray, depth, T = synth.sequence(8, 3)
    T = T.astype(np.float32)
    ex = ORBextractor()
    cam = make_camera(K["fx"], K["fy"], K["cx"], K["cy"], K["bf"], K["depth_factor"])
    out = track_orb_batch(ex, gray, depth, T, cam, make_track_params(30.0, 0.9, False))
    ref = oracle_chain(orc, gray, depth, T, K, th=30.0, ori=False)
    for b, (kps, desc, ur, z, assign, nm) in enumerate(ref):
        n = out["n"][b]
        assert np.array_equal(out["assign"][b, :n], assign) and out["nmatches"][b] == nm


def test_combined_frontend_vs_oracle_chain(orc):
    """cfg 4 shaped: ORB + lines + both frame-to-frame matchers for a batch spanning several chunks."""
    from psl_slam_b200 import Context, ORBextractor, default_config, make_camera, make_track_params, synth, \
        track_frontend_batch
    K = synth.ICL
    gray, depth, T = synth.sequence(9, 4)
    T = T.astype(np.float32)
    cfg = default_config()
    cfg.chunk_frames, cfg.line_chunk_frames = 3, 2
    ex = ORBextractor(ctx=Context(cfg))
    cam = make_camera(K["fx"], K["fy"], K["cx"], K["cy"], K["bf"], K["depth_factor"])
    out = track_frontend_batch(ex, gray, depth, T, cam, make_track_params(15.0, 0.9, True), 0.95)
    ref = oracle_chain(orc, gray, depth, T, K)
    bounds = np.array([0, 0, gray.shape[2], gray.shape[1]], np.float32)
    prev = None
    for b, (kps, desc, ur, z, assign, nm) in enumerate(ref):
        n = out["n"][b]
        assert n == len(kps) and np.array_equal(out["desc"][b, :n], desc)
        assert np.array_equal(out["assign"][b, :n], assign) and out["nmatches"][b] == nm
        kl, ld, eq, _ = orc.line_extract(gray[b], 200)
        nl = out["nl"][b]
        assert nl == len(kl) and nl > 30
        assert out["kl"][b, :nl].tobytes() == kl.tobytes() and np.array_equal(out["ldesc"][b, :nl], ld)
        assert np.array_equal(out["lineeq"][b, :nl], eq)
        if prev is None:
            assert out["line_nmatches"][b] == 0 and np.all(out["line_assign"][b, :nl] == -1)
        else:
            la, ln = orc.line_search_geom(prev[0], prev[1], np.ones(len(prev[0]), np.uint8), kl, ld, bounds, 0.95)
            assert np.array_equal(out["line_assign"][b, :nl], la) and out["line_nmatches"][b] == ln and ln > 10
        prev = (kl, ld)
    # the CPU-baseline harness computes the same counts
    cam6 = np.array([K["fx"], K["fy"], K["cx"], K["cy"], K["bf"], np.float32(1.0) / np.float32(K["depth_factor"])], np.float32)
    n2, nm2, nl2, lnm2 = orc.frontend_batch_mt(gray, depth, T[:, :3, :4].reshape(len(T), 12), cam6, nthreads=2)
    assert np.array_equal(n2, out["n"]) and np.array_equal(nm2, out["nmatches"])
    assert np.array_equal(nl2, out["nl"]) and np.array_equal(lnm2, out["line_nmatches"])


def test_input_conversion_vs_cv2_golden():
    """Tracking::GrabImageRGBD's cvtColor / convertTo (K0) against real cv2 4.13 outputs."""
    from conftest import load_golden
    from psl_slam_b200 import Context, convert_rgbd, default_config, synth
    g = load_golden("convert_rgbd")
    ctx = Context(default_config())
    for key, arr, order in (("gray_rgb", g["rgb"], True), ("gray_bgr", g["rgb"], False), ("gray_rgba", g["rgba"], True),
                            ("gray_bgra", g["rgba"], False)):
        gray, dep = convert_rgbd(ctx, np.stack([arr, arr[::-1].copy()]), order, np.stack([g["depth"], g["depth"]]))
        assert np.array_equal(gray[0], g[key]) and np.array_equal(gray[1], g[key][::-1])
        assert np.array_equal(dep[0], g["depth_f"])
    # the synthetic generator's gray frames are this conversion of its RGB renders
    poster = synth.make_poster(3, 512, 200)
    rgb, _ = synth.render(poster, synth.trajectory(1, 3)[0], 320, 240)
    gray, _ = convert_rgbd(ctx, rgb[None], True, None)
    assert np.array_equal(gray[0], synth.rgb_to_gray(rgb))
    # both kernels (16 pixels per thread for rows of whole 16-pixel groups, 4 per thread otherwise), both channel orders
    rng = np.random.default_rng(11)
    for w in (64, 640, 52, 37):
        img = rng.integers(0, 256, (3, 24, w, 3), dtype=np.uint8)
        c = img.astype(np.uint32)
        for order in (True, False):
            r, b = (c[..., 0], c[..., 2]) if order else (c[..., 2], c[..., 0])
            want = ((r * 9798 + c[..., 1] * 19235 + b * 3735 + 16384) >> 15).astype(np.uint8)
            got, _ = convert_rgbd(ctx, img, order, None)
            assert np.array_equal(got, want), (w, order)


@pytest.mark.parametrize("cam", ["tum1", "tum2", "k1only", "strong"])
def test_undistort_keypoints_vs_cv2_golden(orc, cam):
    """Frame::UndistortKeyPoints / ComputeImageBounds (Frame.cc:1062-1092, 1135-1163) against real cv2.undistortPoints."""
    import ctypes as C

    import torch
    from conftest import load_golden

    from psl_slam_b200 import Context, _lib, default_config, image_bounds, make_distortion, undistort_keypoints
    from psl_slam_b200._lib import KP_DTYPE
    g = load_golden("undistort")
    d = make_distortion(*[float(v) for v in g[f"cam_{cam}"]])
    ctx = Context(default_config())

    def kp_array(xy):
        k = np.zeros(len(xy), KP_DTYPE)
        k["x"], k["y"] = xy[:, 0], xy[:, 1]
        k["size"], k["angle"], k["octave"], k["class_id"] = 31.0, 12.5, 2, -1
        return k

    for src, want in ((g["pts_kp"], g[f"kp_{cam}"]), (g["pts_rnd"], g[f"rnd_{cam}"])):
        k = kp_array(src)
        out = undistort_keypoints(ctx, k, d)
        assert np.array_equal(np.stack([out["x"], out["y"]], 1), want)
        assert out.tobytes() == orc.undistort_keypoints(k, d).tobytes()
    assert np.array_equal(np.array(image_bounds(ctx, 640, 480, d), np.float32), g[f"bounds_{cam}"])
    z = make_distortion(500, 500, 320, 240, 0.0, 0.3, 0.01, 0.01, 0.1)
    k = kp_array(g["pts_rnd"])
    assert undistort_keypoints(ctx, k, z).tobytes() == k.tobytes()
    assert image_bounds(ctx, 640, 480, z) == (0.0, 0.0, 640.0, 480.0)
    assert len(undistort_keypoints(ctx, k[:0], d)) == 0
    # batched device form: 3 frames with different counts, rows beyond a frame's count are left alone
    cap, ns = 900, [900, 0, 417]
    k3 = np.stack([kp_array(g["pts_rnd"][i * 900:(i + 1) * 900]) for i in range(3)])
    dk = torch.from_numpy(k3.view(np.uint8).reshape(3, cap, 28)).cuda()
    dn = torch.tensor(ns, dtype=torch.int32, device="cuda")
    do = torch.full((3, cap, 28), 0xAB, dtype=torch.uint8, device="cuda")
    ctx.check(_lib.lib().psl_undistort_keypoints_dev(ctx.handle, dk.data_ptr(), dn.data_ptr(), cap, 3, C.byref(d), do.data_ptr()))
    ctx.sync()
    got = do.cpu().numpy()
    for b, n in enumerate(ns):
        want = orc.undistort_keypoints(k3[b, :n], d)
        assert got[b, :n].tobytes() == want.tobytes()
        assert (got[b, n:] == 0xAB).all()


def test_pose_after_matching_vs_oracle_chain(orc):
    """The second half of TrackWithMotionModel (Tracking.cc:1193-1240) on the device arrays of the batched front end:
    matches -> MapPoints of the previous frame (UnprojectStereo) -> PoseOptimization, against the same chain composed
    from the oracle's pieces (pose to 1e-6, identical outlier flags and inlier counts)."""
    import ctypes as C

    import torch
    from psl_slam_b200 import ORBextractor, make_camera, make_track_params, synth, track_orb_batch
    from psl_slam_b200._lib import lib
    K = synth.ICL
    B = 5
    gray, depth, T = synth.sequence(9, B)
    T = T.astype(np.float32)
    rng = np.random.default_rng(3)
    T[1:, :3, 3] += rng.normal(0, 0.004, (B - 1, 3)).astype(np.float32)   # the motion model is never exact
    ex = ORBextractor()
    cam = make_camera(K["fx"], K["fy"], K["cx"], K["cy"], K["bf"], K["depth_factor"])
    out = track_orb_batch(ex, gray, depth, T, cam, make_track_params(15.0, 0.9, True))
    cap = ex.cap
    dev = {k: torch.from_numpy(np.ascontiguousarray(out[k]).view(np.uint8).reshape(-1)).cuda()
           for k in ("kps", "u_right", "z", "assign", "n")}
    T12 = np.ascontiguousarray(T[:, :3, :4].reshape(B, 12))
    d_T = torch.from_numpy(T12).cuda()
    d_To = torch.zeros((B, 16), dtype=torch.float32, device="cuda")
    d_out = torch.zeros((B, cap), dtype=torch.uint8, device="cuda")
    d_cnt = torch.zeros(B, dtype=torch.int32, device="cuda")
    ex.ctx.check(lib().psl_track_pose_batch_dev(ex.ctx.handle, dev["kps"].data_ptr(), dev["u_right"].data_ptr(),
                                                dev["z"].data_ptr(), dev["assign"].data_ptr(), dev["n"].data_ptr(), cap, B,
                                                d_T.data_ptr(), C.addressof(cam), d_To.data_ptr(), d_out.data_ptr(),
                                                d_cnt.data_ptr()))
    ex.ctx.sync()
    To, outl, cnt = d_To.cpu().numpy().reshape(B, 4, 4), d_out.cpu().numpy(), d_cnt.cpu().numpy()
    assert cnt[0] == 0 and np.array_equal(To[0], T[0])
    fx, fy, cx, cy, bf = (np.float32(K[k]) for k in ("fx", "fy", "cx", "cy", "bf"))
    inv_sigma2 = ex.GetInverseScaleSigmaSquares()
    f32, f64 = np.float32, np.float64
    for b in range(1, B):
        n, nl = int(out["n"][b]), int(out["n"][b - 1])
        kps, kl = out["kps"][b, :n], out["kps"][b - 1, :nl]
        assign, zl = out["assign"][b, :n], out["z"][b - 1, :nl]
        pts = np.zeros(n, orc.POSE_POINT_DTYPE)
        pts["u"], pts["v"], pts["u_right"] = kps["x"], kps["y"], out["u_right"][b, :n]
        pts["inv_sigma2"] = inv_sigma2[kps["octave"]]
        m = np.nonzero(assign >= 0)[0]
        j = assign[m]
        assert (zl[j] > 0).all()
        # Last.UnprojectStereo(j): fp32 steps, the 3x3 products accumulated in double and rounded once
        Rcw, tcw = T[b - 1, :3, :3], T[b - 1, :3, 3]
        Ow = (-(Rcw.T.astype(f64)) @ tcw.astype(f64)).astype(f32)
        invfx, invfy = f32(1) / fx, f32(1) / fy
        xc = np.stack([((kl["x"][j] - cx) * zl[j]) * invfx, ((kl["y"][j] - cy) * zl[j]) * invfy, zl[j]], 1).astype(f32)
        Xw = (xc.astype(f64) @ Rcw.astype(f64) + Ow.astype(f64)).astype(f32)      # Rwc x = Rcw^T x
        pts["xw"][m], pts["yw"][m], pts["zw"][m] = Xw[:, 0], Xw[:, 1], Xw[:, 2]
        pts["flags"][m] = 1
        wT, wout, wcnt = orc.pose_optimization(T[b], pts, fx, fy, cx, cy, bf)
        assert cnt[b] == wcnt and wcnt > 200, (b, cnt[b], wcnt)
        assert np.array_equal(outl[b, :n], wout), b
        assert np.allclose(To[b], wT, rtol=0, atol=1e-6), b


def test_extract_lsd_chain_on_device_vs_oracle_pieces(orc):
    """Frame::ExtractLSD (Frame.cc:489-645) chained on the device for a batch -- line extraction -> depth conversion ->
    3-D lines -> junctions, nothing in between touching the host -- against the same chain composed from the oracle's
    pieces on the frames one by one; the plane hypotheses (host entry point) then run on the rows of the device arrays."""
    import torch

    from psl_slam_b200 import LINEextractor, make_camera, plane_hypotheses, synth
    from psl_slam_b200._lib import JUNCTION_DTYPE, KEYLINE_DTYPE
    from psl_slam_b200.tracking import extract_lsd_batch_dev
    K = synth.ICL
    B = 3
    gray, depth, _ = synth.sequence(11, B)
    H, W = gray.shape[1:]
    lex = LINEextractor()
    cam = make_camera(K["fx"], K["fy"], K["cx"], K["cy"], K["bf"], K["depth_factor"])
    d_gray, d_depth = torch.from_numpy(gray).cuda(), torch.from_numpy(depth.view(np.int16)).cuda()
    out = extract_lsd_batch_dev(lex, d_gray.data_ptr(), d_depth.data_ptr(), B, W, H, cam, seed=7)
    lex.ctx.sync()
    n = out["n"].cpu().numpy()
    kl_all = out["kl"].cpu().numpy().reshape(B, -1).view(KEYLINE_DTYPE)
    js_all = out["junctions"].cpu().numpy().reshape(B, -1).view(JUNCTION_DTYPE)
    planes_seen = 0
    for b in range(B):
        kl, ld, eq, _ = orc.line_extract(gray[b])
        assert n[b] == len(kl) and kl_all[b, : n[b]].tobytes() == kl.tobytes()
        assert np.array_equal(out["ldesc"][b, : n[b]].cpu().numpy(), ld)
        depf = depth[b].astype(np.float32) * np.float32(cam.depth_factor)   # convertTo(CV_32F, factor): one fp32 product
        assert np.array_equal(out["depth"][b].cpu().numpy(), depf)
        l3, eq3 = orc.lines_3d(kl, depf, cam.fx, cam.fy, cam.cx, cam.cy, 7)
        assert np.array_equal(out["lines3d"][b, : n[b]].cpu().numpy(), l3)
        assert np.array_equal(out["line_eq3"][b, : n[b]].cpu().numpy(), eq3)
        fans, js = orc.line_junctions(kl, l3, W, H)
        nf, nj = int(out["n_fans"][b]), int(out["n_junctions"][b])
        assert nf == len(fans) and nj == len(js)
        assert np.array_equal(out["fans"][b, :nf].cpu().numpy(), fans)
        assert js_all[b, :nj].tobytes() == np.ascontiguousarray(js, JUNCTION_DTYPE).tobytes()
        got = plane_hypotheses(lex.ctx, kl_all[b, : n[b]], out["line_eq3"][b, : n[b]].cpu().numpy(),
                               out["lines3d"][b, : n[b]].cpu().numpy(), js_all[b, :nj])
        want = orc.plane_hypotheses(kl, eq3, l3, js)
        for g, w_ in zip(got, want):
            assert np.array_equal(g, w_)
        planes_seen += len(got[1])
    assert n.min() > 10


def test_bench_sequence_both_extractors_vs_oracle(orc):
    """Frames of the bench's own workload (the GPU-rendered cfg-4 sequence: ~1100 raw LSD segments and ~5 k FAST
    candidates per frame, heavier than the goldens) through both batched extractors, every frame against the oracle."""
    import os
    import sys

    import torch
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    from psl_slam_b200 import KP_DTYPE, LINEextractor, ORBextractor
    from psl_slam_b200._lib import KEYLINE_DTYPE
    B, W, H = 40, 640, 480
    rgb, _, _ = bench.render_sequence_cuda(4, B, W, H, torch.device("cuda"))
    d_gray = bench.gray_cuda(rgb).contiguous()
    gray = d_gray.cpu().numpy()
    ex, lex = ORBextractor(), LINEextractor(chunk_frames=B)
    cap, lcap = ex.cap, lex.cap
    kps = torch.zeros((B, cap, KP_DTYPE.itemsize), dtype=torch.uint8, device="cuda")
    desc = torch.zeros((B, cap, 32), dtype=torch.uint8, device="cuda")
    n = torch.zeros(B, dtype=torch.int32, device="cuda")
    ex.extract_batch_dev(d_gray.data_ptr(), B, W, H, W, W * H, kps.data_ptr(), desc.data_ptr(), n.data_ptr())
    ex.ctx.sync()
    kl = torch.zeros((B, lcap, KEYLINE_DTYPE.itemsize), dtype=torch.uint8, device="cuda")
    ld = torch.zeros((B, lcap, 32), dtype=torch.uint8, device="cuda")
    eq = torch.zeros((B, lcap, 3), dtype=torch.float64, device="cuda")
    nl = torch.zeros(B, dtype=torch.int32, device="cuda")
    lex.extract_batch_dev(d_gray.data_ptr(), B, W, H, W, W * H, kl.data_ptr(), ld.data_ptr(), eq.data_ptr(), None, nl.data_ptr())
    lex.ctx.sync()
    n, nl = n.cpu().numpy(), nl.cpu().numpy()
    kps, desc = kps.cpu().numpy().reshape(B, -1).view(KP_DTYPE), desc.cpu().numpy()
    kl, ld, eq = kl.cpu().numpy().reshape(B, -1).view(KEYLINE_DTYPE), ld.cpu().numpy(), eq.cpu().numpy()
    for b in range(0, B, 3):   # (the oracle takes ~50 ms per frame: every third frame keeps the test short)
        wk, wd = orc.orb_extract(gray[b])
        assert n[b] == len(wk) and kps[b, : n[b]].tobytes() == wk.tobytes() and np.array_equal(desc[b, : n[b]], wd), b
        wl, wld, weq, _ = orc.line_extract(gray[b])
        assert nl[b] == len(wl) and kl[b, : nl[b]].tobytes() == wl.tobytes(), b
        assert np.array_equal(ld[b, : nl[b]], wld) and np.array_equal(eq[b, : nl[b]], weq), b
    assert n.min() > 900 and nl.min() > 40
