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
