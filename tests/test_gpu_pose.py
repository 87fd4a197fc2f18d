"""GPU: Optimizer::PoseOptimization (point edges; "next" row N4) through the C-ABI against the oracle and the goldens of
the independent numpy restatement.  Contract: fp64 tolerance, not bit-exactness — the pose within 1e-6
(float32 entries), identical outlier flags and counts on these cases (no residual sits within 1e-6 of a threshold)."""
import numpy as np
import pytest

from conftest import golden_names, load_golden

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", golden_names("pose_case"))
def test_pose_optimization_vs_golden_and_oracle(orc, name):
    from psl_slam_b200 import Context, PoseOptimization, default_config
    g = load_golden(name)
    fx, fy, cx, cy, bf = (float(v) for v in g["cam"])
    ctx = Context(default_config())
    T, outlier, count = PoseOptimization(ctx, g["Tcw0"], g["pts"], fx, fy, cx, cy, bf)
    wT, wout, wcount = orc.pose_optimization(g["Tcw0"], g["pts"], fx, fy, cx, cy, bf)
    assert count == wcount == int(g["count"])
    assert np.array_equal(outlier, wout) and np.array_equal(outlier, g["outlier"])
    assert np.allclose(T, wT, rtol=0, atol=1e-6) and np.allclose(T, g["Tcw"], rtol=0, atol=1e-6)


def test_pose_optimization_batched_device(orc):
    import torch
    from psl_slam_b200 import Context, default_config
    from psl_slam_b200._lib import POSE_POINT_DTYPE, lib
    names = golden_names("pose_case")
    gs = [load_golden(n) for n in names]
    cap = max(len(g["pts"]) for g in gs) + 7
    B = len(gs)
    pts = np.zeros((B, cap), POSE_POINT_DTYPE)
    T0 = np.zeros((B, 16), np.float32)
    for b, g in enumerate(gs):
        pts[b, : len(g["pts"])] = g["pts"]
        T0[b] = g["Tcw0"].reshape(16)
    fx, fy, cx, cy, bf = (float(v) for v in gs[0]["cam"])
    ctx = Context(default_config())
    d_p = torch.from_numpy(pts.view(np.uint8).reshape(-1)).cuda()
    d_T = torch.from_numpy(T0).cuda()
    d_n = torch.tensor([len(g["pts"]) for g in gs], dtype=torch.int32, device="cuda")
    d_To = torch.zeros_like(d_T)
    d_o = torch.zeros(B * cap, dtype=torch.uint8, device="cuda")
    d_c = torch.zeros(B, dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    import ctypes as C
    ctx.check(lib().psl_pose_optimization_dev(ctx.handle, d_T.data_ptr(), d_p.data_ptr(), d_n.data_ptr(), cap, B, C.c_float(fx),
                                              C.c_float(fy), C.c_float(cx), C.c_float(cy), C.c_float(bf), d_To.data_ptr(),
                                              d_o.data_ptr(), d_c.data_ptr()))
    ctx.sync()
    To, o, c = d_To.cpu().numpy().reshape(B, 4, 4), d_o.cpu().numpy().reshape(B, cap), d_c.cpu().numpy()
    for b, g in enumerate(gs):
        assert c[b] == int(g["count"])
        assert np.array_equal(o[b, : len(g["pts"])], g["outlier"])
        assert np.allclose(To[b], g["Tcw"], rtol=0, atol=1e-6)


@pytest.mark.parametrize("name", golden_names("pose_lil_case"))
def test_pose_optimization_with_lil_edges(orc, name):
    """The complete PoseOptimization (points + EdgeLILSE3ProjectXYZ) through psl_pose_optimization_lil."""
    from psl_slam_b200 import Context, PoseOptimization, default_config
    g = load_golden(name)
    fx, fy, cx, cy, bf = (float(v) for v in g["cam"])
    ctx = Context(default_config())
    T, outlier, count, lout = PoseOptimization(ctx, g["Tcw0"], g["pts"], fx, fy, cx, cy, bf, g["lils"])
    wT, wout, wcount, wlout = orc.pose_optimization(g["Tcw0"], g["pts"], fx, fy, cx, cy, bf, g["lils"])
    assert count == wcount == int(g["count"])
    assert np.array_equal(outlier, wout) and np.array_equal(outlier, g["outlier"])
    assert np.array_equal(lout, wlout) and np.array_equal(lout, g["lil_outlier"])
    assert np.allclose(T, wT, rtol=0, atol=1e-6) and np.allclose(T, g["Tcw"], rtol=0, atol=1e-6)


def test_pose_optimization_lil_batched_device():
    import ctypes as C

    import torch
    from psl_slam_b200 import Context, default_config
    from psl_slam_b200._lib import POSE_LIL_DTYPE, POSE_POINT_DTYPE, lib
    gs = [load_golden(n) for n in golden_names("pose_lil_case")] + [load_golden("pose_case0")]
    B = len(gs)
    cap = max(len(g["pts"]) for g in gs) + 3
    lcap = max(len(g["lils"]) if "lils" in g else 0 for g in gs) + 2
    pts = np.zeros((B, cap), POSE_POINT_DTYPE)
    lils = np.zeros((B, lcap), POSE_LIL_DTYPE)
    T0 = np.zeros((B, 16), np.float32)
    nl = []
    for b, g in enumerate(gs):
        pts[b, : len(g["pts"])] = g["pts"]
        T0[b] = g["Tcw0"].reshape(16)
        if "lils" in g:
            lils[b, : len(g["lils"])] = g["lils"]
        nl.append(len(g["lils"]) if "lils" in g else 0)
    fx, fy, cx, cy, bf = (float(v) for v in gs[0]["cam"])
    ctx = Context(default_config())
    d_p = torch.from_numpy(pts.view(np.uint8).reshape(-1)).cuda()
    d_l = torch.from_numpy(lils.view(np.uint8).reshape(-1)).cuda()
    d_T = torch.from_numpy(T0).cuda()
    d_n = torch.tensor([len(g["pts"]) for g in gs], dtype=torch.int32, device="cuda")
    d_nl = torch.tensor(nl, dtype=torch.int32, device="cuda")
    d_To = torch.zeros_like(d_T)
    d_o = torch.zeros(B * cap, dtype=torch.uint8, device="cuda")
    d_lo = torch.zeros(B * lcap, dtype=torch.uint8, device="cuda")
    d_c = torch.zeros(B, dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    ctx.check(lib().psl_pose_optimization_lil_dev(ctx.handle, d_T.data_ptr(), d_p.data_ptr(), d_n.data_ptr(), cap, d_l.data_ptr(),
                                                  d_nl.data_ptr(), lcap, B, C.c_float(fx), C.c_float(fy), C.c_float(cx),
                                                  C.c_float(cy), C.c_float(bf), d_To.data_ptr(), d_o.data_ptr(),
                                                  d_lo.data_ptr(), d_c.data_ptr()))
    ctx.sync()
    To, o, c = d_To.cpu().numpy().reshape(B, 4, 4), d_o.cpu().numpy().reshape(B, cap), d_c.cpu().numpy()
    lo = d_lo.cpu().numpy().reshape(B, lcap)
    for b, g in enumerate(gs):
        assert c[b] == int(g["count"])
        assert np.array_equal(o[b, : len(g["pts"])], g["outlier"])
        if "lils" in g:
            assert np.array_equal(lo[b, : len(g["lils"])], g["lil_outlier"])
        assert np.allclose(To[b], g["Tcw"], rtol=0, atol=1e-6)
