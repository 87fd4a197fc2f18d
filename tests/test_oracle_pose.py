"""CPU: the pose-optimisation oracle (oracle/c/orc_pose.cpp; "next" row N4, point edges only — oracle side only this
round, there is no product kernel yet) against the goldens of the independent numpy restatement
(oracle/pyref/pose_py.py).  Neither is pinned by the reference (g2o needs Eigen, which is not in the image): the two
restatements pin each other.  Tolerance: 1e-6 on the float32 pose entries (fp64 sums in a different order, float32
output), identical outlier flags and counts."""
import numpy as np
import pytest

from conftest import golden_names, load_golden


@pytest.mark.parametrize("name", golden_names("pose_case"))
def test_pose_optimization_vs_golden(orc, name):
    g = load_golden(name)
    fx, fy, cx, cy, bf = (float(v) for v in g["cam"])
    T, outlier, count = orc.pose_optimization(g["Tcw0"], g["pts"], fx, fy, cx, cy, bf)
    assert count == int(g["count"])
    assert np.array_equal(outlier, g["outlier"])
    assert np.allclose(T, g["Tcw"], rtol=0, atol=1e-6)
    npts = int((g["pts"]["flags"] & 1).sum())
    if npts < 3:     # Optimizer.cc:592-593: nothing is optimised, the pose stays
        assert count == 0 and np.array_equal(T, g["Tcw0"])
    if npts >= 100:  # the optimisation does what it is for: the pose moves from the prior towards the true one
        err0 = np.abs(g["Tcw0"][:3, 3] - g["Ttrue"][:3, 3]).max()
        err1 = np.abs(T[:3, 3] - g["Ttrue"][:3, 3]).max()
        assert err1 < 0.1 * err0 and err1 < 5e-3
        assert outlier[(g["pts"]["flags"] & 1) == 0].sum() == 0


@pytest.mark.parametrize("name", golden_names("pose_lil_case"))
def test_pose_optimization_with_lil_edges_vs_golden(orc, name):
    """The complete PoseOptimization: point edges + EdgeLILSE3ProjectXYZ (Optimizer.cc:619-693, 976-1007; EdgeLIL.h:210-374)."""
    g = load_golden(name)
    fx, fy, cx, cy, bf = (float(v) for v in g["cam"])
    T, outlier, count, lil_outlier = orc.pose_optimization(g["Tcw0"], g["pts"], fx, fy, cx, cy, bf, g["lils"])
    assert count == int(g["count"])
    assert np.array_equal(outlier, g["outlier"]) and np.array_equal(lil_outlier, g["lil_outlier"])
    assert np.allclose(T, g["Tcw"], rtol=0, atol=1e-6)
    assert lil_outlier[(g["lils"]["flags"] & 1) == 0].sum() == 0
    err0 = np.abs(g["Tcw0"][:3, 3] - g["Ttrue"][:3, 3]).max()
    err1 = np.abs(T[:3, 3] - g["Ttrue"][:3, 3]).max()
    assert err1 < 0.2 * err0


def test_lil_edges_change_the_result(orc):
    """Without the structural-line edges the same frame gives another pose (what ADVICE r01 pointed at), and a frame
    whose points alone are too few is only optimised through them."""
    g = load_golden("pose_lil_case2_few_points")
    fx, fy, cx, cy, bf = (float(v) for v in g["cam"])
    T_pts, _, _ = orc.pose_optimization(g["Tcw0"], g["pts"], fx, fy, cx, cy, bf)
    assert np.abs(T_pts - g["Tcw"]).max() > 1e-4
    g = load_golden("pose_lil_case1_lines_only")
    T_pts, _, c = orc.pose_optimization(g["Tcw0"], g["pts"], fx, fy, cx, cy, bf)
    assert c == 0 and np.array_equal(T_pts, g["Tcw0"]) and int(g["count"]) == 9
