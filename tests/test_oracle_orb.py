"""CPU: the self-contained oracle (oracle/c) against the cv2-primitive goldens (tests/golden)."""
import numpy as np
import pytest

from conftest import golden_names, load_golden

ORB = golden_names("orb_")


def _params(orc, g):
    nf, nl, ini, mn = [int(v) for v in g["params"]]
    return orc.params(nf, float(g["scale_factor"]), nl, ini, mn)


def test_tables_vga(orc):
    scale, inv, quota, umax = orc.orb_tables(orc.params())
    assert quota.tolist() == [217, 181, 151, 126, 105, 87, 73, 60]
    assert umax.tolist() == [15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3]
    assert scale[1] == np.float32(1.2) and abs(float(scale[7]) - 3.5831816196) < 1e-6


def test_fast_atan2_known_answers(orc):
    assert orc.fast_atan2(0.0, 0.0) == 0.0
    assert orc.fast_atan2(0.0, 5.0) == 0.0
    assert abs(orc.fast_atan2(1.0, 1.0) - 45.0) < 0.02
    assert abs(orc.fast_atan2(-1.0, -1.0) - 225.0) < 0.02


@pytest.mark.parametrize("name", ORB)
def test_resize_chain_and_blur(orc, name):
    g = load_golden(name)
    p = _params(orc, g)
    _, inv, _, _ = orc.orb_tables(p)
    img = g["image"]
    lvl = int(g["level_idx"])
    cur = img
    for l in range(1, lvl + 1):
        w = int(np.rint(np.float32(img.shape[1]) * inv[l]))
        h = int(np.rint(np.float32(img.shape[0]) * inv[l]))
        cur = orc.resize_linear(cur, w, h)
    assert np.array_equal(cur, g["level_img"])
    if "level_blur" in g:
        assert np.array_equal(orc.gauss_blur(cur, 7), g["level_blur"])


@pytest.mark.parametrize("name", ORB)
def test_stage_candidates_and_octree(orc, name):
    g = load_golden(name)
    p = _params(orc, g)
    _, inv, quota, _ = orc.orb_tables(p)
    cur = g["image"]
    for l in range(p.nlevels):
        if l:
            w = int(np.rint(np.float32(g["image"].shape[1]) * inv[l]))
            h = int(np.rint(np.float32(g["image"].shape[0]) * inv[l]))
            cur = orc.resize_linear(cur, w, h)
        cands = orc.fast_cells(cur, p.ini_th, p.min_th)
        assert np.array_equal(cands, g[f"cands_{l}"]), f"level {l} candidates"
        if len(cands):
            sel = orc.octree(cands, 16, cur.shape[1] - 16, 16, cur.shape[0] - 16, int(quota[l]))
            gs = g[f"sel_{l}"]
            assert np.array_equal(sel[:, 0] + 16, gs[:, 0]) and np.array_equal(sel[:, 1] + 16, gs[:, 1])
            assert np.array_equal(sel[:, 2], gs[:, 2])


@pytest.mark.parametrize("name", ORB)
def test_orb_extract_end_to_end(orc, name):
    g = load_golden(name)
    kps, desc = orc.orb_extract(g["image"], _params(orc, g))
    gk = g["kps"]
    assert len(kps) == len(gk)
    for i, f in enumerate(["x", "y", "size", "angle", "response"]):
        assert np.array_equal(kps[f], gk[:, i]), f
    assert np.array_equal(kps["octave"], g["octave"])
    assert np.all(kps["class_id"] == -1)
    assert np.array_equal(desc, g["desc"])


def test_empty_image(orc):
    kps, desc = orc.orb_extract(np.zeros((0, 0), np.uint8).reshape(0, 0))
    assert len(kps) == 0 and desc.shape == (0, 32)


def test_gray_conversion_formula_vs_cv2_golden():
    """The Q15 RGB->Y restatement used by the synthetic generator against real cv2.cvtColor."""
    from conftest import load_golden
    from psl_slam_b200 import synth
    g = load_golden("convert_rgbd")
    assert np.array_equal(synth.rgb_to_gray(g["rgb"]), g["gray_rgb"])
    assert np.array_equal(synth.rgb_to_gray(g["rgb"][..., ::-1]), g["gray_bgr"])
    assert np.array_equal(synth.rgb_to_gray(g["rgba"][..., :3]), g["gray_rgba"])
    assert np.array_equal((g["depth"].astype(np.float32) * g["factor"]).astype(np.float32), g["depth_f"])
