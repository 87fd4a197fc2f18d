"""GPU: BASELINE config 5 shaped case — 1920x1080 frames, 4000 ORB features, 12 pyramid levels, dense line output —
through the C-ABI against the oracle (bit-exact), plus size-independent properties of the full-size outputs."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def hd_frames():
    from psl_slam_b200 import synth
    poster = synth.make_poster(5, 4096, 9000)
    T = synth.trajectory(2, 5)
    out = []
    for i in range(2):
        rgb, dep = synth.render(poster, T[i], 1920, 1080, metres_per_px=0.0011, noise_seed=50 + i)
        out.append((synth.rgb_to_gray(rgb), dep))
    return out, T


def test_hd_orb_vs_oracle(orc, hd_frames):
    from psl_slam_b200 import ORBextractor
    (g0, _), (g1, _) = hd_frames[0]
    ex = ORBextractor(4000, 1.2, 12, 20, 7, max_width=1920, max_height=1080, chunk_frames=2, max_candidates=262144)
    kps, desc, n = ex.extract_batch(np.stack([g0, g1]))
    p = orc.params(4000, 1.2, 12, 20, 7)
    for b, g in enumerate((g0, g1)):
        okps, odesc = orc.orb_extract(g, p)
        assert n[b] == len(okps) and 4000 <= n[b] <= 4000 + 3 * 12
        for f in okps.dtype.names:
            assert np.array_equal(kps[b, : n[b]][f], okps[f]), f
        assert np.array_equal(desc[b, : n[b]], odesc)
        # properties: octaves ascending, coordinates inside the image minus the 19-px border, no duplicates per level
        k = kps[b, : n[b]]
        assert np.all(np.diff(k["octave"]) >= 0) and k["octave"].max() == 11
        assert k["x"].min() >= 19 and k["x"].max() <= 1920 - 19 and k["y"].min() >= 19 and k["y"].max() <= 1080 - 19
        key = np.stack([k["octave"].astype(np.float64), k["x"], k["y"]], 1)
        assert len(np.unique(key, axis=0)) == len(k)


def test_hd_lines_vs_oracle(orc, hd_frames):
    from psl_slam_b200 import LINEextractor
    (g0, _), _ = hd_frames[0]
    ex = LINEextractor(1, 1.2, 1000, 0.0, max_width=1920, max_height=1080, chunk_frames=1, max_raw=32768)
    kl, ld, eq, lbd = ex(g0, with_lbd_floats=True)
    okl, old, oeq, olbd = orc.line_extract(g0, 1000, cap=4096)
    assert len(kl) == len(okl) and len(kl) > 200
    for f in kl.dtype.names:
        assert np.array_equal(kl[f], okl[f]), f
    assert np.array_equal(ld, old) and np.array_equal(eq, oeq) and np.array_equal(lbd, olbd, equal_nan=True)
    # properties: every kept line is longer than 50 px and (merged endpoints are not re-clamped by the reference)
    # within a few pixels of the image; unit-normal line equation through both ends
    assert np.all(kl["line_length"] > 50.0)
    assert kl["start_x"].min() > -5 and kl["end_x"].max() < 1925 and kl["start_y"].min() > -5 and kl["end_y"].max() < 1085
    assert np.allclose(np.hypot(eq[:, 0], eq[:, 1]), 1.0, atol=1e-12)
    res = eq[:, 0] * kl["start_x"] + eq[:, 1] * kl["start_y"] + eq[:, 2]
    assert np.abs(res).max() < 1e-3
