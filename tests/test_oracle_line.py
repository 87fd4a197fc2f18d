"""CPU: the line oracle (oracle/c/orc_lsd.cpp, orc_line.cpp) against goldens made from the real cv2 LSD /
blur / Sobel / clipLine and the independent Python restatement of merge + LBD (oracle/pyref/line_py.py)."""
import numpy as np
import pytest

from conftest import golden_names, load_golden

LINES = golden_names("line_")


def test_lsd_known_answers(orc):
    """analytic images measured with cv2 4.13."""
    im = np.full((480, 640), 60, np.uint8)
    im[:, 200:] = 180
    assert np.array_equal(orc.lsd_detect(im), np.array([[199.375, 0.625, 199.375, 478.125]], np.float32))
    im = np.full((480, 640), 60, np.uint8)
    im[120:360, 100:400] = 200
    seg = orc.lsd_detect(im)
    assert seg.shape == (4, 4)
    assert np.allclose(seg[0], [398.125, 119.373, 100.625, 119.373], atol=1e-3)  # direction encodes polarity
    assert len(orc.lsd_detect(np.full((100, 100), 7, np.uint8))) == 0


@pytest.mark.parametrize("name", LINES)
def test_lsd_vs_cv2(orc, name):
    g = load_golden(name)
    assert np.array_equal(orc.lsd_detect(g["image"]), g["lsd_raw"])


def test_line_iterator_count_known_answers(orc):
    assert orc.line_iterator_count(640, 480, 10.4, 10.4, 20.6, 13.0) == 12      # rounds to (10,10)-(21,13)
    assert orc.line_iterator_count(640, 480, -50.0, 100.0, 50.0, 100.0) == 51   # clipped at x = 0
    assert orc.line_iterator_count(640, 480, -50.0, -10.0, -5.0, -2.0) == 0     # entirely outside
    assert orc.line_iterator_count(640, 480, 5.0, 5.0, 5.0, 5.0) == 1


@pytest.mark.parametrize("name", LINES)
def test_merge_keylines_lbd(orc, name):
    g = load_golden(name)
    img = g["image"]
    h, w = img.shape
    merged = orc.merge_lines_lsd(orc.clamp_segments(g["lsd_raw"], w, h))
    assert np.array_equal(merged, g["merged"])
    kl = orc.make_keylines(merged, w, h, int(g["nfeatures"]))
    gk = g["keylines"]
    assert len(kl) == len(gk)
    for f in kl.dtype.names:
        assert np.array_equal(kl[f], gk[f]), f
    if len(kl):
        dx, dy = orc.lbd_gradients(img)
        des, bits = orc.lbd_descriptors(dx, dy, kl)
        assert np.array_equal(des, g["lbd"], equal_nan=True)
        assert np.array_equal(bits, g["ldesc"])


@pytest.mark.parametrize("name", LINES)
def test_line_extract_end_to_end(orc, name):
    g = load_golden(name)
    kl, ld, eq, lbd = orc.line_extract(g["image"], int(g["nfeatures"]))
    assert len(kl) == len(g["keylines"])
    for f in kl.dtype.names:
        assert np.array_equal(kl[f], g["keylines"][f]), f
    assert np.array_equal(ld, g["ldesc"]) and np.array_equal(lbd, g["lbd"], equal_nan=True)
    assert np.array_equal(eq, g["lineeq"])
    if len(kl):
        assert np.all(kl["line_length"] > 50.0) and np.all(kl["octave"] == 0)
        assert np.array_equal(kl["class_id"], np.arange(len(kl)))


def test_merge_empty_and_single(orc):
    assert orc.merge_lines_lsd(np.zeros((0, 4), np.float32)).shape == (0, 4)
    one = np.array([[10, 10, 200, 12]], np.float32)
    out = orc.merge_lines_lsd(one)
    assert out.shape == (1, 4) and np.allclose(out, one, atol=1e-3)
    # two collinear halves merge into one long segment
    two = np.array([[10, 100, 150, 100], [160, 100.5, 300, 100.5]], np.float32)
    out = orc.merge_lines_lsd(two)
    assert out.shape == (1, 4) and abs(out[0, 2] - out[0, 0]) > 280
