"""GPU: the CUDA line extractor (LSD + merge + LBD) through the C-ABI against the oracle and the committed
cv2 goldens.  Integer / byte results and every float field are compared bit-exactly."""
import numpy as np
import pytest

from conftest import golden_names, load_golden

pytestmark = pytest.mark.gpu
LINES = golden_names("line_")


def _same(got, want):
    kl, ld, eq, lbd = got
    okl, old, oeq, olbd = want
    assert len(kl) == len(okl)
    for f in kl.dtype.names:
        assert np.array_equal(kl[f], okl[f]), f
    assert np.array_equal(ld, old)
    assert np.array_equal(eq, oeq)
    assert np.array_equal(lbd, olbd, equal_nan=True)


@pytest.mark.parametrize("name", LINES)
def test_line_vs_golden_and_stages(orc, name):
    from psl_slam_b200 import LINEextractor
    g = load_golden(name)
    img = g["image"]
    h, w = img.shape
    ex = LINEextractor(1, 1.2, int(g["nfeatures"]), 0.0, max_width=w, max_height=h, chunk_frames=2)
    got = ex(img, with_lbd_floats=True)
    _same(got, (g["keylines"], g["ldesc"], g["lineeq"], g["lbd"]))
    # stages: the 0.8x working image and the raw LSD segments (cv2.createLineSegmentDetector output)
    scaled = orc.lsd_scaled_image(img)
    assert np.array_equal(ex.debug_fetch(4).reshape(scaled.shape), scaled)
    assert np.array_equal(ex.debug_fetch(5), orc.clamp_segments(g["lsd_raw"], w, h))


def test_lsd_known_answers():
    """SURVEY App. B.5 (measured with cv2 4.13): a vertical step edge gives exactly one segment."""
    from psl_slam_b200 import LINEextractor
    ex = LINEextractor(chunk_frames=1)
    im = np.full((480, 640), 60, np.uint8)
    im[:, 200:] = 180
    kl, ld, eq = ex(im)
    assert np.array_equal(ex.debug_fetch(5), np.array([[199.375, 0.625, 199.375, 478.125]], np.float32))
    assert len(kl) == 1 and kl["num_pixels"][0] == 478 and ld.shape == (1, 32)
    assert np.allclose(np.abs(eq[0]), [1.0, 0.0, 199.375])
    kl, ld, eq = ex(np.full((480, 640), 7, np.uint8))
    assert len(kl) == 0 and ld.shape == (0, 32) and eq.shape == (0, 3)
    kl, ld, eq = ex(np.zeros((0, 0), np.uint8))  # empty image: silent return (LineExtractor.cpp:327-328)
    assert len(kl) == 0


def test_batch_mixed_frames_vs_oracle(orc):
    """A batch spanning several chunks: textured, low-texture and flat frames side by side."""
    from psl_slam_b200 import LINEextractor, synth
    gray, _, _ = synth.sequence(6, 3)
    frames = np.stack([gray[0], synth.make_lowtex(11), np.full((480, 640), 100, np.uint8), gray[1],
                       synth.make_lowtex(12), gray[2], synth.make_lowtex(13)])
    ex = LINEextractor(chunk_frames=3)
    kl, ld, eq, lbd, n = ex.extract_batch(frames, with_lbd_floats=True)
    assert n[2] == 0 and n.max() > 20
    for b in range(len(frames)):
        want = orc.line_extract(frames[b], 200)
        _same((kl[b, : n[b]], ld[b, : n[b]], eq[b, : n[b]], lbd[b, : n[b]]), want)
    # same frames, one chunk, strided rows: identical bytes
    padded = np.zeros((len(frames), 480, 704), np.uint8)
    padded[:, :, :640] = frames
    ex2 = LINEextractor(chunk_frames=16)
    kl2, ld2, eq2, n2 = ex2.extract_batch(padded[:, :, :640])
    assert np.array_equal(n, n2)
    for b in range(len(frames)):  # rows past n[b] are unspecified
        assert kl[b, : n[b]].tobytes() == kl2[b, : n[b]].tobytes() and np.array_equal(ld[b, : n[b]], ld2[b, : n[b]])
        assert np.array_equal(eq[b, : n[b]], eq2[b, : n[b]])


def test_capacity_errors_are_loud():
    from psl_slam_b200 import LINEextractor, PslError, synth
    gray, _, _ = synth.sequence(6, 1)
    ex = LINEextractor(max_raw=64, chunk_frames=1)
    with pytest.raises(PslError) as e:
        ex(gray[0])
    assert e.value.code == -3
    with pytest.raises(PslError):
        LINEextractor(max_width=320, max_height=240)(gray[0])


def test_auto_capacity_grows_and_reruns():
    """line_max_raw <= 0: the context owns the bound.  A frame with more raw segments than the starting value makes the
    host-pointer call grow the buffers and run again by itself; the result is the one of a roomy context."""
    from psl_slam_b200 import LINEextractor, synth
    gray, _, _ = synth.sequence(6, 2)
    kl0, ld0, eq0 = LINEextractor()(gray[0])
    ex = LINEextractor(max_raw=-16, chunk_frames=2)
    for g in gray:   # the second call starts with the grown bound
        kl, ld, eq = ex(g)
    kl, ld, eq = ex(gray[0])
    assert kl.tobytes() == kl0.tobytes() and np.array_equal(ld, ld0) and np.array_equal(eq, eq0)


def test_unaligned_device_frames_vs_oracle(orc):
    """Device-pointer entry with an odd base address and row stride: the byte-load form of the blur kernel (LSD
    prologue, LBD) must reproduce the word form and the oracle."""
    import torch

    from psl_slam_b200 import KEYLINE_DTYPE, LINEextractor, synth
    frames = np.stack([synth.sequence(8, 1)[0][0], synth.make_lowtex(31)])
    B, (H, W) = len(frames), frames.shape[1:]
    stride, off = W + 3, 1
    fs = stride * H + 5
    host = np.zeros(B * fs + 16, np.uint8)
    for b in range(B):
        for y in range(H):
            s = off + b * fs + y * stride
            host[s:s + W] = frames[b, y]
    buf = torch.from_numpy(host).cuda()
    ex = LINEextractor(chunk_frames=4)
    cap = ex.cap
    kl = torch.zeros((B, cap, 68), dtype=torch.uint8, device="cuda")
    ld = torch.zeros((B, cap, 32), dtype=torch.uint8, device="cuda")
    eq = torch.zeros((B, cap, 3), dtype=torch.float64, device="cuda")
    n = torch.zeros(B, dtype=torch.int32, device="cuda")
    ex.extract_batch_dev(buf.data_ptr() + off, B, W, H, stride, fs, kl.data_ptr(), ld.data_ptr(), eq.data_ptr(), None,
                         n.data_ptr())
    ex.ctx.sync()
    n_h = n.cpu().numpy()
    for b in range(B):
        okl, old, oeq, _ = orc.line_extract(frames[b], 200)
        assert n_h[b] == len(okl) and n_h[b] > 5
        assert kl[b, : n_h[b]].cpu().numpy().tobytes() == okl.tobytes()
        assert np.array_equal(ld[b, : n_h[b]].cpu().numpy(), old) and np.array_equal(eq[b, : n_h[b]].cpu().numpy(), oeq)


def test_odd_sizes_vs_oracle(orc):
    """Widths that are not multiples of 4 (image and its 0.8x copy): the per-pixel forms of the exact resize and the
    Sobel kernels, and the edge columns of the blur."""
    from psl_slam_b200 import LINEextractor, synth
    base = synth.make_lowtex(41)
    for (w, h) in ((322, 241), (431, 303)):
        img = np.ascontiguousarray(base[:h, :w])
        ex = LINEextractor(max_width=w, max_height=h)
        kl, ld, eq = ex(img)
        okl, old, oeq, _ = orc.line_extract(img, 200)
        assert len(kl) == len(okl) and len(kl) > 3
        assert kl.tobytes() == okl.tobytes() and np.array_equal(ld, old) and np.array_equal(eq, oeq)
