"""GPU: the CUDA ORB extractor through the C-ABI against the oracle and the committed goldens."""
import numpy as np
import pytest

from conftest import golden_names, load_golden

pytestmark = pytest.mark.gpu
ORB = golden_names("orb_")


def _extractor(g):
    from psl_slam_b200 import ORBextractor
    nf, nl, ini, mn = [int(v) for v in g["params"]]
    h, w = g["image"].shape
    return ORBextractor(nf, float(g["scale_factor"]), nl, ini, mn, max_width=w, max_height=h)


def _same(kps, desc, okps, odesc):
    assert len(kps) == len(okps)
    for f in kps.dtype.names:
        assert np.array_equal(kps[f], okps[f]), f
    assert np.array_equal(desc, odesc)


@pytest.mark.parametrize("name", ORB)
def test_orb_vs_golden(name):
    g = load_golden(name)
    ex = _extractor(g)
    kps, desc = ex(g["image"])
    gk = g["kps"]
    assert len(kps) == len(gk)
    for i, f in enumerate(["x", "y", "size", "angle", "response"]):
        assert np.array_equal(kps[f], gk[:, i]), f
    assert np.array_equal(kps["octave"], g["octave"])
    assert np.all(kps["class_id"] == -1)
    assert np.array_equal(desc, g["desc"])


def test_getters_match_oracle(orc):
    from psl_slam_b200 import ORBextractor
    ex = ORBextractor()
    scale, inv, quota, _ = orc.orb_tables(orc.params())
    assert np.array_equal(ex.GetScaleFactors(), scale)
    assert np.array_equal(ex.GetInverseScaleFactors(), inv)
    assert np.array_equal(ex.features_per_level(), quota)
    assert ex.GetLevels() == 8


def test_batch_sequence_vs_oracle(orc):
    """cfg-4 shaped: a batch of consecutive frames, more frames than one chunk."""
    from psl_slam_b200 import ORBextractor, synth
    gray, _, _ = synth.sequence(4, 5)
    ex = ORBextractor(chunk_frames=2)
    kps, desc, n = ex.extract_batch(gray)
    for b in range(len(gray)):
        okps, odesc = orc.orb_extract(gray[b])
        _same(kps[b, : n[b]], desc[b, : n[b]], okps, odesc)


def test_strided_input_and_repeatability(orc):
    from psl_slam_b200 import ORBextractor, synth
    gray, _, _ = synth.sequence(6, 1)
    big = np.zeros((480, 1000), np.uint8)
    big[:, 100:740] = gray[0]
    ex = ORBextractor()
    a = ex(big[:, 100:740])
    b = ex(gray[0])
    _same(a[0], a[1], b[0], b[1])
    okps, odesc = orc.orb_extract(gray[0])
    _same(a[0], a[1], okps, odesc)


def test_random_noise_sizes(orc):
    """Dense-corner stress (every cell full, octree phase 2 everywhere) at odd sizes."""
    from psl_slam_b200 import ORBextractor
    rng = np.random.default_rng(3)
    for (w, h, nf, nl) in [(231, 187, 700, 4), (402, 150, 300, 3), (640, 480, 2000, 8)]:
        img = rng.integers(0, 256, (h, w), dtype=np.uint8)
        ex = ORBextractor(nf, 1.2, nl, 20, 7, max_width=w, max_height=h, max_candidates=131072)
        kps, desc = ex(img)
        okps, odesc = orc.orb_extract(img, orc.params(nf, 1.2, nl, 20, 7))
        _same(kps, desc, okps, odesc)


def test_empty_and_errors():
    from psl_slam_b200 import ORBextractor, PslError
    ex = ORBextractor()
    kps, desc = ex(np.zeros((0, 0), np.uint8))
    assert len(kps) == 0 and desc.shape == (0, 32)
    with pytest.raises(PslError):
        ex(np.zeros((600, 800), np.uint8))  # larger than max_width/max_height
    with pytest.raises(PslError):
        ORBextractor(1000, 1.2, 8, 20, 7, max_width=64, max_height=64)(np.zeros((64, 64), np.uint8))


def test_candidate_pool_overflow_is_loud():
    from psl_slam_b200 import ORBextractor, PslError
    rng = np.random.default_rng(5)
    img = rng.integers(0, 256, (480, 640), dtype=np.uint8)
    # 1024: the first level alone overflows; 32768: a level whose own total fits loses single cells to the other
    # levels' appends (the gather of the octree must not follow the table entries of cells that were dropped)
    for cap in (1024, 32768):
        ex = ORBextractor(max_candidates=cap)
        with pytest.raises(PslError) as e:
            ex(img)
        assert e.value.code == -3
    ex2 = ORBextractor(max_candidates=131072)
    assert len(ex2(img)[0]) >= 1000


def test_auto_candidate_pool_grows_and_reruns():
    """orb_max_candidates <= 0: the pool belongs to the context and doubles until the frame fits (the call is run again
    inside the host-pointer entry point); same key points and descriptors as with a roomy fixed pool."""
    from psl_slam_b200 import ORBextractor
    rng = np.random.default_rng(5)
    img = rng.integers(0, 256, (480, 640), dtype=np.uint8)
    k0, d0 = ORBextractor(max_candidates=131072)(img)
    k1, d1 = ORBextractor(max_candidates=-1024)(img)
    assert k0.tobytes() == k1.tobytes() and np.array_equal(d0, d1)


@pytest.mark.parametrize("name", ORB)
def test_stages_vs_golden(name):
    """Pyramid level, blur, FAST candidates and octree selection, stage by stage."""
    g = load_golden(name)
    ex = _extractor(g)
    ex(g["image"])
    lvl = int(g["level_idx"])
    li = g["level_img"]
    if lvl > 0:
        assert np.array_equal(ex.debug_fetch(0, 0, lvl).reshape(li.shape), li)
    if "level_blur" in g:
        assert np.array_equal(ex.debug_fetch(1, 0, lvl).reshape(li.shape), g["level_blur"])
    for l in range(ex.GetLevels()):
        assert np.array_equal(ex.debug_fetch(2, 0, l), g[f"cands_{l}"]), f"candidates level {l}"
        sel = ex.debug_fetch(3, 0, l)
        gs = g[f"sel_{l}"]
        assert np.array_equal(sel[:, :2] + 16, gs[:, :2]) and np.array_equal(sel[:, 2], gs[:, 2]), f"octree level {l}"


def test_unaligned_device_frames_vs_oracle(orc):
    """Device-pointer entry with a base address and row stride that are not multiples of 4 (nor 16): the byte-load
    forms of the pyramid / blur kernels and the FAST tiles without a TMA descriptor must give the same result as the
    aligned forms (and the oracle).  Low-texture frames also send most cells through the minThFAST fallback list."""
    import torch

    from psl_slam_b200 import KP_DTYPE, ORBextractor, synth
    frames = np.stack([synth.sequence(8, 1)[0][0], synth.make_lowtex(31), synth.sequence(9, 1)[0][0]])
    B, (H, W) = len(frames), frames.shape[1:]
    stride, off = W + 3, 1
    fs = stride * H + 5
    buf = torch.zeros(B * fs + 16, dtype=torch.uint8, device="cuda")
    host = np.zeros(B * fs + 16, np.uint8)
    for b in range(B):
        for y in range(H):
            s = off + b * fs + y * stride
            host[s:s + W] = frames[b, y]
    buf.copy_(torch.from_numpy(host))
    ex = ORBextractor(chunk_frames=2)
    cap = ex.cap
    kps = torch.zeros((B, cap, 28), dtype=torch.uint8, device="cuda")
    desc = torch.zeros((B, cap, 32), dtype=torch.uint8, device="cuda")
    n = torch.zeros(B, dtype=torch.int32, device="cuda")
    ex.extract_batch_dev(buf.data_ptr() + off, B, W, H, stride, fs, kps.data_ptr(), desc.data_ptr(), n.data_ptr())
    ex.ctx.sync()
    n_h, kps_h, desc_h = n.cpu().numpy(), kps.cpu().numpy(), desc.cpu().numpy()
    for b in range(B):
        okps, odesc = orc.orb_extract(frames[b])
        got = kps_h[b, : n_h[b]].copy().view(KP_DTYPE).reshape(-1)
        _same(got, desc_h[b, : n_h[b]], okps, odesc)
    assert n_h[1] < n_h[0]  # the low-texture frame is the sparse one


def test_cuda_equals_reference_build():
    """The CUDA extractor against oracle/_ref: /root/reference/src/ORBextractor.cc compiled unchanged over the OpenCV
    stand-in (tests/test_oracle_ref.py).  Prebuilt in the build container; travels with the snapshot."""
    import ctypes as C
    import os

    from conftest import ROOT, golden_names, load_golden
    from psl_slam_b200._lib import KP_DTYPE
    path = os.path.join(ROOT, "oracle", "_ref", "libpsl_ref_orb.so")
    if not os.path.exists(path):
        pytest.skip("oracle/_ref not built")
    lib = C.CDLL(path)
    lib.ref_orb_extract.restype = C.c_int
    lib.ref_orb_extract.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                    C.c_void_p, C.c_void_p, C.c_int]
    for name in golden_names("orb_"):
        g = load_golden(name)
        nf, nl, ini, mn = [int(v) for v in g["params"]]
        sf = float(g["scale_factor"])
        img = np.ascontiguousarray(g["image"])
        cap = nf + 64 * nl
        rk = np.zeros(cap, KP_DTYPE)
        rd = np.zeros((cap, 32), np.uint8)
        n = lib.ref_orb_extract(nf, sf, nl, ini, mn, img.ctypes.data, img.shape[1], img.shape[0], img.strides[0],
                                rk.ctypes.data, rd.ctypes.data, cap)
        ex = _extractor(g)
        kps, desc = ex(img)
        assert len(kps) == n, name
        assert kps.tobytes() == rk[:n].tobytes() and np.array_equal(desc, rd[:n]), name
