"""TEST INFRASTRUCTURE — writes tests/golden/convert_rgbd.npz with the real cv2 4.13 cvtColor outputs
(run in the build container only; needs cv2):  python -m oracle.pyref.make_convert_golden"""
import os

import cv2
import numpy as np

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "tests", "golden")

if __name__ == "__main__":
    rng = np.random.default_rng(9)
    h, w = 37, 53
    rgb = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    rgba = rng.integers(0, 256, (h, w, 4), dtype=np.uint8)
    d = rng.integers(0, 65536, (h, w), dtype=np.uint16)
    f = np.float32(1.0) / np.float32(5000.0)  # Tracking.cc:142-145
    # imDepth.convertTo(CV_32F, factor): cv::Mat::convertTo is not exposed by cv2; its 16U -> 32F path multiplies in
    # fp32, and fp32 / fp64 scaling agree on every u16 for this factor (re-checked here)
    f32 = (d.astype(np.float32) * f).astype(np.float32)
    all16 = np.arange(65536, dtype=np.uint16)
    assert np.array_equal((all16.astype(np.float32) * f).astype(np.float32),
                          (all16.astype(np.float64) * float(f)).astype(np.float32))
    np.savez_compressed(os.path.join(OUT, "convert_rgbd.npz"), rgb=rgb, rgba=rgba, depth=d, factor=f, depth_f=f32,
                        gray_rgb=cv2.cvtColor(rgb, cv2.COLOR_RGB2GRAY), gray_bgr=cv2.cvtColor(rgb, cv2.COLOR_BGR2GRAY),
                        gray_rgba=cv2.cvtColor(rgba, cv2.COLOR_RGBA2GRAY),
                        gray_bgra=cv2.cvtColor(rgba, cv2.COLOR_BGRA2GRAY))
    print("convert_rgbd.npz written")
