"""TEST INFRASTRUCTURE — writes tests/golden/*.npz (run in the build container only; needs cv2).

    python -m oracle.pyref.make_goldens [orb|match|line|all]

Each fixture stores the seeded input bytes and the outputs of the cv2-primitive
restatement of the reference (oracle/pyref), so that the tests never need cv2 or
/root/reference at run time.
"""
from __future__ import annotations

import os
import sys

import numpy as np

from oracle.pyref import orb_cv2
from psl_slam_b200 import synth

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "tests", "golden")


def orb_cases():
    g, _, _ = synth.sequence(1, 1)
    yield "orb_vga_seed1", g[0], dict(nfeatures=1000, scale_factor=1.2, nlevels=8, ini_th=20, min_th=7)
    rng = np.random.default_rng(11)
    yield "orb_noise_200x160", rng.integers(0, 256, (160, 200), dtype=np.uint8), dict(
        nfeatures=500, scale_factor=1.2, nlevels=4, ini_th=20, min_th=7)
    poster = synth.make_poster(12, 1024, 900)
    rgb, _ = synth.render(poster, synth.trajectory(1, 12)[0], 752, 300, noise_seed=5)
    yield "orb_wide_752x300", synth.rgb_to_gray(rgb), dict(nfeatures=800, scale_factor=1.2, nlevels=6, ini_th=20,
                                                          min_th=7)
    yield "orb_lowtex_320x240", synth.make_lowtex(3, 320, 240), dict(nfeatures=300, scale_factor=1.2, nlevels=5,
                                                                     ini_th=20, min_th=7)
    poster = synth.make_poster(5, 2048)
    rgb, _ = synth.render(poster, synth.trajectory(1, 5)[0], 960, 540, noise_seed=9)
    yield "orb_qhd_960x540_12lvl", synth.rgb_to_gray(rgb), dict(nfeatures=4000, scale_factor=1.2, nlevels=12,
                                                               ini_th=20, min_th=7)
    yield "orb_flat_160x120", np.full((120, 160), 77, np.uint8), dict(nfeatures=100, scale_factor=1.2, nlevels=3,
                                                                     ini_th=20, min_th=7)


def make_orb():
    for name, img, kw in orb_cases():
        P = orb_cv2.OrbParams(**kw)
        st = {}
        kps, octv, desc = orb_cv2.orb_extract(img, P, st)
        arrays = dict(image=img, kps=kps, octave=octv, desc=desc, params=np.array(
            [kw["nfeatures"], kw["nlevels"], kw["ini_th"], kw["min_th"]], np.int32),
            scale_factor=np.float32(kw["scale_factor"]))
        for l in range(P.nlevels):
            arrays[f"cands_{l}"] = st["cands"][l]
            arrays[f"sel_{l}"] = st["selected"][l]
        # one mid-pyramid level image + its blur pin the resize chain and the blur
        l = min(2, P.nlevels - 1)
        arrays["level_idx"] = np.int32(l)
        arrays["level_img"] = st["level_img"][l]
        if l in st.get("blur", {}):
            arrays["level_blur"] = st["blur"][l]
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **arrays)
        print(name, img.shape, "n =", len(kps), "per-level", np.bincount(octv, minlength=P.nlevels).tolist())


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    os.makedirs(OUT, exist_ok=True)
    if what in ("orb", "all"):
        make_orb()
