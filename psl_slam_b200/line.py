"""Host-side mirror of ORB_SLAM2::LINEextractor (add_inc/LineExtractor.h:159-253 in the reference) over the
C-ABI.  Same constructor arguments (LineExtractor.cpp:6-25), same getters, same call semantics:
``extractor(image, mask=None) -> (keylines, descriptors, lineVec2d)`` — LSD, long-line merge, top-N by
response, LBD descriptors and normalised 2-D line equations (LineExtractor.cpp:325-366); an empty image
returns nothing (:327-328).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import KEYLINE_DTYPE
from .orb import Context, _ptr


class LINEextractor:
    """LINEextractor(numOctaves, scale, nLSDFeature, min_line_length) — LineExtractor.cpp:6-25."""

    def __init__(self, numOctaves=1, scale=1.2, nLSDFeature=200, min_line_length=0.0, *, device=0, max_width=640,
                 max_height=480, chunk_frames=0, max_raw=0, ctx: Context | None = None):
        if ctx is None:
            cfg = _lib.default_config()
            cfg.device = device
            cfg.max_width, cfg.max_height = max_width, max_height
            cfg.line_nlevels, cfg.line_scale_factor = numOctaves, scale
            cfg.line_nfeatures, cfg.line_min_length = nLSDFeature, min_line_length
            cfg.line_chunk_frames, cfg.line_max_raw = chunk_frames, max_raw
            ctx = Context(cfg)
        self.ctx = ctx
        L = int(ctx.cfg.line_nlevels)
        s = float(ctx.cfg.line_scale_factor)
        # LineExtractor.cpp:9-24 (float tables built by repeated multiplication)
        self._scale = np.ones(L, np.float32)
        for i in range(1, L):
            self._scale[i] = np.float32(self._scale[i - 1] * np.float32(s))
        self._s2 = self._scale * self._scale
        self._inv = (np.float32(1.0) / self._scale).astype(np.float32)
        self._is2 = (np.float32(1.0) / self._s2).astype(np.float32)
        self.cap = int(ctx.cfg.line_nfeatures)

    # LineExtractor.h:211-233
    def GetLevels(self): return int(self.ctx.cfg.line_nlevels)
    def GetScaleFactor(self): return float(self.ctx.cfg.line_scale_factor)
    def GetScaleFactors(self): return self._scale.copy()
    def GetInverseScaleFactors(self): return self._inv.copy()
    def GetScaleSigmaSquares(self): return self._s2.copy()
    def GetInverseScaleSigmaSquares(self): return self._is2.copy()

    def __call__(self, image: np.ndarray, mask=None, with_lbd_floats: bool = False):
        """operator()(image, mask, keylines, descriptors, lineVec2d)."""
        image = np.asarray(image)
        empty = (np.zeros(0, KEYLINE_DTYPE), np.zeros((0, 32), np.uint8), np.zeros((0, 3), np.float64))
        if image.size == 0:
            return empty + ((np.zeros((0, 72), np.float32),) if with_lbd_floats else ())
        out = self.extract_batch(image[None], with_lbd_floats)
        n = int(out[-1][0])
        res = tuple(a[0, :n].copy() for a in out[:-1])
        return res

    def extract_batch(self, frames: np.ndarray, with_lbd_floats: bool = False):
        """frames: host u8 [B,H,W].  Returns (kl [B,cap], ldesc [B,cap,32], lineeq [B,cap,3](, lbd [B,cap,72]), n [B])."""
        if frames.dtype != np.uint8 or frames.ndim != 3:
            raise ValueError("expected CV_8UC1 frames [B,H,W]")  # assert at LineExtractor.cpp:331
        if frames.strides[2] != 1:
            frames = np.ascontiguousarray(frames)
        B, H, W = frames.shape
        kl = np.zeros((B, self.cap), KEYLINE_DTYPE)
        ld = np.zeros((B, self.cap, 32), np.uint8)
        eq = np.zeros((B, self.cap, 3), np.float64)
        lbd = np.zeros((B, self.cap, 72), np.float32) if with_lbd_floats else None
        n = np.zeros(B, np.int32)
        self.ctx.check(_lib.lib().psl_line_extract_batch(self.ctx.handle, _ptr(frames), B, W, H, frames.strides[1],
                                                         frames.strides[0], _ptr(kl), _ptr(ld), _ptr(eq),
                                                         _ptr(lbd) if with_lbd_floats else None, self.cap, _ptr(n)))
        return (kl, ld, eq, lbd, n) if with_lbd_floats else (kl, ld, eq, n)

    def extract_batch_dev(self, d_gray: int, B, W, H, stride, frame_stride, d_kl: int, d_ldesc: int, d_lineeq: int,
                          d_lbd72: int | None, d_n: int, cap: int | None = None):
        """Device-pointer form (asynchronous on the ctx stream); call ctx.sync() to collect errors."""
        self.ctx.check(_lib.lib().psl_line_extract_batch_dev(self.ctx.handle, d_gray, B, W, H, stride, frame_stride,
                                                             d_kl, d_ldesc, d_lineeq, d_lbd72, cap or self.cap, d_n))

    def debug_fetch(self, what: int, frame: int = 0, nbytes: int = 1 << 22) -> np.ndarray:
        """Intermediates of the last call: 4 = the 0.8x image LSD works on (flat), 5 = raw LSD segments [n,4]
        (after checkLineExtremes, which the merge stage applies in place)."""
        buf = np.empty(nbytes, np.uint8)
        n = C.c_int64()
        self.ctx.check(_lib.lib().psl_debug_fetch(self.ctx.handle, what, frame, 0, _ptr(buf), nbytes, C.addressof(n)))
        if what == 4:
            return buf[: n.value].copy()
        return buf[: n.value * 16].view(np.float32).reshape(-1, 4).copy()
