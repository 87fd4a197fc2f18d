"""Host-side mirror of ORB_SLAM2::ORBextractor (include/ORBextractor.h:45-112 in the reference)
over the C-ABI.  Same constructor arguments, same getters, same call semantics:
``extractor(image, mask=None) -> (keypoints, descriptors)`` with the mask ignored
(ORBextractor.h:57-58) and an empty image returning nothing (ORBextractor.cc:1046-1047).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import KP_DTYPE, PslError


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class Context:
    """Owns one psl_ctx (one GPU, one stream, not re-entrant)."""

    def __init__(self, cfg: _lib.Config):
        self.cfg = cfg
        self._h = C.c_void_p()
        rc = _lib.lib().psl_create(C.byref(cfg), C.byref(self._h))
        if rc != 0:
            raise PslError(rc, "psl_create failed (no sm_100 GPU, or bad config) — there is no CPU fallback")

    def check(self, rc):
        if rc != 0:
            raise PslError(rc, _lib.lib().psl_last_error(self._h).decode())

    @property
    def handle(self):
        return self._h

    def sync(self):
        self.check(_lib.lib().psl_sync(self._h))

    def stream(self) -> int:
        return int(_lib.lib().psl_stream(self._h) or 0)

    STAGES = ["pyramid", "fast", "octree", "blur", "describe", "match_single", "stereo_queries", "grid",
              "candidates", "resolve", "lsd_prologue", "lsd_order", "lsd_grow", "line_merge", "lbd", "line_match"]

    def profile(self, on: bool):
        self.check(_lib.lib().psl_profile_enable(self._h, int(on)))

    def profile_read(self):
        """(ms per stage, launches per stage) accumulated since the last read."""
        ms = np.zeros(16, np.float32)
        ln = np.zeros(16, np.int64)
        self.check(_lib.lib().psl_profile_read(self._h, _ptr(ms), _ptr(ln)))
        return ms, ln

    def launch_count(self) -> int:
        return int(_lib.lib().psl_launch_count(self._h))

    def close(self):
        if self._h:
            _lib.lib().psl_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class ORBextractor:
    """ORBextractor(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST) — ORBextractor.cc:410-470."""

    def __init__(self, nfeatures=1000, scaleFactor=1.2, nlevels=8, iniThFAST=20, minThFAST=7, *, device=0,
                 max_width=640, max_height=480, chunk_frames=0, max_candidates=0, ctx: Context | None = None):
        if ctx is None:
            cfg = _lib.default_config()
            cfg.device = device
            cfg.max_width, cfg.max_height = max_width, max_height
            cfg.orb_nfeatures, cfg.orb_scale_factor, cfg.orb_nlevels = nfeatures, scaleFactor, nlevels
            cfg.orb_ini_th_fast, cfg.orb_min_th_fast = iniThFAST, minThFAST
            cfg.chunk_frames, cfg.orb_max_candidates = chunk_frames, max_candidates
            ctx = Context(cfg)
        self.ctx = ctx
        L = ctx.cfg.orb_nlevels
        self._scale = np.empty(L, np.float32)
        self._inv = np.empty(L, np.float32)
        self._s2 = np.empty(L, np.float32)
        self._is2 = np.empty(L, np.float32)
        self._quota = np.empty(L, np.int32)
        nl = C.c_int32()
        ctx.check(_lib.lib().psl_orb_tables(ctx.handle, C.addressof(nl), _ptr(self._scale), _ptr(self._inv),
                                            _ptr(self._s2), _ptr(self._is2), _ptr(self._quota)))
        self.cap = int(ctx.cfg.orb_nfeatures + 4 * L + 64)

    # ORBextractor.h:63-83
    def GetLevels(self): return int(self.ctx.cfg.orb_nlevels)
    def GetScaleFactor(self): return float(self.ctx.cfg.orb_scale_factor)
    def GetScaleFactors(self): return self._scale.copy()
    def GetInverseScaleFactors(self): return self._inv.copy()
    def GetScaleSigmaSquares(self): return self._s2.copy()
    def GetInverseScaleSigmaSquares(self): return self._is2.copy()
    def features_per_level(self): return self._quota.copy()

    def __call__(self, image: np.ndarray, mask=None):
        """operator()(image, mask, keypoints, descriptors) — ORBextractor.cc:1043-1105."""
        image = np.asarray(image)
        if image.size == 0:
            return np.zeros(0, KP_DTYPE), np.zeros((0, 32), np.uint8)
        kps, desc, n = self.extract_batch(image[None])
        return kps[0, : n[0]].copy(), desc[0, : n[0]].copy()

    def extract_batch(self, frames: np.ndarray):
        """frames: host u8 [B,H,W] (any row stride).  Returns (kps [B,cap], desc [B,cap,32], n [B])."""
        if frames.dtype != np.uint8 or frames.ndim != 3:
            raise ValueError("expected CV_8UC1 frames [B,H,W]")  # assert at ORBextractor.cc:1050
        if frames.strides[2] != 1:
            frames = np.ascontiguousarray(frames)
        B, H, W = frames.shape
        kps = np.zeros((B, self.cap), KP_DTYPE)
        desc = np.zeros((B, self.cap, 32), np.uint8)
        n = np.zeros(B, np.int32)
        self.ctx.check(_lib.lib().psl_orb_extract_batch(self.ctx.handle, _ptr(frames), B, W, H, frames.strides[1],
                                                        frames.strides[0], _ptr(kps), _ptr(desc), self.cap, _ptr(n)))
        return kps, desc, n

    def extract_batch_dev(self, d_gray_ptr: int, B, W, H, stride, frame_stride, d_kps_ptr: int, d_desc_ptr: int,
                          d_n_ptr: int, cap: int | None = None):
        """Device-pointer form (asynchronous on the ctx stream); call ctx.sync() to collect errors."""
        self.ctx.check(_lib.lib().psl_orb_extract_batch_dev(self.ctx.handle, d_gray_ptr, B, W, H, stride,
                                                            frame_stride, d_kps_ptr, d_desc_ptr,
                                                            cap or self.cap, d_n_ptr))

    def debug_fetch(self, what: int, frame: int, level: int, nbytes: int = 1 << 24) -> np.ndarray:
        """Intermediate of the last call (psl_debug_fetch): 0 level image, 1 blur, 2 candidates, 3 selected."""
        buf = np.empty(nbytes, np.uint8)
        n = C.c_int64()
        self.ctx.check(_lib.lib().psl_debug_fetch(self.ctx.handle, what, frame, level, _ptr(buf), nbytes,
                                                  C.addressof(n)))
        if what in (0, 1):
            return buf[: n.value].copy()
        k = buf[: n.value * 4].view(np.uint32)
        return np.stack([k >> 20, (k >> 8) & 0xFFF, k & 0xFF], 1).astype(np.float32)
