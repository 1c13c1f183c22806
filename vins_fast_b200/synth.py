"""Deterministic synthetic stereo sequences of EuRoC shape.

Everything here is integer arithmetic on top of numpy's PCG64 stream, so the
same (seed, width, height, frame) yields the same bytes on every host: the
golden fixtures under tests/golden/ and the GPU-box runs see identical input.

A sequence is one textured canvas viewed through a slowly varying affine
motion (left camera) and a second, slightly different affine (right camera:
about -12 px disparity, 0.3 px vertical offset, 0.2 % scale).  Motion is a
set of triangle waves so features enter and leave the image for as long as
the sequence runs (track counts diversify, ids churn) while the view stays
inside the canvas.

This module is host-side input synthesis only; it is not part of the
reference's hot path (the reference replays a ROS bag,
vins/test/test_feature_tracker.cpp:13-43).
"""
from __future__ import annotations

import numpy as np

Q = 16  # fixed-point fraction bits for warp coordinates
ONE = 1 << Q
MARGIN = 128  # canvas margin around the image, pixels


def _box_blur(a: np.ndarray, width: int, axis: int) -> np.ndarray:
    """Integer running-sum box filter (reflect padding); output keeps the sum (no divide)."""
    r = width // 2
    pad = [(0, 0), (0, 0)]
    pad[axis] = (r, r)
    p = np.pad(a, pad, mode="reflect")
    c = np.cumsum(p, axis=axis, dtype=np.int64)
    zero_shape = list(c.shape)
    zero_shape[axis] = 1
    c = np.concatenate([np.zeros(zero_shape, np.int64), c], axis=axis)
    n = a.shape[axis]
    hi = [slice(None), slice(None)]
    lo = [slice(None), slice(None)]
    hi[axis] = slice(width, width + n)
    lo[axis] = slice(0, n)
    return c[tuple(hi)] - c[tuple(lo)]


def _blur(a: np.ndarray, width: int, passes: int) -> np.ndarray:
    out = a.astype(np.int64)
    for _ in range(passes):
        out = _box_blur(_box_blur(out, width, 0), width, 1)
    return out


def _stretch(a: np.ndarray, top: int) -> np.ndarray:
    lo, hi = int(a.min()), int(a.max())
    return ((a - lo) * top) // max(hi - lo, 1)


def make_canvas(seed: int, width: int, height: int) -> np.ndarray:
    """Band-limited texture, uint8, (height+2*MARGIN) x (width+2*MARGIN)."""
    rng = np.random.default_rng(seed)
    ch, cw = height + 2 * MARGIN, width + 2 * MARGIN
    fine = rng.integers(0, 256, size=(ch, cw), dtype=np.int64)
    fine = _stretch(_blur(fine, 5, 3), 1 << 12)
    coarse = rng.integers(0, 256, size=((ch + 7) // 8, (cw + 7) // 8), dtype=np.int64)
    coarse = np.repeat(np.repeat(coarse, 8, axis=0), 8, axis=1)[:ch, :cw]
    coarse = _stretch(_blur(coarse, 9, 3), 1 << 12)
    mix = 5 * fine + 3 * coarse
    return _stretch(mix, 255).astype(np.uint8)


def _tri(k: int, period: int, amp: int) -> int:
    """Integer triangle wave in [-amp, amp], zero at k=0, period frames."""
    half = period // 2
    ph = (k + half // 2) % period
    v = ph if ph < half else period - ph  # 0..half
    return (2 * v - half) * amp // half


def left_affine(k: int) -> tuple[int, int, int, int, int, int]:
    """Q16 affine (a00,a01,b0,a10,a11,b1): canvas = A*[x,y]+b, relative to MARGIN."""
    a00 = ONE + _tri(k, 160, 1966)          # +-3 % zoom
    a11 = ONE + _tri(k + 11, 160, 1966)
    a01 = _tri(k + 40, 240, 1310)           # +-2 % shear / rotation
    a10 = -_tri(k + 40, 240, 1310)
    b0 = _tri(k, 56, 52 * ONE)              # about 3.7 px/frame
    b1 = _tri(k + 9, 72, 40 * ONE)          # about 2.2 px/frame
    return a00, a01, b0, a10, a11, b1


def right_affine(k: int) -> tuple[int, int, int, int, int, int]:
    a00, a01, b0, a10, a11, b1 = left_affine(k)
    # right(x,y) = left(x + 12 + 0.004*y, y + 0.3): features move ~12 px to the left
    a00r = a00 + a00 // 500
    a01r = a01 + (a00 * 262) // ONE         # 0.004 * a00
    b0r = b0 + 12 * a00 + (a01 * 19661) // ONE
    b1r = b1 + 12 * a10 + (a11 * 19661) // ONE   # 0.3 px
    return a00r, a01r, b0r, a10, a11, b1r


def warp(canvas: np.ndarray, width: int, height: int, aff) -> np.ndarray:
    """Fixed-point bilinear sample of the canvas through a Q16 affine."""
    a00, a01, b0, a10, a11, b1 = aff
    cx, cy = width // 2, height // 2
    x = np.arange(width, dtype=np.int64)[None, :] - cx
    y = np.arange(height, dtype=np.int64)[:, None] - cy
    sx = a00 * x + a01 * y + b0 + ((MARGIN + cx) << Q)
    sy = a10 * x + a11 * y + b1 + ((MARGIN + cy) << Q)
    ix, fx = sx >> Q, sx & (ONE - 1)
    iy, fy = sy >> Q, sy & (ONE - 1)
    ch, cw = canvas.shape
    ix = np.clip(ix, 0, cw - 2)
    iy = np.clip(iy, 0, ch - 2)
    c = canvas.astype(np.int64)
    c00, c01 = c[iy, ix], c[iy, ix + 1]
    c10, c11 = c[iy + 1, ix], c[iy + 1, ix + 1]
    top = c00 * (ONE - fx) + c01 * fx
    bot = c10 * (ONE - fx) + c11 * fx
    v = (top * (ONE - fy) + bot * fy + (1 << (2 * Q - 1))) >> (2 * Q)
    return v.astype(np.uint8)


class StereoSequence:
    """Frame k -> (t, left u8 HxW, right u8 HxW); t = k/20 s (EuRoC rate)."""

    def __init__(self, seed: int = 0, width: int = 752, height: int = 480):
        self.seed, self.width, self.height = seed, width, height
        self.canvas = make_canvas(seed, width, height)

    def frame(self, k: int):
        left = warp(self.canvas, self.width, self.height, left_affine(k))
        right = warp(self.canvas, self.width, self.height, right_affine(k))
        return k / 20.0, np.ascontiguousarray(left), np.ascontiguousarray(right)

    def frames(self, n: int, start: int = 0):
        for k in range(start, start + n):
            yield self.frame(k)
