"""Stacked-pyramid geometry and synthetic pyramids.

The reference consumes a pre-blurred, pre-built, vertically stacked pyramid and leaves its
construction to the caller (reference README.md:28-31, 56-83; level table at demo/demo.cpp:38-47).
This module only describes that layout and fabricates deterministic inputs of that shape for
tests and bench.py -- it is host-side numpy, not part of the device path.
"""
from dataclasses import dataclass
from typing import List

import numpy as np


@dataclass(frozen=True)
class PyramidLayout:
    """Levels stacked top to bottom in one `rows x vstep` uint8 buffer (row pitch = base width)."""
    vstep: int
    widths: List[int]
    heights: List[int]
    row_offsets: List[int]

    @property
    def nlevels(self) -> int:
        return len(self.widths)

    @property
    def rows(self) -> int:
        return self.row_offsets[-1] + self.heights[-1]

    @property
    def stack_bytes(self) -> int:
        return self.rows * self.vstep

    @property
    def level_pixels(self) -> int:
        """Sum of w*h over levels: the unit Mpix/s is quoted on."""
        return int(sum(w * h for w, h in zip(self.widths, self.heights)))

    def classified_pixels(self, border: int) -> int:
        return int(sum(max(0, w - 2 * border) * max(0, h - 2 * border)
                       for w, h in zip(self.widths, self.heights)))


def pyramid_layout(width: int, height: int, nlevels: int = 8, scale: float = 1.2) -> PyramidLayout:
    """round(W/scale^i) x round(H/scale^i): reproduces demo/demo.cpp:38-47 for 640x480."""
    ws, hs, offs, r = [], [], [], 0
    for i in range(nlevels):
        w = int(np.floor(width / scale ** i + 0.5))
        h = int(np.floor(height / scale ** i + 0.5))
        ws.append(w)
        hs.append(h)
        offs.append(r)
        r += h
    return PyramidLayout(vstep=width, widths=ws, heights=hs, row_offsets=offs)


VGA = pyramid_layout(640, 480)
FHD = pyramid_layout(1920, 1080)
UHD = pyramid_layout(3840, 2160)


def default_nshapes(width: int, height: int) -> int:
    """Tuned once against the oracle: ~1000 keypoints at VGA, ~2000 at 1080p, ~4000 at 4K."""
    return max(4, int(NSHAPES_VGA * (width * height / (640.0 * 480.0)) ** (1.0 / 3.0)))


NSHAPES_VGA = 90


def _blur5(img: np.ndarray) -> np.ndarray:
    """Separable [1 4 6 4 1]/16 with reflected borders, in uint16 arithmetic."""
    k = (1, 4, 6, 4, 1)
    a = img.astype(np.uint16)
    p = np.pad(a, ((0, 0), (2, 2)), mode="reflect")
    a = sum(k[i] * p[:, i:i + a.shape[1]] for i in range(5))
    a = (a + 8) >> 4
    p = np.pad(a, ((2, 2), (0, 0)), mode="reflect")
    a = sum(k[i] * p[i:i + a.shape[0], :] for i in range(5))
    return ((a + 8) >> 4).astype(np.uint8)


def _resize_bilinear(img: np.ndarray, w: int, h: int) -> np.ndarray:
    """Pixel-centre-aligned bilinear resize, integer weights /256 (deterministic on any host)."""
    H, W = img.shape
    if (w, h) == (W, H):
        return img.copy()

    def taps(n_out, n_in):
        pos = (np.arange(n_out, dtype=np.int64) * 2 + 1) * n_in * 128 // n_out - 128  # 8.8 fixed point
        pos = np.clip(pos, 0, (n_in - 1) * 256)
        i0 = pos >> 8
        f = (pos & 255).astype(np.int32)
        i1 = np.minimum(i0 + 1, n_in - 1)
        return i0, i1, f

    x0, x1, fx = taps(w, W)
    y0, y1, fy = taps(h, H)
    a = img.astype(np.int32)
    top = a[y0][:, x0] * (256 - fx) + a[y0][:, x1] * fx
    bot = a[y1][:, x0] * (256 - fx) + a[y1][:, x1] * fx
    out = (top * (256 - fy)[:, None] + bot * fy[:, None] + (1 << 15)) >> 16
    return out.astype(np.uint8)


def synthetic_base(width: int, height: int, seed: int, nshapes: int = 0) -> np.ndarray:
    """A blurred scene of random rectangles / triangles / discs on mid-gray plus mild noise.

    `nshapes` (default: default_nshapes()) sets the corner density so that the VGA / 1080p /
    4K pyramids yield on the order of 1000 / 2000 / 4000 keypoints with the demo's parameters
    (FAST 20, Harris 2^15, buckets <4,3>); the oracle's actual count is what tests and bench report.
    """
    rng = np.random.Generator(np.random.PCG64(seed))
    img = np.full((height, width), 128, np.int32)
    yy, xx = np.mgrid[0:height, 0:width]
    n = nshapes if nshapes > 0 else default_nshapes(width, height)
    for _ in range(n):
        kind = rng.integers(0, 3)
        cx, cy = rng.integers(0, width), rng.integers(0, height)
        s = int(rng.integers(5, 5 + max(8, min(width, height) // 14)))
        g = int(rng.integers(0, 256))
        x0, x1 = max(0, cx - s), min(width, cx + s + 1)
        y0, y1 = max(0, cy - s), min(height, cy + s + 1)
        if x1 <= x0 or y1 <= y0:
            continue
        sx, sy = xx[y0:y1, x0:x1] - cx, yy[y0:y1, x0:x1] - cy
        if kind == 0:  # rotated rectangle
            th = rng.uniform(0, np.pi)
            c, sn = np.cos(th), np.sin(th)
            u, v = c * sx + sn * sy, -sn * sx + c * sy
            m = (np.abs(u) <= s * 0.7) & (np.abs(v) <= s * rng.uniform(0.2, 0.7))
        elif kind == 1:  # triangle (half-plane cut of a square)
            th = rng.uniform(0, 2 * np.pi)
            m = (np.abs(sx) <= s * 0.7) & (np.abs(sy) <= s * 0.7) & (np.cos(th) * sx + np.sin(th) * sy <= 0)
        else:  # disc
            m = sx * sx + sy * sy <= (s * 0.6) ** 2
        img[y0:y1, x0:x1][m] = g
    img += rng.integers(-4, 5, size=img.shape)
    return _blur5(np.clip(img, 0, 255).astype(np.uint8))


def build_stack(base: np.ndarray, layout: PyramidLayout) -> np.ndarray:
    """Resize `base` to every level and stack them at pitch `vstep`, zero padding on the right
    of each level (the shape of the reference's fixture demo/input.png)."""
    stack = np.zeros((layout.rows, layout.vstep), np.uint8)
    for w, h, r in zip(layout.widths, layout.heights, layout.row_offsets):
        stack[r:r + h, :w] = _resize_bilinear(base, w, h)
    return stack


def synthetic_stack(layout: PyramidLayout, frame_index: int = 0) -> np.ndarray:
    base = synthetic_base(layout.widths[0], layout.heights[0], 0x5EED0000 + frame_index)
    return build_stack(base, layout)


def noise_stack(layout: PyramidLayout, seed: int = 5489) -> np.ndarray:
    """Adversarial uniform noise (as reference test/TestUtil.cpp:54-62 does for its own tests);
    keeps the zero padding right of each level."""
    rng = np.random.Generator(np.random.PCG64(seed))
    stack = np.zeros((layout.rows, layout.vstep), np.uint8)
    for w, h, r in zip(layout.widths, layout.heights, layout.row_offsets):
        stack[r:r + h, :w] = rng.integers(0, 256, size=(h, w), dtype=np.uint8)
    return stack
