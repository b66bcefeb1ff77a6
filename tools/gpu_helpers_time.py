"""Kernel durations of the pyramid helpers on a 3840x2160 frame (run under ncu --metrics gpu__time_duration.sum)."""
import sys
sys.path.insert(0, ".")
import numpy as np
from pislam_b200 import api
rng = np.random.Generator(np.random.PCG64(1))
img = rng.integers(0, 256, size=(2160, 3840), dtype=np.uint8)
with api.Context(0) as ctx:
    for _ in range(3):
        out = np.zeros_like(img)
        ctx.gaussian5x5(3840, 2160, img, out)
        ctx.bilinear7_8(3840, 2160, img, out)
        ctx.bilinear13_16(3840, 2160, img, out)
print("done")
