"""Deterministic synthetic "CDnet-shaped" sequences: static textured background with a
flickering dynamic region, moving rectangles/discs as foreground, per-pixel uniform noise.
Frame 0 is background only.  Used by tests/ and bench.py; no dataset access is needed (there is no network).
"""
import numpy as np


class SynthSequence:
    def __init__(self, width, height, channels=3, seed=1, n_objects=5, noise=3, fg_area=None):
        """fg_area: None keeps the historical object sizes (half-size 3-7 % of min(W,H): ~2 % of a 16:9 frame is foreground);
        a fraction (SURVEY 8(d) specifies 5-10 %) sizes the objects so that together they cover about that share of the frame."""
        self.w, self.h, self.c, self.seed, self.noise = width, height, channels, seed, noise
        rng = np.random.RandomState(seed)
        yy, xx = np.mgrid[0:height, 0:width].astype(np.float32)
        bg = np.zeros((height, width, channels), np.float32)
        for c in range(channels):
            acc = 110.0 + 40.0 * (xx / max(width - 1, 1)) + 25.0 * (yy / max(height - 1, 1)) * (1 if c != 1 else -1)
            for _ in range(8):
                fx, fy = rng.uniform(0.01, 0.35, 2)
                ph = rng.uniform(0, 2 * np.pi)
                acc = acc + (40.0 / 8) * np.sin(fx * xx + fy * yy + ph)
            bg[..., c] = acc
        self.bg = bg
        # dynamic region: ~10 % of the area, +-12 flicker
        dw, dh = max(width // 3, 1), max(int(height * 0.3), 1)
        self.dyn = (slice(height // 10, height // 10 + dh), slice(width // 2, width // 2 + dw))
        self.objs = []
        for _ in range(n_objects):
            kind = rng.randint(0, 2)
            size = rng.uniform(0.03, 0.07) * min(width, height)
            if fg_area is not None:
                # rectangle of half-size s covers (2s+1)^2, disc of radius s covers pi s^2; shares of 0.6-1.4 x the mean per object
                a = fg_area * width * height / n_objects * (size / (0.05 * min(width, height)))
                size = (np.sqrt(a) - 1) / 2 if kind == 0 else np.sqrt(a / np.pi)
            pos = rng.uniform([0, 0], [width, height])
            vel = rng.uniform(1.0, 3.0, 2) * rng.choice([-1, 1], 2)
            col = rng.randint(0, 256, channels).astype(np.float32)
            self.objs.append((kind, size, pos, vel, col))
        self._yy, self._xx = yy, xx

    def frame(self, t, with_gt=False):
        rng = np.random.RandomState((self.seed * 1000003 + t) & 0x7FFFFFFF)
        f = self.bg.copy()
        f[self.dyn] += 12.0 * np.sin(0.9 * t + 0.15 * self._xx[self.dyn])[..., None]
        gt = np.zeros((self.h, self.w), bool)
        if t > 0:
            for kind, size, pos, vel, col in self.objs:
                cx = (pos[0] + vel[0] * t) % self.w
                cy = (pos[1] + vel[1] * t) % self.h
                x0, x1 = int(max(cx - size, 0)), int(min(cx + size + 1, self.w))
                y0, y1 = int(max(cy - size, 0)), int(min(cy + size + 1, self.h))
                if x1 <= x0 or y1 <= y0:
                    continue
                if kind == 0:
                    m = np.ones((y1 - y0, x1 - x0), bool)
                else:
                    m = (self._xx[y0:y1, x0:x1] - cx) ** 2 + (self._yy[y0:y1, x0:x1] - cy) ** 2 <= size * size
                f[y0:y1, x0:x1][m] = col
                gt[y0:y1, x0:x1] |= m
        f += rng.randint(-self.noise, self.noise + 1, f.shape).astype(np.float32)
        out = np.clip(np.rint(f), 0, 255).astype(np.uint8)
        if self.c == 1:
            out = out[..., 0]
        return (out, gt) if with_gt else out

    def frames(self, n, start=0):
        return np.stack([self.frame(t) for t in range(start, start + n)])
