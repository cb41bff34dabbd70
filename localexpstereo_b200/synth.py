"""Synthetic inputs of SURVEY.md section 8(d): guide images, cost volumes, plane hypotheses.
numpy only (input generation, no hot-path arithmetic)."""
from __future__ import annotations

import numpy as np

f32 = np.float32


def synthetic_image(H, W, seed):
    """8UC3 guide: blurred (sigma ~ 3) uniform noise stretched to 0..255 plus +-8 white noise."""
    rng = np.random.default_rng(seed)
    img = np.empty((H, W, 3), dtype=np.uint8)
    k = np.exp(-0.5 * (np.arange(-9, 10) / 3.0) ** 2)
    k /= k.sum()
    for c in range(3):
        z = rng.random((H + 18, W + 18))
        zz = np.zeros((H + 18, W))
        for i, kk in enumerate(k):
            zz += kk * z[:, i:i + W]
        z = np.zeros((H, W))
        for i, kk in enumerate(k):
            z += kk * zz[i:i + H, :]
        z = (z - z.min()) / max(z.max() - z.min(), 1e-12) * 255.0
        z = z + rng.integers(-8, 9, size=(H, W))
        img[:, :, c] = np.clip(np.rint(z), 0, 255).astype(np.uint8)
    return img


def _create_plane(n, z, x, y):
    """Plane::CreatePlane (Plane.h:14-22) vectorised, float arithmetic."""
    n = n.astype(f32)
    a = (-n[:, 0] / n[:, 2]).astype(f32)
    b = (-n[:, 1] / n[:, 2]).astype(f32)
    c = ((z.astype(f32) - a * x.astype(f32)).astype(f32) - (b * y.astype(f32)).astype(f32)).astype(f32)
    return np.stack([a, b, c, np.zeros_like(a)], axis=1).astype(f32)


def _unit_vectors(rng, n, theta_range):
    """cvutils::getRandomUnitVector (Utilities.hpp:254-261)."""
    th = rng.uniform(0.0, theta_range, n)
    ph = rng.uniform(0.0, 2 * np.pi, n)
    return np.stack([np.sin(th) * np.cos(ph), np.sin(th) * np.sin(ph), np.cos(th)], axis=1)


def synthetic_planes(unit_regions, n_steps, D, seed):
    """[n_steps][n_cells][4] float32.  Step 0: StereoEnergy::createRandomLabel (StereoEnergy.h:120-129)
    anchored at a uniform random pixel of the unit cell; step s >= 1: RandomProposer::getNextProposal
    (Proposer.h:120-148) of the step-0 label with m = s - 1."""
    rng = np.random.default_rng(seed)
    u = np.asarray(unit_regions, dtype=np.int64)
    n = len(u)
    sx = (u[:, 0] + rng.integers(0, 1 << 30, n) % u[:, 2]).astype(f32)
    sy = (u[:, 1] + rng.integers(0, 1 << 30, n) % u[:, 3]).astype(f32)
    MAX = f32(D - 1)
    out = np.zeros((n_steps, n, 4), f32)
    z0 = rng.uniform(0.0, float(MAX), n).astype(f32)
    base = _create_plane(_unit_vectors(rng, n, np.pi / 3), z0, sx, sy)
    out[0] = base
    a, b, c = base[:, 0], base[:, 1], base[:, 2]
    nz = (1.0 / np.sqrt(1.0 + (a * a).astype(np.float64) + (b * b).astype(np.float64))).astype(f32)  # Plane.h:42-50
    normal = np.stack([-a * nz, -b * nz, nz], axis=1).astype(f32)
    zs = ((a * sx + b * sy).astype(f32) + c).astype(f32)
    for s in range(1, n_steps):
        m = s - 1
        dz = f32(MAX * f32(0.5) ** (m + 1))
        lo = np.maximum(f32(0), zs - dz)
        hi = np.minimum(MAX, zs + dz)
        z = (lo + rng.random(n).astype(f32) * (hi - lo)).astype(f32)
        nv = normal + (_unit_vectors(rng, n, np.pi) * (0.5 ** m)).astype(f32)
        nv = nv / np.sqrt((nv.astype(np.float64) ** 2).sum(axis=1, keepdims=True))
        out[s] = _create_plane(nv.astype(f32), z, sx, sy)
    return out
