"""Seeded synthetic point-cloud maps for the TRG hot path.

The reference's prebuilt maps (config/*.yaml:7, shellscripts/download_maps.sh:7-9) are
downloaded from a server and are not available offline, so every test / bench input is
generated here. All generators are deterministic in (shape, seed): numpy PCG64 streams,
float32 output, points emitted in a seeded shuffle (the reference's insertion-order kd-tree
degenerates on raster-ordered input), and every point gets an independent
uniform (x, y) jitter of +-0.2*h so no two points share (x, y).
"""
from __future__ import annotations

import numpy as np


def _value_noise(x: np.ndarray, y: np.ndarray, wavelength: float, lat: np.ndarray) -> np.ndarray:
    """Smoothstep-interpolated lattice noise in [-1, 1] on the given random lattice."""
    n = lat.shape[0]
    fx = x / wavelength
    fy = y / wavelength
    ix = np.floor(fx).astype(np.int64)
    iy = np.floor(fy).astype(np.int64)
    tx = fx - ix
    ty = fy - iy
    tx = tx * tx * (3.0 - 2.0 * tx)
    ty = ty * ty * (3.0 - 2.0 * ty)
    ix = np.clip(ix, 0, n - 2)
    iy = np.clip(iy, 0, n - 2)
    v00 = lat[ix, iy]
    v10 = lat[ix + 1, iy]
    v01 = lat[ix, iy + 1]
    v11 = lat[ix + 1, iy + 1]
    return (v00 * (1 - tx) + v10 * tx) * (1 - ty) + (v01 * (1 - tx) + v11 * tx) * ty


def _lattice(nx: int, ny: int, h: float, rng: np.random.Generator):
    gx, gy = np.meshgrid(np.arange(nx, dtype=np.float64) * h, np.arange(ny, dtype=np.float64) * h,
                         indexing="ij")
    x = gx.ravel() + rng.uniform(-0.2 * h, 0.2 * h, size=nx * ny)
    y = gy.ravel() + rng.uniform(-0.2 * h, 0.2 * h, size=nx * ny)
    return x, y


def _finish(x, y, z, rng: np.random.Generator, shuffle: bool) -> np.ndarray:
    pts = np.stack([x, y, z], axis=1).astype(np.float32)
    if shuffle:
        pts = pts[rng.permutation(pts.shape[0])]
    return np.ascontiguousarray(pts)


def mountain_slab(side: int, h: float = 0.1, seed: int = 2, *, slab: int = 0, n_slabs: int = 1, **kw) -> np.ndarray:
    """ONE side x side mountain map cut into `n_slabs` slabs along x (strong scaling, BASELINE config #3):
    slab k holds the lattice columns [k * side // n_slabs, (k + 1) * side // n_slabs) of the same
    continuous height function (side must be divisible by n_slabs, or the remainder columns are dropped)."""
    nx = side // n_slabs
    return mountain(nx, side, h=h, seed=seed, tile=(slab, 0), world_tiles=(n_slabs, 1), **kw)


def mountain(nx: int, ny: int | None = None, h: float = 0.1, seed: int = 2, *,
             base_wavelength: float = 64.0, amplitude: float = 12.0, octaves: int = 5,
             gain: float = 0.5, noise_sigma: float = 0.01, shuffle: bool = True,
             tile=(0, 0), world_tiles=(1, 1)) -> np.ndarray:
    """C2/C3/C4 generator: fBm value-noise heightfield on a jittered h-lattice.

    `amplitude` is the half-range of the fBm sum: the octave weights are scaled so that the
    summed field lies in [-amplitude, +amplitude] (12 m => ~31 % of TRG edges carry a
    non-zero risk with config/mountain.yaml parameters, ~5.5 nodes/m^2).

    Multi-GPU sharding: the world is `world_tiles` = (tx, ty) tiles of nx x ny lattice points; the
    height is ONE continuous function of world coordinates (the noise lattices depend only on
    `seed` and the world extent), `tile` = (i, j) selects which tile's points are emitted. With the
    default single tile the output is identical to an unsharded map.
    """
    ny = nx if ny is None else ny
    field_rng = np.random.default_rng(seed)
    wx, wy = world_tiles
    extent = max(nx * wx, ny * wy) * h + 1.0
    w = np.array([gain ** i for i in range(octaves)])
    w = w / w.sum() * amplitude
    lattices = []
    for i in range(octaves):
        wl = base_wavelength / (2 ** i)
        n = int(np.ceil(extent / wl)) + 3
        lattices.append(field_rng.uniform(-1.0, 1.0, size=(n, n)))
    rng = np.random.default_rng([seed, 7919 + tile[0], 104729 + tile[1]])
    x, y = _lattice(nx, ny, h, rng)
    x = x + tile[0] * nx * h
    y = y + tile[1] * ny * h
    z = np.zeros_like(x)
    for i in range(octaves):
        z += w[i] * _value_noise(x + 0.5 * h, y + 0.5 * h, base_wavelength / (2 ** i), lattices[i])
    z += rng.normal(0.0, noise_sigma, size=z.shape)
    return _finish(x, y, z, rng, shuffle)


def indoor(nx: int, ny: int | None = None, h: float = 0.2, seed: int = 1, *, room: float = 10.0,
           wall_thickness: float = 0.2, door: float = 1.2, wall_height: float = 2.0,
           floor_sigma: float = 0.005, shuffle: bool = True) -> np.ndarray:
    """C1 generator: flat floor + axis-aligned wall grid with door gaps.

    Walls run along x = k*room and y = k*room (k >= 1), `wall_thickness` thick, sampled as
    z-columns 0..wall_height at h steps; each wall segment between two crossings has a
    `door`-wide gap at its middle.
    """
    ny = nx if ny is None else ny
    rng = np.random.default_rng(seed)
    x, y = _lattice(nx, ny, h, rng)
    z = rng.normal(0.0, floor_sigma, size=x.shape)
    lx, ly = nx * h, ny * h

    def on_wall(u, v):
        # wall lines at u = k*room ; door gap where (v mod room) is within door/2 of room/2
        k = np.round(u / room)
        near = (np.abs(u - k * room) <= wall_thickness * 0.5) & (k >= 1) & (k * room < max(lx, ly) - 1e-6)
        vm = np.mod(v, room)
        gap = np.abs(vm - room * 0.5) <= door * 0.5
        return near & ~gap

    wall = on_wall(x, y) | on_wall(y, x)
    wx, wy = x[wall], y[wall]
    levels = np.arange(h, wall_height + 1e-6, h)
    cols_x = np.repeat(wx, len(levels)) + rng.uniform(-0.2 * h, 0.2 * h, size=wx.size * len(levels))
    cols_y = np.repeat(wy, len(levels)) + rng.uniform(-0.2 * h, 0.2 * h, size=wx.size * len(levels))
    cols_z = np.tile(levels, wx.size) + rng.normal(0.0, floor_sigma, size=wx.size * len(levels))
    x = np.concatenate([x, cols_x])
    y = np.concatenate([y, cols_y])
    z = np.concatenate([z, cols_z])
    return _finish(x, y, z, rng, shuffle)


def stairs(nx: int, ny: int | None = None, h: float = 0.1, seed: int = 5, *, riser: float = 0.18,
           tread: float = 0.3, slab_fraction: float = 0.10, slab_height: float = 2.0,
           shuffle: bool = True) -> np.ndarray:
    """C5 generator: terraces/stairs along x plus overhang slabs above ~10 % of the area."""
    ny = nx if ny is None else ny
    rng = np.random.default_rng(seed)
    x, y = _lattice(nx, ny, h, rng)
    period = 40.0  # stairs go up for 20 m then down for 20 m
    u = np.mod(x, period)
    up = np.where(u < period / 2, u, period - u)
    z = np.floor(up / tread) * riser + rng.normal(0.0, 0.005, size=x.shape)
    # overhang slabs: 4 m x 4 m tiles chosen with probability slab_fraction
    tx = np.floor(x / 4.0).astype(np.int64)
    ty = np.floor(y / 4.0).astype(np.int64)
    tile_rng = np.random.default_rng(seed + 1000)
    ntx, nty = tx.max() + 1, ty.max() + 1
    sel = tile_rng.uniform(size=(ntx, nty)) < slab_fraction
    slab = sel[tx, ty]
    sx = x[slab] + rng.uniform(-0.2 * h, 0.2 * h, size=int(slab.sum()))
    sy = y[slab] + rng.uniform(-0.2 * h, 0.2 * h, size=int(slab.sum()))
    sz = z[slab] + slab_height
    x = np.concatenate([x, sx])
    y = np.concatenate([y, sy])
    z = np.concatenate([z, sz])
    return _finish(x, y, z, rng, shuffle)


def query_pairs(points_bbox, n: int, seed: int = 7) -> np.ndarray:
    """n start/goal pairs uniform in the bbox: rows (sx, sy, gx, gy, gz=0)."""
    (x0, x1), (y0, y1) = points_bbox
    rng = np.random.default_rng(seed)
    q = np.empty((n, 5), dtype=np.float32)
    q[:, 0] = rng.uniform(x0, x1, n)
    q[:, 1] = rng.uniform(y0, y1, n)
    q[:, 2] = rng.uniform(x0, x1, n)
    q[:, 3] = rng.uniform(y0, y1, n)
    q[:, 4] = 0.0
    return q


def bbox(points: np.ndarray):
    return ((float(points[:, 0].min()), float(points[:, 0].max())),
            (float(points[:, 1].min()), float(points[:, 1].max())))
