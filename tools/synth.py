"""Synthetic inputs of the BASELINE configs, shared by bench.py and the tests.

  room_cloud   -- config 1 / 5: an S3DIS-sized room sampled on its surfaces (floor, ceiling, four walls of a
                  ~7 x 5 x 3 m box plus a few axis-aligned furniture boxes, sigma = 5 mm noise), xyz float32 >= 0
                  (the reference subtracts the minimum, utils/data_prepare_s3dis.py:49-50), uint8 rgb, uint8 labels
                  that are piecewise constant per surface.  1 M points -> about 10 points per 0.04 m voxel.
  scan_cloud   -- config 3: a terrestrial-scan-like cloud (ground plane + facades, 200 x 200 x 30 m, 1/r^2 density
                  fall-off from a scanner at the origin: heavy voxels near the scanner, 64-bit voxel keys at 0.06 m),
                  generated ON THE DEVICE with a seeded torch generator so that every rank can hold the same 80 M points
                  without a host array.
  room_sizes   -- config 5: 272 raw room sizes, log-normal around `median`.
"""
import numpy as np

ROOM = np.array([7.0, 5.0, 3.0])
# furniture: (min corner, size), axis-aligned boxes standing on the floor
_FURNITURE = (((1.0, 1.0, 0.0), (1.6, 0.8, 0.75)), ((4.0, 0.5, 0.0), (0.6, 0.6, 1.9)), ((3.0, 3.2, 0.0), (2.0, 1.0, 0.45)),
              ((5.6, 3.6, 0.0), (1.0, 1.0, 1.1)), ((0.3, 3.4, 0.0), (0.5, 1.4, 2.0)))


def _surfaces(scale):
    """List of (origin, edge u, edge v, label) rectangles: room shell + furniture faces."""
    X, Y, Z = ROOM * scale
    s = [((0, 0, 0), (X, 0, 0), (0, Y, 0), 1),   # floor
         ((0, 0, Z), (X, 0, 0), (0, Y, 0), 0),   # ceiling
         ((0, 0, 0), (X, 0, 0), (0, 0, Z), 2), ((0, Y, 0), (X, 0, 0), (0, 0, Z), 2),
         ((0, 0, 0), (0, Y, 0), (0, 0, Z), 2), ((X, 0, 0), (0, Y, 0), (0, 0, Z), 2)]
    for k, (o, d) in enumerate(_FURNITURE):
        o = np.array(o) * scale
        d = np.array(d) * scale
        lab = 7 + k % 6
        s.append((o + (0, 0, d[2]), (d[0], 0, 0), (0, d[1], 0), lab))                  # top
        s.append((o, (d[0], 0, 0), (0, 0, d[2]), lab))
        s.append((o + (0, d[1], 0), (d[0], 0, 0), (0, 0, d[2]), lab))
        s.append((o, (0, d[1], 0), (0, 0, d[2]), lab))
        s.append((o + (d[0], 0, 0), (0, d[1], 0), (0, 0, d[2]), lab))
    return s


def room_cloud(n, seed, scale=1.0, noise=0.005):
    """(xyz float32 (n,3), rgb uint8 (n,3), labels uint8 (n,)) of one synthetic room."""
    rng = np.random.default_rng(seed)
    surf = _surfaces(scale)
    area = np.array([np.linalg.norm(np.cross(u, v)) for _, u, v, _ in surf])
    which = rng.choice(len(surf), size=n, p=area / area.sum())
    a, b = rng.random(n), rng.random(n)
    O = np.array([np.asarray(s[0], float) for s in surf])[which]
    U = np.array([np.asarray(s[1], float) for s in surf])[which]
    V = np.array([np.asarray(s[2], float) for s in surf])[which]
    xyz = O + a[:, None] * U + b[:, None] * V + rng.normal(0.0, noise, (n, 3))
    xyz -= xyz.min(axis=0)
    lab = np.array([s[3] for s in surf], np.uint8)[which]
    rgb = rng.integers(0, 256, (n, 3)).astype(np.uint8)
    return xyz.astype(np.float32), rgb, lab


def room_sizes(n_rooms=272, median=1_000_000, sigma=0.5, seed=100):
    rng = np.random.default_rng(seed)
    return np.maximum((median * np.exp(rng.normal(0.0, sigma, n_rooms))).astype(np.int64), 20_000)


def scan_cloud(n, seed, device, chunk=8_000_000):
    """(xyz float32 (n,3), rgb float32 (n,3), labels int32 (n,)) cuda tensors of one synthetic terrestrial scan."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(int(seed))
    xyz = torch.empty((n, 3), dtype=torch.float32, device=device)
    lab = torch.empty((n,), dtype=torch.int32, device=device)
    for b in range(0, n, chunk):
        m = min(chunk, n - b)
        u = torch.rand((m, 5), generator=g, device=device)
        # log-uniform range: areal density on the ground ~ 1/r^2, r in [0.6 m, 100 m]
        r = 0.6 * torch.pow(torch.tensor(100.0 / 0.6, device=device), u[:, 0])
        th = u[:, 1] * (2.0 * np.pi)
        x, y = r * torch.cos(th), r * torch.sin(th)
        z = 0.03 * (u[:, 2] - 0.5)
        # 30 % of the returns hit a facade: the ray stops at the first of four walls (|x| = 40 or |y| = 55 m)
        wall = u[:, 3] < 0.3
        t = torch.minimum(40.0 / torch.abs(torch.cos(th)).clamp_min(1e-6), 55.0 / torch.abs(torch.sin(th)).clamp_min(1e-6))
        hit = wall & (t < 100.0)
        x = torch.where(hit, t * torch.cos(th), x)
        y = torch.where(hit, t * torch.sin(th), y)
        z = torch.where(hit, 30.0 * u[:, 4] * u[:, 2], z)
        p = torch.stack([x, y, z], dim=1)
        p += 0.004 * torch.randn((m, 3), generator=g, device=device)
        xyz[b:b + m] = p
        lab[b:b + m] = torch.where(hit, 2 + (th * (4.0 / np.pi)).to(torch.int32) % 6, torch.zeros_like(lab[b:b + m]))
        del u, r, th, x, y, z, wall, t, hit, p
    rgb = torch.randint(0, 256, (n, 3), generator=g, device=device, dtype=torch.int32).to(torch.float32)
    return xyz, rgb, lab


def room_cloud_device(n, seed, device, scale=1.0, noise=0.005):
    """Device-side version of room_cloud (same surfaces, torch generator): (xyz f32, rgb f32, labels i32) cuda tensors."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(int(seed))
    surf = _surfaces(scale)
    area = torch.tensor([float(np.linalg.norm(np.cross(u, v))) for _, u, v, _ in surf], device=device)
    which = torch.multinomial(area / area.sum(), n, replacement=True, generator=g)
    O = torch.tensor(np.array([np.asarray(s[0], float) for s in surf]), dtype=torch.float32, device=device)[which]
    U = torch.tensor(np.array([np.asarray(s[1], float) for s in surf]), dtype=torch.float32, device=device)[which]
    V = torch.tensor(np.array([np.asarray(s[2], float) for s in surf]), dtype=torch.float32, device=device)[which]
    ab = torch.rand((n, 2), generator=g, device=device)
    xyz = O + ab[:, :1] * U + ab[:, 1:] * V + noise * torch.randn((n, 3), generator=g, device=device)
    xyz -= xyz.min(dim=0).values
    lab = torch.tensor([s[3] for s in surf], dtype=torch.int32, device=device)[which]
    rgb = torch.randint(0, 256, (n, 3), generator=g, device=device, dtype=torch.int32).to(torch.float32)
    return xyz.contiguous(), rgb, lab
