"""Seeded synthetic planar-patch clouds at the centroid-level boundary.

A pair is what `Piecewise_ICP` (reference src/Registration.cpp:618-700) holds after
`PatchGenerationAndRefinement` + `calBPandCTSTD` (:653-664): patch centroids (CT), six boundary
points per source patch (BP, ordered Xmax,Xmin,Ymax,Ymin,Zmax,Zmin -- src/Segmentation.cpp:295-300),
per-patch sigmas (CTstd1 = sigma/n, BPstd2 = sigma -- src/Segmentation.cpp:316-319), target patch
normals, the patch point lists of the source and the two full clouds.

Generator: jittered grid (spacing s, jitter U(-0.3s, 0.3s)) on a sinusoidal terrain with slopes up
to ~45 deg, 10 % of the cells moved onto vertical wall strips so all 6 DoF are observable.  The
source is an independently re-jittered sample of the same surface, `changed` of its patches are
displaced 1-10 cm along the normal, then a rigid motion is applied.
"""
import numpy as np

SEED_TARGET, SEED_SOURCE, SEED_CHANGE = 20250606, 20250607, 20250608
# magnitude of data/data_synthetic/defined_transformations.txt:6-10
DEFAULT_MOTION = (0.013, 0.007, 0.001, 0.008, 0.004, 0.008)

_TERRAIN = [  # (A, f, g, phi, psi)
    (1.20, 0.35, 0.27, 0.3, 1.1),
    (0.45, 0.90, 1.10, 2.0, 0.4),
    (0.15, 2.30, 1.90, 0.7, 2.6),
]


def _terrain(x, y):
    z = np.zeros_like(x, dtype=np.float64)
    zx = np.zeros_like(z)
    zy = np.zeros_like(z)
    for A, f, g, ph, ps in _TERRAIN:
        sx, cx = np.sin(f * x + ph), np.cos(f * x + ph)
        sy, cy = np.sin(g * y + ps), np.cos(g * y + ps)
        z += A * sx * sy
        zx += A * f * cx * sy
        zy += A * g * sx * cy
    return z, zx, zy


def rigid_matrix(rx, ry, rz, tx, ty, tz):
    """R = Rz*Ry*Rx as in PCL's constructTransformationMatrix; float64 4x4."""
    ca, sa, cb, sb, cg, sg = np.cos(rx), np.sin(rx), np.cos(ry), np.sin(ry), np.cos(rz), np.sin(rz)
    T = np.eye(4)
    T[0, :3] = [cg * cb, -sg * ca + cg * sb * sa, sg * sa + cg * sb * ca]
    T[1, :3] = [sg * cb, cg * ca + sg * sb * sa, -cg * sa + sg * sb * ca]
    T[2, :3] = [-sb, cb * sa, cb * ca]
    T[:3, 3] = [tx, ty, tz]
    return T


def _surface(n_side_x, n_side_y, s, rng):
    """One jittered sample of the surface: points, unit normals, two in-plane axes (float64)."""
    ix, iy = np.meshgrid(np.arange(n_side_x), np.arange(n_side_y), indexing="xy")
    ix = ix.ravel()
    iy = iy.ravel()
    n = ix.size
    jx = rng.uniform(-0.3 * s, 0.3 * s, n)
    jy = rng.uniform(-0.3 * s, 0.3 * s, n)
    x = (ix - n_side_x / 2) * s + jx
    y = (iy - n_side_y / 2) * s + jy
    z, zx, zy = _terrain(x, y)
    nrm = np.stack([-zx, -zy, np.ones(n)], 1)
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    pts = np.stack([x, y, z], 1)
    # vertical wall strips: 5 % of the cells on x-walls (normal +x), 5 % on y-walls (normal +y)
    wx = (ix % 100) < 5
    wy = ((iy % 100) >= 50) & ((iy % 100) < 55) & ~wx
    if wx.any():
        x0 = ((ix[wx] // 100) * 100 - n_side_x / 2) * s
        z0, _, _ = _terrain(x0, y[wx])
        pts[wx] = np.stack([x0, y[wx], z0 + 0.6 + (ix[wx] % 100) * s + jx[wx]], 1)
        nrm[wx] = [1.0, 0.0, 0.0]
    if wy.any():
        y0 = ((iy[wy] // 100) * 100 + 50 - n_side_y / 2) * s
        z0, _, _ = _terrain(x[wy], y0)
        pts[wy] = np.stack([x[wy], y0, z0 + 0.6 + (iy[wy] % 100 - 50) * s + jy[wy]], 1)
        nrm[wy] = [0.0, 1.0, 0.0]
    # in-plane axes
    ref = np.where(np.abs(nrm[:, [2]]) < 0.9, [[0.0, 0.0, 1.0]], [[1.0, 0.0, 0.0]])
    u = np.cross(nrm, ref)
    u /= np.linalg.norm(u, axis=1, keepdims=True)
    v = np.cross(nrm, u)
    return pts, nrm, u, v


def make_pair(n, seed=SEED_TARGET, s=0.05, motion=DEFAULT_MOTION, changed=0.2, pts_per_patch=8,
              with_clouds=True, Res=0.005, DTmin=0.004):
    """Centroid-level pair with about `n` patches per cloud (n is rounded to a grid).

    Returns a dict of float32/int32 arrays: ct1, nrm1, ctstd1, ct2, bp2, bpstd2, patch_off2,
    patch_pts2, cloud1, cloud2 (+ scalars Res1/Res2/SVRes1/SVRes2/DTmin and the ground truth
    `T_true` mapping the source back onto the target).
    """
    side = int(round(np.sqrt(n)))
    nx, ny = side, max(1, int(round(n / side)))
    rng1 = np.random.default_rng(np.random.PCG64(seed))
    rng2 = np.random.default_rng(np.random.PCG64(seed + 1))
    rng3 = np.random.default_rng(np.random.PCG64(seed + 2))

    p1, n1v, u1, v1 = _surface(nx, ny, s, rng1)
    p2, n2v, u2, v2 = _surface(nx, ny, s, rng2)
    N1, N2 = len(p1), len(p2)

    # target: centroid noise 0.2 mm, noisy analytic normals, sigma ~ U(0.5,1.5) mm, n ~ U{60..100}
    ct1 = p1 + rng1.normal(0, 2e-4, p1.shape)
    nn1 = n1v + rng1.normal(0, 1e-3, n1v.shape)
    nn1 /= np.linalg.norm(nn1, axis=1, keepdims=True)
    sig1 = rng1.uniform(5e-4, 1.5e-3, N1)
    cnt1 = rng1.integers(60, 101, N1)
    ctstd1 = sig1 / cnt1

    # source: changes along the normal, then the rigid motion
    ct2 = p2 + rng2.normal(0, 2e-4, p2.shape)
    sig2 = rng2.uniform(5e-4, 1.5e-3, N2)
    is_changed = rng3.random(N2) < changed
    disp = rng3.uniform(0.01, 0.10, N2) * rng3.choice([-1.0, 1.0], N2) * is_changed
    ct2 = ct2 + disp[:, None] * n2v
    half = s / 2
    # six boundary points: +-u, +-v in plane, +-1 mm along the normal
    offs = np.stack([half * u2, -half * u2, half * v2, -half * v2, 1e-3 * n2v, -1e-3 * n2v], 1)
    bp2 = ct2[:, None, :] + offs                                  # (N2, 6, 3)

    k = int(pts_per_patch)
    ang = (np.arange(k) + 0.5) * (2 * np.pi / k)
    ring = (0.8 * half) * np.stack([np.cos(ang), np.sin(ang)], 1)  # (k,2)

    def patch_points(ct, u, v, nv, sig, rng):
        pp = (ct[:, None, :] + ring[None, :, [0]] * u[:, None, :] + ring[None, :, [1]] * v[:, None, :]
              + (rng.normal(0, 1, (len(ct), k, 1)) * sig[:, None, None]) * nv[:, None, :])
        return pp

    out = {}
    T_mov = rigid_matrix(*motion)           # motion applied to the source
    R, t = T_mov[:3, :3], T_mov[:3, 3]
    mv = lambda a: a @ R.T + t
    if with_clouds:
        pp1 = patch_points(ct1, u1, v1, n1v, sig1, rng1).reshape(-1, 3)
        pp2 = patch_points(ct2, u2, v2, n2v, sig2, rng2)
        pp2 = mv(pp2.reshape(-1, 3))
        out["cloud1"] = pp1.astype(np.float32)
        out["patch_pts2"] = pp2.astype(np.float32)
        out["cloud2"] = out["patch_pts2"].copy()
        out["patch_off2"] = (np.arange(N2 + 1) * k).astype(np.int32)
    else:
        out["cloud1"] = np.zeros((0, 3), np.float32)
        out["cloud2"] = np.zeros((0, 3), np.float32)
        out["patch_pts2"] = np.zeros((0, 3), np.float32)
        out["patch_off2"] = np.zeros(N2 + 1, np.int32)
    out["ct1"] = ct1.astype(np.float32)
    out["nrm1"] = nn1.astype(np.float32)
    out["ctstd1"] = ctstd1.astype(np.float32)
    out["ct2"] = mv(ct2).astype(np.float32)
    out["bp2"] = mv(bp2.reshape(-1, 3)).astype(np.float32)
    out["bpstd2"] = sig2.astype(np.float32)
    out["changed"] = is_changed
    out["Res1"] = out["Res2"] = float(Res)
    out["SVRes1"] = out["SVRes2"] = float(s)
    out["DTmin"] = float(DTmin)
    out["T_true"] = np.linalg.inv(T_mov)
    return out
