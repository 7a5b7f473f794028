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


# ---------------------------------------------------------------------------------------------
# Dense synthetic scans (PCD files) for the file-level drivers PiecewiseICP_pair_call / _4D_call
# ---------------------------------------------------------------------------------------------
def make_scan(extent=4.0, spacing=0.01, seed=1, motion=None, bump=None, noise=3e-4):
    """Dense scan of the synthetic terrain: jittered grid of `spacing` over a square of side
    `extent` [m], range noise along the normal; optional local change `bump` = (cx, cy, radius,
    height) and rigid `motion` (rx, ry, rz, tx, ty, tz) applied to the whole scan."""
    rng = np.random.default_rng(np.random.PCG64(seed))
    n = int(round(extent / spacing))
    ix, iy = np.meshgrid(np.arange(n), np.arange(n), indexing="xy")
    x = (ix.ravel() - n / 2) * spacing + rng.uniform(-0.3 * spacing, 0.3 * spacing, n * n)
    y = (iy.ravel() - n / 2) * spacing + rng.uniform(-0.3 * spacing, 0.3 * spacing, n * n)
    z, zx, zy = _terrain(3.0 * x, 3.0 * y)            # a steeper, smaller-scale version of the terrain
    z = z / 6.0
    nrm = np.stack([-zx / 2.0, -zy / 2.0, np.ones_like(z)], 1)
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    pts = np.stack([x, y, z], 1) + rng.normal(0, noise, (n * n, 1)) * nrm
    if bump is not None:
        cx, cy, rad, hgt = bump
        d2 = (x - cx) ** 2 + (y - cy) ** 2
        pts[:, 2] += hgt * np.exp(-d2 / (2 * (rad / 2.5) ** 2)) * (d2 < rad * rad)
    if motion is not None:
        T = rigid_matrix(*motion)
        pts = pts @ T[:3, :3].T + T[:3, 3]
    return pts.astype(np.float32)


def write_pcd(path, xyz):
    xyz = np.ascontiguousarray(xyz, np.float32)
    with open(path, "wb") as f:
        f.write(("# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS x y z\nSIZE 4 4 4\nTYPE F F F\n"
                 "COUNT 1 1 1\nWIDTH %d\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS %d\nDATA binary\n" % (len(xyz), len(xyz))).encode())
        f.write(xyz.tobytes())


def write_config(path, p1, p2, res=0.01, sv=0.1, dtinit=0.05, dtmin=0.004, manual_res=1, manual_dt=1, crlf=True):
    """The 11-line positional configuration file of the reference (configuration_files/*.txt)."""
    lines = [f"FolderFilePath1: {p1}", f"FolderFilePath2: {p2}", f"isSetResSVsize:{manual_res}", f"PCres1:{res}",
             f"PCres2:{res}", f"SVsize1:{sv}", f"SVsize2:{sv}", f"isSetDTinit:{manual_dt}", f"DTinit:{dtinit}",
             f"DisThrhdmin:{dtmin}", "isVisual:0"]
    with open(path, "w", newline="") as f:
        f.write(("\r\n" if crlf else "\n").join(lines))      # no trailing newline, like the shipped files


def make_series(folder, n_epochs=4, extent=3.0, spacing=0.01, seed=100):
    """Writes Epoch_001.pcd .. Epoch_00N.pcd (each with its own rigid motion and a growing local
    change) and defined_transformations.txt (ground truth: epoch -> reference)."""
    import os
    os.makedirs(folder, exist_ok=True)
    rng = np.random.default_rng(seed)
    gt = []
    for e in range(n_epochs):
        motion = None if e == 0 else tuple(np.concatenate([rng.uniform(-0.004, 0.004, 3), rng.uniform(-0.01, 0.01, 3)]))
        bump = None if e == 0 else (0.4, -0.3, 0.35, 0.02 * e)
        pts = make_scan(extent, spacing, seed + e, motion, bump)
        write_pcd(os.path.join(folder, "Epoch_%03d.pcd" % (e + 1)), pts)
        gt.append(np.eye(4) if motion is None else np.linalg.inv(rigid_matrix(*motion)))
    with open(os.path.join(folder, "..", "defined_transformations.txt") if False else os.path.join(os.path.dirname(folder.rstrip("/")), "defined_transformations.txt"), "w") as f:
        for e, T in enumerate(gt):
            f.write("%d\n" % (e + 1))
            for r in range(4):
                f.write(" ".join("%.12f" % v for v in T[r]) + "\n")
    return gt
