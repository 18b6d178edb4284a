"""Synthetic inputs for the five BASELINE.json configs: planar-scene frames,
tracker maps, bundle-adjustment graphs.  Pure numpy, fixed seeds, no I/O.  This is data
generation (plumbing), not the hot path; it never imports oracle/ — callers that need FAST corners
to build a map pass a ``detect`` callable (tests pass the oracle's, bench.py the product's).
"""
from __future__ import annotations

import numpy as np

CAMERA_PARAMS = np.array([1.0803, 1.43987, 0.519983, 0.548655, 0.244943])  # config/camera.cfg:7
TEX_SIZE = 2048
PLANE_EXTENT = 4.0  # world units covered by the texture, centred on the origin


# ---------------------------------------------------------------------------------------------
# small SE3 / camera helpers (float64 numpy; used for data generation only)
# ---------------------------------------------------------------------------------------------
def so3_exp(w):
    w = np.asarray(w, float)
    th = np.linalg.norm(w)
    K = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])
    if th < 1e-9:
        return np.eye(3) + K
    return np.eye(3) + np.sin(th) / th * K + (1 - np.cos(th)) / th**2 * (K @ K)


def se3_exp(mu):
    """SE3 exponential, mu = (translation part, rotation vector); returns (R, t)."""
    mu = np.asarray(mu, float)
    u, w = mu[:3], mu[3:]
    th = np.linalg.norm(w)
    K = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])
    if th < 1e-9:
        V = np.eye(3) + 0.5 * K
    else:
        V = np.eye(3) + (1 - np.cos(th)) / th**2 * K + (th - np.sin(th)) / th**3 * (K @ K)
    return so3_exp(w), V @ u


def se3_to12(R, t):
    return np.concatenate([np.asarray(R, float).reshape(9), np.asarray(t, float).reshape(3)])


def se3_from12(p):
    p = np.asarray(p, float)
    return p[:9].reshape(3, 3), p[9:12].copy()


def se3_mul(a, b):
    return a[0] @ b[0], a[0] @ b[1] + a[1]


class AtanCamera:
    """numpy restatement of the ATAN/FOV camera model used only to *generate* data."""

    def __init__(self, width, height, params=CAMERA_PARAMS):
        p = np.asarray(params, float)
        self.W, self.H, self.p = width, height, p
        self.focal = np.array([width * p[0], height * p[1]])
        self.center = np.array([width * p[2] - 0.5, height * p[3] - 0.5])
        self.w = p[4]
        self.tan2 = 2.0 * np.tan(self.w / 2.0)

    def unproject(self, im):
        im = np.asarray(im, float)
        d = (im - self.center) / self.focal
        dr = np.sqrt((d * d).sum(-1))
        r = np.tan(dr * self.w) / self.tan2
        f = np.where(dr > 0.01, r / np.maximum(dr, 1e-300), 1.0)
        return d * f[..., None]

    def project(self, cam):
        cam = np.asarray(cam, float)
        r = np.sqrt((cam * cam).sum(-1))
        fac = np.where(r < 0.001, 1.0, np.arctan(r * self.tan2) / (self.w * np.maximum(r, 1e-300)))
        return self.center + self.focal * (fac[..., None] * cam)


# ---------------------------------------------------------------------------------------------
# scene, trajectory, renderer
# ---------------------------------------------------------------------------------------------
def make_texture(seed=20260101, size=TEX_SIZE, n_rects=6000):
    rng = np.random.default_rng(seed)
    tex = np.full((size, size), 128, np.uint8)
    xs = rng.integers(0, size, n_rects)
    ys = rng.integers(0, size, n_rects)
    ws = rng.integers(8, 97, n_rects)
    hs = rng.integers(8, 97, n_rects)
    gs = rng.integers(0, 256, n_rects)
    for x, y, w, h, g in zip(xs, ys, ws, hs, gs):
        tex[y:y + h, x:x + w] = g
    # one 3x3 box blur
    t = np.pad(tex.astype(np.uint32), 1, mode="edge")
    acc = np.zeros((size, size), np.uint32)
    for dy in range(3):
        for dx in range(3):
            acc += t[dy:dy + size, dx:dx + size]
    return (acc // 9).astype(np.uint8)


def trajectory(n_frames=256, seed=20260101, height=1.0, tilt_deg=15.0, step=0.01, yaw_deg=0.2):
    """Camera-from-world poses (n,12): height 1 above the z=0 plane, tilted, drifting and yawing."""
    rng = np.random.default_rng(seed)
    heading = rng.uniform(0, 2 * np.pi)
    start = rng.uniform(-0.3, 0.3, 2)
    poses = np.zeros((n_frames, 12))
    tilt = np.deg2rad(tilt_deg)
    for f in range(n_frames):
        yaw = heading + np.deg2rad(yaw_deg) * f
        Rz = np.array([[np.cos(yaw), -np.sin(yaw), 0], [np.sin(yaw), np.cos(yaw), 0], [0, 0, 1]])
        a = np.pi - tilt
        Rx = np.array([[1, 0, 0], [0, np.cos(a), -np.sin(a)], [0, np.sin(a), np.cos(a)]])
        Rwc = Rz @ Rx  # camera axes in world coordinates; camera looks along its +z (downwards)
        c = np.array([start[0] + step * f * np.cos(heading), start[1] + step * f * np.sin(heading), height])
        R = Rwc.T
        poses[f] = se3_to12(R, -R @ c)
    return poses


def render_frame(tex, cam: AtanCamera, pose12):
    """Inverse-map every pixel through UnProject -> plane z=0 -> texture (bilinear), u8."""
    R, t = se3_from12(pose12)
    c = -R.T @ t
    v, u = np.mgrid[0:cam.H, 0:cam.W]
    ray_c = cam.unproject(np.stack([u, v], -1).astype(float))
    ray_c = np.concatenate([ray_c, np.ones(ray_c.shape[:-1] + (1,))], -1)
    ray_w = ray_c @ R  # R^T applied to each row vector
    s = -c[2] / np.where(np.abs(ray_w[..., 2]) < 1e-12, -1e-12, ray_w[..., 2])
    X = c[0] + s * ray_w[..., 0]
    Y = c[1] + s * ray_w[..., 1]
    size = tex.shape[0]
    tx = (X / PLANE_EXTENT + 0.5) * size - 0.5
    ty = (Y / PLANE_EXTENT + 0.5) * size - 0.5
    ok = (s > 0) & (tx >= 0) & (ty >= 0) & (tx < size - 1) & (ty < size - 1)
    txc = np.clip(tx, 0, size - 1.001)
    tyc = np.clip(ty, 0, size - 1.001)
    x0 = txc.astype(np.int64)
    y0 = tyc.astype(np.int64)
    fx = txc - x0
    fy = tyc - y0
    T = tex.astype(np.float64)
    val = (1 - fy) * ((1 - fx) * T[y0, x0] + fx * T[y0, x0 + 1]) + fy * ((1 - fx) * T[y0 + 1, x0] + fx * T[y0 + 1, x0 + 1])
    img = np.where(ok, val + 0.5, 128.0)
    return np.clip(img, 0, 255).astype(np.uint8)


def render_sequence(width, height, n_frames, seed=20260101, tex=None, params=CAMERA_PARAMS):
    tex = make_texture() if tex is None else tex
    cam = AtanCamera(width, height, params)
    poses = trajectory(n_frames, seed)
    frames = np.stack([render_frame(tex, cam, p) for p in poses])
    return frames, poses


# ---------------------------------------------------------------------------------------------
# tracker map
# ---------------------------------------------------------------------------------------------
def shi_tomasi(im, x, y, half=3):
    """ImageProcess::ShiTomasiScoreAtPoint (reference src/ImageProcess.cc:20-47)."""
    p = im[y - half - 1:y + half + 2, x - half - 1:x + half + 2].astype(np.float64)
    dx = p[1:-1, 2:] - p[1:-1, :-2]
    dy = p[2:, 1:-1] - p[:-2, 1:-1]
    n = (2 * half + 1) ** 2
    xx, yy, xy = (dx * dx).sum() / (2.0 * n), (dy * dy).sum() / (2.0 * n), (dx * dy).sum() / (2.0 * n)
    return 0.5 * (xx + yy - np.sqrt((xx + yy) ** 2 - 4 * (xx * yy - xy * xy)))


def build_map(frames, poses, detect, cam: AtanCamera, kf_indices=(0, 64, 128, 192),
              per_level=(600, 250, 100, 50), seed=20260101, min_score=70.0):
    """Map points from FAST corners of the source keyframes.

    detect(image) -> list over 4 levels of (pixels HxW u8, corners (n,2) int32 [x,y]).
    Returns (keyframe_images, map dict of SoA arrays ready for Tracker.set_map()).
    World position = corner ray ∩ plane z=0; pixel-right/down vectors exactly as
    MapPoint::RefreshPixelVectors (reference src/Map.cc:40-65) with v3Normal_NC=(0,0,-1) and
    +-1-level-pixel rays as in MapMaker::AddPointEpipolar (src/MapMaker.cc:651-668).
    """
    rng = np.random.default_rng(seed)
    kf_indices = [k for k in kf_indices if k < len(frames)]
    cands = [[] for _ in range(4)]
    for kid, fi in enumerate(kf_indices):
        levels = detect(frames[fi])
        for l, (pix, corners) in enumerate(levels):
            h, w = pix.shape
            for (x, y) in corners:
                if not (x >= 10 and y >= 10 and x < w - 10 and y < h - 10):
                    continue
                if shi_tomasi(pix, int(x), int(y)) > min_score:
                    cands[l].append((kid, int(x), int(y)))
    out = dict(world_pos=[], pixel_right_w=[], pixel_down_w=[], src_kf=[], src_level=[], ir_center=[])
    normal = np.array([0.0, 0.0, -1.0])
    for l in range(4):
        c = cands[l]
        if not c:
            continue
        pick = rng.permutation(len(c))[:per_level[l]]
        for i in sorted(pick):
            kid, x, y = c[i]
            R, t = se3_from12(poses[kf_indices[kid]])
            sc = 1 << l
            root = np.array([(x + 0.5) * sc - 0.5, (y + 0.5) * sc - 0.5])
            def ray(px):
                v = np.append(cam.unproject(px), 1.0)
                return v / np.linalg.norm(v)
            centre, right, down = ray(root), ray(root + [sc, 0]), ray(root + [0, sc])
            cw = -R.T @ t
            rw = R.T @ centre
            if abs(rw[2]) < 1e-9:
                continue
            s = -cw[2] / rw[2]
            if s <= 0:
                continue
            world = cw + s * rw
            # RefreshPixelVectors
            plane_pt_c = R @ world + t
            cam_height = abs(plane_pt_c @ normal)
            pr, rr, dr = abs(centre @ normal), abs(right @ normal), abs(down @ normal)
            c_on = centre * cam_height / pr
            r_on = right * cam_height / rr
            d_on = down * cam_height / dr
            out["world_pos"].append(world)
            out["pixel_right_w"].append(R.T @ (r_on - c_on))
            out["pixel_down_w"].append(R.T @ (d_on - c_on))
            out["src_kf"].append(kid)
            out["src_level"].append(l)
            out["ir_center"].append((x, y))
    m = {k: np.asarray(v) for k, v in out.items()}
    for k in ("world_pos", "pixel_right_w", "pixel_down_w"):
        m[k] = m[k].reshape(-1, 3).astype(np.float64)
    m["src_kf"] = m["src_kf"].astype(np.int32)
    m["src_level"] = m["src_level"].astype(np.int32)
    m["ir_center"] = m["ir_center"].reshape(-1, 2).astype(np.int32)
    return [frames[i] for i in kf_indices], m


def perturb_pose(pose12, rng, sigma=0.005):
    xi = rng.normal(0, sigma, 6)
    return se3_to12(*se3_mul(se3_exp(xi), se3_from12(pose12)))


# ---------------------------------------------------------------------------------------------
# bundle-adjustment graphs (configs C3 / C4)
# ---------------------------------------------------------------------------------------------
def make_ba_graph(n_cams=50, n_points=5000, n_meas=20000, seed=42, width=640, height=480,
                  params=CAMERA_PARAMS, outlier_frac=0.02):
    """Synthetic keyframe/point graph: points on a 4x4x0.5 slab, cameras on a
    serpentine grid 1.0 above looking down +-20 deg, each point seen by its k nearest cameras that
    see it (every point >= 2 observations, totals exact), pyramid level ~ {.5,.25,.15,.1},
    sigma^2 = 4^level, pixel noise N(0,(0.5*2^level)^2), 2% gross outliers, perturbed initial
    state, camera 0 fixed.  Measurements are inserted camera-major then by point."""
    rng = np.random.default_rng(seed)
    cam = AtanCamera(width, height, params)
    # slab side: 4 units unless there are so few cameras that the requested mean
    # track length cannot be reached with this lens at height 1 (footprint ~0.63 units^2): then the
    # slab shrinks until every point has >= 2 observers and the totals fit.
    mean_k = n_meas / n_points
    side = min(4.0, float(np.sqrt(n_cams * 0.63 / (2.5 * mean_k))))
    g = int(np.ceil(np.sqrt(n_cams)))
    base = int(np.ceil(mean_k))
    for attempt in range(40):
        hs = side / 2
        pts = np.column_stack([rng.uniform(-hs, hs, n_points), rng.uniform(-hs, hs, n_points), rng.uniform(-0.25, 0.25, n_points)])
        cam_R, cam_t, cam_c = [], [], []
        for i in range(n_cams):
            r, cidx = divmod(i, g)
            if r % 2:
                cidx = g - 1 - cidx
            c = np.array([-0.8 * hs + 1.6 * hs * cidx / max(g - 1, 1), -0.8 * hs + 1.6 * hs * r / max(g - 1, 1), 1.0])
            ax, ay, az = np.deg2rad(rng.uniform(-20, 20, 2)).tolist() + [rng.uniform(0, 2 * np.pi)]
            Rx = np.array([[1, 0, 0], [0, np.cos(np.pi + ax), -np.sin(np.pi + ax)], [0, np.sin(np.pi + ax), np.cos(np.pi + ax)]])
            Ry = np.array([[np.cos(ay), 0, np.sin(ay)], [0, 1, 0], [-np.sin(ay), 0, np.cos(ay)]])
            Rz = np.array([[np.cos(az), -np.sin(az), 0], [np.sin(az), np.cos(az), 0], [0, 0, 1]])
            R = (Rz @ Ry @ Rx).T
            cam_R.append(R); cam_t.append(-R @ c); cam_c.append(c)
        cam_R, cam_t, cam_c = np.array(cam_R), np.array(cam_t), np.array(cam_c)

        def visibility(P):
            uv_ = np.zeros((n_cams, len(P), 2))
            vis_ = np.zeros((n_cams, len(P)), bool)
            for j in range(n_cams):
                pc = P @ cam_R[j].T + cam_t[j]
                z = pc[:, 2]
                ok = z > 0.1
                ip = pc[:, :2] / np.where(ok, z, 1.0)[:, None]
                q = cam.project(ip)
                uv_[j] = q
                vis_[j] = ok & (q[:, 0] >= 8) & (q[:, 1] >= 8) & (q[:, 0] < width - 8) & (q[:, 1] < height - 8) & ((ip * ip).sum(1) < 1.0)
            return uv_, vis_

        uv, vis = visibility(pts)
        for _ in range(20):  # re-draw the points nobody (or only one camera) sees
            lonely = np.flatnonzero(vis.sum(0) < 2)
            if len(lonely) == 0:
                break
            pts[lonely] = np.column_stack([rng.uniform(-hs, hs, len(lonely)), rng.uniform(-hs, hs, len(lonely)), rng.uniform(-0.25, 0.25, len(lonely))])
            uv_l, vis_l = visibility(pts[lonely])
            uv[:, lonely], vis[:, lonely] = uv_l, vis_l
        nvis = vis.sum(0)
        level_cap = np.minimum(nvis, 2 * base + 2)
        if nvis.min() >= 2 and int(level_cap.sum()) >= n_meas:
            break
        side *= 0.85
    else:
        raise ValueError("could not build a graph with the requested sizes")
    d2 = ((pts[None, :, :2] - cam_c[:, None, :2]) ** 2).sum(-1)
    d2 = np.where(vis, d2, np.inf)
    order = np.argsort(d2, axis=0)  # cameras by distance, per point
    # k per point: 2 each, then hand out the remainder round-robin up to the cap
    k = np.full(n_points, 2)
    remaining = n_meas - int(k.sum())
    if remaining < 0:
        raise ValueError("n_meas must be at least 2 per point")
    prio = rng.permutation(n_points)
    while remaining > 0:
        for i in prio:
            if remaining == 0:
                break
            if k[i] < level_cap[i]:
                k[i] += 1; remaining -= 1
    mc, mp = [], []
    for i in range(n_points):
        for j in order[:k[i], i]:
            mc.append(j); mp.append(i)
    mc, mp = np.array(mc, np.int32), np.array(mp, np.int32)
    o = np.lexsort((mp, mc))  # camera-major, then point
    mc, mp = mc[o], mp[o]
    n = len(mc)
    level = rng.choice(4, n, p=[0.5, 0.25, 0.15, 0.1])
    noise = rng.normal(0, 1, (n, 2)) * (0.5 * 2.0 ** level)[:, None]
    meas_uv = uv[mc, mp] + noise
    outl = rng.random(n) < outlier_frac
    meas_uv[outl] += rng.uniform(-20, 20, (int(outl.sum()), 2))
    pts0 = pts + rng.normal(0, 0.01, pts.shape)
    se3 = np.zeros((n_cams, 12))
    for j in range(n_cams):
        Rt = (cam_R[j], cam_t[j])
        if j > 0:
            Rt = se3_mul(se3_exp(rng.normal(0, 0.005, 6)), Rt)
        se3[j] = se3_to12(*Rt)
    fixed = np.zeros(n_cams, np.int32)
    fixed[0] = 1
    return dict(cam_se3=se3, cam_fixed=fixed, points=pts0, meas_cam=mc, meas_point=mp, meas_uv=meas_uv,
                meas_sigma_sq=(4.0 ** level), true_points=pts,
                true_se3=np.array([se3_to12(cam_R[j], cam_t[j]) for j in range(n_cams)]), width=width, height=height)
