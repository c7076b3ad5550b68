"""Seeded synthetic inputs for the hot path.

BA: UAV-style nadir survey (serpentine track at height 100 over a z in U(-5,5) ground
plane, 80 % forward / 60 % side overlap, N(0, 3 deg) attitude noise), exact projection +
N(0, 0.5 px) noise, 2 % gross outliers, initial state = truth perturbed, gauge as in
SequentialMapper::adjust_global_bundle (src/sfm/sequential_mapper.cc:1095-1097: image 0
FIXED, image 1 FIXED_X).  Matching: unit-norm SURF-like descriptors where image j shares
60 % of image i's descriptors (+N(0, 0.05), renormalised, permuted).
"""
import numpy as np

from ._abi import MM_INTR_STRIDE, MM_MODEL_CATA, MM_MODEL_OPENCV, MM_MODEL_PINHOLE
from .ba import FlatProblem

IMG_W, IMG_H = 1280.0, 960.0
HEIGHT = 100.0

INTRINSICS = {
    MM_MODEL_PINHOLE: [1000.0, 1000.0, 640.0, 480.0],
    MM_MODEL_OPENCV: [1000.0, 1000.0, 640.0, 480.0, -0.1, 0.02, 1e-3, 1e-3],
    MM_MODEL_CATA: [1000.0, 1000.0, 640.0, 480.0, -0.1, 0.02, 1e-3, 1e-3, 0.2],
}


def _rodrigues(rvec):
    """rotation matrices of a batch of angle-axis vectors [n,3] -> [n,3,3]."""
    rvec = np.asarray(rvec, dtype=np.float64).reshape(-1, 3)
    th = np.linalg.norm(rvec, axis=1)
    k = rvec / np.maximum(th, 1e-300)[:, None]
    K = np.zeros((len(rvec), 3, 3))
    K[:, 0, 1], K[:, 0, 2] = -k[:, 2], k[:, 1]
    K[:, 1, 0], K[:, 1, 2] = k[:, 2], -k[:, 0]
    K[:, 2, 0], K[:, 2, 1] = -k[:, 1], k[:, 0]
    s, c = np.sin(th)[:, None, None], np.cos(th)[:, None, None]
    R = np.eye(3)[None] + s * K + (1 - c) * (K @ K)
    R[th == 0] = np.eye(3)
    return R


def project(model, intr, xc):
    """world2image of camera-frame points xc [n,3] (src/base3d/camera_models.h:111-302)."""
    x, y, z = xc[:, 0], xc[:, 1], xc[:, 2]
    if model == MM_MODEL_CATA:
        z = z + intr[8] * np.sqrt(x * x + y * y + z * z)
    u, v = x / z, y / z
    if model != MM_MODEL_PINHOLE:
        k1, k2, p1, p2 = intr[4:8]
        u2, uv, v2 = u * u, u * v, v * v
        r2 = u2 + v2
        rad = k1 * r2 + k2 * r2 * r2
        du = u * rad + 2 * p1 * uv + p2 * (r2 + 2 * u2)
        dv = v * rad + 2 * p2 * uv + p1 * (r2 + 2 * v2)
        u, v = u + du, v + dv
    return np.stack([intr[0] * u + intr[2], intr[1] * v + intr[3]], axis=1)


def grid_shape(n_img):
    """strips x images-per-strip for a roughly 2:1 survey block."""
    best = None
    for strips in range(1, n_img + 1):
        if n_img % strips == 0:
            per = n_img // strips
            score = abs(per / strips - 2.0)
            if best is None or score < best[0]:
                best = (score, strips, per)
    return best[1], best[2]


def make_ba_problem(n_img, n_obs_target, track_len, model=MM_MODEL_PINHOLE, seed=0xBA5E,
                    refine_camera_params=False, n_cam=1, exact_track=False,
                    outlier_frac=0.02, noise_px=0.5, perturb=(0.01, 0.1, 0.2),
                    models=None):
    """Returns (FlatProblem at the perturbed initial state, dict of ground truth).

    `models`: optional list of model codes, one per camera (mixed PINHOLE+OPENCV rigs)."""
    from scipy.spatial.transform import Rotation
    rng = np.random.default_rng(seed)
    strips, per = grid_shape(n_img)
    foot_x, foot_y = IMG_W / 1000.0 * HEIGHT, IMG_H / 1000.0 * HEIGHT
    dx, dy = 0.2 * foot_x, 0.4 * foot_y
    s_idx = np.repeat(np.arange(strips), per)
    k_idx = np.tile(np.arange(per), strips)
    k_ser = np.where(s_idx % 2 == 0, k_idx, per - 1 - k_idx)       # serpentine
    C = np.stack([k_ser * dx, s_idx * dy, np.full(n_img, HEIGHT)], axis=1)
    att = Rotation.from_rotvec(rng.normal(0, np.deg2rad(3.0), (n_img, 3)))
    R_nadir = np.diag([1.0, -1.0, -1.0])
    R = att.as_matrix() @ R_nadir[None]
    rvec = Rotation.from_matrix(R).as_rotvec()
    tvec = -np.einsum("nij,nj->ni", R, C)

    if models is None:
        models = [model] * n_cam
    n_cam = len(models)
    cam_model = np.array(models, dtype=np.int32)
    intr = np.zeros((n_cam, MM_INTR_STRIDE))
    for c, m in enumerate(models):
        intr[c, :len(INTRINSICS[m])] = INTRINSICS[m]
    img_cam = (np.arange(n_img) % n_cam).astype(np.int32)

    n_pt_target = int(np.ceil(n_obs_target / track_len))
    x0, x1 = -0.3 * foot_x, (per - 1) * dx + 0.3 * foot_x
    y0, y1 = -0.3 * foot_y, (strips - 1) * dy + 0.3 * foot_y
    cam_of = np.full((strips, per), -1, dtype=np.int64)
    cam_of[s_idx, k_ser] = np.arange(n_img)
    wk, ws = 4, 2                                   # candidate window (+-) along / across

    pts_l, oi_l, op_l, oxy_l = [], [], [], []
    n_pt = 0
    oversample = 1.6
    while n_pt < n_pt_target:
        m = int((n_pt_target - n_pt) * oversample) + 64
        m = min(m, 400000)
        X = np.stack([rng.uniform(x0, x1, m), rng.uniform(y0, y1, m), rng.uniform(-5, 5, m)], axis=1)
        kc = np.rint(X[:, 0] / dx).astype(np.int64)
        sc = np.rint(X[:, 1] / dy).astype(np.int64)
        ok_l, img_l, uv_l = [], [], []
        for ds in range(-ws, ws + 1):
            for dk in range(-wk, wk + 1):
                s, k = sc + ds, kc + dk
                inb = (s >= 0) & (s < strips) & (k >= 0) & (k < per)
                img = cam_of[np.clip(s, 0, strips - 1), np.clip(k, 0, per - 1)]
                Xc = np.einsum("nij,nj->ni", R[img], X) + tvec[img]
                uv = np.zeros((m, 2))
                for c in range(n_cam):
                    sel = img_cam[img] == c
                    if sel.any():
                        uv[sel] = project(int(cam_model[c]), intr[c], Xc[sel])
                vis = inb & (Xc[:, 2] > 1.0) & (uv[:, 0] >= 0) & (uv[:, 0] < IMG_W) & (uv[:, 1] >= 0) & (uv[:, 1] < IMG_H)
                vis &= rng.random(m) < 0.7
                ok_l.append(vis); img_l.append(img); uv_l.append(uv)
        ok = np.stack(ok_l, axis=1); img = np.stack(img_l, axis=1); uv = np.stack(uv_l, axis=1)
        # random truncation of each track to `track_len` (exact) or a length drawn around it
        prio = rng.random(ok.shape)
        prio[~ok] = 2.0
        order = np.argsort(prio, axis=1)
        nvis = ok.sum(axis=1)
        if exact_track:
            want = np.full(m, track_len)
            keep_pt = nvis >= track_len
        else:
            want = np.clip(rng.poisson(track_len - 2.0, m) + 2, 2, None)
            want = np.minimum(want, nvis)
            keep_pt = nvis >= 2
        rank = np.empty_like(order)
        np.put_along_axis(rank, order, np.arange(ok.shape[1])[None, :].repeat(m, 0), axis=1)
        sel = ok & (rank < want[:, None]) & keep_pt[:, None]
        ids = np.nonzero(keep_pt)[0]
        if n_pt + len(ids) > n_pt_target:
            ids = ids[: n_pt_target - n_pt]
            keep2 = np.zeros(m, dtype=bool); keep2[ids] = True
            sel &= keep2[:, None]
        new_index = np.full(m, -1, dtype=np.int64); new_index[ids] = n_pt + np.arange(len(ids))
        pi, ci = np.nonzero(sel)
        pts_l.append(X[ids]); oi_l.append(img[pi, ci]); op_l.append(new_index[pi]); oxy_l.append(uv[pi, ci])
        n_pt += len(ids)

    pts = np.concatenate(pts_l); obs_img = np.concatenate(oi_l); obs_pt = np.concatenate(op_l)
    obs_xy = np.concatenate(oxy_l)
    obs_xy = obs_xy + rng.normal(0, noise_px, obs_xy.shape)
    n_obs = len(obs_xy)
    out = rng.random(n_obs) < outlier_frac
    obs_xy[out] = np.stack([rng.uniform(0, IMG_W, out.sum()), rng.uniform(0, IMG_H, out.sum())], axis=1)
    # residual-block order of the reference: image-major
    order = np.lexsort((obs_pt, obs_img))
    obs_xy, obs_img, obs_pt = obs_xy[order], obs_img[order], obs_pt[order]

    truth = {"poses": np.concatenate([rvec, tvec], axis=1), "intr": intr.copy(), "pts": pts.copy(),
             "grid": (strips, per)}
    poses0 = truth["poses"].copy()
    poses0[:, :3] += rng.normal(0, perturb[0], (n_img, 3))
    poses0[:, 3:] += rng.normal(0, perturb[1], (n_img, 3))
    pts0 = pts + rng.normal(0, perturb[2], pts.shape)
    pose_const = np.zeros((n_img, 4), dtype=np.uint8)
    pose_const[0, :] = 1                      # BA_POSE_FIXED
    if n_img > 1:
        pose_const[1, 1] = 1                  # BA_POSE_FIXED_X
    # the fixed images keep their true pose (they define the datum)
    poses0[0] = truth["poses"][0]
    if n_img > 1:
        poses0[1, 3] = truth["poses"][1, 3]
    intr_const = np.full(n_cam, 0 if refine_camera_params else 1, dtype=np.uint8)
    flat = FlatProblem(poses0, pose_const, img_cam, intr, cam_model, intr_const, pts0,
                       np.zeros(len(pts0), dtype=np.uint8), obs_xy, obs_img.astype(np.int32),
                       obs_pt.astype(np.int32))
    return flat, truth


BA_CONFIGS = {
    # BASELINE.json configs -> generator arguments
    "cfg1": dict(n_img=20, n_obs_target=12000, track_len=4, seed=0xBA5E + 1),
    "cfg2": dict(n_img=500, n_obs_target=1_000_000, track_len=4, seed=0xBA5E + 2),
    "cfg4": dict(n_img=5000, n_obs_target=10_000_000, track_len=5, seed=0xBA5E + 4, exact_track=True),
    # configs[4]: mixed multi-camera (PINHOLE + OPENCV) 1000-image sequence, intrinsics refined as the mapper does by default
    "cfg5": dict(n_img=1000, n_obs_target=2_000_000, track_len=4, seed=0xBA5E + 5, models=[1, 2], refine_camera_params=True),
    "tiny": dict(n_img=8, n_obs_target=1500, track_len=4, seed=0xBA5E + 9),
    "small": dict(n_img=40, n_obs_target=40000, track_len=4, seed=0xBA5E + 10),
}


def make_descriptors(n_images, n_feat, k, seed=0xF00D, shared=0.6, noise=0.05):
    """Returns (desc [n_images, n_feat, k] fp32 unit-norm, xy [n_images, n_feat, 2] fp32).
    Image i+1 shares `shared` of image i's descriptors (noisy, renormalised, permuted)."""
    rng = np.random.default_rng(seed)
    desc = np.empty((n_images, n_feat, k), dtype=np.float32)
    xy = np.stack([rng.uniform(0, IMG_W, (n_images, n_feat)), rng.uniform(0, IMG_H, (n_images, n_feat))],
                  axis=2).astype(np.float32)
    d = rng.normal(size=(n_feat, k))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    desc[0] = d.astype(np.float32)
    n_sh = int(round(shared * n_feat))
    for i in range(1, n_images):
        prev = desc[i - 1].astype(np.float64)
        src = rng.permutation(n_feat)[:n_sh]
        a = prev[src] + rng.normal(0, noise, (n_sh, k))
        b = rng.normal(size=(n_feat - n_sh, k))
        cur = np.concatenate([a, b])
        cur /= np.linalg.norm(cur, axis=1, keepdims=True)
        perm = rng.permutation(n_feat)
        desc[i] = cur[perm].astype(np.float32)
        # shared keypoints move a little (for the max_distance mask path)
        pos = np.concatenate([xy[i - 1][src] + rng.normal(0, 3.0, (n_sh, 2)), xy[i][n_sh:]])
        xy[i] = pos[perm].astype(np.float32)
    return desc, xy
