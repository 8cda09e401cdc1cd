"""CPU restatement of the reference's local-map candidate step.  TEST INFRASTRUCTURE ONLY: nothing
under boslam_b200/ may import this module.

Follows reference slam/tracking.py:96-128 (Tracker._track_local_map):
    :101  pose = g2o.SE3Quat(frame.R, frame.t)
    :102  pixel = self.cam.cam_map(pose * mp.pt3d)
    :103  0 <= pixel[0] < width and 0 <= pixel[1] < height
    :104  np.dot(frame.see_vector, mp.nf()) < cos60          (kept exactly as written)
    :107-110  append descriptor / 3-D point in edge order
    :119-121  np.stack, crossCheck match, distance <= d_hamming_max
    :126-128  inds_frame, inds -> kp_arr[inds_frame], pts3d[inds]

The projection arithmetic lives in third-party code that is absent from /root/reference and not
installed here (g2opy -> g2o::SE3Quat / CameraParameters over Eigen; reference pins no version,
docker/Dockerfile builds g2opy master), so its published algorithm is restated:
    Eigen  Quaternion(Matrix3)            (Shepperd's method, Eigen/src/Geometry/Quaternion.h)
    g2o    SE3Quat(R, t): normalizeRotation  (w >= 0, unit norm; g2o/types/slam3d/se3quat.h)
    Eigen  q * v = v + w * uv + q.vec x uv,  uv = 2 * (q.vec x v)
    g2o    SE3Quat::map = _r * xyz + _t;  CameraParameters::cam_map = project2d(x) * f + c
**Parity unpinned against g2o itself** (no reference test or fixture touches this path):
the oracle pins the CUDA path bit for bit, and differs from a g2o build at most in the last ulp of a
pixel, i.e. only for points within ~1e-13 px of the image border or of the cos60 threshold.
Every operation below is a separate IEEE fp64 operation in a fixed order (no BLAS, no fused
multiply-add), which is what the CUDA kernel issues with __dmul_rn / __dadd_rn / __ddiv_rn.
"""
import math

import numpy as np

from . import hamming_oracle as orc


def quaternion_from_rotation(R):
    m = np.asarray(R, np.float64)
    tr = float(m[0, 0]) + float(m[1, 1]) + float(m[2, 2])
    q = [0.0] * 4  # x y z w
    if tr > 0.0:
        t = math.sqrt(tr + 1.0)
        q[3] = 0.5 * t
        t = 0.5 / t
        q[0] = (float(m[2, 1]) - float(m[1, 2])) * t
        q[1] = (float(m[0, 2]) - float(m[2, 0])) * t
        q[2] = (float(m[1, 0]) - float(m[0, 1])) * t
    else:
        i = 0
        if m[1, 1] > m[0, 0]:
            i = 1
        if m[2, 2] > m[i, i]:
            i = 2
        j, k = (i + 1) % 3, (i + 2) % 3
        t = math.sqrt(float(m[i, i]) - float(m[j, j]) - float(m[k, k]) + 1.0)
        q[i] = 0.5 * t
        t = 0.5 / t
        q[3] = (float(m[k, j]) - float(m[j, k])) * t
        q[j] = (float(m[j, i]) + float(m[i, j])) * t
        q[k] = (float(m[k, i]) + float(m[i, k])) * t
    if q[3] < 0.0:
        q = [-c for c in q]
    n = math.sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3])
    return q[3] / n, q[0] / n, q[1] / n, q[2] / n


def project(R, t, pts):
    """pixels float64[M, 2] of cam_map(SE3Quat(R, t) * X) - intrinsics applied by the caller."""
    qw, qx, qy, qz = (np.float64(c) for c in quaternion_from_rotation(R))
    X = np.asarray(pts, np.float64)
    x0, x1, x2 = X[:, 0], X[:, 1], X[:, 2]
    ux = qy * x2 - qz * x1
    uy = qz * x0 - qx * x2
    uz = qx * x1 - qy * x0
    vx, vy, vz = ux + ux, uy + uy, uz + uz
    cx = qy * vz - qz * vy
    cy = qz * vx - qx * vz
    cz = qx * vy - qy * vx
    rx = (x0 + qw * vx) + cx
    ry = (x1 + qw * vy) + cy
    rz = (x2 + qw * vz) + cz
    tt = np.asarray(t, np.float64).ravel()
    return rx + tt[0], ry + tt[1], rz + tt[2]


def visible(R, t, see_vector, pts, normals, fx, fy, cx, cy, width, height, cos_max):
    with np.errstate(divide="ignore", invalid="ignore"):
        x, y, z = project(R, t, pts)
        u = (x / z) * np.float64(fx) + np.float64(cx)
        v = (y / z) * np.float64(fy) + np.float64(cy)
    N = np.asarray(normals, np.float64)
    s = np.asarray(see_vector, np.float64).ravel()
    dot = (s[0] * N[:, 0] + s[1] * N[:, 1]) + s[2] * N[:, 2]
    ok = (0.0 <= u) & (u < float(width)) & (0.0 <= v) & (v < float(height)) & (dot < np.float64(cos_max))
    return ok, np.stack([u, v], axis=1)


def track_local_map(store_desc, store_pt3d, store_normal, edges, frame_des, frame_kp, R, t, see_vector,
                    fx, fy, cx, cy, width, height, cos_max, cross_check=True, max_distance=30, strict=False,
                    k=1, ratio=None, window_radius=None):
    """Returns the dict of arrays MapStore.track produces (same names)."""
    edges = np.asarray(edges, np.int64)
    ok, pix = visible(R, t, see_vector, store_pt3d[edges], store_normal[edges], fx, fy, cx, cy, width, height, cos_max)
    vis = np.nonzero(ok)[0].astype(np.int32)
    feats = store_desc[edges[vis]]                       # :119 np.stack(feats)
    pts3d = store_pt3d[edges[vis]]
    vpix = pix[vis]
    kp = np.asarray(frame_kp, np.float64)
    mask = None
    if window_radius is not None:                        # fp32 window predicate, as the matcher evaluates it
        mask = orc.window_mask(kp.astype(np.float32), vpix.astype(np.float32), float(window_radius))
    if len(vis) == 0 or len(frame_des) == 0:
        e = np.zeros(0, np.int32)
        mq, mt, md = e, e, np.zeros(0, np.float32)
    else:
        mq, mt, md = orc.match(frame_des, feats, k=k, ratio=ratio, cross_check_=cross_check, mask=mask,
                               max_distance=max_distance, strict=strict)
    return {"visible_edges": vis, "visible_pixels": vpix, "inds_frame": mq, "inds": mt, "distance": md,
            "edges": vis[mt], "pts3d": pts3d[mt], "kp": kp[mq]}


def keyframe_vote(edge_kf, visible_edges, inds, inliers, top=100):
    """reference slam/tracking.py:142-154: ids_matching_kfs is the keyframe id of every VISIBLE edge (appended at :108
    for the edges that pass :103-104), and the vote is Counter(ids_matching_kfs[inds[inliers]]).most_common(top)."""
    from collections import Counter
    ids_matching_kfs = np.asarray(edge_kf)[np.asarray(visible_edges)]
    inds = np.asarray(inds)
    inliers = np.asarray(inliers, dtype=np.int64).flatten()
    return [(int(k), int(c)) for k, c in Counter(ids_matching_kfs[inds[inliers]].tolist()).most_common(top)]
