"""CPU restatement of the SMPL layer -- TEST INFRASTRUCTURE (oracle/__init__.py).

**Parity unpinned**: the arithmetic lives in third-party ``smplx`` (unpinned in the
reference's requirements.txt:9; ``models/smpl.py:7`` constrains it to the
``ModelOutput``-era API, i.e. ~0.1.13 as pinned by SPIN), which is absent from
/root/reference and from this image.  This file restates smplx's published
``lbs`` / ``batch_rodrigues`` / ``batch_rigid_transform`` / ``VertexJointSelector``
and the reference's own wrapper (``models/smpl.py:27-46``), anchored on the
reference call sites demo.py:148, eval.py:146,172,193, smpl_regressor.py:160,176.

numpy, any float dtype (fp64 = ground truth, fp32 = "what the reference computes").
"""
import numpy as np

# reference constants.py:73-91 (JOINT_MAP) applied to constants.py:15-69 (JOINT_NAMES)
JOINT_MAP_49 = np.array([
    24, 12, 17, 19, 21, 16, 18, 20, 0, 2, 5, 8, 1, 4, 7, 25, 26, 27, 28, 29, 30, 31, 32, 33, 34,
    8, 5, 45, 46, 4, 7, 21, 19, 17, 16, 18, 20, 47, 48, 49, 50, 51, 52, 53, 24, 26, 25, 28, 27],
    dtype=np.int64)
# reference constants.py:97-100
J24_TO_J17 = [0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 18, 14, 16, 17]
J24_TO_J19 = J24_TO_J17[:14] + [19, 20, 21, 22, 23]
# reference constants.py:95-96
H36M_TO_J17 = [6, 5, 4, 1, 2, 3, 16, 15, 14, 11, 12, 13, 8, 10, 0, 7, 9]
H36M_TO_J14 = H36M_TO_J17[:14]


def batch_rodrigues_smplx(aa):
    """smplx.lbs.batch_rodrigues: angle=|v+1e-8|, K=skew(v/angle), R=I+sin*K+(1-cos)*K@K."""
    aa = np.asarray(aa)
    dt = aa.dtype
    angle = np.linalg.norm(aa + dt.type(1e-8), axis=1, keepdims=True)
    d = aa / angle
    c = np.cos(angle)[:, None]
    s = np.sin(angle)[:, None]
    rx, ry, rz = d[:, 0], d[:, 1], d[:, 2]
    z = np.zeros_like(rx)
    K = np.stack([z, -rz, ry, rz, z, -rx, -ry, rx, z], 1).reshape(-1, 3, 3)
    I = np.eye(3, dtype=dt)[None]
    return (I + s * K + (dt.type(1) - c) * (K @ K)).astype(dt)


def batch_rodrigues_quat(aa):
    """reference utils/geometry.py:9-45 (quaternion route)."""
    aa = np.asarray(aa)
    dt = aa.dtype
    l1 = np.linalg.norm(aa + dt.type(1e-8), axis=1, keepdims=True)
    n = aa / l1
    half = l1 * dt.type(0.5)
    q = np.concatenate([np.cos(half), np.sin(half) * n], 1)
    q = q / np.linalg.norm(q, axis=1, keepdims=True)
    w, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    w2, x2, y2, z2 = w * w, x * x, y * y, z * z
    wx, wy, wz, xy, xz, yz = w * x, w * y, w * z, x * y, x * z, y * z
    R = np.stack([w2 + x2 - y2 - z2, 2 * xy - 2 * wz, 2 * wy + 2 * xz,
                  2 * wz + 2 * xy, w2 - x2 + y2 - z2, 2 * yz - 2 * wx,
                  2 * xz - 2 * wy, 2 * wx + 2 * yz, w2 - x2 - y2 + z2], 1)
    return R.reshape(-1, 3, 3).astype(dt)


def rot6d_to_rotmat(x):
    """reference utils/geometry.py:47-61: columns (b1, b2, b3); F.normalize eps=1e-12."""
    x = np.asarray(x)
    dt = x.dtype
    x = x.reshape(-1, 3, 2)
    a1, a2 = x[:, :, 0], x[:, :, 1]
    nrm = lambda v: v / np.maximum(np.linalg.norm(v, axis=1, keepdims=True), dt.type(1e-12))
    b1 = nrm(a1)
    b2 = nrm(a2 - np.sum(b1 * a2, 1, keepdims=True) * b1)
    b3 = np.cross(b1, b2)
    return np.stack([b1, b2, b3], -1).astype(dt)


def lbs(model, betas, rot_mats, dtype=np.float64):
    """smplx.lbs.lbs with pose2rot=False.  rot_mats [B,24,3,3] -> (verts [B,6890,3],
    posed joints [B,24,3])."""
    f = lambda a: np.asarray(a, dtype=dtype)
    v_template, shapedirs, posedirs = f(model["v_template"]), f(model["shapedirs"]), f(model["posedirs"])
    J_regressor, W = f(model["J_regressor"]), f(model["lbs_weights"])
    parents = np.asarray(model["parents"])
    betas, R = f(betas), f(rot_mats)
    B = betas.shape[0]
    v_shaped = v_template[None] + np.einsum("bl,mkl->bmk", betas, shapedirs)
    J = np.einsum("bik,ji->bjk", v_shaped, J_regressor)
    ident = np.eye(3, dtype=dtype)
    pose_feature = (R[:, 1:] - ident).reshape(B, 207)
    v_posed = v_shaped + (pose_feature @ posedirs).reshape(B, -1, 3)
    # batch_rigid_transform
    rel = J.copy()
    rel[:, 1:] -= J[:, parents[1:]]
    G = np.zeros((B, 24, 4, 4), dtype=dtype)
    local = np.zeros((B, 24, 4, 4), dtype=dtype)
    local[:, :, :3, :3] = R
    local[:, :, :3, 3] = rel
    local[:, :, 3, 3] = 1
    G[:, 0] = local[:, 0]
    for i in range(1, 24):
        G[:, i] = G[:, parents[i]] @ local[:, i]
    posed_joints = G[:, :, :3, 3].copy()
    Jh = np.concatenate([J, np.zeros((B, 24, 1), dtype=dtype)], -1)[..., None]      # [B,24,4,1]
    A = G.copy()
    A[:, :, :, 3:] -= G @ Jh
    T = np.einsum("vj,bjk->bvk", W, A.reshape(B, 24, 16)).reshape(B, -1, 4, 4)
    vh = np.concatenate([v_posed, np.ones((B, v_posed.shape[1], 1), dtype=dtype)], -1)
    verts = np.einsum("bvij,bvj->bvi", T, vh)[..., :3]
    return verts, posed_joints


def smpl_forward(model, betas, body_pose, global_orient, pose2rot=True, dtype=np.float64):
    """The reference's SMPL wrapper (models/smpl.py:27-46) on top of smplx.SMPL.forward.

    Returns dict with vertices [B,6890,3], joints [B,49,3], smpl_joints [B,24,3],
    joints_J19 [B,19,3], full_pose (rot-mats if pose2rot=False else axis-angle [B,72])."""
    betas = np.asarray(betas, dtype=dtype)
    B = betas.shape[0]
    if pose2rot:
        full = np.concatenate([np.asarray(global_orient, dtype=dtype).reshape(B, -1),
                               np.asarray(body_pose, dtype=dtype).reshape(B, -1)], 1)
        R = batch_rodrigues_smplx(full.reshape(-1, 3)).reshape(B, 24, 3, 3)
    else:
        full = np.concatenate([np.asarray(global_orient, dtype=dtype).reshape(B, 1, 3, 3),
                               np.asarray(body_pose, dtype=dtype).reshape(B, 23, 3, 3)], 1)
        R = full
    verts, J = lbs(model, betas, R, dtype=dtype)
    joints45 = np.concatenate([J, verts[:, np.asarray(model["selected_verts"])]], 1)
    extra = np.einsum("jv,bvk->bjk", np.asarray(model["J_regressor_extra"], dtype=dtype), verts)
    joints54 = np.concatenate([joints45, extra], 1)
    joints = joints54[:, JOINT_MAP_49]
    out = {"vertices": verts, "joints": joints, "smpl_joints": joints45[:, :24],
           "joints_J19": joints[:, -24:][:, J24_TO_J19], "full_pose": full, "rot_mats": R}
    if "J_regressor_h36m" in model:
        out["joints_h36m"] = np.einsum("jv,bvk->bjk", np.asarray(model["J_regressor_h36m"], dtype=dtype), verts)
    return out


def mpjpe_h36m(joints_h36m, gt_j14):
    """reference eval.py:202-212: pelvis-centre J17, pick H36M_TO_J14, mean L2 per sample."""
    pelvis = joints_h36m[:, [0]]
    pred = (joints_h36m - pelvis)[:, H36M_TO_J14]
    return np.sqrt(((pred - gt_j14) ** 2).sum(-1)).mean(-1)


def perspective_projection(points, rotation, translation, focal_length, camera_center):
    """reference utils/geometry.py:63-91."""
    p = np.einsum("bij,bkj->bki", rotation, points) + translation[:, None]
    p = p / p[:, :, -1:]
    B = points.shape[0]
    K = np.zeros((B, 3, 3), dtype=points.dtype)
    K[:, 0, 0] = focal_length
    K[:, 1, 1] = focal_length
    K[:, 2, 2] = 1
    K[:, :-1, -1] = camera_center
    return np.einsum("bij,bkj->bki", K, p)[:, :, :-1]
