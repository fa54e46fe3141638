"""The three rotation conversions the hot path uses (utils/transforms/rotation_conversions.py:
rotation_6d_to_matrix :556-577, matrix_to_quaternion :102-161, quaternion_to_matrix :41-70) plus
`get_rot_matrix` / `normalize_rotation` (utils/misc.py:121-160, 327-344) for pose_mode=rot_matrix.
Plain tensor math on whatever device/dtype the input has; used only for API-surface outputs
(`pred_pose_q_wxyz`) -- the aggregation kernel has its own fused float64 versions."""
import torch
import torch.nn.functional as F


def rotation_6d_to_matrix(d6):
    a1, a2 = d6[..., :3], d6[..., 3:]
    b1 = F.normalize(a1, dim=-1)
    b2 = F.normalize(a2 - (b1 * a2).sum(-1, keepdim=True) * b1, dim=-1)
    return torch.stack((b1, b2, torch.cross(b1, b2, dim=-1)), dim=-2)


def get_rot_matrix(batch_pose, pose_mode="rot_matrix"):
    if pose_mode != "rot_matrix":
        raise NotImplementedError("only pose_mode='rot_matrix' is on the accelerated path")
    return rotation_6d_to_matrix(batch_pose).permute(0, 2, 1)


def matrix_to_quaternion(matrix):
    m = matrix.reshape(matrix.shape[:-2] + (9,))
    m00, m01, m02, m10, m11, m12, m20, m21, m22 = torch.unbind(m, dim=-1)
    q_abs = torch.sqrt(torch.clamp_min(torch.stack(
        [1.0 + m00 + m11 + m22, 1.0 + m00 - m11 - m22, 1.0 - m00 + m11 - m22, 1.0 - m00 - m11 + m22], dim=-1), 0.0))
    cand = torch.stack([
        torch.stack([q_abs[..., 0] ** 2, m21 - m12, m02 - m20, m10 - m01], dim=-1),
        torch.stack([m21 - m12, q_abs[..., 1] ** 2, m10 + m01, m02 + m20], dim=-1),
        torch.stack([m02 - m20, m10 + m01, q_abs[..., 2] ** 2, m12 + m21], dim=-1),
        torch.stack([m10 - m01, m20 + m02, m21 + m12, q_abs[..., 3] ** 2], dim=-1)], dim=-2)
    cand = cand / (2.0 * q_abs[..., None].clamp_min(0.1))
    best = q_abs.argmax(dim=-1)
    return torch.gather(cand, -2, best[..., None, None].expand(best.shape + (1, 4))).squeeze(-2)


def quaternion_to_matrix(q):
    r, i, j, k = torch.unbind(q, -1)
    two_s = 2.0 / (q * q).sum(-1)
    o = torch.stack((
        1 - two_s * (j * j + k * k), two_s * (i * j - k * r), two_s * (i * k + j * r),
        two_s * (i * j + k * r), 1 - two_s * (i * i + k * k), two_s * (j * k - i * r),
        two_s * (i * k - j * r), two_s * (j * k + i * r), 1 - two_s * (i * i + j * j)), -1)
    return o.reshape(q.shape[:-1] + (3, 3))


def matrix_to_rotation_6d_cols(matrix):
    """get_pose_representation(rot, 'rot_matrix') (utils/misc.py:163-190): first two COLUMNS."""
    return torch.cat([matrix[..., :, 0], matrix[..., :, 1]], dim=-1)
