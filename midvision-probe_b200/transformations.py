"""Mirror of the reference's ``evals/utils/transformations.py`` (the three functions the correspondence callers
use).  The rigid transform itself runs inside ``mv_k3_score`` on the matching path; these torch forms serve the
callers' host-side glue (3x3 / 3x4 matrices, a handful of points).
"""
import torch

__all__ = ["transform_points_Rt", "so3_relative_angle", "so3_rotation_angle"]


def transform_points_Rt(points, viewpoint, inverse=False):
    """points (..., n, 3) moved by viewpoint (..., 3|4, 4): p R^T + t, or (p - t) R when inverse.
    evals/utils/transformations.py:27-36."""
    R = viewpoint[..., :3, :3]
    t = viewpoint[..., None, :3, 3]
    return (points - t) @ R if inverse else points @ R.transpose(-2, -1) + t


def so3_rotation_angle(R, eps=1e-4):
    """rotation angle (radians) of a batch of 3x3 rotations from the trace.  evals/utils/transformations.py:47-63."""
    if R.dim() != 3 or R.shape[1:] != (3, 3):
        raise ValueError("Input has to be a batch of 3x3 Tensors.")
    tr = R[:, 0, 0] + R[:, 1, 1] + R[:, 2, 2]
    if ((tr < -1.0 - eps) + (tr > 3.0 + eps)).any():
        raise ValueError("A matrix has trace outside valid range [-1-eps,3+eps].")
    return torch.acos(((tr - 1.0) * 0.5).clamp(min=-1, max=1))


def so3_relative_angle(R1, R2, eps=1e-4):
    """angle of R1 R2^T.  evals/utils/transformations.py:39-44."""
    return so3_rotation_angle(torch.bmm(R1, R2.permute(0, 2, 1)), eps=eps)
