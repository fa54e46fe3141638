import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


def rot_from_6d(x):
    """[N,6] (first two columns) -> [N,3,3] float64 numpy, Gram-Schmidt."""
    x = np.asarray(x, dtype=np.float64)
    a1, a2 = x[:, :3], x[:, 3:6]
    b1 = a1 / np.linalg.norm(a1, axis=1, keepdims=True)
    b2 = a2 - (b1 * a2).sum(1, keepdims=True) * b1
    b2 /= np.linalg.norm(b2, axis=1, keepdims=True)
    b3 = np.cross(b1, b2)
    return np.stack([b1, b2, b3], axis=2)


def geodesic_mats(Ra, Rb):
    """rotation angle between Ra and Rb.  ||Ra - Rb||_F = 2 sqrt(2) |sin(theta / 2)|: unlike
    arccos((tr - 1) / 2) this stays accurate for tiny angles (arccos turns a 6e-8 float32 rounding
    of the trace into a 2e-4 rad 'error')."""
    d = np.linalg.norm((np.asarray(Ra, np.float64) - np.asarray(Rb, np.float64)).reshape(-1, 9), axis=1)
    return 2.0 * np.arcsin(np.clip(d / (2.0 * np.sqrt(2.0)), 0.0, 1.0))


def geodesic_6d(x, y):
    return geodesic_mats(rot_from_6d(x), rot_from_6d(y))


def pose_errors(x, y):
    """x, y [N,9] -> (max geodesic rad, max translation L2)."""
    x = np.asarray(x, dtype=np.float64).reshape(-1, 9)
    y = np.asarray(y, dtype=np.float64).reshape(-1, 9)
    return float(geodesic_6d(x[:, :6], y[:, :6]).max()), float(np.linalg.norm(x[:, 6:] - y[:, 6:], axis=1).max())


def ref_ext():
    """The reference's own CUDA extension built into oracle/_ref (None if it cannot be loaded)."""
    d = os.path.join(ROOT, "oracle", "_ref")
    if not os.path.exists(os.path.join(d, "pointnet2_cuda.so")):
        return None
    if d not in sys.path:
        sys.path.insert(0, d)
    try:
        import pointnet2_cuda
        return pointnet2_cuda
    except Exception:
        return None


def rep(a, R):
    return a.unsqueeze(1).repeat(1, R, *([1] * (a.dim() - 1))).view(a.shape[0] * R, *a.shape[1:])
