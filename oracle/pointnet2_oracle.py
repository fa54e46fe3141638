"""TEST INFRASTRUCTURE ONLY -- ctypes front-end of oracle/pointnet2_oracle.c (the CPU
restatement of sampling_gpu.cu / ball_query_gpu.cu / group_points_gpu.cu)."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle_pointnet2.so")
_lib = None


def build():
    src = os.path.join(_HERE, "pointnet2_oracle.c")
    if not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "liboracle_pointnet2.so"])
    return _SO


def _load():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
    return _lib


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def fps_block_size(n):
    return _load().gp_oracle_fps_block_size(int(n))


def furthest_point_sample(xyz, npoint):
    xyz = _f32(xyz)
    B, N, _ = xyz.shape
    idx = np.zeros((B, npoint), dtype=np.int32)
    rc = _load().gp_oracle_fps(_p(xyz), B, N, int(npoint), _p(idx))
    assert rc == 0
    return idx


def gather_operation(features, idx):
    features, idx = _f32(features), _i32(idx)
    B, C, N = features.shape
    m = idx.shape[1]
    out = np.empty((B, C, m), dtype=np.float32)
    _load().gp_oracle_gather(_p(features), _p(idx), B, C, N, m, _p(out))
    return out


def ball_query(radius, nsample, xyz, new_xyz):
    xyz, new_xyz = _f32(xyz), _f32(new_xyz)
    B, N, _ = xyz.shape
    M = new_xyz.shape[1]
    idx = np.zeros((B, M, nsample), dtype=np.int32)
    _load().gp_oracle_ball_query(
        _p(new_xyz), _p(xyz), B, N, M, ctypes.c_float(radius), int(nsample), _p(idx)
    )
    return idx


def grouping_operation(features, idx):
    features, idx = _f32(features), _i32(idx)
    B, C, N = features.shape
    _, M, ns = idx.shape
    out = np.empty((B, C, M, ns), dtype=np.float32)
    _load().gp_oracle_group(_p(features), _p(idx), B, C, N, M, ns, _p(out))
    return out
