"""Energy ranking + pose aggregation on the device.

`sort_poses_by_energy` mirrors networks/reward.py:131-155; `aggregate_pose` mirrors the block that
is copy-pasted in the reference runners (runners/evaluation_single.py:179-215,
evaluation_tracking.py:146-183, infer.py:158-194): keep the top `retain_ratio` hypotheses by energy
(rotation and translation ranked independently), convert to quaternions, eigen-average, DBSCAN on
the pairwise-distance rows, re-average the largest cluster, mean translation -> [bs,4,4].
One kernel launch (gp_aggregate) instead of a host loop with a D2H copy and an sklearn call per object.
"""
import torch

from . import _lib


MAX_HYPOTHESES = 64          # AG_MAXR of csrc/aggregate.cu: hypotheses per object the kernel ranks in one warp
MAX_RETAIN_CLUSTERED = 32    # AG_MAXK: retained hypotheses when DBSCAN runs (neighbourhoods are 32-bit masks)


def _run(poses, energy, retain, clustering, eps, min_samples, want_sorted=False, want_labels=False):
    """Limits of the kernel that the reference's host loop does not have: at most 64 hypotheses per object, at most 32
    retained ones when clustering (64 without).  The evaluation scripts use 50 / 20.  Outside them this raises
    NotImplementedError instead of silently doing something else."""
    R = poses.shape[1]
    if R > MAX_HYPOTHESES:
        raise NotImplementedError(f"aggregation of {R} hypotheses per object: the kernel handles at most {MAX_HYPOTHESES}")
    if retain > (MAX_RETAIN_CLUSTERED if clustering else MAX_HYPOTHESES):
        raise NotImplementedError(f"retain_num={retain}: the kernel keeps at most {MAX_RETAIN_CLUSTERED} hypotheses with "
                                  f"clustering ({MAX_HYPOTHESES} without)")
    if clustering and min_samples < 1:
        # sklearn.cluster.DBSCAN(min_samples=0) raises InvalidParameterError (the reference would crash here)
        raise ValueError(f"clustering_minpts * retain_num = {min_samples} < 1: DBSCAN needs min_samples >= 1")
    poses = _lib.check_cuda(poses.to(torch.float64).contiguous(), "poses", torch.float64)
    energy = _lib.check_cuda(energy.to(poses.device, torch.float32).contiguous(), "energy", torch.float32)
    B, R, D = poses.shape
    if D != 9 or energy.shape != (B, R, 2):
        raise ValueError("poses must be [bs,R,9] and energy [bs,R,2]")
    out = torch.empty((B, 4, 4), dtype=torch.float32, device=poses.device)
    labels = torch.empty((B, retain), dtype=torch.int32, device=poses.device) if want_labels else None
    sorted_p = torch.empty_like(poses) if want_sorted else None
    _lib.call("gp_aggregate", _lib.ptr(poses), _lib.ptr(energy), B, R, int(retain), 1 if clustering else 0,
              float(eps), int(min_samples), _lib.ptr(out), _lib.ptr(labels), _lib.ptr(sorted_p),
              device=poses.device)
    return out, labels, sorted_p


def sort_poses_by_energy(poses, energy):
    """reward.py:131-155 -> (sorted_poses [bs,R,9], sorted_energy [bs,R,2]).  Equal energies are ranked
    stable-descending (torch.sort(descending=True) leaves their order unspecified, and its CPU and CUDA kernels differ)."""
    R = poses.shape[1]
    _, _, sorted_p = _run(poses, energy, min(R, 32), False, 0.0, 1, want_sorted=True)
    sorted_energy = torch.sort(energy.to(poses.device), descending=True, dim=1)[0]
    return sorted_p.to(poses.dtype), sorted_energy


def aggregate_pose(pred_pose, pred_energy, cfg=None, *, eval_repeat_num=None, retain_ratio=0.4, clustering=1,
                   clustering_eps=0.05, clustering_minpts=0.1667, return_labels=False):
    """evaluation_single.py:179-215.  pred_pose [bs,R,9] (f64), pred_energy [bs,R,2] -> [bs,4,4] f32
    (on the device; the reference builds it on the CPU)."""
    if cfg is not None:
        eval_repeat_num = cfg.eval_repeat_num
        retain_ratio, clustering = cfg.retain_ratio, cfg.clustering
        clustering_eps, clustering_minpts = cfg.clustering_eps, cfg.clustering_minpts
    if eval_repeat_num is None:
        eval_repeat_num = pred_pose.shape[1]
    retain_num = int(eval_repeat_num * retain_ratio)
    min_samples = int(clustering_minpts * retain_num)
    out, labels, _ = _run(pred_pose, pred_energy, retain_num, clustering, clustering_eps, min_samples,
                          want_labels=return_labels)
    return (out, labels) if return_labels else out
