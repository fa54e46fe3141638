"""Drop-ins for `cond_ode_sampler` and `cond_pc_sampler`
(networks/gf_algorithms/samplers.py:180-258, 113-177): same signatures, same returned shapes and
dtypes (the ODE sampler returns float64 like scipy's state), same consumption of the global CPU
generator for the initial noise (`prior` is called exactly as the reference calls it).

What changes: the scipy `solve_ivp` host loop with a host<->device round trip per RHS evaluation
(samplers.py:204-234) is one persistent device kernel (gp_scorenet_ode); the 500-step PC loop is
one persistent kernel (gp_scorenet_pc).
"""
import threading

import numpy as np
import torch

from . import _lib
from .scorenet import _object_features

POSE_DIM = 9  # get_pose_dim("rot_matrix"), utils/genpose_utils.py:34-35
MAX_TRAJ = 512  # accepted-step slots recorded when the trajectory is requested (upper bound)
TRAJ_BYTES_BUDGET = 2 << 30   # the recording buffer [slots, N, 9] f64 is capped at this size (never below 64 slots)


def _traj_slots(batch_size):
    """Slots of the trajectory buffer: MAX_TRAJ for small batches, fewer when [slots, N, 9] f64 would exceed
    TRAJ_BYTES_BUDGET (the evaluation settings accept 9 - 62 steps; a longer trajectory raises)."""
    per_slot = max(1, batch_size) * POSE_DIM * 8
    return int(max(64, min(MAX_TRAJ, TRAJ_BYTES_BUDGET // per_slot)))

last_ode_stats = {}  # statistics of the most recent cond_ode_sampler call (nfev, accepted, ...)

# Failure of the device integrator (status -1: step size underflow, -2: attempt cap) must not pass silently, and
# checking it must not stall the stream: every call copies its statistics to pinned host memory asynchronously; the
# copies that have landed are inspected at the next sampler call / ode_stats() and a failed integration raises there.
# (In band, the kernel also turns x_out of a failed integration into NaN.)
_pending_stats = []


def _watch(stats):
    if torch.cuda.is_current_stream_capturing():
        return   # no host allocations / events inside a graph capture; the graph's owner watches the static stats
    host = torch.empty(stats.shape, dtype=stats.dtype, pin_memory=True)
    host.copy_(stats, non_blocking=True)
    ev = torch.cuda.Event()
    ev.record(torch.cuda.current_stream(stats.device))
    _pending_stats.append((ev, host))
    if len(_pending_stats) > 64:
        _pending_stats.pop(0)[0].synchronize()


def check_pending(wait=False):
    """Raise if an earlier device integration failed (only looks at statistics that have already reached the host
    unless wait=True)."""
    if torch.cuda.is_current_stream_capturing():
        return
    while _pending_stats and (wait or _pending_stats[0][0].query()):
        ev, host = _pending_stats.pop(0)
        ev.synchronize()
        status = int(host[_lib.STAT_STATUS])
        if status != 0:
            raise RuntimeError(f"device RK45 integration failed with status {status} "
                               "(-1: step size underflow, -2: attempt cap); its poses are NaN")


def _score_net(score_model):
    net = getattr(score_model, "pose_score_net", score_model)
    if not hasattr(net, "packed"):
        raise TypeError("score_model must be a genpose2_b200 GFObjectPose / PoseScoreNet (no fallback path)")
    return net


def _mlp_mode(score_model):
    return _score_net(score_model).mode_arg()


_staging = {}
_staging_lock = threading.Lock()


def _to_device_async(t, device):
    """Host tensor -> device without blocking the host: a pageable copy is synchronous AND ordered behind everything
    already enqueued on the stream, so the prior draw (made on the CPU with the CPU generator, like the reference)
    would stall the host until the encoder has finished.  Staged through a cached pinned buffer instead."""
    if t.is_cuda:
        return t.to(device)
    # one staging buffer per (device, host thread, shape, dtype): two threads / devices never share one, so a buffer
    # is only reused after the event of ITS previous copy; the cache is bounded (least recently created entry goes)
    key = (str(device), threading.get_ident(), tuple(t.shape), t.dtype)
    with _staging_lock:
        ent = _staging.get(key)
        if ent is None:
            if len(_staging) >= 16:
                _staging.pop(next(iter(_staging)))
            ent = _staging[key] = [torch.empty(t.shape, dtype=t.dtype, pin_memory=True), None]
    buf, ev = ent
    if ev is not None:
        ev.synchronize()   # the previous copy out of this buffer (long finished in practice)
    buf.copy_(t)
    out = buf.to(device, non_blocking=True)
    ev = torch.cuda.Event()
    ev.record(torch.cuda.current_stream(out.device))
    ent[1] = ev
    return out


def cond_ode_sampler(score_model, data, prior, sde_coeff, atol=1e-5, rtol=1e-5, device="cuda", eps=1e-5,
                     T=1.0, num_steps=None, pose_mode="quat_wxyz", denoise=True, init_x=None,
                     return_trajectory=True):
    """-> (xs [N, S, 9] f64, x [N, 9] f64).  num_steps=None: xs holds every accepted step (scipy's res.y);
    num_steps=n: scipy's dense output on t_eval = linspace(T, eps, n) (samplers.py:222-235).
    `return_trajectory=False` (not in the reference) skips recording and returns xs = x[:, None]."""
    if pose_mode != "rot_matrix":
        raise NotImplementedError("accelerated sampler supports pose_mode='rot_matrix' only")
    net = _score_net(score_model)
    check_pending()
    batch_size = data["pts"].shape[0]
    noise = _to_device_async(prior((batch_size, POSE_DIM), T=T), device)  # CPU generator, like samplers.py:197-201
    x0 = noise if init_x is None else init_x + noise
    dev = x0.device
    x0 = x0.to(torch.float64).contiguous()  # scipy casts y0 to float64
    feat, rpo = _object_features(data, batch_size)
    proj = net.project(feat.to(dev))
    center = _lib.check_cuda(data["pts_center"].to(torch.float32).contiguous(), "pts_center", torch.float32)

    x_out = torch.empty((batch_size, POSE_DIM), dtype=torch.float64, device=dev)
    stats = torch.zeros(_lib.GP_STAT_COUNT, dtype=torch.float64, device=dev)
    ws_bytes = _lib.load().gp_scorenet_ode_workspace_bytes(batch_size)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    if num_steps is not None:
        import numpy as np
        t_eval = torch.from_numpy(np.linspace(T, eps, int(num_steps))).to(dev)   # float64, as solve_ivp receives it
        dense = torch.empty((int(num_steps), batch_size, POSE_DIM), dtype=torch.float64, device=dev)
        _lib.call("gp_scorenet_ode_dense", _lib.ptr(net.packed()), _lib.ptr(proj), _lib.ptr(x0), _lib.ptr(center),
                  batch_size, rpo, float(T), float(eps), float(rtol), float(atol), 1 if denoise else 0,
                  _lib.ptr(x_out), _lib.ptr(t_eval), int(num_steps), _lib.ptr(dense), _lib.ptr(stats),
                  _lib.ptr(ws), ws_bytes, _mlp_mode(score_model), device=dev)
        xs = torch.empty((batch_size, int(num_steps), POSE_DIM), dtype=torch.float64, device=dev)
        _lib.call("gp_traj_finalize", _lib.ptr(dense), _lib.ptr(center), int(num_steps), batch_size, _lib.ptr(xs), device=dev)
        last_ode_stats.clear()
        last_ode_stats["device_stats"] = stats
        _watch(stats)
        return xs, x_out
    traj = None
    slots = _traj_slots(batch_size)
    if return_trajectory:
        traj = torch.empty((slots, batch_size, POSE_DIM), dtype=torch.float64, device=dev)
    _lib.call("gp_scorenet_ode", _lib.ptr(net.packed()), _lib.ptr(proj), _lib.ptr(x0), _lib.ptr(center),
              batch_size, rpo, float(T), float(eps), float(rtol), float(atol), 1 if denoise else 0,
              _lib.ptr(x_out), _lib.ptr(traj), slots if traj is not None else 0, _lib.ptr(stats),
              _lib.ptr(ws), ws_bytes, _mlp_mode(score_model), device=dev)
    if return_trajectory:
        st = stats.cpu()
        S = int(st[_lib.STAT_ACCEPTED].item()) + 1
        if S > slots:
            raise RuntimeError(f"trajectory of {S} accepted steps does not fit the {slots} recorded slots "
                               "(raise samplers.TRAJ_BYTES_BUDGET, or pass return_trajectory=False)")
        xs = torch.empty((batch_size, S, POSE_DIM), dtype=torch.float64, device=dev)
        _lib.call("gp_traj_finalize", _lib.ptr(traj), _lib.ptr(center), S, batch_size, _lib.ptr(xs), device=dev)
        _record_stats(st)
    else:
        xs = x_out.unsqueeze(1)
        last_ode_stats.clear()
        last_ode_stats["device_stats"] = stats
        _watch(stats)
    return xs, x_out


def _record_stats(st):
    last_ode_stats.clear()
    last_ode_stats.update(
        nfev=int(st[_lib.STAT_NFEV]), accepted=int(st[_lib.STAT_ACCEPTED]), rejected=int(st[_lib.STAT_REJECTED]),
        status=int(st[_lib.STAT_STATUS]), t_final=float(st[_lib.STAT_T_FINAL]),
        h_initial=float(st[_lib.STAT_H_INITIAL]), h_last=float(st[_lib.STAT_H_LAST]))
    if last_ode_stats["status"] != 0:
        # the reference ignores solve_ivp's status (samplers.py:226-236); here a failed integration is an error
        raise RuntimeError(f"device RK45 integration failed with status {last_ode_stats['status']} "
                           "(-1: step size underflow, -2: attempt cap); its poses are NaN")


def ode_stats():
    """Statistics of the last cond_ode_sampler call as a dict (synchronises if still on device)."""
    if "device_stats" in last_ode_stats:
        st = last_ode_stats["device_stats"].cpu()
        _pending_stats.clear()   # the call being read is the newest one; older ones were checked when it was issued
        _record_stats(st)
    return dict(last_ode_stats)


def cond_pc_sampler(score_model, data, prior, sde_coeff, num_steps=500, snr=0.16, device="cuda", eps=1e-5,
                    pose_mode="quat_wxyz", init_x=None, noise=None):
    """-> (xs [N, num_steps, 9] f32, mean_x [N, 9] f32).  `noise` (not in the reference):
    [num_steps, 2, N, 9] tensor replacing the per-step `torch.randn_like` draws; by default they are
    drawn with torch on `device` in the reference's call order."""
    if pose_mode != "rot_matrix":
        raise NotImplementedError("accelerated sampler supports pose_mode='rot_matrix' only")
    net = _score_net(score_model)
    batch_size = data["pts"].shape[0]
    x0 = _to_device_async(prior((batch_size, POSE_DIM)), device) if init_x is None else init_x
    dev = x0.device
    x0 = x0.to(torch.float32).contiguous()
    time_steps = torch.linspace(1.0, eps, num_steps, device=dev)
    if noise is None:
        noise = torch.stack([torch.stack([torch.randn_like(x0), torch.randn_like(x0)]) for _ in range(num_steps)])
    noise = _lib.check_cuda(noise.to(dev, torch.float32).contiguous(), "noise", torch.float32)
    feat, rpo = _object_features(data, batch_size)
    proj = net.project(feat.to(dev))
    center = _lib.check_cuda(data["pts_center"].to(torch.float32).contiguous(), "pts_center", torch.float32)
    xs = torch.empty((batch_size, num_steps, POSE_DIM), dtype=torch.float32, device=dev)
    mean_x = torch.empty((batch_size, POSE_DIM), dtype=torch.float32, device=dev)
    ws_bytes = _lib.load().gp_scorenet_pc_workspace_bytes(batch_size)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    _lib.call("gp_scorenet_pc", _lib.ptr(net.packed()), _lib.ptr(proj), _lib.ptr(x0), _lib.ptr(noise),
              _lib.ptr(center), _lib.ptr(time_steps), batch_size, rpo, int(num_steps), float(snr), _lib.ptr(xs),
              _lib.ptr(mean_x), _lib.ptr(ws), ws_bytes, _mlp_mode(score_model), device=dev)
    return xs, mean_x
