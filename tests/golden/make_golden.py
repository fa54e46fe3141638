"""Generates tests/golden/*.npz by running the REFERENCE ITSELF (imported read-only from
/root/reference through oracle/ref_shim.py) on seeded synthetic inputs, on CPU, in the build
container.  The fixtures travel; the reference does not.

    python tests/golden/make_golden.py

What runs is the reference's own code: PoseNet.pred_func / get_energy / pred_scale_func
(networks/posenet_agent.py), GFObjectPose.forward (networks/posenet.py), cond_ode_sampler /
cond_pc_sampler (networks/gf_algorithms/samplers.py) with scipy's solve_ivp, PoseScoreNet /
PoseEnergyNet, sort_poses_by_energy (networks/reward.py), average_quaternion_batch
(utils/misc.py), the vendored rotation conversions and sklearn's DBSCAN.  The only things
supplied from outside are (i) weights -- genpose2_b200.synthetic state dicts loaded with
load_state_dict, because the reference zero-initialises its output layers -- (ii) the encoder
features (the reference encoder is CUDA-only; `extract_pts_feature` is patched to return the
given features, as BASELINE.md section 3 prescribes) and (iii) the aggregation block's glue,
which lives in a runner that cannot be imported (runners/evaluation_single.py executes dataset
code at import) and is therefore called function-by-function in the order of
evaluation_single.py:179-215.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from genpose2_b200 import synthetic  # noqa: E402
from oracle import ref_shim  # noqa: E402
from oracle.ref_runner import reference_aggregate  # noqa: E402


def to_np(d):
    out = {}
    for k, v in d.items():
        if isinstance(v, torch.Tensor):
            v = v.detach().cpu().numpy()
        out[k] = np.asarray(v)
    return out


def main():
    ns = ref_shim.load()
    cfg = ns.cfg
    cfg.device = "cpu"
    cfg.sampler_mode = ["ode"]
    torch.manual_seed(0)
    torch.set_num_threads(8)

    score_sd = synthetic.random_gfobjectpose_state_dict(100)
    energy_sd = synthetic.random_gfobjectpose_state_dict(200)
    scale_sd = synthetic.random_scalenet_state_dict(300)

    injected = {}

    def patched_extract(self, data):
        return injected["feat"][self._gp_role]

    ns.posenet.GFObjectPose.extract_pts_feature = patched_extract

    cfg.agent_type = "score"
    score_agent = ns.posenet_agent.PoseNet(cfg)
    score_agent.net.load_state_dict(score_sd)
    score_agent.net._gp_role = "score"
    cfg.agent_type = "energy"
    energy_agent = ns.posenet_agent.PoseNet(cfg)
    energy_agent.net.load_state_dict(energy_sd)
    energy_agent.net._gp_role = "energy"
    cfg.agent_type = "scale"
    scale_agent = ns.posenet_agent.PoseNet(cfg)
    scale_agent.net.load_state_dict(scale_sd)
    cfg.agent_type = "score"

    meta = dict(score_seed=100, energy_seed=200, scale_seed=300)

    # ---- sampler-only cases: reference cond_ode_sampler on CPU --------------------------------
    def run_ode(name, B, R, T0, feat_seed, noise_seed, with_init=False, num_steps=None):
        g = torch.Generator().manual_seed(feat_seed)
        feat = torch.relu(torch.randn(B, 1024, generator=g))
        center = torch.randn(B, 3, generator=g) * 0.1 + torch.tensor([0.0, 0.0, 0.8])
        init = None
        if with_init:
            R0 = synthetic._random_rotations(np.random.default_rng(feat_seed), B)
            init = torch.zeros(B, 9)
            init[:, :3] = torch.from_numpy(R0[:, :, 0]).float()
            init[:, 3:6] = torch.from_numpy(R0[:, :, 1]).float()
            init[:, 6:] = torch.randn(B, 3, generator=g) * 0.02
        rep = lambda a: a.unsqueeze(1).repeat(1, R, 1).view(B * R, -1)
        data = {"pts": torch.zeros(B * R, 4, 3), "pts_feat": rep(feat), "pts_center": rep(center)}
        nfev = [0]
        net = score_agent.net
        orig_forward = net.pose_score_net.forward

        def counting(d):
            nfev[0] += 1
            return orig_forward(d)

        net.pose_score_net.forward = counting
        torch.manual_seed(noise_seed)
        xs, x = ns.samplers.cond_ode_sampler(
            score_model=net, data=data, prior=net.prior_fn, sde_coeff=net.sde_fn, atol=1e-5,
            rtol=1e-5, device="cpu", eps=net.sampling_eps, T=T0, num_steps=num_steps,
            pose_mode="rot_matrix", denoise=True, init_x=None if init is None else rep(init),
        )
        net.pose_score_net.forward = orig_forward
        # the reference against ITSELF with another sgemm summation order (1 thread instead of 8):
        # its own reproducibility bounds what any other implementation can be held to
        torch.set_num_threads(1)
        torch.manual_seed(noise_seed)
        _, x_1t = ns.samplers.cond_ode_sampler(
            score_model=net, data=dict(data), prior=net.prior_fn, sde_coeff=net.sde_fn, atol=1e-5,
            rtol=1e-5, device="cpu", eps=net.sampling_eps, T=T0, num_steps=num_steps,
            pose_mode="rot_matrix", denoise=True, init_x=None if init is None else rep(init),
        )
        torch.set_num_threads(8)
        torch.manual_seed(noise_seed)
        noise = net.prior_fn((B * R, 9), T=T0)
        out = dict(feat=feat, center=center, noise=noise, x=x, x_1thread=x_1t, xs_last=xs[:, -1], xs_first=xs[:, 0],
                   xs_mid=xs[:, xs.shape[1] // 2], S=xs.shape[1], nfev=nfev[0], B=B, R=R, T0=T0,
                   num_steps=-1 if num_steps is None else num_steps)
        if init is not None:
            out["init_x"] = init
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **to_np(out), **meta)
        print(name, "S", xs.shape[1], "nfev", nfev[0])

    run_ode("ode_c1_T1", 1, 50, 1.0, 11, 1)
    run_ode("ode_b4_T055", 4, 50, 0.55, 12, 2)
    run_ode("ode_track_T025", 2, 50, 0.25, 13, 3, with_init=True)
    run_ode("ode_b2_T055_steps20", 2, 10, 0.55, 14, 4, num_steps=20)

    # ---- PC sampler ---------------------------------------------------------------------------
    B, R, steps = 2, 8, 25
    g = torch.Generator().manual_seed(21)
    feat = torch.relu(torch.randn(B, 1024, generator=g))
    center = torch.randn(B, 3, generator=g) * 0.1 + torch.tensor([0.0, 0.0, 0.8])
    rep = lambda a: a.unsqueeze(1).repeat(1, R, 1).view(B * R, -1)
    data = {"pts": torch.zeros(B * R, 4, 3), "pts_feat": rep(feat), "pts_center": rep(center)}
    net = score_agent.net
    torch.manual_seed(5)
    xs, mean_x = ns.samplers.cond_pc_sampler(
        score_model=net, data=data, prior=net.prior_fn, sde_coeff=net.sde_fn, num_steps=steps,
        snr=0.16, device="cpu", eps=net.sampling_eps, pose_mode="rot_matrix", init_x=None)
    torch.manual_seed(5)
    init = net.prior_fn((B * R, 9))
    noises = torch.stack([torch.stack([torch.randn(B * R, 9), torch.randn(B * R, 9)]) for _ in range(steps)])
    np.savez_compressed(os.path.join(HERE, "pc_b2.npz"), **to_np(dict(
        feat=feat, center=center, init=init, noises=noises, xs=xs, mean_x=mean_x, B=B, R=R,
        steps=steps)), **meta)
    print("pc", xs.shape)

    # ---- PC sampler at the reference's default length (samplers.py:118: num_steps=500) -------------------
    # The 2 x 500 noise tensors are not stored: the test regenerates them from the same seed in the same call
    # order (torch's CPU generator is reproducible across machines for a given torch version).
    steps = 500
    torch.manual_seed(6)
    xs, mean_x = ns.samplers.cond_pc_sampler(
        score_model=net, data=data, prior=net.prior_fn, sde_coeff=net.sde_fn, num_steps=steps,
        snr=0.16, device="cpu", eps=net.sampling_eps, pose_mode="rot_matrix", init_x=None)
    torch.set_num_threads(1)   # the reference against itself with another sgemm summation order
    torch.manual_seed(6)
    _, mean_x_1t = ns.samplers.cond_pc_sampler(
        score_model=net, data=dict(data), prior=net.prior_fn, sde_coeff=net.sde_fn, num_steps=steps,
        snr=0.16, device="cpu", eps=net.sampling_eps, pose_mode="rot_matrix", init_x=None)
    torch.set_num_threads(8)
    keep = [0, 1, 10, 100, 250, 400, 499]
    np.savez_compressed(os.path.join(HERE, "pc_b2_500.npz"), **to_np(dict(
        feat=feat, center=center, noise_seed=6, xs_keep=xs[:, keep], keep=np.array(keep), mean_x=mean_x,
        mean_x_1thread=mean_x_1t, B=B, R=R, steps=steps)), **meta)
    print("pc500", xs.shape)

    # ---- full path through the agents: pred_func -> get_energy -> aggregate -> scale ----------
    def run_full(name, B, R, T0, seed, tracking=False):
        g = torch.Generator().manual_seed(seed)
        sfeat = torch.relu(torch.randn(B, 1024, generator=g))
        efeat = torch.relu(torch.randn(B, 1024, generator=g))
        injected["feat"] = {"score": sfeat, "energy": efeat}
        pts, center = synthetic.make_point_clouds(B, 1024, seed=seed)
        init = None
        if tracking:
            R0 = synthetic._random_rotations(np.random.default_rng(seed), B)
            init = torch.zeros(B, 9)
            init[:, :3] = torch.from_numpy(R0[:, :, 0]).float()
            init[:, 3:6] = torch.from_numpy(R0[:, :, 1]).float()
            init[:, 6:] = torch.randn(B, 3, generator=g) * 0.02
        data = {"pts": pts, "pts_center": center}
        torch.manual_seed(seed + 1)
        pred_pose, pred_q = score_agent.pred_func(data=data, repeat_num=R, T0=T0, init_x=init,
                                                  save_path=None)
        torch.manual_seed(seed + 1)
        noise = score_agent.net.prior_fn((B * R, 9), T=T0)
        energy = energy_agent.get_energy(data=data, pose_samples=pred_pose, T=1e-5, mode="test",
                                         extract_feature=True)
        agg, labels = reference_aggregate(ns, pred_pose, energy, R)
        data2 = dict(data)
        data2["pts_feat"] = sfeat
        data2["rgb_feat"] = None
        data2["axes"] = agg[:, :3, :3]
        axes, length = scale_agent.pred_scale_func(data2)
        out = dict(pts=pts, center=center, score_feat=sfeat, energy_feat=efeat, noise=noise,
                   pred_pose=pred_pose, pred_q=pred_q, energy=energy, aggregated_pose=agg,
                   labels=labels, length=length, B=B, R=R, T0=T0)
        if init is not None:
            out["init_x"] = init
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **to_np(out), **meta)
        print(name, "done")

    run_full("full_b3_T055", 3, 50, 0.55, 31)
    run_full("full_track_b2_T025", 2, 50, 0.25, 32, tracking=True)

    # ---- aggregation with real clusters (random-weight runs give none; SURVEY 8c trap 5) ------
    poses = synthetic.make_cluster_quaternion_poses(6, 50, seed=41)
    g = torch.Generator().manual_seed(41)
    energy = torch.randn(6, 50, 2, generator=g)
    # no exact energy ties: the order torch.sort(stable=False) gives equal keys is unspecified (and differs
    # between its CPU and CUDA kernels); the device path documents stable-descending and tests it separately
    agg, labels = reference_aggregate(ns, poses, energy, 50)
    sorted_pose, sorted_energy = ns.reward.sort_poses_by_energy(poses, energy)
    np.savez_compressed(os.path.join(HERE, "aggregate_clusters.npz"), **to_np(dict(
        poses=poses, energy=energy, aggregated_pose=agg, labels=labels, sorted_pose=sorted_pose,
        sorted_energy=sorted_energy)))
    print("aggregate clusters:", [int(l.max()) for l in labels])

    # ---- ScaleNet alone -----------------------------------------------------------------------
    B = 5
    g = torch.Generator().manual_seed(51)
    feat = torch.relu(torch.randn(B, 1024, generator=g))
    axes = torch.from_numpy(synthetic._random_rotations(np.random.default_rng(51), B)).float()
    _, length = scale_agent.pred_scale_func({"pts_feat": feat, "rgb_feat": None, "axes": axes})
    np.savez_compressed(os.path.join(HERE, "scalenet_b5.npz"), **to_np(dict(
        feat=feat, axes=axes, length=length)), **meta)

    # ---- energy alone at T=1e-5 ---------------------------------------------------------------
    B, R = 3, 50
    g = torch.Generator().manual_seed(61)
    efeat = torch.relu(torch.randn(B, 1024, generator=g))
    injected["feat"] = {"score": efeat, "energy": efeat}
    center = torch.randn(B, 3, generator=g) * 0.1 + torch.tensor([0.0, 0.0, 0.8])
    poses = synthetic.make_cluster_quaternion_poses(B, R, seed=61)
    poses[:, :, 6:] += center.unsqueeze(1).double() - torch.tensor([0.0, 0.0, 0.8]).double()
    energy = energy_agent.get_energy(data={"pts": torch.zeros(B, 4, 3), "pts_center": center},
                                     pose_samples=poses, T=1e-5, mode="test", extract_feature=True)
    np.savez_compressed(os.path.join(HERE, "energy_b3.npz"), **to_np(dict(
        feat=efeat, center=center, poses=poses, energy=energy)), **meta)
    print("energy", energy.shape)


if __name__ == "__main__":
    main()
