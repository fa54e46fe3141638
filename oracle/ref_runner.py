"""TEST INFRASTRUCTURE ONLY -- drives the UNMODIFIED reference (imported through oracle/ref_shim.py from
/root/reference in the build container, or from its byte-for-byte copy oracle/_ref/refpkg on the GPU box)
through the per-object pose-generation path, in the order its runner executes it
(runners/evaluation_single.py:98-222):

    score_agent.pred_func -> energy_agent.get_energy(T=1e-5) -> aggregation block -> scale_agent.pred_scale_func

Everything that computes is the reference's own code: PoseNet / GFObjectPose / Pointnet2ClsMSG (with the
reference's CUDA extension when a GPU is present), cond_ode_sampler with scipy's solve_ivp,
PoseScoreNet / PoseEnergyNet, sort_poses_by_energy, average_quaternion_batch, sklearn's DBSCAN, ScaleNet.
Supplied from outside: the weights (the reference zero-initialises its output layers, SURVEY.md 8(c) trap
1) and, optionally, injected encoder features (the reference encoder has no CPU implementation).  The
aggregation block lives in a runner that executes dataset code at import, so its statements are issued
here function by function in the order of evaluation_single.py:179-215.

Used by tests/golden/make_golden.py (fixtures), the -m gpu parity tests and bench.py's reference legs.
Never imported by genpose2_b200/.
"""
import copy

import numpy as np
import torch

from . import ref_shim


def reference_aggregate(ns, pred_pose, pred_energy, repeat_num, retain_ratio=0.4, eps=0.05, minpts=0.1667):
    """evaluation_single.py:179-215, called function by function (see module docstring)."""
    from sklearn.cluster import DBSCAN

    sorted_pose, _ = ns.reward.sort_poses_by_energy(pred_pose, pred_energy)
    bs = pred_pose.shape[0]
    retain_num = int(repeat_num * retain_ratio)
    good_pose = sorted_pose[:, :retain_num, :]
    rot_matrix = ns.misc.get_rot_matrix(good_pose[:, :, :-3].reshape(bs * retain_num, -1), "rot_matrix")
    quat_wxyz = ns.rotconv.matrix_to_quaternion(rot_matrix).reshape(bs, retain_num, -1)
    agg_q = ns.misc.average_quaternion_batch(quat_wxyz)
    all_labels = []
    for j in range(bs):
        pd = 1 - torch.sum(quat_wxyz[j].unsqueeze(0) * quat_wxyz[j].unsqueeze(1), dim=2) ** 2
        labels = DBSCAN(eps=eps, min_samples=int(minpts * retain_num)).fit(pd.cpu().cpu().numpy()).labels_
        all_labels.append(labels)
        if np.any(labels >= 0):
            bins = np.bincount(labels[labels >= 0])
            best = np.argmax(bins)
            agg_q[j] = ns.misc.average_quaternion_batch(quat_wxyz[j, labels == best].unsqueeze(0))[0]
    agg_t = torch.mean(good_pose[:, :, -3:], dim=1)
    out = torch.zeros(bs, 4, 4)
    out[:, 3, 3] = 1
    out[:, :3, :3] = ns.rotconv.quaternion_to_matrix(agg_q)
    out[:, :3, 3] = agg_t
    return out, np.stack(all_labels)


class ReferenceAgents:
    """The reference's three agents (score / energy / scale) on `device`, with the given state dicts."""

    def __init__(self, score_sd, energy_sd, scale_sd, device="cpu", inject_features=False):
        self.ns = ns = ref_shim.load()
        cfg = copy.copy(ns.cfg)
        cfg.device = device
        cfg.sampler_mode = ["ode"]
        self.device = device
        self.injected = {}
        self.inject_features = inject_features

        def agent(kind, sd):
            c = copy.copy(cfg)
            c.agent_type = kind
            a = ns.posenet_agent.PoseNet(c)
            a.net.load_state_dict(sd)
            if inject_features and kind != "scale":
                # instance attribute shadows GFObjectPose.extract_pts_feature (posenet.py:127): forward(mode="pts_feature")
                # calls self.extract_pts_feature(data).  The class itself stays untouched.
                a.net.extract_pts_feature = lambda data, _k=kind: self.injected[_k]
            return a

        self.score_agent = agent("score", score_sd)
        self.energy_agent = agent("energy", energy_sd)
        self.scale_agent = agent("scale", scale_sd)
        self.nfev = 0
        net = self.score_agent.net.pose_score_net
        orig = net.forward

        def counting(d, *a, **k):
            self.nfev += 1
            return orig(d, *a, **k)

        net.forward = counting

    def encoder(self, which="score"):
        return getattr(self, which + "_agent").net.pts_encoder

    @torch.no_grad()
    def full(self, pts, center, R, T0, init_x=None, noise_seed=None, score_feat=None, energy_feat=None,
             stages=None):
        """-> dict(pred_pose, pred_q, energy, aggregated_pose, labels, length, score_feat, nfev).  `noise_seed`
        reseeds the global CPU generator right before pred_func (the prior draws from it, sde.py:34).
        `stages`: optional dict that receives wall-clock seconds per stage."""
        import time

        dev = self.device
        if score_feat is not None:
            self.injected["score"], self.injected["energy"] = score_feat.to(dev), energy_feat.to(dev)
        data = {"pts": pts.to(dev), "pts_center": center.to(dev)}
        init = None if init_x is None else init_x.to(dev)
        if noise_seed is not None:
            torch.manual_seed(noise_seed)
        self.nfev = 0

        def tick():
            if str(dev).startswith("cuda"):
                torch.cuda.synchronize()
            return time.perf_counter()

        t0 = tick()
        pred_pose, pred_q = self.score_agent.pred_func(data=data, repeat_num=R, T0=T0, init_x=init, save_path=None)
        nfev = self.nfev
        t1 = tick()
        energy = self.energy_agent.get_energy(data=data, pose_samples=pred_pose, T=1e-5, mode="test",
                                              extract_feature=True)
        t2 = tick()
        agg, labels = reference_aggregate(self.ns, pred_pose, energy, R)
        t3 = tick()
        data2 = dict(data)
        data2["pts_feat"] = data["pts_feat"]
        data2["rgb_feat"] = None
        data2["axes"] = agg[:, :3, :3].to(dev)
        _, length = self.scale_agent.pred_scale_func(data2)
        t4 = tick()
        if stages is not None:
            stages.update(pred_func=t1 - t0, get_energy=t2 - t1, aggregate=t3 - t2, scale=t4 - t3)
        return dict(pred_pose=pred_pose, pred_q=pred_q, energy=energy, aggregated_pose=agg, labels=labels,
                    length=length, score_feat=data["pts_feat"], nfev=nfev)
