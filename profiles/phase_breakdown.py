"""In-kernel cycle counters of the integrator (stats[8..24] of gp_scorenet_ode, CTA 0), per RHS evaluation:
the cluster evaluator at C2 (64 objects x 50 = 25 tiles) and the one-CTA-per-tile evaluator at the per-GPU share of C5
(1024 objects x 50 = 400 tiles; CTA 0 owns 3 tiles, so its per-evaluation numbers cover 3 tile evaluations)."""
import sys; sys.path.insert(0, '.')
import torch
from genpose2_b200 import samplers
from genpose2_b200.pipeline import PosePipeline
R, T0 = 50, 0.55
for mode in ("bf16", "fp32"):
    pipe = PosePipeline(device="cuda", mlp_mode=mode).load_synthetic_weights()
    net = pipe.score_agent.net
    for B, shape in ((64, "cluster"), (1024, "solo")):
        g = torch.Generator().manual_seed(5)
        feat = torch.relu(torch.randn(B, 1024, generator=g)).cuda()
        center = (torch.randn(B, 3, generator=g) * 0.1 + torch.tensor([0.0, 0.0, 0.8])).cuda()
        N = B * R
        sd = {"pts": torch.empty(N, 0), "pts_center": center.unsqueeze(1).expand(B, R, 3).reshape(N, 3).contiguous(),
              "_gp_pts_feat_obj": feat, "_gp_rows_per_object": R}
        torch.manual_seed(0); noise = net.prior_fn((N, 9), T=T0).cuda()
        net.pose_score_net.eval_shape = shape
        ts = []
        for it in range(3):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            samplers.cond_ode_sampler(net, sd, lambda sh, T=1.0: noise, net.sde_fn, device="cuda", T=T0, pose_mode="rot_matrix", return_trajectory=False)
            e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
        st = samplers.last_ode_stats["device_stats"].cpu().tolist()
        nf = st[0] + 1
        tiles = (N + 127) // 128
        print(f"{mode} {shape} evaluator, {B} objects x {R} = {tiles} tiles: {sorted(ts)[1]:.3f} ms, nfev {nf:.0f}, kernel cycles {st[14]:.0f}, per eval {st[14] / nf:.0f}")
        for n, v in zip(["fwd", "l1", "wait_d1", "epi1", "wait_heads", "epi2"], st[8:14]):
            print(f"   {n:12s} {v / nf:10.0f} cycles/eval")
        for n, v in zip(["stage_tq", "stage_input", "stage_K", "err+gridsync"], st[16:20]):
            print(f"   {n:12s} {v / nf:10.0f} cycles/eval")
        if shape == "cluster":
            for n, v in zip(["x:combine", "x:barrierA", "x:scatter", "x:barrierB", "x:final"], st[20:25]):
                print(f"   {n:12s} {v / nf:10.0f} cycles/eval")
            print(f"   {'mma:wait_w':12s} {st[15] / nf:10.0f} cycles/eval   (MMA issuer waiting for weight chunks of the ring)")
        else:
            print(f"   {'tail':12s} {st[20] / nf:10.0f} cycles/eval")
