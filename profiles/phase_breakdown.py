import sys; sys.path.insert(0,'.')
import torch
from genpose2_b200 import samplers, synthetic, _lib
from genpose2_b200.pipeline import PosePipeline
for mode in ("bf16","fp32"):
    pipe = PosePipeline(device="cuda", mlp_mode=mode).load_synthetic_weights()
    B,R=64,50
    pts, center = synthetic.make_point_clouds(B, 1024, seed=0)
    net = pipe.score_agent.net
    feat = net({"pts": pts.cuda(), "pts_center": center.cuda()}, mode="pts_feature")
    N=B*R
    sd = {"pts": torch.empty(N, 0), "pts_center": center.cuda().unsqueeze(1).expand(B, R, 3).reshape(N, 3).contiguous(), "_gp_pts_feat_obj": feat, "_gp_rows_per_object": R}
    torch.manual_seed(0); noise = net.prior_fn((N, 9), T=0.55)
    for it in range(2):
        samplers.cond_ode_sampler(net, sd, lambda s, T=1.0: noise, net.sde_fn, device="cuda", T=0.55, pose_mode="rot_matrix", return_trajectory=False)
    st = samplers.last_ode_stats["device_stats"].cpu().tolist()
    nf = st[0]+1
    names=["fwd","l1","wait_d1","epi1","wait_heads","epi2"]
    print(mode, "nfev", nf, "kernel cycles", st[14], "per eval", st[14]/nf)
    for n,v in zip(names, st[8:14]): print(f"   {n:12s} {v/nf:10.0f} cycles/eval")
    for n,v in zip(["stage_tq","stage_input","stage_K","err+gridsync"], st[16:20]): print(f"   {n:12s} {v/nf:10.0f} cycles/eval")
    for n,v in zip(["x:combine","x:barrierA","x:scatter","x:barrierB","x:final"], st[20:25]): print(f"   {n:12s} {v/nf:10.0f} cycles/eval")
