"""One step of the C2 workload between cudaProfilerStart/Stop (use with `ncu --profile-from-start off`).

    python profiles/profile_step.py [--what step|sampler] [--objects 64] [--mlp_mode fp32]
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from genpose2_b200 import samplers, synthetic  # noqa: E402
from genpose2_b200.pipeline import PosePipeline  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--what", default="step")
ap.add_argument("--objects", type=int, default=64)
ap.add_argument("--mlp_mode", default="fp32")
ap.add_argument("--T0", type=float, default=0.55)
args = ap.parse_args()

B, R = args.objects, 50
pipe = PosePipeline(device="cuda", mlp_mode=args.mlp_mode).load_synthetic_weights()
pts, center = synthetic.make_point_clouds(B, 1024, seed=0)
data = lambda: {"pts": pts.cuda(), "pts_center": center.cuda()}
torch.manual_seed(0)
if args.what == "step":
    fn = lambda: pipe(data(), repeat_num=R, T0=args.T0)
else:
    net = pipe.score_agent.net
    feat = net(data(), mode="pts_feature")
    N = B * R
    sd = {"pts": torch.empty(N, 0), "pts_center": center.cuda().unsqueeze(1).expand(B, R, 3).reshape(N, 3).contiguous(),
          "_gp_pts_feat_obj": feat, "_gp_rows_per_object": R}
    noise = net.prior_fn((N, 9), T=args.T0)
    fn = lambda: samplers.cond_ode_sampler(net, sd, lambda s, T=1.0: noise, net.sde_fn, device="cuda", T=args.T0,
                                           pose_mode="rot_matrix", return_trajectory=False)
for _ in range(3):
    fn()
torch.cuda.synchronize()
torch.cuda.profiler.start()
fn()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done")
