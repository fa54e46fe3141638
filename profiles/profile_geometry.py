"""The reference-named geometry ops at the encoder's four level sizes, 64 objects (the C2 batch), between
cudaProfilerStart/Stop (use with `ncu --profile-from-start off --set full`).  Launch order, per level k:
fps_kernel<..., 0> (gp_fps, the reference op), fps_kernel<..., 1> (gp_fps_chain, what the encoder calls: levels 2-4
take the FPS-order prefix when level 1 was tie-free), ball_query_kernel<1> (both radii), group_kernel<1> (gp_query_group, scale 1),
group_kernel<0> (gp_group, the bare grouping_operation on the same indices).

Prints one line per launch with the algorithmic bytes DESIGN.md section 4.1-4.3 states, in launch order, so that
make_readme.py can put them beside the ncu durations.
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from genpose2_b200 import pointnet2_utils as pu, synthetic  # noqa: E402
from genpose2_b200.pointnet2 import ClsMSG_CFG_Light as MSG_CFG  # noqa: E402

B = 64
pts, _ = synthetic.make_point_clouds(B, 1024, seed=0)
xyz0 = pts.contiguous().cuda()
chans = [0] + [sum(m[-1] for m in lv) for lv in MSG_CFG["MLPS"]]


def run(record):
    xyz, out, tie = xyz0, [], None
    for k, M in enumerate(MSG_CFG["NPOINTS"]):
        if M is None:
            break
        N, C = xyz.shape[1], chans[k]
        radii, ns = MSG_CFG["RADIUS"][k], MSG_CFG["NSAMPLE"][k]
        feat = torch.randn(B, C, N, device="cuda") if C else None
        pu.furthest_point_sample_gather(xyz, M)
        out.append(("fps", k, B * (12 * N + 4 * M + 12 * M)))
        idx, new_xyz, tie = pu.furthest_point_sample_chain(xyz, M, tie)
        out.append(("fps_chain", k, B * (12 * N + 4 * M + 12 * M)))
        bq = pu.ball_query2(radii, ns, xyz, new_xyz)
        out.append(("ball_query", k, B * (12 * N + 12 * M + 4 * M * (ns[0] + ns[1]))))
        pu.query_group(xyz, new_xyz, feat, bq[0])
        out.append(("query_group", k, B * (12 * N + 4 * C * N + 12 * M + 4 * M * ns[0] + 4 * (C + 3) * M * ns[0])))
        if feat is not None:
            pu.grouping_operation(feat, bq[0])
            out.append(("group", k, B * (4 * C * N + 4 * M * ns[0] + 4 * C * M * ns[0])))
        xyz = new_xyz
    if record:
        for o in out:
            print(json.dumps({"op": o[0], "level": o[1], "algorithmic_bytes": o[2]}))


for _ in range(3):
    run(False)
torch.cuda.synchronize()
torch.cuda.profiler.start()
run(True)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
