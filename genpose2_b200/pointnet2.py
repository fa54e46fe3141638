"""Drop-in for `Pointnet2ClsMSG` (networks/pts_encoder/pointnet2.py:211-252) with the
`ClsMSG_CFG_Light` configuration (pointnet2.py:77-89) and its set-abstraction module
(pointnet2_utils/pointnet2/pointnet2_modules.py:19-74, 77-121).

The module tree reproduces the reference's state-dict keys exactly
(`SA_modules.{k}.mlps.{i}.layer{j}.conv.weight`, `...bn.bn.{weight,bias,running_mean,running_var}`)
so reference checkpoints load with `load_state_dict`.

FPS, gather, ball query and grouping run on the hand-written kernels (libgenpose_b200.so): one
fused FPS+gather launch, one two-radius ball query and one fused query-and-group per scale per
level, instead of the reference's FPS, 3 gathers / transposes, 2 ball queries and 4 groupings.
The geometry (FPS / ball-query indices) depends on xyz only, so `forward` can return it and accept
it back: the score and the energy encoder see the same cloud and share one geometry pass.

The per-scale SharedMLP (conv1x1 + BatchNorm(eval) + ReLU, pytorch_utils.py:5-33) followed by the
max-pool is "next" row f1 of SURVEY.md section 8 and still goes through torch (cuBLAS fp32
matmul with BatchNorm folded into the weights).
"""
from typing import List, Optional

import torch
import torch.nn as nn

from . import pointnet2_utils as pu

ClsMSG_CFG_Light = {
    "NPOINTS": [512, 256, 128, 64, None],
    "RADIUS": [[0.01, 0.02], [0.02, 0.04], [0.04, 0.08], [0.08, 0.16], [None, None]],
    "NSAMPLE": [[16, 32], [16, 32], [16, 32], [16, 32], [None, None]],
    "MLPS": [
        [[16, 16, 32], [32, 32, 64]],
        [[64, 64, 128], [64, 96, 128]],
        [[128, 196, 256], [128, 196, 256]],
        [[256, 256, 512], [256, 384, 512]],
        [[512, 512], [512, 512]],
    ],
    "DP_RATIO": 0.5,
}


class _BatchNorm2d(nn.Module):
    """pytorch_utils.py BatchNorm2d wrapper: child named `bn`."""

    def __init__(self, c):
        super().__init__()
        self.bn = nn.BatchNorm2d(c)


class _Conv2d(nn.Module):
    """pytorch_utils.py Conv2d(_ConvBase): children `conv` (1x1, no bias) and `bn`."""

    def __init__(self, cin, cout):
        super().__init__()
        self.conv = nn.Conv2d(cin, cout, kernel_size=(1, 1), bias=False)
        self.bn = _BatchNorm2d(cout)

    def folded(self):
        bn = self.bn.bn
        scale = bn.weight / torch.sqrt(bn.running_var + bn.eps)
        w = self.conv.weight[:, :, 0, 0] * scale[:, None]
        b = bn.bias - bn.running_mean * scale
        return w, b


class SharedMLP(nn.Module):
    """pytorch_utils.py:5-33: children `layer{i}`."""

    def __init__(self, spec: List[int]):
        super().__init__()
        self.n_layers = len(spec) - 1
        for i in range(self.n_layers):
            self.add_module(f"layer{i}", _Conv2d(spec[i], spec[i + 1]))

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """x (B, C, M, ns) -> (B, Cout, M): conv+BN+ReLU stack then max over ns."""
        B, C, M, ns = x.shape
        h = x.reshape(B, C, M * ns)
        for i in range(self.n_layers):
            w, b = getattr(self, f"layer{i}").folded()
            h = torch.relu_(torch.baddbmm(b[None, :, None], w.unsqueeze(0).expand(B, -1, -1), h))
        return h.view(B, -1, M, ns).amax(dim=3)


class PointnetSAModuleMSG(nn.Module):
    """pointnet2_modules.py:77-121 + forward :19-74 (max_pool, use_xyz=True, bn=True)."""

    def __init__(self, *, npoint, radii, nsamples, mlps):
        super().__init__()
        self.npoint, self.radii, self.nsamples = npoint, radii, nsamples
        self.groupers = nn.ModuleList(
            [pu.QueryAndGroup(r, n) if npoint is not None else pu.GroupAll() for r, n in zip(radii, nsamples)])
        self.mlps = nn.ModuleList([SharedMLP([spec[0] + 3] + spec[1:]) for spec in mlps])

    def forward(self, xyz, features=None, geometry=None):
        """xyz (B,N,3), features (B,C,N) -> new_xyz (B,npoint,3), new_features (B,sum Cout,npoint), geometry"""
        outs = []
        if self.npoint is not None:
            if geometry is None:
                idx, new_xyz = pu.furthest_point_sample_gather(xyz, self.npoint)
                bq = pu.ball_query2(self.radii, self.nsamples, xyz, new_xyz)
                geometry = (idx, new_xyz, bq)
            idx, new_xyz, bq = geometry
            for i in range(len(self.mlps)):
                grouped = pu.query_group(xyz, new_xyz, features, bq[i])
                outs.append(self.mlps[i](grouped))
        else:
            new_xyz = None
            for i in range(len(self.mlps)):
                outs.append(self.mlps[i](self.groupers[i](xyz, None, features)))
        return new_xyz, torch.cat(outs, dim=1), geometry


class Pointnet2ClsMSG(nn.Module):
    def __init__(self, input_channels=0):
        super().__init__()
        cfg = ClsMSG_CFG_Light
        self.SA_modules = nn.ModuleList()
        channel_in = input_channels
        for k in range(len(cfg["NPOINTS"])):
            mlps = [[channel_in] + list(m) for m in cfg["MLPS"][k]]
            self.SA_modules.append(PointnetSAModuleMSG(
                npoint=cfg["NPOINTS"][k], radii=cfg["RADIUS"][k], nsamples=cfg["NSAMPLE"][k], mlps=mlps))
            channel_in = sum(m[-1] for m in mlps)

    @staticmethod
    def _break_up_pc(pc):
        xyz = pc[..., 0:3].contiguous()
        features = pc[..., 3:].transpose(1, 2).contiguous() if pc.size(-1) > 3 else None
        return xyz, features

    def forward(self, pointcloud: torch.Tensor, geometry: Optional[list] = None, return_geometry=False):
        """pointcloud (B, N, 3 + C) -> (B, 1024).  `geometry`: per-level FPS / ball-query results of an
        earlier call on the same cloud (they depend on xyz only)."""
        xyz, features = self._break_up_pc(pointcloud)
        geo_out = []
        for k, sa in enumerate(self.SA_modules):
            g = None if geometry is None else geometry[k]
            xyz, features, g = sa(xyz, features, g)
            geo_out.append(g)
        out = features.squeeze(-1)
        return (out, geo_out) if return_geometry else out
