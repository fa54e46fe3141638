"""Drop-in for `ScaleNet` (networks/scalenet.py:12-56): same constructor, parameter names and
`forward(data)` contract; encode_axes (utils/genpose_utils.py:8-18) and the four Linear layers run
in one kernel (gp_scalenet)."""
import ctypes

import torch
import torch.nn as nn

from . import _lib
from .scorenet import zero_module


class ScaleNet(nn.Module):
    def __init__(self, pts_dim, dino_dim=0, embedding_dim=180):
        super().__init__()
        if dino_dim or embedding_dim != 180 or pts_dim != 1024:
            raise NotImplementedError("accelerated ScaleNet supports pts_dim=1024, dino_dim=0, embedding_dim=180")
        self.pts_dim, self.dino_dim, self.embedding_dim = pts_dim, dino_dim, embedding_dim
        self.act = nn.ReLU(True)
        self.axes_encoder = nn.Sequential(nn.Linear(embedding_dim, 256), self.act, nn.Linear(256, 256), self.act)
        self.fusion_tail_length = nn.Sequential(
            nn.Linear(pts_dim + dino_dim + 256, 256), self.act, zero_module(nn.Linear(256, 3)))

    def _params(self):
        ts = [self.axes_encoder[0].weight, self.axes_encoder[0].bias, self.axes_encoder[2].weight,
              self.axes_encoder[2].bias, self.fusion_tail_length[0].weight, self.fusion_tail_length[0].bias,
              self.fusion_tail_length[2].weight, self.fusion_tail_length[2].bias]
        keep = [t.detach().to(torch.float32).contiguous() for t in ts]
        p = _lib.ScaleNetParams()
        (p.axes_w0, p.axes_b0, p.axes_w1, p.axes_b1, p.tail_w0, p.tail_b0, p.tail_w1, p.tail_b1) = [
            t.data_ptr() for t in keep]
        return p, keep

    def forward(self, data):
        """data{'pts_feat' [bs,1024], 'axes' [bs,3,3]} -> length [bs,3]"""
        axes = data["axes"]
        feat = _lib.check_cuda(data["pts_feat"].to(torch.float32).contiguous(), "pts_feat", torch.float32)
        axes = axes.to(feat.device, torch.float32)
        if axes.stride(-1) != 1:
            axes = axes.contiguous()
        if axes.dim() != 3 or axes.shape[1] != 3 or axes.shape[2] != 3:
            raise ValueError("axes must be [bs,3,3]")
        B = feat.shape[0]
        p, keep = self._params()
        out = torch.empty((B, 3), dtype=torch.float32, device=feat.device)
        # strided view of a [bs,4,4] pose is accepted as is (no copy)
        _lib.call("gp_scalenet", ctypes.byref(p), _lib.ptr(axes), int(axes.stride(0)) if B > 1 else 16,
                  int(axes.stride(1)), _lib.ptr(feat), B, _lib.ptr(out), device=feat.device)
        del keep
        return out

    def loss_fn(self, pred_len, gt_len):
        return torch.mean((pred_len - gt_len) ** 2) * 10000
