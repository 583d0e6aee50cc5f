"""Host-side mirror of the reference's ``models.pointnet.PointNet2Stage`` (reference src/models/pointnet.py:7-38).

Same constructor, same sub-module names (``point_lin1..4``, ``vox_lin1..2`` -> identical ``state_dict`` keys, reference
checkpoints load unchanged), same ``forward(inp_feat, vox2point_idx)``.  Inference (no autograd recording) runs the fused
kernels behind ``lidf_pointnet_forward`` (include/lidf_pointnet.h: the two 128 -> 128 per-point layers on the tcgen05 tensor
cores with split-bf16 operands, the small layers and the per-voxel max on the FMA pipe); while autograd is recording, the same maths runs
on torch ops (``scatter_reduce('amax')`` standing in for torch_scatter) so that training keeps working -- a native
backward does not exist.  The native path covers the shipped dimensions (6 / 32 / 128) on CUDA tensors and raises
otherwise: no silent fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from implicit_depth_b200.extensions.lidf_query.jit import MLP_IMPLS, _chk, lidf_query


class _PointNetParams(C.Structure):
    _fields_ = [(f"{name}_{part}", C.c_void_p) for name in ("point_lin1", "point_lin2", "vox_lin1", "point_lin3", "point_lin4", "vox_lin2")
                for part in ("w", "b")]


def pointnet_forward(state, inp_feat: torch.Tensor, vox2point_idx: torch.Tensor, n_vox: Optional[int] = None,
                     mlp_impl: str = "auto") -> torch.Tensor:
    """``state``: a PointNet2Stage module (ours or the reference's) or its state_dict.  -> occ_voxel_feat [V,128].
    ``mlp_impl``: engine of the two 128 -> 128 per-point layers, "auto" / "tc_bf16x3" (tcgen05, split bf16) or "simt_fp32"."""
    sd = state if isinstance(state, dict) else state.state_dict(keep_vars=True)
    lib = lidf_query.lib
    if not hasattr(lib, "_pn_ready"):
        lib.lidf_pointnet_workspace_bytes.restype = C.c_size_t
        lib.lidf_pointnet_workspace_bytes.argtypes = [C.c_int64, C.c_int64]
        lib.lidf_pointnet_forward_impl.restype = C.c_int
        lib.lidf_pointnet_forward_impl.argtypes = [C.POINTER(_PointNetParams), C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p,
                                                   C.c_void_p, C.c_size_t, C.c_int32, C.c_void_p]
        lib._pn_ready = True
    shapes = {"point_lin1": (32, 6), "point_lin2": (64, 32), "vox_lin1": (64, 64), "point_lin3": (128, 128),
              "point_lin4": (128, 128), "vox_lin2": (128, 128)}
    p = _PointNetParams()
    keep = []
    for name, shp in shapes.items():
        wt, bs = sd[name + ".weight"].detach(), sd[name + ".bias"].detach()
        if tuple(wt.shape) != shp:
            raise RuntimeError(f"pointnet: {name}.weight is {tuple(wt.shape)}, the native path supports pnet_in 6 / pnet_gf 32 / "
                               f"pnet_out 128 only (shipped YAMLs)")
        keep += [wt, bs]
        setattr(p, name + "_w", _chk(wt, name + ".weight", torch.float32))
        setattr(p, name + "_b", _chk(bs, name + ".bias", torch.float32))
    if inp_feat.dim() != 2 or inp_feat.shape[1] != 6 or tuple(vox2point_idx.shape) != (inp_feat.shape[0],):
        raise RuntimeError("inp_feat must be [N,6] and vox2point_idx [N]")
    dev = inp_feat.device
    N = int(inp_feat.shape[0])
    V = int(n_vox) if n_vox is not None else (int(vox2point_idx.max().item()) + 1 if N else 0)   # torch_scatter: max index + 1
    out = torch.empty(V, 128, dtype=torch.float32, device=dev)
    nbytes = int(lib.lidf_pointnet_workspace_bytes(N, V))
    ws = torch.empty(max(nbytes, 1), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        rc = lib.lidf_pointnet_forward_impl(C.byref(p), _chk(inp_feat, "inp_feat", torch.float32),
                                            _chk(vox2point_idx, "vox2point_idx", torch.int64), N, V, out.data_ptr(), ws.data_ptr(),
                                            nbytes, MLP_IMPLS[mlp_impl], C.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
    lidf_query._raise(rc, "lidf_pointnet_forward_impl")
    ws.record_stream(torch.cuda.current_stream(dev))
    return out


def _scatter_rows_max(src, index, n):
    out = torch.zeros(n, src.shape[1], dtype=src.dtype, device=src.device)      # post-ReLU sources: 0 is the identity
    return out.scatter_reduce(0, index.reshape(-1, 1).expand_as(src), src, reduce="amax", include_self=True)


class PointNet2Stage(nn.Module):
    mlp_impl = "auto"   # engine of the two 128 -> 128 per-point layers: "auto" / "tc_bf16x3" (tcgen05) or "simt_fp32"

    def __init__(self, input_channels=6, output_channels=256, gf_dim=64):
        super(PointNet2Stage, self).__init__()
        self.input_channels = input_channels
        self.gf_dim = gf_dim
        self.point_lin1 = nn.Linear(self.input_channels, self.gf_dim, bias=True)
        self.point_lin2 = nn.Linear(self.gf_dim, output_channels // 2, bias=True)
        self.vox_lin1 = nn.Linear(output_channels // 2, output_channels // 2, bias=True)
        self.point_lin3 = nn.Linear(output_channels, output_channels, bias=True)
        self.point_lin4 = nn.Linear(output_channels, output_channels, bias=True)
        self.vox_lin2 = nn.Linear(output_channels, output_channels, bias=True)

    def forward(self, inp_feat, vox2point_idx):
        needs_grad = torch.is_grad_enabled() and (inp_feat.requires_grad or any(p.requires_grad for p in self.parameters()))
        if not needs_grad:
            return pointnet_forward(self, inp_feat.float().contiguous(), vox2point_idx.long().contiguous(), mlp_impl=self.mlp_impl)
        n = int(vox2point_idx.max().item()) + 1
        point_feat2 = F.relu(self.point_lin2(F.relu(self.point_lin1(inp_feat))))
        occ_voxel_feat = F.relu(self.vox_lin1(_scatter_rows_max(point_feat2, vox2point_idx, n)))
        point_feat3 = torch.cat((occ_voxel_feat[vox2point_idx], point_feat2), -1)
        point_feat5 = F.relu(self.point_lin4(F.relu(self.point_lin3(point_feat3))))
        return F.relu(self.vox_lin2(_scatter_rows_max(point_feat5, vox2point_idx, n)))
