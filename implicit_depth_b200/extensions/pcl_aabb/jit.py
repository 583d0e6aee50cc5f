"""``pcl_aabb`` -- drop-in for the reference's native extension of the same name, plus its two fused uses.

Reference: ``from extensions.pcl_aabb.jit import pcl_aabb`` (src/models/pipeline.py:19) and
``pcl_aabb.forward(pos, voxel_bound, bid.int(), occ_vox_bid.int())`` -> mask [V,N] (pipeline.py:305,316,939; binding
src/extensions/pcl_aabb/pcl_aabb_cuda.cpp:20-37).  ``forward`` keeps that signature.  ``pair_label`` and
``end_voxel`` give what the reference computes FROM the dense mask at its two hot call sites, without the mask:
``pcl_mask[vox, ray].float()`` (pipeline.py:305-309) and ``scatter(max)`` of the hit voxel ids into ``end_voxel_id``
(pipeline.py:939-944).

Bound with ctypes to the C ABI of include/lidf_aabb.h.  CUDA tensors only, no fallback.
"""
from __future__ import annotations

import ctypes as C

import torch

from implicit_depth_b200.extensions.lidf_query.jit import _chk, lidf_query


class _PclAabb:
    @staticmethod
    def _stream(dev):
        return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)

    def _args(self, pcl_pos, voxel_bound, pcl_bid, voxel_bid):
        if pcl_pos.dim() != 2 or pcl_pos.shape[1] != 3 or voxel_bound.dim() != 2 or voxel_bound.shape[1] != 6:
            raise RuntimeError("pcl_pos must be [N,3] and voxel_bound [V,6]")
        N, V = int(pcl_pos.shape[0]), int(voxel_bound.shape[0])
        if tuple(pcl_bid.shape) != (N,) or tuple(voxel_bid.shape) != (V,):
            raise RuntimeError("pcl_bid must be [N] and voxel_bid [V]")
        ptrs = (_chk(pcl_pos, "pcl_pos", torch.float32), _chk(voxel_bound, "voxel_bound", torch.float32),
                _chk(pcl_bid, "pcl_bid", torch.int32), _chk(voxel_bid, "voxel_bid", torch.int32))
        return N, V, ptrs

    def forward(self, pcl_pos, voxel_bound, pcl_bid, voxel_bid) -> torch.Tensor:
        """-> mask [V,N] int32, as pcl_aabb_cuda_forward (pcl_aabb_cuda_kernel.cu:49-80)."""
        N, V, ptrs = self._args(pcl_pos, voxel_bound, pcl_bid, voxel_bid)
        dev = pcl_pos.device
        mask = torch.empty(V, N, dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            rc = lidf_query.lib.lidf_pcl_aabb_forward(*ptrs, N, V, mask.data_ptr(), self._stream(dev))
        lidf_query._raise(rc, "lidf_pcl_aabb_forward")
        return mask

    def pair_label(self, pcl_pos, voxel_bound, pcl_bid, voxel_bid, pair_vox, pair_ray) -> torch.Tensor:
        """pcl_label_float [P] = pcl_mask[pair_vox, pair_ray].float() (pipeline.py:305-309)."""
        N, V, ptrs = self._args(pcl_pos, voxel_bound, pcl_bid, voxel_bid)
        dev = pcl_pos.device
        P = int(pair_vox.shape[0])
        label = torch.empty(P, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            rc = lidf_query.lib.lidf_pcl_aabb_pair_label(*ptrs, N, V, _chk(pair_vox, "pair_vox", torch.int64),
                                                         _chk(pair_ray, "pair_ray", torch.int64), P, label.data_ptr(),
                                                         self._stream(dev))
        lidf_query._raise(rc, "lidf_pcl_aabb_pair_label")
        return label

    def end_voxel(self, pcl_pos, voxel_bound, pcl_bid, voxel_bid, end_voxel_id) -> torch.Tensor:
        """In place: end_voxel_id[n] = max(end_voxel_id[n], largest voxel id containing point n) (pipeline.py:939-944)."""
        N, V, ptrs = self._args(pcl_pos, voxel_bound, pcl_bid, voxel_bid)
        if tuple(end_voxel_id.shape) != (N,):
            raise RuntimeError("end_voxel_id must be [N]")
        dev = pcl_pos.device
        with torch.cuda.device(dev):
            rc = lidf_query.lib.lidf_pcl_aabb_end_voxel(*ptrs, N, V, _chk(end_voxel_id, "end_voxel_id", torch.int64),
                                                        self._stream(dev))
        lidf_query._raise(rc, "lidf_pcl_aabb_end_voxel")
        return end_voxel_id


pcl_aabb = _PclAabb()
