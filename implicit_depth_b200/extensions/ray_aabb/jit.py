"""``ray_aabb`` -- drop-in for the reference's native extension of the same name, plus the compact pair list.

Reference: ``from extensions.ray_aabb.jit import ray_aabb`` (src/models/pipeline.py:18) and
``mask, dist = ray_aabb.forward(miss_ray_dir, voxel_bound, miss_bid.int(), occ_vox_bid.int())`` (pipeline.py:277-278;
binding src/extensions/ray_aabb/ray_aabb_cuda.cpp:20-37).  ``forward`` keeps that signature and returns the same two
dense tensors, bit-identical.  ``pairs`` returns what the reference derives from them -- ``torch.nonzero(mask)`` and
``dist[vox, ray]`` (pipeline.py:283-285, :345-346) -- without materialising the 12*V*R-byte slab.

Bound with ctypes to the C ABI of include/lidf_aabb.h (same library as lidf_query).  CUDA tensors only, no fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Tuple

import torch

from implicit_depth_b200.extensions.lidf_query.jit import _chk, lidf_query


class _RayAabb:
    @staticmethod
    def _stream(dev):
        return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)

    def _args(self, ray_dir, voxel_bound, ray_bid, voxel_bid):
        if ray_dir.dim() != 2 or ray_dir.shape[1] != 3 or voxel_bound.dim() != 2 or voxel_bound.shape[1] != 6:
            raise RuntimeError("ray_dir must be [R,3] and voxel_bound [V,6]")
        R, V = int(ray_dir.shape[0]), int(voxel_bound.shape[0])
        if tuple(ray_bid.shape) != (R,) or tuple(voxel_bid.shape) != (V,):
            raise RuntimeError("ray_bid must be [R] and voxel_bid [V]")
        ptrs = (_chk(ray_dir, "ray_dir", torch.float32), _chk(voxel_bound, "voxel_bound", torch.float32),
                _chk(ray_bid, "ray_bid", torch.int32), _chk(voxel_bid, "voxel_bid", torch.int32))
        return R, V, ptrs

    def forward(self, ray_dir, voxel_bound, ray_bid, voxel_bid) -> List[torch.Tensor]:
        """-> [mask [V,R] int32, dist [V,R,2] float32], as ray_aabb_cuda_forward (ray_aabb_cuda_kernel.cu:91-126)."""
        lib = lidf_query.lib
        R, V, ptrs = self._args(ray_dir, voxel_bound, ray_bid, voxel_bid)
        dev = ray_dir.device
        mask = torch.empty(V, R, dtype=torch.int32, device=dev)
        dist = torch.empty(V, R, 2, dtype=torch.float32, device=dev)
        nbytes = int(lib.lidf_ray_aabb_workspace_bytes(R, V))
        if nbytes == 0:
            raise RuntimeError("ray_aabb: problem too large (V * ceil(R/1024) must stay below 2^31)")
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            rc = lib.lidf_ray_aabb_forward(*ptrs, R, V, mask.data_ptr(), dist.data_ptr(), ws.data_ptr(), nbytes, self._stream(dev))
        lidf_query._raise(rc, "lidf_ray_aabb_forward")
        ws.record_stream(torch.cuda.current_stream(dev))
        return [mask, dist]

    def pairs(self, ray_dir, voxel_bound, ray_bid, voxel_bid, order: str = "nonzero", return_ray_start: bool = False):
        """-> (occ_vox_intersect_idx [P] int64, miss_ray_intersect_idx [P] int64, intersect_dist [P,2] float32).
        ``order="nonzero"``: torch.nonzero order (voxel, then ray), the reference's.  ``order="ray"``: the same pairs sorted
        by ray, then voxel -- what ``lidf_query.forward(..., pairs_ray_major=True)`` consumes without regrouping;
        ``return_ray_start`` appends the CSR offsets [R+1] int32.  Synchronises once to learn P, like torch.nonzero."""
        if order == "ray":
            return self._pairs_ray_major(ray_dir, voxel_bound, ray_bid, voxel_bid, return_ray_start)
        if order != "nonzero" or return_ray_start:
            raise ValueError("order must be 'nonzero' or 'ray' (ray_start comes with order='ray')")
        lib = lidf_query.lib
        R, V, ptrs = self._args(ray_dir, voxel_bound, ray_bid, voxel_bid)
        dev = ray_dir.device
        nbytes = int(lib.lidf_ray_aabb_workspace_bytes(R, V))
        if nbytes == 0:
            raise RuntimeError("ray_aabb: problem too large (V * ceil(R/1024) must stay below 2^31)")
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        n = C.c_int64(0)
        with torch.cuda.device(dev):
            st = self._stream(dev)
            rc = lib.lidf_ray_aabb_pairs_count(*ptrs, R, V, ws.data_ptr(), nbytes, C.byref(n), st)
            lidf_query._raise(rc, "lidf_ray_aabb_pairs_count")
            P = int(n.value)
            vox = torch.empty(P, dtype=torch.int64, device=dev)
            ray = torch.empty(P, dtype=torch.int64, device=dev)
            dist = torch.empty(P, 2, dtype=torch.float32, device=dev)
            rc = lib.lidf_ray_aabb_pairs_fill(*ptrs, R, V, ws.data_ptr(), nbytes, P, vox.data_ptr(), ray.data_ptr(),
                                              dist.data_ptr(), st)
            lidf_query._raise(rc, "lidf_ray_aabb_pairs_fill")
        ws.record_stream(torch.cuda.current_stream(dev))
        return vox, ray, dist

    def _pairs_ray_major(self, ray_dir, voxel_bound, ray_bid, voxel_bid, return_ray_start):
        lib = lidf_query.lib
        R, V, ptrs = self._args(ray_dir, voxel_bound, ray_bid, voxel_bid)
        dev = ray_dir.device
        nbytes = int(lib.lidf_ray_aabb_ray_major_workspace_bytes(R, V))
        if nbytes == 0:
            raise RuntimeError("ray_aabb: problem too large (R and V must stay below 2^31)")
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        n = C.c_int64(0)
        with torch.cuda.device(dev):
            st = self._stream(dev)
            rc = lib.lidf_ray_aabb_pairs_ray_major_count(*ptrs, R, V, ws.data_ptr(), nbytes, C.byref(n), st)
            lidf_query._raise(rc, "lidf_ray_aabb_pairs_ray_major_count")
            P = int(n.value)
            vox = torch.empty(P, dtype=torch.int64, device=dev)
            ray = torch.empty(P, dtype=torch.int64, device=dev)
            dist = torch.empty(P, 2, dtype=torch.float32, device=dev)
            ray_start = torch.empty(R + 1, dtype=torch.int32, device=dev) if return_ray_start else None
            rc = lib.lidf_ray_aabb_pairs_ray_major_fill(*ptrs, R, V, ws.data_ptr(), nbytes, P, vox.data_ptr(), ray.data_ptr(),
                                                        dist.data_ptr(), ray_start.data_ptr() if return_ray_start else None, st)
            lidf_query._raise(rc, "lidf_ray_aabb_pairs_ray_major_fill")
        ws.record_stream(torch.cuda.current_stream(dev))
        return (vox, ray, dist, ray_start) if return_ray_start else (vox, ray, dist)


ray_aabb = _RayAabb()
