"""TEST INFRASTRUCTURE -- numpy restatement of the reference's two native geometry ops and of the torch glue around them.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline may import this module; the product path
(implicit_depth_b200/) never does.

Follows, line by line:
  ray_aabb_dense   /root/reference/src/extensions/ray_aabb/ray_aabb_cuda_kernel.cu:10-89 (kernel), :105-106 (zeros)
  ray_aabb_pairs   /root/reference/src/models/pipeline.py:277-285 (nonzero of mask[V,R] -> voxel-major pair list)
                   + :345-346 (dist[vox, ray] lookup)
  pcl_aabb_dense   /root/reference/src/extensions/pcl_aabb/pcl_aabb_cuda_kernel.cu:10-45, :61
  pcl_pair_label   /root/reference/src/models/pipeline.py:305-309 (pcl_mask[occ_vox_intersect_idx, miss_ray_intersect_idx])
  pcl_end_voxel    /root/reference/src/models/pipeline.py:939-944 (nonzero + torch_scatter.scatter(reduce='max', out=...))

Parity pinning: the reference ships no tests for these ops, so the restatement is pinned against the reference's OWN
kernels compiled for the CPU by oracle/build_ref.py (fixtures tests/golden/aabb_*.npz, made by
tests/golden/make_golden_aabb.py) -- bit-exact, including the double-precision reciprocal 1/(d + 1e-12).
All arithmetic below is IEEE single precision element-wise (numpy float32), except that reciprocal.
"""
from __future__ import annotations

import numpy as np


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def ray_inv_dir(ray_dir):
    """float(1 / (double(d) + 1e-12)) -- `1e-12` is a double literal in the kernel (ray_aabb_cuda_kernel.cu:32,48,67)."""
    return (1.0 / (_f32(ray_dir).astype(np.float64) + 1e-12)).astype(np.float32)


def ray_aabb_dense(ray_dir, voxel_bound, ray_bid, voxel_bid):
    """-> mask [V,R] int32, dist [V,R,2] float32 (zeros where no hit / different image)."""
    inv = ray_inv_dir(ray_dir)                                    # [R,3]
    vb = _f32(voxel_bound)                                        # [V,6]
    rb = np.asarray(ray_bid).astype(np.int64); xb = np.asarray(voxel_bid).astype(np.int64)
    V, R = vb.shape[0], inv.shape[0]
    same = xb[:, None] == rb[None, :]                             # :26
    pos = inv >= 0                                                # [R,3]   :33,49,68
    lo, hi = vb[:, None, 0:3], vb[:, None, 3:6]                   # [V,1,3]
    near = np.where(pos[None], lo, hi)                            # bound entered first
    far = np.where(pos[None], hi, lo)
    with np.errstate(invalid="ignore", over="ignore"):
        tmin = (near * inv[None]).astype(np.float32)              # [V,R,3] single-precision products :41-42
        tmax = (far * inv[None]).astype(np.float32)
        tmin_max, tmax_min = tmin[..., 0], tmax[..., 0]
        ok = same.copy()
        for a in (1, 2):
            ok &= ~((tmin_max > tmax[..., a]) | (tmax_min < tmin[..., a]))      # :60, :79
            tmin_max = np.fmax(tmin_max, tmin[..., a])            # fmaxf / fminf :62-63, :81-82
            tmax_min = np.fmin(tmax_min, tmax[..., a])
    mask = ok.astype(np.int32)
    dist = np.zeros((V, R, 2), np.float32)
    dist[..., 0] = np.where(ok, tmin_max, 0)
    dist[..., 1] = np.where(ok, tmax_min, 0)
    return mask, dist


def ray_aabb_pairs(ray_dir, voxel_bound, ray_bid, voxel_bid):
    """pipeline.py:283-285 + :345-346 -> (occ_vox_intersect_idx [P] i64, miss_ray_intersect_idx [P] i64, dist [P,2])."""
    mask, dist = ray_aabb_dense(ray_dir, voxel_bound, ray_bid, voxel_bid)
    vox, ray = np.nonzero(mask)                                   # row-major == torch.nonzero order (voxel-major)
    return vox.astype(np.int64), ray.astype(np.int64), dist[vox, ray]


def pcl_aabb_dense(pcl_pos, voxel_bound, pcl_bid, voxel_bid):
    """-> mask [V,N] int32; closed box test (pcl_aabb_cuda_kernel.cu:30-42)."""
    p = _f32(pcl_pos); vb = _f32(voxel_bound)
    pb = np.asarray(pcl_bid).astype(np.int64); xb = np.asarray(voxel_bid).astype(np.int64)
    ok = xb[:, None] == pb[None, :]
    for a in range(3):
        x = p[None, :, a]
        with np.errstate(invalid="ignore"):
            ok &= ~((x < vb[:, None, a]) | (x > vb[:, None, a + 3]))
    return ok.astype(np.int32)


def pcl_pair_label(pcl_pos, voxel_bound, pcl_bid, voxel_bid, pair_vox, pair_ray):
    """pipeline.py:305-309: pcl_label_float = pcl_mask[vox, ray].float()."""
    return pcl_aabb_dense(pcl_pos, voxel_bound, pcl_bid, voxel_bid)[pair_vox, pair_ray].astype(np.float32)


def pcl_end_voxel(pcl_pos, voxel_bound, pcl_bid, voxel_bid, end_voxel_id):
    """pipeline.py:939-944: end_voxel_id[n] = max(end_voxel_id[n], max{v : point n inside voxel v})."""
    mask = pcl_aabb_dense(pcl_pos, voxel_bound, pcl_bid, voxel_bid)
    out = np.array(end_voxel_id, dtype=np.int64, copy=True)
    vox, pt = np.nonzero(mask)
    np.maximum.at(out, pt, vox)
    return out
