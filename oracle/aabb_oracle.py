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


# --------------------------------------------------------------------------------------------------------------------
# Voxelisation of the valid points (the producer of voxel_bound / occ_vox_bid / revidx):
#   batch_get_occupied_idx   /root/reference/src/utils/point_utils.py:12-76  (overlap = False, the only mode LIDF uses)
#   get_occ_vox_bound        /root/reference/src/models/pipeline.py:162-201
# Pinned by tests/golden/voxel_*.npz (made by tests/golden/make_golden_voxel.py from the reference's own code).
# Arithmetic as torch performs it on float32 tensors with python-float scalars (scalar rounded to fp32, IEEE ops).
# --------------------------------------------------------------------------------------------------------------------
XMIN = (-1.0, -1.0, 0.0)     # src/constants.py:15
XMAX = (1.0, 1.0, 2.0)       # src/constants.py:16


def grid_setup(res: int):
    """pipeline.py:167-173 -> (xmin [3] f32 incl. the half-voxel margin, part_size python float, grid dims rr [3])."""
    xmin = np.array(XMIN, np.float32); xmax = np.array(XMAX, np.float32)
    part = float(np.min(xmax - xmin)) / res
    xmin = (xmin - np.float32(0.5 * part)).astype(np.float32)
    xmax = (xmax + np.float32(0.5 * part)).astype(np.float32)
    rr = np.ceil((xmax - xmin) / np.float32(part)).astype(np.int64)            # point_utils.py:25-27
    return xmin, part, rr


def batch_get_occupied_idx(valid_xyz, valid_bid, xmin, part_size: float, rr):
    """point_utils.py:12-76 -> (occ_bid_global_coord [V,4] i64 sorted unique, revidx [Nv], valid_v_pid [Nv],
    valid_v_rel_coord [Nv,3] f32)."""
    crop = np.float32(part_size)
    v = (_f32(valid_xyz) - _f32(xmin)[None]).astype(np.float32)                 # :23
    coord = np.floor(v / crop).astype(np.int64)                                 # :43 (shift is zero without overlap)
    center = (coord.astype(np.float32) * crop + np.float32(0.5 * part_size)).astype(np.float32)   # :50
    rel = (v - center).astype(np.float32)                                       # :51
    ok = np.ones(v.shape[0], bool)
    for i in range(3):                                                          # :59-61
        ok &= (coord[:, i] >= 0) & (coord[:, i] < rr[i])
    pid = np.nonzero(ok)[0].astype(np.int64)
    rows = np.concatenate((np.asarray(valid_bid).astype(np.int64)[ok, None], coord[ok]), 1)      # :70
    occ, revidx = np.unique(rows, axis=0, return_inverse=True)                  # :73 (sorted, lexicographic)
    return occ.astype(np.int64), revidx.reshape(-1).astype(np.int64), pid, rel[ok]


def get_occ_vox_bound(valid_xyz, valid_bid, res: int):
    """pipeline.py:162-201 -> dict with the data_dict entries it writes."""
    xmin, part, rr = grid_setup(res)
    occ, revidx, pid, rel = batch_get_occupied_idx(valid_xyz, valid_bid, xmin, part, rr)
    bound_min = (xmin[None] + occ[:, 1:].astype(np.float32) * np.float32(part)).astype(np.float32)   # :186
    bound_max = (bound_min + np.float32(part)).astype(np.float32)                                    # :187
    return dict(xmin=xmin, part_size=part, revidx=revidx, valid_v_pid=pid, valid_v_rel_coord=rel, occ_vox_bid=occ[:, 0],
                occ_vox_global_coord=occ[:, 1:], voxel_bound=np.concatenate((bound_min, bound_max), 1), grid=rr)
