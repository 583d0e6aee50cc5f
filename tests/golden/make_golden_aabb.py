"""Golden vectors for the ray_aabb / pcl_aabb rows, produced by the REFERENCE's own kernels (build container only).

    python tests/golden/make_golden_aabb.py        # needs /root/reference; writes tests/golden/aabb_*.npz

oracle/build_ref.py compiles the two reference CUDA kernels for the CPU (their bodies are plain C) into
oracle/_ref/libaabb_ref.so; this script drives them on seeded inputs and stores inputs + outputs.  The torch glue the
reference wraps around the kernels is executed here with the same torch calls:
  torch.nonzero(mask) -> pair list, dist[vox, ray]                      (pipeline.py:283-285, :345-346)
  pcl_mask[vox, ray].float()                                            (pipeline.py:305-309)
  scatter(pred_occ_mask_idx[:,0], pred_occ_mask_idx[:,1], out=end_voxel_id, reduce='max')   (pipeline.py:939-944;
      torch_scatter is absent here: restated as the sequential max loop of its 2.0.x CPU kernel)
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)

from implicit_depth_b200.synthetic import XMIN, grid_part_size, make_rays  # noqa: E402
from oracle import build_ref  # noqa: E402


def grid_voxels(B, V_img, rng, res=8):
    """Distinct cells of the 9^3 grid per image (pipeline.py:167-189)."""
    part = grid_part_size(res); n = res + 1
    bounds, bid = [], []
    for b in range(B):
        cells = rng.permutation(n ** 3)[:V_img]
        c = np.stack((cells // (n * n), (cells // n) % n, cells % n), -1).astype(np.float32)
        lo = (np.array(XMIN, np.float32) - np.float32(0.5 * part)) + c * np.float32(part)
        bounds.append(np.concatenate((lo, lo + np.float32(part)), 1)); bid += [b] * V_img
    return np.concatenate(bounds).astype(np.float32), np.array(bid, np.int32)


def case_grid(seed, B, H, W, V_img):
    rng = np.random.default_rng(seed)
    miss_bid, _, ray_dir = make_rays(B, H, W, "cpu")          # includes the dir.x == 0 column and dir.y == 0 row
    vb, xb = grid_voxels(B, V_img, rng)
    return ray_dir.numpy(), vb, miss_bid.numpy().astype(np.int32), xb


def case_edge(seed):
    """Hand-made: axis-aligned and negative directions, +-0 components, rays grazing faces/edges/corners exactly,
    interleaved (unsorted) image ids, an image id with no rays, a degenerate zero-thickness voxel."""
    rng = np.random.default_rng(seed)
    dirs = [(0, 0, 1), (1, 0, 0), (0, 1, 0), (0, 0, -1), (-1, 0, 0), (0, -1, 0), (-0.0, 0.0, 1.0), (0.0, -0.0, 1.0),
            (1, 1, 1), (-1, 1, 1), (1, -1, 1), (-1, -1, 1), (1, 1, -1), (0.25, 0, 1), (0, 0.25, 1), (0.125, 0.125, 1),
            (0.5, 0, 1), (0, 0.5, 1), (0.5, 0.5, 1), (-0.5, 0.5, 1), (1e-12, 0, 1), (-1e-12, 0, 1), (1e-20, 1e-20, 1),
            (3e-7, -3e-7, 1)]
    d = np.array(dirs, np.float64)
    d = np.concatenate((d, rng.normal(size=(40, 3))))
    d[24:] /= np.linalg.norm(d[24:], axis=1, keepdims=True)        # the hand-made ones stay un-normalised on purpose
    ray_dir = d.astype(np.float32)
    R = ray_dir.shape[0]
    ray_bid = (np.arange(R) % 3).astype(np.int32)                  # interleaved ids 0,1,2
    boxes = [(-0.125, -0.125, 0.5, 0.125, 0.125, 0.75), (0.125, -0.125, 0.5, 0.375, 0.125, 0.75),
             (-0.125, 0.125, 0.5, 0.125, 0.375, 0.75), (0.25, 0.25, 1.0, 0.5, 0.5, 1.25), (-0.5, -0.5, 1.0, -0.25, -0.25, 1.25),
             (0, 0, 0, 0.25, 0.25, 0.25), (-0.25, -0.25, -0.25, 0, 0, 0), (0.5, -0.125, 0.875, 0.75, 0.125, 1.125),
             (0.0, 0.0, 1.0, 0.0, 0.25, 1.25), (-1.125, -1.125, -0.125, 1.125, 1.125, 2.125)]
    vb = np.array(boxes * 3, np.float32)
    voxel_bid = np.repeat(np.array([2, 0, 1], np.int32), len(boxes))      # unsorted; 3 copies, one per image id
    vb = np.concatenate((vb, vb[:2]))
    voxel_bid = np.concatenate((voxel_bid, np.array([7, 7], np.int32)))   # image 7 has no rays
    return ray_dir, vb, ray_bid, voxel_bid


def pcl_points(seed, vb, xb, n):
    """Points inside / outside / exactly on faces shared by neighbouring voxels (closed test -> inside both)."""
    rng = np.random.default_rng(seed)
    pick = rng.integers(0, vb.shape[0], size=n)
    u = rng.random((n, 3)).astype(np.float32)
    snap = rng.random((n, 3)) < 0.25
    u = np.where(snap, np.round(u), u)                                    # 25 % of the coordinates sit on a face
    p = vb[pick, :3] + u * (vb[pick, 3:] - vb[pick, :3])
    far = rng.random(n) < 0.2
    p[far] += rng.normal(size=(int(far.sum()), 3)).astype(np.float32)
    pb = xb[pick].copy()
    wrong = rng.random(n) < 0.1
    pb[wrong] = (pb[wrong] + 1) % (xb.max() + 1)
    return p.astype(np.float32), pb.astype(np.int32)


def main():
    lib = build_ref.build(force=True)
    ref = build_ref.load()
    print("reference kernels compiled for the CPU:", lib)
    cases = {"aabb_grid_2x12x16": case_grid(11, 2, 12, 16, 40), "aabb_grid_1x9x11": case_grid(12, 1, 9, 11, 120),
             "aabb_edge": case_edge(13)}
    for name, (ray_dir, vb, ray_bid, voxel_bid) in cases.items():
        mask, dist = ref.ray_aabb(ray_dir, vb, ray_bid, voxel_bid)
        idx = torch.nonzero(torch.from_numpy(mask).long(), as_tuple=False)          # pipeline.py:283
        vox, ray = idx[:, 0], idx[:, 1]
        pair_dist = torch.from_numpy(dist)[vox, ray]                                 # pipeline.py:345
        # points: per ray a predicted position along the ray (some inside a hit voxel) + free points
        R = ray_dir.shape[0]
        t = np.random.default_rng(5).random(R).astype(np.float32) * np.float32(2.5)
        if vox.numel() > 0:
            mid = (pair_dist[:, 0] + pair_dist[:, 1]).numpy() * np.float32(0.5)
            t[ray.numpy()] = mid                                                     # last pair of each ray wins
        ray_pts = (ray_dir * t[:, None]).astype(np.float32)
        pmask_rays = ref.pcl_aabb(ray_pts, vb, ray_bid, voxel_bid)
        label = torch.from_numpy(pmask_rays).long()[vox, ray].float()                # pipeline.py:305-309
        P = vox.numel()
        start = np.random.default_rng(6).integers(0, vb.shape[0], size=R).astype(np.int64)
        end_voxel = start.copy()
        pidx = torch.nonzero(torch.from_numpy(pmask_rays).long(), as_tuple=False).numpy()
        for v, n in pidx:                                                            # scatter(..., reduce='max', out=)
            if v > end_voxel[n]:
                end_voxel[n] = v
        pts, pts_bid = pcl_points(7, vb, voxel_bid, 300)
        pmask = ref.pcl_aabb(pts, vb, pts_bid, voxel_bid)
        out = os.path.join(HERE, name + ".npz")
        np.savez_compressed(out, ray_dir=ray_dir, voxel_bound=vb, ray_bid=ray_bid, voxel_bid=voxel_bid,
                            ref_mask=mask, ref_dist=dist, ref_pair_vox=vox.numpy(), ref_pair_ray=ray.numpy(),
                            ref_pair_dist=pair_dist.numpy(), ray_pts=ray_pts, ref_pcl_mask_rays=pmask_rays,
                            ref_pair_label=label.numpy(), end_voxel_start=start, ref_end_voxel=end_voxel,
                            pts=pts, pts_bid=pts_bid, ref_pcl_mask=pmask)
        print(f"{name}: V={vb.shape[0]} R={R} pairs={P} hits/ray={P / R:.2f} pcl hits={int(pmask.sum())} "
              f"ray-point hits={int(pmask_rays.sum())} -> {os.path.getsize(out)} bytes")


if __name__ == "__main__":
    main()
