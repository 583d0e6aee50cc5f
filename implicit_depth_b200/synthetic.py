"""Seeded synthetic inputs for the LIDF query hot path.

Produces exactly the ``data_dict`` entries that the reference's ``LIDF.get_embedding`` /
``LIDF.get_pred`` read (reference src/models/pipeline.py:338-466): all pixels are miss rays
(``mask_type: all``, pipeline.py:130-133), ray directions follow ``get_miss_ray``
(pipeline.py:210-219) with the ClearGrasp-synthetic intrinsics
(src/datasets/cleargrasp_synthetic_dataset.py:122,145-148), voxel bounds are cells of the
9^3 grid built in ``get_occ_vox_bound`` (pipeline.py:167-189, constants.py:15-16) and the pair
list is emitted in the reference's voxel-major order (``torch.nonzero`` of a [V,R] mask,
pipeline.py:283-285).
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch

XMIN = (-1.0, -1.0, 0.0)     # reference src/constants.py:15
XMAX = (1.0, 1.0, 2.0)       # reference src/constants.py:16
FOV_X = 1.2112585            # reference src/datasets/cleargrasp_synthetic_dataset.py:122


def grid_part_size(res: int = 8) -> float:
    """pipeline.py:169-170: min(XMAX - XMIN) / res."""
    return min(b - a for a, b in zip(XMIN, XMAX)) / res


def make_rays(B: int, H: int, W: int, device, dtype=torch.float32):
    """All-pixel miss rays.  Returns miss_bid [R] i64, miss_img_ind [R,2] (x,y) i64, miss_ray_dir [R,3]."""
    fx = 0.5 * W / math.tan(0.5 * FOV_X)
    fy = fx
    cx, cy = W / 2.0, H / 2.0
    ys, xs = torch.meshgrid(torch.arange(H, device=device), torch.arange(W, device=device), indexing="ij")
    x_ind = xs.to(dtype).unsqueeze(0).repeat(B, 1, 1)
    y_ind = ys.to(dtype).unsqueeze(0).repeat(B, 1, 1)
    img_ind = torch.stack((x_ind, y_ind), -1).reshape(B * H * W, 2).long()
    cam_x = x_ind - cx
    cam_y = (y_ind - cy) * fx / fy
    cam_z = torch.full_like(cam_x, fx)
    ray_dir = torch.stack((cam_x, cam_y, cam_z), -1)
    ray_dir = ray_dir / torch.norm(ray_dir, dim=-1, keepdim=True)
    miss_bid = torch.arange(B, device=device).repeat_interleave(H * W)
    return miss_bid, img_ind, ray_dir.reshape(B * H * W, 3).contiguous()


def make_inputs(B: int, H: int, W: int, N: int, *, V_img: int = 256, seed: int = 1234,
                device="cpu", ragged: bool = False, res: int = 8, rgb_out: int = 32, pnet_out: int = 128,
                ray_major: bool = False) -> Dict[str, torch.Tensor]:
    """Synthetic hot-path inputs: B images of HxW rays with N (or, if ``ragged``, 0..N) pairs per ray.

    Returned pair arrays are in voxel-major order (sorted by voxel then ray) unless ``ray_major``.
    """
    dev = torch.device(device)
    g = torch.Generator(device=dev).manual_seed(seed)
    R_img = H * W
    R = B * R_img
    V = B * V_img
    part = grid_part_size(res)
    ncell = res + 1
    assert V_img <= ncell ** 3 and N <= V_img

    def rnd(*s):
        return torch.rand(*s, generator=g, device=dev)

    full_rgb_feat = torch.randn(B, rgb_out, H, W, generator=g, device=dev)
    occ_voxel_feat = torch.relu(torch.randn(V, pnet_out, generator=g, device=dev))
    # distinct cells of the 9^3 grid per image -> voxel_bound [V,6]
    cells = torch.stack([torch.randperm(ncell ** 3, generator=g, device=dev)[:V_img] for _ in range(B)]).reshape(-1)
    coord = torch.stack((cells // (ncell * ncell), (cells // ncell) % ncell, cells % ncell), -1).float()
    xmin = torch.tensor(XMIN, device=dev) - 0.5 * part
    bound_min = xmin.unsqueeze(0) + coord * part
    voxel_bound = torch.cat((bound_min, bound_min + part), 1).contiguous()
    occ_vox_bid = torch.arange(B, device=dev).repeat_interleave(V_img)

    miss_bid, miss_img_ind, miss_ray_dir = make_rays(B, H, W, dev)

    ray_l, vox_l = [], []
    for b in range(B):
        keys = rnd(R_img, V_img)
        pick = keys.topk(N, dim=1).indices                   # N distinct voxels per ray, uniform w/o replacement
        rid = torch.arange(R_img, device=dev).unsqueeze(1).expand(R_img, N) + b * R_img
        if ragged:
            cnt = torch.randint(0, N + 1, (R_img, 1), generator=g, device=dev)
            keep = torch.arange(N, device=dev).unsqueeze(0) < cnt
            ray_l.append(rid[keep]); vox_l.append(pick[keep] + b * V_img)
        else:
            ray_l.append(rid.reshape(-1)); vox_l.append(pick.reshape(-1) + b * V_img)
        del keys
    ray = torch.cat(ray_l); vox = torch.cat(vox_l)
    P = ray.shape[0]
    key = (ray * V + vox) if ray_major else (vox * R + ray)
    order = torch.argsort(key)
    ray = ray[order].contiguous(); vox = vox[order].contiguous()
    del key, order
    t_enter = 0.3 + 2.2 * rnd(P)
    t_leave = t_enter + 0.02 + 0.413 * rnd(P)
    return dict(
        bs=B, h=H, w=W, part_size=part,
        full_rgb_feat=full_rgb_feat, occ_voxel_feat=occ_voxel_feat, voxel_bound=voxel_bound,
        occ_vox_bid=occ_vox_bid, miss_bid=miss_bid, miss_img_ind=miss_img_ind, miss_ray_dir=miss_ray_dir,
        occ_vox_intersect_idx=vox, miss_ray_intersect_idx=ray,
        intersect_dist=torch.stack((t_enter, t_leave), -1).contiguous(),
        total_miss_sample_num=R,
    )


def shard_images(B: int, rank: int, world: int):
    """Image range owned by ``rank`` (the reference shards by image: DistributedSampler,
    src/trainers/train_lidf.py:163-164).  Returns (first_image, n_images)."""
    base, rem = divmod(B, world)
    n = base + (1 if rank < rem else 0)
    first = rank * base + min(rank, rem)
    return first, n
