#!/usr/bin/env python
"""The reference's own test-time shape (test_lidf.yaml: batch 1, 320x240, mask_type all, 10^4 valid points, 9^3 grid):
ragged pair counts from real ray/voxel geometry.  Times every native stage of the chain and the fused query call.

    python tools/bench_realistic.py [--batch 1] [--steps 50]
"""
import argparse
import json
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def timed(fn, steps, warmup=5):
    for _ in range(warmup):
        out = fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps, out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--pair-order", default="nonzero", choices=["nonzero", "ray"])
    ap.add_argument("--winner-only", action="store_true")
    args = ap.parse_args()
    from bench import make_decoders
    from implicit_depth_b200.extensions.lidf_query.jit import lidf_query
    from implicit_depth_b200.models.pipeline import LIDF, default_opt
    from implicit_depth_b200.models.pointnet import pointnet_forward
    from implicit_depth_b200.synthetic import make_rays
    dev = torch.device("cuda")
    B, H, W = args.batch, 240, 320
    g = torch.Generator().manual_seed(5)
    n_pts = 10000
    # scene: a tilted noisy plane with a bump, as seen by the camera
    xy = torch.rand(B * n_pts, 2, generator=g) * 1.8 - 0.9
    zz = 1.0 + 0.35 * xy[:, :1] + 0.25 * torch.exp(-8 * (xy ** 2).sum(1, keepdim=True)) + 0.01 * torch.randn(B * n_pts, 1, generator=g)
    valid_xyz = torch.cat((xy, zz), 1).contiguous().to(dev)
    valid_bid = torch.arange(B).repeat_interleave(n_pts).to(dev)
    valid_rgb = torch.rand(B * n_pts, 3, generator=g).to(dev)
    miss_bid, miss_img_ind, miss_ray_dir = (t.to(dev) for t in make_rays(B, H, W, "cpu"))
    full_rgb_feat = torch.randn(B, 32, H, W, generator=g).to(dev)
    off, prob = make_decoders(dev, "IEF")
    lidf = LIDF(default_opt(), dev).to(dev).eval()
    lidf.offset_dec, lidf.prob_dec = off, prob
    lidf.pair_order = args.pair_order
    lidf.winner_only = args.winner_only
    dd = dict(bs=B, h=H, w=W, valid_xyz=valid_xyz, valid_bid=valid_bid, miss_bid=miss_bid, miss_img_ind=miss_img_ind,
              miss_ray_dir=miss_ray_dir, full_rgb_feat=full_rgb_feat, total_miss_sample_num=miss_bid.shape[0], item_path=["scene"])
    res = {}
    with torch.no_grad():
        res["get_occ_vox_bound_ms"], _ = timed(lambda: lidf.get_occ_vox_bound(dd), args.steps)
        res["compute_ray_aabb_ms"], _ = timed(lambda: lidf.compute_ray_aabb(dd), args.steps)
        pn_inp = torch.cat((dd["valid_v_rel_coord"], valid_rgb[dd["valid_v_pid"]]), -1).contiguous()
        res["pointnet_ms"], feat = timed(lambda: pointnet_forward(lidf.pnet_model, pn_inp, dd["revidx"], dd["voxel_bound"].shape[0]), args.steps)
        dd["occ_voxel_feat"] = feat
        res["get_pred_ms"], _ = timed(lambda: lidf.get_pred(dd, "test", 0), args.steps)
        lidf_query.launch_count(reset=True)
        lidf.get_pred(dd, "test", 0)
        res["get_pred_launches"] = lidf_query.launch_count()
        res["decoder_kernel_ms"] = lidf_query.last_mlp_ms()
    P, R, V = int(dd["occ_vox_intersect_idx"].shape[0]), int(miss_bid.shape[0]), int(dd["voxel_bound"].shape[0])
    res.update(workload=f"{B} x {H}x{W} all-pixel rays, {n_pts} valid points/image, 9^3 grid", pair_order=args.pair_order, winner_only=args.winner_only, rays=R, voxels=V, pairs=P,
               pairs_per_ray=P / R, get_pred_points_per_s=P / (res["get_pred_ms"] * 1e-3),
               chain_ms=res["get_occ_vox_bound_ms"] + res["compute_ray_aabb_ms"] + res["pointnet_ms"] + res["get_pred_ms"])
    print(json.dumps(res))


if __name__ == "__main__":
    main()
