#!/usr/bin/env python
"""Measurement of the ray_aabb / pcl_aabb rows on one B200.

    python tools/bench_aabb.py [--workload c2|c3] [--steps 10]

Geometry = the bench workload's: B images of HxW all-pixel rays against 256 occupied cells of the 9^3 grid per image.
Prints one JSON line per op:
  ray_aabb.forward  dense drop-in; HBM-bound: algorithmic bytes = 12*V*R written (mask int32 + dist 2xfp32) + inputs
  ray_aabb.pairs    compact pair list (count + scan + fill + the one sync); unit = ray-voxel tests/s
  reference sequence on the same GPU = dense forward + mask.long() + torch.nonzero + dist[vox, ray] (pipeline.py:277-285,345)
  pcl_aabb.end_voxel / pair_label
CPU baseline: the reference's own kernel compiled for the CPU (oracle/_ref, kind "reference", 1 core) when present,
else the numpy oracle (kind "port"), on a bounded sample of the same rays/voxels.
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

from bench import WORKLOADS, load_peaks  # noqa: E402


def timed(fn, steps, warmup=3):
    for _ in range(warmup):
        fn()
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
    ap.add_argument("--workload", default="c2", choices=["c2", "c3", "tiny"])
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--no-dense", action="store_true")
    args = ap.parse_args()
    from implicit_depth_b200.extensions.pcl_aabb.jit import pcl_aabb
    from implicit_depth_b200.extensions.ray_aabb.jit import ray_aabb
    from implicit_depth_b200.synthetic import make_inputs
    B, H, W, _ = WORKLOADS[args.workload]
    d = make_inputs(B, H, W, 1, V_img=256, seed=1234, device="cuda")
    rd, vb = d["miss_ray_dir"], d["voxel_bound"]
    rb, xb = d["miss_bid"].int(), d["occ_vox_bid"].int()
    R, V = rd.shape[0], vb.shape[0]
    peaks = load_peaks()
    same_image_tests = R * 256
    base = dict(workload=f"{args.workload}: {B} images of {H}x{W} rays x 256 occupied voxels/image (R={R}, V={V})", steps=args.steps)

    ms, (vox, ray, pd) = timed(lambda: ray_aabb.pairs(rd, vb, rb, xb), args.steps)
    P = int(vox.shape[0])
    print(json.dumps(dict(op="ray_aabb.pairs", ms=ms, pairs=P, pairs_per_ray=P / R, value=same_image_tests / (ms * 1e-3),
                          unit="same-image ray-voxel tests/s", out_bytes=P * 24, **base)), flush=True)
    ms2, (vox2, ray2, pd2) = timed(lambda: ray_aabb.pairs(rd, vb, rb, xb, order="ray"), args.steps)
    o = torch.sort(ray, stable=True).indices
    assert torch.equal(vox2, vox[o]) and torch.equal(ray2, ray[o]) and torch.equal(pd2.view(torch.int32), pd[o].view(torch.int32))
    print(json.dumps(dict(op="ray_aabb.pairs(order='ray')  (ray-major list + CSR: the consumer skips its regroup)", ms=ms2, pairs=P,
                          value=same_image_tests / (ms2 * 1e-3), unit="same-image ray-voxel tests/s", out_bytes=P * 24, **base)), flush=True)
    del vox2, ray2, pd2, o
    dense_bytes = 12 * V * R
    if not args.no_dense and dense_bytes < 60e9:
        ms_d, (mask, dist) = timed(lambda: ray_aabb.forward(rd, vb, rb, xb), args.steps)
        algo = dense_bytes + R * 16 + V * 28
        print(json.dumps(dict(op="ray_aabb.forward (dense drop-in)", ms=ms_d,
                              roofline=dict(bound="hbm", achieved=algo / (ms_d * 1e-3) / 1e9, peak=peaks["hbm_gbs"], unit="GB/s",
                                            frac=algo / (ms_d * 1e-3) / 1e9 / peaks["hbm_gbs"], traffic=None,
                                            algorithmic_bytes_per_launch=algo, peak_source=peaks["source"]), **base)), flush=True)

        def ref_seq():
            m, dd = ray_aabb.forward(rd, vb, rb, xb)
            idx = torch.nonzero(m.long(), as_tuple=False)
            return idx[:, 0], idx[:, 1], dd[idx[:, 0], idx[:, 1]]
        ms_r, (v2, r2, d2) = timed(ref_seq, max(2, args.steps // 2))
        assert torch.equal(v2, vox) and torch.equal(r2, ray) and torch.equal(d2.view(torch.int32), pd.view(torch.int32))
        print(json.dumps(dict(op="reference call sequence on this GPU: dense + mask.long() + nonzero + dist[vox,ray]", ms=ms_r,
                              speedup_of_pairs=ms_r / ms, **base)), flush=True)
        del mask, dist
    mid = (rd[ray] * (0.5 * (pd[:, 0:1] + pd[:, 1:2]))).contiguous()
    rbp = rb[ray].contiguous()
    ar = torch.arange(P, device="cuda")
    ms_l, _ = timed(lambda: pcl_aabb.pair_label(mid, vb, rbp, xb, vox, ar), args.steps)
    print(json.dumps(dict(op="pcl_aabb.pair_label", ms=ms_l, points=P, value=P / (ms_l * 1e-3), unit="pairs/s", **base)), flush=True)
    pts = (rd * 1.2).contiguous()
    start = torch.zeros(R, dtype=torch.int64, device="cuda")
    ms_e, _ = timed(lambda: pcl_aabb.end_voxel(pts, vb, rb, xb, start.clone()), args.steps)
    print(json.dumps(dict(op="pcl_aabb.end_voxel (+ clone of the [R] start ids)", ms=ms_e, points=R, value=R / (ms_e * 1e-3), unit="points/s", **base)), flush=True)

    # CPU baseline on a bounded sample: first `rows` image rows of image 0 against image 0's voxels
    from oracle import aabb_oracle as A
    from oracle import build_ref
    ref = build_ref.load()
    rows = max(1, min(H, (1 << 24) // (W * 256)))
    n = rows * W
    srd, svb = rd[:n].cpu().numpy(), vb[:256].cpu().numpy()
    srb, sxb = rb[:n].cpu().numpy(), xb[:256].cpu().numpy()
    t0 = time.perf_counter()
    if ref is not None:
        ref.ray_aabb(srd, svb, srb, sxb); kind = "reference"
    else:
        A.ray_aabb_dense(srd, svb, srb, sxb); kind = "port"
    dt = time.perf_counter() - t0
    print(json.dumps(dict(op="cpu_baseline ray_aabb", kind=kind, cores=1, seconds=dt, value=n * 256 / dt,
                          unit="same-image ray-voxel tests/s", sample=f"{rows}x{W} rays of image 0 x 256 voxels", **base)), flush=True)


if __name__ == "__main__":
    main()
