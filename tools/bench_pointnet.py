#!/usr/bin/env python
"""PointNet2Stage forward: fused kernels vs the stock torch op chain on the same GPU (fp32, TF32 off).

    python tools/bench_pointnet.py [--points 2537600] [--voxels 2048]

Default size = BASELINE config 5's second stage: 8 x 10^4 valid points + 2,457,600 predicted points over 2,048 voxels."""
import argparse
import json
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--points", type=int, default=2537600)
    ap.add_argument("--voxels", type=int, default=2048)
    ap.add_argument("--steps", type=int, default=10)
    args = ap.parse_args()
    from implicit_depth_b200.models.pointnet import PointNet2Stage, pointnet_forward
    torch.backends.cuda.matmul.allow_tf32 = False
    g = torch.Generator().manual_seed(0)
    net = PointNet2Stage(6, 128, 32).cuda().eval()
    N, V = args.points, args.voxels
    inp = torch.cat((0.25 * (torch.rand(N, 3, generator=g) - 0.5), torch.rand(N, 3, generator=g)), 1).cuda()
    idx = torch.randint(0, V, (N,), generator=g).sort().values.cuda()          # spatially coherent, as image / ray order is

    def timed(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            out = fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / args.steps, out

    with torch.no_grad():
        ms, out = timed(lambda: pointnet_forward(net, inp, idx, V))                                   # tcgen05 (default)
        ms_simt, out_simt = timed(lambda: pointnet_forward(net, inp, idx, V, mlp_impl="simt_fp32"))
        for p in net.parameters():
            p.requires_grad_(True)
        with torch.enable_grad():
            pass
        def torch_chain():
            with torch.enable_grad():                                            # route through the module's torch-op branch
                return net(inp, idx).detach()
        ms_t, out_t = timed(torch_chain)
    err = float((out - out_t).abs().max() / out_t.abs().max())
    macs = N * (6 * 32 + 32 * 64 + 128 * 128 * 2) + V * (64 * 64 + 128 * 128)
    err_s = float((out_simt - out_t).abs().max() / out_t.abs().max())
    print(json.dumps(dict(op="PointNet2Stage forward", points=N, voxels=V, ms=ms, engine="tcgen05 split-bf16 (128->128 layers)",
                          simt_fp32_ms=ms_simt, torch_ms=ms_t, speedup=ms_t / ms, speedup_vs_simt=ms_simt / ms,
                          tflops_nominal=2 * macs / (ms * 1e-3) / 1e12, tflops_simt=2 * macs / (ms_simt * 1e-3) / 1e12,
                          max_rel_diff_vs_torch=err, max_rel_diff_simt_vs_torch=err_s)))


if __name__ == "__main__":
    main()
