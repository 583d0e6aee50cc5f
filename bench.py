#!/usr/bin/env python
"""bench.py -- LIDF query-points/sec on synthetic batches (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c3] [--engine auto] [--impl ours|reference]

One "step" = one pass of the hot path: device-resident (full_rgb_feat, occ_voxel_feat, rays,
voxel-major pair list, enter/leave distances, decoder weights) -> the six outputs of LIDF.get_pred written.
Workload at every N: BASELINE config 3 per GPU -- 8 images of 640x480 rays x 64 pairs/ray = 157,286,400 query points,
shipped decoders (IEF n_iter 2 + IMNet), i.e. weak scaling by image (no data-path collective).

Printed JSON (one line, rank 0): value = whole-job points/s with inputs resident in HBM (CUDA-event timed, max over
ranks); e2e = same metric through ``lidf_query.forward_host_async`` with pinned HOST buffers, steps issued back to back
(every step's inputs H2D and all six outputs D2H inside the timed region; ``sync_ms_per_step`` = one isolated synchronous
``forward_host`` call); roofline = nominal decoder FLOPs / decoder-kernel time vs the measured bf16 peak; cpu_baseline (N = 1
only) = the oracle port of the reference (stock torch CPU ops + torchvision roi_align) on the host cores, on a bounded
sample.  ``--workload c1|c2|c4|c5`` select the other BASELINE configs, ``--stage2`` adds config 5's RefineNet stage.

``--impl reference`` times that same CPU port as the reference arm (the reference is a Python/PyTorch program with no
compiled artefact; /root/reference is not available on the GPU box).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

WORKLOADS = {          # name: (B per GPU, H, W, N)   -- BASELINE.json configs
    "c1": (1, 64, 64, 16),
    "c2": (4, 240, 320, 64),
    "c3": (8, 480, 640, 64),
    "c4": (4, 480, 640, 64),
    "c5": (8, 480, 640, 128),
    "tiny": (1, 48, 64, 16),
}
FLOP_IMNET = 279168            # 2*MAC of the reference Linear stack, per query point
FLOP_IEF2 = 574784
FLOP_PER_POINT = {"IEF": FLOP_IEF2 + FLOP_IMNET, "IMNET": 2 * FLOP_IMNET}
def ncu_traffic(workload):
    """dram__bytes_read.sum + dram__bytes_write.sum of ONE k_mlp_tc launch, from the latest committed ncu capture
    (profiles/traffic.json, written when a capture is summarised; names the capture it came from)."""
    try:
        with open(os.path.join(REPO, "profiles", "traffic.json")) as f:
            t = json.load(f).get(workload)
        return (float(t["bytes"]), t["source"]) if t else (None, None)
    except Exception:
        return None, None
ALGO_BYTES_PER_POINT = 80      # decoder kernel: perm 4 + pair_vox 8 + pair_ray 8 + dist 8 + T row share 32 read; 4 + 4 + 12 written
EXEC_MAC_TC = 626688           # bf16 MACs the tcgen05 engine executes per point: 3 passes x (112*256 + 256*128 + 128*64) x 3 products


_T0 = time.perf_counter()


def progress(msg):
    """phase markers on stderr (LIDF_BENCH_VERBOSE=1): the JSON line on stdout stays the only stdout output"""
    if os.environ.get("LIDF_BENCH_VERBOSE"):
        print(f"[bench {time.perf_counter() - _T0:7.1f}s] {msg}", file=sys.stderr, flush=True)


def load_peaks():
    path = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            pk = json.load(f)
        burst = float(pk.get("bf16_tflops", 1590.0))
        return dict(bf16_tflops=burst, bf16_tflops_sustained=float(pk.get("bf16_tflops_sustained", burst)),
                    hbm_gbs=float(pk.get("hbm_gbs", 6650.0)), source="measured (MEASURED_PEAKS.json)")
    return dict(bf16_tflops=1590.0, bf16_tflops_sustained=1400.0, hbm_gbs=6650.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi sampling during the timed region (B200_PROFILING.md clocks line)."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        try:
            self.proc.wait(2)
        except Exception:
            self.proc.kill()
        sm = [int(r[0]) for r in self.rows if len(r) >= 6 and r[0].isdigit()]
        mx = [int(r[1]) for r in self.rows if len(r) >= 6 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 6 for n, v in zip(names, r[2:6]) if v.lower().startswith("active")})
        return dict(sm_mhz=statistics.median(sm) if sm else None, sm_max_mhz=max(mx) if mx else None, reasons=reasons,
                    samples=len(sm))


def make_decoders(device, offdec="IEF"):
    """'trained-like' decoders with the reference's state_dict keys, as modules."""
    from implicit_depth_b200.models import implicit_net as N
    g = torch.Generator().manual_seed(99)
    off = N.IEF(device, 385, 1, gf_dim=64, n_iter=2) if offdec == "IEF" else N.IMNet(385, 1, gf_dim=64)
    prob = N.IMNet(385, 1, gf_dim=64)
    for mod in (off, prob):
        for name, p in mod.named_parameters():
            with torch.no_grad():
                if p.dim() == 2:
                    p.copy_(torch.randn(p.shape, generator=g) / (p.shape[1] ** 0.5))
                else:
                    p.copy_(torch.randn(p.shape, generator=g) * 0.1)
    return off.to(device).eval(), prob.to(device).eval()


def cpu_reference_points_per_s(workload, offdec, budget_pairs, steps=1, warmup=0, threads=None):
    """The reference's op chain on the host cores: oracle port + torchvision roi_align, per pair, fp32.
    Sample = the first ``budget_pairs`` pairs worth of rays of image 0 of the same synthetic workload."""
    from implicit_depth_b200.synthetic import make_inputs
    from oracle import lidf_oracle as O
    try:
        import torchvision.ops as tv_ops
        roi_fn = lambda feat, boxes, out, scale: tv_ops.roi_align(feat, boxes, output_size=out, spatial_scale=scale, aligned=True)
        roi_name = "torchvision.ops.roi_align"
    except Exception:
        roi_fn, roi_name = None, "oracle roi_align restatement"
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    B, H, W, N = WORKLOADS[workload]
    rows = max(1, min(H, budget_pairs // (W * N)))
    d = make_inputs(1, rows, W, N, seed=1234)          # one image strip of `rows` x W rays, N pairs each
    g = torch.Generator().manual_seed(99)
    cfg = dict(O.DEFAULT_CFG, offdec_type=offdec)
    off = O.init_decoder(offdec, 385, mode="trained", generator=g)
    prob = O.init_decoder("IMNET", 385, mode="trained", generator=g)
    P = d["occ_vox_intersect_idx"].shape[0]
    # The reference itself when its tree is available (baseline/_ref/src travels with the repo snapshot; /root/reference/src in
    # the build container): the UNMODIFIED LIDF.get_embedding + LIDF.get_pred on the host cores, one call over the sample as
    # the reference runs it (torch_scatter -> oracle/ref_loader.py's torch shim, the two producers stubbed with the
    # sample's features).  Otherwise the oracle port of the same op chain.
    run, kind = None, "port"
    from oracle import ref_loader
    if ref_loader.find_ref_src() is not None:
        try:
            import contextlib
            cpu = torch.device("cpu")
            with contextlib.redirect_stdout(sys.stderr):
                ref, opt = ref_loader.load({"model.offdec_type": offdec})
                lidf = ref.LIDF(opt, cpu).eval()
            lidf.offset_dec.load_state_dict(off); lidf.prob_dec.load_state_dict(prob)

            class _Fixed(torch.nn.Module):
                def __init__(self, v):
                    super().__init__(); self.v = v

                def forward(self, *a, **kw):
                    return self.v
            R0, V0 = int(d["miss_ray_dir"].shape[0]), int(d["voxel_bound"].shape[0])
            lidf.resnet_model = _Fixed(d["full_rgb_feat"]); lidf.pnet_model = _Fixed(d["occ_voxel_feat"])
            dense = torch.zeros(V0, R0, 2)
            dense[d["occ_vox_intersect_idx"], d["miss_ray_intersect_idx"]] = d["intersect_dist"]
            base = dict(bs=1, h=rows, w=W, dist=dense, occ_vox_intersect_idx=d["occ_vox_intersect_idx"],
                        miss_ray_intersect_idx=d["miss_ray_intersect_idx"], miss_ray_dir=d["miss_ray_dir"],
                        miss_img_ind=d["miss_img_ind"], miss_bid=d["miss_bid"], voxel_bound=d["voxel_bound"],
                        occ_vox_bid=d["occ_vox_bid"], rgb_img=torch.zeros(1, 3, rows, W), valid_rgb=torch.zeros(4, 3),
                        valid_v_pid=torch.zeros(4, dtype=torch.long), valid_v_rel_coord=torch.zeros(4, 3),
                        revidx=torch.zeros(4, dtype=torch.long), part_size=d["part_size"], total_miss_sample_num=R0,
                        item_path=["synthetic"])

            def run():
                dd = dict(base)
                lidf.get_embedding(dd)
                lidf.get_pred(dd, "test", 0)
                return dd["pred_pos"]
            kind = "reference"
            roi_name = "unmodified reference LIDF.get_embedding + get_pred (torch CPU ops, torchvision roi_align, torch_scatter shim)"
        except Exception as e:                                                       # noqa: BLE001
            progress(f"reference tree present but not runnable on the CPU ({type(e).__name__}: {e}); timing the port")
            run, kind = None, "port"
    if run is None:
        run = lambda: O.lidf_query_chunked(d, cfg, off, prob, d["part_size"], chunk_pairs=1 << 18, roi_fn=roi_fn)
    times = []
    with torch.no_grad():
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            run()
            if i >= warmup:
                times.append(time.perf_counter() - t0)
    t = sum(times) / len(times)
    sample = f"{rows}x{W} rays x {N} pairs = {P} points of the {workload} workload, {steps} pass(es), {roi_name}"
    return P / t, t, threads, sample, P, kind


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    pts, t, threads, sample, P, kind = cpu_reference_points_per_s(args.workload, args.offdec, args.cpu_sample_pairs,
                                                                  steps=max(1, args.steps), warmup=min(args.warmup, 1))
    B, H, W, N = WORKLOADS[args.workload]
    line = dict(impl="reference", metric="lidf_query_points_per_sec", value=pts, unit="points/s", n_gpus=args.gpus,
                steps=args.steps, warmup=args.warmup, ms_per_step=t * 1e3, higher_is_better=True, scaling="weak",
                vs_baseline=None, dtype="f32", data="synthetic",
                config=dict(workload=f"{args.workload}: {H}x{W} rays x {N} pairs/ray, decoders {args.offdec}+IMNET "
                                     f"(bounded sample of it, see cpu_baseline.sample)"),
                cpu_baseline=dict(value=pts, unit="points/s", cores=threads, kind=kind, sample=sample, device="cpu"),
                device="cpu (host cores; the same-GPU torch arm is torch_gpu_baseline in the default line)",
                e2e=dict(value=pts, unit="points/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--engine", default="auto", choices=["auto", "simt_fp32", "tc_bf16x3", "tc_bf16x1"])
    ap.add_argument("--offdec", default="IEF", choices=["IEF", "IMNET"])
    ap.add_argument("--cpu-sample-pairs", type=int, default=1 << 19)
    ap.add_argument("--pair-order", default="nonzero", choices=["nonzero", "ray"],
                    help="order of the synthetic pair list: 'nonzero' = the reference's torch.nonzero order (voxel-major; "
                         "the call regroups by ray inside the timed region: the default and the headline), 'ray' = the list "
                         "as lidf_ray_aabb_pairs_ray_major_* emits it (sorted by ray; LidfQueryParams::pairs_ray_major: no regroup)")
    ap.add_argument("--no-winner-only", action="store_true", help="skip the extra winner-only-mode timing (N = 1 only)")
    ap.add_argument("--winner-only", action="store_true",
                    help="--train only: forward + backward in winner-only mode (offset decoder on each ray's arg-max pair; the "
                         "same gradients, see include/lidf_query.h); stated in config")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--stage2", action="store_true",
                    help="also time BASELINE config 5's second stage on this rank's rays: end-voxel lookup + voxel-feature "
                         "gather + the RefineNet decoder tail, forward_times = 2 (reported as 'stage2', not part of 'value')")
    ap.add_argument("--no-torch-gpu-baseline", action="store_true",
                    help="skip timing the stock torch op chain on the same GPU (north_star's '>= 10x the reference PyTorch "
                         "decoder' target; N = 1 only, bounded sample of >= 2^20 points)")
    ap.add_argument("--cuda-graph", action="store_true",
                    help="time the device-resident step as a CUDA-graph replay (lidf_query.make_graphed_forward): for the "
                         "launch-bound small configs (c1)")
    ap.add_argument("--train", action="store_true",
                    help="BASELINE config 4's training step on this rank's images: fused forward + lidf_query_backward + "
                         "NCCL all-reduce of the decoder gradients (DDP's exchange), synthetic upstream gradients")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the lidf_query path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    from implicit_depth_b200.extensions.lidf_query.jit import bind_to_device_numa_node, lidf_query
    from implicit_depth_b200.synthetic import make_inputs

    # one process per GPU: stay on the GPU's NUMA node so the pinned host buffers of the e2e path sit next to its PCIe port
    # (N > 1 only: at N = 1 the process keeps every core for the CPU baseline it also has to time)
    numa_node = bind_to_device_numa_node(local)[0] if world > 1 else None

    B, H, W, N = WORKLOADS[args.workload]
    d = make_inputs(B, H, W, N, seed=1234 + rank, device=dev)      # this rank's images; pair list voxel-major
    off, prob = make_decoders(dev, args.offdec)
    P = int(d["occ_vox_intersect_idx"].shape[0]); R = int(d["miss_ray_dir"].shape[0])
    kw = dict(part_size=d["part_size"], mlp_impl=args.engine)
    if args.pair_order == "ray":                                   # what compute_ray_aabb hands over with pair_order = "ray"
        o = torch.sort(d["miss_ray_intersect_idx"], stable=True).indices
        for k in ("occ_vox_intersect_idx", "miss_ray_intersect_idx", "intersect_dist"):
            d[k] = d[k][o].contiguous()
        del o
        kw["pairs_ray_major"] = True
    ins = [d[k] for k in lidf_query.INPUT_KEYS]
    if args.train:
        return run_train(args, d, ins, off, prob, kw, dev, rank, world, local)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    graphed = lidf_query.make_graphed_forward(*ins, off, prob, **kw) if args.cuda_graph else None

    def step():
        return graphed() if graphed is not None else lidf_query.forward(*ins, off, prob, **kw)

    in_bytes = sum(t.numel() * t.element_size() for t in ins)
    flush_buf = torch.empty(192 << 20, dtype=torch.uint8, device=dev) if in_bytes <= 126e6 else None
    progress(f"inputs ready: P={P} R={R}")
    for _ in range(max(3, args.warmup)):
        out = step()
    del out
    torch.cuda.synchronize()
    progress("warm-up done")
    sampler = ClockSampler(local)
    barrier()
    if rank == 0:
        sampler.start()
    lidf_query.launch_count(reset=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    mlp_ms = []
    if flush_buf is None:
        e0.record()
        for _ in range(args.steps):
            out = step()
            del out
        e1.record()
        barrier()
        total_ms = e0.elapsed_time(e1)
    else:                                            # small workload: flush L2 between the timed iterations, time each step alone
        total_ms = 0.0
        for _ in range(args.steps):
            flush_buf.fill_(1)
            e0.record()
            out = step()
            e1.record()
            del out
            torch.cuda.synchronize()
            total_ms += e0.elapsed_time(e1)
        barrier()
    launches = lidf_query.launch_count()
    # decoder-kernel time: CUDA events recorded by the library around that launch, on the launching stream
    for _ in range(min(3, args.steps)):
        step()
        mlp_ms.append(lidf_query.last_mlp_ms())
    torch.cuda.synchronize()
    if args.cuda_graph:                              # the library's event pair is not usable inside a captured graph
        mlp_ms = [total_ms / args.steps]
    clocks = sampler.stop() if rank == 0 else None
    # ---- extra (never the headline): winner-only mode -- per-ray outputs only, the offset decoder on each ray's arg-max
    # pair (LidfQueryParams::winner_only_offset; pred_pos / max_pair_id / pred_prob_end* bit-identical to the full call) ----
    winner = None
    if not args.no_winner_only and not args.cuda_graph and args.engine != "simt_fp32" and world == 1:
        kw_w = dict(kw, winner_only=True)
        for _ in range(2):
            lidf_query.forward(*ins, off, prob, **kw_w)
        torch.cuda.synchronize()
        w0, w1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        w0.record()
        for _ in range(args.steps):
            o = lidf_query.forward(*ins, off, prob, **kw_w)
            del o
        w1.record()
        torch.cuda.synchronize()
        w_ms = w0.elapsed_time(w1) / args.steps
        winner = dict(value=P / (w_ms * 1e-3), unit="points/s", ms_per_step=w_ms, steps=args.steps,
                      note="NOT the headline: pred_pos / max_pair_id / pred_prob_end / pred_prob_end_softmax only (what the "
                           "reference reads downstream of get_pred), bit-identical to the full call; the probability decoder "
                           "runs over all P pairs, the offset decoder on one row per ray; pred_offset / pair_pred_pos are not produced")
    t = torch.tensor([total_ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step = float(t) / args.steps
    value = P * world / (ms_per_step * 1e-3)

    # ---- e2e: the same call with pinned HOST buffers, H2D + D2H inside the timed region ----------------------
    # Steps are issued back to back through forward_host_async (two sets of pinned output buffers): every step copies its
    # inputs host -> device and all six outputs device -> host; the copies of neighbouring steps overlap the decoder kernel.
    # `sync_ms_per_step` is one isolated, fully synchronous forward_host call (pipeline fill + drain exposed).
    progress(f"device-resident timing done: {ms_per_step:.1f} ms/step")
    e2e = None
    if not args.no_e2e:
        def measure_e2e(host, outputs, n_e2e, kw=kw):
            progress(f"e2e: outputs={outputs} steps={n_e2e}")
            out_a, h2d, d2h = lidf_query.forward_host(host, off, prob, dev, outputs=outputs, **kw)   # warm-up, allocates pinned outputs
            out_b = {k: torch.empty_like(v).pin_memory() for k, v in out_a.items()}
            bufs = (out_a, out_b)
            lidf_query.forward_host(host, off, prob, dev, out_host=out_b, outputs=outputs, **kw)
            barrier()
            t0 = time.perf_counter()
            lidf_query.forward_host(host, off, prob, dev, out_host=out_a, outputs=outputs, **kw)
            sync_ms = (time.perf_counter() - t0) * 1e3
            barrier()

            def stream_steps(n):
                pending = None
                for i in range(n):
                    call = lidf_query.forward_host_async(host, off, prob, dev, out_host=bufs[i & 1], outputs=outputs, **kw)
                    if pending is not None:
                        pending.wait()                                                # step i-1's outputs are on the host
                    pending = call
                pending.wait()
                torch.cuda.synchronize()
            stream_steps(3)                              # warm-up: two steps in flight double the device working set
            barrier()
            t0 = time.perf_counter()
            stream_steps(n_e2e)
            te = torch.tensor([(time.perf_counter() - t0) / n_e2e, sync_ms], device=dev)
            if world > 1:
                dist.all_reduce(te, op=dist.ReduceOp.MAX)
            return dict(value=P * world / float(te[0]), unit="points/s", h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h,
                        ms_per_step=float(te[0]) * 1e3, steps=n_e2e, sync_ms_per_step=float(te[1]), host_numa_node=numa_node)

        # Headline e2e = the inference call a user makes (test_lidf.yaml): every input H2D from pinned host memory -- the four
        # index arrays held as int32 on the host (widened on the device) -- and the results inference reads, pred_pos and
        # max_pair_id, D2H.  `e2e_all_outputs` (N = 1 only) is the round-1 definition: int64 indices in, all six outputs
        # out incl. the four per-pair tensors that only the training losses read.
        host64 = {k: d[k].cpu().pin_memory() for k in lidf_query.INPUT_KEYS + ("occ_vox_bid",)}   # occ_vox_bid stays on the host
        host = {k: (v.to(torch.int32).pin_memory() if k in lidf_query.INDEX_KEYS else v) for k, v in host64.items()}
        e2e = measure_e2e(host, ("pred_pos", "max_pair_id"), max(2, min(args.steps, 5)))
        e2e["mode"] = ("steps issued back to back (forward_host_async); per step: all inputs H2D (index arrays int32 on the host), "
                       "pred_pos + max_pair_id D2H, inside the timed region")
        if world == 1:
            full = measure_e2e(host64, None, 2)
            e2e["all_outputs_int64"] = dict(value=full["value"], ms_per_step=full["ms_per_step"],
                                            h2d_bytes_per_step=full["h2d_bytes_per_step"], d2h_bytes_per_step=full["d2h_bytes_per_step"],
                                            note="round-1 definition: int64 index arrays in, all six outputs out")
            if winner is not None:      # extra, never the headline: the same host-buffer loop in winner-only mode
                we = measure_e2e(host, ("pred_pos", "max_pair_id"), 3, kw=dict(kw, winner_only=True))
                winner["e2e"] = dict(value=we["value"], ms_per_step=we["ms_per_step"], h2d_bytes_per_step=we["h2d_bytes_per_step"],
                                     d2h_bytes_per_step=we["d2h_bytes_per_step"])
        del host, host64

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = load_peaks()
    k_ms = statistics.median(mlp_ms)
    flop_pt = FLOP_PER_POINT[args.offdec]
    achieved = P * flop_pt / (k_ms * 1e-3) / 1e12
    peak = peaks["bf16_tflops_sustained"]           # the kernel is ~the whole of a long step -> sustained figure
    traffic, traffic_src = ncu_traffic(args.workload) if (args.engine != "simt_fp32" and args.offdec == "IEF") else (None, None)
    roofline = dict(bound="tensor", achieved=achieved, peak=peak, unit="TFLOP/s", frac=achieved / peak, traffic=traffic,
                    traffic_source=traffic_src,
                    algorithmic_bytes_per_launch=P * ALGO_BYTES_PER_POINT,
                    kernel="k_mlp_simt" if args.engine == "simt_fp32" else "k_mlp_tc", kernel_ms=k_ms,
                    kernel_share_of_step=k_ms / ms_per_step, flop_per_point_nominal=flop_pt,
                    executed_tflops=(P * 2 * EXEC_MAC_TC / (k_ms * 1e-3) / 1e12) if args.engine != "simt_fp32" else None,
                    peak_source=peaks["source"] + ", bf16_tflops_sustained")
    progress("e2e done")
    cpu = None
    if not args.no_cpu_baseline and world == 1:     # the CPU baseline is an N = 1 figure (rank 0, all host cores)
        pts, tcpu, threads, sample, _, ckind = cpu_reference_points_per_s(args.workload, args.offdec, args.cpu_sample_pairs)
        cpu = dict(value=pts, unit="points/s", cores=threads, kind=ckind, sample=sample, seconds=tcpu, device="cpu")
    line = dict(metric="lidf_query_points_per_sec", value=value, unit="points/s", n_gpus=world, steps=args.steps,
                warmup=max(3, args.warmup), ms_per_step=ms_per_step, higher_is_better=True, scaling="weak",
                vs_baseline=None, dtype="bf16x3 (split bf16 operands, fp32 accumulate)" if args.engine != "simt_fp32" else "f32",
                data="synthetic",
                config=dict(workload=f"{args.workload}: {B} images/GPU of {H}x{W} rays x {N} pairs/ray = {P} points/GPU, "
                                     f"decoders {args.offdec}(n_iter 2)+IMNET, trained-like random weights",
                            engine=args.engine, cuda_graph=bool(args.cuda_graph), l2=("inputs (>3 GB/step) exceed the 126 MB L2; no explicit flush" if in_bytes > 126e6 else
                                f"inputs ({in_bytes / 1e6:.1f} MB) fit in the 126 MB L2: a 192 MB buffer is written between steps to flush it"),
                            pair_order="reference voxel-major (regroup inside the timed region)" if args.pair_order == "nonzero" else
                                       "ray-major, as lidf_ray_aabb_pairs_ray_major_* emits it (pairs_ray_major: no regroup)"),
                clocks=clocks, e2e=e2e, gpu_launches=launches, roofline=roofline, cpu_baseline=cpu)
    if winner is not None:
        line["winner_only_mode"] = winner
    progress("cpu baseline done")
    if not args.no_torch_gpu_baseline and world == 1:
        tg = torch_gpu_baseline(d, off, prob, args, dev)
        progress("torch gpu baseline done")
        line["torch_gpu_baseline"] = tg
        line["vs_torch_gpu"] = dict(value=value / tg["value"], e2e=(e2e["value"] / tg["value"]) if e2e else None,
                                    note="this arm (device-resident / e2e) over the stock torch op chain on the same B200; "
                                         "north_star target >= 10")
    if args.stage2:
        line["stage2"] = stage2(d, step, dev, impl=args.engine)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_train(args, d, ins, off, prob, kw, dev, rank, world, local):
    """BASELINE config 4: one training step of the hot path per rank = fused forward (activations are not kept) +
    lidf_query_backward (recompute + tcgen05 dgrad / wgrad) + all-reduce of the decoder gradients over NCCL (the one
    exchange DDP performs, reference src/trainers/train_lidf.py:120,394).  Upstream gradients dL/d pred_pos and
    dL/d pred_prob_end are synthetic and resident (the reference's compute_loss is outside the path)."""
    import torch.distributed as dist
    from implicit_depth_b200.extensions.lidf_query.jit import lidf_query
    P = int(d["occ_vox_intersect_idx"].shape[0]); R = int(d["miss_ray_dir"].shape[0])
    B, H, W, N = WORKLOADS[args.workload]
    g = torch.Generator(device=dev).manual_seed(77 + rank)
    g_pos = torch.randn(R, 3, generator=g, device=dev) / R
    g_prob = torch.randn(P, 1, generator=g, device=dev) / P
    names = [k for k, _ in off.named_parameters()], [k for k, _ in prob.named_parameters()]
    n_dec_params = sum(p.numel() for p in off.parameters()) + sum(p.numel() for p in prob.parameters())
    full_model = torch.zeros(21641314, device=dev) if world > 1 else None      # the reference's whole DDP payload
    ev = lambda: torch.cuda.Event(enable_timing=True)
    t_ar, t_fwd, t_bwd, t_bwd_tc, t_mlp = [], [], [], [], []

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step(record=False):
        e = [ev() for _ in range(4)]
        e[0].record()
        out = lidf_query.forward(*ins, off, prob, save_for_backward=True, winner_only=args.winner_only, **kw)
        e[1].record()
        res = lidf_query.backward(*ins, off, prob, out, g_pred_pos=g_pos, g_pred_prob_end=g_prob, winner_only=args.winner_only, **kw)
        e[2].record()
        flat = torch.cat([res["offset_dec"][k].reshape(-1) for k in names[0]] + [res["prob_dec"][k].reshape(-1) for k in names[1]])
        if world > 1:
            dist.all_reduce(flat)
            flat /= world
        e[3].record()
        if record:
            t_mlp.append(lidf_query.last_mlp_ms()); t_bwd_tc.append(lidf_query.last_bwd_ms())
            torch.cuda.synchronize()
            t_fwd.append(e[0].elapsed_time(e[1])); t_bwd.append(e[1].elapsed_time(e[2])); t_ar.append(e[2].elapsed_time(e[3]))
        return flat

    for _ in range(max(3, args.warmup)):
        step()
    sampler = ClockSampler(local)
    barrier()
    if rank == 0:
        sampler.start()
    lidf_query.launch_count(reset=True)
    e0, e1 = ev(), ev()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    launches = lidf_query.launch_count()
    total_ms = e0.elapsed_time(e1)
    for _ in range(min(3, args.steps)):
        step(record=True)
    ar_full = None
    if world > 1:
        for _ in range(2):
            dist.all_reduce(full_model)
        a, b = ev(), ev()
        barrier(); a.record()
        for _ in range(5):
            dist.all_reduce(full_model)
        b.record(); torch.cuda.synchronize()
        ar_full = a.elapsed_time(b) / 5
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([total_ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step = float(t) / args.steps
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peaks = load_peaks()
    med = statistics.median
    flop_pt = 3 * FLOP_PER_POINT[args.offdec]                 # forward + backward (dgrad + wgrad) as written by the reference
    k_ms = med(t_mlp) + med(t_bwd_tc)
    achieved = P * flop_pt / (k_ms * 1e-3) / 1e12
    peak = peaks["bf16_tflops_sustained"]
    line = dict(metric="lidf_train_step_points_per_sec", value=P * world / (ms_per_step * 1e-3), unit="points/s", n_gpus=world,
                steps=args.steps, warmup=max(3, args.warmup), ms_per_step=ms_per_step, higher_is_better=True, scaling="weak",
                vs_baseline=None, dtype="bf16x3 (split bf16 operands, fp32 accumulate)", data="synthetic",
                config=dict(workload=f"{args.workload} train: {B} images/GPU of {H}x{W} rays x {N} pairs/ray = {P} points/GPU, "
                                     f"decoders {args.offdec}(n_iter 2)+IMNET; step = fused forward + native backward + "
                                     f"all-reduce of the {n_dec_params} decoder gradients",
                            l2="inputs exceed the 126 MB L2; no explicit flush",
                            mode=("winner-only: offset decoder forward + backward on each ray's arg-max pair (R rows), probability "
                                  "decoder on all P pairs; gradients equal the full backward's (tests/test_gpu_backward.py)")
                                 if args.winner_only else "full: both decoders forward + backward over all P pairs"),
                clocks=clocks, gpu_launches=launches,
                train=dict(forward_ms=med(t_fwd), backward_ms=med(t_bwd), allreduce_decoder_grads_ms=med(t_ar),
                           allreduce_bytes=4 * n_dec_params, allreduce_full_model_86MB_ms=ar_full,
                           forward_mlp_kernel_ms=med(t_mlp), backward_tc_section_ms=med(t_bwd_tc)),
                roofline=dict(bound="tensor", achieved=achieved, peak=peak, unit="TFLOP/s", frac=achieved / peak, traffic=None,
                              kernel="k_mlp_tc + k_mlp_bwd_tc + k_wgrad_pk_tc + k_wgrad_tc", kernel_ms=k_ms, kernel_share_of_step=k_ms / ms_per_step,
                              flop_per_point_nominal=flop_pt, peak_source=peaks["source"] + ", bf16_tflops_sustained",
                              note="nominal forward + backward FLOPs of the reference's Linear stack (3 x forward) over the "
                                   "decoder kernel of the forward plus the backward's tcgen05 section; the backward skips the offset "
                                   "decoder's rows whose upstream gradient is exactly zero (all but one pair per ray: the loss "
                                   "reads pred_pos and pred_prob_end only), so executed work is well below nominal"),
                e2e=None, cpu_baseline=None)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def stage2(d, step, dev, forward_times=2, impl="auto", valid_per_image=10000):
    """RefineNet.get_pred_refine per ray (reference pipeline.py:922-1041) on this repo's kernels: end-voxel lookup
    (pcl_aabb.end_voxel), PointNet2Stage re-run over the valid points + one predicted point per ray (lidf_pointnet_forward),
    cached per-ray ROI feature, decoder tail incl. the occ_voxel_feat[end_voxel_id] gather (lidf_refine_forward), repeated
    forward_times (train_refine.yaml:82).  The torch glue between them (voxel centres, torch.cat of the PointNet inputs,
    pipeline.py:975-1010) is inside the timed region."""
    from implicit_depth_b200.extensions.lidf_query.jit import lidf_query
    from implicit_depth_b200.extensions.pcl_aabb.jit import pcl_aabb
    from implicit_depth_b200.models import implicit_net as N
    from implicit_depth_b200.models.pointnet import PointNet2Stage, pointnet_forward
    g = torch.Generator().manual_seed(7)
    dec = N.IEF(dev, 334, 1, gf_dim=64, n_iter=2)
    with torch.no_grad():
        for p_ in dec.parameters():
            p_.copy_(torch.randn(p_.shape, generator=g) / (p_.shape[1] ** 0.5) if p_.dim() == 2 else torch.randn(p_.shape, generator=g) * 0.1)
    dec = dec.to(dev).eval()
    pnet = PointNet2Stage(6, 128, 32).to(dev).eval()
    out = step()
    roi = lidf_query.roi_align_rays(d["full_rgb_feat"], d["miss_img_ind"], d["miss_bid"], 8)
    vox = d["occ_vox_intersect_idx"]
    dummy = torch.cat((vox, torch.zeros(1, dtype=vox.dtype, device=dev)), 0)
    rb, xb = d["miss_bid"].int(), d["occ_vox_bid"].int()
    R, V = int(d["miss_ray_dir"].shape[0]), int(d["occ_voxel_feat"].shape[0])
    B = int(d["full_rgb_feat"].shape[0])
    nv = valid_per_image * B                                         # grid.valid_sample_num points per image (test_lidf.yaml:55)
    pnet_inp = torch.cat((0.25 * (torch.rand(nv, 3, generator=g) - 0.5), torch.rand(nv, 3, generator=g)), 1).to(dev)
    revidx = torch.randint(0, V, (nv,), generator=g).sort().values.to(dev)
    miss_rgb = torch.rand(R, 3, generator=g).to(dev)
    vb = d["voxel_bound"]
    t_pn = []

    def run(record=False):
        pos = out["pred_pos"]
        with torch.no_grad():
            for _ in range(forward_times):
                end = pcl_aabb.end_voxel(pos, vb, rb, xb, dummy[out["max_pair_id"]].contiguous())
                evb = vb[end]
                pred_inp = torch.cat((pos - (evb[:, :3] + evb[:, 3:]) / 2., miss_rgb), 1)          # pipeline.py:975-984
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                feat = pointnet_forward(pnet, torch.cat((pnet_inp, pred_inp), 0), torch.cat((revidx, end), 0), V)   # :1008-1014
                e1.record()
                if record:
                    t_pn.append((e0, e1))
                pos = lidf_query.refine_forward(pos, d["miss_ray_dir"], None, None, roi, dec, occ_voxel_feat=feat,
                                                end_voxel_id=end, voxel_bound=vb, mlp_impl=impl)
        return pos
    for _ in range(2):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        run(record=True)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    pn_ms = sum(a.elapsed_time(b) for a, b in t_pn) / 3
    return dict(ms=ms, pointnet_ms=pn_ms, rays=R, pointnet_points=nv + R, forward_times=forward_times,
                rays_per_s=R * forward_times / (ms * 1e-3))


def torch_gpu_baseline(d, off, prob, args, dev, rays=1 << 15):
    """north_star's yardstick: 'the reference PyTorch decoder' on the SAME B200.  When the reference tree is available
    (baseline/_ref/src, staged by __graft_entry__.build(), or /root/reference/src) this drives the UNMODIFIED
    ``LIDF.get_embedding`` + ``LIDF.get_pred`` (reference src/models/pipeline.py:338-466) on CUDA -- stock torch ops, fp32,
    TF32 off, torch_scatter replaced by a torch scatter_reduce shim (oracle/ref_loader.py), the two producers replaced by
    modules returning the sample's features -- otherwise the oracle port of the same op chain.  Sample: the first ``rays``
    rays of image 0 (>= 2^20 query points at 64 pairs/ray) in one un-chunked call, as the reference would run them."""
    from oracle import lidf_oracle as O
    from oracle import ref_loader
    import torchvision.ops as tv_ops
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    m = d["miss_ray_intersect_idx"] < rays
    sub = dict(d)
    for k in ("occ_vox_intersect_idx", "miss_ray_intersect_idx", "intersect_dist"):
        sub[k] = d[k][m]
    P = int(sub["occ_vox_intersect_idx"].shape[0])
    B, H, W, N = WORKLOADS[args.workload]
    kind, run = None, None
    if ref_loader.find_ref_src() is not None:
        try:
            import contextlib
            with contextlib.redirect_stdout(sys.stderr):                             # the reference prints from its constructors
                ref, opt = ref_loader.load({"model.offdec_type": args.offdec})
                lidf = ref.LIDF(opt, dev).to(dev).eval()
            lidf.offset_dec.load_state_dict(off.state_dict()); lidf.prob_dec.load_state_dict(prob.state_dict())

            class _Fixed(torch.nn.Module):
                def __init__(self, v):
                    super().__init__(); self.v = v

                def forward(self, *a, **kw):
                    return self.v
            V0 = int((d["occ_vox_bid"] == 0).sum())                                  # image 0's voxels (rays < `rays` are image 0)
            assert rays <= H * W and int(sub["occ_vox_intersect_idx"].max()) < V0
            lidf.resnet_model = _Fixed(d["full_rgb_feat"][:1]); lidf.pnet_model = _Fixed(d["occ_voxel_feat"][:V0])
            dist = torch.zeros(V0, rays, 2, device=dev)
            dist[sub["occ_vox_intersect_idx"], sub["miss_ray_intersect_idx"]] = sub["intersect_dist"]
            base = dict(bs=1, h=H, w=W, dist=dist, occ_vox_intersect_idx=sub["occ_vox_intersect_idx"],
                        miss_ray_intersect_idx=sub["miss_ray_intersect_idx"], miss_ray_dir=d["miss_ray_dir"][:rays],
                        miss_img_ind=d["miss_img_ind"][:rays], miss_bid=d["miss_bid"][:rays], voxel_bound=d["voxel_bound"][:V0],
                        occ_vox_bid=d["occ_vox_bid"][:V0], rgb_img=torch.zeros(1, 3, H, W, device=dev),
                        valid_rgb=torch.zeros(4, 3, device=dev), valid_v_pid=torch.zeros(4, dtype=torch.long, device=dev),
                        valid_v_rel_coord=torch.zeros(4, 3, device=dev), revidx=torch.zeros(4, dtype=torch.long, device=dev),
                        part_size=d["part_size"], total_miss_sample_num=rays, item_path=["synthetic"])

            def run():
                dd = dict(base)
                lidf.get_embedding(dd)
                lidf.get_pred(dd, "test", 0)
                return dd["pred_pos"]
            with torch.no_grad():
                run()
            kind = "reference: unmodified LIDF.get_embedding + get_pred on CUDA (torch_scatter shim, producers stubbed)"
        except Exception as e:                                                       # noqa: BLE001
            kind, run = None, None
            note = f"reference tree present but not runnable here ({type(e).__name__}: {e}); "
        else:
            note = ""
    else:
        note = "reference tree absent; "
    if run is None:
        cfg = dict(O.DEFAULT_CFG, offdec_type=args.offdec)
        offd = {k: v.detach() for k, v in off.state_dict().items()}
        probd = {k: v.detach() for k, v in prob.state_dict().items()}
        roi_fn = lambda feat, boxes, out, scale: tv_ops.roi_align(feat, boxes, output_size=out, spatial_scale=scale, aligned=True)
        run = lambda: O.lidf_query_chunked(sub, cfg, offd, probd, d["part_size"], chunk_pairs=1 << 21, roi_fn=roi_fn)
        kind = note + "port: oracle restatement of the same op chain on torch CUDA ops"
    with torch.no_grad():
        for _ in range(2):
            run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            run()
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    return dict(value=P / (ms * 1e-3), unit="points/s", device="cuda (same GPU)", kind=kind,
                sample=f"{rays} rays x {N} pairs = {P} points, one call, fp32, TF32 off", ms=ms)


if __name__ == "__main__":
    main()
