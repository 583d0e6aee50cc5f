#!/usr/bin/env python
"""Host<->device copy floor of the box: one process per GPU, pinned buffers, H2D and D2H running concurrently on two
streams (what lidf_query.forward_host does under the decoder kernel).  Prints one JSON line (rank 0):
aggregate and per-GPU GB/s for H2D-only, D2H-only and both at once.

    python tools/bench_pcie.py                                    # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tools/bench_pcie.py
"""
import json
import os

import torch


def main():
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    nbytes = 1 << 30
    h_in = torch.empty(nbytes, dtype=torch.uint8).pin_memory(); h_out = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    d_in = torch.empty(nbytes, dtype=torch.uint8, device=dev); d_out = torch.ones(nbytes, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

    def run(h2d, d2h, reps=6):
        def once():
            if h2d:
                with torch.cuda.stream(s1):
                    d_in.copy_(h_in, non_blocking=True)
            if d2h:
                with torch.cuda.stream(s2):
                    h_out.copy_(d_out, non_blocking=True)
        once(); torch.cuda.synchronize()
        if dist:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        s1.wait_event(e0); s2.wait_event(e0)
        for _ in range(reps):
            once()
        torch.cuda.current_stream().wait_stream(s1); torch.cuda.current_stream().wait_stream(s2)
        e1.record(); torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / reps], device=dev)
        if dist:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t)
        per_gpu = (int(h2d) + int(d2h)) * nbytes / (ms * 1e-3) / 1e9
        return dict(ms_per_GiB_round=ms, per_gpu_GBs=per_gpu, aggregate_GBs=per_gpu * world)

    res = dict(n_gpus=world, buffer_bytes=nbytes, h2d_only=run(True, False), d2h_only=run(False, True), both=run(True, True),
               cpus=os.cpu_count(), note="pinned host memory, CUDA-event timed, max over ranks")
    if rank == 0:
        print(json.dumps(res), flush=True)
    if dist:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
