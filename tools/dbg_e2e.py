import sys, time, torch
sys.path.insert(0, ".")
from implicit_depth_b200.extensions.lidf_query.jit import lidf_query
from implicit_depth_b200.synthetic import make_inputs
import bench
variant = sys.argv[1]
dev = torch.device("cuda", 0)
B, H, W, N = bench.WORKLOADS["c2"]
d = make_inputs(B, H, W, N, seed=1234, device=dev)
off, prob = bench.make_decoders(dev, "IEF")
kw = dict(part_size=d["part_size"], mlp_impl="auto")
host = {k: d[k].cpu().pin_memory() for k in lidf_query.INPUT_KEYS + ("occ_vox_bid",)}
if "32" in variant:
    host = {k: (v.to(torch.int32).pin_memory() if k in lidf_query.INDEX_KEYS else v) for k, v in host.items()}
outputs = ("pred_pos", "max_pair_id") if "sub" in variant else None
t0 = time.perf_counter()
out, h2d, d2h = lidf_query.forward_host(host, off, prob, dev, outputs=outputs, pipeline="mono" not in variant, **kw)
print(variant, "ok", time.perf_counter() - t0, h2d, d2h, flush=True)
