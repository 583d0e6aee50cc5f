"""World-size-2 gloo test (CPU): the path shards by image with no data-path collective;
the only exchange is DDP's gradient all-reduce over the decoder parameters.  The product's training path runs on the
sm_100a kernels only, so on this CPU box the op inside the DDP-wrapped module is stood in for by the oracle (test
infrastructure): what is exercised is the host logic -- image sharding, the module surface DDP wraps, gradients landing on
the decoder parameters and being averaged across ranks."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import REPO


def _worker(rank, world, port, q):
    import sys
    sys.path.insert(0, REPO)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from implicit_depth_b200.models.pipeline import LIDF, default_opt
    from implicit_depth_b200.synthetic import make_inputs, shard_images
    torch.manual_seed(0)

    class _Fixed(torch.nn.Module):                                   # stands in for the ResNet / PointNet producers
        def __init__(self, value):
            super().__init__()
            self.value = value

        def forward(self, *a, **kw):
            return self.value

    B = 3
    first, n = shard_images(B, rank, world)
    d = make_inputs(n, 8, 8, 3, V_img=16, seed=100 + first)          # this rank's images only: no exchange
    d["full_rgb_feat"].requires_grad_(True)
    from oracle import lidf_oracle as O

    class _OracleLIDF(LIDF):                                         # CPU stand-in for the fused kernels (oracle = checker)
        def get_pred(self, data_dict, exp_type, epoch):
            cfg = dict(O.DEFAULT_CFG)
            off = dict(self.offset_dec.named_parameters()); prob = dict(self.prob_dec.named_parameters())
            data_dict.update(O.lidf_query(data_dict, cfg, off, prob, data_dict["part_size"], dedup_rays=True))

    lidf = _OracleLIDF(default_opt(), torch.device("cpu"), resnet_model=_Fixed(d["full_rgb_feat"]), pnet_model=_Fixed(d["occ_voxel_feat"]))
    ddp = torch.nn.parallel.DistributedDataParallel(lidf, find_unused_parameters=True)
    dd = dict(d); dd["total_miss_sample_num"] = d["miss_ray_dir"].shape[0]
    dd.update(rgb_img=torch.zeros(n, 3, 8, 8), valid_rgb=torch.zeros(4, 3), valid_v_pid=torch.zeros(4, dtype=torch.long),
              valid_v_rel_coord=torch.zeros(4, 3), revidx=torch.zeros(4, dtype=torch.long))
    lidf.train()
    out = ddp(dd, "test", 0)                                         # through DDP.forward so the reducer hooks are armed
    loss = out["pred_pos"].abs().mean() + out["pred_prob_end"].mean()
    loss.backward()
    grads = [p.grad.clone() for p in lidf.parameters() if p.grad is not None]
    flat = torch.cat([g.reshape(-1) for g in grads])
    # manual all-reduce of the same thing: every rank must hold identical (averaged) decoder grads afterwards
    gathered = [torch.zeros_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    same = all(torch.equal(gathered[0], g) for g in gathered)            # DDP averaged them: identical on every rank
    q.put((rank, n, float(flat.abs().sum()), bool(same)))
    counts = torch.tensor([float(d["occ_vox_intersect_idx"].shape[0])])
    dist.all_reduce(counts)
    q.put(("pairs", rank, float(counts)))
    dist.destroy_process_group()


def test_two_rank_sharding_by_image():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    got = [q.get(timeout=5) for _ in range(4)]
    shards = sorted(x[1] for x in got if x[0] in (0, 1))
    assert shards == [1, 2]                                   # 3 images over 2 ranks
    assert all(x[3] for x in got if x[0] in (0, 1))           # decoder gradients identical after DDP's all-reduce
    assert all(x[2] > 0 for x in got if x[0] in (0, 1))
    totals = {x[2] for x in got if x[0] == "pairs"}
    assert len(totals) == 1                                   # both ranks agree on the global pair count
