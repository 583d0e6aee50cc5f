"""Loader for the UNMODIFIED reference tree (test / bench infrastructure, like the rest of oracle/).

Finds the reference's ``src`` directory -- ``baseline/_ref/src`` (a git-ignored copy made by
``__graft_entry__.build()`` in the build container so that it travels to the GPU box with the snapshot) or
``/root/reference/src`` -- puts it on ``sys.path`` and imports ``models.pipeline`` as it is.  Absent third-party modules are
stubbed, exactly as tests/golden/make_golden.py does on the CPU, but device-generic so the reference runs on the GPU:

  matplotlib / mpl_toolkits           empty modules (visualisation only)
  torch_scatter                       torch restatement of the 2.0.x semantics (scatter_reduce / index_add): empty segment ->
                                      value 0 / arg = len(src), first maximum wins, soft-max eps 1e-12; differentiable
  extensions.{ray_aabb,pcl_aabb}.jit  this repo's drop-in modules (same names, ``forward`` signatures and results as the
                                      reference's JIT-built CUDA extensions; bit-exact per tests/test_gpu_aabb.py)

Nothing in the product imports this file.
"""
from __future__ import annotations

import os
import sys
import types

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CANDIDATES = (os.path.join(REPO, "baseline", "_ref", "src"), "/root/reference/src")


def find_ref_src():
    for c in CANDIDATES:
        if os.path.isfile(os.path.join(c, "models", "pipeline.py")):
            return c
    return None


# ----------------------------------------------------------------------------- torch_scatter shim (any device)
def _lowest(dtype):
    return float("-inf") if dtype.is_floating_point else torch.iinfo(dtype).min


def scatter_max(src, index, dim=-1, out=None, dim_size=None):
    assert src.dim() == 1
    n = int(index.max()) + 1 if dim_size is None else int(dim_size)
    P = src.shape[0]
    mx = torch.full((n,), _lowest(src.dtype), dtype=src.dtype, device=src.device).scatter_reduce(0, index, src.detach(), "amax")
    pos = torch.arange(P, device=src.device)
    cand = torch.where(src.detach() == mx[index], pos, torch.full_like(pos, P))
    arg = torch.full((n,), P, dtype=torch.long, device=src.device).scatter_reduce(0, index, cand, "amin")
    val = torch.where(arg < P, mx, torch.zeros_like(mx))
    return val, arg


def _seg_max(src, index):
    n = int(index.max()) + 1
    return torch.full((n,), float("-inf"), dtype=src.dtype, device=src.device).scatter_reduce(0, index, src.detach(), "amax")


def scatter_softmax(src, index, dim=-1, eps=1e-12):
    rec = (src - _seg_max(src, index)[index]).exp()
    s = torch.zeros(int(index.max()) + 1, dtype=src.dtype, device=src.device).index_add(0, index, rec)
    return rec / (s + eps)[index]


def scatter_log_softmax(src, index, dim=-1, eps=1e-12):
    rec = src - _seg_max(src, index)[index]
    s = torch.zeros(int(index.max()) + 1, dtype=src.dtype, device=src.device).index_add(0, index, rec.exp())
    return rec - (s + eps).log()[index]


def scatter(src, index, dim=-1, out=None, dim_size=None, reduce="sum"):
    if reduce == "max" and out is not None and src.dim() == 1:       # pipeline.py:944: in-place running maximum
        return out.scatter_reduce_(0, index, src, "amax", include_self=True)
    n = int(index.max()) + 1 if dim_size is None else int(dim_size)
    if reduce == "max":                                               # pointnet.py:27,35: per-voxel max of [N,C] rows
        o = torch.full((n,) + tuple(src.shape[1:]), float("-inf"), dtype=src.dtype, device=src.device)
        idx = index.reshape(-1, *([1] * (src.dim() - 1))).expand_as(src)
        o = o.scatter_reduce(0, idx, src, "amax", include_self=True)
        return torch.where(torch.isinf(o), torch.zeros_like(o), o)    # rows without any source stay 0
    if reduce == "sum":
        return torch.zeros((n,) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device).index_add(0, index, src)
    raise NotImplementedError(reduce)


def install_stubs(use_native_extensions: bool = True):
    for name in ("matplotlib", "matplotlib.pyplot", "mpl_toolkits", "mpl_toolkits.mplot3d"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["mpl_toolkits"].mplot3d = sys.modules["mpl_toolkits.mplot3d"]
    ts = types.ModuleType("torch_scatter")
    ts.scatter, ts.scatter_max, ts.scatter_softmax, ts.scatter_log_softmax = scatter, scatter_max, scatter_softmax, scatter_log_softmax
    sys.modules["torch_scatter"] = ts
    if use_native_extensions:
        from implicit_depth_b200.extensions.pcl_aabb.jit import pcl_aabb
        from implicit_depth_b200.extensions.ray_aabb.jit import ray_aabb
        for name, obj in (("ray_aabb", ray_aabb), ("pcl_aabb", pcl_aabb)):
            jit = types.ModuleType(f"extensions.{name}.jit")
            setattr(jit, name, obj)
            sys.modules[f"extensions.{name}.jit"] = jit


def load(overrides=None, yamls=("test_lidf.yaml",), gpu_id=0):
    """-> (reference ``models.pipeline`` module, opt) with default_config.yaml + the given override YAMLs applied."""
    src = find_ref_src()
    if src is None:
        raise FileNotFoundError("reference tree not found (baseline/_ref/src or /root/reference/src)")
    if src not in sys.path:
        sys.path.insert(0, src)
    install_stubs()
    import models.pipeline as ref_pipeline           # the reference, unmodified
    torch.autograd.set_detect_anomaly(False)         # models/implicit_net.py:2 switches it on process-wide as an import side effect
    from opt import Params
    cfg_dir = os.path.join(src, "experiments", "implicit_depth")
    opt = Params(os.path.join(cfg_dir, "default_config.yaml"))
    for y in yamls:
        opt.update(os.path.join(cfg_dir, y))
    opt.gpu_id = gpu_id
    for k, v in (overrides or {}).items():
        sect, key = k.split(".")
        setattr(getattr(opt, sect), key, v)
    return ref_pipeline, opt


def synthetic_batch(B, H, W, seed=0, device="cpu"):
    """A batch dict with the keys of the reference's datasets (src/datasets/cleargrasp_synthetic_dataset.py:229-245):
    a tilted plane with a bump, a rectangular 'transparent' region whose depth is missing in the corrupt inputs."""
    import math
    g = torch.Generator().manual_seed(seed)
    fx = 0.5 * W / math.tan(0.5 * 1.2112585); fy = fx; cx, cy = W / 2.0, H / 2.0
    ys, xs = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32), indexing="ij")
    depth = torch.stack([0.9 + 0.2 * b + 0.25 * (xs / W) + 0.15 * (ys / H)
                         + 0.08 * torch.exp(-(((xs - W / 2) / (0.2 * W)) ** 2 + ((ys - H / 2) / (0.2 * H)) ** 2)) for b in range(B)])
    depth = depth + 0.002 * torch.randn(B, H, W, generator=g)
    xyz = torch.stack(((xs - cx) / fx * depth, (ys - cy) / fy * depth, depth), 1)               # [B,3,H,W]
    corrupt = torch.zeros(B, H, W)
    corrupt[:, H // 4: 3 * H // 4, W // 3: 2 * W // 3] = 1.0
    depth_c = depth * (1 - corrupt)
    xyz_c = xyz * (1 - corrupt).unsqueeze(1)
    batch = dict(rgb=torch.rand(B, 3, H, W, generator=g), xyz=xyz, xyz_corrupt=xyz_c, depth=depth.unsqueeze(1),
                 depth_corrupt=depth_c.unsqueeze(1), corrupt_mask=corrupt.unsqueeze(1), valid_mask=(1 - corrupt).unsqueeze(1),
                 fx=torch.full((B,), fx), fy=torch.full((B,), fy), cx=torch.full((B,), cx), cy=torch.full((B,), cy),
                 item_path=[f"synthetic_{i}" for i in range(B)])
    return {k: (v.to(device) if torch.is_tensor(v) else v) for k, v in batch.items()}
