"""Generate golden vectors by executing the REFERENCE's own code (run in the build container only).

    python tests/golden/make_golden.py            # needs /root/reference (read-only mount)

What runs unmodified from /root/reference/src:
  models.implicit_net  (Embedder / get_embedder / IMNet / IEF)
  models.pipeline.LIDF.get_embedding + LIDF.get_pred            (pipeline.py:338-466)
  models.pipeline.RefineNet.get_pred_refine                     (pipeline.py:922-1030)
  torchvision.ops.roi_align (0.26 here; the reference pins 0.7.0, same aligned=True semantics)

What is stubbed, because it is absent from this container (no network):
  matplotlib / mpl_toolkits : empty modules (only used by visualisation code)
  torch_scatter             : a sequential numpy restatement of the 2.0.x CPU kernels
                              (scatter / scatter_max / scatter_softmax / scatter_log_softmax)
  extensions.{ray_aabb,pcl_aabb}.jit : CUDA-only (CHECK_CUDA); ray_aabb is not called by the hot path,
                              pcl_aabb.forward is restated in torch from pcl_aabb_cuda_kernel.cu:10-45
  resnet_model / pnet_model : replaced by modules that return the given full_rgb_feat /
                              occ_voxel_feat (they are upstream producers, not on the path)

Output: tests/golden/*.npz (inputs + reference outputs, fp32).
"""
import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF_SRC = "/root/reference/src"
sys.path.insert(0, REPO)

from implicit_depth_b200.synthetic import make_inputs  # noqa: E402
from oracle.lidf_oracle import init_decoder  # noqa: E402  (only the weight initialiser, no oracle math)


# ----------------------------------------------------------------------------- stubs
def _np_scatter_max(src, index, dim_size=None):
    s = src.detach().cpu().numpy(); idx = index.detach().cpu().numpy()
    n = int(idx.max()) + 1 if dim_size is None else int(dim_size)
    out = np.full((n,), -np.inf, dtype=s.dtype); arg = np.full((n,), s.shape[0], dtype=np.int64)
    for i in range(s.shape[0]):                     # sequential, strict compare: first max wins
        j = idx[i]
        if s[i] > out[j]:
            out[j] = s[i]; arg[j] = i
    out[arg == s.shape[0]] = 0
    return torch.from_numpy(out), torch.from_numpy(arg)


def _shim_scatter_max(src, index, dim=-1, out=None, dim_size=None):
    assert src.dim() == 1
    return _np_scatter_max(src, index, dim_size)


def _shim_scatter_softmax(src, index, dim=-1, eps=1e-12):
    mx, _ = _np_scatter_max(src, index)
    rec = (src - mx[index]).exp()
    s = torch.zeros_like(mx).index_add_(0, index, rec)
    return rec / (s + eps)[index]


def _shim_scatter_log_softmax(src, index, dim=-1, eps=1e-12):
    mx, _ = _np_scatter_max(src, index)
    rec = src - mx[index]
    s = torch.zeros_like(mx).index_add_(0, index, rec.exp())
    return rec - (s + eps).log()[index]


def _shim_scatter(src, index, dim=-1, out=None, dim_size=None, reduce="sum"):
    if reduce == "max" and out is not None and src.dim() == 1:      # pipeline.py:944 in-place max
        o = out.numpy(); s = src.numpy(); idx = index.numpy()
        for i in range(s.shape[0]):
            if s[i] > o[idx[i]]:
                o[idx[i]] = s[i]
        return out
    if reduce == "max" and out is None and dim == 0 and src.requires_grad:
        # differentiable variant (gradient goes to the arg-max row of every (voxel, channel), as in torch_scatter)
        n = int(index.max()) + 1 if dim_size is None else dim_size
        o = torch.full((n,) + tuple(src.shape[1:]), -float("inf"), dtype=src.dtype)
        o = o.scatter_reduce(0, index.reshape(-1, 1).expand_as(src), src, reduce="amax", include_self=True)
        return torch.where(torch.isinf(o), torch.zeros_like(o), o)
    if reduce == "max" and out is None and dim == 0:                # pointnet.py:27,35: per-voxel max of [N,C] rows
        n = int(index.max()) + 1 if dim_size is None else dim_size  # torch_scatter: rows without any source stay 0
        s = src.detach().numpy(); idx = index.numpy()
        o = np.full((n,) + s.shape[1:], -np.inf, dtype=s.dtype)
        np.maximum.at(o, idx, s)
        o[np.isneginf(o)] = 0
        return torch.from_numpy(o)
    if reduce == "sum":
        n = int(index.max()) + 1 if dim_size is None else dim_size
        return torch.zeros((n,) + src.shape[1:], dtype=src.dtype).index_add_(0, index, src)
    raise NotImplementedError(reduce)


def _pcl_aabb_forward(pcl_pos, voxel_bound, pcl_bid, voxel_bid):
    """torch restatement of pcl_aabb_cuda_kernel.cu:10-45 -> mask [V,N] int32."""
    p = pcl_pos.unsqueeze(0); lo = voxel_bound[:, None, :3]; hi = voxel_bound[:, None, 3:]
    inside = ((p >= lo) & (p <= hi)).all(-1) & (voxel_bid[:, None] == pcl_bid[None, :])
    return inside.int()


def install_stubs():
    for name in ("matplotlib", "matplotlib.pyplot", "mpl_toolkits", "mpl_toolkits.mplot3d"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["mpl_toolkits"].mplot3d = sys.modules["mpl_toolkits.mplot3d"]
    ts = types.ModuleType("torch_scatter")
    ts.scatter = _shim_scatter; ts.scatter_max = _shim_scatter_max
    ts.scatter_softmax = _shim_scatter_softmax; ts.scatter_log_softmax = _shim_scatter_log_softmax
    sys.modules["torch_scatter"] = ts
    for ext in ("ray_aabb", "pcl_aabb"):
        jit = types.ModuleType(f"extensions.{ext}.jit")
        obj = types.SimpleNamespace(forward=_pcl_aabb_forward if ext == "pcl_aabb" else None)
        setattr(jit, ext, obj)
        sys.modules[f"extensions.{ext}.jit"] = jit


class _Fixed(nn.Module):
    """Stands in for resnet_model / pnet_model: returns a fixed tensor, records the last kwargs."""

    def __init__(self, value):
        super().__init__()
        self.value = value
        self.last = None

    def forward(self, *a, **kw):
        self.last = (a, kw)
        return self.value


def build_reference(cfg_over):
    os.chdir(REF_SRC)
    sys.path.insert(0, REF_SRC)
    install_stubs()
    import models.pipeline as ref_pipeline          # the reference, unmodified
    from opt import Params
    opt = Params("experiments/implicit_depth/default_config.yaml")
    opt.update("experiments/implicit_depth/test_refine.yaml")     # superset of test_lidf.yaml (adds refine.*)
    opt.gpu_id = 0
    for k, v in cfg_over.items():
        sect, key = k.split(".")
        setattr(getattr(opt, sect), key, v)
    dev = torch.device("cpu")
    lidf = ref_pipeline.LIDF(opt, dev)
    refine = ref_pipeline.RefineNet(opt, dev)
    return opt, lidf, refine


def dense_dist(d):
    V = d["voxel_bound"].shape[0]; R = d["miss_ray_dir"].shape[0]
    dist = torch.zeros(V, R, 2)
    dist[d["occ_vox_intersect_idx"], d["miss_ray_intersect_idx"]] = d["intersect_dist"]
    return dist


def run_case(name, B, H, W, N, V_img, seed, init_mode, cfg_over, ragged, label_branch=False, with_refine=False,
             store_roi=True):
    torch.manual_seed(seed)
    opt, lidf, refine = build_reference(cfg_over)
    d = make_inputs(B, H, W, N, V_img=V_img, seed=seed, ragged=ragged)
    g = torch.Generator().manual_seed(seed + 7)
    off = init_decoder(opt.model.offdec_type, 385, mode=init_mode, generator=g)
    prob = init_decoder(opt.model.probdec_type, 385, mode=init_mode, generator=g)
    lidf.offset_dec.load_state_dict(off); lidf.prob_dec.load_state_dict(prob)
    lidf.resnet_model = _Fixed(d["full_rgb_feat"]); lidf.pnet_model = _Fixed(d["occ_voxel_feat"])
    lidf.eval()
    R = d["total_miss_sample_num"]
    P = d["occ_vox_intersect_idx"].shape[0]
    dd = dict(bs=B, h=H, w=W, dist=dense_dist(d), occ_vox_intersect_idx=d["occ_vox_intersect_idx"],
              miss_ray_intersect_idx=d["miss_ray_intersect_idx"], miss_ray_dir=d["miss_ray_dir"],
              miss_img_ind=d["miss_img_ind"], miss_bid=d["miss_bid"], voxel_bound=d["voxel_bound"],
              occ_vox_bid=d["occ_vox_bid"],
              rgb_img=torch.zeros(B, 3, H, W), valid_rgb=torch.zeros(4, 3), valid_v_pid=torch.zeros(4, dtype=torch.long),
              valid_v_rel_coord=torch.zeros(4, 3), revidx=torch.zeros(4, dtype=torch.long),
              part_size=d["part_size"], total_miss_sample_num=R, item_path=["synthetic"],
              miss_flat_img_id=(d["miss_img_ind"][:, 1] * W + d["miss_img_ind"][:, 0]))
    save = {k: (v.to(torch.int32) if v.dtype == torch.int64 else v).numpy()
            for k, v in d.items() if isinstance(v, torch.Tensor)}
    save.update(meta_B=B, meta_H=H, meta_W=W, meta_N=N, meta_V_img=V_img, part_size=np.float64(d["part_size"]),
                offdec_type=str(opt.model.offdec_type), n_iter=int(opt.model.n_iter),
                use_sigmoid=bool(opt.model.use_sigmoid), intersect_pos_type=str(opt.model.intersect_pos_type),
                pos_encode=bool(opt.model.pos_encode), multires=int(opt.model.multires),
                multires_views=int(opt.model.multires_views),
                offset_range=np.array(opt.grid.offset_range, dtype=np.float64))
    for k, v in off.items():
        save["offset_dec." + k] = v.numpy()
    for k, v in prob.items():
        save["prob_dec." + k] = v.numpy()
    with torch.no_grad():
        lidf.get_embedding(dd)                                   # reference code, unmodified
        if label_branch:
            lab = (torch.rand(P, generator=g) < 0.15).float()
            dd["pcl_label_float"] = lab
            save["pcl_label_float"] = lab.numpy()
            lidf.get_pred(dd, "train", 0)                        # epoch 0 < maxpool_label_epo (=6)
        else:
            lidf.get_pred(dd, "test", 0)
        for k in ("pair_pred_pos", "max_pair_id", "pred_prob_end", "pred_prob_end_softmax", "pred_pos"):
            save["ref." + k] = dd[k].numpy()
        if store_roi:   # per-pair in the reference; identical for all pairs of a ray -> keep one row per ray
            roi = torch.full((R, dd["intersect_rgb_feat"].shape[1]), float("nan"))
            roi[dd["miss_ray_intersect_idx"]] = dd["intersect_rgb_feat"]
            save["ref.roi_feat_per_ray"] = roi.numpy()
        # pred_offset is not stored by the reference; recover it from the module for the fixture
        inp_embed = torch.cat((dd["intersect_voxel_feat"], dd["intersect_rgb_feat"], dd["intersect_enter_pos_embed"],
                               dd["intersect_leave_pos_embed"], dd["intersect_dir_embed"]), -1)
        save["ref.pred_offset"] = lidf.offset_dec(inp_embed).numpy()
        if with_refine:
            roff = init_decoder(opt.refine.offdec_type, 334, mode=init_mode, generator=g)
            refine.offset_dec.load_state_dict(roff)
            for k, v in roff.items():
                save["refine_dec." + k] = v.numpy()
            vfeat2 = torch.relu(torch.randn(d["occ_voxel_feat"].shape, generator=g))
            pn = _Fixed(vfeat2); refine.pnet_model = pn
            refine.eval()
            opt.mask_type = "all"
            dd["valid_mask"] = torch.ones(B, H, W)
            out = refine.get_pred_refine(dd, dd["pred_pos"], "test", 0)      # reference code, unmodified
            end_voxel_id = pn.last[1]["vox2point_idx"][-R:]                  # final_revidx tail, pipeline.py:1010
            save["refine.occ_voxel_feat"] = vfeat2.numpy()
            save["refine.end_voxel_id"] = end_voxel_id.numpy()
            save["ref.pred_pos_refine"] = out.numpy()
            save["refine.offset_range"] = np.array(opt.refine.offset_range, dtype=np.float64)
            save["refine.n_iter"] = int(opt.refine.n_iter)
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **save)
    print(f"{name}: P={P} R={R} -> {path} ({os.path.getsize(path) / 1e6:.2f} MB)")


if __name__ == "__main__":
    assert os.path.isdir(REF_SRC), "run in the build container: /root/reference must be mounted"
    # c1: BASELINE config 1 -- single 64x64 image, 16 pairs/ray, random-init IMNet decoders
    run_case("c1_imnet_64x64x16", 1, 64, 64, 16, 256, 1234, "reference",
             {"model.offdec_type": "IMNET"}, ragged=False, store_roi=False)
    # shipped YAML decoders (IEF n_iter 2 + IMNet), trained-like weights, ragged pairs incl. empty rays, + refine
    run_case("ief_ragged_2x24x32", 2, 24, 32, 12, 64, 4321, "trained", {}, ragged=True, with_refine=True)
    # option coverage: rel positions, sigmoid head, IEF n_iter 3, GT-label arg-max branch
    run_case("ief_rel_sigmoid_1x16x20", 1, 16, 20, 6, 40, 99, "trained",
             {"model.intersect_pos_type": "rel", "model.use_sigmoid": True, "model.n_iter": 3},
             ragged=True, label_branch=True)
