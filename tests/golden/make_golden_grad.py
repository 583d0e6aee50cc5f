"""Golden GRADIENTS of the LIDF query path, from the REFERENCE's own code under autograd (build container only).

    python tests/golden/make_golden_grad.py       # needs /root/reference; writes tests/golden/grad_*.npz

Runs the reference's unmodified ``LIDF.get_embedding`` + ``LIDF.get_pred`` (src/models/pipeline.py:338-466) in training mode
on an existing fixture's inputs with ``full_rgb_feat`` / ``occ_voxel_feat`` and all decoder parameters requiring grad, forms
the scalar  L = sum(c_pos * pred_pos) + sum(c_prob * pred_prob_end)  with seeded coefficient tensors (the two outputs the
reference's losses differentiate, pipeline.py:472,482), calls backward, and stores every gradient.  This pins the native
backward (lidf_query_backward) and autograd through the oracle.

Two fixtures per case:
  grad_<case>.npz   coefficients as drawn.
  gradk_<case>.npz  the same coefficients, ZEROED on every pair that has a pre-activation closer than KINK_THR to a
                    leaky-ReLU kink (or to a kink of the output clamp) in the reference's own fp32 forward, and on rays whose
                    arg-max pair is such a pair.  The sign of such a pre-activation -- a factor 50 in the local derivative --
                    is decided by rounding noise, so its gradient differs between ANY two implementations / precisions /
                    summation orders (the reference's own fp32 vs fp64 gradients differ by up to 9e-2 on these inputs);
                    on the remaining pairs the gradient is well conditioned and must match to 1e-3.  The margins come from
                    forward hooks on the reference modules' nn.Linear layers (no reference code is modified).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import make_golden as MG  # noqa: E402


KINK_THR = 1e-4


def reference_kink_margin(lidf, dd, use_sigmoid):
    """min over layers / IEF iterations / both decoders of |pre-activation| per pair, from forward hooks on the reference's
    own modules (linear_1..3 feed F.leaky_relu(inplace=True): the hook sees the value before it is overwritten)."""
    P = dd["occ_vox_intersect_idx"].shape[0]
    margin = torch.full((P,), float("inf"))
    handles, acc = [], {}

    def hook_act(_m, _i, out):
        margin.copy_(torch.minimum(margin, out.detach().abs().min(1).values))

    def hook_l4(name):
        def h(_m, _i, out):
            acc[name] = acc.get(name, 0.0) + out.detach()[:, 0]
        return h
    for name, dec in (("offset_dec", lidf.offset_dec), ("prob_dec", lidf.prob_dec)):
        for lin in (dec.linear_1, dec.linear_2, dec.linear_3):
            handles.append(lin.register_forward_hook(hook_act))
        handles.append(dec.linear_4.register_forward_hook(hook_l4(name)))
    with torch.no_grad():
        d2 = dict(dd)
        lidf.get_embedding(d2)
        lidf.get_pred(d2, "train", 100)
    for h in handles:
        h.remove()
    if not use_sigmoid:                                   # leaky clamp max(min(x, .01x+.99), .01x): kinks at 0 and 1
        for name, dec in (("offset_dec", lidf.offset_dec), ("prob_dec", lidf.prob_dec)):
            x = acc[name] + (float(dec.init_offset) if hasattr(dec, "init_offset") else 0.0)
            margin = torch.minimum(margin, torch.minimum(x.abs(), (x - 1).abs()))
    return margin, d2["max_pair_id"]


def run(src_name, out_name, seed, mask_kinks=False):
    z = np.load(os.path.join(HERE, src_name + ".npz"))
    over = {"model.offdec_type": str(z["offdec_type"]), "model.n_iter": int(z["n_iter"]),
            "model.use_sigmoid": bool(z["use_sigmoid"]), "model.intersect_pos_type": str(z["intersect_pos_type"])}
    opt, lidf, _ = MG.build_reference(over)
    t = {k: torch.from_numpy(z[k]) for k in z.files if z[k].ndim > 0}
    off = {k[len("offset_dec."):]: v for k, v in t.items() if k.startswith("offset_dec.")}
    prob = {k[len("prob_dec."):]: v for k, v in t.items() if k.startswith("prob_dec.")}
    lidf.offset_dec.load_state_dict(off); lidf.prob_dec.load_state_dict(prob)
    feat = t["full_rgb_feat"].clone().requires_grad_(True)
    vfeat = t["occ_voxel_feat"].clone().requires_grad_(True)
    lidf.resnet_model = MG._Fixed(feat); lidf.pnet_model = MG._Fixed(vfeat)
    lidf.train()
    B, H, W = int(z["meta_B"]), int(z["meta_H"]), int(z["meta_W"])
    d = {k: t[k] for k in ("voxel_bound", "miss_ray_dir", "intersect_dist")}
    for k in ("occ_vox_bid", "miss_bid", "miss_img_ind", "occ_vox_intersect_idx", "miss_ray_intersect_idx"):
        d[k] = t[k].long()
    R = d["miss_ray_dir"].shape[0]
    dd = dict(bs=B, h=H, w=W, dist=MG.dense_dist(d), occ_vox_intersect_idx=d["occ_vox_intersect_idx"],
              miss_ray_intersect_idx=d["miss_ray_intersect_idx"], miss_ray_dir=d["miss_ray_dir"], miss_img_ind=d["miss_img_ind"],
              miss_bid=d["miss_bid"], voxel_bound=d["voxel_bound"], occ_vox_bid=d["occ_vox_bid"],
              rgb_img=torch.zeros(B, 3, H, W), valid_rgb=torch.zeros(4, 3), valid_v_pid=torch.zeros(4, dtype=torch.long),
              valid_v_rel_coord=torch.zeros(4, 3), revidx=torch.zeros(4, dtype=torch.long),
              part_size=float(z["part_size"]), total_miss_sample_num=R, item_path=["synthetic"])
    keep = None
    if mask_kinks:
        margin, mp = reference_kink_margin(lidf, dd, bool(z["use_sigmoid"]))
        keep = margin > KINK_THR
    lidf.get_embedding(dd)                                   # reference code, unmodified, autograd recording
    lidf.get_pred(dd, "train", 100)                          # epoch >= maxpool_label_epo: arg-max of the predicted soft-max
    g = torch.Generator().manual_seed(seed)
    c_pos = torch.randn(dd["pred_pos"].shape, generator=g)
    c_prob = torch.randn(dd["pred_prob_end"].shape, generator=g)
    if keep is not None:
        P = keep.shape[0]
        mp = dd["max_pair_id"]
        ray_keep = torch.where(mp < P, keep[mp.clamp(max=P - 1)], torch.ones_like(mp, dtype=torch.bool))
        c_pos = c_pos * ray_keep.float().unsqueeze(-1)
        c_prob = c_prob * keep.float().unsqueeze(-1)
    loss = (c_pos * dd["pred_pos"]).sum() + (c_prob * dd["pred_prob_end"]).sum()
    loss.backward()
    save = {"c_pos": c_pos.numpy(), "c_prob": c_prob.numpy(), "loss": np.float64(float(loss)),
            "grad.full_rgb_feat": feat.grad.numpy(), "grad.occ_voxel_feat": vfeat.grad.numpy(),
            "max_pair_id": dd["max_pair_id"].numpy().astype(np.int32)}
    if keep is not None:
        save["kink_keep"] = keep.numpy(); save["kink_thr"] = np.float64(KINK_THR)
        print(out_name, "pairs kept", float(keep.float().mean()))
    for name, mod in (("offset_dec", lidf.offset_dec), ("prob_dec", lidf.prob_dec)):
        for k, p in mod.named_parameters():
            save[f"grad.{name}.{k}"] = p.grad.numpy()
    path = os.path.join(HERE, out_name + ".npz")
    np.savez_compressed(path, **save)
    print(out_name, "loss", float(loss), "|g feat|", float(feat.grad.abs().sum()), "|g vfeat|", float(vfeat.grad.abs().sum()),
          os.path.getsize(path), "bytes")


if __name__ == "__main__":
    run("ief_ragged_2x24x32", "grad_ief_ragged_2x24x32", 41)
    run("ief_rel_sigmoid_1x16x20", "grad_ief_rel_sigmoid_1x16x20", 42)
    run("ief_ragged_2x24x32", "gradk_ief_ragged_2x24x32", 41, mask_kinks=True)
    run("ief_rel_sigmoid_1x16x20", "gradk_ief_rel_sigmoid_1x16x20", 42, mask_kinks=True)
