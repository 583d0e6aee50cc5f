import os
import sys

import numpy as np
import pytest
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

GOLDEN_DIR = os.path.join(REPO, "tests", "golden")
GOLDEN_CASES = ["c1_imnet_64x64x16", "ief_ragged_2x24x32", "ief_rel_sigmoid_1x16x20"]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


def load_golden(name):
    """Load a fixture written by tests/golden/make_golden.py -> (inputs dict, cfg, decoders, reference outputs)."""
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)
    t = {k: torch.from_numpy(z[k]) for k in z.files if z[k].ndim > 0}
    d = {}
    for k in ("full_rgb_feat", "occ_voxel_feat", "voxel_bound", "miss_ray_dir", "intersect_dist"):
        d[k] = t[k]
    for k in ("occ_vox_bid", "miss_bid", "miss_img_ind", "occ_vox_intersect_idx", "miss_ray_intersect_idx"):
        d[k] = t[k].long()
    cfg = dict(pos_encode=bool(z["pos_encode"]), multires=int(z["multires"]), multires_views=int(z["multires_views"]),
               intersect_pos_type=str(z["intersect_pos_type"]), roi_inp_bbox=8, roi_out_bbox=2,
               offdec_type=str(z["offdec_type"]), n_iter=int(z["n_iter"]), probdec_type="IMNET",
               use_sigmoid=bool(z["use_sigmoid"]), offset_range=tuple(float(v) for v in z["offset_range"]),
               scatter_type="Maxpool")
    off = {k[len("offset_dec."):]: v for k, v in t.items() if k.startswith("offset_dec.")}
    prob = {k[len("prob_dec."):]: v for k, v in t.items() if k.startswith("prob_dec.")}
    ref = {k[len("ref."):]: v for k, v in t.items() if k.startswith("ref.")}
    ref["max_pair_id"] = ref["max_pair_id"].long()
    extra = {k: v for k, v in t.items() if k.startswith("refine") or k == "pcl_label_float"}
    if "refine.n_iter" in z.files:
        extra["refine.n_iter"] = int(z["refine.n_iter"])
    return d, cfg, off, prob, float(z["part_size"]), ref, extra


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    """max |a-b| / max(|b|, rms(b)): relative error with a per-tensor scale floor (hard part 1)."""
    a = a.double().reshape(-1); b = b.double().reshape(-1)
    if b.numel() == 0:
        return 0.0
    scale = b.pow(2).mean().sqrt().clamp_min(1e-30)
    return float(((a - b).abs() / torch.maximum(b.abs(), scale)).max())
