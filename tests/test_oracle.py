"""CPU: pin the oracle (oracle/lidf_oracle.py) against outputs of the reference's own code
(tests/golden/*.npz, made by tests/golden/make_golden.py) and against torchvision's roi_align."""
import pytest
import torch

from conftest import GOLDEN_CASES, load_golden, rel_err
from oracle import lidf_oracle as O

TOL = 2e-5   # fp32 oracle vs fp32 reference: same ops, only summation-order noise


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_oracle_matches_reference_outputs(name):
    d, cfg, off, prob, part, ref, extra = load_golden(name)
    out = O.lidf_query(d, cfg, off, prob, part, dedup_rays=True, pcl_label_float=extra.get("pcl_label_float"))
    for k in ("pred_offset", "pred_prob_end", "pair_pred_pos", "pred_prob_end_softmax", "pred_pos"):
        assert out[k].shape == ref[k].shape, k
        assert rel_err(out[k], ref[k]) < TOL, (k, rel_err(out[k], ref[k]))
    assert torch.equal(out["max_pair_id"], ref["max_pair_id"])
    if "roi_feat_per_ray" in ref:
        ray = d["miss_ray_intersect_idx"]
        assert rel_err(out["intersect_rgb_feat"], ref["roi_feat_per_ray"][ray]) < TOL


def test_oracle_per_pair_roi_equals_per_ray_roi():
    d, cfg, off, prob, part, ref, _ = load_golden("ief_rel_sigmoid_1x16x20")
    a = O.get_embedding(d, cfg, dedup_rays=False)["intersect_rgb_feat"]
    b = O.get_embedding(d, cfg, dedup_rays=True)["intersect_rgb_feat"]
    assert torch.equal(a, b)


def test_fp64_oracle_error_budget():
    """fp32 reference vs fp64 oracle: the reference itself is only ~1e-5 accurate; the 1e-3 bar has room."""
    d, cfg, off, prob, part, ref, _ = load_golden("ief_ragged_2x24x32")
    out = O.lidf_query(O.cast_tree(d, torch.float64), cfg, O.cast_tree(off, torch.float64),
                       O.cast_tree(prob, torch.float64), part, dedup_rays=True)
    for k in ("pred_offset", "pred_prob_end", "pair_pred_pos"):
        assert rel_err(ref[k], out[k]) < 1e-4, k


def test_roi_align_restatement_vs_torchvision():
    tv = pytest.importorskip("torchvision.ops")
    g = torch.Generator().manual_seed(0)
    for (B, C, H, W) in [(2, 5, 13, 17), (1, 3, 6, 7), (1, 2, 3, 4)]:
        feat = torch.randn(B, C, H, W, generator=g)
        ys, xs = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
        pix = torch.stack((xs.reshape(-1), ys.reshape(-1)), -1)
        for b in range(B):
            ul = pix - 4; br = pix + 4
            ul = torch.stack((ul[:, 0].clamp(0, W - 1), ul[:, 1].clamp(0, H - 1)), -1)
            br = torch.stack((br[:, 0].clamp(0, W - 1), br[:, 1].clamp(0, H - 1)), -1)
            boxes = torch.cat((torch.full((pix.shape[0], 1), b), ul, br), -1).float()
            want = tv.roi_align(feat, boxes, output_size=2, spatial_scale=1.0, aligned=True)
            got = O.roi_align_aligned(feat, boxes, 2, 1.0)
            assert torch.allclose(got, want, atol=1e-5, rtol=1e-5)
    # fractional boxes too (not produced by the reference, but the restatement is general)
    feat = torch.randn(1, 4, 20, 24, generator=g)
    xy = torch.rand(64, 2, generator=g) * torch.tensor([20.0, 16.0])
    wh = torch.rand(64, 2, generator=g) * 7.5
    boxes = torch.cat((torch.zeros(64, 1), xy, xy + wh), -1)
    want = tv.roi_align(feat, boxes, output_size=2, spatial_scale=1.0, aligned=True)
    assert torch.allclose(O.roi_align_aligned(feat, boxes, 2, 1.0), want, atol=1e-5, rtol=1e-5)


def test_scatter_semantics():
    src = torch.tensor([1.0, 3.0, 3.0, -2.0, 0.5])
    idx = torch.tensor([0, 0, 0, 2, 2])
    mx, arg = O.scatter_max(src, idx, dim_size=4)
    assert mx.tolist() == [3.0, 0.0, 0.5, 0.0]
    assert arg.tolist() == [1, 5, 4, 5]          # first max wins; empty segment -> src.numel()
    sm = O.scatter_softmax(src, idx)
    assert abs(float(sm[:3].sum()) - 1) < 1e-6 and abs(float(sm[3:].sum()) - 1) < 1e-6


def test_embed_layout_and_dims():
    x = torch.tensor([[0.1, -0.2, 0.3]])
    e = O.embed(x, 8)
    assert e.shape == (1, 51) and O.embed_out_dim(8) == 51 and O.embed_out_dim(4) == 27
    assert torch.allclose(e[0, 3:6], torch.sin(x[0])) and torch.allclose(e[0, 6:9], torch.cos(x[0]))
    assert torch.allclose(e[0, 45:48], torch.sin(128 * x[0])) and torch.allclose(e[0, 48:51], torch.cos(128 * x[0]))
    assert O.embed(x, 8, enabled=False) is x


def test_refine_tail_matches_reference():
    d, cfg, off, prob, part, ref, extra = load_golden("ief_ragged_2x24x32")
    rdec = {k[len("refine_dec."):]: v for k, v in extra.items() if k.startswith("refine_dec.")}
    evid = extra["refine.end_voxel_id"].long()
    vb = d["voxel_bound"][evid]
    center = (vb[:, :3] + vb[:, 3:]) / 2
    rgb = O.roi_align_aligned(d["full_rgb_feat"], torch.cat((d["miss_bid"].unsqueeze(-1).float(),
                              (d["miss_img_ind"] - 4).clamp(min=0).float(),
                              torch.stack(((d["miss_img_ind"][:, 0] + 4).clamp(max=d["full_rgb_feat"].shape[3] - 1),
                                           (d["miss_img_ind"][:, 1] + 4).clamp(max=d["full_rgb_feat"].shape[2] - 1)), -1).float()), -1))
    rcfg = dict(O.REFINE_CFG, offset_range=tuple(float(v) for v in extra["refine.offset_range"]),
                n_iter=extra["refine.n_iter"])
    out = O.refine_decoder_tail(ref["pred_pos"], d["miss_ray_dir"], center, extra["refine.occ_voxel_feat"][evid],
                                rgb.reshape(rgb.shape[0], -1), rcfg, rdec)
    assert rel_err(out, ref["pred_pos_refine"]) < TOL


LOSS_CASES = ["loss_ief_ragged_2x24x32", "loss_c1_imnet_64x64x16"]


def load_loss(name):
    import os
    import numpy as np
    from conftest import GOLDEN_DIR
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    t = {k: torch.from_numpy(z[k]) for k in z.files if z[k].ndim > 0}
    t["miss_ray_intersect_idx"] = t["miss_ray_intersect_idx"].long(); t["pcl_label"] = t["pcl_label"].long()
    sc = {k: float(z[k]) for k in ("ref_pos_loss", "ref_prob_loss", "ref_acc", "ref_err", "ref_surf_norm_loss", "ref_smooth_loss",
                                   "ref_angle_err", "ref_loss_net", "B", "H", "W")}
    t["miss_bid"] = t["miss_bid"].long(); t["miss_flat_img_id"] = t["miss_flat_img_id"].long()
    return t, sc, int(z["R"])


@pytest.mark.parametrize("name", LOSS_CASES)
def test_ray_loss_oracle_matches_reference_compute_loss(name):
    """pos_loss / prob_loss / acc / err from the reference's own compute_loss (tests/golden/make_golden_loss.py)."""
    t, sc, R = load_loss(name)
    out = O.ray_loss_stats(t["pred_prob_end"], t["pred_prob_end_softmax"], t["miss_ray_intersect_idx"], t["pcl_label"], R,
                           t["pred_pos"], t["gt_pos"])
    assert torch.equal(out["pred_label"], t["ref_pred_label"]) and torch.equal(out["gt_label"], t["ref_gt_label"])
    assert rel_err(out["log_softmax"], t["ref_log_softmax"]) < 1e-6
    for k in ("pos_loss", "prob_loss", "acc", "err"):
        assert abs(float(out[k]) - sc["ref_" + k]) <= 2e-6 * max(1.0, abs(sc["ref_" + k])), k


POINTNET_CASES = ["pointnet_2000x60", "pointnet_257x3"]


def load_pointnet(name):
    import os
    import numpy as np
    from conftest import GOLDEN_DIR
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    w = {k[2:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("w.")}
    return w, torch.from_numpy(z["inp_feat"]), torch.from_numpy(z["vox2point_idx"]).long(), int(z["V"]), torch.from_numpy(z["ref_out"])


@pytest.mark.parametrize("name", POINTNET_CASES)
def test_pointnet_oracle_matches_reference_module(name):
    w, inp, idx, V, ref = load_pointnet(name)
    out = O.pointnet2stage_forward(w, inp, idx, V)
    assert out.shape == ref.shape == (V, 128)
    assert rel_err(out, ref) < 2e-6


@pytest.mark.parametrize("name", LOSS_CASES)
def test_image_loss_oracle_matches_reference_compute_loss(name):
    """surf_norm_loss / smooth_loss / angle_err and both normal images from the reference's own compute_loss."""
    t, sc, R = load_loss(name)
    out = O.image_loss_stats(t["xyz_flat"], t["miss_bid"], t["miss_flat_img_id"], t["pred_pos"], t["gt_pos"],
                             int(sc["B"]), int(sc["H"]), int(sc["W"]))
    assert torch.equal(out["pred_surf_norm_img"], t["ref_pred_surf_norm_img"]) and torch.equal(out["gt_surf_norm_img"], t["ref_gt_surf_norm_img"])
    for k in ("surf_norm_loss", "smooth_loss", "angle_err"):
        assert abs(float(out[k]) - sc["ref_" + k]) <= 2e-6 * max(1.0, abs(sc["ref_" + k])), k


@pytest.mark.parametrize("name,gname", [("ief_ragged_2x24x32", "grad_ief_ragged_2x24x32"),
                                        ("ief_rel_sigmoid_1x16x20", "grad_ief_rel_sigmoid_1x16x20")])
def test_oracle_is_differentiable_and_matches_reference_gradients(name, gname):
    """Autograd through the oracle restatement reproduces the reference's own gradients (tests/golden/make_golden_grad.py):
    the yardstick for the native backward planned in DESIGN.md section 9."""
    import os
    import numpy as np
    from conftest import GOLDEN_DIR
    d, cfg, off, prob, part, ref, _ = load_golden(name)
    z = np.load(os.path.join(GOLDEN_DIR, gname + ".npz"))
    d = dict(d)
    d["full_rgb_feat"] = d["full_rgb_feat"].clone().requires_grad_(True)
    d["occ_voxel_feat"] = d["occ_voxel_feat"].clone().requires_grad_(True)
    off = {k: v.clone().requires_grad_(True) for k, v in off.items()}
    prob = {k: v.clone().requires_grad_(True) for k, v in prob.items()}
    out = O.lidf_query(d, cfg, off, prob, part, dedup_rays=True)
    loss = (torch.from_numpy(z["c_pos"]) * out["pred_pos"]).sum() + (torch.from_numpy(z["c_prob"]) * out["pred_prob_end"]).sum()
    loss.backward()
    assert rel_err(d["full_rgb_feat"].grad, torch.from_numpy(z["grad.full_rgb_feat"])) < 1e-4
    assert rel_err(d["occ_voxel_feat"].grad, torch.from_numpy(z["grad.occ_voxel_feat"])) < 1e-4
    for mod_name, params in (("offset_dec", off), ("prob_dec", prob)):
        for k, p in params.items():
            assert rel_err(p.grad, torch.from_numpy(z[f"grad.{mod_name}.{k}"])) < 1e-4, (mod_name, k)


def test_factored_layer1_backward_algebra():
    """with linear_1 applied to cat(voxel[vox], roi[ray], PE(pos), PE(dir)[ray]), the gradients of
    the per-voxel / per-ray blocks are segment sums of delta1 followed by small GEMMs.  Checked against autograd in fp64."""
    g = torch.Generator().manual_seed(11)
    P, R, V = 500, 60, 17
    vox = torch.randint(0, V, (P,), generator=g); ray = torch.randint(0, R, (P,), generator=g)
    vf = torch.randn(V, 128, generator=g, dtype=torch.float64, requires_grad=True)
    roi = torch.randn(R, 128, generator=g, dtype=torch.float64, requires_grad=True)
    pdir = torch.randn(R, 27, generator=g, dtype=torch.float64)
    ppos = torch.randn(P, 102, generator=g, dtype=torch.float64)
    W1 = torch.randn(256, 385, generator=g, dtype=torch.float64, requires_grad=True)
    x = torch.cat((vf[vox], roi[ray], ppos, pdir[ray]), -1)                       # pipeline.py:431-433 column order
    z1 = x @ W1.t()
    delta1 = torch.randn(P, 256, generator=g, dtype=torch.float64)                # dL/dz1 arriving from layer 2
    (z1 * delta1).sum().backward()
    Gv = torch.zeros(V, 256, dtype=torch.float64).index_add_(0, vox, delta1)      # per-voxel segment sum
    Gr = torch.zeros(R, 256, dtype=torch.float64).index_add_(0, ray, delta1)      # per-ray segment sum (CSR segments)
    dW1 = torch.cat((Gv.t() @ vf.detach(), Gr.t() @ roi.detach(), delta1.t() @ ppos, Gr.t() @ pdir), 1)
    assert torch.allclose(dW1, W1.grad, rtol=1e-10, atol=1e-10)
    assert torch.allclose(Gv @ W1.detach()[:, :128], vf.grad, rtol=1e-10, atol=1e-10)
    assert torch.allclose(Gr @ W1.detach()[:, 128:256], roi.grad, rtol=1e-10, atol=1e-10)


@pytest.mark.parametrize("name", ["metrics_bs1_64x64", "metrics_bs2_24x32"])
def test_depth_metrics_oracle_matches_reference_compute_loss(name):
    """The nine depth metrics of compute_loss(..., 'test', ...) (pipeline.py:570-618), both batch-size branches, against
    fixtures produced by the reference's unmodified code (make_golden_loss.py: run_metrics) -- incl. the bs == 1 branch's
    cv2.resize(INTER_NEAREST) round trip, restated as an index pick."""
    import os
    import numpy as np
    from conftest import GOLDEN_DIR
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    t = {k: torch.from_numpy(z[k]) for k in z.files if z[k].ndim > 0}
    B, H, W = int(z["B"]), int(z["H"]), int(z["W"])
    if B == 1:
        got = O.depth_metrics_image(t["xyz_flat"], t["xyz_corrupt_flat"], t["corrupt_mask"], t["miss_flat_img_id"].long(), t["pred_pos"], H, W)
    else:
        got = O.depth_metrics_rays(t["pred_pos"], t["gt_pos"])
    for k in ("a1", "a2", "a3", "rmse", "rmse_log", "log10", "abs_rel", "mae", "sq_rel"):
        assert abs(float(got[k]) - float(z["ref_" + k])) <= 2e-6 * max(1.0, abs(float(z["ref_" + k]))), k


def test_resize_nearest_restatement_matches_cv2():
    cv2 = pytest.importorskip("cv2")
    g = torch.Generator().manual_seed(4)
    for H, W in ((64, 64), (240, 320), (97, 131), (480, 640)):
        img = torch.rand(H, W, generator=g)
        want = cv2.resize(img.numpy(), (256, 144), interpolation=cv2.INTER_NEAREST)
        assert torch.equal(O.resize_nearest(img), torch.from_numpy(want)), (H, W)
