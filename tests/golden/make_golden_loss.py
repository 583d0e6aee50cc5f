"""Golden vectors for the ray-wise loss statistics (lidf_ray_loss), produced by the REFERENCE's own compute_loss.

    python tests/golden/make_golden_loss.py        # needs /root/reference; writes tests/golden/loss_*.npz

Runs ``LIDF.compute_loss(data_dict, 'train', epoch)`` (reference src/models/pipeline.py:468-650, unmodified; torch_scatter
and matplotlib stubbed exactly as in make_golden.py) on the outputs stored in an existing fixture plus seeded
``gt_pos`` / ``pcl_label`` / ``xyz_flat``, and stores the four scalars this repo reproduces natively -- pos_loss, prob_loss,
acc, err -- together with the scatter_log_softmax / scatter_max intermediates (computed with the same stub calls).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import make_golden as MG  # noqa: E402


def run(src_name, out_name, seed, zero_frac=0.1, label_frac=0.2):
    z = np.load(os.path.join(HERE, src_name + ".npz"))
    opt, lidf, _ = MG.build_reference({})
    import torch_scatter as ts                                   # the stub installed by build_reference
    g = torch.Generator().manual_seed(seed)
    B, H, W = int(z["meta_B"]), int(z["meta_H"]), int(z["meta_W"])
    ray = torch.from_numpy(z["miss_ray_intersect_idx"]).long()
    P, R = ray.shape[0], z["miss_ray_dir"].shape[0]
    pred_pos = torch.from_numpy(z["ref.pred_pos"])
    gt_pos = pred_pos + 0.05 * torch.randn(R, 3, generator=g)
    gt_pos[torch.rand(R, generator=g) < zero_frac] = 0.0          # zero-depth points are excluded from `err`
    label = (torch.rand(P, generator=g) < label_frac).long()
    img_ind = torch.from_numpy(z["miss_img_ind"]).long()
    dd = dict(bs=B, h=H, w=W, pred_pos=pred_pos, gt_pos=gt_pos,
              pred_prob_end=torch.from_numpy(z["ref.pred_prob_end"]),
              pred_prob_end_softmax=torch.from_numpy(z["ref.pred_prob_end_softmax"]),
              miss_ray_intersect_idx=ray, pcl_label=label, total_miss_sample_num=R,
              miss_bid=torch.from_numpy(z["miss_bid"]).long(), miss_flat_img_id=img_ind[:, 1] * W + img_ind[:, 0],
              xyz_flat=torch.randn(B, H * W, 3, generator=g))
    with torch.no_grad():
        loss = lidf.compute_loss(dd, "train", 0)                 # reference code, unmodified
        lsm = ts.scatter_log_softmax(dd["pred_prob_end"][:, 0], ray)
        _, pred_label = ts.scatter_max(dd["pred_prob_end_softmax"], ray, dim_size=R)
        _, gt_label = ts.scatter_max(label, ray, dim_size=R)
    out = os.path.join(HERE, out_name + ".npz")
    np.savez_compressed(out, pred_prob_end=z["ref.pred_prob_end"], pred_prob_end_softmax=z["ref.pred_prob_end_softmax"],
                        miss_ray_intersect_idx=z["miss_ray_intersect_idx"], pcl_label=label.numpy().astype(np.int32),
                        pred_pos=pred_pos.numpy(), gt_pos=gt_pos.numpy(), R=R, B=B, H=H, W=W,
                        xyz_flat=dd["xyz_flat"].numpy(), miss_bid=z["miss_bid"], miss_flat_img_id=dd["miss_flat_img_id"].numpy().astype(np.int32),
                        ref_pred_surf_norm_img=dd["pred_surf_norm_img"].numpy(), ref_gt_surf_norm_img=dd["gt_surf_norm_img"].numpy(),
                        ref_surf_norm_loss=np.float64(float(loss["surf_norm_loss"])), ref_smooth_loss=np.float64(float(loss["smooth_loss"])),
                        ref_angle_err=np.float64(float(loss["angle_err"])), ref_loss_net=np.float64(float(loss["loss_net"])),
                        ref_log_softmax=lsm.numpy(), ref_pred_label=pred_label.numpy(), ref_gt_label=gt_label.numpy(),
                        **{"ref_" + k: np.float64(float(loss[k])) for k in ("pos_loss", "prob_loss", "acc", "err")})
    print(out_name, {k: float(loss[k]) for k in ("pos_loss", "prob_loss", "acc", "err", "surf_norm_loss", "smooth_loss", "angle_err")}, "P", P, "R", R,
          "empty rays", int((pred_label == P).sum()))


def run_metrics(src_name, out_name, seed):
    """Depth metrics of compute_loss(..., 'test', epoch) (pipeline.py:570-618): the bs != 1 branch (rays) and, for one-image
    fixtures, the bs == 1 branch with its cv2.resize(256x144, INTER_NEAREST) host round trip -- reference code, unmodified."""
    z = np.load(os.path.join(HERE, src_name + ".npz"))
    opt, lidf, _ = MG.build_reference({})
    g = torch.Generator().manual_seed(seed)
    B, H, W = int(z["meta_B"]), int(z["meta_H"]), int(z["meta_W"])
    ray = torch.from_numpy(z["miss_ray_intersect_idx"]).long()
    P, R = ray.shape[0], z["miss_ray_dir"].shape[0]
    pred_pos = torch.from_numpy(z["ref.pred_pos"])
    gt_pos = pred_pos * (1 + 0.08 * torch.randn(R, 1, generator=g))
    gt_pos[torch.rand(R, generator=g) < 0.1] = 0.0
    label = (torch.rand(P, generator=g) < 0.2).long()
    img_ind = torch.from_numpy(z["miss_img_ind"]).long()
    depth = 0.4 + 1.5 * torch.rand(B, H * W, generator=g)
    depth[torch.rand(B, H * W, generator=g) < 0.03] = 0.0                       # holes
    depth[0, 5] = float("nan"); depth[0, 9] = float("inf")                      # the reference zeroes these (:581-582)
    xyz_flat = torch.stack((torch.randn(B, H * W, generator=g), torch.randn(B, H * W, generator=g), depth), -1)
    corrupt = (torch.rand(B, H, W, generator=g) < 0.5).float()
    xyz_corrupt_flat = xyz_flat * (1 - corrupt).reshape(B, H * W, 1)
    xyz_corrupt_flat = torch.where(torch.isfinite(xyz_corrupt_flat), xyz_corrupt_flat, torch.zeros_like(xyz_corrupt_flat))
    dd = dict(bs=B, h=H, w=W, pred_pos=pred_pos, gt_pos=gt_pos, pred_prob_end=torch.from_numpy(z["ref.pred_prob_end"]),
              pred_prob_end_softmax=torch.from_numpy(z["ref.pred_prob_end_softmax"]), miss_ray_intersect_idx=ray,
              pcl_label=label, total_miss_sample_num=R, miss_bid=torch.from_numpy(z["miss_bid"]).long(),
              miss_flat_img_id=img_ind[:, 1] * W + img_ind[:, 0], xyz_flat=xyz_flat, xyz_corrupt_flat=xyz_corrupt_flat,
              corrupt_mask=corrupt)
    with torch.no_grad():
        loss = lidf.compute_loss(dd, "test", 0)                  # reference code, unmodified (cv2 branch when bs == 1)
    keys = ("a1", "a2", "a3", "rmse", "rmse_log", "log10", "abs_rel", "mae", "sq_rel")
    out = os.path.join(HERE, out_name + ".npz")
    np.savez_compressed(out, pred_pos=pred_pos.numpy(), gt_pos=gt_pos.numpy(), xyz_flat=xyz_flat.numpy(),
                        xyz_corrupt_flat=xyz_corrupt_flat.numpy(), corrupt_mask=corrupt.numpy(),
                        miss_flat_img_id=dd["miss_flat_img_id"].numpy().astype(np.int32), B=B, H=H, W=W,
                        **{"ref_" + k: np.float64(float(loss[k])) for k in keys})
    print(out_name, "bs", B, {k: round(float(loss[k]), 6) for k in keys})


if __name__ == "__main__":
    run_metrics("c1_imnet_64x64x16", "metrics_bs1_64x64", 21)
    run_metrics("ief_ragged_2x24x32", "metrics_bs2_24x32", 22)
    if "--metrics-only" in sys.argv:
        sys.exit(0)
    run("ief_ragged_2x24x32", "loss_ief_ragged_2x24x32", 11)
    run("c1_imnet_64x64x16", "loss_c1_imnet_64x64x16", 12, zero_frac=0.0, label_frac=0.05)
