"""CPU oracle for the LIDF per-query-point decoder hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``implicit_depth_b200/`` imports this
file; only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may.  It is a plain restatement, in
stock PyTorch ops, of the reference algorithm, written function by function
against the reference source (paths relative to /root/reference):

  embed                 src/models/implicit_net.py:9-57   (Embedder / get_embedder)
  imnet_forward         src/models/implicit_net.py:81-98  (IMNet.forward)
  ief_forward           src/models/implicit_net.py:129-152 (IEF.forward)
  roi_align_aligned     torchvision.ops.roi_align (third party, torchvision 0.7.0 pinned
                        by README.md:45 / Dockerfile:29; aligned=True, sampling_ratio=-1),
                        call site src/models/pipeline.py:384-387
  scatter_softmax/max   torch_scatter (third party, un-vendored, version unpinned,
                        requirements.txt:11; 2.0.x semantics), call sites
                        src/models/pipeline.py:442,445-450
  get_embedding         src/models/pipeline.py:338-425 (minus resnet_model :370, pnet_model :407)
  get_pred              src/models/pipeline.py:427-466
  refine_decoder_tail   src/models/pipeline.py:1018-1029
  scatter_log_softmax, ray_loss_stats
                        src/models/pipeline.py:472,482-486,553-567 (compute_loss, ray-keyed terms)
  image_gradient, surface_normal, image_loss_stats
                        src/utils/point_utils.py:208-235, src/models/pipeline.py:494-541 (compute_loss, image-space terms)
  pointnet2stage_forward
                        src/models/pointnet.py:22-38
  (the geometry ops -- ray_aabb, pcl_aabb, voxelisation -- are restated in oracle/aabb_oracle.py)

Parity pinning: the reference ships no tests / golden vectors for this path
(SURVEY.md section 4), so this oracle is pinned against outputs of the reference's own
unmodified ``LIDF.get_embedding`` + ``LIDF.get_pred`` executed in the build
container (tests/golden/make_golden.py -> tests/golden/*.npz), of its ``compute_loss``
(make_golden_loss.py) and of its ``PointNet2Stage`` (make_golden_pointnet.py), and against
torchvision's ``roi_align`` (tests/test_oracle.py).

All functions are dtype-generic: pass float64 tensors to get an error budget
reference, float32 to mimic the reference bit-for-bit-ish.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------- #
# implicit_net.py
# --------------------------------------------------------------------------- #


def embed_out_dim(multires: int, enabled: bool = True) -> int:
    """Output width of get_embedder (implicit_net.py:42-57): 3 + 3*2*multires, or 3."""
    return 3 + 6 * multires if enabled else 3


def embed(x: torch.Tensor, multires: int, enabled: bool = True) -> torch.Tensor:
    """Embedder.embed (implicit_net.py:14-39).

    Layout [x, sin(f0 x), cos(f0 x), sin(f1 x), cos(f1 x), ...], f_k = 2**k for
    k = 0..multires-1 (``2.**linspace(0, multires-1, multires)``, log_sampling).
    ``enabled=False`` is get_embedder(i=-1) -> nn.Identity.
    """
    if not enabled:
        return x
    outs = [x]
    freq_bands = 2.0 ** torch.linspace(0.0, multires - 1, steps=multires)
    for freq in freq_bands:
        f = float(freq)
        outs.append(torch.sin(x * f))
        outs.append(torch.cos(x * f))
    return torch.cat(outs, -1)


def _final_act(l4: torch.Tensor, use_sigmoid: bool) -> torch.Tensor:
    """implicit_net.py:93-96 / :147-150."""
    if use_sigmoid:
        return torch.sigmoid(l4)
    return torch.max(torch.min(l4, l4 * 0.01 + 0.99), l4 * 0.01)


def imnet_forward(p: Dict[str, torch.Tensor], x: torch.Tensor, use_sigmoid: bool = False) -> torch.Tensor:
    """IMNet.forward (implicit_net.py:81-98). ``p`` holds state_dict keys linear_{1..4}.{weight,bias}."""
    l1 = F.leaky_relu(F.linear(x, p["linear_1.weight"], p["linear_1.bias"]), 0.02)
    l2 = F.leaky_relu(F.linear(l1, p["linear_2.weight"], p["linear_2.bias"]), 0.02)
    l3 = F.leaky_relu(F.linear(l2, p["linear_3.weight"], p["linear_3.bias"]), 0.02)
    l4 = F.linear(l3, p["linear_4.weight"], p["linear_4.bias"])
    return _final_act(l4, use_sigmoid)


def ief_forward(p: Dict[str, torch.Tensor], x: torch.Tensor, n_iter: int, use_sigmoid: bool = False,
                init_offset: float = 0.001) -> torch.Tensor:
    """IEF.forward (implicit_net.py:129-152). Extra keys offset_enc.{weight,bias}."""
    pred = torch.full((x.shape[0], 1), init_offset, dtype=x.dtype, device=x.device)
    for _ in range(n_iter):
        off_feat = F.linear(pred, p["offset_enc.weight"], p["offset_enc.bias"])
        xc = torch.cat([x, off_feat], 1)
        l1 = F.leaky_relu(F.linear(xc, p["linear_1.weight"], p["linear_1.bias"]), 0.02)
        l2 = F.leaky_relu(F.linear(l1, p["linear_2.weight"], p["linear_2.bias"]), 0.02)
        l3 = F.leaky_relu(F.linear(l2, p["linear_3.weight"], p["linear_3.bias"]), 0.02)
        l4 = F.linear(l3, p["linear_4.weight"], p["linear_4.bias"])
        pred = pred + l4
    return _final_act(pred, use_sigmoid)


def decoder_forward(kind: str, p: Dict[str, torch.Tensor], x: torch.Tensor, n_iter: int,
                    use_sigmoid: bool) -> torch.Tensor:
    if kind.upper() == "IMNET":
        return imnet_forward(p, x, use_sigmoid)
    if kind.upper() == "IEF":
        return ief_forward(p, x, n_iter, use_sigmoid)
    raise NotImplementedError(kind)


# --------------------------------------------------------------------------- #
# torchvision.ops.roi_align (aligned=True, sampling_ratio=-1)
# --------------------------------------------------------------------------- #


def _bilinear(feat: torch.Tensor, bid: torch.Tensor, y: torch.Tensor, x: torch.Tensor) -> torch.Tensor:
    """torchvision bilinear_interpolate for a vector of sample points.

    feat [B,C,H,W]; bid,y,x [R] -> [R,C].  Out-of-range (< -1 or > size) samples give 0.
    """
    B, C, H, W = feat.shape
    dead = (y < -1.0) | (y > H) | (x < -1.0) | (x > W)
    y = torch.clamp(y, min=0.0)
    x = torch.clamp(x, min=0.0)
    y_low = y.to(torch.int64)  # trunc == floor for y >= 0
    x_low = x.to(torch.int64)
    ycap = y_low >= H - 1
    xcap = x_low >= W - 1
    y_low = torch.where(ycap, torch.full_like(y_low, H - 1), y_low)
    x_low = torch.where(xcap, torch.full_like(x_low, W - 1), x_low)
    y_high = torch.where(ycap, y_low, y_low + 1)
    x_high = torch.where(xcap, x_low, x_low + 1)
    y = torch.where(ycap, y_low.to(y.dtype), y)
    x = torch.where(xcap, x_low.to(x.dtype), x)
    # dead samples may hold garbage indices: make them safe
    y_low = y_low.clamp(0, H - 1); y_high = y_high.clamp(0, H - 1)
    x_low = x_low.clamp(0, W - 1); x_high = x_high.clamp(0, W - 1)
    ly = y - y_low.to(y.dtype)
    lx = x - x_low.to(x.dtype)
    hy = 1.0 - ly
    hx = 1.0 - lx
    fl = feat.permute(0, 2, 3, 1)  # [B,H,W,C] view
    v1 = fl[bid, y_low, x_low]
    v2 = fl[bid, y_low, x_high]
    v3 = fl[bid, y_high, x_low]
    v4 = fl[bid, y_high, x_high]
    w1 = (hy * hx).unsqueeze(-1)
    w2 = (hy * lx).unsqueeze(-1)
    w3 = (ly * hx).unsqueeze(-1)
    w4 = (ly * lx).unsqueeze(-1)
    val = w1 * v1 + w2 * v2 + w3 * v3 + w4 * v4
    return torch.where(dead.unsqueeze(-1), torch.zeros_like(val), val)


def roi_align_aligned(feat: torch.Tensor, boxes: torch.Tensor, out_size: int = 2,
                      spatial_scale: float = 1.0) -> torch.Tensor:
    """roi_align(feat, boxes[K,5]=(bid,x1,y1,x2,y2), out_size, spatial_scale, sampling_ratio=-1, aligned=True).

    Returns [K, C, out, out].  Restates torchvision's roi_align kernel: adaptive sampling grid
    ceil(roi/out) per axis, plain sum in (iy, ix) order, divided by max(gh*gw, 1).
    """
    K = boxes.shape[0]
    B, C, H, W = feat.shape
    dt = feat.dtype
    bid = boxes[:, 0].to(torch.int64)
    x1 = boxes[:, 1].to(dt) * spatial_scale - 0.5
    y1 = boxes[:, 2].to(dt) * spatial_scale - 0.5
    x2 = boxes[:, 3].to(dt) * spatial_scale - 0.5
    y2 = boxes[:, 4].to(dt) * spatial_scale - 0.5
    roi_w = x2 - x1
    roi_h = y2 - y1
    bin_w = roi_w / out_size
    bin_h = roi_h / out_size
    gw = torch.ceil(roi_w / out_size).to(torch.int64)
    gh = torch.ceil(roi_h / out_size).to(torch.int64)
    count = torch.clamp(gh * gw, min=1).to(dt)
    gmax_h = int(gh.max().item()) if K else 0
    gmax_w = int(gw.max().item()) if K else 0
    out = torch.zeros(K, C, out_size, out_size, dtype=dt, device=feat.device)
    gwf = torch.clamp(gw, min=1).to(dt)
    ghf = torch.clamp(gh, min=1).to(dt)
    for ph in range(out_size):
        for pw in range(out_size):
            acc = torch.zeros(K, C, dtype=dt, device=feat.device)
            for iy in range(gmax_h):
                y = y1 + ph * bin_h + (iy + 0.5) * bin_h / ghf
                for ix in range(gmax_w):
                    x = x1 + pw * bin_w + (ix + 0.5) * bin_w / gwf
                    live = (iy < gh) & (ix < gw)
                    val = _bilinear(feat, bid, y, x)
                    acc = acc + torch.where(live.unsqueeze(-1), val, torch.zeros_like(val))
            out[:, :, ph, pw] = acc / count.unsqueeze(-1)
    return out


# --------------------------------------------------------------------------- #
# torch_scatter (2.0.x semantics)
# --------------------------------------------------------------------------- #


def scatter_max(src: torch.Tensor, index: torch.Tensor, dim_size: Optional[int] = None
                ) -> Tuple[torch.Tensor, torch.Tensor]:
    """torch_scatter.scatter_max on a 1-D src: (max, argmax); empty segment -> (0, src.numel());
    first maximum wins on ties (CPU kernel iterates in order with a strict compare)."""
    P = src.shape[0]
    if dim_size is None:
        dim_size = int(index.max().item()) + 1 if P else 0
    neg = torch.full((dim_size,), -float("inf"), dtype=src.dtype, device=src.device)
    mx = neg.scatter_reduce(0, index, src, reduce="amax", include_self=True)
    pos = torch.arange(P, device=src.device, dtype=torch.int64)
    cand = torch.where(src == mx[index], pos, torch.full_like(pos, P))
    arg = torch.full((dim_size,), P, dtype=torch.int64, device=src.device)
    arg = arg.scatter_reduce(0, index, cand, reduce="amin", include_self=True)
    empty = arg == P
    mx = torch.where(empty, torch.zeros_like(mx), mx)
    return mx, arg


def scatter_softmax(src: torch.Tensor, index: torch.Tensor, eps: float = 1e-12) -> torch.Tensor:
    """torch_scatter.composite.scatter_softmax: exp(src - segmax) / (segsum + eps)."""
    if src.shape[0] == 0:
        return src.clone()
    mx, _ = scatter_max(src, index)
    rec = (src - mx[index]).exp()
    ssum = torch.zeros_like(mx).index_add_(0, index, rec)
    return rec / (ssum + eps)[index]


def scatter_log_softmax(src: torch.Tensor, index: torch.Tensor, eps: float = 1e-12) -> torch.Tensor:
    """torch_scatter.composite.scatter_log_softmax: (src - segmax) - log(segsum(exp(src - segmax)) + eps)."""
    if src.shape[0] == 0:
        return src.clone()
    mx, _ = scatter_max(src, index)
    rec = src - mx[index]
    ssum = torch.zeros_like(mx).index_add_(0, index, rec.exp())
    return rec - (ssum + eps).log()[index]


def ray_loss_stats(pred_prob_end: torch.Tensor, pred_prob_end_softmax: torch.Tensor, ray: torch.Tensor,
                   pcl_label: torch.Tensor, R: int, pred_pos: Optional[torch.Tensor] = None,
                   gt_pos: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
    """The ray-keyed part of LIDF.compute_loss (hard_neg False, the shipped setting):
      pipeline.py:472      pos_loss = L1Loss()(pred_pos, gt_pos)
      pipeline.py:482-486  prob_loss = mean(-scatter_log_softmax(pred_prob_end[:,0], ray)[nonzero(pcl_label)])
      pipeline.py:553-557  acc = mean(scatter_max(softmax, ray)[1] == scatter_max(pcl_label, ray)[1])
      pipeline.py:560-567  err = sum(||pred_pos - gt_pos||_2 * zero_mask) / sum(zero_mask), zero_mask = (sum|gt_pos| != 0)"""
    logit = pred_prob_end.reshape(-1)
    lsm = scatter_log_softmax(logit, ray)
    idx = torch.nonzero(pcl_label, as_tuple=False).reshape(-1)
    out = dict(log_softmax=lsm, prob_loss=torch.mean(-1 * lsm[idx]))
    _, out["pred_label"] = scatter_max(pred_prob_end_softmax, ray, dim_size=R)
    _, out["gt_label"] = scatter_max(pcl_label.to(pred_prob_end_softmax.dtype), ray, dim_size=R)
    out["acc"] = torch.sum(torch.eq(out["pred_label"], out["gt_label"]).float()) / max(R, 1)
    if gt_pos is not None:
        out["pos_loss"] = F.l1_loss(pred_pos, gt_pos)
        zero_mask = torch.sum(gt_pos.abs(), dim=-1)
        zero_mask[zero_mask != 0] = 1.
        n = torch.sum(zero_mask)
        out["err"] = torch.zeros(()) if float(n) == 0 else \
            torch.sum(torch.sqrt(torch.sum((pred_pos - gt_pos) ** 2, -1)) * zero_mask) / n
    return out


def image_gradient(x: torch.Tensor):
    """point_utils.gradient (/root/reference/src/utils/point_utils.py:208-227): forward differences, zero in the last
    column (dx) / last row (dy).  x: [b,c,h,w]."""
    right = F.pad(x, [0, 1, 0, 0])[:, :, :, 1:]
    bottom = F.pad(x, [0, 0, 0, 1])[:, :, 1:, :]
    dx, dy = right - x, bottom - x
    dx[:, :, :, -1] = 0
    dy[:, :, -1, :] = 0
    return dx, dy


def surface_normal(x: torch.Tensor):
    """point_utils.get_surface_normal (point_utils.py:229-235)."""
    dx, dy = image_gradient(x)
    n = torch.cross(dx, dy, dim=1)
    n = n / (torch.norm(n, dim=1, keepdim=True) + 1e-8)
    return n, dx, dy


def image_loss_stats(xyz_flat: torch.Tensor, miss_bid: torch.Tensor, miss_flat_img_id: torch.Tensor, pred_pos: torch.Tensor,
                     gt_pos: torch.Tensor, bs: int, h: int, w: int) -> Dict[str, torch.Tensor]:
    """The image-space part of LIDF.compute_loss (hard_neg False): pipeline.py:494-541 -- surface-normal loss, angle error
    and smoothness loss at the miss pixels of the point image with pred_pos / gt_pos scattered in.  ``xyz_flat`` [bs,h*w,3]
    is xyz_flat (train) or xyz_corrupt_flat (otherwise), :495-500."""
    gt_pcl = xyz_flat.clone(); pred_pcl = xyz_flat.clone()
    gt_pcl[miss_bid, miss_flat_img_id] = gt_pos                                           # :501
    gt_n, _, _ = surface_normal(gt_pcl.reshape(bs, h, w, 3).permute(0, 3, 1, 2).contiguous())
    gt_sn = gt_n.permute(0, 2, 3, 1).contiguous().reshape(bs, h * w, 3)[miss_bid, miss_flat_img_id]
    pred_pcl[miss_bid, miss_flat_img_id] = pred_pos                                       # :507
    pred_n, dx, dy = surface_normal(pred_pcl.reshape(bs, h, w, 3).permute(0, 3, 1, 2).contiguous())
    pred_sn = pred_n.permute(0, 2, 3, 1).contiguous().reshape(bs, h * w, 3)[miss_bid, miss_flat_img_id]
    cosine_val = F.cosine_similarity(pred_sn, gt_sn, dim=-1)                              # :514
    surf_norm_loss = torch.mean((1 - cosine_val) / 2.)                                    # :515-517
    angle_err = torch.mean(torch.acos(torch.clamp(cosine_val, min=-1, max=1))) / math.pi * 180.   # :523-524
    dxd = torch.sum(dx * dx, 1).reshape(bs, h * w)[miss_bid, miss_flat_img_id]            # :527-529
    dyd = torch.sum(dy * dy, 1).reshape(bs, h * w)[miss_bid, miss_flat_img_id]            # :531-533
    smooth_loss = torch.mean(dxd) + torch.mean(dyd)                                       # :536
    return dict(surf_norm_loss=surf_norm_loss, angle_err=angle_err, smooth_loss=smooth_loss, pred_surf_norm_img=pred_n,
                gt_surf_norm_img=gt_n, cosine_val=cosine_val)


# --------------------------------------------------------------------------- #
# pipeline.py : LIDF.get_embedding + LIDF.get_pred
# --------------------------------------------------------------------------- #

DEFAULT_CFG = dict(
    pos_encode=True, multires=8, multires_views=4, intersect_pos_type="abs",
    roi_inp_bbox=8, roi_out_bbox=2, offdec_type="IEF", n_iter=2, probdec_type="IMNET",
    use_sigmoid=False, offset_range=(0.0, 1.0), scatter_type="Maxpool",
)


def depth_metrics(pred: torch.Tensor, gt: torch.Tensor) -> Dict[str, torch.Tensor]:
    """The nine depth metrics of LIDF.compute_loss (pipeline.py:605-618) over matched 1-D depth tensors."""
    safe_log = lambda x: torch.log(torch.clamp(x, 1e-6, 1e6))      # noqa: E731  (safe_log10 is the same natural log, :607)
    thresh = torch.max(gt / pred, pred / gt)
    return dict(a1=(thresh < 1.05).float().mean(), a2=(thresh < 1.10).float().mean(), a3=(thresh < 1.25).float().mean(),
                rmse=((gt - pred) ** 2).mean().sqrt(), rmse_log=((safe_log(gt) - safe_log(pred)) ** 2).mean().sqrt(),
                log10=(safe_log(gt) - safe_log(pred)).abs().mean(), abs_rel=((gt - pred).abs() / gt).mean(),
                mae=(gt - pred).abs().mean(), sq_rel=((gt - pred) ** 2 / gt).mean())


def resize_nearest(img: torch.Tensor, out_w: int = 256, out_h: int = 144) -> torch.Tensor:
    """cv2.resize(img, (out_w, out_h), interpolation=cv2.INTER_NEAREST) for a 2-D array (OpenCV resizeNN):
    dst(y, x) = src(min(floor(y * H / out_h), H - 1), min(floor(x * W / out_w), W - 1)), double arithmetic."""
    H, W = img.shape
    ys = torch.clamp(torch.floor(torch.arange(out_h, dtype=torch.float64) * (H / out_h)).long(), max=H - 1)
    xs = torch.clamp(torch.floor(torch.arange(out_w, dtype=torch.float64) * (W / out_w)).long(), max=W - 1)
    return img[ys][:, xs]


def depth_metrics_rays(pred_pos: torch.Tensor, gt_pos: torch.Tensor) -> Dict[str, torch.Tensor]:
    """bs != 1 branch (pipeline.py:560-575): rays whose gt_pos is not all-zero, z components."""
    keep = gt_pos.abs().sum(-1) != 0
    return depth_metrics(pred_pos[:, 2][keep], gt_pos[:, 2][keep])


def depth_metrics_image(xyz_flat: torch.Tensor, xyz_corrupt_flat: torch.Tensor, corrupt_mask: torch.Tensor,
                        miss_flat_img_id: torch.Tensor, pred_pos: torch.Tensor, h: int, w: int) -> Dict[str, torch.Tensor]:
    """bs == 1 branch (pipeline.py:576-603) without the host round trip."""
    gt = resize_nearest(xyz_flat[0, :, 2].reshape(h, w))
    gt = torch.where(torch.isnan(gt) | torch.isinf(gt), torch.zeros_like(gt), gt)
    seg = resize_nearest(corrupt_mask.reshape(h, w).to(torch.uint8))
    pz = xyz_corrupt_flat[0, :, 2].clone()
    pz[miss_flat_img_id] = pred_pos[:, 2]
    pred = resize_nearest(pz.reshape(h, w))
    m = (gt > 0) & (seg != 0)
    return depth_metrics(pred[m], gt[m])


def get_embedding(d: Dict[str, torch.Tensor], cfg: dict, *, roi_fn=None, dedup_rays: bool = False
                  ) -> Dict[str, torch.Tensor]:
    """LIDF.get_embedding (pipeline.py:338-425) with resnet_model / pnet_model outputs given.

    Inputs (keys follow the reference's data_dict): full_rgb_feat [B,32,H,W], occ_voxel_feat [V,128],
    occ_vox_intersect_idx [P], miss_ray_intersect_idx [P], intersect_dist [P,2] (== dist[vox,ray],
    pipeline.py:345), miss_ray_dir [R,3], miss_img_ind [R,2] (x,y) i64, miss_bid [R] i64,
    voxel_bound [V,6].
    ``roi_fn`` lets the timed reference arm plug torchvision's roi_align in; default is the restatement.
    ``dedup_rays`` evaluates ROIAlign once per ray (same values; the reference does it per pair).
    """
    feat = d["full_rgb_feat"]
    B, C, H, W = feat.shape
    vox = d["occ_vox_intersect_idx"]
    ray = d["miss_ray_intersect_idx"]
    dist = d["intersect_dist"]
    t_enter, t_leave = dist[:, 0], dist[:, 1]
    dirs = d["miss_ray_dir"][ray]                                   # :348
    enter_pos = dirs * t_enter.unsqueeze(-1)                        # :349
    leave_pos = dirs * t_leave.unsqueeze(-1)                        # :350
    vb = d["voxel_bound"][vox]                                      # :353
    center = (vb[:, :3] + vb[:, 3:]) / 2.0                          # :354
    if cfg["intersect_pos_type"] == "rel":                          # :355-360
        inp_enter, inp_leave = enter_pos - center, leave_pos - center
    else:
        inp_enter, inp_leave = enter_pos, leave_pos
    pe = cfg["pos_encode"]
    enter_embed = embed(inp_enter, cfg["multires"], pe)             # :363
    leave_embed = embed(inp_leave, cfg["multires"], pe)             # :364
    dir_embed = embed(dirs, cfg["multires_views"], pe)              # :365

    def _boxes(img_ind, bid):
        half = cfg["roi_inp_bbox"] // 2
        ul = img_ind - half                                         # :374
        br = img_ind + half                                         # :375
        ul = torch.stack((ul[:, 0].clamp(0, W - 1), ul[:, 1].clamp(0, H - 1)), -1)   # :377-378
        br = torch.stack((br[:, 0].clamp(0, W - 1), br[:, 1].clamp(0, H - 1)), -1)   # :379-380
        return torch.cat((bid.unsqueeze(-1), ul, br), -1).float()   # :381 (.float(): fp32 boxes)

    roi = roi_fn if roi_fn is not None else roi_align_aligned
    if dedup_rays:
        per_ray = roi(feat, _boxes(d["miss_img_ind"], d["miss_bid"]).to(feat.dtype), cfg["roi_out_bbox"], 1.0)
        rgb_feat = per_ray.reshape(per_ray.shape[0], -1)[ray]
    else:
        boxes = _boxes(d["miss_img_ind"][ray], d["miss_bid"][ray]).to(feat.dtype)   # :368-369
        rgb_feat = roi(feat, boxes, cfg["roi_out_bbox"], 1.0)       # :384-387
        rgb_feat = rgb_feat.reshape(rgb_feat.shape[0], -1)          # :389 -> (c, ph, pw) order
    vox_feat = d["occ_voxel_feat"][vox]                             # :410
    return dict(intersect_dir=dirs, intersect_enter_pos=enter_pos, intersect_leave_pos=leave_pos,
                intersect_enter_pos_embed=enter_embed, intersect_leave_pos_embed=leave_embed,
                intersect_dir_embed=dir_embed, intersect_rgb_feat=rgb_feat, intersect_voxel_feat=vox_feat)


def get_pred(d: Dict[str, torch.Tensor], e: Dict[str, torch.Tensor], cfg: dict,
             offset_dec: Dict[str, torch.Tensor], prob_dec: Dict[str, torch.Tensor], part_size: float,
             total_miss_sample_num: int, pcl_label_float: Optional[torch.Tensor] = None
             ) -> Dict[str, torch.Tensor]:
    """LIDF.get_pred (pipeline.py:427-466).  ``pcl_label_float`` given <=> the
    ``exp_type=='train' and epoch < maxpool_label_epo`` branch (:444-446)."""
    ray = d["miss_ray_intersect_idx"]
    inp_embed = torch.cat((e["intersect_voxel_feat"], e["intersect_rgb_feat"], e["intersect_enter_pos_embed"],
                           e["intersect_leave_pos_embed"], e["intersect_dir_embed"]), -1)   # :431-433
    pred_offset = decoder_forward(cfg["offdec_type"], offset_dec, inp_embed, cfg["n_iter"], cfg["use_sigmoid"])
    pred_prob_end = decoder_forward(cfg["probdec_type"], prob_dec, inp_embed, cfg["n_iter"], cfg["use_sigmoid"])
    r0, r1 = cfg["offset_range"]
    scaled = pred_offset * (r1 - r0) + r0                           # :437
    scaled = scaled * math.sqrt(3) * part_size                      # :438 (np.sqrt(3) python scalar)
    pair_pred_pos = e["intersect_enter_pos"] + scaled * e["intersect_dir"]   # :439
    soft = scatter_softmax(pred_prob_end.detach()[:, 0], ray)       # :442
    if pcl_label_float is not None:
        _, max_pair_id = scatter_max(pcl_label_float, ray, dim_size=total_miss_sample_num)   # :445
    else:
        _, max_pair_id = scatter_max(soft, ray, dim_size=total_miss_sample_num)              # :449
    dummy = torch.zeros(1, 3, dtype=pair_pred_pos.dtype, device=pair_pred_pos.device)        # :452
    pred_pos = torch.cat((pair_pred_pos, dummy), 0)[max_pair_id]    # :453-454
    return dict(pred_offset=pred_offset, pred_prob_end=pred_prob_end, pair_pred_pos=pair_pred_pos,
                pred_prob_end_softmax=soft, max_pair_id=max_pair_id, pred_pos=pred_pos)


def lidf_query(d: Dict[str, torch.Tensor], cfg: dict, offset_dec: Dict[str, torch.Tensor],
               prob_dec: Dict[str, torch.Tensor], part_size: float, *, roi_fn=None,
               dedup_rays: bool = False, pcl_label_float: Optional[torch.Tensor] = None
               ) -> Dict[str, torch.Tensor]:
    """get_embedding + get_pred: the whole hot path."""
    e = get_embedding(d, cfg, roi_fn=roi_fn, dedup_rays=dedup_rays)
    R = d["miss_ray_dir"].shape[0]
    out = get_pred(d, e, cfg, offset_dec, prob_dec, part_size, R, pcl_label_float)
    out["intersect_rgb_feat"] = e["intersect_rgb_feat"]
    return out


def lidf_query_chunked(d, cfg, offset_dec, prob_dec, part_size, chunk_pairs: int = 1 << 20, **kw):
    """Same result as lidf_query, evaluated over pair chunks so the P x 385 concat stays small.
    Decoder part is per pair; ray termination runs once on the concatenated logits."""
    P = d["occ_vox_intersect_idx"].shape[0]
    R = d["miss_ray_dir"].shape[0]
    offs, probs, pos = [], [], []
    for s in range(0, P, chunk_pairs):
        sl = slice(s, min(P, s + chunk_pairs))
        dd = dict(d)
        for k in ("occ_vox_intersect_idx", "miss_ray_intersect_idx", "intersect_dist"):
            dd[k] = d[k][sl]
        e = get_embedding(dd, cfg, **kw)
        x = torch.cat((e["intersect_voxel_feat"], e["intersect_rgb_feat"], e["intersect_enter_pos_embed"],
                       e["intersect_leave_pos_embed"], e["intersect_dir_embed"]), -1)
        po = decoder_forward(cfg["offdec_type"], offset_dec, x, cfg["n_iter"], cfg["use_sigmoid"])
        pp = decoder_forward(cfg["probdec_type"], prob_dec, x, cfg["n_iter"], cfg["use_sigmoid"])
        r0, r1 = cfg["offset_range"]
        sc = (po * (r1 - r0) + r0) * math.sqrt(3) * part_size
        offs.append(po); probs.append(pp); pos.append(e["intersect_enter_pos"] + sc * e["intersect_dir"])
    pred_offset = torch.cat(offs); pred_prob_end = torch.cat(probs); pair_pred_pos = torch.cat(pos)
    ray = d["miss_ray_intersect_idx"]
    soft = scatter_softmax(pred_prob_end[:, 0], ray)
    _, max_pair_id = scatter_max(soft, ray, dim_size=R)
    dummy = torch.zeros(1, 3, dtype=pair_pred_pos.dtype, device=pair_pred_pos.device)
    pred_pos = torch.cat((pair_pred_pos, dummy), 0)[max_pair_id]
    return dict(pred_offset=pred_offset, pred_prob_end=pred_prob_end, pair_pred_pos=pair_pred_pos,
                pred_prob_end_softmax=soft, max_pair_id=max_pair_id, pred_pos=pred_pos)


# --------------------------------------------------------------------------- #
# pipeline.py : RefineNet.get_pred_refine decoder tail (:1018-1029)
# --------------------------------------------------------------------------- #

REFINE_CFG = dict(pos_encode=True, multires=8, multires_views=4, intersect_pos_type="abs",
                  offdec_type="IEF", n_iter=2, use_sigmoid=False, offset_range=(-0.2, 0.2))


def refine_decoder_tail(pred_pos: torch.Tensor, miss_ray_dir: torch.Tensor, end_voxel_center: torch.Tensor,
                        voxel_feat_end: torch.Tensor, rgb_feat_end: torch.Tensor, cfg: dict,
                        offset_dec: Dict[str, torch.Tensor]) -> torch.Tensor:
    """pipeline.py:1018-1029: PE(pos[-centre]) | PE(dir) -> concat(voxel, rgb, pos, dir) -> offset_dec ->
    pred_pos + (o*(r1-r0)+r0)*dir.  (No sqrt(3)*part_size factor here, unlike get_pred.)"""
    if cfg["intersect_pos_type"] == "rel":
        enter_pos = pred_pos - end_voxel_center                     # :1020
    else:
        enter_pos = pred_pos                                        # :1022
    pos_embed = embed(enter_pos, cfg["multires"], cfg["pos_encode"])            # :1023
    dir_embed = embed(miss_ray_dir, cfg["multires_views"], cfg["pos_encode"])   # :947
    x = torch.cat((voxel_feat_end, rgb_feat_end, pos_embed, dir_embed), -1)     # :1025-1026
    off = decoder_forward(cfg["offdec_type"], offset_dec, x, cfg["n_iter"], cfg["use_sigmoid"])   # :1027
    r0, r1 = cfg["offset_range"]
    scaled = off * (r1 - r0) + r0                                   # :1028
    return pred_pos + scaled * miss_ray_dir                         # :1029


# --------------------------------------------------------------------------- #
# pointnet.py : PointNet2Stage (the producer of occ_voxel_feat)
# --------------------------------------------------------------------------- #
def _scatter_rows_max(src: torch.Tensor, index: torch.Tensor, n: int) -> torch.Tensor:
    """torch_scatter.scatter(src, index, dim=0, reduce='max'): per-row segment max; rows nobody writes stay 0."""
    out = torch.full((n, src.shape[1]), -float("inf"), dtype=src.dtype)
    out = out.scatter_reduce(0, index.reshape(-1, 1).expand_as(src), src, reduce="amax", include_self=True)
    return torch.where(torch.isinf(out), torch.zeros_like(out), out)


def pointnet2stage_forward(p: Dict[str, torch.Tensor], inp_feat: torch.Tensor, vox2point_idx: torch.Tensor,
                           n_vox: Optional[int] = None) -> torch.Tensor:
    """/root/reference/src/models/pointnet.py:22-38, state_dict keys point_lin{1..4}, vox_lin{1,2}."""
    n = int(vox2point_idx.max()) + 1 if n_vox is None else n_vox
    lin = lambda name, x: F.linear(x, p[name + ".weight"], p[name + ".bias"])
    f1 = F.relu(lin("point_lin1", inp_feat))                                   # :24
    f2 = F.relu(lin("point_lin2", f1))                                         # :25
    v1 = F.relu(lin("vox_lin1", _scatter_rows_max(f2, vox2point_idx, n)))      # :27-28
    f3 = torch.cat((v1[vox2point_idx], f2), -1)                                # :30-31
    f5 = F.relu(lin("point_lin4", F.relu(lin("point_lin3", f3))))              # :32-33
    return F.relu(lin("vox_lin2", _scatter_rows_max(f5, vox2point_idx, n)))    # :35-36


# --------------------------------------------------------------------------- #
# helpers shared by tests / bench
# --------------------------------------------------------------------------- #


def init_decoder(kind: str, inp_dim: int, gf: int = 64, *, mode: str = "reference",
                 generator: Optional[torch.Generator] = None, dtype=torch.float32) -> Dict[str, torch.Tensor]:
    """Random decoder parameters with the reference state_dict keys.

    mode 'reference': implicit_net.py:72-79 / :117-126 (N(0,0.02), linear_4.weight N(1e-5,0.02), zero bias).
    mode 'trained'  : N(0, 1/sqrt(fan_in)) weights and small random biases so activations are O(1)
                      and the softmax / argmax are exercised (SURVEY.md section 7 hard part 1).
    """
    g = generator
    is_ief = kind.upper() == "IEF"
    dims = [(inp_dim + (16 if is_ief else 0), gf * 4), (gf * 4, gf * 2), (gf * 2, gf), (gf, 1)]
    p: Dict[str, torch.Tensor] = {}

    def rn(*shape):
        return torch.randn(*shape, generator=g, dtype=torch.float32)

    for i, (fi, fo) in enumerate(dims, 1):
        if mode == "reference":
            w = rn(fo, fi) * 0.02 + (1e-5 if i == 4 else 0.0)
            b = torch.zeros(fo)
        else:
            w = rn(fo, fi) / math.sqrt(fi)
            b = rn(fo) * 0.1
        p[f"linear_{i}.weight"] = w.to(dtype)
        p[f"linear_{i}.bias"] = b.to(dtype)
    if is_ief:
        if mode == "reference":
            p["offset_enc.weight"] = (rn(16, 1) * 0.02).to(dtype)
            p["offset_enc.bias"] = torch.zeros(16, dtype=dtype)
        else:
            p["offset_enc.weight"] = rn(16, 1).to(dtype)
            p["offset_enc.bias"] = (rn(16) * 0.1).to(dtype)
    return p


def cast_tree(t, dtype):
    if isinstance(t, dict):
        return {k: cast_tree(v, dtype) for k, v in t.items()}
    if isinstance(t, torch.Tensor) and t.is_floating_point():
        return t.to(dtype)
    return t
