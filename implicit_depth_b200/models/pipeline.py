"""Host-side mirror of the hot-path part of the reference's ``models.pipeline`` (reference src/models/pipeline.py).

Only the per-query-point decoder path is re-implemented: ``LIDF.get_embedding`` (pipeline.py:338-425),
``LIDF.get_pred`` (:427-466) and the decoder tail of ``RefineNet.get_pred_refine`` (:1018-1029).  Everything
else in the reference pipeline (data prep, voxelisation, ray generation, ray_aabb, GT, losses, trainers) is out of
scope and is meant to keep running from the reference tree; INTEGRATION.md shows the two-method override.

``LIDFQueryMixin`` carries the two replacement methods with the reference's exact names, arguments and ``data_dict``
contract, so it can be mixed into the reference class::

    class LIDF(LIDFQueryMixin, reference_pipeline.LIDF): pass

``LIDF`` / ``RefineNet`` below are stand-alone holders of the same sub-module attributes (``embed_fn``,
``embeddirs_fn``, ``offset_dec``, ``prob_dec``, ``resnet_model``, ``pnet_model``) for use without the reference tree
(tests, bench); the two producers are injected, they are not part of this path.

Both inference and training run the fused sm_100a kernels through the C ABI -- there is no CPU or eager fallback.
When autograd is recording, ``get_pred`` is ONE autograd node (``_LidfQueryFn``): forward = ``lidf_query_forward``,
backward = ``lidf_query_backward`` (tcgen05 dgrad / wgrad kernels, csrc/lidf_bwd.cuh); the decoder parameters receive
ordinary ``.grad`` tensors, so the reference's DistributedDataParallel wrapper (train_lidf.py:120) works unchanged.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn as nn

from implicit_depth_b200.extensions.lidf_query.jit import lidf_query
from implicit_depth_b200.models import implicit_net as im_net
from implicit_depth_b200.models import pointnet as pnet


XMIN = (-1.0, -1.0, 0.0)     # reference src/constants.py:15
XMAX = (1.0, 1.0, 2.0)       # reference src/constants.py:16


def _multires_of(fn, default):
    return getattr(fn, "multires", default)


class LIDFQueryMixin:
    """Replacement ``get_embedding`` / ``get_pred`` (reference pipeline.py:338-466).

    Reads from ``self``: ``opt.model.{pos_encode, multires, multires_views, intersect_pos_type, roi_inp_bbox,
    roi_out_bbox, n_iter, use_sigmoid, maxpool_label_epo, scatter_type, pnet_pos_type}``, ``opt.grid.offset_range``,
    ``resnet_model``, ``pnet_model``, ``offset_dec``, ``prob_dec`` -- the same attributes the reference uses.
    """

    mlp_impl = "auto"   # "auto" | "tc_bf16x3" | "simt_fp32" | "tc_bf16x1" (see include/lidf_query.h)
    # Order of the pair list compute_ray_aabb writes: "nonzero" = the reference's torch.nonzero order (voxel, then ray;
    # bit-identical index tensors), "ray" = sorted by ray, then voxel -- the same pairs, emitted ray-major by the pair
    # generator so that get_pred skips its regroup (every reader of the pair tensors in the reference is keyed by
    # miss_ray_intersect_idx / occ_vox_intersect_idx and does not depend on their order).
    pair_order = "nonzero"
    # True = get_pred computes the per-ray results only (pred_pos, max_pair_id, pred_prob_end,
    # pred_prob_end_softmax -- everything compute_loss and RefineNet read); the offset decoder then runs on each ray's arg-max
    # pair instead of on all pairs.  data_dict['pair_pred_pos'] / ['pred_offset'] (written by the reference, never read) are
    # not produced.  Same bits in every produced tensor; in training the offset decoder's backward runs over those rows only
    # (every other pair's upstream gradient is exactly zero in the reference's loss).
    winner_only = False

    def get_embedding(self, data_dict):
        """Runs the two upstream producers exactly where the reference does (pipeline.py:370, :400-407) and stores
        their outputs; the per-pair gathers / PE / ROIAlign of the reference happen inside the fused kernel."""
        data_dict['full_rgb_feat'] = self.resnet_model(data_dict['rgb_img'])
        valid_v_rgb = data_dict['valid_rgb'][data_dict['valid_v_pid']]
        if self.opt.model.pnet_pos_type == 'rel':
            pnet_inp = torch.cat((data_dict['valid_v_rel_coord'], valid_v_rgb), -1)
        else:
            raise NotImplementedError('Does not support Pnet pos type: {}'.format(self.opt.model.pnet_pos_type))
        data_dict['occ_voxel_feat'] = self.pnet_model(inp_feat=pnet_inp, vox2point_idx=data_dict['revidx'])

    def _query_kwargs(self, data_dict):
        m = self.opt.model
        return dict(part_size=float(data_dict['part_size']), pos_encode=bool(m.pos_encode), multires=int(m.multires),
                    multires_views=int(m.multires_views), intersect_pos_type=str(m.intersect_pos_type),
                    roi_inp_bbox=int(m.roi_inp_bbox), roi_out_bbox=int(m.roi_out_bbox), n_iter=int(m.n_iter),
                    use_sigmoid=bool(m.use_sigmoid), offset_range=tuple(float(v) for v in self.opt.grid.offset_range))

    def get_occ_vox_bound(self, data_dict):
        """Replacement for ``LIDF.get_occ_vox_bound`` (reference pipeline.py:162-201): same grid set-up, same ``data_dict``
        entries, the (image, cell) ``torch.unique`` replaced by the occupancy-bitmap kernels (``lidf_voxelize_*``)."""
        from implicit_depth_b200.utils import point_utils
        dev = data_dict['valid_xyz'].device
        xmin = torch.Tensor(XMIN).float().to(dev)
        xmax = torch.Tensor(XMAX).float().to(dev)
        min_bb = torch.min(xmax - xmin).item()
        part_size = min_bb / self.opt.grid.res
        xmin = xmin - 0.5 * part_size                                  # half voxel margin on each side, :172-173
        xmax = xmax + 0.5 * part_size
        dims = point_utils._grid_dims(xmin, xmax, part_size)
        n_images = int(data_dict['bs']) if 'bs' in data_dict else int(data_dict['valid_bid'].max().item()) + 1
        occ, voxel_bound, revidx, valid_v_pid, valid_v_rel_coord = point_utils.voxelize(
            data_dict['valid_xyz'].float().contiguous(), data_dict['valid_bid'].long().contiguous(), xmin, part_size, dims,
            n_images)
        if occ.shape[0] == 0:
            print('No occupied voxel', data_dict.get('item_path'))
            return False
        data_dict.update({'xmin': xmin, 'part_size': part_size, 'revidx': revidx, 'valid_v_pid': valid_v_pid,
                          'valid_v_rel_coord': valid_v_rel_coord, 'occ_vox_bid': occ[:, 0],
                          'occ_vox_global_coord': occ[:, 1:], 'voxel_bound': voxel_bound})
        return True

    def compute_ray_aabb(self, data_dict):
        """Replacement for ``LIDF.compute_ray_aabb`` (reference pipeline.py:271-296): the same slab test, but the pair list
        and the per-pair enter/leave distances come straight from the kernel (``ray_aabb.pairs``) instead of
        ``torch.nonzero`` over a dense [V,R] mask plus a ``dist[vox, ray]`` lookup.  Writes ``occ_vox_intersect_idx``,
        ``miss_ray_intersect_idx`` (reference order) and ``intersect_dist`` [P,2]; the dense ``mask`` / ``dist`` entries of
        the reference are not produced -- their only reader is ``get_embedding`` (pipeline.py:345), replaced above."""
        from implicit_depth_b200.extensions.ray_aabb.jit import ray_aabb
        vox, ray, dist = ray_aabb.pairs(data_dict['miss_ray_dir'].contiguous(), data_dict['voxel_bound'].contiguous(),
                                        data_dict['miss_bid'].int().contiguous(), data_dict['occ_vox_bid'].int().contiguous(),
                                        order=self.pair_order)
        if vox.shape[0] == 0:
            print('No miss ray and occ vox intersection pair', data_dict.get('item_path'))
            return False
        data_dict.update({'occ_vox_intersect_idx': vox, 'miss_ray_intersect_idx': ray, 'intersect_dist': dist,
                          'pairs_ray_major': self.pair_order == 'ray'})
        return True

    def compute_pair_label(self, data_dict, gt_pos):
        """``pcl_label`` / ``pcl_label_float`` of ``LIDF.compute_gt`` (reference pipeline.py:303-309) without the dense
        [V,R] point-in-voxel mask."""
        from implicit_depth_b200.extensions.pcl_aabb.jit import pcl_aabb
        lab = pcl_aabb.pair_label(gt_pos.float().contiguous(), data_dict['voxel_bound'].contiguous(),
                                  data_dict['miss_bid'].int().contiguous(), data_dict['occ_vox_bid'].int().contiguous(),
                                  data_dict['occ_vox_intersect_idx'].contiguous(), data_dict['miss_ray_intersect_idx'].contiguous())
        data_dict.update({'gt_pos': gt_pos, 'pcl_label': lab.long(), 'pcl_label_float': lab})

    def compute_ray_loss(self, data_dict):
        """The ray-keyed statistics of ``LIDF.compute_loss`` (reference pipeline.py:472, :482-486, :553-567; shipped setting
        ``hard_neg: False``) from the path's outputs, forward only (evaluation / logging): returns ``pos_loss``, ``prob_loss``,
        ``acc``, ``err`` as 0-dim tensors without a host sync, plus ``log_softmax`` / ``pred_label`` / ``gt_label``.  The
        image-space terms (surface normals, smoothness, depth metrics) stay with the reference code."""
        return lidf_query.ray_loss(data_dict['pred_prob_end'].contiguous(), data_dict['pred_prob_end_softmax'].contiguous(),
                                   data_dict['miss_ray_intersect_idx'].contiguous(),
                                   data_dict['pcl_label_float'].contiguous(), int(data_dict['total_miss_sample_num']),
                                   data_dict['pred_pos'].contiguous(), data_dict['gt_pos'].float().contiguous())

    def compute_loss_eval(self, data_dict, exp_type, epoch):
        """Forward-only mirror of ``LIDF.compute_loss`` (reference pipeline.py:468-650) for evaluation / logging, shipped
        setting ``hard_neg: False``: the ray-keyed terms come from ``lidf_ray_loss``, the image-space terms from
        ``lidf_image_loss``; ``loss_net`` is assembled with the same weights and epoch gates (:538-543).  Returns the
        reference's ``loss_dict`` keys as 0-dim tensors (no host sync).  For ``exp_type != 'train'`` the depth metrics
        (:570-618) are added by ``lidf_depth_metrics_*`` -- both the ``bs != 1`` branch over rays and the ``bs == 1`` branch,
        whose cv2 nearest-neighbour resampling to 256x144 is done on the device instead of through the host.  Training
        needs gradients and keeps using the reference's ``compute_loss``."""
        L = self.opt.loss
        if getattr(L, 'hard_neg', False):
            raise NotImplementedError('compute_loss_eval covers hard_neg: False (the shipped lidf YAMLs)')
        bs, h, w = data_dict['bs'], data_dict['h'], data_dict['w']
        ray = self.compute_ray_loss(data_dict)
        xyz = data_dict['xyz_flat'] if exp_type == 'train' else data_dict['xyz_corrupt_flat']      # :494-500
        img = lidf_query.image_loss(xyz.float().contiguous(), data_dict['miss_bid'].long().contiguous(),
                                    data_dict['miss_flat_img_id'].long().contiguous(), data_dict['pred_pos'].contiguous(),
                                    data_dict['gt_pos'].float().contiguous(), h, w, want_normal_imgs=True)
        loss_net = L.pos_w * ray['pos_loss'] + L.prob_w * ray['prob_loss']
        if L.surf_norm_w > 0 and epoch >= L.surf_norm_epo:
            loss_net = loss_net + L.surf_norm_w * img['surf_norm_loss']
        if L.smooth_w > 0 and epoch >= L.smooth_epo:
            loss_net = loss_net + L.smooth_w * img['smooth_loss']
        data_dict.update({'gt_surf_norm_img': img['gt_surf_norm_img'], 'pred_surf_norm_img': img['pred_surf_norm_img']})
        loss_dict = {'pos_loss': ray['pos_loss'], 'prob_loss': ray['prob_loss'], 'surf_norm_loss': img['surf_norm_loss'],
                     'smooth_loss': img['smooth_loss'], 'loss_net': loss_net, 'acc': ray['acc'], 'err': ray['err'],
                     'angle_err': img['angle_err']}
        if exp_type != 'train':                                                                  # depth metrics, :570-618
            if bs != 1:
                m = lidf_query.depth_metrics(data_dict['pred_pos'].contiguous(), data_dict['gt_pos'].float().contiguous())
            else:       # the reference's cv2.resize(256x144, INTER_NEAREST) host round trip, done on the device
                m = lidf_query.depth_metrics(data_dict['pred_pos'].contiguous(), xyz_flat=data_dict['xyz_flat'].float().contiguous(),
                                             xyz_corrupt_flat=data_dict['xyz_corrupt_flat'].float().contiguous(),
                                             corrupt_mask=data_dict['corrupt_mask'].float().contiguous(),
                                             miss_flat_img_id=data_dict['miss_flat_img_id'].long().contiguous(), h=h, w=w)
            loss_dict.update({k: m[k] for k in ('a1', 'a2', 'a3', 'rmse', 'rmse_log', 'log10', 'abs_rel', 'mae', 'sq_rel')})
        return loss_dict

    def get_pred(self, data_dict, exp_type, epoch):
        if self.opt.model.scatter_type != 'Maxpool':
            raise NotImplementedError('Does not support Scatter Type: {}'.format(self.opt.model.scatter_type))
        use_label = exp_type == 'train' and epoch < self.opt.model.maxpool_label_epo       # pipeline.py:444
        dist = data_dict['dist'] if 'dist' in data_dict else data_dict['intersect_dist']
        args = (data_dict['full_rgb_feat'].float().contiguous(), data_dict['occ_voxel_feat'].float().contiguous(),
                data_dict['miss_ray_dir'].contiguous(), data_dict['miss_img_ind'].long().contiguous(),
                data_dict['miss_bid'].long().contiguous(), data_dict['voxel_bound'].contiguous(),
                data_dict['occ_vox_intersect_idx'].contiguous(), data_dict['miss_ray_intersect_idx'].contiguous(),
                dist.contiguous())
        kw = dict(pcl_label_float=data_dict['pcl_label_float'].contiguous() if use_label else None,
                  mlp_impl=self.mlp_impl, pairs_ray_major=bool(data_dict.get('pairs_ray_major', False)),
                  **self._query_kwargs(data_dict))
        needs_grad = torch.is_grad_enabled() and (
            args[0].requires_grad or args[1].requires_grad
            or any(p.requires_grad for p in self.offset_dec.parameters())
            or any(p.requires_grad for p in self.prob_dec.parameters()))
        if needs_grad:
            # training: the same fused forward, recorded as ONE autograd node whose backward is lidf_query_backward
            # (tcgen05 dgrad / wgrad kernels); DDP sees ordinary .grad tensors on the decoder parameters
            out = lidf_query_autograd(args, self.offset_dec, self.prob_dec, kw,
                                      winner_only=bool(self.winner_only) and self.mlp_impl != "simt_fp32")
        else:
            out = lidf_query.forward(*args, self.offset_dec, self.prob_dec, want_roi_feat=True,
                                     winner_only=bool(self.winner_only) and self.mlp_impl != "simt_fp32", **kw)
        assert out['pred_pos'].shape[0] == data_dict['total_miss_sample_num']
        data_dict.update({k: out[k] for k in ('pair_pred_pos', 'max_pair_id', 'pred_prob_end', 'pred_prob_end_softmax', 'pred_pos',
                                              'pred_offset', 'roi_feat_per_ray') if k in out})


class _LidfQueryFn(torch.autograd.Function):
    """``lidf_query.forward`` as one autograd node (reference: the op chain of pipeline.py:338-466 recorded op by op).
    Differentiable outputs: pred_offset, pred_prob_end, pair_pred_pos, pred_pos.  pred_prob_end_softmax and max_pair_id
    come from a detached soft-max in the reference (:442) and carry no gradient; neither does the cached ROI feature."""

    @staticmethod
    def forward(ctx, meta, full_rgb_feat, occ_voxel_feat, *params):
        n_off = len(meta['off_keys'])
        off = dict(zip(meta['off_keys'], params[:n_off]))
        prob = dict(zip(meta['prob_keys'], params[n_off:]))
        wo = bool(meta.get('winner_only', False))
        out = lidf_query.forward(full_rgb_feat, occ_voxel_feat, *meta['index_args'], off, prob, want_roi_feat=True,
                                 save_for_backward=True, winner_only=wo, **meta['kw'])
        ctx.meta = meta
        ctx.fwd_out = {k: out[k] for k in ('pred_offset', 'pred_prob_end', 'max_pair_id', 'ief_iter', 'pred_offset_ray') if k in out}
        ctx.save_for_backward(full_rgb_feat, occ_voxel_feat, *params)
        if wo:      # per-pair offset outputs do not exist in this mode: empty placeholders keep the node's signature
            e0, e1 = out['pred_prob_end'].new_empty(0), out['pred_prob_end'].new_empty(0)
            ctx.mark_non_differentiable(out['pred_prob_end_softmax'], out['max_pair_id'], out['roi_feat_per_ray'], e0, e1)
            return (e0, out['pred_prob_end'], e1, out['pred_prob_end_softmax'], out['max_pair_id'], out['pred_pos'],
                    out['roi_feat_per_ray'])
        ctx.mark_non_differentiable(out['pred_prob_end_softmax'], out['max_pair_id'], out['roi_feat_per_ray'])
        return (out['pred_offset'], out['pred_prob_end'], out['pair_pred_pos'], out['pred_prob_end_softmax'],
                out['max_pair_id'], out['pred_pos'], out['roi_feat_per_ray'])

    @staticmethod
    def backward(ctx, g_off, g_prob, g_pair, _g_soft, _g_arg, g_pos, _g_roi):
        meta = ctx.meta
        full_rgb_feat, occ_voxel_feat, *params = ctx.saved_tensors
        n_off = len(meta['off_keys'])
        off = dict(zip(meta['off_keys'], params[:n_off]))
        prob = dict(zip(meta['prob_keys'], params[n_off:]))
        kw = {k: v for k, v in meta['kw'].items() if k != 'pcl_label_float'}
        wo = bool(meta.get('winner_only', False))
        res = lidf_query.backward(full_rgb_feat, occ_voxel_feat, *meta['index_args'], off, prob, ctx.fwd_out,
                                  g_pred_pos=g_pos, g_pred_prob_end=g_prob, g_pred_offset=None if wo else g_off,
                                  g_pair_pred_pos=None if wo else g_pair, winner_only=wo,
                                  need_feat_grad=ctx.needs_input_grad[1], need_vox_grad=ctx.needs_input_grad[2], **kw)
        grads = [res['offset_dec'][k] for k in meta['off_keys']] + [res['prob_dec'][k] for k in meta['prob_keys']]
        grads = [g if need else None for g, need in zip(grads, ctx.needs_input_grad[3:])]
        return (None, res['full_rgb_feat'], res['occ_voxel_feat'], *grads)


def lidf_query_autograd(args, offset_dec, prob_dec, kw, winner_only=False):
    """Training entry: ``args`` = the nine tensor inputs of ``lidf_query.forward`` (features first), decoders as modules.
    There is no CPU / eager fallback: the tensors must live on a CUDA device."""
    full_rgb_feat, occ_voxel_feat = args[0], args[1]
    if not full_rgb_feat.is_cuda:
        raise RuntimeError('lidf_query: the training path runs on the sm_100a kernels only (no CPU fallback)')
    off_sd = dict(offset_dec.named_parameters())
    prob_sd = dict(prob_dec.named_parameters())
    meta = dict(off_keys=list(off_sd), prob_keys=list(prob_sd), index_args=tuple(a.detach() for a in args[2:]), kw=kw,
                winner_only=bool(winner_only))
    names = ('pred_offset', 'pred_prob_end', 'pair_pred_pos', 'pred_prob_end_softmax', 'max_pair_id', 'pred_pos',
             'roi_feat_per_ray')
    outs = _LidfQueryFn.apply(meta, full_rgb_feat, occ_voxel_feat, *off_sd.values(), *prob_sd.values())
    out = dict(zip(names, outs))
    if winner_only:
        out.pop('pred_offset'); out.pop('pair_pred_pos')
    return out


class LIDF(LIDFQueryMixin, nn.Module):
    """Stand-alone holder with the reference's attribute names (pipeline.py:39-89); producers are injected."""

    def __init__(self, opt, device, resnet_model=None, pnet_model=None):
        super().__init__()
        self.opt = opt
        self.device = device
        m = opt.model
        if m.pos_encode:
            self.embed_fn, embed_ch = im_net.get_embedder(m.multires)
            self.embeddirs_fn, embeddirs_ch = im_net.get_embedder(m.multires_views)
        else:
            self.embed_fn, embed_ch = im_net.get_embedder(m.multires, i=-1)
            self.embeddirs_fn, embeddirs_ch = im_net.get_embedder(m.multires_views, i=-1)
        self.resnet_model = resnet_model          # cuDNN backbone: out of scope, injected (reference pipeline.py:50-54)
        if pnet_model is None and m.pnet_model_type == 'twostage':                                   # pipeline.py:56-58
            pnet_model = pnet.PointNet2Stage(input_channels=m.pnet_in, output_channels=m.pnet_out, gf_dim=m.pnet_gf).to(device)
        self.pnet_model = pnet_model
        dec_inp_dim = m.pnet_out + m.rgb_out * (m.roi_out_bbox ** 2) + 2 * embed_ch + embeddirs_ch   # pipeline.py:64-65
        if m.offdec_type == 'IMNET':
            self.offset_dec = im_net.IMNet(inp_dim=dec_inp_dim, out_dim=1, gf_dim=m.imnet_gf,
                                           use_sigmoid=m.use_sigmoid).to(device)
        elif m.offdec_type == 'IEF':
            self.offset_dec = im_net.IEF(device, inp_dim=dec_inp_dim, out_dim=1, gf_dim=m.imnet_gf, n_iter=m.n_iter,
                                         use_sigmoid=m.use_sigmoid).to(device)
        else:
            raise NotImplementedError('Does not support Offset Decoder Type: {}'.format(m.offdec_type))
        if m.probdec_type == 'IMNET':
            self.prob_dec = im_net.IMNet(inp_dim=dec_inp_dim, out_dim=1, gf_dim=m.imnet_gf,
                                         use_sigmoid=m.use_sigmoid).to(device)
        else:
            raise NotImplementedError('Does not support Prob Decoder Type: {}'.format(m.probdec_type))

    def forward(self, data_dict, exp_type='test', epoch=0):
        """Hot-path segment of the reference forward (pipeline.py:706-708): get_embedding then get_pred."""
        self.get_embedding(data_dict)
        self.get_pred(data_dict, exp_type, epoch)
        return data_dict


class RefineDecoderMixin:
    """Decoder tail of ``RefineNet.get_pred_refine`` (reference pipeline.py:1018-1029) as one fused call."""

    mlp_impl = "auto"

    def refine_end_voxel(self, data_dict, pred_pos):
        """``end_voxel_id`` of ``RefineNet.get_pred_refine`` (reference pipeline.py:939-944): the voxel of each ray's
        arg-max pair, raised to the largest occupied voxel that contains the predicted point -- one kernel instead of
        the dense pcl_aabb mask + nonzero + torch_scatter.scatter(reduce='max')."""
        from implicit_depth_b200.extensions.pcl_aabb.jit import pcl_aabb
        vox = data_dict['occ_vox_intersect_idx']
        dummy = torch.cat((vox, torch.zeros(1, dtype=vox.dtype, device=vox.device)), 0)      # concat_dummy, :924,:942
        end_voxel_id = dummy[data_dict['max_pair_id']].contiguous()
        return pcl_aabb.end_voxel(pred_pos.float().contiguous(), data_dict['voxel_bound'].contiguous(),
                                  data_dict['miss_bid'].int().contiguous(), data_dict['occ_vox_bid'].int().contiguous(),
                                  end_voxel_id)

    def refine_decoder_tail(self, data_dict, pred_pos, end_voxel_id, occ_voxel_feat, rgb_feat_per_ray):
        """pipeline.py:1016-1029: gather of the end-voxel feature + decoder + position update, one fused call (the gather
        and the voxel centre for 'rel' positions happen inside the kernel)."""
        r = self.opt.refine
        needs_grad = torch.is_grad_enabled() and (occ_voxel_feat.requires_grad or pred_pos.requires_grad
                                                  or any(p.requires_grad for p in self.offset_dec.parameters()))
        if needs_grad:
            # training (train_refine.yaml optimises pnet_model + offset_dec): the same maths on differentiable torch ops,
            # reference pipeline.py:1016-1029, because the native backward does not exist yet
            end_voxel_bound = data_dict['voxel_bound'][end_voxel_id]
            end_voxel_center = (end_voxel_bound[:, :3] + end_voxel_bound[:, 3:]) / 2.
            inp_pos = pred_pos - end_voxel_center if r.intersect_pos_type == 'rel' else pred_pos
            inp_embed = torch.cat((occ_voxel_feat[end_voxel_id], rgb_feat_per_ray, self.embed_fn(inp_pos),
                                   self.embeddirs_fn(data_dict['miss_ray_dir'])), -1)
            pred_refine_offset = self.offset_dec(inp_embed)
            r0, r1 = r.offset_range
            return pred_pos + (pred_refine_offset * (r1 - r0) + r0) * data_dict['miss_ray_dir']
        return lidf_query.refine_forward(
            pred_pos.contiguous(), data_dict['miss_ray_dir'].contiguous(), None, None, rgb_feat_per_ray.contiguous(),
            self.offset_dec, occ_voxel_feat=occ_voxel_feat.float().contiguous(), end_voxel_id=end_voxel_id.long().contiguous(),
            voxel_bound=data_dict['voxel_bound'].contiguous(),
            pos_encode=bool(r.pos_encode), multires=int(r.multires), multires_views=int(r.multires_views),
            intersect_pos_type=str(r.intersect_pos_type), n_iter=int(r.n_iter), use_sigmoid=bool(r.use_sigmoid),
            offset_range=tuple(float(v) for v in r.offset_range), mlp_impl=self.mlp_impl)


    def get_pred_refine(self, data_dict, pred_pos, exp_type, cur_iter):
        """Replacement for ``RefineNet.get_pred_refine`` (reference pipeline.py:922-1030), inference path: the same steps in
        the same order with this repo's kernels -- end-voxel lookup (``pcl_aabb.end_voxel``), the per-ray ROI feature cached
        by stage 1 (``data_dict['roi_feat_per_ray']``) instead of a second ``roi_align`` (:952-970), the PointNet re-run over
        the valid points plus the predicted points (``self.pnet_model``; ``models.pointnet.PointNet2Stage`` runs it natively),
        and the fused decoder tail.  The training-time perturbation (:926-937) is reproduced as written."""
        bs, h, w = data_dict['bs'], data_dict['h'], data_dict['w']
        r = self.opt.refine
        if exp_type == 'train' and r.perturb and cur_iter == 0 and np.random.random() < r.perturb_prob:
            prob = np.random.random()
            if prob < 0.5:
                noise = np.random.random() * (0 + 0.05) - 0.05
            elif prob < 0.8:
                noise = np.random.random() * (0.05 - 0)
            elif prob < 0.9:
                noise = np.random.random() * (-0.05 + 0.1) - 0.1
            else:
                noise = np.random.random() * (0.1 - 0.05) + 0.05
            pred_pos = pred_pos + noise * data_dict['miss_ray_dir']
        end_voxel_id = self.refine_end_voxel(data_dict, pred_pos)                                   # :939-944
        rgb_img_flat = data_dict['rgb_img'].permute(0, 2, 3, 1).contiguous().reshape(bs, -1, 3)     # :972-973
        miss_rgb = rgb_img_flat[data_dict['miss_bid'], data_dict['miss_flat_img_id']]
        end_voxel_bound = data_dict['voxel_bound'][end_voxel_id]                                    # :975-977
        end_voxel_center = (end_voxel_bound[:, :3] + end_voxel_bound[:, 3:]) / 2.
        if r.pnet_pos_type == 'rel':                                                                # :979-984
            pred_inp = torch.cat((pred_pos - end_voxel_center, miss_rgb), 1)
        else:
            pred_inp = torch.cat((pred_pos, miss_rgb), 1)
        if exp_type != 'train' and self.opt.mask_type == 'all' and r.use_all_pix == False:          # noqa: E712  (:985-995)
            zero_pixel_idx = torch.nonzero(1 - data_dict['valid_mask'], as_tuple=False)
            sel = (zero_pixel_idx[:, 0], zero_pixel_idx[:, 1], zero_pixel_idx[:, 2])
            new_pred_inp = pred_inp.reshape(bs, h, w, pred_inp.shape[-1])[sel]
            new_end_voxel_id = end_voxel_id.reshape(bs, h, w)[sel]
        else:
            new_pred_inp, new_end_voxel_id = pred_inp, end_voxel_id
        valid_v_rgb = data_dict['valid_rgb'][data_dict['valid_v_pid']]                              # :999-1007
        if r.pnet_pos_type == 'rel':
            pnet_inp = torch.cat((data_dict['valid_v_rel_coord'], valid_v_rgb), -1)
        elif r.pnet_pos_type == 'abs':
            pnet_inp = torch.cat((data_dict['valid_xyz'][data_dict['valid_v_pid']], valid_v_rgb), -1)
        else:
            raise NotImplementedError('Does not support Pnet pos type: {}'.format(r.pnet_pos_type))
        final_pnet_inp = torch.cat((pnet_inp, new_pred_inp), 0)                                     # :1008-1010
        final_revidx = torch.cat((data_dict['revidx'], new_end_voxel_id), 0)
        occ_voxel_feat = self.pnet_model(inp_feat=final_pnet_inp, vox2point_idx=final_revidx)       # :1011-1014
        roi = data_dict.get('roi_feat_per_ray')
        if roi is None:     # stage 1 ran on the stock reference LIDF (no cached per-ray feature): recompute it, pipeline.py:952-970
            roi = lidf_query.roi_align_rays(data_dict['full_rgb_feat'].float().contiguous(), data_dict['miss_img_ind'].long().contiguous(),
                                            data_dict['miss_bid'].long().contiguous(), int(self.opt.model.roi_inp_bbox))
        return self.refine_decoder_tail(data_dict, pred_pos, end_voxel_id, occ_voxel_feat, roi)


class RefineNet(RefineDecoderMixin, nn.Module):
    """Stand-alone holder of RefineNet.offset_dec (pipeline.py:722-758)."""

    def __init__(self, opt, device, pnet_model=None):
        super().__init__()
        self.opt = opt
        self.device = device
        r, m = opt.refine, opt.model
        if r.pos_encode:
            self.embed_fn, embed_ch = im_net.get_embedder(r.multires)
            self.embeddirs_fn, embeddirs_ch = im_net.get_embedder(r.multires_views)
        else:
            self.embed_fn, embed_ch = im_net.get_embedder(r.multires, i=-1)
            self.embeddirs_fn, embeddirs_ch = im_net.get_embedder(r.multires_views, i=-1)
        if pnet_model is None and r.pnet_model_type == 'twostage':                                   # pipeline.py:733-735
            pnet_model = pnet.PointNet2Stage(input_channels=r.pnet_in, output_channels=r.pnet_out, gf_dim=r.pnet_gf).to(device)
        self.pnet_model = pnet_model
        dec_inp_dim = r.pnet_out + embed_ch + embeddirs_ch + m.rgb_out * (m.roi_out_bbox ** 2)   # pipeline.py:740-742
        if r.offdec_type == 'IMNET':
            self.offset_dec = im_net.IMNet(inp_dim=dec_inp_dim, out_dim=1, gf_dim=r.imnet_gf,
                                           use_sigmoid=r.use_sigmoid).to(device)
        elif r.offdec_type == 'IEF':
            self.offset_dec = im_net.IEF(device, inp_dim=dec_inp_dim, out_dim=1, gf_dim=r.imnet_gf, n_iter=r.n_iter,
                                         use_sigmoid=r.use_sigmoid).to(device)
        else:
            raise NotImplementedError('Does not support Offset Decoder Type: {}'.format(r.offdec_type))


class _NS:
    """Tiny attribute namespace standing in for the reference's ``Params`` (src/opt.py) in tests / bench."""

    def __init__(self, **kw):
        self.__dict__.update(kw)


def default_opt(**over):
    """opt.model / opt.grid / opt.refine with the values of the shipped YAMLs (train_lidf.yaml, train_refine.yaml)."""
    model = dict(pos_encode=True, multires=8, multires_views=4, intersect_pos_type='abs', rgb_in=3, rgb_out=32,
                 roi_inp_bbox=8, roi_out_bbox=2, pnet_in=6, pnet_out=128, pnet_gf=32, pnet_pos_type='rel',
                 offdec_type='IEF', n_iter=2, probdec_type='IMNET', imnet_gf=64, scatter_type='Maxpool', pnet_model_type='twostage',
                 use_sigmoid=False, maxpool_label_epo=6)
    refine = dict(pos_encode=True, multires=8, multires_views=4, intersect_pos_type='abs', pnet_out=128,
                  offdec_type='IEF', n_iter=2, imnet_gf=64, use_sigmoid=False, offset_range=[-0.2, 0.2], forward_times=2,
                  pnet_in=6, pnet_gf=32, pnet_pos_type='rel', pnet_model_type='twostage', perturb=False, perturb_prob=0.,
                  use_all_pix=True)
    grid = dict(res=8, offset_range=[0., 1.])
    loss = dict(pos_loss_type='single', prob_loss_type='ray', hard_neg=False, hard_neg_ratio=None, pos_w=100.0, prob_w=0.5,
                surf_norm_w=10.0, surf_norm_epo=0, smooth_w=0.0, smooth_epo=0)
    for k, v in over.items():
        sect, key = k.split('.')
        {'model': model, 'refine': refine, 'grid': grid, 'loss': loss}[sect][key] = v
    return _NS(model=_NS(**model), refine=_NS(**refine), grid=_NS(**grid), loss=_NS(**loss), gpu_id=0, mask_type='all')
