"""GPU parity for the ray_aabb / pcl_aabb rows (include/lidf_aabb.h), through the C ABI: bit-exact against
(1) golden vectors from the reference's own kernels, (2) the numpy oracle on seeded inputs incl. empty / ragged /
multi-tile shapes, (3) at BASELINE config-2 size, the compact pair list against torch.nonzero of our dense output."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN_DIR
from oracle import aabb_oracle as A

pytestmark = pytest.mark.gpu
AABB_CASES = ["aabb_grid_2x12x16", "aabb_grid_1x9x11", "aabb_edge"]


def _ext():
    from implicit_depth_b200.extensions.pcl_aabb.jit import pcl_aabb
    from implicit_depth_b200.extensions.ray_aabb.jit import ray_aabb
    return ray_aabb, pcl_aabb


def _c(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a))
    return (t.to(dtype) if dtype is not None else t).cuda()


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def _check_ray(ray_dir, vb, rb, xb, want_mask, want_dist, want_pairs):
    ray_aabb, _ = _ext()
    args = (_c(ray_dir, torch.float32), _c(vb, torch.float32), _c(rb, torch.int32), _c(xb, torch.int32))
    mask, dist = ray_aabb.forward(*args)
    assert mask.dtype == torch.int32 and dist.dtype == torch.float32
    assert np.array_equal(mask.cpu().numpy(), want_mask)
    assert np.array_equal(bits(dist.cpu().numpy()), bits(want_dist))
    vox, ray, pd = ray_aabb.pairs(*args)
    assert vox.dtype == torch.int64 and ray.dtype == torch.int64
    assert np.array_equal(vox.cpu().numpy(), want_pairs[0]) and np.array_equal(ray.cpu().numpy(), want_pairs[1])
    assert np.array_equal(bits(pd.cpu().numpy()), bits(want_pairs[2]))
    # ray-major list (lidf_ray_aabb_pairs_ray_major_*): the same pairs, stably re-sorted by ray (voxel stays ascending
    # within a ray because the nonzero order is voxel-major), bit-identical distances, plus its CSR offsets
    vox2, ray2, pd2, start = ray_aabb.pairs(*args, order="ray", return_ray_start=True)
    o = np.argsort(want_pairs[1], kind="stable")
    assert np.array_equal(vox2.cpu().numpy(), want_pairs[0][o]) and np.array_equal(ray2.cpu().numpy(), want_pairs[1][o])
    assert np.array_equal(bits(pd2.cpu().numpy()), bits(want_pairs[2][o]))
    R = ray_dir.shape[0]
    assert start.dtype == torch.int32 and np.array_equal(start.cpu().numpy(), np.searchsorted(want_pairs[1][o], np.arange(R + 1)))


@pytest.mark.parametrize("name", AABB_CASES)
def test_ray_aabb_golden_reference_kernel_outputs(name):
    z = dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz")))
    _check_ray(z["ray_dir"], z["voxel_bound"], z["ray_bid"], z["voxel_bid"], z["ref_mask"], z["ref_dist"],
               (z["ref_pair_vox"], z["ref_pair_ray"], z["ref_pair_dist"]))


@pytest.mark.parametrize("name", AABB_CASES)
def test_pcl_aabb_golden_reference_kernel_outputs(name):
    _, pcl_aabb = _ext()
    z = dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz")))
    vb, xb = _c(z["voxel_bound"]), _c(z["voxel_bid"], torch.int32)
    m = pcl_aabb.forward(_c(z["pts"]), vb, _c(z["pts_bid"], torch.int32), xb)
    assert np.array_equal(m.cpu().numpy(), z["ref_pcl_mask"])
    rp, rb = _c(z["ray_pts"]), _c(z["ray_bid"], torch.int32)
    assert np.array_equal(pcl_aabb.forward(rp, vb, rb, xb).cpu().numpy(), z["ref_pcl_mask_rays"])
    lab = pcl_aabb.pair_label(rp, vb, rb, xb, _c(z["ref_pair_vox"]), _c(z["ref_pair_ray"]))
    assert np.array_equal(lab.cpu().numpy(), z["ref_pair_label"])
    end = pcl_aabb.end_voxel(rp, vb, rb, xb, _c(z["end_voxel_start"]))
    assert np.array_equal(end.cpu().numpy(), z["ref_end_voxel"])


@pytest.mark.parametrize("R,V,nimg", [(1, 1, 1), (1023, 7, 2), (1024, 3, 1), (1025, 33, 3), (5000, 300, 4), (3, 700, 2)])
def test_ray_and_pcl_aabb_vs_oracle_seeded(R, V, nimg):
    _, pcl_aabb = _ext()
    rng = np.random.default_rng(R * 1000 + V)
    d = rng.normal(size=(R, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    d[rng.random((R, 3)) < 0.03] = 0.0
    d = d.astype(np.float32)
    lo = rng.uniform(-1, 1, size=(V, 3)); vb = np.concatenate((lo, lo + rng.uniform(0, 1.2, size=(V, 3))), 1).astype(np.float32)
    rb = np.sort(rng.integers(0, nimg, size=R)).astype(np.int32)
    xb = rng.integers(0, nimg, size=V).astype(np.int32)              # voxel image ids unsorted on purpose
    mask, dist = A.ray_aabb_dense(d, vb, rb, xb)
    _check_ray(d, vb, rb, xb, mask, dist, A.ray_aabb_pairs(d, vb, rb, xb))
    pts = rng.uniform(-1.2, 2.2, size=(R, 3)).astype(np.float32)
    args = (_c(pts), _c(vb), _c(rb), _c(xb))
    assert np.array_equal(pcl_aabb.forward(*args).cpu().numpy(), A.pcl_aabb_dense(pts, vb, rb, xb))
    start = rng.integers(0, V, size=R).astype(np.int64)
    assert np.array_equal(pcl_aabb.end_voxel(*args, _c(start)).cpu().numpy(), A.pcl_end_voxel(pts, vb, rb, xb, start))
    vox, ray, _ = A.ray_aabb_pairs(d, vb, rb, xb)
    if vox.size:
        assert np.array_equal(pcl_aabb.pair_label(*args, _c(vox), _c(ray)).cpu().numpy(), A.pcl_pair_label(pts, vb, rb, xb, vox, ray))


def test_aabb_empty_inputs_and_no_hits():
    ray_aabb, pcl_aabb = _ext()
    f = lambda *s: torch.zeros(*s, dtype=torch.float32, device="cuda")
    i = lambda *s: torch.zeros(*s, dtype=torch.int32, device="cuda")
    for R, V in [(0, 4), (5, 0), (0, 0)]:
        mask, dist = ray_aabb.forward(f(R, 3), f(V, 6), i(R), i(V))
        assert tuple(mask.shape) == (V, R) and tuple(dist.shape) == (V, R, 2)
        vox, ray, pd = ray_aabb.pairs(f(R, 3), f(V, 6), i(R), i(V))
        assert vox.numel() == 0 and ray.numel() == 0 and tuple(pd.shape) == (0, 2)
        vox, ray, pd, start = ray_aabb.pairs(f(R, 3), f(V, 6), i(R), i(V), order="ray", return_ray_start=True)
        assert vox.numel() == 0 and ray.numel() == 0 and tuple(pd.shape) == (0, 2) and int(start.abs().sum()) == 0
        assert tuple(pcl_aabb.forward(f(R, 3), f(V, 6), i(R), i(V)).shape) == (V, R)
    # rays and voxels of different images never pair up
    d = torch.tensor([[0., 0., 1.]] * 2000, device="cuda")
    vb = torch.tensor([[-1., -1., 0.5, 1., 1., 1.]] * 3, device="cuda")
    vox, ray, pd = ray_aabb.pairs(d, vb, i(2000), i(3) + 1)
    assert vox.numel() == 0
    mask, dist = ray_aabb.forward(d, vb, i(2000), i(3) + 1)
    assert int(mask.sum()) == 0 and float(dist.abs().sum()) == 0.0
    vox, ray, pd = ray_aabb.pairs(d, vb, i(2000), i(3))
    assert vox.numel() == 6000 and torch.equal(pd, torch.tensor([[0.5, 1.0]], device="cuda").expand(6000, 2))
    with pytest.raises(RuntimeError, match="must be a CUDA tensor"):
        ray_aabb.forward(d.cpu(), vb, i(2000), i(3))
    with pytest.raises(RuntimeError, match="must be torch.int32"):
        ray_aabb.forward(d, vb, i(2000).long(), i(3))


def test_ray_aabb_pairs_equal_nonzero_of_dense_at_config2_size():
    """4 images of 320x240 all-pixel rays against 256 occupied cells of the 9^3 grid each (the bench geometry):
    the compact list must be exactly torch.nonzero(mask) / dist[vox, ray] of the dense drop-in, and a random slice of the
    dense output must match the oracle."""
    from implicit_depth_b200.synthetic import make_inputs
    ray_aabb, pcl_aabb = _ext()
    d = make_inputs(4, 240, 320, 1, V_img=256, seed=77, device="cuda")
    rd, vb = d["miss_ray_dir"], d["voxel_bound"]
    rb, xb = d["miss_bid"].int(), d["occ_vox_bid"].int()
    mask, dist = ray_aabb.forward(rd, vb, rb, xb)
    vox, ray, pd = ray_aabb.pairs(rd, vb, rb, xb)
    idx = torch.nonzero(mask.long(), as_tuple=False)                      # pipeline.py:281-285
    assert idx.shape[0] == vox.shape[0] and idx.shape[0] > 300000
    assert torch.equal(idx[:, 0], vox) and torch.equal(idx[:, 1], ray)
    assert torch.equal(dist[vox, ray].view(torch.int32), pd.view(torch.int32))   # pipeline.py:345, bit patterns
    assert torch.equal(rb[ray], xb[vox]) and bool((pd[:, 1] >= pd[:, 0]).all())
    vox2, ray2, pd2, start = ray_aabb.pairs(rd, vb, rb, xb, order="ray", return_ray_start=True)      # ray-major: a stable re-sort
    o = torch.sort(ray, stable=True).indices
    assert torch.equal(vox2, vox[o]) and torch.equal(ray2, ray[o]) and torch.equal(pd2.view(torch.int32), pd[o].view(torch.int32))
    assert torch.equal(start.long(), torch.searchsorted(ray2, torch.arange(rd.shape[0] + 1, device="cuda")))
    sel_v = torch.arange(0, vb.shape[0], 37, device="cuda"); sel_r = torch.arange(0, rd.shape[0], 41, device="cuda")
    om, od = A.ray_aabb_dense(rd[sel_r].cpu().numpy(), vb[sel_v].cpu().numpy(), rb[sel_r].cpu().numpy(), xb[sel_v].cpu().numpy())
    assert np.array_equal(mask[sel_v][:, sel_r].cpu().numpy(), om)
    assert np.array_equal(bits(dist[sel_v][:, sel_r].cpu().numpy()), bits(od))
    # points in the middle of each pair's [enter, leave] interval lie inside that pair's voxel -> label 1 everywhere
    mid = rd[ray] * (0.5 * (pd[:, 0:1] + pd[:, 1:2]))
    lab = pcl_aabb.pair_label(mid.contiguous(), vb, rb[ray].contiguous(), xb, vox, torch.arange(vox.shape[0], device="cuda"))
    assert float(lab.mean()) > 0.999
    # end voxel of those points, starting from 0, is >= the pair's own voxel and contains the point
    end = pcl_aabb.end_voxel(mid.contiguous(), vb, rb[ray].contiguous(), xb, torch.zeros_like(vox))
    ok = lab > 0
    assert bool((end[ok] >= vox[ok]).all())
    lab2 = pcl_aabb.pair_label(mid.contiguous(), vb, rb[ray].contiguous(), xb, end, torch.arange(vox.shape[0], device="cuda"))
    assert bool((lab2[ok] == 1).all())


def test_pipeline_mixin_real_geometry_chain_vs_oracle():
    """compute_ray_aabb -> compute_pair_label -> get_pred -> refine_end_voxel on real ray/voxel geometry (all-pixel rays
    against occupied cells of the 9^3 grid: ragged pair counts incl. rays without any pair), against the oracle chain
    aabb_oracle -> lidf_oracle."""
    from conftest import rel_err
    from implicit_depth_b200.models.pipeline import LIDF, RefineNet, default_opt
    from implicit_depth_b200.synthetic import make_inputs
    from oracle import lidf_oracle as O
    d = make_inputs(2, 30, 40, 1, V_img=60, seed=91)                # pair arrays of the generator are discarded below
    vox, ray, pd = A.ray_aabb_pairs(d["miss_ray_dir"].numpy(), d["voxel_bound"].numpy(), d["miss_bid"].numpy(), d["occ_vox_bid"].numpy())
    R = d["miss_ray_dir"].shape[0]
    assert 0 < np.unique(ray).size < R                              # some rays have no pair
    do = dict(d, occ_vox_intersect_idx=torch.from_numpy(vox), miss_ray_intersect_idx=torch.from_numpy(ray),
              intersect_dist=torch.from_numpy(pd))
    g = torch.Generator().manual_seed(92)
    off = O.init_decoder("IEF", 385, mode="trained", generator=g)
    prob = O.init_decoder("IMNET", 385, mode="trained", generator=g)
    want = O.lidf_query(do, dict(O.DEFAULT_CFG), off, prob, d["part_size"], dedup_rays=True)

    lidf = LIDF(default_opt(), torch.device("cuda")).cuda().eval()
    lidf.offset_dec.load_state_dict(off); lidf.prob_dec.load_state_dict(prob)
    dd = {k: (v.cuda() if isinstance(v, torch.Tensor) else v) for k, v in d.items()
          if k not in ("occ_vox_intersect_idx", "miss_ray_intersect_idx", "intersect_dist")}
    with torch.no_grad():
        assert lidf.compute_ray_aabb(dd) is True
        assert np.array_equal(dd["occ_vox_intersect_idx"].cpu().numpy(), vox)
        assert np.array_equal(dd["miss_ray_intersect_idx"].cpu().numpy(), ray)
        assert np.array_equal(bits(dd["intersect_dist"].cpu().numpy()), bits(pd))
        lidf.get_pred(dd, "test", 0)
    for k in ("pred_offset", "pred_prob_end", "pair_pred_pos", "pred_prob_end_softmax"):
        assert rel_err(dd[k].cpu(), want[k]) < 1e-3, k
    same = dd["max_pair_id"].cpu() == want["max_pair_id"]
    assert float(same.float().mean()) > 0.99
    empty = want["max_pair_id"] == vox.shape[0]
    assert bool(empty.any()) and torch.equal(dd["max_pair_id"].cpu()[empty], want["max_pair_id"][empty])
    # GT labels for points placed at the predicted positions, and the refine stage's end voxel
    gt_pos = dd["pred_pos"].clone()
    lidf.compute_pair_label(dd, gt_pos)
    lab = A.pcl_pair_label(gt_pos.cpu().numpy(), d["voxel_bound"].numpy(), d["miss_bid"].numpy(), d["occ_vox_bid"].numpy(), vox, ray)
    assert np.array_equal(dd["pcl_label_float"].cpu().numpy(), lab) and dd["pcl_label"].dtype == torch.int64
    refine = RefineNet(default_opt(), torch.device("cuda"))
    end = refine.refine_end_voxel(dd, dd["pred_pos"])
    start = np.concatenate((vox, [0]))[dd["max_pair_id"].cpu().numpy()]
    want_end = A.pcl_end_voxel(dd["pred_pos"].cpu().numpy(), d["voxel_bound"].numpy(), d["miss_bid"].numpy(), d["occ_vox_bid"].numpy(), start)
    assert np.array_equal(end.cpu().numpy(), want_end)
    # the same chain with the pair list emitted ray-major (pair_order = "ray": no regroup inside get_pred): per-ray results
    # bit-identical, per-pair results a permutation, labels / end voxels follow the permuted index tensors
    lidf.pair_order = "ray"
    d2 = {k: (v.cuda() if isinstance(v, torch.Tensor) else v) for k, v in d.items()
          if k not in ("occ_vox_intersect_idx", "miss_ray_intersect_idx", "intersect_dist")}
    with torch.no_grad():
        assert lidf.compute_ray_aabb(d2) is True and d2["pairs_ray_major"] is True
        lidf.get_pred(d2, "test", 0)
    from implicit_depth_b200.extensions.lidf_query.jit import lidf_query
    lidf_query.check_index_errors()
    o = torch.sort(dd["miss_ray_intersect_idx"], stable=True).indices
    assert torch.equal(d2["miss_ray_intersect_idx"], dd["miss_ray_intersect_idx"][o]) and torch.equal(d2["occ_vox_intersect_idx"], dd["occ_vox_intersect_idx"][o])
    assert torch.equal(d2["pred_pos"], dd["pred_pos"]) and torch.equal(d2["pred_prob_end"], dd["pred_prob_end"][o])
    lidf.compute_pair_label(d2, gt_pos)
    assert torch.equal(d2["pcl_label_float"], dd["pcl_label_float"][o])
    assert torch.equal(refine.refine_end_voxel(d2, d2["pred_pos"]), end)


# ---------------------------------------------------------------------------------------------------------------
VOXEL_CASES = ["voxel_res8_3img", "voxel_res8_unsorted", "voxel_res5"]


@pytest.mark.parametrize("name", VOXEL_CASES)
def test_voxelisation_golden_reference_outputs(name):
    """LIDFQueryMixin.get_occ_vox_bound against the reference's own get_occ_vox_bound, bit for bit."""
    from implicit_depth_b200.models.pipeline import LIDF, default_opt
    z = dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz")))
    lidf = LIDF(default_opt(**{"grid.res": int(z["res"])}), torch.device("cuda"))
    dd = dict(valid_xyz=_c(z["valid_xyz"]), valid_bid=_c(z["valid_bid"], torch.int64), bs=int(z["B"]), item_path=["golden"])
    assert lidf.get_occ_vox_bound(dd) is True
    assert dd["part_size"] == float(z["part_size"]) and np.array_equal(bits(dd["xmin"].cpu().numpy()), bits(z["xmin"]))
    assert np.array_equal(dd["occ_vox_bid"].cpu().numpy(), z["ref_occ_vox_bid"])
    assert np.array_equal(dd["occ_vox_global_coord"].cpu().numpy(), z["ref_occ_vox_global_coord"])
    assert np.array_equal(bits(dd["voxel_bound"].cpu().numpy()), bits(z["ref_voxel_bound"]))
    assert np.array_equal(dd["revidx"].cpu().numpy(), z["ref_revidx"])
    assert np.array_equal(dd["valid_v_pid"].cpu().numpy(), z["ref_valid_v_pid"])
    assert np.array_equal(bits(dd["valid_v_rel_coord"].cpu().numpy()), bits(z["ref_valid_v_rel_coord"]))
    assert dd["occ_vox_bid"].dtype == torch.int64 and dd["revidx"].dtype == torch.int64


def test_voxelisation_vs_oracle_seeded_and_edge_cases():
    from implicit_depth_b200.utils import point_utils as PU
    rng = np.random.default_rng(8)
    for n, B, res in [(1, 1, 8), (257, 2, 8), (100000, 8, 8), (5000, 3, 3), (4096, 1, 13)]:
        xyz = rng.uniform(-1.5, 2.5, size=(n, 3)).astype(np.float32)
        xyz[rng.random((n, 3)) < 0.1] = np.float32(0.125)                       # on cell faces
        bid = np.sort(rng.integers(0, B, size=n))
        want = A.get_occ_vox_bound(xyz, bid, res)
        xmin, part, rr = A.grid_setup(res)
        occ, bound, revidx, pid, rel = PU.voxelize(_c(xyz), _c(bid, torch.int64), torch.from_numpy(xmin), part, rr.tolist(), B)
        assert np.array_equal(occ[:, 0].cpu().numpy(), want["occ_vox_bid"])
        assert np.array_equal(occ[:, 1:].cpu().numpy(), want["occ_vox_global_coord"])
        assert np.array_equal(bits(bound.cpu().numpy()), bits(want["voxel_bound"]))
        assert np.array_equal(revidx.cpu().numpy(), want["revidx"]) and np.array_equal(pid.cpu().numpy(), want["valid_v_pid"])
        assert np.array_equal(bits(rel.cpu().numpy()), bits(want["valid_v_rel_coord"]))
    # reference signature: batch_id [N,1], python-tuple bounds; no point inside the grid -> empty outputs
    far = torch.full((10, 3), 50.0, device="cuda")
    occ, revidx, pid, rel, idx_grid = PU.batch_get_occupied_idx(far, torch.zeros(10, 1, dtype=torch.int64, device="cuda"),
                                                                xmin=(-1.125, -1.125, -0.125), xmax=(1.125, 1.125, 2.125),
                                                                crop_size=0.25)
    assert occ.shape == (0, 4) and revidx.numel() == 0 and pid.numel() == 0 and tuple(idx_grid.shape) == (9, 9, 9, 3)
    occ, revidx, pid, rel, _ = PU.batch_get_occupied_idx(torch.zeros(0, 3, device="cuda"), torch.zeros(0, 1, dtype=torch.int64, device="cuda"),
                                                         xmin=(-1.125, -1.125, -0.125), xmax=(1.125, 1.125, 2.125), crop_size=0.25)
    assert occ.shape == (0, 4)
    with pytest.raises(NotImplementedError):
        PU.batch_get_occupied_idx(far, torch.zeros(10, 1, dtype=torch.int64, device="cuda"), overlap=True)
    with pytest.raises(RuntimeError, match="must be a CUDA tensor"):
        PU.batch_get_occupied_idx(far.cpu(), torch.zeros(10, 1, dtype=torch.int64))


def test_front_to_back_chain_from_points():
    """The whole native front half feeding the query path, on a synthetic scene: valid points -> get_occ_vox_bound
    (bitmap voxelisation) -> compute_ray_aabb (pair list) -> get_pred (fused decoders + ray termination) ->
    compute_pair_label / compute_ray_loss / refine_end_voxel / refine_decoder_tail -- every stage against its oracle."""
    from conftest import rel_err
    from implicit_depth_b200.models.pipeline import LIDF, RefineNet, default_opt
    from implicit_depth_b200.synthetic import make_rays
    from oracle import lidf_oracle as O
    B, H, W = 2, 36, 48
    g = torch.Generator().manual_seed(101)
    # scene: a tilted plane around z = 1.1 + noise, sampled at 900 valid points per image
    n_pts = 900
    xy = torch.rand(B * n_pts, 2, generator=g) * 1.6 - 0.8
    zz = 1.1 + 0.3 * xy[:, :1] + 0.05 * torch.randn(B * n_pts, 1, generator=g)
    valid_xyz = torch.cat((xy, zz), 1).contiguous()
    valid_bid = torch.arange(B).repeat_interleave(n_pts)
    miss_bid, miss_img_ind, miss_ray_dir = make_rays(B, H, W, "cpu")
    full_rgb_feat = torch.randn(B, 32, H, W, generator=g)

    lidf = LIDF(default_opt(), torch.device("cuda")).cuda().eval()
    off = O.init_decoder("IEF", 385, mode="trained", generator=g); prob = O.init_decoder("IMNET", 385, mode="trained", generator=g)
    lidf.offset_dec.load_state_dict(off); lidf.prob_dec.load_state_dict(prob)
    dd = dict(bs=B, h=H, w=W, valid_xyz=valid_xyz.cuda(), valid_bid=valid_bid.cuda(), miss_bid=miss_bid.cuda(),
              miss_img_ind=miss_img_ind.cuda(), miss_ray_dir=miss_ray_dir.cuda(), full_rgb_feat=full_rgb_feat.cuda(),
              total_miss_sample_num=miss_bid.shape[0], item_path=["scene"])
    with torch.no_grad():
        # 1. voxelisation
        assert lidf.get_occ_vox_bound(dd)
        wv = A.get_occ_vox_bound(valid_xyz.numpy(), valid_bid.numpy(), 8)
        assert np.array_equal(bits(dd["voxel_bound"].cpu().numpy()), bits(wv["voxel_bound"]))
        assert np.array_equal(dd["revidx"].cpu().numpy(), wv["revidx"])
        V = dd["voxel_bound"].shape[0]
        assert V > 20
        # 2. pairs
        assert lidf.compute_ray_aabb(dd)
        vox, ray, pd = A.ray_aabb_pairs(miss_ray_dir.numpy(), wv["voxel_bound"], miss_bid.numpy(), wv["occ_vox_bid"])
        assert np.array_equal(dd["occ_vox_intersect_idx"].cpu().numpy(), vox) and np.array_equal(dd["miss_ray_intersect_idx"].cpu().numpy(), ray)
        assert np.array_equal(bits(dd["intersect_dist"].cpu().numpy()), bits(pd))
        P = vox.shape[0]
        assert P > 1000
        # 3. decoders + termination (the voxel features are a producer's output: random stand-in)
        occ_voxel_feat = torch.relu(torch.randn(V, 128, generator=g))
        dd["occ_voxel_feat"] = occ_voxel_feat.cuda()
        lidf.get_pred(dd, "test", 0)
        do = dict(full_rgb_feat=full_rgb_feat, occ_voxel_feat=occ_voxel_feat, voxel_bound=torch.from_numpy(wv["voxel_bound"]),
                  miss_ray_dir=miss_ray_dir, miss_img_ind=miss_img_ind, miss_bid=miss_bid,
                  occ_vox_intersect_idx=torch.from_numpy(vox), miss_ray_intersect_idx=torch.from_numpy(ray),
                  intersect_dist=torch.from_numpy(pd))
        want = O.lidf_query(do, dict(O.DEFAULT_CFG), off, prob, wv["part_size"], dedup_rays=True)
        for k in ("pred_offset", "pred_prob_end", "pair_pred_pos", "pred_prob_end_softmax"):
            assert rel_err(dd[k].cpu(), want[k]) < 1e-3, k
        same = dd["max_pair_id"].cpu() == want["max_pair_id"]
        assert float(same.float().mean()) > 0.99
        # 4. labels + ray-wise loss statistics, with the GPU's own outputs as inputs to the oracle
        gt_pos = (dd["pred_pos"] + 0.02 * torch.randn(dd["pred_pos"].shape, generator=g).cuda()).contiguous()
        lidf.compute_pair_label(dd, gt_pos)
        lab = A.pcl_pair_label(gt_pos.cpu().numpy(), wv["voxel_bound"], miss_bid.numpy(), wv["occ_vox_bid"], vox, ray)
        assert np.array_equal(dd["pcl_label_float"].cpu().numpy(), lab)
        st = lidf.compute_ray_loss(dd)
        ws = O.ray_loss_stats(dd["pred_prob_end"].cpu(), dd["pred_prob_end_softmax"].cpu(), torch.from_numpy(ray),
                              torch.from_numpy(lab).long(), miss_bid.shape[0], dd["pred_pos"].cpu(), gt_pos.cpu())
        assert torch.equal(st["pred_label"].cpu(), ws["pred_label"]) and torch.equal(st["gt_label"].cpu(), ws["gt_label"])
        for k in ("pos_loss", "prob_loss", "acc", "err"):
            assert abs(float(st[k]) - float(ws[k])) < 1e-5 * max(1.0, abs(float(ws[k]))), k
        # 5. stage 2
        refine = RefineNet(default_opt(), torch.device("cuda")).cuda().eval()
        rdec = O.init_decoder("IEF", 334, mode="trained", generator=g)
        refine.offset_dec.load_state_dict(rdec)
        end = refine.refine_end_voxel(dd, dd["pred_pos"])
        start = np.concatenate((vox, [0]))[dd["max_pair_id"].cpu().numpy()]
        want_end = A.pcl_end_voxel(dd["pred_pos"].cpu().numpy(), wv["voxel_bound"], miss_bid.numpy(), wv["occ_vox_bid"], start)
        assert np.array_equal(end.cpu().numpy(), want_end)
        feat2 = torch.relu(torch.randn(V, 128, generator=g))
        out2 = refine.refine_decoder_tail(dd, dd["pred_pos"], end, feat2.cuda(), dd["roi_feat_per_ray"])
        vb = torch.from_numpy(wv["voxel_bound"])[torch.from_numpy(want_end)]
        want2 = O.refine_decoder_tail(dd["pred_pos"].cpu(), miss_ray_dir, (vb[:, :3] + vb[:, 3:]) / 2, feat2[torch.from_numpy(want_end)],
                                      dd["roi_feat_per_ray"].cpu(), dict(O.REFINE_CFG), rdec)
        assert rel_err(out2.cpu(), want2) < 1e-3


@pytest.mark.parametrize("H,W,B", [(64, 64, 1), (48, 100, 3), (130, 33, 2)])
def test_ray_aabb_tile_culling_is_conservative(H, W, B):
    """Image-ordered all-pixel rays make the (voxel, 1024-ray tile) frustum cull bite (a tile spans a few image rows, most
    voxels project elsewhere): the pair list and the dense slab must still equal the oracle exactly, including voxels that
    touch z <= 0 (never culled), grazing boxes, and boxes stretched across the whole frustum."""
    from implicit_depth_b200.synthetic import make_rays
    rng = np.random.default_rng(H * W + B)
    miss_bid, _, ray_dir = make_rays(B, H, W, "cpu")
    V_img = 90
    lo = np.stack((rng.uniform(-1.2, 1.0, V_img * B), rng.uniform(-1.2, 1.0, V_img * B), rng.uniform(-0.2, 2.0, V_img * B)), 1)
    size = rng.uniform(0.02, 0.5, size=(V_img * B, 3))
    size[::17] = 3.0                                                  # a few huge boxes
    vb = np.concatenate((lo, lo + size), 1).astype(np.float32)
    # boxes whose faces pass exactly through pixel rays: snap a corner onto a ray at depth 1
    d = ray_dir.numpy()
    pick = rng.integers(0, d.shape[0], size=20)
    vb[:20, 0:3] = d[pick] / d[pick][:, 2:3]
    vb[:20, 3:6] = vb[:20, 0:3] + np.float32(0.1)
    xb = np.repeat(np.arange(B), V_img).astype(np.int32)
    rb = miss_bid.numpy().astype(np.int32)
    mask, dist = A.ray_aabb_dense(d, vb, rb, xb)
    assert mask.sum() > 1000
    _check_ray(d, vb, rb, xb, mask, dist, A.ray_aabb_pairs(d, vb, rb, xb))
    # rays not looking down +z in the middle of the list: their tile must not be culled
    d2 = d.copy(); d2[1500 % d.shape[0]] = np.array([0.3, -0.2, -0.9], np.float32); d2[7] = np.array([1.0, 0.0, 0.0], np.float32)
    mask2, dist2 = A.ray_aabb_dense(d2, vb, rb, xb)
    _check_ray(d2, vb, rb, xb, mask2, dist2, A.ray_aabb_pairs(d2, vb, rb, xb))


def test_forward_host_from_geometry_equals_pairs_on_the_host():
    """lidf_query.forward_host_from_geometry: rays + voxel boxes + features from HOST buffers, pairs generated on the device
    (ray-major, no regroup).  Must give what forward_host gives for the same geometry's pair list shipped from the host, with
    a fraction of the H2D bytes; the winner-only mode through it as well."""
    from implicit_depth_b200.extensions.lidf_query.jit import lidf_query as lq
    from implicit_depth_b200.synthetic import make_inputs
    from oracle import lidf_oracle as O
    ray_aabb, _ = _ext()
    d = make_inputs(2, 30, 40, 1, V_img=60, seed=93)
    g = torch.Generator().manual_seed(94)
    off = {k: v.cuda() for k, v in O.init_decoder("IEF", 385, mode="trained", generator=g).items()}
    prob = {k: v.cuda() for k, v in O.init_decoder("IMNET", 385, mode="trained", generator=g).items()}
    vox, ray, dist = ray_aabb.pairs(d["miss_ray_dir"].cuda(), d["voxel_bound"].cuda(), d["miss_bid"].int().cuda(),
                                    d["occ_vox_bid"].int().cuda(), order="ray")
    host = {k: d[k].pin_memory() for k in lq.GEOMETRY_KEYS}
    hostp = dict({k: d[k].pin_memory() for k in lq.INPUT_KEYS if k in lq.GEOMETRY_KEYS},
                 occ_vox_intersect_idx=vox.cpu().pin_memory(), miss_ray_intersect_idx=ray.cpu().pin_memory(),
                 intersect_dist=dist.cpu().pin_memory(), occ_vox_bid=d["occ_vox_bid"])
    want, h2d_p, _ = lq.forward_host(hostp, off, prob, "cuda", part_size=d["part_size"], outputs=("pred_pos", "max_pair_id"))
    got, h2d_g, d2h_g = lq.forward_host_from_geometry(host, off, prob, "cuda", part_size=d["part_size"])
    assert got["n_pairs"] == vox.shape[0] > 0 and h2d_g < h2d_p
    assert torch.equal(got["pred_pos"], want["pred_pos"]) and torch.equal(got["max_pair_id"], want["max_pair_id"])
    got2, _, _ = lq.forward_host_from_geometry(host, off, prob, "cuda", part_size=d["part_size"], winner_only=True,
                                               outputs=("pred_pos", "max_pair_id", "miss_ray_intersect_idx"))
    assert torch.equal(got2["pred_pos"], want["pred_pos"]) and torch.equal(got2["miss_ray_intersect_idx"], ray.cpu())
