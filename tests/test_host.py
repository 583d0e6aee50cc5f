"""CPU tests: host logic, module surface, the C-ABI library loads and exports every declared symbol."""
import ctypes
import os
import re

import pytest
import torch

from conftest import REPO, load_golden, rel_err
from oracle import lidf_oracle as O


def test_library_builds_loads_and_exports_every_declared_symbol():
    from implicit_depth_b200 import build
    from implicit_depth_b200.extensions.lidf_query import jit
    path = build.build()
    assert os.path.exists(path)
    lib = jit.load_library()
    declared = set()
    for name, want in (("lidf_query.h", jit.EXPORTED_SYMBOLS), ("lidf_aabb.h", jit.EXPORTED_SYMBOLS_AABB),
                       ("lidf_pointnet.h", jit.EXPORTED_SYMBOLS_POINTNET)):
        header = open(os.path.join(REPO, "include", name)).read()
        header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
        got = set(re.findall(r"\b(lidf_[a-z_0-9]+)\s*\(", header))
        assert got == set(want), (name, got ^ set(want))
        declared |= got
    assert sorted(os.listdir(os.path.join(REPO, "include"))) == ["lidf_aabb.h", "lidf_pointnet.h", "lidf_query.h"]
    for sym in declared:
        assert hasattr(lib, sym), sym
    assert lib.lidf_query_abi_version() == 4
    assert lib.lidf_query_backward(None, None) == -1 and lib.lidf_query_backward_workspace_bytes(None) == 0
    assert lib.lidf_query_error_string(-2).decode().startswith("unsupported")
    # argument validation needs no GPU: NULL params / empty problem
    assert lib.lidf_query_forward(None, None) == -1
    assert lib.lidf_refine_forward(None, None) == -1
    assert lib.lidf_query_workspace_bytes(None) == 0
    assert lib.lidf_ray_terminate_workspace_bytes(1000, 10) > 4000
    assert lib.lidf_ray_aabb_workspace_bytes(5000, 40) > 5000 * 12 and lib.lidf_ray_aabb_workspace_bytes(-1, 4) == 0
    assert lib.lidf_ray_aabb_forward(None, None, None, None, 5, 5, None, None, None, 0, None) == -1
    assert lib.lidf_pcl_aabb_forward(None, None, None, None, 5, 5, None, None) == -1
    assert lib.lidf_pcl_aabb_end_voxel(None, None, None, None, 0, 5, None, None) == 0       # empty problem: no-op
    lib.lidf_pointnet_workspace_bytes.restype = ctypes.c_size_t
    lib.lidf_pointnet_workspace_bytes.argtypes = [ctypes.c_int64, ctypes.c_int64]
    assert lib.lidf_pointnet_workspace_bytes(1000, 10) >= 10 * 256 * 4 and lib.lidf_pointnet_forward(None, None, None, 0, 0, None, None, 0, None) == -1
    assert lib.lidf_voxelize_workspace_bytes(10000, 8, 9, 9, 9) > 8 * 729 * 8 and lib.lidf_voxelize_workspace_bytes(5, 0, 9, 9, 9) == 0


def test_packed_handover_layout_is_the_canonical_mn_major_umma_layout():
    """The backward hands activations to the wgrad kernel in 64-row groups laid out as `[hi: F/8 feature groups][64 rows][16 B]
    [lo: same]` (csrc/lidf_bwd.cuh: bw_pk_*).  k_wgrad_pk_tc bulk-copies a group and describes it to tcgen05.mma as a Major-MN,
    no-swizzle operand with LBO = 128 B and SBO = 1 KB; CUTLASS documents that canonical layout, in 16-byte units, as
    ((1,n),(8,k)) : ((X,SBO),(1,LBO)) (cute/atom/mma_traits_sm100.hpp, make_umma_desc<Major::MN>).  Restated here and held
    against the library's own offset arithmetic, plus the properties the kernels rely on."""
    from implicit_depth_b200.extensions.lidf_query import jit
    lib = jit.load_library()
    LBO, SBO, ROWS = 128, 1024, 64

    def canonical(m, k):               # byte offset of element (feature m, row k) inside one group's hi (or lo) part
        return (m // 8) * SBO + (k % 8) * 16 + (k // 8) * LBO + (m % 8) * 2
    for F in (64, 128, 256):
        group, lo_off = ROWS * F * 4, (F // 8) * SBO
        for row in (0, 1, 7, 8, 63, 64, 65, 127, 1000):
            for m in (0, 1, 7, 8, F // 2 + 3, F - 1):
                for lo in (0, 1):
                    want = (row // ROWS) * group + lo * lo_off + canonical(m, row % ROWS)
                    assert lib.lidf_pk_offset_bytes(F, row, m, lo) == want, (F, row, m, lo)
        # a warp of k_mlp_bwd_tc (32 consecutive rows, one 16-byte unit each) writes 512 contiguous bytes
        offs = [lib.lidf_pk_offset_bytes(F, 32 + r, 8, 0) for r in range(32)]
        assert offs == list(range(offs[0], offs[0] + 512, 16))
        # hi and lo parts tile the group exactly: no byte is shared, none is left over
        seen = {lib.lidf_pk_offset_bytes(F, r, m, lo) for r in range(ROWS) for m in range(0, F, 8) for lo in (0, 1)}
        assert len(seen) == 2 * ROWS * F // 8 and min(seen) == 0 and max(seen) == group - 16
    # one pipeline stage of each product the backward runs (A features + B features) leaves room for >= 2 stages in 227 KB
    for fa, fb in ((128, 64), (128, 256), (128, 256)):
        assert 2 * ROWS * (fa + fb) * 4 + 1024 <= 227 * 1024
    assert lib.lidf_pk_offset_bytes(100, 0, 0, 0) == -1 and lib.lidf_pk_offset_bytes(128, 0, 128, 0) == -1


def test_no_cpu_fallback_for_aabb_ops():
    from implicit_depth_b200.extensions.pcl_aabb.jit import pcl_aabb
    from implicit_depth_b200.extensions.ray_aabb.jit import ray_aabb
    f, i = torch.zeros(4, 3), torch.zeros(4, dtype=torch.int32)
    with pytest.raises(RuntimeError, match="must be a CUDA tensor"):
        ray_aabb.forward(f, torch.zeros(2, 6), i, i[:2])
    with pytest.raises(RuntimeError, match="must be a CUDA tensor"):
        pcl_aabb.forward(f, torch.zeros(2, 6), i, i[:2])


def test_no_cpu_fallback_for_fused_op():
    from implicit_depth_b200.extensions.lidf_query.jit import lidf_query
    d, cfg, off, prob, part, ref, _ = load_golden("ief_rel_sigmoid_1x16x20")
    with pytest.raises(RuntimeError, match="must be a CUDA tensor"):
        lidf_query.forward(d["full_rgb_feat"], d["occ_voxel_feat"], d["miss_ray_dir"], d["miss_img_ind"], d["miss_bid"],
                           d["voxel_bound"], d["occ_vox_intersect_idx"], d["miss_ray_intersect_idx"], d["intersect_dist"],
                           off, prob, part_size=part)


def test_product_code_does_not_import_oracle():
    for root, _, files in os.walk(os.path.join(REPO, "implicit_depth_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(root, f)).read()
                assert "oracle" not in src.replace("# oracle-free", ""), os.path.join(root, f)


def test_implicit_net_mirror_matches_reference_goldens():
    """Module surface: same state_dict keys as the reference checkpoints, same forward values."""
    from implicit_depth_b200.models import implicit_net as N
    d, cfg, off, prob, part, ref, _ = load_golden("ief_ragged_2x24x32")
    ief = N.IEF(torch.device("cpu"), 385, 1, gf_dim=64, n_iter=2)
    imn = N.IMNet(385, 1, gf_dim=64)
    assert set(ief.state_dict().keys()) == set(off.keys())
    assert set(imn.state_dict().keys()) == set(prob.keys())
    assert sum(p.numel() for p in ief.parameters()) == 144161 and sum(p.numel() for p in imn.parameters()) == 140033
    ief.load_state_dict(off); imn.load_state_dict(prob)
    e = O.get_embedding(d, cfg, dedup_rays=True)
    x = torch.cat((e["intersect_voxel_feat"], e["intersect_rgb_feat"], e["intersect_enter_pos_embed"],
                   e["intersect_leave_pos_embed"], e["intersect_dir_embed"]), -1)
    with torch.no_grad():
        assert rel_err(ief(x), ref["pred_offset"]) < 2e-5
        assert rel_err(imn(x), ref["pred_prob_end"]) < 2e-5
    fn, dim = N.get_embedder(8)
    assert dim == 51 and torch.equal(fn(x[:4, :3]), O.embed(x[:4, :3], 8))
    ident, dim = N.get_embedder(8, i=-1)
    assert dim == 3 and isinstance(ident, torch.nn.Identity)


def test_training_path_has_no_cpu_fallback():
    """While autograd records, get_pred is one autograd node over the sm_100a kernels (lidf_query_forward / _backward):
    on CPU tensors it must fail loudly instead of falling back to torch ops."""
    from implicit_depth_b200.models import pipeline as PL
    d, cfg, off, prob, part, ref, _ = load_golden("ief_ragged_2x24x32")
    lidf = PL.LIDF(PL.default_opt(), torch.device("cpu"))
    lidf.offset_dec.load_state_dict(off); lidf.prob_dec.load_state_dict(prob)
    dd = dict(d); dd.update(total_miss_sample_num=d["miss_ray_dir"].shape[0], part_size=part)
    lidf.train()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        lidf.get_pred(dd, "test", 0)
    assert not hasattr(PL.LIDFQueryMixin, "_get_pred_autograd") and not hasattr(PL, "_ray_terminate_torch")


def test_synthetic_generator_contract():
    from implicit_depth_b200.synthetic import make_inputs, shard_images
    d = make_inputs(2, 12, 16, 5, V_img=32, seed=1)
    R, P = 2 * 12 * 16, 2 * 12 * 16 * 5
    assert d["miss_ray_dir"].shape == (R, 3) and d["occ_vox_intersect_idx"].shape == (P,)
    assert torch.allclose(d["miss_ray_dir"].norm(dim=-1), torch.ones(R), atol=1e-6)
    key = d["occ_vox_intersect_idx"] * R + d["miss_ray_intersect_idx"]
    assert (key[1:] > key[:-1]).all()                        # voxel-major, strictly increasing => distinct pairs
    assert torch.equal(d["occ_vox_bid"][d["occ_vox_intersect_idx"]], d["miss_bid"][d["miss_ray_intersect_idx"]])
    assert (d["intersect_dist"][:, 1] > d["intersect_dist"][:, 0]).all()
    vb = d["voxel_bound"]
    assert torch.allclose(vb[:, 3:] - vb[:, :3], torch.full((64, 3), 0.25))
    assert [shard_images(8, r, 3) for r in range(3)] == [(0, 3), (3, 3), (6, 2)]


def test_image_splits_of_the_reference_layout():
    """Per-image slices of the ray / voxel / pair arrays used by the pipelined host path (rays and voxels image-major,
    pair list voxel-major -- reference pipeline.py:226-262, :283-285)."""
    from implicit_depth_b200.extensions.lidf_query.jit import lidf_query
    from implicit_depth_b200.synthetic import make_inputs
    d = make_inputs(3, 10, 12, 4, V_img=16, seed=3, ragged=True)
    s = lidf_query.image_splits(d, 3)
    assert s["rays"] == [0, 120, 240, 360] and s["voxels"] == [0, 16, 32, 48]
    vox, ray = d["occ_vox_intersect_idx"], d["miss_ray_intersect_idx"]
    for b in range(3):
        p0, p1 = s["pairs"][b], s["pairs"][b + 1]
        assert bool(((vox[p0:p1] >= s["voxels"][b]) & (vox[p0:p1] < s["voxels"][b + 1])).all())
        assert bool(((ray[p0:p1] >= s["rays"][b]) & (ray[p0:p1] < s["rays"][b + 1])).all())
    assert s["pairs"][-1] == vox.shape[0]
    d.pop("occ_vox_bid")
    assert lidf_query.image_splits(d, 3) is None


def test_pointnet_mirror_state_dict_and_autograd_path():
    """models.pointnet.PointNet2Stage mirror: reference checkpoint keys, reference values on the torch (training) path,
    and no CPU fallback for the native (inference) path."""
    from test_oracle import load_pointnet
    from implicit_depth_b200.models.pointnet import PointNet2Stage
    w, inp, idx, V, ref = load_pointnet("pointnet_2000x60")
    net = PointNet2Stage(input_channels=6, output_channels=128, gf_dim=32)
    assert set(net.state_dict().keys()) == set(w.keys())
    net.load_state_dict(w)
    import numpy as np
    from conftest import GOLDEN_DIR
    z = np.load(os.path.join(GOLDEN_DIR, "pointnet_2000x60.npz"))
    inp_g = inp.clone().requires_grad_(True)
    out = net(inp_g, idx)                                    # parameters require grad -> torch path
    assert rel_err(out.detach(), ref) < 2e-6
    (torch.from_numpy(z["c_out"]) * out).sum().backward()    # gradients of the reference module, make_golden_pointnet.py
    assert rel_err(inp_g.grad, torch.from_numpy(z["grad.inp_feat"])) < 1e-5
    for k, p_ in net.named_parameters():
        assert rel_err(p_.grad, torch.from_numpy(z["grad." + k])) < 1e-5, k
    with torch.no_grad(), pytest.raises(RuntimeError, match="must be a CUDA tensor"):
        net(inp, idx)


def test_bench_reference_arm_runs_without_a_gpu_and_prints_the_contract_line():
    """`bench.py --impl reference` is pure CPU (oracle port): one JSON line with the driver's keys; other ranks stay silent."""
    import json
    import subprocess
    import sys
    cmd = [sys.executable, os.path.join(REPO, "bench.py"), "--impl", "reference", "--workload", "tiny", "--steps", "1",
           "--warmup", "0", "--cpu-sample-pairs", "4096"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=dict(os.environ, RANK="0"))
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "lidf_query_points_per_sec" and line["unit"] == "points/s"
    assert line["value"] > 0 and line["higher_is_better"] is True and line["gpu_launches"] == 0
    # "reference" = the unmodified reference classes driven on the host cores (tree present: /root/reference/src here,
    # baseline/_ref/src on the GPU box), "port" = the oracle's restatement of the same op chain (tree absent)
    from oracle import ref_loader
    want_kind = "reference" if ref_loader.find_ref_src() is not None else "port"
    assert line["cpu_baseline"]["kind"] == want_kind and line["cpu_baseline"]["cores"] >= 1 and "sample" in line["cpu_baseline"]
    assert line["e2e"] == dict(value=line["value"], unit="points/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0)
    quiet = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=dict(os.environ, RANK="1", WORLD_SIZE="2"))
    assert quiet.returncode == 0 and quiet.stdout.strip() == ""


def test_refine_training_path_cpu_matches_reference_output_and_has_gradients():
    """RefineDecoderMixin.refine_decoder_tail while autograd records (train_refine.yaml trains pnet_model + offset_dec):
    torch ops, same values as the reference's get_pred_refine, gradients reach the decoder and the voxel features."""
    from implicit_depth_b200.models.pipeline import RefineNet, default_opt
    d, cfg, off, prob, part, ref, extra = load_golden("ief_ragged_2x24x32")
    rdec = {k[len("refine_dec."):]: v for k, v in extra.items() if k.startswith("refine_dec.")}
    refine = RefineNet(default_opt(), torch.device("cpu"))
    refine.offset_dec.load_state_dict(rdec)
    refine.train()
    vfeat = extra["refine.occ_voxel_feat"].clone().requires_grad_(True)
    roi = O.get_embedding(d, cfg, dedup_rays=True)["intersect_rgb_feat"]
    R = d["miss_ray_dir"].shape[0]
    roi_per_ray = torch.zeros(R, 128); roi_per_ray[d["miss_ray_intersect_idx"]] = roi
    roi_per_ray = torch.where(torch.isnan(ref["roi_feat_per_ray"]), roi_per_ray, ref["roi_feat_per_ray"])   # rays without pairs
    # rays without a pair have no ROI row in the fixture; compute them with the oracle's roi_align
    boxes_missing = torch.isnan(ref["roi_feat_per_ray"]).any(1)
    if boxes_missing.any():
        pix = d["miss_img_ind"].float(); H, W = d["full_rgb_feat"].shape[2:]
        ul = torch.stack(((pix[:, 0] - 4).clamp(0, W - 1), (pix[:, 1] - 4).clamp(0, H - 1)), -1)
        br = torch.stack(((pix[:, 0] + 4).clamp(0, W - 1), (pix[:, 1] + 4).clamp(0, H - 1)), -1)
        boxes = torch.cat((d["miss_bid"].float().unsqueeze(-1), ul, br), -1)
        full = O.roi_align_aligned(d["full_rgb_feat"], boxes, 2, 1.0).reshape(R, -1)
        roi_per_ray = torch.where(boxes_missing.unsqueeze(1), full, roi_per_ray)
    out = refine.refine_decoder_tail(dict(d), ref["pred_pos"], extra["refine.end_voxel_id"].long(), vfeat, roi_per_ray)
    assert rel_err(out.detach(), ref["pred_pos_refine"]) < 2e-5
    out.abs().sum().backward()
    assert vfeat.grad is not None and float(vfeat.grad.abs().sum()) > 0
    assert refine.offset_dec.linear_1.weight.grad is not None


def test_integration_doc_names_every_exported_symbol():
    from implicit_depth_b200.extensions.lidf_query import jit
    doc = open(os.path.join(REPO, "INTEGRATION.md")).read()
    for sym in jit.EXPORTED_SYMBOLS + jit.EXPORTED_SYMBOLS_AABB + jit.EXPORTED_SYMBOLS_POINTNET:
        assert sym in doc, f"INTEGRATION.md does not mention {sym}"


def test_forward_rejects_inconsistent_shapes_before_touching_the_gpu():
    from implicit_depth_b200.extensions.lidf_query.jit import lidf_query
    d, cfg, off, prob, part, ref, _ = load_golden("ief_rel_sigmoid_1x16x20")
    args = [d[k] for k in lidf_query.INPUT_KEYS]
    bad = list(args); bad[3] = d["miss_img_ind"][:-1]                       # one ray short
    with pytest.raises(RuntimeError, match="miss_img_ind must have shape"):
        lidf_query.forward(*bad, off, prob, part_size=part)
    bad = list(args); bad[5] = d["voxel_bound"][:, :5].contiguous()
    with pytest.raises(RuntimeError, match="voxel_bound must have shape"):
        lidf_query.forward(*bad, off, prob, part_size=part)
    bad = list(args); bad[7] = d["miss_ray_intersect_idx"][:-2]
    with pytest.raises(RuntimeError, match="miss_ray_intersect_idx must have shape"):
        lidf_query.forward(*bad, off, prob, part_size=part)
