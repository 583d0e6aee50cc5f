"""Debug helper (GPU): per-tensor errors of the native backward vs goldens / oracle autograd, wgrad kernel errors."""
import os, sys
import numpy as np
import torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "tests"))
from conftest import GOLDEN_DIR, load_golden, rel_err
import test_gpu_backward as T
from implicit_depth_b200.extensions.lidf_query.jit import lidf_query

dev = torch.device("cuda", 0)
for rows, M, N in [(64, 128, 64), (128, 128, 64), (1000, 128, 256), (4097, 256, 112), (333, 256, 32), (70000, 256, 128), (640, 128, 64)]:
    g = torch.Generator().manual_seed(rows + M + N)
    A = torch.randn(rows, M, generator=g).to(dev); B = torch.randn(rows, N, generator=g).to(dev)
    C = lidf_query.wgrad_selftest(A, B)
    want = A.double().t() @ B.double()
    print("wgrad", rows, M, N, "rel_err %.3e" % rel_err(C.cpu(), want.cpu()), "max abs err %.3e" % float((C.double() - want).abs().max()),
          "rms", float(want.pow(2).mean().sqrt()), flush=True)

def report(res, want, tag):
    print("==", tag)
    print("  full_rgb_feat  %.3e" % rel_err(res["full_rgb_feat"].cpu(), want["full_rgb_feat"]))
    print("  occ_voxel_feat %.3e" % rel_err(res["occ_voxel_feat"].cpu(), want["occ_voxel_feat"]))
    for mod in ("offset_dec", "prob_dec"):
        for k, w in want[mod].items():
            if w is None:
                print("  %s.%s want None, got absmax %.3e" % (mod, k, float(res[mod][k].abs().max()))); continue
            r = res[mod][k].cpu()
            if k == "linear_1.weight":
                D = w.shape[1]
                for name, sl in (("vox", slice(0, 128)), ("rgb", slice(128, 256)), ("pos", slice(256, 358)), ("dir", slice(358, 385)), ("enc", slice(385, D))):
                    if sl.start < D:
                        print("  %s.%s[%s] %.3e" % (mod, k, name, rel_err(r[:, sl], w[:, sl])))
            else:
                print("  %s.%s %.3e" % (mod, k, rel_err(r, w)))

for name, gname in [("ief_ragged_2x24x32", "grad_ief_ragged_2x24x32"), ("ief_rel_sigmoid_1x16x20", "grad_ief_rel_sigmoid_1x16x20")]:
    d, cfg, off, prob, part, ref, _ = load_golden(name)
    z = np.load(os.path.join(GOLDEN_DIR, gname + ".npz"))
    coef = dict(pred_pos=torch.from_numpy(z["c_pos"]), pred_prob_end=torch.from_numpy(z["c_prob"]))
    out, res = T._native_grads(d, cfg, off, prob, part, coef)
    print(name, "max_pair_id equal:", bool(torch.equal(out["max_pair_id"].cpu(), torch.from_numpy(z["max_pair_id"]).long())))
    want = dict(full_rgb_feat=torch.from_numpy(z["grad.full_rgb_feat"]), occ_voxel_feat=torch.from_numpy(z["grad.occ_voxel_feat"]),
                offset_dec={k: torch.from_numpy(z[f"grad.offset_dec.{k}"]) for k in off},
                prob_dec={k: torch.from_numpy(z[f"grad.prob_dec.{k}"]) for k in prob})
    report(res, want, name + " vs reference autograd golden")
    want2 = T._oracle_grads(d, cfg, off, prob, part, coef, out["max_pair_id"].cpu())
    report(res, want2, name + " vs oracle fp64 autograd")

for offdec, n_iter, rel, sig, chunk in [("IMNET", 1, False, False, 0), ("IEF", 2, False, False, 0), ("IEF", 2, False, False, 1024), ("IEF", 3, True, True, 2048)]:
    d, cfg, off, prob, part, g = T._seeded_case(2, 20, 28, 7, 24, seed=300 + n_iter, offdec=offdec, n_iter=n_iter, rel=rel, sigmoid=sig)
    P, R = d["occ_vox_intersect_idx"].shape[0], d["miss_ray_dir"].shape[0]
    coef = dict(pred_pos=torch.randn(R, 3, generator=g), pred_prob_end=torch.randn(P, 1, generator=g),
                pred_offset=0.3 * torch.randn(P, 1, generator=g), pair_pred_pos=0.2 * torch.randn(P, 3, generator=g))
    out, res = T._native_grads(d, cfg, off, prob, part, coef, chunk_rows=chunk)
    mp = out["max_pair_id"].cpu()
    margin = T.kink_margin(d, cfg, off, prob)
    for thr in (0.0, 3e-5, 3e-4):
        cm, kept = T.mask_coefficients(coef, margin, d["miss_ray_intersect_idx"], mp, thr)
        out, res = T._native_grads(d, cfg, off, prob, part, cm, chunk_rows=chunk)
        want = T._oracle_grads(d, cfg, off, prob, part, cm, mp)
        report(res, want, f"seeded {offdec} n_iter {n_iter} rel {rel} sig {sig} chunk {chunk} P {P} kink thr {thr} kept {kept:.3f}")
