"""Golden vectors for the voxelisation row, produced by the REFERENCE's own code (build container only).

    python tests/golden/make_golden_voxel.py       # needs /root/reference; writes tests/golden/voxel_*.npz

Runs ``LIDF.get_occ_vox_bound`` (reference src/models/pipeline.py:162-201, which calls
``point_utils.batch_get_occupied_idx``, src/utils/point_utils.py:12-76) unmodified on seeded points: inside / outside the
9^3 grid, exactly on cell faces, images without any point, unsorted image ids.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import make_golden as MG  # noqa: E402


def points(seed, n, B, skip_image=None, sort_bid=True):
    g = torch.Generator().manual_seed(seed)
    xyz = torch.rand(n, 3, generator=g) * torch.tensor([2.8, 2.8, 2.8]) + torch.tensor([-1.4, -1.4, -0.4])
    snap = torch.rand(n, 3, generator=g) < 0.15                     # 15 % of the coordinates sit exactly on a cell face
    xyz = torch.where(snap, torch.round((xyz + 0.125) / 0.25) * 0.25 - 0.125, xyz)
    bid = torch.randint(0, B, (n,), generator=g)
    if skip_image is not None:
        bid[bid == skip_image] = (skip_image + 1) % B
    if sort_bid:
        bid, order = torch.sort(bid, stable=True)
        xyz = xyz[order]
    return xyz.contiguous(), bid


def run(name, seed, n, B, res, **kw):
    opt, lidf, _ = MG.build_reference({"grid.res": res})
    xyz, bid = points(seed, n, B, **kw)
    dd = dict(valid_xyz=xyz, valid_bid=bid, item_path=["synthetic"])
    ok = lidf.get_occ_vox_bound(dd)                                 # reference code, unmodified
    assert ok
    out = os.path.join(HERE, name + ".npz")
    np.savez_compressed(out, valid_xyz=xyz.numpy(), valid_bid=bid.numpy().astype(np.int32), B=B, res=res,
                        part_size=np.float64(dd["part_size"]), xmin=dd["xmin"].numpy(),
                        ref_occ_vox_bid=dd["occ_vox_bid"].numpy().astype(np.int32),
                        ref_occ_vox_global_coord=dd["occ_vox_global_coord"].numpy().astype(np.int32),
                        ref_voxel_bound=dd["voxel_bound"].numpy(), ref_revidx=dd["revidx"].numpy().astype(np.int32),
                        ref_valid_v_pid=dd["valid_v_pid"].numpy().astype(np.int32),
                        ref_valid_v_rel_coord=dd["valid_v_rel_coord"].numpy())
    print(name, "points", n, "inside", dd["valid_v_pid"].shape[0], "voxels", dd["occ_vox_bid"].shape[0], "part", dd["part_size"])


if __name__ == "__main__":
    run("voxel_res8_3img", 21, 4000, 3, 8, skip_image=1)
    run("voxel_res8_unsorted", 22, 1500, 2, 8, sort_bid=False)
    run("voxel_res5", 23, 2500, 2, 5)                               # part_size 0.4: not a power of two
