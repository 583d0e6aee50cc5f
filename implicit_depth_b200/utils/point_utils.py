"""Host-side mirror of the one hot-path-adjacent function of the reference's ``utils.point_utils``:
``batch_get_occupied_idx`` (reference src/utils/point_utils.py:12-76), the voxelisation of the valid points that produces
``occ_vox_bid`` / ``voxel_bound`` / ``revidx`` for the query path.  Same name, arguments and return tuple; the work is
done by the occupancy-bitmap kernels behind ``lidf_voxelize_*`` (include/lidf_aabb.h) instead of ``torch.unique`` over
(image, cell) rows.  CUDA tensors only, ``overlap=False`` only (the only mode LIDF uses, pipeline.py:178), no fallback.
"""
from __future__ import annotations

import ctypes as C

import torch

from implicit_depth_b200.extensions.lidf_query.jit import _chk, lidf_query


def _grid_dims(xmin: torch.Tensor, xmax: torch.Tensor, crop_size: float):
    r = torch.ceil((xmax - xmin) / crop_size)                         # point_utils.py:25
    return [int(x) for x in r.long().tolist()]


def voxelize(v: torch.Tensor, batch_id: torch.Tensor, xmin: torch.Tensor, crop_size: float, dims, n_images: int):
    """-> (occ_bid_global_coord [V,4] i64, voxel_bound [V,6] f32, revidx [Nv] i64, valid_v_pid [Nv] i64,
    valid_v_rel_coord [Nv,3] f32).  Synchronises once to learn V and Nv (the reference syncs in torch.unique)."""
    lib = lidf_query.lib
    dev = v.device
    Np = int(v.shape[0])
    bid = batch_id.reshape(-1)
    x = [float(t) for t in xmin.detach().cpu().tolist()]
    args = (_chk(v, "v", torch.float32), _chk(bid, "batch_id", torch.int64), Np, int(n_images), x[0], x[1], x[2],
            float(crop_size), float(0.5 * crop_size), int(dims[0]), int(dims[1]), int(dims[2]))
    nbytes = int(lib.lidf_voxelize_workspace_bytes(Np, int(n_images), int(dims[0]), int(dims[1]), int(dims[2])))
    if nbytes == 0:
        raise RuntimeError("voxelize: invalid grid / problem too large")
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    nv, ni = C.c_int64(0), C.c_int64(0)
    with torch.cuda.device(dev):
        st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        rc = lib.lidf_voxelize_count(*args, ws.data_ptr(), nbytes, C.byref(nv), C.byref(ni), st)
        lidf_query._raise(rc, "lidf_voxelize_count")
        V, Nv = int(nv.value), int(ni.value)
        occ = torch.empty(V, 4, dtype=torch.int64, device=dev)
        bound = torch.empty(V, 6, dtype=torch.float32, device=dev)
        revidx = torch.empty(Nv, dtype=torch.int64, device=dev)
        pid = torch.empty(Nv, dtype=torch.int64, device=dev)
        rel = torch.empty(Nv, 3, dtype=torch.float32, device=dev)
        if V > 0:                                                   # V == 0 implies Nv == 0: nothing to write
            rc = lib.lidf_voxelize_fill(*args, ws.data_ptr(), nbytes, occ.data_ptr(), bound.data_ptr(), revidx.data_ptr(),
                                        pid.data_ptr(), rel.data_ptr(), st)
            lidf_query._raise(rc, "lidf_voxelize_fill")
    ws.record_stream(torch.cuda.current_stream(dev))
    return occ, bound, revidx, pid, rel


def batch_get_occupied_idx(v, batch_id, xmin=(0., 0., 0.), xmax=(1., 1., 1.), crop_size=.125, overlap=False, n_images=None):
    """Reference signature (point_utils.py:12-15) and return tuple
    ``(occ_bid_global_coord, revidx, valid_v_pid, valid_v_rel_coord, idx_grid)``.  ``n_images`` (extra, optional) avoids a
    device->host read of ``batch_id.max()``."""
    if overlap:
        raise NotImplementedError("overlap=True is not used by LIDF (pipeline.py:178) and not implemented")
    if not torch.is_tensor(xmin):
        xmin = torch.Tensor(xmin).float().to(v.device)
    if not torch.is_tensor(xmax):
        xmax = torch.Tensor(xmax).float().to(v.device)
    dims = _grid_dims(xmin, xmax, crop_size)
    if n_images is None:
        n_images = int(batch_id.max().item()) + 1 if batch_id.numel() else 1
    occ, _, revidx, pid, rel = voxelize(v.float().contiguous(), batch_id.long().contiguous(), xmin, crop_size, dims, n_images)
    idx_grid = torch.stack(torch.meshgrid(torch.arange(dims[0]), torch.arange(dims[1]), torch.arange(dims[2]), indexing="ij"),
                           dim=-1).to(v.device)                   # only its shape is read downstream
    return occ, revidx, pid, rel, idx_grid
