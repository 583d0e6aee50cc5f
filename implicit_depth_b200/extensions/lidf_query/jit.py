"""``lidf_query`` extension loader -- same convention as the reference's native extensions.

The reference loads its CUDA ops at import time as ``from extensions.ray_aabb.jit import ray_aabb`` and calls
``ray_aabb.forward(...)`` (reference src/extensions/ray_aabb/jit.py:1-3, src/models/pipeline.py:18,277).
This module does the same for the fused query op::

    from extensions.lidf_query.jit import lidf_query
    outs = lidf_query.forward(full_rgb_feat, occ_voxel_feat, ...)

It binds the C ABI of ``include/lidf_query.h`` with ctypes (no torch types cross the boundary: device pointers,
sizes and the current CUDA stream handle).  The library is built in-tree by ``implicit_depth_b200/build.py`` (nvcc,
sm_100a).  There is NO fallback: if the library cannot be loaded, or the tensors are not CUDA tensors, the call
raises -- exactly like the reference's ``CHECK_CUDA`` (ray_aabb_cuda.cpp:16-18).
"""
from __future__ import annotations

import ctypes as C
import math
import os
from typing import Dict, Optional, Sequence

import torch

from implicit_depth_b200 import build as _build

MLP_IMPLS = {"auto": 0, "simt_fp32": 1, "tc_bf16x3": 2, "tc_bf16x1": 3}


class _Decoder(C.Structure):
    _fields_ = [("kind", C.c_int32), ("n_iter", C.c_int32), ("inp_dim", C.c_int32), ("use_sigmoid", C.c_int32),
                ("init_offset", C.c_float),
                ("w1", C.c_void_p), ("b1", C.c_void_p), ("w2", C.c_void_p), ("b2", C.c_void_p),
                ("w3", C.c_void_p), ("b3", C.c_void_p), ("w4", C.c_void_p), ("b4", C.c_void_p),
                ("w_enc", C.c_void_p), ("b_enc", C.c_void_p)]


class _QueryParams(C.Structure):
    _fields_ = [("P", C.c_int64), ("R", C.c_int64), ("V", C.c_int64),
                ("B", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
                ("full_rgb_feat", C.c_void_p), ("occ_voxel_feat", C.c_void_p),
                ("miss_ray_dir", C.c_void_p), ("miss_img_ind", C.c_void_p), ("miss_bid", C.c_void_p),
                ("voxel_bound", C.c_void_p), ("pair_vox", C.c_void_p), ("pair_ray", C.c_void_p),
                ("pair_dist", C.c_void_p), ("dense_dist", C.c_void_p), ("pcl_label_float", C.c_void_p),
                ("pos_encode", C.c_int32), ("multires", C.c_int32), ("multires_views", C.c_int32),
                ("intersect_pos_rel", C.c_int32), ("roi_inp_bbox", C.c_int32),
                ("offset_range0", C.c_float), ("offset_range1", C.c_float), ("part_size", C.c_float),
                ("offset_dec", _Decoder), ("prob_dec", _Decoder), ("mlp_impl", C.c_int32),
                ("pred_offset", C.c_void_p), ("pred_prob_end", C.c_void_p), ("pair_pred_pos", C.c_void_p),
                ("pred_prob_end_softmax", C.c_void_p), ("max_pair_id", C.c_void_p), ("pred_pos", C.c_void_p),
                ("roi_feat_per_ray", C.c_void_p), ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t),
                ("ief_iter_out", C.c_void_p), ("index_error", C.c_void_p),
                ("weight_cache", C.c_void_p), ("weight_cache_bytes", C.c_size_t), ("weight_cache_valid", C.c_int32),
                ("pairs_ray_major", C.c_int32), ("winner_only_offset", C.c_int32), ("pred_offset_ray", C.c_void_p)]


class _DecoderGrad(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in ("w1", "b1", "w2", "b2", "w3", "b3", "w4", "b4", "w_enc", "b_enc")]


class _BackwardParams(C.Structure):
    _fields_ = [("fwd", _QueryParams), ("ief_iter", C.c_void_p),
                ("g_pred_pos", C.c_void_p), ("g_pred_prob_end", C.c_void_p), ("g_pred_offset", C.c_void_p),
                ("g_pair_pred_pos", C.c_void_p), ("g_full_rgb_feat", C.c_void_p), ("g_occ_voxel_feat", C.c_void_p),
                ("g_offset_dec", _DecoderGrad), ("g_prob_dec", _DecoderGrad), ("chunk_rows", C.c_int64),
                ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t)]


class _RefineParams(C.Structure):
    _fields_ = [("R", C.c_int64), ("pred_pos", C.c_void_p), ("miss_ray_dir", C.c_void_p),
                ("end_voxel_center", C.c_void_p), ("voxel_feat_end", C.c_void_p), ("rgb_feat_end", C.c_void_p),
                ("pos_encode", C.c_int32), ("multires", C.c_int32), ("multires_views", C.c_int32),
                ("intersect_pos_rel", C.c_int32), ("offset_range0", C.c_float), ("offset_range1", C.c_float),
                ("offset_dec", _Decoder), ("mlp_impl", C.c_int32), ("pred_pos_refine", C.c_void_p),
                ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t),
                ("V", C.c_int64), ("occ_voxel_feat", C.c_void_p), ("end_voxel_id", C.c_void_p), ("voxel_bound", C.c_void_p)]


ABI_VERSION = 4
EXPORTED_SYMBOLS = [
    "lidf_query_abi_version", "lidf_query_struct_size", "lidf_query_error_string", "lidf_query_last_cuda_error",
    "lidf_query_workspace_bytes", "lidf_query_weight_cache_bytes", "lidf_query_forward", "lidf_refine_workspace_bytes", "lidf_refine_forward",
    "lidf_roi_align_rays", "lidf_ray_terminate_workspace_bytes", "lidf_ray_terminate", "lidf_query_launch_count",
    "lidf_query_last_mlp_ms", "lidf_tc_selftest", "lidf_ray_loss_workspace_bytes", "lidf_ray_loss",
    "lidf_image_loss_workspace_bytes", "lidf_image_loss",
    "lidf_query_backward_workspace_bytes", "lidf_query_backward", "lidf_query_last_bwd_ms",
    "lidf_wgrad_selftest_scratch_bytes", "lidf_wgrad_selftest", "lidf_wgrad_pk_selftest_scratch_bytes", "lidf_wgrad_pk_selftest", "lidf_pk_offset_bytes",
    "lidf_depth_metrics_workspace_bytes", "lidf_depth_metrics_rays", "lidf_depth_metrics_image",
]
# include/lidf_pointnet.h (bound by models/pointnet.py)
EXPORTED_SYMBOLS_POINTNET = ["lidf_pointnet_workspace_bytes", "lidf_pointnet_forward", "lidf_pointnet_forward_impl"]
# include/lidf_aabb.h (bound by extensions/ray_aabb/jit.py and extensions/pcl_aabb/jit.py)
EXPORTED_SYMBOLS_AABB = [
    "lidf_ray_aabb_workspace_bytes", "lidf_ray_aabb_forward", "lidf_ray_aabb_pairs_count", "lidf_ray_aabb_pairs_fill",
    "lidf_ray_aabb_ray_major_workspace_bytes", "lidf_ray_aabb_pairs_ray_major_count", "lidf_ray_aabb_pairs_ray_major_fill",
    "lidf_pcl_aabb_forward", "lidf_pcl_aabb_pair_label", "lidf_pcl_aabb_end_voxel",
    "lidf_voxelize_workspace_bytes", "lidf_voxelize_count", "lidf_voxelize_fill",
]


def load_library(build_if_needed: bool = True) -> C.CDLL:
    """dlopen csrc/liblidf_query.so (building it first when nvcc is available and it is missing/stale)."""
    path = _build.LIB_PATH
    if build_if_needed and _build.find_nvcc() is not None:
        path = _build.build()
    if not os.path.exists(path):
        raise RuntimeError(f"{path} is missing and nvcc is not available to build it; lidf_query has no fallback path")
    lib = C.CDLL(path)
    lib.lidf_query_abi_version.restype = C.c_int
    lib.lidf_query_error_string.restype = C.c_char_p
    lib.lidf_query_error_string.argtypes = [C.c_int]
    lib.lidf_query_last_cuda_error.restype = C.c_char_p
    lib.lidf_query_workspace_bytes.restype = C.c_size_t
    lib.lidf_query_workspace_bytes.argtypes = [C.POINTER(_QueryParams)]
    lib.lidf_query_weight_cache_bytes.restype = C.c_size_t
    lib.lidf_query_weight_cache_bytes.argtypes = [C.POINTER(_QueryParams)]
    lib.lidf_query_forward.restype = C.c_int
    lib.lidf_query_forward.argtypes = [C.POINTER(_QueryParams), C.c_void_p]
    lib.lidf_query_backward_workspace_bytes.restype = C.c_size_t
    lib.lidf_query_backward_workspace_bytes.argtypes = [C.POINTER(_BackwardParams)]
    lib.lidf_query_backward.restype = C.c_int
    lib.lidf_query_backward.argtypes = [C.POINTER(_BackwardParams), C.c_void_p]
    lib.lidf_query_last_bwd_ms.restype = C.c_float
    lib.lidf_wgrad_selftest_scratch_bytes.restype = C.c_size_t
    lib.lidf_wgrad_selftest_scratch_bytes.argtypes = [C.c_int32, C.c_int32]
    lib.lidf_wgrad_selftest.restype = C.c_int
    lib.lidf_wgrad_selftest.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]
    lib.lidf_wgrad_pk_selftest_scratch_bytes.restype = C.c_size_t
    lib.lidf_wgrad_pk_selftest_scratch_bytes.argtypes = [C.c_int64, C.c_int32, C.c_int32]
    lib.lidf_pk_offset_bytes.restype = C.c_int64
    lib.lidf_pk_offset_bytes.argtypes = [C.c_int32, C.c_int64, C.c_int32, C.c_int32]
    lib.lidf_wgrad_pk_selftest.restype = C.c_int
    lib.lidf_wgrad_pk_selftest.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]
    lib.lidf_depth_metrics_workspace_bytes.restype = C.c_size_t
    lib.lidf_depth_metrics_workspace_bytes.argtypes = [C.c_int64, C.c_int32, C.c_int32]
    lib.lidf_depth_metrics_rays.restype = C.c_int
    lib.lidf_depth_metrics_rays.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
    lib.lidf_depth_metrics_image.restype = C.c_int
    lib.lidf_depth_metrics_image.argtypes = [C.c_void_p] * 5 + [C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
    lib.lidf_refine_workspace_bytes.restype = C.c_size_t
    lib.lidf_refine_workspace_bytes.argtypes = [C.POINTER(_RefineParams)]
    lib.lidf_refine_forward.restype = C.c_int
    lib.lidf_refine_forward.argtypes = [C.POINTER(_RefineParams), C.c_void_p]
    lib.lidf_roi_align_rays.restype = C.c_int
    lib.lidf_roi_align_rays.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_int64,
                                        C.c_int32, C.c_void_p, C.c_void_p]
    lib.lidf_ray_terminate_workspace_bytes.restype = C.c_size_t
    lib.lidf_ray_terminate_workspace_bytes.argtypes = [C.c_int64, C.c_int64]
    lib.lidf_ray_terminate.restype = C.c_int
    lib.lidf_ray_terminate.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64,
                                       C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
    lib.lidf_ray_loss_workspace_bytes.restype = C.c_size_t
    lib.lidf_ray_loss_workspace_bytes.argtypes = [C.c_int64, C.c_int64]
    lib.lidf_ray_loss.restype = C.c_int
    lib.lidf_ray_loss.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p,
                                  C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
    lib.lidf_image_loss_workspace_bytes.restype = C.c_size_t
    lib.lidf_image_loss_workspace_bytes.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.c_int64]
    lib.lidf_image_loss.restype = C.c_int
    lib.lidf_image_loss.argtypes = [C.c_void_p] * 5 + [C.c_int32, C.c_int32, C.c_int32, C.c_int64] + [C.c_void_p] * 4 + [C.c_size_t, C.c_void_p]
    lib.lidf_query_last_mlp_ms.restype = C.c_float
    lib.lidf_tc_selftest.restype = C.c_int
    lib.lidf_tc_selftest.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]
    lib.lidf_query_launch_count.restype = C.c_int64
    lib.lidf_query_launch_count.argtypes = [C.c_int]
    vp, i64, sz = C.c_void_p, C.c_int64, C.c_size_t
    lib.lidf_ray_aabb_workspace_bytes.restype = sz
    lib.lidf_ray_aabb_workspace_bytes.argtypes = [i64, i64]
    lib.lidf_ray_aabb_forward.restype = C.c_int
    lib.lidf_ray_aabb_forward.argtypes = [vp, vp, vp, vp, i64, i64, vp, vp, vp, sz, vp]
    lib.lidf_ray_aabb_pairs_count.restype = C.c_int
    lib.lidf_ray_aabb_pairs_count.argtypes = [vp, vp, vp, vp, i64, i64, vp, sz, C.POINTER(C.c_int64), vp]
    lib.lidf_ray_aabb_pairs_fill.restype = C.c_int
    lib.lidf_ray_aabb_pairs_fill.argtypes = [vp, vp, vp, vp, i64, i64, vp, sz, i64, vp, vp, vp, vp]
    lib.lidf_ray_aabb_ray_major_workspace_bytes.restype = sz
    lib.lidf_ray_aabb_ray_major_workspace_bytes.argtypes = [i64, i64]
    lib.lidf_ray_aabb_pairs_ray_major_count.restype = C.c_int
    lib.lidf_ray_aabb_pairs_ray_major_count.argtypes = [vp, vp, vp, vp, i64, i64, vp, sz, C.POINTER(C.c_int64), vp]
    lib.lidf_ray_aabb_pairs_ray_major_fill.restype = C.c_int
    lib.lidf_ray_aabb_pairs_ray_major_fill.argtypes = [vp, vp, vp, vp, i64, i64, vp, sz, i64, vp, vp, vp, vp, vp]
    lib.lidf_pcl_aabb_forward.restype = C.c_int
    lib.lidf_pcl_aabb_forward.argtypes = [vp, vp, vp, vp, i64, i64, vp, vp]
    lib.lidf_pcl_aabb_pair_label.restype = C.c_int
    lib.lidf_pcl_aabb_pair_label.argtypes = [vp, vp, vp, vp, i64, i64, vp, vp, i64, vp, vp]
    lib.lidf_pcl_aabb_end_voxel.restype = C.c_int
    lib.lidf_pcl_aabb_end_voxel.argtypes = [vp, vp, vp, vp, i64, i64, vp, vp]
    i32, f32 = C.c_int32, C.c_float
    vox_common = [vp, vp, i64, i32, f32, f32, f32, f32, f32, i32, i32, i32, vp, sz]
    lib.lidf_voxelize_workspace_bytes.restype = sz
    lib.lidf_voxelize_workspace_bytes.argtypes = [i64, i32, i32, i32, i32]
    lib.lidf_voxelize_count.restype = C.c_int
    lib.lidf_voxelize_count.argtypes = vox_common + [C.POINTER(C.c_int64), C.POINTER(C.c_int64), vp]
    lib.lidf_voxelize_fill.restype = C.c_int
    lib.lidf_voxelize_fill.argtypes = vox_common + [vp, vp, vp, vp, vp, vp]
    if lib.lidf_query_abi_version() != ABI_VERSION:
        raise RuntimeError("liblidf_query.so ABI version mismatch")
    lib.lidf_query_struct_size.restype = C.c_size_t
    lib.lidf_query_struct_size.argtypes = [C.c_int]
    for which, st in enumerate((_Decoder, _QueryParams, _RefineParams, _BackwardParams)):
        if lib.lidf_query_struct_size(which) != C.sizeof(st):
            raise RuntimeError(f"ctypes layout of {st.__name__} does not match the compiled library")
    return lib


def bind_to_device_numa_node(device_index: int):
    """Pin the calling process to the CPUs of the NUMA node the GPU hangs off (sysfs), so that the pinned host buffers of
    ``forward_host`` are allocated next to the GPU's PCIe root port.  One process per GPU (the reference's mp.spawn /
    torchrun model) otherwise lands on arbitrary sockets and the H2D / D2H streams of 8 ranks cross the socket
    interconnect.  Returns (node, previous_affinity) or (None, previous_affinity) when the topology cannot be read."""
    prev = os.sched_getaffinity(0)
    try:
        props = torch.cuda.get_device_properties(device_index)
        if all(hasattr(props, a) for a in ("pci_domain_id", "pci_bus_id", "pci_device_id")):
            bus = f"{props.pci_domain_id:04x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0"
        else:
            import subprocess
            out = subprocess.run(["nvidia-smi", "-i", str(device_index), "--query-gpu=pci.bus_id", "--format=csv,noheader"],
                                 capture_output=True, text=True).stdout.strip().lower()
            bus = out[-12:] if len(out) >= 12 else out             # nvidia-smi prints an 8-digit domain
        with open(f"/sys/bus/pci/devices/{bus}/numa_node") as f:
            node = int(f.read().strip())
        if node < 0:
            return None, prev
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = set()
            for part in f.read().strip().split(","):
                a, _, b = part.partition("-")
                cpus.update(range(int(a), int(b or a) + 1))
        cpus &= prev
        if not cpus:
            return None, prev
        os.sched_setaffinity(0, cpus)
        return node, prev
    except Exception:
        return None, prev


def _chk(t: Optional[torch.Tensor], name: str, dtype, optional: bool = False):
    """The reference's CHECK_INPUT (ray_aabb_cuda.cpp:16-18) plus a dtype check."""
    if t is None:
        if optional:
            return 0
        raise RuntimeError(f"{name} must not be None")
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor")
    if not t.is_contiguous():
        raise RuntimeError(f"{name} must be contiguous")
    if t.dtype != dtype:
        raise RuntimeError(f"{name} must be {dtype}, got {t.dtype}")
    return t.data_ptr()


def decoder_state(dec) -> Dict[str, torch.Tensor]:
    """Accept an IMNet/IEF module (ours or the reference's) or a plain state-dict-like mapping."""
    if isinstance(dec, dict):
        return dec
    return {k: v for k, v in dec.state_dict(keep_vars=True).items()}


def _decoder_struct(sd: Dict[str, torch.Tensor], n_iter: int, use_sigmoid: bool, keep: list, name: str) -> _Decoder:
    d = _Decoder()
    is_ief = "offset_enc.weight" in sd
    d.kind = 1 if is_ief else 0
    d.n_iter = int(n_iter) if is_ief else 1
    d.use_sigmoid = int(bool(use_sigmoid))
    d.init_offset = 0.001                                  # IEF.init_offset, implicit_net.py:104
    w1 = sd["linear_1.weight"]
    if tuple(w1.shape[:1]) != (256,) or tuple(sd["linear_2.weight"].shape) != (128, 256) \
            or tuple(sd["linear_3.weight"].shape) != (64, 128) or tuple(sd["linear_4.weight"].shape) != (1, 64):
        raise RuntimeError(f"{name}: lidf_query supports imnet_gf=64, out_dim=1 decoders only (shipped YAMLs)")
    d.inp_dim = int(w1.shape[1]) - (16 if is_ief else 0)
    for field, key in (("w1", "linear_1.weight"), ("b1", "linear_1.bias"), ("w2", "linear_2.weight"),
                       ("b2", "linear_2.bias"), ("w3", "linear_3.weight"), ("b3", "linear_3.bias"),
                       ("w4", "linear_4.weight"), ("b4", "linear_4.bias")):
        t = sd[key].detach()
        keep.append(t)
        setattr(d, field, _chk(t, f"{name}.{key}", torch.float32))
    if is_ief:
        for field, key in (("w_enc", "offset_enc.weight"), ("b_enc", "offset_enc.bias")):
            t = sd[key].detach()
            keep.append(t)
            setattr(d, field, _chk(t, f"{name}.{key}", torch.float32))
    return d


class _HostCall:
    """One in-flight ``forward_host_async`` batch."""

    def __init__(self, lq, host, offset_dec, prob_dec, dev, out_host, kw, h2d, d2h, outputs):
        self.lq, self.host, self.offset_dec, self.prob_dec, self.dev = lq, host, offset_dec, prob_dec, dev
        self.out_host, self.kw, self.h2d, self.d2h, self.outputs = out_host, kw, h2d, d2h, outputs
        self.pending = None
        self.finished = False

    def run_monolithic(self):
        lq, dev = self.lq, self.dev
        ins = [lq._widen(k, self.host[k].to(dev, non_blocking=True)) for k in lq.INPUT_KEYS]
        out = lq.forward(*ins, self.offset_dec, self.prob_dec, **self.kw)
        for k in self.outputs:
            self.out_host[k].copy_(out[k], non_blocking=True)
        done = torch.cuda.Event(); done.record(torch.cuda.current_stream(dev))
        self.pending = (done, None, (ins, out), None)

    def wait(self):
        """Block until the outputs are in ``out_host``.  If a pipelined batch turns out not to be image-contiguous
        (device-side range check), it is redone as one monolithic call."""
        if not self.finished:
            done, bad_host, _live, _bad = self.pending
            done.synchronize()
            if bad_host is not None and int(bad_host[0]) != 0:
                self.run_monolithic()
                self.pending[0].synchronize()
            self.pending = None
            self.finished = True
        return self.out_host, self.h2d, self.d2h


class _LidfQuery:
    """Object with a ``forward`` like the reference's pybind modules (ray_aabb.forward, pcl_aabb.forward)."""

    def __init__(self):
        self._lib = None

    @property
    def lib(self) -> C.CDLL:
        if self._lib is None:
            self._lib = load_library()
        return self._lib

    def _raise(self, rc: int, what: str):
        if rc != 0:
            msg = self.lib.lidf_query_error_string(rc).decode()
            cu = self.lib.lidf_query_last_cuda_error().decode()
            raise RuntimeError(f"{what} failed: {msg}" + (f" [{cu}]" if rc == -4 else ""))

    def launch_count(self, reset: bool = False) -> int:
        return int(self.lib.lidf_query_launch_count(1 if reset else 0))

    def last_mlp_ms(self) -> float:
        """Device time of the decoder kernel in the latest forward (CUDA events on the launching stream)."""
        return float(self.lib.lidf_query_last_mlp_ms())

    def tc_selftest(self, A: torch.Tensor, W: torch.Tensor, variant: int = 0) -> torch.Tensor:
        """D = A @ W.T through the engine's tcgen05 primitives (A [128,32], W [128,32] fp32 CUDA)."""
        dev = A.device
        D = torch.empty(128, 128, dtype=torch.float32, device=dev)
        scratch = torch.empty(16384, dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            rc = self.lib.lidf_tc_selftest(_chk(A, "A", torch.float32), _chk(W, "W", torch.float32), D.data_ptr(),
                                           scratch.data_ptr(), int(variant),
                                           C.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
        self._raise(rc, "lidf_tc_selftest")
        torch.cuda.synchronize(dev)
        return D

    INPUT_KEYS = ("full_rgb_feat", "occ_voxel_feat", "miss_ray_dir", "miss_img_ind", "miss_bid", "voxel_bound",
                  "occ_vox_intersect_idx", "miss_ray_intersect_idx", "intersect_dist")
    OUTPUT_KEYS = ("pred_offset", "pred_prob_end", "pair_pred_pos", "pred_prob_end_softmax", "max_pair_id", "pred_pos")

    INDEX_KEYS = ("miss_img_ind", "miss_bid", "occ_vox_intersect_idx", "miss_ray_intersect_idx")

    @staticmethod
    def _widen(key, t):
        """Host index arrays may be int32 (half the H2D bytes of the reference's int64 ``torch.nonzero`` output); the
        kernels read int64, so they are widened on the device right after the copy."""
        return t.long() if (key in _LidfQuery.INDEX_KEYS and t.dtype == torch.int32) else t

    def image_splits(self, host: Dict[str, torch.Tensor], B: int):
        """Per-image offsets into the ray / voxel / pair arrays ([B+1] python lists) for the pipelined host path.
        The reference keeps rays and voxels image-major (``miss_bid`` / ``occ_vox_bid`` sorted: pipeline.py:226-262,
        point_utils.py:44-60) and its pair list voxel-major (``torch.nonzero`` of mask[V,R], pipeline.py:283-285), so
        every image owns one contiguous slice of each array.  Found by binary search (O(B log P) on the host);
        ``forward_host`` verifies on the device that each slice really is self-contained."""
        if "occ_vox_bid" not in host:
            return None
        edges = torch.arange(B + 1, dtype=torch.int64)
        rays = torch.searchsorted(host["miss_bid"], edges.to(host["miss_bid"].dtype)).tolist()
        voxels = torch.searchsorted(host["occ_vox_bid"].to(torch.int64), edges).tolist()
        pairs = torch.searchsorted(host["occ_vox_intersect_idx"],
                                   torch.tensor(voxels, dtype=host["occ_vox_intersect_idx"].dtype)).tolist()
        return dict(rays=rays, voxels=voxels, pairs=pairs)

    def forward_host(self, host: Dict[str, torch.Tensor], offset_dec, prob_dec, device, out_host=None,
                     pipeline: bool = True, min_chunk_pairs: int = 1 << 22, outputs=None, **kw):
        """Same call with HOST buffers (pinned for async copies): H2D of every input, the fused forward, D2H of every
        output, then a synchronise.  Returns (out_host dict, h2d_bytes, d2h_bytes).

        The path shards by image, so with ``pipeline`` the batch is cut into groups of whole
        images and run as a three-stage pipeline on three CUDA streams -- H2D of group i+1 and D2H of group i-1 overlap
        the decoder kernels of group i.  Needs ``host['occ_vox_bid']`` (the reference's data_dict key) next to
        INPUT_KEYS to find the per-image slices; without it, or when the arrays are not image-contiguous, the whole
        batch goes through one copy-in / compute / copy-out sequence."""
        return self.forward_host_async(host, offset_dec, prob_dec, device, out_host=out_host, pipeline=pipeline,
                                       min_chunk_pairs=min_chunk_pairs, outputs=outputs, **kw).wait()

    def forward_host_async(self, host: Dict[str, torch.Tensor], offset_dec, prob_dec, device, out_host=None,
                           pipeline: bool = True, min_chunk_pairs: int = 1 << 22, outputs=None, **kw) -> "_HostCall":
        """``forward_host`` without the final synchronise: enqueues the whole batch and returns a handle whose ``wait()``
        gives (out_host, h2d_bytes, d2h_bytes).  Back-to-back batches (serving) keep the GPU busy across batch boundaries:
        the H2D of batch k+1 runs under the tail of batch k, its D2H drains under the head of batch k+2.  Give every
        in-flight batch its own ``out_host`` buffers.

        ``outputs``: which of OUTPUT_KEYS are copied back (default: all six).  Inference reads ``pred_pos`` (+
        ``max_pair_id`` for RefineNet, reference pipeline.py:942); the four per-pair tensors are 24 B per query point of
        D2H that only the training losses look at.  The index arrays in ``host`` may be int32 (see ``_widen``)."""
        outputs = tuple(outputs) if outputs is not None else self.OUTPUT_KEYS
        if kw.get("winner_only") and any(k in ("pred_offset", "pair_pred_pos") for k in outputs):
            raise RuntimeError("forward_host(winner_only=True): pred_offset / pair_pred_pos are not produced in that mode; "
                               "ask for outputs=('pred_pos', 'max_pair_id', ...)")
        dev = torch.device(device)
        B = int(host["full_rgb_feat"].shape[0])
        P = int(host["occ_vox_intersect_idx"].shape[0]); R = int(host["miss_ray_dir"].shape[0])
        h2d = sum(host[k].numel() * host[k].element_size() for k in self.INPUT_KEYS)
        if host["intersect_dist"].dim() != 2 or kw.get("pcl_label_float") is not None:
            pipeline = False                                # dense dist[V,R,2] / label branch: monolithic call
        splits = self.image_splits(host, B) if (pipeline and B > 1) else None
        groups = []
        if splits is not None:
            b0 = 0
            for b in range(1, B + 1):
                if splits["pairs"][b] - splits["pairs"][b0] >= min_chunk_pairs or b == B:
                    groups.append((b0, b)); b0 = b
        if out_host is None:
            f32 = dict(dtype=torch.float32, pin_memory=True)
            shapes = dict(pred_offset=((P, 1), f32), pred_prob_end=((P, 1), f32), pair_pred_pos=((P, 3), f32),
                          pred_prob_end_softmax=((P,), f32), max_pair_id=((R,), dict(dtype=torch.int64, pin_memory=True)),
                          pred_pos=((R, 3), f32))
            out_host = {k: torch.empty(*shapes[k][0], **shapes[k][1]) for k in outputs}
        d2h = sum(out_host[k].numel() * out_host[k].element_size() for k in outputs)
        call = _HostCall(self, host, offset_dec, prob_dec, dev, out_host, kw, h2d, d2h, outputs)
        if len(groups) > 1:
            self._enqueue_host_pipeline(call, splits, groups)
        else:
            call.run_monolithic()
        return call

    GEOMETRY_KEYS = ("full_rgb_feat", "occ_voxel_feat", "miss_ray_dir", "miss_img_ind", "miss_bid", "voxel_bound", "occ_vox_bid")

    def forward_host_from_geometry(self, host: Dict[str, torch.Tensor], offset_dec, prob_dec, device, outputs=("pred_pos", "max_pair_id"),
                                   **kw):
        """The call a real pipeline makes when the pair list does not exist on the host yet (reference
        ``compute_ray_aabb`` -> ``get_embedding`` -> ``get_pred``, pipeline.py:271-296,338-466): HOST buffers of the two feature
        tensors, the rays and the occupied voxels' boxes (``GEOMETRY_KEYS``; ``miss_bid`` / ``occ_vox_bid`` int32 or int64) go
        to the device, the pairs are generated THERE, ray-major (``ray_aabb.pairs(order="ray")``: no regroup in the forward),
        and the requested outputs come back.  H2D per query point: nothing -- the 16-24 bytes per pair of index / distance
        arrays never cross PCIe.  Returns (out_host dict incl. the device-side pair count ``n_pairs``, h2d_bytes, d2h_bytes).
        ``max_pair_id`` indexes the ray-major list, which is also returned on request (``outputs`` may name
        ``occ_vox_intersect_idx`` / ``miss_ray_intersect_idx`` / ``intersect_dist``)."""
        from implicit_depth_b200.extensions.ray_aabb.jit import ray_aabb
        dev = torch.device(device)
        h2d = sum(host[k].numel() * host[k].element_size() for k in self.GEOMETRY_KEYS)
        g = {k: host[k].to(dev, non_blocking=True) for k in self.GEOMETRY_KEYS}
        vox, ray, dist = ray_aabb.pairs(g["miss_ray_dir"], g["voxel_bound"], g["miss_bid"].int(), g["occ_vox_bid"].int(), order="ray")
        out = self.forward(g["full_rgb_feat"], g["occ_voxel_feat"], g["miss_ray_dir"], g["miss_img_ind"].long(), g["miss_bid"].long(),
                           g["voxel_bound"], vox, ray, dist, offset_dec, prob_dec, pairs_ray_major=True, **kw)
        out.update(occ_vox_intersect_idx=vox, miss_ray_intersect_idx=ray, intersect_dist=dist)
        res = {k: out[k].to("cpu", non_blocking=True) for k in outputs}
        torch.cuda.current_stream(dev).synchronize()
        res["n_pairs"] = int(vox.shape[0])
        d2h = sum(v.numel() * v.element_size() for k, v in res.items() if k != "n_pairs")
        return res, h2d, d2h

    def _enqueue_host_pipeline(self, call: "_HostCall", splits, groups) -> None:
        """Three-stage pipeline over image groups; nothing here waits on the host."""
        host, dev, out_host, kw = call.host, call.dev, call.out_host, call.kw
        P = int(host["occ_vox_intersect_idx"].shape[0])
        if getattr(self, "_host_streams", None) is None or self._host_streams[0].device != dev:
            self._host_streams = tuple(torch.cuda.Stream(dev) for _ in range(3))
        s_in, s_cmp, s_out = self._host_streams
        cur = torch.cuda.current_stream(dev)
        for s in (s_in, s_cmp, s_out):
            s.wait_stream(cur)
        bad = torch.zeros(1, dtype=torch.int64, device=dev)
        live = []                                           # keeps every staged tensor alive until wait()
        for (b0, b1) in groups:
            r0, r1 = splits["rays"][b0], splits["rays"][b1]
            v0, v1 = splits["voxels"][b0], splits["voxels"][b1]
            p0, p1 = splits["pairs"][b0], splits["pairs"][b1]
            with torch.cuda.stream(s_in):
                up = lambda t: t.to(dev, non_blocking=True)
                ins = dict(full_rgb_feat=up(host["full_rgb_feat"][b0:b1]), occ_voxel_feat=up(host["occ_voxel_feat"][v0:v1]),
                           miss_ray_dir=up(host["miss_ray_dir"][r0:r1]), miss_img_ind=up(host["miss_img_ind"][r0:r1]),
                           miss_bid=up(host["miss_bid"][r0:r1]), voxel_bound=up(host["voxel_bound"][v0:v1]),
                           occ_vox_intersect_idx=up(host["occ_vox_intersect_idx"][p0:p1]),
                           miss_ray_intersect_idx=up(host["miss_ray_intersect_idx"][p0:p1]),
                           intersect_dist=up(host["intersect_dist"][p0:p1]))
                ev_in = torch.cuda.Event(); ev_in.record(s_in)
            with torch.cuda.stream(s_cmp):
                s_cmp.wait_event(ev_in)
                raw = {k: ins[k] for k in self.INDEX_KEYS}      # the uploaded (possibly int32) tensors live on s_in's pool:
                live.append(raw)                                # keep them until wait(), or the next group's upload reuses them
                for k in self.INDEX_KEYS:
                    ins[k] = self._widen(k, raw[k])
                # re-base the group's indices to its own slices and check that the slice is self-contained
                pv, pr, mb = ins["occ_vox_intersect_idx"], ins["miss_ray_intersect_idx"], ins["miss_bid"]
                pv.sub_(v0); pr.sub_(r0); mb.sub_(b0)
                if p1 > p0:
                    lo_v, hi_v = torch.aminmax(pv); lo_r, hi_r = torch.aminmax(pr)
                    bad += ((lo_v < 0) | (hi_v >= v1 - v0) | (lo_r < 0) | (hi_r >= r1 - r0)).to(torch.int64)
                    pv.clamp_(0, max(v1 - v0 - 1, 0)); pr.clamp_(0, max(r1 - r0 - 1, 0))   # never fault; `bad` discards
                if r1 > r0:
                    lo_b, hi_b = torch.aminmax(mb)
                    bad += ((lo_b < 0) | (hi_b >= b1 - b0)).to(torch.int64)
                    mb.clamp_(0, b1 - b0 - 1)
                out = self.forward(*[ins[k] for k in self.INPUT_KEYS], call.offset_dec, call.prob_dec, **kw)
                mp = out["max_pair_id"]                     # local pair ids -> ids in the whole list; empty ray -> P
                out["max_pair_id"] = torch.where(mp == p1 - p0, P, mp + p0)
                ev_cmp = torch.cuda.Event(); ev_cmp.record(s_cmp)
            with torch.cuda.stream(s_out):
                s_out.wait_event(ev_cmp)
                for k in call.outputs:
                    sl = slice(r0, r1) if k in ("max_pair_id", "pred_pos") else slice(p0, p1)
                    out_host[k][sl].copy_(out[k], non_blocking=True)
            live.append((ins, out))
        with torch.cuda.stream(s_out):
            bad_host = torch.empty(1, dtype=torch.int64, pin_memory=True)
            bad_host.copy_(bad, non_blocking=True)
            done = torch.cuda.Event(); done.record(s_out)
        call.pending = (done, bad_host, live, bad)

    # ------------------------------------------------------------------ fused get_embedding + get_pred
    def _query_params(self, full_rgb_feat, occ_voxel_feat, miss_ray_dir, miss_img_ind, miss_bid, voxel_bound,
                      occ_vox_intersect_idx, miss_ray_intersect_idx, dist, offset_dec, prob_dec, keep, *,
                      part_size, pos_encode=True, multires=8, multires_views=4, intersect_pos_type="abs", roi_inp_bbox=8,
                      roi_out_bbox=2, n_iter=2, use_sigmoid=False, offset_range=(0.0, 1.0), pcl_label_float=None,
                      mlp_impl="auto", pairs_ray_major=False) -> _QueryParams:
        """Input half of the parameter block (shared by forward and backward): shape / dtype / device checks included."""
        if roi_out_bbox != 2 or tuple(full_rgb_feat.shape[1:2]) != (32,) or occ_voxel_feat.shape[-1] != 128:
            raise RuntimeError("lidf_query supports rgb_out=32, roi_out_bbox=2, pnet_out=128 only (shipped YAMLs)")
        P = int(occ_vox_intersect_idx.shape[0]); R = int(miss_ray_dir.shape[0]); V = int(occ_voxel_feat.shape[0])
        B, _, H, W = (int(s) for s in full_rgb_feat.shape)
        # shapes the kernels index with (the reference's CHECK_INPUT does not look at shapes; an out-of-bounds read would)
        for name, t, shape in (("occ_voxel_feat", occ_voxel_feat, (V, 128)), ("miss_ray_dir", miss_ray_dir, (R, 3)),
                               ("miss_img_ind", miss_img_ind, (R, 2)), ("miss_bid", miss_bid, (R,)),
                               ("voxel_bound", voxel_bound, (V, 6)), ("occ_vox_intersect_idx", occ_vox_intersect_idx, (P,)),
                               ("miss_ray_intersect_idx", miss_ray_intersect_idx, (P,))):
            if tuple(t.shape) != shape:
                raise RuntimeError(f"{name} must have shape {shape}, got {tuple(t.shape)}")
        if pcl_label_float is not None and tuple(pcl_label_float.shape) != (P,):
            raise RuntimeError(f"pcl_label_float must have shape {(P,)}, got {tuple(pcl_label_float.shape)}")
        p = _QueryParams()
        p.P, p.R, p.V, p.B, p.H, p.W = P, R, V, B, H, W
        p.full_rgb_feat = _chk(full_rgb_feat, "full_rgb_feat", torch.float32)
        p.occ_voxel_feat = _chk(occ_voxel_feat, "occ_voxel_feat", torch.float32)
        p.miss_ray_dir = _chk(miss_ray_dir, "miss_ray_dir", torch.float32)
        p.miss_img_ind = _chk(miss_img_ind, "miss_img_ind", torch.int64)
        p.miss_bid = _chk(miss_bid, "miss_bid", torch.int64)
        p.voxel_bound = _chk(voxel_bound, "voxel_bound", torch.float32)
        p.pair_vox = _chk(occ_vox_intersect_idx, "occ_vox_intersect_idx", torch.int64)
        p.pair_ray = _chk(miss_ray_intersect_idx, "miss_ray_intersect_idx", torch.int64)
        if dist.dim() == 2:
            if tuple(dist.shape) != (P, 2):
                raise RuntimeError("per-pair dist must be [P,2]")
            p.pair_dist = _chk(dist, "dist", torch.float32)
        else:
            if tuple(dist.shape) != (V, R, 2):
                raise RuntimeError("dense dist must be [V,R,2]")
            p.dense_dist = _chk(dist, "dist", torch.float32)
        p.pcl_label_float = _chk(pcl_label_float, "pcl_label_float", torch.float32, optional=True)
        p.pos_encode = int(bool(pos_encode)); p.multires = int(multires); p.multires_views = int(multires_views)
        p.intersect_pos_rel = int(intersect_pos_type == "rel"); p.roi_inp_bbox = int(roi_inp_bbox)
        p.offset_range0, p.offset_range1 = float(offset_range[0]), float(offset_range[1])
        p.part_size = float(part_size)
        p.offset_dec = _decoder_struct(decoder_state(offset_dec), n_iter, use_sigmoid, keep, "offset_dec")
        p.prob_dec = _decoder_struct(decoder_state(prob_dec), n_iter, use_sigmoid, keep, "prob_dec")
        p.mlp_impl = MLP_IMPLS[mlp_impl]
        p.pairs_ray_major = int(bool(pairs_ray_major))
        return p

    use_weight_cache = True

    def _attach_weight_cache(self, p: _QueryParams, keep: list, dev) -> None:
        """Packed decoder weights persist across calls in a device buffer owned by this wrapper (one per device, stream
        and engine).  They are re-packed only when a decoder tensor was replaced or modified in place (``data_ptr`` /
        ``_version`` of every tensor -- optimizer steps, ``load_state_dict`` and ``nn.init`` all bump the version) or a
        setting that enters the packing changed.  The cache entry keeps the decoder tensors alive, so an address can not be
        recycled by a different tensor while it is cached.  In-place writes through ``param.data`` bypass the version
        counter: call ``invalidate_weight_cache()`` after those (or set ``use_weight_cache = False``)."""
        if not self.use_weight_cache:
            return
        stream = torch.cuda.current_stream(dev).cuda_stream
        slot = (dev.index if dev.index is not None else torch.cuda.current_device(), stream, int(p.mlp_impl))
        key = tuple((t.data_ptr(), t._version, tuple(t.shape)) for t in keep) + (
            int(p.pos_encode), int(p.multires), int(p.multires_views), int(p.offset_dec.kind), int(p.offset_dec.n_iter),
            float(p.offset_dec.init_offset))
        nbytes = int(self.lib.lidf_query_weight_cache_bytes(C.byref(p)))
        if nbytes == 0:
            return
        caches = self.__dict__.setdefault("_wcaches", {})
        buf, old_key, _alive = caches.get(slot, (None, None, None))
        if buf is None or buf.numel() < nbytes:
            buf, old_key = torch.empty(nbytes, dtype=torch.uint8, device=dev), None
        p.weight_cache, p.weight_cache_bytes = buf.data_ptr(), buf.numel()
        p.weight_cache_valid = int(old_key == key)
        caches[slot] = (buf, key, list(keep))            # strong refs: the storages behind `key` stay allocated

    def invalidate_weight_cache(self) -> None:
        self.__dict__.pop("_wcaches", None)

    def check_index_errors(self, wait: bool = True) -> None:
        """Raise if an earlier ``forward`` saw an out-of-range pair_ray / pair_vox / miss_bid value.  The kernels clamp
        such indices (no out-of-bounds access) and set a device flag that is copied to pinned host memory right behind
        the call; it is looked at here -- without a sync when ``wait`` is false (every later ``forward`` does that), after
        waiting for the copy when true.  The reference trips a device-side assert in the same situation."""
        pend, self._pending_flags = getattr(self, "_pending_flags", []), []
        for flag_host, ev in pend:
            if wait:
                ev.synchronize()
            if not ev.query():
                self._pending_flags.append((flag_host, ev))
            elif int(flag_host[0]) != 0:
                self._pending_flags = []
                raise RuntimeError("lidf_query_forward: index out of range in miss_ray_intersect_idx / occ_vox_intersect_idx "
                                   f"/ miss_bid, or a pair list passed as pairs_ray_major that is not sorted by ray (flag "
                                   f"{int(flag_host[0])}: 1 ray, 2 voxel / image, 4 order); the affected outputs are meaningless")

    def forward(self, full_rgb_feat, occ_voxel_feat, miss_ray_dir, miss_img_ind, miss_bid, voxel_bound,
                occ_vox_intersect_idx, miss_ray_intersect_idx, dist, offset_dec, prob_dec, *,
                part_size: float, pos_encode: bool = True, multires: int = 8, multires_views: int = 4,
                intersect_pos_type: str = "abs", roi_inp_bbox: int = 8, roi_out_bbox: int = 2,
                n_iter: int = 2, use_sigmoid: bool = False, offset_range: Sequence[float] = (0.0, 1.0),
                pcl_label_float: Optional[torch.Tensor] = None, mlp_impl: str = "auto",
                want_roi_feat: bool = False, save_for_backward: bool = False,
                check_indices: bool = False, pairs_ray_major: bool = False,
                winner_only: bool = False) -> Dict[str, torch.Tensor]:
        """Everything LIDF.get_embedding + LIDF.get_pred compute after the ResNet / PointNet producers
        (reference src/models/pipeline.py:338-466).  ``dist`` is either the per-pair [P,2] enter/leave distances or
        the reference's dense [V,R,2] tensor.  Returns the data_dict entries of pipeline.py:460-466 (+ pred_offset).
        ``save_for_backward`` adds ``ief_iter`` (the IEF offsets between iterations, which ``backward`` needs);
        ``check_indices`` waits for the call and raises on an out-of-range index (otherwise the next call reports it).
        ``pairs_ray_major``: the caller vouches that ``miss_ray_intersect_idx`` is non-decreasing (what
        ``ray_aabb.pairs(..., order="ray")`` emits); the in-call regroup by ray is then a binary search per ray.  A list
        that is not sorted raises like an out-of-range index (flag bit 2).
        ``winner_only``: for callers that read the per-ray results only (everything downstream of get_pred in the reference
        does: compute_loss and RefineNet use pred_pos, max_pair_id, pred_prob_end, pred_prob_end_softmax; pair_pred_pos and
        pred_offset are written to data_dict and never read).  The probability decoder runs over all pairs, the rays are
        terminated, and the offset decoder runs on each ray's arg-max pair only -- ``pred_pos`` comes out bit-identical to the
        full call at a third of the decoder work (64 pairs per ray).  ``pred_offset`` / ``pair_pred_pos`` are not returned."""
        capturing = torch.cuda.is_current_stream_capturing() if full_rgb_feat.is_cuda else False
        if not capturing:
            self.check_index_errors(wait=False)
        dev = full_rgb_feat.device
        keep: list = []
        p = self._query_params(full_rgb_feat, occ_voxel_feat, miss_ray_dir, miss_img_ind, miss_bid, voxel_bound,
                               occ_vox_intersect_idx, miss_ray_intersect_idx, dist, offset_dec, prob_dec, keep,
                               part_size=part_size, pos_encode=pos_encode, multires=multires, multires_views=multires_views,
                               intersect_pos_type=intersect_pos_type, roi_inp_bbox=roi_inp_bbox, roi_out_bbox=roi_out_bbox,
                               n_iter=n_iter, use_sigmoid=use_sigmoid, offset_range=offset_range,
                               pcl_label_float=pcl_label_float, mlp_impl=mlp_impl, pairs_ray_major=pairs_ray_major)
        P, R = int(p.P), int(p.R)
        f32 = dict(dtype=torch.float32, device=dev)
        out = dict(pred_prob_end=torch.empty(P, 1, **f32), pred_prob_end_softmax=torch.empty(P, **f32),
                   max_pair_id=torch.empty(R, dtype=torch.int64, device=dev), pred_pos=torch.empty(R, 3, **f32))
        if winner_only:
            p.winner_only_offset = 1
            if save_for_backward:                            # the winners' offsets, by ray: what backward(winner_only=True) reads
                out["pred_offset_ray"] = torch.empty(R, **f32)
                p.pred_offset_ray = out["pred_offset_ray"].data_ptr()
        else:
            out.update(pred_offset=torch.empty(P, 1, **f32), pair_pred_pos=torch.empty(P, 3, **f32))
        if want_roi_feat:
            out["roi_feat_per_ray"] = torch.empty(R, 128, **f32)
            p.roi_feat_per_ray = out["roi_feat_per_ray"].data_ptr()
        if save_for_backward:
            n_it = int(p.offset_dec.n_iter) if p.offset_dec.kind == 1 else 1
            out["ief_iter"] = torch.empty(max(n_it - 1, 0), R if winner_only else P, **f32)
            p.ief_iter_out = out["ief_iter"].data_ptr() if n_it > 1 and P > 0 else None
        flag = torch.zeros(1, dtype=torch.int32, device=dev)
        p.index_error = flag.data_ptr()
        self._attach_weight_cache(p, keep, dev)
        for k in ("pred_offset", "pred_prob_end", "pair_pred_pos", "pred_prob_end_softmax", "max_pair_id", "pred_pos"):
            if k in out:
                setattr(p, k, out[k].data_ptr())
        nbytes = int(self.lib.lidf_query_workspace_bytes(C.byref(p)))
        if nbytes == 0:
            raise RuntimeError("lidf_query: unsupported configuration (lidf_query_workspace_bytes returned 0)")
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        p.workspace, p.workspace_bytes = ws.data_ptr(), nbytes
        with torch.cuda.device(dev):
            cur = torch.cuda.current_stream(dev)
            rc = self.lib.lidf_query_forward(C.byref(p), C.c_void_p(cur.cuda_stream))
            self._raise(rc, "lidf_query_forward")
            if capturing:                       # inside a CUDA graph: no host-side bookkeeping; the flag stays on the device
                out["index_error"] = flag
                return out
            flag_host = torch.empty(1, dtype=torch.int32, pin_memory=True)
            flag_host.copy_(flag, non_blocking=True)
            ev = torch.cuda.Event(); ev.record(cur)
        ws.record_stream(cur); flag.record_stream(cur)
        if not hasattr(self, "_pending_flags"):
            self._pending_flags = []
        self._pending_flags.append((flag_host, ev))
        if check_indices:
            self.check_index_errors(wait=True)
        return out

    def make_graphed_forward(self, full_rgb_feat, occ_voxel_feat, miss_ray_dir, miss_img_ind, miss_bid, voxel_bound,
                             occ_vox_intersect_idx, miss_ray_intersect_idx, dist, offset_dec, prob_dec, **kw):
        """Inference with fixed shapes (same image size, same pair count -- e.g. BASELINE config 1, or a serving loop over
        equally sized batches): the whole ``forward`` -- ~15 launches, 0.3 ms of Python / launch latency for 0.07 ms of
        kernels at config 1 -- is captured once into a CUDA graph and replayed with one launch.  Returns
        ``run(*nine_input_tensors) -> out`` (pass nothing to re-run on the static buffers, ``run.inputs``); the outputs are
        the graph's static tensors, overwritten by the next replay.  The packed decoder weights are taken from the weight
        cache (filled by the warm-up on the capture stream), so the graph holds no packing kernels; after an in-place
        update of a decoder tensor call ``run.repack()`` (one eager call that re-packs into the same buffer), after
        replacing a tensor build a new graph."""
        dev = full_rgb_feat.device
        static = [t.clone() for t in (full_rgb_feat, occ_voxel_feat, miss_ray_dir, miss_img_ind, miss_bid, voxel_bound,
                                      occ_vox_intersect_idx, miss_ray_intersect_idx, dist)]
        kw.pop("check_indices", None)
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(2):                                          # warm-up outside the capture (function attributes, allocator)
                self.forward(*static, offset_dec, prob_dec, **kw)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.check_index_errors(wait=True)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=side):                      # same stream as the warm-up: its weight cache is valid
            out = self.forward(*static, offset_dec, prob_dec, **kw)

        def run(*new_inputs):
            for dst, src in zip(static, new_inputs):
                if src is not dst:
                    dst.copy_(src, non_blocking=True)
            graph.replay()
            return out

        def repack():
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                self.forward(*static, offset_dec, prob_dec, **kw)
            torch.cuda.current_stream(dev).wait_stream(side)
        run.inputs, run.graph, run.repack = static, graph, repack
        return run

    def backward(self, full_rgb_feat, occ_voxel_feat, miss_ray_dir, miss_img_ind, miss_bid, voxel_bound,
                 occ_vox_intersect_idx, miss_ray_intersect_idx, dist, offset_dec, prob_dec, fwd_out, *,
                 g_pred_pos=None, g_pred_prob_end=None, g_pred_offset=None, g_pair_pred_pos=None,
                 need_feat_grad: bool = True, need_vox_grad: bool = True, chunk_rows: int = 0, winner_only: bool = False, **kw):
        """Backward of ``forward`` (include/lidf_query.h: lidf_query_backward): what torch autograd does for the reference's
        get_embedding + get_pred under ``loss_net.backward()`` (reference src/trainers/train_lidf.py:394).  ``fwd_out`` is
        the dict ``forward(..., save_for_backward=True)`` returned; ``kw`` are the same scalar settings.  Returns
        ``dict(full_rgb_feat=..., occ_voxel_feat=..., offset_dec={state_dict key: grad}, prob_dec={...})``.
        ``winner_only``: ``fwd_out`` comes from ``forward(..., winner_only=True, save_for_backward=True)``; the offset decoder's
        backward then runs over one row per ray (the only rows whose upstream gradient is non-zero when the loss reads
        ``pred_pos`` and ``pred_prob_end``, as the reference's does); ``g_pred_offset`` / ``g_pair_pred_pos`` must be None."""
        dev = full_rgb_feat.device
        keep: list = []
        kw.pop("want_roi_feat", None); kw.pop("save_for_backward", None); kw.pop("check_indices", None)
        bp = _BackwardParams()
        bp.fwd = self._query_params(full_rgb_feat, occ_voxel_feat, miss_ray_dir, miss_img_ind, miss_bid, voxel_bound,
                                    occ_vox_intersect_idx, miss_ray_intersect_idx, dist, offset_dec, prob_dec, keep, **kw)
        P, R, V = int(bp.fwd.P), int(bp.fwd.R), int(bp.fwd.V)
        f32 = dict(dtype=torch.float32, device=dev)
        if winner_only:
            if g_pred_offset is not None or g_pair_pred_pos is not None:
                raise RuntimeError("backward(winner_only=True): pred_offset / pair_pred_pos are not outputs of that mode")
            bp.fwd.winner_only_offset = 1
            bp.fwd.pred_offset_ray = _chk(fwd_out["pred_offset_ray"], "pred_offset_ray", torch.float32)
        else:
            bp.fwd.pred_offset = _chk(fwd_out["pred_offset"], "pred_offset", torch.float32)
        bp.fwd.pred_prob_end = _chk(fwd_out["pred_prob_end"], "pred_prob_end", torch.float32)
        bp.fwd.max_pair_id = _chk(fwd_out["max_pair_id"], "max_pair_id", torch.int64)
        n_it = int(bp.fwd.offset_dec.n_iter) if bp.fwd.offset_dec.kind == 1 else 1
        if n_it > 1 and P > 0:
            it = fwd_out.get("ief_iter")
            if it is None or tuple(it.shape) != (n_it - 1, R if winner_only else P):
                raise RuntimeError("backward needs fwd_out['ief_iter'] [n_iter-1, P] ([n_iter-1, R] in winner-only mode): "
                                   "call forward(save_for_backward=True)")
            bp.ief_iter = _chk(it, "ief_iter", torch.float32)
        for name, t, shape in (("g_pred_pos", g_pred_pos, (R, 3)), ("g_pred_prob_end", g_pred_prob_end, (P, 1)),
                               ("g_pred_offset", g_pred_offset, (P, 1)), ("g_pair_pred_pos", g_pair_pred_pos, (P, 3))):
            if t is not None:
                if tuple(t.shape) != shape and tuple(t.shape) != shape[:1]:
                    raise RuntimeError(f"{name} must have shape {shape}, got {tuple(t.shape)}")
                t = t.contiguous(); keep.append(t)
                setattr(bp, name, _chk(t, name, torch.float32))
        res = dict(full_rgb_feat=torch.empty_like(full_rgb_feat) if need_feat_grad else None,
                   occ_voxel_feat=torch.empty_like(occ_voxel_feat) if need_vox_grad else None)
        bp.g_full_rgb_feat = res["full_rgb_feat"].data_ptr() if need_feat_grad else None
        bp.g_occ_voxel_feat = res["occ_voxel_feat"].data_ptr() if need_vox_grad else None
        names = (("w1", "linear_1.weight"), ("b1", "linear_1.bias"), ("w2", "linear_2.weight"), ("b2", "linear_2.bias"),
                 ("w3", "linear_3.weight"), ("b3", "linear_3.bias"), ("w4", "linear_4.weight"), ("b4", "linear_4.bias"),
                 ("w_enc", "offset_enc.weight"), ("b_enc", "offset_enc.bias"))
        for field, dec in (("g_offset_dec", offset_dec), ("g_prob_dec", prob_dec)):
            sd = decoder_state(dec)
            gd, gs = _DecoderGrad(), {}
            for f, key in names:
                if key in sd:
                    gs[key] = torch.empty(sd[key].shape, **f32)
                    setattr(gd, f, gs[key].data_ptr())
            setattr(bp, field, gd)
            res[field[2:]] = gs
        bp.chunk_rows = int(chunk_rows)
        nbytes = int(self.lib.lidf_query_backward_workspace_bytes(C.byref(bp)))
        if nbytes == 0:
            raise RuntimeError("lidf_query_backward: unsupported configuration")
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        bp.workspace, bp.workspace_bytes = ws.data_ptr(), nbytes
        with torch.cuda.device(dev):
            cur = torch.cuda.current_stream(dev)
            rc = self.lib.lidf_query_backward(C.byref(bp), C.c_void_p(cur.cuda_stream))
        self._raise(rc, "lidf_query_backward")
        ws.record_stream(cur)
        return res

    def last_bwd_ms(self) -> float:
        """Device time of the backward's tcgen05 section in the latest ``backward`` (CUDA events on the launching stream)."""
        return float(self.lib.lidf_query_last_bwd_ms())

    def wgrad_selftest(self, A: torch.Tensor, B: torch.Tensor, packed: bool = False) -> torch.Tensor:
        """C = A.T @ B through k_wgrad_tc (A [rows,M], M in {128,256}; B [rows,N], N % 16 == 0, N <= 256), or with
        ``packed`` through the backward's packed hand-over path (k_pk_pack_rows -> k_wgrad_pk_tc, M = 128)."""
        dev = A.device
        rows, M = (int(v) for v in A.shape); N = int(B.shape[1])
        out = torch.empty(M, N, dtype=torch.float32, device=dev)
        nbytes = self.lib.lidf_wgrad_pk_selftest_scratch_bytes(rows, M, N) if packed else self.lib.lidf_wgrad_selftest_scratch_bytes(M, N)
        scratch = torch.empty(int(nbytes), dtype=torch.uint8, device=dev)
        fn = self.lib.lidf_wgrad_pk_selftest if packed else self.lib.lidf_wgrad_selftest
        with torch.cuda.device(dev):
            rc = fn(_chk(A, "A", torch.float32), _chk(B, "B", torch.float32), out.data_ptr(), rows, M, N,
                    scratch.data_ptr(), C.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
        self._raise(rc, "lidf_wgrad_pk_selftest" if packed else "lidf_wgrad_selftest")
        torch.cuda.synchronize(dev)
        return out

    # ------------------------------------------------------------------ RefineNet decoder tail
    def refine_forward(self, pred_pos, miss_ray_dir, end_voxel_center, voxel_feat_end, rgb_feat_end, offset_dec, *,
                       pos_encode: bool = True, multires: int = 8, multires_views: int = 4,
                       intersect_pos_type: str = "abs", n_iter: int = 2, use_sigmoid: bool = False,
                       offset_range: Sequence[float] = (-0.2, 0.2), mlp_impl: str = "auto",
                       occ_voxel_feat: Optional[torch.Tensor] = None, end_voxel_id: Optional[torch.Tensor] = None,
                       voxel_bound: Optional[torch.Tensor] = None) -> torch.Tensor:
        """pipeline.py:1018-1029 for all rays at once.

        Either pass the gathered ``voxel_feat_end`` [R,128] (+ ``end_voxel_center`` for 'rel'): fp32 FFMA engine.  Or pass
        what the reference holds just before that gather -- ``occ_voxel_feat`` [V,128], ``end_voxel_id`` [R] (and
        ``voxel_bound`` [V,6] for 'rel') -- and the decoder runs on the tcgen05 engine (``mlp_impl`` "auto")."""
        dev = pred_pos.device
        R = int(pred_pos.shape[0])
        keep: list = []
        p = _RefineParams()
        p.R = R
        p.pred_pos = _chk(pred_pos, "pred_pos", torch.float32)
        p.miss_ray_dir = _chk(miss_ray_dir, "miss_ray_dir", torch.float32)
        ungathered = occ_voxel_feat is not None and end_voxel_id is not None
        if ungathered and mlp_impl == "simt_fp32" and voxel_feat_end is None:
            voxel_feat_end = occ_voxel_feat[end_voxel_id].contiguous()           # the reference's gather, pipeline.py:1016
            if intersect_pos_type == "rel" and end_voxel_center is None:
                vb = voxel_bound[end_voxel_id]
                end_voxel_center = ((vb[:, :3] + vb[:, 3:]) / 2.).contiguous()
        p.end_voxel_center = _chk(end_voxel_center, "end_voxel_center", torch.float32, optional=True)
        p.voxel_feat_end = _chk(voxel_feat_end, "voxel_feat_end", torch.float32, optional=ungathered)
        p.rgb_feat_end = _chk(rgb_feat_end, "rgb_feat_end", torch.float32)
        if ungathered:
            if tuple(end_voxel_id.shape) != (R,) or occ_voxel_feat.dim() != 2 or occ_voxel_feat.shape[1] != 128:
                raise RuntimeError("occ_voxel_feat must be [V,128] and end_voxel_id [R]")
            p.V = int(occ_voxel_feat.shape[0])
            p.occ_voxel_feat = _chk(occ_voxel_feat, "occ_voxel_feat", torch.float32)
            p.end_voxel_id = _chk(end_voxel_id, "end_voxel_id", torch.int64)
            p.voxel_bound = _chk(voxel_bound, "voxel_bound", torch.float32, optional=intersect_pos_type != "rel")
        p.pos_encode = int(bool(pos_encode)); p.multires = int(multires); p.multires_views = int(multires_views)
        p.intersect_pos_rel = int(intersect_pos_type == "rel")
        p.offset_range0, p.offset_range1 = float(offset_range[0]), float(offset_range[1])
        p.offset_dec = _decoder_struct(decoder_state(offset_dec), n_iter, use_sigmoid, keep, "offset_dec")
        p.mlp_impl = MLP_IMPLS[mlp_impl]
        out = torch.empty(R, 3, dtype=torch.float32, device=dev)
        p.pred_pos_refine = out.data_ptr()
        nbytes = int(self.lib.lidf_refine_workspace_bytes(C.byref(p)))
        if nbytes == 0:
            raise RuntimeError("lidf_refine: unsupported configuration")
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        p.workspace, p.workspace_bytes = ws.data_ptr(), nbytes
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            rc = self.lib.lidf_refine_forward(C.byref(p), C.c_void_p(stream))
        self._raise(rc, "lidf_refine_forward")
        ws.record_stream(torch.cuda.current_stream(dev))
        return out

    # ------------------------------------------------------------------ stand-alone pieces
    def roi_align_rays(self, full_rgb_feat, miss_img_ind, miss_bid, roi_inp_bbox: int = 8) -> torch.Tensor:
        dev = full_rgb_feat.device
        B, Cc, H, W = (int(s) for s in full_rgb_feat.shape)
        if Cc != 32:
            raise RuntimeError("rgb_out must be 32")
        R = int(miss_bid.shape[0])
        out = torch.empty(R, 128, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            rc = self.lib.lidf_roi_align_rays(_chk(full_rgb_feat, "full_rgb_feat", torch.float32), B, H, W,
                                              _chk(miss_img_ind, "miss_img_ind", torch.int64),
                                              _chk(miss_bid, "miss_bid", torch.int64), R, int(roi_inp_bbox),
                                              out.data_ptr(), C.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
        self._raise(rc, "lidf_roi_align_rays")
        return out

    def ray_terminate(self, pred_prob_end, miss_ray_intersect_idx, pair_pred_pos, R: int,
                      pcl_label_float: Optional[torch.Tensor] = None):
        dev = pair_pred_pos.device
        P = int(miss_ray_intersect_idx.shape[0])
        logit = pred_prob_end.reshape(-1)
        soft = torch.empty(P, dtype=torch.float32, device=dev)
        arg = torch.empty(R, dtype=torch.int64, device=dev)
        pos = torch.empty(R, 3, dtype=torch.float32, device=dev)
        nbytes = int(self.lib.lidf_ray_terminate_workspace_bytes(P, R))
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            rc = self.lib.lidf_ray_terminate(_chk(logit, "pred_prob_end", torch.float32),
                                             _chk(miss_ray_intersect_idx, "miss_ray_intersect_idx", torch.int64),
                                             _chk(pair_pred_pos, "pair_pred_pos", torch.float32),
                                             _chk(pcl_label_float, "pcl_label_float", torch.float32, optional=True),
                                             P, R, soft.data_ptr(), arg.data_ptr(), pos.data_ptr(), ws.data_ptr(), nbytes,
                                             C.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
        self._raise(rc, "lidf_ray_terminate")
        ws.record_stream(torch.cuda.current_stream(dev))
        return soft, arg, pos

    def ray_loss(self, pred_prob_end, pred_prob_end_softmax, miss_ray_intersect_idx, pcl_label_float, R: int,
                 pred_pos: Optional[torch.Tensor] = None, gt_pos: Optional[torch.Tensor] = None):
        """The ray-keyed torch_scatter part of LIDF.compute_loss (pipeline.py:482-486, :553-557) and the per-ray position
        errors (:472, :560-567) in two kernels.  Returns a dict: log_softmax [P], pred_label / gt_label [R], and 0-dim
        tensors prob_loss, acc (+ pos_loss, err when gt_pos is given) computed from the device-side sums (no host sync)."""
        dev = pred_prob_end_softmax.device
        P = int(miss_ray_intersect_idx.shape[0])
        logit = pred_prob_end.reshape(-1)
        lsm = torch.empty(P, dtype=torch.float32, device=dev)
        pl = torch.empty(R, dtype=torch.int64, device=dev)
        gl = torch.empty(R, dtype=torch.int64, device=dev)
        stats = torch.empty(6, dtype=torch.float64, device=dev)
        nbytes = int(self.lib.lidf_ray_loss_workspace_bytes(P, R))
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            rc = self.lib.lidf_ray_loss(_chk(logit, "pred_prob_end", torch.float32),
                                        _chk(pred_prob_end_softmax, "pred_prob_end_softmax", torch.float32),
                                        _chk(miss_ray_intersect_idx, "miss_ray_intersect_idx", torch.int64),
                                        _chk(pcl_label_float, "pcl_label_float", torch.float32), P, R,
                                        _chk(pred_pos, "pred_pos", torch.float32, optional=True),
                                        _chk(gt_pos, "gt_pos", torch.float32, optional=True),
                                        lsm.data_ptr(), pl.data_ptr(), gl.data_ptr(), stats.data_ptr(), ws.data_ptr(), nbytes,
                                        C.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
        self._raise(rc, "lidf_ray_loss")
        ws.record_stream(torch.cuda.current_stream(dev))
        out = dict(log_softmax=lsm, pred_label=pl, gt_label=gl, stats=stats,
                   prob_loss=(stats[0] / stats[1]).float(), acc=(stats[2] / max(R, 1)).float())
        if gt_pos is not None:
            out["pos_loss"] = (stats[3] / max(3 * R, 1)).float()
            out["err"] = torch.where(stats[5] > 0, stats[4] / stats[5].clamp_min(1), torch.zeros_like(stats[4])).float()
        return out

    def image_loss(self, xyz_flat, miss_bid, miss_flat_img_id, pred_pos, gt_pos, h: int, w: int, want_normal_imgs: bool = False):
        """The image-space terms of LIDF.compute_loss (pipeline.py:494-541): returns 0-dim tensors surf_norm_loss, angle_err
        (degrees), smooth_loss (no host sync) and, on request, the two [B,3,H,W] surface-normal images."""
        dev = xyz_flat.device
        B, R = int(xyz_flat.shape[0]), int(miss_bid.shape[0])
        if tuple(xyz_flat.shape) != (B, h * w, 3):
            raise RuntimeError("xyz_flat must be [B, h*w, 3]")
        stats = torch.empty(6, dtype=torch.float64, device=dev)
        imgs = [torch.empty(B, 3, h, w, dtype=torch.float32, device=dev) for _ in range(2)] if want_normal_imgs else [None, None]
        nbytes = int(self.lib.lidf_image_loss_workspace_bytes(B, h, w, R))
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            rc = self.lib.lidf_image_loss(_chk(xyz_flat, "xyz_flat", torch.float32), _chk(miss_bid, "miss_bid", torch.int64),
                                          _chk(miss_flat_img_id, "miss_flat_img_id", torch.int64),
                                          _chk(pred_pos, "pred_pos", torch.float32), _chk(gt_pos, "gt_pos", torch.float32),
                                          B, int(h), int(w), R, imgs[0].data_ptr() if want_normal_imgs else None,
                                          imgs[1].data_ptr() if want_normal_imgs else None, stats.data_ptr(), ws.data_ptr(), nbytes,
                                          C.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
        self._raise(rc, "lidf_image_loss")
        ws.record_stream(torch.cuda.current_stream(dev))
        n = max(R, 1)
        out = dict(stats=stats, surf_norm_loss=(stats[0] / n).float(), angle_err=(stats[1] / n * (180.0 / math.pi)).float(),
                   smooth_loss=((stats[2] + stats[3]) / n).float())
        if want_normal_imgs:
            out["pred_surf_norm_img"], out["gt_surf_norm_img"] = imgs
        return out


METRIC_KEYS = ("a1", "a2", "a3", "rmse", "rmse_log", "log10", "abs_rel", "mae", "sq_rel")


def _metrics_from_stats(stats: torch.Tensor) -> Dict[str, torch.Tensor]:
    """stats[12] (include/lidf_query.h) -> the reference's nine depth metrics as 0-dim fp32 tensors (no host sync)."""
    n = stats[0]
    out = {"a1": stats[1] / n, "a2": stats[2] / n, "a3": stats[3] / n, "rmse": (stats[4] / n).sqrt(),
           "rmse_log": (stats[5] / n).sqrt(), "log10": stats[6] / n, "abs_rel": stats[7] / n, "mae": stats[8] / n,
           "sq_rel": stats[9] / n}
    return {k: v.float() for k, v in out.items()}


def _depth_metrics(self, pred_pos, gt_pos=None, *, xyz_flat=None, xyz_corrupt_flat=None, corrupt_mask=None,
                   miss_flat_img_id=None, h: int = 0, w: int = 0):
    """Depth metrics of LIDF.compute_loss for exp_type != 'train' (reference pipeline.py:570-618).  With ``gt_pos``: the
    bs != 1 branch over rays.  With ``xyz_flat`` / ``xyz_corrupt_flat`` [1,H*W,3], ``corrupt_mask`` [1,H,W] and
    ``miss_flat_img_id``: the bs == 1 branch (cv2 nearest-neighbour resampling to 256x144) entirely on the device."""
    dev = pred_pos.device
    R = int(pred_pos.shape[0])
    stats = torch.empty(12, dtype=torch.float64, device=dev)
    nbytes = int(self.lib.lidf_depth_metrics_workspace_bytes(R, int(h), int(w)))
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        if gt_pos is not None:
            rc = self.lib.lidf_depth_metrics_rays(_chk(pred_pos, "pred_pos", torch.float32), _chk(gt_pos, "gt_pos", torch.float32), R,
                                                  stats.data_ptr(), ws.data_ptr(), nbytes, st)
        else:
            if tuple(xyz_flat.shape) != (1, h * w, 3) or tuple(xyz_corrupt_flat.shape) != (1, h * w, 3) or corrupt_mask.numel() != h * w:
                raise RuntimeError("depth_metrics (bs == 1): xyz_flat / xyz_corrupt_flat must be [1,h*w,3], corrupt_mask [1,h,w]")
            rc = self.lib.lidf_depth_metrics_image(_chk(xyz_flat, "xyz_flat", torch.float32),
                                                   _chk(xyz_corrupt_flat, "xyz_corrupt_flat", torch.float32),
                                                   _chk(corrupt_mask, "corrupt_mask", torch.float32),
                                                   _chk(miss_flat_img_id, "miss_flat_img_id", torch.int64),
                                                   _chk(pred_pos, "pred_pos", torch.float32), R, int(h), int(w),
                                                   stats.data_ptr(), ws.data_ptr(), nbytes, st)
    self._raise(rc, "lidf_depth_metrics")
    ws.record_stream(torch.cuda.current_stream(dev))
    out = _metrics_from_stats(stats)
    out["stats"] = stats
    return out


_LidfQuery.depth_metrics = _depth_metrics
lidf_query = _LidfQuery()
