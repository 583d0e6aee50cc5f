"""CPU: pin oracle/aabb_oracle.py (numpy restatement of ray_aabb / pcl_aabb + their torch glue) against golden vectors
produced by the reference's OWN kernels compiled for the CPU (tests/golden/aabb_*.npz, made by
tests/golden/make_golden_aabb.py through oracle/build_ref.py).  Bit-exact: integer masks, and fp32 distances compared
as bit patterns."""
import os

import numpy as np
import pytest

from conftest import GOLDEN_DIR
from oracle import aabb_oracle as A
from oracle import build_ref

AABB_CASES = ["aabb_grid_2x12x16", "aabb_grid_1x9x11", "aabb_edge"]


def load(name):
    return dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz")))


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


@pytest.mark.parametrize("name", AABB_CASES)
def test_ray_aabb_oracle_is_bit_exact_against_reference_kernel(name):
    z = load(name)
    mask, dist = A.ray_aabb_dense(z["ray_dir"], z["voxel_bound"], z["ray_bid"], z["voxel_bid"])
    assert mask.dtype == np.int32 and np.array_equal(mask, z["ref_mask"])
    assert np.array_equal(bits(dist), bits(z["ref_dist"]))
    vox, ray, pd = A.ray_aabb_pairs(z["ray_dir"], z["voxel_bound"], z["ray_bid"], z["voxel_bid"])
    assert np.array_equal(vox, z["ref_pair_vox"]) and np.array_equal(ray, z["ref_pair_ray"])
    assert np.array_equal(bits(pd), bits(z["ref_pair_dist"]))
    assert mask.sum() > 0 and (z["ray_bid"][ray] == z["voxel_bid"][vox]).all()


@pytest.mark.parametrize("name", AABB_CASES)
def test_pcl_aabb_oracle_is_bit_exact_against_reference_kernel(name):
    z = load(name)
    m = A.pcl_aabb_dense(z["pts"], z["voxel_bound"], z["pts_bid"], z["voxel_bid"])
    assert np.array_equal(m, z["ref_pcl_mask"]) and m.sum() > 0
    m2 = A.pcl_aabb_dense(z["ray_pts"], z["voxel_bound"], z["ray_bid"], z["voxel_bid"])
    assert np.array_equal(m2, z["ref_pcl_mask_rays"])
    lab = A.pcl_pair_label(z["ray_pts"], z["voxel_bound"], z["ray_bid"], z["voxel_bid"], z["ref_pair_vox"], z["ref_pair_ray"])
    assert np.array_equal(lab, z["ref_pair_label"])
    end = A.pcl_end_voxel(z["ray_pts"], z["voxel_bound"], z["ray_bid"], z["voxel_bid"], z["end_voxel_start"])
    assert np.array_equal(end, z["ref_end_voxel"])
    assert (end >= z["end_voxel_start"]).all() and (end != z["end_voxel_start"]).any()


def test_edge_fixture_really_contains_the_edge_cases():
    z = load("aabb_edge")
    d = z["ray_dir"]
    assert (d == 0).any() and np.signbit(d[d == 0]).any()            # +0 and -0 direction components
    assert not (np.diff(z["voxel_bid"]) >= 0).all()                  # unsorted image ids
    assert 7 in z["voxel_bid"] and 7 not in z["ray_bid"]             # an image without rays
    assert z["ref_mask"][z["voxel_bid"] == 7].sum() == 0
    vb = z["voxel_bound"]
    assert ((vb[:, 3:] - vb[:, :3]) == 0).any()                      # zero-thickness voxel


@pytest.mark.skipif(not os.path.exists(build_ref.LIB), reason="oracle/_ref not built (needs /root/reference)")
def test_oracle_vs_reference_kernel_on_fresh_random_inputs():
    """When oracle/_ref/libaabb_ref.so is present (build container, or shipped to the GPU box), compare directly."""
    ref = build_ref.load()
    rng = np.random.default_rng(123)
    for trial in range(4):
        R, V = int(rng.integers(1, 700)), int(rng.integers(1, 60))
        d = rng.normal(size=(R, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
        d[rng.random((R, 3)) < 0.05] = 0.0
        lo = rng.uniform(-1, 1, size=(V, 3)); vb = np.concatenate((lo, lo + rng.uniform(0, 1.5, size=(V, 3))), 1)
        rb = rng.integers(0, 3, size=R); xb = rng.integers(0, 3, size=V)
        m, dist = A.ray_aabb_dense(d, vb, rb, xb)
        rm, rd = ref.ray_aabb(d, vb, rb, xb)
        assert np.array_equal(m, rm) and np.array_equal(bits(dist), bits(rd))
        pts = rng.uniform(-1.2, 2.2, size=(R, 3))
        assert np.array_equal(A.pcl_aabb_dense(pts, vb, rb, xb), ref.pcl_aabb(pts, vb, rb, xb))


VOXEL_CASES = ["voxel_res8_3img", "voxel_res8_unsorted", "voxel_res5"]


@pytest.mark.parametrize("name", VOXEL_CASES)
def test_voxelisation_oracle_is_bit_exact_against_reference_code(name):
    z = load(name)
    out = A.get_occ_vox_bound(z["valid_xyz"], z["valid_bid"], int(z["res"]))
    assert out["part_size"] == float(z["part_size"]) and np.array_equal(bits(out["xmin"]), bits(z["xmin"]))
    assert np.array_equal(out["occ_vox_bid"], z["ref_occ_vox_bid"])
    assert np.array_equal(out["occ_vox_global_coord"], z["ref_occ_vox_global_coord"])
    assert np.array_equal(bits(out["voxel_bound"]), bits(z["ref_voxel_bound"]))
    assert np.array_equal(out["revidx"], z["ref_revidx"]) and np.array_equal(out["valid_v_pid"], z["ref_valid_v_pid"])
    assert np.array_equal(bits(out["valid_v_rel_coord"]), bits(z["ref_valid_v_rel_coord"]))
    assert 0 < out["valid_v_pid"].size < z["valid_xyz"].shape[0]          # some points fall outside the grid


@pytest.mark.skipif(not os.path.exists(build_ref.LIB), reason="oracle/_ref not built (needs /root/reference)")
def test_oracle_vs_reference_kernel_on_non_finite_and_extreme_inputs():
    """inf / nan / denormal / huge components in directions, boxes and points: the restatement must follow the reference
    kernel's comparisons and fmaxf / fminf exactly (NaN compares false, fmax ignores a NaN operand)."""
    ref = build_ref.load()
    rng = np.random.default_rng(7)
    special = np.array([0.0, -0.0, np.inf, -np.inf, np.nan, 1e-45, -1e-45, 1e-38, 3e38, -3e38, 1e-12, -1e-12, 1.0, -1.0], np.float32)
    for trial in range(6):
        R, V = 300, 40
        d = rng.normal(size=(R, 3)).astype(np.float32)
        vb = np.concatenate((rng.uniform(-1, 1, (V, 3)), rng.uniform(1, 2, (V, 3))), 1).astype(np.float32)
        pts = rng.uniform(-1.5, 2.5, size=(R, 3)).astype(np.float32)
        for arr, frac in ((d, 0.15), (vb, 0.08), (pts, 0.15)):
            m = rng.random(arr.shape) < frac
            arr[m] = rng.choice(special, size=int(m.sum()))
        rb = rng.integers(0, 2, size=R); xb = rng.integers(0, 2, size=V)
        with np.errstate(all="ignore"):
            m, dist = A.ray_aabb_dense(d, vb, rb, xb)
        rm, rd = ref.ray_aabb(d, vb, rb, xb)
        assert np.array_equal(m, rm), trial
        hit = rm.astype(bool)
        assert np.array_equal(bits(dist)[hit], bits(rd)[hit]) or np.array_equal(np.nan_to_num(dist[hit], nan=7.0), np.nan_to_num(rd[hit], nan=7.0))
        assert np.array_equal(A.pcl_aabb_dense(pts, vb, rb, xb), ref.pcl_aabb(pts, vb, rb, xb)), trial
