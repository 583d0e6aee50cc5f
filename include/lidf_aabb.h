/* lidf_aabb.h -- C ABI of the B200-native ray/point vs voxel-box tests that feed the LIDF query path.
 *
 * Replaces the reference's two native extensions and the torch glue directly around them:
 *
 *   lidf_ray_aabb_forward      <- ray_aabb_forward / ray_aabb_cuda_forward
 *                                 (src/extensions/ray_aabb/ray_aabb_cuda.cpp:20-32, ray_aabb_cuda_kernel.cu:10-126)
 *   lidf_ray_aabb_pairs_*      <- the same test + torch.nonzero(mask) + dist[vox, ray]
 *                                 (src/models/pipeline.py:277-285, :345-346) without the dense [V,R] slab
 *   lidf_pcl_aabb_forward      <- pcl_aabb_forward / pcl_aabb_cuda_forward
 *                                 (src/extensions/pcl_aabb/pcl_aabb_cuda.cpp:20-32, pcl_aabb_cuda_kernel.cu:10-80)
 *   lidf_pcl_aabb_pair_label   <- pcl_mask[occ_vox_intersect_idx, miss_ray_intersect_idx].float()  (pipeline.py:305-309)
 *   lidf_pcl_aabb_end_voxel    <- pcl_aabb + nonzero + scatter(reduce='max', out=end_voxel_id)     (pipeline.py:939-944)
 *
 * Same conventions as the reference bindings: forward only, borrowed read-only contiguous device arrays, fp32
 * geometry, int32 image ids (the reference casts with .int(), pipeline.py:278).  Outputs go to caller-owned device
 * buffers (every element is written; no pre-zeroing needed).  Results are bit-identical to the reference kernels,
 * including the double-precision reciprocal 1/(d + 1e-12).  Return 0 or a negative LIDF_ERR_* code (lidf_query.h).
 */
#ifndef LIDF_AABB_H_
#define LIDF_AABB_H_

#include <stddef.h>
#include <stdint.h>

#include "lidf_query.h"

#ifdef __cplusplus
extern "C" {
#endif

/* scratch for any of the ray_aabb calls below (reciprocal directions, per-ray-block image-id ranges, pair counts) */
size_t lidf_ray_aabb_workspace_bytes(int64_t R, int64_t V);

/* dense drop-in: mask [V,R] int32, dist [V,R,2] fp32 */
int lidf_ray_aabb_forward(const float* ray_dir /*[R,3]*/, const float* voxel_bound /*[V,6]*/,
                          const int32_t* ray_bid /*[R]*/, const int32_t* voxel_bid /*[V]*/, int64_t R, int64_t V,
                          int32_t* mask, float* dist, void* workspace, size_t workspace_bytes, lidf_stream_t stream);

/* compact pair list in the reference's order (sorted by voxel, then ray == torch.nonzero of mask[V,R]):
 * step 1 counts and leaves the scanned counts in the workspace; *n_pairs_host (host memory) receives the number of
 * pairs after a stream synchronise -- the same sync torch.nonzero does.  Step 2 writes the P pairs; it must see the
 * same inputs and the untouched workspace of step 1. */
int lidf_ray_aabb_pairs_count(const float* ray_dir, const float* voxel_bound, const int32_t* ray_bid,
                              const int32_t* voxel_bid, int64_t R, int64_t V, void* workspace, size_t workspace_bytes,
                              int64_t* n_pairs_host, lidf_stream_t stream);
int lidf_ray_aabb_pairs_fill(const float* ray_dir, const float* voxel_bound, const int32_t* ray_bid,
                             const int32_t* voxel_bid, int64_t R, int64_t V, void* workspace, size_t workspace_bytes,
                             int64_t P, int64_t* pair_vox /*[P]*/, int64_t* pair_ray /*[P]*/, float* pair_dist /*[P,2]*/,
                             lidf_stream_t stream);

/* The same pairs sorted by ray, then voxel (every ray's pairs adjacent, ascending voxel id): the order the consumer of
 * the list works in -- lidf_query_forward with LidfQueryParams::pairs_ray_major = 1 skips its regroup.  Same two-step
 * protocol with its own workspace; step 2 can also hand out the CSR offsets ray_start [R+1] int32 (NULL to skip). */
size_t lidf_ray_aabb_ray_major_workspace_bytes(int64_t R, int64_t V);
int lidf_ray_aabb_pairs_ray_major_count(const float* ray_dir, const float* voxel_bound, const int32_t* ray_bid,
                                        const int32_t* voxel_bid, int64_t R, int64_t V, void* workspace,
                                        size_t workspace_bytes, int64_t* n_pairs_host, lidf_stream_t stream);
int lidf_ray_aabb_pairs_ray_major_fill(const float* ray_dir, const float* voxel_bound, const int32_t* ray_bid,
                                       const int32_t* voxel_bid, int64_t R, int64_t V, void* workspace,
                                       size_t workspace_bytes, int64_t P, int64_t* pair_vox /*[P]*/, int64_t* pair_ray /*[P]*/,
                                       float* pair_dist /*[P,2]*/, int32_t* ray_start /*[R+1] or NULL*/, lidf_stream_t stream);

/* dense drop-in: mask [V,N] int32 */
int lidf_pcl_aabb_forward(const float* pcl_pos /*[N,3]*/, const float* voxel_bound /*[V,6]*/,
                          const int32_t* pcl_bid /*[N]*/, const int32_t* voxel_bid /*[V]*/, int64_t N, int64_t V,
                          int32_t* mask, lidf_stream_t stream);
/* label[i] = inside(pcl_pos[pair_ray[i]], voxel pair_vox[i]) as 0.0f / 1.0f */
int lidf_pcl_aabb_pair_label(const float* pcl_pos, const float* voxel_bound, const int32_t* pcl_bid,
                             const int32_t* voxel_bid, int64_t N, int64_t V, const int64_t* pair_vox,
                             const int64_t* pair_ray, int64_t P, float* label /*[P]*/, lidf_stream_t stream);
/* end_voxel_id[n] = max(end_voxel_id[n], largest v with point n inside voxel v); in/out [N] int64 */
int lidf_pcl_aabb_end_voxel(const float* pcl_pos, const float* voxel_bound, const int32_t* pcl_bid,
                            const int32_t* voxel_bid, int64_t N, int64_t V, int64_t* end_voxel_id, lidf_stream_t stream);

/* Voxelisation of the valid points: point_utils.batch_get_occupied_idx (src/utils/point_utils.py:12-76, overlap = False)
 * as called by LIDF.get_occ_vox_bound (src/models/pipeline.py:162-201).  Grid: origin xmin[3], cubic cells of edge
 * part_size, n[3] cells per axis, B images; half_part = (float)(0.5 * part_size) as the reference forms it in double.
 * Step 1 marks occupied cells and in-grid points, scans both, and returns the two counts (one sync, where the reference
 * syncs on torch.unique); step 2 writes
 *   occ_vox_bid_global_coord [V,4] int64 (image, cx, cy, cz), sorted as torch.unique(dim=0) sorts them
 *   voxel_bound [V,6] = [xmin + c*part, that + part]            (pipeline.py:186-188)
 *   revidx [Nv] int64, valid_v_pid [Nv] int64, valid_v_rel_coord [Nv,3]   (point_utils.py:66-74) */
size_t lidf_voxelize_workspace_bytes(int64_t Np, int32_t B, int32_t nx, int32_t ny, int32_t nz);
int lidf_voxelize_count(const float* valid_xyz /*[Np,3]*/, const int64_t* valid_bid /*[Np]*/, int64_t Np, int32_t B,
                        float xmin0, float xmin1, float xmin2, float part_size, float half_part, int32_t nx, int32_t ny,
                        int32_t nz, void* workspace, size_t workspace_bytes, int64_t* n_voxels_host,
                        int64_t* n_inside_host, lidf_stream_t stream);
int lidf_voxelize_fill(const float* valid_xyz, const int64_t* valid_bid, int64_t Np, int32_t B, float xmin0, float xmin1,
                       float xmin2, float part_size, float half_part, int32_t nx, int32_t ny, int32_t nz, void* workspace,
                       size_t workspace_bytes, int64_t* occ_vox_bid_global_coord, float* voxel_bound, int64_t* revidx,
                       int64_t* valid_v_pid, float* valid_v_rel_coord, lidf_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* LIDF_AABB_H_ */
