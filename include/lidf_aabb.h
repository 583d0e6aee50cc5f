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

#ifdef __cplusplus
}
#endif
#endif /* LIDF_AABB_H_ */
