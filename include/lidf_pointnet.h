/* lidf_pointnet.h -- C ABI of the PointNet2Stage forward, the producer of occ_voxel_feat for the LIDF query path.
 *
 * Replaces `PointNet2Stage.forward(inp_feat, vox2point_idx)` (reference src/models/pointnet.py:22-38) as it is called by
 * `LIDF.get_embedding` (src/models/pipeline.py:400-407) and, with the predicted points appended, by
 * `RefineNet.get_pred_refine` (pipeline.py:1000-1014).  Dimensions of the shipped YAMLs only (train_lidf.yaml:43-46:
 * pnet_in 6, pnet_gf 32, pnet_out 128).  Weights are the nn.Linear tensors of the state_dict ([out,in] row-major).
 * Forward only, fp32, borrowed read-only device arrays; 0 or a negative LIDF_ERR_* code (lidf_query.h).
 */
#ifndef LIDF_POINTNET_H_
#define LIDF_POINTNET_H_

#include <stddef.h>
#include <stdint.h>

#include "lidf_query.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct LidfPointNet {
  const float* point_lin1_w; const float* point_lin1_b;   /* [32,6],    [32]  */
  const float* point_lin2_w; const float* point_lin2_b;   /* [64,32],   [64]  */
  const float* vox_lin1_w;   const float* vox_lin1_b;     /* [64,64],   [64]  */
  const float* point_lin3_w; const float* point_lin3_b;   /* [128,128], [128] */
  const float* point_lin4_w; const float* point_lin4_b;   /* [128,128], [128] */
  const float* vox_lin2_w;   const float* vox_lin2_b;     /* [128,128], [128] */
} LidfPointNet;

size_t lidf_pointnet_workspace_bytes(int64_t N, int64_t V);
/* inp_feat [N,6] (relative xyz | rgb), vox2point_idx [N] int64 in [0,V) -> occ_voxel_feat [V,128]; voxels that own no
 * point get relu(bias) of the two voxel layers applied to zeros, exactly as torch_scatter leaves their rows at 0 */
int lidf_pointnet_forward(const LidfPointNet* w, const float* inp_feat, const int64_t* vox2point_idx, int64_t N, int64_t V,
                          float* occ_voxel_feat, void* workspace, size_t workspace_bytes, lidf_stream_t stream);

/* Same call with the engine of the two 128 -> 128 per-point layers (97 % of the MACs) chosen explicitly:
 * LIDF_MLP_AUTO / LIDF_MLP_TC_BF16X3 = tcgen05 tensor cores, split-bf16 operands with fp32 accumulation (the decoder's
 * arithmetic, ~2^-16 relative), LIDF_MLP_SIMT_FP32 = fp32 FMA kernels.  lidf_pointnet_forward is the AUTO form. */
int lidf_pointnet_forward_impl(const LidfPointNet* w, const float* inp_feat, const int64_t* vox2point_idx, int64_t N, int64_t V,
                               float* occ_voxel_feat, void* workspace, size_t workspace_bytes, int32_t mlp_impl,
                               lidf_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* LIDF_POINTNET_H_ */
