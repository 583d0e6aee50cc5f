/* lidf_query.h -- C ABI of the B200-native LIDF per-query-point decoder ("lidf_query").
 *
 * This is the drop-in boundary for ONE hot path of NVlabs/implicit_depth: everything
 * LIDF.get_embedding + LIDF.get_pred do after the ResNet / PointNet feature producers have run
 * (reference src/models/pipeline.py:338-466), plus the RefineNet decoder tail
 * (src/models/pipeline.py:1018-1029).  The reference has no native implementation of this path --
 * it is a chain of stock PyTorch / torchvision / torch_scatter ops -- so there is no existing FFI to
 * mirror symbol-for-symbol.  The entry points below follow the conventions of the reference's own
 * native extensions (src/extensions/ray_aabb/ray_aabb_cuda.cpp:20-37,
 * src/extensions/pcl_aabb/pcl_aabb_cuda.cpp:20-37): forward-only, inputs borrowed and read-only,
 * contiguous row-major device arrays, int/float scalars by value; outputs are written into
 * caller-owned device buffers (the Python wrapper allocates them with torch, like the reference's
 * torch::zeros inside ray_aabb_cuda_forward).  Errors come back as negative codes (no exceptions
 * cross the C ABI); the wrapper raises RuntimeError, matching TORCH_CHECK.
 *
 * Plain C: pointers, sizes and a stream handle.  No torch types.  All pointers are DEVICE pointers
 * unless a name ends in _host.  All float arrays are fp32, index arrays int64 (what torch.nonzero
 * hands the reference, pipeline.py:283-285).
 */
#ifndef LIDF_QUERY_H_
#define LIDF_QUERY_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LIDF_QUERY_ABI_VERSION 4

/* fixed by the shipped YAMLs (train_lidf.yaml:36-57): rgb_out 32 x roi_out_bbox 2^2, pnet_out 128, imnet_gf 64 */
#define LIDF_RGB_CH 32
#define LIDF_ROI_OUT 2
#define LIDF_RGB_DIM 128
#define LIDF_VOX_DIM 128
#define LIDF_H1 256
#define LIDF_H2 128
#define LIDF_H3 64
#define LIDF_IEF_ENC 16

typedef struct CUstream_st* lidf_stream_t; /* == cudaStream_t */

enum { LIDF_OK = 0, LIDF_ERR_NULL = -1, LIDF_ERR_UNSUPPORTED = -2, LIDF_ERR_WORKSPACE = -3,
       LIDF_ERR_CUDA = -4, LIDF_ERR_ARG = -5, LIDF_ERR_NO_SM100 = -6 };

enum { LIDF_DEC_IMNET = 0, LIDF_DEC_IEF = 1 };

/* which MLP engine runs the decoders */
enum { LIDF_MLP_AUTO = 0,      /* tcgen05 bf16x3 (split-precision, ~1e-5 rel.) on sm_100          */
       LIDF_MLP_SIMT_FP32 = 1, /* fp32 FFMA kernel: bit-level close to the reference, slow       */
       LIDF_MLP_TC_BF16X3 = 2, /* tcgen05, 3 bf16 products per MAC (hi*hi + hi*lo + lo*hi)        */
       LIDF_MLP_TC_BF16X1 = 3  /* tcgen05, 1 bf16 product per MAC: fast, ~3e-3 rel. -- NOT parity  */ };

/* One decoder = reference IMNet (src/models/implicit_net.py:60-98) or IEF (:100-152).
 * Weights are the PyTorch nn.Linear tensors as they sit in the state_dict: weight [out,in] row-major. */
typedef struct LidfDecoder {
  int32_t kind;            /* LIDF_DEC_IMNET | LIDF_DEC_IEF */
  int32_t n_iter;          /* IEF iterations (implicit_net.py:133); ignored for IMNet */
  int32_t inp_dim;         /* D (385 for LIDF, 334 for RefineNet); linear_1.weight is [256, D (+16 if IEF)] */
  int32_t use_sigmoid;     /* implicit_net.py:93-96 */
  float init_offset;       /* IEF.init_offset = 0.001 (implicit_net.py:104) */
  const float* w1; const float* b1;   /* linear_1 */
  const float* w2; const float* b2;   /* linear_2 [128,256] */
  const float* w3; const float* b3;   /* linear_3 [64,128] */
  const float* w4; const float* b4;   /* linear_4 [1,64] */
  const float* w_enc; const float* b_enc; /* IEF offset_enc [16,1],[16]; NULL for IMNet */
} LidfDecoder;

/* LIDF.get_embedding + LIDF.get_pred (pipeline.py:338-466) */
typedef struct LidfQueryParams {
  int64_t P;               /* ray-voxel intersection pairs (query points) */
  int64_t R;               /* miss rays  (total_miss_sample_num) */
  int64_t V;               /* occupied voxels in the batch */
  int32_t B, H, W;         /* feature-map batch / height / width */
  /* producers' outputs (not on the path): */
  const float* full_rgb_feat;   /* [B,32,H,W]  resnet_model(rgb_img), pipeline.py:370 */
  const float* occ_voxel_feat;  /* [V,128]     pnet_model(...),       pipeline.py:407 */
  /* rays: */
  const float* miss_ray_dir;    /* [R,3] unit */
  const int64_t* miss_img_ind;  /* [R,2] (x,y) */
  const int64_t* miss_bid;      /* [R] image id */
  const float* voxel_bound;     /* [V,6] min xyz, max xyz */
  /* pair list (occ_vox_intersect_idx, miss_ray_intersect_idx), any order; the reference's is voxel-major: */
  const int64_t* pair_vox;      /* [P] */
  const int64_t* pair_ray;      /* [P] */
  /* enter/leave distance per pair: EITHER pair_dist [P,2] OR the reference's dense dist [V,R,2]
   * (pipeline.py:345 looks it up as dist[vox,ray]) */
  const float* pair_dist;
  const float* dense_dist;
  const float* pcl_label_float; /* [P] or NULL: non-NULL selects the GT-label arg-max branch, pipeline.py:444-446 */
  /* opt.model.* / opt.grid.* scalars read on the path: */
  int32_t pos_encode;           /* model.pos_encode  */
  int32_t multires;             /* model.multires (8)       */
  int32_t multires_views;       /* model.multires_views (4) */
  int32_t intersect_pos_rel;    /* model.intersect_pos_type == 'rel' */
  int32_t roi_inp_bbox;         /* model.roi_inp_bbox (8); roi_out_bbox is fixed at 2 */
  float offset_range0, offset_range1;   /* grid.offset_range */
  float part_size;              /* data_dict['part_size'] (pipeline.py:170) */
  LidfDecoder offset_dec;       /* IMNet or IEF */
  LidfDecoder prob_dec;         /* IMNet */
  int32_t mlp_impl;             /* LIDF_MLP_* */
  /* outputs (data_dict entries, pipeline.py:460-466): */
  float* pred_offset;           /* [P]   offset_dec output before scaling (pipeline.py:434) */
  float* pred_prob_end;         /* [P]   prob_dec output (logit after the leaky clamp / sigmoid) */
  float* pair_pred_pos;         /* [P,3] */
  float* pred_prob_end_softmax; /* [P]   */
  int64_t* max_pair_id;         /* [R]   arg-max pair per ray; P for rays without a pair */
  float* pred_pos;              /* [R,3] */
  float* roi_feat_per_ray;      /* [R,128] optional (NULL to skip): ROIAlign feature, reused by RefineNet (pipeline.py:964) */
  void* workspace; size_t workspace_bytes;   /* >= lidf_query_workspace_bytes() */
  /* ABI 3, both optional (NULL to skip): */
  float* ief_iter_out;          /* [n_iter-1][P] IEF offset after iteration k < n_iter-1 (implicit_net.py:146), at the
                                 * original pair index: what lidf_query_backward needs to re-run an iteration */
  int32_t* index_error;         /* [1] device flag, set non-zero when a pair_ray / pair_vox / miss_bid value is out of
                                 * range.  Such entries are clamped into range (nothing is read or written out of
                                 * bounds) and the results for them are meaningless -- the reference trips a device-side
                                 * assert in the same situation. */
  void* weight_cache;           /* optional caller-owned device buffer, >= lidf_query_weight_cache_bytes(p): holds the   */
  size_t weight_cache_bytes;    /* packed decoder weights (k-major fp32 copies, bf16 hi|lo MMA streams, folded biases)   */
  int32_t weight_cache_valid;   /* between calls.  0: this call packs into it; 1: the ~20 packing launches are skipped   */
                                /* (the caller vouches that no decoder tensor changed since the call that filled it).    */
  /* ABI 4: */
  int32_t pairs_ray_major;      /* 1: the caller vouches that pair_ray is non-decreasing (the order lidf_ray_aabb_pairs_ray_major_*
                                 * emits: by ray, then voxel).  The voxel-major -> ray-major regroup (count / scan / scatter /
                                 * segment sort over all P pairs) is replaced by one binary search per ray; outputs are the same
                                 * tensors at the same pair indices.  A list that is not sorted sets bit 2 of *index_error. */
  int32_t winner_only_offset;   /* 1: "winner-only" mode for callers that read the per-ray results only.  Every reader of this
                                 * path's outputs in the reference (compute_loss pipeline.py:468-650, RefineNet :939-944) looks at
                                 * pred_pos, max_pair_id, pred_prob_end and pred_prob_end_softmax; pred_offset and pair_pred_pos
                                 * are stored in data_dict (:460-466) and never read again.  In this mode the probability
                                 * decoder runs over all P pairs, the rays are terminated, and the offset decoder runs on ONE
                                 * row per ray -- its arg-max pair -- writing pred_pos directly: the same bits as the full
                                 * call (a row's arithmetic does not depend on its tile), a third of the decoder work at 64
                                 * pairs per ray.  pred_offset / pair_pred_pos are NOT written (may be NULL); needs the tcgen05
                                 * engine.  Training: ief_iter_out is then [n_iter-1][R] (by ray) and pred_offset_ray [R]
                                 * receives the offset decoder's output of each ray's winner; lidf_query_backward, given the
                                 * same flag and those two arrays, runs the offset decoder's backward over the R winner rows
                                 * only (every other row's upstream gradient is exactly zero in the reference: pos_loss
                                 * reaches the decoder through pred_pos = pair_pred_pos[max_pair_id] alone, pipeline.py:449-454);
                                 * g_pred_offset / g_pair_pred_pos must then be NULL. */
  float* pred_offset_ray;       /* [R] optional (winner-only mode): forward output / backward input */
} LidfQueryParams;

/* Gradients of one decoder's parameters: fp32 device buffers with the shapes of the LidfDecoder tensors; every buffer is
 * OVERWRITTEN (not accumulated).  w_enc / b_enc are ignored for IMNet. */
typedef struct LidfDecoderGrad {
  float* w1; float* b1; float* w2; float* b2; float* w3; float* b3; float* w4; float* b4; float* w_enc; float* b_enc;
} LidfDecoderGrad;

/* Backward of lidf_query_forward: what torch autograd does for LIDF.get_embedding + LIDF.get_pred in training
 * (pipeline.py:338-466 under loss_net.backward(), src/trainers/train_lidf.py:394).  `fwd` is the forward call's parameter
 * block: its inputs are re-read and its outputs pred_offset, pred_prob_end, max_pair_id (and ief_iter_out, handed in as
 * `ief_iter`) are READ; fwd.workspace / roi_feat_per_ray / index_error are ignored.  Upstream gradients may be NULL
 * (= zero).  The soft-max that picks max_pair_id is detached in the reference (:442), so pred_prob_end_softmax and
 * max_pair_id carry no gradient.  Runs on the tcgen05 engine (split-bf16, fp32 accumulate) whatever fwd.mlp_impl says. */
typedef struct LidfQueryBackwardParams {
  LidfQueryParams fwd;
  const float* ief_iter;          /* [n_iter-1][P]; required when offset_dec is an IEF with n_iter > 1 */
  const float* g_pred_pos;        /* [R,3] */
  const float* g_pred_prob_end;   /* [P]   */
  const float* g_pred_offset;     /* [P]   */
  const float* g_pair_pred_pos;   /* [P,3] */
  float* g_full_rgb_feat;         /* [B,32,H,W] or NULL */
  float* g_occ_voxel_feat;        /* [V,128]    or NULL */
  LidfDecoderGrad g_offset_dec, g_prob_dec;
  int64_t chunk_rows;             /* pairs whose activations are resident at once (0 = auto, ~2 M); multiple of 128 */
  void* workspace; size_t workspace_bytes;   /* >= lidf_query_backward_workspace_bytes() */
} LidfQueryBackwardParams;

/* RefineNet.get_pred_refine decoder tail (pipeline.py:1018-1029): per RAY
 * x = [voxel_feat_end | rgb_feat_end | PE(pos [- centre]) | PE(dir)] -> offset_dec -> pred_pos + (o*(r1-r0)+r0)*dir */
typedef struct LidfRefineParams {
  int64_t R;
  const float* pred_pos;        /* [R,3] */
  const float* miss_ray_dir;    /* [R,3] */
  const float* end_voxel_center;/* [R,3] (only read when intersect_pos_rel) */
  const float* voxel_feat_end;  /* [R,128] occ_voxel_feat[end_voxel_id] */
  const float* rgb_feat_end;    /* [R,128] ROIAlign feature per ray */
  int32_t pos_encode, multires, multires_views, intersect_pos_rel;
  float offset_range0, offset_range1;   /* refine.offset_range */
  LidfDecoder offset_dec;
  int32_t mlp_impl;
  float* pred_pos_refine;       /* [R,3] */
  void* workspace; size_t workspace_bytes;
  /* Optional (ABI 2): the voxel features UN-gathered, as the reference holds them right before the gather
   * `occ_voxel_feat[end_voxel_id]` (pipeline.py:1016): occ_voxel_feat [V,128] + end_voxel_id [R] (+ voxel_bound [V,6], read
   * instead of end_voxel_center when intersect_pos_rel).  With these set, voxel_feat_end may be NULL and the decoder runs on
   * the tcgen05 engine (per-voxel layer-1 term + gather, as in lidf_query_forward) unless mlp_impl = LIDF_MLP_SIMT_FP32. */
  int64_t V;
  const float* occ_voxel_feat;
  const int64_t* end_voxel_id;
  const float* voxel_bound;
} LidfRefineParams;

int lidf_query_abi_version(void);
/* sizeof() of the parameter structs as compiled (0: LidfDecoder, 1: LidfQueryParams, 2: LidfRefineParams,
 * 3: LidfQueryBackwardParams) so an FFI
 * binding can verify its own struct layout */
size_t lidf_query_struct_size(int which);
const char* lidf_query_error_string(int code);
/* last CUDA error text recorded by this library on the calling thread ("" if none) */
const char* lidf_query_last_cuda_error(void);

size_t lidf_query_workspace_bytes(const LidfQueryParams* p);
size_t lidf_query_weight_cache_bytes(const LidfQueryParams* p);
int lidf_query_forward(const LidfQueryParams* p, lidf_stream_t stream);

size_t lidf_query_backward_workspace_bytes(const LidfQueryBackwardParams* p);
int lidf_query_backward(const LidfQueryBackwardParams* p, lidf_stream_t stream);
/* device time (ms) of the backward's tcgen05 section (k_mlp_bwd_tc + k_wgrad_pk_tc + k_wgrad_tc launches and the segment sums between them) in the most recent
 * lidf_query_backward on the calling thread; synchronises; < 0 if none */
float lidf_query_last_bwd_ms(void);

/* unit test of the wgrad kernel: C[M,N] = A^T B, A [rows,M] (M = 128 or 256), B [rows,N] (N % 16 == 0, <= 256), fp32
 * device arrays, C [M,N] row-major; scratch >= lidf_wgrad_selftest_scratch_bytes(M, N) */
size_t lidf_wgrad_selftest_scratch_bytes(int32_t M, int32_t N);
int lidf_wgrad_selftest(const float* A, const float* B, float* C, int64_t rows, int32_t M, int32_t N, void* scratch,
                        lidf_stream_t stream);
/* the same product through the packed hand-over path of the backward (k_pk_pack_rows -> k_wgrad_pk_tc: operands already
 * split into bf16 hi | lo and laid out for the MN-major UMMA descriptor, fetched by TMA bulk copies); M = 128 */
size_t lidf_wgrad_pk_selftest_scratch_bytes(int64_t rows, int32_t M, int32_t N);
/* byte offset of element (row, feature) of a [rows, F] tensor in that packed layout (lo != 0: the lo part); host
 * arithmetic only, < 0 on bad arguments */
int64_t lidf_pk_offset_bytes(int32_t F, int64_t row, int32_t feature, int32_t lo);
int lidf_wgrad_pk_selftest(const float* A, const float* B, float* C, int64_t rows, int32_t M, int32_t N, void* scratch,
                           lidf_stream_t stream);

size_t lidf_refine_workspace_bytes(const LidfRefineParams* p);
int lidf_refine_forward(const LidfRefineParams* p, lidf_stream_t stream);

/* stand-alone pieces (same kernels the fused call uses), exposed for tests and for callers that
 * already hold some of the intermediates: */
/* torchvision.ops.roi_align(feat, boxes(pix +- bbox/2 clamped), output_size=2, spatial_scale=1, aligned=True)
 * evaluated once per ray -> [R,128] in (c,ph,pw) order (pipeline.py:374-389) */
int lidf_roi_align_rays(const float* full_rgb_feat, int32_t B, int32_t H, int32_t W,
                        const int64_t* miss_img_ind, const int64_t* miss_bid, int64_t R,
                        int32_t roi_inp_bbox, float* roi_feat_per_ray, lidf_stream_t stream);
/* torch_scatter.scatter_softmax + scatter_max + pred_pos gather (pipeline.py:441-454) on an arbitrary-order pair list */
size_t lidf_ray_terminate_workspace_bytes(int64_t P, int64_t R);
int lidf_ray_terminate(const float* pred_prob_end, const int64_t* pair_ray, const float* pair_pred_pos,
                       const float* pcl_label_float, int64_t P, int64_t R,
                       float* pred_prob_end_softmax, int64_t* max_pair_id, float* pred_pos,
                       void* workspace, size_t workspace_bytes, lidf_stream_t stream);

/* The torch_scatter calls of LIDF.compute_loss that act on the path's outputs, keyed by ray (pipeline.py:482-486
 * scatter_log_softmax cross-entropy; :553-557 scatter_max labels and accuracy) plus the per-ray position errors (:472 L1,
 * :560-567 L2 over rays whose gt_pos is not all-zero).  pcl_label_float [P] holds 0/1 (pipeline.py:308-309).
 * Outputs: log_softmax [P], pred_label [R], gt_label [R] (int64, P for a ray without pairs), stats [6] (device, double):
 *   {sum of -log_softmax over labelled pairs, #labelled pairs, #rays with pred_label == gt_label,
 *    sum |pred_pos - gt_pos| over R*3, sum of masked L2 errors, #masked rays};
 * prob_loss = stats[0]/stats[1], acc = stats[2]/R, pos_loss = stats[3]/(3R), err = stats[4]/stats[5].
 * gt_pos / pred_pos may be NULL (stats[3..5] = 0). */
size_t lidf_ray_loss_workspace_bytes(int64_t P, int64_t R);
int lidf_ray_loss(const float* pred_prob_end, const float* pred_prob_end_softmax, const int64_t* pair_ray,
                  const float* pcl_label_float, int64_t P, int64_t R, const float* pred_pos, const float* gt_pos,
                  float* log_softmax, int64_t* pred_label, int64_t* gt_label, double* stats,
                  void* workspace, size_t workspace_bytes, lidf_stream_t stream);

/* The image-space terms of LIDF.compute_loss (pipeline.py:494-541; point_utils.gradient / get_surface_normal,
 * src/utils/point_utils.py:208-235): xyz_flat [B,H*W,3] (xyz_flat for 'train', xyz_corrupt_flat otherwise) with pred_pos /
 * gt_pos [R,3] scattered in at (miss_bid, miss_flat_img_id), forward-difference surface normals, and at the R miss pixels
 *   stats [6] (device, double) = {sum (1 - cos)/2, sum acos(clamp(cos)) [rad], sum |dx|^2, sum |dy|^2, 0, 0}
 * so that surf_norm_loss = stats[0]/R, angle_err = stats[1]/R * 180/pi, smooth_loss = (stats[2] + stats[3])/R.
 * Optional outputs (NULL to skip): the two full normal images [B,3,H,W] the reference stores in data_dict (:609-611). */
size_t lidf_image_loss_workspace_bytes(int32_t B, int32_t H, int32_t W, int64_t R);
int lidf_image_loss(const float* xyz_flat, const int64_t* miss_bid, const int64_t* miss_flat_img_id, const float* pred_pos,
                    const float* gt_pos, int32_t B, int32_t H, int32_t W, int64_t R, float* pred_surf_norm_img,
                    float* gt_surf_norm_img, double* stats, void* workspace, size_t workspace_bytes, lidf_stream_t stream);

/* Depth metrics of LIDF.compute_loss for exp_type != 'train' (pipeline.py:570-618).  stats[12] (device, double) =
 *   {n, #(thresh < 1.05), #(< 1.10), #(< 1.25), sum (gt-pred)^2, sum (log gt - log pred)^2,
 *    sum |log gt - log pred|, sum |gt-pred|/gt, sum |gt-pred|, sum (gt-pred)^2/gt, 0, 0}
 * so that a1 = s1/n, a2 = s2/n, a3 = s3/n, rmse = sqrt(s4/n), rmse_log = sqrt(s5/n), log10 = s6/n (natural log, as the
 * reference computes it), abs_rel = s7/n, mae = s8/n, sq_rel = s9/n.
 * _rays  (bs != 1, :571-575): rays whose gt_pos is not all-zero; depth = z of pred_pos / gt_pos [R,3].
 * _image (bs == 1, :576-603): the reference's host round trip -- gt depth (z of xyz_flat [H*W,3], nan/inf -> 0), corrupt_mask
 *   [H,W] (fp32 0/1) and the predicted depth image (z of xyz_corrupt_flat with pred_pos scattered in at miss_flat_img_id)
 *   resampled to 256x144 by cv2.resize(INTER_NEAREST) -- as one device-side nearest-neighbour pick; valid = gt > 0 and mask. */
size_t lidf_depth_metrics_workspace_bytes(int64_t n_rays, int32_t H, int32_t W);
int lidf_depth_metrics_rays(const float* pred_pos, const float* gt_pos, int64_t R, double* stats, void* workspace,
                            size_t workspace_bytes, lidf_stream_t stream);
int lidf_depth_metrics_image(const float* xyz_flat, const float* xyz_corrupt_flat, const float* corrupt_mask,
                             const int64_t* miss_flat_img_id, const float* pred_pos, int64_t R, int32_t H, int32_t W,
                             double* stats, void* workspace, size_t workspace_bytes, lidf_stream_t stream);

/* unit test of the tcgen05 primitives the decoder engine is built from (tcgen05.st operand staging, TMA weight chunks,
 * TS-mode tcgen05.mma with the 3-product bf16 split, tcgen05.ld): D[128,128] = A[128,32] * W[128,32]^T, fp32 device
 * arrays; scratch >= 16 KB device memory.  variant 0 = the layout the engine uses; 1 = LBO/SBO swapped (must be wrong). */
int lidf_tc_selftest(const float* A, const float* W, float* D, void* scratch, int32_t variant, lidf_stream_t stream);

/* device time (ms) of the dominant decoder kernel (k_mlp_tc / k_mlp_simt) in the most recent lidf_query_forward on
 * the calling thread, measured with CUDA events recorded around that launch on the caller's stream; synchronises on
 * the stop event.  < 0 if no decoder kernel has been launched. */
float lidf_query_last_mlp_ms(void);

/* number of kernels this library launched on the calling thread since the last reset (bench.py's gpu_launches) */
int64_t lidf_query_launch_count(int reset);

#ifdef __cplusplus
}
#endif
#endif /* LIDF_QUERY_H_ */
