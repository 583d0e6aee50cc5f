// lidf_query.cu -- C-ABI entry points (include/lidf_query.h) and host-side launch plan.
//
// Launch plan of lidf_query_forward (all on the caller's stream, no host sync, no allocation):
//   1. regroup  : k_count_pairs -> exclusive scan -> k_fill_perm -> k_sort_segments   (pairs -> CSR by ray)
//   2. k_roi_align_rays                       ROIAlign feature once per ray             (pipeline.py:374-389)
//   3. weight packing                         k-major fp32 copies / bf16 hi-lo stream images
//   4. k_rowprep (rays [, voxels])            per-ray / per-voxel part of linear_1
//   5. decoder kernel                         k_mlp_tc (tcgen05, default) or k_mlp_simt (fp32)
//   6. k_ray_terminate                        scatter_softmax + scatter_max + pred_pos  (pipeline.py:441-454)
#include <cuda_runtime.h>
#include <limits.h>
#include <math.h>
#include <stdio.h>
#include <string.h>

#include "lidf_common.cuh"
#include "lidf_prep.cuh"
#include "lidf_simt.cuh"
#include "lidf_tc.cuh"
#include "lidf_bwd.cuh"

static thread_local char g_cuda_err[256] = "";
static thread_local int64_t g_launches = 0;
// timing events of the decoder kernel: one pair per device ordinal (an event belongs to the device it was created on)
#define LIDF_MAX_DEVICES 64
static thread_local cudaEvent_t g_ev_mlp_dev[LIDF_MAX_DEVICES][2] = {};
static thread_local cudaEvent_t* g_ev_mlp = g_ev_mlp_dev[0];
static thread_local bool g_ev_valid = false;

static void mlp_event(int which, cudaStream_t st) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= LIDF_MAX_DEVICES) { cudaGetLastError(); return; }
  g_ev_mlp = g_ev_mlp_dev[dev];
  if (!g_ev_mlp[0] && (cudaEventCreate(&g_ev_mlp[0]) != cudaSuccess || cudaEventCreate(&g_ev_mlp[1]) != cudaSuccess)) {
    cudaGetLastError(); g_ev_mlp[0] = g_ev_mlp[1] = nullptr; return;
  }
  if (cudaEventRecord(g_ev_mlp[which], st) != cudaSuccess) { cudaGetLastError(); g_ev_valid = false; return; }   // e.g. inside a graph capture
  if (which == 1) g_ev_valid = true;
}

#define LIDF_LAUNCH_CHECK()                                                        \
  do {                                                                             \
    ++g_launches;                                                                  \
    cudaError_t e__ = cudaGetLastError();                                          \
    if (e__ != cudaSuccess) {                                                      \
      snprintf(g_cuda_err, sizeof(g_cuda_err), "%s:%d %s", __FILE__, __LINE__, cudaGetErrorString(e__)); \
      return LIDF_ERR_CUDA;                                                        \
    }                                                                              \
  } while (0)
#define LIDF_CUDA(call)                                                            \
  do {                                                                             \
    cudaError_t e__ = (call);                                                      \
    if (e__ != cudaSuccess) {                                                      \
      snprintf(g_cuda_err, sizeof(g_cuda_err), "%s:%d %s", __FILE__, __LINE__, cudaGetErrorString(e__)); \
      return LIDF_ERR_CUDA;                                                        \
    }                                                                              \
  } while (0)

namespace {

int persistent_blocks(int64_t items, int per_sm);   // grid of a persistent kernel: min(items, SM count x per_sm)

struct Bump {  // workspace carving, 256-byte aligned
  char* base; size_t off;
  template <typename T> T* take(size_t n) {
    off = (size_t)lidf_align_up((int64_t)off, 256);
    T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
    off += n * sizeof(T);
    return p;
  }
};

inline int pad16(int k) { return (k + 15) / 16 * 16; }

struct CsrBufs { int* cnt; int* ray_start; int* block_sums; int* perm; int nb; };

CsrBufs carve_csr(Bump& b, int64_t P, int64_t R) {
  CsrBufs c;
  const int64_t n = R + 1;
  c.nb = (int)((n + LIDF_SCAN_BLOCK * LIDF_SCAN_ITEMS - 1) / (LIDF_SCAN_BLOCK * LIDF_SCAN_ITEMS));
  c.cnt = b.take<int>(n + 1);          // +1: error flag lives at cnt[n]
  c.ray_start = b.take<int>(n);
  c.block_sums = b.take<int>(c.nb);
  c.perm = b.take<int>(P > 0 ? P : 1);
  return c;
}

int build_csr(const CsrBufs& c, const int64_t* pair_ray, int64_t P, int64_t R, cudaStream_t st, bool sort_segments = true,
              const int64_t* check_vox = nullptr, int64_t V = 0, bool ray_major = false) {
  const int64_t n = R + 1;
  if (ray_major) {                                   // caller's list is sorted by ray: no regroup, see k_ray_start_sorted
    LIDF_CUDA(cudaMemsetAsync(c.cnt + n, 0, sizeof(int), st));
    k_ray_start_sorted<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(pair_ray, P, R, c.ray_start);
    LIDF_LAUNCH_CHECK();
    if (P > 0) {
      k_check_sorted_pairs<<<(unsigned)((P + 255) / 256), 256, 0, st>>>(pair_ray, P, R, c.perm, c.cnt + n, check_vox, V);
      LIDF_LAUNCH_CHECK();
    }
    return LIDF_OK;
  }
  LIDF_CUDA(cudaMemsetAsync(c.cnt, 0, sizeof(int) * (n + 1), st));
  if (P > 0) {
    k_count_pairs<<<(unsigned)((P + 255) / 256), 256, 0, st>>>(pair_ray, P, R, c.cnt, c.cnt + n, check_vox, V);
    LIDF_LAUNCH_CHECK();
  }
  k_scan_partial<<<c.nb, LIDF_SCAN_BLOCK, 0, st>>>(c.cnt, n, c.ray_start, c.block_sums);
  LIDF_LAUNCH_CHECK();
  k_scan_blocksums<<<1, 1024, 0, st>>>(c.block_sums, c.nb);
  LIDF_LAUNCH_CHECK();
  k_scan_add<<<c.nb, LIDF_SCAN_BLOCK, 0, st>>>(c.ray_start, n, c.block_sums, nullptr);
  LIDF_LAUNCH_CHECK();
  if (P > 0) {
    LIDF_CUDA(cudaMemsetAsync(c.cnt, 0, sizeof(int) * n, st));
    k_fill_perm<<<(unsigned)((P + 255) / 256), 256, 0, st>>>(pair_ray, P, R, c.ray_start, c.cnt, c.perm);
    LIDF_LAUNCH_CHECK();
    if (sort_segments) {
      k_sort_segments<<<(unsigned)((R * 32 + 255) / 256), 256, 0, st>>>(c.ray_start, R, c.perm);
      LIDF_LAUNCH_CHECK();
    }
  }
  return LIDF_OK;
}

int check_decoder(const LidfDecoder& d, int D) {
  if (!d.w1 || !d.b1 || !d.w2 || !d.b2 || !d.w3 || !d.b3 || !d.w4 || !d.b4) return LIDF_ERR_NULL;
  if (d.kind != LIDF_DEC_IMNET && d.kind != LIDF_DEC_IEF) return LIDF_ERR_UNSUPPORTED;
  if (d.kind == LIDF_DEC_IEF && (!d.w_enc || !d.b_enc || d.n_iter < 1 || d.n_iter > 8)) return LIDF_ERR_ARG;
  if (d.inp_dim != D) return LIDF_ERR_UNSUPPORTED;
  return LIDF_OK;
}

// AUTO = the tcgen05 engine where its operand layout applies (NeRF encoding with multires 8, the shipped YAMLs), else the
// fp32 FMA engine, which takes any multires <= LIDF_MAX_MULTIRES and pos_encode off.  Both are CUDA kernels of this library;
// an explicitly requested engine is never substituted.
int resolve_impl(int mlp_impl, int* out, int pos_encode = 1, int multires = 8) {
  int impl = mlp_impl == LIDF_MLP_AUTO ? ((pos_encode && multires == 8) ? LIDF_MLP_TC_BF16X3 : LIDF_MLP_SIMT_FP32) : mlp_impl;
  if (impl != LIDF_MLP_SIMT_FP32 && impl != LIDF_MLP_TC_BF16X3 && impl != LIDF_MLP_TC_BF16X1) return LIDF_ERR_ARG;
  *out = impl;
  return LIDF_OK;
}

// ---- SIMT engine: packed fp32 weights --------------------------------------------------------
struct SimtPack {
  // shared row-prep matrices
  float* Wt_row; int KR; int Ntot;      // per-ray (or per refine-row) part of linear_1, [KR][Ntot]
  float* bias_row;                      // [Ntot]
  float* Wt_vox;                        // [128][Ntot] (LIDF only)
  float* Wt_pe[2]; float* Wt2[2]; float* Wt3[2]; float* u[2];
  int KP;
};

SimtPack carve_simt_pack(Bump& b, int n_dec, int KR, int KP, bool with_vox) {
  SimtPack s;
  s.KR = KR; s.KP = KP; s.Ntot = 256 * n_dec;
  s.Wt_row = b.take<float>((size_t)KR * s.Ntot);
  s.bias_row = b.take<float>(s.Ntot);
  s.Wt_vox = with_vox ? b.take<float>((size_t)128 * s.Ntot) : nullptr;
  for (int d = 0; d < 2; ++d) {
    if (d < n_dec) {
      s.Wt_pe[d] = b.take<float>((size_t)KP * 256);
      s.Wt2[d] = b.take<float>(256 * 128);
      s.Wt3[d] = b.take<float>(128 * 64);
      s.u[d] = b.take<float>(256);
    } else { s.Wt_pe[d] = s.Wt2[d] = s.Wt3[d] = s.u[d] = nullptr; }
  }
  return s;
}

int pack_wt(const float* w, int ldw, int col0, int K, int N, float* dst, int ldd, int k_off, int n_off, cudaStream_t st) {
  if (K <= 0) return LIDF_OK;
  k_pack_wt<<<(K * N + 255) / 256, 256, 0, st>>>(w, ldw, col0, K, N, dst, ldd, k_off, n_off);
  LIDF_LAUNCH_CHECK();
  return LIDF_OK;
}

}  // namespace

// =================================================================================================
extern "C" int lidf_query_abi_version(void) { return LIDF_QUERY_ABI_VERSION; }
extern "C" size_t lidf_query_struct_size(int which) {
  return which == 0 ? sizeof(LidfDecoder) : which == 1 ? sizeof(LidfQueryParams) : which == 2 ? sizeof(LidfRefineParams)
         : which == 3 ? sizeof(LidfQueryBackwardParams) : 0;
}

extern "C" const char* lidf_query_error_string(int code) {
  switch (code) {
    case LIDF_OK: return "ok";
    case LIDF_ERR_NULL: return "null pointer argument";
    case LIDF_ERR_UNSUPPORTED: return "unsupported configuration (decoder dims / kinds outside the shipped YAML family)";
    case LIDF_ERR_WORKSPACE: return "workspace too small";
    case LIDF_ERR_CUDA: return "CUDA error (see lidf_query_last_cuda_error)";
    case LIDF_ERR_ARG: return "invalid argument";
    case LIDF_ERR_NO_SM100: return "tcgen05 engine needs an sm_100 device";
    default: return "unknown error";
  }
}
extern "C" const char* lidf_query_last_cuda_error(void) { return g_cuda_err; }
extern "C" int lidf_tc_selftest(const float* A, const float* W, float* D, void* scratch, int32_t variant, lidf_stream_t stream) {
  if (!A || !W || !D || !scratch) return LIDF_ERR_NULL;
  return tc_selftest(A, W, D, (uint8_t*)scratch, variant, stream, &g_launches, g_cuda_err, sizeof(g_cuda_err));
}
extern "C" float lidf_query_last_mlp_ms(void) {
  if (!g_ev_valid) return -1.f;
  float ms = -1.f;
  if (cudaEventSynchronize(g_ev_mlp[1]) != cudaSuccess) return -1.f;
  if (cudaEventElapsedTime(&ms, g_ev_mlp[0], g_ev_mlp[1]) != cudaSuccess) return -1.f;
  return ms;
}
extern "C" int64_t lidf_query_launch_count(int reset) {
  int64_t v = g_launches;
  if (reset) g_launches = 0;
  return v;
}

// -------------------------------------------------------------------------------------------------
extern "C" int lidf_roi_align_rays(const float* feat, int32_t B, int32_t H, int32_t W, const int64_t* img_ind,
                                   const int64_t* bid, int64_t R, int32_t roi_inp_bbox, float* out, lidf_stream_t stream) {
  if (!feat || !img_ind || !bid || !out) return LIDF_ERR_NULL;
  if (R <= 0) return LIDF_OK;
  k_roi_align_rays<<<(unsigned)((R + 31) / 32), LIDF_ROI_THREADS, 0, stream>>>(feat, nullptr, B, H, W, img_ind, bid, R,
                                                                                roi_inp_bbox / 2, out, nullptr, nullptr);
  LIDF_LAUNCH_CHECK();
  return LIDF_OK;
}

namespace {
// 3-D tensor map (W, H, B * 32) over full_rgb_feat [B,32,H,W] fp32 for k_box4_tma.  cuTensorMapEncodeTiled is a driver entry
// point; it is looked up through the runtime so that the library keeps linking against libcudart only.
bool make_feat_tensor_map(CUtensorMap* tmap, const float* feat, int B, int H, int W) {
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static thread_local EncodeFn encode = nullptr;
  static thread_local bool looked_up = false;
  if (!looked_up) {
    looked_up = true;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      encode = reinterpret_cast<EncodeFn>(fn);
    else
      cudaGetLastError();
  }
  if (!encode || ((uintptr_t)feat & 15) != 0) return false;
  const cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B * LIDF_RGB_CH};
  const cuuint64_t strides[2] = {(cuuint64_t)W * 4, (cuuint64_t)W * H * 4};          // bytes, dims 1 and 2
  const cuuint32_t box[3] = {BOX_SW, BOX_SH, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  return encode(tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(feat), dims, strides, box, estr,
                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// ROIAlign per ray through the 4x4 box-sum map (dense ray sets: the map costs a quarter of a ray per pixel)
int roi_align_rays_box(const float* feat, float* box, int* border_list, int* border_count, int B, int H, int W,
                       const int64_t* img_ind, const int64_t* bid, int64_t R, int roi_inp_bbox, float* out, cudaStream_t st) {
  const int64_t n = (int64_t)B * LIDF_RGB_CH * H * W;
  CUtensorMap tmap;
  if (W % 4 == 0 && (int64_t)B * LIDF_RGB_CH <= 65535 && make_feat_tensor_map(&tmap, feat, B, H, W)) {
    // feature tiles staged by TMA (cp.async.bulk.tensor, one 68 x 35 tile per CTA)
    k_box4_tma<<<dim3((W + BOX_TW - 1) / BOX_TW, (H + BOX_TH - 1) / BOX_TH, B * LIDF_RGB_CH), 256, 0, st>>>(tmap, H, W, box);
  } else {
    k_box4<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(feat, n, H, W, box);
  }
  LIDF_LAUNCH_CHECK();
  LIDF_CUDA(cudaMemsetAsync(border_count, 0, sizeof(int), st));
  k_roi_align_rays<<<(unsigned)((R + 31) / 32), LIDF_ROI_THREADS, 0, st>>>(feat, box, B, H, W, img_ind, bid, R,
                                                                            roi_inp_bbox / 2, out, border_list, border_count);
  LIDF_LAUNCH_CHECK();
  k_roi_align_border<<<persistent_blocks((R + 31) / 32, 8), LIDF_ROI_THREADS, 0, st>>>(feat, B, H, W, img_ind, bid,
                                                                                        roi_inp_bbox / 2, out, border_list,
                                                                                        border_count);
  LIDF_LAUNCH_CHECK();
  return LIDF_OK;
}
}  // namespace

extern "C" size_t lidf_ray_terminate_workspace_bytes(int64_t P, int64_t R) {
  Bump b{nullptr, 0};
  carve_csr(b, P, R);
  return b.off + 256;
}

extern "C" int lidf_ray_terminate(const float* logit, const int64_t* pair_ray, const float* pair_pred_pos,
                                  const float* label, int64_t P, int64_t R, float* soft, int64_t* max_pair_id,
                                  float* pred_pos, void* ws, size_t ws_bytes, lidf_stream_t stream) {
  if (!max_pair_id || !pred_pos || !ws) return LIDF_ERR_NULL;
  if (P > 0 && (!logit || !pair_ray || !pair_pred_pos || !soft)) return LIDF_ERR_NULL;
  if (P >= INT_MAX || R >= INT_MAX / 32) return LIDF_ERR_UNSUPPORTED;
  if (ws_bytes < lidf_ray_terminate_workspace_bytes(P, R)) return LIDF_ERR_WORKSPACE;
  if (R <= 0) return LIDF_OK;
  Bump b{(char*)ws, 0};
  CsrBufs c = carve_csr(b, P, R);
  int rc = build_csr(c, pair_ray, P, R, stream);
  if (rc) return rc;
  k_ray_terminate<<<(unsigned)(((R + 3) / 4 * 32 + 255) / 256), 256, 0, stream>>>(logit, pair_pred_pos, label, c.ray_start, c.perm,
                                                                        P, R, soft, max_pair_id, pred_pos);
  LIDF_LAUNCH_CHECK();
  return LIDF_OK;
}

#define LIDF_LOSS_BLOCKS 592      // 4 x 148: blocks of the fixed-order reduction
extern "C" size_t lidf_ray_loss_workspace_bytes(int64_t P, int64_t R) {
  Bump b{nullptr, 0};
  carve_csr(b, P, R);
  b.take<float>((size_t)(R > 0 ? R : 1) * LIDF_LOSS_NSTAT);
  b.take<double>((size_t)LIDF_LOSS_BLOCKS * LIDF_LOSS_NSTAT);
  b.take<unsigned>(1);
  return b.off + 256;
}

extern "C" int lidf_ray_loss(const float* logit, const float* soft, const int64_t* pair_ray, const float* label, int64_t P,
                             int64_t R, const float* pred_pos, const float* gt_pos, float* log_softmax, int64_t* pred_label,
                             int64_t* gt_label, double* stats, void* ws, size_t ws_bytes, lidf_stream_t stream) {
  if (!pred_label || !gt_label || !stats || !ws) return LIDF_ERR_NULL;
  if (P > 0 && (!logit || !soft || !pair_ray || !label || !log_softmax)) return LIDF_ERR_NULL;
  if ((gt_pos != nullptr) != (pred_pos != nullptr)) return LIDF_ERR_ARG;
  if (P < 0 || R < 0) return LIDF_ERR_ARG;
  if (P >= INT_MAX || R >= INT_MAX / 32) return LIDF_ERR_UNSUPPORTED;
  if (ws_bytes < lidf_ray_loss_workspace_bytes(P, R)) return LIDF_ERR_WORKSPACE;
  cudaStream_t st = stream;
  LIDF_CUDA(cudaMemsetAsync(stats, 0, sizeof(double) * LIDF_LOSS_NSTAT, st));
  if (R == 0) return LIDF_OK;
  Bump b{(char*)ws, 0};
  CsrBufs c = carve_csr(b, P, R);
  float* ray_part = b.take<float>((size_t)R * LIDF_LOSS_NSTAT);
  double* block_part = b.take<double>((size_t)LIDF_LOSS_BLOCKS * LIDF_LOSS_NSTAT);
  unsigned* done = b.take<unsigned>(1);
  int rc = build_csr(c, pair_ray, P, R, st);
  if (rc) return rc;
  LIDF_CUDA(cudaMemsetAsync(done, 0, sizeof(unsigned), st));
  k_ray_loss<<<(unsigned)((R * 32 + 255) / 256), 256, 0, st>>>(logit, soft, label, c.ray_start, c.perm, P, R, pred_pos, gt_pos,
                                                                log_softmax, pred_label, gt_label, ray_part);
  LIDF_LAUNCH_CHECK();
  const int nb = (int)(R < LIDF_LOSS_BLOCKS ? R : LIDF_LOSS_BLOCKS);
  k_ray_loss_reduce<<<nb, 256, 0, st>>>(ray_part, R, block_part, done, stats);
  LIDF_LAUNCH_CHECK();
  return LIDF_OK;
}

extern "C" size_t lidf_image_loss_workspace_bytes(int32_t B, int32_t H, int32_t W, int64_t R) {
  if (B <= 0 || H <= 0 || W <= 0 || R < 0) return 0;
  Bump b{nullptr, 0};
  b.take<float>((size_t)B * H * W * 3); b.take<float>((size_t)B * H * W * 3);
  b.take<float>((size_t)(R > 0 ? R : 1) * LIDF_LOSS_NSTAT);
  b.take<double>((size_t)LIDF_LOSS_BLOCKS * LIDF_LOSS_NSTAT);
  b.take<unsigned>(1);
  return b.off + 256;
}

extern "C" int lidf_image_loss(const float* xyz_flat, const int64_t* bid, const int64_t* flat, const float* pred_pos,
                               const float* gt_pos, int32_t B, int32_t H, int32_t W, int64_t R, float* pred_img, float* gt_img,
                               double* stats, void* ws, size_t ws_bytes, lidf_stream_t stream) {
  if (!xyz_flat || !stats || !ws) return LIDF_ERR_NULL;
  if (B <= 0 || H <= 0 || W <= 0 || R < 0) return LIDF_ERR_ARG;
  if (R > 0 && (!bid || !flat || !pred_pos || !gt_pos)) return LIDF_ERR_NULL;
  if (ws_bytes < lidf_image_loss_workspace_bytes(B, H, W, R)) return LIDF_ERR_WORKSPACE;
  cudaStream_t st = stream;
  const size_t npix = (size_t)B * H * W;
  Bump b{(char*)ws, 0};
  float* pred_pcl = b.take<float>(npix * 3);
  float* gt_pcl = b.take<float>(npix * 3);
  float* ray_part = b.take<float>((size_t)(R > 0 ? R : 1) * LIDF_LOSS_NSTAT);
  double* block_part = b.take<double>((size_t)LIDF_LOSS_BLOCKS * LIDF_LOSS_NSTAT);
  unsigned* done = b.take<unsigned>(1);
  LIDF_CUDA(cudaMemsetAsync(stats, 0, sizeof(double) * LIDF_LOSS_NSTAT, st));
  LIDF_CUDA(cudaMemcpyAsync(pred_pcl, xyz_flat, sizeof(float) * npix * 3, cudaMemcpyDeviceToDevice, st));   // .clone(), :496-500
  LIDF_CUDA(cudaMemcpyAsync(gt_pcl, xyz_flat, sizeof(float) * npix * 3, cudaMemcpyDeviceToDevice, st));
  if (R > 0) {
    k_img_scatter<<<(unsigned)((R + 255) / 256), 256, 0, st>>>(pred_pos, gt_pos, bid, flat, R, (int64_t)H * W, pred_pcl, gt_pcl);
    LIDF_LAUNCH_CHECK();
    k_img_loss<<<(unsigned)((R + 255) / 256), 256, 0, st>>>(pred_pcl, gt_pcl, bid, flat, R, H, W, ray_part);
    LIDF_LAUNCH_CHECK();
    LIDF_CUDA(cudaMemsetAsync(done, 0, sizeof(unsigned), st));
    const int nb = (int)(R < LIDF_LOSS_BLOCKS ? R : LIDF_LOSS_BLOCKS);
    k_ray_loss_reduce<<<nb, 256, 0, st>>>(ray_part, R, block_part, done, stats);
    LIDF_LAUNCH_CHECK();
  }
  if (pred_img) { k_img_normals<<<(unsigned)((npix + 255) / 256), 256, 0, st>>>(pred_pcl, B, H, W, pred_img); LIDF_LAUNCH_CHECK(); }
  if (gt_img) { k_img_normals<<<(unsigned)((npix + 255) / 256), 256, 0, st>>>(gt_pcl, B, H, W, gt_img); LIDF_LAUNCH_CHECK(); }
  return LIDF_OK;
}

// -------------------------------------------------------------------------------------------------
// depth metrics (pipeline.py:570-618): per-element terms -> two fixed-order reductions -> stats[12]
namespace {
struct MetricPlan { float* partA; float* partB; double* blockA; double* blockB; unsigned* done; float* predz; size_t bytes; };
MetricPlan plan_metrics(int64_t n, int64_t HW, char* base) {
  Bump b{base, 0};
  MetricPlan m;
  m.partA = b.take<float>((size_t)(n > 0 ? n : 1) * 6); m.partB = b.take<float>((size_t)(n > 0 ? n : 1) * 6);
  m.blockA = b.take<double>((size_t)LIDF_LOSS_BLOCKS * 6); m.blockB = b.take<double>((size_t)LIDF_LOSS_BLOCKS * 6);
  m.done = b.take<unsigned>(2);
  m.predz = b.take<float>((size_t)(HW > 0 ? HW : 1));
  m.bytes = b.off + 256;
  return m;
}
int reduce_metrics(const MetricPlan& m, int64_t n, double* stats, cudaStream_t st) {
  LIDF_CUDA(cudaMemsetAsync(m.done, 0, 2 * sizeof(unsigned), st));
  const int nb = (int)(n < LIDF_LOSS_BLOCKS ? n : LIDF_LOSS_BLOCKS);
  k_ray_loss_reduce<<<nb, 256, 0, st>>>(m.partA, n, m.blockA, m.done, stats);
  LIDF_LAUNCH_CHECK();
  k_ray_loss_reduce<<<nb, 256, 0, st>>>(m.partB, n, m.blockB, m.done + 1, stats + 6);
  LIDF_LAUNCH_CHECK();
  return LIDF_OK;
}
}  // namespace

extern "C" size_t lidf_depth_metrics_workspace_bytes(int64_t n_rays, int32_t H, int32_t W) {
  const int64_t n = n_rays > (int64_t)LIDF_METRIC_W * LIDF_METRIC_H ? n_rays : (int64_t)LIDF_METRIC_W * LIDF_METRIC_H;
  return plan_metrics(n, (int64_t)H * W, nullptr).bytes;
}
extern "C" int lidf_depth_metrics_rays(const float* pred_pos, const float* gt_pos, int64_t R, double* stats, void* ws,
                                       size_t ws_bytes, lidf_stream_t stream) {
  if (!stats || !ws) return LIDF_ERR_NULL;
  cudaStream_t st = stream;
  LIDF_CUDA(cudaMemsetAsync(stats, 0, 12 * sizeof(double), st));
  if (R <= 0) return LIDF_OK;
  if (!pred_pos || !gt_pos) return LIDF_ERR_NULL;
  MetricPlan m = plan_metrics(R, 0, (char*)ws);
  if (ws_bytes < m.bytes) return LIDF_ERR_WORKSPACE;
  k_depth_terms_rays<<<(unsigned)((R + 255) / 256), 256, 0, st>>>(pred_pos, gt_pos, R, m.partA, m.partB);
  LIDF_LAUNCH_CHECK();
  return reduce_metrics(m, R, stats, st);
}
extern "C" int lidf_depth_metrics_image(const float* xyz_flat, const float* xyz_corrupt_flat, const float* corrupt_mask,
                                        const int64_t* miss_flat_img_id, const float* pred_pos, int64_t R, int32_t H, int32_t W,
                                        double* stats, void* ws, size_t ws_bytes, lidf_stream_t stream) {
  if (!xyz_flat || !xyz_corrupt_flat || !corrupt_mask || !stats || !ws) return LIDF_ERR_NULL;
  if (H <= 0 || W <= 0 || R < 0 || (R > 0 && (!miss_flat_img_id || !pred_pos))) return LIDF_ERR_ARG;
  cudaStream_t st = stream;
  const int64_t HW = (int64_t)H * W, n = (int64_t)LIDF_METRIC_W * LIDF_METRIC_H;
  MetricPlan m = plan_metrics(n, HW, (char*)ws);
  if (ws_bytes < m.bytes) return LIDF_ERR_WORKSPACE;
  k_depth_pred_image<<<(unsigned)((HW + 255) / 256), 256, 0, st>>>(xyz_corrupt_flat, HW, m.predz);
  LIDF_LAUNCH_CHECK();
  if (R > 0) {
    k_depth_pred_scatter<<<(unsigned)((R + 255) / 256), 256, 0, st>>>(pred_pos, miss_flat_img_id, R, HW, m.predz);
    LIDF_LAUNCH_CHECK();
  }
  k_depth_terms_image<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(xyz_flat, corrupt_mask, m.predz, H, W, m.partA, m.partB);
  LIDF_LAUNCH_CHECK();
  return reduce_metrics(m, n, stats, st);
}

// -------------------------------------------------------------------------------------------------
namespace {

struct QueryPlan {
  int impl, pe_pos, pe_dir, D, KR, KP;
  CsrBufs csr;
  float* roi_feat; float* T; float* Av; float* box4; int* border_list; int* border_count;
  int* win; float* off_ray;                 // winner-only mode: each ray's arg-max pair (-1: none), its offset output
  int* live_list; int* live_count;          // sparse regime: rays that own a pair (row prep works on these only)
  bool roi_sparse;                          // ... and ROIAlign too (only when the per-ray ROI feature is not an output)
  SimtPack sp;
  TcBufs tc;
  size_t bytes;                             // workspace
  size_t cache_bytes;                       // packed-weight region (inside the workspace unless p->weight_cache is given)
  bool packed;                              // the packed weights in the cache are valid: skip the packing launches
};

int plan_query(const LidfQueryParams* p, QueryPlan* q, char* base, bool sizing_cache_only = false, bool allow_sparse = true) {
  int rc = resolve_impl(p->mlp_impl, &q->impl, p->pos_encode, p->multires);
  if (rc) return rc;
  if (p->multires < 0 || p->multires > LIDF_MAX_MULTIRES || p->multires_views < 0 || p->multires_views > LIDF_MAX_MULTIRES)
    return LIDF_ERR_UNSUPPORTED;
  q->pe_pos = lidf_pe_dim(p->multires, p->pos_encode);
  q->pe_dir = lidf_pe_dim(p->multires_views, p->pos_encode);
  q->D = LIDF_VOX_DIM + LIDF_RGB_DIM + 2 * q->pe_pos + q->pe_dir;     // pipeline.py:64-65
  q->KR = pad16(LIDF_RGB_DIM + q->pe_dir);
  q->KP = pad16(2 * q->pe_pos);
  Bump b{base, 0};
  q->csr = carve_csr(b, p->P, p->R);
  q->roi_feat = p->roi_feat_per_ray ? p->roi_feat_per_ray : b.take<float>((size_t)p->R * LIDF_RGB_DIM);
  // dense ray sets (at least a third of the pixels) go through the 4x4 box-sum map
  const bool use_box = p->roi_inp_bbox == 8 && p->H >= 9 && p->W >= 9 && (int64_t)p->R * 3 >= (int64_t)p->B * p->H * p->W;
  q->box4 = use_box ? b.take<float>((size_t)p->B * LIDF_RGB_CH * p->H * p->W) : nullptr;   // (nullptr in the sizing pass too)
  q->border_list = use_box ? b.take<int>((size_t)(p->R > 0 ? p->R : 1)) : nullptr;
  q->border_count = use_box ? b.take<int>(1) : nullptr;
  q->T = b.take<float>((size_t)p->R * 512);
  q->Av = b.take<float>((size_t)p->V * 512);
  q->win = p->winner_only_offset ? b.take<int>((size_t)(p->R > 0 ? p->R : 1)) : nullptr;
  q->off_ray = p->winner_only_offset ? b.take<float>((size_t)(p->R > 0 ? p->R : 1)) : nullptr;
  // sparse regime (fewer than 8 pairs per ray on average): per-ray work only for rays
  // that own a pair.  The tensor-core row prep (T) always can; ROIAlign only if its output is not handed to the caller
  // (RefineNet reads the feature of every ray) and the rays do not go through the box-sum map anyway.
  const bool sparse = allow_sparse && p->P < 8 * p->R;
  q->live_list = sparse ? b.take<int>((size_t)(p->R > 0 ? p->R : 1)) : nullptr;
  q->live_count = sparse ? b.take<int>(1) : nullptr;
  q->roi_sparse = sparse && !p->roi_feat_per_ray && !use_box;
  if (q->impl != LIDF_MLP_SIMT_FP32 && q->KP > TC_KPE_MAX) return LIDF_ERR_UNSUPPORTED;
  // packed weights: in the caller's cache buffer when given, else at the end of the workspace
  const bool ext = p->weight_cache != nullptr && base != nullptr;
  Bump c{ext ? (char*)p->weight_cache : base, ext || sizing_cache_only ? 0 : b.off};
  if (sizing_cache_only) c.base = nullptr;
  q->sp = carve_simt_pack(c, 2, q->KR, q->KP, true);      // row-prep matrices are shared by both engines
  if (q->impl != LIDF_MLP_SIMT_FP32) q->tc = carve_tc(c, p->V, 2);
  q->cache_bytes = (ext || sizing_cache_only ? c.off : c.off - b.off) + 256;
  q->packed = ext && p->weight_cache_valid != 0;
  if (ext && p->weight_cache_bytes < q->cache_bytes) return LIDF_ERR_WORKSPACE;
  q->bytes = (ext ? b.off : c.off) + 256;
  return LIDF_OK;
}

}  // namespace


namespace {
// steps 1-4 of the launch plan (shared by the forward and the backward, which recomputes them instead of keeping the
// forward's workspace alive): pair regroup + index range check, ROIAlign per ray, weight packing, row prep (T, A_v)
int run_prep(const LidfQueryParams* p, QueryPlan& q, cudaStream_t st) {
  int rc;
  const int64_t P = p->P, R = p->R, V = p->V;
  // 1. regroup (flags out-of-range pair_ray in csr.cnt[R + 1]); range check of pair_vox / miss_bid into the same flag
  if ((rc = build_csr(q.csr, p->pair_ray, P, R, st, true, p->pair_vox, V, p->pairs_ray_major != 0))) return rc;
  if (R > 0) {
    k_validate_indices<<<(unsigned)((R + 255) / 256), 256, 0, st>>>(nullptr, 0, V, p->miss_bid, R, p->B, q.csr.cnt + (R + 1));
    LIDF_LAUNCH_CHECK();
  }
  // 2. ROIAlign per ray
  if (q.live_list) {
    LIDF_CUDA(cudaMemsetAsync(q.live_count, 0, sizeof(int), st));
    k_live_rays<<<(unsigned)((R + 255) / 256), 256, 0, st>>>(q.csr.ray_start, R, q.live_list, q.live_count);
    LIDF_LAUNCH_CHECK();
  }
  if (q.box4) {
    if ((rc = roi_align_rays_box(p->full_rgb_feat, q.box4, q.border_list, q.border_count, p->B, p->H, p->W, p->miss_img_ind,
                                 p->miss_bid, R, p->roi_inp_bbox, q.roi_feat, st))) return rc;
  } else if (q.roi_sparse && R > 0) {
    k_roi_align_rays<<<(unsigned)((R + 31) / 32), LIDF_ROI_THREADS, 0, st>>>(p->full_rgb_feat, nullptr, p->B, p->H, p->W,
                                                                              p->miss_img_ind, p->miss_bid, R, p->roi_inp_bbox / 2,
                                                                              q.roi_feat, nullptr, nullptr, q.csr.ray_start);
    LIDF_LAUNCH_CHECK();
  } else if ((rc = lidf_roi_align_rays(p->full_rgb_feat, p->B, p->H, p->W, p->miss_img_ind, p->miss_bid, R, p->roi_inp_bbox,
                                       q.roi_feat, st))) return rc;
  if (P > 0) {
    // 3. weights (skipped when the caller's weight cache already holds them)
    const LidfDecoder* decs[2] = {&p->offset_dec, &p->prob_dec};
    SimtPack& sp = q.sp;
    if (!q.packed) LIDF_CUDA(cudaMemsetAsync(sp.Wt_row, 0, sizeof(float) * (size_t)sp.KR * sp.Ntot, st));
    for (int d = 0; d < 2 && !q.packed; ++d) {
      const LidfDecoder& dc = *decs[d];
      const int ldw = q.D + (dc.kind == LIDF_DEC_IEF ? LIDF_IEF_ENC : 0);
      // per-ray rows: [rgb(128) | PE(dir)]
      if ((rc = pack_wt(dc.w1, ldw, LIDF_VOX_DIM, LIDF_RGB_DIM, 256, sp.Wt_row, sp.Ntot, 0, 256 * d, st))) return rc;
      if ((rc = pack_wt(dc.w1, ldw, LIDF_VOX_DIM + LIDF_RGB_DIM + 2 * q.pe_pos, q.pe_dir, 256, sp.Wt_row, sp.Ntot,
                        LIDF_RGB_DIM, 256 * d, st))) return rc;
      k_pack_bias1<<<1, 256, 0, st>>>(dc.w1, ldw, q.D, dc.b1, dc.w_enc, dc.b_enc, dc.kind == LIDF_DEC_IEF,
                                      dc.init_offset, sp.bias_row + 256 * d, sp.u[d]);
      LIDF_LAUNCH_CHECK();
      if ((rc = pack_wt(dc.w1, ldw, 0, LIDF_VOX_DIM, 256, sp.Wt_vox, sp.Ntot, 0, 256 * d, st))) return rc;
      if (q.impl == LIDF_MLP_SIMT_FP32) {
        LIDF_CUDA(cudaMemsetAsync(sp.Wt_pe[d], 0, sizeof(float) * (size_t)sp.KP * 256, st));
        if ((rc = pack_wt(dc.w1, ldw, LIDF_VOX_DIM + LIDF_RGB_DIM, 2 * q.pe_pos, 256, sp.Wt_pe[d], 256, 0, 0, st))) return rc;
        if ((rc = pack_wt(dc.w2, 256, 0, 256, 128, sp.Wt2[d], 128, 0, 0, st))) return rc;
        if ((rc = pack_wt(dc.w3, 128, 0, 128, 64, sp.Wt3[d], 64, 0, 0, st))) return rc;
      }
    }
    // 4. row prep: per-ray term T[R][512] and per-voxel term A_v[V][512] (both decoders)
    {
      RowPrepArgs a{};
      a.featA = q.roi_feat; a.featB = nullptr; a.dirs = p->miss_ray_dir; a.rows = R;
      a.multires_views = p->multires_views; a.pos_encode = p->pos_encode;
      a.Wt = sp.Wt_row; a.Kpad = sp.KR; a.Ntot = sp.Ntot; a.bias = sp.bias_row; a.out = q.T;
      const size_t smem = sizeof(float) * ((size_t)LIDF_SIMT_BM * a.Kpad + LIDF_KC * 128);
      LIDF_CUDA(cudaFuncSetAttribute(k_rowprep, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
      // tensor-core row prep for the shipped layout ([roi 128 | PE(dir) 27] -> K = 160, both decoders -> N = 512)
      const bool rp_tc = q.impl != LIDF_MLP_SIMT_FP32 && sp.KR == 16 * RP_KSTEPS && sp.Ntot == 512 && p->pos_encode &&
                         p->multires_views == 4 && tc_device_ok();
      if (rp_tc) {
        if ((rc = tc_rowprep_forward(q.roi_feat, p->miss_ray_dir, R, sp.Wt_row, sp.Ntot, sp.bias_row, q.tc.rp_wstream, q.T,
                                     q.impl, st, &g_launches, g_cuda_err, sizeof(g_cuda_err), q.packed, q.live_list,
                                     q.live_count))) return rc;
      } else {
        k_rowprep<<<dim3((unsigned)((R + LIDF_SIMT_BM - 1) / LIDF_SIMT_BM), a.Ntot / 128), LIDF_SIMT_THREADS, smem, st>>>(a);
        LIDF_LAUNCH_CHECK();
      }
      if (V > 0) {
        a.featA = p->occ_voxel_feat; a.dirs = nullptr; a.rows = V; a.Wt = sp.Wt_vox; a.Kpad = 128; a.bias = nullptr;
        a.out = q.Av;
        const size_t smem2 = sizeof(float) * ((size_t)LIDF_SIMT_BM * 128 + LIDF_KC * 128);
        k_rowprep<<<dim3((unsigned)((V + LIDF_SIMT_BM - 1) / LIDF_SIMT_BM), a.Ntot / 128), LIDF_SIMT_THREADS, smem2, st>>>(a);
        LIDF_LAUNCH_CHECK();
      }
    }
  }
  return LIDF_OK;
}
}  // namespace

extern "C" size_t lidf_query_workspace_bytes(const LidfQueryParams* p) {
  if (!p) return 0;
  QueryPlan q;
  if (plan_query(p, &q, nullptr) != LIDF_OK) return 0;
  return q.bytes;
}
extern "C" size_t lidf_query_weight_cache_bytes(const LidfQueryParams* p) {
  if (!p) return 0;
  QueryPlan q;
  if (plan_query(p, &q, nullptr, true) != LIDF_OK) return 0;
  return q.cache_bytes;
}

extern "C" int lidf_query_forward(const LidfQueryParams* p, lidf_stream_t stream) {
  if (!p) return LIDF_ERR_NULL;
  if (p->P < 0 || p->R < 0 || p->V < 0 || p->B <= 0 || p->H <= 0 || p->W <= 0) return LIDF_ERR_ARG;
  if (p->P >= INT_MAX || p->R >= INT_MAX / 32 || p->V >= INT_MAX / 512) return LIDF_ERR_UNSUPPORTED;
  if (!p->max_pair_id || !p->pred_pos || !p->workspace) return LIDF_ERR_NULL;
  if (p->R > 0 && (!p->full_rgb_feat || !p->miss_ray_dir || !p->miss_img_ind || !p->miss_bid)) return LIDF_ERR_NULL;
  if (p->P > 0) {
    if (!p->occ_voxel_feat || !p->voxel_bound || !p->pair_vox || !p->pair_ray) return LIDF_ERR_NULL;
    if (!p->pair_dist && !p->dense_dist) return LIDF_ERR_NULL;
    if (!p->pred_prob_end || !p->pred_prob_end_softmax) return LIDF_ERR_NULL;
    if (!p->winner_only_offset && (!p->pred_offset || !p->pair_pred_pos)) return LIDF_ERR_NULL;
  }
  if (p->roi_inp_bbox < 0) return LIDF_ERR_ARG;
  if (p->prob_dec.kind != LIDF_DEC_IMNET) return LIDF_ERR_UNSUPPORTED;   // pipeline.py:81-85
  QueryPlan q;
  int rc = plan_query(p, &q, (char*)p->workspace);
  if (rc) return rc;
  if (p->winner_only_offset && q.impl == LIDF_MLP_SIMT_FP32) return LIDF_ERR_UNSUPPORTED;
  if (p->workspace_bytes < q.bytes) return LIDF_ERR_WORKSPACE;
  if ((rc = check_decoder(p->offset_dec, q.D))) return rc;
  if ((rc = check_decoder(p->prob_dec, q.D))) return rc;
  cudaStream_t st = stream;
  const int64_t P = p->P, R = p->R, V = p->V;
  if (R == 0) return LIDF_OK;

  if ((rc = run_prep(p, q, st))) return rc;
  if (p->index_error) { k_publish_flag<<<1, 1, 0, st>>>(q.csr.cnt + (R + 1), p->index_error); LIDF_LAUNCH_CHECK(); }
  if (P > 0) {
    const LidfDecoder* decs[2] = {&p->offset_dec, &p->prob_dec};
    SimtPack& sp = q.sp;
    // 5. decoders
    if (q.impl == LIDF_MLP_SIMT_FP32) {
      SimtMlpArgs a{};
      a.rows = P; a.refine = 0; a.perm = q.csr.perm; a.pair_vox = p->pair_vox; a.pair_ray = p->pair_ray;
      a.pair_dist = p->pair_dist; a.dense_dist = p->dense_dist; a.R = R; a.V = V; a.ray_dir = p->miss_ray_dir;
      a.voxel_bound = p->voxel_bound; a.rel = p->intersect_pos_rel; a.pos_encode = p->pos_encode;
      a.multires = p->multires; a.KP = sp.KP; a.n_dec = 2;
      a.o_iter = decs[0]->kind == LIDF_DEC_IEF ? p->ief_iter_out : nullptr;
      float* outs[2] = {p->pred_offset, p->pred_prob_end};
      for (int d = 0; d < 2; ++d) {
        const LidfDecoder& dc = *decs[d];
        SimtDecoder& s = a.dec[d];
        s.kind = dc.kind; s.n_pass = dc.kind == LIDF_DEC_IEF ? dc.n_iter : 1; s.use_sigmoid = dc.use_sigmoid;
        s.Wt_pe = sp.Wt_pe[d]; s.Wt2 = sp.Wt2[d]; s.b2 = dc.b2; s.Wt3 = sp.Wt3[d]; s.b3 = dc.b3; s.w4 = dc.w4; s.b4 = dc.b4;
        s.u = dc.kind == LIDF_DEC_IEF ? sp.u[d] : nullptr; s.o0 = dc.init_offset;
        s.addA = q.Av; s.ldA = 512; s.offA = 256 * d; s.addB = q.T; s.ldB = 512; s.offB = 256 * d; s.out = outs[d];
      }
      a.r0 = p->offset_range0; a.r1 = p->offset_range1; a.scale = (float)sqrt(3.0); a.scale2 = p->part_size;
      a.pos_out = p->pair_pred_pos;
      const size_t smem = simt_mlp_smem_bytes(sp.KP);
      LIDF_CUDA(cudaFuncSetAttribute(k_mlp_simt, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      mlp_event(0, st);
      k_mlp_simt<<<(unsigned)((P + LIDF_SIMT_BM - 1) / LIDF_SIMT_BM), LIDF_SIMT_THREADS, smem, st>>>(a);
      mlp_event(1, st);
      LIDF_LAUNCH_CHECK();
    } else if (p->winner_only_offset) {
      // 5'. winner-only mode: probability decoder over all pairs -> ray termination -> offset decoder on each ray's winner
      if ((rc = tc_query_forward(p, q.tc, q.csr.perm, q.T, q.Av, q.sp.u[0], q.pe_pos, q.D, q.impl, st, &g_launches, g_cuda_err,
                                 sizeof(g_cuda_err), mlp_event, 1))) return rc;
      k_ray_terminate<<<(unsigned)(((R + 3) / 4 * 32 + 255) / 256), 256, 0, st>>>(p->pred_prob_end, nullptr, p->pcl_label_float,
                                                                         q.csr.ray_start, q.csr.perm, P, R,
                                                                         p->pred_prob_end_softmax, p->max_pair_id, p->pred_pos, q.win);
      LIDF_LAUNCH_CHECK();
      LIDF_CUDA(cudaMemsetAsync(p->pred_pos, 0, sizeof(float) * 3 * (size_t)R, st));      // rays without a pair: zeros (pipeline.py:452)
      return tc_query_forward(p, q.tc, q.win, q.T, q.Av, q.sp.u[0], q.pe_pos, q.D, q.impl, st, &g_launches, g_cuda_err,
                              sizeof(g_cuda_err), mlp_event, 2, p->pred_offset_ray ? p->pred_offset_ray : q.off_ray);
    } else {
      if ((rc = tc_query_forward(p, q.tc, q.csr.perm, q.T, q.Av, q.sp.u[0], q.pe_pos, q.D, q.impl, st, &g_launches, g_cuda_err,
                                 sizeof(g_cuda_err), mlp_event))) return rc;
    }
  } else if (p->winner_only_offset) {
    LIDF_CUDA(cudaMemsetAsync(p->pred_pos, 0, sizeof(float) * 3 * (size_t)R, st));
  }
  // 6. ray termination
  k_ray_terminate<<<(unsigned)(((R + 3) / 4 * 32 + 255) / 256), 256, 0, st>>>(p->pred_prob_end, p->pair_pred_pos, p->pcl_label_float,
                                                                     q.csr.ray_start, q.csr.perm, P, R,
                                                                     p->pred_prob_end_softmax, p->max_pair_id, p->pred_pos);
  LIDF_LAUNCH_CHECK();
  return LIDF_OK;
}


// -------------------------------------------------------------------------------------------------
// Backward of lidf_query_forward (lidf_bwd.cuh).  Launch plan, all on the caller's stream:
//   run_prep (CSR, ROIAlign, T, A_v recomputed) -> weight streams (forward + transposed chunks) -> k_bwd_seed
//   per decoder, per chunk of ray-major pairs, IEF passes in reverse order:
//       k_mlp_bwd_tc -> k_wgrad_tc (dW3, dW2) ; after the passes: k_wgrad_tc (dW1[:,pos]), G_r / G_v segment sums
//   per decoder: wgrad finish, column sums, dW1[:,rgb|dir|vox] = G^T [roi | PE(dir) | occ_voxel_feat] (k_wgrad_tc)
//   d roi = G_r W1[:,rgb] -> k_roi_align_backward ; d occ_voxel_feat = G_v W1[:,vox]
// -------------------------------------------------------------------------------------------------
// timing events of the backward's tcgen05 section, one pair per device ordinal like the forward's (an event belongs to the
// device it was created on; a failed create / record only disables the timing, never the call)
static thread_local cudaEvent_t g_ev_bwd_dev[LIDF_MAX_DEVICES][2] = {};
static thread_local cudaEvent_t* g_ev_bwd = g_ev_bwd_dev[0];
static thread_local bool g_ev_bwd_valid = false;
static void bwd_event(int which, cudaStream_t st) {
  int dev = 0;
  if (which == 0) g_ev_bwd_valid = false;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= LIDF_MAX_DEVICES) { cudaGetLastError(); return; }
  g_ev_bwd = g_ev_bwd_dev[dev];
  if (!g_ev_bwd[0] && (cudaEventCreate(&g_ev_bwd[0]) != cudaSuccess || cudaEventCreate(&g_ev_bwd[1]) != cudaSuccess)) {
    cudaGetLastError(); g_ev_bwd[0] = g_ev_bwd[1] = nullptr; return;
  }
  if (cudaEventRecord(g_ev_bwd[which], st) != cudaSuccess) { cudaGetLastError(); g_ev_bwd_valid = false; return; }
  if (which == 1) g_ev_bwd_valid = true;
}

namespace {
struct BwdPlan {
  QueryPlan q;
  uint8_t* wbwd;                 // [2][20][8 KB]
  float* g[2];                   // [P] dL/d o of the running IEF iteration, per decoder
  int64_t chunk_rows;
  float *h1, *h2, *d1, *d1pk, *d2, *d3, *pe;
  int64_t* vkeys; CsrBufs vcsr;
  float *Gr, *Gv, *pedir, *droi, *Wg;
  float* partial; size_t partial_floats; int n_cta;
  float* colpart; float* du; float* dc;
  int* win; int* iota; float* g0r;          // winner-only backward: row list / identity CSR / seeds of the offset decoder, by ray
  size_t bytes;
};

int plan_backward(const LidfQueryBackwardParams* bp, BwdPlan* b, char* base) {
  LidfQueryParams fp = bp->fwd;
  fp.mlp_impl = LIDF_MLP_TC_BF16X3; fp.roi_feat_per_ray = nullptr; fp.weight_cache = nullptr; fp.weight_cache_valid = 0;
  fp.winner_only_offset = 0;                      // (the forward's scratch of that mode is not needed here)
  int rc = plan_query(&fp, &b->q, base, false, false);
  fp.winner_only_offset = bp->fwd.winner_only_offset;    // dense per-ray rows: the wgrad GEMMs read every ROI / T row
  if (rc) return rc;
  const int64_t P = fp.P, R = fp.R, V = fp.V;
  Bump bm{base, b->q.bytes};
  b->wbwd = bm.take<uint8_t>((size_t)2 * BW_CHUNKS_BWD * TC_CHUNK_BYTES);
  for (int d = 0; d < 2; ++d) b->g[d] = bm.take<float>((size_t)(P > 0 ? P : 1));
  int64_t cr = bp->chunk_rows > 0 ? bp->chunk_rows : ((int64_t)1 << 21);
  cr = (cr + 127) / 128 * 128;
  const int64_t p_pad = (P + 127) / 128 * 128;
  if (cr > p_pad) cr = p_pad;
  if (cr < 128) cr = 128;
  b->chunk_rows = cr;
  b->h1 = bm.take<float>((size_t)cr * LIDF_H1); b->h2 = bm.take<float>((size_t)cr * LIDF_H2);
  b->d1 = bm.take<float>((size_t)cr * LIDF_H1); b->d2 = bm.take<float>((size_t)cr * LIDF_H2);
  b->d3 = bm.take<float>((size_t)cr * LIDF_H3); b->pe = bm.take<float>((size_t)cr * BW_PE_F);
  b->d1pk = bm.take<float>((size_t)cr * LIDF_H1);
  b->vkeys = bm.take<int64_t>((size_t)cr);
  b->vcsr = carve_csr(bm, cr, V > 0 ? V : 1);
  b->Gr = bm.take<float>((size_t)(R > 0 ? R : 1) * 512); b->Gv = bm.take<float>((size_t)(V > 0 ? V : 1) * 512);
  b->pedir = bm.take<float>((size_t)(R > 0 ? R : 1) * 32);
  b->droi = bm.take<float>((size_t)(R > 0 ? R : 1) * 128);
  b->Wg = bm.take<float>((size_t)512 * 128);
  b->n_cta = 148;
  b->partial_floats = (size_t)256 * 128 + 256 * 32 + 256 * 128;                // row level: rgb | dir | vox  (>= dW3^T | dW2 | dW1[:,pos])
  b->partial = bm.take<float>(b->partial_floats * b->n_cta);
  b->colpart = bm.take<float>((size_t)b->n_cta * TC_ROW_WARPS * 32 * BW_COLPART);
  b->du = bm.take<float>(256); b->dc = bm.take<float>(256);
  const bool wo = fp.winner_only_offset != 0;
  b->win = bm.take<int>((size_t)(R > 0 ? R : 1));
  b->iota = bm.take<int>((size_t)R + 1);
  b->g0r = wo ? bm.take<float>((size_t)(R > 0 ? R : 1)) : nullptr;
  b->bytes = bm.off + 256;
  return LIDF_OK;
}

int sm_count() {
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
  return sms;
}

size_t wgrad_smem_bytes(int M, int N) { return 1024 + 2 * ((size_t)(M / 128) * 4 * 8192 + (size_t)4 * N * 64); }

// C partial slices += A^T B over `rows` rows
int launch_wgrad(const float* A, int lda, int M, const float* B, int ldb, int N, int n_valid, int64_t rows, float* partial,
                 int n_cta, cudaStream_t st) {
  if (rows <= 0) return LIDF_OK;
  if ((M != 128 && M != 256) || N % 16 || N < 16 || N > 256 || (M / 128) * N > 512) return LIDF_ERR_ARG;
  WgArgs a{};
  a.A = A; a.lda = lda; a.M = M; a.B = B; a.ldb = ldb; a.N = N; a.n_valid = n_valid; a.rows = rows; a.partial = partial;
  const int64_t groups = (rows + WG_ROWS - 1) / WG_ROWS;
  const int grid = (int)(groups < n_cta ? groups : n_cta);
  const size_t smem = wgrad_smem_bytes(M, N);
  LIDF_CUDA(cudaFuncSetAttribute(k_wgrad_tc<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_wgrad_tc<3><<<grid, WG_THREADS, smem, st>>>(a);
  LIDF_LAUNCH_CHECK();
  return LIDF_OK;
}
// same contract for operands in the packed hand-over layout (lidf_bwd.cuh: bw_pk_*): C[128][N] partial slices += A^T B over
// `groups` 64-row groups
int launch_wgrad_pk(const float* A, int FA, const float* B, int FB, int64_t groups, float* partial, int n_cta, cudaStream_t st) {
  if (groups <= 0) return LIDF_OK;
  if (FA != 128 || FB % 16 || FB < 16 || FB > 256) return LIDF_ERR_ARG;
  WgPkArgs a{};
  a.A = reinterpret_cast<const uint8_t*>(A); a.FA = FA; a.B = reinterpret_cast<const uint8_t*>(B); a.FB = FB;
  a.groups = groups; a.partial = partial;
  const size_t stage = bw_pk_group_bytes(FA) + bw_pk_group_bytes(FB);
  int ns = (int)((227 * 1024 - 1024) / stage);
  a.n_stages = ns > WPK_MAX_STAGES ? WPK_MAX_STAGES : ns;
  if (a.n_stages < 2) return LIDF_ERR_ARG;
  const size_t smem = 1024 + (size_t)a.n_stages * stage;
  const int grid = (int)(groups < n_cta ? groups : n_cta);
  LIDF_CUDA(cudaFuncSetAttribute(k_wgrad_pk_tc<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_wgrad_pk_tc<3><<<grid, WPK_THREADS, smem, st>>>(a);
  LIDF_LAUNCH_CHECK();
  return LIDF_OK;
}
int finish_wgrad(const float* partial, int n_cta, int M, int N, int mode, float* dst, int ldd, int n_keep, int pe_pos,
                 int ones_col, float* extra, cudaStream_t st) {
  WgFinishArgs f{partial, n_cta, M, N, mode, dst, ldd, n_keep, pe_pos, ones_col, extra};
  k_wgrad_finish<<<(M * N + 255) / 256, 256, 0, st>>>(f);
  LIDF_LAUNCH_CHECK();
  return LIDF_OK;
}
}  // namespace

extern "C" size_t lidf_query_backward_workspace_bytes(const LidfQueryBackwardParams* bp) {
  if (!bp) return 0;
  BwdPlan b;
  if (plan_backward(bp, &b, nullptr) != LIDF_OK) return 0;
  return b.bytes;
}

extern "C" float lidf_query_last_bwd_ms(void) {
  if (!g_ev_bwd_valid) return -1.f;
  float ms = -1.f;
  if (cudaEventSynchronize(g_ev_bwd[1]) != cudaSuccess) return -1.f;
  if (cudaEventElapsedTime(&ms, g_ev_bwd[0], g_ev_bwd[1]) != cudaSuccess) return -1.f;
  return ms;
}

extern "C" size_t lidf_wgrad_selftest_scratch_bytes(int32_t M, int32_t N) { return (size_t)148 * M * N * sizeof(float) + 256; }
extern "C" int lidf_wgrad_selftest(const float* A, const float* B, float* C, int64_t rows, int32_t M, int32_t N, void* scratch,
                                   lidf_stream_t stream) {
  if (!A || !B || !C || !scratch) return LIDF_ERR_NULL;
  if (!tc_device_ok()) return LIDF_ERR_NO_SM100;
  cudaStream_t st = stream;
  float* partial = (float*)scratch;
  LIDF_CUDA(cudaMemsetAsync(partial, 0, (size_t)148 * M * N * sizeof(float), st));
  int rc = launch_wgrad(A, M, M, B, N, N, N, rows, partial, 148, st);
  if (rc) return rc;
  return finish_wgrad(partial, 148, M, N, 1, C, N, N, 0, -1, nullptr, st);
}

// Byte offset of element (row, feature) of a [rows, F] tensor in the packed hand-over layout of the backward (hi part, or lo
// part with lo != 0); < 0 on bad arguments.  Host arithmetic only: lets a CPU test hold the layout against the canonical
// MN-major UMMA layout the wgrad kernel's descriptors assume.
extern "C" int64_t lidf_pk_offset_bytes(int32_t F, int64_t row, int32_t feature, int32_t lo) {
  if (F <= 0 || F % 8 || row < 0 || feature < 0 || feature >= F) return -1;
  return (int64_t)((size_t)(row / BW_PK_ROWS) * bw_pk_group_bytes(F) + (lo ? bw_pk_lo_offset(F) : 0) +
                   (size_t)(feature >> 3) * BW_PK_FG_BYTES + (size_t)(row % BW_PK_ROWS) * 16 + (size_t)(feature & 7) * 2);
}

extern "C" size_t lidf_wgrad_pk_selftest_scratch_bytes(int64_t rows, int32_t M, int32_t N) {
  const size_t groups = (size_t)((rows + BW_PK_ROWS - 1) / BW_PK_ROWS);
  return (size_t)148 * M * N * sizeof(float) + groups * (bw_pk_group_bytes(M) + bw_pk_group_bytes(N)) + 1024;
}
extern "C" int lidf_wgrad_pk_selftest(const float* A, const float* B, float* C, int64_t rows, int32_t M, int32_t N, void* scratch,
                                      lidf_stream_t stream) {
  if (!A || !B || !C || !scratch) return LIDF_ERR_NULL;
  if (rows <= 0 || M != 128 || N % 16 || N < 16 || N > 256) return LIDF_ERR_ARG;
  if (!tc_device_ok()) return LIDF_ERR_NO_SM100;
  cudaStream_t st = stream;
  const int64_t groups = (rows + BW_PK_ROWS - 1) / BW_PK_ROWS;
  float* partial = (float*)scratch;
  uint8_t* Apk = (uint8_t*)scratch + ((size_t)148 * M * N * sizeof(float) + 255) / 256 * 256;
  uint8_t* Bpk = Apk + (size_t)groups * bw_pk_group_bytes(M);
  LIDF_CUDA(cudaMemsetAsync(partial, 0, (size_t)148 * M * N * sizeof(float), st));
  k_pk_pack_rows<<<(unsigned)((groups * BW_PK_ROWS * (M / 8) + 255) / 256), 256, 0, st>>>(A, rows, M, Apk);
  LIDF_LAUNCH_CHECK();
  k_pk_pack_rows<<<(unsigned)((groups * BW_PK_ROWS * (N / 8) + 255) / 256), 256, 0, st>>>(B, rows, N, Bpk);
  LIDF_LAUNCH_CHECK();
  int rc = launch_wgrad_pk((const float*)Apk, M, (const float*)Bpk, N, groups, partial, 148, st);
  if (rc) return rc;
  return finish_wgrad(partial, 148, M, N, 1, C, N, N, 0, -1, nullptr, st);
}

extern "C" int lidf_query_backward(const LidfQueryBackwardParams* bp, lidf_stream_t stream) {
  if (!bp) return LIDF_ERR_NULL;
  LidfQueryParams fp = bp->fwd;
  fp.mlp_impl = LIDF_MLP_TC_BF16X3; fp.roi_feat_per_ray = nullptr; fp.index_error = nullptr; fp.ief_iter_out = nullptr;
  fp.weight_cache = nullptr; fp.weight_cache_valid = 0;
  const LidfQueryParams* p = &fp;
  if (p->P < 0 || p->R < 0 || p->V < 0 || p->B <= 0 || p->H <= 0 || p->W <= 0) return LIDF_ERR_ARG;
  if (p->P >= INT_MAX || p->R >= INT_MAX / 32 || p->V >= INT_MAX / 512) return LIDF_ERR_UNSUPPORTED;
  if (!bp->workspace) return LIDF_ERR_NULL;
  if (p->R > 0 && (!p->full_rgb_feat || !p->miss_ray_dir || !p->miss_img_ind || !p->miss_bid || !p->max_pair_id)) return LIDF_ERR_NULL;
  if (p->P > 0) {
    if (!p->occ_voxel_feat || !p->voxel_bound || !p->pair_vox || !p->pair_ray) return LIDF_ERR_NULL;
    if (!p->pair_dist && !p->dense_dist) return LIDF_ERR_NULL;
    if (!p->pred_prob_end) return LIDF_ERR_NULL;
    if (!p->winner_only_offset && !p->pred_offset) return LIDF_ERR_NULL;
    if (p->winner_only_offset && (!p->pred_offset_ray || bp->g_pred_offset || bp->g_pair_pred_pos)) return LIDF_ERR_ARG;
  }
  const bool wo = p->winner_only_offset != 0;
  // The offset decoder's upstream gradient is non-zero on ONE pair per ray unless the caller differentiates pred_offset /
  // pair_pred_pos themselves (the reference's loss never does: pos_loss sees pred_pos = pair_pred_pos[max_pair_id],
  // pipeline.py:449-454,472).  Rows with a zero upstream gradient contribute exactly zero to every gradient, so the offset
  // decoder's backward then runs over the R winner rows only -- also after a full (not winner-only) forward.
  const bool rows_by_ray = wo || (!bp->g_pred_offset && !bp->g_pair_pred_pos);
  if (p->prob_dec.kind != LIDF_DEC_IMNET) return LIDF_ERR_UNSUPPORTED;
  if (!tc_device_ok()) return LIDF_ERR_NO_SM100;
  BwdPlan b;
  int rc = plan_backward(bp, &b, (char*)bp->workspace);
  if (rc) return rc;
  if (bp->workspace_bytes < b.bytes) return LIDF_ERR_WORKSPACE;
  QueryPlan& q = b.q;
  if (!p->pos_encode || p->multires != 8 || q.pe_pos != 51 || p->multires_views != 4) return LIDF_ERR_UNSUPPORTED;
  if ((rc = check_decoder(p->offset_dec, q.D))) return rc;
  if ((rc = check_decoder(p->prob_dec, q.D))) return rc;
  const LidfDecoder* decs[2] = {&p->offset_dec, &p->prob_dec};
  const LidfDecoderGrad* grads[2] = {&bp->g_offset_dec, &bp->g_prob_dec};
  for (int d = 0; d < 2; ++d) {
    const LidfDecoderGrad& gd = *grads[d];
    if (!gd.w1 || !gd.b1 || !gd.w2 || !gd.b2 || !gd.w3 || !gd.b3 || !gd.w4 || !gd.b4) return LIDF_ERR_NULL;
    if (decs[d]->kind == LIDF_DEC_IEF && (!gd.w_enc || !gd.b_enc)) return LIDF_ERR_NULL;
  }
  const int n_pass0 = decs[0]->kind == LIDF_DEC_IEF ? decs[0]->n_iter : 1;
  if (n_pass0 > 1 && p->P > 0 && !bp->ief_iter) return LIDF_ERR_NULL;
  cudaStream_t st = stream;
  const int64_t P = p->P, R = p->R, V = p->V;
  const int ldw[2] = {q.D + (decs[0]->kind == LIDF_DEC_IEF ? LIDF_IEF_ENC : 0), q.D + (decs[1]->kind == LIDF_DEC_IEF ? LIDF_IEF_ENC : 0)};

  // every output is overwritten
  for (int d = 0; d < 2; ++d) {
    const LidfDecoderGrad& gd = *grads[d];
    LIDF_CUDA(cudaMemsetAsync(gd.w1, 0, sizeof(float) * 256 * (size_t)ldw[d], st));
    LIDF_CUDA(cudaMemsetAsync(gd.b1, 0, sizeof(float) * 256, st));
    LIDF_CUDA(cudaMemsetAsync(gd.w2, 0, sizeof(float) * 128 * 256, st));
    LIDF_CUDA(cudaMemsetAsync(gd.b2, 0, sizeof(float) * 128, st));
    LIDF_CUDA(cudaMemsetAsync(gd.w3, 0, sizeof(float) * 64 * 128, st));
    LIDF_CUDA(cudaMemsetAsync(gd.b3, 0, sizeof(float) * 64, st));
    LIDF_CUDA(cudaMemsetAsync(gd.w4, 0, sizeof(float) * 64, st));
    LIDF_CUDA(cudaMemsetAsync(gd.b4, 0, sizeof(float), st));
    if (decs[d]->kind == LIDF_DEC_IEF) {
      LIDF_CUDA(cudaMemsetAsync(gd.w_enc, 0, sizeof(float) * LIDF_IEF_ENC, st));
      LIDF_CUDA(cudaMemsetAsync(gd.b_enc, 0, sizeof(float) * LIDF_IEF_ENC, st));
    }
  }
  if (bp->g_full_rgb_feat) LIDF_CUDA(cudaMemsetAsync(bp->g_full_rgb_feat, 0, sizeof(float) * (size_t)p->B * LIDF_RGB_CH * p->H * p->W, st));
  if (bp->g_occ_voxel_feat && V > 0) LIDF_CUDA(cudaMemsetAsync(bp->g_occ_voxel_feat, 0, sizeof(float) * (size_t)V * 128, st));
  if (R == 0 || P == 0) return LIDF_OK;

  if ((rc = run_prep(p, q, st))) return rc;
  for (int d = 0; d < 2; ++d) {
    k_pack_tc_weights<<<(TC_CHUNKS_PER_DEC * 2048 + 255) / 256, 256, 0, st>>>(
        decs[d]->w1, ldw[d], q.pe_pos, decs[d]->w2, decs[d]->w3, q.tc.wstream + (size_t)d * TC_CHUNKS_PER_DEC * TC_CHUNK_BYTES, 0);
    LIDF_LAUNCH_CHECK();
    k_pack_tc_weights_bwd<<<(BW_CHUNKS_BWD * 2048 + 255) / 256, 256, 0, st>>>(decs[d]->w2, decs[d]->w3,
                                                                             b.wbwd + (size_t)d * BW_CHUNKS_BWD * TC_CHUNK_BYTES);
    LIDF_LAUNCH_CHECK();
  }
  LIDF_CUDA(cudaMemsetAsync(b.Gr, 0, sizeof(float) * (size_t)R * 512, st));
  LIDF_CUDA(cudaMemsetAsync(b.Gv, 0, sizeof(float) * (size_t)V * 512, st));
  {
    BwSeedArgs a{};
    a.P = P; a.R = R; a.pair_ray = p->pair_ray; a.max_pair_id = p->max_pair_id; a.ray_dir = p->miss_ray_dir;
    a.pred_offset = p->pred_offset; a.pred_prob_end = p->pred_prob_end;
    a.g_pred_pos = bp->g_pred_pos; a.g_pred_prob_end = bp->g_pred_prob_end; a.g_pred_offset = bp->g_pred_offset;
    a.g_pair_pred_pos = bp->g_pair_pred_pos;
    a.scale = (float)((double)(p->offset_range1 - p->offset_range0) * sqrt(3.0) * (double)p->part_size);
    a.sig0 = decs[0]->use_sigmoid; a.sig1 = decs[1]->use_sigmoid; a.g0 = b.g[0]; a.g1 = b.g[1];
    if (wo) {           // the offset decoder's seeds live by ray (below); the per-pair pass only forms the probability decoder's
      a.g_pred_pos = nullptr; a.pred_offset = p->pred_prob_end; a.g0 = b.g[0];
    }
    k_bwd_seed<<<(unsigned)((P + 255) / 256), 256, 0, st>>>(a);
    LIDF_LAUNCH_CHECK();
    if (rows_by_ray) {
      k_bwd_winner_rows<<<(unsigned)((R + 1 + 255) / 256), 256, 0, st>>>(p->max_pair_id, P, R, b.win, b.iota);
      LIDF_LAUNCH_CHECK();
    }
    if (wo) {
      k_bwd_seed_rays<<<(unsigned)((R + 255) / 256), 256, 0, st>>>(bp->g_pred_pos, p->miss_ray_dir, p->pred_offset_ray, b.win, R,
                                                                    a.scale, a.sig0, b.g0r);
      LIDF_LAUNCH_CHECK();
    }
  }
  k_bwd_pedir<<<(unsigned)((R + 255) / 256), 256, 0, st>>>(p->miss_ray_dir, R, p->multires_views, p->pos_encode, b.pedir);
  LIDF_LAUNCH_CHECK();

  const int sms = sm_count();
  const int n_cta = sms < b.n_cta ? sms : b.n_cta;
  const size_t bw_smem = sizeof(BwSmem) + 1024;
  LIDF_CUDA(cudaFuncSetAttribute(k_mlp_bwd_tc<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bw_smem));
  bwd_event(0, st);
  float* part_w3 = b.partial;                                   // [n_cta][128*64]
  float* part_w2 = part_w3 + (size_t)b.n_cta * 128 * 64;        // [n_cta][128*256]
  float* part_w1 = part_w2 + (size_t)b.n_cta * 128 * 256;       // [n_cta][128*256]  (PE^T delta1)
  for (int d = 0; d < 2; ++d) {
    const LidfDecoder& dc = *decs[d];
    const LidfDecoderGrad& gd = *grads[d];
    const bool ief = dc.kind == LIDF_DEC_IEF;
    const int n_pass = ief ? dc.n_iter : 1;
    LIDF_CUDA(cudaMemsetAsync(b.partial, 0, sizeof(float) * b.partial_floats * b.n_cta, st));
    LIDF_CUDA(cudaMemsetAsync(b.colpart, 0, sizeof(float) * (size_t)b.n_cta * TC_ROW_WARPS * 32 * BW_COLPART, st));
    // rows of this decoder's backward: all P pairs in ray-major order, or (winner-only mode, offset decoder) one row per ray
    const bool by_ray = rows_by_ray && d == 0;
    const bool by_slot = wo && d == 0;              // seeds / saved offsets indexed by ray (winner-only forward) or by pair
    const int64_t n_dom = by_ray ? R : P;
    const int* row_perm = by_ray ? b.win : q.csr.perm;
    const int* row_ray_start = by_ray ? b.iota : q.csr.ray_start;
    for (int64_t s0 = 0; s0 < n_dom; s0 += b.chunk_rows) {
      const int n_rows = (int)((n_dom - s0) < b.chunk_rows ? (n_dom - s0) : b.chunk_rows);
      const int n_tiles = (n_rows + 127) / 128;
      for (int it = n_pass - 1; it >= 0; --it) {
        BwArgs a{};
        a.P = P; a.s0 = s0; a.n_rows = n_rows; a.n_tiles = n_tiles;
        a.by_slot = by_slot ? 1 : 0;
        a.perm = row_perm; a.pair_vox = p->pair_vox; a.pair_ray = p->pair_ray; a.pair_dist = p->pair_dist;
        a.dense_dist = p->dense_dist; a.R = R; a.V = V; a.ray_dir = p->miss_ray_dir; a.voxel_bound = p->voxel_bound;
        a.rel = p->intersect_pos_rel; a.Av = q.Av; a.T = q.T; a.dcol = 256 * d;
        a.wfwd = q.tc.wstream + (size_t)d * TC_CHUNKS_PER_DEC * TC_CHUNK_BYTES;
        a.wbwd = b.wbwd + (size_t)d * BW_CHUNKS_BWD * TC_CHUNK_BYTES;
        a.u = ief ? q.sp.u[d] : nullptr; a.b2 = dc.b2; a.b3 = dc.b3; a.w4 = dc.w4;
        a.is_ief = ief ? 1 : 0; a.it = it; a.o0 = ief ? dc.init_offset : 0.f;
        a.o_in = (ief && it > 0) ? bp->ief_iter + (size_t)(it - 1) * (by_slot ? R : P) : nullptr;   // [n_iter-1][P], or [..][R] by ray
        a.g = by_slot ? b.g0r : b.g[d];
        a.h1 = b.h1; a.h2 = b.h2; a.d1 = b.d1; a.d2 = b.d2; a.d3 = b.d3;
        a.pe = it == n_pass - 1 ? b.pe : nullptr;
        a.d1pk = it == 0 ? b.d1pk : nullptr;
        a.d1_accumulate = it == n_pass - 1 ? 0 : 1;
        a.colpart = b.colpart;
        const int grid = n_tiles < n_cta ? n_tiles : n_cta;
        k_mlp_bwd_tc<3><<<grid, TC_THREADS, bw_smem, st>>>(a);
        LIDF_LAUNCH_CHECK();
        // B1 wrote h1, h2, delta2, delta3 of all n_tiles x 128 rows in the packed layout (dead rows: delta = 0)
        if ((rc = launch_wgrad_pk(b.h2, LIDF_H2, b.d3, LIDF_H3, 2 * (int64_t)n_tiles, part_w3, n_cta, st))) return rc;   // dW3^T
        if ((rc = launch_wgrad_pk(b.d2, LIDF_H2, b.h1, LIDF_H1, 2 * (int64_t)n_tiles, part_w2, n_cta, st))) return rc;   // dW2
      }
      if ((rc = launch_wgrad_pk(b.pe, BW_PE_F, b.d1pk, LIDF_H1, 2 * (int64_t)n_tiles, part_w1, n_cta, st))) return rc;    // dW1[:,pos]^T
      k_segsum_rays<<<(unsigned)((R * 32 + 255) / 256), 256, 0, st>>>(b.d1, s0, n_rows, row_ray_start, R, b.Gr, 256 * d);
      LIDF_LAUNCH_CHECK();
      k_chunk_vox_keys<<<(n_rows + 255) / 256, 256, 0, st>>>(row_perm, p->pair_vox, s0, n_rows, V, b.vkeys);
      LIDF_LAUNCH_CHECK();
      if ((rc = build_csr(b.vcsr, b.vkeys, n_rows, V, st, false))) return rc;
      k_segsum_vox<<<(n_rows + 63) / 64, 64, 0, st>>>(b.d1, b.vcsr.perm, b.vcsr.ray_start, V, n_rows, b.Gv, 256 * d);
      LIDF_LAUNCH_CHECK();
    }
    if ((rc = finish_wgrad(part_w3, n_cta, 128, 64, 0, gd.w3, 0, 0, 0, -1, nullptr, st))) return rc;
    if ((rc = finish_wgrad(part_w2, n_cta, 128, 256, 1, gd.w2, 256, 256, 0, -1, nullptr, st))) return rc;
    if ((rc = finish_wgrad(part_w1, n_cta, 128, 256, 4, gd.w1, ldw[d], 0, q.pe_pos, -1, nullptr, st))) return rc;
    k_bwd_colpart_finish<<<1, 512, 0, st>>>(b.colpart, n_cta, gd.b2, gd.b3, gd.w4, gd.b4, ief ? b.du : nullptr);
    LIDF_LAUNCH_CHECK();
    // row-level weight gradients of layer 1 from the segment sums: dW1[:,rgb] = G_r^T roi, dW1[:,dir] = G_r^T PE(dir)
    // (+ its ones column = sum_r G_r = db1 / d c), dW1[:,vox] = G_v^T occ_voxel_feat
    LIDF_CUDA(cudaMemsetAsync(b.partial, 0, sizeof(float) * b.partial_floats * b.n_cta, st));
    float* part_rgb = b.partial;                                // [n_cta][256*128]
    float* part_dir = part_rgb + (size_t)b.n_cta * 256 * 128;   // [n_cta][256*32]
    float* part_vox = part_dir + (size_t)b.n_cta * 256 * 32;    // [n_cta][256*128]
    if ((rc = launch_wgrad(b.Gr + 256 * d, 512, 256, q.roi_feat, 128, 128, 128, R, part_rgb, n_cta, st))) return rc;
    if ((rc = launch_wgrad(b.Gr + 256 * d, 512, 256, b.pedir, 32, 32, 32, R, part_dir, n_cta, st))) return rc;
    if ((rc = launch_wgrad(b.Gv + 256 * d, 512, 256, p->occ_voxel_feat, 128, 128, 128, V, part_vox, n_cta, st))) return rc;
    if ((rc = finish_wgrad(part_rgb, n_cta, 256, 128, 1, gd.w1 + LIDF_VOX_DIM, ldw[d], 128, 0, -1, nullptr, st))) return rc;
    if ((rc = finish_wgrad(part_dir, n_cta, 256, 32, 3, gd.w1 + LIDF_VOX_DIM + LIDF_RGB_DIM + 2 * q.pe_pos, ldw[d], q.pe_dir, 0, 27,
                           b.dc, st))) return rc;
    if ((rc = finish_wgrad(part_vox, n_cta, 256, 128, 1, gd.w1, ldw[d], 128, 0, -1, nullptr, st))) return rc;
    k_bwd_finish_decoder<<<1, 256, 0, st>>>(dc.w1, ldw[d], q.D, dc.w_enc, dc.b_enc, ief ? 1 : 0, b.du, b.dc, gd.w1, gd.b1,
                                            gd.w_enc, gd.b_enc);
    LIDF_LAUNCH_CHECK();
  }
  bwd_event(1, st);
  // feature gradients: d roi = G_r [W1_off[:,rgb]; W1_prob[:,rgb]] -> ROIAlign^T ; d occ_voxel_feat = G_v [W1[:,vox]; ...]
  const size_t gemm_smem = sizeof(float) * ((size_t)LIDF_SIMT_BM * 512 + LIDF_KC * 128);
  LIDF_CUDA(cudaFuncSetAttribute(k_bwd_rows_gemm, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gemm_smem));
  if (bp->g_full_rgb_feat) {
    for (int d = 0; d < 2; ++d) {
      k_bwd_pack_rows<<<(256 * 128 + 255) / 256, 256, 0, st>>>(decs[d]->w1, ldw[d], LIDF_VOX_DIM, 256, b.Wg, 256 * d);
      LIDF_LAUNCH_CHECK();
    }
    k_bwd_rows_gemm<<<(unsigned)((R + LIDF_SIMT_BM - 1) / LIDF_SIMT_BM), LIDF_SIMT_THREADS, gemm_smem, st>>>(b.Gr, R, b.Wg, b.droi);
    LIDF_LAUNCH_CHECK();
    k_roi_align_backward<<<(unsigned)((R + 31) / 32), LIDF_ROI_THREADS, 0, st>>>(b.droi, p->B, p->H, p->W, p->miss_img_ind,
                                                                                  p->miss_bid, R, p->roi_inp_bbox / 2,
                                                                                  bp->g_full_rgb_feat);
    LIDF_LAUNCH_CHECK();
  }
  if (bp->g_occ_voxel_feat && V > 0) {
    for (int d = 0; d < 2; ++d) {
      k_bwd_pack_rows<<<(256 * 128 + 255) / 256, 256, 0, st>>>(decs[d]->w1, ldw[d], 0, 256, b.Wg, 256 * d);
      LIDF_LAUNCH_CHECK();
    }
    k_bwd_rows_gemm<<<(unsigned)((V + LIDF_SIMT_BM - 1) / LIDF_SIMT_BM), LIDF_SIMT_THREADS, gemm_smem, st>>>(b.Gv, V, b.Wg,
                                                                                                             bp->g_occ_voxel_feat);
    LIDF_LAUNCH_CHECK();
  }
  return LIDF_OK;
}

// -------------------------------------------------------------------------------------------------
// RefineNet decoder tail (pipeline.py:1018-1029).  Per ray; every layer-1 input except PE(pos) is folded into the
// row-prep GEMM: T'[r] = W1[:,0:128] vox_end[r] + W1[:,128:256] rgb_end[r] + W1[:,dir] PE(dir_r) + b1 (+ IEF const).
// -------------------------------------------------------------------------------------------------
namespace {
struct RefinePlan { int impl, pe_pos, pe_dir, D, KR, KP; bool tc; float* T; float* Av; float* scratch; SimtPack sp; TcBufs tcb; size_t bytes; };
int plan_refine(const LidfRefineParams* p, RefinePlan* q, char* base) {
  int rc = resolve_impl(p->mlp_impl, &q->impl, p->pos_encode, p->multires);
  if (rc) return rc;
  if (p->multires < 0 || p->multires > LIDF_MAX_MULTIRES || p->multires_views < 0 || p->multires_views > LIDF_MAX_MULTIRES)
    return LIDF_ERR_UNSUPPORTED;
  q->pe_pos = lidf_pe_dim(p->multires, p->pos_encode);
  q->pe_dir = lidf_pe_dim(p->multires_views, p->pos_encode);
  q->D = LIDF_VOX_DIM + LIDF_RGB_DIM + q->pe_pos + q->pe_dir;          // pipeline.py:740-742
  q->KR = pad16(LIDF_VOX_DIM + LIDF_RGB_DIM + q->pe_dir);
  q->KP = pad16(q->pe_pos);
  Bump b{base, 0};
  // tcgen05 path: needs the un-gathered voxel features (per-voxel layer-1 term + in-kernel gather) and the shipped encodings
  q->tc = q->impl != LIDF_MLP_SIMT_FP32 && p->occ_voxel_feat && p->end_voxel_id && p->V > 0 && p->pos_encode &&
          p->multires == 8 && p->multires_views == 4 && (!p->intersect_pos_rel || p->voxel_bound);
  if (q->tc) {
    q->KR = pad16(LIDF_RGB_DIM + q->pe_dir);                           // per-ray rows: [rgb(128) | PE(dir)], as in stage 1
    q->T = b.take<float>((size_t)p->R * 512);
    q->Av = b.take<float>((size_t)p->V * 512);
    q->scratch = b.take<float>((size_t)p->R);
    q->sp = carve_simt_pack(b, 2, q->KR, q->KP, true);
    q->tcb = carve_tc(b, p->V, 1);
  } else {
    q->T = b.take<float>((size_t)p->R * 256);
    q->Av = nullptr; q->scratch = nullptr;
    q->sp = carve_simt_pack(b, 1, q->KR, q->KP, false);
  }
  q->bytes = b.off + 256;
  return LIDF_OK;
}

// tcgen05 path of the refine tail: same factoring as lidf_query_forward with one decoder in slot 0 (slot 1 zero)
int refine_forward_tc(const LidfRefineParams* p, RefinePlan& q, cudaStream_t st) {
  int rc;
  const LidfDecoder& dc = p->offset_dec;
  const int ldw = q.D + (dc.kind == LIDF_DEC_IEF ? LIDF_IEF_ENC : 0);
  SimtPack& sp = q.sp;                                                  // Ntot = 512
  LIDF_CUDA(cudaMemsetAsync(sp.Wt_row, 0, sizeof(float) * (size_t)sp.KR * sp.Ntot, st));
  LIDF_CUDA(cudaMemsetAsync(sp.bias_row, 0, sizeof(float) * sp.Ntot, st));
  LIDF_CUDA(cudaMemsetAsync(sp.Wt_vox, 0, sizeof(float) * (size_t)128 * sp.Ntot, st));
  if ((rc = pack_wt(dc.w1, ldw, LIDF_VOX_DIM, LIDF_RGB_DIM, 256, sp.Wt_row, sp.Ntot, 0, 0, st))) return rc;              // rgb
  if ((rc = pack_wt(dc.w1, ldw, LIDF_VOX_DIM + LIDF_RGB_DIM + q.pe_pos, q.pe_dir, 256, sp.Wt_row, sp.Ntot, LIDF_RGB_DIM, 0,
                    st))) return rc;                                                                                     // PE(dir)
  k_pack_bias1<<<1, 256, 0, st>>>(dc.w1, ldw, q.D, dc.b1, dc.w_enc, dc.b_enc, dc.kind == LIDF_DEC_IEF, dc.init_offset,
                                  sp.bias_row, sp.u[0]);
  LIDF_LAUNCH_CHECK();
  if ((rc = pack_wt(dc.w1, ldw, 0, LIDF_VOX_DIM, 256, sp.Wt_vox, sp.Ntot, 0, 0, st))) return rc;                         // voxel
  if ((rc = tc_rowprep_forward(p->rgb_feat_end, p->miss_ray_dir, p->R, sp.Wt_row, sp.Ntot, sp.bias_row, q.tcb.rp_wstream, q.T,
                               q.impl, st, &g_launches, g_cuda_err, sizeof(g_cuda_err)))) return rc;
  {
    RowPrepArgs a{};
    a.featA = p->occ_voxel_feat; a.featB = nullptr; a.dirs = nullptr; a.rows = p->V;
    a.multires_views = p->multires_views; a.pos_encode = p->pos_encode;
    a.Wt = sp.Wt_vox; a.Kpad = 128; a.Ntot = sp.Ntot; a.bias = nullptr; a.out = q.Av;
    const size_t smem = sizeof(float) * ((size_t)LIDF_SIMT_BM * 128 + LIDF_KC * 128);
    LIDF_CUDA(cudaFuncSetAttribute(k_rowprep, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    k_rowprep<<<dim3((unsigned)((p->V + LIDF_SIMT_BM - 1) / LIDF_SIMT_BM), a.Ntot / 128), LIDF_SIMT_THREADS, smem, st>>>(a);
    LIDF_LAUNCH_CHECK();
  }
  return tc_refine_forward(p, q.tcb, q.T, q.Av, sp.u[0], q.scratch, q.pe_pos, q.D, q.impl, st, &g_launches, g_cuda_err,
                           sizeof(g_cuda_err));
}
}  // namespace

extern "C" size_t lidf_refine_workspace_bytes(const LidfRefineParams* p) {
  if (!p) return 0;
  RefinePlan q;
  if (plan_refine(p, &q, nullptr) != LIDF_OK) return 0;
  return q.bytes;
}

extern "C" int lidf_refine_forward(const LidfRefineParams* p, lidf_stream_t stream) {
  if (!p) return LIDF_ERR_NULL;
  if (p->R < 0) return LIDF_ERR_ARG;
  if (p->R >= INT_MAX / 512) return LIDF_ERR_UNSUPPORTED;
  if (p->R == 0) return LIDF_OK;
  if (!p->pred_pos || !p->miss_ray_dir || !p->rgb_feat_end || !p->pred_pos_refine || !p->workspace) return LIDF_ERR_NULL;
  if (p->V < 0 || p->V >= INT_MAX / 512) return LIDF_ERR_ARG;
  RefinePlan q;
  int rc = plan_refine(p, &q, (char*)p->workspace);
  if (rc) return rc;
  if (p->workspace_bytes < q.bytes) return LIDF_ERR_WORKSPACE;
  if ((rc = check_decoder(p->offset_dec, q.D))) return rc;
  cudaStream_t st = stream;
  if (q.tc) return refine_forward_tc(p, q, st);
  if (!p->voxel_feat_end) return LIDF_ERR_NULL;
  if (p->intersect_pos_rel && !p->end_voxel_center) return LIDF_ERR_NULL;
  const LidfDecoder& dc = p->offset_dec;
  const int ldw = q.D + (dc.kind == LIDF_DEC_IEF ? LIDF_IEF_ENC : 0);
  SimtPack& sp = q.sp;
  LIDF_CUDA(cudaMemsetAsync(sp.Wt_row, 0, sizeof(float) * (size_t)sp.KR * sp.Ntot, st));
  if ((rc = pack_wt(dc.w1, ldw, 0, 256, 256, sp.Wt_row, 256, 0, 0, st))) return rc;                        // vox | rgb
  if ((rc = pack_wt(dc.w1, ldw, 256 + q.pe_pos, q.pe_dir, 256, sp.Wt_row, 256, 256, 0, st))) return rc;    // PE(dir)
  k_pack_bias1<<<1, 256, 0, st>>>(dc.w1, ldw, q.D, dc.b1, dc.w_enc, dc.b_enc, dc.kind == LIDF_DEC_IEF, dc.init_offset,
                                  sp.bias_row, sp.u[0]);
  LIDF_LAUNCH_CHECK();
  LIDF_CUDA(cudaMemsetAsync(sp.Wt_pe[0], 0, sizeof(float) * (size_t)sp.KP * 256, st));
  if ((rc = pack_wt(dc.w1, ldw, 256, q.pe_pos, 256, sp.Wt_pe[0], 256, 0, 0, st))) return rc;
  if ((rc = pack_wt(dc.w2, 256, 0, 256, 128, sp.Wt2[0], 128, 0, 0, st))) return rc;
  if ((rc = pack_wt(dc.w3, 128, 0, 128, 64, sp.Wt3[0], 64, 0, 0, st))) return rc;
  {
    RowPrepArgs a{};
    a.featA = p->voxel_feat_end; a.featB = p->rgb_feat_end; a.dirs = p->miss_ray_dir; a.rows = p->R;
    a.multires_views = p->multires_views; a.pos_encode = p->pos_encode;
    a.Wt = sp.Wt_row; a.Kpad = sp.KR; a.Ntot = 256; a.bias = sp.bias_row; a.out = q.T;
    const size_t smem = sizeof(float) * ((size_t)LIDF_SIMT_BM * a.Kpad + LIDF_KC * 128);
    LIDF_CUDA(cudaFuncSetAttribute(k_rowprep, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    k_rowprep<<<dim3((unsigned)((p->R + LIDF_SIMT_BM - 1) / LIDF_SIMT_BM), 2), LIDF_SIMT_THREADS, smem, st>>>(a);
    LIDF_LAUNCH_CHECK();
  }
  // The refine tail is ~1% of stage-1 work: it runs on the fp32 engine for every mlp_impl.
  SimtMlpArgs a{};
  a.rows = p->R; a.refine = 1; a.ray_dir = p->miss_ray_dir; a.pos_in = p->pred_pos; a.center_in = p->end_voxel_center;
  a.rel = p->intersect_pos_rel; a.pos_encode = p->pos_encode; a.multires = p->multires; a.KP = sp.KP; a.n_dec = 1;
  SimtDecoder& s = a.dec[0];
  s.kind = dc.kind; s.n_pass = dc.kind == LIDF_DEC_IEF ? dc.n_iter : 1; s.use_sigmoid = dc.use_sigmoid;
  s.Wt_pe = sp.Wt_pe[0]; s.Wt2 = sp.Wt2[0]; s.b2 = dc.b2; s.Wt3 = sp.Wt3[0]; s.b3 = dc.b3; s.w4 = dc.w4; s.b4 = dc.b4;
  s.u = dc.kind == LIDF_DEC_IEF ? sp.u[0] : nullptr; s.o0 = dc.init_offset;
  s.addA = nullptr; s.addB = q.T; s.ldB = 256; s.offB = 0; s.out = nullptr;
  a.r0 = p->offset_range0; a.r1 = p->offset_range1; a.scale = 1.f; a.scale2 = 1.f; a.pos_out = p->pred_pos_refine;
  const size_t smem = simt_mlp_smem_bytes(sp.KP);
  LIDF_CUDA(cudaFuncSetAttribute(k_mlp_simt, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_mlp_simt<<<(unsigned)((p->R + LIDF_SIMT_BM - 1) / LIDF_SIMT_BM), LIDF_SIMT_THREADS, smem, st>>>(a);
  LIDF_LAUNCH_CHECK();
  return LIDF_OK;
}

// -------------------------------------------------------------------------------------------------
// Ray / point vs voxel-box tests (include/lidf_aabb.h): the producers of the pair list the query path consumes.
// -------------------------------------------------------------------------------------------------
#include "lidf_aabb.cuh"

namespace {
struct AabbPlan { int64_t RB, M; int nb; float* inv; int2* tile_bid; float4* tile_tan; int* cnt; int* start; int* block_sums; int* items; int* n_items; size_t bytes; };

int plan_aabb(int64_t R, int64_t V, AabbPlan* q, char* base) {
  if (R < 0 || V < 0) return LIDF_ERR_ARG;
  q->RB = (R + AABB_TILE - 1) / AABB_TILE;
  q->M = q->RB * V;
  if (q->M >= INT_MAX - 2 || R >= ((int64_t)1 << 40)) return LIDF_ERR_UNSUPPORTED;
  const int64_t n = q->M + 1;
  q->nb = (int)((n + LIDF_SCAN_BLOCK * LIDF_SCAN_ITEMS - 1) / (LIDF_SCAN_BLOCK * LIDF_SCAN_ITEMS));
  Bump b{base, 0};
  q->inv = b.take<float>((size_t)(R > 0 ? R : 1) * 3);
  q->tile_bid = b.take<int2>((size_t)(q->RB > 0 ? q->RB : 1));
  q->tile_tan = b.take<float4>((size_t)(q->RB > 0 ? q->RB : 1));
  q->cnt = b.take<int>((size_t)n);
  q->start = b.take<int>((size_t)n);
  q->block_sums = b.take<int>((size_t)q->nb);
  q->items = b.take<int>((size_t)(q->M > 0 ? q->M : 1));
  q->n_items = b.take<int>(1);
  q->bytes = b.off + 256;
  return LIDF_OK;
}

int persistent_blocks(int64_t items, int per_sm) {
  int dev = 0, sms = 0;                                      // per call: the current device may differ between calls
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0)
    sms = 148;
  const int64_t g = (int64_t)sms * per_sm;                   // a multiple of the SM count (148 on B200)
  return (int)(items < g ? (items > 0 ? items : 1) : g);
}

int aabb_ray_prep(const AabbPlan& q, const float* ray_dir, const int32_t* ray_bid, int64_t R, cudaStream_t st) {
  k_aabb_ray_prep<<<(unsigned)q.RB, AABB_THREADS, 0, st>>>(ray_dir, ray_bid, R, q.inv, q.tile_bid, q.tile_tan);
  LIDF_LAUNCH_CHECK();
  return LIDF_OK;
}
}  // namespace

extern "C" size_t lidf_ray_aabb_workspace_bytes(int64_t R, int64_t V) {
  AabbPlan q;
  if (plan_aabb(R, V, &q, nullptr) != LIDF_OK) return 0;
  return q.bytes;
}

extern "C" int lidf_ray_aabb_forward(const float* ray_dir, const float* voxel_bound, const int32_t* ray_bid,
                                     const int32_t* voxel_bid, int64_t R, int64_t V, int32_t* mask, float* dist,
                                     void* workspace, size_t workspace_bytes, lidf_stream_t stream) {
  AabbPlan q;
  int rc = plan_aabb(R, V, &q, (char*)workspace);
  if (rc) return rc;
  if (R == 0 || V == 0) return LIDF_OK;
  if (!ray_dir || !voxel_bound || !ray_bid || !voxel_bid || !mask || !dist || !workspace) return LIDF_ERR_NULL;
  if (workspace_bytes < q.bytes) return LIDF_ERR_WORKSPACE;
  cudaStream_t st = stream;
  if ((rc = aabb_ray_prep(q, ray_dir, ray_bid, R, st))) return rc;
  // 16-byte accesses need R % 4 == 0 (every mask/dist row then starts 16-byte aligned) and aligned base pointers
  const bool vec = R % 4 == 0 && ((uintptr_t)mask % 16 == 0) && ((uintptr_t)dist % 16 == 0) && ((uintptr_t)ray_bid % 16 == 0);
  if (vec)
    k_aabb_dense<true><<<persistent_blocks(q.M, 8), AABB_THREADS, 0, st>>>(q.inv, voxel_bound, ray_bid, voxel_bid, q.tile_bid,
                                                                            q.tile_tan, R, q.RB, q.M, mask,
                                                                            reinterpret_cast<float2*>(dist));
  else
    k_aabb_dense<false><<<persistent_blocks(q.M, 8), AABB_THREADS, 0, st>>>(q.inv, voxel_bound, ray_bid, voxel_bid, q.tile_bid,
                                                                             q.tile_tan, R, q.RB, q.M, mask,
                                                                             reinterpret_cast<float2*>(dist));
  LIDF_LAUNCH_CHECK();
  return LIDF_OK;
}

extern "C" int lidf_ray_aabb_pairs_count(const float* ray_dir, const float* voxel_bound, const int32_t* ray_bid,
                                         const int32_t* voxel_bid, int64_t R, int64_t V, void* workspace,
                                         size_t workspace_bytes, int64_t* n_pairs_host, lidf_stream_t stream) {
  if (!n_pairs_host) return LIDF_ERR_NULL;
  *n_pairs_host = 0;
  AabbPlan q;
  int rc = plan_aabb(R, V, &q, (char*)workspace);
  if (rc) return rc;
  if (R == 0 || V == 0) return LIDF_OK;
  if (!ray_dir || !voxel_bound || !ray_bid || !voxel_bid || !workspace) return LIDF_ERR_NULL;
  if (workspace_bytes < q.bytes) return LIDF_ERR_WORKSPACE;
  cudaStream_t st = stream;
  if ((rc = aabb_ray_prep(q, ray_dir, ray_bid, R, st))) return rc;
  const int64_t n = q.M + 1;
  LIDF_CUDA(cudaMemsetAsync(q.cnt, 0, sizeof(int) * (size_t)n, st));
  LIDF_CUDA(cudaMemsetAsync(q.n_items, 0, sizeof(int), st));
  k_aabb_items<<<(unsigned)((q.M + 255) / 256), 256, 0, st>>>(voxel_bound, voxel_bid, q.tile_bid, q.tile_tan, q.RB, q.M, q.items,
                                                               q.n_items);
  LIDF_LAUNCH_CHECK();
  k_aabb_count<<<persistent_blocks(q.M, 8), AABB_THREADS, 0, st>>>(q.inv, voxel_bound, ray_bid, voxel_bid, q.items, q.n_items, R,
                                                                    q.RB, q.cnt);
  LIDF_LAUNCH_CHECK();
  k_scan_partial<<<q.nb, LIDF_SCAN_BLOCK, 0, st>>>(q.cnt, n, q.start, q.block_sums);
  LIDF_LAUNCH_CHECK();
  k_scan_blocksums<<<1, 1024, 0, st>>>(q.block_sums, q.nb);
  LIDF_LAUNCH_CHECK();
  k_scan_add<<<q.nb, LIDF_SCAN_BLOCK, 0, st>>>(q.start, n, q.block_sums, nullptr);
  LIDF_LAUNCH_CHECK();
  int total = 0;
  LIDF_CUDA(cudaMemcpyAsync(&total, q.start + q.M, sizeof(int), cudaMemcpyDeviceToHost, st));
  LIDF_CUDA(cudaStreamSynchronize(st));
  if (total < 0) return LIDF_ERR_UNSUPPORTED;              // >= 2^31 pairs
  *n_pairs_host = total;
  return LIDF_OK;
}

extern "C" int lidf_ray_aabb_pairs_fill(const float* ray_dir, const float* voxel_bound, const int32_t* ray_bid,
                                        const int32_t* voxel_bid, int64_t R, int64_t V, void* workspace,
                                        size_t workspace_bytes, int64_t P, int64_t* pair_vox, int64_t* pair_ray,
                                        float* pair_dist, lidf_stream_t stream) {
  (void)ray_dir;
  AabbPlan q;
  int rc = plan_aabb(R, V, &q, (char*)workspace);
  if (rc) return rc;
  if (P < 0) return LIDF_ERR_ARG;
  if (R == 0 || V == 0 || P == 0) return LIDF_OK;
  if (!voxel_bound || !ray_bid || !voxel_bid || !workspace || !pair_vox || !pair_ray || !pair_dist) return LIDF_ERR_NULL;
  if (workspace_bytes < q.bytes) return LIDF_ERR_WORKSPACE;
  cudaStream_t st = stream;
  k_aabb_fill<<<persistent_blocks(q.M, 8), AABB_THREADS, 0, st>>>(q.inv, voxel_bound, ray_bid, voxel_bid, q.items, q.n_items, R,
                                                                   q.RB, q.start, pair_vox, pair_ray,
                                                                   reinterpret_cast<float2*>(pair_dist));
  LIDF_LAUNCH_CHECK();
  return LIDF_OK;
}

namespace {
struct AabbRmPlan { int* cnt; int* start; int* block_sums; int nb; size_t bytes; };
int plan_aabb_rm(int64_t R, int64_t V, AabbRmPlan* q, char* base) {
  if (R < 0 || V < 0) return LIDF_ERR_ARG;
  if (R >= INT_MAX - 2 || V >= INT_MAX - 2) return LIDF_ERR_UNSUPPORTED;
  const int64_t n = R + 1;
  q->nb = (int)((n + LIDF_SCAN_BLOCK * LIDF_SCAN_ITEMS - 1) / (LIDF_SCAN_BLOCK * LIDF_SCAN_ITEMS));
  Bump b{base, 0};
  q->cnt = b.take<int>((size_t)n);
  q->start = b.take<int>((size_t)n);
  q->block_sums = b.take<int>((size_t)q->nb);
  q->bytes = b.off + 256;
  return LIDF_OK;
}
}  // namespace

extern "C" size_t lidf_ray_aabb_ray_major_workspace_bytes(int64_t R, int64_t V) {
  AabbRmPlan q;
  if (plan_aabb_rm(R, V, &q, nullptr) != LIDF_OK) return 0;
  return q.bytes;
}

extern "C" int lidf_ray_aabb_pairs_ray_major_count(const float* ray_dir, const float* voxel_bound, const int32_t* ray_bid,
                                                   const int32_t* voxel_bid, int64_t R, int64_t V, void* workspace,
                                                   size_t workspace_bytes, int64_t* n_pairs_host, lidf_stream_t stream) {
  if (!n_pairs_host) return LIDF_ERR_NULL;
  *n_pairs_host = 0;
  AabbRmPlan q;
  int rc = plan_aabb_rm(R, V, &q, (char*)workspace);
  if (rc) return rc;
  if (R == 0) return LIDF_OK;
  if (!workspace) return LIDF_ERR_NULL;
  if (workspace_bytes < q.bytes) return LIDF_ERR_WORKSPACE;
  cudaStream_t st = stream;
  const int64_t n = R + 1;
  LIDF_CUDA(cudaMemsetAsync(q.cnt, 0, sizeof(int) * (size_t)n, st));
  if (V > 0) {
    if (!ray_dir || !voxel_bound || !ray_bid || !voxel_bid) return LIDF_ERR_NULL;
    k_aabb_ray_major<false><<<(unsigned)((R + AABB_RM_THREADS - 1) / AABB_RM_THREADS), AABB_RM_THREADS, 0, st>>>(
        ray_dir, voxel_bound, ray_bid, voxel_bid, R, V, q.cnt, nullptr, nullptr, nullptr, nullptr);
    LIDF_LAUNCH_CHECK();
  }
  k_scan_partial<<<q.nb, LIDF_SCAN_BLOCK, 0, st>>>(q.cnt, n, q.start, q.block_sums);
  LIDF_LAUNCH_CHECK();
  k_scan_blocksums<<<1, 1024, 0, st>>>(q.block_sums, q.nb);
  LIDF_LAUNCH_CHECK();
  k_scan_add<<<q.nb, LIDF_SCAN_BLOCK, 0, st>>>(q.start, n, q.block_sums, nullptr);
  LIDF_LAUNCH_CHECK();
  int total = 0;
  LIDF_CUDA(cudaMemcpyAsync(&total, q.start + R, sizeof(int), cudaMemcpyDeviceToHost, st));
  LIDF_CUDA(cudaStreamSynchronize(st));
  if (total < 0) return LIDF_ERR_UNSUPPORTED;              // >= 2^31 pairs
  *n_pairs_host = total;
  return LIDF_OK;
}

extern "C" int lidf_ray_aabb_pairs_ray_major_fill(const float* ray_dir, const float* voxel_bound, const int32_t* ray_bid,
                                                  const int32_t* voxel_bid, int64_t R, int64_t V, void* workspace,
                                                  size_t workspace_bytes, int64_t P, int64_t* pair_vox, int64_t* pair_ray,
                                                  float* pair_dist, int32_t* ray_start, lidf_stream_t stream) {
  AabbRmPlan q;
  int rc = plan_aabb_rm(R, V, &q, (char*)workspace);
  if (rc) return rc;
  if (P < 0) return LIDF_ERR_ARG;
  cudaStream_t st = stream;
  if (R == 0) {
    if (ray_start) LIDF_CUDA(cudaMemsetAsync(ray_start, 0, sizeof(int), st));
    return LIDF_OK;
  }
  if (!workspace) return LIDF_ERR_NULL;
  if (workspace_bytes < q.bytes) return LIDF_ERR_WORKSPACE;
  if (ray_start) LIDF_CUDA(cudaMemcpyAsync(ray_start, q.start, sizeof(int) * (size_t)(R + 1), cudaMemcpyDeviceToDevice, st));
  if (V == 0 || P == 0) return LIDF_OK;
  if (!ray_dir || !voxel_bound || !ray_bid || !voxel_bid || !pair_vox || !pair_ray || !pair_dist) return LIDF_ERR_NULL;
  k_aabb_ray_major<true><<<(unsigned)((R + AABB_RM_THREADS - 1) / AABB_RM_THREADS), AABB_RM_THREADS, 0, st>>>(
      ray_dir, voxel_bound, ray_bid, voxel_bid, R, V, nullptr, q.start, pair_vox, pair_ray, reinterpret_cast<float2*>(pair_dist));
  LIDF_LAUNCH_CHECK();
  return LIDF_OK;
}

extern "C" int lidf_pcl_aabb_forward(const float* pcl_pos, const float* voxel_bound, const int32_t* pcl_bid,
                                     const int32_t* voxel_bid, int64_t N, int64_t V, int32_t* mask, lidf_stream_t stream) {
  if (N < 0 || V < 0) return LIDF_ERR_ARG;
  if (N == 0 || V == 0) return LIDF_OK;
  if (!pcl_pos || !voxel_bound || !pcl_bid || !voxel_bid || !mask) return LIDF_ERR_NULL;
  const int64_t NB = (N + AABB_TILE - 1) / AABB_TILE, M = NB * V;
  const bool vec = N % 4 == 0 && ((uintptr_t)mask % 16 == 0) && ((uintptr_t)pcl_pos % 16 == 0) && ((uintptr_t)pcl_bid % 16 == 0);
  if (vec)
    k_pcl_dense<true><<<persistent_blocks(M, 8), AABB_THREADS, 0, (cudaStream_t)stream>>>(pcl_pos, voxel_bound, pcl_bid, voxel_bid, N, NB, M, mask);
  else
    k_pcl_dense<false><<<persistent_blocks(M, 8), AABB_THREADS, 0, (cudaStream_t)stream>>>(pcl_pos, voxel_bound, pcl_bid, voxel_bid, N, NB, M, mask);
  LIDF_LAUNCH_CHECK();
  return LIDF_OK;
}

extern "C" int lidf_pcl_aabb_pair_label(const float* pcl_pos, const float* voxel_bound, const int32_t* pcl_bid,
                                        const int32_t* voxel_bid, int64_t N, int64_t V, const int64_t* pair_vox,
                                        const int64_t* pair_ray, int64_t P, float* label, lidf_stream_t stream) {
  if (N < 0 || V < 0 || P < 0) return LIDF_ERR_ARG;
  if (P == 0) return LIDF_OK;
  if (!pcl_pos || !voxel_bound || !pcl_bid || !voxel_bid || !pair_vox || !pair_ray || !label) return LIDF_ERR_NULL;
  k_pcl_pair_label<<<(unsigned)((P + 255) / 256), 256, 0, (cudaStream_t)stream>>>(pcl_pos, voxel_bound, pcl_bid, voxel_bid, N, V,
                                                                                  pair_vox, pair_ray, P, label);
  LIDF_LAUNCH_CHECK();
  return LIDF_OK;
}

extern "C" int lidf_pcl_aabb_end_voxel(const float* pcl_pos, const float* voxel_bound, const int32_t* pcl_bid,
                                       const int32_t* voxel_bid, int64_t N, int64_t V, int64_t* end_voxel_id,
                                       lidf_stream_t stream) {
  if (N < 0 || V < 0) return LIDF_ERR_ARG;
  if (N == 0 || V == 0) return LIDF_OK;
  if (!pcl_pos || !voxel_bound || !pcl_bid || !voxel_bid || !end_voxel_id) return LIDF_ERR_NULL;
  k_pcl_end_voxel<<<(unsigned)((N + AABB_THREADS - 1) / AABB_THREADS), AABB_THREADS, 0, (cudaStream_t)stream>>>(
      pcl_pos, voxel_bound, pcl_bid, voxel_bid, N, V, end_voxel_id);
  LIDF_LAUNCH_CHECK();
  return LIDF_OK;
}

// ---- voxelisation ------------------------------------------------------------------------------------------------------
namespace {
struct VoxPlan { int64_t ncell; int nb_c, nb_p; int* cell_flag; int* cell_rank; int* inside; int* pt_rank; int* bs_c; int* bs_p; size_t bytes; };
int plan_vox(int64_t Np, int B, int nx, int ny, int nz, VoxPlan* q, char* base) {
  if (Np < 0 || B <= 0 || nx <= 0 || ny <= 0 || nz <= 0) return LIDF_ERR_ARG;
  q->ncell = (int64_t)B * nx * ny * nz;
  if (q->ncell >= INT_MAX - 2 || Np >= INT_MAX - 2) return LIDF_ERR_UNSUPPORTED;
  const int64_t per = LIDF_SCAN_BLOCK * LIDF_SCAN_ITEMS;
  q->nb_c = (int)((q->ncell + 1 + per - 1) / per);
  q->nb_p = (int)((Np + 1 + per - 1) / per);
  Bump b{base, 0};
  q->cell_flag = b.take<int>((size_t)q->ncell + 1);
  q->cell_rank = b.take<int>((size_t)q->ncell + 1);
  q->inside = b.take<int>((size_t)Np + 1);
  q->pt_rank = b.take<int>((size_t)Np + 1);
  q->bs_c = b.take<int>((size_t)q->nb_c);
  q->bs_p = b.take<int>((size_t)q->nb_p);
  q->bytes = b.off + 256;
  return LIDF_OK;
}
VoxGrid make_grid(int B, float x0, float x1, float x2, float part, float half_part, int nx, int ny, int nz) {
  VoxGrid g;
  g.xmin[0] = x0; g.xmin[1] = x1; g.xmin[2] = x2; g.crop = part; g.half_crop = half_part;
  g.n[0] = nx; g.n[1] = ny; g.n[2] = nz; g.B = B;
  return g;
}
int scan_ints(const int* in, int64_t n, int* out, int* block_sums, int nb, cudaStream_t st) {
  k_scan_partial<<<nb, LIDF_SCAN_BLOCK, 0, st>>>(in, n, out, block_sums);
  LIDF_LAUNCH_CHECK();
  k_scan_blocksums<<<1, 1024, 0, st>>>(block_sums, nb);
  LIDF_LAUNCH_CHECK();
  k_scan_add<<<nb, LIDF_SCAN_BLOCK, 0, st>>>(out, n, block_sums, nullptr);
  LIDF_LAUNCH_CHECK();
  return LIDF_OK;
}
}  // namespace

extern "C" size_t lidf_voxelize_workspace_bytes(int64_t Np, int32_t B, int32_t nx, int32_t ny, int32_t nz) {
  VoxPlan q;
  if (plan_vox(Np, B, nx, ny, nz, &q, nullptr) != LIDF_OK) return 0;
  return q.bytes;
}

extern "C" int lidf_voxelize_count(const float* xyz, const int64_t* bid, int64_t Np, int32_t B, float x0, float x1, float x2,
                                   float part, float half_part, int32_t nx, int32_t ny, int32_t nz, void* ws, size_t ws_bytes,
                                   int64_t* n_vox_host, int64_t* n_inside_host, lidf_stream_t stream) {
  if (!n_vox_host || !n_inside_host) return LIDF_ERR_NULL;
  *n_vox_host = 0; *n_inside_host = 0;
  VoxPlan q;
  int rc = plan_vox(Np, B, nx, ny, nz, &q, (char*)ws);
  if (rc) return rc;
  if (!(part > 0.f)) return LIDF_ERR_ARG;
  if (Np == 0) return LIDF_OK;
  if (!xyz || !bid || !ws) return LIDF_ERR_NULL;
  if (ws_bytes < q.bytes) return LIDF_ERR_WORKSPACE;
  cudaStream_t st = stream;
  const VoxGrid g = make_grid(B, x0, x1, x2, part, half_part, nx, ny, nz);
  LIDF_CUDA(cudaMemsetAsync(q.cell_flag, 0, sizeof(int) * (size_t)(q.ncell + 1), st));
  LIDF_CUDA(cudaMemsetAsync(q.inside + Np, 0, sizeof(int), st));
  k_vox_mark<<<(unsigned)((Np + 255) / 256), 256, 0, st>>>(xyz, bid, Np, g, q.cell_flag, q.inside);
  LIDF_LAUNCH_CHECK();
  if ((rc = scan_ints(q.cell_flag, q.ncell + 1, q.cell_rank, q.bs_c, q.nb_c, st))) return rc;
  if ((rc = scan_ints(q.inside, Np + 1, q.pt_rank, q.bs_p, q.nb_p, st))) return rc;
  int tot[2] = {0, 0};
  LIDF_CUDA(cudaMemcpyAsync(&tot[0], q.cell_rank + q.ncell, sizeof(int), cudaMemcpyDeviceToHost, st));
  LIDF_CUDA(cudaMemcpyAsync(&tot[1], q.pt_rank + Np, sizeof(int), cudaMemcpyDeviceToHost, st));
  LIDF_CUDA(cudaStreamSynchronize(st));
  *n_vox_host = tot[0]; *n_inside_host = tot[1];
  return LIDF_OK;
}

extern "C" int lidf_voxelize_fill(const float* xyz, const int64_t* bid, int64_t Np, int32_t B, float x0, float x1, float x2,
                                  float part, float half_part, int32_t nx, int32_t ny, int32_t nz, void* ws, size_t ws_bytes,
                                  int64_t* occ, float* bound, int64_t* revidx, int64_t* pid, float* rel, lidf_stream_t stream) {
  VoxPlan q;
  int rc = plan_vox(Np, B, nx, ny, nz, &q, (char*)ws);
  if (rc) return rc;
  if (Np == 0) return LIDF_OK;
  if (!xyz || !bid || !ws || !occ || !bound || !revidx || !pid || !rel) return LIDF_ERR_NULL;
  if (ws_bytes < q.bytes) return LIDF_ERR_WORKSPACE;
  cudaStream_t st = stream;
  const VoxGrid g = make_grid(B, x0, x1, x2, part, half_part, nx, ny, nz);
  k_vox_fill_voxels<<<(unsigned)((q.ncell + 255) / 256), 256, 0, st>>>(q.cell_flag, q.cell_rank, q.ncell, g, occ, bound);
  LIDF_LAUNCH_CHECK();
  k_vox_fill_points<<<(unsigned)((Np + 255) / 256), 256, 0, st>>>(xyz, bid, Np, g, q.inside, q.pt_rank, q.cell_rank, revidx, pid, rel);
  LIDF_LAUNCH_CHECK();
  return LIDF_OK;
}

// ---- PointNet2Stage forward (include/lidf_pointnet.h) ------------------------------------------------------------------
#include "lidf_pointnet.cuh"

namespace {
struct PnPlan { float* vmax1; float* vf1; float* vmax2; int* err; uint8_t* wpack; size_t bytes; };
int plan_pn(int64_t N, int64_t V, PnPlan* q, char* base) {
  if (N < 0 || V < 0) return LIDF_ERR_ARG;
  if (V >= INT_MAX / 256 || N >= ((int64_t)1 << 40)) return LIDF_ERR_UNSUPPORTED;
  Bump b{base, 0};
  const size_t v = (size_t)(V > 0 ? V : 1);
  q->vmax1 = b.take<float>(v * PN_C1);
  q->vmax2 = b.take<float>(v * PN_C2);      // contiguous with vmax1: one memset clears both
  q->vf1 = b.take<float>(v * PN_C1);
  q->err = b.take<int>(1);
  q->wpack = b.take<uint8_t>((size_t)2 * PNT_KSTEPS * TC_CHUNK_BYTES);     // point_lin3 / point_lin4 as bf16 hi | lo MMA chunks
  q->bytes = b.off + 256;
  return LIDF_OK;
}
}  // namespace

extern "C" size_t lidf_pointnet_workspace_bytes(int64_t N, int64_t V) {
  PnPlan q;
  if (plan_pn(N, V, &q, nullptr) != LIDF_OK) return 0;
  return q.bytes;
}

extern "C" int lidf_pointnet_forward(const LidfPointNet* wp, const float* inp, const int64_t* idx, int64_t N, int64_t V,
                                     float* out, void* ws, size_t ws_bytes, lidf_stream_t stream) {
  return lidf_pointnet_forward_impl(wp, inp, idx, N, V, out, ws, ws_bytes, LIDF_MLP_AUTO, stream);
}

extern "C" int lidf_pointnet_forward_impl(const LidfPointNet* wp, const float* inp, const int64_t* idx, int64_t N, int64_t V,
                                          float* out, void* ws, size_t ws_bytes, int32_t mlp_impl, lidf_stream_t stream) {
  if (!wp) return LIDF_ERR_NULL;
  if (mlp_impl != LIDF_MLP_AUTO && mlp_impl != LIDF_MLP_SIMT_FP32 && mlp_impl != LIDF_MLP_TC_BF16X3) return LIDF_ERR_ARG;
  const bool use_tc = mlp_impl != LIDF_MLP_SIMT_FP32;
  if (use_tc && !tc_device_ok()) return LIDF_ERR_NO_SM100;
  PnPlan q;
  int rc = plan_pn(N, V, &q, (char*)ws);
  if (rc) return rc;
  if (V == 0) return LIDF_OK;
  if (!out || !ws || (N > 0 && (!inp || !idx))) return LIDF_ERR_NULL;
  if (ws_bytes < q.bytes) return LIDF_ERR_WORKSPACE;
  PnWeights w{wp->point_lin1_w, wp->point_lin1_b, wp->point_lin2_w, wp->point_lin2_b, wp->vox_lin1_w, wp->vox_lin1_b,
              wp->point_lin3_w, wp->point_lin3_b, wp->point_lin4_w, wp->point_lin4_b, wp->vox_lin2_w, wp->vox_lin2_b};
  const float* const* all = &w.w_p1;
  for (int i = 0; i < 12; ++i)
    if (!all[i]) return LIDF_ERR_NULL;
  cudaStream_t st = stream;
  LIDF_CUDA(cudaMemsetAsync(q.vmax1, 0, (size_t)((char*)q.vf1 - (char*)q.vmax1), st));     // vmax1 and vmax2
  LIDF_CUDA(cudaMemsetAsync(q.err, 0, sizeof(int), st));
  const size_t sm1 = pn_stage1_smem(), sm2 = pn_stage2_smem();
  LIDF_CUDA(cudaFuncSetAttribute(k_pn_stage1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm1));
  LIDF_CUDA(cudaFuncSetAttribute(k_pn_stage2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm2));
  if (N > 0) {
    const int64_t t1 = (N + PN1_THREADS - 1) / PN1_THREADS;
    k_pn_stage1<<<persistent_blocks(t1, 2), PN1_THREADS, sm1, st>>>(inp, idx, N, V, w, q.vmax1, q.err);
    LIDF_LAUNCH_CHECK();
  }
  k_pn_vox<PN_C1><<<(unsigned)((V + 7) / 8), PN_C1, 0, st>>>(q.vmax1, w.w_v1, w.b_v1, V, q.vf1);
  LIDF_LAUNCH_CHECK();
  if (N > 0 && use_tc) {
    // the two 128 -> 128 layers on the tensor cores (split bf16, fp32 accumulate), weights resident in shared memory
    k_pn_pack_tc<<<(PNT_KSTEPS * 2048 + 255) / 256, 256, 0, st>>>(w.w_p3, q.wpack);
    LIDF_LAUNCH_CHECK();
    k_pn_pack_tc<<<(PNT_KSTEPS * 2048 + 255) / 256, 256, 0, st>>>(w.w_p4, q.wpack + (size_t)PNT_KSTEPS * TC_CHUNK_BYTES);
    LIDF_LAUNCH_CHECK();
    const size_t smt = sizeof(PnTcSmem) + 1024;
    LIDF_CUDA(cudaFuncSetAttribute(k_pn_stage2_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smt));
    k_pn_stage2_tc<<<persistent_blocks((N + 127) / 128, 1), PNT_THREADS, smt, st>>>(inp, idx, N, V, w, q.wpack, q.vf1, q.vmax2);
    LIDF_LAUNCH_CHECK();
  } else if (N > 0) {
    const int64_t t2 = (N + PN2_TP - 1) / PN2_TP;
    k_pn_stage2<<<persistent_blocks(t2, 1), PN2_THREADS, sm2, st>>>(inp, idx, N, V, w, q.vf1, q.vmax2);
    LIDF_LAUNCH_CHECK();
  }
  k_pn_vox<PN_C2><<<(unsigned)((V + 7) / 8), PN_C2, 0, st>>>(q.vmax2, w.w_v2, w.b_v2, V, out);
  LIDF_LAUNCH_CHECK();
  return LIDF_OK;
}
