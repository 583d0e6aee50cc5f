// lidf_simt.cuh -- fp32 FFMA engine for the LIDF decoders (LIDF_MLP_SIMT_FP32) and the shared
// per-ray / per-voxel layer-1 precompute ("row prep") used by both engines.
//
// Algebra (exact up to fp32 summation order): linear_1 acts on a concat, so
//   W1 x = W1[:,0:128] vox[v]  +  W1[:,128:256] rgb[r] + W1[:,dir] PE(dir_r) + b1   +  W1[:,pos] PE(enter,leave)
//          `---- per voxel ----'  `------------------- per ray -------------------'     `------ per pair ------'
// and for IEF (implicit_net.py:133-146) offset_enc -> linear_1[:, D:D+16] has no non-linearity in between:
//   W1[:,D:] (w_enc o + b_enc) = u o + c,  u = W1[:,D:] w_enc,  c = W1[:,D:] b_enc
// so the static part of layer 1 is evaluated once and every IEF iteration adds the rank-1 term u*(o - o0).
#pragma once
#include "lidf_common.cuh"
#include "lidf_prep.cuh"

#define LIDF_SIMT_THREADS 256
#define LIDF_SIMT_BM 64
#define LIDF_KC 16

// ------------------------------------------------------------------------------------------------
// weight packing: dst[(k_off + k) * ldd + n_off + n] = w[n * ldw + col0 + k]   (k < K), k-major copies
// ------------------------------------------------------------------------------------------------
__global__ void k_pack_wt(const float* __restrict__ w, int ldw, int col0, int K, int N, float* __restrict__ dst,
                          int ldd, int k_off, int n_off) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= K * N) return;
  int n = idx / K, k = idx % K;   // consecutive threads read consecutive k (coalesced source rows)
  dst[(size_t)(k_off + k) * ldd + n_off + n] = w[(size_t)n * ldw + col0 + k];
}

// layer-1 bias for the per-ray term and the IEF rank-1 vector:
//   bias_out[n] = b1[n] (+ c_n + u_n * o0),  u_out[n] = u_n
__global__ void k_pack_bias1(const float* __restrict__ w1, int ldw, int D, const float* __restrict__ b1,
                             const float* __restrict__ w_enc, const float* __restrict__ b_enc, int is_ief, float o0,
                             float* __restrict__ bias_out, float* __restrict__ u_out) {
  int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= LIDF_H1) return;
  float b = b1[n], u = 0.f, c = 0.f;
  if (is_ief) {
    for (int m = 0; m < LIDF_IEF_ENC; ++m) {
      float w = w1[(size_t)n * ldw + D + m];
      u += w * w_enc[m];
      c += w * b_enc[m];
    }
    b += c + u * o0;
  }
  bias_out[n] = b;
  if (u_out) u_out[n] = u;
}

// ------------------------------------------------------------------------------------------------
// Block GEMM on an smem-resident X tile: acc[i][j] = sum_k Xs[row_i][k] * Wt[k][col_j]
//   thread (rg = tid/32, cl = tid%32) owns rows rg*8 .. rg*8+7 and cols cl + 32 j  (j < N/32)
//   Wt is k-major in global memory (L2 resident), staged through Wtile[KC][N] in smem.
// ------------------------------------------------------------------------------------------------
template <int N>
__device__ __forceinline__ void simt_gemm_tile(const float* __restrict__ Xs, int ldx, int K,
                                               const float* __restrict__ Wt, int ldw, float* __restrict__ Wtile,
                                               float (&acc)[8][N / 32]) {
  const int tid = threadIdx.x, rg = tid >> 5, cl = tid & 31;
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < N / 32; ++j) acc[i][j] = 0.f;
  for (int k0 = 0; k0 < K; k0 += LIDF_KC) {
    __syncthreads();
    for (int idx = tid; idx < LIDF_KC * N / 4; idx += LIDF_SIMT_THREADS) {
      const int kk = idx / (N / 4), c4 = idx % (N / 4);
      reinterpret_cast<float4*>(Wtile)[idx] =
          __ldg(reinterpret_cast<const float4*>(Wt + (size_t)(k0 + kk) * ldw) + c4);
    }
    __syncthreads();
#pragma unroll
    for (int kk4 = 0; kk4 < LIDF_KC; kk4 += 4) {
      float4 xv[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) xv[i] = *reinterpret_cast<const float4*>(Xs + (size_t)(rg * 8 + i) * ldx + k0 + kk4);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float w[N / 32];
#pragma unroll
        for (int j = 0; j < N / 32; ++j) w[j] = Wtile[(kk4 + q) * N + cl + 32 * j];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float x = q == 0 ? xv[i].x : q == 1 ? xv[i].y : q == 2 ? xv[i].z : xv[i].w;
#pragma unroll
          for (int j = 0; j < N / 32; ++j) acc[i][j] = fmaf(x, w[j], acc[i][j]);
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Row prep: out[row][n] = bias[n] + sum_k X[row][k] Wt[k][n],  X[row] = [featA[row](128) | featB[row](128) | PE(dir[row])]
//   rays   : featA = ROI feature, dirs = ray dir            -> T[R][Ntot]      (Ntot = 256 per decoder)
//   voxels : featA = occ_voxel_feat                         -> A_v[V][Ntot]    (SIMT engine only)
//   refine : featA = voxel_feat_end, featB = rgb_feat_end, dirs
// grid (ceil(rows/64), Ntot/128)
// ------------------------------------------------------------------------------------------------
struct RowPrepArgs {
  const float* featA; const float* featB; const float* dirs;
  int64_t rows;
  int multires_views, pos_encode;
  const float* Wt; int Kpad; int Ntot;      // Wt [Kpad][Ntot], rows beyond the real K are zero
  const float* bias;                        // [Ntot] or NULL
  float* out;                               // [rows][Ntot]
};

__global__ void __launch_bounds__(LIDF_SIMT_THREADS) k_rowprep(const RowPrepArgs a) {
  extern __shared__ __align__(16) float smem[];
  float* Xs = smem;                              // [64][Kpad]
  float* Wtile = smem + LIDF_SIMT_BM * a.Kpad;   // [16][128]
  const int tid = threadIdx.x;
  const int64_t row0 = (int64_t)blockIdx.x * LIDF_SIMT_BM;
  const int n0 = blockIdx.y * 128;
  const int ldx = a.Kpad;
  int koff = 0;
  for (int f = 0; f < 2; ++f) {
    const float* src = f == 0 ? a.featA : a.featB;
    if (!src) continue;
    for (int idx = tid; idx < LIDF_SIMT_BM * 32; idx += LIDF_SIMT_THREADS) {
      const int r = idx >> 5, c4 = idx & 31;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (row0 + r < a.rows) v = __ldg(reinterpret_cast<const float4*>(src + (size_t)(row0 + r) * 128) + c4);
      *reinterpret_cast<float4*>(Xs + (size_t)r * ldx + koff + c4 * 4) = v;
    }
    koff += 128;
  }
  // zero the tail (PE slots + padding), then PE(dir)
  for (int idx = tid; idx < LIDF_SIMT_BM * (ldx - koff); idx += LIDF_SIMT_THREADS) {
    const int r = idx / (ldx - koff), c = idx % (ldx - koff);
    Xs[(size_t)r * ldx + koff + c] = 0.f;
  }
  __syncthreads();
  if (a.dirs && tid < LIDF_SIMT_BM && row0 + tid < a.rows) {
    const float* d = a.dirs + (size_t)(row0 + tid) * 3;
    float* xr = Xs + (size_t)tid * ldx + koff;
    lidf_pe3(d[0], d[1], d[2], a.multires_views, a.pos_encode, [&](int j, float v) { xr[j] = v; });
  }
  float acc[8][4];
  simt_gemm_tile<128>(Xs, ldx, a.Kpad, a.Wt + n0, a.Ntot, Wtile, acc);
  const int rg = tid >> 5, cl = tid & 31;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int64_t row = row0 + rg * 8 + i;
    if (row >= a.rows) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + cl + 32 * j;
      a.out[(size_t)row * a.Ntot + n] = acc[i][j] + (a.bias ? a.bias[n] : 0.f);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Fused per-row decoder, fp32.  One block = 64 rows (pairs in ray-major order, or rays for the refine tail).
// ------------------------------------------------------------------------------------------------
struct SimtDecoder {
  int kind, n_pass, use_sigmoid;      // n_pass = 1 (IMNet) or n_iter (IEF)
  const float* Wt_pe;                 // [KP][256]
  const float* Wt2; const float* b2;  // [256][128], [128]
  const float* Wt3; const float* b3;  // [128][64], [64]
  const float* w4; const float* b4;   // [64], [1]
  const float* u;                     // [256] IEF rank-1 vector (NULL for IMNet)
  float o0;                           // IEF init offset
  const float* addA; int ldA; int offA;   // per-voxel term A_v[vox][offA + n] (NULL in refine mode)
  const float* addB; int ldB; int offB;   // per-ray   term T[ray][offB + n]
  float* out;                             // [rows] written at the original pair index (NULL -> not stored)
};

struct SimtMlpArgs {
  int64_t rows;                 // P (pairs) or R (refine)
  int refine;                   // 0: LIDF pairs, 1: RefineNet rays
  // pair mode
  const int* perm;              // sorted slot -> original pair (NULL = identity)
  const int64_t* pair_vox; const int64_t* pair_ray;
  const float* pair_dist; const float* dense_dist; int64_t R; int64_t V;
  const float* ray_dir; const float* voxel_bound;
  // refine mode
  const float* pos_in; const float* center_in;
  int rel, pos_encode, multires, KP;
  int n_dec;
  SimtDecoder dec[2];           // LIDF: [0] = offset_dec, [1] = prob_dec; refine: [0] = offset_dec
  float r0, r1, scale;          // pos = enter + ((off*(r1-r0)+r0)*scale_a)*scale_b * dir
  float scale2;
  float* pos_out;               // pair_pred_pos [P,3] / pred_pos_refine [R,3]
  float* o_iter;                // [n_pass-1][rows] or NULL: dec[0]'s running IEF offset after every iteration but the last
};

__global__ void __launch_bounds__(LIDF_SIMT_THREADS) k_mlp_simt(const SimtMlpArgs a) {
  extern __shared__ __align__(16) float smem[];
  const int KP = a.KP;
  float* Xpe = smem;                                  // [64][KP]
  float* H1pre = Xpe + LIDF_SIMT_BM * KP;             // [64][256]
  float* H1 = H1pre + LIDF_SIMT_BM * LIDF_H1;         // [64][256]   (H3 [64][68] aliases it)
  float* H2 = H1 + LIDF_SIMT_BM * LIDF_H1;            // [64][128]
  float* Wtile = H2 + LIDF_SIMT_BM * LIDF_H2;         // [16][256]
  float* H3 = H1;
  __shared__ int s_orig[LIDF_SIMT_BM], s_vox[LIDF_SIMT_BM], s_ray[LIDF_SIMT_BM];
  __shared__ float s_dir[LIDF_SIMT_BM][3], s_base[LIDF_SIMT_BM][3], s_pin[LIDF_SIMT_BM][2][3];
  __shared__ float s_o[LIDF_SIMT_BM], s_res[2][LIDF_SIMT_BM];

  const int tid = threadIdx.x, rg = tid >> 5, cl = tid & 31;
  const int64_t row0 = (int64_t)blockIdx.x * LIDF_SIMT_BM;
  const int npos = a.refine ? 1 : 2;
  const int pe = lidf_pe_dim(a.multires, a.pos_encode);

  for (int idx = tid; idx < LIDF_SIMT_BM * KP; idx += LIDF_SIMT_THREADS) Xpe[idx] = 0.f;
  if (tid < LIDF_SIMT_BM) {
    const int64_t s = row0 + tid;
    int orig = -1, vox = 0, ray = 0;
    float d[3] = {0.f, 0.f, 0.f}, base[3] = {0.f, 0.f, 0.f}, p0[3] = {0.f, 0.f, 0.f}, p1[3] = {0.f, 0.f, 0.f};
    if (s < a.rows) {
      if (!a.refine) {
        orig = a.perm ? a.perm[s] : (int)s;
        vox = (int)lidf_clamp_idx(a.pair_vox[orig], a.V); ray = (int)lidf_clamp_idx(a.pair_ray[orig], a.R);
        float t0, t1;
        if (a.pair_dist) { t0 = a.pair_dist[2 * (size_t)orig]; t1 = a.pair_dist[2 * (size_t)orig + 1]; }
        else { const size_t o = ((size_t)vox * a.R + ray) * 2; t0 = a.dense_dist[o]; t1 = a.dense_dist[o + 1]; }
        float c[3] = {0.f, 0.f, 0.f};
        if (a.rel) for (int k = 0; k < 3; ++k)
          c[k] = (a.voxel_bound[(size_t)vox * 6 + k] + a.voxel_bound[(size_t)vox * 6 + 3 + k]) / 2.0f;
        for (int k = 0; k < 3; ++k) {
          d[k] = a.ray_dir[(size_t)ray * 3 + k];
          base[k] = d[k] * t0;                     // intersect_enter_pos, pipeline.py:349
          p0[k] = base[k] - c[k];
          p1[k] = d[k] * t1 - c[k];                // intersect_leave_pos (- centre), pipeline.py:350,357
        }
      } else {
        orig = (int)s; ray = (int)s;
        for (int k = 0; k < 3; ++k) {
          d[k] = a.ray_dir[(size_t)s * 3 + k];
          base[k] = a.pos_in[(size_t)s * 3 + k];
          p0[k] = base[k] - (a.rel ? a.center_in[(size_t)s * 3 + k] : 0.f);
        }
      }
    }
    s_orig[tid] = orig; s_vox[tid] = vox; s_ray[tid] = ray;
    for (int k = 0; k < 3; ++k) { s_dir[tid][k] = d[k]; s_base[tid][k] = base[k]; s_pin[tid][0][k] = p0[k]; s_pin[tid][1][k] = p1[k]; }
  }
  __syncthreads();
  {  // positional encoding: 4 threads per row split the (position, frequency) work items
    const int r = tid & 63, part = tid >> 6;
    if (s_orig[r] >= 0) {
      const int nfreq = a.pos_encode ? a.multires : 0;
      const int nwork = npos * (nfreq + 1);
      for (int w = part; w < nwork; w += 4) {
        const int pi = w / (nfreq + 1), k = w % (nfreq + 1) - 1;
        const float x0 = s_pin[r][pi][0], x1 = s_pin[r][pi][1], x2 = s_pin[r][pi][2];
        float* xr = Xpe + (size_t)r * KP + pi * pe;
        if (k < 0) { xr[0] = x0; xr[1] = x1; xr[2] = x2; }
        else {
          const float f = (float)(1 << k);
          float s0, c0, s1, c1, s2, c2;
          sincosf(x0 * f, &s0, &c0); sincosf(x1 * f, &s1, &c1); sincosf(x2 * f, &s2, &c2);
          const int b = 3 + 6 * k;
          xr[b] = s0; xr[b + 1] = s1; xr[b + 2] = s2; xr[b + 3] = c0; xr[b + 4] = c1; xr[b + 5] = c2;
        }
      }
    }
  }
  __syncthreads();

  for (int di = 0; di < a.n_dec; ++di) {
    const SimtDecoder& D = a.dec[di];
    {  // layer 1, static part: H1pre = Xpe * W_pe + A_v[vox] + T[ray]
      float acc[8][8];
      simt_gemm_tile<256>(Xpe, KP, KP, D.Wt_pe, LIDF_H1, Wtile, acc);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int r = rg * 8 + i;
        const bool ok = s_orig[r] >= 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int n = cl + 32 * j;
          float v = acc[i][j];
          if (ok) {
            if (D.addA) v += __ldg(D.addA + (size_t)s_vox[r] * D.ldA + D.offA + n);
            v += __ldg(D.addB + (size_t)s_ray[r] * D.ldB + D.offB + n);
          }
          H1pre[r * LIDF_H1 + n] = v;
        }
      }
    }
    if (tid < LIDF_SIMT_BM) s_o[tid] = D.kind == LIDF_DEC_IEF ? D.o0 : 0.f;   // running pred_offset (implicit_net.py:132)
    __syncthreads();
    for (int it = 0; it < D.n_pass; ++it) {
      for (int idx = tid; idx < LIDF_SIMT_BM * LIDF_H1; idx += LIDF_SIMT_THREADS) {
        const int r = idx >> 8, n = idx & 255;
        float v = H1pre[idx];
        if (D.u) v = fmaf(D.u[n], s_o[r] - D.o0, v);   // T already holds u*o0 + c
        H1[idx] = lidf_leaky(v);
      }
      {
        float acc[8][4];
        simt_gemm_tile<128>(H1, LIDF_H1, LIDF_H1, D.Wt2, LIDF_H2, Wtile, acc);
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int n = cl + 32 * j;
            H2[(rg * 8 + i) * LIDF_H2 + n] = lidf_leaky(acc[i][j] + D.b2[n]);
          }
      }
      {
        float acc[8][2];
        simt_gemm_tile<64>(H2, LIDF_H2, LIDF_H2, D.Wt3, LIDF_H3, Wtile, acc);
        // gemm_tile's leading __syncthreads ordered all H1 reads before this point; H3 aliases H1
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const int n = cl + 32 * j;
            H3[(rg * 8 + i) * 68 + n] = lidf_leaky(acc[i][j] + D.b3[n]);
          }
      }
      __syncthreads();
      {  // layer 4: 4 threads per row, 16 terms each
        const int r = tid >> 2, q = tid & 3;
        float p = 0.f;
#pragma unroll
        for (int j = 0; j < 16; ++j) p = fmaf(H3[r * 68 + q * 16 + j], D.w4[q * 16 + j], p);
        p += __shfl_xor_sync(0xffffffffu, p, 1);
        p += __shfl_xor_sync(0xffffffffu, p, 2);
        if (q == 0) {
          p += D.b4[0];
          if (D.kind == LIDF_DEC_IEF) s_o[r] += p;      // pred_offset += l4 (implicit_net.py:146)
          else s_o[r] = p;
          if (a.o_iter && di == 0 && D.kind == LIDF_DEC_IEF && it + 1 < D.n_pass && s_orig[r] >= 0)
            a.o_iter[(size_t)it * a.rows + s_orig[r]] = s_o[r];
        }
      }
      __syncthreads();
    }
    if (tid < LIDF_SIMT_BM) {
      float v = s_o[tid];
      v = lidf_final_act(v, D.use_sigmoid);
      s_res[di][tid] = v;
      if (s_orig[tid] >= 0 && D.out) D.out[s_orig[tid]] = v;
    }
    __syncthreads();
  }
  if (tid < LIDF_SIMT_BM && s_orig[tid] >= 0) {
    // pipeline.py:437-439 (scale = sqrt(3), scale2 = part_size) / pipeline.py:1028-1029 (scale = scale2 = 1 skipped)
    float sc = s_res[0][tid] * (a.r1 - a.r0) + a.r0;
    if (!a.refine) { sc = sc * a.scale; sc = sc * a.scale2; }
    for (int k = 0; k < 3; ++k) a.pos_out[(size_t)s_orig[tid] * 3 + k] = s_base[tid][k] + sc * s_dir[tid][k];
  }
}

__host__ inline size_t simt_mlp_smem_bytes(int KP) {
  return sizeof(float) * (size_t)(LIDF_SIMT_BM * KP + 2 * LIDF_SIMT_BM * LIDF_H1 + LIDF_SIMT_BM * LIDF_H2 + LIDF_KC * 256);
}
