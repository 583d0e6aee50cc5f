// lidf_pointnet.cuh -- PointNet2Stage forward (reference src/models/pointnet.py:7-38), the producer of occ_voxel_feat,
// fused into four fp32 kernels (include/lidf_pointnet.h).  Shapes of the shipped YAMLs: 6 -> 32 -> 64 per point,
// per-voxel max, 64 -> 64 per voxel, [voxel 64 | point 64] -> 128 -> 128 per point, per-voxel max, 128 -> 128 per voxel.
//
// The reference runs six cuBLAS GEMMs with their [N,C] activations in HBM plus two torch_scatter max reductions (one
// atomic per point and channel).  Here the per-point activations never leave the SM: a tile of points goes through its
// layers in shared memory / registers, and the per-voxel max is first reduced over runs of consecutive points with the
// same voxel inside the tile (points arrive spatially sorted: image order / ray order), so that only one atomic per
// (run, channel) reaches L2.  All values entering a max are post-ReLU (>= 0), so float max == signed-int max on the bit
// patterns and a zero-initialised buffer reproduces torch_scatter's "rows without a source stay 0".
#pragma once
#include "lidf_pointnet.h"
#include "lidf_common.cuh"

#define PN_IN 6
#define PN_GF 32
#define PN_C1 64
#define PN_C2 128

struct PnWeights {
  const float *w_p1, *b_p1, *w_p2, *b_p2, *w_v1, *b_v1, *w_p3, *b_p3, *w_p4, *b_p4, *w_v2, *b_v2;
};

// f2 = relu(W2 relu(W1 x + b1) + b2) for one point; weights in shared memory ([out][in] as PyTorch stores them, every
// lane reads the same address -> broadcast)
template <int NJ>
__device__ __forceinline__ void pn_point_mlp(const float* __restrict__ x, const float* __restrict__ sW1, const float* __restrict__ sb1,
                                             const float* __restrict__ sW2, const float* __restrict__ sb2, int j0, float (&f2)[NJ]) {
  float f1[PN_GF];
#pragma unroll
  for (int j = 0; j < PN_GF; ++j) {
    float a = sb1[j];
#pragma unroll
    for (int k = 0; k < PN_IN; ++k) a = fmaf(sW1[j * PN_IN + k], x[k], a);
    f1[j] = fmaxf(a, 0.f);
  }
#pragma unroll
  for (int j = 0; j < NJ; ++j) {
    float a = sb2[j0 + j];
    const float4* w = reinterpret_cast<const float4*>(sW2 + (j0 + j) * PN_GF);
#pragma unroll
    for (int k = 0; k < PN_GF / 4; ++k) {
      const float4 ww = w[k];
      a = fmaf(ww.x, f1[4 * k], a); a = fmaf(ww.y, f1[4 * k + 1], a); a = fmaf(ww.z, f1[4 * k + 2], a); a = fmaf(ww.w, f1[4 * k + 3], a);
    }
    f2[j] = fmaxf(a, 0.f);
  }
}

// run-reduced scatter-max of a tile: O[c][p] (channel-major, pitch `pitch`) for points p in [0, np) with voxel s_idx[p]
// (-1 = padding); thread (c, part) scans its slice of points and flushes one atomicMax per run of equal voxel ids
__device__ __forceinline__ void pn_tile_scatter_max(const float* __restrict__ O, int pitch, const int* __restrict__ s_idx, int np, int C,
                                                    int nparts, float* __restrict__ vmax) {
  for (int t = threadIdx.x; t < C * nparts; t += blockDim.x) {
    const int c = t % C, part = t / C;
    const int per = (np + nparts - 1) / nparts, p0 = part * per, p1 = min(np, p0 + per);
    int cur = -1;
    float m = 0.f;
    for (int p = p0; p < p1; ++p) {
      const int v = s_idx[p];
      if (v != cur) {
        if (cur >= 0) atomicMax(reinterpret_cast<int*>(vmax + (size_t)cur * C + c), __float_as_int(m));
        cur = v; m = 0.f;
      }
      m = fmaxf(m, O[c * pitch + p]);
    }
    if (cur >= 0) atomicMax(reinterpret_cast<int*>(vmax + (size_t)cur * C + c), __float_as_int(m));
  }
}

// ---- stage 1: per point 6 -> 32 -> 64, per-voxel max into vmax1 [V][64] (zeroed by the caller) ------------------------
#define PN1_THREADS 256
#define PN1_PITCH (PN1_THREADS + 1)
__global__ void __launch_bounds__(PN1_THREADS) k_pn_stage1(const float* __restrict__ inp, const int64_t* __restrict__ idx, int64_t N,
                                                           int64_t V, PnWeights w, float* __restrict__ vmax1, int* __restrict__ err) {
  extern __shared__ float pn_smem[];
  float* sW1 = pn_smem;                         // [32][6]
  float* sb1 = sW1 + PN_GF * PN_IN;             // [32]
  float* sW2 = sb1 + PN_GF;                     // [64][32]
  float* sb2 = sW2 + PN_C1 * PN_GF;             // [64]
  float* O = sb2 + PN_C1;                       // [64][257]
  int* s_idx = reinterpret_cast<int*>(O + PN_C1 * PN1_PITCH);
  for (int i = threadIdx.x; i < PN_GF * PN_IN; i += blockDim.x) sW1[i] = w.w_p1[i];
  for (int i = threadIdx.x; i < PN_GF; i += blockDim.x) sb1[i] = w.b_p1[i];
  for (int i = threadIdx.x; i < PN_C1 * PN_GF; i += blockDim.x) sW2[i] = w.w_p2[i];
  for (int i = threadIdx.x; i < PN_C1; i += blockDim.x) sb2[i] = w.b_p2[i];
  __syncthreads();
  for (int64_t base = (int64_t)blockIdx.x * PN1_THREADS; base < N; base += (int64_t)gridDim.x * PN1_THREADS) {
    const int64_t i = base + threadIdx.x;
    int v = -1;
    if (i < N) {
      const int64_t vv = idx[i];
      if (vv < 0 || vv >= V) atomicOr(err, 1); else v = (int)vv;
    }
    s_idx[threadIdx.x] = v;
    if (v >= 0) {
      float x[PN_IN], f2[PN_C1];
#pragma unroll
      for (int k = 0; k < PN_IN; ++k) x[k] = inp[i * PN_IN + k];
      pn_point_mlp<PN_C1>(x, sW1, sb1, sW2, sb2, 0, f2);
#pragma unroll
      for (int j = 0; j < PN_C1; ++j) O[j * PN1_PITCH + threadIdx.x] = f2[j];
    }
    __syncthreads();
    const int np = (int)min((int64_t)PN1_THREADS, N - base);
    pn_tile_scatter_max(O, PN1_PITCH, s_idx, np, PN_C1, 4, vmax1);
    __syncthreads();
  }
}

// ---- per-voxel linear + ReLU: out[v][j] = relu(b[j] + sum_k W[j][k] in[v][k]), C x C, one block per 8 voxels ----------
template <int C>
__global__ void __launch_bounds__(C) k_pn_vox(const float* __restrict__ in, const float* __restrict__ W, const float* __restrict__ b,
                                              int64_t V, float* __restrict__ out) {
  __shared__ float s_in[8][C];
  const int64_t v0 = (int64_t)blockIdx.x * 8;
  const int nv = (int)min((int64_t)8, V - v0);
  for (int i = threadIdx.x; i < nv * C; i += C) s_in[i / C][i % C] = in[v0 * C + i];
  __syncthreads();
  const int j = threadIdx.x;
  float acc[8];
#pragma unroll
  for (int r = 0; r < 8; ++r) acc[r] = b[j];
  const float4* wr = reinterpret_cast<const float4*>(W + (size_t)j * C);
  for (int k = 0; k < C / 4; ++k) {
    const float4 ww = __ldg(wr + k);
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      acc[r] = fmaf(ww.x, s_in[r][4 * k], acc[r]); acc[r] = fmaf(ww.y, s_in[r][4 * k + 1], acc[r]);
      acc[r] = fmaf(ww.z, s_in[r][4 * k + 2], acc[r]); acc[r] = fmaf(ww.w, s_in[r][4 * k + 3], acc[r]);
    }
  }
  for (int r = 0; r < nv; ++r) out[(v0 + r) * C + j] = fmaxf(acc[r], 0.f);
}

// ---- stage 2: per point [vf1[voxel] (64) | f2 (64)] -> 128 -> 128, per-voxel max into vmax2 [V][128] ------------------
// Tile = 64 points.  Activations channel-major in shared memory (X, H: [128][64 + 4]); weights k-major ([k][j], transposed
// once per block); thread (pg, jg) of 16 x 16 owns 4 points x 8 outputs: 3 LDS.128 per 32 FFMA.
#define PN2_THREADS 256
#define PN2_TP 64
#define PN2_PITCH (PN2_TP + 4)
__device__ __forceinline__ void pn_tile_gemm(const float* __restrict__ Xs, const float* __restrict__ Wt, const float* __restrict__ bias,
                                             int pg, int jg, float (&acc)[4][8]) {
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[a][c] = bias[8 * jg + c];
#pragma unroll 4
  for (int k = 0; k < PN_C2; ++k) {
    const float4 xa = *reinterpret_cast<const float4*>(Xs + k * PN2_PITCH + 4 * pg);
    const float4 w0 = *reinterpret_cast<const float4*>(Wt + k * PN_C2 + 8 * jg);
    const float4 w1 = *reinterpret_cast<const float4*>(Wt + k * PN_C2 + 8 * jg + 4);
    const float xs[4] = {xa.x, xa.y, xa.z, xa.w};
    const float ws[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int c = 0; c < 8; ++c) acc[a][c] = fmaf(xs[a], ws[c], acc[a][c]);
  }
}

__global__ void __launch_bounds__(PN2_THREADS, 1) k_pn_stage2(const float* __restrict__ inp, const int64_t* __restrict__ idx, int64_t N,
                                                              int64_t V, PnWeights w, const float* __restrict__ vf1,
                                                              float* __restrict__ vmax2) {
  extern __shared__ float pn_smem[];
  float* sW1 = pn_smem;                          // [32][6]
  float* sb1 = sW1 + PN_GF * PN_IN;              // [32]
  float* sW2 = sb1 + PN_GF;                      // [64][32]
  float* sb2 = sW2 + PN_C1 * PN_GF;              // [64]
  float* sb3 = sb2 + PN_C1;                      // [128]
  float* sb4 = sb3 + PN_C2;                      // [128]
  float* W3t = sb4 + PN_C2;                      // [128 k][128 j]
  float* W4t = W3t + PN_C2 * PN_C2;
  float* X = W4t + PN_C2 * PN_C2;                // [128][68]
  float* H = X + PN_C2 * PN2_PITCH;              // [128][68]
  int* s_idx = reinterpret_cast<int*>(H + PN_C2 * PN2_PITCH);
  for (int i = threadIdx.x; i < PN_GF * PN_IN; i += blockDim.x) sW1[i] = w.w_p1[i];
  for (int i = threadIdx.x; i < PN_GF; i += blockDim.x) sb1[i] = w.b_p1[i];
  for (int i = threadIdx.x; i < PN_C1 * PN_GF; i += blockDim.x) sW2[i] = w.w_p2[i];
  for (int i = threadIdx.x; i < PN_C1; i += blockDim.x) sb2[i] = w.b_p2[i];
  for (int i = threadIdx.x; i < PN_C2; i += blockDim.x) { sb3[i] = w.b_p3[i]; sb4[i] = w.b_p4[i]; }
  for (int i = threadIdx.x; i < PN_C2 * PN_C2; i += blockDim.x) {
    const int j = i / PN_C2, k = i % PN_C2;      // coalesced read of W[j][k], transposed store
    W3t[k * PN_C2 + j] = w.w_p3[i];
    W4t[k * PN_C2 + j] = w.w_p4[i];
  }
  __syncthreads();
  const int pg = threadIdx.x & 15, jg = threadIdx.x >> 4;
  const int n_tiles = (int)((N + PN2_TP - 1) / PN2_TP);
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t base = (int64_t)tile * PN2_TP;
    const int np = (int)min((int64_t)PN2_TP, N - base);
    // inputs of the tile: thread (p, q) builds a quarter of the 64 point features and copies a quarter of the voxel row
    {
      const int p = threadIdx.x & 63, q = threadIdx.x >> 6;
      const int64_t i = base + p;
      int v = -1;
      if (p < np) { const int64_t vv = idx[i]; if (vv >= 0 && vv < V) v = (int)vv; }
      if (q == 0) s_idx[p] = v;
      if (v >= 0) {
        float x[PN_IN], f2[16];
#pragma unroll
        for (int k = 0; k < PN_IN; ++k) x[k] = inp[i * PN_IN + k];
        pn_point_mlp<16>(x, sW1, sb1, sW2, sb2, 16 * q, f2);
        const float4* vr = reinterpret_cast<const float4*>(vf1 + (size_t)v * PN_C1 + 16 * q);
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4) {
          const float4 g = __ldg(vr + c4);
          X[(16 * q + 4 * c4 + 0) * PN2_PITCH + p] = g.x; X[(16 * q + 4 * c4 + 1) * PN2_PITCH + p] = g.y;
          X[(16 * q + 4 * c4 + 2) * PN2_PITCH + p] = g.z; X[(16 * q + 4 * c4 + 3) * PN2_PITCH + p] = g.w;
        }
#pragma unroll
        for (int c = 0; c < 16; ++c) X[(PN_C1 + 16 * q + c) * PN2_PITCH + p] = f2[c];     // cat((voxel, point), -1), pointnet.py:31
      } else {
#pragma unroll
        for (int c = 0; c < 16; ++c) { X[(16 * q + c) * PN2_PITCH + p] = 0.f; X[(PN_C1 + 16 * q + c) * PN2_PITCH + p] = 0.f; }
      }
    }
    __syncthreads();
    float acc[4][8];
    pn_tile_gemm(X, W3t, sb3, pg, jg, acc);                                   // point_lin3 + ReLU
#pragma unroll
    for (int c = 0; c < 8; ++c)
      *reinterpret_cast<float4*>(H + (8 * jg + c) * PN2_PITCH + 4 * pg) =
          make_float4(fmaxf(acc[0][c], 0.f), fmaxf(acc[1][c], 0.f), fmaxf(acc[2][c], 0.f), fmaxf(acc[3][c], 0.f));
    __syncthreads();
    pn_tile_gemm(H, W4t, sb4, pg, jg, acc);                                   // point_lin4 + ReLU
#pragma unroll
    for (int c = 0; c < 8; ++c)
      *reinterpret_cast<float4*>(X + (8 * jg + c) * PN2_PITCH + 4 * pg) =
          make_float4(fmaxf(acc[0][c], 0.f), fmaxf(acc[1][c], 0.f), fmaxf(acc[2][c], 0.f), fmaxf(acc[3][c], 0.f));
    __syncthreads();
    pn_tile_scatter_max(X, PN2_PITCH, s_idx, np, PN_C2, 2, vmax2);
    __syncthreads();
  }
}

// ---- stage 2 on the tensor cores (tcgen05) ------------------------------------------------------------------------------
// The two 128 -> 128 per-point layers are 97 % of PointNet2Stage's MACs (32,768 of 33,984 per point).  Same split-bf16
// arithmetic and primitives as the decoder engine (lidf_tc.cuh): x = hi + lo, three products per MAC, fp32 accumulators
// in TMEM.  Persistent CTA per SM, tile = 128 points (row = TMEM lane = point):
//   16 row warps (TMEM quadrant q, column group g): build the layer-3 operand [voxel feature (64) | point feature (64)]
//     in shared memory (UMMA canonical K-major layout; the 6 -> 32 -> 64 point MLP stays on the FMA pipe, 16 outputs per
//     thread), run the two epilogues (bias + ReLU; layer 3's output is re-split and written back to TMEM as the layer-4
//     operand) and the per-voxel run-reduced max of the tile
//   warp 16: one thread loads BOTH packed weight matrices once (2 x 64 KB stay resident in shared memory for the CTA's
//     lifetime -- no weight streaming) and issues the 2 x 24 MMAs of a tile (layer 3 SS-mode, layer 4 TS-mode).
// TMEM: [0,128) layer-3 accumulator, [128,256) layer-4 operand (hi | lo), [256,384) layer-4 accumulator.
// The operand tile doubles as the fp32 [channel][point] staging buffer of the scatter-max once layer 3 has consumed it.
#define PNT_ROW_WARPS 16
#define PNT_THREADS ((PNT_ROW_WARPS + 1) * 32)
#define PNT_KSTEPS 8

struct PnTcSmem {
  uint8_t w[2][PNT_KSTEPS][TC_CHUNK_BYTES];      // point_lin3, point_lin4: per k-step [hi: kg0 | kg1][lo: kg0 | kg1], N = 128
  uint8_t x[2][PNT_KSTEPS * 4096];                // layer-3 operand [hi | lo][k-step][kgroup][128 rows][16 B]; later O[128][128] fp32
  float W1[PN_GF * PN_IN], b1[PN_GF], W2[PN_C1 * PN_GF], b2[PN_C1], b3[PN_C2], b4[PN_C2];
  int idx[128];
  uint64_t w_full, a_ready, d3_full, a4_ready, d4_full;
  uint32_t tmem_base;
};

// W [128 out][128 in] fp32 row-major -> 8 k-step chunks of 8 KB in the UMMA canonical layout, bf16 hi | lo
__global__ void k_pn_pack_tc(const float* __restrict__ W, uint8_t* __restrict__ chunks) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= PNT_KSTEPS * 2048) return;
  const int c = idx / 2048, r = idx % 2048, n = r / 16, kk = r % 16;
  const float w = W[n * PN_C2 + 16 * c + kk];
  const __nv_bfloat16 hi = __float2bfloat16_rn(w);
  const __nv_bfloat16 lo = __float2bfloat16_rn(w - __bfloat162float(hi));
  const size_t off = (size_t)(kk >> 3) * 128 * 16 + (size_t)n * 16 + (kk & 7) * 2;
  *reinterpret_cast<__nv_bfloat16*>(chunks + (size_t)c * TC_CHUNK_BYTES + off) = hi;
  *reinterpret_cast<__nv_bfloat16*>(chunks + (size_t)c * TC_CHUNK_BYTES + 128 * 32 + off) = lo;
}

__global__ void __launch_bounds__(PNT_THREADS, 1) k_pn_stage2_tc(const float* __restrict__ inp, const int64_t* __restrict__ idx,
                                                                 int64_t N, int64_t V, PnWeights w, const uint8_t* __restrict__ wpack,
                                                                 const float* __restrict__ vf1, float* __restrict__ vmax2) {
  extern __shared__ __align__(1024) uint8_t pn_smem_raw[];
  PnTcSmem& S = *reinterpret_cast<PnTcSmem*>(pn_smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < PN_GF * PN_IN; i += PNT_THREADS) S.W1[i] = w.w_p1[i];
  for (int i = tid; i < PN_GF; i += PNT_THREADS) S.b1[i] = w.b_p1[i];
  for (int i = tid; i < PN_C1 * PN_GF; i += PNT_THREADS) S.W2[i] = w.w_p2[i];
  for (int i = tid; i < PN_C1; i += PNT_THREADS) S.b2[i] = w.b_p2[i];
  for (int i = tid; i < PN_C2; i += PNT_THREADS) { S.b3[i] = w.b_p3[i]; S.b4[i] = w.b_p4[i]; }
  if (tid == 0) {
    tc::mbar_init(&S.w_full, 1);
    tc::mbar_init(&S.a_ready, PNT_ROW_WARPS);
    tc::mbar_init(&S.d3_full, 1);
    tc::mbar_init(&S.a4_ready, PNT_ROW_WARPS);
    tc::mbar_init(&S.d4_full, 1);
    tc::fence_barrier_init();
  }
  if (warp == PNT_ROW_WARPS) tc::tmem_alloc(&S.tmem_base, 512);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = S.tmem_base;
  const int n_tiles = (int)((N + 127) / 128);
  const int n_my_tiles = n_tiles > (int)blockIdx.x ? (n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  constexpr uint32_t COL_D3 = 0, COL_A4 = 128, COL_D4 = 256;

  if (warp == PNT_ROW_WARPS) {
    // ================================ weight load (once) + MMA issue, one thread ================================
    if (tc::elect_one()) {
      tc::mbar_arrive_expect_tx(&S.w_full, 2 * PNT_KSTEPS * TC_CHUNK_BYTES);
#pragma unroll 1
      for (int c = 0; c < 2 * PNT_KSTEPS; ++c)
        tc::bulk_g2s(&S.w[0][0][0] + (size_t)c * TC_CHUNK_BYTES, wpack + (size_t)c * TC_CHUNK_BYTES, TC_CHUNK_BYTES, &S.w_full);
      tc::mbar_wait(&S.w_full, 0);
      constexpr uint32_t idesc = tc::make_idesc(128);
      const uint64_t w3d = tc::make_bdesc(tc::smem_u32(&S.w[0][0][0]), 2048u, 128u);
      const uint64_t w4d = tc::make_bdesc(tc::smem_u32(&S.w[1][0][0]), 2048u, 128u);
      const uint64_t ad = tc::make_bdesc(tc::smem_u32(S.x[0]), 2048u, 128u);
      for (int t = 0; t < n_my_tiles; ++t) {
        const uint32_t ph = (uint32_t)t & 1u;
        tc::mbar_wait(&S.a_ready, ph);
        tc::fence_after_sync();
#pragma unroll 1
        for (int ks = 0; ks < PNT_KSTEPS; ++ks) {                                // layer 3: operand in shared memory
          const uint64_t bhi = w3d + ((uint32_t)(ks * TC_CHUNK_BYTES) >> 4), ahi = ad + ((uint32_t)(ks * 4096) >> 4);
          tc::mma_ss(tmem + COL_D3, ahi, bhi, idesc, ks == 0 ? 0u : 1u);
          tc::mma_ss(tmem + COL_D3, ahi + ((uint32_t)(PNT_KSTEPS * 4096) >> 4), bhi, idesc, 1u);
          tc::mma_ss(tmem + COL_D3, ahi, bhi + (4096u >> 4), idesc, 1u);
        }
        tc::commit(&S.d3_full);
        tc::mbar_wait(&S.a4_ready, ph);
        tc::fence_after_sync();
#pragma unroll 1
        for (int ks = 0; ks < PNT_KSTEPS; ++ks) {                                // layer 4: operand in TMEM (8 cols hi | 8 cols lo)
          const uint64_t bhi = w4d + ((uint32_t)(ks * TC_CHUNK_BYTES) >> 4);
          const uint32_t acol = tmem + COL_A4 + 16 * ks;
          tc::mma_ts(tmem + COL_D4, acol, bhi, idesc, ks == 0 ? 0u : 1u);
          tc::mma_ts(tmem + COL_D4, acol + 8, bhi, idesc, 1u);
          tc::mma_ts(tmem + COL_D4, acol, bhi + (4096u >> 4), idesc, 1u);
        }
        tc::commit(&S.d4_full);
      }
    }
  } else {
    // ================================ row warps ================================
    const int q = warp & 3, g = warp >> 2;
    const int row = q * 32 + lane;
    const uint32_t lane_addr = tmem + ((uint32_t)(q * 32) << 16);
    const uint32_t x_hi = tc::smem_u32(S.x[0]) + row * 16, x_lo = tc::smem_u32(S.x[1]) + row * 16;
    float* const O = reinterpret_cast<float*>(&S.x[0][0]);
    auto st_x = [&](int ks, const uint32_t* wd) {
      tc::st_shared_v4(x_hi + ks * 4096, wd[0], wd[1], wd[2], wd[3]);
      tc::st_shared_v4(x_hi + ks * 4096 + 2048, wd[4], wd[5], wd[6], wd[7]);
      tc::st_shared_v4(x_lo + ks * 4096, wd[8], wd[9], wd[10], wd[11]);
      tc::st_shared_v4(x_lo + ks * 4096 + 2048, wd[12], wd[13], wd[14], wd[15]);
    };
    for (int t = 0; t < n_my_tiles; ++t) {
      const int tile = (int)blockIdx.x + t * (int)gridDim.x;
      const int64_t base = (int64_t)tile * 128;
      const int np = (int)min((int64_t)128, N - base);
      const uint32_t ph = (uint32_t)t & 1u;
      // ---- layer-3 operand: cat((voxel feature, point feature), -1), pointnet.py:31; this thread owns 16 + 16 of the 128
      {
        const int64_t i = base + row;
        int v = -1;
        if (row < np) { const int64_t vv = idx[i]; if (vv >= 0 && vv < V) v = (int)vv; }
        if (g == 0) S.idx[row] = v;
        uint32_t wd[16];
        {                                                                    // point feature: 6 -> 32 -> 64, outputs 16 g .. 16 g + 15
          float f2[16];
          if (v >= 0) {
            float x[PN_IN];
#pragma unroll
            for (int k = 0; k < PN_IN; ++k) x[k] = inp[i * PN_IN + k];
            pn_point_mlp<16>(x, S.W1, S.b1, S.W2, S.b2, 16 * g, f2);
          } else {
#pragma unroll
            for (int c = 0; c < 16; ++c) f2[c] = 0.f;
          }
          tc::split16(f2, wd);
          st_x(4 + g, wd);
        }
        {                                                                    // voxel feature of the point's voxel, columns 16 g .. 16 g + 15
          float xv[16];
#pragma unroll
          for (int c = 0; c < 16; ++c) xv[c] = 0.f;
          if (v >= 0) {
            const float4* vr = reinterpret_cast<const float4*>(vf1 + (size_t)v * PN_C1 + 16 * g);
#pragma unroll
            for (int c4 = 0; c4 < 4; ++c4) {
              const float4 gq = __ldg(vr + c4);
              xv[4 * c4] = gq.x; xv[4 * c4 + 1] = gq.y; xv[4 * c4 + 2] = gq.z; xv[4 * c4 + 3] = gq.w;
            }
          }
          tc::split16(xv, wd);
          st_x(g, wd);
        }
      }
      tc::fence_proxy_async();
      tc::fence_before_sync();                                               // also orders the previous tile's TMEM loads
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&S.a_ready);
      // ---- layer-3 epilogue: relu(acc + b3) -> hi | lo -> layer-4 operand in TMEM
      tc::mbar_wait(&S.d3_full, ph);
      tc::fence_after_sync();
      {
        uint32_t r[32];
        tc::tmem_ld32(lane_addr + COL_D3 + 32 * g, r);
        tc::wait_ld();
#pragma unroll
        for (int s16 = 0; s16 < 2; ++s16) {
          float h[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) h[j] = fmaxf(__uint_as_float(r[16 * s16 + j]) + S.b3[32 * g + 16 * s16 + j], 0.f);
          uint32_t wd[16];
          tc::split16(h, wd);
          tc::tmem_st16(lane_addr + COL_A4 + 32 * g + 16 * s16, wd);
        }
        tc::wait_st();
      }
      tc::fence_before_sync();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&S.a4_ready);
      // ---- layer-4 epilogue: relu(acc + b4) -> O[channel][point] (XOR-swizzled columns: conflict-free both ways)
      tc::mbar_wait(&S.d4_full, ph);
      tc::fence_after_sync();
      // O overlays the layer-3 operand tile.  Every warp's operand stores are ordered before this point through the mbarrier
      // chain (a_ready -> MMAs -> d3_full -> a4_ready -> MMAs -> d4_full); the explicit barrier states that for tools that do
      // not follow mbarriers (compute-sanitizer racecheck) at the price of one bar.sync per tile.
      asm volatile("bar.sync 1, %0;" ::"n"(PNT_ROW_WARPS * 32) : "memory");
      {
        uint32_t r[32];
        tc::tmem_ld32(lane_addr + COL_D4 + 32 * g, r);
        tc::wait_ld();
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int c = 32 * g + j;
          O[c * 128 + (row ^ j)] = fmaxf(__uint_as_float(r[j]) + S.b4[c], 0.f);
        }
      }
      asm volatile("bar.sync 1, %0;" ::"n"(PNT_ROW_WARPS * 32) : "memory");
      // ---- per-voxel max of the tile: thread (channel c, quarter part) scans 32 points, one atomic per run of equal voxel ids
      {
        const int c = tid & 127, part = tid >> 7;
        const int p0 = 32 * part, p1 = min(np, p0 + 32);
        int cur = -1;
        float m = 0.f;
        for (int p = p0; p < p1; ++p) {
          const int v = S.idx[p];
          if (v != cur) {
            if (cur >= 0) atomicMax(reinterpret_cast<int*>(vmax2 + (size_t)cur * PN_C2 + c), __float_as_int(m));
            cur = v; m = 0.f;
          }
          m = fmaxf(m, O[c * 128 + (p ^ (c & 31))]);
        }
        if (cur >= 0) atomicMax(reinterpret_cast<int*>(vmax2 + (size_t)cur * PN_C2 + c), __float_as_int(m));
      }
      asm volatile("bar.sync 1, %0;" ::"n"(PNT_ROW_WARPS * 32) : "memory");    // O / idx are overwritten by the next tile's build
    }
  }
  __syncwarp();
  tc::fence_before_sync();
  __syncthreads();
  if (warp == PNT_ROW_WARPS) tc::tmem_dealloc(tmem, 512);
}

inline size_t pn_stage1_smem() { return sizeof(float) * (PN_GF * PN_IN + PN_GF + PN_C1 * PN_GF + PN_C1 + PN_C1 * PN1_PITCH) + sizeof(int) * PN1_THREADS; }
inline size_t pn_stage2_smem() {
  return sizeof(float) * (PN_GF * PN_IN + PN_GF + PN_C1 * PN_GF + PN_C1 + 2 * PN_C2 + 2 * PN_C2 * PN_C2 + 2 * PN_C2 * PN2_PITCH) +
         sizeof(int) * PN2_TP;
}
