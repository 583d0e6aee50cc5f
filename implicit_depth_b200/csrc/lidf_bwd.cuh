// lidf_bwd.cuh -- native backward of the LIDF query path (LIDF.get_embedding + LIDF.get_pred under autograd,
// reference src/models/pipeline.py:338-466, decoders src/models/implicit_net.py:81-98,129-152), sm_100a only.
//
// What reaches the path from the losses (pipeline.py:472,482): dL/d pred_pos [R,3] (through the arg-max gather: only the
// winning pair of a ray gets a gradient; the soft-max that selects it is detached, :442), dL/d pred_prob_end [P] and,
// for completeness, dL/d pred_offset / dL/d pair_pred_pos.  The backward mirrors the forward's factored layer 1:
//
//   z1 = W1[:,pos] PE(enter,leave) + A_v[vox] + T[ray] (+ u o_in)       A_v = W1[:,vox] occ_voxel_feat, T = per-ray term
//
//   k_bwd_seed     per pair : dL/d o_last of both decoders (offset scaling, arg-max gather, final activation)
//   k_mlp_bwd_tc   "B1", tcgen05, one decoder pass per launch over a chunk of ray-major pairs: recomputes the forward
//                  (activations are never stored by the forward), then the two dgrad GEMMs as TS-mode MMAs with
//                  transposed weight chunks; h1, h2, delta1..3 (and PE once) leave the SM for the wgrad kernel already
//                  split into bf16 hi | lo and in that kernel's operand layout (bw_pk_*; delta1 also as fp32 rows for the
//                  segment sums); bias / w4 / IEF-u gradients are column sums done with shuffles; the IEF feedback
//                  dL/d o_{k-1} = dL/d o_k + u . delta1 is carried between the passes in a [P] array.
//   k_wgrad_pk_tc  "B2", tcgen05: C[128,N] += A^T B over the rows of a chunk (K = pair index), operands fetched packed by
//                  TMA bulk copies (MN-major descriptors), 3 products, accumulator resident in TMEM over all row groups
//                  of the CTA, per-CTA partial slices reduced in a fixed order afterwards (reproducible).
//   k_wgrad_tc     the same product for operands that only exist as fp32 rows (the row-level terms: G_r, G_v, ROI feature,
//                  occ_voxel_feat): loader warps convert to bf16 hi/lo into the K-major layout.
//   k_segsum_*     G_r[ray] += delta1, G_v[vox] += delta1: the factored terms turn the two widest weight gradients into
//                  segment sums followed by small GEMMs (dW1[:,rgb|dir] = G_r^T [roi | PE(dir)], dW1[:,vox] = G_v^T feat,
//                  d roi = G_r W1[:,rgb], d occ_voxel_feat = G_v W1[:,vox])
//   k_roi_align_backward   d roi -> d full_rgb_feat (transpose of torchvision's aligned ROIAlign sampling)
#pragma once
#include "lidf_tc.cuh"

#define BW_CHUNKS_BWD 20               // 4 (W3^T, N = 128, K = 64) + 8 + 8 (W2^T halves, N = 128, K = 128)
#define BW_CHUNKS_FWD TC_CHUNKS_PER_DEC
#define BW_PE_F 128                    // feature count of the handed-over layer-1 MMA operand (K order of tc_a1_col; 112 + 16 zeros)
#define BW_COLPART 8                   // floats per (cta, warp, lane) slot of the column-sum partials

// mbarrier waits of the backward kernels: the tile-serial kernels hand over between the row warps and the MMA issuer a dozen
// times per tile, so the wake-up latency matters more than the power a parked warp saves -> short suspend hint.
#ifndef BW_WAIT_HINT_NS
#define BW_WAIT_HINT_NS 20
#endif
__device__ __forceinline__ void bw_wait_a(uint32_t bar_saddr, uint32_t parity) {
  uint32_t spins = 0;
  for (;;) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar_saddr), "r"(parity), "r"((uint32_t)BW_WAIT_HINT_NS)
        : "memory");
    if (ok) break;
    if (++spins > (1u << 26)) __trap();
  }
}
__device__ __forceinline__ void bw_wait(uint64_t* bar, uint32_t parity) { bw_wait_a(tc::smem_u32(bar), parity); }

// ------------------------------------------------------------------------------------------------ seed
// dL/d o_last per pair and decoder, at the ORIGINAL pair index.
//   pair_pred_pos = enter + ((off (r1-r0) + r0) sqrt3 part) dir      (pipeline.py:437-439)
//   pred_pos[ray] = pair_pred_pos[max_pair_id[ray]]                  (:452-454)
// final activation (implicit_net.py:93-96,148-151) differentiated from its output y: sigmoid y(1-y); leaky clamp
// max(min(x, .01x+.99), .01x): 1 for 0 < y < 1, else 0.01.
__device__ __forceinline__ float lidf_final_act_grad(float y, int use_sigmoid) {
  if (use_sigmoid) return y * (1.0f - y);
  return (y > 0.f && y < 1.f) ? 1.0f : 0.01f;
}

struct BwSeedArgs {
  int64_t P, R;
  const int64_t* pair_ray; const int64_t* max_pair_id; const float* ray_dir;
  const float* pred_offset; const float* pred_prob_end;
  const float* g_pred_pos; const float* g_pred_prob_end; const float* g_pred_offset; const float* g_pair_pred_pos;
  float scale;                         // (r1 - r0) sqrt(3) part_size
  int sig0, sig1;
  float* g0; float* g1;                // [P]
};
__global__ void k_bwd_seed(const BwSeedArgs a) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.P) return;
  int64_t ray = a.pair_ray[i];
  ray = ray < 0 ? 0 : (ray >= a.R ? a.R - 1 : ray);
  float gp[3] = {0.f, 0.f, 0.f};
  if (a.g_pair_pred_pos) { gp[0] = a.g_pair_pred_pos[3 * i]; gp[1] = a.g_pair_pred_pos[3 * i + 1]; gp[2] = a.g_pair_pred_pos[3 * i + 2]; }
  if (a.g_pred_pos && a.max_pair_id[ray] == i) {
    gp[0] += a.g_pred_pos[3 * ray]; gp[1] += a.g_pred_pos[3 * ray + 1]; gp[2] += a.g_pred_pos[3 * ray + 2];
  }
  const float dot = gp[0] * a.ray_dir[3 * ray] + gp[1] * a.ray_dir[3 * ray + 1] + gp[2] * a.ray_dir[3 * ray + 2];
  float go = dot * a.scale;
  if (a.g_pred_offset) go += a.g_pred_offset[i];
  a.g0[i] = go * lidf_final_act_grad(a.pred_offset[i], a.sig0);
  a.g1[i] = a.g_pred_prob_end ? a.g_pred_prob_end[i] * lidf_final_act_grad(a.pred_prob_end[i], a.sig1) : 0.f;
}

// ------------------------------------------------------------------------------------------------ weight packing
// transposed chunks for the two dgrad GEMMs, same [hi: kg0 N x 16 B | kg1][lo] layout as the forward chunks (N = 128):
//   chunk c < 4          : delta2pre = delta3 W3     -> B[n][k] = w3[k][n],        k = 16 c + kk        (K = 64)
//   chunk 4 + 8 h + s    : delta1pre[:,128h..] = delta2 W2[:,128h..] -> B[n][k] = w2[k][128 h + n], k = 16 s + kk
__global__ void k_pack_tc_weights_bwd(const float* __restrict__ w2, const float* __restrict__ w3, uint8_t* __restrict__ stream) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= BW_CHUNKS_BWD * 2048) return;
  const int c = idx / 2048, r = idx % 2048, n = r / 16, kk = r % 16;
  float w;
  if (c < 4) w = w3[(size_t)(16 * c + kk) * LIDF_H2 + n];
  else { const int h = (c - 4) / 8, s = (c - 4) % 8; w = w2[(size_t)(16 * s + kk) * LIDF_H1 + 128 * h + n]; }
  const __nv_bfloat16 hi = __float2bfloat16_rn(w);
  const __nv_bfloat16 lo = __float2bfloat16_rn(w - __bfloat162float(hi));
  const size_t off = (size_t)(kk >> 3) * 128 * 16 + (size_t)n * 16 + (kk & 7) * 2;
  uint8_t* base = stream + (size_t)c * TC_CHUNK_BYTES;
  *reinterpret_cast<__nv_bfloat16*>(base + off) = hi;
  *reinterpret_cast<__nv_bfloat16*>(base + 128 * 32 + off) = lo;
}

// ------------------------------------------------------------------------------------------------ B1
struct BwArgs {
  int64_t P; int64_t s0; int n_rows; int n_tiles;   // chunk = sorted positions [s0, s0 + n_rows)
  const int* perm; const int64_t* pair_vox; const int64_t* pair_ray;
  const float* pair_dist; const float* dense_dist; int64_t R; int64_t V;
  const float* ray_dir; const float* voxel_bound; int rel;
  const float* Av; const float* T; int dcol;        // this decoder's 256 columns of the [.,512] tables
  const uint8_t* wfwd; const uint8_t* wbwd;
  const float* u; const float* b2; const float* b3; const float* w4;
  int is_ief; int it; float o0;
  const float* o_in;                                // [P] offset fed to this iteration (original index); NULL: o0
  float* g;                                         // [P] in: dL/d o_it, out (IEF, it > 0): dL/d o_{it-1}
  float* h1; float* h2; float* d2; float* d3;       // chunk-local, PACKED split-bf16 hand-over to k_wgrad_pk_tc (layout: bw_pk_*)
  float* pe;                                        // packed too (F = BW_PE_F): the layer-1 MMA operand; NULL = do not write
  float* d1;                                        // chunk-local fp32 rows [.,256] for the ray / voxel segment sums (summed over IEF passes)
  float* d1pk;                                      // the same sum, packed, for dW1[:,pos]; non-NULL on the last pass processed (it = 0)
  int d1_accumulate;
  float* colpart;                                   // [grid][16][32][BW_COLPART], accumulated
  int by_slot;                                      // winner-only backward of the offset decoder: rows = rays (perm = each ray's
                                                    // arg-max pair, -1: none), g / o_in indexed by the row's slot s0 + row
};

// ---- packed hand-over B1 -> k_wgrad_pk_tc ("PK" layout) ------------------------------------------------------------
// The wgrad GEMMs contract over the ROW index (K = pair), and B1 owns one row per thread.  A tensor X[rows][F] is handed
// over already split into bf16 hi | lo and already in the shared-memory layout the MN-major UMMA descriptor wants, so that
// the wgrad kernel only has to bulk-copy it (TMA) -- no conversion, no transpose, no register staging on either side:
//   group G = row / 64 (one pipeline stage of the wgrad kernel), 64 F 4 bytes each:
//     [hi: F/8 feature groups][64 rows][16 B = 8 consecutive features of that row, bf16]  then  [lo: same]
// i.e. in UMMA terms ((8 features),(8 rows, 8 row groups)) with LBO = 128 B between 8-row core matrices and SBO = 1 KB
// between feature groups (cute: "Major-MN, INTERLEAVE: ((1,n),(8,k)):((X,SBO),(1,LBO))" in 16-byte units).
// Same bytes per element as fp32 (2 + 2), same arithmetic (the split is the one B1 makes for its own TMEM operands).
#define BW_PK_ROWS 64
#define BW_PK_FG_BYTES (BW_PK_ROWS * 16)          // one feature group: 64 rows x 16 B
__host__ __device__ inline size_t bw_pk_group_bytes(int F) { return (size_t)BW_PK_ROWS * F * 4; }
__host__ __device__ inline size_t bw_pk_lo_offset(int F) { return (size_t)(F / 8) * BW_PK_FG_BYTES; }
// address of the 16-byte hi unit holding features [f0, f0 + 8) (f0 % 8 == 0) of chunk row `rl`
__device__ __forceinline__ uint8_t* bw_pk_ptr(float* tensor, int F, int64_t rl, int f0) {
  return reinterpret_cast<uint8_t*>(tensor) + (size_t)(rl / BW_PK_ROWS) * bw_pk_group_bytes(F) + (size_t)(f0 >> 3) * BW_PK_FG_BYTES +
         (size_t)(rl % BW_PK_ROWS) * 16;
}

#define BW_SLOTS 8                     // weight ring: 8 x 8 KB chunks
#define BW_STAGE_PITCH 36              // floats per row of a staging tile (32 + 4: conflict-free 128-bit accesses)
struct BwSmem {
  uint8_t w[BW_SLOTS][TC_CHUNK_BYTES];
  uint8_t a1[2][TC_A1S_PART_BYTES];
  float u[LIDF_H1]; float b2[LIDF_H2]; float b3[LIDF_H3]; float w4[LIDF_H3];
  float part[2][4][128];
  float stage[TC_ROW_WARPS][32 * BW_STAGE_PITCH];   // per-warp transpose tile: rows <-> lanes, so that global accesses are whole rows
  uint64_t w_full[BW_SLOTS], w_empty[BW_SLOTS];
  uint64_t a1_ready, a1_free, x_full[2], x_done[2], y_full, y_done, z_full, z_done, y2_full, y2_done, x2_full[2], x2_done;
  uint32_t tmem_base;
};

// column sums over the 32 lanes of a warp: v[c] of every lane -> lane l returns sum over lanes of v[l]  (31 shuffles)
__device__ __forceinline__ float bw_colreduce32(float (&v)[32], int lane) {
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) {
    const bool up = (lane & s) != 0;
#pragma unroll
    for (int i = 0; i < s; ++i) {
      const float send = up ? v[i] : v[i + s];
      const float keep = up ? v[i + s] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
    }
  }
  return v[0];
}

template <int NPROD>
__device__ __forceinline__ void bw_ts128(uint32_t tmem, uint64_t wd128, uint32_t dcol, uint32_t acol, uint32_t off, bool first) {
  constexpr uint32_t idesc = tc::make_idesc(128);
  const uint64_t bhi = wd128 + (off >> 4);
  tc::mma_ts(tmem + dcol, tmem + acol, bhi, idesc, first ? 0u : 1u);
  if (NPROD == 3) {
    tc::mma_ts(tmem + dcol, tmem + acol + 8, bhi, idesc, 1u);
    tc::mma_ts(tmem + dcol, tmem + acol, bhi + (4096u >> 4), idesc, 1u);
  }
}
template <int NPROD>
__device__ __forceinline__ void bw_ts64(uint32_t tmem, uint64_t wd64, uint32_t dcol, uint32_t acol, uint32_t off, bool first) {
  constexpr uint32_t idesc = tc::make_idesc(64);
  const uint64_t bhi = wd64 + (off >> 4);
  tc::mma_ts(tmem + dcol, tmem + acol, bhi, idesc, first ? 0u : 1u);
  if (NPROD == 3) {
    tc::mma_ts(tmem + dcol, tmem + acol + 8, bhi, idesc, 1u);
    tc::mma_ts(tmem + dcol, tmem + acol, bhi + (2048u >> 4), idesc, 1u);
  }
}

// One decoder pass, forward recompute + dgrad, tile-serial (the kernel is a latency chain of seven epilogues and five MMA
// phases per tile, not bound by the tensor pipe).  Same roles and TMEM plan as k_mlp_tc: Z [0,64), X0 [128,256),
// X1 [256,384), Y [384,512).
// Row-major fp32 traffic (the A_v / T gathers, delta1) goes through a per-warp shared-memory transpose tile: a thread owns a
// ROW (TMEM lane) but a coalesced access wants 8 lanes per 128-byte row segment -- the first version stored 16 bytes per
// lane into 32 different lines per instruction and sat at 70 % L1TEX throughput / 26 % of the HBM rate (profiles/r2a); two
// CTAs per SM did not help.  The packed hand-over tensors need no transpose (see bw_pk_*).
//   MMA order per tile : L1 -> X0, X1 | L2 -> Y | L3 -> Z | D2: delta3 (in Z) W3 -> Y | D1: delta2 (in Y) W2 -> X0, X1
//   row warps per tile : E1 x2 (h1, mask1) | [next tile's operand] | E2 (h2, mask2) | E3 (delta3) | Ed2 (delta2) | Ed1 x2
template <int NPROD>
__global__ void __launch_bounds__(TC_THREADS, 1) k_mlp_bwd_tc(const __grid_constant__ BwArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  BwSmem& S = *reinterpret_cast<BwSmem*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < LIDF_H1; i += TC_THREADS) S.u[i] = a.u ? a.u[i] : 0.f;
  for (int i = tid; i < LIDF_H2; i += TC_THREADS) S.b2[i] = a.b2[i];
  for (int i = tid; i < LIDF_H3; i += TC_THREADS) { S.b3[i] = a.b3[i]; S.w4[i] = a.w4[i]; }
  if (tid == 0) {
    for (int i = 0; i < BW_SLOTS; ++i) { tc::mbar_init(&S.w_full[i], 1); tc::mbar_init(&S.w_empty[i], 1); }
    tc::mbar_init(&S.a1_ready, TC_ROW_WARPS); tc::mbar_init(&S.a1_free, 1);
    for (int i = 0; i < 2; ++i) {
      tc::mbar_init(&S.x_full[i], 1); tc::mbar_init(&S.x_done[i], TC_ROW_WARPS); tc::mbar_init(&S.x2_full[i], 1);
    }
    tc::mbar_init(&S.y_full, 1); tc::mbar_init(&S.y_done, TC_ROW_WARPS);
    tc::mbar_init(&S.z_full, 1); tc::mbar_init(&S.z_done, TC_ROW_WARPS);
    tc::mbar_init(&S.y2_full, 1); tc::mbar_init(&S.y2_done, TC_ROW_WARPS);
    tc::mbar_init(&S.x2_done, TC_ROW_WARPS);
    tc::fence_barrier_init();
  }
  if (warp == TC_ROW_WARPS) tc::tmem_alloc(&S.tmem_base, 512);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = S.tmem_base;
  const int n_my_tiles = (a.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  constexpr int NCH = BW_CHUNKS_FWD + BW_CHUNKS_BWD;

  if (warp == TC_ROW_WARPS + 1) {
    // ================================ weight loader ================================
    if (tc::elect_one()) {
      uint32_t fill = 0;
      for (int t = 0; t < n_my_tiles; ++t) {
        for (int c = 0; c < NCH; ++c, ++fill) {
          const uint32_t slot = fill % BW_SLOTS;
          if (fill >= BW_SLOTS) bw_wait(&S.w_empty[slot], ((fill / BW_SLOTS) & 1u) ^ 1u);
          tc::mbar_arrive_expect_tx(&S.w_full[slot], TC_CHUNK_BYTES);
          const uint8_t* src = c < BW_CHUNKS_FWD ? a.wfwd + (size_t)c * TC_CHUNK_BYTES
                                                 : a.wbwd + (size_t)(c - BW_CHUNKS_FWD) * TC_CHUNK_BYTES;
          tc::bulk_g2s(S.w[slot], src, TC_CHUNK_BYTES, &S.w_full[slot]);
        }
      }
    }
  } else if (warp == TC_ROW_WARPS) {
    // ================================ MMA issuer ================================
    if (tc::elect_one()) {
      const uint32_t full0 = tc::smem_u32(&S.w_full[0]), empty0 = tc::smem_u32(&S.w_empty[0]);
      const uint64_t wd128 = tc::make_bdesc(tc::smem_u32(S.w[0]), 2048u, 128u);
      const uint64_t wd64 = tc::make_bdesc(tc::smem_u32(S.w[0]), 1024u, 128u);
      const uint64_t ad = tc::make_bdesc(tc::smem_u32(S.a1[0]), 2048u, 128u);
      constexpr uint32_t id128 = tc::make_idesc(128);
      uint32_t fill = 0;
      auto acquire = [&]() {
        const uint32_t slot = fill % BW_SLOTS;
        bw_wait_a(full0 + 8 * slot, (fill / BW_SLOTS) & 1u);
        ++fill;
        return slot;
      };
      for (int t = 0; t < n_my_tiles; ++t) {
        const uint32_t ph = (uint32_t)t & 1u;
        bw_wait(&S.a1_ready, ph);
        if (t > 0) bw_wait(&S.x2_done, ph ^ 1u);           // Ed1 of the previous tile has read X0 / X1
        tc::fence_after_sync();
        for (int hf = 0; hf < 2; ++hf) {                          // layer 1 (SS): PE operand x W1[:,pos] -> X0 / X1
          const uint32_t dcol = hf ? TC_COL_X1 : TC_COL_X0;
          for (int s = 0; s < TC_K1_STEPS; ++s) {
            const uint32_t slot = acquire();
            const uint64_t bhi = wd128 + ((slot * TC_CHUNK_BYTES) >> 4);
            const uint64_t ahi = ad + ((uint32_t)(s * 4096) >> 4);
            tc::mma_ss(tmem + dcol, ahi, bhi, id128, s == 0 ? 0u : 1u);
            if (NPROD == 3) {
              tc::mma_ss(tmem + dcol, ahi + ((uint32_t)TC_A1S_PART_BYTES >> 4), bhi, id128, 1u);
              tc::mma_ss(tmem + dcol, ahi, bhi + (4096u >> 4), id128, 1u);
            }
            tc::commit_a(empty0 + 8 * slot);
          }
          tc::commit(&S.x_full[hf]);
        }
        tc::commit(&S.a1_free);
        for (int kh = 0; kh < 2; ++kh) {                          // layer 2 (TS): h1 (in X) x W2 -> Y
          bw_wait(&S.x_done[kh], ph);
          tc::fence_after_sync();
          for (int s = 0; s < 8; ++s) {
            const uint32_t slot = acquire();
            bw_ts128<NPROD>(tmem, wd128, TC_COL_Y, (kh ? TC_COL_X1 : TC_COL_X0) + 16 * s, slot * TC_CHUNK_BYTES, kh == 0 && s == 0);
            tc::commit_a(empty0 + 8 * slot);
          }
        }
        tc::commit(&S.y_full);
        bw_wait(&S.y_done, ph);                             // layer 3 (TS): h2 (in Y) x W3 -> Z
        tc::fence_after_sync();
        for (int c = 0; c < 4; ++c) {
          const uint32_t slot = acquire();
          for (int j = 0; j < 2; ++j)
            bw_ts64<NPROD>(tmem, wd64, TC_COL_Z, TC_COL_Y + 16 * (2 * c + j), slot * TC_CHUNK_BYTES + j * 4096, c == 0 && j == 0);
          tc::commit_a(empty0 + 8 * slot);
        }
        tc::commit(&S.z_full);
        bw_wait(&S.z_done, ph);                             // dgrad 2 (TS): delta3 (in Z, K = 64) x W3^T -> Y
        tc::fence_after_sync();
        for (int s = 0; s < 4; ++s) {
          const uint32_t slot = acquire();
          bw_ts128<NPROD>(tmem, wd128, TC_COL_Y, TC_COL_Z + 16 * s, slot * TC_CHUNK_BYTES, s == 0);
          tc::commit_a(empty0 + 8 * slot);
        }
        tc::commit(&S.y2_full);
        bw_wait(&S.y2_done, ph);                            // dgrad 1 (TS): delta2 (in Y) x W2^T -> X0, X1
        tc::fence_after_sync();
        for (int hf = 0; hf < 2; ++hf) {
          for (int s = 0; s < 8; ++s) {
            const uint32_t slot = acquire();
            bw_ts128<NPROD>(tmem, wd128, hf ? TC_COL_X1 : TC_COL_X0, TC_COL_Y + 16 * s, slot * TC_CHUNK_BYTES, s == 0);
            tc::commit_a(empty0 + 8 * slot);
          }
          tc::commit(&S.x2_full[hf]);
        }
      }
    }
  } else if (warp < TC_ROW_WARPS) {
    // ================================ row warps ================================
    const int q = warp & 3, g = warp >> 2;
    const int row = q * 32 + lane;
    const uint32_t lane_addr = tmem + ((uint32_t)(q * 32) << 16);
    const uint32_t a1row_hi = tc::smem_u32(S.a1[0]) + row * 16, a1row_lo = tc::smem_u32(S.a1[1]) + row * 16;
    auto st_a1 = [&](int sk, const uint32_t* w) {
      tc::st_shared_v4(a1row_hi + sk * 4096, w[0], w[1], w[2], w[3]);
      tc::st_shared_v4(a1row_hi + sk * 4096 + 2048, w[4], w[5], w[6], w[7]);
      tc::st_shared_v4(a1row_lo + sk * 4096, w[8], w[9], w[10], w[11]);
      tc::st_shared_v4(a1row_lo + sk * 4096 + 2048, w[12], w[13], w[14], w[15]);
    };

    // ---- coalesced global <-> register-row transposes through this warp's staging tile
    float* const stg = S.stage[warp];
    const int tr = lane >> 3, tc4 = 4 * (lane & 7);              // instruction j touches rows 4 j + tr, floats [tc4, tc4 + 4)
    // rows (lanes) x 32 floats -> global rows gbase + r * pitch (+ col)
    auto store32 = [&](const float* x, float* gbase, int pitch) {
#pragma unroll
      for (int i = 0; i < 8; ++i)
        *reinterpret_cast<float4*>(stg + lane * BW_STAGE_PITCH + 4 * i) = make_float4(x[4 * i], x[4 * i + 1], x[4 * i + 2], x[4 * i + 3]);
      __syncwarp();
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int r = 4 * j + tr;
        *reinterpret_cast<float4*>(gbase + (size_t)r * pitch + tc4) = *reinterpret_cast<const float4*>(stg + r * BW_STAGE_PITCH + tc4);
      }
      __syncwarp();
    };
    // packed hand-over (bw_pk_*): 16 features [f0, f0 + 16) of this lane's row, w = 8 hi words | 8 lo words as split16 makes
    // them.  One instruction covers the warp's 32 consecutive rows x 16 B = 512 contiguous bytes: no transpose needed.
    auto store_pk16 = [&](float* tensor, int F, int tile_local, int f0, const uint32_t* w) {
      uint8_t* p = bw_pk_ptr(tensor, F, (int64_t)tile_local * 128 + row, f0);
      *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
      *reinterpret_cast<uint4*>(p + BW_PK_FG_BYTES) = make_uint4(w[4], w[5], w[6], w[7]);
      p += bw_pk_lo_offset(F);
      *reinterpret_cast<uint4*>(p) = make_uint4(w[8], w[9], w[10], w[11]);
      *reinterpret_cast<uint4*>(p + BW_PK_FG_BYTES) = make_uint4(w[12], w[13], w[14], w[15]);
    };
    auto load32 = [&](float* x, const float* gbase, int pitch) {    // global rows -> this lane's row
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int r = 4 * j + tr;
        *reinterpret_cast<float4*>(stg + r * BW_STAGE_PITCH + tc4) = *reinterpret_cast<const float4*>(gbase + (size_t)r * pitch + tc4);
      }
      __syncwarp();
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 v = *reinterpret_cast<const float4*>(stg + lane * BW_STAGE_PITCH + 4 * i);
        x[4 * i] = v.x; x[4 * i + 1] = v.y; x[4 * i + 2] = v.z; x[4 * i + 3] = v.w;
      }
      __syncwarp();
    };
    // Layer-1 addend A_v[vox][col ..) + T[ray][col ..) of this lane's row.  The A_v part is gathered ONE EPILOGUE AHEAD of its
    // use: 8 lanes copy 128 B of a row with cp.async straight into the staging tile (row indices via shuffles), so the
    // latency runs under the previous epilogue / the tail of the previous tile.  The T part is read by the lane for its own
    // row (lanes of one ray share the address); its lines were prefetched into L2 during the previous tile.
    const uint32_t stg_sa = tc::smem_u32(stg);
    auto gather_issue = [&](int vox_, int col) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int r = 4 * j + tr;
        const int vr = __shfl_sync(0xffffffffu, vox_, r);
        tc::cp_async16(stg_sa + (uint32_t)(r * BW_STAGE_PITCH + tc4) * 4u, a.Av + (size_t)vr * 512 + col + tc4);
      }
      tc::cp_async_commit();
    };
    auto gather_read = [&](float* x, int ray_, bool valid_, int col) {
      float4 tt[8];
      const float4* tp = reinterpret_cast<const float4*>(a.T + (size_t)ray_ * 512 + col);
#pragma unroll
      for (int i = 0; i < 8; ++i) tt[i] = valid_ ? __ldg(tp + i) : make_float4(0.f, 0.f, 0.f, 0.f);
      tc::cp_async_wait_all();
      __syncwarp();
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 v = *reinterpret_cast<const float4*>(stg + lane * BW_STAGE_PITCH + 4 * i);
        x[4 * i] = valid_ ? v.x + tt[i].x : 0.f; x[4 * i + 1] = valid_ ? v.y + tt[i].y : 0.f;
        x[4 * i + 2] = valid_ ? v.z + tt[i].z : 0.f; x[4 * i + 3] = valid_ ? v.w + tt[i].w : 0.f;
      }
      __syncwarp();                                                // the tile may be overwritten by the next gather / store
    };
    struct RowMeta { int orig, vox, ray; float t0, t1; bool valid; };
    auto load_meta = [&](int tile_local) {
      RowMeta m{0, 0, 0, 0.f, 0.f, false};
      const int rl = tile_local * 128 + row;
      if (rl < a.n_rows && a.perm[a.s0 + rl] >= 0) {
        m.valid = true;
        m.orig = a.perm[a.s0 + rl];
        m.vox = (int)lidf_clamp_idx(a.pair_vox[m.orig], a.V);      // out-of-range indices are flagged by the forward
        m.ray = (int)lidf_clamp_idx(a.pair_ray[m.orig], a.R);
        if (a.pair_dist) { const float2 t = *reinterpret_cast<const float2*>(a.pair_dist + 2 * (size_t)m.orig); m.t0 = t.x; m.t1 = t.y; }
        else { const size_t o = ((size_t)m.vox * a.R + m.ray) * 2; m.t0 = a.dense_dist[o]; m.t1 = a.dense_dist[o + 1]; }
      }
      return m;
    };
    // layer-1 MMA operand (identical to k_mlp_tc::build_a1) + optional fp32 copy of it for the dW1[:,pos] wgrad
    auto build_a1 = [&](const RowMeta& m, int tile_local) {
      float dir[3] = {0.f, 0.f, 0.f}, pe[3] = {0.f, 0.f, 0.f}, pl[3] = {0.f, 0.f, 0.f};
      if (m.valid) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          dir[k] = a.ray_dir[(size_t)m.ray * 3 + k];
          const float c = a.rel ? (a.voxel_bound[(size_t)m.vox * 6 + k] + a.voxel_bound[(size_t)m.vox * 6 + 3 + k]) / 2.0f : 0.f;
          pe[k] = dir[k] * m.t0 - c;
          pl[k] = dir[k] * m.t1 - c;
        }
      }
      if (g < 2) {
        float v[48];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const float pin = g == 0 ? pe[c] : pl[c];
#pragma unroll
          for (int fg = 0; fg < 2; ++fg) {
            float sn, cs;
            sincosf(pin * (fg ? 16.0f : 1.0f), &sn, &cs);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
              const int k = 4 * fg + kk;
              v[6 * k + c] = sn; v[6 * k + 3 + c] = cs;
              const float s2 = 2.0f * sn * cs, c2 = 1.0f - 2.0f * sn * sn;
              sn = s2; cs = c2;
            }
          }
        }
        if (!m.valid) {
#pragma unroll
          for (int k = 0; k < 48; ++k) v[k] = 0.f;               // rows past the chunk contribute nothing to the wgrad
        }
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          uint32_t w[16];
          tc::split16(v + 16 * j, w);
          st_a1(3 * g + j, w);
          if (a.pe) store_pk16(a.pe, BW_PE_F, tile_local, 16 * (3 * g + j), w);     // the same words, for dW1[:,pos]
        }
      } else if (g == 2) {
        float x[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) x[k] = 0.f;
#pragma unroll
        for (int k = 0; k < 3; ++k) { x[k] = pe[k]; x[3 + k] = pl[k]; }
        uint32_t w[16];
        tc::split16(x, w);
        st_a1(6, w);
        if (a.pe) store_pk16(a.pe, BW_PE_F, tile_local, 96, w);
      }
      else if (a.pe) {                                            // column group 3: the zero k-step that pads K to 128
        uint32_t w[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) w[k] = 0u;
        store_pk16(a.pe, BW_PE_F, tile_local, 112, w);
      }
      tc::fence_proxy_async();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&S.a1_ready);
    };

    float acc_k0 = 0.f, acc_b2 = 0.f, acc_du0 = 0.f, acc_du1 = 0.f, acc_b4 = 0.f;
    uint32_t par = 0;
    RowMeta cur = load_meta((int)blockIdx.x);
    build_a1(cur, (int)blockIdx.x);
    if (n_my_tiles > 0) gather_issue(cur.vox, a.dcol + 32 * g);
    for (int t = 0; t < n_my_tiles; ++t) {
      const int tile_local = (int)blockIdx.x + t * (int)gridDim.x;
      const uint32_t ph = (uint32_t)t & 1u;
      const bool has_next = t + 1 < n_my_tiles;
      const size_t wrow0 = (size_t)tile_local * 128 + q * 32;    // first of this warp's 32 chunk rows
      const bool valid = cur.valid;
      const int64_t gidx = a.by_slot ? a.s0 + (int64_t)tile_local * 128 + row : (int64_t)cur.orig;
      const float gin = valid ? a.g[gidx] : 0.f;
      const float oin = (valid && a.o_in) ? a.o_in[gidx] : a.o0;
      const float delta = oin - a.o0;
      const bool rank1 = a.is_ief && a.it > 0;
      RowMeta nxt{0, 0, 0, 0.f, 0.f, false};
      if (has_next) nxt = load_meta(tile_local + (int)gridDim.x);
      uint32_t mask1a = 0u, mask1b = 0u, mask2 = 0u;
      // ---- E1 x 2: h1 = leaky(acc + A_v + T (+ u delta)) -> HBM row, sign mask, bf16 hi|lo in place
#pragma unroll 1
      for (int hf = 0; hf < 2; ++hf) {
        const int n0 = 128 * hf + 32 * g;
        float xo[32];
        gather_read(xo, cur.ray, valid, a.dcol + n0);             // A_v + T, before waiting for the accumulator
        if (hf == 0) gather_issue(cur.vox, a.dcol + 128 + 32 * g);   // second half's A_v rows: in flight during this epilogue
        const uint32_t xcol = (hf ? TC_COL_X1 : TC_COL_X0) + 32 * g;
        bw_wait(&S.x_full[hf], ph);
        tc::fence_after_sync();
        uint32_t r[32];
        tc::tmem_ld32(lane_addr + xcol, r);
        tc::wait_ld();
        uint32_t mk = 0u;
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          float v = __uint_as_float(r[e]) + xo[e];
          if (rank1) v = fmaf(S.u[n0 + e], delta, v);
          const float x = fmaxf(v, LIDF_LEAKY * v);
          xo[e] = x;
          mk |= (x > 0.f ? 1u : 0u) << e;
        }
        if (hf == 0) mask1a = mk; else mask1b = mk;
#pragma unroll
        for (int s16 = 0; s16 < 2; ++s16) {
          uint32_t w[16];
          tc::split16(xo + 16 * s16, w);
          tc::tmem_st16(lane_addr + xcol + 16 * s16, w);
          store_pk16(a.h1, LIDF_H1, tile_local, n0 + 16 * s16, w);   // the same words go to the wgrad kernel (fire and forget)
        }
        tc::wait_st();
        tc::fence_before_sync();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&S.x_done[hf]);
      }
      // ---- operand of the next tile (its last reader, L1 of this tile, has retired once a1_free completes)
      if (has_next) {
        if (nxt.valid)                                            // next tile's T lines -> L2 (8 lines per ray and decoder, one per lane)
          asm volatile("prefetch.global.L2 [%0];" ::"l"(a.T + (size_t)nxt.ray * 512 + a.dcol + 32 * (row & 7)));
        bw_wait(&S.a1_free, ph);
        build_a1(nxt, tile_local + (int)gridDim.x);
      }
      // ---- E2: h2 = leaky(acc + b2)
      {
        bw_wait(&S.y_full, ph);
        tc::fence_after_sync();
        uint32_t r[32];
        tc::tmem_ld32(lane_addr + TC_COL_Y + 32 * g, r);
        tc::wait_ld();
        float xo[32];
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          const float v = __uint_as_float(r[e]) + S.b2[32 * g + e];
          const float x = fmaxf(v, LIDF_LEAKY * v);
          xo[e] = x;
          mask2 |= (x > 0.f ? 1u : 0u) << e;
        }
#pragma unroll
        for (int s16 = 0; s16 < 2; ++s16) {
          uint32_t w[16];
          tc::split16(xo + 16 * s16, w);
          tc::tmem_st16(lane_addr + TC_COL_Y + 32 * g + 16 * s16, w);
          store_pk16(a.h2, LIDF_H2, tile_local, 32 * g + 16 * s16, w);
        }
        tc::wait_st();
        tc::fence_before_sync();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&S.y_done);
      }
      // ---- E3: z3 -> delta3 = g w4 leaky'(z3) (this thread: columns [16 g, 16 g + 16)), in place as the D2 operand
      {
        bw_wait(&S.z_full, ph);
        tc::fence_after_sync();
        uint32_t r[16];
        tc::tmem_ld16(lane_addr + TC_COL_Z + 16 * g, r);
        tc::wait_ld();
        float d3[16], red[32];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int n = 16 * g + j;
          const float z = __uint_as_float(r[j]) + S.b3[n];
          const bool pos = z > 0.f;
          const float h3 = pos ? z : LIDF_LEAKY * z;
          d3[j] = gin * S.w4[n] * (pos ? 1.0f : LIDF_LEAKY);
          red[j] = d3[j];
          red[16 + j] = gin * h3;
        }
        uint32_t w[16];
        tc::split16(d3, w);
        tc::tmem_st16(lane_addr + TC_COL_Z + 16 * g, w);
        store_pk16(a.d3, LIDF_H3, tile_local, 16 * g, w);
        tc::wait_st();
        tc::fence_before_sync();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&S.z_done);
        acc_k0 += bw_colreduce32(red, lane);                      // lanes 0-15: db3[16 g + l], 16-31: dw4[16 g + l - 16]
        if (g == 0) acc_b4 += gin;
      }
      // ---- Ed2: delta2 = (delta3 W3) leaky'(z2)
      {
        bw_wait(&S.y2_full, ph);
        tc::fence_after_sync();
        uint32_t r[32];
        tc::tmem_ld32(lane_addr + TC_COL_Y + 32 * g, r);
        tc::wait_ld();
        float d2[32];
#pragma unroll
        for (int e = 0; e < 32; ++e) d2[e] = __uint_as_float(r[e]) * (((mask2 >> e) & 1u) ? 1.0f : LIDF_LEAKY);
#pragma unroll
        for (int s16 = 0; s16 < 2; ++s16) {
          uint32_t w[16];
          tc::split16(d2 + 16 * s16, w);
          tc::tmem_st16(lane_addr + TC_COL_Y + 32 * g + 16 * s16, w);
          store_pk16(a.d2, LIDF_H2, tile_local, 32 * g + 16 * s16, w);
        }
        tc::wait_st();
        tc::fence_before_sync();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&S.y2_done);
        acc_b2 += bw_colreduce32(d2, lane);                       // db2[32 g + l]
      }
      // ---- Ed1 x 2: delta1 = (delta2 W2) leaky'(z1) -> HBM (summed over the IEF passes), du, IEF feedback
      float fb = 0.f;
#pragma unroll 1
      for (int hf = 0; hf < 2; ++hf) {
        const int n0 = 128 * hf + 32 * g;
        bw_wait(&S.x2_full[hf], ph);
        tc::fence_after_sync();
        uint32_t r[32];
        tc::tmem_ld32(lane_addr + (hf ? TC_COL_X1 : TC_COL_X0) + 32 * g, r);
        tc::wait_ld();
        if (hf == 1) {                                            // X0 / X1 are in registers: layer 1 of the next tile may overwrite them
          tc::fence_before_sync();                                // while this tile's delta1 is still being stored
          __syncwarp();
          if (lane == 0) tc::mbar_arrive(&S.x2_done);
        }
        const uint32_t mk = hf == 0 ? mask1a : mask1b;
        float d1[32];
#pragma unroll
        for (int e = 0; e < 32; ++e) d1[e] = __uint_as_float(r[e]) * (((mk >> e) & 1u) ? 1.0f : LIDF_LEAKY);
        float* dwarp = a.d1 + wrow0 * LIDF_H1 + n0;
        if (a.d1_accumulate) {
          float old[32];
          load32(old, dwarp, LIDF_H1);
#pragma unroll
          for (int e = 0; e < 32; ++e) old[e] += d1[e];
          store32(old, dwarp, LIDF_H1);
          if (a.d1pk) {
#pragma unroll
            for (int s16 = 0; s16 < 2; ++s16) {
              uint32_t w[16];
              tc::split16(old + 16 * s16, w);
              store_pk16(a.d1pk, LIDF_H1, tile_local, n0 + 16 * s16, w);
            }
          }
        } else {
          store32(d1, dwarp, LIDF_H1);
          if (a.d1pk) {
#pragma unroll
            for (int s16 = 0; s16 < 2; ++s16) {
              uint32_t w[16];
              tc::split16(d1 + 16 * s16, w);
              store_pk16(a.d1pk, LIDF_H1, tile_local, n0 + 16 * s16, w);
            }
          }
        }
        if (a.is_ief) {
#pragma unroll
          for (int e = 0; e < 32; ++e) { fb = fmaf(S.u[n0 + e], d1[e], fb); d1[e] *= oin; }
          const float du = bw_colreduce32(d1, lane);              // du[n0 + l] += sum_rows delta1 o_in
          if (hf == 0) acc_du0 += du; else acc_du1 += du;
        }
      }
      if (has_next) gather_issue(nxt.vox, a.dcol + 32 * g);       // first half of the next tile (the staging tile is free again)
      if (rank1) {                                                // dL/d o_{it-1} = dL/d o_it + u . delta1
        S.part[par][g][row] = fb;
        tc::bar_quadrant(q);
        if (g == 0 && valid)
          a.g[gidx] = gin + (((S.part[par][0][row] + S.part[par][1][row]) + S.part[par][2][row]) + S.part[par][3][row]);
        par ^= 1;
      }
      cur = nxt;
    }
    float* cp = a.colpart + (((size_t)blockIdx.x * TC_ROW_WARPS + warp) * 32 + lane) * BW_COLPART;
    cp[0] += acc_k0; cp[1] += acc_b2; cp[2] += acc_du0; cp[3] += acc_du1; cp[4] += acc_b4;
  }
  __syncwarp();
  tc::fence_before_sync();
  __syncthreads();
  if (warp == TC_ROW_WARPS) tc::tmem_dealloc(tmem, 512);
}

// column-sum partials -> db2, db3, dw4, db4, du (fixed summation order: cta, then quadrant)
__global__ void k_bwd_colpart_finish(const float* __restrict__ colpart, int n_cta, float* __restrict__ db2, float* __restrict__ db3,
                                     float* __restrict__ dw4, float* __restrict__ db4, float* __restrict__ du) {
  const int t = threadIdx.x;                 // 512 threads: (g = t / 128, kind/col by t % 128)
  auto sum = [&](int g, int l, int k) {
    float s = 0.f;
    for (int c = 0; c < n_cta; ++c)
      for (int q = 0; q < 4; ++q) s += colpart[(((size_t)c * TC_ROW_WARPS + (q + 4 * g)) * 32 + l) * BW_COLPART + k];
    return s;
  };
  if (t < 128) { const int g = t / 32, l = t % 32; if (l < 16) db3[16 * g + l] = sum(g, l, 0); else dw4[16 * g + l - 16] = sum(g, l, 0); }
  else if (t < 256) { const int g = (t - 128) / 32, l = t % 32; db2[32 * g + l] = sum(g, l, 1); }
  else if (t < 384) { const int g = (t - 256) / 32, l = t % 32; if (du) du[32 * g + l] = sum(g, l, 2); }
  else if (t < 512) { const int g = (t - 384) / 32, l = t % 32; if (du) du[128 + 32 * g + l] = sum(g, l, 3); }
  if (t == 0) {
    float s = 0.f;
    for (int c = 0; c < n_cta; ++c)
      for (int q = 0; q < 4; ++q)
        for (int l = 0; l < 32; ++l) s += colpart[(((size_t)c * TC_ROW_WARPS + q) * 32 + l) * BW_COLPART + 4];
    db4[0] = s;
  }
}

// ------------------------------------------------------------------------------------------------ B2
// C[M][N] += sum_rows A[row][m] B[row][n].  M = 128 MB (MB = 1, 2), N a multiple of 16 <= 256.  Persistent CTAs, each
// owning a contiguous range of 64-row groups; 16 loader warps convert fp32 -> bf16 hi/lo into the K-major UMMA layout
// (K = row index: lane = feature, 8 rows -> one 16-byte core-matrix row), double-buffered stages; one thread issues
// 4 k-steps x MB x 3 MMAs per stage into accumulators that stay in TMEM for the whole kernel.
#define WG_ROWS 64
#define WG_LOAD_WARPS 16
#define WG_THREADS ((WG_LOAD_WARPS + 1) * 32)
struct WgArgs {
  const float* A; int lda; int M;
  const float* B; int ldb; int N; int n_valid;      // columns of B >= n_valid are read as zero (N padded to 16)
  int64_t rows;
  float* partial;                                   // [grid][M * N], accumulated
};
struct WgSmemHdr { uint64_t full[2], empty[2], done; uint32_t tmem_base; };

template <int NPROD>
__global__ void __launch_bounds__(WG_THREADS, 1) k_wgrad_tc(const __grid_constant__ WgArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  WgSmemHdr& S = *reinterpret_cast<WgSmemHdr*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int MB = a.M / 128, N = a.N;
  const uint32_t a_bytes = (uint32_t)MB * 4 * 8192, b_bytes = 4u * (uint32_t)N * 64u;   // per stage
  const uint32_t stage_bytes = a_bytes + b_bytes;
  const uint32_t data0 = (tc::smem_u32(smem_raw) + 1024u);                                // stages start 1 KB in
  if (tid == 0) {
    for (int i = 0; i < 2; ++i) { tc::mbar_init(&S.full[i], WG_LOAD_WARPS); tc::mbar_init(&S.empty[i], 1); }
    tc::mbar_init(&S.done, 1);
    tc::fence_barrier_init();
  }
  if (warp == WG_LOAD_WARPS) tc::tmem_alloc(&S.tmem_base, 512);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = S.tmem_base;
  const int64_t n_groups = (a.rows + WG_ROWS - 1) / WG_ROWS;
  const int64_t per = (n_groups + gridDim.x - 1) / gridDim.x;
  const int64_t g0 = (int64_t)blockIdx.x * per, g1 = g0 + per < n_groups ? g0 + per : n_groups;
  const int n_it = g1 > g0 ? (int)(g1 - g0) : 0;

  if (warp == WG_LOAD_WARPS) {
    if (tc::elect_one() && n_it > 0) {
      const uint32_t idesc = tc::make_idesc(N);
      for (int it = 0; it < n_it; ++it) {
        const int buf = it & 1;
        bw_wait(&S.full[buf], (uint32_t)(it >> 1) & 1u);
        tc::fence_after_sync();
        const uint32_t sa = data0 + buf * stage_bytes, sb = sa + a_bytes;
#pragma unroll 1
        for (int ks = 0; ks < 4; ++ks) {
          const uint64_t bhi = tc::make_bdesc(sb + ks * (uint32_t)N * 64u, (uint32_t)N * 16u, 128u);
          const uint64_t blo = tc::make_bdesc(sb + ks * (uint32_t)N * 64u + (uint32_t)N * 32u, (uint32_t)N * 16u, 128u);
          for (int mb = 0; mb < MB; ++mb) {
            const uint32_t abase = sa + (uint32_t)(mb * 4 + ks) * 8192u;
            const uint64_t ahi = tc::make_bdesc(abase, 2048u, 128u), alo = tc::make_bdesc(abase + 4096u, 2048u, 128u);
            const uint32_t acc = (it > 0 || ks > 0) ? 1u : 0u;
            tc::mma_ss(tmem + mb * N, ahi, bhi, idesc, acc);
            if (NPROD == 3) {
              tc::mma_ss(tmem + mb * N, alo, bhi, idesc, 1u);
              tc::mma_ss(tmem + mb * N, ahi, blo, idesc, 1u);
            }
          }
        }
        tc::commit(&S.empty[buf]);
      }
      tc::commit(&S.done);
    }
  } else {
    // ---- loaders: unit = (8-row group rg, 32-feature block); A blocks first, then B blocks
    const int a_blocks = a.M / 32, b_blocks = (N + 31) / 32, n_units = (a_blocks + b_blocks) * 8;
    // A warp's units of a stage come in rounds of three (24 independent 128-byte loads in flight before any conversion: the
    // kernel is bound by memory-level parallelism, not by instruction issue).  The first round of stage it + 1 is issued
    // before the hand-over of stage it, so its latency runs under the fence, the arrive and the wait for the buffer.
    auto unit_load = [&](int64_t row0, int un, float (&x)[8]) {
      const int blk = un >> 3, rg = un & 7;
      const bool isA = blk < a_blocks;
      const int f = (isA ? blk : blk - a_blocks) * 32 + lane;
      const float* src = isA ? a.A : a.B;
      const int ld = isA ? a.lda : a.ldb;
      const bool fok = isA ? true : (f < a.n_valid);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int64_t rr = row0 + 8 * rg + i;
        x[i] = (fok && rr < a.rows) ? __ldg(src + (size_t)rr * ld + f) : 0.f;
      }
    };
    auto unit_store = [&](uint32_t sa, uint32_t sb, int un, const float (&x)[8]) {
      const int blk = un >> 3, rg = un & 7;
      const bool isA = blk < a_blocks;
      const int f = (isA ? blk : blk - a_blocks) * 32 + lane;
      uint32_t h[4], l[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) tc::split2(x[2 * j], x[2 * j + 1], h[j], l[j]);
      const int ks = rg >> 1, kg = rg & 1;
      if (isA) {
        const uint32_t base = sa + (uint32_t)((f >> 7) * 4 + ks) * 8192u + kg * 2048u + (uint32_t)(f & 127) * 16u;
        tc::st_shared_v4(base, h[0], h[1], h[2], h[3]);
        tc::st_shared_v4(base + 4096u, l[0], l[1], l[2], l[3]);
      } else if (f < N) {
        const uint32_t base = sb + ks * (uint32_t)N * 64u + kg * (uint32_t)N * 16u + (uint32_t)f * 16u;
        tc::st_shared_v4(base, h[0], h[1], h[2], h[3]);
        tc::st_shared_v4(base + (uint32_t)N * 32u, l[0], l[1], l[2], l[3]);
      }
    };
    float x0[8], x1[8], x2[8];
    auto round_load = [&](int64_t row0, int un) {
      unit_load(row0, un, x0);
      if (un + WG_LOAD_WARPS < n_units) unit_load(row0, un + WG_LOAD_WARPS, x1);
      if (un + 2 * WG_LOAD_WARPS < n_units) unit_load(row0, un + 2 * WG_LOAD_WARPS, x2);
    };
    auto round_store = [&](uint32_t sa, uint32_t sb, int un) {
      unit_store(sa, sb, un, x0);
      if (un + WG_LOAD_WARPS < n_units) unit_store(sa, sb, un + WG_LOAD_WARPS, x1);
      if (un + 2 * WG_LOAD_WARPS < n_units) unit_store(sa, sb, un + 2 * WG_LOAD_WARPS, x2);
    };
    if (n_it > 0 && warp < n_units) round_load(g0 * WG_ROWS, warp);
    for (int it = 0; it < n_it; ++it) {
      const int buf = it & 1;
      if (it >= 2) bw_wait(&S.empty[buf], (uint32_t)((it >> 1) - 1) & 1u);
      const uint32_t sa = data0 + buf * stage_bytes, sb = sa + a_bytes;
      const int64_t row0 = (g0 + it) * WG_ROWS;
      if (warp < n_units) round_store(sa, sb, warp);                          // round 0: loaded one stage ahead
      for (int un = warp + 3 * WG_LOAD_WARPS; un < n_units; un += 3 * WG_LOAD_WARPS) {
        round_load(row0, un);
        round_store(sa, sb, un);
      }
      if (it + 1 < n_it && warp < n_units) round_load(row0 + WG_ROWS, warp);
      tc::fence_proxy_async();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&S.full[buf]);
    }
    // ---- epilogue: TMEM -> this CTA's partial slice (accumulated across launches)
    if (n_it > 0) {
      bw_wait(&S.done, 0);
      tc::fence_after_sync();
      const int q = warp & 3, cg = warp >> 2;
      const uint32_t lane_addr = tmem + ((uint32_t)(q * 32) << 16);
      float* part = a.partial + (size_t)blockIdx.x * a.M * N;
      const int nblk = MB * N / 16;
      for (int blk = cg; blk < nblk; blk += 4) {
        uint32_t r[16];
        tc::tmem_ld16(lane_addr + 16 * blk, r);
        tc::wait_ld();
        const int col = 16 * blk, mb = col / N, n = col % N;
        float* dst = part + (size_t)(mb * 128 + q * 32 + lane) * N + n;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float4 o = *reinterpret_cast<float4*>(dst + 4 * i);
          o.x += __uint_as_float(r[4 * i]); o.y += __uint_as_float(r[4 * i + 1]);
          o.z += __uint_as_float(r[4 * i + 2]); o.w += __uint_as_float(r[4 * i + 3]);
          *reinterpret_cast<float4*>(dst + 4 * i) = o;
        }
      }
    }
  }
  __syncwarp();
  tc::fence_before_sync();
  __syncthreads();
  if (warp == WG_LOAD_WARPS) tc::tmem_dealloc(tmem, 512);
}

// ------------------------------------------------------------------------------------------------ B2, packed operands
// C[128][N] += sum_rows A[row][m] B[row][n] with both operands in the PK layout (above): a stage = one 64-row group of A
// and of B, fetched by two bulk copies (TMA) straight into the MN-major operand position; 4 k-steps x 3 products per
// stage; accumulator resident in TMEM over all of the CTA's groups; per-CTA partial slices as in k_wgrad_tc.
//   warps 0-7 : epilogue (TMEM -> partial slice), idle until the last MMA has retired
//   warp 8    : producer, one thread: ring of n_stages (2-4, whatever fits in shared memory)
//   warp 9    : MMA issuer, one thread
// The kernel moves 4 (FA + FB) bytes per row and does nothing else with them: it is bound by HBM reads.
#define WPK_EPI_WARPS 8
#define WPK_THREADS ((WPK_EPI_WARPS + 2) * 32)
#define WPK_MAX_STAGES 4
struct WgPkArgs {
  const uint8_t* A; int FA;           // FA = 128 (M)
  const uint8_t* B; int FB;           // FB = N, a multiple of 16, <= 256
  int64_t groups;                     // 64-row groups (whole tiles: B1 writes every row of a tile, dead rows as zeros in delta)
  int n_stages;
  float* partial;                     // [grid][128 * N], accumulated
};
struct WgPkSmemHdr { uint64_t full[WPK_MAX_STAGES], empty[WPK_MAX_STAGES], done; uint32_t tmem_base; };

namespace tc {
__device__ __forceinline__ void bulk_g2s_a(uint32_t dst_saddr, const void* src_gmem, uint32_t bytes, uint32_t bar_saddr) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_saddr),
               "l"(src_gmem), "r"(bytes), "r"(bar_saddr)
               : "memory");
}
// instruction descriptor with BOTH operands MN-major (bits 15, 16: cute::UMMA::InstrDescriptor a_major_ / b_major_)
__host__ __device__ constexpr uint32_t make_idesc_mn(int N) { return make_idesc(N) | (1u << 15) | (1u << 16); }
}  // namespace tc

template <int NPROD>
__global__ void __launch_bounds__(WPK_THREADS, 1) k_wgrad_pk_tc(const __grid_constant__ WgPkArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  WgPkSmemHdr& S = *reinterpret_cast<WgPkSmemHdr*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int N = a.FB, NS = a.n_stages;
  const uint32_t a_bytes = (uint32_t)bw_pk_group_bytes(a.FA), b_bytes = (uint32_t)bw_pk_group_bytes(a.FB);
  const uint32_t stage_bytes = a_bytes + b_bytes;
  const uint32_t data0 = tc::smem_u32(smem_raw) + 1024u;
  if (tid == 0) {
    for (int i = 0; i < WPK_MAX_STAGES; ++i) { tc::mbar_init(&S.full[i], 1); tc::mbar_init(&S.empty[i], 1); }
    tc::mbar_init(&S.done, 1);
    tc::fence_barrier_init();
  }
  if (warp == WPK_EPI_WARPS + 1) tc::tmem_alloc(&S.tmem_base, 512);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = S.tmem_base;
  const int64_t per = (a.groups + gridDim.x - 1) / gridDim.x;
  const int64_t g0 = (int64_t)blockIdx.x * per, g1 = g0 + per < a.groups ? g0 + per : a.groups;
  const int n_it = g1 > g0 ? (int)(g1 - g0) : 0;

  if (warp == WPK_EPI_WARPS) {
    // ---- producer
    if (tc::elect_one()) {
      for (int it = 0; it < n_it; ++it) {
        const int s = it % NS;
        if (it >= NS) bw_wait(&S.empty[s], (uint32_t)(it / NS - 1) & 1u);
        const uint32_t bar = tc::smem_u32(&S.full[s]);
        tc::mbar_arrive_expect_tx(&S.full[s], stage_bytes);
        tc::bulk_g2s_a(data0 + s * stage_bytes, a.A + (size_t)(g0 + it) * a_bytes, a_bytes, bar);
        tc::bulk_g2s_a(data0 + s * stage_bytes + a_bytes, a.B + (size_t)(g0 + it) * b_bytes, b_bytes, bar);
      }
    }
  } else if (warp == WPK_EPI_WARPS + 1) {
    // ---- MMA issuer: D[m][n] (+)= sum over 16 rows of A[row][m] B[row][n], operands MN-major
    if (tc::elect_one() && n_it > 0) {
      const uint32_t idesc = tc::make_idesc_mn(N);
      const uint32_t a_lo = (uint32_t)bw_pk_lo_offset(a.FA), b_lo = (uint32_t)bw_pk_lo_offset(a.FB);
      for (int it = 0; it < n_it; ++it) {
        const int s = it % NS;
        bw_wait(&S.full[s], (uint32_t)(it / NS) & 1u);
        tc::fence_after_sync();
        const uint32_t sa = data0 + s * stage_bytes, sb = sa + a_bytes;
#pragma unroll 1
        for (int ks = 0; ks < BW_PK_ROWS / 16; ++ks) {
          // k-step = 16 rows = two 8-row core matrices 128 B apart (LBO); feature groups 1 KB apart (SBO)
          const uint64_t ahi = tc::make_bdesc(sa + ks * 256u, 128u, (uint32_t)BW_PK_FG_BYTES);
          const uint64_t alo = tc::make_bdesc(sa + a_lo + ks * 256u, 128u, (uint32_t)BW_PK_FG_BYTES);
          const uint64_t bhi = tc::make_bdesc(sb + ks * 256u, 128u, (uint32_t)BW_PK_FG_BYTES);
          const uint64_t blo = tc::make_bdesc(sb + b_lo + ks * 256u, 128u, (uint32_t)BW_PK_FG_BYTES);
          tc::mma_ss(tmem, ahi, bhi, idesc, (it > 0 || ks > 0) ? 1u : 0u);
          if (NPROD == 3) {
            tc::mma_ss(tmem, alo, bhi, idesc, 1u);
            tc::mma_ss(tmem, ahi, blo, idesc, 1u);
          }
        }
        tc::commit(&S.empty[s]);
      }
      tc::commit(&S.done);
    }
  } else if (n_it > 0) {
    // ---- epilogue: TMEM -> this CTA's partial slice (accumulated across launches)
    bw_wait(&S.done, 0);
    tc::fence_after_sync();
    const int q = warp & 3, cg = warp >> 2;
    const uint32_t lane_addr = tmem + ((uint32_t)(q * 32) << 16);
    float* part = a.partial + (size_t)blockIdx.x * 128 * N;
    for (int blk = cg; blk < N / 16; blk += WPK_EPI_WARPS / 4) {
      uint32_t r[16];
      tc::tmem_ld16(lane_addr + 16 * blk, r);
      tc::wait_ld();
      float* dst = part + (size_t)(q * 32 + lane) * N + 16 * blk;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float4 o = *reinterpret_cast<float4*>(dst + 4 * i);
        o.x += __uint_as_float(r[4 * i]); o.y += __uint_as_float(r[4 * i + 1]);
        o.z += __uint_as_float(r[4 * i + 2]); o.w += __uint_as_float(r[4 * i + 3]);
        *reinterpret_cast<float4*>(dst + 4 * i) = o;
      }
    }
  }
  __syncwarp();
  tc::fence_before_sync();
  __syncthreads();
  if (warp == WPK_EPI_WARPS + 1) tc::tmem_dealloc(tmem, 512);
}

// fp32 rows [rows][F] -> PK layout over ceil(rows / 64) groups, rows past the end as zeros (self test of k_wgrad_pk_tc)
__global__ void k_pk_pack_rows(const float* __restrict__ X, int64_t rows, int F, uint8_t* __restrict__ out) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;        // one thread per (padded row, feature group)
  const int fgs = F / 8;
  const int64_t rows_pad = (rows + BW_PK_ROWS - 1) / BW_PK_ROWS * BW_PK_ROWS;
  if (idx >= rows_pad * fgs) return;
  const int64_t r = idx / fgs;
  const int fg = (int)(idx % fgs);
  float x[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) x[k] = r < rows ? X[(size_t)r * F + 8 * fg + k] : 0.f;
  uint32_t h[4], l[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) tc::split2(x[2 * j], x[2 * j + 1], h[j], l[j]);
  uint8_t* p = bw_pk_ptr(reinterpret_cast<float*>(out), F, r, 8 * fg);
  *reinterpret_cast<uint4*>(p) = make_uint4(h[0], h[1], h[2], h[3]);
  *reinterpret_cast<uint4*>(p + bw_pk_lo_offset(F)) = make_uint4(l[0], l[1], l[2], l[3]);
}

// partial slices -> gradient tensor.  mode: 0 dst[n * 128 + m] (dW3 = C^T), 1 dst[m * ldd + n] for n < n_keep (dense rows
// of a [M, ldd] weight starting at dst), 2 layer-1 PE columns: dst[m * ldd + tc_a1_col(n)] (4: its transpose, dst[n * ldd +
// tc_a1_col(m)]), 3 as 1 plus column `ones_col`
// of C -> extra[m] (column sums through the ones column of the PE(dir) block)
struct WgFinishArgs { const float* partial; int n_cta; int M, N; int mode; float* dst; int ldd; int n_keep; int pe_pos; int ones_col; float* extra; };
__global__ void k_wgrad_finish(const WgFinishArgs a) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= a.M * a.N) return;
  const int m = idx / a.N, n = idx % a.N;
  float s = 0.f;
  for (int c = 0; c < a.n_cta; ++c) s += a.partial[(size_t)c * a.M * a.N + idx];
  if (a.mode == 0) a.dst[(size_t)n * 128 + m] = s;
  else if (a.mode == 2) { const int col = tc_a1_col(n, a.pe_pos, 0); if (col >= 0) a.dst[(size_t)m * a.ldd + col] = s; }
  else if (a.mode == 4) { const int col = tc_a1_col(m, a.pe_pos, 0); if (col >= 0) a.dst[(size_t)n * a.ldd + col] = s; }   // C = PE^T delta1
  else {
    if (n < a.n_keep) a.dst[(size_t)m * a.ldd + n] = s;
    if (a.mode == 3 && n == a.ones_col) a.extra[m] = s;
  }
}

// ------------------------------------------------------------------------------------------------ segment sums
// G_r[ray][dcol .. dcol + 256) += sum over the ray's pairs inside the chunk of delta1 (chunk rows are ray-major)
__global__ void k_segsum_rays(const float* __restrict__ d1, int64_t s0, int n_rows, const int* __restrict__ ray_start, int64_t R,
                              float* __restrict__ G, int dcol) {
  const int64_t ray = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (ray >= R) return;
  int64_t s = ray_start[ray], e = ray_start[ray + 1];
  if (s < s0) s = s0;
  if (e > s0 + n_rows) e = s0 + n_rows;
  if (e <= s) return;
  float4 acc0 = make_float4(0.f, 0.f, 0.f, 0.f), acc1 = acc0;
  for (int64_t i = s; i < e; ++i) {
    const float4* p = reinterpret_cast<const float4*>(d1 + (size_t)(i - s0) * LIDF_H1 + 8 * lane);
    const float4 v0 = __ldg(p), v1 = __ldg(p + 1);
    acc0.x += v0.x; acc0.y += v0.y; acc0.z += v0.z; acc0.w += v0.w;
    acc1.x += v1.x; acc1.y += v1.y; acc1.z += v1.z; acc1.w += v1.w;
  }
  float4* o = reinterpret_cast<float4*>(G + (size_t)ray * 512 + dcol + 8 * lane);
  float4 t0 = o[0], t1 = o[1];
  t0.x += acc0.x; t0.y += acc0.y; t0.z += acc0.z; t0.w += acc0.w;
  t1.x += acc1.x; t1.y += acc1.y; t1.z += acc1.z; t1.w += acc1.w;
  o[0] = t0; o[1] = t1;
}
// chunk-local voxel keys for the CSR-by-voxel of a chunk
__global__ void k_chunk_vox_keys(const int* __restrict__ perm, const int64_t* __restrict__ pair_vox, int64_t s0, int n_rows,
                                 int64_t V, int64_t* __restrict__ keys) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_rows) return;
  const int o = perm[s0 + i];
  int64_t v = o >= 0 ? pair_vox[o] : 0;              // slot without a row (winner-only list): its delta1 row is zero
  keys[i] = v < 0 ? 0 : (v >= V ? V - 1 : v);
}
// winner-only backward: the offset decoder's row list (one row per ray) and its seeds, by ray
__global__ void k_bwd_winner_rows(const int64_t* __restrict__ max_pair_id, int64_t P, int64_t R, int* __restrict__ win,
                                  int* __restrict__ iota) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r > R) return;
  iota[r] = (int)r;
  if (r < R) { const int64_t m = max_pair_id[r]; win[r] = (m >= 0 && m < P) ? (int)m : -1; }
}
__global__ void k_bwd_seed_rays(const float* __restrict__ g_pred_pos, const float* __restrict__ ray_dir, const float* __restrict__ off_ray,
                                const int* __restrict__ win, int64_t R, float scale, int sig0, float* __restrict__ g0r) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  float go = 0.f;
  if (g_pred_pos && win[r] >= 0) {
    const float dot = g_pred_pos[3 * r] * ray_dir[3 * r] + g_pred_pos[3 * r + 1] * ray_dir[3 * r + 1] + g_pred_pos[3 * r + 2] * ray_dir[3 * r + 2];
    go = dot * scale * lidf_final_act_grad(off_ray[r], sig0);
  }
  g0r[r] = go;
}
// G_v[vox][dcol ..) += delta1 rows grouped by voxel: block = 64 consecutive positions of the voxel-sorted row list
// (order[], seg_start[V+1]); a running sum is flushed with atomics whenever the voxel changes (segments are long, so
// nearly always once per block).  64 threads x 4 columns.
__global__ void __launch_bounds__(64) k_segsum_vox(const float* __restrict__ d1, const int* __restrict__ order,
                                                   const int* __restrict__ seg_start, int64_t V, int n_rows,
                                                   float* __restrict__ G, int dcol) {
  __shared__ int s_v;
  const int p0 = blockIdx.x * 64, p1 = min(p0 + 64, n_rows);
  if (threadIdx.x == 0) {                       // voxel of position p0: last v with seg_start[v] <= p0
    int lo = 0, hi = (int)V;
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (seg_start[mid] <= p0) lo = mid; else hi = mid; }
    s_v = lo;
  }
  __syncthreads();
  int v = s_v;
  int vend = seg_start[v + 1];
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  bool dirty = false;
  auto flush = [&]() {
    if (!dirty) return;
    dirty = false;
    float* o = G + (size_t)v * 512 + dcol + 4 * threadIdx.x;
    atomicAdd(o, acc.x); atomicAdd(o + 1, acc.y); atomicAdd(o + 2, acc.z); atomicAdd(o + 3, acc.w);
    acc = make_float4(0.f, 0.f, 0.f, 0.f);
  };
  for (int p = p0; p < p1; ++p) {
    while (p >= vend) { flush(); ++v; vend = seg_start[v + 1]; }
    const float4 x = __ldg(reinterpret_cast<const float4*>(d1 + (size_t)order[p] * LIDF_H1) + threadIdx.x);
    acc.x += x.x; acc.y += x.y; acc.z += x.z; acc.w += x.w;
    dirty = true;
  }
  flush();
}

// ------------------------------------------------------------------------------------------------ small pieces
// PE(dir) per ray as the B operand of the dW1[:,dir] wgrad: [R][32] = [27 values | 1.0 | 0 x 4]; the ones column turns
// the same GEMM into the column sums of G_r (= db1 and the IEF constant's gradient).
__global__ void k_bwd_pedir(const float* __restrict__ dirs, int64_t R, int multires_views, int pos_encode, float* __restrict__ out) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  float x[32];
#pragma unroll
  for (int k = 0; k < 32; ++k) x[k] = 0.f;
  lidf_pe3(dirs[3 * r], dirs[3 * r + 1], dirs[3 * r + 2], multires_views, pos_encode, [&](int j, float v) { if (j < 27) x[j] = v; });
  x[27] = 1.0f;
#pragma unroll
  for (int i = 0; i < 8; ++i) *reinterpret_cast<float4*>(out + r * 32 + 4 * i) = make_float4(x[4 * i], x[4 * i + 1], x[4 * i + 2], x[4 * i + 3]);
}
// Wg[(row0 + n)][k] = w[n * ldw + col0 + k]: rows of a weight slice as a dense k-major [.,128] matrix
__global__ void k_bwd_pack_rows(const float* __restrict__ w, int ldw, int col0, int n_rows, float* __restrict__ dst, int row0) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_rows * 128) return;
  const int n = idx / 128, k = idx % 128;
  dst[(size_t)(row0 + n) * 128 + k] = w[(size_t)n * ldw + col0 + k];
}
// out[rows][128] = X[rows][512] Wg[512][128] (fp32 FFMA; d roi = G_r W1[:,rgb], d occ_voxel_feat = G_v W1[:,vox])
__global__ void __launch_bounds__(LIDF_SIMT_THREADS) k_bwd_rows_gemm(const float* __restrict__ X, int64_t rows, const float* __restrict__ Wg,
                                                                     float* __restrict__ out) {
  extern __shared__ __align__(16) float smem[];
  float* Xs = smem;                              // [64][512]
  float* Wtile = smem + LIDF_SIMT_BM * 512;      // [16][128]
  const int tid = threadIdx.x;
  const int64_t row0 = (int64_t)blockIdx.x * LIDF_SIMT_BM;
  for (int idx = tid; idx < LIDF_SIMT_BM * 128; idx += LIDF_SIMT_THREADS) {
    const int r = idx >> 7, c4 = idx & 127;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row0 + r < rows) v = __ldg(reinterpret_cast<const float4*>(X + (size_t)(row0 + r) * 512) + c4);
    *reinterpret_cast<float4*>(Xs + (size_t)r * 512 + c4 * 4) = v;
  }
  float acc[8][4];
  simt_gemm_tile<128>(Xs, 512, 512, Wg, 128, Wtile, acc);
  const int rg = tid >> 5, cl = tid & 31;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int64_t r = row0 + rg * 8 + i;
    if (r < rows)
#pragma unroll
      for (int j = 0; j < 4; ++j) out[(size_t)r * 128 + cl + 32 * j] = acc[i][j];
  }
}

// transpose of roi_channel_general: every sample's taps receive w * g / count (zero-weight taps skipped)
__global__ void __launch_bounds__(LIDF_ROI_THREADS)
k_roi_align_backward(const float* __restrict__ droi, int B, int H, int W, const int64_t* __restrict__ img_ind,
                     const int64_t* __restrict__ bid, int64_t R, int half, float* __restrict__ dfeat) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t ray = (int64_t)blockIdx.x * 32 + lane;
  if (ray >= R) return;
  const int px = (int)img_ind[2 * ray], py = (int)img_ind[2 * ray + 1];
  int b = (int)bid[ray];
  b = b < 0 ? 0 : (b >= B ? B - 1 : b);
  const RoiBox rb = roi_box(px, py, half, H, W);
  for (int c = warp; c < LIDF_RGB_CH; c += LIDF_ROI_THREADS / 32) {
    float* fc = dfeat + ((size_t)b * LIDF_RGB_CH + c) * H * W;
    const float4 g4 = __ldg(reinterpret_cast<const float4*>(droi + (size_t)ray * LIDF_RGB_DIM + 4 * c));
    const float gs[4] = {g4.x, g4.y, g4.z, g4.w};
#pragma unroll
    for (int ph = 0; ph < 2; ++ph)
#pragma unroll
      for (int pw = 0; pw < 2; ++pw) {
        const float gv = gs[ph * 2 + pw] / rb.count;
        for (int iy = 0; iy < rb.gh; ++iy) {
          const float y = rb.sh + ph * rb.bh + (iy + 0.5f) * rb.bh / (float)rb.gh;
          int ylo, yhi; float ly, hy; bool ydead;
          roi_tap(y, H, ylo, yhi, ly, hy, ydead);
          for (int ix = 0; ix < rb.gw; ++ix) {
            const float x = rb.sw + pw * rb.bw + (ix + 0.5f) * rb.bw / (float)rb.gw;
            int xlo, xhi; float lx, hx; bool xdead;
            roi_tap(x, W, xlo, xhi, lx, hx, xdead);
            if (ydead || xdead) continue;
            const float w1 = hy * hx, w2 = hy * lx, w3 = ly * hx, w4 = ly * lx;
            if (w1 != 0.f) atomicAdd(fc + (size_t)ylo * W + xlo, w1 * gv);
            if (w2 != 0.f) atomicAdd(fc + (size_t)ylo * W + xhi, w2 * gv);
            if (w3 != 0.f) atomicAdd(fc + (size_t)yhi * W + xlo, w3 * gv);
            if (w4 != 0.f) atomicAdd(fc + (size_t)yhi * W + xhi, w4 * gv);
          }
        }
      }
  }
}

// decoder-level leftovers: db1 = dc; IEF: dW1[:, D + j] = du w_enc[j] + dc b_enc[j], dw_enc[j] = sum_m W1[m][D + j] du[m],
// db_enc[j] = sum_m W1[m][D + j] dc[m]   (u = W1[:,D:] w_enc, c = W1[:,D:] b_enc; z1 = ... + u o_in + c)
__global__ void k_bwd_finish_decoder(const float* __restrict__ w1, int ldw, int D, const float* __restrict__ w_enc,
                                     const float* __restrict__ b_enc, int is_ief, const float* __restrict__ du,
                                     const float* __restrict__ dc, float* __restrict__ dw1, float* __restrict__ db1,
                                     float* __restrict__ dw_enc, float* __restrict__ db_enc) {
  const int t = threadIdx.x;                     // 256 threads
  db1[t] = dc[t];
  if (!is_ief) return;
  for (int j = 0; j < LIDF_IEF_ENC; ++j) dw1[(size_t)t * ldw + D + j] = du[t] * w_enc[j] + dc[t] * b_enc[j];
  __syncthreads();
  if (t < LIDF_IEF_ENC) {
    float s0 = 0.f, s1 = 0.f;
    for (int m = 0; m < LIDF_H1; ++m) { const float w = w1[(size_t)m * ldw + D + t]; s0 += w * du[m]; s1 += w * dc[m]; }
    dw_enc[t] = s0; db_enc[t] = s1;
  }
}
