// lidf_tc.cuh -- tcgen05 (5th-gen tensor core) engine for the LIDF decoders, sm_100a only.
//
// One persistent CTA per SM walks tiles of 128 ray-major pairs (row = TMEM lane = query point) and runs every decoder
// pass on the tile while all activations stay on chip:
//
//   warps 0-15 "row warps"  : 4 warps per TMEM lane quadrant, each owning 32 of every 128 columns.  They build the
//                             layer-1 A operand (voxel-feature gather + positional encoding) and run the epilogues
//                             TMEM -> regs (bias / per-ray term / leaky) -> bf16 hi|lo -> TMEM (in place)
//   warp 16    MMA issuer   : ONE elected thread issues every tcgen05.mma of the CTA from a fully static schedule
//                             (weight-ring slot, mbarrier parity and all descriptors are compile-time constants);
//                             completion via tcgen05.commit -> mbarrier
//   warp 17    weight loader: ONE elected thread streams the pre-packed bf16 weight chunks (UMMA canonical K-major
//                             layout) global/L2 -> smem ring with cp.async.bulk (TMA) + mbarrier complete_tx
//
// Precision: every fp32 operand x is split x = hi + lo (two bf16); a MAC is the 3 products hi*hi + lo*hi + hi*lo
// accumulated in fp32 (~2^-16 relative), which is what keeps the result within 1e-3 of the fp32 reference (a single
// bf16 or fp16 product is 3e-2 / 4e-3 off on the same inputs).
//
// Layer 1 is evaluated in its factored form (exact algebra):
//   W1 x = W1[:,pos] PE(enter, leave)   <- per pair, K = 102 (padded to 112): the only part that goes through the MMA
//        + A_v[vox]                     <- per voxel, fp32 row-prep GEMM, gathered in the layer-1 epilogue
//        + T[ray]                       <- per ray (ROI feature, PE(dir), bias, IEF constant), same
// The gather is coalesced: 8 lanes fetch one row's 128 B of (A_v + T), the warp transposes through a private smem
// staging tile, and every lane (= TMEM lane = pair) reads back its own 32 values.
// Layer-1 MMA operand (K = 112): sincos enter(48) | sincos leave(48) | xyz enter, xyz leave, 10 x 0, in shared memory
// in the UMMA canonical layout (SS-mode MMA).
// TMEM plan (512 columns x 128 lanes x 32 bit):
//   [  0, 64) Z  : layer-3 accumulator (64 fp32)          [64,128) unused
//   [128,256) X0 : layer-1 output half 0, fp32 accumulator -> converted IN PLACE to the layer-2 operand (K half 0)
//   [256,384) X1 : layer-1 output half 1, same
//   [384,512) Y  : layer-2 accumulator (128 fp32) -> converted in place to the layer-3 operand
// MMA issue order per pass p (the tensor pipe executes in order; Ex = epilogue of the row warps):
//   S2(p): L1 half 1 -> X1   | E0(p) converts X0 meanwhile
//   S1(p): L2 K-half 0 -> Y  | E1(p) converts X1 meanwhile
//   S3(p): L2 K-half 1 -> Y
//   S0(p+1): L1 half 0 of the NEXT pass (possibly of the next tile) -> X0   | E2(p) converts Y meanwhile
//   S4(p): L3 -> Z           | E3(p): layer-3 epilogue + layer-4 dot product
// The operand of the next tile is built by the row warps between E1 and E2 of a tile's last pass, as soon as the last
// layer-1 MMA of the tile has retired, so the tensor pipe does not drain at tile boundaries.
// Weight ring: 5 slots x 16 KB.  A pass consumes 20 fills (4 + 4 + 4 + 4 + 4) = exactly 4 ring rotations, so the slot
// and parity of every fill are the same in every pass.
#pragma once
#include <cuda_bf16.h>

#include "lidf_common.cuh"
#include "lidf_prep.cuh"

#define TC_ROW_WARPS 16
#define TC_ROW_THREADS (TC_ROW_WARPS * 32)
#define TC_THREADS ((TC_ROW_WARPS + 2) * 32)
#define TC_CHUNK_BYTES 8192
#define TC_CHUNKS_PER_DEC 34     // 7 (L1 half 0) + 7 (L1 half 1) + 8 (L2 K-half 0) + 8 (L2 K-half 1) + 4 (L3)
#define TC_STAGES 5
#define TC_STAGE_BYTES 16384
#define TC_FILLS_PER_PASS 20
#define TC_K1_STEPS 7             // k-steps of the layer-1 MMA (K = 112)
#define TC_COL_Z 0                // layer-3 accumulator
#define TC_COL_X0 128             // layer-1 output half 0: accumulator -> (in place) layer-2 operand
#define TC_COL_X1 256             // layer-1 output half 1
#define TC_COL_Y 384              // layer-2 accumulator -> (in place) layer-3 operand
#define TC_A1S_PART_BYTES (TC_K1_STEPS * 4096)    // one part (hi or lo): [kstep][kgroup(2)][128 rows][16 B]
#define TC_STAGE_PITCH 36         // floats per row of the per-warp gather staging tile (32 + 4: conflict-free LDS.128)
#define TC_KPE_MAX 112            // widest per-pair PE block (2 x PE(pos)) the layer-1 operand layout holds
#define TC_MAX_PASSES 9
#define TC_SPIN_LIMIT (1u << 22)
// Timing experiments only (results are garbage when non-zero; tools/dbg_modes.sh, profiles/r1l_modes.md): bit 0 = do not
// issue the MMAs, bit 1 = epilogues skip the TMEM loads, bit 2 = epilogues skip math + TMEM stores, bit 3 = operand build
// skips sincos.  Never set in a product build.
#ifndef TC_DEBUG_MODE
#define TC_DEBUG_MODE 0
#endif
// Energy experiments (results are garbage when non-zero; profiles/r2i_energy.md): bit 0 = the weight ring is filled from L2
// only during the CTA's first pass (later fills just complete the barrier), bit 1 = layer 1 issued as N = 256 MMAs (half the
// instructions), bit 2 = no layer-1 MMAs for IEF iterations > 0, bit 3 = only the hi*hi product in layer 3, bit 4 = every row reads T[0]
// (L1-resident: no DRAM latency on the per-ray term), bit 5 = every row gathers A_v[0].
#ifndef TC_EXP
#define TC_EXP 0
#endif
// suspend-time hint (ns) of mbarrier.try_wait: the warp is parked by the hardware instead of spinning through the loop
#ifndef TC_WAIT_HINT_NS
#define TC_WAIT_HINT_NS 1000
#endif

// ------------------------------------------------------------------------------------------------ PTX wrappers
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"((uint32_t)TC_WAIT_HINT_NS)
      : "memory");
  return ok != 0;
}
// bounded spin: a protocol bug traps (-> CUDA error) instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > TC_SPIN_LIMIT) __trap();
  }
}
// TMA bulk copy global -> shared, completion counted on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// true on exactly one (converged) lane of the warp
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// D[tmem] (+)= A[tmem] * B[smem]^T, kind::f16 (bf16 in, fp32 accumulate), M = 128.  (SASS: UTCHMMA)
__device__ __forceinline__ void mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  if (TC_DEBUG_MODE & 1) return;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n\t}"
      :
      : "r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(0u)
      : "memory");
}

// D[tmem] (+)= A[smem] * B[smem]^T
__device__ __forceinline__ void mma_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  if (TC_DEBUG_MODE & 1) return;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}"
      :
      : "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(0u)
      : "memory");
}
// make generic-proxy smem writes visible to the async proxy (tensor core operand fetch)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void st_shared_v4(uint32_t saddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// instruction descriptor (cute::UMMA::InstrDescriptor): c=F32 [4,6), a=BF16 [7,10), b=BF16 [10,13), K-major A/B,
// N>>3 at [17,23), M>>4 at [24,29)
__host__ __device__ constexpr uint32_t make_idesc(int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor), K-major, SWIZZLE_NONE ("interleave"):
// core matrix = 8 rows x 16 B stored as 128 contiguous bytes; SBO = byte stride between 8-row groups,
// LBO = byte stride between the two core matrices along K.  Layout used here: [kgroup(2)][N rows][16 B].
__device__ __forceinline__ uint64_t make_bdesc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46);
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  if (TC_DEBUG_MODE & 2) {
#pragma unroll
    for (int i = 0; i < 32; ++i) r[i] = taddr + i;
    return;
  }
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* r) {
  if (TC_DEBUG_MODE & 4) return;
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}

// x = hi + lo with hi = bf16_rn(x), lo = bf16_rn(x - hi).  Packs two elements per word, EVEN element in the low half
// (element order inside a 32-bit TMEM column / smem word is little-endian in K).
__device__ __forceinline__ void split2(float x_even, float x_odd, uint32_t& hi, uint32_t& lo) {
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(x_odd), "f"(x_even));
  const float re = x_even - __uint_as_float(hi << 16);
  const float ro = x_odd - __uint_as_float(hi & 0xFFFF0000u);
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(ro), "f"(re));
}
// one k-step (16 elements) -> 8 hi words | 8 lo words
__device__ __forceinline__ void split16(const float* x, uint32_t* out16) {
#pragma unroll
  for (int j = 0; j < 8; ++j) split2(x[2 * j], x[2 * j + 1], out16[j], out16[8 + j]);
}

}  // namespace tc

// ------------------------------------------------------------------------------------------------ packing kernels
// element e of the layer-1 MMA operand (K index) -> column of linear_1.weight; -1 = zero padding.
// single_pos (RefineNet decoder tail, pipeline.py:1018-1023): only one encoded position sits at column 256; the
// "leave" slots of the operand get zero weights.
__host__ __device__ inline int tc_a1_col(int e, int pe_pos, int single_pos = 0) {
  const int base = LIDF_VOX_DIM + LIDF_RGB_DIM;             // PE(enter) starts at column 256 (pipeline.py:431-433)
  if (e < 48) return base + 3 + e;                          // sin/cos part of PE(enter)
  if (e < 96) return single_pos ? -1 : base + pe_pos + 3 + (e - 48);   // sin/cos part of PE(leave)
  if (e < 99) return base + (e - 96);                       // raw enter xyz
  if (e < 102) return single_pos ? -1 : base + pe_pos + (e - 99);      // raw leave xyz
  return -1;
}

// weight stream of one decoder: 34 chunks x 8 KB.  Chunk with N rows (128, or 64 for layer 3) holds k-steps of
// [hi: kg0 N x 16 B | kg1 N x 16 B][lo: kg0 | kg1]; an N=128 chunk is one k-step, an N=64 chunk two.
__global__ void k_pack_tc_weights(const float* __restrict__ w1, int ldw1, int pe_pos, const float* __restrict__ w2,
                                  const float* __restrict__ w3, uint8_t* __restrict__ stream, int single_pos) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= TC_CHUNKS_PER_DEC * 2048) return;
  const int c = idx / 2048, r = idx % 2048;
  float w = 0.f;
  int n, kk, N, ksub = 0;
  if (c < 30) {
    N = 128; n = r / 16; kk = r % 16;
    if (c < 14) {                                           // layer 1, output half 0 / 1
      const int half = c >= 7, s = half ? c - 7 : c;
      const int col = tc_a1_col(16 * s + kk, pe_pos, single_pos);
      if (col >= 0) w = w1[(size_t)(half * 128 + n) * ldw1 + col];
    } else {                                                // layer 2, K half 0 / 1
      const int half = c >= 22, s = half ? c - 22 : c - 14;
      w = w2[(size_t)n * LIDF_H1 + half * 128 + 16 * s + kk];
    }
  } else {                                                  // layer 3: N = 64, two k-steps per chunk
    N = 64; ksub = r / 1024; const int rr = r % 1024; n = rr / 16; kk = rr % 16;
    w = w3[(size_t)n * LIDF_H2 + 16 * (2 * (c - 30) + ksub) + kk];
  }
  const __nv_bfloat16 hi = __float2bfloat16_rn(w);
  const __nv_bfloat16 lo = __float2bfloat16_rn(w - __bfloat162float(hi));
  const int kg = kk >> 3, e = kk & 7;
  const size_t kstep_bytes = (size_t)N * 64;                // hi (N*32) + lo (N*32)
  uint8_t* base = stream + (size_t)c * TC_CHUNK_BYTES + ksub * kstep_bytes;
  const size_t off = (size_t)kg * N * 16 + (size_t)n * 16 + e * 2;
  *reinterpret_cast<__nv_bfloat16*>(base + off) = hi;
  *reinterpret_cast<__nv_bfloat16*>(base + (size_t)N * 32 + off) = lo;
}

// ------------------------------------------------------------------------------------------------ self test
// D[128][128] = A[128][32] * W[128][32]^T through the exact primitives the engine uses (tcgen05.st A operand, TMA'd
// weight chunks, TS-mode MMA with 3 split products, tcgen05.ld).  variant bit0 swaps LBO/SBO (must then be wrong).
__global__ void __launch_bounds__(128) k_tc_selftest(const uint8_t* __restrict__ chunks, const float* __restrict__ A,
                                                     float* __restrict__ D, int variant) {
  __shared__ __align__(1024) uint8_t s_w[2 * TC_CHUNK_BYTES];
  __shared__ __align__(8) uint64_t s_full, s_done;
  __shared__ uint32_t s_tmem;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) { tc::mbar_init(&s_full, 1); tc::mbar_init(&s_done, 1); tc::fence_barrier_init(); }
  if (warp == 0) tc::tmem_alloc(&s_tmem, 256);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = s_tmem;
  if (tid == 0) {
    tc::mbar_arrive_expect_tx(&s_full, 2 * TC_CHUNK_BYTES);
    tc::bulk_g2s(s_w, chunks, TC_CHUNK_BYTES, &s_full);
    tc::bulk_g2s(s_w + TC_CHUNK_BYTES, chunks + TC_CHUNK_BYTES, TC_CHUNK_BYTES, &s_full);
  }
  {
    float x[32];
#pragma unroll
    for (int k = 0; k < 32; ++k) x[k] = A[tid * 32 + k];
    uint32_t w[16];
    const uint32_t row_addr = tmem + ((uint32_t)(warp * 32) << 16);
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      tc::split16(x + 16 * s, w);
      tc::tmem_st16(row_addr + 16 * s, w);
    }
    tc::wait_st();
  }
  tc::fence_before_sync();
  __syncthreads();
  if (tid == 0) {
    tc::fence_after_sync();
    tc::mbar_wait(&s_full, 0);
    const uint32_t idesc = tc::make_idesc(128);
    const uint32_t lbo = (variant & 1) ? 128u : 128u * 16u, sbo = (variant & 1) ? 128u * 16u : 128u;
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      const uint32_t whi = tc::smem_u32(s_w + s * TC_CHUNK_BYTES), wlo = whi + 128 * 32;
      const uint32_t a_hi = tmem + 16 * s, a_lo = a_hi + 8;
      tc::mma_ts(tmem + 128, a_hi, tc::make_bdesc(whi, lbo, sbo), idesc, s > 0);
      tc::mma_ts(tmem + 128, a_lo, tc::make_bdesc(whi, lbo, sbo), idesc, 1);
      tc::mma_ts(tmem + 128, a_hi, tc::make_bdesc(wlo, lbo, sbo), idesc, 1);
    }
    tc::commit(&s_done);
  }
  tc::mbar_wait(&s_done, 0);
  tc::fence_after_sync();
  {
    const uint32_t row_addr = tmem + ((uint32_t)(warp * 32) << 16) + 128;
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
      uint32_t r[32];
      tc::tmem_ld32(row_addr + 32 * c, r);
      tc::wait_ld();
#pragma unroll
      for (int j = 0; j < 32; ++j) D[tid * 128 + 32 * c + j] = __uint_as_float(r[j]);
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 256);
  (void)lane;
}

// pack W[128][32] (fp32, row-major) into two N=128 chunks for the self test
__global__ void k_tc_selftest_pack(const float* __restrict__ W, uint8_t* __restrict__ chunks) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= 2 * 2048) return;
  const int c = idx / 2048, r = idx % 2048, n = r / 16, kk = r % 16;
  const float w = W[n * 32 + 16 * c + kk];
  const __nv_bfloat16 hi = __float2bfloat16_rn(w);
  const __nv_bfloat16 lo = __float2bfloat16_rn(w - __bfloat162float(hi));
  const size_t off = (size_t)(kk >> 3) * 128 * 16 + (size_t)n * 16 + (kk & 7) * 2;
  *reinterpret_cast<__nv_bfloat16*>(chunks + (size_t)c * TC_CHUNK_BYTES + off) = hi;
  *reinterpret_cast<__nv_bfloat16*>(chunks + (size_t)c * TC_CHUNK_BYTES + 128 * 32 + off) = lo;
}

// ------------------------------------------------------------------------------------------------ the engine
namespace tc {
__device__ __forceinline__ void mbar_wait_a(uint32_t bar_saddr, uint32_t parity) {
  uint32_t spins = 0;
  for (;;) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar_saddr), "r"(parity), "r"((uint32_t)TC_WAIT_HINT_NS)
        : "memory");
    if (ok) break;
    if (++spins > TC_SPIN_LIMIT) __trap();
  }
}
__device__ __forceinline__ void commit_a(uint32_t bar_saddr) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_saddr) : "memory");
}
__device__ __forceinline__ void mbar_arrive_a(uint32_t bar_saddr) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_saddr) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  if (TC_DEBUG_MODE & 2) {
#pragma unroll
    for (int i = 0; i < 16; ++i) r[i] = taddr + i;
    return;
  }
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
// packed fp32 pairs (Blackwell FADD2 / FMUL2 / FFMA2): same IEEE results as the scalar ops, half the instructions
__device__ __forceinline__ uint64_t pack2(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
// (v0, v1) -> x = leaky(v) -> bf16 hi pair | lo pair  (hi = bf16_rn(x), lo = bf16_rn(x - hi); element 0 in the low half)
__device__ __forceinline__ void leaky_split2(uint64_t v, uint32_t& hi, uint32_t& lo) {
  const uint64_t m = mul2(v, pack2(LIDF_LEAKY, LIDF_LEAKY));
  float v0, v1, m0, m1;
  unpack2(v, v0, v1);
  unpack2(m, m0, m1);
  const float x0 = fmaxf(v0, m0), x1 = fmaxf(v1, m1);
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(x1), "f"(x0));
  const uint64_t h = pack2(__uint_as_float(hi << 16), __uint_as_float(hi & 0xFFFF0000u));
  float r0, r1;
  unpack2(fma2(h, pack2(-1.0f, -1.0f), pack2(x0, x1)), r0, r1);
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(r1), "f"(r0));
}
// 16-byte asynchronous copy global -> shared (LDGSTS), no register staging
__device__ __forceinline__ void cp_async16(uint32_t dst_saddr, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst_saddr), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
// named barrier over the 4 row warps of one TMEM lane quadrant (ids 1..4)
__device__ __forceinline__ void bar_quadrant(int q) { asm volatile("bar.sync %0, 128;" ::"r"(q + 1) : "memory"); }
}  // namespace tc

struct TcArgs {
  int64_t P; int n_tiles;
  const int* perm; const int64_t* pair_vox; const int64_t* pair_ray;
  const float* pair_dist; const float* dense_dist; int64_t R; int64_t V;
  const float* ray_dir; const float* voxel_bound; int rel;
  const float* pos_in;             // RefineNet tail: row = ray, position given ([P,3]); pair_ray / dist are not read
  const float* Av;                 // [V][512] per-voxel layer-1 term, decoder d at column 256 d
  const float* T;                  // [R][512] per-ray layer-1 term, decoder d at column 256 d
  const uint8_t* wstream;          // [2][34][8192]
  const float* u;                  // [256] IEF rank-1 vector of decoder 0 (NULL if IMNet)
  const float* b2[2]; const float* b3[2]; const float* w4[2]; const float* b4[2];
  int kind[2]; int n_pass[2]; int use_sigmoid[2];
  int npt;                         // passes per tile
  uint32_t pass_dec_mask;          // bit p = decoder of pass p: the two decoders are interleaved so that consecutive passes
  uint64_t pass_it_pack;           // are independent wherever possible (IEF iteration k+1 needs iteration k); 4 bits per
                                   // pass = its iteration index.  Packed so that the kernel indexes them with shifts.
  float o0, r0, r1, sqrt3, part;
  float* out[2];                   // pred_offset, pred_prob_end  (written at the original pair index)
  float* pos_out;                  // pair_pred_pos [P,3]
  int n_prod;                      // bf16 products per MAC: 3 (hi*hi + lo*hi + hi*lo) or 1
  float* o_iter;                   // [n_pass[0]-1][P] or NULL: decoder 0's running IEF offset after every iteration but the last
  int out_by_slot;                 // 1: outputs are written at the row's slot (tile * 128 + row) instead of its original pair
                                   // index, and perm[slot] < 0 marks a slot without a row (winner-only pass: slot = ray)
};

struct TcSmem {
  uint8_t w[TC_STAGES][TC_STAGE_BYTES];
  uint8_t a1[2][TC_A1S_PART_BYTES];  // layer-1 MMA operand: [hi|lo][kstep(7)][kgroup(2)][128 rows][16 B]
  float stage[TC_ROW_WARPS][32 * TC_STAGE_PITCH];   // per-warp transpose tile of the gathered (A_v + T) rows
  float u[LIDF_H1];
  float b2[2][LIDF_H2];
  float b3[2][LIDF_H3];
  float w4[2][LIDF_H3];
  float b4[2];
  float part[2][4][128];             // [parity][column group][row] layer-4 partial sums
  int m_orig[2][128];                // [tile parity][row] original pair index
  float m_geo[2][6][128];            // [tile parity][enter xyz, dir xyz][row]
  uint64_t w_full[TC_STAGES], w_empty[TC_STAGES];
  uint64_t a1_ready, a1_free, x_full[2], x_done[2], y_full, y_done, z_full, z_free;
  uint32_t tmem_base;
};

// everything the MMA issuer thread needs; descriptors are (base + compile-time constant)
struct TcIssue {
  uint32_t tmem;
  uint32_t full0, empty0;            // smem addresses of w_full[0], w_empty[0]
  uint64_t wd128, wd64;              // B descriptors of ring slot 0 for N = 128 / N = 64 chunks
  uint64_t ad;                       // A descriptor of k-step 0 of the hi part of the smem operand tile
};

// one k-step (K = 16), A operand in TMEM (8 cols hi | 8 cols lo), N = 128: B = [hi 4 KB][lo 4 KB] at byte offset off
template <int NPROD>
__device__ __forceinline__ void tc_k_ts128(const TcIssue& c, uint32_t dcol, uint32_t acol, uint32_t off, bool first) {
  constexpr uint32_t idesc = tc::make_idesc(128);
  const uint64_t bhi = c.wd128 + (off >> 4);
  tc::mma_ts(c.tmem + dcol, c.tmem + acol, bhi, idesc, first ? 0u : 1u);
  if (NPROD == 3) {
    tc::mma_ts(c.tmem + dcol, c.tmem + acol + 8, bhi, idesc, 1u);
    tc::mma_ts(c.tmem + dcol, c.tmem + acol, bhi + (4096u >> 4), idesc, 1u);
  }
}
// layer 3: A in TMEM, N = 64: B = [hi 2 KB][lo 2 KB] at byte offset off
template <int NPROD>
__device__ __forceinline__ void tc_k_ts64(const TcIssue& c, uint32_t dcol, uint32_t acol, uint32_t off, bool first) {
  constexpr uint32_t idesc = tc::make_idesc(64);
  const uint64_t bhi = c.wd64 + (off >> 4);
  tc::mma_ts(c.tmem + dcol, c.tmem + acol, bhi, idesc, first ? 0u : 1u);
  if (NPROD == 3 && !(TC_EXP & 8)) {
    tc::mma_ts(c.tmem + dcol, c.tmem + acol + 8, bhi, idesc, 1u);
    tc::mma_ts(c.tmem + dcol, c.tmem + acol, bhi + (2048u >> 4), idesc, 1u);
  }
}

// Fill GI (index in the static schedule) lives in ring slot GI % 7 and is the (GI / 7)-th use of that slot.
#define TC_SLOT(gi) ((gi) % TC_STAGES)
#define TC_PAR(gi) ((uint32_t)(((gi) / TC_STAGES) & 1))

// layer 1, one output half (7 k-steps = 3 fills of 2 + 1 fill of 1), A operand in shared memory
template <int NPROD, int GI0>
__device__ __forceinline__ void tc_issue_l1(const TcIssue& c, uint32_t dcol, bool exp_skip = false, bool exp_n256 = false) {
  const uint32_t idesc = (TC_EXP & 2) && exp_n256 ? tc::make_idesc(256) : tc::make_idesc(128);
#pragma unroll
  for (int sg = 0; sg < 4; ++sg) {
    const int gi = GI0 + sg, slot = TC_SLOT(gi);
    tc::mbar_wait_a(c.full0 + 8 * slot, TC_PAR(gi));
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int ks = 2 * sg + j;
      if (ks < TC_K1_STEPS && !((TC_EXP & 6) && exp_skip)) {
        const uint64_t bhi = c.wd128 + ((uint32_t)(slot * TC_STAGE_BYTES + j * TC_CHUNK_BYTES) >> 4);
        const uint64_t ahi = c.ad + ((uint32_t)(ks * 4096) >> 4);
        tc::mma_ss(c.tmem + dcol, ahi, bhi, idesc, ks == 0 ? 0u : 1u);
        if (NPROD == 3) {
          tc::mma_ss(c.tmem + dcol, ahi + ((uint32_t)TC_A1S_PART_BYTES >> 4), bhi, idesc, 1u);
          tc::mma_ss(c.tmem + dcol, ahi, bhi + (4096u >> 4), idesc, 1u);
        }
      }
    }
    tc::commit_a(c.empty0 + 8 * slot);
  }
}
// layer 2, one K half (8 k-steps = 4 fills), operand = converted layer-1 half at acol, accumulator Y
template <int NPROD, int GI0>
__device__ __forceinline__ void tc_issue_l2(const TcIssue& c, uint32_t acol, bool first_half) {
#pragma unroll
  for (int sg = 0; sg < 4; ++sg) {
    const int gi = GI0 + sg, slot = TC_SLOT(gi);
    tc::mbar_wait_a(c.full0 + 8 * slot, TC_PAR(gi));
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int ks = 2 * sg + j;
      tc_k_ts128<NPROD>(c, TC_COL_Y, acol + 16 * ks, (uint32_t)(slot * TC_STAGE_BYTES + j * TC_CHUNK_BYTES),
                        first_half && ks == 0);
    }
    tc::commit_a(c.empty0 + 8 * slot);
  }
}
// layer 3 (K = 128, N = 64: 4 fills of 8 KB = 2 k-steps each), operand = converted Y, accumulator Z
template <int NPROD, int GI0>
__device__ __forceinline__ void tc_issue_l3(const TcIssue& c) {
#pragma unroll
  for (int sg = 0; sg < 4; ++sg) {
    const int gi = GI0 + sg, slot = TC_SLOT(gi);
    tc::mbar_wait_a(c.full0 + 8 * slot, TC_PAR(gi));
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int ks = 2 * sg + j;
      tc_k_ts64<NPROD>(c, TC_COL_Z, TC_COL_Y + 16 * ks, (uint32_t)(slot * TC_STAGE_BYTES + j * 4096), ks == 0);
    }
    tc::commit_a(c.empty0 + 8 * slot);
  }
}
// consume fills without issuing MMAs (keeps the static schedule aligned on the very last pass of the CTA)
template <int GI0, int N>
__device__ __forceinline__ void tc_drain(const TcIssue& c) {
#pragma unroll
  for (int sg = 0; sg < N; ++sg) {
    const int gi = GI0 + sg, slot = TC_SLOT(gi);
    tc::mbar_wait_a(c.full0 + 8 * slot, TC_PAR(gi));
    tc::mbar_arrive_a(c.empty0 + 8 * slot);
  }
}

// loader side of the same static schedule: N fills starting at GI0; `pairs` fills are 16 KB, the rest 8 KB
template <int GI0, int N, int PAIRS>
__device__ __forceinline__ void tc_load_seg(TcSmem& S, const uint8_t* src, bool ring_primed, bool no_stream = false) {
#pragma unroll
  for (int sg = 0; sg < N; ++sg) {
    const int gi = GI0 + sg, slot = TC_SLOT(gi);
    const uint32_t bytes = sg < PAIRS ? TC_STAGE_BYTES : TC_CHUNK_BYTES;
    // previous use of this slot was fill gi - 7: wait for its release (skip while the ring fills for the first time)
    if (ring_primed || gi >= TC_STAGES) tc::mbar_wait(&S.w_empty[slot], TC_PAR(gi) ^ 1u);
    if ((TC_EXP & 1) && ring_primed && no_stream) { tc::mbar_arrive(&S.w_full[slot]); src += bytes; continue; }
    tc::mbar_arrive_expect_tx(&S.w_full[slot], bytes);
    tc::bulk_g2s(S.w[slot], src, bytes, &S.w_full[slot]);
    src += bytes;
  }
}

template <int NPROD>
__global__ void __launch_bounds__(TC_THREADS, 1) k_mlp_tc(const __grid_constant__ TcArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  TcSmem& S = *reinterpret_cast<TcSmem*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  // ---- one-time setup -------------------------------------------------------------------------------------
  for (int i = tid; i < LIDF_H1; i += TC_THREADS) S.u[i] = a.u ? a.u[i] : 0.f;
  for (int i = tid; i < 2 * LIDF_H2; i += TC_THREADS) S.b2[i / LIDF_H2][i % LIDF_H2] = a.b2[i / LIDF_H2][i % LIDF_H2];
  for (int i = tid; i < 2 * LIDF_H3; i += TC_THREADS) {
    S.b3[i / LIDF_H3][i % LIDF_H3] = a.b3[i / LIDF_H3][i % LIDF_H3];
    S.w4[i / LIDF_H3][i % LIDF_H3] = a.w4[i / LIDF_H3][i % LIDF_H3];
  }
  if (tid < 2) S.b4[tid] = a.b4[tid][0];
  if (tid == 0) {
    for (int i = 0; i < TC_STAGES; ++i) { tc::mbar_init(&S.w_full[i], 1); tc::mbar_init(&S.w_empty[i], 1); }
    tc::mbar_init(&S.a1_ready, TC_ROW_WARPS);
    tc::mbar_init(&S.a1_free, 1);
    for (int i = 0; i < 2; ++i) { tc::mbar_init(&S.x_full[i], 1); tc::mbar_init(&S.x_done[i], TC_ROW_WARPS); }
    tc::mbar_init(&S.y_full, 1);
    tc::mbar_init(&S.y_done, TC_ROW_WARPS);
    tc::mbar_init(&S.z_full, 1);
    tc::mbar_init(&S.z_free, TC_ROW_WARPS);
    tc::fence_barrier_init();
  }
  if (warp == TC_ROW_WARPS) tc::tmem_alloc(&S.tmem_base, 512);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = S.tmem_base;
  const int npt = a.npt;                                                       // passes per tile
  const uint32_t dec_mask = a.pass_dec_mask;
  const uint64_t it_pack = a.pass_it_pack;
  auto pass_dec = [&](int p) { return (int)((dec_mask >> p) & 1u); };
  auto pass_it = [&](int p) { return (int)((it_pack >> (4 * p)) & 15u); };
  const bool ief0 = a.kind[0] == LIDF_DEC_IEF, ief1 = a.kind[1] == LIDF_DEC_IEF;
  const int npass0 = a.n_pass[0], npass1 = a.n_pass[1], sig0 = a.use_sigmoid[0], sig1 = a.use_sigmoid[1];
  const int n_my_tiles = (a.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int total_passes = n_my_tiles * npt;
  // weight-stream segments of one decoder (chunk offsets)
  constexpr int SEG_L1H0 = 0, SEG_L1H1 = 7, SEG_L2K0 = 14, SEG_L2K1 = 22, SEG_L3 = 30;
  // static fill schedule: prologue S0 = fills 0..3; a pass = fills 4..23 (+ 20 per pass: same slots, same parities)
  constexpr int GI_S2 = 4, GI_S1 = 8, GI_S3 = 12, GI_S0 = 16, GI_S4 = 20;

  if (warp == TC_ROW_WARPS + 1) {
    // ================================ weight loader (TMA), one thread ================================
    if (tc::elect_one()) {
      auto seg = [&](int d, int chunk) { return a.wstream + ((size_t)d * TC_CHUNKS_PER_DEC + chunk) * TC_CHUNK_BYTES; };
      tc_load_seg<0, 4, 3>(S, seg(pass_dec(0), SEG_L1H0), false);
      int p = 0;
      for (int gp = 0; gp < total_passes; ++gp) {
        const int d = pass_dec(p);
        const int pn = p + 1 == npt ? 0 : p + 1;
        const int dn = pass_dec(pn);
        tc_load_seg<GI_S2, 4, 3>(S, seg(d, SEG_L1H1), gp > 0, gp > 0);
        tc_load_seg<GI_S1, 4, 4>(S, seg(d, SEG_L2K0), true, gp > 0);
        tc_load_seg<GI_S3, 4, 4>(S, seg(d, SEG_L2K1), true, gp > 0);
        tc_load_seg<GI_S0, 4, 3>(S, seg(dn, SEG_L1H0), true, gp > 0);          // drained unused after the last pass
        tc_load_seg<GI_S4, 4, 0>(S, seg(d, SEG_L3), true, gp > 0);
        p = pn;
      }
    }
  } else if (warp == TC_ROW_WARPS) {
    // ================================ MMA issuer, one thread ================================
    if (tc::elect_one()) {
      TcIssue c;
      c.tmem = tmem;
      c.full0 = tc::smem_u32(&S.w_full[0]);
      c.empty0 = tc::smem_u32(&S.w_empty[0]);
      c.wd128 = tc::make_bdesc(tc::smem_u32(S.w[0]), 2048u, 128u);
      c.wd64 = tc::make_bdesc(tc::smem_u32(S.w[0]), 1024u, 128u);
      c.ad = tc::make_bdesc(tc::smem_u32(S.a1[0]), 2048u, 128u);
      tc::mbar_wait(&S.a1_ready, 0);
      tc::fence_after_sync();
      tc_issue_l1<NPROD, 0>(c, TC_COL_X0, false, true);                        // S0 of the first pass
      tc::commit(&S.x_full[0]);
      int p = 0;
      uint32_t tl = 0;                                                         // tiles whose operand has been consumed
      for (int gp = 0; gp < total_passes; ++gp) {
        const uint32_t ph = (uint32_t)gp & 1u;
        const bool last_of_tile = p + 1 == npt;
        // S2(p): layer 1, output half 1 -> X1 (its previous reader S3(p-1) precedes it in the in-order pipe)
        tc_issue_l1<NPROD, GI_S2>(c, TC_COL_X1, (TC_EXP & 2) || ((TC_EXP & 4) && pass_it(p) > 0));
        tc::commit(&S.x_full[1]);
        if (last_of_tile) tc::commit(&S.a1_free);                             // last reader of this tile's layer-1 operand
        // S1(p): layer 2, K half 0 (X0 converted in place by E0) -> Y
        tc::mbar_wait(&S.x_done[0], ph);
        tc::fence_after_sync();
        tc_issue_l2<NPROD, GI_S1>(c, TC_COL_X0, true);
        // S3(p): layer 2, K half 1 (X1 converted by E1) -> Y
        tc::mbar_wait(&S.x_done[1], ph);
        tc::fence_after_sync();
        tc_issue_l2<NPROD, GI_S3>(c, TC_COL_X1, false);
        tc::commit(&S.y_full);
        // S0(p+1): layer 1, output half 0 of the next pass -> X0 (free: S1(p) precedes it in the in-order pipe)
        if (gp + 1 < total_passes) {
          if (last_of_tile) {
            ++tl;
            tc::mbar_wait(&S.a1_ready, tl & 1u);                               // operand of the next tile
            tc::fence_after_sync();
          }
          tc_issue_l1<NPROD, GI_S0>(c, TC_COL_X0, (TC_EXP & 4) && pass_it(last_of_tile ? 0 : p + 1) > 0, true);
          tc::commit(&S.x_full[0]);
        } else {
          tc_drain<GI_S0, 4>(c);
        }
        // S4(p): layer 3, operand = Y converted by E2, accumulator Z (E3 of the previous pass must have read Z)
        tc::mbar_wait(&S.y_done, ph);
        if (gp > 0) tc::mbar_wait(&S.z_free, ph ^ 1u);
        tc::fence_after_sync();
        tc_issue_l3<NPROD, GI_S4>(c);
        tc::commit(&S.z_full);
        p = last_of_tile ? 0 : p + 1;
      }
    }
  } else if (warp < TC_ROW_WARPS) {
    // ================================ row warps: operand build + epilogues ================================
    const int q = warp & 3, g = warp >> 2;            // TMEM lane quadrant, column group
    const int row = q * 32 + lane;
    const uint32_t lane_addr = tmem + ((uint32_t)(q * 32) << 16);
    const uint32_t a1row_hi = tc::smem_u32(S.a1[0]) + row * 16, a1row_lo = tc::smem_u32(S.a1[1]) + row * 16;
    // store one k-step (8 hi words | 8 lo words) of the smem operand tile: hi/lo x kgroup 0/1, 16 B each
    auto st_a1 = [&](int sk, const uint32_t* w) {
      tc::st_shared_v4(a1row_hi + sk * 4096, w[0], w[1], w[2], w[3]);
      tc::st_shared_v4(a1row_hi + sk * 4096 + 2048, w[4], w[5], w[6], w[7]);
      tc::st_shared_v4(a1row_lo + sk * 4096, w[8], w[9], w[10], w[11]);
      tc::st_shared_v4(a1row_lo + sk * 4096 + 2048, w[12], w[13], w[14], w[15]);
    };
    struct RowMeta { int orig, vox, ray; float t0, t1; bool valid; };
    // two-stage load (the second stage depends on the first): issued a pass apart so neither latency is exposed
    auto load_orig = [&](int tile) {
      const int64_t s = (int64_t)tile * 128 + row;
      return s < a.P ? (a.perm ? a.perm[s] : (int)s) : -1;
    };
    auto load_meta = [&](int tile, int orig_) {
      RowMeta m{0, 0, 0, 0.f, 0.f, false};
      const int64_t s = (int64_t)tile * 128 + row;
      if (s < a.P && orig_ >= 0) {
        m.valid = true;
        m.orig = orig_;
        // out-of-range indices are clamped (and reported by k_count_pairs / k_validate_indices), never dereferenced
        m.vox = (int)lidf_clamp_idx(a.pair_vox[m.orig], a.V);
        m.ray = a.pos_in ? m.orig : (int)lidf_clamp_idx(a.pair_ray[m.orig], a.R);
        if (a.pos_in) { }
        else if (a.pair_dist) { const float2 t = *reinterpret_cast<const float2*>(a.pair_dist + 2 * (size_t)m.orig); m.t0 = t.x; m.t1 = t.y; }
        else { const size_t o = ((size_t)m.vox * a.R + m.ray) * 2; m.t0 = a.dense_dist[o]; m.t1 = a.dense_dist[o + 1]; }
      }
      return m;
    };
    // layer-1 MMA operand of one tile.  Column group 0 encodes the enter position (k-steps 0-2), 1 the leave position
    // (k-steps 3-5), 2 writes the raw xyz k-step (6) and parks the output metadata in smem.
    auto build_a1 = [&](const RowMeta& m, int buf, int tile_) {
      float dir[3] = {0.f, 0.f, 0.f}, enter[3] = {0.f, 0.f, 0.f}, pe[3] = {0.f, 0.f, 0.f}, pl[3] = {0.f, 0.f, 0.f};
      if (m.valid) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          dir[k] = a.ray_dir[(size_t)m.ray * 3 + k];
          const float c = a.rel ? (a.voxel_bound[(size_t)m.vox * 6 + k] + a.voxel_bound[(size_t)m.vox * 6 + 3 + k]) / 2.0f : 0.f;
          enter[k] = a.pos_in ? a.pos_in[(size_t)m.orig * 3 + k] : dir[k] * m.t0;
          pe[k] = enter[k] - c;
          pl[k] = a.pos_in ? 0.f : dir[k] * m.t1 - c;
        }
      }
      if (g < 2) {
        // positional encoding: accurate sincos at f = 1 and f = 16, three exact-form double-angle steps after each
        // (sin 2x = 2 s c, cos 2x = 1 - 2 s^2)
        float v[48];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const float pin = g == 0 ? pe[c] : pl[c];
#pragma unroll
          for (int fg = 0; fg < 2; ++fg) {
            float sn, cs;
            if (TC_DEBUG_MODE & 8) { sn = pin; cs = pin * 0.5f; } else sincosf(pin * (fg ? 16.0f : 1.0f), &sn, &cs);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
              const int k = 4 * fg + kk;
              v[6 * k + c] = sn; v[6 * k + 3 + c] = cs;
              const float s2 = 2.0f * sn * cs, c2 = 1.0f - 2.0f * sn * sn;
              sn = s2; cs = c2;
            }
          }
        }
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          uint32_t w[16];
          tc::split16(v + 16 * j, w);
          st_a1(3 * g + j, w);
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
        S.m_orig[buf][row] = a.out_by_slot ? (m.valid ? tile_ * 128 + row : -1) : m.orig;
#pragma unroll
        for (int k = 0; k < 3; ++k) { S.m_geo[buf][k][row] = enter[k]; S.m_geo[buf][3 + k][row] = dir[k]; }
      }
      tc::fence_proxy_async();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&S.a1_ready);
    };
    // Gather of A_v[vox][col .. col + 32) for this warp's 32 rows, one epilogue ahead of its use: 8 lanes copy one row's
    // 128 B (4 rows per instruction, fully coalesced) with cp.async straight into the warp's private staging tile; the
    // consumer waits, and every lane (= TMEM lane = pair) reads back its own 32 values.
    float* const stage_w = S.stage[warp];
    const uint32_t stage_sa = tc::smem_u32(stage_w);
    auto gather_issue = [&](int vox_, int col) {
      const float* src0 = a.Av + col + 4 * (lane & 7);
      const float* srcs[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) srcs[i] = src0 + (size_t)((TC_EXP & 32) ? 0 : __shfl_sync(0xffffffffu, vox_, 4 * i + (lane >> 3))) * 512;
      const uint32_t dst0 = stage_sa + (uint32_t)((lane >> 3) * TC_STAGE_PITCH + 4 * (lane & 7)) * 4u;
#pragma unroll
      for (int i = 0; i < 8; ++i) tc::cp_async16(dst0 + (uint32_t)(4 * i * TC_STAGE_PITCH) * 4u, srcs[i]);
      tc::cp_async_commit();
    };
    auto gather_read = [&](float4 (&t)[8]) {
      tc::cp_async_wait_all();
      __syncwarp();
#pragma unroll
      for (int i = 0; i < 8; ++i) t[i] = *reinterpret_cast<const float4*>(stage_w + lane * TC_STAGE_PITCH + 4 * i);
      __syncwarp();                                                            // tile may be overwritten by the next gather
    };
    // Layer-1 epilogue of output half hf (X0 / X1) of a pass of decoder d: this thread converts columns [32 g, 32 g + 32):
    // x = leaky(acc + A_v[vox] + T[ray] (+ u * delta)) -> bf16 hi | lo, in place.  Also issues the gather of the NEXT
    // layer-1 epilogue in program order (next_col < 0: none).
    auto epi_l1 = [&](int hf, int d, bool rank1, float delta, int ray_, bool valid_, uint32_t ph, int next_vox, int next_col) {
      const int n0 = 128 * hf + 32 * g;
      float4 tt[8];
      {
        const float4* tp = reinterpret_cast<const float4*>(a.T + (size_t)((TC_EXP & 16) ? 0 : ray_) * 512 + 256 * d + n0);
#pragma unroll
        for (int i = 0; i < 8; ++i) tt[i] = valid_ ? __ldg(tp + i) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      float4 t[8];
      gather_read(t);
      if (next_col >= 0) gather_issue(next_vox, next_col);
      const uint32_t xcol = (hf ? TC_COL_X1 : TC_COL_X0) + 32 * g;
      tc::mbar_wait(&S.x_full[hf], ph);
      tc::fence_after_sync();
      uint64_t tp2[16];                                                        // T + A_v, as packed pairs
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        tp2[2 * i] = tc::add2(tc::pack2(t[i].x, t[i].y), tc::pack2(tt[i].x, tt[i].y));
        tp2[2 * i + 1] = tc::add2(tc::pack2(t[i].z, t[i].w), tc::pack2(tt[i].z, tt[i].w));
      }
      uint32_t r[32];
      tc::tmem_ld32(lane_addr + xcol, r);
      tc::wait_ld();
      const uint64_t delta2 = tc::pack2(delta, delta);
#pragma unroll
      for (int s16 = 0; s16 < 2; ++s16) {
        uint32_t w[16];
        if (rank1) {
          const float4* up = reinterpret_cast<const float4*>(&S.u[n0 + 16 * s16]);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int e = 16 * s16 + 2 * j;
            const float4 uv = up[j >> 1];
            const uint64_t u2 = (j & 1) ? tc::pack2(uv.z, uv.w) : tc::pack2(uv.x, uv.y);
            const uint64_t v = tc::add2(tc::pack2(__uint_as_float(r[e]), __uint_as_float(r[e + 1])), tp2[e >> 1]);
            tc::leaky_split2(tc::fma2(u2, delta2, v), w[j], w[8 + j]);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int e = 16 * s16 + 2 * j;
            tc::leaky_split2(tc::add2(tc::pack2(__uint_as_float(r[e]), __uint_as_float(r[e + 1])), tp2[e >> 1]), w[j], w[8 + j]);
          }
        }
        tc::tmem_st16(lane_addr + xcol + 16 * s16, w);
      }
      tc::wait_st();
      tc::fence_before_sync();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&S.x_done[hf]);
    };

    // Layer-2 epilogue on Y of a pass of decoder d: x = leaky(acc + b2) -> bf16 hi | lo, in place.
    auto epi_l2 = [&](int d, uint32_t ph) {
      tc::mbar_wait(&S.y_full, ph);
      tc::fence_after_sync();
      uint32_t r[32];
      tc::tmem_ld32(lane_addr + TC_COL_Y + 32 * g, r);
      tc::wait_ld();
      const float4* bp = reinterpret_cast<const float4*>(&S.b2[d][32 * g]);
#pragma unroll
      for (int s16 = 0; s16 < 2; ++s16) {
        uint32_t w[16];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int e = 16 * s16 + 2 * j;
          const float4 bv = bp[e >> 2];
          const uint64_t b2v = (j & 1) ? tc::pack2(bv.z, bv.w) : tc::pack2(bv.x, bv.y);
          tc::leaky_split2(tc::add2(tc::pack2(__uint_as_float(r[e]), __uint_as_float(r[e + 1])), b2v), w[j], w[8 + j]);
        }
        tc::tmem_st16(lane_addr + TC_COL_Y + 32 * g + 16 * s16, w);
      }
      tc::wait_st();
      tc::fence_before_sync();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&S.y_done);
    };
    // Layer-3 epilogue + layer 4 (64-term dot product, 16 terms per column group) on Z of pass (d, it) of the tile whose
    // output metadata sits in buffer `buf`; updates the decoder's running value and, after its last pass, writes the
    // decoder's outputs (column group 0: pred_offset, 1: pred_prob_end, 2: pair_pred_pos).
    uint32_t par = 0;
    float o_a = 0.f, o_b = 0.f;                                // running IEF offsets of decoder 0, 1
    auto epi_l3 = [&](int d, int it, uint32_t ph, int tile_, int buf) {
      tc::mbar_wait(&S.z_full, ph);
      tc::fence_after_sync();
      uint32_t r[16];
      tc::tmem_ld16(lane_addr + TC_COL_Z + 16 * g, r);
      tc::wait_ld();
      tc::fence_before_sync();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&S.z_free);                               // Z may be overwritten by the next pass
      float partial = 0.f;
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int n = 16 * g + j;
        partial = fmaf(lidf_leaky(__uint_as_float(r[j]) + S.b3[d][n]), S.w4[d][n], partial);
      }
      S.part[par][g][row] = partial;
      tc::bar_quadrant(q);                                                     // the 4 warps that share these rows
      const float l4 = ((S.part[par][0][row] + S.part[par][1][row]) + S.part[par][2][row]) + S.part[par][3][row] + S.b4[d];
      par ^= 1;
      const bool is_ief = d == 0 ? ief0 : ief1;
      const float prev = it == 0 ? (is_ief ? a.o0 : 0.f) : (d == 0 ? o_a : o_b);
      const float onew = is_ief ? prev + l4 : l4;
      if (d == 0) o_a = onew; else o_b = onew;
      const bool row_live = (int64_t)tile_ * 128 + row < a.P && (!a.out_by_slot || S.m_orig[buf][row] >= 0);
      if (a.o_iter && d == 0 && g == 0 && it + 1 < npass0 && row_live)
        a.o_iter[(size_t)it * a.P + S.m_orig[buf][row]] = onew;
      if (it + 1 == (d == 0 ? npass0 : npass1) && row_live) {
        const float res = lidf_final_act(onew, d == 0 ? sig0 : sig1);
        const int orig = S.m_orig[buf][row];
        if (d == 0) {
          if (g == 0) a.out[0][orig] = res;
          else if (g == 2) {
            float sc = res * (a.r1 - a.r0) + a.r0;           // pipeline.py:437-439
            sc = sc * a.sqrt3;
            sc = sc * a.part;
#pragma unroll
            for (int k = 0; k < 3; ++k)
              a.pos_out[(size_t)orig * 3 + k] = __fadd_rn(S.m_geo[buf][k][row], __fmul_rn(sc, S.m_geo[buf][3 + k][row]));
          }
        } else if (g == 1) {
          a.out[1][orig] = res;
        }
      }
    };

    // Program order of a row warp (E0/E1 = layer-1 epilogues, E2 = layer 2, E3 = layer 3 + 4 + outputs):
    //   E0(first) | E1(p) E3(p-1) [build next tile's operand] E2(p) E0(p+1) | E1(p+1) E3(p) ...
    // E3 is deferred into the slot where the row warps would otherwise wait for layer 2 of the next pass, and E0 of the
    // next pass runs before it (its accumulator is ready earlier and the tensor pipe needs it sooner) -- unless the next
    // pass is the next IEF iteration of the same decoder, which needs E3's result: then E3 runs right after E2.
    RowMeta cur = load_meta((int)blockIdx.x, load_orig((int)blockIdx.x));
    build_a1(cur, 0, (int)blockIdx.x);
    int ray = cur.ray, vox = cur.vox;
    bool valid = cur.valid;
    uint32_t gp = 0, tl = 0;
    bool pend = false;                                         // E3 of the previous pass is deferred (its arguments are re-derived)
    gather_issue(vox, 256 * pass_dec(0) + 32 * g);
    epi_l1(0, pass_dec(0), false, 0.f, ray, valid, 0u, vox, 256 * pass_dec(0) + 128 + 32 * g);
    for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x, ++tl) {
      const int next_tile = tile + (int)gridDim.x;
      const bool has_next = next_tile < a.n_tiles;
      RowMeta nxt{0, 0, 0, 0.f, 0.f, false};
      int orig_n = -1;
      for (int p = 0; p < npt; ++p, ++gp) {
        const int d = pass_dec(p), it = pass_it(p);
        const uint32_t ph = gp & 1u;
        const bool is_ief = d == 0 ? ief0 : ief1;
        const bool last = p + 1 == npt;
        const bool has_next_pass = !last || has_next;
        const int pn = last ? 0 : p + 1;
        const int dn = pass_dec(pn), itn = pass_it(pn);
        const bool dep = !last && dn == d;                                    // next pass continues this decoder's IEF loop
        if (has_next && p == 0) orig_n = load_orig(next_tile);               // prefetch, stage 1
        if (has_next && p == (npt >= 2 ? npt - 2 : 0)) nxt = load_meta(next_tile, orig_n);   // stage 2: first used in E1 of the last pass
        const int vox_n = last ? nxt.vox : vox, ray_n = last ? nxt.ray : ray;
        const bool valid_n = last ? nxt.valid : valid;
        // ---- E1(p)
        epi_l1(1, d, is_ief && it > 0, (d == 0 ? o_a : o_b) - a.o0, ray, valid, ph, vox_n,
               has_next_pass ? 256 * dn + 32 * g : -1);
        // ---- deferred E3(p-1): fills the wait for layer 2 of this pass
        if (pend) {
          const int pp = p == 0 ? npt - 1 : p - 1;                           // the previous pass: of this tile, or the last one of the previous tile
          epi_l3(pass_dec(pp), pass_it(pp), ph ^ 1u, p == 0 ? tile - (int)gridDim.x : tile, (int)((p == 0 ? tl + 1u : tl) & 1u));
          pend = false;
        }
        // ---- operand of the next tile: its last reader (S2 of this pass) has retired once a1_free completes
        if (last && has_next) {
          tc::mbar_wait(&S.a1_free, tl & 1u);
          build_a1(nxt, (int)((tl + 1) & 1u), next_tile);
        }
        // ---- E2(p)
        epi_l2(d, ph);
        if (has_next_pass && !dep) {
          // ---- E0(p+1) first, E3(p) deferred
          epi_l1(0, dn, (dn == 0 ? ief0 : ief1) && itn > 0, (dn == 0 ? o_a : o_b) - a.o0, ray_n, valid_n, ph ^ 1u, vox_n,
                 256 * dn + 128 + 32 * g);
          pend = true;
        } else {
          // ---- E3(p) now; then E0(p+1) of the dependent pass
          epi_l3(d, it, ph, tile, (int)(tl & 1u));
          if (has_next_pass)
            epi_l1(0, dn, true, (dn == 0 ? o_a : o_b) - a.o0, ray_n, valid_n, ph ^ 1u, vox_n, 256 * dn + 128 + 32 * g);
        }
      }
      ray = nxt.ray; vox = nxt.vox;
      valid = nxt.valid;
    }
  }
  // ---- teardown ----
  __syncwarp();
  tc::fence_before_sync();
  __syncthreads();
  if (warp == TC_ROW_WARPS) tc::tmem_dealloc(tmem, 512);
}

// ------------------------------------------------------------------------------------------------ per-ray row prep
// T[ray][512] = [roi(128) | PE(dir)(27) | 0 x 5] W_row^T + bias for both decoders (the per-ray layer-1 term) on the tensor
// cores with the same split-bf16 arithmetic: 128 rays per tile, K = 160 (10 k-steps), N = 4 x 128 (one TMEM quarter per
// N-chunk).  8 row warps (TMEM quadrant q, column half h), warp 8 = MMA issuer, warp 9 = weight loader.
// Software pipeline over tiles (the kernel moves 2.5 KB per ray: HBM-bound, so the point is to keep loads, MMAs and
// stores of neighbouring tiles in flight together):
//   row warps : build(t + 1) into the other operand buffer  |  epilogue(t): 4 x [TMEM quarter -> + bias -> 2 KB rows to HBM]
//   issuer    : MMAs of tile t + 1, chunk c, start as soon as epilogue(t) has drained quarter c (d_free[c])
//   loader    : streams the 320 KB of packed weights per tile through a 4 x 8 KB ring (slot / parity from running counters)
// Output rows leave through a per-warp shared-memory transpose tile (16-byte chunks XOR-swizzled by row: conflict-free both
// ways) so that 8 lanes write one 128-byte row segment -- a lane storing 16 bytes into its own row touches 32 half-filled
// sectors per instruction.
#define RP_KSTEPS 10
#define RP_ROW_WARPS 8
#define RP_THREADS ((RP_ROW_WARPS + 2) * 32)
#define RP_A_PART_BYTES (RP_KSTEPS * 4096)
#define RP_STAGES 4               // ring of 4 x 8 KB (one k-step per fill)
#define RP_FILLS_PER_TILE 40

struct TcRowPrepArgs {
  const float* feat;               // [rows][128] ROI feature per ray
  const float* dirs;               // [rows][3]
  int64_t rows; int n_tiles;
  const uint8_t* wstream;          // [4 N-chunks][10 k-steps][8 KB]
  const float* bias;               // [512]
  float* out;                      // [rows][512]
  const int* row_list;             // optional: work on rows row_list[0 .. *n_list) only (sparse regime: rays that own a pair)
  const int* n_list;
};
struct TcRowPrepSmem {
  uint8_t w[RP_STAGES][TC_CHUNK_BYTES];
  uint8_t a[2][2][RP_A_PART_BYTES];  // [tile parity][hi|lo][kstep(10)][kgroup(2)][128 rows][16 B]
  float stage[RP_ROW_WARPS][32 * 32];   // per-warp transpose tile of the epilogue: [row][8 x 16 B, chunk ^ (row & 7)]
  uint64_t w_full[RP_STAGES], w_empty[RP_STAGES];
  uint64_t a_ready[2], a_free[2], d_full[4], d_free[4];
  uint32_t tmem_base;
};

// weight stream: chunk (c, s) = N rows 128 c .. 128 c + 127, K elements 16 s .. 16 s + 15 of Wt [Kpad][Ntot]
__global__ void k_pack_rowprep_tc(const float* __restrict__ Wt, int Ntot, uint8_t* __restrict__ stream) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= 4 * RP_KSTEPS * 2048) return;
  const int ch = idx / 2048, r = idx % 2048, c = ch / RP_KSTEPS, ks = ch % RP_KSTEPS, n = r / 16, kk = r % 16;
  const float w = Wt[(size_t)(16 * ks + kk) * Ntot + 128 * c + n];
  const __nv_bfloat16 hi = __float2bfloat16_rn(w);
  const __nv_bfloat16 lo = __float2bfloat16_rn(w - __bfloat162float(hi));
  const size_t off = (size_t)(kk >> 3) * 128 * 16 + (size_t)n * 16 + (kk & 7) * 2;
  uint8_t* base = stream + (size_t)ch * TC_CHUNK_BYTES;
  *reinterpret_cast<__nv_bfloat16*>(base + off) = hi;
  *reinterpret_cast<__nv_bfloat16*>(base + 128 * 32 + off) = lo;
}

template <int NPROD>
__global__ void __launch_bounds__(RP_THREADS, 1) k_rowprep_tc(const __grid_constant__ TcRowPrepArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  TcRowPrepSmem& S = *reinterpret_cast<TcRowPrepSmem*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int i = 0; i < RP_STAGES; ++i) { tc::mbar_init(&S.w_full[i], 1); tc::mbar_init(&S.w_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { tc::mbar_init(&S.a_ready[i], RP_ROW_WARPS); tc::mbar_init(&S.a_free[i], 1); }
    for (int i = 0; i < 4; ++i) { tc::mbar_init(&S.d_full[i], 1); tc::mbar_init(&S.d_free[i], RP_ROW_WARPS); }
    tc::fence_barrier_init();
  }
  if (warp == RP_ROW_WARPS) tc::tmem_alloc(&S.tmem_base, 512);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = S.tmem_base;
  // the row list pays only when it drops a good part of the rows: going through it costs a dependent load per row and
  // scatters the tiles; with more than 3/4 of the rows listed the kernel walks all rows instead (CTA-uniform decision)
  const int* const row_list = (a.row_list && (int64_t)*a.n_list * 4 < a.rows * 3) ? a.row_list : nullptr;
  const int64_t n_rows = row_list ? (int64_t)*a.n_list : a.rows;
  const int n_tiles = row_list ? (int)((n_rows + 127) / 128) : a.n_tiles;
  const int n_my_tiles = n_tiles > (int)blockIdx.x ? (n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

  if (warp == RP_ROW_WARPS + 1) {
    // ---- weight loader: 40 fills of 8 KB per tile (4 N-chunks x 10 k-steps), ring of 4; slot / wrap count run over all tiles
    if (tc::elect_one()) {
      uint32_t slot = 0, use = 0;                      // use = how many times the ring has wrapped (parity of the slot's phase)
      for (int t = 0; t < n_my_tiles; ++t) {
#pragma unroll 1
        for (int f = 0; f < RP_FILLS_PER_TILE; ++f) {
          if (use > 0) tc::mbar_wait(&S.w_empty[slot], (use - 1u) & 1u);
          tc::mbar_arrive_expect_tx(&S.w_full[slot], TC_CHUNK_BYTES);
          tc::bulk_g2s(S.w[slot], a.wstream + (size_t)f * TC_CHUNK_BYTES, TC_CHUNK_BYTES, &S.w_full[slot]);
          if (++slot == RP_STAGES) { slot = 0; ++use; }
        }
      }
    }
  } else if (warp == RP_ROW_WARPS) {
    // ---- MMA issuer
    if (tc::elect_one()) {
      constexpr uint32_t idesc = tc::make_idesc(128);
      const uint64_t wd = tc::make_bdesc(tc::smem_u32(S.w[0]), 2048u, 128u);
      uint32_t slot = 0, use = 0;
      for (int t = 0; t < n_my_tiles; ++t) {
        const uint32_t buf = (uint32_t)t & 1u, uph = ((uint32_t)t >> 1) & 1u, ph = (uint32_t)t & 1u;
        const uint64_t ad = tc::make_bdesc(tc::smem_u32(S.a[buf][0]), 2048u, 128u);
        tc::mbar_wait(&S.a_ready[buf], uph);
        tc::fence_after_sync();
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          if (t > 0) { tc::mbar_wait(&S.d_free[c], ph ^ 1u); tc::fence_after_sync(); }
#pragma unroll 1
          for (int ks = 0; ks < RP_KSTEPS; ++ks) {
            tc::mbar_wait(&S.w_full[slot], use & 1u);
            const uint64_t bhi = wd + ((uint32_t)(slot * TC_CHUNK_BYTES) >> 4);
            const uint64_t ahi = ad + ((uint32_t)(ks * 4096) >> 4);
            tc::mma_ss(tmem + 128 * c, ahi, bhi, idesc, ks == 0 ? 0u : 1u);
            if (NPROD == 3) {
              tc::mma_ss(tmem + 128 * c, ahi + ((uint32_t)RP_A_PART_BYTES >> 4), bhi, idesc, 1u);
              tc::mma_ss(tmem + 128 * c, ahi, bhi + (4096u >> 4), idesc, 1u);
            }
            tc::commit(&S.w_empty[slot]);
            if (++slot == RP_STAGES) { slot = 0; ++use; }
          }
          tc::commit(&S.d_full[c]);
        }
        tc::commit(&S.a_free[buf]);
      }
    }
  } else {
    // ---- row warps: operand build (one tile ahead) + epilogue
    const int q = warp & 3, h = warp >> 2;
    const int row = q * 32 + lane;
    const uint32_t lane_addr = tmem + ((uint32_t)(q * 32) << 16);
    auto build = [&](int t) {
      const int tile = (int)blockIdx.x + t * (int)gridDim.x;
      const int64_t row0 = (int64_t)tile * 128;
      const uint32_t buf = (uint32_t)t & 1u;
      const uint32_t a_hi = tc::smem_u32(S.a[buf][0]), a_lo = tc::smem_u32(S.a[buf][1]);
      if (t >= 2) tc::mbar_wait(&S.a_free[buf], (((uint32_t)t >> 1) - 1u) & 1u);   // MMAs of tile t - 2 have read this buffer
      // feature part (k-steps 0-7): one instruction = 8 rows x one k-step, 4 lanes x 16 B per row; 8 loads in flight per thread
      {
        const int r8 = lane >> 2, j4 = lane & 3;
#pragma unroll 1
        for (int it0 = warp; it0 < 16 * 8; it0 += 8 * RP_ROW_WARPS) {
          float4 v[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int it = it0 + u * RP_ROW_WARPS, rg = it >> 3, ks = it & 7, r = 8 * rg + r8;
            v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (row0 + r < n_rows) {
              const int64_t src = row_list ? (int64_t)row_list[row0 + r] : row0 + r;
              v[u] = __ldg(reinterpret_cast<const float4*>(a.feat + (size_t)src * 128 + 16 * ks) + j4);
            }
          }
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int it = it0 + u * RP_ROW_WARPS, rg = it >> 3, ks = it & 7, r = 8 * rg + r8;
            uint32_t h0, l0, h1, l1;
            tc::split2(v[u].x, v[u].y, h0, l0);
            tc::split2(v[u].z, v[u].w, h1, l1);
            const uint32_t off = (uint32_t)(ks * 4096 + (j4 >> 1) * 2048 + r * 16 + 8 * (j4 & 1));
            asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(a_hi + off), "r"(h0), "r"(h1) : "memory");
            asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(a_lo + off), "r"(l0), "r"(l1) : "memory");
          }
        }
      }
      // PE(dir) part (k-steps 8, 9): Embedder order [x, sin f0 x, cos f0 x, sin f1 x, ...], f = 1, 2, 4, 8; 27 values
      if (h == 0) {
        float x[32];
#pragma unroll
        for (int k = 0; k < 32; ++k) x[k] = 0.f;
        if (row0 + row < n_rows) {
          const float* dp = a.dirs + (size_t)(row_list ? (int64_t)row_list[row0 + row] : row0 + row) * 3;
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            const float dv = dp[c];
            x[c] = dv;
            float sn, cs;
            sincosf(dv, &sn, &cs);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              x[3 + 6 * k + c] = sn; x[3 + 6 * k + 3 + c] = cs;
              const float s2 = 2.0f * sn * cs, c2 = 1.0f - 2.0f * sn * sn;
              sn = s2; cs = c2;
            }
          }
        }
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          uint32_t w[16];
          tc::split16(x + 16 * j, w);
          const uint32_t off = (uint32_t)((8 + j) * 4096 + row * 16);
          tc::st_shared_v4(a_hi + off, w[0], w[1], w[2], w[3]);
          tc::st_shared_v4(a_hi + off + 2048, w[4], w[5], w[6], w[7]);
          tc::st_shared_v4(a_lo + off, w[8], w[9], w[10], w[11]);
          tc::st_shared_v4(a_lo + off + 2048, w[12], w[13], w[14], w[15]);
        }
      }
      tc::fence_proxy_async();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&S.a_ready[buf]);
    };
    if (n_my_tiles > 0) build(0);
    for (int t = 0; t < n_my_tiles; ++t) {
      if (t + 1 < n_my_tiles) build(t + 1);
      const int tile = (int)blockIdx.x + t * (int)gridDim.x;
      const int64_t row0 = (int64_t)tile * 128;
      const uint32_t ph = (uint32_t)t & 1u;
      // epilogue: this thread pulls columns [64 h, 64 h + 64) of every N-chunk of its row out of TMEM; the warp transposes
      // them through its staging tile so that lane l stores 16 bytes of row 4 i + (l >> 3), i = 0..7 (8 lanes = 128 B of a row)
      float* orows[8];
      bool olive[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int64_t rr = row0 + 32 * q + 4 * i + (lane >> 3);
        olive[i] = rr < n_rows;
        orows[i] = a.out + (size_t)(olive[i] && row_list ? (int64_t)row_list[rr] : rr) * 512 + 64 * h + 4 * (lane & 7);
      }
      float* const stg = S.stage[warp];
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        tc::mbar_wait(&S.d_full[c], ph);
        tc::fence_after_sync();
        uint32_t r0[32], r1[32];
        tc::tmem_ld32(lane_addr + 128 * c + 64 * h, r0);
        tc::tmem_ld32(lane_addr + 128 * c + 64 * h + 32, r1);
        tc::wait_ld();
        tc::fence_before_sync();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&S.d_free[c]);                          // the quarter is in registers: release it first
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          const uint32_t* r = k ? r1 : r0;
#pragma unroll
          for (int j = 0; j < 8; ++j)
            *reinterpret_cast<uint4*>(stg + lane * 32 + 4 * (j ^ (lane & 7))) = make_uint4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
          __syncwarp();
          const float4 b4 = a.bias ? __ldg(reinterpret_cast<const float4*>(a.bias + 128 * c + 64 * h + 32 * k) + (lane & 7))
                                   : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int rr = 4 * i + (lane >> 3);
            float4 o = *reinterpret_cast<const float4*>(stg + rr * 32 + 4 * ((lane & 7) ^ (rr & 7)));
            o.x += b4.x; o.y += b4.y; o.z += b4.z; o.w += b4.w;
            if (olive[i]) __stcs(reinterpret_cast<float4*>(orows[i] + 128 * c + 32 * k), o);
          }
          __syncwarp();                                                        // the tile is rewritten by the next round
        }
      }
    }
  }
  __syncwarp();
  tc::fence_before_sync();
  __syncthreads();
  if (warp == RP_ROW_WARPS) tc::tmem_dealloc(tmem, 512);
}

// ------------------------------------------------------------------------------------------------ host side
struct TcBufs { uint8_t* wstream; uint8_t* rp_wstream; };

template <typename B>
inline TcBufs carve_tc(B& b, int64_t V, int n_dec) {
  TcBufs t;
  (void)V;
  t.wstream = b.template take<uint8_t>((size_t)n_dec * TC_CHUNKS_PER_DEC * TC_CHUNK_BYTES);
  t.rp_wstream = b.template take<uint8_t>((size_t)4 * RP_KSTEPS * TC_CHUNK_BYTES);
  return t;
}

#define TC_LAUNCH_CHECK()                                                                         \
  do {                                                                                            \
    ++(*launches);                                                                                \
    cudaError_t e__ = cudaGetLastError();                                                         \
    if (e__ != cudaSuccess) {                                                                     \
      snprintf(errbuf, errlen, "%s:%d %s", __FILE__, __LINE__, cudaGetErrorString(e__));          \
      return LIDF_ERR_CUDA;                                                                       \
    }                                                                                             \
  } while (0)

inline int tc_device_ok() {
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return 0;
  return major == 10;
}

struct TcArgs;
inline int tc_launch(TcArgs& a, cudaStream_t st, int64_t* launches, char* errbuf, size_t errlen,
                     void (*mlp_event)(int, cudaStream_t));

// decoders + pair_pred_pos for all P pairs; T = per-ray layer-1 term [R][512], u = IEF rank-1 vector of offset_dec
// phase: 0 = both decoders over all P pairs; 1 = the probability decoder only (winner-only mode, first half);
// 2 = the offset decoder over one row per ray (winner-only mode, second half): perm = win [R] (the ray's arg-max pair, -1 for
// a ray without pairs), outputs by ray: pred_pos [R,3] directly and the offset into off_ray [R].
inline int tc_query_forward(const LidfQueryParams* p, const TcBufs& tb, const int* perm, const float* T, const float* Av,
                            const float* u,
                            int pe_pos, int D, int impl, cudaStream_t st, int64_t* launches, char* errbuf, size_t errlen,
                            void (*mlp_event)(int, cudaStream_t), int phase = 0, float* off_ray = nullptr) {
  if (!p->pos_encode || p->multires != 8 || pe_pos != 51) return LIDF_ERR_UNSUPPORTED;   // A1 layout is built for PE(8)
  if (!tc_device_ok()) return LIDF_ERR_NO_SM100;
  const LidfDecoder* decs[2] = {&p->offset_dec, &p->prob_dec};
  for (int d = 0; d < 2 && phase != 2 && !(p->weight_cache && p->weight_cache_valid); ++d) {
    const int ldw = D + (decs[d]->kind == LIDF_DEC_IEF ? LIDF_IEF_ENC : 0);
    k_pack_tc_weights<<<(TC_CHUNKS_PER_DEC * 2048 + 255) / 256, 256, 0, st>>>(
        decs[d]->w1, ldw, pe_pos, decs[d]->w2, decs[d]->w3, tb.wstream + (size_t)d * TC_CHUNKS_PER_DEC * TC_CHUNK_BYTES, 0);
    TC_LAUNCH_CHECK();
  }
  TcArgs a{};
  a.P = p->P; a.n_tiles = (int)((p->P + 127) / 128);
  a.perm = perm; a.pair_vox = p->pair_vox; a.pair_ray = p->pair_ray; a.pair_dist = p->pair_dist;
  a.dense_dist = p->dense_dist; a.R = p->R; a.V = p->V; a.ray_dir = p->miss_ray_dir; a.voxel_bound = p->voxel_bound;
  a.rel = p->intersect_pos_rel; a.Av = Av; a.T = T; a.wstream = tb.wstream;
  a.o_iter = decs[0]->kind == LIDF_DEC_IEF ? p->ief_iter_out : nullptr;
  a.u = decs[0]->kind == LIDF_DEC_IEF ? u : nullptr;
  for (int d = 0; d < 2; ++d) {
    a.b2[d] = decs[d]->b2; a.b3[d] = decs[d]->b3; a.w4[d] = decs[d]->w4; a.b4[d] = decs[d]->b4;
    a.kind[d] = decs[d]->kind; a.n_pass[d] = decs[d]->kind == LIDF_DEC_IEF ? decs[d]->n_iter : 1;
    a.use_sigmoid[d] = decs[d]->use_sigmoid;
  }
  a.o0 = decs[0]->init_offset; a.r0 = p->offset_range0; a.r1 = p->offset_range1;
  a.sqrt3 = (float)sqrt(3.0); a.part = p->part_size;
  a.out[0] = p->pred_offset; a.out[1] = p->pred_prob_end; a.pos_out = p->pair_pred_pos;
  a.n_prod = impl == LIDF_MLP_TC_BF16X1 ? 1 : 3;
  if (phase == 1) {
    a.n_pass[0] = 0; a.o_iter = nullptr;
  } else if (phase == 2) {
    a.n_pass[1] = 0;
    a.P = p->R; a.n_tiles = (int)((p->R + 127) / 128);
    a.out_by_slot = 1; a.out[0] = off_ray; a.pos_out = p->pred_pos;      // a.o_iter (if any) is [n_iter-1][R], by ray
  }
  return tc_launch(a, st, launches, errbuf, errlen, phase == 2 ? nullptr : mlp_event);
}

// pass schedule + launch of k_mlp_tc (a.n_pass[] / a.kind[] set by the caller)
inline int tc_launch(TcArgs& a, cudaStream_t st, int64_t* launches, char* errbuf, size_t errlen,
                     void (*mlp_event)(int, cudaStream_t)) {
  if (a.n_pass[0] + a.n_pass[1] > TC_MAX_PASSES || a.n_pass[0] + a.n_pass[1] < 1) return LIDF_ERR_UNSUPPORTED;
  {  // interleave the two decoders' passes: d0 it0, d1 it0, d0 it1, d1 it1, ... then whatever is left of the longer one
    int it[2] = {0, 0};
    a.npt = 0; a.pass_dec_mask = 0; a.pass_it_pack = 0;
    while (it[0] < a.n_pass[0] || it[1] < a.n_pass[1])
      for (int d = 0; d < 2; ++d)
        if (it[d] < a.n_pass[d]) {
          a.pass_dec_mask |= (uint32_t)d << a.npt;
          a.pass_it_pack |= (uint64_t)it[d]++ << (4 * a.npt);
          ++a.npt;
        }
  }
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = a.n_tiles < sms ? a.n_tiles : sms;
  const size_t smem = sizeof(TcSmem) + 1024;
  auto kern = a.n_prod == 3 ? k_mlp_tc<3> : k_mlp_tc<1>;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
    snprintf(errbuf, errlen, "cudaFuncSetAttribute(k_mlp_tc, %zu) failed: %s", smem, cudaGetErrorString(cudaGetLastError()));
    return LIDF_ERR_CUDA;
  }
  if (mlp_event) mlp_event(0, st);
  kern<<<grid, TC_THREADS, smem, st>>>(a);
  if (mlp_event) mlp_event(1, st);
  TC_LAUNCH_CHECK();
  return LIDF_OK;
}

// RefineNet decoder tail on the same engine: row = ray, "voxel of the pair" = end_voxel_id[ray], one decoder (its IEF passes
// run back to back), layer-1 MMA operand = PE(pos [- centre]) in the enter slots (leave slots carry zero weights),
// pred_pos_refine = pos + (o (r1 - r0) + r0) dir  (pipeline.py:1018-1029: no sqrt(3) part_size factor here).
inline int tc_refine_forward(const LidfRefineParams* p, const TcBufs& tb, const float* T, const float* Av, const float* u,
                             float* scratch_out, int pe_pos, int D, int impl, cudaStream_t st, int64_t* launches,
                             char* errbuf, size_t errlen) {
  if (!p->pos_encode || p->multires != 8 || pe_pos != 51) return LIDF_ERR_UNSUPPORTED;
  if (!tc_device_ok()) return LIDF_ERR_NO_SM100;
  const LidfDecoder& dc = p->offset_dec;
  const int ldw = D + (dc.kind == LIDF_DEC_IEF ? LIDF_IEF_ENC : 0);
  k_pack_tc_weights<<<(TC_CHUNKS_PER_DEC * 2048 + 255) / 256, 256, 0, st>>>(dc.w1, ldw, pe_pos, dc.w2, dc.w3, tb.wstream, 1);
  TC_LAUNCH_CHECK();
  TcArgs a{};
  a.P = p->R; a.n_tiles = (int)((p->R + 127) / 128);
  a.perm = nullptr; a.pair_vox = p->end_voxel_id; a.pair_ray = nullptr; a.pair_dist = nullptr; a.dense_dist = nullptr;
  a.R = p->R; a.V = p->V; a.ray_dir = p->miss_ray_dir; a.voxel_bound = p->voxel_bound; a.rel = p->intersect_pos_rel;
  a.pos_in = p->pred_pos; a.Av = Av; a.T = T; a.wstream = tb.wstream;
  a.u = dc.kind == LIDF_DEC_IEF ? u : nullptr;
  for (int d = 0; d < 2; ++d) { a.b2[d] = dc.b2; a.b3[d] = dc.b3; a.w4[d] = dc.w4; a.b4[d] = dc.b4; a.use_sigmoid[d] = dc.use_sigmoid; }
  a.kind[0] = dc.kind; a.n_pass[0] = dc.kind == LIDF_DEC_IEF ? dc.n_iter : 1;
  a.kind[1] = LIDF_DEC_IMNET; a.n_pass[1] = 0;
  a.o0 = dc.init_offset; a.r0 = p->offset_range0; a.r1 = p->offset_range1; a.sqrt3 = 1.f; a.part = 1.f;
  a.out[0] = scratch_out; a.out[1] = scratch_out; a.pos_out = p->pred_pos_refine;
  a.n_prod = impl == LIDF_MLP_TC_BF16X1 ? 1 : 3;
  return tc_launch(a, st, launches, errbuf, errlen, nullptr);
}

// per-ray layer-1 term on the tensor cores; Wt = sp.Wt_row [160][512], wstream scratch = 4 x 10 x 8 KB
inline int tc_rowprep_forward(const float* roi_feat, const float* dirs, int64_t R, const float* Wt, int Ntot, const float* bias,
                              uint8_t* wstream, float* T, int impl, cudaStream_t st, int64_t* launches, char* errbuf,
                              size_t errlen, bool packed = false, const int* row_list = nullptr, const int* n_list = nullptr) {
  if (!packed) {
    k_pack_rowprep_tc<<<(4 * RP_KSTEPS * 2048 + 255) / 256, 256, 0, st>>>(Wt, Ntot, wstream);
    TC_LAUNCH_CHECK();
  }
  TcRowPrepArgs a{};
  a.feat = roi_feat; a.dirs = dirs; a.rows = R; a.n_tiles = (int)((R + 127) / 128);
  a.wstream = wstream; a.bias = bias; a.out = T; a.row_list = row_list; a.n_list = n_list;
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = a.n_tiles < sms ? a.n_tiles : sms;
  const size_t smem = sizeof(TcRowPrepSmem) + 1024;
  auto kern = impl == LIDF_MLP_TC_BF16X1 ? k_rowprep_tc<1> : k_rowprep_tc<3>;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
    snprintf(errbuf, errlen, "cudaFuncSetAttribute(k_rowprep_tc, %zu) failed: %s", smem, cudaGetErrorString(cudaGetLastError()));
    return LIDF_ERR_CUDA;
  }
  kern<<<grid, RP_THREADS, smem, st>>>(a);
  TC_LAUNCH_CHECK();
  return LIDF_OK;
}

// self test entry (exported through lidf_query.cu): D = A W^T, A [128][32], W [128][32], D [128][128], all device fp32
inline int tc_selftest(const float* A, const float* W, float* D, uint8_t* scratch16k, int variant, cudaStream_t st,
                       int64_t* launches, char* errbuf, size_t errlen) {
  if (!tc_device_ok()) return LIDF_ERR_NO_SM100;
  k_tc_selftest_pack<<<16, 256, 0, st>>>(W, scratch16k);
  TC_LAUNCH_CHECK();
  k_tc_selftest<<<1, 128, 0, st>>>(scratch16k, A, D, variant);
  TC_LAUNCH_CHECK();
  return LIDF_OK;
}
