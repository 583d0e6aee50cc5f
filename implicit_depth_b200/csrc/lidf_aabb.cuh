// lidf_aabb.cuh -- ray / point vs voxel-box tests (include/lidf_aabb.h), sm_100a.
//
// Bit-identical to the reference kernels (src/extensions/ray_aabb/ray_aabb_cuda_kernel.cu:10-89,
// src/extensions/pcl_aabb/pcl_aabb_cuda_kernel.cu:10-45).  What differs is the work and the traffic:
//   * the reference evaluates three double-precision reciprocals 1/(d+1e-12) per (voxel, ray); here once per ray
//     (k_aabb_ray_prep), the slab test itself is 6 single-precision products + compares;
//   * the reference launches V x ceil(R/1024) blocks that mostly exit on `ray_bid != voxel_bid`; here a persistent grid
//     (a multiple of the SM count) walks (voxel, 1024-ray tile) work items and skips a tile whose image-id range
//     cannot contain the voxel's image after two loads;
//   * the pair list is produced directly (count -> scan -> fill, in torch.nonzero order) instead of writing and
//     re-reading the dense 12*V*R-byte mask/dist slab; the dense outputs exist only as a drop-in.
// All kernels are HBM/L2-bound byte movers: coalesced 4/8-byte accesses per thread, streaming stores for write-once
// outputs, no tensor-core work.
#pragma once
#include "lidf_aabb.h"
#include "lidf_common.cuh"
#include "lidf_prep.cuh"

#define AABB_THREADS 256
#define AABB_RPT 4                              // rays (points) per thread
#define AABB_TILE (AABB_THREADS * AABB_RPT)      // 1024 rays per work item -- the reference's block size
#define AABB_WARPS (AABB_THREADS / 32)

struct AabbBox { float lo[3], hi[3]; };

__device__ __forceinline__ AabbBox aabb_load_box(const float* __restrict__ voxel_bound, int64_t v) {
  AabbBox b;
  const float* p = voxel_bound + v * 6;
  b.lo[0] = __ldg(p + 0); b.lo[1] = __ldg(p + 1); b.lo[2] = __ldg(p + 2);
  b.hi[0] = __ldg(p + 3); b.hi[1] = __ldg(p + 4); b.hi[2] = __ldg(p + 5);
  return b;
}

// ray_aabb_cuda_kernel.cu:28-83 with the reciprocal direction already formed.  __fmul_rn keeps the six products
// plain IEEE multiplies (the reference has no multiply-add here either).
__device__ __forceinline__ bool aabb_slab(float ix, float iy, float iz, const AabbBox& b, float& t_enter, float& t_leave) {
  const float txmin = __fmul_rn(ix >= 0 ? b.lo[0] : b.hi[0], ix), txmax = __fmul_rn(ix >= 0 ? b.hi[0] : b.lo[0], ix);
  const float tymin = __fmul_rn(iy >= 0 ? b.lo[1] : b.hi[1], iy), tymax = __fmul_rn(iy >= 0 ? b.hi[1] : b.lo[1], iy);
  float tmin_max = txmin, tmax_min = txmax;
  if ((tmin_max > tymax) || (tmax_min < tymin)) return false;
  tmin_max = fmaxf(tmin_max, tymin);
  tmax_min = fminf(tmax_min, tymax);
  const float tzmin = __fmul_rn(iz >= 0 ? b.lo[2] : b.hi[2], iz), tzmax = __fmul_rn(iz >= 0 ? b.hi[2] : b.lo[2], iz);
  if ((tmin_max > tzmax) || (tmax_min < tzmin)) return false;
  t_enter = fmaxf(tmin_max, tzmin);
  t_leave = fminf(tmax_min, tzmax);
  return true;
}

// pcl_aabb_cuda_kernel.cu:28-42 (closed box; the comparisons are written exactly as the reference's so NaN behaves alike)
__device__ __forceinline__ bool aabb_inside(float x, float y, float z, const AabbBox& b) {
  if ((x < b.lo[0]) || (x > b.hi[0])) return false;
  if ((y < b.lo[1]) || (y > b.hi[1])) return false;
  if ((z < b.lo[2]) || (z > b.hi[2])) return false;
  return true;
}

// ---- pre-pass: reciprocal directions + image-id range + direction bounds of every 1024-ray tile ---------------------
// tile_tan[rb] = (min dx/dz, max dx/dz, min dy/dz, max dy/dz) over the tile's rays, or (-inf, +inf, -inf, +inf) when some
// ray of the tile does not look down +z: a conservative "frustum" of the tile used to cull (voxel, tile) work items.
__global__ void __launch_bounds__(AABB_THREADS) k_aabb_ray_prep(const float* __restrict__ ray_dir, const int32_t* __restrict__ ray_bid,
                                                                int64_t R, float* __restrict__ inv, int2* __restrict__ tile_bid,
                                                                float4* __restrict__ tile_tan) {
  __shared__ int s_min, s_max, s_bad;
  __shared__ float s_t[AABB_WARPS][4];
  if (threadIdx.x == 0) { s_min = INT_MAX; s_max = INT_MIN; s_bad = 0; }
  __syncthreads();
  int lo = INT_MAX, hi = INT_MIN, bad = 0;
  float t[4] = {INFINITY, -INFINITY, INFINITY, -INFINITY};
  const int64_t base = (int64_t)blockIdx.x * AABB_TILE;
#pragma unroll
  for (int j = 0; j < AABB_RPT; ++j) {
    const int64_t r = base + j * AABB_THREADS + threadIdx.x;
    if (r < R) {
      const int b = ray_bid[r];
      lo = min(lo, b); hi = max(hi, b);
      float d[3];
#pragma unroll
      for (int a = 0; a < 3; ++a) {     // float(1 / (double(d) + 1e-12)): ray_aabb_cuda_kernel.cu:32,48,67
        d[a] = ray_dir[r * 3 + a];
        inv[r * 3 + a] = __double2float_rn(1.0 / ((double)d[a] + 1e-12));
      }
      if (d[2] > 1e-6f && isfinite(d[0]) && isfinite(d[1])) {
        const float tx = d[0] / d[2], ty = d[1] / d[2];
        t[0] = fminf(t[0], tx); t[1] = fmaxf(t[1], tx); t[2] = fminf(t[2], ty); t[3] = fmaxf(t[3], ty);
      } else {
        bad = 1;
      }
    }
  }
  lo = __reduce_min_sync(0xffffffffu, lo); hi = __reduce_max_sync(0xffffffffu, hi);
  bad = __any_sync(0xffffffffu, bad);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    t[0] = fminf(t[0], __shfl_xor_sync(0xffffffffu, t[0], o)); t[1] = fmaxf(t[1], __shfl_xor_sync(0xffffffffu, t[1], o));
    t[2] = fminf(t[2], __shfl_xor_sync(0xffffffffu, t[2], o)); t[3] = fmaxf(t[3], __shfl_xor_sync(0xffffffffu, t[3], o));
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMin(&s_min, lo); atomicMax(&s_max, hi);
    if (bad) atomicOr(&s_bad, 1);
#pragma unroll
    for (int k = 0; k < 4; ++k) s_t[threadIdx.x >> 5][k] = t[k];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    tile_bid[blockIdx.x] = make_int2(s_min, s_max);
    float4 o = make_float4(INFINITY, -INFINITY, INFINITY, -INFINITY);
    for (int w = 0; w < AABB_WARPS; ++w) {
      o.x = fminf(o.x, s_t[w][0]); o.y = fmaxf(o.y, s_t[w][1]); o.z = fminf(o.z, s_t[w][2]); o.w = fmaxf(o.w, s_t[w][3]);
    }
    if (s_bad) o = make_float4(-INFINITY, INFINITY, -INFINITY, INFINITY);
    tile_tan[blockIdx.x] = o;
  }
}

// Can any ray of the tile meet the box?  A line through the origin with direction d (dz > 0) meets a box with z in
// [zlo, zhi], zlo > 0, only at points with x/z in [min(xlo/zlo, xlo/zhi), max(xhi/zlo, xhi/zhi)] (same for y), and x/z is
// dx/dz on the whole line.  The interval is widened by 1e-4 (relative + absolute), four orders of magnitude above the fp32
// rounding of either side, so the cull never drops a pair the slab test would report.  Boxes that touch z <= 0 and tiles
// with rays not looking down +z are never culled.
__device__ __forceinline__ bool aabb_tile_may_hit(const float4 tt, const AabbBox& b) {
  if (!(b.lo[2] > 1e-6f) || !(b.hi[2] >= b.lo[2]) || !(b.hi[0] >= b.lo[0]) || !(b.hi[1] >= b.lo[1])) return true;
  const float ilo = 1.0f / b.lo[2], ihi = 1.0f / b.hi[2];
  const float xl = fminf(b.lo[0] * ilo, b.lo[0] * ihi), xh = fmaxf(b.hi[0] * ilo, b.hi[0] * ihi);
  const float yl = fminf(b.lo[1] * ilo, b.lo[1] * ihi), yh = fmaxf(b.hi[1] * ilo, b.hi[1] * ihi);
  const float mx = 1e-4f * (1.0f + fmaxf(fabsf(xl), fabsf(xh))), my = 1e-4f * (1.0f + fmaxf(fabsf(yl), fabsf(yh)));
  return !(tt.y < xl - mx || tt.x > xh + mx || tt.w < yl - my || tt.z > yh + my);
}

// ---- compact pair list: live work items --------------------------------------------------------------------------
// One thread per (voxel, ray-tile) item: keep it if the tile's image-id range contains the voxel's image and the tile's
// direction bounds can meet the box.  Survivors are appended to `items` (warp-aggregated atomic; the order is irrelevant:
// an item's output position comes from the scan of cnt[], indexed by the item id).  count / fill then walk only the
// survivors -- at the bench geometry 4.9 M items shrink to a few 10^5, and the per-item latency chain (two dependent
// loads + a 64-bit division) disappears from the persistent loops.
__global__ void k_aabb_items(const float* __restrict__ voxel_bound, const int32_t* __restrict__ voxel_bid,
                             const int2* __restrict__ tile_bid, const float4* __restrict__ tile_tan, int64_t RB, int64_t M,
                             int* __restrict__ items, int* __restrict__ n_items) {
  const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  bool live = false;
  if (w < M) {
    const int64_t v = w / RB, rb = w - v * RB;
    const int vbid = __ldg(voxel_bid + v);
    const int2 tb = __ldg(tile_bid + rb);
    if (!(vbid < tb.x || vbid > tb.y)) live = aabb_tile_may_hit(__ldg(tile_tan + rb), aabb_load_box(voxel_bound, v));
  }
  const unsigned m = __ballot_sync(0xffffffffu, live);
  if (m == 0) return;
  const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
  int base = 0;
  if (lane == leader) base = atomicAdd(n_items, __popc(m));
  base = __shfl_sync(0xffffffffu, base, leader);
  if (live) items[base + __popc(m & ((1u << lane) - 1u))] = (int)w;
}

// ---- compact pair list: count (cnt[] zeroed by the caller; only live items are visited) ---------------------------------
__global__ void __launch_bounds__(AABB_THREADS) k_aabb_count(const float* __restrict__ inv, const float* __restrict__ voxel_bound,
                                                             const int32_t* __restrict__ ray_bid, const int32_t* __restrict__ voxel_bid,
                                                             const int* __restrict__ items, const int* __restrict__ n_items,
                                                             int64_t R, int64_t RB, int* __restrict__ cnt) {
  const int n = *n_items;
  for (int it = blockIdx.x; it < n; it += gridDim.x) {
    const int64_t w = items[it];
    const int64_t v = w / RB, rb = w - v * RB;
    const int vbid = __ldg(voxel_bid + v);
    const AabbBox box = aabb_load_box(voxel_bound, v);
    int total = 0;
#pragma unroll
    for (int j = 0; j < AABB_RPT; ++j) {
      const int64_t r = rb * AABB_TILE + j * AABB_THREADS + threadIdx.x;
      bool hit = false;
      if (r < R && ray_bid[r] == vbid) {
        float t0, t1;
        hit = aabb_slab(inv[r * 3 + 0], inv[r * 3 + 1], inv[r * 3 + 2], box, t0, t1);
      }
      total += __syncthreads_count(hit);
    }
    if (threadIdx.x == 0) cnt[w] = total;
  }
}

// ---- compact pair list: fill (torch.nonzero order: voxel, then ray) --------------------------------------------------
__global__ void __launch_bounds__(AABB_THREADS) k_aabb_fill(const float* __restrict__ inv, const float* __restrict__ voxel_bound,
                                                            const int32_t* __restrict__ ray_bid, const int32_t* __restrict__ voxel_bid,
                                                            const int* __restrict__ items, const int* __restrict__ n_items,
                                                            int64_t R, int64_t RB, const int* __restrict__ start,
                                                            int64_t* __restrict__ pair_vox, int64_t* __restrict__ pair_ray,
                                                            float2* __restrict__ pair_dist) {
  __shared__ int s_off[AABB_RPT * AABB_WARPS];                       // 32 (j, warp) groups, in ray order
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n = *n_items;
  for (int it = blockIdx.x; it < n; it += gridDim.x) {
    const int64_t w = items[it];
    const int base = __ldg(start + w);
    if (__ldg(start + w + 1) == base) continue;                     // nothing to write for this (voxel, tile)
    const int64_t v = w / RB, rb = w - v * RB;
    const int vbid = __ldg(voxel_bid + v);
    const AabbBox box = aabb_load_box(voxel_bound, v);
    unsigned ball[AABB_RPT];
    float t0[AABB_RPT], t1[AABB_RPT];
#pragma unroll
    for (int j = 0; j < AABB_RPT; ++j) {
      const int64_t r = rb * AABB_TILE + j * AABB_THREADS + threadIdx.x;
      bool hit = false;
      if (r < R && ray_bid[r] == vbid) hit = aabb_slab(inv[r * 3 + 0], inv[r * 3 + 1], inv[r * 3 + 2], box, t0[j], t1[j]);
      ball[j] = __ballot_sync(0xffffffffu, hit);
      if (lane == 0) s_off[j * AABB_WARPS + warp] = __popc(ball[j]);
    }
    __syncthreads();
    if (warp == 0) {                                                 // exclusive prefix over the 32 groups
      const int c = s_off[lane];
      int inc = c;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
      s_off[lane] = inc - c;
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < AABB_RPT; ++j) {
      if ((ball[j] >> lane) & 1u) {
        const int64_t r = rb * AABB_TILE + j * AABB_THREADS + threadIdx.x;
        const int64_t pos = (int64_t)base + s_off[j * AABB_WARPS + warp] + __popc(ball[j] & ((1u << lane) - 1u));
        pair_vox[pos] = v;
        pair_ray[pos] = r;
        pair_dist[pos] = make_float2(t0[j], t1[j]);
      }
    }
    __syncthreads();                                                 // s_off is reused by the next work item
  }
}

// ---- compact pair list, RAY-MAJOR (sorted by ray, then voxel) + its CSR ------------------------------------------------
// The consumer of the pair list (lidf_query_forward) wants all pairs of a ray adjacent; emitting them that way makes its
// regroup (count / scan / scatter / segment sort over all P pairs) a binary search per ray.  One thread per ray, one block
// per 256 consecutive rays.  The block forms the image-id range and the direction bounds ("frustum") of its rays, then
// walks the voxels in tiles of 256: thread i keeps voxel v0 + i if its image id lies in the range and the frustum can meet
// the box (aabb_tile_may_hit, the same conservative test as the voxel-major generator); survivors are compacted IN ORDER
// into shared memory, and every ray runs the exact slab test (aabb_slab: bit-identical enter / leave distances) against
// that short list only.  Pass 1 (FILL = false) counts per ray; after a scan, pass 2 writes each ray's pairs contiguously
// in ascending voxel order -- no atomics, deterministic.
#define AABB_RM_THREADS 256
template <bool FILL>
__global__ void __launch_bounds__(AABB_RM_THREADS) k_aabb_ray_major(const float* __restrict__ ray_dir, const float* __restrict__ voxel_bound,
                                                                    const int32_t* __restrict__ ray_bid, const int32_t* __restrict__ voxel_bid,
                                                                    int64_t R, int64_t V, int* __restrict__ cnt, const int* __restrict__ start,
                                                                    int64_t* __restrict__ pair_vox, int64_t* __restrict__ pair_ray,
                                                                    float2* __restrict__ pair_dist) {
  __shared__ float s_box[AABB_RM_THREADS * 6];
  __shared__ int s_vid[AABB_RM_THREADS], s_vb[AABB_RM_THREADS];
  __shared__ int s_woff[AABB_RM_THREADS / 32 + 1];
  __shared__ float s_t[AABB_RM_THREADS / 32][4];
  __shared__ int s_bmin, s_bmax, s_bad;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t r = (int64_t)blockIdx.x * AABB_RM_THREADS + threadIdx.x;
  const bool act = r < R;
  if (threadIdx.x == 0) { s_bmin = INT_MAX; s_bmax = INT_MIN; s_bad = 0; }
  __syncthreads();
  float inv[3] = {0.f, 0.f, 0.f};
  int bid = 0;
  {
    int lo = INT_MAX, hi = INT_MIN, bad = 0;
    float t[4] = {INFINITY, -INFINITY, INFINITY, -INFINITY};
    if (act) {
      bid = ray_bid[r];
      lo = hi = bid;
      float d[3];
#pragma unroll
      for (int a = 0; a < 3; ++a) {     // float(1 / (double(d) + 1e-12)): ray_aabb_cuda_kernel.cu:32,48,67
        d[a] = ray_dir[r * 3 + a];
        inv[a] = __double2float_rn(1.0 / ((double)d[a] + 1e-12));
      }
      if (d[2] > 1e-6f && isfinite(d[0]) && isfinite(d[1])) {
        const float tx = d[0] / d[2], ty = d[1] / d[2];
        t[0] = t[1] = tx; t[2] = t[3] = ty;
      } else {
        bad = 1;
      }
    }
    lo = __reduce_min_sync(0xffffffffu, lo); hi = __reduce_max_sync(0xffffffffu, hi);
    bad = __any_sync(0xffffffffu, bad);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      t[0] = fminf(t[0], __shfl_xor_sync(0xffffffffu, t[0], o)); t[1] = fmaxf(t[1], __shfl_xor_sync(0xffffffffu, t[1], o));
      t[2] = fminf(t[2], __shfl_xor_sync(0xffffffffu, t[2], o)); t[3] = fmaxf(t[3], __shfl_xor_sync(0xffffffffu, t[3], o));
    }
    if (lane == 0) {
      atomicMin(&s_bmin, lo); atomicMax(&s_bmax, hi);
      if (bad) atomicOr(&s_bad, 1);
#pragma unroll
      for (int k = 0; k < 4; ++k) s_t[warp][k] = t[k];
    }
  }
  __syncthreads();
  const int bmin = s_bmin, bmax = s_bmax;
  float4 fr = make_float4(INFINITY, -INFINITY, INFINITY, -INFINITY);
#pragma unroll
  for (int w = 0; w < AABB_RM_THREADS / 32; ++w) {
    fr.x = fminf(fr.x, s_t[w][0]); fr.y = fmaxf(fr.y, s_t[w][1]); fr.z = fminf(fr.z, s_t[w][2]); fr.w = fmaxf(fr.w, s_t[w][3]);
  }
  if (s_bad) fr = make_float4(-INFINITY, INFINITY, -INFINITY, INFINITY);
  int count = 0;
  int64_t pos = (FILL && act) ? (int64_t)start[r] : 0;
  for (int64_t v0 = 0; v0 < V; v0 += AABB_RM_THREADS) {
    const int64_t v = v0 + threadIdx.x;
    bool live = false;
    AabbBox box;
    int vb = 0;
    if (v < V) {
      vb = __ldg(voxel_bid + v);
      if (vb >= bmin && vb <= bmax) { box = aabb_load_box(voxel_bound, v); live = aabb_tile_may_hit(fr, box); }
    }
    const unsigned m = __ballot_sync(0xffffffffu, live);
    if (lane == 0) s_woff[warp] = __popc(m);
    __syncthreads();
    if (threadIdx.x == 0) {                                           // exclusive prefix over the 8 warps; total in s_woff[8]
      int run = 0;
#pragma unroll
      for (int w = 0; w < AABB_RM_THREADS / 32; ++w) { const int c = s_woff[w]; s_woff[w] = run; run += c; }
      s_woff[AABB_RM_THREADS / 32] = run;
    }
    __syncthreads();
    if (live) {
      const int i = s_woff[warp] + __popc(m & ((1u << lane) - 1u));
      s_vid[i] = (int)threadIdx.x; s_vb[i] = vb;
#pragma unroll
      for (int k = 0; k < 3; ++k) { s_box[i * 6 + k] = box.lo[k]; s_box[i * 6 + 3 + k] = box.hi[k]; }
    }
    __syncthreads();
    const int n = s_woff[AABB_RM_THREADS / 32];
    if (act) {
      for (int i = 0; i < n; ++i) {
        if (s_vb[i] != bid) continue;
        AabbBox b;
#pragma unroll
        for (int k = 0; k < 3; ++k) { b.lo[k] = s_box[i * 6 + k]; b.hi[k] = s_box[i * 6 + 3 + k]; }
        float t0, t1;
        if (!aabb_slab(inv[0], inv[1], inv[2], b, t0, t1)) continue;
        if (FILL) {
          pair_vox[pos] = v0 + s_vid[i];
          pair_ray[pos] = r;
          pair_dist[pos] = make_float2(t0, t1);
          ++pos;
        } else {
          ++count;
        }
      }
    }
    __syncthreads();                                                  // the lists are rebuilt by the next voxel tile
  }
  if (!FILL && act) cnt[r] = count;
}

// ---- dense drop-in: mask[V,R] int32, dist[V,R,2] -- every element written once with streaming stores -----------------
// Thread t of a work item owns the 4 consecutive rays 4t .. 4t+3 of the tile: one 16-byte mask store and two 16-byte
// dist stores per thread (a warp writes 512 B + 1 KB contiguous), 16-byte loads of the image ids and reciprocal
// directions.  VEC = false (R not a multiple of 4: rows are not 16-byte aligned) falls back to scalar accesses.
template <bool VEC>
__global__ void __launch_bounds__(AABB_THREADS) k_aabb_dense(const float* __restrict__ inv, const float* __restrict__ voxel_bound,
                                                             const int32_t* __restrict__ ray_bid, const int32_t* __restrict__ voxel_bid,
                                                             const int2* __restrict__ tile_bid, const float4* __restrict__ tile_tan,
                                                             int64_t R, int64_t RB, int64_t M, int* __restrict__ mask,
                                                             float2* __restrict__ dist) {
  for (int64_t w = blockIdx.x; w < M; w += gridDim.x) {
    const int64_t v = w / RB, rb = w - v * RB;
    const int vbid = __ldg(voxel_bid + v);
    const int2 tb = __ldg(tile_bid + rb);
    const bool live = !(vbid < tb.x || vbid > tb.y) && aabb_tile_may_hit(__ldg(tile_tan + rb), aabb_load_box(voxel_bound, v));
    const int64_t r0 = rb * AABB_TILE + 4 * (int64_t)threadIdx.x;
    if (r0 >= R) continue;
    int hit[4] = {0, 0, 0, 0};
    float t0[4] = {0.f, 0.f, 0.f, 0.f}, t1[4] = {0.f, 0.f, 0.f, 0.f};
    if (live) {
      const AabbBox box = aabb_load_box(voxel_bound, v);
      if (VEC) {
        const int4 b4 = *reinterpret_cast<const int4*>(ray_bid + r0);
        const int bid[4] = {b4.x, b4.y, b4.z, b4.w};
        if (bid[0] == vbid || bid[1] == vbid || bid[2] == vbid || bid[3] == vbid) {
          const float4 i0 = *reinterpret_cast<const float4*>(inv + r0 * 3), i1 = *reinterpret_cast<const float4*>(inv + r0 * 3 + 4),
                       i2 = *reinterpret_cast<const float4*>(inv + r0 * 3 + 8);
          const float iv[12] = {i0.x, i0.y, i0.z, i0.w, i1.x, i1.y, i1.z, i1.w, i2.x, i2.y, i2.z, i2.w};
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (bid[j] == vbid && aabb_slab(iv[3 * j], iv[3 * j + 1], iv[3 * j + 2], box, t0[j], t1[j])) hit[j] = 1;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int64_t r = r0 + j;
          if (r < R && ray_bid[r] == vbid && aabb_slab(inv[r * 3 + 0], inv[r * 3 + 1], inv[r * 3 + 2], box, t0[j], t1[j])) hit[j] = 1;
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (!hit[j]) { t0[j] = 0.f; t1[j] = 0.f; }
    }
    int* mp = mask + v * R + r0;
    float2* dp = dist + v * R + r0;
    if (VEC) {
      __stcs(reinterpret_cast<int4*>(mp), make_int4(hit[0], hit[1], hit[2], hit[3]));
      __stcs(reinterpret_cast<float4*>(dp), make_float4(t0[0], t1[0], t0[1], t1[1]));
      __stcs(reinterpret_cast<float4*>(dp) + 1, make_float4(t0[2], t1[2], t0[3], t1[3]));
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (r0 + j < R) { __stcs(mp + j, hit[j]); __stcs(dp + j, make_float2(t0[j], t1[j])); }
    }
  }
}

// ---- point-in-box: dense drop-in mask[V,N], same thread mapping ------------------------------------------------------
template <bool VEC>
__global__ void __launch_bounds__(AABB_THREADS) k_pcl_dense(const float* __restrict__ pcl_pos, const float* __restrict__ voxel_bound,
                                                            const int32_t* __restrict__ pcl_bid, const int32_t* __restrict__ voxel_bid,
                                                            int64_t N, int64_t NB, int64_t M, int* __restrict__ mask) {
  for (int64_t w = blockIdx.x; w < M; w += gridDim.x) {
    const int64_t v = w / NB, nb = w - v * NB;
    const int64_t n0 = nb * AABB_TILE + 4 * (int64_t)threadIdx.x;
    if (n0 >= N) continue;
    const int vbid = __ldg(voxel_bid + v);
    const AabbBox box = aabb_load_box(voxel_bound, v);
    int in[4] = {0, 0, 0, 0};
    if (VEC) {
      const int4 b4 = *reinterpret_cast<const int4*>(pcl_bid + n0);
      const int bid[4] = {b4.x, b4.y, b4.z, b4.w};
      if (bid[0] == vbid || bid[1] == vbid || bid[2] == vbid || bid[3] == vbid) {
        const float4 p0 = *reinterpret_cast<const float4*>(pcl_pos + n0 * 3), p1 = *reinterpret_cast<const float4*>(pcl_pos + n0 * 3 + 4),
                     p2 = *reinterpret_cast<const float4*>(pcl_pos + n0 * 3 + 8);
        const float pv[12] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w, p2.x, p2.y, p2.z, p2.w};
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (bid[j] == vbid && aabb_inside(pv[3 * j], pv[3 * j + 1], pv[3 * j + 2], box)) in[j] = 1;
      }
      __stcs(reinterpret_cast<int4*>(mask + v * N + n0), make_int4(in[0], in[1], in[2], in[3]));
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int64_t n = n0 + j;
        if (n >= N) break;
        const bool inside = pcl_bid[n] == vbid && aabb_inside(pcl_pos[n * 3 + 0], pcl_pos[n * 3 + 1], pcl_pos[n * 3 + 2], box);
        __stcs(mask + v * N + n, inside ? 1 : 0);
      }
    }
  }
}

// label[i] = pcl_mask[pair_vox[i], pair_ray[i]] without the mask (pipeline.py:305-309)
__global__ void k_pcl_pair_label(const float* __restrict__ pcl_pos, const float* __restrict__ voxel_bound,
                                 const int32_t* __restrict__ pcl_bid, const int32_t* __restrict__ voxel_bid, int64_t N, int64_t V,
                                 const int64_t* __restrict__ pair_vox, const int64_t* __restrict__ pair_ray, int64_t P,
                                 float* __restrict__ label) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P) return;
  const int64_t v = pair_vox[i], n = pair_ray[i];
  if (v < 0 || v >= V || n < 0 || n >= N) { label[i] = 0.f; return; }      // index outside the mask: never a hit
  bool in = false;
  if (__ldg(pcl_bid + n) == __ldg(voxel_bid + v))
    in = aabb_inside(__ldg(pcl_pos + n * 3 + 0), __ldg(pcl_pos + n * 3 + 1), __ldg(pcl_pos + n * 3 + 2), aabb_load_box(voxel_bound, v));
  label[i] = in ? 1.f : 0.f;
}

// end_voxel_id[n] = max(end_voxel_id[n], largest v containing point n)  (pipeline.py:939-944).
// One thread per point; voxel boxes go through shared memory in tiles of 256, highest ids first, so a thread stops at
// its first hit and a block stops as soon as every thread is settled.  A voxel tile none of whose image ids falls into
// the block's image-id range is skipped before its boxes are even loaded (points and voxels are image-major in practice,
// so 7 of 8 tiles go that way at batch 8).
#define AABB_VTILE 256
__global__ void __launch_bounds__(AABB_THREADS) k_pcl_end_voxel(const float* __restrict__ pcl_pos, const float* __restrict__ voxel_bound,
                                                                const int32_t* __restrict__ pcl_bid, const int32_t* __restrict__ voxel_bid,
                                                                int64_t N, int64_t V, int64_t* __restrict__ end_voxel_id) {
  __shared__ float s_box[AABB_VTILE * 6];
  __shared__ int s_bid[AABB_VTILE];
  __shared__ int s_bmin, s_bmax;
  const int64_t n = (int64_t)blockIdx.x * AABB_THREADS + threadIdx.x;
  const bool act = n < N;
  float x = 0.f, y = 0.f, z = 0.f;
  int bid = 0;
  int64_t cur = V;                                                   // inactive threads are settled from the start
  if (act) { x = pcl_pos[n * 3 + 0]; y = pcl_pos[n * 3 + 1]; z = pcl_pos[n * 3 + 2]; bid = pcl_bid[n]; cur = end_voxel_id[n]; }
  if (threadIdx.x == 0) { s_bmin = INT_MAX; s_bmax = INT_MIN; }
  __syncthreads();
  {
    const int lo = __reduce_min_sync(0xffffffffu, act ? bid : INT_MAX), hi = __reduce_max_sync(0xffffffffu, act ? bid : INT_MIN);
    if ((threadIdx.x & 31) == 0) { atomicMin(&s_bmin, lo); atomicMax(&s_bmax, hi); }
  }
  __syncthreads();
  const int bmin = s_bmin, bmax = s_bmax;
  const int64_t cur0 = cur;
  bool done = !act;
  for (int64_t hi = V; hi > 0; hi -= AABB_VTILE) {
    const int64_t lo = hi > AABB_VTILE ? hi - AABB_VTILE : 0;
    const int cntv = (int)(hi - lo);
    done = done || (cur >= hi - 1);                                  // nothing above `cur` left in this or lower tiles
    if (__syncthreads_and(done)) break;
    bool mine = false;
    for (int i = threadIdx.x; i < cntv; i += AABB_THREADS) {
      const int vb = voxel_bid[lo + i];
      s_bid[i] = vb;
      mine = mine || (vb >= bmin && vb <= bmax);
    }
    if (!__syncthreads_or(mine)) continue;                           // no voxel of this tile belongs to an image of this block
    for (int i = threadIdx.x; i < cntv * 6; i += AABB_THREADS) s_box[i] = voxel_bound[lo * 6 + i];
    __syncthreads();
    if (!done) {
      for (int i = cntv - 1; i >= 0; --i) {
        const int64_t v = lo + i;
        if (v <= cur) { done = true; break; }
        if (s_bid[i] != bid) continue;
        const float* b = s_box + i * 6;
        if ((x < b[0]) || (x > b[3])) continue;
        if ((y < b[1]) || (y > b[4])) continue;
        if ((z < b[2]) || (z > b[5])) continue;
        cur = v; done = true; break;
      }
    }
  }
  if (act && cur != cur0) end_voxel_id[n] = cur;
}

// ------------------------------------------------------------------------------------------------
// Voxelisation of the valid points: batch_get_occupied_idx (src/utils/point_utils.py:12-76, overlap = False) +
// get_occ_vox_bound (src/models/pipeline.py:162-201).  The reference concatenates (image id, cell) rows and calls
// torch.unique(dim=0) -- a sort of up to 10^4 x B rows.  The grid has only nx*ny*nz (9^3) cells per image, so the sorted
// unique list is an occupancy bitmap + a prefix sum: mark -> scan -> fill, no sort.  Row order of torch.unique is
// lexicographic in (image, cx, cy, cz) = ascending flat cell index.  fp32 operations are the reference's, one by one
// (__f*_rn keeps nvcc from contracting the multiply-adds torch performs as two roundings).
// ------------------------------------------------------------------------------------------------
struct VoxGrid { float xmin[3]; float crop, half_crop; int n[3]; int B; };

__device__ __forceinline__ bool vox_locate(const float* __restrict__ xyz, int64_t i, int bid, const VoxGrid& g, float (&v)[3],
                                           int (&c)[3]) {
  bool ok = bid >= 0 && bid < g.B;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    v[k] = __fsub_rn(xyz[i * 3 + k], g.xmin[k]);                       // point_utils.py:23
    // :43.  True IEEE division = what the reference computes on the CPU (where the voxel goldens were made).  torch's CUDA
    // kernel for `tensor / python_scalar` multiplies by the reciprocal instead; for the shipped grid.res = 8 (crop 0.25, a power
    // of two) both are exact and agree; for a non-power-of-two crop a point within 1 ulp of a cell face may land in the
    // neighbouring cell under the GPU reference -- parity here is defined against exact division.
    const float q = floorf(__fdiv_rn(v[k], g.crop));
    ok = ok && q >= 0.f && q < (float)g.n[k];                          // :59-61 (also rejects NaN)
    c[k] = ok ? (int)q : 0;
  }
  return ok;
}

__global__ void k_vox_mark(const float* __restrict__ xyz, const int64_t* __restrict__ bid, int64_t Np, VoxGrid g,
                           int* __restrict__ cell_flag, int* __restrict__ inside) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Np) return;
  float v[3]; int c[3];
  const int b = (int)bid[i];
  const bool ok = vox_locate(xyz, i, b, g, v, c);
  inside[i] = ok ? 1 : 0;
  if (ok) cell_flag[((int64_t)b * g.n[0] + c[0]) * g.n[1] * g.n[2] + c[1] * g.n[2] + c[2]] = 1;   // benign same-value race
}

__global__ void k_vox_fill_voxels(const int* __restrict__ cell_flag, const int* __restrict__ cell_rank, int64_t ncell_total,
                                  VoxGrid g, int64_t* __restrict__ occ, float* __restrict__ bound) {
  const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= ncell_total || !cell_flag[w]) return;
  const int64_t j = cell_rank[w];
  const int per = g.n[0] * g.n[1] * g.n[2];
  const int b = (int)(w / per), r = (int)(w - (int64_t)b * per);
  const int c[3] = {r / (g.n[1] * g.n[2]), (r / g.n[2]) % g.n[1], r % g.n[2]};
  occ[j * 4 + 0] = b;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    occ[j * 4 + 1 + k] = c[k];
    const float lo = __fadd_rn(g.xmin[k], __fmul_rn((float)c[k], g.crop));      // pipeline.py:186
    bound[j * 6 + k] = lo;
    bound[j * 6 + 3 + k] = __fadd_rn(lo, g.crop);                                 // :187
  }
}

__global__ void k_vox_fill_points(const float* __restrict__ xyz, const int64_t* __restrict__ bid, int64_t Np, VoxGrid g,
                                  const int* __restrict__ inside, const int* __restrict__ pt_rank, const int* __restrict__ cell_rank,
                                  int64_t* __restrict__ revidx, int64_t* __restrict__ pid, float* __restrict__ rel) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Np || !inside[i]) return;
  float v[3]; int c[3];
  const int b = (int)bid[i];
  vox_locate(xyz, i, b, g, v, c);
  const int64_t j = pt_rank[i];
  pid[j] = i;                                                                    // point_utils.py:66
  revidx[j] = cell_rank[((int64_t)b * g.n[0] + c[0]) * g.n[1] * g.n[2] + c[1] * g.n[2] + c[2]];   // :73 inverse index
#pragma unroll
  for (int k = 0; k < 3; ++k)                                                    // :50-51 centre, relative coordinate
    rel[j * 3 + k] = __fsub_rn(v[k], __fadd_rn(__fmul_rn((float)c[k], g.crop), g.half_crop));
}
