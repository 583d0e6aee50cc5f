// lidf_prep.cuh -- pair regroup (voxel-major -> ray-major CSR), ROIAlign per ray, ray termination.
#pragma once
#include <cuda.h>          // CUtensorMap (types only: cuTensorMapEncodeTiled is fetched through cudaGetDriverEntryPoint)

#include "lidf_common.cuh"

// ------------------------------------------------------------------------------------------------
// Pair regroup.  The reference's pair list is torch.nonzero of a [V,R] mask (pipeline.py:283-285),
// i.e. sorted by voxel then ray.  The decoder kernels and the ray termination want all pairs of a
// ray adjacent, so we build a CSR by ray: ray_start[R+1] and perm[P] (sorted slot -> original pair).
// Outputs are always written back at the ORIGINAL pair index, so results are order-independent.
// ------------------------------------------------------------------------------------------------
// Out-of-range keys are clamped into range (so every pair lands in some segment and nothing downstream indexes out of
// bounds) and reported through *err; every kernel that reads pair_ray / pair_vox / miss_bid again clamps the same way.
__device__ __forceinline__ int64_t lidf_clamp_idx(int64_t v, int64_t n) { return v < 0 ? 0 : (v >= n ? n - 1 : v); }
__global__ void k_count_pairs(const int64_t* __restrict__ pair_ray, int64_t P, int64_t R, int* __restrict__ cnt,
                              int* __restrict__ err, const int64_t* __restrict__ pair_vox = nullptr, int64_t V = 0) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P) return;
  int64_t r = pair_ray[i];
  if (r < 0 || r >= R) { atomicOr(err, 1); r = lidf_clamp_idx(r, R); }
  atomicAdd(cnt + r, 1);
  if (pair_vox) {                                   // range check of the voxel index in the same pass over the pair list
    const int64_t v = pair_vox[i];
    if (v < 0 || v >= V) atomicOr(err, 2);
  }
}
// Pair list already ray-major (pair_ray non-decreasing, LidfQueryParams::pairs_ray_major): the CSR is a binary search per
// ray (ray_start[r] = first i with pair_ray[i] >= r, ray_start[R] = P) and perm is the identity.  k_check_sorted_pairs is the
// pass over the list that keeps the range checks of k_count_pairs and verifies the order (err bit 2).
__global__ void k_ray_start_sorted(const int64_t* __restrict__ pair_ray, int64_t P, int64_t R, int* __restrict__ ray_start) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r > R) return;
  int64_t lo = 0, hi = P;                             // lower bound of r (clamped keys: < 0 counts as ray 0, >= R as R - 1)
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (lidf_clamp_idx(__ldg(pair_ray + mid), R) < r) lo = mid + 1; else hi = mid;
  }
  ray_start[r] = (int)(r == R ? P : lo);
}
__global__ void k_check_sorted_pairs(const int64_t* __restrict__ pair_ray, int64_t P, int64_t R, int* __restrict__ perm,
                                     int* __restrict__ err, const int64_t* __restrict__ pair_vox, int64_t V) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P) return;
  const int64_t r = pair_ray[i];
  int bad = (r < 0 || r >= R) ? 1 : 0;
  if (i > 0 && lidf_clamp_idx(pair_ray[i - 1], R) > lidf_clamp_idx(r, R)) bad |= 4;
  if (pair_vox) { const int64_t v = pair_vox[i]; if (v < 0 || v >= V) bad |= 2; }
  perm[i] = (int)i;
  if (bad) atomicOr(err, bad);
}
// range check of the other index arrays the kernels dereference (pair_vox over P, miss_bid over R)
__global__ void k_validate_indices(const int64_t* __restrict__ pair_vox, int64_t P, int64_t V, const int64_t* __restrict__ bid,
                                   int64_t R, int64_t B, int* __restrict__ err) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  bool bad = false;
  if (pair_vox && i < P) { const int64_t v = pair_vox[i]; bad |= v < 0 || v >= V; }
  if (bid && i < R) { const int64_t b = bid[i]; bad |= b < 0 || b >= B; }
  if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) atomicOr(err, 2);
}
// rays that own at least one pair, appended in arbitrary order (every consumer treats rows independently)
__global__ void k_live_rays(const int* __restrict__ ray_start, int64_t R, int* __restrict__ list, int* __restrict__ count) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = r < R && ray_start[r + 1] > ray_start[r];
  const unsigned m = __ballot_sync(0xffffffffu, live);
  if (!m) return;
  const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
  int base = 0;
  if (lane == leader) base = atomicAdd(count, __popc(m));
  base = __shfl_sync(0xffffffffu, base, leader);
  if (live) list[base + __popc(m & ((1u << lane) - 1u))] = (int)r;
}
__global__ void k_publish_flag(const int* __restrict__ err, int32_t* __restrict__ out) { *out = *err; }

// exclusive scan of int32 counts, 3 kernels (block partials, scan of partials, add back)
#define LIDF_SCAN_BLOCK 1024
#define LIDF_SCAN_ITEMS 4
__global__ void k_scan_partial(const int* __restrict__ in, int64_t n, int* __restrict__ out, int* __restrict__ block_sums) {
  __shared__ int warp_sums[32];
  const int64_t base = ((int64_t)blockIdx.x * LIDF_SCAN_BLOCK + threadIdx.x) * LIDF_SCAN_ITEMS;
  int v[LIDF_SCAN_ITEMS];
  int tsum = 0;
#pragma unroll
  for (int j = 0; j < LIDF_SCAN_ITEMS; ++j) { v[j] = (base + j < n) ? in[base + j] : 0; tsum += v[j]; }
  int inc = tsum;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
  if (lane == 31) warp_sums[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    int w = warp_sums[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += t; }
    warp_sums[lane] = w;
  }
  __syncthreads();
  int excl = inc - tsum + (warp ? warp_sums[warp - 1] : 0);
#pragma unroll
  for (int j = 0; j < LIDF_SCAN_ITEMS; ++j) { if (base + j < n) out[base + j] = excl; excl += v[j]; }
  if (threadIdx.x == LIDF_SCAN_BLOCK - 1) block_sums[blockIdx.x] = excl;
}
// single block: exclusive scan of block_sums in place (nb <= a few thousand)
__global__ void k_scan_blocksums(int* __restrict__ block_sums, int nb) {
  __shared__ int carry;
  __shared__ int warp_sums[32];
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int b0 = 0; b0 < nb; b0 += 1024) {
    int i = b0 + threadIdx.x;
    int v = i < nb ? block_sums[i] : 0;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
    if (lane == 31) warp_sums[warp] = inc;
    __syncthreads();
    if (warp == 0) {
      int w = warp_sums[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += t; }
      warp_sums[lane] = w;
    }
    __syncthreads();
    int excl = inc - v + (warp ? warp_sums[warp - 1] : 0) + carry;
    if (i < nb) block_sums[i] = excl;
    __syncthreads();
    if (threadIdx.x == 1023) carry = excl + v;
    __syncthreads();
  }
}
__global__ void k_scan_add(int* __restrict__ out, int64_t n, const int* __restrict__ block_sums, int* __restrict__ total_slot) {
  const int64_t base = ((int64_t)blockIdx.x * LIDF_SCAN_BLOCK + threadIdx.x) * LIDF_SCAN_ITEMS;
  const int add = block_sums[blockIdx.x];
#pragma unroll
  for (int j = 0; j < LIDF_SCAN_ITEMS; ++j)
    if (base + j < n) out[base + j] += add;
  (void)total_slot;
}

// perm fill: slot = ray_start[ray] + (running cursor); cursor array must be zero on entry
__global__ void k_fill_perm(const int64_t* __restrict__ pair_ray, int64_t P, int64_t R, const int* __restrict__ ray_start,
                            int* __restrict__ cursor, int* __restrict__ perm) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P) return;
  int r = (int)lidf_clamp_idx(pair_ray[i], R);
  int slot = ray_start[r] + atomicAdd(cursor + r, 1);
  perm[slot] = (int)i;
}
// Make each ray's segment ascending in original pair index (the atomics above fill it in arbitrary order),
// so that every floating-point reduction over a ray runs in a reproducible order.
// One warp per ray; segments up to 128 pairs are sorted in registers (bitonic network, shuffles for
// strides < 32), longer ones by odd-even transposition in place.
template <int NPL>
__device__ __forceinline__ void warp_bitonic_sort(int (&v)[NPL], int lane) {
  constexpr int N = 32 * NPL;
#pragma unroll
  for (int k = 2; k <= N; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
      if (j >= 32) {
        const int qj = j >> 5;
#pragma unroll
        for (int q = 0; q < NPL; ++q) {
          if ((q & qj) == 0) {
            const bool up = (((q << 5) + lane) & k) == 0;
            const int a = v[q], b = v[q | qj];
            if ((a > b) == up) { v[q] = b; v[q | qj] = a; }
          }
        }
      } else {
#pragma unroll
        for (int q = 0; q < NPL; ++q) {
          const int o = __shfl_xor_sync(0xffffffffu, v[q], j);
          const bool up = (((q << 5) + lane) & k) == 0;
          const bool lower = (lane & j) == 0;
          v[q] = (lower == up) ? min(v[q], o) : max(v[q], o);
        }
      }
    }
  }
}
template <int NPL>
__device__ __forceinline__ void sort_segment_regs(int* __restrict__ perm, int s, int n, int lane) {
  int v[NPL];
#pragma unroll
  for (int q = 0; q < NPL; ++q) v[q] = (q * 32 + lane < n) ? perm[s + q * 32 + lane] : 0x7fffffff;
  warp_bitonic_sort<NPL>(v, lane);
#pragma unroll
  for (int q = 0; q < NPL; ++q) if (q * 32 + lane < n) perm[s + q * 32 + lane] = v[q];
}
__global__ void k_sort_segments(const int* __restrict__ ray_start, int64_t R, int* __restrict__ perm) {
  const int64_t ray = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (ray >= R) return;
  const int s = ray_start[ray], e = ray_start[ray + 1];
  const int n = e - s;
  if (n <= 1) return;
  if (n <= 32) { sort_segment_regs<1>(perm, s, n, lane); return; }
  if (n <= 64) { sort_segment_regs<2>(perm, s, n, lane); return; }
  if (n <= 128) { sort_segment_regs<4>(perm, s, n, lane); return; }
  // Segments beyond a few thousand pairs do not occur in this workload (a ray meets at most a few hundred voxels); they
  // appear only when out-of-range indices were clamped onto one ray (reported through the index_error flag).  The
  // transposition sort is quadratic, so such a segment is left in fill order: results stay valid, only the order of its
  // floating-point reductions is then not reproducible.
  if (n > 4096) return;
  for (int pass = 0; pass < n; ++pass) {   // long segment (rare): odd-even transposition, n passes
    const int off = pass & 1;
    for (int i = off + 2 * lane; i + 1 < n; i += 64) {
      int a = perm[s + i], b = perm[s + i + 1];
      if (a > b) { perm[s + i] = b; perm[s + i + 1] = a; }
    }
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------------------
// ROIAlign per ray: torchvision.ops.roi_align(feat, boxes, output_size=2, spatial_scale=1, sampling_ratio=-1,
// aligned=True) with boxes = (bid, clamp(x-h), clamp(y-h), clamp(x+h), clamp(y+h)), h = roi_inp_bbox//2
// (reference src/models/pipeline.py:374-389).  The feature depends only on the ray, so it is evaluated once per ray
// (the reference recomputes it for every pair).  Output [R,128] in (c, ph, pw) order (pipeline.py:389).
// One lane per ray (consecutive rays = consecutive x -> coalesced rows), warps loop over channels.
// ------------------------------------------------------------------------------------------------
struct RoiAxis {  // per axis, per output bin: sample taps
  float start, bin;
  int grid;
};

__device__ __forceinline__ void roi_tap(float p, int size, int& lo, int& hi, float& l, float& h, bool& dead) {
  // torchvision bilinear_interpolate, one axis
  dead = (p < -1.0f) || (p > (float)size);
  if (p <= 0.f) p = 0.f;
  lo = (int)p;
  if (lo >= size - 1) { hi = lo = size - 1; p = (float)lo; } else { hi = lo + 1; }
  l = p - (float)lo;
  h = 1.0f - l;
}

#define LIDF_ROI_THREADS 256
// one channel of one ray: the four bins of torchvision's aligned ROIAlign (adaptive sampling grid gh x gw per bin)
__device__ __forceinline__ void roi_channel_general(const float* __restrict__ fc, int H, int W, float sw, float sh, float bw,
                                                    float bh, int gw, int gh, float count, float (&o)[4]) {
#pragma unroll
  for (int ph = 0; ph < 2; ++ph) {
#pragma unroll
    for (int pw = 0; pw < 2; ++pw) {
      float acc = 0.f;
      for (int iy = 0; iy < gh; ++iy) {
        const float y = sh + ph * bh + (iy + 0.5f) * bh / (float)gh;
        int ylo, yhi; float ly, hy; bool ydead;
        roi_tap(y, H, ylo, yhi, ly, hy, ydead);
        for (int ix = 0; ix < gw; ++ix) {
          const float x = sw + pw * bw + (ix + 0.5f) * bw / (float)gw;
          int xlo, xhi; float lx, hx; bool xdead;
          roi_tap(x, W, xlo, xhi, lx, hx, xdead);
          float val = 0.f;
          if (!(ydead || xdead)) {
            const float w1 = hy * hx, w2 = hy * lx, w3 = ly * hx, w4 = ly * lx;
            // same expression order as torchvision: w1*v1 + w2*v2 + w3*v3 + w4*v4; zero-weight taps are skipped
            // (v finite => identical result) which makes interior pixels one load per sample.
            val = w1 * __ldg(fc + (size_t)ylo * W + xlo);
            if (w2 != 0.f) val += w2 * __ldg(fc + (size_t)ylo * W + xhi);
            if (w3 != 0.f) val += w3 * __ldg(fc + (size_t)yhi * W + xlo);
            if (w4 != 0.f) val += w4 * __ldg(fc + (size_t)yhi * W + xhi);
          }
          acc += val;
        }
      }
      o[ph * 2 + pw] = acc / count;
    }
  }
}

// NC channels of one ray at once (planes fc, fc + cstride, ...): the sample positions and bilinear weights depend on the ray
// only, so they are formed once per sample and the NC x (1..4) loads of a sample are independent of each other.  Per channel
// the accumulation order is exactly roi_channel_general's (ph, pw, iy, ix; w1 v1 + w2 v2 + w3 v3 + w4 v4) -> same bits.
template <int NC>
__device__ __forceinline__ void roi_channels_general(const float* __restrict__ fc, size_t cstride, int H, int W, float sw, float sh,
                                                     float bw, float bh, int gw, int gh, float count, float (&o)[NC][4]) {
#pragma unroll
  for (int ph = 0; ph < 2; ++ph) {
#pragma unroll
    for (int pw = 0; pw < 2; ++pw) {
      float acc[NC];
#pragma unroll
      for (int k = 0; k < NC; ++k) acc[k] = 0.f;
      for (int iy = 0; iy < gh; ++iy) {
        const float y = sh + ph * bh + (iy + 0.5f) * bh / (float)gh;
        int ylo, yhi; float ly, hy; bool ydead;
        roi_tap(y, H, ylo, yhi, ly, hy, ydead);
        for (int ix = 0; ix < gw; ++ix) {
          const float x = sw + pw * bw + (ix + 0.5f) * bw / (float)gw;
          int xlo, xhi; float lx, hx; bool xdead;
          roi_tap(x, W, xlo, xhi, lx, hx, xdead);
          if (ydead || xdead) {
#pragma unroll
            for (int k = 0; k < NC; ++k) acc[k] += 0.f;
            continue;
          }
          const float w1 = hy * hx, w2 = hy * lx, w3 = ly * hx, w4 = ly * lx;
          const float* p1 = fc + (size_t)ylo * W + xlo; const float* p2 = fc + (size_t)ylo * W + xhi;
          const float* p3 = fc + (size_t)yhi * W + xlo; const float* p4 = fc + (size_t)yhi * W + xhi;
          float val[NC];
#pragma unroll
          for (int k = 0; k < NC; ++k) val[k] = w1 * __ldg(p1 + k * cstride);
          if (w2 != 0.f) {
#pragma unroll
            for (int k = 0; k < NC; ++k) val[k] += w2 * __ldg(p2 + k * cstride);
          }
          if (w3 != 0.f) {
#pragma unroll
            for (int k = 0; k < NC; ++k) val[k] += w3 * __ldg(p3 + k * cstride);
          }
          if (w4 != 0.f) {
#pragma unroll
            for (int k = 0; k < NC; ++k) val[k] += w4 * __ldg(p4 + k * cstride);
          }
#pragma unroll
          for (int k = 0; k < NC; ++k) acc[k] += val[k];
        }
      }
#pragma unroll
      for (int k = 0; k < NC; ++k) o[k][ph * 2 + pw] = acc[k] / count;
    }
  }
}

// per-ray box set-up shared by the two ROI kernels (boxes are built from integers then .float(), pipeline.py:374-381)
struct RoiBox { float sw, sh, bw, bh, count; int gw, gh; };
__device__ __forceinline__ RoiBox roi_box(int px, int py, int half, int H, int W) {
  const float x1 = (float)min(max(px - half, 0), W - 1), x2 = (float)min(max(px + half, 0), W - 1);
  const float y1 = (float)min(max(py - half, 0), H - 1), y2 = (float)min(max(py + half, 0), H - 1);
  RoiBox r;
  r.sw = x1 - 0.5f; r.sh = y1 - 0.5f;
  const float rw = (x2 - 0.5f) - r.sw, rh = (y2 - 0.5f) - r.sh;
  r.bw = rw / 2.0f; r.bh = rh / 2.0f;
  r.gw = (int)ceilf(rw / 2.0f); r.gh = (int)ceilf(rh / 2.0f);
  r.count = (float)max(r.gh * r.gw, 1);
  return r;
}

// Block = 32 rays (lane) x 8 channel groups (warp).  Results go through a shared tile [32 rays][128 (+4)] so that the global
// stores are whole 512-byte rows (one warp instruction = one ray's row).
// BOX != nullptr (dense ray sets): interior rays (8x8 box not clamped: every sample sits on a pixel centre with weight 1)
// read their four bins from the 4x4 box-sum map, which was accumulated in the same order -> bit-identical to the general
// path.  Rays in the 4-pixel border band need the general path (up to 16 bilinear samples per bin); they are only a few
// per cent of the rays but sit in 10-20 % of the warps, so instead of dragging those warps through the slow path with a
// handful of active lanes they are appended to `border_list` and done by k_roi_align_border with full warps.
// BOX == nullptr: every ray takes the general path here.
__global__ void __launch_bounds__(LIDF_ROI_THREADS)
k_roi_align_rays(const float* __restrict__ feat, const float* __restrict__ box, int B, int H, int W,
                 const int64_t* __restrict__ img_ind, const int64_t* __restrict__ bid, int64_t R, int half,
                 float* __restrict__ out, int* __restrict__ border_list, int* __restrict__ border_count,
                 const int* __restrict__ ray_start = nullptr) {
  __shared__ float s_tile[32][LIDF_RGB_DIM + 4];
  __shared__ unsigned s_deferred;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t ray_raw = (int64_t)blockIdx.x * 32 + lane;
  const int64_t ray = ray_raw < R ? ray_raw : R - 1;                 // tail lanes recompute the last ray (never stored)
  // sparse regime (ray_start given): a ray without pairs is never looked at by the decoders -> no feature for it
  const bool dead = ray_start != nullptr && ray_start[ray + 1] == ray_start[ray];
  if (ray_start != nullptr && __ballot_sync(0xffffffffu, !dead && ray_raw < R) == 0u) return;   // whole warp column idle
  const int px = (int)img_ind[2 * ray], py = (int)img_ind[2 * ray + 1];
  const int b = (int)lidf_clamp_idx(bid[ray], B);
  const bool interior = box != nullptr && half == 4 && px - 4 >= 0 && px + 4 <= W - 1 && py - 4 >= 0 && py + 4 <= H - 1;
  const bool defer = box != nullptr && border_list != nullptr && !interior && !dead;
  const unsigned dmask = __ballot_sync(0xffffffffu, defer && ray_raw < R);
  const unsigned deadmask = __ballot_sync(0xffffffffu, dead);
  if (warp == 0) {
    if (lane == 0) s_deferred = dmask | deadmask;                    // neither group is stored by this kernel
    if (dmask) {                                                     // warp-aggregated append (order is irrelevant)
      const int leader = __ffs(dmask) - 1;
      int base = 0;
      if (lane == leader) base = atomicAdd(border_count, __popc(dmask));
      base = __shfl_sync(0xffffffffu, base, leader);
      if ((dmask >> lane) & 1u) border_list[base + __popc(dmask & ((1u << lane) - 1u))] = (int)ray_raw;
    }
  }
  const size_t plane0 = (size_t)b * LIDF_RGB_CH * H * W;
  if (!defer && !dead) {
    const RoiBox rb = roi_box(px, py, half, H, W);
    for (int c = warp; c < LIDF_RGB_CH; c += LIDF_ROI_THREADS / 32) {
      float o[4];
      if (interior) {
        const float* bc = box + plane0 + (size_t)c * H * W + (size_t)(py - 4) * W + (px - 4);
        // interior: gh = gw = 4, count = 16 -- dividing by a power of two is the same IEEE result as multiplying by 1/16
        o[0] = __ldg(bc) * 0.0625f;
        o[1] = __ldg(bc + 4) * 0.0625f;
        o[2] = __ldg(bc + (size_t)4 * W) * 0.0625f;
        o[3] = __ldg(bc + (size_t)4 * W + 4) * 0.0625f;
      } else {
        roi_channel_general(feat + plane0 + (size_t)c * H * W, H, W, rb.sw, rb.sh, rb.bw, rb.bh, rb.gw, rb.gh, rb.count, o);
      }
      *reinterpret_cast<float4*>(&s_tile[lane][c * 4]) = make_float4(o[0], o[1], o[2], o[3]);
    }
  }
  __syncthreads();
  const int64_t ray0 = (int64_t)blockIdx.x * 32;
  const unsigned deferred = s_deferred;
  for (int r = warp; r < 32; r += LIDF_ROI_THREADS / 32)
    if (ray0 + r < R && !((deferred >> r) & 1u))
      *reinterpret_cast<float4*>(out + (size_t)(ray0 + r) * LIDF_RGB_DIM + 4 * lane) = *reinterpret_cast<const float4*>(&s_tile[r][4 * lane]);
}

// the deferred border rays, 32 per block iteration, general path only (same arithmetic as above -> same bits)
__global__ void __launch_bounds__(LIDF_ROI_THREADS)
k_roi_align_border(const float* __restrict__ feat, int B, int H, int W, const int64_t* __restrict__ img_ind,
                   const int64_t* __restrict__ bid, int half, float* __restrict__ out, const int* __restrict__ border_list,
                   const int* __restrict__ border_count) {
  __shared__ float s_tile[32][LIDF_RGB_DIM + 4];
  __shared__ int s_ray[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n = *border_count;
  for (int chunk = blockIdx.x; chunk * 32 < n; chunk += gridDim.x) {
    const int idx = chunk * 32 + lane;
    const int64_t ray = border_list[idx < n ? idx : n - 1];
    if (warp == 0) s_ray[lane] = idx < n ? (int)ray : -1;
    const int px = (int)img_ind[2 * ray], py = (int)img_ind[2 * ray + 1];
    const int b = (int)lidf_clamp_idx(bid[ray], B);
    const RoiBox rb = roi_box(px, py, half, H, W);
    const size_t plane0 = (size_t)b * LIDF_RGB_CH * H * W;
    {                                                                // this warp's 4 channels (warp, warp + 8, ...) together
      constexpr int NC = LIDF_RGB_CH / (LIDF_ROI_THREADS / 32);
      float o[NC][4];
      roi_channels_general<NC>(feat + plane0 + (size_t)warp * H * W, (size_t)(LIDF_ROI_THREADS / 32) * H * W, H, W, rb.sw, rb.sh,
                               rb.bw, rb.bh, rb.gw, rb.gh, rb.count, o);
#pragma unroll
      for (int k = 0; k < NC; ++k)
        *reinterpret_cast<float4*>(&s_tile[lane][(warp + k * (LIDF_ROI_THREADS / 32)) * 4]) = make_float4(o[k][0], o[k][1], o[k][2], o[k][3]);
    }
    __syncthreads();
    for (int r = warp; r < 32; r += LIDF_ROI_THREADS / 32)
      if (s_ray[r] >= 0)
        *reinterpret_cast<float4*>(out + (size_t)s_ray[r] * LIDF_RGB_DIM + 4 * lane) = *reinterpret_cast<const float4*>(&s_tile[r][4 * lane]);
    __syncthreads();
  }
}

// 4x4 box sums of every feature plane, accumulated exactly like the interior case of roi_channel_general: acc = 0, then
// += f[y + iy][x + ix] for iy (outer), ix (inner).  Positions whose box leaves the plane are never read.
__global__ void k_box4(const float* __restrict__ feat, int64_t n, int H, int W, float* __restrict__ box) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n) return;
  const int x = (int)(idx % W), y = (int)((idx / W) % H);
  if (x > W - 4 || y > H - 4) return;
  const float* f = feat + idx;
  float acc = 0.f;
#pragma unroll
  for (int iy = 0; iy < 4; ++iy)
#pragma unroll
    for (int ix = 0; ix < 4; ++ix) acc += __ldg(f + (size_t)iy * W + ix);
  box[idx] = acc;
}

// The same box sums with the feature tiles staged by TMA: the map is a 3-D tensor (W, H, B * 32 planes); one CTA owns a
// 64 x 32 output tile of one plane, one elected thread issues ONE cp.async.bulk.tensor (SASS: UTMALDG) for the 68 x 35
// input tile (out-of-range elements arrive as zeros and belong to outputs that are never written), the completion is
// counted on an mbarrier, and all 256 threads then read their 16 taps from shared memory in exactly the order of k_box4
// (iy outer, ix inner, sequential adds) -> bit-identical.  Needs W % 4 == 0 (global strides of a tensor map are multiples
// of 16 bytes); other widths take k_box4.
#define BOX_TW 64
#define BOX_TH 32
#define BOX_SW 68                    // 64 + 3 halo, padded to a multiple of 4 floats (TMA inner box = 272 bytes)
#define BOX_SH 35
__global__ void __launch_bounds__(256) k_box4_tma(const __grid_constant__ CUtensorMap tmap, int H, int W, float* __restrict__ box) {
  __shared__ __align__(128) float s_tile[BOX_SH][BOX_SW];
  __shared__ __align__(8) uint64_t s_bar;
  const int x0 = blockIdx.x * BOX_TW, y0 = blockIdx.y * BOX_TH, plane = blockIdx.z;
  const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&s_bar), dst = (uint32_t)__cvta_generic_to_shared(&s_tile[0][0]);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)(BOX_SW * BOX_SH * 4)) : "memory");
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(&tmap)), "r"(x0), "r"(y0), "r"(plane), "r"(bar)
        : "memory");
  }
  {
    uint32_t ok = 0, spins = 0;
    while (!ok) {
      asm volatile(
          "{\n\t.reg .pred p;\n\t"
          "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
          "selp.b32 %0, 1, 0, p;\n\t}"
          : "=r"(ok)
          : "r"(bar), "r"(0u)
          : "memory");
      if (!ok && ++spins > (1u << 24)) __trap();
    }
  }
  const int tx = threadIdx.x & 63, ty0 = threadIdx.x >> 6;               // 64 columns x 4 row phases
  const int x = x0 + tx;
  float* bp = box + (size_t)plane * H * W;
#pragma unroll 2
  for (int ty = ty0; ty < BOX_TH; ty += 4) {
    const int y = y0 + ty;
    if (x > W - 4 || y > H - 4) continue;
    float acc = 0.f;
#pragma unroll
    for (int iy = 0; iy < 4; ++iy)
#pragma unroll
      for (int ix = 0; ix < 4; ++ix) acc += s_tile[ty + iy][tx + ix];
    bp[(size_t)y * W + x] = acc;
  }
}

// ------------------------------------------------------------------------------------------------
// Ray termination (reference src/models/pipeline.py:441-454):
//   soft = scatter_softmax(pred_prob_end, ray)         exp(x - max) / (sum + 1e-12)
//   max_pair_id = scatter_max(soft | pcl_label, ray)   first maximum (lowest pair index) wins, empty ray -> P
//   pred_pos = cat(pair_pred_pos, 0)[max_pair_id]
// A warp owns 4 consecutive rays.  Rays with at most 8 pairs (the reference's real regime: ~2 pairs per ray) are done by the
// warp's four 8-lane groups at once, one pair per lane; longer rays take the whole warp, one after the other.  All reductions
// are shuffle trees in a fixed order; for <= 8 pairs the 8-lane tree adds exactly what the 32-lane tree adds (the upper levels
// only add zeros / compare against -inf), so both paths give the same bits.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void ray_terminate_full_warp(const float* __restrict__ logit, const float* __restrict__ pair_pred_pos,
                                                        const float* __restrict__ label, const int* __restrict__ perm, int64_t P,
                                                        int64_t ray, int s, int e, int lane, float* __restrict__ soft,
                                                        int64_t* __restrict__ max_pair_id, float* __restrict__ pred_pos,
                                                        int* __restrict__ win) {
  float m = -INFINITY;
  for (int i = s + lane; i < e; i += 32) m = fmaxf(m, logit[perm ? perm[i] : i]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  // sum in segment order: lane-strided partials, then a fixed shuffle tree (reproducible)
  float sum = 0.f;
  for (int i = s + lane; i < e; i += 32) sum += expf(logit[perm ? perm[i] : i] - m);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float denom = sum + 1e-12f;
  float best = -INFINITY;
  int best_id = 0x7fffffff;
  for (int i = s + lane; i < e; i += 32) {
    const int id = perm ? perm[i] : i;
    const float sv = expf(logit[id] - m) / denom;
    soft[id] = sv;
    const float key = label ? label[id] : sv;
    if (key > best || (key == best && id < best_id)) { best = key; best_id = id; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, best_id, o);
    if (ob > best || (ob == best && oi < best_id)) { best = ob; best_id = oi; }
  }
  if (lane == 0) { max_pair_id[ray] = (e > s) ? (int64_t)best_id : P; if (win) win[ray] = (e > s) ? best_id : -1; }
  if (lane < 3 && !win) pred_pos[ray * 3 + lane] = (e > s) ? pair_pred_pos[(int64_t)best_id * 3 + lane] : 0.f;
}

// win != nullptr (winner-only mode): pred_pos is not gathered here (the offset decoder writes it for the winners
// afterwards); win [R] receives each ray's arg-max pair, -1 for a ray without pairs.
__global__ void k_ray_terminate(const float* __restrict__ logit, const float* __restrict__ pair_pred_pos,
                                const float* __restrict__ label, const int* __restrict__ ray_start,
                                const int* __restrict__ perm, int64_t P, int64_t R, float* __restrict__ soft,
                                int64_t* __restrict__ max_pair_id, float* __restrict__ pred_pos, int* __restrict__ win = nullptr) {
  const int64_t ray0 = (((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5) * 4;
  const int lane = threadIdx.x & 31, gi = lane >> 3, gl = lane & 7;
  if (ray0 >= R) return;
  const int64_t ray = ray0 + gi;
  const bool have = ray < R;
  const int s = have ? ray_start[ray] : 0, e = have ? ray_start[ray + 1] : 0;
  const bool small = have && e - s <= 8;
  {  // ---- the four 8-lane groups: rays with <= 8 pairs, lane gl holds pair s + gl
    const bool act = small && s + gl < e;
    const int id = act ? (perm ? perm[s + gl] : s + gl) : 0x7fffffff;
    const float x = act ? logit[id] : -INFINITY;
    float m = x;
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    const float ex = act ? expf(x - m) : 0.f;
    float sum = ex;
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float denom = sum + 1e-12f;
    float best = -INFINITY;
    int best_id = 0x7fffffff;
    if (act) {
      const float sv = ex / denom;
      soft[id] = sv;
      best = label ? label[id] : sv;
      best_id = id;
    }
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, best_id, o);
      if (ob > best || (ob == best && oi < best_id)) { best = ob; best_id = oi; }
    }
    if (small) {
      if (gl == 0) { max_pair_id[ray] = (e > s) ? (int64_t)best_id : P; if (win) win[ray] = (e > s) ? best_id : -1; }
      if (gl < 3 && !win) pred_pos[ray * 3 + gl] = (e > s) ? pair_pred_pos[(int64_t)best_id * 3 + gl] : 0.f;
    }
  }
  // ---- rays with more than 8 pairs: the whole warp, one ray at a time
  unsigned longm = __ballot_sync(0xffffffffu, have && !small && gl == 0);
  while (longm) {
    const int src = __ffs(longm) - 1;
    longm &= longm - 1;
    const int ls = __shfl_sync(0xffffffffu, s, src), le = __shfl_sync(0xffffffffu, e, src);
    ray_terminate_full_warp(logit, pair_pred_pos, label, perm, P, ray0 + (src >> 3), ls, le, lane, soft, max_pair_id, pred_pos, win);
  }
}

// ------------------------------------------------------------------------------------------------
// Ray-wise loss statistics on the path's outputs: the torch_scatter calls of LIDF.compute_loss that act on per-pair
// data keyed by ray (pipeline.py:482-486 scatter_log_softmax cross-entropy, :553-557 scatter_max labels / accuracy) plus
// the two per-ray position errors (:472 L1, :560-567 masked L2).  One warp per ray over the CSR built by build_csr;
// per-ray partials are reduced in a fixed order (reproducible), in double precision.
//   log_softmax[i] = (x_i - max_ray) - log(sum_ray exp(x - max_ray) + 1e-12)          (torch_scatter 2.0.x, eps 1e-12)
//   pred_label[r]  = first arg-max of pred_prob_end_softmax over the ray's pairs, P for a ray without pairs
//   gt_label[r]    = first arg-max of pcl_label over the ray's pairs (first pair of the ray when no pair is labelled)
// ray_part[r] = {sum of -log_softmax over labelled pairs, #labelled pairs, pred_label == gt_label, sum |pred - gt| (3 terms),
//                ||pred - gt||_2 * mask, mask}  with mask = (sum |gt| != 0)
// ------------------------------------------------------------------------------------------------
#define LIDF_LOSS_NSTAT 6
__global__ void k_ray_loss(const float* __restrict__ logit, const float* __restrict__ soft, const float* __restrict__ label,
                           const int* __restrict__ ray_start, const int* __restrict__ perm, int64_t P, int64_t R,
                           const float* __restrict__ pred_pos, const float* __restrict__ gt_pos,
                           float* __restrict__ log_softmax, int64_t* __restrict__ pred_label, int64_t* __restrict__ gt_label,
                           float* __restrict__ ray_part) {
  const int64_t ray = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (ray >= R) return;
  const int s = ray_start[ray], e = ray_start[ray + 1];
  float m = -INFINITY;
  for (int i = s + lane; i < e; i += 32) m = fmaxf(m, logit[perm[i]]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  float sum = 0.f;
  for (int i = s + lane; i < e; i += 32) sum += expf(logit[perm[i]] - m);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float lognorm = logf(sum + 1e-12f);
  float ce = 0.f, nlab = 0.f;
  float pbest = -INFINITY, gbest = -INFINITY;
  int pid = 0x7fffffff, gid = 0x7fffffff;
  for (int i = s + lane; i < e; i += 32) {
    const int id = perm[i];
    const float l = (logit[id] - m) - lognorm;
    log_softmax[id] = l;
    const float lab = label[id];
    if (lab != 0.f) { ce -= l; nlab += 1.f; }
    const float sv = soft[id];
    if (sv > pbest || (sv == pbest && id < pid)) { pbest = sv; pid = id; }
    if (lab > gbest || (lab == gbest && id < gid)) { gbest = lab; gid = id; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    ce += __shfl_xor_sync(0xffffffffu, ce, o);
    nlab += __shfl_xor_sync(0xffffffffu, nlab, o);
    const float ob = __shfl_xor_sync(0xffffffffu, pbest, o); const int oi = __shfl_xor_sync(0xffffffffu, pid, o);
    if (ob > pbest || (ob == pbest && oi < pid)) { pbest = ob; pid = oi; }
    const float gb = __shfl_xor_sync(0xffffffffu, gbest, o); const int gi = __shfl_xor_sync(0xffffffffu, gid, o);
    if (gb > gbest || (gb == gbest && gi < gid)) { gbest = gb; gid = gi; }
  }
  if (lane == 0) {
    const int64_t pl = (e > s) ? (int64_t)pid : P, gl = (e > s) ? (int64_t)gid : P;
    pred_label[ray] = pl; gt_label[ray] = gl;
    float* o = ray_part + ray * LIDF_LOSS_NSTAT;
    o[0] = ce; o[1] = nlab; o[2] = pl == gl ? 1.f : 0.f;
    float l1 = 0.f, l2 = 0.f, mask = 0.f;
    if (gt_pos) {
      float az = 0.f, sq = 0.f;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const float g = gt_pos[ray * 3 + k], d = pred_pos[ray * 3 + k] - g;
        l1 += fabsf(d); sq += d * d; az += fabsf(g);
      }
      mask = az != 0.f ? 1.f : 0.f;
      l2 = sqrtf(sq) * mask;
    }
    o[3] = l1; o[4] = l2; o[5] = mask;
  }
}
// fixed-order reduction of ray_part[R][6] -> stats[6] (double): each block sums a contiguous slab of rays (thread-strided
// partials, shuffle tree), then the last block to finish adds the block partials in block order.
__global__ void k_ray_loss_reduce(const float* __restrict__ ray_part, int64_t R, double* __restrict__ block_part,
                                  unsigned* __restrict__ done, double* __restrict__ stats) {
  __shared__ double sh[8][LIDF_LOSS_NSTAT];
  __shared__ bool last;
  const int64_t per = (R + gridDim.x - 1) / gridDim.x, r0 = (int64_t)blockIdx.x * per, r1 = r0 + per < R ? r0 + per : R;
  double acc[LIDF_LOSS_NSTAT];
#pragma unroll
  for (int k = 0; k < LIDF_LOSS_NSTAT; ++k) acc[k] = 0.0;
  for (int64_t r = r0 + threadIdx.x; r < r1; r += blockDim.x)
#pragma unroll
    for (int k = 0; k < LIDF_LOSS_NSTAT; ++k) acc[k] += (double)ray_part[r * LIDF_LOSS_NSTAT + k];
#pragma unroll
  for (int k = 0; k < LIDF_LOSS_NSTAT; ++k)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], o);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0)
#pragma unroll
    for (int k = 0; k < LIDF_LOSS_NSTAT; ++k) sh[warp][k] = acc[k];
  __syncthreads();
  if (threadIdx.x < LIDF_LOSS_NSTAT) {
    double t = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += sh[w][threadIdx.x];
    block_part[(size_t)blockIdx.x * LIDF_LOSS_NSTAT + threadIdx.x] = t;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) last = atomicAdd(done, 1u) == gridDim.x - 1;
  __syncthreads();
  if (last && threadIdx.x < LIDF_LOSS_NSTAT) {
    double t = 0.0;
    for (unsigned b = 0; b < gridDim.x; ++b) t += ((volatile double*)block_part)[(size_t)b * LIDF_LOSS_NSTAT + threadIdx.x];
    stats[threadIdx.x] = t;
    if (threadIdx.x == 0) *done = 0;
  }
}

// ------------------------------------------------------------------------------------------------
// Image-space loss terms of LIDF.compute_loss (pipeline.py:494-541): surface-normal loss, angle error and smoothness loss
// at the miss pixels of the point image with pred_pos / gt_pos scattered in (point_utils.gradient / get_surface_normal,
// src/utils/point_utils.py:208-235).  The reference builds two full [B,3,H,W] normal images and three gradient images
// and then gathers R pixels from each; here one thread per miss pixel evaluates its two normals from the 3 neighbouring
// points directly.  ray_part[r] = {(1 - cos)/2, acos(clamp(cos)), |dx|^2, |dy|^2, 0, 0} -> k_ray_loss_reduce.
// ------------------------------------------------------------------------------------------------
__global__ void k_img_scatter(const float* __restrict__ pred_pos, const float* __restrict__ gt_pos, const int64_t* __restrict__ bid,
                              const int64_t* __restrict__ flat, int64_t R, int64_t HW, float* __restrict__ pred_pcl,
                              float* __restrict__ gt_pcl) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  const int64_t o = (bid[r] * HW + flat[r]) * 3;
#pragma unroll
  for (int k = 0; k < 3; ++k) { pred_pcl[o + k] = pred_pos[r * 3 + k]; gt_pcl[o + k] = gt_pos[r * 3 + k]; }
}

// forward differences (zero in the last column / row), cross product, n / (|n| + 1e-8)
__device__ __forceinline__ void img_normal_at(const float* __restrict__ pcl, int64_t b, int y, int x, int H, int W, float (&n)[3],
                                              float (&dx)[3], float (&dy)[3]) {
  const float* p = pcl + ((b * H + y) * (int64_t)W + x) * 3;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    dx[k] = x == W - 1 ? 0.f : __fsub_rn(p[3 + k], p[k]);
    dy[k] = y == H - 1 ? 0.f : __fsub_rn(p[(int64_t)W * 3 + k], p[k]);
  }
  // torch.cross evaluates a_i b_j - a_j b_i as fma(a_i, b_j, -round(a_j b_i)) (compiler contraction, CPU and CUDA builds
  // alike).  It matters: where dx == dy (both neighbours are rays without a pair, pred_pos = 0) the result is the rounding
  // error of one product instead of 0, and the normalisation below blows it up to a near-unit vector.
  n[0] = fmaf(dx[1], dy[2], -__fmul_rn(dx[2], dy[1]));
  n[1] = fmaf(dx[2], dy[0], -__fmul_rn(dx[0], dy[2]));
  n[2] = fmaf(dx[0], dy[1], -__fmul_rn(dx[1], dy[0]));
  const float len = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(n[0], n[0]), __fmul_rn(n[1], n[1])), __fmul_rn(n[2], n[2])));
  const float d = __fadd_rn(len, 1e-8f);
#pragma unroll
  for (int k = 0; k < 3; ++k) n[k] = __fdiv_rn(n[k], d);
}

__global__ void k_img_loss(const float* __restrict__ pred_pcl, const float* __restrict__ gt_pcl, const int64_t* __restrict__ bid,
                           const int64_t* __restrict__ flat, int64_t R, int H, int W, float* __restrict__ ray_part) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  const int64_t b = bid[r], f = flat[r];
  const int y = (int)(f / W), x = (int)(f - (int64_t)y * W);
  float pn[3], gn[3], dx[3], dy[3], t0[3], t1[3];
  img_normal_at(gt_pcl, b, y, x, H, W, gn, t0, t1);
  img_normal_at(pred_pcl, b, y, x, H, W, pn, dx, dy);
  // F.cosine_similarity (torch 2.x): sum((a / max(|a|, eps)) * (b / max(|b|, eps))), eps = 1e-8
  const float na = fmaxf(sqrtf(pn[0] * pn[0] + pn[1] * pn[1] + pn[2] * pn[2]), 1e-8f);
  const float nb = fmaxf(sqrtf(gn[0] * gn[0] + gn[1] * gn[1] + gn[2] * gn[2]), 1e-8f);
  float cosv = 0.f;
#pragma unroll
  for (int k = 0; k < 3; ++k) cosv += (pn[k] / na) * (gn[k] / nb);
  float* o = ray_part + r * LIDF_LOSS_NSTAT;
  o[0] = (1.f - cosv) / 2.f;
  o[1] = acosf(fminf(fmaxf(cosv, -1.f), 1.f));
  o[2] = dx[0] * dx[0] + dx[1] * dx[1] + dx[2] * dx[2];
  o[3] = dy[0] * dy[0] + dy[1] * dy[1] + dy[2] * dy[2];
  o[4] = 0.f; o[5] = 0.f;
}

// optional: the full normal image [B,3,H,W] (data_dict['pred_surf_norm_img'] / ['gt_surf_norm_img'], pipeline.py:609-611)
__global__ void k_img_normals(const float* __restrict__ pcl, int B, int H, int W, float* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t HW = (int64_t)H * W;
  if (i >= B * HW) return;
  const int64_t b = i / HW, f = i - b * HW;
  const int y = (int)(f / W), x = (int)(f - (int64_t)y * W);
  float n[3], dx[3], dy[3];
  img_normal_at(pcl, b, y, x, H, W, n, dx, dy);
#pragma unroll
  for (int k = 0; k < 3; ++k) out[(b * 3 + k) * HW + f] = n[k];
}

// ------------------------------------------------------------------------------------------------
// Depth metrics of LIDF.compute_loss for exp_type != 'train' (reference src/models/pipeline.py:570-618): a1/a2/a3, rmse,
// rmse_log, log10 (sic: natural log, :607), abs_rel, mae, sq_rel over a set of (pred, gt) depth pairs.
//   rays  (bs != 1, :571-575): the rays whose gt_pos is not all-zero, depth = z of pred_pos / gt_pos
//   image (bs == 1, :576-603): the reference pulls gt depth, the corrupt mask and the predicted depth image to the host and
//     resamples them to 256x144 with cv2.resize(INTER_NEAREST); the same nearest-neighbour pick is done here on the device:
//     dst(y, x) <- src(min(floor(y * H / 144), H - 1), min(floor(x * W / 256), W - 1)) (OpenCV resizeNN, double arithmetic);
//     valid = gt_depth > 0 (after nan / inf -> 0) and corrupt_mask != 0.
// Per-element terms are written as two [n][6] slabs for the fixed-order reduction k_ray_loss_reduce:
//   A = {valid, thresh < 1.05, < 1.10, < 1.25, (gt-pred)^2, (log gt - log pred)^2}
//   B = {|log gt - log pred|, |gt-pred| / gt, |gt-pred|, (gt-pred)^2 / gt, 0, 0}
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void depth_metric_terms(float pred, float gt, bool valid, float* A, float* B) {
  if (!valid) {
#pragma unroll
    for (int k = 0; k < 6; ++k) { A[k] = 0.f; B[k] = 0.f; }
    return;
  }
  const float thresh = fmaxf(gt / pred, pred / gt);
  const float d = gt - pred;
  const float lg = logf(fminf(fmaxf(gt, 1e-6f), 1e6f)) - logf(fminf(fmaxf(pred, 1e-6f), 1e6f));
  A[0] = 1.f; A[1] = thresh < 1.05f ? 1.f : 0.f; A[2] = thresh < 1.10f ? 1.f : 0.f; A[3] = thresh < 1.25f ? 1.f : 0.f;
  A[4] = d * d; A[5] = lg * lg;
  B[0] = fabsf(lg); B[1] = fabsf(d) / gt; B[2] = fabsf(d); B[3] = d * d / gt; B[4] = 0.f; B[5] = 0.f;
}
__global__ void k_depth_terms_rays(const float* __restrict__ pred_pos, const float* __restrict__ gt_pos, int64_t R,
                                   float* __restrict__ partA, float* __restrict__ partB) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  const float gx = gt_pos[3 * r], gy = gt_pos[3 * r + 1], gz = gt_pos[3 * r + 2];
  const bool valid = (fabsf(gx) + fabsf(gy) + fabsf(gz)) != 0.f;             // zero_mask, :560-568
  depth_metric_terms(pred_pos[3 * r + 2], gz, valid, partA + r * 6, partB + r * 6);
}
// z channel of xyz_corrupt_flat with pred_pos scattered in at the miss pixels (:590-592)
__global__ void k_depth_pred_image(const float* __restrict__ xyz_corrupt_flat, int64_t HW, float* __restrict__ predz) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < HW) predz[i] = xyz_corrupt_flat[3 * i + 2];
}
__global__ void k_depth_pred_scatter(const float* __restrict__ pred_pos, const int64_t* __restrict__ flat, int64_t R, int64_t HW,
                                     float* __restrict__ predz) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  const int64_t f = flat[r];
  if (f >= 0 && f < HW) predz[f] = pred_pos[3 * r + 2];
}
#define LIDF_METRIC_W 256
#define LIDF_METRIC_H 144
__global__ void k_depth_terms_image(const float* __restrict__ xyz_flat, const float* __restrict__ corrupt_mask,
                                    const float* __restrict__ predz, int H, int W, float* __restrict__ partA,
                                    float* __restrict__ partB) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= LIDF_METRIC_W * LIDF_METRIC_H) return;
  const int y = i / LIDF_METRIC_W, x = i % LIDF_METRIC_W;
  const int sy = min((int)floor((double)y * ((double)H / (double)LIDF_METRIC_H)), H - 1);
  const int sx = min((int)floor((double)x * ((double)W / (double)LIDF_METRIC_W)), W - 1);
  const size_t s = (size_t)sy * W + sx;
  float gt = xyz_flat[3 * s + 2];
  if (isnan(gt) || isinf(gt)) gt = 0.f;
  const bool valid = gt > 0.f && ((unsigned char)corrupt_mask[s]) != 0;       // seg_mask.astype(np.uint8), :584-588
  depth_metric_terms(predz[s], gt, valid, partA + (size_t)i * 6, partB + (size_t)i * 6);
}
