// lidf_common.cuh -- shared device helpers for the lidf_query kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "lidf_query.h"

#define LIDF_MAX_MULTIRES 10
#define LIDF_LEAKY 0.02f

// ------------------------------------------------------------------------------------------------
// Positional encoding, reference src/models/implicit_net.py:14-39:
//   [x(3), sin(2^0 x)(3), cos(2^0 x)(3), sin(2^1 x)(3), cos(2^1 x)(3), ...]   (3 + 6*L values)
// x*2^k is exact in fp32; arguments reach ~450 rad so the accurate sincosf is used, not __sinf.
// ------------------------------------------------------------------------------------------------
__host__ __device__ inline int lidf_pe_dim(int multires, int pos_encode) { return pos_encode ? 3 + 6 * multires : 3; }

template <typename Store>
__device__ __forceinline__ void lidf_pe3(float x0, float x1, float x2, int L, int pos_encode, Store&& st) {
  st(0, x0); st(1, x1); st(2, x2);
  if (!pos_encode) return;
  float f = 1.0f;
  for (int k = 0; k < L; ++k) {
    float s0, c0, s1, c1, s2, c2;
    sincosf(x0 * f, &s0, &c0);
    sincosf(x1 * f, &s1, &c1);
    sincosf(x2 * f, &s2, &c2);
    const int b = 3 + 6 * k;
    st(b + 0, s0); st(b + 1, s1); st(b + 2, s2);
    st(b + 3, c0); st(b + 4, c1); st(b + 5, c2);
    f *= 2.0f;
  }
}

__device__ __forceinline__ float lidf_leaky(float x) { return fmaxf(x, LIDF_LEAKY * x); }

// implicit_net.py:93-96: sigmoid or max(min(x, 0.01x+0.99), 0.01x)
__device__ __forceinline__ float lidf_final_act(float x, int use_sigmoid) {
  if (use_sigmoid) return 1.0f / (1.0f + expf(-x));
  return fmaxf(fminf(x, x * 0.01f + 0.99f), x * 0.01f);
}

__host__ __device__ inline int64_t lidf_align_up(int64_t x, int64_t a) { return (x + a - 1) / a * a; }
