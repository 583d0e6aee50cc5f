// temporary stub
#pragma once
#include "lidf_common.cuh"
#define TC_KPE_MAX 112
struct TcBufs { void* p; };
template <typename B> inline TcBufs carve_tc(B& b, int64_t V, int n_dec) { return TcBufs{nullptr}; }
inline int tc_query_forward(const LidfQueryParams*, const TcBufs&, const int*, const float*, const float*, int, int, int,
                            cudaStream_t, int64_t*, char*, size_t, void (*)(int, cudaStream_t)) { return LIDF_ERR_UNSUPPORTED; }
