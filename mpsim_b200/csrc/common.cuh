// Shared device helpers and host-side error plumbing for libmpsim_b200 (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include "../../include/mpsim_b200.h"

typedef float2 cf;   // complex64

__device__ __forceinline__ cf cf_make(float re, float im) { return make_float2(re, im); }
__device__ __forceinline__ cf cf_conj(cf a) { return make_float2(a.x, -a.y); }
__device__ __forceinline__ cf cf_add(cf a, cf b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ cf cf_sub(cf a, cf b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ cf cf_scale(float s, cf a) { return make_float2(s * a.x, s * a.y); }
__device__ __forceinline__ cf cf_mul(cf a, cf b) {
    return make_float2(fmaf(a.x, b.x, -(a.y * b.y)), fmaf(a.x, b.y, a.y * b.x));
}
// acc + a*b   (split-real: 4 FFMA)
__device__ __forceinline__ cf cf_fma(cf a, cf b, cf acc) {
    acc.x = fmaf(a.x, b.x, acc.x);
    acc.x = fmaf(-a.y, b.y, acc.x);
    acc.y = fmaf(a.x, b.y, acc.y);
    acc.y = fmaf(a.y, b.x, acc.y);
    return acc;
}
// acc + conj(a)*b
__device__ __forceinline__ cf cf_fma_conja(cf a, cf b, cf acc) {
    acc.x = fmaf(a.x, b.x, acc.x);
    acc.x = fmaf(a.y, b.y, acc.x);
    acc.y = fmaf(a.x, b.y, acc.y);
    acc.y = fmaf(-a.y, b.x, acc.y);
    return acc;
}
__device__ __forceinline__ float cf_abs2(cf a) { return fmaf(a.x, a.x, a.y * a.y); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---- host side --------------------------------------------------------------------------
void mpsb_set_error(const char* fmt, ...);

#define MPSB_ARG(cond, ...)                     \
    do {                                        \
        if (!(cond)) {                          \
            mpsb_set_error(__VA_ARGS__);        \
            return -1;                          \
        }                                       \
    } while (0)

#define MPSB_CUDA(call)                                                              \
    do {                                                                             \
        cudaError_t e__ = (call);                                                    \
        if (e__ != cudaSuccess) {                                                    \
            mpsb_set_error("%s failed: %s", #call, cudaGetErrorString(e__));         \
            return (int)e__;                                                         \
        }                                                                            \
    } while (0)

#define MPSB_LAUNCH_CHECK(name)                                                      \
    do {                                                                             \
        cudaError_t e__ = cudaGetLastError();                                        \
        if (e__ != cudaSuccess) {                                                    \
            mpsb_set_error("launch of %s failed: %s", name, cudaGetErrorString(e__)); \
            return (int)e__;                                                         \
        }                                                                            \
    } while (0)

// library-owned state (stream pool, pinned read-back slots) is kept per device ordinal and guarded by
// one mutex (api.cu)
#define MPSB_MAX_DEVICES 16
#ifdef __cplusplus
#include <mutex>
std::recursive_mutex& mpsb_lib_mutex();
int mpsb_current_device(int* dev);
const char* mpsb_env(const char* name);      // cached developer switches (api.cu)
#endif

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// internal launchers (defined in the .cu files, used by api.cu)
int launch_cgemm(const cf* A, int64_t a_rs, int64_t a_cs, int conj_a, int64_t a_bs,
                 const cf* B, int64_t b_rs, int64_t b_cs, int conj_b, int64_t b_bs,
                 cf* C, int64_t c_ld, int64_t c_bs, int M, int N, int K, int nbatch,
                 cudaStream_t st);
int launch_theta(const mpsb_gate2_desc* descs, int ndesc, int nbatch, int d, int chiL, int chiM,
                 int chiR, int transpose_out, cf* out, int64_t out_job_stride, cudaStream_t st);
size_t tc_theta_workspace_floats(int njobs, int chiL, int chiM, int chiR);
size_t tc_cgemm_workspace_floats(int njobs, int M, int N, int K);
int launch_theta_tc(const mpsb_gate2_desc* descs, int ndesc, int nbatch, int chiL, int chiM, int chiR,
                    int transpose_out, cf* out, int64_t out_job_stride, float* work, cudaStream_t st);
int launch_cgemm_tc(const cf* A, int64_t a_bs, const cf* B, int64_t b_bs, cf* C, int64_t c_ld, int64_t c_bs,
                    int M, int N, int K, int nbatch, float* work, cudaStream_t st);
int launch_svd_small(const cf* X, int64_t x_job_stride, int njobs, int nv, int L, int k,
                     int left_canonical, const mpsb_gate2_desc* descs, int ndesc, int nbatch,
                     cf* left, int64_t left_stride, cf* right, int64_t right_stride,
                     float* svals, int64_t svals_stride, int32_t* info, cf* zglobal,
                     cudaStream_t st);
size_t svd_large_workspace_elems(int nv, int L);
// one shape group of a multi-stream large-SVD call (svd_large.cu: launch_svd_large_multi)
struct LargeMultiJob {
    cf* X; int64_t x_job_stride; int njobs, nv, L, k, left_canonical;
    const mpsb_gate2_desc* descs; int nbatch; int32_t* info; cf* work; cudaStream_t st; int pin_slot;
};
int launch_svd_large_multi(const LargeMultiJob* jobs, int n);
int svd_large_padded_rows(int nv);
int launch_svd_large(cf* X, int64_t x_job_stride, int njobs, int nv, int L, int k,
                     int left_canonical, const mpsb_gate2_desc* descs, int ndesc, int nbatch,
                     cf* left, int64_t left_stride, cf* right, int64_t right_stride,
                     float* svals, int64_t svals_stride, int32_t* info, cf* work,
                     cudaStream_t st);
