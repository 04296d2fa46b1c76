// Micro-benchmark: sweep-engine variants for svd_small.cu (one-sided Jacobi on the rows of a 128 x 128
// complex64 matrix resident in shared memory, one CTA per matrix).  Same rotation formulas and the same
// circle-method block tournament in every variant; what changes is the register/thread blocking and the
// instruction mix:
//   V0  512 threads, 4-row blocks (8 rows per warp), interleaved complex, FFMA      (= production, round 1)
//   V2  512 threads, 4-row blocks, PLANAR re/im rows, packed fma.rn.f32x2 (FFMA2), LDS.128
//   V3 1024 threads, 2-row blocks (4 rows per warp), planar + FFMA2, block-wide barrier per round
//   V4 1024 threads, 2-row blocks, planar + FFMA2, grouped named barriers (edges e, e+15, e+30 share an id)
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o build/jacobi_variants scripts/micro/jacobi_variants.cu
//   ./build/jacobi_variants [nmat=148] [max_sweeps=30]
// Prints per variant: ms per launch, sweeps, cycles per sweep of CTA 0, and the check of matrix 0 against a
// float64 Jacobi on the host.  Not part of the product library.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <complex>
#include <algorithm>
#include <cuda_runtime.h>

typedef float2 cf;
constexpr int N = 128;
constexpr float TOL2 = 3e-6f * 3e-6f, BIG2 = 1e-8f;

__device__ __forceinline__ void rot_params(float a, float b, float gr, float gi, float g2,
                                           float& c, float& sr, float& si, float& tg) {
    float rg = rsqrtf(g2);
    float zeta = (a - b) * (0.5f * rg);
    float az = fminf(fabsf(zeta), 1e18f);
    float z2 = fmaf(az, az, 1.0f);
    float t = copysignf(__fdividef(1.0f, az + z2 * rsqrtf(z2)), zeta);
    float ct = (t * rsqrtf(fmaf(t, t, 1.0f))) * rg;
    sr = ct * gr;
    si = ct * gi;
    float h = fmaf(sr, sr, si * si);
    if (h < 0.0625f) {
        float poly = fmaf(h, fmaf(h, fmaf(h, fmaf(h, 0.02734375f, 0.0390625f), 0.0625f), 0.125f), 0.5f);
        c = fmaf(-h, poly, 1.0f);
    } else {
        c = sqrtf(fmaf(-sr, sr, fmaf(-si, si, 1.0f)));
    }
    tg = t * (g2 * rg);
}

__device__ __forceinline__ void named_barrier_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ------------------------------------------------------------------------------------------------
// V0: production structure (interleaved complex, FFMA)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void rot_apply(float c, float sr, float si, cf& p, cf& q) {
    cf np_, nq_;
    np_.x = fmaf(c, p.x, fmaf(sr, q.x, -(si * q.y)));
    np_.y = fmaf(c, p.y, fmaf(sr, q.y, si * q.x));
    nq_.x = fmaf(c, q.x, -fmaf(sr, p.x, si * p.y));
    nq_.y = fmaf(c, q.y, fmaf(si, p.x, -(sr * p.y)));
    p = np_;
    q = nq_;
}

// transposed reduction of 4 values: afterwards lane l holds the full sum of value (l >> 3)
__device__ __forceinline__ float reduce4(const float (&g)[4], int lane) {
    const bool h16 = (lane & 16) != 0, h8 = (lane & 8) != 0;
    float k0 = h16 ? g[2] : g[0], k1 = h16 ? g[3] : g[1];
    float s0 = h16 ? g[0] : g[2], s1 = h16 ? g[1] : g[3];
    k0 += __shfl_xor_sync(0xffffffffu, s0, 16);
    k1 += __shfl_xor_sync(0xffffffffu, s1, 16);
    float k = h8 ? k1 : k0, sd = h8 ? k0 : k1;
    k += __shfl_xor_sync(0xffffffffu, sd, 8);
    k += __shfl_xor_sync(0xffffffffu, k, 4);
    k += __shfl_xor_sync(0xffffffffu, k, 2);
    k += __shfl_xor_sync(0xffffffffu, k, 1);
    return k;
}

#define TICK(k) do { if (PROF) { const unsigned now_ = (unsigned)clock(); acc[k] += now_ - tlast; tlast = now_; } } while (0)

template <bool PROF, int A0, int A1, int A2, int A3, int B0, int B1, int B2, int B3>
__device__ __forceinline__ int sub_round_v0(cf (&y)[8][4], float (&a)[8], int lane, bool& big, unsigned (&acc)[8], unsigned& tlast) {
    constexpr int PA[4] = {A0, A1, A2, A3}, PB[4] = {B0, B1, B2, B3};
    float gr[4], gi[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float r = 0.f, m = 0.f;
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            cf p = y[PA[i]][t], q = y[PB[i]][t];
            r = fmaf(p.x, q.x, r); r = fmaf(p.y, q.y, r);
            m = fmaf(p.y, q.x, m); m = fmaf(-p.x, q.y, m);
        }
        gr[i] = r; gi[i] = m;
    }
    if (PROF) { asm volatile("" :: "f"(gr[0]), "f"(gr[1]), "f"(gr[2]), "f"(gr[3]), "f"(gi[0]), "f"(gi[1]), "f"(gi[2]), "f"(gi[3])); }
    TICK(0);
    const float mgr = reduce4(gr, lane), mgi = reduce4(gi, lane);
    if (PROF) { asm volatile("" :: "f"(mgr), "f"(mgi)); }
    TICK(1);
    const int sel = lane >> 3;
    float ap = a[PA[0]], aq = a[PB[0]];
#pragma unroll
    for (int i = 1; i < 4; ++i) if (sel == i) { ap = a[PA[i]]; aq = a[PB[i]]; }
    const float g2 = fmaf(mgr, mgr, mgi * mgi), apq = ap * aq;
    float c = 1.f, sr = 0.f, si = 0.f, tg = 0.f;
    const bool dorot = (g2 > TOL2 * apq) && (g2 > 1e-30f);
    big = big || (dorot && g2 > BIG2 * apq);
    if (dorot) rot_params(ap, aq, mgr, mgi, g2, c, sr, si, tg);
    if (PROF) { asm volatile("" :: "f"(c), "f"(sr), "f"(si), "f"(tg)); }
    TICK(2);
    const unsigned bal = __ballot_sync(0xffffffffu, dorot);
    const unsigned flags = (bal & 1u) | ((bal >> 7) & 2u) | ((bal >> 14) & 4u) | ((bal >> 21) & 8u);
    if (flags == 0u) return 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        if (flags & (1u << i)) {
            const float ci = __shfl_sync(0xffffffffu, c, 8 * i);
            const float sri = __shfl_sync(0xffffffffu, sr, 8 * i);
            const float sii = __shfl_sync(0xffffffffu, si, 8 * i);
            const float tgi = __shfl_sync(0xffffffffu, tg, 8 * i);
#pragma unroll
            for (int t = 0; t < 4; ++t) rot_apply(ci, sri, sii, y[PA[i]][t], y[PB[i]][t]);
            a[PA[i]] = fmaxf(a[PA[i]] + tgi, 0.f);
            a[PB[i]] = fmaxf(a[PB[i]] - tgi, 0.f);
        }
    }
    if (PROF) {
#pragma unroll
        for (int i = 0; i < 8; ++i) asm volatile("" :: "f"(y[i][0].x), "f"(y[i][3].y));
    }
    TICK(3);
    return __popc(flags);
}

// pairwise named barriers between neighbouring groups (production scheme, ngroups <= 16)
__device__ __forceinline__ void neighbour_sync(int g, int ngroups) {
    const bool has_r = g + 1 < ngroups, has_l = g > 0;
    if (g & 1) {
        if (has_l) named_barrier_sync(g, 64);
        if (has_r) named_barrier_sync(g + 1, 64);
    } else {
        if (has_r) named_barrier_sync(g + 1, 64);
        if (has_l) named_barrier_sync(g, 64);
    }
}

template <bool PROF>
__global__ void __launch_bounds__(512, 1) jacobi_v0(const cf* X, cf* Yout, int* info, long long* clk, int max_sweeps) {
    unsigned acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    unsigned tlast = 0;
    extern __shared__ float4 smem_raw[];
    constexpr int LS = N + 4, NT = 512, NW = 16;
    cf* Ys = (cf*)smem_raw;
    float* nrm = (float*)(Ys + N * LS);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const cf* Xj = X + (size_t)blockIdx.x * N * N;
    for (int e = tid; e < N * N; e += NT) Ys[(e / N) * LS + (e % N)] = Xj[e];
    __syncthreads();
    constexpr int nb = N / 4, mcirc = nb - 1, nrounds = nb - 1, ngroups = nb / 2;
    int sweeps = 0, status = 1;
    long long t0 = clock64();
    for (int sweep = 0; sweep < max_sweeps; ++sweep) {
        for (int i = warp; i < N; i += NW) {
            float s2 = 0.f;
#pragma unroll
            for (int t = 0; t < 4; ++t) { cf v = Ys[i * LS + lane + 32 * t]; s2 = fmaf(v.x, v.x, fmaf(v.y, v.y, s2)); }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s2 += __shfl_xor_sync(0xffffffffu, s2, o);
            if (lane == 0) nrm[i] = s2;
        }
        __syncthreads();
        bool big = false;
        for (int r = 0; r < nrounds; ++r) {
            const int g = warp;
            int I, J;
            if (g == 0) { I = mcirc; J = r; }
            else { I = (r + g) % mcirc; J = (r - g + mcirc) % mcirc; }
            cf* rowA = Ys + (4 * I) * LS + lane;
            cf* rowB = Ys + (4 * J) * LS + lane;
            cf v[8][4];
            float a[8];
            if (PROF) tlast = (unsigned)clock();
#pragma unroll
            for (int i = 0; i < 4; ++i) {
#pragma unroll
                for (int t = 0; t < 4; ++t) { v[i][t] = rowA[i * LS + 32 * t]; v[4 + i][t] = rowB[i * LS + 32 * t]; }
                a[i] = nrm[4 * I + i];
                a[4 + i] = nrm[4 * J + i];
            }
            if (PROF) {
#pragma unroll
                for (int i = 0; i < 8; ++i) asm volatile("" :: "f"(v[i][0].x), "f"(v[i][3].y), "f"(a[i]));
            }
            TICK(4);
            int nrot = 0;
            if (r == 0) {
                nrot += sub_round_v0<PROF, 0, 2, 4, 6, 1, 3, 5, 7>(v, a, lane, big, acc, tlast);
                nrot += sub_round_v0<PROF, 0, 1, 4, 5, 2, 3, 6, 7>(v, a, lane, big, acc, tlast);
                nrot += sub_round_v0<PROF, 0, 1, 4, 5, 3, 2, 7, 6>(v, a, lane, big, acc, tlast);
            }
            nrot += sub_round_v0<PROF, 0, 1, 2, 3, 4, 5, 6, 7>(v, a, lane, big, acc, tlast);
            nrot += sub_round_v0<PROF, 0, 1, 2, 3, 5, 6, 7, 4>(v, a, lane, big, acc, tlast);
            nrot += sub_round_v0<PROF, 0, 1, 2, 3, 6, 7, 4, 5>(v, a, lane, big, acc, tlast);
            nrot += sub_round_v0<PROF, 0, 1, 2, 3, 7, 4, 5, 6>(v, a, lane, big, acc, tlast);
            if (nrot) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
#pragma unroll
                    for (int t = 0; t < 4; ++t) { rowA[i * LS + 32 * t] = v[i][t]; rowB[i * LS + 32 * t] = v[4 + i][t]; }
                }
                float am = a[0];
#pragma unroll
                for (int i = 1; i < 8; ++i) if (lane == i) am = a[i];
                if (lane < 8) nrm[lane < 4 ? 4 * I + lane : 4 * J + lane - 4] = am;
            }
            TICK(5);
            if (r + 1 < nrounds) neighbour_sync(g, ngroups);
            TICK(6);
        }
        sweeps = sweep + 1;
        if (!__syncthreads_or(big ? 1 : 0)) { status = 0; break; }
    }
    long long t1 = clock64();
    cf* Yj = Yout + (size_t)blockIdx.x * N * N;
    for (int e = tid; e < N * N; e += NT) Yj[e] = Ys[(e / N) * LS + (e % N)];
    if (tid == 0) {
        info[2 * blockIdx.x] = status; info[2 * blockIdx.x + 1] = sweeps;
        if (blockIdx.x == 0) clk[0] = t1 - t0;
    }
    if (PROF && blockIdx.x == 0 && lane == 0) {
        for (int i = 0; i < 8; ++i) clk[8 + warp * 8 + i] = acc[i];
    }
}

// ------------------------------------------------------------------------------------------------
// planar rows + packed FFMA2.  Row i: re[0..127] at Yp[i * RS], im[0..127] at Yp[i * RS + 128];
// lane l holds elements 4l .. 4l+3 of a row as two float2 of real and two of imaginary parts.
// ------------------------------------------------------------------------------------------------
constexpr int RS = 2 * N;        // floats per planar row

struct Row { float2 re[2], im[2]; };

__device__ __forceinline__ void row_load(Row& r, const float* p, int lane) {
    const float4 a = *reinterpret_cast<const float4*>(p + 4 * lane);
    const float4 b = *reinterpret_cast<const float4*>(p + N + 4 * lane);
    r.re[0] = make_float2(a.x, a.y); r.re[1] = make_float2(a.z, a.w);
    r.im[0] = make_float2(b.x, b.y); r.im[1] = make_float2(b.z, b.w);
}
__device__ __forceinline__ void row_store(const Row& r, float* p, int lane) {
    *reinterpret_cast<float4*>(p + 4 * lane) = make_float4(r.re[0].x, r.re[0].y, r.re[1].x, r.re[1].y);
    *reinterpret_cast<float4*>(p + N + 4 * lane) = make_float4(r.im[0].x, r.im[0].y, r.im[1].x, r.im[1].y);
}

// partial Gram entry <p, q> = sum p conj(q) over this lane's 4 elements
__device__ __forceinline__ void gram_part(const Row& p, const Row& q, float& gr, float& gi) {
    float2 r2 = make_float2(0.f, 0.f), m2 = r2, n2 = r2;
#pragma unroll
    for (int t = 0; t < 2; ++t) {
        r2 = __ffma2_rn(p.re[t], q.re[t], r2);
        r2 = __ffma2_rn(p.im[t], q.im[t], r2);
        m2 = __ffma2_rn(p.im[t], q.re[t], m2);
        n2 = __ffma2_rn(p.re[t], q.im[t], n2);
    }
    gr = r2.x + r2.y;
    gi = (m2.x - n2.x) + (m2.y - n2.y);
}

__device__ __forceinline__ void rot_apply_p(float c, float sr, float si, Row& p, Row& q) {
    const float2 C2 = make_float2(c, c), S2 = make_float2(sr, sr), I2 = make_float2(si, si);
    const float2 NS2 = make_float2(-sr, -sr), NI2 = make_float2(-si, -si);
#pragma unroll
    for (int t = 0; t < 2; ++t) {
        const float2 pre = p.re[t], pim = p.im[t], qre = q.re[t], qim = q.im[t];
        p.re[t] = __ffma2_rn(C2, pre, __ffma2_rn(S2, qre, __fmul2_rn(NI2, qim)));
        p.im[t] = __ffma2_rn(C2, pim, __ffma2_rn(S2, qim, __fmul2_rn(I2, qre)));
        q.re[t] = __ffma2_rn(C2, qre, __ffma2_rn(NS2, pre, __fmul2_rn(NI2, pim)));
        q.im[t] = __ffma2_rn(C2, qim, __ffma2_rn(I2, pre, __fmul2_rn(NS2, pim)));
    }
}

template <int A0, int A1, int A2, int A3, int B0, int B1, int B2, int B3>
__device__ __forceinline__ int sub_round_p4(Row (&y)[8], float (&a)[8], int lane, bool& big) {
    constexpr int PA[4] = {A0, A1, A2, A3}, PB[4] = {B0, B1, B2, B3};
    float gr[4], gi[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) gram_part(y[PA[i]], y[PB[i]], gr[i], gi[i]);
    const float mgr = reduce4(gr, lane), mgi = reduce4(gi, lane);
    const int sel = lane >> 3;
    float ap = a[PA[0]], aq = a[PB[0]];
#pragma unroll
    for (int i = 1; i < 4; ++i) if (sel == i) { ap = a[PA[i]]; aq = a[PB[i]]; }
    const float g2 = fmaf(mgr, mgr, mgi * mgi), apq = ap * aq;
    float c = 1.f, sr = 0.f, si = 0.f, tg = 0.f;
    const bool dorot = (g2 > TOL2 * apq) && (g2 > 1e-30f);
    big = big || (dorot && g2 > BIG2 * apq);
    if (dorot) rot_params(ap, aq, mgr, mgi, g2, c, sr, si, tg);
    const unsigned bal = __ballot_sync(0xffffffffu, dorot);
    const unsigned flags = (bal & 1u) | ((bal >> 7) & 2u) | ((bal >> 14) & 4u) | ((bal >> 21) & 8u);
    if (flags == 0u) return 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        if (flags & (1u << i)) {
            const float ci = __shfl_sync(0xffffffffu, c, 8 * i);
            const float sri = __shfl_sync(0xffffffffu, sr, 8 * i);
            const float sii = __shfl_sync(0xffffffffu, si, 8 * i);
            const float tgi = __shfl_sync(0xffffffffu, tg, 8 * i);
            rot_apply_p(ci, sri, sii, y[PA[i]], y[PB[i]]);
            a[PA[i]] = fmaxf(a[PA[i]] + tgi, 0.f);
            a[PB[i]] = fmaxf(a[PB[i]] - tgi, 0.f);
        }
    }
    return __popc(flags);
}

__device__ __forceinline__ void load_planar(float* Yp, const cf* Xj, int tid, int nt) {
    for (int e = tid; e < N * N; e += nt) {
        const int i = e / N, c = e % N;
        const cf v = Xj[e];
        Yp[i * RS + c] = v.x; Yp[i * RS + N + c] = v.y;
    }
}
__device__ __forceinline__ void store_planar(const float* Yp, cf* Yj, int tid, int nt) {
    for (int e = tid; e < N * N; e += nt) {
        const int i = e / N, c = e % N;
        Yj[e] = make_float2(Yp[i * RS + c], Yp[i * RS + N + c]);
    }
}
__device__ __forceinline__ void refresh_norms(const float* Yp, float* nrm, int warp, int lane, int nw) {
    for (int i = warp; i < N; i += nw) {
        Row r; row_load(r, Yp + i * RS, lane);
        float2 s2 = make_float2(0.f, 0.f);
#pragma unroll
        for (int t = 0; t < 2; ++t) { s2 = __ffma2_rn(r.re[t], r.re[t], s2); s2 = __ffma2_rn(r.im[t], r.im[t], s2); }
        float s = s2.x + s2.y;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) nrm[i] = s;
    }
}

__global__ void __launch_bounds__(512, 1) jacobi_v2(const cf* X, cf* Yout, int* info, long long* clk, int max_sweeps) {
    extern __shared__ float4 smem_raw[];
    constexpr int NT = 512, NW = 16;
    float* Yp = (float*)smem_raw;
    float* nrm = Yp + N * RS;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    load_planar(Yp, X + (size_t)blockIdx.x * N * N, tid, NT);
    __syncthreads();
    constexpr int nb = N / 4, mcirc = nb - 1, nrounds = nb - 1, ngroups = nb / 2;
    int sweeps = 0, status = 1;
    long long t0 = clock64();
    for (int sweep = 0; sweep < max_sweeps; ++sweep) {
        refresh_norms(Yp, nrm, warp, lane, NW);
        __syncthreads();
        bool big = false;
        for (int r = 0; r < nrounds; ++r) {
            const int g = warp;
            int I, J;
            if (g == 0) { I = mcirc; J = r; }
            else { I = (r + g) % mcirc; J = (r - g + mcirc) % mcirc; }
            float* rowA = Yp + (4 * I) * RS;
            float* rowB = Yp + (4 * J) * RS;
            Row v[8];
            float a[8];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                row_load(v[i], rowA + i * RS, lane);
                row_load(v[4 + i], rowB + i * RS, lane);
                a[i] = nrm[4 * I + i];
                a[4 + i] = nrm[4 * J + i];
            }
            int nrot = 0;
            if (r == 0) {
                nrot += sub_round_p4<0, 2, 4, 6, 1, 3, 5, 7>(v, a, lane, big);
                nrot += sub_round_p4<0, 1, 4, 5, 2, 3, 6, 7>(v, a, lane, big);
                nrot += sub_round_p4<0, 1, 4, 5, 3, 2, 7, 6>(v, a, lane, big);
            }
            nrot += sub_round_p4<0, 1, 2, 3, 4, 5, 6, 7>(v, a, lane, big);
            nrot += sub_round_p4<0, 1, 2, 3, 5, 6, 7, 4>(v, a, lane, big);
            nrot += sub_round_p4<0, 1, 2, 3, 6, 7, 4, 5>(v, a, lane, big);
            nrot += sub_round_p4<0, 1, 2, 3, 7, 4, 5, 6>(v, a, lane, big);
            if (nrot) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    row_store(v[i], rowA + i * RS, lane);
                    row_store(v[4 + i], rowB + i * RS, lane);
                }
                float am = a[0];
#pragma unroll
                for (int i = 1; i < 8; ++i) if (lane == i) am = a[i];
                if (lane < 8) nrm[lane < 4 ? 4 * I + lane : 4 * J + lane - 4] = am;
            }
            if (r + 1 < nrounds) neighbour_sync(g, ngroups);
        }
        sweeps = sweep + 1;
        if (!__syncthreads_or(big ? 1 : 0)) { status = 0; break; }
    }
    long long t1 = clock64();
    store_planar(Yp, Yout + (size_t)blockIdx.x * N * N, tid, NT);
    if (tid == 0) {
        info[2 * blockIdx.x] = status; info[2 * blockIdx.x + 1] = sweeps;
        if (blockIdx.x == 0) clk[0] = t1 - t0;
    }
}

__device__ unsigned g_stagger_ns;

// V7: V2 with half of the warps of every scheduler delayed at the start of each round
__global__ void __launch_bounds__(512, 1) jacobi_v7(const cf* X, cf* Yout, int* info, long long* clk, int max_sweeps) {
    extern __shared__ float4 smem_raw[];
    constexpr int NT = 512, NW = 16;
    float* Yp = (float*)smem_raw;
    float* nrm = Yp + N * RS;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    load_planar(Yp, X + (size_t)blockIdx.x * N * N, tid, NT);
    __syncthreads();
    constexpr int nb = N / 4, mcirc = nb - 1, nrounds = nb - 1, ngroups = nb / 2;
    int sweeps = 0, status = 1;
    long long t0 = clock64();
    for (int sweep = 0; sweep < max_sweeps; ++sweep) {
        refresh_norms(Yp, nrm, warp, lane, NW);
        __syncthreads();
        bool big = false;
        for (int r = 0; r < nrounds; ++r) {
            // STAGGER: the warps of a scheduler (w, w+4, w+8, w+12) run the rounds in lock-step -- all four in
            // the FMA-bound apply, then all four in the reduction / parameter latency chains.  Half of them
            // (w in 4-7, 12-15: two per scheduler) start every round g_stagger_ns later, so that one pair's FMA
            // phases fall into the other pair's latency phases.
            if (g_stagger_ns && ((warp >> 2) & 1)) __nanosleep(g_stagger_ns);
            const int g = warp;
            int I, J;
            if (g == 0) { I = mcirc; J = r; }
            else { I = (r + g) % mcirc; J = (r - g + mcirc) % mcirc; }
            float* rowA = Yp + (4 * I) * RS;
            float* rowB = Yp + (4 * J) * RS;
            Row v[8];
            float a[8];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                row_load(v[i], rowA + i * RS, lane);
                row_load(v[4 + i], rowB + i * RS, lane);
                a[i] = nrm[4 * I + i];
                a[4 + i] = nrm[4 * J + i];
            }
            int nrot = 0;
            if (r == 0) {
                nrot += sub_round_p4<0, 2, 4, 6, 1, 3, 5, 7>(v, a, lane, big);
                nrot += sub_round_p4<0, 1, 4, 5, 2, 3, 6, 7>(v, a, lane, big);
                nrot += sub_round_p4<0, 1, 4, 5, 3, 2, 7, 6>(v, a, lane, big);
            }
            nrot += sub_round_p4<0, 1, 2, 3, 4, 5, 6, 7>(v, a, lane, big);
            nrot += sub_round_p4<0, 1, 2, 3, 5, 6, 7, 4>(v, a, lane, big);
            nrot += sub_round_p4<0, 1, 2, 3, 6, 7, 4, 5>(v, a, lane, big);
            nrot += sub_round_p4<0, 1, 2, 3, 7, 4, 5, 6>(v, a, lane, big);
            if (nrot) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    row_store(v[i], rowA + i * RS, lane);
                    row_store(v[4 + i], rowB + i * RS, lane);
                }
                float am = a[0];
#pragma unroll
                for (int i = 1; i < 8; ++i) if (lane == i) am = a[i];
                if (lane < 8) nrm[lane < 4 ? 4 * I + lane : 4 * J + lane - 4] = am;
            }
            if (r + 1 < nrounds) neighbour_sync(g, ngroups);
        }
        sweeps = sweep + 1;
        if (!__syncthreads_or(big ? 1 : 0)) { status = 0; break; }
    }
    long long t1 = clock64();
    store_planar(Yp, Yout + (size_t)blockIdx.x * N * N, tid, NT);
    if (tid == 0) {
        info[2 * blockIdx.x] = status; info[2 * blockIdx.x + 1] = sweeps;
        if (blockIdx.x == 0) clk[0] = t1 - t0;
    }
}

// ------------------------------------------------------------------------------------------------
// V8: V2 with the FMA-heavy apply phases of the two warp pairs of a scheduler forced to ALTERNATE.  Warps w and
// w ^ 4 share a scheduler; each passes a token (two mbarriers per pair, one arrival per apply) so that one of
// them applies its rotations while the other is in its reduction / parameter latency chain, instead of all
// four warps of the scheduler contending for the FMA pipe at once and then idling at once.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32_(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init_(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32_(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_(unsigned long long* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32_(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait_(unsigned long long* bar, unsigned parity) {
    unsigned ok = 0;
    while (!ok) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(smem_u32_(bar)), "r"(parity) : "memory");
    }
}

template <int A0, int A1, int A2, int A3, int B0, int B1, int B2, int B3>
__device__ __forceinline__ int sub_round_t4(Row (&y)[8], float (&a)[8], int lane, bool& big,
                                            unsigned long long* mine, unsigned long long* partner, bool lower, unsigned& k) {
    constexpr int PA[4] = {A0, A1, A2, A3}, PB[4] = {B0, B1, B2, B3};
    float gr[4], gi[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) gram_part(y[PA[i]], y[PB[i]], gr[i], gi[i]);
    const float mgr = reduce4(gr, lane), mgi = reduce4(gi, lane);
    const int sel = lane >> 3;
    float ap = a[PA[0]], aq = a[PB[0]];
#pragma unroll
    for (int i = 1; i < 4; ++i) if (sel == i) { ap = a[PA[i]]; aq = a[PB[i]]; }
    const float g2 = fmaf(mgr, mgr, mgi * mgi), apq = ap * aq;
    float c = 1.f, sr = 0.f, si = 0.f, tg = 0.f;
    const bool dorot = (g2 > TOL2 * apq) && (g2 > 1e-30f);
    big = big || (dorot && g2 > BIG2 * apq);
    if (dorot) rot_params(ap, aq, mgr, mgi, g2, c, sr, si, tg);
    const unsigned bal = __ballot_sync(0xffffffffu, dorot);
    const unsigned flags = (bal & 1u) | ((bal >> 7) & 2u) | ((bal >> 14) & 4u) | ((bal >> 21) & 8u);
    // token: the lower warp's k-th apply waits for the partner's (k-1)-th, the partner's k-th for the lower's k-th
    if (lower) { if (k > 0) mbar_wait_(partner, (k - 1) & 1u); }
    else mbar_wait_(partner, k & 1u);
    int n = 0;
    if (flags) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            if (flags & (1u << i)) {
                const float ci = __shfl_sync(0xffffffffu, c, 8 * i);
                const float sri = __shfl_sync(0xffffffffu, sr, 8 * i);
                const float sii = __shfl_sync(0xffffffffu, si, 8 * i);
                const float tgi = __shfl_sync(0xffffffffu, tg, 8 * i);
                rot_apply_p(ci, sri, sii, y[PA[i]], y[PB[i]]);
                a[PA[i]] = fmaxf(a[PA[i]] + tgi, 0.f);
                a[PB[i]] = fmaxf(a[PB[i]] - tgi, 0.f);
            }
        }
        n = __popc(flags);
    }
    __syncwarp();
    if (lane == 0) mbar_arrive_(mine);
    ++k;
    return n;
}

__global__ void __launch_bounds__(512, 1) jacobi_v8(const cf* X, cf* Yout, int* info, long long* clk, int max_sweeps) {
    extern __shared__ float4 smem_raw[];
    __shared__ __align__(8) unsigned long long tok[16];
    constexpr int NT = 512, NW = 16;
    float* Yp = (float*)smem_raw;
    float* nrm = Yp + N * RS;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid < 16) mbar_init_(&tok[tid], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    load_planar(Yp, X + (size_t)blockIdx.x * N * N, tid, NT);
    __syncthreads();
    unsigned long long* mine = &tok[warp];
    unsigned long long* partner = &tok[warp ^ 4];
    const bool lower = (warp & 4) == 0;
    unsigned k = 0;
    constexpr int nb = N / 4, mcirc = nb - 1, nrounds = nb - 1, ngroups = nb / 2;
    int sweeps = 0, status = 1;
    long long t0 = clock64();
    for (int sweep = 0; sweep < max_sweeps; ++sweep) {
        refresh_norms(Yp, nrm, warp, lane, NW);
        __syncthreads();
        bool big = false;
        for (int r = 0; r < nrounds; ++r) {
            const int g = warp;
            int I, J;
            if (g == 0) { I = mcirc; J = r; }
            else { I = (r + g) % mcirc; J = (r - g + mcirc) % mcirc; }
            float* rowA = Yp + (4 * I) * RS;
            float* rowB = Yp + (4 * J) * RS;
            Row v[8];
            float a[8];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                row_load(v[i], rowA + i * RS, lane);
                row_load(v[4 + i], rowB + i * RS, lane);
                a[i] = nrm[4 * I + i];
                a[4 + i] = nrm[4 * J + i];
            }
            int nrot = 0;
            if (r == 0) {
                nrot += sub_round_t4<0, 2, 4, 6, 1, 3, 5, 7>(v, a, lane, big, mine, partner, lower, k);
                nrot += sub_round_t4<0, 1, 4, 5, 2, 3, 6, 7>(v, a, lane, big, mine, partner, lower, k);
                nrot += sub_round_t4<0, 1, 4, 5, 3, 2, 7, 6>(v, a, lane, big, mine, partner, lower, k);
            }
            nrot += sub_round_t4<0, 1, 2, 3, 4, 5, 6, 7>(v, a, lane, big, mine, partner, lower, k);
            nrot += sub_round_t4<0, 1, 2, 3, 5, 6, 7, 4>(v, a, lane, big, mine, partner, lower, k);
            nrot += sub_round_t4<0, 1, 2, 3, 6, 7, 4, 5>(v, a, lane, big, mine, partner, lower, k);
            nrot += sub_round_t4<0, 1, 2, 3, 7, 4, 5, 6>(v, a, lane, big, mine, partner, lower, k);
            if (nrot) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    row_store(v[i], rowA + i * RS, lane);
                    row_store(v[4 + i], rowB + i * RS, lane);
                }
                float am = a[0];
#pragma unroll
                for (int i = 1; i < 8; ++i) if (lane == i) am = a[i];
                if (lane < 8) nrm[lane < 4 ? 4 * I + lane : 4 * J + lane - 4] = am;
            }
            if (r + 1 < nrounds) neighbour_sync(g, ngroups);
        }
        sweeps = sweep + 1;
        if (!__syncthreads_or(big ? 1 : 0)) { status = 0; break; }
    }
    long long t1 = clock64();
    store_planar(Yp, Yout + (size_t)blockIdx.x * N * N, tid, NT);
    if (tid == 0) {
        info[2 * blockIdx.x] = status; info[2 * blockIdx.x + 1] = sweeps;
        if (blockIdx.x == 0) clk[0] = t1 - t0;
    }
}

// ------------------------------------------------------------------------------------------------
// V5: the Gram entries of a block pair are formed ONCE per round.  One pass over the 8 rows gives the
// 4 x 4 cross block C[i][j] = <A_i, B_j> (128 FFMA2 per lane), ONE transposed reduction of its 32 real
// numbers leaves entry (i, j) on lane pair 8i + 2j (+1: im), and the four sub-rounds then run on those
// scalars: rotation (A_i, B_(i+s)&3) from the current C[i][j] and the running squared norms, after which
// every other entry is rescaled by the cosines of the two rotations that touched its rows ("lite" update:
// the terms it drops are second order in the rotation angles; the rotations stay exactly unitary, only
// their angles are approximate; convergence measured on circuit thetas: same number of sweeps).  The 16
// rotations are then applied to the rows in one uninterrupted FFMA2 stream.  Per round: one reduction
// chain instead of four, and long pure-FMA phases that the other warps of the scheduler can hide behind.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float2 neg2(float2 a) { return make_float2(-a.x, -a.y); }

__device__ __forceinline__ void gram_part_n(const Row& p, const Row& q, float& gr, float& gi) {
    float2 r2 = make_float2(0.f, 0.f), m2 = r2;
#pragma unroll
    for (int t = 0; t < 2; ++t) {
        r2 = __ffma2_rn(p.re[t], q.re[t], r2);
        r2 = __ffma2_rn(p.im[t], q.im[t], r2);
        m2 = __ffma2_rn(p.im[t], q.re[t], m2);
        m2 = __ffma2_rn(neg2(p.re[t]), q.im[t], m2);
    }
    gr = r2.x + r2.y;
    gi = m2.x + m2.y;
}

// transposed reduction of 32 values over the 32 lanes: afterwards lane l holds the full sum of v[l]
__device__ __forceinline__ float reduce32(float (&v)[32], int lane) {
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
        const bool hi = (lane & d) != 0;
#pragma unroll
        for (int k = 0; k < d; ++k) {
            const float keep = hi ? v[k + d] : v[k];
            const float send = hi ? v[k] : v[k + d];
            v[k] = keep + __shfl_xor_sync(0xffffffffu, send, d);
        }
    }
    return v[0];
}

template <bool PROF>
__global__ void __launch_bounds__(512, 1) jacobi_v5(const cf* X, cf* Yout, int* info, long long* clk, int max_sweeps) {
    unsigned acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    unsigned tlast = 0;
    extern __shared__ float4 smem_raw[];
    constexpr int NT = 512, NW = 16;
    float* Yp = (float*)smem_raw;
    float* nrm = Yp + N * RS;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    load_planar(Yp, X + (size_t)blockIdx.x * N * N, tid, NT);
    __syncthreads();
    constexpr int nb = N / 4, mcirc = nb - 1, nrounds = nb - 1, ngroups = nb / 2;
    const int li = lane >> 3, lj = (lane >> 1) & 3;
    const bool im_lane = (lane & 1) != 0;
    int sweeps = 0, status = 1;
    long long t0 = clock64();
    for (int sweep = 0; sweep < max_sweeps; ++sweep) {
        refresh_norms(Yp, nrm, warp, lane, NW);
        __syncthreads();
        bool big = false;
        for (int r = 0; r < nrounds; ++r) {
            const int g = warp;
            int I, J;
            if (g == 0) { I = mcirc; J = r; }
            else { I = (r + g) % mcirc; J = (r - g + mcirc) % mcirc; }
            float* rowA = Yp + (4 * I) * RS;
            float* rowB = Yp + (4 * J) * RS;
            Row v[8];
            if (PROF) tlast = (unsigned)clock();
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                row_load(v[i], rowA + i * RS, lane);
                row_load(v[4 + i], rowB + i * RS, lane);
            }
            if (PROF) {
#pragma unroll
                for (int i = 0; i < 8; ++i) asm volatile("" :: "f"(v[i].re[0].x), "f"(v[i].im[1].y));
            }
            TICK(4);
            int nrot = 0;
            if (r == 0) {
                // the 12 rotations inside the two blocks: three exact sub-rounds
                float a[8];
#pragma unroll
                for (int i = 0; i < 4; ++i) { a[i] = nrm[4 * I + i]; a[4 + i] = nrm[4 * J + i]; }
                nrot += sub_round_p4<0, 2, 4, 6, 1, 3, 5, 7>(v, a, lane, big);
                nrot += sub_round_p4<0, 1, 4, 5, 2, 3, 6, 7>(v, a, lane, big);
                nrot += sub_round_p4<0, 1, 4, 5, 3, 2, 7, 6>(v, a, lane, big);
                float am = a[0];
#pragma unroll
                for (int i = 1; i < 8; ++i) if (lane == i) am = a[i];
                if (lane < 8) nrm[lane < 4 ? 4 * I + lane : 4 * J + lane - 4] = am;
                __syncwarp();
            }
            // ---- cross block: Gram once ----
            float vals[32];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) gram_part_n(v[i], v[4 + j], vals[8 * i + 2 * j], vals[8 * i + 2 * j + 1]);
            if (PROF) {
#pragma unroll
                for (int i = 0; i < 32; ++i) asm volatile("" :: "f"(vals[i]));
            }
            TICK(0);
            const float mine = reduce32(vals, lane);
            const float other = __shfl_xor_sync(0xffffffffu, mine, 1);
            float gr = im_lane ? other : mine, gi = im_lane ? mine : other;
            if (PROF) { asm volatile("" :: "f"(gr), "f"(gi)); }
            TICK(1);
            float aA = nrm[4 * I + li], aB = nrm[4 * J + lj];
            float rc[4], rsr[4], rsi[4];
            unsigned rflags = 0;
#pragma unroll
            for (int s = 0; s < 4; ++s) {
                const bool active = lj == ((li + s) & 3);
                const float g2 = fmaf(gr, gr, gi * gi), apq = aA * aB;
                const bool dorot = active && (g2 > TOL2 * apq) && (g2 > 1e-30f);
                big = big || (dorot && g2 > BIG2 * apq);
                float c = 1.f, sr = 0.f, si = 0.f, tg = 0.f;
                const unsigned bal = __ballot_sync(0xffffffffu, dorot);
                if (bal) {                                     // warp-uniform
                    if (dorot) rot_params(aA, aB, gr, gi, g2, c, sr, si, tg);
                    const int srcA = 8 * li + 2 * ((li + s) & 3);          // rotation of row A_li in this sub-round
                    const int srcB = 8 * ((lj - s) & 3) + 2 * lj;          // rotation of row B_lj
                    const float cA = __shfl_sync(0xffffffffu, c, srcA), tA = __shfl_sync(0xffffffffu, tg, srcA);
                    const float cB = __shfl_sync(0xffffffffu, c, srcB), tB = __shfl_sync(0xffffffffu, tg, srcB);
                    aA = fmaxf(aA + tA, 0.f);
                    aB = fmaxf(aB - tB, 0.f);
                    const float sc = cA * cB;
                    gr *= sc; gi *= sc;
#pragma unroll
                    for (int i = 0; i < 4; ++i) rflags |= ((bal >> (8 * i + 2 * ((i + s) & 3))) & 1u) << (4 * s + i);
                }
                rc[s] = c; rsr[s] = sr; rsi[s] = si;
            }
            if (PROF) { asm volatile("" :: "f"(rc[3]), "f"(rsr[3]), "f"(rsi[3]), "f"(aA), "f"(aB), "r"(rflags)); }
            TICK(2);
            if (rflags) {
#pragma unroll
                for (int s = 0; s < 4; ++s)
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        if (rflags & (1u << (4 * s + i))) {    // warp-uniform
                            const int src = 8 * i + 2 * ((i + s) & 3);
                            const float ci = __shfl_sync(0xffffffffu, rc[s], src);
                            const float sri = __shfl_sync(0xffffffffu, rsr[s], src);
                            const float sii = __shfl_sync(0xffffffffu, rsi[s], src);
                            rot_apply_p(ci, sri, sii, v[i], v[4 + ((i + s) & 3)]);
                        }
                    }
                nrot += __popc(rflags);
                if ((lane & 7) == 0) nrm[4 * I + li] = aA;
                if (lane < 8 && !im_lane) nrm[4 * J + lj] = aB;
            }
            if (PROF) {
#pragma unroll
                for (int i = 0; i < 8; ++i) asm volatile("" :: "f"(v[i].re[0].x), "f"(v[i].im[1].y));
            }
            TICK(3);
            if (nrot) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    row_store(v[i], rowA + i * RS, lane);
                    row_store(v[4 + i], rowB + i * RS, lane);
                }
            }
            TICK(5);
            if (r + 1 < nrounds) neighbour_sync(g, ngroups);
            TICK(6);
        }
        sweeps = sweep + 1;
        if (!__syncthreads_or(big ? 1 : 0)) { status = 0; break; }
    }
    long long t1 = clock64();
    store_planar(Yp, Yout + (size_t)blockIdx.x * N * N, tid, NT);
    if (tid == 0) {
        info[2 * blockIdx.x] = status; info[2 * blockIdx.x + 1] = sweeps;
        if (blockIdx.x == 0) clk[0] = t1 - t0;
    }
    if (PROF && blockIdx.x == 0 && lane == 0) {
        for (int i = 0; i < 8; ++i) clk[8 + warp * 8 + i] = acc[i];
    }
}

// ------------------------------------------------------------------------------------------------
// V6: V2 with FAST (scaled) rotations.  Row p is stored as y_p with a log2 scale l_p: true row = 2^l_p y_p.
// The rotation [c s; -conj(s) c] of the true rows becomes two complex axpys on the stored rows,
//   y_p <- y_p + alpha y_q,   y_q <- y_q - beta y_p(old),   alpha = (s/c) 2^(l_q - l_p),  beta = conj(s/c) 2^(l_p - l_q),
// and both scales take the factor c (l += log2 c): 8 FMA per element pair instead of 12.  Gram entries,
// thresholds, rotation parameters and the running squared norms are those of the TRUE rows.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float ex2_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float lg2_approx(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_approx_(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

__device__ __forceinline__ void fast_apply_p(float ar, float ai, float br, float bi, Row& p, Row& q) {
    const float2 AR = make_float2(ar, ar), AI = make_float2(ai, ai), NAI = make_float2(-ai, -ai);
    const float2 NBR = make_float2(-br, -br), BI = make_float2(bi, bi), NBI = make_float2(-bi, -bi);
#pragma unroll
    for (int t = 0; t < 2; ++t) {
        const float2 pre = p.re[t], pim = p.im[t], qre = q.re[t], qim = q.im[t];
        p.re[t] = __ffma2_rn(AR, qre, __ffma2_rn(NAI, qim, pre));      // p + alpha q
        p.im[t] = __ffma2_rn(AR, qim, __ffma2_rn(AI, qre, pim));
        q.re[t] = __ffma2_rn(NBR, pre, __ffma2_rn(BI, pim, qre));      // q - beta p(old)
        q.im[t] = __ffma2_rn(NBR, pim, __ffma2_rn(NBI, pre, qim));
    }
}

// the squared norms and log2 scales of the rows stay in shared memory (nrm / lsc, indexed by row): the lanes of
// octet i fetch those of their pair and lane 8 i writes the updated values back (no per-lane copies of all 8)
template <int A0, int A1, int A2, int A3, int B0, int B1, int B2, int B3>
__device__ __forceinline__ int sub_round_f4(Row (&y)[8], float* nrm, float* lsc, int rowI, int rowJ, int lane, bool& big) {
    constexpr int PA[4] = {A0, A1, A2, A3}, PB[4] = {B0, B1, B2, B3};
    const int sel = lane >> 3;
    int xa = PA[0], xb = PB[0];
#pragma unroll
    for (int i = 1; i < 4; ++i) if (sel == i) { xa = PA[i]; xb = PB[i]; }
    const int rp = xa < 4 ? rowI + xa : rowJ + xa - 4, rq = xb < 4 ? rowI + xb : rowJ + xb - 4;
    const float ap = nrm[rp], aq = nrm[rq], lp = lsc[rp], lq = lsc[rq];
    float gr[4], gi[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) gram_part(y[PA[i]], y[PB[i]], gr[i], gi[i]);
    float mgr = reduce4(gr, lane), mgi = reduce4(gi, lane);
    const float dd = ex2_approx(lp + lq);
    mgr *= dd; mgi *= dd;                                      // Gram entry of the true rows
    const float g2 = fmaf(mgr, mgr, mgi * mgi), apq = ap * aq;
    float c = 1.f, sr = 0.f, si = 0.f, tg = 0.f;
    const bool dorot = (g2 > TOL2 * apq) && (g2 > 1e-30f);
    big = big || (dorot && g2 > BIG2 * apq);
    float ar = 0.f, ai = 0.f, br = 0.f, bi = 0.f;
    if (dorot) {
        rot_params(ap, aq, mgr, mgi, g2, c, sr, si, tg);
        float ic = rcp_approx_(c);
        ic = ic * fmaf(-c, ic, 2.0f);                          // Newton: c in [0.707, 1]
        const float tr = sr * ic, ti = si * ic;                // t = s / c
        const float eq = ex2_approx(lq - lp), ep = ex2_approx(lp - lq);
        ar = tr * eq; ai = ti * eq;
        br = tr * ep; bi = -ti * ep;
        const float h = fmaf(sr, sr, si * si);                 // c^2 = 1 - h
        float lc;
        if (h < 0.0625f) lc = -0.72134752f * h * fmaf(h, fmaf(h, fmaf(h, 0.25f, 0.33333334f), 0.5f), 1.0f);
        else lc = lg2_approx(c);
        if ((lane & 7) == 0) {
            nrm[rp] = fmaxf(ap + tg, 0.f); nrm[rq] = fmaxf(aq - tg, 0.f);
            lsc[rp] = lp + lc; lsc[rq] = lq + lc;
        }
    }
    const unsigned bal = __ballot_sync(0xffffffffu, dorot);
    const unsigned flags = (bal & 1u) | ((bal >> 7) & 2u) | ((bal >> 14) & 4u) | ((bal >> 21) & 8u);
    if (flags == 0u) return 0;
    __syncwarp();                                              // the updates of nrm / lsc are visible to the next sub-round
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        if (flags & (1u << i)) {
            const float ari = __shfl_sync(0xffffffffu, ar, 8 * i), aii = __shfl_sync(0xffffffffu, ai, 8 * i);
            const float bri = __shfl_sync(0xffffffffu, br, 8 * i), bii = __shfl_sync(0xffffffffu, bi, 8 * i);
            fast_apply_p(ari, aii, bri, bii, y[PA[i]], y[PB[i]]);
        }
    }
    return __popc(flags);
}

__global__ void __launch_bounds__(512, 1) jacobi_v6(const cf* X, cf* Yout, int* info, long long* clk, int max_sweeps) {
    extern __shared__ float4 smem_raw[];
    constexpr int NT = 512, NW = 16;
    float* Yp = (float*)smem_raw;
    float* nrm = Yp + N * RS;
    float* lsc = nrm + N;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    load_planar(Yp, X + (size_t)blockIdx.x * N * N, tid, NT);
    for (int i = tid; i < N; i += NT) lsc[i] = 0.f;
    __syncthreads();
    constexpr int nb = N / 4, mcirc = nb - 1, nrounds = nb - 1, ngroups = nb / 2;
    int sweeps = 0, status = 1;
    long long t0 = clock64();
    for (int sweep = 0; sweep < max_sweeps; ++sweep) {
        // true squared norms: 4^l |y|^2
        for (int i = warp; i < N; i += NW) {
            Row r; row_load(r, Yp + i * RS, lane);
            float2 s2 = make_float2(0.f, 0.f);
#pragma unroll
            for (int t = 0; t < 2; ++t) { s2 = __ffma2_rn(r.re[t], r.re[t], s2); s2 = __ffma2_rn(r.im[t], r.im[t], s2); }
            float sum = s2.x + s2.y;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
            if (lane == 0) nrm[i] = sum * ex2_approx(2.0f * lsc[i]);
        }
        __syncthreads();
        bool big = false;
        for (int r = 0; r < nrounds; ++r) {
            const int g = warp;
            int I, J;
            if (g == 0) { I = mcirc; J = r; }
            else { I = (r + g) % mcirc; J = (r - g + mcirc) % mcirc; }
            float* rowA = Yp + (4 * I) * RS;
            float* rowB = Yp + (4 * J) * RS;
            Row v[8];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                row_load(v[i], rowA + i * RS, lane);
                row_load(v[4 + i], rowB + i * RS, lane);
            }
            int nrot = 0;
            if (r == 0) {
                nrot += sub_round_f4<0, 2, 4, 6, 1, 3, 5, 7>(v, nrm, lsc, 4 * I, 4 * J, lane, big);
                nrot += sub_round_f4<0, 1, 4, 5, 2, 3, 6, 7>(v, nrm, lsc, 4 * I, 4 * J, lane, big);
                nrot += sub_round_f4<0, 1, 4, 5, 3, 2, 7, 6>(v, nrm, lsc, 4 * I, 4 * J, lane, big);
            }
            nrot += sub_round_f4<0, 1, 2, 3, 4, 5, 6, 7>(v, nrm, lsc, 4 * I, 4 * J, lane, big);
            nrot += sub_round_f4<0, 1, 2, 3, 5, 6, 7, 4>(v, nrm, lsc, 4 * I, 4 * J, lane, big);
            nrot += sub_round_f4<0, 1, 2, 3, 6, 7, 4, 5>(v, nrm, lsc, 4 * I, 4 * J, lane, big);
            nrot += sub_round_f4<0, 1, 2, 3, 7, 4, 5, 6>(v, nrm, lsc, 4 * I, 4 * J, lane, big);
            if (nrot) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    row_store(v[i], rowA + i * RS, lane);
                    row_store(v[4 + i], rowB + i * RS, lane);
                }
            }
            if (r + 1 < nrounds) neighbour_sync(g, ngroups);
        }
        sweeps = sweep + 1;
        if (!__syncthreads_or(big ? 1 : 0)) { status = 0; break; }
    }
    long long t1 = clock64();
    __syncthreads();
    // back to true rows
    for (int i = warp; i < N; i += NW) {
        const float d = ex2_approx(lsc[i]);
        for (int c = lane; c < 2 * N; c += 32) Yp[i * RS + c] *= d;
    }
    __syncthreads();
    store_planar(Yp, Yout + (size_t)blockIdx.x * N * N, tid, NT);
    if (tid == 0) {
        info[2 * blockIdx.x] = status; info[2 * blockIdx.x + 1] = sweeps;
        if (blockIdx.x == 0) clk[0] = t1 - t0;
    }
}

// ---- 1024 threads, 2-row blocks -----------------------------------------------------------------
// sub-round of 2 disjoint rotations (A0,B0), (A1,B1): after the 16-step a half-warp owns one pair,
// after the 8-step a quarter owns its re or im part; one more shuffle fetches the other part.
template <int A0, int B0, int A1, int B1>
__device__ __forceinline__ int sub_round_p2(Row (&y)[4], float (&a)[4], int lane, bool& big) {
    constexpr int PA[2] = {A0, A1}, PB[2] = {B0, B1};
    float gr[2], gi[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) gram_part(y[PA[i]], y[PB[i]], gr[i], gi[i]);
    const bool h16 = (lane & 16) != 0, h8 = (lane & 8) != 0;
    float kr = h16 ? gr[1] : gr[0], ki = h16 ? gi[1] : gi[0];
    const float sr_ = h16 ? gr[0] : gr[1], si_ = h16 ? gi[0] : gi[1];
    kr += __shfl_xor_sync(0xffffffffu, sr_, 16);
    ki += __shfl_xor_sync(0xffffffffu, si_, 16);
    float k = h8 ? ki : kr;
    const float sd = h8 ? kr : ki;
    k += __shfl_xor_sync(0xffffffffu, sd, 8);
    k += __shfl_xor_sync(0xffffffffu, k, 4);
    k += __shfl_xor_sync(0xffffffffu, k, 2);
    k += __shfl_xor_sync(0xffffffffu, k, 1);
    const float other = __shfl_xor_sync(0xffffffffu, k, 8);
    const float mgr = h8 ? other : k, mgi = h8 ? k : other;
    const float ap = h16 ? a[PA[1]] : a[PA[0]], aq = h16 ? a[PB[1]] : a[PB[0]];
    const float g2 = fmaf(mgr, mgr, mgi * mgi), apq = ap * aq;
    float c = 1.f, sr = 0.f, si = 0.f, tg = 0.f;
    const bool dorot = (g2 > TOL2 * apq) && (g2 > 1e-30f);
    big = big || (dorot && g2 > BIG2 * apq);
    if (dorot) rot_params(ap, aq, mgr, mgi, g2, c, sr, si, tg);
    const unsigned bal = __ballot_sync(0xffffffffu, dorot);
    const unsigned flags = (bal & 1u) | ((bal >> 15) & 2u);
    if (flags == 0u) return 0;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        if (flags & (1u << i)) {
            const float ci = __shfl_sync(0xffffffffu, c, 16 * i);
            const float sri = __shfl_sync(0xffffffffu, sr, 16 * i);
            const float sii = __shfl_sync(0xffffffffu, si, 16 * i);
            const float tgi = __shfl_sync(0xffffffffu, tg, 16 * i);
            rot_apply_p(ci, sri, sii, y[PA[i]], y[PB[i]]);
            a[PA[i]] = fmaxf(a[PA[i]] + tgi, 0.f);
            a[PB[i]] = fmaxf(a[PB[i]] - tgi, 0.f);
        }
    }
    return __popc(flags);
}

template <int GROUPED>
__global__ void __launch_bounds__(1024, 1) jacobi_v3(const cf* X, cf* Yout, int* info, long long* clk, int max_sweeps) {
    extern __shared__ float4 smem_raw[];
    constexpr int NT = 1024, NW = 32;
    float* Yp = (float*)smem_raw;
    float* nrm = Yp + N * RS;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    load_planar(Yp, X + (size_t)blockIdx.x * N * N, tid, NT);
    __syncthreads();
    constexpr int nb = N / 2, mcirc = nb - 1, nrounds = nb - 1;      // 64 two-row blocks, 32 groups = NW
    int sweeps = 0, status = 1;
    long long t0 = clock64();
    for (int sweep = 0; sweep < max_sweeps; ++sweep) {
        refresh_norms(Yp, nrm, warp, lane, NW);
        __syncthreads();
        bool big = false;
        for (int r = 0; r < nrounds; ++r) {
            const int g = warp;
            int I, J;
            if (g == 0) { I = mcirc; J = r; }
            else { I = (r + g) % mcirc; J = (r - g + mcirc) % mcirc; }
            float* rowA = Yp + (2 * I) * RS;
            float* rowB = Yp + (2 * J) * RS;
            Row v[4];
            float a[4];
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                row_load(v[i], rowA + i * RS, lane);
                row_load(v[2 + i], rowB + i * RS, lane);
                a[i] = nrm[2 * I + i];
                a[2 + i] = nrm[2 * J + i];
            }
            int nrot = 0;
            if (r == 0) nrot += sub_round_p2<0, 1, 2, 3>(v, a, lane, big);
            nrot += sub_round_p2<0, 2, 1, 3>(v, a, lane, big);
            nrot += sub_round_p2<0, 3, 1, 2>(v, a, lane, big);
            if (nrot) {
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    row_store(v[i], rowA + i * RS, lane);
                    row_store(v[2 + i], rowB + i * RS, lane);
                }
                float am = a[0];
#pragma unroll
                for (int i = 1; i < 4; ++i) if (lane == i) am = a[i];
                if (lane < 4) nrm[lane < 2 ? 2 * I + lane : 2 * J + lane - 2] = am;
            }
            if (GROUPED) {
                // group g exchanges blocks with groups g-1 and g+1 only; edge e = (e, e+1), e = 0..30.
                // Even edges are the FIRST barrier of both their warps, odd edges the second, so edges may
                // share a hardware barrier only with edges of the same parity (else: circular wait):
                // e and e+16 share id 1 + e (4 warps); edge 15 joins edges 1 and 17 (6 warps).
                if (r + 1 < nrounds) {
                    const int el = g - 1, er = g;                    // left edge, right edge
                    auto edge_sync = [&](int e) {
                        const int id = e == 15 ? 1 : (e & 15);
                        named_barrier_sync(1 + id, id == 1 ? 192 : 128);
                    };
                    if (g & 1) { if (el >= 0) edge_sync(el); if (er <= 30) edge_sync(er); }
                    else { if (er <= 30) edge_sync(er); if (el >= 0) edge_sync(el); }
                }
            } else {
                __syncthreads();
            }
        }
        sweeps = sweep + 1;
        if (!__syncthreads_or(big ? 1 : 0)) { status = 0; break; }
    }
    long long t1 = clock64();
    store_planar(Yp, Yout + (size_t)blockIdx.x * N * N, tid, NT);
    if (tid == 0) {
        info[2 * blockIdx.x] = status; info[2 * blockIdx.x + 1] = sweeps;
        if (blockIdx.x == 0) clk[0] = t1 - t0;
    }
}

// float64 reference on the host: scalar cyclic one-sided Jacobi on the rows
static std::vector<double> host_singular_values(const std::vector<std::complex<double>>& M) {
    std::vector<std::complex<double>> Y = M;
    for (int sweep = 0; sweep < 60; ++sweep) {
        int nrot = 0;
        for (int p = 0; p < N; ++p)
            for (int q = p + 1; q < N; ++q) {
                double a = 0, b = 0; std::complex<double> g = 0;
                for (int c = 0; c < N; ++c) { a += std::norm(Y[p * N + c]); b += std::norm(Y[q * N + c]); g += Y[p * N + c] * std::conj(Y[q * N + c]); }
                double ag = std::abs(g);
                if (!(ag > 1e-14 * std::sqrt(a * b)) || ag == 0) continue;
                double zeta = (a - b) / (2 * ag), t = (zeta >= 0 ? 1.0 : -1.0) / (std::fabs(zeta) + std::sqrt(1 + zeta * zeta));
                double cc = 1 / std::sqrt(1 + t * t);
                std::complex<double> s = cc * t * g / ag;
                for (int c = 0; c < N; ++c) {
                    std::complex<double> yp = Y[p * N + c], yq = Y[q * N + c];
                    Y[p * N + c] = cc * yp + s * yq;
                    Y[q * N + c] = cc * yq - std::conj(s) * yp;
                }
                ++nrot;
            }
        if (!nrot) break;
    }
    std::vector<double> sv(N);
    for (int p = 0; p < N; ++p) { double a = 0; for (int c = 0; c < N; ++c) a += std::norm(Y[p * N + c]); sv[p] = std::sqrt(a); }
    std::sort(sv.begin(), sv.end(), std::greater<double>());
    return sv;
}

typedef void (*kern_t)(const cf*, cf*, int*, long long*, int);

int main(int argc, char** argv) {
    const int nmat = argc > 1 ? atoi(argv[1]) : 148, max_sweeps = argc > 2 ? atoi(argv[2]) : 30;
    std::vector<cf> h((size_t)nmat * N * N);
    srand(7);
    for (auto& v : h) { v.x = (rand() / (float)RAND_MAX - 0.5f); v.y = (rand() / (float)RAND_MAX - 0.5f); }
    cf *dX, *dY; int* dinfo; long long* dclk;
    cudaMalloc(&dX, h.size() * sizeof(cf)); cudaMalloc(&dY, h.size() * sizeof(cf));
    cudaMalloc(&dinfo, nmat * 2 * sizeof(int)); cudaMalloc(&dclk, 256 * sizeof(long long));
    cudaMemcpy(dX, h.data(), h.size() * sizeof(cf), cudaMemcpyHostToDevice);
    std::vector<std::complex<double>> M(N * N);
    for (int e = 0; e < N * N; ++e) M[e] = std::complex<double>(h[e].x, h[e].y);
    const std::vector<double> ref = host_singular_values(M);

    struct Variant { const char* name; kern_t k; int nt; int smem; };
    const int smem_v0 = (N * (N + 4)) * (int)sizeof(cf) + N * (int)sizeof(float);
    const int smem_p = N * RS * (int)sizeof(float) + N * (int)sizeof(float);
    Variant vs[] = {
        {"V0  512 thr, 4-row blocks, interleaved FFMA (production)", jacobi_v0<false>, 512, smem_v0},
        {"V0p the same with per-phase clocks", jacobi_v0<true>, 512, smem_v0},
        {"V2  512 thr, 4-row blocks, planar FFMA2", jacobi_v2, 512, smem_p},
        {"V8  V2 + apply phases of scheduler-mates alternate (mbarrier token)", jacobi_v8, 512, smem_p},
        {"V6  512 thr, 4-row blocks, planar FFMA2, fast (scaled) rotations", jacobi_v6, 512, smem_p + N * (int)sizeof(float)},
        {"V5  512 thr, 4-row blocks, planar FFMA2, Gram once per round (lite update)", jacobi_v5<false>, 512, smem_p},
        {"V5p the same with per-phase clocks (gram | reduce32 | 4 scalar sub-rounds | apply | load | store | barrier)", jacobi_v5<true>, 512, smem_p},
        {"V3 1024 thr, 2-row blocks, planar FFMA2, __syncthreads per round", jacobi_v3<0>, 1024, smem_p},
        {"V4 1024 thr, 2-row blocks, planar FFMA2, grouped named barriers", jacobi_v3<1>, 1024, smem_p},
    };
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const unsigned stag_list[] = {0u, 150u, 300u, 500u, 800u, 1200u};
    for (unsigned stag : stag_list) {
        cudaMemcpyToSymbol(g_stagger_ns, &stag, sizeof(unsigned));
        cudaFuncSetAttribute(jacobi_v7, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_p);
        float ms = 0;
        for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(e0);
            jacobi_v7<<<nmat, 512, smem_p>>>(dX, dY, dinfo, dclk, max_sweeps);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            cudaEventElapsedTime(&ms, e0, e1);
        }
        std::vector<int> info(nmat * 2); long long clk[256];
        cudaMemcpy(info.data(), dinfo, info.size() * sizeof(int), cudaMemcpyDeviceToHost);
        cudaMemcpy(clk, dclk, sizeof(clk), cudaMemcpyDeviceToHost);
        printf("V7  V2 + stagger %4u ns: %.3f ms per launch, CTA 0: %.0f cycles per sweep (%d sweeps)\n", stag, ms, (double)clk[0] / info[1], info[1]);
    }
    for (const Variant& v : vs) {
        cudaFuncSetAttribute(v.k, cudaFuncAttributeMaxDynamicSharedMemorySize, v.smem);
        float ms = 0;
        for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(e0);
            v.k<<<nmat, v.nt, v.smem>>>(dX, dY, dinfo, dclk, max_sweeps);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            cudaEventElapsedTime(&ms, e0, e1);
        }
        cudaError_t err = cudaDeviceSynchronize();
        if (err != cudaSuccess) { printf("%s: CUDA error: %s\n", v.name, cudaGetErrorString(err)); return 1; }
        std::vector<int> info(nmat * 2); long long clk[256];
        cudaMemcpy(info.data(), dinfo, info.size() * sizeof(int), cudaMemcpyDeviceToHost);
        cudaMemcpy(clk, dclk, sizeof(clk), cudaMemcpyDeviceToHost);
        std::vector<cf> y((size_t)N * N);
        cudaMemcpy(y.data(), dY, y.size() * sizeof(cf), cudaMemcpyDeviceToHost);
        double sw = 0; int bad = 0;
        for (int j = 0; j < nmat; ++j) { sw += info[2 * j + 1]; bad += info[2 * j] != 0; }
        std::vector<double> got(N);
        double maxdot = 0;
        for (int p = 0; p < N; ++p) { double a = 0; for (int c = 0; c < N; ++c) a += (double)y[p * N + c].x * y[p * N + c].x + (double)y[p * N + c].y * y[p * N + c].y; got[p] = std::sqrt(a); }
        for (int p = 0; p < N; ++p)
            for (int q = p + 1; q < N; ++q) {
                std::complex<double> g = 0;
                for (int c = 0; c < N; ++c) g += std::complex<double>(y[p * N + c].x, y[p * N + c].y) * std::conj(std::complex<double>(y[q * N + c].x, y[q * N + c].y));
                maxdot = std::max(maxdot, std::abs(g) / (got[p] * got[q] + 1e-300));
            }
        std::sort(got.begin(), got.end(), std::greater<double>());
        double maxerr = 0;
        for (int p = 0; p < N; ++p) maxerr = std::max(maxerr, std::fabs(got[p] - ref[p]) / ref[0]);
        printf("%s\n    %.3f ms per launch (%d matrices), mean sweeps %.2f, not converged %d, CTA 0: %.0f cycles per sweep; "
               "matrix 0: sigma err %.2e sigma_max, max |cos| %.2e\n",
               v.name, ms, nmat, sw / nmat, bad, (double)clk[0] / info[1], maxerr, maxdot);
        if (v.name[2] == 'p') {
            printf("    per-warp cycles of CTA 0 (gram | reduce | params | bcast+apply | load | store | barrier), total %lld:\n", clk[0]);
            for (int w : {0, 1, 5, 10, 15}) {
                printf("      warp %2d:", w);
                for (int i = 0; i < 7; ++i) printf(" %9lld", clk[8 + w * 8 + i]);
                printf("\n");
            }
        }
    }
    return 0;
}
