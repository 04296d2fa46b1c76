// Micro-benchmark for DESIGN.md section 9, lead 1: the sweep engine of svd_small.cu with 1 024 threads.
//
// svd_small_kernel runs one-sided Jacobi sweeps on a 128 x 128 complex64 matrix in shared memory
// with 16 warps, each holding a pair of 4-row blocks (8 rows) in registers; ncu shows the issue slots
// 59 % active: the four warps of a scheduler are in the same latency phase (shuffle reduction,
// MUFU rotation parameters) at the same time.  This program measures the same sweep with 32 warps
// of 2-row blocks (4 rows per warp, 32 data registers per thread, 8 warps per scheduler): same
// rotation formulas, same circle-method tournament (on 64 blocks: 63 rounds of 2 cross sub-rounds
// of 2 rotations per warp), block-wide barrier per round (only 16 named barriers exist, the
// production kernel's pairwise ones do not scale to 32 groups).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o jacobi1024 scripts/micro/jacobi1024.cu
//   ./jacobi1024 [nmat=148] [max_sweeps=30]
// Prints: ms per launch, sweeps, cycles per sweep of CTA 0 (production kernel: ~345 k cycles per
// sweep, scripts/prof_svd.py), and the check against a float64 Jacobi on the host (singular values,
// row orthogonality).  Not part of the product library.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <complex>
#include <algorithm>
#include <cuda_runtime.h>

typedef float2 cf;
constexpr int N = 128;           // rows = columns
constexpr int LS = N + 1;        // shared-memory row stride (elements)
constexpr int NT = 1024, NW = NT / 32;
constexpr int EPL = N / 32;      // elements per lane per row
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

__device__ __forceinline__ void rot_apply(float c, float sr, float si, cf& p, cf& q) {
    cf np_, nq_;
    np_.x = fmaf(c, p.x, fmaf(sr, q.x, -(si * q.y)));
    np_.y = fmaf(c, p.y, fmaf(sr, q.y, si * q.x));
    nq_.x = fmaf(c, q.x, -fmaf(sr, p.x, si * p.y));
    nq_.y = fmaf(c, q.y, fmaf(si, p.x, -(sr * p.y)));
    p = np_;
    q = nq_;
}

// One sub-round on the 4 rows of a warp: 2 disjoint pairs (A0, B0), (A1, B1).  The four partial sums
// (re, im of both Gram entries) are reduced transposed: after the 16-step a half-warp owns one
// pair, after the 8-step a quarter owns its re or im part; 6 shuffles + 1 to fetch the other part.
template <int A0, int B0, int A1, int B1>
__device__ __forceinline__ int sub_round2(cf (&y)[4][EPL], float (&a)[4], int lane, bool& big) {
    constexpr int PA[2] = {A0, A1}, PB[2] = {B0, B1};
    float gr[2], gi[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        float r = 0.f, m = 0.f;
#pragma unroll
        for (int t = 0; t < EPL; ++t) {
            cf p = y[PA[i]][t], q = y[PB[i]][t];
            r = fmaf(p.x, q.x, r); r = fmaf(p.y, q.y, r);
            m = fmaf(p.y, q.x, m); m = fmaf(-p.x, q.y, m);
        }
        gr[i] = r; gi[i] = m;
    }
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
#pragma unroll
            for (int t = 0; t < EPL; ++t) rot_apply(ci, sri, sii, y[PA[i]][t], y[PB[i]][t]);
            a[PA[i]] = fmaxf(a[PA[i]] + tgi, 0.f);
            a[PB[i]] = fmaxf(a[PB[i]] - tgi, 0.f);
        }
    }
    return __popc(flags);
}

__global__ void __launch_bounds__(NT, 1) jacobi1024_kernel(const cf* X, cf* Yout, int* info, long long* clk, int max_sweeps) {
    extern __shared__ float4 smem_raw[];
    cf* Ys = (cf*)smem_raw;                       // [N][LS]
    float* nrm = (float*)(Ys + N * LS);           // [N]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const cf* Xj = X + (size_t)blockIdx.x * N * N;
    for (int e = tid; e < N * N; e += NT) Ys[(e / N) * LS + (e % N)] = Xj[e];
    __syncthreads();
    constexpr int nb = N / 2, mcirc = nb - 1, nrounds = nb - 1;      // 64 two-row blocks, 32 groups = NW
    int sweeps = 0, status = 1;
    long long t0 = clock64(), tfirst = 0;
    for (int sweep = 0; sweep < max_sweeps; ++sweep) {
        for (int i = warp; i < N; i += NW) {
            float s2 = 0.f;
#pragma unroll
            for (int t = 0; t < EPL; ++t) { cf v = Ys[i * LS + lane + 32 * t]; s2 = fmaf(v.x, v.x, fmaf(v.y, v.y, s2)); }
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
            cf* rowA = Ys + (2 * I) * LS + lane;
            cf* rowB = Ys + (2 * J) * LS + lane;
            cf v[4][EPL];
            float a[4];
#pragma unroll
            for (int i = 0; i < 2; ++i) {
#pragma unroll
                for (int t = 0; t < EPL; ++t) { v[i][t] = rowA[i * LS + 32 * t]; v[2 + i][t] = rowB[i * LS + 32 * t]; }
                a[i] = nrm[2 * I + i];
                a[2 + i] = nrm[2 * J + i];
            }
            int nrot = 0;
            if (r == 0) nrot += sub_round2<0, 1, 2, 3>(v, a, lane, big);      // inside the two blocks
            nrot += sub_round2<0, 2, 1, 3>(v, a, lane, big);
            nrot += sub_round2<0, 3, 1, 2>(v, a, lane, big);
            if (nrot) {
#pragma unroll
                for (int i = 0; i < 2; ++i) {
#pragma unroll
                    for (int t = 0; t < EPL; ++t) { rowA[i * LS + 32 * t] = v[i][t]; rowB[i * LS + 32 * t] = v[2 + i][t]; }
                }
                float am = a[0];
#pragma unroll
                for (int i = 1; i < 4; ++i) if (lane == i) am = a[i];
                if (lane < 4) nrm[lane < 2 ? 2 * I + lane : 2 * J + lane - 2] = am;
            }
            __syncthreads();
        }
        sweeps = sweep + 1;
        if (sweep == 0 && blockIdx.x == 0 && tid == 0) tfirst = clock64() - t0;
        if (!__syncthreads_or(big ? 1 : 0)) { status = 0; break; }
    }
    long long t1 = clock64();
    cf* Yj = Yout + (size_t)blockIdx.x * N * N;
    for (int e = tid; e < N * N; e += NT) Yj[e] = Ys[(e / N) * LS + (e % N)];
    if (tid == 0) {
        info[2 * blockIdx.x] = status; info[2 * blockIdx.x + 1] = sweeps;
        if (blockIdx.x == 0) { clk[0] = t1 - t0; clk[1] = tfirst; }
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

int main(int argc, char** argv) {
    const int nmat = argc > 1 ? atoi(argv[1]) : 148, max_sweeps = argc > 2 ? atoi(argv[2]) : 30;
    std::vector<cf> h((size_t)nmat * N * N);
    srand(7);
    for (auto& v : h) { v.x = (rand() / (float)RAND_MAX - 0.5f); v.y = (rand() / (float)RAND_MAX - 0.5f); }
    cf *dX, *dY; int* dinfo; long long* dclk;
    cudaMalloc(&dX, h.size() * sizeof(cf)); cudaMalloc(&dY, h.size() * sizeof(cf));
    cudaMalloc(&dinfo, nmat * 2 * sizeof(int)); cudaMalloc(&dclk, 2 * sizeof(long long));
    cudaMemcpy(dX, h.data(), h.size() * sizeof(cf), cudaMemcpyHostToDevice);
    const int smem = (N * LS) * (int)sizeof(cf) + N * (int)sizeof(float);
    cudaFuncSetAttribute(jacobi1024_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms = 0;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        jacobi1024_kernel<<<nmat, NT, smem>>>(dX, dY, dinfo, dclk, max_sweeps);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
    }
    cudaError_t err = cudaDeviceSynchronize();
    if (err != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(err)); return 1; }
    std::vector<int> info(nmat * 2); long long clk[2];
    cudaMemcpy(info.data(), dinfo, info.size() * sizeof(int), cudaMemcpyDeviceToHost);
    cudaMemcpy(clk, dclk, sizeof(clk), cudaMemcpyDeviceToHost);
    std::vector<cf> y((size_t)N * N);
    cudaMemcpy(y.data(), dY, y.size() * sizeof(cf), cudaMemcpyDeviceToHost);
    double sw = 0; int bad = 0;
    for (int j = 0; j < nmat; ++j) { sw += info[2 * j + 1]; bad += info[2 * j] != 0; }
    printf("%d matrices of %d x %d, %d threads per CTA: %.3f ms per launch, mean sweeps %.2f, not converged %d\n", nmat, N, N, NT, ms, sw / nmat, bad);
    printf("CTA 0: %lld cycles in %d sweeps = %.0f cycles per sweep (first, all-rotating sweep: %lld); production kernel (16 warps of 8 rows): ~345 000\n",
           clk[0], info[1], (double)clk[0] / info[1], clk[1]);
    // check matrix 0
    std::vector<std::complex<double>> M(N * N);
    for (int e = 0; e < N * N; ++e) M[e] = std::complex<double>(h[e].x, h[e].y);
    std::vector<double> ref = host_singular_values(M), got(N);
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
    printf("matrix 0: max |sigma - sigma_ref| / sigma_max = %.2e, max |cos(row_p, row_q)| = %.2e\n", maxerr, maxdot);
    return 0;
}
