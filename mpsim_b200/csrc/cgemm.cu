// Batched complex64 GEMM with explicit element strides (split-real FFMA, shared-memory tiled).
//
// C[M][N] = op(A)[M][K] . op(B)[K][N]   (op = optional conjugation; transposition is expressed
// through the strides).  This is the building block of the whole-chain contractions that
// replace tensornetwork's contract_between on the norm / wavefunction paths
// (mpsim/core.py:489-500, 543-561) and of the large-chi block-Jacobi updates.  The theta
// contraction of the gate path has its own fused kernel (theta.cu).
#include "common.cuh"

namespace {

constexpr int BM = 64, BN = 64, BK = 16, GT = 256;

template <bool CONJ_A, bool CONJ_B>
__global__ void __launch_bounds__(GT)
cgemm_kernel(const cf* __restrict__ A, int64_t a_rs, int64_t a_cs, int64_t a_bs,
             const cf* __restrict__ B, int64_t b_rs, int64_t b_cs, int64_t b_bs,
             cf* __restrict__ C, int64_t c_ld, int64_t c_bs, int M, int N, int K) {
    __shared__ cf As[BK][BM + 1];
    __shared__ cf Bs[BK][BN + 1];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;       // 16 x 16 threads, 4 x 4 outputs each
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;      // M on grid.x: tall-skinny products (wavefunction chain) are not capped at 65 535 tiles
    A += (int64_t)blockIdx.z * a_bs;
    B += (int64_t)blockIdx.z * b_bs;
    C += (int64_t)blockIdx.z * c_bs;

    // Choose the load mapping so that consecutive threads walk the unit-stride direction.
    const bool a_k_fast = (a_cs == 1);   // A row-major: K contiguous
    const bool b_n_fast = (b_cs == 1);   // B row-major: N contiguous

    cf acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = cf_make(0.f, 0.f);

    for (int k0 = 0; k0 < K; k0 += BK) {
#pragma unroll
        for (int it = 0; it < (BM * BK) / GT; ++it) {
            int e = tid + it * GT;
            int mm, kk;
            if (a_k_fast) { kk = e % BK; mm = e / BK; } else { mm = e % BM; kk = e / BM; }
            int gm = m0 + mm, gk = k0 + kk;
            cf v = cf_make(0.f, 0.f);
            if (gm < M && gk < K) v = A[(int64_t)gm * a_rs + (int64_t)gk * a_cs];
            if (CONJ_A) v.y = -v.y;
            As[kk][mm] = v;
        }
#pragma unroll
        for (int it = 0; it < (BN * BK) / GT; ++it) {
            int e = tid + it * GT;
            int nn, kk;
            if (b_n_fast) { nn = e % BN; kk = e / BN; } else { kk = e % BK; nn = e / BK; }
            int gn = n0 + nn, gk = k0 + kk;
            cf v = cf_make(0.f, 0.f);
            if (gn < N && gk < K) v = B[(int64_t)gk * b_rs + (int64_t)gn * b_cs];
            if (CONJ_B) v.y = -v.y;
            Bs[kk][nn] = v;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            cf a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx + 16 * j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = cf_fma(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int gm = m0 + ty * 4 + i;
        if (gm >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int gn = n0 + tx + 16 * j;
            if (gn < N) C[(int64_t)gm * c_ld + gn] = acc[i][j];
        }
    }
}

}  // namespace

int launch_cgemm(const cf* A, int64_t a_rs, int64_t a_cs, int conj_a, int64_t a_bs,
                 const cf* B, int64_t b_rs, int64_t b_cs, int conj_b, int64_t b_bs,
                 cf* C, int64_t c_ld, int64_t c_bs, int M, int N, int K, int nbatch,
                 cudaStream_t st) {
    if (M <= 0 || N <= 0 || nbatch <= 0) return 0;
    dim3 grid((M + BM - 1) / BM, (N + BN - 1) / BN, nbatch);
    MPSB_ARG(grid.y <= 65535 && grid.z <= 65535, "cgemm: grid too large (N=%d, nbatch=%d)", N, nbatch);
#define GO(CA, CB) cgemm_kernel<CA, CB><<<grid, GT, 0, st>>>(A, a_rs, a_cs, a_bs, B, b_rs, b_cs, b_bs, C, c_ld, c_bs, M, N, K)
    if (conj_a && conj_b) GO(true, true);
    else if (conj_a) GO(true, false);
    else if (conj_b) GO(false, true);
    else GO(false, false);
#undef GO
    MPSB_LAUNCH_CHECK("cgemm_kernel");
    return 0;
}
