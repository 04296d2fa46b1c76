// Theta contraction with the two-qudit gate folded into the epilogue.
//
//   theta[(l,o1)][(o2,r)] = sum_{p,q} G[o1][o2][p][q] * sum_m A[l][p][m] * B[m][q][r]
//
// replaces tn.contract_between(A_i, A_j) + tn.flatten_edges_between + tn.contract
// (mpsim/core.py:1060-1068) and the row/column grouping of core.py:1095-1102, with
// rows = (left bond, gate edge 0) -- a row permutation of the reference's (gate edge 0, left
// bond) that leaves singular values and V unchanged and makes U land directly in the
// [chiL][d][k] site layout.
//
// One GEMM (d*chiL x chiM) . (chiM x d*chiR); each thread owns 2 l-values x 2 r-values x all
// d*d (p,q) combinations, so the d^2 x d^2 gate is applied in registers before the store.
// Split-real FFMA tiles staged through shared memory (round 1; the tcgen05/TMA variant is
// the next step for this kernel).
#include "common.cuh"

namespace {

constexpr int TL = 32;   // l-values per tile
constexpr int TR = 32;   // r-values per tile
constexpr int TK = 16;   // m-values per stage
constexpr int TT = 256;  // threads

template <int D>
__global__ void __launch_bounds__(TT)
theta_kernel(const mpsb_gate2_desc* __restrict__ descs, int nbatch, int chiL, int chiM, int chiR,
             int transpose_out, cf* __restrict__ out, int64_t out_job_stride) {
    __shared__ cf As[TK][TL * D + 1];
    __shared__ cf Bs[TK][TR * D + 1];
    __shared__ cf Gs[D * D * D * D];

    const int job = blockIdx.z;
    const int di = job / nbatch, bi = job % nbatch;
    const mpsb_gate2_desc dsc = descs[di];
    const cf* __restrict__ A = (const cf*)dsc.site_l + (int64_t)bi * dsc.bs_site_l;
    const cf* __restrict__ B = (const cf*)dsc.site_r + (int64_t)bi * dsc.bs_site_r;
    const cf* __restrict__ G = (const cf*)dsc.gate + (int64_t)bi * dsc.bs_gate;
    cf* __restrict__ O = out + (int64_t)job * out_job_stride;

    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int l0 = blockIdx.y * TL, r0 = blockIdx.x * TR;
    const int mrows = chiL * D;      // rows of A as a matrix, also rows of theta
    const int ncols = chiR * D;      // cols of theta

    if (tid < D * D * D * D) Gs[tid] = G[tid];
    __syncthreads();

    cf acc[2][2][D][D];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int p = 0; p < D; ++p)
#pragma unroll
                for (int q = 0; q < D; ++q) acc[i][j][p][q] = cf_make(0.f, 0.f);

    for (int k0 = 0; k0 < chiM; k0 += TK) {
        // A tile: rows (l,p) = l0*D .. (l0+TL)*D, K contiguous in memory
        for (int e = tid; e < TL * D * TK; e += TT) {
            int kk = e % TK, rr = e / TK;
            int grow = l0 * D + rr, gk = k0 + kk;
            cf v = cf_make(0.f, 0.f);
            if (grow < mrows && gk < chiM) v = A[(int64_t)grow * chiM + gk];
            As[kk][rr] = v;
        }
        // B tile: row m, columns (q, r) with r contiguous in memory
        for (int e = tid; e < TR * D * TK; e += TT) {
            int rr = e % TR, t = e / TR;
            int q = t % D, kk = t / D;
            int gr = r0 + rr, gk = k0 + kk;
            cf v = cf_make(0.f, 0.f);
            if (gr < chiR && gk < chiM) v = B[(int64_t)gk * ncols + (int64_t)q * chiR + gr];
            Bs[kk][q * TR + rr] = v;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < TK; ++kk) {
            cf a[2][D], b[D][2];
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int p = 0; p < D; ++p) a[i][p] = As[kk][(ty * 2 + i) * D + p];
#pragma unroll
            for (int q = 0; q < D; ++q)
#pragma unroll
                for (int j = 0; j < 2; ++j) b[q][j] = Bs[kk][q * TR + tx + 16 * j];
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int j = 0; j < 2; ++j)
#pragma unroll
                    for (int p = 0; p < D; ++p)
#pragma unroll
                        for (int q = 0; q < D; ++q)
                            acc[i][j][p][q] = cf_fma(a[i][p], b[q][j], acc[i][j][p][q]);
        }
        __syncthreads();
    }

    // epilogue: apply the gate in registers and store
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        int l = l0 + ty * 2 + i;
        if (l >= chiL) continue;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            int r = r0 + tx + 16 * j;
            if (r >= chiR) continue;
#pragma unroll
            for (int o1 = 0; o1 < D; ++o1)
#pragma unroll
                for (int o2 = 0; o2 < D; ++o2) {
                    cf t = cf_make(0.f, 0.f);
#pragma unroll
                    for (int p = 0; p < D; ++p)
#pragma unroll
                        for (int q = 0; q < D; ++q)
                            t = cf_fma(Gs[((o1 * D + o2) * D + p) * D + q], acc[i][j][p][q], t);
                    int row = l * D + o1, col = o2 * chiR + r;
                    if (transpose_out) O[(int64_t)col * mrows + row] = t;
                    else O[(int64_t)row * ncols + col] = t;
                }
        }
    }
}

}  // namespace

int launch_theta(const mpsb_gate2_desc* descs, int ndesc, int nbatch, int d, int chiL, int chiM,
                 int chiR, int transpose_out, cf* out, int64_t out_job_stride, cudaStream_t st) {
    int njobs = ndesc * nbatch;
    if (njobs <= 0 || chiL <= 0 || chiR <= 0) return 0;
    dim3 grid((chiR + TR - 1) / TR, (chiL + TL - 1) / TL, njobs);
    MPSB_ARG(grid.z <= 65535, "theta: too many jobs in one call (%d > 65535)", njobs);
    switch (d) {
        case 2: theta_kernel<2><<<grid, TT, 0, st>>>(descs, nbatch, chiL, chiM, chiR, transpose_out, out, out_job_stride); break;
        case 3: theta_kernel<3><<<grid, TT, 0, st>>>(descs, nbatch, chiL, chiM, chiR, transpose_out, out, out_job_stride); break;
        case 4: theta_kernel<4><<<grid, TT, 0, st>>>(descs, nbatch, chiL, chiM, chiR, transpose_out, out, out_job_stride); break;
        default: MPSB_ARG(false, "theta: qudit dimension %d not supported on device (2..4)", d);
    }
    MPSB_LAUNCH_CHECK("theta_kernel");
    return 0;
}
