// Complex64 GEMM on the 5th-generation tensor cores (tcgen05, sm_100a) by 3xTF32 splitting, with
// the two-qudit gate folded into the epilogue: the theta contraction of mpsim/core.py:1060-1068
// for large bond dimensions (d = 2, chi >= 64).
//
//   T[(l,p)][(q,r)] = sum_m A[l][p][m] B[m][q][r]            (complex, M = 2 chiL, N = 2 chiR, K = chiM)
//   theta[(l,o1)][(o2,r)] = sum_{p,q} G[o1][o2][p][q] T[(l,p)][(q,r)]
//
// Complex product on a real tensor core: with A' = A viewed as real [M][2K] (re/im interleaved
// along K, exactly the memory layout of the site tensor) and the real matrix
//   Bt[(n,0)][(k,0)] =  Br[k][n]   Bt[(n,0)][(k,1)] = -Bi[k][n]
//   Bt[(n,1)][(k,0)] =  Bi[k][n]   Bt[(n,1)][(k,1)] =  Br[k][n]
// the real GEMM D = A' Bt^T has D[m][(n,0)] = Re C[m][n], D[m][(n,1)] = Im C[m][n]: the
// accumulator row IS the interleaved complex row.  8 M N K real flops, the same as the complex
// count, no wasted work.
// 3xTF32: every fp32 operand is split x = hi + lo (hi = x rounded to TF32, lo = x - hi, exact) by
// the operand-preparation kernels, which also build Bt (K-major, so that both operands use the
// canonical K-major SWIZZLE_128B shared-memory layout); the kernel accumulates
// hi*hi + hi*lo + lo*hi in the fp32 TMEM accumulator (the dropped lo*lo term is ~2^-22 relative).
// The tensor core truncates (round-toward-zero) once per accumulating MMA, which shrinks a long
// accumulation systematically: measured rms error 2.8e-7 / 9.5e-7 / 4.5e-6 / 8.3e-6 of max|C| at
// K = 64 / 256 / 1024 / 2048 when a whole K loop runs in one accumulator (scripts/tc_accuracy.py).
// The K loop is therefore cut into chunks of KCH k-blocks: each chunk accumulates in TMEM from zero
// and the epilogue warps add the chunks in fp32 registers with round-to-nearest.  Round 1 used chunks
// of 4 k-blocks (64 complex k, 48 MMAs): rms error fine, but the truncation is a systematic SHRINK --
// a 14-qubit GHZ + QFT circuit (about 700 applications, nothing truncated) lost 2.6e-4 of its norm,
// 3.7e-7 per application, against +3e-5 with the FFMA theta kernel.  With one k-block per chunk
// (16 complex k, 12 MMAs; the two TMEM accumulators still alternate, so draining chunk c overlaps the
// MMAs of chunk c+1) the loss is 3e-5, the noise level of the FFMA kernel; the price is one TMEM drain
// per k-block: theta at chi = 1024 1.42 -> 1.54 ms (0.76 -> 0.71 of the complex tensor-core peak), chi =
// 256 0.271 -> 0.300 ms (scripts/norm_bias_ghz_qft.py, bench.py; MPSB_TC_KCH=n overrides for A/B).
//
// Kernel structure (persistent, one CTA per SM, 640 threads):
//   warp 0   : TMA producer -- per k-block four boxes (A_hi, A_lo: 128 x 32 fp32; B_hi, B_lo:
//              256 x 32 fp32) land in a 2-stage ring of 96 KB stages, mbarrier complete_tx;
//   warp 1   : MMA issuer -- one lane issues 12 tcgen05.mma.kind::tf32 (M=128, N=256, K=8) per
//              stage (3 operand pairings x 4 k-steps), tcgen05.commit frees the stage;
//              (it also allocates the 512 TMEM columns = two 128 x 256 fp32 accumulators, so the
//              drain of chunk c overlaps the MMAs of chunk c+1);
//   warps 4-19: epilogue -- thread = accumulator row (l,p) x one 64-column quarter; per chunk
//              tcgen05.ld + fp32 add into 64 registers; at the end the partner row (l,p^1) comes
//              by shuffle, the gate is applied in registers and 128-byte row segments are stored
//              straight into the SVD input matrix (or its transpose).
#include "common.cuh"
#include "tc_ptx.cuh"
#include <cuda.h>

namespace {
using namespace tcx;

constexpr int BM = 128;          // accumulator rows per tile (TMEM lanes)
constexpr int BNR = 256;         // accumulator columns per tile = 128 complex columns
constexpr int BK = 32;           // fp32 per k-block = one 128-byte swizzle row = 16 complex k
constexpr int NSTAGE = 2;
constexpr int A_TILE_BYTES = BM * BK * 4;        // 16 KB
constexpr int B_TILE_BYTES = BNR * BK * 4;       // 32 KB
constexpr int STAGE_BYTES = 2 * A_TILE_BYTES + 2 * B_TILE_BYTES;     // 96 KB
constexpr int TC_THREADS = 640;        // producer, MMA, 2 idle warps (registers are allocated per 4 warps) + 16 epilogue warps
constexpr int CPT = 64;                // accumulator columns per epilogue thread
constexpr int KCH = 1;           // k-blocks per accumulation chunk (see the header: truncation bias)
constexpr int RB = 64;           // r values per tile (two column halves of 32, each with q = 0 | q = 1)
// kind::tf32, fp32 accumulate, A and B K-major, M = 128, N = 256
constexpr uint32_t IDESC_TF32 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BNR >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

struct TcParams {
    int njobs, nbatch;           // jobs = ndesc * nbatch
    int mtiles, ntiles, nkb;     // tiles per job, k-blocks
    int kch;                     // k-blocks per accumulation chunk
    int M, chiR;                 // valid rows (2 chiL) and r values (mode 1) / valid complex columns (mode 0)
    int mode;                    // 1: theta epilogue (gate, optional transpose); 0: plain complex C store
    int transpose_out;
    const mpsb_gate2_desc* descs;
    cf* out; int64_t out_job_stride; int64_t out_ld;
};

__global__ void __launch_bounds__(TC_THREADS, 1)
tc_cgemm_kernel(const __grid_constant__ CUtensorMap map_ahi, const __grid_constant__ CUtensorMap map_alo,
                const __grid_constant__ CUtensorMap map_bhi, const __grid_constant__ CUtensorMap map_blo, TcParams P) {
    extern __shared__ __align__(1024) uint8_t smem_dyn[];
    __shared__ __align__(8) uint64_t full_bar[NSTAGE], empty_bar[NSTAGE], acc_full[2], acc_empty[2];
    __shared__ uint32_t tmem_slot;
    // dynamic shared memory is only guaranteed 16-byte aligned: align the ring to 1024 by hand
    uint8_t* ring = (uint8_t*)(((uintptr_t)smem_dyn + 1023) & ~(uintptr_t)1023);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ntile_total = P.njobs * P.mtiles * P.ntiles;

    if (warp == 1 && lane == 0) {
        for (int s = 0; s < NSTAGE; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&acc_full[a], 1); mbar_init(&acc_empty[a], 16); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < ntile_total; tile += gridDim.x) {
                const int job = tile / (P.mtiles * P.ntiles), rem = tile % (P.mtiles * P.ntiles);
                const int mt = rem / P.ntiles, nt = rem % P.ntiles;
                for (int kb = 0; kb < P.nkb; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t* st = ring + (size_t)stage * STAGE_BYTES;
                    mbar_expect_tx(&full_bar[stage], STAGE_BYTES);
                    tma_load_3d(st, &map_ahi, &full_bar[stage], kb * BK, mt * BM, job);
                    tma_load_3d(st + A_TILE_BYTES, &map_alo, &full_bar[stage], kb * BK, mt * BM, job);
                    tma_load_3d(st + 2 * A_TILE_BYTES, &map_bhi, &full_bar[stage], kb * BK, nt * BNR, job);
                    tma_load_3d(st + 2 * A_TILE_BYTES + B_TILE_BYTES, &map_blo, &full_bar[stage], kb * BK, nt * BNR, job);
                    if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            int acc = 0; uint32_t acc_phase = 0;
            for (int tile = blockIdx.x; tile < ntile_total; tile += gridDim.x) {
                for (int kb0 = 0; kb0 < P.nkb; kb0 += P.kch) {
                    mbar_wait(&acc_empty[acc], acc_phase ^ 1);
                    tc_fence_after();
                    const uint32_t tmem_d = tmem_base + (uint32_t)acc * BNR;
                    const int kb1 = min(kb0 + P.kch, P.nkb);
                    for (int kb = kb0; kb < kb1; ++kb) {
                        mbar_wait(&full_bar[stage], phase);
                        tc_fence_after();
                        const uint32_t sa = smem_u32(ring + (size_t)stage * STAGE_BYTES);
                        const uint64_t ahi = umma_desc_k128(sa), alo = umma_desc_k128(sa + A_TILE_BYTES);
                        const uint64_t bhi = umma_desc_k128(sa + 2 * A_TILE_BYTES), blo = umma_desc_k128(sa + 2 * A_TILE_BYTES + B_TILE_BYTES);
#pragma unroll
                        for (int k = 0; k < BK / 8; ++k) {
                            const uint64_t adv = (uint64_t)((k * 32) >> 4);      // 8 tf32 = 32 bytes along K inside the swizzle row
                            tc_mma_tf32(tmem_d, ahi + adv, blo + adv, IDESC_TF32, ((kb - kb0) | k) != 0);
                            tc_mma_tf32(tmem_d, alo + adv, bhi + adv, IDESC_TF32, 1);
                            tc_mma_tf32(tmem_d, ahi + adv, bhi + adv, IDESC_TF32, 1);
                        }
                        tc_commit(&empty_bar[stage]);
                        if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
                    }
                    tc_commit(&acc_full[acc]);
                    if (++acc == 2) { acc = 0; acc_phase ^= 1; }
                }
            }
        }
    } else if (warp >= 4) {
        // ===== epilogue =====
        const int wq = warp & 3;                      // TMEM lane quarter this warp may read
        const int cq = (warp - 4) >> 2;               // which 64 accumulator columns
        int acc = 0; uint32_t acc_phase = 0;
        for (int tile = blockIdx.x; tile < ntile_total; tile += gridDim.x) {
            const int job = tile / (P.mtiles * P.ntiles), rem = tile % (P.mtiles * P.ntiles);
            const int mt = rem / P.ntiles, nt = rem % P.ntiles;
            const int row = mt * BM + wq * 32 + lane;               // GEMM row = 2 l + p
            const bool row_ok = row < P.M;
            float sum[CPT];
#pragma unroll
            for (int i = 0; i < CPT; ++i) sum[i] = 0.f;
            for (int kb0 = 0; kb0 < P.nkb; kb0 += P.kch) {
                mbar_wait(&acc_full[acc], acc_phase);
                tc_fence_after();
                const uint32_t trow = tmem_base + (uint32_t)acc * BNR + (uint32_t)cq * CPT + ((uint32_t)(wq * 32) << 16);
#pragma unroll
                for (int c = 0; c < CPT / 16; ++c) {
                    float v[16];
                    tc_ld16(trow + c * 16, v);
                    tc_ld_wait();
#pragma unroll
                    for (int i = 0; i < 16; ++i) sum[c * 16 + i] += v[i];
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&acc_empty[acc]);
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
            cf* O = P.out + (int64_t)job * P.out_job_stride;
            if (P.mode == 1) {
                // columns of this quarter: [q = 0: 16 r x (re, im)][q = 1: 16 r x (re, im)]
                const int di = job / P.nbatch, bi = job % P.nbatch;
                const mpsb_gate2_desc dsc = P.descs[di];
                const cf* G = (const cf*)dsc.gate + (int64_t)bi * dsc.bs_gate;
                const int p = row & 1;
                cf g[2][2][2];                                          // G[o1 = p][o2][p'][q]
#pragma unroll
                for (int o2 = 0; o2 < 2; ++o2)
#pragma unroll
                    for (int pp = 0; pp < 2; ++pp)
#pragma unroll
                        for (int q = 0; q < 2; ++q) g[o2][pp][q] = G[((p * 2 + o2) * 2 + pp) * 2 + q];
                const int mrows = P.M, ncols = 2 * P.chiR;
#pragma unroll
                for (int rc = 0; rc < 4; ++rc) {                    // 4 r values per step
                    float w0[8], w1[8];                             // the partner row (l, p ^ 1)
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        w0[i] = __shfl_xor_sync(0xffffffffu, sum[rc * 8 + i], 1);
                        w1[i] = __shfl_xor_sync(0xffffffffu, sum[32 + rc * 8 + i], 1);
                    }
                    const int rbase = nt * RB + cq * 16 + rc * 4;
#pragma unroll
                    for (int o2 = 0; o2 < 2; ++o2) {
                        cf o[4];
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            // T[p'][q][r]: p' == p is this thread's row, p' != p the partner's
                            const cf own0 = cf_make(sum[rc * 8 + 2 * i], sum[rc * 8 + 2 * i + 1]);
                            const cf own1 = cf_make(sum[32 + rc * 8 + 2 * i], sum[32 + rc * 8 + 2 * i + 1]);
                            const cf oth0 = cf_make(w0[2 * i], w0[2 * i + 1]), oth1 = cf_make(w1[2 * i], w1[2 * i + 1]);
                            const cf t00 = p ? oth0 : own0, t01 = p ? oth1 : own1;     // p' = 0, q = 0 | 1
                            const cf t10 = p ? own0 : oth0, t11 = p ? own1 : oth1;     // p' = 1
                            cf s = cf_mul(g[o2][0][0], t00);
                            s = cf_fma(g[o2][0][1], t01, s);
                            s = cf_fma(g[o2][1][0], t10, s);
                            s = cf_fma(g[o2][1][1], t11, s);
                            o[i] = s;
                        }
                        if (row_ok) {
                            if (!P.transpose_out) {
                                cf* dst = O + (int64_t)row * ncols + (int64_t)o2 * P.chiR + rbase;
                                if (rbase + 4 <= P.chiR && (P.chiR & 1) == 0) {
#pragma unroll
                                    for (int i = 0; i < 4; i += 2)
                                        *reinterpret_cast<float4*>(dst + i) = make_float4(o[i].x, o[i].y, o[i + 1].x, o[i + 1].y);
                                } else {
#pragma unroll
                                    for (int i = 0; i < 4; ++i) if (rbase + i < P.chiR) dst[i] = o[i];
                                }
                            } else {
#pragma unroll
                                for (int i = 0; i < 4; ++i)
                                    if (rbase + i < P.chiR) O[((int64_t)o2 * P.chiR + rbase + i) * mrows + row] = o[i];
                            }
                        }
                    }
                }
            } else {
                const int nbase = nt * (BNR / 2) + cq * (CPT / 2);
                if (row_ok) {
#pragma unroll
                    for (int i = 0; i < CPT / 2; ++i)
                        if (nbase + i < P.chiR) O[(int64_t)row * P.out_ld + nbase + i] = cf_make(sum[2 * i], sum[2 * i + 1]);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

// ---- operand preparation ----------------------------------------------------------------------

struct PrepParams {
    const mpsb_gate2_desc* descs; int nbatch;     // mode 1: operands come from the descriptor table
    const cf* A; int64_t a_bs; const cf* B; int64_t b_bs;     // mode 0: dense A [M][K], B [K][N]
    int mode;
    int Mc, K, Nc;               // complex sizes: A is Mc x K, B is K x Nc
    int Mp, Kp, Np;              // padded real sizes: A' [Mp][Kp], Bt [Np][Kp]
    int chiR;                    // mode 1: Nc = 2 chiR, column n = q chiR + r goes to the q-interleaved tile order
    float* ahi; float* alo; float* bhi; float* blo;
};

// A' = A as real [Mp][Kp], split.  One thread per complex element (two floats of a row).
__global__ void tc_prep_a_kernel(PrepParams P) {
    const int job = blockIdx.z;
    const cf* A;
    if (P.mode == 1) {
        const mpsb_gate2_desc d = P.descs[job / P.nbatch];
        A = (const cf*)d.site_l + (int64_t)(job % P.nbatch) * d.bs_site_l;
    } else A = P.A + (int64_t)job * P.a_bs;
    const int kc = blockIdx.x * blockDim.x + threadIdx.x;      // complex k
    const int m = blockIdx.y;
    if (2 * kc >= P.Kp) return;
    cf v = cf_make(0.f, 0.f);
    if (m < P.Mc && kc < P.K) v = A[(int64_t)m * P.K + kc];
    float2 hi, lo;
    split_tf32(v.x, hi.x, lo.x);
    split_tf32(v.y, hi.y, lo.y);
    const size_t o = ((size_t)job * P.Mp + m) * P.Kp + 2 * kc;
    *reinterpret_cast<float2*>(P.ahi + o) = hi;
    *reinterpret_cast<float2*>(P.alo + o) = lo;
}

// Bt [Np][Kp] from B [K][Nc] through a 32 x 32 shared-memory transpose, split.  Rows of Bt that
// belong to padding columns are not written here (the caller clears the buffer when there are any).
__global__ void tc_prep_b_kernel(PrepParams P) {
    __shared__ cf tile[32][33];
    const int job = blockIdx.z;
    const cf* B;
    if (P.mode == 1) {
        const mpsb_gate2_desc d = P.descs[job / P.nbatch];
        B = (const cf*)d.site_r + (int64_t)(job % P.nbatch) * d.bs_site_r;
    } else B = P.B + (int64_t)job * P.b_bs;
    const int n0 = blockIdx.x * 32, k0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int k = k0 + i, n = n0 + threadIdx.x;
        tile[i][threadIdx.x] = (k < P.K && n < P.Nc) ? B[(int64_t)k * P.Nc + n] : cf_make(0.f, 0.f);
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int n = n0 + i, kc = k0 + threadIdx.x;
        if (n >= P.Nc || 2 * kc >= P.Kp) continue;
        int npos = n;                              // complex column position in the tile order
        if (P.mode == 1) {                         // n = q chiR + r  ->  tile of 64 r: [quarter][q][16 r]
            const int q = n / P.chiR, r = n % P.chiR;
            npos = (r / RB) * (2 * RB) + ((r % RB) / 16) * 32 + q * 16 + (r % 16);
        }
        const cf v = tile[threadIdx.x][i];
        float hr, lr, hi_, li;
        split_tf32(v.x, hr, lr);
        split_tf32(v.y, hi_, li);
        // row (n,0): (Br, -Bi)   row (n,1): (Bi, Br)
        const size_t o0 = ((size_t)job * P.Np + 2 * npos) * P.Kp + 2 * kc, o1 = o0 + P.Kp;
        *reinterpret_cast<float2*>(P.bhi + o0) = make_float2(hr, -hi_);
        *reinterpret_cast<float2*>(P.blo + o0) = make_float2(lr, -li);
        *reinterpret_cast<float2*>(P.bhi + o1) = make_float2(hi_, hr);
        *reinterpret_cast<float2*>(P.blo + o1) = make_float2(li, lr);
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

// fp32 tensor [njobs][rows][Kp], box [1][box_rows][32], 128-byte swizzle
int make_map(CUtensorMap* map, float* base, int njobs, int rows, int Kp, int box_rows) {
    EncodeTiledFn enc = get_encode_fn();
    MPSB_ARG(enc != nullptr, "tc_gemm: cuTensorMapEncodeTiled is not available from the driver");
    cuuint64_t dims[3] = {(cuuint64_t)Kp, (cuuint64_t)rows, (cuuint64_t)njobs};
    cuuint64_t strides[2] = {(cuuint64_t)Kp * 4, (cuuint64_t)Kp * 4 * (cuuint64_t)rows};
    cuuint32_t box[3] = {(cuuint32_t)BK, (cuuint32_t)box_rows, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    MPSB_ARG(r == CUDA_SUCCESS, "tc_gemm: cuTensorMapEncodeTiled failed with %d", (int)r);
    return 0;
}

struct TcLayout { int Mp, Kp, Np, chiRp, mtiles, ntiles, nkb; size_t a_floats, b_floats; };

// mode 1: Mc = 2 chiL, K = chiM, r range chiR;  mode 0: Mc x K times K x Nc
TcLayout tc_layout(int mode, int Mc, int K, int Nc_or_chiR) {
    TcLayout lo;
    lo.Mp = (Mc + BM - 1) / BM * BM;
    lo.Kp = (2 * K + BK - 1) / BK * BK;
    if (mode == 1) {
        lo.chiRp = (Nc_or_chiR + RB - 1) / RB * RB;
        lo.Np = 4 * lo.chiRp;
    } else {
        lo.chiRp = 0;
        lo.Np = (2 * Nc_or_chiR + BNR - 1) / BNR * BNR;
    }
    lo.mtiles = lo.Mp / BM; lo.ntiles = lo.Np / BNR; lo.nkb = lo.Kp / BK;
    lo.a_floats = (size_t)lo.Mp * lo.Kp; lo.b_floats = (size_t)lo.Np * lo.Kp;
    return lo;
}

int run_tc(int mode, int njobs, int nbatch, const mpsb_gate2_desc* descs, const cf* A, int64_t a_bs, const cf* B, int64_t b_bs,
           int Mc, int K, int Nc_or_chiR, int transpose_out, cf* out, int64_t out_job_stride, int64_t out_ld,
           float* work, cudaStream_t st) {
    const TcLayout lo = tc_layout(mode, Mc, K, Nc_or_chiR);
    float* ahi = work; float* alo = ahi + lo.a_floats * njobs;
    float* bhi = alo + lo.a_floats * njobs; float* blo = bhi + lo.b_floats * njobs;
    PrepParams pp;
    pp.descs = descs; pp.nbatch = nbatch > 0 ? nbatch : 1; pp.A = A; pp.a_bs = a_bs; pp.B = B; pp.b_bs = b_bs; pp.mode = mode;
    pp.Mc = Mc; pp.K = K; pp.Nc = mode == 1 ? 2 * Nc_or_chiR : Nc_or_chiR;
    pp.Mp = lo.Mp; pp.Kp = lo.Kp; pp.Np = lo.Np; pp.chiR = Nc_or_chiR;
    pp.ahi = ahi; pp.alo = alo; pp.bhi = bhi; pp.blo = blo;
    MPSB_ARG(njobs <= 65535, "tc_gemm: njobs %d > 65535", njobs);
    {
        dim3 grid((lo.Kp / 2 + 127) / 128, lo.Mp, njobs);
        tc_prep_a_kernel<<<grid, 128, 0, st>>>(pp);
        MPSB_LAUNCH_CHECK("tc_prep_a_kernel");
    }
    if (2 * pp.Nc != lo.Np) {
        // rows of Bt for padding columns are never written by the transpose kernel below: clear them
        MPSB_CUDA(cudaMemsetAsync(bhi, 0, 2 * lo.b_floats * njobs * sizeof(float), st));
    }
    {
        dim3 grid((pp.Nc + 31) / 32, (lo.Kp / 2 + 31) / 32, njobs);
        tc_prep_b_kernel<<<grid, dim3(32, 8), 0, st>>>(pp);
        MPSB_LAUNCH_CHECK("tc_prep_b_kernel");
    }
    CUtensorMap mahi, malo, mbhi, mblo;
    int rc;
    if ((rc = make_map(&mahi, ahi, njobs, lo.Mp, lo.Kp, BM))) return rc;
    if ((rc = make_map(&malo, alo, njobs, lo.Mp, lo.Kp, BM))) return rc;
    if ((rc = make_map(&mbhi, bhi, njobs, lo.Np, lo.Kp, BNR))) return rc;
    if ((rc = make_map(&mblo, blo, njobs, lo.Np, lo.Kp, BNR))) return rc;
    TcParams P;
    P.njobs = njobs; P.nbatch = nbatch > 0 ? nbatch : 1;
    P.mtiles = lo.mtiles; P.ntiles = lo.ntiles; P.nkb = lo.nkb;
    P.kch = KCH;
    if (const char* e = mpsb_env("MPSB_TC_KCH")) { const int v = atoi(e); if (v >= 1 && v <= 64) P.kch = v; }
    P.M = Mc; P.chiR = Nc_or_chiR; P.mode = mode; P.transpose_out = transpose_out;
    P.descs = descs; P.out = out; P.out_job_stride = out_job_stride; P.out_ld = out_ld;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int ntile_total = njobs * lo.mtiles * lo.ntiles;
    const int grid = ntile_total < sms ? ntile_total : sms;
    const size_t smem = (size_t)NSTAGE * STAGE_BYTES + 1024;
    MPSB_CUDA(cudaFuncSetAttribute(tc_cgemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    tc_cgemm_kernel<<<grid, TC_THREADS, smem, st>>>(mahi, malo, mbhi, mblo, P);
    MPSB_LAUNCH_CHECK("tc_cgemm_kernel");
    return 0;
}

}  // namespace

size_t tc_theta_workspace_floats(int njobs, int chiL, int chiM, int chiR) {
    const TcLayout lo = tc_layout(1, 2 * chiL, chiM, chiR);
    return 2 * (lo.a_floats + lo.b_floats) * (size_t)njobs;
}

size_t tc_cgemm_workspace_floats(int njobs, int M, int N, int K) {
    const TcLayout lo = tc_layout(0, M, K, N);
    return 2 * (lo.a_floats + lo.b_floats) * (size_t)njobs;
}

// theta for d = 2 on the tensor cores; same output contract as launch_theta
int launch_theta_tc(const mpsb_gate2_desc* descs, int ndesc, int nbatch, int chiL, int chiM, int chiR,
                    int transpose_out, cf* out, int64_t out_job_stride, float* work, cudaStream_t st) {
    const int njobs = ndesc * nbatch;
    if (njobs <= 0 || chiL <= 0 || chiR <= 0) return 0;
    MPSB_ARG(work != nullptr, "theta_tc: workspace missing");
    return run_tc(1, njobs, nbatch, descs, nullptr, 0, nullptr, 0, 2 * chiL, chiM, chiR, transpose_out, out,
                  out_job_stride, 0, work, st);
}

// C [M][N] = A [M][K] . B [K][N], dense row-major complex64, batched
int launch_cgemm_tc(const cf* A, int64_t a_bs, const cf* B, int64_t b_bs, cf* C, int64_t c_ld, int64_t c_bs,
                    int M, int N, int K, int nbatch, float* work, cudaStream_t st) {
    if (nbatch <= 0 || M <= 0 || N <= 0) return 0;
    MPSB_ARG(work != nullptr, "cgemm_tc: workspace missing");
    return run_tc(0, nbatch, 1, nullptr, A, a_bs, B, b_bs, M, K, N, 0, C, c_bs, c_ld, work, st);
}
