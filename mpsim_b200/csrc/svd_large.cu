// Truncated SVD for d*chi > 128: block one-sided Jacobi with the matrices resident in HBM/L2.
//
// Same contract and the same split/absorb identity as svd_small.cu (rows of X are the vectors;
// Z X0 == X always; isometry = conj(Z), weighted factor = X), but the rotations are blocked so
// that the heavy steps are GEMM-shaped:
//   rows are grouped in blocks of B = P/2; block pairs follow the circle-method tournament;
//   for every pair of a round (all pairs of all jobs in ONE launch per step):
//     bj_gram_evd : G = Xp Xp^H  (P x L) . (L x P), then ONE cyclic pass of two-sided Hermitian
//                   Jacobi rotations on G in shared memory -> Q (P x P, near the identity)
//     bj_apply    : Xp <- Q Xp,  Zp <- Q Zp            (P x P) . (P x (L + nv)), in place
//   a job stops rotating (its launches become no-ops) after a sweep in which no pair rotated.
// Replaces tn.split_node_full_svd -> np.linalg.svd for 2chi in {256, 512, 2048}
// (mpsim/core.py:1132-1152).  The Gram step runs on FFMA tiles (bj_gram_evd_kernel) or, for rows of
// >= 1024 entries in full launches, on tcgen05 3xTF32 (bj_gram_tc_kernel, then the pass alone); the
// apply on tcgen05 3xTF32 for full launches -- rows by TMA into a raw ring and converter warps
// (bj_apply_tma_kernel), loader warps where the rows are not 16-byte aligned (bj_apply_tc_kernel) --
// and on FFMA tiles for one- and two-matrix launches (bj_apply_kernel).
#include "common.cuh"
#include "tc_ptx.cuh"
#include <vector>
#include <stdlib.h>
#include <sched.h>

namespace {

constexpr int P = 32;            // rows per block pair
constexpr int BLK = P / 2;       // rows per block
constexpr int LT = 256;          // threads
constexpr int CT = 128;          // columns per apply tile
constexpr int MAX_OUTER = 64;
constexpr float ABS_ETA = 3e-7f;     // see bj_gram_evd_kernel

struct Misc {                    // per job, lives in the workspace
    int active, rot, sweeps;
    unsigned gmax_next;          // max Gram diagonal seen in this sweep (float bits; values >= 0)
    float gmax;                  // the same from the previous sweep (~ sigma_max^2)
    int pad[3];
};

struct LargeParams {
    cf* X; int64_t x_stride;     // [nvp][L]
    cf* Z; int64_t z_stride;     // [nvp][nvp]
    cf* Q; int64_t g_stride;     // [npairs][P][P] per job
    int* rotflag;                // [njobs][npairs]: did this round's pass rotate anything in the pair?
    int* gcount;                 // [njobs][npairs]: arrival counter of the split Gram (zero between launches)
    cf* gpart;                   // [njobs][npairs][nsplit][P*P] partial Gram matrices (nsplit > 1 only)
    int nsplit;                  // CTAs sharing one pair's Gram (column ranges)
    int gram_tc;                 // the Grams come from bj_gram_tc_kernel: gpart = [njobs][npairs][re | im][P*P] floats
    float* sigma; int* perm; int64_t s_stride;   // [nvp]
    Misc* misc;
    int nv, L, nvp, nb, npairs;
    float tol2, eta2;
    long long* clk;              // phase clocks of CTA (0,0) (MPSB_LARGE_CLOCKS=1, profiling only) or nullptr
};
#define BJ_CLK(i) do { if (p.clk && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) p.clk[i] = clock64(); } while (0)

__device__ __forceinline__ void pair_blocks(int nb, int r, int g, int& I, int& J) {
    if (nb == 2) { I = 0; J = 1; return; }
    const int m = nb - 1;
    if (g == 0) { I = m; J = r; }
    else { I = (r + g) % m; J = (r - g + m) % m; }
}

// Programmatic dependent launch: the round kernels are launched with
// cudaLaunchAttributeProgrammaticStreamSerialization, so a grid is scheduled while its predecessor
// in the stream drains and only this instruction waits for the predecessor's memory to be visible.
// A single matrix runs ~600 dependent rounds of two ~8 us kernels: the launch latency hidden this
// way is a fixed share of each of them.
// launch_dependents right after the wait: the next grid becomes resident (and blocks in its own
// wait) while this one computes, never more than one grid ahead.
__device__ __forceinline__ void griddep_wait() {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

__device__ __forceinline__ int pair_row(int I, int J, int t) { return t < BLK ? I * BLK + t : J * BLK + (t - BLK); }

__global__ void bj_init_kernel(LargeParams p) {
    const int job = blockIdx.y;
    cf* X = p.X + (size_t)job * p.x_stride;
    cf* Z = p.Z + (size_t)job * p.z_stride;
    const size_t nz = (size_t)p.nvp * p.nvp;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < nz; e += (size_t)gridDim.x * blockDim.x) {
        size_t i = e / p.nvp, c = e - i * p.nvp;
        Z[e] = cf_make(i == c ? 1.f : 0.f, 0.f);
    }
    const size_t npad = (size_t)(p.nvp - p.nv) * p.L;
    cf* Xpad = X + (size_t)p.nv * p.L;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < npad; e += (size_t)gridDim.x * blockDim.x)
        Xpad[e] = cf_make(0.f, 0.f);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        Misc& m = p.misc[job];
        m.active = 1; m.rot = 0; m.sweeps = 0; m.gmax_next = 0u; m.gmax = 0.f;
    }
    if (blockIdx.x == 0)
        for (int g = threadIdx.x; g < p.npairs; g += blockDim.x) p.gcount[(size_t)job * p.npairs + g] = 0;
}

__device__ __forceinline__ float rsqrt_approx(float x) {
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__device__ __forceinline__ void evd_rot_params(float a, float b, float gr, float gi, float g2,
                                               float& c, float& sr, float& si) {
    // On the critical path of every rotation set (16-31 dependent sets per launch).  The same two-MUFU form
    // as svd_small.cu: d = a - b, h = sqrt(d^2 + 4|g|^2), w = h + |d|, |s|^2 = 2|g|^2 / (h w); the Gram
    // matrix is scaled to max|G| in [1, 2) and a rotation needs |g|^2 > 1e-30, so the arguments of the
    // approximations stay in range.  The angle only has to be close to the annihilating one; unitarity comes
    // from c being DERIVED from s (and from the Newton-Schulz step on Q).
    const float d = a - b;
    const float u = fmaf(d, d, 4.0f * g2);
    const float h = u * rsqrt_approx(u);
    const float w = h + fabsf(d);
    const float k = copysignf(1.41421356f, d) * rsqrt_approx(h * w);
    sr = gr * k;
    si = gi * k;
    const float hh = fmaf(sr, sr, si * si);
    if (hh < 0.0625f) {
        const float poly = fmaf(hh, fmaf(hh, fmaf(hh, fmaf(hh, 0.02734375f, 0.0390625f), 0.0625f), 0.125f), 0.5f);
        c = fmaf(-hh, poly, 1.0f);
    } else {
        const float y = fmaf(-sr, sr, fmaf(-si, si, 1.0f));
        const float r0 = rsqrt_approx(y);
        const float c0 = y * r0;
        c = fmaf(0.5f * r0, fmaf(-c0, c0, y), c0);
    }
}

// 8-byte cp.async (one complex64; always aligned) with zero fill when !valid
__device__ __forceinline__ void cp_async8(void* smem, const void* gmem, bool valid) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    const int sz = valid ? 8 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(s), "l"(gmem), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

// One launch per round for all (job, pair) problems, a CTA of 256 threads each:
//   1. Gram  G = Xp Xp^H  (G[i][j] = sum_c Xp[i][c] conj(Xp[j][c])): column tiles of GK staged in
//      shared memory by a two-stage cp.async ring; the CTA is four groups of 64 threads, each
//      group covers the whole 32 x 32 output in 4 x 4 register tiles (rows ty + 8 i, tx + 8 j:
//      conflict-free LDS, 8 LDS per 64 FFMA) over its quarter of every tile; the four partial
//      sums are added in a fixed order;
//   2. ONE cyclic pass of two-sided Hermitian Jacobi rotations over the pairs of the two 16-row
//      blocks (see below), G and Q in shared memory: every warp derives the 16 rotations of a set
//      redundantly (no exchange), then applies two of them (rows of G and Q, then columns of G);
//   3. one Newton-Schulz step on Q, Q to global, "rotated" flag for bj_apply_kernel.
// (First versions: three kernels -- Gram with 2 x 2 tiles, a 256-thread EVD with a converged inner
// iteration, later a warp-per-problem EVD kernel -- and a G round trip through L2; then this kernel
// with a single-warp pass, which left 7 of 8 warps at a barrier: profiles/r1_svd_large_ncu_full.txt.)
// rotation `idx` (0..15) of set t of a pass over the two 16-row blocks of a pair: in the first
// round of a sweep the 15 intra-block sets come first (circle method inside each block, 8 pairs
// per block), then the 16 cross sets (i, 16 + (i + s) mod 16)
__device__ __forceinline__ void set_pair(int t, int first_round, int idx, int& pp, int& qq) {
    if (first_round && t < BLK - 1) {
        const int blk = idx >> 3, i = idx & 7, m = BLK - 1;
        int a_, b_;
        if (i == 0) { a_ = m; b_ = t; }
        else { a_ = t + i; a_ = a_ >= m ? a_ - m : a_; b_ = t - i; b_ = b_ < 0 ? b_ + m : b_; }      // (t +- i) mod 15
        pp = blk * BLK + min(a_, b_); qq = blk * BLK + max(a_, b_);
    } else {
        const int sft = first_round ? t - (BLK - 1) : t;
        pp = idx; qq = BLK + ((idx + sft) & (BLK - 1));
    }
}

constexpr int GK = 64;                       // columns per Gram tile
constexpr int XS_LD = GK + 1;
constexpr int GRAM_MAX_STAGES = 4;           // a small batch is bound by the latency of its tile loads: deeper ring
constexpr int gram_smem_bytes(int nst) { return (nst * P * XS_LD + 2 * P * (P + 1)) * (int)sizeof(cf); }

// (ty | tx << 4) of the 36 thread tiles with ty <= tx (see bj_gram_evd_kernel)
__constant__ unsigned char c_tri[36] = {0, 16, 32, 48, 64, 80, 96, 112, 17, 33, 49, 65, 81, 97, 113, 34, 50, 66, 82, 98, 114, 51, 67, 83, 99, 115, 68, 84, 100, 116, 85, 101, 117, 102, 118, 119};

template <int NTHR, bool SYM>                // NTHR: 256 (512: timing experiments); SYM: see the Gram loop
__global__ void __launch_bounds__(NTHR, NTHR == 512 ? 1 : 3) bj_gram_evd_kernel(LargeParams p, int round, int first_round, int nst) {
    constexpr int NGS = SYM ? NTHR / 36 : NTHR / 64, NWARP = NTHR / 32;
    const int job = blockIdx.y, g = blockIdx.x;
    griddep_wait();
    if (!p.misc[job].active) return;
    extern __shared__ float4 gram_smem[];
    cf* Xbuf = reinterpret_cast<cf*>(gram_smem);
    cf (*Gs)[P + 1] = reinterpret_cast<cf (*)[P + 1]>(Xbuf + nst * P * XS_LD);
    cf (*Qs)[P + 1] = Gs + P;
    int I, J;
    pair_blocks(p.nb, round, g, I, J);
    const cf* X = p.X + (size_t)job * p.x_stride;
    // Thread tiles: 4 x 4 entries, rows ty + 8 i, columns tx + 8 j (conflict-free LDS); the threads
    // of a group cover G once and the NGS groups share the columns of every staged tile.
    // SYM: G is Hermitian, so only the 36 thread tiles with ty <= tx are computed, by NGS = 7
    // groups of 36 threads, and the rest is mirrored afterwards: ~9 columns of 16 complex FMAs
    // per thread and tile instead of 16 (this loop is FFMA-issue bound: 79 % of the issue slots in
    // profiles/r1_svd_large_ncu_full.txt).  Measured: 50 x 512^2 194 -> 179 ms, 4 x 2048^2
    // 1.02 -> 0.94 s, configs[2] 722 -> 766 applications/s.  CTAs that see a single tile (split
    // Gram of one 256 x 256 matrix) keep the full 8 x 8 grid: the mirrored variant was 4 % slower
    // there (8.14 -> 8.47 ms per solve).
    constexpr int TPG = SYM ? 36 : 64;
    const bool gact = threadIdx.x < NGS * TPG;
    const int grp = threadIdx.x / TPG;
    const int tyx = SYM ? c_tri[threadIdx.x % TPG] : ((threadIdx.x >> 3) & 7) | ((threadIdx.x & 7) << 4);
    const int ty = tyx & 15, tx = tyx >> 4;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    auto issue = [&](int c0, cf* buf) {
        for (int e = threadIdx.x; e < P * GK; e += NTHR) {
            const int r = e / GK, c = e % GK;
            const bool valid = c0 + c < p.L;
            cp_async8(&buf[r * XS_LD + c], valid ? X + (size_t)pair_row(I, J, r) * p.L + c0 + c : X, valid);
        }
        cp_async_commit();
    };
    if (p.gram_tc) {
        // G was formed on the tensor cores (bj_gram_tc_kernel): the upper triangle of the two planes, mirrored
        const float* gp = reinterpret_cast<const float*>(p.gpart) + ((size_t)job * p.npairs + g) * (2 * P * P);
        for (int e = threadIdx.x; e < P * P; e += NTHR) {
            const int r = e / P, c = e % P;
            const int eu = r <= c ? e : c * P + r;
            const float re = __ldcg(gp + eu), im = r == c ? 0.f : __ldcg(gp + P * P + eu);
            Gs[r][c] = cf_make(re, r <= c ? im : -im);
        }
        __syncthreads();
    } else {
    cf acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = cf_make(0.f, 0.f);
    // this CTA's column tiles: all of them, or one of nsplit contiguous ranges (blockIdx.z)
    const int ntile_all = (p.L + GK - 1) / GK;
    const int tile_lo = (int)(((long long)ntile_all * blockIdx.z) / p.nsplit);
    const int ntile = (int)(((long long)ntile_all * (blockIdx.z + 1)) / p.nsplit) - tile_lo;
    BJ_CLK(0);
    for (int t = 0; t < nst - 1 && t < ntile; ++t) issue((tile_lo + t) * GK, Xbuf + t * (P * XS_LD));
    for (int t = 0; t < ntile; ++t) {
        const cf* Xs = Xbuf + (t % nst) * (P * XS_LD);
        if (t + nst - 1 < ntile) issue((tile_lo + t + nst - 1) * GK, Xbuf + ((t + nst - 1) % nst) * (P * XS_LD));
        const int pending = min(ntile - 1 - t, nst - 1);     // groups that may still be in flight
        if (pending >= 3) cp_async_wait<3>();
        else if (pending == 2) cp_async_wait<2>();
        else if (pending == 1) cp_async_wait<1>();
        else cp_async_wait<0>();
        __syncthreads();
#pragma unroll 2
        for (int n = grp; n < (gact ? GK : 0); n += NGS) {
            // SYM: column n of the tile in 8 x 8 transposed order -- the two groups that share a
            // warp read columns 8 apart, i.e. the other half of the shared-memory banks (row stride 65)
            const int c = SYM ? ((n & 7) << 3) | (n >> 3) : n;
            cf a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = Xs[(ty + 8 * i) * XS_LD + c];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = Xs[(tx + 8 * j) * XS_LD + c];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = cf_fma_conja(b[j], a[i], acc[i][j]);     // a * conj(b)
        }
        __syncthreads();
    }
    BJ_CLK(1);
    if (SYM && NGS * 36 * 16 <= nst * P * XS_LD) {       // block-uniform
        // the partial tiles of the NGS groups go through the (now free) tile ring, laid out
        // [group][tile entry][thread tile] so that a warp stores consecutive elements; every
        // thread then adds the NGS partial sums of its G entries in a fixed order (one barrier
        // instead of NGS) and the mirrored entries G[r][c] = conj(G[c][r]) are filled on the way
        cf* part = Xbuf;
        const int u36 = threadIdx.x % 36;
        if (gact) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) part[(grp * 16 + i * 4 + j) * 36 + u36] = acc[i][j];
        }
        __syncthreads();
        for (int e = threadIdx.x; e < P * P; e += NTHR) {
            const int r = e / P, c = e % P;
            const bool up = (r & 7) <= (c & 7);
            const int rr = up ? r : c, cc = up ? c : r;
            const int ty_ = rr & 7, tx_ = cc & 7;
            const int uu = ty_ * 8 - ty_ * (ty_ - 1) / 2 + (tx_ - ty_);
            const int idx = (rr >> 3) * 4 + (cc >> 3);
            cf sum = part[idx * 36 + uu];
            for (int q = 1; q < NGS; ++q) sum = cf_add(sum, part[(q * 16 + idx) * 36 + uu]);
            Gs[r][c] = up ? sum : cf_conj(sum);
        }
        __syncthreads();
    } else {
        for (int q = 0; q < NGS; ++q) {
            if (grp == q) {                   // (threads beyond the last group have grp == NGS)
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        cf& d = Gs[ty + 8 * i][tx + 8 * j];
                        d = q == 0 ? acc[i][j] : cf_add(d, acc[i][j]);
                    }
            }
            __syncthreads();
        }
        if (SYM) {
            // the other triangle of the thread-tile grid: G[r][c] = conj(G[c][r]) where (r mod 8) > (c mod 8)
            for (int e = threadIdx.x; e < P * P; e += NTHR) {
                const int r = e / P, c = e % P;
                if ((r & 7) > (c & 7)) Gs[r][c] = cf_conj(Gs[c][r]);
            }
            __syncthreads();
        }
    }
    if (p.nsplit > 1) {
        // A launch with fewer pairs than SMs (one or two matrices) spreads each pair's Gram over
        // nsplit CTAs: partial sums go to global memory, the CTA that arrives last adds them in a
        // fixed order and carries on alone (one pair's 32 x 32 x L Gram is otherwise bound by the
        // FFMA issue rate of a single SM: 17 k of the 45 k cycles of this kernel at L = 256).
        const size_t pr = (size_t)job * p.npairs + g;
        cf* mine = p.gpart + (pr * p.nsplit + blockIdx.z) * (P * P);
        for (int e = threadIdx.x; e < P * P; e += NTHR) mine[e] = Gs[e / P][e % P];
        __threadfence();
        __syncthreads();
        __shared__ int s_last;
        if (threadIdx.x == 0) {
            const int ticket = atomicAdd(&p.gcount[pr], 1);
            s_last = ticket == p.nsplit - 1;
            if (s_last) p.gcount[pr] = 0;                    // ready for the next launch
        }
        __syncthreads();
        if (!s_last) return;
        __threadfence();
        const cf* all = p.gpart + pr * p.nsplit * (P * P);
        for (int e = threadIdx.x; e < P * P; e += NTHR) {
            cf sum = __ldcg(&all[e]);
            for (int z = 1; z < p.nsplit; ++z) sum = cf_add(sum, __ldcg(&all[(size_t)z * (P * P) + e]));
            Gs[e / P][e % P] = sum;
        }
        __syncthreads();
    }
    }   // !p.gram_tc
    // scale by a power of two so that max|G| is in [1, 2)   (every warp derives the same factor)
    float mx = 0.f;
    for (int i = 0; i < P; ++i) { cf v = Gs[i][lane]; mx = fmaxf(mx, fmaxf(fabsf(v.x), fabsf(v.y))); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    int ex = 0;
    if (mx > 0.f && isfinite(mx)) (void)frexpf(mx, &ex);
    const float sc = mx > 0.f ? ldexpf(1.0f, 1 - ex) : 1.0f;
    __syncthreads();
#pragma unroll
    for (int ii = 0; ii < P / NWARP; ++ii) {
        const int i = warp + NWARP * ii;
        cf v = Gs[i][lane];
        Gs[i][lane] = cf_make(v.x * sc, v.y * sc);
        Qs[i][lane] = cf_make(i == lane ? 1.f : 0.f, 0.f);
    }
    __syncthreads();
    // Rotation criterion: |G_pq| > tol sqrt(G_pp G_qq)  AND  |G_pq| > eta sigma_max max(s_p, s_q).
    // The second (absolute) part matters for graded spectra: every GEMM that mixes a large row
    // into a small one leaves ~eps*sigma_max of noise in it, so |G_pq| of a small pair carries
    // noise ~eps*sigma_max*max(s_p, s_q) and the relative criterion alone is never met in fp32.
    // A row far below sigma_max keeps a component of up to eta*sigma_max along each larger row,
    // i.e. an absolute error of ~eta*sqrt(nv)*sigma_max in the smallest singular values: measured
    // 6e-6 / 1.9e-5 sigma_max at nv = 512 / 2048 with eta = 1e-6, 6e-7 / 4e-6 with 2.5e-7 at the
    // same sweep count (parity bound: 1e-5 sigma_max); eta = 3e-7 is ~5 eps.  Without
    // preconditioning the cyclic block method converges only linearly (~2x per sweep) until then:
    // ~10 sweeps for flat spectra, ~20 for graded ones (tests/_jacobi_model.py notes).
    const float gmax_s = p.misc[job].gmax * sc;
    const float eta2g = p.eta2 * gmax_s;
    BJ_CLK(2);
    // ONE pass over the pairs of the two 16-row blocks, each rotation computed from the current
    // (two-sidedly updated) Gram entries: the 16 cross sets (i, 16 + (i+s) mod 16), preceded in the
    // first round of an outer sweep by the 15 intra-block sets (circle method inside each block).
    // This is the scalar cyclic Jacobi sweep of svd_small.cu carried out on the Gram matrix: the
    // outer iteration needs the same ~9 sweeps, but a pair step costs 16 (31) rotation sets instead
    // of the 5 x 31 of a fully converged inner eigen-decomposition.
    // A set is 16 disjoint rotations J_i on rows/columns (p_i, q_i): G <- J G J^H splits into 16 x 16
    // independent 2 x 2 blocks  B_ij = G[{p_i,q_i}][{p_j,q_j}] <- J_i B_ij J_j^H, one per thread
    // (thread = (i, j)); Q <- J Q likewise (thread (i, j): rows p_i, q_i, columns 2j, 2j+1).  The
    // sixteen lanes of warp 0 derive the rotations; the (p, q) of every set come from a table built
    // once per launch.  Two barriers per set.  (First version: rows then columns, 2 rotations per warp, 3 barriers per set and
    // ~1 400 cycles per set on the critical path -- half of this kernel for a single matrix.)
    float4* prm = reinterpret_cast<float4*>(Xbuf);           // [16] (c, s.re, s.im, rotate?), tile ring is free now
    unsigned short* pairtab = reinterpret_cast<unsigned short*>(prm + 16);      // [nsets][16]  p | q << 8
    int total_rot = 0;
    const int nsets = (first_round ? (BLK - 1) : 0) + BLK;
    for (int e = threadIdx.x; e < nsets * 16; e += NTHR) {
        int pp, qq;
        set_pair(e >> 4, first_round, e & 15, pp, qq);
        pairtab[e] = (unsigned short)(pp | (qq << 8));
    }
    __syncthreads();
    const int ri = (threadIdx.x >> 4) & 15, rj = threadIdx.x & 15;
    for (int t = 0; t < nsets; ++t) {
        const int pqi = pairtab[t * 16 + ri], pqj = pairtab[t * 16 + rj];
        const int pi_ = pqi & 0xff, qi_ = pqi >> 8, pj_ = pqj & 0xff, qj_ = pqj >> 8;
        int dorot = 0;
        if (threadIdx.x < 16) {                  // rotation rj of the set (lanes 0-15 of warp 0)
            const float a = Gs[pj_][pj_].x, b = Gs[qj_][qj_].x;
            const cf gg = Gs[pj_][qj_];
            const float g2 = cf_abs2(gg);
            float c = 1.f, sr = 0.f, si = 0.f;
            dorot = (g2 > p.tol2 * a * b && g2 > eta2g * fmaxf(a, b) && g2 > 1e-30f) ? 1 : 0;
            if (dorot) evd_rot_params(a, b, gg.x, gg.y, g2, c, sr, si);
            prm[rj] = make_float4(c, sr, si, dorot ? 1.f : 0.f);
        }
        const int nrot = __syncthreads_count(dorot);
        if (nrot == 0) continue;                 // block-uniform: nothing to rotate in this set
        total_rot += nrot;
        const float4 Ji = prm[ri], Jj = prm[rj];
        if (threadIdx.x < 256 && (Ji.w != 0.f || Jj.w != 0.f)) {
            cf g00 = Gs[pi_][pj_], g01 = Gs[pi_][qj_], g10 = Gs[qi_][pj_], g11 = Gs[qi_][qj_];
            // left: [x; y] <- J_i [x; y]:  x' = c x + s y,  y' = c y - conj(s) x
            {
                const float c = Ji.x, sr = Ji.y, si = Ji.z;
                cf n00, n01, n10, n11;
                n00.x = fmaf(c, g00.x, fmaf(sr, g10.x, -(si * g10.y)));
                n00.y = fmaf(c, g00.y, fmaf(sr, g10.y, si * g10.x));
                n10.x = fmaf(c, g10.x, -fmaf(sr, g00.x, si * g00.y));
                n10.y = fmaf(c, g10.y, fmaf(si, g00.x, -(sr * g00.y)));
                n01.x = fmaf(c, g01.x, fmaf(sr, g11.x, -(si * g11.y)));
                n01.y = fmaf(c, g01.y, fmaf(sr, g11.y, si * g11.x));
                n11.x = fmaf(c, g11.x, -fmaf(sr, g01.x, si * g01.y));
                n11.y = fmaf(c, g11.y, fmaf(si, g01.x, -(sr * g01.y)));
                g00 = n00; g01 = n01; g10 = n10; g11 = n11;
            }
            // right: [x, y] <- [x, y] J_j^H:  x' = c x + conj(s) y,  y' = c y - s x
            {
                const float c = Jj.x, sr = Jj.y, si = Jj.z;
                cf n00, n01, n10, n11;
                n00.x = fmaf(c, g00.x, fmaf(sr, g01.x, si * g01.y));
                n00.y = fmaf(c, g00.y, fmaf(sr, g01.y, -(si * g01.x)));
                n01.x = fmaf(c, g01.x, -fmaf(sr, g00.x, -(si * g00.y)));
                n01.y = fmaf(c, g01.y, -fmaf(sr, g00.y, si * g00.x));
                n10.x = fmaf(c, g10.x, fmaf(sr, g11.x, si * g11.y));
                n10.y = fmaf(c, g10.y, fmaf(sr, g11.y, -(si * g11.x)));
                n11.x = fmaf(c, g11.x, -fmaf(sr, g10.x, -(si * g10.y)));
                n11.y = fmaf(c, g11.y, -fmaf(sr, g10.y, si * g10.x));
                g00 = n00; g01 = n01; g10 = n10; g11 = n11;
            }
            Gs[pi_][pj_] = g00; Gs[pi_][qj_] = g01; Gs[qi_][pj_] = g10; Gs[qi_][qj_] = g11;
        }
        if (threadIdx.x < 256 && Ji.w != 0.f) {
            const float c = Ji.x, sr = Ji.y, si = Ji.z;
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int col = 2 * rj + u;
                const cf x = Qs[pi_][col], y = Qs[qi_][col];
                cf nx, ny;
                nx.x = fmaf(c, x.x, fmaf(sr, y.x, -(si * y.y)));
                nx.y = fmaf(c, x.y, fmaf(sr, y.y, si * y.x));
                ny.x = fmaf(c, y.x, -fmaf(sr, x.x, si * x.y));
                ny.y = fmaf(c, y.y, fmaf(si, x.x, -(sr * x.y)));
                Qs[pi_][col] = nx; Qs[qi_][col] = ny;
            }
        }
        __syncthreads();
    }
    BJ_CLK(3);
    if (warp == 0) {
        float lam = Gs[lane][lane].x;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) lam = fmaxf(lam, __shfl_xor_sync(0xffffffffu, lam, o));
        if (lane == 0) {
            atomicMax(&p.misc[job].gmax_next, __float_as_uint(fmaxf(lam, 0.f) / sc));
            if (total_rot > 0) atomicAdd(&p.misc[job].rot, 1);
            p.rotflag[(size_t)job * p.npairs + g] = total_rot > 0 ? 1 : 0;
        }
    }
    if (total_rot == 0) return;              // Q = I: bj_apply_kernel skips this pair (same count in every warp)
    __syncthreads();                         // lam has been read before R overwrites G
    // Q is written UNSORTED: small-angle rotations started from the identity keep Q close to
    // the identity, which the cyclic block method needs to converge (sorting the rows by
    // eigenvalue is a permutation far from the identity and makes the outer iteration cycle).
    // One Newton-Schulz step Q <- (3 Q - (Q Q^H) Q) / 2: the product of ~10^3 fp32 rotations is
    // unitary only to ~2e-6, and that error would random-walk into every row norm (= singular
    // value) over the ~400 block rotations of a solve; after the step Q is unitary to ~1e-7.
    // R = Q Q^H into Gs, then Qo = 1.5 Q - 0.5 R Q   (thread = column `lane` of rows warp + 8 ii).
    {
        constexpr int RW = P / NWARP;
        cf r[RW];
#pragma unroll
        for (int ii = 0; ii < RW; ++ii) r[ii] = cf_make(0.f, 0.f);
#pragma unroll 8
        for (int k = 0; k < P; ++k) {
            const cf qj = Qs[lane][k];
#pragma unroll
            for (int ii = 0; ii < RW; ++ii) r[ii] = cf_fma_conja(qj, Qs[warp + NWARP * ii][k], r[ii]);   // Q_ik conj(Q_jk)
        }
#pragma unroll
        for (int ii = 0; ii < RW; ++ii) Gs[warp + NWARP * ii][lane] = r[ii];
    }
    __syncthreads();
    {
        cf* Qo = p.Q + (size_t)job * p.g_stride + (size_t)g * P * P;
        constexpr int RW = P / NWARP;
        cf t[RW];
#pragma unroll
        for (int ii = 0; ii < RW; ++ii) t[ii] = cf_make(0.f, 0.f);
#pragma unroll 8
        for (int k = 0; k < P; ++k) {
            const cf qk = Qs[k][lane];
#pragma unroll
            for (int ii = 0; ii < RW; ++ii) t[ii] = cf_fma(Gs[warp + NWARP * ii][k], qk, t[ii]);
        }
#pragma unroll
        for (int ii = 0; ii < RW; ++ii) {
            const int i = warp + NWARP * ii;
            const cf q = Qs[i][lane];
            Qo[i * P + lane] = cf_make(1.5f * q.x - 0.5f * t[ii].x, 1.5f * q.y - 0.5f * t[ii].y);
        }
    }
    BJ_CLK(4);
}

// rows of the pair, columns of [X | Z]:  T <- Q T   (in place).  A CTA walks NT_APPLY column
// tiles of CT columns with a two-stage cp.async ring (the loads of tile t+1 overlap the FFMAs of
// tile t; the first version loaded, synchronised, computed and stored one tile per CTA and spent
// its time on the long scoreboard: FMA pipe 39 % -- profiles/r1_svd_large_ncu_full.txt).  A thread
// owns two columns (c, c + CT/2) and 8 of the 32 output rows: per k two LDS.64 of T and four
// broadcast LDS.128 of Q^T for 64 FFMA.
constexpr int NT_APPLY = 4;                                  // at most; fewer for small batches (large_begin)
constexpr int TS_LD = CT + 2;                                 // row stride of a staged tile (elements)
constexpr int APPLY_SMEM = (P * (P + 2) + 2 * P * TS_LD) * (int)sizeof(cf);

__global__ void __launch_bounds__(LT) bj_apply_kernel(LargeParams p, int round, int ntx, int ntot, int nt_cta) {
    const int job = blockIdx.z, g = blockIdx.y;
    griddep_wait();
    if (!p.misc[job].active) return;
    if (!p.rotflag[(size_t)job * p.npairs + g]) return;      // no rotation in this pair: Q = I
    extern __shared__ float4 apply_smem[];
    cf (*Qt)[P + 2] = reinterpret_cast<cf (*)[P + 2]>(apply_smem);          // Qt[k][i] = Q[i][k]
    cf* Tbuf = reinterpret_cast<cf*>(apply_smem) + P * (P + 2);
    int I, J;
    pair_blocks(p.nb, round, g, I, J);
    const int tile0 = blockIdx.x * nt_cta;
    const int nt = min(nt_cta, ntot - tile0);
    cf* const Xb = p.X + (size_t)job * p.x_stride;
    cf* const Zb = p.Z + (size_t)job * p.z_stride;
    auto issue = [&](int tile, cf* buf) {
        cf* base; int ld, c0, ncol;
        if (tile < ntx) { base = Xb; ld = p.L; c0 = tile * CT; ncol = p.L; }
        else { base = Zb; ld = p.nvp; c0 = (tile - ntx) * CT; ncol = p.nvp; }
        for (int e = threadIdx.x; e < P * CT; e += LT) {
            const int r = e / CT, c = e - r * CT;
            const bool valid = c0 + c < ncol;
            cp_async8(&buf[r * TS_LD + c], valid ? base + (size_t)pair_row(I, J, r) * ld + c0 + c : base, valid);
        }
        cp_async_commit();
    };
    issue(tile0, Tbuf);
    const cf* Q = p.Q + (size_t)job * p.g_stride + (size_t)g * P * P;
    for (int e = threadIdx.x; e < P * P; e += LT) Qt[e % P][e / P] = Q[e];
    constexpr int HC = CT / 2;                               // 64 column pairs x 4 row groups
    const int c = threadIdx.x % HC, rg = threadIdx.x / HC;
    constexpr int RPT = P / (LT / HC);                       // 8 rows per thread
    for (int t = 0; t < nt; ++t) {
        const cf* Ts = Tbuf + (t & 1) * (P * TS_LD);
        if (t + 1 < nt) { issue(tile0 + t + 1, Tbuf + ((t + 1) & 1) * (P * TS_LD)); cp_async_wait<1>(); }
        else cp_async_wait<0>();
        __syncthreads();
        cf acc0[RPT], acc1[RPT];
#pragma unroll
        for (int i = 0; i < RPT; ++i) { acc0[i] = cf_make(0.f, 0.f); acc1[i] = cf_make(0.f, 0.f); }
#pragma unroll 4
        for (int k = 0; k < P; ++k) {
            const cf t0 = Ts[k * TS_LD + c], t1 = Ts[k * TS_LD + c + HC];
            const float4* qrow = reinterpret_cast<const float4*>(&Qt[k][rg * RPT]);
#pragma unroll
            for (int i2 = 0; i2 < RPT / 2; ++i2) {
                const float4 q2 = qrow[i2];
                const cf qa = cf_make(q2.x, q2.y), qb = cf_make(q2.z, q2.w);
                acc0[2 * i2] = cf_fma(qa, t0, acc0[2 * i2]);
                acc1[2 * i2] = cf_fma(qa, t1, acc1[2 * i2]);
                acc0[2 * i2 + 1] = cf_fma(qb, t0, acc0[2 * i2 + 1]);
                acc1[2 * i2 + 1] = cf_fma(qb, t1, acc1[2 * i2 + 1]);
            }
        }
        const int tile = tile0 + t;
        cf* base; int ld, c0, ncol;
        if (tile < ntx) { base = Xb; ld = p.L; c0 = tile * CT; ncol = p.L; }
        else { base = Zb; ld = p.nvp; c0 = (tile - ntx) * CT; ncol = p.nvp; }
#pragma unroll
        for (int i = 0; i < RPT; ++i) {
            cf* row = base + (size_t)pair_row(I, J, rg * RPT + i) * ld + c0;
            if (c0 + c < ncol) row[c] = acc0[i];
            if (c0 + c + HC < ncol) row[c + HC] = acc1[i];
        }
        __syncthreads();                                     // the other stage is refilled next iteration
    }
}

// ---- the same block rotation on the tensor cores ----------------------------------------------
// The apply of launches with at least two tiles per SM (large_begin; MPSB_LARGE_TC_APPLY=0/1 forces
// the choice for A/B timing); bj_apply_kernel above serves the one- and two-matrix launches.
// Measured on B200 (profiles/r1_tc_apply_runs.txt): 50 x 512^2 178 -> 145 ms per solve batch (the
// apply itself 0.198 -> ~0.093 ms per round), configs[2] 766 -> 892 applications/s; all GPU parity
// tests green with it (sigma within 1e-6 sigma_max of LAPACK at 512^2, within the 1e-5 bound at 2048^2).
//
// out[i][c] = sum_k Q[i][k] T[k][c] for the 32 rows of a pair and a tile of 128 columns, as ONE real
// 3xTF32 product per tile with the columns of the tile as MMA rows:
//   D[m][2i+u] = sum_kk A[m][kk] B[2i+u][kk],   m = column in the tile, kk = 2k + t  (K' = 64)
//   A[m][2k+t]  = t ? Im T[k][m] : Re T[k][m]
//   B[2i][2k] = Qr, B[2i][2k+1] = -Qi, B[2i+1][2k] = Qi, B[2i+1][2k+1] = Qr        (Q = Q[i][k])
// so that row m of the accumulator is the interleaved complex column m of the output: TMEM lane m
// stores out[i][c0+m] = (D[m][2i], D[m][2i+1]) for i = 0..31, a warp writing 256 contiguous bytes
// per row.  M = 128, N = 64, K' = 64: 2 k-blocks x 4 k-steps x 3 operand pairings = 24
// tcgen05.mma.kind::tf32 per tile (hi*lo + lo*hi + hi*hi, fp32 accumulation in TMEM).
// The rows of X are contiguous along m, the MMA wants K-major operands, and 3xTF32 needs the
// hi/lo split: the loader warps read the tile from global memory (coalesced), split it in
// registers and write A_hi / A_lo straight into the canonical K-major SWIZZLE_128B layout that TMA
// would produce (row m at m * 128 bytes, 16-byte chunk j stored at j ^ (m & 7)), then
// fence.proxy.async + mbarrier.  A separate split pass through HBM would triple the traffic of a
// round; this way a tile is read once and written once (64 KB per tile and SM: HBM bound).
// Roles (416 threads, persistent, one CTA per SM, contiguous range of (job, pair, tile) items per
// CTA so that B = the embedding of Q is rebuilt only when the pair changes):
//   warps 0-3  epilogue: tcgen05.ld of the 64 accumulator columns of lane m, global stores;
//   warp  4    MMA issuer (one lane), two 64-column accumulators in TMEM so that the drain of a
//              tile overlaps the MMAs of the next;
//   warps 5-12 loaders / splitters, 2-stage ring of 96 KB stages (A_hi, A_lo, B_hi, B_lo).
constexpr int TA_M = 128;                               // columns per tile = MMA rows (TMEM lanes)
constexpr int TA_N = 2 * P;                             // 64 accumulator columns: (i, re | im)
constexpr int TA_KB_A = TA_M * 128;                     // bytes of one k-block (32 fp32) of A
constexpr int TA_KB_B = TA_N * 128;                     // ... of B
constexpr int TA_A_BYTES = 2 * TA_KB_A;                 // K' = 64 fp32 = 2 k-blocks
constexpr int TA_B_BYTES = 2 * TA_KB_B;
constexpr int TA_STAGE_BYTES = 2 * TA_A_BYTES + 2 * TA_B_BYTES;     // 96 KB
constexpr int TA_STAGES = 2;
constexpr int TA_SMEM = TA_STAGES * TA_STAGE_BYTES + 1024;           // + slack to align the ring to 1024
constexpr int TC_APPLY_MAX_ROUNDS = 3200;               // see large_enqueue_sweep
constexpr int TA_EPI_WARPS = 4, TA_LOAD_WARPS = 8;
constexpr int TA_LOAD_THREADS = TA_LOAD_WARPS * 32;
constexpr int TA_THREADS = (TA_EPI_WARPS + 1 + TA_LOAD_WARPS) * 32;
static_assert(CT == TA_M, "the tensor-core apply uses the tile width of bj_apply_kernel");

struct TaItem { int job, g, tile; bool skip; };

__device__ __forceinline__ TaItem ta_item(const LargeParams& p, long long it, int ntot) {
    TaItem w;
    const long long pr = it / ntot;
    w.tile = (int)(it - pr * ntot);
    w.job = (int)(pr / p.npairs);
    w.g = (int)(pr - (long long)w.job * p.npairs);
    w.skip = !p.misc[w.job].active || !p.rotflag[pr];
    return w;
}

// the same with the two flag loads (global memory, ~1 us of latency on a role's serial path) done once per
// (job, pair) instead of once per tile
struct TaCursor { long long pr = -1; bool skip = false; };
__device__ __forceinline__ TaItem ta_item_cached(const LargeParams& p, long long it, int ntot, TaCursor& c) {
    TaItem w;
    const long long pr = it / ntot;
    w.tile = (int)(it - pr * ntot);
    w.job = (int)(pr / p.npairs);
    w.g = (int)(pr - (long long)w.job * p.npairs);
    if (pr != c.pr) { c.pr = pr; c.skip = !p.misc[w.job].active || !p.rotflag[pr]; }
    w.skip = c.skip;
    return w;
}

// Operand stores of the loader warps.  The compiler emits generic ST.E.64 for them (the ring pointer
// is derived from an aligned cast); building with -DTA_STS=1 (MPSB_NVCC_EXTRA, csrc/build.py) uses
// explicit st.shared.v2.f32 instead -- a variant to A/B on hardware, not yet measured.
#ifndef TA_STS
#define TA_STS 0
#endif
// -DTA_DELTA=1 (variant, not yet measured): the tensor cores compute (Q - I) T and the epilogue adds
// T in fp32.  The accumulator truncates toward zero once per accumulating MMA; with Q itself the
// dominant term q_ii t_i sits in the accumulator while ~12 more MMAs are added, a systematic shrink
// of ~3e-7 per apply (x several thousand applies per solve, removed at the end by the
// Newton-Schulz step on Z).  With Q - I only the O(angle) correction is accumulated, so the
// truncation error scales with the rotation angle and vanishes as the sweeps converge; the price
// is one read of the tile by the epilogue threads (L2 hits: the loaders just fetched it).
#ifndef TA_DELTA
#define TA_DELTA 0
#endif
__device__ __forceinline__ void ta_store(uint8_t* base, int off, float a, float b) {
#if TA_STS
    asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(tcx::smem_u32(base) + (uint32_t)off), "f"(a), "f"(b) : "memory");
#else
    *reinterpret_cast<float2*>(base + off) = make_float2(a, b);
#endif
}

// byte offset of the (re, im) pair of pair-row k in row r of a K-major SWIZZLE_128B operand
__device__ __forceinline__ int ta_offset(int r, int k, int kb_bytes) {
    return (k >> 4) * kb_bytes + r * 128 + ((((k & 15) >> 1) ^ (r & 7)) << 4) + ((k & 1) << 3);
}

__global__ void __launch_bounds__(TA_THREADS, 1) bj_apply_tc_kernel(LargeParams p, int round, int ntx, int ntot, int njobs) {
    using namespace tcx;
    extern __shared__ __align__(1024) uint8_t ta_smem[];
    __shared__ __align__(8) uint64_t full_bar[TA_STAGES], empty_bar[TA_STAGES], acc_full[2], acc_empty[2];
    __shared__ uint32_t tmem_slot;
    uint8_t* ring = (uint8_t*)(((uintptr_t)ta_smem + 1023) & ~(uintptr_t)1023);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    griddep_wait();
    if (warp == TA_EPI_WARPS && lane == 0) {
        for (int s = 0; s < TA_STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&acc_full[a], 1); mbar_init(&acc_empty[a], TA_EPI_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(2u * TA_N) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;

    const long long total = (long long)njobs * p.npairs * ntot;
    const long long per = (total + gridDim.x - 1) / gridDim.x;
    const long long it0 = per * blockIdx.x, it1 = min(total, it0 + per);

    if (warp > TA_EPI_WARPS) {
        // ===== loaders: global -> registers (hi/lo split) -> swizzled K-major operands =====
        const int lt = threadIdx.x - (TA_EPI_WARPS + 1) * 32, lw = lt >> 5;
        long long held0 = -1, held1 = -1;                  // pair whose B each stage holds
        int n = 0;
        for (long long it = it0; it < it1; ++it) {
            const TaItem w = ta_item(p, it, ntot);
            if (w.skip) continue;
            const int s = n & 1;
            uint8_t* st = ring + (size_t)s * TA_STAGE_BYTES;
            int I, J;
            pair_blocks(p.nb, round, w.g, I, J);
            const cf* base; int ld, c0, ncol;
            if (w.tile < ntx) { base = p.X + (size_t)w.job * p.x_stride; ld = p.L; c0 = w.tile * TA_M; ncol = p.L; }
            else { base = p.Z + (size_t)w.job * p.z_stride; ld = p.nvp; c0 = (w.tile - ntx) * TA_M; ncol = p.nvp; }
            // all 16 loads of this thread first (the shared-memory stores below may alias them as far as
            // the compiler knows: interleaved, every load waited for the previous store -- 8 us per tile).
            // Measured and not adopted (round 2): a second register tile so that the loads of the NEXT tile
            // are in flight while this one is split and stored -- 50 x 512^2 145 -> 157 ms, configs[2]
            // 890 -> 820 applications/s (the loader warps spill at the 128-register cap of a 416-thread CTA).
            constexpr int RPW = P / TA_LOAD_WARPS, CPL = TA_M / 32;
            cf v[RPW][CPL];
#pragma unroll
            for (int rr = 0; rr < RPW; ++rr) {
                const cf* row = base + (size_t)pair_row(I, J, lw * RPW + rr) * ld + c0;
#pragma unroll
                for (int j = 0; j < CPL; ++j) v[rr][j] = c0 + lane + 32 * j < ncol ? __ldcg(row + lane + 32 * j) : cf_make(0.f, 0.f);
            }
            const long long pr = (long long)w.job * p.npairs + w.g;
            const bool newq = (s ? held1 : held0) != pr;
            cf qv[P * P / TA_LOAD_THREADS];
            if (newq) {
                const cf* Q = p.Q + (size_t)w.job * p.g_stride + (size_t)w.g * P * P;
#pragma unroll
                for (int j = 0; j < P * P / TA_LOAD_THREADS; ++j) qv[j] = __ldcg(Q + lt + TA_LOAD_THREADS * j);
            }
            mbar_wait(&empty_bar[s], ((n >> 1) & 1) ^ 1);    // (the loads above are in flight while the stage drains)
#pragma unroll
            for (int rr = 0; rr < RPW; ++rr) {
                const int k = lw * RPW + rr;
#pragma unroll
                for (int j = 0; j < CPL; ++j) {
                    const int m = lane + 32 * j;
                    float rh, rl, ih, il;
                    split_tf32(v[rr][j].x, rh, rl);
                    split_tf32(v[rr][j].y, ih, il);
                    const int off = ta_offset(m, k, TA_KB_A);
                    ta_store(st, off, rh, ih);
                    ta_store(st + TA_A_BYTES, off, rl, il);
                }
            }
            if (newq) {
                if (s) held1 = pr; else held0 = pr;
                uint8_t* bh = st + 2 * TA_A_BYTES;
                uint8_t* bl = bh + TA_B_BYTES;
#pragma unroll
                for (int j = 0; j < P * P / TA_LOAD_THREADS; ++j) {
                    const int e = lt + TA_LOAD_THREADS * j, i = e >> 5, k = e & 31;
                    float rh, rl, ih, il;
                    split_tf32(qv[j].x - (TA_DELTA && i == k ? 1.0f : 0.0f), rh, rl);      // TA_DELTA: Q - I
                    split_tf32(qv[j].y, ih, il);
                    const int o0 = ta_offset(2 * i, k, TA_KB_B), o1 = ta_offset(2 * i + 1, k, TA_KB_B);
                    ta_store(bh, o0, rh, -ih);
                    ta_store(bl, o0, rl, -il);
                    ta_store(bh, o1, ih, rh);
                    ta_store(bl, o1, il, rl);
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy stores -> UMMA reads
            named_bar_sync(1, TA_LOAD_THREADS);
            if (lt == 0) mbar_arrive(&full_bar[s]);
            ++n;
        }
    } else if (warp == TA_EPI_WARPS) {
        // ===== MMA issuer =====
        if (lane == 0) {
            constexpr uint32_t idesc = idesc_tf32(TA_M, TA_N);
            int n = 0;
            for (long long it = it0; it < it1; ++it) {
                if (ta_item(p, it, ntot).skip) continue;
                const int s = n & 1;
                const uint32_t ph = (n >> 1) & 1;
                mbar_wait(&acc_empty[s], ph ^ 1);
                mbar_wait(&full_bar[s], ph);
                tc_fence_after();
                const uint32_t sa = smem_u32(ring + (size_t)s * TA_STAGE_BYTES);
                const uint32_t tmem_d = tmem_base + (uint32_t)s * TA_N;
#pragma unroll
                for (int kb = 0; kb < 2; ++kb) {
                    const uint64_t ahi = umma_desc_k128(sa + kb * TA_KB_A), alo = umma_desc_k128(sa + TA_A_BYTES + kb * TA_KB_A);
                    const uint64_t bhi = umma_desc_k128(sa + 2 * TA_A_BYTES + kb * TA_KB_B);
                    const uint64_t blo = umma_desc_k128(sa + 2 * TA_A_BYTES + TA_B_BYTES + kb * TA_KB_B);
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {
                        const uint64_t adv = (uint64_t)((ks * 32) >> 4);      // 8 tf32 = 32 bytes along K inside the swizzle row
                        tc_mma_tf32(tmem_d, ahi + adv, blo + adv, idesc, (kb | ks) != 0);
                        tc_mma_tf32(tmem_d, alo + adv, bhi + adv, idesc, 1);
                        tc_mma_tf32(tmem_d, ahi + adv, bhi + adv, idesc, 1);
                    }
                }
                tc_commit(&empty_bar[s]);
                tc_commit(&acc_full[s]);
                ++n;
            }
        }
    } else {
        // ===== epilogue: TMEM lane m = column c0 + m of the tile =====
        int n = 0;
        for (long long it = it0; it < it1; ++it) {
            const TaItem w = ta_item(p, it, ntot);
            if (w.skip) continue;
            const int a = n & 1;
            mbar_wait(&acc_full[a], (n >> 1) & 1);
            tc_fence_after();
            const uint32_t trow = tmem_base + (uint32_t)a * TA_N + ((uint32_t)(warp * 32) << 16);
            float d[TA_N];
#pragma unroll
            for (int c = 0; c < TA_N / 16; ++c) {
                float v[16];
                tc_ld16(trow + c * 16, v);
                tc_ld_wait();
#pragma unroll
                for (int i = 0; i < 16; ++i) d[c * 16 + i] = v[i];
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[a]);
            int I, J;
            pair_blocks(p.nb, round, w.g, I, J);
            cf* base; int ld, c0, ncol;
            if (w.tile < ntx) { base = p.X + (size_t)w.job * p.x_stride; ld = p.L; c0 = w.tile * TA_M; ncol = p.L; }
            else { base = p.Z + (size_t)w.job * p.z_stride; ld = p.nvp; c0 = (w.tile - ntx) * TA_M; ncol = p.nvp; }
            const int col = c0 + warp * 32 + lane;
            if (col < ncol) {
#pragma unroll
                for (int i = 0; i < P; ++i) {
                    cf* dst = base + (size_t)pair_row(I, J, i) * ld + col;
#if TA_DELTA
                    const cf t = __ldcg(dst);                    // this thread owns (i, col): no hazard
                    *dst = cf_make(t.x + d[2 * i], t.y + d[2 * i + 1]);
#else
                    *dst = cf_make(d[2 * i], d[2 * i + 1]);
#endif
                }
            }
            ++n;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2u * TA_N) : "memory");
    }
}

// ---- the tensor-core apply with TMA-fed raw tiles ------------------------------------------------
// bj_apply_tc_kernel above is latency bound (ncu: 3.1 TB/s, tensor pipe 17 %, the loader warps on the long
// scoreboard): its loaders hold one tile of global loads in registers, i.e. ~32 KB in flight per SM, and
// the proxy fence after their shared-memory stores waits for whatever they have in flight, so they cannot
// run ahead.  Here the rows come in by TMA -- two boxes of 16 rows x 128 complex columns per tile, blocks
// I and J of the pair -- into a ring of four RAW 32 KB tiles (128 KB in flight per SM, no registers, no
// thread waiting), and the converter warps read a raw tile from shared memory, split it and write the
// swizzled K-major operands exactly as before.  Shared memory: two A stages (hi + lo, 64 KB each) so that the
// conversion of tile n+1 overlaps the MMAs of tile n, ONE copy of B (the embedding of Q changes once per pair,
// i.e. every ntx + ntz tiles: the converters then wait for the MMAs of the previous tile), two raw tiles.
// (First version: one 96 KB operand stage and four raw tiles -- 119 -> 108 us per round at 50 x 512^2; the
// kernel was then bound by the convert -> MMA -> convert chain of the single stage, not by the loads.)
// Roles (448 threads): warps 0-3 epilogue, warp 4 MMA issuer, warp 5 TMA producer, warps 6-13 converters.
constexpr int TB_RAW_BYTES = P * TA_M * (int)sizeof(cf);                 // 32 KB: 32 rows x 128 columns
constexpr int TB_RAW_STAGES = 2;
constexpr int TB_A_STAGE = 2 * TA_A_BYTES;                               // A_hi, A_lo of one tile: 64 KB
constexpr int TB_OP_BYTES = 2 * TB_A_STAGE + 2 * TA_B_BYTES;             // two A stages + B_hi, B_lo: 160 KB
constexpr int TB_SMEM = TB_OP_BYTES + TB_RAW_STAGES * TB_RAW_BYTES + 1024;
constexpr int TB_CONV_WARPS = 8, TB_CONV_THREADS = TB_CONV_WARPS * 32;
constexpr int TB_THREADS = (TA_EPI_WARPS + 2 + TB_CONV_WARPS) * 32;      // 448
static_assert(TB_SMEM <= 232448, "shared memory of bj_apply_tma_kernel");

__global__ void __launch_bounds__(TB_THREADS, 1)
bj_apply_tma_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_z,
                    LargeParams p, int round, int ntx, int ntot, int njobs) {
    using namespace tcx;
    extern __shared__ __align__(1024) uint8_t tb_smem[];
    __shared__ __align__(8) uint64_t raw_full[TB_RAW_STAGES], raw_empty[TB_RAW_STAGES], op_full[2], op_empty[2], acc_full[2], acc_empty[2];
    __shared__ uint32_t tmem_slot;
    uint8_t* op = (uint8_t*)(((uintptr_t)tb_smem + 1023) & ~(uintptr_t)1023);
    uint8_t* raw = op + TB_OP_BYTES;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    griddep_wait();
    if (warp == TA_EPI_WARPS && lane == 0) {
        for (int s = 0; s < TB_RAW_STAGES; ++s) { mbar_init(&raw_full[s], 1); mbar_init(&raw_empty[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&op_full[a], 1); mbar_init(&op_empty[a], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&acc_full[a], 1); mbar_init(&acc_empty[a], TA_EPI_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(2u * TA_N) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;

    const long long total = (long long)njobs * p.npairs * ntot;
    const long long per = (total + gridDim.x - 1) / gridDim.x;
    const long long it0 = per * blockIdx.x, it1 = min(total, it0 + per);

    if (warp == TA_EPI_WARPS + 1) {
        // ===== TMA producer =====
        if (lane == 0) {
            unsigned n = 0;
            TaCursor cur;
            for (long long it = it0; it < it1; ++it) {
                const TaItem w = ta_item_cached(p, it, ntot, cur);
                if (w.skip) continue;
                const int s = (int)(n % TB_RAW_STAGES);
                mbar_wait(&raw_empty[s], ((n / TB_RAW_STAGES) & 1) ^ 1);
                int I, J;
                pair_blocks(p.nb, round, w.g, I, J);
                const bool isx = w.tile < ntx;
                const CUtensorMap* map = isx ? &map_x : &map_z;
                const int c0f = 2 * TA_M * (isx ? w.tile : w.tile - ntx);         // first float of the tile's columns
                uint8_t* dst = raw + (size_t)s * TB_RAW_BYTES;
                mbar_expect_tx(&raw_full[s], TB_RAW_BYTES);
                tma_load_3d(dst, map, &raw_full[s], c0f, I * BLK, w.job);
                tma_load_3d(dst + TB_RAW_BYTES / 2, map, &raw_full[s], c0f, J * BLK, w.job);
                ++n;
            }
        }
    } else if (warp > TA_EPI_WARPS + 1) {
        // ===== converters: raw tile -> registers (hi/lo split) -> swizzled K-major operands =====
        const int lt = threadIdx.x - (TA_EPI_WARPS + 2) * 32, lw = lt >> 5;
        long long held = -1;                               // pair whose B the operand stage holds
        unsigned n = 0;
        TaCursor cur;
        for (long long it = it0; it < it1; ++it) {
            const TaItem w = ta_item_cached(p, it, ntot, cur);
            if (w.skip) continue;
            const int s = (int)(n % TB_RAW_STAGES);
            const long long pr = (long long)w.job * p.npairs + w.g;
            const bool newq = held != pr;
            cf qv[P * P / TB_CONV_THREADS];
            if (newq) {
                const cf* Q = p.Q + (size_t)w.job * p.g_stride + (size_t)w.g * P * P;
#pragma unroll
                for (int j = 0; j < P * P / TB_CONV_THREADS; ++j) qv[j] = __ldcg(Q + lt + TB_CONV_THREADS * j);
            }
            constexpr int RPW = P / TB_CONV_WARPS, CPL = TA_M / 32;
            cf v[RPW][CPL];
            mbar_wait(&raw_full[s], (n / TB_RAW_STAGES) & 1);
            const cf* rt = reinterpret_cast<const cf*>(raw + (size_t)s * TB_RAW_BYTES);
#pragma unroll
            for (int rr = 0; rr < RPW; ++rr)
#pragma unroll
                for (int j = 0; j < CPL; ++j) v[rr][j] = rt[(lw * RPW + rr) * TA_M + lane + 32 * j];
            const int s2 = (int)(n & 1);
            uint8_t* as = op + (size_t)s2 * TB_A_STAGE;
            mbar_wait(&op_empty[s2], ((n >> 1) & 1) ^ 1);   // the MMAs of tile n - 2 have read this A stage
            if (newq && n > 0) mbar_wait(&op_empty[s2 ^ 1], ((n - 1) >> 1) & 1);    // ... and those of tile n - 1 the old B
            // pair rows k, k + 1 of one column are adjacent in the operand row (8 bytes each): one 16-byte store
            // per pair, and the 8 lanes of a swizzle period cover 128 contiguous bytes (conflict free; the 8-byte
            // stores of bj_apply_tc_kernel are 2-way conflicted: 46 % of its shared-memory wavefronts in ncu)
#pragma unroll
            for (int rp = 0; rp < RPW / 2; ++rp) {
                const int k = lw * RPW + 2 * rp;
#pragma unroll
                for (int j = 0; j < CPL; ++j) {
                    const int m = lane + 32 * j;
                    float4 hi, lo;
                    split_tf32(v[2 * rp][j].x, hi.x, lo.x);
                    split_tf32(v[2 * rp][j].y, hi.y, lo.y);
                    split_tf32(v[2 * rp + 1][j].x, hi.z, lo.z);
                    split_tf32(v[2 * rp + 1][j].y, hi.w, lo.w);
                    const int off = ta_offset(m, k, TA_KB_A);
                    *reinterpret_cast<float4*>(as + off) = hi;
                    *reinterpret_cast<float4*>(as + TA_A_BYTES + off) = lo;
                }
            }
            if (newq) {
                held = pr;
                uint8_t* bh = op + 2 * TB_A_STAGE;
                uint8_t* bl = bh + TA_B_BYTES;
#pragma unroll
                for (int j = 0; j < P * P / TB_CONV_THREADS; ++j) {
                    const int e = lt + TB_CONV_THREADS * j, i = e >> 5, k = e & 31;
                    float rh, rl, ih, il;
                    split_tf32(qv[j].x, rh, rl);
                    split_tf32(qv[j].y, ih, il);
                    const int o0 = ta_offset(2 * i, k, TA_KB_B), o1 = ta_offset(2 * i + 1, k, TA_KB_B);
                    ta_store(bh, o0, rh, -ih);
                    ta_store(bl, o0, rl, -il);
                    ta_store(bh, o1, ih, rh);
                    ta_store(bl, o1, il, rl);
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy stores -> UMMA reads
            named_bar_sync(1, TB_CONV_THREADS);
            if (lt == 0) { mbar_arrive(&op_full[s2]); mbar_arrive(&raw_empty[s]); }
            ++n;
        }
    } else if (warp == TA_EPI_WARPS) {
        // ===== MMA issuer =====
        if (lane == 0) {
            constexpr uint32_t idesc = idesc_tf32(TA_M, TA_N);
            const uint32_t sa = smem_u32(op);
            unsigned n = 0;
            TaCursor cur;
            for (long long it = it0; it < it1; ++it) {
                if (ta_item_cached(p, it, ntot, cur).skip) continue;
                const int a = n & 1;
                mbar_wait(&acc_empty[a], ((n >> 1) & 1) ^ 1);
                mbar_wait(&op_full[a], (n >> 1) & 1);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + (uint32_t)a * TA_N;
                const uint32_t sA = sa + (uint32_t)a * TB_A_STAGE, sB = sa + 2 * TB_A_STAGE;
#pragma unroll
                for (int kb = 0; kb < 2; ++kb) {
                    const uint64_t ahi = umma_desc_k128(sA + kb * TA_KB_A), alo = umma_desc_k128(sA + TA_A_BYTES + kb * TA_KB_A);
                    const uint64_t bhi = umma_desc_k128(sB + kb * TA_KB_B);
                    const uint64_t blo = umma_desc_k128(sB + TA_B_BYTES + kb * TA_KB_B);
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {
                        const uint64_t adv = (uint64_t)((ks * 32) >> 4);
                        tc_mma_tf32(tmem_d, ahi + adv, blo + adv, idesc, (kb | ks) != 0);
                        tc_mma_tf32(tmem_d, alo + adv, bhi + adv, idesc, 1);
                        tc_mma_tf32(tmem_d, ahi + adv, bhi + adv, idesc, 1);
                    }
                }
                tc_commit(&op_empty[a]);
                tc_commit(&acc_full[a]);
                ++n;
            }
        }
    } else {
        // ===== epilogue: TMEM lane m = column c0 + m of the tile =====
        unsigned n = 0;
        TaCursor cur;
        for (long long it = it0; it < it1; ++it) {
            const TaItem w = ta_item_cached(p, it, ntot, cur);
            if (w.skip) continue;
            const int a = n & 1;
            mbar_wait(&acc_full[a], (n >> 1) & 1);
            tc_fence_after();
            const uint32_t trow = tmem_base + (uint32_t)a * TA_N + ((uint32_t)(warp * 32) << 16);
            float d[TA_N];
#pragma unroll
            for (int c = 0; c < TA_N / 16; ++c) {
                float v[16];
                tc_ld16(trow + c * 16, v);
                tc_ld_wait();
#pragma unroll
                for (int i = 0; i < 16; ++i) d[c * 16 + i] = v[i];
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[a]);
            int I, J;
            pair_blocks(p.nb, round, w.g, I, J);
            cf* base; int ld, c0, ncol;
            if (w.tile < ntx) { base = p.X + (size_t)w.job * p.x_stride; ld = p.L; c0 = w.tile * TA_M; ncol = p.L; }
            else { base = p.Z + (size_t)w.job * p.z_stride; ld = p.nvp; c0 = (w.tile - ntx) * TA_M; ncol = p.nvp; }
            const int col = c0 + warp * 32 + lane;
            if (col < ncol) {
#pragma unroll
                for (int i = 0; i < P; ++i)
                    base[(size_t)pair_row(I, J, i) * ld + col] = cf_make(d[2 * i], d[2 * i + 1]);
            }
            ++n;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2u * TA_N) : "memory");
    }
}

typedef CUresult (*EncodeTiledFnL)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
// complex64 rows as fp32: tensor [njobs][rows][2 ncols] with the given strides (in complex numbers), box
// [1][16 rows][256 floats], no swizzle (raw tiles)
static int make_row_map(CUtensorMap* map, cf* base, int njobs, int rows, int ncols, int64_t row_stride, int64_t job_stride) {
    static EncodeTiledFnL enc = nullptr;
    if (!enc) {
        void* fp = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            enc = (EncodeTiledFnL)fp;
    }
    if (!enc) return 1;
    cuuint64_t dims[3] = {(cuuint64_t)(2 * ncols), (cuuint64_t)rows, (cuuint64_t)njobs};
    cuuint64_t strides[2] = {(cuuint64_t)row_stride * 8, (cuuint64_t)job_stride * 8};
    cuuint32_t box[3] = {(cuuint32_t)(2 * TA_M), (cuuint32_t)BLK, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : 2;
}

// ---- the pair Grams on the tensor cores ---------------------------------------------------------
// G = Xp Xp^H of a block pair is a REAL product of the rows as they lie in memory: with r_i the
// interleaved (re, im) row i (2L floats) and r~_i = (im, -re) of the same row,
//   Re G_ij = r_i . r_j,      Im G_ij = r~_i . r_j        (G_ij = sum_c x_ic conj(x_jc)).
// The contraction index is the contiguous one, so both operands are K-major as they come.  One MMA
// tile carries TWO pairs:  A = [r(1) | r(2) | r~(1) | r~(2)]  (4 x 32 rows = M 128),  B = rows 0..63
// of the same tile (N = 64), D = A B^T:  Re G(1) = D[0:32, 0:32], Re G(2) = D[32:64, 32:64],
// Im G(1) = D[64:96, 0:32], Im G(2) = D[96:128, 32:64] (the cross-pair blocks are computed and
// dropped: the tensor pipe has the time).  3xTF32: hi.lo + lo.hi + hi.hi per k-step, fp32
// accumulation in TMEM over the whole row length.  The Gram only steers the rotation angles -- the
// cyclic pass and the apply keep the state unitary -- so the per-MMA truncation of the accumulator
// (diagonal entries come out ~1e-5 low at L = 2048) is harmless here; it is not chunked.
// Loader warps read 32 complex columns of the 64 rows per stage (coalesced 16-byte loads), split and
// write r and r~ straight into the K-major SWIZZLE_128B layout (row at r * 128 B per 32-float
// k-block, 16-byte chunk j at j ^ (r & 7)): 16 KB read, 64 KB written, 24 MMAs per stage.
// Persistent, one CTA per SM, items (job, two pairs) strided over the CTAs; the Grams go to global
// memory as two planes of 32 x 32 floats per pair (p.gpart) for the pass (bj_gram_evd_kernel with
// p.gram_tc set: no Gram of its own).
// Used for rows of at least 1024 entries (chi >= 512).  Measured on B200 (profiles/r2_ab_tc_gram.txt): per
// round, 4 x 2048^2: Gram + pass 120 us fused on FFMA -> 80 + 22 us (solve 766 -> 708 ms, the chi = 1024 leg
// of bench.py 7.7 -> 8.4 applications/s); 50 x 512^2: 119 us -> 64 + 51 us and the chi = 256 circuit
// 908 -> 872 applications/s (a third launch per round), so the fused FFMA kernel keeps the shorter rows.
// What bounds it (ncu, profiles/r2_gram_tc_ncu_full.txt: tensor pipe 36 %, DRAM 1.6 TB/s, loader warps on the
// long scoreboard): fence.proxy.async is MEMBAR.ALL.CTA + FENCE.VIEW.ASYNC, i.e. it waits for every load
// the thread has in flight, so a loader cannot prefetch across its own fence; with one chunk of loads in
// flight per loader group (registers) an SM has 48 KB outstanding.  Raw row tiles by TMA into a deeper ring with
// converter warps (what bj_apply_tma_kernel does) were tried here too and bought nothing (66 / 83 us): with the
// loads out of the way the kernel sits on its shared-memory traffic.  The MMAs themselves are
// shared-memory heavy (per k-step and pair 3 x (64 + 32) rows x 32 B = 9.2 KB of operand reads: a 32 x 32
// Gram has no reuse to speak of), which puts the floor near 35 us per round at 50 x 512^2.
constexpr int TG_KC = 32;                               // complex columns per stage (2 k-blocks of 32 floats)
constexpr int TG_KB = 128 * 128;                        // bytes of one k-block of the 128-row operand
constexpr int TG_OP_BYTES = 2 * TG_KB;                  // hi (or lo) of a stage
constexpr int TG_STAGE_BYTES = 2 * TG_OP_BYTES;         // 64 KB
constexpr int TG_STAGES = 3;
constexpr int TG_SMEM = TG_STAGES * TG_STAGE_BYTES + 1024;
constexpr int TG_N = 64;
constexpr int TG_GROUPS = TG_STAGES;                   // loader groups: group q owns stage q (chunks q, q + 3, ...)
constexpr int TG_GT = 64;                               // threads per loader group
constexpr int TG_THREADS = (TA_EPI_WARPS + 1) * 32 + TG_GROUPS * TG_GT;      // 352

struct TgItem { int job, g0, npair; bool skip; };
__device__ __forceinline__ TgItem tg_item(const LargeParams& p, long long it, int ngroups) {
    TgItem w;
    w.job = (int)(it / ngroups);
    const int grp = (int)(it - (long long)w.job * ngroups);
    w.g0 = 2 * grp;
    w.npair = min(2, p.npairs - w.g0);
    w.skip = !p.misc[w.job].active;
    return w;
}

__global__ void __launch_bounds__(TG_THREADS, 1) bj_gram_tc_kernel(LargeParams p, int round, int njobs) {
    using namespace tcx;
    extern __shared__ __align__(1024) uint8_t tg_smem[];
    __shared__ __align__(8) uint64_t full_bar[TG_STAGES], empty_bar[TG_STAGES], acc_full[2], acc_empty[2];
    __shared__ uint32_t tmem_slot;
    uint8_t* ring = (uint8_t*)(((uintptr_t)tg_smem + 1023) & ~(uintptr_t)1023);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    griddep_wait();
    if (warp == TA_EPI_WARPS && lane == 0) {
        for (int s = 0; s < TG_STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&acc_full[a], 1); mbar_init(&acc_empty[a], TA_EPI_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(2u * TG_N) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;

    const int ngroups = (p.npairs + 1) / 2;
    const long long total = (long long)njobs * ngroups;
    const int nchunk = (p.L + TG_KC - 1) / TG_KC;

    if (warp > TA_EPI_WARPS) {
        // ===== loaders: TG_GROUPS groups of 64 threads; group q fills stage q, i.e. every TG_GROUPS-th chunk =====
        // (as many groups as stages: a stage has ONE writer, so its empty/full phases advance in step with that
        // group -- four groups on three stages let a group lap a slower one and overwrite a stage in use)
        // fence.proxy.async compiles to MEMBAR.ALL.CTA + FENCE.VIEW.ASYNC: it waits for every load the thread
        // has in flight, so a thread cannot prefetch across its own fence (the first version of this kernel
        // did, 4 chunks deep, and ran at the latency of one chunk at a time: 8 of 8 loader warps on the long
        // scoreboard in ncu).  Here a group issues the loads of its NEXT chunk right after signalling the
        // current one and the other groups' loads are in flight while it splits, stores and fences.
        const int lt = threadIdx.x - (TA_EPI_WARPS + 1) * 32;
        const int grp = lt / TG_GT, gt = lt % TG_GT;
        const int c4 = gt & 7, rsub = gt >> 3;              // 8 threads per row (2 x 16 bytes each), rows rsub + 8 rr
        constexpr int RR = 32 / (TG_GT / 8);                // row slots per pair and thread (4)
        float4 v[2][RR][2];
        // chunk number n (counted over all items of this CTA) -> (item, chunk); group q owns n = q mod TG_GROUPS
        long long it = blockIdx.x;
        int ch = -1;                                        // position of the walk
        unsigned n = 0;                                     // number of the chunk (it, ch) once ch >= 0
        TgItem w = {};
        bool have = false;
        auto next_item = [&]() {                            // first non-skipped item at or after `it`
            for (; it < total; it += gridDim.x) {
                w = tg_item(p, it, ngroups);
                if (!w.skip) return true;
            }
            return false;
        };
        // advance to this group's next chunk; returns false at the end
        auto advance = [&](int steps) {
            while (steps > 0) {
                if (ch < 0) {
                    if (!next_item()) return false;
                    ch = 0;
                    --steps;
                    if (have) ++n; else have = true;
                    continue;
                }
                const int room = nchunk - 1 - ch;
                if (room >= steps) { ch += steps; n += steps; steps = 0; }
                else { steps -= room; ch = -1; n += room; it += gridDim.x; }
            }
            return true;
        };
        auto fetch = [&]() {
            int I0, J0, I1, J1;
            pair_blocks(p.nb, round, min(w.g0, p.npairs - 1), I0, J0);
            pair_blocks(p.nb, round, min(w.g0 + 1, p.npairs - 1), I1, J1);
            const cf* X = p.X + (size_t)w.job * p.x_stride;
#pragma unroll
            for (int q = 0; q < 2; ++q)
#pragma unroll
                for (int rr = 0; rr < RR; ++rr) {
                    const cf* row = X + (size_t)pair_row(q ? I1 : I0, q ? J1 : J0, rsub + 8 * rr) * p.L;
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int c0 = ch * TG_KC + 2 * (c4 + 8 * h);
                        const bool ok = q < w.npair && c0 + 1 < p.L;
                        v[q][rr][h] = ok ? __ldcg(reinterpret_cast<const float4*>(row + c0)) : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                }
        };
        bool live = advance(grp + 1);                       // chunk number grp
        if (live) fetch();
        while (live) {
            const int s = (int)(n % TG_STAGES);
            uint8_t* st = ring + (size_t)s * TG_STAGE_BYTES;
            mbar_wait(&empty_bar[s], ((n / TG_STAGES) & 1) ^ 1);
#pragma unroll
            for (int q = 0; q < 2; ++q)
#pragma unroll
                for (int rr = 0; rr < RR; ++rr)
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        // floats 4 (c4 + 8 h) .. + 3 of the stage's 64: k-block h, 16-byte chunk c4
                        const int kb = h, j = c4;
                        const int r0 = 32 * q + rsub + 8 * rr, r1 = 64 + r0;
                        const float4 x = v[q][rr][h];
                        float4 hi, lo, thi, tlo;
                        split_tf32(x.x, hi.x, lo.x); split_tf32(x.y, hi.y, lo.y);
                        split_tf32(x.z, hi.z, lo.z); split_tf32(x.w, hi.w, lo.w);
                        thi = make_float4(hi.y, -hi.x, hi.w, -hi.z);      // r~ = (im, -re)
                        tlo = make_float4(lo.y, -lo.x, lo.w, -lo.z);
                        const int o0 = kb * TG_KB + r0 * 128 + ((j ^ (r0 & 7)) << 4);
                        const int o1 = kb * TG_KB + r1 * 128 + ((j ^ (r1 & 7)) << 4);
                        *reinterpret_cast<float4*>(st + o0) = hi;
                        *reinterpret_cast<float4*>(st + TG_OP_BYTES + o0) = lo;
                        *reinterpret_cast<float4*>(st + o1) = thi;
                        *reinterpret_cast<float4*>(st + TG_OP_BYTES + o1) = tlo;
                    }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            named_bar_sync(1 + grp, TG_GT);
            if (gt == 0) mbar_arrive(&full_bar[s]);
            live = advance(TG_GROUPS);
            if (live) fetch();
        }
    } else if (warp == TA_EPI_WARPS) {
        // ===== MMA issuer =====
        if (lane == 0) {
            constexpr uint32_t idesc = idesc_tf32(128, TG_N);
            unsigned n = 0;
            int na = 0;
            for (long long it = blockIdx.x; it < total; it += gridDim.x) {
                if (tg_item(p, it, ngroups).skip) continue;
                const int a = na & 1;
                mbar_wait(&acc_empty[a], ((na >> 1) & 1) ^ 1);
                const uint32_t tmem_d = tmem_base + (uint32_t)a * TG_N;
                for (int ch = 0; ch < nchunk; ++ch, ++n) {
                    const int s = (int)(n % TG_STAGES);
                    mbar_wait(&full_bar[s], (uint32_t)(n / TG_STAGES) & 1);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(ring + (size_t)s * TG_STAGE_BYTES);
#pragma unroll
                    for (int kb = 0; kb < 2; ++kb) {
                        const uint64_t hi = umma_desc_k128(sa + kb * TG_KB), lo = umma_desc_k128(sa + TG_OP_BYTES + kb * TG_KB);
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks) {
                            const uint64_t adv = (uint64_t)((ks * 32) >> 4);
                            tc_mma_tf32(tmem_d, hi + adv, lo + adv, idesc, (ch | kb | ks) != 0);
                            tc_mma_tf32(tmem_d, lo + adv, hi + adv, idesc, 1);
                            tc_mma_tf32(tmem_d, hi + adv, hi + adv, idesc, 1);
                        }
                    }
                    tc_commit(&empty_bar[s]);
                }
                tc_commit(&acc_full[a]);
                ++na;
            }
        }
    } else {
        // ===== epilogue: warp 0 Re G(1), warp 1 Re G(2), warp 2 Im G(1), warp 3 Im G(2) =====
        int na = 0;
        for (long long it = blockIdx.x; it < total; it += gridDim.x) {
            const TgItem w = tg_item(p, it, ngroups);
            if (w.skip) continue;
            const int a = na & 1;
            mbar_wait(&acc_full[a], (na >> 1) & 1);
            tc_fence_after();
            const int q = warp & 1, plane = warp >> 1;
            const uint32_t trow = tmem_base + (uint32_t)a * TG_N + (uint32_t)(32 * q) + ((uint32_t)(warp * 32) << 16);
            float d[32];
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                float v[16];
                tc_ld16(trow + c * 16, v);
                tc_ld_wait();
#pragma unroll
                for (int i = 0; i < 16; ++i) d[c * 16 + i] = v[i];
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[a]);
            if (q < w.npair) {
                float* out = reinterpret_cast<float*>(p.gpart) + (((size_t)w.job * p.npairs + w.g0 + q) * 2 + plane) * (P * P) + lane * P;
#pragma unroll
                for (int c = 0; c < 8; ++c)
                    reinterpret_cast<float4*>(out)[c] = make_float4(d[4 * c], d[4 * c + 1], d[4 * c + 2], d[4 * c + 3]);
            }
            ++na;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2u * TG_N) : "memory");
    }
}

// R <- 3/2 I - 1/2 R   (Newton-Schulz factor)
__global__ void bj_ns_kernel(cf* R, int64_t stride, int n) {
    cf* r = R + (size_t)blockIdx.y * stride;
    const size_t tot = (size_t)n * n;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < tot; e += (size_t)gridDim.x * blockDim.x) {
        size_t i = e / n, j = e - i * n;
        cf v = r[e];
        r[e] = cf_make((i == j ? 1.5f : 0.f) - 0.5f * v.x, -0.5f * v.y);
    }
}

__global__ void bj_sweep_end_kernel(LargeParams p, int njobs) {
    griddep_wait();
    int job = blockIdx.x * blockDim.x + threadIdx.x;
    if (job >= njobs) return;
    Misc& m = p.misc[job];
    if (m.active) {
        m.sweeps += 1; m.active = m.rot > 0 ? 1 : 0; m.rot = 0;
        m.gmax = __uint_as_float(m.gmax_next); m.gmax_next = 0u;
    }
}

// sigma_i = |X_i| ; padding rows (those whose Z row lives in the padded columns) get -1
__global__ void __launch_bounds__(LT) bj_sigma_kernel(LargeParams p) {
    const int job = blockIdx.y;
    const int row = blockIdx.x * (LT / 32) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= p.nvp) return;
    const cf* x = p.X + (size_t)job * p.x_stride + (size_t)row * p.L;
    const cf* z = p.Z + (size_t)job * p.z_stride + (size_t)row * p.nvp;
    float s2 = 0.f, pad = 0.f;
    for (int c = lane; c < p.L; c += 32) s2 += cf_abs2(x[c]);
    for (int c = p.nv + lane; c < p.nvp; c += 32) pad += cf_abs2(z[c]);
    s2 = warp_sum(s2); pad = warp_sum(pad);
    if (lane == 0) p.sigma[(size_t)job * p.s_stride + row] = pad > 0.5f ? -1.0f : sqrtf(s2);
}

__global__ void bj_rank_kernel(LargeParams p) {
    const int job = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.nvp) return;
    const float* s = p.sigma + (size_t)job * p.s_stride;
    float si = s[i];
    int rk = 0;
    for (int j = 0; j < p.nvp; ++j) { float sj = s[j]; rk += (sj > si) || (sj == si && j < i); }
    p.perm[(size_t)job * p.s_stride + rk] = i;
}

struct OutParams {
    const mpsb_gate2_desc* descs; int nbatch;
    cf* left; int64_t left_stride; cf* right; int64_t right_stride;
    float* svals; int64_t svals_stride; int32_t* info;
    int k, lc;
};

__global__ void __launch_bounds__(LT) bj_write_kernel(LargeParams p, OutParams o) {
    const int job = blockIdx.y;
    cf *left, *right; float* sv;
    if (o.descs) {
        int di = job / o.nbatch, bi = job % o.nbatch;
        const mpsb_gate2_desc d = o.descs[di];
        left = (cf*)d.out_l + (size_t)bi * d.bs_out_l;
        right = (cf*)d.out_r + (size_t)bi * d.bs_out_r;
        sv = d.svals ? d.svals + (size_t)bi * d.bs_svals : nullptr;
    } else {
        left = o.left + (size_t)job * o.left_stride;
        right = o.right + (size_t)job * o.right_stride;
        sv = o.svals ? o.svals + (size_t)job * o.svals_stride : nullptr;
    }
    const cf* X = p.X + (size_t)job * p.x_stride;
    const cf* Z = p.Z + (size_t)job * p.z_stride;
    const int* perm = p.perm + (size_t)job * p.s_stride;
    const float* sig = p.sigma + (size_t)job * p.s_stride;
    const int k = o.k, nv = p.nv, L = p.L;
    const size_t stride = (size_t)gridDim.x * blockDim.x, t0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (o.lc) {
        for (size_t e = t0; e < (size_t)k * L; e += stride) { size_t j = e / L, c = e - j * L; right[e] = X[(size_t)perm[j] * L + c]; }
        for (size_t e = t0; e < (size_t)nv * k; e += stride) { size_t a = e / k, j = e - a * k; left[e] = cf_conj(Z[(size_t)perm[j] * p.nvp + a]); }
    } else {
        for (size_t e = t0; e < (size_t)L * k; e += stride) { size_t a = e / k, j = e - a * k; left[e] = X[(size_t)perm[j] * L + a]; }
        for (size_t e = t0; e < (size_t)k * nv; e += stride) { size_t j = e / nv, b = e - j * nv; right[e] = cf_conj(Z[(size_t)perm[j] * p.nvp + b]); }
    }
    const int mn = nv < L ? nv : L;
    if (sv) for (size_t j = t0; j < (size_t)mn; j += stride) sv[j] = fmaxf(sig[perm[j]], 0.f);
    if (o.info && t0 == 0) {
        o.info[2 * job] = p.misc[job].active ? 1 : 0;      // still rotating at the sweep limit
        o.info[2 * job + 1] = p.misc[job].sweeps;
    }
}

struct LargeLayout { int nvp, nb, npairs; size_t z, g, s, misc, m0, total; };

LargeLayout large_layout(int nv, int L) {
    (void)L;
    LargeLayout lo;
    lo.nvp = (nv + P - 1) / P * P;
    lo.nb = lo.nvp / BLK;
    lo.npairs = lo.nb / 2;
    lo.z = align_up((size_t)lo.nvp * lo.nvp, 16);
    lo.g = align_up((size_t)lo.npairs * P * P, 16);
    lo.s = align_up((size_t)lo.nvp, 16);              // floats / ints, counted in cf units below
    lo.misc = 16;
    lo.m0 = align_up((size_t)nv * L, 16);
    // Z, R, Z2 (3 z) + G, Q (2 g) + sigma|perm (s cf = 2 s words) + misc + copy of the input
    lo.total = 3 * lo.z + 2 * lo.g + lo.s + lo.misc + lo.m0;
    return lo;
}

// ---------------------------------------------------------------------------------------
// Host driver.  A solve is a resumable object (LargeRun): begin -> enqueue sweeps -> finish, so
// that several shape groups of one circuit layer can be in flight on different streams at once
// (mpsb_apply_gate2_layer): a group of one or two matrices is bound by the latency of its ~600
// dependent rounds (~27 us each), not by throughput, and hides completely behind the layer's
// main group.  Convergence is read back without ever blocking a stream: every CHECK_EVERY sweeps
// the per-job flags are copied to pinned host memory behind an event; the host polls the events
// and keeps up to MAX_CHECKS checks (sweeps already queued behind them) outstanding per run.
// ---------------------------------------------------------------------------------------
constexpr int CHECK_EVERY = 2;
constexpr int MAX_CHECKS = 2;
constexpr int PIN_SLOTS = 9;                 // 8 pool streams + the caller's stream

struct PinSlot {
    Misc* host = nullptr;                    // [MAX_CHECKS][jobs_cap]
    size_t jobs_cap = 0;
    cudaEvent_t ev[MAX_CHECKS] = {nullptr, nullptr};
};
static PinSlot g_pin[MPSB_MAX_DEVICES][PIN_SLOTS];      // guarded by mpsb_lib_mutex()

struct LargeRun {
    LargeParams p;
    LargeLayout lo;
    OutParams o;
    cf *Rbuf, *Z2, *M0;
    int njobs, nv, L, nrounds, ntx, ntz, max_outer, skip, nt_cta, pdl, gram_stages, gram_threads, tc_apply, tc_gram, tma_apply;
    CUtensorMap map_x, map_z;                // rows of X and Z for bj_apply_tma_kernel
    cudaStream_t st;
    PinSlot* pin;
    int sweeps_queued = 0;
    int checks_out = 0, check_head = 0;      // ring of outstanding checks
    bool done = false;
};

static int large_begin(LargeRun& r, cf* X, int64_t x_job_stride, int njobs, int nv, int L, int k,
                       int left_canonical, const mpsb_gate2_desc* descs, int nbatch,
                       cf* left, int64_t left_stride, cf* right, int64_t right_stride,
                       float* svals, int64_t svals_stride, int32_t* info, cf* work,
                       cudaStream_t st, int pin_slot) {
    MPSB_ARG(work != nullptr, "svd_large: workspace missing");
    MPSB_ARG(njobs <= 65535, "svd_large: njobs %d > 65535", njobs);
    MPSB_ARG(pin_slot >= 0 && pin_slot < PIN_SLOTS, "svd_large: bad pin slot");
    LargeLayout lo = large_layout(nv, L);
    MPSB_ARG(x_job_stride >= (int64_t)lo.nvp * L, "svd_large: X stride too small for the padded rows");
    LargeParams& p = r.p;
    p.X = X; p.x_stride = x_job_stride;
    // workspace: [njobs][Z] [njobs][flags] [njobs][Q] [njobs][sigma|perm] [njobs][misc] R Z2 M0
    cf* w = work;
    p.Z = w; p.z_stride = (int64_t)lo.z; w += lo.z * njobs;
    p.rotflag = (int*)w; p.gcount = p.rotflag + (size_t)njobs * lo.npairs; w += lo.g * njobs;   // (region of the former G buffer)
    p.Q = w; w += lo.g * njobs;
    p.g_stride = (int64_t)lo.g;
    p.sigma = (float*)w; p.perm = (int*)((float*)w + (size_t)lo.s * njobs); p.s_stride = (int64_t)lo.s; w += lo.s * njobs;
    p.misc = (Misc*)w; w += lo.misc * njobs;
    r.Rbuf = w; w += lo.z * njobs;
    // split Gram for launches with fewer pair-CTAs than SMs; the partial sums borrow R (used only at the end)
    p.gpart = r.Rbuf;
    p.nsplit = 1;
    {
        const int ntile_all = (L + GK - 1) / GK;
        for (int ns = 4; ns > 1; ns >>= 1)
            if (ns <= ntile_all && (long long)njobs * lo.npairs * ns <= 148 && (size_t)lo.npairs * ns * P * P <= lo.z) { p.nsplit = ns; break; }
        if (const char* e = mpsb_env("MPSB_LARGE_NSPLIT")) {   // timing experiments only
            int ns = atoi(e);
            if (ns >= 1 && ns <= ntile_all && (size_t)lo.npairs * ns * P * P <= lo.z) p.nsplit = ns;
        }
    }
    r.Z2 = w; w += lo.z * njobs;
    r.M0 = w;
    p.nv = nv; p.L = L; p.nvp = lo.nvp; p.nb = lo.nb; p.npairs = lo.npairs;
    // the Gram entries carry rounding noise ~ eps*sqrt(L)*sqrt(G_ii G_jj): keep the threshold above it
    float tol = 3e-6f;
    float floor_ = 4.0f * 5.96e-8f * sqrtf((float)L);
    if (floor_ > tol) tol = floor_;
    p.tol2 = tol * tol;
    float eta = ABS_ETA;
    if (const char* e = mpsb_env("MPSB_LARGE_ETA")) eta = (float)atof(e);          // experiments only
    p.eta2 = eta * eta;
    p.clk = nullptr;
    if (mpsb_env("MPSB_LARGE_CLOCKS")) {                        // profiling only: one 64-byte buffer per process
        static long long* clk_buf = nullptr;
        if (!clk_buf) MPSB_CUDA(cudaMalloc((void**)&clk_buf, 8 * sizeof(long long)));
        p.clk = clk_buf;
        MPSB_CUDA(cudaMemsetAsync(p.clk, 0, 8 * sizeof(long long), st));
    }
    r.lo = lo; r.njobs = njobs; r.nv = nv; r.L = L; r.st = st;
    r.nrounds = lo.nb > 2 ? lo.nb - 1 : 1;
    r.ntx = (L + CT - 1) / CT; r.ntz = (lo.nvp + CT - 1) / CT;
    // column tiles per apply CTA: as many as keep one full wave of CTAs (3 per SM) in the launch;
    // a small batch is latency bound and wants the shortest CTAs
    r.nt_cta = 1;
    for (int nt = NT_APPLY; nt > 1; nt >>= 1)
        if ((long long)njobs * lo.npairs * ((r.ntx + r.ntz + nt - 1) / nt) >= 3 * 148) { r.nt_cta = nt; break; }
    // Gram tile ring: two stages when several CTAs share an SM (they hide each other's loads), four
    // when the launch has fewer CTAs than SMs
    r.gram_stages = (long long)njobs * lo.npairs >= 2 * 148 ? 2 : GRAM_MAX_STAGES;     // (no measurable effect either way)
    if (const char* e = mpsb_env("MPSB_LARGE_STAGES")) r.gram_stages = atoi(e) >= 2 && atoi(e) <= GRAM_MAX_STAGES ? atoi(e) : r.gram_stages;
    // 512 threads do not help a launch with fewer CTAs than SMs either (measured: a single 256 x 256
    // solve 9.4 ms against 8.9 ms): one pair's Gram is bound by the FFMA issue rate of ONE SM, and the
    // pass by its ~1 200-cycle dependent chain per rotation set (parameters 500, update 600)
    r.gram_threads = 256;
    if (const char* e = mpsb_env("MPSB_LARGE_GTHREADS")) r.gram_threads = atoi(e) == 512 ? 512 : 256;   // experiments only
    r.pdl = 1;
    if (const char* e = mpsb_env("MPSB_LARGE_NT")) r.nt_cta = atoi(e) > 0 ? atoi(e) : r.nt_cta;   // timing experiments only
    if (const char* e = mpsb_env("MPSB_LARGE_PDL")) r.pdl = atoi(e);
    // the tensor-core apply (bj_apply_tc_kernel) is persistent, one CTA per SM: taken when there are
    // at least two tiles per SM, otherwise the FFMA kernel (a single 256 x 256 matrix is 32 tiles)
    r.tc_apply = (long long)njobs * lo.npairs * (r.ntx + r.ntz) >= 2 * 148;
    if (const char* e = mpsb_env("MPSB_LARGE_TC_APPLY")) r.tc_apply = atoi(e) != 0;       // A/B timing
    // the TMA-fed variant of the tensor-core apply needs 16-byte aligned rows (bj_apply_tma_kernel)
    r.tma_apply = 0;
    // (and rows at least one box wide: shorter ones keep the loader-warp kernel)
    if (r.tc_apply && L % 2 == 0 && L >= TA_M && lo.nvp >= TA_M && x_job_stride % 2 == 0 &&
        (reinterpret_cast<uintptr_t>(X) & 15) == 0 && (reinterpret_cast<uintptr_t>(p.Z) & 15) == 0) {
        if (make_row_map(&r.map_x, X, njobs, lo.nvp, L, L, x_job_stride) == 0 &&
            make_row_map(&r.map_z, p.Z, njobs, lo.nvp, lo.nvp, lo.nvp, (int64_t)lo.z) == 0)
            r.tma_apply = 1;
    }
    if (const char* e = mpsb_env("MPSB_LARGE_TMA_APPLY")) r.tma_apply = r.tma_apply && atoi(e) != 0;   // A/B timing
    // ... and so is the tensor-core Gram (bj_gram_tc_kernel; 16-byte row loads: even row length and stride)
    r.tc_gram = r.tc_apply && L >= 1024 && L % 2 == 0 && x_job_stride % 2 == 0 && (reinterpret_cast<uintptr_t>(X) & 15) == 0;
    if (const char* e = mpsb_env("MPSB_LARGE_TC_GRAM"))                                          // A/B timing
        r.tc_gram = atoi(e) != 0 && L % 2 == 0 && x_job_stride % 2 == 0 && (reinterpret_cast<uintptr_t>(X) & 15) == 0;
    p.gram_tc = r.tc_gram ? 1 : 0;
    if (r.tc_gram) p.nsplit = 1;
    r.skip = 0; r.max_outer = MAX_OUTER;
    if (const char* e = mpsb_env("MPSB_LARGE_SKIP")) r.skip = atoi(e);
    if (const char* e = mpsb_env("MPSB_LARGE_SWEEPS")) r.max_outer = atoi(e);
    OutParams& o = r.o;
    o.descs = descs; o.nbatch = nbatch > 0 ? nbatch : 1;
    o.left = left; o.left_stride = left_stride; o.right = right; o.right_stride = right_stride;
    o.svals = svals; o.svals_stride = svals_stride; o.info = info; o.k = k; o.lc = left_canonical;
    // pinned read-back area of this slot (a slot is only ever written from one stream at a time)
    int dev = 0;
    if (int rc = mpsb_current_device(&dev)) return rc;
    PinSlot& pin = g_pin[dev][pin_slot];
    if (pin.jobs_cap < (size_t)njobs) {
        if (pin.host) { MPSB_CUDA(cudaDeviceSynchronize()); cudaFreeHost(pin.host); pin.host = nullptr; }
        pin.jobs_cap = (size_t)njobs > 4096 ? (size_t)njobs : 4096;
        MPSB_CUDA(cudaHostAlloc((void**)&pin.host, MAX_CHECKS * pin.jobs_cap * sizeof(Misc), cudaHostAllocDefault));
    }
    for (int i = 0; i < MAX_CHECKS; ++i)
        if (!pin.ev[i]) MPSB_CUDA(cudaEventCreateWithFlags(&pin.ev[i], cudaEventDisableTiming));
    r.pin = &pin;
    static bool attrs_set[MPSB_MAX_DEVICES] = {};
    bool& attrs = attrs_set[dev];
    if (!attrs) {
        MPSB_CUDA(cudaFuncSetAttribute(bj_apply_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, APPLY_SMEM));
        MPSB_CUDA(cudaFuncSetAttribute(bj_apply_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TA_SMEM));
        MPSB_CUDA(cudaFuncSetAttribute(bj_gram_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TG_SMEM));
        MPSB_CUDA(cudaFuncSetAttribute(bj_apply_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TB_SMEM));
        MPSB_CUDA(cudaFuncSetAttribute(bj_gram_evd_kernel<256, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, gram_smem_bytes(GRAM_MAX_STAGES)));
        MPSB_CUDA(cudaFuncSetAttribute(bj_gram_evd_kernel<256, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, gram_smem_bytes(GRAM_MAX_STAGES)));
        MPSB_CUDA(cudaFuncSetAttribute(bj_gram_evd_kernel<512, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, gram_smem_bytes(GRAM_MAX_STAGES)));
        attrs = true;
    }
    // keep the input: the weighted factor is recomputed from it at the end (see large_finish)
    MPSB_CUDA(cudaMemcpy2DAsync(r.M0, lo.m0 * sizeof(cf), X, (size_t)x_job_stride * sizeof(cf),
                                (size_t)nv * L * sizeof(cf), njobs, cudaMemcpyDeviceToDevice, st));
    bj_init_kernel<<<dim3(148, njobs), 256, 0, st>>>(p);
    MPSB_LAUNCH_CHECK("bj_init_kernel");
    return 0;
}

// one sweep onto the run's stream, followed every CHECK_EVERY sweeps by a flag read-back
static int large_enqueue_sweep(LargeRun& r) {
    const LargeParams& p = r.p;
    const LargeLayout& lo = r.lo;
    cudaStream_t st = r.st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = r.pdl ? 1 : 0;
    cudaLaunchConfig_t cfg = {};
    cfg.stream = st; cfg.attrs = attr; cfg.numAttrs = 1;
    // The tensor-core accumulator truncates toward zero: every apply shrinks the rows of X and Z by
    // ~5e-7 (emulated and consistent with scripts/tc_accuracy.py), which the final Newton-Schulz step
    // on Z squares away as long as the total stays near 1e-3 -- the regime the 2048^2 LAPACK
    // comparison covers (25 sweeps x 127 rounds).  Solves that run longer finish on the FFMA kernel.
    const bool tc_apply = r.tc_apply && (long long)r.sweeps_queued * r.nrounds < TC_APPLY_MAX_ROUNDS;
    for (int rd = 0; rd < r.nrounds; ++rd) {
        if (!(r.skip & 1) && r.tc_gram) {
            const long long items = (long long)r.njobs * ((lo.npairs + 1) / 2);
            cfg.gridDim = dim3((unsigned)(items < 148 ? items : 148));
            cfg.blockDim = dim3(TG_THREADS); cfg.dynamicSmemBytes = TG_SMEM;
            MPSB_CUDA(cudaLaunchKernelEx(&cfg, bj_gram_tc_kernel, p, rd, r.njobs));
            cfg.gridDim = dim3(lo.npairs, r.njobs, 1); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = gram_smem_bytes(2);
            MPSB_CUDA(cudaLaunchKernelEx(&cfg, bj_gram_evd_kernel<256, false>, p, rd, rd == 0 ? 1 : 0, 2));
        } else if (!(r.skip & 1)) {
            cfg.gridDim = dim3(lo.npairs, r.njobs, p.nsplit); cfg.dynamicSmemBytes = gram_smem_bytes(r.gram_stages);
            if (r.gram_threads == 512) {
                cfg.blockDim = dim3(512);
                MPSB_CUDA(cudaLaunchKernelEx(&cfg, bj_gram_evd_kernel<512, false>, p, rd, rd == 0 ? 1 : 0, r.gram_stages));
            } else if (((p.L + GK - 1) / GK) / p.nsplit >= 2) {      // at least two column tiles per CTA
                cfg.blockDim = dim3(256);
                MPSB_CUDA(cudaLaunchKernelEx(&cfg, bj_gram_evd_kernel<256, true>, p, rd, rd == 0 ? 1 : 0, r.gram_stages));
            } else {
                cfg.blockDim = dim3(256);
                MPSB_CUDA(cudaLaunchKernelEx(&cfg, bj_gram_evd_kernel<256, false>, p, rd, rd == 0 ? 1 : 0, r.gram_stages));
            }
        }
        if (!(r.skip & 4) && tc_apply && r.tma_apply) {
            const long long items = (long long)r.njobs * lo.npairs * (r.ntx + r.ntz);
            cfg.gridDim = dim3((unsigned)(items < 148 ? items : 148));
            cfg.blockDim = dim3(TB_THREADS); cfg.dynamicSmemBytes = TB_SMEM;
            MPSB_CUDA(cudaLaunchKernelEx(&cfg, bj_apply_tma_kernel, r.map_x, r.map_z, p, rd, r.ntx, r.ntx + r.ntz, r.njobs));
        } else if (!(r.skip & 4) && tc_apply) {
            const long long items = (long long)r.njobs * lo.npairs * (r.ntx + r.ntz);
            cfg.gridDim = dim3((unsigned)(items < 148 ? items : 148));
            cfg.blockDim = dim3(TA_THREADS); cfg.dynamicSmemBytes = TA_SMEM;
            MPSB_CUDA(cudaLaunchKernelEx(&cfg, bj_apply_tc_kernel, p, rd, r.ntx, r.ntx + r.ntz, r.njobs));
        } else if (!(r.skip & 4)) {
            cfg.gridDim = dim3((r.ntx + r.ntz + r.nt_cta - 1) / r.nt_cta, lo.npairs, r.njobs);
            cfg.blockDim = dim3(LT); cfg.dynamicSmemBytes = APPLY_SMEM;
            MPSB_CUDA(cudaLaunchKernelEx(&cfg, bj_apply_kernel, p, rd, r.ntx, r.ntx + r.ntz, r.nt_cta));
        }
    }
    cfg.gridDim = dim3((r.njobs + 127) / 128); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = 0;
    MPSB_CUDA(cudaLaunchKernelEx(&cfg, bj_sweep_end_kernel, p, r.njobs));
    MPSB_LAUNCH_CHECK("bj_round kernels");
    r.sweeps_queued += 1;
    if (r.sweeps_queued % CHECK_EVERY == 0 && r.sweeps_queued >= 4) {
        const int slot = (r.check_head + r.checks_out) % MAX_CHECKS;
        MPSB_CUDA(cudaMemcpyAsync(r.pin->host + (size_t)slot * r.pin->jobs_cap, p.misc, sizeof(Misc) * (size_t)r.njobs,
                                  cudaMemcpyDeviceToHost, st));
        MPSB_CUDA(cudaEventRecord(r.pin->ev[slot], st));
        r.checks_out += 1;
    }
    return 0;
}

// Advance a run as far as possible without blocking (block = true: wait for its oldest check).
// Returns 0 or an error; sets r.done once every job has converged or the sweep limit is queued.
static int large_advance(LargeRun& r, bool block) {
    while (!r.done) {
        if (r.checks_out == MAX_CHECKS || (r.checks_out > 0 && r.sweeps_queued >= r.max_outer)) {
            cudaEvent_t ev = r.pin->ev[r.check_head];
            if (block) MPSB_CUDA(cudaEventSynchronize(ev));
            else {
                cudaError_t q = cudaEventQuery(ev);
                if (q == cudaErrorNotReady) return 0;
                MPSB_CUDA(q);
            }
            const Misc* m = r.pin->host + (size_t)r.check_head * r.pin->jobs_cap;
            bool any = false;
            for (int j = 0; j < r.njobs; ++j) any = any || m[j].active != 0;
            r.check_head = (r.check_head + 1) % MAX_CHECKS;
            r.checks_out -= 1;
            if (!any) { r.done = true; break; }
            continue;
        }
        if (r.sweeps_queued >= r.max_outer) { r.done = true; break; }
        int rc = large_enqueue_sweep(r);
        if (rc) return rc;
    }
    return 0;
}

static int large_finish(LargeRun& r) {
    LargeParams& p = r.p;
    const LargeLayout& lo = r.lo;
    cudaStream_t st = r.st;
    const int njobs = r.njobs;
    // Every block rotation is a 32-term fp32 GEMM, so after several hundred of them Z has drifted
    // from unitarity by ~1e-5 (and X = Z M with it), which would show up 1:1 in the singular
    // values.  Re-orthonormalise Z with one Newton-Schulz step, Z <- (3/2 I - 1/2 Z Z^H) Z, and
    // recompute X = Z M from the saved input: the split stays an exact projection, the
    // accumulated error is gone, and the rows of X stay orthogonal to second order.
    int rc = launch_cgemm(p.Z, lo.nvp, 1, 0, (int64_t)lo.z, p.Z, 1, lo.nvp, 1, (int64_t)lo.z,
                          r.Rbuf, lo.nvp, (int64_t)lo.z, lo.nvp, lo.nvp, lo.nvp, njobs, st);
    if (rc) return rc;
    bj_ns_kernel<<<dim3(148, njobs), 256, 0, st>>>(r.Rbuf, (int64_t)lo.z, lo.nvp);
    rc = launch_cgemm(r.Rbuf, lo.nvp, 1, 0, (int64_t)lo.z, p.Z, lo.nvp, 1, 0, (int64_t)lo.z,
                      r.Z2, lo.nvp, (int64_t)lo.z, lo.nvp, lo.nvp, lo.nvp, njobs, st);
    if (rc) return rc;
    rc = launch_cgemm(r.Z2, lo.nvp, 1, 0, (int64_t)lo.z, r.M0, r.L, 1, 0, (int64_t)lo.m0,
                      p.X, r.L, p.x_stride, lo.nvp, r.L, r.nv, njobs, st);
    if (rc) return rc;
    p.Z = r.Z2;
    bj_sigma_kernel<<<dim3((lo.nvp + LT / 32 - 1) / (LT / 32), njobs), LT, 0, st>>>(p);
    bj_rank_kernel<<<dim3((lo.nvp + 127) / 128, njobs), 128, 0, st>>>(p);
    bj_write_kernel<<<dim3(32, njobs), LT, 0, st>>>(p, r.o);
    MPSB_LAUNCH_CHECK("bj_write_kernel");
    if (p.clk) {
        long long h[8];
        MPSB_CUDA(cudaMemcpy(h, p.clk, sizeof(h), cudaMemcpyDeviceToHost));
        fprintf(stderr, "bj_gram_evd phases (cycles, CTA 0 of the last launch): gram %lld  reduce+scale %lld  pass %lld  flags+NS %lld\n",
                h[1] - h[0], h[2] - h[1], h[3] - h[2], h[4] - h[3]);
    }
    return 0;
}

}  // namespace

int svd_large_padded_rows(int nv) { return (nv + P - 1) / P * P; }

size_t svd_large_workspace_elems(int nv, int L) { return large_layout(nv, L).total; }

int launch_svd_large(cf* X, int64_t x_job_stride, int njobs, int nv, int L, int k,
                     int left_canonical, const mpsb_gate2_desc* descs, int ndesc, int nbatch,
                     cf* left, int64_t left_stride, cf* right, int64_t right_stride,
                     float* svals, int64_t svals_stride, int32_t* info, cf* work,
                     cudaStream_t st) {
    (void)ndesc;
    if (njobs <= 0) return 0;
    std::lock_guard<std::recursive_mutex> lock(mpsb_lib_mutex());
    LargeRun r;
    int rc = large_begin(r, X, x_job_stride, njobs, nv, L, k, left_canonical, descs, nbatch, left, left_stride,
                         right, right_stride, svals, svals_stride, info, work, st, PIN_SLOTS - 1);
    if (rc) return rc;
    rc = large_advance(r, true);
    if (rc) return rc;
    return large_finish(r);
}

// Several independent solves (different shapes, disjoint outputs), each on its own stream:
// begin all, keep every stream fed by polling, finish each as soon as it has converged.

int launch_svd_large_multi(const LargeMultiJob* jobs, int n) {
    std::lock_guard<std::recursive_mutex> lock(mpsb_lib_mutex());
    std::vector<LargeRun> runs(n);
    std::vector<char> finished(n, 0);
    for (int i = 0; i < n; ++i) {
        const LargeMultiJob& j = jobs[i];
        int rc = large_begin(runs[i], j.X, j.x_job_stride, j.njobs, j.nv, j.L, j.k, j.left_canonical, j.descs, j.nbatch,
                             nullptr, 0, nullptr, 0, nullptr, 0, j.info, j.work, j.st, j.pin_slot);
        if (rc) return rc;
    }
    int left = n;
    unsigned spins = 0;
    while (left > 0) {
        for (int i = 0; i < n; ++i) {
            if (finished[i]) continue;
            int rc = large_advance(runs[i], left == 1);      // the last one may block instead of spinning
            if (rc) return rc;
            if (runs[i].done) {
                rc = large_finish(runs[i]);
                if (rc) return rc;
                finished[i] = 1;
                left -= 1;
            }
        }
        if ((++spins & 15u) == 0) sched_yield();
    }
    return 0;
}
