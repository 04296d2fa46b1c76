// Truncated SVD for d*chi <= 128: one CTA per matrix, everything resident in shared memory.
//
// Replaces tn.split_node_full_svd -> np.linalg.svd (LAPACK zgesdd) + slice [:k] + the two
// absorb contractions of mpsim/core.py:1132-1152.
//
// Input  X [nv][L] : the theta matrix oriented so that its ROWS are the vectors to
//                    orthogonalise (left-canonical: X = theta, nv = d*chiL, L = d*chiR;
//                    otherwise X = theta^T).  X = U S V^H below.
// Method (CPU model of the same algorithm: tests/_jacobi_model.py):
//   0. Y = X scaled by an exact power of two so that max|x| is in [1, 2).
//   1. Householder QR preconditioning, R only (Y <- R): Jacobi on R converges in ~8 sweeps
//      independently of how graded the spectrum is (plain Jacobi on X needs 10-25).
//   2. One-sided Jacobi on the rows of Y (Y <- J Y, J unitary, never formed).  Rows are visited by
//      a block tournament: 4-row blocks are paired by the circle method; a warp owns a pair of
//      blocks, keeps its 8 rows in registers (lane owns elements lane, lane+32, ...) and performs
//      the 16 cross rotations (plus the 12 intra-block ones in the first round of a sweep) with
//      transposed warp-shuffle reductions; squared row norms are cached in shared memory and
//      refreshed once per sweep.  At convergence row j of Y is sigma_j v_j^H.
//   3. sigma_j = |Y_j|, stable descending rank sort (ties keep the lower index: this is what
//      reproduces the reference on Bell + maxsvals=1, README.md:48-53), keep the first k.
//   4. Isometry without accumulating J and without dividing by small numbers that matter:
//        W = X V_k  (columns sigma_j u_j, rescaled by 1/sigma_j; columns with sigma_j ~ 0 are set to 0),
//        Q = the first k columns of the unitary of a Householder QR of W (formed in place).
//      Q is a product of reflectors applied to [I_k; 0], hence orthonormal to rounding even in
//      the null directions (like LAPACK's U), and because the columns of W come in descending
//      sigma order, span(Q[:, :j]) = span(u_1..u_j) for every j.
//   5. Split/absorb:  P = Q^H X  (k x L), so that  Q P  is the exact orthogonal projection of X
//      onto span(U_k) whether or not Jacobi converged:
//        left-canonical : left = Q        (U)     right = P      (S.Vh)
//        otherwise      : left = P^T      (U.S)   right = Q^T    (Vh; X was theta^T)
#include "common.cuh"
#include <stdlib.h>

namespace {

constexpr int ST = 512;          // threads per CTA
constexpr int NW = ST / 32;      // warps: one 8-row group each at nv = 128
constexpr int EPL = 4;           // elements per lane per row (row length <= 128)
constexpr int XCH = 32;          // staging chunk of X: columns in step 4, rows in step 5
constexpr int XS_ELEMS = 128 * (XCH + 1);   // >= XCH * (128 + 4)
constexpr float BIG2 = 1e-4f * 1e-4f;       // a sweep without a rotation above this is the last

// Profiling build only (MPSB_NVCC_EXTRA=-DMPSB_PROFILE, scripts/prof_svd.py): phase boundaries of
// CTA 0 (clock64), read back through mpsb_debug_phase_clocks.  The release library carries neither
// the marks nor the export.
#ifdef MPSB_PROFILE
__device__ long long g_phase_clk[32];
#define PHASE_MARK(i) do { if (blockIdx.x == 0 && threadIdx.x == 0) g_phase_clk[i] = clock64(); } while (0)
#else
#define PHASE_MARK(i) do { } while (0)
#endif

struct SvdSmallParams {
    const cf* X; int64_t x_stride;
    int nv, L, k, lc;
    int nvp, LS, LC;             // padded rows, smem row stride, columns touched by lanes
    const mpsb_gate2_desc* descs; int nbatch;     // output mode A (descs != nullptr)
    cf* left; int64_t left_stride; cf* right; int64_t right_stride;   // output mode B
    float* svals; int64_t svals_stride;
    int32_t* info;
    int max_sweeps; float tol2; int do_qr; int use_ns;
};

// (c, s, t|g|) of [[c, s], [-conj(s), c]] diagonalising [[a, g], [conj(g), b]]; s first, then
// c = sqrt(1 - |s|^2) derived from s (series near 1) so the rotation is unitary to rounding
// WITHOUT bias -- see tests/_jacobi_model.py:rotation_params.
__device__ __forceinline__ void rot_params(float a, float b, float gr, float gi, float g2,
                                           float& c, float& sr, float& si, float& tg) {
    // The tangent only steers convergence, so it is built from MUFU approximations (rsqrt, rcp:
    // ~2 ulp) instead of an IEEE division and square root -- every lane runs this code once per
    // four rotations.  Unitarity does not depend on it: c is derived from the s actually used.
    float rg = rsqrtf(g2);
    float zeta = (a - b) * (0.5f * rg);
    float az = fminf(fabsf(zeta), 1e18f);                    // keeps az^2 finite
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
        c = sqrtf(fmaf(-sr, sr, fmaf(-si, si, 1.0f)));       // large rotations (first sweeps only)
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

// One sub-round on the rows of a group: 4 disjoint pairs (A_i, B_i).  The four Gram entries are
// reduced together; lane l then computes the rotation of pair (l >> 3) only and the parameters
// are exchanged by shuffles (4x fewer scalar instructions than every lane doing all four).
template <int A0, int A1, int A2, int A3, int B0, int B1, int B2, int B3>
__device__ __forceinline__ int sub_round_y(cf (&y)[8][EPL], float (&a)[8], float tol2, int lane, bool& big) {
    constexpr int PA[4] = {A0, A1, A2, A3}, PB[4] = {B0, B1, B2, B3};
    float gr[4], gi[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float r = 0.f, m = 0.f;
#pragma unroll
        for (int t = 0; t < EPL; ++t) {
            cf p = y[PA[i]][t], q = y[PB[i]][t];
            r = fmaf(p.x, q.x, r); r = fmaf(p.y, q.y, r);
            m = fmaf(p.y, q.x, m); m = fmaf(-p.x, q.y, m);
        }
        gr[i] = r; gi[i] = m;
    }
    // Transposed reduction: 12 shuffles instead of 40.  After the 16- and 8-steps every lane
    // owns ONE of the four pairs (pair index = lane >> 3), the 4/2/1 butterfly finishes it.
    const bool h16 = (lane & 16) != 0, h8 = (lane & 8) != 0;
    float mgr, mgi;
    {
        float k0 = h16 ? gr[2] : gr[0], k1 = h16 ? gr[3] : gr[1];
        float s0 = h16 ? gr[0] : gr[2], s1 = h16 ? gr[1] : gr[3];
        k0 += __shfl_xor_sync(0xffffffffu, s0, 16);
        k1 += __shfl_xor_sync(0xffffffffu, s1, 16);
        float k = h8 ? k1 : k0, sd = h8 ? k0 : k1;
        k += __shfl_xor_sync(0xffffffffu, sd, 8);
        k += __shfl_xor_sync(0xffffffffu, k, 4);
        k += __shfl_xor_sync(0xffffffffu, k, 2);
        k += __shfl_xor_sync(0xffffffffu, k, 1);
        mgr = k;
    }
    {
        float k0 = h16 ? gi[2] : gi[0], k1 = h16 ? gi[3] : gi[1];
        float s0 = h16 ? gi[0] : gi[2], s1 = h16 ? gi[1] : gi[3];
        k0 += __shfl_xor_sync(0xffffffffu, s0, 16);
        k1 += __shfl_xor_sync(0xffffffffu, s1, 16);
        float k = h8 ? k1 : k0, sd = h8 ? k0 : k1;
        k += __shfl_xor_sync(0xffffffffu, sd, 8);
        k += __shfl_xor_sync(0xffffffffu, k, 4);
        k += __shfl_xor_sync(0xffffffffu, k, 2);
        k += __shfl_xor_sync(0xffffffffu, k, 1);
        mgi = k;
    }
    // this lane's pair: index lane >> 3 (all 8 lanes of an octet hold bitwise identical sums)
    const int sel = lane >> 3;
    float ap = a[PA[0]], aq = a[PB[0]];
#pragma unroll
    for (int i = 1; i < 4; ++i) {
        if (sel == i) { ap = a[PA[i]]; aq = a[PB[i]]; }
    }
    const float g2 = fmaf(mgr, mgr, mgi * mgi);
    const float apq = ap * aq;
    float c = 1.f, sr = 0.f, si = 0.f, tg = 0.f;
    const bool dorot = (g2 > tol2 * apq) && (g2 > 1e-30f);
    big = big || (dorot && g2 > BIG2 * apq);
    if (dorot) rot_params(ap, aq, mgr, mgi, g2, c, sr, si, tg);
    const unsigned bal = __ballot_sync(0xffffffffu, dorot);
    const unsigned flags = (bal & 1u) | ((bal >> 7) & 2u) | ((bal >> 14) & 4u) | ((bal >> 21) & 8u);
    if (flags == 0u) return 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        if (flags & (1u << i)) {                       // warp-uniform
            const float ci = __shfl_sync(0xffffffffu, c, 8 * i);
            const float sri = __shfl_sync(0xffffffffu, sr, 8 * i);
            const float sii = __shfl_sync(0xffffffffu, si, 8 * i);
            const float tgi = __shfl_sync(0xffffffffu, tg, 8 * i);
#pragma unroll
            for (int t = 0; t < EPL; ++t) rot_apply(ci, sri, sii, y[PA[i]][t], y[PB[i]][t]);
            a[PA[i]] = fmaxf(a[PA[i]] + tgi, 0.f);
            a[PB[i]] = fmaxf(a[PB[i]] - tgi, 0.f);
        }
    }
    return __popc(flags);
}

__device__ __forceinline__ void named_barrier_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// threads per column for the Householder passes: the largest power of two <= 32 with ncols * tpc <= ST
__device__ __forceinline__ int threads_per_col(int ncols) {
    int t = 32;
    while (t > 1 && ncols * t > ST) t >>= 1;
    return t;
}

// Householder reduction of A [nrows][ncols] (row stride LS), register resident: tpc threads share
// a column (thread (col, sub) keeps rows sub, sub + tpc, ... of its column in registers for the
// whole factorisation), so a step only moves one column through shared memory (the first
// implementation streamed the matrix through shared memory twice per step).
// Step j:  every warp reads the published column j (zero above and on the diagonal; x0 aside),
// reduces its norm and derives (v0, tau) redundantly -- no single warp is on the critical path and
// one barrier per step suffices; the columns > j are updated in registers over warp-uniform chunks
// of 8 register slots (the zeros of the published column make per-lane predicates unnecessary,
// only the chunk holding row j patches in v0); the owners of column j+1 publish it.
//   MODE 0 : A <- R (alpha on the diagonal, zeros below); the reflectors are dropped.
//   MODE 1 : A <- Q = H_0 H_1 ... H_{k-1} [I_k; 0], the first k = ncols columns of the unitary
//            (forward pass keeps v_j in column j, the backward pass forms Q in the same registers).
// H_j = I - tau_j v_j v_j^H.  Shared memory: vbuf [2][VB], scal [8], tau_arr / v0_arr [>= ncols].
constexpr int RMAX = 32;         // register slots per thread: nrows <= 128 with tpc >= 4
constexpr int RCH = 8;           // slots per chunk
constexpr int VB = 256;          // > 31 + 32 * 7: every slot of every chunk maps inside the buffer

#define HH_CHUNKS(C0, C1, BODY)                                        \
    _Pragma("unroll") for (int c_ = 0; c_ < RMAX / RCH; ++c_) {        \
        if (c_ >= (C0) && c_ < (C1)) {                                 \
            _Pragma("unroll") for (int u_ = 0; u_ < RCH; ++u_) {       \
                const int t = c_ * RCH + u_;                           \
                BODY                                                   \
            }                                                          \
        }                                                              \
    }
// chunk CD carries the diagonal row (BODY_D), the chunks after it are plain (BODY)
#define HH_CHUNKS_D(CD, C1, BODY_D, BODY)                              \
    _Pragma("unroll") for (int c_ = 0; c_ < RMAX / RCH; ++c_) {        \
        if (c_ == (CD)) {                                              \
            _Pragma("unroll") for (int u_ = 0; u_ < RCH; ++u_) {       \
                const int t = c_ * RCH + u_;                           \
                BODY_D                                                 \
            }                                                          \
        } else if (c_ > (CD) && c_ < (C1)) {                           \
            _Pragma("unroll") for (int u_ = 0; u_ < RCH; ++u_) {       \
                const int t = c_ * RCH + u_;                           \
                BODY                                                   \
            }                                                          \
        }                                                              \
    }

template <int MODE>
__device__ __noinline__ void householder(cf* A, int LS, int nrows, int ncols, int nsteps, cf* vbuf, float* scal,
                                         float* tau_arr, cf* v0_arr) {
    const int tid = threadIdx.x, lane = tid & 31;
    if (nsteps <= 0 && MODE == 0) return;
    const int tpc = threads_per_col(ncols);
    const int col = tid / tpc, sub = tid % tpc;
    const bool mine = col < ncols;
    const int c1 = ((nrows + tpc - 1) / tpc + RCH - 1) / RCH;        // chunks in use (warp-uniform)
    const unsigned gmask = tpc == 32 ? 0xffffffffu : (((1u << tpc) - 1u) << (lane & ~(tpc - 1)));
    cf a[RMAX];
#pragma unroll
    for (int t = 0; t < RMAX; ++t) {
        const int row = sub + tpc * t;
        a[t] = (mine && row < nrows) ? A[(size_t)row * LS + col] : cf_make(0.f, 0.f);
    }
    for (int i = tid; i < 2 * VB; i += ST) vbuf[i] = cf_make(0.f, 0.f);
    __syncthreads();

    // the tpc owners of column jn publish it: zeros down to and including row jn, x0 aside
    auto publish = [&](int jn, cf* vn, float* sc) {
        const int cj = jn / (RCH * tpc);
        HH_CHUNKS(cj > 0 ? cj - 1 : 0, c1, {
            const int row = sub + tpc * t;
            vn[tpc * t] = row > jn ? a[t] : cf_make(0.f, 0.f);
            if (row == jn) { sc[0] = a[t].x; sc[1] = a[t].y; }
        })
    };
    // w = v^H a over my rows, reduced over the column's threads; then a -= tau v w.
    // Each chunk's 8 reflector entries are loaded into registers first so that the shared-memory
    // latency is paid once per chunk, not once per row.
    auto reflect_column = [&](const cf* vb, int j, cf v0, float tau) {
        const int cd = j / (RCH * tpc);
        cf w0 = cf_make(0.f, 0.f), w1 = cf_make(0.f, 0.f);
#pragma unroll
        for (int c_ = 0; c_ < RMAX / RCH; ++c_) {
            if (c_ >= cd && c_ < c1) {
                cf vv[RCH];
#pragma unroll
                for (int u_ = 0; u_ < RCH; ++u_) vv[u_] = vb[tpc * (c_ * RCH + u_)];
                if (c_ == cd) {
#pragma unroll
                    for (int u_ = 0; u_ < RCH; ++u_) if (sub + tpc * (c_ * RCH + u_) == j) vv[u_] = v0;
                }
#pragma unroll
                for (int u_ = 0; u_ < RCH; ++u_) {
                    if (u_ & 1) w1 = cf_fma_conja(vv[u_], a[c_ * RCH + u_], w1);
                    else w0 = cf_fma_conja(vv[u_], a[c_ * RCH + u_], w0);
                }
            }
        }
        cf w = cf_add(w0, w1);
        for (int o = tpc >> 1; o > 0; o >>= 1) {
            w.x += __shfl_xor_sync(gmask, w.x, o);
            w.y += __shfl_xor_sync(gmask, w.y, o);
        }
        const cf tw = cf_scale(-tau, w);
#pragma unroll
        for (int c_ = 0; c_ < RMAX / RCH; ++c_) {
            if (c_ >= cd && c_ < c1) {
                cf vv[RCH];
#pragma unroll
                for (int u_ = 0; u_ < RCH; ++u_) vv[u_] = vb[tpc * (c_ * RCH + u_)];
                if (c_ == cd) {
#pragma unroll
                    for (int u_ = 0; u_ < RCH; ++u_) if (sub + tpc * (c_ * RCH + u_) == j) vv[u_] = v0;
                }
#pragma unroll
                for (int u_ = 0; u_ < RCH; ++u_) a[c_ * RCH + u_] = cf_fma(vv[u_], tw, a[c_ * RCH + u_]);
            }
        }
    };

    if (mine && col == 0 && nsteps > 0) publish(0, vbuf + sub, scal);
    __syncthreads();
    cf my_alpha = cf_make(0.f, 0.f), my_v0 = cf_make(0.f, 0.f);
    bool my_reflect = false;
    for (int j = 0; j < nsteps; ++j) {
        const int cur = j & 1, nxt = cur ^ 1;
        const cf* vbl = vbuf + cur * VB;
        float tail2 = 0.f;
#pragma unroll
        for (int m = 0; m < 4; ++m) tail2 += cf_abs2(vbl[lane + 32 * m]);
        tail2 = warp_sum(tail2);
        cf x0 = cf_make(scal[cur * 4 + 0], scal[cur * 4 + 1]);
        float ax0sq = cf_abs2(x0);
        // |x0|^2 below ~1e-30 is a denormal-range number with few significant bits: the
        // phase x0/|x0| would be off by 1e-4 and the reflector no longer unitary (seen on
        // GHZ circuits).  The matrix is scaled to max|x| in [1,2), so such an x0 is noise.
        if (ax0sq < 1e-30f) { x0 = cf_make(0.f, 0.f); ax0sq = 0.f; }
        // skip the reflector when the column is already reduced (keeps diagonal inputs, and with
        // them the order of tied singular values, untouched), or so small that 1/|x|^2 overflows
        const bool reflect = tail2 > 0.f && tail2 + ax0sq > 1e-30f;
        cf v0 = cf_make(0.f, 0.f), alpha = cf_make(0.f, 0.f);
        float tau = 0.f;
        if (reflect) {
            float ax0 = sqrtf(ax0sq);
            float normx = sqrtf(tail2 + ax0sq);
            cf ph = ax0 > 0.f ? cf_scale(1.0f / ax0, x0) : cf_make(1.f, 0.f);
            alpha = cf_scale(-normx, ph);
            v0 = cf_sub(x0, alpha);
            tau = 1.0f / (normx * (normx + ax0));
        }
        if (MODE == 1 && tid == 0) { tau_arr[j] = tau; v0_arr[j] = v0; }
        if (mine && col == j) { my_alpha = alpha; my_v0 = v0; my_reflect = reflect; }
        if (mine && col > j) {
            if (reflect) reflect_column(vbl + sub, j, v0, tau);
            if (col == j + 1 && j + 1 < nsteps) publish(j + 1, vbuf + nxt * VB + sub, scal + nxt * 4);
        }
        __syncthreads();
    }
    if (MODE == 1) {
        // backward: Q <- H_j Q for j = k-1 .. 0, in the same registers
        const int k = ncols;
        if (mine && col == k - 1) publish(k - 1, vbuf + ((k - 1) & 1) * VB + sub, scal);
        __syncthreads();
        for (int j = k - 1; j >= 0; --j) {
            const float tau = tau_arr[j];
            const cf v0 = v0_arr[j];
            if (mine && col > j) {
                if (tau != 0.f) reflect_column(vbuf + (j & 1) * VB + sub, j, v0, tau);
            } else if (mine && col == j) {
                // H_j e_j = e_j - tau conj(v_j[0]) v_j ; rows above j still hold R from the forward pass
                const cf f = (tau != 0.f) ? cf_scale(-tau, cf_conj(v0)) : cf_make(0.f, 0.f);
                HH_CHUNKS(0, c1, {
                    const int row = sub + tpc * t;
                    cf q = cf_make(0.f, 0.f);
                    if (row >= j && row < nrows) {
                        q = cf_mul(f, row == j ? v0 : a[t]);
                        if (row == j) q.x += 1.0f;
                    }
                    a[t] = q;
                })
            } else if (mine && j > 0 && col == j - 1) {
                publish(j - 1, vbuf + ((j - 1) & 1) * VB + sub, scal);
            }
            __syncthreads();
        }
    }
#pragma unroll
    for (int t = 0; t < RMAX; ++t) {
        const int row = sub + tpc * t;
        if (mine && row < nrows) {
            cf v = a[t];
            if (MODE == 0 && my_reflect && row >= col) v = (row == col) ? my_alpha : cf_make(0.f, 0.f);
            A[(size_t)row * LS + col] = v;
        }
    }
    __syncthreads();
}

// Orthonormalisation of W [nv][k] (k <= 64) by one Newton-Schulz step, Q = W (3/2 I - 1/2 W^H W):
// two small GEMMs instead of the 2k barrier-separated Householder steps.  W = X V_k S^-1 is
// orthonormal up to E = W^H W - I with |E_ij| ~ 3e-6 sigma_i/sigma_j; the step squares that
// (||Q^H Q - I|| <= 3/4 ||E||^2) and keeps span(Q) = span(W).  Only taken when ||E||_F < 2e-3
// (measured on the data, block-uniform); otherwise W is left untouched and the caller runs the
// Householder path, which also handles zero / noise columns.  M: scratch of 64 x 64.
__device__ bool newton_schulz_q(cf* W, int LS, int nv, int k, cf* M, float* red) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    {
        const int j = tid >> 3, jb = (tid & 7) * 8;            // G[j][jb .. jb+7]
        cf acc[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) acc[u] = cf_make(0.f, 0.f);
        if (j < k && jb < k) {
            for (int a = 0; a < nv; ++a) {
                const cf* wr = W + (size_t)a * LS;
                const cf wj = wr[j];
                const float4* w4 = reinterpret_cast<const float4*>(wr + jb);      // rows are 16-byte aligned (LS even)
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const float4 v = w4[u];
                    acc[2 * u] = cf_fma_conja(wj, cf_make(v.x, v.y), acc[2 * u]);
                    acc[2 * u + 1] = cf_fma_conja(wj, cf_make(v.z, v.w), acc[2 * u + 1]);
                }
            }
        }
        float e2 = 0.f;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const bool ok = j < k && jb + u < k;
            cf e = acc[u];
            if (j == jb + u) e.x -= 1.0f;
            if (ok) e2 += cf_abs2(e);
            // M = 3/2 I - 1/2 G  (zero outside the k x k block)
            M[j * 64 + jb + u] = ok ? cf_make((j == jb + u ? 1.5f : 0.f) - 0.5f * acc[u].x, -0.5f * acc[u].y) : cf_make(0.f, 0.f);
        }
        e2 = warp_sum(e2);
        if (lane == 0) red[warp] = e2;
    }
    __syncthreads();
    float e2 = 0.f;
#pragma unroll
    for (int w = 0; w < NW; ++w) e2 += red[w];
    if (!(e2 < 2e-3f * 2e-3f)) { __syncthreads(); return false; }       // also catches NaN
    {
        const int a = tid >> 2, jq = (tid & 3) * 16;           // Q[a][jq .. jq+15]
        cf acc[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) acc[u] = cf_make(0.f, 0.f);
        if (a < nv && jq < k) {
            const cf* wr = W + (size_t)a * LS;
            for (int jp = 0; jp < k; ++jp) {
                const cf w = wr[jp];
                const float4* m4 = reinterpret_cast<const float4*>(M + jp * 64 + jq);
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const float4 v = m4[u];
                    acc[2 * u] = cf_fma(w, cf_make(v.x, v.y), acc[2 * u]);
                    acc[2 * u + 1] = cf_fma(w, cf_make(v.z, v.w), acc[2 * u + 1]);
                }
            }
        }
        __syncthreads();                                       // every read of W is done
        if (a < nv) {
#pragma unroll
            for (int u = 0; u < 16; ++u)
                if (jq + u < k) W[(size_t)a * LS + jq + u] = acc[u];
        }
    }
    __syncthreads();
    return true;
}

__global__ void __launch_bounds__(ST, 1) svd_small_kernel(SvdSmallParams P) {
    extern __shared__ float4 smem_raw[];
    const int job = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nv = P.nv, L = P.L, nvp = P.nvp, LS = P.LS, k = P.k;

    cf* Ys = (cf*)smem_raw;                        // [nvp][LS]   Y, later W / Q
    cf* Xs = Ys + (size_t)nvp * LS;                // [XS_ELEMS]  staging of X for steps 4, 5
    cf* vbuf = Xs + XS_ELEMS;                      // [2][VB]
    cf* v0_arr = vbuf + 2 * VB;                    // [nvp]
    float* sig = (float*)(v0_arr + nvp);           // [nvp]  sigma (scaled units)
    float* nrm = sig + nvp;                        // [nvp]  cached squared row norms
    float* tau_arr = nrm + nvp;                    // [nvp]
    int* perm = (int*)(tau_arr + nvp);             // [nvp]
    float* scal = (float*)(perm + nvp);            // [NW]  (also the QR scalars: 8 used)

    PHASE_MARK(0);
    // ---- step 0: load X scaled by an exact power of two so that max|x| is in [1, 2) ----------
    // (the guards below are absolute, and products of numerical zeros would otherwise
    //  underflow in the Gram entries; LAPACK scales for the same reason)
    const cf* X = P.X + (size_t)job * P.x_stride;
    float mx = 0.f;
    for (int e = tid; e < nv * L; e += ST) { cf v = X[e]; mx = fmaxf(mx, fmaxf(fabsf(v.x), fabsf(v.y))); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (lane == 0) scal[warp] = mx;
    __syncthreads();
    mx = 0.f;
#pragma unroll
    for (int w = 0; w < NW; ++w) mx = fmaxf(mx, scal[w]);
    int ex = 0;
    if (mx > 0.f && isfinite(mx)) (void)frexpf(mx, &ex);      // mx = f * 2^ex, f in [0.5, 1)
    const float scale_in = mx > 0.f ? ldexpf(1.0f, 1 - ex) : 1.0f;
    const float scale_out = mx > 0.f ? ldexpf(1.0f, ex - 1) : 1.0f;
    __syncthreads();
    for (int e = tid; e < nvp * LS; e += ST) {
        int i = e / LS, c = e - i * LS;
        cf v = cf_make(0.f, 0.f);
        if (i < nv && c < L) { v = X[(size_t)i * L + c]; v.x *= scale_in; v.y *= scale_in; }
        Ys[e] = v;
    }
    __syncthreads();

    PHASE_MARK(1);
    // ---- step 1: Householder QR, R only ------------------------------------------------------
    if (P.do_qr) householder<0>(Ys, LS, nv, L, min(nv - 1, L), vbuf, scal, nullptr, nullptr);

    PHASE_MARK(2);
    // ---- step 2: one-sided Jacobi on the rows of Y -------------------------------------------
    const int nact = P.do_qr ? min(nv, L) : nv;    // rows >= L of R are exactly zero
    const int nb = 2 * ((nact + 7) / 8);           // 4-row blocks (even count)
    const int mcirc = nb - 1;
    const int ngroups = nb / 2;
    const int nrounds = nb > 2 ? nb - 1 : 1;
    const int ylanes = P.LC / 32;
    int sweeps = 0, status = 0;
    if (nact >= 2) {
        status = 1;
        for (int sweep = 0; sweep < P.max_sweeps; ++sweep) {
            // refresh the cached squared norms (they are updated by +-t|g| within the sweep)
            for (int i = warp; i < 4 * nb; i += NW) {
                const cf* yr = Ys + (size_t)i * LS;
                float s2 = 0.f;
#pragma unroll
                for (int t = 0; t < EPL; ++t)
                    if (t < ylanes) s2 += cf_abs2(yr[lane + 32 * t]);
                s2 = warp_sum(s2);
                if (lane == 0) nrm[i] = s2;
            }
            __syncthreads();
            bool big = false;
            // Rounds are separated by PAIRWISE named barriers, not by a block-wide one: in the circle
            // method group g takes its two blocks of the next round from groups g-1 and g+1 only, so
            // a warp syncs with its two neighbours (edge (g, g+1) = barrier 1 + g, even warps right
            // edge first, odd warps left edge first) and the warps drift apart by whole rounds; their
            // shuffle / MUFU latencies then overlap instead of all of them stalling in the same phase.
            // (Polling per-block round counters in shared memory was tried and was 3x slower.)
            for (int r = 0; r < nrounds; ++r) {
                for (int g = warp; g < ngroups; g += NW) {
                    int I, Jb;
                    if (nb == 2) { I = 0; Jb = 1; }
                    else if (g == 0) { I = mcirc; Jb = r; }
                    else { I = (r + g) % mcirc; Jb = (r - g + mcirc) % mcirc; }
                    cf* rowA = Ys + (size_t)(4 * I) * LS + lane;
                    cf* rowB = Ys + (size_t)(4 * Jb) * LS + lane;
                    cf v[8][EPL];
                    float a[8];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
#pragma unroll
                        for (int t = 0; t < EPL; ++t) {
                            v[i][t] = (t < ylanes) ? rowA[(size_t)i * LS + 32 * t] : cf_make(0.f, 0.f);
                            v[4 + i][t] = (t < ylanes) ? rowB[(size_t)i * LS + 32 * t] : cf_make(0.f, 0.f);
                        }
                        a[i] = nrm[4 * I + i];
                        a[4 + i] = nrm[4 * Jb + i];
                    }
                    int nrot = 0;
                    if (r == 0) {
                        nrot += sub_round_y<0, 2, 4, 6, 1, 3, 5, 7>(v, a, P.tol2, lane, big);
                        nrot += sub_round_y<0, 1, 4, 5, 2, 3, 6, 7>(v, a, P.tol2, lane, big);
                        nrot += sub_round_y<0, 1, 4, 5, 3, 2, 7, 6>(v, a, P.tol2, lane, big);
                    }
                    nrot += sub_round_y<0, 1, 2, 3, 4, 5, 6, 7>(v, a, P.tol2, lane, big);
                    nrot += sub_round_y<0, 1, 2, 3, 5, 6, 7, 4>(v, a, P.tol2, lane, big);
                    nrot += sub_round_y<0, 1, 2, 3, 6, 7, 4, 5>(v, a, P.tol2, lane, big);
                    nrot += sub_round_y<0, 1, 2, 3, 7, 4, 5, 6>(v, a, P.tol2, lane, big);
                    if (nrot) {
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
#pragma unroll
                            for (int t = 0; t < EPL; ++t) {
                                if (t < ylanes) {
                                    rowA[(size_t)i * LS + 32 * t] = v[i][t];
                                    rowB[(size_t)i * LS + 32 * t] = v[4 + i][t];
                                }
                            }
                        }
                        float am = a[0];
#pragma unroll
                        for (int i = 1; i < 8; ++i) if (lane == i) am = a[i];
                        if (lane < 8) nrm[lane < 4 ? 4 * I + lane : 4 * Jb + lane - 4] = am;
                    }
                    if (r + 1 < nrounds) {
                        const bool has_r = g + 1 < ngroups, has_l = g > 0;
                        if (g & 1) {
                            if (has_l) named_barrier_sync(g, 64);
                            if (has_r) named_barrier_sync(g + 1, 64);
                        } else {
                            if (has_r) named_barrier_sync(g + 1, 64);
                            if (has_l) named_barrier_sync(g, 64);
                        }
                    }
                }
            }
            sweeps = sweep + 1;
            if (!__syncthreads_or(big ? 1 : 0)) { status = 0; break; }
        }
    }

    PHASE_MARK(3);
    // ---- step 3: singular values, stable descending sort --------------------------------------
    __syncthreads();
    for (int i = warp; i < nvp; i += NW) {
        float s2 = 0.f;
        if (i < nv) {
            const cf* yr = Ys + (size_t)i * LS;
            for (int c = lane; c < L; c += 32) s2 += cf_abs2(yr[c]);
        }
        s2 = warp_sum(s2);
        if (lane == 0) { sig[i] = sqrtf(s2); perm[i] = i; }
    }
    __syncthreads();
    int myrank = -1;
    if (tid < nv) {
        float si = sig[tid];
        int rank = 0;
        for (int j = 0; j < nv; ++j) {
            float sj = sig[j];
            rank += (sj > si) || (sj == si && j < tid);
        }
        myrank = rank;
    }
    __syncthreads();
    if (myrank >= 0) perm[myrank] = tid;
    __syncthreads();

    cf *left, *right; float* sv;
    if (P.descs) {
        int di = job / P.nbatch, bi = job % P.nbatch;
        const mpsb_gate2_desc d = P.descs[di];
        left = (cf*)d.out_l + (size_t)bi * d.bs_out_l;
        right = (cf*)d.out_r + (size_t)bi * d.bs_out_r;
        sv = d.svals ? d.svals + (size_t)bi * d.bs_svals : nullptr;
    } else {
        left = P.left + (size_t)job * P.left_stride;
        right = P.right + (size_t)job * P.right_stride;
        sv = P.svals ? P.svals + (size_t)job * P.svals_stride : nullptr;
    }
    if (sv) for (int j = tid; j < min(nv, L); j += ST) sv[j] = scale_out * sig[perm[j]];
    if (P.info && tid == 0) { P.info[2 * job] = status; P.info[2 * job + 1] = sweeps; }
    if (k <= 0) return;

    PHASE_MARK(4);
    // ---- step 4: W = (s X) Y_k^H diag(1/sigma'^2)  (columns ~ u_j), then Q from its QR --------
    // thread tile: rows a = lane + 32 i, columns j = warp + 16 jj
    {
        cf acc[4][8];
        int pj[8];
        float rinv[8];
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
            const int j = warp + NW * jj;
            pj[jj] = j < k ? perm[j] : 0;
            const float s = j < k ? sig[pj[jj]] : 0.f;
            // sigma' below 1e-12 (max|x| is in [1,2)) is noise of noise: leave the column to the
            // QR, which completes the basis with an arbitrary orthonormal vector (as LAPACK does)
            rinv[jj] = s > 1e-12f ? 1.0f / s : 0.f;
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[i][jj] = cf_make(0.f, 0.f);
        }
        const int njj = warp < k ? (k - warp + NW - 1) / NW : 0;
        for (int c0 = 0; c0 < L; c0 += XCH) {
            const int cw = min(XCH, L - c0);
            for (int e = tid; e < 128 * XCH; e += ST) {
                const int a_ = e / XCH, cc = e - a_ * XCH;
                cf v = cf_make(0.f, 0.f);
                if (a_ < nv && cc < cw) { v = X[(size_t)a_ * L + c0 + cc]; v.x *= scale_in; v.y *= scale_in; }
                Xs[a_ * (XCH + 1) + cc] = v;
            }
            __syncthreads();
            for (int cc = 0; cc < cw; ++cc) {
                cf xa[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) xa[i] = Xs[(lane + 32 * i) * (XCH + 1) + cc];
#pragma unroll
                for (int jj = 0; jj < 8; ++jj) {
                    if (jj < njj) {
                        const cf yb = Ys[(size_t)pj[jj] * LS + c0 + cc];
#pragma unroll
                        for (int i = 0; i < 4; ++i) acc[i][jj] = cf_fma_conja(yb, xa[i], acc[i][jj]);   // x conj(y)
                    }
                }
            }
            __syncthreads();
        }
        // all reads of Y are done (barrier above): overwrite it with W [nv][k]
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
            const int j = warp + NW * jj;
            if (j < k) {
                const float r = rinv[jj];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int a_ = lane + 32 * i;
                    if (a_ < nv) {
                        cf w = acc[i][jj];
                        w.x = (w.x * r) * r; w.y = (w.y * r) * r;
                        Ys[(size_t)a_ * LS + j] = w;
                    }
                }
            }
        }
        __syncthreads();
    }
    PHASE_MARK(5);
    {
        bool done = false;
        if (k <= 64 && P.use_ns) done = newton_schulz_q(Ys, LS, nv, k, Xs, scal);
        if (!done) householder<1>(Ys, LS, nv, k, k, vbuf, scal, tau_arr, v0_arr);
    }
    PHASE_MARK(6);

    // ---- step 5: P = Q^H X (k x L) and the outputs ---------------------------------------------
    // thread tile: columns c = lane + 32 ci, rows j = warp + 16 jj
    {
        cf acc[8][4];
#pragma unroll
        for (int jj = 0; jj < 8; ++jj)
#pragma unroll
            for (int ci = 0; ci < 4; ++ci) acc[jj][ci] = cf_make(0.f, 0.f);
        const int njj = warp < k ? (k - warp + NW - 1) / NW : 0;
        const int XLS = P.LC + 4;
        for (int a0 = 0; a0 < nv; a0 += XCH) {
            const int ah = min(XCH, nv - a0);
            for (int e = tid; e < ah * P.LC; e += ST) {
                const int al = e / P.LC, c = e - al * P.LC;
                Xs[al * XLS + c] = c < L ? X[(size_t)(a0 + al) * L + c] : cf_make(0.f, 0.f);
            }
            __syncthreads();
            for (int al = 0; al < ah; ++al) {
                cf xb[4];
#pragma unroll
                for (int ci = 0; ci < 4; ++ci) xb[ci] = (ci < ylanes) ? Xs[al * XLS + lane + 32 * ci] : cf_make(0.f, 0.f);
                const cf* qrow = Ys + (size_t)(a0 + al) * LS + warp;
#pragma unroll
                for (int jj = 0; jj < 8; ++jj) {
                    if (jj < njj) {
                        const cf q = qrow[NW * jj];
#pragma unroll
                        for (int ci = 0; ci < 4; ++ci) acc[jj][ci] = cf_fma_conja(q, xb[ci], acc[jj][ci]);
                    }
                }
            }
            __syncthreads();
        }
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
            const int j = warp + NW * jj;
            if (j < k) {
#pragma unroll
                for (int ci = 0; ci < 4; ++ci) {
                    const int c = lane + 32 * ci;
                    if (c < L) {
                        if (P.lc) right[(size_t)j * L + c] = acc[jj][ci];      // S.Vh  [k][L]
                        else left[(size_t)c * k + j] = acc[jj][ci];           // U.S   [L][k]
                    }
                }
            }
        }
    }
    PHASE_MARK(7);
    if (P.lc) {
        // left [nv][k] = Q
        for (int e = tid; e < nv * k; e += ST) { int a_ = e / k, j = e - a_ * k; left[e] = Ys[(size_t)a_ * LS + j]; }
    } else {
        // right [k][nv] = Q^T
        for (int e = tid; e < k * nv; e += ST) { int j = e / nv, b_ = e - j * nv; right[e] = Ys[(size_t)b_ * LS + j]; }
    }
    PHASE_MARK(8);
}

struct Layout { int nvp, LC, LS; size_t smem; };

Layout make_layout(int nv, int L) {
    Layout lo;
    lo.nvp = (nv + 7) / 8 * 8;
    lo.LC = (L + 31) / 32 * 32;
    lo.LS = lo.LC + 4;           // rows 16-byte aligned; 4 consecutive rows x 4 columns hit 16 distinct 8-byte banks
    lo.smem = ((size_t)lo.nvp * lo.LS + XS_ELEMS + 2 * (size_t)VB + lo.nvp) * 8 + (size_t)lo.nvp * 16 + NW * 4 + 64;
    return lo;
}

}  // namespace

#ifdef MPSB_PROFILE
extern "C" int mpsb_debug_phase_clocks(long long* out32) {
    return (int)cudaMemcpyFromSymbol(out32, g_phase_clk, sizeof(long long) * 32);
}
#endif

int launch_svd_small(const cf* X, int64_t x_job_stride, int njobs, int nv, int L, int k,
                     int left_canonical, const mpsb_gate2_desc* descs, int ndesc, int nbatch,
                     cf* left, int64_t left_stride, cf* right, int64_t right_stride,
                     float* svals, int64_t svals_stride, int32_t* info, cf* zglobal,
                     cudaStream_t st) {
    (void)ndesc; (void)zglobal;
    if (njobs <= 0) return 0;
    MPSB_ARG(nv >= 1 && L >= 1 && nv <= MPSB_MAX_SMALL_DIM && L <= MPSB_MAX_SMALL_DIM,
             "svd_small: shape %d x %d outside [1, %d]", nv, L, MPSB_MAX_SMALL_DIM);
    MPSB_ARG(k >= 0 && k <= (nv < L ? nv : L), "svd_small: k=%d out of range for %d x %d", k, nv, L);
    Layout lo = make_layout(nv, L);
    SvdSmallParams P;
    P.X = X; P.x_stride = x_job_stride;
    P.nv = nv; P.L = L; P.k = k; P.lc = left_canonical;
    P.nvp = lo.nvp; P.LS = lo.LS; P.LC = lo.LC;
    P.descs = descs; P.nbatch = nbatch > 0 ? nbatch : 1;
    P.left = left; P.left_stride = left_stride; P.right = right; P.right_stride = right_stride;
    P.svals = svals; P.svals_stride = svals_stride;
    P.info = info;
    P.max_sweeps = 30;
    P.tol2 = 3e-6f * 3e-6f;
    P.do_qr = 1;
    P.use_ns = 1;
    // debugging knobs (not part of the ABI)
    if (const char* e = mpsb_env("MPSB_SVD_MAX_SWEEPS")) P.max_sweeps = atoi(e);
    if (const char* e = mpsb_env("MPSB_SVD_NO_QR")) P.do_qr = atoi(e) ? 0 : 1;
    if (const char* e = mpsb_env("MPSB_SVD_NO_NS")) P.use_ns = atoi(e) ? 0 : 1;
    MPSB_CUDA(cudaFuncSetAttribute(svd_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lo.smem));
    svd_small_kernel<<<njobs, ST, lo.smem, st>>>(P);
    MPSB_LAUNCH_CHECK("svd_small_kernel");
    return 0;
}
