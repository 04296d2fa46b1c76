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
constexpr int XCH = 32;          // staging chunk of X: columns in step 4, rows in step 5
constexpr int XS_ELEMS = 128 * (XCH + 1);   // >= XCH * (128 + 4)
constexpr float BIG2 = 1e-3f * 1e-3f;       // a sweep without a rotation above this cosine is the last (see jacobi_sweeps)

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
    int max_sweeps; float tol2; float big2; int do_qr; int use_ns; int colsort;
};

// MUFU approximations without the denormal / range wrappers of rsqrtf() and __fdividef(): the
// arguments below are bounded away from 0 and infinity (the matrix is scaled to max|x| in [1, 2)
// and a rotation needs |g|^2 > 1e-30)
__device__ __forceinline__ float rsqrt_approx(float x) {
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// (c, s, t|g|) of [[c, s], [-conj(s), c]] diagonalising [[a, g], [conj(g), b]].
// With d = a - b, h = sqrt(d^2 + 4|g|^2), w = h + |d|:  |s|^2 = 2|g|^2 / (h w),  c^2 = w / (2h),
// t|g| = sign(d) 2|g|^2 / w  (the same rotation as t = sign(zeta) / (|zeta| + sqrt(1 + zeta^2)),
// zeta = d / 2|g|).  s is built from TWO dependent MUFU approximations (rsqrt of h^2, rsqrt of h w;
// ~2 ulp each) -- this runs on the critical path of every rotation set and the first version's four
// dependent MUFU + IEEE sqrt cost ~450 cycles per set (scripts/micro/jacobi_variants.cu) -- and
// c = sqrt(1 - |s|^2) is then DERIVED from the s actually used, so the rotation is unitary to
// rounding WITHOUT bias whatever the error of the angle (tests/_jacobi_model.py:rotation_params):
// series near 1, one Newton step on the approximation otherwise (|s|^2 <= 1/2).
__device__ __forceinline__ void rot_params(float a, float b, float gr, float gi, float g2,
                                           float& c, float& sr, float& si, float& tg) {
    const float d = a - b;
    const float u = fmaf(d, d, 4.0f * g2);
    const float h = u * rsqrt_approx(u);
    const float w = h + fabsf(d);
    const float k = copysignf(1.41421356f, d) * rsqrt_approx(h * w);
    sr = gr * k;
    si = gi * k;
    tg = copysignf(2.0f * g2 * rcp_approx(w), d);
    const float hh = fmaf(sr, sr, si * si);
    if (hh < 0.0625f) {
        const float poly = fmaf(hh, fmaf(hh, fmaf(hh, fmaf(hh, 0.02734375f, 0.0390625f), 0.0625f), 0.125f), 0.5f);
        c = fmaf(-hh, poly, 1.0f);
    } else {
        const float y = fmaf(-sr, sr, fmaf(-si, si, 1.0f));     // in [~1/2, 15/16]
        const float r0 = rsqrt_approx(y);
        const float c0 = y * r0;
        c = fmaf(0.5f * r0, fmaf(-c0, c0, y), c0);               // Newton: c0 + (y - c0^2) / (2 c0)
    }
}

// ---- sweep engine: rows in PLANAR layout, packed fma.rn.f32x2 ------------------------------------
// Row i of Y lives at (float*)(Ys + i * LS): W real parts, then W imaginary parts (W = LC = 64 or
// 128).  Lane l holds elements 2 NP l .. 2 NP l + 2 NP - 1 of a row (NP = W / 64 float2 pairs of
// real and of imaginary parts): one LDS.128 (NP = 2) or LDS.64 per plane, and every FFMA2 works on
// two elements (half the issue slots of the scalar FFMA version for the same FMA-pipe time).
template <int NP> struct PRow { float2 re[NP], im[NP]; };

template <int NP>
__device__ __forceinline__ void prow_load(PRow<NP>& r, const float* p, int lane) {
    if (NP == 2) {
        const float4 x = *reinterpret_cast<const float4*>(p + 4 * lane);
        const float4 y = *reinterpret_cast<const float4*>(p + 128 + 4 * lane);
        r.re[0] = make_float2(x.x, x.y); r.re[NP - 1] = make_float2(x.z, x.w);
        r.im[0] = make_float2(y.x, y.y); r.im[NP - 1] = make_float2(y.z, y.w);
    } else {
        r.re[0] = *reinterpret_cast<const float2*>(p + 2 * lane);
        r.im[0] = *reinterpret_cast<const float2*>(p + 64 + 2 * lane);
    }
}
template <int NP>
__device__ __forceinline__ void prow_store(const PRow<NP>& r, float* p, int lane) {
    if (NP == 2) {
        *reinterpret_cast<float4*>(p + 4 * lane) = make_float4(r.re[0].x, r.re[0].y, r.re[NP - 1].x, r.re[NP - 1].y);
        *reinterpret_cast<float4*>(p + 128 + 4 * lane) = make_float4(r.im[0].x, r.im[0].y, r.im[NP - 1].x, r.im[NP - 1].y);
    } else {
        *reinterpret_cast<float2*>(p + 2 * lane) = r.re[0];
        *reinterpret_cast<float2*>(p + 64 + 2 * lane) = r.im[0];
    }
}

__device__ __forceinline__ float2 neg2(float2 v) { return make_float2(-v.x, -v.y); }

// this lane's part of <p, q> = sum p conj(q)
template <int NP>
__device__ __forceinline__ void gram_part(const PRow<NP>& p, const PRow<NP>& q, float& gr, float& gi) {
    float2 r2 = make_float2(0.f, 0.f), m2 = r2;
#pragma unroll
    for (int t = 0; t < NP; ++t) {
        r2 = __ffma2_rn(p.re[t], q.re[t], r2);
        r2 = __ffma2_rn(p.im[t], q.im[t], r2);
        m2 = __ffma2_rn(p.im[t], q.re[t], m2);
        m2 = __ffma2_rn(neg2(p.re[t]), q.im[t], m2);       // the negation folds into the FFMA2 operand
    }
    gr = r2.x + r2.y;
    gi = m2.x + m2.y;
}

// [p; q] <- [[c, s], [-conj(s), c]] [p; q]; the same nesting of roundings as the scalar form
//   re p' = fma(c, p.re, fma(s.re, q.re, -(s.im q.im)))   etc.
template <int NP>
__device__ __forceinline__ void rot_apply(float c, float sr, float si, PRow<NP>& p, PRow<NP>& q) {
    const float2 C2 = make_float2(c, c), S2 = make_float2(sr, sr), I2 = make_float2(si, si);
    const float2 NS2 = make_float2(-sr, -sr), NI2 = make_float2(-si, -si);
#pragma unroll
    for (int t = 0; t < NP; ++t) {
        const float2 pre = p.re[t], pim = p.im[t], qre = q.re[t], qim = q.im[t];
        p.re[t] = __ffma2_rn(C2, pre, __ffma2_rn(S2, qre, __fmul2_rn(NI2, qim)));
        p.im[t] = __ffma2_rn(C2, pim, __ffma2_rn(S2, qim, __fmul2_rn(I2, qre)));
        q.re[t] = __ffma2_rn(C2, qre, __ffma2_rn(NS2, pre, __fmul2_rn(NI2, pim)));
        q.im[t] = __ffma2_rn(C2, qim, __ffma2_rn(I2, pre, __fmul2_rn(NS2, pim)));
    }
}

// transposed reduction of 4 values: afterwards every lane of octet i holds the full sum of g[i]
__device__ __forceinline__ float reduce4(const float (&g)[4], int lane) {
    const bool h16 = (lane & 16) != 0, h8 = (lane & 8) != 0;
    float k0 = h16 ? g[2] : g[0], k1 = h16 ? g[3] : g[1];
    const float s0 = h16 ? g[0] : g[2], s1 = h16 ? g[1] : g[3];
    k0 += __shfl_xor_sync(0xffffffffu, s0, 16);
    k1 += __shfl_xor_sync(0xffffffffu, s1, 16);
    float k = h8 ? k1 : k0;
    const float sd = h8 ? k0 : k1;
    k += __shfl_xor_sync(0xffffffffu, sd, 8);
    k += __shfl_xor_sync(0xffffffffu, k, 4);
    k += __shfl_xor_sync(0xffffffffu, k, 2);
    k += __shfl_xor_sync(0xffffffffu, k, 1);
    return k;
}

// One sub-round on the 8 rows of a group: 4 disjoint pairs (A_i, B_i).  The four Gram entries are
// reduced together (transposed: 12 shuffles instead of 40); the lanes of octet i compute the rotation
// of pair i and the parameters are exchanged by shuffles.
// Measured and NOT adopted (scripts/micro/jacobi_variants.cu V5; production A/B in DESIGN.md 4.1): forming
// the 16 cross Gram entries of a block pair once per round and updating them by the cosines of the
// rotations ("lite" update) -- one reduction chain per round instead of four, 30 % fewer instructions,
// 4 % faster per sweep, but two more sweeps on graded spectra.
template <int NP, int A0, int A1, int A2, int A3, int B0, int B1, int B2, int B3>
__device__ __forceinline__ int sub_round(PRow<NP> (&y)[8], float (&a)[8], float tol2, float big2, int lane, bool& big) {
    constexpr int PA[4] = {A0, A1, A2, A3}, PB[4] = {B0, B1, B2, B3};
    float gr[4], gi[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) gram_part<NP>(y[PA[i]], y[PB[i]], gr[i], gi[i]);
    const float mgr = reduce4(gr, lane), mgi = reduce4(gi, lane);
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
    big = big || (dorot && g2 > big2 * apq);
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
            rot_apply<NP>(ci, sri, sii, y[PA[i]], y[PB[i]]);
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

// Step 2 of the kernel: one-sided Jacobi sweeps on the first 4 nb rows of Y (nb even), which are
// converted IN PLACE from interleaved complex to the planar layout of the sweep engine (a row keeps
// its slot of LS complex numbers: W = LC reals, then W imaginaries) and stay planar afterwards.
// 4-row blocks are paired by the circle method; a warp owns a pair of blocks per round, keeps its 8
// rows in registers and does the 16 cross rotations of the pair (plus the 12 inside the two blocks in
// the first round of a sweep) in sub-rounds of 4 disjoint rotations.  Squared row norms are cached in
// shared memory, refreshed once per sweep and updated by +-t|g| in between.
// Rounds are separated by PAIRWISE named barriers, not by a block-wide one: in the circle method
// group g takes its two blocks of the next round from groups g-1 and g+1 only, so a warp syncs with
// its two neighbours (edge (g, g+1) = barrier 1 + g, even warps right edge first, odd warps left edge
// first).  (Polling per-block round counters in shared memory was tried and was 3x slower.)
// A sweep in which no rotation exceeded cos 1e-3 is the last one: on the thetas of the chi = 64 workload a
// sweep whose largest rotation was eps leaves cosines of ~10 eps^2 (8.6e-3 -> 7.4e-4 -> 3e-6, the fp32 floor),
// so after such a sweep the rows are orthogonal to ~1e-5, i.e. sigma to 1e-10 and the kept subspace to 1e-5
// of an angle.  (1e-4, the first version: one more sweep on 35 % of the solves for identical singular values
// and backward error; 3e-3 fails the 1e-5 singular-value test on one fixture.)
template <int NP>
__device__ __noinline__ void jacobi_sweeps(cf* Ys, int LS, float* nrm, int nb, int max_sweeps, float tol2, float big2,
                                           int& sweeps_out, int& status_out) {
    constexpr int W = 64 * NP;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int mcirc = nb - 1, ngroups = nb / 2, nrounds = nb > 2 ? nb - 1 : 1;
    // interleaved -> planar, a row at a time (all of a row is in registers before any of it is written)
    for (int i = warp; i < 4 * nb; i += NW) {
        float* row = reinterpret_cast<float*>(Ys + (size_t)i * LS);
        if (NP == 2) {
            const float4 x0 = reinterpret_cast<const float4*>(row)[2 * lane];
            const float4 x1 = reinterpret_cast<const float4*>(row)[2 * lane + 1];
            __syncwarp();
            *reinterpret_cast<float4*>(row + 4 * lane) = make_float4(x0.x, x0.z, x1.x, x1.z);
            *reinterpret_cast<float4*>(row + W + 4 * lane) = make_float4(x0.y, x0.w, x1.y, x1.w);
        } else {
            const float4 x0 = reinterpret_cast<const float4*>(row)[lane];
            __syncwarp();
            *reinterpret_cast<float2*>(row + 2 * lane) = make_float2(x0.x, x0.z);
            *reinterpret_cast<float2*>(row + W + 2 * lane) = make_float2(x0.y, x0.w);
        }
    }
    __syncthreads();
    int sweeps = 0, status = 1;
    for (int sweep = 0; sweep < max_sweeps; ++sweep) {
        // refresh the cached squared norms
        for (int i = warp; i < 4 * nb; i += NW) {
            PRow<NP> r;
            prow_load<NP>(r, reinterpret_cast<const float*>(Ys + (size_t)i * LS), lane);
            float2 s2 = make_float2(0.f, 0.f);
#pragma unroll
            for (int t = 0; t < NP; ++t) { s2 = __ffma2_rn(r.re[t], r.re[t], s2); s2 = __ffma2_rn(r.im[t], r.im[t], s2); }
            const float sum = warp_sum(s2.x + s2.y);
            if (lane == 0) nrm[i] = sum;
        }
        __syncthreads();
        bool big = false;
        for (int r = 0; r < nrounds; ++r) {
            const int g = warp;
            if (g < ngroups) {
                int I, Jb;
                if (nb == 2) { I = 0; Jb = 1; }
                else if (g == 0) { I = mcirc; Jb = r; }
                else { I = (r + g) % mcirc; Jb = (r - g + mcirc) % mcirc; }
                float* rowA = reinterpret_cast<float*>(Ys + (size_t)(4 * I) * LS);
                float* rowB = reinterpret_cast<float*>(Ys + (size_t)(4 * Jb) * LS);
                const int rs = 2 * LS;                           // row stride in floats
                PRow<NP> v[8];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    prow_load<NP>(v[i], rowA + i * rs, lane);
                    prow_load<NP>(v[4 + i], rowB + i * rs, lane);
                }
                float a[8];
#pragma unroll
                for (int i = 0; i < 4; ++i) { a[i] = nrm[4 * I + i]; a[4 + i] = nrm[4 * Jb + i]; }
                int nrot = 0;
                if (r == 0) {                                    // inside the two blocks
                    nrot += sub_round<NP, 0, 2, 4, 6, 1, 3, 5, 7>(v, a, tol2, big2, lane, big);
                    nrot += sub_round<NP, 0, 1, 4, 5, 2, 3, 6, 7>(v, a, tol2, big2, lane, big);
                    nrot += sub_round<NP, 0, 1, 4, 5, 3, 2, 7, 6>(v, a, tol2, big2, lane, big);
                }
                nrot += sub_round<NP, 0, 1, 2, 3, 4, 5, 6, 7>(v, a, tol2, big2, lane, big);
                nrot += sub_round<NP, 0, 1, 2, 3, 5, 6, 7, 4>(v, a, tol2, big2, lane, big);
                nrot += sub_round<NP, 0, 1, 2, 3, 6, 7, 4, 5>(v, a, tol2, big2, lane, big);
                nrot += sub_round<NP, 0, 1, 2, 3, 7, 4, 5, 6>(v, a, tol2, big2, lane, big);
                if (nrot) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        prow_store<NP>(v[i], rowA + i * rs, lane);
                        prow_store<NP>(v[4 + i], rowB + i * rs, lane);
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
    sweeps_out = sweeps;
    status_out = status;
}

// Orthonormalisation of W [nv][k] (k <= 64) by Newton-Schulz steps, W <- W (3/2 I - 1/2 W^H W): two
// small GEMMs per step instead of the 2k barrier-separated Householder steps.  W = X V_k S^-1 is
// orthonormal up to E = W^H W - I with |E_ij| ~ 3e-6 sigma_i / sigma_j (a few 1e-3 in Frobenius norm
// on the thetas of a chi = 64 circuit); a step squares that (||E'|| <= 3/4 ||E||^2) and keeps
// span(W), and the split only depends on span(Q) (Q Q^H X is the projection either way; the basis
// inside the span is a gauge choice).  Steps are taken while ||E||_F < 0.05 (measured on the data,
// block-uniform), the last one from ||E||_F < 1e-3 (-> below 1e-6), at most three; a W further from
// orthonormal (zero / noise columns, sigma_k ~ 1e-6 sigma_1) is left to the Householder path, which
// completes the basis as LAPACK does.  Returns true when W is orthonormal.  M: scratch of 64 x 64.
__device__ bool newton_schulz_q(cf* W, int LS, int nv, int k, cf* M, float* red) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int step = 0; step < 3; ++step) {
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
        if (!(e2 < 0.05f * 0.05f)) { __syncthreads(); return false; }       // also catches NaN
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
            __syncthreads();                                       // every read of W (and of red) is done
            if (a < nv) {
#pragma unroll
                for (int u = 0; u < 16; ++u)
                    if (jq + u < k) W[(size_t)a * LS + jq + u] = acc[u];
            }
        }
        __syncthreads();
        if (e2 < 1e-3f * 1e-3f) return true;                       // this step brought ||E||_F below 1e-6
    }
    return true;                                                   // 0.05 -> 1.9e-3 -> 2.7e-6 -> 5e-12
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
    unsigned char* colinv = (unsigned char*)(scal + NW + 16);    // [128] column of X held by column c of Y

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
    for (int c = tid; c < 128; c += ST) colinv[c] = (unsigned char)c;
    __syncthreads();
    // Columns in order of descending norm before the QR (the cheap part of column pivoting: the rows
    // of R then come out graded and Jacobi needs ~0.5 sweeps fewer on the thetas of a truncated
    // circuit; a full pivoted QR gains no more, scripts/exp_sweeps_preconditioning.py).  Stable, so
    // equal columns -- diagonal and tied inputs -- stay where they are.  Y keeps the permuted
    // columns to the end; step 4 reads the columns of X in the same order.
    if (P.colsort && L >= 2) {
        float* cnp = reinterpret_cast<float*>(Xs);             // [nparts][L] partial squared column norms
        float* cn = cnp + ST;                                  // [L]
        unsigned char* crank = reinterpret_cast<unsigned char*>(cn + 128);
        // thread (c, part) sums rows part, part + nparts, ... of column c  (L <= 128 < ST); the parts are
        // added in a fixed order (run-to-run reproducible ranks)
        const int nparts = ST / L;
        {
            const int c = tid % L, part = tid / L;
            if (part < nparts) {
                float s2 = 0.f;
                for (int i = part; i < nv; i += nparts) s2 += cf_abs2(Ys[(size_t)i * LS + c]);
                cnp[part * L + c] = s2;
            }
        }
        __syncthreads();
        if (tid < L) {
            float s2 = 0.f;
            for (int part = 0; part < nparts; ++part) s2 += cnp[part * L + tid];
            cn[tid] = s2;
        }
        __syncthreads();
        if (tid < L) {
            const float mine = cn[tid];
            int rank = 0;
            for (int j = 0; j < L; ++j) { const float o = cn[j]; rank += (o > mine) || (o == mine && j < tid); }
            crank[tid] = (unsigned char)rank;
            colinv[rank] = (unsigned char)tid;
        }
        __syncthreads();
        // permute the columns of Y in place, a row per warp pass (the whole row is in registers
        // before any of it is written)
        for (int i = warp; i < nv; i += NW) {
            cf* yr = Ys + (size_t)i * LS;
            cf v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) { const int c = lane + 32 * u; v[u] = c < L ? yr[c] : cf_make(0.f, 0.f); }
            __syncwarp();
#pragma unroll
            for (int u = 0; u < 4; ++u) { const int c = lane + 32 * u; if (c < L) yr[crank[c]] = v[u]; }
        }
        __syncthreads();
    }

    PHASE_MARK(1);
    // ---- step 1: Householder QR, R only ------------------------------------------------------
    // (Measured and not adopted: the factorisation in panels of 32 steps on the shrinking trailing matrix,
    // so that the owners of finished columns do not idle -- 540 k -> 563 k cycles: a step is bound by its
    // ~2 000-cycle scalar / publish / barrier chain, not by the column update.)
    if (P.do_qr) householder<0>(Ys, LS, nv, L, min(nv - 1, L), vbuf, scal, nullptr, nullptr);

    PHASE_MARK(2);
    // ---- step 2: one-sided Jacobi on the rows of Y -------------------------------------------
    const int nact = P.do_qr ? min(nv, L) : nv;    // rows >= L of R are exactly zero
    const int nb = 2 * ((nact + 7) / 8);           // 4-row blocks (even count)
    int sweeps = 0, status = 0;
    if (nact >= 2) {
        if (P.LC == 128) jacobi_sweeps<2>(Ys, LS, nrm, nb, P.max_sweeps, P.tol2, P.big2, sweeps, status);
        else jacobi_sweeps<1>(Ys, LS, nrm, nb, P.max_sweeps, P.tol2, P.big2, sweeps, status);
    }
    const int W = P.LC;                            // rows < 4 nb are planar from here on (the others are zero)
    const int nplanar = nact >= 2 ? 4 * nb : 0;

    PHASE_MARK(3);
    // ---- step 3: singular values, stable descending sort --------------------------------------
    __syncthreads();
    for (int i = warp; i < nvp; i += NW) {
        float s2 = 0.f;
        if (i < nv && i < nplanar) {
            const float* yr = reinterpret_cast<const float*>(Ys + (size_t)i * LS);
            for (int c = lane; c < L; c += 32) s2 = fmaf(yr[c], yr[c], fmaf(yr[W + c], yr[W + c], s2));
        } else if (i < nv) {
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
                if (a_ < nv && cc < cw) { v = X[(size_t)a_ * L + colinv[c0 + cc]]; v.x *= scale_in; v.y *= scale_in; }
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
                        cf yb;
                        if (pj[jj] < nplanar) {
                            const float* yr = reinterpret_cast<const float*>(Ys + (size_t)pj[jj] * LS);
                            yb = cf_make(yr[c0 + cc], yr[W + c0 + cc]);
                        } else {
                            yb = Ys[(size_t)pj[jj] * LS + c0 + cc];
                        }
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
        const int ylanes = P.LC / 32;
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
    lo.LC = L <= 64 ? 64 : 128;  // row width of the planar sweep layout (lanes cover 2 or 4 columns)
    lo.LS = lo.LC + 4;           // rows 16-byte aligned; 4 consecutive rows x 4 columns hit 16 distinct 8-byte banks
    lo.smem = ((size_t)lo.nvp * lo.LS + XS_ELEMS + 2 * (size_t)VB + lo.nvp) * 8 + (size_t)lo.nvp * 16 + NW * 4 + 64 + 128;
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
    P.big2 = BIG2;
    P.do_qr = 1;
    P.use_ns = 1;
    P.colsort = 1;
    // debugging knobs (not part of the ABI)
    if (const char* e = mpsb_env("MPSB_SVD_MAX_SWEEPS")) P.max_sweeps = atoi(e);
    if (const char* e = mpsb_env("MPSB_SVD_LAST_COS")) { float c = (float)atof(e); P.big2 = c * c; }
    if (const char* e = mpsb_env("MPSB_SVD_NO_COLSORT")) P.colsort = atoi(e) ? 0 : 1;
    if (const char* e = mpsb_env("MPSB_SVD_NO_QR")) P.do_qr = atoi(e) ? 0 : 1;
    if (const char* e = mpsb_env("MPSB_SVD_NO_NS")) P.use_ns = atoi(e) ? 0 : 1;
    MPSB_CUDA(cudaFuncSetAttribute(svd_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lo.smem));
    svd_small_kernel<<<njobs, ST, lo.smem, st>>>(P);
    MPSB_LAUNCH_CHECK("svd_small_kernel");
    return 0;
}
