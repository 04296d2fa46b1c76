// Truncated SVD for d*chi <= 128: one CTA per matrix, everything resident in shared memory.
//
// Replaces tn.split_node_full_svd -> np.linalg.svd (LAPACK zgesdd) + slice [:k] + the two
// absorb contractions of mpsim/core.py:1132-1152.
//
// Input  X [nv][L] : the theta matrix oriented so that its ROWS are the vectors to
//                    orthogonalise (left-canonical: X = theta, nv = d*chiL, L = d*chiR;
//                    otherwise X = theta^T).
// Method (CPU model with identical arithmetic: tests/_jacobi_model.py):
//   1. Householder QR preconditioning  X = Q R ; Y <- R, Z <- Q^H.  Jacobi on R converges in
//      ~9 sweeps independently of how graded the spectrum is (plain Jacobi on X needs 15-25).
//   2. One-sided Jacobi on the rows of Y, the same 2x2 unitaries accumulated into Z, so that
//      Z X0 == Y always.  Rows are visited by a block tournament: 4-row blocks are paired by the
//      circle method; a warp owns a pair of blocks, keeps its 8 rows of Y in registers
//      (lane owns elements lane, lane+32, ...), performs the 16 cross rotations (plus the 12
//      intra-block ones in the first round of a sweep) with warp-shuffle reductions, logs them
//      to shared memory and replays them on its 8 rows of Z in the same registers (two phases
//      keep the kernel under 128 registers, so 16 warps are resident instead of 8).
//   3. sigma_j = |Y_j|, stable descending rank sort (ties keep the lower index: this is what
//      reproduces the reference on Bell + maxsvals=1, README.md:48-53), keep the first k.
//   4. Split/absorb without any division: the isometry is conj(Z) (a product of unitaries, so
//      orthonormal even for zero singular values, like LAPACK's), the weighted factor is Y:
//        left-canonical : left  = Z_k^H   (U)     right = Y_k      (S.Vh)
//        otherwise      : left  = Y_k^T   (U.S)   right = conj(Z_k) (Vh)
//      left.right == exact rank-k projection of theta whether or not Jacobi converged.
#include "common.cuh"
#include <stdlib.h>

namespace {

constexpr int ST = 512;          // threads per CTA
constexpr int NW = ST / 32;      // warps: one 8-row group each at nv = 128
constexpr int EPL = 4;           // elements per lane per row (row length <= 128)
constexpr int NSLOT = 28;        // rotations per group step: 12 intra-block + 16 cross

struct SvdSmallParams {
    const cf* X; int64_t x_stride;
    int nv, L, k, lc;
    int nvp, LS, ZS, LC, ZC;     // padded rows, smem strides, columns touched by lanes
    int nz_smem;                 // Z rows [0, nz_smem) live in shared memory, the rest in zg
    cf* zg; int64_t zg_stride;
    const mpsb_gate2_desc* descs; int nbatch;     // output mode A (descs != nullptr)
    cf* left; int64_t left_stride; cf* right; int64_t right_stride;   // output mode B
    float* svals; int64_t svals_stride;
    int32_t* info;
    int max_sweeps; float tol2; int do_qr;
};

// (c, s, t|g|) of [[c, s], [-conj(s), c]] diagonalising [[a, g], [conj(g), b]]; s first, then
// c = sqrt(1 - |s|^2) derived from s (series near 1) so the rotation is unitary to rounding
// WITHOUT bias -- see tests/_jacobi_model.py:rotation_params.
__device__ __forceinline__ void rot_params(float a, float b, float gr, float gi, float g2,
                                           float& c, float& sr, float& si, float& tg) {
    float rg = rsqrtf(g2);
    float zeta = (a - b) * (0.5f * rg);
    float t = copysignf(1.0f, zeta) / (fabsf(zeta) + sqrtf(fmaf(zeta, zeta, 1.0f)));
    float ct = (t * rsqrtf(fmaf(t, t, 1.0f))) * rg;
    sr = ct * gr;
    si = ct * gi;
    float h = fmaf(sr, sr, si * si);
    float poly = fmaf(h, fmaf(h, fmaf(h, fmaf(h, 0.02734375f, 0.0390625f), 0.0625f), 0.125f), 0.5f);
    float c_series = fmaf(-h, poly, 1.0f);
    float c_sqrt = sqrtf(fmaf(-sr, sr, fmaf(-si, si, 1.0f)));
    c = (h < 0.0625f) ? c_series : c_sqrt;
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

// One sub-round on the Y rows of a group: 4 disjoint pairs (A_i, B_i).  The four Gram entries are
// reduced together; lane l then computes the rotation of pair (l >> 3) only and the parameters
// are exchanged by shuffles (4x fewer scalar instructions than every lane doing all four).
// The rotations are logged to rot[slot..slot+3] = (c, s.re, s.im, rotated?) for the Z phase.
template <int A0, int A1, int A2, int A3, int B0, int B1, int B2, int B3>
__device__ __forceinline__ int sub_round_y(cf (&y)[8][EPL], float (&a)[8], float tol2,
                                           float4* rot, int slot, int lane) {
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
    float c = 1.f, sr = 0.f, si = 0.f, tg = 0.f;
    const bool dorot = (g2 > tol2 * ap * aq) && (g2 > 1e-30f);
    if (dorot) rot_params(ap, aq, mgr, mgi, g2, c, sr, si, tg);
    if ((lane & 7) == 0) rot[slot + sel] = make_float4(c, sr, si, dorot ? 1.f : 0.f);
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

// The same sub-round replayed on the Z rows from the logged rotations.
template <int A0, int A1, int A2, int A3, int B0, int B1, int B2, int B3>
__device__ __forceinline__ void sub_round_z(cf (&z)[8][EPL], const float4* rot, int slot) {
    constexpr int PA[4] = {A0, A1, A2, A3}, PB[4] = {B0, B1, B2, B3};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float4 r = rot[slot + i];                // broadcast load
        if (r.w != 0.f) {
#pragma unroll
            for (int t = 0; t < EPL; ++t) rot_apply(r.x, r.y, r.z, z[PA[i]][t], z[PB[i]][t]);
        }
    }
}

__global__ void __launch_bounds__(ST, 1) svd_small_kernel(SvdSmallParams P) {
    extern __shared__ float4 smem_raw[];
    const int job = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nv = P.nv, L = P.L, nvp = P.nvp, LS = P.LS, ZS = P.ZS;

    float4* rotbuf = smem_raw;                     // [NW][NSLOT] (16-byte aligned first)
    cf* Ys = (cf*)(rotbuf + NW * NSLOT);           // [nvp][LS]
    cf* Zs = Ys + (size_t)nvp * LS;                // [nz_smem][ZS]
    cf* vbuf = Zs + (size_t)P.nz_smem * ZS;        // [2][nvp]
    float* sig = (float*)(vbuf + 2 * nvp);         // [nvp]
    int* perm = (int*)(sig + nvp);                 // [nvp]
    float* scal = (float*)(perm + nvp);            // [NW] (also the QR scalars: 8 used)
    int* cnt = (int*)(scal + NW);                  // [2]
    cf* zg = P.zg + (size_t)job * P.zg_stride;
    const int nzs = P.nz_smem;
    auto zrow = [&](int i) -> cf* { return i < nzs ? Zs + (size_t)i * ZS : zg + (size_t)(i - nzs) * ZS; };

    // ---- phase 0: load X scaled by an exact power of two so that max|x| is in [1, 2), Z = I ----
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
    for (int e = tid; e < nv * ZS; e += ST) {
        int i = e / ZS, c = e - i * ZS;
        zrow(i)[c] = cf_make(c == i ? 1.f : 0.f, 0.f);
    }
    if (tid < 2) cnt[tid] = 0;
    __syncthreads();

    // ---- phase 1: Householder QR, two threads per column of [Y | Z] (rows split even/odd) ----
    const int J = P.do_qr ? min(nv - 1, L) : 0;
    if (J > 0) {
        if (warp == 0) {
            float t2 = 0.f;
            for (int i = lane; i < nv; i += 32) {
                cf v = Ys[(size_t)i * LS];
                vbuf[i] = v;
                if (i > 0) t2 += cf_abs2(v);
            }
            t2 = warp_sum(t2);
            if (lane == 0) { cf x0 = Ys[0]; scal[0] = x0.x; scal[1] = x0.y; scal[2] = t2; }
        }
        __syncthreads();
        const int colid = tid >> 1, half = tid & 1;
        const bool isY = colid < L;
        const bool isZ = !isY && (colid - L) < nv;
        const int col = isY ? colid : colid - L;
        for (int j = 0; j < J; ++j) {
            const int cur = j & 1, nxt = cur ^ 1;
            const cf* vb = vbuf + cur * nvp;
            cf* vn = vbuf + nxt * nvp;
            cf x0 = cf_make(scal[cur * 4 + 0], scal[cur * 4 + 1]);
            const float tail2 = scal[cur * 4 + 2];
            const bool record = isY && (col == j + 1) && (j + 1 < J);
            float ax0sq = cf_abs2(x0);
            // |x0|^2 below ~1e-30 is a denormal-range number with few significant bits: the
            // phase x0/|x0| would be off by 1e-4 and the reflector no longer unitary (seen on
            // GHZ circuits).  The matrix is scaled to max|x| in [1,2), so such an x0 is noise.
            if (ax0sq < 1e-30f) { x0 = cf_make(0.f, 0.f); ax0sq = 0.f; }
            // skip the reflector when the column is already reduced, or so small that
            // 1/|x|^2 would overflow (QR is only a preconditioner: any unitary Z is valid)
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
            const bool upd = reflect && ((isY && col > j) || isZ);
            // This thread's rows are j + half, j + half + 2, ...; row j (v0 instead of vb[j]) belongs
            // to half == 0.  A Z column is split at nzs into its shared-memory and its spilled part
            // so that the loops run on plain pointers with constant strides.
            const int ifirst = j + half + (half == 0 ? 2 : 0);      // first row handled with vb[]
            const int zsplit = min(nv, max(ifirst, nzs + ((nzs ^ ifirst) & 1)));   // first spilled row of my parity
            cf* prow_j = isY ? &Ys[(size_t)j * LS + col] : &zrow(j)[col];
            // pass 1: w = v^H A[:, col]
            cf w = cf_make(0.f, 0.f);
            if (upd) {
                if (half == 0) w = cf_fma_conja(v0, *prow_j, w);
                if (isY) {
                    const cf* a = Ys + (size_t)ifirst * LS + col;
                    for (int i = ifirst; i < nv; i += 2, a += 2 * LS) w = cf_fma_conja(vb[i], *a, w);
                } else {
                    const cf* a = Zs + (size_t)ifirst * ZS + col;
                    int i = ifirst;
                    for (; i < zsplit; i += 2, a += 2 * ZS) w = cf_fma_conja(vb[i], *a, w);
                    const cf* b = zg + (size_t)(i - nzs) * ZS + col;
                    for (; i < nv; i += 2, b += 2 * ZS) w = cf_fma_conja(vb[i], *b, w);
                }
            }
            w.x += __shfl_xor_sync(0xffffffffu, w.x, 1);
            w.y += __shfl_xor_sync(0xffffffffu, w.y, 1);
            // pass 2: A[:, col] -= tau v w ; the owner of column j+1 also records the next reflector
            float t2 = 0.f;
            if (upd) {
                const cf tw = cf_scale(-tau, w);
                if (half == 0) *prow_j = cf_fma(v0, tw, *prow_j);
                if (record) {
                    cf* a = Ys + (size_t)ifirst * LS + col;
                    if (half == 1) {                       // row j+1 is mine: it is the next x0
                        // (ifirst == j + 1 for half == 1)
                    }
                    for (int i = ifirst; i < nv; i += 2, a += 2 * LS) {
                        cf nvl = cf_fma(vb[i], tw, *a);
                        *a = nvl;
                        vn[i] = nvl;
                        if (i > j + 1) t2 += cf_abs2(nvl);
                        else { scal[nxt * 4 + 0] = nvl.x; scal[nxt * 4 + 1] = nvl.y; }
                    }
                } else if (isY) {
                    cf* a = Ys + (size_t)ifirst * LS + col;
                    for (int i = ifirst; i < nv; i += 2, a += 2 * LS) *a = cf_fma(vb[i], tw, *a);
                } else {
                    cf* a = Zs + (size_t)ifirst * ZS + col;
                    int i = ifirst;
                    for (; i < zsplit; i += 2, a += 2 * ZS) *a = cf_fma(vb[i], tw, *a);
                    cf* b = zg + (size_t)(i - nzs) * ZS + col;
                    for (; i < nv; i += 2, b += 2 * ZS) *b = cf_fma(vb[i], tw, *b);
                }
            } else if (reflect && isY && col == j) {
                for (int i = j + half; i < nv; i += 2) Ys[(size_t)i * LS + j] = (i == j) ? alpha : cf_make(0.f, 0.f);
            } else if (!reflect && record) {
                for (int i = j + 1 + half; i < nv; i += 2) {
                    cf v = Ys[(size_t)i * LS + col];
                    vn[i] = v;
                    if (i > j + 1) t2 += cf_abs2(v);
                    else { scal[nxt * 4 + 0] = v.x; scal[nxt * 4 + 1] = v.y; }
                }
            }
            t2 += __shfl_xor_sync(0xffffffffu, t2, 1);
            if (record && half == 0) scal[nxt * 4 + 2] = t2;
            __syncthreads();
        }
    }

    // ---- phase 2: one-sided Jacobi on the rows of Y ------------------------------------------
    const int nact = P.do_qr ? min(nv, L) : nv;    // rows >= L of R are exactly zero
    const int nb = 2 * ((nact + 7) / 8);           // 4-row blocks (even count)
    const int mcirc = nb - 1;
    const int ngroups = nb / 2;
    const int nrounds = nb > 2 ? nb - 1 : 1;
    const int ylanes = P.LC / 32, zlanes = P.ZC / 32;
    float4* rot = rotbuf + warp * NSLOT;
    int sweeps = 0, status = 0;
    if (nact >= 2) {
        status = 1;
        for (int sweep = 0; sweep < P.max_sweeps; ++sweep) {
            int my_rot = 0;
            for (int r = 0; r < nrounds; ++r) {
                for (int g = warp; g < ngroups; g += NW) {
                    int I, Jb;
                    if (nb == 2) { I = 0; Jb = 1; }
                    else if (g == 0) { I = mcirc; Jb = r; }
                    else { I = (r + g) % mcirc; Jb = (r - g + mcirc) % mcirc; }
                    int rows[8];
#pragma unroll
                    for (int i = 0; i < 4; ++i) { rows[i] = 4 * I + i; rows[4 + i] = 4 * Jb + i; }
                    cf v[8][EPL];                  // first the Y rows, later re-used for the Z rows
                    float a[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const cf* yr = Ys + (size_t)rows[i] * LS;
                        float s2 = 0.f;
#pragma unroll
                        for (int t = 0; t < EPL; ++t) {
                            v[i][t] = (t < ylanes) ? yr[lane + 32 * t] : cf_make(0.f, 0.f);
                            s2 += cf_abs2(v[i][t]);
                        }
                        a[i] = s2;
                    }
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
                        for (int i = 0; i < 8; ++i) a[i] += __shfl_xor_sync(0xffffffffu, a[i], o);
                    int nrot = 0;
                    const bool intra = (r == 0);
                    if (intra) {
                        nrot += sub_round_y<0, 2, 4, 6, 1, 3, 5, 7>(v, a, P.tol2, rot, 0, lane);
                        nrot += sub_round_y<0, 1, 4, 5, 2, 3, 6, 7>(v, a, P.tol2, rot, 4, lane);
                        nrot += sub_round_y<0, 1, 4, 5, 3, 2, 7, 6>(v, a, P.tol2, rot, 8, lane);
                    }
                    nrot += sub_round_y<0, 1, 2, 3, 4, 5, 6, 7>(v, a, P.tol2, rot, 12, lane);
                    nrot += sub_round_y<0, 1, 2, 3, 5, 6, 7, 4>(v, a, P.tol2, rot, 16, lane);
                    nrot += sub_round_y<0, 1, 2, 3, 6, 7, 4, 5>(v, a, P.tol2, rot, 20, lane);
                    nrot += sub_round_y<0, 1, 2, 3, 7, 4, 5, 6>(v, a, P.tol2, rot, 24, lane);
                    if (nrot) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            cf* yr = Ys + (size_t)rows[i] * LS;
#pragma unroll
                            for (int t = 0; t < EPL; ++t)
                                if (t < ylanes) yr[lane + 32 * t] = v[i][t];
                        }
                        // Z phase: same registers, rotations replayed from the log
                        __syncwarp();
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const bool ok = rows[i] < nv;
                            const cf* zr = ok ? zrow(rows[i]) : Ys;
#pragma unroll
                            for (int t = 0; t < EPL; ++t)
                                v[i][t] = (ok && t < zlanes) ? zr[lane + 32 * t] : cf_make(0.f, 0.f);
                        }
                        if (intra) {
                            sub_round_z<0, 2, 4, 6, 1, 3, 5, 7>(v, rot, 0);
                            sub_round_z<0, 1, 4, 5, 2, 3, 6, 7>(v, rot, 4);
                            sub_round_z<0, 1, 4, 5, 3, 2, 7, 6>(v, rot, 8);
                        }
                        sub_round_z<0, 1, 2, 3, 4, 5, 6, 7>(v, rot, 12);
                        sub_round_z<0, 1, 2, 3, 5, 6, 7, 4>(v, rot, 16);
                        sub_round_z<0, 1, 2, 3, 6, 7, 4, 5>(v, rot, 20);
                        sub_round_z<0, 1, 2, 3, 7, 4, 5, 6>(v, rot, 24);
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            if (rows[i] < nv) {
                                cf* zr = zrow(rows[i]);
#pragma unroll
                                for (int t = 0; t < EPL; ++t)
                                    if (t < zlanes) zr[lane + 32 * t] = v[i][t];
                            }
                        }
                        my_rot += nrot;
                        __syncwarp();              // the log is rewritten by the next group step
                    }
                }
                __syncthreads();
            }
            if (lane == 0 && my_rot) atomicAdd(&cnt[sweep & 1], my_rot);
            __syncthreads();
            int total = cnt[sweep & 1];
            if (tid == 0) cnt[(sweep + 1) & 1] = 0;
            sweeps = sweep + 1;
            if (total == 0) { status = 0; break; }
            // the reset of the other counter is ordered before its next use by the barriers above
        }
    }

    // ---- phase 3: singular values, stable descending sort, split/absorb -----------------------
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

    const int k = P.k;
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
    if (P.lc) {
        // right [k][L] = Y_k ; left [nv][k] = conj(Z_k)^T
        for (int e = tid; e < k * L; e += ST) { int j = e / L, c = e - j * L; right[e] = cf_scale(scale_out, Ys[(size_t)perm[j] * LS + c]); }
        for (int e = tid; e < nv * k; e += ST) { int a_ = e / k, j = e - a_ * k; left[e] = cf_conj(zrow(perm[j])[a_]); }
    } else {
        // left [L][k] = Y_k^T ; right [k][nv] = conj(Z_k)
        for (int e = tid; e < L * k; e += ST) { int a_ = e / k, j = e - a_ * k; left[e] = cf_scale(scale_out, Ys[(size_t)perm[j] * LS + a_]); }
        for (int e = tid; e < k * nv; e += ST) { int j = e / nv, b_ = e - j * nv; right[e] = cf_conj(zrow(perm[j])[b_]); }
    }
    if (sv) for (int j = tid; j < min(nv, L); j += ST) sv[j] = scale_out * sig[perm[j]];
    if (P.info && tid == 0) { P.info[2 * job] = status; P.info[2 * job + 1] = sweeps; }
}

struct Layout { int nvp, LC, LS, ZC, ZS, nz_smem; size_t smem; };

Layout make_layout(int nv, int L) {
    Layout lo;
    lo.nvp = (nv + 7) / 8 * 8;
    lo.LC = (L + 31) / 32 * 32; lo.LS = lo.LC + 1;
    lo.ZC = (nv + 31) / 32 * 32; lo.ZS = lo.ZC + 1;
    size_t fixed = (size_t)NW * NSLOT * 16 + (size_t)lo.nvp * lo.LS * 8 + (size_t)2 * lo.nvp * 8 + (size_t)lo.nvp * 8 + NW * 4 + 2 * 4 + 64;
    int dev = 0, optin = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess || optin <= 0)
        optin = 232448;      // sm_100: 227 KB (also the value assumed when planning without a device)
    size_t avail = (size_t)optin > fixed ? (size_t)optin - fixed : 0;
    size_t per_row = (size_t)lo.ZS * 8;
    int fit = (int)(avail / per_row);
    lo.nz_smem = fit >= nv ? nv : fit;
    lo.smem = fixed + (size_t)lo.nz_smem * per_row;
    return lo;
}

}  // namespace

size_t svd_small_global_z_elems(int nv, int L) {
    Layout lo = make_layout(nv, L);
    return (size_t)(nv - lo.nz_smem) * lo.ZS;
}

int launch_svd_small(const cf* X, int64_t x_job_stride, int njobs, int nv, int L, int k,
                     int left_canonical, const mpsb_gate2_desc* descs, int ndesc, int nbatch,
                     cf* left, int64_t left_stride, cf* right, int64_t right_stride,
                     float* svals, int64_t svals_stride, int32_t* info, cf* zglobal,
                     cudaStream_t st) {
    (void)ndesc;
    if (njobs <= 0) return 0;
    MPSB_ARG(nv >= 1 && L >= 1 && nv <= MPSB_MAX_SMALL_DIM && L <= MPSB_MAX_SMALL_DIM,
             "svd_small: shape %d x %d outside [1, %d]", nv, L, MPSB_MAX_SMALL_DIM);
    MPSB_ARG(k >= 0 && k <= (nv < L ? nv : L), "svd_small: k=%d out of range for %d x %d", k, nv, L);
    Layout lo = make_layout(nv, L);
    SvdSmallParams P;
    P.X = X; P.x_stride = x_job_stride;
    P.nv = nv; P.L = L; P.k = k; P.lc = left_canonical;
    P.nvp = lo.nvp; P.LS = lo.LS; P.ZS = lo.ZS; P.LC = lo.LC; P.ZC = lo.ZC;
    P.nz_smem = lo.nz_smem;
    P.zg = zglobal; P.zg_stride = (int64_t)(nv - lo.nz_smem) * lo.ZS;
    MPSB_ARG(P.zg_stride == 0 || zglobal != nullptr, "svd_small: Z spill workspace missing");
    P.descs = descs; P.nbatch = nbatch > 0 ? nbatch : 1;
    P.left = left; P.left_stride = left_stride; P.right = right; P.right_stride = right_stride;
    P.svals = svals; P.svals_stride = svals_stride;
    P.info = info;
    P.max_sweeps = 30;
    P.tol2 = 3e-6f * 3e-6f;
    P.do_qr = 1;
    // debugging knobs (not part of the ABI)
    if (const char* e = getenv("MPSB_SVD_MAX_SWEEPS")) P.max_sweeps = atoi(e);
    if (const char* e = getenv("MPSB_SVD_NO_QR")) P.do_qr = atoi(e) ? 0 : 1;
    MPSB_CUDA(cudaFuncSetAttribute(svd_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lo.smem));
    svd_small_kernel<<<njobs, ST, lo.smem, st>>>(P);
    MPSB_LAUNCH_CHECK("svd_small_kernel");
    return 0;
}
