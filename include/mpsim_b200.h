/*
 * mpsim_b200 -- C ABI of the B200-native MPS gate-application path.
 *
 * The reference (grmlarose/mpsim) has no FFI: its boundary for this path is the Python class
 * API above tensornetwork.  Each entry point below replaces the group of
 * tensornetwork/numpy calls cited next to it (paths are into the reference tree).  The Python
 * classes in mpsim_b200/ (MPS, MPSOperation, MPSimulator: same names, arguments and errors as
 * the reference) bind these symbols with ctypes; INTEGRATION.md shows the binding.
 *
 * Conventions
 *   - plain C types and raw DEVICE pointers only; the caller (PyTorch) owns every buffer,
 *     the library never allocates or frees persistent memory;
 *   - complex64 = interleaved float2;  a site tensor is dense row-major [chi_left][d][chi_right];
 *   - every call is asynchronous on the given cudaStream_t (passed as void*);
 *   - return value: 0 ok; <0 argument error (nothing launched, text in mpsb_last_error());
 *     >0 a cudaError_t;
 *   - "batch": the same descriptor list is applied to nbatch independent MPS whose buffers are
 *     laid out at a fixed element stride (bs_*) from the descriptor's base pointers.
 */
#ifndef MPSIM_B200_H
#define MPSIM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MPSB_VERSION 100
#define MPSB_MAX_SMALL_DIM 128   /* d*chi up to which the single-CTA shared-memory SVD is used */

/* One adjacent two-qudit application on sites (i, i+1).
 * Replaces mpsim/core.py:1046-1152 (connect, contract_between, flatten_edges_between, contract,
 * split_node_full_svd, contract_between(s, vdag) / (u, s), write back). */
typedef struct mpsb_gate2_desc {
    const void* site_l;    /* complex64 [chiL][d][chiM] */
    const void* site_r;    /* complex64 [chiM][d][chiR] */
    void*       out_l;     /* complex64 [chiL][d][k]     (may alias site_l) */
    void*       out_r;     /* complex64 [k][d][chiR]     (may alias site_r) */
    const void* gate;      /* complex64 [d][d][d][d]  G[o1][o2][p][q], edge convention of core.py:986-992 */
    float*      svals;     /* float [min(d*chiL, d*chiR)] all singular values, descending; or NULL */
    int64_t bs_site_l, bs_site_r, bs_out_l, bs_out_r, bs_gate, bs_svals;  /* batch strides, elements */
} mpsb_gate2_desc;

/* One one-qudit application.  Replaces mpsim/core.py:819-826 (connect + contract). */
typedef struct mpsb_gate1_desc {
    const void* site;      /* complex64 [chiL][d][chiR] */
    void*       out;       /* same shape (may alias site) */
    const void* gate;      /* complex64 [d][d]  g[o][p]; axis 1 contracts with the site (core.py:773-775) */
    int64_t bs_site, bs_out, bs_gate;
    int32_t chiL, chiR;
} mpsb_gate1_desc;

/* A site of a chain, for the whole-chain contractions. */
typedef struct mpsb_site_ref {
    const void* site;      /* complex64 [chiL][d][chiR] */
    int64_t bs;            /* batch stride, elements */
    int32_t chiL, chiR;
} mpsb_site_ref;

int         mpsb_version(void);
const char* mpsb_last_error(void);           /* thread-local text of the last failure */

/* Device properties the host planner needs (SM count, opt-in shared memory per block). */
int mpsb_device_info(int* sm_count, int* max_smem_optin, int* cc_major, int* cc_minor);

/* Workspace (bytes) mpsb_apply_gate2 needs for ndesc*nbatch applications of one shape. */
size_t mpsb_gate2_workspace_bytes(int ndesc, int nbatch, int d, int chiL, int chiM, int chiR, int k);

/* Apply ndesc*nbatch two-qudit gates of ONE shape class (chiL, chiM, chiR, k, left_canonical):
 * theta contraction with the gate folded in, truncated SVD (kept count k is data independent:
 * k = min(maxsvals, d*chiL, d*chiR), zeros are kept; core.py:1105-1137), split/absorb
 * (left_canonical: left = U, right = S.Vh; else left = U.S, right = Vh; core.py:1140-1145).
 * info (optional, int32[ndesc*nbatch][2]) receives {status, sweeps}; status 0 = converged,
 * 1 = sweep limit reached (result is still an exact rank-k projection of theta). */
int mpsb_apply_gate2(const mpsb_gate2_desc* descs_dev, int ndesc, int nbatch,
                     int d, int chiL, int chiM, int chiR, int k, int left_canonical,
                     void* workspace, size_t workspace_bytes, int32_t* info, void* stream);

/* One shape class of a circuit layer: the arguments of one mpsb_apply_gate2 call. */
typedef struct mpsb_gate2_group {
    const mpsb_gate2_desc* descs_dev;   /* this group's descriptors (device memory) */
    int32_t* info;                      /* device int32 [ndesc*nbatch][2], or NULL */
    int32_t ndesc, chiL, chiM, chiR, k, left_canonical;
} mpsb_gate2_group;

/* All shape classes of ONE layer of disjoint applications (the moment dispatcher's unit,
 * core.py:1221-1247 "TODO: Parallelize") in one call: same result as ngroups mpsb_apply_gate2
 * calls, but the groups run concurrently on library-owned streams forked from / joined to
 * `stream` -- a group of one or two large matrices is latency bound and hides behind the
 * layer's main group.  groups_host is read on the host during the call.  Workspace:
 * mpsb_gate2_layer_workspace_bytes() (the groups' own needs, 256-byte aligned, summed). */
size_t mpsb_gate2_layer_workspace_bytes(const mpsb_gate2_group* groups_host, int ngroups, int nbatch, int d);
int mpsb_apply_gate2_layer(const mpsb_gate2_group* groups_host, int ngroups, int nbatch, int d,
                           void* workspace, size_t workspace_bytes, void* stream);

/* A'[l,o,r] = sum_p g[o,p] A[l,p,r] for ndesc*nbatch sites (core.py:819-826). */
int mpsb_apply_gate1(const mpsb_gate1_desc* descs_dev, int ndesc, int nbatch, int d,
                     int max_site_elems, void* stream);

/* <a|b> by the transfer-matrix chain of core.py:543-561 (sum a * conj(b), as the reference
 * writes it).  out: complex64[nbatch].  workspace >= mpsb_inner_workspace_bytes(). */
size_t mpsb_inner_workspace_bytes(int nbatch, int d, int max_chi_a, int max_chi_b);
int mpsb_inner_products(const mpsb_site_ref* a_host, const mpsb_site_ref* b_host, int nsites,
                        int nbatch, int d, void* out, void* workspace, size_t workspace_bytes,
                        void* stream);

/* site <- factor[b] * site for every site of every batch member (renormalize, core.py:590-594). */
int mpsb_scale_sites(const mpsb_site_ref* sites_dev, int nsites, int nbatch, int d,
                     const float* factors_dev, int max_site_elems, void* stream);

/* Gauge rebalance of every batch member's chain: when the binary exponents of the sites' largest
 * entries are more than spread_log2 apart, multiply each site by a power of two so that they
 * agree; the shifts of a chain sum to zero, so the state is unchanged bit for bit.  No reference
 * counterpart: the reference stores complex128, whose exponent range absorbs the drift of its
 * alternating left/right-canonical sweeps (mpsim/core.py:1348-1360); complex64 storage needs this
 * after a few hundred layers.  shifts_dev: int32 [nsites][nbatch] scratch, holds the applied
 * shifts on return. */
int mpsb_rebalance_sites(const mpsb_site_ref* sites_dev, int nsites, int nbatch, int d, int spread_log2,
                         int* shifts_dev, int max_site_elems, void* stream);

/* Dense wavefunction of one batch member, big-endian (core.py:483-500): out complex64[d^n].
 * workspace >= mpsb_wavefunction_workspace_bytes(). */
size_t mpsb_wavefunction_workspace_bytes(const mpsb_site_ref* sites_host, int nsites, int d);
int mpsb_wavefunction(const mpsb_site_ref* sites_host, int nsites, int d, int batch_index,
                      void* out, void* workspace, size_t workspace_bytes, void* stream);

/* Amplitudes <bits|psi> for nbits given basis states (uint8 [nbits][nsites]) of every batch
 * member: out complex64 [nbatch][nbits]. */
int mpsb_amplitudes(const mpsb_site_ref* sites_dev, int nsites, int nbatch, int d, int max_chi,
                    const uint8_t* bits_dev, int nbits, void* out, void* stream);

/* Building blocks exported for tests and micro-benchmarks (same kernels the calls above use). */
/* C[M][N] = op(A)[M][K] . op(B)[K][N], complex64, element strides given explicitly. */
int mpsb_cgemm(const void* A, int64_t a_rs, int64_t a_cs, int conj_a, int64_t a_bs,
               const void* B, int64_t b_rs, int64_t b_cs, int conj_b, int64_t b_bs,
               void* C, int64_t c_ld, int64_t c_bs, int M, int N, int K, int nbatch, void* stream);
/* Same product on the tensor cores (tcgen05 3xTF32, tc_gemm.cu): dense row-major A [M][K],
 * B [K][N], C [M][N] (row stride c_ld), batch strides in elements. */
size_t mpsb_cgemm_tc_workspace_bytes(int M, int N, int K, int nbatch);
int mpsb_cgemm_tc(const void* A, int64_t a_bs, const void* B, int64_t b_bs, void* C, int64_t c_ld, int64_t c_bs,
                  int M, int N, int K, int nbatch, void* workspace, size_t workspace_bytes, void* stream);
/* theta matrix only: out[job] = row-major [d*chiL][d*chiR] with rows (l,o1), cols (o2,r).
 * With a workspace of mpsb_theta_workspace_bytes() the tensor-core kernel is used where it
 * applies (d = 2, chi >= 32); without one the FFMA kernel. */
size_t mpsb_theta_workspace_bytes(int ndesc, int nbatch, int d, int chiL, int chiM, int chiR);
int mpsb_theta(const mpsb_gate2_desc* descs_dev, int ndesc, int nbatch, int d,
               int chiL, int chiM, int chiR, void* out, void* workspace, size_t workspace_bytes,
               void* stream);
/* SVD of njobs dense row-major complex64 matrices [m][n] (stride m*n): left [m][k], right [k][n],
 * svals [min(m,n)], info [njobs][2]. */
size_t mpsb_svd_workspace_bytes(int njobs, int m, int n);
int mpsb_svd(const void* mats, int njobs, int m, int n, int k, int left_canonical,
             void* left, void* right, float* svals, int32_t* info,
             void* workspace, size_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MPSIM_B200_H */
