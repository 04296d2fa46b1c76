// Truncated SVD for d*chi > 128 (block one-sided Jacobi in global memory) -- see DESIGN.md.
#include "common.cuh"

size_t svd_large_workspace_elems(int nv, int L) {
    (void)nv; (void)L;
    return 0;
}

int launch_svd_large(cf* X, int64_t x_job_stride, int njobs, int nv, int L, int k,
                     int left_canonical, const mpsb_gate2_desc* descs, int ndesc, int nbatch,
                     cf* left, int64_t left_stride, cf* right, int64_t right_stride,
                     float* svals, int64_t svals_stride, int32_t* info, cf* work,
                     cudaStream_t st) {
    (void)X; (void)x_job_stride; (void)njobs; (void)k; (void)left_canonical; (void)descs; (void)ndesc;
    (void)nbatch; (void)left; (void)left_stride; (void)right; (void)right_stride; (void)svals;
    (void)svals_stride; (void)info; (void)work; (void)st;
    MPSB_ARG(false, "svd_large: %d x %d not implemented yet", nv, L);
}
