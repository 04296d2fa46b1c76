// HBM-bound site kernels: one-qudit gate, per-chain scaling, amplitude chains.
#include <climits>
#include "common.cuh"

namespace {

// A'[l][o][r] = sum_p g[o][p] A[l][p][r]     (mpsim/core.py:819-826)
// One thread per (l, r) -- or per (l, r, r+1) with 16-byte accesses when chiR is even (VEC = 2; site
// slots and batch strides are 128-byte aligned): reads the d physical components (each a coalesced
// stream over r), writes d.  Algorithmic traffic: 16*d*chiL*chiR bytes per site.
template <int D>
__device__ __forceinline__ void gate1_vec2(const mpsb_gate1_desc& d, const cf* Ac, cf* Oc, const cf (&g)[D][D]) {
    const float4* __restrict__ A = reinterpret_cast<const float4*>(Ac);
    float4* __restrict__ O = reinterpret_cast<float4*>(Oc);
    const int half = d.chiR >> 1;                              // float4 = two consecutive r
    const int64_t n = (int64_t)d.chiL * half;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t l = e / half;
        const int r2 = (int)(e - l * half);
        const int64_t base = l * D * half + r2;
        float4 a[D];
#pragma unroll
        for (int p = 0; p < D; ++p) a[p] = A[base + (int64_t)p * half];
#pragma unroll
        for (int o = 0; o < D; ++o) {
            cf t0 = cf_make(0.f, 0.f), t1 = cf_make(0.f, 0.f);
#pragma unroll
            for (int p = 0; p < D; ++p) {
                t0 = cf_fma(g[o][p], cf_make(a[p].x, a[p].y), t0);
                t1 = cf_fma(g[o][p], cf_make(a[p].z, a[p].w), t1);
            }
            O[base + (int64_t)o * half] = make_float4(t0.x, t0.y, t1.x, t1.y);
        }
    }
}

template <int D>
__global__ void __launch_bounds__(256)
gate1_kernel(const mpsb_gate1_desc* __restrict__ descs, int nbatch) {
    const int job = blockIdx.y;
    const int di = job / nbatch, bi = job % nbatch;
    const mpsb_gate1_desc d = descs[di];
    const cf* __restrict__ A = (const cf*)d.site + (int64_t)bi * d.bs_site;
    cf* __restrict__ O = (cf*)d.out + (int64_t)bi * d.bs_out;
    const cf* __restrict__ G = (const cf*)d.gate + (int64_t)bi * d.bs_gate;
    cf g[D][D];
#pragma unroll
    for (int o = 0; o < D; ++o)
#pragma unroll
        for (int p = 0; p < D; ++p) g[o][p] = G[o * D + p];
    const int chiR = d.chiR;
    if ((chiR & 1) == 0 && ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(O)) & 15) == 0) {   // block-uniform
        gate1_vec2<D>(d, A, O, g);
        return;
    }
    const int64_t n = (int64_t)d.chiL * chiR;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
        int64_t l = e / chiR;
        int r = (int)(e - l * chiR);
        int64_t base = l * D * chiR + r;
        cf a[D];
#pragma unroll
        for (int p = 0; p < D; ++p) a[p] = A[base + (int64_t)p * chiR];
#pragma unroll
        for (int o = 0; o < D; ++o) {
            cf t = cf_make(0.f, 0.f);
#pragma unroll
            for (int p = 0; p < D; ++p) t = cf_fma(g[o][p], a[p], t);
            O[base + (int64_t)o * chiR] = t;
        }
    }
}

// site <- factor[b] * site     (mpsim/core.py:590-594)
__global__ void __launch_bounds__(256)
scale_kernel(const mpsb_site_ref* __restrict__ sites, int nbatch, int d, const float* __restrict__ factors) {
    const int job = blockIdx.y;
    const int si = job / nbatch, bi = job % nbatch;
    const mpsb_site_ref s = sites[si];
    cf* A = (cf*)s.site + (int64_t)bi * s.bs;
    const float f = factors[bi];
    const int64_t n = (int64_t)s.chiL * d * s.chiR;
    if ((n & 1) == 0 && (reinterpret_cast<uintptr_t>(A) & 15) == 0) {         // 16-byte accesses (block-uniform)
        float4* A4 = reinterpret_cast<float4*>(A);
        for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < (n >> 1); e += (int64_t)gridDim.x * blockDim.x) {
            float4 v = A4[e];
            A4[e] = make_float4(f * v.x, f * v.y, f * v.z, f * v.w);
        }
        return;
    }
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
        cf v = A[e];
        A[e] = cf_make(f * v.x, f * v.y);
    }
}

// <bits|psi> = A_0[:, b0, :] A_1[:, b1, :] ... : one CTA per (bitstring, batch member); the
// running row vector lives in shared memory, each step is a coalesced vector-matrix product.
__global__ void __launch_bounds__(256)
amplitude_kernel(const mpsb_site_ref* __restrict__ sites, int nsites, int nbatch, int d, int max_chi,
                 const uint8_t* __restrict__ bits, int nbits, cf* __restrict__ out) {
    extern __shared__ float4 smem_raw[];
    cf* v0 = (cf*)smem_raw;
    cf* v1 = v0 + max_chi;
    const int bit_id = blockIdx.x, bi = blockIdx.y;
    const uint8_t* mybits = bits + (size_t)bit_id * nsites;
    for (int i = threadIdx.x; i < max_chi; i += blockDim.x) v0[i] = cf_make(i == 0 ? 1.f : 0.f, 0.f);
    __syncthreads();
    cf* cur = v0; cf* nxt = v1;
    for (int s = 0; s < nsites; ++s) {
        const mpsb_site_ref sr = sites[s];
        const cf* A = (const cf*)sr.site + (int64_t)bi * sr.bs;
        const int b = mybits[s];
        const int chiL = sr.chiL, chiR = sr.chiR;
        for (int r = threadIdx.x; r < chiR; r += blockDim.x) {
            cf acc = cf_make(0.f, 0.f);
            for (int l = 0; l < chiL; ++l) acc = cf_fma(cur[l], A[((int64_t)l * d + b) * chiR + r], acc);
            nxt[r] = acc;
        }
        __syncthreads();
        cf* t = cur; cur = nxt; nxt = t;
    }
    if (threadIdx.x == 0) out[(size_t)bi * nbits + bit_id] = cur[0];
}

// ---- gauge rebalance -------------------------------------------------------------------------
// The reference's sweep pattern (even layers left-canonical, odd layers right-canonical,
// mpsim/core.py:1348-1360) leaves the SCALE of the individual site tensors free: only the product
// is fixed, and in a long circuit it migrates geometrically (in a 12-qubit brickwork one site
// reaches 1e-37 and two others 1e18 after 240 layers -- the complex128 reference shows the same
// numbers and has the exponent range for it, complex64 storage does not).  These three kernels
// move powers of two between the sites of one chain so that every site's largest entry has about
// the same exponent.  The shifts of a chain sum to zero and a power-of-two scaling is exact, so the
// state (norm, amplitudes, every contraction) is unchanged bit for bit.

// pass 1: bit pattern of max |re|, |im| over the finite entries of each (site, member)
__global__ void __launch_bounds__(256)
site_maxabs_kernel(const mpsb_site_ref* __restrict__ sites, int nbatch, int d, unsigned* __restrict__ maxbits) {
    const int job = blockIdx.x;
    const int si = job / nbatch, bi = job % nbatch;
    const mpsb_site_ref s = sites[si];
    const float* A = (const float*)((const cf*)s.site + (int64_t)bi * s.bs);
    const int64_t n = 2 * (int64_t)s.chiL * d * s.chiR;
    unsigned mx = 0;
    for (int64_t e = (int64_t)blockIdx.y * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.y * blockDim.x) {
        unsigned b = __float_as_uint(A[e]) & 0x7fffffffu;
        if (b < 0x7f800000u && b > mx) mx = b;
    }
    for (int o = 16; o; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0 && mx) atomicMax(&maxbits[job], mx);
}

// pass 2: one thread per member turns the maxima into shifts (in place, as int).  Nothing moves
// while the exponents of a chain are within `spread` of each other.
__global__ void rebalance_shift_kernel(unsigned* __restrict__ maxbits, int nsites, int nbatch, int spread) {
    const int bi = blockIdx.x * blockDim.x + threadIdx.x;
    if (bi >= nbatch) return;
    int lo = INT_MAX, hi = INT_MIN, cnt = 0;
    long long sum = 0;
    for (int s = 0; s < nsites; ++s) {
        unsigned b = maxbits[(size_t)s * nbatch + bi];
        if (!b) continue;
        int e = ilogbf(__uint_as_float(b));
        lo = min(lo, e); hi = max(hi, e); sum += e; ++cnt;
    }
    const bool act = cnt > 1 && hi - lo > spread;
    // every site to exponent `mean`, the first `rem` of them to mean + 1: the shifts add up to
    // cnt * mean + rem - sum = 0 exactly
    long long mean = 0, rem = 0;
    if (act) {
        mean = sum / cnt;
        if (sum - mean * cnt < 0) --mean;                   // floor division
        rem = sum - mean * cnt;                             // 0 .. cnt - 1
    }
    int seen = 0;
    for (int s = 0; s < nsites; ++s) {
        unsigned b = maxbits[(size_t)s * nbatch + bi];
        int sh = 0;
        if (act && b) {
            sh = (int)mean - ilogbf(__uint_as_float(b)) + (seen < rem ? 1 : 0);
            ++seen;
        }
        ((int*)maxbits)[(size_t)s * nbatch + bi] = sh;
    }
}

// pass 3: site <- 2^shift * site (CTAs of an unshifted site leave at once)
__global__ void __launch_bounds__(256)
rebalance_apply_kernel(const mpsb_site_ref* __restrict__ sites, int nbatch, int d, const int* __restrict__ shifts) {
    const int job = blockIdx.x;
    const int sh = shifts[job];
    if (sh == 0) return;
    const int si = job / nbatch, bi = job % nbatch;
    const mpsb_site_ref s = sites[si];
    float* A = (float*)((cf*)s.site + (int64_t)bi * s.bs);
    const int64_t n = 2 * (int64_t)s.chiL * d * s.chiR;
    for (int64_t e = (int64_t)blockIdx.y * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.y * blockDim.x)
        A[e] = scalbnf(A[e], sh);
}

}  // namespace

int launch_gate1(const mpsb_gate1_desc* descs, int ndesc, int nbatch, int d, int max_site_elems, cudaStream_t st) {
    int njobs = ndesc * nbatch;
    if (njobs <= 0) return 0;
    MPSB_ARG(njobs <= 65535, "gate1: too many jobs in one call (%d > 65535)", njobs);
    int per = max_site_elems / d;
    int bx = (per + 255) / 256;
    if (bx < 1) bx = 1;
    if (bx > 1184) bx = 1184;      // 8 x 148 SMs; the kernel grid-strides beyond that
    dim3 grid(bx, njobs);
    switch (d) {
        case 2: gate1_kernel<2><<<grid, 256, 0, st>>>(descs, nbatch); break;
        case 3: gate1_kernel<3><<<grid, 256, 0, st>>>(descs, nbatch); break;
        case 4: gate1_kernel<4><<<grid, 256, 0, st>>>(descs, nbatch); break;
        case 5: gate1_kernel<5><<<grid, 256, 0, st>>>(descs, nbatch); break;
        default: MPSB_ARG(false, "gate1: qudit dimension %d not supported on device (2..5)", d);
    }
    MPSB_LAUNCH_CHECK("gate1_kernel");
    return 0;
}

int launch_scale(const mpsb_site_ref* sites, int nsites, int nbatch, int d, const float* factors,
                 int max_site_elems, cudaStream_t st) {
    int njobs = nsites * nbatch;
    if (njobs <= 0) return 0;
    MPSB_ARG(njobs <= 65535, "scale: too many jobs in one call (%d > 65535)", njobs);
    int bx = (max_site_elems + 255) / 256;
    if (bx < 1) bx = 1;
    if (bx > 1184) bx = 1184;
    scale_kernel<<<dim3(bx, njobs), 256, 0, st>>>(sites, nbatch, d, factors);
    MPSB_LAUNCH_CHECK("scale_kernel");
    return 0;
}

int launch_rebalance(const mpsb_site_ref* sites, int nsites, int nbatch, int d, int spread, int* shifts,
                     int max_site_elems, cudaStream_t st) {
    const long long njobs = (long long)nsites * nbatch;
    if (njobs <= 0) return 0;
    MPSB_ARG(njobs < (1ll << 31), "rebalance: too many (site, member) pairs (%lld)", njobs);
    int by = (2 * max_site_elems + 2047) / 2048;          // 8 values per thread
    if (by < 1) by = 1;
    if (by > 64) by = 64;
    MPSB_CUDA(cudaMemsetAsync(shifts, 0, (size_t)njobs * sizeof(int), st));
    site_maxabs_kernel<<<dim3((unsigned)njobs, by), 256, 0, st>>>(sites, nbatch, d, (unsigned*)shifts);
    MPSB_LAUNCH_CHECK("site_maxabs_kernel");
    rebalance_shift_kernel<<<(nbatch + 127) / 128, 128, 0, st>>>((unsigned*)shifts, nsites, nbatch, spread);
    MPSB_LAUNCH_CHECK("rebalance_shift_kernel");
    rebalance_apply_kernel<<<dim3((unsigned)njobs, by), 256, 0, st>>>(sites, nbatch, d, shifts);
    MPSB_LAUNCH_CHECK("rebalance_apply_kernel");
    return 0;
}

int launch_amplitudes(const mpsb_site_ref* sites, int nsites, int nbatch, int d, int max_chi,
                      const uint8_t* bits, int nbits, cf* out, cudaStream_t st) {
    if (nbits <= 0 || nbatch <= 0) return 0;
    MPSB_ARG(nbatch <= 65535, "amplitudes: nbatch %d > 65535", nbatch);
    if (max_chi < 1) max_chi = 1;
    size_t smem = (size_t)2 * max_chi * sizeof(cf);
    MPSB_ARG(smem <= 200 * 1024, "amplitudes: max_chi %d too large", max_chi);
    if (smem > 48 * 1024)
        MPSB_CUDA(cudaFuncSetAttribute(amplitude_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    amplitude_kernel<<<dim3(nbits, nbatch), 256, smem, st>>>(sites, nsites, nbatch, d, max_chi, bits, nbits, out);
    MPSB_LAUNCH_CHECK("amplitude_kernel");
    return 0;
}
