// extern "C" surface of libmpsim_b200 (see include/mpsim_b200.h for the contract).
#include "common.cuh"
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>
#include <mutex>
#include <map>
#include <utility>

int launch_gate1(const mpsb_gate1_desc* descs, int ndesc, int nbatch, int d, int max_site_elems, cudaStream_t st);
int launch_scale(const mpsb_site_ref* sites, int nsites, int nbatch, int d, const float* factors,
                 int max_site_elems, cudaStream_t st);
int launch_amplitudes(const mpsb_site_ref* sites, int nsites, int nbatch, int d, int max_chi,
                      const uint8_t* bits, int nbits, cf* out, cudaStream_t st);
int launch_rebalance(const mpsb_site_ref* sites, int nsites, int nbatch, int d, int spread, int* shifts,
                     int max_site_elems, cudaStream_t st);

static thread_local char g_err[512] = "";

void mpsb_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

// Developer switches (MPSB_* environment variables; not part of the ABI): each is read ONCE, at its
// first use, and cached -- no getenv on the call path.  MPSB_DEV_REREAD_ENV=1 (itself read once)
// turns the cache off for tests and A/B scripts that flip a switch between calls.
const char* mpsb_env(const char* name) {
    static const bool reread = [] { const char* e = getenv("MPSB_DEV_REREAD_ENV"); return e && atoi(e) != 0; }();
    if (reread) return getenv(name);
    static std::mutex mu;
    static std::map<std::string, std::pair<bool, std::string>> cache;
    std::lock_guard<std::mutex> lock(mu);
    auto it = cache.find(name);
    if (it == cache.end()) {
        const char* e = getenv(name);
        it = cache.emplace(name, std::make_pair(e != nullptr, std::string(e ? e : ""))).first;
    }
    return it->second.first ? it->second.second.c_str() : nullptr;
}

std::recursive_mutex& mpsb_lib_mutex() {
    static std::recursive_mutex mu;
    return mu;
}

int mpsb_current_device(int* dev) {
    MPSB_CUDA(cudaGetDevice(dev));
    MPSB_ARG(*dev >= 0 && *dev < MPSB_MAX_DEVICES, "device ordinal %d outside [0, %d)", *dev, MPSB_MAX_DEVICES);
    return 0;
}

namespace {

__global__ void set_ones_kernel(cf* E, int64_t stride, int nbatch) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < nbatch) E[(int64_t)b * stride] = cf_make(1.f, 0.f);
}

__global__ void gather_first_kernel(const cf* E, int64_t stride, int nbatch, cf* out) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < nbatch) out[b] = E[(int64_t)b * stride];
}

// out[job][c][r] = in[job][r][c]   (in: rows x cols)
__global__ void transpose_kernel(const cf* __restrict__ in, cf* __restrict__ out, int rows, int cols) {
    __shared__ cf tile[32][33];
    const cf* src = in + (int64_t)blockIdx.z * rows * cols;
    cf* dst = out + (int64_t)blockIdx.z * rows * cols;
    int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        int r = r0 + i, c = c0 + threadIdx.x;
        if (r < rows && c < cols) tile[i][threadIdx.x] = src[(int64_t)r * cols + c];
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        int c = c0 + i, r = r0 + threadIdx.x;
        if (r < rows && c < cols) dst[(int64_t)c * rows + r] = tile[threadIdx.x][i];
    }
}

struct Gate2Plan {
    int m, n, nv, L;
    bool small;
    size_t x_elems, extra_elems, job_elems;
};

// theta runs on the tensor cores (tc_gemm.cu) for qubits once the tile is reasonably filled;
// MPSB_THETA_TC=0/1 forces the choice (debugging / A-B timing)
bool theta_uses_tc(int d, int chiL, int chiM, int chiR) {
    if (const char* e = mpsb_env("MPSB_THETA_TC")) return atoi(e) != 0 && d == 2;
    return d == 2 && chiL >= 32 && chiR >= 32 && chiM >= 16;
}

Gate2Plan plan_gate2(int d, int chiL, int chiR, int lc) {
    Gate2Plan p;
    p.m = d * chiL; p.n = d * chiR;
    p.nv = lc ? p.m : p.n;
    p.L = lc ? p.n : p.m;
    p.small = p.m <= MPSB_MAX_SMALL_DIM && p.n <= MPSB_MAX_SMALL_DIM;
    p.x_elems = p.small ? (size_t)p.m * p.n : (size_t)svd_large_padded_rows(p.nv) * p.L;
    p.extra_elems = p.small ? (size_t)0 : svd_large_workspace_elems(p.nv, p.L);
    p.job_elems = align_up(p.x_elems, 16) + align_up(p.extra_elems, 16);
    return p;
}

}  // namespace

extern "C" {

int mpsb_version(void) { return MPSB_VERSION; }
const char* mpsb_last_error(void) { return g_err; }

int mpsb_device_info(int* sm_count, int* max_smem_optin, int* cc_major, int* cc_minor) {
    int dev = 0;
    MPSB_CUDA(cudaGetDevice(&dev));
    if (sm_count) MPSB_CUDA(cudaDeviceGetAttribute(sm_count, cudaDevAttrMultiProcessorCount, dev));
    if (max_smem_optin) MPSB_CUDA(cudaDeviceGetAttribute(max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    if (cc_major) MPSB_CUDA(cudaDeviceGetAttribute(cc_major, cudaDevAttrComputeCapabilityMajor, dev));
    if (cc_minor) MPSB_CUDA(cudaDeviceGetAttribute(cc_minor, cudaDevAttrComputeCapabilityMinor, dev));
    return 0;
}

size_t mpsb_gate2_workspace_bytes(int ndesc, int nbatch, int d, int chiL, int chiM, int chiR, int k) {
    (void)chiM; (void)k;
    if (ndesc <= 0 || nbatch <= 0 || chiL <= 0 || chiR <= 0) return 0;
    // worst of the two orientations so one workspace serves both canonical forms
    size_t a = plan_gate2(d, chiL, chiR, 1).job_elems, b = plan_gate2(d, chiL, chiR, 0).job_elems;
    size_t tc = d == 2 ? align_up(tc_theta_workspace_floats(ndesc * nbatch, chiL, chiM, chiR) * sizeof(float), 256) : 0;
    return (a > b ? a : b) * sizeof(cf) * (size_t)ndesc * nbatch + 256 + tc;
}

// theta of one shape group onto `st`; returns the plan and where X / the SVD scratch live
static int gate2_theta(const mpsb_gate2_desc* descs_dev, int ndesc, int nbatch, int d, int chiL, int chiM, int chiR,
                       int k, int left_canonical, void* workspace, size_t workspace_bytes, cudaStream_t st,
                       Gate2Plan& p, cf*& X, cf*& extra) {
    MPSB_ARG(descs_dev != nullptr, "apply_gate2: descs is NULL");
    MPSB_ARG(ndesc >= 0 && nbatch >= 0, "apply_gate2: negative counts");
    MPSB_ARG(d >= 2, "apply_gate2: qudit dimension %d < 2", d);
    MPSB_ARG(chiL >= 0 && chiM >= 0 && chiR >= 0, "apply_gate2: negative bond dimension");
    int njobs = ndesc * nbatch;
    int mn = d * (chiL < chiR ? chiL : chiR);
    // k == 0 is legal (maxsvals=0, core_test.py:1093-1101): the sites become empty but the
    // singular values are still computed and reported
    MPSB_ARG(k >= 0 && k <= mn, "apply_gate2: k=%d outside [0, %d]", k, mn);
    MPSB_ARG(njobs <= 65535, "apply_gate2: %d applications in one call (max 65535); split the call", njobs);
    p = plan_gate2(d, chiL, chiR, left_canonical ? 1 : 0);
    const bool tc = theta_uses_tc(d, chiL, chiM, chiR);
    size_t need_svd = align_up(p.job_elems * sizeof(cf) * (size_t)njobs, 256);
    size_t need = need_svd + (tc ? tc_theta_workspace_floats(njobs, chiL, chiM, chiR) * sizeof(float) : 0);
    MPSB_ARG(workspace != nullptr && workspace_bytes >= need, "apply_gate2: workspace %zu B < %zu B", workspace_bytes, need);
    MPSB_ARG(((uintptr_t)workspace & 255) == 0, "apply_gate2: workspace must be 256-byte aligned");
    cf* ws = (cf*)workspace;
    X = ws;                                                  // [njobs][x_elems]
    extra = ws + align_up(p.x_elems, 16) * (size_t)njobs;    // [njobs][extra]
    if (tc)
        return launch_theta_tc(descs_dev, ndesc, nbatch, chiL, chiM, chiR, left_canonical ? 0 : 1, X,
                               (int64_t)align_up(p.x_elems, 16), (float*)((char*)workspace + need_svd), st);
    return launch_theta(descs_dev, ndesc, nbatch, d, chiL, chiM, chiR, left_canonical ? 0 : 1,
                        X, (int64_t)align_up(p.x_elems, 16), st);
}

int mpsb_apply_gate2(const mpsb_gate2_desc* descs_dev, int ndesc, int nbatch,
                     int d, int chiL, int chiM, int chiR, int k, int left_canonical,
                     void* workspace, size_t workspace_bytes, int32_t* info, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    MPSB_ARG(ndesc >= 0 && nbatch >= 0, "apply_gate2: negative counts");
    if ((long long)ndesc * nbatch > 65535 && nbatch <= 65535) {
        // the kernels index (descriptor, member) jobs through one 16-bit grid dimension: a longer call
        // is run as consecutive chunks of descriptors on the same stream and workspace
        const int per = 65535 / nbatch;
        for (int off = 0; off < ndesc; off += per) {
            const int c = ndesc - off < per ? ndesc - off : per;
            int rc = mpsb_apply_gate2(descs_dev + off, c, nbatch, d, chiL, chiM, chiR, k, left_canonical, workspace,
                                      workspace_bytes, info ? info + (size_t)off * nbatch * 2 : nullptr, stream);
            if (rc) return rc;
        }
        return 0;
    }
    int njobs = ndesc * nbatch;
    if (njobs == 0 || chiL == 0 || chiR == 0) return 0;   // empty tensors: nothing to compute
    Gate2Plan p; cf *X, *extra;
    int rc = gate2_theta(descs_dev, ndesc, nbatch, d, chiL, chiM, chiR, k, left_canonical, workspace, workspace_bytes, st,
                         p, X, extra);
    if (rc) return rc;
    if (p.small) {
        return launch_svd_small(X, (int64_t)align_up(p.x_elems, 16), njobs, p.nv, p.L, k, left_canonical ? 1 : 0,
                                descs_dev, ndesc, nbatch, nullptr, 0, nullptr, 0, nullptr, 0, info,
                                p.extra_elems ? extra : nullptr, st);
    }
    return launch_svd_large(X, (int64_t)align_up(p.x_elems, 16), njobs, p.nv, p.L, k, left_canonical ? 1 : 0,
                            descs_dev, ndesc, nbatch, nullptr, 0, nullptr, 0, nullptr, 0, info, extra, st);
}

size_t mpsb_gate2_layer_workspace_bytes(const mpsb_gate2_group* g, int ngroups, int nbatch, int d) {
    size_t tot = 0;
    for (int i = 0; i < ngroups; ++i)
        tot += align_up(mpsb_gate2_workspace_bytes(g[i].ndesc, nbatch, d, g[i].chiL, g[i].chiM, g[i].chiR, g[i].k), 256);
    return tot;
}

// library-owned streams of mpsb_apply_gate2_layer: one pool per device ordinal, created on first use
// under the library mutex, never destroyed.  Calls that use library-owned state (this pool, the pinned
// read-back slots of svd_large.cu) serialise on that mutex: ctypes drops the GIL, so two host threads
// may be inside the library at once.
static const int kPoolStreams = 8;
struct StreamPool {
    cudaStream_t s[kPoolStreams];
    cudaEvent_t fork = nullptr, join[kPoolStreams];
    bool ready = false;
};
static StreamPool g_pools[MPSB_MAX_DEVICES];


int mpsb_apply_gate2_layer(const mpsb_gate2_group* groups, int ngroups, int nbatch, int d,
                           void* workspace, size_t workspace_bytes, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    MPSB_ARG(groups != nullptr || ngroups == 0, "apply_gate2_layer: groups is NULL");
    MPSB_ARG(ngroups >= 0 && ngroups <= 4096, "apply_gate2_layer: %d groups", ngroups);
    if (ngroups == 0 || nbatch == 0) return 0;
    MPSB_ARG(workspace_bytes >= mpsb_gate2_layer_workspace_bytes(groups, ngroups, nbatch, d),
             "apply_gate2_layer: workspace too small");
    std::lock_guard<std::recursive_mutex> lock(mpsb_lib_mutex());
    int dev = 0;
    if (int rc = mpsb_current_device(&dev)) return rc;
    StreamPool& pool = g_pools[dev];
    if (!pool.ready) {
        for (int i = 0; i < kPoolStreams; ++i) {
            MPSB_CUDA(cudaStreamCreateWithFlags(&pool.s[i], cudaStreamNonBlocking));
            MPSB_CUDA(cudaEventCreateWithFlags(&pool.join[i], cudaEventDisableTiming));
        }
        MPSB_CUDA(cudaEventCreateWithFlags(&pool.fork, cudaEventDisableTiming));
        pool.ready = true;
    }
    cudaStream_t* g_pool = pool.s;
    cudaEvent_t g_fork = pool.fork;
    cudaEvent_t* g_join = pool.join;
    // heaviest group first: it stays on the caller's stream, the others go round the pool
    std::vector<int> order(ngroups);
    std::vector<double> cost(ngroups);
    std::vector<size_t> offs(ngroups);
    size_t off = 0;
    for (int i = 0; i < ngroups; ++i) {
        order[i] = i;
        double mx = (double)d * (groups[i].chiL > groups[i].chiR ? groups[i].chiL : groups[i].chiR);
        cost[i] = (double)groups[i].ndesc * nbatch * mx * mx * mx;
        offs[i] = off;
        off += align_up(mpsb_gate2_workspace_bytes(groups[i].ndesc, nbatch, d, groups[i].chiL, groups[i].chiM,
                                                   groups[i].chiR, groups[i].k), 256);
    }
    for (int a = 1; a < ngroups; ++a)                        // insertion sort, descending cost
        for (int b = a; b > 0 && cost[order[b]] > cost[order[b - 1]]; --b) { int t = order[b]; order[b] = order[b - 1]; order[b - 1] = t; }
    if (ngroups > 1) {
        MPSB_CUDA(cudaEventRecord(g_fork, st));
        for (int i = 0; i < kPoolStreams && i < ngroups - 1; ++i) MPSB_CUDA(cudaStreamWaitEvent(g_pool[i], g_fork, 0));
    }
    std::vector<LargeMultiJob> large;
    int rc = 0;
    auto flush_large = [&]() -> int {
        if (large.empty()) return 0;
        int r = launch_svd_large_multi(large.data(), (int)large.size());
        large.clear();
        return r;
    };
    for (int oi = 0; oi < ngroups && rc == 0; ++oi) {
        const mpsb_gate2_group& g = groups[order[oi]];
        const int njobs = g.ndesc * nbatch;
        if (njobs == 0 || g.chiL == 0 || g.chiR == 0) continue;
        const int slot = oi == 0 ? kPoolStreams : (oi - 1) % kPoolStreams;     // pin slot == stream index
        cudaStream_t gs = oi == 0 ? st : g_pool[slot];
        char* ws = (char*)workspace + offs[order[oi]];
        size_t wsb = mpsb_gate2_workspace_bytes(g.ndesc, nbatch, d, g.chiL, g.chiM, g.chiR, g.k);
        Gate2Plan p; cf *X, *extra;
        rc = gate2_theta(g.descs_dev, g.ndesc, nbatch, d, g.chiL, g.chiM, g.chiR, g.k, g.left_canonical, ws, wsb, gs, p, X, extra);
        if (rc) break;
        if (p.small) {
            rc = launch_svd_small(X, (int64_t)align_up(p.x_elems, 16), njobs, p.nv, p.L, g.k, g.left_canonical ? 1 : 0,
                                  g.descs_dev, g.ndesc, nbatch, nullptr, 0, nullptr, 0, nullptr, 0, g.info,
                                  p.extra_elems ? extra : nullptr, gs);
            continue;
        }
        // a stream (and its pinned read-back slot) carries one large solve at a time
        for (const LargeMultiJob& j : large)
            if (j.pin_slot == slot) { rc = flush_large(); break; }
        if (rc) break;
        LargeMultiJob j;
        j.X = X; j.x_job_stride = (int64_t)align_up(p.x_elems, 16); j.njobs = njobs; j.nv = p.nv; j.L = p.L; j.k = g.k;
        j.left_canonical = g.left_canonical ? 1 : 0; j.descs = g.descs_dev; j.nbatch = nbatch; j.info = g.info;
        j.work = extra; j.st = gs; j.pin_slot = slot;
        large.push_back(j);
    }
    if (rc == 0) rc = flush_large();
    if (ngroups > 1) {                                       // join even after an error: never leave the pool detached
        for (int i = 0; i < kPoolStreams && i < ngroups - 1; ++i) {
            cudaEventRecord(g_join[i], g_pool[i]);
            cudaStreamWaitEvent(st, g_join[i], 0);
        }
    }
    return rc;
}

int mpsb_apply_gate1(const mpsb_gate1_desc* descs_dev, int ndesc, int nbatch, int d,
                     int max_site_elems, void* stream) {
    MPSB_ARG(descs_dev != nullptr || ndesc == 0, "apply_gate1: descs is NULL");
    return launch_gate1(descs_dev, ndesc, nbatch, d, max_site_elems, (cudaStream_t)stream);
}

size_t mpsb_inner_workspace_bytes(int nbatch, int d, int max_chi_a, int max_chi_b) {
    if (max_chi_a < 1) max_chi_a = 1;
    if (max_chi_b < 1) max_chi_b = 1;
    size_t e = align_up((size_t)max_chi_a * max_chi_b, 16);
    size_t t = align_up((size_t)max_chi_b * d * max_chi_a, 16);
    return (2 * e + t) * sizeof(cf) * (size_t)nbatch + 256;
}

int mpsb_inner_products(const mpsb_site_ref* a, const mpsb_site_ref* b, int nsites,
                        int nbatch, int d, void* out, void* workspace, size_t workspace_bytes,
                        void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    MPSB_ARG(a && b && nsites >= 1 && nbatch >= 1, "inner_products: bad arguments");
    int ma = 1, mb = 1;
    for (int s = 0; s < nsites; ++s) {
        MPSB_ARG(a[s].chiL >= 0 && a[s].chiR >= 0 && b[s].chiL >= 0 && b[s].chiR >= 0, "inner_products: negative bond");
        if (s > 0) MPSB_ARG(a[s].chiL == a[s - 1].chiR && b[s].chiL == b[s - 1].chiR, "inner_products: bond mismatch at site %d", s);
        ma = a[s].chiL > ma ? a[s].chiL : ma; ma = a[s].chiR > ma ? a[s].chiR : ma;
        mb = b[s].chiL > mb ? b[s].chiL : mb; mb = b[s].chiR > mb ? b[s].chiR : mb;
    }
    MPSB_ARG(a[0].chiL == 1 && b[0].chiL == 1 && a[nsites - 1].chiR == 1 && b[nsites - 1].chiR == 1,
             "inner_products: chain ends must have bond dimension 1");
    MPSB_ARG(workspace_bytes >= mpsb_inner_workspace_bytes(nbatch, d, ma, mb), "inner_products: workspace too small");
    size_t es = align_up((size_t)ma * mb, 16), ts = align_up((size_t)mb * d * ma, 16);
    cf* E0 = (cf*)workspace;
    cf* E1 = E0 + es * nbatch;
    cf* T = E1 + es * nbatch;
    MPSB_CUDA(cudaMemsetAsync(E0, 0, es * nbatch * sizeof(cf), st));
    set_ones_kernel<<<(nbatch + 127) / 128, 128, 0, st>>>(E0, (int64_t)es, nbatch);
    MPSB_LAUNCH_CHECK("set_ones_kernel");
    cf* Ec = E0; cf* En = E1;
    for (int s = 0; s < nsites; ++s) {
        int xa = a[s].chiL, xa2 = a[s].chiR, yb = b[s].chiL, yb2 = b[s].chiR;
        // T[y][(p,x')] = sum_x E[x][y] A[x][(p,x')]
        int rc = launch_cgemm(Ec, 1, yb, 0, (int64_t)es,
                              (const cf*)a[s].site, (int64_t)d * xa2, 1, 0, a[s].bs,
                              T, (int64_t)d * xa2, (int64_t)ts, yb, d * xa2, xa, nbatch, st);
        if (rc) return rc;
        // E'[x'][y'] = sum_{(y,p)} T[(y,p)][x'] conj(B[(y,p)][y'])
        rc = launch_cgemm(T, 1, xa2, 0, (int64_t)ts,
                          (const cf*)b[s].site, yb2, 1, 1, b[s].bs,
                          En, yb2, (int64_t)es, xa2, yb2, yb * d, nbatch, st);
        if (rc) return rc;
        cf* t = Ec; Ec = En; En = t;
    }
    gather_first_kernel<<<(nbatch + 127) / 128, 128, 0, st>>>(Ec, (int64_t)es, nbatch, (cf*)out);
    MPSB_LAUNCH_CHECK("gather_first_kernel");
    return 0;
}

int mpsb_scale_sites(const mpsb_site_ref* sites_dev, int nsites, int nbatch, int d,
                     const float* factors_dev, int max_site_elems, void* stream) {
    MPSB_ARG(sites_dev && factors_dev, "scale_sites: NULL argument");
    return launch_scale(sites_dev, nsites, nbatch, d, factors_dev, max_site_elems, (cudaStream_t)stream);
}

int mpsb_rebalance_sites(const mpsb_site_ref* sites_dev, int nsites, int nbatch, int d, int spread_log2,
                         int* shifts_dev, int max_site_elems, void* stream) {
    MPSB_ARG(sites_dev && shifts_dev, "rebalance_sites: NULL argument");
    MPSB_ARG(spread_log2 >= 0, "rebalance_sites: negative spread");
    return launch_rebalance(sites_dev, nsites, nbatch, d, spread_log2, shifts_dev, max_site_elems, (cudaStream_t)stream);
}

// Dense wavefunction of one chain: contracted from BOTH ends and joined by one product,
//   L [d^sp][chi_sp]   = A_0 A_1 ... A_{sp-1}          (left chain, rows = leading digits)
//   R [chi_sp][d^(n-sp)] = A_sp ... A_{n-1}            (right chain, columns = trailing digits)
//   psi [d^sp][d^(n-sp)] = L . R                        (big-endian index, mpsim/core.py:483-500)
// A one-sided chain materialises d^(i+1) x chi_{i+1} partial products all the way to the last site --
// 16 x the output at n = 24, chi = 64 (the right end of a chain has bonds 64, 32, ..., 2, 1 while the row
// count keeps doubling) -- and ran at 25 GB/s of output (bench.py, round 2); split at the site that
// minimises the bytes of all partial products, the products are ~2 x the output and the join is one GEMM
// with K = chi_sp (on the tcgen05 kernel when the tile fills).
struct WfPlan { int sp; size_t maxL, maxR, tc_floats; bool tc; };

static WfPlan wf_plan(const mpsb_site_ref* s, int n, int d) {
    std::vector<double> Ls(n), Rs(n);          // elements of L after site i, of R from site i
    double rows = 1;
    for (int i = 0; i < n; ++i) { rows *= d; Ls[i] = rows * (s[i].chiR > 0 ? s[i].chiR : 1); }
    double cols = 1;
    for (int i = n - 1; i >= 0; --i) { cols *= d; Rs[i] = cols * (s[i].chiL > 0 ? s[i].chiL : 1); }
    WfPlan p; p.sp = 1;
    double best = -1;
    for (int sp = 1; sp <= n - 1; ++sp) {
        double tot = 0;
        for (int i = 1; i < sp; ++i) tot += Ls[i];           // site 0 is used in place
        for (int i = sp; i < n - 1; ++i) tot += Rs[i];       // the last site is used in place
        if (best < 0 || tot < best) { best = tot; p.sp = sp; }
    }
    p.maxL = 1; p.maxR = 1;
    for (int i = 1; i < p.sp; ++i) if ((size_t)Ls[i] > p.maxL) p.maxL = (size_t)Ls[i];
    for (int i = p.sp; i < n - 1; ++i) if ((size_t)Rs[i] > p.maxR) p.maxR = (size_t)Rs[i];
    double M = 1, N = 1;
    for (int i = 0; i < p.sp; ++i) M *= d;
    for (int i = p.sp; i < n; ++i) N *= d;
    const int K = s[p.sp].chiL;
    p.tc = K >= 16 && M >= 256 && N >= 128 && M <= 0x7fffffff && N <= 0x7fffffff;
    p.tc_floats = p.tc ? tc_cgemm_workspace_floats(1, (int)M, (int)N, K) : 0;
    return p;
}

size_t mpsb_wavefunction_workspace_bytes(const mpsb_site_ref* s, int nsites, int d) {
    if (!s || nsites < 2) return 256;
    const WfPlan p = wf_plan(s, nsites, d);
    return (2 * align_up(p.maxL, 16) + 2 * align_up(p.maxR, 16)) * sizeof(cf) + align_up(p.tc_floats * sizeof(float), 256) + 512;
}

int mpsb_wavefunction(const mpsb_site_ref* s, int nsites, int d, int batch_index,
                      void* out, void* workspace, size_t workspace_bytes, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    MPSB_ARG(s && nsites >= 2 && out, "wavefunction: bad arguments");
    MPSB_ARG(s[0].chiL == 1 && s[nsites - 1].chiR == 1, "wavefunction: chain ends must have bond dimension 1");
    MPSB_ARG(workspace_bytes >= mpsb_wavefunction_workspace_bytes(s, nsites, d), "wavefunction: workspace too small");
    MPSB_ARG(((uintptr_t)workspace & 255) == 0, "wavefunction: workspace must be 256-byte aligned");
    size_t total = 1;
    bool empty = false;
    for (int i = 0; i < nsites; ++i) {
        total *= (size_t)d;
        if (i > 0) MPSB_ARG(s[i].chiL == s[i - 1].chiR, "wavefunction: bond mismatch at site %d", i);
        if (s[i].chiR == 0) empty = true;
    }
    if (empty) {     // a bond of dimension 0 (maxsvals=0, core_test.py:1093-1101): psi = 0
        MPSB_CUDA(cudaMemsetAsync(out, 0, total * sizeof(cf), st));
        return 0;
    }
    const WfPlan p = wf_plan(s, nsites, d);
    cf* w = (cf*)workspace;
    cf* Lb[2] = {w, w + align_up(p.maxL, 16)};
    w += 2 * align_up(p.maxL, 16);
    cf* Rb[2] = {w, w + align_up(p.maxR, 16)};
    w += 2 * align_up(p.maxR, 16);
    float* tcw = (float*)(((uintptr_t)w + 255) & ~(uintptr_t)255);
    auto site = [&](int i) { return (const cf*)s[i].site + (int64_t)batch_index * s[i].bs; };
    // left chain: L_i [d^(i+1)][chi_{i+1}] = L_{i-1} [d^i][chi_i] . A_i [chi_i][d chi_{i+1}]
    const cf* L = site(0);
    size_t rows = d;
    for (int i = 1; i < p.sp; ++i) {
        MPSB_ARG(rows <= 0x7fffffff, "wavefunction: too many qudits");
        cf* dst = Lb[i & 1];
        int rc = launch_cgemm(L, s[i].chiL, 1, 0, 0, site(i), (int64_t)d * s[i].chiR, 1, 0, 0,
                              dst, (int64_t)d * s[i].chiR, 0, (int)rows, d * s[i].chiR, s[i].chiL, 1, st);
        if (rc) return rc;
        L = dst;
        rows *= (size_t)d;
    }
    // right chain: R_i [chi_i][d cols] = A_i [(chi_i, d)][chi_{i+1}] . R_{i+1} [chi_{i+1}][cols]
    const cf* R = site(nsites - 1);
    size_t cols = d;
    for (int i = nsites - 2; i >= p.sp; --i) {
        MPSB_ARG(cols <= 0x7fffffff, "wavefunction: too many qudits");
        cf* dst = Rb[i & 1];
        int rc = launch_cgemm(site(i), s[i].chiR, 1, 0, 0, R, (int64_t)cols, 1, 0, 0,
                              dst, (int64_t)cols, 0, s[i].chiL * d, (int)cols, s[i].chiR, 1, st);
        if (rc) return rc;
        R = dst;
        cols *= (size_t)d;
    }
    MPSB_ARG(rows <= 0x7fffffff && cols <= 0x7fffffff, "wavefunction: too many qudits");
    const int K = s[p.sp].chiL;
    if (p.tc)
        return launch_cgemm_tc(L, 0, R, 0, (cf*)out, (int64_t)cols, 0, (int)rows, (int)cols, K, 1, tcw, st);
    return launch_cgemm(L, K, 1, 0, 0, R, (int64_t)cols, 1, 0, 0, (cf*)out, (int64_t)cols, 0, (int)rows, (int)cols, K, 1, st);
}

int mpsb_amplitudes(const mpsb_site_ref* sites_dev, int nsites, int nbatch, int d, int max_chi,
                    const uint8_t* bits_dev, int nbits, void* out, void* stream) {
    MPSB_ARG(sites_dev && bits_dev && out, "amplitudes: NULL argument");
    return launch_amplitudes(sites_dev, nsites, nbatch, d, max_chi, bits_dev, nbits, (cf*)out, (cudaStream_t)stream);
}

int mpsb_cgemm(const void* A, int64_t a_rs, int64_t a_cs, int conj_a, int64_t a_bs,
               const void* B, int64_t b_rs, int64_t b_cs, int conj_b, int64_t b_bs,
               void* C, int64_t c_ld, int64_t c_bs, int M, int N, int K, int nbatch, void* stream) {
    return launch_cgemm((const cf*)A, a_rs, a_cs, conj_a, a_bs, (const cf*)B, b_rs, b_cs, conj_b, b_bs,
                        (cf*)C, c_ld, c_bs, M, N, K, nbatch, (cudaStream_t)stream);
}

int mpsb_theta(const mpsb_gate2_desc* descs_dev, int ndesc, int nbatch, int d,
               int chiL, int chiM, int chiR, void* out, void* workspace, size_t workspace_bytes,
               void* stream) {
    MPSB_ARG(descs_dev && out, "theta: NULL argument");
    if (theta_uses_tc(d, chiL, chiM, chiR) && workspace != nullptr &&
        workspace_bytes >= mpsb_theta_workspace_bytes(ndesc, nbatch, d, chiL, chiM, chiR)) {
        MPSB_ARG(((uintptr_t)workspace & 255) == 0, "theta: workspace must be 256-byte aligned");
        return launch_theta_tc(descs_dev, ndesc, nbatch, chiL, chiM, chiR, 0, (cf*)out, (int64_t)d * chiL * d * chiR,
                               (float*)workspace, (cudaStream_t)stream);
    }
    return launch_theta(descs_dev, ndesc, nbatch, d, chiL, chiM, chiR, 0, (cf*)out,
                        (int64_t)d * chiL * d * chiR, (cudaStream_t)stream);
}

size_t mpsb_theta_workspace_bytes(int ndesc, int nbatch, int d, int chiL, int chiM, int chiR) {
    if (d != 2 || ndesc <= 0 || nbatch <= 0 || chiL <= 0 || chiM <= 0 || chiR <= 0) return 0;
    return tc_theta_workspace_floats(ndesc * nbatch, chiL, chiM, chiR) * sizeof(float);
}

size_t mpsb_cgemm_tc_workspace_bytes(int M, int N, int K, int nbatch) {
    if (M <= 0 || N <= 0 || K <= 0 || nbatch <= 0) return 0;
    return tc_cgemm_workspace_floats(nbatch, M, N, K) * sizeof(float);
}

int mpsb_cgemm_tc(const void* A, int64_t a_bs, const void* B, int64_t b_bs, void* C, int64_t c_ld, int64_t c_bs,
                  int M, int N, int K, int nbatch, void* workspace, size_t workspace_bytes, void* stream) {
    MPSB_ARG(A && B && C, "cgemm_tc: NULL argument");
    MPSB_ARG(workspace != nullptr && workspace_bytes >= mpsb_cgemm_tc_workspace_bytes(M, N, K, nbatch),
             "cgemm_tc: workspace too small");
    MPSB_ARG(((uintptr_t)workspace & 255) == 0, "cgemm_tc: workspace must be 256-byte aligned");
    return launch_cgemm_tc((const cf*)A, a_bs, (const cf*)B, b_bs, (cf*)C, c_ld, c_bs, M, N, K, nbatch,
                           (float*)workspace, (cudaStream_t)stream);
}

static size_t svd_x_elems(int m, int n) {
    bool small = m <= MPSB_MAX_SMALL_DIM && n <= MPSB_MAX_SMALL_DIM;
    if (small) return (size_t)m * n;
    size_t a = (size_t)svd_large_padded_rows(m) * n, b = (size_t)svd_large_padded_rows(n) * m;
    return a > b ? a : b;
}

size_t mpsb_svd_workspace_bytes(int njobs, int m, int n) {
    if (njobs <= 0 || m <= 0 || n <= 0) return 0;
    bool small = m <= MPSB_MAX_SMALL_DIM && n <= MPSB_MAX_SMALL_DIM;
    size_t ex_a = small ? (size_t)0 : svd_large_workspace_elems(m, n);
    size_t ex_b = small ? (size_t)0 : svd_large_workspace_elems(n, m);
    size_t ex = ex_a > ex_b ? ex_a : ex_b;
    return (align_up(svd_x_elems(m, n), 16) + align_up(ex, 16)) * sizeof(cf) * (size_t)njobs + 256;
}

int mpsb_svd(const void* mats, int njobs, int m, int n, int k, int left_canonical,
             void* left, void* right, float* svals, int32_t* info,
             void* workspace, size_t workspace_bytes, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    MPSB_ARG(mats && left && right, "svd: NULL argument");
    MPSB_ARG(m >= 1 && n >= 1, "svd: empty matrix");
    MPSB_ARG(njobs >= 0 && njobs <= 65535, "svd: njobs %d outside [0, 65535]", njobs);
    if (njobs == 0) return 0;
    MPSB_ARG(workspace_bytes >= mpsb_svd_workspace_bytes(njobs, m, n), "svd: workspace too small");
    int lc = left_canonical ? 1 : 0;
    int nv = lc ? m : n, L = lc ? n : m;
    int mn = m < n ? m : n;
    bool small = m <= MPSB_MAX_SMALL_DIM && n <= MPSB_MAX_SMALL_DIM;
    size_t xs = align_up(svd_x_elems(m, n), 16);
    cf* X = (cf*)workspace;
    cf* extra = X + xs * njobs;
    if (lc) {
        MPSB_CUDA(cudaMemcpy2DAsync(X, xs * sizeof(cf), mats, (size_t)m * n * sizeof(cf), (size_t)m * n * sizeof(cf),
                                    njobs, cudaMemcpyDeviceToDevice, st));
    } else {
        // X[job] = mats[job]^T, written at the padded job stride
        for (int j0 = 0; j0 < njobs; ++j0) {
            dim3 grid((n + 31) / 32, (m + 31) / 32, 1);
            transpose_kernel<<<grid, dim3(32, 8), 0, st>>>((const cf*)mats + (size_t)j0 * m * n, X + (size_t)j0 * xs, m, n);
        }
        MPSB_LAUNCH_CHECK("transpose_kernel");
    }
    if (small)
        return launch_svd_small(X, (int64_t)xs, njobs, nv, L, k, lc, nullptr, 0, 1,
                                (cf*)left, (int64_t)m * k, (cf*)right, (int64_t)k * n, svals, mn, info,
                                extra, st);
    return launch_svd_large(X, (int64_t)xs, njobs, nv, L, k, lc, nullptr, 0, 1,
                            (cf*)left, (int64_t)m * k, (cf*)right, (int64_t)k * n, svals, mn, info,
                            extra, st);
}

}  // extern "C"
