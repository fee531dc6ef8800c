// permanent_kernel.cu -- Nijenhuis-Wilf / Ryser matrix permanent for sm_100a.
//
// The reference (nwPerm.cpp:251-332) walks all 2^(n-1) Gray-code column subsets
// sequentially, updating x_j += +-a[j,k] and accumulating +-prod_j x_j.  Here the
// index range [0, 2^(n-1)) is cut into aligned power-of-two chunks, one per thread:
// a thread re-seeds x from the Gray code of its first index (base_j + sum of the set
// columns), then walks its chunk.  Because chunks are aligned, every thread of a CTA
// flips the SAME column at the same step, so the column is read from shared memory as
// a broadcast; x lives in registers (row count is a template parameter, rows padded
// with x == 1); the running sum is kept as an error-free double-double and reduced in
// a fixed order (lane, warp, CTA, then a finalize kernel over CTAs), so results are
// reproducible and slightly MORE accurate than the reference's single running double.
// Bound: FP64 pipe (n DFMA + n DMUL per subset per thread); the matrix is <= 8 KB.
//
// Rectangular input follows permanentExact (nwPerm.cpp:217-231): pad with ones to
// max(rows, cols), divide by (|rows - cols|)!.
#include "pda_internal.h"
#include "pda_host_stage.h"

#include <algorithm>
#include <vector>

namespace pda {
namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr int PERM_THREADS = 256;

struct dd { double hi, lo; };

__device__ __forceinline__ void two_sum(double a, double b, double& s, double& e) {
    s = a + b;
    const double bb = s - a;
    e = (a - (s - bb)) + (b - bb);
}
__device__ __forceinline__ dd dd_add(dd a, dd b) {
    double s, e;
    two_sum(a.hi, b.hi, s, e);
    e += a.lo + b.lo;
    dd r;
    r.hi = s + e;
    r.lo = e - (r.hi - s);
    return r;
}

// n! for n = 0..32, correctly rounded (tgamma(n+1) of nwPerm.cpp:223)
__constant__ double kFactorial[33] = {
    1.0, 1.0, 2.0, 6.0, 24.0, 120.0, 720.0, 5040.0, 40320.0, 362880.0, 3628800.0, 39916800.0, 479001600.0,
    6227020800.0, 87178291200.0, 1307674368000.0, 20922789888000.0, 355687428096000.0, 6402373705728000.0,
    121645100408832000.0, 2432902008176640000.0, 51090942171709440000.0, 1124000727777607680000.0,
    25852016738884976640000.0, 620448401733239439360000.0, 15511210043330985984000000.0,
    403291461126605635584000000.0, 10888869450418352160768000000.0, 304888344611713860501504000000.0,
    8841761993739701954543616000000.0, 265252859812191058636308480000000.0,
    8222838654177922817725562880000000.0, 263130836933693530167218012160000000.0};

struct PermArgs {
    const double* mats; const int64_t* matOff; const int32_t* rows; const int32_t* cols;
    int64_t nMats;
    // range mode (nMats == 1, square): Gray indices [begin, end), chunk length `chunk`
    unsigned long long begin, end, chunk;
    int rangeMode;
    dd* partial;        // [nMats * ctasPerMat]
    int ctasPerMat;
};

// Walks `len` Gray indices starting at i0 (i0 % len == 0, len a power of two).
template <int NP>
__device__ __forceinline__ dd walk_chunk(const double* __restrict__ sA, const double* __restrict__ sBase,
                                         const unsigned long long i0, const unsigned long long len) {
    double x[NP];
#pragma unroll
    for (int j = 0; j < NP; ++j) x[j] = sBase[j];
    // seed: subset = gray(i0)
    unsigned long long g = i0 ^ (i0 >> 1);
    while (g) {
        const int b = __ffsll((long long)g) - 1;
        g &= g - 1;
        const double* col = sA + b * NP;
#pragma unroll
        for (int j = 0; j < NP; j += 2) {
            const double2 a = *reinterpret_cast<const double2*>(col + j);
            x[j] += a.x;
            x[j + 1] += a.y;
        }
    }
    double hi, lo = 0.0;
    {
        double p0 = 1.0, p1 = 1.0, p2 = 1.0, p3 = 1.0;
#pragma unroll
        for (int j = 0; j < NP; ++j) {
            if ((j & 3) == 0) p0 *= x[j];
            else if ((j & 3) == 1) p1 *= x[j];
            else if ((j & 3) == 2) p2 *= x[j];
            else p3 *= x[j];
        }
        const double prod = (p0 * p1) * (p2 * p3);
        hi = (i0 & 1ULL) ? -prod : prod;
    }
    for (unsigned long long t = 1; t < len; ++t) {
        const unsigned long long i = i0 + t;
        const int k = __ffsll((long long)t) - 1;  // == ctz(i): the bit in which gray(i) and gray(i-1) differ
        const unsigned long long gray = i ^ (i >> 1);
        const double s = ((gray >> k) & 1ULL) ? 1.0 : -1.0;
        const double* col = sA + k * NP;
        double p0 = 1.0, p1 = 1.0, p2 = 1.0, p3 = 1.0;
#pragma unroll
        for (int j = 0; j < NP; j += 2) {
            const double2 a = *reinterpret_cast<const double2*>(col + j);
            x[j] = fma(s, a.x, x[j]);
            x[j + 1] = fma(s, a.y, x[j + 1]);
            if ((j & 3) == 0) { p0 *= x[j]; p1 *= x[j + 1]; }
            else { p2 *= x[j]; p3 *= x[j + 1]; }
        }
        const double prod = (p0 * p1) * (p2 * p3);
        const double term = (i & 1ULL) ? -prod : prod;
        double s2, e;
        two_sum(hi, term, s2, e);
        hi = s2;
        lo += e;
    }
    dd r;
    r.hi = hi;
    r.lo = lo;
    return r;
}

template <int NP>
__global__ void __launch_bounds__(PERM_THREADS) perm_kernel(const PermArgs a) {
    __shared__ __align__(16) double sA[NP * 32];
    __shared__ __align__(16) double sBase[NP];
    __shared__ dd sWarp[PERM_THREADS / 32];
    const int m = blockIdx.y, cta = blockIdx.x, tid = threadIdx.x;
    const int rows = a.rows[m], cols = a.cols[m];
    const int n = rows > cols ? rows : cols;
    dd acc;
    acc.hi = 0.0;
    acc.lo = 0.0;
    if (n >= 1 && n <= NP && n <= PDA_MAX_PERM_DIM) {
        const double* A = a.mats + a.matOff[m];
        // stage the matrix: ones outside the given block (nwPerm.cpp:226-228), zero rows beyond n
        for (int e = tid; e < NP * n; e += PERM_THREADS) {
            const int j = e % NP, k = e / NP;
            double val = 0.0;
            if (j < n) val = (j < rows && k < cols) ? A[j + (size_t)k * rows] : 1.0;
            sA[e] = val;
        }
        __syncthreads();
        if (tid < NP) {
            double b = 1.0;  // padded rows keep x == 1 forever
            if (tid < n) {
                double rs = 0.0;
                for (int k = 0; k < n; ++k) rs += sA[tid + k * NP];
                b = sA[tid + (n - 1) * NP] - rs / 2;  // nwPerm.cpp:289
            }
            sBase[tid] = b;
        }
        __syncthreads();
        unsigned long long begin, end, chunk;
        if (a.rangeMode) { begin = a.begin; end = a.end; chunk = a.chunk; }
        else {
            begin = 0;
            end = 1ULL << (n - 1);
            const unsigned long long threads = (unsigned long long)a.ctasPerMat * PERM_THREADS;
            chunk = end / threads;
            if (chunk < 1) chunk = 1;
        }
        const unsigned long long nChunks = (end - begin) / chunk;
        const unsigned long long gid = (unsigned long long)cta * PERM_THREADS + tid;
        const unsigned long long stride = (unsigned long long)gridDim.x * PERM_THREADS;
        for (unsigned long long c = gid; c < nChunks; c += stride) acc = dd_add(acc, walk_chunk<NP>(sA, sBase, begin + c * chunk, chunk));
    }
    // fixed-order reduction: lanes, then warps
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        dd other;
        other.hi = __shfl_down_sync(FULL, acc.hi, o);
        other.lo = __shfl_down_sync(FULL, acc.lo, o);
        acc = dd_add(acc, other);
    }
    if ((tid & 31) == 0) sWarp[tid >> 5] = acc;
    __syncthreads();
    if (tid == 0) {
        dd t = sWarp[0];
        for (int w = 1; w < PERM_THREADS / 32; ++w) t = dd_add(t, sWarp[w]);
        a.partial[(size_t)m * a.ctasPerMat + cta] = t;
    }
}

// Sums the per-CTA partials in order and applies sign, factor 2 and the rectangular scale.
__global__ void perm_finalize_kernel(const PermArgs a, double* __restrict__ out, int32_t* __restrict__ status,
                                     double* __restrict__ rangePartial) {
    const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= a.nMats) return;
    dd t;
    t.hi = 0.0;
    t.lo = 0.0;
    for (int c = 0; c < a.ctasPerMat; ++c) t = dd_add(t, a.partial[(size_t)m * a.ctasPerMat + c]);
    if (a.rangeMode) {
        rangePartial[0] = t.hi;
        rangePartial[1] = t.lo;
        return;
    }
    const int rows = a.rows[m], cols = a.cols[m];
    const int n = rows > cols ? rows : cols;
    if (n > PDA_MAX_PERM_DIM) { out[m] = 0.0; if (status) status[m] = 1; return; }  // nwPerm.cpp:327-330 throws
    if (status) status[m] = 0;
    if (n == 0) { out[m] = 1.0; return; }  // nwPerm.cpp:261-264
    double p = (double)(4 * (n & 1) - 2) * (t.hi + t.lo);
    if (rows != cols) p = p / kFactorial[rows > cols ? rows - cols : cols - rows];
    out[m] = p;
}

template <int NP>
int launch_np(const PermArgs& a, int gridX, cudaStream_t stream) {
    dim3 grid((unsigned)gridX, (unsigned)a.nMats);
    perm_kernel<NP><<<grid, PERM_THREADS, 0, stream>>>(a);
    PDA_CUDA_TRY(cudaGetLastError());
    return PDA_OK;
}

int dispatch(const PermArgs& a, int maxDim, int gridX, cudaStream_t stream) {
    const int np = std::max(2, (maxDim + 1) / 2 * 2);
    switch (np) {
#define PDA_CASE(N) case N: return launch_np<N>(a, gridX, stream);
        PDA_CASE(2) PDA_CASE(4) PDA_CASE(6) PDA_CASE(8) PDA_CASE(10) PDA_CASE(12) PDA_CASE(14) PDA_CASE(16)
        PDA_CASE(18) PDA_CASE(20) PDA_CASE(22) PDA_CASE(24) PDA_CASE(26) PDA_CASE(28) PDA_CASE(30) PDA_CASE(32)
#undef PDA_CASE
    }
    return fail(PDA_ERR_UNSUPPORTED, "permanent: dimension %d above %d", maxDim, PDA_MAX_PERM_DIM);
}

int pow2_floor(long long x) { int p = 1; while ((long long)p * 2 <= x) p *= 2; return p; }

}  // namespace

// CTAs per matrix: enough CTAs to fill the chip, but at least 64 subsets per thread.
static int ctas_per_matrix(int64_t nMats, int maxDim, int smCount) {
    const long long subsets = 1LL << std::max(0, std::min(maxDim, PDA_MAX_PERM_DIM) - 1);
    long long byWork = std::max(1LL, subsets / ((long long)PERM_THREADS * 64));
    long long byChip = std::max<long long>(1, (2LL * smCount + nMats - 1) / nMats);
    return pow2_floor(std::max(1LL, std::min(byWork, std::min(byChip, 1024LL))));
}

int launch_permanent_batch(const double* mats, const int64_t* matOff, const int32_t* rows, const int32_t* cols,
                           int64_t nMats, int32_t maxDim, double* out, int32_t* status, void* workspace,
                           int64_t workspaceBytes, cudaStream_t stream) {
    DeviceInfo dev;
    PDA_TRY(current_device_info(&dev));
    const int effDim = std::min<int>(maxDim, PDA_MAX_PERM_DIM);
    int cpm = ctas_per_matrix(nMats, effDim, dev.smCount);
    while (cpm > 1 && (int64_t)cpm * nMats * (int64_t)sizeof(dd) > workspaceBytes) cpm >>= 1;
    if ((int64_t)cpm * nMats * (int64_t)sizeof(dd) > workspaceBytes)
        return fail(PDA_ERR_WORKSPACE, "permanent: workspace of %lld B too small (need %lld)", (long long)workspaceBytes,
                    (long long)(nMats * (int64_t)sizeof(dd)));
    PermArgs a = {mats, matOff, rows, cols, nMats, 0, 0, 0, 0, reinterpret_cast<dd*>(workspace), cpm};
    // blockIdx.y is limited to 65535: slice the batch
    for (int64_t m0 = 0; m0 < nMats; m0 += 65535) {
        PermArgs s = a;
        s.matOff = matOff + m0; s.rows = rows + m0; s.cols = cols + m0;
        s.nMats = std::min<int64_t>(65535, nMats - m0);
        s.partial = a.partial + m0 * cpm;
        PDA_TRY(dispatch(s, std::max(effDim, 1), cpm, stream));
    }
    perm_finalize_kernel<<<(unsigned)((nMats + 127) / 128), 128, 0, stream>>>(a, out, status, nullptr);
    PDA_CUDA_TRY(cudaGetLastError());
    return PDA_OK;
}

int launch_permanent_range(const double* A, int32_t n, uint64_t begin, uint64_t end, double* partial,
                           void* workspace, int64_t workspaceBytes, cudaStream_t stream) {
    DeviceInfo dev;
    PDA_TRY(current_device_info(&dev));
    // chunk: largest power of two dividing both ends, capped so that there are enough chunks to fill the chip
    const unsigned long long total = end - begin;
    unsigned long long align = (begin | end) ? ((begin | end) & (~(begin | end) + 1ULL)) : (1ULL << 62);
    unsigned long long chunk = 1;
    const unsigned long long wantThreads = 2ULL * dev.smCount * PERM_THREADS;
    while (chunk * 2 <= align && total / (chunk * 2) >= wantThreads) chunk *= 2;
    while (chunk * 2 <= align && chunk < 64 && total / (chunk * 2) >= 1) chunk *= 2;
    const unsigned long long nChunks = total / chunk;
    int gridX = (int)std::min<unsigned long long>((nChunks + PERM_THREADS - 1) / PERM_THREADS, 8ULL * dev.smCount);
    gridX = std::max(gridX, 1);
    if ((int64_t)gridX * (int64_t)sizeof(dd) + 64 > workspaceBytes)
        return fail(PDA_ERR_WORKSPACE, "permanent_range: workspace too small");
    // the single matrix is described by tiny device-side descriptors at the head of the workspace
    unsigned char* ws = reinterpret_cast<unsigned char*>(workspace);
    int64_t* dOff = reinterpret_cast<int64_t*>(ws);
    int32_t* dRows = reinterpret_cast<int32_t*>(ws + 8);
    int32_t* dCols = reinterpret_cast<int32_t*>(ws + 12);
    struct { int64_t off; int32_t r, c; } desc = {0, n, n};
    PDA_CUDA_TRY(cudaMemcpyAsync(ws, &desc, 16, cudaMemcpyHostToDevice, stream));
    PermArgs a = {A, dOff, dRows, dCols, 1, begin, end, chunk, 1, reinterpret_cast<dd*>(ws + 64), gridX};
    PDA_TRY(dispatch(a, n, gridX, stream));
    perm_finalize_kernel<<<1, 32, 0, stream>>>(a, nullptr, nullptr, partial);
    PDA_CUDA_TRY(cudaGetLastError());
    return PDA_OK;
}

}  // namespace pda

using namespace pda;

extern "C" {

int64_t pda_permanent_workspace_bytes(int64_t nMats) {
    return 64 + 16 * (std::max<int64_t>(nMats, 1) + 1024);
}

int pda_permanent_batch(const double* mats, const int64_t* matOff, const int32_t* rows, const int32_t* cols,
                        int64_t nMats, int32_t maxDim, double* out, int32_t* status,
                        void* workspace, int64_t workspaceBytes, void* stream) {
    if (nMats < 0) return fail(PDA_ERR_INVALID, "permanent: nMats < 0");
    if (nMats == 0) return PDA_OK;
    if (!mats || !matOff || !rows || !cols || !out || !workspace) return fail(PDA_ERR_INVALID, "permanent: NULL argument");
    if (maxDim < 0) return fail(PDA_ERR_INVALID, "permanent: maxDim < 0");
    return launch_permanent_batch(mats, matOff, rows, cols, nMats, maxDim, out, status, workspace, workspaceBytes,
                                  reinterpret_cast<cudaStream_t>(stream));
}

int pda_permanent_batch_host(const double* mats, const int64_t* matOff, const int32_t* rows, const int32_t* cols,
                             int64_t nMats, double* out, int32_t* status, int32_t device) {
    if (nMats < 0) return fail(PDA_ERR_INVALID, "permanent: nMats < 0");
    if (nMats == 0) return PDA_OK;
    if (!mats || !matOff || !rows || !cols || !out) return fail(PDA_ERR_INVALID, "permanent: NULL argument");
    size_t nEl = 0;
    int maxDim = 0;
    for (int64_t i = 0; i < nMats; ++i) {
        if (rows[i] < 0 || cols[i] < 0) return fail(PDA_ERR_INVALID, "permanent: negative dimension");
        nEl = std::max(nEl, (size_t)matOff[i] + (size_t)rows[i] * cols[i]);
        maxDim = std::max(maxDim, std::max(rows[i], cols[i]));
    }
    std::lock_guard<std::mutex> lk(g_hostMu);
    PDA_TRY(check_device(device));
    DeviceInfo dev;
    PDA_TRY(current_device_info(&dev));
    const size_t n = (size_t)nMats;
    const int cpm = ctas_per_matrix(nMats, std::min(maxDim, PDA_MAX_PERM_DIM), dev.smCount);
    const size_t wsBytes = n * cpm * sizeof(double) * 2;
    Stage st(device);
    const size_t oM = st.reserve(nEl * 8), oOff = st.reserve(n * 8), oR = st.reserve(n * 4), oC = st.reserve(n * 4);
    const size_t oOut = st.reserve(n * 8), oSt = st.reserve(n * 4), oWs = st.reserve(wsBytes);
    PDA_TRY(st.commit());
    cudaStream_t s = 0;
    PDA_TRY(h2d(st.at<double>(oM), mats, nEl, s));
    PDA_TRY(h2d(st.at<int64_t>(oOff), matOff, n, s));
    PDA_TRY(h2d(st.at<int32_t>(oR), rows, n, s));
    PDA_TRY(h2d(st.at<int32_t>(oC), cols, n, s));
    PDA_TRY(launch_permanent_batch(st.at<double>(oM), st.at<int64_t>(oOff), st.at<int32_t>(oR), st.at<int32_t>(oC), nMats,
                                   maxDim, st.at<double>(oOut), st.at<int32_t>(oSt), st.at<unsigned char>(oWs),
                                   (int64_t)wsBytes, s));
    PDA_TRY(d2h(out, st.at<double>(oOut), n, s));
    PDA_TRY(d2h(status, st.at<int32_t>(oSt), n, s));
    PDA_CUDA_TRY(cudaStreamSynchronize(s));
    return PDA_OK;
}

int pda_permanent_range(const double* A, int32_t n, uint64_t begin, uint64_t end, double* partial,
                        void* workspace, int64_t workspaceBytes, void* stream) {
    if (!A || !partial || !workspace) return fail(PDA_ERR_INVALID, "permanent_range: NULL argument");
    if (n < 1 || n > PDA_MAX_PERM_DIM) return fail(PDA_ERR_UNSUPPORTED, "permanent_range: n = %d outside 1..%d", n, PDA_MAX_PERM_DIM);
    if (begin > end || end > (1ULL << (n - 1))) return fail(PDA_ERR_INVALID, "permanent_range: bad Gray range");
    if (begin == end) {
        PDA_CUDA_TRY(cudaMemsetAsync(partial, 0, 16, reinterpret_cast<cudaStream_t>(stream)));
        return PDA_OK;
    }
    return launch_permanent_range(A, n, begin, end, partial, workspace, workspaceBytes, reinterpret_cast<cudaStream_t>(stream));
}

int pda_permanent_range_host(const double* A, int32_t n, uint64_t begin, uint64_t end, double* partial, int32_t device) {
    if (!A || !partial) return fail(PDA_ERR_INVALID, "permanent_range: NULL argument");
    if (n < 1 || n > PDA_MAX_PERM_DIM) return fail(PDA_ERR_UNSUPPORTED, "permanent_range: n = %d outside 1..%d", n, PDA_MAX_PERM_DIM);
    std::lock_guard<std::mutex> lk(g_hostMu);
    PDA_TRY(check_device(device));
    DeviceInfo dev;
    PDA_TRY(current_device_info(&dev));
    const size_t wsBytes = 64 + (size_t)8 * dev.smCount * 16;
    Stage st(device);
    const size_t oA = st.reserve((size_t)n * n * 8), oP = st.reserve(16), oWs = st.reserve(wsBytes);
    PDA_TRY(st.commit());
    cudaStream_t s = 0;
    PDA_TRY(h2d(st.at<double>(oA), A, (size_t)n * n, s));
    PDA_TRY(pda_permanent_range(st.at<double>(oA), n, begin, end, st.at<double>(oP), st.at<unsigned char>(oWs), (int64_t)wsBytes, s));
    PDA_TRY(d2h(partial, st.at<double>(oP), 2, s));
    PDA_CUDA_TRY(cudaStreamSynchronize(s));
    return PDA_OK;
}

}  // extern "C"
