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

#ifndef PERM_CHAINS
#define PERM_CHAINS 4
#endif
#ifndef PERM_UNROLL
#define PERM_UNROLL 1
#endif
constexpr int kPermUnroll = PERM_UNROLL;  // subsets per loop trip

struct PermArgs {
    const double* mats; const int64_t* matOff; const int32_t* rows; const int32_t* cols;
    int64_t nMats;
    // range mode (nMats == 1, square): Gray indices [begin, end)
    unsigned long long begin, end;
    int rangeMode;
    int chunk;          // alignment of the per-thread ranges (power of two): below it every thread of a CTA flips
                        // the same column at the same step, so the column read is a shared-memory broadcast
    int threadsPerMat;  // 32..256 (power of two): threads of one CTA that share a matrix
    int ctasPerMat;     // CTAs (blockIdx.x) that share a matrix
    int maxN;           // largest dimension in the batch (sizes the shared-memory slots)
    dd* partial;        // [nMats * ctasPerMat]
    int dpBits;         // matrices whose SMALLER side is <= dpBits go to perm_dp_kernel instead (-1: none)
    // fused finalisation: the CTA that finishes a matrix last sums the per-CTA partials (same fixed order as
    // perm_finalize_kernel) and writes the result, so a large single matrix costs ONE launch
    int fuse;
    unsigned* done;     // [nMats] arrival counters (zero before the launch; the finishing CTA resets its counter)
    double* out; int32_t* status; double* rangePartial;
    int oneDim;         // > 0: a single oneDim x oneDim matrix at mats[0]; rows / cols / matOff are not read (range mode)
};

// Matrices with a small side are not walked by the NW kernel at all: see perm_dp_kernel.
constexpr int PERM_DP_MAX = 10;
__device__ __forceinline__ bool perm_uses_dp(const int rows, const int cols, const int dpBits) {
    return (rows < cols ? rows : cols) <= dpBits;
}

// Leading dimension of the matrix in shared memory.  When the threads of a warp flip DIFFERENT columns (the first
// index of every aligned chunk) each lane reads column k at k*LD + j; with LD == NP the columns of a 24-row matrix
// start only two distinct 128-byte phases apart and those reads serialise (ncu: a third of all shared-memory
// wavefronts of the kernel were bank conflicts).  LD*8 bytes == an odd multiple of 16 modulo 128 spreads eight
// consecutive columns over all eight 16-byte slots of a 128-byte line.
__host__ __device__ constexpr int perm_ld(int np) { return ((np + 2) / 2) % 2 ? np + 2 : np + 4; }

// Walks Gray indices [i0, i1) of the NW sum: seeds x for the subset gray(i0), adds its term, then steps.
// Correct for ANY i0, i1; the column index ctz(i) is warp-uniform whenever it is below log2(chunk).
template <int NP>
__device__ __forceinline__ dd walk_range(const double* __restrict__ sA, const double* __restrict__ sBase,
                                         const unsigned long long i0, const unsigned long long i1) {
    double x[NP];
#pragma unroll
    for (int j = 0; j < NP; ++j) x[j] = sBase[j];
    unsigned long long g = i0 ^ (i0 >> 1);
    while (g) {
        const int b = __ffsll((long long)g) - 1;
        g &= g - 1;
        const double* col = sA + b * perm_ld(NP);
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
#pragma unroll kPermUnroll
    for (unsigned long long i = i0 + 1; i < i1; ++i) {
        const int k = __ffsll((long long)i) - 1;  // the bit in which gray(i) and gray(i-1) differ
        const unsigned long long gray = i ^ (i >> 1);
        const double s = ((gray >> k) & 1ULL) ? 1.0 : -1.0;
        const double* col = sA + k * perm_ld(NP);
        double p[PERM_CHAINS];
#pragma unroll
        for (int c = 0; c < PERM_CHAINS; ++c) p[c] = 1.0;
#pragma unroll
        for (int j = 0; j < NP; j += 2) {
            const double2 a = *reinterpret_cast<const double2*>(col + j);
            x[j] = fma(s, a.x, x[j]);
            x[j + 1] = fma(s, a.y, x[j + 1]);
            p[j % PERM_CHAINS] *= x[j];
            p[(j + 1) % PERM_CHAINS] *= x[j + 1];
        }
#pragma unroll
        for (int w = PERM_CHAINS / 2; w >= 1; w >>= 1)
#pragma unroll
            for (int c = 0; c < w; ++c) p[c] *= p[c + w];
        const double prod = p[0];
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

// Variant of walk_range that keeps column 0 in registers.  In Gray-code order every ODD index flips column 0,
// so half of all steps then need no shared-memory read at all.  That matters because a broadcast LDS.128 still
// writes 512 B into the register file: at one column read per subset the kernel is bound by that return path,
// not by the FP64 pipe.  Requires i0 even (ranges are chunk aligned) -- the loop alternates odd / even indices.
template <int NP>
__device__ __forceinline__ dd walk_range_c0(const double* __restrict__ sA, const double* __restrict__ sBase,
                                            const unsigned long long i0, const unsigned long long i1) {
    double x[NP], c0[NP];
#pragma unroll
    for (int j = 0; j < NP; ++j) { x[j] = sBase[j]; c0[j] = sA[j]; }
    unsigned long long g = i0 ^ (i0 >> 1);
    while (g) {
        const int b = __ffsll((long long)g) - 1;
        g &= g - 1;
        const double* col = sA + b * perm_ld(NP);
#pragma unroll
        for (int j = 0; j < NP; j += 2) {
            const double2 a = *reinterpret_cast<const double2*>(col + j);
            x[j] += a.x;
            x[j + 1] += a.y;
        }
    }
    double hi, lo = 0.0;
    {
        double p[PERM_CHAINS];
#pragma unroll
        for (int c = 0; c < PERM_CHAINS; ++c) p[c] = 1.0;
#pragma unroll
        for (int j = 0; j < NP; ++j) p[j % PERM_CHAINS] *= x[j];
#pragma unroll
        for (int w = PERM_CHAINS / 2; w >= 1; w >>= 1)
#pragma unroll
            for (int c = 0; c < w; ++c) p[c] *= p[c + w];
        hi = p[0];  // i0 is even: sign +
    }
    for (unsigned long long i = i0 + 1; i < i1; i += 2) {
        {   // odd index i: column 0 from registers; gray bit 0 of an odd i is 1 ^ bit1(i)
            const double s = ((i >> 1) & 1ULL) ? -1.0 : 1.0;
            double p[PERM_CHAINS];
#pragma unroll
            for (int c = 0; c < PERM_CHAINS; ++c) p[c] = 1.0;
#pragma unroll
            for (int j = 0; j < NP; ++j) {
                x[j] = fma(s, c0[j], x[j]);
                p[j % PERM_CHAINS] *= x[j];
            }
#pragma unroll
            for (int w = PERM_CHAINS / 2; w >= 1; w >>= 1)
#pragma unroll
                for (int c = 0; c < w; ++c) p[c] *= p[c + w];
            double s2, e;
            two_sum(hi, -p[0], s2, e);  // odd index: sign -
            hi = s2;
            lo += e;
        }
        const unsigned long long ie = i + 1;
        if (ie < i1) {  // even index: column ctz(ie) >= 1 from shared memory
            const int k = __ffsll((long long)ie) - 1;
            const double s = (((ie ^ (ie >> 1)) >> k) & 1ULL) ? 1.0 : -1.0;
            const double* col = sA + k * perm_ld(NP);
            double p[PERM_CHAINS];
#pragma unroll
            for (int c = 0; c < PERM_CHAINS; ++c) p[c] = 1.0;
#pragma unroll
            for (int j = 0; j < NP; j += 2) {
                const double2 a = *reinterpret_cast<const double2*>(col + j);
                x[j] = fma(s, a.x, x[j]);
                x[j + 1] = fma(s, a.y, x[j + 1]);
                p[j % PERM_CHAINS] *= x[j];
                p[(j + 1) % PERM_CHAINS] *= x[j + 1];
            }
#pragma unroll
            for (int w = PERM_CHAINS / 2; w >= 1; w >>= 1)
#pragma unroll
                for (int c = 0; c < w; ++c) p[c] *= p[c + w];
            double s2, e;
            two_sum(hi, p[0], s2, e);  // even index: sign +
            hi = s2;
            lo += e;
        }
    }
    dd r;
    r.hi = hi;
    r.lo = lo;
    return r;
}

// Small matrices: columns 0 AND 1 in registers.  Column 1 flips at every index == 2 (mod 4), so three steps out of
// four then need no shared-memory read.  Requires i0 == 0 (mod 4) and i1 - i0 a multiple of 4 (ranges are chunk aligned
// with chunk >= 4 whenever this variant is selected).
template <int NP>
__device__ __forceinline__ dd walk_range_c01(const double* __restrict__ sA, const double* __restrict__ sBase,
                                             const unsigned long long i0, const unsigned long long i1) {
    double x[NP], c0[NP], c1[NP];
#pragma unroll
    for (int j = 0; j < NP; ++j) { x[j] = sBase[j]; c0[j] = sA[j]; c1[j] = sA[perm_ld(NP) + j]; }
    unsigned long long g = i0 ^ (i0 >> 1);
    while (g) {
        const int b = __ffsll((long long)g) - 1;
        g &= g - 1;
        const double* col = sA + b * perm_ld(NP);
#pragma unroll
        for (int j = 0; j < NP; j += 2) {
            const double2 a = *reinterpret_cast<const double2*>(col + j);
            x[j] += a.x;
            x[j + 1] += a.y;
        }
    }
    auto product = [&]() {
        double p[PERM_CHAINS];
#pragma unroll
        for (int c = 0; c < PERM_CHAINS; ++c) p[c] = 1.0;
#pragma unroll
        for (int j = 0; j < NP; ++j) p[j % PERM_CHAINS] *= x[j];
#pragma unroll
        for (int w = PERM_CHAINS / 2; w >= 1; w >>= 1)
#pragma unroll
            for (int c = 0; c < w; ++c) p[c] *= p[c + w];
        return p[0];
    };
    double hi = product(), lo = 0.0;  // i0 is even: sign +
    auto add = [&](double term) {
        double s2, e;
        two_sum(hi, term, s2, e);
        hi = s2;
        lo += e;
    };
    for (unsigned long long i = i0; i < i1; i += 4) {
        if (i != i0) {  // index i == 0 (mod 4): column ctz(i) >= 2 from shared memory
            const int k = __ffsll((long long)i) - 1;
            const double s = (((i ^ (i >> 1)) >> k) & 1ULL) ? 1.0 : -1.0;
            const double* col = sA + k * perm_ld(NP);
#pragma unroll
            for (int j = 0; j < NP; j += 2) {
                const double2 a = *reinterpret_cast<const double2*>(col + j);
                x[j] = fma(s, a.x, x[j]);
                x[j + 1] = fma(s, a.y, x[j + 1]);
            }
            add(product());
        }
        {   // i + 1 (odd): column 0; gray bit 0 = 1 ^ bit 1 of the index = 1  -> +
#pragma unroll
            for (int j = 0; j < NP; ++j) x[j] += c0[j];
            add(-product());
        }
        {   // i + 2: column 1; gray bit 1 = 1 ^ bit 2 of the index
            const double s = ((i >> 2) & 1ULL) ? -1.0 : 1.0;
#pragma unroll
            for (int j = 0; j < NP; ++j) x[j] = fma(s, c1[j], x[j]);
            add(product());
        }
        {   // i + 3 (odd): column 0; gray bit 0 = 1 ^ bit 1 = 0  -> -
#pragma unroll
            for (int j = 0; j < NP; ++j) x[j] -= c0[j];
            add(-product());
        }
    }
    dd r;
    r.hi = hi;
    r.lo = lo;
    return r;
}

#ifndef PERM_C0_MAX
#define PERM_C0_MAX 32
#endif
#ifndef PERM_C01_MAX
#define PERM_C01_MAX 16
#endif
__host__ __device__ constexpr bool perm_cache01(int np) { return np <= PERM_C01_MAX; }
__host__ __device__ constexpr bool perm_cache0(int np) { return np <= PERM_C0_MAX; }
// resident CTAs per SM the register allocator must leave room for
__host__ __device__ constexpr int perm_min_blocks(int np) {
    if (perm_cache01(np)) return np <= 8 ? 4 : 2;
    return perm_cache0(np) ? (np <= 12 ? 4 : (np <= 16 ? 3 : 2)) : (np <= 16 ? 4 : (np <= 24 ? 3 : 2));
}

// Sums the per-CTA partials of matrix m (lane-strided, then a fixed shuffle tree: the order depends only on ctasPerMat,
// so results are reproducible) and applies sign, factor 2 and the rectangular scale.  One warp.
__device__ __forceinline__ void perm_finish_warp(const PermArgs& a, const int64_t m, double* __restrict__ out,
                                                 int32_t* __restrict__ status, double* __restrict__ rangePartial, const int lane) {
    dd t;
    t.hi = 0.0;
    t.lo = 0.0;
    for (int c = lane; c < a.ctasPerMat; c += 32) {
        const double2 v = __ldcg(reinterpret_cast<const double2*>(a.partial + (size_t)m * a.ctasPerMat + c));  // written by other CTAs
        dd o;
        o.hi = v.x; o.lo = v.y;
        t = dd_add(t, o);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        dd other;
        other.hi = __shfl_down_sync(FULL, t.hi, o);
        other.lo = __shfl_down_sync(FULL, t.lo, o);
        t = dd_add(t, other);
    }
    if (lane != 0) return;
    if (a.rangeMode) {
        rangePartial[0] = t.hi;
        rangePartial[1] = t.lo;
        return;
    }
    const int rows = a.oneDim > 0 ? a.oneDim : a.rows[m], cols = a.oneDim > 0 ? a.oneDim : a.cols[m];
    const int n = rows > cols ? rows : cols;
    if (n > PDA_MAX_PERM_DIM) { out[m] = 0.0; if (status) status[m] = 1; return; }  // nwPerm.cpp:327-330 throws
    if (status) status[m] = 0;
    if (n == 0) { out[m] = 1.0; return; }  // nwPerm.cpp:261-264
    double p = (double)(4 * (n & 1) - 2) * (t.hi + t.lo);
    if (rows != cols) p = p / kFactorial[rows > cols ? rows - cols : cols - rows];
    out[m] = p;
}

template <int NP>
__global__ void __launch_bounds__(PERM_THREADS, perm_min_blocks(NP)) perm_kernel(const PermArgs a) {
    extern __shared__ __align__(16) double smemD[];
    __shared__ dd sWarp[PERM_THREADS / 32];
    const int tid = threadIdx.x, tpm = a.threadsPerMat;
    const int slot = tid / tpm, t = tid % tpm, slots = PERM_THREADS / tpm;
    const int64_t m = (int64_t)blockIdx.y * slots + slot;
    const int cta = blockIdx.x;
    constexpr int LD = perm_ld(NP);
    const int slotDoubles = LD * a.maxN + NP;
    double* sA = smemD + (size_t)slot * slotDoubles;
    double* sBase = sA + LD * a.maxN;
    const bool live = m < a.nMats;
    const int rows = live ? (a.oneDim > 0 ? a.oneDim : a.rows[m]) : 0, cols = live ? (a.oneDim > 0 ? a.oneDim : a.cols[m]) : 0;
    const int n = rows > cols ? rows : cols;
    const bool ok = live && n >= 1 && n <= NP && n <= PDA_MAX_PERM_DIM && n <= a.maxN && !perm_uses_dp(rows, cols, a.dpBits);
    if (ok) {
        const double* A = a.mats + (a.oneDim > 0 ? 0 : a.matOff[m]);
        // stage the matrix: ones outside the given block (nwPerm.cpp:226-228), zero rows beyond n
        for (int e = t; e < NP * n; e += tpm) {
            const int j = e % NP, k = e / NP;
            double val = 0.0;
            if (j < n) val = (j < rows && k < cols) ? A[j + (size_t)k * rows] : 1.0;
            sA[j + k * LD] = val;
        }
    }
    __syncthreads();
    if (ok && t < NP) {
        double b = 1.0;  // padded rows keep x == 1 forever
        if (t < n) {
            double rs = 0.0;
            for (int k = 0; k < n; ++k) rs += sA[t + k * LD];
            b = sA[t + (n - 1) * LD] - rs / 2;  // nwPerm.cpp:289
        }
        sBase[t] = b;
    }
    __syncthreads();
    dd acc;
    acc.hi = 0.0;
    acc.lo = 0.0;
    if (ok) {
        unsigned long long begin = 0, end = 1ULL << (n - 1);
        if (a.rangeMode) { begin = a.begin; end = a.end; }
        // contiguous, chunk-aligned share of [begin, end) for this thread
        const unsigned long long total = end - begin, chunk = (unsigned long long)a.chunk;
        const unsigned long long nChunks = (total + chunk - 1) / chunk;
        const unsigned long long threads = (unsigned long long)a.ctasPerMat * tpm;
        const unsigned long long per = (nChunks + threads - 1) / threads;
        const unsigned long long g = (unsigned long long)cta * tpm + t;
        unsigned long long lo = begin + g * per * chunk, hi = lo + per * chunk;
        if (hi > end) hi = end;
        if (g * per < nChunks && lo < hi) {
            if (perm_cache01(NP) && (lo & 3ULL) == 0 && ((hi - lo) & 3ULL) == 0) acc = walk_range_c01<NP>(sA, sBase, lo, hi);
            else if (perm_cache0(NP) && (lo & 1ULL) == 0) acc = walk_range_c0<NP>(sA, sBase, lo, hi);
            else acc = walk_range<NP>(sA, sBase, lo, hi);
        }
    }
    // fixed-order reduction: lanes, then the warps that share the matrix
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        dd other;
        other.hi = __shfl_down_sync(FULL, acc.hi, o);
        other.lo = __shfl_down_sync(FULL, acc.lo, o);
        acc = dd_add(acc, other);
    }
    if ((tid & 31) == 0) sWarp[tid >> 5] = acc;
    __syncthreads();
    if (live && t == 0) {
        const int w0 = tid >> 5, nw = tpm >> 5;
        dd tot = sWarp[w0];
        for (int w = 1; w < nw; ++w) tot = dd_add(tot, sWarp[w0 + w]);
        a.partial[(size_t)m * a.ctasPerMat + cta] = tot;
    }
    if (a.fuse) {  // the CTA that completes a matrix finalises it: no second launch
        __shared__ int sLast[PERM_THREADS / 32];
        if (live && t == 0) {
            bool last = true;
            if (a.ctasPerMat > 1) {
                __threadfence();  // the partial above is visible before the arrival is counted
                last = atomicAdd(&a.done[m], 1u) == (unsigned)a.ctasPerMat - 1u;
            }
            sLast[slot] = last ? 1 : 0;
        }
        __syncthreads();
        if (live && t < 32 && sLast[slot] && !(a.dpBits >= 0 && perm_uses_dp(rows, cols, a.dpBits))) {
            __threadfence();
            perm_finish_warp(a, m, a.out, a.status, a.rangePartial, t);
            if (t == 0 && a.ctasPerMat > 1) a.done[m] = 0u;  // ready for the next launch on this stream
        }
    }
}

// ---- matrices with a small side: subset dynamic programme ---------------------------------------------------------
// permanentExact pads an m x n matrix (m < n) with ones to n x n and divides by (n-m)! (nwPerm.cpp:223-230); what that
// computes is the sum over all injective maps of the m short-side indices into the n long-side ones.  The NW walk
// over the padded matrix costs n * 2^(n-1) and -- alternating signs over products of row sums -- cancels: on the
// shapes permanentProb produces (3-8 detections against 10-30 landmark rows, entries spanning many decades) the
// reference's own result is good to 1e-8 at best.  The same sum by dynamic programming over subsets of the SHORT side,
// one long-side index at a time,
//     f_j[S] = f_{j-1}[S] + sum_{i in S} B[i, j] * f_{j-1}[S \ {i}],        perm = f_n[all],
// costs n * 2^m * m / 2 multiply-adds and, for the non-negative matrices of this path, adds only non-negative terms:
// no cancellation, relative error ~ (n + m) ulp.  One warp per matrix, f double-buffered in shared memory, states
// strided over the lanes (S ^ (1 << i) keeps the bank of S: conflict-free).  Used whenever the short side is at most
// PERM_DP_MAX (square matrices included); the reference's n > 32 limit is kept (status 1) so callers see the same
// "throws" behaviour.
__device__ __forceinline__ void perm_dp_warp(const PermArgs& a, const int64_t it, const int rows, const int cols,
                                             double* __restrict__ buf, double* __restrict__ out,
                                             int32_t* __restrict__ status, const int lane) {
    const int m = rows < cols ? rows : cols, n = rows < cols ? cols : rows;
    if (n > PDA_MAX_PERM_DIM) { if (lane == 0) { out[it] = 0.0; if (status) status[it] = 1; } return; }  // nwPerm.cpp:327-330
    const int cap = 1 << a.dpBits, states = 1 << m;
    double* cur = buf;
    double* nxt = cur + cap;
    double* sb = cur + 2 * cap;
    for (int S = lane; S < states; S += 32) cur[S] = (S == 0) ? 1.0 : 0.0;
    const double* A = a.mats + a.matOff[it];
    const bool shortRows = rows <= cols;
    for (int j = 0; j < n; ++j) {
        __syncwarp();
        if (lane < m) sb[lane] = shortRows ? A[lane + (size_t)j * rows] : A[j + (size_t)lane * rows];
        __syncwarp();
        for (int S = lane; S < states; S += 32) {
            double acc = cur[S];
            unsigned bits = (unsigned)S;
            while (bits) {
                const int i = __ffs(bits) - 1;
                bits &= bits - 1u;
                acc = fma(sb[i], cur[S ^ (1 << i)], acc);
            }
            nxt[S] = acc;
        }
        double* t = cur; cur = nxt; nxt = t;
    }
    __syncwarp();
    if (lane == 0) { out[it] = cur[states - 1]; if (status) status[it] = 0; }
}

// One warp per matrix: sums the per-CTA partials (lane-strided, then a fixed shuffle tree -- the order depends
// only on ctasPerMat, so results are reproducible) and applies sign, factor 2 and the rectangular scale.
__global__ void perm_finalize_kernel(const PermArgs a, double* __restrict__ out, int32_t* __restrict__ status,
                                     double* __restrict__ rangePartial) {
    extern __shared__ __align__(16) double smemD[];
    const int lane = threadIdx.x & 31;
    const int64_t m = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (m >= a.nMats) return;
    if (a.dpBits >= 0 && !a.rangeMode) {
        const int rows = a.oneDim > 0 ? a.oneDim : a.rows[m], cols = a.oneDim > 0 ? a.oneDim : a.cols[m];
        if (perm_uses_dp(rows, cols, a.dpBits)) {  // the NW kernel skipped this one
            perm_dp_warp(a, m, rows, cols, smemD + (size_t)(threadIdx.x >> 5) * (2 * ((size_t)1 << a.dpBits) + 16), out, status, lane);
            return;
        }
    }
    if (a.fuse) return;  // perm_kernel finalised this one itself
    perm_finish_warp(a, m, out, status, rangePartial, lane);
}

template <int NP>
int launch_np(const PermArgs& a, cudaStream_t stream) {
    const int slots = PERM_THREADS / a.threadsPerMat;
    const size_t smem = (size_t)slots * (perm_ld(NP) * a.maxN + NP) * sizeof(double);
    PDA_CUDA_TRY(cudaFuncSetAttribute(perm_kernel<NP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)a.ctasPerMat, (unsigned)((a.nMats + slots - 1) / slots));
    perm_kernel<NP><<<grid, PERM_THREADS, smem, stream>>>(a);
    PDA_CUDA_TRY(cudaGetLastError());
    return PDA_OK;
}

int dispatch(const PermArgs& a, cudaStream_t stream) {
    const int np = std::max(2, (a.maxN + 1) / 2 * 2);
    switch (np) {
#define PDA_CASE(N) case N: return launch_np<N>(a, stream);
        PDA_CASE(2) PDA_CASE(4) PDA_CASE(6) PDA_CASE(8) PDA_CASE(10) PDA_CASE(12) PDA_CASE(14) PDA_CASE(16)
        PDA_CASE(18) PDA_CASE(20) PDA_CASE(22) PDA_CASE(24) PDA_CASE(26) PDA_CASE(28) PDA_CASE(30) PDA_CASE(32)
#undef PDA_CASE
    }
    return fail(PDA_ERR_UNSUPPORTED, "permanent: dimension %d above %d", a.maxN, PDA_MAX_PERM_DIM);
}

}  // namespace

// Launch shape for a batch whose largest dimension is maxDim: threads that share a matrix (small matrices get
// one warp each, eight to a CTA), CTAs per matrix (a few large matrices are spread over the whole chip, in
// multiples that fill 2 CTAs per SM evenly), and the alignment of the per-thread index ranges.
struct PermShape { int threadsPerMat, ctasPerMat, chunk; };
static PermShape perm_shape(int64_t nMats, int maxDim, int smCount, unsigned long long rangeLen) {
    const int n = std::max(1, std::min(maxDim, PDA_MAX_PERM_DIM));
    const unsigned long long steps = rangeLen ? rangeLen : (1ULL << (n - 1));
    PermShape sh;
    int tpm = 32;
    while (tpm < PERM_THREADS && (unsigned long long)tpm * 64 < steps) tpm *= 2;
    sh.threadsPerMat = tpm;
    const long long slots = PERM_THREADS / tpm;
    const long long ctasForBatch = (nMats + slots - 1) / slots;
    long long cpm = 1;
    if (tpm == PERM_THREADS) {
        // few large matrices: spread each over the chip.  Prefer a CTA count that fills every SM evenly
        // (a multiple of the SM count) while leaving >= 96 subsets per thread; otherwise as many as the work allows.
        const int np = std::max(2, (n + 1) / 2 * 2);
        const long long byWork = (long long)std::max<unsigned long long>(1, steps / ((unsigned long long)PERM_THREADS * 96));
        const long long perSmMax = perm_min_blocks(np);
        long long total = std::min(byWork * ctasForBatch, perSmMax * smCount);           // CTAs for the whole batch
        if (total >= smCount) total = total / smCount * smCount;
        cpm = std::max(1LL, std::min(total / ctasForBatch, 4096LL));
    }
    sh.ctasPerMat = (int)cpm;
    const unsigned long long perThread = std::max<unsigned long long>(1, steps / ((unsigned long long)cpm * tpm));
    int chunk = 1;
    while (chunk < 32 && (unsigned long long)chunk * 2 * 6 <= perThread) chunk *= 2;  // >= 6 chunks per thread keeps the shares even
    sh.chunk = chunk;
    return sh;
}

int launch_permanent_batch(const double* mats, const int64_t* matOff, const int32_t* rows, const int32_t* cols,
                           int64_t nMats, int32_t maxDim, double* out, int32_t* status, void* workspace,
                           int64_t workspaceBytes, cudaStream_t stream, int32_t maxSmall, int32_t minSmall) {
    DeviceInfo dev;
    PDA_TRY(current_device_info(&dev));
    const int effDim = std::max(1, std::min<int>(maxDim, PDA_MAX_PERM_DIM));
    // maxSmall / minSmall: bounds on min(rows, cols) over the batch when the caller knows them (else maxDim / 0)
    const int dpBits = std::min(PERM_DP_MAX, std::max(0, std::min<int>(maxSmall < 0 ? maxDim : maxSmall, maxDim)));
    const bool allDp = (maxSmall >= 0 ? maxSmall : maxDim) <= dpBits;   // nothing for the NW walk
    const bool noneDp = minSmall > dpBits;
    const size_t dpSmem = noneDp ? 0 : (size_t)4 * (2 * ((size_t)1 << dpBits) + 16) * sizeof(double);  // 4 warps per finalize CTA
    if (dpSmem > 48 * 1024) PDA_CUDA_TRY(cudaFuncSetAttribute(perm_finalize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dpSmem));
    if (allDp) {
        PermArgs d = {mats, matOff, rows, cols, nMats, 0, 0, 0, 1, 32, 0, effDim, nullptr, dpBits, 0, nullptr, nullptr, nullptr, nullptr, 0};
        perm_finalize_kernel<<<(unsigned)((nMats + 3) / 4), 128, dpSmem, stream>>>(d, out, status, nullptr);
        PDA_CUDA_TRY(cudaGetLastError());
        return PDA_OK;
    }
    // workspace: arrival counters of the fused finalisation, then the per-CTA partial sums
    const int64_t cntBytes = (nMats * 4 + 15) / 16 * 16;
    const int64_t partBytes = workspaceBytes - cntBytes;
    PermShape sh = perm_shape(nMats, effDim, dev.smCount, 0);
    while (sh.ctasPerMat > 1 && (int64_t)sh.ctasPerMat * nMats * (int64_t)sizeof(dd) > partBytes) sh.ctasPerMat >>= 1;
    if ((int64_t)sh.ctasPerMat * nMats * (int64_t)sizeof(dd) > partBytes)
        return fail(PDA_ERR_WORKSPACE, "permanent: workspace of %lld B too small (need %lld)", (long long)workspaceBytes,
                    (long long)(cntBytes + nMats * (int64_t)sizeof(dd)));
    unsigned* done = reinterpret_cast<unsigned*>(workspace);
    if (sh.ctasPerMat > 1) PDA_CUDA_TRY(cudaMemsetAsync(done, 0, (size_t)nMats * 4, stream));
    PermArgs a = {mats, matOff, rows, cols, nMats, 0, 0, 0, sh.chunk, sh.threadsPerMat, sh.ctasPerMat, effDim,
                  reinterpret_cast<dd*>(reinterpret_cast<unsigned char*>(workspace) + cntBytes), noneDp ? -1 : dpBits,
                  1, done, out, status, nullptr, 0};
    // blockIdx.y is limited to 65535: slice the batch
    const int64_t slots = PERM_THREADS / sh.threadsPerMat, perLaunch = 65535 * slots;
    for (int64_t m0 = 0; m0 < nMats; m0 += perLaunch) {
        PermArgs s = a;
        s.matOff = matOff + m0; s.rows = rows + m0; s.cols = cols + m0;
        s.nMats = std::min<int64_t>(perLaunch, nMats - m0);
        s.partial = a.partial + m0 * sh.ctasPerMat;
        s.done = done + m0;
        s.out = out + m0;
        s.status = status ? status + m0 : nullptr;
        PDA_TRY(dispatch(s, stream));
    }
    if (!noneDp) {  // only the short-sided matrices are left for this kernel: everything else finalised itself
        perm_finalize_kernel<<<(unsigned)((nMats + 3) / 4), 128, dpSmem, stream>>>(a, out, status, nullptr);
        PDA_CUDA_TRY(cudaGetLastError());
    }
    return PDA_OK;
}

int launch_permanent_range(const double* A, int32_t n, uint64_t begin, uint64_t end, double* partial,
                           void* workspace, int64_t workspaceBytes, cudaStream_t stream) {
    DeviceInfo dev;
    PDA_TRY(current_device_info(&dev));
    PermShape sh = perm_shape(1, n, dev.smCount, end - begin);
    // keep the ranges uniform inside a warp: the chunk must divide `begin`
    while (sh.chunk > 1 && (begin % (unsigned long long)sh.chunk) != 0) sh.chunk >>= 1;
    if ((int64_t)sh.ctasPerMat * (int64_t)sizeof(dd) + 64 > workspaceBytes) {
        sh.ctasPerMat = (int)std::max<int64_t>(1, (workspaceBytes - 64) / (int64_t)sizeof(dd));
        if (workspaceBytes < 64 + (int64_t)sizeof(dd)) return fail(PDA_ERR_WORKSPACE, "permanent_range: workspace too small");
    }
    // one launch: the matrix is described in the kernel arguments (oneDim), the CTA that arrives last adds the per-CTA
    // partials up; the arrival counter at the head of the workspace is zeroed by a memset node in front of the kernel
    unsigned char* ws = reinterpret_cast<unsigned char*>(workspace);
    unsigned* done = reinterpret_cast<unsigned*>(ws);
    if (sh.ctasPerMat > 1) PDA_CUDA_TRY(cudaMemsetAsync(done, 0, 4, stream));
    PermArgs a = {A, nullptr, nullptr, nullptr, 1, begin, end, 1, sh.chunk, sh.threadsPerMat, sh.ctasPerMat, n,
                  reinterpret_cast<dd*>(ws + 64), -1, 1, done, nullptr, nullptr, partial, n};
    return dispatch(a, stream);
}

}  // namespace pda

using namespace pda;

extern "C" {

int64_t pda_permanent_workspace_bytes(int64_t nMats) {
    const int64_t n = std::max<int64_t>(nMats, 1);
    return 64 + 16 * (n + 8192) + (n * 4 + 15) / 16 * 16;  // per-CTA partials + arrival counters
}

int pda_permanent_batch(const double* mats, const int64_t* matOff, const int32_t* rows, const int32_t* cols,
                        int64_t nMats, int32_t maxDim, double* out, int32_t* status,
                        void* workspace, int64_t workspaceBytes, void* stream) {
    if (nMats < 0) return fail(PDA_ERR_INVALID, "permanent: nMats < 0");
    if (nMats == 0) return PDA_OK;
    if (!mats || !matOff || !rows || !cols || !out || !workspace) return fail(PDA_ERR_INVALID, "permanent: NULL argument");
    if (maxDim < 0) return fail(PDA_ERR_INVALID, "permanent: maxDim < 0");
    // the dimensions live on the device, so this entry cannot tell whether any matrix has a short side: it walks every
    // matrix the reference's way (NW on the ones-padded square).  The *_host entries, which see the dimensions, send
    // short-sided matrices to the cancellation-free subset programme (perm_dp_warp).
    return launch_permanent_batch(mats, matOff, rows, cols, nMats, maxDim, out, status, workspace, workspaceBytes,
                                  reinterpret_cast<cudaStream_t>(stream), -1, PDA_MAX_DIM + 1);
}

int pda_permanent_batch_host(const double* mats, const int64_t* matOff, const int32_t* rows, const int32_t* cols,
                             int64_t nMats, double* out, int32_t* status, int32_t device) {
    if (nMats < 0) return fail(PDA_ERR_INVALID, "permanent: nMats < 0");
    if (nMats == 0) return PDA_OK;
    if (!mats || !matOff || !rows || !cols || !out) return fail(PDA_ERR_INVALID, "permanent: NULL argument");
    size_t nEl = 0, loEl = ~(size_t)0;  // only [loEl, nEl) of `mats` is staged (a shard passes the whole array)
    int maxDim = 0;
    for (int64_t i = 0; i < nMats; ++i) {
        if (rows[i] < 0 || cols[i] < 0 || matOff[i] < 0) return fail(PDA_ERR_INVALID, "permanent: negative dimension or offset");
        nEl = std::max(nEl, (size_t)matOff[i] + (size_t)rows[i] * cols[i]);
        loEl = std::min(loEl, (size_t)matOff[i]);
        maxDim = std::max(maxDim, std::max(rows[i], cols[i]));
    }
    if (loEl > nEl) loEl = nEl;
    DeviceScope scope;
    PDA_TRY(scope.enter(device));
    DeviceInfo dev;
    PDA_TRY(current_device_info(&dev));
    const size_t n = (size_t)nMats;
    const size_t wsBytes = (size_t)pda_permanent_workspace_bytes(nMats);
    Stage st(device);
    const size_t oM = st.reserve((nEl - loEl) * 8), oOff = st.reserve(n * 8), oR = st.reserve(n * 4), oC = st.reserve(n * 4);
    const size_t oOut = st.reserve(n * 8), oSt = st.reserve(n * 4), oWs = st.reserve(wsBytes);
    PDA_TRY(st.commit());
    cudaStream_t s = 0;
    PDA_TRY(h2d(st.at<double>(oM), mats + loEl, nEl - loEl, s));
    PDA_TRY(h2d(st.at<int64_t>(oOff), matOff, n, s));
    PDA_TRY(h2d(st.at<int32_t>(oR), rows, n, s));
    PDA_TRY(h2d(st.at<int32_t>(oC), cols, n, s));
    int maxSmall = 0, minSmall = PDA_MAX_DIM;
    for (int64_t i = 0; i < nMats; ++i) {
        const int sm = std::min(rows[i], cols[i]);
        maxSmall = std::max(maxSmall, sm); minSmall = std::min(minSmall, sm);
    }
    PDA_TRY(launch_permanent_batch(st.at<double>(oM) - loEl, st.at<int64_t>(oOff), st.at<int32_t>(oR), st.at<int32_t>(oC), nMats,
                                   maxDim, st.at<double>(oOut), st.at<int32_t>(oSt), st.at<unsigned char>(oWs),
                                   (int64_t)wsBytes, s, maxSmall, minSmall));
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
    DeviceScope scope;
    PDA_TRY(scope.enter(device));
    DeviceInfo dev;
    PDA_TRY(current_device_info(&dev));
    const size_t wsBytes = 64 + (size_t)16 * 8192;
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

// ---- multi-device forms: one host thread per device, no data-path collective except for the single
// large permanent, whose 16-byte partial sums are gathered by the host and added in device order ----------------------
int pda_permanent_batch_host_multi(const double* mats, const int64_t* matOff, const int32_t* rows, const int32_t* cols,
                                   int64_t nMats, double* out, int32_t* status, const int32_t* devices, int32_t nDevices) {
    if (!devices || nDevices < 1) return fail(PDA_ERR_INVALID, "permanent (multi): need at least one device");
    if (nMats < 0) return fail(PDA_ERR_INVALID, "permanent: nMats < 0");
    if (nMats == 0) return PDA_OK;
    if (!mats || !matOff || !rows || !cols || !out || !status) return fail(PDA_ERR_INVALID, "permanent: NULL argument");
    return run_sharded(nMats, devices, nDevices, [&](int64_t m0, int64_t m1, int dev) {
        return pda_permanent_batch_host(mats, matOff + m0, rows + m0, cols + m0, m1 - m0, out + m0, status + m0, dev);
    });
}

int pda_permanent_sharded_host(const double* A, int32_t n, const int32_t* devices, int32_t nDevices, double* out) {
    if (!A || !out) return fail(PDA_ERR_INVALID, "permanent_sharded: NULL argument");
    if (!devices || nDevices < 1) return fail(PDA_ERR_INVALID, "permanent_sharded: need at least one device");
    if (n < 1 || n > PDA_MAX_PERM_DIM) return fail(PDA_ERR_UNSUPPORTED, "permanent_sharded: n = %d outside 1..%d", n, PDA_MAX_PERM_DIM);
    // Gray index range [0, 2^(n-1)) in a power-of-two number of equal pieces: every boundary is a multiple of a large
    // power of two, which keeps the kernel's column reads warp-uniform (launch_permanent_range)
    const uint64_t total = 1ULL << (n - 1);
    int parts = 1;
    while (parts * 2 <= nDevices && (uint64_t)parts * 2 <= total) parts *= 2;
    std::vector<double> partial((size_t)parts * 2, 0.0);
    PDA_TRY(run_sharded(parts, devices, parts, [&](int64_t r0, int64_t r1, int dev) {
        for (int64_t r = r0; r < r1; ++r) {
            const int rc = pda_permanent_range_host(A, n, total * (uint64_t)r / parts, total * (uint64_t)(r + 1) / parts,
                                                    partial.data() + 2 * r, dev);
            if (rc) return rc;
        }
        return (int)PDA_OK;
    }));
    // (hi, lo) pairs added in device order: error-free two-sum on the high parts -- the result does not depend on timing
    double hi = 0.0, lo = 0.0;
    for (int r = 0; r < parts; ++r) {
        const double h = partial[2 * (size_t)r], l = partial[2 * (size_t)r + 1];
        const double s2 = hi + h, bb = s2 - hi, e = (hi - (s2 - bb)) + (h - bb);
        hi = s2;
        lo += e + l;
    }
    *out = (double)(4 * (n & 1) - 2) * (hi + lo);  // sign and factor 2 of the NW formula (nwPerm.cpp:326)
    return PDA_OK;
}

}  // extern "C"
