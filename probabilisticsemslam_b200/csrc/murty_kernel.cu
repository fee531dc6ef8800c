// murty_kernel.cu -- batched Murty k-best assignment enumeration for sm_100a.
//
// One warp owns one association problem at a time (persistent warps, atomic work
// cursor).  The algorithm is the reference's (shortestPathCPP.cpp): a shortest-
// augmenting-path LAP on the zero-padded square matrix for the root, then Murty's
// partition where every child inherits the parent's duals and needs ONE Dijkstra
// from the freed column (Miller-Stone-Cox).  What is different is where things live:
//
//   lane l owns rows  l, l+32, ...   (v, col4row, shortestPathCost, pred  in registers)
//   lane l owns cols  l, l+32, ...   (u, row4col                            in registers)
//   shared memory per warp: the shifted cost matrix (real columns only -- the
//     reference's padding columns are exactly 0.0 and are never stored), a mirror of
//     u / row4col for the data-dependent lookups, and the weight accumulators
//   global memory per warp: a node arena (bump allocated, written once, read once)
//     and the binary heap that orders the nodes
//
// The row scan of one Dijkstra step is a single pass over the lane's R rows; the
// "closest row" is a warp arg-min done with three REDUX.MIN on an order-preserving
// 64-bit key, lowest row index winning ties -- exactly the reference's first-minimum-
// in-list-order rule because its Row2Scan lists are always ascending
// (shortestPathCPP.cpp:155-157, 486, 215/345).  Every FP64 expression keeps the
// reference's operand order (this file is compiled with -fmad=false; there is no
// multiply on the path except CDelta*numCol), so row4col, col4row, gains and the
// enumeration order are bit-identical to an IEEE-strict build of the reference.
// The heap replays libstdc++'s __push_heap / __adjust_heap so that exact gain ties
// pop in the same order as std::priority_queue<pMurtyHyp> (shortestPathCPP.cpp:30-42).
#include "pda_internal.h"

#include <math_constants.h>

namespace pda {
namespace {

constexpr unsigned FULL = 0xffffffffu;

// resident CTAs per SM the register allocator must leave room for (4 warps each)
#ifndef PDA_MURTY_MINB
#define PDA_MURTY_MINB 6
#endif

struct __align__(16) HeapEntry {
    double gain;
    int node;
    int pad;
};

// ---- order-preserving 64-bit key for doubles ----------------------------------------------------
__device__ __forceinline__ void to_key(double d, unsigned& khi, unsigned& klo) {
    const unsigned hi = (unsigned)__double2hiint(d), lo = (unsigned)__double2loint(d);
    const unsigned m = (unsigned)((int)hi >> 31);  // all ones for negatives
    khi = hi ^ (m | 0x80000000u);
    klo = lo ^ m;
}
__device__ __forceinline__ double from_key(unsigned khi, unsigned klo) {
    const unsigned m = (khi & 0x80000000u) ? 0u : 0xffffffffu;
    return __hiloint2double((int)(khi ^ (m | 0x80000000u)), (int)(klo ^ m));
}
constexpr unsigned KEY_INF_HI = 0xFFF00000u;  // key of +inf is (0xFFF00000, 0)

__device__ __forceinline__ double warp_min(double x) {
#pragma unroll
    for (int o = 16; o; o >>= 1) { double y = __shfl_xor_sync(FULL, x, o); x = (y < x) ? y : x; }
    return x;
}
__device__ __forceinline__ double warp_max(double x) {
#pragma unroll
    for (int o = 16; o; o >>= 1) { double y = __shfl_xor_sync(FULL, x, o); x = (x < y) ? y : x; }
    return x;
}

// Per-warp view of shared memory.
struct WarpSmem {
    double* C;       // [ld * numCol] shifted cost matrix, real columns only
    double* u;       // [32R] mirror of the working node's column duals
    double* spc;     // [32R] shortestPathCost, published after a scan for the dual update
    double* acc;     // [numCol * (nL+1)] weight accumulators
    short* r4c;      // [32R] mirror of the working node's row4col
    short* pred;     // [32R] predecessor column per row
    unsigned short* c4r;  // [32R] mirror of the working node's col4row (0xffff = free)
};

// The working node, distributed over the warp.
template <int R>
struct Node {
    double v[R];   // row duals           (lane owns rows  lane + 32 s)
    double u[R];   // column duals        (lane owns cols  lane + 32 s)
    int c4r[R];    // column of each owned row (-1 = free)
    int r4c[R];    // row of each owned column (-1 = free)
};

// One relaxation pass of the row scan from column `cur` (shortestPathCPP.cpp:179-195 / 307-325).
// Rows that must not take part carry v == -inf, which makes their reduced cost +inf, so no row mask is
// needed: `t < cand` is simply never true for them.  REAL = the column exists in sm.C; otherwise it is
// one of the reference's zero padding columns, C == +0.0 and delta + 0.0 == delta (delta is never -0.0:
// the staged matrix holds no -0.0 and x - x rounds to +0.0).
template <int R, bool REAL>
__device__ __forceinline__ void relax(const double* __restrict__ Ccol, const int n, const double delta,
                                      const double ucur, const double (&v)[R], const int cur,
                                      double (&cand)[R], int (&pred)[R], const int lane) {
    const double du = delta - ucur;  // used by the padding-column form only
#pragma unroll
    for (int s = 0; s < R; ++s) {
        double t;
        if (REAL) {
            const double c = (lane + 32 * s < n) ? Ccol[lane + 32 * s] : 0.0;
            t = ((delta + c) - ucur) - v[s];
        } else {
            t = du - v[s];
        }
        const bool better = t < cand[s];
        cand[s] = better ? t : cand[s];
        pred[s] = better ? cur : pred[s];
    }
}

constexpr int NAN_HI = 0x7ff80000;  // high word of the NaN that marks a scanned row's candidate

// Fast-forward over the reference's no-op hops.
//
// Most Dijkstra steps of a child solve (93 % on the benchmark shapes) go through rows that are paired with
// zero-cost PADDING columns and change nothing: scanning such a column p from row r offers every other row
// t = (cand[r] - u[p]) - v[row], which is not below what the row already holds, so the reference merely retires
// r and moves to the next-closest row.  This routine proves that for a whole run of such rows at once and
// retires them together, with results bit-identical to stepping through them:
//   stopper  = the closest live row that is NOT paired with a padding column (a free row -- the sink -- or a row
//              whose column has real costs); key order is (cand, row), the reference's first-minimum order
//   F        = live rows paired with padding columns that come before the stopper in that order
//   W        = min over F of fl(cand[r] - u[col(r)])   (what each of those hops would offer, before the row dual)
//   test     : fl(W - v[row]) >= cand[row] for EVERY live row.  Rounding is monotone, so fl(W - v) is the smallest
//              offer any hop of F could make to that row; if even that does not beat its candidate, no hop of F
//              updates anything, in any order (a sufficient condition -- it also covers offers from hops that
//              come after the row, which the reference never makes).
// If the test passes every row of F is scanned at its current candidate (parked in sm.spc), predecessors stay as
// they are, and the stopper is the next row to scan: (closest, delta) are returned so the caller skips its own
// arg-min.  If it fails nothing is changed and the caller steps normally.  Returns 0 = not applied,
// 1 = applied, 2 = applied and nothing finite is left (infeasible).
template <int R>
__device__ __forceinline__ int fast_forward(const int numColReal, const WarpSmem& sm, const Node<R>& nd,
                                            const double (&vEff)[R], const double (&uRow)[R], double (&cand)[R],
                                            int& closest, double& delta, const int lane) {
    // the stopper
    double sb = CUDART_INF;
    int sbs = 0;
#pragma unroll
    for (int s = 0; s < R; ++s)
        if (nd.c4r[s] < numColReal && cand[s] < sb) { sb = cand[s]; sbs = s; }
    unsigned khi, klo;
    to_key(sb, khi, klo);
    const unsigned mhi = __reduce_min_sync(FULL, khi);
    const unsigned mlo = __reduce_min_sync(FULL, (khi == mhi) ? klo : 0xffffffffu);
    const bool win = (khi == mhi) && (klo == mlo);
    const int rT = (int)__reduce_min_sync(FULL, win ? (unsigned)(lane + 32 * sbs) : 0xffffu);
    const double kT = from_key(mhi, mlo);
    // F and what its hops would offer
    unsigned inF = 0u;
    double wmin = CUDART_INF;
    int nF = 0;
#pragma unroll
    for (int s = 0; s < R; ++s) {
        const bool f = nd.c4r[s] >= numColReal && cand[s] < CUDART_INF &&
                       (cand[s] < kT || (cand[s] == kT && lane + 32 * s < rT));
        if (f) {
            inF |= 1u << s;
            const double w = cand[s] - uRow[s];
            wmin = (w < wmin) ? w : wmin;
        }
        nF += __popc(__ballot_sync(FULL, f));
    }
    if (nF < 2) return 0;
    to_key(wmin, khi, klo);
    const unsigned whi = __reduce_min_sync(FULL, khi);
    const unsigned wlo = __reduce_min_sync(FULL, (khi == whi) ? klo : 0xffffffffu);
    const double W = from_key(whi, wlo);
    bool beats = false;
#pragma unroll
    for (int s = 0; s < R; ++s) beats = beats || ((W - vEff[s]) < cand[s]);
    if (__any_sync(FULL, beats)) return 0;
#pragma unroll
    for (int s = 0; s < R; ++s) {
        if ((inF >> s) & 1u) {
            sm.spc[lane + 32 * s] = cand[s];
            cand[s] = __hiloint2double(NAN_HI, __double2loint(cand[s]));
        }
    }
    closest = rT;
    delta = kT;
    return (mhi >= KEY_INF_HI) ? 2 : 1;
}

// One shortest augmenting path from `startCol` over the rows flagged in scanBits
// (bit s = row lane+32s), then the dual update and the flip along the path.
//   shortestPathCPP.cpp:168-226 / 297-356 (scan), :92-106 (duals), :108-116 (flip).
// forbBits hides rows on the first hop only (:310).  numColReal = columns that exist in
// sm.C; columns beyond are the reference's zero padding.  Returns true if infeasible.
//
// cand[s] is the reference's shortestPathCost of a row that is still to be scanned.  Rows outside the
// scan set keep cand == +inf (their v is -inf, so they never relax); a row that HAS been scanned gets
// its cand poisoned to NaN (one high-word write): `t < NaN` is false, so it never relaxes again, the
// arg-min skips it, and "scanned" can be read back from it afterwards.  The cost at which a row was
// scanned (== delta at that moment) is parked in sm.spc by lane 0, where the dual update reads it.
template <int R>
__device__ __forceinline__ bool augment_from(const int startCol, const int numColReal, const int ld,
                                             const WarpSmem& sm, Node<R>& nd, const unsigned scanBits,
                                             const unsigned forbBits, const int lane) {
    double cand[R], vEff[R];
    int pred[R];
#pragma unroll
    for (int s = 0; s < R; ++s) {
        cand[s] = CUDART_INF;
        pred[s] = 0;
        vEff[s] = ((scanBits >> s) & 1u) ? nd.v[s] : -CUDART_INF;
    }
    int cur = startCol, sink;
    double delta = 0.0;
    {   // first hop: the forbidden rows sit this one out (:310)
        double vHop[R];
#pragma unroll
        for (int s = 0; s < R; ++s) vHop[s] = ((forbBits >> s) & 1u) ? -CUDART_INF : vEff[s];
        const double ucur = sm.u[cur];
        if (cur < numColReal) relax<R, true>(sm.C + cur * ld, ld, delta, ucur, vHop, cur, cand, pred, lane);
        else relax<R, false>(nullptr, ld, delta, ucur, vHop, cur, cand, pred, lane);
    }
    // u of the column each owned row is paired with (only rows paired with padding columns use it)
    double uRow[R];
#pragma unroll
    for (int s = 0; s < R; ++s) uRow[s] = (nd.c4r[s] >= 0) ? sm.u[nd.c4r[s]] : 0.0;
    bool padPrev = cur >= numColReal;  // the last relaxation came from a padding column
    for (;;) {
        int closest = 0;
        int ff = 0;
        if (padPrev) ff = fast_forward<R>(numColReal, sm, nd, vEff, uRow, cand, closest, delta, lane);
        if (ff == 2) return true;
        if (ff == 0) {
            // lane-local first minimum (lower slot = lower row wins ties; NaN = already scanned), then the warp arg-min
            double best = cand[0];
            int bs = 0;
#pragma unroll
            for (int s = 1; s < R; ++s) if (cand[s] < best || best != best) { best = cand[s]; bs = s; }
            unsigned khi, klo;
            to_key(best, khi, klo);
            const unsigned mhi = __reduce_min_sync(FULL, khi);
            if (mhi >= KEY_INF_HI) return true;  // minVal == +inf (:197, :327): nothing finite is left
            const unsigned mlo = __reduce_min_sync(FULL, (khi == mhi) ? klo : 0xffffffffu);
            const bool win = (khi == mhi) && (klo == mlo);
            closest = (int)__reduce_min_sync(FULL, win ? (unsigned)(lane + 32 * bs) : 0xffffu);
            delta = from_key(mhi, mlo);
        }
        if (lane == 0) sm.spc[closest] = delta;
#pragma unroll
        for (int s = 0; s < R; ++s)
            if (lane + 32 * s == closest) cand[s] = __hiloint2double(NAN_HI, __double2loint(cand[s]));
        const unsigned next = sm.c4r[closest];
        if (next == 0xffffu) { sink = closest; break; }
        cur = (int)next;
        padPrev = cur >= numColReal;
        const double ucur = sm.u[cur];
        if (padPrev) relax<R, false>(nullptr, ld, delta, ucur, vEff, cur, cand, pred, lane);
        else relax<R, true>(sm.C + cur * ld, ld, delta, ucur, vEff, cur, cand, pred, lane);
    }

    // duals, using row4col as it was before the flip (:92-106).  A column other than startCol was scanned
    // exactly when the row it is paired with was scanned (the sink row is unpaired).
    unsigned rowsDone[R];
#pragma unroll
    for (int s = 0; s < R; ++s) {
        rowsDone[s] = __ballot_sync(FULL, __double2hiint(cand[s]) == NAN_HI);
        sm.pred[lane + 32 * s] = (short)pred[s];
    }
    __syncwarp();
#pragma unroll
    for (int s = 0; s < R; ++s) {
        if ((rowsDone[s] >> lane) & 1u) nd.v[s] = (nd.v[s] - delta) + sm.spc[lane + 32 * s];
        const int c = lane + 32 * s, r = nd.r4c[s];
        bool seen = false;
        if (r >= 0) {
            unsigned w = rowsDone[0];
#pragma unroll
            for (int q = 1; q < R; ++q) if ((r >> 5) == q) w = rowsDone[q];
            seen = (w >> (r & 31)) & 1u;
        }
        if (c == startCol) { nd.u[s] = nd.u[s] + delta; sm.u[c] = nd.u[s]; }
        else if (seen) { nd.u[s] = (nd.u[s] + delta) - sm.spc[r]; sm.u[c] = nd.u[s]; }
    }
    // flip along the predecessor chain (:108-116); sm.r4c still holds the pre-flip pairing
    int r = sink, c;
    do {
        c = sm.pred[r];
        const int h = sm.r4c[c];
#pragma unroll
        for (int s = 0; s < R; ++s) {
            if (lane + 32 * s == r) nd.c4r[s] = c;
            if (lane + 32 * s == c) nd.r4c[s] = r;
        }
        r = h;
    } while (c != startCol);
    __syncwarp();
#pragma unroll
    for (int s = 0; s < R; ++s) {
        sm.r4c[lane + 32 * s] = (short)nd.r4c[s];
        sm.c4r[lane + 32 * s] = (unsigned short)nd.c4r[s];
    }
    __syncwarp();
    return false;
}

// calcGain (:59-80): ascending column order, starting from 0.0.
__device__ __forceinline__ double path_gain(const WarpSmem& sm, const int ld, const int numColGain) {
    double g = 0.0;
#pragma unroll 4
    for (int c = 0; c < numColGain; ++c) g = g + sm.C[c * ld + sm.r4c[c]];
    return g;
}

// ---- the heap: lane 0 only ----------------------------------------------------------------------------
// Entries are 16 bytes and move as one 128-bit access.  The first `topCap` entries (the top levels, which every
// pop walks through) live in the warp's shared memory, the rest in its global arena: a pop's sift-down is a chain
// of dependent loads, and this turns most of its ~10 L2 round trips into shared-memory reads.
struct Heap {
    HeapEntry* top;   // shared memory, entries [0, topCap)
    HeapEntry* deep;  // global arena, entry i at deep[i] (slots below topCap unused)
    int topCap;
    __device__ __forceinline__ HeapEntry get(int i) const { return (i < topCap) ? top[i] : deep[i]; }
    __device__ __forceinline__ void put(int i, const HeapEntry& e) const {
        if (i < topCap) top[i] = e; else deep[i] = e;
    }
};

__device__ __forceinline__ void heap_sift_up(const Heap& h, int hole, const HeapEntry val) {
    while (hole > 0) {
        const int parent = (hole - 1) / 2;
        const HeapEntry par = h.get(parent);
        if (!(par.gain > val.gain)) break;
        h.put(hole, par);
        hole = parent;
    }
    h.put(hole, val);
}
__device__ __forceinline__ void heap_pop(const Heap& h, const int lenBefore) {
    if (lenBefore > 1) {
        const int len = lenBefore - 1;
        const HeapEntry val = h.get(len);
        int hole = 0, child = 0;
        while (child < (len - 1) / 2) {
            child = 2 * (child + 1);
            const HeapEntry right = h.get(child), left = h.get(child - 1);  // both children in one round trip
            const bool takeLeft = right.gain > left.gain;                     // right child wins an exact tie
            if (takeLeft) child--;
            h.put(hole, takeLeft ? left : right);
            hole = child;
        }
        if ((len & 1) == 0 && child == (len - 2) / 2) {
            child = 2 * (child + 1);
            h.put(hole, h.get(child - 1));
            hole = child - 1;
        }
        heap_sift_up(h, hole, val);
    }
}

// ---- node arena -------------------------------------------------------------------------------------
// layout of one stored node (D = geo.nodeDim):  v[D] | u[D] | c4r bytes[D] | r4c bytes[D] | forb words[R] | activeCol
template <int R>
__device__ __forceinline__ void node_store(unsigned char* base, const int D, const int n, const Node<R>& nd,
                                           const unsigned forbBits, const int activeCol, const int lane) {
    double* dv = reinterpret_cast<double*>(base);
    unsigned char* bi = base + 16 * D;
    unsigned* meta = reinterpret_cast<unsigned*>(base + 18 * D);
#pragma unroll
    for (int s = 0; s < R; ++s) {
        const int i = lane + 32 * s;
        if (i < n) {
            dv[i] = nd.v[s];
            dv[D + i] = nd.u[s];
            bi[i] = (unsigned char)nd.c4r[s];
            bi[D + i] = (unsigned char)nd.r4c[s];
        }
        const unsigned w = __ballot_sync(FULL, (forbBits >> s) & 1u);
        if (lane == 0) meta[s] = w;
    }
    if (lane == 0) meta[R] = (unsigned)activeCol;
}
template <int R>
__device__ __forceinline__ void node_load(const unsigned char* base, const int D, const int n, Node<R>& nd,
                                          unsigned& forbBits, int& activeCol, const int lane) {
    const double* dv = reinterpret_cast<const double*>(base);
    const unsigned char* bi = base + 16 * D;
    const unsigned* meta = reinterpret_cast<const unsigned*>(base + 18 * D);
    forbBits = 0u;
#pragma unroll
    for (int s = 0; s < R; ++s) {
        const int i = lane + 32 * s;
        if (i < n) {
            nd.v[s] = dv[i];
            nd.u[s] = dv[D + i];
            nd.c4r[s] = (int)(signed char)bi[i];
            nd.r4c[s] = (int)(signed char)bi[D + i];
        } else {
            nd.v[s] = 0.0; nd.u[s] = 0.0; nd.c4r[s] = -1; nd.r4c[s] = -1;
        }
        forbBits |= ((meta[s] >> lane) & 1u) << s;
    }
    activeCol = (int)meta[R];
}

template <int R>
__device__ __forceinline__ void publish_cols(const WarpSmem& sm, const Node<R>& nd, const int lane) {
    __syncwarp();  // earlier readers of the mirrors (gain, new row) are done before they are overwritten
#pragma unroll
    for (int s = 0; s < R; ++s) {
        sm.u[lane + 32 * s] = nd.u[s];
        sm.r4c[lane + 32 * s] = (short)nd.r4c[s];
        sm.c4r[lane + 32 * s] = (unsigned short)nd.c4r[s];
    }
    __syncwarp();
}

// makeCostMatrixSafe (:534-569): shift so every entry is >= 0; returns the shift.
__device__ __forceinline__ double stage_safe_matrix(const double* Cg, double* Cs, const int numEl,
                                                    const bool maximize, const bool makeSafe, const int lane) {
    if (!makeSafe) {
        for (int i = lane; i < numEl; i += 32) Cs[i] = Cg[i] + 0.0;  // + 0.0: no -0.0 in the staged matrix
        __syncwarp();
        return 0.0;
    }
    double d;
    if (!maximize) {
        d = CUDART_INF;
        for (int i = lane; i < numEl; i += 32) { const double x = Cg[i]; d = (x < d) ? x : d; }
        d = warp_min(d);
        for (int i = lane; i < numEl; i += 32) Cs[i] = (Cg[i] - d) + 0.0;
    } else {
        d = -CUDART_INF;
        for (int i = lane; i < numEl; i += 32) { const double x = Cg[i]; d = (d < x) ? x : d; }
        d = warp_max(d);
        for (int i = lane; i < numEl; i += 32) Cs[i] = (-Cg[i] + d) + 0.0;
    }
    __syncwarp();
    return d;
}

template <int R>
__device__ __forceinline__ void emit(const MurtyArgs& a, const long long p, const int slot, const int n, const int nc,
                                     const Node<R>& nd, const double gainOut, const int lane) {
    if (a.c4rBest) {
        int64_t* o = a.c4rBest + a.c4rOff[p] + (int64_t)slot * n;
#pragma unroll
        for (int s = 0; s < R; ++s) if (lane + 32 * s < n) o[lane + 32 * s] = (int64_t)nd.c4r[s];
    }
    if (a.r4cBest) {
        int64_t* o = a.r4cBest + a.r4cOff[p] + (int64_t)slot * nc;
#pragma unroll
        for (int s = 0; s < R; ++s) if (lane + 32 * s < nc) o[lane + 32 * s] = (int64_t)nd.r4c[s];
    }
    if (a.gainBest && lane == 0) a.gainBest[p * (long long)a.k + slot] = gainOut;
}

// assignmentProb / bruteForceProb accumulation of one hypothesis (assignment.cpp:620-640, 916-937)
template <int R>
__device__ __forceinline__ void add_weight(const MurtyArgs& a, const WarpSmem& sm, const Node<R>& nd, const int nc,
                                           const int nL, const double best, const double gainOut, double& total,
                                           const int lane) {
    if (a.weightMode == PDA_WEIGHTS_GATED && !(best + a.weightGate > gainOut)) return;
    const double w = exp(best - gainOut);
    total += w;
#pragma unroll
    for (int s = 0; s < R; ++s) {
        const int c = lane + 32 * s;
        if (c < nc) {
            const int to = nd.r4c[s] >= nL ? nL : nd.r4c[s];
            sm.acc[c * (nL + 1) + to] += w;
        }
    }
}

template <int R>
__device__ void solve_problem(const MurtyArgs& a, const long long p, const WarpSmem& sm, const Heap& heap,
                              unsigned char* nodes, const int lane) {
    const int n = a.numRow[p], nc = a.numCol[p];
    const int D = a.geo.nodeDim;
    const bool wantW = a.weightMode != PDA_WEIGHTS_NONE;
    const int nL = wantW ? a.nL[p] : 0;
    if (nc < 1 || nc > n || n > 32 * R || (wantW && nL + nc != n)) {
        if (lane == 0) a.nFound[p] = 0;
        return;
    }
    const double* Cg = a.costs + a.costOff[p];

    // single-detection shortcut of assignmentProb / bruteForceProb (assignment.cpp:554-570, 840-856)
    if (wantW && nc == 1) {
        for (int i = lane; i <= nL; i += 32) sm.acc[i] = (Cg[i] < a.weightGate) ? exp(-Cg[i]) : 0.0;
        __syncwarp();
        double norm = 0.0;
        for (int i = 0; i <= nL; ++i) if (Cg[i] < a.weightGate) norm += sm.acc[i];
        norm = 1.0 / norm;
        double* out = a.probs + a.probOff[p];
        for (int i = lane; i <= nL; i += 32) out[i] = sm.acc[i] * norm;
        __syncwarp();
    }

    const bool maximize = a.maximize != 0;
    double CDelta = stage_safe_matrix(Cg, sm.C, n * nc, maximize, true, lane);
    CDelta = CDelta * (double)nc;  // (:583, :664) a separately rounded product
    if (wantW && nc > 1) {
        for (int i = lane; i < nc * (nL + 1); i += 32) sm.acc[i] = 0.0;
    }

    // ---- root: shortestPathCPP on the zero-padded n x n matrix (:119-238) --------------------
    Node<R> nd;
#pragma unroll
    for (int s = 0; s < R; ++s) { nd.v[s] = 0.0; nd.u[s] = 0.0; nd.c4r[s] = -1; nd.r4c[s] = -1; }
    publish_cols<R>(sm, nd, lane);
    unsigned allRows = 0u;
#pragma unroll
    for (int s = 0; s < R; ++s) if (lane + 32 * s < n) allRows |= 1u << s;
    for (int c = 0; c < n; ++c) {
        if (augment_from<R>(c, nc, n, sm, nd, allRows, 0u, lane)) {
            if (lane == 0) a.nFound[p] = 0;
            if (wantW && nc > 1) {  // the reference ends up scaling zeros by 1/0 here
                double* out = a.probs + a.probOff[p];
                for (int i = lane; i < nc * (nL + 1); i += 32) out[i] = CUDART_NAN;
            }
            return;
        }
    }
    double gain = path_gain(sm, n, nc);
    unsigned forb = 0u;
    {
        const int r0 = sm.r4c[0];
#pragma unroll
        for (int s = 0; s < R; ++s) if (lane + 32 * s == r0) forb |= 1u << s;
    }
    int activeCol = 0;

    double gain0Out, cutoffGain = a.cutoff;
    bool cutMax = a.cutMaximize != 0;
    if (!maximize) {
        if (a.cutMode == PDA_CUT_RELATIVE) { cutoffGain = gain + a.cutoff; cutMax = false; }
        gain0Out = gain + CDelta;
    } else {
        if (a.cutMode == PDA_CUT_RELATIVE) { cutoffGain = gain - a.cutoff; cutMax = true; }
        gain0Out = -gain + CDelta;
    }
    const bool cutting = a.cutMode != PDA_CUT_NONE;
    emit<R>(a, p, 0, n, nc, nd, gain0Out, lane);
    double total = 0.0;
    if (wantW && nc > 1) add_weight<R>(a, sm, nd, nc, nL, gain0Out, gain0Out, total, lane);

    int heapLen = 1;   // the root; its state is live in registers, slot 0 of the arena stays unused
    int nNodes = 1;
    int sweep = 1;
    for (; sweep < a.k; ++sweep) {
        // ---- pop the node whose state we hold (it is the heap top) ---------------------------
        if (lane == 0) heap_pop(heap, heapLen);
        heapLen--;
        __syncwarp();

        // ---- split (:455-532) --------------------------------------------------------------
        Node<R> par = nd;
        const unsigned parForb = forb;
        const int a0 = activeCol;
        unsigned inPar = 0u;  // Row2ScanParent: rows paired with columns >= a0
#pragma unroll
        for (int s = 0; s < R; ++s) if (lane + 32 * s < n && par.c4r[s] >= a0) inPar |= 1u << s;
        for (int c = a0; c < nc; ++c) {
            nd = par;
            publish_cols<R>(sm, nd, lane);
            const int r0 = sm.r4c[c];  // the pairing this child must give up
            unsigned hideFirst = 0u;
#pragma unroll
            for (int s = 0; s < R; ++s) {
                if (lane + 32 * s == r0) { nd.c4r[s] = -1; hideFirst |= 1u << s; }
                if (lane + 32 * s == c) nd.r4c[s] = -1;
            }
            if (c == a0) hideFirst = parForb;  // first child inherits every constraint on the active column (:490)
            if (lane == 0) sm.c4r[r0] = 0xffffu;  // the freed row is the only sink of this search
            __syncwarp();
            const bool infeasible = augment_from<R>(c, nc, n, sm, nd, inPar, hideFirst, lane);
            if (!infeasible) {
                const double g = path_gain(sm, n, nc);
                const bool cut = cutting && (cutMax ? (g < cutoffGain) : (g > cutoffGain));
                if (!cut) {
                    unsigned childForb = hideFirst;
                    const int rNew = sm.r4c[c];
#pragma unroll
                    for (int s = 0; s < R; ++s) if (lane + 32 * s == rNew) childForb |= 1u << s;
                    node_store<R>(nodes + (size_t)nNodes * a.geo.nodeStride, D, n, nd, childForb, c, lane);
                    if (lane == 0) {
                        HeapEntry e;
                        e.gain = g; e.node = nNodes; e.pad = 0;
                        heap_sift_up(heap, heapLen, e);
                    }
                    heapLen++;
                    nNodes++;
                }
            }
#pragma unroll
            for (int s = 0; s < R; ++s) if (lane + 32 * s == r0) inPar &= ~(1u << s);  // column c is fixed from here on
        }
        __syncwarp();
        if (heapLen == 0) break;

        // ---- the new top is hypothesis number `sweep` (:703-719) ---------------------------
        const HeapEntry top = heap.get(0);
        __syncwarp();  // every lane has read the top before lane 0 starts the next pop
        gain = top.gain;
        node_load<R>(nodes + (size_t)top.node * a.geo.nodeStride, D, n, nd, forb, activeCol, lane);
        double gainOut;
        bool stop = false;
        if (!maximize) {
            gainOut = gain + CDelta;
            if (a.cutMode == PDA_CUT_RELATIVE && gainOut > gain0Out + a.cutoff) stop = true;
        } else {
            gainOut = -gain + CDelta;
            if (a.cutMode == PDA_CUT_RELATIVE && gainOut < gain0Out - a.cutoff) stop = true;
        }
        emit<R>(a, p, sweep, n, nc, nd, gainOut, lane);
        if (stop) break;
        if (wantW && nc > 1) add_weight<R>(a, sm, nd, nc, nL, gain0Out, gainOut, total, lane);
    }
    if (lane == 0) a.nFound[p] = sweep;
    if (wantW && nc > 1) {
        __syncwarp();
        const double norm = 1.0 / total;
        double* out = a.probs + a.probOff[p];
        for (int i = lane; i < nc * (nL + 1); i += 32) out[i] = sm.acc[i] * norm;
        __syncwarp();
    }
}

__device__ __forceinline__ WarpSmem carve(unsigned char* base, const MurtyGeometry& g) {
    WarpSmem sm;
    const int D = 32 * g.R;
    sm.C = reinterpret_cast<double*>(base);
    sm.u = sm.C + g.cCap;
    sm.spc = sm.u + D;
    sm.acc = sm.spc + D;
    sm.r4c = reinterpret_cast<short*>(sm.acc + g.pCap);
    sm.pred = sm.r4c + D;
    sm.c4r = reinterpret_cast<unsigned short*>(sm.pred + D);
    return sm;
}

template <int R>
__global__ void __launch_bounds__(128, PDA_MURTY_MINB) murty_kernel(const MurtyArgs a) {
    extern __shared__ __align__(16) unsigned char smemRaw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gw = blockIdx.x * (blockDim.x >> 5) + warp;
    if (gw >= a.nWarps) return;
    const WarpSmem sm = carve(smemRaw + (size_t)warp * a.geo.smemPerWarp, a.geo);
    unsigned char* arena = a.arena + (size_t)gw * a.geo.arenaStride;
    Heap heap;
    heap.deep = reinterpret_cast<HeapEntry*>(arena);
    heap.top = reinterpret_cast<HeapEntry*>(smemRaw + (size_t)warp * a.geo.smemPerWarp + a.geo.heapTopOff);
    heap.topCap = a.geo.heapTopCap;
    unsigned char* nodes = arena + a.geo.heapBytes;
    for (;;) {
        unsigned long long p = 0;
        if (lane == 0) p = atomicAdd(a.cursor, 1ULL);
        p = __shfl_sync(FULL, p, 0);
        if ((long long)p >= a.nProblems) break;
        if (a.order) p = (unsigned long long)a.order[p];  // most expensive problems first: a short tail
        solve_problem<R>(a, (long long)p, sm, heap, nodes, lane);
        __syncwarp();
    }
}

// Longest-processing-time-first order for the work cursor.  A problem's cost grows with its number of detections
// (children per pop), so a counting sort by numCol, descending, is enough: the persistent warps then finish on the
// cheapest problems and run dry almost together.  One CTA; ~30 us for 100 000 problems.
__global__ void order_by_cost_kernel(const int32_t* __restrict__ numCol, const long long n, int32_t* __restrict__ order) {
    __shared__ unsigned bucket[PDA_MAX_DIM + 2];
    for (int i = threadIdx.x; i < PDA_MAX_DIM + 2; i += blockDim.x) bucket[i] = 0u;
    __syncthreads();
    for (long long p = threadIdx.x; p < n; p += blockDim.x) {
        int key = numCol[p];
        key = key < 0 ? 0 : (key > PDA_MAX_DIM ? PDA_MAX_DIM : key);
        atomicAdd(&bucket[key], 1u);
    }
    __syncthreads();
    if (threadIdx.x == 0) {  // exclusive prefix over keys in DESCENDING order
        unsigned run = 0;
        for (int key = PDA_MAX_DIM; key >= 0; --key) { const unsigned c = bucket[key]; bucket[key] = run; run += c; }
    }
    __syncthreads();
    for (long long p = threadIdx.x; p < n; p += blockDim.x) {
        int key = numCol[p];
        key = key < 0 ? 0 : (key > PDA_MAX_DIM ? PDA_MAX_DIM : key);
        order[atomicAdd(&bucket[key], 1u)] = (int32_t)p;
    }
}

// ---- plain LAP (assign2D / shortestPathCPP on a rectangular matrix, no padding) ---------------------
template <int R>
__global__ void lap_kernel(const LapArgs a, const int smemPerWarp, const int cCap) {
    extern __shared__ __align__(16) unsigned char smemRaw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long p = (long long)blockIdx.x * (blockDim.x >> 5) + warp;
    if (p >= a.nProblems) return;
    MurtyGeometry g;
    g.R = R; g.cCap = cCap; g.pCap = 0;
    const WarpSmem sm = carve(smemRaw + (size_t)warp * smemPerWarp, g);
    const int n = a.numRow[p], nc = a.numCol[p];
    const int ncGain = a.numCol4Gain ? a.numCol4Gain[p] : nc;
    if (nc < 0 || nc > n || n > 32 * R) { if (lane == 0 && a.feasible) a.feasible[p] = 0; return; }
    double CDelta = stage_safe_matrix(a.costs + a.costOff[p], sm.C, n * nc, a.maximize != 0, a.makeSafe != 0, lane);
    CDelta = CDelta * (double)nc;
    Node<R> nd;
#pragma unroll
    for (int s = 0; s < R; ++s) { nd.v[s] = 0.0; nd.u[s] = 0.0; nd.c4r[s] = -1; nd.r4c[s] = -1; }
    publish_cols<R>(sm, nd, lane);
    unsigned allRows = 0u;
#pragma unroll
    for (int s = 0; s < R; ++s) if (lane + 32 * s < n) allRows |= 1u << s;
    bool infeasible = false;
    for (int c = 0; c < nc && !infeasible; ++c) infeasible = augment_from<R>(c, nc, n, sm, nd, allRows, 0u, lane);
    double gain = -1.0;
    int r0 = -1;
    if (!infeasible) {
        gain = path_gain(sm, n, ncGain);
        if (a.makeSafe) gain = a.maximize ? (-gain + CDelta) : (gain + CDelta);
        if (nc > 0) r0 = sm.r4c[0];
    }
    const int64_t ro = a.rowOff[p], co = a.colOff[p];
#pragma unroll
    for (int s = 0; s < R; ++s) {
        const int i = lane + 32 * s;
        if (i < n) {
            if (a.col4row) a.col4row[ro + i] = nd.c4r[s];
            if (a.v) a.v[ro + i] = nd.v[s];
            if (a.forbidden) a.forbidden[ro + i] = (i == r0) ? 1 : 0;
        }
        if (i < nc) {
            if (a.row4col) a.row4col[co + i] = nd.r4c[s];
            if (a.u) a.u[co + i] = nd.u[s];
        }
    }
    if (lane == 0) {
        if (a.gain) a.gain[p] = gain;
        if (a.feasible) a.feasible[p] = infeasible ? 0 : 1;
    }
}

}  // namespace

// ---- host side ------------------------------------------------------------------------------------
static int round_up(int x, int m) { return (x + m - 1) / m * m; }

int murty_geometry(int32_t k, int32_t maxNumRow, int32_t maxNumCol, bool weights, const DeviceInfo& dev,
                   MurtyGeometry* g) {
    if (k < 1 || maxNumRow < 1 || maxNumCol < 1 || maxNumCol > maxNumRow)
        return fail(PDA_ERR_INVALID, "murty: need k >= 1 and 1 <= maxNumCol <= maxNumRow (got k=%d, %d x %d)", k, maxNumRow, maxNumCol);
    if (maxNumRow > PDA_MAX_DIM)
        return fail(PDA_ERR_UNSUPPORTED, "murty: numRow %d exceeds PDA_MAX_DIM %d", maxNumRow, PDA_MAX_DIM);
    g->R = (maxNumRow + 31) / 32;
    if (g->R == 3) g->R = 4;
    const int D = 32 * g->R;
    g->nodeDim = round_up(maxNumRow, 8);
    g->nodeStride = round_up(18 * g->nodeDim + 4 * (g->R + 1), 16);
    const int64_t nodes = 1 + (int64_t)(k - 1) * maxNumCol;  // every pop creates at most numCol children
    if (nodes > (int64_t)1 << 30) return fail(PDA_ERR_UNSUPPORTED, "murty: k * numCol too large");
    g->maxNodes = (int)nodes;
    g->heapBytes = (int64_t)round_up((int)nodes, 8) * (int64_t)sizeof(HeapEntry);
    g->arenaStride = (g->heapBytes + nodes * g->nodeStride + 255) / 256 * 256;
    g->cCap = round_up(maxNumRow * maxNumCol, 2);
    g->pCap = weights ? round_up(maxNumCol * maxNumRow, 2) : 0;
    const int baseSmem = round_up(8 * (g->cCap + g->pCap + 2 * D) + 3 * 2 * D, 16);
    // leftover shared memory (at the 24 warps per SM the register budget allows) holds the top of the heap
    const int budget = (227 * 1024 - 6 * 1024) / 24;
    int topCap = budget > baseSmem ? (budget - baseSmem) / (int)sizeof(HeapEntry) : 0;
    if (topCap > g->maxNodes) topCap = g->maxNodes;
    if (topCap < 3) topCap = 0;
    g->heapTopOff = baseSmem;
    g->heapTopCap = topCap;
    g->smemPerWarp = baseSmem + topCap * (int)sizeof(HeapEntry);
    if (g->smemPerWarp > dev.maxSmemOptin)
        return fail(PDA_ERR_UNSUPPORTED, "murty: a %d x %d problem needs %d B of shared memory per warp (limit %d)",
                    maxNumRow, maxNumCol, g->smemPerWarp, dev.maxSmemOptin);
    int wpc = 4;
    while (wpc > 1 && wpc * g->smemPerWarp > dev.maxSmemOptin) wpc >>= 1;
    g->warpsPerCta = wpc;
    // resident CTAs per SM, as the occupancy calculator sees this instantiation (registers, shared memory)
    int occ = 0;
    const int threads = 32 * wpc;
    const size_t smem = (size_t)wpc * g->smemPerWarp;
    cudaError_t e = cudaSuccess;
    switch (g->R) {
        case 1:
            e = cudaFuncSetAttribute(murty_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, murty_kernel<1>, threads, smem);
            break;
        case 2:
            e = cudaFuncSetAttribute(murty_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, murty_kernel<2>, threads, smem);
            break;
        default:
            e = cudaFuncSetAttribute(murty_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, murty_kernel<4>, threads, smem);
            break;
    }
    if (e != cudaSuccess) return cuda_fail(e, "occupancy query for murty_kernel");
    g->ctasPerSm = occ < 1 ? 1 : occ;
    return PDA_OK;
}

template <int R>
static int launch_murty_r(const MurtyArgs& a, cudaStream_t stream) {
    const int threads = 32 * a.geo.warpsPerCta;
    const int smem = a.geo.warpsPerCta * a.geo.smemPerWarp;
    PDA_CUDA_TRY(cudaFuncSetAttribute(murty_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const int ctas = (a.nWarps + a.geo.warpsPerCta - 1) / a.geo.warpsPerCta;
    murty_kernel<R><<<ctas, threads, smem, stream>>>(a);
    PDA_CUDA_TRY(cudaGetLastError());
    return PDA_OK;
}

int launch_murty(const MurtyArgs& a, cudaStream_t stream) {
    PDA_CUDA_TRY(cudaMemsetAsync(a.cursor, 0, sizeof(unsigned long long), stream));
    if (a.order) {
        order_by_cost_kernel<<<1, 1024, 0, stream>>>(a.numCol, a.nProblems, a.order);
        PDA_CUDA_TRY(cudaGetLastError());
    }
    switch (a.geo.R) {
        case 1: return launch_murty_r<1>(a, stream);
        case 2: return launch_murty_r<2>(a, stream);
        case 4: return launch_murty_r<4>(a, stream);
    }
    return fail(PDA_ERR_UNSUPPORTED, "murty: unsupported row-slot count %d", a.geo.R);
}

template <int R>
static int launch_lap_r(const LapArgs& a, cudaStream_t stream, const DeviceInfo& dev) {
    const int D = 32 * R;
    const int cCap = round_up(a.maxNumRow * a.maxNumCol, 2);
    const int smemPerWarp = round_up(8 * (cCap + 2 * D) + 6 * D, 16);
    if (smemPerWarp > dev.maxSmemOptin) return fail(PDA_ERR_UNSUPPORTED, "lap: matrix too large for shared memory");
    int wpc = 4;
    while (wpc > 1 && wpc * smemPerWarp > dev.maxSmemOptin) wpc >>= 1;
    const int smem = wpc * smemPerWarp;
    PDA_CUDA_TRY(cudaFuncSetAttribute(lap_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const long long ctas = (a.nProblems + wpc - 1) / wpc;
    lap_kernel<R><<<(unsigned)ctas, 32 * wpc, smem, stream>>>(a, smemPerWarp, cCap);
    PDA_CUDA_TRY(cudaGetLastError());
    return PDA_OK;
}

int launch_lap(const LapArgs& a, cudaStream_t stream) {
    DeviceInfo dev;
    int rc = current_device_info(&dev);
    if (rc) return rc;
    if (a.maxNumRow > PDA_MAX_DIM) return fail(PDA_ERR_UNSUPPORTED, "lap: numRow %d exceeds PDA_MAX_DIM", a.maxNumRow);
    const int R = (a.maxNumRow + 31) / 32;
    if (R <= 1) return launch_lap_r<1>(a, stream, dev);
    if (R == 2) return launch_lap_r<2>(a, stream, dev);
    return launch_lap_r<4>(a, stream, dev);
}

}  // namespace pda
