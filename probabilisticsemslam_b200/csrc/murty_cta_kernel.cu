// murty_cta_kernel.cu -- Murty k-best for FEW problems: one CTA (16 warps) per problem.
//
// The throughput kernel (murty_kernel.cu) gives a problem to a single warp; that is the right shape
// for 100 000 problems, but a single problem -- what the SLAM loop submits once per frame
// (system.cpp:268, slidingWindow.cpp:339) and what compMethods times (comparison.cpp:194-222) --
// then runs at the latency of one warp: ~5 us per sweep, ~1 ms at k = 200.  This kernel keeps the
// same arithmetic (the same augment_from / heap / node format from murty_device.cuh, so results are
// bit-identical to the warp kernel and to the reference) but runs the k-best loop as
// out-of-order execution with in-order commit:
//
//   * a node's split (shortestPathCPP.cpp:455-532) depends only on the node itself, so it can be done
//     BEFORE the node reaches the top of the queue.  Every round, warp 0 picks the cheapest not-yet-split
//     nodes near the top of the heap (the top itself first) and turns each (node, child column) pair
//     into a task; the 16 warps solve the tasks in parallel (one Dijkstra each) and leave the children's
//     gains in a split record;
//   * warp 0 then replays the reference's loop strictly in order -- pop, push the popped node's
//     children in column order, read the new top -- for as long as the top's split record exists.  The
//     heap therefore sees exactly the reference's sequence of push/pop operations, including the
//     order among exactly equal gains.
//   On the KITTI-shaped problems ~97 % of the speculated splits are consumed (a node close to the top
//   is almost always popped within the next few sweeps), and 199 sweeps take ~30 rounds.
//   * lists and weights are written at the end, by all warps, from the recorded pop order; weights are
//     accumulated per table entry in hypothesis order, i.e. in the reference's order (assignment.cpp:620-640).
#include "murty_device.cuh"

namespace pda {
namespace {

constexpr int CTA_WARPS = 16;
constexpr int CTA_RECORDS = 64;             // split records (splits done but not yet committed)
constexpr int CTA_MAXTASKS = 2 * CTA_WARPS;  // child solves per round
constexpr int CTA_SPEC = 10;                 // nodes split per round, at most

struct Task {
    int parent;  // node whose split this child belongs to
    int child;   // arena slot for the child
    short c;     // the child's active column
    short rec;   // split record
};

struct CtaCtl {
    double CDelta, gain0Out, cutoffGain;
    unsigned long long freeRec;  // bit r set = record r is free
    long long problem;
    int heapLen, sweep, nNodes, nTasks, done, nFound, nEmit, uncommitted, cutMax, cutting, feasible;
};

// HeapEntry::pad of this kernel: bits 0-7 activeCol, bits 8-23 split record + 1 (0 = not split yet)
__device__ __forceinline__ int pad_active(int pad) { return pad & 0xff; }
__device__ __forceinline__ int pad_record(int pad) { return (pad >> 8) & 0xffff; }

struct CtaSmem {
    double* C;
    double* acc;
    unsigned char* mirrors;  // CTA_WARPS x mirrorBytes
    double* recGain;         // [CTA_RECORDS][PDA_CTA_MAX_COL]
    int* recBase;            // [CTA_RECORDS]
    Task* tasks;             // [CTA_MAXTASKS]
    CtaCtl* ctl;
    HeapEntry* heapTop;
};

__device__ __forceinline__ CtaSmem carve_cta(unsigned char* base, const MurtyGeometry& g, const CtaGeometry& cg) {
    CtaSmem s;
    s.C = reinterpret_cast<double*>(base);
    s.acc = s.C + g.cCap;
    s.mirrors = reinterpret_cast<unsigned char*>(s.acc + g.pCap);
    s.recGain = reinterpret_cast<double*>(s.mirrors + (size_t)CTA_WARPS * cg.mirrorBytes);
    s.recBase = reinterpret_cast<int*>(s.recGain + CTA_RECORDS * PDA_CTA_MAX_COL);
    s.tasks = reinterpret_cast<Task*>(s.recBase + CTA_RECORDS);
    s.ctl = reinterpret_cast<CtaCtl*>(base + cg.ctlOff);
    s.heapTop = reinterpret_cast<HeapEntry*>(base + cg.heapTopOff);
    return s;
}

__device__ __forceinline__ WarpSmem warp_view(const CtaSmem& s, const CtaGeometry& cg, const int R, const int warp) {
    WarpSmem sm;
    const int D = 32 * R;
    unsigned char* m = s.mirrors + (size_t)warp * cg.mirrorBytes;
    sm.C = s.C;
    sm.acc = s.acc;
    sm.u = reinterpret_cast<double*>(m);
    sm.spc = sm.u + D;
    sm.r4c = reinterpret_cast<short*>(sm.spc + D);
    sm.pred = sm.r4c + D;
    sm.c4r = reinterpret_cast<unsigned short*>(sm.pred + D);
    return sm;
}

// per-problem global scratch behind the heap
struct CtaArena {
    HeapEntry* heapDeep;
    double* orderGain;     // [k] reported gain of hypothesis i
    double* orderW;        // [k] its weight
    int* orderNode;        // [k] arena slot holding hypothesis i
    unsigned char* hypRows;  // [k][PDA_CTA_MAX_COL] row4col of hypothesis i, one byte each
    unsigned char* nodes;
};
__device__ __forceinline__ CtaArena carve_arena(unsigned char* base, const CtaGeometry& cg, const int k) {
    CtaArena A;
    A.heapDeep = reinterpret_cast<HeapEntry*>(base);
    A.orderGain = reinterpret_cast<double*>(base + cg.heapBytes);
    A.orderW = A.orderGain + k;
    A.orderNode = reinterpret_cast<int*>(A.orderW + k);
    A.hypRows = reinterpret_cast<unsigned char*>(A.orderNode + ((k + 3) & ~3));
    A.nodes = base + cg.nodesOff;
    return A;
}

// ---- warp 0: staging, single-detection shortcut, root LAP (shortestPathCPP.cpp:119-238) ---------------
template <int R>
__device__ void root_phase(const MurtyArgs& a, const long long p, const CtaSmem& S, const WarpSmem& sm, const Heap& heap,
                           const CtaArena& A, const int lane) {
    CtaCtl* ctl = S.ctl;
    const int n = a.numRow[p], nc = a.numCol[p];
    const int D = a.geo.nodeDim;
    const bool wantW = a.weightMode != PDA_WEIGHTS_NONE;
    const int nL = wantW ? a.nL[p] : 0;
    const double* Cg = a.costs + a.costOff[p];
    if (wantW && nc == 1) {  // assignment.cpp:554-570, 840-856
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
    CDelta = CDelta * (double)nc;
    Node<R> nd;
#pragma unroll
    for (int s = 0; s < R; ++s) { nd.v[s] = 0.0; nd.u[s] = 0.0; nd.c4r[s] = -1; nd.r4c[s] = -1; }
    publish_cols<R>(sm, nd, lane);
    unsigned allRows = 0u;
#pragma unroll
    for (int s = 0; s < R; ++s) if (lane + 32 * s < n) allRows |= 1u << s;
    for (int c = 0; c < n; ++c) {
        if (augment_from<R>(c, nc, n, sm, nd, allRows, 0u, lane)) {
            if (lane == 0) { a.nFound[p] = 0; ctl->feasible = 0; ctl->done = 1; ctl->nFound = 0; ctl->nEmit = 0; }
            if (wantW && nc > 1) {
                double* out = a.probs + a.probOff[p];
                for (int i = lane; i < nc * (nL + 1); i += 32) out[i] = CUDART_NAN;
            }
            return;
        }
    }
    const double gain = path_gain(sm, n, nc);
    unsigned forb = 0u;
    {
        const int r0 = sm.r4c[0];
#pragma unroll
        for (int s = 0; s < R; ++s) if (lane + 32 * s == r0) forb |= 1u << s;
    }
    double gain0Out, cutoffGain = a.cutoff;
    bool cutMax = a.cutMaximize != 0;
    if (!maximize) {
        if (a.cutMode == PDA_CUT_RELATIVE) { cutoffGain = gain + a.cutoff; cutMax = false; }
        gain0Out = gain + CDelta;
    } else {
        if (a.cutMode == PDA_CUT_RELATIVE) { cutoffGain = gain - a.cutoff; cutMax = true; }
        gain0Out = -gain + CDelta;
    }
    node_store<R>(A.nodes, D, n, nd, forb, 0, lane);
    if (lane == 0) {
        HeapEntry e;
        e.gain = gain; e.node = 0; e.pad = 0;
        heap.put(0, e);
        A.orderNode[0] = 0;
        A.orderGain[0] = gain0Out;
        ctl->CDelta = CDelta; ctl->gain0Out = gain0Out; ctl->cutoffGain = cutoffGain;
        ctl->cutMax = cutMax ? 1 : 0; ctl->cutting = a.cutMode != PDA_CUT_NONE;
        ctl->freeRec = ~0ULL;
        ctl->heapLen = 1; ctl->sweep = 1; ctl->nNodes = 1; ctl->nTasks = 0; ctl->uncommitted = 0;
        ctl->feasible = 1;
        ctl->done = (a.k <= 1) ? 1 : 0;
        ctl->nFound = 1; ctl->nEmit = 1;
    }
}

// ---- warp 0, every round: commit in order, then choose the next splits -------------------------------
__device__ void serial_phase(const MurtyArgs& a, const CtaGeometry& cg, const CtaSmem& S, const Heap& heap,
                             const CtaArena& A, const int nc, const int lane) {
    CtaCtl* ctl = S.ctl;
    if (lane == 0) {
        int heapLen = ctl->heapLen, sweep = ctl->sweep, uncommitted = ctl->uncommitted;
        unsigned long long freeRec = ctl->freeRec;
        const bool maximize = a.maximize != 0;
        const double CDelta = ctl->CDelta, gain0Out = ctl->gain0Out;
        int done = 0;
        while (sweep < a.k) {
            const HeapEntry top = heap.get(0);
            const int rec = pad_record(top.pad);
            if (rec == 0) break;  // the top has not been split yet
            const int a0 = pad_active(top.pad), cnt = nc - a0;
            heap_pop(heap, heapLen);
            heapLen--;
            const int base = S.recBase[rec - 1];
            for (int j = 0; j < cnt; ++j) {  // children in column order (:493-522)
                const double g = S.recGain[(rec - 1) * PDA_CTA_MAX_COL + j];
                if (g == g) {
                    HeapEntry e;
                    e.gain = g; e.node = base + j; e.pad = a0 + j;
                    heap_sift_up(heap, heapLen, e);
                    heapLen++;
                }
            }
            freeRec |= 1ULL << (rec - 1);
            uncommitted -= cnt;
            if (heapLen == 0) { done = 1; ctl->nFound = sweep; ctl->nEmit = sweep; break; }
            const HeapEntry nt = heap.get(0);  // hypothesis number `sweep` (:703-719)
            double gainOut;
            bool stop = false;
            if (!maximize) {
                gainOut = nt.gain + CDelta;
                if (a.cutMode == PDA_CUT_RELATIVE && gainOut > gain0Out + a.cutoff) stop = true;
            } else {
                gainOut = -nt.gain + CDelta;
                if (a.cutMode == PDA_CUT_RELATIVE && gainOut < gain0Out - a.cutoff) stop = true;
            }
            A.orderNode[sweep] = nt.node;
            A.orderGain[sweep] = gainOut;
            if (stop) { done = 1; ctl->nFound = sweep; ctl->nEmit = sweep + 1; break; }
            sweep++;
        }
        if (!done && sweep >= a.k) { done = 1; ctl->nFound = sweep; ctl->nEmit = sweep; }
        ctl->heapLen = heapLen; ctl->sweep = sweep; ctl->uncommitted = uncommitted; ctl->freeRec = freeRec;
        ctl->done = done;
    }
    __syncwarp();
    if (ctl->done) return;

    // ---- choose what to split this round: the cheapest unsplit entries among the first 32 of the heap
    const int heapLen = ctl->heapLen;
    const int m = heapLen < 32 ? heapLen : 32;
    HeapEntry e;
    e.gain = CUDART_INF; e.node = 0; e.pad = 0;
    if (lane < m) e = heap.get(lane);
    double key = (lane < m && pad_record(e.pad) == 0) ? e.gain : CUDART_INF;
    unsigned long long freeRec = ctl->freeRec;
    int nNodes = ctl->nNodes, uncommitted = ctl->uncommitted, nTasks = 0;
    for (int it = 0; it < CTA_SPEC; ++it) {
        unsigned khi, klo;
        to_key(key, khi, klo);
        const unsigned mhi = __reduce_min_sync(FULL, khi);
        if (mhi >= KEY_INF_HI) break;
        const unsigned mlo = __reduce_min_sync(FULL, (khi == mhi) ? klo : 0xffffffffu);
        const bool win = (khi == mhi) && (klo == mlo);
        const int j = (int)__reduce_min_sync(FULL, win ? (unsigned)lane : 0xffffu);
        const int a0 = __shfl_sync(FULL, pad_active(e.pad), j);
        const int parent = __shfl_sync(FULL, e.node, j);
        const int cnt = nc - a0;
        if (it > 0) {  // speculative: needs task room, a spare record (one stays reserved for a top) and arena slack
            if (nTasks + cnt > CTA_MAXTASKS || __popcll(freeRec) < 2 || uncommitted + cnt > cg.specSlack) break;
        }
        const int rec = __ffsll((long long)freeRec) - 1;
        freeRec &= ~(1ULL << rec);
        if (lane == j) {
            e.pad |= (rec + 1) << 8;
            heap.put(j, e);
            key = CUDART_INF;
        }
        if (lane < cnt) {
            Task t;
            t.parent = parent; t.child = nNodes + lane; t.c = (short)(a0 + lane); t.rec = (short)rec;
            S.tasks[nTasks + lane] = t;
        }
        if (lane == 0) S.recBase[rec] = nNodes;
        nNodes += cnt;
        uncommitted += cnt;
        nTasks += cnt;
    }
    if (lane == 0) { ctl->freeRec = freeRec; ctl->nNodes = nNodes; ctl->uncommitted = uncommitted; ctl->nTasks = nTasks; }
}

// ---- any warp: one child of one split (shortestPathUpdateCPP, shortestPathCPP.cpp:240-365) -----------
template <int R>
__device__ void run_task(const MurtyArgs& a, const CtaSmem& S, const WarpSmem& sm, const CtaArena& A, const Task t,
                         const int n, const int nc, const int lane) {
    const int D = a.geo.nodeDim;
    Node<R> nd;
    unsigned parForb;
    int a0;
    node_load<R>(A.nodes + (size_t)t.parent * a.geo.nodeStride, D, n, nd, parForb, a0, lane);
    const int c = t.c;
    unsigned inPar = 0u;  // rows paired with columns >= c: columns a0..c-1 are fixed for this child (:506-508, 525-527)
#pragma unroll
    for (int s = 0; s < R; ++s) if (lane + 32 * s < n && nd.c4r[s] >= c) inPar |= 1u << s;
    publish_cols<R>(sm, nd, lane);
    const int r0 = sm.r4c[c];
    unsigned hideFirst = 0u;
#pragma unroll
    for (int s = 0; s < R; ++s) {
        if (lane + 32 * s == r0) { nd.c4r[s] = -1; hideFirst |= 1u << s; }
        if (lane + 32 * s == c) nd.r4c[s] = -1;
    }
    if (c == a0) hideFirst = parForb;
    if (lane == 0) sm.c4r[r0] = 0xffffu;
    __syncwarp();
    double result = CUDART_NAN;  // NaN = no child (infeasible or cut)
    const bool infeasible = augment_from<R>(c, nc, n, sm, nd, inPar, hideFirst, lane);
    if (!infeasible) {
        const double g = path_gain(sm, n, nc);
        const double cutoffGain = S.ctl->cutoffGain;
        const bool cut = S.ctl->cutting && (S.ctl->cutMax ? (g < cutoffGain) : (g > cutoffGain));
        if (!cut) {
            unsigned childForb = hideFirst;
            const int rNew = sm.r4c[c];
#pragma unroll
            for (int s = 0; s < R; ++s) if (lane + 32 * s == rNew) childForb |= 1u << s;
            node_store<R>(A.nodes + (size_t)t.child * a.geo.nodeStride, D, n, nd, childForb, c, lane);
            result = g;
        }
    }
    if (lane == 0) S.recGain[t.rec * PDA_CTA_MAX_COL + (c - a0)] = result;
    __syncwarp();
}

template <int R>
__device__ void solve_problem_cta(const MurtyArgs& a, const CtaGeometry& cg, const long long p, const CtaSmem& S,
                                  unsigned char* arenaBase) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n = a.numRow[p], nc = a.numCol[p];
    const bool wantW = a.weightMode != PDA_WEIGHTS_NONE;
    const int nL = wantW ? a.nL[p] : 0;
    if (nc < 1 || nc > n || n > 32 * R || nc > PDA_CTA_MAX_COL || (wantW && nL + nc != n)) {
        if (threadIdx.x == 0) a.nFound[p] = 0;
        return;
    }
    const WarpSmem sm = warp_view(S, cg, R, warp);
    const CtaArena A = carve_arena(arenaBase, cg, a.k);
    Heap heap;
    heap.top = S.heapTop;
    heap.deep = A.heapDeep;
    heap.topCap = cg.heapTopCap;
    CtaCtl* ctl = S.ctl;

    if (warp == 0) root_phase<R>(a, p, S, sm, heap, A, lane);
    __syncthreads();
    const bool feasible = ctl->feasible != 0, doneAtRoot = ctl->done != 0;
    __syncthreads();  // everyone has read the root's verdict before warp 0 may change it
    if (!feasible) return;

    while (!doneAtRoot) {
        if (warp == 0) serial_phase(a, cg, S, heap, A, nc, lane);
        __syncthreads();
        if (ctl->done) break;  // uniform: ctl is only written by warp 0 between the second and the first barrier
        const int nT = ctl->nTasks;
        for (int t = warp; t < nT; t += CTA_WARPS) run_task<R>(a, S, sm, A, S.tasks[t], n, nc, lane);
        __syncthreads();
    }

    // ---- lists (hpp:226-231) from the recorded pop order --------------------------------------------
    const int nFound = ctl->nFound, nEmit = ctl->nEmit;
    const int D = a.geo.nodeDim;
    for (int i = warp; i < nEmit; i += CTA_WARPS) {
        const unsigned char* nb = A.nodes + (size_t)A.orderNode[i] * a.geo.nodeStride + 16 * D;
        if (a.c4rBest) {
            int64_t* o = a.c4rBest + a.c4rOff[p] + (int64_t)i * n;
            for (int r = lane; r < n; r += 32) o[r] = (int64_t)(signed char)nb[r];
        }
        if (a.r4cBest) {
            int64_t* o = a.r4cBest + a.r4cOff[p] + (int64_t)i * nc;
            for (int c = lane; c < nc; c += 32) o[c] = (int64_t)(signed char)nb[D + c];
        }
        if (lane < nc) A.hypRows[(size_t)i * PDA_CTA_MAX_COL + lane] = nb[D + lane];
        if (a.gainBest && lane == 0) a.gainBest[p * (long long)a.k + i] = A.orderGain[i];
    }
    if (threadIdx.x == 0) a.nFound[p] = nFound;

    // ---- weights (assignment.cpp:616-648, 910-945): every table entry adds its terms in hypothesis order ---
    if (wantW && nc > 1) {
        const double best = ctl->gain0Out;
        for (int i = threadIdx.x; i < nFound; i += blockDim.x) {
            const double g = A.orderGain[i];
            const bool gatedOut = a.weightMode == PDA_WEIGHTS_GATED && !(best + a.weightGate > g);
            A.orderW[i] = gatedOut ? 0.0 : exp(best - g);
        }
        __syncthreads();
        const int cells = nc * (nL + 1);
        for (int t = threadIdx.x; t < cells; t += blockDim.x) {
            const int c = t / (nL + 1), to = t - c * (nL + 1);
            double acc = 0.0, total = 0.0;
            for (int i = 0; i < nFound; ++i) {
                const double w = A.orderW[i];
                int r = (int)(signed char)A.hypRows[(size_t)i * PDA_CTA_MAX_COL + c];
                r = r >= nL ? nL : r;
                total += w;
                if (r == to) acc += w;
            }
            a.probs[a.probOff[p] + t] = acc * (1.0 / total);
        }
    }
}

template <int R>
__global__ void __launch_bounds__(32 * CTA_WARPS, 1) murty_cta_kernel(const MurtyArgs a, const CtaGeometry cg) {
    extern __shared__ __align__(16) unsigned char smemRaw[];
    if ((int)blockIdx.x >= a.nWarps) return;  // nWarps = arenas available = CTAs allowed to run
    const CtaSmem S = carve_cta(smemRaw, a.geo, cg);
    unsigned char* arena = a.arena + (size_t)blockIdx.x * cg.arenaStride;
    for (;;) {
        __syncthreads();  // the previous problem's readers of ctl are done
        if (threadIdx.x == 0) S.ctl->problem = (long long)atomicAdd(a.cursor, 1ULL);
        __syncthreads();
        const long long p = S.ctl->problem;
        if (p >= a.nProblems) break;
        solve_problem_cta<R>(a, cg, p, S, arena);
    }
}

}  // namespace

// ---- host side ------------------------------------------------------------------------------------
static int round_up_i(int x, int m) { return (x + m - 1) / m * m; }

int murty_cta_geometry(int32_t k, int32_t maxNumRow, int32_t maxNumCol, bool weights, const DeviceInfo& dev,
                       MurtyGeometry* g, CtaGeometry* cg) {
    int rc = murty_geometry(k, maxNumRow, maxNumCol, weights, dev, g);
    if (rc) return rc;
    if (maxNumCol > PDA_CTA_MAX_COL)
        return fail(PDA_ERR_UNSUPPORTED, "murty (CTA path): numCol %d exceeds %d", maxNumCol, PDA_CTA_MAX_COL);
    const int D = 32 * g->R;
    cg->mirrorBytes = round_up_i(22 * D, 16);
    cg->specSlack = 16 * maxNumCol;
    const int64_t nodes = 1 + (int64_t)k * maxNumCol + cg->specSlack;
    if (nodes > (int64_t)1 << 30) return fail(PDA_ERR_UNSUPPORTED, "murty: k * numCol too large");
    cg->maxNodes = (int)nodes;
    cg->heapBytes = (int64_t)round_up_i((int)nodes, 8) * (int64_t)sizeof(HeapEntry);
    const int64_t orderBytes = (int64_t)k * 8 * 2 + (int64_t)((k + 3) & ~3) * 4 + (int64_t)k * PDA_CTA_MAX_COL;
    cg->nodesOff = (cg->heapBytes + orderBytes + 255) / 256 * 256;
    cg->arenaStride = (cg->nodesOff + nodes * g->nodeStride + 255) / 256 * 256;
    int off = 8 * (g->cCap + g->pCap) + CTA_WARPS * cg->mirrorBytes + 8 * CTA_RECORDS * PDA_CTA_MAX_COL +
              4 * CTA_RECORDS + (int)sizeof(Task) * CTA_MAXTASKS;
    off = round_up_i(off, 16);
    cg->ctlOff = off;
    off = round_up_i(off + (int)sizeof(CtaCtl), 16);
    cg->heapTopOff = off;
    int topCap = (dev.maxSmemOptin - off) / (int)sizeof(HeapEntry);
    if (topCap > cg->maxNodes) topCap = cg->maxNodes;
    if (topCap < 32) return fail(PDA_ERR_UNSUPPORTED, "murty (CTA path): problem too large for shared memory");
    cg->heapTopCap = topCap;
    cg->smemBytes = off + topCap * (int)sizeof(HeapEntry);
    return PDA_OK;
}

template <int R>
static int launch_murty_cta_r(const MurtyArgs& a, const CtaGeometry& cg, cudaStream_t stream) {
    PDA_CUDA_TRY(cudaFuncSetAttribute(murty_cta_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, cg.smemBytes));
    const int ctas = (int)(a.nProblems < a.nWarps ? a.nProblems : a.nWarps);
    murty_cta_kernel<R><<<ctas, 32 * CTA_WARPS, cg.smemBytes, stream>>>(a, cg);
    PDA_CUDA_TRY(cudaGetLastError());
    return PDA_OK;
}

int launch_murty_cta(const MurtyArgs& a, const CtaGeometry& cg, cudaStream_t stream) {
    PDA_CUDA_TRY(cudaMemsetAsync(a.cursor, 0, sizeof(unsigned long long), stream));
    switch (a.geo.R) {
        case 1: return launch_murty_cta_r<1>(a, cg, stream);
        case 2: return launch_murty_cta_r<2>(a, cg, stream);
        case 4: return launch_murty_cta_r<4>(a, cg, stream);
    }
    return fail(PDA_ERR_UNSUPPORTED, "murty: unsupported row-slot count %d", a.geo.R);
}

}  // namespace pda
