/* oracle_murty.c -- CPU parity oracle for the Murty k-best / shortest-augmenting-path
 * part of the hot path.  TEST INFRASTRUCTURE ONLY (see oracle_capi.h).
 *
 * A plain-C restatement of the algorithm in the reference's shortestPathCPP.cpp.
 * It is organised the way the CUDA kernel is (row sets as membership flags scanned
 * in ascending row order instead of shifted index lists; an explicit binary heap
 * instead of std::priority_queue; a node arena instead of new/delete) but every
 * floating-point expression keeps the reference's operand order, so results are
 * bit-identical to an IEEE-strict build of the reference:
 *
 *   safe matrix      shortestPathCPP.cpp:534-569, 582-585, 663-666
 *   root LAP         :119-238     (dijkstra_augment with no first-hop mask)
 *   dual update      :82-117
 *   gain             :59-80
 *   child re-solve   :240-365     (one Dijkstra from the un-assigned column)
 *   Murty partition  :455-532
 *   heap order       :30-42 + libstdc++ bits/stl_heap.h __push_heap/__adjust_heap
 *   drivers          :571-644 (kBest2D), :646-733 (kBest2DCutoff), :735-762 (assign2D)
 *
 * Why membership flags reproduce the reference's list order: Row2Scan is always an
 * ascending list (identity at :155-157; qsort at :486; order-preserving memmove at
 * :215, :345, :507, :526), so "first minimum in list order" == "lowest row index".
 */
#include "oracle_capi.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    int64_t* row4col;  /* [n] */
    int64_t* col4row;  /* [n] */
    double* u;         /* [n] column duals */
    double* v;         /* [n] row duals */
    uint8_t* forb;     /* [n] forbiddenActiveRows */
    double gain;
    int64_t activeCol;
    char* slab;
} Node;

typedef struct {
    double* spc;         /* shortestPathCost per row */
    int64_t* pred;       /* predecessor column per row */
    uint8_t* scanned;    /* ScannedRows */
    uint8_t* inScan;     /* membership form of Row2Scan */
    uint8_t* inScanPar;  /* membership form of Row2ScanParent */
    uint8_t* forb;       /* workMem.forbiddenActiveRows */
    int64_t* colOrder;   /* ScannedColIdx */
    double* C;           /* n x n safe matrix */
} Work;

typedef struct { double gain; Node* node; } HeapItem;
typedef struct { HeapItem* a; int64_t len, cap; } Heap;

/* sticky cut state of a ScratchSpace (shortestPathCPP.hpp:84-86, 130-132) */
typedef struct { int toCut; int maximize; double cutoffGain; } CutState;

static _Thread_local int64_t g_pops, g_children, g_iters, g_evals, g_maxHeap;
/* search-distance limit of the pruning model below (INFINITY = the plain reference behaviour) */
static _Thread_local double t_limit = INFINITY;
static _Thread_local int t_limit_hit = 0;

static Node* node_new(int64_t n) {
    Node* s = (Node*)malloc(sizeof(Node));
    size_t bytes = (size_t)n * (2 * sizeof(int64_t) + 2 * sizeof(double) + 1);
    s->slab = (char*)malloc(bytes ? bytes : 1);
    char* p = s->slab;
    s->row4col = (int64_t*)p; p += (size_t)n * sizeof(int64_t);
    s->col4row = (int64_t*)p; p += (size_t)n * sizeof(int64_t);
    s->u = (double*)p; p += (size_t)n * sizeof(double);
    s->v = (double*)p; p += (size_t)n * sizeof(double);
    s->forb = (uint8_t*)p;
    s->gain = 0.0;
    s->activeCol = 0;
    return s;
}
static void node_free(Node* s) { if (s) { free(s->slab); free(s); } }

static void work_init(Work* w, int64_t n) {
    size_t z = (size_t)(n > 0 ? n : 1);
    w->spc = (double*)malloc(z * sizeof(double));
    w->pred = (int64_t*)malloc(z * sizeof(int64_t));
    w->scanned = (uint8_t*)malloc(z);
    w->inScan = (uint8_t*)malloc(z);
    w->inScanPar = (uint8_t*)malloc(z);
    w->forb = (uint8_t*)malloc(z);
    w->colOrder = (int64_t*)malloc(z * sizeof(int64_t));
    w->C = (double*)malloc(z * z * sizeof(double));
}
static void work_free(Work* w) {
    free(w->spc); free(w->pred); free(w->scanned); free(w->inScan);
    free(w->inScanPar); free(w->forb); free(w->colOrder); free(w->C);
}

/* ---- heap: libstdc++ std::priority_queue<pMurtyHyp> mechanics -----------------
 * comp(a,b) == (a.gain > b.gain)  (shortestPathCPP.cpp:35-37), so the top is the
 * smallest gain.  push = push_back + __push_heap; pop = __pop_heap -> __adjust_heap. */
static void heap_sift_up(HeapItem* a, int64_t hole, int64_t top, HeapItem val) {
    int64_t parent = (hole - 1) / 2;
    while (hole > top && a[parent].gain > val.gain) {
        a[hole] = a[parent];
        hole = parent;
        parent = (hole - 1) / 2;
    }
    a[hole] = val;
}
static void heap_push(Heap* h, HeapItem it) {
    if (h->len == h->cap) {
        h->cap = h->cap ? 2 * h->cap : 64;
        h->a = (HeapItem*)realloc(h->a, (size_t)h->cap * sizeof(HeapItem));
    }
    h->len++;
    heap_sift_up(h->a, h->len - 1, 0, it);
    if (h->len > g_maxHeap) g_maxHeap = h->len;
}
static Node* heap_pop(Heap* h) {
    Node* top = h->a[0].node;
    if (h->len > 1) {
        int64_t len = h->len - 1;
        HeapItem val = h->a[len];
        HeapItem* a = h->a;
        int64_t hole = 0, child = 0;
        while (child < (len - 1) / 2) {
            child = 2 * (child + 1);
            if (a[child].gain > a[child - 1].gain) child--;  /* right wins an exact tie */
            a[hole] = a[child];
            hole = child;
        }
        if ((len & 1) == 0 && child == (len - 2) / 2) {
            child = 2 * (child + 1);
            a[hole] = a[child - 1];
            hole = child - 1;
        }
        heap_sift_up(a, hole, 0, val);
    }
    h->len--;
    return top;
}

/* One shortest augmenting path from `startCol`, then dual update and augmentation.
 * firstHopMask (may be NULL) hides rows only while the scan is at startCol (:310).
 * Consumes w->inScan.  Returns 1 if no finite path exists. */
static int dijkstra_augment(Node* s, const double* C, int64_t n, int64_t startCol,
                            const uint8_t* firstHopMask, Work* w) {
    const double INF = INFINITY;
    int64_t nScannedCols = 0, cur = startCol, sink = -1;
    double delta = 0;
    t_limit_hit = 0;
    for (int64_t r = 0; r < n; r++) { w->scanned[r] = 0; w->spc[r] = INF; }

    do {
        double minVal = INF;
        int64_t best = -1;
        w->colOrder[nScannedCols++] = cur;
        g_iters++;
        for (int64_t r = 0; r < n; r++) {
            if (!w->inScan[r]) continue;
            if (firstHopMask && cur == startCol && firstHopMask[r]) continue;
            double red = delta + C[r + cur * n] - s->u[cur] - s->v[r];
            g_evals++;
            if (red < w->spc[r]) { w->pred[r] = cur; w->spc[r] = red; }
            if (w->spc[r] < minVal) { minVal = w->spc[r]; best = r; }
        }
        if (minVal == INF) return 1;
        w->scanned[best] = 1;
        w->inScan[best] = 0;
        delta = w->spc[best];
        if (delta > t_limit) { t_limit_hit = 1; return 2; }  /* pruning model only: this child cannot be among the hypotheses still wanted */
        if (s->col4row[best] == -1) sink = best; else cur = s->col4row[best];
    } while (sink == -1);

    /* duals (:92-106), with row4col as it was before the flip */
    s->u[startCol] = s->u[startCol] + delta;
    for (int64_t i = 1; i < nScannedCols; i++) {
        int64_t c = w->colOrder[i];
        s->u[c] = s->u[c] + delta - w->spc[s->row4col[c]];
    }
    for (int64_t r = 0; r < n; r++)
        if (w->scanned[r]) s->v[r] = s->v[r] - delta + w->spc[r];

    /* flip along the predecessor chain (:108-116) */
    int64_t r = sink, c;
    do {
        c = w->pred[r];
        s->col4row[r] = c;
        int64_t h = s->row4col[c];
        s->row4col[c] = r;
        r = h;
    } while (c != startCol);
    return 0;
}

static double path_gain(const Node* s, const double* C, int64_t n, int64_t nColGain) {
    double g = 0;
    for (int64_t c = 0; c < nColGain; c++) g = g + C[c * n + s->row4col[c]];
    return g;
}

/* shortestPathCPP (:119-238): numColSolve columns of an n-row matrix with leading dim n. */
static int root_solve(Node* s, Work* w, const double* C, int64_t n, int64_t numColSolve, int64_t nColGain) {
    for (int64_t r = 0; r < n; r++) { s->col4row[r] = -1; s->v[r] = 0; s->forb[r] = 0; }
    for (int64_t c = 0; c < numColSolve; c++) { s->u[c] = 0; s->row4col[c] = -1; }
    s->activeCol = 0;
    for (int64_t c = 0; c < numColSolve; c++) {
        for (int64_t r = 0; r < n; r++) w->inScan[r] = 1;
        if (dijkstra_augment(s, C, n, c, NULL, w)) { s->gain = -1; return 1; }
    }
    s->gain = path_gain(s, C, n, nColGain);
    if (numColSolve > 0) s->forb[s->row4col[0]] = 1;
    return 0;
}

static int cut_hyp(const CutState* cs, double gain) {
    if (!cs->toCut) return 0;
    return cs->maximize ? (gain < cs->cutoffGain) : (gain > cs->cutoffGain);
}

/* shortestPathUpdateCPP (:240-365): clone the parent, free (row4col[c], c), re-solve. */
static Node* child_solve(const Node* par, Work* w, int64_t startCol, int64_t numVarCol, int64_t n) {
    Node* s = node_new(n);
    g_children++;
    s->activeCol = startCol;
    memcpy(s->row4col, par->row4col, (size_t)n * sizeof(int64_t));
    memcpy(s->col4row, par->col4row, (size_t)n * sizeof(int64_t));
    memcpy(s->u, par->u, (size_t)n * sizeof(double));
    memcpy(s->v, par->v, (size_t)n * sizeof(double));
    memcpy(s->forb, w->forb, (size_t)n);
    s->col4row[s->row4col[startCol]] = -1;
    s->row4col[startCol] = -1;
    if (dijkstra_augment(s, w->C, n, startCol, w->forb, w)) { s->gain = -1; return s; }
    s->gain = path_gain(s, w->C, n, numVarCol);
    s->forb[s->row4col[startCol]] = 1;
    return s;
}

/* split (:455-532) */
static void split(const Node* par, Heap* heap, Work* w, const CutState* cs, int64_t numVarCol, int64_t n) {
    const int64_t a = par->activeCol;
    memset(w->inScanPar, 0, (size_t)n);
    for (int64_t c = a; c < n; c++) w->inScanPar[par->row4col[c]] = 1;

    /* first child keeps the parent's constraints on the active column (:488-501) */
    memcpy(w->inScan, w->inScanPar, (size_t)n);
    memcpy(w->forb, par->forb, (size_t)n);
    Node* ch = child_solve(par, w, a, numVarCol, n);
    if (ch->gain == -1 || cut_hyp(cs, ch->gain)) node_free(ch);
    else { HeapItem it = {ch->gain, ch}; heap_push(heap, it); }
    w->inScanPar[par->row4col[a]] = 0;  /* column a is now fixed (:506-508) */

    memset(w->forb, 0, (size_t)n);
    for (int64_t c = a + 1; c < numVarCol; c++) {
        memcpy(w->inScan, w->inScanPar, (size_t)n);
        w->forb[par->row4col[c]] = 1;  /* only the parent's own pairing is excluded (:516) */
        ch = child_solve(par, w, c, numVarCol, n);
        if (ch->gain == -1 || cut_hyp(cs, ch->gain)) node_free(ch);
        else { HeapItem it = {ch->gain, ch}; heap_push(heap, it); }
        w->inScanPar[par->row4col[c]] = 0;
        w->forb[par->row4col[c]] = 0;
    }
}

/* makeCostMatrixSafe (:534-569) + zero padding to n x n (:585, :666); returns CDelta*numCol. */
static double make_safe_padded(Work* w, const double* C, int64_t n, int64_t numCol, int maximize) {
    const int64_t numEl = n * numCol;
    double d = C[0];
    if (!maximize) {
        for (int64_t i = 1; i < numEl; i++) if (C[i] < d) d = C[i];
        for (int64_t i = 0; i < numEl; i++) w->C[i] = C[i] - d;
    } else {
        for (int64_t i = 1; i < numEl; i++) if (d < C[i]) d = C[i];
        for (int64_t i = 0; i < numEl; i++) w->C[i] = -C[i] + d;
    }
    for (int64_t i = numEl; i < n * n; i++) w->C[i] = 0;
    return d * (double)numCol;
}

static void emit(const Node* s, int64_t slot, int64_t n, int64_t numCol, int64_t* c4r, int64_t* r4c) {
    memcpy(c4r + slot * n, s->col4row, (size_t)n * sizeof(int64_t));
    memcpy(r4c + slot * numCol, s->row4col, (size_t)numCol * sizeof(int64_t));
}

/* kBest2D / kBest2DCutoff.  useCutoff selects the second driver; cs carries the
 * ScratchSpace's sticky cut state in and out. */
static int64_t kbest_core(int64_t k, int64_t n, int64_t numCol, int maximize, const double* C,
                          int64_t* c4rBest, int64_t* r4cBest, double* gainBest,
                          int useCutoff, double cutoff, CutState* cs) {
    g_pops = g_children = g_iters = g_evals = g_maxHeap = 0;
    if (useCutoff) { cs->toCut = 1; cs->maximize = maximize; }
    Work w;
    work_init(&w, n);
    Heap heap = {NULL, 0, 0};
    Node* cur = node_new(n);
    double CDelta = make_safe_padded(&w, C, n, numCol, maximize);

    if (root_solve(cur, &w, w.C, n, n, numCol)) {
        node_free(cur);
        work_free(&w);
        return 0;
    }
    emit(cur, 0, n, numCol, c4rBest, r4cBest);
    gainBest[0] = cur->gain;
    if (!maximize) {
        if (useCutoff) cs->cutoffGain = gainBest[0] + cutoff;
        gainBest[0] = gainBest[0] + CDelta;
    } else {
        if (useCutoff) cs->cutoffGain = gainBest[0] - cutoff;
        gainBest[0] = -gainBest[0] + CDelta;
    }
    { HeapItem it = {cur->gain, cur}; heap_push(&heap, it); }

    int64_t sweep;
    for (sweep = 1; sweep < k; sweep++) {
        cur = heap_pop(&heap);
        g_pops++;
        split(cur, &heap, &w, cs, numCol, n);
        node_free(cur);
        if (heap.len == 0) break;
        cur = heap.a[0].node;
        emit(cur, sweep, n, numCol, c4rBest, r4cBest);
        gainBest[sweep] = cur->gain;
        if (!maximize) {
            gainBest[sweep] = gainBest[sweep] + CDelta;
            if (useCutoff && gainBest[sweep] > gainBest[0] + cutoff) break;
        } else {
            gainBest[sweep] = -gainBest[sweep] + CDelta;
            if (useCutoff && gainBest[sweep] < gainBest[0] - cutoff) break;
        }
    }
    for (int64_t i = 0; i < heap.len; i++) node_free(heap.a[i].node);
    free(heap.a);
    work_free(&w);
    return sweep;
}

int64_t orc_kbest2d(int64_t k, int64_t numRow, int64_t numCol, int maximize, const double* C,
                    int64_t* col4row, int64_t* row4col, double* gain) {
    CutState cs = {0, 0, 0.0};
    return kbest_core(k, numRow, numCol, maximize, C, col4row, row4col, gain, 0, 0.0, &cs);
}

int64_t orc_kbest2d_cutoff(int64_t k, int64_t numRow, int64_t numCol, int maximize, const double* C,
                           int64_t* col4row, int64_t* row4col, double* gain, double cutoff) {
    CutState cs = {0, 0, 0.0};
    return kbest_core(k, numRow, numCol, maximize, C, col4row, row4col, gain, 1, cutoff, &cs);
}

int64_t orc_kbest2d_after_cutoff(int64_t k, int64_t numRow, int64_t numCol, int maximize, const double* C,
                                 int64_t* col4row, int64_t* row4col, double* gain,
                                 int firstMaximize, const double* firstC, double firstCutoff) {
    CutState cs = {0, 0, 0.0};
    int64_t* c4r = (int64_t*)malloc((size_t)(numRow > 0 ? numRow : 1) * sizeof(int64_t));
    int64_t* r4c = (int64_t*)malloc((size_t)(numCol > 0 ? numCol : 1) * sizeof(int64_t));
    double g;
    kbest_core(1, numRow, numCol, firstMaximize, firstC, c4r, r4c, &g, 1, firstCutoff, &cs);
    free(c4r);
    free(r4c);
    return kbest_core(k, numRow, numCol, maximize, C, col4row, row4col, gain, 0, 0.0, &cs);
}

static void export_node(const Node* s, int64_t numRow, int64_t numCol, int64_t* col4row, int64_t* row4col,
                        double* u, double* v, double* gain) {
    for (int64_t r = 0; r < numRow; r++) { col4row[r] = s->col4row[r]; v[r] = s->v[r]; }
    for (int64_t c = 0; c < numCol; c++) { row4col[c] = s->row4col[c]; u[c] = s->u[c]; }
    *gain = s->gain;
}

int orc_assign2d(int64_t numRow, int64_t numCol, int maximize, const double* C,
                 int64_t* col4row, int64_t* row4col, double* u, double* v, double* gain) {
    Work w;
    work_init(&w, numRow);
    Node* s = node_new(numRow);
    const int64_t numEl = numRow * numCol;
    double d = C[0];
    if (!maximize) {
        for (int64_t i = 1; i < numEl; i++) if (C[i] < d) d = C[i];
        for (int64_t i = 0; i < numEl; i++) w.C[i] = C[i] - d;
    } else {
        for (int64_t i = 1; i < numEl; i++) if (d < C[i]) d = C[i];
        for (int64_t i = 0; i < numEl; i++) w.C[i] = -C[i] + d;
    }
    double CDelta = d * (double)numCol;
    int infeasible = root_solve(s, &w, w.C, numRow, numCol, numCol);
    if (!infeasible) s->gain = maximize ? (-s->gain + CDelta) : (s->gain + CDelta);
    export_node(s, numRow, numCol, col4row, row4col, u, v, gain);
    node_free(s);
    work_free(&w);
    return infeasible ? 0 : 1;
}

int orc_shortest_path(int64_t numRow, int64_t numCol, int64_t numCol4Gain, const double* Cprepared,
                      int64_t* col4row, int64_t* row4col, double* u, double* v, double* gain,
                      uint8_t* forbidden) {
    Work w;
    work_init(&w, numRow);
    Node* s = node_new(numRow);
    int infeasible = root_solve(s, &w, Cprepared, numRow, numCol, numCol4Gain);
    export_node(s, numRow, numCol, col4row, row4col, u, v, gain);
    for (int64_t r = 0; r < numRow; r++) forbidden[r] = s->forb[r];
    node_free(s);
    work_free(&w);
    return infeasible;
}

/* ---- CPU model of the pruning kernel (murty_kernel<R, true>, probabilisticsemslam_b200/csrc/murty_kernel.cu) -------
 * TEST INFRASTRUCTURE: restates the DECISIONS of the CUDA fast path -- the bound T, when and how it is tightened, which
 * children are abandoned or dropped, when a selection counts as tied -- on top of this file's reference arithmetic, so the
 * soundness of those rules can be checked on the CPU against the plain enumeration above (tests/test_pruning_model.py) and
 * its counters against the kernel's (-DPDA_FAST_STATS).  Same constants, same operand order as the kernel:
 *   cap    = roundup32(k + 2*maxCol + 8 + 32) slots; a split that might overflow them bails
 *   m      = k - sweep hypotheses still wanted; tighten when (T == inf ? live >= m : live - m >= trig); trig = live - m + 8
 *   limit  = (T - gain(parent)) + 1e-7 * (T + 1)     a search is abandoned once its distance exceeds it
 *   tie    = a second live entry with the winner's gain bits  =>  bail (the exact kernel redoes the problem)
 * Returns the number of hypotheses, or -2 if the model bails.  stats[6]: children, abandoned, dropped when finished,
 * kept, tightenings, slots at tightening. */
typedef struct { double gain; Node* node; int live; } Slot;

static void model_tighten(Slot* q, int64_t* len, int64_t* live, double* T, int64_t m) {
    double lo = INFINITY, hi = -INFINITY;
    for (int64_t i = 0; i < *len; i++) if (q[i].live) { if (q[i].gain < lo) lo = q[i].gain; if (q[i].gain > hi) hi = q[i].gain; }
    double Tn = hi;
    if (hi > lo) {
        const double width = (hi - lo) * 0.03125;
        const double inv = 1.0 / width;
        int64_t hist[32] = {0};
        for (int64_t i = 0; i < *len; i++) if (q[i].live) {
            int b = (int)((q[i].gain - lo) * inv);
            b = b > 31 ? 31 : (b < 0 ? 0 : b);
            hist[b]++;
        }
        int64_t cum = 0; int bstar = -1;
        for (int b = 0; b < 32; b++) { cum += hist[b]; if (cum >= m) { bstar = b; break; } }
        if (bstar >= 0 && bstar < 31) { Tn = lo + (double)(bstar + 1) * width; if (Tn > hi) Tn = hi; }
    }
    int64_t cnt = 0;
    for (int64_t i = 0; i < *len; i++) if (q[i].live && q[i].gain <= Tn) cnt++;
    if (cnt < m) Tn = hi;
    int64_t out = 0;
    for (int64_t i = 0; i < *len; i++) {
        if (q[i].live && q[i].gain <= Tn) q[out++] = q[i];
        else if (q[i].live) node_free(q[i].node);
    }
    *len = out; *live = out; *T = Tn;
}

int64_t orc_kbest2d_cutoff_pruned(int64_t k, int64_t n, int64_t numCol, int maximize, const double* C,
                                  int64_t* c4rBest, int64_t* r4cBest, double* gainBest, double cutoff,
                                  int64_t maxCol, int64_t* stats) {
    for (int i = 0; i < 6; i++) stats[i] = 0;
    CutState cs = {1, maximize, 0.0};
    Work w;
    work_init(&w, n);
    Node* cur = node_new(n);
    const double CDelta = make_safe_padded(&w, C, n, numCol, maximize);
    if (root_solve(cur, &w, w.C, n, n, numCol)) { node_free(cur); work_free(&w); return 0; }
    emit(cur, 0, n, numCol, c4rBest, r4cBest);
    gainBest[0] = cur->gain;
    if (!maximize) { cs.cutoffGain = gainBest[0] + cutoff; gainBest[0] = gainBest[0] + CDelta; }
    else { cs.cutoffGain = gainBest[0] - cutoff; gainBest[0] = -gainBest[0] + CDelta; }
    const int64_t cap = (k + 2 * maxCol + 8 + 32 + 31) / 32 * 32;
    Slot* q = (Slot*)malloc(sizeof(Slot) * (size_t)(cap + numCol + 1));
    int64_t len = 0, live = 0, trig = 0, sweep, result = -1;
    double T = INFINITY;
    for (sweep = 1; sweep < k; sweep++) {
        const int64_t m = k - sweep;
        if ((T == INFINITY) ? (live >= m) : (live - m >= trig)) {
            stats[4]++; stats[5] += len;
            model_tighten(q, &len, &live, &T, m);
            trig = live - m + 8;
        }
        if (len + numCol > cap) { result = -2; break; }
        const double limit = (T - cur->gain) + 1e-7 * (T + 1.0);
        /* split (:455-532), every child searched under the limit */
        const int64_t a = cur->activeCol;
        memset(w.inScanPar, 0, (size_t)n);
        for (int64_t c = a; c < n; c++) w.inScanPar[cur->row4col[c]] = 1;
        for (int64_t c = a; c < numCol; c++) {
            memcpy(w.inScan, w.inScanPar, (size_t)n);
            if (c == a) memcpy(w.forb, cur->forb, (size_t)n);
            else { memset(w.forb, 0, (size_t)n); w.forb[cur->row4col[c]] = 1; }
            stats[0]++;
            t_limit = limit;
            Node* ch = child_solve(cur, &w, c, numCol, n);
            const int abandoned = (ch->gain == -1 && t_limit_hit);
            t_limit = INFINITY;
            if (abandoned) stats[1]++;
            if (ch->gain == -1 || cut_hyp(&cs, ch->gain)) node_free(ch);
            else if (ch->gain > T) { stats[2]++; node_free(ch); }
            else { stats[3]++; q[len].gain = ch->gain; q[len].node = ch; q[len].live = 1; len++; live++; }
            w.inScanPar[cur->row4col[c]] = 0;
        }
        node_free(cur);
        cur = NULL;
        if (live == 0) break;
        /* take the smallest; a second live entry with the same bits means only the reference's heap can order them */
        int64_t best = -1;
        for (int64_t i = 0; i < len; i++) if (q[i].live && (best < 0 || q[i].gain < q[best].gain)) best = i;
        int tie = 0;
        for (int64_t i = 0; i < len; i++) if (q[i].live && i != best && q[i].gain == q[best].gain) tie = 1;
        if (tie) { result = -2; break; }
        cur = q[best].node;
        q[best].live = 0;
        live--;
        emit(cur, sweep, n, numCol, c4rBest, r4cBest);
        gainBest[sweep] = cur->gain;
        if (!maximize) {
            gainBest[sweep] = gainBest[sweep] + CDelta;
            if (gainBest[sweep] > gainBest[0] + cutoff) break;
        } else {
            gainBest[sweep] = -gainBest[sweep] + CDelta;
            if (gainBest[sweep] < gainBest[0] - cutoff) break;
        }
    }
    if (result != -2) result = sweep;
    for (int64_t i = 0; i < len; i++) if (q[i].live) node_free(q[i].node);
    if (cur) node_free(cur);
    free(q);
    work_free(&w);
    return result;
}

void orc_last_counters(int64_t* pops, int64_t* childSolves, int64_t* dijkstraIters, int64_t* evaluations, int64_t* maxHeap) {
    *pops = g_pops; *childSolves = g_children; *dijkstraIters = g_iters; *evaluations = g_evals; *maxHeap = g_maxHeap;
}
