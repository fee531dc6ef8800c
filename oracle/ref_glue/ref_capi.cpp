// extern "C" door onto the UNMODIFIED reference hot path, built by
// oracle/Makefile into oracle/_ref/libpda_ref_*.so from the sources where they
// lie under /root/reference (shortestPathCPP.cpp whole; assignment.cpp and
// nwPerm.cpp by line range against ref_prelude_*.h).  Nothing of the reference
// is copied into the repository.
//
// TEST INFRASTRUCTURE ONLY: used by tests/ (to pin the oracle restatement and to
// produce tests/golden/*) and by bench.py's cpu_baseline / --impl reference legs.
// The product (libpda_b200.so) never links or loads this.
//
// Every entry point has an `orc_` twin with the same signature in
// oracle/oracle_capi.h (the CPU restatement), so the Python harness can drive
// either through one wrapper.
#include "ref_prelude_assignment.h"

#include <atomic>
#include <cstdint>
#include <cstring>
#include <pthread.h>

namespace {

// The reference keeps k-best lists in stack VLAs (assignment.cpp:586-588, 873-875:
// up to 20000*nRows*8 bytes), so every call runs on a thread with a roomy stack.
const size_t kStackBytes = size_t(512) << 20;

struct Thunk { void (*fn)(void*); void* arg; };
void* thunkMain(void* p) { Thunk* t = static_cast<Thunk*>(p); t->fn(t->arg); return nullptr; }

void runOnBigStack(void (*fn)(void*), void* arg) {
    pthread_attr_t attr;
    pthread_attr_init(&attr);
    pthread_attr_setstacksize(&attr, kStackBytes);
    Thunk t = {fn, arg};
    pthread_t th;
    if (pthread_create(&th, &attr, thunkMain, &t) != 0) { fn(arg); }
    else pthread_join(th, nullptr);
    pthread_attr_destroy(&attr);
}

void flattenProbs(const std::vector<std::vector<double> >& p, double* out) {
    size_t o = 0;
    for (size_t m = 0; m < p.size(); m++)
        for (size_t l = 0; l < p[m].size(); l++) out[o++] = p[m][l];
}

Eigen::MatrixXd wrap(const double* A, int64_t rows, int64_t cols) {
    Eigen::MatrixXd M(rows, cols);
    if (rows * cols > 0) std::memcpy(M.data(), A, sizeof(double) * size_t(rows * cols));
    return M;
}

}  // namespace

extern "C" {

int64_t ref_kbest2d(int64_t k, int64_t numRow, int64_t numCol, int maximize, const double* C,
                    int64_t* col4row, int64_t* row4col, double* gain) {
    ScratchSpace w;
    w.init(size_t(numRow), size_t(numRow));
    return int64_t(kBest2D(size_t(k), size_t(numRow), size_t(numCol), maximize != 0, C, w,
                           reinterpret_cast<ptrdiff_t*>(col4row), reinterpret_cast<ptrdiff_t*>(row4col), gain));
}

int64_t ref_kbest2d_cutoff(int64_t k, int64_t numRow, int64_t numCol, int maximize, const double* C,
                           int64_t* col4row, int64_t* row4col, double* gain, double cutoff) {
    ScratchSpace w;
    w.init(size_t(numRow), size_t(numRow));
    return int64_t(kBest2DCutoff(size_t(k), size_t(numRow), size_t(numCol), maximize != 0, C, w,
                                 reinterpret_cast<ptrdiff_t*>(col4row), reinterpret_cast<ptrdiff_t*>(row4col), gain, cutoff));
}

// kBest2D on a ScratchSpace that an earlier kBest2DCutoff left behind: the
// reference never clears toCut/cutoffGain/maximize (shortestPathCPP.cpp:650-651,
// hpp:84-86), so the second call still prunes.  first* outputs are scratch.
int64_t ref_kbest2d_after_cutoff(int64_t k, int64_t numRow, int64_t numCol, int maximize, const double* C,
                                 int64_t* col4row, int64_t* row4col, double* gain,
                                 int firstMaximize, const double* firstC, double firstCutoff) {
    ScratchSpace w;
    w.init(size_t(numRow), size_t(numRow));
    std::vector<ptrdiff_t> c4r(static_cast<size_t>(numRow), 0), r4c(static_cast<size_t>(numCol), 0);
    double g;
    kBest2DCutoff(1, size_t(numRow), size_t(numCol), firstMaximize != 0, firstC, w, c4r.data(), r4c.data(), &g, firstCutoff);
    return int64_t(kBest2D(size_t(k), size_t(numRow), size_t(numCol), maximize != 0, C, w,
                           reinterpret_cast<ptrdiff_t*>(col4row), reinterpret_cast<ptrdiff_t*>(row4col), gain));
}

int ref_assign2d(int64_t numRow, int64_t numCol, int maximize, const double* C,
                 int64_t* col4row, int64_t* row4col, double* u, double* v, double* gain) {
    ScratchSpace w;
    w.init(size_t(numRow), size_t(numCol));
    const size_t nRz = static_cast<size_t>(numRow), nCz = static_cast<size_t>(numCol);
    MurtyHyp h(nRz, nCz);
    for (int64_t c = 0; c < numCol; c++) h.row4col[c] = -1;
    int ret = assign2D(size_t(numRow), size_t(numCol), maximize != 0, C, w, &h);
    for (int64_t r = 0; r < numRow; r++) { col4row[r] = h.col4row[r]; v[r] = h.v[r]; }
    for (int64_t c = 0; c < numCol; c++) { row4col[c] = h.row4col[c]; u[c] = h.u[c]; }
    *gain = h.gain;
    return ret;
}

int ref_shortest_path(int64_t numRow, int64_t numCol, int64_t numCol4Gain, const double* Cprepared,
                      int64_t* col4row, int64_t* row4col, double* u, double* v, double* gain,
                      uint8_t* forbidden) {
    ScratchSpace w;
    w.init(size_t(numRow), size_t(numCol));
    std::memcpy(w.C, Cprepared, sizeof(double) * size_t(numRow * numCol));
    const size_t nRz = static_cast<size_t>(numRow), nCz = static_cast<size_t>(numCol);
    MurtyHyp h(nRz, nCz);
    for (int64_t c = 0; c < numCol; c++) h.row4col[c] = -1;
    int ret = shortestPathCPP(&h, w, size_t(numRow), size_t(numCol), size_t(numCol4Gain));
    for (int64_t r = 0; r < numRow; r++) { col4row[r] = h.col4row[r]; v[r] = h.v[r]; forbidden[r] = h.forbiddenActiveRows[r] ? 1 : 0; }
    for (int64_t c = 0; c < numCol; c++) { row4col[c] = h.row4col[c]; u[c] = h.u[c]; }
    *gain = h.gain;
    return ret;
}

int64_t ref_condition_costs(const double* costs, int64_t nRows, int64_t nCols, double* outCosts, int64_t* rowIdx) {
    std::vector<double> in(costs, costs + nRows * nCols);
    std::vector<ptrdiff_t> idx;
    std::vector<double> out = conditionCosts(in, size_t(nRows), size_t(nCols), idx);
    std::copy(out.begin(), out.end(), outCosts);
    for (size_t i = 0; i < idx.size(); i++) rowIdx[i] = idx[i];
    return int64_t(idx.size());
}

void ref_to_probs(double* v, int64_t n) {
    std::vector<double> t(v, v + n);
    toProbs(t);
    std::copy(t.begin(), t.end(), v);
}

struct ProbArgs { const double* costs; int64_t nL, nM, k; double* probs; int which; int permOpt; int status; };
static void probBody(void* p) {
    ProbArgs* a = static_cast<ProbArgs*>(p);
    std::vector<double> c(a->costs, a->costs + (a->nL + a->nM) * a->nM);
    try {
        std::vector<std::vector<double> > r;
        if (a->which == 0) r = assignmentProb(c, size_t(a->nL), size_t(a->nM), size_t(a->k));
        else if (a->which == 1) r = bruteForceProb(c, size_t(a->nL), size_t(a->nM));
        else r = permanentProb(c, size_t(a->nL), size_t(a->nM), a->permOpt);
        flattenProbs(r, a->probs);
        a->status = 0;
    } catch (const std::exception&) { a->status = 1; }
}

int ref_assignment_prob(const double* costs, int64_t nL, int64_t nM, int64_t k, double* probs) {
    ProbArgs a = {costs, nL, nM, k, probs, 0, 0, 0};
    runOnBigStack(probBody, &a);
    return a.status;
}
int ref_brute_force_prob(const double* costs, int64_t nL, int64_t nM, double* probs) {
    ProbArgs a = {costs, nL, nM, 0, probs, 1, 0, 0};
    runOnBigStack(probBody, &a);
    return a.status;
}
int ref_permanent_prob(const double* costs, int64_t nL, int64_t nM, int permOpt, double* probs) {
    ProbArgs a = {costs, nL, nM, 0, probs, 2, permOpt, 0};
    runOnBigStack(probBody, &a);
    return a.status;
}

// getAssignmentProbs (assignment.cpp:38-139) takes GTSAM quadrics and cannot be compiled here; this is its
// cost-matrix-level body, lines :57-74, written against the reference's own conditionCosts / assignmentProb /
// permanentProb (the only restated part is the scatter loop :68-74).
struct AssocArgs { const double* costs; int64_t nL, nM, k; int usePerm; double* probs; int status; };
static void assocBody(void* p) {
    AssocArgs* a = static_cast<AssocArgs*>(p);
    const size_t nL = size_t(a->nL), nM = size_t(a->nM);
    a->status = 0;
    if (nM == 0) return;
    if (nL == 0) { for (size_t m = 0; m < nM; m++) a->probs[m] = 1.0; return; }
    try {
        std::vector<double> costMatrix(a->costs, a->costs + (nL + nM) * nM);
        std::vector<ptrdiff_t> rowIdx;
        std::vector<double> conditionedCosts = conditionCosts(costMatrix, nL + nM, nM, rowIdx);
        size_t condL = (conditionedCosts.size() / nM) - nM;
        std::vector<std::vector<double> > conditionedProbs;
        if (a->usePerm) conditionedProbs = permanentProb(conditionedCosts, condL, nM, 1);
        else conditionedProbs = assignmentProb(conditionedCosts, condL, nM, size_t(a->k));
        std::vector<std::vector<double> > probs(nM, std::vector<double>(nL + 1, 0));
        for (size_t m = 0; m < nM; m++) {
            for (size_t l = 0; l < condL; l++) probs[m][rowIdx[l]] = conditionedProbs[m][l];
            probs[m][nL] = conditionedProbs[m][condL];
        }
        flattenProbs(probs, a->probs);
    } catch (const std::exception&) { a->status = 1; }
}
int ref_association_probs(const double* costs, int64_t nL, int64_t nM, int64_t k, int usePerm, double* probs) {
    AssocArgs a = {costs, nL, nM, k, usePerm, probs, 0};
    runOnBigStack(assocBody, &a);
    return a.status;
}

double ref_permanent_exact(const double* A, int64_t rows, int64_t cols, int* status) {
    try { *status = 0; return permanentExact(wrap(A, rows, cols)); }
    catch (const std::exception&) { *status = 1; return 0.0; }
}
double ref_permanent_exact_square(const double* A, int64_t n, int* status) {
    try { *status = 0; return permanentExactSquare(wrap(A, n, n)); }
    catch (const std::exception&) { *status = 1; return 0.0; }
}
double ref_permanent_exact_long(const double* A, int64_t rows, int64_t cols, int* status) {
    try { *status = 0; return double(permanentExactLong(wrap(A, rows, cols))); }
    catch (const std::exception&) { *status = 1; return 0.0; }
}
double ref_conditioned_permanent(const double* A, int64_t rows, int64_t cols, int permOpt, int* status) {
    try { *status = 0; return conditionedPermanent(wrap(A, rows, cols), permOpt); }
    catch (const std::exception&) { *status = 1; return 0.0; }
}

// ---- batch drivers (CPU baseline timing; one ScratchSpace per call, one
// problem per thread at a time -- the reference functions are re-entrant) -----
struct BatchArgs {
    const double* costs; const int64_t* costOff; const int32_t* nL; const int32_t* nM;
    int64_t n; int64_t k; double cutoff;
    double* probs; const int64_t* probOff;
    int64_t* col4row; const int64_t* c4rOff; int64_t* row4col; const int64_t* r4cOff;
    double* gain; int32_t* nFound;
    std::atomic<int64_t>* next;
};
static void batchBody(void* p) {
    BatchArgs* a = static_cast<BatchArgs*>(p);
    for (;;) {
        int64_t i = a->next->fetch_add(1);
        if (i >= a->n) break;
        const int64_t nL = a->nL[i], nM = a->nM[i], nR = nL + nM;
        const double* c = a->costs + a->costOff[i];
        if (a->probs) {
            std::vector<double> cm(c, c + nR * nM);
            flattenProbs(assignmentProb(cm, size_t(nL), size_t(nM), size_t(a->k)), a->probs + a->probOff[i]);
        }
        if (a->gain) {
            ScratchSpace w;
            w.init(size_t(nR), size_t(nR));
            size_t f = kBest2DCutoff(size_t(a->k), size_t(nR), size_t(nM), false, c, w,
                                     reinterpret_cast<ptrdiff_t*>(a->col4row + a->c4rOff[i]),
                                     reinterpret_cast<ptrdiff_t*>(a->row4col + a->r4cOff[i]),
                                     a->gain + i * a->k, a->cutoff);
            a->nFound[i] = int32_t(f);
        }
    }
}
static void* batchThread(void* p) { batchBody(p); return nullptr; }

// Runs problems [0,n) over nThreads host threads; returns wall seconds.
// probs != NULL -> assignmentProb per problem; gain != NULL -> kBest2DCutoff lists.
double ref_batch(const double* costs, const int64_t* costOff, const int32_t* nL, const int32_t* nM,
                 int64_t n, int64_t k, double cutoff, int nThreads,
                 double* probs, const int64_t* probOff,
                 int64_t* col4row, const int64_t* c4rOff, int64_t* row4col, const int64_t* r4cOff,
                 double* gain, int32_t* nFound) {
    std::atomic<int64_t> next(0);
    BatchArgs a = {costs, costOff, nL, nM, n, k, cutoff, probs, probOff, col4row, c4rOff, row4col, r4cOff, gain, nFound, &next};
    if (nThreads < 1) nThreads = 1;
    std::vector<pthread_t> th(static_cast<size_t>(nThreads));
    pthread_attr_t attr;
    pthread_attr_init(&attr);
    pthread_attr_setstacksize(&attr, size_t(64) << 20);
    auto t0 = tic();
    for (int t = 0; t < nThreads; t++) pthread_create(&th[size_t(t)], &attr, batchThread, &a);
    for (int t = 0; t < nThreads; t++) pthread_join(th[size_t(t)], nullptr);
    double dt = toc(t0);
    pthread_attr_destroy(&attr);
    return dt;
}

// n permanents of size dim x dim (column-major, contiguous); returns wall seconds.
struct PermBatch { const double* mats; int64_t dim; int64_t n; double* out; std::atomic<int64_t>* next; };
static void* permThread(void* p) {
    PermBatch* a = static_cast<PermBatch*>(p);
    for (;;) {
        int64_t i = a->next->fetch_add(1);
        if (i >= a->n) break;
        a->out[i] = permanentExactSquare(wrap(a->mats + i * a->dim * a->dim, a->dim, a->dim));
    }
    return nullptr;
}
double ref_permanent_batch(const double* mats, int64_t dim, int64_t n, int nThreads, double* out) {
    std::atomic<int64_t> next(0);
    PermBatch a = {mats, dim, n, out, &next};
    if (nThreads < 1) nThreads = 1;
    std::vector<pthread_t> th(static_cast<size_t>(nThreads));
    auto t0 = tic();
    for (int t = 0; t < nThreads; t++) pthread_create(&th[size_t(t)], nullptr, permThread, &a);
    for (int t = 0; t < nThreads; t++) pthread_join(th[size_t(t)], nullptr);
    return toc(t0);
}

const char* ref_build_flags(void) {
#ifdef PDA_REF_FLAGS
    return PDA_REF_FLAGS;
#else
    return "unknown";
#endif
}

}  // extern "C"
