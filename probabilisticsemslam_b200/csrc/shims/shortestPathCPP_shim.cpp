// shortestPathCPP_shim.cpp -- the reference's shortestPathCPP.hpp entry points as batch-of-one calls
// into libpda_b200.so.  No solver lives here: this file only marshals arguments and reproduces the
// reference's return-value / side-effect conventions (shortestPathCPP.cpp:119-238, 571-762).
#include "shortestPathCPP.hpp"

#include <stdint.h>

#include <algorithm>
#include <stdexcept>
#include <string>
#include <vector>

#include "pda_b200.h"

static int g_device = 0;
void pdaSetDevice(int device) { g_device = device; }
int pdaShimDevice() { return g_device; }

void pdaCheck(int rc, const char* where) {
    if (rc != PDA_OK) throw std::runtime_error(std::string(where) + ": " + pda_last_error());
}

static_assert(sizeof(ptrdiff_t) == sizeof(int64_t), "the C ABI carries indices as int64");

static size_t kbest(const size_t k, const size_t numRow, const size_t numCol, const bool maximize, const double* C,
                    ScratchSpace& workMem, ptrdiff_t* col4rowBest, ptrdiff_t* row4colBest, double* gainBest,
                    int cutMode, double cutoff, bool cutMaximize) {
    if (k == 0) return 0;
    const int64_t costOff = 0, r4cOff = 0, c4rOff = 0;
    const int32_t nr = int32_t(numRow), nc = int32_t(numCol);
    int32_t nFound = 0;
    pdaCheck(pda_murty_batch_host(C, &costOff, &nr, &nc, 1, int32_t(k), cutMode, cutoff, maximize ? 1 : 0, cutMaximize ? 1 : 0,
                                  reinterpret_cast<int64_t*>(row4colBest), &r4cOff, reinterpret_cast<int64_t*>(col4rowBest),
                                  &c4rOff, gainBest, &nFound, PDA_WEIGHTS_NONE, NULL, NULL, NULL, g_device),
             "kBest2D");
    (void)workMem;
    return size_t(nFound);
}

size_t kBest2D(const size_t k, const size_t numRow, const size_t numCol, const bool maximize, const double* C,
               ScratchSpace& workMem, ptrdiff_t* col4rowBest, ptrdiff_t* row4colBest, double* gainBest) {
    // a ScratchSpace that went through kBest2DCutoff keeps pruning (toCut is never cleared by the reference)
    if (workMem.toCut)
        return kbest(k, numRow, numCol, maximize, C, workMem, col4rowBest, row4colBest, gainBest, PDA_CUT_STICKY,
                     workMem.cutoffGain, workMem.maximize);
    return kbest(k, numRow, numCol, maximize, C, workMem, col4rowBest, row4colBest, gainBest, PDA_CUT_NONE, 0.0, false);
}

size_t kBest2DCutoff(const size_t k, const size_t numRow, const size_t numCol, const bool maximize, const double* C,
                     ScratchSpace& workMem, ptrdiff_t* col4rowBest, ptrdiff_t* row4colBest, double* gainBest,
                     double cutoff) {
    workMem.toCut = true;          // shortestPathCPP.cpp:650-651
    workMem.maximize = maximize;
    const size_t found = kbest(k, numRow, numCol, maximize, C, workMem, col4rowBest, row4colBest, gainBest,
                               PDA_CUT_RELATIVE, cutoff, maximize);
    if (found > 0) {
        // Side effect the reference leaves behind: cutoffGain = gain0(shifted) +/- cutoff (:681, :684), where
        // gain0(shifted) is the root hypothesis summed over the shifted matrix in column order (:72-79).
        // Re-derived here from the returned root assignment with the same operand order, so it is bit-exact.
        const size_t numEl = numRow * numCol;
        const double d = maximize ? *std::max_element(C, C + numEl) : *std::min_element(C, C + numEl);
        double g0 = 0;
        for (size_t c = 0; c < numCol; c++) {
            const double e = C[c * numRow + size_t(row4colBest[c])];
            g0 = g0 + (maximize ? (-e + d) : (e - d));
        }
        workMem.cutoffGain = maximize ? g0 - cutoff : g0 + cutoff;
    }
    return found;
}

static int lap(MurtyHyp* sol, const double* C, size_t numRow, size_t numCol, size_t numCol4Gain, bool makeSafe, bool maximize) {
    const int64_t costOff = 0, rowOff = 0, colOff = 0;
    const int32_t nr = int32_t(numRow), nc = int32_t(numCol), ng = int32_t(numCol4Gain);
    std::vector<uint8_t> forb(numRow ? numRow : 1);
    int32_t feasible = 0;
    double gain = 0;
    pdaCheck(pda_lap_batch_host(C, &costOff, &nr, &nc, &ng, 1, makeSafe ? 1 : 0, maximize ? 1 : 0, &rowOff, &colOff,
                                reinterpret_cast<int64_t*>(sol->col4row), reinterpret_cast<int64_t*>(sol->row4col), sol->u,
                                sol->v, forb.data(), &gain, &feasible, g_device),
             "shortestPathCPP");
    for (size_t r = 0; r < numRow; r++) sol->forbiddenActiveRows[r] = forb[r] != 0;
    sol->gain = feasible ? gain : -1;
    sol->activeCol = 0;
    sol->solved = true;
    return feasible;
}

int shortestPathCPP(MurtyHyp* problemSol, ScratchSpace& workMem, const size_t numRow, const size_t numCol,
                    const size_t numCol4Gain) {
    return lap(problemSol, workMem.C, numRow, numCol, numCol4Gain, false, false) ? 0 : 1;
}

int assign2D(const size_t numRow, const size_t numCol, const bool maximize, const double* C, ScratchSpace& workMem,
             MurtyHyp* problemSol) {
    (void)workMem;
    return lap(problemSol, C, numRow, numCol, numCol, true, maximize) ? 1 : 0;
}
