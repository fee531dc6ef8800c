/* shortestPathCPP.hpp -- drop-in replacement for the reference's header of the same name
 * (reference: shortestPathCPP.hpp:22-65 MurtyHyp, :73-142 ScratchSpace, :144-149 assign2D,
 * :178-182 shortestPathCPP, :204-212 kBest2D, :256-265 kBest2DCutoff).
 *
 * Same class names, member names, function names, argument order, ownership and return-value
 * conventions, so comparison.cpp / system.cpp / assignment.cpp style callers compile unchanged.
 * The implementation behind it (probabilisticsemslam_b200/csrc/shims/shortestPathCPP_shim.cpp) has no
 * solver of its own: every call is a batch of one through the C ABI of libpda_b200.so
 * (include/pda_b200.h) and therefore runs on the B200.  Without a CUDA device the functions throw
 * std::runtime_error -- there is no CPU fallback.
 *
 * Conventions kept from the reference:
 *   - C is column-major, numRow x numCol, numRow >= numCol; outputs are caller-allocated:
 *     col4rowBest[numRow*k], row4colBest[numCol*k], gainBest[k]; only the first <return value>
 *     hypotheses are defined (for kBest2DCutoff the slot that triggered the cutoff break is written too).
 *   - kBest2D / kBest2DCutoff return the number of hypotheses found, 0 = infeasible.
 *   - shortestPathCPP returns 1 if infeasible (gain = -1), else 0; assign2D returns 0 if infeasible, 1 otherwise.
 *   - kBest2DCutoff leaves toCut / cutoffGain / maximize set in the ScratchSpace and kBest2D honours
 *     them, exactly like the reference (hpp:84-86, 130-132; cpp:650-651).
 */
#ifndef SPALGS
#define SPALGS
#include <stddef.h>

#include <vector>

class MurtyHyp {
    std::vector<char> storage_;

public:
    ptrdiff_t* col4row;
    ptrdiff_t* row4col;
    double gain;
    double* u;
    double* v;
    size_t activeCol;
    bool* forbiddenActiveRows;
    bool solved;

    MurtyHyp() : col4row(NULL), row4col(NULL), gain(0), u(NULL), v(NULL), activeCol(0), forbiddenActiveRows(NULL), solved(false) {}
    MurtyHyp(const size_t numRow, const size_t numCol) : gain(0), activeCol(0), solved(false) { bind(numRow, numCol); }
    MurtyHyp(const MurtyHyp&) = delete;
    MurtyHyp& operator=(const MurtyHyp&) = delete;

private:
    void bind(size_t numRow, size_t numCol) {
        const size_t idx = sizeof(ptrdiff_t), dbl = sizeof(double);
        storage_.assign(idx * (numRow + numCol) + dbl * (numRow + numCol) + numRow * sizeof(bool) + 16, 0);
        char* p = storage_.data();
        col4row = reinterpret_cast<ptrdiff_t*>(p); p += idx * numRow;
        row4col = reinterpret_cast<ptrdiff_t*>(p); p += idx * numCol;
        u = reinterpret_cast<double*>(p); p += dbl * numCol;
        v = reinterpret_cast<double*>(p); p += dbl * numRow;
        forbiddenActiveRows = reinterpret_cast<bool*>(p);
    }
};

class ScratchSpace {
    std::vector<char> storage_;

public:
    char* buffer;
    double* C;  /* numRow x numCol working copy of the (shifted) cost matrix: input of shortestPathCPP() */
    /* The remaining arrays are the reference solver's private scratch.  They are allocated with the
     * reference's sizes so that code poking at them keeps compiling, but the B200 solver does not use them. */
    size_t* ScannedColIdx;
    bool* ScannedRows;
    size_t* pred;
    double* shortestPathCost;
    ptrdiff_t* Row2ScanParent;
    ptrdiff_t* Row2Scan;
    bool* forbiddenActiveRows;
    bool toCut;
    double cutoffGain;
    bool maximize;

    ScratchSpace() : buffer(NULL), C(NULL), ScannedColIdx(NULL), ScannedRows(NULL), pred(NULL), shortestPathCost(NULL),
                     Row2ScanParent(NULL), Row2Scan(NULL), forbiddenActiveRows(NULL), toCut(false), cutoffGain(0), maximize(false) {}
    ScratchSpace(const size_t numRow, const size_t numCol) { this->init(numRow, numCol); }
    ScratchSpace(const ScratchSpace&) = delete;
    ScratchSpace& operator=(const ScratchSpace&) = delete;

    void init(const size_t numRow, const size_t numCol) {
        const size_t w = sizeof(size_t), d = sizeof(double);
        storage_.assign(numCol * w + 3 * numRow * w + numRow * (1 + numCol) * d + 2 * numRow * sizeof(bool) + 16, 0);
        char* p = storage_.data();
        buffer = p;
        ScannedColIdx = reinterpret_cast<size_t*>(p); p += w * numCol;
        Row2ScanParent = reinterpret_cast<ptrdiff_t*>(p); p += w * numRow;
        Row2Scan = reinterpret_cast<ptrdiff_t*>(p); p += w * numRow;
        pred = reinterpret_cast<size_t*>(p); p += w * numRow;
        shortestPathCost = reinterpret_cast<double*>(p); p += d * numRow;
        C = reinterpret_cast<double*>(p); p += d * numRow * numCol;
        ScannedRows = reinterpret_cast<bool*>(p); p += numRow * sizeof(bool);
        forbiddenActiveRows = reinterpret_cast<bool*>(p);
        toCut = false;
    }

    inline bool cutHyp(double gain) { return toCut ? (maximize ? gain < cutoffGain : gain > cutoffGain) : false; }
};

/* 2D assignment on an unpadded numRow x numCol matrix; results in problemSol (built as MurtyHyp(numRow, numCol)).
 * Returns 1 if solved, 0 if no finite-cost assignment exists. */
int assign2D(const size_t numRow, const size_t numCol, const bool maximize, const double* C, ScratchSpace& workMem,
             MurtyHyp* problemSol);

/* Shortest-augmenting-path LAP on workMem.C, which must already be non-negative ("safe"), minimisation.
 * numCol4Gain = number of leading columns summed into the gain.  Returns 1 if infeasible (gain = -1), else 0. */
int shortestPathCPP(MurtyHyp* problemSol, ScratchSpace& workMem, const size_t numRow, const size_t numCol,
                    const size_t numCol4Gain);

/* Murty k-best.  workMem must have been init(numRow, numRow).  Returns the number of hypotheses found. */
size_t kBest2D(const size_t k, const size_t numRow, const size_t numCol, const bool maximize, const double* C,
               ScratchSpace& workMem, ptrdiff_t* col4rowBest, ptrdiff_t* row4colBest, double* gainBest);

/* Same, but hypotheses more than `cutoff` worse than the best are neither kept nor returned. */
size_t kBest2DCutoff(const size_t k, const size_t numRow, const size_t numCol, const bool maximize, const double* C,
                     ScratchSpace& workMem, ptrdiff_t* col4rowBest, ptrdiff_t* row4colBest, double* gainBest,
                     double cutoff);

template <class T>
void increment(T& x) {
    x++;
}

/* Device used by the batch-of-one shims (default 0). */
void pdaSetDevice(int device);

#endif
