// compmethods_driver.cpp -- a compMethods-style caller (reference: comparison.cpp:32-57 getCosts,
// :146-247 the per-frame call sequence) used to show that the drop-in headers really drop in.
//
// The SAME source is built twice:
//   -DDRIVER_USE_REFERENCE : against the reference's own code (oracle/Makefile, target ref_driver;
//                            binary oracle/_ref/compmethods_ref, built where /root/reference is mounted)
//   default                : against include/shortestPathCPP.hpp + include/assignment.h and
//                            libpda_b200_shims.so, i.e. the B200 path
// and tests/test_gpu_dropin.py compares the two outputs on the same .dat cost-matrix files
// (k-best lists and gains exactly, probabilities to 1e-9).
#ifdef DRIVER_USE_REFERENCE
#include "ref_prelude_assignment.h"
#else
#include "assignment.h"
#include "shortestPathCPP.hpp"
#endif

#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <limits>
#include <sstream>
#include <string>
#include <vector>

// the .dat wire format of saveAssignmentProb (assignment.cpp:821-831): one CSV line per row,
// std::to_string values, "inf" for missing arcs; read like comparison.cpp:39-54
static bool readDat(const std::string& path, std::vector<std::vector<double> >& rows) {
    std::ifstream f(path.c_str());
    if (!f) return false;
    std::string line;
    rows.clear();
    while (std::getline(f, line)) {
        if (line.empty()) continue;
        rows.push_back(std::vector<double>());
        std::stringstream ss(line);
        std::string val;
        while (std::getline(ss, val, ',')) {
            if (val[0] == 'i') rows.back().push_back(std::numeric_limits<double>::infinity());
            else rows.back().push_back(std::stod(val));
        }
    }
    return !rows.empty();
}

static void printProbs(const char* tag, const std::vector<std::vector<double> >& p) {
    for (size_t m = 0; m < p.size(); m++) {
        std::printf("%s %zu", tag, m);
        for (size_t l = 0; l < p[m].size(); l++) std::printf(" %.17g", p[m][l]);
        std::printf("\n");
    }
}

int main(int argc, char** argv) {
    size_t k = 200;
    std::vector<std::string> files;
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        if (a == "--k" && i + 1 < argc) k = size_t(std::atoi(argv[++i]));
        else files.push_back(a);
    }
    for (size_t fi = 0; fi < files.size(); fi++) {
        std::vector<std::vector<double> > costs;
        if (!readDat(files[fi], costs)) { std::printf("frame %zu unreadable\n", fi); continue; }
        const size_t nRows = costs.size(), nCols = costs[0].size();
        const size_t nM = nCols, nL = nRows - nM;
        std::vector<double> unrolled(nCols * nRows);
        for (size_t r = 0; r < nRows; r++)
            for (size_t c = 0; c < nCols; c++) unrolled[c * nRows + r] = costs[r][c];

        std::vector<ptrdiff_t> rowIdx;
        std::vector<double> cond = conditionCosts(unrolled, nL + nM, nM, rowIdx);
        const size_t condL = (cond.size() / nM) - nM;
        std::printf("frame %zu nL %zu nM %zu condL %zu rowIdx", fi, nL, nM, condL);
        for (size_t i = 0; i < rowIdx.size(); i++) std::printf(" %td", rowIdx[i]);
        std::printf("\n");

        if (nM > 1) {  // the k-best lists themselves (what assignmentProb consumes internally)
            const size_t nR = condL + nM;
            ScratchSpace workMem;
            workMem.init(nR, nR);
            std::vector<ptrdiff_t> c4r(nR * k), r4c(nM * k);
            std::vector<double> gains(k);
            const size_t found = kBest2DCutoff(k, nR, nM, false, cond.data(), workMem, c4r.data(), r4c.data(), gains.data(), 42.0);
            std::printf("kbest found %zu\n", found);
            for (size_t i = 0; i < found; i++) {
                std::printf("h %zu %.17g :", i, gains[i]);
                for (size_t c = 0; c < nM; c++) std::printf(" %td", r4c[i * nM + c]);
                std::printf(" |");
                for (size_t r = 0; r < nR; r++) std::printf(" %td", c4r[i * nR + r]);
                std::printf("\n");
            }
        }
        printProbs("truth", bruteForceProb(cond, condL, nM));
        const size_t ks[5] = {1, 20, 100, 200, k};
        for (int i = 0; i < 5; i++) {
            char tag[32];
            std::snprintf(tag, sizeof(tag), "k%zu", ks[i]);
            printProbs(tag, assignmentProb(cond, condL, nM, ks[i]));
        }
        if (condL + nM < 32) printProbs("permExact", permanentProb(cond, condL, nM, 1));  // comparison.cpp:231-235
    }
    return 0;
}
