// moments_driver.cpp -- exercises the C++ drop-in forms that have no line-for-line twin in the CPU reference build:
// computeQuadricCostMatrixRaw, getAssignmentProbsFromMoments (include/assignment.h) and permanentApproximationRaw
// (include/nwPerm.h).  Reads "nL nM nonassign k" and the flattened moments from stdin, prints full-precision results;
// tests/test_gpu_dropin.py compares them with the Python face of the same library and with the CPU oracle.
#include <cstdio>
#include <vector>

#include "assignment.h"

int main() {
    size_t nL, nM, k;
    double nonassign;
    if (scanf("%zu %zu %lf %zu", &nL, &nM, &nonassign, &k) != 4) return 2;
    std::vector<double> lm(3 * nL), lc(9 * nL), mm(3 * nM), mc(9 * nM);
    for (double& x : lm) if (scanf("%lf", &x) != 1) return 2;
    for (double& x : lc) if (scanf("%lf", &x) != 1) return 2;
    for (double& x : mm) if (scanf("%lf", &x) != 1) return 2;
    for (double& x : mc) if (scanf("%lf", &x) != 1) return 2;
    const std::vector<double> costs = computeQuadricCostMatrixRaw(lm, lc, mm, mc, nonassign);
    printf("costs");
    for (double c : costs) printf(" %.17g", c);
    printf("\n");
    const std::vector<std::vector<double> > probs = getAssignmentProbsFromMoments(lm, lc, mm, mc, nonassign, k);
    for (size_t m = 0; m < probs.size(); m++) {
        printf("probs %zu", m);
        for (double p : probs[m]) printf(" %.17g", p);
        printf("\n");
    }
    // a 6 x 6 matrix of ones has permanent 720
    const std::vector<double> ones(36, 1.0);
    printf("approx %.17g\n", permanentApproximationRaw(ones.data(), 6, 6, 300));
    return 0;
}
