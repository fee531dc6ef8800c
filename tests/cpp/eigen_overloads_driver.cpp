// eigen_overloads_driver.cpp -- calls the reference-TYPED overloads of the drop-in headers (const Eigen::MatrixXd&,
// std::vector<Eigen::Vector3d>, const semConsts&), which exist only when an <Eigen/Core> is on the include path.  Built
// against tests/cpp/eigen_stub (a stand-in with Eigen's storage order), so these signatures are compiled in every build
// and run on the GPU box by tests/test_gpu_dropin.py.  Prints full-precision values, one tagged line each.
#include <cstdio>
#include <stdexcept>
#include <vector>

#include "assignment.h"
#include "nwPerm.h"

#ifndef PDA_HAVE_EIGEN
#error "the Eigen-typed overloads were not enabled: <Eigen/Core> not found on the include path"
#endif

// the one member of the reference's semConsts (constsUtils.h:24-42) this path reads
struct semConsts { double NONASSIGN_QUADRIC; size_t k; };

int main() {
    // permanentExact / Square / Long on a 5 x 5 and a 3 x 5 matrix (nwPerm.h:22-25 of the reference)
    Eigen::MatrixXd A(5, 5), B(3, 5);
    for (int j = 0; j < 5; j++) for (int i = 0; i < 5; i++) A(i, j) = 0.25 + 0.5 * ((7 * i + 3 * j) % 5);
    for (int j = 0; j < 5; j++) for (int i = 0; i < 3; i++) B(i, j) = 1.0 + 0.125 * ((5 * i + 2 * j) % 7);
    printf("permanentExact %.17g\n", permanentExact(A));
    printf("permanentExactSquare %.17g\n", permanentExactSquare(A));
    printf("permanentExactLong %.17Lg\n", permanentExactLong(A));
    printf("permanentExactRect %.17g\n", permanentExact(B));
    printf("permanentExactLongRect %.17Lg\n", permanentExactLong(B));
    printf("conditionedPermanent %.17g\n", conditionedPermanent(B, 1));
    printf("conditionedPermanentLong %.17g\n", conditionedPermanent(B, 2));
    const int dev0 = 0;
    printf("permanentExactSharded %.17g\n", permanentExactSharded(A, &dev0, 1));
    bool threw = false;
    try { Eigen::MatrixXd big(33, 33); permanentExact(big); } catch (const std::runtime_error&) { threw = true; }
    printf("throwsAbove32 %d\n", threw ? 1 : 0);

    // computeQuadricCostMatrix with the reference's argument types, runConsts included (assignment.h:31-32)
    std::vector<Eigen::Vector3d> m1(4), m2(2);
    std::vector<Eigen::Matrix<double, 3, 3> > c1(4), c2(2);
    for (size_t i = 0; i < m1.size(); i++) {
        for (int d = 0; d < 3; d++) m1[i](d) = 2.0 * (double)i + 0.3 * d;
        for (int d = 0; d < 3; d++) c1[i](d, d) = 0.5 + 0.1 * (double)i + 0.05 * d;
        c1[i](0, 1) = c1[i](1, 0) = 0.02;
    }
    for (size_t i = 0; i < m2.size(); i++) {
        for (int d = 0; d < 3; d++) m2[i](d) = 2.0 * (double)i + 0.4 + 0.2 * d;
        for (int d = 0; d < 3; d++) c2[i](d, d) = 0.4 + 0.07 * d;
        c2[i](1, 2) = c2[i](2, 1) = -0.03;
    }
    semConsts runConsts = {10.0, 200};
    const std::vector<double> viaConsts = computeQuadricCostMatrix(m1, c1, m2, c2, runConsts);
    const std::vector<double> viaDouble = computeQuadricCostMatrix(m1, c1, m2, c2, 10.0);
    printf("quadricCosts");
    for (double x : viaConsts) printf(" %.17g", x);
    printf("\nquadricCostsSame %d\n", viaConsts == viaDouble ? 1 : 0);
    return 0;
}
