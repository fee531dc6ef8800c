/* oracle_quadric.c -- CPU restatement of the cost-matrix builders in front of the association path.
 * TEST INFRASTRUCTURE ONLY (see oracle_capi.h).
 *
 *   orc_quadric_covs         getCovs                  assignment.cpp:693-703
 *   orc_quadric_cost_matrix  computeQuadricCostMatrix assignment.cpp:705-722
 *   orc_association_from_moments  getAssignmentProbs  assignment.cpp:38-74 (usePerm == 0) from the moments on
 *
 * PARITY UNPINNED for the 3x3 solve: the reference calls Eigen's `(cov1+cov2).ldlt().solve(d)` (:716-717) and
 * Eigen (a CMake dependency of the reference, find_package(Eigen3), not vendored and not installed here) cannot be
 * compiled into oracle/_ref.  ldlt3_solve below restates the published algorithm of Eigen 3.4
 * (Eigen/src/Cholesky/LDLT.h: ldlt_inplace<Lower>::unblocked and LDLT::_solve_impl) as general loops over size = 3;
 * tests additionally check it against numpy.linalg.solve.  Everything else in this file is plain arithmetic.
 */
#include <float.h>
#include <math.h>
#include <stdlib.h>

#include "oracle_capi.h"

/* In-place LDLT of the lower triangle of the column-major n x n matrix A (n <= 4) with diagonal pivoting; returns the
 * transpositions.  Follows LDLT.h statement by statement, including the pivot search on the not yet updated diagonal. */
static void ldlt_lower_inplace(double* A, int n, int* tr, int* allZero) {
#define M(i, j) A[(i) + (j) * n]
    double temp[4];
    *allZero = 0;
    for (int k = 0; k < n; k++) {
        int idx = k;
        double best = fabs(M(k, k));
        for (int i = k + 1; i < n; i++)
            if (fabs(M(i, i)) > best) { best = fabs(M(i, i)); idx = i; } /* maxCoeff: first maximum */
        tr[k] = idx;
        if (k != idx) {
            const int s = n - idx - 1;
            for (int j = 0; j < k; j++) { double t = M(k, j); M(k, j) = M(idx, j); M(idx, j) = t; }             /* row heads   */
            for (int i = n - s; i < n; i++) { double t = M(i, k); M(i, k) = M(i, idx); M(i, idx) = t; }          /* column tails */
            { double t = M(k, k); M(k, k) = M(idx, idx); M(idx, idx) = t; }
            for (int i = k + 1; i < idx; i++) { double t = M(i, k); M(i, k) = M(idx, i); M(idx, i) = t; }
        }
        const int rs = n - k - 1;
        if (k > 0) {
            for (int j = 0; j < k; j++) temp[j] = M(j, j) * M(k, j);
            double dot = M(k, 0) * temp[0];
            for (int j = 1; j < k; j++) dot = dot + M(k, j) * temp[j];
            M(k, k) -= dot;
            for (int i = k + 1; i < n; i++) {
                double acc = M(i, 0) * temp[0];
                for (int j = 1; j < k; j++) acc = acc + M(i, j) * temp[j];
                M(i, k) -= acc;
            }
        }
        const double akk = M(k, k);
        const int valid = fabs(akk) > 0.0;
        if (k == 0 && !valid) { /* the matrix is entirely zero */
            for (int j = 0; j < n; j++) tr[j] = j;
            *allZero = 1;
            return;
        }
        if (rs > 0 && valid)
            for (int i = k + 1; i < n; i++) M(i, k) /= akk;
    }
#undef M
}

static void ldlt3_solve(const double* S /* 3x3 column-major, symmetric */, const double* d, double* x) {
    double A[9];
    int tr[3], zero;
    for (int i = 0; i < 9; i++) A[i] = S[i];
    ldlt_lower_inplace(A, 3, tr, &zero);
    double y[3] = {d[0], d[1], d[2]};
    for (int k = 0; k < 3; k++) { double t = y[k]; y[k] = y[tr[k]]; y[tr[k]] = t; }            /* dst = P b */
    if (!zero)
        for (int j = 0; j < 3; j++)                                                            /* unit-lower solve, column by column */
            for (int i = j + 1; i < 3; i++) y[i] -= y[j] * A[i + 3 * j];
    for (int i = 0; i < 3; i++) {                                                              /* pseudo-inverse of D */
        const double dii = zero ? 0.0 : A[i + 3 * i];
        y[i] = (fabs(dii) > DBL_MIN) ? y[i] / dii : 0.0;
    }
    if (!zero)
        for (int i = 2; i >= 0; i--) {                                                         /* unit-upper (L^T) solve, row by row */
            if (i == 2) continue;
            double acc = A[(i + 1) + 3 * i] * y[i + 1];
            for (int j = i + 2; j < 3; j++) acc = acc + A[j + 3 * i] * y[j];
            y[i] -= acc;
        }
    for (int k = 2; k >= 0; k--) { double t = y[k]; y[k] = y[tr[k]]; y[tr[k]] = t; }           /* dst = P^T dst */
    x[0] = y[0]; x[1] = y[1]; x[2] = y[2];
}

void orc_quadric_covs(const double* Q, int64_t n, double* covs) {
    for (int64_t q = 0; q < n; q++) {
        const double* m = Q + 16 * q; /* symmetric 4x4 */
        double* o = covs + 9 * q;
        const double q03 = m[3], q13 = m[7], q23 = m[11];
        o[0] = m[0] + pow(q03, 2); o[1] = m[1] + q03 * q13; o[2] = m[2] + q03 * q23;
        o[3] = m[1] + q03 * q13;   o[4] = m[5] + pow(q13, 2); o[5] = m[6] + q13 * q23;
        o[6] = m[2] + q03 * q23;   o[7] = m[6] + q13 * q23;   o[8] = m[10] + pow(q23, 2);
    }
}

void orc_quadric_cost_matrix(const double* landMean, const double* landCov, int64_t nL, const double* measMean,
                             const double* measCov, int64_t nM, double nonassign, double* costs) {
    const int64_t nRows = nL + nM;
    for (int64_t i = 0; i < nRows * nM; i++) costs[i] = INFINITY;
    for (int64_t col = 0; col < nM; col++) {
        for (int64_t row = 0; row < nL; row++) {
            double d[3], S[9], x[3];
            for (int i = 0; i < 3; i++) d[i] = landMean[3 * row + i] - measMean[3 * col + i];
            for (int i = 0; i < 9; i++) S[i] = landCov[9 * row + i] + measCov[9 * col + i];
            ldlt3_solve(S, d, x);
            costs[col * nRows + row] = d[0] * x[0] + (d[1] * x[1] + d[2] * x[2]);
        }
        costs[col * nRows + nL + col] = nonassign;
    }
}

int orc_association_from_moments(const double* landMean, const double* landCov, int64_t nL, const double* measMean,
                                 const double* measCov, int64_t nM, double nonassign, int64_t k, double* probs) {
    if (nM <= 0) return 0;
    if (nL == 0) { for (int64_t m = 0; m < nM; m++) probs[m] = 1.0; return 0; }
    double* costs = (double*)malloc((size_t)((nL + nM) * nM) * sizeof(double));
    orc_quadric_cost_matrix(landMean, landCov, nL, measMean, measCov, nM, nonassign, costs);
    const int rc = orc_association_probs(costs, nL, nM, k, 0, probs);
    free(costs);
    return rc;
}
