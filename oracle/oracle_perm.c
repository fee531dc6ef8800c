/* oracle_perm.c -- CPU parity oracle for the matrix permanent.  TEST INFRASTRUCTURE
 * ONLY (see oracle_capi.h).
 *
 * Restates, on raw column-major arrays, the reference's Nijenhuis-Wilf / Ryser
 * Gray-code permanent:
 *   permanentExactSquare  nwPerm.cpp:251-332
 *   permanentExact        nwPerm.cpp:217-231  (rectangular: pad with ones, divide by (|m-n|)!)
 *   permanentExactLong    nwPerm.cpp:386-400  (same double kernel, final divide in long double)
 * The walk is kept strictly sequential with the reference's operand order so the
 * result is bit-identical to an IEEE-strict build of the reference.
 */
#include "oracle_capi.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* Sequential NW walk over all 2^(n-1) Gray-code column subsets. 1 <= n <= 32. */
static double nw_walk(const double* a, int n) {
    double x[32];
    double p = 1.0;
    for (int j = 0; j < n; j++) {
        double rs = 0.0;
        for (int k = 0; k < n; k++) rs += a[j + k * n];
        x[j] = a[j + (n - 1) * n] - rs / 2;
        p *= x[j];
    }
    const uint64_t last = ((uint64_t)1 << (n - 1)) - 1;
    for (uint64_t i = 1; i <= last; i++) {
        /* the bit in which gray(i) differs from gray(i-1) is the lowest set bit of i */
        const int k = __builtin_ctzll(i);
        const uint64_t gray = i ^ (i >> 1);
        const double s = ((gray >> k) & 1) ? +1.0 : -1.0;
        const double* col = a + (size_t)k * (size_t)n;
        double prod = 1.0;
        for (int j = 0; j < n; j++) {
            x[j] += s * col[j];
            prod *= x[j];
        }
        p += ((i & 1) ? -1.0 : 1.0) * prod;
    }
    return (double)(4 * (n & 1) - 2) * p;
}

double orc_permanent_exact_square(const double* A, int64_t n, int* status) {
    *status = 0;
    if (n == 0) return 1.0;
    if (n > 32) { *status = 1; return 0.0; }  /* nwPerm.cpp:327-330 throws */
    return nw_walk(A, (int)n);
}

/* pad an m x n matrix with ones to dim x dim (nwPerm.cpp:226-228) */
static double* pad_with_ones(const double* A, int64_t m, int64_t n, int64_t dim) {
    double* P = (double*)malloc((size_t)(dim * dim) * sizeof(double));
    for (int64_t i = 0; i < dim * dim; i++) P[i] = 1.0;
    for (int64_t c = 0; c < n; c++)
        for (int64_t r = 0; r < m; r++) P[r + c * dim] = A[r + c * m];
    return P;
}

double orc_permanent_exact(const double* A, int64_t rows, int64_t cols, int* status) {
    if (rows == cols) return orc_permanent_exact_square(A, rows, status);
    const int m = (int)rows, n = (int)cols;
    const double scale = tgamma(abs(m - n) + 1);
    const int64_t dim = rows > cols ? rows : cols;
    double* P = pad_with_ones(A, rows, cols, dim);
    double r = orc_permanent_exact_square(P, dim, status);
    free(P);
    return *status ? 0.0 : r / scale;
}

static long double permanent_exact_long(const double* A, int64_t rows, int64_t cols, int* status) {
    if (rows == cols) return orc_permanent_exact_square(A, rows, status);
    const int m = (int)rows, n = (int)cols;
    const long double scale = tgamma(abs(m - n) + 1);
    const int64_t dim = rows > cols ? rows : cols;
    double* P = pad_with_ones(A, rows, cols, dim);
    double r = orc_permanent_exact_square(P, dim, status);
    free(P);
    return *status ? 0.0L : r / scale;
}

double orc_permanent_exact_long(const double* A, int64_t rows, int64_t cols, int* status) {
    return (double)permanent_exact_long(A, rows, cols, status);
}

/* conditionedPermanent (assignment.cpp:325-435): drop all-zero rows and columns,
 * scale every kept column by 1/sqrt(max * smallest-nonzero), take the permanent of
 * the transpose, undo the scaling; if cancellation made it negative, retry
 * untransposed (:409-419).
 *
 * Deviation, documented: with an all-zero COLUMN the reference reads colsIdx[] past
 * its end and addresses Ascaled by the uncompacted column index (:384-392, undefined
 * behaviour).  Here zero columns are compacted away, which is
 * what the code evidently intends; parity tests avoid such inputs. */
/* permOpt == 0 (Huber's approximation, apprxIter = 300, assignment.cpp:10/:401) draws from the counter-based stream of
 * oracle_perm_approx.c; the stream's matrix index is set by the caller (the item index m*(nL+1)+l of permanentProb, 0 for
 * a plain conditionedPermanent call), mirroring how the CUDA pipeline numbers its items. */
static __thread int64_t g_approxIndex = 0;
static uint64_t g_approxSeed = 20260217ULL;
void orc_set_approx_stream(uint64_t seed, int64_t index) { g_approxSeed = seed; g_approxIndex = index; }

double orc_conditioned_permanent(const double* A, int64_t rows, int64_t cols, int permOpt, int* status) {
    *status = 0;
    if (permOpt < 0 || permOpt > 2) { *status = 1; return 0.0; }  /* the reference throws (:406) */
    int64_t* keepC = (int64_t*)malloc((size_t)(cols > 0 ? cols : 1) * sizeof(int64_t));
    int64_t* keepR = (int64_t*)malloc((size_t)(rows > 0 ? rows : 1) * sizeof(int64_t));
    double* colScale = (double*)malloc((size_t)(cols > 0 ? cols : 1) * sizeof(double));
    int64_t nKC = 0, nKR = 0;
    for (int64_t c = 0; c < cols; c++) {
        double mx = A[c * rows], mn = 1;
        for (int64_t r = 0; r < rows; r++) {
            double e = A[r + c * rows];
            if (e > mx) mx = e;
            if (e > 0 && e < mn) mn = e;
        }
        if (mx > 0) { colScale[nKC] = 1.0 / pow(mx * mn, 0.5); keepC[nKC++] = c; }
    }
    for (int64_t r = 0; r < rows; r++) {
        double mx = A[r];
        for (int64_t c = 1; c < cols; c++) if (A[r + c * rows] > mx) mx = A[r + c * rows];
        if (mx > 0) keepR[nKR++] = r;
    }
    double scaleFactor = 1;
    /* S = scaled, compacted matrix (nKR x nKC); St = its transpose (nKC x nKR) */
    double* S = (double*)malloc((size_t)(nKR * nKC > 0 ? nKR * nKC : 1) * sizeof(double));
    double* St = (double*)malloc((size_t)(nKR * nKC > 0 ? nKR * nKC : 1) * sizeof(double));
    for (int64_t j = 0; j < nKC; j++) {
        scaleFactor *= colScale[j];
        for (int64_t i = 0; i < nKR; i++) {
            double e = colScale[j] * A[keepR[i] + keepC[j] * rows];
            S[i + j * nKR] = e;
            St[j + i * nKC] = e;
        }
    }
    double result;
    if (permOpt == 0) {
        const int64_t dim = nKC > nKR ? nKC : nKR;
        if (dim > 32) *status = 1;  /* device limit of the approximation kernel (the reference has none) */
        result = *status ? 0.0 : orc_permanent_approx(St, nKC, nKR, 300, g_approxSeed, g_approxIndex, NULL) / scaleFactor;
    } else if (permOpt == 1) {
        result = orc_permanent_exact(St, nKC, nKR, status) / scaleFactor;
        if (!*status && result < 0) result = orc_permanent_exact(S, nKR, nKC, status) / scaleFactor;
    } else {
        result = (double)(permanent_exact_long(St, nKC, nKR, status) / scaleFactor);
        if (!*status && result < 0) result = (double)(permanent_exact_long(S, nKR, nKC, status) / scaleFactor);
    }
    free(keepC); free(keepR); free(colScale); free(S); free(St);
    return *status ? 0.0 : result;
}
