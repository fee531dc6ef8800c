/* nwPerm.h -- drop-in replacement for the exact-permanent part of the reference's nwPerm.h
 * (reference: nwPerm.h:22-25; nwPerm.cpp:217-231 permanentExact, :251-332 permanentExactSquare,
 * :386-400 permanentExactLong).  Implemented as a batch of one over libpda_b200.so.
 *
 * With Eigen available the reference's own signatures (const Eigen::MatrixXd&) are declared; the raw
 * column-major forms below them are always available and are what the Eigen forms call.
 * Matrices above dimension 32 throw std::runtime_error with the reference's message (nwPerm.cpp:329).
 *
 * Huber's randomised approximation (permanentApproximation / permanentApproximationSquare, nwPerm.cpp:126-211, with
 * their sinkhorn / hl_factor / pickRowFromProbs helpers) is provided as well.  The reference draws from an unseeded
 * process-wide rand(); here the draws are counter-based and seeded (pda_b200.h: pda_permanent_approx_batch), so the
 * estimate is reproducible but agrees with the reference statistically, not sample for sample.  permanentFastest
 * (nwPerm.cpp:19-33, never called by the reference) is the obvious switch over the two and is provided inline.
 */
#ifndef sensSLAM_perm
#define sensSLAM_perm

#include <stddef.h>

#if defined(__has_include)
#if __has_include(<eigen3/Eigen/Core>)
#include <eigen3/Eigen/Core>
#define PDA_HAVE_EIGEN 1
#elif __has_include(<Eigen/Core>)
#include <Eigen/Core>
#define PDA_HAVE_EIGEN 1
#endif
#endif

/* raw forms: A is rows x cols, column-major */
double permanentExactRaw(const double* A, size_t rows, size_t cols);
long double permanentExactLongRaw(const double* A, size_t rows, size_t cols);
double permanentApproximationRaw(const double* A, size_t rows, size_t cols, size_t iterations);
/* ONE n x n matrix over several GPUs of this process: the Gray range of the NW walk (nwPerm.cpp:294-323) split across
 * `nDevices` CUDA device ordinals, partial sums combined on the host in device order (pda_permanent_sharded_host) */
double permanentExactShardedRaw(const double* A, size_t n, const int* devices, size_t nDevices);

#ifdef PDA_HAVE_EIGEN
inline double permanentExact(const Eigen::MatrixXd& A) { return permanentExactRaw(A.data(), size_t(A.rows()), size_t(A.cols())); }
inline double permanentExactSquare(const Eigen::MatrixXd& A) { return permanentExactRaw(A.data(), size_t(A.rows()), size_t(A.cols())); }
inline long double permanentExactLong(const Eigen::MatrixXd& A) { return permanentExactLongRaw(A.data(), size_t(A.rows()), size_t(A.cols())); }
inline double permanentApproximation(const Eigen::MatrixXd& A, size_t iterations) { return permanentApproximationRaw(A.data(), size_t(A.rows()), size_t(A.cols()), iterations); }
inline double permanentApproximationSquare(const Eigen::MatrixXd& A, size_t iterations) { return permanentApproximationRaw(A.data(), size_t(A.rows()), size_t(A.cols()), iterations); }
inline double permanentExactSharded(const Eigen::MatrixXd& A, const int* devices, size_t nDevices) { return permanentExactShardedRaw(A.data(), size_t(A.rows()), devices, nDevices); }
inline double permanentFastest(const Eigen::MatrixXd& A) {  /* nwPerm.cpp:19-33 */
    return (A.rows() > A.cols() ? A.rows() : A.cols()) <= 20 ? permanentExact(A) : permanentApproximation(A, 300);
}
#endif

#endif
