/* nwPerm.h -- drop-in replacement for the exact-permanent part of the reference's nwPerm.h
 * (reference: nwPerm.h:22-25; nwPerm.cpp:217-231 permanentExact, :251-332 permanentExactSquare,
 * :386-400 permanentExactLong).  Implemented as a batch of one over libpda_b200.so.
 *
 * With Eigen available the reference's own signatures (const Eigen::MatrixXd&) are declared; the raw
 * column-major forms below them are always available and are what the Eigen forms call.
 * Matrices above dimension 32 throw std::runtime_error with the reference's message (nwPerm.cpp:329).
 *
 * Not provided: Huber's randomised approximation (permanentApproximation*, sinkhorn, hl_factor,
 * permanentFastest; nwPerm.cpp:19-211) -- it is driven by an unseeded rand() and is outside the
 * accelerated path; keep the reference's nwPerm.cpp for those symbols.
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

#ifdef PDA_HAVE_EIGEN
inline double permanentExact(const Eigen::MatrixXd& A) { return permanentExactRaw(A.data(), size_t(A.rows()), size_t(A.cols())); }
inline double permanentExactSquare(const Eigen::MatrixXd& A) { return permanentExactRaw(A.data(), size_t(A.rows()), size_t(A.cols())); }
inline long double permanentExactLong(const Eigen::MatrixXd& A) { return permanentExactLongRaw(A.data(), size_t(A.rows()), size_t(A.cols())); }
#endif

#endif
