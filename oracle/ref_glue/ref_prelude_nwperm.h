// Prelude for the translation unit assembled at build time from
// /root/reference/nwPerm.cpp lines 217-231, 251-332 and 386-400
// (permanentExact, permanentExactSquare, permanentExactLong).
// TEST INFRASTRUCTURE ONLY.
#ifndef PDA_REF_PRELUDE_NWPERM
#define PDA_REF_PRELUDE_NWPERM
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <stdexcept>
#include <Eigen/Core>  // oracle/ref_glue/eigen_shim
double permanentExact(const Eigen::MatrixXd& A);
long double permanentExactLong(const Eigen::MatrixXd& A);
double permanentExactSquare(const Eigen::MatrixXd& A);
#endif
