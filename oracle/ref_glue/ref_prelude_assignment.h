// Prelude for the translation unit that oracle/Makefile assembles, at build time
// and outside the repository, from line ranges of /root/reference/assignment.cpp
// (9-11, 28-36, 145-290, 292-323, 325-435, 439-525, 527-542, 547-683, 835-964).
// It supplies what the reference gets from assignment.h / constsUtils.h /
// nwPerm.h without dragging in GTSAM, OpenCV or Eigen proper (none installed).
// TEST INFRASTRUCTURE ONLY.
#ifndef PDA_REF_PRELUDE_ASSIGNMENT
#define PDA_REF_PRELUDE_ASSIGNMENT

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstddef>
#include <iostream>
#include <limits>
#include <numeric>
#include <stdexcept>
#include <string>
#include <vector>

#include <Eigen/Core>           // oracle/ref_glue/eigen_shim
#include "shortestPathCPP.hpp"  // the reference's own header (-I/root/reference)

// constsUtils.h:10, 18-21
#define inf_d std::numeric_limits<double>::infinity()
inline std::chrono::high_resolution_clock::time_point tic()
{ return std::chrono::high_resolution_clock::now(); }
inline double toc(const std::chrono::high_resolution_clock::time_point& t2)
{ return std::chrono::duration<double>(std::chrono::high_resolution_clock::now() - t2).count(); }

// assignment.h:11-43 (the GTSAM-free subset)
std::vector<std::vector<double> > assignmentProb(const std::vector<double>& costMatrix, size_t nL, size_t nM, size_t k);
std::vector<std::vector<double> > permanentProb(std::vector<double> costMatrix, size_t nL, size_t nM, int permOpt);
void setupAssgnMatrix(Eigen::MatrixXd& subProbs, const Eigen::MatrixXd& elProbs, size_t col);
double conditionedPermanent(const Eigen::MatrixXd& A, int permOpt);
void toProbs(std::vector<double>& costMatrix);
std::vector<double> conditionCosts(const std::vector<double>& costs, size_t nRows, size_t nCols, std::vector<ptrdiff_t>& rowIdxOut);
std::vector<std::vector<double> > bruteForceProb(const std::vector<double>& costMatrix, size_t nL, size_t nM);
double permWAssignments(const Eigen::MatrixXd& A);  // declared by the reference, defined nowhere

// nwPerm.h:18-25
double permanentApproximation(const Eigen::MatrixXd& A, size_t iterations);  // Huber path: out of scope, stubbed in ref_capi.cpp
double permanentExact(const Eigen::MatrixXd& A);
long double permanentExactLong(const Eigen::MatrixXd& A);
double permanentExactSquare(const Eigen::MatrixXd& A);

#endif
