// Definitions for the two symbols the extracted reference code refers to but that lie outside the
// parity oracle's scope.  TEST INFRASTRUCTURE ONLY (oracle/_ref builds).
#include "ref_prelude_assignment.h"

// ---- symbols the extracted reference code refers to but that are out of scope
double permWAssignments(const Eigen::MatrixXd&) {
    // assignment.h:41 declares it, no definition exists anywhere in the reference;
    // it is only reachable from `if(verbose)` blocks with verbose == false.
    throw std::runtime_error("permWAssignments: not defined by the reference");
}
double permanentApproximation(const Eigen::MatrixXd&, size_t) {
    // Huber's randomised approximation (nwPerm.cpp:126-211, unseeded rand()):
    // out of scope per SURVEY.md section 2.
    throw std::runtime_error("permanentApproximation: out of scope for the parity oracle");
}

