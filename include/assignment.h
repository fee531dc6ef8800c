/* assignment.h -- drop-in replacement for the numeric half of the reference's assignment.h
 * (reference: assignment.h:11-43; assignment.cpp:145-290 permanentProb, :325-435 conditionedPermanent,
 * :439-525 conditionCosts, :527-542 toProbs, :547-683 assignmentProb, :835-964 bruteForceProb).
 * Same names, same std::vector-based signatures, same return shapes: probs[nM][nL+1], the last entry
 * of each row being the non-assignment probability.  Everything runs as a batch of one on the B200
 * through libpda_b200.so; the *Batch forms underneath are what a multi-frame caller should use.
 *
 * Not here: the GTSAM / OpenCV typed entry points of the reference header (getAssignmentProbs, asgnBB,
 * computeQuadricCostMatrix, computeBBCostMatrix, getMeans, getCovs, saveAssignmentProb).  They only build
 * cost matrices and then call the functions below; keep the reference's definitions for them.
 */
#ifndef sensSLAM_assignment
#define sensSLAM_assignment

#include <stddef.h>

#include <vector>

#include "nwPerm.h"

std::vector<std::vector<double> > assignmentProb(const std::vector<double>& costMatrix, size_t nL, size_t nM, size_t k);

/* permOpt: 1 exact, 2 "long" (same double kernel, as in the reference); 0 (Huber approximation) and anything
 * else throw std::runtime_error. */
std::vector<std::vector<double> > permanentProb(std::vector<double> costMatrix, size_t nL, size_t nM, int permOpt);

std::vector<std::vector<double> > bruteForceProb(const std::vector<double>& costMatrix, size_t nL, size_t nM);

std::vector<double> conditionCosts(const std::vector<double>& costs, size_t nRows, size_t nCols,
                                   std::vector<ptrdiff_t>& rowIdxOut);

void toProbs(std::vector<double>& costMatrix);

/* The body of getAssignmentProbs (assignment.cpp:57-74) from the cost matrix on: conditionCosts, then
 * assignmentProb(k) or (usePerm) permanentProb(.., 1) on the conditioned problem, then the weights scattered back to
 * the original landmark indices.  costMatrix is what computeQuadricCostMatrix returns ((nL+nM) x nM, column-major).
 * In the reference's getAssignmentProbs, replace lines :57-74 by a call to this. */
std::vector<std::vector<double> > getAssignmentProbsFromCosts(const std::vector<double>& costMatrix, size_t nL, size_t nM,
                                                              size_t k, bool usePerm);

/* asgnBB (reference assignment.h:21, assignment.cpp:724-775) on raw boxes: five doubles per box
 * (xmin, ymin, xmax, ymax, xOffset), nonassign = runConsts.NONASSIGN_BOUNDBOX.  Returns, per left box, the index of
 * the right box it is paired with or -1.  A boundBox-typed asgnBB is a three-line wrapper over this. */
std::vector<int> asgnBBRaw(const std::vector<double>& boxesL, const std::vector<double>& boxesR, double nonassign);

/* raw form of conditionedPermanent: A is rows x cols, column-major */
double conditionedPermanentRaw(const double* A, size_t rows, size_t cols, int permOpt);
#ifdef PDA_HAVE_EIGEN
inline double conditionedPermanent(const Eigen::MatrixXd& A, int permOpt) {
    return conditionedPermanentRaw(A.data(), size_t(A.rows()), size_t(A.cols()), permOpt);
}
#endif

/* ---- batch forms (one call, many frames): what the GPU is built for --------------------------------
 * costs[p] is problem p's column-major (nL[p]+nM[p]) x nM[p] matrix. Returns probs[p][m][l]. */
std::vector<std::vector<std::vector<double> > > assignmentProbBatch(const std::vector<std::vector<double> >& costs,
                                                                    const std::vector<size_t>& nL,
                                                                    const std::vector<size_t>& nM, size_t k);

#endif
