/* oracle_bbox.c -- CPU parity oracle for the stereo box association.
 * TEST INFRASTRUCTURE ONLY (see oracle_capi.h).  Restates
 *   boundBox::IoU        boundBox.h:62-75      (the x offset shifts only the box the method is called on)
 *   computeBBCostMatrix  assignment.cpp:777-797 ((nR+nL) x nL scores, -inf background, min of the two IoUs, dummy diagonal)
 *   asgnBB               assignment.cpp:724-775 (k = 1, MAXIMISE; a left box paired with a dummy row gets -1)
 * A box is five doubles: xmin, ymin, xmax, ymax, xOffset. */
#include "oracle_capi.h"

#include <math.h>
#include <stdlib.h>

static double box_area(const double* b) { return (b[2] - b[0]) * (b[3] - b[1]); }

static double box_iou(const double* self, const double* other) {
    const double l = fmax(self[0] + self[4], other[0]);
    const double r = fmin(self[2] + self[4], other[2]);
    const double t = fmax(self[1], other[1]);
    const double b = fmin(self[3], other[3]);
    if (l >= r || t >= b) return 0;
    const double inter = (r - l) * (b - t);
    return inter / (box_area(self) + box_area(other) - inter);
}

void orc_bb_cost_matrix(const double* boxesL, int64_t nL, const double* boxesR, int64_t nR, double nonassign, double* out) {
    const int64_t nRows = nR + nL;
    for (int64_t i = 0; i < nRows * nL; i++) out[i] = -INFINITY;
    for (int64_t c = 0; c < nL; c++) {
        for (int64_t r = 0; r < nR; r++) {
            const double i1 = box_iou(boxesR + 5 * r, boxesL + 5 * c), i2 = box_iou(boxesL + 5 * c, boxesR + 5 * r);
            out[c * nRows + r] = (i2 < i1) ? i2 : i1;  /* std::min(iou1, iou2) */
        }
        out[c * nRows + nR + c] = nonassign;
    }
}

void orc_asgn_bb(const double* boxesL, int64_t nL, const double* boxesR, int64_t nR, double nonassign, int32_t* out) {
    for (int64_t c = 0; c < nL; c++) out[c] = -1;
    if (nL == 0 || nR == 0) return;
    const int64_t nRows = nR + nL;
    double* C = (double*)malloc((size_t)(nRows * nL) * sizeof(double));
    int64_t* c4r = (int64_t*)malloc((size_t)nRows * sizeof(int64_t));
    int64_t* r4c = (int64_t*)malloc((size_t)nL * sizeof(int64_t));
    double g;
    orc_bb_cost_matrix(boxesL, nL, boxesR, nR, nonassign, C);
    if (orc_kbest2d(1, nRows, nL, 1, C, c4r, r4c, &g) > 0)
        for (int64_t c = 0; c < nL; c++) if (r4c[c] < nR) out[c] = (int32_t)r4c[c];
    free(C); free(c4r); free(r4c);
}
