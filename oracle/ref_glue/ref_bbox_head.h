// Head of the translation unit that oracle/Makefile assembles for the stereo box association
// (asgnBB, computeBBCostMatrix: assignment.cpp:724-797) -- TEST INFRASTRUCTURE ONLY.
// The reference's boundBox wraps gtsam_quadrics::AlignedBox2 and OpenCV types (boundBox.h:5-7), neither installed;
// this stand-in keeps exactly the members that IoU and the two functions touch.  The body of IoU itself is NOT
// restated: oracle/Makefile splices boundBox.h lines 62-75 in right after this file, inside the class.
#include <algorithm>
#include <cstddef>
#include <iostream>
#include <limits>
#include <vector>

#include "shortestPathCPP.hpp"  // the reference's own header (-I/root/reference)

#define inf_d std::numeric_limits<double>::infinity()

struct RefAlignedBox2 {
    double x0, y0, x1, y1;
    double xmin() const { return x0; }
    double ymin() const { return y0; }
    double xmax() const { return x1; }
    double ymax() const { return y1; }
    double width() const { return x1 - x0; }
    double height() const { return y1 - y0; }
};

struct semConsts { double NONASSIGN_BOUNDBOX; };

class boundBox {
public:
    RefAlignedBox2 aBox;
    double xOffset;
    double xmin() const { return aBox.xmin(); }
    double xmax() const { return aBox.xmax(); }
    double ymin() const { return aBox.ymin(); }
    double ymax() const { return aBox.ymax(); }
    double area() const { return aBox.width() * aBox.height(); }
    // ---- boundBox.h:62-75 (double IoU(const boundBox& other) const {...}) is spliced in below ----
