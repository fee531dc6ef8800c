};  // class boundBox

std::vector<int> asgnBB(const std::vector<boundBox>& bbL, const std::vector<boundBox>& bbR, const semConsts& runConsts);
std::vector<double> computeBBCostMatrix(const std::vector<boundBox>& bbL, const std::vector<boundBox>& bbR, const semConsts& runConsts);

// C door (boxes are 5 doubles each: xmin, ymin, xmax, ymax, xOffset)
static std::vector<boundBox> refBoxes(const double* b, long n) {
    std::vector<boundBox> v(static_cast<size_t>(n));
    for (long i = 0; i < n; i++) { v[i].aBox = RefAlignedBox2{b[5 * i], b[5 * i + 1], b[5 * i + 2], b[5 * i + 3]}; v[i].xOffset = b[5 * i + 4]; }
    return v;
}
extern "C" void ref_asgn_bb(const double* boxesL, long nL, const double* boxesR, long nR, double nonassign, int* out) {
    semConsts c = {nonassign};
    std::vector<int> a = asgnBB(refBoxes(boxesL, nL), refBoxes(boxesR, nR), c);
    for (size_t i = 0; i < a.size(); i++) out[i] = a[i];
}
extern "C" void ref_bb_cost_matrix(const double* boxesL, long nL, const double* boxesR, long nR, double nonassign, double* out) {
    semConsts c = {nonassign};
    std::vector<double> m = computeBBCostMatrix(refBoxes(boxesL, nL), refBoxes(boxesR, nR), c);
    std::copy(m.begin(), m.end(), out);
}
// ---- assignment.cpp:724-797 (asgnBB, computeBBCostMatrix) is spliced in below ----
