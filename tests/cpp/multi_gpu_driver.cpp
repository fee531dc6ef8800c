// multi_gpu_driver.cpp -- a C++ caller of the multi-device entry points (include/assignment.h, include/nwPerm.h over
// pda_*_host_multi): the frames x window batch of the SLAM loop (slidingWindow.cpp:260-339, system.cpp:268) spread over
// the GPUs of ONE process, no Python anywhere.  Usage:
//     multi_gpu_b200 <problems> <k> <permDim> <dev0,dev1,...>
// A device may be listed more than once (its slices then queue on that device's lock), so the sharding logic can be
// checked on a single-GPU box.  Prints one JSON line: throughputs on the first device alone and on all of them, and
// whether the sharded results are bit-identical to the single-device ones.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "assignment.h"
#include "nwPerm.h"
#include "pda_b200.h"

static uint64_t mix(uint64_t z) {
    z ^= z >> 30; z *= 0xBF58476D1CE4E5B9ULL; z ^= z >> 27; z *= 0x94D049BB133111EBULL; z ^= z >> 31;
    return z;
}
static const uint64_t GAMMA = 0x9E3779B97F4A7C15ULL, SEED = 20260217ULL;
static double u01(uint64_t x) { return (double)(x >> 11) * 0x1p-53; }

// generator G1 of SURVEY.md 8d (the same stream as probabilisticsemslam_b200/synth.py:g1_dense)
static void g1(size_t p, std::vector<double>& C, size_t& nL, size_t& nM) {
    const uint64_t base = (GAMMA * (uint64_t)(p + 1)) ^ SEED;
    nM = 3 + (size_t)(mix(base + GAMMA) % 6);
    nL = 30;
    const size_t n = nL + nM;
    C.assign(n * nM, INFINITY);
    for (size_t c = 0; c < nM; c++) {
        for (size_t r = 0; r < nL; r++) C[r + c * n] = 40.0 * u01(mix(base + (uint64_t)(1 + c * nL + r + 1) * GAMMA));
        C[nL + c + c * n] = 10.0;
    }
}

static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

int main(int argc, char** argv) {
    const size_t nProb = argc > 1 ? (size_t)atoll(argv[1]) : 2000;
    const size_t k = argc > 2 ? (size_t)atoll(argv[2]) : 200;
    const size_t permDim = argc > 3 ? (size_t)atoll(argv[3]) : 24;
    std::vector<int> devices;
    {
        std::string s = argc > 4 ? argv[4] : "0";
        size_t pos = 0;
        while (pos <= s.size()) {
            const size_t q = s.find(',', pos);
            devices.push_back(atoi(s.substr(pos, q == std::string::npos ? std::string::npos : q - pos).c_str()));
            if (q == std::string::npos) break;
            pos = q + 1;
        }
    }
    // ---- the batch in the flat layout of the C ABI, in page-locked buffers every device may use in place ------------
    std::vector<int64_t> costOff(nProb), probOff(nProb);
    std::vector<int32_t> nr(nProb), nc(nProb), nl(nProb);
    std::vector<size_t> nL(nProb), nM(nProb);
    size_t nCost = 0, nPr = 0;
    {
        std::vector<double> tmp;
        for (size_t p = 0; p < nProb; p++) {
            g1(p, tmp, nL[p], nM[p]);
            costOff[p] = (int64_t)nCost; probOff[p] = (int64_t)nPr;
            nr[p] = (int32_t)(nL[p] + nM[p]); nc[p] = (int32_t)nM[p]; nl[p] = (int32_t)nL[p];
            nCost += tmp.size(); nPr += nM[p] * (nL[p] + 1);
        }
    }
    double* costs = (double*)pda_host_alloc((int64_t)nCost * 8);
    double* probs1 = (double*)pda_host_alloc((int64_t)nPr * 8);
    double* probsN = (double*)pda_host_alloc((int64_t)nPr * 8);
    int32_t* found = (int32_t*)pda_host_alloc((int64_t)nProb * 4);
    if (!costs || !probs1 || !probsN || !found) { fprintf(stderr, "pda_host_alloc: %s\n", pda_last_error()); return 2; }
    {
        std::vector<double> tmp;
        size_t a, b;
        for (size_t p = 0; p < nProb; p++) { g1(p, tmp, a, b); memcpy(costs + costOff[p], tmp.data(), tmp.size() * 8); }
    }
    std::vector<int32_t> dev32(devices.begin(), devices.end());
    auto run = [&](double* probs, const int32_t* dev, int nDev) {
        const int rc = pda_murty_batch_host_multi(costs, costOff.data(), nr.data(), nc.data(), (int64_t)nProb, (int32_t)k,
                                                  PDA_CUT_RELATIVE, 42.0, 0, 0, NULL, NULL, NULL, NULL, NULL, found,
                                                  PDA_WEIGHTS_GATED, probs, probOff.data(), nl.data(), dev, nDev);
        if (rc) { fprintf(stderr, "pda_murty_batch_host_multi: %s\n", pda_last_error()); exit(2); }
    };
    run(probsN, dev32.data(), (int)dev32.size());  // warm-up (arenas, streams, module load) on every device
    double tSingle = 1e30, tMulti = 1e30;
    for (int rep = 0; rep < 3; rep++) {
        double t0 = now();
        run(probs1, dev32.data(), 1);
        tSingle = std::min(tSingle, now() - t0);
        t0 = now();
        run(probsN, dev32.data(), (int)dev32.size());
        tMulti = std::min(tMulti, now() - t0);
    }
    bool same = memcmp(probs1, probsN, nPr * 8) == 0;
    double checksum = 0;
    for (size_t p = 0; p < nProb; p++) checksum += probsN[probOff[p]];

    // ---- the std::vector face of the same call (include/assignment.h), on a slice ------------------------------------
    {
        const size_t m = std::min<size_t>(nProb, 1500);
        std::vector<std::vector<double> > vc(m);
        std::vector<size_t> vL(nL.begin(), nL.begin() + (long)m), vM(nM.begin(), nM.begin() + (long)m);
        for (size_t p = 0; p < m; p++) vc[p].assign(costs + costOff[p], costs + costOff[p] + (size_t)nr[p] * nc[p]);
        const std::vector<std::vector<std::vector<double> > > tabs = assignmentProbBatch(vc, vL, vM, k, devices);
        for (size_t p = 0; same && p < m; p++)
            for (size_t c = 0; same && c < vM[p]; c++)
                same = memcmp(tabs[p][c].data(), probsN + probOff[p] + c * (vL[p] + 1), sizeof(double) * (vL[p] + 1)) == 0;
    }
    const std::vector<int> one(1, devices[0]);

    // ---- permanent weights of small gated problems (config 4), sharded the same way -----------------------------
    const size_t nPerm = std::min<size_t>(nProb, 256);
    std::vector<std::vector<double> > pc(nPerm);
    std::vector<size_t> pL(nPerm), pM(nPerm);
    for (size_t p = 0; p < nPerm; p++) {  // 3..5 detections x 9 landmarks: 11..13 rows
        const uint64_t base = (GAMMA * (uint64_t)(p + 777)) ^ SEED;
        pM[p] = 3 + (size_t)(mix(base + GAMMA) % 3); pL[p] = 9;
        const size_t n = pL[p] + pM[p];
        pc[p].assign(n * pM[p], INFINITY);
        for (size_t c = 0; c < pM[p]; c++) {
            for (size_t r = 0; r < pL[p]; r++) pc[p][r + c * n] = 30.0 * u01(mix(base + (uint64_t)(2 + c * pL[p] + r) * GAMMA));
            pc[p][pL[p] + c + c * n] = 10.0;
        }
    }
    const std::vector<std::vector<std::vector<double> > > pp1 = permanentProbBatch(pc, pL, pM, 1, one);
    const std::vector<std::vector<std::vector<double> > > ppN = permanentProbBatch(pc, pL, pM, 1, devices);
    bool samePerm = true;
    for (size_t p = 0; samePerm && p < nPerm; p++)
        for (size_t m = 0; samePerm && m < pM[p]; m++)
            samePerm = memcmp(pp1[p][m].data(), ppN[p][m].data(), sizeof(double) * (pL[p] + 1)) == 0;

    // ---- ONE permanent, Gray range split over the devices ---------------------------------------------------------
    std::vector<double> A(permDim * permDim);
    for (size_t i = 0; i < A.size(); i++) A[i] = u01(mix(((GAMMA * 4243ULL) ^ SEED) + (uint64_t)(i + 1) * GAMMA));
    const double whole = permanentExactRaw(A.data(), permDim, permDim);
    permanentExactShardedRaw(A.data(), permDim, devices.data(), devices.size());
    double t0 = now();
    const double sharded = permanentExactShardedRaw(A.data(), permDim, devices.data(), devices.size());
    const double tPerm = now() - t0;

    printf("{\"problems\": %zu, \"k\": %zu, \"devices\": %zu, \"single_problems_per_s\": %.1f, \"multi_problems_per_s\": %.1f, "
           "\"speedup\": %.3f, \"bit_identical\": %s, \"checksum\": %.17g, \"permanent_prob_bit_identical\": %s, "
           "\"perm_dim\": %zu, \"perm_whole\": %.17g, \"perm_sharded\": %.17g, \"perm_rel_diff\": %.3g, \"perm_sharded_ms\": %.4f}\n",
           nProb, k, devices.size(), nProb / tSingle, nProb / tMulti, tSingle / tMulti, same ? "true" : "false", checksum,
           samePerm ? "true" : "false", permDim, whole, sharded, std::fabs(sharded - whole) / std::fabs(whole), tPerm * 1e3);
    pda_host_free(costs); pda_host_free(probs1); pda_host_free(probsN); pda_host_free(found);
    return (same && samePerm) ? 0 : 1;
}
