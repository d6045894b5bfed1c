#include "ORBmatcher_accel.h"

#include <cmath>
#include <cstring>
#include <stdexcept>
#include <string>

#include "rumi_orb.h"

namespace ORB_SLAM3 {

static std::vector<uint8_t> rows32(const cv::Mat& m) {
    std::vector<uint8_t> out(32 * (size_t)m.rows);
    for (int i = 0; i < m.rows; ++i) std::memcpy(out.data() + 32 * (size_t)i, m.ptr(i), 32);
    return out;
}

ORBmatcherAccel::ORBmatcherAccel(float nnratio, int device) : ctx(nullptr), mfNNratio(nnratio) {
    if (rumi_match_create(&ctx, device) != RUMI_OK)
        throw std::runtime_error(std::string("ORBmatcherAccel: ") + rumi_last_error());
}
ORBmatcherAccel::~ORBmatcherAccel() { rumi_match_destroy(ctx); }

int ORBmatcherAccel::DescriptorDistance(const cv::Mat& a, const cv::Mat& b) {
    return rumi_descriptor_distance(a.ptr(0), b.ptr(0));
}

void ORBmatcherAccel::Top2(const cv::Mat& Q, const cv::Mat& T, std::vector<int>& idx1, std::vector<uint16_t>& d1,
                           std::vector<uint16_t>& d2) {
    const std::vector<uint8_t> q = rows32(Q), t = rows32(T);
    idx1.assign(Q.rows, -1); d1.assign(Q.rows, 256); d2.assign(Q.rows, 256);
    if (Q.rows == 0) return;
    if (rumi_hamming_top2(ctx, q.data(), Q.rows, t.data(), T.rows, idx1.data(), d1.data(), d2.data()) != RUMI_OK)
        throw std::runtime_error(std::string("ORBmatcherAccel: ") + rumi_last_error());
}

int ORBmatcherAccel::MatchRatio(const cv::Mat& Q, const cv::Mat& T, std::vector<int>& matches12, int th) {
    std::vector<int> idx;
    std::vector<uint16_t> d1, d2;
    Top2(Q, T, idx, d1, d2);
    matches12.assign(Q.rows, -1);
    int n = 0;
    for (int q = 0; q < Q.rows; ++q) {
        const int bestDist1 = d1[q], bestDist2 = d2[q];
        if (idx[q] >= 0 && bestDist1 <= th &&
            static_cast<float>(bestDist1) < mfNNratio * static_cast<float>(bestDist2)) {   // ORBmatcher.cc:290-291
            matches12[q] = idx[q];
            ++n;
        }
    }
    return n;
}

int ORBmatcherAccel::MatchKeyFramePairs(const std::vector<cv::Mat>& descA, const std::vector<cv::Mat>& descB,
                                        std::vector<std::vector<int>>& matches12, int th) {
    if (descA.size() != descB.size()) throw std::runtime_error("ORBmatcherAccel: pair lists differ in length");
    std::vector<uint8_t> Q, T;
    std::vector<int32_t> segs;
    for (size_t p = 0; p < descA.size(); ++p) {
        const std::vector<uint8_t> a = rows32(descA[p]), b = rows32(descB[p]);
        segs.push_back((int32_t)(Q.size() / 32)); segs.push_back(descA[p].rows);
        segs.push_back((int32_t)(T.size() / 32)); segs.push_back(descB[p].rows);
        Q.insert(Q.end(), a.begin(), a.end());
        T.insert(T.end(), b.begin(), b.end());
    }
    const int nq = (int)(Q.size() / 32), nt = (int)(T.size() / 32);
    std::vector<int32_t> idx(nq);
    std::vector<uint16_t> d1(nq), d2(nq);
    if (rumi_hamming_top2_pairs(ctx, Q.data(), nq, T.data(), nt, segs.data(), (int)descA.size(), idx.data(), d1.data(),
                                d2.data()) != RUMI_OK)
        throw std::runtime_error(std::string("ORBmatcherAccel: ") + rumi_last_error());
    matches12.assign(descA.size(), std::vector<int>());
    int n = 0;
    for (size_t p = 0; p < descA.size(); ++p) {
        matches12[p].assign(descA[p].rows, -1);
        for (int q = 0; q < descA[p].rows; ++q) {
            const int g = segs[4 * p] + q, bestDist1 = d1[g], bestDist2 = d2[g];
            if (idx[g] >= 0 && bestDist1 <= th &&
                static_cast<float>(bestDist1) < mfNNratio * static_cast<float>(bestDist2)) {   // ORBmatcher.cc:290-291
                matches12[p][q] = idx[g];
                ++n;
            }
        }
    }
    return n;
}

void ORBmatcherAccel::StereoBest1(const std::vector<cv::KeyPoint>& keysL, const cv::Mat& descL,
                                  const std::vector<cv::KeyPoint>& keysR, const cv::Mat& descR,
                                  const std::vector<float>& scaleFactors, int nRows, float minD, float maxD,
                                  std::vector<int>& bestIdxR, std::vector<uint16_t>& bestDist) {
    static_assert(sizeof(cv::KeyPoint) == sizeof(rumi_kp), "cv::KeyPoint layout");
    const std::vector<uint8_t> l = rows32(descL), r = rows32(descR);
    bestIdxR.assign(keysL.size(), -1); bestDist.assign(keysL.size(), TH_HIGH);
    if (keysL.empty()) return;
    if (rumi_stereo_best1(ctx, reinterpret_cast<const rumi_kp*>(keysL.data()), l.data(), (int)keysL.size(),
                          reinterpret_cast<const rumi_kp*>(keysR.data()), r.data(), (int)keysR.size(),
                          scaleFactors.data(), (int)scaleFactors.size(), nRows, minD, maxD, bestIdxR.data(),
                          bestDist.data()) != RUMI_OK)
        throw std::runtime_error(std::string("ORBmatcherAccel: ") + rumi_last_error());
}

int ORBmatcherAccel::ComputeStereoMatches(rumi_orb* extractorLeft, rumi_orb* extractorRight,
                                          const std::vector<cv::KeyPoint>& keysL, const cv::Mat& descL,
                                          const std::vector<cv::KeyPoint>& keysR, const cv::Mat& descR, float mbf,
                                          float mb, std::vector<float>& mvuRight, std::vector<float>& mvDepth) {
    const std::vector<uint8_t> l = rows32(descL), r = rows32(descR);
    mvuRight.assign(keysL.size(), -1.0f);
    mvDepth.assign(keysL.size(), -1.0f);
    int n = 0;
    if (keysL.empty()) return 0;
    if (rumi_stereo_match(ctx, extractorLeft, extractorRight, reinterpret_cast<const rumi_kp*>(keysL.data()), l.data(),
                          (int)keysL.size(), reinterpret_cast<const rumi_kp*>(keysR.data()), r.data(),
                          (int)keysR.size(), mbf, mb, mvuRight.data(), mvDepth.data(), &n) != RUMI_OK)
        throw std::runtime_error(std::string("ORBmatcherAccel: ") + rumi_last_error());
    return n;
}

// R/lib_src/ORBmatcher.cc:1795-1828
static void ComputeThreeMaxima(std::vector<int>* histo, const int L, int& ind1, int& ind2, int& ind3) {
    int max1 = 0, max2 = 0, max3 = 0;
    for (int i = 0; i < L; i++) {
        const int s = (int)histo[i].size();
        if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
        else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
        else if (s > max3) { max3 = s; ind3 = i; }
    }
    if (max2 < 0.1f * (float)max1) { ind2 = -1; ind3 = -1; }
    else if (max3 < 0.1f * (float)max1) { ind3 = -1; }
}

long long ORBmatcherAccel::NodeBlocks(const cv::Mat& descA,
                                      const std::vector<std::pair<unsigned, std::vector<unsigned> > >& fvA,
                                      const cv::Mat& descB,
                                      const std::vector<std::pair<unsigned, std::vector<unsigned> > >& fvB,
                                      std::vector<int32_t>& aIdx, std::vector<int32_t>& bIdx, std::vector<int32_t>& segs,
                                      std::vector<uint16_t>& dist) {
    // common nodes in ascending order == the nodes the reference's lower_bound walk visits (:218-221, :345-353)
    long long ndist = 0;
    size_t a = 0, b = 0;
    while (a < fvA.size() && b < fvB.size()) {
        if (fvA[a].first == fvB[b].first) {
            const std::vector<unsigned>&ia = fvA[a].second, &ib = fvB[b].second;
            segs.push_back((int32_t)aIdx.size()); segs.push_back((int32_t)ia.size());
            segs.push_back((int32_t)bIdx.size()); segs.push_back((int32_t)ib.size());
            segs.push_back((int32_t)ndist);
            aIdx.insert(aIdx.end(), ia.begin(), ia.end());
            bIdx.insert(bIdx.end(), ib.begin(), ib.end());
            ndist += (long long)ia.size() * (long long)ib.size();
            ++a; ++b;
        } else if (fvA[a].first < fvB[b].first) ++a;
        else ++b;
    }
    if (segs.empty() || ndist == 0) return 0;
    const std::vector<uint8_t> da = rows32(descA), db = rows32(descB);
    dist.resize((size_t)ndist);
    if (rumi_bow_node_distances(ctx, da.data(), descA.rows, db.data(), descB.rows, aIdx.data(), (int)aIdx.size(),
                                bIdx.data(), (int)bIdx.size(), segs.data(), (int)(segs.size() / 5), dist.data(),
                                ndist) != RUMI_OK)
        throw std::runtime_error(std::string("ORBmatcherAccel: ") + rumi_last_error());
    return ndist;
}

int ORBmatcherAccel::SearchByBoWKF(const cv::Mat& desc1, const std::vector<float>& angle1,
                                   const std::vector<uint8_t>& valid1,
                                   const std::vector<std::pair<unsigned, std::vector<unsigned> > >& featVec1,
                                   const cv::Mat& desc2, const std::vector<float>& angle2,
                                   const std::vector<uint8_t>& valid2,
                                   const std::vector<std::pair<unsigned, std::vector<unsigned> > >& featVec2,
                                   bool checkOrientation, std::vector<int>& match12) {
    match12.assign(desc1.rows, -1);
    std::vector<bool> vbMatched2(desc2.rows, false);
    std::vector<int32_t> aIdx, bIdx, segs;
    std::vector<uint16_t> dist;
    if (NodeBlocks(desc1, featVec1, desc2, featVec2, aIdx, bIdx, segs, dist) == 0) return 0;
    int nmatches = 0;
    std::vector<int> rotHist[HISTO_LENGTH];
    for (int i = 0; i < HISTO_LENGTH; i++) rotHist[i].reserve(500);
    const float factor = 1.0f / HISTO_LENGTH;
    for (size_t s = 0; s < segs.size(); s += 5) {
        const int a0 = segs[s], ac = segs[s + 1], b0 = segs[s + 2], bc = segs[s + 3];
        const uint16_t* block = dist.data() + segs[s + 4];
        for (int i1 = 0; i1 < ac; i1++) {
            const int idx1 = aIdx[a0 + i1];
            if (!valid1[idx1]) continue;                                          // :713-717
            int bestDist1 = 256, bestIdx2 = -1, bestDist2 = 256;
            for (int i2 = 0; i2 < bc; i2++) {
                const int idx2 = bIdx[b0 + i2];
                if (vbMatched2[idx2] || !valid2[idx2]) continue;                  // :735-739
                const int d = block[(size_t)i1 * bc + i2];
                if (d < bestDist1) { bestDist2 = bestDist1; bestDist1 = d; bestIdx2 = idx2; }
                else if (d < bestDist2) { bestDist2 = d; }
            }
            if (bestDist1 < TH_LOW) {                                             // strict '<' in this overload (:756)
                if (static_cast<float>(bestDist1) < mfNNratio * static_cast<float>(bestDist2)) {
                    match12[idx1] = bestIdx2;
                    vbMatched2[bestIdx2] = true;
                    if (checkOrientation) {
                        float rot = angle1[idx1] - angle2[bestIdx2];
                        if (rot < 0.0) rot += 360.0f;
                        int bin = (int)std::round(rot * factor);
                        if (bin == HISTO_LENGTH) bin = 0;
                        rotHist[bin].push_back(idx1);
                    }
                    nmatches++;
                }
            }
        }
    }
    if (checkOrientation) {
        int ind1 = -1, ind2 = -1, ind3 = -1;
        ComputeThreeMaxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
        for (int i = 0; i < HISTO_LENGTH; i++) {
            if (i == ind1 || i == ind2 || i == ind3) continue;
            for (size_t j = 0, jend = rotHist[i].size(); j < jend; j++) { match12[rotHist[i][j]] = -1; nmatches--; }
        }
    }
    return nmatches;
}

int ORBmatcherAccel::SearchByBoW(const cv::Mat& descKF, const std::vector<float>& angleKF,
                                 const std::vector<uint8_t>& kfValid,
                                 const std::vector<std::pair<unsigned, std::vector<unsigned> > >& featVecKF,
                                 const cv::Mat& descF, const std::vector<float>& angleF,
                                 const std::vector<std::pair<unsigned, std::vector<unsigned> > >& featVecF,
                                 bool checkOrientation, std::vector<int>& matchF) {
    matchF.assign(descF.rows, -1);
    std::vector<int32_t> aIdx, bIdx, segs;
    std::vector<uint16_t> dist;
    if (NodeBlocks(descKF, featVecKF, descF, featVecF, aIdx, bIdx, segs, dist) == 0) return 0;

    int nmatches = 0;
    std::vector<int> rotHist[HISTO_LENGTH];
    for (int i = 0; i < HISTO_LENGTH; i++) rotHist[i].reserve(500);
    const float factor = 1.0f / HISTO_LENGTH;
    for (size_t s = 0; s < segs.size(); s += 5) {
        const int a0 = segs[s], ac = segs[s + 1], b0 = segs[s + 2], bc = segs[s + 3];
        const uint16_t* block = dist.data() + segs[s + 4];
        for (int iKF = 0; iKF < ac; iKF++) {
            const int realIdxKF = aIdx[a0 + iKF];
            if (!kfValid[realIdxKF]) continue;                                   // :229-233
            int bestDist1 = 256, bestIdxF = -1, bestDist2 = 256;
            for (int iF = 0; iF < bc; iF++) {
                const int realIdxF = bIdx[b0 + iF];
                if (matchF[realIdxF] >= 0) continue;                              // :249
                const int d = block[(size_t)iKF * bc + iF];
                if (d < bestDist1) { bestDist2 = bestDist1; bestDist1 = d; bestIdxF = realIdxF; }
                else if (d < bestDist2) { bestDist2 = d; }
            }
            if (bestDist1 <= TH_LOW) {
                if (static_cast<float>(bestDist1) < mfNNratio * static_cast<float>(bestDist2)) {
                    matchF[bestIdxF] = realIdxKF;
                    if (checkOrientation) {
                        float rot = angleKF[realIdxKF] - angleF[bestIdxF];
                        if (rot < 0.0) rot += 360.0f;
                        int bin = (int)std::round(rot * factor);
                        if (bin == HISTO_LENGTH) bin = 0;
                        rotHist[bin].push_back(bestIdxF);
                    }
                    nmatches++;
                }
            }
        }
    }
    if (checkOrientation) {
        int ind1 = -1, ind2 = -1, ind3 = -1;
        ComputeThreeMaxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
        for (int i = 0; i < HISTO_LENGTH; i++) {
            if (i == ind1 || i == ind2 || i == ind3) continue;
            for (size_t j = 0, jend = rotHist[i].size(); j < jend; j++) { matchF[rotHist[i][j]] = -1; nmatches--; }
        }
    }
    return nmatches;
}

}  // namespace ORB_SLAM3
