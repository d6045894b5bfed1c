#include "ORBmatcher_accel.h"

#include <algorithm>
#include <climits>
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

int ORBmatcherAccel::SearchForTriangulation(const cv::Mat& desc1, const std::vector<float>& angle1,
                                            const std::vector<uint8_t>& hasMapPoint1, const std::vector<uint8_t>& stereo1,
                                            const std::vector<std::pair<unsigned, std::vector<unsigned> > >& featVec1,
                                            const cv::Mat& desc2, const std::vector<float>& angle2,
                                            const std::vector<uint8_t>& hasMapPoint2, const std::vector<uint8_t>& stereo2,
                                            const std::vector<cv::KeyPoint>& keys2,
                                            const std::vector<std::pair<unsigned, std::vector<unsigned> > >& featVec2,
                                            const std::vector<float>& scaleFactors2, cv::Point2f ep,
                                            const std::function<bool(size_t, size_t)>& epipolarOk, bool bOnlyStereo, bool bCoarse,
                                            bool checkOrientation, std::vector<std::pair<size_t, size_t> >& vMatchedPairs) {
    vMatchedPairs.clear();
    std::vector<int> vMatches12(desc1.rows, -1);
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
            const size_t idx1 = aIdx[a0 + i1];
            if (hasMapPoint1[idx1]) continue;                                     // :862-867
            const bool bStereo1 = stereo1[idx1] != 0;
            if (bOnlyStereo && !bStereo1) continue;
            int bestDist = TH_LOW, bestIdx2 = -1;
            for (int i2 = 0; i2 < bc; i2++) {
                const size_t idx2 = bIdx[b0 + i2];
                if (hasMapPoint2[idx2]) continue;                                 // :889-893 (vbMatched2 is never set there)
                const bool bStereo2 = stereo2[idx2] != 0;
                if (bOnlyStereo && !bStereo2) continue;
                const int d = block[(size_t)i1 * bc + i2];
                if (d > TH_LOW || d > bestDist) continue;
                const cv::KeyPoint& kp2 = keys2[idx2];
                if (!bStereo1 && !bStereo2) {                                     // too close to the epipole (:910-916)
                    const float distex = ep.x - kp2.pt.x, distey = ep.y - kp2.pt.y;
                    if (distex * distex + distey * distey < 100 * scaleFactors2[kp2.octave]) continue;
                }
                if (bCoarse || epipolarOk(idx1, idx2)) { bestIdx2 = (int)idx2; bestDist = d; }
            }
            if (bestIdx2 >= 0) {
                vMatches12[idx1] = bestIdx2;
                nmatches++;
                if (checkOrientation) {
                    float rot = angle1[idx1] - angle2[bestIdx2];
                    if (rot < 0.0) rot += 360.0f;
                    int bin = (int)std::round(rot * factor);
                    if (bin == HISTO_LENGTH) bin = 0;
                    rotHist[bin].push_back((int)idx1);
                }
            }
        }
    }
    if (checkOrientation) {
        int ind1 = -1, ind2 = -1, ind3 = -1;
        ComputeThreeMaxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
        for (int i = 0; i < HISTO_LENGTH; i++) {
            if (i == ind1 || i == ind2 || i == ind3) continue;
            for (size_t j = 0, jend = rotHist[i].size(); j < jend; j++) { vMatches12[rotHist[i][j]] = -1; nmatches--; }
        }
    }
    vMatchedPairs.reserve(nmatches);
    for (size_t i = 0, iend = vMatches12.size(); i < iend; i++)
        if (vMatches12[i] >= 0) vMatchedPairs.push_back(std::make_pair(i, (size_t)vMatches12[i]));
    return nmatches;
}

int ORBmatcherAccel::SearchByBoW(const cv::Mat& descKF, const std::vector<float>& angleKF,
                                 const std::vector<uint8_t>& kfValid,
                                 const std::vector<std::pair<unsigned, std::vector<unsigned> > >& featVecKF,
                                 const cv::Mat& descF, const std::vector<float>& angleF,
                                 const std::vector<std::pair<unsigned, std::vector<unsigned> > >& featVecF,
                                 bool checkOrientation, std::vector<int>& matchF, int Nleft) {
    matchF.assign(descF.rows, -1);
    std::vector<int32_t> aIdx, bIdx, segs;
    std::vector<uint16_t> dist;
    if (NodeBlocks(descKF, featVecKF, descF, featVecF, aIdx, bIdx, segs, dist) == 0) return 0;

    int nmatches = 0;
    std::vector<int> rotHist[HISTO_LENGTH];
    for (int i = 0; i < HISTO_LENGTH; i++) rotHist[i].reserve(500);
    const float factor = 1.0f / HISTO_LENGTH;
    auto accept = [&](int realIdxKF, int idxF) {
        matchF[idxF] = realIdxKF;
        if (checkOrientation) {
            float rot = angleKF[realIdxKF] - angleF[idxF];
            if (rot < 0.0) rot += 360.0f;
            int bin = (int)std::round(rot * factor);
            if (bin == HISTO_LENGTH) bin = 0;
            rotHist[bin].push_back(idxF);
        }
        nmatches++;
    };
    for (size_t s = 0; s < segs.size(); s += 5) {
        const int a0 = segs[s], ac = segs[s + 1], b0 = segs[s + 2], bc = segs[s + 3];
        const uint16_t* block = dist.data() + segs[s + 4];
        for (int iKF = 0; iKF < ac; iKF++) {
            const int realIdxKF = aIdx[a0 + iKF];
            if (!kfValid[realIdxKF]) continue;                                   // :229-233
            int bestDist1 = 256, bestIdxF = -1, bestDist2 = 256;
            int bestDist1R = 256, bestIdxFR = -1, bestDist2R = 256;              // right fisheye camera (F.Nleft != -1, :258-283)
            for (int iF = 0; iF < bc; iF++) {
                const int realIdxF = bIdx[b0 + iF];
                if (matchF[realIdxF] >= 0) continue;                              // :249
                const int d = block[(size_t)iKF * bc + iF];
                if (Nleft == -1 || realIdxF < Nleft) {
                    if (d < bestDist1) { bestDist2 = bestDist1; bestDist1 = d; bestIdxF = realIdxF; }
                    else if (d < bestDist2) { bestDist2 = d; }
                } else {
                    if (d < bestDist1R) { bestDist2R = bestDist1R; bestDist1R = d; bestIdxFR = realIdxF; }
                    else if (d < bestDist2R) { bestDist2R = d; }
                }
            }
            if (bestDist1 <= TH_LOW) {
                if (static_cast<float>(bestDist1) < mfNNratio * static_cast<float>(bestDist2)) accept(realIdxKF, bestIdxF);
                if (bestDist1R <= TH_LOW) accept(realIdxKF, bestIdxFR);           // :314-316 ("|| true": no ratio test on the right)
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

void ORBmatcherAccel::StereoFishEyeMatches(const cv::Mat& descLeft, int monoLeft, const cv::Mat& descRight, int monoRight,
                                           std::vector<std::pair<int, int> >& pairs) {
    pairs.clear();
    const int nq = descLeft.rows - monoLeft, nt = descRight.rows - monoRight;
    if (nq <= 0 || nt < 2) return;                     // knnMatch returns fewer than 2 neighbours: "size() >= 2" fails (:1146)
    cv::Mat Q(nq, 32, CV_8U), T(nt, 32, CV_8U);
    for (int i = 0; i < nq; ++i) std::memcpy(Q.ptr(i), descLeft.ptr(monoLeft + i), 32);
    for (int i = 0; i < nt; ++i) std::memcpy(T.ptr(i), descRight.ptr(monoRight + i), 32);
    std::vector<int> idx;
    std::vector<uint16_t> d1, d2;
    Top2(Q, T, idx, d1, d2);
    for (int q = 0; q < nq; ++q)
        if (idx[q] >= 0 && (float)d1[q] < (float)d2[q] * 0.7)                    // DMatch::distance is float, 0.7 a double
            pairs.push_back(std::make_pair(q + monoLeft, idx[q] + monoRight));
}

// ---------------------------------------------------------------------------------------------------------------
// key-point grid
FrameGridAccel::FrameGridAccel(const std::vector<cv::KeyPoint>& keysUn, float minX, float minY, float maxX, float maxY)
    : mnMinX(minX), mnMinY(minY), mnMaxX(maxX), mnMaxY(maxY), keys(keysUn) {
    mfGridElementWidthInv = static_cast<float>(kCols) / static_cast<float>(maxX - minX);        // Frame.cc:98-99
    mfGridElementHeightInv = static_cast<float>(kRows) / static_cast<float>(maxY - minY);
    for (size_t i = 0; i < keys.size(); ++i) {                                                  // AssignFeaturesToGrid
        const int posX = (int)std::round((keys[i].pt.x - mnMinX) * mfGridElementWidthInv);      // PosInGrid
        const int posY = (int)std::round((keys[i].pt.y - mnMinY) * mfGridElementHeightInv);
        if (posX < 0 || posX >= kCols || posY < 0 || posY >= kRows) continue;
        mGrid[posX][posY].push_back(i);
    }
}

std::vector<size_t> FrameGridAccel::GetFeaturesInArea(const float& x, const float& y, const float& r, const int minLevel,
                                                      const int maxLevel) const {
    std::vector<size_t> vIndices;
    const int nMinCellX = std::max(0, (int)std::floor((x - mnMinX - r) * mfGridElementWidthInv));
    if (nMinCellX >= kCols) return vIndices;
    const int nMaxCellX = std::min(kCols - 1, (int)std::ceil((x - mnMinX + r) * mfGridElementWidthInv));
    if (nMaxCellX < 0) return vIndices;
    const int nMinCellY = std::max(0, (int)std::floor((y - mnMinY - r) * mfGridElementHeightInv));
    if (nMinCellY >= kRows) return vIndices;
    const int nMaxCellY = std::min(kRows - 1, (int)std::ceil((y - mnMinY + r) * mfGridElementHeightInv));
    if (nMaxCellY < 0) return vIndices;
    const bool bCheckLevels = (minLevel > 0) || (maxLevel >= 0);
    for (int ix = nMinCellX; ix <= nMaxCellX; ix++)
        for (int iy = nMinCellY; iy <= nMaxCellY; iy++)
            for (size_t j : mGrid[ix][iy]) {
                const cv::KeyPoint& kpUn = keys[j];
                if (bCheckLevels) {
                    if (kpUn.octave < minLevel) continue;
                    if (maxLevel >= 0 && kpUn.octave > maxLevel) continue;
                }
                if (std::fabs(kpUn.pt.x - x) < r && std::fabs(kpUn.pt.y - y) < r) vIndices.push_back(j);
            }
    return vIndices;
}

namespace {
// ORBmatcher::ComputeThreeMaxima (R/lib_src/ORBmatcher.cc:1795-1826)
void three_maxima(const std::vector<int>* histo, int L, int& ind1, int& ind2, int& ind3) {
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
}  // namespace

int ORBmatcherAccel::SearchForInitialization(const std::vector<cv::KeyPoint>& keys1, const cv::Mat& desc1,
                                             const std::vector<cv::KeyPoint>& keys2, const cv::Mat& desc2,
                                             const FrameGridAccel& grid2, std::vector<cv::Point2f>& vbPrevMatched,
                                             std::vector<int>& vnMatches12, int windowSize, bool checkOrientation) {
    int nmatches = 0;
    vnMatches12 = std::vector<int>(keys1.size(), -1);
    std::vector<int> rotHist[HISTO_LENGTH];
    const float factor = 1.0f / HISTO_LENGTH;
    std::vector<int> vMatchedDistance(keys2.size(), INT_MAX), vnMatches21(keys2.size(), -1);
    // candidate lists of every level-0 key point of F1 (:596-602), all distances in one launch
    std::vector<int> query;
    std::vector<int32_t> off(1, 0), idx;
    std::vector<uint8_t> Q;
    for (size_t i1 = 0; i1 < keys1.size(); i1++) {
        const int level1 = keys1[i1].octave;
        if (level1 > 0) continue;
        const std::vector<size_t> v = grid2.GetFeaturesInArea(vbPrevMatched[i1].x, vbPrevMatched[i1].y, windowSize, level1, level1);
        query.push_back((int)i1);
        idx.insert(idx.end(), v.begin(), v.end());
        off.push_back((int32_t)idx.size());
        Q.insert(Q.end(), desc1.ptr((int)i1), desc1.ptr((int)i1) + 32);
    }
    std::vector<uint16_t> dist(idx.size() + 1);
    const std::vector<uint8_t> T = rows32(desc2);
    if (!query.empty() && rumi_hamming_candidates(ctx, Q.data(), (int)query.size(), T.data(), desc2.rows, off.data(), idx.data(),
                                                  dist.data(), nullptr, nullptr, nullptr, nullptr) != RUMI_OK)
        throw std::runtime_error(std::string("ORBmatcherAccel: ") + rumi_last_error());
    for (size_t qi = 0; qi < query.size(); ++qi) {
        const int i1 = query[qi];
        int bestDist = INT_MAX, bestDist2 = INT_MAX, bestIdx2 = -1;
        for (int p = off[qi]; p < off[qi + 1]; ++p) {
            const int i2 = idx[p], d = dist[p];
            if (vMatchedDistance[i2] <= d) continue;                                            // :617
            if (d < bestDist) { bestDist2 = bestDist; bestDist = d; bestIdx2 = i2; }
            else if (d < bestDist2) bestDist2 = d;
        }
        if (bestDist <= TH_LOW) {
            if (bestDist < (float)bestDist2 * mfNNratio) {
                if (vnMatches21[bestIdx2] >= 0) { vnMatches12[vnMatches21[bestIdx2]] = -1; nmatches--; }
                vnMatches12[i1] = bestIdx2;
                vnMatches21[bestIdx2] = i1;
                vMatchedDistance[bestIdx2] = bestDist;
                nmatches++;
                if (checkOrientation) {
                    float rot = keys1[i1].angle - keys2[bestIdx2].angle;
                    if (rot < 0.0) rot += 360.0f;
                    int bin = (int)std::round(rot * factor);
                    if (bin == HISTO_LENGTH) bin = 0;
                    rotHist[bin].push_back(i1);
                }
            }
        }
    }
    if (checkOrientation) {
        int ind1 = -1, ind2 = -1, ind3 = -1;
        three_maxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
        for (int i = 0; i < HISTO_LENGTH; i++) {
            if (i == ind1 || i == ind2 || i == ind3) continue;
            for (int idx1 : rotHist[i])
                if (vnMatches12[idx1] >= 0) { vnMatches12[idx1] = -1; nmatches--; }
        }
    }
    for (size_t i1 = 0; i1 < vnMatches12.size(); i1++)
        if (vnMatches12[i1] >= 0) vbPrevMatched[i1] = keys2[vnMatches12[i1]].pt;
    return nmatches;
}

int ORBmatcherAccel::SearchByProjection(const std::vector<cv::KeyPoint>& keysF, const cv::Mat& descF,
                                        const FrameGridAccel& gridF, const std::vector<float>& scaleFactors,
                                        const std::vector<cv::Point2f>& proj, const std::vector<int>& level,
                                        const std::vector<float>& viewCos, const cv::Mat& descMP,
                                        const std::vector<uint8_t>& hasObservations, float th, std::vector<int>& frameMatch) {
    const int nMP = (int)proj.size();
    frameMatch.assign(keysF.size(), -1);
    const bool bFactor = th != 1.0;
    std::vector<int32_t> off(1, 0), idx;
    for (int iMP = 0; iMP < nMP; ++iMP) {
        const int nPredictedLevel = level[iMP];
        float r = viewCos[iMP] > 0.998 ? 2.5f : 4.0f;                                           // RadiusByViewingCos
        if (bFactor) r *= th;
        const std::vector<size_t> v = gridF.GetFeaturesInArea(proj[iMP].x, proj[iMP].y, r * scaleFactors[nPredictedLevel],
                                                              nPredictedLevel - 1, nPredictedLevel);
        idx.insert(idx.end(), v.begin(), v.end());
        off.push_back((int32_t)idx.size());
    }
    std::vector<uint16_t> dist(idx.size() + 1);
    const std::vector<uint8_t> Q = rows32(descMP), T = rows32(descF);
    if (nMP > 0 && rumi_hamming_candidates(ctx, Q.data(), nMP, T.data(), descF.rows, off.data(), idx.data(), dist.data(), nullptr,
                                           nullptr, nullptr, nullptr) != RUMI_OK)
        throw std::runtime_error(std::string("ORBmatcherAccel: ") + rumi_last_error());
    int nmatches = 0;
    for (int iMP = 0; iMP < nMP; ++iMP) {
        int bestDist = 256, bestLevel = -1, bestDist2 = 256, bestLevel2 = -1, bestIdx = -1;
        for (int p = off[iMP]; p < off[iMP + 1]; ++p) {
            const int j = idx[p], d = dist[p];
            if (frameMatch[j] >= 0 && hasObservations[frameMatch[j]]) continue;                 // :86-88
            if (d < bestDist) { bestDist2 = bestDist; bestDist = d; bestLevel2 = bestLevel; bestLevel = keysF[j].octave; bestIdx = j; }
            else if (d < bestDist2) { bestLevel2 = keysF[j].octave; bestDist2 = d; }
        }
        if (bestDist <= TH_HIGH && bestIdx >= 0) {
            if (bestLevel == bestLevel2 && bestDist > mfNNratio * bestDist2) continue;
            if (bestLevel != bestLevel2 || bestDist <= mfNNratio * bestDist2) { frameMatch[bestIdx] = iMP; nmatches++; }
        }
    }
    return nmatches;
}

int ORBmatcherAccel::SearchByProjection(const std::vector<cv::KeyPoint>& keysF, const cv::Mat& descF,
                                        const FrameGridAccel& gridF, const std::vector<float>& scaleFactors,
                                        const std::vector<cv::Point2f>& proj, const std::vector<int>& level,
                                        const std::vector<float>& viewCos, const cv::Mat& descMP,
                                        const std::vector<uint8_t>& hasObservations, float th, const LocalPointsExtras& ex,
                                        std::vector<int>& frameMatch) {
    const int nMP = (int)proj.size(), nL = (int)keysF.size();
    const bool fisheye = ex.keysRight && ex.gridRight && !ex.keysRight->empty();
    const int nR = fisheye ? (int)ex.keysRight->size() : 0, Nleft = fisheye ? nL : -1;
    const bool bFactor = th != 1.0;
    auto inV = [&](int i) { return ex.inView ? (*ex.inView)[i] != 0 : true; };
    auto inVR = [&](int i) { return fisheye && ex.inViewR && (*ex.inViewR)[i] != 0; };
    // candidate lists of every window of both cameras: queries [0, nMP) = left halves, [nMP, 2 nMP) = right halves
    std::vector<int32_t> off(1, 0), idx;
    std::vector<float> radius(nMP, 0.0f);
    for (int i = 0; i < nMP; ++i) {
        if (inV(i)) {
            float r = viewCos[i] > 0.998 ? 2.5f : 4.0f;                                          // RadiusByViewingCos
            if (bFactor) r *= th;
            radius[i] = r * scaleFactors[level[i]];
            const std::vector<size_t> v = gridF.GetFeaturesInArea(proj[i].x, proj[i].y, radius[i], level[i] - 1, level[i]);
            idx.insert(idx.end(), v.begin(), v.end());
        }
        off.push_back((int32_t)idx.size());
    }
    for (int i = 0; i < nMP; ++i) {
        if (inVR(i) && (*ex.levelR)[i] != -1) {
            const int lv = (*ex.levelR)[i];
            const float r = (*ex.viewCosR)[i] > 0.998 ? 2.5f : 4.0f;                             // no th factor here (:129)
            const std::vector<size_t> v = ex.gridRight->GetFeaturesInArea((*ex.projR)[i].x, (*ex.projR)[i].y, r * scaleFactors[lv],
                                                                          lv - 1, lv);
            for (size_t j : v) idx.push_back((int32_t)j + nL);                                   // descriptor row of a right key point
        }
        off.push_back((int32_t)idx.size());
    }
    std::vector<uint8_t> Q = rows32(descMP);
    if (fisheye) Q.insert(Q.end(), Q.begin(), Q.begin() + (size_t)32 * nMP);                     // the right halves ask with the same descriptors
    std::vector<uint16_t> dist;
    CandidateDistances(Q, fisheye ? 2 * nMP : nMP, descF, fisheye ? off : std::vector<int32_t>(off.begin(), off.begin() + nMP + 1),
                       idx, dist);
    // state of F.mvpMapPoints[j]: -2 = held a point with observations on entry, -1 = none, >= 0 = stored by this call
    std::vector<int> st(nL + nR, -1);
    if (ex.occupied) for (int j = 0; j < nL + nR; ++j) if ((*ex.occupied)[j]) st[j] = -2;
    auto taken = [&](int j) { return st[j] == -2 || (st[j] >= 0 && hasObservations[st[j]]); };
    int nmatches = 0;
    for (int iMP = 0; iMP < nMP; ++iMP) {
        if (!inV(iMP) && !inVR(iMP)) continue;
        if (inV(iMP)) {
            int bestDist = 256, bestLevel = -1, bestDist2 = 256, bestLevel2 = -1, bestIdx = -1;
            for (int p = off[iMP]; p < off[iMP + 1]; ++p) {
                const int j = idx[p], d = dist[p];
                if (taken(j)) continue;                                                          // :80-82
                if (Nleft == -1 && ex.uRight && (*ex.uRight)[j] > 0) {                           // :84-88
                    const float er = fabs((*ex.projR)[iMP].x - (*ex.uRight)[j]);
                    if (er > radius[iMP]) continue;
                }
                if (d < bestDist) { bestDist2 = bestDist; bestDist = d; bestLevel2 = bestLevel; bestLevel = keysF[j].octave; bestIdx = j; }
                else if (d < bestDist2) { bestLevel2 = keysF[j].octave; bestDist2 = d; }
            }
            if (bestDist <= TH_HIGH && bestIdx >= 0) {
                if (bestLevel == bestLevel2 && bestDist > mfNNratio * bestDist2) continue;       // leaves the right half out too
                if (bestLevel != bestLevel2 || bestDist <= mfNNratio * bestDist2) {
                    st[bestIdx] = iMP;
                    if (Nleft != -1 && ex.leftToRight && (*ex.leftToRight)[bestIdx] != -1) {     // :114-118
                        st[(*ex.leftToRight)[bestIdx] + Nleft] = iMP;
                        nmatches++;
                    }
                    nmatches++;
                }
            }
        }
        if (Nleft != -1 && inVR(iMP) && (*ex.levelR)[iMP] != -1) {
            const int q = nMP + iMP;
            if (off[q] == off[q + 1]) continue;
            int bestDist = 256, bestLevel = -1, bestDist2 = 256, bestLevel2 = -1, bestIdx = -1;
            for (int p = off[q]; p < off[q + 1]; ++p) {
                const int j = idx[p], d = dist[p];                                               // j = Nleft + right index
                if (taken(j)) continue;
                const int oct = (*ex.keysRight)[j - Nleft].octave;
                if (d < bestDist) { bestDist2 = bestDist; bestDist = d; bestLevel2 = bestLevel; bestLevel = oct; bestIdx = j - Nleft; }
                else if (d < bestDist2) { bestLevel2 = oct; bestDist2 = d; }
            }
            if (bestDist <= TH_HIGH && bestIdx >= 0) {
                if (bestLevel == bestLevel2 && bestDist > mfNNratio * bestDist2) continue;
                if (ex.rightToLeft && (*ex.rightToLeft)[bestIdx] != -1) { st[(*ex.rightToLeft)[bestIdx]] = iMP; nmatches++; }   // :175-179
                st[bestIdx + Nleft] = iMP;
                nmatches++;
            }
        }
    }
    frameMatch.resize(nL + nR);
    for (int j = 0; j < nL + nR; ++j) frameMatch[j] = st[j] >= 0 ? st[j] : -1;
    return nmatches;
}

void ORBmatcherAccel::CandidateDistances(const std::vector<uint8_t>& Q, int nq, const cv::Mat& T, const std::vector<int32_t>& off,
                                         const std::vector<int32_t>& idx, std::vector<uint16_t>& dist) {
    dist.assign(idx.size() + 1, 0);
    const std::vector<uint8_t> Tr = rows32(T);
    if (nq > 0 && rumi_hamming_candidates(ctx, Q.data(), nq, Tr.data(), T.rows, off.data(), idx.data(), dist.data(), nullptr,
                                          nullptr, nullptr, nullptr) != RUMI_OK)
        throw std::runtime_error(std::string("ORBmatcherAccel: ") + rumi_last_error());
}

namespace {
// the rotation-consistency epilogue the projection matchers share (ORBmatcher.cc:1661-1680, :1776-1791): every entry of a
// histogram bin outside the three fullest ones is cleared and counted, duplicates included
int clear_rotation_outliers(std::vector<int>* rotHist, int L, std::vector<int>& match) {
    int ind1 = -1, ind2 = -1, ind3 = -1, removed = 0;
    three_maxima(rotHist, L, ind1, ind2, ind3);
    for (int i = 0; i < L; i++) {
        if (i == ind1 || i == ind2 || i == ind3) continue;
        for (int j : rotHist[i]) { match[j] = -1; removed++; }
    }
    return removed;
}
int rotation_bin(float angleA, float angleB, int L) {
    float rot = angleA - angleB;
    if (rot < 0.0) rot += 360.0f;
    int bin = (int)std::round(rot * (1.0f / L));
    if (bin == L) bin = 0;
    return bin;
}
}  // namespace

int ORBmatcherAccel::SearchByProjectionLastFrame(const std::vector<cv::KeyPoint>& keysC, const cv::Mat& descC,
                                                 const FrameGridAccel& gridC, const std::vector<float>& scaleFactors,
                                                 const std::vector<float>& uRight, const std::vector<uint8_t>& occupied,
                                                 float mbf, const std::vector<uint8_t>& valid,
                                                 const std::vector<cv::Point2f>& uv, const std::vector<float>& invzc,
                                                 const std::vector<int>& octaveLast, const std::vector<float>& angleLast,
                                                 const cv::Mat& descMP, const std::vector<uint8_t>& mpHasObservations, float th,
                                                 bool bForward, bool bBackward, bool checkOrientation,
                                                 std::vector<int>& curMatch) {
    curMatch.assign(keysC.size(), -1);
    std::vector<int> query;
    std::vector<float> radii;
    std::vector<int32_t> off(1, 0), idx;
    std::vector<uint8_t> Q;
    for (size_t i = 0; i < uv.size(); i++) {
        if (!valid[i]) continue;
        if (invzc[i] < 0) continue;                                                              // :1529-1530
        if (uv[i].x < gridC.mnMinX || uv[i].x > gridC.mnMaxX) continue;
        if (uv[i].y < gridC.mnMinY || uv[i].y > gridC.mnMaxY) continue;
        const int nLastOctave = octaveLast[i];
        const float radius = th * scaleFactors[nLastOctave];
        const std::vector<size_t> v = bForward    ? gridC.GetFeaturesInArea(uv[i].x, uv[i].y, radius, nLastOctave)
                                      : bBackward ? gridC.GetFeaturesInArea(uv[i].x, uv[i].y, radius, 0, nLastOctave)
                                                  : gridC.GetFeaturesInArea(uv[i].x, uv[i].y, radius, nLastOctave - 1, nLastOctave + 1);
        if (v.empty()) continue;
        query.push_back((int)i);
        radii.push_back(radius);
        idx.insert(idx.end(), v.begin(), v.end());
        off.push_back((int32_t)idx.size());
        Q.insert(Q.end(), descMP.ptr((int)i), descMP.ptr((int)i) + 32);
    }
    std::vector<uint16_t> dist;
    CandidateDistances(Q, (int)query.size(), descC, off, idx, dist);
    int nmatches = 0;
    std::vector<int> rotHist[HISTO_LENGTH];
    for (size_t qi = 0; qi < query.size(); ++qi) {
        const int i = query[qi];
        int bestDist = 256, bestIdx2 = -1;
        for (int p = off[qi]; p < off[qi + 1]; ++p) {
            const int i2 = idx[p];
            if ((!occupied.empty() && occupied[i2]) || (curMatch[i2] >= 0 && mpHasObservations[curMatch[i2]])) continue;   // :1563-1565
            if (!uRight.empty() && uRight[i2] > 0) {                                             // :1567-1572
                const float ur = uv[i].x - mbf * invzc[i];
                const float er = std::fabs(ur - uRight[i2]);
                if (er > radii[qi]) continue;
            }
            const int d = dist[p];
            if (d < bestDist) { bestDist = d; bestIdx2 = i2; }
        }
        if (bestDist <= TH_HIGH) {
            curMatch[bestIdx2] = i;
            nmatches++;
            if (checkOrientation) rotHist[rotation_bin(angleLast[i], keysC[bestIdx2].angle, HISTO_LENGTH)].push_back(bestIdx2);
        }
    }
    if (checkOrientation) nmatches -= clear_rotation_outliers(rotHist, HISTO_LENGTH, curMatch);
    return nmatches;
}

int ORBmatcherAccel::SearchByProjectionLastFrameFisheye(
    const std::vector<cv::KeyPoint>& keysC, const std::vector<cv::KeyPoint>& keysRight, const cv::Mat& descC,
    const FrameGridAccel& gridC, const FrameGridAccel& gridRight, const std::vector<float>& scaleFactors,
    const std::vector<uint8_t>& occupied, const std::vector<uint8_t>& valid, const std::vector<cv::Point2f>& uv,
    const std::vector<cv::Point2f>& uvR, const std::vector<float>& invzc, const std::vector<int>& octaveLast,
    const std::vector<float>& angleLast, const cv::Mat& descMP, const std::vector<uint8_t>& mpHasObservations, float th,
    bool bForward, bool bBackward, bool checkOrientation, std::vector<int>& curMatch) {
    const int Nleft = (int)keysC.size();
    curMatch.assign(keysC.size() + keysRight.size(), -1);
    auto window = [&](const FrameGridAccel& g, const cv::Point2f& p, float radius, int nLastOctave) {
        return bForward    ? g.GetFeaturesInArea(p.x, p.y, radius, nLastOctave)
               : bBackward ? g.GetFeaturesInArea(p.x, p.y, radius, 0, nLastOctave)
                           : g.GetFeaturesInArea(p.x, p.y, radius, nLastOctave - 1, nLastOctave + 1);
    };
    // queries [0, nq): left windows; [nq, 2 nq): the right windows of the same last-frame points
    std::vector<int> query;
    std::vector<int32_t> offL(1, 0), idxL, offR(1, 0), idxR;
    std::vector<uint8_t> Q;
    for (size_t i = 0; i < uv.size(); i++) {
        if (!valid[i]) continue;
        if (invzc[i] < 0) continue;                                                              // :1529-1530
        if (uv[i].x < gridC.mnMinX || uv[i].x > gridC.mnMaxX) continue;
        if (uv[i].y < gridC.mnMinY || uv[i].y > gridC.mnMaxY) continue;
        const int nLastOctave = octaveLast[i];
        const float radius = th * scaleFactors[nLastOctave];
        const std::vector<size_t> v = window(gridC, uv[i], radius, nLastOctave);
        if (v.empty()) continue;                                                                 // :1552-1553: leaves the right half out too
        const std::vector<size_t> vr = window(gridRight, uvR[i], radius, nLastOctave);
        query.push_back((int)i);
        idxL.insert(idxL.end(), v.begin(), v.end());
        offL.push_back((int32_t)idxL.size());
        for (size_t j : vr) idxR.push_back((int32_t)j + Nleft);                                  // descriptor row of a right key point
        offR.push_back((int32_t)idxR.size());
        Q.insert(Q.end(), descMP.ptr((int)i), descMP.ptr((int)i) + 32);
    }
    const int nq = (int)query.size();
    std::vector<int32_t> off(offL), idx(idxL);
    for (int k = 1; k <= nq; ++k) off.push_back(offL[nq] + offR[k]);
    idx.insert(idx.end(), idxR.begin(), idxR.end());
    Q.insert(Q.end(), Q.begin(), Q.begin() + (size_t)32 * nq);
    std::vector<uint16_t> dist;
    CandidateDistances(Q, 2 * nq, descC, off, idx, dist);
    int nmatches = 0;
    std::vector<int> rotHist[HISTO_LENGTH];
    for (int qi = 0; qi < nq; ++qi) {
        const int i = query[qi];
        for (int side = 0; side < 2; ++side) {
            const int ql = side * nq + qi;
            int bestDist = 256, bestIdx2 = -1;
            for (int p = off[ql]; p < off[ql + 1]; ++p) {
                const int i2 = idx[p];                                                           // right camera: Nleft + index
                if ((!occupied.empty() && occupied[i2]) || (curMatch[i2] >= 0 && mpHasObservations[curMatch[i2]])) continue;
                const int d = dist[p];
                if (d < bestDist) { bestDist = d; bestIdx2 = i2; }
            }
            if (bestDist <= TH_HIGH) {
                curMatch[bestIdx2] = i;
                nmatches++;
                if (checkOrientation) {
                    const float angleCF = side ? keysRight[bestIdx2 - Nleft].angle : keysC[bestIdx2].angle;
                    rotHist[rotation_bin(angleLast[i], angleCF, HISTO_LENGTH)].push_back(bestIdx2);
                }
            }
        }
    }
    if (checkOrientation) nmatches -= clear_rotation_outliers(rotHist, HISTO_LENGTH, curMatch);
    return nmatches;
}

int ORBmatcherAccel::SearchByProjectionKeyFrame(const std::vector<cv::KeyPoint>& keysC, const cv::Mat& descC,
                                                const FrameGridAccel& gridC, const std::vector<float>& scaleFactors,
                                                const std::vector<uint8_t>& occupied, const std::vector<uint8_t>& valid,
                                                const std::vector<cv::Point2f>& uv, const std::vector<float>& dist3D,
                                                const std::vector<float>& minDistance, const std::vector<float>& maxDistance,
                                                const std::vector<int>& predictedLevel, const std::vector<float>& angleKF,
                                                const cv::Mat& descMP, float th, int ORBdist, bool checkOrientation,
                                                std::vector<int>& curMatch) {
    curMatch.assign(keysC.size(), -1);
    std::vector<int> query;
    std::vector<int32_t> off(1, 0), idx;
    std::vector<uint8_t> Q;
    for (size_t i = 0; i < uv.size(); i++) {
        if (!valid[i]) continue;
        if (uv[i].x < gridC.mnMinX || uv[i].x > gridC.mnMaxX) continue;
        if (uv[i].y < gridC.mnMinY || uv[i].y > gridC.mnMaxY) continue;
        if (dist3D[i] < minDistance[i] || dist3D[i] > maxDistance[i]) continue;                  // :1724-1725
        const int nPredictedLevel = predictedLevel[i];
        const float radius = th * scaleFactors[nPredictedLevel];
        const std::vector<size_t> v = gridC.GetFeaturesInArea(uv[i].x, uv[i].y, radius, nPredictedLevel - 1, nPredictedLevel + 1);
        if (v.empty()) continue;
        query.push_back((int)i);
        idx.insert(idx.end(), v.begin(), v.end());
        off.push_back((int32_t)idx.size());
        Q.insert(Q.end(), descMP.ptr((int)i), descMP.ptr((int)i) + 32);
    }
    std::vector<uint16_t> dist;
    CandidateDistances(Q, (int)query.size(), descC, off, idx, dist);
    int nmatches = 0;
    std::vector<int> rotHist[HISTO_LENGTH];
    for (size_t qi = 0; qi < query.size(); ++qi) {
        const int i = query[qi];
        int bestDist = 256, bestIdx2 = -1;
        for (int p = off[qi]; p < off[qi + 1]; ++p) {
            const int i2 = idx[p];
            if ((!occupied.empty() && occupied[i2]) || curMatch[i2] >= 0) continue;              // :1742-1743
            const int d = dist[p];
            if (d < bestDist) { bestDist = d; bestIdx2 = i2; }
        }
        if (bestDist <= ORBdist) {
            curMatch[bestIdx2] = i;
            nmatches++;
            if (checkOrientation) rotHist[rotation_bin(angleKF[i], keysC[bestIdx2].angle, HISTO_LENGTH)].push_back(bestIdx2);
        }
    }
    if (checkOrientation) nmatches -= clear_rotation_outliers(rotHist, HISTO_LENGTH, curMatch);
    return nmatches;
}

int ORBmatcherAccel::FuseSearch(const std::vector<cv::KeyPoint>& keysK, const cv::Mat& descK, const FrameGridAccel& gridK,
                                const std::vector<float>& scaleFactors, const std::vector<float>& invLevelSigma2,
                                const std::vector<float>& uRight, const std::vector<uint8_t>& valid,
                                const std::vector<cv::Point2f>& uv, const std::vector<float>& ur,
                                const std::vector<float>& dist3D, const std::vector<float>& minDistance,
                                const std::vector<float>& maxDistance, const std::vector<int>& predictedLevel,
                                const cv::Mat& descMP, float th, std::vector<int>& bestIdx, std::vector<int>& bestDist,
                                int thDist) {
    const int nMPs = (int)uv.size();
    bestIdx.assign(nMPs, -1);
    bestDist.assign(nMPs, 256);
    std::vector<int> query;
    std::vector<int32_t> off(1, 0), idx;
    std::vector<uint8_t> Q;
    for (int i = 0; i < nMPs; i++) {
        if (!valid[i]) continue;
        const float u = uv[i].x, v = uv[i].y;
        if (!(u >= gridK.mnMinX && u < gridK.mnMaxX && v >= gridK.mnMinY && v < gridK.mnMaxY)) continue;    // KeyFrame::IsInImage
        if (dist3D[i] < minDistance[i] || dist3D[i] > maxDistance[i]) continue;
        const int nPredictedLevel = predictedLevel[i];
        const float radius = th * scaleFactors[nPredictedLevel];
        const std::vector<size_t> vIndices = gridK.GetFeaturesInArea(u, v, radius);
        if (vIndices.empty()) continue;
        // the gates that precede DescriptorDistance (:1107-1141) thin the list out; the kernel scans what is left
        for (size_t j : vIndices) {
            const cv::KeyPoint& kp = keysK[j];
            const int kpLevel = kp.octave;
            if (kpLevel < nPredictedLevel - 1 || kpLevel > nPredictedLevel) continue;
            if (uRight[j] >= 0) {
                const float ex = u - kp.pt.x, ey = v - kp.pt.y, er = ur[i] - uRight[j];
                const float e2 = ex * ex + ey * ey + er * er;
                if (e2 * invLevelSigma2[kpLevel] > 7.8) continue;
            } else {
                const float ex = u - kp.pt.x, ey = v - kp.pt.y;
                const float e2 = ex * ex + ey * ey;
                if (e2 * invLevelSigma2[kpLevel] > 5.99) continue;
            }
            idx.push_back((int32_t)j);
        }
        query.push_back(i);
        off.push_back((int32_t)idx.size());
        Q.insert(Q.end(), descMP.ptr(i), descMP.ptr(i) + 32);
    }
    const int nq = (int)query.size();
    if (nq == 0) return 0;
    std::vector<uint16_t> dist(idx.size() + 1), d1(nq), d2(nq);
    std::vector<int32_t> i1(nq), i2(nq);
    const std::vector<uint8_t> T = rows32(descK);
    if (rumi_hamming_candidates(ctx, Q.data(), nq, T.data(), descK.rows, off.data(), idx.data(), dist.data(), i1.data(), d1.data(),
                                i2.data(), d2.data()) != RUMI_OK)
        throw std::runtime_error(std::string("ORBmatcherAccel: ") + rumi_last_error());
    int nFused = 0;
    for (int q = 0; q < nq; ++q) {
        bestDist[query[q]] = d1[q];
        if (d1[q] <= thDist) { bestIdx[query[q]] = i1[q]; nFused++; }
    }
    return nFused;
}

int ORBmatcherAccel::FuseSearchSim3(const std::vector<cv::KeyPoint>& keysK, const cv::Mat& descK, const FrameGridAccel& gridK,
                                    const std::vector<float>& scaleFactors, const std::vector<uint8_t>& valid,
                                    const std::vector<cv::Point2f>& uv, const std::vector<float>& dist3D,
                                    const std::vector<float>& minDistance, const std::vector<float>& maxDistance,
                                    const std::vector<int>& predictedLevel, const cv::Mat& descMP, float th,
                                    std::vector<int>& bestIdx, std::vector<int>& bestDist, int thDist) {
    // zero inverse sigma and "no right coordinate" switch the chi-square gates of FuseSearch off (0 > 5.99 never holds)
    const std::vector<float> zeroSigma(scaleFactors.size(), 0.0f), noRight(keysK.size(), -1.0f), ur(uv.size(), 0.0f);
    return FuseSearch(keysK, descK, gridK, scaleFactors, zeroSigma, noRight, valid, uv, ur, dist3D, minDistance, maxDistance,
                      predictedLevel, descMP, th, bestIdx, bestDist, thDist);
}

int ORBmatcherAccel::SearchBySim3(const std::vector<cv::KeyPoint>& keys1, const cv::Mat& desc1, const FrameGridAccel& grid1,
                                  const std::vector<cv::KeyPoint>& keys2, const cv::Mat& desc2, const FrameGridAccel& grid2,
                                  const std::vector<float>& scaleFactors1, const std::vector<float>& scaleFactors2,
                                  const std::vector<uint8_t>& valid1, const std::vector<cv::Point2f>& uv12,
                                  const std::vector<float>& dist12, const std::vector<float>& min1,
                                  const std::vector<float>& max1, const std::vector<int>& level12,
                                  const std::vector<uint8_t>& valid2, const std::vector<cv::Point2f>& uv21,
                                  const std::vector<float>& dist21, const std::vector<float>& min2,
                                  const std::vector<float>& max2, const std::vector<int>& level21, float th,
                                  std::vector<int>& match12) {
    std::vector<int> vnMatch1, vnMatch2, dd;
    FuseSearchSim3(keys2, desc2, grid2, scaleFactors2, valid1, uv12, dist12, min1, max1, level12, desc1, th, vnMatch1, dd, TH_HIGH);
    FuseSearchSim3(keys1, desc1, grid1, scaleFactors1, valid2, uv21, dist21, min2, max2, level21, desc2, th, vnMatch2, dd, TH_HIGH);
    match12.assign(keys1.size(), -1);
    int nFound = 0;
    for (size_t i1 = 0; i1 < keys1.size(); i1++) {                                // :1483-1494
        const int idx2 = vnMatch1[i1];
        if (idx2 >= 0 && vnMatch2[idx2] == (int)i1) { match12[i1] = idx2; nFound++; }
    }
    return nFound;
}

int ORBmatcherAccel::SearchByProjectionSim3(const std::vector<cv::KeyPoint>& keysK, const cv::Mat& descK,
                                            const FrameGridAccel& gridK, const std::vector<float>& scaleFactors,
                                            const std::vector<uint8_t>& occupied, const std::vector<uint8_t>& valid,
                                            const std::vector<cv::Point2f>& uv, const std::vector<float>& dist3D,
                                            const std::vector<float>& minDistance, const std::vector<float>& maxDistance,
                                            const std::vector<int>& predictedLevel, const cv::Mat& descMP, int th,
                                            float ratioHamming, std::vector<int>& kfMatch) {
    kfMatch.assign(keysK.size(), -1);
    std::vector<int> query;
    std::vector<int32_t> off(1, 0), idx;
    std::vector<uint8_t> Q;
    for (size_t iMP = 0; iMP < uv.size(); iMP++) {
        if (!valid[iMP]) continue;
        const float u = uv[iMP].x, v = uv[iMP].y;
        if (!(u >= gridK.mnMinX && u < gridK.mnMaxX && v >= gridK.mnMinY && v < gridK.mnMaxY)) continue;    // KeyFrame::IsInImage
        if (dist3D[iMP] < minDistance[iMP] || dist3D[iMP] > maxDistance[iMP]) continue;
        const float radius = th * scaleFactors[predictedLevel[iMP]];
        const std::vector<size_t> vIndices = gridK.GetFeaturesInArea(u, v, radius);
        if (vIndices.empty()) continue;
        query.push_back((int)iMP);
        idx.insert(idx.end(), vIndices.begin(), vIndices.end());
        off.push_back((int32_t)idx.size());
        Q.insert(Q.end(), descMP.ptr((int)iMP), descMP.ptr((int)iMP) + 32);
    }
    std::vector<uint16_t> dist;
    CandidateDistances(Q, (int)query.size(), descK, off, idx, dist);
    int nmatches = 0;
    for (size_t qi = 0; qi < query.size(); ++qi) {
        const int iMP = query[qi], nPredictedLevel = predictedLevel[iMP];
        int bestDist = 256, bestIdx = -1;
        for (int p = off[qi]; p < off[qi + 1]; ++p) {
            const int j = idx[p];
            if ((!occupied.empty() && occupied[j]) || kfMatch[j] >= 0) continue;                 // :443-444
            const int kpLevel = keysK[j].octave;
            if (kpLevel < nPredictedLevel - 1 || kpLevel > nPredictedLevel) continue;
            const int d = dist[p];
            if (d < bestDist) { bestDist = d; bestIdx = j; }
        }
        if (bestDist <= TH_LOW * ratioHamming) { kfMatch[bestIdx] = iMP; nmatches++; }
    }
    return nmatches;
}

int ORBmatcherAccel::AssociateSubmap(struct rumi_orb* extractor, const std::vector<cv::Mat>& images1,
                                     const std::vector<std::vector<cv::KeyPoint>>& keys1,
                                     const std::vector<std::vector<uint8_t>>& valid1, const std::vector<cv::Mat>& images2,
                                     const std::vector<std::vector<cv::KeyPoint>>& keys2,
                                     const std::vector<std::vector<uint8_t>>& valid2, std::vector<std::vector<int>>& match12,
                                     int th) {
    const size_t np = images1.size();
    if (images2.size() != np || keys1.size() != np || keys2.size() != np || valid1.size() != np || valid2.size() != np)
        throw std::runtime_error("ORBmatcherAccel::AssociateSubmap: argument lists differ in length");
    match12.assign(np, std::vector<int>());
    if (np == 0) return 0;
    // descriptors of the key points that carry a map point, one batched CloudFrameComputeDescriptors per side
    auto describe = [&](const std::vector<cv::Mat>& images, const std::vector<std::vector<cv::KeyPoint>>& keys,
                        const std::vector<std::vector<uint8_t>>& valid, std::vector<std::vector<int>>& sel,
                        std::vector<int32_t>& kpOff, std::vector<uint8_t>& desc) {
        const int w = images[0].cols, h = images[0].rows;
        std::vector<uint8_t> packed((size_t)w * h * np);
        std::vector<rumi_kp> k;
        sel.assign(np, std::vector<int>());
        kpOff.assign(1, 0);
        for (size_t p = 0; p < np; ++p) {
            if (images[p].cols != w || images[p].rows != h) throw std::runtime_error("AssociateSubmap: image shapes differ");
            for (int y = 0; y < h; ++y) std::memcpy(packed.data() + ((size_t)p * h + y) * w, images[p].ptr(y), (size_t)w);
            for (size_t i = 0; i < keys[p].size(); ++i)
                if (valid[p][i]) {
                    rumi_kp r;
                    std::memcpy(&r, &keys[p][i], sizeof(r));
                    k.push_back(r);
                    sel[p].push_back((int)i);
                }
            kpOff.push_back((int32_t)k.size());
        }
        desc.assign(32 * k.size() + 32, 0);
        if (rumi_orb_describe_batch(extractor, packed.data(), (int)np, w, h, (size_t)w, (size_t)w * h, k.data(), kpOff.data(),
                                    desc.data()) < 0)
            throw std::runtime_error(std::string("ORBmatcherAccel: ") + rumi_last_error());
    };
    std::vector<std::vector<int>> sel1, sel2;
    std::vector<int32_t> off1, off2;
    std::vector<uint8_t> Q, T;
    describe(images1, keys1, valid1, sel1, off1, Q);
    describe(images2, keys2, valid2, sel2, off2, T);
    std::vector<int32_t> segs;
    for (size_t p = 0; p < np; ++p) {
        segs.push_back(off1[p]); segs.push_back(off1[p + 1] - off1[p]);
        segs.push_back(off2[p]); segs.push_back(off2[p + 1] - off2[p]);
    }
    const int nq = off1[np], nt = off2[np];
    std::vector<int32_t> idx(nq + 1);
    std::vector<uint16_t> d1(nq + 1), d2(nq + 1);
    if (rumi_hamming_top2_pairs(ctx, Q.data(), nq, T.data(), nt, segs.data(), (int)np, idx.data(), d1.data(), d2.data()) != RUMI_OK)
        throw std::runtime_error(std::string("ORBmatcherAccel: ") + rumi_last_error());
    int total = 0;
    for (size_t p = 0; p < np; ++p) {
        match12[p].assign(keys1[p].size(), -1);
        for (int q = 0; q < off1[p + 1] - off1[p]; ++q) {
            const int g = off1[p] + q, bestDist1 = d1[g], bestDist2 = d2[g];
            if (idx[g] >= 0 && bestDist1 <= th &&
                static_cast<float>(bestDist1) < mfNNratio * static_cast<float>(bestDist2)) {   // ORBmatcher.cc:290-291
                match12[p][sel1[p][q]] = sel2[p][idx[g]];
                ++total;
            }
        }
    }
    return total;
}

}  // namespace ORB_SLAM3
