// ORACLE -- TEST INFRASTRUCTURE ONLY (used by tests/, __graft_entry__.smoke() and bench.py's CPU baseline legs).
// CPU restatement of the candidate-list matchers of the reference and of the feature grid that feeds them:
//   Frame::AssignFeaturesToGrid / PosInGrid / GetFeaturesInArea      R/lib_src/Frame.cc:441-466, 752-767, 695-750
//   ORBmatcher::SearchForInitialization                              R/lib_src/ORBmatcher.cc:581-680
//   ORBmatcher::SearchByProjection(Frame&, vpMapPoints, th, ...)     R/lib_src/ORBmatcher.cc:39-189 (mono frame, Nleft == -1)
//   ORBmatcher::SearchByProjection(CurrentFrame, LastFrame, th, bMono)   R/lib_src/ORBmatcher.cc:1498-1684 (Nleft == -1)
//   ORBmatcher::SearchByProjection(CurrentFrame, pKF, sAlreadyFound, th, ORBdist)   R/lib_src/ORBmatcher.cc:1685-1794
//   ORBmatcher::Fuse(pKF, vpMapPoints, th, bRight)  (matching core)   R/lib_src/ORBmatcher.cc:1015-1181 (NLeft == -1)
//   ORBmatcher::SearchByProjection(pKF, Scw, vpPoints, [vpPointsKFs,] vpMatched, ...)   R/lib_src/ORBmatcher.cc:372-471, 473-580
//   CloudMerging's pixel-distance key-point association              R/lib_src/CloudMerging.cc:503-551
// Pinned against the UNMODIFIED reference functions compiled into oracle/_ref/librefframe.so
// (tests/test_ref_frame_pin.py).  FRAME_GRID_COLS = 64, FRAME_GRID_ROWS = 48 (R/include/cloud_edge_slam_lib/Frame.h:42-43).
#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

namespace {

struct KP { float x, y, size, angle, response; int32_t octave, class_id; };
static_assert(sizeof(KP) == 28, "cv::KeyPoint layout");

constexpr int kGridCols = 64, kGridRows = 48;
constexpr int TH_HIGH = 100, TH_LOW = 50, HISTO_LENGTH = 30;      // ORBmatcher.cc:31-33

// ORBmatcher::DescriptorDistance (ORBmatcher.cc:1830-1844)
inline int descriptor_distance(const uint8_t* a, const uint8_t* b) {
    int dist = 0;
    for (int i = 0; i < 8; ++i) {
        uint32_t x, y;
        std::memcpy(&x, a + 4 * i, 4); std::memcpy(&y, b + 4 * i, 4);
        uint32_t v = x ^ y;
        v = v - ((v >> 1) & 0x55555555u);
        v = (v & 0x33333333u) + ((v >> 2) & 0x33333333u);
        dist += (int)((((v + (v >> 4)) & 0xF0F0F0Fu) * 0x1010101u) >> 24);
    }
    return dist;
}

struct Grid {
    float minX, minY, wInv, hInv;
    const KP* kps; int n;
    std::vector<int> cell[kGridCols][kGridRows];
    Grid(const KP* k, int nn, int x0, int y0, int x1, int y1) : kps(k), n(nn) {
        minX = (float)x0; minY = (float)y0;
        wInv = static_cast<float>(kGridCols) / static_cast<float>(x1 - x0);      // Frame.cc:98-99
        hInv = static_cast<float>(kGridRows) / static_cast<float>(y1 - y0);
        for (int i = 0; i < n; ++i) {                                            // AssignFeaturesToGrid :455-465
            const int px = (int)std::round((kps[i].x - minX) * wInv), py = (int)std::round((kps[i].y - minY) * hInv);   // PosInGrid
            if (px < 0 || px >= kGridCols || py < 0 || py >= kGridRows) continue;
            cell[px][py].push_back(i);
        }
    }
    // GetFeaturesInArea :695-750 (Nleft == -1)
    void query(float x, float y, float r, int minLevel, int maxLevel, std::vector<int>& out) const {
        out.clear();
        const int nMinCellX = std::max(0, (int)std::floor((x - minX - r) * wInv));
        if (nMinCellX >= kGridCols) return;
        const int nMaxCellX = std::min(kGridCols - 1, (int)std::ceil((x - minX + r) * wInv));
        if (nMaxCellX < 0) return;
        const int nMinCellY = std::max(0, (int)std::floor((y - minY - r) * hInv));
        if (nMinCellY >= kGridRows) return;
        const int nMaxCellY = std::min(kGridRows - 1, (int)std::ceil((y - minY + r) * hInv));
        if (nMaxCellY < 0) return;
        const bool bCheckLevels = (minLevel > 0) || (maxLevel >= 0);
        for (int ix = nMinCellX; ix <= nMaxCellX; ++ix)
            for (int iy = nMinCellY; iy <= nMaxCellY; ++iy)
                for (int j : cell[ix][iy]) {
                    const KP& kp = kps[j];
                    if (bCheckLevels) {
                        if (kp.octave < minLevel) continue;
                        if (maxLevel >= 0 && kp.octave > maxLevel) continue;
                    }
                    const float dx = kp.x - x, dy = kp.y - y;
                    if (std::fabs(dx) < r && std::fabs(dy) < r) out.push_back(j);
                }
    }
};

// ORBmatcher::ComputeThreeMaxima (ORBmatcher.cc:1795-1826)
void three_maxima(const std::vector<int>* histo, int L, int& ind1, int& ind2, int& ind3) {
    int max1 = 0, max2 = 0, max3 = 0;
    for (int i = 0; i < L; ++i) {
        const int s = (int)histo[i].size();
        if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
        else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
        else if (s > max3) { max3 = s; ind3 = i; }
    }
    if (max2 < 0.1f * (float)max1) { ind2 = -1; ind3 = -1; }
    else if (max3 < 0.1f * (float)max1) { ind3 = -1; }
}

}  // namespace

extern "C" {

int mo_features_in_area(const void* kps, int n, int minX, int minY, int maxX, int maxY, float x, float y, float r,
                        int minLevel, int maxLevel, int32_t* out, int cap) {
    Grid g((const KP*)kps, n, minX, minY, maxX, maxY);
    std::vector<int> v;
    g.query(x, y, r, minLevel, maxLevel, v);
    for (size_t i = 0; i < v.size() && (int)i < cap; ++i) out[i] = v[i];
    return (int)v.size();
}

// CSR candidate lists of many queries on one grid: off[nq + 1], idx[cap]; returns the total (may exceed cap: call again).
int mo_candidate_lists(const void* kps, int n, int minX, int minY, int maxX, int maxY, const float* qxy, const float* qr,
                       const int32_t* qMinLevel, const int32_t* qMaxLevel, int nq, int32_t* off, int32_t* idx, int cap) {
    Grid g((const KP*)kps, n, minX, minY, maxX, maxY);
    std::vector<int> v;
    int tot = 0;
    for (int q = 0; q < nq; ++q) {
        off[q] = tot;
        g.query(qxy[2 * q], qxy[2 * q + 1], qr[q], qMinLevel ? qMinLevel[q] : -1, qMaxLevel ? qMaxLevel[q] : -1, v);
        for (int j : v) { if (tot < cap) idx[tot] = j; ++tot; }
    }
    off[nq] = tot;
    return tot;
}

// ORBmatcher::SearchForInitialization (ORBmatcher.cc:581-680).  prev: vbPrevMatched [n1][2], updated in place.
int mo_search_for_initialization(const void* k1v, const uint8_t* d1, int n1, const void* k2v, const uint8_t* d2, int n2,
                                 int minX, int minY, int maxX, int maxY, float* prev, int windowSize, float ratio, int checkOri,
                                 int32_t* vnMatches12) {
    const KP* k1 = (const KP*)k1v; const KP* k2 = (const KP*)k2v;
    Grid g(k2, n2, minX, minY, maxX, maxY);
    int nmatches = 0;
    for (int i = 0; i < n1; ++i) vnMatches12[i] = -1;
    std::vector<int> rotHist[HISTO_LENGTH];
    const float factor = 1.0f / HISTO_LENGTH;
    std::vector<int> vMatchedDistance(n2, INT_MAX), vnMatches21(n2, -1), vIndices2;
    for (int i1 = 0; i1 < n1; ++i1) {
        const int level1 = k1[i1].octave;
        if (level1 > 0) continue;
        g.query(prev[2 * i1], prev[2 * i1 + 1], (float)windowSize, level1, level1, vIndices2);
        if (vIndices2.empty()) continue;
        int bestDist = INT_MAX, bestDist2 = INT_MAX, bestIdx2 = -1;
        for (int i2 : vIndices2) {
            const int dist = descriptor_distance(d1 + 32 * (size_t)i1, d2 + 32 * (size_t)i2);
            if (vMatchedDistance[i2] <= dist) continue;
            if (dist < bestDist) { bestDist2 = bestDist; bestDist = dist; bestIdx2 = i2; }
            else if (dist < bestDist2) bestDist2 = dist;
        }
        if (bestDist <= TH_LOW) {
            if (bestDist < (float)bestDist2 * ratio) {
                if (vnMatches21[bestIdx2] >= 0) { vnMatches12[vnMatches21[bestIdx2]] = -1; nmatches--; }
                vnMatches12[i1] = bestIdx2;
                vnMatches21[bestIdx2] = i1;
                vMatchedDistance[bestIdx2] = bestDist;
                nmatches++;
                if (checkOri) {
                    float rot = k1[i1].angle - k2[bestIdx2].angle;
                    if (rot < 0.0) rot += 360.0f;
                    int bin = (int)std::round(rot * factor);
                    if (bin == HISTO_LENGTH) bin = 0;
                    rotHist[bin].push_back(i1);
                }
            }
        }
    }
    if (checkOri) {
        int ind1 = -1, ind2 = -1, ind3 = -1;
        three_maxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
        for (int i = 0; i < HISTO_LENGTH; ++i) {
            if (i == ind1 || i == ind2 || i == ind3) continue;
            for (int idx1 : rotHist[i])
                if (vnMatches12[idx1] >= 0) { vnMatches12[idx1] = -1; nmatches--; }
        }
    }
    for (int i1 = 0; i1 < n1; ++i1)
        if (vnMatches12[i1] >= 0) { prev[2 * i1] = k2[vnMatches12[i1]].x; prev[2 * i1 + 1] = k2[vnMatches12[i1]].y; }
    return nmatches;
}

// ORBmatcher::SearchByProjection(Frame&, vpMapPoints, th, false, ..) on a mono frame without depth (Nleft == -1, mvuRight
// <= 0): ORBmatcher.cc:39-118.  frameMatch[j] = map point assigned to frame feature j (or -1).
int mo_search_by_projection(const void* kFv, const uint8_t* dF, int nF, const float* scaleFactors, int minX, int minY, int maxX,
                            int maxY, const float* proj, const int32_t* level, const float* viewCos, const uint8_t* dMP,
                            const uint8_t* hasObs, int nMP, float th, float ratio, int32_t* frameMatch) {
    const KP* kF = (const KP*)kFv;
    Grid g(kF, nF, minX, minY, maxX, maxY);
    for (int j = 0; j < nF; ++j) frameMatch[j] = -1;
    int nmatches = 0;
    const bool bFactor = th != 1.0;
    std::vector<int> vIndices;
    for (int iMP = 0; iMP < nMP; ++iMP) {
        const int nPredictedLevel = level[iMP];
        float r = viewCos[iMP] > 0.998 ? 2.5f : 4.0f;                   // RadiusByViewingCos :191-196
        if (bFactor) r *= th;
        g.query(proj[2 * iMP], proj[2 * iMP + 1], r * scaleFactors[nPredictedLevel], nPredictedLevel - 1, nPredictedLevel, vIndices);
        if (vIndices.empty()) continue;
        int bestDist = 256, bestLevel = -1, bestDist2 = 256, bestLevel2 = -1, bestIdx = -1;
        for (int idx : vIndices) {
            if (frameMatch[idx] >= 0 && hasObs[frameMatch[idx]]) continue;           // F.mvpMapPoints[idx]->Observations() > 0
            const int dist = descriptor_distance(dMP + 32 * (size_t)iMP, dF + 32 * (size_t)idx);
            if (dist < bestDist) { bestDist2 = bestDist; bestDist = dist; bestLevel2 = bestLevel; bestLevel = kF[idx].octave; bestIdx = idx; }
            else if (dist < bestDist2) { bestLevel2 = kF[idx].octave; bestDist2 = dist; }
        }
        if (bestDist <= TH_HIGH) {
            if (bestLevel == bestLevel2 && bestDist > ratio * bestDist2) continue;
            if (bestLevel != bestLevel2 || bestDist <= ratio * bestDist2) { frameMatch[bestIdx] = iMP; nmatches++; }
        }
    }
    return nmatches;
}

// The WHOLE of ORBmatcher::SearchByProjection(Frame&, vpMapPoints, th, ..) (ORBmatcher.cc:39-189) as Tracking::SearchLocalPoints
// meets it: a frame that already holds matches (occupied[j] = F.mvpMapPoints[j] is a point with observations on entry,
// :80-82), rectified stereo / RGB-D frames (uRight[j] = F.mvuRight[j] > 0: the right-image gate of :84-88 against
// mTrackProjXR) and stereo-fisheye rigs (nR > 0: Nleft = nL, the right camera's key points kR, descriptor rows nL + i, its own
// grid, the second half of the loop :125-185, and the cross assignments through l2r = mvLeftToRightMatch / r2l =
// mvRightToLeftMatch, :114-118 / :175-179).  Per map point: inView / inViewR = mbTrackInView / mbTrackInViewR,
// proj = (mTrackProjX, mTrackProjY), projR = (mTrackProjXR, mTrackProjYR), level / levelR, viewCos / viewCosR, descriptor,
// hasObs = Observations() > 0.  bFarPoints and isBad() are the caller's filter.  frameMatch[j] (nL + nR entries) = map point
// this call stored in F.mvpMapPoints[j], -1 = untouched.  Nullable: kR, uRight, occupied, l2r, r2l, inViewR and the *R arrays.
int mo_search_by_projection_ex(const void* kLv, int nL, const void* kRv, int nR, const uint8_t* dF, const float* scaleFactors,
                               int minX, int minY, int maxX, int maxY, const float* uRight, const uint8_t* occupied,
                               const int32_t* l2r, const int32_t* r2l, const uint8_t* inView, const uint8_t* inViewR,
                               const float* proj, const float* projR, const int32_t* level, const int32_t* levelR,
                               const float* viewCos, const float* viewCosR, const uint8_t* dMP, const uint8_t* hasObs, int nMP,
                               float th, float ratio, int32_t* frameMatch) {
    const KP *kL = (const KP*)kLv, *kR = (const KP*)kRv;
    const int Nleft = nR > 0 ? nL : -1, N = nL + nR;
    Grid g(kL, nL, minX, minY, maxX, maxY), gR(kR, nR, minX, minY, maxX, maxY);
    // state of F.mvpMapPoints[j]: -2 = a point with observations that was there on entry, -1 = none (or one without
    // observations), >= 0 = map point stored by this call
    std::vector<int> st(N, -1);
    if (occupied) for (int j = 0; j < N; ++j) if (occupied[j]) st[j] = -2;
    auto taken = [&](int j) { return st[j] == -2 || (st[j] >= 0 && hasObs[st[j]]); };
    int nmatches = 0;
    const bool bFactor = th != 1.0;
    std::vector<int> vIndices;
    for (int iMP = 0; iMP < nMP; ++iMP) {
        const bool inV = inView ? inView[iMP] != 0 : true, inVR = inViewR ? inViewR[iMP] != 0 : false;
        if (!inV && !inVR) continue;
        if (inV) {
            const int nPredictedLevel = level[iMP];
            float r = viewCos[iMP] > 0.998 ? 2.5f : 4.0f;                 // RadiusByViewingCos :191-196
            if (bFactor) r *= th;
            g.query(proj[2 * iMP], proj[2 * iMP + 1], r * scaleFactors[nPredictedLevel], nPredictedLevel - 1, nPredictedLevel, vIndices);
            if (!vIndices.empty()) {
                int bestDist = 256, bestLevel = -1, bestDist2 = 256, bestLevel2 = -1, bestIdx = -1;
                for (int idx : vIndices) {
                    if (taken(idx)) continue;
                    if (Nleft == -1 && uRight && uRight[idx] > 0) {
                        const float er = std::fabs(projR[2 * iMP] - uRight[idx]);
                        if (er > r * scaleFactors[nPredictedLevel]) continue;
                    }
                    const int dist = descriptor_distance(dMP + 32 * (size_t)iMP, dF + 32 * (size_t)idx);
                    if (dist < bestDist) { bestDist2 = bestDist; bestDist = dist; bestLevel2 = bestLevel; bestLevel = kL[idx].octave; bestIdx = idx; }
                    else if (dist < bestDist2) { bestLevel2 = kL[idx].octave; bestDist2 = dist; }
                }
                if (bestDist <= TH_HIGH) {
                    if (bestLevel == bestLevel2 && bestDist > ratio * bestDist2) continue;      // skips the right half too
                    if (bestLevel != bestLevel2 || bestDist <= ratio * bestDist2) {
                        st[bestIdx] = iMP;
                        if (Nleft != -1 && l2r && l2r[bestIdx] != -1) { st[l2r[bestIdx] + Nleft] = iMP; nmatches++; }
                        nmatches++;
                    }
                }
            }
        }
        if (Nleft != -1 && inVR) {
            const int nPredictedLevel = levelR[iMP];
            if (nPredictedLevel != -1) {
                const float r = viewCosR[iMP] > 0.998 ? 2.5f : 4.0f;      // no th factor on this side (:129)
                gR.query(projR[2 * iMP], projR[2 * iMP + 1], r * scaleFactors[nPredictedLevel], nPredictedLevel - 1, nPredictedLevel, vIndices);
                if (vIndices.empty()) continue;
                int bestDist = 256, bestLevel = -1, bestDist2 = 256, bestLevel2 = -1, bestIdx = -1;
                for (int idx : vIndices) {
                    if (taken(idx + Nleft)) continue;
                    const int dist = descriptor_distance(dMP + 32 * (size_t)iMP, dF + 32 * (size_t)(idx + Nleft));
                    if (dist < bestDist) { bestDist2 = bestDist; bestDist = dist; bestLevel2 = bestLevel; bestLevel = kR[idx].octave; bestIdx = idx; }
                    else if (dist < bestDist2) { bestLevel2 = kR[idx].octave; bestDist2 = dist; }
                }
                if (bestDist <= TH_HIGH) {
                    if (bestLevel == bestLevel2 && bestDist > ratio * bestDist2) continue;
                    if (r2l && r2l[bestIdx] != -1) { st[r2l[bestIdx]] = iMP; nmatches++; }
                    st[bestIdx + Nleft] = iMP;
                    nmatches++;
                }
            }
        }
    }
    for (int j = 0; j < N; ++j) frameMatch[j] = st[j] >= 0 ? st[j] : -1;
    return nmatches;
}

// ORBmatcher::SearchByProjection(Frame& CurrentFrame, const Frame& LastFrame, th, bMono) (ORBmatcher.cc:1498-1684), the
// matcher of Tracking::TrackWithMotionModel, for frames without a second fisheye camera (Nleft == -1).  Per feature i of
// the last frame: valid[i] = it has a map point and is not an outlier (:1518-1520); uv[i] = projection of that point into
// the current frame, invz[i] = 1 / depth (:1522-1534, the caller's pose and camera model); octave[i], angleLast[i] = its
// key point (:1541, :1601); dMP[i] = the map point's descriptor; mpHasObs[i] = Observations() > 0.  Current frame: key
// points, descriptors, grid bounds, scale factors, uRight (mvuRight, <= 0: none), occupied[j] = mvpMapPoints[j] already
// holds a point with observations.  forward / backward = bForward / bBackward (:1513-1514).  curMatch[j] = last-frame
// feature whose map point ends up in CurrentFrame.mvpMapPoints[j], or -1.
int mo_search_by_projection_last(const void* kCv, const uint8_t* dC, int nC, const float* scaleFactors, int minX, int minY,
                                 int maxX, int maxY, const float* uRight, const uint8_t* occupied, float mbf,
                                 const uint8_t* valid, const float* uv, const float* invz, const int32_t* octave,
                                 const float* angleLast, const uint8_t* dMP, const uint8_t* mpHasObs, int nL, float th,
                                 int forward, int backward, int checkOri, int32_t* curMatch) {
    const KP* kC = (const KP*)kCv;
    Grid g(kC, nC, minX, minY, maxX, maxY);
    for (int j = 0; j < nC; ++j) curMatch[j] = -1;
    int nmatches = 0;
    std::vector<int> rotHist[HISTO_LENGTH];
    const float factor = 1.0f / HISTO_LENGTH;
    std::vector<int> vIndices2;
    for (int i = 0; i < nL; ++i) {
        if (!valid[i]) continue;
        const float invzc = invz[i];
        if (invzc < 0) continue;
        const float u = uv[2 * i], v = uv[2 * i + 1];
        if (u < (float)minX || u > (float)maxX) continue;
        if (v < (float)minY || v > (float)maxY) continue;
        const int nLastOctave = octave[i];
        const float radius = th * scaleFactors[nLastOctave];
        if (forward) g.query(u, v, radius, nLastOctave, -1, vIndices2);
        else if (backward) g.query(u, v, radius, 0, nLastOctave, vIndices2);
        else g.query(u, v, radius, nLastOctave - 1, nLastOctave + 1, vIndices2);
        if (vIndices2.empty()) continue;
        int bestDist = 256, bestIdx2 = -1;
        for (int i2 : vIndices2) {
            if (occupied[i2] || (curMatch[i2] >= 0 && mpHasObs[curMatch[i2]])) continue;      // :1563-1565
            if (uRight && uRight[i2] > 0) {                                                   // :1567-1572
                const float ur = u - mbf * invzc;
                const float er = std::fabs(ur - uRight[i2]);
                if (er > radius) continue;
            }
            const int dist = descriptor_distance(dMP + 32 * (size_t)i, dC + 32 * (size_t)i2);
            if (dist < bestDist) { bestDist = dist; bestIdx2 = i2; }
        }
        if (bestDist <= TH_HIGH) {
            curMatch[bestIdx2] = i;
            nmatches++;
            if (checkOri) {
                float rot = angleLast[i] - kC[bestIdx2].angle;
                if (rot < 0.0) rot += 360.0f;
                int bin = (int)std::round(rot * factor);
                if (bin == HISTO_LENGTH) bin = 0;
                rotHist[bin].push_back(bestIdx2);
            }
        }
    }
    if (checkOri) {
        int ind1 = -1, ind2 = -1, ind3 = -1;
        three_maxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
        for (int i = 0; i < HISTO_LENGTH; ++i) {
            if (i == ind1 || i == ind2 || i == ind3) continue;
            for (int j : rotHist[i]) { curMatch[j] = -1; nmatches--; }                        // :1672-1675 (no "still set" test)
        }
    }
    return nmatches;
}

// The same function on a stereo-fisheye CURRENT frame (CurrentFrame.Nleft = nC != -1, ORBmatcher.cc:1602-1656 in addition):
// kR = mvKeysRight (descriptor rows nC + i, their own grid), uvR[i] = projection of last-frame point i into the right camera
// (GetRelativePoseTrl() * x3Dc through the camera model: the caller's).  The left half loses its mvuRight gate (Nleft != -1),
// the right half is a best-1 search without bounds test that only runs when the left window was not empty (the `continue`
// of :1552-1553 leaves the iteration).  occupied has nC + nR entries, so has curMatch.  angleLast / octave are per last-frame
// feature whichever camera it came from.
int mo_search_by_projection_last_fisheye(const void* kCv, int nC, const void* kRv, int nR, const uint8_t* dC,
                                         const float* scaleFactors, int minX, int minY, int maxX, int maxY,
                                         const uint8_t* occupied, const uint8_t* valid, const float* uv, const float* uvR,
                                         const float* invz, const int32_t* octave, const float* angleLast, const uint8_t* dMP,
                                         const uint8_t* mpHasObs, int nL, float th, int forward, int backward, int checkOri,
                                         int32_t* curMatch) {
    const KP *kC = (const KP*)kCv, *kR = (const KP*)kRv;
    Grid g(kC, nC, minX, minY, maxX, maxY), gR(kR, nR, minX, minY, maxX, maxY);
    const int N = nC + nR, Nleft = nC;
    for (int j = 0; j < N; ++j) curMatch[j] = -1;
    auto taken = [&](int j) { return (occupied && occupied[j]) || (curMatch[j] >= 0 && mpHasObs[curMatch[j]]); };
    int nmatches = 0;
    std::vector<int> rotHist[HISTO_LENGTH];
    const float factor = 1.0f / HISTO_LENGTH;
    auto vote = [&](int i, float angleCF, int slot) {
        float rot = angleLast[i] - angleCF;
        if (rot < 0.0) rot += 360.0f;
        int bin = (int)std::round(rot * factor);
        if (bin == HISTO_LENGTH) bin = 0;
        rotHist[bin].push_back(slot);
    };
    std::vector<int> vIndices2;
    for (int i = 0; i < nL; ++i) {
        if (!valid[i]) continue;
        if (invz[i] < 0) continue;
        const float u = uv[2 * i], v = uv[2 * i + 1];
        if (u < (float)minX || u > (float)maxX) continue;
        if (v < (float)minY || v > (float)maxY) continue;
        const int nLastOctave = octave[i];
        const float radius = th * scaleFactors[nLastOctave];
        if (forward) g.query(u, v, radius, nLastOctave, -1, vIndices2);
        else if (backward) g.query(u, v, radius, 0, nLastOctave, vIndices2);
        else g.query(u, v, radius, nLastOctave - 1, nLastOctave + 1, vIndices2);
        if (vIndices2.empty()) continue;
        int bestDist = 256, bestIdx2 = -1;
        for (int i2 : vIndices2) {
            if (taken(i2)) continue;
            const int dist = descriptor_distance(dMP + 32 * (size_t)i, dC + 32 * (size_t)i2);
            if (dist < bestDist) { bestDist = dist; bestIdx2 = i2; }
        }
        if (bestDist <= TH_HIGH) {
            curMatch[bestIdx2] = i;
            nmatches++;
            if (checkOri) vote(i, kC[bestIdx2].angle, bestIdx2);
        }
        // right camera (:1602-1656)
        const float ur = uvR[2 * i], vr = uvR[2 * i + 1];
        if (forward) gR.query(ur, vr, radius, nLastOctave, -1, vIndices2);
        else if (backward) gR.query(ur, vr, radius, 0, nLastOctave, vIndices2);
        else gR.query(ur, vr, radius, nLastOctave - 1, nLastOctave + 1, vIndices2);
        bestDist = 256; bestIdx2 = -1;
        for (int i2 : vIndices2) {
            if (taken(i2 + Nleft)) continue;
            const int dist = descriptor_distance(dMP + 32 * (size_t)i, dC + 32 * (size_t)(i2 + Nleft));
            if (dist < bestDist) { bestDist = dist; bestIdx2 = i2; }
        }
        if (bestDist <= TH_HIGH) {
            curMatch[bestIdx2 + Nleft] = i;
            nmatches++;
            if (checkOri) vote(i, kR[bestIdx2].angle, bestIdx2 + Nleft);
        }
    }
    if (checkOri) {
        int ind1 = -1, ind2 = -1, ind3 = -1;
        three_maxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
        for (int i = 0; i < HISTO_LENGTH; ++i) {
            if (i == ind1 || i == ind2 || i == ind3) continue;
            for (int j : rotHist[i]) { curMatch[j] = -1; nmatches--; }
        }
    }
    return nmatches;
}

// ORBmatcher::SearchByProjection(Frame& CurrentFrame, KeyFrame* pKF, sAlreadyFound, th, ORBdist) (ORBmatcher.cc:1685-1794),
// the matcher of Tracking::Relocalization.  Per key-frame feature i: valid[i] = map point present, not bad, not in
// sAlreadyFound (:1703-1704); uv[i] = its projection, dist3D[i] = |x3Dw - Ow| with the window [minDist, maxDist]
// (:1717-1725), level[i] = PredictScale (:1727), angleKF[i] = pKF->mvKeysUn[i].angle, dMP[i].  A current-frame feature that
// holds ANY map point (occupied[j], or assigned earlier in this call) is skipped (:1742-1743).
int mo_search_by_projection_kf(const void* kCv, const uint8_t* dC, int nC, const float* scaleFactors, int minX, int minY,
                               int maxX, int maxY, const uint8_t* occupied, const uint8_t* valid, const float* uv,
                               const float* dist3D, const float* minDist, const float* maxDist, const int32_t* level,
                               const float* angleKF, const uint8_t* dMP, int nK, float th, int orbDist, int checkOri,
                               int32_t* curMatch) {
    const KP* kC = (const KP*)kCv;
    Grid g(kC, nC, minX, minY, maxX, maxY);
    for (int j = 0; j < nC; ++j) curMatch[j] = -1;
    int nmatches = 0;
    std::vector<int> rotHist[HISTO_LENGTH];
    const float factor = 1.0f / HISTO_LENGTH;
    std::vector<int> vIndices2;
    for (int i = 0; i < nK; ++i) {
        if (!valid[i]) continue;
        const float u = uv[2 * i], v = uv[2 * i + 1];
        if (u < (float)minX || u > (float)maxX) continue;
        if (v < (float)minY || v > (float)maxY) continue;
        if (dist3D[i] < minDist[i] || dist3D[i] > maxDist[i]) continue;
        const int nPredictedLevel = level[i];
        const float radius = th * scaleFactors[nPredictedLevel];
        g.query(u, v, radius, nPredictedLevel - 1, nPredictedLevel + 1, vIndices2);
        if (vIndices2.empty()) continue;
        int bestDist = 256, bestIdx2 = -1;
        for (int i2 : vIndices2) {
            if (occupied[i2] || curMatch[i2] >= 0) continue;
            const int dist = descriptor_distance(dMP + 32 * (size_t)i, dC + 32 * (size_t)i2);
            if (dist < bestDist) { bestDist = dist; bestIdx2 = i2; }
        }
        if (bestDist <= orbDist) {
            curMatch[bestIdx2] = i;
            nmatches++;
            if (checkOri) {
                float rot = angleKF[i] - kC[bestIdx2].angle;
                if (rot < 0.0) rot += 360.0f;
                int bin = (int)std::round(rot * factor);
                if (bin == HISTO_LENGTH) bin = 0;
                rotHist[bin].push_back(bestIdx2);
            }
        }
    }
    if (checkOri) {
        int ind1 = -1, ind2 = -1, ind3 = -1;
        three_maxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
        for (int i = 0; i < HISTO_LENGTH; ++i) {
            if (i == ind1 || i == ind2 || i == ind3) continue;
            for (int j : rotHist[i]) { curMatch[j] = -1; nmatches--; }
        }
    }
    return nmatches;
}

// ORBmatcher::Fuse(KeyFrame* pKF, vpMapPoints, th, false) (ORBmatcher.cc:1015-1181), the matcher of
// LocalMapping::SearchInNeighbors, for key frames without a second fisheye camera: everything up to the decision "map point
// i fuses with key-frame feature bestIdx" (:1147).  What follows (Replace / AddObservation / AddMapPoint, :1148-1160) edits
// the map and stays with the caller; it does not feed back into later iterations of the search.
// Per map point i: valid[i] = present, not bad, not already in pKF, depth >= 0, viewing angle test passed (:1045-1100, the
// caller's pose / camera / normal); uv[i] = projection, ur[i] = uv.x - bf * invz (:1074), dist3D[i] vs [minDist, maxDist],
// level[i] = PredictScale.  Key frame: mvKeysUn, descriptors, mvuRight (< 0: mono feature), mvInvLevelSigma2, grid.
// bestIdx[i] = key-frame feature (bestDist <= TH_LOW) or -1; returns nFused.
int mo_fuse_search(const void* kKv, const uint8_t* dK, int nK, const float* scaleFactors, const float* invLevelSigma2, int minX,
                   int minY, int maxX, int maxY, const float* uRight, const uint8_t* valid, const float* uv, const float* ur,
                   const float* dist3D, const float* minDist, const float* maxDist, const int32_t* level, const uint8_t* dMP,
                   int nMP, float th, int thDist, int32_t* bestIdxOut, int32_t* bestDistOut) {
    const KP* kK = (const KP*)kKv;
    Grid g(kK, nK, minX, minY, maxX, maxY);
    std::vector<int> vIndices;
    int nFused = 0;
    for (int i = 0; i < nMP; ++i) {
        bestIdxOut[i] = -1; bestDistOut[i] = 256;
        if (!valid[i]) continue;
        const float u = uv[2 * i], v = uv[2 * i + 1];
        if (!(u >= (float)minX && u < (float)maxX && v >= (float)minY && v < (float)maxY)) continue;   // KeyFrame::IsInImage
        if (dist3D[i] < minDist[i] || dist3D[i] > maxDist[i]) continue;
        const int nPredictedLevel = level[i];
        const float radius = th * scaleFactors[nPredictedLevel];
        g.query(u, v, radius, -1, -1, vIndices);                       // KeyFrame::GetFeaturesInArea: no level filter
        if (vIndices.empty()) continue;
        int bestDist = 256, bestIdx = -1;
        for (int idx : vIndices) {
            const KP& kp = kK[idx];
            const int kpLevel = kp.octave;
            if (kpLevel < nPredictedLevel - 1 || kpLevel > nPredictedLevel) continue;
            if (uRight[idx] >= 0) {                                    // stereo feature: 3-dof reprojection error (:1116-1128)
                const float ex = u - kp.x, ey = v - kp.y, er = ur[i] - uRight[idx];
                const float e2 = ex * ex + ey * ey + er * er;
                if (e2 * invLevelSigma2[kpLevel] > 7.8) continue;
            } else {
                const float ex = u - kp.x, ey = v - kp.y;
                const float e2 = ex * ex + ey * ey;
                if (e2 * invLevelSigma2[kpLevel] > 5.99) continue;
            }
            const int dist = descriptor_distance(dMP + 32 * (size_t)i, dK + 32 * (size_t)idx);
            if (dist < bestDist) { bestDist = dist; bestIdx = idx; }
        }
        bestDistOut[i] = bestDist;
        if (bestDist <= thDist) { bestIdxOut[i] = bestIdx; nFused++; }        // TH_LOW in Fuse, TH_HIGH in SearchBySim3
    }
    return nFused;
}

// ORBmatcher::SearchByProjection(KeyFrame* pKF, Sim3 Scw, vpPoints, vpMatched, th, ratioHamming) (ORBmatcher.cc:372-471) and
// its overload that also records the points' key frames (:473-580) -- the matchers of LoopClosing (same search).  Per
// candidate map point i: valid[i] = not bad, not in the initial vpMatched, depth >= 0, viewing-angle test passed; uv[i],
// dist[i] vs [minDist, maxDist], level[i] = PredictScale.  occupied[j] = vpMatched[j] != NULL on entry; a feature matched
// earlier in this call is skipped as well (:443-444).  kfMatch[j] = map point stored in vpMatched[j] by this call, or -1.
int mo_search_by_projection_sim3(const void* kKv, const uint8_t* dK, int nK, const float* scaleFactors, int minX, int minY, int maxX,
                                 int maxY, const uint8_t* occupied, const uint8_t* valid, const float* uv, const float* dist3D,
                                 const float* minDist, const float* maxDist, const int32_t* level, const uint8_t* dMP, int nMP,
                                 int th, float ratioHamming, int32_t* kfMatch) {
    const KP* kK = (const KP*)kKv;
    Grid g(kK, nK, minX, minY, maxX, maxY);
    for (int j = 0; j < nK; ++j) kfMatch[j] = -1;
    std::vector<int> vIndices;
    int nmatches = 0;
    for (int i = 0; i < nMP; ++i) {
        if (!valid[i]) continue;
        const float u = uv[2 * i], v = uv[2 * i + 1];
        if (!(u >= (float)minX && u < (float)maxX && v >= (float)minY && v < (float)maxY)) continue;
        if (dist3D[i] < minDist[i] || dist3D[i] > maxDist[i]) continue;
        const int nPredictedLevel = level[i];
        const float radius = th * scaleFactors[nPredictedLevel];
        g.query(u, v, radius, -1, -1, vIndices);
        if (vIndices.empty()) continue;
        int bestDist = 256, bestIdx = -1;
        for (int idx : vIndices) {
            if (occupied[idx] || kfMatch[idx] >= 0) continue;
            const int kpLevel = kK[idx].octave;
            if (kpLevel < nPredictedLevel - 1 || kpLevel > nPredictedLevel) continue;
            const int dist = descriptor_distance(dMP + 32 * (size_t)i, dK + 32 * (size_t)idx);
            if (dist < bestDist) { bestDist = dist; bestIdx = idx; }
        }
        if (bestDist <= TH_LOW * ratioHamming) { kfMatch[bestIdx] = i; nmatches++; }
    }
    return nmatches;
}

// ORBmatcher::SearchBySim3(pKF1, pKF2, vpMatches12, S12, th) (ORBmatcher.cc:1293-1497): the map points of each key frame are
// searched in the other one (window around the Sim3 projection, [level - 1, level], best distance <= TH_HIGH) and a pair is
// kept when both directions agree (:1483-1494).  Each direction is the Fuse search without gates and with TH_HIGH.
// valid1[i1] = pKF1's feature has a map point that is not bad and not matched yet, depth >= 0 in camera 2; uv12 / dist12 /
// level12 = its projection into image 2; likewise valid2 / uv21 / ... ; match12[i1] = feature of key frame 2 or -1.
int mo_search_by_sim3(const void* k1v, const uint8_t* d1, int n1, const void* k2v, const uint8_t* d2, int n2,
                      const float* scaleFactors, int minX, int minY, int maxX, int maxY, const uint8_t* valid1, const float* uv12,
                      const float* dist12, const float* min1, const float* max1, const int32_t* level12, const uint8_t* valid2,
                      const float* uv21, const float* dist21, const float* min2, const float* max2, const int32_t* level21,
                      float th, int32_t* match12) {
    std::vector<float> zeroSigma(64, 0.0f), noRight1(n1 > 0 ? n1 : 1, -1.0f), noRight2(n2 > 0 ? n2 : 1, -1.0f);
    std::vector<float> ur1(n1 > 0 ? n1 : 1, 0.0f), ur2(n2 > 0 ? n2 : 1, 0.0f);
    std::vector<int32_t> vnMatch1(n1 > 0 ? n1 : 1), vnMatch2(n2 > 0 ? n2 : 1), dd1(n1 > 0 ? n1 : 1), dd2(n2 > 0 ? n2 : 1);
    mo_fuse_search(k2v, d2, n2, scaleFactors, zeroSigma.data(), minX, minY, maxX, maxY, noRight2.data(), valid1, uv12, ur1.data(),
                   dist12, min1, max1, level12, d1, n1, th, TH_HIGH, vnMatch1.data(), dd1.data());
    mo_fuse_search(k1v, d1, n1, scaleFactors, zeroSigma.data(), minX, minY, maxX, maxY, noRight1.data(), valid2, uv21, ur2.data(),
                   dist21, min2, max2, level21, d2, n2, th, TH_HIGH, vnMatch2.data(), dd2.data());
    int nFound = 0;
    for (int i1 = 0; i1 < n1; ++i1) {
        match12[i1] = -1;
        const int idx2 = vnMatch1[i1];
        if (idx2 >= 0 && vnMatch2[idx2] == i1) { match12[i1] = idx2; nFound++; }
    }
    return nFound;
}

// CloudMerging.cc:503-551 for one matched key-frame pair: per key point of key frame 1 the nearest key point of key
// frame 2 in PIXEL distance (< tol, both with a map point) among GetFeaturesInArea(u, v, tol).  match12[i] = j or -1.
int mo_associate_pixels(const void* k1v, const uint8_t* valid1, int n1, const void* k2v, const uint8_t* valid2, int n2,
                        int minX, int minY, int maxX, int maxY, float tol, int32_t* match12) {
    const KP* k1 = (const KP*)k1v; const KP* k2 = (const KP*)k2v;
    Grid g(k2, n2, minX, minY, maxX, maxY);
    std::vector<int> v;
    int matchNum = 0;
    for (int i = 0; i < n1; ++i) {
        match12[i] = -1;
        const float u = k1[i].x, vv = k1[i].y;
        g.query(u, vv, tol, -1, -1, v);
        if (v.empty()) continue;
        float best_dist = tol;
        int best_idx = -1;
        for (int j : v) {
            const float delta = (float)std::sqrt(std::pow(u - k2[j].x, 2) + std::pow(vv - k2[j].y, 2));
            if (delta < best_dist && valid1[i] && valid2[j]) { best_idx = j; best_dist = delta; }
        }
        if (best_idx == -1) continue;
        ++matchNum;
        match12[i] = best_idx;
    }
    return matchNum;
}

}  // extern "C"
