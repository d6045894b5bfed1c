// ORACLE -- TEST INFRASTRUCTURE ONLY (used by tests/, __graft_entry__.smoke() and bench.py's CPU arm; never by the
// product path).
//
// CPU restatement of the bag-of-words side of the matching path:
//   * DBoW2 vocabulary tree descent -- TemplatedVocabulary::transform, R/Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h
//     :1218-1258 (per feature) and :1128-1200 (BowVector / FeatureVector assembly = Frame::ComputeBoW),
//     tree built like loadFromTextFile (:1338-1421), distance = FORB::distance (FORB.cpp:81-101);
//   * ORBmatcher::SearchByBoW(KeyFrame*, Frame&, ...) -- R/lib_src/ORBmatcher.cc:198-370, the Nleft == -1 branch, and
//     SearchByBoW(KeyFrame*, KeyFrame*, ...) -- :682-804, with ComputeThreeMaxima (:1795-1828);
//   * ORBmatcher::SearchForTriangulation -- :806-1013, key frames without a second camera.
// Parity: PINNED.  The tree descent and the vector assembly are checked against the UNMODIFIED reference DBoW2
// compiled over oracle/cvstub (oracle/_ref/librefbow.so, tests/test_bow_oracle.py) and against the frozen vectors
// tests/golden/bow_kats.npz.  SearchByBoW (both overloads) and SearchForTriangulation are checked against the reference's
// own functions compiled over class stand-ins (oracle/_ref/librefframe.so, tests/test_ref_frame_pin.py).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <map>
#include <vector>

namespace {

struct Node {
    std::vector<int> children;
    int parent = 0, word = 0;
    double weight = 0.0;
    uint8_t desc[32];
};
struct Vocab {
    int k = 0, L = 0, scoring = 0, weighting = 0;
    std::vector<Node> nodes;       // node 0 = root
    int nwords = 0;
};

// FORB.cpp:81-101 (bit-parallel popcount per 32-bit word)
int forb_distance(const uint8_t* a, const uint8_t* b) {
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

// TemplatedVocabulary.h:1218-1258
void transform_one(const Vocab& v, const uint8_t* f, int levelsup, int& word, double& weight, int& nid) {
    const int nid_level = v.L - levelsup;
    nid = 0;                                           // root when nid_level <= 0 (:1228)
    int final_id = 0, current_level = 0;
    do {
        ++current_level;
        const std::vector<int>& nodes = v.nodes[final_id].children;
        final_id = nodes[0];
        double best_d = forb_distance(f, v.nodes[final_id].desc);
        for (size_t i = 1; i < nodes.size(); ++i) {
            const double d = forb_distance(f, v.nodes[nodes[i]].desc);
            if (d < best_d) { best_d = d; final_id = nodes[i]; }
        }
        if (current_level == nid_level) nid = final_id;
    } while (!v.nodes[final_id].children.empty());
    word = v.nodes[final_id].word;
    weight = v.nodes[final_id].weight;
}

}  // namespace

extern "C" {

// nodes 1..nnodes-1 in id order: parent id, leaf flag, 32-byte descriptor, weight  (loadFromTextFile :1377-1417)
void* bow_oracle_create(int k, int L, int scoring, int weighting, int nnodes, const int32_t* parent,
                        const uint8_t* is_leaf, const uint8_t* desc, const double* weight) {
    Vocab* v = new Vocab();
    v->k = k; v->L = L; v->scoring = scoring; v->weighting = weighting;
    v->nodes.resize(nnodes);
    for (int nid = 1; nid < nnodes; ++nid) {
        Node& n = v->nodes[nid];
        n.parent = parent[nid];
        v->nodes[parent[nid]].children.push_back(nid);
        std::memcpy(n.desc, desc + 32 * (size_t)nid, 32);
        n.weight = weight[nid];
        if (is_leaf[nid]) n.word = v->nwords++;
    }
    return v;
}
void bow_oracle_free(void* p) { delete static_cast<Vocab*>(p); }

int bow_oracle_transform(void* p, const uint8_t* desc, int n, int levelsup, int32_t* word, double* weight,
                         int32_t* node) {
    const Vocab& v = *static_cast<Vocab*>(p);
    for (int i = 0; i < n; ++i) {
        int w, nid; double val;
        transform_one(v, desc + 32 * (size_t)i, levelsup, w, val, nid);
        word[i] = w; weight[i] = val; node[i] = nid;
    }
    return 0;
}

// TemplatedVocabulary.h:1128-1200 flattened like oracle/ref_bow_shim.cpp's refbow_vectors
int bow_oracle_vectors(void* p, const uint8_t* desc, int n, int levelsup, int32_t* bow_ids, double* bow_vals,
                       int32_t* fv_nodes, int32_t* fv_off, int32_t* fv_idx, int* nfv) {
    const Vocab& v = *static_cast<Vocab*>(p);
    std::map<unsigned, double> bv;
    std::map<unsigned, std::vector<unsigned>> fv;
    // ScoringObject.h:73-89: every scoring but DOT_PRODUCT (5) normalises; L2_NORM (1) with L2, the others with L1
    const bool must = v.scoring != 5;
    const bool l2 = v.scoring == 1;
    const bool tf = v.weighting == 1 || v.weighting == 0;     // BowVector.h: TF_IDF = 0, TF = 1, IDF = 2, BINARY = 3
    for (int i = 0; i < n; ++i) {
        int w, nid; double val;
        transform_one(v, desc + 32 * (size_t)i, levelsup, w, val, nid);
        if (val > 0) {
            if (tf) bv[(unsigned)w] += val;                               // addWeight (BowVector.cpp:34-46)
            else if (!bv.count((unsigned)w)) bv[(unsigned)w] = val;       // addIfNotExist (:50-58)
            fv[(unsigned)nid].push_back((unsigned)i);                     // addFeature (FeatureVector.cpp)
        }
    }
    if (tf && !bv.empty() && !must) {
        const double nd = (double)bv.size();
        for (auto& e : bv) e.second /= nd;
    }
    if (must) {                                                           // BowVector::normalize (:62-84)
        double norm = 0.0;
        if (!l2) { for (auto& e : bv) norm += std::fabs(e.second); }
        else { for (auto& e : bv) norm += e.second * e.second; norm = std::sqrt(norm); }
        if (norm > 0.0) for (auto& e : bv) e.second /= norm;
    }
    int k = 0;
    for (auto& e : bv) { bow_ids[k] = (int32_t)e.first; bow_vals[k] = e.second; ++k; }
    int m = 0, o = 0;
    for (auto& e : fv) {
        fv_nodes[m] = (int32_t)e.first; fv_off[m] = o;
        for (unsigned idx : e.second) fv_idx[o++] = (int32_t)idx;
        ++m;
    }
    fv_off[m] = o;
    *nfv = m;
    return k;
}

// ORBmatcher::ComputeThreeMaxima, R/lib_src/ORBmatcher.cc:1787-1828
static void three_maxima(const std::vector<int>* histo, int L, int& ind1, int& ind2, int& ind3) {
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

// ORBmatcher::SearchByBoW(KeyFrame*, Frame&, vpMapPointMatches), R/lib_src/ORBmatcher.cc:198-370, F.Nleft == -1.
// Feature vectors are flattened (sorted node ids, offsets, feature indices).  kf_valid[i] != 0 <=> the keyframe
// feature has a map point that is not bad (:227-233).  match_f[j] = keyframe feature matched to frame feature j or -1
// (the reference stores the MapPoint*).  Returns nmatches.
// n_left = F.Nleft: -1 for a mono / rectified frame; otherwise features [0, n_left) belong to the left and [n_left, nf)
// to the right fisheye camera, each side keeps its own best / second best and the right one is accepted without a ratio
// test (":316 ... || true"), only when the left one passed TH_LOW (:290-340).
int bow_oracle_search_by_bow_nleft(const uint8_t* desc_kf, const float* angle_kf, const uint8_t* kf_valid,
                             const int32_t* kf_nodes, const int32_t* kf_off, const int32_t* kf_idx, int kf_nnodes,
                             const uint8_t* desc_f, const float* angle_f, int nf,
                             const int32_t* f_nodes, const int32_t* f_off, const int32_t* f_idx, int f_nnodes,
                             float nnratio, int check_ori, int th_low, int n_left, int32_t* match_f) {
    const int HISTO_LENGTH = 30;
    for (int j = 0; j < nf; ++j) match_f[j] = -1;
    int nmatches = 0;
    std::vector<int> rotHist[HISTO_LENGTH];
    const float factor = 1.0f / HISTO_LENGTH;
    int a = 0, b = 0;
    while (a < kf_nnodes && b < f_nnodes) {
        if (kf_nodes[a] == f_nodes[b]) {
            for (int iKF = kf_off[a]; iKF < kf_off[a + 1]; ++iKF) {
                const int realIdxKF = kf_idx[iKF];
                if (!kf_valid[realIdxKF]) continue;
                const uint8_t* dKF = desc_kf + 32 * (size_t)realIdxKF;
                int bestDist1 = 256, bestIdxF = -1, bestDist2 = 256;
                int bestDist1R = 256, bestIdxFR = -1, bestDist2R = 256;
                for (int iF = f_off[b]; iF < f_off[b + 1]; ++iF) {
                    const int realIdxF = f_idx[iF];
                    if (match_f[realIdxF] >= 0) continue;                         // :249
                    const int dist = forb_distance(dKF, desc_f + 32 * (size_t)realIdxF);
                    if (n_left == -1 || realIdxF < n_left) {
                        if (dist < bestDist1) { bestDist2 = bestDist1; bestDist1 = dist; bestIdxF = realIdxF; }
                        else if (dist < bestDist2) { bestDist2 = dist; }
                    } else {
                        if (dist < bestDist1R) { bestDist2R = bestDist1R; bestDist1R = dist; bestIdxFR = realIdxF; }
                        else if (dist < bestDist2R) { bestDist2R = dist; }
                    }
                }
                auto accept = [&](int idxF) {
                    match_f[idxF] = realIdxKF;
                    if (check_ori) {
                        float rot = angle_kf[realIdxKF] - angle_f[idxF];
                        if (rot < 0.0) rot += 360.0f;
                        int bin = (int)std::round(rot * factor);
                        if (bin == HISTO_LENGTH) bin = 0;
                        rotHist[bin].push_back(idxF);
                    }
                    nmatches++;
                };
                if (bestDist1 <= th_low) {
                    if ((float)bestDist1 < nnratio * (float)bestDist2) accept(bestIdxF);
                    if (bestDist1R <= th_low) accept(bestIdxFR);                  // ":316 ... || true": no ratio test on the right
                }
            }
            ++a; ++b;
        } else if (kf_nodes[a] < f_nodes[b]) {
            a = (int)(std::lower_bound(kf_nodes + a, kf_nodes + kf_nnodes, f_nodes[b]) - kf_nodes);
        } else {
            b = (int)(std::lower_bound(f_nodes + b, f_nodes + f_nnodes, kf_nodes[a]) - f_nodes);
        }
    }
    if (check_ori) {
        int ind1 = -1, ind2 = -1, ind3 = -1;
        three_maxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
        for (int i = 0; i < HISTO_LENGTH; i++) {
            if (i == ind1 || i == ind2 || i == ind3) continue;
            for (size_t j = 0; j < rotHist[i].size(); j++) { match_f[rotHist[i][j]] = -1; nmatches--; }
        }
    }
    return nmatches;
}

int bow_oracle_search_by_bow(const uint8_t* desc_kf, const float* angle_kf, const uint8_t* kf_valid,
                             const int32_t* kf_nodes, const int32_t* kf_off, const int32_t* kf_idx, int kf_nnodes,
                             const uint8_t* desc_f, const float* angle_f, int nf,
                             const int32_t* f_nodes, const int32_t* f_off, const int32_t* f_idx, int f_nnodes,
                             float nnratio, int check_ori, int th_low, int32_t* match_f) {
    return bow_oracle_search_by_bow_nleft(desc_kf, angle_kf, kf_valid, kf_nodes, kf_off, kf_idx, kf_nnodes, desc_f, angle_f, nf,
                                          f_nodes, f_off, f_idx, f_nnodes, nnratio, check_ori, th_low, -1, match_f);
}

// ORBmatcher::SearchByBoW(KeyFrame*, KeyFrame*, vpMatches12), R/lib_src/ORBmatcher.cc:682-804 (NLeft == -1 for both).
// valid1 / valid2: the feature has a map point that is not bad.  match12[i1] = feature of keyframe 2 whose map point the
// reference stores in vpMatches12[i1], or -1.  Differences to the KeyFrame-Frame overload: strict '<' against TH_LOW
// (:756), keyframe-2 features need a valid map point (:735-739), and a keyframe-2 feature stays "matched" (:759) even
// when the rotation histogram later drops the match.
int bow_oracle_search_by_bow_kf(const uint8_t* desc1, const float* angle1, const uint8_t* valid1, int n1,
                                const int32_t* nodes1, const int32_t* off1, const int32_t* idx1, int nnodes1,
                                const uint8_t* desc2, const float* angle2, const uint8_t* valid2, int n2,
                                const int32_t* nodes2, const int32_t* off2, const int32_t* idx2, int nnodes2,
                                float nnratio, int check_ori, int th_low, int32_t* match12) {
    const int HISTO_LENGTH = 30;
    for (int i = 0; i < n1; ++i) match12[i] = -1;
    std::vector<char> matched2(n2 > 0 ? n2 : 1, 0);
    std::vector<int> rotHist[HISTO_LENGTH];
    const float factor = 1.0f / HISTO_LENGTH;
    int nmatches = 0, a = 0, b = 0;
    while (a < nnodes1 && b < nnodes2) {
        if (nodes1[a] == nodes2[b]) {
            for (int i1 = off1[a]; i1 < off1[a + 1]; ++i1) {
                const int id1 = idx1[i1];
                if (!valid1[id1]) continue;
                const uint8_t* d1 = desc1 + 32 * (size_t)id1;
                int bestDist1 = 256, bestIdx2 = -1, bestDist2 = 256;
                for (int i2 = off2[b]; i2 < off2[b + 1]; ++i2) {
                    const int id2 = idx2[i2];
                    if (matched2[id2] || !valid2[id2]) continue;
                    const int dist = forb_distance(d1, desc2 + 32 * (size_t)id2);
                    if (dist < bestDist1) { bestDist2 = bestDist1; bestDist1 = dist; bestIdx2 = id2; }
                    else if (dist < bestDist2) { bestDist2 = dist; }
                }
                if (bestDist1 < th_low) {
                    if ((float)bestDist1 < nnratio * (float)bestDist2) {
                        match12[id1] = bestIdx2;
                        matched2[bestIdx2] = 1;
                        if (check_ori) {
                            float rot = angle1[id1] - angle2[bestIdx2];
                            if (rot < 0.0) rot += 360.0f;
                            int bin = (int)std::round(rot * factor);
                            if (bin == HISTO_LENGTH) bin = 0;
                            rotHist[bin].push_back(id1);
                        }
                        nmatches++;
                    }
                }
            }
            ++a; ++b;
        } else if (nodes1[a] < nodes2[b]) {
            a = (int)(std::lower_bound(nodes1 + a, nodes1 + nnodes1, nodes2[b]) - nodes1);
        } else {
            b = (int)(std::lower_bound(nodes2 + b, nodes2 + nnodes2, nodes1[a]) - nodes2);
        }
    }
    if (check_ori) {
        int ind1 = -1, ind2 = -1, ind3 = -1;
        three_maxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
        for (int i = 0; i < HISTO_LENGTH; i++) {
            if (i == ind1 || i == ind2 || i == ind3) continue;
            for (size_t j = 0; j < rotHist[i].size(); j++) { match12[rotHist[i][j]] = -1; nmatches--; }
        }
    }
    return nmatches;
}

// MapPoint::ComputeDistinctiveDescriptors, R/lib_src/MapPoint.cc:392-426, for one map point with N observed descriptors.
// (Restated; MapPoint.cc needs KeyFrame / Map and cannot be compiled here.)  Returns BestIdx, *median = BestMedian.
int bow_oracle_distinctive(const uint8_t* desc, int N, int* median) {
    if (N <= 0) { *median = -1; return -1; }
    std::vector<std::vector<float>> D(N, std::vector<float>(N));
    for (int i = 0; i < N; i++) {
        D[i][i] = 0;
        for (int j = i + 1; j < N; j++) {
            const int distij = forb_distance(desc + 32 * (size_t)i, desc + 32 * (size_t)j);   // == ORBmatcher::DescriptorDistance
            D[i][j] = (float)distij;
            D[j][i] = (float)distij;
        }
    }
    int BestMedian = 2147483647, BestIdx = 0;
    for (int i = 0; i < N; i++) {
        std::vector<int> vDists(D[i].begin(), D[i].end());
        std::sort(vDists.begin(), vDists.end());
        const int med = vDists[(size_t)(0.5 * (N - 1))];
        if (med < BestMedian) { BestMedian = med; BestIdx = i; }
    }
    *median = BestMedian;
    return BestIdx;
}

// ORBmatcher::SearchForTriangulation(pKF1, pKF2, vMatchedPairs, bOnlyStereo, bCoarse), R/lib_src/ORBmatcher.cc:806-1013, the
// matcher of LocalMapping::CreateNewMapPoints, for key frames without a second fisheye camera (mpCamera2 == NULL).
// has_mp*[i]: the feature already has a map point (:862-867, :889-893); stereo*[i]: mvuRight[i] >= 0; (x2, y2, octave2) of
// the key points of key frame 2; (epx, epy) = the epipole in image 2 (:815-819, the caller's poses and camera);
// epi_ok[i1 * n2 + i2] = what pCamera1->epipolarConstrain(...) returns for the pair (:957; evaluated by the reference only
// for pairs that survive the distance tests -- it is a pure function of the pair, so a table is equivalent).
// match12[i1] = feature of key frame 2 or -1.  (vbMatched2 is never set by the reference, :852 / :890: kept that way.)
int bow_oracle_search_for_triangulation(const uint8_t* desc1, const float* angle1, const uint8_t* has_mp1, const uint8_t* stereo1,
                                        int n1, const int32_t* nodes1, const int32_t* off1, const int32_t* idx1, int nnodes1,
                                        const uint8_t* desc2, const float* angle2, const uint8_t* has_mp2, const uint8_t* stereo2,
                                        const float* x2, const float* y2, const int32_t* octave2, int n2, const int32_t* nodes2,
                                        const int32_t* off2, const int32_t* idx2, int nnodes2, const float* scale_factors2,
                                        float epx, float epy, int only_stereo, int coarse, const uint8_t* epi_ok, int check_ori,
                                        int32_t* match12) {
    const int HISTO_LENGTH = 30, TH_LOW = 50;
    for (int i = 0; i < n1; ++i) match12[i] = -1;
    std::vector<int> rotHist[HISTO_LENGTH];
    const float factor = 1.0f / HISTO_LENGTH;
    int nmatches = 0, a = 0, b = 0;
    while (a < nnodes1 && b < nnodes2) {
        if (nodes1[a] == nodes2[b]) {
            for (int i1 = off1[a]; i1 < off1[a + 1]; ++i1) {
                const int id1 = idx1[i1];
                if (has_mp1[id1]) continue;
                const bool bStereo1 = stereo1[id1] != 0;
                if (only_stereo && !bStereo1) continue;
                const uint8_t* d1 = desc1 + 32 * (size_t)id1;
                int bestDist = TH_LOW, bestIdx2 = -1;
                for (int i2 = off2[b]; i2 < off2[b + 1]; ++i2) {
                    const int id2 = idx2[i2];
                    if (has_mp2[id2]) continue;
                    const bool bStereo2 = stereo2[id2] != 0;
                    if (only_stereo && !bStereo2) continue;
                    const int dist = forb_distance(d1, desc2 + 32 * (size_t)id2);
                    if (dist > TH_LOW || dist > bestDist) continue;
                    if (!bStereo1 && !bStereo2) {
                        const float distex = epx - x2[id2], distey = epy - y2[id2];
                        if (distex * distex + distey * distey < 100 * scale_factors2[octave2[id2]]) continue;
                    }
                    if (coarse || epi_ok[(size_t)id1 * n2 + id2]) { bestIdx2 = id2; bestDist = dist; }
                }
                if (bestIdx2 >= 0) {
                    match12[id1] = bestIdx2;
                    nmatches++;
                    if (check_ori) {
                        float rot = angle1[id1] - angle2[bestIdx2];
                        if (rot < 0.0) rot += 360.0f;
                        int bin = (int)std::round(rot * factor);
                        if (bin == HISTO_LENGTH) bin = 0;
                        rotHist[bin].push_back(id1);
                    }
                }
            }
            ++a; ++b;
        } else if (nodes1[a] < nodes2[b]) {
            ++a;                                   // lower_bound on a sorted map = advance to the first id >= the other's
        } else {
            ++b;
        }
    }
    if (check_ori) {
        int ind1 = -1, ind2 = -1, ind3 = -1;
        three_maxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
        for (int i = 0; i < HISTO_LENGTH; ++i) {
            if (i == ind1 || i == ind2 || i == ind3) continue;
            for (int id1 : rotHist[i]) { match12[id1] = -1; nmatches--; }
        }
    }
    return nmatches;
}

}  // extern "C"
