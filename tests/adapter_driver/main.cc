// Test driver: runs the reference-signature C++ classes (rumi_slam_b200/adapter) on raw image files and dumps what a
// Frame constructor would receive, so that tests/test_adapter_gpu.py can compare it with the oracle.
//   adapter_driver <out.bin> <w> <h> <nfeatures> <lap0> <lap1> <left.raw> [<right.raw> <mbf> <mb>]
// out.bin: int32 mono, nkp, then nkp x 28 B keypoints, nkp x 32 B descriptors, int32 nlevels, per level int32 w, h and
// the level's pixels (mvImagePyramid); with a right image: the same block for the right frame, then int32 nmatches,
// nL floats mvuRight, nL floats mvDepth (Frame::ComputeStereoMatches through ORBmatcherAccel), int32 nmatches + nL int32
// vnMatches12 of SearchForInitialization(left, right, window 100), int32 total, int32 n1, n1 int32 of AssociateSubmap.
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <stdexcept>
#include <thread>
#include <vector>

#include "ORBextractor.h"
#include "ORBmatcher_accel.h"

static cv::Mat load(const char* path, int w, int h) {
    cv::Mat m(h, w, CV_8UC1);
    FILE* f = std::fopen(path, "rb");
    if (!f) throw std::runtime_error("cannot open image");
    for (int y = 0; y < h; ++y)
        if (std::fread(m.ptr(y), 1, (size_t)w, f) != (size_t)w) throw std::runtime_error("short image file");
    std::fclose(f);
    return m;
}

struct Result {
    int mono = 0;
    std::vector<cv::KeyPoint> keys;
    cv::Mat desc;
};

static void dump(FILE* f, ORB_SLAM3::ORBextractor& ex, const Result& r) {
    const int nkp = (int)r.keys.size();
    std::fwrite(&r.mono, 4, 1, f);
    std::fwrite(&nkp, 4, 1, f);
    std::fwrite(r.keys.data(), 28, (size_t)nkp, f);
    for (int i = 0; i < nkp; ++i) std::fwrite(r.desc.ptr(i), 1, 32, f);
    const int nl = ex.GetLevels();
    std::fwrite(&nl, 4, 1, f);
    for (int l = 0; l < nl; ++l) {
        const cv::Mat& m = ex.mvImagePyramid[l];
        std::fwrite(&m.cols, 4, 1, f);
        std::fwrite(&m.rows, 4, 1, f);
        for (int y = 0; y < m.rows; ++y) std::fwrite(m.ptr(y), 1, (size_t)m.cols, f);
    }
}

int main(int argc, char** argv) {
    if (argc < 8) return 2;
    try {
        const int w = std::atoi(argv[2]), h = std::atoi(argv[3]), nf = std::atoi(argv[4]);
        std::vector<int> lap = {std::atoi(argv[5]), std::atoi(argv[6])};
        const bool stereo = argc >= 11;
        cv::Mat left = load(argv[7], w, h), right;
        if (stereo) right = load(argv[8], w, h);
        ORB_SLAM3::ORBextractor exL(nf, 1.2f, 8, 20, 7), exR(nf, 1.2f, 8, 20, 7);
        Result L, R;
        // Frame.cc:116-119: the two extractors run in two threads
        std::thread tl([&] { L.mono = exL(left, cv::Mat(), L.keys, L.desc, lap); });
        std::thread tr([&] { if (stereo) R.mono = exR(right, cv::Mat(), R.keys, R.desc, lap); });
        tl.join(); tr.join();
        FILE* f = std::fopen(argv[1], "wb");
        if (!f) return 3;
        dump(f, exL, L);
        if (stereo) {
            dump(f, exR, R);
            ORB_SLAM3::ORBmatcherAccel m;
            std::vector<float> uR, depth;
            const int n = m.ComputeStereoMatches(exL.Handle(), exR.Handle(), L.keys, L.desc, R.keys, R.desc,
                                                 (float)std::atof(argv[9]), (float)std::atof(argv[10]), uR, depth);
            std::fwrite(&n, 4, 1, f);
            std::fwrite(uR.data(), 4, uR.size(), f);
            std::fwrite(depth.data(), 4, depth.size(), f);
            // candidate-list matcher through the C++ adapter: SearchForInitialization(left -> right, window 100)
            ORB_SLAM3::ORBmatcherAccel mi(0.9f);
            ORB_SLAM3::FrameGridAccel grid(R.keys, 0.f, 0.f, (float)w, (float)h);
            std::vector<cv::Point2f> prev;
            for (const cv::KeyPoint& k : L.keys) prev.push_back(k.pt);
            std::vector<int> m12;
            const int ni = mi.SearchForInitialization(L.keys, L.desc, R.keys, R.desc, grid, prev, m12, 100, true);
            std::fwrite(&ni, 4, 1, f);
            std::fwrite(m12.data(), 4, m12.size(), f);
            // descriptor-based association of ONE key-frame pair (level-0 key points, every one valid)
            std::vector<std::vector<cv::KeyPoint>> k1(1), k2(1);
            for (const cv::KeyPoint& k : L.keys) if (k.octave == 0) k1[0].push_back(k);
            for (const cv::KeyPoint& k : R.keys) if (k.octave == 0) k2[0].push_back(k);
            std::vector<std::vector<uint8_t>> v1(1, std::vector<uint8_t>(k1[0].size(), 1)), v2(1, std::vector<uint8_t>(k2[0].size(), 1));
            std::vector<std::vector<int>> a12;
            ORB_SLAM3::ORBmatcherAccel ma(0.75f);
            const int na = ma.AssociateSubmap(exL.Handle(), std::vector<cv::Mat>(1, left), k1, v1, std::vector<cv::Mat>(1, right), k2,
                                              v2, a12);
            const int n1 = (int)k1[0].size();
            std::fwrite(&na, 4, 1, f);
            std::fwrite(&n1, 4, 1, f);
            std::fwrite(a12[0].data(), 4, a12[0].size(), f);
            // the projection matchers of TrackWithMotionModel / Relocalization: last frame / key frame = left image, current
            // frame = right image; deterministic synthetic projections (the test rebuilds the same numbers)
            const int nL = (int)L.keys.size();
            std::vector<uint8_t> valid(nL), hasObs(nL), noneU8;
            std::vector<cv::Point2f> uv(nL);
            std::vector<float> invz(nL), angle(nL), d3(nL), dmin(nL, 0.5f), dmax(nL, 12.0f), noneF, sf = exL.GetScaleFactors();
            std::vector<int> octave(nL);
            for (int i = 0; i < nL; ++i) {
                valid[i] = i % 7 != 0; hasObs[i] = i % 5 != 0;
                uv[i] = cv::Point2f(L.keys[i].pt.x - 2.0f, L.keys[i].pt.y + 1.0f);
                d3[i] = 1.0f + (float)(i % 13);
                invz[i] = 1.0f / d3[i];
                angle[i] = L.keys[i].angle; octave[i] = L.keys[i].octave;
            }
            std::vector<int> cmLast, cmKF;
            const int nl = mi.SearchByProjectionLastFrame(R.keys, R.desc, grid, sf, noneF, noneU8, 0.f, valid, uv, invz, octave, angle,
                                                          L.desc, hasObs, 15.0f, false, false, true, cmLast);
            const int nk = mi.SearchByProjectionKeyFrame(R.keys, R.desc, grid, sf, noneU8, valid, uv, d3, dmin, dmax, octave, angle,
                                                         L.desc, 10.0f, 100, true, cmKF);
            const int nR = (int)R.keys.size();
            std::fwrite(&nl, 4, 1, f); std::fwrite(&nk, 4, 1, f); std::fwrite(&nR, 4, 1, f);
            std::fwrite(cmLast.data(), 4, cmLast.size(), f);
            std::fwrite(cmKF.data(), 4, cmKF.size(), f);
            // the matching core of Fuse: key frame = right image (mono features); map point i sits half a pixel off the key
            // point (7 i) mod nR and carries that key point's descriptor and level
            std::vector<float> invS2(sf.size()), uRnone(nR, -1.0f), urMP(nR), d3F(nR), dminF(nR, 0.0f), dmaxF(nR, 1e9f);
            for (size_t l = 0; l < sf.size(); ++l) invS2[l] = 1.0f / (sf[l] * sf[l]);
            std::vector<cv::Point2f> uvF(nR);
            std::vector<int> lvF(nR);
            std::vector<uint8_t> validF(nR);
            cv::Mat descF(nR, 32, CV_8U);
            for (int i = 0; i < nR; ++i) {
                const int j = (int)(((long long)i * 7) % nR);
                uvF[i] = cv::Point2f(R.keys[j].pt.x + 0.5f, R.keys[j].pt.y - 0.5f);
                lvF[i] = R.keys[j].octave; validF[i] = i % 9 != 0; d3F[i] = 1.0f + (float)(i % 13);
                urMP[i] = uvF[i].x - 40.0f / d3F[i];
                std::memcpy(descF.ptr(i), R.desc.ptr(j), 32);
            }
            std::vector<int> fIdx, fDist;
            const int nf = mi.FuseSearch(R.keys, R.desc, grid, sf, invS2, uRnone, validF, uvF, urMP, d3F, dminF, dmaxF, lvF, descF,
                                         3.0f, fIdx, fDist);
            std::fwrite(&nf, 4, 1, f);
            std::fwrite(fIdx.data(), 4, fIdx.size(), f);
            std::fwrite(fDist.data(), 4, fDist.size(), f);
            // SearchForTriangulation left -> right: "vocabulary node" of a feature = its first descriptor byte mod 16,
            // deterministic map-point / stereo flags and epipolar table (the test rebuilds the same inputs)
            auto featvec = [](const cv::Mat& d) {
                std::vector<std::pair<unsigned, std::vector<unsigned> > > fv(16);
                for (unsigned n = 0; n < 16; ++n) fv[n].first = n;
                for (int i = 0; i < d.rows; ++i) fv[d.ptr(i)[0] % 16].second.push_back((unsigned)i);
                return fv;
            };
            std::vector<uint8_t> has1(nL), has2(nR), st1(nL), st2(nR);
            std::vector<float> a1(nL), a2(nR);
            for (int i = 0; i < nL; ++i) { has1[i] = i % 4 == 0; st1[i] = i % 3 == 0; a1[i] = L.keys[i].angle; }
            for (int i = 0; i < nR; ++i) { has2[i] = i % 5 == 0; st2[i] = i % 2 == 0; a2[i] = R.keys[i].angle; }
            std::vector<std::pair<size_t, size_t> > pairs;
            const int nt = mi.SearchForTriangulation(L.desc, a1, has1, st1, featvec(L.desc), R.desc, a2, has2, st2, R.keys,
                                                     featvec(R.desc), sf, cv::Point2f(376.f, 240.f),
                                                     [](size_t i1, size_t i2) { return (i1 * 31 + i2 * 17) % 5 != 0; }, false, false,
                                                     true, pairs);
            std::vector<int> m12t(nL, -1);
            for (const auto& pr : pairs) m12t[pr.first] = (int)pr.second;
            std::fwrite(&nt, 4, 1, f);
            std::fwrite(m12t.data(), 4, m12t.size(), f);
            // the loop-closing matchers on the Fuse scene above (every third feature of the key frame already matched)
            std::vector<uint8_t> occ(nR);
            for (int j = 0; j < nR; ++j) occ[j] = j % 3 == 0;
            std::vector<int> kmS, fIdxS, fDistS;
            const int ns = mi.SearchByProjectionSim3(R.keys, R.desc, grid, sf, occ, validF, uvF, d3F, dminF, dmaxF, lvF, descF, 4, 0.8f, kmS);
            const int nfs = mi.FuseSearchSim3(R.keys, R.desc, grid, sf, validF, uvF, d3F, dminF, dmaxF, lvF, descF, 4.0f, fIdxS, fDistS);
            std::fwrite(&ns, 4, 1, f); std::fwrite(&nfs, 4, 1, f);
            std::fwrite(kmS.data(), 4, kmS.size(), f);
            std::fwrite(fIdxS.data(), 4, fIdxS.size(), f);
            // SearchBySim3 between the right image and itself shifted: points of "key frame 1" (= right image) project half a
            // pixel off their own key point in "key frame 2" (= right image) and vice versa -> mutual matches
            std::vector<uint8_t> v1S(nR), v2S(nR);
            std::vector<cv::Point2f> uv1S(nR), uv2S(nR);
            std::vector<int> lvS(nR);
            for (int i = 0; i < nR; ++i) {
                v1S[i] = i % 4 != 0; v2S[i] = i % 5 != 0; lvS[i] = R.keys[i].octave;
                uv1S[i] = cv::Point2f(R.keys[i].pt.x + 0.5f, R.keys[i].pt.y - 0.5f);
                uv2S[i] = cv::Point2f(R.keys[i].pt.x - 0.25f, R.keys[i].pt.y + 0.25f);
            }
            std::vector<int> m12S;
            const int nsim = mi.SearchBySim3(R.keys, R.desc, grid, R.keys, R.desc, grid, sf, sf, v1S, uv1S, d3F, dminF, dmaxF, lvS, v2S,
                                             uv2S, d3F, dminF, dmaxF, lvS, 7.5f, m12S);
            std::fwrite(&nsim, 4, 1, f);
            std::fwrite(m12S.data(), 4, m12S.size(), f);
        }
        // the empty-image contract of operator() (ORBextractor.cc:1017)
        cv::Mat empty; Result E;
        const int e = exL(empty, cv::Mat(), E.keys, E.desc, lap);
        std::fwrite(&e, 4, 1, f);
        std::fclose(f);
    } catch (const std::exception& e) {
        std::fprintf(stderr, "adapter_driver: %s\n", e.what());
        return 1;
    }
    return 0;
}
