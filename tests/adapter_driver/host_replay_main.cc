// TEST INFRASTRUCTURE (CPU, no GPU): executes the HOST logic of ORBmatcherAccel::SearchByProjection(..., LocalPointsExtras, ...)
// -- window construction on both grids, the order-dependent acceptance replay, the stereo-fisheye cross assignments -- and
// writes the result for tests/test_adapter_host_replay.py, which compares it with the oracle pinned to the reference function.
// The one thing the GPU does on that path, the Hamming distances of the candidate lists, comes from the three C-ABI stand-ins
// BELOW, which exist only in this test binary (they shadow the library's entry points at link time).  The product has no
// such path: librumi_orb.so fails without a device.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "ORBmatcher_accel.h"
#include "rumi_orb.h"

extern "C" {
int rumi_match_create(rumi_match** out, int) { *out = reinterpret_cast<rumi_match*>(new int(0)); return 0; }
void rumi_match_destroy(rumi_match* m) { delete reinterpret_cast<int*>(m); }
int rumi_hamming_candidates(rumi_match*, const uint8_t* Q, int nq, const uint8_t* T, int, const int32_t* off, const int32_t* idx,
                            uint16_t* dist, int32_t*, uint16_t*, int32_t*, uint16_t*) {
    for (int q = 0; q < nq; ++q)
        for (int p = off[q]; p < off[q + 1]; ++p) {
            int d = 0;
            for (int b = 0; b < 32; ++b) d += __builtin_popcount(Q[32 * (size_t)q + b] ^ T[32 * (size_t)idx[p] + b]);
            dist[p] = (uint16_t)d;
        }
    return 0;
}
}

namespace {
struct Reader {
    FILE* f;
    template <class T> std::vector<T> arr() {
        int32_t n = 0;
        if (std::fread(&n, 4, 1, f) != 1) std::exit(3);
        std::vector<T> v(n > 0 ? n : 0);
        if (n > 0 && std::fread(v.data(), sizeof(T), n, f) != (size_t)n) std::exit(3);
        return v;
    }
};
struct KpRec { float x, y, size, angle, response; int32_t octave, class_id; };
std::vector<cv::KeyPoint> keypoints(const std::vector<KpRec>& r) {
    std::vector<cv::KeyPoint> k(r.size());
    for (size_t i = 0; i < r.size(); ++i) {
        k[i].pt.x = r[i].x; k[i].pt.y = r[i].y; k[i].size = r[i].size; k[i].angle = r[i].angle; k[i].response = r[i].response;
        k[i].octave = r[i].octave; k[i].class_id = r[i].class_id;
    }
    return k;
}
cv::Mat rows(const std::vector<uint8_t>& d) {
    cv::Mat m((int)(d.size() / 32), 32, CV_8U);
    if (!d.empty()) std::memcpy(m.data, d.data(), d.size());
    return m;
}
std::vector<cv::Point2f> points(const std::vector<float>& v) {
    std::vector<cv::Point2f> p(v.size() / 2);
    for (size_t i = 0; i < p.size(); ++i) { p[i].x = v[2 * i]; p[i].y = v[2 * i + 1]; }
    return p;
}
}  // namespace

// usage: host_replay <scenario.bin> <out.bin>.  Scenario = int32-counted arrays in the order read below (count 0 = absent).
int main(int argc, char** argv) {
    if (argc < 3) return 2;
    Reader r{std::fopen(argv[1], "rb")};
    if (!r.f) return 2;
    const std::vector<float> par = r.arr<float>();                    // th, ratio, minX, minY, maxX, maxY [, mode, fwd, bwd, checkOri]
    if (par.size() > 6 && par[6] == 1.0f) {                          // SearchByProjectionLastFrameFisheye
        const std::vector<cv::KeyPoint> kC = keypoints(r.arr<KpRec>()), kR = keypoints(r.arr<KpRec>());
        const cv::Mat dC = rows(r.arr<uint8_t>()), dMP = rows(r.arr<uint8_t>());
        const std::vector<float> sf = r.arr<float>();
        const std::vector<uint8_t> occupied = r.arr<uint8_t>(), valid = r.arr<uint8_t>(), hasObs = r.arr<uint8_t>();
        const std::vector<cv::Point2f> uv = points(r.arr<float>()), uvR = points(r.arr<float>());
        const std::vector<float> invz = r.arr<float>(), angleLast = r.arr<float>();
        const std::vector<int> octave = r.arr<int>();
        std::fclose(r.f);
        ORB_SLAM3::FrameGridAccel gC(kC, par[2], par[3], par[4], par[5]), gR(kR, par[2], par[3], par[4], par[5]);
        ORB_SLAM3::ORBmatcherAccel m(par[1]);
        std::vector<int> cm;
        const int n = m.SearchByProjectionLastFrameFisheye(kC, kR, dC, gC, gR, sf, occupied, valid, uv, uvR, invz, octave, angleLast,
                                                           dMP, hasObs, par[0], par[7] != 0, par[8] != 0, par[9] != 0, cm);
        FILE* o = std::fopen(argv[2], "wb");
        const int32_t head[3] = {n, (int32_t)cm.size(), -1};
        std::fwrite(head, 4, 3, o);
        std::fwrite(cm.data(), 4, cm.size(), o);
        std::fclose(o);
        return 0;
    }
    const std::vector<cv::KeyPoint> kL = keypoints(r.arr<KpRec>()), kR = keypoints(r.arr<KpRec>());
    const cv::Mat dF = rows(r.arr<uint8_t>()), dMP = rows(r.arr<uint8_t>());
    const std::vector<float> sf = r.arr<float>();
    const std::vector<cv::Point2f> proj = points(r.arr<float>()), projR = points(r.arr<float>());
    const std::vector<int> level = r.arr<int>(), levelR = r.arr<int>();
    const std::vector<float> viewCos = r.arr<float>(), viewCosR = r.arr<float>();
    const std::vector<uint8_t> hasObs = r.arr<uint8_t>(), occupied = r.arr<uint8_t>(), inView = r.arr<uint8_t>(),
                               inViewR = r.arr<uint8_t>();
    const std::vector<float> uRight = r.arr<float>();
    const std::vector<int> l2r = r.arr<int>(), r2l = r.arr<int>();
    std::fclose(r.f);

    ORB_SLAM3::FrameGridAccel gL(kL, par[2], par[3], par[4], par[5]), gR(kR, par[2], par[3], par[4], par[5]);
    ORB_SLAM3::ORBmatcherAccel m(par[1]);
    ORB_SLAM3::ORBmatcherAccel::LocalPointsExtras ex;
    if (!occupied.empty()) ex.occupied = &occupied;
    if (!uRight.empty()) ex.uRight = &uRight;
    if (!projR.empty()) ex.projR = &projR;
    if (!kR.empty()) { ex.keysRight = &kR; ex.gridRight = &gR; }
    if (!inView.empty()) ex.inView = &inView;
    if (!inViewR.empty()) ex.inViewR = &inViewR;
    if (!levelR.empty()) ex.levelR = &levelR;
    if (!viewCosR.empty()) ex.viewCosR = &viewCosR;
    if (!l2r.empty()) ex.leftToRight = &l2r;
    if (!r2l.empty()) ex.rightToLeft = &r2l;
    std::vector<int> fm, fm0;
    const int n = m.SearchByProjection(kL, dF, gL, sf, proj, level, viewCos, dMP, hasObs, par[0], ex, fm);
    // the original entry point on the same inputs (only meaningful when no extras are set: both must agree then)
    int n0 = -1;
    if (kR.empty() && inView.empty()) n0 = m.SearchByProjection(kL, dF, gL, sf, proj, level, viewCos, dMP, hasObs, par[0], fm0);
    FILE* o = std::fopen(argv[2], "wb");
    const int32_t head[3] = {n, (int32_t)fm.size(), n0};
    std::fwrite(head, 4, 3, o);
    std::fwrite(fm.data(), 4, fm.size(), o);
    const int32_t n0s = (int32_t)fm0.size();
    std::fwrite(&n0s, 4, 1, o);
    std::fwrite(fm0.data(), 4, fm0.size(), o);
    std::fclose(o);
    return 0;
}
