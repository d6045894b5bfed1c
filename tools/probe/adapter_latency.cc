// Latency of ORB_SLAM3::ORBextractor::operator() through the C++ adapter, with and without the mvImagePyramid download.
// build: g++ -std=c++14 -O2 -I oracle/cvstub -I rumi_slam_b200/adapter -I include tools/probe/adapter_latency.cc \
//        rumi_slam_b200/adapter/ORBextractor.cc -L rumi_slam_b200 -lrumi_orb -Wl,-rpath,$PWD/rumi_slam_b200 -o /tmp/adapter_latency
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <vector>

#include "ORBextractor.h"
#include "rumi_orb.h"

int main() {
    const int w = 640, h = 480;
    cv::Mat img(h, w, CV_8U);
    unsigned s = 12345;
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            s = s * 1664525u + 1013904223u;
            img.ptr(y)[x] = (unsigned char)(((x / 16 + y / 16) % 2 ? 160 : 60) + (s >> 28));
        }
    ORB_SLAM3::ORBextractor ex(1000, 1.2f, 8, 20, 7);
    std::vector<int> lap = {0, 0};
    for (int mode = 0; mode < 3; ++mode) {
        ex.SetPyramidDownload(mode != 1);
        if (mode == 2) rumi_orb_set_pyramid_staging(ex.Handle(), 0);      // the pre-staging behaviour: one device round trip per level
        std::vector<double> t;
        for (int i = 0; i < 220; ++i) {
            std::vector<cv::KeyPoint> k; cv::Mat d;
            const auto t0 = std::chrono::steady_clock::now();
            ex(img, cv::Mat(), k, d, lap);
            const auto t1 = std::chrono::steady_clock::now();
            if (i >= 20) t.push_back(std::chrono::duration<double, std::milli>(t1 - t0).count());
        }
        std::sort(t.begin(), t.end());
        std::printf("operator() through the C++ adapter, pyramid download %s: median %.3f ms  p90 %.3f ms\n", mode == 0 ? "on (staged)" : mode == 1 ? "off" : "on (unstaged)",
                    t[t.size() / 2], t[t.size() * 9 / 10]);
    }
    return 0;
}
