"""The C++ adapter (reference-signature ORBextractor / ORBmatcher core over the C ABI) compiles and links against
librumi_orb.so.  OpenCV headers are absent in this image, so the build uses the minimal cv stand-in of
oracle/cvstub (test infrastructure); with a real OpenCV the same sources build unchanged."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_adapter_builds_and_links(tmp_path):
    from rumi_slam_b200 import _lib
    _lib.lib()
    ad = os.path.join(ROOT, "rumi_slam_b200", "adapter")
    main = tmp_path / "main.cc"
    main.write_text("""
#include "ORBextractor.h"
#include "ORBmatcher_accel.h"
#include <cstdio>
int main() {
    try {
        ORB_SLAM3::ORBextractor ex(1000, 1.2f, 8, 20, 7);
        std::vector<cv::KeyPoint> k; cv::Mat d; std::vector<int> lap = {0, 0};
        cv::Mat empty;
        int r = ex(empty, cv::Mat(), k, d, lap);
        std::printf("levels %d empty-> %d\\n", ex.GetLevels(), r);
    } catch (const std::exception& e) { std::printf("no device: %s\\n", e.what()); }
    unsigned char a[32] = {0}, b[32]; for (int i = 0; i < 32; ++i) b[i] = 255;
    cv::Mat ma(1, 32, CV_8U), mb(1, 32, CV_8U);
    for (int i = 0; i < 32; ++i) { ma.ptr(0)[i] = a[i]; mb.ptr(0)[i] = b[i]; }
    std::printf("dist %d\\n", ORB_SLAM3::ORBmatcherAccel::DescriptorDistance(ma, mb));
    return 0;
}
""")
    exe = tmp_path / "adapter_test"
    cmd = ["g++", "-std=c++14", "-O1", "-I", os.path.join(ROOT, "oracle", "cvstub"), "-I", ad,
           "-I", os.path.join(ROOT, "include"), str(main), os.path.join(ad, "ORBextractor.cc"),
           os.path.join(ad, "ORBmatcher_accel.cc"), "-o", str(exe), "-L", os.path.join(ROOT, "rumi_slam_b200"),
           "-lrumi_orb", "-Wl,-rpath," + os.path.join(ROOT, "rumi_slam_b200")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    assert "dist 256" in out.stdout
    assert "no device" in out.stdout or "empty-> -1" in out.stdout
