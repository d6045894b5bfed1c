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
#include "SparsePyrLK_accel.h"
#include <cstdio>
int main() {
    try {
        SparsePyrLKAccel flow(cv::Size(31, 31), 2, 20, 0.03);
        std::vector<cv::Point2f> old(1, cv::Point2f(20.f, 20.f)), next; std::vector<uchar> status; std::vector<float> err;
        cv::Mat fa(64, 64, CV_8U), fb(64, 64, CV_8U);
        flow.calc(fa, fb, old, next, status, err);
        std::printf("flow status %d\\n", (int)status[0]);
    } catch (const std::exception& e) { std::printf("no device: %s\\n", e.what()); }
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
           os.path.join(ad, "ORBmatcher_accel.cc"), os.path.join(ad, "SparsePyrLK_accel.cc"), "-o", str(exe), "-L", os.path.join(ROOT, "rumi_slam_b200"),
           "-lrumi_orb", "-Wl,-rpath," + os.path.join(ROOT, "rumi_slam_b200")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    assert "dist 256" in out.stdout
    assert "no device" in out.stdout or "empty-> -1" in out.stdout


def test_vocabulary_adapter_builds_against_reference_dbow2(tmp_path):
    """ORBVocabularyAccel fills the reference's own DBoW2::BowVector / FeatureVector containers, so it is compiled
    against R/Thirdparty/DBoW2 (present in the dev container only) over the cv / boost stand-ins."""
    import pytest
    dbow = "/root/reference/src/rumi-slam/Thirdparty/DBoW2"
    if not os.path.exists(os.path.join(dbow, "DBoW2", "BowVector.h")):
        pytest.skip("reference sources absent")
    from rumi_slam_b200 import _lib
    _lib.lib()
    ad = os.path.join(ROOT, "rumi_slam_b200", "adapter")
    voc = tmp_path / "voc.txt"
    voc.write_text("2 1 0 0\n0 1 " + " ".join(["0"] * 32) + " 1.0\n0 1 " + " ".join(["255"] * 32) + " 2.0\n")
    main = tmp_path / "main.cc"
    main.write_text("""
#include "ORBVocabulary_accel.h"
#include "ORBmatcher_accel.h"
#include <cstdio>
int main(int argc, char** argv) {
    try {
        ORB_SLAM3::ORBVocabularyAccel voc;
        bool ok = voc.loadFromTextFile(argv[1]);
        std::vector<cv::Mat> feats(3, cv::Mat(1, 32, CV_8U));
        for (int i = 0; i < 3; ++i) for (int b = 0; b < 32; ++b) feats[i].ptr(0)[b] = i == 1 ? 255 : 0;
        DBoW2::BowVector bv; DBoW2::FeatureVector fv;
        voc.transform(feats, bv, fv, 0);
        std::printf("loaded %d words %u bow %zu fv %zu\\n", (int)ok, voc.size(), bv.size(), fv.size());
        ORB_SLAM3::ORBmatcherAccel m(0.75f);
        std::vector<std::pair<unsigned, std::vector<unsigned> > > fva, fvb;
        std::vector<int> match;
        cv::Mat d(3, 32, CV_8U);
        m.SearchByBoW(d, std::vector<float>(3, 0.f), std::vector<uint8_t>(3, 1), fva, d, std::vector<float>(3, 0.f), fvb,
                      true, match);
    } catch (const std::exception& e) { std::printf("no device: %s\\n", e.what()); }
    return 0;
}
""")
    exe = tmp_path / "voc_test"
    cmd = ["g++", "-std=c++14", "-O1", "-w", "-I", os.path.join(ROOT, "oracle", "cvstub"),
           "-I", os.path.join(ROOT, "oracle", "booststub"), "-I", dbow, "-I", ad, "-I", os.path.join(ROOT, "include"),
           str(main), os.path.join(ad, "ORBVocabulary_accel.cc"), os.path.join(ad, "ORBmatcher_accel.cc"),
           os.path.join(dbow, "DBoW2", "BowVector.cpp"), os.path.join(dbow, "DBoW2", "FeatureVector.cpp"),
           "-o", str(exe), "-L", os.path.join(ROOT, "rumi_slam_b200"), "-lrumi_orb",
           "-Wl,-rpath," + os.path.join(ROOT, "rumi_slam_b200")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    out = subprocess.run([str(exe), str(voc)], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    assert "no device" in out.stdout or "loaded 1 words 2 bow 2 fv 1" in out.stdout
