// ORB_SLAM3::ORBextractor over librumi_orb.so -- see ORBextractor.h.  Marshals cv::Mat / std::vector<cv::KeyPoint>
// to the plain buffers of include/rumi_orb.h; no image arithmetic happens here.
#include "ORBextractor.h"

#include <cassert>
#include <cstring>
#include <stdexcept>
#include <string>

#include "rumi_orb.h"

namespace ORB_SLAM3 {

static int g_defaultDevice = 0;
void ORBextractor::SetDefaultDevice(int device) { g_defaultDevice = device; }

static_assert(sizeof(cv::KeyPoint) == sizeof(rumi_kp), "cv::KeyPoint must be the 28-byte record of rumi_kp");

ORBextractor::ORBextractor(int _nfeatures, float _scaleFactor, int _nlevels, int _iniThFAST, int _minThFAST)
    : nfeatures(_nfeatures), scaleFactor(_scaleFactor), nlevels(_nlevels), iniThFAST(_iniThFAST),
      minThFAST(_minThFAST), handle(nullptr), downloadPyramid(true) {
    if (rumi_orb_create(&handle, _nfeatures, _scaleFactor, _nlevels, _iniThFAST, _minThFAST, g_defaultDevice, 1) != RUMI_OK)
        throw std::runtime_error(std::string("ORBextractor: ") + rumi_last_error());
    mvScaleFactor.resize(nlevels); mvInvScaleFactor.resize(nlevels);
    mvLevelSigma2.resize(nlevels); mvInvLevelSigma2.resize(nlevels);
    mnFeaturesPerLevel.resize(nlevels);
    rumi_orb_tables(handle, mvScaleFactor.data(), mvInvScaleFactor.data(), mvLevelSigma2.data(),
                    mvInvLevelSigma2.data(), mnFeaturesPerLevel.data());
    mvImagePyramid.resize(nlevels);
    rumi_orb_set_pyramid_staging(handle, 1);               // downloadPyramid is on by default
}

ORBextractor::~ORBextractor() { rumi_orb_destroy(handle); }

void ORBextractor::SetPyramidDownload(bool on) {
    downloadPyramid = on;
    rumi_orb_set_pyramid_staging(handle, on ? 1 : 0);     // the levels then come down behind the kernels of every frame
}

bool ORBextractor::Begin(cv::InputArray _image, std::vector<int>& vLappingArea) {
    pendingEmpty = _image.empty();
    if (pendingEmpty) return false;                                  // :1017
    pendingImage = _image.getMat();
    assert(pendingImage.type() == CV_8UC1);                          // :1021
    const int rc = rumi_orb_extract_begin(handle, pendingImage.data, pendingImage.cols, pendingImage.rows,
                                          (size_t)pendingImage.step, vLappingArea[0], vLappingArea[1]);
    if (rc == RUMI_ERR_EMPTY) { pendingEmpty = true; return false; }
    if (rc != RUMI_OK) throw std::runtime_error(std::string("ORBextractor: ") + rumi_last_error());
    return true;
}

int ORBextractor::End(std::vector<cv::KeyPoint>& _keypoints, cv::OutputArray _descriptors) {
    if (pendingEmpty) return -1;
    const int cap = rumi_orb_frame_capacity(handle, pendingImage.cols, pendingImage.rows);
    if (cap < 0) throw std::runtime_error(std::string("ORBextractor: ") + rumi_last_error());
    kpBuf.resize(sizeof(rumi_kp) * (size_t)cap);
    descBuf.resize(32 * (size_t)cap);
    int nkp = 0, mono = 0;
    const int rc = rumi_orb_extract_end(handle, reinterpret_cast<rumi_kp*>(kpBuf.data()), descBuf.data(), cap, &nkp, &mono);
    pendingImage = cv::Mat();
    if (rc != RUMI_OK) throw std::runtime_error(std::string("ORBextractor: ") + rumi_last_error());
    if (nkp == 0) {
        _descriptors.release();                                      // :1035-1036
    } else {
        _descriptors.create(nkp, 32, CV_8U);                         // :1038
        cv::Mat d = _descriptors.getMat();
        for (int i = 0; i < nkp; ++i) std::memcpy(d.ptr(i), descBuf.data() + 32 * (size_t)i, 32);
    }
    _keypoints = std::vector<cv::KeyPoint>(nkp);                     // :1044
    if (nkp) std::memcpy(static_cast<void*>(_keypoints.data()), kpBuf.data(), sizeof(rumi_kp) * (size_t)nkp);
    if (downloadPyramid) {
        for (int l = 0; l < nlevels; ++l) {
            int w = 0, h = 0;
            rumi_orb_pyramid_level(handle, l, nullptr, 0, &w, &h);
            mvImagePyramid[l].create(h, w, CV_8UC1);
            rumi_orb_pyramid_level(handle, l, mvImagePyramid[l].data, (size_t)mvImagePyramid[l].step, &w, &h);
        }
    }
    return mono;                                                     // :1090
}

int ORBextractor::operator()(cv::InputArray _image, cv::InputArray /*_mask*/, std::vector<cv::KeyPoint>& _keypoints,
                             cv::OutputArray _descriptors, std::vector<int>& vLappingArea) {
    if (!Begin(_image, vLappingArea)) return -1;
    return End(_keypoints, _descriptors);
}

int ORBextractor::CloudFrameComputeDescriptors(cv::InputArray _image, const std::vector<cv::KeyPoint>& _keypoints,
                                               cv::OutputArray _descriptors) {
    if (_image.empty()) return -1;                                   // :990-991
    cv::Mat image = _image.getMat();
    assert(image.type() == CV_8UC1);
    const int n = (int)_keypoints.size();
    if (n == 0) {
        _descriptors.release();
        return 0;
    }
    _descriptors.create(n, 32, CV_8U);
    cv::Mat d = _descriptors.getMat();
    descBuf.resize(32 * (size_t)n);
    const int rc = rumi_orb_describe(handle, image.data, image.cols, image.rows, (size_t)image.step,
                                     reinterpret_cast<const rumi_kp*>(_keypoints.data()), n, descBuf.data());
    if (rc < 0) throw std::runtime_error(std::string("ORBextractor: ") + rumi_last_error());
    for (int i = 0; i < n; ++i) std::memcpy(d.ptr(i), descBuf.data() + 32 * (size_t)i, 32);
    return n;
}

}  // namespace ORB_SLAM3
