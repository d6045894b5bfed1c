// Drop-in replacement header for R/include/cloud_edge_slam_lib/ORBextractor.h (ORB_SLAM3::ORBextractor).
// Same class name, namespace, constructor, call operator, getters and public mvImagePyramid, so Frame / KeyFrame /
// Tracking / KFDSample (R/lib_src/Frame.cc:473-479, R/lib_src/KFDSample.cc:113,153, R/lib_src/Tracking.cc:575-581)
// compile and link unchanged; every call is forwarded to the C ABI of librumi_orb.so (include/rumi_orb.h).
// There is no CPU implementation behind it: without a CUDA device construction throws std::runtime_error.
#ifndef ORBEXTRACTOR_H
#define ORBEXTRACTOR_H

#include <list>
#include <vector>
#include <opencv2/opencv.hpp>

struct rumi_orb;

namespace ORB_SLAM3 {

class ORBextractor {
public:
    enum { HARRIS_SCORE = 0, FAST_SCORE = 1 };

    ORBextractor(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST);
    ~ORBextractor();
    ORBextractor(const ORBextractor&) = delete;
    ORBextractor& operator=(const ORBextractor&) = delete;

    int CloudFrameComputeDescriptors(cv::InputArray _image, const std::vector<cv::KeyPoint>& _keypoints,
                                     cv::OutputArray _descriptors);

    // Mask is ignored, exactly as in the reference (R/lib_src/ORBextractor.cc:1014-1091).
    int operator()(cv::InputArray _image, cv::InputArray _mask, std::vector<cv::KeyPoint>& _keypoints,
                   cv::OutputArray _descriptors, std::vector<int>& vLappingArea);

    // operator() in two halves (rumi_orb_extract_begin / _end): Begin enqueues the frame and returns, End waits and fills the
    // outputs.  Lets ONE thread keep the left and the right extractor of a stereo frame busy instead of the two std::threads
    // of Frame::Frame (R/lib_src/Frame.cc:116-119).  Begin returns false for an empty image (End then returns -1).
    bool Begin(cv::InputArray _image, std::vector<int>& vLappingArea);
    int End(std::vector<cv::KeyPoint>& _keypoints, cv::OutputArray _descriptors);

    int inline GetLevels() { return nlevels; }
    float inline GetScaleFactor() { return scaleFactor; }
    std::vector<float> inline GetScaleFactors() { return mvScaleFactor; }
    std::vector<float> inline GetInverseScaleFactors() { return mvInvScaleFactor; }
    std::vector<float> inline GetScaleSigmaSquares() { return mvLevelSigma2; }
    std::vector<float> inline GetInverseScaleSigmaSquares() { return mvInvLevelSigma2; }

    // Filled after every operator() call (read by Frame::ComputeStereoMatches, R/lib_src/Frame.cc:834,918-932).
    // Monocular callers that never read it can switch the device-to-host copy off.
    std::vector<cv::Mat> mvImagePyramid;
    void SetPyramidDownload(bool on);
    // C-ABI handle (lets ORBmatcherAccel::ComputeStereoMatches use the device-resident pyramid in place)
    rumi_orb* Handle() const { return handle; }

    // GPU selection for multi-GPU hosts (process-wide default for extractors constructed afterwards).
    static void SetDefaultDevice(int device);

protected:
    int nfeatures;
    double scaleFactor;
    int nlevels;
    int iniThFAST;
    int minThFAST;
    std::vector<int> mnFeaturesPerLevel;
    std::vector<float> mvScaleFactor;
    std::vector<float> mvInvScaleFactor;
    std::vector<float> mvLevelSigma2;
    std::vector<float> mvInvLevelSigma2;

private:
    rumi_orb* handle;
    bool downloadPyramid;
    std::vector<unsigned char> kpBuf, descBuf;
    cv::Mat pendingImage;                // keeps the frame of Begin alive until End
    bool pendingEmpty = false;
};

}  // namespace ORB_SLAM3

#endif
