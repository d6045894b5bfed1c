// Accelerated ORBVocabulary (R/include/cloud_edge_slam_lib/ORBVocabulary.h:
//   typedef DBoW2::TemplatedVocabulary<DBoW2::FORB::TDescriptor, DBoW2::FORB> ORBVocabulary;)
// Same loadFromTextFile / transform / size interface as the DBoW2 template the reference instantiates
// (R/Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:135-151, :241), so Frame::ComputeBoW / KeyFrame::ComputeBoW
// (mpORBvocabulary->transform(vCurrentDesc, mBowVec, mFeatVec, 4)) compile against it unchanged.  The tree lives on
// the device; the descent of every feature runs there (rumi_bow_transform), the BowVector / FeatureVector maps are
// DBoW2's own containers filled in the reference's order.
#ifndef ORBVOCABULARY_ACCEL_H
#define ORBVOCABULARY_ACCEL_H

#include <string>
#include <vector>
#include <opencv2/core/core.hpp>

#include "DBoW2/BowVector.h"
#include "DBoW2/FeatureVector.h"

struct rumi_vocab;

namespace ORB_SLAM3 {

class ORBVocabularyAccel {
public:
    explicit ORBVocabularyAccel(int device = 0);
    ~ORBVocabularyAccel();
    ORBVocabularyAccel(const ORBVocabularyAccel&) = delete;
    ORBVocabularyAccel& operator=(const ORBVocabularyAccel&) = delete;

    // TemplatedVocabulary::loadFromTextFile (:1338-1421): "k L scoring weighting" + one node per line
    bool loadFromTextFile(const std::string& filename);
    // number of words (TemplatedVocabulary::size)
    unsigned int size() const;
    bool empty() const { return size() == 0; }

    // TemplatedVocabulary::transform(features, BowVector&, FeatureVector&, levelsup) (:1128-1200)
    void transform(const std::vector<cv::Mat>& features, DBoW2::BowVector& v, DBoW2::FeatureVector& fv,
                   int levelsup) const;
    // TemplatedVocabulary::transform(features, BowVector&) (:1066-1122)
    void transform(const std::vector<cv::Mat>& features, DBoW2::BowVector& v) const;

private:
    void descend(const std::vector<cv::Mat>& features, int levelsup, std::vector<int>& word, std::vector<double>& weight,
                 std::vector<int>& node) const;
    rumi_vocab* voc;
    int device, m_k, m_L;
    DBoW2::ScoringType m_scoring;
    DBoW2::WeightingType m_weighting;
};

}  // namespace ORB_SLAM3
#endif
