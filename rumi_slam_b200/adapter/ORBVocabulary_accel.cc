#include "ORBVocabulary_accel.h"

#include <cstring>
#include <fstream>
#include <sstream>
#include <stdexcept>

#include "rumi_orb.h"

namespace ORB_SLAM3 {

ORBVocabularyAccel::ORBVocabularyAccel(int dev)
    : voc(nullptr), device(dev), m_k(0), m_L(0), m_scoring(DBoW2::L1_NORM), m_weighting(DBoW2::TF_IDF) {}
ORBVocabularyAccel::~ORBVocabularyAccel() { rumi_vocab_destroy(voc); }

bool ORBVocabularyAccel::loadFromTextFile(const std::string& filename) {
    std::ifstream f(filename.c_str());
    if (!f.is_open()) return false;
    std::string s;
    std::getline(f, s);
    std::stringstream ss(s);
    int n1 = -1, n2 = -1;
    ss >> m_k >> m_L >> n1 >> n2;
    if (m_k < 0 || m_k > 20 || m_L < 1 || m_L > 10 || n1 < 0 || n1 > 5 || n2 < 0 || n2 > 3) return false;   // :1359
    m_scoring = (DBoW2::ScoringType)n1;
    m_weighting = (DBoW2::WeightingType)n2;
    std::vector<int32_t> parent(1, 0);
    std::vector<uint8_t> leaf(1, 0), desc(32, 0);
    std::vector<double> weight(1, 0.0);
    while (std::getline(f, s)) {
        if (s.empty()) continue;
        std::stringstream sn(s);
        int pid = 0, isLeaf = 0;
        sn >> pid >> isLeaf;
        parent.push_back(pid);
        leaf.push_back(isLeaf > 0 ? 1 : 0);
        for (int i = 0; i < 32; ++i) { int b = 0; sn >> b; desc.push_back((uint8_t)b); }   // FORB::fromString
        double w = 0.0;
        sn >> w;
        weight.push_back(w);
    }
    rumi_vocab_destroy(voc);
    voc = nullptr;
    if (rumi_vocab_create(&voc, device, m_k, m_L, (int)parent.size(), parent.data(), leaf.data(), desc.data(),
                          weight.data()) != RUMI_OK)
        throw std::runtime_error(std::string("ORBVocabularyAccel: ") + rumi_last_error());
    return true;
}

unsigned int ORBVocabularyAccel::size() const { return (unsigned int)rumi_vocab_words(voc); }

void ORBVocabularyAccel::descend(const std::vector<cv::Mat>& features, int levelsup, std::vector<int>& word,
                                 std::vector<double>& weight, std::vector<int>& node) const {
    const int n = (int)features.size();
    std::vector<uint8_t> rows(32 * (size_t)n);
    for (int i = 0; i < n; ++i) std::memcpy(rows.data() + 32 * (size_t)i, features[i].ptr(0), 32);
    word.assign(n, 0); weight.assign(n, 0.0); node.assign(n, 0);
    if (n == 0) return;
    if (rumi_bow_transform(voc, rows.data(), n, levelsup, word.data(), weight.data(), node.data()) != RUMI_OK)
        throw std::runtime_error(std::string("ORBVocabularyAccel: ") + rumi_last_error());
}

void ORBVocabularyAccel::transform(const std::vector<cv::Mat>& features, DBoW2::BowVector& v,
                                   DBoW2::FeatureVector& fv, int levelsup) const {
    v.clear();
    fv.clear();
    if (empty()) return;
    // ScoringObject.h:73-89: every scoring object but DotProduct normalises, L2Scoring with L2, the others with L1
    const bool must = m_scoring != DBoW2::DOT_PRODUCT;
    const DBoW2::LNorm norm = m_scoring == DBoW2::L2_NORM ? DBoW2::L2 : DBoW2::L1;
    std::vector<int> word, node;
    std::vector<double> weight;
    descend(features, levelsup, word, weight, node);
    const bool tf = m_weighting == DBoW2::TF || m_weighting == DBoW2::TF_IDF;
    for (unsigned int i = 0; i < features.size(); ++i) {
        if (weight[i] > 0) {                                   // not stopped
            if (tf) v.addWeight((DBoW2::WordId)word[i], weight[i]);
            else v.addIfNotExist((DBoW2::WordId)word[i], weight[i]);
            fv.addFeature((DBoW2::NodeId)node[i], i);
        }
    }
    if (tf && !v.empty() && !must) {
        const double nd = v.size();
        for (DBoW2::BowVector::iterator vit = v.begin(); vit != v.end(); vit++) vit->second /= nd;
    }
    if (must) v.normalize(norm);
}

void ORBVocabularyAccel::transform(const std::vector<cv::Mat>& features, DBoW2::BowVector& v) const {
    DBoW2::FeatureVector fv;
    transform(features, v, fv, 0);
}

}  // namespace ORB_SLAM3
