// ORACLE -- TEST INFRASTRUCTURE ONLY.
// C entry points around the UNMODIFIED reference DBoW2 (R/Thirdparty/DBoW2/DBoW2/{TemplatedVocabulary.h, FORB.cpp,
// BowVector.cpp, FeatureVector.cpp, ScoringObject.cpp}), compiled where it lies over oracle/cvstub + oracle/booststub
// (see oracle/Makefile target `refbow`).  Used to pin oracle/bow_oracle.cpp: the vocabulary is loaded with the
// reference's own loadFromTextFile and features descend the tree through the reference's own transform().
#include <cstdint>
#include <cstring>
#include <sstream>
#include <vector>

#include <opencv2/core/core.hpp>
#include <opencv2/core/persistence_stub.hpp>
#include "DBoW2/FORB.h"
#include "DBoW2/TemplatedVocabulary.h"

typedef DBoW2::TemplatedVocabulary<DBoW2::FORB::TDescriptor, DBoW2::FORB> RefVoc;

struct RefVocAccess : public RefVoc {
    void one(const cv::Mat& f, DBoW2::WordId& w, DBoW2::WordValue& v, DBoW2::NodeId* nid, int levelsup) const {
        RefVoc::transform(f, w, v, nid, levelsup);        // protected in the reference
    }
};

extern "C" {

void* refbow_load(const char* path) {
    RefVocAccess* v = new RefVocAccess();
    if (!v->loadFromTextFile(path)) { delete v; return nullptr; }
    return v;
}
void refbow_free(void* p) { delete static_cast<RefVocAccess*>(p); }
int refbow_size(void* p) { return (int)static_cast<RefVocAccess*>(p)->size(); }

// per-feature descent: word id, weight, node id at (L - levelsup)
int refbow_transform(void* p, const uint8_t* desc, int n, int levelsup, int32_t* word, double* weight, int32_t* node) {
    const RefVocAccess* v = static_cast<RefVocAccess*>(p);
    for (int i = 0; i < n; ++i) {
        cv::Mat f(1, 32, CV_8U);
        std::memcpy(f.ptr(0), desc + 32 * (size_t)i, 32);
        DBoW2::WordId w; DBoW2::WordValue val; DBoW2::NodeId nid = 0;
        v->one(f, w, val, &nid, levelsup);
        word[i] = (int32_t)w; weight[i] = val; node[i] = (int32_t)nid;
    }
    return 0;
}

// Frame::ComputeBoW (R/lib_src/Frame.cc): transform(features, BowVector, FeatureVector, levelsup), flattened.
// bow_ids/bow_vals: capacity n; fv_nodes / fv_off (capacity n + 1) / fv_idx (capacity n).  Returns #words, *nfv = #nodes.
int refbow_vectors(void* p, const uint8_t* desc, int n, int levelsup, int32_t* bow_ids, double* bow_vals,
                   int32_t* fv_nodes, int32_t* fv_off, int32_t* fv_idx, int* nfv) {
    const RefVocAccess* v = static_cast<RefVocAccess*>(p);
    std::vector<cv::Mat> feats;
    for (int i = 0; i < n; ++i) {
        cv::Mat f(1, 32, CV_8U);
        std::memcpy(f.ptr(0), desc + 32 * (size_t)i, 32);
        feats.push_back(f);
    }
    DBoW2::BowVector bv;
    DBoW2::FeatureVector fv;
    v->transform(feats, bv, fv, levelsup);
    int k = 0;
    for (DBoW2::BowVector::const_iterator it = bv.begin(); it != bv.end(); ++it, ++k) { bow_ids[k] = (int32_t)it->first; bow_vals[k] = it->second; }
    int m = 0, o = 0;
    for (DBoW2::FeatureVector::const_iterator it = fv.begin(); it != fv.end(); ++it, ++m) {
        fv_nodes[m] = (int32_t)it->first; fv_off[m] = o;
        for (size_t j = 0; j < it->second.size(); ++j) fv_idx[o++] = (int32_t)it->second[j];
    }
    fv_off[m] = o;
    *nfv = m;
    return k;
}

}  // extern "C"
