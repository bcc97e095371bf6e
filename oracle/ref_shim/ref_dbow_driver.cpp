// ref_dbow_driver.cpp — TEST INFRASTRUCTURE.  C entry points around the reference's own DBoW2 (Thirdparty/DBoW2, compiled where
// it lies): ORBVocabulary = TemplatedVocabulary<FORB::TDescriptor, FORB> (include/ORBVocabulary.h), its text loader
// (TemplatedVocabulary.h:1338-1420) and transform(features, BowVector&, FeatureVector&, levelsup) (:1175-1215 -> :1218-1259),
// the call FrameKTL::ComputeBoW / KeyFrame::ComputeBoW make (src/FrameKTL.cc:439-446, src/KeyFrame.cc:203-210).
#include <stdint.h>
#include <string.h>
#include <vector>
#include "ORBVocabulary.h"

extern "C" {

void* refv_load(const char* path)
{
    USLAM::ORBVocabulary* v = new USLAM::ORBVocabulary();
    if (!v->loadFromTextFile(path)) { delete v; return 0; }
    return v;
}
void refv_destroy(void* v) { delete (USLAM::ORBVocabulary*)v; }
int refv_size(void* v) { return (int)((USLAM::ORBVocabulary*)v)->size(); }
// transform n descriptors; bow_*: word id / value pairs in map order; fv_node / fv_feat: (node id, feature index) pairs in map
// order, features in insertion order.  Returns 0; counts through n_bow / n_fv; -2 if a capacity is too small.
int refv_transform(void* v_, const uint8_t* desc, int n, int levelsup, int32_t* bow_word, double* bow_value, int* n_bow, int cap_bow,
                   int32_t* fv_node, int32_t* fv_feat, int* n_fv, int cap_fv)
{
    USLAM::ORBVocabulary* v = (USLAM::ORBVocabulary*)v_;
    std::vector<cv::Mat> feats((size_t)n);
    for (int i = 0; i < n; i++) { feats[(size_t)i] = cv::Mat(1, 32, CV_8UC1); memcpy(feats[(size_t)i].data, desc + (size_t)i * 32, 32); }
    DBoW2::BowVector bow; DBoW2::FeatureVector fv;
    v->transform(feats, bow, fv, levelsup);
    int nb = 0, nf = 0;
    for (DBoW2::BowVector::const_iterator it = bow.begin(); it != bow.end(); ++it, ++nb)
        if (nb < cap_bow) { bow_word[nb] = (int32_t)it->first; bow_value[nb] = it->second; }
    for (DBoW2::FeatureVector::const_iterator it = fv.begin(); it != fv.end(); ++it)
        for (size_t j = 0; j < it->second.size(); j++, nf++)
            if (nf < cap_fv) { fv_node[nf] = (int32_t)it->first; fv_feat[nf] = (int32_t)it->second[j]; }
    *n_bow = nb; *n_fv = nf;
    return (nb > cap_bow || nf > cap_fv) ? -2 : 0;
}

}  // extern "C"
