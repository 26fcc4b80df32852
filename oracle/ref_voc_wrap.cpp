// oracle/ref_voc_wrap.cpp -- TEST INFRASTRUCTURE.  C entry points around the reference's OWN vendored DBoW2 (Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h,
// FORB.cpp, BowVector.cpp, FeatureVector.cpp, ScoringObject.cpp, DUtils/Random.cpp), compiled unmodified from /root/reference against oracle/matchshim
// into oracle/_ref/libref_voc.so (oracle/Makefile).  ORBVocabulary is the reference's typedef (include/ORBVocabulary.h:31).  The reference snapshot ships
// no vocabulary file, so tests write synthetic trees in the ORBvoc.txt format and load them through the reference's own loadFromTextFile.
// Pins oracle/bow_oracle.cpp (tests/test_oracle_voc_vs_ref.py, tests/golden/voc_ref.npz).
#include "TemplatedVocabulary.h"
#include "FORB.h"
#include <cstring>

typedef DBoW2::TemplatedVocabulary<DBoW2::FORB::TDescriptor, DBoW2::FORB> ORBVocabulary;

extern "C" {

void* ref_voc_load(const char* txt_path) {                        // System.cc:80 mpVocabulary->loadFromTextFile(strVocFile)
    ORBVocabulary* v = new ORBVocabulary();
    if (!v->loadFromTextFile(txt_path)) { delete v; return nullptr; }
    return v;
}
void ref_voc_free(void* h) { delete (ORBVocabulary*)h; }
int ref_voc_size(void* h) { return (int)((ORBVocabulary*)h)->size(); }

// Frame::ComputeBoW (src/Frame.cc:348-355): transform(vCurrentDesc, mBowVec, mFeatVec, levelsup).  Outputs as oracle_voc_vectors: bow_words / bow_values
// (ascending word id), fv_nodes / fv_start / fv_items (ascending node id, feature indices in push order), counts2 = {#words, #nodes}.
int ref_voc_transform(void* h, const uint8_t* desc, int n, int levelsup, int32_t* bow_words, double* bow_values, int32_t* fv_nodes, int32_t* fv_start,
                      int32_t* fv_items, int32_t* counts2) {
    std::vector<cv::Mat> feats(n);
    for (int i = 0; i < n; i++) feats[i] = cv::Mat(1, 32, CV_8U, (void*)(desc + 32 * (size_t)i)).clone();
    DBoW2::BowVector bv; DBoW2::FeatureVector fv;
    ((ORBVocabulary*)h)->transform(feats, bv, fv, levelsup);
    int nw = 0;
    for (DBoW2::BowVector::const_iterator it = bv.begin(); it != bv.end(); ++it, ++nw) { bow_words[nw] = (int32_t)it->first; bow_values[nw] = it->second; }
    int nn = 0, ni = 0;
    for (DBoW2::FeatureVector::const_iterator it = fv.begin(); it != fv.end(); ++it, ++nn) {
        fv_nodes[nn] = (int32_t)it->first; fv_start[nn] = ni;
        for (size_t j = 0; j < it->second.size(); j++) fv_items[ni++] = (int32_t)it->second[j];
    }
    fv_start[nn] = ni;
    counts2[0] = nw; counts2[1] = nn;
    return 0;
}

// TemplatedVocabulary::score (L1 by default, ScoringObject.cpp) between the BowVectors of two descriptor sets
double ref_voc_score(void* h, const uint8_t* d1, int n1, const uint8_t* d2, int n2) {
    std::vector<cv::Mat> f1(n1), f2(n2);
    for (int i = 0; i < n1; i++) f1[i] = cv::Mat(1, 32, CV_8U, (void*)(d1 + 32 * (size_t)i)).clone();
    for (int i = 0; i < n2; i++) f2[i] = cv::Mat(1, 32, CV_8U, (void*)(d2 + 32 * (size_t)i)).clone();
    DBoW2::BowVector a, b; DBoW2::FeatureVector fa, fb;
    ((ORBVocabulary*)h)->transform(f1, a, fa, 4); ((ORBVocabulary*)h)->transform(f2, b, fb, 4);
    return ((ORBVocabulary*)h)->score(a, b);
}

}  // extern "C"
