// C entry points over the reference's own DBoW2 (ThirdParty/DBoW2/DBoW2/TemplatedVocabulary.h, FORB.cpp, BowVector.cpp,
// FeatureVector.cpp, ScoringObject.cpp), compiled from /root/reference where it lies into oracle/_ref/libdbow_ref.so
// (oracle/Makefile, target dbow_ref).  TEST INFRASTRUCTURE: used by tests/golden/make_golden_bow.py to produce the golden
// vectors that pin oracle/bow_oracle.c, and by nothing in the product.
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "DBoW2/FORB.h"
#include "DBoW2/TemplatedVocabulary.h"

typedef DBoW2::TemplatedVocabulary<DBoW2::FORB::TDescriptor, DBoW2::FORB> OrbVocabulary;

static std::vector<cv::Mat> to_features(const uint8_t* desc, int n)
{
    std::vector<cv::Mat> f((size_t)n);
    for (int i = 0; i < n; i++) {
        f[i].create(1, 32, CV_8U);
        std::memcpy(f[i].ptr<unsigned char>(), desc + (size_t)i * 32, 32);
    }
    return f;
}

static DBoW2::BowVector to_bow(const uint32_t* words, const double* vals, int n)
{
    DBoW2::BowVector v;
    for (int i = 0; i < n; i++) v.insert(v.end(), DBoW2::BowVector::value_type(words[i], vals[i]));
    return v;
}

extern "C" {

void* dbowref_load_text(const char* path)
{
    OrbVocabulary* voc = new OrbVocabulary();
    if (!voc->loadFromTextFile(path)) { delete voc; return nullptr; }
    return voc;
}

void dbowref_free(void* p) { delete static_cast<OrbVocabulary*>(p); }

void dbowref_info(void* p, int* k, int* L, int* nwords, int* scoring, int* weighting)
{
    OrbVocabulary* voc = static_cast<OrbVocabulary*>(p);
    *k = voc->getBranchingFactor(); *L = voc->getDepthLevels(); *nwords = (int)voc->size();
    *scoring = (int)voc->getScoringType(); *weighting = (int)voc->getWeightingType();
}

// transform(features, BowVector&, FeatureVector&, levelsup): the two maps flattened in iteration order
void dbowref_transform(void* p, const uint8_t* desc, int n, int levelsup, uint32_t* words, double* vals, int* nbow,
                       uint32_t* fv_nodes, int32_t* fv_offsets, uint32_t* fv_feats, int* nfv)
{
    OrbVocabulary* voc = static_cast<OrbVocabulary*>(p);
    DBoW2::BowVector v;
    DBoW2::FeatureVector fv;
    voc->transform(to_features(desc, n), v, fv, levelsup);
    int i = 0;
    for (DBoW2::BowVector::const_iterator it = v.begin(); it != v.end(); ++it, ++i) { words[i] = it->first; vals[i] = it->second; }
    *nbow = i;
    int j = 0, o = 0;
    for (DBoW2::FeatureVector::const_iterator it = fv.begin(); it != fv.end(); ++it, ++j) {
        fv_nodes[j] = it->first;
        fv_offsets[j] = o;
        for (size_t q = 0; q < it->second.size(); q++) fv_feats[o++] = it->second[q];
    }
    fv_offsets[j] = o;
    *nfv = j;
}

// transform(features, BowVector&): the same without the feature vector
void dbowref_transform_bow(void* p, const uint8_t* desc, int n, uint32_t* words, double* vals, int* nbow)
{
    OrbVocabulary* voc = static_cast<OrbVocabulary*>(p);
    DBoW2::BowVector v;
    voc->transform(to_features(desc, n), v);
    int i = 0;
    for (DBoW2::BowVector::const_iterator it = v.begin(); it != v.end(); ++it, ++i) { words[i] = it->first; vals[i] = it->second; }
    *nbow = i;
}

// transform(feature) -> word id, getWordWeight, getParentNode
void dbowref_words(void* p, const uint8_t* desc, int n, int levelsup, uint32_t* word, double* weight, uint32_t* parent)
{
    OrbVocabulary* voc = static_cast<OrbVocabulary*>(p);
    std::vector<cv::Mat> f = to_features(desc, n);
    for (int i = 0; i < n; i++) {
        word[i] = voc->transform(f[i]);
        weight[i] = voc->getWordWeight(word[i]);
        parent[i] = voc->getParentNode(word[i], levelsup);
    }
}

double dbowref_score(void* p, const uint32_t* w1, const double* v1, int n1, const uint32_t* w2, const double* v2, int n2)
{
    return static_cast<OrbVocabulary*>(p)->score(to_bow(w1, v1, n1), to_bow(w2, v2, n2));
}

int dbowref_stop_words(void* p, double min_weight) { return static_cast<OrbVocabulary*>(p)->stopWords(min_weight); }

}  // extern "C"
