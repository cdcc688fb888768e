// C entry point around the reference's own DBoW2 transform lines (see shim_bow/bow_shim.h and build_bow_ref.sh).
// TEST INFRASTRUCTURE: the checker of orbx_compute_bow, never on the product path.
#include "bow_shim.h"

#include <chrono>

static double g_last_transform_seconds = 0;
// wall time of the last voc.transform() call alone (building the stub's tree from the flat arrays is not the reference's cost)
extern "C" double bowref_last_transform_seconds() { return g_last_transform_seconds; }

namespace {
struct L1Stub : public DBoW2::GeneralScoring {       // L1Scoring's mustNormalize (ScoringObject.h) without ScoringObject.cpp
    virtual double score(const DBoW2::BowVector&, const DBoW2::BowVector&) const { return 0; }
    virtual bool mustNormalize(DBoW2::LNorm& norm) const { norm = DBoW2::L1; return true; }
};
}  // namespace

// Vocabulary as flat arrays: node i has children child_items[child_start[i] .. child_start[i+1]) (in order), a 32-byte
// descriptor, a weight and, for leaves, a word id.  ORBvoc: k = 10, L = 6, TF_IDF weighting, L1 scoring.
// Outputs: the BowVector as (word id, value) pairs in map order, the FeatureVector as (node id, feature index) pairs in map
// / push_back order.  Returns the number of BowVector entries; *n_fv receives the number of FeatureVector pairs.
extern "C" int bowref_transform(int n_nodes, const int* child_start, const int* child_items, const unsigned char* node_desc,
                                const double* node_weight, const int* node_word, int L, int n, const unsigned char* desc, int levelsup,
                                unsigned int* bow_ids, double* bow_values, unsigned int* fv_nodes, unsigned int* fv_features,
                                int* n_fv) {
    using namespace DBoW2;
    TemplatedVocabulary<FORB::TDescriptor, FORB> voc;
    L1Stub scoring;
    voc.m_k = 10; voc.m_L = L; voc.m_weighting = TF_IDF; voc.m_scoring = L1_NORM; voc.m_scoring_object = &scoring;
    voc.m_nodes.resize(n_nodes);
    for (int i = 0; i < n_nodes; ++i) {
        voc.m_nodes[i].id = i;
        voc.m_nodes[i].weight = node_weight[i];
        voc.m_nodes[i].word_id = (WordId)node_word[i];
        memcpy(voc.m_nodes[i].descriptor.bytes, node_desc + 32 * (size_t)i, 32);
        voc.m_nodes[i].children.assign(child_items + child_start[i], child_items + child_start[i + 1]);
        for (int c = child_start[i]; c < child_start[i + 1]; ++c) voc.m_nodes[child_items[c]].parent = i;
    }
    for (int i = 0; i < n_nodes; ++i)
        if (voc.m_nodes[i].isLeaf()) voc.m_words.push_back(&voc.m_nodes[i]);
    std::vector<cv::Mat> feats(n);
    for (int i = 0; i < n; ++i) memcpy(feats[i].bytes, desc + 32 * (size_t)i, 32);
    BowVector v;
    FeatureVector fv;
    const std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    voc.transform(feats, v, fv, levelsup);
    g_last_transform_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    int k = 0;
    for (BowVector::const_iterator it = v.begin(); it != v.end(); ++it, ++k) { bow_ids[k] = it->first; bow_values[k] = it->second; }
    int m = 0;
    for (FeatureVector::const_iterator it = fv.begin(); it != fv.end(); ++it)
        for (size_t j = 0; j < it->second.size(); ++j, ++m) { fv_nodes[m] = it->first; fv_features[m] = it->second[j]; }
    *n_fv = m;
    return k;
}
