// Stub around the reference's OWN lines of DBoW2's bag-of-words transform -- TEST INFRASTRUCTURE (oracle/), not product code.
//
// oracle/build_bow_ref.sh compiles, from where they lie under /root/reference/Thirdparty/DBoW2/DBoW2: BowVector.cpp and
// FeatureVector.cpp whole (they need nothing but their own headers), FORB::distance (FORB.cpp:81-101) and the two
// TemplatedVocabulary::transform members Frame::ComputeBoW reaches (TemplatedVocabulary.h:1126-1194, :1217-1259; call site
// src/Frame.cc:395-402) against this header: a cv::Mat that is a 32-byte row, and the members of TemplatedVocabulary those
// lines touch.  Written from scratch.
#ifndef ORBX_ORACLE_BOW_SHIM_H
#define ORBX_ORACLE_BOW_SHIM_H
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "BowVector.h"           // the reference's headers (-I Thirdparty/DBoW2/DBoW2): typedefs, BowVector, LNorm, WeightingType
#include "FeatureVector.h"
#include "ScoringObject.h"

using namespace std;             // TemplatedVocabulary.h says `vector` unqualified

namespace cv {
struct Mat {                     // one descriptor: 1 x 32 CV_8U
    unsigned char bytes[32];
    int cols;
    Mat() : cols(32) { memset(bytes, 0, 32); }
    template <typename T> const T* ptr() const { return reinterpret_cast<const T*>(bytes); }
};
}  // namespace cv

namespace DBoW2 {

struct FORB {
    typedef cv::Mat TDescriptor;
    static int distance(const TDescriptor& a, const TDescriptor& b);
};

template <class TDescriptor, class F>
class TemplatedVocabulary {
public:
    struct Node {                                   // TemplatedVocabulary.h:297-329, the fields transform() reads
        NodeId id;
        WordValue weight;
        vector<NodeId> children;
        NodeId parent;
        TDescriptor descriptor;
        WordId word_id;
        Node() : id(0), weight(0), parent(0), word_id(0) {}
        inline bool isLeaf() const { return children.empty(); }
    };
    virtual ~TemplatedVocabulary() {}
    virtual inline bool empty() const { return m_words.empty(); }           // (:1009-1012)
    virtual void transform(const std::vector<TDescriptor>& features, BowVector& v, FeatureVector& fv, int levelsup) const;
    virtual void transform(const TDescriptor& feature, WordId& id, WordValue& weight, NodeId* nid = NULL, int levelsup = 0) const;
    int m_k, m_L;
    WeightingType m_weighting;
    ScoringType m_scoring;
    GeneralScoring* m_scoring_object;
    std::vector<Node> m_nodes;
    std::vector<Node*> m_words;
};

}  // namespace DBoW2
#endif
