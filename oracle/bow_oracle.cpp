// oracle/bow_oracle.cpp -- TEST INFRASTRUCTURE (CPU checker), not product code.
//
// Restatement of the DBoW2 descriptor -> word / node descent the reference runs in Frame::ComputeBoW:
//   TemplatedVocabulary::loadFromTextFile   Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1338-1425 (node and word numbering: nodes in
//                                           file order behind the root 0, children in file order, word ids in order of leaf appearance)
//   TemplatedVocabulary::transform(feature, word_id, weight, nid, levelsup)   :1218-1259 (first minimum wins, `d < best_d`)
//   TemplatedVocabulary::transform(features, BowVector, FeatureVector, levelsup)   :1127-1194 for TF_IDF weighting + L1 scoring (the
//                                           ORBvoc.txt configuration): v.addWeight per word, then v.normalize(L1); fv.addFeature(nid, i)
//   FORB::distance                          Thirdparty/DBoW2/DBoW2/FORB.cpp:81-101
// The reference snapshot carries no vocabulary file, so the tests build synthetic trees in the text file's node order.
// Pinned by the reference's OWN vendored DBoW2, compiled unmodified on oracle/vocshim into oracle/_ref/libref_voc.so (oracle/ref_voc_wrap.cpp):
// identical word ids, node ids, feature order and TF-IDF / L1 doubles (tests/test_oracle_voc_vs_ref.py, tests/golden/voc_ref.npz).
#include <math.h>
#include <stdint.h>
#include <string.h>
#include <map>
#include <vector>
#include "oracle.h"

namespace {
int dist256(const uint8_t* a, const uint8_t* b) {
    int d = 0;
    for (int i = 0; i < 32; i++) d += __builtin_popcount((unsigned)(a[i] ^ b[i]));
    return d;
}
struct Voc {
    int L;
    std::vector<std::vector<int> > children;
    std::vector<int> word_id;
    std::vector<char> leaf;
};
// nodes 1..n_nodes-1 in file order: parent[i], is_leaf[i] (entry 0 is the root and is ignored)
Voc build(const int32_t* parent, const uint8_t* is_leaf, int n_nodes, int L) {
    Voc v; v.L = L;
    v.children.resize(n_nodes); v.word_id.assign(n_nodes, 0); v.leaf.assign(n_nodes, 0);
    int nwords = 0;
    for (int i = 1; i < n_nodes; i++) {
        v.children[parent[i]].push_back(i);
        if (is_leaf[i]) { v.word_id[i] = nwords++; v.leaf[i] = 1; }
    }
    return v;
}
}  // namespace

extern "C" {

// per-feature descent: word_id[n], weight[n], node_id[n] (node at level L - levelsup, 0 = root when that level is <= 0)
void oracle_voc_transform(const int32_t* parent, const uint8_t* is_leaf, const uint8_t* node_desc, const double* node_weight, int n_nodes, int L,
                          const uint8_t* feat, int n, int levelsup, int32_t* word_id, double* weight, int32_t* node_id) {
    const Voc v = build(parent, is_leaf, n_nodes, L);
    const int nid_level = L - levelsup;
    for (int f = 0; f < n; f++) {
        int nid = 0, final_id = 0, level = 0;
        do {
            ++level;
            const std::vector<int>& ch = v.children[final_id];
            final_id = ch[0];
            int best = dist256(feat + 32 * (size_t)f, node_desc + 32 * (size_t)final_id);
            for (size_t c = 1; c < ch.size(); c++) {
                const int d = dist256(feat + 32 * (size_t)f, node_desc + 32 * (size_t)ch[c]);
                if (d < best) { best = d; final_id = ch[c]; }
            }
            if (level == nid_level) nid = final_id;
        } while (!v.leaf[final_id]);
        word_id[f] = v.word_id[final_id]; weight[f] = node_weight[final_id]; node_id[f] = nid;
    }
}

// BowVector (TF_IDF + L1) and FeatureVector from the per-feature results: sorted by key like the std::maps of the reference.
// bow_words / bow_values [<= n], fv_nodes [<= n], fv_start [<= n + 1], fv_items [n]; counts2 = {#words, #nodes}
void oracle_voc_vectors(const int32_t* word_id, const double* weight, const int32_t* node_id, int n,
                        int32_t* bow_words, double* bow_values, int32_t* fv_nodes, int32_t* fv_start, int32_t* fv_items, int32_t* counts2) {
    std::map<int, double> bow;
    std::map<int, std::vector<int> > fv;
    for (int i = 0; i < n; i++) {
        if (weight[i] > 0) {                                   // not stopped
            std::map<int, double>::iterator it = bow.lower_bound(word_id[i]);      // BowVector::addWeight
            if (it != bow.end() && !(bow.key_comp()(word_id[i], it->first))) it->second += weight[i];
            else bow.insert(it, std::make_pair(word_id[i], weight[i]));
            fv[node_id[i]].push_back(i);
        }
    }
    double norm = 0.0;                                         // BowVector::normalize(L1)
    for (std::map<int, double>::iterator it = bow.begin(); it != bow.end(); ++it) norm += fabs(it->second);
    if (norm > 0.0) for (std::map<int, double>::iterator it = bow.begin(); it != bow.end(); ++it) it->second /= norm;
    int nb = 0;
    for (std::map<int, double>::iterator it = bow.begin(); it != bow.end(); ++it) { bow_words[nb] = it->first; bow_values[nb] = it->second; nb++; }
    int nn = 0, run = 0;
    for (std::map<int, std::vector<int> >::iterator it = fv.begin(); it != fv.end(); ++it) {
        fv_nodes[nn] = it->first; fv_start[nn] = run;
        for (size_t k = 0; k < it->second.size(); k++) fv_items[run++] = it->second[k];
        nn++;
    }
    fv_start[nn] = run;
    counts2[0] = nb; counts2[1] = nn;
}

}  // extern "C"
