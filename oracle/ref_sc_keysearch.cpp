// TEST INFRASTRUCTURE.  Thin C wrapper around the REFERENCE's own ring-key tree, compiled where it lies
// (/root/reference/src/global_fusion/include/Scancontext/{nanoflann.hpp, KDTreeVectorOfVectorsAdaptor.h}) exactly as
// SCManager::detectLoopClosureID builds and queries it (Scancontext.h:229-250: InvKeyTree(PC_NUM_RING, keys, 10 /* max leaf */),
// KNNResultSet<float>(k), findNeighbors(..., SearchParams(10))).  Output: oracle/_ref/libref_sckeys.so (git-ignored).  It pins
// the oracle's brute-force ring-key search (orc_scancontext.hpp) index-for-index.
#include <KDTreeVectorOfVectorsAdaptor.h>
#include <cstddef>
#include <vector>

typedef std::vector<std::vector<float> > KeyMat;
typedef KDTreeVectorOfVectorsAdaptor<KeyMat, float> InvKeyTree;

extern "C" void ref_sc_key_knn(const float* keys, int n, int dim, const float* query, int k, long long* idx, float* d2) {
  KeyMat mat((size_t)n, std::vector<float>((size_t)dim));
  for (int i = 0; i < n; ++i)
    for (int d = 0; d < dim; ++d) mat[i][d] = keys[(size_t)i * dim + d];
  InvKeyTree tree(dim, mat, 10);
  std::vector<size_t> ci((size_t)k, 0);
  std::vector<float> cd((size_t)k, 0.f);
  nanoflann::KNNResultSet<float> rs((size_t)k);
  rs.init(&ci[0], &cd[0]);
  tree.index->findNeighbors(rs, query, nanoflann::SearchParams(10));
  for (int j = 0; j < k; ++j) { idx[j] = (long long)ci[j]; d2[j] = cd[j]; }
}
