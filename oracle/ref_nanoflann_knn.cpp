// TEST INFRASTRUCTURE.  Thin C wrapper that compiles the REFERENCE's own vendored kd-tree
// (/root/reference/src/global_fusion/include/Scancontext/nanoflann.hpp, v1.3.2 — a FLANN
// KDTreeSingleIndex derivative) where it lies; nothing from the reference is copied into this repo.
// Output goes to oracle/_ref/libref_nanoflann.so (git-ignored; travels to the GPU box).  It pins the
// oracle's restated kd-tree (orc_pipeline.hpp: KdTree) index-for-index, tie behaviour included.
#include <nanoflann.hpp>
#include <cstddef>

namespace {
struct CloudAdaptor {
  const float* pts;  // stride 4
  size_t n;
  inline size_t kdtree_get_point_count() const { return n; }
  inline float kdtree_get_pt(const size_t idx, const size_t dim) const { return pts[4 * idx + dim]; }
  template <class BBOX> bool kdtree_get_bbox(BBOX&) const { return false; }
};
typedef nanoflann::KDTreeSingleIndexAdaptor<nanoflann::L2_Simple_Adaptor<float, CloudAdaptor>, CloudAdaptor, 3, int> Tree;
}  // namespace

extern "C" void ref_nanoflann_knn(const float* map, int m, const float* q, int nq, int k, int* idx, float* d2) {
  CloudAdaptor ad{map, (size_t)m};
  Tree tree(3, ad, nanoflann::KDTreeSingleIndexAdaptorParams(15));  // leaf 15 = pcl::KdTreeFLANN's KDTreeSingleIndexParams(15)
  tree.buildIndex();
  for (int i = 0; i < nq; ++i) {
    nanoflann::KNNResultSet<float, int> rs(k);
    rs.init(idx + (size_t)k * i, d2 + (size_t)k * i);
    tree.findNeighbors(rs, q + 4 * (size_t)i, nanoflann::SearchParams(32, 0.f, true));
  }
}
