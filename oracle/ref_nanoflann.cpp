// ORACLE/_ref — TEST INFRASTRUCTURE ONLY.
// Thin C wrapper that instantiates the REFERENCE'S OWN vendored kd-tree
// (/root/reference/third_parties/nanoflann/include/nanoflann/nanoflann.hpp, v1.5.0, std-only) exactly the
// way /root/reference/third_parties/nanoflann/include/nanoflann/pcl_adaptor.hpp:12-58 does
// (metric_L2_Simple, Dim = 3, IndexType = size_t, default leaf size 10), minus the PCL types, which are
// absent in this container. It is compiled from the sources where they lie (see Makefile target `ref`);
// nothing from the reference is copied into this repository. Output: oracle/_ref/libref_nanoflann.so.
// Used (a) to cross-check the oracle's kNN restatement and (b) as the kd-tree of the CPU baseline.
#include <nanoflann/nanoflann.hpp>
#include <cstddef>
#include <cstdint>
#include <vector>

namespace {
template <typename Scalar>
struct StridedAdaptor {
  const float* p = nullptr;
  size_t n = 0, stride = 0;
  inline size_t kdtree_get_point_count() const { return n; }
  inline Scalar kdtree_get_pt(const size_t idx, int dim) const { return p[idx * stride + dim]; }
  template <class BBOX> bool kdtree_get_bbox(BBOX&) const { return false; }
};
template <typename Scalar>
struct Tree {
  using Adaptor = StridedAdaptor<Scalar>;
  using metric_t = typename nanoflann::metric_L2_Simple::traits<Scalar, Adaptor>::distance_t;
  using kdtree_t = nanoflann::KDTreeSingleIndexAdaptor<metric_t, Adaptor, 3, size_t>;
  Adaptor adaptor;
  kdtree_t tree;
  Tree() : tree(3, adaptor) {}
  void set(const float* p, size_t n, size_t stride) { adaptor.p = p; adaptor.n = n; adaptor.stride = stride; tree.buildIndex(); }
  size_t knn(Scalar* q, int k, size_t* idx, Scalar* d2) const {
    nanoflann::KNNResultSet<Scalar> rs(k);
    rs.init(idx, d2);
    tree.findNeighbors(rs, q);
    return rs.size();
  }
};
}  // namespace

extern "C" {
void* refnf_create_f64() { return new Tree<double>(); }
void refnf_destroy_f64(void* h) { delete static_cast<Tree<double>*>(h); }
void refnf_build_f64(void* h, const float* pts, size_t n, size_t stride_f) { static_cast<Tree<double>*>(h)->set(pts, n, stride_f); }
size_t refnf_knn_f64(void* h, const double* q, int k, size_t* idx, double* d2) {
  double qq[3] = {q[0], q[1], q[2]};
  return static_cast<Tree<double>*>(h)->knn(qq, k, idx, d2);
}
// batch: idx_out int64 [nq*k] (-1 padded), d2_out [nq*k]
void refnf_knn_batch_f64(const float* pts, size_t n, size_t stride_f, const double* queries, size_t nq, int k, int64_t* idx_out,
                         double* d2_out) {
  Tree<double> t;
  t.set(pts, n, stride_f);
  std::vector<size_t> idx(k);
  std::vector<double> d2(k);
  for (size_t i = 0; i < nq; i++) {
    double q[3] = {queries[i * 3], queries[i * 3 + 1], queries[i * 3 + 2]};
    size_t c = t.knn(q, k, idx.data(), d2.data());
    for (int j = 0; j < k; j++) { idx_out[i * k + j] = size_t(j) < c ? int64_t(idx[j]) : -1; d2_out[i * k + j] = size_t(j) < c ? d2[j] : -1.0; }
  }
}
void refnf_knn_batch_f32(const float* pts, size_t n, size_t stride_f, const double* queries, size_t nq, int k, int64_t* idx_out,
                         double* d2_out) {
  Tree<float> t;
  t.set(pts, n, stride_f);
  std::vector<size_t> idx(k);
  std::vector<float> d2(k);
  for (size_t i = 0; i < nq; i++) {
    float q[3] = {float(queries[i * 3]), float(queries[i * 3 + 1]), float(queries[i * 3 + 2])};
    size_t c = t.knn(q, k, idx.data(), d2.data());
    for (int j = 0; j < k; j++) { idx_out[i * k + j] = size_t(j) < c ? int64_t(idx[j]) : -1; d2_out[i * k + j] = size_t(j) < c ? double(d2[j]) : -1.0; }
  }
}
}
