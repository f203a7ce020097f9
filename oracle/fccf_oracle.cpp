// fccf_oracle.cpp — CPU ORACLE for the FCCF-PCR registration path.
//
// *** TEST INFRASTRUCTURE ONLY. ***  Nothing in the product (fccf_pcr_b200/, include/, the FCCF
// CLI) may include, link or call this file.  Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs use it, and only as the checker / CPU baseline.
//
// *** PARITY UNPINNED. ***  The reference (samsgood0310/FCCF-PCR, FCCF.cpp) has no tests, golden
// vectors or fixtures, and cannot be built here (PCL 1.10 / Eigen 3.3.7 / Ceres 1.14 / FLANN 1.9.1
// are not vendored and not installed).  This file restates FCCF.cpp line by line and restates the
// published algorithms of those libraries from their documented behaviour.
// Every function cites the reference lines (FCCF.cpp:NNN) or the library routine it follows.
//
// Build: g++ -std=c++14 -O3 -ffp-contract=off (no -ffast-math, no -march=native) — the reference's
// CMakeLists.txt:5-10 uses plain -O3, i.e. SSE2 scalar IEEE arithmetic without FMA contraction.
//
// Arithmetic conventions restated from Eigen 3.3 (App. A.4): 3-element reductions (dot, norm,
// 3x3 coefficient products) are evaluated as  a0 + (a1 + a2)  (redux_novec_unroller splits the
// range in halves); 4x4 float products accumulate k = 0..3 sequentially (packet pmadd chain).

#include <algorithm>
#include <cfloat>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

namespace orc {

// ------------------------------------------------------------------------------------------
// parameters: FCCF.cpp:120-176 (file-scope globals), reference defaults
// ------------------------------------------------------------------------------------------
struct Params {
  float parameter_l1 = 0.5f, parameter_l2 = 1.0f, parameter_k1 = 5.0f, parameter_k2 = 2.0f;
  float normal_vector_threshold1 = 5.0f, normal_vector_threshold2 = 8.0f;
  float face_voxel_size = 1.0f;
  float voxel_point_threshold = 5;
  float curvature_threshold = 0.05f;
  float select_plane_number = 15;
  float quick_verify_angel_threshold = 10.0f, quick_verify_distance_threshold = 2.0f;
  float required_optimize_plane = 4.0f;
  float fine_verify_voxel_size = 0.5f;
  float fine_verify_number = 4;
  float included_angle_same_threshold = 5.0f;
  float included_angle_min_threshold = 30.0f, included_angle_max_threshold = 150.0f;
  float third_plane_threshold = 0.5f;
  float third_plane_normal_threshold = 5.0f;
  float cluster_number_threshold = 10;
  float cluster_angel_threshold = 2.0f, cluster_distance_threshold = 0.8f;
  float seclct_cluster_number = 200;
  float rough_threshold_gl = 2;
  int emulate_pcl_overflow = 1;  // VoxelGrid int32 bail-out (App. A.1)
};

// float transcendentals inside pcl::eigen33 (atan2/cos/sin on float arguments).  The reference calls
// the platform's glibc float routines, whose last-ulp behaviour depends on the glibc version and on
// the CPU (ifunc FMA variants; <= 2.40: ~0.56 ulp polynomial kernels, >= 2.41: CORE-MATH, correctly
// rounded).  Default 0: correctly rounded (evaluate in double, round once) — the limit every glibc
// converges to and what the CUDA path computes, so the stages downstream compare bit for bit.
// 1: this platform's atan2f/cosf/sinf (tests bound the effect of the switch on the final transform).
static int g_libm_float = 0;
static inline float f_atan2(float y, float x) { return g_libm_float ? std::atan2(y, x) : (float)std::atan2((double)y, (double)x); }
static inline float f_cos(float x) { return g_libm_float ? std::cos(x) : (float)std::cos((double)x); }
static inline float f_sin(float x) { return g_libm_float ? std::sin(x) : (float)std::sin((double)x); }

struct P3 { float x, y, z; };
struct V3f { float x, y, z; float& operator[](int i){ return (&x)[i]; } float operator[](int i) const { return (&x)[i]; } };
struct V3d { double x, y, z; };
struct M3f { float m[3][3]; };
struct M4f { float m[4][4]; };

// ---- Eigen 3.3 fixed-size float helpers (App. A.4) -------------------------------------------
static inline float sum3(float a, float b, float c) { return a + (b + c); }
static inline double sum3d(double a, double b, double c) { return a + (b + c); }
static inline float dot(const V3f& a, const V3f& b) { return sum3(a.x * b.x, a.y * b.y, a.z * b.z); }
static inline double dotd(const V3d& a, const V3d& b) { return sum3d(a.x * b.x, a.y * b.y, a.z * b.z); }
static inline V3f cross(const V3f& a, const V3f& b) {
  return V3f{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
static inline void normalize(V3f& v) {  // MatrixBase::normalize: divide only if squaredNorm > 0
  float z = dot(v, v);
  if (z > 0.0f) { float n = std::sqrt(z); v.x /= n; v.y /= n; v.z /= n; }
}
static inline M3f identity3() { M3f r; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r.m[i][j] = (i == j) ? 1.f : 0.f; return r; }
static inline M4f identity4() { M4f r; for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) r.m[i][j] = (i == j) ? 1.f : 0.f; return r; }
static inline M3f mul(const M3f& a, const M3f& b) {
  M3f r;
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++)
    r.m[i][j] = sum3(a.m[i][0] * b.m[0][j], a.m[i][1] * b.m[1][j], a.m[i][2] * b.m[2][j]);
  return r;
}
static inline V3f mul(const M3f& a, const V3f& v) {
  V3f r;
  for (int i = 0; i < 3; i++) r[i] = sum3(a.m[i][0] * v.x, a.m[i][1] * v.y, a.m[i][2] * v.z);
  return r;
}
static inline M4f mul(const M4f& a, const M4f& b) {  // 4x4 packet product: sequential k
  M4f r;
  for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) {
    float s = a.m[i][0] * b.m[0][j];
    s = a.m[i][1] * b.m[1][j] + s;
    s = a.m[i][2] * b.m[2][j] + s;
    s = a.m[i][3] * b.m[3][j] + s;
    r.m[i][j] = s;
  }
  return r;
}
static inline M3f transpose(const M3f& a) { M3f r; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r.m[i][j] = a.m[j][i]; return r; }
// Eigen compute_inverse<Matrix3f>: cofactors, det from column 0, multiply by 1/det
static inline float cof(const M3f& m, int i, int j) {
  int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
  return m.m[i1][j1] * m.m[i2][j2] - m.m[i1][j2] * m.m[i2][j1];
}
static inline M3f inverse(const M3f& m) {
  float c0 = cof(m, 0, 0), c1 = cof(m, 1, 0), c2 = cof(m, 2, 0);
  float det = sum3(c0 * m.m[0][0], c1 * m.m[1][0], c2 * m.m[2][0]);
  float invdet = 1.0f / det;
  M3f r;
  r.m[0][0] = c0 * invdet; r.m[0][1] = c1 * invdet; r.m[0][2] = c2 * invdet;
  r.m[1][0] = cof(m, 0, 1) * invdet; r.m[1][1] = cof(m, 1, 1) * invdet; r.m[1][2] = cof(m, 2, 1) * invdet;
  r.m[2][0] = cof(m, 0, 2) * invdet; r.m[2][1] = cof(m, 1, 2) * invdet; r.m[2][2] = cof(m, 2, 2) * invdet;
  return r;
}
struct Quatf { float w, x, y, z; };
// Eigen quaternionbase_assign_impl<Matrix3f>
static inline Quatf quat_from_matrix(const M3f& mat) {
  Quatf q; float c[4];  // c = coeffs (x,y,z,w)
  float t = sum3(mat.m[0][0], mat.m[1][1], mat.m[2][2]);
  if (t > 0.0f) {
    t = std::sqrt(t + 1.0f);
    c[3] = 0.5f * t;
    t = 0.5f / t;
    c[0] = (mat.m[2][1] - mat.m[1][2]) * t;
    c[1] = (mat.m[0][2] - mat.m[2][0]) * t;
    c[2] = (mat.m[1][0] - mat.m[0][1]) * t;
  } else {
    int i = 0;
    if (mat.m[1][1] > mat.m[0][0]) i = 1;
    if (mat.m[2][2] > mat.m[i][i]) i = 2;
    int j = (i + 1) % 3, k = (j + 1) % 3;
    t = std::sqrt(mat.m[i][i] - mat.m[j][j] - mat.m[k][k] + 1.0f);
    c[i] = 0.5f * t;
    t = 0.5f / t;
    c[3] = (mat.m[k][j] - mat.m[j][k]) * t;
    c[j] = (mat.m[j][i] + mat.m[i][j]) * t;
    c[k] = (mat.m[k][i] + mat.m[i][k]) * t;
  }
  q.x = c[0]; q.y = c[1]; q.z = c[2]; q.w = c[3];
  return q;
}
// Eigen QuaternionBase::toRotationMatrix (no normalisation)
static inline M3f quat_to_matrix(const Quatf& q) {
  float tx = 2.f * q.x, ty = 2.f * q.y, tz = 2.f * q.z;
  float twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
  float txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
  float tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
  M3f r;
  r.m[0][0] = 1.f - (tyy + tzz); r.m[0][1] = txy - twz; r.m[0][2] = txz + twy;
  r.m[1][0] = txy + twz; r.m[1][1] = 1.f - (txx + tzz); r.m[1][2] = tyz - twx;
  r.m[2][0] = txz - twy; r.m[2][1] = tyz + twx; r.m[2][2] = 1.f - (txx + tyy);
  return r;
}
// Eigen QuaternionBase::_transformVector
static inline V3f quat_rotate(const Quatf& q, const V3f& v) {
  V3f qv{q.x, q.y, q.z};
  V3f uv = cross(qv, v);
  uv.x += uv.x; uv.y += uv.y; uv.z += uv.z;
  V3f c2 = cross(qv, uv);
  return V3f{(v.x + q.w * uv.x) + c2.x, (v.y + q.w * uv.y) + c2.y, (v.z + q.w * uv.z) + c2.z};
}
// pcl::detail::Transformer<float> (PCL 1.10 common/impl/transforms.hpp, SSE2 path)
static inline V3f tf_se3(const M4f& T, const V3f& p) {
  V3f r;
  for (int i = 0; i < 3; i++) r[i] = p.x * T.m[i][0] + (p.y * T.m[i][1] + (p.z * T.m[i][2] + T.m[i][3]));
  return r;
}
static inline V3f tf_so3(const M4f& T, const V3f& n) {
  V3f r;
  for (int i = 0; i < 3; i++) r[i] = n.x * T.m[i][0] + (n.y * T.m[i][1] + n.z * T.m[i][2]);
  return r;
}

// ------------------------------------------------------------------------------------------
// blobs: named intermediates for stage-wise parity tests
// ------------------------------------------------------------------------------------------
enum { DT_F32 = 0, DT_F64 = 1, DT_I32 = 2, DT_I64 = 3 };
struct Blob { int dtype; std::vector<char> data; };
struct Ctx {
  Params p;
  std::map<std::string, Blob> blobs;
  int keep_blobs = 1;
  double t_pipeline_s = 0, t_total_s = 0;
  // last leftover clouds (for the fine_verify micro-benchmark)
  std::vector<P3> sub1, sub2;
  template <class T> void put(const std::string& name, int dt, const T* p, size_t n) {
    if (!keep_blobs) return;
    Blob b; b.dtype = dt; b.data.resize(n * sizeof(T));
    if (n) memcpy(b.data.data(), p, n * sizeof(T));
    blobs[name] = std::move(b);
  }
  void put_f32(const std::string& n, const std::vector<float>& v) { put(n, DT_F32, v.data(), v.size()); }
  void put_f64(const std::string& n, const std::vector<double>& v) { put(n, DT_F64, v.data(), v.size()); }
  void put_i32(const std::string& n, const std::vector<int>& v) { put(n, DT_I32, v.data(), v.size()); }
  void put_i64(const std::string& n, const std::vector<int64_t>& v) { put(n, DT_I64, v.data(), v.size()); }
};

// ------------------------------------------------------------------------------------------
// pcl::VoxelGrid<PointXYZ>::applyFilter (PCL 1.10 filters/impl/voxel_grid.hpp), App. A.1.
// Call sites FCCF.cpp:1668-1678 and 1377-1387.  Within-cell order is fixed to ascending
// original index (std::sort in PCL is unstable; a stable sort is one of its valid outcomes).
// Non-finite points are skipped (non-dense branch) — this also covers removeNaNFromPointCloud
// (FCCF.cpp:1374-1375) for the bail-out case.
// ------------------------------------------------------------------------------------------
static void voxel_grid(const Params& prm, const std::vector<P3>& in, float leaf, std::vector<P3>& out,
                       std::vector<int64_t>* cell_out, std::vector<int>* cnt_out) {
  out.clear(); if (cell_out) cell_out->clear(); if (cnt_out) cnt_out->clear();
  float inv = 1.0f / leaf;
  float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  size_t nfin = 0;
  for (const P3& p : in) {
    if (!std::isfinite(p.x) || !std::isfinite(p.y) || !std::isfinite(p.z)) continue;
    nfin++;
    mn[0] = std::min(mn[0], p.x); mn[1] = std::min(mn[1], p.y); mn[2] = std::min(mn[2], p.z);
    mx[0] = std::max(mx[0], p.x); mx[1] = std::max(mx[1], p.y); mx[2] = std::max(mx[2], p.z);
  }
  if (nfin == 0) return;
  int64_t dx = (int64_t)((mx[0] - mn[0]) * inv) + 1;
  int64_t dy = (int64_t)((mx[1] - mn[1]) * inv) + 1;
  int64_t dz = (int64_t)((mx[2] - mn[2]) * inv) + 1;
  if (prm.emulate_pcl_overflow && (dx * dy * dz) > (int64_t)INT32_MAX) {
    // "Leaf size is too small for the input dataset" -> output = input
    int64_t k = 0;
    for (const P3& p : in) {
      if (!std::isfinite(p.x) || !std::isfinite(p.y) || !std::isfinite(p.z)) { k++; continue; }
      out.push_back(p);
      if (cell_out) cell_out->push_back(k);
      if (cnt_out) cnt_out->push_back(1);
      k++;
    }
    return;
  }
  int minb[3], maxb[3]; int64_t div[3];
  for (int a = 0; a < 3; a++) {
    minb[a] = (int)std::floor(mn[a] * inv);
    maxb[a] = (int)std::floor(mx[a] * inv);
    div[a] = (int64_t)maxb[a] - minb[a] + 1;
  }
  std::vector<std::pair<int64_t, int>> iv; iv.reserve(nfin);
  for (size_t i = 0; i < in.size(); i++) {
    const P3& p = in[i];
    if (!std::isfinite(p.x) || !std::isfinite(p.y) || !std::isfinite(p.z)) continue;
    int ijk0 = (int)(std::floor(p.x * inv) - (float)minb[0]);
    int ijk1 = (int)(std::floor(p.y * inv) - (float)minb[1]);
    int ijk2 = (int)(std::floor(p.z * inv) - (float)minb[2]);
    int64_t idx = (int64_t)ijk0 + (int64_t)ijk1 * div[0] + (int64_t)ijk2 * div[0] * div[1];
    iv.emplace_back(idx, (int)i);
  }
  std::stable_sort(iv.begin(), iv.end(), [](const std::pair<int64_t, int>& a, const std::pair<int64_t, int>& b) { return a.first < b.first; });
  size_t index = 0;
  while (index < iv.size()) {
    size_t i = index + 1;
    while (i < iv.size() && iv[i].first == iv[index].first) ++i;
    // CentroidPoint<PointXYZ>: float Vector3f running sum, then / n
    float sx = 0, sy = 0, sz = 0;
    for (size_t li = index; li < i; li++) { const P3& p = in[iv[li].second]; sx += p.x; sy += p.y; sz += p.z; }
    float n = (float)(i - index);
    out.push_back(P3{sx / n, sy / n, sz / n});
    if (cell_out) cell_out->push_back(iv[index].first);
    if (cnt_out) cnt_out->push_back((int)(i - index));
    index = i;
  }
}

// ------------------------------------------------------------------------------------------
// pcl::octree::OctreePointCloudSearch restated (PCL 1.10 octree_pointcloud.hpp), App. A.2:
// adoptBoundingBoxToPoint / getKeyBitSize / genOctreeKeyforPoint / getOccupiedVoxelCenters /
// voxelSearch.  Used at FCCF.cpp:475-484 (res 1.0) and 792-805 (res 0.5).
// ------------------------------------------------------------------------------------------
struct Octree {
  double res; double mn[3], mx[3]; int depth = 0; bool defined = false;
  // result: voxels in getOccupiedVoxelCenters (DFS) order
  std::vector<uint32_t> key;       // 3 per voxel, final absolute keys
  std::vector<int> start;          // V+1 offsets into pidx
  std::vector<int> pidx;           // point indices, ascending inside a voxel
  uint32_t key0[3];                // final key of the first inserted point (lattice anchor)
};
static void oct_adopt(Octree& o, const P3& p) {
  const float minValue = std::numeric_limits<float>::epsilon();
  const float pp[3] = {p.x, p.y, p.z};
  while (true) {
    bool lo[3], up[3]; bool any = false;
    for (int a = 0; a < 3; a++) { lo[a] = (pp[a] < o.mn[a]); up[a] = (pp[a] >= o.mx[a]); any = any || lo[a] || up[a]; }
    if (any || !o.defined) {
      if (o.defined) {
        double side = (double)(1u << o.depth) * o.res;
        for (int a = 0; a < 3; a++) if (!up[a]) o.mn[a] -= side;
        o.depth++;
        side = (double)(1u << o.depth) * o.res - minValue;
        for (int a = 0; a < 3; a++) o.mx[a] = o.mn[a] + side;
      } else {
        for (int a = 0; a < 3; a++) { o.mn[a] = pp[a] - o.res / 2; o.mx[a] = pp[a] + o.res / 2; }
        // getKeyBitSize() with leaf_count_ == 0
        unsigned mk[3];
        for (int a = 0; a < 3; a++) mk[a] = (unsigned)std::ceil((o.mx[a] - o.mn[a] - minValue) / o.res);
        unsigned maxv = std::max(std::max(std::max(mk[0], mk[1]), mk[2]), 2u);
        o.depth = (int)std::max(std::min(32u, (unsigned)std::ceil(std::log2((double)maxv) - minValue)), 0u);
        double side = (double)(1u << o.depth) * o.res;
        for (int a = 0; a < 3; a++) {
          double over = (side - (o.mx[a] - o.mn[a])) / 2.0;
          if (over > minValue) { o.mn[a] -= over; o.mx[a] += over; }
        }
        o.defined = true;
      }
    } else break;
  }
}
static void octree_build(Octree& o, const std::vector<P3>& pts, double res) {
  o.res = res; o.defined = false; o.depth = 0;
  o.key.clear(); o.start.clear(); o.pidx.clear();
  for (const P3& p : pts) oct_adopt(o, p);  // all points finite here
  size_t n = pts.size();
  std::vector<std::pair<uint64_t, int>> mv(n);
  std::vector<uint32_t> k3(3 * n);
  for (size_t i = 0; i < n; i++) {
    uint32_t kx = (uint32_t)(((double)pts[i].x - o.mn[0]) / o.res);
    uint32_t ky = (uint32_t)(((double)pts[i].y - o.mn[1]) / o.res);
    uint32_t kz = (uint32_t)(((double)pts[i].z - o.mn[2]) / o.res);
    k3[3 * i] = kx; k3[3 * i + 1] = ky; k3[3 * i + 2] = kz;
    uint64_t code = 0;  // child index = (xbit<<2)|(ybit<<1)|zbit, MSB first
    for (int b = o.depth - 1; b >= 0; b--)
      code = (code << 3) | (uint64_t)((((kx >> b) & 1u) << 2) | (((ky >> b) & 1u) << 1) | ((kz >> b) & 1u));
    mv[i] = std::make_pair(code, (int)i);
  }
  if (n) { o.key0[0] = k3[0]; o.key0[1] = k3[1]; o.key0[2] = k3[2]; }
  std::stable_sort(mv.begin(), mv.end(), [](const std::pair<uint64_t, int>& a, const std::pair<uint64_t, int>& b) { return a.first < b.first; });
  size_t i = 0;
  while (i < n) {
    size_t j = i;
    o.start.push_back((int)o.pidx.size());
    int f = mv[i].second;
    o.key.push_back(k3[3 * f]); o.key.push_back(k3[3 * f + 1]); o.key.push_back(k3[3 * f + 2]);
    while (j < n && mv[j].first == mv[i].first) { o.pidx.push_back(mv[j].second); j++; }
    i = j;
  }
  o.start.push_back((int)o.pidx.size());
}

// ------------------------------------------------------------------------------------------
// pcl::eigen33 (smallest eigenpair) + computeRoots/computeRoots2 (PCL 1.10 common/impl/eigen.hpp),
// pcl::solvePlaneParameters, computeMeanAndCovarianceMatrix (single-pass raw moments) — App. A.3.
// ------------------------------------------------------------------------------------------
static void compute_roots2(float b, float c, float roots[3]) {
  roots[0] = 0.f;
  float d = (float)((double)(b * b) - 4.0 * (double)c);
  if (d < 0.0) d = 0.0f;
  float sd = std::sqrt(d);
  roots[2] = 0.5f * (b + sd);
  roots[1] = 0.5f * (b - sd);
}
static void compute_roots(const float m[3][3], float roots[3]) {
  float c0 = m[0][0] * m[1][1] * m[2][2] + 2.f * m[0][1] * m[0][2] * m[1][2] - m[0][0] * m[1][2] * m[1][2]
           - m[1][1] * m[0][2] * m[0][2] - m[2][2] * m[0][1] * m[0][1];
  float c1 = m[0][0] * m[1][1] - m[0][1] * m[0][1] + m[0][0] * m[2][2] - m[0][2] * m[0][2] + m[1][1] * m[2][2] - m[1][2] * m[1][2];
  float c2 = m[0][0] + m[1][1] + m[2][2];
  if (std::fabs(c0) < FLT_EPSILON) { compute_roots2(c2, c1, roots); return; }
  const float s_inv3 = (float)(1.0 / 3.0);
  const float s_sqrt3 = std::sqrt(3.0f);
  float c2_over_3 = c2 * s_inv3;
  float a_over_3 = (c1 - c2 * c2_over_3) * s_inv3;
  if (a_over_3 > 0.f) a_over_3 = 0.f;
  float half_b = 0.5f * (c0 + c2_over_3 * (2.f * c2_over_3 * c2_over_3 - c1));
  float q = half_b * half_b + a_over_3 * a_over_3 * a_over_3;
  if (q > 0.f) q = 0.f;
  float rho = std::sqrt(-a_over_3);
  float theta = f_atan2(std::sqrt(-q), half_b) * s_inv3;
  float cos_theta = f_cos(theta);
  float sin_theta = f_sin(theta);
  roots[0] = c2_over_3 + 2.f * rho * cos_theta;
  roots[1] = c2_over_3 - rho * (cos_theta + s_sqrt3 * sin_theta);
  roots[2] = c2_over_3 - rho * (cos_theta - s_sqrt3 * sin_theta);
  if (roots[0] >= roots[1]) std::swap(roots[0], roots[1]);
  if (roots[1] >= roots[2]) { std::swap(roots[1], roots[2]); if (roots[0] >= roots[1]) std::swap(roots[0], roots[1]); }
  if (roots[0] <= 0) compute_roots2(c2, c1, roots);
}
static void eigen33_smallest(const float mat[3][3], float& eigenvalue, V3f& evec) {
  float scale = 0.f;
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) scale = std::max(scale, std::fabs(mat[i][j]));
  if (scale <= FLT_MIN) scale = 1.0f;
  float s[3][3];
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) s[i][j] = mat[i][j] / scale;
  float roots[3];
  compute_roots(s, roots);
  eigenvalue = roots[0] * scale;
  s[0][0] -= roots[0]; s[1][1] -= roots[0]; s[2][2] -= roots[0];
  V3f r0{s[0][0], s[0][1], s[0][2]}, r1{s[1][0], s[1][1], s[1][2]}, r2{s[2][0], s[2][1], s[2][2]};
  V3f v1 = cross(r0, r1), v2 = cross(r0, r2), v3 = cross(r1, r2);
  float l1 = dot(v1, v1), l2 = dot(v2, v2), l3 = dot(v3, v3);
  if (l1 >= l2 && l1 >= l3) { float d = std::sqrt(l1); evec = V3f{v1.x / d, v1.y / d, v1.z / d}; }
  else if (l2 >= l1 && l2 >= l3) { float d = std::sqrt(l2); evec = V3f{v2.x / d, v2.y / d, v2.z / d}; }
  else { float d = std::sqrt(l3); evec = V3f{v3.x / d, v3.y / d, v3.z / d}; }
}
// compute3DCentroid(cloud, indices) + NormalEstimation::computePointNormal(cloud, indices, ...)
static void plane_fit(const std::vector<P3>& cloud, const int* idx, int n, float centroid[3], float normal[3], float& curvature) {
  float accu[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (int k = 0; k < n; k++) {
    const P3& p = cloud[idx[k]];
    accu[0] += p.x * p.x; accu[1] += p.x * p.y; accu[2] += p.x * p.z;
    accu[3] += p.y * p.y; accu[4] += p.y * p.z; accu[5] += p.z * p.z;
    accu[6] += p.x; accu[7] += p.y; accu[8] += p.z;
  }
  float fn = (float)n;
  for (int k = 0; k < 9; k++) accu[k] /= fn;
  centroid[0] = accu[6]; centroid[1] = accu[7]; centroid[2] = accu[8];
  float c[3][3];
  c[0][0] = accu[0] - accu[6] * accu[6];
  c[0][1] = accu[1] - accu[6] * accu[7];
  c[0][2] = accu[2] - accu[6] * accu[8];
  c[1][1] = accu[3] - accu[7] * accu[7];
  c[1][2] = accu[4] - accu[7] * accu[8];
  c[2][2] = accu[5] - accu[8] * accu[8];
  c[1][0] = c[0][1]; c[2][0] = c[0][2]; c[2][1] = c[1][2];
  float ev; V3f evec;
  eigen33_smallest(c, ev, evec);
  normal[0] = evec.x; normal[1] = evec.y; normal[2] = evec.z;
  float eig_sum = c[0][0] + c[1][1] + c[2][2];
  curvature = (eig_sum != 0) ? std::fabs(ev / eig_sum) : 0.f;
}

// ------------------------------------------------------------------------------------------
// FCCF.cpp:369-407
// ------------------------------------------------------------------------------------------
static float compute_normal_angel(float x1, float y1, float z1, float x2, float y2, float z2) {
  V3d n1{x1, y1, z1}, n2{x2, y2, z2};
  float n1n3 = (float)dotd(n1, n2);
  float cos_theta = (float)((double)n1n3 / (std::sqrt(dotd(n1, n1)) * std::sqrt(dotd(n2, n2))));
  float theta = (float)(std::acos((double)cos_theta) * 180 / M_PI);  // double acos, rounded to float (§8c)
  return theta;
}
static bool compare_normal(float x1, float y1, float z1, float x2, float y2, float z2, float thr) {
  float theta = compute_normal_angel(x1, y1, z1, x2, y2, z2);
  if (theta > thr) return false; else return true;
}
static bool compare_plane(float nx1, float ny1, float nz1, float cx1, float cy1, float cz1,
                          float nx2, float ny2, float nz2, float cx2, float cy2, float cz2, float l, float k) {
  V3d n1{nx1, ny1, nz1}, n2{nx2, ny2, nz2};
  float vl = std::sqrt((cx1 - cx2) * (cx1 - cx2) + (cy1 - cy2) * (cy1 - cy2) + (cz1 - cz2) * (cz1 - cz2));
  V3d n3{(cx1 - cx2) / vl, (cy1 - cy2) / vl, (cz1 - cz2) / vl};
  float n1n3 = (float)std::fabs(dotd(n1, n3));
  float n2n3 = (float)std::fabs(dotd(n2, n3));
  float thr = l / (k * vl + 1);
  return (n1n3 < thr && n2n3 < thr);
}

struct VoxelNode { float cx, cy, cz, nx, ny, nz; int size; bool alloc; int id; };
struct FaceNode {
  float cx, cy, cz, nx, ny, nz; float size; bool alloc;
  std::vector<VoxelNode> vox;
  int id;  // stage-1 creation index
};
static void face_average(FaceNode& f) {  // FCCF.cpp:563-586 / 619-642 (recomputed from scratch)
  float s = 0, ax = 0, ay = 0, az = 0, bx = 0, by = 0, bz = 0;
  for (const VoxelNode& v : f.vox) {
    s = s + v.size;
    ax = ax + v.cx * v.size; ay = ay + v.cy * v.size; az = az + v.cz * v.size;
    bx = bx + v.nx * v.size; by = by + v.ny * v.size; bz = bz + v.nz * v.size;
  }
  f.size = s; f.cx = ax / s; f.cy = ay / s; f.cz = az / s; f.nx = bx / s; f.ny = by / s; f.nz = bz / s;
}
// FCCF.cpp:409-427 (exchange sort, descending by number of voxels)
static void range_face(std::vector<FaceNode>& fv) {
  for (size_t i = 0; i + 1 < fv.size(); i++)
    for (size_t j = i + 1; j < fv.size(); j++)
      if (fv[i].vox.size() < fv[j].vox.size()) std::swap(fv[i], fv[j]);
}

// FCCF.cpp:470-678
static void face_extrate(Ctx& C, const std::string& tag, const std::vector<P3>& cloud, std::vector<FaceNode>& faces,
                         std::vector<P3>& sub, std::vector<double>& theta_vec) {
  const Params& P = C.p;
  // compute3DCentroid(*cloud_src): float running sums in index order
  float cc[3] = {0, 0, 0};
  for (const P3& p : cloud) { cc[0] += p.x; cc[1] += p.y; cc[2] += p.z; }
  float fn = (float)cloud.size();
  cc[0] /= fn; cc[1] /= fn; cc[2] /= fn;
  Octree oct; octree_build(oct, cloud, (double)P.face_voxel_size);
  int V = (int)oct.start.size() - 1; if (V < 0) V = 0;
  std::vector<VoxelNode> vv;
  std::vector<int> vflag(V, 0), vcnt(V, 0);
  std::vector<float> vplane(V * 8, 0.f);
  sub.clear();
  for (int v = 0; v < V; v++) {
    int n = oct.start[v + 1] - oct.start[v];
    vcnt[v] = n;
    if ((float)n > P.voxel_point_threshold) {
      float cen[3], nrm[3], curv;
      plane_fit(cloud, &oct.pidx[oct.start[v]], n, cen, nrm, curv);
      vplane[v * 8 + 0] = cen[0]; vplane[v * 8 + 1] = cen[1]; vplane[v * 8 + 2] = cen[2];
      vplane[v * 8 + 6] = curv; vplane[v * 8 + 7] = (float)n;
      if (curv < P.curvature_threshold) {
        VoxelNode t; t.cx = cen[0]; t.cy = cen[1]; t.cz = cen[2]; t.size = n;
        V3f to{cen[0] - cc[0], cen[1] - cc[1], cen[2] - cc[2]}; V3f nv{nrm[0], nrm[1], nrm[2]};
        if (dot(to, nv) < 0) { t.nx = nrm[0]; t.ny = nrm[1]; t.nz = nrm[2]; }
        else { t.nx = -nrm[0]; t.ny = -nrm[1]; t.nz = -nrm[2]; }
        t.alloc = false; t.id = (int)vv.size();
        vplane[v * 8 + 3] = t.nx; vplane[v * 8 + 4] = t.ny; vplane[v * 8 + 5] = t.nz;
        vv.push_back(t); vflag[v] = 1;
      } else {
        vplane[v * 8 + 3] = nrm[0]; vplane[v * 8 + 4] = nrm[1]; vplane[v * 8 + 5] = nrm[2];
        for (int k = oct.start[v]; k < oct.start[v + 1]; k++) sub.push_back(cloud[oct.pidx[k]]);
        vflag[v] = 2;
      }
    }
  }
  if (C.keep_blobs) {
    std::vector<float> ccv(cc, cc + 3); C.put_f32("cloud_centroid" + tag, ccv);
    std::vector<double> om(oct.mn, oct.mn + 3); C.put_f64("oct_min" + tag, om);
    std::vector<int> od(1, oct.depth); C.put_i32("oct_depth" + tag, od);
    std::vector<int> k(oct.key.begin(), oct.key.end()); C.put_i32("vox_key" + tag, k);
    C.put_i32("vox_cnt" + tag, vcnt); C.put_i32("vox_flag" + tag, vflag); C.put_f32("vox_plane" + tag, vplane);
    C.put_i32("vox_pidx" + tag, oct.pidx);
    std::vector<float> sb; for (auto& p : sub) { sb.push_back(p.x); sb.push_back(p.y); sb.push_back(p.z); }
    C.put_f32("sub" + tag, sb);
    std::vector<float> pv; for (auto& t : vv) { pv.push_back(t.cx); pv.push_back(t.cy); pv.push_back(t.cz); pv.push_back(t.nx); pv.push_back(t.ny); pv.push_back(t.nz); pv.push_back((float)t.size); }
    C.put_f32("pvox" + tag, pv);
  }
  // stage 1: FCCF.cpp:536-593
  std::vector<FaceNode> grow;
  for (size_t i1 = 0; i1 < vv.size(); i1++) {
    if (vv[i1].alloc == false) {
      FaceNode f; vv[i1].alloc = true; f.vox.push_back(vv[i1]);
      f.size = vv[i1].size; f.nx = vv[i1].nx; f.ny = vv[i1].ny; f.nz = vv[i1].nz; f.cx = vv[i1].cx; f.cy = vv[i1].cy; f.cz = vv[i1].cz;
      for (size_t i2 = 0; i2 < vv.size(); i2++) {
        if (vv[i2].alloc == false) {
          bool same = compare_normal(f.nx, f.ny, f.nz, vv[i2].nx, vv[i2].ny, vv[i2].nz, P.normal_vector_threshold1);
          bool cop = compare_plane(f.nx, f.ny, f.nz, f.cx, f.cy, f.cz, vv[i2].nx, vv[i2].ny, vv[i2].nz, vv[i2].cx, vv[i2].cy, vv[i2].cz, P.parameter_l1, P.parameter_k1);
          if (same && cop) { f.vox.push_back(vv[i2]); vv[i2].alloc = true; face_average(f); }
        }
      }
      f.alloc = false; f.id = (int)grow.size();
      grow.push_back(f);
    }
  }
  if (C.keep_blobs) {
    std::vector<int> lab(vv.size(), -1);
    for (auto& f : grow) for (auto& v : f.vox) lab[v.id] = f.id;
    C.put_i32("grow_label" + tag, lab);
  }
  // stage 2: FCCF.cpp:595-648
  for (size_t i1 = 0; i1 < grow.size(); i1++) {
    if (grow[i1].alloc == false) {
      bool newadd = true;
      while (newadd) {
        newadd = false;
        for (size_t i2 = 0; i2 < grow.size(); i2++) {
          if (i2 != i1 && grow[i2].alloc == false) {
            FaceNode& a = grow[i1]; FaceNode& b = grow[i2];
            bool same = compare_normal(a.nx, a.ny, a.nz, b.nx, b.ny, b.nz, P.normal_vector_threshold2);
            bool cop = compare_plane(a.nx, a.ny, a.nz, a.cx, a.cy, a.cz, b.nx, b.ny, b.nz, b.cx, b.cy, b.cz, P.parameter_l2, P.parameter_k2);
            if (same && cop) {
              newadd = true; b.alloc = true;
              for (auto& v : b.vox) a.vox.push_back(v);
              face_average(a);
            }
          }
        }
      }
    }
  }
  if (C.keep_blobs) {
    std::vector<int> lab(vv.size(), -1);
    for (auto& f : grow) if (!f.alloc) for (auto& v : f.vox) lab[v.id] = f.id;
    C.put_i32("merge_label" + tag, lab);
  }
  range_face(grow);  // FCCF.cpp:650
  // FCCF.cpp:652-677
  std::vector<FaceNode> chose; theta_vec.clear();
  int sel = 0;
  for (size_t i1 = 0; i1 < grow.size(); i1++) {
    if (grow[i1].alloc == false) {
      chose.push_back(grow[i1]);
      double ts = 0; const FaceNode& f = grow[i1];
      for (size_t i = 0; i < f.vox.size(); i++) {
        double th = compute_normal_angel(f.nx, f.ny, f.nz, f.vox[i].nx, f.vox[i].ny, f.vox[i].nz);
        ts += std::fabs(th);
      }
      ts /= f.vox.size();
      theta_vec.push_back(ts);
      sel++;
    }
    if ((float)sel > P.select_plane_number) break;
  }
  faces.swap(chose);
  if (C.keep_blobs) {
    std::vector<float> fp; std::vector<int> fid, fnv, fvox, foff;
    for (auto& f : faces) {
      fp.push_back(f.cx); fp.push_back(f.cy); fp.push_back(f.cz); fp.push_back(f.nx); fp.push_back(f.ny); fp.push_back(f.nz); fp.push_back(f.size);
      fid.push_back(f.id); fnv.push_back((int)f.vox.size()); foff.push_back((int)fvox.size());
      for (auto& v : f.vox) fvox.push_back(v.id);
    }
    foff.push_back((int)fvox.size());
    C.put_f32("face_plane" + tag, fp); C.put_i32("face_id" + tag, fid); C.put_i32("face_nvox" + tag, fnv);
    C.put_i32("face_vox" + tag, fvox); C.put_i32("face_off" + tag, foff);
    C.put_f64("face_theta" + tag, theta_vec);
    std::vector<int> nf(1, (int)grow.size()); C.put_i32("n_stage1_faces" + tag, nf);
  }
}

struct FaceBase { int i1, i2; float angel; };
// FCCF.cpp:429-468.  A NaN roughness matches none of the four branches in the reference (the type
// list then goes out of step with the pair list, which is UB downstream); here type 3 = "none".
static void select_base(const Params& P, std::vector<FaceBase>& base, const std::vector<FaceNode>& faces, std::vector<int>& type_index, const std::vector<double>& theta) {
  float tmin = P.included_angle_min_threshold, tmax = P.included_angle_max_threshold;
  double th1 = P.rough_threshold_gl;
  for (int a = 0; a < (int)faces.size(); a++)
    for (int b = 0; b < (int)faces.size(); b++)
      if (a < b) {
        float angel = compute_normal_angel(faces[a].nx, faces[a].ny, faces[a].nz, faces[b].nx, faces[b].ny, faces[b].nz);
        if (tmin < angel && angel < tmax) {
          base.push_back(FaceBase{a, b, angel});
          if (theta[a] <= th1 && theta[b] <= th1) type_index.push_back(0);
          else if (theta[a] > th1 && theta[b] > th1) type_index.push_back(1);
          else if (theta[a] <= th1 && theta[b] > th1) type_index.push_back(2);
          else if (theta[a] > th1 && theta[b] <= th1) type_index.push_back(2);
          else type_index.push_back(3);
        }
      }
}

static M3f rodrigues(float c, float s, const V3f& r) {  // cos*I + (1-cos)*r r^T + sin*[r]x (FCCF.cpp:850-868)
  M3f I = identity3(), R;
  float rx[3][3] = {{0, -r.z, r.y}, {r.z, 0, -r.x}, {-r.y, r.x, 0}};
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++)
    R.m[i][j] = (c * I.m[i][j] + (1 - c) * (r[i] * r[j])) + s * rx[i][j];
  return R;
}

// FCCF.cpp:841-1018
static void computer_transform(const Params& P, std::vector<std::vector<M4f>>& tv, int index11, int index12, int index21, int index22,
                               const std::vector<FaceNode>& f1, const std::vector<FaceNode>& f2, int type) {
  M4f T = identity4();
  V3f n1{f1[index11].nx, f1[index11].ny, f1[index11].nz}, m1{f1[index12].nx, f1[index12].ny, f1[index12].nz};
  V3f n2{f2[index21].nx, f2[index21].ny, f2[index21].nz}, m2{f2[index22].nx, f2[index22].ny, f2[index22].nz};
  V3f r1 = cross(n2, n1); normalize(r1);
  float n2dn1 = dot(n2, n1);
  V3f r1cn2 = cross(r1, n2);
  float r1cn2dn1 = dot(r1cn2, n1);
  M3f R1 = rodrigues(n2dn1, r1cn2dn1, r1);
  m2 = mul(R1, m2);
  V3f r2 = n1;
  float m2dm1 = dot(m2, m1), m2dr2 = dot(m2, r2), m1dr2 = dot(m1, r2);
  V3f r2cm2 = cross(r2, m2);
  float r2cm2dm1 = dot(r2cm2, m1);
  float cos2 = (m2dm1 - (m2dr2 * m1dr2)) / (1 - (m2dr2 * m1dr2));
  float sin2 = (r2cm2dm1) / (1 - (m2dr2 * m1dr2));
  M3f R2 = rodrigues(cos2, sin2, r2);
  M3f rot = mul(R2, R1);
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) T.m[i][j] = rot.m[i][j];
  std::vector<int> three1;
  V3f n1cm1 = cross(n1, m1); normalize(n1cm1);
  float chose = P.third_plane_threshold;
  for (int k = 0; k < (int)f1.size(); k++)
    if (k != index11 && k != index12) {
      V3f nt{f1[k].nx, f1[k].ny, f1[k].nz};
      if (std::fabs(dot(n1cm1, nt)) > chose) three1.push_back(k);
    }
  V3f n2cm2 = cross(n2, m2); normalize(n2cm2);
  bool getthree = false;
  if (!three1.empty()) {
    std::vector<V3f> cp(f2.size()), cn(f2.size());
    for (size_t k = 0; k < f2.size(); k++) {  // transformPointCloudWithNormals, FCCF.cpp:948
      cp[k] = tf_se3(T, V3f{f2[k].cx, f2[k].cy, f2[k].cz});
      cn[k] = tf_so3(T, V3f{f2[k].nx, f2[k].ny, f2[k].nz});
    }
    float angthr = P.third_plane_normal_threshold;
    for (int k3 : three1) {
      for (int k2 = 0; k2 < (int)f2.size(); k2++) {
        if (k2 != index21 && k2 != index22) {
          float a3 = compute_normal_angel(f1[k3].nx, f1[k3].ny, f1[k3].nz, cn[k2].x, cn[k2].y, cn[k2].z);
          if (a3 < angthr && std::fabs(dot(n2cm2, cn[k2])) > chose) {
            getthree = true;
            V3f k1{f1[k3].nx, f1[k3].ny, f1[k3].nz}, kk2 = cn[k2];
            V3f c11{f1[index11].cx, f1[index11].cy, f1[index11].cz}, c12{f1[index12].cx, f1[index12].cy, f1[index12].cz}, c13{f1[k3].cx, f1[k3].cy, f1[k3].cz};
            V3f c21{f2[index21].cx, f2[index21].cy, f2[index21].cz}, c22{f2[index22].cx, f2[index22].cy, f2[index22].cz}, c23 = cp[k2];
            float d11 = dot(c11, n1), d12 = dot(c12, m1), d13 = dot(c13, k1);
            float d21 = dot(c21, n2), d22 = dot(c22, m2), d23 = dot(c23, kk2);
            V3f D{d11 - d21, d12 - d22, d13 - d23};
            M3f A; A.m[0][0] = n1.x; A.m[0][1] = n1.y; A.m[0][2] = n1.z; A.m[1][0] = m1.x; A.m[1][1] = m1.y; A.m[1][2] = m1.z; A.m[2][0] = k1.x; A.m[2][1] = k1.y; A.m[2][2] = k1.z;
            M3f AT = transpose(A);
            V3f Tt = mul(mul(inverse(mul(AT, A)), AT), D);
            T.m[0][3] = Tt.x; T.m[1][3] = Tt.y; T.m[2][3] = Tt.z;
            tv[type].push_back(T);
          }
        }
      }
    }
  }
  if (!getthree) {
    const FaceNode &a = f1[index11], &b = f1[index12], &c = f2[index21], &d = f2[index22];
    float sx = (a.cx * a.size + b.cx * b.size) / (a.size + b.size);
    float sy = (a.cy * a.size + b.cy * b.size) / (a.size + b.size);
    float sz = (a.cz * a.size + b.cz * b.size) / (a.size + b.size);
    float tx = (c.cx * c.size + d.cx * d.size) / (c.size + d.size);
    float ty = (c.cy * c.size + d.cy * d.size) / (c.size + d.size);
    float tz = (c.cz * c.size + d.cz * d.size) / (c.size + d.size);
    V3f tc = mul(rot, V3f{tx, ty, tz});
    T.m[0][3] = sx - tc.x; T.m[1][3] = sy - tc.y; T.m[2][3] = sz - tc.z;
    tv[type].push_back(T);
  }
}

struct QT { float qw, qx, qy, qz, tx, ty, tz; bool alloc; };

// two-axis rotation construction shared by FCCF.cpp:1148-1196 and 1306-1354
static M3f rotation_from_axes(V3f nt1, V3f nt2) {
  V3f ns1{1, 0, 0}, ns2{0, 1, 0};
  V3f r1 = cross(ns1, nt1); normalize(r1);
  float c1 = dot(nt1, ns1);
  float s1 = dot(nt1, cross(r1, ns1));
  M3f R1 = rodrigues(c1, s1, r1);
  ns2 = mul(R1, ns2);
  V3f r2 = nt1;
  float ns2dnt2 = dot(ns2, nt2), ns2dr2 = dot(ns2, r2), nt2dr2 = dot(nt2, r2);
  V3f r2cns2 = cross(r2, ns2);
  float r2cns2dnt2 = dot(r2cns2, nt2);
  float c2 = (ns2dnt2 - (ns2dr2 * nt2dr2)) / (1 - (ns2dr2 * nt2dr2));
  float s2 = (r2cns2dnt2) / (1 - (ns2dr2 * nt2dr2));
  M3f R2 = rodrigues(c2, s2, r2);
  return mul(R2, R1);
}

// FCCF.cpp:1040-1231.  Neighbour search = pcl::KdTreeFLANN::radiusSearch (App. A.6): squared L2
// (flann::L2_Simple, sequential float), strictly < float(radius*radius), sorted by (dist, index).
static void transform_cluster(Ctx& C, const std::string& tag, std::vector<QT>& v, std::vector<QT>& fine, int cluster_num) {
  const Params& P = C.p;
  int n = (int)v.size();
  if ((float)n <= P.cluster_number_threshold) {
    if (n == 0) fine.push_back(QT{1, 0, 0, 0, 0, 0, 0, true});
    else for (auto& q : v) fine.push_back(q);
    return;
  }
  float dq_thr = P.cluster_angel_threshold;
  double rad = (double)P.cluster_distance_threshold;
  float r2 = (float)(rad * rad);
  std::vector<std::vector<QT>> clusters;
  std::vector<int> seed_ids;
  // x-sorted index for a windowed exact search (any exact search returns the same set)
  auto xkey = [&](int a) { double x = (double)v[a].tx; return (x == x) ? x : (double)INFINITY; };
  std::vector<int> order(n); for (int i = 0; i < n; i++) order[i] = i;
  std::sort(order.begin(), order.end(), [&](int a, int b) { double xa = xkey(a), xb = xkey(b); return xa < xb || (xa == xb && a < b); });
  std::vector<double> xs(n); for (int i = 0; i < n; i++) xs[i] = xkey(order[i]);
  double rr = rad * 1.0001 + 1e-6;
  for (int i1 = 0; i1 < n; i1++) {
    if (i1 != n - 1) {
      if (v[i1].alloc == false) {
        std::vector<QT> cl;
        std::vector<std::pair<float, int>> nb;
        bool finite_x = std::isfinite(v[i1].tx);
        size_t lo = 0, hi = (size_t)n;
        if (finite_x) {
          lo = std::lower_bound(xs.begin(), xs.end(), (double)v[i1].tx - rr) - xs.begin();
          hi = std::upper_bound(xs.begin(), xs.end(), (double)v[i1].tx + rr) - xs.begin();
        }
        for (size_t k = lo; k < hi; k++) {
          int j = order[k];
          float d0 = v[i1].tx - v[j].tx, d1 = v[i1].ty - v[j].ty, d2 = v[i1].tz - v[j].tz;
          float d = 0.f; d += d0 * d0; d += d1 * d1; d += d2 * d2;
          if (d < r2) nb.emplace_back(d, j);
        }
        std::sort(nb.begin(), nb.end());
        Quatf Q1{v[i1].qw, v[i1].qx, v[i1].qy, v[i1].qz};
        V3f p1 = quat_rotate(Q1, V3f{1, 0, 0});
        for (auto& pr : nb) {
          int j = pr.second;
          Quatf Q2{v[j].qw, v[j].qx, v[j].qy, v[j].qz};
          V3f p2 = quat_rotate(Q2, V3f{1, 0, 0});
          float dq = compute_normal_angel(p1.x, p1.y, p1.z, p2.x, p2.y, p2.z);
          if (dq < dq_thr) { v[j].alloc = true; cl.push_back(v[j]); }
        }
        clusters.push_back(cl); seed_ids.push_back(i1);
      }
    }
  }
  // range_cluster FCCF.cpp:1020-1038 (exchange sort by size, descending)
  std::vector<int> perm(clusters.size()); for (size_t i = 0; i < perm.size(); i++) perm[i] = (int)i;
  {
    std::vector<int> sz(clusters.size()); for (size_t i = 0; i < sz.size(); i++) sz[i] = (int)clusters[i].size();
    int K = (int)sz.size();
    for (int i = 0; i < K; i++) for (int j = i + 1; j < K; j++)
      if (sz[i] < sz[j]) { std::swap(sz[i], sz[j]); std::swap(perm[i], perm[j]); }
  }
  if (C.keep_blobs) {
    std::vector<int> s, z; for (size_t i = 0; i < perm.size(); i++) { s.push_back(seed_ids[perm[i]]); z.push_back((int)clusters[perm[i]].size()); }
    C.put_i32("cluster_seed_sorted" + tag, s); C.put_i32("cluster_size_sorted" + tag, z);
  }
  int clusternum = (int)clusters[perm[0]].size();
  bool stop = false;
  for (size_t ci = 0; ci < perm.size(); ci++) {
    const std::vector<QT>& cl = clusters[perm[ci]];
    if (stop == false) {
      if ((int)cl.size() >= clusternum) {
        float ax = 0, ay = 0, az = 0;
        for (auto& q : cl) { ax = ax + q.tx; ay = ay + q.ty; az = az + q.tz; }
        ax = ax / cl.size(); ay = ay / cl.size(); az = az / cl.size();
        // average_normal FCCF.cpp:325-367
        float s1x = 0, s1y = 0, s1z = 0, s2x = 0, s2y = 0, s2z = 0;
        for (auto& q : cl) {
          Quatf Q{q.qw, q.qx, q.qy, q.qz};
          V3f a = quat_rotate(Q, V3f{1, 0, 0}), b = quat_rotate(Q, V3f{0, 1, 0});
          s1x = s1x + a.x; s1y = s1y + a.y; s1z = s1z + a.z; s2x = s2x + b.x; s2y = s2y + b.y; s2z = s2z + b.z;
        }
        V3f a1{s1x / cl.size(), s1y / cl.size(), s1z / cl.size()}, a2{s2x / cl.size(), s2y / cl.size(), s2z / cl.size()};
        normalize(a1); normalize(a2);
        M3f R = rotation_from_axes(a1, a2);
        Quatf q = quat_from_matrix(R);
        fine.push_back(QT{q.w, q.x, q.y, q.z, ax, ay, az, true});
        if (fine.size() > (size_t)cluster_num) break;
      } else {
        if ((double)fine.size() < (cluster_num / 2.0)) {
          stop = false; clusternum--;
          if (clusternum < 2) break;
        } else stop = true;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// ceres::Solve restated (Ceres 1.14 trust_region_minimizer.cc, levenberg_marquardt_strategy.cc,
// dense_qr_solver.cc, local_parameterization.cc) for the problem built at FCCF.cpp:210-230 with
// the cost functor at FCCF.cpp:178-208.  App. A.5.  Analytic derivatives of the same expression
// the Jets differentiate.
// ------------------------------------------------------------------------------------------
struct PairFace { int i1, i2; float w; V3f p1, n1, p2, n2; };
static inline void crossd(const double a[3], const double b[3], double o[3]) {
  o[0] = a[1] * b[2] - a[2] * b[1]; o[1] = a[2] * b[0] - a[0] * b[2]; o[2] = a[0] * b[1] - a[1] * b[0];
}
static inline double dot3d(const double a[3], const double b[3]) { return a[0] * b[0] + (a[1] * b[1] + a[2] * b[2]); }
// f(q,a) = a + w*uv + u x uv, uv = 2 (u x a)  with Jacobian wrt (x,y,z,w)
static void rot_with_jac(const double q[4], const double a[3], double f[3], double J[3][4]) {
  const double* u = q; double w = q[3];
  double uv[3]; crossd(u, a, uv); uv[0] += uv[0]; uv[1] += uv[1]; uv[2] += uv[2];
  double c2[3]; crossd(u, uv, c2);
  for (int i = 0; i < 3; i++) f[i] = (a[i] + w * uv[i]) + c2[i];
  if (!J) return;
  for (int k = 0; k < 3; k++) {
    double e[3] = {0, 0, 0}; e[k] = 1.0;
    double A[3]; crossd(e, a, A); A[0] *= 2; A[1] *= 2; A[2] *= 2;
    double t1[3], t2[3]; crossd(e, uv, t1); crossd(u, A, t2);
    for (int i = 0; i < 3; i++) J[i][k] = w * A[i] + t1[i] + t2[i];
  }
  for (int i = 0; i < 3; i++) J[i][3] = uv[i];
}
// Row layout used for every reduction below: 32 rows (row l = pair l/2, residual l%2; rows past
// 2*npairs are zero) plus 6 damping rows of the augmented system held "under" rows 0..5.  Every
// sum over rows is the xor-butterfly order  v[l] += v[l^16], v[l^8], ... v[l^1]  — Ceres/Eigen's
// own blocked summation order is unknowable here, so the oracle fixes this one (it is the order a
// 32-lane warp reduction produces, which lets the CUDA path agree to the last bit of the sums).
// g_lm_sequential = 1 (test switch "lm_sequential"): plain row-order sums v[0] + v[1] + ... instead, to MEASURE how
// much that choice matters (tests/test_oracle_kat.py::test_oracle_switches_do_not_move_the_result).
static int g_lm_sequential = 0;
static double bfly32(const double* v) {
  if (g_lm_sequential) { double s = 0.0; for (int l = 0; l < 32; l++) s += v[l]; return s; }
  double a[16];   // lane 0 of the butterfly only needs the halving tree
  for (int l = 0; l < 16; l++) a[l] = v[l] + v[l + 16];
  for (int o = 8; o; o >>= 1) for (int l = 0; l < o; l++) a[l] = a[l] + a[l + o];
  return a[0];
}
// residual rows and, optionally, the local (tangent-space) Jacobian rows (6 columns)
static bool lm_evaluate(const std::vector<PairFace>& pf, const double x[7], double r[32], double (*J)[6]) {
  int n = (int)pf.size();
  const double* q = x; const double* t = x + 4;
  // EigenQuaternionParameterization::ComputeJacobian (4x3 row-major)
  double Pm[4][3] = {{q[3], q[2], -q[1]}, {-q[2], q[3], q[0]}, {q[1], -q[0], q[3]}, {-q[0], -q[1], -q[2]}};
  for (int l = 0; l < 32; l++) { r[l] = 0.0; if (J) for (int c = 0; c < 6; c++) J[l][c] = 0.0; }
  for (int k = 0; k < n && k < 16; k++) {
    double n1[3] = {pf[k].n1.x, pf[k].n1.y, pf[k].n1.z}, n2[3] = {pf[k].n2.x, pf[k].n2.y, pf[k].n2.z};
    double p1[3] = {pf[k].p1.x, pf[k].p1.y, pf[k].p1.z}, p2[3] = {pf[k].p2.x, pf[k].p2.y, pf[k].p2.z};
    double w = (double)pf[k].w;
    double n2r[3], p2r[3], Jn[3][4], Jp[3][4];
    rot_with_jac(q, n2, n2r, J ? Jn : nullptr);
    rot_with_jac(q, p2, p2r, J ? Jp : nullptr);
    for (int i = 0; i < 3; i++) p2r[i] += t[i];
    double c[3]; crossd(n1, n2r, c);
    double nc = std::sqrt(dot3d(c, c));
    double d = dot3d(n1, p1) - dot3d(n2r, p2r);
    double sd = std::sqrt(d * d);
    r[2 * k] = w * nc; r[2 * k + 1] = w * sd;
    if (!std::isfinite(r[2 * k]) || !std::isfinite(r[2 * k + 1])) return false;
    if (J) {
      double Ja0[7], Ja1[7];
      for (int col = 0; col < 4; col++) {
        double dn[3] = {Jn[0][col], Jn[1][col], Jn[2][col]}, dp[3] = {Jp[0][col], Jp[1][col], Jp[2][col]};
        double dc[3]; crossd(n1, dn, dc);
        Ja0[col] = w * (dot3d(c, dc) / nc);
        double dd = -(dot3d(dn, p2r) + dot3d(n2r, dp));
        Ja1[col] = w * (d * dd / sd);
      }
      for (int col = 0; col < 3; col++) { Ja0[4 + col] = 0.0; Ja1[4 + col] = w * (d * (-n2r[col]) / sd); }
      double* j0 = J[2 * k]; double* j1 = J[2 * k + 1];
      for (int lc = 0; lc < 3; lc++) {
        double s0 = 0, s1 = 0;
        for (int a = 0; a < 4; a++) { s0 += Ja0[a] * Pm[a][lc]; s1 += Ja1[a] * Pm[a][lc]; }
        j0[lc] = s0; j1[lc] = s1;
      }
      for (int lc = 0; lc < 3; lc++) { j0[3 + lc] = Ja0[4 + lc]; j1[3 + lc] = Ja1[4 + lc]; }
      for (int lc = 0; lc < 6; lc++) if (!std::isfinite(j0[lc]) || !std::isfinite(j1[lc])) return false;
    }
  }
  return true;
}
static void lm_plus(const double x[7], const double delta[6], double out[7]) {
  double nd = std::sqrt(delta[0] * delta[0] + delta[1] * delta[1] + delta[2] * delta[2]);
  if (nd > 0.0) {
    double s = std::sin(nd) / nd;
    double dq[4] = {s * delta[0], s * delta[1], s * delta[2], std::cos(nd)};  // x,y,z,w
    const double* b = x;  // delta_q * x  (Eigen quaternion product)
    out[3] = dq[3] * b[3] - dq[0] * b[0] - dq[1] * b[1] - dq[2] * b[2];
    out[0] = dq[3] * b[0] + dq[0] * b[3] + dq[1] * b[2] - dq[2] * b[1];
    out[1] = dq[3] * b[1] + dq[1] * b[3] + dq[2] * b[0] - dq[0] * b[2];
    out[2] = dq[3] * b[2] + dq[2] * b[3] + dq[0] * b[1] - dq[1] * b[0];
  } else { out[0] = x[0]; out[1] = x[1]; out[2] = x[2]; out[3] = x[3]; }
  for (int i = 0; i < 3; i++) out[4 + i] = x[4 + i] + delta[3 + i];
}
// min || [A; B] y - [b; 0] || by Householder QR (DENSE_QR).  A: 32 main rows, B: 6 damping rows.
// Two algebraically identical ways to apply the reflections:
//  * g_lm_textbook = 1 (test switch "lm_textbook"): the loop as Eigen's HouseholderQR states it — the column norm,
//    then v.v, then one v.a_j per remaining column (three dependent reductions per column), back substitution by
//    division, std::pow for the trust-region cubic;
//  * default: per column ONE set of independent reductions g_j = a_k . a_j (j >= k, and the right-hand side), from
//    which norm = sqrt(g_k), v.v = g_k + alpha^2 - 2 alpha a_kk and v.a_j = g_j - alpha a_kj follow; reciprocals of
//    the diagonal first, then multiplications; u*u*u.  This is the form the CUDA path runs (a third of the dependent
//    shuffle depth), so that trajectories agree to the last bit there.  What the choice is worth is MEASURED by
//    tests/test_oracle_kat.py::test_oracle_switches_do_not_move_the_result.
static int g_lm_textbook = 0;
static bool qr_solve6_textbook(double A[32][6], double b[32], double B[6][6], double y[6]) {
  double bb[6] = {0, 0, 0, 0, 0, 0};
  double tmp[32];
  for (int k = 0; k < 6; k++) {
    double mk[32], ak[32];
    for (int l = 0; l < 32; l++) { mk[l] = (l >= k) ? A[l][k] : 0.0; ak[l] = (l < 6) ? B[l][k] : 0.0; }
    for (int l = 0; l < 32; l++) tmp[l] = mk[l] * mk[l] + ak[l] * ak[l];
    double nrm = std::sqrt(bfly32(tmp));
    if (nrm == 0.0) return false;
    double akk = A[k][k];
    double alpha = (akk > 0) ? -nrm : nrm;
    double v0 = akk - alpha;
    double vm[32];
    for (int l = 0; l < 32; l++) vm[l] = (l == k) ? v0 : mk[l];
    for (int l = 0; l < 32; l++) tmp[l] = vm[l] * vm[l] + ak[l] * ak[l];
    double vtv = bfly32(tmp);
    if (vtv == 0.0) return false;
    double beta = 2.0 / vtv;
    for (int j = k + 1; j < 6; j++) {
      for (int l = 0; l < 32; l++) tmp[l] = vm[l] * A[l][j] + ak[l] * ((l < 6) ? B[l][j] : 0.0);
      double s = bfly32(tmp); s *= beta;
      for (int l = 0; l < 32; l++) A[l][j] -= s * vm[l];
      for (int l = 0; l < 6; l++) B[l][j] -= s * ak[l];
    }
    {
      for (int l = 0; l < 32; l++) tmp[l] = vm[l] * b[l] + ak[l] * ((l < 6) ? bb[l] : 0.0);
      double s = bfly32(tmp); s *= beta;
      for (int l = 0; l < 32; l++) b[l] -= s * vm[l];
      for (int l = 0; l < 6; l++) bb[l] -= s * ak[l];
    }
    A[k][k] = alpha;
  }
  for (int k = 5; k >= 0; k--) {
    double s = b[k]; for (int j = k + 1; j < 6; j++) s -= A[k][j] * y[j];
    y[k] = s / A[k][k];
    if (!std::isfinite(y[k])) return false;
  }
  return true;
}
static bool qr_solve6(double A[32][6], double b[32], double B[6][6], double y[6]) {
  if (g_lm_textbook) return qr_solve6_textbook(A, b, B, y);
  double bb[6] = {0, 0, 0, 0, 0, 0};
  double tmp[32];
  for (int k = 0; k < 6; k++) {
    double mk[32], ak[32], g[7];
    for (int l = 0; l < 32; l++) { mk[l] = (l >= k) ? A[l][k] : 0.0; ak[l] = (l < 6) ? B[l][k] : 0.0; }
    for (int j = k; j < 6; j++) {
      for (int l = 0; l < 32; l++) tmp[l] = mk[l] * A[l][j] + ak[l] * ((l < 6) ? B[l][j] : 0.0);
      g[j] = bfly32(tmp);
    }
    for (int l = 0; l < 32; l++) tmp[l] = mk[l] * b[l] + ak[l] * ((l < 6) ? bb[l] : 0.0);
    g[6] = bfly32(tmp);
    double nrm = std::sqrt(g[k]);
    if (nrm == 0.0) return false;
    double akk = A[k][k];
    double alpha = (akk > 0) ? -nrm : nrm;
    double vtv = (g[k] + alpha * alpha) - 2.0 * (alpha * akk);
    if (vtv == 0.0) return false;
    double beta = 2.0 / vtv;
    double vm[32];
    for (int l = 0; l < 32; l++) vm[l] = (l == k) ? (akk - alpha) : mk[l];
    for (int j = k + 1; j < 6; j++) {
      double s = (g[j] - alpha * A[k][j]) * beta;
      for (int l = 0; l < 32; l++) A[l][j] -= s * vm[l];
      for (int l = 0; l < 6; l++) B[l][j] -= s * ak[l];
    }
    {
      double s = (g[6] - alpha * b[k]) * beta;
      for (int l = 0; l < 32; l++) b[l] -= s * vm[l];
      for (int l = 0; l < 6; l++) bb[l] -= s * ak[l];
    }
    A[k][k] = alpha;
  }
  double dinv[6];
  for (int k = 0; k < 6; k++) dinv[k] = 1.0 / A[k][k];
  bool ok = true;
  for (int k = 5; k >= 0; k--) {
    double s = b[k]; for (int j = k + 1; j < 6; j++) s -= A[k][j] * y[j];
    y[k] = s * dinv[k];
    ok = ok && std::isfinite(y[k]);
  }
  return ok;
}
static void ceres_refine(M4f& newT, const std::vector<PairFace>& pf, int* iters_out) {
  const int max_iter = 50;
  const double function_tolerance = 1e-6, gradient_tolerance = 1e-10, parameter_tolerance = 1e-8;
  const double min_relative_decrease = 1e-3, min_radius = 1e-32, max_radius = 1e16;
  const double min_diag = 1e-6, max_diag = 1e32;
  double x[7] = {0, 0, 0, 1, 0, 0, 0};
  double r[32], J[32][6], rc[32], A[32][6], B[6][6], bb[32], tmp[32];
  double radius = 1e4, decrease_factor = 2.0; bool reuse_diag = false;
  double scale[6], diag[6];
  int iter = 0;
  auto colsum = [&](int c, bool with_r) { for (int l = 0; l < 32; l++) tmp[l] = with_r ? J[l][c] * r[l] : J[l][c] * J[l][c]; return bfly32(tmp); };
  auto half_sq = [&](const double* v) { for (int l = 0; l < 32; l++) tmp[l] = v[l] * v[l]; return 0.5 * bfly32(tmp); };
  bool ok = lm_evaluate(pf, x, r, J);
  if (ok) {
    double cost = half_sq(r);
    double g[6];
    auto grad_and_scale = [&](bool first) {
      for (int c = 0; c < 6; c++) g[c] = colsum(c, true);
      if (first) for (int c = 0; c < 6; c++) scale[c] = 1.0 / (1.0 + std::sqrt(colsum(c, false)));
      for (int l = 0; l < 32; l++) for (int c = 0; c < 6; c++) J[l][c] *= scale[c];
    };
    auto grad_max_norm = [&]() {
      double ng[6], xp[7]; for (int c = 0; c < 6; c++) ng[c] = -g[c];
      lm_plus(x, ng, xp);
      double mxn = 0; for (int i = 0; i < 7; i++) mxn = std::fmax(mxn, std::fabs(x[i] - xp[i]));
      return mxn;
    };
    grad_and_scale(true);
    double gmax = grad_max_norm();
    double x_norm = 0; for (int i = 0; i < 7; i++) x_norm += x[i] * x[i]; x_norm = std::sqrt(x_norm);
    int invalid = 0;
    bool step_successful = true;
    while (true) {
      if (iter >= max_iter) break;
      if (step_successful && gmax <= gradient_tolerance) break;
      if (radius < min_radius) break;
      iter++;
      step_successful = false;
      // LevenbergMarquardtStrategy::ComputeStep
      if (!reuse_diag) for (int c = 0; c < 6; c++) diag[c] = std::fmin(std::fmax(colsum(c, false), min_diag), max_diag);
      for (int l = 0; l < 32; l++) { for (int c = 0; c < 6; c++) A[l][c] = J[l][c]; bb[l] = r[l]; }
      for (int i = 0; i < 6; i++) for (int c = 0; c < 6; c++) B[i][c] = (i == c) ? std::sqrt(diag[c] / radius) : 0.0;
      double step[6];
      bool solved = qr_solve6(A, bb, B, step);
      reuse_diag = true;
      bool valid = false; double model_change = 0;
      if (solved) {
        for (int c = 0; c < 6; c++) step[c] = -step[c];
        for (int l = 0; l < 32; l++) { double mr = 0; for (int c = 0; c < 6; c++) mr += J[l][c] * step[c]; tmp[l] = mr * (r[l] + mr / 2.0); }
        model_change = -bfly32(tmp);
        valid = (model_change > 0.0);
      }
      if (!valid) {
        invalid++;
        if (invalid >= 5) break;
        radius = radius / decrease_factor; decrease_factor *= 2.0; reuse_diag = true;
        continue;
      }
      invalid = 0;
      double delta[6]; for (int c = 0; c < 6; c++) delta[c] = step[c] * scale[c];
      double xc[7]; lm_plus(x, delta, xc);
      double cand_cost;
      if (lm_evaluate(pf, xc, rc, nullptr)) cand_cost = half_sq(rc);
      else cand_cost = DBL_MAX;
      double sn = 0; for (int i = 0; i < 7; i++) sn += (x[i] - xc[i]) * (x[i] - xc[i]); sn = std::sqrt(sn);
      if (sn <= parameter_tolerance * (x_norm + parameter_tolerance)) break;
      double cost_change = cost - cand_cost;
      if (std::fabs(cost_change) <= function_tolerance * cost) break;
      double rel = cost_change / model_change;
      if (rel > min_relative_decrease) {
        for (int i = 0; i < 7; i++) x[i] = xc[i];
        x_norm = 0; for (int i = 0; i < 7; i++) x_norm += x[i] * x[i]; x_norm = std::sqrt(x_norm);
        cost = cand_cost;
        if (!lm_evaluate(pf, x, r, J)) break;
        grad_and_scale(false);
        gmax = grad_max_norm();
        step_successful = true;
        { const double u = 2.0 * rel - 1.0; radius = radius / std::fmax(1.0 / 3.0, 1.0 - (g_lm_textbook ? std::pow(u, 3) : u * u * u)); }
        radius = std::fmin(max_radius, radius);
        decrease_factor = 2.0; reuse_diag = false;
      } else {
        radius = radius / decrease_factor; decrease_factor *= 2.0; reuse_diag = true;
      }
    }
  }
  if (iters_out) *iters_out = iter;
  Quatf qf; qf.w = (float)x[3]; qf.x = (float)x[0]; qf.y = (float)x[1]; qf.z = (float)x[2];  // FCCF.cpp:231-236
  M3f R = quat_to_matrix(qf);
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) newT.m[i][j] = R.m[i][j];
  newT.m[0][3] = (float)x[4]; newT.m[1][3] = (float)x[5]; newT.m[2][3] = (float)x[6];
}

// FCCF.cpp:680-783
static float quick_verify(Ctx& C, M4f& T, const std::vector<FaceNode>& f1, const std::vector<FaceNode>& f2, std::vector<int>* pairs_out, int* iters_out) {
  const Params& P = C.p;
  int fs1 = 0, fs2 = 0;
  std::vector<V3f> p1, n1, p2, n2;
  for (auto& f : f1) { p1.push_back(V3f{f.cx, f.cy, f.cz}); n1.push_back(V3f{f.nx, f.ny, f.nz}); fs1 = (int)(fs1 + f.size); }
  for (auto& f : f2) { p2.push_back(tf_se3(T, V3f{f.cx, f.cy, f.cz})); n2.push_back(tf_so3(T, V3f{f.nx, f.ny, f.nz})); fs2 = (int)(fs2 + f.size); }
  std::vector<PairFace> pf;
  for (int a = 0; a < (int)f1.size(); a++) {
    std::vector<int> cand; bool find = false;
    for (int b = 0; b < (int)f2.size(); b++) {
      float angel = compute_normal_angel(n1[a].x, n1[a].y, n1[a].z, n2[b].x, n2[b].y, n2[b].z);
      V3d dn1{n1[a].x, n1[a].y, n1[a].z}, dn2{n2[b].x, n2[b].y, n2[b].z}, c1{p1[a].x, p1[a].y, p1[a].z}, c2{p2[b].x, p2[b].y, p2[b].z};
      float d1 = (float)dotd(dn1, c1), d2 = (float)dotd(dn2, c2);
      float dist = (float)std::fabs(d1 - d2);
      if (angel < P.quick_verify_angel_threshold && dist < P.quick_verify_distance_threshold) { find = true; cand.push_back(b); }
    }
    float size1 = f1[a].size; int best = 0; float best_imp = 0, best_score = 0;
    for (int b : cand) {
      float size2 = f2[b].size;
      float mn = size1 < size2 ? size1 : size2, mx = size1 > size2 ? size1 : size2;
      float cs = mn / mx;
      float ci = (2 * mn) / (fs1 + fs2);
      if (cs > best_score) { best_imp = ci; best_score = cs; best = b; }
    }
    if (find) pf.push_back(PairFace{a, best, best_imp, p1[a], n1[a], p2[best], n2[best]});
  }
  if (pairs_out) { pairs_out->clear(); for (auto& p : pf) { pairs_out->push_back(p.i1); pairs_out->push_back(p.i2); } }
  int iters = -1;
  if ((float)pf.size() >= P.required_optimize_plane) {
    M4f nt = identity4();
    ceres_refine(nt, pf, &iters);
    T = mul(nt, T);
  }
  if (iters_out) *iters_out = iters;
  float score = 0;
  for (auto& p : pf) score = score + p.w;
  return score;
}

// FCCF.cpp:785-839.  counts_out (optional): rows (Lx,Ly,Lz,s,t) for every voxel holding both kinds,
// L = final key minus the final key of the first static point (hypothesis-independent lattice).
static float fine_verify(const Params& P, const M4f& T, const std::vector<P3>& src, const std::vector<P3>& tgt, std::vector<int>* counts_out) {
  std::vector<P3> fuse; fuse.reserve(src.size() + tgt.size());
  for (auto& p : src) fuse.push_back(p);
  for (auto& p : tgt) { V3f q = tf_se3(T, V3f{p.x, p.y, p.z}); fuse.push_back(P3{q.x, q.y, q.z}); }
  Octree oct; octree_build(oct, fuse, (double)P.fine_verify_voxel_size);
  int V = (int)oct.start.size() - 1; if (V < 0) V = 0;
  float thr = 1, similar = 0, allin = 0;
  int ns = (int)src.size();
  for (int v = 0; v < V; v++) {
    float s = 0, t = 0;
    for (int k = oct.start[v]; k < oct.start[v + 1]; k++) { if (oct.pidx[k] < ns) s++; else t++; }
    allin = allin + s + t;
    if (s >= thr && t >= thr) {
      float mn = s < t ? s : t, mx = s > t ? s : t;
      similar = similar + (s + t) * (mn / mx);
      if (counts_out) {
        counts_out->push_back((int)oct.key[3 * v] - (int)oct.key0[0]); counts_out->push_back((int)oct.key[3 * v + 1] - (int)oct.key0[1]);
        counts_out->push_back((int)oct.key[3 * v + 2] - (int)oct.key0[2]); counts_out->push_back((int)s); counts_out->push_back((int)t);
      }
    }
  }
  return similar / allin;
}

struct TScore { M4f T; float score, score2; int centre; };
// FCCF.cpp:1233-1251
static void score_range(std::vector<TScore>& v) {
  for (size_t i = 0; i + 1 < v.size(); i++) for (size_t j = i + 1; j < v.size(); j++) if (v[i].score < v[j].score) std::swap(v[i], v[j]);
}
struct HighScore { QT qt; float score; };

// FCCF.cpp:1291-1368 (+ weight_normal 1253-1289)
static void fuse_answer(M4f& best, const std::vector<HighScore>& hs, float sum_score) {
  float ax = 0, ay = 0, az = 0;
  for (auto& h : hs) { ax = ax + h.qt.tx * (h.score / sum_score); ay = ay + h.qt.ty * (h.score / sum_score); az = az + h.qt.tz * (h.score / sum_score); }
  float s1x = 0, s1y = 0, s1z = 0, s2x = 0, s2y = 0, s2z = 0;
  for (auto& h : hs) {
    Quatf Q{h.qt.qw, h.qt.qx, h.qt.qy, h.qt.qz};
    V3f a = quat_rotate(Q, V3f{1, 0, 0}), b = quat_rotate(Q, V3f{0, 1, 0});
    float w = h.score / sum_score;
    s1x = s1x + a.x * w; s1y = s1y + a.y * w; s1z = s1z + a.z * w; s2x = s2x + b.x * w; s2y = s2y + b.y * w; s2z = s2z + b.z * w;
  }
  V3f a1{s1x, s1y, s1z}, a2{s2x, s2y, s2z};
  normalize(a1); normalize(a2);
  M3f R = rotation_from_axes(a1, a2);
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) best.m[i][j] = R.m[i][j];
  best.m[0][3] = ax; best.m[1][3] = ay; best.m[2][3] = az;
}

static void put_points(Ctx& C, const std::string& name, const std::vector<P3>& v) {
  if (!C.keep_blobs) return;
  C.put(name, DT_F32, (const float*)v.data(), v.size() * 3);
}

// FCCF.cpp:1370-1608.  `source`/`target` are the function's parameters: cloud 1 / cloud 2.
// select_base x 2 (FCCF.cpp:1406,1409), the pair-descriptor match loop (1412-1427) and computer_transform (841-1018):
// the three hypothesis pools in push_back order.  Blobs: base1/2, base_angle1/2, matches.
static void stage_hypotheses(Ctx& C, const std::vector<FaceNode>& f1, const std::vector<double>& th1, const std::vector<FaceNode>& f2,
                             const std::vector<double>& th2, std::vector<std::vector<M4f>>& tv) {
  const Params& P = C.p;
  std::vector<FaceBase> b1, b2; std::vector<int> ty1, ty2;
  select_base(P, b1, f1, ty1, th1);
  select_base(P, b2, f2, ty2, th2);
  if (C.keep_blobs) {
    for (int c = 0; c < 2; c++) {
      auto& b = c ? b2 : b1; auto& ty = c ? ty2 : ty1;
      std::vector<int> bi; std::vector<float> ba;
      for (size_t i = 0; i < b.size(); i++) { bi.push_back(b[i].i1); bi.push_back(b[i].i2); bi.push_back(ty[i]); ba.push_back(b[i].angel); }
      C.put_i32(c ? "base2" : "base1", bi); C.put_f32(c ? "base_angle2" : "base_angle1", ba);
    }
  }
  float angthr = P.included_angle_same_threshold;
  std::vector<int> matches;
  for (size_t i1 = 0; i1 < b1.size(); i1++)
    for (size_t i2 = 0; i2 < b2.size(); i2++)
      if ((std::fabs(b1[i1].angel - b2[i2].angel)) < angthr && ty1[i1] == ty2[i2] && ty1[i1] < 3) {
        size_t before = tv[ty1[i1]].size();
        computer_transform(P, tv, b1[i1].i1, b1[i1].i2, b2[i2].i1, b2[i2].i2, f1, f2, ty1[i1]);
        matches.push_back((int)i1); matches.push_back((int)i2); matches.push_back((int)(tv[ty1[i1]].size() - before));
      }
  C.put_i32("matches", matches);
}

// matrix -> quaternion of one pool (FCCF.cpp:1439-1462).  Blobs: hyp<t>, hyp_qt<t>.
static void pool_to_qt(Ctx& C, const char* tg, const std::vector<M4f>& pool, std::vector<QT>& qv) {
  for (auto& M : pool) {
    M3f R; for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) R.m[a][b] = M.m[a][b];
    Quatf q = quat_from_matrix(R);
    qv.push_back(QT{q.w, q.x, q.y, q.z, M.m[0][3], M.m[1][3], M.m[2][3], false});
  }
  if (C.keep_blobs) {
    std::vector<float> h; for (auto& M : pool) for (int a = 0; a < 3; a++) for (int b = 0; b < 4; b++) h.push_back(M.m[a][b]);
    C.put_f32(std::string("hyp") + tg, h);
    std::vector<float> hq; for (auto& q : qv) { hq.push_back(q.qw); hq.push_back(q.qx); hq.push_back(q.qy); hq.push_back(q.qz); hq.push_back(q.tx); hq.push_back(q.ty); hq.push_back(q.tz); }
    C.put_f32(std::string("hyp_qt") + tg, hq);
  }
}

// cluster_num (FCCF.cpp:1465) + transform_cluster (1466) of one pool.  Blob: centre<t> (+ the seed / size lists).
static int pool_cluster(Ctx& C, const char* tg, std::vector<QT>& qv, int tnum, std::vector<QT>& fine) {
  const Params& P = C.p;
  float cnf = P.seclct_cluster_number * qv.size() / tnum;
  int cluster_num = (cnf == cnf) ? (int)cnf : INT32_MIN;  // NaN cast: cvttss2si gives INT_MIN
  transform_cluster(C, tg, qv, fine, cluster_num);
  if (C.keep_blobs) {
    std::vector<float> ce; for (auto& q : fine) { ce.push_back(q.qw); ce.push_back(q.qx); ce.push_back(q.qy); ce.push_back(q.qz); ce.push_back(q.tx); ce.push_back(q.ty); ce.push_back(q.tz); }
    C.put_f32(std::string("centre") + tg, ce);
  }
  return cluster_num;
}

// Per-type best of the fine-verified hypotheses by s1 / sum(s1) + s2 / sum(s2) with the sums over ALL types (Q19),
// the 0.8 gate and fuse_answer (FCCF.cpp:1546-1606).  cand[t]: the fine-verified hypotheses of type t in rank order.
// Blob: type_best.
static void stage_fuse(Ctx& C, const std::vector<std::vector<TScore>>& cand, M4f& best) {
  float score_sum = 0, score1_sum = 0, score2_sum = 0;
  for (int i = 0; i < 3; i++) for (auto& ts : cand[i]) { score2_sum += ts.score2; score1_sum += ts.score; }
  std::vector<HighScore> tmp3; float best_best = 0;
  std::vector<float> type_best;
  for (int i = 0; i < 3; i++) {
    float best_score = 0; M4f tb = identity4();
    for (auto& ts : cand[i]) {
      float score = ts.score / score1_sum + ts.score2 / score2_sum;
      if (score > best_score) { best_score = score; tb = ts.T; }
    }
    if (best_best < best_score) best_best = best_score;
    M3f R; for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) R.m[a][b] = tb.m[a][b];
    Quatf q = quat_from_matrix(R);
    HighScore h; h.qt = QT{q.w, q.x, q.y, q.z, tb.m[0][3], tb.m[1][3], tb.m[2][3], false}; h.score = best_score;
    tmp3.push_back(h);
    type_best.push_back(best_score); for (int a = 0; a < 3; a++) for (int b = 0; b < 4; b++) type_best.push_back(tb.m[a][b]);
  }
  C.put_f32("type_best", type_best);
  std::vector<HighScore> hs;
  for (auto& h : tmp3) if (h.score > best_best * 0.8) { hs.push_back(h); score_sum += h.score; }
  fuse_answer(best, hs, score_sum);
}

static void computer_transform_guess(Ctx& C, const std::vector<P3>& source, const std::vector<P3>& target, float leaf, M4f& best) {
  const Params& P = C.p;
  std::vector<P3> c1, c2;
  std::vector<int64_t> cell; std::vector<int> cnt;
  voxel_grid(P, source, leaf, c1, &cell, &cnt);   // FCCF.cpp:1377-1381 (NaN removal folded in)
  if (C.keep_blobs) { put_points(C, "vg2_xyz1", c1); C.put_i64("vg2_cell1", cell); C.put_i32("vg2_cnt1", cnt); }
  voxel_grid(P, target, leaf, c2, &cell, &cnt);   // FCCF.cpp:1383-1387
  if (C.keep_blobs) { put_points(C, "vg2_xyz2", c2); C.put_i64("vg2_cell2", cell); C.put_i32("vg2_cnt2", cnt); }
  std::vector<FaceNode> f1, f2; std::vector<double> th1, th2;
  C.sub1.clear(); C.sub2.clear();
  face_extrate(C, "1", c1, f1, C.sub1, th1);
  face_extrate(C, "2", c2, f2, C.sub2, th2);
  std::vector<std::vector<M4f>> tv(3);
  stage_hypotheses(C, f1, th1, f2, th2, tv);
  int tnum = (int)(tv[0].size() + tv[1].size() + tv[2].size());
  std::vector<std::vector<TScore>> ctv(3);
  int analyse_max = (int)P.fine_verify_number;
  std::vector<int> n_hyp, n_centres, cluster_nums;
  for (int i = 0; i < 3; i++) {
    char tg[8]; snprintf(tg, sizeof tg, "%d", i);
    std::vector<QT> qv;
    pool_to_qt(C, tg, tv[i], qv);
    std::vector<QT> fine;
    int cluster_num = pool_cluster(C, tg, qv, tnum, fine);
    n_hyp.push_back((int)tv[i].size()); n_centres.push_back((int)fine.size()); cluster_nums.push_back(cluster_num);
    std::vector<float> qs, qT; std::vector<int> qpairs, qpoff, qiters;
    int ci = 0;
    for (auto& q : fine) {
      M3f R = quat_to_matrix(Quatf{q.qw, q.qx, q.qy, q.qz});
      TScore ts; ts.T = identity4(); ts.centre = ci++;
      for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) ts.T.m[a][b] = R.m[a][b];
      ts.T.m[0][3] = q.tx; ts.T.m[1][3] = q.ty; ts.T.m[2][3] = q.tz;
      std::vector<int> pr; int it;
      ts.score = quick_verify(C, ts.T, f1, f2, &pr, &it); ts.score2 = 0;
      ctv[i].push_back(ts);
      if (C.keep_blobs) { qs.push_back(ts.score); for (int a = 0; a < 4; a++) for (int b = 0; b < 4; b++) qT.push_back(ts.T.m[a][b]); qpoff.push_back((int)qpairs.size() / 2); for (int v : pr) qpairs.push_back(v); qiters.push_back(it); }
    }
    if (C.keep_blobs) { qpoff.push_back((int)qpairs.size() / 2); C.put_f32(std::string("qv_score") + tg, qs); C.put_f32(std::string("qv_T") + tg, qT); C.put_i32(std::string("qv_pairs") + tg, qpairs); C.put_i32(std::string("qv_pair_off") + tg, qpoff); C.put_i32(std::string("qv_iters") + tg, qiters); }
    score_range(ctv[i]);
    int asum = 0;
    std::vector<float> topT, tops1, tops2; std::vector<int> topc, fvc, fvoff;
    for (auto& ts : ctv[i]) {
      if (asum < analyse_max) {
        asum++;
        std::vector<int> counts;
        ts.score2 = fine_verify(P, ts.T, C.sub1, C.sub2, C.keep_blobs ? &counts : nullptr);
        if (C.keep_blobs) { for (int a = 0; a < 4; a++) for (int b = 0; b < 4; b++) topT.push_back(ts.T.m[a][b]); tops1.push_back(ts.score); tops2.push_back(ts.score2); topc.push_back(ts.centre); fvoff.push_back((int)fvc.size() / 5); for (int v : counts) fvc.push_back(v); }
      } else break;
    }
    if (C.keep_blobs) { fvoff.push_back((int)fvc.size() / 5); C.put_f32(std::string("top_T") + tg, topT); C.put_f32(std::string("top_s1") + tg, tops1); C.put_f32(std::string("top_s2") + tg, tops2); C.put_i32(std::string("top_centre") + tg, topc); C.put_i32(std::string("fv_counts") + tg, fvc); C.put_i32(std::string("fv_off") + tg, fvoff); }
  }
  C.put_i32("n_hyp", n_hyp); C.put_i32("n_centres", n_centres); C.put_i32("cluster_num", cluster_nums);
  std::vector<std::vector<TScore>> cand(3);
  for (int i = 0; i < 3; i++) { int asum = 0; for (auto& ts : ctv[i]) { if (asum < analyse_max) { asum++; cand[i].push_back(ts); } else break; } }
  stage_fuse(C, cand, best);
}

// main(): FCCF.cpp:1646-1690 (argv order: SRC, TAR; pipeline called with (TAR, SRC), Q1)
static void run_main(Ctx& C, const std::vector<P3>& src, const std::vector<P3>& tar, float leaf, M4f& T) {
  auto t0 = std::chrono::steady_clock::now();
  std::vector<P3> cs, ct; std::vector<int64_t> cell; std::vector<int> cnt;
  voxel_grid(C.p, src, leaf, cs, &cell, &cnt);
  if (C.keep_blobs) { put_points(C, "vg1_xyz2", cs); C.put_i64("vg1_cell2", cell); C.put_i32("vg1_cnt2", cnt); }
  voxel_grid(C.p, tar, leaf, ct, &cell, &cnt);
  if (C.keep_blobs) { put_points(C, "vg1_xyz1", ct); C.put_i64("vg1_cell1", cell); C.put_i32("vg1_cnt1", cnt); }
  T = identity4();
  auto t1 = std::chrono::steady_clock::now();
  computer_transform_guess(C, ct, cs, leaf, T);
  auto t2 = std::chrono::steady_clock::now();
  C.t_pipeline_s = std::chrono::duration<double>(t2 - t1).count();
  C.t_total_s = std::chrono::duration<double>(t2 - t0).count();
}

}  // namespace orc

// ==========================================================================================
// C interface for tests (ctypes)
// ==========================================================================================
using namespace orc;
static std::vector<P3> to_points(const float* xyz, int64_t n) {
  std::vector<P3> v((size_t)n);
  if (n) memcpy(v.data(), xyz, (size_t)n * 12);
  return v;
}
extern "C" {
void* orc_create() { return new Ctx(); }
void orc_destroy(void* c) { delete (Ctx*)c; }
void orc_keep_blobs(void* c, int k) { ((Ctx*)c)->keep_blobs = k; }
int orc_set_param(void* c, const char* name, double v) {
  Params& p = ((Ctx*)c)->p; std::string n(name);
#define SP(f) if (n == #f) { p.f = (float)v; return 0; }
  SP(parameter_l1) SP(parameter_l2) SP(parameter_k1) SP(parameter_k2) SP(normal_vector_threshold1) SP(normal_vector_threshold2)
  SP(face_voxel_size) SP(voxel_point_threshold) SP(curvature_threshold) SP(select_plane_number) SP(quick_verify_angel_threshold)
  SP(quick_verify_distance_threshold) SP(required_optimize_plane) SP(fine_verify_voxel_size) SP(fine_verify_number)
  SP(included_angle_same_threshold) SP(included_angle_min_threshold) SP(included_angle_max_threshold) SP(third_plane_threshold)
  SP(third_plane_normal_threshold) SP(cluster_number_threshold) SP(cluster_angel_threshold) SP(cluster_distance_threshold)
  SP(seclct_cluster_number) SP(rough_threshold_gl)
#undef SP
  if (n == "emulate_pcl_overflow") { p.emulate_pcl_overflow = (int)v; return 0; }
  if (n == "libm_float") { g_libm_float = (int)v; return 0; }   // process-wide (test switch)
  if (n == "lm_sequential") { g_lm_sequential = (int)v; return 0; }   // process-wide (test switch)
  if (n == "lm_textbook") { g_lm_textbook = (int)v; return 0; }       // process-wide (test switch)
  return -1;
}
// full program: src = argv[1], tar = argv[2]; T row-major 16 floats
int orc_register(void* c, const float* src, int64_t ns, const float* tar, int64_t nt, float leaf, float* T16) {
  Ctx& C = *(Ctx*)c; C.blobs.clear();
  M4f T; run_main(C, to_points(src, ns), to_points(tar, nt), leaf, T);
  for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) T16[i * 4 + j] = T.m[i][j];
  return 0;
}
double orc_time_pipeline(void* c) { return ((Ctx*)c)->t_pipeline_s; }
double orc_time_total(void* c) { return ((Ctx*)c)->t_total_s; }
// stage entry points
int64_t orc_voxelgrid(void* c, const float* xyz, int64_t n, float leaf, float* out_xyz, int64_t* out_cell, int* out_cnt) {
  Ctx& C = *(Ctx*)c; std::vector<P3> out; std::vector<int64_t> cell; std::vector<int> cnt;
  voxel_grid(C.p, to_points(xyz, n), leaf, out, &cell, &cnt);
  if (out_xyz && !out.empty()) memcpy(out_xyz, out.data(), out.size() * 12);
  if (out_cell && !cell.empty()) memcpy(out_cell, cell.data(), cell.size() * 8);
  if (out_cnt && !cnt.empty()) memcpy(out_cnt, cnt.data(), cnt.size() * 4);
  return (int64_t)out.size();
}
// octree voxelisation: returns V; keys 3 per voxel; start V+1; pidx n; mn 3 doubles; depth
int orc_octree(const float* xyz, int64_t n, double res, int* keys, int* start, int* pidx, double* mn, int* depth) {
  Octree o; octree_build(o, to_points(xyz, n), res);
  int V = (int)o.start.size() - 1; if (V < 0) V = 0;
  for (size_t i = 0; i < o.key.size(); i++) keys[i] = (int)o.key[i];
  for (size_t i = 0; i < o.start.size(); i++) start[i] = o.start[i];
  for (size_t i = 0; i < o.pidx.size(); i++) pidx[i] = o.pidx[i];
  for (int a = 0; a < 3; a++) mn[a] = o.mn[a];
  *depth = o.depth;
  return V;
}
// face extraction on an (already downsampled) cloud; blobs with tag "1"
int orc_face_extract(void* c, const float* xyz, int64_t n) {
  Ctx& C = *(Ctx*)c; C.blobs.clear();
  std::vector<FaceNode> f; std::vector<P3> sub; std::vector<double> th;
  face_extrate(C, "1", to_points(xyz, n), f, sub, th);
  return (int)f.size();
}
void orc_plane_fit(const float* xyz, int n, float* out8) {
  std::vector<P3> v = to_points(xyz, n); std::vector<int> idx(n); for (int i = 0; i < n; i++) idx[i] = i;
  float cen[3], nrm[3], curv; plane_fit(v, idx.data(), n, cen, nrm, curv);
  out8[0] = cen[0]; out8[1] = cen[1]; out8[2] = cen[2]; out8[3] = nrm[0]; out8[4] = nrm[1]; out8[5] = nrm[2]; out8[6] = curv; out8[7] = (float)n;
}
float orc_normal_angle(float x1, float y1, float z1, float x2, float y2, float z2) { return compute_normal_angel(x1, y1, z1, x2, y2, z2); }
void orc_quat_from_matrix(const float* R9, float* q4) { M3f R; memcpy(R.m, R9, 36); Quatf q = quat_from_matrix(R); q4[0] = q.w; q4[1] = q.x; q4[2] = q.y; q4[3] = q.z; }
void orc_quat_to_matrix(const float* q4, float* R9) { M3f R = quat_to_matrix(Quatf{q4[0], q4[1], q4[2], q4[3]}); memcpy(R9, R.m, 36); }
// fine_verify of one hypothesis (T row-major 16) for static cloud s1 / moving cloud s2.
// counts: capacity rows of 5 ints; returns score; *nrows = rows written
float orc_fine_verify(void* c, const float* T16, const float* s1, int64_t n1, const float* s2, int64_t n2, int* counts, int cap_rows, int* nrows) {
  Ctx& C = *(Ctx*)c; M4f T; memcpy(T.m, T16, 64);
  std::vector<int> cv;
  float s = fine_verify(C.p, T, to_points(s1, n1), to_points(s2, n2), counts ? &cv : nullptr);
  if (counts) { int rows = (int)cv.size() / 5; if (rows > cap_rows) rows = cap_rows; memcpy(counts, cv.data(), (size_t)rows * 20); if (nrows) *nrows = (int)cv.size() / 5; }
  return s;
}
// time nrep fine_verify evaluations (hypotheses: nhyp row-major 4x4, cycled); returns seconds
double orc_bench_fine_verify(void* c, const float* T16, int nhyp, const float* s1, int64_t n1, const float* s2, int64_t n2, int nrep, float* checksum) {
  Ctx& C = *(Ctx*)c; std::vector<P3> a = to_points(s1, n1), b = to_points(s2, n2);
  float acc = 0;
  auto t0 = std::chrono::steady_clock::now();
  for (int r = 0; r < nrep; r++) { M4f T; memcpy(T.m, T16 + 16 * (r % nhyp), 64); acc += fine_verify(C.p, T, a, b, nullptr); }
  auto t1 = std::chrono::steady_clock::now();
  if (checksum) *checksum = acc;
  return std::chrono::duration<double>(t1 - t0).count();
}
// quick_verify of one hypothesis given plane tables (F x 7: c, n, size); T updated in place
float orc_quick_verify(void* c, float* T16, const float* planes1, int F1, const float* planes2, int F2, int* pairs, int* npairs, int* iters) {
  Ctx& C = *(Ctx*)c; std::vector<FaceNode> f1(F1), f2(F2);
  for (int i = 0; i < F1; i++) { const float* p = planes1 + 7 * i; f1[i].cx = p[0]; f1[i].cy = p[1]; f1[i].cz = p[2]; f1[i].nx = p[3]; f1[i].ny = p[4]; f1[i].nz = p[5]; f1[i].size = p[6]; }
  for (int i = 0; i < F2; i++) { const float* p = planes2 + 7 * i; f2[i].cx = p[0]; f2[i].cy = p[1]; f2[i].cz = p[2]; f2[i].nx = p[3]; f2[i].ny = p[4]; f2[i].nz = p[5]; f2[i].size = p[6]; }
  M4f T; memcpy(T.m, T16, 64); std::vector<int> pr; int it;
  float s = quick_verify(C, T, f1, f2, &pr, &it);
  memcpy(T16, T.m, 64);
  if (pairs) memcpy(pairs, pr.data(), pr.size() * 4);
  if (npairs) *npairs = (int)pr.size() / 2;
  if (iters) *iters = it;
  return s;
}
static std::vector<FaceNode> to_faces(const float* planes, int F) {
  std::vector<FaceNode> f(F);
  for (int i = 0; i < F; i++) { const float* p = planes + 7 * i; f[i].cx = p[0]; f[i].cy = p[1]; f[i].cz = p[2]; f[i].nx = p[3]; f[i].ny = p[4]; f[i].nz = p[5]; f[i].size = p[6]; f[i].alloc = false; f[i].id = i; }
  return f;
}
// select_base + match loop + computer_transform + matrix -> quaternion on two plane tables (F x 7: c, n, size) with their
// roughness values; results as blobs base1/2, base_angle1/2, matches, n_hyp, hyp<t>, hyp_qt<t>
int orc_hypotheses(void* c, const float* planes1, const double* theta1, int F1, const float* planes2, const double* theta2, int F2) {
  Ctx& C = *(Ctx*)c; C.blobs.clear();
  std::vector<FaceNode> f1 = to_faces(planes1, F1), f2 = to_faces(planes2, F2);
  std::vector<double> th1(theta1, theta1 + F1), th2(theta2, theta2 + F2);
  std::vector<std::vector<M4f>> tv(3);
  stage_hypotheses(C, f1, th1, f2, th2, tv);
  std::vector<int> n_hyp;
  for (int i = 0; i < 3; i++) { char tg[8]; snprintf(tg, sizeof tg, "%d", i); std::vector<QT> qv; pool_to_qt(C, tg, tv[i], qv); n_hyp.push_back((int)tv[i].size()); }
  C.put_i32("n_hyp", n_hyp);
  return 0;
}
// cluster_num + transform_cluster of three pools given as (qw qx qy qz tx ty tz) rows, pools concatenated; results as
// blobs centre<t>, n_centres, cluster_num, cluster_seed_sorted<t>, cluster_size_sorted<t>
int orc_cluster(void* c, const float* qt7, const int* n_hyp3) {
  Ctx& C = *(Ctx*)c; C.blobs.clear();
  int tnum = n_hyp3[0] + n_hyp3[1] + n_hyp3[2];
  std::vector<int> n_centres, cluster_nums; size_t off = 0;
  for (int i = 0; i < 3; i++) {
    char tg[8]; snprintf(tg, sizeof tg, "%d", i);
    std::vector<QT> qv, fine;
    for (int k = 0; k < n_hyp3[i]; k++) { const float* q = qt7 + 7 * (off + k); qv.push_back(QT{q[0], q[1], q[2], q[3], q[4], q[5], q[6], false}); }
    off += n_hyp3[i];
    cluster_nums.push_back(pool_cluster(C, tg, qv, tnum, fine)); n_centres.push_back((int)fine.size());
  }
  C.put_i32("n_centres", n_centres); C.put_i32("cluster_num", cluster_nums);
  return 0;
}
// per-type best + 0.8 gate + fuse_answer on up to 3 x k fine-verified candidates (row-major 4x4, quick score s1, fine
// score s2; n_top3[t] candidates of type t in rank order); blob type_best
int orc_fuse(void* c, const float* top_T, const float* s1, const float* s2, const int* n_top3, int k, float* T16) {
  Ctx& C = *(Ctx*)c; C.blobs.clear();
  std::vector<std::vector<TScore>> cand(3);
  for (int t = 0; t < 3; t++) for (int j = 0; j < n_top3[t]; j++) {
    TScore ts; memcpy(ts.T.m, top_T + 16 * ((size_t)t * k + j), 64); ts.score = s1[t * k + j]; ts.score2 = s2[t * k + j]; ts.centre = j;
    cand[t].push_back(ts);
  }
  M4f best = identity4();
  stage_fuse(C, cand, best);
  memcpy(T16, best.m, 64);
  return 0;
}
// stdout of main() (FCCF.cpp:1667, 1687): default ostream float formatting + Eigen's default IOFormat
// (precision = stream's, every column right-aligned to the widest coefficient, " " and "\n" separators)
int orc_format_output(float leaf, const float* T16, char* buf, int cap) {
  // operator<<(float) with the default precision 6 and no flags is printf's %g
  std::string s; char c[64];
  snprintf(c, sizeof c, "Leaf size : %g\n", (double)leaf); s += c;
  s += "Transformation: \n";
  size_t width = 0;
  for (int k = 0; k < 16; k++) { snprintf(c, sizeof c, "%g", (double)T16[k]); width = std::max(width, strlen(c)); }
  for (int i = 0; i < 4; i++) {
    if (i) s += "\n";
    for (int j = 0; j < 4; j++) { if (j) s += " "; snprintf(c, sizeof c, "%*g", (int)width, (double)T16[4 * i + j]); s += c; }
  }
  s += "\n";
  if ((int)s.size() + 1 > cap) return -1;
  memcpy(buf, s.c_str(), s.size() + 1);
  return (int)s.size();
}
// blobs
int64_t orc_blob_bytes(void* c, const char* name) { Ctx& C = *(Ctx*)c; auto it = C.blobs.find(name); return it == C.blobs.end() ? -1 : (int64_t)it->second.data.size(); }
int orc_blob_dtype(void* c, const char* name) { Ctx& C = *(Ctx*)c; auto it = C.blobs.find(name); return it == C.blobs.end() ? -1 : it->second.dtype; }
int orc_blob_copy(void* c, const char* name, void* dst) { Ctx& C = *(Ctx*)c; auto it = C.blobs.find(name); if (it == C.blobs.end()) return -1; if (!it->second.data.empty()) memcpy(dst, it->second.data.data(), it->second.data.size()); return 0; }
int orc_blob_names(void* c, char* buf, int cap) {
  Ctx& C = *(Ctx*)c; std::string s; for (auto& kv : C.blobs) { s += kv.first; s += "\n"; }
  if ((int)s.size() + 1 > cap) return -(int)s.size() - 1;
  memcpy(buf, s.c_str(), s.size() + 1); return (int)s.size();
}
}
