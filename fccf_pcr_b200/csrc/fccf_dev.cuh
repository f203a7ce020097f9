// fccf_dev.cuh — device-side arithmetic shared by the kernels.
//
// Everything here is compiled with -fmad=false: the reference is built for baseline x86-64
// (CMakeLists.txt:5-10: -O3, no -march=native), i.e. IEEE single/double without FMA contraction,
// and the parity contract (BASELINE.json) is bit-exact decisions against that arithmetic.
// Evaluation orders follow Eigen 3.3's fixed-size kernels (3-term reductions are a0 + (a1 + a2))
// and PCL 1.10's SSE transform helper (x*c0 + (y*c1 + (z*c2 + c3))).
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

namespace fccf {

struct f3 { float x, y, z; };
struct d3 { double x, y, z; };
struct m3 { float m[3][3]; };
struct q4 { float w, x, y, z; };

__device__ __forceinline__ f3 mk3(float x, float y, float z) { f3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ float get(const f3& v, int i) { return i == 0 ? v.x : (i == 1 ? v.y : v.z); }
__device__ __forceinline__ float sum3(float a, float b, float c) { return a + (b + c); }
__device__ __forceinline__ double sum3d(double a, double b, double c) { return a + (b + c); }
__device__ __forceinline__ float dot(const f3& a, const f3& b) { return sum3(a.x * b.x, a.y * b.y, a.z * b.z); }
__device__ __forceinline__ f3 cross(const f3& a, const f3& b) {
  return mk3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
__device__ __forceinline__ void normalize(f3& v) {  // Eigen normalize(): only if squaredNorm > 0
  float z = dot(v, v);
  if (z > 0.0f) { float n = sqrtf(z); v.x = v.x / n; v.y = v.y / n; v.z = v.z / n; }
}

// FCCF.cpp:369-377 compute_normal_angel: double dot/norms, quotient rounded to float, acos in
// double, degrees, rounded to float.
__device__ __forceinline__ float normal_angle(float x1, float y1, float z1, float x2, float y2, float z2) {
  double a0 = x1, a1 = y1, a2 = z1, b0 = x2, b1 = y2, b2 = z2;
  float n1n3 = (float)sum3d(a0 * b0, a1 * b1, a2 * b2);
  double na = sqrt(sum3d(a0 * a0, a1 * a1, a2 * a2)), nb = sqrt(sum3d(b0 * b0, b1 * b1, b2 * b2));
  float cos_theta = (float)((double)n1n3 / (na * nb));
  return (float)(acos((double)cos_theta) * 180 / 3.14159265358979323846);
}
// The float cosine compute_normal_angel feeds to acos, and threshold tests on it.  theta(c) =
// (float)(acos((double)c) * 180 / pi) is monotone non-increasing in the float c, so
//     theta <  thr   <=>   cut_lt <= c <= 1          (cut_lt = smallest float with theta(c) <  thr)
//     theta >  thr   <=>   -1 <= c < cut_le          (cut_le = smallest float with theta(c) <= thr)
// and theta is NaN exactly when c is NaN or outside [-1, 1] (every comparison with it is false, Q9).
// The cuts are found once on the host by bisection over the float bit patterns with the same
// expression (AngleCuts in fccf_internal.h), which removes the acos from every pairwise test.
__device__ __forceinline__ float normal_cos_n(float x1, float y1, float z1, double na, float x2, float y2, float z2, double nb) {
  double a0 = x1, a1 = y1, a2 = z1, b0 = x2, b1 = y2, b2 = z2;
  float n1n3 = (float)sum3d(a0 * b0, a1 * b1, a2 * b2);
  return (float)((double)n1n3 / (na * nb));
}
__device__ __forceinline__ double normal_norm(float x, float y, float z) { double a0 = x, a1 = y, a2 = z; return sqrt(sum3d(a0 * a0, a1 * a1, a2 * a2)); }
__device__ __forceinline__ float normal_cos(float x1, float y1, float z1, float x2, float y2, float z2) {
  return normal_cos_n(x1, y1, z1, normal_norm(x1, y1, z1), x2, y2, z2, normal_norm(x2, y2, z2));
}
__device__ __forceinline__ bool angle_lt(float c, float cut_lt) { return c >= cut_lt && c <= 1.0f; }          // compute_normal_angel(...) < thr
__device__ __forceinline__ bool angle_not_gt(float c, float cut_le) { return !(c >= -1.0f && c < cut_le); }  // !(compute_normal_angel(...) > thr)
// FCCF.cpp:379-389 with the threshold given as its cosine cut
__device__ __forceinline__ bool compare_normal_cut(float x1, float y1, float z1, float x2, float y2, float z2, float cut_le) {
  return angle_not_gt(normal_cos(x1, y1, z1, x2, y2, z2), cut_le);
}
// FCCF.cpp:379-389
__device__ __forceinline__ bool compare_normal(float x1, float y1, float z1, float x2, float y2, float z2, float thr) {
  float th = normal_angle(x1, y1, z1, x2, y2, z2);
  return !(th > thr);
}
// FCCF.cpp:391-407
__device__ __forceinline__ bool compare_plane(float nx1, float ny1, float nz1, float cx1, float cy1, float cz1,
                                              float nx2, float ny2, float nz2, float cx2, float cy2, float cz2, float l, float k) {
  float dx = cx1 - cx2, dy = cy1 - cy2, dz = cz1 - cz2;
  float vl = sqrtf(dx * dx + dy * dy + dz * dz);
  double e0 = (double)(dx / vl), e1 = (double)(dy / vl), e2 = (double)(dz / vl);
  float n1n3 = (float)fabs(sum3d((double)nx1 * e0, (double)ny1 * e1, (double)nz1 * e2));
  float n2n3 = (float)fabs(sum3d((double)nx2 * e0, (double)ny2 * e1, (double)nz2 * e2));
  float thr = l / (k * vl + 1);
  return (n1n3 < thr && n2n3 < thr);
}

__device__ __forceinline__ m3 mul33(const m3& a, const m3& b) {
  m3 r;
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) r.m[i][j] = sum3(a.m[i][0] * b.m[0][j], a.m[i][1] * b.m[1][j], a.m[i][2] * b.m[2][j]);
  return r;
}
__device__ __forceinline__ f3 mul3v(const m3& a, const f3& v) {
  return mk3(sum3(a.m[0][0] * v.x, a.m[0][1] * v.y, a.m[0][2] * v.z), sum3(a.m[1][0] * v.x, a.m[1][1] * v.y, a.m[1][2] * v.z),
             sum3(a.m[2][0] * v.x, a.m[2][1] * v.y, a.m[2][2] * v.z));
}
// cos*I + (1-cos)*r r^T + sin*[r]x   (FCCF.cpp:850-868)
__device__ __forceinline__ m3 rodrigues(float c, float s, const f3& r) {
  m3 R;
  float rv[3] = {r.x, r.y, r.z};
  float rx[3][3] = {{0.f, -r.z, r.y}, {r.z, 0.f, -r.x}, {-r.y, r.x, 0.f}};
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) R.m[i][j] = (c * ((i == j) ? 1.f : 0.f) + (1 - c) * (rv[i] * rv[j])) + s * rx[i][j];
  return R;
}
// Eigen compute_inverse<Matrix3f>
__device__ __forceinline__ float cof33(const m3& m, int i, int j) {
  int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
  return m.m[i1][j1] * m.m[i2][j2] - m.m[i1][j2] * m.m[i2][j1];
}
__device__ __forceinline__ m3 inverse33(const m3& m) {
  float c0 = cof33(m, 0, 0), c1 = cof33(m, 1, 0), c2 = cof33(m, 2, 0);
  float det = sum3(c0 * m.m[0][0], c1 * m.m[1][0], c2 * m.m[2][0]);
  float invdet = 1.0f / det;
  m3 r;
  r.m[0][0] = c0 * invdet; r.m[0][1] = c1 * invdet; r.m[0][2] = c2 * invdet;
  r.m[1][0] = cof33(m, 0, 1) * invdet; r.m[1][1] = cof33(m, 1, 1) * invdet; r.m[1][2] = cof33(m, 2, 1) * invdet;
  r.m[2][0] = cof33(m, 0, 2) * invdet; r.m[2][1] = cof33(m, 1, 2) * invdet; r.m[2][2] = cof33(m, 2, 2) * invdet;
  return r;
}
// Eigen Quaternionf(Matrix3f)
__device__ __forceinline__ q4 quat_from_matrix(const m3& mat) {
  float c[4];
  float t = sum3(mat.m[0][0], mat.m[1][1], mat.m[2][2]);
  if (t > 0.0f) {
    t = sqrtf(t + 1.0f);
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
    t = sqrtf(mat.m[i][i] - mat.m[j][j] - mat.m[k][k] + 1.0f);
    c[i] = 0.5f * t;
    t = 0.5f / t;
    c[3] = (mat.m[k][j] - mat.m[j][k]) * t;
    c[j] = (mat.m[j][i] + mat.m[i][j]) * t;
    c[k] = (mat.m[k][i] + mat.m[i][k]) * t;
  }
  q4 q; q.x = c[0]; q.y = c[1]; q.z = c[2]; q.w = c[3];
  return q;
}
// Eigen toRotationMatrix (no normalisation)
__device__ __forceinline__ m3 quat_to_matrix(const q4& q) {
  float tx = 2.f * q.x, ty = 2.f * q.y, tz = 2.f * q.z;
  float twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
  float txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
  float tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
  m3 r;
  r.m[0][0] = 1.f - (tyy + tzz); r.m[0][1] = txy - twz; r.m[0][2] = txz + twy;
  r.m[1][0] = txy + twz; r.m[1][1] = 1.f - (txx + tzz); r.m[1][2] = tyz - twx;
  r.m[2][0] = txz - twy; r.m[2][1] = tyz + twx; r.m[2][2] = 1.f - (txx + tyy);
  return r;
}
// Eigen _transformVector
__device__ __forceinline__ f3 quat_rotate(const q4& q, const f3& v) {
  f3 qv = mk3(q.x, q.y, q.z);
  f3 uv = cross(qv, v);
  uv.x += uv.x; uv.y += uv.y; uv.z += uv.z;
  f3 c2 = cross(qv, uv);
  return mk3((v.x + q.w * uv.x) + c2.x, (v.y + q.w * uv.y) + c2.y, (v.z + q.w * uv.z) + c2.z);
}
// pcl::detail::Transformer<float>::se3 / so3 on a row-major 3x4 (first 12 floats of a 4x4)
__device__ __forceinline__ f3 tf_se3(const float* T, const f3& p) {
  return mk3(p.x * T[0] + (p.y * T[1] + (p.z * T[2] + T[3])), p.x * T[4] + (p.y * T[5] + (p.z * T[6] + T[7])),
             p.x * T[8] + (p.y * T[9] + (p.z * T[10] + T[11])));
}
__device__ __forceinline__ f3 tf_so3(const float* T, const f3& n) {
  return mk3(n.x * T[0] + (n.y * T[1] + n.z * T[2]), n.x * T[4] + (n.y * T[5] + n.z * T[6]), n.x * T[8] + (n.y * T[9] + n.z * T[10]));
}
// two-axis rotation construction shared by FCCF.cpp:1148-1196 and 1306-1354
__device__ __forceinline__ m3 rotation_from_axes(f3 nt1, f3 nt2) {
  f3 ns1 = mk3(1, 0, 0), ns2 = mk3(0, 1, 0);
  f3 r1 = cross(ns1, nt1); normalize(r1);
  float c1 = dot(nt1, ns1);
  float s1 = dot(nt1, cross(r1, ns1));
  m3 R1 = rodrigues(c1, s1, r1);
  ns2 = mul3v(R1, ns2);
  f3 r2 = nt1;
  float ns2dnt2 = dot(ns2, nt2), ns2dr2 = dot(ns2, r2), nt2dr2 = dot(nt2, r2);
  f3 r2cns2 = cross(r2, ns2);
  float r2cns2dnt2 = dot(r2cns2, nt2);
  float c2 = (ns2dnt2 - (ns2dr2 * nt2dr2)) / (1 - (ns2dr2 * nt2dr2));
  float s2 = (r2cns2dnt2) / (1 - (ns2dr2 * nt2dr2));
  m3 R2 = rodrigues(c2, s2, r2);
  return mul33(R2, R1);
}

// ordered-int encoding of finite floats for atomicMin/atomicMax
__device__ __forceinline__ int f2ord(float f) { int i = __float_as_int(f); return i >= 0 ? i : (i ^ 0x7fffffff); }
__device__ __forceinline__ float ord2f(int i) { return __int_as_float(i >= 0 ? i : (i ^ 0x7fffffff)); }

// Warp-cooperative emulation of the reference's exchange sorts (range_face FCCF.cpp:409-427,
// range_cluster 1020-1038, score_range 1233-1251):  for i: for j>i: if key[i] < key[j] swap.
// Pass i moves the first maximum of the suffix to position i and shifts every strict
// left-to-right record of the suffix to the next record's position.  Called by one full warp;
// key[]/perm[] live in shared or global memory.  `less(a,b)` is the reference's comparison.
// npass: only the first npass passes (positions [0, npass) are then final, the rest is a permutation of the tail)
template <typename K, typename Less>
__device__ void warp_exchange_sort(K* key, int* perm, int n, Less less, int npass = 0x7fffffff) {
  const unsigned full = 0xffffffffu;
  int lane = threadIdx.x & 31;
  for (int i = 0; i + 1 < n && i < npass; i++) {
    K cur = key[i]; int curp = perm[i];
    bool moved = false;
    for (int base = i + 1; base < n; base += 32) {
      int j = base + lane;
      bool in = j < n;
      K v = in ? key[j] : cur; int vp = in ? perm[j] : 0;
      // records inside this window, resolved in lane order
      unsigned todo = __ballot_sync(full, in && less(cur, v));
      K wcur = cur; int wcurp = curp;
      while (todo) {
        int l = __ffs(todo) - 1;
        K lv = __shfl_sync(full, v, l); int lvp = __shfl_sync(full, vp, l);
        if (lane == l) { key[j] = wcur; perm[j] = wcurp; }
        wcur = lv; wcurp = lvp; moved = true;
        unsigned above = (l == 31) ? 0u : (full << (l + 1));
        todo = __ballot_sync(full, in && less(wcur, v)) & above;
      }
      cur = wcur; curp = wcurp;
    }
    if (moved && lane == 0) { key[i] = cur; perm[i] = curp; }
    __syncwarp();
  }
}

// Block-cooperative form of the same exchange sort for keys with a total order (ints, non-NaN floats).
// Pass i rotates the chain of strict left-to-right maxima of key[i..n): the maximum (first occurrence)
// goes to position i and every other record slot receives the previous record.  With the running
// maximum found by a block-wide prefix-max scan of (ordered key << 32 | ~position) a pass costs a few
// barriers, and the sort stops as soon as key[i..n) is non-increasing (all later passes are no-ops) —
// after about as many passes as there are elements above the minimum.  All threads of the block call
// it; key/perm in shared or global memory; s_scratch: 34 u64 of shared memory.
__device__ __forceinline__ unsigned ord_key(int v) { return (unsigned)v ^ 0x80000000u; }
__device__ __forceinline__ unsigned ord_key(float f) { unsigned b = __float_as_uint(f + 0.0f); return (b & 0x80000000u) ? ~b : (b | 0x80000000u); }
template <typename K>
__device__ void block_exchange_sort_passes(K* key, int* perm, int n, unsigned long long* s_scratch, int nt_) {
  const int t = threadIdx.x, nt = nt_, lane = t & 31, warp = t >> 5, nw = nt >> 5;
  unsigned long long* s_w = s_scratch;                 // [32] exclusive prefix max over warps
  unsigned long long* s_tot = s_scratch + 32;          // [1]  chunk maximum
  unsigned long long* s_carry = s_scratch + 33;        // [1]  running maximum (packed) ...
  K* s_ck = (K*)(s_scratch + 34);                      //      ... and the record it stands for
  int* s_cp = (int*)(s_scratch + 35);
  for (int i = 0; i + 1 < n; i++) {
    // stop when key[i..n) is non-increasing: every remaining pass would be a no-op
    int asc = 0;
    for (int j = i + t; j + 1 < n; j += nt) asc |= (key[j] < key[j + 1]) ? 1 : 0;
    if (!__syncthreads_or(asc)) break;
    if (t == 0) { *s_carry = ((unsigned long long)ord_key(key[i]) << 32) | (unsigned long long)(0xffffffffu - (unsigned)i); *s_ck = key[i]; *s_cp = perm[i]; }
    __syncthreads();
    for (int base = i + 1; base < n; base += nt) {
      const int j = base + t;
      const bool in = j < n;
      K my_k = K(); int my_p = 0;
      if (in) { my_k = key[j]; my_p = perm[j]; }
      const unsigned long long mine = in ? (((unsigned long long)ord_key(my_k) << 32) | (unsigned long long)(0xffffffffu - (unsigned)j)) : 0ull;
      unsigned long long inc = mine;                     // inclusive prefix max inside the warp
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) { unsigned long long y = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d && y > inc) inc = y; }
      if (lane == 31) s_w[warp] = inc;
      __syncthreads();
      if (warp == 0) {
        unsigned long long w = lane < nw ? s_w[lane] : 0ull, wi = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { unsigned long long y = __shfl_up_sync(0xffffffffu, wi, d); if (lane >= d && y > wi) wi = y; }
        unsigned long long ex = __shfl_up_sync(0xffffffffu, wi, 1);
        if (lane == 0) ex = 0ull;
        unsigned long long tot = __shfl_sync(0xffffffffu, wi, 31);
        s_w[lane] = ex;
        if (lane == 0) *s_tot = tot;
      }
      __syncthreads();
      const unsigned long long carry = *s_carry, tot = *s_tot;
      unsigned long long before = __shfl_up_sync(0xffffffffu, inc, 1);
      if (lane == 0) before = 0ull;
      if (s_w[warp] > before) before = s_w[warp];
      const bool from_carry = carry > before;            // the running maximum before j lies in an earlier chunk (or at i)
      if (from_carry) before = carry;
      const bool record = in && ((mine >> 32) > (before >> 32));   // strictly greater key
      K src_k = K(); int src_p = 0;
      if (record) {
        if (from_carry) { src_k = *s_ck; src_p = *s_cp; }
        else { const int pp = (int)(0xffffffffu - (unsigned)(before & 0xffffffffull)); src_k = key[pp]; src_p = perm[pp]; }
      }
      __syncthreads();                                   // all reads of this chunk are done
      if (record) { key[j] = src_k; perm[j] = src_p; }
      if (in && mine == tot && (tot >> 32) > (carry >> 32)) { *s_carry = mine; *s_ck = my_k; *s_cp = my_p; }   // this chunk's last record
      __syncthreads();
    }
    if (t == 0) { key[i] = *s_ck; perm[i] = *s_cp; }
    __syncthreads();
  }
}

// Pipelined form: positions are cut into tiles of XS_B; pass i works on tile q in global step q + i, so
// pass i+1 trails pass i by one tile and every compare/swap sees exactly the values of the sequential
// loops; one thread per pass (cur/perm in registers), one barrier per step.  Only the first
// P = #{key > min key} passes can move anything (afterwards key[P..n) holds only minimum keys), so the
// whole sort takes ceil(n / XS_B) + P steps instead of P passes of several barriers each.  Keys need a
// total order (ints, non-NaN floats).
#define XS_B 4
template <typename K>
// nt_: threads of the block that call (all of them still alive; a block whose surplus warps have exited passes the rest)
__device__ void block_exchange_sort(K* key, int* perm, int n, unsigned long long* s_scratch, int nt_ = 0) {
  const int t = threadIdx.x, nt = nt_ > 0 ? nt_ : (int)blockDim.x, lane = t & 31, warp = t >> 5, nw = nt >> 5;
  if (n < 2) return;
  // P: number of keys above the minimum
  unsigned* s_u = (unsigned*)s_scratch;              // [32] + [1]
  unsigned mn = 0xffffffffu;
  for (int j = t; j < n; j += nt) mn = min(mn, ord_key(key[j]));
  for (int o = 16; o; o >>= 1) mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
  if (lane == 0) s_u[warp] = mn;
  __syncthreads();
  if (warp == 0) { unsigned v = lane < nw ? s_u[lane] : 0xffffffffu; for (int o = 16; o; o >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o)); if (lane == 0) s_u[32] = v; }
  __syncthreads();
  mn = s_u[32];
  __syncthreads();
  int cnt = 0;
  for (int j = t; j < n; j += nt) cnt += (ord_key(key[j]) > mn) ? 1 : 0;
  for (int o = 16; o; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  if (lane == 0) s_u[warp] = (unsigned)cnt;
  __syncthreads();
  if (warp == 0) { unsigned v = lane < nw ? s_u[lane] : 0u; for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o); if (lane == 0) s_u[32] = v; }
  __syncthreads();
  int P = (int)s_u[32];
  __syncthreads();
  // A rare minimum (a handful of empty clusters among hundreds of singletons) would keep P near n although
  // almost every late pass is a no-op.  Then the pipelined passes stop at thr = the second smallest value
  // (P = #{key > thr}): afterwards key[P..n) holds only values <= thr, a pass whose key[i] == thr moves
  // nothing, and the few passes that start from a smaller key are run by one thread below, each ending as
  // soon as its running maximum reaches thr.
  unsigned thr = mn;
  if (P > 0 && (n - P) * 8 <= n) {
    unsigned m2 = 0xffffffffu;
    for (int j = t; j < n; j += nt) { unsigned o = ord_key(key[j]); if (o > mn) m2 = min(m2, o); }
    for (int o = 16; o; o >>= 1) m2 = min(m2, __shfl_xor_sync(0xffffffffu, m2, o));
    if (lane == 0) s_u[warp] = m2;
    __syncthreads();
    if (warp == 0) { unsigned v = lane < nw ? s_u[lane] : 0xffffffffu; for (int o = 16; o; o >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o)); if (lane == 0) s_u[32] = v; }
    __syncthreads();
    thr = s_u[32];
    __syncthreads();
    int c2 = 0;
    for (int j = t; j < n; j += nt) c2 += (ord_key(key[j]) > thr) ? 1 : 0;
    for (int o = 16; o; o >>= 1) c2 += __shfl_xor_sync(0xffffffffu, c2, o);
    if (lane == 0) s_u[warp] = (unsigned)c2;
    __syncthreads();
    if (warp == 0) { unsigned v = lane < nw ? s_u[lane] : 0u; for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o); if (lane == 0) s_u[32] = v; }
    __syncthreads();
    P = (int)s_u[32];
    __syncthreads();
  }
  if (P > n - 1) P = n - 1;
  if (P > 3 * nt) { block_exchange_sort_passes(key, perm, n, s_scratch, nt); return; }
  if (P <= 0 && thr == mn) return;
  // passes owned by this thread: t, t + nt, t + 2 nt.  Only the warps that own a pass take part in the
  // step loop (named barrier 1 over nact threads); the rest wait at the closing block barrier.
  const int nact = P >= nt ? nt : ((P + 31) & ~31);
  const int nslot = (P + nt - 1) / nt;
  if (P > 0 && t < nact) {
    K cur[3]; int curp[3];
#pragma unroll
    for (int u = 0; u < 3; u++) { cur[u] = K(); curp[u] = 0; }
    const int qmax = (n - 1) / XS_B;
    const int last = qmax + P - 1;
    for (int s = 0; s <= last; s++) {
#pragma unroll
      for (int u = 0; u < 3; u++) {
        const int i = t + u * nt;
        if (u < nslot && i < P) {
          const int q = s - i, q0 = (i + 1) / XS_B;
          if (q >= q0 && q <= qmax) {
            if (q == q0) { cur[u] = key[i]; curp[u] = perm[i]; }
            const int jb = q * XS_B;
            // the whole tile (keys and payloads) is fetched first, the compare/swap chain then runs on
            // registers, and only the positions it changed are written back
            K v[XS_B]; int vp[XS_B]; unsigned changed = 0u;
#pragma unroll
            for (int e = 0; e < XS_B; e++) { const int j = jb + e; if (j > i && j < n) { v[e] = key[j]; vp[e] = perm[j]; } }
#pragma unroll
            for (int e = 0; e < XS_B; e++) {
              const int j = jb + e;
              if (j > i && j < n && cur[u] < v[e]) {
                const K tk = v[e]; const int tp = vp[e];
                v[e] = cur[u]; vp[e] = curp[u]; cur[u] = tk; curp[u] = tp; changed |= 1u << e;
              }
            }
#pragma unroll
            for (int e = 0; e < XS_B; e++) if (changed & (1u << e)) { key[jb + e] = v[e]; perm[jb + e] = vp[e]; }
            if (q == qmax) { key[i] = cur[u]; perm[i] = curp[u]; }
          }
        }
      }
      asm volatile("bar.sync 1, %0;" ::"r"(nact) : "memory");
    }
  }
  __syncthreads();
  // passes P..n-2 that start below thr (see above), in order, by one thread
  if (thr != mn) {
    if (t == 0) {
      for (int i = P; i + 1 < n; i++) {
        K cur = key[i];
        if (ord_key(cur) >= thr) continue;
        int curp = perm[i];
        for (int j = i + 1; j < n; j++) {
          const K v = key[j];
          if (cur < v) { const int vp = perm[j]; key[j] = cur; perm[j] = curp; cur = v; curp = vp; if (ord_key(cur) >= thr) break; }
        }
        key[i] = cur; perm[i] = curp;
      }
    }
    __syncthreads();
  }
}

}  // namespace fccf
