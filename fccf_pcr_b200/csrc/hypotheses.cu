// hypotheses.cu — coplane-pair construction, pair-descriptor matching and closed-form hypothesis
// generation (FCCF.cpp:429-468 select_base, 1412-1427 match loop, 841-1018 computer_transform,
// 1439-1462 matrix -> quaternion), as count -> scan -> emit so that every pool keeps the
// reference's push_back order:
//   base_pairs    one CTA: base pairs of both clouds (ordered ballot compaction of the 16x16 upper triangle)
//   match_test    one thread per (pair of cloud 1, pair of cloud 2): the descriptor test
//   match_compact one CTA: ordered list of the matched (pair,pair) indices
//   hyp_count     one WARP per match: the (third plane of cloud 1) x (third plane of cloud 2) combinations
//                 are split over the lanes in the reference's loop order, ballots count them
//   match_scan    ordered scan of the counts per roughness type (three saturating 32-bit counters)
//   emit_hyp      one warp per match: every lane with a passing combination solves its translation and
//                 writes the hypothesis (3x4 + quaternion/translation) at pool offset + ballot rank
#include "fccf_dev.cuh"
#include "fccf_internal.h"
#include <vector>

namespace fccf {

struct HypArgs {
  PipeState* st;
  int* match_cnt; int* match_off; int* mlist;     // mlist: ordered indices of the matched (pair,pair) entries
  float* hyp_T; float* hyp_qt;
  int cap_hyp;
  float tmin, tmax, rough, same_thr, third_thr, third_cut;   // third_cut: cosine cut of third_plane_normal_threshold (strict <)
};

struct Plane { f3 c, n; float size; };
__device__ __forceinline__ Plane load_plane(const FaceTable& f, int i) {
  Plane p; p.c = mk3(f.plane[i][0], f.plane[i][1], f.plane[i][2]); p.n = mk3(f.plane[i][3], f.plane[i][4], f.plane[i][5]); p.size = f.plane[i][6];
  return p;
}

// FCCF.cpp:841-1018 for one matched pair of base pairs, by one warp.  EMIT=false: only count the
// hypotheses this match pushes.  The rotation (two Rodrigues steps) is computed by every lane; the
// third-plane combinations c = k3 * F2 + k2 (the reference's loop nest 936-1000) are tested 32 at a time,
// and a combination's output slot is its ballot rank, i.e. the reference's push_back order.
template <bool EMIT>
__device__ int hyp_generate_warp(const FaceTable& f1, const FaceTable& f2, int i11, int i12, int i21, int i22,
                                 float third_thr, float third_cut, float* outT, float* outQ, int lane) {
  Plane P11 = load_plane(f1, i11), P12 = load_plane(f1, i12), P21 = load_plane(f2, i21), P22 = load_plane(f2, i22);
  f3 n1 = P11.n, m1 = P12.n, n2 = P21.n, m2 = P22.n;
  f3 r1 = cross(n2, n1); normalize(r1);
  float n2dn1 = dot(n2, n1);
  f3 r1cn2 = cross(r1, n2);
  float r1cn2dn1 = dot(r1cn2, n1);
  m3 R1 = rodrigues(n2dn1, r1cn2dn1, r1);
  m2 = mul3v(R1, m2);
  f3 r2 = n1;
  float m2dm1 = dot(m2, m1), m2dr2 = dot(m2, r2), m1dr2 = dot(m1, r2);
  f3 r2cm2 = cross(r2, m2);
  float r2cm2dm1 = dot(r2cm2, m1);
  float cos2 = (m2dm1 - (m2dr2 * m1dr2)) / (1 - (m2dr2 * m1dr2));
  float sin2 = (r2cm2dm1) / (1 - (m2dr2 * m1dr2));
  m3 R2 = rodrigues(cos2, sin2, r2);
  m3 rot = mul33(R2, R1);
  float T[12];
#pragma unroll
  for (int i = 0; i < 3; i++) { T[4 * i] = rot.m[i][0]; T[4 * i + 1] = rot.m[i][1]; T[4 * i + 2] = rot.m[i][2]; T[4 * i + 3] = 0.f; }
  q4 q;
  if (EMIT) q = quat_from_matrix(rot);
  f3 n1cm1 = cross(n1, m1); normalize(n1cm1);
  f3 n2cm2 = cross(n2, m2); normalize(n2cm2);
  int count = 0;
  const int F1 = f1.F, F2 = f2.F, NC = F1 * F2;
  // Per-plane parts of the combination test, once per match: lane l < 16 holds candidate l of cloud 1
  // (normal, its double norm, "far enough from n1 x m1") and candidate l of cloud 2 (rotated normal, its
  // double norm, "far enough from n2 x m2"); the combinations fetch them with shuffles.
  f3 a_n = mk3(0, 0, 0), b_n = mk3(0, 0, 0); double a_nn = 1.0, b_nn = 1.0; int a_ok = 0, b_ok = 0;
  if (lane < F1) {
    a_n = mk3(f1.plane[lane][3], f1.plane[lane][4], f1.plane[lane][5]);
    a_nn = normal_norm(a_n.x, a_n.y, a_n.z);
    a_ok = (lane != i11 && lane != i12 && fabsf(dot(n1cm1, a_n)) > third_thr) ? 1 : 0;
  }
  if (lane < F2) {
    b_n = tf_so3(T, mk3(f2.plane[lane][3], f2.plane[lane][4], f2.plane[lane][5]));     // transformPointCloudWithNormals (FCCF.cpp:948), translation still 0
    b_nn = normal_norm(b_n.x, b_n.y, b_n.z);
    b_ok = (lane != i21 && lane != i22 && fabsf(dot(n2cm2, b_n)) > third_thr) ? 1 : 0;
  }
  for (int cb = 0; cb < NC; cb += 32) {
    const int c = cb + lane;
    const int cc = c < NC ? c : 0;
    const int k3 = cc / F2, k2 = cc - k3 * F2;
    const f3 pn = mk3(__shfl_sync(0xffffffffu, a_n.x, k3), __shfl_sync(0xffffffffu, a_n.y, k3), __shfl_sync(0xffffffffu, a_n.z, k3));
    const double n13 = __shfl_sync(0xffffffffu, a_nn, k3);
    const int ok3 = __shfl_sync(0xffffffffu, a_ok, k3);
    const f3 cn = mk3(__shfl_sync(0xffffffffu, b_n.x, k2), __shfl_sync(0xffffffffu, b_n.y, k2), __shfl_sync(0xffffffffu, b_n.z, k2));
    const double n2n = __shfl_sync(0xffffffffu, b_nn, k2);
    const int ok2 = __shfl_sync(0xffffffffu, b_ok, k2);
    bool ok = false;
    if (c < NC && ok3 && ok2) {
      // compute_normal_angel(k1, k2) < third_plane_normal_threshold (FCCF.cpp:949,958) through its cosine cut
      float c3 = normal_cos_n(pn.x, pn.y, pn.z, n13, cn.x, cn.y, cn.z, n2n);
      ok = angle_lt(c3, third_cut);
    }
    Plane P13, P2;
    if (EMIT && ok) { P13 = load_plane(f1, k3); P2 = load_plane(f2, k2); }
    const unsigned bal = __ballot_sync(0xffffffffu, ok);
    if (EMIT && ok) {
      const int slot = count + __popc(bal & ((1u << lane) - 1u));
      f3 c23 = tf_se3(T, P2.c);
      f3 k1 = P13.n, kk2 = cn;
      float d11 = dot(P11.c, n1), d12 = dot(P12.c, m1), d13 = dot(P13.c, k1);
      float d21 = dot(P21.c, n2), d22 = dot(P22.c, m2), d23 = dot(c23, kk2);
      f3 D = mk3(d11 - d21, d12 - d22, d13 - d23);
      m3 Am; Am.m[0][0] = n1.x; Am.m[0][1] = n1.y; Am.m[0][2] = n1.z; Am.m[1][0] = m1.x; Am.m[1][1] = m1.y; Am.m[1][2] = m1.z;
      Am.m[2][0] = k1.x; Am.m[2][1] = k1.y; Am.m[2][2] = k1.z;
      m3 AT;
#pragma unroll
      for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) AT.m[i][j] = Am.m[j][i];
      f3 Tt = mul3v(mul33(inverse33(mul33(AT, Am)), AT), D);
      float* o = outT + (size_t)slot * 12;
#pragma unroll
      for (int i = 0; i < 12; i++) o[i] = T[i];
      o[3] = Tt.x; o[7] = Tt.y; o[11] = Tt.z;
      float* oq = outQ + (size_t)slot * 8;
      oq[0] = q.w; oq[1] = q.x; oq[2] = q.y; oq[3] = q.z; oq[4] = Tt.x; oq[5] = Tt.y; oq[6] = Tt.z; oq[7] = 0.f;
    }
    count += __popc(bal);
  }
  if (count == 0) {
    if (EMIT && lane == 0) {
      float sa = P11.size, sb = P12.size, sc = P21.size, sd = P22.size;
      float sx = (P11.c.x * sa + P12.c.x * sb) / (sa + sb);
      float sy = (P11.c.y * sa + P12.c.y * sb) / (sa + sb);
      float sz = (P11.c.z * sa + P12.c.z * sb) / (sa + sb);
      float tx = (P21.c.x * sc + P22.c.x * sd) / (sc + sd);
      float ty = (P21.c.y * sc + P22.c.y * sd) / (sc + sd);
      float tz = (P21.c.z * sc + P22.c.z * sd) / (sc + sd);
      f3 tc = mul3v(rot, mk3(tx, ty, tz));
#pragma unroll
      for (int i = 0; i < 12; i++) outT[i] = T[i];
      outT[3] = sx - tc.x; outT[7] = sy - tc.y; outT[11] = sz - tc.z;
      outQ[0] = q.w; outQ[1] = q.x; outQ[2] = q.y; outQ[3] = q.z; outQ[4] = outT[3]; outQ[5] = outT[7]; outQ[6] = outT[11]; outQ[7] = 0.f;
    }
    count = 1;
  }
  return count;
}

__global__ void __launch_bounds__(256) base_pairs_kernel(const HypArgs* __restrict__ AB) {
  FCCF_PDL_ENTER();
  const HypArgs& A = AB[blockIdx.z];
  PipeState* st = A.st;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  __shared__ int s_w[8];
  // ---- select_base for both clouds (FCCF.cpp:429-468) ----
  for (int c = 0; c < 2; c++) {
    const FaceTable& f = st->ft[c];
    BaseTable& B = st->base[c];
    const int F = f.F;
    bool ok = false; float angel = 0.f; int a = t >> 4, b = t & 15;
    if (t < 256 && a < b && b < F) {
      angel = normal_angle(f.plane[a][3], f.plane[a][4], f.plane[a][5], f.plane[b][3], f.plane[b][4], f.plane[b][5]);
      ok = (A.tmin < angel && angel < A.tmax);
    }
    unsigned bal = __ballot_sync(0xffffffffu, ok);
    if (lane == 0) s_w[warp] = __popc(bal);
    __syncthreads();
    int off = 0;
    for (int w2 = 0; w2 < warp; w2++) off += s_w[w2];
    if (ok) {
      int k = off + __popc(bal & ((1u << lane) - 1u));
      B.i[k] = a; B.j[k] = b; B.angle[k] = angel;
      double th1 = (double)A.rough, ta = f.theta[a], tb = f.theta[b];
      int ty = 3;
      if (ta <= th1 && tb <= th1) ty = 0;
      else if (ta > th1 && tb > th1) ty = 1;
      else if (ta <= th1 && tb > th1) ty = 2;
      else if (ta > th1 && tb <= th1) ty = 2;
      B.type[k] = ty;
    }
    if (t == 0) { int tot = 0; for (int w2 = 0; w2 < 8; w2++) tot += s_w[w2]; B.B = tot; }
    __syncthreads();
  }
}

// ---- match loop (FCCF.cpp:1415-1427): one thread per (pair of cloud 1, pair of cloud 2): descriptor test
// (included angle within 5 degrees, same roughness type).  match_cnt: 0 / 1 = matched (counted later) ----
__global__ void __launch_bounds__(128) match_test_kernel(const HypArgs* __restrict__ AB) {
  FCCF_PDL_ENTER();
  const HypArgs& A = AB[blockIdx.z];
  PipeState* st = A.st;
  const int B1 = st->base[0].B, B2 = st->base[1].B;
  const int NM = B1 * B2;
  const BaseTable &b1 = st->base[0], &b2 = st->base[1];
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < NM; idx += gridDim.x * blockDim.x) {
    int i1 = idx / B2, i2 = idx - i1 * B2;
    A.match_cnt[idx] = (fabsf(b1.angle[i1] - b2.angle[i2]) < A.same_thr && b1.type[i1] == b2.type[i2] && b1.type[i1] < 3) ? 1 : 0;
  }
}

// ---- ordered list of the matched entries (one CTA) ----
// Every thread owns a run of consecutive entries (at most 15: B1 x B2 <= FCCF_MAXMATCH): it counts its run, ONE block
// scan places the runs, it writes its entries — two barriers for the whole list instead of three per 1024 entries.
__global__ void __launch_bounds__(1024) match_compact_kernel(const HypArgs* __restrict__ AB) {
  FCCF_PDL_ENTER();
  const HypArgs& A = AB[blockIdx.z];
  PipeState* st = A.st;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  __shared__ int s_w[32];
  const int NM = st->base[0].B * st->base[1].B;
  const int per = (NM + 1023) / 1024;
  const int b0 = min(NM, t * per), b1 = min(NM, b0 + per);
  int cnt = 0;
  for (int k = b0; k < b1; k++) cnt += (A.match_cnt[k] != 0) ? 1 : 0;
  int inc = cnt;
  for (int d = 1; d < 32; d <<= 1) { const int y = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += y; }
  if (lane == 31) s_w[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    const int wv = s_w[lane]; int wi = wv;
    for (int d = 1; d < 32; d <<= 1) { const int y = __shfl_up_sync(0xffffffffu, wi, d); if (lane >= d) wi += y; }
    s_w[lane] = wi - wv;
    if (lane == 31) { st->n_match = NM; st->tickets[22] = wi; }     // tickets[22]: number of matched entries
  }
  __syncthreads();
  int off = s_w[warp] + inc - cnt;
  for (int k = b0; k < b1; k++) if (A.match_cnt[k] != 0) A.mlist[off++] = k;
}

// ---- hypotheses per match: one warp per matched entry ----
__global__ void __launch_bounds__(128) hyp_count_kernel(const HypArgs* __restrict__ AB) {
  FCCF_PDL_ENTER();
  const HypArgs& A = AB[blockIdx.z];
  PipeState* st = A.st;
  const int lane = threadIdx.x & 31;
  const int B2 = st->base[1].B, M = st->tickets[22];
  const BaseTable &b1 = st->base[0], &b2 = st->base[1];
  for (int m = blockIdx.x * 4 + (threadIdx.x >> 5); m < M; m += gridDim.x * 4) {
    const int idx = A.mlist[m];
    const int i1 = idx / B2, i2 = idx - i1 * B2;
    int cnt = hyp_generate_warp<false>(st->ft[0], st->ft[1], b1.i[i1], b1.j[i1], b2.i[i2], b2.j[i2], A.third_thr, A.third_cut, nullptr, nullptr, lane);
    if (lane == 0) A.match_cnt[idx] = cnt;
  }
}

// ---- ordered scan of the counts per type: pool offsets in the reference's push_back order ----
// Three 32-bit counters side by side (one per roughness type), each saturating at cap_hyp + 1: a pool that
// outgrows the arena is reported (ST_HYP_OVERFLOW) instead of wrapping into its neighbour.
struct u3 { u32 a[3]; };
__device__ __forceinline__ u32 sat_add(u32 x, u32 y, u32 lim) { const u32 s = x + y; return (s < x || s > lim) ? lim : s; }
__global__ void __launch_bounds__(1024) match_scan_kernel(const HypArgs* __restrict__ AB) {
  FCCF_PDL_ENTER();
  const HypArgs& A = AB[blockIdx.z];
  PipeState* st = A.st;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  __shared__ u32 s_w[3][32];
  __shared__ u32 s_tot[3];
  const int B1 = st->base[0].B, B2 = st->base[1].B;
  const int NM = B1 * B2;
  const u32 lim = (u32)A.cap_hyp + 1u;
  // every thread owns a run of consecutive entries (at most 15); saturating sums are associative (min(lim, a + b) of
  // non-negative counts), so the run totals go through one block scan per type and the run is walked once more
  const int per = (NM + 1023) / 1024;
  const int b0 = min(NM, t * per), b1 = min(NM, b0 + per);
  u32 x[3] = {0u, 0u, 0u};
  for (int k = b0; k < b1; k++) { const int cnt = A.match_cnt[k]; if (cnt > 0) { const int ty = st->base[0].type[k / B2]; x[ty] = sat_add(x[ty], (u32)cnt, lim); } }
  u32 inc[3] = {x[0], x[1], x[2]};
#pragma unroll
  for (int k = 0; k < 3; k++) {
    for (int d = 1; d < 32; d <<= 1) { u32 y = __shfl_up_sync(0xffffffffu, inc[k], d); if (lane >= d) inc[k] = sat_add(inc[k], y, lim); }
    if (lane == 31) s_w[k][warp] = inc[k];
  }
  __syncthreads();
  if (warp < 3) {
    u32 wv = s_w[warp][lane], wi = wv;
    for (int d = 1; d < 32; d <<= 1) { u32 y = __shfl_up_sync(0xffffffffu, wi, d); if (lane >= d) wi = sat_add(wi, y, lim); }
    u32 ex = __shfl_up_sync(0xffffffffu, wi, 1);
    s_w[warp][lane] = lane ? ex : 0u;
    if (lane == 31) s_tot[warp] = wi;
  }
  __syncthreads();
  u32 run[3];
#pragma unroll
  for (int k = 0; k < 3; k++) {
    // exclusive prefix of this thread's run: warps in front + lanes in front (inclusive scan of the lane in front)
    u32 before = __shfl_up_sync(0xffffffffu, inc[k], 1);
    if (lane == 0) before = 0u;
    run[k] = sat_add(s_w[k][warp], before, lim);
  }
  for (int k = b0; k < b1; k++) {
    const int cnt = A.match_cnt[k];
    if (cnt > 0) { const int ty = st->base[0].type[k / B2]; A.match_off[k] = (int)run[ty]; run[ty] = sat_add(run[ty], (u32)cnt, lim); }
  }
  if (t == 0) {
    u32 n0 = s_tot[0], n1 = s_tot[1], n2 = s_tot[2];
    if ((unsigned long long)n0 + n1 + n2 > (unsigned long long)A.cap_hyp) { atomicOr(&st->status, ST_HYP_OVERFLOW); n0 = n1 = n2 = 0; }
    st->n_hyp[0] = (int)n0; st->n_hyp[1] = (int)n1; st->n_hyp[2] = (int)n2;
    st->hyp_off[0] = 0; st->hyp_off[1] = (int)n0; st->hyp_off[2] = (int)(n0 + n1); st->hyp_off[3] = (int)(n0 + n1 + n2);
    st->n_match = NM;
  }
}

__global__ void __launch_bounds__(128) emit_hyp_kernel(const HypArgs* __restrict__ AB) {
  FCCF_PDL_ENTER();
  const HypArgs& A = AB[blockIdx.z];
  PipeState* st = A.st;
  const int lane = threadIdx.x & 31;
  const int B2 = st->base[1].B, M = st->tickets[22];
  if (st->hyp_off[3] == 0) return;
  const BaseTable &b1 = st->base[0], &b2 = st->base[1];
  for (int m = blockIdx.x * 4 + (threadIdx.x >> 5); m < M; m += gridDim.x * 4) {
    const int idx = A.mlist[m];
    const int i1 = idx / B2, i2 = idx - i1 * B2;
    const int ty = b1.type[i1];
    const size_t off = (size_t)st->hyp_off[ty] + A.match_off[idx];
    hyp_generate_warp<true>(st->ft[0], st->ft[1], b1.i[i1], b1.j[i1], b2.i[i2], b2.j[i2], A.third_thr, A.third_cut, A.hyp_T + off * 12, A.hyp_qt + off * 8, lane);
  }
}

void launch_hypotheses(cudaStream_t s, const Batch& b, uint64_t* launches) {
  const int G = b.G;
  std::vector<HypArgs> As(G);
  for (int g = 0; g < G; g++) {
    const Work& w = b.w[g]; const HypWS& h = w.h;
    HypArgs& A = As[g];
    A.st = w.st; A.match_cnt = h.match_cnt; A.match_off = h.match_off; A.mlist = h.c_state;   // c_state is free until clustering
    A.hyp_T = h.hyp_T; A.hyp_qt = h.hyp_qt; A.cap_hyp = h.cap_hyp;
    A.tmin = b.p.included_angle_min_threshold; A.tmax = b.p.included_angle_max_threshold; A.rough = b.p.rough_threshold_gl;
    A.same_thr = b.p.included_angle_same_threshold; A.third_thr = b.p.third_plane_threshold; A.third_cut = b.cuts.third_lt;
  }
  const HypArgs* dA = b.tab->put(As.data(), G);
  const int nbw = grid_x((FCCF_MAXMATCH + 3) / 4, G);     // warp-per-match kernels: 4 warps per CTA, strided over the list
  klaunch(base_pairs_kernel, dim3(dim3(1, 1, G)), dim3(256), 0, s, dA);
  klaunch(match_test_kernel, dim3(dim3(grid_x((FCCF_MAXMATCH + 127) / 128, G), 1, G)), dim3(128), 0, s, dA);
  klaunch(match_compact_kernel, dim3(dim3(1, 1, G)), dim3(1024), 0, s, dA);
  klaunch(hyp_count_kernel, dim3(dim3(nbw, 1, G)), dim3(128), 0, s, dA);
  klaunch(match_scan_kernel, dim3(dim3(1, 1, G)), dim3(1024), 0, s, dA);
  klaunch(emit_hyp_kernel, dim3(dim3(nbw, 1, G)), dim3(128), 0, s, dA);
  if (launches) *launches += 6;
}

}  // namespace fccf
