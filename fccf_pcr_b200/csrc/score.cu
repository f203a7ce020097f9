// score.cu — fine_verify (FCCF.cpp:785-839) as "hypothesis scoring", per-type best + gate
// (FCCF.cpp:1546-1605) and fuse_answer (1291-1368).
//
// The reference rebuilds a pcl octree (0.5 m) over static + transformed moving leftover points for
// every hypothesis and counts both kinds per voxel.  The voxel lattice is anchored on the first
// static point (pcl's first insert), so it does not depend on the hypothesis: a static
// open-addressing hash of the occupied static voxels (64-bit packed lattice coordinates, linear
// probing) is built once; scoring a hypothesis is then, per moving point, a float32 3x4 transform
// (pcl::transformPointCloud's SSE order), a float64 lattice key, one probe and one counter
// increment.  One CTA scores one hypothesis at a time with its counters in shared memory (global
// fallback for very large tables); the per-voxel integer counts (s,t) are exactly the reference's.
#include "fccf_dev.cuh"
#include "fccf_internal.h"

namespace fccf {

#define SC_EMPTY 0xffffffffffffffffull
#define SC_OFF (1 << 20)
#define SC_THREADS 512
#define SC_SMEM_SLOTS 32768      // 128 KB of u32 counters

__device__ __forceinline__ u64 sc_hash(u64 k) { k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ull; k ^= k >> 33; return k; }
__device__ __forceinline__ bool sc_key(const ScoreState* ss, double inv_res_is_div, float x, float y, float z, u64& key) {
  // lattice coordinate = floor((double(p) - mn) / res); pcl: key = (unsigned)((p - min)/res) on a box whose
  // min is mn shifted by whole voxels
  long long lx = (long long)floor(((double)x - ss->mn[0]) / inv_res_is_div);
  long long ly = (long long)floor(((double)y - ss->mn[1]) / inv_res_is_div);
  long long lz = (long long)floor(((double)z - ss->mn[2]) / inv_res_is_div);
  lx += SC_OFF; ly += SC_OFF; lz += SC_OFF;
  if ((unsigned long long)lx >= (1ull << 21) || (unsigned long long)ly >= (1ull << 21) || (unsigned long long)lz >= (1ull << 21)) return false;
  key = ((u64)lx << 42) | ((u64)ly << 21) | (u64)lz;
  return true;
}

struct ScArgs {
  ScoreWS ws;
  const float* s1; const int* n1p; const int* n2p;
  const float* s2;
  const float* T; int n_hyp; const int* n_top;   // n_top != nullptr: pipeline layout [3][FCCF_TOPK]
  float* scores;
  int* rows; int cap_rows; int* nrows;
  float res;
};

__global__ void score_setup_kernel(const __grid_constant__ ScArgs A) {
  ScoreState* ss = A.ws.ss;
  int n1 = *A.n1p, n2 = *A.n2p;
  ss->n1 = n1; ss->n2 = n2; ss->used = 0;
  double res = (double)A.res;
  if (n1 > 0) {
    // first insert into an empty pcl octree: box = p0 +- res/2, then padded to depth 1 (getKeyBitSize)
    const float minValue = 1.1920928955078125e-07f;
    for (int a = 0; a < 3; a++) {
      double mn = (double)A.s1[a] - res / 2, mx = (double)A.s1[a] + res / 2;
      double side = 2.0 * res;
      double over = (side - (mx - mn)) / 2.0;
      if (over > minValue) mn -= over;
      ss->mn[a] = mn;
    }
  } else { ss->mn[0] = ss->mn[1] = ss->mn[2] = 0.0; }
  int want = 64;
  while (want < 2 * n1 && want < A.ws.cap_hash) want <<= 1;
  if (want > A.ws.cap_hash) want = A.ws.cap_hash;
  ss->cap_eff = want;
}
__global__ void __launch_bounds__(256) score_clear_kernel(const __grid_constant__ ScArgs A) {
  const int cap = A.ws.ss->cap_eff;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < cap; i += gridDim.x * blockDim.x) { A.ws.keys[i] = SC_EMPTY; A.ws.s_cnt[i] = 0; }
}
__global__ void __launch_bounds__(256) score_insert_kernel(const __grid_constant__ ScArgs A) {
  const ScoreState* ss = A.ws.ss;
  const int n1 = ss->n1, cap = ss->cap_eff;
  const unsigned mask = (unsigned)cap - 1u;
  const double res = (double)A.res;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n1; i += gridDim.x * blockDim.x) {
    u64 key;
    if (!sc_key(ss, res, A.s1[3 * i], A.s1[3 * i + 1], A.s1[3 * i + 2], key)) { atomicOr(A.ws.status, ST_HASH_FULL); continue; }
    unsigned h = (unsigned)sc_hash(key) & mask;
    for (int probe = 0; probe < cap; probe++) {
      u64 prev = atomicCAS((unsigned long long*)&A.ws.keys[h], SC_EMPTY, key);
      if (prev == SC_EMPTY || prev == key) { atomicAdd(&A.ws.s_cnt[h], 1); break; }
      h = (h + 1) & mask;
      if (probe == cap - 1) atomicOr(A.ws.status, ST_HASH_FULL);
    }
  }
}

// scores one hypothesis with the whole CTA; cnt: cap counters (shared or global)
__device__ float score_one(const ScArgs& A, const float* T16, int* cnt, float* s_red) {
  const ScoreState* ss = A.ws.ss;
  const int cap = ss->cap_eff, n2 = ss->n2, n1 = ss->n1;
  const unsigned mask = (unsigned)cap - 1u;
  const int t = threadIdx.x;
  const double res = (double)A.res;
  float T[12];
#pragma unroll
  for (int i = 0; i < 12; i++) T[i] = T16[i];
  for (int h = t; h < cap; h += SC_THREADS) cnt[h] = 0;
  __syncthreads();
  const u64* __restrict__ keys = A.ws.keys;
  for (int i = t; i < n2; i += SC_THREADS) {
    f3 q = tf_se3(T, mk3(A.s2[3 * i], A.s2[3 * i + 1], A.s2[3 * i + 2]));
    u64 key;
    if (!sc_key(ss, res, q.x, q.y, q.z, key)) continue;
    unsigned h = (unsigned)sc_hash(key) & mask;
    while (true) {
      u64 k = __ldg(&keys[h]);
      if (k == key) { atomicAdd(&cnt[h], 1); break; }
      if (k == SC_EMPTY) break;
      h = (h + 1) & mask;
    }
  }
  __syncthreads();
  float part = 0.f;
  for (int h = t; h < cap; h += SC_THREADS) {
    int tt = cnt[h];
    if (tt > 0) {
      float fs = (float)A.ws.s_cnt[h], ft = (float)tt;
      float mn = fs < ft ? fs : ft, mx = fs > ft ? fs : ft;
      part = part + (fs + ft) * (mn / mx);
    }
  }
  for (int o = 16; o; o >>= 1) part = part + __shfl_xor_sync(0xffffffffu, part, o);
  if ((t & 31) == 0) s_red[t >> 5] = part;
  __syncthreads();
  float tot = 0.f;
  if (t < 32) {
    float v = (t < SC_THREADS / 32) ? s_red[t] : 0.f;
    for (int o = 16; o; o >>= 1) v = v + __shfl_xor_sync(0xffffffffu, v, o);
    tot = v;
  }
  __syncthreads();
  return tot / (float)(n1 + n2);   // valid on warp 0
}

extern __shared__ int sc_dyn[];
template <bool SMEM>
__global__ void __launch_bounds__(SC_THREADS) score_kernel(const __grid_constant__ ScArgs A) {
  __shared__ float s_red[SC_THREADS / 32];
  const int cap = A.ws.ss->cap_eff;
  const bool use_smem = SMEM && cap <= SC_SMEM_SLOTS;
  if (!use_smem && (int)blockIdx.x >= A.ws.t_rows) return;
  int* cnt = use_smem ? sc_dyn : (A.ws.t_cnt + (size_t)blockIdx.x * A.ws.cap_hash);
  const int stride = use_smem ? gridDim.x : min((int)gridDim.x, A.ws.t_rows);
  for (int h = blockIdx.x; h < A.n_hyp; h += stride) {
    if (A.n_top) { int ty = h / FCCF_TOPK, k = h - ty * FCCF_TOPK; if (k >= A.n_top[ty]) continue; }
    float sc = score_one(A, A.T + (size_t)h * 16, cnt, s_red);
    if (threadIdx.x == 0) A.scores[h] = sc;
  }
}

__global__ void __launch_bounds__(SC_THREADS) score_dump_kernel(const __grid_constant__ ScArgs A) {
  __shared__ float s_red[SC_THREADS / 32];
  const ScoreState* ss = A.ws.ss;
  int* cnt = A.ws.t_cnt;
  score_one(A, A.T, cnt, s_red);
  __syncthreads();
  const int cap = ss->cap_eff;
  for (int h = threadIdx.x; h < cap; h += SC_THREADS) {
    int tt = cnt[h];
    if (tt > 0) {
      int r = atomicAdd(A.nrows, 1);
      if (r < A.cap_rows) {
        u64 k = A.ws.keys[h];
        // rows are relative to the voxel of the first static point (lattice coordinate 1)
        A.rows[5 * r] = (int)((k >> 42) & 0x1fffff) - SC_OFF - 1; A.rows[5 * r + 1] = (int)((k >> 21) & 0x1fffff) - SC_OFF - 1; A.rows[5 * r + 2] = (int)(k & 0x1fffff) - SC_OFF - 1;
        A.rows[5 * r + 3] = A.ws.s_cnt[h]; A.rows[5 * r + 4] = tt;
      }
    }
  }
}

static void fill_common(ScArgs& A, const fccf_params& p, const ScoreWS& ws) {
  A.ws = ws; A.res = p.fine_verify_voxel_size; A.s1 = nullptr; A.n1p = nullptr; A.n2p = nullptr; A.s2 = nullptr; A.T = nullptr; A.n_hyp = 0; A.n_top = nullptr;
  A.scores = nullptr; A.rows = nullptr; A.cap_rows = 0; A.nrows = nullptr;
}

void launch_score_build(cudaStream_t s, const fccf_params& p, const float* d_s1, const int* d_n1, const int* d_n2, int cap_points, const ScoreWS& ws, uint64_t* launches) {
  ScArgs A; fill_common(A, p, ws);
  A.s1 = d_s1; A.n1p = d_n1; A.n2p = d_n2;
  score_setup_kernel<<<1, 1, 0, s>>>(A);
  int nb = (ws.cap_hash + 255) / 256; if (nb > 1184) nb = 1184;
  score_clear_kernel<<<nb, 256, 0, s>>>(A);
  int nbi = (cap_points + 255) / 256; if (nbi > 1184) nbi = 1184; if (nbi < 1) nbi = 1;
  score_insert_kernel<<<nbi, 256, 0, s>>>(A);
  if (launches) *launches += 3;
}

static void score_launch(cudaStream_t s, ScArgs& A, int nblocks) {
  static bool attr = false;
  if (!attr) { cudaFuncSetAttribute(score_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SC_SMEM_SLOTS * 4); attr = true; }
  score_kernel<true><<<nblocks, SC_THREADS, SC_SMEM_SLOTS * 4, s>>>(A);
}

void launch_score_list(cudaStream_t s, const fccf_params& p, const float* d_T16, int n_hyp, const float* d_s2, const ScoreWS& ws, float* d_scores, uint64_t* launches) {
  if (n_hyp <= 0) return;
  ScArgs A; fill_common(A, p, ws);
  A.T = d_T16; A.n_hyp = n_hyp; A.s2 = d_s2; A.scores = d_scores;
  int nb = n_hyp < 148 ? n_hyp : 148;
  score_launch(s, A, nb);
  if (launches) *launches += 1;
}

void launch_score_dump(cudaStream_t s, const fccf_params& p, const float* d_T16, const float* d_s2, const ScoreWS& ws, int* d_rows, int cap_rows, int* d_nrows, uint64_t* launches) {
  ScArgs A; fill_common(A, p, ws);
  A.T = d_T16; A.n_hyp = 1; A.s2 = d_s2; A.rows = d_rows; A.cap_rows = cap_rows; A.nrows = d_nrows;
  score_dump_kernel<<<1, SC_THREADS, 0, s>>>(A);
  if (launches) *launches += 1;
}

// ---------------------------------------------------------------------------------------------
struct FuseArgs { PipeState* st; const float* top_T; const float* top_s1; const float* top_s2; float fine_number; };

// FCCF.cpp:1546-1606 + fuse_answer 1291-1368
__global__ void fuse_kernel(const __grid_constant__ FuseArgs A) {
  PipeState* st = A.st;
  float score_sum = 0.f, score1_sum = 0.f, score2_sum = 0.f;
  for (int ty = 0; ty < 3; ty++)
    for (int k = 0; k < st->n_top[ty]; k++) { score2_sum += A.top_s2[ty * FCCF_TOPK + k]; score1_sum += A.top_s1[ty * FCCF_TOPK + k]; }
  float best_best = 0.f;
  float hs_score[3]; q4 hs_q[3]; f3 hs_t[3];
  for (int ty = 0; ty < 3; ty++) {
    float best_score = 0.f;
    float tb[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
    for (int k = 0; k < st->n_top[ty]; k++) {
      float score = A.top_s1[ty * FCCF_TOPK + k] / score1_sum + A.top_s2[ty * FCCF_TOPK + k] / score2_sum;
      if (score > best_score) { best_score = score; for (int i = 0; i < 12; i++) tb[i] = A.top_T[((size_t)ty * FCCF_TOPK + k) * 16 + i]; }
    }
    if (best_best < best_score) best_best = best_score;
    m3 R; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) R.m[i][j] = tb[4 * i + j];
    hs_q[ty] = quat_from_matrix(R); hs_t[ty] = mk3(tb[3], tb[7], tb[11]); hs_score[ty] = best_score;
    st->type_best[ty][0] = best_score; for (int i = 0; i < 12; i++) st->type_best[ty][1 + i] = tb[i];
  }
  bool keep[3];
  for (int ty = 0; ty < 3; ty++) { keep[ty] = ((double)hs_score[ty] > (double)best_best * 0.8); if (keep[ty]) score_sum += hs_score[ty]; }
  float ax = 0, ay = 0, az = 0;
  for (int ty = 0; ty < 3; ty++) if (keep[ty]) {
    float w = hs_score[ty] / score_sum;
    ax = ax + hs_t[ty].x * w; ay = ay + hs_t[ty].y * w; az = az + hs_t[ty].z * w;
  }
  float s1x = 0, s1y = 0, s1z = 0, s2x = 0, s2y = 0, s2z = 0;
  for (int ty = 0; ty < 3; ty++) if (keep[ty]) {
    f3 a = quat_rotate(hs_q[ty], mk3(1, 0, 0)), b = quat_rotate(hs_q[ty], mk3(0, 1, 0));
    float w = hs_score[ty] / score_sum;
    s1x = s1x + a.x * w; s1y = s1y + a.y * w; s1z = s1z + a.z * w; s2x = s2x + b.x * w; s2y = s2y + b.y * w; s2z = s2z + b.z * w;
  }
  f3 a1 = mk3(s1x, s1y, s1z), a2 = mk3(s2x, s2y, s2z);
  normalize(a1); normalize(a2);
  m3 R = rotation_from_axes(a1, a2);
  float* T = st->T_final;
  for (int i = 0; i < 3; i++) { T[4 * i] = R.m[i][0]; T[4 * i + 1] = R.m[i][1]; T[4 * i + 2] = R.m[i][2]; }
  T[3] = ax; T[7] = ay; T[11] = az;
  T[12] = 0.f; T[13] = 0.f; T[14] = 0.f; T[15] = 1.f;
}

void launch_fine_verify_fuse(cudaStream_t s, const Work& w, const HypWS& h, uint64_t* launches) {
  PipeState* st = w.st;
  ScoreWS ws; ws.keys = h.fv_keys; ws.s_cnt = h.fv_s; ws.t_cnt = h.fv_t; ws.cap_hash = h.cap_hash; ws.t_rows = 3 * FCCF_TOPK; ws.ss = &st->fv; ws.status = &st->status;
  int cap_pts = w.c[0].cap;
  launch_score_build(s, w.p, w.c[0].sub, &st->oct[0].S, &st->oct[1].S, cap_pts, ws, launches);
  ScArgs A; fill_common(A, w.p, ws);
  A.T = h.top_T; A.n_hyp = 3 * FCCF_TOPK; A.n_top = st->n_top; A.s2 = w.c[1].sub; A.scores = h.top_s2;
  score_launch(s, A, 3 * FCCF_TOPK);
  FuseArgs F; F.st = st; F.top_T = h.top_T; F.top_s1 = h.top_s1; F.top_s2 = h.top_s2; F.fine_number = w.p.fine_verify_number;
  fuse_kernel<<<1, 1, 0, s>>>(F);
  if (launches) *launches += 2;
}

}  // namespace fccf
