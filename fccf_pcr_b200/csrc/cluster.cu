// cluster.cu — transform_cluster (FCCF.cpp:1040-1231) for the three roughness pools.
//
// The reference seeds clusters greedily in index order (the last hypothesis never seeds,
// FCCF.cpp:1084), gathers with pcl::KdTreeFLANN::radiusSearch (squared L2 < r^2, sorted by
// (distance, index)) plus a 2-degree test on the rotated x axis, sorts clusters by size with an
// exchange sort (range_cluster, 1020-1038) and emits averaged centres with an adaptive size
// cut-off (1123-1229).  Here:
//   cluster_prep    rotated x axis per hypothesis and a (type, x) sort key
//   radix sort      (sort.cu) -> hypotheses ordered by translation x inside each pool, so the
//                   exact radius test only visits an x-window
//   cluster_kernel  one CTA per pool: the greedy seeding is the lexicographically-first
//                   independent set of the neighbour relation, resolved in parallel rounds
//                   (a hypothesis is a seed once all earlier neighbours are known non-seeds);
//                   then cluster sizes, the exact exchange sort (warp_exchange_sort), the
//                   cut-off walk, and one warp per emitted cluster for the (distance,index)
//                   ordered float32 averaging.
#include "fccf_dev.cuh"
#include "fccf_internal.h"
#include <vector>

namespace fccf {

struct ClArgs {
  PipeState* st;
  const float* hyp_qt; float* hyp_ax; double* hyp_an;
  u64* keys; const u32* order;     // order: sorted position -> global hypothesis index
  float* xs;                        // x in sorted order (global index space)
  int *state, *size, *seeds, *perm, *key, *members;
  float* mdist;
  float* centre;
  float thr_n, ang_cut, rad, sel_num;   // ang_cut: cosine cut of cluster_angel_threshold (strict <)
  int* nbits;                       // device word: key width for the sort
  int cap_hyp;
  int* nbl; int* deg;               // 3 x CL_SMEM_N x CL_NB neighbour lists and 3 x CL_SMEM_N neighbour counts (shared-memory pools)
};

__global__ void __launch_bounds__(256) cluster_prep_kernel(const ClArgs* __restrict__ AB) {
  FCCF_PDL_ENTER();
  const ClArgs& A = AB[blockIdx.z];
  PipeState* st = A.st;
  const int n = st->hyp_off[3];
  if (blockIdx.x == 0 && threadIdx.x == 0) *A.nbits = 34;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
  const float* q = A.hyp_qt + (size_t)i * 8;
  q4 Q; Q.w = q[0]; Q.x = q[1]; Q.y = q[2]; Q.z = q[3];
  f3 ax = quat_rotate(Q, mk3(1, 0, 0));
  float* o = A.hyp_ax + (size_t)i * 4;
  o[0] = ax.x; o[1] = ax.y; o[2] = ax.z; o[3] = 0.f;
  A.hyp_an[i] = normal_norm(ax.x, ax.y, ax.z);
  int ty = (i >= st->hyp_off[2]) ? 2 : ((i >= st->hyp_off[1]) ? 1 : 0);
  float x = q[4];
  u32 k;
  if (x != x) k = 0xffffffffu;
  else { u32 b = __float_as_uint(x); k = (b & 0x80000000u) ? ~b : (b | 0x80000000u); }
  A.keys[i] = ((u64)ty << 32) | (u64)k;
  }
}

__global__ void __launch_bounds__(256) cluster_xs_kernel(const ClArgs* __restrict__ AB) {
  FCCF_PDL_ENTER();
  const ClArgs& A = AB[blockIdx.z];
  const int n = A.st->hyp_off[3];
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
    float x = A.hyp_qt[(size_t)A.order[k] * 8 + 4];
    A.xs[k] = (x != x) ? CUDART_INF_F : x;
  }
}

// ti/tj: translations (3 floats), ai/aj: rotated x axes (3 floats), ni/nj: their double norms
__device__ __forceinline__ bool cl_neigh(const float* ti, const float* ai, double ni, const float* tj, const float* aj, double nj, float r2, float ang_cut, float* dist) {
  float d0 = ti[0] - tj[0], d1 = ti[1] - tj[1], d2 = ti[2] - tj[2];
  float d = 0.f; d += d0 * d0; d += d1 * d1; d += d2 * d2;    // flann::L2_Simple
  if (!(d < r2)) return false;
  if (dist) *dist = d;
  return angle_lt(normal_cos_n(ai[0], ai[1], ai[2], ni, aj[0], aj[1], aj[2], nj), ang_cut);   // compute_normal_angel(...) < 2 degrees (FCCF.cpp:1110)
}
__device__ __forceinline__ float cl_dist2(const float* ti, const float* tj) {
  float d0 = ti[0] - tj[0], d1 = ti[1] - tj[1], d2 = ti[2] - tj[2];
  float d = 0.f; d += d0 * d0; d += d1 * d1; d += d2 * d2;    // flann::L2_Simple
  return d;
}
__device__ __forceinline__ bool cl_angle_ok(const float* ai, double ni, const float* aj, double nj, float ang_cut) {
  return angle_lt(normal_cos_n(ai[0], ai[1], ai[2], ni, aj[0], aj[1], aj[2], nj), ang_cut);   // compute_normal_angel(...) < 2 degrees (FCCF.cpp:1110)
}
__device__ __forceinline__ void cl_window(const float* xs, int n, float x, double rr, int& lo, int& hi) {
  if (!isfinite(x)) { lo = 0; hi = 0; return; }   // a non-finite translation has no neighbour (d is NaN/inf)
  double a = (double)x - rr, b = (double)x + rr;
  int l = 0, h = n;
  while (l < h) { int m = (l + h) >> 1; if ((double)xs[m] < a) l = m + 1; else h = m; }
  lo = l; h = n;
  while (l < h) { int m = (l + h) >> 1; if ((double)xs[m] <= b) l = m + 1; else h = m; }
  hi = l;
}

#define CL_WSCR 128   // per-warp member scratch (clusters up to this size are averaged by one warp)

// ordered float32 averaging of one cluster whose members (local indices) are sorted by (dist, index)
__device__ void cl_emit_centre(const float* qt, const int* mem, int m, float* out, int lane) {
  float acc = 0.f;
  if (lane < 9) {
    for (int k = 0; k < m; k++) {
      const float* q = qt + (size_t)mem[k] * 8;
      float v;
      if (lane < 3) v = q[4 + lane];
      else {
        q4 Q; Q.w = q[0]; Q.x = q[1]; Q.y = q[2]; Q.z = q[3];
        f3 r = quat_rotate(Q, lane < 6 ? mk3(1, 0, 0) : mk3(0, 1, 0));
        v = get(r, (lane - 3) % 3);
      }
      acc = acc + v;
    }
    acc = acc / (float)m;
  }
  float s[9];
#pragma unroll
  for (int k = 0; k < 9; k++) s[k] = __shfl_sync(0xffffffffu, acc, k);
  if (lane == 0) {
    f3 a1 = mk3(s[3], s[4], s[5]), a2 = mk3(s[6], s[7], s[8]);
    normalize(a1); normalize(a2);
    m3 R = rotation_from_axes(a1, a2);
    q4 q = quat_from_matrix(R);
    out[0] = q.w; out[1] = q.x; out[2] = q.y; out[3] = q.z; out[4] = s[0]; out[5] = s[1]; out[6] = s[2]; out[7] = 0.f;
  }
}

#define CL_NB 24               // neighbours listed per hypothesis
#define CL_SMEM_N 2816        // pools up to this many hypotheses are clustered out of shared memory
#define CL_SMEM_BYTES (CL_SMEM_N * 64)
extern __shared__ __align__(16) unsigned char cl_dyn[];

__global__ void __launch_bounds__(1024) cluster_kernel(const ClArgs* __restrict__ AB) {
  FCCF_PDL_ENTER();
  const ClArgs& A = AB[blockIdx.z];
  PipeState* st = A.st;
  const int ty = blockIdx.x;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int n = st->n_hyp[ty], base = st->hyp_off[ty], tnum = st->hyp_off[3];
  const float* qt = A.hyp_qt + (size_t)base * 8;
  float* centre = A.centre + (size_t)ty * FCCF_MAXCENTRE * 8;
  __shared__ int s_flag, s_K, s_E, s_special;
  __shared__ unsigned long long s_sort[40];
  __shared__ int s_emit[FCCF_MAXCENTRE];
  __shared__ int s_mem[32][CL_WSCR];
  __shared__ float s_md[32][CL_WSCR];
  // cluster_num (FCCF.cpp:1465): int(200.0f * size / total); NaN (0/0) casts to INT_MIN on x86
  float cnf = A.sel_num * (float)n / (float)tnum;
  int cluster_num = (cnf == cnf) ? (int)cnf : (int)0x80000000;
  if (t == 0) st->cluster_num[ty] = cluster_num;
  if ((float)n <= A.thr_n) {   // FCCF.cpp:1043-1063
    if (n == 0) {
      if (t == 0) { centre[0] = 1.f; for (int k = 1; k < 8; k++) centre[k] = 0.f; st->n_centre[ty] = 1; st->n_seeds[ty] = 0; }
    } else {
      // cluster_number_threshold <= FCCF_MAXCENTRE is checked on the host; clamp anyway and report it
      const int m = n < FCCF_MAXCENTRE ? n : FCCF_MAXCENTRE;
      for (int k = t; k < m * 8; k += 1024) centre[k] = qt[k];
      if (t == 0) { st->n_centre[ty] = m; st->n_seeds[ty] = 0; if (m < n) atomicOr(&st->status, ST_CENTRE_OVERFLOW); }
    }
    return;
  }
  // Working set of the pool: translations, rotated x axes and their norms, the x-sorted order, the seed
  // state and the sort arrays.  Pools of up to CL_SMEM_N hypotheses keep all of it in shared memory (the
  // greedy rounds poll `state` of every earlier neighbour); larger pools work on the global arrays.
  const bool sm = n <= CL_SMEM_N;
  const double* an; const float* tr; const float* ax; const float* xs; const int* order; int ts, as, obase;
  int* state; int* key; int* perm; float* d_trs = nullptr;
  if (sm) {
    double* d_an = (double*)cl_dyn;
    float* d_tr = (float*)(d_an + CL_SMEM_N); float* d_ax = d_tr + 3 * CL_SMEM_N; float* d_xs = d_ax + 3 * CL_SMEM_N;
    int* d_ord = (int*)(d_xs + CL_SMEM_N); state = d_ord + CL_SMEM_N; key = state + CL_SMEM_N; perm = key + CL_SMEM_N;
    d_trs = (float*)(perm + CL_SMEM_N);     // translations once more, in x-sorted order: a window is contiguous (conflict-free sweeps)
    for (int i = t; i < n; i += 1024) {
      const float* q = qt + (size_t)i * 8; const float* a = A.hyp_ax + (size_t)(base + i) * 4;
      d_tr[3 * i] = q[4]; d_tr[3 * i + 1] = q[5]; d_tr[3 * i + 2] = q[6];
      d_ax[3 * i] = a[0]; d_ax[3 * i + 1] = a[1]; d_ax[3 * i + 2] = a[2];
      d_an[i] = A.hyp_an[base + i];
      d_xs[i] = A.xs[base + i];
      d_ord[i] = (int)A.order[base + i] - base;
      const float* qs = qt + (size_t)d_ord[i] * 8;
      d_trs[3 * i] = qs[4]; d_trs[3 * i + 1] = qs[5]; d_trs[3 * i + 2] = qs[6];
    }
    an = d_an; tr = d_tr; ax = d_ax; xs = d_xs; order = d_ord; ts = 3; as = 3; obase = 0;
  } else {
    an = A.hyp_an + base; tr = qt + 4; ax = A.hyp_ax + (size_t)base * 4; xs = A.xs + base; order = (const int*)(A.order + base); ts = 8; as = 4; obase = base;
    state = A.state + base; key = A.key + base; perm = A.perm + base;
  }
  int* size = A.size + base; int* seeds = A.seeds + base;
  const double rad = (double)A.rad;
  const float r2 = (float)(rad * rad);
  const double rr = rad * 1.0001 + 1e-6;
  for (int i = t; i < n; i += 1024) state[i] = (i == n - 1) ? 2 : 0;
  __syncthreads();
#define CL_MARK(k) if (ty == 0 && t == 0) st->prof[k] = clock64();
#define CL_NEIGH(i, j, dp) cl_neigh(tr + (size_t)(i) * ts, ax + (size_t)(i) * as, an[i], tr + (size_t)(j) * ts, ax + (size_t)(j) * as, an[j], r2, A.ang_cut, dp)
  CL_MARK(0)
  int rounds = 0;
  // ---- neighbour lists (shared-memory pools) ----
  // ALL neighbours of every hypothesis, found once.  One warp per hypothesis: its lanes sweep the x-window
  // with the cheap float radius test, ballots append the candidates (i, j) to a per-warp queue, and the
  // expensive FP64 angle test runs on full batches of 32 queued candidates (of one or several hypotheses),
  // so neither the window sweep nor the angle test wastes lanes.  Lists of up to CL_NB neighbours are kept
  // (nbl), the count always (deg; bit 30: no list); the seeding rounds, the cluster sizes and the centre
  // averaging below read them instead of re-scanning windows.
  int* nbl = A.nbl + (size_t)ty * CL_SMEM_N * CL_NB; int* deg = A.deg + (size_t)ty * CL_SMEM_N;
  if (sm) {
    for (int i = t; i < n; i += 1024) deg[i] = 0;
    __syncthreads();
    int* qi = s_mem[warp]; int* qj = (int*)s_md[warp];       // queue of candidates inside the radius (CL_WSCR = 128 entries)
    int qn = 0;
    const unsigned lt = (1u << lane) - 1u;
    // angle test of queue entries [0, nf): every passing lane appends j to the list of its i
    auto flush = [&](int nf) {
      for (int b0 = 0; b0 < nf; b0 += 32) {
        const int e = b0 + lane;
        bool ok = false; int i = -1, j = -1;
        if (e < nf) { i = qi[e]; j = qj[e]; ok = cl_angle_ok(ax + (size_t)i * 3, an[i], ax + (size_t)j * 3, an[j], A.ang_cut); }
        const unsigned pm = __ballot_sync(0xffffffffu, ok);
        if (ok) {
          const unsigned peers = __match_any_sync(pm, i);       // passing lanes of the same hypothesis
          const int leader = __ffs(peers) - 1;
          int old = 0;
          if (lane == leader) { old = deg[i]; deg[i] = old + __popc(peers); }
          old = __shfl_sync(peers, old, leader);
          const int p = old + __popc(peers & lt);
          if (p < CL_NB) nbl[(size_t)i * CL_NB + p] = j;
        }
        __syncwarp();
      }
    };
    for (int g0 = warp; g0 < n; g0 += 32 * 32) {                // 32 hypotheses of this warp at a time
      const int il = g0 + 32 * lane;                            // lane l looks up the window of hypothesis il
      int lo_l = 0, hi_l = 0;
      if (il < n) cl_window(xs, n, tr[(size_t)il * 3], rr, lo_l, hi_l);
      for (int l2 = 0; l2 < 32; l2++) {
        const int i = g0 + 32 * l2;
        if (i >= n) break;
        const int lo = __shfl_sync(0xffffffffu, lo_l, l2), hi = __shfl_sync(0xffffffffu, hi_l, l2);
        const float ti[3] = {tr[(size_t)i * 3], tr[(size_t)i * 3 + 1], tr[(size_t)i * 3 + 2]};
        for (int k0 = lo; k0 < hi; k0 += 32) {
          const int k = k0 + lane;
          int j = -1;
          bool cand = false;
          if (k < hi) cand = cl_dist2(ti, d_trs + (size_t)k * 3) < r2;
          const unsigned bal = __ballot_sync(0xffffffffu, cand);
          if (cand) { j = order[k]; const int p = qn + __popc(bal & lt); qi[p] = i; qj[p] = j; }
          qn += __popc(bal);
          if (qn >= CL_WSCR - 32) {                             // the next sweep step may add 32 more
            __syncwarp();
            const int nf = qn & ~31, rem = qn - nf;
            flush(nf);
            int ri = 0, rj = 0;
            if (lane < rem) { ri = qi[nf + lane]; rj = qj[nf + lane]; }
            __syncwarp();
            if (lane < rem) { qi[lane] = ri; qj[lane] = rj; }
            qn = rem;
            __syncwarp();
          }
        }
      }
    }
    __syncwarp();
    flush(qn);
    __syncthreads();
    for (int i = t; i < n; i += 1024) if (deg[i] > CL_NB) deg[i] |= 0x40000000;
    __syncthreads();
  }
  CL_MARK(9)
  // ---- greedy seeding as parallel rounds ----
  // A hypothesis is a seed iff none of its EARLIER neighbours is one (FCCF.cpp:1084-1121 in index order).
  while (true) {
    if (t == 0) s_flag = 0;
    __syncthreads();
    if (sm) {
#pragma unroll 1
      for (int u = 0; u < 3; u++) {
        const int i = t + u * 1024;
        if (i >= n || ((volatile int*)state)[i] != 0) continue;
        bool found_seed = false, all_dec = true;
        const int dg = deg[i];
        if (!(dg & 0x40000000)) {
          for (int c = 0; c < dg; c++) {
            int j = nbl[(size_t)i * CL_NB + c];
            if (j >= i) continue;
            int sj = ((volatile int*)state)[j];
            if (sj == 1) { found_seed = true; break; }
            if (sj == 0) all_dec = false;
          }
        } else {
          int lo, hi; cl_window(xs, n, tr[(size_t)i * ts], rr, lo, hi);
          for (int k = lo; k < hi; k++) {
            int j = order[k] - obase;
            if (j >= i) continue;
            int sj = ((volatile int*)state)[j];
            if (sj == 2) continue;
            if (CL_NEIGH(i, j, nullptr)) {
              if (sj == 1) { found_seed = true; break; }
              all_dec = false;
            }
          }
        }
        if (found_seed) state[i] = 2;
        else if (all_dec) state[i] = 1;
        else s_flag = 1;
      }
    } else {
      for (int i = t; i < n; i += 1024) {
        if (((volatile int*)state)[i] != 0) continue;
        int lo, hi; cl_window(xs, n, tr[(size_t)i * ts], rr, lo, hi);
        bool found_seed = false, all_dec = true;
        for (int k = lo; k < hi; k++) {
          int j = order[k] - obase;
          if (j >= i) continue;
          int sj = ((volatile int*)state)[j];
          if (sj == 2) continue;
          if (CL_NEIGH(i, j, nullptr)) {
            if (sj == 1) { found_seed = true; break; }
            all_dec = false;
          }
        }
        if (found_seed) state[i] = 2;
        else if (all_dec) state[i] = 1;
        else s_flag = 1;
      }
    }
    __syncthreads();
    rounds++;
    if (!s_flag) break;
    __syncthreads();
  }
  CL_MARK(1)
  if (ty == 0 && t == 0) st->prof[8] = rounds;
  // ---- seeds in index order + cluster sizes ----
  if (t == 0) s_K = 0;
  __syncthreads();
  {
    __shared__ int s_w[32];
    for (int i0 = 0; i0 < n; i0 += 1024) {
      int i = i0 + t;
      bool is = (i < n) && state[i] == 1;
      unsigned b = __ballot_sync(0xffffffffu, is);
      if (lane == 0) s_w[warp] = __popc(b);
      __syncthreads();
      int off = s_K;
      for (int w2 = 0; w2 < warp; w2++) off += s_w[w2];
      if (is) seeds[off + __popc(b & ((1u << lane) - 1u))] = i;
      __syncthreads();
      if (t == 0) { int tot = 0; for (int w2 = 0; w2 < 32; w2++) tot += s_w[w2]; s_K += tot; }
      __syncthreads();
    }
  }
  const int K = s_K;
  if (t == 0) st->n_seeds[ty] = K;
  // cluster sizes: the neighbour count of the seed (shared-memory pools), else one warp per seed with
  // the window split over its lanes
  if (sm) {
    for (int k = t; k < K; k += 1024) { int cnt = deg[seeds[k]] & 0x3fffffff; size[k] = cnt; key[k] = cnt; perm[k] = k; }
  } else
  for (int k = warp; k < K; k += 32) {
    int i = seeds[k];
    int lo, hi; cl_window(xs, n, tr[(size_t)i * ts], rr, lo, hi);
    int cnt = 0;
    for (int kk = lo + lane; kk < hi; kk += 32) {
      int j = order[kk] - obase;
      if (CL_NEIGH(i, j, nullptr)) cnt++;
    }
    for (int o = 16; o; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if (lane == 0) { size[k] = cnt; key[k] = cnt; perm[k] = k; }
  }
  __syncthreads();
  CL_MARK(2)
  // ---- range_cluster: exchange sort by size ----
  block_exchange_sort(key, perm, K, s_sort);
  __syncthreads();
  if (sm) { for (int k = t; k < K; k += 1024) { A.key[base + k] = key[k]; A.perm[base + k] = perm[k]; } }   // debug blobs read the global copies
  CL_MARK(3)
  // ---- adaptive cut-off walk (FCCF.cpp:1123-1229) ----
  if (t == 0) {
    int clusternum = key[0];
    int emitted = 0, special = 0;
    const double half = cluster_num / 2.0;
    int kn = key[0], pn = perm[0];                       // the next entry is fetched while this one is decided
    for (int ci = 0; ci < K; ci++) {
      const int kc = kn, pc = pn;
      if (ci + 1 < K) { kn = key[ci + 1]; pn = perm[ci + 1]; }
      if (kc >= clusternum) {
        if (emitted >= FCCF_MAXCENTRE) { atomicOr(&st->status, ST_CENTRE_OVERFLOW); break; }
        s_emit[emitted++] = pc;
        special |= (kc == 0 || kc > CL_WSCR);     // clusters the warp path below does not finish (empty / above CL_WSCR members)
        if (cluster_num >= 0 && emitted > cluster_num) break;
      } else {
        if ((double)emitted < half) { clusternum--; if (clusternum < 2) break; }
        else break;   // stop = true: nothing further is emitted
      }
    }
    s_E = emitted;
    st->n_centre[ty] = emitted;
    s_special = special;
  }
  __syncthreads();
  const int E = s_E;
  CL_MARK(4)
  // ---- centres: small clusters by one warp each ----
  for (int e = warp; e < E; e += 32) {
    int k = s_emit[e]; int i = seeds[k]; int m = size[k];
    if (m > CL_WSCR || m == 0) continue;
    int lo = 0, hi = 0;
    int cntw = 0;
    if (sm && !(deg[i] & 0x40000000)) {     // members from the neighbour list (m <= CL_NB <= 32), distances recomputed
      if (lane < m) { int j = nbl[(size_t)i * CL_NB + lane]; s_mem[warp][lane] = j; s_md[warp][lane] = cl_dist2(tr + (size_t)i * 3, tr + (size_t)j * 3); }
    } else cl_window(xs, n, tr[(size_t)i * ts], rr, lo, hi);
    for (int k0 = lo; k0 < hi; k0 += 32) {
      int kk = k0 + lane; bool ok = false; float d = 0.f; int j = -1;
      if (kk < hi) { j = order[kk] - obase; ok = CL_NEIGH(i, j, &d); }
      unsigned b = __ballot_sync(0xffffffffu, ok);
      if (ok) { int p = cntw + __popc(b & ((1u << lane) - 1u)); s_mem[warp][p] = j; s_md[warp][p] = d; }
      cntw += __popc(b);
    }
    __syncwarp();
    // rank sort by (dist, index)
    int myj[CL_WSCR / 32]; int myr[CL_WSCR / 32];
#pragma unroll
    for (int u = 0; u < CL_WSCR / 32; u++) {
      int p = u * 32 + lane; myj[u] = -1; myr[u] = 0;
      if (p < m) {
        float d = s_md[warp][p]; int j = s_mem[warp][p]; int r = 0;
        for (int o = 0; o < m; o++) { float d2 = s_md[warp][o]; int j2 = s_mem[warp][o]; if (d2 < d || (d2 == d && j2 < j)) r++; }
        myj[u] = j; myr[u] = r;
      }
    }
    __syncwarp();
#pragma unroll
    for (int u = 0; u < CL_WSCR / 32; u++) if (myj[u] >= 0) s_mem[warp][myr[u]] = myj[u];
    __syncwarp();
    cl_emit_centre(qt, s_mem[warp], m, centre + (size_t)e * 8, lane);
    __syncwarp();
  }
  __syncthreads();
  CL_MARK(5)
  // ---- big clusters: whole block gathers into global scratch, rank sort, warp 0 averages ----
  int* gmem = A.members + base; float* gmd = A.mdist + base;
  int* gsorted = A.members + A.cap_hyp + base;   // second half of the member scratch
  __shared__ int s_cnt, s_w2[32];
  for (int e = 0; e < (s_special ? E : 0); e++) {
    int k = s_emit[e]; int i = seeds[k]; int m = size[k];
    if (m == 0) { if (t == 0) { float* o = centre + (size_t)e * 8; float nanv = CUDART_NAN_F; for (int u = 0; u < 8; u++) o[u] = nanv; } continue; }
    if (m <= CL_WSCR) continue;
    int lo, hi; cl_window(xs, n, tr[(size_t)i * ts], rr, lo, hi);
    if (t == 0) s_cnt = 0;
    __syncthreads();
    for (int k0 = lo; k0 < hi; k0 += 1024) {
      int kk = k0 + t; bool ok = false; float d = 0.f; int j = -1;
      if (kk < hi) { j = order[kk] - obase; ok = CL_NEIGH(i, j, &d); }
      unsigned b = __ballot_sync(0xffffffffu, ok);
      if (lane == 0) s_w2[warp] = __popc(b);
      __syncthreads();
      int off = s_cnt;
      for (int w2 = 0; w2 < warp; w2++) off += s_w2[w2];
      if (ok) { int p = off + __popc(b & ((1u << lane) - 1u)); gmem[p] = j; gmd[p] = d; }
      __syncthreads();
      if (t == 0) { int tot = 0; for (int w2 = 0; w2 < 32; w2++) tot += s_w2[w2]; s_cnt += tot; }
      __syncthreads();
    }
    for (int p = t; p < m; p += 1024) {
      float d = gmd[p]; int j = gmem[p]; int r = 0;
      for (int o = 0; o < m; o++) { float d2 = gmd[o]; int j2 = gmem[o]; if (d2 < d || (d2 == d && j2 < j)) r++; }
      gsorted[r] = j;
    }
    __syncthreads();
    if (warp == 0) cl_emit_centre(qt, gsorted, m, centre + (size_t)e * 8, lane);
    __syncthreads();
  }
  CL_MARK(6)
}

int cluster_nbl_ints() { return 3 * CL_SMEM_N * CL_NB; }
int cluster_deg_ints() { return 3 * CL_SMEM_N; }
void cluster_init_attributes() { cudaFuncSetAttribute(cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, CL_SMEM_BYTES); }

void launch_cluster(cudaStream_t s, const Batch& b, uint64_t* launches) {
  const int G = b.G;
  std::vector<ClArgs> As(G); std::vector<SortJobs> abs_(G), bas_(G);
  int cap = 1;
  for (int g = 0; g < G; g++) {
    const Work& w = b.w[g]; const HypWS& h = w.h;
    ClArgs& A = As[g]; SortJobs& ab = abs_[g]; SortJobs& ba = bas_[g];
    memset(&A, 0, sizeof A); memset(&ab, 0, sizeof ab); memset(&ba, 0, sizeof ba);
    PipeState* st = w.st;
    A.st = st; A.hyp_qt = h.hyp_qt; A.hyp_ax = h.hyp_ax; A.hyp_an = h.hyp_an; A.keys = h.ckeyA; A.order = h.cidxA; A.xs = (float*)h.c_mdist + h.cap_hyp;
    A.state = h.c_state; A.size = h.c_size; A.seeds = h.c_seeds; A.perm = h.c_perm; A.key = h.c_key; A.members = h.c_members; A.mdist = h.c_mdist;
    A.centre = h.centre;
    A.thr_n = b.p.cluster_number_threshold; A.ang_cut = b.cuts.cluster_lt; A.rad = b.p.cluster_distance_threshold; A.sel_num = b.p.seclct_cluster_number;
    A.nbits = &st->tickets[20]; A.cap_hyp = h.cap_hyp; A.nbl = h.c_nbl; A.deg = h.c_deg;
    if (h.cap_hyp > cap) cap = h.cap_hyp;
    SortJob j; j.kin = h.ckeyA; j.kout = h.ckeyB; j.vin = h.cidxA; j.vout = h.cidxB; j.n = &st->hyp_off[3]; j.nbits = &st->tickets[20]; j.hist = h.chist; j.ticket = &st->tickets[21]; j.miss = &st->status;
    ab.j[0] = j; ab.j[1] = j; ab.j[2] = j;
    SortJob k = j; k.kin = h.ckeyB; k.kout = h.ckeyA; k.vin = h.cidxB; k.vout = h.cidxA; ba.j[0] = k; ba.j[1] = k; ba.j[2] = k;
  }
  const ClArgs* dA = b.tab->put(As.data(), G);
  const SortJobs* dab = b.tab->put(abs_.data(), G); const SortJobs* dba = b.tab->put(bas_.data(), G);
  klaunch(cluster_prep_kernel, dim3(dim3(grid_x((cap + 255) / 256, G), 1, G)), dim3(256), 0, s, dA);
  if (launches) *launches += 1;
  // 34-bit keys: 6 passes of 6 bits (even pass count: result back in ckeyA / cidxA)
  launch_sort(s, dab, dba, 1, G, cap, 6, 8, launches, b.lean);
  klaunch(cluster_xs_kernel, dim3(dim3(grid_x((cap + 255) / 256, G), 1, G)), dim3(256), 0, s, dA);
  klaunch(cluster_kernel, dim3(dim3(3, 1, G)), dim3(1024), CL_SMEM_BYTES, s, dA);
  if (launches) *launches += 2;
}

}  // namespace fccf
