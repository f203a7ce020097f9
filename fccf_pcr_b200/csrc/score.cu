// score.cu — fine_verify (FCCF.cpp:785-839) as "hypothesis scoring", the block-reduced argmax over a
// score list, per-type best + gate (FCCF.cpp:1546-1605) and fuse_answer (1291-1368).
//
// The reference rebuilds a pcl octree (0.5 m) over static + transformed moving leftover points for
// every hypothesis and counts both kinds per voxel.  The voxel lattice is anchored on the first
// static point (pcl's first insert), so it does not depend on the hypothesis.  Here, once per pair:
//   score_bbox    lattice coordinates l = floor((double(p) - mn) / res) of the static points, their
//                 bounding box, and from it a compact 32-bit voxel key ((cx*dy + cy)*dz + cz) plus
//                 exact float bounds of the box (a float q is inside iff lo_f <= q < hi_f)
//   sort+segments (sort.cu) occupied static voxels in ascending key order: voxel id, static count
//   score_insert  open-addressing hash  compact key -> voxel id  (linear probing, load <= 0.5)
// and per hypothesis (score_kernel): per moving point a float32 3x4 transform in pcl's SSE order, six
// float compares against the box, a float64 lattice key, one probe and one counter increment; then
// sum over the occupied voxels in id order of (s+t)*min/max.  A CTA scores SC_NH hypotheses per
// sweep over the moving cloud with the hash and the counters in shared memory (global-memory
// fallback for tables that do not fit).  Voxel ids are the sort order, so per-voxel integer counts
// (s,t) are exactly the reference's and the float32 score does not depend on the hash layout.
#include "fccf_dev.cuh"
#include "fccf_internal.h"
#include <algorithm>
#include <vector>

namespace fccf {

#define SC_EMPTY32 0xffffffffu
#define SC_LIM (1 << 20)         // |lattice coordinate| bound of the static cloud
#define SC_THREADS 1024
#define SC_NH 4                  // global-table kernel: hypotheses per sweep over the moving cloud
#define SC_WARPS 32              // shared-memory kernel: one hypothesis per warp
#define SC_DYN_BYTES 204800      // dynamic shared memory of the shared-memory kernel (200 KB)
#define SC_DENSE_CELLS 32768     // bounding boxes up to this many cells use a dense cell -> voxel id table

struct ScArgs {
  ScoreWS ws;
  const float* s1; const int* n1p; const int* n2p;
  const float* s2;
  const float* T; int n_hyp; const int* n_top; int topk;   // n_top != nullptr: pipeline layout [3][topk]
  float* scores;
  int* rows; int cap_rows; int* nrows;
  float res;
};

__device__ __forceinline__ unsigned sc_hash32(u32 k) { return k * 0x9E3779B1u; }

// lattice origin: first insert into an empty pcl octree: box = p0 +- res/2, padded to depth 1
__global__ void score_setup_kernel(const ScArgs* __restrict__ AB) {
  FCCF_PDL_ENTER();
  const ScArgs& A = AB[blockIdx.z];
  ScoreState* ss = A.ws.ss;
  if (threadIdx.x != 0) return;
  int n1 = *A.n1p, n2 = *A.n2p;
  if (n1 > A.ws.cap_points) { n1 = 0; atomicOr(A.ws.status, ST_HASH_FULL); }
  ss->n1 = n1; ss->n2 = n2; ss->n_occ = 0; ss->n_keys = n1; ss->nbits = 1; ss->cap_eff = 64; ss->mode = 2;
  double res = (double)A.res;
  if (n1 > 0) {
    const float minValue = 1.1920928955078125e-07f;
    for (int a = 0; a < 3; a++) {
      double mn = (double)A.s1[a] - res / 2, mx = (double)A.s1[a] + res / 2;
      double side = 2.0 * res;
      double over = (side - (mx - mn)) / 2.0;
      if (over > minValue) mn -= over;
      ss->mn[a] = mn;
    }
  } else { ss->mn[0] = ss->mn[1] = ss->mn[2] = 0.0; }
  for (int a = 0; a < 3; a++) { ss->lmin[a] = 0x7fffffff; ss->lmax[a] = (int)0x80000000; ss->dims[a] = 1; ss->lo_f[a] = 0.f; ss->hi_f[a] = 0.f; }
  for (int a = 0; a < 4; a++) ss->tickets[a] = 0;
}

__device__ __forceinline__ bool sc_lattice(const ScoreState* ss, double res, float x, float y, float z, int l[3]) {
  double v[3] = {floor(((double)x - ss->mn[0]) / res), floor(((double)y - ss->mn[1]) / res), floor(((double)z - ss->mn[2]) / res)};
  bool ok = true;
#pragma unroll
  for (int a = 0; a < 3; a++) { ok = ok && (v[a] > -(double)SC_LIM) && (v[a] < (double)SC_LIM); l[a] = ok ? (int)v[a] : 0; }
  return ok;
}

// bounding box of the static lattice coordinates; the last block derives the compact key layout
__global__ void __launch_bounds__(256) score_bbox_kernel(const ScArgs* __restrict__ AB) {
  FCCF_PDL_ENTER();
  const ScArgs& A = AB[blockIdx.z];
  ScoreState* ss = A.ws.ss;
  const int n1 = ss->n1;
  const double res = (double)A.res;
  int mn[3] = {0x7fffffff, 0x7fffffff, 0x7fffffff}, mx[3] = {(int)0x80000000, (int)0x80000000, (int)0x80000000};
  bool bad = false;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n1; i += gridDim.x * blockDim.x) {
    int l[3];
    if (!sc_lattice(ss, res, A.s1[3 * i], A.s1[3 * i + 1], A.s1[3 * i + 2], l)) { bad = true; continue; }
#pragma unroll
    for (int a = 0; a < 3; a++) { mn[a] = min(mn[a], l[a]); mx[a] = max(mx[a], l[a]); }
  }
  for (int o = 16; o; o >>= 1) {
#pragma unroll
    for (int a = 0; a < 3; a++) { mn[a] = min(mn[a], __shfl_xor_sync(0xffffffffu, mn[a], o)); mx[a] = max(mx[a], __shfl_xor_sync(0xffffffffu, mx[a], o)); }
  }
  if ((threadIdx.x & 31) == 0 && mn[0] != 0x7fffffff) {
#pragma unroll
    for (int a = 0; a < 3; a++) { atomicMin(&ss->lmin[a], mn[a]); atomicMax(&ss->lmax[a], mx[a]); }
  }
  if (bad) atomicOr(A.ws.status, ST_HASH_FULL);
  __threadfence();
  __syncthreads();
  __shared__ int s_last;
  if (threadIdx.x == 0) s_last = (atomicAdd(&ss->tickets[0], 1) == (int)gridDim.x - 1);
  __syncthreads();
  if (!s_last || threadIdx.x != 0) return;
  __threadfence();
  ss->tickets[0] = 0;
  if (n1 <= 0) { ss->n_keys = 0; return; }
  int lo[3], hi[3];
  for (int a = 0; a < 3; a++) { lo[a] = atomicAdd(&ss->lmin[a], 0); hi[a] = atomicAdd(&ss->lmax[a], 0); }
  if (lo[0] == 0x7fffffff || (atomicAdd(A.ws.status, 0) & ST_HASH_FULL)) { ss->n_keys = 0; ss->n1 = 0; return; }   // static cloud outside the supported lattice range
  unsigned long long tot = 1;
  for (int a = 0; a < 3; a++) { ss->dims[a] = hi[a] - lo[a] + 1; tot *= (unsigned long long)ss->dims[a]; }
  if (tot >= 0xffffffffull) { atomicOr(A.ws.status, ST_HASH_FULL); ss->n_keys = 0; ss->n1 = 0; return; }
  int nbits = 64 - __clzll(tot);
  ss->nbits = nbits < 1 ? 1 : nbits;
  // a float q lies in lattice cells [lo, hi] iff  Lo <= double(q) < Hi  with  Lo = mn + lo*res,
  // Hi = mn + (hi+1)*res (both exact: res is a small binary fraction in practice; otherwise the
  // per-point key below still decides) iff  lo_f <= q < hi_f  with both rounded UP to float
  for (int a = 0; a < 3; a++) {
    double Lo = ss->mn[a] + (double)lo[a] * res, Hi = ss->mn[a] + (double)(hi[a] + 1) * res;
    ss->lo_f[a] = __double2float_ru(Lo);
    ss->hi_f[a] = __double2float_ru(Hi);
  }
}

__device__ __forceinline__ u32 sc_compact(const ScoreState* ss, const int l[3]) {
  return (u32)(((l[0] - ss->lmin[0]) * ss->dims[1] + (l[1] - ss->lmin[1])) * ss->dims[2] + (l[2] - ss->lmin[2]));
}

__global__ void __launch_bounds__(256) score_keys_kernel(const ScArgs* __restrict__ AB) {
  FCCF_PDL_ENTER();
  const ScArgs& A = AB[blockIdx.z];
  const ScoreState* ss = A.ws.ss;
  const int n = ss->n_keys;
  const double res = (double)A.res;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    int l[3];
    sc_lattice(ss, res, A.s1[3 * i], A.s1[3 * i + 1], A.s1[3 * i + 2], l);
    ((u32*)A.ws.keyA)[i] = sc_compact(ss, l);
  }
}

// Table plan from the number of occupied voxels (device-side, every thread computes the same):
//   cap   hash slots (power of two >= 2 * n_occ)
//   mode  0: dense cell -> id table + per-warp 16-bit counters in shared memory
//         1: hash + per-warp 16-bit counters in shared memory
//         2: hash and 32-bit counters in global memory (score_kernel)
struct ScPlan { int cap, mode, row_words, nocc_pad, tab_bytes; unsigned cells; int fast, s2_off; };   // fast: 32-bit counters + the moving cloud staged as float4 at s2_off
__device__ __forceinline__ ScPlan sc_plan(const ScoreState* ss, int cap_hash) {
  ScPlan P;
  const int nocc = ss->n_occ;
  int cap = 64;
  while (cap < 2 * nocc && cap < cap_hash) cap <<= 1;
  P.cap = cap;
  P.cells = (unsigned)ss->dims[0] * (unsigned)ss->dims[1] * (unsigned)ss->dims[2];
  P.nocc_pad = (nocc + 63) & ~63;
  P.row_words = P.nocc_pad / 2;
  size_t fixed = (size_t)4 * P.nocc_pad + (size_t)SC_WARPS * 4 * P.row_words;
  bool small = nocc < 65535 && ss->n2 < 65536;      // 16-bit ids and counters
  P.mode = 2; P.tab_bytes = 0;
  if (small && P.cells <= SC_DENSE_CELLS && fixed + ((2 * (size_t)P.cells + 15) & ~(size_t)15) <= SC_DYN_BYTES) { P.mode = 0; P.tab_bytes = (int)((2 * P.cells + 15) & ~15u); }
  else if (small && fixed + (size_t)6 * cap <= SC_DYN_BYTES) { P.mode = 1; P.tab_bytes = 6 * cap; }
  // room for the fast sweep?  one 32-bit counter per voxel (no half-word select / shift in front of the atomic)
  // and the moving cloud as float4 in shared memory (one LDS.128 per point instead of three strided loads)
  P.fast = 0; P.s2_off = 0;
  if (P.mode != 2) {
    const size_t wide = (size_t)P.tab_bytes + (size_t)4 * P.nocc_pad + (size_t)SC_WARPS * 4 * P.nocc_pad;
    if (wide + (size_t)16 * ss->n2 <= SC_DYN_BYTES) { P.fast = 1; P.s2_off = (int)wide; }
  }
  return P;
}

// per occupied voxel (ascending key): key and static count; clears the hash and the dense table
__global__ void __launch_bounds__(256) score_table_kernel(const ScArgs* __restrict__ AB) {
  FCCF_PDL_ENTER();
  const ScArgs& A = AB[blockIdx.z];
  ScoreState* ss = A.ws.ss;
  const int nocc = ss->n_occ;
  const ScPlan P = sc_plan(ss, A.ws.cap_hash);
  if (blockIdx.x == 0 && threadIdx.x == 0) { ss->cap_eff = P.cap; ss->mode = P.mode; }
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < P.cap; i += gridDim.x * blockDim.x) A.ws.hkey[i] = SC_EMPTY32;
  if (P.mode == 0) for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < (P.cells + 1) / 2; i += gridDim.x * blockDim.x) ((u32*)A.ws.dense)[i] = 0xffffffffu;
  for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < nocc; v += gridDim.x * blockDim.x) {
    int b = A.ws.seg_start[v], e = A.ws.seg_start[v + 1];
    A.ws.vkey[v] = ((const u32*)A.ws.keyA)[b];
    A.ws.s_cnt[v] = e - b;
  }
}
__global__ void __launch_bounds__(256) score_insert_kernel(const ScArgs* __restrict__ AB) {
  FCCF_PDL_ENTER();
  const ScArgs& A = AB[blockIdx.z];
  const ScoreState* ss = A.ws.ss;
  const int nocc = ss->n_occ, cap = ss->cap_eff, mode = ss->mode;
  const unsigned mask = (unsigned)cap - 1u;
  int shift = 32 - (31 - __clz(cap));
  for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < nocc; v += gridDim.x * blockDim.x) {
    u32 key = A.ws.vkey[v];
    if (mode == 0) A.ws.dense[key] = (unsigned short)v;       // compact key == cell index
    unsigned h = (sc_hash32(key) >> shift) & mask;
    for (int probe = 0; probe < cap; probe++) {
      u32 prev = atomicCAS(&A.ws.hkey[h], SC_EMPTY32, key);
      if (prev == SC_EMPTY32) { A.ws.hid[h] = (u32)v; break; }
      h = (h + 1) & mask;
    }
  }
}

// ---- scoring ---------------------------------------------------------------------------------
// voxel id of the transformed point, or -1
template <bool POW2>
__device__ __forceinline__ int sc_lookup(const float* __restrict__ T, float px, float py, float pz, const float* lo_f, const float* hi_f, const double* mnd,
                                         double res_or_inv, const int* lmin, int dx, int dy, int dz, const u32* __restrict__ hkey, const u32* __restrict__ hid,
                                         unsigned mask, int shift) {
  // pcl::transformPointCloud (SSE order): x*c0 + (y*c1 + (z*c2 + c3))
  float qx = px * T[0] + (py * T[1] + (pz * T[2] + T[3]));
  float qy = px * T[4] + (py * T[5] + (pz * T[6] + T[7]));
  float qz = px * T[8] + (py * T[9] + (pz * T[10] + T[11]));
  if (!(qx >= lo_f[0] && qx < hi_f[0] && qy >= lo_f[1] && qy < hi_f[1] && qz >= lo_f[2] && qz < hi_f[2])) return -1;
  double vx, vy, vz;
  if (POW2) { vx = ((double)qx - mnd[0]) * res_or_inv; vy = ((double)qy - mnd[1]) * res_or_inv; vz = ((double)qz - mnd[2]) * res_or_inv; }   // exact: 1/res is a power of two
  else { vx = ((double)qx - mnd[0]) / res_or_inv; vy = ((double)qy - mnd[1]) / res_or_inv; vz = ((double)qz - mnd[2]) / res_or_inv; }
  // floor of a small non-huge double: round-down add of 1.5 * 2^52 leaves the integer in the low word
  int lx = __double2loint(__dadd_rd(vx, 6755399441055744.0)), ly = __double2loint(__dadd_rd(vy, 6755399441055744.0)), lz = __double2loint(__dadd_rd(vz, 6755399441055744.0));
  u32 cx = (u32)(lx - lmin[0]), cy = (u32)(ly - lmin[1]), cz = (u32)(lz - lmin[2]);
  if (!(cx < (u32)dx && cy < (u32)dy && cz < (u32)dz)) return -1;      // belt and braces for inexact box bounds
  u32 key = (cx * (u32)dy + cy) * (u32)dz + cz;
  unsigned h = (sc_hash32(key) >> shift) & mask;
  while (true) {
    u32 k = hkey[h];
    if (k == key) return (int)hid[h];
    if (k == SC_EMPTY32) return -1;
    h = (h + 1) & mask;
  }
}

// deterministic block sum (fixed tree): valid in thread 0
__device__ __forceinline__ float sc_block_sum(float part, float* s_red) {
  const int t = threadIdx.x;
  for (int o = 16; o; o >>= 1) part = part + __shfl_xor_sync(0xffffffffu, part, o);
  if ((t & 31) == 0) s_red[t >> 5] = part;
  __syncthreads();
  float tot = 0.f;
  if (t == 0) { for (int w = 0; w < SC_THREADS / 32; w++) tot = tot + s_red[w]; }
  __syncthreads();
  return tot;
}

extern __shared__ __align__(16) unsigned char sc_dyn[];

// ---- shared-memory kernel (plan modes 0 and 1): one hypothesis per warp ----------------------------
// Each CTA keeps the static voxel table (dense cell -> id, or hash) and the static counts in shared
// memory; each of its 32 warps scores one hypothesis at a time with the 3x4 transform in registers and
// a private row of 16-bit counters (two per 32-bit word, shared-memory atomics), sweeping the moving
// cloud 32 points at a time.  No block-wide synchronisation after the table load.
template <bool POW2>
__global__ void __launch_bounds__(SC_THREADS, 1) score_warp_kernel(const ScArgs* __restrict__ AB) {
  FCCF_PDL_ENTER();
  const ScArgs& A = AB[blockIdx.z];
  const ScoreState* ss = A.ws.ss;
  const int mode = ss->mode;
  if (mode == 2) return;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int nocc = ss->n_occ, n1 = ss->n1, n2 = ss->n2;
  const ScPlan P = sc_plan(ss, A.ws.cap_hash);
  unsigned short* tab16 = (unsigned short*)sc_dyn;                       // mode 0: cells x u16
  u32* hkey = (u32*)sc_dyn; unsigned short* hid16 = (unsigned short*)(hkey + P.cap);   // mode 1
  int* s_cnt = (int*)(sc_dyn + P.tab_bytes);
  u32* row = (u32*)(s_cnt + P.nocc_pad) + (size_t)warp * P.row_words;
  if (mode == 0) { for (unsigned i = t; i < (P.cells + 1) / 2; i += SC_THREADS) ((u32*)tab16)[i] = ((const u32*)A.ws.dense)[i]; }
  else { for (int i = t; i < P.cap; i += SC_THREADS) { hkey[i] = A.ws.hkey[i]; hid16[i] = (unsigned short)A.ws.hid[i]; } }
  for (int i = t; i < P.nocc_pad; i += SC_THREADS) s_cnt[i] = i < nocc ? A.ws.s_cnt[i] : 0;
  __syncthreads();
  const float lo0 = ss->lo_f[0], lo1 = ss->lo_f[1], lo2 = ss->lo_f[2], hi0 = ss->hi_f[0], hi1 = ss->hi_f[1], hi2 = ss->hi_f[2];
  const double m0 = ss->mn[0], m1 = ss->mn[1], m2 = ss->mn[2];
  const int l0 = ss->lmin[0], l1 = ss->lmin[1], l2 = ss->lmin[2];
  const u32 dx = (u32)ss->dims[0], dy = (u32)ss->dims[1], dz = (u32)ss->dims[2];
  const unsigned mask = (unsigned)P.cap - 1u;
  const int shift = 32 - (31 - __clz(P.cap));
  const double rr = POW2 ? 1.0 / (double)A.res : (double)A.res;
  const float* __restrict__ s2 = A.s2;
  // ---- fast sweep: power-of-two lattice with 1/res >= 1 and room in shared memory (ScPlan::fast) ----
  // The 1/res scale is folded into the transform rows and the box (exact: a power of two commutes with every
  // rounding), the moving cloud is read as one float4 per point out of shared memory, and every voxel has a
  // 32-bit counter.  Same counts, same summation order of the score.
  if (POW2 && P.fast && A.res <= 1.0f) {
    const float rf = 1.0f / A.res;
    float4* s2v = (float4*)(sc_dyn + P.s2_off);
    for (int i = t; i < n2; i += SC_THREADS) s2v[i] = make_float4(s2[3 * i], s2[3 * i + 1], s2[3 * i + 2], 0.f);
    __syncthreads();
    u32* roww = (u32*)(s_cnt + P.nocc_pad) + (size_t)warp * P.nocc_pad;
    const float slo0 = lo0 * rf, slo1 = lo1 * rf, slo2 = lo2 * rf, shi0 = hi0 * rf, shi1 = hi1 * rf, shi2 = hi2 * rf;
    const double sm0 = m0 * rr, sm1 = m1 * rr, sm2 = m2 * rr;
    for (int h = blockIdx.x * SC_WARPS + warp; h < A.n_hyp; h += gridDim.x * SC_WARPS) {
      if (A.n_top) { int ty = h / A.topk, kk = h - ty * A.topk; if (kk >= A.n_top[ty]) continue; }
      const float4* Tp = (const float4*)(A.T + (size_t)h * 16);
      float4 r0 = __ldg(Tp), r1 = __ldg(Tp + 1), r2 = __ldg(Tp + 2);
      r0.x *= rf; r0.y *= rf; r0.z *= rf; r0.w *= rf; r1.x *= rf; r1.y *= rf; r1.z *= rf; r1.w *= rf; r2.x *= rf; r2.y *= rf; r2.z *= rf; r2.w *= rf;
      for (int i = lane; i < P.nocc_pad; i += 32) roww[i] = 0u;
      __syncwarp();
#pragma unroll 4
      for (int i = lane; i < n2; i += 32) {
        const float4 p = s2v[i];
        const float qx = p.x * r0.x + (p.y * r0.y + (p.z * r0.z + r0.w));
        const float qy = p.x * r1.x + (p.y * r1.y + (p.z * r1.z + r1.w));
        const float qz = p.x * r2.x + (p.y * r2.y + (p.z * r2.z + r2.w));
        if (qx >= slo0 && qx < shi0 && qy >= slo1 && qy < shi1 && qz >= slo2 && qz < shi2) {
          const u32 cx = (u32)(__double2loint(__dadd_rd((double)qx - sm0, 6755399441055744.0)) - l0);
          const u32 cy = (u32)(__double2loint(__dadd_rd((double)qy - sm1, 6755399441055744.0)) - l1);
          const u32 cz = (u32)(__double2loint(__dadd_rd((double)qz - sm2, 6755399441055744.0)) - l2);
          if (cx < dx && cy < dy && cz < dz) {
            const u32 key = (cx * dy + cy) * dz + cz;
            u32 id = 0xffffu;
            if (mode == 0) id = tab16[key];
            else {
              unsigned hh = (sc_hash32(key) >> shift) & mask;
              while (true) {
                u32 k = hkey[hh];
                if (k == key) { id = hid16[hh]; break; }
                if (k == SC_EMPTY32) break;
                hh = (hh + 1) & mask;
              }
            }
            if (id != 0xffffu) atomicAdd(&roww[id], 1u);
          }
        }
      }
      __syncwarp();
      float part = 0.f;
      for (int w = lane; w < P.row_words; w += 32) {      // the order of the half-word rows: voxels 2w, 2w+1 per step
#pragma unroll
        for (int e = 0; e < 2; e++) {
          const int tt = (int)roww[2 * w + e];
          if (tt > 0) {
            float fs = (float)s_cnt[2 * w + e], ft = (float)tt;
            float mn = fs < ft ? fs : ft, mx = fs > ft ? fs : ft;
            part = part + (fs + ft) * (mn / mx);
          }
        }
      }
      for (int o = 16; o; o >>= 1) part = part + __shfl_xor_sync(0xffffffffu, part, o);
      if (lane == 0) A.scores[h] = part / (float)(n1 + n2);
      __syncwarp();
    }
    return;
  }
  for (int h = blockIdx.x * SC_WARPS + warp; h < A.n_hyp; h += gridDim.x * SC_WARPS) {
    if (A.n_top) { int ty = h / A.topk, kk = h - ty * A.topk; if (kk >= A.n_top[ty]) continue; }
    const float4* Tp = (const float4*)(A.T + (size_t)h * 16);
    const float4 r0 = __ldg(Tp), r1 = __ldg(Tp + 1), r2 = __ldg(Tp + 2);
    for (int i = lane; i < P.row_words; i += 32) row[i] = 0u;
    __syncwarp();
#pragma unroll 4
    for (int i = lane; i < n2; i += 32) {
      const float px = s2[3 * i], py = s2[3 * i + 1], pz = s2[3 * i + 2];
      // pcl::transformPointCloud (SSE order): x*c0 + (y*c1 + (z*c2 + c3))
      const float qx = px * r0.x + (py * r0.y + (pz * r0.z + r0.w));
      const float qy = px * r1.x + (py * r1.y + (pz * r1.z + r1.w));
      const float qz = px * r2.x + (py * r2.y + (pz * r2.z + r2.w));
      if (qx >= lo0 && qx < hi0 && qy >= lo1 && qy < hi1 && qz >= lo2 && qz < hi2) {
        double vx, vy, vz;
        if (POW2) { vx = ((double)qx - m0) * rr; vy = ((double)qy - m1) * rr; vz = ((double)qz - m2) * rr; }
        else { vx = ((double)qx - m0) / rr; vy = ((double)qy - m1) / rr; vz = ((double)qz - m2) / rr; }
        const u32 cx = (u32)(__double2loint(__dadd_rd(vx, 6755399441055744.0)) - l0);
        const u32 cy = (u32)(__double2loint(__dadd_rd(vy, 6755399441055744.0)) - l1);
        const u32 cz = (u32)(__double2loint(__dadd_rd(vz, 6755399441055744.0)) - l2);
        if (cx < dx && cy < dy && cz < dz) {
          const u32 key = (cx * dy + cy) * dz + cz;
          u32 id = 0xffffu;
          if (mode == 0) id = tab16[key];
          else {
            unsigned hh = (sc_hash32(key) >> shift) & mask;
            while (true) {
              u32 k = hkey[hh];
              if (k == key) { id = hid16[hh]; break; }
              if (k == SC_EMPTY32) break;
              hh = (hh + 1) & mask;
            }
          }
          if (id != 0xffffu) atomicAdd(&row[id >> 1], (id & 1u) ? 0x10000u : 1u);
        }
      }
    }
    __syncwarp();
    float part = 0.f;
    for (int w = lane; w < P.row_words; w += 32) {
      const u32 c2 = row[w];
      if (c2) {
#pragma unroll
        for (int e = 0; e < 2; e++) {
          const int tt = (int)((c2 >> (16 * e)) & 0xffffu);
          if (tt > 0) {
            float fs = (float)s_cnt[2 * w + e], ft = (float)tt;
            float mn = fs < ft ? fs : ft, mx = fs > ft ? fs : ft;
            part = part + (fs + ft) * (mn / mx);
          }
        }
      }
    }
    for (int o = 16; o; o >>= 1) part = part + __shfl_xor_sync(0xffffffffu, part, o);
    if (lane == 0) A.scores[h] = part / (float)(n1 + n2);
    __syncwarp();
  }
}

// ---- global-table kernel (plan mode 2) ------------------------------------------------------------
// One CTA scores groups of SC_NH hypotheses; the hash is read from global memory (L2) and the 32-bit
// counters live in the CTA's rows of t_cnt.
template <bool POW2>
__global__ void __launch_bounds__(SC_THREADS) score_kernel(const ScArgs* __restrict__ AB) {
  FCCF_PDL_ENTER();
  const ScArgs& A = AB[blockIdx.z];
  const ScoreState* ss = A.ws.ss;
  if (ss->mode != 2) return;                      // the launcher issues both kernels; one of them runs
  const int t = threadIdx.x;
  const int nocc = ss->n_occ, cap = ss->cap_eff, n1 = ss->n1, n2 = ss->n2;
  __shared__ float s_T[SC_NH][12];
  __shared__ float s_red[SC_THREADS / 32];
  __shared__ float s_lo[3], s_hi[3];
  __shared__ double s_mn[3];
  __shared__ int s_lmin[3];
  if ((int)blockIdx.x * SC_NH + SC_NH > A.ws.t_rows) return;
  const u32* hkey = A.ws.hkey; const u32* hid = A.ws.hid; const int* s_cnt = A.ws.s_cnt;
  int* cnt = A.ws.t_cnt + (size_t)blockIdx.x * SC_NH * A.ws.cap_points;
  const int cstride = A.ws.cap_points;
  if (t < 3) { s_lo[t] = ss->lo_f[t]; s_hi[t] = ss->hi_f[t]; s_mn[t] = ss->mn[t]; s_lmin[t] = ss->lmin[t]; }
  const int dx = ss->dims[0], dy = ss->dims[1], dz = ss->dims[2];
  const unsigned mask = (unsigned)cap - 1u;
  const int shift = 32 - (31 - __clz(cap));
  const double rr = POW2 ? 1.0 / (double)A.res : (double)A.res;
  const int ngroups = (A.n_hyp + SC_NH - 1) / SC_NH;
  const int gstride = min((int)gridDim.x, A.ws.t_rows / SC_NH);
  const float* __restrict__ s2 = A.s2;
  for (int g = blockIdx.x; g < ngroups; g += gstride) {
    const int h0 = g * SC_NH;
    unsigned actm = 0;
#pragma unroll
    for (int k = 0; k < SC_NH; k++) {
      int h = h0 + k;
      bool a = h < A.n_hyp;
      if (a && A.n_top) { int ty = h / A.topk, kk = h - ty * A.topk; a = kk < A.n_top[ty]; }
      if (a) actm |= 1u << k;
    }
    if (!actm) continue;
    __syncthreads();
    if (t < SC_NH * 12) { int k = t / 12, e = t - k * 12; s_T[k][e] = (h0 + k < A.n_hyp) ? A.T[(size_t)(h0 + k) * 16 + e] : 0.f; }
#pragma unroll
    for (int k = 0; k < SC_NH; k++) for (int i = t; i < nocc; i += SC_THREADS) cnt[(size_t)k * cstride + i] = 0;
    __syncthreads();
    for (int i = t; i < n2; i += SC_THREADS) {
      float px = s2[3 * i], py = s2[3 * i + 1], pz = s2[3 * i + 2];
#pragma unroll
      for (int k = 0; k < SC_NH; k++) {
        if (!((actm >> k) & 1u)) continue;
        int id = sc_lookup<POW2>(s_T[k], px, py, pz, s_lo, s_hi, s_mn, rr, s_lmin, dx, dy, dz, hkey, hid, mask, shift);
        if (id >= 0) atomicAdd(&cnt[(size_t)k * cstride + id], 1);
      }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < SC_NH; k++) {
      if (!((actm >> k) & 1u)) continue;             // uniform across the block
      float part = 0.f;
      for (int v = t; v < nocc; v += SC_THREADS) {
        int tt = cnt[(size_t)k * cstride + v];
        if (tt > 0) {
          float fs = (float)s_cnt[v], ft = (float)tt;
          float mn = fs < ft ? fs : ft, mx = fs > ft ? fs : ft;
          part = part + (fs + ft) * (mn / mx);
        }
      }
      float tot = sc_block_sum(part, s_red);
      if (t == 0) A.scores[h0 + k] = tot / (float)(n1 + n2);
    }
  }
}

// per-voxel rows (Lx, Ly, Lz, s, t) of ONE hypothesis, L relative to the voxel of the first static point
template <bool POW2>
__global__ void __launch_bounds__(SC_THREADS) score_dump_kernel(const ScArgs* __restrict__ AB) {
  FCCF_PDL_ENTER();
  const ScArgs& A = AB[blockIdx.z];
  const ScoreState* ss = A.ws.ss;
  const int t = threadIdx.x;
  const int nocc = ss->n_occ, cap = ss->cap_eff, n2 = ss->n2;
  int* cnt = A.ws.t_cnt;
  __shared__ float s_T[12];
  __shared__ float s_lo[3], s_hi[3];
  __shared__ double s_mn[3];
  __shared__ int s_lmin[3];
  if (t < 12) s_T[t] = A.T[t];
  if (t < 3) { s_lo[t] = ss->lo_f[t]; s_hi[t] = ss->hi_f[t]; s_mn[t] = ss->mn[t]; s_lmin[t] = ss->lmin[t]; }
  for (int i = t; i < nocc; i += SC_THREADS) cnt[i] = 0;
  __syncthreads();
  const int dx = ss->dims[0], dy = ss->dims[1], dz = ss->dims[2];
  const unsigned mask = (unsigned)cap - 1u;
  const int shift = 32 - (31 - __clz(cap));
  const double rr = POW2 ? 1.0 / (double)A.res : (double)A.res;
  for (int i = t; i < n2; i += SC_THREADS) {
    int id = sc_lookup<POW2>(s_T, A.s2[3 * i], A.s2[3 * i + 1], A.s2[3 * i + 2], s_lo, s_hi, s_mn, rr, s_lmin, dx, dy, dz, A.ws.hkey, A.ws.hid, mask, shift);
    if (id >= 0) atomicAdd(&cnt[id], 1);
  }
  __syncthreads();
  for (int v = t; v < nocc; v += SC_THREADS) {
    int tt = cnt[v];
    if (tt > 0) {
      int r = atomicAdd(A.nrows, 1);
      if (r < A.cap_rows) {
        u32 k = A.ws.vkey[v];
        int cz = (int)(k % (u32)dz); k /= (u32)dz; int cy = (int)(k % (u32)dy); int cx = (int)(k / (u32)dy);
        // lattice coordinate of the first static point is 1 on every axis
        A.rows[5 * r] = cx + ss->lmin[0] - 1; A.rows[5 * r + 1] = cy + ss->lmin[1] - 1; A.rows[5 * r + 2] = cz + ss->lmin[2] - 1;
        A.rows[5 * r + 3] = A.ws.s_cnt[v]; A.rows[5 * r + 4] = tt;
      }
    }
  }
}

static void fill_common(ScArgs& A, const fccf_params& p, const ScoreWS& ws) {
  A.ws = ws; A.res = p.fine_verify_voxel_size; A.s1 = nullptr; A.n1p = nullptr; A.n2p = nullptr; A.s2 = nullptr; A.T = nullptr; A.n_hyp = 0; A.n_top = nullptr;
  A.topk = 1; A.scores = nullptr; A.rows = nullptr; A.cap_rows = 0; A.nrows = nullptr;
}
static bool is_pow2_res(float res) { int e; float m = frexpf(res, &e); return m == 0.5f && res > 0.f; }

size_t score_ws_layout(ScoreWS* ws, char* base, int cap_points, int t_rows) {
  size_t off = 0;
  auto take = [&](size_t bytes) -> char* { char* p = base ? base + off : nullptr; off += (bytes + 255) & ~(size_t)255; return p; };
  size_t c = (size_t)(cap_points < 1024 ? 1024 : cap_points), nb = c / RS_TILE + 2;
  int cap_hash = 1024; while ((size_t)cap_hash < 2 * c) cap_hash <<= 1;
  ScoreWS w;
  w.cap_points = (int)c; w.cap_hash = cap_hash; w.t_rows = t_rows;
  w.keyA = (u64*)take(8 * c); w.keyB = (u64*)take(8 * c); w.idxA = (u32*)take(4 * c); w.idxB = (u32*)take(4 * c);
  w.hist = (u32*)take(nb * 1024); w.segblk = (u32*)take(nb * 4); w.seg_start = (int*)take(4 * (c + 1));
  w.vkey = (u32*)take(4 * c); w.s_cnt = (int*)take(4 * c);
  w.hkey = (u32*)take(4 * (size_t)cap_hash); w.hid = (u32*)take(4 * (size_t)cap_hash);
  w.dense = (unsigned short*)take(2 * (size_t)SC_DENSE_CELLS + 64);
  w.t_cnt = (int*)take(4 * c * (size_t)t_rows);
  w.ss = nullptr; w.status = nullptr;
  if (ws) *ws = w;
  return off + 256;
}

void launch_score_build(cudaStream_t s, const fccf_params& p, const ScoreBuildJob* jobs, int G, ArgTable& tab, uint64_t* launches, bool lean) {
  std::vector<ScArgs> As(G); std::vector<SortJobs> abs_(G), bas_(G); std::vector<SegJobs> sjs(G);
  int cap = 1, cap_hash = 1;
  for (int g = 0; g < G; g++) {
    const ScoreWS& ws = jobs[g].ws;
    ScArgs& A = As[g]; fill_common(A, p, ws);
    A.s1 = jobs[g].s1; A.n1p = jobs[g].n1; A.n2p = jobs[g].n2;
    if (ws.cap_points > cap) cap = ws.cap_points;
    if (ws.cap_hash > cap_hash) cap_hash = ws.cap_hash;
    SortJob j; j.kin = ws.keyA; j.kout = ws.keyB; j.vin = ws.idxA; j.vout = ws.idxB; j.n = &ws.ss->n_keys; j.nbits = &ws.ss->nbits; j.hist = ws.hist; j.ticket = &ws.ss->tickets[1]; j.miss = ws.status;
    abs_[g].j[0] = j; abs_[g].j[1] = j; abs_[g].j[2] = j;
    SortJob k = j; k.kin = ws.keyB; k.kout = ws.keyA; k.vin = ws.idxB; k.vout = ws.idxA;
    bas_[g].j[0] = k; bas_[g].j[1] = k; bas_[g].j[2] = k;
    SegJob sg; sg.keys = ws.keyA; sg.n = &ws.ss->n_keys; sg.seg_start = ws.seg_start; sg.nseg = &ws.ss->n_occ; sg.blk = ws.segblk; sg.ticket = &ws.ss->tickets[2];
    sjs[g].j[0] = sg; sjs[g].j[1] = sg; sjs[g].j[2] = sg;
  }
  const ScArgs* dA = tab.put(As.data(), G);
  const SortJobs* dab = tab.put(abs_.data(), G); const SortJobs* dba = tab.put(bas_.data(), G); const SegJobs* dsj = tab.put(sjs.data(), G);
  int nb = (cap + 255) / 256; if (nb > 1184) nb = 1184;
  nb = grid_x(nb, G);
  klaunch(score_setup_kernel, dim3(dim3(1, 1, G)), dim3(32), 0, s, dA);
  klaunch(score_bbox_kernel, dim3(dim3(nb, 1, G)), dim3(256), 0, s, dA);
  klaunch(score_keys_kernel, dim3(dim3(nb, 1, G)), dim3(256), 0, s, dA);
  if (launches) *launches += 3;
  launch_sort(s, dab, dba, 1, G, cap, 4, 4, launches, lean);
  launch_segments(s, dsj, 1, G, cap, 4, launches);
  int nbh = (cap_hash + 255) / 256; if (nbh > 1184) nbh = 1184;
  nbh = grid_x(nbh, G);
  klaunch(score_table_kernel, dim3(dim3(nbh, 1, G)), dim3(256), 0, s, dA);
  klaunch(score_insert_kernel, dim3(dim3(nb, 1, G)), dim3(256), 0, s, dA);
  if (launches) *launches += 2;
}

void score_init_attributes() {
  cudaFuncSetAttribute(score_warp_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SC_DYN_BYTES);
  cudaFuncSetAttribute(score_warp_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SC_DYN_BYTES);
}
// As: G per-lane argument blocks with the same n_hyp
static void score_launch(cudaStream_t s, const std::vector<ScArgs>& As, ArgTable& tab, uint64_t* launches) {
  const int G = (int)As.size();
  const int n_hyp = As[0].n_hyp;
  if (n_hyp < 1) return;
  int t_rows = As[0].ws.t_rows;
  for (int g = 1; g < G; g++) t_rows = std::min(t_rows, As[g].ws.t_rows);
  int nbw = (n_hyp + SC_WARPS - 1) / SC_WARPS; if (nbw > 148) nbw = 148;
  int ngroups = (n_hyp + SC_NH - 1) / SC_NH;
  int nbg = t_rows / SC_NH; if (nbg > ngroups) nbg = ngroups; if (nbg > 148) nbg = 148; if (nbg < 1) nbg = 1;
  const ScArgs* dA = tab.put(As.data(), G);
  // which kernel does the work is a device-side fact (table plan); the other one exits at once
  if (is_pow2_res(As[0].res)) {
    klaunch(score_warp_kernel<true>, dim3(dim3(nbw, 1, G)), dim3(SC_THREADS), SC_DYN_BYTES, s, dA);
    klaunch(score_kernel<true>, dim3(dim3(nbg, 1, G)), dim3(SC_THREADS), 0, s, dA);
  } else {
    klaunch(score_warp_kernel<false>, dim3(dim3(nbw, 1, G)), dim3(SC_THREADS), SC_DYN_BYTES, s, dA);
    klaunch(score_kernel<false>, dim3(dim3(nbg, 1, G)), dim3(SC_THREADS), 0, s, dA);
  }
  if (launches) *launches += 2;
}

void launch_score_list(cudaStream_t s, const fccf_params& p, const float* d_T16, int n_hyp, const float* d_s2, const ScoreWS& ws, float* d_scores, ArgTable& tab, uint64_t* launches) {
  if (n_hyp <= 0) return;
  std::vector<ScArgs> As(1); ScArgs& A = As[0]; fill_common(A, p, ws);
  A.T = d_T16; A.n_hyp = n_hyp; A.s2 = d_s2; A.scores = d_scores;
  score_launch(s, As, tab, launches);
}

void launch_score_dump(cudaStream_t s, const fccf_params& p, const float* d_T16, const float* d_s2, const ScoreWS& ws, int* d_rows, int cap_rows, int* d_nrows, ArgTable& tab, uint64_t* launches) {
  ScArgs A; fill_common(A, p, ws);
  A.T = d_T16; A.n_hyp = 1; A.s2 = d_s2; A.rows = d_rows; A.cap_rows = cap_rows; A.nrows = d_nrows;
  const ScArgs* dA = tab.put(&A, 1);
  if (is_pow2_res(A.res)) klaunch(score_dump_kernel<true>, dim3(1), dim3(SC_THREADS), 0, s, dA);
  else klaunch(score_dump_kernel<false>, dim3(1), dim3(SC_THREADS), 0, s, dA);
  if (launches) *launches += 1;
}

// ---------------------------------------------------------------------------------------------
// Block-reduced argmax over a score list: packed = (order-preserving int of the score) << 32 |
// (0xFFFFFFFF - global index); the signed 64-bit maximum is the highest score, smallest index among
// ties (the reference's strict `>` first-maximum scans, FCCF.cpp:1559).  NaN ranks below everything.
// One atomicMax per block; across GPUs the same word goes through one 8-byte all-reduce (max).
__device__ __forceinline__ long long pack_score(float sc, long long gidx) {
  int key;
  if (sc != sc) key = (int)0x80000000;
  else { sc = sc + 0.0f; int b = __float_as_int(sc); key = b >= 0 ? b : (b ^ 0x7fffffff); }
  return ((long long)key << 32) | (long long)(0xffffffffll - gidx);
}
__global__ void __launch_bounds__(256) score_best_kernel(const float* __restrict__ scores, int n, long long index_base, long long* out) {
  FCCF_PDL_ENTER();
  long long best = (long long)0x8000000000000000ull;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    long long p = pack_score(scores[i], index_base + i);
    best = p > best ? p : best;
  }
  for (int o = 16; o; o >>= 1) { long long q = __shfl_xor_sync(0xffffffffu, best, o); best = q > best ? q : best; }
  __shared__ long long s_b[8];
  if ((threadIdx.x & 31) == 0) s_b[threadIdx.x >> 5] = best;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; w++) best = s_b[w] > best ? s_b[w] : best;
    atomicMax(out, best);
  }
}
void launch_score_best(cudaStream_t s, const float* d_scores, int n, long long index_base, long long* d_out, uint64_t* launches) {
  cudaMemsetAsync(d_out, 0, 8, s);
  // LLONG_MIN = 0x8000...: set the top byte
  static const unsigned char top = 0x80;
  cudaMemsetAsync((char*)d_out + 7, top, 1, s);
  if (n <= 0) return;
  int nb = (n + 255) / 256; if (nb > 592) nb = 592;
  klaunch(score_best_kernel, dim3(nb), dim3(256), 0, s, d_scores, n, index_base, d_out);
  if (launches) *launches += 1;
}

// ---------------------------------------------------------------------------------------------
struct FuseArgs { PipeState* st; const float* top_T; const float* top_s1; const float* top_s2; float fine_number; int topk; };

// FCCF.cpp:1546-1606 + fuse_answer 1291-1368
__global__ void fuse_kernel(const FuseArgs* __restrict__ AB) {
  FCCF_PDL_ENTER();
  const FuseArgs& A = AB[blockIdx.x];
  PipeState* st = A.st;
  float score_sum = 0.f, score1_sum = 0.f, score2_sum = 0.f;
  for (int ty = 0; ty < 3; ty++)
    for (int k = 0; k < st->n_top[ty]; k++) { score2_sum += A.top_s2[ty * A.topk + k]; score1_sum += A.top_s1[ty * A.topk + k]; }
  float best_best = 0.f;
  float hs_score[3]; q4 hs_q[3]; f3 hs_t[3];
  for (int ty = 0; ty < 3; ty++) {
    float best_score = 0.f;
    float tb[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
    for (int k = 0; k < st->n_top[ty]; k++) {
      float score = A.top_s1[ty * A.topk + k] / score1_sum + A.top_s2[ty * A.topk + k] / score2_sum;
      if (score > best_score) { best_score = score; for (int i = 0; i < 12; i++) tb[i] = A.top_T[((size_t)ty * A.topk + k) * 16 + i]; }
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

// per-type best + gate + fusion alone (stand-alone stage entry point fccf_fuse)
void launch_fuse(cudaStream_t s, PipeState* st, const float* top_T, const float* top_s1, const float* top_s2, float fine_number, int topk, ArgTable& tab, uint64_t* launches) {
  FuseArgs F; F.st = st; F.top_T = top_T; F.top_s1 = top_s1; F.top_s2 = top_s2; F.fine_number = fine_number; F.topk = topk;
  klaunch(fuse_kernel, dim3(1), dim3(1), 0, s, tab.put(&F, 1));
  if (launches) *launches += 1;
}

// fine verify in two parts: the static voxel table of cloud 1's leftover points (depends only on the plane
// stage, so a captured graph runs it on a branch beside hypotheses / clustering / quick verify), then the
// scoring of the selected centres and the fusion.
static ScoreWS fv_ws(const Work& w) { ScoreWS ws = w.h.fv; ws.ss = &w.st->fv; ws.status = &w.st->status; return ws; }

void launch_fine_verify_build(cudaStream_t s, const Batch& b, uint64_t* launches) {
  const int G = b.G;
  std::vector<ScoreBuildJob> jobs(G);
  for (int g = 0; g < G; g++) {
    const Work& w = b.w[g]; PipeState* st = w.st;
    jobs[g].s1 = w.c[0].sub; jobs[g].n1 = &st->oct[0].S; jobs[g].n2 = &st->oct[1].S; jobs[g].ws = fv_ws(w);
  }
  launch_score_build(s, b.p, jobs.data(), G, *b.tab, launches, b.lean);
}

void launch_fine_verify_fuse(cudaStream_t s, const Batch& b, uint64_t* launches) {
  const int G = b.G;
  std::vector<ScArgs> As(G); std::vector<FuseArgs> Fs(G);
  for (int g = 0; g < G; g++) {
    const Work& w = b.w[g]; const HypWS& h = w.h;
    PipeState* st = w.st;
    ScArgs& A = As[g]; fill_common(A, b.p, fv_ws(w));
    A.T = h.top_T; A.n_hyp = 3 * fccf_topk(b.p); A.n_top = st->n_top; A.topk = fccf_topk(b.p); A.s2 = w.c[1].sub; A.scores = h.top_s2;
    FuseArgs& F = Fs[g]; F.st = st; F.top_T = h.top_T; F.top_s1 = h.top_s1; F.top_s2 = h.top_s2; F.fine_number = b.p.fine_verify_number; F.topk = fccf_topk(b.p);
  }
  score_launch(s, As, *b.tab, launches);
  klaunch(fuse_kernel, dim3(G), dim3(1), 0, s, b.tab->put(Fs.data(), G));
  if (launches) *launches += 1;
}

}  // namespace fccf
