// voxelgrid_fast.cu — pcl::VoxelGrid<PointXYZ>::filter (FCCF.cpp:1668-1678 / 1377-1387) for clouds
// that fit ONE thread-block cluster: the whole filter of a cloud is one pass of a cluster of CS CTAs
// (CS x 32 warps), so the raw cloud crosses HBM once and every intermediate stays on chip or in a
// per-cluster scratch that never leaves L2.  Same results, bit for bit, as the generic path
// (voxelgrid.cu + sort.cu): cell index ijk = (int)(floorf(p * inv) - (float)min_b), one output per
// occupied cell in ascending cell order, float32 running sum in ascending point index, / n.
//
//   A  bounding box: warp gw = rank*32 + w owns the contiguous run [gw*RL, (gw+1)*RL) of the cloud
//      (128-bit loads staged through a per-warp shared-memory tile -> stride-3 reads), warp / CTA
//      reduction, exchange of the CTA results through DISTRIBUTED SHARED MEMORY, grid set-up
//      computed redundantly by every CTA (pcl 1.10 applyFilter, incl. the int32 bail-out test)
//   B  cell c of every point: row (slot) = c / DIG, owner digit d = (c + ROT * slot) % DIG (DIG = CS*32
//      owner warps; the row-dependent rotation spreads walls and floors evenly over the owners).  Rank of
//      the point among the points of its digit in index order = running count of (warp, d), bumped with
//      per-bit ballots, + prefix over the warps of the CTA (shared memory) + prefix over the CTAs (the
//      owner warp of d reads / writes the CTA tables through DSMEM) + prefix over the digits.  The
//      rank word (slot, digit, rank inside the warp's run) goes to the scratch, 4 bytes per point.
//   B2 every point is written as (x, y, z, slot) at its rank into the queue of its digit: the queue
//      of a digit lists its points in ascending index
//   C  owner warp (digit d = w*CS + rank) walks its queue 32 entries at a time: lanes of one round
//      that hit the same cell are found with ballots over the slot bits and applied in lane order,
//      so every cell's float32 sum runs in ascending point index (pcl::CentroidPoint order);
//      accumulators are compact (first touch) behind a dense slot -> id table, all in shared memory
//   D  occupancy masks of the DIG cells of each row exchanged through DSMEM, row prefix, and the
//      owners write centroid / cell / count at the cell's rank among the occupied cells
// Clouds the cluster cannot hold (pcl's bail-out, more than MAXCELLS cells, more than VF_ACC occupied
// cells for one owner warp, more points than the scratch) raise ST_VG_FAST_MISS and the host re-runs
// the registration through the generic kernels.
#include <cooperative_groups.h>
#include "fccf_dev.cuh"
#include "fccf_internal.h"
#include <cstdlib>
#include <vector>

namespace cg = cooperative_groups;

namespace fccf {

#define VF_T 1024
#define VF_NW 32
#define VF_ACC 160                    // occupied cells per owner warp
#define VF_ROT 37                     // digit rotation per row (odd)
#define VF_NMAX 524032                // points per cloud the cluster path takes

struct VFArgs {
  const CallArgs* call;      // leaf; raw clouds when in[c] == nullptr (stage 0)
  const float* in[2];
  const int* n_in[2];
  VGState* st[2];
  float* out[2]; long long* cell[2]; int* cnt[2];
  int* status;
  int emulate;
  long long* prof;           // clock64() marks of cluster 0 / CTA 0 (debug blob "prof", slots 22..31); nullptr: none
};

// compile-time geometry of a cluster of CS CTAs
template <int CS> struct VfGeo {
  static constexpr int LOGCS = CS == 8 ? 3 : 2;
  static constexpr int DB = 5 + LOGCS;                  // digit bits
  static constexpr int DIG = VF_NW * CS;                // owner warps = digits
  static constexpr int MAXSLOT = CS == 8 ? 768 : 1024;  // rows of DIG cells
  static constexpr int TABSTRIDE = MAXSLOT + 2;         // u16 row stride of the slot -> id table (odd number of words: conflict-free column reads)
  static constexpr int MAXCELLS = DIG * MAXSLOT;
  // shared-memory layout (dynamic)
  static constexpr int OFF_ACC = 0;                                  // float4 [32][VF_ACC]
  static constexpr int OFF_TAB = OFF_ACC + VF_NW * VF_ACC * 16;      // u16 [32][TABSTRIDE]
  static constexpr int OFF_CNTW = OFF_TAB + VF_NW * TABSTRIDE * 2;   // u32 [32][DIG]
  static constexpr int OFF_MASK = OFF_CNTW + VF_NW * DIG * 4;        // u32 [CS][MAXSLOT]
  static constexpr int OFF_OWN = OFF_MASK + CS * MAXSLOT * 4;        // u32 [MAXSLOT]
  static constexpr int OFF_ROWB = OFF_OWN + MAXSLOT * 4;             // u32 [MAXSLOT] occupied cells before the row
  static constexpr int OFF_ROWP = OFF_ROWB + MAXSLOT * 4;            // u32 [MAXSLOT] occupied digits below the row's first digit | row total << 16
  static constexpr int OFF_QST = OFF_ROWP + MAXSLOT * 4;             // u32 [DIG] queue start of every digit
  static constexpr int OFF_CCNT = OFF_QST + DIG * 4;                 // u32 [DIG] this CTA's count per digit
  static constexpr int OFF_CBASE = OFF_CCNT + DIG * 4;               // u32 [DIG] this CTA's base inside the digit's queue
  static constexpr int OFF_STG = OFF_CBASE + DIG * 4;                // float4 [32][24]
  static constexpr int OFF_TOT = OFF_STG + VF_NW * 24 * 16;          // u32 [32] entries of the digits this CTA owns
  static constexpr int OFF_RED = OFF_TOT + VF_NW * 4;                // int [32][8]
  static constexpr int OFF_CTA = OFF_RED + VF_NW * 8 * 4;            // int [8]
  static constexpr int OFF_GLOB = OFF_CTA + 32;                      // int [8]
  static constexpr int OFF_FAIL = OFF_GLOB + 32;                     // int [8]
  static constexpr int SMEM_BYTES = OFF_FAIL + 32;
};

extern __shared__ __align__(16) unsigned char vf_smem[];

// Points [first, first + 32*nr) of a cloud, one round of 32 at a time.  The first `nfull` rounds are whole and
// 16-byte aligned: lane l < 24 fetches 16 bytes of the round's 384, three rounds ahead, the round is staged in the
// warp's shared-memory tile and read back as x y z.  The last (partial) round, or every round of a cloud
// that is not 16-byte aligned, takes scalar loads.
struct VfLoader {
  const float4* src;       // this lane's 16 bytes of round 0
  const float* p;
  float4* st4; const float* stx;
  int first, n, nfull, lane; float4 v0, v1, v2;
  __device__ __forceinline__ float4 fetch(int k) const {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (k < nfull && lane < 24) v = __ldg(src + k * 24);
    return v;
  }
  __device__ __forceinline__ void init(const float* p_, int first_, int n_, int nr_, int lane_, float* stg_) {
    p = p_; first = first_; n = n_; lane = lane_;
    st4 = (float4*)stg_ + lane_; stx = stg_ + 3 * lane_;
    src = (const float4*)p_ + ((size_t)first_ * 3) / 4 + lane_;
    nfull = ((((size_t)p_) & 15) == 0 && first_ < n_) ? min(nr_, (n_ - first_) >> 5) : 0;
    v0 = fetch(0); v1 = fetch(1); v2 = fetch(2);
  }
  __device__ __forceinline__ bool get(int k, float& x, float& y, float& z) {
    bool valid = true;
    if (k < nfull) {
      if (lane < 24) *st4 = v0;
      __syncwarp();
      x = stx[0]; y = stx[1]; z = stx[2];
      __syncwarp();
    } else {
      const int i = first + k * 32 + lane;
      valid = i < n;
      x = y = z = 0.f;
      if (valid) { const float* q = p + 3 * (size_t)i; x = q[0]; y = q[1]; z = q[2]; }
    }
    v0 = v1; v1 = v2; v2 = fetch(k + 3);
    return valid;
  }
};

// Lanes that take part (`in`) and whose key has the same low `nbits` bits as this lane's: per-bit ballots
// (match.any was measured slower on sm_100a: 2800 against 2000 cycles per round of phase B).
__device__ __forceinline__ unsigned vf_group(bool in, int v, int nbits) {
  unsigned p = __ballot_sync(0xffffffffu, in);
  for (int b = 0; b < nbits; b++) {
    const int m = (v << (31 - b)) >> 31;                       // 0 / -1
    const unsigned bal = __ballot_sync(0xffffffffu, m != 0);
    p &= ~(bal ^ (unsigned)m);
  }
  return p;
}

template <int CS>
__global__ void __launch_bounds__(VF_T, 1) vg_fast_kernel(const VFArgs* __restrict__ AB, int ncloud, int nitems, u32* __restrict__ info_all,
                                                          float4* __restrict__ queue_all, int stride) {
  FCCF_PDL_ENTER();
  typedef VfGeo<CS> G;
  constexpr int DIG = G::DIG, DB = G::DB, MAXSLOT = G::MAXSLOT, TABSTRIDE = G::TABSTRIDE;
  cg::cluster_group cluster = cg::this_cluster();
  const int r = (int)cluster.block_rank();
  const int cid = blockIdx.x / CS, ncl = gridDim.x / CS;
  const int t = threadIdx.x, lane = t & 31, w = t >> 5;
  const unsigned lt = (1u << lane) - 1u;
  float4* acc = (float4*)(vf_smem + G::OFF_ACC);
  unsigned short* tab = (unsigned short*)(vf_smem + G::OFF_TAB);
  u32* cntw = (u32*)(vf_smem + G::OFF_CNTW);
  u32* allmask = (u32*)(vf_smem + G::OFF_MASK);
  u32* m_own = (u32*)(vf_smem + G::OFF_OWN);
  u32* s_rowb = (u32*)(vf_smem + G::OFF_ROWB);
  u32* s_rowp = (u32*)(vf_smem + G::OFF_ROWP);
  u32* s_qst = (u32*)(vf_smem + G::OFF_QST);
  u32* s_ccnt = (u32*)(vf_smem + G::OFF_CCNT);
  u32* s_cbase = (u32*)(vf_smem + G::OFF_CBASE);
  float* stgw = (float*)(vf_smem + G::OFF_STG) + w * 96;
  u32* s_tot = (u32*)(vf_smem + G::OFF_TOT);
  int* s_red = (int*)(vf_smem + G::OFF_RED);
  int* s_cta = (int*)(vf_smem + G::OFF_CTA);
  int* s_glob = (int*)(vf_smem + G::OFF_GLOB);
  int* s_fail = (int*)(vf_smem + G::OFF_FAIL);
  u32* info = info_all + (size_t)cid * stride;      // rank word of every point, in index order
  float4* queue = queue_all + (size_t)cid * stride;
  const int gw = r * VF_NW + w;
  u32* cw = cntw + w * DIG;                          // this warp's counter row
  unsigned short* tw = tab + w * TABSTRIDE;          // this (owner) warp's slot -> id row
  float4* aw = acc + w * VF_ACC;

  for (int item = cid; item < nitems; item += ncl) {
    const VFArgs& A = AB[item / ncloud];
    const int c = item % ncloud;
    VGState* st = A.st[c];
    const int n = *A.n_in[c];
    const float* p = A.in[c] ? A.in[c] : A.call->raw[c];
    const int RL = (((n + DIG - 1) / DIG) + 31) & ~31;       // run of one warp, a multiple of 32 points
    const int first = gw * RL;
    const int nr = first < n ? (min(n - first, RL) + 31) / 32 : 0;
    VfLoader L;
#define VF_MARK(k) if (A.prof && c == 0 && r == 0 && t == 0) A.prof[k] = clock64();
    VF_MARK(0)

    // ---- A: bounding box of the finite points -------------------------------------------------
    {
      int mn0 = 0x7fffffff, mn1 = 0x7fffffff, mn2 = 0x7fffffff, mx0 = (int)0x80000000, mx1 = (int)0x80000000, mx2 = (int)0x80000000, nf = 0;
      L.init(p, first, n, nr, lane, stgw);
      for (int k = 0; k < nr; k++) {
        float x, y, z;
        const bool valid = L.get(k, x, y, z);
        if (valid && isfinite(x) && isfinite(y) && isfinite(z)) {
          nf++;
          const int ox = f2ord(x), oy = f2ord(y), oz = f2ord(z);
          mn0 = min(mn0, ox); mn1 = min(mn1, oy); mn2 = min(mn2, oz);
          mx0 = max(mx0, ox); mx1 = max(mx1, oy); mx2 = max(mx2, oz);
        }
      }
      mn0 = __reduce_min_sync(0xffffffffu, mn0); mn1 = __reduce_min_sync(0xffffffffu, mn1); mn2 = __reduce_min_sync(0xffffffffu, mn2);
      mx0 = __reduce_max_sync(0xffffffffu, mx0); mx1 = __reduce_max_sync(0xffffffffu, mx1); mx2 = __reduce_max_sync(0xffffffffu, mx2);
      nf = __reduce_add_sync(0xffffffffu, nf);
      if (lane == 0) { int* q = s_red + w * 8; q[0] = mn0; q[1] = mn1; q[2] = mn2; q[3] = mx0; q[4] = mx1; q[5] = mx2; q[6] = nf; }
    }
    VF_MARK(1)
    // warp-private tables of this item (nothing of the previous item is read any more: closing cluster barrier below)
#pragma unroll
    for (int i = lane; i < DIG; i += 32) cw[i] = 0u;
    if (t < 8) s_fail[t] = 0;
    __syncthreads();
    if (w == 0) {
      int v[7];
#pragma unroll
      for (int k = 0; k < 7; k++) v[k] = s_red[lane * 8 + k];
#pragma unroll
      for (int k = 0; k < 3; k++) { v[k] = __reduce_min_sync(0xffffffffu, v[k]); v[3 + k] = __reduce_max_sync(0xffffffffu, v[3 + k]); }
      v[6] = __reduce_add_sync(0xffffffffu, v[6]);
      if (lane == 0) {
#pragma unroll
        for (int k = 0; k < 7; k++) s_cta[k] = v[k];
      }
    }
    cluster.sync();                                                                       // #1
    if (w == 0) {
      // lane = 8 * (CTA in a group of four) + value: one DSMEM read per lane and group, then a butterfly over the groups
      int v = 0;
      const int k = lane & 7;
#pragma unroll
      for (int r0 = 0; r0 < CS; r0 += 4) {
        const int rr = r0 + (lane >> 3);
        const int x = (k < 7) ? cluster.map_shared_rank(s_cta, rr)[k] : 0;
        if (r0 == 0) v = x; else v = (k < 3) ? min(v, x) : (k < 6 ? max(v, x) : v + x);
      }
#pragma unroll
      for (int o = 8; o < 32; o <<= 1) { const int y = __shfl_xor_sync(0xffffffffu, v, o); v = (k < 3) ? min(v, y) : (k < 6 ? max(v, y) : v + y); }
      if (lane < 7) s_glob[lane] = v;
    }
    __syncthreads();
    // grid set-up (pcl 1.10 voxel_grid.hpp applyFilter), the same in every thread of the cluster
    const int nfin = s_glob[6];
    const float inv = 1.0f / A.call->leaf;
    int bail = 0, minb[3] = {0, 0, 0};
    long long div[3] = {1, 1, 1}, total = 0;
    bool fast = n <= stride && n <= VF_NMAX;
    if (nfin > 0) {
      float fmn[3], fmx[3];
#pragma unroll
      for (int a = 0; a < 3; a++) { fmn[a] = ord2f(s_glob[a]); fmx[a] = ord2f(s_glob[3 + a]); }
      const long long dx = (long long)((fmx[0] - fmn[0]) * inv) + 1, dy = (long long)((fmx[1] - fmn[1]) * inv) + 1, dz = (long long)((fmx[2] - fmn[2]) * inv) + 1;
      bail = (A.emulate && (dx * dy * dz) > 2147483647LL) ? 1 : 0;
      if (!bail) {
#pragma unroll
        for (int a = 0; a < 3; a++) {
          const int lo = (int)floorf(fmn[a] * inv), hi = (int)floorf(fmx[a] * inv);
          minb[a] = lo; div[a] = (long long)hi - lo + 1;
        }
        if (div[0] > G::MAXCELLS || div[1] > G::MAXCELLS || div[2] > G::MAXCELLS) fast = false;
        else { total = div[0] * div[1] * div[2]; if (total > G::MAXCELLS) fast = false; }
      } else fast = false;
    }
    if (!fast) {
      if (r == 0 && t == 0) { atomicOr(A.status, ST_VG_FAST_MISS); st->n_in = n; st->n_out = 0; st->n_finite = 0; }
      cluster.sync();                                                                     // closing barrier of the item
      continue;
    }
    if (r == 0 && t == 0) {
      st->n_in = n; st->n_finite = nfin; st->inv = inv; st->bail = 0; st->total = total;
      for (int a = 0; a < 3; a++) { st->mn[a] = s_glob[a]; st->mx[a] = s_glob[3 + a]; st->minb[a] = minb[a]; st->div[a] = div[a]; }
      const int nb = 64 - __clzll(total);
      st->nbits = nb < 1 ? 1 : nb;
      if (nfin == 0) st->n_out = 0;
    }
    if (nfin == 0) { cluster.sync(); continue; }
    const int nslots = (int)((total + DIG - 1) / DIG);
    const int sbits = nslots > 1 ? 32 - __clz(nslots - 1) : 0;
    // this owner warp's slot -> id row (read again in phase C; no other warp touches it before phase D)
    for (int i = lane; i < (nslots + 1) / 2; i += 32) ((u32*)tw)[i] = 0xffffffffu;

    VF_MARK(2)
    // ---- B: cell of every point, rank inside (warp, digit) ------------------------------------
    {
      const float mb0 = (float)minb[0], mb1 = (float)minb[1], mb2 = (float)minb[2];
      const int d0 = (int)div[0], d01 = (int)(div[0] * div[1]);
      u32* infow = info + first + lane;
      L.init(p, first, n, nr, lane, stgw);
      for (int k = 0; k < nr; k++) {
        float x, y, z;
        const bool valid = L.get(k, x, y, z);
        const bool ok = valid && isfinite(x) && isfinite(y) && isfinite(z);
        int ci = 0;
        if (ok) {
          const int i0 = (int)(floorf(x * inv) - mb0);
          const int i1 = (int)(floorf(y * inv) - mb1);
          const int i2 = (int)(floorf(z * inv) - mb2);
          ci = i0 + i1 * d0 + i2 * d01;
        }
        const int slot = ci >> DB;
        const int d = (ci + VF_ROT * slot) & (DIG - 1);
        const unsigned peers = vf_group(ok, d, DB);
        // running count of (warp, digit): every lane reads it, the first lane of each group bumps it
        const u32 old = cw[d];
        __syncwarp();
        u32 word = 0xffffffffu;
        if (ok) {
          if ((peers & lt) == 0u) cw[d] = old + (u32)__popc(peers);
          word = ((u32)slot << 22) | ((u32)d << 14) | (old + (u32)__popc(peers & lt));
        }
        if (valid) infow[k * 32] = word;
        __syncwarp();
      }
    }
    VF_MARK(3)
    __syncthreads();
    // prefix over the warps of this CTA, per digit
    if (t < DIG) {
      u32 run = 0;
#pragma unroll 8
      for (int w2 = 0; w2 < VF_NW; w2++) { const u32 v = cntw[w2 * DIG + t]; cntw[w2 * DIG + t] = run; run += v; }
      s_ccnt[t] = run;
    }
    cluster.sync();                                                                       // #2
    // owner warp of digit d = w*CS + r: prefix over the CTAs, written back into their tables
    {
      const int d = w * CS + r;
      u32 v = 0;
      if (lane < CS) v = cluster.map_shared_rank(s_ccnt, lane)[d];
      u32 inc = v;
#pragma unroll
      for (int o = 1; o < CS; o <<= 1) { const u32 y = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += y; }
      if (lane < CS) cluster.map_shared_rank(s_cbase, lane)[d] = inc - v;
      const u32 tot = __shfl_sync(0xffffffffu, inc, CS - 1);
      if (lane == 0) s_tot[w] = tot;
    }
    cluster.sync();                                                                       // #3
    // queue start of every digit: exclusive prefix of the DIG totals (digit d is owned by CTA d % CS, warp d / CS)
    if (t < DIG) s_qst[t] = cluster.map_shared_rank(s_tot, t & (CS - 1))[t >> G::LOGCS];
    __syncthreads();
    if (w == 0) {
      constexpr int PER = DIG / 32;
      u32 v[PER], s = 0;
#pragma unroll
      for (int k = 0; k < PER; k++) { v[k] = s_qst[lane * PER + k]; s += v[k]; }
      u32 inc = s;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const u32 y = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += y; }
      u32 run = inc - s;
#pragma unroll
      for (int k = 0; k < PER; k++) { s_qst[lane * PER + k] = run; run += v[k]; }
    }
    __syncthreads();
    // this warp's row becomes the absolute queue position of its first point of every digit
#pragma unroll
    for (int i = lane; i < DIG; i += 32) cw[i] += s_qst[i] + s_cbase[i];
    __syncwarp();

    VF_MARK(4)
    // ---- B2: points into the queues of their digits (the cloud is read a third time, out of L2) ----
    {
      const u32* infow = info + first + lane;
      L.init(p, first, n, nr, lane, stgw);
      u32 w0 = 0xffffffffu, w1 = 0xffffffffu;
      const int lim = n - first - lane;                  // rounds k with k*32 < lim hold a point of this lane
      if (0 < lim) w0 = __ldcg(infow);
      if (32 < lim) w1 = __ldcg(infow + 32);
      for (int k = 0; k < nr; k++) {
        float x, y, z;
        L.get(k, x, y, z);
        const u32 word = w0;
        w0 = w1; w1 = 0xffffffffu;
        if ((k + 2) * 32 < lim) w1 = __ldcg(infow + (k + 2) * 32);
        if (word != 0xffffffffu) {
          const u32 pos = cw[(word >> 14) & 255u] + (word & 0x3fffu);
          queue[pos] = make_float4(x, y, z, __int_as_float((int)(word >> 22)));
        }
      }
    }
    VF_MARK(5)
    __threadfence();
    cluster.sync();                                                                       // #4
    VF_MARK(6)

    // ---- C: in-order accumulation by the owner warps ------------------------------------------
    {
      const int d = w * CS + r;
      const int cq = (int)s_tot[w];
      const float4* qw = queue + s_qst[d] + lane;
      const int lim = cq - lane;
      int nalloc = 0;
      float4 qn = make_float4(0.f, 0.f, 0.f, 0.f), qn2 = qn;
      if (0 < lim) qn = __ldcg(qw);
      if (32 < lim) qn2 = __ldcg(qw + 32);
      for (int k0 = 0; k0 < cq; k0 += 32) {
        const float4 q = qn;
        const bool valid = k0 < lim;
        qn = qn2;
        if (k0 + 64 < lim) qn2 = __ldcg(qw + k0 + 64);
        const int slot = valid ? __float_as_int(q.w) : 0;
        const unsigned peers = vf_group(valid, slot, sbits);
        const int rk = valid ? __popc(peers & lt) : -1;
        int id = valid ? (int)tw[slot] : 0;
        const bool isnew = valid && rk == 0 && id == 0xffff;
        const unsigned nm = __ballot_sync(0xffffffffu, isnew);
        if (nm) {
          if (isnew) {
            id = nalloc + __popc(nm & lt);
            if (id < VF_ACC) { tw[slot] = (unsigned short)id; aw[id] = make_float4(0.f, 0.f, 0.f, __int_as_float(0)); }
          }
          nalloc += __popc(nm);
          if (nalloc > VF_ACC) { if (lane == 0) s_fail[0] = 1; break; }     // warp-uniform
          __syncwarp();
          if (valid) id = (int)tw[slot];
        }
        const int maxrk = __reduce_max_sync(0xffffffffu, rk);
        for (int s2 = 0; s2 <= maxrk; s2++) {
          if (rk == s2) {
            float4 a = aw[id];
            a.x += q.x; a.y += q.y; a.z += q.z; a.w = __int_as_float(__float_as_int(a.w) + 1);
            aw[id] = a;
          }
          __syncwarp();
        }
      }
    }
    VF_MARK(7)
    __syncthreads();
    // ---- D: occupancy rows, ranks, output -----------------------------------------------------
    for (int s = w; s < nslots; s += VF_NW) {
      const unsigned b = __ballot_sync(0xffffffffu, tab[lane * TABSTRIDE + s] != 0xffff);
      if (lane == 0) m_own[s] = b;
    }
    cluster.sync();                                                                       // #5
    for (int i = t; i < CS * nslots; i += VF_T) { const int rr = i / nslots, s = i - rr * nslots; allmask[rr * MAXSLOT + s] = cluster.map_shared_rank(m_own, rr)[s]; }
    if (t < CS) s_glob[t] = cluster.map_shared_rank(s_fail, t)[0];
    __syncthreads();
    int failed = 0;
#pragma unroll
    for (int k = 0; k < CS; k++) failed |= s_glob[k];
    if (failed) {
      if (r == 0 && t == 0) { atomicOr(A.status, ST_VG_FAST_MISS); st->n_out = 0; }
      cluster.sync();
      continue;
    }
    // occupied digits below digit x of row s (digit x = warp wx * CS + CTA rx)
    auto below_digit = [&](int s, int x) -> u32 {
      const int wx = x >> G::LOGCS, rx = x & (CS - 1);
      const unsigned lo = (1u << wx) - 1u, upto = (2u << wx) - 1u;
      u32 cnt2 = 0;
#pragma unroll
      for (int rr = 0; rr < CS; rr++) cnt2 += (u32)__popc(allmask[rr * MAXSLOT + s] & (rr < rx ? upto : lo));
      return cnt2;
    };
    // per row: its total and the occupied digits below its first digit (the digit of cell j = 0)
    for (int s = t; s < nslots; s += VF_T) {
      u32 tot = 0;
#pragma unroll
      for (int rr = 0; rr < CS; rr++) tot += (u32)__popc(allmask[rr * MAXSLOT + s]);
      s_rowp[s] = below_digit(s, (VF_ROT * s) & (DIG - 1)) | (tot << 16);
    }
    __syncthreads();
    if (w == 0) {
      // exclusive prefix of the occupied cells per row
      const int per = (nslots + 31) / 32;
      u32 s = 0;
      for (int k = 0; k < per; k++) { const int row = lane * per + k; if (row < nslots) s += s_rowp[row] >> 16; }
      u32 inc = s;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const u32 y = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += y; }
      u32 run = inc - s;
      for (int k = 0; k < per; k++) { const int row = lane * per + k; if (row < nslots) { s_rowb[row] = run; run += s_rowp[row] >> 16; } }
      const u32 tot = __shfl_sync(0xffffffffu, inc, 31);
      if (lane == 0 && r == 0) st->n_out = (int)tot;
    }
    __syncthreads();
    {
      float* out = A.out[c]; long long* cell = A.cell[c]; int* cnt = A.cnt[c];
      for (int s = w; s < nslots; s += VF_NW) {
        const int id = (int)tab[lane * TABSTRIDE + s];
        if (id != 0xffff) {
          // cells of row s in ascending order: j = 0 .. DIG-1 sits at digit (j + o) % DIG, o = ROT*s % DIG
          const int d = lane * CS + r, o = (VF_ROT * s) & (DIG - 1);
          const u32 rp = s_rowp[s], po = rp & 0xffffu, tot = rp >> 16;
          const u32 pd = below_digit(s, d);
          const u32 pos = s_rowb[s] + (d >= o ? pd - po : pd + tot - po);
          const float4 a = acc[lane * VF_ACC + id];
          const int m = __float_as_int(a.w);
          const float fn = (float)m;
          out[3 * (size_t)pos] = a.x / fn; out[3 * (size_t)pos + 1] = a.y / fn; out[3 * (size_t)pos + 2] = a.z / fn;
          cell[pos] = (long long)(s * DIG + ((d - o) & (DIG - 1)));
          cnt[pos] = m;
        }
      }
    }
    VF_MARK(8)
    cluster.sync();                                                                       // closing barrier of the item
    VF_MARK(9)
  }
}

// Co-resident clusters of each cluster size (cudaOccupancyMaxActiveClusters; B200: 15 clusters of 8 = 120 SMs,
// 33 clusters of 4 = 132 SMs).  A launch with few clouds takes clusters of 8 (a cloud finishes sooner: 195k
// against 342k cycles for 200k points), a launch with more clouds than 8-CTA clusters fit takes clusters of 4.
static int g_vf_max8 = -1, g_vf_max4 = -1, g_vf_force = 0;
static size_t g_vf_l2_budget = (size_t)80 << 20;

template <int CS> static int vf_query() {
  if (cudaFuncSetAttribute(vg_fast_kernel<CS>, cudaFuncAttributeMaxDynamicSharedMemorySize, VfGeo<CS>::SMEM_BYTES) != cudaSuccess) { cudaGetLastError(); return 0; }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(CS * 64); cfg.blockDim = dim3(VF_T); cfg.dynamicSmemBytes = VfGeo<CS>::SMEM_BYTES;
  cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = CS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  int nc = 0;
  if (cudaOccupancyMaxActiveClusters(&nc, vg_fast_kernel<CS>, &cfg) != cudaSuccess) { cudaGetLastError(); nc = 0; }
  return nc > 64 ? 64 : nc;
}

int vg_fast_init() {
  g_vf_force = 0;
  if (const char* e = getenv("FCCF_VF_CS")) g_vf_force = atoi(e);
  g_vf_max8 = vf_query<8>(); g_vf_max4 = vf_query<4>();
  { int dev = 0, l2 = 0; cudaGetDevice(&dev); if (cudaDeviceGetAttribute(&l2, cudaDevAttrL2CacheSize, dev) == cudaSuccess && l2 > 0) g_vf_l2_budget = (size_t)l2 / 10 * 8; }
  if (const char* e = getenv("FCCF_VF_L2_MB")) g_vf_l2_budget = (size_t)atoi(e) << 20;
  if (g_vf_max8 < 1 || g_vf_max4 < 1) { g_vf_max8 = g_vf_max4 = 0; }
  return vg_fast_max_clusters();
}
int vg_fast_max_clusters() { return g_vf_max8 > g_vf_max4 ? g_vf_max8 : g_vf_max4; }
int vg_fast_nmax() { return VF_NMAX; }

// VoxelGrid stage `stage` of both clouds of every lane by clusters; scratch: `ncl` x `stride` records + queue entries
cudaError_t launch_voxelgrid_fast(cudaStream_t s, const Batch& b, int stage, int ncloud, const VgFastScratch& sc, uint64_t* launches) {
  const int G = b.G;
  std::vector<VFArgs> As(G);
  for (int g = 0; g < G; g++) {
    const Work& w = b.w[g];
    VFArgs& A = As[g]; memset(&A, 0, sizeof A);
    PipeState* st = w.st;
    for (int c = 0; c < 2; c++) {
      const int cc = c < ncloud ? c : 0;
      const CloudWS& cw = w.c[cc];
      A.in[c] = (stage == 0) ? nullptr : cw.vg_xyz[0];
      A.n_in[c] = (stage == 0) ? &st->vg[0][cc].n_in : &st->vg[0][cc].n_out;
      A.st[c] = &st->vg[stage][cc];
      A.out[c] = cw.vg_xyz[stage]; A.cell[c] = cw.vg_cell[stage]; A.cnt[c] = cw.vg_cnt[stage];
    }
    A.call = &st->call; A.status = &st->status; A.emulate = b.p.emulate_pcl_overflow;
    A.prof = (stage == 0) ? st->prof + 22 : nullptr;
  }
  const VFArgs* dA = b.tab->put(As.data(), G);
  const int nitems = G * ncloud;
  // Clusters in flight are capped so that what they keep between their phases (per point 12 B of the cloud, the
  // 4-byte rank word and the 16-byte queue entry) stays inside L2: a scratch that spills turns every 16-byte
  // queue store into DRAM read-modify-write traffic (measured: 94 B of DRAM traffic per point with 33 clusters
  // of 200k-point clouds in flight, 3x slower scatter).  Stage 1 works on the stage-0 output (a fraction of
  // the raw cloud; the count is device-side, so the capacity / 4 stands in for it).
  const size_t per_cloud = (size_t)32 * (size_t)(stage == 0 ? sc.stride : (sc.stride + 3) / 4);
  int allowed = (int)(g_vf_l2_budget / (per_cloud ? per_cloud : 1));
  if (allowed < 2) allowed = 2;
  // more clouds than 8-CTA clusters fit: clusters of 4.  Capped by the L2 budget they take ~60 SMs and leave the rest to
  // the other launch sequences in flight (the single-CTA stages of the rotating groups): measured best for batches
  // (0.0330 ms/registration against 0.0357 with 15 clusters of 8 and 0.0344 with 33 uncapped clusters of 4).
  int cs = nitems > g_vf_max8 ? 4 : 8;
  if (g_vf_force == 4 || g_vf_force == 8) cs = g_vf_force;
  int ncl = cs == 8 ? g_vf_max8 : g_vf_max4;
  if (ncl > allowed) ncl = allowed;
  if (ncl > sc.ncl) ncl = sc.ncl;
  if (ncl > nitems) ncl = nitems;
  if (ncl < 1) ncl = 1;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(cs * ncl); cfg.blockDim = dim3(VF_T); cfg.dynamicSmemBytes = cs == 8 ? VfGeo<8>::SMEM_BYTES : VfGeo<4>::SMEM_BYTES; cfg.stream = s;
  cudaLaunchAttribute at[2]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = fccf_pdl_on() ? 2 : 1;
  cudaError_t e = cs == 8 ? cudaLaunchKernelEx(&cfg, vg_fast_kernel<8>, dA, ncloud, nitems, sc.info, sc.queue, sc.stride)
                          : cudaLaunchKernelEx(&cfg, vg_fast_kernel<4>, dA, ncloud, nitems, sc.info, sc.queue, sc.stride);
  if (launches) *launches += 1;
  return e;
}

}  // namespace fccf
