// sort.cu — stable LSD radix sort of (64-bit key, 32-bit value) pairs and segment-head
// extraction, with device-side element counts and key widths (no host round trip).
//
// Used for: the VoxelGrid cell sort (std::sort over cloud_point_index_idx, PCL 1.10
// voxel_grid.hpp, called at FCCF.cpp:1377-1387/1668-1678), the octree DFS (Morton) ordering of
// FCCF.cpp:475-484, and the x-ordering of hypothesis translations for the radius search of
// FCCF.cpp:1075-1091.  Stability (ascending original index inside a cell) is what fixes the
// float32 summation order of every centroid/covariance downstream, so it is part of parity.
//
// One pass = two kernels:  rs_hist (per-tile digit histograms; the last tile to finish turns
// them into global scatter offsets) and rs_scatter (stable ranking by warp match + per-(round,
// warp) digit counters, then a direct scatter).  Tiles are 2048 keys; grid.y batches up to
// three independent sort jobs (both clouds of a pair) and grid.z the lanes (registrations in
// flight) in one launch.
#include "fccf_internal.h"

namespace fccf {

// Passes that do work: ceil(nbits / 8) rounded up to an even count (so that the result lands where an even
// pass count leaves it) — keys of up to 16 bits (every indoor-scale voxel grid) need only the first two of
// the launched passes; the rest exit at once.  np (launched passes, even) must cover the widest key the
// caller can produce: 4 for 32-bit keys, 6 for the 34-bit cluster keys, 8 for 64-bit cell keys.
__device__ __forceinline__ int rs_active(int nbits, int np) {
  int na = (nbits + 7) / 8;
  if (na < 2) na = 2;
  na += na & 1;
  return ((np & 1) || na > np) ? np : na;
}
__device__ __forceinline__ int rs_bpp(int nbits, int np) {
  int na = rs_active(nbits, np);
  int b = (nbits + na - 1) / na;
  return b < 1 ? 1 : (b > 8 ? 8 : b);
}

// Jobs of up to RS_SMALL keys are sorted by ONE CTA in shared memory before the passes (bitonic network over
// (key, original index): the same stable order), the result written where an even pass count leaves it (back
// in kin / vin); the pass kernels of such a job exit at once.
// A 1k-hypothesis pool costs one ~8 us launch instead of twelve dependent pass kernels.
// (RS_SMALL = 8192 was measured: the one-CTA network takes longer than four passes over all SMs on 8k octree keys)
extern __shared__ __align__(16) unsigned char rs_small_dyn[];
template <typename KT> struct RsSmallBytes { static constexpr int value = RS_SMALL * (sizeof(KT) == 4 ? 8 : 10); };
template <typename KT>
__global__ void __launch_bounds__(1024) rs_small_kernel(const SortJobs* __restrict__ JB, int small_only) {
  FCCF_PDL_ENTER();
  const SortJob& j = JB[blockIdx.z].j[blockIdx.y];
  const int n = *j.n;
  const int t = threadIdx.x;
  if (n > RS_SMALL) { if (small_only && j.miss && t == 0) atomicOr(j.miss, ST_SORT_MISS); return; }
  KT* kin = (KT*)j.kin;
  u32* vin = const_cast<u32*>(j.vin);
  int N = 2;
  while (N < n) N <<= 1;
  // Bitonic network: exchanges at distance >= 32 go through shared memory (one barrier each), the up to five exchanges at
  // distances 16 .. 1 that end every merge step stay inside a warp and run on registers with shuffles (element i is
  // handled by thread i mod 1024, so partners at distance < 32 sit in the same warp) — 28 barriers instead of 66 for 2048 keys.
  if (sizeof(KT) == 4) {
    // 32-bit keys: one 64-bit word (key, index) per element, a single compare per exchange
    u64* sk = (u64*)rs_small_dyn;
    for (int i = t; i < N; i += 1024) sk[i] = i < n ? (((u64)kin[i] << 32) | (u64)(u32)i) : ~(u64)0;
    __syncthreads();
    for (int k = 2; k <= N; k <<= 1) {
      int jj = k >> 1;
      for (; jj >= 32 || (N < 64 && jj > 0); jj >>= 1) {
        for (int q = t; q < (N >> 1); q += 1024) {
          const int lo = ((q & ~(jj - 1)) << 1) | (q & (jj - 1)), hi = lo | jj;
          const bool up = (lo & k) == 0;
          const u64 a = sk[lo], b = sk[hi];
          if ((a > b) == up) { sk[lo] = b; sk[hi] = a; }
        }
        __syncthreads();
      }
      if (jj > 0) {
        for (int i = t; i < N; i += 1024) {
          u64 v = sk[i];
          const bool up = (i & k) == 0;
          for (int j2 = jj; j2 > 0; j2 >>= 1) {
            const u64 o = __shfl_xor_sync(0xffffffffu, v, j2);
            const bool keep_min = (((i & j2) == 0) == up);
            v = keep_min ? (v < o ? v : o) : (v > o ? v : o);
          }
          sk[i] = v;
        }
        __syncthreads();
      }
    }
    for (int i = t; i < n; i += 1024) { const u64 v = sk[i]; kin[i] = (KT)(v >> 32); vin[i] = (u32)v; }
  } else {
    KT* sk = (KT*)rs_small_dyn;
    unsigned short* sv = (unsigned short*)(sk + RS_SMALL);
    for (int i = t; i < N; i += 1024) { sk[i] = i < n ? kin[i] : (KT)~(KT)0; sv[i] = (unsigned short)i; }
    __syncthreads();
    for (int k = 2; k <= N; k <<= 1) {
      int jj = k >> 1;
      for (; jj >= 32 || (N < 64 && jj > 0); jj >>= 1) {
        for (int q = t; q < (N >> 1); q += 1024) {
          const int lo = ((q & ~(jj - 1)) << 1) | (q & (jj - 1)), hi = lo | jj;
          const bool up = (lo & k) == 0;
          const KT a = sk[lo], b = sk[hi];
          const unsigned short ia = sv[lo], ib = sv[hi];
          const bool gt = a > b || (a == b && ia > ib);
          if (gt == up) { sk[lo] = b; sk[hi] = a; sv[lo] = ib; sv[hi] = ia; }
        }
        __syncthreads();
      }
      if (jj > 0) {
        for (int i = t; i < N; i += 1024) {
          KT v = sk[i]; u32 vi = sv[i];
          const bool up = (i & k) == 0;
          for (int j2 = jj; j2 > 0; j2 >>= 1) {
            const KT o = __shfl_xor_sync(0xffffffffu, v, j2);
            const u32 oi = __shfl_xor_sync(0xffffffffu, vi, j2);
            const bool less = v < o || (v == o && vi < oi);
            const bool keep_min = (((i & j2) == 0) == up);
            if (keep_min != less) { v = o; vi = oi; }
          }
          sk[i] = v; sv[i] = (unsigned short)vi;
        }
        __syncthreads();
      }
    }
    for (int i = t; i < n; i += 1024) { kin[i] = sk[i]; vin[i] = (u32)sv[i]; }
  }
}

template <typename KT>
__global__ void __launch_bounds__(RS_T) rs_hist_kernel(const SortJobs* __restrict__ JB, int pass, int np) {
  FCCF_PDL_ENTER();
  const SortJob& j = JB[blockIdx.z].j[blockIdx.y];
  const KT* __restrict__ kin = (const KT*)j.kin;
  const int n = *j.n;
  const int nact = (n + RS_TILE - 1) / RS_TILE;
  if (n <= RS_SMALL || pass >= rs_active(*j.nbits, np)) return;   // small jobs are sorted by rs_small_kernel
  const int t = threadIdx.x;
  __shared__ u32 h[256];
  __shared__ int s_last;
  const int bpp = rs_bpp(*j.nbits, np), shift = pass * bpp;
  const u32 mask = (1u << bpp) - 1u;
  for (int tile = blockIdx.x; tile < nact; tile += gridDim.x) {
    h[t] = 0;
    __syncthreads();
    const int base = tile * RS_TILE;
#pragma unroll
    for (int r = 0; r < RS_I; r++) {
      int i = base + r * RS_T + t;
      if (i < n) atomicAdd(&h[(u32)(kin[i] >> shift) & mask], 1u);
    }
    __syncthreads();
    j.hist[(size_t)tile * 256 + t] = h[t];
    __threadfence();
    __syncthreads();
    if (t == 0) s_last = (atomicAdd(j.ticket, 1) == nact - 1);
    __syncthreads();
    if (!s_last) continue;
    __threadfence();
    // last tile: hist[b][d] -> exclusive prefix over tiles; row nact = exclusive prefix over digits
    u32 run = 0;
    for (int b0 = 0; b0 < nact; b0 += 16) {     // sixteen independent loads in flight per thread (10M-point clouds: ~5000 tiles)
      u32 v[16];
#pragma unroll
      for (int u = 0; u < 16; u++) v[u] = (b0 + u < nact) ? __ldcg(&j.hist[(size_t)(b0 + u) * 256 + t]) : 0u;
#pragma unroll
      for (int u = 0; u < 16; u++) if (b0 + u < nact) { j.hist[(size_t)(b0 + u) * 256 + t] = run; run += v[u]; }
    }
    h[t] = run;
    __syncthreads();
    if (t == 0) {
      u32 acc = 0;
      for (int d = 0; d < 256; d++) { u32 v = h[d]; h[d] = acc; acc += v; }
      *j.ticket = 0;
    }
    __syncthreads();
    j.hist[(size_t)nact * 256 + t] = h[t];
    __syncthreads();
  }
}

template <typename KT>
__global__ void __launch_bounds__(RS_T, 4) rs_scatter_kernel(const SortJobs* __restrict__ JB, int pass, int np, int identity) {
  FCCF_PDL_ENTER();
  const SortJob& j = JB[blockIdx.z].j[blockIdx.y];
  const KT* __restrict__ kin = (const KT*)j.kin; KT* __restrict__ kout = (KT*)j.kout;
  const int n = *j.n;
  const int nact = (n + RS_TILE - 1) / RS_TILE;
  if (n <= RS_SMALL || pass >= rs_active(*j.nbits, np)) return;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  // Warp w owns the 256 consecutive keys [w * 256, (w + 1) * 256) of the tile, in 8 rounds of 32: the stable
  // order of the tile is (warp, round, lane), so ONE private counter row per warp, bumped round after round
  // by the warp itself, gives every key its rank inside (warp, digit) without block barriers; the block
  // then only scans 8 rows per digit.  The tile is staged in digit order in shared memory so that the
  // global writes are coalesced runs.
  __shared__ u32 cntw[RS_T / 32][256];
  __shared__ __align__(16) unsigned char s_stage[RS_TILE * (sizeof(KT) + 4)];
  KT* skey = (KT*)s_stage; u32* sval = (u32*)(s_stage + RS_TILE * sizeof(KT));
  __shared__ u32 gbase[256];      // global offset of the digit, minus the digit's start inside the tile
  __shared__ u32 dstart[256];     // start of the digit inside the tile
  __shared__ u32 s_ws[RS_T / 32];
  const int bpp = rs_bpp(*j.nbits, np), shift = pass * bpp;
  const u32 mask = (1u << bpp) - 1u;
  const unsigned lt = (1u << lane) - 1u;
  for (int tile = blockIdx.x; tile < nact; tile += gridDim.x) {
#pragma unroll
    for (int w = 0; w < RS_T / 32; w++) cntw[w][t] = 0u;
    const int base = tile * RS_TILE + warp * (RS_I * 32);
    const u32 gb = j.hist[(size_t)tile * 256 + t] + j.hist[(size_t)nact * 256 + t];
    __syncthreads();
    KT key[RS_I]; u32 val[RS_I]; int dig[RS_I]; int off[RS_I];
    // all loads of the warp's 256 keys first (eight keys and values in flight per thread), then the ranking
#pragma unroll
    for (int r = 0; r < RS_I; r++) {
      const int i = base + r * 32 + lane;
      key[r] = (i < n) ? kin[i] : (KT)0;
      val[r] = (i < n) ? (identity ? (u32)i : j.vin[i]) : 0u;
    }
#pragma unroll
    for (int r = 0; r < RS_I; r++) {
      const int i = base + r * 32 + lane;
      const bool valid = i < n;
      const unsigned am = __ballot_sync(0xffffffffu, valid);
      dig[r] = -1; off[r] = 0;
      if (valid) {
        dig[r] = (int)((u32)(key[r] >> shift) & mask);
        // lanes of this round with the same digit: one ballot per digit bit
        unsigned peers = am;
        for (int bb = 0; bb < bpp; bb++) { const unsigned bal = __ballot_sync(am, (dig[r] >> bb) & 1); peers &= ((dig[r] >> bb) & 1) ? bal : ~bal; }
        const int leader = __ffs(peers) - 1;
        u32 old = 0;
        if (lane == leader) { old = cntw[warp][dig[r]]; cntw[warp][dig[r]] = old + (u32)__popc(peers); }
        old = __shfl_sync(peers, old, leader);
        off[r] = (int)old + __popc(peers & lt);
      }
      __syncwarp();
    }
    __syncthreads();
    // digit t: exclusive prefix over the warps' rows, then over the digits of the tile
    u32 run = 0;
#pragma unroll
    for (int w = 0; w < RS_T / 32; w++) { const u32 c = cntw[w][t]; cntw[w][t] = run; run += c; }
    u32 inc = run;
    for (int d = 1; d < 32; d <<= 1) { u32 y = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += y; }
    if (lane == 31) s_ws[warp] = inc;
    __syncthreads();
    u32 woff = 0;
    for (int w = 0; w < warp; w++) woff += s_ws[w];
    const u32 ds = woff + inc - run;
    dstart[t] = ds; gbase[t] = gb - ds;
    __syncthreads();
#pragma unroll
    for (int r = 0; r < RS_I; r++) if (dig[r] >= 0) { const int lp = (int)(dstart[dig[r]] + cntw[warp][dig[r]]) + off[r]; skey[lp] = key[r]; sval[lp] = val[r]; }
    __syncthreads();
    const int ntile = min(RS_TILE, n - tile * RS_TILE);
    for (int q = t; q < ntile; q += RS_T) {
      const KT k = skey[q];
      const u32 pos = gbase[(u32)(k >> shift) & mask] + (u32)q;
      kout[pos] = k;
      j.vout[pos] = sval[q];
    }
    __syncthreads();
  }
}

void sort_init_attributes() {
  cudaFuncSetAttribute(rs_small_kernel<u32>, cudaFuncAttributeMaxDynamicSharedMemorySize, RsSmallBytes<u32>::value);
  cudaFuncSetAttribute(rs_small_kernel<u64>, cudaFuncAttributeMaxDynamicSharedMemorySize, RsSmallBytes<u64>::value);
}
void launch_sort(cudaStream_t s, const SortJobs* ab, const SortJobs* ba, int njobs, int G, int cap, int np, int key_bytes, uint64_t* launches, bool small_only) {
  dim3 grid(grid_x((cap + RS_TILE - 1) / RS_TILE, G, njobs), njobs, G);
  if (key_bytes == 4) klaunch(rs_small_kernel<u32>, dim3(dim3(1, njobs, G)), dim3(1024), (size_t)RsSmallBytes<u32>::value, s, ab, small_only ? 1 : 0);
  else klaunch(rs_small_kernel<u64>, dim3(dim3(1, njobs, G)), dim3(1024), (size_t)RsSmallBytes<u64>::value, s, ab, small_only ? 1 : 0);
  if (launches) *launches += 1;
  if (small_only) return;
  for (int p = 0; p < np; p++) {
    const SortJobs* J = (p & 1) ? ba : ab;
    if (key_bytes == 4) {
      klaunch(rs_hist_kernel<u32>, dim3(grid), dim3(RS_T), 0, s, J, p, np);
      klaunch(rs_scatter_kernel<u32>, dim3(grid), dim3(RS_T), 0, s, J, p, np, p == 0 ? 1 : 0);
    } else {
      klaunch(rs_hist_kernel<u64>, dim3(grid), dim3(RS_T), 0, s, J, p, np);
      klaunch(rs_scatter_kernel<u64>, dim3(grid), dim3(RS_T), 0, s, J, p, np, p == 0 ? 1 : 0);
    }
    if (launches) *launches += 2;
  }
}

// ------------------------------------------------------------------------------------------
// segment heads of a sorted key array (one segment per distinct key), order preserving
// ------------------------------------------------------------------------------------------
template <typename KT>
__device__ __forceinline__ bool seg_is_head(const KT* keys, int i) { return i == 0 || keys[i] != keys[i - 1]; }

template <typename KT>
__global__ void __launch_bounds__(RS_T) seg_count_kernel(const SegJobs* __restrict__ JB) {
  FCCF_PDL_ENTER();
  const SegJob& j = JB[blockIdx.z].j[blockIdx.y];
  const KT* __restrict__ keys = (const KT*)j.keys;
  const int n = *j.n;
  const int nact = (n + RS_TILE - 1) / RS_TILE;
  const int t = threadIdx.x;
  if (nact == 0) { if (blockIdx.x == 0 && t == 0) { *j.nseg = 0; j.seg_start[0] = 0; } return; }
  __shared__ int s_cnt[RS_T / 32];
  __shared__ int s_last;
  __shared__ u32 sc[RS_T];
  __shared__ u32 carry;
  for (int tile = blockIdx.x; tile < nact; tile += gridDim.x) {
    const int base = tile * RS_TILE;
    int c = 0;
#pragma unroll
    for (int r = 0; r < RS_I; r++) { int i = base + r * RS_T + t; if (i < n && seg_is_head(keys, i)) c++; }
    for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((t & 31) == 0) s_cnt[t >> 5] = c;
    __syncthreads();
    if (t == 0) {
      int tot = 0;
      for (int w = 0; w < RS_T / 32; w++) tot += s_cnt[w];
      j.blk[tile] = (u32)tot;
      __threadfence();
      s_last = (atomicAdd(j.ticket, 1) == nact - 1);
    }
    __syncthreads();
    if (!s_last) continue;
    __threadfence();
    // exclusive scan over the tile counts (chunks of RS_T with a carry)
    if (t == 0) carry = 0;
    __syncthreads();
    for (int b0 = 0; b0 < nact; b0 += RS_T) {
      int b = b0 + t;
      u32 v = (b < nact) ? __ldcg(&j.blk[b]) : 0u;
      sc[t] = v;
      __syncthreads();
      for (int o = 1; o < RS_T; o <<= 1) { u32 a = (t >= o) ? sc[t - o] : 0u; __syncthreads(); sc[t] += a; __syncthreads(); }
      if (b < nact) j.blk[b] = carry + sc[t] - v;
      __syncthreads();
      if (t == RS_T - 1) carry += sc[t];
      __syncthreads();
    }
    if (t == 0) { *j.nseg = (int)carry; j.seg_start[carry] = n; *j.ticket = 0; }
    __syncthreads();
  }
}

template <typename KT>
__global__ void __launch_bounds__(RS_T) seg_write_kernel(const SegJobs* __restrict__ JB) {
  FCCF_PDL_ENTER();
  const SegJob& j = JB[blockIdx.z].j[blockIdx.y];
  const KT* __restrict__ keys = (const KT*)j.keys;
  const int n = *j.n;
  const int nact = (n + RS_TILE - 1) / RS_TILE;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  __shared__ int s_w[RS_T / 32];
  __shared__ int s_run;
  for (int tile = blockIdx.x; tile < nact; tile += gridDim.x) {
    if (t == 0) s_run = (int)j.blk[tile];
    __syncthreads();
    const int base = tile * RS_TILE;
    for (int r = 0; r < RS_I; r++) {
      int i = base + r * RS_T + t;
      bool head = (i < n) && seg_is_head(keys, i);
      unsigned b = __ballot_sync(0xffffffffu, head);
      if (lane == 0) s_w[warp] = __popc(b);
      __syncthreads();
      int off = s_run;
      for (int w = 0; w < warp; w++) off += s_w[w];
      if (head) j.seg_start[off + __popc(b & ((1u << lane) - 1u))] = i;
      __syncthreads();
      if (t == 0) { int tot = 0; for (int w = 0; w < RS_T / 32; w++) tot += s_w[w]; s_run += tot; }
      __syncthreads();
    }
  }
}

void launch_segments(cudaStream_t s, const SegJobs* jobs, int njobs, int G, int cap, int key_bytes, uint64_t* launches) {
  dim3 grid(grid_x((cap + RS_TILE - 1) / RS_TILE, G, njobs), njobs, G);
  if (key_bytes == 4) { klaunch(seg_count_kernel<u32>, dim3(grid), dim3(RS_T), 0, s, jobs); klaunch(seg_write_kernel<u32>, dim3(grid), dim3(RS_T), 0, s, jobs); }
  else { klaunch(seg_count_kernel<u64>, dim3(grid), dim3(RS_T), 0, s, jobs); klaunch(seg_write_kernel<u64>, dim3(grid), dim3(RS_T), 0, s, jobs); }
  if (launches) *launches += 2;
}

}  // namespace fccf
