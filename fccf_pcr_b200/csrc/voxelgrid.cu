// voxelgrid.cu — pcl::VoxelGrid<PointXYZ>::filter as called at FCCF.cpp:1668-1678 (main) and
// FCCF.cpp:1377-1387 (computer_transform_guess), restated for the GPU:
//   vg_minmax   bounding box of the finite points (getMinMax3D) + grid set-up by the last block
//               (inverse leaf, int32-overflow bail-out, min_b/div_b, key width)
//   vg_keys     one 64-bit cell key per point: ijk = (int)(floor(p*inv) - (float)min_b),
//               key = ix + dx*(iy + dy*iz)  (PCL's idx, widened to 64 bit); bail-out: key = index
//   radix sort  (sort.cu) by key, stable => ascending original index inside a cell
//   segments    one segment per occupied cell
//   vg_centroid per cell, float32 running sum in sorted order then / n (pcl::CentroidPoint)
// HBM traffic per launch is 12 B/point read (+8 B key write); everything else is key traffic.
#include "fccf_dev.cuh"
#include "fccf_internal.h"
#include <vector>

namespace fccf {

struct VGArgs {
  const CallArgs* call;      // leaf; raw clouds when in[c] == nullptr (stage 0)
  const float* in[2];        // packed xyz (nullptr: call->raw[c])
  const int* n_in[2];        // device-side input count
  VGState* st[2];
  u64* keys[2];
  int* ticket[2];
  int emulate;
};

__global__ void __launch_bounds__(256) vg_minmax_kernel(const VGArgs* __restrict__ AB) {
  FCCF_PDL_ENTER();
  const VGArgs& A = AB[blockIdx.z];
  const int c = blockIdx.y;
  const int n = *A.n_in[c];
  const float* p = A.in[c] ? A.in[c] : A.call->raw[c];
  int mn[3] = {0x7fffffff, 0x7fffffff, 0x7fffffff}, mx[3] = {(int)0x80000000, (int)0x80000000, (int)0x80000000};
  int nf = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    float x = p[3 * i], y = p[3 * i + 1], z = p[3 * i + 2];
    if (isfinite(x) && isfinite(y) && isfinite(z)) {
      nf++;
      int ox = f2ord(x), oy = f2ord(y), oz = f2ord(z);
      mn[0] = min(mn[0], ox); mn[1] = min(mn[1], oy); mn[2] = min(mn[2], oz);
      mx[0] = max(mx[0], ox); mx[1] = max(mx[1], oy); mx[2] = max(mx[2], oz);
    }
  }
  for (int o = 16; o; o >>= 1) {
    nf += __shfl_xor_sync(0xffffffffu, nf, o);
#pragma unroll
    for (int a = 0; a < 3; a++) {
      mn[a] = min(mn[a], __shfl_xor_sync(0xffffffffu, mn[a], o));
      mx[a] = max(mx[a], __shfl_xor_sync(0xffffffffu, mx[a], o));
    }
  }
  VGState* st = A.st[c];
  if ((threadIdx.x & 31) == 0 && nf > 0) {
    atomicAdd(&st->n_finite, nf);
#pragma unroll
    for (int a = 0; a < 3; a++) { atomicMin(&st->mn[a], mn[a]); atomicMax(&st->mx[a], mx[a]); }
  }
  __threadfence();
  __syncthreads();
  __shared__ int s_last;
  if (threadIdx.x == 0) s_last = (atomicAdd(A.ticket[c], 1) == (int)gridDim.x - 1);
  __syncthreads();
  if (!s_last || threadIdx.x != 0) return;
  __threadfence();
  // grid set-up (PCL 1.10 voxel_grid.hpp applyFilter)
  *A.ticket[c] = 0;
  st->n_in = n;
  int nfin = atomicAdd(&st->n_finite, 0);
  float inv = 1.0f / A.call->leaf;
  st->inv = inv;
  if (nfin == 0) { st->bail = 0; st->total = 0; st->nbits = 1; st->n_finite = 0; return; }
  float fmn[3], fmx[3];
  for (int a = 0; a < 3; a++) { fmn[a] = ord2f(atomicAdd(&st->mn[a], 0)); fmx[a] = ord2f(atomicAdd(&st->mx[a], 0)); }
  long long dx = (long long)((fmx[0] - fmn[0]) * inv) + 1;
  long long dy = (long long)((fmx[1] - fmn[1]) * inv) + 1;
  long long dz = (long long)((fmx[2] - fmn[2]) * inv) + 1;
  int bail = (A.emulate && (dx * dy * dz) > 2147483647LL) ? 1 : 0;
  st->bail = bail;
  long long total;
  if (bail) total = n;
  else {
    for (int a = 0; a < 3; a++) {
      int lo = (int)floorf(fmn[a] * inv), hi = (int)floorf(fmx[a] * inv);
      st->minb[a] = lo;
      st->div[a] = (long long)hi - lo + 1;
    }
    total = st->div[0] * st->div[1] * st->div[2];
  }
  st->total = total;
  int nbits = 64 - __clzll(total);   // non-finite points get key = total
  st->nbits = nbits < 1 ? 1 : nbits;
}

template <typename KT>
__global__ void __launch_bounds__(256) vg_keys_kernel(const VGArgs* __restrict__ AB) {
  FCCF_PDL_ENTER();
  const VGArgs& A = AB[blockIdx.z];
  const int c = blockIdx.y;
  const VGState* st = A.st[c];
  const int n = st->n_in;
  const float* p = A.in[c] ? A.in[c] : A.call->raw[c];
  const float inv = st->inv;
  const int bail = st->bail;
  const float mb0 = (float)st->minb[0], mb1 = (float)st->minb[1], mb2 = (float)st->minb[2];
  const long long d0 = st->div[0], d01 = st->div[0] * st->div[1];
  const u64 total = (u64)st->total;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    float x = p[3 * i], y = p[3 * i + 1], z = p[3 * i + 2];
    u64 key;
    if (!(isfinite(x) && isfinite(y) && isfinite(z))) key = total;
    else if (bail) key = (u64)i;
    else {
      int i0 = (int)(floorf(x * inv) - mb0);
      int i1 = (int)(floorf(y * inv) - mb1);
      int i2 = (int)(floorf(z * inv) - mb2);
      key = (u64)((long long)i0 + (long long)i1 * d0 + (long long)i2 * d01);
    }
    ((KT*)A.keys[c])[i] = (KT)key;
  }
}

struct VGOut {
  const CallArgs* call;
  const float* in[2];
  const u64* keys[2];
  const u32* idx[2];
  const int* seg_start[2];
  VGState* st[2];
  float* out[2];
  long long* cell[2];
  int* cnt[2];
  u32* big[2];               // cells of more than VG_BIG points, left to vg_centroid_big_kernel (count: st->pad)
};
#define VG_BIG 192

// one thread per occupied cell: in-order float32 running sum (pcl::CentroidPoint<PointXYZ>)
template <typename KT>
__global__ void __launch_bounds__(128) vg_centroid_kernel(const VGOut* __restrict__ AB) {
  FCCF_PDL_ENTER();
  const VGOut& A = AB[blockIdx.z];
  const int c = blockIdx.y;
  VGState* st = A.st[c];
  const int nseg = st->n_out;
  const float* p = A.in[c] ? A.in[c] : A.call->raw[c];
  const u32* idx = A.idx[c];
  for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < nseg; s += gridDim.x * blockDim.x) {
    const int b = A.seg_start[c][s], e = A.seg_start[c][s + 1];
    if (e - b > VG_BIG) { A.big[c][atomicAdd(&st->pad, 1)] = (u32)s; continue; }     // a warp sums it (below)
    float sx = 0.f, sy = 0.f, sz = 0.f;
    // the sum is sequential (in index order, like pcl::CentroidPoint) but the gathers are not: eight points
    // of the cell are fetched at once before they are added
    for (int k = b; k < e; k += 8) {
      u32 id[8]; float x[8], y[8], z[8];
#pragma unroll
      for (int u = 0; u < 8; u++) id[u] = (k + u < e) ? idx[k + u] : 0u;
#pragma unroll
      for (int u = 0; u < 8; u++) if (k + u < e) { x[u] = p[3 * (size_t)id[u]]; y[u] = p[3 * (size_t)id[u] + 1]; z[u] = p[3 * (size_t)id[u] + 2]; }
#pragma unroll
      for (int u = 0; u < 8; u++) if (k + u < e) { sx += x[u]; sy += y[u]; sz += z[u]; }
    }
    float fn = (float)(e - b);
    A.out[c][3 * s] = sx / fn; A.out[c][3 * s + 1] = sy / fn; A.out[c][3 * s + 2] = sz / fn;
    A.cell[c][s] = (long long)((const KT*)A.keys[c])[b];
    A.cnt[c][s] = e - b;
  }
}

// Cells of many points (a coarse leaf on a dense cloud: 10M points in a few hundred cells, BASELINE config 5) would keep
// ONE thread adding tens of thousands of gathered points.  Here a CTA takes such a cell: its 8 warps gather 8 x 32
// points of it per step into one half of a double-buffered shared-memory tile set while lanes 0..2 of warp 0 add the
// other half's x / y / z in order — one dependent float add per point, which is the floor pcl::CentroidPoint's
// sequential sum allows (a 32k-point cell: ~0.15 ms instead of the 7 ms of one warp with one gather in flight).
template <typename KT>
__global__ void __launch_bounds__(256) vg_centroid_big_kernel(const VGOut* __restrict__ AB) {
  FCCF_PDL_ENTER();
  const VGOut& A = AB[blockIdx.z];
  const int c = blockIdx.y;
  VGState* st = A.st[c];
  const int nbig = st->pad;
  if (nbig == 0) return;
  const float* p = A.in[c] ? A.in[c] : A.call->raw[c];
  const u32* idx = A.idx[c];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __shared__ float tile[2][8][96];
  for (int k = blockIdx.x; k < nbig; k += gridDim.x) {
    const int s = (int)A.big[c][k];
    const int b = A.seg_start[c][s], e = A.seg_start[c][s + 1];
    float acc = 0.f;
    // first step
    { const int q = b + warp * 32 + lane; if (q < e) { const float* r = p + 3 * (size_t)idx[q]; float* t = &tile[0][warp][3 * lane]; t[0] = r[0]; t[1] = r[1]; t[2] = r[2]; } }
    __syncthreads();
    int cur = 0;
    for (int k0 = b; k0 < e; k0 += 256) {
      // gather the next step's point of this thread (loads in flight while warp 0 adds)
      const int q = k0 + 256 + warp * 32 + lane;
      float r0 = 0.f, r1 = 0.f, r2 = 0.f;
      if (q < e) { const float* r = p + 3 * (size_t)idx[q]; r0 = r[0]; r1 = r[1]; r2 = r[2]; }
      if (warp == 0 && lane < 3) {
        const int m = min(256, e - k0);
        const float* t = &tile[cur][0][lane];
        int j = 0;
        for (; j + 8 <= m; j += 8) {
          float v[8];
#pragma unroll
          for (int u = 0; u < 8; u++) v[u] = t[3 * (j + u)];
#pragma unroll
          for (int u = 0; u < 8; u++) acc += v[u];
        }
        for (; j < m; j++) acc += t[3 * j];
      }
      float* nt = &tile[cur ^ 1][warp][3 * lane];
      nt[0] = r0; nt[1] = r1; nt[2] = r2;
      __syncthreads();
      cur ^= 1;
    }
    const float fn = (float)(e - b);
    if (warp == 0 && lane < 3) A.out[c][3 * (size_t)s + lane] = acc / fn;
    if (threadIdx.x == 0) { A.cell[c][s] = (long long)((const KT*)A.keys[c])[b]; A.cnt[c][s] = e - b; }
    __syncthreads();
  }
}

struct InitArgs { PipeState* st; };
__global__ void init_state_kernel(const InitArgs* __restrict__ AB, const CallArgs* __restrict__ calls) {
  FCCF_PDL_ENTER();
  PipeState* st = AB[blockIdx.x].st;
  int t = threadIdx.x;
  if (t == 0) {
    const CallArgs call = calls[blockIdx.x];
    st->call = call;
    const int n0 = call.n0, n1 = call.n1;
    for (int s = 0; s < 2; s++) for (int c = 0; c < 2; c++) {
      VGState& v = st->vg[s][c];
      v.n_in = 0; v.n_finite = 0; v.n_out = 0; v.bail = 0; v.total = 0; v.nbits = 1; v.pad = 0;
      for (int a = 0; a < 3; a++) { v.mn[a] = 0x7fffffff; v.mx[a] = (int)0x80000000; v.minb[a] = 0; v.div[a] = 1; }
    }
    st->vg[0][0].n_in = n0; st->vg[0][1].n_in = n1;
    st->status = 0;
    st->n_match = 0;
    for (int i = 0; i < 3; i++) { st->n_hyp[i] = 0; st->n_centre[i] = 0; st->n_top[i] = 0; st->cluster_num[i] = 0; }
  }
  if (t < 64) st->tickets[t] = 0;
}

void launch_init_state(cudaStream_t s, const Batch& b, const CallArgs* d_calls, uint64_t* launches) {
  std::vector<InitArgs> I(b.G);
  for (int g = 0; g < b.G; g++) I[g].st = b.w[g].st;
  klaunch(init_state_kernel, dim3(b.G), dim3(64), 0, s, b.tab->put(I.data(), b.G), d_calls);
  if (launches) *launches += 1;
}

// stage 0: raw -> vg_xyz[0]; stage 1: vg_xyz[0] -> vg_xyz[1]
void launch_voxelgrid(cudaStream_t s, const Batch& b, int stage, int ncloud, uint64_t* launches) {
  const int G = b.G;
  std::vector<VGArgs> As(G); std::vector<VGOut> Os(G); std::vector<SortJobs> abs_(G), bas_(G); std::vector<SegJobs> sjs(G);
  int cap = 1;
  for (int g = 0; g < G; g++) {
    const Work& w = b.w[g];
    VGArgs& A = As[g]; VGOut& O = Os[g]; SortJobs& ab = abs_[g]; SortJobs& ba = bas_[g]; SegJobs& sj = sjs[g];
    memset(&A, 0, sizeof A); memset(&O, 0, sizeof O); memset(&ab, 0, sizeof ab); memset(&ba, 0, sizeof ba); memset(&sj, 0, sizeof sj);
    PipeState* st = w.st;
    for (int c = 0; c < ncloud; c++) {
      const CloudWS& cw = w.c[c];
      A.in[c] = (stage == 0) ? nullptr : cw.vg_xyz[0];
      A.n_in[c] = (stage == 0) ? &st->vg[0][c].n_in : &st->vg[0][c].n_out;
      A.st[c] = &st->vg[stage][c];
      A.keys[c] = cw.keyA;
      A.ticket[c] = &st->tickets[0 + c];
      SortJob j; j.miss = nullptr; j.kin = cw.keyA; j.kout = cw.keyB; j.vin = cw.idxA; j.vout = cw.idxB; j.n = &st->vg[stage][c].n_in; j.nbits = &st->vg[stage][c].nbits;
      j.hist = cw.hist; j.ticket = &st->tickets[2 + c];
      ab.j[c] = j;
      SortJob k = j; k.kin = cw.keyB; k.kout = cw.keyA; k.vin = cw.idxB; k.vout = cw.idxA;
      ba.j[c] = k;
      SegJob sg; sg.keys = cw.keyA; sg.n = &st->vg[stage][c].n_finite; sg.seg_start = cw.seg_start; sg.nseg = &st->vg[stage][c].n_out; sg.blk = cw.segblk; sg.ticket = &st->tickets[4 + c];
      sj.j[c] = sg;
      O.in[c] = A.in[c]; O.keys[c] = cw.keyA; O.idx[c] = cw.idxA; O.seg_start[c] = cw.seg_start; O.st[c] = &st->vg[stage][c];
      O.out[c] = cw.vg_xyz[stage]; O.cell[c] = cw.vg_cell[stage]; O.cnt[c] = cw.vg_cnt[stage];
      O.big[c] = cw.hist;            // the sort's histogram buffer is free by then: (cap / 2048 + 2) * 256 entries >= cap / VG_BIG
      if (cw.cap > cap) cap = cw.cap;
    }
    for (int c = ncloud; c < 2; c++) { O.big[c] = O.big[0]; A.in[c] = A.in[0]; A.n_in[c] = A.n_in[0]; A.st[c] = A.st[0]; A.keys[c] = A.keys[0]; A.ticket[c] = A.ticket[0]; }
    A.call = &st->call; O.call = &st->call; A.emulate = b.p.emulate_pcl_overflow;
  }
  const VGArgs* dA = b.tab->put(As.data(), G); const VGOut* dO = b.tab->put(Os.data(), G);
  const SortJobs* dab = b.tab->put(abs_.data(), G); const SortJobs* dba = b.tab->put(bas_.data(), G); const SegJobs* dsj = b.tab->put(sjs.data(), G);
  int nb_mm = (cap + 256 * 8 - 1) / (256 * 8);
  if (nb_mm > 592) nb_mm = 592;
  klaunch(vg_minmax_kernel, dim3(dim3(grid_x(nb_mm, G, ncloud), ncloud, G)), dim3(256), 0, s, dA);
  // With pcl::VoxelGrid's int32 overflow bail-out emulated, every key (cell index, point index when bailing
  // out, "non-finite" = number of cells) fits 32 bits: the sort moves 4-byte keys.  Without it cells may
  // need the full 64 bits.
  const int kb = b.p.emulate_pcl_overflow ? 4 : 8;
  if (kb == 4) klaunch(vg_keys_kernel<u32>, dim3(dim3(grid_x((cap + 255) / 256, G, ncloud), ncloud, G)), dim3(256), 0, s, dA);
  else klaunch(vg_keys_kernel<u64>, dim3(dim3(grid_x((cap + 255) / 256, G, ncloud), ncloud, G)), dim3(256), 0, s, dA);
  if (launches) *launches += 2;
  launch_sort(s, dab, dba, ncloud, G, cap, kb == 4 ? 4 : 8, kb, launches);   // result back in keyA / idxA (even pass counts)
  launch_segments(s, dsj, ncloud, G, cap, kb, launches);
  // more CTAs of this kernel per lane: fewer lanes in flight at once, so that the clouds being gathered from stay in L2
  const dim3 gc(grid_x((cap + 127) / 128, G, ncloud, 16384), ncloud, G);
  if (kb == 4) klaunch(vg_centroid_kernel<u32>, dim3(gc), dim3(128), 0, s, dO);
  else klaunch(vg_centroid_kernel<u64>, dim3(gc), dim3(128), 0, s, dO);
  const dim3 gb(grid_x(148 * 2, G, ncloud, 4096), ncloud, G);
  if (kb == 4) klaunch(vg_centroid_big_kernel<u32>, dim3(gb), dim3(256), 0, s, dO);
  else klaunch(vg_centroid_big_kernel<u64>, dim3(gb), dim3(256), 0, s, dO);
  if (launches) *launches += 2;
}

}  // namespace fccf
