// api.cu — the C-ABI of libfccf (include/fccf.h): context, device workspace, the registration
// pipeline (main() + computer_transform_guess, FCCF.cpp:1646-1690 / 1370-1608) as a sequence of
// kernel launches on one stream with all sizes kept on the device, stage entry points for the
// parity tests, and read-back of stage intermediates ("debug blobs").
// No CPU fallback: every entry point needs a CUDA device.
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <string>
#include <vector>
#include <climits>
#include <cstddef>
#include <memory>
#include <thread>
#include "fccf_internal.h"
#include "hostcopy.h"

using namespace fccf;

namespace fccf {
bool nccl_allreduce_max_i64(const std::vector<int>& devs, const std::vector<cudaStream_t>& streams, const std::vector<const long long*>& words_dev, long long* out);
}

#define CK(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess) {                                                                       \
      char b_[512]; snprintf(b_, sizeof b_, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
      ctx->err = b_;                                                                               \
      return FCCF_ERR_CUDA;                                                                        \
    }                                                                                              \
  } while (0)

struct Arena {
  char* base = nullptr; size_t size = 0, off = 0;
  template <class T> T* take(size_t n) {
    size_t bytes = (n * sizeof(T) + 255) & ~(size_t)255;
    T* p = (T*)(base + off); off += bytes; return p;
  }
};

// One registration in flight: its workspaces and its slot in the group's state arrays.
struct Lane {
  int cap_pts = 0;
  Arena cloud_arena[2];
  float* d_raw[2] = {nullptr, nullptr};
  CloudWS c[2];
  PipeState* d_st = nullptr;     // slot of Group::d_st_all
  PipeState* h_st = nullptr;     // slot of Group::h_st_all (pinned)
  Arena hyp_arena;
  HypWS h;
  int pair = -1;
};

struct GraphRec { int G = 0; bool fast = false, lean = false; cudaGraphExec_t exec = nullptr; int launches = 0; };

// A group = the lanes that run as ONE batched launch sequence on one stream: every kernel of the
// pipeline is launched once with grid.z = G and serves G independent registrations (most stages of one
// registration are single-CTA, order-dependent greedy loops, so G of them fill the other SMs), and the
// whole sequence is one CUDA graph per G.  A context owns group 0 (single-pair and stage entry points
// use its lane 0) and up to three more groups for fccf_register_batch*, which rotates through them:
// while some groups compute, the next one receives its clouds (FCCF_GROUPS overrides the count).
struct Group {
  cudaStream_t stream = nullptr;
  cudaStream_t stream2 = nullptr;   // second capture stream: the graph branch of the fine-verify table build
  cudaEvent_t fork = nullptr, join = nullptr, fork2 = nullptr, join2 = nullptr;
  std::vector<Lane> lanes;
  PipeState* d_st_all = nullptr; PipeState* h_st_all = nullptr;
  CallArgs* d_calls = nullptr; CallArgs* h_calls = nullptr;
  ArgTable tab;
  std::vector<GraphRec> graphs;
  uint64_t params_epoch = 0;
  cudaEvent_t ev[6];
  cudaEvent_t sev[8];
  bool busy = false, had_h2d = false;
  int G = 0;                     // lanes of the sequence in flight
  bool fast = false;             // the sequence in flight uses the cluster VoxelGrid (voxelgrid_fast.cu)
  bool stage_timed = false;      // the sequence in flight records the per-stage events
  bool lean = false;             // the sequence in flight has no radix pass launches behind its hypothesis / fine-verify sorts
  float leaf = 0.f;
  VgFastScratch vf;              // its per-cluster scratch (L2-resident)
  size_t last_h2d = 0;
  uint64_t launches = 0, l0 = 0;
};

struct fccf_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;   // = groups[0]->stream
  fccf_params p;
  std::string err;
  uint64_t launches = 0;           // kernels launched outside the groups (stand-alone stage entry points)
  std::vector<Group*> groups;
  int max_lanes = 64;
  bool use_graph = true;
  // Cluster VoxelGrid: tried first; a cloud it cannot hold (ST_VG_FAST_MISS) makes the host re-run the sequence
  // through the generic kernels, and calls with a leaf <= the one that missed go generic directly for a while.
  bool vg_fast = false;
  float fast_miss_leaf = 0.f; int fast_miss_ttl = 0;
  // Sequences without the radix-pass launches of the hypothesis / fine-verify sorts ("lean"): used once a registration
  // has shown lists well inside the one-CTA sort, dropped for a while when one meets a longer list (ST_SORT_MISS: re-run)
  bool lean_ok = false; int lean_ttl = 0;
  uint64_t fast_runs = 0, fast_misses = 0;
  uint64_t params_epoch = 1;
  bool stage_timing = false;     // per-stage timing events inside the captured sequences (fccf_set_stage_timing)
  int cap_hyp = 1 << 18;
  bool have_run = false;
  float leaf = 0.f;
  ArgTable itab;                   // immediate-mode argument table of the stand-alone stage entry points
  // stand-alone scoring
  float *d_sc_s1 = nullptr, *d_sc_s2 = nullptr, *d_sc_T = nullptr, *d_sc_scores = nullptr;
  size_t sc_cap1 = 0, sc_cap2 = 0, sc_capT = 0;
  Arena sc_arena; ScoreWS sc_ws; int sc_cap_hash = 0; ScoreState* d_sc_ss = nullptr; int* d_sc_n = nullptr;
  size_t sc_n2 = 0, sc_nhyp = 0;
  long long* d_sc_best = nullptr;
  std::unique_ptr<CopyPool> pool;   // pageable host inputs: pinned staging chunks filled by worker threads (hostcopy.h), created on first use
  uint64_t staged_bytes = 0;
  Group& G0() { return *groups[0]; }
  Lane& L0() { return groups[0]->lanes[0]; }
};

bool& fccf::fccf_pdl_flag() { static thread_local bool on = false; return on; }
// programmatic dependent launches: for captured sequences of few lanes (latency); FCCF_PDL=0 / 1 forces them off / on
static bool pdl_wanted(int G) {
  static int mode = -2;
  if (mode == -2) { const char* e = getenv("FCCF_PDL"); mode = e ? atoi(e) : -1; }
  if (mode >= 0) return mode != 0;
  return false;   // measured on the 200k pair: 1.092 ms with, 1.085 ms without (the graph already keeps the gaps short)
}

static size_t cloud_bytes(int cap) {
  size_t c = (size_t)cap, nb = c / RS_TILE + 2;
  size_t b = 0;
  auto add = [&](size_t x) { b += (x + 255) & ~(size_t)255; };
  add(8 * c); add(8 * c); add(4 * c); add(4 * c); add(nb * 256 * 4); add(nb * 4);
  for (int s = 0; s < 2; s++) { add(12 * c); add(8 * c); add(4 * c); }
  add(4 * (c + 1)); add(4 * (c + 1)); add(48 * c); add(4 * c); add(32 * c); add(12 * c);
  for (int k = 0; k < 9; k++) add(4 * (c + 1));
  add(64 * c); add(4 * c); add(4 * 64);
  return b + 4096;
}
static void cloud_carve(Arena& a, CloudWS& w, int cap) {
  size_t c = (size_t)cap, nb = c / RS_TILE + 2;
  w.cap = cap;
  w.keyA = a.take<u64>(c); w.keyB = a.take<u64>(c); w.idxA = a.take<u32>(c); w.idxB = a.take<u32>(c);
  w.hist = a.take<u32>(nb * 256); w.segblk = a.take<u32>(nb);
  for (int s = 0; s < 2; s++) { w.vg_xyz[s] = a.take<float>(3 * c); w.vg_cell[s] = a.take<long long>(c); w.vg_cnt[s] = a.take<int>(c); }
  w.seg_start = a.take<int>(c + 1); w.vox_start = a.take<int>(c + 1);
  w.vox_rec = a.take<float>(12 * c); w.vox_aux = a.take<int>(c); w.pvox = a.take<float>(8 * c); w.sub = a.take<float>(3 * c);
  w.grow_label = a.take<int>(c + 1); w.merge_label = a.take<int>(c + 1); w.next = a.take<int>(c + 1);
  w.fhead = a.take<int>(c + 1); w.ftail = a.take<int>(c + 1); w.fnvox = a.take<int>(c + 1); w.falloc = a.take<int>(c + 1);
  w.fperm = a.take<int>(c + 1); w.fkey = a.take<int>(c + 1);
  w.fstat = a.take<float>(16 * c); w.face_vox = a.take<int>(c); w.face_off = a.take<int>(64);
}

static int table_create(fccf_ctx* ctx, ArgTable& t, size_t bytes, cudaStream_t s) {
  t.cap = bytes; t.off = 0; t.stream = s;
  if (cudaMalloc(&t.d, bytes) != cudaSuccess || cudaMallocHost(&t.h, bytes) != cudaSuccess) { ctx->err = "argument table allocation failed"; return FCCF_ERR_CUDA; }
  return FCCF_OK;
}
static void table_destroy(ArgTable& t) {
  if (t.d) cudaFree(t.d);
  if (t.h) cudaFreeHost(t.h);
  t.d = t.h = nullptr;
}

static void group_drop_graphs(Group* g) {
  for (GraphRec& r : g->graphs) if (r.exec) cudaGraphExecDestroy(r.exec);
  g->graphs.clear();
  g->tab.off = 0; g->tab.overflow = false;
}

static int group_create(fccf_ctx* ctx, Group** out) {
  Group* g = new Group();
  const int nl = ctx->max_lanes;
  bool ok = cudaStreamCreateWithFlags(&g->stream, cudaStreamNonBlocking) == cudaSuccess;
  ok = ok && cudaStreamCreateWithFlags(&g->stream2, cudaStreamNonBlocking) == cudaSuccess;
  ok = ok && cudaEventCreateWithFlags(&g->fork, cudaEventDisableTiming) == cudaSuccess && cudaEventCreateWithFlags(&g->join, cudaEventDisableTiming) == cudaSuccess;
  ok = ok && cudaEventCreateWithFlags(&g->fork2, cudaEventDisableTiming) == cudaSuccess && cudaEventCreateWithFlags(&g->join2, cudaEventDisableTiming) == cudaSuccess;
  ok = ok && cudaMalloc(&g->d_st_all, sizeof(PipeState) * nl) == cudaSuccess && cudaMallocHost(&g->h_st_all, sizeof(PipeState) * nl) == cudaSuccess;
  ok = ok && cudaMalloc(&g->d_calls, sizeof(CallArgs) * nl) == cudaSuccess && cudaMallocHost(&g->h_calls, sizeof(CallArgs) * nl) == cudaSuccess;
  if (!ok) { delete g; ctx->err = "state allocation failed"; cudaGetLastError(); return FCCF_ERR_CUDA; }
  cudaMemset(g->d_st_all, 0, sizeof(PipeState) * nl);
  memset(g->h_st_all, 0, sizeof(PipeState) * nl);
  // one capture of G lanes takes about 10 KB of argument blocks per lane (+ alignment)
  if (table_create(ctx, g->tab, (size_t)nl * 3 * 12288 + (1 << 20), g->stream) != FCCF_OK) { delete g; return FCCF_ERR_CUDA; }
  g->lanes.resize(nl);
  for (int l = 0; l < nl; l++) { g->lanes[l].d_st = g->d_st_all + l; g->lanes[l].h_st = g->h_st_all + l; }
  for (int i = 0; i < 6; i++) cudaEventCreate(&g->ev[i]);
  for (int i = 0; i < 8; i++) cudaEventCreate(&g->sev[i]);
  *out = g;
  return FCCF_OK;
}
static void group_destroy(Group* g) {
  if (!g) return;
  if (g->stream) cudaStreamSynchronize(g->stream);
  group_drop_graphs(g);
  for (Lane& L : g->lanes) {
    for (int c = 0; c < 2; c++) { if (L.cloud_arena[c].base) cudaFree(L.cloud_arena[c].base); if (L.d_raw[c]) cudaFree(L.d_raw[c]); }
    if (L.hyp_arena.base) cudaFree(L.hyp_arena.base);
  }
  if (g->vf.info) cudaFree(g->vf.info);
  if (g->vf.queue) cudaFree(g->vf.queue);
  if (g->d_st_all) cudaFree(g->d_st_all);
  if (g->h_st_all) cudaFreeHost(g->h_st_all);
  if (g->d_calls) cudaFree(g->d_calls);
  if (g->h_calls) cudaFreeHost(g->h_calls);
  table_destroy(g->tab);
  for (int i = 0; i < 6; i++) cudaEventDestroy(g->ev[i]);
  for (int i = 0; i < 8; i++) cudaEventDestroy(g->sev[i]);
  if (g->fork) cudaEventDestroy(g->fork);
  if (g->join) cudaEventDestroy(g->join);
  if (g->fork2) cudaEventDestroy(g->fork2);
  if (g->join2) cudaEventDestroy(g->join2);
  if (g->stream2) cudaStreamDestroy(g->stream2);
  if (g->stream) cudaStreamDestroy(g->stream);
  delete g;
}

static void lane_release(Lane* L) {
  for (int c = 0; c < 2; c++) {
    if (L->cloud_arena[c].base) cudaFree(L->cloud_arena[c].base);
    if (L->d_raw[c]) cudaFree(L->d_raw[c]);
    L->cloud_arena[c] = Arena(); L->d_raw[c] = nullptr;
    memset(&L->c[c], 0, sizeof L->c[c]);
  }
  if (L->hyp_arena.base) cudaFree(L->hyp_arena.base);
  L->hyp_arena = Arena();
  memset(&L->h, 0, sizeof L->h);
  L->cap_pts = 0;
}

// workspaces of lane `li` of group `g` for clouds of up to max(n0, n1) points.  On failure the lane owns
// nothing and reports capacity 0 (no pointer to freed memory survives), and the group's graphs are dropped.
static int ensure_capacity(fccf_ctx* ctx, Group* g, int li, size_t n0, size_t n1) {
  Lane* L = &g->lanes[li];
  size_t need = std::max(n0, n1);
  if (need < 1024) need = 1024;
  if ((size_t)L->cap_pts >= need) return FCCF_OK;
  if (need > (size_t)1 << 30) { ctx->err = "cloud too large"; return FCCF_ERR_ARG; }
  int cap = (int)((need + 4095) & ~(size_t)4095);
  CK(cudaStreamSynchronize(g->stream));
  group_drop_graphs(g);            // captured graphs hold the old pointers
  lane_release(L);
  for (int c = 0; c < 2; c++) {
    size_t bytes = cloud_bytes(cap);
    if (cudaMalloc(&L->cloud_arena[c].base, bytes) != cudaSuccess || cudaMalloc(&L->d_raw[c], (size_t)cap * 12) != cudaSuccess) {
      cudaGetLastError(); lane_release(L); ctx->err = "out of device memory (cloud workspace)"; return FCCF_ERR_CUDA;
    }
    L->cloud_arena[c].size = bytes;
    cloud_carve(L->cloud_arena[c], L->c[c], cap);
    L->c[c].raw = L->d_raw[c];
  }
  // fine-verify hash sized for the leftover cloud (<= cap points)
  HypWS& h = L->h;
  size_t ch = (size_t)ctx->cap_hyp, nbh = ch / RS_TILE + 2;
  size_t bytes = 0;
  auto add = [&](size_t x) { bytes += (x + 255) & ~(size_t)255; };
  add(4 * FCCF_MAXMATCH); add(4 * FCCF_MAXMATCH); add(48 * ch); add(32 * ch); add(16 * ch); add(8 * ch);
  add(8 * ch); add(8 * ch); add(4 * ch); add(4 * ch); add(nbh * 1024);
  for (int k = 0; k < 5; k++) add(4 * ch);
  add(8 * ch); add(8 * ch);
  add(4 * (size_t)cluster_nbl_ints()); add(4 * (size_t)cluster_deg_ints());
  size_t nc = 3 * FCCF_MAXCENTRE, ntop = 3 * FCCF_TOPK;
  add(nc * 32); add(nc * 64); add(nc * 4); add(nc * 4); add(nc * 128); add(nc * 4); add(nc * 4);
  add(ntop * 64); add(ntop * 4); add(ntop * 4); add(ntop * 4);
  size_t fv_bytes = score_ws_layout(nullptr, nullptr, cap, 48);   // 48 counter rows: hypotheses in flight of the global-table kernel
  add(fv_bytes);
  bytes += 8192;
  if (cudaMalloc(&L->hyp_arena.base, bytes) != cudaSuccess) { cudaGetLastError(); lane_release(L); ctx->err = "out of device memory (hypothesis workspace)"; return FCCF_ERR_CUDA; }
  L->hyp_arena.size = bytes;
  Arena& a = L->hyp_arena;
  h.cap_hyp = ctx->cap_hyp;
  h.match_cnt = a.take<int>(FCCF_MAXMATCH); h.match_off = a.take<int>(FCCF_MAXMATCH);
  h.hyp_T = a.take<float>(12 * ch); h.hyp_qt = a.take<float>(8 * ch); h.hyp_ax = a.take<float>(4 * ch); h.hyp_an = a.take<double>(ch);
  h.ckeyA = a.take<u64>(ch); h.ckeyB = a.take<u64>(ch); h.cidxA = a.take<u32>(ch); h.cidxB = a.take<u32>(ch); h.chist = a.take<u32>(nbh * 256);
  h.c_state = a.take<int>(ch); h.c_size = a.take<int>(ch); h.c_seeds = a.take<int>(ch); h.c_perm = a.take<int>(ch); h.c_key = a.take<int>(ch);
  h.c_members = a.take<int>(2 * ch); h.c_mdist = a.take<float>(2 * ch);
  h.c_nbl = a.take<int>((size_t)cluster_nbl_ints()); h.c_deg = a.take<int>((size_t)cluster_deg_ints());
  h.centre = a.take<float>(nc * 8); h.qv_T = a.take<float>(nc * 16); h.qv_score = a.take<float>(nc); h.qv_npair = a.take<int>(nc);
  h.qv_pairs = a.take<int>(nc * 32); h.qv_iters = a.take<int>(nc); h.rank_perm = a.take<int>(nc);
  h.top_T = a.take<float>(ntop * 16); h.top_s1 = a.take<float>(ntop); h.top_s2 = a.take<float>(ntop); h.top_centre = a.take<int>(ntop);
  score_ws_layout(&h.fv, a.take<char>(fv_bytes), cap, 48);
  L->cap_pts = cap;
  return FCCF_OK;
}

// per-cluster scratch of the cluster VoxelGrid for clouds of up to n points (shared by the lanes of the group)
static int ensure_fast_scratch(fccf_ctx* ctx, Group* g, size_t n) {
  if (!ctx->vg_fast) return FCCF_OK;
  size_t want = std::min<size_t>((n + 4095) & ~(size_t)4095, (size_t)((vg_fast_nmax() + 4095) & ~4095));
  if (want < 4096) want = 4096;
  if ((size_t)g->vf.stride >= want) return FCCF_OK;
  CK(cudaStreamSynchronize(g->stream));
  group_drop_graphs(g);
  if (g->vf.info) cudaFree(g->vf.info);
  if (g->vf.queue) cudaFree(g->vf.queue);
  g->vf = VgFastScratch();
  int ncl = std::min(vg_fast_max_clusters(), 2 * ctx->max_lanes);
  if (ncl < 1) return FCCF_OK;
  if (cudaMalloc(&g->vf.info, (size_t)ncl * want * 4) != cudaSuccess || cudaMalloc(&g->vf.queue, (size_t)ncl * want * 16) != cudaSuccess) {
    cudaGetLastError();
    if (g->vf.info) cudaFree(g->vf.info);
    g->vf = VgFastScratch();
    ctx->err = "out of device memory (cluster VoxelGrid scratch)"; return FCCF_ERR_CUDA;
  }
  g->vf.stride = (int)want; g->vf.ncl = ncl;
  return FCCF_OK;
}

extern "C" {

void fccf_default_params(fccf_params* p) {
  memset(p, 0, sizeof *p);
  p->parameter_l1 = 0.5f; p->parameter_l2 = 1.0f; p->parameter_k1 = 5.0f; p->parameter_k2 = 2.0f;
  p->normal_vector_threshold1 = 5.0f; p->normal_vector_threshold2 = 8.0f;
  p->face_voxel_size = 1.0f; p->voxel_point_threshold = 5; p->curvature_threshold = 0.05f; p->select_plane_number = 15;
  p->quick_verify_angel_threshold = 10.0f; p->quick_verify_distance_threshold = 2.0f; p->required_optimize_plane = 4.0f;
  p->fine_verify_voxel_size = 0.5f; p->fine_verify_number = 4;
  p->included_angle_same_threshold = 5.0f; p->included_angle_min_threshold = 30.0f; p->included_angle_max_threshold = 150.0f;
  p->third_plane_threshold = 0.5f; p->third_plane_normal_threshold = 5.0f;
  p->cluster_number_threshold = 10; p->cluster_angel_threshold = 2.0f; p->cluster_distance_threshold = 0.8f;
  p->seclct_cluster_number = 200; p->rough_threshold_gl = 2;
  p->emulate_pcl_overflow = 1;
}

// Values the kernels cannot run with (fixed-size tables, divisions): rejected up front with FCCF_ERR_ARG.
static const char* params_problem(const fccf_params& p) {
  if (!(p.face_voxel_size > 0.f) || !(p.fine_verify_voxel_size > 0.f)) return "face_voxel_size and fine_verify_voxel_size must be > 0";
  if (!(p.cluster_number_threshold >= 0.f) || p.cluster_number_threshold > (float)FCCF_MAXCENTRE) return "cluster_number_threshold must be in [0, 256] (cluster centre capacity per roughness type)";
  if (!(p.select_plane_number >= 0.f) || p.select_plane_number > (float)(FCCF_MAXF - 1)) return "select_plane_number must be in [0, 15] (the reference keeps select_plane_number + 1 <= 16 planes)";
  if (!(p.fine_verify_number >= 0.f)) return "fine_verify_number must be >= 0";
  if (!(p.seclct_cluster_number >= 0.f)) return "seclct_cluster_number must be >= 0";
  if (!(p.cluster_distance_threshold >= 0.f)) return "cluster_distance_threshold must be >= 0";
  if (p.batch_lanes < 0) return "batch_lanes must be >= 0";
  if (p.emulate_pcl_overflow != 0 && p.emulate_pcl_overflow != 1) return "emulate_pcl_overflow must be 0 or 1";
  return nullptr;
}
static thread_local std::string g_create_error;

fccf_ctx* fccf_create(int device, const fccf_params* params) {
  g_create_error.clear();
  // One hardware work queue per in-flight launch sequence: the driver default of 8 connections serialises the
  // rotating groups of fccf_register_batch.  Read when the process creates its CUDA context, so it only takes
  // effect if no CUDA call came first; an explicit setting of the caller is kept.
  setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0);
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev) { g_create_error = "no usable CUDA device (libfccf has no CPU path)"; return nullptr; }
  if (cudaSetDevice(device) != cudaSuccess) { g_create_error = "cudaSetDevice failed"; return nullptr; }
  fccf_params p0;
  if (params) p0 = *params; else fccf_default_params(&p0);
  if (const char* why = params_problem(p0)) { g_create_error = std::string("bad parameter: ") + why; return nullptr; }
  fccf_ctx* ctx = new fccf_ctx();
  ctx->device = device;
  ctx->p = p0;
  if (ctx->p.batch_lanes > 0) ctx->max_lanes = ctx->p.batch_lanes > 1024 ? 1024 : ctx->p.batch_lanes;
  if (const char* e = getenv("FCCF_NO_GRAPH")) ctx->use_graph = !(e[0] == '1');
  if (const char* e = getenv("FCCF_STAGE_EVENTS")) ctx->stage_timing = (e[0] == '1');
  sort_init_attributes();
  planes_init_attributes();
  score_init_attributes();
  cluster_init_attributes();
  ctx->vg_fast = vg_fast_init() > 0;
  if (const char* e = getenv("FCCF_NO_VG_FAST")) if (e[0] == '1') ctx->vg_fast = false;
  Group* g = nullptr;
  if (group_create(ctx, &g) != FCCF_OK) { delete ctx; return nullptr; }
  ctx->groups.push_back(g);
  ctx->stream = g->stream;
  if (table_create(ctx, ctx->itab, 1 << 20, ctx->stream) != FCCF_OK) { group_destroy(g); delete ctx; return nullptr; }
  return ctx;
}

void fccf_destroy(fccf_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  for (Group* g : ctx->groups) if (g && g->stream) cudaStreamSynchronize(g->stream);
  ctx->pool.reset();
  for (Group* g : ctx->groups) group_destroy(g);
  table_destroy(ctx->itab);
  if (ctx->sc_arena.base) cudaFree(ctx->sc_arena.base);
  if (ctx->d_sc_s1) cudaFree(ctx->d_sc_s1);
  if (ctx->d_sc_s2) cudaFree(ctx->d_sc_s2);
  if (ctx->d_sc_T) cudaFree(ctx->d_sc_T);
  if (ctx->d_sc_scores) cudaFree(ctx->d_sc_scores);
  if (ctx->d_sc_ss) cudaFree(ctx->d_sc_ss);
  if (ctx->d_sc_n) cudaFree(ctx->d_sc_n);
  if (ctx->d_sc_best) cudaFree(ctx->d_sc_best);
  delete ctx;
}

const char* fccf_last_error(const fccf_ctx* ctx) {
  if (ctx) return ctx->err.c_str();
  return g_create_error.empty() ? "no context (no usable CUDA device)" : g_create_error.c_str();   // why the last fccf_create of this thread failed
}
int fccf_set_params(fccf_ctx* ctx, const fccf_params* params) {
  if (!ctx || !params) return FCCF_ERR_ARG;
  if (const char* why = params_problem(*params)) { ctx->err = std::string("bad parameter: ") + why; return FCCF_ERR_ARG; }
  ctx->fast_miss_leaf = 0.f; ctx->fast_miss_ttl = 0; ctx->lean_ok = false; ctx->lean_ttl = 0;
  int lanes = ctx->p.batch_lanes;
  ctx->p = *params; ctx->p.batch_lanes = lanes;   // the lane count is fixed at creation
  ctx->params_epoch++;                            // captured graphs hold the old values: re-capture lazily
  return FCCF_OK;
}
int fccf_set_stage_timing(fccf_ctx* ctx, int on) {
  if (!ctx) return FCCF_ERR_ARG;
  if (ctx->stage_timing != (on != 0)) { ctx->stage_timing = (on != 0); ctx->params_epoch++; }   // the events are nodes of the captured sequences
  return FCCF_OK;
}
// library-level integers for tools (not part of include/fccf.h)
int fccf_debug_int(const char* name) { if (name && !strcmp(name, "vg_fast_max_clusters")) return vg_fast_max_clusters(); return -1; }
uint64_t fccf_launch_count(const fccf_ctx* ctx) {
  if (!ctx) return 0;
  uint64_t n = ctx->launches;
  for (const Group* g : ctx->groups) n += g->launches;
  return n;
}
void* fccf_stream_handle(const fccf_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

}  // extern "C"

namespace fccf {
// theta(c) = (float)(acos((double)c) * 180 / pi) as in compute_normal_angel (FCCF.cpp:375-376), monotone
// non-increasing in c: bisection over the ordered float bit patterns of [-1, 1]
float angle_cut(float thr_deg, bool strict) {
  auto theta = [](float c) { return (float)(acos((double)c) * 180 / 3.14159265358979323846); };
  auto pass = [&](float c) { float th = theta(c); return strict ? (th < thr_deg) : (th <= thr_deg); };
  auto ord = [](float f) { int32_t i; memcpy(&i, &f, 4); return i >= 0 ? (int64_t)i : (int64_t)(i ^ 0x7fffffff); };
  auto unord = [](int64_t o) { int32_t i = (int32_t)o; if (i < 0) i ^= 0x7fffffff; float f; memcpy(&f, &i, 4); return f; };
  if (!pass(1.0f)) return 2.0f;          // nothing passes
  int64_t lo = ord(-1.0f), hi = ord(1.0f);   // invariant: pass(unord(hi))
  if (pass(-1.0f)) return -1.0f;
  while (hi - lo > 1) { int64_t mid = lo + (hi - lo) / 2; if (pass(unord(mid))) hi = mid; else lo = mid; }
  return unord(hi);
}
AngleCuts make_angle_cuts(const fccf_params& p) {
  AngleCuts c;
  c.third_lt = angle_cut(p.third_plane_normal_threshold, true);
  c.qv_lt = angle_cut(p.quick_verify_angel_threshold, true);
  c.cluster_lt = angle_cut(p.cluster_angel_threshold, true);
  c.grow1_le = angle_cut(p.normal_vector_threshold1, false);
  c.grow2_le = angle_cut(p.normal_vector_threshold2, false);
  return c;
}
}  // namespace fccf

// device buffers of a stand-alone entry point, released on every return path
struct DevTmp {
  std::vector<void*> ptrs;
  ~DevTmp() { for (void* p : ptrs) if (p) cudaFree(p); }
  template <class T> cudaError_t get(T** p, size_t bytes) { *p = nullptr; cudaError_t e = cudaMalloc((void**)p, bytes ? bytes : 4); if (e == cudaSuccess) ptrs.push_back(*p); return e; }
};

static Work lane_work(const Lane& L) {
  Work w; w.c[0] = L.c[0]; w.c[1] = L.c[1]; w.h = L.h; w.st = L.d_st;
  return w;
}
// Batch over lanes [0, G) of a group; `ws` must outlive the launches issued with it.
static Batch make_batch(fccf_ctx* ctx, Group* g, int G, std::vector<Work>& ws, ArgTable* tab) {
  ws.resize(G);
  for (int l = 0; l < G; l++) ws[l] = lane_work(g->lanes[l]);
  Batch b; b.w = ws.data(); b.G = G; b.tab = tab; b.p = ctx->p; b.cuts = make_angle_cuts(ctx->p);
  return b;
}

// lane 0 of group 0 as a batch of one, with the immediate-mode argument table (stand-alone stage entry points)
static Batch single_batch(fccf_ctx* ctx, std::vector<Work>& ws) {
  ctx->itab.off = 0; ctx->itab.overflow = false; ctx->itab.immediate = true; ctx->itab.stream = ctx->stream;
  return make_batch(ctx, ctx->groups[0], 1, ws, &ctx->itab);
}
static int set_single_call(fccf_ctx* ctx, int n0, int n1, float leaf) {
  Group& g = ctx->G0();
  CallArgs* hc = g.h_calls; hc->n0 = n0; hc->n1 = n1; hc->leaf = leaf; hc->pad = 0; hc->raw[0] = ctx->L0().d_raw[0]; hc->raw[1] = ctx->L0().d_raw[1];
  CK(cudaMemcpyAsync(g.d_calls, hc, sizeof(CallArgs), cudaMemcpyHostToDevice, ctx->stream));
  return FCCF_OK;
}

static int check_status(fccf_ctx* ctx, int st) {
  st &= ~(ST_VG_FAST_MISS | ST_SORT_MISS);          // not errors (handled by the caller: re-run of the full sequence)
  if (st == 0) return FCCF_OK;
  char b[256];
  snprintf(b, sizeof b, "device status 0x%x:%s%s%s%s%s", st, (st & ST_OCT_DEPTH) ? " octree deeper than the 32-bit Morton key" : "",
           (st & ST_HYP_OVERFLOW) ? " hypothesis pool capacity exceeded" : "", (st & ST_CENTRE_OVERFLOW) ? " cluster centre capacity exceeded" : "",
           (st & ST_HASH_FULL) ? " fine-verify lattice out of range / hash full" : "", (st & ST_KEYBITS) ? " sort key too wide" : "");
  ctx->err = b;
  return FCCF_ERR_CAPACITY;
}

// Every kernel of main() + computer_transform_guess for G lanes and the read-back of their result
// blocks, in stream order on the group's stream.  All sizes are device-side and the per-call values
// (point counts, leaf, raw-cloud pointers) are read from the group's call array, so the sequence is
// identical for every call with the same G: it is captured into a CUDA graph once and replayed.
static int group_pipeline(fccf_ctx* ctx, Group* g, int G, ArgTable* tab, uint64_t* launches, bool capturing, bool fast, bool lean = false) {
  cudaStream_t s = g->stream;
  // inside a capture a plain cudaEventRecord is only a dependency; the External flag makes a timing node
  const bool stage_events = ctx->stage_timing;
  auto rec = [&](cudaEvent_t e) { if (!stage_events && e != g->ev[2] && e != g->ev[3]) return cudaSuccess; return capturing ? cudaEventRecordWithFlags(e, s, cudaEventRecordExternal) : cudaEventRecord(e, s); };
  std::vector<Work> ws;
  Batch b = make_batch(ctx, g, G, ws, tab);
  if (capturing) { b.side = g->stream2; b.side_fork = g->fork2; b.side_join = g->join2; }
  b.lean = lean;
  struct PdlScope { bool prev; PdlScope(bool on) : prev(fccf_pdl_flag()) { fccf_pdl_flag() = on; } ~PdlScope() { fccf_pdl_flag() = prev; } } pdl_scope(capturing && pdl_wanted(G));
  launch_init_state(s, b, g->d_calls, launches);
  if (fast) CK(launch_voxelgrid_fast(s, b, 0, 2, g->vf, launches));   // main(): FCCF.cpp:1668-1678, one cluster per cloud
  else launch_voxelgrid(s, b, 0, 2, launches);
  CK(rec(g->ev[2]));
  if (fast) CK(launch_voxelgrid_fast(s, b, 1, 2, g->vf, launches));   // FCCF.cpp:1377-1387
  else launch_voxelgrid(s, b, 1, 2, launches);
  CK(rec(g->sev[0]));
  launch_planes(s, b, 2, 1, launches);               // FCCF.cpp:1400-1401
  CK(rec(g->sev[1]));
  // The fine-verify voxel table needs only the leftover cloud of the plane stage: in a captured graph it is
  // a branch of its own (second capture stream forked / joined by events), next to the mostly single-CTA
  // hypothesis, clustering and quick-verify kernels.
  if (capturing) {
    CK(cudaEventRecord(g->fork, s));
    CK(cudaStreamWaitEvent(g->stream2, g->fork, 0));
    launch_fine_verify_build(g->stream2, b, launches);
    CK(cudaEventRecord(g->join, g->stream2));
  }
  launch_hypotheses(s, b, launches);                 // FCCF.cpp:1406-1427, 1439-1462
  CK(rec(g->sev[2]));
  launch_cluster(s, b, launches);                    // FCCF.cpp:1464-1466
  CK(rec(g->sev[3]));
  launch_quick_verify(s, b, launches);               // FCCF.cpp:1468-1494
  CK(rec(g->sev[4]));
  if (capturing) CK(cudaStreamWaitEvent(s, g->join, 0));
  else launch_fine_verify_build(s, b, launches);
  launch_fine_verify_fuse(s, b, launches);           // FCCF.cpp:1499-1606
  CK(rec(g->ev[3]));
  CK(cudaMemcpyAsync(g->h_st_all, g->d_st_all, sizeof(PipeState) * (size_t)G, cudaMemcpyDeviceToHost, s));
  if (tab->overflow) { ctx->err = "argument table overflow"; return FCCF_ERR_CAPACITY; }
  return FCCF_OK;
}

static int group_graph(fccf_ctx* ctx, Group* g, int G, bool fast, bool lean, GraphRec** out) {
  if (g->params_epoch != ctx->params_epoch) { CK(cudaStreamSynchronize(g->stream)); group_drop_graphs(g); g->params_epoch = ctx->params_epoch; }
  for (GraphRec& r : g->graphs) if (r.G == G && r.fast == fast && r.lean == lean) { *out = &r; return FCCF_OK; }
  // room for one more capture?  (about 10 KB of argument blocks per lane)
  if (g->graphs.size() >= 8 || g->tab.off + (size_t)G * 12288 + 65536 > g->tab.cap) { CK(cudaStreamSynchronize(g->stream)); group_drop_graphs(g); }
  cudaStream_t s = g->stream;
  GraphRec r; r.G = G; r.fast = fast; r.lean = lean;
  size_t off0 = g->tab.off;
  g->tab.immediate = false;
  cudaGraph_t graph = nullptr;
  uint64_t cnt = 0;
  CK(cudaStreamBeginCapture(s, cudaStreamCaptureModeRelaxed));
  int rc = group_pipeline(ctx, g, G, &g->tab, &cnt, true, fast, lean);
  cudaError_t e = cudaStreamEndCapture(s, &graph);
  if (rc != FCCF_OK || e != cudaSuccess || !graph) {
    if (graph) cudaGraphDestroy(graph);
    if (rc == FCCF_OK) ctx->err = std::string("graph capture failed: ") + cudaGetErrorString(e);
    cudaGetLastError();
    g->tab.off = off0;
    return rc != FCCF_OK ? rc : FCCF_ERR_CUDA;
  }
  e = cudaGraphInstantiate(&r.exec, graph, 0);
  cudaGraphDestroy(graph);
  if (e != cudaSuccess) { ctx->err = std::string("graph instantiation failed: ") + cudaGetErrorString(e); g->tab.off = off0; return FCCF_ERR_CUDA; }
  // the argument blocks of this capture, uploaded once (stream-ordered before the first replay)
  CK(cudaMemcpyAsync(g->tab.d + off0, g->tab.h + off0, g->tab.off - off0, cudaMemcpyHostToDevice, s));
  r.launches = (int)cnt;
  g->graphs.push_back(r);
  *out = &g->graphs.back();
  return FCCF_OK;
}

// The batched pipeline of the G lanes whose call blocks are in g->d_calls (graph replay, or plain launches).
static int group_launch(fccf_ctx* ctx, Group* g, int G, bool fast, bool lean) {
  cudaStream_t s = g->stream;
  if (ctx->use_graph) {
    GraphRec* r = nullptr;
    int rc = group_graph(ctx, g, G, fast, lean, &r);
    if (rc) return rc;
    CK(cudaGraphLaunch(r->exec, s));
    g->launches += (uint64_t)r->launches;
  } else {
    g->tab.immediate = true; g->tab.off = 0; g->tab.overflow = false;
    int rc = group_pipeline(ctx, g, G, &g->tab, &g->launches, false, fast, lean);
    if (rc) return rc;
  }
  CK(cudaEventRecord(g->ev[4], s));
  return FCCF_OK;
}

// Enqueues G whole registrations on the group's stream (no host synchronisation): optional H2D of the
// raw clouds, the per-call blocks, then the batched pipeline (graph replay).  tar/src: host pointers
// (host_in) or device pointers; pair index of lane l = pair0 + l.
static int group_enqueue(fccf_ctx* ctx, Group* g, int G, int pair0, const float* const* src, const size_t* n_src, const float* const* tar, const size_t* n_tar,
                         float leaf, bool host_in) {
  cudaStream_t s = g->stream;
  g->l0 = g->launches; g->had_h2d = host_in; g->G = G; g->last_h2d = 0;
  bool staged = false;
  static const bool stage_pageable = !(getenv("FCCF_NO_STAGING") && getenv("FCCF_NO_STAGING")[0] == '1');
  CK(cudaEventRecord(g->ev[0], s));
  for (int l = 0; l < G; l++) {
    Lane& L = g->lanes[l];
    L.pair = pair0 + l;
    if (host_in) {
      const float* hp[2] = {tar[l], src[l]}; const size_t hn[2] = {n_tar[l], n_src[l]};
      for (int c = 0; c < 2; c++) {
        if (!hn[c]) continue;
        // pinned (or registered) memory: one DMA straight from the caller's buffer; pageable memory (the
        // reference's std::vector / pcl storage): through the context's pinned staging chunks
        if (!stage_pageable || host_pointer_is_pinned(hp[c])) CK(cudaMemcpyAsync(L.d_raw[c], hp[c], hn[c] * 12, cudaMemcpyHostToDevice, s));
        else {
          if (!ctx->pool) {
            int nw = 0;
            if (const char* e = getenv("FCCF_COPY_THREADS")) nw = atoi(e);
            if (nw <= 0) { int hw = (int)std::thread::hardware_concurrency(), nd = 1; cudaGetDeviceCount(&nd); nw = std::max(2, std::min(8, hw / std::max(1, nd))); }
            ctx->pool.reset(new CopyPool(ctx->device, nw, (size_t)8 << 20));   // 8 MiB chunks: per-chunk driver overhead dominates below ~4 MiB (17 GB/s at 2 MiB, 31 GB/s at 8 MiB)
            if (!ctx->pool->ok()) { ctx->pool.reset(); ctx->err = "pinned staging allocation failed"; cudaGetLastError(); return FCCF_ERR_CUDA; }
          }
          ctx->pool->add(L.d_raw[c], hp[c], hn[c] * 12);
          staged = true; ctx->staged_bytes += hn[c] * 12;
        }
      }
      g->last_h2d += (n_tar[l] + n_src[l]) * 12;
    }
    CallArgs& c = g->h_calls[l];
    c.n0 = (int)n_tar[l]; c.n1 = (int)n_src[l]; c.leaf = leaf; c.pad = 0;
    c.raw[0] = host_in ? L.d_raw[0] : tar[l]; c.raw[1] = host_in ? L.d_raw[1] : src[l];
  }
  if (staged) CK(ctx->pool->flush_into(s));      // every staged chunk issued; the group's stream waits for the workers' DMAs
  CK(cudaMemcpyAsync(g->d_calls, g->h_calls, sizeof(CallArgs) * (size_t)G, cudaMemcpyHostToDevice, s));
  CK(cudaEventRecord(g->ev[1], s));
  // cluster VoxelGrid first, unless a cloud cannot fit its scratch or this leaf missed recently
  size_t nmax = 0;
  for (int l = 0; l < G; l++) nmax = std::max(nmax, std::max(n_src[l], n_tar[l]));
  bool fast = ctx->vg_fast && g->vf.ncl > 0 && nmax <= (size_t)g->vf.stride && nmax <= (size_t)vg_fast_nmax();
  if (fast && ctx->fast_miss_ttl > 0 && leaf <= ctx->fast_miss_leaf) { fast = false; ctx->fast_miss_ttl--; }
  static const bool no_lean = getenv("FCCF_NO_LEAN") && getenv("FCCF_NO_LEAN")[0] == '1';
  const bool lean = ctx->lean_ok && ctx->use_graph && !no_lean;
  g->fast = fast; g->leaf = leaf; g->stage_timed = ctx->stage_timing; g->lean = lean;
  int rc = group_launch(ctx, g, G, fast, lean);
  if (rc) return rc;
  g->busy = true;
  return FCCF_OK;
}

// Waits for the group's registrations, returns their matrices (T_out indexed by pair) and adds the
// group's device timings to *tm.  Returns the worst status of its lanes.
static int group_finish(fccf_ctx* ctx, Group* g, float* T_out, fccf_timing* tm) {
  g->busy = false;                 // whatever happens below, the group is drained or abandoned
  CK(cudaStreamSynchronize(g->stream));
  CK(cudaGetLastError());
  {
    // Two parts of a sequence are speculative: the cluster VoxelGrid (a cloud it cannot hold raises ST_VG_FAST_MISS) and
    // the "lean" sorts without radix passes (a list beyond the one-CTA sort raises ST_SORT_MISS).  The call blocks and
    // the raw clouds are still on the device, so a miss is answered by replaying the sequence without that part.
    if (g->fast) ctx->fast_runs++;
    bool smiss_any = false;
    for (int attempt = 0; attempt < 3; attempt++) {      // a replay can bring the other miss to light (clouds that the cluster VoxelGrid dropped have no lists)
      bool vmiss = false, smiss = false;
      for (int l = 0; l < g->G; l++) { const int stt = g->lanes[l].h_st->status; vmiss = vmiss || (stt & ST_VG_FAST_MISS); smiss = smiss || (stt & ST_SORT_MISS); }
      vmiss = vmiss && g->fast; smiss = smiss && g->lean;
      if (vmiss) { ctx->fast_misses++; ctx->fast_miss_leaf = std::max(ctx->fast_miss_leaf, g->leaf); ctx->fast_miss_ttl = 64; }
      if (smiss) { ctx->lean_ok = false; ctx->lean_ttl = 64; smiss_any = true; }
      if (!vmiss && !smiss) break;
      g->fast = g->fast && !vmiss; g->lean = g->lean && !smiss;
      int rc = group_launch(ctx, g, g->G, g->fast, g->lean);
      if (rc) return rc;
      CK(cudaStreamSynchronize(g->stream));
      CK(cudaGetLastError());
    }
    const bool smiss = smiss_any;
    // lean sequences from the next call on, once every lane's lists sit well inside the one-CTA sort
    if (!smiss) {
      int longest = 0;
      for (int l = 0; l < g->G; l++) { const PipeState* hs = g->lanes[l].h_st; longest = std::max(longest, std::max(hs->hyp_off[3], hs->oct[0].S)); }
      if (ctx->lean_ttl > 0) ctx->lean_ttl--;
      else ctx->lean_ok = longest <= RS_SMALL - RS_SMALL / 16;
    }
  }
  int worst = FCCF_OK;
  for (int l = 0; l < g->G; l++) {
    Lane& L = g->lanes[l];
    for (int i = 0; i < 16; i++) T_out[16 * (size_t)L.pair + i] = L.h_st->T_final[i];
    int rc = check_status(ctx, L.h_st->status);
    if (rc) worst = rc;
  }
  if (tm) {
    float v = 0.f;
    if (g->had_h2d) { cudaEventElapsedTime(&v, g->ev[0], g->ev[1]); tm->h2d_ms += v; }
    cudaEventElapsedTime(&v, g->ev[1], g->ev[2]); tm->downsample_ms += v; tm->stage_ms[0] += v;
    cudaEventElapsedTime(&v, g->ev[2], g->ev[3]); tm->pipeline_ms += v;
    cudaEventElapsedTime(&v, g->ev[3], g->ev[4]); tm->d2h_ms += v;
    cudaEventElapsedTime(&v, g->ev[0], g->ev[4]); tm->total_ms += v;
    tm->n_launches += (int)(g->launches - g->l0);
    tm->h2d_bytes += (unsigned long long)g->last_h2d;
    tm->d2h_bytes += sizeof(PipeState) * (unsigned long long)g->G;
    if (g->stage_timed) {
      cudaEventElapsedTime(&v, g->ev[2], g->sev[0]); tm->stage_ms[1] += v;
      for (int k = 0; k < 4; k++) { cudaEventElapsedTime(&v, g->sev[k], g->sev[k + 1]); tm->stage_ms[2 + k] += v; }
      cudaEventElapsedTime(&v, g->sev[4], g->ev[3]); tm->stage_ms[6] += v;
    }
  }
  return worst;
}

// Waits for every group that still has work in flight (an earlier call that returned on an error) and marks
// it idle, without touching any result buffer: after this no copy from a caller's host buffer is pending and
// no lane refers to an old batch.
static void drain_groups(fccf_ctx* ctx) {
  for (Group* g : ctx->groups) {
    cudaStreamSynchronize(g->stream);
    g->busy = false; g->G = 0;
    for (Lane& L : g->lanes) L.pair = -1;
  }
  cudaGetLastError();
}

static int register_one(fccf_ctx* ctx, const float* src, size_t n_src, const float* tar, size_t n_tar, float leaf, float T_out[16], fccf_timing* timing, bool host_in) {
  if (!ctx) return FCCF_ERR_NO_DEVICE;
  if (!T_out || (!src && n_src) || (!tar && n_tar) || !(leaf > 0.f)) { ctx->err = "bad argument"; return FCCF_ERR_ARG; }
  CK(cudaSetDevice(ctx->device));
  Group* g = ctx->groups[0];
  drain_groups(ctx);
  int rc = ensure_capacity(ctx, g, 0, n_tar, n_src);
  if (rc) return rc;
  if ((rc = ensure_fast_scratch(ctx, g, std::max(n_tar, n_src)))) return rc;
  rc = group_enqueue(ctx, g, 1, 0, &src, &n_src, &tar, &n_tar, leaf, host_in);
  if (rc) { drain_groups(ctx); return rc; }
  fccf_timing tm; memset(&tm, 0, sizeof tm);
  rc = group_finish(ctx, g, T_out, &tm);
  if (timing) *timing = tm;
  ctx->have_run = true; ctx->leaf = leaf;
  return rc;
}

// Batch of independent pairs: chunks of up to max_lanes pairs, each chunk one batched launch sequence
// on a group; the groups rotate so that the H2D of one chunk overlaps the compute of the others.
// timing: device times summed over the chunks (each covers its whole chunk), total_ms = device time
// of the whole batch, stage_ms[7] = host wall clock.
static int register_many(fccf_ctx* ctx, int n_pairs, const float* const* src, const size_t* n_src, const float* const* tar, const size_t* n_tar,
                         float leaf, float* T_out, fccf_timing* timing, bool host_in) {
  if (!ctx) return FCCF_ERR_NO_DEVICE;
  if (n_pairs < 0 || (n_pairs && (!src || !tar || !n_src || !n_tar || !T_out)) || !(leaf > 0.f)) { ctx->err = "bad argument"; return FCCF_ERR_ARG; }
  CK(cudaSetDevice(ctx->device));
  fccf_timing acc; memset(&acc, 0, sizeof acc);
  if (n_pairs == 0) { if (timing) *timing = acc; return FCCF_OK; }
  const int nl = std::min(ctx->max_lanes, n_pairs);
  // Chunk sizes.  Host inputs: the copies of a chunk are what the next launch sequence waits for, and the
  // compute of the LAST chunk is the only part of the batch that no copy overlaps, so the tail tapers
  // (nl, ..., nl, nl/2, nl/4, nl/4).  Device-resident inputs: full chunks.
  std::vector<int> chunk;
  for (int rem = n_pairs; rem > 0;) {
    int c = std::min(nl, rem);
    if (host_in && nl >= 4 && !getenv("FCCF_NO_TAPER")) c = std::min(nl, std::max(nl / 4, rem / 2));
    if (c > rem) c = rem;
    chunk.push_back(c); rem -= c;
  }
  const int nchunks = (int)chunk.size();
  int max_groups = 4;
  if (const char* e = getenv("FCCF_GROUPS")) { max_groups = atoi(e); if (max_groups < 1) max_groups = 1; if (max_groups > 8) max_groups = 8; }
  const int ngroups = std::min(nchunks, max_groups);
  size_t nmax = 0;
  for (int b = 0; b < n_pairs; b++) { if ((!src[b] && n_src[b]) || (!tar[b] && n_tar[b])) { ctx->err = "bad argument"; return FCCF_ERR_ARG; } nmax = std::max(nmax, std::max(n_src[b], n_tar[b])); }
  drain_groups(ctx);               // nothing of an earlier (failed) call may still be in flight
  while ((int)ctx->groups.size() < ngroups) { Group* g = nullptr; int rc = group_create(ctx, &g); if (rc) return rc; ctx->groups.push_back(g); }
  // a lane is sized for the largest cloud of the pairs it will actually receive (pair p0 + l of every chunk of its group)
  {
    int p0c = 0;
    std::vector<std::vector<size_t>> need(ngroups);
    for (int k = 0; k < nchunks; k++) {
      std::vector<size_t>& nd = need[k % ngroups];
      if ((int)nd.size() < chunk[k]) nd.resize(chunk[k], 0);
      for (int l = 0; l < chunk[k]; l++) nd[l] = std::max(nd[l], std::max(n_src[p0c + l], n_tar[p0c + l]));
      p0c += chunk[k];
    }
    for (int gi = 0; gi < ngroups; gi++) {
      for (size_t l = 0; l < need[gi].size(); l++) { int rc = ensure_capacity(ctx, ctx->groups[gi], (int)l, need[gi][l], need[gi][l]); if (rc) return rc; }
      int rc = ensure_fast_scratch(ctx, ctx->groups[gi], nmax); if (rc) return rc;
    }
  }
  // device time of the whole batch: an event on group 0's stream before the first enqueue, and one
  // after it has waited for the last sequence of the other groups
  Group* Z = ctx->groups[0];
  auto t0 = std::chrono::steady_clock::now();
  CK(cudaEventRecord(Z->sev[6], Z->stream));
  int worst = FCCF_OK;
  float enq_ms = 0.f;
  int p0 = 0;
  for (int k = 0; k < nchunks; k++) {
    Group* g = ctx->groups[k % ngroups];
    if (g->busy) { int rc = group_finish(ctx, g, T_out, &acc); if (rc == FCCF_ERR_CUDA || rc == FCCF_ERR_ARG) { drain_groups(ctx); return rc; } if (rc) worst = rc; }
    const int G = chunk[k];
    auto te0 = std::chrono::steady_clock::now();
    int rc = group_enqueue(ctx, g, G, p0, src + p0, n_src + p0, tar + p0, n_tar + p0, leaf, host_in);
    enq_ms += std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - te0).count();
    if (rc) { drain_groups(ctx); return rc; }
    p0 += G;
  }
  if (getenv("FCCF_DEBUG_TIMING")) fprintf(stderr, "[fccf] batch of %d in %d chunk(s): host enqueue %.3f ms total\n", n_pairs, nchunks, enq_ms);
  for (int gi = 1; gi < ngroups; gi++) if (ctx->groups[gi]->busy) CK(cudaStreamWaitEvent(Z->stream, ctx->groups[gi]->ev[4], 0));
  CK(cudaEventRecord(Z->sev[7], Z->stream));
  for (int gi = 0; gi < ngroups; gi++) {
    Group* g = ctx->groups[gi];
    if (g->busy) { int rc = group_finish(ctx, g, T_out, &acc); if (rc == FCCF_ERR_CUDA || rc == FCCF_ERR_ARG) { drain_groups(ctx); return rc; } if (rc) worst = rc; }
  }
  CK(cudaEventSynchronize(Z->sev[7]));
  cudaEventElapsedTime(&acc.total_ms, Z->sev[6], Z->sev[7]);
  acc.stage_ms[7] = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
  if (timing) *timing = acc;
  ctx->have_run = true; ctx->leaf = leaf;    // blobs: the last pair that ran on lane 0 of group 0
  return worst;
}

extern "C" {

int fccf_register(fccf_ctx* ctx, const float* src_xyz, size_t n_src, const float* tar_xyz, size_t n_tar, float leaf, float T_out[16], fccf_timing* timing) {
  return register_one(ctx, src_xyz, n_src, tar_xyz, n_tar, leaf, T_out, timing, true);
}

int fccf_register_device(fccf_ctx* ctx, const float* d_src_xyz, size_t n_src, const float* d_tar_xyz, size_t n_tar, float leaf, float T_out[16], fccf_timing* timing) {
  return register_one(ctx, d_src_xyz, n_src, d_tar_xyz, n_tar, leaf, T_out, timing, false);
}

int fccf_register_batch(fccf_ctx* ctx, int n_pairs, const float* const* src_xyz, const size_t* n_src, const float* const* tar_xyz, const size_t* n_tar,
                        float leaf, float* T_out, fccf_timing* timing) {
  return register_many(ctx, n_pairs, src_xyz, n_src, tar_xyz, n_tar, leaf, T_out, timing, true);
}

int fccf_register_batch_device(fccf_ctx* ctx, int n_pairs, const float* const* d_src_xyz, const size_t* n_src, const float* const* d_tar_xyz, const size_t* n_tar,
                               float leaf, float* T_out, fccf_timing* timing) {
  return register_many(ctx, n_pairs, d_src_xyz, n_src, d_tar_xyz, n_tar, leaf, T_out, timing, false);
}

int fccf_voxelgrid(fccf_ctx* ctx, const float* xyz, size_t n, float leaf, float* out_xyz, int64_t* out_cell, int32_t* out_cnt, size_t* n_out) {
  if (!ctx) return FCCF_ERR_NO_DEVICE;
  if (!n_out || (!xyz && n) || !(leaf > 0.f)) { ctx->err = "bad argument"; return FCCF_ERR_ARG; }
  CK(cudaSetDevice(ctx->device));
  drain_groups(ctx);
  int rc = ensure_capacity(ctx, ctx->groups[0], 0, n, 0);
  if (rc) return rc;
  if ((rc = ensure_fast_scratch(ctx, ctx->groups[0], n))) return rc;
  cudaStream_t s = ctx->stream;
  CK(cudaStreamSynchronize(s));
  if (n) CK(cudaMemcpyAsync(ctx->L0().d_raw[0], xyz, n * 12, cudaMemcpyHostToDevice, s));
  Group& g = ctx->G0();
  // the cluster kernel first; a cloud it cannot hold raises ST_VG_FAST_MISS and the generic kernels run
  bool fast = ctx->vg_fast && g.vf.ncl > 0 && n <= (size_t)g.vf.stride && n <= (size_t)vg_fast_nmax() && !getenv("FCCF_VG_GENERIC");
  for (int attempt = 0; attempt < 2; attempt++) {
    std::vector<Work> ws; Batch w = single_batch(ctx, ws);
    if ((rc = set_single_call(ctx, (int)n, 0, leaf))) return rc;
    launch_init_state(s, w, g.d_calls, &ctx->launches);
    if (fast) CK(launch_voxelgrid_fast(s, w, 0, 1, g.vf, &ctx->launches));
    else launch_voxelgrid(s, w, 0, 1, &ctx->launches);
    CK(cudaMemcpyAsync(ctx->L0().h_st, ctx->L0().d_st, sizeof(PipeState), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    CK(cudaGetLastError());
    if (fast) ctx->fast_runs++;
    if (!fast || !(ctx->L0().h_st->status & ST_VG_FAST_MISS)) break;
    ctx->fast_misses++;
    fast = false;
  }
  size_t m = (size_t)ctx->L0().h_st->vg[0][0].n_out;
  *n_out = m;
  if (m && out_xyz) CK(cudaMemcpy(out_xyz, ctx->L0().c[0].vg_xyz[0], m * 12, cudaMemcpyDeviceToHost));
  if (m && out_cell) CK(cudaMemcpy(out_cell, ctx->L0().c[0].vg_cell[0], m * 8, cudaMemcpyDeviceToHost));
  if (m && out_cnt) CK(cudaMemcpy(out_cnt, ctx->L0().c[0].vg_cnt[0], m * 4, cudaMemcpyDeviceToHost));
  CK(cudaGetLastError());
  ctx->have_run = false;
  return FCCF_OK;
}

int fccf_extract_planes(fccf_ctx* ctx, const float* xyz, size_t n, int32_t* n_faces) {
  if (!ctx) return FCCF_ERR_NO_DEVICE;
  if (!xyz && n) { ctx->err = "bad argument"; return FCCF_ERR_ARG; }
  CK(cudaSetDevice(ctx->device));
  int rc = ensure_capacity(ctx, ctx->groups[0], 0, n, 0);
  if (rc) return rc;
  cudaStream_t s = ctx->stream;
  CK(cudaStreamSynchronize(s));
  std::vector<Work> ws; Batch w = single_batch(ctx, ws);
  if ((rc = set_single_call(ctx, 0, 0, 1.0f))) return rc;
  launch_init_state(s, w, ctx->G0().d_calls, &ctx->launches);
  if (n) CK(cudaMemcpyAsync(ctx->L0().c[0].vg_xyz[1], xyz, n * 12, cudaMemcpyHostToDevice, s));
  int nn = (int)n;
  CK(cudaMemcpyAsync(&ctx->L0().d_st->vg[1][0].n_out, &nn, 4, cudaMemcpyHostToDevice, s));
  launch_planes(s, w, 1, 1, &ctx->launches);
  CK(cudaMemcpyAsync(ctx->L0().h_st, ctx->L0().d_st, sizeof(PipeState), cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  CK(cudaGetLastError());
  if (n_faces) *n_faces = ctx->L0().h_st->ft[0].F;
  ctx->have_run = true;
  return check_status(ctx, ctx->L0().h_st->status);
}

static int score_prepare(fccf_ctx* ctx, const float* T, size_t n_hyp, const float* s1, size_t n1, const float* s2, size_t n2) {
  cudaStream_t s = ctx->stream;
  CK(cudaStreamSynchronize(s));
  auto grow = [&](float*& p, size_t& cap, size_t need) -> int {
    if (need <= cap && p) return FCCF_OK;
    if (p) CK(cudaFree(p));
    p = nullptr; cap = std::max<size_t>(need, 1024);
    CK(cudaMalloc(&p, cap * 4));
    return FCCF_OK;
  };
  int rc;
  if ((rc = grow(ctx->d_sc_s1, ctx->sc_cap1, 3 * n1 + 3))) return rc;
  if ((rc = grow(ctx->d_sc_s2, ctx->sc_cap2, 3 * n2 + 3))) return rc;
  size_t oldT = ctx->sc_capT;
  if ((rc = grow(ctx->d_sc_T, ctx->sc_capT, 16 * n_hyp + 16))) return rc;
  if (ctx->sc_capT != oldT || !ctx->d_sc_scores) { if (ctx->d_sc_scores) CK(cudaFree(ctx->d_sc_scores)); CK(cudaMalloc(&ctx->d_sc_scores, ctx->sc_capT / 16 * 4 + 64)); }
  if ((int)std::max<size_t>(n1, 1024) > ctx->sc_cap_hash) {     // sc_cap_hash: capacity in static points of the stand-alone table
    if (ctx->sc_arena.base) CK(cudaFree(ctx->sc_arena.base));
    ctx->sc_arena = Arena();
    int capn = (int)std::max<size_t>(n1, 1024);
    int rows = (int)std::min<size_t>(148 * 4, std::max<size_t>(4, ((size_t)1 << 30) / ((size_t)capn * 4)));
    rows &= ~3;
    size_t bytes = score_ws_layout(nullptr, nullptr, capn, rows);
    CK(cudaMalloc(&ctx->sc_arena.base, bytes));
    score_ws_layout(&ctx->sc_ws, ctx->sc_arena.base, capn, rows);
    ctx->sc_cap_hash = capn;
  }
  if (!ctx->d_sc_ss) { CK(cudaMalloc(&ctx->d_sc_ss, sizeof(ScoreState))); CK(cudaMalloc(&ctx->d_sc_n, 64)); }
  ctx->sc_ws.ss = ctx->d_sc_ss; ctx->sc_ws.status = &ctx->L0().d_st->status;
  int nn[2] = {(int)n1, (int)n2};
  CK(cudaMemsetAsync(&ctx->L0().d_st->status, 0, 4, s));
  CK(cudaMemcpyAsync(ctx->d_sc_n, nn, 8, cudaMemcpyHostToDevice, s));
  if (n1) CK(cudaMemcpyAsync(ctx->d_sc_s1, s1, n1 * 12, cudaMemcpyHostToDevice, s));
  if (n2) CK(cudaMemcpyAsync(ctx->d_sc_s2, s2, n2 * 12, cudaMemcpyHostToDevice, s));
  if (n_hyp) CK(cudaMemcpyAsync(ctx->d_sc_T, T, n_hyp * 64, cudaMemcpyHostToDevice, s));
  ctx->itab.off = 0; ctx->itab.overflow = false; ctx->itab.immediate = true; ctx->itab.stream = s;
  ScoreBuildJob job; job.s1 = ctx->d_sc_s1; job.n1 = ctx->d_sc_n; job.n2 = ctx->d_sc_n + 1; job.ws = ctx->sc_ws;
  launch_score_build(s, ctx->p, &job, 1, ctx->itab, &ctx->launches);
  ctx->sc_n2 = n2; ctx->sc_nhyp = n_hyp;
  return FCCF_OK;
}

int fccf_score_hypotheses(fccf_ctx* ctx, const float* T, size_t n_hyp, const float* s1_xyz, size_t n1, const float* s2_xyz, size_t n2, float* scores) {
  if (!ctx) return FCCF_ERR_NO_DEVICE;
  if ((!T && n_hyp) || !scores) { ctx->err = "bad argument"; return FCCF_ERR_ARG; }
  CK(cudaSetDevice(ctx->device));
  int rc = score_prepare(ctx, T, n_hyp, s1_xyz, n1, s2_xyz, n2);
  if (rc) return rc;
  launch_score_list(ctx->stream, ctx->p, ctx->d_sc_T, (int)n_hyp, ctx->d_sc_s2, ctx->sc_ws, ctx->d_sc_scores, ctx->itab, &ctx->launches);
  if (n_hyp) CK(cudaMemcpyAsync(scores, ctx->d_sc_scores, n_hyp * 4, cudaMemcpyDeviceToHost, ctx->stream));
  int st = 0;
  CK(cudaMemcpyAsync(&st, &ctx->L0().d_st->status, 4, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  CK(cudaGetLastError());
  ctx->L0().h_st->status = st;
  return check_status(ctx, ctx->L0().h_st->status);
}

int fccf_score_hypotheses_bench(fccf_ctx* ctx, const float* T, size_t n_hyp, const float* s1_xyz, size_t n1, const float* s2_xyz, size_t n2, int repeat,
                                float* scores, float* kernel_ms) {
  if (!ctx) return FCCF_ERR_NO_DEVICE;
  if ((!T && n_hyp) || repeat < 1) { ctx->err = "bad argument"; return FCCF_ERR_ARG; }
  CK(cudaSetDevice(ctx->device));
  int rc = score_prepare(ctx, T, n_hyp, s1_xyz, n1, s2_xyz, n2);
  if (rc) return rc;
  cudaStream_t s = ctx->stream;
  // every launch takes one argument block of the immediate table: rewind it to the mark each time (the block
  // is stream-ordered before its kernel and identical from launch to launch), so no repeat count overflows it
  const size_t mark = ctx->itab.off;
  for (int w = 0; w < 3; w++) { ctx->itab.off = mark; launch_score_list(s, ctx->p, ctx->d_sc_T, (int)n_hyp, ctx->d_sc_s2, ctx->sc_ws, ctx->d_sc_scores, ctx->itab, &ctx->launches); }
  CK(cudaStreamSynchronize(s));
  CK(cudaEventRecord(ctx->G0().ev[0], s));
  for (int r = 0; r < repeat; r++) { ctx->itab.off = mark; launch_score_list(s, ctx->p, ctx->d_sc_T, (int)n_hyp, ctx->d_sc_s2, ctx->sc_ws, ctx->d_sc_scores, ctx->itab, &ctx->launches); }
  CK(cudaEventRecord(ctx->G0().ev[1], s));
  CK(cudaStreamSynchronize(s));
  if (ctx->itab.overflow) { ctx->err = "argument table overflow"; return FCCF_ERR_CAPACITY; }
  float ms = 0; cudaEventElapsedTime(&ms, ctx->G0().ev[0], ctx->G0().ev[1]);
  if (kernel_ms) *kernel_ms = ms / repeat;
  if (scores && n_hyp) CK(cudaMemcpy(scores, ctx->d_sc_scores, n_hyp * 4, cudaMemcpyDeviceToHost));
  CK(cudaGetLastError());
  return FCCF_OK;
}

int fccf_score_best(fccf_ctx* ctx, size_t index_base, int64_t* packed_host, int64_t* packed_device) {
  if (!ctx) return FCCF_ERR_NO_DEVICE;
  if (!ctx->d_sc_scores || (!packed_host && !packed_device)) { ctx->err = "bad argument / no previous fccf_score_hypotheses call"; return FCCF_ERR_ARG; }
  CK(cudaSetDevice(ctx->device));
  if (!ctx->d_sc_best) CK(cudaMalloc(&ctx->d_sc_best, 8));
  launch_score_best(ctx->stream, ctx->d_sc_scores, (int)ctx->sc_nhyp, (long long)index_base, ctx->d_sc_best, &ctx->launches);
  if (packed_device) CK(cudaMemcpyAsync(packed_device, ctx->d_sc_best, 8, cudaMemcpyDeviceToDevice, ctx->stream));
  long long h = 0;
  if (packed_host) CK(cudaMemcpyAsync(&h, ctx->d_sc_best, 8, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  if (packed_host) *packed_host = (int64_t)h;
  CK(cudaGetLastError());
  return FCCF_OK;
}

int fccf_score_counts(fccf_ctx* ctx, size_t hyp, int32_t* rows, size_t cap_rows, size_t* n_rows) {
  if (!ctx) return FCCF_ERR_NO_DEVICE;
  if (!rows || !n_rows || !ctx->d_sc_ss) { ctx->err = "bad argument / no previous fccf_score_hypotheses call"; return FCCF_ERR_ARG; }
  if (hyp >= ctx->sc_nhyp) { ctx->err = "hypothesis index beyond the last scored list"; return FCCF_ERR_ARG; }
  CK(cudaSetDevice(ctx->device));
  DevTmp tmp;
  int* d_rows = nullptr; int* d_n = nullptr;
  CK(tmp.get(&d_rows, std::max<size_t>(cap_rows, 1) * 20));
  CK(tmp.get(&d_n, 4));
  CK(cudaMemsetAsync(d_n, 0, 4, ctx->stream));
  ctx->itab.off = 0; ctx->itab.overflow = false; ctx->itab.immediate = true; ctx->itab.stream = ctx->stream;
  launch_score_dump(ctx->stream, ctx->p, ctx->d_sc_T + 16 * hyp, ctx->d_sc_s2, ctx->sc_ws, d_rows, (int)cap_rows, d_n, ctx->itab, &ctx->launches);
  int n = 0;
  CK(cudaMemcpyAsync(&n, d_n, 4, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  *n_rows = (size_t)n;
  size_t m = std::min<size_t>((size_t)n, cap_rows);
  if (m) CK(cudaMemcpy(rows, d_rows, m * 20, cudaMemcpyDeviceToHost));
  CK(cudaGetLastError());
  return FCCF_OK;
}

int fccf_quick_verify(fccf_ctx* ctx, float* T, size_t n_hyp, const float* planes1, int f1, const float* planes2, int f2, float* scores,
                      int32_t* pair_count, int32_t* pairs, int32_t* iters) {
  if (!ctx) return FCCF_ERR_NO_DEVICE;
  if (!T || !scores || f1 < 0 || f2 < 0 || f1 > FCCF_MAXF || f2 > FCCF_MAXF) { ctx->err = "bad argument"; return FCCF_ERR_ARG; }
  CK(cudaSetDevice(ctx->device));
  if (n_hyp == 0) return FCCF_OK;
  float *dT = nullptr, *dp1 = nullptr, *dp2 = nullptr, *dsc = nullptr; int *dnp = nullptr, *dpr = nullptr, *dit = nullptr;
  std::vector<float> p1(8 * FCCF_MAXF, 0.f), p2(8 * FCCF_MAXF, 0.f);
  for (int i = 0; i < f1; i++) for (int k = 0; k < 7; k++) p1[8 * i + k] = planes1[7 * i + k];
  for (int i = 0; i < f2; i++) for (int k = 0; k < 7; k++) p2[8 * i + k] = planes2[7 * i + k];
  DevTmp tmp;
  CK(tmp.get(&dT, n_hyp * 64)); CK(tmp.get(&dp1, p1.size() * 4)); CK(tmp.get(&dp2, p2.size() * 4)); CK(tmp.get(&dsc, n_hyp * 4));
  CK(tmp.get(&dnp, n_hyp * 4)); CK(tmp.get(&dpr, n_hyp * 128)); CK(tmp.get(&dit, n_hyp * 4));
  CK(cudaMemcpy(dT, T, n_hyp * 64, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dp1, p1.data(), p1.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dp2, p2.data(), p2.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemset(dpr, 0xff, n_hyp * 128));
  launch_quick_verify_list(ctx->stream, ctx->p, dT, (int)n_hyp, dp1, f1, dp2, f2, dsc, dnp, dpr, dit, &ctx->launches);
  CK(cudaStreamSynchronize(ctx->stream));
  CK(cudaMemcpy(T, dT, n_hyp * 64, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(scores, dsc, n_hyp * 4, cudaMemcpyDeviceToHost));
  if (pair_count) CK(cudaMemcpy(pair_count, dnp, n_hyp * 4, cudaMemcpyDeviceToHost));
  if (pairs) CK(cudaMemcpy(pairs, dpr, n_hyp * 128, cudaMemcpyDeviceToHost));
  if (iters) CK(cudaMemcpy(iters, dit, n_hyp * 4, cudaMemcpyDeviceToHost));
  CK(cudaGetLastError());
  return FCCF_OK;
}

// ---------------------------------------------------------------------------------------------
// stage intermediates of the last run
// ---------------------------------------------------------------------------------------------
int fccf_debug_blob(fccf_ctx* ctx, const char* name_c, void* dst, size_t cap_bytes, size_t* bytes, int* dtype) {
  if (!ctx) return FCCF_ERR_NO_DEVICE;
  if (name_c && !strcmp(name_c, "vg_fast")) {      // [sequences run through the cluster VoxelGrid, of which re-run through the generic kernels]
    long long v[2] = {(long long)ctx->fast_runs, (long long)ctx->fast_misses};
    if (bytes) *bytes = sizeof v;
    if (dtype) *dtype = FCCF_I64;
    if (dst) { if (cap_bytes < sizeof v) { ctx->err = "blob buffer too small"; return FCCF_ERR_ARG; } memcpy(dst, v, sizeof v); }
    return FCCF_OK;
  }
  if (!name_c || !ctx->have_run) { ctx->err = "no run to inspect"; return FCCF_ERR_ARG; }
  CK(cudaSetDevice(ctx->device));
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->itab.off = 0; ctx->itab.overflow = false; ctx->itab.immediate = true; ctx->itab.stream = ctx->stream;
  std::string name(name_c);
  const PipeState& st = *ctx->L0().h_st;
  std::vector<char> out; int dt = FCCF_F32;
  cudaError_t fetch_err = cudaSuccess;
  auto fetch = [&](const void* dptr, size_t nbytes) -> std::vector<char> {
    std::vector<char> v(nbytes);
    if (nbytes) { cudaError_t e = cudaMemcpy(v.data(), dptr, nbytes, cudaMemcpyDeviceToHost); if (e != cudaSuccess) fetch_err = e; }
    return v;
  };
  auto put = [&](const void* p, size_t nbytes, int d) { out.assign((const char*)p, (const char*)p + nbytes); dt = d; };
  char last = name.empty() ? 0 : name.back();
  std::string stem = name.substr(0, name.size() ? name.size() - 1 : 0);
  int ci = (last == '1') ? 0 : ((last == '2') ? 1 : -1);       // cloud tag
  int ti = (last >= '0' && last <= '2') ? last - '0' : -1;      // type tag
  bool ok = true;
  if ((stem == "vg1_xyz" || stem == "vg2_xyz") && ci >= 0) { int sgi = stem[2] - '1'; out = fetch(ctx->L0().c[ci].vg_xyz[sgi], (size_t)st.vg[sgi][ci].n_out * 12); dt = FCCF_F32; }
  else if ((stem == "vg1_cell" || stem == "vg2_cell") && ci >= 0) { int sgi = stem[2] - '1'; out = fetch(ctx->L0().c[ci].vg_cell[sgi], (size_t)st.vg[sgi][ci].n_out * 8); dt = FCCF_I64; }
  else if ((stem == "vg1_cnt" || stem == "vg2_cnt") && ci >= 0) { int sgi = stem[2] - '1'; out = fetch(ctx->L0().c[ci].vg_cnt[sgi], (size_t)st.vg[sgi][ci].n_out * 4); dt = FCCF_I32; }
  else if (stem == "cloud_centroid" && ci >= 0) put(st.oct[ci].cc, 12, FCCF_F32);
  else if (stem == "oct_min" && ci >= 0) put(st.oct[ci].mn, 24, FCCF_F64);
  else if (stem == "oct_depth" && ci >= 0) put(&st.oct[ci].depth, 4, FCCF_I32);
  else if ((stem == "vox_key" || stem == "vox_cnt" || stem == "vox_flag" || stem == "vox_plane") && ci >= 0) {
    int V = st.oct[ci].V;
    std::vector<char> rec = fetch(ctx->L0().c[ci].vox_rec, (size_t)V * 48);
    const float* r = (const float*)rec.data();
    if (stem == "vox_key") { std::vector<int> k(3 * (size_t)V); for (int v = 0; v < V; v++) for (int a = 0; a < 3; a++) memcpy(&k[3 * v + a], &r[12 * v + 9 + a], 4); put(k.data(), k.size() * 4, FCCF_I32); }
    else if (stem == "vox_cnt") { std::vector<int> k(V); for (int v = 0; v < V; v++) k[v] = (int)r[12 * v + 7]; put(k.data(), k.size() * 4, FCCF_I32); }
    else if (stem == "vox_flag") { std::vector<int> k(V); for (int v = 0; v < V; v++) k[v] = (int)r[12 * v + 8]; put(k.data(), k.size() * 4, FCCF_I32); }
    else { std::vector<float> k(8 * (size_t)V); for (int v = 0; v < V; v++) { for (int a = 0; a < 7; a++) k[8 * v + a] = r[12 * v + a]; k[8 * v + 7] = r[12 * v + 8] > 0 ? r[12 * v + 7] : 0.f; } put(k.data(), k.size() * 4, FCCF_F32); }
  }
  else if (stem == "vox_pidx" && ci >= 0) { out = fetch(ctx->L0().c[ci].idxA, (size_t)st.oct[ci].n * 4); dt = FCCF_I32; }
  else if (stem == "sub" && ci >= 0) { out = fetch(ctx->L0().c[ci].sub, (size_t)st.oct[ci].S * 12); dt = FCCF_F32; }
  else if (stem == "pvox" && ci >= 0) {
    int Vp = st.oct[ci].Vp; std::vector<char> raw = fetch(ctx->L0().c[ci].pvox, (size_t)Vp * 32); const float* r = (const float*)raw.data();
    std::vector<float> k(7 * (size_t)Vp); for (int v = 0; v < Vp; v++) for (int a = 0; a < 7; a++) k[7 * v + a] = r[8 * v + a];
    put(k.data(), k.size() * 4, FCCF_F32);
  }
  else if (stem == "grow_label" && ci >= 0) { out = fetch(ctx->L0().c[ci].grow_label, (size_t)st.oct[ci].Vp * 4); dt = FCCF_I32; }
  else if (stem == "merge_label" && ci >= 0) { out = fetch(ctx->L0().c[ci].merge_label, (size_t)st.oct[ci].Vp * 4); dt = FCCF_I32; }
  else if (stem == "n_stage1_faces" && ci >= 0) put(&st.oct[ci].F1, 4, FCCF_I32);
  else if (stem == "face_plane" && ci >= 0) { std::vector<float> k; for (int f = 0; f < st.ft[ci].F; f++) for (int a = 0; a < 7; a++) k.push_back(st.ft[ci].plane[f][a]); put(k.data(), k.size() * 4, FCCF_F32); }
  else if (stem == "face_nvox" && ci >= 0) { std::vector<int> k; for (int f = 0; f < st.ft[ci].F; f++) k.push_back((int)st.ft[ci].plane[f][7]); put(k.data(), k.size() * 4, FCCF_I32); }
  else if (stem == "face_id" && ci >= 0) put(st.ft[ci].id, (size_t)st.ft[ci].F * 4, FCCF_I32);
  else if (stem == "face_theta" && ci >= 0) put(st.ft[ci].theta, (size_t)st.ft[ci].F * 8, FCCF_F64);
  else if (stem == "face_off" && ci >= 0) { out = fetch(ctx->L0().c[ci].face_off, (size_t)(st.ft[ci].F + 1) * 4); dt = FCCF_I32; }
  else if (stem == "face_vox" && ci >= 0) {
    std::vector<char> off = fetch(ctx->L0().c[ci].face_off, (size_t)(st.ft[ci].F + 1) * 4);
    int tot = ((const int*)off.data())[st.ft[ci].F];
    out = fetch(ctx->L0().c[ci].face_vox, (size_t)tot * 4); dt = FCCF_I32;
  }
  else if (stem == "base" && ci >= 0) { std::vector<int> k; const BaseTable& b = st.base[ci]; for (int i = 0; i < b.B; i++) { k.push_back(b.i[i]); k.push_back(b.j[i]); k.push_back(b.type[i]); } put(k.data(), k.size() * 4, FCCF_I32); }
  else if (stem == "base_angle" && ci >= 0) put(st.base[ci].angle, (size_t)st.base[ci].B * 4, FCCF_F32);
  else if (name == "matches") {
    int NM = st.n_match, B2 = st.base[1].B;
    std::vector<char> raw = fetch(ctx->L0().h.match_cnt, (size_t)NM * 4); const int* mc = (const int*)raw.data();
    std::vector<int> k; for (int i = 0; i < NM; i++) if (mc[i] > 0) { k.push_back(i / B2); k.push_back(i % B2); k.push_back(mc[i]); }
    put(k.data(), k.size() * 4, FCCF_I32);
  }
  else if (name == "n_hyp") put(st.n_hyp, 12, FCCF_I32);
  else if (name == "n_centres") put(st.n_centre, 12, FCCF_I32);
  else if (name == "cluster_num") put(st.cluster_num, 12, FCCF_I32);
  else if (stem == "hyp" && ti >= 0) { out = fetch(ctx->L0().h.hyp_T + (size_t)st.hyp_off[ti] * 12, (size_t)st.n_hyp[ti] * 48); dt = FCCF_F32; }
  else if (stem == "hyp_qt" && ti >= 0) {
    int n = st.n_hyp[ti]; std::vector<char> raw = fetch(ctx->L0().h.hyp_qt + (size_t)st.hyp_off[ti] * 8, (size_t)n * 32); const float* r = (const float*)raw.data();
    std::vector<float> k(7 * (size_t)n); for (int i = 0; i < n; i++) for (int a = 0; a < 7; a++) k[7 * i + a] = r[8 * i + a];
    put(k.data(), k.size() * 4, FCCF_F32);
  }
  else if ((stem == "cluster_seed_sorted" || stem == "cluster_size_sorted") && ti >= 0) {
    int K = st.n_seeds[ti]; size_t base = (size_t)st.hyp_off[ti];
    std::vector<char> pk = fetch(ctx->L0().h.c_key + base, (size_t)K * 4), pp = fetch(ctx->L0().h.c_perm + base, (size_t)K * 4), ps = fetch(ctx->L0().h.c_seeds + base, (size_t)K * 4);
    if (stem == "cluster_size_sorted") { out = pk; }
    else { std::vector<int> k(K); for (int i = 0; i < K; i++) k[i] = ((const int*)ps.data())[((const int*)pp.data())[i]]; put(k.data(), k.size() * 4, FCCF_I32); }
    dt = FCCF_I32;
  }
  else if (stem == "centre" && ti >= 0) {
    int n = st.n_centre[ti]; std::vector<char> raw = fetch(ctx->L0().h.centre + (size_t)ti * FCCF_MAXCENTRE * 8, (size_t)n * 32); const float* r = (const float*)raw.data();
    std::vector<float> k(7 * (size_t)n); for (int i = 0; i < n; i++) for (int a = 0; a < 7; a++) k[7 * i + a] = r[8 * i + a];
    put(k.data(), k.size() * 4, FCCF_F32);
  }
  else if (stem == "qv_score" && ti >= 0) { out = fetch(ctx->L0().h.qv_score + (size_t)ti * FCCF_MAXCENTRE, (size_t)st.n_centre[ti] * 4); dt = FCCF_F32; }
  else if (stem == "qv_T" && ti >= 0) { out = fetch(ctx->L0().h.qv_T + (size_t)ti * FCCF_MAXCENTRE * 16, (size_t)st.n_centre[ti] * 64); dt = FCCF_F32; }
  else if (stem == "qv_iters" && ti >= 0) { out = fetch(ctx->L0().h.qv_iters + (size_t)ti * FCCF_MAXCENTRE, (size_t)st.n_centre[ti] * 4); dt = FCCF_I32; }
  else if ((stem == "qv_pairs" || stem == "qv_pair_off") && ti >= 0) {
    int n = st.n_centre[ti];
    std::vector<char> np = fetch(ctx->L0().h.qv_npair + (size_t)ti * FCCF_MAXCENTRE, (size_t)n * 4), pr = fetch(ctx->L0().h.qv_pairs + (size_t)ti * FCCF_MAXCENTRE * 32, (size_t)n * 128);
    std::vector<int> pairs, off;
    for (int i = 0; i < n; i++) { off.push_back((int)pairs.size() / 2); int c = ((const int*)np.data())[i]; for (int k = 0; k < 2 * c; k++) pairs.push_back(((const int*)pr.data())[32 * i + k]); }
    off.push_back((int)pairs.size() / 2);
    if (stem == "qv_pairs") put(pairs.data(), pairs.size() * 4, FCCF_I32); else put(off.data(), off.size() * 4, FCCF_I32);
  }
  else if (stem == "top_T" && ti >= 0) { out = fetch(ctx->L0().h.top_T + (size_t)ti * fccf_topk(ctx->p) * 16, (size_t)st.n_top[ti] * 64); dt = FCCF_F32; }
  else if (stem == "top_s1" && ti >= 0) { out = fetch(ctx->L0().h.top_s1 + (size_t)ti * fccf_topk(ctx->p), (size_t)st.n_top[ti] * 4); dt = FCCF_F32; }
  else if (stem == "top_s2" && ti >= 0) { out = fetch(ctx->L0().h.top_s2 + (size_t)ti * fccf_topk(ctx->p), (size_t)st.n_top[ti] * 4); dt = FCCF_F32; }
  else if (stem == "top_centre" && ti >= 0) { out = fetch(ctx->L0().h.top_centre + (size_t)ti * fccf_topk(ctx->p), (size_t)st.n_top[ti] * 4); dt = FCCF_I32; }
  else if ((stem == "fv_counts" || stem == "fv_off") && ti >= 0) {
    // per-voxel (s,t) of every fine-verified hypothesis of this type, rows sorted lexicographically
    ScoreWS ws = ctx->L0().h.fv; ws.ss = &ctx->L0().d_st->fv; ws.status = &ctx->L0().d_st->status;
    int cap_rows = std::max(st.fv.n_occ, 1);
    int* d_rows = nullptr; int* d_n = nullptr;
    DevTmp tmp;
    CK(tmp.get(&d_rows, (size_t)cap_rows * 20)); CK(tmp.get(&d_n, 4));
    std::vector<int> all, off;
    for (int k = 0; k < st.n_top[ti]; k++) {
      CK(cudaMemsetAsync(d_n, 0, 4, ctx->stream));
      launch_score_dump(ctx->stream, ctx->p, ctx->L0().h.top_T + ((size_t)ti * fccf_topk(ctx->p) + k) * 16, ctx->L0().c[1].sub, ws, d_rows, cap_rows, d_n, ctx->itab, &ctx->launches);
      int n = 0; CK(cudaMemcpyAsync(&n, d_n, 4, cudaMemcpyDeviceToHost, ctx->stream)); CK(cudaStreamSynchronize(ctx->stream));
      n = std::min(n, cap_rows);
      std::vector<int> rows(5 * (size_t)n); if (n) CK(cudaMemcpy(rows.data(), d_rows, (size_t)n * 20, cudaMemcpyDeviceToHost));
      std::vector<int> order(n); for (int i = 0; i < n; i++) order[i] = i;
      std::sort(order.begin(), order.end(), [&](int a, int b) { return std::lexicographical_compare(&rows[5 * a], &rows[5 * a + 3], &rows[5 * b], &rows[5 * b + 3]); });
      off.push_back((int)all.size() / 5);
      for (int i = 0; i < n; i++) for (int a = 0; a < 5; a++) all.push_back(rows[5 * order[i] + a]);
    }
    off.push_back((int)all.size() / 5);
    if (stem == "fv_counts") put(all.data(), all.size() * 4, FCCF_I32); else put(off.data(), off.size() * 4, FCCF_I32);
  }
  else if (name == "type_best") put(st.type_best, sizeof st.type_best, FCCF_F32);
  else if (name == "prof") put(st.prof, sizeof st.prof, FCCF_I64);
  else if (name == "final_T") put(st.T_final, 64, FCCF_F32);
  else ok = false;
  if (!ok) { ctx->err = "unknown blob: " + name; return FCCF_ERR_ARG; }
  if (fetch_err != cudaSuccess) { ctx->err = std::string("blob read-back failed: ") + cudaGetErrorString(fetch_err); cudaGetLastError(); return FCCF_ERR_CUDA; }
  if (bytes) *bytes = out.size();
  if (dtype) *dtype = dt;
  if (dst) { if (out.size() > cap_bytes) { ctx->err = "blob buffer too small"; return FCCF_ERR_ARG; } if (!out.empty()) memcpy(dst, out.data(), out.size()); }
  CK(cudaGetLastError());
  return FCCF_OK;
}

int fccf_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

// contiguous [lo, hi) of n ordered items for part i of k (the first n % k parts get one more)
static void part_range(size_t n, int i, int k, size_t* lo, size_t* hi) {
  size_t q = n / (size_t)k, r = n % (size_t)k;
  *lo = (size_t)i * q + std::min<size_t>((size_t)i, r);
  *hi = *lo + q + ((size_t)i < r ? 1 : 0);
}

int fccf_register_batch_multi(fccf_ctx* const* ctxs, int n_ctx, int n_pairs, const float* const* src_xyz, const size_t* n_src,
                              const float* const* tar_xyz, const size_t* n_tar, float leaf, float* T_out, fccf_timing* timing) {
  if (!ctxs || n_ctx < 1) return FCCF_ERR_ARG;
  for (int i = 0; i < n_ctx; i++) if (!ctxs[i]) return FCCF_ERR_NO_DEVICE;
  if (n_ctx == 1) return register_many(ctxs[0], n_pairs, src_xyz, n_src, tar_xyz, n_tar, leaf, T_out, timing, true);
  if (n_pairs < 0 || (n_pairs && (!src_xyz || !tar_xyz || !n_src || !n_tar || !T_out))) { ctxs[0]->err = "bad argument"; return FCCF_ERR_ARG; }
  std::vector<int> rc(n_ctx, FCCF_OK);
  std::vector<std::thread> th;
  for (int i = 0; i < n_ctx; i++) {
    th.emplace_back([&, i]() {
      size_t lo, hi; part_range((size_t)n_pairs, i, n_ctx, &lo, &hi);
      fccf_timing tm; memset(&tm, 0, sizeof tm);
      rc[i] = register_many(ctxs[i], (int)(hi - lo), src_xyz + lo, n_src + lo, tar_xyz + lo, n_tar + lo, leaf, T_out + 16 * lo, &tm, true);
      if (timing) timing[i] = tm;
    });
  }
  for (std::thread& t : th) t.join();
  int worst = FCCF_OK;
  for (int i = 0; i < n_ctx; i++) if (rc[i] != FCCF_OK && (worst == FCCF_OK || rc[i] != FCCF_ERR_CAPACITY)) worst = rc[i];
  return worst;
}

int fccf_score_sharded(fccf_ctx* const* ctxs, int n_ctx, const float* T, size_t n_hyp, const float* s1_xyz, size_t n1, const float* s2_xyz, size_t n2,
                       float* scores, float* best_score, int64_t* best_index, int* used_nccl) {
  if (!ctxs || n_ctx < 1) return FCCF_ERR_ARG;
  for (int i = 0; i < n_ctx; i++) if (!ctxs[i]) return FCCF_ERR_NO_DEVICE;
  if (!T && n_hyp) { ctxs[0]->err = "bad argument"; return FCCF_ERR_ARG; }
  std::vector<float> tmp;
  if (!scores) { tmp.resize(std::max<size_t>(n_hyp, 1)); scores = tmp.data(); }
  std::vector<int> rc(n_ctx, FCCF_OK);
  auto work = [&](int i) {
    fccf_ctx* ctx = ctxs[i];
    size_t lo, hi; part_range(n_hyp, i, n_ctx, &lo, &hi);
    rc[i] = fccf_score_hypotheses(ctx, T + 16 * lo, hi - lo, s1_xyz, n1, s2_xyz, n2, scores + lo);
    if (rc[i] != FCCF_OK && rc[i] != FCCF_ERR_CAPACITY) return;
    cudaSetDevice(ctx->device);
    if (!ctx->d_sc_best && cudaMalloc(&ctx->d_sc_best, 8) != cudaSuccess) { cudaGetLastError(); rc[i] = FCCF_ERR_CUDA; return; }
    // block-reduced argmax of this shard: packed (score, global index) word, left on the device for the all-reduce
    launch_score_best(ctx->stream, ctx->d_sc_scores, (int)(hi - lo), (long long)lo, ctx->d_sc_best, &ctx->launches);
  };
  if (n_ctx == 1) work(0);
  else {
    std::vector<std::thread> th;
    for (int i = 0; i < n_ctx; i++) th.emplace_back(work, i);
    for (std::thread& t : th) t.join();
  }
  for (int i = 0; i < n_ctx; i++) if (rc[i] != FCCF_OK && rc[i] != FCCF_ERR_CAPACITY) { if (i) ctxs[0]->err = ctxs[i]->err; return rc[i]; }
  std::vector<long long> words(n_ctx, LLONG_MIN);
  bool nccl = false;
  if (n_ctx > 1) {
    std::vector<int> devs; std::vector<cudaStream_t> streams; std::vector<const long long*> wd;
    bool distinct = true;
    for (int i = 0; i < n_ctx; i++) { for (int d : devs) distinct = distinct && d != ctxs[i]->device; devs.push_back(ctxs[i]->device); streams.push_back(ctxs[i]->stream); wd.push_back(ctxs[i]->d_sc_best); }
    if (distinct) nccl = nccl_allreduce_max_i64(devs, streams, wd, words.data());     // one 8-byte all-reduce(max) over NVLink
  }
  if (!nccl) {
    for (int i = 0; i < n_ctx; i++) {
      fccf_ctx* ctx = ctxs[i];
      CK(cudaSetDevice(ctx->device));
      CK(cudaMemcpyAsync(&words[i], ctx->d_sc_best, 8, cudaMemcpyDeviceToHost, ctx->stream));
      CK(cudaStreamSynchronize(ctx->stream));
    }
    long long m = *std::max_element(words.begin(), words.end());
    for (long long& w : words) w = m;
  }
  const long long best = words[0];
  if (used_nccl) *used_nccl = nccl ? 1 : 0;
  const int key = (int)(best >> 32);
  if (best_index) *best_index = n_hyp ? (int64_t)(0xffffffffll - (best & 0xffffffffll)) : -1;
  if (best_score) {
    float sc;
    if (key == INT_MIN || !n_hyp) sc = std::nanf("");
    else { int b = key >= 0 ? key : (key ^ 0x7fffffff); memcpy(&sc, &b, 4); }
    *best_score = sc;
  }
  return FCCF_OK;
}

// ---------------------------------------------------------------------------------------------
// stand-alone entry points of the small stages: inputs go straight into lane 0's state block, the
// stage's own kernels run, results are read through fccf_debug_blob
// ---------------------------------------------------------------------------------------------
static int stage_begin(fccf_ctx* ctx) {
  CK(cudaSetDevice(ctx->device));
  drain_groups(ctx);
  int rc = ensure_capacity(ctx, ctx->groups[0], 0, 1024, 1024);
  if (rc) return rc;
  cudaStream_t s = ctx->stream;
  CK(cudaStreamSynchronize(s));
  if ((rc = set_single_call(ctx, 0, 0, 1.0f))) return rc;
  return FCCF_OK;
}
static int stage_end(fccf_ctx* ctx) {
  cudaStream_t s = ctx->stream;
  CK(cudaMemcpyAsync(ctx->L0().h_st, ctx->L0().d_st, sizeof(PipeState), cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  CK(cudaGetLastError());
  ctx->have_run = true;
  return check_status(ctx, ctx->L0().h_st->status);
}
static void fill_face_table(FaceTable& ft, const float* planes, const double* theta, int f) {
  memset(&ft, 0, sizeof ft);
  ft.F = f;
  for (int i = 0; i < f; i++) {
    for (int k = 0; k < 7; k++) ft.plane[i][k] = planes[7 * i + k];
    ft.plane[i][7] = 1.f; ft.theta[i] = theta ? theta[i] : 0.0; ft.id[i] = i;
  }
}

int fccf_hypotheses(fccf_ctx* ctx, const float* planes1, const double* theta1, int f1, const float* planes2, const double* theta2, int f2, int32_t n_hyp[3]) {
  if (!ctx) return FCCF_ERR_NO_DEVICE;
  if (f1 < 0 || f2 < 0 || f1 > FCCF_MAXF || f2 > FCCF_MAXF || (f1 && !planes1) || (f2 && !planes2)) { ctx->err = "bad argument"; return FCCF_ERR_ARG; }
  int rc = stage_begin(ctx);
  if (rc) return rc;
  cudaStream_t s = ctx->stream;
  std::vector<Work> ws; Batch w = single_batch(ctx, ws);
  launch_init_state(s, w, ctx->G0().d_calls, &ctx->launches);
  FaceTable ft[2];
  fill_face_table(ft[0], planes1, theta1, f1); fill_face_table(ft[1], planes2, theta2, f2);
  CK(cudaMemcpyAsync(&ctx->L0().d_st->ft[0], ft, sizeof ft, cudaMemcpyHostToDevice, s));
  CK(cudaStreamSynchronize(s));        // ft lives on this stack frame
  launch_hypotheses(s, w, &ctx->launches);
  rc = stage_end(ctx);
  if (n_hyp) for (int i = 0; i < 3; i++) n_hyp[i] = ctx->L0().h_st->n_hyp[i];
  return rc;
}

int fccf_base_pairs(fccf_ctx* ctx, const float* planes, const double* theta, int f, int32_t* pairs, float* angles, int32_t* n_pairs) {
  if (!ctx) return FCCF_ERR_NO_DEVICE;
  if (!n_pairs) { ctx->err = "bad argument"; return FCCF_ERR_ARG; }
  int rc = fccf_hypotheses(ctx, planes, theta, f, nullptr, nullptr, 0, nullptr);      // an empty second table: no match, no hypothesis
  if (rc) return rc;
  const BaseTable& b = ctx->L0().h_st->base[0];
  *n_pairs = b.B;
  for (int i = 0; i < b.B; i++) {
    if (pairs) { pairs[3 * i] = b.i[i]; pairs[3 * i + 1] = b.j[i]; pairs[3 * i + 2] = b.type[i]; }
    if (angles) angles[i] = b.angle[i];
  }
  return FCCF_OK;
}

int fccf_cluster(fccf_ctx* ctx, const float* qt7, const int32_t n_hyp[3], int32_t n_centres[3]) {
  if (!ctx) return FCCF_ERR_NO_DEVICE;
  if (!n_hyp || n_hyp[0] < 0 || n_hyp[1] < 0 || n_hyp[2] < 0) { ctx->err = "bad argument"; return FCCF_ERR_ARG; }
  const long long tot = (long long)n_hyp[0] + n_hyp[1] + n_hyp[2];
  if (tot > ctx->cap_hyp || (tot && !qt7)) { ctx->err = "bad argument / more hypotheses than the arena holds"; return FCCF_ERR_ARG; }
  int rc = stage_begin(ctx);
  if (rc) return rc;
  cudaStream_t s = ctx->stream;
  std::vector<Work> ws; Batch w = single_batch(ctx, ws);
  launch_init_state(s, w, ctx->G0().d_calls, &ctx->launches);
  std::vector<float> q8((size_t)8 * (size_t)std::max<long long>(tot, 1), 0.f);
  for (long long i = 0; i < tot; i++) for (int k = 0; k < 7; k++) q8[8 * i + k] = qt7[7 * i + k];
  int hdr[7] = {n_hyp[0], n_hyp[1], n_hyp[2], 0, n_hyp[0], n_hyp[0] + n_hyp[1], (int)tot};      // n_hyp[3], hyp_off[4]
  static_assert(offsetof(PipeState, hyp_off) == offsetof(PipeState, n_hyp) + 12, "n_hyp and hyp_off are adjacent");
  if (tot) CK(cudaMemcpyAsync(ctx->L0().h.hyp_qt, q8.data(), (size_t)tot * 32, cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(&ctx->L0().d_st->n_hyp[0], hdr, sizeof hdr, cudaMemcpyHostToDevice, s));
  CK(cudaStreamSynchronize(s));
  launch_cluster(s, w, &ctx->launches);
  rc = stage_end(ctx);
  if (n_centres) for (int i = 0; i < 3; i++) n_centres[i] = ctx->L0().h_st->n_centre[i];
  return rc;
}

int fccf_fuse(fccf_ctx* ctx, const float* top_T, const float* s1, const float* s2, const int32_t n_top[3], int k, float T_out[16]) {
  if (!ctx) return FCCF_ERR_NO_DEVICE;
  if (!top_T || !s1 || !s2 || !n_top || !T_out || k < 1 || k > FCCF_TOPK || n_top[0] < 0 || n_top[1] < 0 || n_top[2] < 0 || n_top[0] > k || n_top[1] > k || n_top[2] > k) { ctx->err = "bad argument"; return FCCF_ERR_ARG; }
  int rc = stage_begin(ctx);
  if (rc) return rc;
  cudaStream_t s = ctx->stream;
  std::vector<Work> ws; Batch w = single_batch(ctx, ws);
  launch_init_state(s, w, ctx->G0().d_calls, &ctx->launches);
  HypWS& h = ctx->L0().h;
  CK(cudaMemcpyAsync(h.top_T, top_T, (size_t)3 * k * 64, cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(h.top_s1, s1, (size_t)3 * k * 4, cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(h.top_s2, s2, (size_t)3 * k * 4, cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(&ctx->L0().d_st->n_top[0], n_top, 12, cudaMemcpyHostToDevice, s));
  CK(cudaStreamSynchronize(s));
  launch_fuse(s, ctx->L0().d_st, h.top_T, h.top_s1, h.top_s2, ctx->p.fine_verify_number, k, ctx->itab, &ctx->launches);
  rc = stage_end(ctx);
  for (int i = 0; i < 16; i++) T_out[i] = ctx->L0().h_st->T_final[i];
  return rc;
}

}  // extern "C"
