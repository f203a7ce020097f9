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
#include "fccf_internal.h"

using namespace fccf;

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

// One registration in flight: its own stream, workspaces and result block.  A context owns lane 0
// (the single-pair entry points and the stage entry points) and creates further lanes on demand for
// fccf_register_batch*, which keeps several independent pairs in flight at once: most kernels of one
// registration are single-CTA (order-dependent greedy stages), so concurrent lanes fill the other SMs.
struct Lane {
  cudaStream_t stream = nullptr;
  int cap_pts = 0;
  Arena cloud_arena[2];
  float* d_raw[2] = {nullptr, nullptr};
  CloudWS c[2];
  PipeState* d_st = nullptr;
  PipeState* h_st = nullptr;     // pinned
  Arena hyp_arena;
  HypWS h;
  cudaEvent_t ev[6];
  cudaEvent_t sev[8];
  size_t last_h2d = 0;
  uint64_t launches = 0, l0 = 0;
  bool busy = false; int pair = -1; bool had_h2d = false;
  CallArgs* h_call = nullptr;    // pinned
  cudaGraphExec_t gexec = nullptr;   // the whole pipeline of this lane, captured once per workspace / parameter set
  int graph_launches = 0;        // kernel nodes in the graph
  uint64_t params_epoch = 0;
};

struct fccf_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;   // = lanes[0]->stream
  fccf_params p;
  std::string err;
  uint64_t launches = 0;           // kernels launched outside the lanes (stand-alone stage entry points)
  std::vector<Lane*> lanes;
  int max_lanes = 8;
  bool use_graph = true;
  uint64_t params_epoch = 1;
  int cap_hyp = 1 << 18;
  bool have_run = false;
  float leaf = 0.f;
  // stand-alone scoring
  float *d_sc_s1 = nullptr, *d_sc_s2 = nullptr, *d_sc_T = nullptr, *d_sc_scores = nullptr;
  size_t sc_cap1 = 0, sc_cap2 = 0, sc_capT = 0;
  Arena sc_arena; ScoreWS sc_ws; int sc_cap_hash = 0; ScoreState* d_sc_ss = nullptr; int* d_sc_n = nullptr;
  size_t sc_n2 = 0, sc_nhyp = 0;
  long long* d_sc_best = nullptr;
  // lane-0 aliases used by the stage entry points and the blob reader
  Lane& L0() { return *lanes[0]; }
};

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

static int lane_create(fccf_ctx* ctx, Lane** out) {
  Lane* L = new Lane();
  if (cudaStreamCreateWithFlags(&L->stream, cudaStreamNonBlocking) != cudaSuccess) { delete L; ctx->err = "cudaStreamCreate failed"; return FCCF_ERR_CUDA; }
  if (cudaMalloc(&L->d_st, sizeof(PipeState)) != cudaSuccess || cudaMallocHost(&L->h_st, sizeof(PipeState)) != cudaSuccess ||
      cudaMallocHost(&L->h_call, sizeof(CallArgs)) != cudaSuccess) { delete L; ctx->err = "state allocation failed"; return FCCF_ERR_CUDA; }
  cudaMemset(L->d_st, 0, sizeof(PipeState));
  for (int i = 0; i < 6; i++) cudaEventCreate(&L->ev[i]);
  for (int i = 0; i < 8; i++) cudaEventCreate(&L->sev[i]);
  *out = L;
  return FCCF_OK;
}
static void lane_destroy(Lane* L) {
  if (!L) return;
  if (L->stream) cudaStreamSynchronize(L->stream);
  for (int c = 0; c < 2; c++) { if (L->cloud_arena[c].base) cudaFree(L->cloud_arena[c].base); if (L->d_raw[c]) cudaFree(L->d_raw[c]); }
  if (L->hyp_arena.base) cudaFree(L->hyp_arena.base);
  if (L->d_st) cudaFree(L->d_st);
  if (L->h_st) cudaFreeHost(L->h_st);
  if (L->h_call) cudaFreeHost(L->h_call);
  if (L->gexec) cudaGraphExecDestroy(L->gexec);
  for (int i = 0; i < 6; i++) cudaEventDestroy(L->ev[i]);
  for (int i = 0; i < 8; i++) cudaEventDestroy(L->sev[i]);
  if (L->stream) cudaStreamDestroy(L->stream);
  delete L;
}

static int ensure_capacity(fccf_ctx* ctx, Lane* L, size_t n0, size_t n1) {
  size_t need = std::max(n0, n1);
  if (need < 1024) need = 1024;
  if ((size_t)L->cap_pts >= need) return FCCF_OK;
  if (need > (size_t)1 << 30) { ctx->err = "cloud too large"; return FCCF_ERR_ARG; }
  int cap = (int)((need + 4095) & ~(size_t)4095);
  CK(cudaStreamSynchronize(L->stream));
  if (L->gexec) { cudaGraphExecDestroy(L->gexec); L->gexec = nullptr; }
  for (int c = 0; c < 2; c++) {
    if (L->cloud_arena[c].base) CK(cudaFree(L->cloud_arena[c].base));
    if (L->d_raw[c]) CK(cudaFree(L->d_raw[c]));
    L->cloud_arena[c] = Arena();
    L->d_raw[c] = nullptr;
    size_t bytes = cloud_bytes(cap);
    CK(cudaMalloc(&L->cloud_arena[c].base, bytes));
    L->cloud_arena[c].size = bytes;
    CK(cudaMalloc(&L->d_raw[c], (size_t)cap * 12));
    cloud_carve(L->cloud_arena[c], L->c[c], cap);
    L->c[c].raw = L->d_raw[c];
  }
  // fine-verify hash sized for the leftover cloud (<= cap points)
  if (L->hyp_arena.base) CK(cudaFree(L->hyp_arena.base));
  L->hyp_arena = Arena();
  HypWS& h = L->h;
  size_t ch = (size_t)ctx->cap_hyp, nbh = ch / RS_TILE + 2;
  size_t bytes = 0;
  auto add = [&](size_t x) { bytes += (x + 255) & ~(size_t)255; };
  add(4 * FCCF_MAXMATCH); add(4 * FCCF_MAXMATCH); add(48 * ch); add(32 * ch); add(16 * ch); add(8 * ch);
  add(8 * ch); add(8 * ch); add(4 * ch); add(4 * ch); add(nbh * 1024);
  for (int k = 0; k < 5; k++) add(4 * ch);
  add(8 * ch); add(8 * ch);
  size_t nc = 3 * FCCF_MAXCENTRE, ntop = 3 * FCCF_TOPK;
  add(nc * 32); add(nc * 64); add(nc * 4); add(nc * 4); add(nc * 128); add(nc * 4); add(nc * 4);
  add(ntop * 64); add(ntop * 4); add(ntop * 4); add(ntop * 4);
  size_t fv_bytes = score_ws_layout(nullptr, nullptr, cap, (int)ntop);
  add(fv_bytes);
  bytes += 8192;
  CK(cudaMalloc(&L->hyp_arena.base, bytes));
  L->hyp_arena.size = bytes;
  Arena& a = L->hyp_arena;
  h.cap_hyp = ctx->cap_hyp;
  h.match_cnt = a.take<int>(FCCF_MAXMATCH); h.match_off = a.take<int>(FCCF_MAXMATCH);
  h.hyp_T = a.take<float>(12 * ch); h.hyp_qt = a.take<float>(8 * ch); h.hyp_ax = a.take<float>(4 * ch); h.hyp_an = a.take<double>(ch);
  h.ckeyA = a.take<u64>(ch); h.ckeyB = a.take<u64>(ch); h.cidxA = a.take<u32>(ch); h.cidxB = a.take<u32>(ch); h.chist = a.take<u32>(nbh * 256);
  h.c_state = a.take<int>(ch); h.c_size = a.take<int>(ch); h.c_seeds = a.take<int>(ch); h.c_perm = a.take<int>(ch); h.c_key = a.take<int>(ch);
  h.c_members = a.take<int>(2 * ch); h.c_mdist = a.take<float>(2 * ch);
  h.centre = a.take<float>(nc * 8); h.qv_T = a.take<float>(nc * 16); h.qv_score = a.take<float>(nc); h.qv_npair = a.take<int>(nc);
  h.qv_pairs = a.take<int>(nc * 32); h.qv_iters = a.take<int>(nc); h.rank_perm = a.take<int>(nc);
  h.top_T = a.take<float>(ntop * 16); h.top_s1 = a.take<float>(ntop); h.top_s2 = a.take<float>(ntop); h.top_centre = a.take<int>(ntop);
  score_ws_layout(&h.fv, a.take<char>(fv_bytes), cap, (int)ntop);
  L->cap_pts = cap;
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

fccf_ctx* fccf_create(int device, const fccf_params* params) {
  // Concurrent lanes need one hardware work queue each; the driver's default is 8 per device, which
  // serialises lanes that share a queue (measured: 16 lanes 1.40 -> 0.63 ms/registration with 32).
  // Only effective if the CUDA context of this process has not been created yet; never overrides.
  setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0);
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev) return nullptr;
  if (cudaSetDevice(device) != cudaSuccess) return nullptr;
  fccf_ctx* ctx = new fccf_ctx();
  ctx->device = device;
  if (params) ctx->p = *params; else fccf_default_params(&ctx->p);
  if (ctx->p.batch_lanes > 0) ctx->max_lanes = ctx->p.batch_lanes > 256 ? 256 : ctx->p.batch_lanes;
  if (const char* e = getenv("FCCF_NO_GRAPH")) ctx->use_graph = !(e[0] == '1');
  score_init_attributes();
  Lane* L = nullptr;
  if (lane_create(ctx, &L) != FCCF_OK) { delete ctx; return nullptr; }
  ctx->lanes.push_back(L);
  ctx->stream = L->stream;
  return ctx;
}

void fccf_destroy(fccf_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  for (Lane* L : ctx->lanes) lane_destroy(L);
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

const char* fccf_last_error(const fccf_ctx* ctx) { return ctx ? ctx->err.c_str() : "no context (no usable CUDA device)"; }
int fccf_set_params(fccf_ctx* ctx, const fccf_params* params) {
  if (!ctx || !params) return FCCF_ERR_ARG;
  ctx->p = *params; ctx->params_epoch++;     // captured graphs hold the old values: re-capture lazily
  return FCCF_OK;
}
uint64_t fccf_launch_count(const fccf_ctx* ctx) {
  if (!ctx) return 0;
  uint64_t n = ctx->launches;
  for (const Lane* L : ctx->lanes) n += L->launches;
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

static Work make_work(fccf_ctx* ctx, Lane* L, float leaf) {
  Work w; w.c[0] = L->c[0]; w.c[1] = L->c[1]; w.st = L->d_st; w.p = ctx->p; w.cuts = make_angle_cuts(ctx->p); w.leaf = leaf;
  return w;
}

static int check_status(fccf_ctx* ctx, int st) {
  if (st == 0) return FCCF_OK;
  char b[256];
  snprintf(b, sizeof b, "device status 0x%x:%s%s%s%s%s", st, (st & ST_OCT_DEPTH) ? " octree deeper than the 32-bit Morton key" : "",
           (st & ST_HYP_OVERFLOW) ? " hypothesis pool capacity exceeded" : "", (st & ST_CENTRE_OVERFLOW) ? " cluster centre capacity exceeded" : "",
           (st & ST_HASH_FULL) ? " fine-verify lattice out of range / hash full" : "", (st & ST_KEYBITS) ? " sort key too wide" : "");
  ctx->err = b;
  return FCCF_ERR_CAPACITY;
}

// Every kernel of main() + computer_transform_guess and the read-back of the result block, in stream
// order on the lane's stream.  All sizes are device-side and the per-call values (point counts, leaf,
// raw-cloud pointers) are read from st->call, so the sequence is identical for every call on this lane:
// it is captured into a CUDA graph once and replayed afterwards.
static int lane_pipeline(fccf_ctx* ctx, Lane* L, uint64_t* launches, bool capturing) {
  cudaStream_t s = L->stream;
  // inside a capture a plain cudaEventRecord is only a dependency; the External flag makes a timing node
  auto rec = [&](cudaEvent_t e) { return capturing ? cudaEventRecordWithFlags(e, s, cudaEventRecordExternal) : cudaEventRecord(e, s); };
  Work w = make_work(ctx, L, 0.f);
  launch_init_state(s, L->d_st, launches);
  launch_voxelgrid(s, w, 0, 2, launches);            // main(): FCCF.cpp:1668-1678
  CK(rec(L->ev[2]));
  launch_voxelgrid(s, w, 1, 2, launches);            // FCCF.cpp:1377-1387
  CK(rec(L->sev[0]));
  launch_planes(s, w, 2, 1, launches);               // FCCF.cpp:1400-1401
  CK(rec(L->sev[1]));
  launch_hypotheses(s, w, L->h, launches);           // FCCF.cpp:1406-1427, 1439-1462
  CK(rec(L->sev[2]));
  launch_cluster(s, w, L->h, launches);              // FCCF.cpp:1464-1466
  CK(rec(L->sev[3]));
  launch_quick_verify(s, w, L->h, launches);         // FCCF.cpp:1468-1494
  CK(rec(L->sev[4]));
  launch_fine_verify_fuse(s, w, L->h, launches);     // FCCF.cpp:1499-1606
  CK(rec(L->ev[3]));
  CK(cudaMemcpyAsync(L->h_st, L->d_st, sizeof(PipeState), cudaMemcpyDeviceToHost, s));
  return FCCF_OK;
}

// Enqueues one whole registration on the lane's stream (no host synchronisation): the per-call block,
// optional H2D of both raw clouds, then the pipeline (graph replay).  tar/src: host pointers (host_in)
// or device pointers.
static int lane_enqueue(fccf_ctx* ctx, Lane* L, const float* src, size_t n_src, const float* tar, size_t n_tar, float leaf, bool host_in) {
  cudaStream_t s = L->stream;
  L->l0 = L->launches; L->had_h2d = host_in;
  CK(cudaEventRecord(L->ev[0], s));
  if (host_in) {
    if (n_tar) CK(cudaMemcpyAsync(L->d_raw[0], tar, n_tar * 12, cudaMemcpyHostToDevice, s));
    if (n_src) CK(cudaMemcpyAsync(L->d_raw[1], src, n_src * 12, cudaMemcpyHostToDevice, s));
    L->last_h2d = (n_tar + n_src) * 12;
  } else L->last_h2d = 0;
  L->h_call->n0 = (int)n_tar; L->h_call->n1 = (int)n_src; L->h_call->leaf = leaf; L->h_call->pad = 0;
  L->h_call->raw[0] = host_in ? L->d_raw[0] : tar; L->h_call->raw[1] = host_in ? L->d_raw[1] : src;
  CK(cudaMemcpyAsync(&L->d_st->call, L->h_call, sizeof(CallArgs), cudaMemcpyHostToDevice, s));
  CK(cudaEventRecord(L->ev[1], s));
  if (ctx->use_graph) {
    if (!L->gexec || L->params_epoch != ctx->params_epoch) {
      if (L->gexec) { cudaGraphExecDestroy(L->gexec); L->gexec = nullptr; }
      cudaGraph_t g = nullptr;
      uint64_t cnt = 0;
      CK(cudaStreamBeginCapture(s, cudaStreamCaptureModeRelaxed));
      int rc = lane_pipeline(ctx, L, &cnt, true);
      cudaError_t e = cudaStreamEndCapture(s, &g);
      if (rc != FCCF_OK || e != cudaSuccess || !g) { if (g) cudaGraphDestroy(g); ctx->err = std::string("graph capture failed: ") + cudaGetErrorString(e); cudaGetLastError(); return FCCF_ERR_CUDA; }
      e = cudaGraphInstantiate(&L->gexec, g, 0);
      cudaGraphDestroy(g);
      if (e != cudaSuccess) { L->gexec = nullptr; ctx->err = std::string("graph instantiation failed: ") + cudaGetErrorString(e); return FCCF_ERR_CUDA; }
      L->graph_launches = (int)cnt; L->params_epoch = ctx->params_epoch;
    }
    CK(cudaGraphLaunch(L->gexec, s));
    L->launches += (uint64_t)L->graph_launches;
  } else {
    int rc = lane_pipeline(ctx, L, &L->launches, false);
    if (rc) return rc;
  }
  CK(cudaEventRecord(L->ev[4], s));
  L->busy = true;
  return FCCF_OK;
}

// Waits for the lane's registration, returns the matrix and (optionally) its device timings.
static int lane_finish(fccf_ctx* ctx, Lane* L, float T_out[16], fccf_timing* tm) {
  CK(cudaStreamSynchronize(L->stream));
  CK(cudaGetLastError());
  L->busy = false;
  for (int i = 0; i < 16; i++) T_out[i] = L->h_st->T_final[i];
  if (tm) {
    memset(tm, 0, sizeof *tm);
    if (L->had_h2d) cudaEventElapsedTime(&tm->h2d_ms, L->ev[0], L->ev[1]);
    cudaEventElapsedTime(&tm->downsample_ms, L->ev[1], L->ev[2]);
    cudaEventElapsedTime(&tm->pipeline_ms, L->ev[2], L->ev[3]);
    cudaEventElapsedTime(&tm->d2h_ms, L->ev[3], L->ev[4]);
    cudaEventElapsedTime(&tm->total_ms, L->ev[0], L->ev[4]);
    tm->n_launches = (int)(L->launches - L->l0);
    tm->h2d_bytes = (unsigned long long)L->last_h2d;
    tm->d2h_bytes = sizeof(PipeState);
    tm->stage_ms[0] = tm->downsample_ms;
    cudaEventElapsedTime(&tm->stage_ms[1], L->ev[2], L->sev[0]);
    for (int k = 0; k < 4; k++) cudaEventElapsedTime(&tm->stage_ms[2 + k], L->sev[k], L->sev[k + 1]);
    cudaEventElapsedTime(&tm->stage_ms[6], L->sev[4], L->ev[3]);
  }
  return check_status(ctx, L->h_st->status);
}

static int register_one(fccf_ctx* ctx, const float* src, size_t n_src, const float* tar, size_t n_tar, float leaf, float T_out[16], fccf_timing* timing, bool host_in) {
  if (!ctx) return FCCF_ERR_NO_DEVICE;
  if (!T_out || (!src && n_src) || (!tar && n_tar) || !(leaf > 0.f)) { ctx->err = "bad argument"; return FCCF_ERR_ARG; }
  CK(cudaSetDevice(ctx->device));
  Lane* L = ctx->lanes[0];
  int rc = ensure_capacity(ctx, L, n_tar, n_src);
  if (rc) return rc;
  rc = lane_enqueue(ctx, L, src, n_src, tar, n_tar, leaf, host_in);
  if (rc) return rc;
  rc = lane_finish(ctx, L, T_out, timing);
  ctx->have_run = true; ctx->leaf = leaf;
  return rc;
}

// Batch of independent pairs: up to max_lanes registrations in flight, one lane (stream + workspace)
// each; pair b waits only for the lane it reuses.  timing: device times summed over the pairs, and
// total_ms = host wall clock of the whole batch.
static int register_many(fccf_ctx* ctx, int n_pairs, const float* const* src, const size_t* n_src, const float* const* tar, const size_t* n_tar,
                         float leaf, float* T_out, fccf_timing* timing, bool host_in) {
  if (!ctx) return FCCF_ERR_NO_DEVICE;
  if (n_pairs < 0 || (n_pairs && (!src || !tar || !n_src || !n_tar || !T_out)) || !(leaf > 0.f)) { ctx->err = "bad argument"; return FCCF_ERR_ARG; }
  CK(cudaSetDevice(ctx->device));
  fccf_timing acc; memset(&acc, 0, sizeof acc);
  if (n_pairs == 0) { if (timing) *timing = acc; return FCCF_OK; }
  int nl = std::min(ctx->max_lanes, n_pairs);
  size_t nmax = 0;
  for (int b = 0; b < n_pairs; b++) { if ((!src[b] && n_src[b]) || (!tar[b] && n_tar[b])) { ctx->err = "bad argument"; return FCCF_ERR_ARG; } nmax = std::max(nmax, std::max(n_src[b], n_tar[b])); }
  while ((int)ctx->lanes.size() < nl) { Lane* L = nullptr; int rc = lane_create(ctx, &L); if (rc) return rc; ctx->lanes.push_back(L); }
  for (int l = 0; l < nl; l++) { int rc = ensure_capacity(ctx, ctx->lanes[l], nmax, nmax); if (rc) return rc; }
  // device time of the whole batch: an event on lane 0 before the first enqueue, and one on lane 0
  // after it has waited for the last registration of every lane
  Lane* Z = ctx->lanes[0];
  auto t0 = std::chrono::steady_clock::now();
  CK(cudaEventRecord(Z->sev[6], Z->stream));
  int worst = FCCF_OK;
  float enq_ms = 0.f;
  auto collect = [&](Lane* L) -> int {
    fccf_timing t;
    int rc = lane_finish(ctx, L, T_out + 16 * (size_t)L->pair, &t);
    acc.h2d_ms += t.h2d_ms; acc.downsample_ms += t.downsample_ms; acc.pipeline_ms += t.pipeline_ms; acc.d2h_ms += t.d2h_ms; acc.n_launches += t.n_launches;
    acc.h2d_bytes += t.h2d_bytes; acc.d2h_bytes += t.d2h_bytes; for (int k = 0; k < 8; k++) acc.stage_ms[k] += t.stage_ms[k];
    return rc;
  };
  for (int b = 0; b < n_pairs; b++) {
    Lane* L = ctx->lanes[b % nl];
    if (L->busy) { int rc = collect(L); if (rc == FCCF_ERR_CUDA || rc == FCCF_ERR_ARG) return rc; if (rc) worst = rc; }
    L->pair = b;
    auto te0 = std::chrono::steady_clock::now();
    int rc = lane_enqueue(ctx, L, src[b], n_src[b], tar[b], n_tar[b], leaf, host_in);
    enq_ms += std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - te0).count();
    if (rc) return rc;
  }
  if (getenv("FCCF_DEBUG_TIMING")) fprintf(stderr, "[fccf] batch of %d: host enqueue %.3f ms total (%.3f ms/pair)\n", n_pairs, enq_ms, enq_ms / n_pairs);
  for (int l = 1; l < nl; l++) if (ctx->lanes[l]->busy) CK(cudaStreamWaitEvent(Z->stream, ctx->lanes[l]->ev[4], 0));
  CK(cudaEventRecord(Z->sev[7], Z->stream));
  for (int l = 0; l < nl; l++) if (ctx->lanes[l]->busy) { int rc = collect(ctx->lanes[l]); if (rc == FCCF_ERR_CUDA || rc == FCCF_ERR_ARG) return rc; if (rc) worst = rc; }
  CK(cudaEventSynchronize(Z->sev[7]));
  cudaEventElapsedTime(&acc.total_ms, Z->sev[6], Z->sev[7]);
  acc.stage_ms[7] = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
  if (timing) *timing = acc;
  ctx->have_run = true; ctx->leaf = leaf;    // blobs: the last pair that ran on lane 0
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
  int rc = ensure_capacity(ctx, ctx->lanes[0], n, 0);
  if (rc) return rc;
  cudaStream_t s = ctx->stream;
  if (n) CK(cudaMemcpyAsync(ctx->L0().d_raw[0], xyz, n * 12, cudaMemcpyHostToDevice, s));
  Work w = make_work(ctx, ctx->lanes[0], leaf);
  { CallArgs* hc = ctx->L0().h_call; hc->n0 = (int)n; hc->n1 = 0; hc->leaf = leaf; hc->pad = 0; hc->raw[0] = ctx->L0().d_raw[0]; hc->raw[1] = ctx->L0().d_raw[1];
    CK(cudaMemcpyAsync(&ctx->L0().d_st->call, hc, sizeof(CallArgs), cudaMemcpyHostToDevice, s)); }
  launch_init_state(s, ctx->L0().d_st, &ctx->launches);
  launch_voxelgrid(s, w, 0, 1, &ctx->launches);
  CK(cudaMemcpyAsync(ctx->L0().h_st, ctx->L0().d_st, sizeof(PipeState), cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
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
  int rc = ensure_capacity(ctx, ctx->lanes[0], n, 0);
  if (rc) return rc;
  cudaStream_t s = ctx->stream;
  Work w = make_work(ctx, ctx->lanes[0], 1.0f);
  { CallArgs* hc = ctx->L0().h_call; hc->n0 = 0; hc->n1 = 0; hc->leaf = 1.0f; hc->pad = 0; hc->raw[0] = ctx->L0().d_raw[0]; hc->raw[1] = ctx->L0().d_raw[1];
    CK(cudaMemcpyAsync(&ctx->L0().d_st->call, hc, sizeof(CallArgs), cudaMemcpyHostToDevice, s)); }
  launch_init_state(s, ctx->L0().d_st, &ctx->launches);
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
  launch_score_build(s, ctx->p, ctx->d_sc_s1, ctx->d_sc_n, ctx->d_sc_n + 1, ctx->sc_ws, &ctx->launches);
  ctx->sc_n2 = n2; ctx->sc_nhyp = n_hyp;
  return FCCF_OK;
}

int fccf_score_hypotheses(fccf_ctx* ctx, const float* T, size_t n_hyp, const float* s1_xyz, size_t n1, const float* s2_xyz, size_t n2, float* scores) {
  if (!ctx) return FCCF_ERR_NO_DEVICE;
  if ((!T && n_hyp) || !scores) { ctx->err = "bad argument"; return FCCF_ERR_ARG; }
  CK(cudaSetDevice(ctx->device));
  int rc = score_prepare(ctx, T, n_hyp, s1_xyz, n1, s2_xyz, n2);
  if (rc) return rc;
  launch_score_list(ctx->stream, ctx->p, ctx->d_sc_T, (int)n_hyp, ctx->d_sc_s2, ctx->sc_ws, ctx->d_sc_scores, &ctx->launches);
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
  for (int w = 0; w < 3; w++) launch_score_list(s, ctx->p, ctx->d_sc_T, (int)n_hyp, ctx->d_sc_s2, ctx->sc_ws, ctx->d_sc_scores, &ctx->launches);
  CK(cudaStreamSynchronize(s));
  CK(cudaEventRecord(ctx->L0().ev[0], s));
  for (int r = 0; r < repeat; r++) launch_score_list(s, ctx->p, ctx->d_sc_T, (int)n_hyp, ctx->d_sc_s2, ctx->sc_ws, ctx->d_sc_scores, &ctx->launches);
  CK(cudaEventRecord(ctx->L0().ev[1], s));
  CK(cudaStreamSynchronize(s));
  float ms = 0; cudaEventElapsedTime(&ms, ctx->L0().ev[0], ctx->L0().ev[1]);
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
  CK(cudaSetDevice(ctx->device));
  int* d_rows = nullptr; int* d_n = nullptr;
  CK(cudaMalloc(&d_rows, std::max<size_t>(cap_rows, 1) * 20));
  CK(cudaMalloc(&d_n, 4));
  CK(cudaMemsetAsync(d_n, 0, 4, ctx->stream));
  launch_score_dump(ctx->stream, ctx->p, ctx->d_sc_T + 16 * hyp, ctx->d_sc_s2, ctx->sc_ws, d_rows, (int)cap_rows, d_n, &ctx->launches);
  int n = 0;
  CK(cudaMemcpyAsync(&n, d_n, 4, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  *n_rows = (size_t)n;
  size_t m = std::min<size_t>((size_t)n, cap_rows);
  if (m) CK(cudaMemcpy(rows, d_rows, m * 20, cudaMemcpyDeviceToHost));
  cudaFree(d_rows); cudaFree(d_n);
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
  CK(cudaMalloc(&dT, n_hyp * 64)); CK(cudaMalloc(&dp1, p1.size() * 4)); CK(cudaMalloc(&dp2, p2.size() * 4)); CK(cudaMalloc(&dsc, n_hyp * 4));
  CK(cudaMalloc(&dnp, n_hyp * 4)); CK(cudaMalloc(&dpr, n_hyp * 128)); CK(cudaMalloc(&dit, n_hyp * 4));
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
  cudaFree(dT); cudaFree(dp1); cudaFree(dp2); cudaFree(dsc); cudaFree(dnp); cudaFree(dpr); cudaFree(dit);
  CK(cudaGetLastError());
  return FCCF_OK;
}

// ---------------------------------------------------------------------------------------------
// stage intermediates of the last run
// ---------------------------------------------------------------------------------------------
int fccf_debug_blob(fccf_ctx* ctx, const char* name_c, void* dst, size_t cap_bytes, size_t* bytes, int* dtype) {
  if (!ctx) return FCCF_ERR_NO_DEVICE;
  if (!name_c || !ctx->have_run) { ctx->err = "no run to inspect"; return FCCF_ERR_ARG; }
  CK(cudaSetDevice(ctx->device));
  CK(cudaStreamSynchronize(ctx->stream));
  std::string name(name_c);
  const PipeState& st = *ctx->L0().h_st;
  std::vector<char> out; int dt = FCCF_F32;
  auto fetch = [&](const void* dptr, size_t nbytes) -> std::vector<char> {
    std::vector<char> v(nbytes);
    if (nbytes) cudaMemcpy(v.data(), dptr, nbytes, cudaMemcpyDeviceToHost);
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
  else if (stem == "top_T" && ti >= 0) { out = fetch(ctx->L0().h.top_T + (size_t)ti * FCCF_TOPK * 16, (size_t)st.n_top[ti] * 64); dt = FCCF_F32; }
  else if (stem == "top_s1" && ti >= 0) { out = fetch(ctx->L0().h.top_s1 + (size_t)ti * FCCF_TOPK, (size_t)st.n_top[ti] * 4); dt = FCCF_F32; }
  else if (stem == "top_s2" && ti >= 0) { out = fetch(ctx->L0().h.top_s2 + (size_t)ti * FCCF_TOPK, (size_t)st.n_top[ti] * 4); dt = FCCF_F32; }
  else if (stem == "top_centre" && ti >= 0) { out = fetch(ctx->L0().h.top_centre + (size_t)ti * FCCF_TOPK, (size_t)st.n_top[ti] * 4); dt = FCCF_I32; }
  else if ((stem == "fv_counts" || stem == "fv_off") && ti >= 0) {
    // per-voxel (s,t) of every fine-verified hypothesis of this type, rows sorted lexicographically
    ScoreWS ws = ctx->L0().h.fv; ws.ss = &ctx->L0().d_st->fv; ws.status = &ctx->L0().d_st->status;
    int cap_rows = std::max(st.fv.n_occ, 1);
    int* d_rows = nullptr; int* d_n = nullptr;
    CK(cudaMalloc(&d_rows, (size_t)cap_rows * 20)); CK(cudaMalloc(&d_n, 4));
    std::vector<int> all, off;
    for (int k = 0; k < st.n_top[ti]; k++) {
      CK(cudaMemsetAsync(d_n, 0, 4, ctx->stream));
      launch_score_dump(ctx->stream, ctx->p, ctx->L0().h.top_T + ((size_t)ti * FCCF_TOPK + k) * 16, ctx->L0().c[1].sub, ws, d_rows, cap_rows, d_n, &ctx->launches);
      int n = 0; CK(cudaMemcpyAsync(&n, d_n, 4, cudaMemcpyDeviceToHost, ctx->stream)); CK(cudaStreamSynchronize(ctx->stream));
      n = std::min(n, cap_rows);
      std::vector<int> rows(5 * (size_t)n); if (n) CK(cudaMemcpy(rows.data(), d_rows, (size_t)n * 20, cudaMemcpyDeviceToHost));
      std::vector<int> order(n); for (int i = 0; i < n; i++) order[i] = i;
      std::sort(order.begin(), order.end(), [&](int a, int b) { return std::lexicographical_compare(&rows[5 * a], &rows[5 * a + 3], &rows[5 * b], &rows[5 * b + 3]); });
      off.push_back((int)all.size() / 5);
      for (int i = 0; i < n; i++) for (int a = 0; a < 5; a++) all.push_back(rows[5 * order[i] + a]);
    }
    off.push_back((int)all.size() / 5);
    cudaFree(d_rows); cudaFree(d_n);
    if (stem == "fv_counts") put(all.data(), all.size() * 4, FCCF_I32); else put(off.data(), off.size() * 4, FCCF_I32);
  }
  else if (name == "type_best") put(st.type_best, sizeof st.type_best, FCCF_F32);
  else if (name == "prof") put(st.prof, sizeof st.prof, FCCF_I64);
  else if (name == "final_T") put(st.T_final, 64, FCCF_F32);
  else ok = false;
  if (!ok) { ctx->err = "unknown blob: " + name; return FCCF_ERR_ARG; }
  if (bytes) *bytes = out.size();
  if (dtype) *dtype = dt;
  if (dst) { if (out.size() > cap_bytes) { ctx->err = "blob buffer too small"; return FCCF_ERR_ARG; } if (!out.empty()) memcpy(dst, out.data(), out.size()); }
  CK(cudaGetLastError());
  return FCCF_OK;
}

}  // extern "C"
