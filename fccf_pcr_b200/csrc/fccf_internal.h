// fccf_internal.h — device workspace layout and kernel launchers of libfccf (not part of the ABI).
#pragma once
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>
#include "../../include/fccf.h"

namespace fccf {

typedef unsigned long long u64;
typedef unsigned int u32;

#define FCCF_MAXF 16          // planes kept per cloud (FCCF.cpp:141,670: 16)
#define FCCF_MAXBASE 120      // C(16,2)
#define FCCF_MAXMATCH (FCCF_MAXBASE * FCCF_MAXBASE)
#define FCCF_MAXCENTRE 256    // per roughness type
#define FCCF_TOPK 256         // capacity: fine-verified hypotheses per type (reference default: 4; 256 = every centre, the
                              // exhaustive-scoring mode of SURVEY.md f2: fine_verify_number >= FCCF_MAXCENTRE)
// slots per type actually used: the top arrays are laid out [3][topk] with the run-time stride
inline int fccf_topk(const fccf_params& p) { int k = (int)p.fine_verify_number; return k < 1 ? 1 : (k > FCCF_TOPK ? FCCF_TOPK : k); }

// radix sort tile
#define RS_T 256
#define RS_I 8
#define RS_TILE (RS_T * RS_I)

// status bits written by kernels (checked by the host after the final synchronise)
enum { ST_OCT_DEPTH = 1, ST_HYP_OVERFLOW = 2, ST_KEYBITS = 4, ST_CENTRE_OVERFLOW = 8, ST_HASH_FULL = 16, ST_CLUSTER_MEMBERS = 32,
       ST_VG_FAST_MISS = 64,     // not an error: the cluster VoxelGrid (voxelgrid_fast.cu) could not hold a cloud, the host re-runs the generic kernels
       ST_SORT_MISS = 128 };     // not an error: a sequence captured without radix passes (small sorts only) met a longer list, the host re-runs the full one
#define RS_SMALL 4096          // sort jobs of up to this many keys are done by one CTA in shared memory (sort.cu)

struct VGState {
  int n_in, n_finite;
  int mn[3], mx[3];          // ordered-int encoded float bounds
  int bail;                  // pcl::VoxelGrid int32 overflow bail-out (output = input)
  int minb[3];
  long long div[3];
  long long total;           // number of cells (or n_in when bailing out) = key of non-finite points
  int nbits;
  float inv;
  int n_out;                 // number of output points
  int pad;                   // generic path: number of cells left to the warp-per-cell centroid kernel (voxelgrid.cu: VG_BIG)
};
struct OctState {
  double mn[3], mx[3];
  int depth, nbits;
  int n;                     // points
  int V, Vp, S, F1;
  float cc[3];               // whole-cloud centroid (FCCF.cpp:473)
  int pad;
};
struct FaceTable {
  int F, pad;
  float plane[FCCF_MAXF][8]; // cx cy cz nx ny nz size nvox
  double theta[FCCF_MAXF];
  int id[FCCF_MAXF];         // stage-1 face id
};
struct BaseTable {
  int B;
  int i[FCCF_MAXBASE], j[FCCF_MAXBASE], type[FCCF_MAXBASE];
  float angle[FCCF_MAXBASE];
};
struct ScoreState {          // fine-verify lattice + static voxel table set-up (device)
  double mn[3];              // lattice origin: voxel faces lie at mn + k*res (mn = first static point - res)
  int lmin[3], lmax[3];      // bounding box of the static cloud in lattice cells
  int dims[3];
  float lo_f[3], hi_f[3];    // a float q lies inside the box iff lo_f <= q < hi_f
  int n1, n2, n_occ, cap_eff, nbits, n_keys, mode, pad;
  int tickets[4];
};
struct CallArgs {            // per-call values of one lane, copied to the device before the (graph-replayed) pipeline runs
  int n0, n1;                // raw point counts: cloud "1" (TAR file), cloud "2" (SRC file)
  float leaf; int pad;
  const float* raw[2];       // raw clouds (device)
};
struct PipeState {
  CallArgs call;
  VGState vg[2][2];          // [stage: 0 main(), 1 computer_transform_guess][cloud: 0 = "1" (TAR file), 1 = "2" (SRC file)]
  OctState oct[2];
  FaceTable ft[2];
  BaseTable base[2];
  int n_match;
  int n_hyp[3], hyp_off[4];
  int cluster_num[3];
  int n_centre[3];
  int n_seeds[3];
  int n_top[3];
  int tickets[64];
  int status;
  ScoreState fv;
  float type_best[3][13];    // per type: best score + 3x4
  float T_final[16];
  long long prof[32];        // clock64() marks of the single-CTA kernels (debug blob "prof")
};

struct SortJob {
  const void* kin; void* kout; const u32* vin; u32* vout;   // keys: u32 or u64 (launch_sort's key_bytes); pass 0 uses the element index as value
  const int* n; const int* nbits;
  u32* hist;     // [(nblocks_cap + 1) * 256]
  int* ticket;
  int* miss;     // status word that takes ST_SORT_MISS when a pass-less launch meets a job beyond RS_SMALL (nullptr: none)
};
struct SortJobs { SortJob j[3]; };
struct SegJob {
  const void* keys; const int* n; int* seg_start; int* nseg; u32* blk; int* ticket;
};
struct SegJobs { SegJob j[3]; };

// ---- batched launches -------------------------------------------------------------------------
// A launch covers G "lanes" (independent registrations in flight): grid.z = G and every CTA reads its
// lane's argument block from a device table, A = table[blockIdx.z].  The table is filled on the host
// while the launch sequence is issued: in immediate mode each block is copied right before its kernel
// (stand-alone stage entry points); while a CUDA graph is being captured nothing is copied and the
// whole table is uploaded once after the capture (the graph only holds the table's device address).
struct ArgTable {
  char* h = nullptr;         // pinned host mirror
  char* d = nullptr;
  size_t cap = 0, off = 0;
  cudaStream_t stream = nullptr;
  bool immediate = true;
  bool overflow = false;
  template <class T> const T* put(const T* v, int G) {
    size_t bytes = (sizeof(T) * (size_t)G + 255) & ~(size_t)255;
    if (off + bytes > cap) { overflow = true; return (const T*)d; }
    memcpy(h + off, v, sizeof(T) * (size_t)G);
    if (immediate) cudaMemcpyAsync(d + off, h + off, sizeof(T) * (size_t)G, cudaMemcpyHostToDevice, stream);
    const T* r = (const T*)(d + off);
    off += bytes;
    return r;
  }
};

// grid.x of a launch whose CTAs stride over up to nb_cap blocks of work per job, when G lanes x njobs
// jobs share the launch: the element counts are device-side, so grids are sized for capacity, and a
// batched launch must not drown the GPU in CTAs that only find out that they have nothing to do.
inline int grid_x(int nb_cap, int G, int njobs = 1, int budget = 4096) {
  int lim = budget / (G * njobs);
  if (lim < 8) lim = 8;
  if (nb_cap < 1) nb_cap = 1;
  return nb_cap < lim ? nb_cap : lim;
}

// Stable LSD radix sort of (key,value) pairs with a device-side element count and key width;
// `np` passes of ceil(nbits/np) <= 8 bits.  Result ends in (kout,vout) of the last pass; the
// launcher ping-pongs between the two buffer sets given in `a` and `b` (np even: result in a).
// jobs_ab / jobs_ba: device tables of G entries.
// small_only: only the one-CTA sort is launched (sequences captured for lists known to be short; longer ones raise ST_SORT_MISS)
void launch_sort(cudaStream_t s, const SortJobs* jobs_ab, const SortJobs* jobs_ba, int njobs, int G, int cap, int np, int key_bytes, uint64_t* launches, bool small_only = false);
// segment heads of a sorted key array: seg_start[0..nseg], nseg
void launch_segments(cudaStream_t s, const SegJobs* jobs, int njobs, int G, int cap, int key_bytes, uint64_t* launches);

// per-cloud device buffers
struct CloudWS {
  int cap;                   // capacity in points
  const float* raw;          // raw cloud (device)
  u64 *keyA, *keyB; u32 *idxA, *idxB; u32 *hist, *segblk;
  float* vg_xyz[2]; long long* vg_cell[2]; int* vg_cnt[2];
  int* seg_start;
  int* vox_start;            // V+1
  float* vox_rec;            // V x 12: cx cy cz nx ny nz curv count flag kx ky kz
  int* vox_aux;              // V: planar rank / leftover offset
  float* pvox;               // Vp x 8: cx cy cz nx ny nz size voxel-index
  float* sub;                // S x 3 leftover cloud
  int *grow_label, *merge_label, *next, *fhead, *ftail, *fnvox, *falloc, *fperm, *fkey;
  float* fstat;              // F1 x 16: avg(7) + sums(7)
  int* face_vox;             // member lists of the selected faces (debug)
  int* face_off;
};
// Cosine cuts of the angle thresholds (see fccf_dev.cuh: angle_lt / angle_not_gt): found on the host by
// bisection over the float bit patterns with compute_normal_angel's own expression.
struct AngleCuts { float third_lt, qv_lt, cluster_lt, grow1_le, grow2_le; };
float angle_cut(float thr_deg, bool strict);   // smallest float c with theta(c) < thr (strict) or <= thr
AngleCuts make_angle_cuts(const fccf_params& p);


struct ScoreWS {
  int cap_points, cap_hash, t_rows;
  u64 *keyA, *keyB; u32 *idxA, *idxB; u32 *hist, *segblk; int* seg_start;   // voxel-key sort of the static cloud
  u32* vkey; int* s_cnt;     // per occupied static voxel (ascending compact key): key, static point count
  u32 *hkey, *hid;           // open-addressing hash: compact key -> voxel id
  unsigned short* dense;     // dense cell -> voxel id table (bounding boxes of <= 32768 cells)
  int* t_cnt;                // t_rows x cap_points moving-point counters (when the table exceeds shared memory)
  ScoreState* ss; int* status;
};
// carves a ScoreWS out of `base` (nullptr: only computes the size); returns the bytes needed
size_t score_ws_layout(ScoreWS* ws, char* base, int cap_points, int t_rows);

// hypotheses / clustering / verification workspace
struct HypWS {
  int cap_hyp;
  int* match_cnt;            // FCCF_MAXMATCH: hypotheses of (b1,b2), 0 if no match
  int* match_off;            // offset inside its type pool
  float* hyp_T;              // cap_hyp x 12 (3x4 row-major), pools concatenated: type0, type1, type2
  float* hyp_qt;             // cap_hyp x 8: qw qx qy qz tx ty tz pad
  float* hyp_ax;             // cap_hyp x 4: rotated x axis
  double* hyp_an;            // cap_hyp: its norm (double), for the cosine form of the 2-degree test
  u64 *ckeyA, *ckeyB; u32 *cidxA, *cidxB; u32* chist;
  int* c_state;              // per hypothesis: seed state
  int* c_size;               // per hypothesis: cluster size if seed
  int* c_seeds;              // compacted seed ids (per type region)
  int* c_perm; int* c_key;
  int* c_members; float* c_mdist;   // scratch for emitted clusters
  int* c_nbl; int* c_deg;           // neighbour lists / counts of the shared-memory clustering path
  float* centre;             // 3 x FCCF_MAXCENTRE x 8 (qw qx qy qz tx ty tz pad)
  float* qv_T;               // 3 x FCCF_MAXCENTRE x 16 refined
  float* qv_score;           // 3 x FCCF_MAXCENTRE
  int* qv_npair; int* qv_pairs; int* qv_iters;   // per centre: count, 16 x 2, iterations
  int* rank_perm;            // 3 x FCCF_MAXCENTRE
  float* top_T;              // 3 x FCCF_TOPK x 16
  float* top_s1; float* top_s2; int* top_centre;
  ScoreWS fv;                // fine verify: static voxel table of cloud 1's leftover points
};
struct Work {                 // one lane: its workspaces and state block
  CloudWS c[2];
  HypWS h;
  PipeState* st;
};
struct Batch {                // the lanes one batched launch sequence covers
  const Work* w; int G;
  ArgTable* tab;
  fccf_params p;
  AngleCuts cuts;
  // inside a stream capture: a second capture stream and an event pair for work that may run beside the main
  // sequence (nullptr outside captures: everything is launched in order on the one stream)
  cudaStream_t side = nullptr; cudaEvent_t side_fork = nullptr, side_join = nullptr;
  bool lean = false;     // the hypothesis and fine-verify sorts are known to be short: no radix pass launches behind their small sort
};


// ---- launches -------------------------------------------------------------------------------------
// Every kernel of the library starts with FCCF_PDL_ENTER(): it lets the grids that depend on it be scheduled at
// once (griddepcontrol.launch_dependents) and then waits until everything it depends on has completed and is
// visible (griddepcontrol.wait).  A launch made while `fccf_pdl_on()` holds carries the programmatic stream
// serialization attribute, so inside a captured sequence the next kernel's CTAs are already resident, spinning in
// the wait, when the previous grid drains: the ~2 us launch gap between dependent kernels of a registration
// (60+ per pair) shrinks to the drain itself.  Both instructions are no-ops for ordinary launches.  The wait is
// executed before ANY exit path, so completion stays transitive along the chain.
#ifdef __CUDACC__
#define FCCF_PDL_ENTER() do { asm volatile("griddepcontrol.launch_dependents;"); asm volatile("griddepcontrol.wait;" ::: "memory"); } while (0)
bool& fccf_pdl_flag();
inline bool fccf_pdl_on() { return fccf_pdl_flag(); }
template <typename... KA, typename... A>
inline cudaError_t klaunch(void (*k)(KA...), dim3 g, dim3 b, size_t smem, cudaStream_t s, A... a) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = g; cfg.blockDim = b; cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = fccf_pdl_on() ? 1 : 0;
  const cudaError_t e = cudaLaunchKernelEx(&cfg, k, KA(a)...);
  if (e != cudaSuccess && getenv("FCCF_DEBUG_LAUNCH")) fprintf(stderr, "libfccf: launch failed (%s): grid %u %u %u, block %u, dynamic smem %zu\n", cudaGetErrorString(e), g.x, g.y, g.z, b.x, smem);
  return e;
}
#endif

// copies calls[lane] into every lane's state block and resets the per-registration counters
void launch_init_state(cudaStream_t s, const Batch& b, const CallArgs* d_calls, uint64_t* launches);
void cluster_init_attributes();
int cluster_nbl_ints();
int cluster_deg_ints();
void sort_init_attributes();
void planes_init_attributes();
void score_init_attributes();   // one-time function attributes (not allowed inside a stream capture)
// VoxelGrid stage `stage` (0: on raw clouds, 1: on the stage-0 output) for both clouds
void launch_voxelgrid(cudaStream_t s, const Batch& b, int stage, int ncloud, uint64_t* launches);
// The same stage by one thread-block cluster per cloud (voxelgrid_fast.cu).  scratch: per resident cluster
// `stride` u32 + `stride` float4 (sized to stay in L2).  Clouds it cannot hold raise ST_VG_FAST_MISS.
struct VgFastScratch { u32* info = nullptr; float4* queue = nullptr; int stride = 0; int ncl = 0; };
int vg_fast_init();             // one-time attributes; returns the number of co-resident clusters (0: unavailable)
int vg_fast_max_clusters();
int vg_fast_nmax();             // largest cloud (points) the cluster path takes
cudaError_t launch_voxelgrid_fast(cudaStream_t s, const Batch& b, int stage, int ncloud, const VgFastScratch& sc, uint64_t* launches);
// face_extrate for both clouds (input: vg_xyz[1] with st->vg[1][c].n_out points)
void launch_planes(cudaStream_t s, const Batch& b, int ncloud, int src_stage, uint64_t* launches);
void launch_hypotheses(cudaStream_t s, const Batch& b, uint64_t* launches);
void launch_cluster(cudaStream_t s, const Batch& b, uint64_t* launches);
void launch_quick_verify(cudaStream_t s, const Batch& b, uint64_t* launches);
void launch_fine_verify_build(cudaStream_t s, const Batch& b, uint64_t* launches);   // static voxel table (needs only the plane stage)
void launch_fine_verify_fuse(cudaStream_t s, const Batch& b, uint64_t* launches);    // scoring of the selected centres + fusion

// stand-alone stage entry points (C-ABI helpers)
void launch_quick_verify_list(cudaStream_t s, const fccf_params& p, float* d_T16, int n, const float* d_planes1, int f1,
                              const float* d_planes2, int f2, float* d_score, int* d_npair, int* d_pairs, int* d_iters, uint64_t* launches);
// lattice + static hash of the static leftover cloud (n1/n2 are device-side counts), G lanes at once
struct ScoreBuildJob { const float* s1; const int* n1; const int* n2; ScoreWS ws; };
void launch_score_build(cudaStream_t s, const fccf_params& p, const ScoreBuildJob* jobs, int G, ArgTable& tab, uint64_t* launches, bool lean = false);
// scores n_hyp hypotheses (row-major 4x4 each) -> d_scores
void launch_score_list(cudaStream_t s, const fccf_params& p, const float* d_T16, int n_hyp, const float* d_s2, const ScoreWS& ws, float* d_scores, ArgTable& tab, uint64_t* launches);
// packed (score, index) maximum of a device score list -> *d_out (8 bytes)
void launch_score_best(cudaStream_t s, const float* d_scores, int n, long long index_base, long long* d_out, uint64_t* launches);
void launch_fuse(cudaStream_t s, PipeState* st, const float* top_T, const float* top_s1, const float* top_s2, float fine_number, int topk, ArgTable& tab, uint64_t* launches);
// per-voxel (s,t) rows of one hypothesis: rows of 5 ints, *d_nrows rows
void launch_score_dump(cudaStream_t s, const fccf_params& p, const float* d_T16, const float* d_s2, const ScoreWS& ws, int* d_rows, int cap_rows, int* d_nrows, ArgTable& tab, uint64_t* launches);

}  // namespace fccf
