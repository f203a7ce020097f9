/* fccf.h — C-ABI of libfccf, the B200-native FCCF-PCR registration path.
 *
 * Drop-in boundary for the reference's operator
 *     void computer_transform_guess(PointCloud<PointXYZ>::Ptr source, Ptr target, Matrix4f& best)
 * (/root/reference/FCCF.cpp:1370) and for the program around it, main() (FCCF.cpp:1646-1690).
 * Plain pointers and sizes only; no C++/torch types cross this boundary.  All compute runs in
 * hand-written sm_100a CUDA kernels; there is no CPU fallback: every entry point fails with a
 * non-zero status when no CUDA device is usable.
 *
 * Ownership: the caller owns every host pointer; a context owns its device memory, stream and
 * CUDA graph.  Threading: one context per host thread per GPU; a context is not thread-safe.
 * Errors: int status (0 = ok), text via fccf_last_error().
 */
#ifndef FCCF_H_
#define FCCF_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct fccf_ctx fccf_ctx;

/* Every file-scope tunable of the reference (FCCF.cpp:126-176) with its default. */
typedef struct fccf_params {
  float parameter_l1, parameter_l2, parameter_k1, parameter_k2;      /* FCCF.cpp:126-129 */
  float normal_vector_threshold1, normal_vector_threshold2;          /* :131-132 */
  float face_voxel_size;                                              /* :134 */
  float voxel_point_threshold;                                        /* :136 */
  float curvature_threshold;                                          /* :138 */
  float select_plane_number;                                          /* :141 */
  float quick_verify_angel_threshold, quick_verify_distance_threshold; /* :145-146 */
  float required_optimize_plane;                                      /* :147 */
  float fine_verify_voxel_size;                                       /* :150 */
  float fine_verify_number;                                           /* :151 */
  float included_angle_same_threshold;                                /* :156 */
  float included_angle_min_threshold, included_angle_max_threshold;   /* :157-158 */
  float third_plane_threshold;                                        /* :160 */
  float third_plane_normal_threshold;                                 /* :162 */
  float cluster_number_threshold;                                     /* :166 */
  float cluster_angel_threshold, cluster_distance_threshold;          /* :167-168 */
  float seclct_cluster_number;                                        /* :171 */
  float rough_threshold_gl;                                           /* :175 */
  int emulate_pcl_overflow; /* 1: reproduce pcl::VoxelGrid's int32 bail-out (output = input) */
  int batch_lanes;          /* registrations per batched launch sequence of fccf_register_batch* (0: default 64); fixed at fccf_create */
  int reserved[2];
} fccf_params;

typedef struct fccf_timing {
  float h2d_ms;       /* host->device copies of both raw clouds */
  float downsample_ms;/* main()'s two VoxelGrid filters, FCCF.cpp:1668-1678 (outside the reference's clock) */
  float pipeline_ms;  /* computer_transform_guess, the reference's clock() region FCCF.cpp:1682-1684 */
  float d2h_ms;       /* result read-back */
  float total_ms;     /* end to end, host pointers in -> matrix out */
  int n_launches;     /* kernels launched for this registration */
  unsigned long long h2d_bytes, d2h_bytes; /* bytes copied host->device / device->host */
  float stage_ms[8];  /* voxelgrid(main), voxelgrid(pipeline), planes, hypotheses, cluster, quick_verify, fine_verify+fuse;
                         [1]..[6] are filled only while stage timing is on (fccf_set_stage_timing: six more event nodes
                         in every captured sequence, ~27 us per registration); [7]: batch entry points only, host wall
                         clock of the whole batch in ms */
} fccf_timing;

enum { FCCF_OK = 0, FCCF_ERR_CUDA = 1, FCCF_ERR_ARG = 2, FCCF_ERR_CAPACITY = 3, FCCF_ERR_NO_DEVICE = 4 };
enum { FCCF_F32 = 0, FCCF_F64 = 1, FCCF_I32 = 2, FCCF_I64 = 3 };

/* replaces: the global initialisers FCCF.cpp:126-176 */
void fccf_default_params(fccf_params* p);

/* Creates a context on CUDA device `device`.  NULL if there is no usable device (no fallback). */
fccf_ctx* fccf_create(int device, const fccf_params* params);
void fccf_destroy(fccf_ctx* ctx);
const char* fccf_last_error(const fccf_ctx* ctx);
int fccf_set_params(fccf_ctx* ctx, const fccf_params* params);
/* per-stage device times (fccf_timing.stage_ms[1..6]): off by default (the reference prints one clock() figure,
 * FCCF.cpp:1682-1686); FCCF_STAGE_EVENTS=1 in the environment turns it on at fccf_create */
int fccf_set_stage_timing(fccf_ctx* ctx, int on);

/* replaces: main() from the loaded clouds on (FCCF.cpp:1667-1687): VoxelGrid(leaf) on each cloud,
 * then computer_transform_guess(cloud_tar, cloud_src, T) — note the reference's swapped argument
 * order (FCCF.cpp:1683).  src = argv[1], tar = argv[2]; xyz are packed float32 triples in host
 * memory.  T_out: row-major 4x4 mapping src into tar's frame.  timing may be NULL. */
int fccf_register(fccf_ctx* ctx, const float* src_xyz, size_t n_src, const float* tar_xyz, size_t n_tar,
                  float leaf, float T_out[16], fccf_timing* timing);

/* Same, with both raw clouds already resident in device memory (device pointers). */
int fccf_register_device(fccf_ctx* ctx, const float* d_src_xyz, size_t n_src, const float* d_tar_xyz, size_t n_tar,
                         float leaf, float T_out[16], fccf_timing* timing);

/* Batch of independent pairs (BASELINE config 4): pair b uses src[b]/tar[b]; T_out is B x 16.  The batch
 * is cut into chunks of up to params.batch_lanes pairs; a chunk runs as ONE launch sequence (every kernel
 * launched once for all its pairs, one CUDA graph), and up to four chunks rotate on their own streams so
 * that the host->device copies of one overlap the compute of the others.  Results are bit-identical to
 * fccf_register on each pair.  timing: device times summed over the chunks (each figure covers its whole
 * chunk; divide by the number of pairs for the amortised per-registration figure), total_ms = device time
 * of the whole batch (CUDA events spanning every chunk), stage_ms[7] = host wall clock of the call. */
int fccf_register_batch(fccf_ctx* ctx, int n_pairs, const float* const* src_xyz, const size_t* n_src,
                        const float* const* tar_xyz, const size_t* n_tar, float leaf, float* T_out, fccf_timing* timing);
/* Same with every cloud already resident in device memory (arrays of device pointers, held on the host). */
int fccf_register_batch_device(fccf_ctx* ctx, int n_pairs, const float* const* d_src_xyz, const size_t* n_src,
                               const float* const* d_tar_xyz, const size_t* n_tar, float leaf, float* T_out, fccf_timing* timing);

/* Batch of independent pairs over SEVERAL contexts, one per GPU (BASELINE config 4 inside the library): the pairs
 * are cut into n_ctx contiguous blocks, block i is registered by ctxs[i] on its own host thread; no data-path
 * collective (the unit of work is one independent main() run, FCCF.cpp:1646-1690).  Host pointers may be pinned
 * or pageable (std::vector / pcl storage): pageable clouds go through the context's pinned staging chunks.
 * timing: NULL or an array of n_ctx records (one per context, as fccf_register_batch fills it). */
int fccf_register_batch_multi(fccf_ctx* const* ctxs, int n_ctx, int n_pairs, const float* const* src_xyz, const size_t* n_src,
                              const float* const* tar_xyz, const size_t* n_tar, float leaf, float* T_out, fccf_timing* timing);

/* Number of CUDA devices this process can use (0: none). */
int fccf_device_count(void);

/* replaces: pcl::VoxelGrid<PointXYZ>::filter as called at FCCF.cpp:1668-1678 / 1377-1387.
 * out_xyz: capacity n points; out_cell (int64 linear cell index) / out_cnt may be NULL. */
int fccf_voxelgrid(fccf_ctx* ctx, const float* xyz, size_t n, float leaf, float* out_xyz, int64_t* out_cell,
                   int32_t* out_cnt, size_t* n_out);

/* replaces: face_extrate (FCCF.cpp:470-678) on one already-downsampled cloud; results are read
 * through fccf_debug_blob with the tag "1" (vox_key1, vox_cnt1, vox_plane1, face_plane1, ...). */
int fccf_extract_planes(fccf_ctx* ctx, const float* xyz, size_t n, int32_t* n_faces);

/* replaces: fine_verify (FCCF.cpp:785-839) for H hypotheses at once ("hypotheses scored"): static
 * leftover cloud s1, moving leftover cloud s2, T = H row-major 4x4.  scores: H floats.
 * counts (optional): per hypothesis the per-voxel overlap table is kept on the device and can be
 * read with fccf_score_counts. */
int fccf_score_hypotheses(fccf_ctx* ctx, const float* T, size_t n_hyp, const float* s1_xyz, size_t n1,
                          const float* s2_xyz, size_t n2, float* scores);
/* Device-resident variant used for throughput measurement: uploads once, then scores `repeat`
 * times; returns the average kernel milliseconds per launch in *kernel_ms. */
int fccf_score_hypotheses_bench(fccf_ctx* ctx, const float* T, size_t n_hyp, const float* s1_xyz, size_t n1,
                                const float* s2_xyz, size_t n2, int repeat, float* scores, float* kernel_ms);
/* replaces: the best-hypothesis scan (strict `>` first maximum, FCCF.cpp:1559) over the scores of the
 * last fccf_score_hypotheses[_bench] call, as a block-reduced argmax on the device.  The result is one
 * packed int64: (order-preserving int32 of the float score) << 32 | (0xFFFFFFFF - (index_base + i));
 * its signed maximum over several GPUs (one 8-byte all-reduce) is the global best, smallest index
 * among ties.  packed_host and/or packed_device (a device pointer, e.g. an NCCL buffer) receive it. */
int fccf_score_best(fccf_ctx* ctx, size_t index_base, int64_t* packed_host, int64_t* packed_device);

/* Hypothesis scoring sharded over several contexts / GPUs (BASELINE config 3; fine_verify FCCF.cpp:785-839 and the
 * first-maximum scan 1559): the ordered list T is cut into n_ctx contiguous ranges, every context builds the same
 * static voxel table (s1) and holds the same moving cloud (s2), scores its range, reduces it to one packed
 * (score, global index) word on the device, and ONE 8-byte ncclAllReduce(max, int64) over NVLink gives every GPU the
 * global best (smallest index among ties).  libnccl.so.2 is loaded at run time; when it is missing (or two contexts
 * share a device) the n_ctx words are compared on the host and *used_nccl is 0.  scores: NULL or n_hyp floats. */
int fccf_score_sharded(fccf_ctx* const* ctxs, int n_ctx, const float* T, size_t n_hyp, const float* s1_xyz, size_t n1,
                       const float* s2_xyz, size_t n2, float* scores, float* best_score, int64_t* best_index, int* used_nccl);

/* Per-voxel overlap counts of hypothesis `hyp` of the last fccf_score_hypotheses call: rows of
 * (Lx, Ly, Lz, s, t) for voxels holding both static and moving points; returns rows in *n_rows. */
int fccf_score_counts(fccf_ctx* ctx, size_t hyp, int32_t* rows, size_t cap_rows, size_t* n_rows);

/* Exhaustive scoring: set params.fine_verify_number >= 256 — every cluster centre is refined
 * and fine-verified instead of the reference's 4 per type; no separate entry point is needed. */

/* replaces: quick_verify + ceres_refine (FCCF.cpp:680-783, 210-249) for H hypotheses given two plane
 * tables (F x 7 floats: centroid, normal, point size).  T is updated in place (refined);
 * scores: H floats; pair_count/pairs (optional): per hypothesis up to 16 (i1,i2) pairs; iters optional. */
int fccf_quick_verify(fccf_ctx* ctx, float* T, size_t n_hyp, const float* planes1, int f1, const float* planes2, int f2,
                      float* scores, int32_t* pair_count, int32_t* pairs, int32_t* iters);

/* replaces: select_base (FCCF.cpp:429-468) on one plane table (F x 7 floats: centroid, normal, point size) with its
 * roughness values theta (F doubles, FCCF.cpp:660-667): pairs = rows of (i, j, type), angles = included angles. */
int fccf_base_pairs(fccf_ctx* ctx, const float* planes, const double* theta, int f, int32_t* pairs, float* angles, int32_t* n_pairs);

/* replaces: select_base x 2, the pair-descriptor match loop (FCCF.cpp:1412-1427), computer_transform (841-1018) and
 * the matrix -> quaternion conversion (1439-1462) on two plane tables.  n_hyp: hypotheses per roughness type; the
 * pools themselves are read with fccf_debug_blob: base1/2, base_angle1/2, matches, hyp0..2 (3x4), hyp_qt0..2. */
int fccf_hypotheses(fccf_ctx* ctx, const float* planes1, const double* theta1, int f1, const float* planes2, const double* theta2, int f2,
                    int32_t n_hyp[3]);

/* replaces: cluster_num (FCCF.cpp:1465) + transform_cluster (1040-1231) for the three pools given as rows of
 * (qw qx qy qz tx ty tz), pools concatenated in type order.  Centres etc. through fccf_debug_blob: centre0..2,
 * n_centres, cluster_num, cluster_seed_sorted<t>, cluster_size_sorted<t>. */
int fccf_cluster(fccf_ctx* ctx, const float* qt7, const int32_t n_hyp[3], int32_t n_centres[3]);

/* replaces: the per-type best by s1/sum(s1) + s2/sum(s2), the 0.8 gate and fuse_answer (FCCF.cpp:1546-1606,
 * 1291-1368) on up to 3 x k fine-verified candidates: top_T [3][k] row-major 4x4, s1 / s2 [3][k] (quick / fine
 * score), n_top[t] candidates of type t in rank order.  Blob type_best: per type (score, 3x4). */
int fccf_fuse(fccf_ctx* ctx, const float* top_T, const float* s1, const float* s2, const int32_t n_top[3], int k, float T_out[16]);

/* Stage intermediates of the last fccf_register / fccf_extract_planes call, copied to host memory.
 * Returns FCCF_ERR_ARG for an unknown name; *bytes = size needed (also when dst is NULL). */
int fccf_debug_blob(fccf_ctx* ctx, const char* name, void* dst, size_t cap_bytes, size_t* bytes, int* dtype);

/* The CUDA stream (cudaStream_t) every kernel of this context is launched on, so that a caller can
 * bracket calls with its own events.  NULL without a context. */
void* fccf_stream_handle(const fccf_ctx* ctx);

/* Number of kernels this library has launched in the context so far. */
uint64_t fccf_launch_count(const fccf_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* FCCF_H_ */
