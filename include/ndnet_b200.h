/* include/ndnet_b200.h — C ABI of libndnet_b200.so (B200 / sm_100a).
 *
 * Two groups of entry points:
 *
 * (1) LEGACY, host-pointer symbols with the exact signatures of the reference's libndnet.so, so that
 *     /root/reference/ndnet/preprocessing/ndt_legacy.py:28-43,153-163,186-232 binds to this library
 *     unchanged (only the path in its LoadLibrary call changes).  Each replaces:
 *       ndt_downsample       core_legacy/include/ndnet_core/ndt.h:100-110   (src/ndt.c:119-222)
 *       prune_nds            core_legacy/include/ndnet_core/ndt.h:59-62     (src/ndt.c:28-73)
 *       to_point_cloud       core_legacy/include/ndnet_core/ndt.h:76-82     (src/ndt.c:75-117)
 *       free_nds             core_legacy/include/ndnet_core/ndt.h:116       (src/ndt.c:224-239)
 *       free_kl_divergences  core_legacy/include/ndnet_core/kullback_leibler.h:74 (src/kullback_leibler.c:204-207)
 *       print_matrix         core_legacy/include/ndnet_core/matrix.h:40     (src/matrix.c:28-35)
 *     The two struct pointers are opaque tokens here (the reference's Python never dereferences them,
 *     ndt_legacy.py:5-25 declares inert `__fields__`).  All work runs on the GPU; there is no CPU path.
 *     Error convention as the reference: negative int + a line on stderr; additionally -100 - cudaError
 *     for CUDA failures.
 *
 * (2) BATCHED, device-pointer entry points (ours): one call voxelises, searches, prunes and packs a whole
 *     batch of clouds on a CUDA stream, and runs the PointNet/NDT-Net forward.  Replaces the per-cloud
 *     Python loop of ndnet/preprocessing/ndtnet_preprocessing.py:27-63.
 *
 * Plain pointers and sizes only; no torch types.
 */
#ifndef NDNET_B200_H
#define NDNET_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------ (1) legacy ABI (host pointers) */
struct normal_distribution_t; /* opaque */
struct kl_divergence_t;       /* opaque */

int ndt_downsample(double *point_cloud, unsigned short point_dim, unsigned long num_points,
                   unsigned int *len_x, unsigned int *len_y, unsigned int *len_z,
                   double *offset_x, double *offset_y, double *offset_z,
                   double *voxel_size,
                   unsigned short *classes, unsigned short num_classes,
                   unsigned long num_desired_points,
                   double *downsampled_point_cloud, unsigned long *num_downsampled_points,
                   double *covariances,
                   unsigned short *downsampled_classes,
                   struct normal_distribution_t **nd_array, unsigned long *num_valid_nds,
                   struct kl_divergence_t **kl_divergences, unsigned long *num_kl_divergences);

int prune_nds(struct normal_distribution_t *nd_array,
              unsigned int len_x, unsigned int len_y, unsigned int len_z,
              unsigned long num_desired_nds, unsigned long *num_valid_nds,
              struct kl_divergence_t *kl_divergences, unsigned long *num_kl_divergences);

int to_point_cloud(struct normal_distribution_t *nd_array,
                   unsigned int len_x, unsigned int len_y, unsigned int len_z,
                   double offset_x, double offset_y, double offset_z,
                   double voxel_size,
                   double *point_cloud, unsigned long *num_points,
                   double *covariances,
                   unsigned short *classes);

void free_nds(struct normal_distribution_t *nd_array, unsigned long num_nds);
void free_kl_divergences(struct kl_divergence_t *kl_divergences);
void print_matrix(double *matrix, int rows, int cols);

/* ------------------------------------------------------------------ (2) batched ABI (device pointers) */

typedef struct ndnet_b200_ctx ndnet_b200_ctx;

#define NDNET_B200_F32 0
#define NDNET_B200_F64 1

/* flags for ndnet_b200_downsample_batch */
#define NDNET_B200_NAN_TO_NUM 1u /* NaN/+-inf -> 0 in the f32 features (ndtnet_preprocessing.py:66-69) */
#define NDNET_B200_LABELS_U8 2u  /* `labels` holds one BYTE per point instead of the reference's uint16 (num_classes <= 255):
                                    fewer bytes to move when the scans come from the host */
#define NDNET_B200_TEXTBOOK_KL 4u /* the algorithm the reference's README documents (README.md:6) instead of what its compiled
                                    core does: population covariances (off-diagonal sums divided once by the voxel's count,
                                    nothing factorised in place), the Kullback-Leibler divergence of the two Gaussians
                                    (mean term included), the LEAST divergent distributions removed first.  Default off:
                                    parity with the reference is defined on the legacy behaviour. */

/* Per-cloud record written by ndnet_b200_downsample_batch (device memory, one per cloud). */
typedef struct ndnet_b200_cloud_info {
    int32_t status;        /* 0 ok; -1 grid too large (stands in for malloc failure, ndt.c:151-155);
                              -3 voxel-size search did not converge (ndt.c:191-194);
                              -5 a coordinate is NaN: refused, outputs zero (undefined behaviour in the reference,
                                 voxel.c:89-91; +-inf coordinates end in -1: the grid cannot be held) */
    int32_t prune_status;  /* 0, or -2 when the walk hit the end of the list (ndt.c:53-56) */
    int32_t evaluations;   /* estimate passes the search made (<= 15) */
    uint32_t len[3];       /* accepted grid */
    uint32_t num_voxels;   /* occupied voxels before pruning (num_valid_nds before prune) */
    uint32_t num_valid;    /* after pruning */
    uint32_t num_kl;       /* divergence-list entries before pruning */
    uint32_t num_kl_after; /* list entries that remain after the walk */
    uint32_t num_out;      /* rows written (clamped to num_desired) */
    uint32_t num_survivors;/* rows the reference would have written (may exceed num_desired, A15) */
    double voxel_size;
    double offset[3];
    double limits[6];      /* max x,y,z, min x,y,z */
} ndnet_b200_cloud_info;

/* Create / destroy a context (owns the device workspace; one per host thread / stream). */
int ndnet_b200_create(ndnet_b200_ctx **ctx, int device);
void ndnet_b200_destroy(ndnet_b200_ctx *ctx);
const char *ndnet_b200_last_error(const ndnet_b200_ctx *ctx);
const char *ndnet_b200_version(void);

/* NDT voxelise + voxel-size search + neighbour pseudo-KL prune + 12-D feature pack for B clouds.
 *   points      device, [B, N, 3] contiguous, dtype f32 or f64 (NDNET_B200_F32/F64)
 *   labels      device, [B, N] uint16 point classes, or NULL
 *   num_classes classes are 0..num_classes inclusive (normal_distributions.c:115,158)
 *   num_desired n_desired_nds (D)
 *   out_feat    device, [B, D, 12] f32: mean(3) then row-major 3x3 "covariance"(9); rows >= num_out are 0
 *   out_feat64  device, [B, D, 12] f64 or NULL (bit-exact values, no nan_to_num)
 *   out_labels  device, [B, D] uint16 or NULL
 *   out_voxel   device, [B, D] int32 linear voxel index of each row (-1 for padding) or NULL
 *   info        device, [B] ndnet_b200_cloud_info or NULL
 *   stream      cudaStream_t (as void*); all work is enqueued, nothing synchronises
 * Returns 0 or -100-cudaError / -200-.. argument errors. */
int ndnet_b200_downsample_batch(ndnet_b200_ctx *ctx, const void *points, int dtype, const uint16_t *labels,
                                int B, long N, int num_classes, long num_desired, unsigned flags,
                                float *out_feat, double *out_feat64, uint16_t *out_labels, int32_t *out_voxel,
                                ndnet_b200_cloud_info *info, void *stream);

/* Same, from/to HOST buffers (pinned or pageable): H2D of the points (+labels), the kernels, D2H of the
 * features (+labels, info), then a stream synchronise.  This is the call the reference-facing Python
 * wrapper makes for CPU tensors and the one bench.py times as `e2e`. */
int ndnet_b200_downsample_batch_host(ndnet_b200_ctx *ctx, const void *points, int dtype, const uint16_t *labels,
                                     int B, long N, int num_classes, long num_desired, unsigned flags,
                                     float *out_feat, double *out_feat64, uint16_t *out_labels, int32_t *out_voxel,
                                     ndnet_b200_cloud_info *info, void *stream);

/* One-hot rows <-> class tags on the device, for callers that hold labels the way the reference's Python entry does
 * (ndnet/preprocessing/ndtnet_preprocessing.py:34 `argmax` of [rows, width] one-hot rows -> tag; :55-57 tag -> one-hot
 * row).  onehot: device f32 [rows, width]; labels: device uint16 [rows].  The first maximal entry wins (numpy / torch
 * argmax).  Enqueued on `stream`. */
int ndnet_b200_onehot_to_labels(const float *onehot, long rows, int width, uint16_t *labels, void *stream);
int ndnet_b200_labels_to_onehot(const uint16_t *labels, long rows, int width, float *onehot, void *stream);

/* Debug/inspection: copy per-point voxel ids of the last batch ([B,N] int32, -1 = never voxelised) to a
 * device buffer.  Used by the parity tests ("voxel ids and per-voxel membership bit-exact"). */
int ndnet_b200_keep_point_voxels(ndnet_b200_ctx *ctx, int enable);   /* off by default (costs 4 B/point of HBM writes) */
int ndnet_b200_last_point_voxels(ndnet_b200_ctx *ctx, int32_t *out_dev, void *stream);
/* Debug/inspection: the pre-prune divergence list of cloud b of the last batch, in list order.
 * Host buffers of capacity `cap`; returns the number of entries or a negative error.  The batched calls only sort the head
 * of the list that the prune walk can reach; ndnet_b200_keep_kl_list(ctx, 1) BEFORE the batch makes them sort and keep all
 * of it (the legacy ndt_downsample always does: its handles carry the list for prune_nds). */
int ndnet_b200_keep_kl_list(ndnet_b200_ctx *ctx, int enable);
long ndnet_b200_last_kl_list(ndnet_b200_ctx *ctx, int b, double *div, int32_t *p_voxel, int32_t *q_voxel, long cap);

/* Instrumentation for bench.py.  launch_count: kernels launched by this library since it was loaded.
 * stage_timing(1) makes every ndnet_b200_downsample_batch record CUDA events between its stages on the
 * launching stream (and synchronise at the end); stage_times returns the accumulated milliseconds for
 * {limits, search, rank, offsets, scatter, stats, kl, select} and the number of batches measured. */
long ndnet_b200_launch_count(void);
/* Self test: compares the statistics kernel's reciprocal+FMA division by an integer count with the IEEE
 * division on n pseudo-random operand pairs; returns the number of mismatching results (0 expected). */
long ndnet_b200_selftest_div(long n, unsigned seed);
/* Test hook: the next workspace allocation of this context fails as if the device were out of memory (the call that
 * triggers it returns the allocation error; the one after must allocate afresh and succeed). */
int ndnet_b200_test_fail_next_reserve(ndnet_b200_ctx *ctx);
int ndnet_b200_stage_timing(ndnet_b200_ctx *ctx, int enable);
/* Of the last ndnet_b200_downsample_batch on this context, averaged over its clouds: how many voxel-size guesses needed a
 * pass over the points, and how many guesses were evaluated in all (a grid with fewer cells than num_desired is decided
 * without reading the points).  Synchronises the device.  bench.py divides the search stage by the former. */
int ndnet_b200_last_search_passes(ndnet_b200_ctx *ctx, double *mean_passes, double *mean_evaluations);
int ndnet_b200_stage_times(ndnet_b200_ctx *ctx, double *ms, int cap, long *runs);

/* ---- PointNet / NDT-Net forward (ndnet/models/ndtnet.py:33-62,112-164,181-196,218-243) ------------ */

typedef struct ndnet_b200_model ndnet_b200_model;

/* kinds */
#define NDNET_B200_NDTNET_CLS 0     /* ndnet/models/ndtnet.py:166-196  */
#define NDNET_B200_NDTNET_SEG 1     /* ndnet/models/ndtnet.py:198-243  */
#define NDNET_B200_POINTNET_CLS 2   /* ndnet/models/pointnet.py:137-167 */
#define NDNET_B200_POINTNET_SEG 3   /* ndnet/models/pointnet.py:169-214 */

/* Build a model from a flat list of named fp32 host tensors (the reference state_dict: names follow
 * ndtnet.py module attributes, e.g. "feature_extractor.t1.conv1.weight").  BatchNorm is folded with its
 * running statistics (eval mode).  feature_dim / num_classes are read from the tensor shapes. */
int ndnet_b200_model_create(ndnet_b200_ctx *ctx, ndnet_b200_model **model, int kind, int n_tensors,
                            const char *const *names, const float *const *data, const int64_t *const *shapes,
                            const int *ndims);
void ndnet_b200_model_destroy(ndnet_b200_model *model);
/* Segmentation head layers 1 + 2 (ndtnet.py:231-232) run as ONE kernel by default: the 512-wide activation stays in tensor
 * memory / shared memory.  enable = 0 runs them as two GEMMs (the activation goes through HBM and can be tapped as "head.l1"). */
int ndnet_b200_model_set_fused_head(ndnet_b200_model *model, int enable);
/* floats per input row the model expects: 12 for NDT-Net (mean + covariance), point_dim for PointNet */
int ndnet_b200_model_input_dim(const ndnet_b200_model *model);

/* Forward over B clouds of D distributions.  feat: device [B, D, input_dim] f32.
 * cls: out device [B, num_classes] f32 (softmax);  seg: out device [B, D, num_classes+1] f32 (log_softmax). */
int ndnet_b200_model_forward(ndnet_b200_ctx *ctx, ndnet_b200_model *model, const float *feat, int B, int D,
                             float *out, void *stream);

/* Inspection (parity tests): one internal activation of the LAST ndnet_b200_model_forward of `model` on `ctx`, converted
 * to fp32 into the device buffer `out` of capacity `cap` floats (out == NULL: size query).  Names: "t1" [B,d,d] and "t2"
 * [B,64,64] (the T-Net transforms, ndtnet.py:33-62), "t1.pool" / "t2.pool" [B,1024], "trunk.l1" [B,N,64] (ndtnet.py:149),
 * "trunk.xt2" [B,N,64] (:153-155), "trunk.pool" [B,F] (max over points of :161), "head.l1/l2/l3" [B,N,512/256/128]
 * (:231-233; segmentation) or "trunk.l2" [B,N,128] (classification).  Returns the element count or a negative error. */
long ndnet_b200_model_tap(ndnet_b200_ctx *ctx, ndnet_b200_model *model, const char *name, float *out, long cap, void *stream);

/* The whole hot path in one call from HOST buffers (what bench.py times as `e2e`): H2D of points (+labels),
 * NDT (NaN/inf -> 0 as ndtnet_preprocessing.py:66-69), network forward, D2H of the result
 * (seg: [B, D, num_classes+1] log-probabilities; cls: [B, num_classes] probabilities), stream synchronise.
 * out_elems_per_cloud = D*(num_classes+1) or num_classes of the MODEL. */
int ndnet_b200_infer_host(ndnet_b200_ctx *ctx, ndnet_b200_model *model, const void *points, int dtype,
                          const uint16_t *labels, int B, long N, int num_classes, long num_desired, float *out_host,
                          long out_elems_per_cloud, void *stream);
/* Same with one byte per point label (13 instead of 14 bytes per point cross PCIe; num_classes <= 255). */
int ndnet_b200_infer_host_u8(ndnet_b200_ctx *ctx, ndnet_b200_model *model, const void *points, int dtype,
                             const uint8_t *labels, int B, long N, int num_classes, long num_desired, float *out_host,
                             long out_elems_per_cloud, void *stream);
/* Asynchronous form for a caller that keeps several batches in flight (a serving loop): returns once the copies and kernels
 * are enqueued; `out_host` is valid after ndnet_b200_infer_wait (or a synchronise of `stream`).  Consecutive calls overlap -
 * the copies of one batch travel while the kernels of the previous one drain - so every call needs its own `out_host`, and the
 * input buffers must stay untouched until the wait.  labels_u8 != 0: `labels` holds one byte per point. */
int ndnet_b200_infer_host_async(ndnet_b200_ctx *ctx, ndnet_b200_model *model, const void *points, int dtype, const void *labels,
                                int labels_u8, int B, long N, int num_classes, long num_desired, float *out_host,
                                long out_elems_per_cloud, void *stream);
int ndnet_b200_infer_wait(ndnet_b200_ctx *ctx, void *stream);
/* Same with DEVICE buffers in and out, asynchronous (the caller's stream waits for the result). */
int ndnet_b200_infer_device(ndnet_b200_ctx *ctx, ndnet_b200_model *model, const void *points, int dtype,
                            const uint16_t *labels, int B, long N, int num_classes, long num_desired, float *out_dev,
                            long out_elems_per_cloud, void *stream);
/* Both infer calls cut the batch into chunks of at most `chunk` scans and run the chunks round-robin on `lanes`
 * internal streams (each with its own workspace), so that host<->device copies overlap kernels and the
 * latency-bound tail of one chunk (the sequential per-voxel statistics) overlaps the bulk work of another.
 * Defaults: 2 lanes, 64 scans.  The lane count is fixed at the first infer call. */
int ndnet_b200_set_pipeline(ndnet_b200_ctx *ctx, int lanes, int chunk);
/* Chunk size of ndnet_b200_infer_device only (default 128): with the scans already in HBM there are no copies to hide,
 * and fewer, larger chunks run faster (2048 scans per call: 64-scan chunks 92 k clouds/s, 128: 108 k, 256: 127 k, 512 and
 * above: 132-135 k; bench.py uses 512). */
int ndnet_b200_set_device_chunk(ndnet_b200_ctx *ctx, int chunk);
/* Staggered lanes (on != 0): a chunk's front - bounding box, voxel-size search, voxel assignment, the kernels that stream
 * every point from HBM - starts behind the front of the chunk before it, so that the front of chunk k runs beside the
 * statistics / divergences / selection / network of chunk k-1 instead of all lanes walking through the same stage side by
 * side.  The environment variable NDNET_B200_STAGGER=0|1 overrides the setting. */
int ndnet_b200_set_stagger(ndnet_b200_ctx *ctx, int on);
/* CUDA-graph replay of the NDT chain inside ndnet_b200_downsample_batch.  mode -1 (default): batches of at most ~4 M points
 * (32 scans of 120 k), where the ~55 dependent launches of the chain are bound by launch latency; 0: never; 1: every batch.
 * A shape gets its graph the second time it is seen; the graph runs on buffers the context owns (the scans are copied in and
 * the results out, device to device, around one cudaGraphLaunch), results are bit-identical to the direct launches.  A call
 * whose stream is itself being captured launches directly into the caller's graph.  NDNET_B200_NDT_GRAPH=-1|0|1 overrides. */
int ndnet_b200_set_ndt_graph(ndnet_b200_ctx *ctx, int mode);

/* ------------------------------------------------------------------ (3) ASCII-PLY ingest
 * Replaces the per-line Python loop of /root/reference/ndnet/datasets/CARLA_Seg.py:96-183 (`get_data_pcl`):
 * skip `num_header_lines` lines (:115), then per line float(data[0..2]) and int(data[-1]) (:117-123), refuse a tag
 * above n_classes (:127-128); points are held as float32 (:173), tags as uint16 (:146).  `text` is the file's bytes
 * (host, or device when text_on_device != 0).  The parsed cloud stays resident on the GPU behind the handle; its
 * buffers come from the device's stream-ordered pool on `stream`, which must outlive the handle.
 * Returns 0, or (first offending line in *bad_line, 0-based in the file; the tag in *bad_value for -303):
 *   -301  a data line has fewer than three tokens                (IndexError in the reference)
 *   -302  a token is not a float()/int() literal                  (ValueError)
 *   -303  class tag > n_classes                                   (ValueError "Class tag {tag} out of bounds", :128)
 *   -304  literal outside what is converted exactly: inf/nan/underscores, more than 19 significant digits,
 *         |decimal exponent| > 55 — refused, never approximated
 *   -305  negative class tag                                      (OverflowError at :146)
 *   -306  lone '\r' line ends or non-ASCII bytes
 * Decimal -> double is correctly rounded (CPython float()), then rounded to float32 as torch's .float(). */
typedef struct ndnet_b200_ply ndnet_b200_ply;
int ndnet_b200_ply_load(int device, const char *text, size_t nbytes, int text_on_device, int num_header_lines,
                        int n_classes, void *stream, ndnet_b200_ply **out, unsigned long *num_points, long *bad_line,
                        long *bad_value);
long ndnet_b200_ply_num_points(const ndnet_b200_ply *ply);
/* Gathers rows `indexes[0..n)` (host int64, or device when indexes_on_device != 0; NULL = every point in file
 * order) into DEVICE outputs, any of which may be NULL: points [n,3] f32, labels [n] u16, one-hot [n, n_classes+1]
 * f32 (:141-147,173-179).  -307 when an index is out of range. */
int ndnet_b200_ply_sample(ndnet_b200_ply *ply, const int64_t *indexes, size_t n, int indexes_on_device,
                          float *out_points, uint16_t *out_labels, float *out_onehot, void *stream);
void ndnet_b200_ply_free(ndnet_b200_ply *ply);

/* ------------------------------------------------------------------ (4) training step of the network
 * Train-mode forward and backward of the reference's four networks, recognised from the state_dict keys and shapes:
 * NDTNetSegmentation / NDTNetClassification (/root/reference/ndnet/models/ndtnet.py:33-62,112-164,181-196,218-243) and
 * PointNetSegmentation / PointNetClassification (pointnet.py:65-214, any point_dim <= 64), as run by the reference's
 * training loops (/root/reference/tools/train.py:66-76): BatchNorm with batch statistics over all rows (running
 * statistics updated in place, momentum 0.1, eps 1e-5), fp32 throughout.
 * `names`/`shapes` describe the module's parameters AND buffers (state_dict keys); `tensors[i]` / `grads[i]` of the
 * forward/backward calls are DEVICE pointers in that same order (grads[i] may be NULL for buffers; int64
 * num_batches_tracked buffers are passed through the same array and incremented).  The loss stays with the caller:
 * forward writes the network's output - segmentation: log-probabilities [B,N,C+1]; classification: probabilities
 * [B,num_classes] - and backward takes dL/d(output) of the same shape and OVERWRITES each grads[i].  `feat` is
 * [B,N,12] = [mean | covariance] for the NDT networks and [B,N,point_dim] for PointNet (ndnet_b200_trainer_info).
 * B >= 2 (BatchNorm over the FC layers of the T-Nets needs more than one cloud). */
typedef struct ndnet_b200_trainer ndnet_b200_trainer;
int ndnet_b200_trainer_create(int device, int n_tensors, const char *const *names, const int64_t *const *shapes,
                              const int *ndims, ndnet_b200_trainer **out);
/* kind: 0 NDTNetClassification, 1 NDTNetSegmentation, 2 PointNetClassification, 3 PointNetSegmentation; width of a row of
 * `feat`; outputs per row (segmentation) or per cloud (classification).  Any pointer may be NULL. */
int ndnet_b200_trainer_info(const ndnet_b200_trainer *t, int *kind, int *point_width, int *outputs);
int ndnet_b200_trainer_forward(ndnet_b200_trainer *t, const float *feat /* [B,N,point_width] */, int B, int N,
                               float *const *tensors, float *out_logp, int update_running_stats, void *stream);
int ndnet_b200_trainer_backward(ndnet_b200_trainer *t, const float *dlogp, float *const *tensors, float *const *grads,
                                void *stream);
/* Same, with the gradients of all parameters written into ONE flat device buffer: parameter i (a weight, bias or
 * BatchNorm scale/shift) starts at offsets[i] floats (16-byte aligned; -1 for buffers), total length = the return value
 * of ndnet_b200_trainer_grad_layout. */
int ndnet_b200_trainer_backward_flat(ndnet_b200_trainer *t, const float *dlogp, float *const *tensors, float *flat_grads,
                                     void *stream);
long ndnet_b200_trainer_grad_layout(const ndnet_b200_trainer *t, long *offsets, int n_tensors);
/* Gradient buckets for an all-reduce that overlaps the backward pass (/root/reference/tools/train.py:72-81 has one process;
 * data-parallel training adds the collective).  The flat layout follows the order in which backward finishes the parameters:
 * bucket 0 = segmentation head, 1 = trunk conv2/conv3 + feature T-Net, 2 = conv1 + input T-Net; bucket i covers floats
 * [begin, end) of the flat buffer.  After ndnet_b200_trainer_backward_flat returns (everything is only enqueued),
 * ndnet_b200_trainer_bucket_ready(t, i, flat, side) makes the stream `side` wait for the event recorded when bucket i became
 * final - so a collective launched on `side` runs while the rest of the backward still executes.  With CUDA graphs the
 * gradients are produced in a library-owned buffer; set_deferred_copy(1) drops the whole-buffer copy at the end of
 * backward_flat and bucket_ready copies each bucket into `flat` on `side` instead. */
int ndnet_b200_trainer_num_buckets(const ndnet_b200_trainer *t);
int ndnet_b200_trainer_bucket_range(const ndnet_b200_trainer *t, int i, long *begin, long *end);
int ndnet_b200_trainer_set_deferred_copy(ndnet_b200_trainer *t, int enable);
int ndnet_b200_trainer_bucket_ready(ndnet_b200_trainer *t, int i, float *flat_grads, void *side_stream);
/* enable = 1: the ~100 / ~250 kernel launches of a forward / backward pass are captured into CUDA graphs (one per shape
 * and pointer set; the first pass of a configuration runs eagerly) and replayed; inputs and outputs are staged through
 * library-owned buffers so the replayed pointers never change.  Default 0. */
int ndnet_b200_trainer_set_graph(ndnet_b200_trainer *t, int enable);
/* tf32 = 1: the large GEMMs (forward, dgrad, wgrad) run on the tensor cores (tcgen05 kind::tf32, operands read from the
 * fp32 buffers, fp32 accumulation); tf32 = 0 (default): fp32 FMA everywhere (the parity configuration). */
int ndnet_b200_trainer_set_precision(ndnet_b200_trainer *t, int tf32);
/* Test hook: C[M,N] (+)= A[M,K] . B[N,K]^T (+ bias) through the training GEMM kernels, device pointers, row strides in
 * floats.  mode 0 = fp32 FMA kernel, 1 = tcgen05 TF32 kernel (-206 when the shape/alignment is not eligible for it). */
int ndnet_b200_debug_train_gemm(int mode, const float *A, long lda, const float *B, long ldb, float *C, long ldc, int M, int N,
                                int K, const float *bias, int accumulate, void *stream);
const char *ndnet_b200_trainer_last_error(const ndnet_b200_trainer *t);
/* Test hook: copies one internal activation/gradient buffer of the last pass ("h3.dA", "t2.c1.Y", "c1.A", "t1.T", ...)
 * to `out` (device, may be NULL to query the size); returns the element count or -200. */
long ndnet_b200_trainer_debug_buffer(ndnet_b200_trainer *t, const char *name, float *out, void *stream);
void ndnet_b200_trainer_destroy(ndnet_b200_trainer *t);

#ifdef __cplusplus
}
#endif
#endif /* NDNET_B200_H */
