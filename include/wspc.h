/*
 * wspc.h — C ABI of the B200-native (sm_100a) hot path for weakly-supervised
 * point-cloud segmentation (drop-in for the DGCNN EdgeConv stack + weak-sup
 * losses of alex-xun-xu/WeakSupPointCloudSeg).
 *
 * The reference has no FFI of its own: its "native layer" is the set of stock
 * TensorFlow-1.14 ops behind the Python functions cited next to each entry
 * point below (paths relative to the reference checkout).  Every entry point
 * replaces one such call site.
 *
 * Conventions (all entry points):
 *   - plain device pointers + sizes; no torch / C++ types in any signature;
 *   - returns WSPC_OK (0) or a negative WSPC_ERR_* code, never throws/exits;
 *     wspc_last_error() gives a thread-local human-readable message;
 *   - no allocation, no host synchronisation, no default-stream use inside:
 *     the caller owns every buffer, passes scratch as (workspace, bytes) sized
 *     by the matching *_workspace_bytes(), and passes the stream to launch on;
 *   - all tensors are dense row-major fp32 / int32 unless a leading dimension
 *     (ld*) says otherwise; base pointers must be 16-byte aligned;
 *   - built for sm_100a only; on any other device every call returns
 *     WSPC_ERR_ARCH (there is no fallback path).
 */
#ifndef WSPC_H_
#define WSPC_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WSPC_OK 0
#define WSPC_ERR_INVALID (-1)   /* bad argument (shape, alignment, null)   */
#define WSPC_ERR_WORKSPACE (-2) /* workspace too small                     */
#define WSPC_ERR_CUDA (-3)      /* a CUDA runtime call / launch failed     */
#define WSPC_ERR_ARCH (-4)      /* device is not sm_100                    */

typedef struct CUstream_st* wspc_stream_t; /* == cudaStream_t */

/* distance arithmetic flavours */
#define WSPC_DIST_TFUTIL 0 /* (sq_i + (-2*dot_ij)) + sq_j            tf_util.py:652-657         */
#define WSPC_DIST_SMOOTH 1 /* max((sq_i + sq_j) - 2*dot_ij, 0)       SmoothConstraint.py:144-148 */

int wspc_version(void);
const char* wspc_last_error(void);
/* number of kernel launches issued through this library by the calling
 * process since load (bench.py reports the per-step delta as gpu_launches). */
uint64_t wspc_launch_count(void);

/* ------------------------------------------------------------------ kNN --- */
/* Fused pairwise distance + k-nearest selection; the N x N matrix is never
 * materialised.  Replaces tf_util.pairwise_distance + tf_util.knn
 * (Networks/dgcnn/utils/tf_util.py:638-671) and, with WSPC_DIST_SMOOTH, the
 * Dmat/top_k block of Util/SmoothConstraint.py:141-154.
 *   x    : (B, N, ldx) fp32; the D feature channels start at column `coff`
 *   idx  : (B, N, k) int32, nearest first, ties -> lower index (tf.nn.top_k)
 *   dist : (B, N, k) fp32 distances of the selected neighbours, or NULL
 * Arithmetic is the canonical fp32 chain of SURVEY.md App. A-1 (sequential
 * fmaf over channels) so indices are bit-exact against oracle/knn_oracle.c.
 * Limits: 1 <= k <= 64, k <= N, 1 <= D <= 128. */
size_t wspc_knn_workspace_bytes(int B, int N, int D);
int wspc_knn_fused(const float* x, int B, int N, int ldx, int coff, int D, int k, int flavour,
                   int32_t* idx, float* dist, void* workspace, size_t workspace_bytes,
                   wspc_stream_t stream);

/* Unfused API-compat pair (same arithmetic): adj (B, N, N) fp32.
 * tf_util.pairwise_distance (tf_util.py:638-657). */
int wspc_pairwise_distance(const float* x, int B, int N, int ldx, int coff, int D, int flavour,
                           float* adj, void* workspace, size_t workspace_bytes,
                           wspc_stream_t stream);
/* tf_util.knn (tf_util.py:660-671) == tf.nn.top_k(-adj, k).indices: the k
 * smallest entries of each of `rows` rows of length `ncols`, ascending, ties
 * -> lower index.  vals may be NULL. */
int wspc_topk_rows(const float* adj, long long rows, int ncols, int k, int32_t* idx, float* vals,
                   wspc_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* WSPC_H_ */
