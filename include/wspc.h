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

/* Kernel selection for wspc_knn_fused: 0 = auto (D <= 64, k <= 64: tcgen05 distances with a proven error
 * margin + exact fp32 re-scoring and a per-row exact fallback; otherwise the CUDA-core kernel), 1 = CUDA-core
 * kernel only.  Both paths return bit-identical results; returns the previous setting. */
int wspc_set_knn_path(int path);
/* Telemetry (synchronises the device): rows of the last tensor-core wspc_knn_fused call on `workspace` that
 * were recomputed by the exact per-row fallback (margin overflow or error-bound self-check). */
int wspc_knn_fallback_rows(const void* workspace, int B, int N, int D, int* rows_out);

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

/* ------------------------------------------------ 1x1-conv (shared MLP) --- */
/* The reference's EdgeConv / per-point layers are tf_util.conv2d with a [1,1]
 * kernel followed by batch_norm_for_conv2d(is_dist=True) and ReLU
 * (Networks/dgcnn/utils/tf_util.py:115-173,502-535,577-592) applied to
 * get_edge_feature's output (tf_util.py:674-706) or to the previous layer.
 * Here a layer is one GEMM whose A operand is *synthesised on load* from what is
 * already in HBM (so edge features and normalised activations never round-trip
 * through memory) and whose epilogue folds the bias, the BN batch statistics,
 * the ReLU mask of the backward pass or the gather's scatter-add gradient.     */

/* A-operand loader modes */
#define WSPC_OP_PLAIN 0     /* a[row, c]                                                            */
#define WSPC_OP_BNRELU 1    /* relu(p[row,c]*sc[c] + sh[c]) (* dmask[row,c] * dscale)   tf_util.py:167-172,631-635 */
#define WSPC_OP_EDGE 2      /* row=(point i, slot r): [x_i | x_idx[i,r] - x_i], C = 2*Cx            tf_util.py:696-705 */
#define WSPC_OP_DY 3        /* c1[c]*p[row,c] + c2[c] + c3[c]*y[row,c]  (BN backward as an affine map; c1==NULL: p) */
#define WSPC_OP_DY_SPARSE 4 /* as DY with G[row,c] = (amax[cloud,c]==row%npts) ? dg[cloud,c] : 0   (max_pool2d grad) */
#define WSPC_OP_DY_MAXK 5   /* as DY with G synthesised from the max over k (tf.reduce_max grad, equal split among ties):
                               row=(point i, slot r); a = relu(y*sc+sh); p = MS (points, 2C) = [max_r a | dout/ties] from
                               wspc_maxk_bnrelu_bwd_stats; G = (a == MS[i,c] && MS[i,c] > 0) ? MS[i,C+c] : 0 */
#define WSPC_OP_IMG 6       /* a[row, c] as the pre-split image written by wspc_rows_image (p = image); accepted by
                             * wspc_conv1x1_rows_ws / wspc_conv1x1_pool_fwd where wspc_rows_image_supported says so */

typedef struct wspc_operand {
  const float* p;   /* PLAIN: matrix; BNRELU: pre-BN activation; EDGE: point features; DY: upstream gradient G;
                       DY_MAXK: MS (points, 2C) */
  long long ld;     /* leading dimension of p (floats) */
  int C;            /* logical channels */
  const float* sc;  /* BNRELU, DY_MAXK: gamma*rsqrt(var+eps) of the layer whose activation is formed */
  const float* sh;  /* BNRELU, DY_MAXK: beta - mean*sc */
  const float* dmask; /* BNRELU: optional dropout mask (rows, C) of 0/1 floats, or NULL */
  float dscale;     /* BNRELU: 1/keep_prob */
  const int32_t* idx; /* EDGE: (points, k) neighbour ids, local to the cloud */
  int k;            /* EDGE, DY_MAXK: neighbours (rows) per point */
  int npts;         /* EDGE / DY_SPARSE: points per cloud */
  const float* y;   /* DY*: pre-BN output of the layer being differentiated */
  long long ldy;
  const float* c1;  /* DY*: per-channel affine of the BN backward (wspc_bn_bwd_coeffs) */
  const float* c2;
  const float* c3;
  const float* dg;  /* DY_SPARSE: (clouds, C) gradient w.r.t. the max-pooled feature, already ReLU-gated */
  const int32_t* amax; /* DY_SPARSE: (clouds, C) first arg-max point */
} wspc_operand_t;

/* epilogue modes */
#define WSPC_EPI_STORE 0          /* out = acc + bias (+ rowbias[cloud])                                     */
#define WSPC_EPI_STORE_STATS 1    /* as STORE, and stats += per-column (sum, sum of squares)  -> tf.nn.moments */
#define WSPC_EPI_RELUMASK_STATS 2 /* out = acc*[relu(yprev*scp+shp)>0](*dmask*dscale); stats += (sum, sum*yprev) */
#define WSPC_EPI_ACCUM 3          /* out += acc                                                               */
#define WSPC_EPI_EDGE_SCATTER 4   /* acc=[dE_c|dE_d]: dx[i]+=dE_c-dE_d, dx[idx]+=dE_d (Gather/Concat/Sub grads) */

typedef struct wspc_epilogue {
  float* out;
  long long ldo;
  const float* bias;     /* (N) or NULL */
  const float* rowbias;  /* (clouds, ldrb) per-cloud additive row or NULL (folded tiled global feature) */
  int rb_rows;           /* rows per cloud for rowbias */
  long long ldrb;
  double* stats;         /* (2, N) fp64 accumulators, caller zeroes them */
  const float* yprev;    /* RELUMASK: pre-BN activation of the producing layer (rows, ldyp) */
  long long ldyp;
  const float* scp;
  const float* shp;
  const float* dmask;    /* RELUMASK: optional dropout mask of the producing layer's output */
  float dscale;
  float* dx;             /* EDGE_SCATTER: (points, lddx) gradient accumulator (atomics) */
  long long lddx;
  const int32_t* idx;
  int k;
  int npts;
} wspc_epilogue_t;

/* out(M,N) = A(M,K) * Bm(K,N) with A read through `a_mode` and the result written through
 * `epi_mode`.  Bm is row-major (K,N) with leading dimension ldb, or, if b_transposed, stored
 * as (N,K) row-major (i.e. the layer's own weight matrix used for the data gradient).
 * Forward:  A = PLAIN/BNRELU/EDGE, Bm = W(Cin,Cout)           tf_util.py:160-165
 * Backward: A = DY/DY_SPARSE,      Bm = W^T (b_transposed=1)   Conv2DBackpropInput [TF] */
int wspc_conv1x1_rows(const wspc_operand_t* A, int a_mode, const float* Bm, long long ldb, int b_transposed,
                      long long M, int N, int K, const wspc_epilogue_t* epi, int epi_mode,
                      wspc_stream_t stream);

/* Same operation with caller-provided scratch: when K is processed in several chunks (K > 128) the tensor-core
 * kernel stages a pre-split (bf16 hi/lo) image of the weights with one bulk copy per chunk instead of re-splitting
 * the fp32 weights for every 128-row tile; the image lives in `workspace` (wspc_conv1x1_rows_workspace_bytes(N, K)
 * bytes, 16-byte aligned; 0 means no scratch is needed).  workspace == NULL selects the in-kernel split. */
size_t wspc_conv1x1_rows_workspace_bytes(int N, int K);
int wspc_conv1x1_rows_ws(const wspc_operand_t* A, int a_mode, const float* Bm, long long ldb, int b_transposed,
                         long long M, int N, int K, const wspc_epilogue_t* epi, int epi_mode, void* workspace,
                         size_t workspace_bytes, wspc_stream_t stream);

/* Kernel selection for wspc_conv1x1_rows: 0 = auto (tcgen05 tensor-core kernel for eligible shapes, CUDA-core
 * kernel otherwise), 1 = CUDA-core kernel only (used by the tests to A/B the two device paths).  Returns the
 * previous setting.  Both paths are sm_100a CUDA; neither is a CPU fallback. */
int wspc_set_gemm_path(int path);

/* dW(K1,K2) = sum_rows A(row,:)^T dY(row,:), db(K2) = sum_rows dY(row,:)   (Conv2DBackpropFilter,
 * BiasAddGrad [TF]).  A through a_mode (PLAIN/BNRELU/EDGE), dY through g_mode (DY/DY_SPARSE).
 * Deterministic: row slabs are reduced in a fixed order in fp64.  db may be NULL. */
size_t wspc_conv1x1_wgrad_workspace_bytes(int K1, int K2);
int wspc_conv1x1_wgrad(const wspc_operand_t* A, int a_mode, const wspc_operand_t* G, int g_mode, long long M,
                       float* dW, float* db, void* workspace, size_t workspace_bytes, wspc_stream_t stream);

/* Weight gradient and data gradient of one conv2d in a single pass over the rows (Conv2DBackpropFilter + BiasAddGrad +
 * Conv2DBackpropInput [TF]): dW(K1,K2), db(K2) as wspc_conv1x1_wgrad; dA(M,K1) = dY * W^T with W the layer's (K1,K2) weight
 * matrix (leading dimension ldw), written through `epi` exactly as wspc_conv1x1_rows would with WSPC_EPI_RELUMASK_STATS.
 * The tensor-core kernel stages G / y / the previous activation once for both products (A = BNRELU, G = DY or DY_MAXK,
 * K1 <= 64, K2 <= 256); other shapes run the two separate kernels.  Workspace as wspc_conv1x1_wgrad_workspace_bytes. */
int wspc_conv1x1_bwd_fused(const wspc_operand_t* A, int a_mode, const wspc_operand_t* G, int g_mode, long long M,
                           const float* W, long long ldw, const wspc_epilogue_t* epi, float* dW, float* db,
                           void* workspace, size_t workspace_bytes, wspc_stream_t stream);

/* ------------------------------------------------- batch norm + pooling --- */
/* batch_norm_dist_template (tf_util.py:502-535).  stats = (2,C) fp64 (sum, sum of squares) produced by
 * WSPC_EPI_STORE_STATS over `rows` rows.  training: mean/biased var from stats, pop <- pop*decay +
 * batch*(1-decay) (:524-525); else the population statistics are used (:530).  Outputs the folded
 * affine sc = gamma*rsqrt(var+eps), sh = beta - mean*sc consumed by WSPC_OP_BNRELU, and (optionally)
 * mean / invstd for the backward pass. */
int wspc_bn_finalize(const double* stats, int C, double rows, const float* gamma, const float* beta, float eps,
                     float decay, int training, float* pop_mean, float* pop_var, float* sc, float* sh,
                     float* save_mean, float* save_invstd, wspc_stream_t stream);
/* BN backward folded into dy = c1*G + c2 + c3*y; stats = (2,C) fp64 (sum G, sum G*y).
 * Also emits dgamma, dbeta. */
int wspc_bn_bwd_coeffs(const double* stats, int C, double rows, const float* gamma, const float* mean,
                       const float* invstd, float* c1, float* c2, float* c3, float* dgamma, float* dbeta,
                       wspc_stream_t stream);
/* out[p, c] = max_r relu(y[p,r,c]*sc[c]+sh[c])  == tf.reduce_max(relu(bn(.)), axis=-2)
 * (DGCNN_S3DIS.py:46,62,78; transform_nets.py:27).  y (P,k,C) dense, out strided by ldo. */
int wspc_maxk_bnrelu_fwd(const float* y, const float* sc, const float* sh, long long P, int k, int C, float* out,
                         long long ldo, wspc_stream_t stream);
/* gradient of the above w.r.t. the BN output, already ReLU-masked: G (P,k,C); equal split among ties
 * [TF _MinOrMaxGrad]; stats (2,C) += (sum G, sum G*y). */
int wspc_maxk_bnrelu_bwd(const float* y, const float* sc, const float* sh, const float* out, long long ldo,
                         const float* dout, long long lddo, long long P, int k, int C, float* G, double* stats,
                         wspc_stream_t stream);
/* Statistics-only variant: writes MS (P, 2C) = [max over k of relu(bn(y)) | dout / #ties] and the BN-backward sums instead of
 * the (P*k, C) gradient; the consumers synthesise G on load (WSPC_OP_DY_MAXK, wspc_edge_combine_bwd_maxk). */
int wspc_maxk_bnrelu_bwd_stats(const float* y, const float* sc, const float* sh, const float* out, long long ldo,
                               const float* dout, long long lddo, long long P, int k, int C, float* MS, double* stats,
                               wspc_stream_t stream);
/* g[b,c] = max_n relu(bn(y[b,n,c])), amax = first arg-max  == tf_util.max_pool2d([N,1]) (tf_util.py:357-380) */
int wspc_maxn_bnrelu_fwd(const float* y, const float* sc, const float* sh, int B, int N, int C, float* g,
                         int32_t* amax, wspc_stream_t stream);
/* conv2d 1x1 -> BN -> ReLU -> max_pool2d([N,1]) (adj_conv7 + the global feature, DGCNN_S3DIS.py:80-85) WITHOUT writing the
 * (M, N) conv output: one pass of the warp-specialised tcgen05 GEMM accumulates the BN sums into `stats` (2, N; the caller
 * zeroes them) and, per (cloud, column), a packed key (order-preserving bits of y or -y | ~row) of the row with the largest
 * (gamma >= 0) or smallest (gamma < 0) pre-BN value -- BN and ReLU are monotone per column, so that row is max_pool2d's first
 * arg-max.  `keys` is (M / npts, N) and is cleared by the call.  wspc_conv1x1_pool_supported says whether a shape is eligible
 * (K > 128, K % 32 == 0, npts % 128 == 0, >= 592 row tiles); otherwise run wspc_conv1x1_rows + wspc_maxn_bnrelu_fwd.
 * workspace: wspc_conv1x1_rows_workspace_bytes(N, K).  wspc_maxn_from_keys then gives g = relu(bn(y*)), amax and y* (B, C). */
int wspc_conv1x1_pool_supported(long long M, int N, int K, int npts);
int wspc_conv1x1_pool_fwd(const wspc_operand_t* A, int a_mode, const float* W, long long ldw, long long M, int N, int K,
                          int npts, const float* bias, const float* gamma, double* stats, unsigned long long* keys,
                          void* workspace, size_t workspace_bytes, wspc_stream_t stream);
int wspc_maxn_from_keys(const unsigned long long* keys, const float* gamma, const float* sc, const float* sh, int B, int C,
                        float* g, int32_t* amax, float* ymax, wspc_stream_t stream);
/* Pre-split operand image of a (M, K) fp32 matrix that several GEMMs read (the concatenated EdgeConv features feed adj_conv7
 * and seg/conv1, DGCNN_S3DIS.py:80,93): per 128-row tile and 32-channel chunk the bf16 hi and lo halves in the tensor core's
 * shared-memory layout, so the warp-specialised GEMM stages its A operand with one bulk copy per chunk and no CUDA-core work.
 * K % 32 == 0, x 16-byte aligned rows.  wspc_rows_image_supported(M, N, K): 1 if wspc_conv1x1_rows_ws takes WSPC_OP_IMG for
 * an (M, K) x (K, N) product. */
size_t wspc_rows_image_bytes(long long M, int K);
int wspc_rows_image_supported(long long M, int N, int K);
int wspc_rows_image(const float* x, long long ldx, long long M, int K, void* image, wspc_stream_t stream);
/* dg = dgin*[g>0]; stats (2,C) = (sum, sum*y at the arg-max rows): the sparse gradient of max_pool2d */
int wspc_maxn_bwd_gate(const float* g, const float* dgin, const int32_t* amax, const float* y, int B, int N, int C,
                       float* dg, double* stats, wspc_stream_t stream);
/* S[b,c] = sum_n dY[b,n,c], dY through a WSPC_OP_DY operand: gradient of the tiled global feature (tf.tile grad) */
int wspc_cloud_colsum(const wspc_operand_t* G, int B, int N, float* S, wspc_stream_t stream);

/* ----------------------------------------------------- head + weak losses --- */
/* softmax + masked CE + Siamese + inexact (MIL) + manifold smoothness, values and dZ in one call.
 * S3DIS_DGCNN_trainer.py:85-102,120-137 / ShapeNet_DGCNN_trainer.py:85-100,115-133;
 * Util/SmoothConstraint.py:155-165 (the kNN graph comes from wspc_knn_fused with WSPC_DIST_SMOOTH).
 *   Z (B,N,C) logits, Y (B,N,C) one-hot float, Mask (B,N); sm_idx/sm_dist (B,N,knn) or NULL
 *   full=1: loss = seg + siamese + inexact + smooth (ramp-up gate open); full=0: seg only ("Plain");
 *   full=2: all four VALUES, gradient of the seg term only (Full graph with the gate closed, S3DIS_DGCNN_trainer.py:100-102)
 *   P (B,N,C) softmax out; dZ (B,N,C) d loss / d Z (required if want_grad or full)
 *   losses[5] = {seg, siamese, inexact, smooth, total}            C <= 64 */
size_t wspc_head_losses_workspace_bytes(int B, int N, int C);
int wspc_head_losses(const float* Z, const float* Y, const float* Mask, const int32_t* sm_idx, const float* sm_dist,
                     int B, int N, int C, int knn, float gamma, float siam_w, int full, int want_grad, float* P,
                     float* dZ, float* losses, void* workspace, size_t workspace_bytes, wspc_stream_t stream);

/* ------------------------------------------- factored EdgeConv first layer --- */
/* The first conv2d of every EdgeConv block reads e_ij = [x_i | x_j - x_i] (tf_util.get_edge_feature,
 * tf_util.py:674-706; DGCNN_S3DIS.py:34-39,50-55,66-71; DGCNN_ShapeNet.py:24-37; transform_nets.py:17-20) with
 * W = [W1; W2] (2*Cx, Cout).  y_ij = x_i (W1 - W2) + x_j W2 + b = u_i + v_j + b, so the P*k-row GEMM becomes one
 * P-row GEMM [u | v] = X Wc (wspc_conv1x1_rows) plus the gather-add below; the backward pass is the transpose.
 * Cout must be 64.
 *   wspc_edge_split_weights : Wc (Cx, 2*Cout) = [W1 - W2 | W2]
 *   wspc_edge_combine_fwd   : y (P*k, Cout) = U[i] + V[cloud(i)*npts + idx[i,j]] + bias; stats[2][Cout] (fp64, may be
 *                             NULL) += column sums / sums of squares (batch-norm moments, tf_util.py:521-522)
 *   wspc_edge_combine_bwd   : dy = c1*G + c2 + c3*y (c1 == NULL: dy = G); DUV[i, 0:Cout] = sum_j dy_ij;
 *                             DUV[cloud(i)*npts + idx[i,j], Cout:2*Cout] += dy_ij (red.global.add.v4, caller zeroes)
 *   wspc_edge_merge_wgrad   : dW (2*Cx, Cout) from dWc (Cx, 2*Cout): dW1 = dWa, dW2 = dWb - dWa; db = dbc[0:Cout] */
int wspc_edge_split_weights(const float* W, int Cx, int Cout, float* Wc, wspc_stream_t stream);
int wspc_edge_combine_fwd(const float* UV, long long ldu, const int32_t* idx, const float* bias, long long P, int k,
                          int npts, int Cout, float* y, double* stats, wspc_stream_t stream);
/* as wspc_edge_combine_fwd, additionally writing MM (P, 2*Cout) = [max_r y | min_r y] over each point's k rows; when the
 * layer's activation goes straight into tf.reduce_max over k (adj_conv5, DGCNN_S3DIS.py:66-78) wspc_maxk_from_extrema
 * turns it into max_r relu(sc*y_r + sh) without re-reading y (bit-identical: fmaf(y, sc, sh) is monotone in y). */
int wspc_edge_combine_fwd_extrema(const float* UV, long long ldu, const int32_t* idx, const float* bias, long long P, int k,
                                  int npts, int Cout, float* y, double* stats, float* MM, wspc_stream_t stream);
int wspc_maxk_from_extrema(const float* MM, const float* sc, const float* sh, long long P, int C, float* out, long long ldo,
                           wspc_stream_t stream);
int wspc_edge_combine_bwd(const float* G, const float* y, const float* c1, const float* c2, const float* c3,
                          const int32_t* idx, long long P, int k, int npts, int Cout, float* DUV, long long ldd,
                          wspc_stream_t stream);
/* as wspc_edge_combine_bwd with G synthesised from the max over k (WSPC_OP_DY_MAXK; MS from wspc_maxk_bnrelu_bwd_stats) */
int wspc_edge_combine_bwd_maxk(const float* y, const float* c1, const float* c2, const float* c3, const float* sc,
                               const float* sh, const float* MS, const int32_t* idx, long long P, int k, int npts, int Cout,
                               float* DUV, long long ldd, wspc_stream_t stream);
int wspc_edge_merge_wgrad(const float* dWc, const float* dbc, int Cx, int Cout, float* dW, float* db,
                          wspc_stream_t stream);

/* ----------------------------------------- fused EdgeConv blocks (no (B*N*k, C) tensor in HBM) --- */
/* One EdgeConv block of the reference graphs is  get_edge_feature -> conv2d + BN + ReLU [-> conv2d + BN + ReLU] ->
 * tf.reduce_max over k  (tf_util.py:674-706,115-173,502-535; DGCNN_S3DIS.py:32-46,48-62,64-78; DGCNN_ShapeNet.py:32-78).
 * With the factored first layer (above) every edge row y1_ij = u_i + v_j + b1 is a function of two rows of the
 * L2-resident UV (P, 128) = [u | v], so forward and backward recompute the edge tensors on chip instead of storing
 * them.  All channel counts are 64.  P*k < 2^31, P % npts == 0.
 *
 *   wspc_edge_gather_stats      one gather sweep over the edges: stats (2,64) fp64 += (sum y1, sum y1^2)  [tf.nn.moments];
 *                               MM (P,128) = [max_j y1 | min_j y1] (single-conv block: feeds wspc_maxk_from_extrema);
 *                               SS (P,128) = [S_i = sum_j v_j | SU_p += u_i for every edge i->p] and deg (P) += 1 per
 *                               incoming edge (caller zeroes SS[:,64:] and deg; needed by the backward pass only).
 *                               Any of stats / MM / (SS, deg) may be NULL.
 *   wspc_edgeconv2_fwd          two-conv block: a1 = relu(bn1(y1)) (sc1/sh1 from wspc_bn_finalize) -> y2 = a1 W2 + b2 on
 *                               tcgen05 (bf16 hi/lo split, 3 passes, fp32 accumulate) -> MM (P,128) = [max_j y2 | min_j y2]
 *                               and stats2 (2,64) fp64 += (sum y2, sum y2^2) (NULL at inference); 2 <= k <= 128.
 *                               reduce_max(relu(bn2(y2))) = wspc_maxk_from_extrema(MM, sc2, sh2).
 *   wspc_maxk_extrema_bwd_prep  MS (P,128) = [out > 0 ? out : -1 | out > 0 ? dout : 0] and the BN-2 backward sums
 *                               stats (2,64) += (sum G, sum G*y2) of the max-over-k gradient G (it lives on the extremal rows).
 *   wspc_edgeconv2_bwd          recomputes a1, y2 per tile (bit-identical to the forward), G = tie-split reduce_max gradient
 *                               [TF _MinOrMaxGrad], dy2 = c1 G + c2 + c3 y2 (wspc_bn_bwd_coeffs of layer 2);
 *                               dW2 (64,64) = a1^T dy2;  g1 = (dy2 W2^T) * [a1 > 0];
 *                               TS (P,128): TS[i,0:64] = sum_j g1_ij,  TS[p,64:128] += g1_ij for every edge i->p
 *                               (red.global.add, caller zeroes that half).
 *   wspc_edge1_bwd              single-conv block: the same TS from (out, dout) of the max over k directly (k <= 64).
 *   wspc_edge_bwd_stats         BN-1 backward sums from TS: bstats (2,64) fp64 += (sum g1, sum g1*y1).
 *   wspc_edge_bwd_finalize      DUV (P,128) = [du | dv] of dy1 = c1 g1 + c2 + c3 y1 (closed form from TS, SS, deg, UV);
 *                               the P-row GEMMs of the factored layer (wspc_conv1x1_wgrad / _rows) finish the block.
 *   wspc_bn_bias_grad           db = c1 sum G + rows c2 + c3 sum y  (bias of a conv followed by BN; fstats = forward moments) */
int wspc_edge_gather_stats(const float* UV, long long ldu, const int32_t* idx, const float* bias, long long P, int k,
                           int npts, int Cout, double* stats, float* MM, float* SS, float* deg, wspc_stream_t stream);
int wspc_edgeconv2_fwd(const float* UV, long long ldu, const int32_t* idx, const float* bias1, const float* sc1,
                       const float* sh1, const float* W2, const float* bias2, long long P, int k, int npts, int C1, int C2,
                       double* stats2, float* MM, wspc_stream_t stream);
int wspc_maxk_extrema_bwd_prep(const float* MM, const float* sc, const float* out, long long ldo, const float* dout,
                               long long lddo, long long P, int C, float* MS, double* stats, wspc_stream_t stream);
size_t wspc_edgeconv2_bwd_workspace_bytes(void);
int wspc_edgeconv2_bwd(const float* UV, long long ldu, const int32_t* idx, const float* bias1, const float* sc1,
                       const float* sh1, const float* W2, const float* bias2, const float* sc2, const float* sh2,
                       const float* c1, const float* c2, const float* c3, const float* MS, long long P, int k, int npts,
                       int C1, int C2, float* TS, float* dW2, void* workspace, size_t workspace_bytes,
                       wspc_stream_t stream);
/* The same two calls with a routing export for parity tests (NULL = off): which rows attained each pooled maximum.
 *   edgeconv2: routing_out (P*k, 2) uint32, bit c of word h = row attains the maximum of channel 32 h + c (k >= 8)
 *   edge1:     routing_out (P, 64) uint64, bit j = row j of the point attains the maximum of that channel */
int wspc_edgeconv2_bwd_ex(const float* UV, long long ldu, const int32_t* idx, const float* bias1, const float* sc1,
                          const float* sh1, const float* W2, const float* bias2, const float* sc2, const float* sh2,
                          const float* c1, const float* c2, const float* c3, const float* MS, long long P, int k, int npts,
                          int C1, int C2, float* TS, float* dW2, uint32_t* routing_out, void* workspace,
                          size_t workspace_bytes, wspc_stream_t stream);
int wspc_edge1_bwd_ex(const float* UV, long long ldu, const int32_t* idx, const float* bias, const float* sc, const float* sh,
                      const float* out, long long ldo, const float* dout, long long lddo, long long P, int k, int npts,
                      int Cout, float* TS, uint64_t* routing_out, wspc_stream_t stream);
int wspc_edge1_bwd(const float* UV, long long ldu, const int32_t* idx, const float* bias, const float* sc, const float* sh,
                   const float* out, long long ldo, const float* dout, long long lddo, long long P, int k, int npts,
                   int Cout, float* TS, wspc_stream_t stream);
int wspc_edge_bwd_stats(const float* TS, const float* UV, long long ldu, const float* bias, long long P, int Cout,
                        double* bstats, wspc_stream_t stream);
int wspc_edge_bwd_finalize(const float* TS, const float* SS, const float* deg, const float* UV, long long ldu,
                           const float* bias, const float* c1, const float* c2, const float* c3, long long P, int k,
                           int Cout, float* DUV, long long ldd, wspc_stream_t stream);
/* p[r, col0:col0+ncols] = 0 for r < rows of a (rows, ld) fp32 matrix (the scatter halves of SS / TS above) */
int wspc_zero_cols(float* p, long long ld, int col0, int ncols, long long rows, wspc_stream_t stream);
int wspc_bn_bias_grad(const double* fstats, const double* bstats, const float* c1, const float* c2, const float* c3, int C,
                      double rows, float* db, wspc_stream_t stream);

/* ------------------- backward of conv2d 1x1 -> BN -> ReLU -> max over the points of a cloud --- */
/* adj_conv7 + maxpool (DGCNN_S3DIS.py:80-85, DGCNN_ShapeNet.py:80-85) and tconv3 + tmaxpool (transform_nets.py:29-34).
 * max_pool2d's gradient is sparse (one point per cloud and channel), the BN backward is affine (dy = c1*G + c2 + c3*y)
 * and y = A W + b is linear in the layer input A (P, cin), so
 *   dA = A (W diag(c3) W^T) + 1 (W t)^T + sparse(c1*G) W^T,   dW = (A^T A) W diag(c3) + (A^T 1) t^T + A^T sparse(c1*G),
 *   db = c3 * (W^T A^T 1) + P t + 1^T sparse(c1*G),           t = c2 + c3*b,
 * i.e. (P,cin)x(cin,cin) GEMMs through wspc_conv1x1_rows / wspc_conv1x1_wgrad instead of (P,cout) ones, and no read of
 * y.  These entry points are the glue:
 *   wspc_poolconv_coeffs   : t (cout), Wsc (cin,cout) = W diag(c3)
 *   wspc_poolconv_sparse   : dx[b*N + amax[b,c], :] += c1[c] dg[b,c] W[:,c] (atomics; dx may be NULL);
 *                            sW (cin,cout) = A^T sparse(c1*G), sdb (cout) = 1^T sparse(c1*G)  (deterministic)
 *   wspc_poolconv_finalize : dW = T + colsum t^T + sW with T = (A^T A) Wsc, colsum = A^T 1;  db as above (may be NULL) */
int wspc_poolconv_coeffs(const float* W, const float* b, const float* c2, const float* c3, int cin, int cout, float* t,
                         float* Wsc, wspc_stream_t stream);
int wspc_poolconv_sparse(const float* dg, const int32_t* amax, const float* c1, const float* W, const float* A,
                         long long lda, int B, int N, int cin, int cout, float* dx, long long lddx, float* sW, float* sdb,
                         wspc_stream_t stream);
/* out[row, 0:C] = v[0:C] for every row (the constant row of dA when no earlier GEMM writes dA) */
int wspc_fill_rows(float* out, long long ldo, const float* v, long long rows, int C, wspc_stream_t stream);
int wspc_poolconv_finalize(const float* T, const float* colsum, const float* t, const float* sW, const float* sdb,
                           const float* Wsc, int cin, int cout, double rows, float* dW, float* db, wspc_stream_t stream);

/* ------------------------------------------------------------ optimiser --- */
/* tf.train.AdamOptimizer update on flat fp32 buffers (eps not bias-corrected);
 * lr_t = lr*sqrt(1-b2^t)/(1-b1^t) is computed by the caller; gscale multiplies g (1/world_size). */
int wspc_adam_tf(float* p, const float* g, float* m, float* v, long long n, float lr_t, float b1, float b2,
                 float eps, float gscale, wspc_stream_t stream);
/* tf.nn.dropout keep-mask (0/1 floats) from a Philox-4x32-10 stream: keep iff floor(keep + U) == 1 */
int wspc_dropout_mask(float* mask, long long n, float keep, uint64_t seed, uint64_t offset, wspc_stream_t stream);
int wspc_zero(void* ptr, size_t bytes, wspc_stream_t stream);

/* X' = X (T + I) per cloud: point_cloud_transformed = tf.matmul(point_cloud, transform)
 * (ShapeNet/DGCNN_ShapeNet.py:29) with the identity of transform_nets.py:51 added when add_eye != 0.
 * X, Xt (B,N,3); T, dT (B,3,3).  bwd: dT[b] = X[b]^T dXt[b]. */
int wspc_transform_points_fwd(const float* X, const float* T, int B, int N, int add_eye, float* Xt, wspc_stream_t stream);
int wspc_transform_points_bwd(const float* X, const float* dXt, int B, int N, float* dT, wspc_stream_t stream);

/* Stand-alone manifold smoothness loss on a kNN graph (Util/SmoothConstraint.py:155-165):
 * loss = mean_{b,n,r} exp(-dist/gamma) * mean_c (Z[b,n,c] - Z[b,idx[b,n,r],c])^2 ; dZ (zeroed by caller) or NULL. */
int wspc_smooth_loss(const float* Z, const int32_t* idx, const float* dist, int B, int N, int C, int knn, float gamma,
                     float* dZ, float* loss, void* workspace, size_t workspace_bytes, wspc_stream_t stream);
/* The sibling variants of Util/SmoothConstraint.py on the same kernel.  flags:
 *   WSPC_SMOOTH_SUM_C      sum_c instead of mean_c of the squared differences (:31, :118-122, :215)
 *   WSPC_SMOOTH_WEIGHTS    `dist` already holds the edge weights W (Loss_SpatialSmooth, :9-33); gamma is ignored
 *   WSPC_SMOOTH_GLOBAL_SS  Loss_SpatialSmooth_SelfContain (:64-65): every weight multiplies the sum of squared differences over
 *                          ALL edges and channels: loss = (sum W)(sum ss) / (B N knn); forward only (dZ must be NULL)
 * idx_match (B,N,knn) or NULL: an edge slot counts only where idx_match == idx (knn_mask = equal(Ind_xyz, Ind_rgb), :113). */
#define WSPC_SMOOTH_SUM_C 1
#define WSPC_SMOOTH_WEIGHTS 2
#define WSPC_SMOOTH_GLOBAL_SS 4
int wspc_smooth_loss_ex(const float* Z, const int32_t* idx, const float* dist, const int32_t* idx_match, int B, int N, int C,
                        int knn, float gamma, int flags, float* dZ, float* loss, void* workspace, size_t workspace_bytes,
                        wspc_stream_t stream);

/* ------------------------------------------------- test-time label propagation --- */
/* Lsym = D^-1/2 (diag(d + 1e-8) - W) D^-1/2, W = exp(-scale_xyz d_xyz) * exp(-scale_rgb d_rgb)
 * (Tool.TF_Computation.LaplacianMatSym_XYZRGB_DirectComp, Util/Tool.py:435-468; scales 1e3 / 1e1).
 * X (B,N,D1), RGB (B,N,D2), D <= 3; deg_ws (B,N) scratch; Lout (B,N,N). */
int wspc_laplacian_sym(const float* X, const float* RGB, int B, int N, int D1, int D2, float scale_xyz,
                       float scale_rgb, float* deg_ws, float* Lout, wspc_stream_t stream);
/* LabelPropagation_TF.SolveLabelProp (Util/ProbLabelPropagation.py:19-23,38-57) for B independent blocks at once
 * (the reference's Test loop calls it block by block, S3DIS_DGCNN_trainer.py:541-544):
 *   w = 1 - H_2(G)/log_2 K;  Y = beta (alpha L + beta diag(w) + 1e-5 I)^-1 diag(w) G;  Yprob = Y / sum_k Y.
 * Each SPD system is solved by Jacobi-preconditioned CG on all K right-hand sides; the matrix alpha L + diag is never formed
 * (q = alpha L p + d p).  A block stops when every class column has |r| <= tol |b|; the decision is taken ON THE DEVICE per
 * block and the call never waits for the GPU: iterations are enqueued in chunks, the host only looks (cudaEventQuery) at
 * convergence counters that have already landed in pinned memory and stops enqueuing once all B blocks are done.
 *   L (B,N,N), G (B,N,K) -> Y, Yprob (B,N,K), w (B,N); 2 <= K <= 64, any N (N % 4 == 0 takes the vector loads)
 *   iters (B) iterations used; resid (B) max_c |r_c|/|b_c| at exit; done (B) 1 = converged, 0 = stopped at max_iter
 *   (device arrays: the caller checks them when it reads the result). */
size_t wspc_lp_blocks_workspace_bytes(int B, int N, int K, int max_iter);
int wspc_lp_blocks(const float* L, const float* G, int B, int N, int K, float alpha, float beta, int max_iter, float tol,
                   float* Y, float* Yprob, float* w, int32_t* iters, float* resid, int32_t* done, void* workspace,
                   size_t workspace_bytes, wspc_stream_t stream);
/* One system = wspc_lp_blocks with B = 1.  iters_out != NULL asks for the iteration count on the host and costs one
 * copy + stream wait at the END of the call; NULL keeps it asynchronous.  max_iter is capped at 4096. */
size_t wspc_lp_solve_workspace_bytes(int N, int K);
int wspc_lp_solve(const float* L, const float* G, int N, int K, float alpha, float beta, int max_iter, float tol,
                  float* Y, float* Yprob, float* w, int* iters_out, void* workspace, size_t workspace_bytes,
                  wspc_stream_t stream);

/* ----------------------------------------------------------- unfused API ops --- */
/* edge=0: out[b,n,r,:] = X[b, idx[b,n,r], :]            Tool.batch_gather_v1 (Util/Tool.py:72-104)
 * edge=1: out[b,n,r,:] = [X[b,n,:] | X[b,idx,:]-X[b,n,:]]  tf_util.get_edge_feature (tf_util.py:674-706) */
int wspc_gather(const float* X, const int32_t* idx, int B, int N, int k, int C, long long ldx, int edge, float* out,
                wspc_stream_t stream);
/* out = y*sc + sh (relu optional): tf.nn.batch_normalization with the folded affine of wspc_bn_finalize */
int wspc_bn_apply(const float* y, const float* sc, const float* sh, long long rows, int C, int relu, float* out,
                  wspc_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* WSPC_H_ */
