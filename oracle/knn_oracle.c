/*
 * knn_oracle.c — CPU restatement of the reference's distance + kNN path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under weaksuppointcloudseg_b200/ may
 * import, link or call this file; only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs do, and only as the checker
 * or the timed CPU baseline.
 *
 * PINNING: the reference (alex-xun-xu/WeakSupPointCloudSeg) ships no tests or
 * golden vectors for this path and its arithmetic lives in TensorFlow 1.14
 * (README.md:21), which is absent from /root/reference and cannot be installed
 * here.  This file restates the *published* semantics of the TF ops at the
 * reference's call sites; it is checked against the reference's own
 * tf_util.pairwise_distance / knn and SmoothConstraint code executed on the
 * op-level TF stand-in (tests/golden/ref_unit_ops.npz,
 * tests/test_reference_golden_cpu.py): distances to 1e-5 of the matrix scale,
 * neighbour lists equal except on last-bit ties (TensorFlow's matmul summation
 * order is unspecified, so bit-level index parity with a TF run is undefined;
 * "bit-exact kNN" in this repository means bit-exact against the canonical
 * fp32 arithmetic fixed in SURVEY.md App. A and implemented here):
 *
 *   pairwise distance, tf_util flavour   Networks/dgcnn/utils/tf_util.py:652-657
 *       inner = -2 * (X X^T); sq = sum(x^2); D = (sq_i + inner_ij) + sq_j
 *   pairwise distance, smooth flavour    Util/SmoothConstraint.py:144-148
 *       D = (X2_i + Y2_j) - 2*XY_ij ; D<0 -> 0     (also Util/Tool.py:444-448)
 *   kNN                                  Networks/dgcnn/utils/tf_util.py:669-671,
 *                                        Util/SmoothConstraint.py:154
 *       tf.nn.top_k(-D, k): k smallest D, ascending, ties -> lower index
 *
 * Canonical arithmetic: every dot product (and the squared norm, which is the
 * dot of a row with itself) is one sequential chain acc = fmaf(a_c, b_c, acc)
 * over c = 0..D-1 starting from +0, all other operations are single fp32
 * operations rounded to nearest.  Build with -ffp-contract=off (see Makefile).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define FLAVOUR_TFUTIL 0
#define FLAVOUR_SMOOTH 1

static inline float dot_chain(const float* a, const float* b, int D) {
  float acc = 0.0f;
  for (int c = 0; c < D; ++c) acc = fmaf(a[c], b[c], acc);
  return acc;
}

static inline float dist_value(int flavour, float sqi, float sqj, float dot) {
  if (flavour == FLAVOUR_TFUTIL) {
    volatile float inner = -2.0f * dot;   /* tf_util.py:654 */
    volatile float t = sqi + inner;       /* tf_util.py:657, left to right */
    return t + sqj;
  } else {
    volatile float t = sqi + sqj;         /* SmoothConstraint.py:147  X_2 + Y_2 */
    volatile float two_xy = 2.0f * dot;
    float d = t - two_xy;
    return d > 0.0f ? d : 0.0f;           /* SmoothConstraint.py:148 */
  }
}

/* adj: (B, N, N) */
int oracle_pairwise_distance(const float* x, int B, int N, int ldx, int coff, int D, int flavour, float* adj) {
#pragma omp parallel for collapse(2) schedule(static)
  for (int b = 0; b < B; ++b) {
    for (int i = 0; i < N; ++i) {
      const float* xb = x + (size_t)b * N * ldx + coff;
      const float* xi = xb + (size_t)i * ldx;
      const float sqi = dot_chain(xi, xi, D);
      float* row = adj + ((size_t)b * N + i) * N;
      for (int j = 0; j < N; ++j) {
        const float* xj = xb + (size_t)j * ldx;
        row[j] = dist_value(flavour, sqi, dot_chain(xj, xj, D), dot_chain(xi, xj, D));
      }
    }
  }
  return 0;
}

/* k smallest of one row, ascending, ties -> lower index (stable). Scanning j in
 * ascending order with a strict '<' test keeps earlier indices ahead of later
 * equal values, which is exactly tf.nn.top_k's tie rule on -D. */
static void select_row(const float* row, int n, int k, int32_t* idx, float* val) {
  int m = 0; /* current list length */
  for (int j = 0; j < n; ++j) {
    const float d = row[j];
    if (m == k && !(d < val[k - 1])) continue;
    int p = (m < k) ? m : k - 1;
    while (p > 0 && d < val[p - 1]) {
      val[p] = val[p - 1];
      idx[p] = idx[p - 1];
      --p;
    }
    val[p] = d;
    idx[p] = j;
    if (m < k) ++m;
  }
}

int oracle_topk_rows(const float* adj, long long rows, int ncols, int k, int32_t* idx, float* vals) {
  if (k < 1 || k > ncols) return -1;
#pragma omp parallel
  {
    float* tmp = (float*)malloc(sizeof(float) * (size_t)k);
#pragma omp for schedule(static)
    for (long long r = 0; r < rows; ++r) {
      select_row(adj + (size_t)r * ncols, ncols, k, idx + (size_t)r * k, tmp);
      if (vals) memcpy(vals + (size_t)r * k, tmp, sizeof(float) * (size_t)k);
    }
    free(tmp);
  }
  return 0;
}

/* fused restatement (no N x N buffer): idx (B,N,k), dist (B,N,k) or NULL */
int oracle_knn(const float* x, int B, int N, int ldx, int coff, int D, int k, int flavour, int32_t* idx,
               float* dist) {
  if (k < 1 || k > N) return -1;
  float* sq = (float*)malloc(sizeof(float) * (size_t)B * N);
  if (!sq) return -2;
#pragma omp parallel for schedule(static)
  for (long long r = 0; r < (long long)B * N; ++r) {
    const float* xr = x + (size_t)r * ldx + coff;
    sq[r] = dot_chain(xr, xr, D);
  }
#pragma omp parallel
  {
    float* row = (float*)malloc(sizeof(float) * (size_t)N);
    float* tmp = (float*)malloc(sizeof(float) * (size_t)k);
#pragma omp for collapse(2) schedule(dynamic, 16)
    for (int b = 0; b < B; ++b) {
      for (int i = 0; i < N; ++i) {
        const float* xb = x + (size_t)b * N * ldx + coff;
        const float* xi = xb + (size_t)i * ldx;
        const float sqi = sq[(size_t)b * N + i];
        for (int j = 0; j < N; ++j)
          row[j] = dist_value(flavour, sqi, sq[(size_t)b * N + j], dot_chain(xi, xb + (size_t)j * ldx, D));
        select_row(row, N, k, idx + ((size_t)b * N + i) * k, tmp);
        if (dist) memcpy(dist + ((size_t)b * N + i) * k, tmp, sizeof(float) * (size_t)k);
      }
    }
    free(row);
    free(tmp);
  }
  free(sq);
  return 0;
}
