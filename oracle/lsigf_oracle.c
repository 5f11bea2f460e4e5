/*
 * Plain-C restatement of the LSIGF path on the CPU (float64).  TEST INFRASTRUCTURE — NOT PRODUCT CODE.
 * Only tests/, __graft_entry__ (build + smoke) and bench.py's CPU legs may build, load or call this file.
 *
 * Follows alegnn/utils/graphML.py:83-176 step by step, in the reference's own [B, G, N] layout, with every S_e held in
 * CSR (row i lists the non-zeros S_e[i, j]) so that it also runs where the dense E x N x N GSO cannot exist:
 *     z_{e,0} = x                                   graphML.py:152-154
 *     z_{e,k} = z_{e,k-1} S_e   (ROW-vector shift)  graphML.py:158-161 :  (zS)[c, j] = sum_i z[c, i] S[i, j]
 *     y[b,f,n] = sum_{e,k,g} h[f,e,k,g] z_{e,k}[b,g,n] + bias     graphML.py:170-175
 * and the gradients autograd derives from it.  A second, independent checker beside the numpy oracle
 * (oracle/lsigf_oracle.py); both are pinned to fixtures produced by the unmodified reference (tests/golden).
 *
 * Build: gcc -O2 -shared -fPIC -o oracle/_build/liblsigf_oracle.so oracle/lsigf_oracle.c   (oracle/c_oracle.py does it)
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* out[c, j] = sum_i in[c, i] * S[i, j]   for c < C  (in, out: [C, N] row-major) */
static void shift_rows(int64_t N, int64_t C, const int64_t* rowptr, const int32_t* col, const double* val,
                       const double* in, double* out) {
  memset(out, 0, (size_t)(C * N) * sizeof(double));
  for (int64_t i = 0; i < N; ++i) {
    for (int64_t p = rowptr[i]; p < rowptr[i + 1]; ++p) {
      const int64_t j = col[p];
      const double v = val[p];
      for (int64_t c = 0; c < C; ++c) out[c * N + j] += in[c * N + i] * v;
    }
  }
}

/* out[c, i] = sum_j S[i, j] * in[c, j]   (shift with S^T in the row-vector convention) */
static void shift_rows_transposed(int64_t N, int64_t C, const int64_t* rowptr, const int32_t* col, const double* val,
                                  const double* in, double* out) {
  for (int64_t c = 0; c < C; ++c) {
    for (int64_t i = 0; i < N; ++i) {
      double acc = 0.0;
      for (int64_t p = rowptr[i]; p < rowptr[i + 1]; ++p) acc += val[p] * in[c * N + col[p]];
      out[c * N + i] = acc;
    }
  }
}

/* y [B,F,N]; h [F,E,K,G]; x [B,G,N]; bias NULL / [F] (bias_per_node = 0) / [F,N] (bias_per_node = 1) */
int lsigf_oracle_forward(int64_t N, int E, int K, int G, int F, int B, const int64_t* const* rowptr,
                         const int32_t* const* col, const double* const* val, const double* h, const double* x,
                         const double* bias, int bias_per_node, double* y) {
  const int64_t C = (int64_t)B * G;
  double* z = (double*)malloc((size_t)(C * N) * sizeof(double));
  double* zn = (double*)malloc((size_t)(C * N) * sizeof(double));
  if (!z || !zn) { free(z); free(zn); return -1; }
  for (int64_t b = 0; b < B; ++b)
    for (int f = 0; f < F; ++f)
      for (int64_t n = 0; n < N; ++n)
        y[(b * F + f) * N + n] = bias ? (bias_per_node ? bias[(int64_t)f * N + n] : bias[f]) : 0.0;
  for (int e = 0; e < E; ++e) {
    memcpy(z, x, (size_t)(C * N) * sizeof(double));             /* k = 0: the same x for every e (:154) */
    for (int k = 0; k < K; ++k) {
      if (k > 0) {
        shift_rows(N, C, rowptr[e], col[e], val[e], z, zn);
        double* t = z; z = zn; zn = t;
      }
      for (int64_t b = 0; b < B; ++b)
        for (int f = 0; f < F; ++f)
          for (int g = 0; g < G; ++g) {
            const double w = h[(((int64_t)f * E + e) * K + k) * G + g];
            const double* zr = z + (b * G + g) * N;
            double* yr = y + (b * F + f) * N;
            for (int64_t n = 0; n < N; ++n) yr[n] += w * zr[n];
          }
    }
  }
  free(z); free(zn);
  return 0;
}

/* dy [B,F,N] -> dh [F,E,K,G], dx [B,G,N], db ([F] or [F,N]; may be NULL) */
int lsigf_oracle_backward(int64_t N, int E, int K, int G, int F, int B, const int64_t* const* rowptr,
                          const int32_t* const* col, const double* const* val, const double* h, const double* x,
                          const double* dy, int bias_per_node, double* dh, double* dx, double* db) {
  const int64_t C = (int64_t)B * G, CF = (int64_t)B * F;
  double* z = (double*)malloc((size_t)(C * N) * sizeof(double));
  double* zn = (double*)malloc((size_t)(C * N) * sizeof(double));
  double* v = (double*)malloc((size_t)(CF * N) * sizeof(double));
  double* vn = (double*)malloc((size_t)(CF * N) * sizeof(double));
  if (!z || !zn || !v || !vn) { free(z); free(zn); free(v); free(vn); return -1; }
  memset(dx, 0, (size_t)(C * N) * sizeof(double));
  for (int e = 0; e < E; ++e) {
    memcpy(z, x, (size_t)(C * N) * sizeof(double));
    memcpy(v, dy, (size_t)(CF * N) * sizeof(double));
    for (int k = 0; k < K; ++k) {
      if (k > 0) {
        shift_rows(N, C, rowptr[e], col[e], val[e], z, zn);                 /* z_k = z_{k-1} S_e        */
        shift_rows_transposed(N, CF, rowptr[e], col[e], val[e], v, vn);     /* v_k = v_{k-1} S_e^T      */
        double* t = z; z = zn; zn = t;
        t = v; v = vn; vn = t;
      }
      for (int f = 0; f < F; ++f)
        for (int g = 0; g < G; ++g) {
          double acc = 0.0;                                                  /* dh = <dy_f, z_{e,k,g}>   */
          const double w = h[(((int64_t)f * E + e) * K + k) * G + g];
          for (int64_t b = 0; b < B; ++b) {
            const double* zr = z + (b * G + g) * N;
            const double* dyr = dy + (b * F + f) * N;
            const double* vr = v + (b * F + f) * N;
            double* dxr = dx + (b * G + g) * N;
            for (int64_t n = 0; n < N; ++n) {
              acc += dyr[n] * zr[n];
              dxr[n] += w * vr[n];                                           /* dx += h * (dy S^T^k)     */
            }
          }
          dh[(((int64_t)f * E + e) * K + k) * G + g] = acc;
        }
    }
  }
  if (db) {
    if (bias_per_node) {
      for (int f = 0; f < F; ++f)
        for (int64_t n = 0; n < N; ++n) {
          double s = 0.0;
          for (int64_t b = 0; b < B; ++b) s += dy[(b * F + f) * N + n];
          db[(int64_t)f * N + n] = s;
        }
    } else {
      for (int f = 0; f < F; ++f) {
        double s = 0.0;
        for (int64_t b = 0; b < B; ++b)
          for (int64_t n = 0; n < N; ++n) s += dy[(b * F + f) * N + n];
        db[f] = s;
      }
    }
  }
  free(z); free(zn); free(v); free(vn);
  return 0;
}
