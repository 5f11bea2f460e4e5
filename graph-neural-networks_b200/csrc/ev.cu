// Edge-variant graph filter (EVGF), sparse execution — the variant row of the path.
//
// Reference: EVGF(S, x, b), alegnn/utils/graphML.py:389-488, called by EdgeVariantGF.forward (:2670-2698) with
// Phi = weightEV * sparsityPatternFull (:2676).  Per (f, e, g) it runs the chain
//     u_0 = Phi^(0) x_g ,  u_k = Phi^(k) u_{k-1}      (COLUMN convention Phi·x, graphML.py:464,475)
// and sums u_k over k, g, e.  The reference stores every Phi^(k)_{feg} as a dense N x N matrix; only the entries on
// the masked sparsity pattern of |S_e| + I are live (:2620-2663).  Here the live entries are stored per non-zero of
// one CSR pattern (rowptr, col) over a COMPACT node set A of NA nodes (for the hybrid layer: the M selected nodes and
// the nodes adjacent to them — every other row and column of Phi is identically zero), one call per edge feature e:
//     w    [F, K, G, nnz]   Phi^(k) on the pattern (k = 0 included: its off-diagonal entries are simply zero)
//     xA   [B, G, NA]       input restricted to A
//     S    [F, G, B, NA]    sum_k u_k   (the caller sums over g and e and adds the bias)
// chains c = (f*G + g)*B + b ; states u_k [c, NA] are kept for the backward pass.
#include "common.cuh"

namespace b200gf {
namespace ev {

// u_k[c, i] = sum_{idx in row i} w_k[f, g, idx] * prev[col[idx]] ;  S (+)= u_k.   k = 0 reads x instead of u_{-1}
template <typename T>
__global__ void hop_kernel(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col, const T* __restrict__ w,
                           const T* __restrict__ uprev, const T* __restrict__ xA, T* __restrict__ ucur, T* __restrict__ S,
                           int64_t NA, int B, int G, int K, int k, int64_t nnz, int64_t total) {
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t c = t / NA;
    const int i = (int)(t - c * NA);
    const int b = (int)(c % B);
    const int64_t fg = c / B;
    const int64_t f = fg / G;
    const int g = (int)(fg - f * G);
    const T* __restrict__ wk = w + ((f * K + k) * G + g) * nnz;
    const T* __restrict__ up = k == 0 ? xA + ((int64_t)b * G + g) * NA : uprev + c * NA;
    T acc = T(0);
    for (int64_t idx = rowptr[i]; idx < rowptr[i + 1]; ++idx) acc = fma(wk[idx], up[col[idx]], acc);
    ucur[t] = acc;
    S[t] = k == 0 ? acc : S[t] + acc;
  }
}

// lam_{K-1}[c, i] = dyA[b, f, i]
template <typename T>
__global__ void adjoint_init_kernel(const T* __restrict__ dyA, T* __restrict__ lam, int64_t NA, int B, int G, int F,
                                    int64_t total) {
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t c = t / NA;
    const int i = (int)(t - c * NA);
    const int b = (int)(c % B);
    const int64_t f = (c / B) / G;
    lam[t] = dyA[((int64_t)b * F + f) * NA + i];
  }
}

// lam_{k-1}[c, j] = dyA[b, f, j] + sum_{it in rowT(j)} w_k[perm[it]] * lam_k[c, colT[it]]        (k >= 1)
template <typename T>
__global__ void adjoint_hop_kernel(const int64_t* __restrict__ rowptrT, const int32_t* __restrict__ colT,
                                   const int64_t* __restrict__ perm, const T* __restrict__ w, const T* __restrict__ dyA,
                                   const T* __restrict__ lam_k, T* __restrict__ lam_km1, int64_t NA, int B, int G, int F,
                                   int K, int k, int64_t nnz, int64_t total) {
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t c = t / NA;
    const int j = (int)(t - c * NA);
    const int b = (int)(c % B);
    const int64_t fg = c / B;
    const int64_t f = fg / G;
    const int g = (int)(fg - f * G);
    const T* __restrict__ wk = w + ((f * K + k) * G + g) * nnz;
    const T* __restrict__ lk = lam_k + c * NA;
    T acc = dyA[((int64_t)b * F + f) * NA + j];
    for (int64_t it = rowptrT[j]; it < rowptrT[j + 1]; ++it) acc = fma(wk[perm[it]], lk[colT[it]], acc);
    lam_km1[t] = acc;
  }
}

// dw_k[f, g, idx(i, j)] = sum_b lam_k[(fg, b), i] * prev[(fg, b), j]      prev = u_{k-1}, or x for k = 0
template <typename T>
__global__ void wgrad_kernel(const int32_t* __restrict__ rowidx, const int32_t* __restrict__ col, const T* __restrict__ lam_k,
                             const T* __restrict__ uprev, const T* __restrict__ xA, T* __restrict__ dw, int64_t NA, int B,
                             int G, int K, int k, int64_t nnz, int64_t total /* F*G*nnz */) {
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t fg = t / nnz;
    const int64_t idx = t - fg * nnz;
    const int64_t f = fg / G;
    const int g = (int)(fg - f * G);
    const int i = rowidx[idx], j = col[idx];
    T acc = T(0);
    for (int b = 0; b < B; ++b) {
      const int64_t c = fg * B + b;
      const T p = k == 0 ? xA[((int64_t)b * G + g) * NA + j] : uprev[c * NA + j];
      acc = fma(lam_k[c * NA + i], p, acc);
    }
    dw[((f * K + k) * G + g) * nnz + idx] = acc;
  }
}

// dxA[b, g, j] = sum_f sum_{it in rowT(j)} w_0[f, g, perm[it]] * lam_0[((f*G + g)*B + b), colT[it]]
template <typename T>
__global__ void xgrad_kernel(const int64_t* __restrict__ rowptrT, const int32_t* __restrict__ colT,
                             const int64_t* __restrict__ perm, const T* __restrict__ w, const T* __restrict__ lam0,
                             T* __restrict__ dxA, int64_t NA, int B, int G, int F, int K, int64_t nnz,
                             int64_t total /* B*G*NA */) {
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t bg = t / NA;
    const int j = (int)(t - bg * NA);
    const int b = (int)(bg / G);
    const int g = (int)(bg - (int64_t)b * G);
    T acc = T(0);
    for (int f = 0; f < F; ++f) {
      const T* __restrict__ w0 = w + (((int64_t)f * K + 0) * G + g) * nnz;
      const T* __restrict__ l0 = lam0 + (((int64_t)f * G + g) * B + b) * NA;
      for (int64_t it = rowptrT[j]; it < rowptrT[j + 1]; ++it) acc = fma(w0[perm[it]], l0[colT[it]], acc);
    }
    dxA[t] = acc;
  }
}

inline int grid_for(int64_t total) { return (int)imin64((total + 255) / 256, 148 * 16); }

template <typename T>
int forward_t(int64_t NA, int B, int G, int F, int K, const int64_t* rowptr, const int32_t* col, int64_t nnz, const T* w,
              const T* xA, T* states, T* S, cudaStream_t st) {
  const int64_t total = (int64_t)F * G * B * NA;
  if (total == 0) return B200GF_OK;
  for (int k = 0; k < K; ++k) {
    hop_kernel<T><<<grid_for(total), 256, 0, st>>>(rowptr, col, w, k > 0 ? states + (int64_t)(k - 1) * total : nullptr, xA,
                                                   states + (int64_t)k * total, S, NA, B, G, K, k, nnz, total);
    LAUNCH_CHECK();
  }
  return B200GF_OK;
}

template <typename T>
int backward_t(int64_t NA, int B, int G, int F, int K, const int32_t* rowidx, const int32_t* col, const int64_t* rowptrT,
               const int32_t* colT, const int64_t* perm, int64_t nnz, const T* w, const T* xA, const T* states,
               const T* dyA, T* lam /* 2 x [chains, NA] */, T* dw, T* dxA, cudaStream_t st) {
  const int64_t total = (int64_t)F * G * B * NA;
  if (total == 0) return B200GF_OK;
  T* cur = lam;
  T* nxt = lam + total;
  adjoint_init_kernel<T><<<grid_for(total), 256, 0, st>>>(dyA, cur, NA, B, G, F, total);
  LAUNCH_CHECK();
  const int64_t tw = (int64_t)F * G * nnz;
  for (int k = K - 1; k >= 0; --k) {
    if (tw > 0) {
      wgrad_kernel<T><<<grid_for(tw), 256, 0, st>>>(rowidx, col, cur, k > 0 ? states + (int64_t)(k - 1) * total : nullptr,
                                                    xA, dw, NA, B, G, K, k, nnz, tw);
      LAUNCH_CHECK();
    }
    if (k == 0) break;
    adjoint_hop_kernel<T><<<grid_for(total), 256, 0, st>>>(rowptrT, colT, perm, w, dyA, cur, nxt, NA, B, G, F, K, k, nnz,
                                                           total);
    LAUNCH_CHECK();
    T* tmp = cur; cur = nxt; nxt = tmp;
  }
  const int64_t tx = (int64_t)B * G * NA;
  xgrad_kernel<T><<<grid_for(tx), 256, 0, st>>>(rowptrT, colT, perm, w, cur, dxA, NA, B, G, F, K, nnz, tx);
  LAUNCH_CHECK();
  return B200GF_OK;
}

}  // namespace ev
}  // namespace b200gf

extern "C" {

int b200gf_ev_forward(int dtype, int64_t NA, int B, int G, int F, int K, const int64_t* rowptr, const int32_t* col,
                      int64_t nnz, const void* w, const void* xA, void* states, void* S, void* stream) {
  using namespace b200gf;
  if (NA < 0 || B <= 0 || G <= 0 || F <= 0 || K <= 0 || nnz < 0) return B200GF_EINVAL;
  if (!rowptr || !col || !xA || !states || !S || (nnz > 0 && !w)) return B200GF_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == B200GF_F32)
    return ev::forward_t<float>(NA, B, G, F, K, rowptr, col, nnz, (const float*)w, (const float*)xA, (float*)states,
                                (float*)S, st);
  if (dtype == B200GF_F64)
    return ev::forward_t<double>(NA, B, G, F, K, rowptr, col, nnz, (const double*)w, (const double*)xA, (double*)states,
                                 (double*)S, st);
  return B200GF_EUNSUPPORTED;
}

int b200gf_ev_backward(int dtype, int64_t NA, int B, int G, int F, int K, const int32_t* rowidx, const int32_t* col,
                       const int64_t* rowptrT, const int32_t* colT, const int64_t* perm, int64_t nnz, const void* w,
                       const void* xA, const void* states, const void* dyA, void* lam, void* dw, void* dxA, void* stream) {
  using namespace b200gf;
  if (NA < 0 || B <= 0 || G <= 0 || F <= 0 || K <= 0 || nnz < 0) return B200GF_EINVAL;
  if (!rowidx || !col || !rowptrT || !colT || !perm || !xA || !states || !dyA || !lam || !dxA) return B200GF_EINVAL;
  if (nnz > 0 && (!w || !dw)) return B200GF_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == B200GF_F32)
    return ev::backward_t<float>(NA, B, G, F, K, rowidx, col, rowptrT, colT, perm, nnz, (const float*)w, (const float*)xA,
                                 (const float*)states, (const float*)dyA, (float*)lam, (float*)dw, (float*)dxA, st);
  if (dtype == B200GF_F64)
    return ev::backward_t<double>(NA, B, G, F, K, rowidx, col, rowptrT, colT, perm, nnz, (const double*)w,
                                  (const double*)xA, (const double*)states, (const double*)dyA, (double*)lam, (double*)dw,
                                  (double*)dxA, st);
  return B200GF_EUNSUPPORTED;
}

}  // extern "C"
