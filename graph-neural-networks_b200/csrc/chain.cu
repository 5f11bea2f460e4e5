// All K-1 shifts of one LSIGF call in ONE launch, for graphs small enough that a column slab of the signal lives in
// shared memory (BASELINE.json config 3: N = 1682; the 50..1000-node graphs of the reference's examples; every time step
// of the recurrent layers, graphML.py:1461) — SURVEY.md §8d "cfg3: single persistent kernel for all hops", VERDICT r1 #8.
//
//   z_0 = src,  z_k = A z_{k-1}  (k = 1 .. n_hops),   A = CSR gather operator of the plan, all z_k written to global memory
//   (the tap contraction reads them), but every hop READS its source from shared memory.
//
// A shift never mixes feature columns, so the matrix is cut into slabs of W columns (W * sizeof(T) = 64 bytes): one CTA
// owns one slab for the whole chain, keeps z_{k-1} and z_k of the slab in two shared-memory buffers (2 N W s bytes
// <= 227 KB: N <= ~1770 at 64-byte rows) and synchronises with __syncthreads between hops — no grid-wide barrier, no
// re-launch, no L2 round trip for the gathered rows (the separate hop kernels are L2-bandwidth-bound here: 24.6 us per
// hop at N = 1682, C = 2048).  1024 threads: warp per row, lanes = 8 neighbour groups x 4 lanes x 16 bytes, the row's
// col/val arrive with one coalesced load issued one row ahead (software prefetch), partial sums of the 8 groups are
// folded with shuffles.
#include <cstdlib>

#include "common.cuh"

namespace b200gf {
namespace chain {

constexpr int THREADS = 1024;
constexpr int MAX_HOPS = 15;

template <typename T>
struct Dsts { T* p[MAX_HOPS]; };

template <typename T, typename IDX>
__global__ void __launch_bounds__(THREADS, 1)
hop_chain_kernel(const IDX* __restrict__ rowptr, const int32_t* __restrict__ col, const T* __restrict__ val,
                 const T* __restrict__ src, int64_t src_ld, const __grid_constant__ Dsts<T> dsts, int64_t dst_ld, int n_rows,
                 int C, int n_hops) {
  constexpr int VEC = 16 / sizeof(T);      // elements per 16-byte lane vector
  constexpr int LPR = 4;                   // lanes per row: 64-byte slab rows
  constexpr int W = LPR * VEC;             // columns per slab
  constexpr int S = 32 / LPR;              // neighbours gathered concurrently by one warp
  using V16 = typename std::conditional<sizeof(T) == 4, float4, double2>::type;
  extern __shared__ __align__(16) unsigned char smem_chain[];
  V16* buf0 = reinterpret_cast<V16*>(smem_chain);
  V16* buf1 = buf0 + (size_t)n_rows * LPR;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int c0 = blockIdx.x * W;
  const int sub = lane / LPR, ch = lane % LPR;
  const bool col_ok = c0 + ch * VEC < C;   // C is a multiple of VEC (checked by the launcher)

  // slab of the source -> buf0
  for (int i = tid; i < n_rows * LPR; i += THREADS) {
    const int r = i / LPR, cc = i % LPR;
    V16 v;
    if (c0 + cc * VEC < C) v = *reinterpret_cast<const V16*>(src + (int64_t)r * src_ld + c0 + cc * VEC);
    else if constexpr (sizeof(T) == 4) v = make_float4(0.f, 0.f, 0.f, 0.f);
    else v = make_double2(0.0, 0.0);
    buf0[i] = v;
  }
  __syncthreads();

  constexpr int NW = THREADS / 32;
  for (int h = 0; h < n_hops; ++h) {
    const V16* __restrict__ in = (h & 1) ? buf1 : buf0;
    V16* __restrict__ out = (h & 1) ? buf0 : buf1;
    T* __restrict__ gdst = dsts.p[h];
    // prefetch of the first row's entries
    int row = warp;
    IDX beg = 0; int len = 0; int32_t c = 0; T v = T(0);
    if (row < n_rows) {
      beg = rowptr[row]; len = (int)(rowptr[row + 1] - beg);
      if (lane < len) { c = col[beg + lane]; v = val[beg + lane]; }
    }
    while (row < n_rows) {
      // next row's window (issued before this row's gathers: hides the rowptr -> col/val latency)
      const int nrow = row + NW;
      IDX nbeg = 0; int nlen = 0; int32_t nc = 0; T nv = T(0);
      if (nrow < n_rows) {
        nbeg = rowptr[nrow]; nlen = (int)(rowptr[nrow + 1] - nbeg);
        if (lane < nlen) { nc = col[nbeg + lane]; nv = val[nbeg + lane]; }
      }
      T acc[VEC];
#pragma unroll
      for (int i = 0; i < VEC; ++i) acc[i] = T(0);
      for (int b0 = 0; b0 < len; b0 += 32) {
        if (b0 > 0) {                        // rows longer than 32 entries: fetch the following window
          c = 0; v = T(0);
          if (b0 + lane < len) { c = col[beg + b0 + lane]; v = val[beg + b0 + lane]; }
        }
        const int cnt = min(len - b0, 32);
        for (int j = 0; j < cnt; j += S) {
          const int jj = j + sub;
          const int32_t cc = __shfl_sync(0xffffffffu, c, jj & 31);
          const T ww = __shfl_sync(0xffffffffu, v, jj & 31);
          if (jj < cnt) {
            const V16 d = in[(size_t)cc * LPR + ch];
            if constexpr (sizeof(T) == 4) {
              acc[0] = fma(ww, d.x, acc[0]); acc[1] = fma(ww, d.y, acc[1]); acc[2] = fma(ww, d.z, acc[2]); acc[3] = fma(ww, d.w, acc[3]);
            } else {
              acc[0] = fma(ww, d.x, acc[0]); acc[1] = fma(ww, d.y, acc[1]);
            }
          }
        }
      }
#pragma unroll
      for (int off = LPR; off < 32; off <<= 1)
#pragma unroll
        for (int i = 0; i < VEC; ++i) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], off);
      if (sub == 0) {
        V16 o;
        if constexpr (sizeof(T) == 4) o = make_float4(acc[0], acc[1], acc[2], acc[3]);
        else o = make_double2(acc[0], acc[1]);
        out[(size_t)row * LPR + ch] = o;
        if (col_ok) *reinterpret_cast<V16*>(gdst + (int64_t)row * dst_ld + c0 + ch * VEC) = o;
      }
      row = nrow; beg = nbeg; len = nlen; c = nc; v = nv;
    }
    __syncthreads();
  }
}

template <typename T>
static int launch_t(const CsrDev& A, int64_t n_rows, const T* src, int64_t src_ld, T* const* dst, int64_t dst_ld, int C,
                    int n_hops, cudaStream_t st) {
  constexpr int VEC = 16 / sizeof(T), W = 4 * VEC;
  Dsts<T> d{};
  for (int h = 0; h < n_hops; ++h) d.p[h] = dst[h];
  const size_t smem = (size_t)2 * n_rows * W * sizeof(T);
  const int slabs = (C + W - 1) / W;
  if (A.rowptr32) {
    auto kern = hop_chain_kernel<T, int32_t>;
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    kern<<<slabs, THREADS, smem, st>>>(A.rowptr32, A.col, reinterpret_cast<const T*>(A.val), src, src_ld, d, dst_ld, (int)n_rows, C, n_hops);
  } else {
    auto kern = hop_chain_kernel<T, int64_t>;
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    kern<<<slabs, THREADS, smem, st>>>(A.rowptr, A.col, reinterpret_cast<const T*>(A.val), src, src_ld, d, dst_ld, (int)n_rows, C, n_hops);
  }
  LAUNCH_CHECK();
  return B200GF_OK;
}

}  // namespace chain

static size_t chain_max_bytes() {
  static const size_t v = [] {
    const char* e = getenv("B200GF_CHAIN_MAX_BYTES");   // 0 disables the chain kernel (A/B timing in tests)
    return e ? (size_t)atoll(e) : (size_t)2 << 20;
  }();
  return v;
}

// square operators only (the chain feeds its own output back); every operand 16-byte aligned with C a whole number of
// 16-byte vectors; the two slab buffers must fit the 227 KB of shared memory
bool hop_chain_eligible(int dtype, const CsrDev& A, int64_t n_rows, int64_t n_cols, const void* src, int64_t src_ld,
                        void* const* dst, int64_t dst_ld, int C, int n_hops) {
  if (n_rows != n_cols || n_rows <= 0 || n_hops < 2 || n_hops > chain::MAX_HOPS) return false;   // one hop: the plain kernel
  const size_t es = dtype_size(dtype);
  // Used only where it wins: launch-latency-bound chains.  At BASELINE config 3 (N = 1682, C = 2048: 14 MB per hop source)
  // the separate hop kernels are L2-bandwidth-bound at 24.6 us per hop and this kernel, with one 1024-thread CTA per SM
  // and its index loads missing the (shared-memory-sized-down) L1, took 50 us per hop (profiles/r2_bench_er1m_1gpu_b.json).
  if ((size_t)n_rows * (size_t)C * es > chain_max_bytes()) return false;
  const int VEC = (int)(16 / es);
  if ((size_t)2 * n_rows * 64 > (size_t)227 * 1024 - 1024) return false;
  if (C % VEC != 0 || src_ld % VEC != 0 || dst_ld % VEC != 0 || (reinterpret_cast<uintptr_t>(src) & 15)) return false;
  for (int h = 0; h < n_hops; ++h)
    if (!dst[h] || (reinterpret_cast<uintptr_t>(dst[h]) & 15)) return false;
  return true;
}

int launch_hop_chain(int dtype, const CsrDev& A, int64_t n_rows, const void* src, int64_t src_ld, void* const* dst,
                     int64_t dst_ld, int C, int n_hops, cudaStream_t st) {
  if (dtype == B200GF_F32)
    return chain::launch_t<float>(A, n_rows, (const float*)src, src_ld, reinterpret_cast<float* const*>(dst), dst_ld, C, n_hops, st);
  if (dtype == B200GF_F64)
    return chain::launch_t<double>(A, n_rows, (const double*)src, src_ld, reinterpret_cast<double* const*>(dst), dst_ld, C, n_hops, st);
  return B200GF_EUNSUPPORTED;
}

}  // namespace b200gf

extern "C" int b200gf_hop_chain(const b200gf_plan* plan, int e, int direction, const void* src, int64_t src_ld,
                                void* const* dst, int64_t dst_ld, int C, int n_hops, void* stream) {
  using namespace b200gf;
  if (!plan || !src || !dst || e < 0 || e >= plan->E || C <= 0 || n_hops < 1) return B200GF_EINVAL;
  if (direction != B200GF_HOP_FWD && direction != B200GF_HOP_BWD) return B200GF_EINVAL;
  if (direction == B200GF_HOP_BWD && !plan->has_bwd) return B200GF_EINVAL;
  const CsrDev& A = direction == B200GF_HOP_FWD ? plan->fwd[e] : plan->bwd[e];
  if (!hop_chain_eligible(plan->dtype, A, plan->n_rows, plan->n_cols, src, src_ld, dst, dst_ld, C, n_hops))
    return B200GF_EUNSUPPORTED;
  return launch_hop_chain(plan->dtype, A, plan->n_rows, src, src_ld, dst, dst_ld, C, n_hops, (cudaStream_t)stream);
}
