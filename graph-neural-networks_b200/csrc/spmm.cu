// One graph shift ("hop") of the LSIGF path:  dst[r, :] = sum_j A[r, j] * src[j, :]
// with A a CSR gather operator (S_e^T for the forward x·S of reference graphML.py:159, S_e for backward)
// and src/dst node-major [rows, ld] feature matrices.  This replaces the reference's dense broadcast-batched
// GEMM torch.matmul(x, S) and is the HBM-bound kernel the roofline in bench.py is quoted on.
//
// Mapping (sm_100a, 148 SMs):
//   * one warp per (row, 128-byte..512-byte column chunk) work item, grid-stride over items in
//     chunk-major order, so that at any time all resident warps gather from the same column slab
//     (keeps the gathered slab L2-resident when N * chunk_bytes fits the 126 MB L2);
//   * L lanes x 16-byte vectors cover the chunk; the 32/L lane groups each take a different neighbour,
//     so one warp-wide LDG.128 fetches 32/L whole neighbour rows (fully used 32-byte sectors);
//   * U independent LDG.128 per lane are issued before the FMAs (memory-level parallelism);
//   * col/val of a row are read once, coalesced (lane i holds entry i) and broadcast with SHFL;
//   * next row's col/val and the row after's rowptr are prefetched while the current row is gathered.
#include "common.cuh"
#include "spmm_kernels.cuh"

namespace b200gf {

// Library configuration, chosen from tools/spmm_sweep.cu on B200 (profiles/r1_spmm_sweep.md): 256 threads,
// registers capped for 6 resident blocks/SM (48 warps: the kernel is latency-bound below that), gathered rows loaded
// with an L2 evict_last policy and no L1 allocation (DRAM reads 7.9 GB -> 6.3 GB per hop at N=1M, C=64), no
// next-row prefetch (it costs registers and measured slower).
template <typename T, int VEC, int L, int U>
static int launch_one(int sm_count, const CsrDev& A, int64_t n_rows, const T* src, int64_t src_ld, T* dst,
                      int64_t dst_ld, int C, cudaStream_t st) {
  constexpr int THREADS = 256, MINB = sizeof(T) == 4 ? 6 : 4, HINT = 3;
  constexpr bool PF = false;
  auto kern = spmm_hop_kernel<T, VEC, L, U, THREADS, MINB, HINT, PF>;
  const int n_chunks = (C + L * VEC - 1) / (L * VEC);
  const int64_t n_items = n_rows * n_chunks;
  if (n_items == 0) return B200GF_OK;
  const int wpb = THREADS / 32;
  int occ = 0;
  CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, THREADS, 0));
  if (occ < 1) occ = 1;
  int64_t blocks = (n_items + wpb - 1) / wpb;
  const int64_t cap = (int64_t)sm_count * occ;  // one resident wave: persistent warps stride over the items
  if (blocks > cap) blocks = cap;
  kern<<<(unsigned)blocks, THREADS, 0, st>>>(A.rowptr, A.col, reinterpret_cast<const T*>(A.val), src, src_ld, dst,
                                             dst_ld, n_rows, C, n_chunks, 1.0f);
  LAUNCH_CHECK();
  return B200GF_OK;
}

// narrow rows (<= 128 bytes): several rows per warp (spmm_hop_multirow_kernel); geometry from the small-C sweep
// (profiles/r1_spmm_sweep_smallC.log: C = 8 0.39 -> 0.19 ms, C = 16 0.48 -> 0.34 ms, C = 32 0.75 -> 0.68 ms at N = 1M)
template <typename T, int VEC, int L, int GS, int U, int MINB>
static int launch_multirow(int sm_count, const CsrDev& A, int64_t n_rows, const T* src, int64_t src_ld, T* dst,
                           int64_t dst_ld, int C, cudaStream_t st) {
  constexpr int THREADS = 256, HINT = 3;
  auto kern = spmm_hop_multirow_kernel<T, VEC, L, GS, U, THREADS, MINB, HINT>;
  if (n_rows == 0) return B200GF_OK;
  constexpr int rows_per_block = (THREADS / 32) * (32 / GS);
  int occ = 0;
  CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, THREADS, 0));
  if (occ < 1) occ = 1;
  int64_t blocks = (n_rows + rows_per_block - 1) / rows_per_block;
  const int64_t cap = (int64_t)sm_count * occ;
  if (blocks > cap) blocks = cap;
  kern<<<(unsigned)blocks, THREADS, 0, st>>>(A.rowptr, A.col, reinterpret_cast<const T*>(A.val), src, src_ld, dst,
                                             dst_ld, n_rows, C);
  LAUNCH_CHECK();
  return B200GF_OK;
}

template <typename T>
static int launch_typed(int sm_count, const CsrDev& A, int64_t n_rows, const void* src_, int64_t src_ld,
                        void* dst_, int64_t dst_ld, int C, cudaStream_t st) {
  constexpr int VEC = 16 / sizeof(T);
  const T* src = reinterpret_cast<const T*>(src_);
  T* dst = reinterpret_cast<T*>(dst_);
  const int Cv = (C + VEC - 1) / VEC * VEC;
  const bool vec_ok = (src_ld % VEC == 0) && (dst_ld % VEC == 0) && (Cv <= src_ld) && (Cv <= dst_ld) &&
                      ((reinterpret_cast<uintptr_t>(src) & 15) == 0) && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0);
  if (!vec_ok) return launch_one<T, 1, 32, 4>(sm_count, A, n_rows, src, src_ld, dst, dst_ld, C, st);
  const int nv = Cv / VEC;  // 16-byte vectors per row
  constexpr int MB = sizeof(T) == 4 ? 6 : 4;
  if (nv <= 1) return launch_multirow<T, VEC, 1, 8, 1, MB>(sm_count, A, n_rows, src, src_ld, dst, dst_ld, C, st);
  if (nv <= 2) return launch_multirow<T, VEC, 2, 8, 2, MB>(sm_count, A, n_rows, src, src_ld, dst, dst_ld, C, st);
  if (nv <= 4) return launch_multirow<T, VEC, 4, 16, 2, sizeof(T) == 4 ? 8 : 4>(sm_count, A, n_rows, src, src_ld, dst, dst_ld, C, st);
  if (nv <= 8) return launch_multirow<T, VEC, 8, 32, 4, MB>(sm_count, A, n_rows, src, src_ld, dst, dst_ld, C, st);
  if (nv <= 16) return launch_one<T, VEC, 16, 4>(sm_count, A, n_rows, src, src_ld, dst, dst_ld, C, st);
  return launch_one<T, VEC, 32, 4>(sm_count, A, n_rows, src, src_ld, dst, dst_ld, C, st);
}

int launch_hop(int dtype, int sm_count, const CsrDev& A, int64_t n_rows, const void* src, int64_t src_ld,
               void* dst, int64_t dst_ld, int C, cudaStream_t st) {
  if (C <= 0 || src_ld < C || dst_ld < C) return B200GF_EINVAL;
  if (dtype == B200GF_F32) return launch_typed<float>(sm_count, A, n_rows, src, src_ld, dst, dst_ld, C, st);
  if (dtype == B200GF_F64) return launch_typed<double>(sm_count, A, n_rows, src, src_ld, dst, dst_ld, C, st);
  return B200GF_EUNSUPPORTED;
}

}  // namespace b200gf

extern "C" int b200gf_hop(const b200gf_plan* plan, int e, int direction, const void* src, int64_t src_ld,
                          void* dst, int64_t dst_ld, int C, void* stream) {
  if (!plan || !src || !dst || e < 0 || e >= plan->E) return B200GF_EINVAL;
  if (direction != B200GF_HOP_FWD && direction != B200GF_HOP_BWD) return B200GF_EINVAL;
  if (direction == B200GF_HOP_BWD && !plan->has_bwd) return B200GF_EINVAL;
  const b200gf::CsrDev& A = direction == B200GF_HOP_FWD ? plan->fwd[e] : plan->bwd[e];
  return b200gf::launch_hop(plan->dtype, plan->sm_count, A, plan->n_rows, src, src_ld, dst, dst_ld, C,
                            (cudaStream_t)stream);
}
