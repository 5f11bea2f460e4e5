// Kernels of the graph shift ("hop"); included by spmm.cu (the library) and tools/spmm_sweep.cu (tuning sweeps).
//
//   dst[r, :] = sum_j A[r, j] * src[j, :]        A = CSR gather operator, src/dst node-major [rows, ld]
//
// Mapping (sm_100a, 148 SMs):
//   * one warp per (row, column chunk) work item, grid-stride over items in chunk-major order, so that at any
//     time all resident warps gather from the same column slab (keeps it L2-resident when N * chunk_bytes fits);
//   * L lanes x 16-byte vectors cover the chunk; the 32/L lane groups each take a different neighbour, so one
//     warp-wide LDG.128 fetches 32/L whole neighbour rows (every 32-byte sector fully used);
//   * U independent LDG.128 per lane are in flight before the FMAs (memory-level parallelism);
//   * col/val of a row are read once, coalesced (lane i holds entry i), and broadcast with SHFL;
//   * PF (template flag, off in the library): prefetch of the next row's col/val and of the rowptr pair after it while
//     the current row is gathered — measured slower (registers / issue slots), kept for the sweep tool.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <type_traits>

namespace b200gf {

template <typename T, int VEC>
struct Acc {
  T v[VEC];
  __device__ __forceinline__ void zero() {
#pragma unroll
    for (int i = 0; i < VEC; ++i) v[i] = T(0);
  }
};

// L2 cache policy: `frac` of the touched lines (chosen by address hash) get evict_last priority, the rest stay
// evict_unchanged.  With the gathered feature matrix larger than the L2, a sticky subset that fits turns an LRU
// thrash (measured 6 % L2 hit rate) into hits on that subset.
__device__ __forceinline__ uint64_t evict_last_policy(float frac) {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, %1;" : "=l"(pol) : "f"(frac));
  return pol;
}

// HINT 0: ld.global.nc (default policy)   1: + L1::no_allocate (gathered rows have no L1 reuse)
// HINT 2: L1::no_allocate + L2::evict_last on the gathered rows (fight for L2 residency of the feature slab)
template <typename T, int VEC, int HINT>
__device__ __forceinline__ Acc<T, VEC> load_vec(const T* p, uint64_t pol) {
  Acc<T, VEC> a;
  if constexpr (VEC == 4 && sizeof(T) == 4) {
    float4 t;
    if constexpr (HINT == 0) {
      t = __ldg(reinterpret_cast<const float4*>(p));
    } else if constexpr (HINT == 1) {
      asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                   : "=f"(t.x), "=f"(t.y), "=f"(t.z), "=f"(t.w) : "l"(p));
    } else {
      asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
                   : "=f"(t.x), "=f"(t.y), "=f"(t.z), "=f"(t.w) : "l"(p), "l"(pol));
    }
    a.v[0] = t.x; a.v[1] = t.y; a.v[2] = t.z; a.v[3] = t.w;
  } else if constexpr (VEC * sizeof(T) == 32) {
    // sm_100 256-bit load (LDG.E.256): one lane fetches a whole 32-byte sector.  HINT 0/1: L1::no_allocate only;
    // HINT 3: L2::evict_last as a plain qualifier (no policy register); other HINTs: the policy in `pol`.
    if constexpr (sizeof(T) == 4) {
      unsigned r[8];
      if constexpr (HINT == 3)
        asm volatile("ld.global.nc.L1::no_allocate.L2::evict_last.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "l"(p));
      else if constexpr (HINT >= 2)
        asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8], %9;"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "l"(p), "l"(pol));
      else
        asm volatile("ld.global.nc.L1::no_allocate.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "l"(p));
#pragma unroll
      for (int i = 0; i < 8; ++i) a.v[i] = __uint_as_float(r[i]);
    } else {
      double d[4];
      if constexpr (HINT == 3)
        asm volatile("ld.global.nc.L1::no_allocate.L2::evict_last.v4.b64 {%0,%1,%2,%3}, [%4];"
                     : "=d"(d[0]), "=d"(d[1]), "=d"(d[2]), "=d"(d[3]) : "l"(p));
      else if constexpr (HINT >= 2)
        asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.b64 {%0,%1,%2,%3}, [%4], %5;"
                     : "=d"(d[0]), "=d"(d[1]), "=d"(d[2]), "=d"(d[3]) : "l"(p), "l"(pol));
      else
        asm volatile("ld.global.nc.L1::no_allocate.v4.b64 {%0,%1,%2,%3}, [%4];"
                     : "=d"(d[0]), "=d"(d[1]), "=d"(d[2]), "=d"(d[3]) : "l"(p));
#pragma unroll
      for (int i = 0; i < 4; ++i) a.v[i] = d[i];
    }
  } else if constexpr (VEC == 2) {
    double2 t;
    if constexpr (HINT == 0) {
      t = __ldg(reinterpret_cast<const double2*>(p));
    } else if constexpr (HINT == 1) {
      asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(t.x), "=d"(t.y) : "l"(p));
    } else {
      asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.f64 {%0,%1}, [%2], %3;"
                   : "=d"(t.x), "=d"(t.y) : "l"(p), "l"(pol));
    }
    a.v[0] = t.x; a.v[1] = t.y;
  } else {
    a.v[0] = __ldg(p);
  }
  return a;
}

// SH 0: default store   1: st.global.cs (streaming: the result row is not re-read by this kernel)
template <typename T, int VEC, int SH>
__device__ __forceinline__ void store_vec(T* p, const Acc<T, VEC>& a) {
  if constexpr (VEC * sizeof(T) == 32 && sizeof(T) == 4) {
    // one 256-bit store per lane (STG.E.256): a row leaves as whole contiguous sectors — two 16-byte stores per lane
    // would interleave half-sector writes across the warp, which costs on NVLink peer stores (fused all-gather epilogue)
    if constexpr (SH == 1)
      asm volatile("st.global.cs.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(a.v[0]), "f"(a.v[1]), "f"(a.v[2]),
                   "f"(a.v[3]), "f"(a.v[4]), "f"(a.v[5]), "f"(a.v[6]), "f"(a.v[7]) : "memory");
    else
      asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(a.v[0]), "f"(a.v[1]), "f"(a.v[2]),
                   "f"(a.v[3]), "f"(a.v[4]), "f"(a.v[5]), "f"(a.v[6]), "f"(a.v[7]) : "memory");
  } else if constexpr (VEC * sizeof(T) == 32) {
    if constexpr (SH == 1)
      asm volatile("st.global.cs.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(a.v[0]), "d"(a.v[1]), "d"(a.v[2]), "d"(a.v[3]) : "memory");
    else
      asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(a.v[0]), "d"(a.v[1]), "d"(a.v[2]), "d"(a.v[3]) : "memory");
  } else if constexpr (VEC == 4) {
    if constexpr (SH == 1) __stcs(reinterpret_cast<float4*>(p), make_float4(a.v[0], a.v[1], a.v[2], a.v[3]));
    else *reinterpret_cast<float4*>(p) = make_float4(a.v[0], a.v[1], a.v[2], a.v[3]);
  } else if constexpr (VEC == 2) {
    if constexpr (SH == 1) __stcs(reinterpret_cast<double2*>(p), make_double2(a.v[0], a.v[1]));
    else *reinterpret_cast<double2*>(p) = make_double2(a.v[0], a.v[1]);
  } else {
    p[0] = a.v[0];
  }
}

// streaming reads of the CSR arrays (each entry is used once per hop)
template <typename V>
__device__ __forceinline__ V ld_stream(const V* p) { return __ldcs(p); }

constexpr unsigned FULL = 0xffffffffu;

// Fused hop + collective (feature-sharded multi-GPU path): besides the local result row, each computed row slice is
// stored straight into the row-local contraction operand of the rank that owns that node row, through NVLink peer
// pointers (st.global on IPC-mapped memory).  Row r goes to peer r / rows_per_peer, local row r % rows_per_peer;
// local column (b, g) = b*gl + g lands at b*stride_b + out_col + g.  n_peers == 0 turns it off.
constexpr int MAX_PEERS = 16;
template <typename T>
struct ScatterArgs {
  T* peer[MAX_PEERS];
  int64_t rows_per_peer;
  int64_t out_ld;
  int64_t out_col;
  int64_t stride_b;
  int gl;
  int n_peers;
};

template <typename T, int VEC>
__device__ __forceinline__ void scatter_store(const ScatterArgs<T>& sc, int64_t row, int cbase, const Acc<T, VEC>& acc) {
  const int64_t q = row / sc.rows_per_peer;
  const int64_t lr = row - q * sc.rows_per_peer;
  const int b = cbase / sc.gl;
  const int g = cbase - b * sc.gl;
  T* o = sc.peer[q] + lr * sc.out_ld + (int64_t)b * sc.stride_b + sc.out_col + g;
  store_vec<T, VEC, 0>(o, acc);
}

template <typename T, int VEC, int L, int U, int THREADS, int MINB, int HINT, bool PF, int SH = 0>
__global__ void __launch_bounds__(THREADS, MINB)
spmm_hop_kernel(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                const T* __restrict__ val, const T* __restrict__ src, int64_t src_ld,
                T* __restrict__ dst, int64_t dst_ld, int64_t n_rows, int C, int n_chunks, float l2_frac,
                const ScatterArgs<T> sc) {
  uint64_t pol = 0;
  if constexpr (HINT == 2) pol = evict_last_policy(l2_frac);
  if constexpr (HINT == 3) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  if constexpr (HINT == 4) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 0.5;" : "=l"(pol));
  if constexpr (HINT == 5) asm volatile("createpolicy.fractional.L2::evict_last.L2::evict_first.b64 %0, 0.5;" : "=l"(pol));
  constexpr int S = 32 / L;  // neighbours gathered concurrently by one warp
  const int lane = threadIdx.x & 31;
  const int sub = lane / L;
  const int cl = lane % L;
  const int64_t n_warps = (int64_t)gridDim.x * (THREADS >> 5);
  const int64_t n_items = n_rows * n_chunks;
  int64_t item = (int64_t)blockIdx.x * (THREADS >> 5) + (threadIdx.x >> 5);
  if (item >= n_items) return;

  // software pipeline state: current row (beg, end, c, v), next row (nbeg, nend)
  int64_t row = item % n_rows;
  int64_t beg = __ldg(rowptr + row), end = __ldg(rowptr + row + 1);
  int32_t c = 0;
  T v = T(0);
  if (beg + lane < end) { c = ld_stream(col + beg + lane); v = ld_stream(val + beg + lane); }
  int64_t nbeg = 0, nend = 0;
  if (PF && item + n_warps < n_items) {
    const int64_t nrow = (item + n_warps) % n_rows;
    nbeg = __ldg(rowptr + nrow); nend = __ldg(rowptr + nrow + 1);
  }

  while (true) {
    const int chunk = (int)(item / n_rows);
    row = item - (int64_t)chunk * n_rows;
    const int cbase = chunk * (L * VEC) + cl * VEC;
    const bool col_ok = cbase < C;
    const T* __restrict__ srcc = src + cbase;

    // prefetch: next row's first 32 (col, val), and the rowptr pair of the row after that
    const int64_t next = item + n_warps;
    const bool has_next = next < n_items;
    int32_t nc = 0;
    T nv = T(0);
    int64_t nnbeg = 0, nnend = 0;
    if (PF) {
      if (has_next && nbeg + lane < nend) { nc = ld_stream(col + nbeg + lane); nv = ld_stream(val + nbeg + lane); }
      if (next + n_warps < n_items) {
        const int64_t nnrow = (next + n_warps) % n_rows;
        nnbeg = __ldg(rowptr + nnrow); nnend = __ldg(rowptr + nnrow + 1);
      }
    }

    Acc<T, VEC> acc;
    acc.zero();
    for (int64_t base = beg; base < end; base += 32) {
      if (base != beg) {  // rows longer than 32: fetch the following entries (not prefetched)
        c = 0; v = T(0);
        if (base + lane < end) { c = ld_stream(col + base + lane); v = ld_stream(val + base + lane); }
      }
      const int cnt = (int)((end - base) < 32 ? (end - base) : 32);
#pragma unroll 1
      for (int j = 0; j < cnt; j += S * U) {
        Acc<T, VEC> buf[U];
        T w[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int jj = j + u * S + sub;
          const int32_t cc = __shfl_sync(FULL, c, jj & 31);
          const T ww = __shfl_sync(FULL, v, jj & 31);
          const bool ok = (jj < cnt) && col_ok;
          w[u] = ok ? ww : T(0);
          if (ok) buf[u] = load_vec<T, VEC, HINT>(srcc + (int64_t)cc * src_ld, pol);
          else buf[u].zero();
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
#pragma unroll
          for (int i = 0; i < VEC; ++i) acc.v[i] = fma(w[u], buf[u].v[i], acc.v[i]);
        }
      }
    }
    // fold the S neighbour groups together
#pragma unroll
    for (int off = L; off < 32; off <<= 1) {
#pragma unroll
      for (int i = 0; i < VEC; ++i) acc.v[i] += __shfl_xor_sync(FULL, acc.v[i], off);
    }
    if (sub == 0 && col_ok) {
      store_vec<T, VEC, SH>(dst + row * dst_ld + cbase, acc);
      if (sc.n_peers > 0 && cbase < C) scatter_store<T, VEC>(sc, row, cbase, acc);
    }

    if (!has_next) break;
    item = next;
    if (PF) {
      beg = nbeg; end = nend; c = nc; v = nv;
      nbeg = nnbeg; nend = nnend;
    } else {
      row = item % n_rows;
      beg = __ldg(rowptr + row); end = __ldg(rowptr + row + 1);
      c = 0; v = T(0);
      if (beg + lane < end) { c = ld_stream(col + beg + lane); v = ld_stream(val + beg + lane); }
    }
  }
}


// ---------------------------------------------------------------------------------------------------
// Round-2 hop kernels.
//
// (1) spmm_hop_v2_kernel: the register-path (LDG.128) kernel above with the register pressure removed, so that the
//     40-register / 48-warp configuration no longer spills: 32-bit row and offset arithmetic (IDX = int32 rowptr when
//     nnz < 2^31), leading dimensions as 32-bit (IMAD.WIDE forms the 64-bit address), chunk loop outside the
//     grid-stride row loop (no 64-bit item / n_rows division), and the NVLink scatter arguments compiled out of the
//     plain instantiation (SCATTER = false carries an empty struct).
//
// (2) A variant that stages the gathered rows in shared memory with cp.async and software-pipelines the index chain three
//     units deep was built and measured (1.14-1.15 ms against 1.04 ms for (1), profiles/r2_spmm_sweep1_c64.log); it lives
//     with the sweep tool (tools/spmm_async_variant.cuh), not in the library.
// ---------------------------------------------------------------------------------------------------
// Fused hop + all-gather (node-sharded multi-GPU path): the rank computes rows [row0, row0 + n_rows) of the next hop's
// source matrix and every finished row is written into the full-height matrix of EVERY rank while the gather of the
// following rows is still in flight — either with one multimem.st through the NVSwitch multicast address `mc`
// (egress n_rows*C*s per hop) or, without multicast, with one NVLink peer store per rank (egress (P-1)/P*N*C*s).
template <typename T>
struct BcastArgs {
  T* peer[MAX_PEERS];     // full-height destination [n_total, out_ld] of every rank (own one included)
  T* mc;                  // multicast alias of the same buffer, or nullptr
  int64_t row0;           // global index of this rank's first row
  int64_t out_ld;
  int n_peers;
};

template <int VEC>
__device__ __forceinline__ void multimem_store(float* p, const Acc<float, VEC>& a) {
#pragma unroll
  for (int i = 0; i < VEC; i += 4)
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p + i), "f"(a.v[i]), "f"(a.v[i + 1]),
                 "f"(a.v[i + 2]), "f"(a.v[i + 3]) : "memory");
}
template <int VEC>
__device__ __forceinline__ void multimem_store(double* p, const Acc<double, VEC>& a) {
  // multimem.st has no f64 vector form; a store moves bits, so two doubles travel as one .v4.f32
#pragma unroll
  for (int i = 0; i < VEC; i += 2)
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p + i),
                 "f"(__int_as_float(__double2loint(a.v[i]))), "f"(__int_as_float(__double2hiint(a.v[i]))),
                 "f"(__int_as_float(__double2loint(a.v[i + 1]))), "f"(__int_as_float(__double2hiint(a.v[i + 1]))) : "memory");
}

template <typename T, int VEC>
__device__ __forceinline__ void bcast_store(const BcastArgs<T>& bc, int64_t row, int cbase, const Acc<T, VEC>& acc) {
  const int64_t off = (bc.row0 + row) * bc.out_ld + cbase;
  if (bc.mc != nullptr) {
    multimem_store<VEC>(bc.mc + off, acc);
  } else {
    for (int q = 0; q < bc.n_peers; ++q) store_vec<T, VEC, 0>(bc.peer[q] + off, acc);
  }
}

// epilogue of the v2 / async kernels.  MODE 0: plain hop; 1: feature-sharded scatter (ScatterArgs); 2: row broadcast.
//                               3 (2-D process grid): both — the row is all-gathered inside the rank's column group AND its
//                               slice is delivered to the row's contraction owner inside the rank's row group.
constexpr int EPI_NONE = 0, EPI_SCATTER = 1, EPI_BCAST = 2, EPI_GRID = 3;
template <typename T, int MODE>
struct ScatterParam {};
template <typename T>
struct ScatterParam<T, EPI_SCATTER> { ScatterArgs<T> a; };
template <typename T>
struct ScatterParam<T, EPI_BCAST> { BcastArgs<T> a; };
template <typename T>
struct ScatterParam<T, EPI_GRID> { ScatterArgs<T> s; BcastArgs<T> a; };

// shared epilogue of the v2 kernels
template <typename T, int VEC, int MODE, int SH>
__device__ __forceinline__ void hop_epilogue(const ScatterParam<T, MODE>& sp, T* __restrict__ dst, int dst_ld, int row, int cbase,
                                             const Acc<T, VEC>& acc) {
  if constexpr (MODE == EPI_BCAST) {
    bcast_store<T, VEC>(sp.a, row, cbase, acc);              // the rank's own copy is one of the destinations
  } else if constexpr (MODE == EPI_GRID) {
    if (sp.a.n_peers > 0) bcast_store<T, VEC>(sp.a, row, cbase, acc);   // n_peers == 0: last hop of a chain, no all-gather
    else if (dst != nullptr) store_vec<T, VEC, SH>(dst + (int64_t)row * dst_ld + cbase, acc);
    scatter_store<T, VEC>(sp.s, row, cbase, acc);
  } else {
    store_vec<T, VEC, SH>(dst + (int64_t)row * dst_ld + cbase, acc);
    if constexpr (MODE == EPI_SCATTER) {
      if (sp.a.n_peers > 0) scatter_store<T, VEC>(sp.a, row, cbase, acc);
    }
  }
}

// HINT: 0/1 no L2 policy; 3 evict_last on every gathered line; 2 evict_last on the fraction l2_frac of the lines (by
// address hash), the rest unchanged; 6 the same with evict_first on the rest.  SH = 1: streaming stores of the result.
template <typename T, typename IDX, int VEC, int L, int U, int THREADS, int MINB, int HINT, int SCATTER, int SH = 0>
__global__ void __launch_bounds__(THREADS, MINB)
spmm_hop_v2_kernel(const IDX* __restrict__ rowptr, const int32_t* __restrict__ col, const T* __restrict__ val,
                   const T* __restrict__ src, int src_ld, T* __restrict__ dst, int dst_ld, int n_rows, int C,
                   float l2_frac, const ScatterParam<T, SCATTER> sp) {
  uint64_t pol = 0;
  if constexpr (HINT == 3) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  if constexpr (HINT == 2) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, %1;" : "=l"(pol) : "f"(l2_frac));
  if constexpr (HINT == 6)
    asm volatile("createpolicy.fractional.L2::evict_last.L2::evict_first.b64 %0, %1;" : "=l"(pol) : "f"(l2_frac));
  constexpr int S = 32 / L;
  const int lane = threadIdx.x & 31;
  const int sub = lane / L;
  const int cl = lane % L;
  const int warp0 = blockIdx.x * (THREADS >> 5) + (threadIdx.x >> 5);
  {
    // column chunk = blockIdx.y: the block scheduler hands out all CTAs of chunk 0 first, so the resident warps gather
    // from one column slab at a time (chunk-major order) without a chunk loop in the kernel
    const int cbase = (int)blockIdx.y * (L * VEC) + cl * VEC;
    const bool col_ok = cbase < C;
    const T* __restrict__ srcc = src + cbase;
    for (int row = warp0; row < n_rows;) {
      IDX p = __ldg(rowptr + row);
      const IDX end = __ldg(rowptr + row + 1);
      Acc<T, VEC> acc;
      acc.zero();
      for (; p < end; p += 32) {
        int32_t c = 0;
        T v = T(0);
        const int cnt = (int)min((IDX)32, end - p);
        if (lane < cnt) { c = ld_stream(col + p + lane); v = ld_stream(val + p + lane); }
#pragma unroll 1
        for (int j = 0; j < cnt; j += S * U) {
          Acc<T, VEC> buf[U];
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const int jj = j + u * S + sub;
            const int32_t cc = __shfl_sync(FULL, c, jj & 31);
            if ((jj < cnt) && col_ok) buf[u] = load_vec<T, VEC, HINT>(srcc + (int64_t)cc * src_ld, pol);
            else buf[u].zero();
          }
#pragma unroll
          for (int u = 0; u < U; ++u) {   // weights are shuffled after the loads are in flight (saves U registers)
            const int jj = j + u * S + sub;
            const T ww = __shfl_sync(FULL, v, jj & 31);
            const T w = jj < cnt ? ww : T(0);
#pragma unroll
            for (int i = 0; i < VEC; ++i) acc.v[i] = fma(w, buf[u].v[i], acc.v[i]);
          }
        }
      }
#pragma unroll
      for (int off = L; off < 32; off <<= 1) {
#pragma unroll
        for (int i = 0; i < VEC; ++i) acc.v[i] += __shfl_xor_sync(FULL, acc.v[i], off);
      }
      // lane id and grid size are re-read here (volatile) so that the compiler does not park the loop-invariant store
      // predicate and row stride in local memory to fit the 40-register budget (it did: 24-byte "spill")
      unsigned ln, gdx;
      asm volatile("mov.u32 %0, %%laneid;" : "=r"(ln));
      asm volatile("mov.u32 %0, %%nctaid.x;" : "=r"(gdx));
      if (ln < (unsigned)L && cbase < C) hop_epilogue<T, VEC, SCATTER, SH>(sp, dst, dst_ld, row, cbase, acc);
      row += (int)gdx * (THREADS >> 5);
    }
  }
}

// Narrow rows with the v2 treatment (32-bit index math, 32-byte lanes, epilogue selected at compile time): a warp works on
// RPW = 32/GS consecutive rows at once, each group of GS lanes owns one row and its S = GS/L sub-groups of L lanes gather S
// neighbours per load.  C = 8 floats: L = 1 — one lane fetches a whole 32-byte neighbour row.
template <typename T, typename IDX, int VEC, int L, int GS, int U, int THREADS, int MINB, int HINT, int SCATTER>
__global__ void __launch_bounds__(THREADS, MINB)
spmm_hop_multirow_v2_kernel(const IDX* __restrict__ rowptr, const int32_t* __restrict__ col, const T* __restrict__ val,
                            const T* __restrict__ src, int src_ld, T* __restrict__ dst, int dst_ld, int n_rows, int C,
                            const ScatterParam<T, SCATTER> sp) {
  static_assert(GS % L == 0 && GS <= 32 && (GS / L) * U <= GS, "bad multirow geometry");
  constexpr int RPW = 32 / GS;
  constexpr int S = GS / L;
  uint64_t pol = 0;
  if constexpr (HINT >= 2) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  const int lane = threadIdx.x & 31;
  const int grp = lane / GS, gl = lane % GS;
  const int sub = gl / L, cl = gl % L;
  const int cbase = cl * VEC;
  const bool col_ok = cbase < C;
  const T* __restrict__ srcc = src + cbase;
  const int n_warps = gridDim.x * (THREADS >> 5);
  for (int row0 = (blockIdx.x * (THREADS >> 5) + (threadIdx.x >> 5)) * RPW; row0 < n_rows; row0 += n_warps * RPW) {
    const int row = row0 + grp;
    const bool row_ok = row < n_rows;
    const IDX beg = row_ok ? __ldg(rowptr + row) : 0;
    const int len = row_ok ? (int)(__ldg(rowptr + row + 1) - beg) : 0;
    int maxlen = len;  // warp-uniform trip count: the longest of the RPW rows
#pragma unroll
    for (int off = GS; off < 32; off <<= 1) maxlen = max(maxlen, __shfl_xor_sync(FULL, maxlen, off));
    Acc<T, VEC> acc;
    acc.zero();
    for (int b0 = 0; b0 < maxlen; b0 += GS) {
      int32_t c = 0;
      T v = T(0);
      if (b0 + gl < len) { c = ld_stream(col + beg + b0 + gl); v = ld_stream(val + beg + b0 + gl); }
      const int cnt = len - b0;           // entries of my row left in this chunk (may be <= 0)
      const int maxcnt = maxlen - b0;
#pragma unroll
      for (int j = 0; j < GS; j += S * U) {
        if (j >= maxcnt) break;           // warp-uniform
        Acc<T, VEC> buf[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int jj = j + u * S + sub;
          const int32_t cc = __shfl_sync(FULL, c, jj, GS);
          if ((jj < cnt) && col_ok) buf[u] = load_vec<T, VEC, HINT>(srcc + (int64_t)cc * src_ld, pol);
          else buf[u].zero();
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int jj = j + u * S + sub;
          const T ww = __shfl_sync(FULL, v, jj, GS);
          const T w = jj < cnt ? ww : T(0);
#pragma unroll
          for (int i = 0; i < VEC; ++i) acc.v[i] = fma(w, buf[u].v[i], acc.v[i]);
        }
      }
    }
#pragma unroll
    for (int off = L; off < GS; off <<= 1) {
#pragma unroll
      for (int i = 0; i < VEC; ++i) acc.v[i] += __shfl_xor_sync(FULL, acc.v[i], off);
    }
    if (sub == 0 && col_ok && row_ok) hop_epilogue<T, VEC, SCATTER, 0>(sp, dst, dst_ld, row, cbase, acc);
  }
}

// Narrow feature rows (C * sizeof(T) <= 128 bytes): one warp per row leaves most lanes idle and the kernel
// latency-bound on the rowptr -> col/val -> gather dependency chain (measured 0.38 ms at C = 8 vs 1.27 ms at C = 64,
// N = 1M).  Here a warp works on RPW = 32/GS consecutive rows at once: each group of GS lanes owns one row, reads
// its col/val GS entries at a time, and its S = GS/L sub-groups of L lanes gather S neighbours per load.
template <typename T, int VEC, int L, int GS, int U, int THREADS, int MINB, int HINT>
__global__ void __launch_bounds__(THREADS, MINB)
spmm_hop_multirow_kernel(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                         const T* __restrict__ val, const T* __restrict__ src, int64_t src_ld,
                         T* __restrict__ dst, int64_t dst_ld, int64_t n_rows, int C, const ScatterArgs<T> sc) {
  static_assert(GS % L == 0 && GS <= 32 && (GS / L) * U <= GS, "bad multirow geometry");
  constexpr int RPW = 32 / GS;
  constexpr int S = GS / L;
  uint64_t pol = 0;
  if constexpr (HINT >= 2) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  const int lane = threadIdx.x & 31;
  const int grp = lane / GS, gl = lane % GS;
  const int sub = gl / L, cl = gl % L;
  const int cbase = cl * VEC;
  const bool col_ok = cbase < C;
  const T* __restrict__ srcc = src + cbase;
  const int64_t n_warps = (int64_t)gridDim.x * (THREADS >> 5);
  const int64_t warp_global = (int64_t)blockIdx.x * (THREADS >> 5) + (threadIdx.x >> 5);
  for (int64_t row0 = warp_global * RPW; row0 < n_rows; row0 += n_warps * RPW) {
    const int64_t row = row0 + grp;
    const bool row_ok = row < n_rows;
    const int64_t beg = row_ok ? __ldg(rowptr + row) : 0;
    const int len = row_ok ? (int)(__ldg(rowptr + row + 1) - beg) : 0;
    int maxlen = len;  // warp-uniform trip count: the longest of the RPW rows
#pragma unroll
    for (int off = GS; off < 32; off <<= 1) maxlen = max(maxlen, __shfl_xor_sync(FULL, maxlen, off));
    Acc<T, VEC> acc;
    acc.zero();
    for (int b0 = 0; b0 < maxlen; b0 += GS) {
      int32_t c = 0;
      T v = T(0);
      if (b0 + gl < len) { c = ld_stream(col + beg + b0 + gl); v = ld_stream(val + beg + b0 + gl); }
      const int cnt = len - b0;           // entries of my row left in this chunk (may be <= 0)
      const int maxcnt = maxlen - b0;
#pragma unroll
      for (int j = 0; j < GS; j += S * U) {
        if (j >= maxcnt) break;           // warp-uniform
        Acc<T, VEC> buf[U];
        T w[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int jj = j + u * S + sub;
          const int32_t cc = __shfl_sync(FULL, c, jj, GS);
          const T ww = __shfl_sync(FULL, v, jj, GS);
          const bool ok = (jj < cnt) && col_ok;
          w[u] = ok ? ww : T(0);
          if (ok) buf[u] = load_vec<T, VEC, HINT>(srcc + (int64_t)cc * src_ld, pol);
          else buf[u].zero();
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
#pragma unroll
          for (int i = 0; i < VEC; ++i) acc.v[i] = fma(w[u], buf[u].v[i], acc.v[i]);
        }
      }
    }
#pragma unroll
    for (int off = L; off < GS; off <<= 1) {
#pragma unroll
      for (int i = 0; i < VEC; ++i) acc.v[i] += __shfl_xor_sync(FULL, acc.v[i], off);
    }
    if (sub == 0 && col_ok && row_ok) {
      store_vec<T, VEC, 0>(dst + row * dst_ld + cbase, acc);
      if (sc.n_peers > 0) scatter_store<T, VEC>(sc, row, cbase, acc);
    }
  }
}

}  // namespace b200gf
