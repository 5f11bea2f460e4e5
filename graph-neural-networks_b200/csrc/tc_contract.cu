// Tensor-core tap contraction for FP32 (sm_100a: TMA + tcgen05.mma kind::tf32 + TMEM), error-compensated ("3xTF32").
//
//   out[r, q] = bias[q] + sum_t sum_p Z_t[r, p] * W_t[p, q]          r = node * B + b
//
// replaces the reference's [B,N,EKG] x [EKG,F] torch.matmul + bias (alegnn/utils/graphML.py:170-175).
// A single TF32 pass misses the 1e-4 tolerance over E*K*G = 320 terms (SURVEY.md §7 hard part 3), so every operand
// is split x = hi + lo with hi = the 19 leading bits (exact in TF32) and the product is accumulated in FP32 in TMEM as
//   hi*hi + lo*hi + hi*lo            (the dropped lo*lo term is ~2^-22 relative).
//
// One persistent CTA per SM, 320 threads, warp-specialised:
//   warp 0        TMA producer: per 32-float k-chunk, one 128 x 32 tile of Z_t (128-byte swizzle) + the matching
//                 Q x 32 tiles of W_hi and W_lo (pre-split, K-major) into a 4-stage shared-memory ring
//   warps 2-5     splitter: writes lo = z - hi(z) next to the Z tile (same swizzled layout: the transform is
//                 element-wise); the tile itself serves as the hi operand (the MMA truncates FP32 to TF32, which is
//                 exactly hi), fence.proxy.async, then hands the stage to
//   warp 1        MMA issuer: one elected lane issues 3 x 4 tcgen05.mma (M=128, N=Q, K=8) per stage into one of
//                 two TMEM accumulators (so the epilogue of tile i overlaps the MMAs of tile i+1); tcgen05.commit
//                 releases the stage to the producer and, after the last chunk, the accumulator to
//   warps 6-9     epilogue: tcgen05.ld (32x32b), + bias, 16-byte stores of whole output rows.
#include <cuda.h>

#include <cstdlib>

#include "common.cuh"

namespace b200gf {
namespace tc {

constexpr int BM = 128;       // rows per tile  (UMMA M)
constexpr int BK = 32;        // floats per k-chunk = one 128-byte swizzle row
constexpr int UMMA_K = 8;     // tf32: 32 bytes per instruction along K
constexpr int MAX_T = 16;     // terms (tensor maps travel as kernel parameters)
constexpr int THREADS = 320;
constexpr int A_BYTES = BM * BK * 4;  // 16 KB
constexpr unsigned SPIN_LIMIT = 1u << 24;  // bounded waits: a protocol bug traps instead of hanging the GPU

struct Params {
  CUtensorMap a_map[MAX_T];
  CUtensorMap bhi_map, blo_map;
  const float* bias;
  float* out;
  int64_t out_ld;
  int64_t R;        // total rows = n_rows * B
  int64_t n_rows;
  int T, P, Q, B;
  int num_tiles;
  int stages;
  int bias_per_node;
  int raw_hi;        // leave the TMA tile in place as the hi operand: the tensor core reads the top 19 bits of an FP32 value
                     // (truncation), which IS hi — measured bit-identical to rewriting it (profiles/README.md), 16 KB less
                     // shared-memory traffic per stage.  B200GF_TC_RAWHI=0 restores the rewrite.
  int relu;          // epilogue activation: out = max(out, 0)  (fused GraphFilter -> ReLU layer, architectures.py:287)
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  unsigned spins = 0;
  while (true) {
    asm volatile(
        "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
        : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (done) break;
    if (++spins > SPIN_LIMIT) __trap();
  }
}

__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint32_t bar, uint32_t dst, int x, int y) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(dst), "l"(map), "r"(bar), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ void tma_load_3d(const CUtensorMap* map, uint32_t bar, uint32_t dst, int x, int y, int z) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
               ::"r"(dst), "l"(map), "r"(bar), "r"(x), "r"(y), "r"(z) : "memory");
}

// K-major, 128-byte-swizzled operand tile: 8-row groups 1024 bytes apart (SBO), LBO unused (=1), descriptor v1
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}

__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}"
               ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ float tf32_hi(float v) { return __uint_as_float(__float_as_uint(v) & 0xFFFFE000u); }

__global__ void __launch_bounds__(THREADS, 1) tc_contract_kernel(const __grid_constant__ Params prm) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  // 1024-byte alignment for the 128-byte swizzle atoms
  unsigned char* smem = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int Q = prm.Q;
  const int S = prm.stages;
  const int b_bytes = Q * BK * 4;
  const int stage_bytes = 2 * A_BYTES + 2 * b_bytes;
  unsigned char* bar_base = smem + (size_t)S * stage_bytes;
  uint64_t* full = (uint64_t*)bar_base;            // [S] TMA landed
  uint64_t* split = full + S;                      // [S] hi/lo written
  uint64_t* empty = split + S;                     // [S] MMAs done with the stage
  uint64_t* tmem_full = empty + S;                 // [2]
  uint64_t* tmem_empty = tmem_full + 2;            // [2]
  uint32_t* tmem_ptr = (uint32_t*)(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cpt = prm.P / BK;                      // k-chunks per term
  const int chunks = prm.T * cpt;
  int tmem_cols = 32;
  while (tmem_cols < 2 * Q) tmem_cols <<= 1;

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(smem_u32(full + s), 1);
      mbar_init(smem_u32(split + s), 128);
      mbar_init(smem_u32(empty + s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(smem_u32(tmem_full + a), 1);
      mbar_init(smem_u32(tmem_empty + a), 128);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {  // TMEM allocation (whole warp), address lands in shared memory
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "r"(tmem_cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < prm.num_tiles; tile += gridDim.x) {
        for (int t = 0; t < prm.T; ++t) {
          for (int pc = 0; pc < cpt; ++pc, ++it) {
            const int s = it % S;
            const uint32_t ph = (it / S) & 1;
            mbar_wait(smem_u32(empty + s), ph ^ 1);
            unsigned char* st = smem + (size_t)s * stage_bytes;
            const uint32_t bar = smem_u32(full + s);
            mbar_expect_tx(bar, A_BYTES + 2 * b_bytes);
            tma_load_2d(&prm.a_map[t], bar, smem_u32(st), pc * BK, tile * BM);
            tma_load_3d(&prm.bhi_map, bar, smem_u32(st + 2 * A_BYTES), pc * BK, 0, t);
            tma_load_3d(&prm.blo_map, bar, smem_u32(st + 2 * A_BYTES + b_bytes), pc * BK, 0, t);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    // instruction descriptor: D=F32 (bit 4), A=B=TF32 (2 at bits 7 and 10), both K-major, N>>3 at 17, M>>4 at 24
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(Q >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
    uint32_t it = 0, tile_iter = 0;
    for (int tile = blockIdx.x; tile < prm.num_tiles; tile += gridDim.x, ++tile_iter) {
      const int a = tile_iter & 1;
      const uint32_t aph = (tile_iter >> 1) & 1;
      mbar_wait(smem_u32(tmem_empty + a), aph ^ 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t d_tmem = tmem_base + (uint32_t)(a * Q);
      for (int c = 0; c < chunks; ++c, ++it) {
        const int s = it % S;
        const uint32_t ph = (it / S) & 1;
        mbar_wait(smem_u32(split + s), ph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (lane == 0) {
          const uint32_t st = smem_u32(smem + (size_t)s * stage_bytes);
#pragma unroll
          for (int j = 0; j < BK / UMMA_K; ++j) {
            const uint64_t a_hi = umma_desc(st + j * UMMA_K * 4);
            const uint64_t a_lo = umma_desc(st + A_BYTES + j * UMMA_K * 4);
            const uint64_t b_hi = umma_desc(st + 2 * A_BYTES + j * UMMA_K * 4);
            const uint64_t b_lo = umma_desc(st + 2 * A_BYTES + b_bytes + j * UMMA_K * 4);
            umma_tf32(d_tmem, a_hi, b_hi, idesc, (c > 0 || j > 0) ? 1u : 0u);
            umma_tf32(d_tmem, a_lo, b_hi, idesc, 1u);
            umma_tf32(d_tmem, a_hi, b_lo, idesc, 1u);
          }
          umma_commit(smem_u32(empty + s));                      // stage free once these MMAs retire
          if (c == chunks - 1) umma_commit(smem_u32(tmem_full + a));  // accumulator complete
        }
        __syncwarp();
      }
    }
  } else if (warp < 6) {
    // ------------------------------------------------------------------ splitter (128 threads)
    const int tid = threadIdx.x - 64;
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < prm.num_tiles; tile += gridDim.x) {
      for (int c = 0; c < chunks; ++c, ++it) {
        const int s = it % S;
        const uint32_t ph = (it / S) & 1;
        mbar_wait(smem_u32(full + s), ph);
        float4* A = (float4*)(smem + (size_t)s * stage_bytes);
        float4* Alo = (float4*)(smem + (size_t)s * stage_bytes + A_BYTES);
#pragma unroll
        for (int i = 0; i < A_BYTES / 16 / 128; ++i) {
          const int idx = i * 128 + tid;
          const float4 v = A[idx];
          float4 hi, lo;
          hi.x = tf32_hi(v.x); hi.y = tf32_hi(v.y); hi.z = tf32_hi(v.z); hi.w = tf32_hi(v.w);
          lo.x = v.x - hi.x; lo.y = v.y - hi.y; lo.z = v.z - hi.z; lo.w = v.w - hi.w;
          if (!prm.raw_hi) A[idx] = hi;
          Alo[idx] = lo;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic writes -> visible to the tensor core
        mbar_arrive(smem_u32(split + s));
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (128 threads)
    const int q4 = warp & 3;  // TMEM lane quarter this warp may read
    uint32_t tile_iter = 0;
    for (int tile = blockIdx.x; tile < prm.num_tiles; tile += gridDim.x, ++tile_iter) {
      const int a = tile_iter & 1;
      const uint32_t aph = (tile_iter >> 1) & 1;
      mbar_wait(smem_u32(tmem_full + a), aph);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int64_t r = (int64_t)tile * BM + q4 * 32 + lane;
      const bool row_ok = r < prm.R;
      const int64_t n = row_ok ? r / prm.B : 0;
      const int b = row_ok ? (int)(r - n * prm.B) : 0;
      float* orow = prm.out + n * prm.out_ld + (int64_t)b * Q;
      for (int c0 = 0; c0 < Q; c0 += 16) {
        const uint32_t taddr = tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(a * Q + c0);
        uint32_t v[16];
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
              "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
            : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (row_ok) {
#pragma unroll
          for (int j = 0; j < 16; j += 4) {
            float4 o;
            o.x = __uint_as_float(v[j]); o.y = __uint_as_float(v[j + 1]);
            o.z = __uint_as_float(v[j + 2]); o.w = __uint_as_float(v[j + 3]);
            if (prm.bias) {
              const int q = c0 + j;
              if (prm.bias_per_node) {
                o.x += prm.bias[(int64_t)q * prm.n_rows + n]; o.y += prm.bias[(int64_t)(q + 1) * prm.n_rows + n];
                o.z += prm.bias[(int64_t)(q + 2) * prm.n_rows + n]; o.w += prm.bias[(int64_t)(q + 3) * prm.n_rows + n];
              } else {
                o.x += __ldg(prm.bias + q); o.y += __ldg(prm.bias + q + 1);
                o.z += __ldg(prm.bias + q + 2); o.w += __ldg(prm.bias + q + 3);
              }
            }
            if (prm.relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
            *reinterpret_cast<float4*>(orow + c0 + j) = o;
          }
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      mbar_arrive(smem_u32(tmem_empty + a));
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols));
  }
}

// taps h[F,E,K,G] -> K-major hi / lo operand tiles Whi/Wlo [t][Q][P]
//   forward  (to_input = 0): Q = F, P = G,  W[t][f][g] = tap(t, f, g)
//   backward (to_input = 1): Q = G, P = F,  W[t][g][f] = tap(t, f, g)
__global__ void pack_taps_split_kernel(const float* __restrict__ h, float* __restrict__ Whi, float* __restrict__ Wlo,
                                       int F, int E, int K, int G, int to_input) {
  const int Tn = 1 + E * (K - 1);
  const int64_t total = (int64_t)Tn * G * F;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int t = (int)(i / ((int64_t)G * F));
    const int rem = (int)(i - (int64_t)t * G * F);
    int f, g;
    if (to_input) { g = rem / F; f = rem % F; }
    else { f = rem / G; g = rem % G; }
    float val;
    if (t == 0) {
      val = 0.f;
      for (int e = 0; e < E; ++e) val += h[(((int64_t)f * E + e) * K + 0) * G + g];
    } else {
      const int e = (t - 1) / (K - 1), k = (t - 1) % (K - 1) + 1;
      val = h[(((int64_t)f * E + e) * K + k) * G + g];
    }
    const float hi = tf32_hi(val);
    Whi[i] = hi;
    Wlo[i] = val - hi;
  }
}

// generic W[T][P][Q] (p-major, the b200gf_tap_contract layout) -> Whi/Wlo [t][Q][P]
__global__ void split_w_kernel(const float* __restrict__ W, float* __restrict__ Whi, float* __restrict__ Wlo, int T, int P,
                               int Q) {
  const int64_t total = (int64_t)T * P * Q;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int t = (int)(i / ((int64_t)P * Q));
    const int rem = (int)(i - (int64_t)t * P * Q);
    const int q = rem / P, p = rem % P;
    const float val = W[((int64_t)t * P + p) * Q + q];
    const float hi = tf32_hi(val);
    Whi[i] = hi;
    Wlo[i] = val - hi;
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
    else
      (void)cudaGetLastError();
  }
  return fn;
}

static int stages_for(int Q) {
  const int stage_bytes = 2 * A_BYTES + 2 * Q * BK * 4;
  int s = (200 * 1024) / stage_bytes;
  if (s > 4) s = 4;
  return s;
}

}  // namespace tc

size_t tc_contract_scratch_bytes(int T, int P, int Q) { return align_up((size_t)2 * T * P * Q * sizeof(float), 256); }

bool tc_contract_eligible(int dtype, int64_t n_rows, int B, int P, int Q, int T, const void* const* zs,
                          const int64_t* z_ld, const void* out, int64_t out_ld, int accumulate) {
  if (dtype != B200GF_F32 || accumulate) return false;
  if (T < 1 || T > tc::MAX_T) return false;
  if (P % tc::BK != 0 || Q % 16 != 0 || Q < 16 || Q > 256) return false;
  if (tc::stages_for(Q) < 2) return false;
  if (n_rows * B < tc::BM) return false;                        // tiny problems: the FMA kernel is launch-bound anyway
  if (n_rows * B > (int64_t)INT32_MAX) return false;             // TMA coordinates are 32-bit
  if ((reinterpret_cast<uintptr_t>(out) & 15) != 0 || out_ld % 4 != 0) return false;
  for (int t = 0; t < T; ++t) {
    if (z_ld[t] != (int64_t)B * P) return false;                 // rows (n, b) must be uniformly P apart
    if ((reinterpret_cast<uintptr_t>(zs[t]) & 15) != 0) return false;
  }
  return tc::get_encode() != nullptr;
}

int launch_pack_taps_split(const void* h, void* whi_wlo, int F, int E, int K, int G, int to_input, cudaStream_t st) {
  const int64_t total = (int64_t)(1 + E * (K - 1)) * G * F;
  float* hi = (float*)whi_wlo;
  float* lo = hi + total;
  const int blocks = (int)imin64((total + 255) / 256, 1184);
  tc::pack_taps_split_kernel<<<blocks, 256, 0, st>>>((const float*)h, hi, lo, F, E, K, G, to_input);
  LAUNCH_CHECK();
  return B200GF_OK;
}

int launch_split_w(const void* W, void* whi_wlo, int T, int P, int Q, cudaStream_t st) {
  const int64_t total = (int64_t)T * P * Q;
  float* hi = (float*)whi_wlo;
  float* lo = hi + total;
  const int blocks = (int)imin64((total + 255) / 256, 1184);
  tc::split_w_kernel<<<blocks, 256, 0, st>>>((const float*)W, hi, lo, T, P, Q);
  LAUNCH_CHECK();
  return B200GF_OK;
}

// whi_wlo: device [2][T][Q][P] (hi then lo), K-major; everything else as launch_tap_contract
int launch_tc_contract(int sm_count, int64_t n_rows, int B, int P, int Q, int T, const void* const* zs,
                       const void* whi_wlo, const void* bias, int bias_per_node, void* out, int64_t out_ld,
                       cudaStream_t st, int act) {
  using namespace tc;
  EncodeTiledFn encode = get_encode();
  if (!encode) return B200GF_EUNSUPPORTED;
  Params prm;
  const int64_t R = n_rows * B;
  for (int t = 0; t < T; ++t) {
    const cuuint64_t dims[2] = {(cuuint64_t)P, (cuuint64_t)R};
    const cuuint64_t strides[1] = {(cuuint64_t)P * 4};
    const cuuint32_t box[2] = {BK, BM};
    const cuuint32_t estr[2] = {1, 1};
    if (encode(&prm.a_map[t], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(zs[t]), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return B200GF_EINVAL;
  }
  for (int t = T; t < MAX_T; ++t) prm.a_map[t] = prm.a_map[0];
  {
    const cuuint64_t dims[3] = {(cuuint64_t)P, (cuuint64_t)Q, (cuuint64_t)T};
    const cuuint64_t strides[2] = {(cuuint64_t)P * 4, (cuuint64_t)P * Q * 4};
    const cuuint32_t box[3] = {BK, (cuuint32_t)Q, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    float* hi = (float*)const_cast<void*>(whi_wlo);
    float* lo = hi + (int64_t)T * P * Q;
    if (encode(&prm.bhi_map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, hi, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return B200GF_EINVAL;
    if (encode(&prm.blo_map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, lo, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return B200GF_EINVAL;
  }
  prm.bias = (const float*)bias;
  prm.out = (float*)out;
  prm.out_ld = out_ld;
  prm.R = R;
  prm.n_rows = n_rows;
  prm.T = T; prm.P = P; prm.Q = Q; prm.B = B;
  prm.num_tiles = (int)((R + BM - 1) / BM);
  prm.stages = stages_for(Q);
  prm.bias_per_node = bias_per_node;
  prm.relu = act;
  {
    static const int raw = [] { const char* e = getenv("B200GF_TC_RAWHI"); return (e && e[0] == '0') ? 0 : 1; }();
    prm.raw_hi = raw;
  }
  const int stage_bytes = 2 * A_BYTES + 2 * Q * BK * 4;
  const size_t smem = (size_t)prm.stages * stage_bytes + 1024 /*alignment slack*/ + 256 /*barriers*/;
  CUDA_TRY(cudaFuncSetAttribute(tc_contract_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  const int grid = prm.num_tiles < sm_count ? prm.num_tiles : sm_count;
  tc_contract_kernel<<<grid, THREADS, smem, st>>>(prm);
  LAUNCH_CHECK();
  return B200GF_OK;
}

}  // namespace b200gf
