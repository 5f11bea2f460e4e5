// Probe for the next backward-pass kernel: tap gradient on tensor cores.
//
//     dW[t][p][q] = sum_{r < R} A[r, p] * V_t[r, q]          (FP32 in/out, 3xTF32: hi*hi + lo*hi + hi*lo)
//
// NOT part of libb200gf.so.  Standalone: builds with
//     nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o /tmp/tapgrad_tc_probe tools/tapgrad_tc_probe.cu
// and runs a correctness check against a double-precision CPU sum plus a timing loop at the headline size
// (R = 1M, P = Q = 64, T = 5).  The shipped kernel it would replace is tap_grad_multi_kernel (FMA, 1.37 ms).
//
// Formulation.  Both operands have the reduction index r as the STRIDED one in memory (rows of A / V_t are contiguous in
// p / q), i.e. they are "MN-major" for the tensor core.  Instead of MN-major descriptors, the loader warps — which have
// to rewrite every element anyway for the hi/lo split — store the tiles TRANSPOSED into the K-major, 128-byte-swizzled
// layout that tc_contract.cu already uses (same shared-memory descriptors, same instruction descriptor form):
//     A-operand (M = 128, K = 32):  row m = tt*64 + q  holds  V_{2j+tt}[r0 .. r0+31, q]      (two terms per MMA)
//     B-operand (N =  64, K = 32):  row p              holds  A[r0 .. r0+31, p]
//     D_j[(tt, q), p] (TMEM, 64 columns per pair j) += A-operand x B-operand^T
// A lane owns one r (one k index) and scatters the 4 values of each float4 to 4 consecutive rows: for a fixed row the
// 32 lanes of a store hit the 32 different banks of that 128-byte swizzled row, so the transposing stores are
// conflict-free.  One persistent CTA per SM walks its row chunk in steps of 32 rows through a 2-stage ring
// (112 KB per stage: 3 pairs x (hi, lo) x 16 KB + (hi, lo) x 8 KB), one lane issues 36 tcgen05.mma per step, and the
// per-CTA partial sums go to partial[cta][t][p][q] for the deterministic second pass the library already has.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

namespace probe {

constexpr int P = 64, Q = 64;       // this probe fixes the headline shape
constexpr int BK = 32;              // rows per step = one 128-byte swizzled k-row
constexpr int UMMA_K = 8;
constexpr int MAX_T = 6;
constexpr int PAIRS = MAX_T / 2;
constexpr int LOADER_WARPS = 8;
constexpr int THREADS = 32 * (1 + LOADER_WARPS);
constexpr int VT_BYTES = 128 * BK * 4;   // 16 KB: one (pair, hi|lo) A-operand tile
constexpr int AT_BYTES = P * BK * 4;     //  8 KB: one (hi|lo) B-operand tile
constexpr int STAGE_BYTES = PAIRS * 2 * VT_BYTES + 2 * AT_BYTES;   // 112 KB
constexpr int STAGES = 2;
constexpr unsigned SPIN_LIMIT = 1u << 24;

struct Params {
  const float* A; int64_t a_ld;
  const float* V[MAX_T]; int64_t v_ld[MAX_T];
  float* partial;             // [gridDim.x][T][P][Q]
  int64_t R, rows_per_cta;
  int T;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
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
// K-major, 128-byte swizzle, 8-row groups 1024 bytes apart — identical to tc_contract.cu (validated on B200)
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
__device__ __forceinline__ void sts_f32(uint32_t addr, float v) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ float tf32_hi(float v) { return __uint_as_float(__float_as_uint(v) & 0xFFFFE000u); }

// byte offset of element (row m, k) inside a K-major 128B-swizzled tile (what TMA SWIZZLE_128B would have produced)
__device__ __forceinline__ uint32_t swz(int m, int k) {
  return (uint32_t)((m >> 3) * 1024 + (m & 7) * 128 + ((((k >> 2) ^ (m & 7)) & 7) << 4) + ((k & 3) << 2));
}

__global__ void __launch_bounds__(THREADS, 1) tapgrad_tc_kernel(const __grid_constant__ Params prm) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* split = (uint64_t*)(smem + (size_t)STAGES * STAGE_BYTES);   // [STAGES] tiles written
  uint64_t* empty = split + STAGES;                                      // [STAGES] MMAs retired
  uint64_t* done = empty + STAGES;                                       // accumulators complete
  uint32_t* tmem_ptr = (uint32_t*)(done + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int T = prm.T;
  const int pairs = (T + 1) / 2;
  constexpr int tmem_cols = 256;   // PAIRS * 64 = 192 accumulator columns

  // zero both stages once: the unused half of an odd last pair is never written afterwards
  for (int i = threadIdx.x; i < STAGES * STAGE_BYTES / 16; i += THREADS) ((float4*)smem)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(smem_u32(split + s), 32 * LOADER_WARPS);
      mbar_init(smem_u32(empty + s), 1);
    }
    mbar_init(smem_u32(done), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "r"(tmem_cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");          // the zero fill is read by the tensor core
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_ptr;

  const int64_t row0 = (int64_t)blockIdx.x * prm.rows_per_cta;
  int64_t row1 = row0 + prm.rows_per_cta;
  if (row1 > prm.R) row1 = prm.R;
  const int steps = row1 > row0 ? (int)((row1 - row0 + BK - 1) / BK) : 0;

  if (warp == 0) {
    // ------------------------------------------------------------------ MMA issuer
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(P >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    for (int it = 0; it < steps; ++it) {
      const int s = it & 1;
      const uint32_t ph = (it >> 1) & 1;
      mbar_wait(smem_u32(split + s), ph);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (lane == 0) {
        const uint32_t st = smem_u32(smem + (size_t)s * STAGE_BYTES);
        const uint32_t at = st + PAIRS * 2 * VT_BYTES;
        for (int j = 0; j < pairs; ++j) {
          const uint32_t vt = st + j * 2 * VT_BYTES;
          const uint32_t d_tmem = tmem_base + (uint32_t)(j * P);
#pragma unroll
          for (int kk = 0; kk < BK / UMMA_K; ++kk) {
            const uint64_t v_hi = umma_desc(vt + kk * UMMA_K * 4);
            const uint64_t v_lo = umma_desc(vt + VT_BYTES + kk * UMMA_K * 4);
            const uint64_t a_hi = umma_desc(at + kk * UMMA_K * 4);
            const uint64_t a_lo = umma_desc(at + AT_BYTES + kk * UMMA_K * 4);
            umma_tf32(d_tmem, v_hi, a_hi, idesc, (it > 0 || kk > 0) ? 1u : 0u);
            umma_tf32(d_tmem, v_lo, a_hi, idesc, 1u);
            umma_tf32(d_tmem, v_hi, a_lo, idesc, 1u);
          }
        }
        umma_commit(smem_u32(empty + s));
        if (it == steps - 1) umma_commit(smem_u32(done));
      }
      __syncwarp();
    }
  } else {
    // ------------------------------------------------------------------ loaders: global rows -> hi/lo, transposed + swizzled
    const int lw = warp - 1;                        // 0 .. LOADER_WARPS-1
    const int n4 = (P + T * Q) / 4;                 // float4 per row of [A | V_0 .. V_{T-1}]
    for (int it = 0; it < steps; ++it) {
      const int s = it & 1;
      const uint32_t ph = (it >> 1) & 1;
      mbar_wait(smem_u32(empty + s), ph ^ 1);
      const uint32_t st = smem_u32(smem + (size_t)s * STAGE_BYTES);
      const uint32_t at = st + PAIRS * 2 * VT_BYTES;
      const int64_t r = row0 + (int64_t)it * BK + lane;          // this lane's row = k index `lane`
      const bool ok = r < row1;
      for (int j4 = lw; j4 < n4; j4 += LOADER_WARPS) {
        const int c = j4 * 4;                       // column in the concatenated row
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        uint32_t hi_tile;
        int m0;
        uint32_t lo_off;
        if (c < P) {
          if (ok) v = __ldg(reinterpret_cast<const float4*>(prm.A + r * prm.a_ld + c));
          hi_tile = at; m0 = c; lo_off = AT_BYTES;
        } else {
          const int t = (c - P) / Q, q = (c - P) % Q;
          if (ok) v = __ldg(reinterpret_cast<const float4*>(prm.V[t] + r * prm.v_ld[t] + q));
          hi_tile = st + (t >> 1) * 2 * VT_BYTES; m0 = (t & 1) * 64 + q; lo_off = VT_BYTES;
        }
        const float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float h = tf32_hi(e[i]);
          const uint32_t off = swz(m0 + i, lane);
          sts_f32(hi_tile + off, h);
          sts_f32(hi_tile + lo_off + off, e[i] - h);
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_arrive(smem_u32(split + s));
    }
    // ------------------------------------------------------------------ epilogue: TMEM -> partial[cta][t][p][q]
    if (steps > 0) {
      mbar_wait(smem_u32(done), 0);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    const int quad = warp & 3;                      // TMEM lanes 32*quad .. +31 are readable by this warp
    const int half = lw >> 2;                       // two warps share a quadrant: alternate 16-column groups
    const int m = quad * 32 + lane;                 // accumulator row = (tt, q)
    float* out = prm.partial + (size_t)blockIdx.x * T * P * Q;
    for (int g = half; g < pairs * (P / 16); g += 2) {
      const int j = g / (P / 16), p0 = (g % (P / 16)) * 16;
      uint32_t v[16];
      if (steps > 0) {
        const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(j * P + p0);
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
              "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
            : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = 0u;
      }
      const int t = 2 * j + (m >> 6), q = m & 63;
      if (t < T) {
#pragma unroll
        for (int i = 0; i < 16; ++i) out[((size_t)t * P + (p0 + i)) * Q + q] = __uint_as_float(v[i]);   // lanes: 32 consecutive q
      }
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols));
  }
}

// second pass (deterministic): dW[i] = sum_c partial[c][i]
__global__ void reduce_kernel(const float* __restrict__ partial, float* __restrict__ dW, int n_chunks, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float acc = 0.f;
    for (int c = 0; c < n_chunks; ++c) acc += partial[(size_t)c * n + i];
    dW[i] = acc;
  }
}

}  // namespace probe

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 1; } } while (0)

static int run_case(int64_t R, int T, int reps, bool check) {
  using namespace probe;
  int dev = 0, sms = 0;
  CK(cudaGetDevice(&dev));
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  std::vector<float> hA((size_t)R * P), hV((size_t)T * R * Q);
  uint32_t seed = 12345u + (uint32_t)R;
  auto rnd = [&]() { seed = seed * 1664525u + 1013904223u; return ((seed >> 8) & 0xFFFF) / 32768.0f - 1.0f; };
  for (auto& x : hA) x = rnd();
  for (auto& x : hV) x = rnd();
  float *dA, *dV, *dPart, *dW;
  CK(cudaMalloc(&dA, hA.size() * 4));
  CK(cudaMalloc(&dV, hV.size() * 4));
  CK(cudaMemcpy(dA, hA.data(), hA.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dV, hV.data(), hV.size() * 4, cudaMemcpyHostToDevice));
  int grid = sms;
  int64_t rpc = ((R + grid - 1) / grid + BK - 1) / BK * BK;
  if (rpc < BK) rpc = BK;
  grid = (int)((R + rpc - 1) / rpc);
  const size_t n = (size_t)T * P * Q;
  CK(cudaMalloc(&dPart, (size_t)grid * n * 4));
  CK(cudaMalloc(&dW, n * 4));
  Params prm;
  prm.A = dA; prm.a_ld = P; prm.R = R; prm.rows_per_cta = rpc; prm.T = T; prm.partial = dPart;
  for (int t = 0; t < MAX_T; ++t) { prm.V[t] = dV + (size_t)(t < T ? t : 0) * R * Q; prm.v_ld[t] = Q; }
  const size_t smem = (size_t)STAGES * STAGE_BYTES + 1024 + 64;
  CK(cudaFuncSetAttribute(tapgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int i = 0; i < reps + 1; ++i) {
    CK(cudaEventRecord(e0));
    tapgrad_tc_kernel<<<grid, THREADS, smem>>>(prm);
    reduce_kernel<<<64, 256>>>(dPart, dW, grid, (int64_t)n);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    CK(cudaGetLastError());
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (i > 0 && ms < best) best = ms;
  }
  printf("R=%lld T=%d grid=%d rows/cta=%lld smem=%zu: %.3f ms (kernel + reduce, best of %d)\n", (long long)R, T, grid,
         (long long)rpc, smem, best, reps);
  int rc = 0;
  if (check) {
    std::vector<float> hW(n);
    CK(cudaMemcpy(hW.data(), dW, n * 4, cudaMemcpyDeviceToHost));
    double maxerr = 0, maxref = 0;
    for (int t = 0; t < T; ++t)
      for (int p = 0; p < P; ++p)
        for (int q = 0; q < Q; ++q) {
          double acc = 0;
          for (int64_t r = 0; r < R; ++r) acc += (double)hA[(size_t)r * P + p] * (double)hV[((size_t)t * R + r) * Q + q];
          maxerr = fmax(maxerr, fabs(acc - hW[((size_t)t * P + p) * Q + q]));
          maxref = fmax(maxref, fabs(acc));
        }
    printf("   max |err| / max |ref| = %.3e  (%s, 3xTF32 should be ~1e-6)\n", maxerr / maxref, maxerr / maxref < 1e-5 ? "OK" : "FAIL");
    rc = maxerr / maxref < 1e-5 ? 0 : 2;
  }
  cudaFree(dA); cudaFree(dV); cudaFree(dPart); cudaFree(dW);
  return rc;
}

int main(int argc, char** argv) {
  int rc = 0;
  rc |= run_case(4096 + 17, 5, 1, true);     // ragged tail, odd T (zero half of the last pair)
  rc |= run_case(777, 2, 1, true);           // fewer rows than CTAs x 32: idle CTAs write zero partials
  rc |= run_case(150000, 6, 1, true);
  if (argc > 1) rc |= run_case(1000000, 5, 5, false);   // headline size, timing only
  printf(rc ? "PROBE FAILED\n" : "PROBE OK\n");
  return rc;
}
