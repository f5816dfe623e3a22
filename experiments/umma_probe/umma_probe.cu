// umma_probe — stand-alone probe of the tcgen05 conventions the round-2 MLP tile kernel depends on (sm_100a).
//
// STATUS: compiles for sm_100a (SASS contains UTCMMA / UTCBAR / LDTM / STTM); NOT YET EXECUTED on a GPU — round 1 ran
// out of GPU minutes.  It is not part of libtensoir_b200.so and nothing in the product imports it.
//
//   make -C experiments/umma_probe && timeout 60 experiments/umma_probe/umma_probe
//
// What it checks / measures
//   1. SS form: D[128x128] (fp32, TMEM) = A[128xK] * B[128xK]^T with bf16 operands written BY THREADS into shared memory
//      in the no-swizzle K-major canonical layout (8 rows x 16 B core matrices), compared with a CPU product.  This is
//      the layout the fused MLP would use for its resident split-bf16 weights.  `--swap` exchanges the roles of the
//      leading / stride byte offsets in the descriptor in case the convention is the other way round.
//   2. TS form: the same product with A read from TMEM (written with tcgen05.st, two bf16 per 32-bit column) — the way
//      hidden activations would be handed from one layer's epilogue to the next layer's MMA without touching smem.
//   3. Throughput: cycles per 128x128x16 MMA for the 3-term error-compensated product (hi*hi + hi*lo + lo*hi), one and
//      two accumulators in flight, and cycles for a 4-warp tcgen05.ld drain of a 128x128 fp32 accumulator.
// Every mbarrier wait is bounded: a wrong descriptor makes the probe report TIMEOUT instead of hanging the GPU.
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace {

constexpr int M = 128;   // rows of A = TMEM lanes
constexpr int N = 128;   // rows of B = accumulator columns

#define CUDA_OK(x)                                                                          \
  do {                                                                                      \
    cudaError_t e_ = (x);                                                                   \
    if (e_ != cudaSuccess) {                                                                \
      std::printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);   \
      std::exit(2);                                                                         \
    }                                                                                       \
  } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// Byte offset of element (r, k) of an [R x K] bf16 operand: K-chunk-major, each chunk = R rows x 16 B, so a warp whose
// lanes own consecutive rows writes 512 contiguous bytes (no bank conflicts) and
//   stride between 8-row groups (SBO) = 128 B, stride between K chunks (LBO) = R * 16 B.
__host__ __device__ inline uint32_t operand_offset(int r, int k, int R) {
  return (uint32_t)(k >> 3) * (uint32_t)(R * 16) + (uint32_t)r * 16u + (uint32_t)(k & 7) * 2u;
}

// Shared-memory matrix descriptor, SWIZZLE_NONE, sm_100 version bits (bit 46).
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFFu);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}

// kind::f16 instruction descriptor: D fp32, A/B bf16, both K-major, dense.
__host__ __device__ constexpr uint32_t make_idesc(int m, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

__device__ __forceinline__ void mma_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void bar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
// bounded wait; returns false on timeout
__device__ __forceinline__ bool bar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t a = smem_u32(bar);
  for (int spin = 0; spin < (1 << 22); ++spin) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(done)
        : "r"(a), "r"(parity)
        : "memory");
    if (done) return true;
  }
  return false;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void proxy_fence() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

template <int COLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* slot) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(COLS)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_free(uint32_t base) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(COLS) : "memory");
}

// 32 lanes x 32 consecutive fp32 columns: thread `lane` of warp w receives row 32*(w%4)+lane.
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// 32 lanes x 8 columns store (8 x 32-bit = 16 bf16 of one row = one K=16 MMA slice of a TMEM-resident A operand)
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};\n" ::"r"(taddr), "r"(v[0]),
               "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------------------------------
// 1 + 2: correctness.  mode 0 = A from shared memory, mode 1 = A from TMEM.
__global__ void __launch_bounds__(128) umma_check(const __nv_bfloat16* __restrict__ A, const __nv_bfloat16* __restrict__ B,
                                                  float* __restrict__ D, int K, int swap, int mode, int* status) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  uint8_t* sA = smem;
  uint8_t* sB = smem + (size_t)M * K * 2;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  for (int i = tid; i < M * K; i += blockDim.x) {
    const int r = i / K, k = i % K;
    *reinterpret_cast<__nv_bfloat16*>(sA + operand_offset(r, k, M)) = A[i];
    *reinterpret_cast<__nv_bfloat16*>(sB + operand_offset(r, k, N)) = B[i];
  }
  proxy_fence();   // generic-proxy writes -> visible to the tensor core's async proxy
  if (warp == 0) tmem_alloc<512>(&tmem_slot);
  if (tid == 0) {
    bar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t tmem_d = tmem;            // columns [0,128): accumulator
  const uint32_t tmem_a = tmem + 256;      // columns [256, 256 + K/2): A operand for the TS form

  if (mode == 1) {
    // row = 32*warp + lane; pack consecutive-K pairs into 32-bit columns
    const int row = warp * 32 + lane;
    for (int ks = 0; ks < K / 16; ++ks) {
      uint32_t v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const __nv_bfloat16 lo = A[row * K + ks * 16 + 2 * j], hi = A[row * K + ks * 16 + 2 * j + 1];
        v[j] = (uint32_t)__bfloat16_as_ushort(lo) | ((uint32_t)__bfloat16_as_ushort(hi) << 16);
      }
      tmem_st8(tmem_a + ((uint32_t)(warp * 32) << 16) + ks * 8, v);
    }
    tmem_st_wait();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  }

  if (tid == 0) {
    const uint32_t idesc = make_idesc(M, N);
    const uint32_t lbo = swap ? 128u : (uint32_t)(M * 16), sbo = swap ? (uint32_t)(M * 16) : 128u;
    for (int ks = 0; ks < K / 16; ++ks) {
      const uint32_t koff = (uint32_t)ks * 2u * (uint32_t)(M * 16);   // two K chunks per MMA
      const uint64_t bdesc = make_desc(smem_u32(sB) + koff, lbo, sbo);
      if (mode == 0) {
        const uint64_t adesc = make_desc(smem_u32(sA) + koff, lbo, sbo);
        mma_ss(tmem_d, adesc, bdesc, idesc, ks > 0);
      } else {
        mma_ts(tmem_d, tmem_a + ks * 8, bdesc, idesc, ks > 0);
      }
    }
    mma_commit(&bar);
  }
  const bool ok = bar_wait(&bar, 0);
  tc_fence_after();
  if (!ok) {
    if (tid == 0) *status = 1;
  } else {
    const int row = warp * 32 + lane;
    for (int c = 0; c < N / 32; ++c) {
      uint32_t v[32];
      tmem_ld32(tmem_d + ((uint32_t)(warp * 32) << 16) + c * 32, v);
#pragma unroll
      for (int j = 0; j < 32; ++j) D[row * N + c * 32 + j] = __uint_as_float(v[j]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_free<512>(tmem);
}

// ---------------------------------------------------------------------------------------------------------------------
// 3: throughput of the 3-term split product for one layer-0 sized tile (K = 160 -> 10 k-steps x 3 MMAs), `accs`
// accumulators in flight (1 or 2), and the cost of draining one accumulator with 4 warps.
__global__ void __launch_bounds__(128) umma_rate(int iters, int accs, long long* cycles_mma, long long* cycles_ld,
                                                 float* sink, int* status) {
  constexpr int K = 160;
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar[2];
  __shared__ uint32_t tmem_slot;
  // hi / lo planes of A and B: 4 x 128 x 160 x 2 B = 160 KB
  uint8_t* plane[4];
  for (int i = 0; i < 4; ++i) plane[i] = smem + (size_t)i * M * K * 2;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < 4 * M * K / 2; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;   // small bf16
  proxy_fence();
  if (warp == 0) tmem_alloc<512>(&tmem_slot);
  if (tid == 0) {
    bar_init(&bar[0], 1);
    bar_init(&bar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  bool ok = true;
  long long t0 = 0, t1 = 0;
  if (tid == 0) {
    const uint32_t idesc = make_idesc(M, N);
    const uint32_t lbo = M * 16, sbo = 128;
    t0 = clock64();
    uint32_t phase[2] = {0, 0};
    for (int it = 0; it < iters; ++it) {
      const int a = it % accs;
      if (it >= accs) {   // accumulator `a` is about to be overwritten: its previous product must have completed
        ok = ok && bar_wait(&bar[a], phase[a]);
        phase[a] ^= 1;
      }
      const uint32_t d = tmem + a * 128;
      int first = 1;
      for (int ks = 0; ks < K / 16; ++ks) {
        const uint32_t koff = (uint32_t)ks * 2u * lbo;
        const uint64_t ah = make_desc(smem_u32(plane[0]) + koff, lbo, sbo), al = make_desc(smem_u32(plane[1]) + koff, lbo, sbo);
        const uint64_t bh = make_desc(smem_u32(plane[2]) + koff, lbo, sbo), bl = make_desc(smem_u32(plane[3]) + koff, lbo, sbo);
        mma_ss(d, ah, bh, idesc, !first);
        mma_ss(d, ah, bl, idesc, 1);
        mma_ss(d, al, bh, idesc, 1);
        first = 0;
      }
      mma_commit(&bar[a]);
    }
    for (int a = 0; a < accs && a < iters; ++a) ok = ok && bar_wait(&bar[a], phase[a]);
    t1 = clock64();
    if (blockIdx.x == 0) *cycles_mma = t1 - t0;
    if (!ok) *status = 1;
  }
  __syncthreads();
  tc_fence_after();
  // drain timing: 4 warps read the 128 x 128 accumulator
  const long long l0 = clock64();
  float s = 0.f;
  for (int c = 0; c < N / 32; ++c) {
    uint32_t v[32];
    tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c * 32, v);
#pragma unroll
    for (int j = 0; j < 32; ++j) s += __uint_as_float(v[j]);
  }
  const long long l1 = clock64();
  if (blockIdx.x == 0 && tid == 0) *cycles_ld = l1 - l0;
  if (s == 123.456f) sink[0] = s;
  (void)lane;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_free<512>(tmem);
}

float bf16_round(float x) { return __bfloat162float(__float2bfloat16(x)); }

int run_check(int K, int swap, int mode) {
  std::vector<float> a(M * K), b(N * K);
  std::vector<__nv_bfloat16> ha(M * K), hb(N * K);
  uint32_t rng = 12345u;
  auto next = [&]() {
    rng = rng * 1664525u + 1013904223u;
    return ((rng >> 8) & 0xFFFF) / 65536.f - 0.5f;
  };
  for (int i = 0; i < M * K; ++i) { a[i] = bf16_round(next()); ha[i] = __float2bfloat16(a[i]); }
  for (int i = 0; i < N * K; ++i) { b[i] = bf16_round(next()); hb[i] = __float2bfloat16(b[i]); }
  __nv_bfloat16 *dA, *dB;
  float* dD;
  int* dS;
  CUDA_OK(cudaMalloc(&dA, ha.size() * 2));
  CUDA_OK(cudaMalloc(&dB, hb.size() * 2));
  CUDA_OK(cudaMalloc(&dD, M * N * 4));
  CUDA_OK(cudaMalloc(&dS, 4));
  CUDA_OK(cudaMemcpy(dA, ha.data(), ha.size() * 2, cudaMemcpyHostToDevice));
  CUDA_OK(cudaMemcpy(dB, hb.data(), hb.size() * 2, cudaMemcpyHostToDevice));
  CUDA_OK(cudaMemset(dD, 0, M * N * 4));
  CUDA_OK(cudaMemset(dS, 0, 4));
  const size_t smem = (size_t)(M + N) * K * 2;
  CUDA_OK(cudaFuncSetAttribute(umma_check, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  umma_check<<<1, 128, smem>>>(dA, dB, dD, K, swap, mode, dS);
  CUDA_OK(cudaDeviceSynchronize());
  std::vector<float> d(M * N);
  int st = 0;
  CUDA_OK(cudaMemcpy(d.data(), dD, M * N * 4, cudaMemcpyDeviceToHost));
  CUDA_OK(cudaMemcpy(&st, dS, 4, cudaMemcpyDeviceToHost));
  double worst = 0.0;
  for (int i = 0; i < M; ++i)
    for (int j = 0; j < N; ++j) {
      double ref = 0.0;
      for (int k = 0; k < K; ++k) ref += (double)a[i * K + k] * (double)b[j * K + k];
      worst = std::fmax(worst, std::fabs(ref - (double)d[i * N + j]));
    }
  const bool pass = st == 0 && worst < 1e-3;
  std::printf("check  %s  K=%3d swap=%d  max|err| = %.3e  %s\n", mode ? "A in TMEM" : "A in smem", K, swap, worst,
              st ? "TIMEOUT" : (pass ? "PASS" : "FAIL"));
  cudaFree(dA); cudaFree(dB); cudaFree(dD); cudaFree(dS);
  return pass ? 0 : 1;
}

void run_rate(int accs) {
  long long *dC, *dL;
  float* dSink;
  int* dS;
  CUDA_OK(cudaMalloc(&dC, 8));
  CUDA_OK(cudaMalloc(&dL, 8));
  CUDA_OK(cudaMalloc(&dSink, 4));
  CUDA_OK(cudaMalloc(&dS, 4));
  CUDA_OK(cudaMemset(dS, 0, 4));
  const size_t smem = (size_t)4 * M * 160 * 2;
  CUDA_OK(cudaFuncSetAttribute(umma_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int iters = 200, ctas = 148;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  umma_rate<<<ctas, 128, smem>>>(8, accs, dC, dL, dSink, dS);   // warm-up
  cudaEventRecord(e0);
  umma_rate<<<ctas, 128, smem>>>(iters, accs, dC, dL, dSink, dS);
  cudaEventRecord(e1);
  CUDA_OK(cudaDeviceSynchronize());
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  long long cyc = 0, ld = 0;
  int st = 0;
  CUDA_OK(cudaMemcpy(&cyc, dC, 8, cudaMemcpyDeviceToHost));
  CUDA_OK(cudaMemcpy(&ld, dL, 8, cudaMemcpyDeviceToHost));
  CUDA_OK(cudaMemcpy(&st, dS, 4, cudaMemcpyDeviceToHost));
  const double mmas = (double)iters * 30.0;
  const double flops = mmas * 2.0 * 128 * 128 * 16 * ctas;
  std::printf("rate   accumulators=%d  %s  %.1f cycles per 128x128x16 MMA, drain of one accumulator %lld cycles, "
              "%.1f TFLOP/s bf16 over %d CTAs (kernel %.3f ms incl. setup)\n",
              accs, st ? "TIMEOUT" : "ok", (double)cyc / mmas, ld, flops / (ms * 1e-3) / 1e12, ctas, ms);
  cudaFree(dC); cudaFree(dL); cudaFree(dSink); cudaFree(dS);
}

}  // namespace

int main(int argc, char** argv) {
  int swap = 0;
  for (int i = 1; i < argc; ++i)
    if (!std::strcmp(argv[i], "--swap")) swap = 1;
  cudaDeviceProp prop;
  CUDA_OK(cudaGetDeviceProperties(&prop, 0));
  std::printf("device %s sm_%d%d, %d SMs\n", prop.name, prop.major, prop.minor, prop.multiProcessorCount);
  int bad = 0;
  bad += run_check(16, swap, 0);
  bad += run_check(64, swap, 0);
  bad += run_check(160, swap, 0);
  bad += run_check(64, swap, 1);
  if (!bad) {
    run_rate(1);
    run_rate(2);
  }
  std::printf(bad ? "PROBE FAILED\n" : "PROBE OK\n");
  return bad ? 1 : 0;
}
