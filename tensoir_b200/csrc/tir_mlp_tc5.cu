// Blackwell-native appearance MLP (sm_100a): the secondary-ray appearance head of compute_radiance
// (models/relight_utils.py:803-834) = compute_appfeature (tensoRF_rotated_lights.py:197-224) + MLPRender_Fea
// (tensorBase_rotated_lights.py:122-146) on the compacted appearance-sample list, with basis_mat and the two 128-wide
// layers on the 5th-generation tensor cores:
//
//   * tcgen05.mma (kind::f16, M = 128 samples x N = 128 units x K = 16) issued by ONE thread; accumulator in TMEM;
//   * the A operand (MLP input / hidden activations, error-compensated split BF16: hi + lo) lives in TMEM as well
//     (".ts" form): the epilogue threads write the next layer's A with tcgen05.st straight from their registers, so
//     activations never touch shared memory; B operands (split-BF16 weights, 147 KB) stay resident in shared memory in
//     the canonical no-swizzle K-major layout for the whole persistent CTA;
//   * every product is hi*hi + hi*lo + lo*hi accumulated in fp32 (the dropped lo*lo term is 2^-16 relative);
//   * two warpgroups of 128 threads (thread = sample row = TMEM lane) ping-pong through ONE issuer thread: while one
//     group's MMAs run or its accumulator drains, the other gathers (108 x 256-bit loads per sample; the products
//     plane*line*light go straight into TMEM as the A operand of the basis_mat GEMM) — the L2-latency-bound gather
//     overlaps the tensor work;
//   * three MMA stages per 128-sample tile: basis_mat (K = 144, N = 32), layer 0 (K = 160, N = 128), layer 1 (K = 128);
//   * TMEM map (512 columns): P0 [0,160) / P1 [160,320) = A operand of group 0 / 1 (hi | lo halves), Q [320,448) = the
//     shared fp32 accumulator, handed from group to group with mbarriers.
//   The last layer (128 -> 3/4, < 1 % of the FLOPs) runs in fp32 on the CUDA cores.
//
// The descriptor / TMEM-operand conventions are the ones experiments/umma_probe validated on a B200
// (gpurun_out/r2_c1/umma_probe.txt: SS and TS forms exact to 1e-6, 77 cycles per 128x128x16 MMA).
// Every mbarrier wait is bounded: a protocol error raises p.error instead of hanging the device.
#include <cuda_bf16.h>
#include "tir_device.cuh"
#include "tir_internal.h"

using namespace tir;

namespace {

constexpr int AC = 48;
constexpr int K0 = 3 * AC;        // 144
constexpr int F = 27;
constexpr int HID = 128;
constexpr int IN = 150;
constexpr int K1 = 160;           // IN padded to k16
constexpr int ROWS = 128;         // samples per tile = TMEM lanes
constexpr int NCONS = 256;        // two consumer warpgroups
constexpr int NTHREADS = NCONS + 32;   // + the issuer warp
constexpr uint32_t P_COLS = 160, Q_OFF = 320;

struct Smem5 {
  // B operands, canonical K-major no-swizzle layout: byte offset(n, k) = (k >> 3) * (128 * 16) + n * 16 + (k & 7) * 2
  __align__(1024) uint8_t w0h[HID * K1 * 2];
  uint8_t w0l[HID * K1 * 2];
  uint8_t w1h[HID * HID * 2];
  uint8_t w1l[HID * HID * 2];
  uint8_t bsh[32 * K0 * 2];       // basis_mat as a B operand: [n = feature (27 -> 32)][k = product channel], 32-row atoms
  uint8_t bsl[32 * K0 * 2];
  float w2[4 * HID];
  float b0[HID], b1[HID], b2[4];
  float lmean[K0];
  unsigned long long bar_a[2];    // consumer group g -> issuer: A operand of the next layer is in P_g
  unsigned long long bar_d[2];    // issuer -> group g: the layer's MMAs have completed (tcgen05.commit)
  unsigned long long bar_q;       // draining group -> issuer: Q is free again
  uint32_t tmem_base;
};

struct Tc5Params {
  TirField f;
  TirMlp mlp;
  // sample-list mode
  const TirAppSample* samples;
  const uint32_t* sample_count;
  int64_t max_samples;
  const float* ray_dirs;
  int n_dirs;
  const int32_t* light_idx;
  float* rgb_out;
  // points mode
  const float* pts_xn;
  const float* pts_x;
  int64_t n_points;
  float* out;
  int act;
  int light_mode;     // 0 none, 1 indexed row, 2 mean row
  int* error;         // device flag: set when a bounded wait timed out
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t operand_offset(int r, int k) {
  return (uint32_t)(k >> 3) * (uint32_t)(ROWS * 16) + (uint32_t)r * 16u + (uint32_t)(k & 7) * 2u;
}
// SWIZZLE_NONE descriptor: LBO = bytes between K chunks (= operand rows * 16), SBO = 128 B between 8-row groups
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo = ROWS * 16) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFFu);
  d |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((128u >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
__host__ __device__ constexpr uint32_t make_idesc(int m, int n) {     // kind::f16: D fp32, A/B bf16, K-major, dense
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void mma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void mma_commit(unsigned long long* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void bar_init(unsigned long long* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void bar_arrive(unsigned long long* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// bounded wait; on timeout raises the error flag and returns false
__device__ __forceinline__ bool bar_wait(unsigned long long* bar, uint32_t parity, int* error) {
  const uint32_t a = smem_u32(bar);
  for (int spin = 0; spin < (1 << 24); ++spin) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(done)
        : "r"(a), "r"(parity)
        : "memory");
    if (done) return true;
    if ((spin & 1023) == 1023 && *reinterpret_cast<volatile int*>(error) != 0) return false;
  }
  atomicExch(error, 1);
  return false;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void proxy_fence() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

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
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};\n" ::"r"(taddr), "r"(v[0]),
               "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// 256-bit read-only load (sm_100: LDG.E.256): one full 32-byte sector per lane and request.  The 128-bit form fetched
// every sector twice (two requests per sector, second one usually an L1 miss: 88 KB of L1 vs ~200 KB of loads in flight),
// which put the kernel on the L2 bandwidth wall at 2x the algorithmic bytes.
struct F8 { float4 a, b; };
__device__ __forceinline__ F8 ldg8(const float* p) {
  F8 r;
  asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(r.a.x), "=f"(r.a.y), "=f"(r.a.z), "=f"(r.a.w), "=f"(r.b.x), "=f"(r.b.y), "=f"(r.b.z), "=f"(r.b.w)
               : "l"(p));
  return r;
}

// fp32 pair -> packed split-BF16 words: hi = (bf16(a) | bf16(b) << 16), lo = the residuals
__device__ __forceinline__ void split_pack(float a, float b, uint32_t& hi, uint32_t& lo) {
  const __nv_bfloat16 ah = __float2bfloat16_rn(a), bh = __float2bfloat16_rn(b);
  const __nv_bfloat16 al = __float2bfloat16_rn(a - __bfloat162float(ah)), bl = __float2bfloat16_rn(b - __bfloat162float(bh));
  hi = (uint32_t)__bfloat16_as_ushort(ah) | ((uint32_t)__bfloat16_as_ushort(bh) << 16);
  lo = (uint32_t)__bfloat16_as_ushort(al) | ((uint32_t)__bfloat16_as_ushort(bl) << 16);
}

// element `c` of the 160-wide MLP input row [feat 27 | x 3 | sin PE(feat) 54 | cos 54 | sin PE(x) 6 | cos 6 | 0 x 10]
// (tensorBase:12-17, :136-142) from the cached sin / cos of the base angles; frequency 2 via the double-angle
// identities (agrees with sin(2x) / cos(2x) to 1 ulp-level rounding, far inside the 5e-5 parity bar).
// Written so that, fully unrolled with a compile-time `c`, every array index is a constant.
#define TC5_INPUT_ELEM(c, feat, xv, sf, cf, sx, cx)                                                         \
  ((c) < F ? feat[(c) < F ? (c) : 0]                                                                        \
   : (c) < 30 ? xv[(c) < 30 && (c) >= F ? (c) - F : 0]                                                      \
   : (c) < 84 ? ((((c) - 30) & 1) ? 2.f * sf[((c) - 30) >> 1 < F ? ((c) - 30) >> 1 : 0] * cf[((c) - 30) >> 1 < F ? ((c) - 30) >> 1 : 0] \
                                  : sf[((c) - 30) >> 1 < F ? ((c) - 30) >> 1 : 0])                           \
   : (c) < 138 ? ((((c) - 84) & 1) ? 1.f - 2.f * sf[((c) - 84) >> 1 < F ? ((c) - 84) >> 1 : 0] * sf[((c) - 84) >> 1 < F ? ((c) - 84) >> 1 : 0] \
                                   : cf[((c) - 84) >> 1 < F ? ((c) - 84) >> 1 : 0])                          \
   : (c) < 144 ? ((((c) - 138) & 1) ? 2.f * sx[((c) - 138) >> 1 < 3 ? ((c) - 138) >> 1 : 0] * cx[((c) - 138) >> 1 < 3 ? ((c) - 138) >> 1 : 0] \
                                    : sx[((c) - 138) >> 1 < 3 ? ((c) - 138) >> 1 : 0])                       \
   : (c) < 150 ? ((((c) - 144) & 1) ? 1.f - 2.f * sx[((c) - 144) >> 1 < 3 ? ((c) - 144) >> 1 : 0] * sx[((c) - 144) >> 1 < 3 ? ((c) - 144) >> 1 : 0] \
                                    : cx[((c) - 144) >> 1 < 3 ? ((c) - 144) >> 1 : 0])                       \
               : 0.f)

template <bool POINTS>
__global__ void __launch_bounds__(NTHREADS, 1) app_mlp_tc5_kernel(const Tc5Params p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  Smem5& s = *reinterpret_cast<Smem5*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const TirMlp& mlp = p.mlp;
  const int out_dim = mlp.out_dim;

  // ---- stage the split-BF16 weights in the tensor core's K-major core-matrix layout (once per persistent CTA)
  for (int i = tid; i < HID * K1; i += NTHREADS) {
    const int n = i / K1, k = i % K1;
    const float v = k < IN ? __ldg(mlp.w0 + n * IN + k) : 0.f;
    const __nv_bfloat16 h = __float2bfloat16_rn(v), l = __float2bfloat16_rn(v - __bfloat162float(h));
    *reinterpret_cast<__nv_bfloat16*>(s.w0h + operand_offset(n, k)) = h;
    *reinterpret_cast<__nv_bfloat16*>(s.w0l + operand_offset(n, k)) = l;
  }
  for (int i = tid; i < HID * HID; i += NTHREADS) {
    const int n = i / HID, k = i % HID;
    const float v = __ldg(mlp.w1 + i);
    const __nv_bfloat16 h = __float2bfloat16_rn(v), l = __float2bfloat16_rn(v - __bfloat162float(h));
    *reinterpret_cast<__nv_bfloat16*>(s.w1h + operand_offset(n, k)) = h;
    *reinterpret_cast<__nv_bfloat16*>(s.w1l + operand_offset(n, k)) = l;
  }
  for (int i = tid; i < 32 * K0; i += NTHREADS) {
    const int n = i / K0, k = i % K0;
    const float v = n < F ? __ldg(mlp.basis + n * K0 + k) : 0.f;
    const __nv_bfloat16 h = __float2bfloat16_rn(v), l = __float2bfloat16_rn(v - __bfloat162float(h));
    const uint32_t off = (uint32_t)(k >> 3) * (32u * 16u) + (uint32_t)n * 16u + (uint32_t)(k & 7) * 2u;
    *reinterpret_cast<__nv_bfloat16*>(s.bsh + off) = h;
    *reinterpret_cast<__nv_bfloat16*>(s.bsl + off) = l;
  }
  for (int i = tid; i < 4 * HID; i += NTHREADS) s.w2[i] = (i / HID) < out_dim ? __ldg(mlp.w2 + i) : 0.f;
  for (int i = tid; i < HID; i += NTHREADS) { s.b0[i] = __ldg(mlp.b0 + i); s.b1[i] = __ldg(mlp.b1 + i); }
  if (tid < 4) s.b2[tid] = tid < out_dim ? __ldg(mlp.b2 + tid) : 0.f;
  if (p.light_mode == 2)
    for (int c = tid; c < K0; c += NTHREADS) {
      float a = 0.f;
      for (int l = 0; l < mlp.n_lights; ++l) a += __ldg(mlp.light_line + (size_t)l * K0 + c);
      s.lmean[c] = a / (float)mlp.n_lights;
    }
  proxy_fence();   // generic-proxy writes of the B operands -> visible to the tensor core's async proxy
  if (warp == NCONS / 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s.tmem_base)),
                 "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    bar_init(&s.bar_a[0], ROWS); bar_init(&s.bar_a[1], ROWS);
    bar_init(&s.bar_d[0], 1); bar_init(&s.bar_d[1], 1);
    bar_init(&s.bar_q, ROWS);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s.tmem_base;

  const int64_t total = POINTS ? p.n_points
                               : (int64_t)min((unsigned long long)*p.sample_count, (unsigned long long)p.max_samples);
  const int64_t n_tiles = (total + ROWS - 1) / ROWS;
  const int64_t n_pairs = (n_tiles + 1) / 2;

  if (warp == NCONS / 32) {
    // =================================================== issuer ===================================================
    if (lane == 0) {
      const uint32_t idesc = make_idesc(ROWS, HID), idesc_b = make_idesc(ROWS, 32);
      uint32_t ph_a[2] = {0, 0}, ph_q = 1;     // a fresh mbarrier passes a wait on parity 1: Q starts out free
      bool ok = true;
      for (int64_t pair = blockIdx.x; pair < n_pairs && ok; pair += gridDim.x) {
        // stage 0: basis_mat (K = 144, N = 32), 1: layer 0 (K = 160, N = 128), 2: layer 1 (K = 128, N = 128)
        for (int stage = 0; stage < 3 && ok; ++stage)
          for (int g = 0; g < 2 && ok; ++g) {
            ok = ok && bar_wait(&s.bar_a[g], ph_a[g], p.error);
            ph_a[g] ^= 1;
            ok = ok && bar_wait(&s.bar_q, ph_q, p.error);
            ph_q ^= 1;
            if (!ok) break;
            tc_fence_after();
            const uint32_t pa = tmem + (uint32_t)g * P_COLS;
            const int kdim = stage == 0 ? K0 : (stage == 1 ? K1 : HID);
            const uint32_t lo_off = kdim / 2;                                // hi half | lo half of the A operand
            const uint8_t* bh = stage == 0 ? s.bsh : (stage == 1 ? s.w0h : s.w1h);
            const uint8_t* bl = stage == 0 ? s.bsl : (stage == 1 ? s.w0l : s.w1l);
            const uint32_t lbo = stage == 0 ? 32u * 16u : (uint32_t)(ROWS * 16);
            const uint32_t id = stage == 0 ? idesc_b : idesc;
            for (int ks = 0; ks < kdim / 16; ++ks) {
              const uint32_t koff = (uint32_t)ks * 2u * lbo;                 // two 8-element K chunks per MMA
              const uint64_t dh = make_desc(smem_u32(bh) + koff, lbo), dl = make_desc(smem_u32(bl) + koff, lbo);
              mma_ts(tmem + Q_OFF, pa + lo_off + ks * 8, dh, id, ks > 0);    // small terms first
              mma_ts(tmem + Q_OFF, pa + ks * 8, dl, id, 1);
              mma_ts(tmem + Q_OFF, pa + ks * 8, dh, id, 1);
            }
            mma_commit(&s.bar_d[g]);
          }
      }
    }
  } else {
    // ================================================== consumers =================================================
    const int g = warp >> 2;                    // warpgroup
    const int r = tid & (ROWS - 1);             // sample row = TMEM lane
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t tp = tmem + lane_base + (uint32_t)g * P_COLS;
    const uint32_t tq = tmem + lane_base + Q_OFF;
    uint32_t ph_d = 0;
    bool ok = true;
    for (int64_t pair = blockIdx.x; pair < n_pairs; pair += gridDim.x) {
      const int64_t i = (pair * 2 + g) * ROWS + r;
      const bool live = i < total;
      // ---- sample
      float xn[3] = {0.f, 0.f, 0.f}, xv[3] = {0.f, 0.f, 0.f}, wgt = 0.f;
      int ray = 0, li = 0;
      if (live) {
        if (POINTS) {
          xn[0] = p.pts_xn[i * 3]; xn[1] = p.pts_xn[i * 3 + 1]; xn[2] = p.pts_xn[i * 3 + 2];
          xv[0] = p.pts_x[i * 3]; xv[1] = p.pts_x[i * 3 + 1]; xv[2] = p.pts_x[i * 3 + 2];
          li = p.light_idx ? p.light_idx[i] : 0;
        } else {
          const TirAppSample sm = p.samples[i];
          xn[0] = sm.xn[0]; xn[1] = sm.xn[1]; xn[2] = sm.xn[2]; wgt = sm.weight; ray = sm.ray;
          const int64_t di = p.n_dirs > 0 ? (int64_t)(ray % p.n_dirs) : (int64_t)ray;
          xv[0] = __ldg(p.ray_dirs + di * 3); xv[1] = __ldg(p.ray_dirs + di * 3 + 1); xv[2] = __ldg(p.ray_dirs + di * 3 + 2);
          li = p.light_idx ? __ldg(p.light_idx + (p.n_dirs > 0 ? ray / p.n_dirs : ray)) : 0;
        }
      }
      // ---- gather: products (plane * line * light) go straight into P_g as the split-BF16 A operand of the basis GEMM
#pragma unroll 1
      for (int k = 0; k < 3; ++k) {
        const int m0 = kMat0[k], m1 = kMat1[k], v = kVec[k];
        const Bilinear b = bilinear_setup(xn[m0], xn[m1], p.f.grid[m0], p.f.grid[m1]);
        const Linear1 l = linear_setup(xn[v], p.f.grid[v]);
        const float* P00 = p.f.aplane[k] + (size_t)b.o00 * AC;
        const float* P01 = p.f.aplane[k] + (size_t)b.o01 * AC;
        const float* P10 = p.f.aplane[k] + (size_t)b.o10 * AC;
        const float* P11 = p.f.aplane[k] + (size_t)b.o11 * AC;
        const float* L0 = p.f.aline[k] + (size_t)l.o0 * AC;
        const float* L1 = p.f.aline[k] + (size_t)l.o1 * AC;
        const float* lrow = p.light_mode == 1 ? (mlp.light_line + (size_t)li * K0) : nullptr;
#pragma unroll 1
        for (int ks = 0; ks < AC / 16; ++ks) {        // 16 channels = one K step = 24 independent LDG.128
          uint32_t hi[8], lo[8];
#pragma unroll
          for (int h8 = 0; h8 < 2; ++h8) {
            const int c = ks * 16 + h8 * 8;
            float x[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            if (live) {
              const F8 t00 = ldg8(P00 + c), t01 = ldg8(P01 + c), t10 = ldg8(P10 + c), t11 = ldg8(P11 + c);
              const F8 u0 = ldg8(L0 + c), u1 = ldg8(L1 + c);
              const float4 pa = bilerp4(t00.a, t01.a, t10.a, t11.a, b), pb = bilerp4(t00.b, t01.b, t10.b, t11.b, b);
              const float4 la = lerp4(u0.a, u1.a, l), lb = lerp4(u0.b, u1.b, l);
              x[0] = __fmul_rn(pa.x, la.x); x[1] = __fmul_rn(pa.y, la.y); x[2] = __fmul_rn(pa.z, la.z); x[3] = __fmul_rn(pa.w, la.w);
              x[4] = __fmul_rn(pb.x, lb.x); x[5] = __fmul_rn(pb.y, lb.y); x[6] = __fmul_rn(pb.z, lb.z); x[7] = __fmul_rn(pb.w, lb.w);
              const int col = k * AC + c;
              if (p.light_mode == 1) {            // (plane * line) * light  (tensoRF_rotated_lights.py:222)
                const F8 lc = ldg8(lrow + col);
                x[0] = __fmul_rn(x[0], lc.a.x); x[1] = __fmul_rn(x[1], lc.a.y); x[2] = __fmul_rn(x[2], lc.a.z); x[3] = __fmul_rn(x[3], lc.a.w);
                x[4] = __fmul_rn(x[4], lc.b.x); x[5] = __fmul_rn(x[5], lc.b.y); x[6] = __fmul_rn(x[6], lc.b.z); x[7] = __fmul_rn(x[7], lc.b.w);
              } else if (p.light_mode == 2) {
#pragma unroll
                for (int e = 0; e < 8; ++e) x[e] = __fmul_rn(x[e], s.lmean[col + e]);
              }
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) split_pack(x[2 * e], x[2 * e + 1], hi[h8 * 4 + e], lo[h8 * 4 + e]);
          }
          tmem_st8(tp + (k * (AC / 16) + ks) * 8, hi);
          tmem_st8(tp + K0 / 2 + (k * (AC / 16) + ks) * 8, lo);
        }
      }
      tmem_st_wait();
      tc_fence_before();
      bar_arrive(&s.bar_a[g]);
      // ---- basis_mat on the tensor cores: feat = products @ basis^T (27 of 32 accumulator columns)
      float feat[F];
      {
        ok = ok && bar_wait(&s.bar_d[g], ph_d, p.error);
        ph_d ^= 1;
        tc_fence_after();
        uint32_t v[32];
        tmem_ld32(tq, v);
#pragma unroll
        for (int f = 0; f < F; ++f) feat[f] = __uint_as_float(v[f]);
        tc_fence_before();
        bar_arrive(&s.bar_q);        // Q drained
      }
      // ---- positional encoding -> the 160-wide input row, written as the split-BF16 A operand into P_g
      {
        float sf[F], cf[F], sx[3], cx[3];
#pragma unroll
        for (int f = 0; f < F; ++f) sincosf(feat[f], &sf[f], &cf[f]);
#pragma unroll
        for (int d = 0; d < 3; ++d) sincosf(xv[d], &sx[d], &cx[d]);
#pragma unroll
        for (int ks = 0; ks < K1 / 16; ++ks) {
          uint32_t hi[8], lo[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float a = TC5_INPUT_ELEM(ks * 16 + 2 * j, feat, xv, sf, cf, sx, cx);
            const float b = TC5_INPUT_ELEM(ks * 16 + 2 * j + 1, feat, xv, sf, cf, sx, cx);
            split_pack(a, b, hi[j], lo[j]);
          }
          tmem_st8(tp + ks * 8, hi);
          tmem_st8(tp + K1 / 2 + ks * 8, lo);
        }
      }
      tmem_st_wait();
      tc_fence_before();
      bar_arrive(&s.bar_a[g]);
      // ---- layer 0 epilogue: bias + ReLU, hidden activations go back to P_g as the next A operand
      ok = ok && bar_wait(&s.bar_d[g], ph_d, p.error);
      ph_d ^= 1;
      tc_fence_after();
#pragma unroll 1
      for (int c = 0; c < HID / 32; ++c) {
        uint32_t v[32];
        tmem_ld32(tq + c * 32, v);
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          uint32_t hi[8], lo[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int n = c * 32 + half * 16 + 2 * j;
            const float a = fmaxf(__uint_as_float(v[half * 16 + 2 * j]) + s.b0[n], 0.f);
            const float b = fmaxf(__uint_as_float(v[half * 16 + 2 * j + 1]) + s.b0[n + 1], 0.f);
            split_pack(a, b, hi[j], lo[j]);
          }
          tmem_st8(tp + (c * 2 + half) * 8, hi);
          tmem_st8(tp + HID / 2 + (c * 2 + half) * 8, lo);
        }
      }
      tmem_st_wait();
      tc_fence_before();
      bar_arrive(&s.bar_q);          // Q drained
      bar_arrive(&s.bar_a[g]);       // A operand of layer 1 ready
      // ---- layer 1 epilogue + output layer (128 -> out_dim, fp32 on the CUDA cores)
      ok = ok && bar_wait(&s.bar_d[g], ph_d, p.error);
      ph_d ^= 1;
      tc_fence_after();
      float o[4] = {s.b2[0], s.b2[1], s.b2[2], s.b2[3]};
#pragma unroll 1
      for (int c = 0; c < HID / 32; ++c) {
        uint32_t v[32];
        tmem_ld32(tq + c * 32, v);
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const int n = c * 32 + j;
          const float h0 = fmaxf(__uint_as_float(v[j]) + s.b1[n], 0.f), h1 = fmaxf(__uint_as_float(v[j + 1]) + s.b1[n + 1], 0.f);
          const float h2 = fmaxf(__uint_as_float(v[j + 2]) + s.b1[n + 2], 0.f), h3 = fmaxf(__uint_as_float(v[j + 3]) + s.b1[n + 3], 0.f);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float4 w4 = *reinterpret_cast<const float4*>(s.w2 + q * HID + n);
            o[q] = fmaf(h3, w4.w, fmaf(h2, w4.z, fmaf(h1, w4.y, fmaf(h0, w4.x, o[q]))));
          }
        }
      }
      tc_fence_before();
      bar_arrive(&s.bar_q);          // Q drained
      if (live && ok) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          if (q < out_dim) {
            const float y = p.act == 0 ? 1.f / (1.f + expf(-o[q])) : tanhf(o[q]);
            if (POINTS) p.out[i * out_dim + q] = y;
            else atomicAdd(p.rgb_out + (int64_t)ray * 3 + q, __fmul_rn(wgt, y));
          }
        }
      }
      if (!ok) break;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == NCONS / 32) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
  }
}

int check_shapes(const TirField* f, const TirMlp* m) {
  if (f->aC != AC) return TIR_ERR_SHAPE;
  if (m->feat_dim != F || m->hidden != HID || m->pe_feat != 2 || m->pe_x != 2) return TIR_ERR_SHAPE;
  if (m->out_dim < 1 || m->out_dim > 4) return TIR_ERR_SHAPE;
  if (!m->w0 || !m->b0 || !m->w1 || !m->b1 || !m->w2 || !m->b2 || !m->basis) return TIR_ERR_NULL;
  return TIR_OK;
}

int* error_flag() {
  static int* flag = nullptr;
  if (!flag) {
    if (cudaMalloc(&flag, sizeof(int)) != cudaSuccess) return nullptr;
    cudaMemset(flag, 0, sizeof(int));
  }
  return flag;
}

template <bool POINTS>
int launch5(Tc5Params& p, int64_t max_items, cudaStream_t stream) {
  static bool configured[2] = {false, false};
  const int smem = (int)sizeof(Smem5) + 1024;
  if (!configured[POINTS]) {
    cudaError_t e = cudaFuncSetAttribute(app_mlp_tc5_kernel<POINTS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return (int)e;
    configured[POINTS] = true;
  }
  p.error = error_flag();
  if (!p.error) return TIR_ERR_NULL;
  const int64_t pairs = ((max_items + ROWS - 1) / ROWS + 1) / 2;
  const int blocks = (int)(pairs < 148 ? (pairs > 0 ? pairs : 1) : 148);
  app_mlp_tc5_kernel<POINTS><<<blocks, NTHREADS, smem, stream>>>(p);
  return (int)cudaGetLastError();
}

}  // namespace

// 0 = no tcgen05 protocol error so far (host read, synchronises the device; tests / diagnostics only)
extern "C" int tir_mlp_tc5_error(void) {
  int* flag = error_flag();
  int v = -1;
  if (!flag || cudaMemcpy(&v, flag, sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
  return v;
}

extern "C" int tir_app_mlp_tc5(const TirField* field, const TirMlp* mlp, const TirAppSample* samples,
                               const uint32_t* sample_count, int64_t max_samples, const float* ray_dirs,
                               int32_t n_dirs, const int32_t* light_idx, float* rgb_out, void* stream) {
  if (!field || !mlp || !samples || !sample_count || !ray_dirs || !rgb_out) return TIR_ERR_NULL;
  int rc = check_shapes(field, mlp);
  if (rc) return rc;
  if (mlp->out_dim != 3) return TIR_ERR_SHAPE;
  Tc5Params p{};
  p.f = *field; p.mlp = *mlp; p.samples = samples; p.sample_count = sample_count; p.max_samples = max_samples;
  p.ray_dirs = ray_dirs; p.n_dirs = n_dirs; p.light_idx = light_idx; p.rgb_out = rgb_out; p.act = 0;
  p.light_mode = mlp->light_line ? 1 : 0;
  return launch5<false>(p, max_samples, (cudaStream_t)stream);
}

extern "C" int tir_app_mlp_points_tc5(const TirField* field, const TirMlp* mlp, const float* xn, const float* x_in,
                                      const int32_t* light_idx, int64_t n, int32_t act, float* out, void* stream) {
  if (n <= 0) return TIR_OK;
  if (!field || !mlp || !xn || !x_in || !out) return TIR_ERR_NULL;
  int rc = check_shapes(field, mlp);
  if (rc) return rc;
  if (act != 0 && act != 1) return TIR_ERR_CONFIG;
  Tc5Params p{};
  p.f = *field; p.mlp = *mlp; p.pts_xn = xn; p.pts_x = x_in; p.n_points = n; p.light_idx = light_idx; p.out = out;
  p.act = act; p.light_mode = mlp->light_line ? 1 : 0;
  return launch5<true>(p, n, (cudaStream_t)stream);
}
