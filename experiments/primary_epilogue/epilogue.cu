// CUDA side of the per-ray epilogue (experiments; not linked into libtensoir_b200.so yet).  One thread per ray; the two
// scalar smoothness losses (torch.mean over rays) are block-reduced and accumulated with one atomic per block.
// Math in epilogue_body.h (validated on the CPU by test_epilogue_host.py).  STATUS: compiles for sm_100a, not yet run.
#include <cuda_runtime.h>
#include <stdint.h>

#include "epilogue_body.h"

namespace {

__device__ __forceinline__ EpiIn load_ray(int64_t r, const float* __restrict__ packed, const float* __restrict__ acc,
                                          const float* __restrict__ depth, const float* __restrict__ rays, float fresnel0,
                                          int bg) {
  EpiIn in;
#pragma unroll
  for (int k = 0; k < 14; ++k) in.P[k] = packed[r * 14 + k];
  in.acc = acc[r]; in.depth = depth[r]; in.dz = rays[r * 6 + 5]; in.fresnel0 = fresnel0; in.bg = bg;
  return in;
}

// out [n,18] in EpiOut order; acc_mask [n] (acc > 0.5); losses[2] += mean(albedo cost), mean(roughness cost)
__global__ void epilogue_fwd_kernel(int64_t n, const float* packed, const float* acc, const float* depth,
                                    const float* rays, float fresnel0, int bg, float* __restrict__ out,
                                    uint8_t* __restrict__ acc_mask, float* __restrict__ losses) {
  float s_ac = 0.f, s_rc = 0.f;
  for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x) {
    const EpiIn in = load_ray(r, packed, acc, depth, rays, fresnel0, bg);
    EpiOut o;
    epi_forward(in, o);
    const float* f = reinterpret_cast<const float*>(&o);
#pragma unroll
    for (int k = 0; k < 18; ++k) out[r * 18 + k] = f[k];
    if (acc_mask) acc_mask[r] = in.acc > 0.5f ? 1 : 0;
    s_ac += o.ac; s_rc += o.rc;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s_ac += __shfl_xor_sync(0xffffffffu, s_ac, o);
    s_rc += __shfl_xor_sync(0xffffffffu, s_rc, o);
  }
  if ((threadIdx.x & 31) == 0 && losses) {
    atomicAdd(losses + 0, s_ac / (float)n);
    atomicAdd(losses + 1, s_rc / (float)n);
  }
}

// g_out [n,18] (columns 16,17 ignored: the scalar losses' gradients come in g_losses[2]) -> g_packed [n,14], g_acc, g_depth
__global__ void epilogue_bwd_kernel(int64_t n, const float* packed, const float* acc, const float* depth,
                                    const float* rays, float fresnel0, int bg, const float* __restrict__ g_out,
                                    const float* __restrict__ g_losses, float* __restrict__ g_packed,
                                    float* __restrict__ g_acc, float* __restrict__ g_depth) {
  const float g_ac = g_losses ? g_losses[0] / (float)n : 0.f, g_rc = g_losses ? g_losses[1] / (float)n : 0.f;
  for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x) {
    const EpiIn in = load_ray(r, packed, acc, depth, rays, fresnel0, bg);
    EpiOut g;
    float* f = reinterpret_cast<float*>(&g);
#pragma unroll
    for (int k = 0; k < 16; ++k) f[k] = g_out[r * 18 + k];
    g.ac = g_ac; g.rc = g_rc;
    float gP[14], ga, gd;
    epi_backward(in, g, gP, &ga, &gd);
#pragma unroll
    for (int k = 0; k < 14; ++k) g_packed[r * 14 + k] = gP[k];
    g_acc[r] = ga;
    g_depth[r] = gd;
  }
}

inline int blocks_for(int64_t n) {
  int64_t b = (n + 127) / 128;
  return (int)(b < 148 * 8 ? (b > 0 ? b : 1) : 148 * 8);
}

}  // namespace

extern "C" int tir_epilogue_fwd(int64_t n, const float* packed, const float* acc, const float* depth, const float* rays,
                                float fresnel0, int bg, float* out, uint8_t* acc_mask, float* losses, void* stream) {
  if (n <= 0) return 0;
  if (!packed || !acc || !depth || !rays || !out) return -1;
  epilogue_fwd_kernel<<<blocks_for(n), 128, 0, (cudaStream_t)stream>>>(n, packed, acc, depth, rays, fresnel0, bg, out,
                                                                       acc_mask, losses);
  return (int)cudaGetLastError();
}

extern "C" int tir_epilogue_bwd(int64_t n, const float* packed, const float* acc, const float* depth, const float* rays,
                                float fresnel0, int bg, const float* g_out, const float* g_losses, float* g_packed,
                                float* g_acc, float* g_depth, void* stream) {
  if (n <= 0) return 0;
  if (!packed || !acc || !depth || !rays || !g_out || !g_packed || !g_acc || !g_depth) return -1;
  epilogue_bwd_kernel<<<blocks_for(n), 128, 0, (cudaStream_t)stream>>>(n, packed, acc, depth, rays, fresnel0, bg, g_out,
                                                                       g_losses, g_packed, g_acc, g_depth);
  return (int)cudaGetLastError();
}
