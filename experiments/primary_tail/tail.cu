// CUDA side of the fused primary tail (experiments; not linked into libtensoir_b200.so yet).
// One thread per appearance sample; the per-ray sums are float atomics (an appearance list of a 4096-ray batch has
// ~2e4 rows, 14 atomics each, spread over 4096 x 14 addresses).  The per-sample math lives in tail_body.h and is
// validated on the CPU by test_tail_host.py; what remains GPU-only here is indexing.
//
// STATUS: compiles for sm_100a; not yet executed on a GPU.
#include <cuda_runtime.h>
#include <stdint.h>

#include "tail_body.h"

namespace {

__device__ __forceinline__ TailSample load_sample(int64_t i, const float* __restrict__ w, const int64_t* __restrict__ ray,
                                                  const float* __restrict__ rgb, const float* __restrict__ brdf,
                                                  const float* __restrict__ brdfj, const float* __restrict__ vn,
                                                  const float* __restrict__ dn, const float* __restrict__ viewdirs) {
  TailSample s;
  s.w = w[i];
  const int64_t r = ray[i];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    s.rgb[c] = rgb[i * 3 + c];
    s.vn[c] = vn[i * 3 + c];
    s.dn[c] = dn ? dn[i * 3 + c] : 0.f;
    s.vd[c] = __ldg(viewdirs + r * 3 + c);
  }
  const float4 b = *reinterpret_cast<const float4*>(brdf + i * 4), bj = *reinterpret_cast<const float4*>(brdfj + i * 4);
  s.brdf[0] = b.x; s.brdf[1] = b.y; s.brdf[2] = b.z; s.brdf[3] = b.w;
  s.brdfj[0] = bj.x; s.brdfj[1] = bj.y; s.brdfj[2] = bj.z; s.brdfj[3] = bj.w;
  return s;
}

__global__ void tail_fwd_kernel(int64_t n, const float* w, const int64_t* ray, const float* rgb, const float* brdf,
                                const float* brdfj, const float* vn, const float* dn, const float* viewdirs,
                                float* __restrict__ packed) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const TailSample s = load_sample(i, w, ray, rgb, brdf, brdfj, vn, dn, viewdirs);
    if (s.w == 0.f) continue;                      // padding rows of the static-capacity list contribute nothing
    float v[TAIL_CH];
    tail_channels(s, dn != nullptr, v);
    float* dst = packed + ray[i] * TAIL_CH;
#pragma unroll
    for (int k = 0; k < TAIL_CH; ++k) atomicAdd(dst + k, s.w * v[k]);
  }
}

__global__ void tail_bwd_kernel(int64_t n, const float* w, const int64_t* ray, const float* rgb, const float* brdf,
                                const float* brdfj, const float* vn, const float* dn, const float* viewdirs,
                                const float* __restrict__ g_packed, float* g_w, float* g_rgb, float* g_brdf,
                                float* g_brdfj, float* g_vn, float* g_dn) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const TailSample s = load_sample(i, w, ray, rgb, brdf, brdfj, vn, dn, viewdirs);
    float G[TAIL_CH];
    const float* src = g_packed + ray[i] * TAIL_CH;
#pragma unroll
    for (int k = 0; k < TAIL_CH; ++k) G[k] = __ldg(src + k);
    TailGrad g;
    tail_backward_sample(s, dn != nullptr, G, g);
    g_w[i] = g.w;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      g_rgb[i * 3 + c] = g.rgb[c];
      g_vn[i * 3 + c] = g.vn[c];
      if (g_dn) g_dn[i * 3 + c] = g.dn[c];
    }
    *reinterpret_cast<float4*>(g_brdf + i * 4) = make_float4(g.brdf[0], g.brdf[1], g.brdf[2], g.brdf[3]);
    *reinterpret_cast<float4*>(g_brdfj + i * 4) = make_float4(g.brdfj[0], g.brdfj[1], g.brdfj[2], g.brdfj[3]);
  }
}

inline int blocks_for(int64_t n) {
  int64_t b = (n + 255) / 256;
  return (int)(b < 148 * 8 ? (b > 0 ? b : 1) : 148 * 8);
}

}  // namespace

// packed [n_rays, 14] must be zero-initialised by the caller.  Returns 0 or a cudaError_t.
extern "C" int tir_tail_fwd(int64_t n, const float* w, const int64_t* ray, const float* rgb, const float* brdf,
                            const float* brdfj, const float* vn, const float* dn, const float* viewdirs, float* packed,
                            void* stream) {
  if (n <= 0) return 0;
  if (!w || !ray || !rgb || !brdf || !brdfj || !vn || !viewdirs || !packed) return -1;
  tail_fwd_kernel<<<blocks_for(n), 256, 0, (cudaStream_t)stream>>>(n, w, ray, rgb, brdf, brdfj, vn, dn, viewdirs, packed);
  return (int)cudaGetLastError();
}

extern "C" int tir_tail_bwd(int64_t n, const float* w, const int64_t* ray, const float* rgb, const float* brdf,
                            const float* brdfj, const float* vn, const float* dn, const float* viewdirs,
                            const float* g_packed, float* g_w, float* g_rgb, float* g_brdf, float* g_brdfj, float* g_vn,
                            float* g_dn, void* stream) {
  if (n <= 0) return 0;
  if (!w || !ray || !rgb || !brdf || !brdfj || !vn || !viewdirs || !g_packed || !g_w || !g_rgb || !g_brdf || !g_brdfj ||
      !g_vn)
    return -1;
  tail_bwd_kernel<<<blocks_for(n), 256, 0, (cudaStream_t)stream>>>(n, w, ray, rgb, brdf, brdfj, vn, dn, viewdirs,
                                                                   g_packed, g_w, g_rgb, g_brdf, g_brdfj, g_vn, g_dn);
  return (int)cudaGetLastError();
}
