// Host build of the fused primary tail (test harness for tail_body.h; same argument lists as the CUDA entry points).
#include <stdint.h>

#include "tail_body.h"

static TailSample load(int64_t i, const float* w, const int64_t* ray, const float* rgb, const float* brdf,
                       const float* brdfj, const float* vn, const float* dn, const float* viewdirs) {
  TailSample s;
  s.w = w[i];
  for (int c = 0; c < 3; ++c) {
    s.rgb[c] = rgb[i * 3 + c];
    s.vn[c] = vn[i * 3 + c];
    s.dn[c] = dn ? dn[i * 3 + c] : 0.f;
    s.vd[c] = viewdirs[ray[i] * 3 + c];
  }
  for (int c = 0; c < 4; ++c) {
    s.brdf[c] = brdf[i * 4 + c];
    s.brdfj[c] = brdfj[i * 4 + c];
  }
  return s;
}

extern "C" void tail_forward_host(int64_t n, const float* w, const int64_t* ray, const float* rgb, const float* brdf,
                                  const float* brdfj, const float* vn, const float* dn, const float* viewdirs,
                                  float* packed /* [n_rays, 14], zero-initialised by the caller */) {
  for (int64_t i = 0; i < n; ++i) {
    const TailSample s = load(i, w, ray, rgb, brdf, brdfj, vn, dn, viewdirs);
    float v[TAIL_CH];
    tail_channels(s, dn != nullptr, v);
    for (int k = 0; k < TAIL_CH; ++k) packed[ray[i] * TAIL_CH + k] += s.w * v[k];
  }
}

extern "C" void tail_backward_host(int64_t n, const float* w, const int64_t* ray, const float* rgb, const float* brdf,
                                   const float* brdfj, const float* vn, const float* dn, const float* viewdirs,
                                   const float* g_packed, float* g_w, float* g_rgb, float* g_brdf, float* g_brdfj,
                                   float* g_vn, float* g_dn) {
  for (int64_t i = 0; i < n; ++i) {
    const TailSample s = load(i, w, ray, rgb, brdf, brdfj, vn, dn, viewdirs);
    TailGrad g;
    tail_backward_sample(s, dn != nullptr, g_packed + ray[i] * TAIL_CH, g);
    g_w[i] = g.w;
    for (int c = 0; c < 3; ++c) {
      g_rgb[i * 3 + c] = g.rgb[c];
      g_vn[i * 3 + c] = g.vn[c];
      if (g_dn) g_dn[i * 3 + c] = g.dn[c];
    }
    for (int c = 0; c < 4; ++c) {
      g_brdf[i * 4 + c] = g.brdf[c];
      g_brdfj[i * 4 + c] = g.brdfj[c];
    }
  }
}
