// Host build of the per-ray epilogue (test harness for epilogue_body.h).
#include <stdint.h>

#include "epilogue_body.h"

extern "C" void epilogue_forward_host(int64_t n, const float* packed, const float* acc, const float* depth,
                                      const float* rays, float fresnel0, int bg, float* out /* [n, 18] */) {
  for (int64_t r = 0; r < n; ++r) {
    EpiIn in;
    for (int k = 0; k < 14; ++k) in.P[k] = packed[r * 14 + k];
    in.acc = acc[r]; in.depth = depth[r]; in.dz = rays[r * 6 + 5]; in.fresnel0 = fresnel0; in.bg = bg;
    EpiOut o;
    epi_forward(in, o);
    const float* f = reinterpret_cast<const float*>(&o);
    for (int k = 0; k < 18; ++k) out[r * 18 + k] = f[k];
  }
}

extern "C" void epilogue_backward_host(int64_t n, const float* packed, const float* acc, const float* depth,
                                       const float* rays, float fresnel0, int bg, const float* g_out /* [n, 18] */,
                                       float* g_packed, float* g_acc, float* g_depth) {
  for (int64_t r = 0; r < n; ++r) {
    EpiIn in;
    for (int k = 0; k < 14; ++k) in.P[k] = packed[r * 14 + k];
    in.acc = acc[r]; in.depth = depth[r]; in.dz = rays[r * 6 + 5]; in.fresnel0 = fresnel0; in.bg = bg;
    EpiOut g;
    float* f = reinterpret_cast<float*>(&g);
    for (int k = 0; k < 18; ++k) f[k] = g_out[r * 18 + k];
    epi_backward(in, g, g_packed + r * 14, g_acc + r, g_depth + r);
  }
}
