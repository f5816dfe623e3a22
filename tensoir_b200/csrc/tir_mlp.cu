// Appearance gather (3 x 48-channel plane*line products) -> * light_line -> basis_mat (144 -> 27) ->
// positional encoding -> 150 -> 128 -> 128 -> {3,4} MLP, on a compacted list of samples.
// Replaces compute_appfeature / compute_intrinfeature (tensoRF_rotated_lights.py:167-224) + MLPRender_Fea /
// MLPBRDF_PEandFeature (tensorBase:122-146, :182-208).
//
// v1: fp32 SIMT.  Persistent CTAs keep all weights resident in shared memory (162 KB) and stream tiles of
// 48 samples through two ping-pong activation buffers stored k-major ([k][sample]) so that the register-tiled
// GEMM inner loops read 128-bit, conflict-free operands.  (1e-4 relative parity rules out plain TF32/BF16
// tensor-core math; an error-compensated tcgen05 path is the planned replacement.)
#include "tir_device.cuh"

using namespace tir;

namespace {

constexpr int AC = 48;            // appearance channels / orientation
constexpr int K0 = 3 * AC;        // 144
constexpr int F = 27;             // app_dim
constexpr int FP = 28;            // padded
constexpr int HID = 128;
constexpr int IN = 150;           // 27 + 3 + 54 + 54 + 6 + 6
constexpr int INP = 152;
constexpr int M = 48;             // samples per tile
constexpr int NT = 192;           // threads: (M/4) x (HID/8) register tiles of 4 x 8

struct SmemLayout {
  float w0t[INP * HID];     // [k][h]
  float w1t[HID * HID];     // [k][h]
  float basist[K0 * FP];    // [k][f]
  float w2t[HID * 4];       // [k][o]
  float b0[HID];
  float b1[HID];
  float b2[4];
  float bufA[INP * M];      // [k][m]
  float bufB[K0 * M];       // [k][m]
  float xn[M][3];
  float xv[M][3];           // view dir (radiance head) or position (BRDF / normal heads)
  float wgt[M];
  int ray[M];
  int light[M];
};

struct MlpParams {
  TirField f;
  TirMlp mlp;
  // list source
  const TirAppSample* samples;
  const uint32_t* sample_count;
  int64_t max_samples;
  const float* ray_dirs;
  int n_dirs;
  const int32_t* light_idx;
  float* rgb_out;
  // point source
  const float* pts_xn;
  const float* pts_x;
  int64_t n_points;
  float* out;
  int act;        // 0 sigmoid, 1 tanh
};

template <bool POINTS>
__global__ void __launch_bounds__(NT, 1) app_mlp_kernel(const MlpParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SmemLayout& s = *reinterpret_cast<SmemLayout*>(smem_raw);
  const int tid = threadIdx.x;
  const TirMlp& mlp = p.mlp;
  const int out_dim = mlp.out_dim;

  // ---- stage weights (transposed to k-major) once per CTA
  for (int i = tid; i < INP * HID; i += NT) {
    int k = i / HID, h = i % HID;
    s.w0t[i] = (k < IN) ? __ldg(mlp.w0 + h * IN + k) : 0.f;
  }
  for (int i = tid; i < HID * HID; i += NT) {
    int k = i / HID, h = i % HID;
    s.w1t[i] = __ldg(mlp.w1 + h * HID + k);
  }
  for (int i = tid; i < K0 * FP; i += NT) {
    int k = i / FP, ff = i % FP;
    s.basist[i] = (ff < F) ? __ldg(mlp.basis + ff * K0 + k) : 0.f;
  }
  for (int i = tid; i < HID * 4; i += NT) {
    int k = i / 4, o = i % 4;
    s.w2t[i] = (o < out_dim) ? __ldg(mlp.w2 + o * HID + k) : 0.f;
  }
  for (int i = tid; i < HID; i += NT) { s.b0[i] = __ldg(mlp.b0 + i); s.b1[i] = __ldg(mlp.b1 + i); }
  if (tid < 4) s.b2[tid] = (tid < out_dim) ? __ldg(mlp.b2 + tid) : 0.f;
  __syncthreads();

  const int64_t total = POINTS ? p.n_points
                               : (int64_t)min((unsigned long long)*p.sample_count, (unsigned long long)p.max_samples);
  const int64_t n_tiles = (total + M - 1) / M;

  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t base = tile * M;
    // ---- sample metadata
    if (tid < M) {
      const int64_t i = base + tid;
      float xn0 = 0.f, xn1 = 0.f, xn2 = 0.f, v0 = 0.f, v1 = 0.f, v2 = 0.f, w = 0.f;
      int ray = -1, li = 0;
      if (i < total) {
        if (POINTS) {
          xn0 = p.pts_xn[i * 3 + 0]; xn1 = p.pts_xn[i * 3 + 1]; xn2 = p.pts_xn[i * 3 + 2];
          v0 = p.pts_x[i * 3 + 0]; v1 = p.pts_x[i * 3 + 1]; v2 = p.pts_x[i * 3 + 2];
          ray = (int)i; w = 1.f;
          li = p.light_idx ? p.light_idx[i] : 0;
        } else {
          const TirAppSample sm = p.samples[i];
          xn0 = sm.xn[0]; xn1 = sm.xn[1]; xn2 = sm.xn[2]; w = sm.weight; ray = sm.ray;
          const int64_t di = p.n_dirs > 0 ? (int64_t)(ray % p.n_dirs) : (int64_t)ray;
          v0 = __ldg(p.ray_dirs + di * 3 + 0); v1 = __ldg(p.ray_dirs + di * 3 + 1); v2 = __ldg(p.ray_dirs + di * 3 + 2);
          li = p.light_idx ? __ldg(p.light_idx + (p.n_dirs > 0 ? ray / p.n_dirs : ray)) : 0;
        }
      }
      s.xn[tid][0] = xn0; s.xn[tid][1] = xn1; s.xn[tid][2] = xn2;
      s.xv[tid][0] = v0; s.xv[tid][1] = v1; s.xv[tid][2] = v2;
      s.wgt[tid] = w; s.ray[tid] = ray; s.light[tid] = li;
    }
    __syncthreads();

    // ---- phase 1: gather. 4 threads per sample, 12 channels of each orientation per thread -> bufB[k][m]
    {
      const int m = tid >> 2, qd = tid & 3;
      const float xn[3] = {s.xn[m][0], s.xn[m][1], s.xn[m][2]};
      const float* lrow = mlp.light_line ? (mlp.light_line + (size_t)s.light[m] * K0) : nullptr;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const int m0 = kMat0[k], m1 = kMat1[k], v = kVec[k];
        const Bilinear b = bilinear_setup(xn[m0], xn[m1], p.f.grid[m0], p.f.grid[m1]);
        const Linear1 l = linear_setup(xn[v], p.f.grid[v]);
        const float* P = p.f.aplane[k];
        const float* L = p.f.aline[k];
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          const int c = qd * 12 + j * 4;
          const float4 pv = bilerp4(ldg4(P + (size_t)b.o00 * AC + c), ldg4(P + (size_t)b.o01 * AC + c),
                                    ldg4(P + (size_t)b.o10 * AC + c), ldg4(P + (size_t)b.o11 * AC + c), b);
          const float4 lv = lerp4(ldg4(L + (size_t)l.o0 * AC + c), ldg4(L + (size_t)l.o1 * AC + c), l);
          float4 lc = make_float4(1.f, 1.f, 1.f, 1.f);
          if (lrow) lc = ldg4(lrow + k * AC + c);
          const int kk = k * AC + c;
          // (plane * line) * light  (tensoRF_rotated_lights.py:222)
          s.bufB[(kk + 0) * M + m] = lrow ? __fmul_rn(__fmul_rn(pv.x, lv.x), lc.x) : __fmul_rn(pv.x, lv.x);
          s.bufB[(kk + 1) * M + m] = lrow ? __fmul_rn(__fmul_rn(pv.y, lv.y), lc.y) : __fmul_rn(pv.y, lv.y);
          s.bufB[(kk + 2) * M + m] = lrow ? __fmul_rn(__fmul_rn(pv.z, lv.z), lc.z) : __fmul_rn(pv.z, lv.z);
          s.bufB[(kk + 3) * M + m] = lrow ? __fmul_rn(__fmul_rn(pv.w, lv.w), lc.w) : __fmul_rn(pv.w, lv.w);
        }
      }
    }
    __syncthreads();

    // ---- phase 2: basis_mat (144 -> 27) + MLP input assembly into bufA[k][m]
    //      layout [feat 27 | x 3 | sin PE(feat) 54 | cos 54 | sin PE(x) 6 | cos 6]  (tensorBase:12-17, :136-142)
    {
      const int m = tid % M, fg = tid / M;   // fg in 0..3, 7 features each
      float a[7];
#pragma unroll
      for (int j = 0; j < 7; ++j) a[j] = 0.f;
      for (int k = 0; k < K0; ++k) {
        const float x = s.bufB[k * M + m];
        const float* bw = s.basist + k * FP + fg * 7;
#pragma unroll
        for (int j = 0; j < 7; ++j) a[j] = fmaf(x, bw[j], a[j]);
      }
#pragma unroll
      for (int j = 0; j < 7; ++j) {
        const int ff = fg * 7 + j;
        if (ff < F) {
          const float v = a[j];
          s.bufA[ff * M + m] = v;
          float s1, c1, s2, c2;
          sincosf(v, &s1, &c1);
          sincosf(__fmul_rn(v, 2.f), &s2, &c2);
          s.bufA[(30 + 2 * ff) * M + m] = s1; s.bufA[(31 + 2 * ff) * M + m] = s2;
          s.bufA[(84 + 2 * ff) * M + m] = c1; s.bufA[(85 + 2 * ff) * M + m] = c2;
        }
      }
      if (fg == 0) {
#pragma unroll
        for (int d = 0; d < 3; ++d) {
          const float v = s.xv[m][d];
          s.bufA[(27 + d) * M + m] = v;
          float s1, c1, s2, c2;
          sincosf(v, &s1, &c1);
          sincosf(__fmul_rn(v, 2.f), &s2, &c2);
          s.bufA[(138 + 2 * d) * M + m] = s1; s.bufA[(139 + 2 * d) * M + m] = s2;
          s.bufA[(144 + 2 * d) * M + m] = c1; s.bufA[(145 + 2 * d) * M + m] = c2;
        }
        s.bufA[150 * M + m] = 0.f; s.bufA[151 * M + m] = 0.f;
      }
    }
    __syncthreads();

    // ---- phases 3/4: two hidden layers, 4 samples x 8 units per thread
    const int mg = tid % 12, hg = tid / 12;
    auto layer = [&](const float* __restrict__ act, const float* __restrict__ wt, const float* __restrict__ bias,
                     int K, float* __restrict__ outb) {
      float accv[4][8];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) accv[i][j] = 0.f;
#pragma unroll 2
      for (int k = 0; k < K; ++k) {
        const float4 av = *reinterpret_cast<const float4*>(act + k * M + mg * 4);
        const float4 wa = *reinterpret_cast<const float4*>(wt + k * HID + hg * 8);
        const float4 wb = *reinterpret_cast<const float4*>(wt + k * HID + hg * 8 + 4);
        const float a4[4] = {av.x, av.y, av.z, av.w};
        const float w8[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 8; ++j) accv[i][j] = fmaf(a4[i], w8[j], accv[i][j]);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float bj = bias[hg * 8 + j];
        float4 o;
        o.x = fmaxf(accv[0][j] + bj, 0.f); o.y = fmaxf(accv[1][j] + bj, 0.f);
        o.z = fmaxf(accv[2][j] + bj, 0.f); o.w = fmaxf(accv[3][j] + bj, 0.f);
        *reinterpret_cast<float4*>(outb + (hg * 8 + j) * M + mg * 4) = o;
      }
    };
    layer(s.bufA, s.w0t, s.b0, INP, s.bufB);
    __syncthreads();
    layer(s.bufB, s.w1t, s.b1, HID, s.bufA);
    __syncthreads();

    // ---- phase 5: output layer + activation + composite
    if (tid < M * 4) {
      const int m = tid % M, o = tid / M;
      if (o < out_dim) {
        float a = 0.f;
        for (int k = 0; k < HID; ++k) a = fmaf(s.bufA[k * M + m], s.w2t[k * 4 + o], a);
        a += s.b2[o];
        const float y = p.act == 0 ? 1.f / (1.f + expf(-a)) : tanhf(a);
        const int64_t i = base + m;
        if (i < total) {
          if (POINTS) p.out[i * out_dim + o] = y;
          else atomicAdd(p.rgb_out + (int64_t)s.ray[m] * 3 + o, __fmul_rn(s.wgt[m], y));
        }
      }
    }
    __syncthreads();
  }
}

int check_shapes(const TirField* f, const TirMlp* m) {
  if (f->aC != AC) return TIR_ERR_SHAPE;
  if (m->feat_dim != F || m->hidden != HID || m->pe_feat != 2 || m->pe_x != 2) return TIR_ERR_SHAPE;
  if (m->out_dim < 1 || m->out_dim > 4) return TIR_ERR_SHAPE;
  if (!m->w0 || !m->b0 || !m->w1 || !m->b1 || !m->w2 || !m->b2 || !m->basis) return TIR_ERR_NULL;
  return TIR_OK;
}

template <bool POINTS>
int launch(const MlpParams& p, int64_t max_items, cudaStream_t stream) {
  static bool configured[2] = {false, false};
  const int smem = (int)sizeof(SmemLayout);
  if (!configured[POINTS]) {
    cudaError_t e = cudaFuncSetAttribute(app_mlp_kernel<POINTS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return (int)e;
    configured[POINTS] = true;
  }
  int64_t tiles = (max_items + M - 1) / M;
  int blocks = (int)(tiles < 148 ? (tiles > 0 ? tiles : 1) : 148);
  app_mlp_kernel<POINTS><<<blocks, NT, smem, stream>>>(p);
  return (int)cudaGetLastError();
}

}  // namespace

extern "C" int tir_app_mlp(const TirField* field, const TirMlp* mlp, const TirAppSample* samples,
                           const uint32_t* sample_count, int64_t max_samples, const float* ray_dirs, int32_t n_dirs,
                           const int32_t* light_idx, float* rgb_out, void* stream) {
  if (!field || !mlp || !samples || !sample_count || !ray_dirs || !rgb_out) return TIR_ERR_NULL;
  int rc = check_shapes(field, mlp);
  if (rc) return rc;
  if (mlp->out_dim != 3) return TIR_ERR_SHAPE;
  MlpParams p{};
  p.f = *field; p.mlp = *mlp; p.samples = samples; p.sample_count = sample_count; p.max_samples = max_samples;
  p.ray_dirs = ray_dirs; p.n_dirs = n_dirs; p.light_idx = light_idx; p.rgb_out = rgb_out; p.act = 0;
  return launch<false>(p, max_samples, (cudaStream_t)stream);
}

extern "C" int tir_app_mlp_points(const TirField* field, const TirMlp* mlp, const float* xn, const float* x_in,
                                  const int32_t* light_idx, int64_t n, int32_t act, float* out, void* stream) {
  if (!field || !mlp || !xn || !x_in || !out) return TIR_ERR_NULL;
  int rc = check_shapes(field, mlp);
  if (rc) return rc;
  if (act != 0 && act != 1) return TIR_ERR_CONFIG;
  if (n <= 0) return TIR_OK;
  MlpParams p{};
  p.f = *field; p.mlp = *mlp; p.pts_xn = xn; p.pts_x = x_in; p.n_points = n; p.light_idx = light_idx;
  p.out = out; p.act = act;
  return launch<true>(p, n, (cudaStream_t)stream);
}
