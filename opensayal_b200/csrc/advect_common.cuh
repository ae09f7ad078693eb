// advect_common.cuh — device functions of the semi-Lagrangian advection shared by the plain kernels
// (kernels_basic.cu) and the shared-memory tile kernels (advect_tile.cu): the reference's samplers
// Fluid::get_general_velocity_x/_y (/root/reference/src/fluid.cu:418-539) and Fluid::interpolate_smoke
// (fluid.cu:644-716), reading global memory.  Every operation is an explicit round-to-nearest intrinsic.
#ifndef SAYAL_ADVECT_COMMON_CUH
#define SAYAL_ADVECT_COMMON_CUH

#include "sayal_internal.h"

// the tile kernels call these as out-of-line fallbacks (define SAYAL_SAMPLER_ATTR as __noinline__ before including)
#ifndef SAYAL_SAMPLER_ATTR
#define SAYAL_SAMPLER_ATTR
#endif

namespace sayal {

struct View {
  const float* __restrict__ u;
  const float* __restrict__ v;
  const float* __restrict__ smoke;
  const uint8_t* __restrict__ flags;
  int32_t* overflow;
};

__device__ __forceinline__ int f2i_rz(float x) { return __float2int_rz(x); }  // cvt.rzi.s32.f32: saturating, NaN -> 0

template <int HC>
__device__ __forceinline__ float div_h(float x, float hf) { return HC == 1 ? x : __fdiv_rn(x, hf); }
template <int HC>
__device__ __forceinline__ float mul_h(int k, int h) { return HC == 1 ? (float)k : (float)(k * h); }
template <int HC>
__device__ __forceinline__ float pos_half(int k, int h) {
  float c = __fadd_rn((float)k, 0.5f);
  return HC == 1 ? c : __fmul_rn(c, (float)h);
}

// Fluid::is_valid_fluid (fluid.cu:360-362) with all bounds tests, for the base cell of a sample.  On success
// *k is the cell's index and its 3x3 neighbourhood is addressable.
__device__ __forceinline__ bool base_fluid(const Grid& g, const View& w, int i, int j, long* k) {
  if (i < 0 || j < 0 || i >= g.W || j >= g.H) return false;
  int lr = (g.H - 1 - j) - g.row_base;
  if (lr < g.valid_lo || lr >= g.valid_hi) {  // a slab's back-trace left its ghost rows: report, do not guess
    atomicAdd(w.overflow, 1);
    return false;
  }
  long kk = (long)lr * g.pitch + i;
  if (w.flags[kk] & FL_SOLID) return false;
  if (lr < g.valid_lo + 1 || lr > g.valid_hi - 2) {  // fluid cell on the first/last local row: only possible in a slab
    atomicAdd(w.overflow, 1);
    return false;
  }
  *k = kk;
  return true;
}

__device__ __forceinline__ bool open_tap(const View& w, long k) { return !(w.flags[k] & FL_SOLID); }

// Fluid::get_general_velocity_y (fluid.cu:418-477).  Memory row of (i, j+1) is k - pitch.
template <int HC>
__device__ SAYAL_SAMPLER_ATTR float general_velocity_y(const Grid& g, const View& w, float x, float y) {
  const float hf = (float)g.h;
  const double half = HC == 1 ? 0.5 : (double)g.h / 2.0;
  int i = f2i_rz(div_h<HC>(x, hf)), j = f2i_rz(div_h<HC>(y, hf));
  long k;
  if (!base_fluid(g, w, i, j, &k)) return 0.f;
  const long up = -(long)g.pitch;
  float in_x = __fsub_rn(x, mul_h<HC>(i, g.h));
  float in_y = __fsub_rn(y, mul_h<HC>(j, g.h));
  float w_y = __fsub_rn(1.0f, div_h<HC>(in_y, hf));
  float n_y = __fsub_rn(1.0f, w_y);
  float avg = 0.f;
  if ((double)in_x < half) {
    float d_x = (float)__dsub_rn(half, (double)in_x);
    float w_x = __fsub_rn(1.0f, div_h<HC>(d_x, hf));
    float n_x = __fsub_rn(1.0f, w_x);
    avg = __fmaf_rn(__fmul_rn(w_y, w_x), w.v[k], avg);
    if (open_tap(w, k - 1)) avg = __fmaf_rn(__fmul_rn(w_y, n_x), w.v[k - 1], avg);
    if (open_tap(w, k - 1 + up)) avg = __fmaf_rn(__fmul_rn(n_y, n_x), w.v[k - 1 + up], avg);
    if (open_tap(w, k + up)) avg = __fmaf_rn(__fmul_rn(n_y, w_x), w.v[k + up], avg);
  } else {
    float d_x = (float)__dsub_rn((double)in_x, half);
    float w_x = __fsub_rn(1.0f, div_h<HC>(d_x, hf));
    float n_x = __fsub_rn(1.0f, w_x);
    avg = __fmaf_rn(__fmul_rn(w_y, w_x), w.v[k], avg);
    if (open_tap(w, k + up)) avg = __fmaf_rn(__fmul_rn(n_y, w_x), w.v[k + up], avg);
    if (open_tap(w, k + 1 + up)) avg = __fmaf_rn(__fmul_rn(n_y, n_x), w.v[k + 1 + up], avg);
    if (open_tap(w, k + 1)) avg = __fmaf_rn(__fmul_rn(w_y, n_x), w.v[k + 1], avg);
  }
  return avg;
}

// Fluid::get_general_velocity_x (fluid.cu:479-539).  Memory row of (i, j-1) is k + pitch.
template <int HC>
__device__ SAYAL_SAMPLER_ATTR float general_velocity_x(const Grid& g, const View& w, float x, float y) {
  const float hf = (float)g.h;
  const double half = HC == 1 ? 0.5 : (double)g.h / 2.0;
  int i = f2i_rz(div_h<HC>(x, hf)), j = f2i_rz(div_h<HC>(y, hf));
  long k;
  if (!base_fluid(g, w, i, j, &k)) return 0.f;
  const long up = -(long)g.pitch, down = (long)g.pitch;
  float in_x = __fsub_rn(x, mul_h<HC>(i, g.h));
  float in_y = __fsub_rn(y, mul_h<HC>(j, g.h));
  float w_x = __fsub_rn(1.0f, div_h<HC>(in_x, hf));
  float n_x = __fsub_rn(1.0f, w_x);
  float avg = 0.f;
  if ((double)in_y <= half) {  // note <= here, < in _y (fluid.cu:493 vs 432)
    float d_y = (float)__dsub_rn(half, (double)in_y);
    float w_y = __fsub_rn(1.0f, div_h<HC>(d_y, hf));
    float n_y = __fsub_rn(1.0f, w_y);
    avg = __fmaf_rn(__fmul_rn(w_y, w_x), w.u[k], avg);
    if (open_tap(w, k + 1)) avg = __fmaf_rn(__fmul_rn(w_y, n_x), w.u[k + 1], avg);
    if (open_tap(w, k + down)) avg = __fmaf_rn(__fmul_rn(n_y, w_x), w.u[k + down], avg);
    if (open_tap(w, k + 1 + down)) avg = __fmaf_rn(__fmul_rn(n_y, n_x), w.u[k + 1 + down], avg);
  } else {
    float d_y = (float)__dsub_rn((double)in_y, half);
    float w_y = __fsub_rn(1.0f, div_h<HC>(d_y, hf));
    float n_y = __fsub_rn(1.0f, w_y);
    avg = __fmaf_rn(__fmul_rn(w_y, w_x), w.u[k], avg);
    if (open_tap(w, k + up)) avg = __fmaf_rn(__fmul_rn(n_y, w_x), w.u[k + up], avg);
    if (open_tap(w, k + 1)) avg = __fmaf_rn(__fmul_rn(w_y, n_x), w.u[k + 1], avg);
    if (open_tap(w, k + 1 + up)) avg = __fmaf_rn(__fmul_rn(n_y, n_x), w.u[k + 1 + up], avg);
  }
  return avg;
}

// Fluid::is_valid_fluid for an arbitrary tap of interpolate_smoke, whose base cell may be anything
__device__ __forceinline__ bool fluid_at(const Grid& g, const View& w, int i, int j, long* k) {
  if (i < 0 || j < 0 || i >= g.W || j >= g.H) return false;
  int lr = (g.H - 1 - j) - g.row_base;
  if (lr < g.valid_lo || lr >= g.valid_hi) {
    atomicAdd(w.overflow, 1);
    return false;
  }
  *k = (long)lr * g.pitch + i;
  return !(w.flags[*k] & FL_SOLID);
}

// Fluid::interpolate_smoke (fluid.cu:644-716): inverse-distance weights over the quadrant's four centres
template <int HC>
__device__ SAYAL_SAMPLER_ATTR float interpolate_smoke(const Grid& g, const View& w, float x, float y) {
  const float hf = (float)g.h;
  const double half = HC == 1 ? 0.5 : (double)g.h / 2.0;
  int i = f2i_rz(div_h<HC>(x, hf)), j = f2i_rz(div_h<HC>(y, hf));
  float in_x = __fsub_rn(x, mul_h<HC>(i, g.h));
  float in_y = __fsub_rn(y, mul_h<HC>(j, g.h));
  int di = ((double)in_x < half) ? -1 : 1;
  int dj = ((double)in_y < half) ? -1 : 1;
  // distances to the centres of (i,j), (i+di,j), (i,j+dj), (i+di,j+dj)  (helper.cuh:77-81)
  float dx0 = __fsub_rn(x, pos_half<HC>(i, g.h)), dx1 = __fsub_rn(x, pos_half<HC>(i + di, g.h));
  float dy0 = __fsub_rn(y, pos_half<HC>(j, g.h)), dy1 = __fsub_rn(y, pos_half<HC>(j + dj, g.h));
  float yy0 = __fmul_rn(dy0, dy0), yy1 = __fmul_rn(dy1, dy1);
  float dist[4] = {__fsqrt_rn(__fmaf_rn(dx0, dx0, yy0)), __fsqrt_rn(__fmaf_rn(dx1, dx1, yy0)),
                   __fsqrt_rn(__fmaf_rn(dx0, dx0, yy1)), __fsqrt_rn(__fmaf_rn(dx1, dx1, yy1))};
  float inv[4];
#pragma unroll
  for (int t = 0; t < 4; t++)  // float inv = 1.0 / (distance + 1e-6): FP64 (fluid.cu:690-693); rcp.rn == 1.0/x
    inv[t] = (float)__drcp_rn(__dadd_rn((double)dist[t], 1e-6));
  float sum_inv = __fadd_rn(__fadd_rn(__fadd_rn(inv[0], inv[1]), inv[2]), inv[3]);
  float avg = 0.f;
  // the common case: base cell is an interior cell => no bounds tests
  bool interior = i >= 1 && j >= 1 && i <= g.W - 2 && j <= g.H - 2;
  int lr = (g.H - 1 - j) - g.row_base;
  if (interior && lr >= g.valid_lo + 1 && lr <= g.valid_hi - 2) {
    long k = (long)lr * g.pitch + i;
    long kj = -(long)dj * g.pitch;  // (i, j+dj)
#pragma unroll
    for (int t = 0; t < 4; t++) {
      long kk = k + ((t & 1) ? di : 0) + ((t & 2) ? kj : 0);
      float wt = __fdiv_rn(inv[t], sum_inv);
      if (open_tap(w, kk)) avg = __fmaf_rn(wt, w.smoke[kk], avg);
    }
  } else {
#pragma unroll
    for (int t = 0; t < 4; t++) {
      int ti = i + ((t & 1) ? di : 0), tj = j + ((t & 2) ? dj : 0);
      float wt = __fdiv_rn(inv[t], sum_inv);
      long kk;
      if (fluid_at(g, w, ti, tj, &kk)) avg = __fmaf_rn(wt, w.smoke[kk], avg);
    }
  }
  return avg;
}

// avg / count for count in 1..4 (fluid.cu:386, 413): x * 0.5 and x * 0.25 are the correctly rounded quotients,
// only count == 3 needs a real division
__device__ __forceinline__ float div_count(float x, int count) {
  if (count == 3) return __fdiv_rn(x, 3.0f);
  return __fmul_rn(x, count == 4 ? 0.25f : (count == 2 ? 0.5f : 1.0f));
}


}  // namespace sayal

#endif
