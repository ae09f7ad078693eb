// visual.cu — what the reference's renderer computes from the fluid state each frame, without SDL, plus the
// viscous diffusion stage.  Reference: GraphicsHandler::update (/root/reference/src/graphics_handler.cu:463-478):
//   update_fluid_pixels   (:214-302)  RGBA8888 frame from is_solid / smoke / pressure  -> render_pixels_kernel
//   update_traces         (:358-421)  path lines, Fluid::trace (fluid.cu:16-36)        -> path_lines_kernel
//   update_velocity_arrows(:304-356)  one arrow per arrows.distance cells (:168-200)   -> arrows_kernel
// and Fluid::apply_diffusion (fluid.cu:167-190)                                          -> diffusion_half_sweep_kernel
//
// The reference blocks the simulation on every one of these (kernel, cudaMemcpyAsync to pageable memory,
// cudaDeviceSynchronize — three times per frame).  Here the frame is rendered on the step stream into one of two
// device buffers and copied to pinned host memory on a second stream; the step stream only ever waits for the copy
// that last used the buffer it is about to overwrite (two frames ago).
//
// Roofline: HBM.  render: read smoke 4 B (+ p 4 B) + flags 1 B, write 4 B per cell.  diffusion: 8 B per cell per
// sweep (read + write u; neighbours come from L1/L2).
#include <cmath>
#include <cstdio>

#include "advect_common.cuh"

namespace sayal {

#define VIS_LAUNCH_CHECK(s, what)                                          \
  do {                                                                     \
    cudaError_t e__ = cudaGetLastError();                                  \
    if (e__ != cudaSuccess) {                                              \
      char m__[256];                                                       \
      snprintf(m__, sizeof m__, "%s: %s", what, cudaGetErrorString(e__)); \
      return set_error(SAYAL_ECUDA, m__);                                  \
    }                                                                      \
    (s)->launches++;                                                       \
  } while (0)

namespace {

// static_cast<uint8_t>(float) as nvcc emits it for the reference: F2I.U32.TRUNC (saturating, NaN -> 0), low byte
__device__ __forceinline__ unsigned f2u8(float x) { return __float2uint_rz(x) & 255u; }
__device__ __forceinline__ unsigned map_rgba(unsigned r, unsigned g, unsigned b, unsigned a) {
  return r << 24 | g << 16 | b << 8 | a;  // helper.cu:45-47
}
__device__ __forceinline__ float from_ordered_dev(int e) { return __int_as_float(e ^ ((e >> 31) & 0x7fffffff)); }

// hsv_to_rgb (helper.cu:3-43)
__device__ __forceinline__ unsigned hsv_pixel(float h, float s, float v) {
  float c = __fmul_rn(v, s);
  float x = __fmul_rn(c, __fsub_rn(1.0f, fabsf(__fsub_rn(fmodf(__fdiv_rn(h, 60.0f), 2.0f), 1.0f))));
  float m = __fsub_rn(v, c);
  float r_, g_, b_;
  if (h < 60.f) { r_ = c; g_ = x; b_ = 0.f; }
  else if (h < 120.f) { r_ = x; g_ = c; b_ = 0.f; }
  else if (h < 180.f) { r_ = 0.f; g_ = c; b_ = x; }
  else if (h < 240.f) { r_ = 0.f; g_ = x; b_ = c; }
  else if (h < 300.f) { r_ = x; g_ = 0.f; b_ = c; }
  else { r_ = c; g_ = 0.f; b_ = x; }
  return map_rgba(f2u8(__fmul_rn(__fadd_rn(r_, m), 255.0f)), f2u8(__fmul_rn(__fadd_rn(g_, m), 255.0f)),
                  f2u8(__fmul_rn(__fadd_rn(b_, m), 255.0f)), 255u);
}

__device__ __forceinline__ float clampf(float x, float lo, float hi) { return x < lo ? lo : (x > hi ? hi : x); }

// update_fluid_pixels_kernel (graphics_handler.cu:258-283).  One thread per 4 cells: 16-byte loads of smoke / p, one
// 4-byte load of flags, one 16-byte store.  `pixels` has pitch W (the host layout), rows [own_lo, own_hi).
__global__ void __launch_bounds__(256)
render_pixels_kernel(Grid g, const uint8_t* __restrict__ flags, const float* __restrict__ smoke,
                     const float* __restrict__ p, const int32_t* __restrict__ range, int enable_pressure,
                     int enable_smoke, uint32_t* __restrict__ pixels) {
  const int x4 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  const int lr = g.own_lo + blockIdx.y * blockDim.y + threadIdx.y;
  if (x4 >= g.W || lr >= g.own_hi) return;
  const size_t k = (size_t)lr * g.pitch + x4;
  const unsigned fl = *reinterpret_cast<const unsigned*>(flags + k);
  float sm[4] = {1.f, 1.f, 1.f, 1.f}, pr[4] = {0.f, 0.f, 0.f, 0.f};
  if (enable_smoke) {
    float4 t = *reinterpret_cast<const float4*>(smoke + k);
    sm[0] = t.x; sm[1] = t.y; sm[2] = t.z; sm[3] = t.w;
  }
  float mn = 0.f, mx = 0.f;
  if (enable_pressure) {
    float4 t = *reinterpret_cast<const float4*>(p + k);
    pr[0] = t.x; pr[1] = t.y; pr[2] = t.z; pr[3] = t.w;
    mn = from_ordered_dev(range[0]);
    mx = from_ordered_dev(range[1]);
  }
  uint32_t* out = pixels + (size_t)(lr - g.own_lo) * g.W + x4;
#pragma unroll
  for (int c = 0; c < 4; c++) {
    if (x4 + c >= g.W) break;
    unsigned px;
    if ((fl >> (8 * c)) & FL_SOLID) {
      px = map_rgba(80, 80, 80, 255);
    } else if (enable_pressure) {
      float norm_p;
      if (enable_smoke) {  // update_smoke_and_pressure (:221-238)
        norm_p = 0.f;
        if (pr[c] < 0.f && mn != 0.f) norm_p = __fdiv_rn(-pr[c], mn);
        else if (mx != 0.f) norm_p = __fdiv_rn(pr[c], mx);
      } else {             // update_pressure_pixel (:240-256)
        norm_p = pr[c] < 0.f ? __fdiv_rn(-pr[c], mn) : __fdiv_rn(pr[c], mx);
      }
      norm_p = clampf(norm_p, -1.0f, 1.0f);
      px = hsv_pixel(__fmul_rn(__fsub_rn(1.0f, norm_p), 120.0f), 1.0f, enable_smoke ? sm[c] : 1.0f);
    } else if (enable_smoke) {  // update_smoke_pixels (:214-219)
      unsigned col = (255u - f2u8(__fmul_rn(sm[c], 255.0f))) & 255u;
      px = map_rgba(255, col, col, 255);
    } else {
      continue;  // the reference leaves the pixel untouched
    }
    out[c] = px;
  }
}

// Fluid::trace (fluid.cu:16-36) from the centre of every `dist`-th cell (update_traces_kernel, :391-402)
__global__ void path_lines_kernel(Grid g, View w, int dist, int len, float d_t, int nx, int ny, int32_t* __restrict__ xs,
                                  int32_t* __restrict__ ys) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y * blockDim.y + threadIdx.y;
  if (a >= nx || b >= ny) return;
  const int i = a * dist, j = b * dist;
  int32_t* lx = xs + ((size_t)(ny - 1 - b) * nx + a) * len;
  int32_t* ly = ys + ((size_t)(ny - 1 - b) * nx + a) * len;
  const int lr = (g.H - 1 - j) - g.row_base;
  bool solid = true;
  if (lr >= 0 && lr < g.local_rows) solid = w.flags[(size_t)lr * g.pitch + i] & FL_SOLID;
  if (solid) {
    for (int k = 0; k < len; k++) lx[k] = ly[k] = -1;
    return;
  }
  float px = pos_half<0>(i, g.h), py = pos_half<0>(j, g.h);
  lx[0] = f2i_rz(roundf(px));
  ly[0] = g.H - 1 - f2i_rz(roundf(py));
  for (int k = 1; k < len; k++) {
    float vx = general_velocity_x<0>(g, w, px, py), vy = general_velocity_y<0>(g, w, px, py);
    px = __fmaf_rn(vx, d_t, px);
    py = __fmaf_rn(vy, d_t, py);
    lx[k] = f2i_rz(roundf(px));
    ly[k] = g.H - 1 - f2i_rz(roundf(py));
  }
}

// cosf / sinf / atan2f of the source evaluated in double and rounded: the same bits as the CPU restatement
__device__ __forceinline__ float cosf_cr(float x) { return (float)cos((double)x); }
__device__ __forceinline__ float sinf_cr(float x) { return (float)sin((double)x); }

// update_center_velocity_arrow_at (graphics_handler.cu:317-335) + make_arrow_data (:168-200)
__global__ void arrows_kernel(Grid g, View w, int dist, int cs, float length_multiplier, float threshold, float head_len,
                              int nx, int ny, sayal_arrow* __restrict__ out) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y * blockDim.y + threadIdx.y;
  if (a >= nx || b >= ny) return;
  const int i = a * dist, j = b * dist;
  sayal_arrow ar = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  sayal_arrow* dst = out + (size_t)(ny - 1 - b) * nx + a;
  const int lr = (g.H - 1 - j) - g.row_base;
  bool solid = true;
  if (lr >= 0 && lr < g.local_rows) solid = w.flags[(size_t)lr * g.pitch + i] & FL_SOLID;
  if (!solid) {
    const float x = (float)(((double)i + 0.5) * (double)cs);
    const float y = (float)(((double)(g.H - j - 1) + 0.5) * (double)cs);
    const float qy = __fsub_rn((float)(g.H * cs), y);
    const float vx = general_velocity_x<0>(g, w, x, qy), vy = general_velocity_y<0>(g, w, x, qy);
    const float angle = (float)atan2((double)vy, (double)vx);
    float length = __fsqrt_rn(__fmaf_rn(vx, vx, __fmul_rn(vy, vy)));
    if (!(length < threshold)) {
      const float head_angle = (float)(M_PI / 8);
      ar.valid = 1;
      ar.start_x = f2i_rz(x);
      ar.start_y = f2i_rz(y);
      length = __fmul_rn(length, length_multiplier);
      ar.end_x = ar.start_x + f2i_rz(__fmul_rn(length, cosf_cr(angle)));
      ar.end_y = ar.start_y + f2i_rz(__fmul_rn(-length, sinf_cr(angle)));
      ar.left_head_end_x = ar.end_x + f2i_rz(__fmul_rn(-head_len, cosf_cr(__fadd_rn(angle, head_angle))));
      ar.left_head_end_y = ar.end_y + f2i_rz(__fmul_rn(head_len, sinf_cr(__fadd_rn(angle, head_angle))));
      ar.right_head_end_x = ar.end_x + f2i_rz(__fmul_rn(-head_len, cosf_cr(__fsub_rn(head_angle, angle))));
      ar.right_head_end_y = ar.end_y + f2i_rz(__fmul_rn(-head_len, sinf_cr(__fsub_rn(head_angle, angle))));
    }
  }
  *dst = ar;
}

// One colour of one diffusion sweep (fluid.cu:176-183): interior cells with (i + j + colour) even.
__global__ void __launch_bounds__(256)
diffusion_half_sweep_kernel(Grid g, float a, float denom, int colour, float* __restrict__ u) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int lr = blockIdx.y * blockDim.y + threadIdx.y;
  if (lr < 1 || lr >= g.local_rows - 1) return;
  const int j = g.H - 1 - (g.row_base + lr);
  if (j < 1 || j > g.H - 2) return;
  const int i = 2 * t + 1 + ((j + 1 + colour) & 1);
  if (i > g.W - 2) return;
  const size_t k = (size_t)lr * g.pitch + i;
  // (i-1, j) + (i+1, j) + (i, j-1) + (i, j+1); memory row of j-1 is lr+1
  float sum = __fadd_rn(__fadd_rn(__fadd_rn(u[k - 1], u[k + 1]), u[k + g.pitch]), u[k - g.pitch]);
  u[k] = __fdiv_rn(__fmaf_rn(a, sum, u[k]), denom);
}

}  // namespace

int launch_diffusion(Sim* s, int iterations, float d_t) {
  if (iterations <= 0 || s->cfg.viscosity == 0.f) return SAYAL_OK;
  const float a = (s->cfg.viscosity * d_t) / (float)(s->g.h * s->g.h);
  const float denom = fmaf(4.0f, a, 1.0f);
  dim3 block(64, 4);
  dim3 grid((s->g.W / 2 + block.x) / block.x, (s->g.local_rows + block.y - 1) / block.y);
  for (int it = 0; it < iterations; it++)
    for (int colour = 0; colour < 2; colour++) {
      diffusion_half_sweep_kernel<<<grid, block, 0, s->stream>>>(s->g, a, denom, colour, s->u);
      VIS_LAUNCH_CHECK(s, "diffusion_half_sweep_kernel");
    }
  return SAYAL_OK;
}

int launch_render_pixels(Sim* s, uint32_t* d_pixels) {
  dim3 block(64, 4);
  dim3 grid(((s->g.W + 3) / 4 + block.x - 1) / block.x, (s->g.own_hi - s->g.own_lo + block.y - 1) / block.y);
  render_pixels_kernel<<<grid, block, 0, s->stream>>>(s->g, s->flags, s->smoke, s->p, s->d_range + (is_linked(s) ? 2 : 0), s->ph.enable_pressure,
                                                     s->ph.enable_smoke, d_pixels);
  VIS_LAUNCH_CHECK(s, "render_pixels_kernel");
  return SAYAL_OK;
}

int launch_path_lines(Sim* s, int dist, int len, float d_t, int nx, int ny, int32_t* d_xs, int32_t* d_ys) {
  View w{s->u, s->v, s->smoke, s->flags, s->d_overflow};
  dim3 block(32, 4);
  dim3 grid((nx + block.x - 1) / block.x, (ny + block.y - 1) / block.y);
  path_lines_kernel<<<grid, block, 0, s->stream>>>(s->g, w, dist, len, d_t, nx, ny, d_xs, d_ys);
  VIS_LAUNCH_CHECK(s, "path_lines_kernel");
  return SAYAL_OK;
}

int launch_arrows(Sim* s, const sayal_visual* v, int nx, int ny, sayal_arrow* d_out) {
  View w{s->u, s->v, s->smoke, s->flags, s->d_overflow};
  dim3 block(32, 4);
  dim3 grid((nx + block.x - 1) / block.x, (ny + block.y - 1) / block.y);
  arrows_kernel<<<grid, block, 0, s->stream>>>(s->g, w, v->arrows_distance, v->cell_pixel_size, v->arrows_length_multiplier,
                                              v->arrows_disable_threshold, (float)v->arrows_head_length, nx, ny, d_out);
  VIS_LAUNCH_CHECK(s, "arrows_kernel");
  return SAYAL_OK;
}

// Load this file's kernels now: CUDA loads a kernel lazily at its first launch, and that load can wait for the device
// to drain — which never happens while a linked slab on the same device spins for rows this thread has yet to enqueue.
int preload_visual() {
  cudaFuncAttributes fa;
  cudaError_t e = cudaSuccess;
  if (e == cudaSuccess) e = cudaFuncGetAttributes(&fa, render_pixels_kernel);
  if (e == cudaSuccess) e = cudaFuncGetAttributes(&fa, diffusion_half_sweep_kernel);
  if (e == cudaSuccess) e = cudaFuncGetAttributes(&fa, path_lines_kernel);
  if (e == cudaSuccess) e = cudaFuncGetAttributes(&fa, arrows_kernel);
  return e == cudaSuccess ? SAYAL_OK : set_error(SAYAL_ECUDA, cudaGetErrorString(e));
}

}  // namespace sayal
