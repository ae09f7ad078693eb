// kernels_basic.cu — the stages of Fluid::update (/root/reference/src/fluid.cu:770-795) other than the
// temporally blocked projection: cell flags, external forces, plain red/black half-sweeps (used for
// pressure runs and as the in-library cross-check of the tiled kernel), pressure range, boundary
// extrapolation and the fused semi-Lagrangian advection of u, v and smoke (+ decay).
//
// Arithmetic contract: IEEE fp32/fp64 with explicitly placed FMAs — every operation below
// is a round-to-nearest intrinsic, so neither --fmad nor --use_fast_math can change a bit.  The CPU
// oracle (oracle/sayal_oracle.c) states the same sequence; tests require bit equality with it.
#include <cmath>
#include <cstdio>

#include "sayal_internal.h"
#include "advect_common.cuh"

namespace sayal {

#define SAYAL_LAUNCH_CHECK(s, what)                                                     \
  do {                                                                                  \
    cudaError_t e__ = cudaGetLastError();                                               \
    if (e__ != cudaSuccess) {                                                           \
      char m__[256];                                                                    \
      snprintf(m__, sizeof m__, "%s: %s", what, cudaGetErrorString(e__));               \
      return set_error(SAYAL_ECUDA, m__);                                               \
    }                                                                                   \
    (s)->launches++;                                                                    \
  } while (0)

__constant__ float c_inv_s[8] = {0.0f, 1.0f, 0.5f, 1.0f / 3.0f, 0.25f, 0.f, 0.f, 0.f};

static inline dim3 cell_grid(const Grid& g, int rows, dim3 block, int cells_per_thread_x = 1) {
  return dim3((g.W + block.x * cells_per_thread_x - 1) / (block.x * cells_per_thread_x),
              (rows + block.y - 1) / block.y);
}

// ---------------------------------------------------------------------------------------------------
// Cell flags: Fluid::init_device_memory (fluid.cu:99-161), evaluated per cell on the device.
// sqrt/pow on ints in the reference are double precision; dx*dx+dy*dy is exact and __dsqrt_rn is
// correctly rounded, so this matches the host formula bit for bit.
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool solid_formula(const Grid& g, const Phys& p, int i, int j) {
  if (i == 0 || j == 0 || j == g.H - 1) return true;
  if (!p.enable_drain && i == g.W - 1) return true;
  if (p.obstacle_enable) {
    double dx = (double)(i - p.obstacle_cx), dy = (double)(j - p.obstacle_cy);
    if (__dsqrt_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy))) < (double)p.obstacle_radius) return true;
  }
  if (i < p.wt_pipe_length &&
      (j == g.H / 2 - p.wt_pipe_height / 2 - 1 || j == g.H / 2 + p.wt_pipe_height / 2 + 1))
    return true;
  return false;
}

__device__ __forceinline__ bool open_formula(const Grid& g, const Phys& p, int i, int j) {
  return i >= 0 && j >= 0 && i < g.W && j < g.H && !solid_formula(g, p, i, j);
}

__global__ void build_flags_kernel(Grid g, Phys p, uint8_t* __restrict__ flags) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  int lr = blockIdx.y * blockDim.y + threadIdx.y;
  if (i >= g.pitch || lr >= g.local_rows) return;
  uint8_t f = FL_SOLID;  // pad columns read as solid: the advection taps rely on it
  if (i < g.W) {
    f = 0;
    int j = g.H - 1 - (g.row_base + lr);
    bool solid = solid_formula(g, p, i, j);
    if (solid) {
      f = FL_SOLID;
    } else if (i > 0 && j > 0 && i < g.W - 1 && j < g.H - 1) {  // fluid.cu:267, 276
      int l = open_formula(g, p, i - 1, j), r = open_formula(g, p, i + 1, j);
      int b = open_formula(g, p, i, j - 1), t = open_formula(g, p, i, j + 1);
      int ts = l + r + b + t;
      if (ts > 0) f = (uint8_t)((l ? FL_L : 0) | (r ? FL_R : 0) | (b ? FL_B : 0) | (t ? FL_T : 0) | (ts << 4));
    }
  }
  flags[(size_t)lr * g.pitch + i] = f;
}

__global__ void export_masks_kernel(Grid g, Phys p, int32_t* __restrict__ is_solid, int32_t* __restrict__ total_s) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  int lr = blockIdx.y * blockDim.y + threadIdx.y;
  if (i >= g.W || lr >= g.local_rows) return;
  int j = g.H - 1 - (g.row_base + lr);
  size_t k = (size_t)lr * g.W + i;  // exported arrays are dense (reference layout)
  is_solid[k] = solid_formula(g, p, i, j) ? 1 : 0;
  total_s[k] = open_formula(g, p, i - 1, j) + open_formula(g, p, i + 1, j) + open_formula(g, p, i, j - 1) +
               open_formula(g, p, i, j + 1);
}

int launch_build_flags(Sim* s) {
  dim3 block(64, 4);
  dim3 grid((s->g.pitch + 63) / 64, (s->g.local_rows + 3) / 4);
  build_flags_kernel<<<grid, block, 0, s->stream>>>(s->g, s->ph, s->flags);
  SAYAL_LAUNCH_CHECK(s, "build_flags_kernel");
  return SAYAL_OK;
}

int launch_export_masks(Sim* s) {
  dim3 block(64, 4);
  export_masks_kernel<<<cell_grid(s->g, s->g.local_rows, block), block, 0, s->stream>>>(s->g, s->ph, s->d_is_solid,
                                                                                      s->d_total_s);
  SAYAL_LAUNCH_CHECK(s, "export_masks_kernel");
  return SAYAL_OK;
}

// ---------------------------------------------------------------------------------------------------
// External forces: Fluid::apply_external_forces_at (fluid.cu:308-349), every cell (H14).
// ---------------------------------------------------------------------------------------------------

__global__ void forces_kernel(Grid g, Phys p, ForceArgs a, float* __restrict__ u, float* __restrict__ v,
                              float* __restrict__ smoke) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  int lr = blockIdx.y * blockDim.y + threadIdx.y;
  if (i >= g.W || lr >= g.local_rows) return;
  int j = g.H - 1 - (g.row_base + lr);
  size_t k = (size_t)lr * g.pitch + i;
  float vv = v[k];
  bool inlet = i <= p.wt_smoke_length && i != 0 && j >= a.band_lo && j <= a.band_hi;
  bool touch_u = inlet || p.drag_coeff != 0.f || a.src_active;
  float uu = touch_u ? u[k] : 0.f;
  if (inlet) {
    uu = p.wt_speed;
    bool on = p.wt_smoke_count == 1 ? (j >= a.smoke_lo && j <= a.smoke_hi)
                                    : (a.period != 0 && (a.anchor - j) % a.period < p.wt_smoke_height);
    if (on) smoke[k] = p.wt_smoke;
  }
  if (p.drag_coeff != 0.f) {
    uu = __fmul_rn(uu, a.damping);
    vv = __fmul_rn(vv, a.damping);
  }
  if (a.src_active) {
    int dx = i - a.src_x, dy = j - a.src_y;
    if (dx * dx + dy * dy < 1600) {
      if (a.src_smoke != 0.f) smoke[k] = a.src_smoke;
      uu = __fmaf_rn(a.src_velocity, (float)dx, uu);
      vv = __fmaf_rn(a.src_velocity, (float)dy, vv);
    }
  }
  vv = __fmaf_rn(p.g, a.d_t, vv);
  v[k] = vv;
  if (touch_u) u[k] = uu;
}

int launch_forces(Sim* s, const sayal_source* src, float d_t, bool may_fuse) {
  const Phys& p = s->ph;
  const int H = s->g.H;
  ForceArgs a;
  a.band_lo = H / 2 - p.wt_height / 2;
  a.band_hi = H / 2 + p.wt_height / 2;
  a.smoke_lo = H / 2 - p.wt_smoke_height / 2;
  a.smoke_hi = H / 2 + p.wt_smoke_height / 2;
  // fluid.cu:312-315; H11: integer division by zero is only reachable when the value is unused
  int spacing = (p.wt_smoke_count - 1) != 0 ? (p.wt_height - p.wt_smoke_count * p.wt_smoke_height) / (p.wt_smoke_count - 1) : 0;
  a.period = spacing + p.wt_smoke_height;
  a.anchor = H / 2 + p.wt_height / 2;
  a.damping = p.drag_coeff != 0.f ? expf(-p.drag_coeff * d_t) : 1.0f;
  a.src_active = src && src->active;
  a.src_x = src ? src->x : 0;
  a.src_y = src ? src->y : 0;
  a.src_smoke = src ? src->smoke : 0.f;
  a.src_velocity = src ? src->velocity : 0.f;
  a.d_t = d_t;
  // No drag and no interactive source: what is left (v += g d_t, the inlet) is folded into the load of the
  // step's first tiled projection pass (projection_pack.cu) when the caller runs one next.
  if (may_fuse && s->fuse_forces && s->projection_kernel == 1 && !p.enable_pressure && s->cfg.proj_n > 0 &&
      s->cfg.viscosity == 0.f && p.drag_coeff == 0.f && !a.src_active) {
    s->fuse_args = a;
    s->fuse_pending = 1;
    return SAYAL_OK;
  }
  dim3 block(128, 2);
  forces_kernel<<<cell_grid(s->g, s->g.local_rows, block), block, 0, s->stream>>>(s->g, p, a, s->u, s->v, s->smoke);
  SAYAL_LAUNCH_CHECK(s, "forces_kernel");
  return SAYAL_OK;
}

int launch_zero_pressure(Sim* s) {  // Fluid::zero_pressure_at (fluid.cu:212-214)
  cudaError_t e = cudaMemsetAsync(s->p, 0, sizeof(float) * (size_t)s->g.pitch * s->g.local_rows, s->stream);
  if (e != cudaSuccess) return set_error(SAYAL_ECUDA, cudaGetErrorString(e));
  return SAYAL_OK;
}

// ---------------------------------------------------------------------------------------------------
// Plain projection half-sweep: Fluid::apply_projection_at (fluid.cu:229-262), one thread per cell of the
// active colour.  (i + j) even first, then odd (fluid.cu:264-295).
// ---------------------------------------------------------------------------------------------------
__global__ void projection_half_sweep_kernel(Grid g, float o, float density, float inv_dt, int pressure, int colour,
                                             const uint8_t* __restrict__ flags, float* __restrict__ u,
                                             float* __restrict__ v, float* __restrict__ p) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  int lr = blockIdx.y * blockDim.y + threadIdx.y;
  // the cell above (j+1) is local row lr-1: the first local row of a slab cannot be updated here
  if (lr >= g.local_rows || lr < 1) return;
  int j = g.H - 1 - (g.row_base + lr);
  int i = 2 * t + ((j + colour) & 1);
  if (i >= g.W - 1 || i <= 0) return;
  size_t k = (size_t)lr * g.pitch + i;
  unsigned f = flags[k];
  unsigned ts = (f >> 4) & 7u;
  if (ts == 0) return;  // solid, border, or enclosed (H7)
  size_t kt = k - g.pitch;
  float uu = u[k], ur = u[k + 1], vv = v[k], vt = v[kt];
  float d = __fsub_rn(__fadd_rn(__fsub_rn(ur, uu), vt), vv);
  float vd = __fmul_rn(o, __fmul_rn(d, c_inv_s[ts]));
  // update_pressure_at (fluid.cu:225-226): ((vd * density) * cell_size) * (1/d_t), accumulated with one FMA
  if (pressure) p[k] = __fmaf_rn(__fmul_rn(__fmul_rn(vd, density), (float)g.h), inv_dt, p[k]);
  if (f & FL_L) u[k] = __fadd_rn(uu, vd);
  if (f & FL_R) u[k + 1] = __fsub_rn(ur, vd);
  if (f & FL_B) v[k] = __fadd_rn(vv, vd);
  if (f & FL_T) v[kt] = __fsub_rn(vt, vd);
}

int launch_projection_plain(Sim* s, int iterations, float d_t) {
  dim3 block(64, 4);
  dim3 grid((s->g.W / 2 + 1 + block.x - 1) / block.x, (s->g.local_rows + block.y - 1) / block.y);
  float inv_dt = 1.0f / d_t;
  for (int it = 0; it < iterations; it++) {
    for (int colour = 0; colour < 2; colour++) {
      projection_half_sweep_kernel<<<grid, block, 0, s->stream>>>(s->g, s->ph.o, s->ph.density, inv_dt,
                                                                  s->ph.enable_pressure, colour, s->flags, s->u, s->v,
                                                                  s->p);
      SAYAL_LAUNCH_CHECK(s, "projection_half_sweep_kernel");
    }
  }
  return SAYAL_OK;
}

// ---------------------------------------------------------------------------------------------------
// Pressure range: thrust::reduce min / max over all cells (fluid.cu:778-787, H12), as one pass with
// ordered-int atomics.
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ int ordered_int(float f) {
  int b = __float_as_int(f);
  return b ^ ((b >> 31) & 0x7fffffff);
}

__global__ void range_init_kernel(int32_t* range) {
  range[0] = 0x7f800000;                    // +inf
  range[1] = (int)0xff800000 ^ 0x7fffffff;  // ordered(-inf)
}

__global__ void pressure_range_kernel(Grid g, const float* __restrict__ p, int32_t* range) {
  int mn = 0x7f800000, mx = (int)0xff800000 ^ 0x7fffffff;
  size_t n = (size_t)(g.own_hi - g.own_lo) * g.W;
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (size_t)gridDim.x * blockDim.x) {
    int lr = g.own_lo + (int)(t / g.W), i = (int)(t % g.W);
    float x = p[(size_t)lr * g.pitch + i];
    if (x == x) {
      int e = ordered_int(x);
      mn = min(mn, e);
      mx = max(mx, e);
    }
  }
  for (int off = 16; off > 0; off >>= 1) {
    mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, off));
    mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, off));
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMin(&range[0], mn);
    atomicMax(&range[1], mx);
  }
}

int launch_pressure_range(Sim* s) {
  range_init_kernel<<<1, 1, 0, s->stream>>>(s->d_range);
  SAYAL_LAUNCH_CHECK(s, "range_init_kernel");
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, s->device);
  pressure_range_kernel<<<sms * 4, 256, 0, s->stream>>>(s->g, s->p, s->d_range);
  SAYAL_LAUNCH_CHECK(s, "pressure_range_kernel");
  return SAYAL_OK;
}

// ---------------------------------------------------------------------------------------------------
// Boundary extrapolation: Fluid::apply_extrapolation_at (fluid.cu:720-733) in its canonical order
// (all j-rules, then all i-rules; H4), written in closed form so that no thread reads a value another
// thread writes:   u(1,j)=0;  u(i,0)=u(i,1), u(i,H-1)=u(i,H-2) for i!=1;
//                  v(i,1)=0;  v(0,j)=v(1,j), v(W-1,j)=v(W-2,j) for j!=1.
// ---------------------------------------------------------------------------------------------------
__global__ void extrapolation_kernel(Grid g, float* __restrict__ u, float* __restrict__ v) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int W = g.W, H = g.H;
  if (t < W) {
    int i = t;
    int lr0 = (H - 1) - g.row_base;      // row j = 0
    int lr1 = (H - 2) - g.row_base;      // row j = 1
    int lrT = 0 - g.row_base;            // row j = H-1
    int lrT1 = 1 - g.row_base;           // row j = H-2
    if (lr0 >= 0 && lr0 < g.local_rows && lr1 >= 0 && lr1 < g.local_rows)
      u[(size_t)lr0 * g.pitch + i] = (i == 1) ? 0.f : u[(size_t)lr1 * g.pitch + i];
    if (lr1 >= 0 && lr1 < g.local_rows) v[(size_t)lr1 * g.pitch + i] = 0.f;
    if (lrT >= 0 && lrT < g.local_rows && lrT1 >= 0 && lrT1 < g.local_rows)
      u[(size_t)lrT * g.pitch + i] = (i == 1) ? 0.f : u[(size_t)lrT1 * g.pitch + i];
  } else if (t < W + g.local_rows) {
    int lr = t - W;
    int j = H - 1 - (g.row_base + lr);
    size_t row = (size_t)lr * g.pitch;
    u[row + 1] = 0.f;
    v[row + 0] = (j == 1) ? 0.f : v[row + 1];
    v[row + W - 1] = (j == 1) ? 0.f : v[row + W - 2];
  }
}

int launch_extrapolation(Sim* s) {
  int n = s->g.W + s->g.local_rows;
  extrapolation_kernel<<<(n + 255) / 256, 256, 0, s->stream>>>(s->g, s->u, s->v);
  SAYAL_LAUNCH_CHECK(s, "extrapolation_kernel");
  return SAYAL_OK;
}

// ---------------------------------------------------------------------------------------------------
// Semi-Lagrangian advection (fluid.cu:364-716).  Velocity advection writes u', v' and smoke advection
// (+decay) writes smoke' into the back buffers; the host swaps pointers instead of running the reference's
// copy-back kernels (fluid.cu:614-617, 569-571).
//
// Two things keep the instruction count down without changing a bit of the result:
//   * HC = 1 specialises cell_size == 1 (every shipped config): x / 1.0f == x, (k + 0.5) * 1 == k + 0.5, so
//     all divisions by the cell size disappear.  HC = 0 keeps the general integer cell size.
//   * Once the base cell of a sample is a fluid cell it is an interior cell (walls are always solid), so
//     its 3x3 neighbourhood is inside the array; the one exception, column W when the drain is open, reads
//     a pad byte (or column 0 of the next row) that build_flags marks solid.  Taps therefore need no
//     bounds test, only the solid bit.
// ---------------------------------------------------------------------------------------------------
template <int HC>
__global__ void __launch_bounds__(256)
advect_velocity_kernel(Grid g, View w, float d_t, float* __restrict__ u_out, float* __restrict__ v_out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  int lr = g.own_lo + blockIdx.y * blockDim.y + threadIdx.y;
  if (i >= g.W || lr >= g.own_hi) return;
  int r = g.row_base + lr;
  int j = g.H - 1 - r;
  long k = (long)lr * g.pitch + i;
  const long up = -(long)g.pitch, down = (long)g.pitch;
  // which neighbours exist at all (index_is_valid, fluid.cu:356-358); rows must also be held locally
  const bool has_l = i > 0, has_r = i < g.W - 1;
  const bool has_up = j < g.H - 1 && lr > g.valid_lo, has_dn = j > 0 && lr < g.valid_hi - 1;
  if ((j < g.H - 1 && lr == g.valid_lo) || (j > 0 && lr == g.valid_hi - 1)) atomicAdd(w.overflow, 1);
  float uk = w.u[k], vk = w.v[k];
  // get_vertical_edge_velocity (fluid.cu:364-389)
  float avg_v = vk;
  int count = 1;
  if (has_l && has_up && open_tap(w, k - 1 + up)) { avg_v = __fadd_rn(avg_v, w.v[k - 1 + up]); count++; }
  if (has_up && open_tap(w, k + up)) { avg_v = __fadd_rn(avg_v, w.v[k + up]); count++; }
  if (has_l && open_tap(w, k - 1)) { avg_v = __fadd_rn(avg_v, w.v[k - 1]); count++; }
  avg_v = div_count(avg_v, count);
  float px = __fmaf_rn(-uk, d_t, mul_h<HC>(i, g.h));
  float py = __fmaf_rn(-avg_v, d_t, pos_half<HC>(j, g.h));
  u_out[k] = general_velocity_x<HC>(g, w, px, py);
  // get_horizontal_edge_velocity (fluid.cu:391-416)
  float avg_u = uk;
  count = 1;
  if (has_r && open_tap(w, k + 1)) { avg_u = __fadd_rn(avg_u, w.u[k + 1]); count++; }
  if (has_dn && open_tap(w, k + down)) { avg_u = __fadd_rn(avg_u, w.u[k + down]); count++; }
  if (has_r && has_dn && open_tap(w, k + 1 + down)) { avg_u = __fadd_rn(avg_u, w.u[k + 1 + down]); count++; }
  avg_u = div_count(avg_u, count);
  px = __fmaf_rn(-avg_u, d_t, pos_half<HC>(i, g.h));
  py = __fmaf_rn(-vk, d_t, mul_h<HC>(j, g.h));
  v_out[k] = general_velocity_y<HC>(g, w, px, py);
}

template <int HC>
__global__ void __launch_bounds__(256)
advect_smoke_kernel(Grid g, View w, float d_t, int enable_decay, float decay_rate, float* __restrict__ smoke_out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  int lr = g.own_lo + blockIdx.y * blockDim.y + threadIdx.y;
  if (i >= g.W || lr >= g.own_hi) return;
  int j = g.H - 1 - (g.row_base + lr);
  long k = (long)lr * g.pitch + i;
  // apply_smoke_advection_at (fluid.cu:560-567): runs on the velocity field AFTER velocity advection
  float cx = pos_half<HC>(i, g.h), cy = pos_half<HC>(j, g.h);
  float vx = general_velocity_x<HC>(g, w, cx, cy), vy = general_velocity_y<HC>(g, w, cx, cy);
  float sm = interpolate_smoke<HC>(g, w, __fmaf_rn(-vx, d_t, cx), __fmaf_rn(-vy, d_t, cy));
  if (enable_decay) {  // decay_smoke_at (fluid.cu:758-762)
    float t = __fmaf_rn(-decay_rate, d_t, sm);
    sm = (float)fmax((double)t, 0.0);
  }
  smoke_out[k] = sm;
}

int launch_advect(Sim* s, float d_t, bool velocity, bool smoke) {
  dim3 block(64, 4);
  dim3 grid((s->g.W + block.x - 1) / block.x, (s->g.own_hi - s->g.own_lo + block.y - 1) / block.y);
  View w{s->u, s->v, s->smoke, s->flags, s->d_overflow};
  if (velocity) {
    if (s->g.h == 1) advect_velocity_kernel<1><<<grid, block, 0, s->stream>>>(s->g, w, d_t, s->u_buf, s->v_buf);
    else advect_velocity_kernel<0><<<grid, block, 0, s->stream>>>(s->g, w, d_t, s->u_buf, s->v_buf);
    SAYAL_LAUNCH_CHECK(s, "advect_velocity_kernel");
  }
  if (smoke) {
    if (s->g.h == 1)
      advect_smoke_kernel<1><<<grid, block, 0, s->stream>>>(s->g, w, d_t, s->ph.enable_decay, s->ph.decay_rate, s->smoke_buf);
    else
      advect_smoke_kernel<0><<<grid, block, 0, s->stream>>>(s->g, w, d_t, s->ph.enable_decay, s->ph.decay_rate, s->smoke_buf);
    SAYAL_LAUNCH_CHECK(s, "advect_smoke_kernel");
  }
  return SAYAL_OK;
}

// Fluid::get_general_velocity (fluid.cu:541-545) at arbitrary points
__global__ void sample_velocity_kernel(Grid g, View w, int n, const float* __restrict__ xs, const float* __restrict__ ys,
                                       float* __restrict__ ou, float* __restrict__ ov) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  ou[t] = general_velocity_x<0>(g, w, xs[t], ys[t]);
  ov[t] = general_velocity_y<0>(g, w, xs[t], ys[t]);
}

int launch_sample_velocity(Sim* s, int n, const float* d_xs, const float* d_ys, float* d_ou, float* d_ov) {
  View w{s->u, s->v, s->smoke, s->flags, s->d_overflow};
  sample_velocity_kernel<<<(n + 127) / 128, 128, 0, s->stream>>>(s->g, w, n, d_xs, d_ys, d_ou, d_ov);
  SAYAL_LAUNCH_CHECK(s, "sample_velocity_kernel");
  return SAYAL_OK;
}

// ---------------------------------------------------------------------------------------------------
// Slab edge rows <-> packed device buffer (U, V, SMOKE blocks of nrows*W floats, in field_mask order).
// ---------------------------------------------------------------------------------------------------
__global__ void pack_rows_kernel(Grid g, int local_row0, int nrows, int nfields, float* f0, float* f1, float* f2,
                                 float* __restrict__ buf, int unpack) {
  size_t per = (size_t)nrows * g.W;
  size_t total = per * nfields;
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
    int f = (int)(t / per);
    size_t r = t % per;
    int row = (int)(r / g.W), i = (int)(r % g.W);
    float* field = f == 0 ? f0 : (f == 1 ? f1 : f2);
    size_t k = (size_t)(local_row0 + row) * g.pitch + i;
    if (unpack) field[k] = buf[t];
    else buf[t] = field[k];
  }
}

int launch_pack_rows(Sim* s, int local_row0, int nrows, int field_mask, float* dev_buf, bool unpack) {
  float* fields[3];
  int nf = 0;
  if (field_mask & 1) fields[nf++] = s->u;
  if (field_mask & 2) fields[nf++] = s->v;
  if (field_mask & 4) fields[nf++] = s->smoke;
  if (field_mask & 8) fields[nf++] = s->p;
  if (nf == 0 || nf > 3) return set_error(SAYAL_EINVAL, "pack_rows: field_mask must select 1..3 fields");
  if (local_row0 < 0 || nrows <= 0 || local_row0 + nrows > s->g.local_rows)
    return set_error(SAYAL_EINVAL, "pack_rows: row range outside the slab");
  for (int k = nf; k < 3; k++) fields[k] = fields[0];
  size_t total = (size_t)nrows * s->g.W * nf;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  pack_rows_kernel<<<blocks, 256, 0, s->stream>>>(s->g, local_row0, nrows, nf, fields[0], fields[1], fields[2], dev_buf,
                                                  unpack ? 1 : 0);
  SAYAL_LAUNCH_CHECK(s, "pack_rows_kernel");
  return SAYAL_OK;
}

// Load this file's kernels now: CUDA loads a kernel lazily at its first launch, and that load can wait for the device
// to drain — which never happens while a linked slab on the same device spins for rows this thread has yet to enqueue.
int preload_basic() {
  cudaFuncAttributes fa;
  cudaError_t e = cudaSuccess;
  if (e == cudaSuccess) e = cudaFuncGetAttributes(&fa, forces_kernel);
  if (e == cudaSuccess) e = cudaFuncGetAttributes(&fa, projection_half_sweep_kernel);
  if (e == cudaSuccess) e = cudaFuncGetAttributes(&fa, range_init_kernel);
  if (e == cudaSuccess) e = cudaFuncGetAttributes(&fa, pressure_range_kernel);
  if (e == cudaSuccess) e = cudaFuncGetAttributes(&fa, extrapolation_kernel);
  if (e == cudaSuccess) e = cudaFuncGetAttributes(&fa, advect_velocity_kernel<0>);
  if (e == cudaSuccess) e = cudaFuncGetAttributes(&fa, advect_velocity_kernel<1>);
  if (e == cudaSuccess) e = cudaFuncGetAttributes(&fa, advect_smoke_kernel<0>);
  if (e == cudaSuccess) e = cudaFuncGetAttributes(&fa, advect_smoke_kernel<1>);
  if (e == cudaSuccess) e = cudaFuncGetAttributes(&fa, sample_velocity_kernel);
  if (e == cudaSuccess) e = cudaFuncGetAttributes(&fa, pack_rows_kernel);
  if (e == cudaSuccess) e = cudaFuncGetAttributes(&fa, build_flags_kernel);
  if (e == cudaSuccess) e = cudaFuncGetAttributes(&fa, export_masks_kernel);
  return e == cudaSuccess ? SAYAL_OK : set_error(SAYAL_ECUDA, cudaGetErrorString(e));
}

}  // namespace sayal
