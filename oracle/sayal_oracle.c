/*
 * sayal_oracle.c — CPU restatement of OpenSayal's per-step simulation path, Fluid::update()
 * (/root/reference/src/fluid.cu:770-795).
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / reference legs may load it.  The shipped path is the CUDA library under
 * opensayal_b200/csrc and never calls into this file.
 *
 * Pin status: the reference holds no tests, golden vectors or known-answer fixtures for this path
 *  — "parity unpinned" by vectors.  The pin is execution of the reference's own
 * CUDA code (oracle/_ref, built from /root/reference/src by oracle/Makefile) on the GPU box:
 * tests/test_reference_parity.py compares this restatement with it (flags bit-exact, fields within
 * 1e-5 relative L2 per step), and the mask counts derived in SURVEY.md §8a-M are checked as
 * known answers in tests/test_oracle.py.
 *
 * Arithmetic: IEEE fp32 / fp64 exactly where the source has float / double, with the fused
 * multiply-adds nvcc's default contraction produces written out as fmaf() and every other
 * contraction disabled (-ffp-contract=off).  The reference's Release build additionally uses
 * --use_fast_math (CMakeLists.txt:21), so it differs from this file by a few ulp per operation
 * (MUFU.RCP division, approximate sqrt): that is the 1e-5 tolerance, see DESIGN.md.
 * Two deliberate restatements of that fast-math binary, both exact to <= 1 ulp of the source:
 *   - projection divides by total_s through a reciprocal table (the SASS is `d * MUFU.RCP(s)`),
 *   - pressure accumulates with one FMA on a host-computed 1/d_t (SASS: FFMA with rcp(d_t)).
 *
 * Layout: every array is W*H in the reference's order, index(i,j) = (H-1-j)*W + i (fluid.cu:163).
 */
#define _GNU_SOURCE
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "sayal.h"

typedef struct oracle_sim {
  sayal_config c;
  int W, H, h; /* h = (int)cell_size, Fluid::cell_size is int (fluid.cuh:46) */
  float *u, *v, *p, *smoke, *u_buf, *v_buf, *smoke_buf;
  int32_t *is_solid, *total_s;
  float min_p, max_p;
  int threads;
} oracle_sim;

#define IDX(s, i, j) ((size_t)((s)->H - 1 - (j)) * (size_t)(s)->W + (size_t)(i))

/* float -> int as the GPU's cvt.rzi.s32.f32 does it (saturating, NaN -> 0); plain C casts are UB
 * out of range.  Used for `int i = x / cell_size` (fluid.cu:419-420, 480-481, 645-646). */
static inline int f2i(float x) {
  if (x != x) return 0;
  if (x >= 2147483648.0f) return INT32_MAX;
  if (x <= -2147483648.0f) return INT32_MIN;
  return (int)x;
}

/* Fluid::index_is_valid (fluid.cu:356-358) */
static inline int index_is_valid(const oracle_sim* s, int i, int j) {
  return i < s->W && j < s->H && i >= 0 && j >= 0;
}
/* Fluid::is_valid_fluid (fluid.cu:360-362) */
static inline int is_valid_fluid(const oracle_sim* s, int i, int j) {
  return index_is_valid(s, i, j) && !s->is_solid[IDX(s, i, j)];
}

/* ---- masks: Fluid::init_device_memory (fluid.cu:99-161) -------------------------------------- */
void oracle_build_masks(const sayal_config* c, int32_t* is_solid, int32_t* total_s) {
  const int W = c->width, H = c->height;
  for (int j = 0; j < H; j++) {
    for (int i = 0; i < W; i++) {
      /* fluid.cu:113-124.  std::pow / std::sqrt on ints and a float radius: evaluated in double. */
      double dx = (double)(i - c->obstacle_center_x), dy = (double)(j - c->obstacle_center_y);
      int disc = c->obstacle_enable && sqrt(pow(dx, 2) + pow(dy, 2)) < (double)c->obstacle_radius;
      int pipe = i < c->wt_pipe_length && (j == H / 2 - c->wt_pipe_height / 2 - 1 ||
                                           j == H / 2 + c->wt_pipe_height / 2 + 1);
      int solid = i == 0 || j == 0 || j == H - 1 || (!c->enable_drain && i == W - 1) || disc || pipe;
      is_solid[(size_t)(H - 1 - j) * W + i] = solid;
    }
  }
  /* fluid.cu:127-142; H2: the reference increments malloc'd memory, the specification is a count */
  for (int j = 0; j < H; j++) {
    for (int i = 0; i < W; i++) {
      int n = 0;
      if (i - 1 >= 0 && !is_solid[(size_t)(H - 1 - j) * W + (i - 1)]) n++;
      if (i + 1 < W && !is_solid[(size_t)(H - 1 - j) * W + (i + 1)]) n++;
      if (j - 1 >= 0 && !is_solid[(size_t)(H - 1 - (j - 1)) * W + i]) n++;
      if (j + 1 < H && !is_solid[(size_t)(H - 1 - (j + 1)) * W + i]) n++;
      total_s[(size_t)(H - 1 - j) * W + i] = n;
    }
  }
}

/* ---- lifetime ------------------------------------------------------------------------------- */
oracle_sim* oracle_create(const sayal_config* c) {
  oracle_sim* s = (oracle_sim*)calloc(1, sizeof(oracle_sim));
  if (!s) return NULL;
  s->c = *c;
  s->W = c->width;
  s->H = c->height;
  s->h = (int)c->cell_size;
  s->threads = 1;
  size_t n = (size_t)s->W * s->H;
  s->u = (float*)calloc(n, 4);
  s->v = (float*)calloc(n, 4);
  s->p = (float*)calloc(n, 4);
  s->smoke = (float*)calloc(n, 4);
  s->u_buf = (float*)calloc(n, 4);
  s->v_buf = (float*)calloc(n, 4);
  s->smoke_buf = (float*)calloc(n, 4);
  s->is_solid = (int32_t*)calloc(n, 4);
  s->total_s = (int32_t*)calloc(n, 4);
  oracle_build_masks(c, s->is_solid, s->total_s);
  return s;
}

void oracle_destroy(oracle_sim* s) {
  if (!s) return;
  free(s->u); free(s->v); free(s->p); free(s->smoke);
  free(s->u_buf); free(s->v_buf); free(s->smoke_buf);
  free(s->is_solid); free(s->total_s);
  free(s);
}

void oracle_set_threads(oracle_sim* s, int t) { s->threads = t < 1 ? 1 : t; }

void* oracle_field(oracle_sim* s, int field) {
  switch (field) {
    case SAYAL_U: return s->u;
    case SAYAL_V: return s->v;
    case SAYAL_P: return s->p;
    case SAYAL_SMOKE: return s->smoke;
    case SAYAL_IS_SOLID: return s->is_solid;
    case SAYAL_TOTAL_S: return s->total_s;
  }
  return NULL;
}

/* ---- forces: Fluid::apply_external_forces_at (fluid.cu:308-349) ------------------------------- */
void oracle_forces(oracle_sim* s, const sayal_source* src, float d_t) {
  const sayal_config* c = &s->c;
  const int W = s->W, H = s->H;
  const int ph = c->wt_pipe_height, sh = c->wt_smoke_height, cnt = c->wt_smoke_count;
  /* fluid.cu:312-315.  H11: the GPU's integer division by zero does not trap; the value is only
   * consumed when cnt != 1, so 0 is a faithful guard. */
  const int smoke_spacing = (cnt - 1) != 0 ? (ph - cnt * sh) / (cnt - 1) : 0;
  const int period = smoke_spacing + sh;
  /* expf is evaluated once on the host and applied per cell (fluid.cu:332) */
  const float damping = c->drag_coeff != 0 ? expf(-c->drag_coeff * d_t) : 1.0f;
  const int active = src && src->active;
#pragma omp parallel for num_threads(s->threads) schedule(static)
  for (int j = 0; j < H; j++) {
    for (int i = 0; i < W; i++) {
      size_t k = IDX(s, i, j);
      if (i <= c->wt_smoke_length && i != 0 && j >= H / 2 - ph / 2 && j <= H / 2 + ph / 2) {
        s->u[k] = c->wt_speed;
        if ((cnt == 1 && j >= H / 2 - sh / 2 && j <= H / 2 + sh / 2) ||
            (cnt != 1 && period != 0 && (H / 2 + ph / 2 - j) % period < sh)) {
          s->smoke[k] = c->wt_smoke;
        }
      }
      if (c->drag_coeff != 0) {
        s->u[k] *= damping;
        s->v[k] *= damping;
      }
      if (active) {
        int dx = i - src->x, dy = j - src->y;
        if (dx * dx + dy * dy < 40 * 40) {
          if (src->smoke != 0) s->smoke[k] = src->smoke;
          s->u[k] = fmaf(src->velocity, (float)dx, s->u[k]);
          s->v[k] = fmaf(src->velocity, (float)dy, s->v[k]);
        }
      }
      s->v[k] = fmaf(c->g, d_t, s->v[k]); /* fluid.cu:348, contracted by nvcc */
    }
  }
}

/* ---- Fluid::zero_pressure_at (fluid.cu:212-214) ------------------------------------------------ */
void oracle_zero_pressure(oracle_sim* s) { memset(s->p, 0, (size_t)s->W * s->H * 4); }

/* ---- projection: Fluid::apply_projection_at (fluid.cu:229-262), host loop (fluid.cu:282-295) -- */
static const float INV_S[5] = {0.0f, 1.0f, 0.5f, 1.0f / 3.0f, 0.25f};

static void projection_half_sweep(oracle_sim* s, int colour, float d_t) {
  const int W = s->W, H = s->H;
  const float o = s->c.proj_o;
  const int pressure = s->c.enable_pressure;
  const float inv_dt = 1.0f / d_t;
  const float dens = s->c.density, hf = (float)s->h;
#pragma omp parallel for num_threads(s->threads) schedule(static)
  for (int j = 1; j < H - 1; j++) {
    /* even kernel: i = 2*t + (j%2) (fluid.cu:266) => (i+j) even; odd kernel (fluid.cu:275) => odd */
    int i0 = ((j + colour) & 1);
    for (int i = i0; i < W - 1; i += 2) {
      if (i <= 0) continue;
      size_t k = IDX(s, i, j);
      if (s->is_solid[k]) continue;
      size_t kr = IDX(s, i + 1, j), kt = IDX(s, i, j + 1);
      float u = s->u[k], v = s->v[k], top_v = s->v[kt], right_u = s->u[kr];
      float divergence = ((right_u - u) + top_v) - v; /* fluid.cu:239, left to right */
      int ts = s->total_s[k];
      if (ts <= 0 || ts > 4) continue; /* H7: enclosed cell, reference would produce inf/NaN */
      float velocity_diff = o * (divergence * INV_S[ts]); /* fluid.cu:241 */
      if (pressure) { /* fluid.cu:225-226 */
        s->p[k] = fmaf((velocity_diff * dens) * hf, inv_dt, s->p[k]);
      }
      if (!s->is_solid[IDX(s, i - 1, j)]) s->u[k] += velocity_diff;
      if (!s->is_solid[kr]) s->u[kr] -= velocity_diff;
      if (!s->is_solid[IDX(s, i, j - 1)]) s->v[k] += velocity_diff;
      if (!s->is_solid[kt]) s->v[kt] -= velocity_diff;
    }
  }
}

void oracle_projection(oracle_sim* s, int iterations, float d_t) {
  for (int it = 0; it < iterations; it++) {
    projection_half_sweep(s, 0, d_t);
    projection_half_sweep(s, 1, d_t);
  }
}

/* ---- thrust::reduce min / max over ALL cells (fluid.cu:778-787, H12) ------------------------- */
void oracle_pressure_range(oracle_sim* s) {
  float mn = INFINITY, mx = -INFINITY;
  size_t n = (size_t)s->W * s->H;
  for (size_t k = 0; k < n; k++) {
    if (s->p[k] < mn) mn = s->p[k];
    if (s->p[k] > mx) mx = s->p[k];
  }
  s->min_p = mn;
  s->max_p = mx;
}
void oracle_get_pressure_range(oracle_sim* s, float* mn, float* mx) { *mn = s->min_p; *mx = s->max_p; }

/* ---- extrapolation: Fluid::apply_extrapolation_at (fluid.cu:720-733) --------------------------
 * The reference kernel races at four faces (H4).  Canonical order: all j-rules, then all i-rules. */
void oracle_extrapolation(oracle_sim* s) {
  const int W = s->W, H = s->H;
  for (int i = 0; i < W; i++) {
    s->u[IDX(s, i, 0)] = s->u[IDX(s, i, 1)];
    s->v[IDX(s, i, 1)] = 0;
    s->u[IDX(s, i, H - 1)] = s->u[IDX(s, i, H - 2)];
  }
  for (int j = 0; j < H; j++) {
    s->v[IDX(s, 0, j)] = s->v[IDX(s, 1, j)];
    s->u[IDX(s, 1, j)] = 0;
    s->v[IDX(s, W - 1, j)] = s->v[IDX(s, W - 2, j)];
  }
}

/* ---- Fluid::get_vertical_edge_velocity (fluid.cu:364-389): velocity at the u-face of (i,j) ---- */
static inline void vertical_edge_velocity(const oracle_sim* s, int i, int j, float* ou, float* ov) {
  float avg_v = s->v[IDX(s, i, j)];
  int count = 1;
  if (is_valid_fluid(s, i - 1, j + 1)) { avg_v += s->v[IDX(s, i - 1, j + 1)]; count++; }
  if (is_valid_fluid(s, i, j + 1)) { avg_v += s->v[IDX(s, i, j + 1)]; count++; }
  if (is_valid_fluid(s, i - 1, j)) { avg_v += s->v[IDX(s, i - 1, j)]; count++; }
  *ou = s->u[IDX(s, i, j)];
  *ov = avg_v / (float)count;
}

/* ---- Fluid::get_horizontal_edge_velocity (fluid.cu:391-416): velocity at the v-face of (i,j) -- */
static inline void horizontal_edge_velocity(const oracle_sim* s, int i, int j, float* ou, float* ov) {
  float avg_u = s->u[IDX(s, i, j)];
  int count = 1;
  if (is_valid_fluid(s, i + 1, j)) { avg_u += s->u[IDX(s, i + 1, j)]; count++; }
  if (is_valid_fluid(s, i, j - 1)) { avg_u += s->u[IDX(s, i, j - 1)]; count++; }
  if (is_valid_fluid(s, i + 1, j - 1)) { avg_u += s->u[IDX(s, i + 1, j - 1)]; count++; }
  *ou = avg_u / (float)count;
  *ov = s->v[IDX(s, i, j)];
}

/* ---- Fluid::get_general_velocity_y (fluid.cu:418-477) ---------------------------------------- */
static float general_velocity_y(const oracle_sim* s, float x, float y) {
  const float hf = (float)s->h;
  const double half = (double)s->h / 2.0;
  int i = f2i(x / hf), j = f2i(y / hf);
  if (!is_valid_fluid(s, i, j)) return 0;
  float in_x = x - (float)(i * s->h);
  float in_y = y - (float)(j * s->h);
  float avg_v = 0;
  float w_y = 1.0f - in_y / hf;
  if ((double)in_x < half) { /* columns (i, i-1); fluid.cu:432-452 */
    float d_x = (float)(half - (double)in_x);
    float w_x = 1.0f - d_x / hf;
    avg_v = fmaf(w_y * w_x, s->v[IDX(s, i, j)], avg_v);
    if (is_valid_fluid(s, i - 1, j)) avg_v = fmaf(w_y * (1.0f - w_x), s->v[IDX(s, i - 1, j)], avg_v);
    if (is_valid_fluid(s, i - 1, j + 1))
      avg_v = fmaf((1.0f - w_y) * (1.0f - w_x), s->v[IDX(s, i - 1, j + 1)], avg_v);
    if (is_valid_fluid(s, i, j + 1)) avg_v = fmaf((1.0f - w_y) * w_x, s->v[IDX(s, i, j + 1)], avg_v);
  } else { /* columns (i, i+1); fluid.cu:454-474 */
    float d_x = (float)((double)in_x - half);
    float w_x = 1.0f - d_x / hf;
    avg_v = fmaf(w_y * w_x, s->v[IDX(s, i, j)], avg_v);
    if (is_valid_fluid(s, i, j + 1)) avg_v = fmaf((1.0f - w_y) * w_x, s->v[IDX(s, i, j + 1)], avg_v);
    if (is_valid_fluid(s, i + 1, j + 1))
      avg_v = fmaf((1.0f - w_y) * (1.0f - w_x), s->v[IDX(s, i + 1, j + 1)], avg_v);
    if (is_valid_fluid(s, i + 1, j)) avg_v = fmaf(w_y * (1.0f - w_x), s->v[IDX(s, i + 1, j)], avg_v);
  }
  return avg_v;
}

/* ---- Fluid::get_general_velocity_x (fluid.cu:479-539) ---------------------------------------- */
static float general_velocity_x(const oracle_sim* s, float x, float y) {
  const float hf = (float)s->h;
  const double half = (double)s->h / 2.0;
  int i = f2i(x / hf), j = f2i(y / hf);
  if (!is_valid_fluid(s, i, j)) return 0;
  float in_x = x - (float)(i * s->h);
  float in_y = y - (float)(j * s->h);
  float avg_u = 0;
  float w_x = 1.0f - in_x / hf;
  if ((double)in_y <= half) { /* rows (j, j-1); fluid.cu:493-513 — note <= here, < in _y */
    float d_y = (float)(half - (double)in_y);
    float w_y = 1.0f - d_y / hf;
    avg_u = fmaf(w_y * w_x, s->u[IDX(s, i, j)], avg_u);
    if (is_valid_fluid(s, i + 1, j)) avg_u = fmaf(w_y * (1.0f - w_x), s->u[IDX(s, i + 1, j)], avg_u);
    if (is_valid_fluid(s, i, j - 1)) avg_u = fmaf((1.0f - w_y) * w_x, s->u[IDX(s, i, j - 1)], avg_u);
    if (is_valid_fluid(s, i + 1, j - 1))
      avg_u = fmaf((1.0f - w_y) * (1.0f - w_x), s->u[IDX(s, i + 1, j - 1)], avg_u);
  } else { /* rows (j, j+1); fluid.cu:516-536 */
    float d_y = (float)((double)in_y - half);
    float w_y = 1.0f - d_y / hf;
    avg_u = fmaf(w_y * w_x, s->u[IDX(s, i, j)], avg_u);
    if (is_valid_fluid(s, i, j + 1)) avg_u = fmaf((1.0f - w_y) * w_x, s->u[IDX(s, i, j + 1)], avg_u);
    if (is_valid_fluid(s, i + 1, j)) avg_u = fmaf(w_y * (1.0f - w_x), s->u[IDX(s, i + 1, j)], avg_u);
    if (is_valid_fluid(s, i + 1, j + 1))
      avg_u = fmaf((1.0f - w_y) * (1.0f - w_x), s->u[IDX(s, i + 1, j + 1)], avg_u);
  }
  return avg_u;
}

/* Fluid::get_general_velocity (fluid.cu:541-545), exported for sampler tests */
void oracle_sample_velocity(oracle_sim* s, int n, const float* xs, const float* ys, float* ou, float* ov) {
  for (int k = 0; k < n; k++) {
    ou[k] = general_velocity_x(s, xs[k], ys[k]);
    ov[k] = general_velocity_y(s, xs[k], ys[k]);
  }
}

/* positions: fluid.cu:547-558; (k + 0.5) * cell_size is exact in fp32 for every grid we accept */
static inline float pos_half(int k, int h) { return ((float)k + 0.5f) * (float)h; }
static inline float pos_int(int k, int h) { return (float)(k * h); }

/* ---- Fluid::apply_velocity_advection_at (fluid.cu:598-612), all cells (H14) ------------------ */
void oracle_advect_velocity(oracle_sim* s, float d_t) {
  const int W = s->W, H = s->H;
#pragma omp parallel for num_threads(s->threads) schedule(static)
  for (int j = 0; j < H; j++) {
    for (int i = 0; i < W; i++) {
      float eu, ev;
      vertical_edge_velocity(s, i, j, &eu, &ev);
      float px = fmaf(-eu, d_t, pos_int(i, s->h)), py = fmaf(-ev, d_t, pos_half(j, s->h));
      s->u_buf[IDX(s, i, j)] = general_velocity_x(s, px, py);
      horizontal_edge_velocity(s, i, j, &eu, &ev);
      px = fmaf(-eu, d_t, pos_half(i, s->h));
      py = fmaf(-ev, d_t, pos_int(j, s->h));
      s->v_buf[IDX(s, i, j)] = general_velocity_y(s, px, py);
    }
  }
  /* update_velocity_advection_at (fluid.cu:614-617): copy back == swap */
  float* t = s->u; s->u = s->u_buf; s->u_buf = t;
  t = s->v; s->v = s->v_buf; s->v_buf = t;
}

/* ---- Fluid::interpolate_smoke (fluid.cu:644-716) ---------------------------------------------- */
static float interpolate_smoke(const oracle_sim* s, float x, float y) {
  const float hf = (float)s->h;
  const double half = (double)s->h / 2.0;
  int i = f2i(x / hf), j = f2i(y / hf);
  float in_x = x - (float)(i * s->h);
  float in_y = y - (float)(j * s->h);
  int di = ((double)in_x < half) ? -1 : 1;
  int dj = ((double)in_y < half) ? -1 : 1;
  /* taps in the reference's order: (i,j), (i+di,j), (i,j+dj), (i+di,j+dj) (fluid.cu:651-674) */
  int ti[4] = {i, i + di, i, i + di};
  int tj[4] = {j, j, j + dj, j + dj};
  float inv[4];
  for (int k = 0; k < 4; k++) {
    float cx = pos_half(ti[k], s->h), cy = pos_half(tj[k], s->h);
    float ddx = x - cx, ddy = y - cy;
    float dist = sqrtf(fmaf(ddx, ddx, ddy * ddy)); /* helper.cuh:77-81 */
    inv[k] = (float)(1.0 / ((double)dist + 1e-6));   /* fluid.cu:690-693, FP64 */
  }
  float sum_inv = ((inv[0] + inv[1]) + inv[2]) + inv[3];
  float avg = 0;
  for (int k = 0; k < 4; k++) {
    float w = inv[k] / sum_inv;
    if (is_valid_fluid(s, ti[k], tj[k])) avg = fmaf(w, s->smoke[IDX(s, ti[k], tj[k])], avg);
  }
  return avg;
}

/* ---- Fluid::apply_smoke_advection_at (fluid.cu:560-567) + decay_smoke_at (fluid.cu:758-762) --- */
void oracle_advect_smoke(oracle_sim* s, float d_t) {
  const int W = s->W, H = s->H;
#pragma omp parallel for num_threads(s->threads) schedule(static)
  for (int j = 0; j < H; j++) {
    for (int i = 0; i < W; i++) {
      float cx = pos_half(i, s->h), cy = pos_half(j, s->h);
      float vx = general_velocity_x(s, cx, cy), vy = general_velocity_y(s, cx, cy);
      float px = fmaf(-vx, d_t, cx), py = fmaf(-vy, d_t, cy);
      s->smoke_buf[IDX(s, i, j)] = interpolate_smoke(s, px, py);
    }
  }
  float* t = s->smoke; s->smoke = s->smoke_buf; s->smoke_buf = t;
}

void oracle_decay_smoke(oracle_sim* s, float d_t) {
  if (!s->c.smoke_enable_decay) return;
  size_t n = (size_t)s->W * s->H;
  const float rate = s->c.smoke_decay_rate;
  for (size_t k = 0; k < n; k++) {
    float t = fmaf(-rate, d_t, s->smoke[k]);
    s->smoke[k] = (float)fmax((double)t, 0.0); /* fluid.cu:760-761, double max */
  }
}

/* ---- Fluid::apply_diffusion_at (fluid.cu:176-183), host loop (fluid.cu:185-190) ------------------
 * The reference sweeps u in place, n launches, in whatever order the blocks run (H1: a data race).  The
 * specification restated here fixes the order: cells with (i+j) even, then (i+j) odd — the same order the
 * CUDA path uses; against the reference's racy binary the two differ by O(a^2) per sweep. */
void oracle_diffusion(oracle_sim* s, int iterations, float d_t) {
  const int W = s->W, H = s->H;
  const float a = (s->c.viscosity * d_t) / (float)(s->h * s->h); /* fluid.cu:177, square(int) */
  const float denom = fmaf(4.0f, a, 1.0f);                         /* 1 + 4 * a */
  for (int it = 0; it < iterations; it++) {
    for (int colour = 0; colour < 2; colour++) {
#pragma omp parallel for num_threads(s->threads) schedule(static)
      for (int j = 1; j < H - 1; j++) {
        for (int i = 1 + ((j + 1 + colour) & 1); i < W - 1; i += 2) {
          size_t k = IDX(s, i, j);
          float sum = ((s->u[IDX(s, i - 1, j)] + s->u[IDX(s, i + 1, j)]) + s->u[IDX(s, i, j - 1)]) +
                      s->u[IDX(s, i, j + 1)];
          s->u[k] = fmaf(a, sum, s->u[k]) / denom;
        }
      }
    }
  }
}

/* ---- Fluid::update (fluid.cu:770-795) ------------------------------------------------------------ */
void oracle_step(oracle_sim* s, const sayal_source* src, float d_t) {
  oracle_forces(s, src, d_t);
  if (s->c.enable_pressure) oracle_zero_pressure(s);
  if (s->c.viscosity != 0) oracle_diffusion(s, s->c.proj_n, d_t); /* fluid.cu:775-777 */
  oracle_projection(s, s->c.proj_n, d_t);
  if (s->c.enable_pressure) oracle_pressure_range(s);
  oracle_extrapolation(s);
  oracle_advect_velocity(s, d_t);
  if (s->c.enable_smoke && s->c.wt_smoke != 0) {
    oracle_advect_smoke(s, d_t);
    oracle_decay_smoke(s, d_t);
  }
}

/* ---- Fluid::trace (fluid.cu:16-36) under GraphicsHandler::update_traces (graphics_handler.cu:358-421) ----
 * static_cast<int>(round(x)): roundf (half away from zero), then the GPU's saturating conversion. */
void oracle_path_lines(oracle_sim* s, const sayal_visual* vis, float d_t, int32_t* xs, int32_t* ys) {
  const int dist = vis->path_line_distance, len = vis->path_line_length;
  const int nx = s->W / dist, ny = s->H / dist;
  for (int b = 0; b < ny; b++) {
    for (int a = 0; a < nx; a++) {
      const int i = a * dist, j = b * dist;
      int32_t* lx = xs + ((size_t)(ny - 1 - b) * nx + a) * len;
      int32_t* ly = ys + ((size_t)(ny - 1 - b) * nx + a) * len;
      if (s->is_solid[IDX(s, i, j)]) {
        for (int k = 0; k < len; k++) lx[k] = ly[k] = -1;
        continue;
      }
      float px = pos_half(i, s->h), py = pos_half(j, s->h);
      lx[0] = f2i(roundf(px));
      ly[0] = s->H - 1 - f2i(roundf(py));
      for (int k = 1; k < len; k++) {
        float vx = general_velocity_x(s, px, py), vy = general_velocity_y(s, px, py);
        px = fmaf(vx, d_t, px); /* position + velocity * d_t, contracted by nvcc */
        py = fmaf(vy, d_t, py);
        lx[k] = f2i(roundf(px));
        ly[k] = s->H - 1 - f2i(roundf(py));
      }
    }
  }
}

/* ---- update_center_velocity_arrow_at (graphics_handler.cu:317-335) + make_arrow_data (:168-200) ------
 * The float transcendental calls of the source (atan2, cos, sin on float => atan2f, cosf, sinf) are evaluated in
 * double and rounded to float: correctly rounded up to a 2^-29 chance, hence the same bits on the CPU and on the
 * GPU.  (The reference's binary uses the fast-math approximations; its arrow tips may differ by one pixel.) */
static inline float cosf_cr(float x) { return (float)cos((double)x); }
static inline float sinf_cr(float x) { return (float)sin((double)x); }

void oracle_arrows(oracle_sim* s, const sayal_visual* vis, sayal_arrow* out) {
  const int dist = vis->arrows_distance, cs = vis->cell_pixel_size;
  const int nx = s->W / dist, ny = s->H / dist;
  const float head_angle = (float)(M_PI / 8); /* ARROW_HEAD_ANGLE, config.hpp:6 */
  const float head_len = (float)vis->arrows_head_length;
  for (int b = 0; b < ny; b++) {
    for (int a = 0; a < nx; a++) {
      const int i = a * dist, j = b * dist;
      sayal_arrow* ar = out + (size_t)(ny - 1 - b) * nx + a;
      memset(ar, 0, sizeof *ar);
      if (s->is_solid[IDX(s, i, j)]) continue;
      float x = (float)(((double)i + 0.5) * (double)cs);
      float y = (float)(((double)(s->H - j - 1) + 0.5) * (double)cs);
      float qy = (float)(s->H * cs) - y;
      float vx = general_velocity_x(s, x, qy), vy = general_velocity_y(s, x, qy);
      float angle = (float)atan2((double)vy, (double)vx);
      float length = sqrtf(fmaf(vx, vx, vy * vy));
      if (length < vis->arrows_disable_threshold) continue;
      ar->valid = 1;
      int sx = f2i(x), sy = f2i(y); /* make_arrow_data(int x, int y, ...) */
      ar->start_x = sx;
      ar->start_y = sy;
      length *= vis->arrows_length_multiplier;
      int x_off = f2i(length * cosf_cr(angle));
      int y_off = f2i(-length * sinf_cr(angle));
      ar->end_x = sx + x_off;
      ar->end_y = sy + y_off;
      int hx = f2i(-head_len * cosf_cr(angle + head_angle));
      int hy = f2i(head_len * sinf_cr(angle + head_angle));
      ar->left_head_end_x = ar->end_x + hx;
      ar->left_head_end_y = ar->end_y + hy;
      hx = f2i(-head_len * cosf_cr(head_angle - angle));
      hy = f2i(-head_len * sinf_cr(head_angle - angle));
      ar->right_head_end_x = ar->end_x + hx;
      ar->right_head_end_y = ar->end_y + hy;
    }
  }
}

/* ---- frame read-back: GraphicsHandler::update_fluid_pixels (graphics_handler.cu:214-302) with
 * hsv_to_rgb / map_rgba / clamp (helper.cu:3-51).  Pixel (x = i, y = H-1-j) sits at y*W + x, i.e. the pixel
 * array has the same layout as the fields.  static_cast<uint8_t>(float) is what nvcc emits for it:
 * cvt.rzi.u32.f32 (saturating, NaN -> 0) and the low 8 bits. */
static inline uint32_t f2u8(float x) {
  uint32_t u;
  if (!(x > 0.0f)) u = 0; /* negative, -0, NaN */
  else if (x >= 4294967296.0f) u = 0xffffffffu;
  else u = (uint32_t)x;
  return u & 255u;
}

static inline uint32_t map_rgba(uint32_t r, uint32_t g, uint32_t b, uint32_t a) { return r << 24 | g << 16 | b << 8 | a; }

static inline float clampf(float x, float lo, float hi) { return (x < lo) ? lo : (x > hi) ? hi : x; }

static void hsv_to_rgb(float h, float s, float v, uint32_t* r, uint32_t* g, uint32_t* b) {
  float c = v * s;
  float x = c * (1.0f - fabsf(fmodf(h / 60.0f, 2.0f) - 1.0f));
  float m = v - c;
  float r_, g_, b_;
  if (h < 60) { r_ = c; g_ = x; b_ = 0; }
  else if (h < 120) { r_ = x; g_ = c; b_ = 0; }
  else if (h < 180) { r_ = 0; g_ = c; b_ = x; }
  else if (h < 240) { r_ = 0; g_ = x; b_ = c; }
  else if (h < 300) { r_ = x; g_ = 0; b_ = c; }
  else { r_ = c; g_ = 0; b_ = x; }
  *r = f2u8((r_ + m) * 255.0f);
  *g = f2u8((g_ + m) * 255.0f);
  *b = f2u8((b_ + m) * 255.0f);
}

void oracle_render_pixels(oracle_sim* s, uint32_t* pixels) {
  const size_t n = (size_t)s->W * s->H;
  const int P = s->c.enable_pressure != 0, S = s->c.enable_smoke != 0;
  const float mn = s->min_p, mx = s->max_p;
  for (size_t k = 0; k < n; k++) {
    if (s->is_solid[k]) { pixels[k] = map_rgba(80, 80, 80, 255); continue; }
    if (P) {
      float pr = s->p[k], norm_p;
      if (S) { /* update_smoke_and_pressure (:221-238) */
        norm_p = 0;
        if (pr < 0 && mn != 0) norm_p = -pr / mn;
        else if (mx != 0) norm_p = pr / mx;
      } else { /* update_pressure_pixel (:240-256) */
        if (pr < 0) norm_p = -pr / mn;
        else norm_p = pr / mx;
      }
      norm_p = clampf(norm_p, -1.0f, 1.0f);
      float hue = (1.0f - norm_p) * 120.0f;
      uint32_t r, g, b;
      hsv_to_rgb(hue, 1.0f, S ? s->smoke[k] : 1.0f, &r, &g, &b);
      pixels[k] = map_rgba(r, g, b, 255);
    } else if (S) { /* update_smoke_pixels (:214-219) */
      uint32_t color = (255u - f2u8(s->smoke[k] * 255.0f)) & 255u;
      pixels[k] = map_rgba(255, color, color, 255);
    }
    /* neither: the reference leaves the pixel untouched */
  }
}

int oracle_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
