/*
 * ref_shim.cu — headless C wrapper around the UNMODIFIED reference `class Fluid`
 * (/root/reference/inc/fluid.cuh, /root/reference/src/fluid.cu).
 *
 * TEST INFRASTRUCTURE.  Built by oracle/Makefile together with the reference's own sources, compiled
 * from where they lie under /root/reference with the reference's Release flags (CMakeLists.txt:21)
 * plus -gencode arch=compute_100,code=sm_100, into oracle/_ref/libsayal_ref.so.  No reference source is
 * copied into this repository; this file only *includes* the reference headers and drives the
 * public surface of Fluid the way src/main.cu:49,97 does (Fluid fluid(config); fluid.update(source, d_t)).
 *
 * Used to (1) pin the CPU oracle and the B200-native path against the reference's real arithmetic on
 * the GPU box, (2) time the reference on the same B200 (bench.py --impl reference).
 */
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <new>

#include <cuda_runtime.h>

#include "config_parser.hpp"
#include "fluid.cuh"
#include "sayal.h"

struct ref_sim {
  Fluid* fluid;
  int W, H;
  cudaEvent_t e0, e1;
};

static Config to_reference_config(const sayal_config* c) {
  Config cfg;
  std::memset(&cfg, 0, sizeof(cfg));
  cfg.thread.openMP.thread_count = 1;
  cfg.thread.cuda.block_size_x = c->block_size_x > 0 ? c->block_size_x : 64;
  cfg.thread.cuda.block_size_y = c->block_size_y > 0 ? c->block_size_y : 1;
  cfg.sim.height = c->height;
  cfg.sim.width = c->width;
  cfg.sim.cell_pixel_size = 1;
  cfg.sim.cell_size = c->cell_size;
  cfg.sim.enable_drain = c->enable_drain != 0;
  cfg.sim.enable_pressure = c->enable_pressure != 0;
  cfg.sim.enable_smoke = c->enable_smoke != 0;
  cfg.sim.enable_interactive = c->enable_interactive != 0;
  cfg.sim.projection.n = c->proj_n;
  cfg.sim.projection.o = c->proj_o;
  cfg.sim.wind_tunnel.pipe_height = c->wt_pipe_height;
  cfg.sim.wind_tunnel.pipe_length = c->wt_pipe_length;
  cfg.sim.wind_tunnel.smoke_length = c->wt_smoke_length;
  cfg.sim.wind_tunnel.smoke_height = c->wt_smoke_height;
  cfg.sim.wind_tunnel.smoke_count = c->wt_smoke_count;
  cfg.sim.wind_tunnel.speed = c->wt_speed;
  cfg.sim.wind_tunnel.smoke = c->wt_smoke;
  cfg.sim.physics.g = c->g;
  cfg.sim.time.d_t = c->d_t;
  cfg.sim.time.enable_real_time = c->enable_real_time != 0;
  cfg.sim.time.real_time_multiplier = c->real_time_multiplier;
  cfg.sim.smoke.enable_decay = c->smoke_enable_decay != 0;
  cfg.sim.smoke.decay_rate = c->smoke_decay_rate;
  cfg.sim.obstacle.enable = c->obstacle_enable != 0;
  cfg.sim.obstacle.center_x = c->obstacle_center_x;
  cfg.sim.obstacle.center_y = c->obstacle_center_y;
  cfg.sim.obstacle.radius = c->obstacle_radius;
  cfg.fluid.density = c->density;
  cfg.fluid.drag_coeff = c->drag_coeff;
  cfg.fluid.viscosity = c->viscosity;
  return cfg;
}

static float* field_ptr(ref_sim* s, int field) {
  switch (field) {
    case SAYAL_U: return s->fluid->d_vel_x;
    case SAYAL_V: return s->fluid->d_vel_y;
    case SAYAL_P: return s->fluid->d_pressure;
    case SAYAL_SMOKE: return s->fluid->d_smoke;
    case SAYAL_IS_SOLID: return reinterpret_cast<float*>(s->fluid->d_is_solid);
    case SAYAL_TOTAL_S: return reinterpret_cast<float*>(s->fluid->d_total_s);
  }
  return nullptr;
}

extern "C" {

/* H2: Fluid::init_device_memory malloc()s total_s and only ever ++'s it (fluid.cu:102-103, 127-142), so the
 * reference is correct only when malloc returns fresh zero pages — true in its own short-lived process, not in
 * a long-lived test process whose heap holds dirty free chunks (observed on the GPU box: garbage total_s and a
 * 6e-2 error in u).  oracle/Makefile links this library with -Wl,--wrap=malloc, which redirects the malloc
 * calls of the objects linked here (the reference's fluid.o among them, unmodified) to this zeroing version. */
void* __wrap_malloc(size_t n) { return std::calloc(1, n ? n : 1); }

int ref_create(const sayal_config* c, int device, ref_sim** out) {
  if (cudaSetDevice(device) != cudaSuccess) return SAYAL_ECUDA;
  ref_sim* s = new (std::nothrow) ref_sim();
  if (!s) return SAYAL_ENOMEM;
  s->W = c->width;
  s->H = c->height;
  s->fluid = new Fluid(to_reference_config(c));
  /* the reference leaves d_pressure uninitialised (fluid.cu:88); give both sides the same start */
  cudaMemset(s->fluid->d_pressure, 0, sizeof(float) * (size_t)s->W * s->H);
  cudaEventCreate(&s->e0);
  cudaEventCreate(&s->e1);
  if (cudaDeviceSynchronize() != cudaSuccess || cudaGetLastError() != cudaSuccess) return SAYAL_ECUDA;
  *out = s;
  return SAYAL_OK;
}

void ref_destroy(ref_sim* s) {
  if (!s) return;
  delete s->fluid;
  cudaEventDestroy(s->e0);
  cudaEventDestroy(s->e1);
  delete s;
}

/* main.cu:97 — fluid.update(source, d_t); includes the reference's own cudaDeviceSynchronize */
int ref_step(ref_sim* s, const sayal_source* src, float d_t) {
  Source source;
  source.active = src && src->active;
  source.smoke = src ? src->smoke : 0.f;
  source.velocity = src ? src->velocity : 0.f;
  source.position = Vector2d<int>(src ? src->x : 0, src ? src->y : 0);
  s->fluid->update(source, d_t);
  return cudaGetLastError() == cudaSuccess ? SAYAL_OK : SAYAL_ECUDA;
}

/* `steps` updates timed with CUDA events on the legacy default stream the reference launches on */
int ref_run_timed(ref_sim* s, int steps, float d_t, float* elapsed_ms) {
  Source source;
  source.active = false;
  source.smoke = 0.f;
  source.velocity = 0.f;
  source.position = Vector2d<int>(0, 0);
  cudaEventRecord(s->e0, 0);
  for (int k = 0; k < steps; k++) s->fluid->update(source, d_t);
  cudaEventRecord(s->e1, 0);
  cudaEventSynchronize(s->e1);
  if (elapsed_ms) cudaEventElapsedTime(elapsed_ms, s->e0, s->e1);
  return cudaGetLastError() == cudaSuccess ? SAYAL_OK : SAYAL_ECUDA;
}

int ref_get_field(ref_sim* s, int field, void* host_dst) {
  float* p = field_ptr(s, field);
  if (!p) return SAYAL_EINVAL;
  return cudaMemcpy(host_dst, p, sizeof(float) * (size_t)s->W * s->H, cudaMemcpyDeviceToHost) == cudaSuccess
             ? SAYAL_OK
             : SAYAL_ECUDA;
}

int ref_set_field(ref_sim* s, int field, const void* host_src) {
  float* p = field_ptr(s, field);
  if (!p) return SAYAL_EINVAL;
  return cudaMemcpy(p, host_src, sizeof(float) * (size_t)s->W * s->H, cudaMemcpyHostToDevice) == cudaSuccess
             ? SAYAL_OK
             : SAYAL_ECUDA;
}

int ref_pressure_range(ref_sim* s, float* mn, float* mx) {
  *mn = s->fluid->min_pressure;
  *mx = s->fluid->max_pressure;
  return SAYAL_OK;
}

}  // extern "C"
