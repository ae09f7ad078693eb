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

#include <vector>

#include "config_parser.hpp"
#include "fluid.cuh"
/* GraphicsHandler keeps its device buffers private; the shim zeroes d_arrow_data once (see ref_gfx_create).  The
 * access specifier does not change the class layout, and graphics_handler.cu itself is compiled untouched. */
#define private public
#include "graphics_handler.cuh"
#undef private
#include "sayal.h"

struct ref_sim {
  Fluid* fluid;
  int W, H;
  cudaEvent_t e0, e1;
};

static Config to_reference_config(const sayal_config* c, const sayal_visual* vis = nullptr) {
  Config cfg;
  std::memset(&cfg, 0, sizeof(cfg));
  if (vis) {
    cfg.visual.arrows.enable = vis->arrows_enable != 0;
    cfg.visual.arrows.distance = vis->arrows_distance;
    cfg.visual.arrows.length_multiplier = vis->arrows_length_multiplier;
    cfg.visual.arrows.disable_threshold = vis->arrows_disable_threshold;
    cfg.visual.arrows.head_length = vis->arrows_head_length;
    cfg.visual.arrows.color = {vis->arrows_color[0], vis->arrows_color[1], vis->arrows_color[2], vis->arrows_color[3]};
    cfg.visual.path_line.enable = vis->path_line_enable != 0;
    cfg.visual.path_line.length = vis->path_line_length;
    cfg.visual.path_line.distance = vis->path_line_distance;
    cfg.visual.path_line.color = {vis->path_line_color[0], vis->path_line_color[1], vis->path_line_color[2],
                                  vis->path_line_color[3]};
  } else {  // GraphicsHandler divides by these
    cfg.visual.arrows.distance = 20;
    cfg.visual.path_line.distance = 20;
    cfg.visual.path_line.length = 20;
  }
  cfg.thread.openMP.thread_count = 1;
  cfg.thread.cuda.block_size_x = c->block_size_x > 0 ? c->block_size_x : 64;
  cfg.thread.cuda.block_size_y = c->block_size_y > 0 ? c->block_size_y : 1;
  cfg.sim.height = c->height;
  cfg.sim.width = c->width;
  cfg.sim.cell_pixel_size = vis ? vis->cell_pixel_size : 1;
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

/* ---- the reference's own renderer, headless -----------------------------------------------------------------
 * graphics_handler.cu is compiled unmodified next to fluid.cu.  The SDL entry points it calls are defined below
 * as recording stubs: window / renderer / texture creation hand back dummy non-null handles, SDL_UpdateTexture
 * keeps the RGBA frame, SDL_RenderDrawLines keeps every path line, SDL_RenderDrawLine keeps every arrow segment.
 * GraphicsHandler::update(fluid, d_t) (graphics_handler.cu:463-478) therefore runs exactly as in the reference's
 * main loop (main.cu:98) and what it would have drawn is what the tests compare against. */
struct gfx_record {
  std::vector<uint32_t> pixels;
  std::vector<int32_t> polyline_points;  /* x, y pairs; every polyline has `count` points */
  std::vector<int32_t> polyline_counts;
  std::vector<int32_t> segments;         /* x1, y1, x2, y2 */
  int tex_w = 0, tex_h = 0;
};
static gfx_record g_rec;
static int g_dummy_handles[4];

struct ref_gfx {
  GraphicsHandler* gh;
};

int ref_gfx_create(const sayal_config* c, const sayal_visual* vis, ref_gfx** out) {
  ref_gfx* g = new (std::nothrow) ref_gfx();
  if (!g) return SAYAL_ENOMEM;
  g_rec.tex_w = c->width;
  g_rec.tex_h = c->height;
  g->gh = new GraphicsHandler(to_reference_config(c, vis));
  /* H15: update_center_velocity_arrow_at returns without writing for solid cells (graphics_handler.cu:320-322), so
   * those entries of the cudaMalloc'd d_arrow_data are never initialised and the host loop draws whatever `valid`
   * byte it finds.  Zero the buffer once so that the reference's output is deterministic (same idea as H2). */
  cudaMemset(g->gh->d_arrow_data, 0, sizeof(ArrowData) * (size_t)g->gh->arrow_data_width * g->gh->arrow_data_height);
  if (cudaDeviceSynchronize() != cudaSuccess || cudaGetLastError() != cudaSuccess) return SAYAL_ECUDA;
  *out = g;
  return SAYAL_OK;
}

void ref_gfx_destroy(ref_gfx* g) {
  if (!g) return;
  delete g->gh;
  delete g;
}

/* graphics.update(fluid, d_t) (main.cu:98) */
int ref_gfx_update(ref_gfx* g, ref_sim* s, float d_t) {
  g_rec.pixels.clear();
  g_rec.polyline_points.clear();
  g_rec.polyline_counts.clear();
  g_rec.segments.clear();
  g->gh->update(*s->fluid, d_t);
  return cudaGetLastError() == cudaSuccess ? SAYAL_OK : SAYAL_ECUDA;
}

int64_t ref_gfx_pixels(uint32_t* dst, int64_t capacity) {
  int64_t n = (int64_t)g_rec.pixels.size();
  if (dst && capacity >= n) std::memcpy(dst, g_rec.pixels.data(), sizeof(uint32_t) * n);
  return n;
}
int64_t ref_gfx_polylines(int32_t* points_dst, int64_t capacity, int32_t* n_lines) {
  int64_t n = (int64_t)g_rec.polyline_points.size();
  if (points_dst && capacity >= n) std::memcpy(points_dst, g_rec.polyline_points.data(), sizeof(int32_t) * n);
  if (n_lines) *n_lines = (int32_t)g_rec.polyline_counts.size();
  return n;
}
int64_t ref_gfx_segments(int32_t* dst, int64_t capacity) {
  int64_t n = (int64_t)g_rec.segments.size();
  if (dst && capacity >= n) std::memcpy(dst, g_rec.segments.data(), sizeof(int32_t) * n);
  return n;
}

/* Fluid::get_general_velocity on the device (fluid.cu:541-545), for sampler parity */
__global__ void ref_sample_kernel(Fluid* f, int n, const float* xs, const float* ys, float* ou, float* ov) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  Vector2d<float> v = f->get_general_velocity(xs[t], ys[t]);
  ou[t] = v.get_x();
  ov[t] = v.get_y();
}
int ref_sample_velocity(ref_sim* s, int n, const float* xs, const float* ys, float* ou, float* ov) {
  float* d = nullptr;
  if (cudaMalloc(&d, sizeof(float) * 4 * (size_t)n) != cudaSuccess) return SAYAL_ENOMEM;
  cudaMemcpy(d, xs, sizeof(float) * n, cudaMemcpyHostToDevice);
  cudaMemcpy(d + n, ys, sizeof(float) * n, cudaMemcpyHostToDevice);
  ref_sample_kernel<<<(n + 127) / 128, 128>>>(s->fluid->d_this, n, d, d + n, d + 2 * (size_t)n, d + 3 * (size_t)n);
  cudaMemcpy(ou, d + 2 * (size_t)n, sizeof(float) * n, cudaMemcpyDeviceToHost);
  cudaMemcpy(ov, d + 3 * (size_t)n, sizeof(float) * n, cudaMemcpyDeviceToHost);
  cudaFree(d);
  return cudaGetLastError() == cudaSuccess ? SAYAL_OK : SAYAL_ECUDA;
}

/* ---- SDL stubs (signatures from lib/sdl/include) ------------------------------------------------------------ */
int SDL_Init(Uint32) { return 0; }
void SDL_Quit(void) {}
const char* SDL_GetError(void) { return "headless stub"; }
SDL_Window* SDL_CreateWindow(const char*, int, int, int, int, Uint32) { return reinterpret_cast<SDL_Window*>(&g_dummy_handles[0]); }
SDL_Renderer* SDL_CreateRenderer(SDL_Window*, int, Uint32) { return reinterpret_cast<SDL_Renderer*>(&g_dummy_handles[1]); }
SDL_Texture* SDL_CreateTexture(SDL_Renderer*, Uint32, int, int w, int h) {
  g_rec.tex_w = w;
  g_rec.tex_h = h;
  return reinterpret_cast<SDL_Texture*>(&g_dummy_handles[2]);
}
SDL_PixelFormat* SDL_AllocFormat(Uint32) { return reinterpret_cast<SDL_PixelFormat*>(&g_dummy_handles[3]); }
void SDL_FreeFormat(SDL_PixelFormat*) {}
void SDL_DestroyWindow(SDL_Window*) {}
void SDL_DestroyRenderer(SDL_Renderer*) {}
void SDL_DestroyTexture(SDL_Texture*) {}
int SDL_RenderClear(SDL_Renderer*) { return 0; }
int SDL_RenderCopy(SDL_Renderer*, SDL_Texture*, const SDL_Rect*, const SDL_Rect*) { return 0; }
void SDL_RenderPresent(SDL_Renderer*) {}
int SDL_SetRenderDrawColor(SDL_Renderer*, Uint8, Uint8, Uint8, Uint8) { return 0; }
int SDL_UpdateTexture(SDL_Texture*, const SDL_Rect*, const void* pixels, int pitch) {
  const char* src = static_cast<const char*>(pixels);
  g_rec.pixels.resize((size_t)g_rec.tex_w * g_rec.tex_h);
  for (int y = 0; y < g_rec.tex_h; y++)
    std::memcpy(g_rec.pixels.data() + (size_t)y * g_rec.tex_w, src + (size_t)y * pitch, sizeof(uint32_t) * g_rec.tex_w);
  return 0;
}
int SDL_RenderDrawLine(SDL_Renderer*, int x1, int y1, int x2, int y2) {
  g_rec.segments.insert(g_rec.segments.end(), {x1, y1, x2, y2});
  return 0;
}
int SDL_RenderDrawLines(SDL_Renderer*, const SDL_Point* points, int count) {
  for (int k = 0; k < count; k++) g_rec.polyline_points.insert(g_rec.polyline_points.end(), {points[k].x, points[k].y});
  g_rec.polyline_counts.push_back(count);
  return 0;
}

}  // extern "C"
