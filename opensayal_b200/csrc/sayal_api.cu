// sayal_api.cu — the C ABI of include/sayal.h over the CUDA step path.  Host-side mirror of
// Fluid::Fluid / ~Fluid / update (/root/reference/src/fluid.cu:41-97, 770-795).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>

#include "sayal_internal.h"

// measurement aids of sayal_stream_delay / sayal_stream_hold (defined here so that create_impl can preload them)
__global__ void stream_delay_kernel(long long ns) {
  long long t0, t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  do {
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  } while (t - t0 < ns);
}

__global__ void stream_gate_kernel(const volatile unsigned* gate, unsigned ticket) {
  long long t0, t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  while ((int)(*gate - ticket) < 0) {
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    if (t - t0 > 20000000000ll) break;
    __nanosleep(200);
  }
}

namespace sayal {

static thread_local std::string g_last_error;

int set_error(int code, const char* msg) {
  g_last_error = msg ? msg : "";
  return code;
}

#define CUDA_TRY(expr)                                                                   \
  do {                                                                                   \
    cudaError_t e__ = (expr);                                                            \
    if (e__ != cudaSuccess) {                                                            \
      char m__[320];                                                                     \
      snprintf(m__, sizeof m__, "%s failed: %s", #expr, cudaGetErrorString(e__));        \
      return set_error(SAYAL_ECUDA, m__);                                                \
    }                                                                                    \
  } while (0)

#define TRY(expr)            \
  do {                       \
    int r__ = (expr);        \
    if (r__ != SAYAL_OK) return r__; \
  } while (0)

static size_t field_elems(const Sim* s) { return (size_t)s->g.pitch * s->g.local_rows; }

static void free_sim(Sim* s) {
  if (!s) return;
  cudaSetDevice(s->device);
  for (int k = 0; k < 4; k++)
    if (s->graph[k]) cudaGraphExecDestroy(s->graph[k]);
  if (s->p) cudaFree(s->p);
  if (s->smoke_block) cudaFree(s->smoke_block);  // smoke, smoke_buf
  if (s->vel_block) cudaFree(s->vel_block);  // u, v, u_buf, v_buf
  for (int k = 0; k < s->n_orders; k++)
    if (s->orders[k].order) cudaFree(s->orders[k].order);
  if (s->flags) cudaFree(s->flags);
  if (s->geo) cudaFree(s->geo);
  if (s->d_is_solid) cudaFree(s->d_is_solid);
  if (s->d_total_s) cudaFree(s->d_total_s);
  if (s->d_range) cudaFree(s->d_range);
  if (s->d_overflow) cudaFree(s->d_overflow);
  if (s->d_resident) cudaFree(s->d_resident);
  if (s->d_box) cudaFree(s->d_box);
  if (s->h_resident_error) cudaFreeHost(s->h_resident_error);
  if (s->d_timeline) cudaFree(s->d_timeline);
  for (int k = 0; k < 2; k++) {
    if (s->ipc_opened[k]) cudaIpcCloseMemHandle(s->ipc_opened[k]);
    if (s->ipc_opened_vel[k]) cudaIpcCloseMemHandle(s->ipc_opened_vel[k]);
  }
  if (s->link_block) cudaFree(s->link_block);
  if (s->link_counters) cudaFree(s->link_counters);
  if (s->h_link_error) cudaFreeHost(s->h_link_error);
  if (s->h_gate) cudaFreeHost(s->h_gate);
  if (s->ev_range) cudaEventDestroy(s->ev_range);

  for (int k = 0; k < Sim::kFrames; k++) {
    if (s->d_frame[k]) cudaFree(s->d_frame[k]);
    if (s->h_frame[k]) cudaFreeHost(s->h_frame[k]);
    if (s->ev_rendered[k]) cudaEventDestroy(s->ev_rendered[k]);
    if (s->ev_copied[k]) cudaEventDestroy(s->ev_copied[k]);
  }
  if (s->copy_stream) cudaStreamDestroy(s->copy_stream);
  for (int k = 0; k < Sim::kDbgEvents; k++)
    if (s->dbg_ev[k]) cudaEventDestroy(s->dbg_ev[k]);
  if (s->ev_fork) cudaEventDestroy(s->ev_fork);
  if (s->ev_join) cudaEventDestroy(s->ev_join);
  if (s->aux_stream) cudaStreamDestroy(s->aux_stream);
  if (s->stream) cudaStreamDestroy(s->stream);
  delete s;
}

static void invalidate_graphs(Sim* s) {
  for (int k = 0; k < 4; k++)
    if (s->graph[k]) {
      cudaGraphExecDestroy(s->graph[k]);
      s->graph[k] = nullptr;
    }
}

static int create_impl(const sayal_config* c, int device, const sayal_slab* slab, Sim** out) {
  if (!c || !out) return set_error(SAYAL_EINVAL, "sayal_create: null argument");
  if (c->width < 4 || c->height < 4) return set_error(SAYAL_EINVAL, "sayal_create: width and height must be >= 4");
  if ((int64_t)c->width * c->height >= (int64_t)1 << 31)
    return set_error(SAYAL_EINVAL, "sayal_create: width*height must fit int32 (reference limit, fluid.cu:163)");
  if ((int)c->cell_size < 1) return set_error(SAYAL_EINVAL, "sayal_create: cell_size must be an integer >= 1 (fluid.cuh:46)");
  if (c->proj_n < 0) return set_error(SAYAL_EINVAL, "sayal_create: projection.n must be >= 0");
  int ndev = 0;
  CUDA_TRY(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev) return set_error(SAYAL_EINVAL, "sayal_create: no such CUDA device");
  CUDA_TRY(cudaSetDevice(device));

  Sim* s = new (std::nothrow) Sim();
  if (!s) return set_error(SAYAL_ENOMEM, "sayal_create: out of host memory");
  std::memset(s, 0, sizeof(Sim));
  s->cfg = *c;
  s->device = device;
  Grid& g = s->g;
  g.W = c->width;
  g.H = c->height;
  g.pitch = (c->width + 3) & ~3;
  g.h = (int)c->cell_size;
  if (slab && slab->rows > 0) {
    if (slab->global_height != c->height || slab->row0 < 0 || slab->row0 + slab->rows > c->height || slab->halo < 0) {
      delete s;
      return set_error(SAYAL_EINVAL, "sayal_create_slab: slab does not fit the domain");
    }
    // the `halo` owned rows next to an interior edge are what a neighbour receives as its ghost rows
    if (slab->rows < slab->halo && slab->rows != c->height) {
      delete s;
      return set_error(SAYAL_EINVAL, "sayal_create_slab: a slab must own at least `halo` rows");
    }
    int lo = slab->row0 - slab->halo, hi = slab->row0 + slab->rows + slab->halo;
    if (lo < 0) lo = 0;
    if (hi > c->height) hi = c->height;
    g.row_base = lo;
    g.local_rows = hi - lo;
    g.own_lo = slab->row0 - lo;
    g.own_hi = g.own_lo + slab->rows;
    s->slab_halo = slab->halo;
  } else {
    g.row_base = 0;
    g.local_rows = c->height;
    g.own_lo = 0;
    g.own_hi = c->height;
  }
  g.valid_lo = 0;
  g.valid_hi = g.local_rows;
  Phys& p = s->ph;
  p.o = c->proj_o;
  p.density = c->density;
  p.g = c->g;
  p.drag_coeff = c->drag_coeff;
  p.wt_speed = c->wt_speed;
  p.wt_smoke = c->wt_smoke;
  p.wt_height = c->wt_pipe_height;
  p.wt_smoke_length = c->wt_smoke_length;
  p.wt_smoke_count = c->wt_smoke_count;
  p.wt_smoke_height = c->wt_smoke_height;
  p.enable_pressure = c->enable_pressure != 0;
  p.enable_smoke = c->enable_smoke != 0;
  p.enable_decay = c->smoke_enable_decay != 0;
  p.decay_rate = c->smoke_decay_rate;
  p.enable_drain = c->enable_drain != 0;
  p.obstacle_enable = c->obstacle_enable != 0;
  p.obstacle_cx = c->obstacle_center_x;
  p.obstacle_cy = c->obstacle_center_y;
  p.wt_pipe_length = c->wt_pipe_length;
  p.wt_pipe_height = c->wt_pipe_height;
  p.obstacle_radius = c->obstacle_radius;

  s->projection_kernel = 1;
  s->temporal_block = 0;  // 0 = auto
  s->use_graph = 1;
  s->use_pdl = 1;
  s->fuse_forces = 1;
  s->fuse_extrapolation = 1;
  s->shrink_window = 1;
  s->proj_depth = -1;
  s->order_tiles = 1;
  s->split_tiles = 1;
  s->tune_depth = -1;
  s->advect_kernel = 2;
  s->overlap_exchange = 1;
  s->slab_push = 1;
  s->resident = 1;
  s->advect_margin = 16;
  s->autotune = 1;
  s->plan_variant = -1, s->n_plans = 0;
  s->force_variant = -1;
  // experiment / profiling overrides (never change results): SAYAL_AUTOTUNE=0, SAYAL_TEMPORAL_BLOCK=T,
  // SAYAL_TILE_ROWS=8|10|12, SAYAL_PROJECTION_KERNEL=0|1, SAYAL_USE_GRAPH=0|1
  if (const char* e = getenv("SAYAL_AUTOTUNE")) s->autotune = atoi(e) != 0;
  if (const char* e = getenv("SAYAL_TEMPORAL_BLOCK")) { int t = atoi(e); if (t >= 0 && t <= tiled_max_temporal_block()) s->temporal_block = t; }
  if (const char* e = getenv("SAYAL_TILE_ROWS")) { int r = atoi(e); if (r == 8 || r == 10 || r == 12) s->force_variant = (r - 8) / 2; }
  if (const char* e = getenv("SAYAL_PROJECTION_KERNEL")) s->projection_kernel = atoi(e) != 0;
  if (const char* e = getenv("SAYAL_USE_GRAPH")) s->use_graph = atoi(e) != 0;
  if (const char* e = getenv("SAYAL_RESIDENT")) { int r = atoi(e); if (r >= 0 && r <= 2) s->resident = r; }
  if (const char* e = getenv("SAYAL_ADVECT_KERNEL")) { int k = atoi(e); if (k >= 0 && k <= 2) s->advect_kernel = k; }
  if (const char* e = getenv("SAYAL_ADVECT_MARGIN")) { int m = atoi(e); if (m >= 2) s->advect_margin = m; }
  if (const char* e = getenv("SAYAL_OVERLAP_EXCHANGE")) s->overlap_exchange = atoi(e) != 0;
  if (const char* e = getenv("SAYAL_USE_PDL")) s->use_pdl = atoi(e) != 0;
  if (const char* e = getenv("SAYAL_FUSE_FORCES")) s->fuse_forces = atoi(e) != 0;
  if (const char* e = getenv("SAYAL_ORDER_TILES")) s->order_tiles = atoi(e) != 0;
  if (const char* e = getenv("SAYAL_SPLIT_TILES")) s->split_tiles = atoi(e) != 0;
  if (const char* e = getenv("SAYAL_SHRINK_WINDOW")) s->shrink_window = atoi(e) != 0;
  if (const char* e = getenv("SAYAL_FUSE_EXTRAPOLATION")) s->fuse_extrapolation = atoi(e) != 0;
  if (const char* e = getenv("SAYAL_SLAB_PUSH")) s->slab_push = atoi(e) != 0;

  auto fail = [&](int code) {
    free_sim(s);
    return code;
  };
  cudaError_t e = cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking);
  if (e != cudaSuccess) return fail(set_error(SAYAL_ECUDA, cudaGetErrorString(e)));
  if (s->slab_halo > 0) {  // slabs: a second stream for exchanges that overlap interior compute
    int prio_lo = 0, prio_hi = 0;
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);  // exchanges overtake the compute they run under
    e = cudaStreamCreateWithPriority(&s->aux_stream, cudaStreamNonBlocking, prio_hi);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s->ev_fork, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s->ev_join, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s->ev_range, cudaEventDisableTiming);
    if (e != cudaSuccess) return fail(set_error(SAYAL_ECUDA, cudaGetErrorString(e)));
  }
  size_t bytes = field_elems(s) * sizeof(float);
  e = cudaMalloc(&s->p, bytes);
  if (e != cudaSuccess) return fail(set_error(SAYAL_ENOMEM, cudaGetErrorString(e)));
  cudaMemsetAsync(s->p, 0, bytes, s->stream);
  // smoke and its back buffer: one allocation with a guard of one row + one cell before, between and behind the
  // arrays.  advect_smoke_geo_kernel loads the four taps of a sample unconditionally and drops the closed ones by a
  // select; with a base cell on the domain's border a closed tap's address lies up to one row + one cell outside.
  {
    const size_t guard = (((size_t)s->g.pitch + 1) * sizeof(float) + 255) & ~(size_t)255;
    const size_t stride = ((bytes + 255) & ~(size_t)255) + guard;
    e = cudaMalloc(&s->smoke_block, guard + 2 * stride);
    if (e != cudaSuccess) return fail(set_error(SAYAL_ENOMEM, cudaGetErrorString(e)));
    cudaMemsetAsync(s->smoke_block, 0, guard + 2 * stride, s->stream);
    s->smoke = reinterpret_cast<float*>(reinterpret_cast<char*>(s->smoke_block) + guard);
    s->smoke_buf = reinterpret_cast<float*>(reinterpret_cast<char*>(s->smoke_block) + guard + stride);
  }
  // u, v and their back buffers: one allocation (one IPC handle for a neighbouring slab), arrays 256-byte aligned
  s->vel_stride = (bytes + 255) & ~(size_t)255;
  e = cudaMalloc(&s->vel_block, 4 * s->vel_stride);
  if (e != cudaSuccess) return fail(set_error(SAYAL_ENOMEM, cudaGetErrorString(e)));
  cudaMemsetAsync(s->vel_block, 0, 4 * s->vel_stride, s->stream);
  {
    float** vel[] = {&s->u, &s->v, &s->u_buf, &s->v_buf};
    for (int k = 0; k < 4; k++) *vel[k] = reinterpret_cast<float*>(reinterpret_cast<char*>(s->vel_block) + k * s->vel_stride);
  }
  // + slack marked solid: the tap (W, j) of the last row reads one byte past the array when pitch == W
  e = cudaMalloc(&s->flags, field_elems(s) + 64);
  if (e != cudaSuccess) return fail(set_error(SAYAL_ENOMEM, cudaGetErrorString(e)));
  cudaMemsetAsync(s->flags, FL_SOLID, field_elems(s) + 64, s->stream);
  e = cudaMalloc(&s->d_range, 4 * sizeof(int32_t));
  if (e != cudaSuccess) return fail(set_error(SAYAL_ENOMEM, cudaGetErrorString(e)));
  cudaMemsetAsync(s->d_range, 0, 4 * sizeof(int32_t), s->stream);  // min = max = 0 until the first step (fluid.cuh:61-62 are uninitialised there)
  e = cudaMalloc(&s->d_overflow, sizeof(int32_t));
  if (e != cudaSuccess) return fail(set_error(SAYAL_ENOMEM, cudaGetErrorString(e)));
  cudaMemsetAsync(s->d_overflow, 0, sizeof(int32_t), s->stream);
  // resident projection (projection_pack.cu): epoch + ticket; error word on the host
  e = cudaMalloc(&s->d_resident, 8 * sizeof(unsigned));
  if (e != cudaSuccess) return fail(set_error(SAYAL_ENOMEM, cudaGetErrorString(e)));
  cudaMemsetAsync(s->d_resident, 0, 8 * sizeof(unsigned), s->stream);
  e = cudaHostAlloc(reinterpret_cast<void**>(&s->h_resident_error), sizeof(int), cudaHostAllocMapped);
  if (e != cudaSuccess) return fail(set_error(SAYAL_ENOMEM, cudaGetErrorString(e)));
  *s->h_resident_error = 0;
  e = cudaHostGetDevicePointer(reinterpret_cast<void**>(&s->d_resident_error), s->h_resident_error, 0);
  if (e != cudaSuccess) return fail(set_error(SAYAL_ECUDA, cudaGetErrorString(e)));
  int r = launch_build_flags(s);
  if (r != SAYAL_OK) return fail(r);
  e = cudaMalloc(&s->geo, field_elems(s) * sizeof(uint16_t));
  if (e != cudaSuccess) return fail(set_error(SAYAL_ENOMEM, cudaGetErrorString(e)));
  r = launch_build_geo(s);
  if (r != SAYAL_OK) return fail(r);
  r = tiled_preload();
  if (r == SAYAL_OK) r = preload_basic();
  if (r == SAYAL_OK) r = preload_advect();
  if (r == SAYAL_OK) r = preload_slab();
  if (r == SAYAL_OK) r = preload_visual();
  if (r == SAYAL_OK) {  // this file's own kernels
    cudaFuncAttributes fa;
    if (cudaFuncGetAttributes(&fa, stream_delay_kernel) != cudaSuccess || cudaFuncGetAttributes(&fa, stream_gate_kernel) != cudaSuccess)
      r = set_error(SAYAL_ECUDA, "preload: stream kernels");
  }
  if (r != SAYAL_OK) return fail(r);
  e = cudaStreamSynchronize(s->stream);
  if (e != cudaSuccess) return fail(set_error(SAYAL_ECUDA, cudaGetErrorString(e)));
  *out = s;
  return SAYAL_OK;
}

// profiling only: stage boundary `k` of the running step, on `stream` (nothing while a graph is being captured)
static void dbg_mark(Sim* s, int k, cudaStream_t stream) {
  if (!s->debug_events || k < 0 || k >= Sim::kDbgEvents) return;
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  cudaStreamIsCapturing(s->stream, &cap);
  if (cap != cudaStreamCaptureStatusNone) return;
  if (!s->dbg_ev[k] && cudaEventCreate(&s->dbg_ev[k]) != cudaSuccess) return;
  cudaEventRecord(s->dbg_ev[k], stream);
  s->dbg_marked |= 1u << k;
}

static void swap_ptr(float*& a, float*& b) {
  float* t = a;
  a = b;
  b = t;
}

static int projection(Sim* s, int iterations, float d_t) {
  if (iterations <= 0) return SAYAL_OK;
  if (s->projection_kernel == 1) return launch_projection_tiled(s, iterations, d_t);
  return launch_projection_plain(s, iterations, d_t);
}

static int advect_velocity(Sim* s, float d_t) {
  TRY(s->advect_kernel == 2 ? launch_advect_geo(s, d_t, true, false)
      : s->advect_kernel == 1 ? launch_advect_tile(s, d_t, true, false) : launch_advect(s, d_t, true, false));
  swap_ptr(s->u, s->u_buf);  // update_velocity_advection_at (fluid.cu:614-617) as a pointer swap
  swap_ptr(s->v, s->v_buf);
  s->parity ^= 1;
  return SAYAL_OK;
}

static int advect_smoke(Sim* s, float d_t) {
  TRY(s->advect_kernel == 2 ? launch_advect_geo(s, d_t, false, true)
      : s->advect_kernel == 1 ? launch_advect_tile(s, d_t, false, true) : launch_advect(s, d_t, false, true));
  swap_ptr(s->smoke, s->smoke_buf);  // update_smoke_advection_at (fluid.cu:569-571)
  s->parity ^= 2;
  return SAYAL_OK;
}

// Local rows that are exact when the ghost rows are valid to depth `depth` beyond the owned rows.
static void set_valid_depth(Sim* s, int depth) {
  Grid& g = s->g;
  g.valid_lo = g.own_lo - depth < 0 ? 0 : g.own_lo - depth;
  g.valid_hi = g.own_hi + depth > g.local_rows ? g.local_rows : g.own_hi + depth;
}

static int advect_rows(Sim* s, float d_t, bool smoke, int lo, int hi) {
  if (s->advect_kernel == 2 && s->g.h == 1) return launch_advect_geo_rows(s, d_t, smoke, lo, hi);
  // the other kernels advect [own_lo, own_hi): narrow / widen that window for the call
  Grid saved = s->g;
  s->g.own_lo = lo;
  s->g.own_hi = hi;
  int r = s->advect_kernel == 1 ? launch_advect_tile(s, d_t, !smoke, smoke) : launch_advect(s, d_t, !smoke, smoke);
  s->g.own_lo = saved.own_lo;
  s->g.own_hi = saved.own_hi;
  return r;
}

// Advection of a linked slab.  `ghost` extra rows beyond the owned ones are advected too (the smoke sampler
// reads the NEW velocity one row above the cell).  With `exchange_mask` the stage ends the step: the rows next
// to the slab edges are advected first and the exchange of the new arrays runs on the aux stream while the
// interior rows are still being advected.
static int advect_linked(Sim* s, float d_t, bool smoke, int ghost, int exchange_mask) {
  const Grid& g = s->g;
  const int halo = s->slab_halo;
  const int lo = g.own_lo - ghost < 0 ? 0 : g.own_lo - ghost;
  const int hi = g.own_hi + ghost > g.local_rows ? g.local_rows : g.own_hi + ghost;
  const bool overlap = exchange_mask && s->overlap_exchange && s->aux_stream && g.own_hi - g.own_lo >= 4 * halo;
  auto swap_fields = [&]() {
    if (smoke) {
      swap_ptr(s->smoke, s->smoke_buf);
      s->parity ^= 2;
    } else {
      swap_ptr(s->u, s->u_buf);
      swap_ptr(s->v, s->v_buf);
      s->parity ^= 1;
    }
  };
  if (overlap) {
    TRY(advect_rows(s, d_t, smoke, lo, g.own_lo + halo));
    TRY(advect_rows(s, d_t, smoke, g.own_hi - halo, hi));
    CUDA_TRY(cudaEventRecord(s->ev_fork, s->stream));
    CUDA_TRY(cudaStreamWaitEvent(s->aux_stream, s->ev_fork, 0));
    dbg_mark(s, smoke ? 4 : 3, s->stream);  // edge rows advected
    // The exchange is enqueued BEFORE the interior rows: its few fat CTAs must find room while the SMs are empty
    // (behind a grid of thousands of small blocks they would only be placed when that grid drains, i.e. the
    // exchange would run after the advection instead of under it).  It reads the edge rows the two kernels above
    // wrote into the back buffers and writes their ghost rows; the interior kernel touches neither.
    swap_fields();  // the exchange sees the new arrays ...
    int r = launch_slab_exchange_on(s, exchange_mask, s->aux_stream);
    swap_fields();  // ... the interior kernel the old ones
    if (r != SAYAL_OK) return r;
    CUDA_TRY(cudaEventRecord(s->ev_join, s->aux_stream));
    dbg_mark(s, 5, s->aux_stream);  // exchange done
    TRY(advect_rows(s, d_t, smoke, g.own_lo + halo, g.own_hi - halo));
    swap_fields();
    dbg_mark(s, 6, s->stream);  // interior rows advected
    CUDA_TRY(cudaStreamWaitEvent(s->stream, s->ev_join, 0));
  } else {
    TRY(advect_rows(s, d_t, smoke, lo, hi));
    swap_fields();
    if (exchange_mask) TRY(launch_slab_exchange(s, exchange_mask));
  }
  return SAYAL_OK;
}

bool is_linked(const Sim* s);
bool push_mode(const Sim* s);

static int step_impl(Sim* s, const sayal_source* src, float d_t) {
  const bool linked = is_linked(s);
  const int skip = s->debug_skip;  // profiling only: marginal cost of a stage inside the replayed graph
  s->dbg_marked = 0;
  dbg_mark(s, 0, s->stream);
  if (!(skip & 1)) TRY(launch_forces(s, src, d_t, true));  // may defer to the first projection pass (fuse_forces)
  if (s->ph.enable_pressure) TRY(launch_zero_pressure(s));
  // apply_diffusion (fluid.cu:775-777): n sweeps over u when viscosity != 0, in a fixed red-black order (H1).
  // A sweep has the same one-row dependency radius per colour as a projection half-sweep, so linked slabs chunk it
  // the same way: halo/2 sweeps, then an exchange of u.
  if (s->cfg.viscosity != 0.f) {
    if (!linked) {
      TRY(launch_diffusion(s, s->cfg.proj_n, d_t));
    } else {
      for (int done = 0; done < s->cfg.proj_n;) {
        int k = s->cfg.proj_n - done < s->slab_halo / 2 ? s->cfg.proj_n - done : s->slab_halo / 2;
        TRY(launch_diffusion(s, k, d_t));
        TRY(launch_slab_exchange(s, 1));
        done += k;
      }
    }
  }
  // y-slab (linked): ghost rows are exact to depth D beyond the owned rows; every SOR iteration costs two rows
  // of depth, an exchange restores D = halo.  Exchanges happen only when the next operation needs more depth
  // than is left, and once at the end of the step (u, v and smoke together).
  int D = s->slab_halo;
  // the pass that ends the projection also applies the boundary extrapolation (option fuse_extrapolation)
  const bool fold_extrap = s->fuse_extrapolation && s->projection_kernel == 1 && !s->ph.enable_pressure && s->cfg.proj_n > 0;
  s->fuse_extrap = 0;
  if (!linked) {
    s->fuse_extrap = fold_extrap && !(skip & 2);
    if (!(skip & 16)) TRY(projection(s, s->cfg.proj_n, d_t));
  } else if (push_mode(s)) {
    // Push mode: every pass consumes 2 it ghost rows and its edge tiles store the slab's new edge rows straight into
    // the neighbours' ghost rows (projection_pack.cu), so the whole projection is ONE call with thin windows and the
    // ghost rows are exact to the full halo again when it returns.
    s->fuse_extrap = fold_extrap;
    s->push_active = 1;
    int r = (skip & 16) ? SAYAL_OK : projection(s, s->cfg.proj_n, d_t);
    s->push_active = 0;
    if (r != SAYAL_OK) return r;
    D = s->slab_halo;
  } else {
    for (int done = 0; done < s->cfg.proj_n;) {
      if (D < 2) {
        TRY(launch_slab_exchange(s, 1 | 2));
        D = s->slab_halo;
      }
      int k = s->cfg.proj_n - done < D / 2 ? s->cfg.proj_n - done : D / 2;
      s->fuse_extrap = fold_extrap && done + k == s->cfg.proj_n;
      s->proj_depth = s->shrink_window ? D : -1;  // sweep only the ghost rows that are still exact (projection_pack.cu)
      if (!(skip & 16)) TRY(projection(s, k, d_t));
      s->proj_depth = -1;
      D -= 2 * k;
      done += k;
    }
  }
  dbg_mark(s, 1, s->stream);  // projection done
  bool range_forked = false;
  if (s->ph.enable_pressure) {
    TRY(launch_pressure_range(s));
    s->range_valid = false;
    if (linked) {  // one min / max for the whole frame: chain reduction on the aux stream, joined at the end of the step
      CUDA_TRY(cudaEventRecord(s->ev_range, s->stream));
      CUDA_TRY(cudaStreamWaitEvent(s->aux_stream, s->ev_range, 0));
      TRY(launch_slab_range_reduce(s, s->aux_stream));
      CUDA_TRY(cudaEventRecord(s->ev_range, s->aux_stream));
      range_forked = true;
    }
  }
  if (s->fuse_extrap != 2 && !(skip & 2)) TRY(launch_extrapolation(s));
  s->fuse_extrap = 0;
  if (!linked) {
    if (!(skip & 4)) TRY(advect_velocity(s, d_t));
    if (s->ph.enable_smoke && s->ph.wt_smoke != 0.f && !(skip & 8)) TRY(advect_smoke(s, d_t));  // decay fused (fluid.cu:792)
  } else {
    const bool smoke = s->ph.enable_smoke && s->ph.wt_smoke != 0.f && !(skip & 8);
    const int end_mask = (skip & 32) ? 0 : (1 | 2);  // profiling only: leave the end-of-step exchange out
    // velocity is advected on the owned rows and one ghost row each side; its gathers reach advect_margin rows
    if (D < s->advect_margin + 1) {
      TRY(launch_slab_exchange(s, 1 | 2));
      D = s->slab_halo;
    }
    set_valid_depth(s, D);
    int r = (skip & 4) ? SAYAL_OK : advect_linked(s, d_t, false, 1, smoke ? 0 : end_mask);
    dbg_mark(s, 2, s->stream);  // velocity advected
    if (r == SAYAL_OK && smoke) {
      // new velocity: exact on the owned rows +- 1; smoke ghosts: exact as deep as exchanges carry smoke
      // (advect_margin + 2 rows, slab_exchange.cu) — a gather beyond that counts as halo overflow
      const int smoke_depth = s->slab_halo < s->advect_margin + 2 ? s->slab_halo : s->advect_margin + 2;
      set_valid_depth(s, D < smoke_depth ? D : smoke_depth);
      r = advect_linked(s, d_t, true, 0, end_mask ? (end_mask | 4) : 0);
    }
    s->g.valid_lo = 0;
    s->g.valid_hi = s->g.local_rows;
    if (r != SAYAL_OK) return r;
  }
  if (range_forked) CUDA_TRY(cudaStreamWaitEvent(s->stream, s->ev_range, 0));
  dbg_mark(s, 7, s->stream);  // step done
  return SAYAL_OK;
}

static bool is_slab(const Sim* s) { return s->g.local_rows != s->g.H; }
bool is_linked(const Sim* s) { return s->link_block && (s->link.peer_recv[0] || s->link.peer_recv[1]); }

// Do the projection passes of a linked step push their edge rows themselves?  Needs the tiled kernel, the
// neighbours' velocity arrays (every link made through this library has them) and at least two ghost rows.
bool push_mode(const Sim* s) {
  if (!is_linked(s) || !s->slab_push || s->projection_kernel != 1 || s->cfg.proj_n <= 0 || s->slab_halo < 2) return false;
  if (s->ph.enable_pressure) return false;  // the pressure instantiation has no push code: exchange kernels between chunks
  for (int d = 0; d < 2; d++)
    if (s->link.peer_words[d] && !s->peer_vel[d][0]) return false;
  return true;
}

// The sticky link error of a slab sim (mapped host word, written by the kernels), as an error code.
static int link_status(const Sim* s, const char* where) {
  if (s->h_resident_error && *reinterpret_cast<volatile int*>(s->h_resident_error) != 0) {
    char m[256];
    snprintf(m, sizeof m, "%s: a tile of the resident projection waited in vain for its neighbours; fields are not valid", where);
    return set_error(SAYAL_ECUDA, m);
  }
  if (!s->h_link_error || *reinterpret_cast<volatile int*>(s->h_link_error) == LINK_OK) return SAYAL_OK;
  char m[256];
  snprintf(m, sizeof m, "%s: slab link broken (%s); fields are not valid", where,
           *s->h_link_error == LINK_PLAN_MISMATCH ? "neighbours split the projection into different passes"
                                                  : "a neighbour did not answer within 2 s");
  return set_error(SAYAL_ELINK, m);
}

}  // namespace sayal

using namespace sayal;

struct sayal_sim {
  Sim impl;
};
static inline Sim* S(sayal_sim* p) { return reinterpret_cast<Sim*>(p); }

extern "C" {

int sayal_abi_version(void) { return SAYAL_ABI_VERSION; }
const char* sayal_last_error(void) { return g_last_error.c_str(); }

int sayal_create(const sayal_config* cfg, int32_t device, sayal_sim** out) {
  Sim* s = nullptr;
  int r = create_impl(cfg, device, nullptr, &s);
  if (r == SAYAL_OK) *out = reinterpret_cast<sayal_sim*>(s);
  return r;
}

int sayal_create_slab(const sayal_config* cfg, int32_t device, const sayal_slab* slab, sayal_sim** out) {
  Sim* s = nullptr;
  int r = create_impl(cfg, device, slab, &s);
  if (r == SAYAL_OK) *out = reinterpret_cast<sayal_sim*>(s);
  return r;
}

void sayal_destroy(sayal_sim* sim) { free_sim(S(sim)); }

int sayal_step(sayal_sim* sim, const sayal_source* src, float d_t) {
  if (!sim) return set_error(SAYAL_EINVAL, "sayal_step: null sim");
  Sim* s = S(sim);
  if (is_slab(s) && !is_linked(s)) return set_error(SAYAL_EINVAL, "sayal_step: a slab sim must be linked to its neighbours first (sayal_slab_ipc_connect / sayal_slab_connect_local), or stepped stage by stage");
  if (is_linked(s) && s->slab_halo < s->advect_margin + 1) return set_error(SAYAL_EINVAL, "sayal_step: linked slabs need halo >= advect_margin + 1 (default 17)");
  TRY(link_status(s, "sayal_step"));
  CUDA_TRY(cudaSetDevice(s->device));
  s->steps_done++;
  return step_impl(s, src, d_t);
}

int sayal_run(sayal_sim* sim, int32_t steps, float d_t) {
  if (!sim) return set_error(SAYAL_EINVAL, "sayal_run: null sim");
  Sim* s = S(sim);
  if (is_slab(s) && !is_linked(s)) return set_error(SAYAL_EINVAL, "sayal_run: a slab sim must be linked to its neighbours first");
  if (is_linked(s) && s->slab_halo < s->advect_margin + 1) return set_error(SAYAL_EINVAL, "sayal_run: linked slabs need halo >= advect_margin + 1 (default 17)");
  if (steps < 0) return set_error(SAYAL_EINVAL, "sayal_run: steps < 0");
  CUDA_TRY(cudaSetDevice(s->device));
  if (s->graph_dt != d_t) {
    invalidate_graphs(s);
    s->graph_dt = d_t;
  }
  TRY(link_status(s, "sayal_run"));
  if (s->projection_kernel == 1) {  // timing is not capturable: choose every plan the step will use first
    if (!is_linked(s)) {
      TRY(tiled_prepare(s, s->cfg.proj_n));
    } else if (push_mode(s)) {
      s->push_active = 1;
      int r = tiled_prepare_windows(s, s->cfg.proj_n, -1);
      s->push_active = 0;
      if (r != SAYAL_OK) return r;
    } else {
      int D = s->slab_halo;
      for (int done = 0; done < s->cfg.proj_n;) {  // the chunk sizes step_impl will use
        if (D < 2) D = s->slab_halo;
        int k = s->cfg.proj_n - done < D / 2 ? s->cfg.proj_n - done : D / 2;
        TRY(tiled_prepare_windows(s, k, s->shrink_window ? D : -1));
        D -= 2 * k;
        done += k;
      }
    }
  }
  int remaining = steps;
  // One graph per starting parity holds ONE step; replaying it is followed by the same pointer swaps on
  // the host that capture performed, so the next replay (or eager call) sees the right front buffers.
  while (s->use_graph && remaining > 0) {
    int slot = s->parity & 3;
    if (!s->graph[slot]) {
      cudaGraph_t graph = nullptr;
      int64_t before = s->launches;
      CUDA_TRY(cudaStreamBeginCapture(s->stream, cudaStreamCaptureModeThreadLocal));
      int r = step_impl(s, nullptr, d_t);
      cudaError_t e = cudaStreamEndCapture(s->stream, &graph);
      s->graph_launches[slot] = s->launches - before;
      s->launches = before;  // capturing launches nothing
      if (r != SAYAL_OK) {
        if (graph) cudaGraphDestroy(graph);
        return r;
      }
      if (e != cudaSuccess) return set_error(SAYAL_ECUDA, cudaGetErrorString(e));
      e = cudaGraphInstantiate(&s->graph[slot], graph, 0);
      cudaGraphDestroy(graph);
      if (e != cudaSuccess) return set_error(SAYAL_ECUDA, cudaGetErrorString(e));
      s->graph_post[slot] = {s->u, s->v, s->u_buf, s->v_buf, s->smoke, s->smoke_buf, s->parity};
    }
    CUDA_TRY(cudaGraphLaunch(s->graph[slot], s->stream));
    const Sim::PtrState& ps = s->graph_post[slot];
    s->u = ps.u; s->v = ps.v; s->u_buf = ps.u_buf; s->v_buf = ps.v_buf;
    s->smoke = ps.smoke; s->smoke_buf = ps.smoke_buf; s->parity = ps.parity;
    s->launches += s->graph_launches[slot];
    if (s->ph.enable_pressure) s->range_valid = false;
    s->steps_done++;
    remaining--;
  }
  for (; remaining > 0; remaining--) {
    s->steps_done++;
    TRY(step_impl(s, nullptr, d_t));
  }
  return SAYAL_OK;
}

int sayal_sync(sayal_sim* sim) {
  if (!sim) return set_error(SAYAL_EINVAL, "sayal_sync: null sim");
  CUDA_TRY(cudaStreamSynchronize(S(sim)->stream));
  return link_status(S(sim), "sayal_sync");
}


static int field_info(Sim* s, int field, void** ptr, size_t* elem, bool* dense) {
  *elem = 4;
  *dense = false;
  switch (field) {
    case SAYAL_U: *ptr = s->u; return SAYAL_OK;
    case SAYAL_V: *ptr = s->v; return SAYAL_OK;
    case SAYAL_P: *ptr = s->p; return SAYAL_OK;
    case SAYAL_SMOKE: *ptr = s->smoke; return SAYAL_OK;
    case SAYAL_IS_SOLID:
    case SAYAL_TOTAL_S: {
      if (!s->d_is_solid) {  // built on demand: the step path itself only needs the 1-byte flags
        size_t n = (size_t)s->g.W * s->g.local_rows * sizeof(int32_t);
        CUDA_TRY(cudaMalloc(&s->d_is_solid, n));
        CUDA_TRY(cudaMalloc(&s->d_total_s, n));
        TRY(launch_export_masks(s));
      }
      *ptr = field == SAYAL_IS_SOLID ? s->d_is_solid : s->d_total_s;
      *dense = true;
      return SAYAL_OK;
    }
  }
  return set_error(SAYAL_EINVAL, "unknown field id");
}

int sayal_get_field(sayal_sim* sim, int32_t field, void* host_dst) {
  if (!sim || !host_dst) return set_error(SAYAL_EINVAL, "sayal_get_field: null argument");
  Sim* s = S(sim);
  CUDA_TRY(cudaSetDevice(s->device));
  void* p;
  size_t elem;
  bool dense;
  TRY(field_info(s, field, &p, &elem, &dense));
  size_t src_pitch = (dense ? s->g.W : s->g.pitch) * elem;
  const char* src = (const char*)p + (size_t)s->g.own_lo * src_pitch;
  const size_t row_bytes = s->g.W * elem, rows = s->g.own_hi - s->g.own_lo;
  // in stream order behind the step that produced the field; one contiguous copy when rows are not padded
  if (src_pitch == row_bytes) CUDA_TRY(cudaMemcpyAsync(host_dst, src, row_bytes * rows, cudaMemcpyDeviceToHost, s->stream));
  else CUDA_TRY(cudaMemcpy2DAsync(host_dst, row_bytes, src, src_pitch, row_bytes, rows, cudaMemcpyDeviceToHost, s->stream));
  CUDA_TRY(cudaStreamSynchronize(s->stream));
  return link_status(s, "sayal_get_field");
}

// Batched forms of the two calls above: every copy is enqueued on the sim's stream back to back and the call waits
// once (sayal_get_fields) or not at all beyond what the runtime needs (sayal_set_fields with pinned memory returns
// when the copies are enqueued; the caller must keep the buffers untouched until sayal_sync or any synchronous call).
static int copy_field(Sim* s, int field, void* host, bool to_host) {
  void* p = nullptr;
  size_t elem;
  bool dense;
  TRY(field_info(s, field, &p, &elem, &dense));
  const size_t dev_pitch = (dense ? s->g.W : s->g.pitch) * elem;
  char* dev = (char*)p + (size_t)s->g.own_lo * dev_pitch;
  const size_t row_bytes = s->g.W * elem, rows = s->g.own_hi - s->g.own_lo;
  if (to_host) {
    if (dev_pitch == row_bytes) CUDA_TRY(cudaMemcpyAsync(host, dev, row_bytes * rows, cudaMemcpyDeviceToHost, s->stream));
    else CUDA_TRY(cudaMemcpy2DAsync(host, row_bytes, dev, dev_pitch, row_bytes, rows, cudaMemcpyDeviceToHost, s->stream));
  } else {
    if (dev_pitch == row_bytes) CUDA_TRY(cudaMemcpyAsync(dev, host, row_bytes * rows, cudaMemcpyHostToDevice, s->stream));
    else CUDA_TRY(cudaMemcpy2DAsync(dev, dev_pitch, host, row_bytes, row_bytes, rows, cudaMemcpyHostToDevice, s->stream));
  }
  return SAYAL_OK;
}

// Device-to-device forms for a caller that already lives on the GPU (a renderer, a torch tensor): `dev` holds the
// owned rows in the reference layout (row pitch W), on the same device; ordered on the sim's stream, asynchronous.
static int copy_field_device(Sim* s, int field, void* dev_buf, bool to_buf) {
  void* p = nullptr;
  size_t elem;
  bool dense;
  TRY(field_info(s, field, &p, &elem, &dense));
  const size_t dev_pitch = (dense ? s->g.W : s->g.pitch) * elem;
  char* mine = (char*)p + (size_t)s->g.own_lo * dev_pitch;
  const size_t row_bytes = s->g.W * elem, rows = s->g.own_hi - s->g.own_lo;
  if (to_buf) CUDA_TRY(cudaMemcpy2DAsync(dev_buf, row_bytes, mine, dev_pitch, row_bytes, rows, cudaMemcpyDeviceToDevice, s->stream));
  else CUDA_TRY(cudaMemcpy2DAsync(mine, dev_pitch, dev_buf, row_bytes, row_bytes, rows, cudaMemcpyDeviceToDevice, s->stream));
  return SAYAL_OK;
}

int sayal_get_field_device(sayal_sim* sim, int32_t field, void* dev_dst) {
  if (!sim || !dev_dst) return set_error(SAYAL_EINVAL, "sayal_get_field_device: null argument");
  Sim* s = S(sim);
  CUDA_TRY(cudaSetDevice(s->device));
  return copy_field_device(s, field, dev_dst, true);
}

int sayal_set_field_device(sayal_sim* sim, int32_t field, const void* dev_src) {
  if (!sim || !dev_src) return set_error(SAYAL_EINVAL, "sayal_set_field_device: null argument");
  if (field < SAYAL_U || field > SAYAL_SMOKE) return set_error(SAYAL_EINVAL, "sayal_set_field_device: only U, V, P, SMOKE are writable");
  Sim* s = S(sim);
  CUDA_TRY(cudaSetDevice(s->device));
  return copy_field_device(s, field, const_cast<void*>(dev_src), false);
}

int sayal_get_fields(sayal_sim* sim, int32_t n, const int32_t* fields, void* const* host_dsts) {
  if (!sim || n < 0 || (n > 0 && (!fields || !host_dsts))) return set_error(SAYAL_EINVAL, "sayal_get_fields: bad argument");
  Sim* s = S(sim);
  CUDA_TRY(cudaSetDevice(s->device));
  for (int k = 0; k < n; k++) {
    if (!host_dsts[k]) return set_error(SAYAL_EINVAL, "sayal_get_fields: null destination");
    TRY(copy_field(s, fields[k], host_dsts[k], true));
  }
  CUDA_TRY(cudaStreamSynchronize(s->stream));
  return link_status(s, "sayal_get_fields");
}

int sayal_set_fields(sayal_sim* sim, int32_t n, const int32_t* fields, const void* const* host_srcs) {
  if (!sim || n < 0 || (n > 0 && (!fields || !host_srcs))) return set_error(SAYAL_EINVAL, "sayal_set_fields: bad argument");
  Sim* s = S(sim);
  CUDA_TRY(cudaSetDevice(s->device));
  for (int k = 0; k < n; k++) {
    if (!host_srcs[k]) return set_error(SAYAL_EINVAL, "sayal_set_fields: null source");
    if (fields[k] < SAYAL_U || fields[k] > SAYAL_SMOKE) return set_error(SAYAL_EINVAL, "sayal_set_fields: only U, V, P, SMOKE are writable");
    TRY(copy_field(s, fields[k], const_cast<void*>(host_srcs[k]), false));
  }
  return SAYAL_OK;
}

int sayal_set_field(sayal_sim* sim, int32_t field, const void* host_src) {
  if (!sim || !host_src) return set_error(SAYAL_EINVAL, "sayal_set_field: null argument");
  Sim* s = S(sim);
  if (field < SAYAL_U || field > SAYAL_SMOKE) return set_error(SAYAL_EINVAL, "sayal_set_field: only U, V, P, SMOKE are writable (masks derive from the config)");
  CUDA_TRY(cudaSetDevice(s->device));
  void* p;
  size_t elem;
  bool dense;
  TRY(field_info(s, field, &p, &elem, &dense));
  char* dst = (char*)p + (size_t)s->g.own_lo * s->g.pitch * elem;
  const size_t row_bytes = s->g.W * elem, rows = s->g.own_hi - s->g.own_lo;
  // In stream order: the copy waits for the steps already enqueued, later steps wait for it.  The host buffer
  // may be reused when the call returns (pageable memory is staged by the runtime; for pinned memory we wait).
  if (s->g.pitch * elem == row_bytes) CUDA_TRY(cudaMemcpyAsync(dst, host_src, row_bytes * rows, cudaMemcpyHostToDevice, s->stream));
  else CUDA_TRY(cudaMemcpy2DAsync(dst, s->g.pitch * elem, host_src, row_bytes, row_bytes, rows, cudaMemcpyHostToDevice, s->stream));
  CUDA_TRY(cudaStreamSynchronize(s->stream));
  return SAYAL_OK;
}

int sayal_device_ptr(sayal_sim* sim, int32_t field, void** dev_ptr, int64_t* pitch_elems, int32_t* first_row,
                     int32_t* n_rows) {
  if (!sim || !dev_ptr) return set_error(SAYAL_EINVAL, "sayal_device_ptr: null argument");
  Sim* s = S(sim);
  CUDA_TRY(cudaSetDevice(s->device));
  void* p;
  size_t elem;
  bool dense;
  TRY(field_info(s, field, &p, &elem, &dense));
  *dev_ptr = p;
  if (pitch_elems) *pitch_elems = dense ? s->g.W : s->g.pitch;
  if (first_row) *first_row = s->g.row_base;
  if (n_rows) *n_rows = s->g.local_rows;
  return SAYAL_OK;
}

static float from_ordered(int32_t e) {
  int32_t b = e ^ ((e >> 31) & 0x7fffffff);
  float f;
  std::memcpy(&f, &b, 4);
  return f;
}

int sayal_pressure_range(sayal_sim* sim, float* min_p, float* max_p) {
  if (!sim) return set_error(SAYAL_EINVAL, "sayal_pressure_range: null sim");
  Sim* s = S(sim);
  CUDA_TRY(cudaSetDevice(s->device));
  if (!s->range_valid) {
    int32_t r[2];
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    TRY(link_status(s, "sayal_pressure_range"));
    // linked slabs: the range of the whole domain (reduced along the chain by the step), like the reference's one pair
    CUDA_TRY(cudaMemcpy(r, s->d_range + (is_linked(s) ? 2 : 0), sizeof r, cudaMemcpyDeviceToHost));
    s->min_p = from_ordered(r[0]);
    s->max_p = from_ordered(r[1]);
    s->range_valid = true;
  }
  if (min_p) *min_p = s->min_p;
  if (max_p) *max_p = s->max_p;
  return SAYAL_OK;
}

int sayal_sample_velocity(sayal_sim* sim, int32_t n, const float* xs, const float* ys, float* out_u, float* out_v) {
  if (!sim || n < 0 || (n > 0 && (!xs || !ys || !out_u || !out_v)))
    return set_error(SAYAL_EINVAL, "sayal_sample_velocity: bad argument");
  if (n == 0) return SAYAL_OK;
  Sim* s = S(sim);
  CUDA_TRY(cudaSetDevice(s->device));
  float* d = nullptr;
  CUDA_TRY(cudaMalloc(&d, sizeof(float) * 4 * (size_t)n));
  cudaMemcpyAsync(d, xs, sizeof(float) * n, cudaMemcpyHostToDevice, s->stream);
  cudaMemcpyAsync(d + n, ys, sizeof(float) * n, cudaMemcpyHostToDevice, s->stream);
  int r = launch_sample_velocity(s, n, d, d + n, d + 2 * (size_t)n, d + 3 * (size_t)n);
  if (r == SAYAL_OK) {
    cudaMemcpyAsync(out_u, d + 2 * (size_t)n, sizeof(float) * n, cudaMemcpyDeviceToHost, s->stream);
    cudaMemcpyAsync(out_v, d + 3 * (size_t)n, sizeof(float) * n, cudaMemcpyDeviceToHost, s->stream);
  }
  cudaError_t e = cudaStreamSynchronize(s->stream);
  cudaFree(d);
  if (r != SAYAL_OK) return r;
  if (e != cudaSuccess) return set_error(SAYAL_ECUDA, cudaGetErrorString(e));
  return SAYAL_OK;
}

// ---- stages ---------------------------------------------------------------------------------------
#define STAGE_PROLOGUE(name)                                          \
  if (!sim) return set_error(SAYAL_EINVAL, name ": null sim");        \
  Sim* s = S(sim);                                                    \
  CUDA_TRY(cudaSetDevice(s->device));

int sayal_stage_forces(sayal_sim* sim, const sayal_source* src, float d_t) {
  STAGE_PROLOGUE("sayal_stage_forces");
  return launch_forces(s, src, d_t);
}
int sayal_stage_zero_pressure(sayal_sim* sim) {
  STAGE_PROLOGUE("sayal_stage_zero_pressure");
  return launch_zero_pressure(s);
}
int sayal_stage_projection(sayal_sim* sim, int32_t iterations, float d_t) {
  STAGE_PROLOGUE("sayal_stage_projection");
  TRY(projection(s, iterations, d_t));
  if (s->ph.enable_pressure) {
    TRY(launch_pressure_range(s));
    s->range_valid = false;
  }
  return SAYAL_OK;
}
int sayal_stage_extrapolation(sayal_sim* sim) {
  STAGE_PROLOGUE("sayal_stage_extrapolation");
  return launch_extrapolation(s);
}
int sayal_stage_advect_velocity(sayal_sim* sim, float d_t) {
  STAGE_PROLOGUE("sayal_stage_advect_velocity");
  return advect_velocity(s, d_t);
}
int sayal_stage_advect_smoke(sayal_sim* sim, float d_t) {
  STAGE_PROLOGUE("sayal_stage_advect_smoke");
  return advect_smoke(s, d_t);
}

int sayal_stage_diffusion(sayal_sim* sim, int32_t iterations, float d_t) {
  STAGE_PROLOGUE("sayal_stage_diffusion");
  if (iterations < 0) return set_error(SAYAL_EINVAL, "sayal_stage_diffusion: iterations < 0");
  return launch_diffusion(s, iterations, d_t);
}

// ---- renderer-side consumers (visual.cu) --------------------------------------------------------------
static size_t frame_pixels(const Sim* s) { return (size_t)s->g.W * (s->g.own_hi - s->g.own_lo); }

static int frame_ring_init(Sim* s) {
  if (s->copy_stream) return SAYAL_OK;
  CUDA_TRY(cudaStreamCreateWithFlags(&s->copy_stream, cudaStreamNonBlocking));
  for (int k = 0; k < Sim::kFrames; k++) {
    CUDA_TRY(cudaMalloc(&s->d_frame[k], frame_pixels(s) * sizeof(uint32_t)));
    CUDA_TRY(cudaMemsetAsync(s->d_frame[k], 0, frame_pixels(s) * sizeof(uint32_t), s->stream));
    CUDA_TRY(cudaHostAlloc(&s->h_frame[k], frame_pixels(s) * sizeof(uint32_t), cudaHostAllocDefault));
    CUDA_TRY(cudaEventCreateWithFlags(&s->ev_rendered[k], cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&s->ev_copied[k], cudaEventDisableTiming));
  }
  return SAYAL_OK;
}

int sayal_frame_submit(sayal_sim* sim) {
  STAGE_PROLOGUE("sayal_frame_submit");
  TRY(frame_ring_init(s));
  if (s->frame_pending >= 2) return set_error(SAYAL_EBUSY, "sayal_frame_submit: two frames outstanding, acquire one first");
  const int k = s->frame_head;
  // the device frame may still be the source of the copy submitted three frames ago
  if (s->frame_used[k]) CUDA_TRY(cudaStreamWaitEvent(s->stream, s->ev_copied[k], 0));
  TRY(launch_render_pixels(s, s->d_frame[k]));
  CUDA_TRY(cudaEventRecord(s->ev_rendered[k], s->stream));
  CUDA_TRY(cudaStreamWaitEvent(s->copy_stream, s->ev_rendered[k], 0));
  CUDA_TRY(cudaMemcpyAsync(s->h_frame[k], s->d_frame[k], frame_pixels(s) * sizeof(uint32_t), cudaMemcpyDeviceToHost,
                           s->copy_stream));
  CUDA_TRY(cudaEventRecord(s->ev_copied[k], s->copy_stream));
  s->frame_used[k] = true;
  s->frame_step[k] = s->steps_done;
  s->frame_head = (s->frame_head + 1) % Sim::kFrames;
  s->frame_pending++;
  return SAYAL_OK;
}

int sayal_frame_acquire(sayal_sim* sim, const uint32_t** pixels, int64_t* step_index) {
  STAGE_PROLOGUE("sayal_frame_acquire");
  if (!pixels) return set_error(SAYAL_EINVAL, "sayal_frame_acquire: null argument");
  if (s->frame_pending <= 0) return set_error(SAYAL_EINVAL, "sayal_frame_acquire: no frame was submitted");
  const int k = (s->frame_head + Sim::kFrames - s->frame_pending) % Sim::kFrames;  // the oldest outstanding slot
  CUDA_TRY(cudaEventSynchronize(s->ev_copied[k]));
  *pixels = s->h_frame[k];
  if (step_index) *step_index = s->frame_step[k];
  s->frame_pending--;
  return SAYAL_OK;
}

int sayal_render_pixels(sayal_sim* sim, uint32_t* host_dst) {
  STAGE_PROLOGUE("sayal_render_pixels");
  if (!host_dst) return set_error(SAYAL_EINVAL, "sayal_render_pixels: null destination");
  uint32_t* d = nullptr;
  const size_t bytes = frame_pixels(s) * sizeof(uint32_t);
  CUDA_TRY(cudaMalloc(&d, bytes));
  // cells the reference leaves untouched (neither smoke nor pressure enabled) keep the caller's pixels
  cudaMemcpyAsync(d, host_dst, bytes, cudaMemcpyHostToDevice, s->stream);
  int r = launch_render_pixels(s, d);
  if (r == SAYAL_OK) cudaMemcpyAsync(host_dst, d, bytes, cudaMemcpyDeviceToHost, s->stream);
  cudaError_t e = cudaStreamSynchronize(s->stream);
  cudaFree(d);
  if (r != SAYAL_OK) return r;
  if (e != cudaSuccess) return set_error(SAYAL_ECUDA, cudaGetErrorString(e));
  return SAYAL_OK;
}

int sayal_arrows(sayal_sim* sim, const sayal_visual* v, sayal_arrow* host_dst, int32_t capacity, int32_t* n_x, int32_t* n_y) {
  STAGE_PROLOGUE("sayal_arrows");
  if (!v || v->arrows_distance < 1 || v->cell_pixel_size < 1) return set_error(SAYAL_EINVAL, "sayal_arrows: arrows.distance and cell_pixel_size must be >= 1");
  const int nx = s->g.W / v->arrows_distance, ny = s->g.H / v->arrows_distance;
  if (n_x) *n_x = nx;
  if (n_y) *n_y = ny;
  if (!host_dst) return SAYAL_OK;  // size query
  if ((int64_t)capacity < (int64_t)nx * ny) return set_error(SAYAL_EINVAL, "sayal_arrows: destination too small");
  if (nx * ny == 0) return SAYAL_OK;
  sayal_arrow* d = nullptr;
  const size_t bytes = sizeof(sayal_arrow) * (size_t)nx * ny;
  CUDA_TRY(cudaMalloc(&d, bytes));
  int r = launch_arrows(s, v, nx, ny, d);
  if (r == SAYAL_OK) cudaMemcpyAsync(host_dst, d, bytes, cudaMemcpyDeviceToHost, s->stream);
  cudaError_t e = cudaStreamSynchronize(s->stream);
  cudaFree(d);
  if (r != SAYAL_OK) return r;
  if (e != cudaSuccess) return set_error(SAYAL_ECUDA, cudaGetErrorString(e));
  return SAYAL_OK;
}

int sayal_path_lines(sayal_sim* sim, const sayal_visual* v, float d_t, int32_t* host_x, int32_t* host_y, int32_t capacity,
                     int32_t* n_x, int32_t* n_y) {
  STAGE_PROLOGUE("sayal_path_lines");
  if (!v || v->path_line_distance < 1 || v->path_line_length < 1) return set_error(SAYAL_EINVAL, "sayal_path_lines: path_line.distance and .length must be >= 1");
  const int nx = s->g.W / v->path_line_distance, ny = s->g.H / v->path_line_distance, len = v->path_line_length;
  if (n_x) *n_x = nx;
  if (n_y) *n_y = ny;
  if (!host_x && !host_y) return SAYAL_OK;  // size query
  if (!host_x || !host_y) return set_error(SAYAL_EINVAL, "sayal_path_lines: null destination");
  const int64_t total = (int64_t)nx * ny * len;
  if ((int64_t)capacity < total) return set_error(SAYAL_EINVAL, "sayal_path_lines: destination too small");
  if (total == 0) return SAYAL_OK;
  int32_t* d = nullptr;
  CUDA_TRY(cudaMalloc(&d, sizeof(int32_t) * 2 * (size_t)total));
  int r = launch_path_lines(s, v->path_line_distance, len, d_t, nx, ny, d, d + total);
  if (r == SAYAL_OK) {
    cudaMemcpyAsync(host_x, d, sizeof(int32_t) * total, cudaMemcpyDeviceToHost, s->stream);
    cudaMemcpyAsync(host_y, d + total, sizeof(int32_t) * total, cudaMemcpyDeviceToHost, s->stream);
  }
  cudaError_t e = cudaStreamSynchronize(s->stream);
  cudaFree(d);
  if (r != SAYAL_OK) return r;
  if (e != cudaSuccess) return set_error(SAYAL_ECUDA, cudaGetErrorString(e));
  return SAYAL_OK;
}

// ---- options --------------------------------------------------------------------------------------
int sayal_set_option(sayal_sim* sim, const char* key, int64_t value) {
  if (!sim || !key) return set_error(SAYAL_EINVAL, "sayal_set_option: null argument");
  Sim* s = S(sim);
  invalidate_graphs(s);
  if (!strcmp(key, "projection_kernel")) {
    if (value < 0 || value > 1) return set_error(SAYAL_EINVAL, "projection_kernel must be 0 or 1");
    s->projection_kernel = (int)value;
    s->plan_variant = -1, s->n_plans = 0;
  } else if (!strcmp(key, "temporal_block")) {
    if (value < 0 || value > tiled_max_temporal_block()) return set_error(SAYAL_EINVAL, "temporal_block out of range");
    s->temporal_block = (int)value;
    s->plan_variant = -1, s->n_plans = 0;
  } else if (!strcmp(key, "tile_rows_per_warp")) {  // 0 = any, else 8 / 10 / 12
    if (value != 0 && value != 8 && value != 10 && value != 12) return set_error(SAYAL_EINVAL, "tile_rows_per_warp must be 0, 8, 10 or 12");
    s->force_variant = value == 0 ? -1 : (int)(value - 8) / 2;
    s->plan_variant = -1, s->n_plans = 0;
  } else if (!strcmp(key, "autotune")) {
    s->autotune = value != 0;
    s->plan_variant = -1, s->n_plans = 0;
  } else if (!strcmp(key, "use_graph")) {
    s->use_graph = value != 0;
  } else if (!strcmp(key, "advect_kernel")) {
    if (value < 0 || value > 2) return set_error(SAYAL_EINVAL, "advect_kernel must be 0, 1 or 2");
    s->advect_kernel = (int)value;
  } else if (!strcmp(key, "advect_margin")) {
    if (value < 2) return set_error(SAYAL_EINVAL, "advect_margin must be >= 2");
    s->advect_margin = (int)value;
  } else if (!strcmp(key, "overlap_exchange")) {
    s->overlap_exchange = value != 0;
  } else if (!strcmp(key, "use_pdl")) {
    s->use_pdl = value != 0;
  } else if (!strcmp(key, "fuse_forces")) {
    s->fuse_forces = value != 0;
  } else if (!strcmp(key, "fuse_extrapolation")) {
    s->fuse_extrapolation = value != 0;
  } else if (!strcmp(key, "shrink_window")) {
    s->shrink_window = value != 0;
  } else if (!strcmp(key, "split_tiles")) {
    s->split_tiles = value != 0;
    s->plan_variant = -1, s->n_plans = 0;
  } else if (!strcmp(key, "debug_events")) {
    s->debug_events = value != 0;
  } else if (!strcmp(key, "resident")) {
    if (value < 0 || value > 2) return set_error(SAYAL_EINVAL, "resident must be 0 (off), 1 (candidate) or 2 (resident plans only)");
    s->resident = (int)value;
    s->n_plans = 0;  // plans are chosen with or without resident candidates
    invalidate_graphs(s);
  } else if (!strcmp(key, "slab_push")) {
    s->slab_push = value != 0;
    s->plan_variant = -1, s->n_plans = 0;
  } else if (!strcmp(key, "order_tiles")) {
    s->order_tiles = value != 0;
    s->plan_variant = -1, s->n_plans = 0;
  } else if (!strcmp(key, "debug_skip")) {
    s->debug_skip = (int)value;
  } else if (!strcmp(key, "debug_timeline")) {  // profiling only: per-CTA phase timestamps (sayal_debug_timeline)
    if (value && !s->d_timeline) {
      s->timeline_cap = 5 * 65536;
      CUDA_TRY(cudaMalloc(&s->d_timeline, s->timeline_cap * sizeof(long long)));
    } else if (!value && s->d_timeline) {
      cudaFree(s->d_timeline);
      s->d_timeline = nullptr;
    }
  } else {
    return set_error(SAYAL_EINVAL, "sayal_set_option: unknown key");
  }
  return SAYAL_OK;
}

int sayal_get_option(sayal_sim* sim, const char* key, int64_t* value) {
  if (!sim || !key || !value) return set_error(SAYAL_EINVAL, "sayal_get_option: null argument");
  Sim* s = S(sim);
  if (!strcmp(key, "projection_kernel")) *value = s->projection_kernel;
  else if (!strcmp(key, "temporal_block")) *value = s->temporal_block;
  else if (!strcmp(key, "use_graph")) *value = s->use_graph;
  else if (!strcmp(key, "use_pdl")) *value = s->use_pdl;
  else if (!strcmp(key, "fuse_forces")) *value = s->fuse_forces;
  else if (!strcmp(key, "order_tiles")) *value = s->order_tiles;
  else if (!strcmp(key, "shrink_window")) *value = s->shrink_window;
  else if (!strcmp(key, "slab_push")) *value = s->slab_push;
  else if (!strcmp(key, "resident")) *value = s->resident;
  else if (!strcmp(key, "split_tiles")) *value = s->split_tiles;
  else if (!strcmp(key, "plan_resident")) *value = s->plan_resident;
  else if (!strcmp(key, "push_mode")) *value = push_mode(s) ? 1 : 0;
  else if (!strcmp(key, "fuse_extrapolation")) *value = s->fuse_extrapolation;
  else if (!strcmp(key, "advect_kernel")) *value = s->advect_kernel;
  else if (!strcmp(key, "autotune")) *value = s->autotune;
  else if (!strcmp(key, "plan_temporal_block")) *value = s->plan_variant >= 0 ? s->plan_T : 0;
  else if (!strcmp(key, "plan_rows_per_warp")) *value = s->plan_variant >= 0 ? 8 + 2 * s->plan_variant : 0;
  else if (!strcmp(key, "halo_overflow")) {
    int32_t v = 0;
    CUDA_TRY(cudaSetDevice(s->device));
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    CUDA_TRY(cudaMemcpy(&v, s->d_overflow, sizeof v, cudaMemcpyDeviceToHost));
    *value = v;
  } else if (!strcmp(key, "link_error")) {
    int32_t v = 0;
    if (s->link_block) {
      CUDA_TRY(cudaSetDevice(s->device));
      CUDA_TRY(cudaStreamSynchronize(s->stream));
      v = *reinterpret_cast<volatile int*>(s->h_link_error);
    }
    *value = v;
  } else if (!strcmp(key, "pitch")) *value = s->g.pitch;
  else if (!strcmp(key, "local_rows")) *value = s->g.local_rows;
  else if (!strcmp(key, "own_lo")) *value = s->g.own_lo;
  else if (!strcmp(key, "own_hi")) *value = s->g.own_hi;
  else return set_error(SAYAL_EINVAL, "sayal_get_option: unknown key");
  return SAYAL_OK;
}

int sayal_plan_log(sayal_sim* sim, char* buf, int32_t capacity) {
  if (!sim) return set_error(SAYAL_EINVAL, "sayal_plan_log: null sim");
  Sim* s = S(sim);
  const int n = (int)strlen(s->plan_log);
  if (buf && capacity > 0) {
    const int m = n < capacity - 1 ? n : capacity - 1;
    std::memcpy(buf, s->plan_log, m);
    buf[m] = 0;
  }
  return n;
}

int sayal_debug_timeline(sayal_sim* sim, int64_t* host_dst, int32_t max_tiles, int32_t* n_tiles) {
  if (!sim || !host_dst || !n_tiles) return set_error(SAYAL_EINVAL, "sayal_debug_timeline: null argument");
  Sim* s = S(sim);
  *n_tiles = 0;
  if (!s->d_timeline || s->timeline_tiles <= 0) return SAYAL_OK;
  int n = s->timeline_tiles < max_tiles ? s->timeline_tiles : max_tiles;
  CUDA_TRY(cudaSetDevice(s->device));
  CUDA_TRY(cudaStreamSynchronize(s->stream));
  CUDA_TRY(cudaMemcpy(host_dst, s->d_timeline, sizeof(long long) * 5 * (size_t)n, cudaMemcpyDeviceToHost));
  *n_tiles = n;
  return SAYAL_OK;
}

// Measurement aid: hold the sim's stream for `microseconds` (one thread spinning on %globaltimer), so that a caller
// can enqueue a whole timed region before the device starts on it and host jitter cannot drain the queue.
int sayal_stream_delay(sayal_sim* sim, int64_t microseconds) {
  if (!sim || microseconds < 0 || microseconds > 1000000) return set_error(SAYAL_EINVAL, "sayal_stream_delay: 0..1e6 us");
  Sim* s = S(sim);
  CUDA_TRY(cudaSetDevice(s->device));
  stream_delay_kernel<<<1, 1, 0, s->stream>>>(microseconds * 1000);
  CUDA_TRY(cudaGetLastError());
  return SAYAL_OK;
}

int sayal_debug_tile_list(sayal_sim* sim, int32_t iterations_per_pass, int32_t* out, int32_t capacity, int32_t* n_tiles) {
  if (!sim || !out || !n_tiles || capacity < 0) return set_error(SAYAL_EINVAL, "sayal_debug_tile_list: bad argument");
  Sim* s = S(sim);
  CUDA_TRY(cudaSetDevice(s->device));
  int n = tiled_debug_tile_list(s, iterations_per_pass, out, capacity);
  if (n < 0) return set_error(SAYAL_ECUDA, "sayal_debug_tile_list: copy failed");
  *n_tiles = n;
  return SAYAL_OK;
}

int sayal_debug_stage_times(sayal_sim* sim, float* ms_out, int32_t capacity) {
  if (!sim || !ms_out) return set_error(SAYAL_EINVAL, "sayal_debug_stage_times: null argument");
  Sim* s = S(sim);
  CUDA_TRY(cudaSetDevice(s->device));
  CUDA_TRY(cudaStreamSynchronize(s->stream));
  if (s->aux_stream) CUDA_TRY(cudaStreamSynchronize(s->aux_stream));
  for (int k = 0; k < capacity; k++) {
    ms_out[k] = -1.f;
    if (k < Sim::kDbgEvents && (s->dbg_marked & 1u) && (s->dbg_marked & (1u << k)))
      if (cudaEventElapsedTime(&ms_out[k], s->dbg_ev[0], s->dbg_ev[k]) != cudaSuccess) { cudaGetLastError(); ms_out[k] = -1.f; }
  }
  return SAYAL_OK;
}

int sayal_debug_link_words(sayal_sim* sim, uint32_t* host_dst) {
  if (!sim || !host_dst) return set_error(SAYAL_EINVAL, "sayal_debug_link_words: null argument");
  Sim* s = S(sim);
  if (!s->link_block || !s->link_counters) return set_error(SAYAL_EINVAL, "sayal_debug_link_words: the sim has no slab link");
  CUDA_TRY(cudaSetDevice(s->device));
  CUDA_TRY(cudaStreamSynchronize(s->stream));
  CUDA_TRY(cudaMemcpy(host_dst, s->link_block, LW_WORDS * sizeof(unsigned), cudaMemcpyDeviceToHost));
  CUDA_TRY(cudaMemcpy(host_dst + LW_WORDS, s->link_counters, 32 * sizeof(unsigned), cudaMemcpyDeviceToHost));
  return SAYAL_OK;
}

// Host-released gate: a one-thread kernel on the sim's stream spins on a word in mapped host memory until
// sayal_stream_release stores the matching ticket.  Everything enqueued behind it starts exactly when the host says
// so — a measured region can be enqueued completely, the ranks of a multi-GPU run can meet at a host barrier, and
// only then do the devices begin.  The spin gives up after 20 s (a forgotten release must not wedge the GPU).
int sayal_stream_hold(sayal_sim* sim) {
  if (!sim) return set_error(SAYAL_EINVAL, "sayal_stream_hold: null sim");
  Sim* s = S(sim);
  CUDA_TRY(cudaSetDevice(s->device));
  if (!s->h_gate) {
    CUDA_TRY(cudaHostAlloc(reinterpret_cast<void**>(&s->h_gate), sizeof(unsigned), cudaHostAllocMapped));
    *s->h_gate = 0;
    CUDA_TRY(cudaHostGetDevicePointer(reinterpret_cast<void**>(&s->d_gate), s->h_gate, 0));
  }
  s->gate_ticket++;
  stream_gate_kernel<<<1, 1, 0, s->stream>>>(s->d_gate, s->gate_ticket);
  CUDA_TRY(cudaGetLastError());
  return SAYAL_OK;
}

int sayal_stream_release(sayal_sim* sim) {
  if (!sim) return set_error(SAYAL_EINVAL, "sayal_stream_release: null sim");
  Sim* s = S(sim);
  if (!s->h_gate) return set_error(SAYAL_EINVAL, "sayal_stream_release: no gate was set");
  __atomic_store_n(s->h_gate, s->gate_ticket, __ATOMIC_RELEASE);
  return SAYAL_OK;
}

int sayal_debug_pass_plans(int32_t pitch, int32_t local_rows, int32_t own_lo, int32_t own_hi, int32_t rows_per_warp,
                           int32_t temporal_block, int32_t iterations, int32_t ghost_depth, int32_t* out, int32_t capacity,
                           int32_t* n_passes) {
  if (!out || !n_passes || capacity < 1) return set_error(SAYAL_EINVAL, "sayal_debug_pass_plans: null argument");
  int n = tiled_debug_pass_plans(pitch, local_rows, own_lo, own_hi, rows_per_warp, temporal_block, iterations, ghost_depth,
                                 out, capacity);
  if (n < 0) return set_error(SAYAL_EINVAL, "sayal_debug_pass_plans: no such plan (rows per warp 8/10/12, T 1..16, pitch % 4 == 0, push mode: 2 T <= halo)");
  *n_passes = n;
  return SAYAL_OK;
}

int64_t sayal_launch_count(sayal_sim* sim) { return sim ? S(sim)->launches : 0; }
void* sayal_stream(sayal_sim* sim) { return sim ? (void*)S(sim)->stream : nullptr; }

// ---- slab links ---------------------------------------------------------------------------------------
static_assert(sizeof(LinkInfo) <= SAYAL_LINK_INFO_BYTES, "LinkInfo must fit the opaque blob of sayal.h");
static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");

int sayal_slab_ipc_export(sayal_sim* sim, void* info_out) {
  STAGE_PROLOGUE("sayal_slab_ipc_export");
  if (!info_out) return set_error(SAYAL_EINVAL, "sayal_slab_ipc_export: null argument");
  LinkInfo info;
  TRY(slab_link_export(s, &info));
  cudaIpcMemHandle_t h;
  CUDA_TRY(cudaIpcGetMemHandle(&h, s->link_block));
  std::memcpy(info.link_handle, &h, sizeof h);
  CUDA_TRY(cudaIpcGetMemHandle(&h, s->vel_block));
  std::memcpy(info.vel_handle, &h, sizeof h);
  std::memset(info_out, 0, SAYAL_LINK_INFO_BYTES);
  std::memcpy(info_out, &info, sizeof info);
  return SAYAL_OK;
}

int sayal_slab_ipc_connect(sayal_sim* sim, int32_t side, const void* info_in) {
  STAGE_PROLOGUE("sayal_slab_ipc_connect");
  if (!info_in || (side != 0 && side != 1)) return set_error(SAYAL_EINVAL, "sayal_slab_ipc_connect: bad argument");
  if (s->ipc_opened[side]) return set_error(SAYAL_EINVAL, "sayal_slab_ipc_connect: that side is already connected");
  LinkInfo info;
  std::memcpy(&info, info_in, sizeof info);
  cudaIpcMemHandle_t h;
  void *p = nullptr, *pv = nullptr;
  std::memcpy(&h, info.link_handle, sizeof h);
  CUDA_TRY(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  std::memcpy(&h, info.vel_handle, sizeof h);
  cudaError_t e = cudaIpcOpenMemHandle(&pv, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) {
    cudaIpcCloseMemHandle(p);
    return set_error(SAYAL_ECUDA, cudaGetErrorString(e));
  }
  int r = slab_link_connect_info(s, side, &info, p, pv);
  if (r != SAYAL_OK) {
    cudaIpcCloseMemHandle(p);
    cudaIpcCloseMemHandle(pv);
    return r;
  }
  s->ipc_opened[side] = p;
  s->ipc_opened_vel[side] = pv;
  invalidate_graphs(s);
  s->plan_variant = -1, s->n_plans = 0;  // linked slabs tune their plans under the chain's constraints
  return SAYAL_OK;
}

int sayal_slab_connect_local(sayal_sim* sim, int32_t side, sayal_sim* neighbour) {
  STAGE_PROLOGUE("sayal_slab_connect_local");
  if (!neighbour || (side != 0 && side != 1)) return set_error(SAYAL_EINVAL, "sayal_slab_connect_local: bad argument");
  Sim* n = S(neighbour);
  CUDA_TRY(cudaSetDevice(n->device));
  LinkInfo info;
  TRY(slab_link_export(n, &info));
  CUDA_TRY(cudaSetDevice(s->device));
  if (n->device != s->device) {
    int can = 0;
    CUDA_TRY(cudaDeviceCanAccessPeer(&can, s->device, n->device));
    if (!can) return set_error(SAYAL_EINVAL, "sayal_slab_connect_local: no peer access between the two devices");
    cudaError_t e = cudaDeviceEnablePeerAccess(n->device, 0);
    if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return set_error(SAYAL_ECUDA, cudaGetErrorString(e));
    cudaGetLastError();
  }
  TRY(slab_link_connect_info(s, side, &info, n->link_block, n->vel_block));
  invalidate_graphs(s);
  s->plan_variant = -1, s->n_plans = 0;
  return SAYAL_OK;
}

int sayal_slab_exchange(sayal_sim* sim, int32_t field_mask) {
  STAGE_PROLOGUE("sayal_slab_exchange");
  return launch_slab_exchange(s, field_mask);
}

// ---- slab edge rows ---------------------------------------------------------------------------------
// side 0 = low memory rows (towards r = 0), side 1 = high memory rows.
int sayal_slab_pack_edge(sayal_sim* sim, int32_t side, int32_t nrows, int32_t field_mask, void* dev_buf) {
  STAGE_PROLOGUE("sayal_slab_pack_edge");
  if (!dev_buf) return set_error(SAYAL_EINVAL, "sayal_slab_pack_edge: null buffer");
  int row0 = side == 0 ? s->g.own_lo : s->g.own_hi - nrows;  // the owned rows next to that side
  return launch_pack_rows(s, row0, nrows, field_mask, (float*)dev_buf, false);
}

int sayal_slab_unpack_ghost(sayal_sim* sim, int32_t side, int32_t nrows, int32_t field_mask, const void* dev_buf) {
  STAGE_PROLOGUE("sayal_slab_unpack_ghost");
  if (!dev_buf) return set_error(SAYAL_EINVAL, "sayal_slab_unpack_ghost: null buffer");
  int row0 = side == 0 ? s->g.own_lo - nrows : s->g.own_hi;  // the ghost rows beyond that side
  return launch_pack_rows(s, row0, nrows, field_mask, (float*)const_cast<void*>(dev_buf), true);
}

}  // extern "C"
