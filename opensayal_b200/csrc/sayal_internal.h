// sayal_internal.h — shared declarations of the B200-native step path (not part of the public ABI).
#ifndef SAYAL_INTERNAL_H
#define SAYAL_INTERNAL_H

#include <cstdint>
#include <cstddef>

#include "sayal.h"

#ifdef __CUDACC__
#include <cuda_runtime.h>
#else
typedef struct CUstream_st* cudaStream_t;
typedef struct CUgraphExec_st* cudaGraphExec_t;
typedef struct CUevent_st* cudaEvent_t;
#endif

namespace sayal {

// thread-local last-error slot behind sayal_last_error(); returns `code`
int set_error(int code, const char* msg);

// Cell-flag byte (1 B/cell instead of the reference's two int32 arrays, fluid.cuh:71-72).
//   bit 0..3  the four face updates of apply_projection_at (fluid.cu:247-261) that are enabled for this
//             cell: L = left neighbour not solid, R, B (j-1), T (j+1); all zero unless the cell is active
//   bit 4..6  total_s (fluid.cu:127-142) if the cell is active (interior and not solid), else 0
//   bit 7     is_solid (fluid.cu:113-124)
// A fully open interior fluid cell is 0x4F.
enum : uint8_t { FL_L = 1, FL_R = 2, FL_B = 4, FL_T = 8, FL_SOLID = 0x80, FL_OPEN = 0x4F };

// Geometry of the arrays one sim holds.  Memory row r = H-1-j (fluid.cu:163-165); a y-slab holds
// global rows [row_base, row_base + local_rows), of which [own_lo, own_hi) (local indices) are owned.
struct Grid {
  int W, H;        // global cells
  int pitch;       // elements per row in every array (multiple of 4)
  int row_base;    // global memory row of local row 0
  int local_rows;  // rows held (owned + ghost)
  int own_lo, own_hi;
  int valid_lo, valid_hi;  // local rows whose data is exact right now (slabs: ghost rows go stale between exchanges);
                           // the advection samplers count a gather that leaves this range as halo overflow
  int h;           // Fluid::cell_size, an int (fluid.cuh:46)
};

// Scalars of Fluid (fluid.cuh:40-59) that kernels need, passed by value (the reference re-loads
// them from a device copy of the object in every thread).
struct Phys {
  float o, density, g, drag_coeff;
  float wt_speed, wt_smoke;
  int wt_height, wt_smoke_length, wt_smoke_count, wt_smoke_height;
  int enable_pressure, enable_smoke, enable_decay;
  float decay_rate;
  // mask formula inputs (fluid.cu:113-124)
  int enable_drain, obstacle_enable, obstacle_cx, obstacle_cy, wt_pipe_length, wt_pipe_height;
  float obstacle_radius;
};

// Host-evaluated parameters of apply_external_forces_at (fluid.cu:308-349) for one step.
struct ForceArgs {
  int band_lo, band_hi;    // H/2 - ph/2 .. H/2 + ph/2
  int smoke_lo, smoke_hi;  // count == 1 band
  int period, anchor;      // count != 1: (anchor - j) % period < smoke_height
  float damping;           // expf(-drag*dt), evaluated once on the host (fluid.cu:332)
  int src_active, src_x, src_y;
  float src_smoke, src_velocity;
  float d_t;
};

// Device-side view of a slab's links to its two neighbours (slab_exchange.cu); side 0 = low memory rows.
struct SlabLinkDev {
  float* peer_recv[2];    // where my edge rows land in the neighbour's block (null: no neighbour on that side)
  unsigned* peer_flag[2]; // the neighbour's flag word for my pushes
  float* my_recv[2];      // where the neighbours' rows land here (2 parities each)
  unsigned* my_flag;      // [2], written by the neighbours
  unsigned *send_seq, *recv_seq, *ticket;  // private counters, advanced by the kernels
  int* link_error;        // raised when a wait gave up
  size_t stage_elems;     // floats per receive buffer: halo * W * 3
};

struct Sim {
  sayal_config cfg;
  Grid g;
  Phys ph;
  int device;
  cudaStream_t stream;
  // fp32 fields, local_rows x pitch
  float *u, *v, *p, *smoke, *u_buf, *v_buf, *smoke_buf;
  uint8_t* flags;
  uint16_t* geo;                    // static per-cell geometry word of the tile advection (advect_tile.cu)
  int32_t *d_is_solid, *d_total_s;  // built on demand for get_field / device_ptr
  int32_t* d_range;                 // ordered-int min / max of pressure
  int32_t* d_overflow;              // count of back-traces that left the local rows (slab runs)
  long long* d_timeline;            // profiling only: per-CTA phase timestamps of the last projection pass
  size_t timeline_cap;              // capacity in int64
  int timeline_tiles;
  float min_p, max_p;
  bool range_valid;
  // options
  int projection_kernel;  // 0 plain half-sweeps, 1 register tile (packed pairs + profile multipliers)
  int temporal_block;     // iterations per pass of the tiled kernel
  int use_graph;
  int advect_kernel;      // 0 plain per-cell kernels, 1 shared-memory tiles, 2 direct with geometry words (cell_size 1)
  int use_pdl;            // programmatic dependent launch between projection passes
  int fuse_forces;        // option: fold the forces into the load of the step's first projection pass (default 1)
  int fuse_extrapolation; // option: fold the boundary extrapolation into the store of the step's last projection pass (default 1)
  int fuse_pending;       // the step's forces have not been applied yet: the next tiled pass applies them
  ForceArgs fuse_args;
  int shrink_window;      // option: linked slabs sweep only the ghost rows that are still exact (default 1)
  int proj_depth;         // linked slabs: ghost rows still exact when the next projection call starts (-1: all rows)
  int debug_skip;         // profiling only (option "debug_skip"): bit mask of step stages to leave out (wrong results!)
  int fuse_extrap;        // 1: the next tiled projection call ends the step's projection and also extrapolates; 2: it did
  int autotune;           // time candidate tile plans on first use
  int force_variant;      // -1 = any tile variant
  int plan_variant, plan_T;  // tile plan of the last projection (projection_pack.cu)
  static constexpr int kMaxPlans = 8;
  struct Plan { int iterations, variant, T; } plans[kMaxPlans];  // one per iteration count seen
  int n_plans;
  // issue order of the tiles (most expensive first) per tile geometry, built on first use (projection_pack.cu)
  int order_tiles;        // option (default 1)
  static constexpr int kMaxOrders = 64;
  struct TileOrder { int variant, it, row_lo, row_hi; int* order; } orders[kMaxOrders];
  int n_orders;
  // CUDA graph cache for sayal_run: one single-step graph per starting buffer parity, with the
  // pointer assignment the step leaves behind (the step swaps front and back buffers)
  cudaGraphExec_t graph[4];   // indexed by `parity`
  int64_t graph_launches[4];  // kernels one replay of graph[k] stands for
  struct PtrState { float *u, *v, *u_buf, *v_buf, *smoke, *smoke_buf; int parity; } graph_post[4];
  float graph_dt;
  int parity;  // bit 0: (u,v) and their back buffers are swapped; bit 1: smoke and its back buffer are
  int64_t launches;
  // y-slab links (slab_exchange.cu)
  int slab_halo;            // ghost rows per interior side as requested at creation (0: whole domain)
  void* link_block;         // neighbour-writable block (flags + receive areas), IPC-exportable
  unsigned* link_counters;  // private
  SlabLinkDev link;
  void* ipc_opened[2];      // peer blocks opened with cudaIpcOpenMemHandle (closed on destroy)
  cudaStream_t aux_stream;  // exchange kernels run here, concurrently with interior compute
  cudaEvent_t ev_fork, ev_join;
  int overlap_exchange;     // option: overlap exchanges with interior compute (default 1)
  int advect_margin;        // ghost depth the advection may gather from: ceil(max|velocity| d_t) + 2 (default 16)
  // frame read-back ring (visual.cu / sayal_frame_*): three device frames, three pinned host frames, a copy
  // stream; at most two frames outstanding, so an acquired frame survives one further submit
  static constexpr int kFrames = 3;
  int64_t steps_done;       // updates enqueued since creation
  uint32_t* d_frame[kFrames];
  uint32_t* h_frame[kFrames];  // cudaHostAlloc
  cudaEvent_t ev_rendered[kFrames], ev_copied[kFrames];
  cudaStream_t copy_stream;
  int64_t frame_step[kFrames];
  int frame_head, frame_pending;  // next slot to submit into, frames submitted and not yet acquired
  bool frame_used[kFrames];       // ev_copied[k] has been recorded at least once
};

// ---- kernels_basic.cu ---------------------------------------------------------------------------
int launch_build_flags(Sim* s);
int launch_export_masks(Sim* s);
int launch_forces(Sim* s, const sayal_source* src, float d_t, bool may_fuse = false);
int launch_zero_pressure(Sim* s);
int launch_projection_plain(Sim* s, int iterations, float d_t);
int launch_pressure_range(Sim* s);
int launch_extrapolation(Sim* s);
int launch_advect(Sim* s, float d_t, bool velocity, bool smoke);
int launch_sample_velocity(Sim* s, int n, const float* d_xs, const float* d_ys, float* d_ou, float* d_ov);
int launch_pack_rows(Sim* s, int local_row0, int nrows, int field_mask, float* dev_buf, bool unpack);

// ---- advect_tile.cu -------------------------------------------------------------------------------
int launch_build_geo(Sim* s);
int launch_advect_tile(Sim* s, float d_t, bool velocity, bool smoke);
int launch_advect_geo(Sim* s, float d_t, bool velocity, bool smoke);
int launch_advect_geo_rows(Sim* s, float d_t, bool smoke, int row_lo, int row_hi);  // local rows [row_lo, row_hi)

// ---- slab_exchange.cu -----------------------------------------------------------------------------
int slab_link_alloc(Sim* s);
int slab_link_connect(Sim* s, int side, void* peer_block, size_t peer_stage_elems);
size_t slab_link_stage_elems(const Sim* s);
int launch_slab_exchange(Sim* s, int field_mask);
int launch_slab_exchange_on(Sim* s, int field_mask, cudaStream_t stream);

// ---- visual.cu -------------------------------------------------------------------------------------
int launch_diffusion(Sim* s, int iterations, float d_t);
int launch_render_pixels(Sim* s, uint32_t* d_pixels);  // owned rows, pitch W
int launch_path_lines(Sim* s, int dist, int len, float d_t, int nx, int ny, int32_t* d_xs, int32_t* d_ys);
int launch_arrows(Sim* s, const sayal_visual* v, int nx, int ny, sayal_arrow* d_out);

// ---- projection_pack.cu --------------------------------------------------------------------------
int launch_projection_tiled(Sim* s, int iterations, float d_t);
int tiled_max_temporal_block();
int tiled_debug_pass_plans(int pitch, int local_rows, int own_lo, int own_hi, int rows_per_warp, int T, int iterations,
                           int ghost_depth, int32_t* out, int capacity);  // host only; 11 int32 per pass, see sayal.h
int tiled_preload();  // load every kernel variant now (never lazily in the middle of a linked step)
int tiled_prepare(Sim* s, int iterations);  // choose the tile plan (may time candidates; not capturable)
int tiled_prepare_windows(Sim* s, int iterations, int ghost_depth);  // + the issue orders of a linked slab's row windows


}  // namespace sayal

#endif
