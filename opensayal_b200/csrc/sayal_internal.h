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
// Every sim owns one neighbour-writable block: 64 control words followed by the receive areas.  Control words:
//   [0], [1]   exchange flags: sequence number of the last exchange the neighbour on that side has pushed
//   [2], [3]   pass flags (projection passes that push their edge rows themselves, projection_pack.cu)
//   [4], [5]   plan signature of the neighbour's pushes (cross-check: both sides must split the projection alike)
//   [8..10]    pressure range travelling down the chain (from side 0): sequence, ordered min, ordered max
//   [12..14]   global pressure range travelling back up (from side 1)
enum : int { LW_XFLAG = 0, LW_PFLAG = 2, LW_PITER = 4, LW_RANGE_DOWN = 8, LW_RANGE_UP = 12, LW_WORDS = 64 };
// link_error values (sticky; the first one wins)
enum : int { LINK_OK = 0, LINK_TIMEOUT = 1, LINK_PLAN_MISMATCH = 2 };

#ifdef __CUDACC__
// system-scope flag primitives of the neighbour protocols (slab_exchange.cu, projection_pack.cu)
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ long long now_ns() {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
constexpr long long kSpinLimitNs = 2000000000ll;
// The sticky error word of a slab link lives in DEVICE memory (every spin iteration of every waiting warp reads it:
// as a mapped host word those reads crossed PCIe by the hundred and stretched every wait); the first error is
// mirrored once into a mapped host word whose address sits two ints behind it, where the host reads it for free.
__device__ __forceinline__ void raise_link_error(int* link_error, int code) {
  if (atomicCAS(link_error, 0, code) == 0) {
    int* host = *reinterpret_cast<int**>(link_error + 2);
    if (host) *reinterpret_cast<volatile int*>(host) = code;
  }
  __threadfence_system();
}
// Spin until *word >= target (wrap-safe).  Gives up after two seconds, or at once when the link is already broken,
// raising the sticky error word; returns false then.
__device__ __forceinline__ bool spin_until(const unsigned* word, unsigned target, int* link_error) {
  if ((int)(ld_acquire_sys(word) - target) >= 0) return true;
  const long long t0 = now_ns();
  while ((int)(ld_acquire_sys(word) - target) < 0) {
    if (now_ns() - t0 > kSpinLimitNs || *reinterpret_cast<volatile int*>(link_error) != 0) {
      raise_link_error(link_error, 1);
      return false;
    }
  }
  return true;
}
#endif

struct SlabLinkDev {
  float* peer_recv[2];    // where my edge rows land in the neighbour's block (null: no neighbour on that side)
  unsigned* peer_words[2];// the neighbour's control words
  float* my_recv[2];      // where the neighbours' rows land here (2 parities each)
  unsigned* my_words;     // my control words, written by the neighbours
  unsigned *send_seq, *recv_seq, *ticket;  // private counters, advanced by the kernels
  unsigned *range_seq, *step_seq, *push_ticket;
  int* link_error;        // device word, raised when a wait gave up (sticky: later waits return at once; see raise_link_error)
  size_t stage_elems;     // floats per receive buffer: halo * W * 3
};

// What two neighbouring slabs tell each other once (opaque to callers: SAYAL_LINK_INFO_BYTES of include/sayal.h).
struct LinkInfo {
  unsigned char link_handle[64];  // cudaIpcMemHandle_t of the link block
  unsigned char vel_handle[64];   // cudaIpcMemHandle_t of the velocity block
  int64_t stage_elems;
  int64_t vel_stride;             // bytes between the four velocity arrays
  int32_t W, pitch, local_rows, own_lo, own_hi, halo, advect_margin, parity;
  int32_t row_base, global_height, abi, reserved;
};

struct Sim {
  sayal_config cfg;
  Grid g;
  Phys ph;
  int device;
  cudaStream_t stream;
  // fp32 fields, local_rows x pitch.  u, v and their back buffers live in ONE allocation (vel_block: four arrays
  // vel_stride bytes apart, in the order u, v, u_buf, v_buf as created) so that a neighbouring slab can map all of
  // them with one IPC handle and store its edge rows straight into our ghost rows (projection_pack.cu).
  float *u, *v, *p, *smoke, *u_buf, *v_buf, *smoke_buf;
  void* vel_block;
  void* smoke_block;  // smoke, smoke_buf with guard rows around each (see create_impl)
  size_t vel_stride;
  uint8_t* flags;
  uint16_t* geo;                    // static per-cell geometry word of the tile advection (advect_tile.cu)
  int32_t *d_is_solid, *d_total_s;  // built on demand for get_field / device_ptr
  int32_t* d_range;                 // ordered-int min / max of pressure: [0], [1] of the owned rows; [2], [3] of the
                                    // whole domain (linked slabs: reduced along the chain, slab_exchange.cu)
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
  // profiling only (option "debug_events", eager steps): CUDA events at the stage boundaries of the last step
  static constexpr int kDbgEvents = 10;
  cudaEvent_t dbg_ev[kDbgEvents];
  int debug_events;
  unsigned dbg_marked;    // bit k: dbg_ev[k] was recorded in the last step
  int debug_skip;         // profiling only (option "debug_skip"): bit mask of step stages to leave out (wrong results!)
  int fuse_extrap;        // 1: the next tiled projection call ends the step's projection and also extrapolates; 2: it did
  int autotune;           // time candidate tile plans on first use
  int force_variant;      // -1 = any tile variant
  int plan_variant, plan_T;  // tile plan of the last projection (projection_pack.cu)
  static constexpr int kMaxPlans = 8;
  struct Plan { int iterations, variant, T, push, resident, resident_ok; } plans[kMaxPlans];  // one per iteration count seen
  // resident projection (projection_pack.cu, ResidentArgs): the whole projection in one cooperative launch
  int resident;            // option (default 1): consider resident plans when the tiles fit on the GPU at once
  int plan_resident;       // the plan of the last projection is a resident one
  unsigned* d_resident;    // [0] epoch the exchange tags count from, [1] ticket of finished tiles
  void* d_box;             // the four mailbox arrays of the ring exchange (8 bytes per cell each), allocated on demand
  int *h_resident_error, *d_resident_error;  // mapped host word raised by a tile whose neighbour never showed up
  int n_plans;
  char plan_log[2048];     // candidates of the last tuning (sayal_plan_log)
  // issue order of the tiles (most expensive first) per tile geometry, built on first use (projection_pack.cu)
  int order_tiles;        // option (default 1)
  static constexpr int kMaxOrders = 256;
  struct TileOrder { int variant, it, row_lo, row_hi, edge_first; int* order; int n_descs; } orders[kMaxOrders];  // edge_first -1: `order` holds a tile list (TileDesc)
  int tune_depth;         // ghost depth the plan tuner times its candidates with (-1: whole-array passes)
  int split_tiles;        // option (default 1): whole-domain passes run over an explicit tile list (projection_pack.cu)
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
  void* link_block;         // neighbour-writable block (control words + receive areas), IPC-exportable
  unsigned* link_counters;  // private
  SlabLinkDev link;
  int* h_link_error;        // mapped host mirror of link.link_error, written by the kernel that raises it: read after a sync at no cost
  void* ipc_opened[2];      // peer link blocks opened with cudaIpcOpenMemHandle (closed on destroy)
  void* ipc_opened_vel[2];  // peer velocity blocks, likewise
  // in-pass push (projection_pack.cu): the neighbours' four velocity arrays as addressable from this device, in
  // their creation order, and where my edge rows land in them
  float* peer_vel[2][4];
  int peer_ghost_row0[2];   // local row (in the neighbour's arrays) of the first row I push to that side
  int slab_push;            // option (default 1): projection passes push their edge rows themselves, one thin
                            // exchange per pass instead of a deep halo recomputed through the whole step
  int push_active;          // set by step_impl for the projection call it is about to make
  cudaEvent_t ev_range;     // the chain reduction of the pressure range runs on aux_stream
  // host-released gate (sayal_stream_hold / sayal_stream_release): a one-thread kernel spinning on a mapped word
  unsigned* h_gate;
  unsigned* d_gate;
  unsigned gate_ticket;     // value the next release writes
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

bool is_linked(const Sim* s);  // sayal_api.cu: the slab has at least one neighbour
bool push_mode(const Sim* s);  // projection passes push their edge rows themselves (projection_pack.cu)

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
size_t slab_link_stage_elems(const Sim* s);
int launch_slab_exchange(Sim* s, int field_mask);
int launch_slab_exchange_on(Sim* s, int field_mask, cudaStream_t stream);
int launch_slab_range_reduce(Sim* s, cudaStream_t stream);  // d_range[0..1] of every slab -> d_range[2..3] everywhere
int launch_slab_push_wait(Sim* s, int passes, int signature);              // the neighbours' pushes of a step's last pass have landed
struct LinkInfo;                                             // what neighbours trade once (sayal_slab_ipc_export)
int slab_link_export(Sim* s, LinkInfo* out);
int slab_link_connect_info(Sim* s, int side, const LinkInfo* info, void* peer_block, void* peer_vel);

// ---- visual.cu -------------------------------------------------------------------------------------
int launch_diffusion(Sim* s, int iterations, float d_t);
int launch_render_pixels(Sim* s, uint32_t* d_pixels);  // owned rows, pitch W
int launch_path_lines(Sim* s, int dist, int len, float d_t, int nx, int ny, int32_t* d_xs, int32_t* d_ys);
int launch_arrows(Sim* s, const sayal_visual* v, int nx, int ny, sayal_arrow* d_out);

// ---- projection_pack.cu --------------------------------------------------------------------------
int launch_projection_tiled(Sim* s, int iterations, float d_t);
int tiled_max_temporal_block();
int tiled_debug_pass_plans(int pitch, int local_rows, int own_lo, int own_hi, int rows_per_warp, int T, int iterations,
                           int ghost_depth, int32_t* out, int capacity);  // host only; 15 int32 per pass, see sayal.h
int tiled_debug_tile_list(Sim* s, int it, int32_t* out, int capacity);  // the plan's explicit tile list (sayal_debug_tile_list)
int tiled_push_temporal_block(int iterations, int halo);  // push mode: the T every rank of a chain uses
int preload_basic();   // each file's kernels, loaded at sayal_create (see tiled_preload)
int preload_advect();
int preload_slab();
int preload_visual();
int tiled_preload();  // load every kernel variant now (never lazily in the middle of a linked step)
int tiled_prepare(Sim* s, int iterations);  // choose the tile plan (may time candidates; not capturable)
int tiled_prepare_windows(Sim* s, int iterations, int ghost_depth);  // + the issue orders of a linked slab's row windows


}  // namespace sayal

#endif
