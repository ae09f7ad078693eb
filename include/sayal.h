/*
 * sayal.h — C ABI of the B200-native OpenSayal step path.
 *
 * The reference (gopmur/OpenSayal) has no FFI: its de-facto boundary is the public surface of
 * `class Fluid` (inc/fluid.cuh:15-116) as used by main (src/main.cu:49,97) and by the renderer
 * (src/graphics_handler.cu:269-283).  Every entry point below names the reference interface it
 * replaces.  Plain pointers and sizes only; no C++/torch types.  All functions return 0 on success
 * or a negative SAYAL_E* code (the reference never checks an error and calls exit(); we never do).
 *
 * Field layout handed across this boundary is the reference's: row-major W x H, flipped y,
 *     index(i, j) = (H - 1 - j) * W + i          (src/fluid.cu:163-165)
 * fp32 for U/V/P/SMOKE, int32 for IS_SOLID/TOTAL_S.
 *
 * Threading contract: one host thread per sayal_sim; distinct sims are independent.
 */
#ifndef SAYAL_H
#define SAYAL_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SAYAL_ABI_VERSION 3

/* error codes */
#define SAYAL_OK 0
#define SAYAL_EINVAL (-1)   /* bad argument / unsupported configuration */
#define SAYAL_ECUDA (-2)    /* a CUDA call failed; see sayal_last_error() */
#define SAYAL_EIO (-3)      /* config file missing / unreadable */
#define SAYAL_EPARSE (-4)   /* config file is not valid JSON / wrong value type */
#define SAYAL_ENOMEM (-5)
#define SAYAL_EBUSY (-6)    /* frame ring full: acquire a frame first */
#define SAYAL_ELINK (-7)    /* a neighbouring slab did not answer (or runs another plan): fields are not valid */

/* fields, for sayal_get_field / sayal_set_field / sayal_device_ptr */
enum sayal_field {
  SAYAL_U = 0,        /* Fluid::d_vel_x   (fluid.cuh:65) */
  SAYAL_V = 1,        /* Fluid::d_vel_y   (fluid.cuh:66) */
  SAYAL_P = 2,        /* Fluid::d_pressure (fluid.cuh:64) */
  SAYAL_SMOKE = 3,    /* Fluid::d_smoke   (fluid.cuh:67) */
  SAYAL_IS_SOLID = 4, /* Fluid::d_is_solid (fluid.cuh:71), int32 0/1 */
  SAYAL_TOTAL_S = 5,  /* Fluid::d_total_s  (fluid.cuh:72), int32 0..4 */
  SAYAL_FIELD_COUNT = 6
};

/*
 * The simulation subset of the reference's `Config` (inc/config_parser.hpp:9-114) — exactly the
 * members the Fluid constructor and init_device_memory read (src/fluid.cu:41-61, 111-124), with the
 * defaults of src/config_parser.cpp:21-119.  JSON key in the trailing comment.
 */
typedef struct sayal_config {
  int32_t width;            /* sim.width             (1920) */
  int32_t height;           /* sim.height            (1080) */
  float cell_size;          /* sim.cell_size         (1.0; Fluid stores it as int, fluid.cuh:46) */
  int32_t enable_drain;     /* sim.enable_drain      (true) */
  int32_t enable_pressure;  /* sim.enable_pressure   (false) */
  int32_t enable_smoke;     /* sim.enable_smoke      (true) */
  int32_t enable_interactive; /* sim.enable_interactive (false) — carried, unused headless */
  int32_t proj_n;           /* sim.projection.n      (50) */
  float proj_o;             /* sim.projection.o      (1.9) */
  int32_t wt_pipe_height;   /* sim.wind_tunnel.pipe_height  (height/4) */
  int32_t wt_pipe_length;   /* hard-coded 0 in config_parser.cpp:64 */
  int32_t wt_smoke_length;  /* sim.wind_tunnel.smoke_length (1) */
  int32_t wt_smoke_height;  /* sim.wind_tunnel.smoke_height (height/4) */
  int32_t wt_smoke_count;   /* sim.wind_tunnel.smoke_count  (1) */
  float wt_speed;           /* sim.wind_tunnel.speed (0) */
  float wt_smoke;           /* sim.wind_tunnel.smoke (1) */
  float g;                  /* sim.physics.g         (0) */
  float d_t;                /* sim.time.d_t          (0.05) */
  int32_t enable_real_time; /* sim.time.enable_real_time (false) — carried for the caller */
  float real_time_multiplier; /* sim.time.real_time_multiplier (1) */
  int32_t smoke_enable_decay; /* sim.smoke.enable_decay (false) */
  float smoke_decay_rate;   /* sim.smoke.decay_rate  (0.05) */
  int32_t obstacle_enable;  /* sim.obstacle.enable   (true) */
  int32_t obstacle_center_x; /* sim.obstacle.center_x (width/2) */
  int32_t obstacle_center_y; /* sim.obstacle.center_y (height/2) */
  float obstacle_radius;    /* sim.obstacle.radius   (min(width,height)/30) */
  float density;            /* fluid.density         (1) */
  float drag_coeff;         /* fluid.drag_coeff      (0) */
  float viscosity;          /* fluid.viscosity       (0.001 in the reference) */
  int32_t block_size_x;     /* thread.cuda.block_size_x (64) — carried for the reference shim only */
  int32_t block_size_y;     /* thread.cuda.block_size_y (1) */
} sayal_config;

/* struct Source (inc/fluid.cuh:8-13): the mouse impulse passed by value to Fluid::update. */
typedef struct sayal_source {
  int32_t active;
  float smoke;
  float velocity;
  int32_t x; /* position.x, cell units */
  int32_t y; /* position.y, cell units, j (bottom-up) */
} sayal_source;

/*
 * Optional y-slab description for multi-GPU runs (no reference equivalent).  A slab owns
 * the memory rows [row0, row0+rows) of a global_height-row domain and carries `halo` ghost rows on each
 * interior side.  All-zero / NULL means "whole domain on this GPU".
 */
typedef struct sayal_slab {
  int32_t global_height; /* H of the whole domain */
  int32_t row0;          /* first owned memory row (memory row r = H-1-j) */
  int32_t rows;          /* owned rows */
  int32_t halo;          /* ghost rows kept above and below (clipped at the domain edge) */
} sayal_slab;

typedef struct sayal_sim sayal_sim;

/* ---- configuration ------------------------------------------------------------------------- */
/* Fill `cfg` with the reference defaults for a width x height grid (config_parser.cpp:21-119). */
int sayal_config_defaults(int32_t width, int32_t height, sayal_config* cfg);
/* ConfigParser::parse() (config_parser.cpp:15-185) incl. get_or's dotted/nested lookup
 * (config_parser.hpp:131-151).  Missing keys take defaults; a missing file is SAYAL_EIO and invalid
 * JSON SAYAL_EPARSE instead of the reference's uncaught exception. */
int sayal_config_load(const char* json_path, sayal_config* cfg);
int sayal_config_parse(const char* json_text, size_t len, sayal_config* cfg);

/* ---- lifetime ------------------------------------------------------------------------------ */
/* Fluid::Fluid(Config) (fluid.cu:41-71): allocate device state on `device`, build is_solid/total_s
 * (fluid.cu:99-161), zero u, v, smoke (and p, which the reference leaves uninitialised). */
int sayal_create(const sayal_config* cfg, int32_t device, sayal_sim** out);
/* Same, for one y-slab of a larger domain. */
int sayal_create_slab(const sayal_config* cfg, int32_t device, const sayal_slab* slab, sayal_sim** out);
/* Fluid::~Fluid() (fluid.cu:73-84). */
void sayal_destroy(sayal_sim* sim);

/* ---- stepping ------------------------------------------------------------------------------ */
/* Fluid::update(Source, d_t) (fluid.cu:770-795) without the trailing cudaDeviceSynchronize:
 * enqueues one step on the sim's stream and returns.  `src` may be NULL (inactive source). */
int sayal_step(sayal_sim* sim, const sayal_source* src, float d_t);
/* Headless batch: `steps` updates with an inactive source (CUDA-graph replay). */
int sayal_run(sayal_sim* sim, int32_t steps, float d_t);
/* The cudaDeviceSynchronize() of fluid.cu:794, restricted to this sim's stream. */
int sayal_sync(sayal_sim* sim);

/* ---- state access -------------------------------------------------------------------------- */
/* Copy a whole field device->host / host->device in the reference layout (W*H elements of 4 bytes).
 * For a slab sim the buffer holds the owned rows only (rows*W elements). Synchronous. */
int sayal_get_field(sayal_sim* sim, int32_t field, void* host_dst);
int sayal_set_field(sayal_sim* sim, int32_t field, const void* host_src);
/* The same for `n` fields at once (what main.cu would do around a batch of updates): the copies are enqueued on the
 * sim's stream back to back.  sayal_get_fields waits once, after the last copy.  sayal_set_fields does not wait at
 * all when the sources are pinned host memory — the caller keeps them untouched until the next synchronous call
 * (sayal_sync, sayal_get_field(s), ...); pageable sources are staged by the runtime before the call returns. */
int sayal_get_fields(sayal_sim* sim, int32_t n, const int32_t* fields, void* const* host_dsts);
int sayal_set_fields(sayal_sim* sim, int32_t n, const int32_t* fields, const void* const* host_srcs);
/* Device-to-device forms for a caller that already lives on the GPU (a renderer, an interop tensor): `dev_*` holds
 * the owned rows in the reference layout (row pitch W elements) on the sim's device.  Ordered on the sim's stream,
 * asynchronous: the caller orders its own streams against sayal_stream() / sayal_sync(). */
int sayal_get_field_device(sayal_sim* sim, int32_t field, void* dev_dst);
int sayal_set_field_device(sayal_sim* sim, int32_t field, const void* dev_src);
/* Zero-copy view for a renderer (replaces dereferencing Fluid::d_* on device,
 * graphics_handler.cu:269-283): device pointer to memory row 0 of the local array, its pitch in
 * elements, the first global memory row it holds and the number of rows held (incl. ghost rows). */
int sayal_device_ptr(sayal_sim* sim, int32_t field, void** dev_ptr, int64_t* pitch_elems,
                     int32_t* first_row, int32_t* n_rows);
/* Fluid::min_pressure / max_pressure (fluid.cuh:61-62, fluid.cu:778-787), of the last step.
 * Synchronises the stream. Only meaningful when enable_pressure.  Linked slabs return the range of the WHOLE
 * domain (reduced along the chain of slabs inside the step), the one pair the reference's frame is coloured with
 * (graphics_handler.cu:288-289); sayal_render_pixels / sayal_frame_submit use the same pair. */
int sayal_pressure_range(sayal_sim* sim, float* min_p, float* max_p);
/* Fluid::get_general_velocity(x, y) (fluid.cu:541-545) at `n` host-supplied points. */
int sayal_sample_velocity(sayal_sim* sim, int32_t n, const float* xs, const float* ys, float* out_u,
                          float* out_v);

/* ---- staged access (tests, multi-GPU drivers) ----------------------------------------------- */
/* Individual stages of Fluid::update on the sim's stream, in the reference's order
 * (fluid.cu:771-793).  sayal_step == forces, [zero_pressure], projection(n), [pressure range],
 * extrapolation, velocity advection, [smoke advection + decay]. */
int sayal_stage_forces(sayal_sim* sim, const sayal_source* src, float d_t);
int sayal_stage_zero_pressure(sayal_sim* sim);
int sayal_stage_projection(sayal_sim* sim, int32_t iterations, float d_t);
int sayal_stage_extrapolation(sayal_sim* sim);
int sayal_stage_advect_velocity(sayal_sim* sim, float d_t);
int sayal_stage_advect_smoke(sayal_sim* sim, float d_t);

/* ---- diffusion (fluid.cu:167-190) ------------------------------------------------------------ */
/* `iterations` sweeps of  u = (u + a (uW + uE + uS + uN)) / (1 + 4a),  a = viscosity d_t / cell_size^2, over the
 * interior cells, u only — the reference's apply_diffusion, which Fluid::update runs n times when
 * fluid.viscosity != 0 (fluid.cu:775-777).  The reference sweeps in place in whatever order the blocks happen
 * to run (a data race); here the order is fixed: cells with (i+j) even, then (i+j) odd.  The two differ by
 * O(a^2) per sweep (a = 5e-5 with the shipped defaults), far inside the 1e-5 parity tolerance.
 * sayal_step / sayal_run run this stage between the forces and the projection when viscosity != 0. */
int sayal_stage_diffusion(sayal_sim* sim, int32_t iterations, float d_t);

/* ---- the renderer's consumers of the state (graphics_handler.cu) ------------------------------------------
 * Everything GraphicsHandler::update (graphics_handler.cu:463-478) computes from the Fluid, without SDL: the
 * RGBA8888 frame, the velocity arrows and the path lines.  A headless caller (sayal_run --frames) or a patched
 * main.cu draws / stores them; stepping never waits for the read-back unless the caller asks for the result. */

/* The members of `Config` GraphicsHandler reads (graphics_handler.cu:99-121; defaults config_parser.cpp:40,120-181). */
typedef struct sayal_visual {
  int32_t cell_pixel_size;         /* sim.cell_pixel_size          (1)  */
  int32_t arrows_enable;           /* visual.arrows.enable         (false) */
  int32_t arrows_distance;         /* visual.arrows.distance       (20) */
  float arrows_length_multiplier;  /* visual.arrows.length_multiplier (0.1) */
  float arrows_disable_threshold;  /* visual.arrows.disable_threshold (0) */
  int32_t arrows_head_length;      /* visual.arrows.head_length    (5)  */
  int32_t path_line_enable;        /* visual.path_line.enable      (false) */
  int32_t path_line_length;        /* visual.path_line.length      (20) */
  int32_t path_line_distance;      /* visual.path_line.distance    (20) */
  int32_t arrows_color[4];         /* visual.arrows.color.{r,g,b,a}    (0,0,0,255) — carried for the caller */
  int32_t path_line_color[4];      /* visual.path_line.color.{r,g,b,a} (0,0,0,255) */
} sayal_visual;
int sayal_visual_defaults(sayal_visual* v);
int sayal_visual_load(const char* json_path, sayal_visual* v);
int sayal_visual_parse(const char* json_text, size_t len, sayal_visual* v);

/* struct ArrowData (graphics_handler.cuh:17-27), pixel coordinates; `valid` 0 for solid cells and for
 * arrows shorter than the threshold (the reference leaves those entries untouched / stale). */
typedef struct sayal_arrow {
  int32_t start_x, start_y, end_x, end_y;
  int32_t right_head_end_x, right_head_end_y, left_head_end_x, left_head_end_y;
  int32_t valid;
} sayal_arrow;

/* update_fluid_pixels_kernel + the copy to the host (graphics_handler.cu:258-302): one uint32 per cell,
 * r<<24 | g<<16 | b<<8 | a, pixel (x = i, y = H-1-j) at y*W + x — the layout of the fields.  Solid cells are
 * (80,80,80); smoke only: (255, c, c) with c = 255 - uint8(smoke*255); with pressure: HSV hue from the pressure
 * normalised by the last step's min / max (read on the device, no host round trip), value from smoke.
 * Synchronous; a slab sim writes its owned rows. */
int sayal_render_pixels(sayal_sim* sim, uint32_t* host_dst);
/* The same frame without stalling the step stream: submit renders the state as of the work enqueued so far into
 * one of three device buffers and copies it to pinned host memory on a second stream; stepping may continue at
 * once.  acquire waits for the OLDEST submitted frame only and returns the library-owned pinned buffer plus the
 * number of updates it shows; the buffer stays valid until the second submit after the acquire.  At most two
 * frames may be outstanding: a third submit returns SAYAL_EBUSY.  Cells the reference leaves untouched (neither
 * smoke nor pressure enabled) read 0. */
int sayal_frame_submit(sayal_sim* sim);
int sayal_frame_acquire(sayal_sim* sim, const uint32_t** pixels, int64_t* step_index);
/* update_center_velocity_arrow (graphics_handler.cu:304-356) + make_arrow_data (:168-200): one arrow per
 * arrows_distance cells, (W / distance) x (H / distance) entries, entry (a, b) — cell (a*distance, b*distance) —
 * at (n_y - 1 - b) * n_x + a (indx_arrow_data, :12-14).  host_dst holds `capacity` entries. */
int sayal_arrows(sayal_sim* sim, const sayal_visual* v, sayal_arrow* host_dst, int32_t capacity, int32_t* n_x,
                 int32_t* n_y);
/* update_traces (graphics_handler.cu:358-421) over Fluid::trace (fluid.cu:16-36): from the centre of every
 * path_line_distance-th cell, `length` points of forward Euler through get_general_velocity with step d_t, as
 * rounded pixel coordinates (x, H-1-y).  Line (a, b) occupies [((n_y-1-b) * n_x + a) * length, +length) of host_x /
 * host_y (indx_traces, :8-11); lines starting in a solid cell are all -1. */
int sayal_path_lines(sayal_sim* sim, const sayal_visual* v, float d_t, int32_t* host_x, int32_t* host_y,
                     int32_t capacity, int32_t* n_x, int32_t* n_y);

/* Tuning / introspection.  set: "projection_kernel" (0 = plain half-sweeps, 1 = register-tile temporally
 * blocked), "temporal_block" (iterations per pass, 0 = choose), "tile_rows_per_warp" (0 = choose, 8/10/12),
 * "autotune" (time candidate tile plans on first use), "use_graph", "use_pdl", "fuse_forces" / "fuse_extrapolation"
 * (fold those stages into the first / last projection pass), "order_tiles" (issue expensive tiles first),
 * "split_tiles" (whole-domain passes run over an explicit tile list: obstacle tiles cut in two, sorted by cost),
 * "resident" (0 = never, 1 = time the resident plans — the whole projection in one cooperative launch — with the
 * others, 2 = resident plans only), "advect_kernel", "advect_margin", "overlap_exchange", "slab_push" (linked slabs:
 * passes push their own edge rows); "debug_timeline" / "debug_events" / "debug_skip" are profiling aids (the last
 * leaves stages out and does change results).  get: the same plus "plan_temporal_block", "plan_rows_per_warp",
 * "plan_resident", "push_mode", "halo_overflow", "link_error", "pitch", "local_rows", "own_lo", "own_hi".  No other
 * option changes results. */
int sayal_set_option(sayal_sim* sim, const char* key, int64_t value);
int sayal_get_option(sayal_sim* sim, const char* key, int64_t* value);
/* The candidates the tile-plan tuner timed for the last projection it planned, as text (one "rows T model ms" line
 * each, the chosen one marked): copies at most `capacity` bytes incl. the terminator; returns the full length. */
int sayal_plan_log(sayal_sim* sim, char* buf, int32_t capacity);
/* Profiling only (option "debug_timeline" = 1): per-CTA timestamps of the last projection pass, 5 int64 per tile
 * {entry, tile loaded, sweeps done, stores issued (globaltimer ns), SM id}. */
int sayal_debug_timeline(sayal_sim* sim, int64_t* host_dst, int32_t max_tiles, int32_t* n_tiles);
/* Diagnostics: the explicit tile list the current plan's whole-array passes of `iterations_per_pass` iterations run
 * over (option "split_tiles": obstacle tiles cut in two, last tile row aligned to the bottom wall, sorted by cost), 5
 * int32 per tile: first column, first row held, end of the rows held, first row written, end of the rows written; the
 * tile writes columns [X0 == 0 ? 0 : X0 + halo_x, X0 + 128 >= pitch ? pitch : X0 + 128 - halo_x).  *n_tiles = 0 when no
 * such list has been built (plan not chosen yet, option off, or the pass uses the regular grid). */
int sayal_debug_tile_list(sayal_sim* sim, int32_t iterations_per_pass, int32_t* out, int32_t capacity, int32_t* n_tiles);
/* Profiling only (option "debug_events" = 1, eager sayal_step): milliseconds from the start of the last step to its
 * stage boundaries — [1] projection done, [2] velocity advected, [3] / [4] edge rows of the velocity / smoke advected
 * (linked slabs), [5] end-of-step exchange done (aux stream), [6] interior rows advected, [7] step done; -1 = not
 * reached in that step. */
int sayal_debug_stage_times(sayal_sim* sim, float* ms_out, int32_t capacity);
/* Diagnostics of a slab link: the 64 neighbour-written control words of this sim followed by its 32 private counters
 * (layout: csrc/sayal_internal.h), copied after the sim's stream has drained. */
int sayal_debug_link_words(sayal_sim* sim, uint32_t* host_dst /* 96 words */);
/* Measurement aid: hold the sim's stream for `microseconds` (<= 1e6) with a one-thread spin kernel, so that a whole
 * timed region can be enqueued before the device starts on it (host launch jitter then cannot drain the queue). */
int sayal_stream_delay(sayal_sim* sim, int64_t microseconds);
/* The same with the host deciding when: sayal_stream_hold enqueues a one-thread kernel that spins on a word of
 * mapped host memory; whatever is enqueued behind it starts when sayal_stream_release is called (or after 20 s).
 * Multi-GPU runs enqueue their whole region on every rank, meet at a host barrier and release together. */
int sayal_stream_hold(sayal_sim* sim);
int sayal_stream_release(sayal_sim* sim);
/* Host-only introspection (no CUDA call, no sim): the passes the tiled projection takes for `iterations` iterations
 * with temporal block T on an array of `pitch` x `local_rows` cells whose rows [own_lo, own_hi) are owned; ghost_depth
 * < 0: every pass sweeps all rows, else pass k sweeps the owned rows +- (ghost_depth - 2 * iterations done before it).
 * ghost_depth <= -2: push mode with halo = -ghost_depth (pass k sweeps the owned rows +- 2 x its iterations on every
 * side where the array holds ghost rows, and writes the owned rows only).
 * Per pass 15 int32: iterations, row_lo, row_hi, halo_x, halo_y, stride_x, stride_y, tiles_x, tiles_y, tile_w, tile_h,
 * write_lo, write_hi, pushing tiles towards side 0, towards side 1.
 * Tile (a, b) covers columns [a stride_x, +tile_w) and rows [row_lo + b stride_y, +tile_h) and writes the part that
 * is at least a halo away from every edge that has a neighbouring tile, clipped to [write_lo, write_hi). */
int sayal_debug_pass_plans(int32_t pitch, int32_t local_rows, int32_t own_lo, int32_t own_hi, int32_t rows_per_warp,
                           int32_t temporal_block, int32_t iterations, int32_t ghost_depth, int32_t* out,
                           int32_t capacity, int32_t* n_passes);
/* Number of kernels this library has launched on behalf of `sim` since creation. */
int64_t sayal_launch_count(sayal_sim* sim);
/* CUDA stream of the sim (cudaStream_t as void*), for event timing by the caller. */
void* sayal_stream(sayal_sim* sim);

/* Slab runs: rows [row0-halo, row0) and [row0+rows, row0+rows+halo) are ghost rows.  These copy ghost /
 * edge rows between the sim's device arrays and caller-provided DEVICE buffers (packed U then V then
 * SMOKE, `nrows` rows of W floats each) on the sim's stream — the caller moves the buffers between GPUs
 * (NCCL send/recv or peer copy).  side: 0 = low-row side (top of the picture), 1 = high-row side. */
int sayal_slab_pack_edge(sayal_sim* sim, int32_t side, int32_t nrows, int32_t field_mask, void* dev_buf);
int sayal_slab_unpack_ghost(sayal_sim* sim, int32_t side, int32_t nrows, int32_t field_mask,
                            const void* dev_buf);

/* ---- slab links: ghost rows over NVLink peer memory, no host in the loop (slab_exchange.cu, projection_pack.cu) ---
 * Each slab sim owns one neighbour-writable device block (control words + receive areas) and keeps u, v and their
 * back buffers in one allocation.  Processes trade one opaque blob once (export on the owner, connect on the
 * neighbour; `side` 0 = the neighbour holding the rows above mine in memory, 1 = below) — it carries the CUDA IPC
 * handles of both allocations and the slab's geometry; slabs living in one process connect directly.  Link before
 * the first step.  Once a slab has a neighbour, sayal_step / sayal_run run the whole slab schedule on the sim's
 * stream, graph-captured by sayal_run:
 *   push mode (default; option "slab_push"): every projection pass sweeps the owned rows plus 2 x its iterations of
 *     ghost rows, and the tiles that produce the slab's edge rows store them a second time straight into the
 *     neighbour's ghost rows and publish a flag the neighbour's next pass waits on — compute and exchange are one
 *     kernel, a step needs halo >= max(2 T, advect_margin + 2) ghost rows (T = iterations per pass, at most 8);
 *   otherwise: ghost rows lose two rows of validity per SOR iteration and are refreshed by an exchange kernel only
 *     when the next operation needs more depth than is left (halo >= 2 n + advect_margin + 2: never inside a step).
 * Either way one exchange at the end of the step carries u, v (halo rows) and smoke (advect_margin + 2 rows), hidden
 * under the interior smoke advection.  The ranks must issue the same sequence of steps and use the same halo,
 * advect_margin and temporal_block.  A neighbour that does not answer within 2 s, or that splits the projection
 * into different passes, raises a sticky error: the next sayal_sync / sayal_run / sayal_get_field(s) returns
 * SAYAL_ELINK and the fields are not valid. */
#define SAYAL_LINK_INFO_BYTES 256
int sayal_slab_ipc_export(sayal_sim* sim, void* info_out /* SAYAL_LINK_INFO_BYTES */);
int sayal_slab_ipc_connect(sayal_sim* sim, int32_t side, const void* info /* SAYAL_LINK_INFO_BYTES */);
int sayal_slab_connect_local(sayal_sim* sim, int32_t side, sayal_sim* neighbour);
/* One exchange of the edge rows of the fields in field_mask (1 = U, 2 = V: `halo` rows; 4 = SMOKE: advect_margin + 2
 * rows) with both neighbours. */
int sayal_slab_exchange(sayal_sim* sim, int32_t field_mask);

const char* sayal_last_error(void);
int sayal_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* SAYAL_H */
