// projection_pack.cu — the pressure projection of Fluid::update (/root/reference/src/fluid.cu:229-295) as a
// register-resident, temporally blocked red-black SOR kernel for sm_100a, second generation.
//
// What the reference does: 2n launches per step, each touching u, v, is_solid(int32) x5 and total_s(int32) for
// half the cells from global memory (fluid.cu:264-295).
//
// What this kernel does: one launch ("pass") advances a whole tile by T full iterations (2T half-sweeps)
// without touching global memory in between.
//   * A CTA owns a 128-column x (NW warps * RY rows) tile.  Lane l of warp w keeps cells x = X0 + 4l .. 4l+3 of
//     rows Y0 + w*RY .. +RY-1 in registers, as PACKED PAIRS: {u0,u2}, {u1,u3}, {v0,v2}, {v1,v3}.  The two cells
//     of one colour in a lane are columns (0,2) or (1,3), so one Blackwell packed-fp32 instruction
//     (add/mul/fma.rn.f32x2 -> FADD2/FMUL2/FFMA2) updates both: half the issue slots of the scalar form, same
//     IEEE result per element.
//   * Columns (1,3) need u of the next lane's column 0: one __shfl_down brings it, one __shfl_up returns the
//     updated face.  Vertically adjacent warps share one row of v faces through shared memory (colour-split
//     float2 slots, conflict-free), one barrier per half-sweep.
//   * Cell masks without branches: on the lanes' common path the face updates of apply_projection_at (fluid.cu:247-261)
//     are fma(e, m, face) with m in {+-1, 0} and the divide by total_s is a multiply by inv in {0, 1, 1/2, 1/3, 1/4}
//     (inv = 0 switches a cell off).  Each lane keeps ONE "profile" of these multipliers (the flags of its
//     reference row) in registers; every row whose flags equal the profile — all rows of an open tile, and all
//     rows of a tile that only touches the left/right wall — costs exactly the same instructions as an open
//     row.  Warps with rows that differ (top/bottom wall, obstacle rim) run the "bits" path: per row one byte of face
//     bits and inv from a 256-entry shared-memory table, the four face updates as predicated scalar adds (three
//     registers per row in flight instead of ten multipliers: these warps are latency-bound and pace the pass).
//     fma(e, +-1, x) == x +- e exactly and a closed face is not touched, like the reference's `if`.
//   * Tiles overlap by a halo of 2T cells (rounded up to 4 in x).  Errors from the missing neighbours travel
//     one cell per half-sweep, so after 2T half-sweeps everything at least 2T cells inside the tile is exactly
//     what the global sweep order produces; only that part is written, to the OTHER buffer (ping-pong).
// The update is order-independent within a colour (cells of one colour share no face), so tiling does not
// change a single bit: tests require equality with the plain half-sweep kernel and with the CPU oracle.
//
// Roofline: HBM.  Algorithmic bytes are 17 B per cell per iteration (r/w u, v + 1 flag byte); one pass moves
// (1 + halo overhead) * 9 B in and 8 B out per cell for T iterations.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "sayal_internal.h"

namespace sayal {

namespace {

typedef unsigned long long u64;
constexpr unsigned FULL = 0xffffffffu;
constexpr int TW = 128;  // tile width: 32 lanes x 4 cells

// ---- packed fp32 pairs (element 0 = low register) ------------------------------------------------------
__device__ __forceinline__ u64 pk(float lo, float hi) {
  u64 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ float lo(u64 v) {
  float a, b;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
  return a;
}
__device__ __forceinline__ float hi(u64 v) {
  float a, b;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
  return b;
}
__device__ __forceinline__ u64 add2(u64 a, u64 b) {
  u64 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ u64 sub2(u64 a, u64 b) {
  u64 r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ u64 mul2(u64 a, u64 b) {
  u64 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) {
  u64 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ u64 lds64(const float* p) {
  float2 t = *reinterpret_cast<const float2*>(p);
  return pk(t.x, t.y);
}
__device__ __forceinline__ void sts64(float* p, u64 v) { *reinterpret_cast<float2*>(p) = make_float2(lo(v), hi(v)); }

// The profile multipliers of one pair of cells.  nR is stored negated: face -= e  ==  fma(e, -1, face).
struct Mult {
  u64 inv, mL, nR;
};

__device__ __forceinline__ float inv_of(unsigned nibble) {
  int s = __popc(nibble & 15u);  // total_s of an active cell == number of open faces (build_flags_kernel)
  return s == 0 ? 0.f : (s == 1 ? 1.0f : (s == 2 ? 0.5f : (s == 3 ? 1.0f / 3.0f : 0.25f)));
}
__device__ __forceinline__ float bit_pos(unsigned nibble, unsigned bit) { return (nibble & bit) ? 1.0f : 0.0f; }
__device__ __forceinline__ float bit_neg(unsigned nibble, unsigned bit) { return (nibble & bit) ? -1.0f : 0.0f; }

// Rows that differ from the lane's profile (walls, an obstacle's rim) carry, per colour, the face nibbles of their
// two cells as one byte (cell a low, cell c high: FL_L, FL_R, FL_B, FL_T each) and fetch 1 / total_s of both cells from
// a 256-entry table in shared memory.  The four face updates of apply_projection_at (fluid.cu:247-261) are then
// PREDICATED scalar adds on that byte: a closed face is simply not touched, like the reference's `if`.  (The first
// generation multiplied by {+-1, 0} fetched from a 48-byte table entry: ten registers of multipliers per row in
// flight, so ptxas could overlap hardly any rows and these warps — bound by the latency of a row's chain, not by issue
// slots — paced every pass.  Byte + two floats: three registers per row.)
__device__ __forceinline__ unsigned pair_index(unsigned fl, int c) {
  unsigned t = (fl >> (8 * c)) & 0x000f000fu;  // CASE 0 -> flag bytes 0 and 2, CASE 1 -> bytes 1 and 3
  return (t | (t >> 12)) & 0xffu;
}
__device__ __forceinline__ float add_if(float x, float e, unsigned bits, unsigned bit) {
  asm("{\n\t.reg .pred p;\n\t.reg .b32 t;\n\tand.b32 t, %2, %3;\n\tsetp.ne.u32 p, t, 0;\n\t@p add.rn.f32 %0, %0, %1;\n\t}"
      : "+f"(x) : "f"(e), "r"(bits), "r"(bit));
  return x;
}
__device__ __forceinline__ float sub_if(float x, float e, unsigned bits, unsigned bit) {
  asm("{\n\t.reg .pred p;\n\t.reg .b32 t;\n\tand.b32 t, %2, %3;\n\tsetp.ne.u32 p, t, 0;\n\t@p sub.rn.f32 %0, %0, %1;\n\t}"
      : "+f"(x) : "f"(e), "r"(bits), "r"(bit));
  return x;
}

// In-pass push of a linked y-slab (no reference equivalent: OpenSayal is single-GPU).  A pass of `it`
// iterations needs 2 it exact ghost rows and leaves none; instead of carrying 2 n + margin ghost rows through the
// whole step (recomputed by every slab), the tiles that produce the slab's edge rows store them a second time —
// straight into the neighbour's ghost rows of ITS output arrays, over NVLink — and the last of them publishes a pass
// flag.  The neighbour's next pass waits for that flag only in the warps that load ghost rows.  Ping-pong makes
// the protocol acknowledgement-free: pass k of the neighbour reads the buffer I will write in pass k + 1, and I
// can only start pass k + 1 after its pass-k flag, which it publishes after those reads.
struct PushArgs {
  int on;                    // bit d: a neighbour on side d (0 = low memory rows) takes this pass's edge rows
  int pass_index;            // 0: the ghost rows came with the end-of-step exchange (stream order), nothing to wait for
  int signature;             // (iterations << 8) | passes of the step's plan: both sides must agree
  int debug;                 // profiling only (SAYAL_DEBUG_PUSH; results are wrong): 1 = no peer stores, 2 = no system
                             // fence before the ticket, 4 = no wait for the neighbour's flag
  int src_lo[2], src_hi[2];  // my local rows that travel to side d
  int dst_row0[2];           // row of src_lo[d] in the neighbour's arrays
  int n_pushers[2];          // tiles whose written rows meet [src_lo, src_hi): the last one publishes the flag
  float* peer_u[2];          // the neighbour's output arrays of this pass
  float* peer_v[2];
  unsigned* peer_words[2];   // control words of the neighbours / mine (sayal_internal.h)
  const unsigned* my_words;
  unsigned* ticket;          // [2]
  const unsigned* step_seq;
  int* link_error;
};

// Resident mode (whole domains whose tiles all fit on the GPU at once, one per SM): ONE launch runs the whole
// projection.  Every tile stays in its CTA's registers from the first iteration to the last; between sweep blocks of
// `T` iterations the CTAs trade only what the trapezoid argument says has gone stale — the ring of owned cells within
// one halo of a tile edge — instead of storing and re-loading the whole domain once per pass.  No grid barrier and no
// flags: the rings travel as 8-byte words {value, tag} (a naturally aligned 8-byte store is single-copy atomic), the
// tag counts exchanges, and a reader simply re-reads a word until its tag is the one it waits for — the latency of an
// exchange is one store and one load through L2, with no fence in between.  A word sits at its cell's own place in a
// full-size mailbox array; two mailboxes alternate (exchange parity) so that a ring is never stored over one a
// neighbour may still be reading: a tile can only store exchange k + 2 after it has read its neighbours' exchange
// k + 1, which they stored after they had read its exchange k.
struct ResidentArgs {
  int blocks;        // sweep blocks (exchanges + 1); 0 = not resident
  unsigned* epoch;   // [0] value the tags of this launch count from, [1] ticket of finished tiles
  int* error;        // mapped host word, raised if a neighbour never shows up (the launch is cooperative: cannot happen)
  u64* box_u[2];     // mailboxes by exchange parity, pitch x local_rows words each
  u64* box_v[2];
};

// A pass may run over an explicit list of tiles instead of the regular grid (whole domains, no push): the tiles whose
// warps mostly take the table path (an obstacle's rim) are cut in two along y — each half keeps fewer warps busy and
// finishes with the open tiles instead of after them (a pass ends when its slowest tile does) — and the last tile row
// is moved up so that the bottom wall's two rows are the last two rows of a warp (MODE 4 of half_sweep).
struct TileDesc {
  int X0, Y0;    // first column / local row held
  int y_end;     // rows [Y0, y_end) are held (at most TH of them): rows beyond are treated like rows outside the array
  int vy0, vy1;  // rows written
  int pad[3];
};

struct PackArgs {
  Grid g;
  const float* __restrict__ u_in;
  const float* __restrict__ v_in;
  float* __restrict__ u_out;
  float* __restrict__ v_out;
  const uint8_t* __restrict__ flags;
  float o;
  int iters;     // iterations in this pass (<= T)
  int halo_x;    // 2T rounded up to a multiple of 4
  int halo_y;    // 2T
  int stride_x;  // TW - 2*halo_x
  int stride_y;  // TH - 2*halo_y
  long long* timeline;  // profiling only (option "debug_timeline"): 5 x int64 per CTA, or null
  // External forces folded into the load of a step's first pass (apply_external_forces_at, fluid.cu:308-349, in
  // the form it takes with no drag and no interactive source: v += g d_t everywhere, u = speed and smoke = value
  // on the inlet cells).  Tiles overlap, so a cell's forces are evaluated by every tile that loads it — the
  // same operation on the same input, hence the same bits — and the pass writes every cell exactly once.
  float* p;              // pressure (PRESSURE variants): updated in place, see the kernel
  float density, hf, inv_dt;  // update_pressure_at (fluid.cu:225-226): p += ((e * density) * cell_size) * (1 / d_t)
  // Row window of this pass: local rows [row_lo, row_hi).  A whole domain passes (0, local_rows).  A linked slab
  // passes its owned rows plus the ghost rows that are still exact: ghost rows lose two rows of validity per
  // iteration, and sweeping rows that are already wrong is wasted work.  Rows outside the window are treated
  // like rows outside the array (not loaded, never written); the window's outer edge behaves like a tile edge
  // without a neighbour, which is what costs the next two rows per iteration.
  int row_lo, row_hi;
  int write_lo, write_hi;  // rows this pass may write: the window, or (push mode) the owned rows only — the ghost
                           // rows of the output arrays belong to the neighbours' pushes
  PushArgs push;
  ResidentArgs res;
  int tiles_x;           // tiles per tile row (the grid is one-dimensional: blockIdx.x -> order -> tile)
  const int* order;      // tiles sorted by cost, most expensive first (tile_order_kernel), or null for row-major
  const TileDesc* descs; // or: the pass's tiles as an explicit list, in issue order (one CTA each)
  int extrap_on;  // the step's last pass: apply_extrapolation_at (fluid.cu:720-733) on the tile before it is stored
  int force_on;
  float force_g, force_dt, wt_speed, wt_smoke;
  int inlet_len, band_lo, band_hi, smoke_lo, smoke_hi, smoke_count, smoke_height, period, anchor;
  float* smoke;
};

// forces for the four cells (x..x+3, memory row lr) a lane has just loaded.  Registers only: a store here would
// sit between the tile's loads and serialise them (the inlet smoke is written by forces_inlet_smoke afterwards).
__device__ __forceinline__ bool inlet_band(const PackArgs& a, int lr, bool* smoke_on) {
  const Grid& g = a.g;
  const int j = g.H - 1 - (g.row_base + lr);
  *smoke_on = a.smoke_count == 1 ? (j >= a.smoke_lo && j <= a.smoke_hi)
                                 : (a.period != 0 && (a.anchor - j) % a.period < a.smoke_height);
  return j >= a.band_lo && j <= a.band_hi;
}

__device__ __forceinline__ void forces_on_load(const PackArgs& a, int x, int lr, float4& uu, float4& vv) {
  const Grid& g = a.g;
  if (x + 3 < g.W) {
    vv.x = __fmaf_rn(a.force_g, a.force_dt, vv.x);
    vv.y = __fmaf_rn(a.force_g, a.force_dt, vv.y);
    vv.z = __fmaf_rn(a.force_g, a.force_dt, vv.z);
    vv.w = __fmaf_rn(a.force_g, a.force_dt, vv.w);
  } else {  // the lane that straddles W: pad columns are not cells
    if (x < g.W) vv.x = __fmaf_rn(a.force_g, a.force_dt, vv.x);
    if (x + 1 < g.W) vv.y = __fmaf_rn(a.force_g, a.force_dt, vv.y);
    if (x + 2 < g.W) vv.z = __fmaf_rn(a.force_g, a.force_dt, vv.z);
  }
  if (x <= a.inlet_len) {  // wind-tunnel inlet (fluid.cu:318-330): columns 1..smoke_length of the pipe band
    bool on;
    if (inlet_band(a, lr, &on)) {
      if (x + 0 != 0 && x + 0 <= a.inlet_len && x + 0 < g.W) uu.x = a.wt_speed;
      if (x + 1 <= a.inlet_len && x + 1 < g.W) uu.y = a.wt_speed;
      if (x + 2 <= a.inlet_len && x + 2 < g.W) uu.z = a.wt_speed;
      if (x + 3 <= a.inlet_len && x + 3 < g.W) uu.w = a.wt_speed;
    }
  }
}

// smoke = wind_tunnel.smoke on the inlet cells of the smoke bands (fluid.cu:321-329) for one lane's row
__device__ __forceinline__ void forces_inlet_smoke(const PackArgs& a, int x, int lr) {
  const Grid& g = a.g;
  bool on;
  if (x > a.inlet_len || !inlet_band(a, lr, &on) || !on) return;
  float* sm = a.smoke + (size_t)lr * g.pitch + x;
#pragma unroll
  for (int c = 0; c < 4; c++)
    if (x + c != 0 && x + c <= a.inlet_len && x + c < g.W) sm[c] = a.wt_smoke;
}

__device__ __forceinline__ long long globaltimer() {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// Tagged words of the resident ring exchange: {value, tag} pairs, two per 16-byte access (each 64-bit element of a
// .v2.b64 access is single-copy atomic; relaxed gpu-scope accesses go through L2, never a stale L1 line).
__device__ __forceinline__ void box_store2(u64* p, float a, float b, unsigned tag) {
  const u64 t = (u64)tag << 32;
  asm volatile("st.relaxed.gpu.global.v2.b64 [%0], {%1, %2};" ::"l"(p), "l"(t | __float_as_uint(a)), "l"(t | __float_as_uint(b)) : "memory");
}
__device__ __forceinline__ void box_load2(const u64* p, u64& a, u64& b) {
  asm volatile("ld.relaxed.gpu.global.v2.b64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
}

// Geometry of a resident tile's halo (held, owned by a neighbour): a top and a bottom band of whole rows and a left and
// a right band beside the owned rows.  The halo is fetched by ALL threads of the CTA in 2-cell chunks (one 16-byte
// tagged load each, every load independent) into a staging area in shared memory laid out band by band; the threads
// that hold halo cells then pick theirs up with one 16-byte shared load per row and field.
struct HaloBands {
  int X0, Y0, ox0, ox1, oy0, oy1;  // tile origin, owned columns [ox0, ox1) and rows [oy0, oy1)
  int tw2, hl2, hr2;               // chunks per row of the whole-row bands / the left band / the right band
  int base_b, base_l, base_r, total;  // first chunk of the bottom / left / right band, chunks per field
  // chunk index of cells (lr, x), (lr, x + 1); x - X0 is even
  __device__ __forceinline__ int encode(int lr, int x) const {
    if (lr < oy0) return (lr - Y0) * tw2 + ((x - X0) >> 1);
    if (lr >= oy1) return base_b + (lr - oy1) * tw2 + ((x - X0) >> 1);
    if (x < ox0) return base_l + (lr - oy0) * hl2 + ((x - X0) >> 1);
    return base_r + (lr - oy0) * hr2 + ((x - ox1) >> 1);
  }
  // local row and first column of chunk c
  __device__ __forceinline__ void decode(int c, int& lr, int& x) const {
    if (c < base_b) { lr = Y0 + c / tw2; x = X0 + 2 * (c % tw2); }
    else if (c < base_l) { c -= base_b; lr = oy1 + c / tw2; x = X0 + 2 * (c % tw2); }
    else if (c < base_r) { c -= base_l; lr = oy0 + c / hl2; x = X0 + 2 * (c % hl2); }
    else { c -= base_r; lr = oy0 + c / hr2; x = ox1 + 2 * (c % hr2); }
  }
};

// One row of one half-sweep.  C = 0: columns (0,2) are the active colour, C = 1: columns (1,3).
// vt / vb are the v faces above / below the row for the active columns.
// pressure factors of one step, packed
struct PMul {
  u64 density, hf, inv_dt;
};

// update_pressure_at (fluid.cu:225-226) for the two cells just updated: p += ((e * density) * h) * (1 / d_t) as one
// FMA, the pressure pair living in shared memory.  An inactive cell has e = +-0 and leaves its p unchanged.
__device__ __forceinline__ void pressure_add(float* pp, u64 e, const PMul& pm) {
  sts64(pp, fma2(mul2(mul2(e, pm.density), pm.hf), pm.inv_dt, lds64(pp)));
}

// Profile rows: the lane's multipliers (inv, mL, nR) in registers, both v faces open.
template <int C, bool PRESSURE>
__device__ __forceinline__ void row_step(u64& U02, u64& U13, u64& vb, u64& vt, const Mult& m, u64 o2, int lane, float* pp,
                                         const PMul& pm) {
  if (C == 0) {
    u64 d = sub2(add2(sub2(U13, U02), vt), vb);
    u64 e = mul2(o2, mul2(d, m.inv));
    if (PRESSURE) pressure_add(pp, e, pm);
    U02 = fma2(e, m.mL, U02);
    U13 = fma2(e, m.nR, U13);
    vb = add2(vb, e);
    vt = sub2(vt, e);
  } else {
    // right faces of columns 1 and 3: own column 2 and the next lane's column 0
    float un = __shfl_down_sync(FULL, lo(U02), 1);
    u64 uR = pk(hi(U02), un);
    u64 d = sub2(add2(sub2(uR, U13), vt), vb);
    u64 e = mul2(o2, mul2(d, m.inv));
    if (PRESSURE) pressure_add(pp, e, pm);
    U13 = fma2(e, m.mL, U13);
    uR = fma2(e, m.nR, uR);
    vb = add2(vb, e);
    vt = sub2(vt, e);
    // the updated face of the next lane's column 0 travels back; lane 0 has no left neighbour in the tile
    float back = __shfl_up_sync(FULL, hi(uR), 1);
    U02 = pk(lane == 0 ? lo(U02) : back, lo(uR));
  }
}

// Rows off the profile: inv = 1 / total_s of the two cells (0 switches a cell off), bits = their face nibbles
// (cell a: bits 0..3, cell c: bits 4..7); every face update is a predicated scalar add.
template <int C, bool PRESSURE>
__device__ __forceinline__ void row_step_bits(u64& U02, u64& U13, u64& vb, u64& vt, u64 inv, unsigned bits, u64 o2, int lane,
                                              float* pp, const PMul& pm) {
  u64 uR = 0;
  u64 d;
  if (C == 0) {
    d = sub2(add2(sub2(U13, U02), vt), vb);
  } else {
    float un = __shfl_down_sync(FULL, lo(U02), 1);
    uR = pk(hi(U02), un);
    d = sub2(add2(sub2(uR, U13), vt), vb);
  }
  const u64 e = mul2(o2, mul2(d, inv));
  if (PRESSURE) pressure_add(pp, e, pm);
  const float ea = lo(e), ec = hi(e);
  u64& left = C == 0 ? U02 : U13;   // the left faces of the two cells
  u64& right = C == 0 ? U13 : uR;   // their right faces
  left = pk(add_if(lo(left), ea, bits, FL_L), add_if(hi(left), ec, bits, FL_L << 4));
  right = pk(sub_if(lo(right), ea, bits, FL_R), sub_if(hi(right), ec, bits, FL_R << 4));
  vb = pk(add_if(lo(vb), ea, bits, FL_B), add_if(hi(vb), ec, bits, FL_B << 4));
  vt = pk(sub_if(lo(vt), ea, bits, FL_T), sub_if(hi(vt), ec, bits, FL_T << 4));
  if (C == 1) {
    float back = __shfl_up_sync(FULL, hi(uR), 1);
    U02 = pk(lane == 0 ? lo(U02) : back, lo(uR));
  }
}

// One half-sweep over the RY rows of this warp.  Q0 = case of row 0; the case alternates with the row.
// MODE 0: every row uses the lane's profile multipliers.  MODE 1 (a warp with rows that differ from the profile: walls,
// an obstacle's rim): every row takes the bits path; s_idx holds, per row, the two cases' table indices (16 bits each).
// No per-row branch in either mode (a branch per row keeps ptxas from interleaving the rows).
template <int RY, int Q0, int MODE, bool PRESSURE>
__device__ __forceinline__ void half_sweep(u64 (&U02)[RY], u64 (&U13)[RY], u64 (&V02)[RY - 1], u64 (&V13)[RY - 1],
                                           const Mult (&prof)[2], u64 o2, int lane, float* sv_top, float* sv_bot,
                                           const unsigned* s_idx, const float2* inv_tab, float* sp_warp,
                                           const PMul& pm) {
#pragma unroll
  for (int r = 0; r < RY; r++) {
    const int c = Q0 ^ (r & 1);
    float* pp = sp_warp + r * 128 + 64 * c;  // this lane's pressure pair of row r for the active colour
    u64 vt, vb;
    if (r == 0) vt = lds64(sv_top + 64 * c);
    else vt = c ? V13[r - 1] : V02[r - 1];
    if (r == RY - 1) vb = lds64(sv_bot + 64 * c);
    else vb = c ? V13[r] : V02[r];
    if (MODE == 1) {
      const unsigned word = s_idx[r * 32];
      const unsigned bits = c ? (word >> 16) : (word & 0xffffu);
      const u64 inv = lds64(reinterpret_cast<const float*>(inv_tab + bits));
      if (c == 0) row_step_bits<0, PRESSURE>(U02[r], U13[r], vb, vt, inv, bits, o2, lane, pp, pm);
      else row_step_bits<1, PRESSURE>(U02[r], U13[r], vb, vt, inv, bits, o2, lane, pp, pm);
    } else {
      if (c == 0) row_step<0, PRESSURE>(U02[r], U13[r], vb, vt, prof[0], o2, lane, pp, pm);
      else row_step<1, PRESSURE>(U02[r], U13[r], vb, vt, prof[1], o2, lane, pp, pm);
    }
    if (r == 0) sts64(sv_top + 64 * c, vt);
    else if (c) V13[r - 1] = vt;
    else V02[r - 1] = vt;
    if (r == RY - 1) sts64(sv_bot + 64 * c, vb);
    else if (c) V13[r] = vb;
    else V02[r] = vb;
  }
}

// FORCES: the step's first pass, which also applies the external forces while loading (a separate instantiation,
// so the other passes keep their register allocation).
// EXTRAP: the step's last pass, which also applies the boundary extrapolation before storing.
// PRESSURE (enable_pressure): the tile's pressure lives in dynamic shared memory (TH x 128 floats, packed like the v
// rows) and accumulates one FMA per cell update.  p is updated IN PLACE in global memory: a cell's p is written
// only by the tile that owns it, which read it when the pass began; what other tiles read in their halo never
// leaves them, and nothing else depends on p.
// RESIDENT: the whole projection in one launch, tiles kept in registers between sweep blocks (ResidentArgs); forces and
// extrapolation are compiled in and switched by a.force_on / a.extrap_on.
// PUSH: the pass of a linked slab that stores its edge rows into the neighbours (PushArgs); compiled out of every other
// instantiation (the store phase runs once per tile, cold: every instruction less is an instruction-cache line less).
template <int RY, int NW, bool FORCES, bool EXTRAP, bool PRESSURE, bool RESIDENT = false, bool PUSH = false>
__global__ void __launch_bounds__(NW * 32, 1) projection_pack_kernel(PackArgs a) {
  extern __shared__ __align__(16) float sp[];  // PRESSURE only
  constexpr int TH = RY * NW;
  // shared v rows: sv[0] is the v row above the tile (read-only halo), sv[w+1] is the last row of warp w.
  // Layout per row: [columns (0,2): 64 floats][columns (1,3): 64 floats]; lane l owns float2 at 2l of each.
  __shared__ __align__(16) float sv[NW + 1][128];
  __shared__ float2 inv_tab[256];  // 1 / total_s of the two cells of a face-nibble pair (bits path of half_sweep)
  __shared__ unsigned s_off_all[NW][RY][32];

  const Grid& g = a.g;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  // Tiles with walls or obstacle rims run the table path and take up to 1.6x as long as open tiles.  With more
  // tiles than SMs the pass ends when the last tile does, so the expensive tiles are issued first (the order is
  // computed once per tile geometry from the flags) and the open ones fill in behind them.
  const int tile = a.order ? a.order[blockIdx.x] : (int)blockIdx.x;
  const int tile_y = tile / a.tiles_x, tile_x = tile - tile_y * a.tiles_x;
  int X0 = tile_x * a.stride_x, Y0 = a.row_lo + tile_y * a.stride_y;
  int y_end = a.row_hi;  // held rows end here (a listed tile may end earlier: TileDesc)
  int dvy0 = 0, dvy1 = 0;
  if (!RESIDENT && a.descs) {
    const int4 d = __ldg(reinterpret_cast<const int4*>(a.descs + blockIdx.x));
    const int d4 = __ldg(&a.descs[blockIdx.x].vy1);
    X0 = d.x; Y0 = d.y; y_end = d.z; dvy0 = d.w; dvy1 = d4;
  }
  const int x = X0 + 4 * lane;
  const int lr0 = Y0 + w * RY;
  long long* tl = a.timeline && !RESIDENT ? a.timeline + 5 * (size_t)tile : nullptr;
  if (tl && threadIdx.x == 0) {
    unsigned smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    tl[0] = globaltimer();
    tl[4] = smid;
  }

  if (threadIdx.x < 256) inv_tab[threadIdx.x] = make_float2(inv_of(threadIdx.x & 15u), inv_of(threadIdx.x >> 4));
  // Programmatic dependent launch: let the next pass's CTAs be scheduled as SMs drain, and do not read what the
  // previous kernel wrote before it has completed.  Both are no-ops for a plain launch.
  asm volatile("griddepcontrol.launch_dependents;");
  asm volatile("griddepcontrol.wait;" ::: "memory");

  u64 U02[RY], U13[RY], V02[RY - 1], V13[RY - 1];
  unsigned fl[RY];
  u64 vlast02, vlast13;

  // rows this tile will write (everything >= halo away from an edge that has a neighbouring tile, clipped to the
  // write window), and — linked slabs in push mode — whether they include edge rows a neighbour is waiting for
  const int vy0 = !RESIDENT && a.descs ? dvy0 : max(Y0 == a.row_lo ? a.row_lo : Y0 + a.halo_y, a.write_lo);
  const int vy1 = !RESIDENT && a.descs ? dvy1 : min(Y0 + TH >= a.row_hi ? a.row_hi : Y0 + TH - a.halo_y, a.write_hi);
  const bool push0 = PUSH && (a.push.on & 1) && vy0 < a.push.src_hi[0] && vy1 > a.push.src_lo[0];
  const bool push1 = PUSH && (a.push.on & 2) && vy0 < a.push.src_hi[1] && vy1 > a.push.src_lo[1];
  if (PUSH && a.push.on && a.push.pass_index > 0) {
    // The neighbour's previous pass stored its edge rows into my ghost rows: the warps that load such rows (and
    // warp 0 of a tile that pushes, which must not overwrite rows the neighbour may still be reading) wait for its
    // flag; every other warp goes straight to its loads and the wait hides behind them.
    const unsigned target = (*a.push.step_seq << 10) + (unsigned)a.push.pass_index;
    const bool need0 = (a.push.on & 1) && (lr0 - (w == 0 ? 1 : 0) < a.write_lo || (w == 0 && push0));
    const bool need1 = (a.push.on & 2) && (lr0 + RY > a.write_hi || (w == 0 && push1));
    if ((need0 || need1) && !(a.push.debug & 4)) {
      if (lane == 0) {
        if (need0) spin_until(a.push.my_words + LW_PFLAG + 0, target, a.push.link_error);
        if (need1) spin_until(a.push.my_words + LW_PFLAG + 1, target, a.push.link_error);
      }
      __syncwarp();
    }
  }

  const bool col_ok = x < g.pitch;  // pitch is a multiple of 4: a lane's four columns are in or out together
#pragma unroll
  for (int r = 0; r < RY; r++) {
    int lr = lr0 + r;
    float4 uu = make_float4(0.f, 0.f, 0.f, 0.f), vv = uu;
    unsigned f = 0;
    if (col_ok && lr < y_end) {
      size_t k = (size_t)lr * g.pitch + x;
      // .cg: read through L2 — ghost rows may have been stored by a neighbouring GPU since this SM last saw them,
      // and a tile reads every element exactly once anyway
      uu = __ldcg(reinterpret_cast<const float4*>(a.u_in + k));
      vv = __ldcg(reinterpret_cast<const float4*>(a.v_in + k));
      f = *reinterpret_cast<const unsigned*>(a.flags + k);
    }
    U02[r] = pk(uu.x, uu.z);
    U13[r] = pk(uu.y, uu.w);
    if (r < RY - 1) {
      V02[r] = pk(vv.x, vv.z);
      V13[r] = pk(vv.y, vv.w);
    } else {
      vlast02 = pk(vv.x, vv.z);
      vlast13 = pk(vv.y, vv.w);
    }
    fl[r] = f;
  }
  float* sp_warp = sp + (size_t)(w * RY) * 128 + 2 * lane;
  PMul pm;
  pm.density = pk(a.density, a.density);
  pm.hf = pk(a.hf, a.hf);
  pm.inv_dt = pk(a.inv_dt, a.inv_dt);
  if (PRESSURE) {
#pragma unroll
    for (int r = 0; r < RY; r++) {
      float4 pv = make_float4(0.f, 0.f, 0.f, 0.f);
      if (col_ok && lr0 + r < y_end) pv = *reinterpret_cast<const float4*>(a.p + (size_t)(lr0 + r) * g.pitch + x);
      sts64(sp_warp + r * 128, pk(pv.x, pv.z));
      sts64(sp_warp + r * 128 + 64, pk(pv.y, pv.w));
    }
  }
  if (FORCES && (!RESIDENT || a.force_on)) {
    // Forces on the register tile, AFTER the loop above: anything with control flow between the loads would keep
    // the compiler from issuing them back to back (measured: +5 us per pass).
    const u64 G2 = pk(a.force_g, a.force_g), DT2 = pk(a.force_dt, a.force_dt);
    const bool whole = x + 3 < g.W;  // the lane that straddles W updates only its real columns
#pragma unroll
    for (int r = 0; r < RY; r++) {
      if (!(col_ok && lr0 + r < y_end)) continue;
      u64& p02 = r < RY - 1 ? V02[r < RY - 1 ? r : 0] : vlast02;
      u64& p13 = r < RY - 1 ? V13[r < RY - 1 ? r : 0] : vlast13;
      if (whole) {
        p02 = fma2(G2, DT2, p02);
        p13 = fma2(G2, DT2, p13);
      } else {
        float4 uu = make_float4(lo(U02[r]), lo(U13[r]), hi(U02[r]), hi(U13[r]));
        float4 vv = make_float4(lo(p02), lo(p13), hi(p02), hi(p13));
        forces_on_load(a, x, lr0 + r, uu, vv);
        p02 = pk(vv.x, vv.z);
        p13 = pk(vv.y, vv.w);
      }
    }
    if (x <= a.inlet_len) {  // inlet: u = speed, smoke = value; tiles overlap and every holder stores the same value
#pragma unroll
      for (int r = 0; r < RY; r++) {
        if (!(col_ok && lr0 + r < y_end)) continue;
        float4 uu = make_float4(lo(U02[r]), lo(U13[r]), hi(U02[r]), hi(U13[r]));
        float4 vv = make_float4(0.f, 0.f, 0.f, 0.f);
        forces_on_load(a, x, lr0 + r, uu, vv);  // (only its inlet part matters here)
        U02[r] = pk(uu.x, uu.z);
        U13[r] = pk(uu.y, uu.w);
        forces_inlet_smoke(a, x, lr0 + r);
      }
    }
  }
  // The tile's last column has no column to its right: that cell only lends its faces (carrier).
  if (lane == 31) {
#pragma unroll
    for (int r = 0; r < RY; r++) fl[r] &= 0x00ffffffu;
  }
  // The first row of a slab's local array has no row above it (its top faces are not held): carrier as well.
  // (On a whole domain that row is the top wall, already inactive.)
  if (lr0 == a.row_lo) fl[0] = 0;

  // v faces above the tile's first row
  float* sv_top = &sv[w][2 * lane];
  float* sv_bot = &sv[w + 1][2 * lane];
  sts64(sv_bot, vlast02);
  sts64(sv_bot + 64, vlast13);
  if (w == 0) {
    float4 vv = make_float4(0.f, 0.f, 0.f, 0.f);
    if (col_ok && Y0 - 1 >= a.row_lo && Y0 - 1 < y_end) {
      vv = __ldcg(reinterpret_cast<const float4*>(a.v_in + (size_t)(Y0 - 1) * g.pitch + x));
      if (FORCES && (!RESIDENT || a.force_on)) {
        float4 unused = make_float4(0.f, 0.f, 0.f, 0.f);
        forces_on_load(a, x, Y0 - 1, unused, vv);
      }
    }
    sts64(sv_top, pk(vv.x, vv.z));
    sts64(sv_top + 64, pk(vv.y, vv.w));
  }

  // Profile = the flags of the lane's middle row; a row that differs from it in any lane of the warp is irregular
  // (bit r of irr_rows, warp-uniform).
  const unsigned pf = fl[RY / 2];
  unsigned irr_rows = 0;
  unsigned* s_off = &s_off_all[w][0][lane];
#pragma unroll
  for (int r = 0; r < RY; r++) {
    if (__any_sync(FULL, fl[r] != pf)) irr_rows |= 1u << r;
    s_off[r * 32] = pair_index(fl[r], 0) | (pair_index(fl[r], 1) << 16);
  }
  Mult prof[2];
#pragma unroll
  for (int c = 0; c < 2; c++) {
    unsigned na = (pf >> (8 * c)) & 15u, nc = (pf >> (8 * c + 16)) & 15u;
    prof[c].inv = pk(inv_of(na), inv_of(nc));
    prof[c].mL = pk(bit_pos(na, FL_L), bit_pos(nc, FL_L));
    prof[c].nR = pk(bit_neg(na, FL_R), bit_neg(nc, FL_R));
  }
  // A profile row must have all its B/T faces open for the unmasked v update to be right; if the middle row is
  // itself next to a horizontal boundary, every row goes through the table.
  {
    bool prof_ok = true;
#pragma unroll
    for (int k = 0; k < 4; k++) {
      unsigned n = (pf >> (8 * k)) & 15u;
      if (n != 0 && (n & (FL_B | FL_T)) != (FL_B | FL_T)) prof_ok = false;
    }
    if (__any_sync(FULL, !prof_ok)) irr_rows = (1u << RY) - 1u;
  }
  // which half_sweep the warp runs.  (Warps beyond the rows of a short listed tile hold zeros with inactive flags and
  // run the profile loop on them: a third, barrier-only loop — or any branch around the half-sweeps — makes ptxas
  // shuffle ~20 more registers per half-sweep of the profile loop, 12 % on the 8-row kernel.)
  const int sweep_mode = irr_rows == 0 ? 0 : 1;
  const u64 o2 = pk(a.o, a.o);
  __syncthreads();
  if (tl && threadIdx.x == 0) tl[1] = globaltimer();

  // Colour of column 0 in row 0 of this warp: cell (i, j) belongs to half-sweep `c` iff (i + j + c) is even
  // (fluid.cu:266, 275).  x is a multiple of 4, so the case of row r in half-sweep c is (j0 - r + c) & 1.
  const int j0 = g.H - 1 - (g.row_base + lr0);
  const int q = j0 & 1;  // case of row 0 in the first (even) half-sweep: 0 -> columns (0,2)
  static_assert(RY % 2 == 0, "rows per warp must be even: q must be uniform across the CTA's warps");
  __shared__ unsigned s_res_base;
  if (RESIDENT && threadIdx.x == 0) s_res_base = __ldcg(a.res.epoch);  // advanced by the last tile to finish, i.e. after
                                                                      // every tile has read it; visible to the CTA after
                                                                      // the first barrier of the sweeps
  const int nblk = RESIDENT ? a.res.blocks : 1;
  const int it_base = a.iters / nblk, it_long = a.iters % nblk;
  for (int blk = 0; blk < nblk; blk++) {
    const int its = it_base + (blk < it_long ? 1 : 0);
    // profiling only (option "debug_timeline"): six stamps per tile and block — sweeps begin, sweeps done, ring stored,
    // flag published, neighbours' flags seen, halo loaded
    long long* rtl = RESIDENT && a.timeline && threadIdx.x == 0 ? a.timeline + ((size_t)tile * nblk + blk) * 6 : nullptr;
    if (rtl) rtl[0] = globaltimer();
    if (sweep_mode == 0) {
      for (int it = 0; it < its; it++) {
        if (q == 0) {
          half_sweep<RY, 0, 0, PRESSURE>(U02, U13, V02, V13, prof, o2, lane, sv_top, sv_bot, s_off, inv_tab, sp_warp, pm);
          __syncthreads();
          half_sweep<RY, 1, 0, PRESSURE>(U02, U13, V02, V13, prof, o2, lane, sv_top, sv_bot, s_off, inv_tab, sp_warp, pm);
          __syncthreads();
        } else {
          half_sweep<RY, 1, 0, PRESSURE>(U02, U13, V02, V13, prof, o2, lane, sv_top, sv_bot, s_off, inv_tab, sp_warp, pm);
          __syncthreads();
          half_sweep<RY, 0, 0, PRESSURE>(U02, U13, V02, V13, prof, o2, lane, sv_top, sv_bot, s_off, inv_tab, sp_warp, pm);
          __syncthreads();
        }
      }
    } else if (sweep_mode == 1) {
      for (int it = 0; it < its; it++) {
        if (q == 0) {
          half_sweep<RY, 0, 1, PRESSURE>(U02, U13, V02, V13, prof, o2, lane, sv_top, sv_bot, s_off, inv_tab, sp_warp, pm);
          __syncthreads();
          half_sweep<RY, 1, 1, PRESSURE>(U02, U13, V02, V13, prof, o2, lane, sv_top, sv_bot, s_off, inv_tab, sp_warp, pm);
          __syncthreads();
        } else {
          half_sweep<RY, 1, 1, PRESSURE>(U02, U13, V02, V13, prof, o2, lane, sv_top, sv_bot, s_off, inv_tab, sp_warp, pm);
          __syncthreads();
          half_sweep<RY, 0, 1, PRESSURE>(U02, U13, V02, V13, prof, o2, lane, sv_top, sv_bot, s_off, inv_tab, sp_warp, pm);
          __syncthreads();
        }
      }
    }
    if (rtl) rtl[1] = globaltimer();
    if (!RESIDENT || blk == nblk - 1) break;
    // ---- ring exchange `blk`: my ring out, the neighbours' rings (my halo) in (protocol: ResidentArgs).
    // Everything the exchange needs is recomputed here from the special registers (read through volatile asm, so that
    // the compiler cannot share it with the prologue): nothing of it stays live across the sweeps, whose loop has no
    // register to spare — spilled, it was re-read from local memory row by row, each read a trip to L2.
    {
      int rt_tile, rt_tid;
      asm volatile("mov.u32 %0, %%ctaid.x;" : "=r"(rt_tile));
      asm volatile("mov.u32 %0, %%tid.x;" : "=r"(rt_tid));
      const int e_ty = rt_tile / a.tiles_x, e_tx = rt_tile - e_ty * a.tiles_x;
      const int eX0 = e_tx * a.stride_x, eY0 = a.row_lo + e_ty * a.stride_y;
      const int ex = eX0 + 4 * (rt_tid & 31), elr0 = eY0 + (rt_tid >> 5) * RY;
      const int xe = eX0 + TW < g.pitch ? eX0 + TW : g.pitch, ye = eY0 + TH < a.row_hi ? eY0 + TH : a.row_hi;  // held extent
      HaloBands hb;
      hb.X0 = eX0; hb.Y0 = eY0;
      hb.ox0 = eX0 == 0 ? 0 : eX0 + a.halo_x;
      hb.ox1 = eX0 + TW >= g.pitch ? g.pitch : eX0 + TW - a.halo_x;
      hb.oy0 = eY0 == a.row_lo ? a.row_lo : eY0 + a.halo_y;
      hb.oy1 = eY0 + TH >= a.row_hi ? a.row_hi : eY0 + TH - a.halo_y;
      hb.tw2 = (xe - eX0) >> 1; hb.hl2 = (hb.ox0 - eX0) >> 1; hb.hr2 = (xe - hb.ox1) >> 1;
      hb.base_b = (hb.oy0 - eY0) * hb.tw2;
      hb.base_l = hb.base_b + (ye - hb.oy1) * hb.tw2;
      hb.base_r = hb.base_l + (hb.oy1 - hb.oy0) * hb.hl2;
      hb.total = hb.base_r + (hb.oy1 - hb.oy0) * hb.hr2;
      // ring: owned cells within one halo of an owned edge that has a neighbour (what the neighbours' halos hold);
      // halo: held cells a neighbour owns
      const bool lane_held = ex < xe, lane_owned = ex >= hb.ox0 && ex < hb.ox1;
      const bool lane_ring = lane_owned && ((eX0 != 0 && ex < hb.ox0 + a.halo_x) || (eX0 + TW < g.pitch && ex >= hb.ox1 - a.halo_x));
      const int ring_t = eY0 != a.row_lo ? hb.oy0 + a.halo_y : hb.oy0;       // owned rows below this one are not top ring
      const int ring_b = eY0 + TH < a.row_hi ? hb.oy1 - a.halo_y : hb.oy1;  // owned rows from this one on are bottom ring

      const int par = blk & 1;
      const unsigned tag = s_res_base + 1u + (unsigned)blk;
      u64* bu = a.res.box_u[par];
      u64* bv = a.res.box_v[par];
      vlast02 = lds64(sv_bot);  // my last row of v faces lives in shared memory during the sweeps
      vlast13 = lds64(sv_bot + 64);
      if (lane_owned) {
#pragma unroll
        for (int r = 0; r < RY; r++) {
          const int lr = elr0 + r;
          if (lr < hb.oy0 || lr >= hb.oy1 || !(lane_ring || lr < ring_t || lr >= ring_b)) continue;
          const size_t k = (size_t)lr * g.pitch + ex;
          const u64 p02 = r < RY - 1 ? V02[r < RY - 1 ? r : 0] : vlast02;
          const u64 p13 = r < RY - 1 ? V13[r < RY - 1 ? r : 0] : vlast13;
          box_store2(bu + k, lo(U02[r]), lo(U13[r]), tag);
          box_store2(bu + k + 2, hi(U02[r]), hi(U13[r]), tag);
          box_store2(bv + k, lo(p02), lo(p13), tag);
          box_store2(bv + k + 2, hi(p02), hi(p13), tag);
        }
      }
      if (rtl) rtl[2] = globaltimer();
      // every thread fetches its share of the halo's chunks (u then v), four loads in flight, re-reading a chunk until
      // both of its tags are this exchange's
      float* stage = sp;
      const int total2 = 2 * hb.total;
      long long t0 = 0;
      for (int c0 = rt_tid; c0 < total2; c0 += 4 * NW * 32) {
        u64 w0[4], w1[4];
        const u64* src[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
          const int c = c0 + k * NW * 32;
          src[k] = nullptr;
          if (c < total2) {
            int lr, xx;
            hb.decode(c < hb.total ? c : c - hb.total, lr, xx);
            src[k] = (c < hb.total ? bu : bv) + (size_t)lr * g.pitch + xx;
            box_load2(src[k], w0[k], w1[k]);
          }
        }
#pragma unroll
        for (int k = 0; k < 4; k++) {
          if (!src[k]) continue;
          unsigned polls = 0;
          while ((unsigned)(w0[k] >> 32) != tag || (unsigned)(w1[k] >> 32) != tag) {
            box_load2(src[k], w0[k], w1[k]);
            if ((++polls & 4095u) == 0) {  // half a second without an answer: a broken launch, do not wedge the GPU
              const long long t = globaltimer();
              if (t0 == 0) t0 = t;
              if (t - t0 > 500000000ll) {
                *reinterpret_cast<volatile int*>(a.res.error) = 1;
                __threadfence_system();
                break;
              }
            }
          }
          *reinterpret_cast<float2*>(stage + 2 * (c0 + k * NW * 32)) =
              make_float2(__uint_as_float((unsigned)w0[k]), __uint_as_float((unsigned)w1[k]));
        }
      }
      __syncthreads();
      if (rtl) rtl[4] = globaltimer();
      if (lane_held) {
#pragma unroll
        for (int r = 0; r < RY; r++) {
          const int lr = elr0 + r;
          if (lr >= ye || (lane_owned && lr >= hb.oy0 && lr < hb.oy1)) continue;
          const int c = hb.encode(lr, ex);
          const float4 uu = *reinterpret_cast<const float4*>(stage + 2 * c);
          const float4 vv = *reinterpret_cast<const float4*>(stage + 2 * (hb.total + c));
          U02[r] = pk(uu.x, uu.z);
          U13[r] = pk(uu.y, uu.w);
          if (r < RY - 1) {
            V02[r < RY - 1 ? r : 0] = pk(vv.x, vv.z);
            V13[r < RY - 1 ? r : 0] = pk(vv.y, vv.w);
          } else {
            vlast02 = pk(vv.x, vv.z);
            vlast13 = pk(vv.y, vv.w);
          }
        }
      }
      sts64(sv_bot, vlast02);
      sts64(sv_bot + 64, vlast13);
      __syncthreads();  // the warp below reads these faces in its next half-sweep; the staging area is free again
      if (rtl) rtl[5] = globaltimer();
    }
  }

  vlast02 = lds64(sv_bot);
  vlast13 = lds64(sv_bot + 64);
  if (tl && threadIdx.x == 0) tl[2] = globaltimer();

  // Boundary extrapolation folded into the step's last pass: the closed form of the canonical order (all
  // j-rules, then all i-rules; see extrapolation_kernel in kernels_basic.cu) applied to the registers.  Every
  // source value sits in the same tile as its destination: rows j = 0, 1 (and H-1, H-2) are the last (first) two
  // rows of the domain, inside the un-haloed edge of the tiles that hold them; columns 0, 1 share a lane.
  // Only tiles that hold a border row or column have anything to extrapolate (uniform over the CTA): rows j = H-1,
  // H-2 (memory rows 0, 1 of the domain), j = 1, 0, columns 0, 1 and W-1.
  const bool extrap_tile = EXTRAP && (!RESIDENT || a.extrap_on) &&
                           (X0 == 0 || (((g.W - 1) & ~3) >= X0 && ((g.W - 1) & ~3) < X0 + TW) || g.row_base + Y0 <= 1 ||
                            (g.H - 2 - g.row_base < Y0 + TH && g.H - 1 - g.row_base >= Y0));
  if (extrap_tile) {
    const int lrA = g.H - 2 - g.row_base, lrB = lrA + 1;  // memory rows of j = 1 and j = 0
    // u(i, H-1) = u(i, H-2): memory rows 0 and 1, both rows of warp 0 of the top tiles
    if (g.row_base == 0 && Y0 == 0 && w == 0 && y_end >= 2) {
      U02[0] = U02[1];
      U13[0] = U13[1];
    }
    // u(i, 0) = u(i, 1): the two rows may sit in different warps, so row j = 1 travels through shared memory
    if (lrA >= Y0 && lrA >= 0 && lrB < y_end && lrB < Y0 + TH) {  // uniform over the CTA
      float4* scratch = reinterpret_cast<float4*>(&sv[0][0]);
#pragma unroll
      for (int r = 0; r < RY; r++)
        if (lr0 + r == lrA) scratch[lane] = make_float4(lo(U02[r]), lo(U13[r]), hi(U02[r]), hi(U13[r]));
      __syncthreads();
#pragma unroll
      for (int r = 0; r < RY; r++)
        if (lr0 + r == lrB) {
          float4 t = scratch[lane];
          U02[r] = pk(t.x, t.z);
          U13[r] = pk(t.y, t.w);
        }
    }
    const int xl = (g.W - 1) & ~3, cl = (g.W - 1) & 3;  // lane and component of column W-1
#pragma unroll
    for (int r = 0; r < RY; r++) {
      u64 p02 = r < RY - 1 ? V02[r < RY - 1 ? r : 0] : vlast02;
      u64 p13 = r < RY - 1 ? V13[r < RY - 1 ? r : 0] : vlast13;
      if (lr0 + r == lrA) {  // v(i, 1) = 0
        p02 = 0ull;
        p13 = 0ull;
      } else {
        if (x == 0) p02 = pk(lo(p13), hi(p02));  // v(0, j) = v(1, j)
        if (cl == 0) {                           // v(W-1, j) = v(W-2, j): column W-2 is the previous lane's last
          float prev = __shfl_up_sync(FULL, hi(p13), 1);
          if (x == xl) p02 = pk(prev, hi(p02));
        } else if (x == xl) {
          if (cl == 1) p13 = pk(lo(p02), hi(p13));
          else if (cl == 2) p02 = pk(lo(p02), lo(p13));
          else p13 = pk(lo(p13), hi(p02));
        }
      }
      if (x == 0) U13[r] = pk(0.f, hi(U13[r]));  // u(1, j) = 0
      if (r < RY - 1) {
        V02[r < RY - 1 ? r : 0] = p02;
        V13[r < RY - 1 ? r : 0] = p13;
      } else {
        vlast02 = p02;
        vlast13 = p13;
      }
    }
  }

  // write the part of the tile that is exact: everything >= halo away from an edge that has a neighbour
  const int vx0 = X0 == 0 ? 0 : X0 + a.halo_x;
  const int vx1 = X0 + TW >= g.pitch ? g.pitch : X0 + TW - a.halo_x;
  if (x >= vx0 && x < vx1) {
#pragma unroll
    for (int r = 0; r < RY; r++) {
      int lr = lr0 + r;
      if (lr >= vy0 && lr < vy1) {
        size_t k = (size_t)lr * g.pitch + x;
        const float4 uo = make_float4(lo(U02[r]), lo(U13[r]), hi(U02[r]), hi(U13[r]));
        *reinterpret_cast<float4*>(a.u_out + k) = uo;
        u64 p02 = r < RY - 1 ? V02[r < RY - 1 ? r : 0] : vlast02;
        u64 p13 = r < RY - 1 ? V13[r < RY - 1 ? r : 0] : vlast13;
        const float4 vo = make_float4(lo(p02), lo(p13), hi(p02), hi(p13));
        *reinterpret_cast<float4*>(a.v_out + k) = vo;
        if (PRESSURE) {
          u64 q02 = lds64(sp_warp + r * 128), q13 = lds64(sp_warp + r * 128 + 64);
          *reinterpret_cast<float4*>(a.p + k) = make_float4(lo(q02), lo(q13), hi(q02), hi(q13));
        }
        if (!PUSH || (a.push.debug & 1)) continue;
        if (push0 && lr >= a.push.src_lo[0] && lr < a.push.src_hi[0]) {  // my edge rows = the neighbour's ghost rows
          size_t kp = (size_t)(lr - a.push.src_lo[0] + a.push.dst_row0[0]) * g.pitch + x;
          *reinterpret_cast<float4*>(a.push.peer_u[0] + kp) = uo;
          *reinterpret_cast<float4*>(a.push.peer_v[0] + kp) = vo;
        }
        if (push1 && lr >= a.push.src_lo[1] && lr < a.push.src_hi[1]) {
          size_t kp = (size_t)(lr - a.push.src_lo[1] + a.push.dst_row0[1]) * g.pitch + x;
          *reinterpret_cast<float4*>(a.push.peer_u[1] + kp) = uo;
          *reinterpret_cast<float4*>(a.push.peer_v[1] + kp) = vo;
        }
      }
    }
  }
  if (PUSH && (push0 || push1)) {  // uniform over the CTA
    __syncthreads();     // every thread's stores are issued before thread 0 fences at system scope
    if (threadIdx.x == 0) {
      if (!(a.push.debug & 2)) __threadfence_system();
      const unsigned published = (*a.push.step_seq << 10) + (unsigned)a.push.pass_index + 1u;
#pragma unroll
      for (int d = 0; d < 2; d++) {
        if (!(d == 0 ? push0 : push1)) continue;
        if (atomicAdd(&a.push.ticket[d], 1u) == (unsigned)a.push.n_pushers[d] - 1u) {  // the side's last tile
          a.push.ticket[d] = 0;
          unsigned* w = a.push.peer_words[d];
          w[LW_PITER + (1 - d)] = (unsigned)a.push.signature;
          __threadfence_system();
          st_release_sys(w + LW_PFLAG + (1 - d), published);
        }
      }
    }
  }
  if (RESIDENT) {  // the last tile to finish moves the epoch past every tag of this launch
    if (threadIdx.x == 0 && atomicAdd(a.res.epoch + 1, 1u) == gridDim.x - 1) {
      a.res.epoch[1] = 0;
      __threadfence();
      a.res.epoch[0] = s_res_base + (unsigned)nblk + 1u;
    }
  }
  if (tl) {
    __syncthreads();
    if (threadIdx.x == 0) tl[3] = globaltimer();
  }
}

// Cost class of every tile of one geometry: the number of its warps that take the table path (same test as the
// pack kernel: a row whose flags differ from the warp's middle row, or a middle row next to a horizontal
// boundary).  One CTA per tile, thread layout as in the pack kernel.
// edge_first (push mode): the first / last tile row (bit 0 / bit 1) produces the rows a neighbouring GPU waits for
// and goes to the front of the issue order whatever it costs.
__global__ void tile_cost_kernel(Grid g, const uint8_t* __restrict__ flags, int ry, int stride_x, int stride_y, int tiles_x,
                                 int tiles_y, int edge_first, int row_lo, int row_hi, int* __restrict__ cost) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int tile = blockIdx.x, tile_y = tile / tiles_x, tile_x = tile - tile_y * tiles_x;
  const int x = tile_x * stride_x + 4 * lane, lr0 = row_lo + tile_y * stride_y + w * ry;
  auto flag_word = [&](int r) -> unsigned {
    int lr = lr0 + r;
    unsigned f = 0;
    if (x < g.pitch && lr < row_hi) f = *reinterpret_cast<const unsigned*>(flags + (size_t)lr * g.pitch + x);
    if (lane == 31) f &= 0x00ffffffu;
    if (lr == row_lo) f = 0;
    return f;
  };
  const unsigned pf = flag_word(ry / 2);
  bool irr = false;
  for (int r = 0; r < ry; r++)
    if (flag_word(r) != pf) irr = true;
  for (int k = 0; k < 4; k++) {
    unsigned n = (pf >> (8 * k)) & 15u;
    if (n != 0 && (n & (FL_B | FL_T)) != (FL_B | FL_T)) irr = true;
  }
  irr = __any_sync(FULL, irr);
  __shared__ int s_count;
  if (threadIdx.x == 0) s_count = 0;
  __syncthreads();
  if (lane == 0 && irr) atomicAdd(&s_count, 1);
  __syncthreads();
  const bool edge = ((edge_first & 1) && tile_y == 0) || ((edge_first & 2) && tile_y == tiles_y - 1);
  if (threadIdx.x == 0) cost[tile] = edge ? 32 : s_count;
}

// Counting sort of the tiles by cost, most expensive first (one CTA; costs are 0..32; ties in any order).
__global__ void tile_order_kernel(const int* __restrict__ cost, int tiles, int* __restrict__ order) {
  __shared__ int hist[33], start[33];
  if (threadIdx.x < 33) hist[threadIdx.x] = 0;
  __syncthreads();
  for (int t = threadIdx.x; t < tiles; t += blockDim.x) atomicAdd(&hist[min(cost[t], 32)], 1);
  __syncthreads();
  if (threadIdx.x == 0) {
    int at = 0;
    for (int c = 32; c >= 0; c--) { start[c] = at; at += hist[c]; }
  }
  __syncthreads();
  for (int t = threadIdx.x; t < tiles; t += blockDim.x) order[atomicAdd(&start[min(cost[t], 32)], 1)] = t;
}

// Sweep cost of every listed tile, in tenths of an open warp: a warp costs 10 (profile path), 11 (wall rows from the
// table), 16 (every row from the table) or nothing (beyond the tile's rows); warp w issues on scheduler w & 3 and a
// half-sweep ends when the busiest scheduler does, so the tile's cost is the largest of the four sums.  cost[2 t] =
// that, cost[2 t + 1] = the number of warps on the all-table path.  One CTA per tile, thread layout of the pack kernel.
__global__ void desc_cost_kernel(Grid g, const uint8_t* __restrict__ flags, int ry, int row_lo, const TileDesc* __restrict__ descs,
                                 int* __restrict__ cost) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const TileDesc d = descs[blockIdx.x];
  const int x = d.X0 + 4 * lane, lr0 = d.Y0 + w * ry;
  auto flag_word = [&](int r) -> unsigned {
    int lr = lr0 + r;
    unsigned f = 0;
    if (x < g.pitch && lr < d.y_end) f = *reinterpret_cast<const unsigned*>(flags + (size_t)lr * g.pitch + x);
    if (lane == 31) f &= 0x00ffffffu;
    if (lr == row_lo) f = 0;
    return f;
  };
  const unsigned pf = flag_word(ry / 2);
  unsigned irr_rows = 0;
  for (int r = 0; r < ry; r++)
    if (__any_sync(FULL, flag_word(r) != pf)) irr_rows |= 1u << r;
  bool prof_ok = true;
  for (int k = 0; k < 4; k++) {
    unsigned n = (pf >> (8 * k)) & 15u;
    if (n != 0 && (n & (FL_B | FL_T)) != (FL_B | FL_T)) prof_ok = false;
  }
  if (__any_sync(FULL, !prof_ok)) irr_rows = (1u << ry) - 1u;
  const int units = lr0 >= d.y_end ? 0 : irr_rows == 0 ? 10 : ((irr_rows & ~3u) == 0 || (irr_rows & ~(3u << (ry - 2))) == 0) ? 11 : 16;
  __shared__ int s_sched[4], s_table;
  if (threadIdx.x < 4) s_sched[threadIdx.x] = 0;
  if (threadIdx.x == 0) s_table = 0;
  __syncthreads();
  if (lane == 0) {
    atomicAdd(&s_sched[w & 3], units);
    if (units == 16) atomicAdd(&s_table, 1);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    cost[2 * blockIdx.x] = max(max(s_sched[0], s_sched[1]), max(s_sched[2], s_sched[3]));
    cost[2 * blockIdx.x + 1] = s_table;
  }
}

struct Variant {
  int ry, nw;
  void (*kernel[6])(PackArgs);  // indexed by (forces on load) | (extrapolation before store) << 1; [4] = with pressure;
                                // [5] = resident (whole projection in one launch, forces / extrapolation by flag)
  void (*push_kernel[4])(PackArgs);  // the first four again, for passes that push their edge rows (linked slabs)
};
#define SAYAL_PACK_VARIANT(RY, NW)                                                                                        \
  {RY, NW, {projection_pack_kernel<RY, NW, false, false, false>, projection_pack_kernel<RY, NW, true, false, false>,     \
            projection_pack_kernel<RY, NW, false, true, false>, projection_pack_kernel<RY, NW, true, true, false>,       \
            projection_pack_kernel<RY, NW, false, false, true>, projection_pack_kernel<RY, NW, true, true, false, true>},  \
           {projection_pack_kernel<RY, NW, false, false, false, false, true>,                                            \
            projection_pack_kernel<RY, NW, true, false, false, false, true>,                                             \
            projection_pack_kernel<RY, NW, false, true, false, false, true>,                                             \
            projection_pack_kernel<RY, NW, true, true, false, false, true>}}
const Variant kVariants[] = {SAYAL_PACK_VARIANT(8, 16), SAYAL_PACK_VARIANT(10, 16), SAYAL_PACK_VARIANT(12, 16)};
constexpr int kNumVariants = sizeof(kVariants) / sizeof(kVariants[0]);
constexpr int kMaxT = 16;
constexpr int kMaxPushT = 8;   // push mode: iterations per pass at most (the halo allowing)

int tiles_for(int extent, int tile, int stride) {
  if (extent <= tile) return 1;
  return (extent - tile + stride - 1) / stride + 1;
}

struct Geometry {
  int halo_x, halo_y, stride_x, stride_y, tiles_x, tiles_y;
};

bool geometry(const Grid& g, const Variant& v, int T, Geometry* out, int rows = -1) {
  if (rows < 0) rows = g.local_rows;
  int th = v.ry * v.nw;
  out->halo_y = 2 * T;
  out->halo_x = (2 * T + 3) & ~3;
  out->stride_x = TW - 2 * out->halo_x;
  out->stride_y = th - 2 * out->halo_y;
  if (out->stride_x < TW / 4 || out->stride_y < th / 4) return false;  // keep at least a quarter of the tile useful
  out->tiles_x = tiles_for(g.pitch, TW, out->stride_x);
  out->tiles_y = tiles_for(rows, th, out->stride_y);
  return true;
}

// model of one projection of n iterations, in arbitrary units: passes x waves x (load/store + T sweeps) per row
double model_cost(const Grid& g, const Variant& v, int T, int n, int sms) {
  Geometry q;
  if (!geometry(g, v, T, &q)) return 1e30;
  int passes = (n + T - 1) / T;
  long tiles = (long)q.tiles_x * q.tiles_y;
  long waves = (tiles + sms - 1) / sms;
  return (double)passes * waves * v.ry * (0.45 + 0.13 * T);
}

size_t pressure_smem(const Variant& v) { return (size_t)v.ry * v.nw * 128 * sizeof(float); }

int launch_pass(Sim* s, const Variant& v, const PackArgs& a, dim3 grid, cudaStream_t stream) {
  cudaLaunchConfig_t lc = {};
  lc.gridDim = grid;
  lc.blockDim = dim3(v.nw * 32);
  lc.dynamicSmemBytes = a.p ? pressure_smem(v) : 0;
  lc.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  lc.attrs = attr;
  lc.numAttrs = s->use_pdl ? 1 : 0;
  const int which = (a.force_on ? 1 : 0) | (a.extrap_on ? 2 : 0);
  if (a.push.on && a.p) return set_error(SAYAL_EINVAL, "projection push: not available with enable_pressure");
  cudaError_t e = cudaLaunchKernelEx(&lc, a.p ? v.kernel[4] : a.push.on ? v.push_kernel[which] : v.kernel[which], a);
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) {
    char m[256];
    snprintf(m, sizeof m, "projection_pack_kernel: %s", cudaGetErrorString(e));
    return set_error(SAYAL_ECUDA, m);
  }
  s->launches++;
  return SAYAL_OK;
}

// Device array with the tiles of geometry (variant, iterations-per-pass) in issue order; built on first use
// (two small kernels on the sim's stream) and cached.  Returns null — row-major order — when the cache is full,
// the option is off, or the stream is being captured and the geometry has not been seen before.
const int* tile_order(Sim* s, int variant, int it, const Geometry& q, int row_lo, int row_hi, int edge_first = 0) {
  if (!s->order_tiles) return nullptr;
  for (int k = 0; k < s->n_orders; k++)
    if (s->orders[k].variant == variant && s->orders[k].it == it && s->orders[k].row_lo == row_lo &&
        s->orders[k].row_hi == row_hi && s->orders[k].edge_first == edge_first)
      return s->orders[k].order;
  const int tiles = q.tiles_x * q.tiles_y;
  if (s->n_orders == Sim::kMaxOrders || tiles <= 1) return nullptr;
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  cudaStreamIsCapturing(s->stream, &cap);
  if (cap != cudaStreamCaptureStatusNone) return nullptr;
  int* buf = nullptr;
  if (cudaMalloc(&buf, sizeof(int) * 2 * (size_t)tiles) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  const Variant& v = kVariants[variant];
  tile_cost_kernel<<<tiles, v.nw * 32, 0, s->stream>>>(s->g, s->flags, v.ry, q.stride_x, q.stride_y, q.tiles_x, q.tiles_y,
                                                      edge_first, row_lo, row_hi, buf + tiles);
  tile_order_kernel<<<1, 1024, 0, s->stream>>>(buf + tiles, tiles, buf);
  if (cudaGetLastError() != cudaSuccess) {
    cudaFree(buf);
    return nullptr;
  }
  s->orders[s->n_orders++] = {variant, it, row_lo, row_hi, edge_first, buf, 0};
  return buf;
}

// The explicit tile list of a whole-domain pass (TileDesc): the regular grid with (1) its last tile row moved up so
// that the array's last two rows end a warp, (2) every tile whose busiest scheduler carries much more than four open
// warps' worth of work (an obstacle rim: warps on the all-table path) cut in two along y, (3) the tiles sorted by
// cost, most expensive first.  Built once per geometry outside graph capture (costs are read back) and cached with
// the issue orders; null when the option is off, the cache is full or a graph is being captured.
const TileDesc* tile_descs(Sim* s, int variant, int it, const Geometry& q, int* n_out) {
  *n_out = 0;
  if (!s->split_tiles) return nullptr;
  const int rows = s->g.local_rows;
  for (int k = 0; k < s->n_orders; k++)
    if (s->orders[k].variant == variant && s->orders[k].it == it && s->orders[k].row_lo == 0 && s->orders[k].row_hi == rows &&
        s->orders[k].edge_first == -1) {
      *n_out = s->orders[k].n_descs;
      return reinterpret_cast<const TileDesc*>(s->orders[k].order);
    }
  if (s->n_orders == Sim::kMaxOrders) return nullptr;
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  cudaStreamIsCapturing(s->stream, &cap);
  if (cap != cudaStreamCaptureStatusNone) return nullptr;
  const Variant& v = kVariants[variant];
  const int th = v.ry * v.nw;
  std::vector<TileDesc> tiles;
  for (int ty = 0; ty < q.tiles_y; ty++)
    for (int tx = 0; tx < q.tiles_x; tx++) {
      TileDesc d = {};
      d.X0 = tx * q.stride_x;
      d.Y0 = ty * q.stride_y;
      d.vy0 = ty == 0 ? 0 : d.Y0 + q.halo_y;
      d.vy1 = d.Y0 + th >= rows ? rows : d.Y0 + th - q.halo_y;
      d.y_end = d.Y0 + th < rows ? d.Y0 + th : rows;
      if (ty == q.tiles_y - 1 && ty > 0) d.Y0 -= (v.ry - (rows - d.Y0) % v.ry) % v.ry;  // the last rows end a warp
      tiles.push_back(d);
    }
  int* d_buf = nullptr;   // descriptors, then two ints of cost per tile (sized for every tile being split)
  const size_t cap_tiles = 2 * tiles.size();
  if (cudaMalloc(&d_buf, cap_tiles * (sizeof(TileDesc) + 2 * sizeof(int))) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  TileDesc* d_descs = reinterpret_cast<TileDesc*>(d_buf);
  int* d_cost = reinterpret_cast<int*>(d_descs + cap_tiles);
  std::vector<int> cost;
  auto evaluate = [&]() -> bool {
    cost.assign(2 * tiles.size(), 0);
    if (cudaMemcpyAsync(d_descs, tiles.data(), tiles.size() * sizeof(TileDesc), cudaMemcpyHostToDevice, s->stream) != cudaSuccess) return false;
    desc_cost_kernel<<<(int)tiles.size(), v.nw * 32, 0, s->stream>>>(s->g, s->flags, v.ry, 0, d_descs, d_cost);
    if (cudaMemcpyAsync(cost.data(), d_cost, cost.size() * sizeof(int), cudaMemcpyDeviceToHost, s->stream) != cudaSuccess) return false;
    return cudaStreamSynchronize(s->stream) == cudaSuccess;
  };
  bool ok = evaluate();
  if (ok) {  // cut the expensive tiles in two (at most an eighth of the tiles)
    std::vector<TileDesc> out;
    int cuts = 0;
    const int max_cuts = (int)tiles.size() / 8 + 1;
    for (size_t t = 0; t < tiles.size(); t++) {
      const TileDesc& d = tiles[t];
      const int owned = d.vy1 - d.vy0;
      if (cost[2 * t] > 46 && cost[2 * t + 1] >= 3 && owned >= 4 * v.ry && cuts < max_cuts) {
        const int mid = d.vy0 + owned / 2;
        TileDesc a = d, b = d;
        a.vy1 = mid;
        a.y_end = mid + q.halo_y < d.y_end ? mid + q.halo_y : d.y_end;
        b.vy0 = mid;
        b.Y0 = mid - q.halo_y > d.Y0 ? mid - q.halo_y : d.Y0;
        out.push_back(a);
        out.push_back(b);
        cuts++;
      } else {
        out.push_back(d);
      }
    }
    if (cuts) {
      tiles.swap(out);
      ok = evaluate();
    }
  }
  if (ok) {  // most expensive first
    std::vector<int> idx(tiles.size());
    for (size_t t = 0; t < idx.size(); t++) idx[t] = (int)t;
    std::stable_sort(idx.begin(), idx.end(), [&](int x, int y) { return cost[2 * x] > cost[2 * y]; });
    std::vector<TileDesc> sorted(tiles.size());
    for (size_t t = 0; t < idx.size(); t++) sorted[t] = tiles[idx[t]];
    ok = cudaMemcpyAsync(d_descs, sorted.data(), sorted.size() * sizeof(TileDesc), cudaMemcpyHostToDevice, s->stream) == cudaSuccess &&
         cudaStreamSynchronize(s->stream) == cudaSuccess;
  }
  if (!ok) {
    cudaGetLastError();
    cudaFree(d_buf);
    return nullptr;
  }
  s->orders[s->n_orders++] = {variant, it, 0, rows, -1, d_buf, (int)tiles.size()};
  *n_out = (int)tiles.size();
  return d_descs;
}

// The passes of one projection call: ceil(iterations / T) passes of nearly equal size (25 at T = 10 -> 9, 8, 8 rather
// than 10, 10, 5: the same number of loads and stores, but narrower halos on every pass).  Row window of pass k:
//   whole domain (ghost_depth < 0, push_sides == 0): every row;
//   linked slab, deep halo (ghost_depth >= 0): the owned rows plus the ghost rows still exact when the pass starts;
//   linked slab, push mode (push_sides != 0): the owned rows plus 2 it ghost rows on every side that has a neighbour —
//   exactly what `it` iterations consume; the pass writes the owned rows only and its edge tiles store the rows
//   [own_lo, own_lo + halo) / [own_hi - halo, own_hi) into the neighbours' ghost rows as well (PushArgs).
// run_passes launches this list; the order cache and sayal_debug_pass_plans (the CPU tests of the covering
// invariants) read the same list.
struct PassPlan {
  int iterations, row_lo, row_hi, write_lo, write_hi;
  int src_lo[2], src_hi[2], n_pushers[2];
  Geometry q;
};

int make_pass_plans(const Grid& g, const Variant& v, int T, int iterations, int ghost_depth, int push_sides, int halo,
                    PassPlan* out, int capacity) {
  if (iterations <= 0 || T <= 0) return 0;
  const int passes = (iterations + T - 1) / T;
  const int base = iterations / passes, longer = iterations % passes;
  const int th = v.ry * v.nw;
  int done = 0, n = 0;
  for (int pass = 0; pass < passes && n < capacity; pass++) {
    PassPlan p = {};
    p.iterations = base + (pass < longer ? 1 : 0);
    p.row_lo = 0;
    p.row_hi = g.local_rows;
    if (push_sides) {
      if (2 * p.iterations > halo) return -1;  // a pass may not consume more ghost rows than a push delivers
      if (push_sides & 1) p.row_lo = g.own_lo - 2 * p.iterations < 0 ? 0 : g.own_lo - 2 * p.iterations;
      if (push_sides & 2) p.row_hi = g.own_hi + 2 * p.iterations > g.local_rows ? g.local_rows : g.own_hi + 2 * p.iterations;
    } else if (ghost_depth >= 0) {
      const int depth = ghost_depth - 2 * done;
      p.row_lo = g.own_lo - depth < 0 ? 0 : g.own_lo - depth;
      p.row_hi = g.own_hi + depth > g.local_rows ? g.local_rows : g.own_hi + depth;
    }
    p.write_lo = (push_sides & 1) ? g.own_lo : p.row_lo;
    p.write_hi = (push_sides & 2) ? g.own_hi : p.row_hi;
    if (!geometry(g, v, p.iterations, &p.q, p.row_hi - p.row_lo)) return -1;
    p.src_lo[0] = g.own_lo; p.src_hi[0] = g.own_lo + halo;
    p.src_lo[1] = g.own_hi - halo; p.src_hi[1] = g.own_hi;
    for (int d = 0; d < 2; d++) {
      p.n_pushers[d] = 0;
      if (!(push_sides & (1 << d))) continue;
      for (int ty = 0; ty < p.q.tiles_y; ty++) {  // the kernel's vy0 / vy1 / push0 / push1
        const int y0 = p.row_lo + ty * p.q.stride_y;
        int vy0 = y0 == p.row_lo ? p.row_lo : y0 + p.q.halo_y;
        int vy1 = y0 + th >= p.row_hi ? p.row_hi : y0 + th - p.q.halo_y;
        if (vy0 < p.write_lo) vy0 = p.write_lo;
        if (vy1 > p.write_hi) vy1 = p.write_hi;
        if (vy0 < p.src_hi[d] && vy1 > p.src_lo[d]) p.n_pushers[d] += p.q.tiles_x;
      }
    }
    out[n++] = p;
    done += p.iterations;
  }
  return n;
}

constexpr int kMaxPasses = 256;

// which sides of a linked slab take pushes (0: the call does not push)
int push_sides_of(const Sim* s) {
  if (!s->push_active) return 0;
  return (s->link.peer_words[0] ? 1 : 0) | (s->link.peer_words[1] ? 2 : 0);
}

// Resident mode (ResidentArgs): is it possible for this sim / plan, and the launch itself.
constexpr size_t kResidentStageMax = 176 * 1024;  // dynamic shared memory a resident launch may ask for

bool resident_possible(const Sim* s) {
  return s->resident && s->d_resident && !s->ph.enable_pressure && s->g.local_rows == s->g.H;
}

// bytes of shared staging for the halo of one tile: 2 fields x 4 bytes per halo cell (upper bound: an interior tile)
size_t resident_stage_bytes(const Variant& v, const Geometry& q) {
  const size_t held = (size_t)TW * v.ry * v.nw, owned = (size_t)q.stride_x * q.stride_y;
  return 8 * (held - owned);
}

bool resident_geometry(const Sim* s, const Variant& v, int T, Geometry* q) {
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, s->device);
  return geometry(s->g, v, T, q) && q->tiles_x * q->tiles_y <= sms && resident_stage_bytes(v, *q) <= kResidentStageMax;
}

// The four mailbox arrays (two parities x u, v; 8 bytes per cell), allocated the first time a resident plan is
// considered (never while a graph is being captured: plans are chosen before).
bool resident_mailboxes(Sim* s) {
  if (s->d_box) return true;
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  cudaStreamIsCapturing(s->stream, &cap);
  if (cap != cudaStreamCaptureStatusNone) return false;
  const size_t words = (size_t)s->g.pitch * s->g.local_rows;
  if (cudaMalloc(&s->d_box, 4 * words * sizeof(u64)) != cudaSuccess) {
    cudaGetLastError();
    s->d_box = nullptr;
    return false;
  }
  cudaMemsetAsync(s->d_box, 0, 4 * words * sizeof(u64), s->stream);  // tag 0: never a tag of an exchange
  return true;
}

int run_resident(Sim* s, int variant, int T, int iterations, float d_t, bool with_forces, bool with_extrap) {
  const Variant& v = kVariants[variant];
  Geometry q;
  if (!resident_possible(s) || !resident_geometry(s, v, T, &q))
    return set_error(SAYAL_EINVAL, "resident projection: the tiles of this plan do not fit on the GPU at once");
  PackArgs a = {};
  a.g = s->g;
  a.row_lo = a.write_lo = 0;
  a.row_hi = a.write_hi = s->g.local_rows;
  a.u_in = s->u; a.v_in = s->v; a.u_out = s->u_buf; a.v_out = s->v_buf;
  a.flags = s->flags;
  a.o = s->ph.o;
  a.iters = iterations;
  a.halo_x = q.halo_x; a.halo_y = q.halo_y; a.stride_x = q.stride_x; a.stride_y = q.stride_y;
  a.density = s->ph.density;
  a.hf = (float)s->g.h;
  a.inv_dt = 1.0f / d_t;
  a.extrap_on = with_extrap;
  a.force_on = with_forces;
  a.smoke = s->smoke;
  if (a.force_on) {
    const ForceArgs& f = s->fuse_args;
    const Phys& ph = s->ph;
    a.force_g = ph.g; a.force_dt = f.d_t; a.wt_speed = ph.wt_speed; a.wt_smoke = ph.wt_smoke;
    a.inlet_len = ph.wt_smoke_length; a.band_lo = f.band_lo; a.band_hi = f.band_hi;
    a.smoke_lo = f.smoke_lo; a.smoke_hi = f.smoke_hi; a.smoke_count = ph.wt_smoke_count;
    a.smoke_height = ph.wt_smoke_height; a.period = f.period; a.anchor = f.anchor;
  }
  a.tiles_x = q.tiles_x;
  if (!s->d_box) return set_error(SAYAL_EINVAL, "resident projection: no mailboxes (plan chosen without them?)");
  a.res.blocks = (iterations + T - 1) / T;
  a.res.epoch = s->d_resident;
  a.res.error = s->d_resident_error;
  {
    const size_t words = (size_t)s->g.pitch * s->g.local_rows;
    u64* box = reinterpret_cast<u64*>(s->d_box);
    a.res.box_u[0] = box; a.res.box_v[0] = box + words;
    a.res.box_u[1] = box + 2 * words; a.res.box_v[1] = box + 3 * words;
  }
  {  // profiling only: 6 stamps per tile and block, read back through sayal_debug_timeline as rows of 5 int64
    const size_t stamps = (size_t)q.tiles_x * q.tiles_y * a.res.blocks * 6;
    a.timeline = s->d_timeline && stamps <= s->timeline_cap ? s->d_timeline : nullptr;
    s->timeline_tiles = a.timeline ? (int)((stamps + 4) / 5) : 0;
  }
  cudaLaunchConfig_t lc = {};
  lc.gridDim = dim3(q.tiles_x * q.tiles_y);
  lc.blockDim = dim3(v.nw * 32);
  lc.dynamicSmemBytes = resident_stage_bytes(v, q);
  lc.stream = s->stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;  // every tile on an SM at once: the tiles wait for each other
  attr[0].val.cooperative = 1;
  lc.attrs = attr;
  lc.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&lc, v.kernel[5], a);
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) {
    char m[256];
    snprintf(m, sizeof m, "projection_pack_kernel (resident): %s", cudaGetErrorString(e));
    return set_error(SAYAL_ECUDA, m);
  }
  s->launches++;
  float* t = s->u; s->u = s->u_buf; s->u_buf = t;
  t = s->v; s->v = s->v_buf; s->v_buf = t;
  s->parity ^= 1;
  return SAYAL_OK;
}

int run_passes(Sim* s, int variant, int T, int iterations, float d_t, bool with_forces = false, bool with_extrap = false,
               int ghost_depth = -1, int push_sides = 0, bool resident = false) {
  if (resident) return run_resident(s, variant, T, iterations, d_t, with_forces, with_extrap);
  const Variant& v = kVariants[variant];
  static thread_local PassPlan plans[kMaxPasses];
  const int passes = make_pass_plans(s->g, v, T, iterations, ghost_depth, push_sides, s->slab_halo, plans, kMaxPasses);
  if (passes < 0 || passes != (iterations + T - 1) / T)
    return set_error(SAYAL_EINVAL, push_sides ? "projection tile: temporal block too large for the tile or the slab's halo"
                                              : "projection tile: temporal block too large for the tile");
  const int signature = (iterations << 8) | passes;
  for (int pass = 0; pass < passes; pass++) {
    const PassPlan& pl = plans[pass];
    const Geometry& q = pl.q;
    const int it = pl.iterations;
    PackArgs a = {};
    a.g = s->g;
    a.row_lo = pl.row_lo;
    a.row_hi = pl.row_hi;
    a.write_lo = pl.write_lo;
    a.write_hi = pl.write_hi;
    a.u_in = s->u;
    a.v_in = s->v;
    a.u_out = s->u_buf;
    a.v_out = s->v_buf;
    a.flags = s->flags;
    a.o = s->ph.o;
    a.iters = it;
    a.halo_x = q.halo_x;
    a.halo_y = q.halo_y;
    a.stride_x = q.stride_x;
    a.stride_y = q.stride_y;
    a.p = s->ph.enable_pressure ? s->p : nullptr;
    a.density = s->ph.density;
    a.hf = (float)s->g.h;
    a.inv_dt = 1.0f / d_t;
    a.extrap_on = with_extrap && pass == passes - 1;
    a.force_on = with_forces && pass == 0;
    a.smoke = s->smoke;
    if (a.force_on) {
      const ForceArgs& f = s->fuse_args;
      const Phys& ph = s->ph;
      a.force_g = ph.g; a.force_dt = f.d_t; a.wt_speed = ph.wt_speed; a.wt_smoke = ph.wt_smoke;
      a.inlet_len = ph.wt_smoke_length; a.band_lo = f.band_lo; a.band_hi = f.band_hi;
      a.smoke_lo = f.smoke_lo; a.smoke_hi = f.smoke_hi; a.smoke_count = ph.wt_smoke_count;
      a.smoke_height = ph.wt_smoke_height; a.period = f.period; a.anchor = f.anchor;
    }
    if (push_sides) {
      PushArgs& pu = a.push;
      pu.on = push_sides;
      {
        static const int dbg = getenv("SAYAL_DEBUG_PUSH") ? atoi(getenv("SAYAL_DEBUG_PUSH")) : 0;
        pu.debug = dbg;
      }
      pu.pass_index = pass;
      pu.signature = signature;
      // the neighbour's output arrays of this pass: it ping-pongs in step with us, so its output is the array with
      // the creation index of ours (0..3 = u, v, u_buf, v_buf as allocated)
      const int iu = (int)(((char*)s->u_buf - (char*)s->vel_block) / (ptrdiff_t)s->vel_stride);
      const int iv = (int)(((char*)s->v_buf - (char*)s->vel_block) / (ptrdiff_t)s->vel_stride);
      for (int d = 0; d < 2; d++) {
        pu.src_lo[d] = pl.src_lo[d]; pu.src_hi[d] = pl.src_hi[d];
        pu.dst_row0[d] = s->peer_ghost_row0[d];
        pu.n_pushers[d] = pl.n_pushers[d];
        pu.peer_u[d] = s->peer_vel[d][iu];
        pu.peer_v[d] = s->peer_vel[d][iv];
        pu.peer_words[d] = s->link.peer_words[d];
        if ((push_sides & (1 << d)) && (pl.n_pushers[d] <= 0 || !pu.peer_u[d] || !pu.peer_v[d]))
          return set_error(SAYAL_EINVAL, "projection push: no tile produces the slab's edge rows (slab thinner than its halo?)");
      }
      pu.my_words = s->link.my_words;
      pu.ticket = s->link.push_ticket;
      pu.step_seq = s->link.step_seq;
      pu.link_error = s->link.link_error;
    }
    a.timeline = s->d_timeline;  // the last pass wins: profile single passes
    if (s->d_timeline && (size_t)q.tiles_x * q.tiles_y * 5 > s->timeline_cap) a.timeline = nullptr;
    s->timeline_tiles = a.timeline ? q.tiles_x * q.tiles_y : 0;
    a.tiles_x = q.tiles_x;
    int n_descs = 0;
    if (!push_sides && ghost_depth < 0 && pl.row_lo == 0 && pl.row_hi == s->g.local_rows)
      a.descs = tile_descs(s, variant, it, q, &n_descs);
    if (a.descs && a.timeline) s->timeline_tiles = n_descs;
    if (!a.descs) a.order = tile_order(s, variant, it, q, pl.row_lo, pl.row_hi, push_sides);
    int r = launch_pass(s, v, a, dim3(a.descs ? n_descs : q.tiles_x * q.tiles_y), s->stream);
    if (r != SAYAL_OK) return r;
    float* t = s->u; s->u = s->u_buf; s->u_buf = t;  // ping-pong: neighbouring tiles still read the old halo
    t = s->v; s->v = s->v_buf; s->v_buf = t;
    s->parity ^= 1;
  }
  if (push_sides) return launch_slab_push_wait(s, passes, signature);
  return SAYAL_OK;
}

}  // namespace

int tiled_max_temporal_block() { return kMaxT; }

// CUDA loads a kernel lazily at its first launch, and loading may wait for the device to drain.  A linked slab
// must never meet that wait in the middle of a step (its neighbour may be spinning for it), so every variant is
// loaded when the sim is created.
int tiled_preload() {
  {
    cudaFuncAttributes fa;
    if (cudaFuncGetAttributes(&fa, tile_cost_kernel) != cudaSuccess || cudaFuncGetAttributes(&fa, tile_order_kernel) != cudaSuccess ||
        cudaFuncGetAttributes(&fa, desc_cost_kernel) != cudaSuccess)
      return set_error(SAYAL_ECUDA, "preload: tile order kernels");
  }
  for (int v = 0; v < kNumVariants; v++)
    for (int m = 0; m < 6; m++) {
      cudaFuncAttributes fa;
      cudaError_t e = cudaFuncGetAttributes(&fa, kVariants[v].kernel[m]);
      if (e == cudaSuccess && m == 4)
        e = cudaFuncSetAttribute(kVariants[v].kernel[4], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pressure_smem(kVariants[v]));
      if (e == cudaSuccess && m == 5)
        e = cudaFuncSetAttribute(kVariants[v].kernel[5], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kResidentStageMax);
      if (e == cudaSuccess && m < 4) e = cudaFuncGetAttributes(&fa, kVariants[v].push_kernel[m]);
      if (e != cudaSuccess) return set_error(SAYAL_ECUDA, cudaGetErrorString(e));
    }
  return SAYAL_OK;
}

// Choose (tile variant, temporal block) for `iterations` SOR iterations on this grid: rank all candidates
// with the wave-quantisation model, then time the best few on the live arrays (state saved and restored;
// every candidate produces the same bits, so the choice never changes results).
int tiled_prepare(Sim* s, int iterations) {
  if (iterations <= 0) return SAYAL_OK;
  const int push = push_sides_of(s) ? 1 : 0;
  const int want_resident = resident_possible(s) && !push && resident_mailboxes(s) ? 1 : 0;
  for (int k = 0; k < s->n_plans; k++)
    if (s->plans[k].iterations == iterations && s->plans[k].push == push && s->plans[k].resident_ok == want_resident) {
      // (slab runs alternate between chunk sizes: keep every plan)
      s->plan_variant = s->plans[k].variant;
      s->plan_T = s->plans[k].T;
      s->plan_resident = s->plans[k].resident;
      return SAYAL_OK;
    }
  // the temporal block is given (option), dictated by the chain (push mode: every rank must split alike), or tuned
  int forced_T = 0;
  if (s->temporal_block > 0) forced_T = s->temporal_block < iterations ? s->temporal_block : iterations;
  else if (push) forced_T = tiled_push_temporal_block(iterations, s->slab_halo);
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, s->device);
  struct Cand { int variant, T; double cost; float ms; int resident; };
  Cand cands[2 * kNumVariants * kMaxT];
  int nc = 0;
  for (int v = 0; v < kNumVariants; v++)
    for (int T = 1; T <= kMaxT && T <= iterations; T++) {
      if (forced_T > 0 && T != forced_T) continue;
      if (s->force_variant >= 0 && v != s->force_variant) continue;
      {  // run_passes splits evenly: T and ceil(n / passes(T)) describe the same plan, keep the canonical one
        int passes = (iterations + T - 1) / T;
        if (forced_T <= 0 && (iterations + passes - 1) / passes != T) continue;
      }
      double c = model_cost(s->g, kVariants[v], T, iterations, sms);
      if (c < 1e29 && !(want_resident && s->resident == 2)) cands[nc++] = {v, T, c, 0.f, 0};  // resident = 2: resident plans only
      Geometry q;
      if (want_resident && resident_geometry(s, kVariants[v], T, &q)) {
        // one load / store of the tile, `blocks` sweep blocks, blocks - 1 ring exchanges of about a microsecond
        const int blocks = (iterations + T - 1) / T;
        cands[nc++] = {v, T, kVariants[v].ry * (0.45 + 0.13 * iterations) + 1.25 * (blocks - 1), 0.f, 1};
      }
    }
  if (nc == 0) return set_error(SAYAL_EINVAL, "projection tile: no feasible tile plan");
  for (int a = 0; a < nc; a++)  // selection sort by model cost
    for (int b = a + 1; b < nc; b++)
      if (cands[b].cost < cands[a].cost) { Cand t = cands[a]; cands[a] = cands[b]; cands[b] = t; }
  int best = 0;
  bool timed = false;
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  cudaStreamIsCapturing(s->stream, &cap);
  // time the model's best 8 candidates always, the rest while the tuning budget (about 0.25 s) lasts
  int ntime = nc;
  if (s->autotune && cap == cudaStreamCaptureStatusNone && ntime > 1) {
    size_t bytes = sizeof(float) * (size_t)s->g.pitch * s->g.local_rows;
    float *su = nullptr, *sv = nullptr, *spres = nullptr;  // the timed runs advance u, v (and accumulate into p)
    const bool save_p = s->ph.enable_pressure;
    if (cudaMalloc(&su, bytes) == cudaSuccess && cudaMalloc(&sv, bytes) == cudaSuccess &&
        (!save_p || cudaMalloc(&spres, bytes) == cudaSuccess)) {
      cudaEvent_t e0, e1;
      cudaEventCreate(&e0);
      cudaEventCreate(&e1);
      cudaMemcpyAsync(su, s->u, bytes, cudaMemcpyDeviceToDevice, s->stream);
      cudaMemcpyAsync(sv, s->v, bytes, cudaMemcpyDeviceToDevice, s->stream);
      if (save_p) cudaMemcpyAsync(spres, s->p, bytes, cudaMemcpyDeviceToDevice, s->stream);
      int64_t launches = s->launches;
      int parity = s->parity;
      const int orders_before = s->n_orders;  // the candidates' issue orders / tile lists are dropped after the tuning
      float *u0 = s->u, *v0 = s->v, *ub0 = s->u_buf, *vb0 = s->v_buf;
      // one timed run of a candidate (ms), or a negative value on failure
      auto time_once = [&](const Cand& c) -> float {
        cudaEventRecord(e0, s->stream);
        // (a linked slab's candidates are timed over the row windows its passes will sweep: s->tune_depth)
        int r = run_passes(s, c.variant, c.T, iterations, 1.0f, false, false, c.resident ? -1 : s->tune_depth, 0, c.resident != 0);
        cudaEventRecord(e1, s->stream);
        if (r != SAYAL_OK || cudaEventSynchronize(e1) != cudaSuccess) return -1.f;
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        return ms;
      };
      // An idle GPU takes a few milliseconds to reach its working clocks: candidates timed first would look slow.
      // Run the model's favourite until ~5 ms of device time have passed before any measurement counts.
      float spent_ms = 0.f;
      for (int k = 0; k < 200 && spent_ms < 5.f; k++) {
        float ms = time_once(cands[0]);
        if (ms < 0.f) break;
        spent_ms += ms;
      }
      spent_ms = 0.f;
      for (int c = 0; c < ntime; c++) {
        cands[c].ms = 1e30f;
        if (c >= 8 && spent_ms > 250.f) continue;
        for (int rep = 0; rep < 3; rep++) {  // first repetition warms the instruction cache
          float ms = time_once(cands[c]);
          if (ms < 0.f) { cands[c].ms = 1e30f; break; }
          spent_ms += ms;
          if (rep > 0 && ms < cands[c].ms) cands[c].ms = ms;
        }
      }
      // final: the three fastest again, interleaved, so that a drifting clock cannot favour one of them
      for (int a = 0; a < nc; a++)
        for (int b = a + 1; b < nc; b++)
          if (cands[b].ms < cands[a].ms) { Cand t = cands[a]; cands[a] = cands[b]; cands[b] = t; }
      const int finalists = nc < 3 ? nc : 3;
      for (int round = 0; round < 3; round++)
        for (int c = 0; c < finalists; c++) {
          if (cands[c].ms > 1e29f) continue;
          float ms = time_once(cands[c]);
          if (ms > 0.f && ms < cands[c].ms) cands[c].ms = ms;
        }
      best = 0;
      for (int c = 1; c < finalists; c++)
        if (cands[c].ms < cands[best].ms) best = c;
      // Reproducibility: finalists within 1.5 % of the fastest are a tie on this hardware (run-to-run noise is of that
      // order); among them the plan the model ranks best wins, so the same grid gets the same plan on every box.
      {
        const float limit = cands[best].ms * 1.015f;
        for (int c = 0; c < finalists; c++)
          if (cands[c].ms <= limit && cands[c].cost < cands[best].cost) best = c;
      }
      // A resident plan needs every SM for itself (cooperative launch) and its timing in isolation flatters it: it
      // must beat the best multi-pass finalist by 3 % to be chosen.
      if (cands[best].resident)
        for (int c = 0; c < finalists; c++)
          if (!cands[c].resident && cands[c].ms <= cands[best].ms * 1.03f && (cands[best].resident || cands[c].ms < cands[best].ms)) best = c;
      timed = true;
      // the cache of issue orders holds one entry per geometry a candidate swept (a slab's row windows: one per pass):
      // keep only what was there before — the chosen plan's entries are rebuilt right after (below / by the caller)
      cudaStreamSynchronize(s->stream);
      for (int k = orders_before; k < s->n_orders; k++)
        if (s->orders[k].order) cudaFree(s->orders[k].order);
      s->n_orders = orders_before;
      // restore state and bookkeeping: tuning is invisible
      s->u = u0; s->v = v0; s->u_buf = ub0; s->v_buf = vb0;
      s->parity = parity;
      s->launches = launches;
      cudaMemcpyAsync(s->u, su, bytes, cudaMemcpyDeviceToDevice, s->stream);
      cudaMemcpyAsync(s->v, sv, bytes, cudaMemcpyDeviceToDevice, s->stream);
      if (save_p) cudaMemcpyAsync(s->p, spres, bytes, cudaMemcpyDeviceToDevice, s->stream);
      cudaStreamSynchronize(s->stream);
      cudaEventDestroy(e0);
      cudaEventDestroy(e1);
    }
    if (su) cudaFree(su);
    if (sv) cudaFree(sv);
    if (spres) cudaFree(spres);
    cudaGetLastError();
  }
  if (!timed)  // a resident plan is only ever chosen on a measurement (its model is rough): untimed, take the best other
    for (int c = 0; c < nc; c++)
      if (!cands[c].resident) { best = c; break; }
  s->plan_variant = cands[best].variant;
  s->plan_T = cands[best].T;
  s->plan_resident = cands[best].resident;
  {  // what was considered, for the bench line and for anyone who wonders why this plan (sayal_plan_log)
    int at = snprintf(s->plan_log, sizeof s->plan_log, "n=%d %s%s: rows T[R = resident] model ms\n", iterations, timed ? "timed" : "model only",
                      push ? ", push mode" : "");
    for (int c = 0; c < nc && at < (int)sizeof s->plan_log - 48; c++) {
      if (timed && cands[c].ms > 1e29f) continue;
      at += snprintf(s->plan_log + at, sizeof s->plan_log - at, "%c %d %d%s %.1f %.4f\n", c == best ? '*' : ' ',
                     kVariants[cands[c].variant].ry, cands[c].T, cands[c].resident ? "R" : "", cands[c].cost,
                     timed ? cands[c].ms : 0.f);
    }
  }
  {  // the issue orders of the plan's pass geometries, now (they cannot be built while a graph is captured)
    const int passes = (iterations + s->plan_T - 1) / s->plan_T;
    for (int it = iterations / passes; !s->plan_resident && it <= (iterations + passes - 1) / passes; it++) {
      Geometry q;
      if (it > 0 && geometry(s->g, kVariants[s->plan_variant], it, &q)) {
        int n_descs = 0;
        if (!tile_descs(s, s->plan_variant, it, q, &n_descs)) tile_order(s, s->plan_variant, it, q, 0, s->g.local_rows);
      }
    }
  }
  if (s->n_plans == Sim::kMaxPlans) s->n_plans = 0;  // full: start over (never happens with <= 8 chunk sizes)
  s->plans[s->n_plans++] = {iterations, s->plan_variant, s->plan_T, push, s->plan_resident, want_resident};
  return SAYAL_OK;
}

// Build (outside graph capture) the tile issue orders a linked slab's projection call of `iterations` iterations
// will use: the same pass list as run_passes.  ghost_depth: exact ghost rows at the start of a deep-halo call;
// ignored in push mode (s->push_active).
int tiled_prepare_windows(Sim* s, int iterations, int ghost_depth) {
  // the tuner times its candidates over the windows the passes will really sweep (whole-array passes take the tile
  // list with its cut tiles, windows the regular grid: a plan that wins on one can lose on the other)
  s->tune_depth = push_sides_of(s) ? s->slab_halo : ghost_depth;
  int r = tiled_prepare(s, iterations);
  s->tune_depth = -1;
  const int sides = push_sides_of(s);
  if (r != SAYAL_OK || iterations <= 0 || (ghost_depth < 0 && !sides)) return r;
  static thread_local PassPlan plans[kMaxPasses];
  const int n = make_pass_plans(s->g, kVariants[s->plan_variant], s->plan_T, iterations, ghost_depth, sides, s->slab_halo, plans,
                                kMaxPasses);
  for (int k = 0; k < n; k++)
    tile_order(s, s->plan_variant, plans[k].iterations, plans[k].q, plans[k].row_lo, plans[k].row_hi, sides);
  return SAYAL_OK;
}

// Host-only (no CUDA call): the pass list of a projection on a grid described by numbers, for sayal_debug_pass_plans.
// ghost_depth >= 0: deep-halo slab; -1: every row; <= -2: push mode with halo = -ghost_depth (a side has a neighbour
// when the array holds rows beyond the owned ones on that side).
int tiled_debug_pass_plans(int pitch, int local_rows, int own_lo, int own_hi, int rows_per_warp, int T, int iterations,
                           int ghost_depth, int32_t* out, int capacity) {
  int variant = -1;
  for (int v = 0; v < kNumVariants; v++)
    if (kVariants[v].ry == rows_per_warp) variant = v;
  if (variant < 0 || pitch < 4 || (pitch & 3) || local_rows < 1 || T < 1 || T > kMaxT || own_lo < 0 || own_hi > local_rows)
    return -1;
  Grid g = {};
  g.W = pitch; g.H = local_rows; g.pitch = pitch; g.local_rows = local_rows; g.own_lo = own_lo; g.own_hi = own_hi; g.h = 1;
  static thread_local PassPlan plans[kMaxPasses];
  const int halo = ghost_depth <= -2 ? -ghost_depth : 0;
  const int sides = ghost_depth <= -2 ? ((own_lo > 0 ? 1 : 0) | (own_hi < local_rows ? 2 : 0)) : 0;
  const int n = make_pass_plans(g, kVariants[variant], T, iterations, ghost_depth, sides, halo, plans,
                                capacity < kMaxPasses ? capacity : kMaxPasses);
  for (int k = 0; k < n; k++) {
    const PassPlan& p = plans[k];
    const int32_t row[15] = {p.iterations, p.row_lo, p.row_hi, p.q.halo_x, p.q.halo_y, p.q.stride_x, p.q.stride_y,
                             p.q.tiles_x, p.q.tiles_y, TW, kVariants[variant].ry * kVariants[variant].nw,
                             p.write_lo, p.write_hi, p.n_pushers[0], p.n_pushers[1]};
    for (int c = 0; c < 15; c++) out[15 * k + c] = row[c];
  }
  return n;
}

// Diagnostics: the explicit tile list of the plan's passes of `it` iterations over the whole array, as 5 int32 per tile
// (X0, Y0, y_end, vy0, vy1), in issue order; returns the number of tiles, 0 if no such list has been built, -1 on error.
int tiled_debug_tile_list(Sim* s, int it, int32_t* out, int capacity) {
  for (int k = 0; k < s->n_orders; k++) {
    const Sim::TileOrder& o = s->orders[k];
    if (o.edge_first != -1 || o.variant != s->plan_variant || o.it != it || o.row_lo != 0 || o.row_hi != s->g.local_rows) continue;
    const int n = o.n_descs < capacity ? o.n_descs : capacity;
    std::vector<TileDesc> host(n);
    if (cudaStreamSynchronize(s->stream) != cudaSuccess ||
        cudaMemcpy(host.data(), o.order, n * sizeof(TileDesc), cudaMemcpyDeviceToHost) != cudaSuccess)
      return -1;
    for (int t = 0; t < n; t++) {
      const int32_t row[5] = {host[t].X0, host[t].Y0, host[t].y_end, host[t].vy0, host[t].vy1};
      for (int c = 0; c < 5; c++) out[5 * t + c] = row[c];
    }
    return o.n_descs;
  }
  return 0;
}

int launch_projection_tiled(Sim* s, int iterations, float d_t) {
  int r = tiled_prepare(s, iterations);
  if (r != SAYAL_OK) return r;
  const bool with_forces = s->fuse_pending != 0, with_extrap = s->fuse_extrap != 0;
  s->fuse_pending = 0;
  s->fuse_extrap = with_extrap ? 2 : 0;  // 2 = done: the caller skips the extrapolation kernel
  const int depth = s->proj_depth;
  s->proj_depth = -1;
  return run_passes(s, s->plan_variant, s->plan_T, iterations, d_t, with_forces, with_extrap, depth, push_sides_of(s),
                    s->plan_resident != 0);
}

// Push mode: the temporal block every rank of a chain uses for `iterations` iterations with `halo` ghost rows.  It
// must not depend on anything rank-local (neighbours pair their passes one to one), so it is a function of these
// two numbers only: the even split of `iterations` into passes of at most min(halo / 2, 8) iterations.
int tiled_push_temporal_block(int iterations, int halo) {
  int cap = halo / 2 < kMaxPushT ? halo / 2 : kMaxPushT;
  if (cap < 1 || iterations <= 0) return 0;
  const int passes = (iterations + cap - 1) / cap;
  return (iterations + passes - 1) / passes;
}

}  // namespace sayal
