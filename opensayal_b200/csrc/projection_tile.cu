// projection_tile.cu — the pressure projection of Fluid::update (/root/reference/src/fluid.cu:229-295) as
// a register-resident, temporally blocked red-black SOR kernel for sm_100a.
//
// What the reference does: 2n launches per step, each touching u, v, is_solid(int32) x5 and
// total_s(int32) for half the cells from global memory (fluid.cu:264-295).
//
// What this kernel does: one launch ("pass") advances a whole tile by T full iterations (2T half-sweeps)
// without touching global memory in between.
//   * A CTA owns a TW x TH tile = 128 columns x (NW warps * RY rows).  Lane l of warp w keeps cells
//     x = X0 + 4l .. 4l+3 of rows Y0 + w*RY .. +RY-1 in registers: u[RY][4], v[RY-1][4] and the four
//     1-byte cell flags of every row in one 32-bit register.  Loads/stores are float4 (512 B contiguous
//     per warp instruction).
//   * Inside a row the two cells of the active colour either own both u-faces (columns 0,2: no exchange)
//     or need the neighbour lane's first u (columns 1,3): one __shfl_down for u(i+1,j) and one __shfl_up
//     carrying the velocity correction back — the warp-shuffle edge exchange.
//   * Between vertically adjacent warps the only shared data is one row of v faces; it lives in shared
//     memory, split by colour so that every access is a conflict-free 8-byte vector, and both warps
//     read-modify-write disjoint columns of it in each half-sweep.  One __syncthreads per half-sweep.
//   * Rows in which every active cell is a fully open fluid cell (flag byte 0x4F, s = 4) take a
//     branch-free path with the constant 0.25; any other row takes the general masked path.  The choice
//     is warp-uniform and precomputed per row and colour.
//   * Tiles overlap by a halo of 2T cells (rounded up to 4 in x).  Errors from the missing neighbours
//     travel one cell per half-sweep, so after 2T half-sweeps everything at least 2T cells inside the
//     tile is exactly what the global sweep order produces; only that part is written, to the OTHER
//     buffer (ping-pong), because neighbouring tiles still read the old halo.
// The update is order-independent within a colour (cells of one colour share no face), so tiling does
// not change a single bit: tests/test_projection_tiled.py requires equality with the plain half-sweep
// kernel and with the CPU oracle.
//
// Roofline: HBM.  Algorithmic bytes are 17 B per cell per iteration (r/w u, v + 1 flag byte); one pass
// moves (1 + halo overhead) * 9 B in and 8 B out per cell for T iterations.
#include <cstdio>

#include "sayal_internal.h"

namespace sayal {

namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr int TW = 128;  // tile width: 32 lanes x 4 cells

__constant__ float c_inv_s_tile[8] = {0.0f, 1.0f, 0.5f, 1.0f / 3.0f, 0.25f, 0.f, 0.f, 0.f};

struct TileArgs {
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
};

// One row, fully open cells only.  CASE 0: columns 0 and 2 are the active colour; CASE 1: columns 1, 3.
template <int CASE>
__device__ __forceinline__ void row_fast(float (&u)[4], float (&vb)[4], float (&vt)[4], float o, int lane) {
  if (CASE == 0) {
    float d0 = __fsub_rn(__fadd_rn(__fsub_rn(u[1], u[0]), vt[0]), vb[0]);
    float d2 = __fsub_rn(__fadd_rn(__fsub_rn(u[3], u[2]), vt[2]), vb[2]);
    float e0 = __fmul_rn(o, __fmul_rn(d0, 0.25f));
    float e2 = __fmul_rn(o, __fmul_rn(d2, 0.25f));
    u[0] = __fadd_rn(u[0], e0);
    u[1] = __fsub_rn(u[1], e0);
    vb[0] = __fadd_rn(vb[0], e0);
    vt[0] = __fsub_rn(vt[0], e0);
    u[2] = __fadd_rn(u[2], e2);
    u[3] = __fsub_rn(u[3], e2);
    vb[2] = __fadd_rn(vb[2], e2);
    vt[2] = __fsub_rn(vt[2], e2);
  } else {
    float ur = __shfl_down_sync(FULL, u[0], 1);  // u(i+1, j) of column 3 lives in the next lane
    float d1 = __fsub_rn(__fadd_rn(__fsub_rn(u[2], u[1]), vt[1]), vb[1]);
    float d3 = __fsub_rn(__fadd_rn(__fsub_rn(ur, u[3]), vt[3]), vb[3]);
    float e1 = __fmul_rn(o, __fmul_rn(d1, 0.25f));
    float e3 = __fmul_rn(o, __fmul_rn(d3, 0.25f));
    u[1] = __fadd_rn(u[1], e1);
    u[2] = __fsub_rn(u[2], e1);
    vb[1] = __fadd_rn(vb[1], e1);
    vt[1] = __fsub_rn(vt[1], e1);
    u[3] = __fadd_rn(u[3], e3);
    vb[3] = __fadd_rn(vb[3], e3);
    vt[3] = __fsub_rn(vt[3], e3);
    float el = __shfl_up_sync(FULL, e3, 1);  // the correction of that shared face travels back
    if (lane == 0) el = 0.f;
    u[0] = __fsub_rn(u[0], el);
  }
}

// One row, any mix of solid / border / partly enclosed cells (fluid.cu:229-262 with the flag byte).
template <int CASE>
__device__ __forceinline__ void row_general(float (&u)[4], float (&vb)[4], float (&vt)[4], unsigned fl, float o,
                                            int lane) {
  if (CASE == 0) {
#pragma unroll
    for (int k = 0; k < 4; k += 2) {
      unsigned f = (fl >> (8 * k)) & 0xffu;
      float d = __fsub_rn(__fadd_rn(__fsub_rn(u[k + 1], u[k]), vt[k]), vb[k]);
      float e = __fmul_rn(o, __fmul_rn(d, c_inv_s_tile[(f >> 4) & 7u]));
      if (f & FL_L) u[k] = __fadd_rn(u[k], e);
      if (f & FL_R) u[k + 1] = __fsub_rn(u[k + 1], e);
      if (f & FL_B) vb[k] = __fadd_rn(vb[k], e);
      if (f & FL_T) vt[k] = __fsub_rn(vt[k], e);
    }
  } else {
    float ur = __shfl_down_sync(FULL, u[0], 1);
    unsigned f1 = (fl >> 8) & 0xffu, f3 = (fl >> 24) & 0xffu;
    float d1 = __fsub_rn(__fadd_rn(__fsub_rn(u[2], u[1]), vt[1]), vb[1]);
    float d3 = __fsub_rn(__fadd_rn(__fsub_rn(ur, u[3]), vt[3]), vb[3]);
    float e1 = __fmul_rn(o, __fmul_rn(d1, c_inv_s_tile[(f1 >> 4) & 7u]));
    float e3 = __fmul_rn(o, __fmul_rn(d3, c_inv_s_tile[(f3 >> 4) & 7u]));
    if (f1 & FL_L) u[1] = __fadd_rn(u[1], e1);
    if (f1 & FL_R) u[2] = __fsub_rn(u[2], e1);
    if (f1 & FL_B) vb[1] = __fadd_rn(vb[1], e1);
    if (f1 & FL_T) vt[1] = __fsub_rn(vt[1], e1);
    if (f3 & FL_L) u[3] = __fadd_rn(u[3], e3);
    if (f3 & FL_B) vb[3] = __fadd_rn(vb[3], e3);
    if (f3 & FL_T) vt[3] = __fsub_rn(vt[3], e3);
    float send = (f3 & FL_R) ? e3 : 0.f;  // x - (+0) == x exactly, so an unmasked subtract is a no-op
    float el = __shfl_up_sync(FULL, send, 1);
    if (lane == 0) el = 0.f;
    u[0] = __fsub_rn(u[0], el);
  }
}

template <int CASE>
__device__ __forceinline__ void row_update(float (&u)[4], float (&vb)[4], float (&vt)[4], unsigned fl, bool fast,
                                           float o, int lane) {
  if (fast) row_fast<CASE>(u, vb, vt, o, lane);
  else row_general<CASE>(u, vb, vt, fl, o, lane);
}

// One half-sweep over the RY rows of this warp.  Q0 = case of row 0; the case alternates with the row.
// sv_top / sv_bot point at this lane's float2 slots of the shared boundary rows above row 0 and below row
// RY-1, already offset to the colour that is active in that row.
template <int RY, int Q0>
__device__ __forceinline__ void half_sweep(float (&u)[RY][4], float (&v)[RY - 1][4], const unsigned (&fl)[RY],
                                           unsigned fast_a, unsigned fast_b, float o, int lane, float* sv_top,
                                           float* sv_bot) {
  constexpr int QL = Q0 ^ ((RY - 1) & 1);  // case of the last row
  // ---- row 0: its top faces are the shared row above this warp
  {
    float vt[4];
    float2 t = *reinterpret_cast<float2*>(sv_top + 64 * Q0);
    vt[Q0] = t.x;
    vt[Q0 + 2] = t.y;
    vt[1 - Q0] = 0.f;
    vt[3 - Q0] = 0.f;
    bool fast = ((Q0 ? fast_b : fast_a) & 1u) != 0;
    row_update<Q0>(u[0], v[0], vt, fl[0], fast, o, lane);
    *reinterpret_cast<float2*>(sv_top + 64 * Q0) = make_float2(vt[Q0], vt[Q0 + 2]);
  }
  // ---- interior rows: both face rows are registers
#pragma unroll
  for (int r = 1; r < RY - 1; r++) {
    if (((r & 1) ^ Q0) == 0) {
      bool fast = ((fast_a >> r) & 1u) != 0;
      row_update<0>(u[r], v[r], v[r - 1], fl[r], fast, o, lane);
    } else {
      bool fast = ((fast_b >> r) & 1u) != 0;
      row_update<1>(u[r], v[r], v[r - 1], fl[r], fast, o, lane);
    }
  }
  // ---- last row: its bottom faces are the shared row below this warp
  {
    float vb[4];
    float2 t = *reinterpret_cast<float2*>(sv_bot + 64 * QL);
    vb[QL] = t.x;
    vb[QL + 2] = t.y;
    vb[1 - QL] = 0.f;
    vb[3 - QL] = 0.f;
    bool fast = (((QL ? fast_b : fast_a) >> (RY - 1)) & 1u) != 0;
    row_update<QL>(u[RY - 1], vb, v[RY - 2], fl[RY - 1], fast, o, lane);
    *reinterpret_cast<float2*>(sv_bot + 64 * QL) = make_float2(vb[QL], vb[QL + 2]);
  }
}

// One half-sweep when every row of this warp is fully open for both colours: no flags, no branches, and
// the compiler is free to interleave the independent rows.  `skip0` is set for the tile's first warp, whose
// row 0 is the carrier row.  Lane 0 does not zero the correction it receives from "lane -1": a warp can only
// be all-open away from the left wall (X0 > 0), where column X0 is halo and is never written back.
template <int RY, int Q0>
__device__ __forceinline__ void half_sweep_open(float (&u)[RY][4], float (&v)[RY - 1][4], float o, bool skip0,
                                                float* sv_top, float* sv_bot) {
  constexpr int QL = Q0 ^ ((RY - 1) & 1);
  if (!skip0) {
    float vt[4];
    float2 t = *reinterpret_cast<float2*>(sv_top + 64 * Q0);
    vt[Q0] = t.x; vt[Q0 + 2] = t.y; vt[1 - Q0] = 0.f; vt[3 - Q0] = 0.f;
    row_fast<Q0>(u[0], v[0], vt, o, 1);
    *reinterpret_cast<float2*>(sv_top + 64 * Q0) = make_float2(vt[Q0], vt[Q0 + 2]);
  }
#pragma unroll
  for (int r = 1; r < RY - 1; r++) {
    if (((r & 1) ^ Q0) == 0) row_fast<0>(u[r], v[r], v[r - 1], o, 1);
    else row_fast<1>(u[r], v[r], v[r - 1], o, 1);
  }
  {
    float vb[4];
    float2 t = *reinterpret_cast<float2*>(sv_bot + 64 * QL);
    vb[QL] = t.x; vb[QL + 2] = t.y; vb[1 - QL] = 0.f; vb[3 - QL] = 0.f;
    row_fast<QL>(u[RY - 1], vb, v[RY - 2], o, 1);
    *reinterpret_cast<float2*>(sv_bot + 64 * QL) = make_float2(vb[QL], vb[QL + 2]);
  }
}

template <int RY, int NW>
__global__ void __launch_bounds__(NW * 32, 1) projection_tile_kernel(TileArgs a) {
  constexpr int TH = RY * NW;
  // shared v rows: sv[0] is a dummy above warp 0, sv[w+1] is the last row of warp w.
  // Layout per row: [colour 0: 64 floats][colour 1: 64 floats]; lane l owns float2 at 2l of each.
  __shared__ __align__(16) float sv[NW + 1][128];

  const Grid& g = a.g;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int X0 = blockIdx.x * a.stride_x, Y0 = blockIdx.y * a.stride_y;
  const int x = X0 + 4 * lane;
  const int lr0 = Y0 + w * RY;

  float u[RY][4], v[RY - 1][4];
  unsigned fl[RY];
  float vlast[4];

  const bool col_ok = x < g.pitch;  // pitch is a multiple of 4: a lane's four columns are in or out together
#pragma unroll
  for (int r = 0; r < RY; r++) {
    int lr = lr0 + r;
    float4 uu = make_float4(0.f, 0.f, 0.f, 0.f), vv = uu;
    unsigned f = 0;
    if (col_ok && lr < g.local_rows) {
      size_t k = (size_t)lr * g.pitch + x;
      uu = *reinterpret_cast<const float4*>(a.u_in + k);
      vv = *reinterpret_cast<const float4*>(a.v_in + k);
      f = *reinterpret_cast<const unsigned*>(a.flags + k);
    }
    u[r][0] = uu.x; u[r][1] = uu.y; u[r][2] = uu.z; u[r][3] = uu.w;
    if (r < RY - 1) {
      v[r][0] = vv.x; v[r][1] = vv.y; v[r][2] = vv.z; v[r][3] = vv.w;
    } else {
      vlast[0] = vv.x; vlast[1] = vv.y; vlast[2] = vv.z; vlast[3] = vv.w;
    }
    fl[r] = f;
  }
  // Carriers: the tile's first row has no row above and its last column no column to the right; those
  // cells only lend their faces to their neighbours.
  if (w == 0) fl[0] = 0;
  if (lane == 31) {
#pragma unroll
    for (int r = 0; r < RY; r++) fl[r] &= 0x00ffffffu;
  }
  // Per row: is every active cell of colour-case A (columns 0,2) / B (columns 1,3) fully open?
  unsigned fast_a = 0, fast_b = 0;
  {
    const unsigned open_a = (unsigned)FL_OPEN | ((unsigned)FL_OPEN << 16);
    const unsigned open_b = ((unsigned)FL_OPEN << 8) | ((unsigned)FL_OPEN << 24);
    // The carrier column may compute garbage in the open path — but only where it is halo, i.e. the tile has a
    // right neighbour.  A tile that ends exactly on the last column keeps the test (that column is a border
    // cell, so its rows take the general path).
    const unsigned mask_b = (lane == 31 && X0 + TW < g.pitch) ? 0x0000ff00u : 0xff00ff00u;
#pragma unroll
    for (int r = 0; r < RY; r++) {
      bool oka = (fl[r] & 0x00ff00ffu) == open_a;
      bool okb = (fl[r] & mask_b) == (open_b & mask_b);
      if (__all_sync(FULL, oka)) fast_a |= 1u << r;
      if (__all_sync(FULL, okb)) fast_b |= 1u << r;
    }
    if (w == 0) {  // carrier row: never fast (its flags were cleared above, this is for clarity)
      fast_a &= ~1u;
      fast_b &= ~1u;
    }
  }

  // publish the last row's v into the shared boundary row, colour-split
  float* sv_top = &sv[w][2 * lane];
  float* sv_bot = &sv[w + 1][2 * lane];
  *reinterpret_cast<float2*>(sv_bot) = make_float2(vlast[0], vlast[2]);
  *reinterpret_cast<float2*>(sv_bot + 64) = make_float2(vlast[1], vlast[3]);
  if (w == 0) {
    *reinterpret_cast<float2*>(sv_top) = make_float2(0.f, 0.f);
    *reinterpret_cast<float2*>(sv_top + 64) = make_float2(0.f, 0.f);
  }
  __syncthreads();

  // Colour of column 0 in row 0 of this warp: cell (i, j) belongs to half-sweep `c` iff (i + j + c) is even
  // (fluid.cu:266, 275).  x is a multiple of 4, so the case of row r in half-sweep c is (j0 - r + c) & 1.
  const int j0 = g.H - 1 - (g.row_base + lr0);
  const int q = j0 & 1;  // case of row 0 in the first (even) half-sweep: 0 -> columns 0,2
  static_assert(RY % 2 == 0, "rows per warp must be even: q must be uniform across the CTA's warps");
  const unsigned all_rows = (1u << RY) - 1u;
  const unsigned need = w == 0 ? (all_rows & ~1u) : all_rows;
  const bool all_open = ((fast_a & fast_b) & need) == need;  // warp-uniform
  const bool skip0 = w == 0;
  if (all_open) {
    for (int it = 0; it < a.iters; it++) {
      if (q == 0) {
        half_sweep_open<RY, 0>(u, v, a.o, skip0, sv_top, sv_bot);
        __syncthreads();
        half_sweep_open<RY, 1>(u, v, a.o, skip0, sv_top, sv_bot);
        __syncthreads();
      } else {
        half_sweep_open<RY, 1>(u, v, a.o, skip0, sv_top, sv_bot);
        __syncthreads();
        half_sweep_open<RY, 0>(u, v, a.o, skip0, sv_top, sv_bot);
        __syncthreads();
      }
    }
  } else {
    for (int it = 0; it < a.iters; it++) {
      // The flags never change, and the compiler knows it: left alone it hoists every per-cell mask and
      // reciprocal out of this loop and spills them.  Make the flag registers opaque once per iteration.
#pragma unroll
      for (int r = 0; r < RY; r++) asm volatile("" : "+r"(fl[r]));
      asm volatile("" : "+r"(fast_a), "+r"(fast_b));
      if (q == 0) {
        half_sweep<RY, 0>(u, v, fl, fast_a, fast_b, a.o, lane, sv_top, sv_bot);
        __syncthreads();
        half_sweep<RY, 1>(u, v, fl, fast_a, fast_b, a.o, lane, sv_top, sv_bot);
        __syncthreads();
      } else {
        half_sweep<RY, 1>(u, v, fl, fast_a, fast_b, a.o, lane, sv_top, sv_bot);
        __syncthreads();
        half_sweep<RY, 0>(u, v, fl, fast_a, fast_b, a.o, lane, sv_top, sv_bot);
        __syncthreads();
      }
    }
  }

  {
    float2 e = *reinterpret_cast<float2*>(sv_bot), o2 = *reinterpret_cast<float2*>(sv_bot + 64);
    vlast[0] = e.x; vlast[2] = e.y; vlast[1] = o2.x; vlast[3] = o2.y;
  }

  // write the part of the tile that is exact: everything >= halo away from an edge that has a neighbour
  const int vx0 = X0 == 0 ? 0 : X0 + a.halo_x;
  const int vx1 = X0 + TW >= g.pitch ? g.pitch : X0 + TW - a.halo_x;
  const int vy0 = Y0 == 0 ? 0 : Y0 + a.halo_y;
  const int vy1 = Y0 + TH >= g.local_rows ? g.local_rows : Y0 + TH - a.halo_y;
  if (x >= vx0 && x < vx1) {
#pragma unroll
    for (int r = 0; r < RY; r++) {
      int lr = lr0 + r;
      if (lr >= vy0 && lr < vy1) {
        size_t k = (size_t)lr * g.pitch + x;
        *reinterpret_cast<float4*>(a.u_out + k) = make_float4(u[r][0], u[r][1], u[r][2], u[r][3]);
        if (r < RY - 1)
          *reinterpret_cast<float4*>(a.v_out + k) = make_float4(v[r][0], v[r][1], v[r][2], v[r][3]);
        else
          *reinterpret_cast<float4*>(a.v_out + k) = make_float4(vlast[0], vlast[1], vlast[2], vlast[3]);
      }
    }
  }
}

struct Variant {
  int ry, nw;
  void (*kernel)(TileArgs);
};
const Variant kVariants[] = {
    {8, 16, projection_tile_kernel<8, 16>},
    {10, 16, projection_tile_kernel<10, 16>},
    {12, 16, projection_tile_kernel<12, 16>},
};
constexpr int kNumVariants = sizeof(kVariants) / sizeof(kVariants[0]);
constexpr int kMaxT = 16;

int tiles_for(int extent, int tile, int stride) {
  if (extent <= tile) return 1;
  return (extent - tile + stride - 1) / stride + 1;
}

struct Geometry {
  int halo_x, halo_y, stride_x, stride_y, tiles_x, tiles_y;
};

bool geometry(const Grid& g, const Variant& v, int T, Geometry* out) {
  int th = v.ry * v.nw;
  out->halo_y = 2 * T;
  out->halo_x = (2 * T + 3) & ~3;
  out->stride_x = TW - 2 * out->halo_x;
  out->stride_y = th - 2 * out->halo_y;
  if (out->stride_x < TW / 4 || out->stride_y < th / 4) return false;  // keep at least a quarter of the tile useful
  out->tiles_x = tiles_for(g.pitch, TW, out->stride_x);
  out->tiles_y = tiles_for(g.local_rows, th, out->stride_y);
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

int run_passes(Sim* s, int variant, int T, int iterations) {
  const Variant& v = kVariants[variant];
  int done = 0;
  while (done < iterations) {
    int it = iterations - done < T ? iterations - done : T;
    Geometry q;
    if (!geometry(s->g, v, it, &q)) return set_error(SAYAL_EINVAL, "projection tile: temporal block too large for the tile");
    TileArgs a;
    a.g = s->g;
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
    v.kernel<<<dim3(q.tiles_x, q.tiles_y), v.nw * 32, 0, s->stream>>>(a);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
      char m[256];
      snprintf(m, sizeof m, "projection_tile_kernel: %s", cudaGetErrorString(e));
      return set_error(SAYAL_ECUDA, m);
    }
    s->launches++;
    float* t = s->u; s->u = s->u_buf; s->u_buf = t;  // ping-pong: neighbouring tiles still read the old halo
    t = s->v; s->v = s->v_buf; s->v_buf = t;
    s->parity ^= 1;
    done += it;
  }
  return SAYAL_OK;
}

}  // namespace

int tiled_max_temporal_block() { return kMaxT; }

// Choose (tile variant, temporal block) for `iterations` SOR iterations on this grid: rank all candidates
// with the wave-quantisation model, then time the best few on the live arrays (state saved and restored;
// every candidate produces the same bits, so the choice never changes results).
int tiled_prepare(Sim* s, int iterations) {
  if (s->ph.enable_pressure || iterations <= 0) return SAYAL_OK;
  if (s->plan_iterations == iterations && s->plan_variant >= 0) return SAYAL_OK;
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, s->device);
  struct Cand { int variant, T; double cost; float ms; };
  Cand cands[kNumVariants * kMaxT];
  int nc = 0;
  for (int v = 0; v < kNumVariants; v++)
    for (int T = 1; T <= kMaxT && T <= iterations; T++) {
      if (s->temporal_block > 0 && T != (s->temporal_block < iterations ? s->temporal_block : iterations)) continue;
      if (s->force_variant >= 0 && v != s->force_variant) continue;
      double c = model_cost(s->g, kVariants[v], T, iterations, sms);
      if (c < 1e29) cands[nc++] = {v, T, c, 0.f};
    }
  if (nc == 0) return set_error(SAYAL_EINVAL, "projection tile: no feasible tile plan");
  for (int a = 0; a < nc; a++)  // selection sort by model cost
    for (int b = a + 1; b < nc; b++)
      if (cands[b].cost < cands[a].cost) { Cand t = cands[a]; cands[a] = cands[b]; cands[b] = t; }
  int best = 0;
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  cudaStreamIsCapturing(s->stream, &cap);
  int ntime = nc < 6 ? nc : 6;
  if (s->autotune && cap == cudaStreamCaptureStatusNone && ntime > 1) {
    size_t bytes = sizeof(float) * (size_t)s->g.pitch * s->g.local_rows;
    float *su = nullptr, *sv = nullptr;
    if (cudaMalloc(&su, bytes) == cudaSuccess && cudaMalloc(&sv, bytes) == cudaSuccess) {
      cudaEvent_t e0, e1;
      cudaEventCreate(&e0);
      cudaEventCreate(&e1);
      cudaMemcpyAsync(su, s->u, bytes, cudaMemcpyDeviceToDevice, s->stream);
      cudaMemcpyAsync(sv, s->v, bytes, cudaMemcpyDeviceToDevice, s->stream);
      int64_t launches = s->launches;
      int parity = s->parity;
      float *u0 = s->u, *v0 = s->v, *ub0 = s->u_buf, *vb0 = s->v_buf;
      float best_ms = 1e30f;
      for (int c = 0; c < ntime; c++) {
        float ms_min = 1e30f;
        for (int rep = 0; rep < 3; rep++) {  // first repetition warms the instruction cache
          cudaEventRecord(e0, s->stream);
          int r = run_passes(s, cands[c].variant, cands[c].T, iterations);
          cudaEventRecord(e1, s->stream);
          if (r != SAYAL_OK || cudaEventSynchronize(e1) != cudaSuccess) { ms_min = 1e30f; break; }
          float ms = 0.f;
          cudaEventElapsedTime(&ms, e0, e1);
          if (rep > 0 && ms < ms_min) ms_min = ms;
        }
        cands[c].ms = ms_min;
        if (ms_min < best_ms) { best_ms = ms_min; best = c; }
      }
      // restore state and bookkeeping: tuning is invisible
      s->u = u0; s->v = v0; s->u_buf = ub0; s->v_buf = vb0;
      s->parity = parity;
      s->launches = launches;
      cudaMemcpyAsync(s->u, su, bytes, cudaMemcpyDeviceToDevice, s->stream);
      cudaMemcpyAsync(s->v, sv, bytes, cudaMemcpyDeviceToDevice, s->stream);
      cudaStreamSynchronize(s->stream);
      cudaEventDestroy(e0);
      cudaEventDestroy(e1);
    }
    if (su) cudaFree(su);
    if (sv) cudaFree(sv);
    cudaGetLastError();
  }
  s->plan_variant = cands[best].variant;
  s->plan_T = cands[best].T;
  s->plan_iterations = iterations;
  return SAYAL_OK;
}

int launch_projection_tiled(Sim* s, int iterations, float d_t) {
  if (s->ph.enable_pressure) return launch_projection_plain(s, iterations, d_t);  // pressure accumulates per cell: plain path
  int r = tiled_prepare(s, iterations);
  if (r != SAYAL_OK) return r;
  return run_passes(s, s->plan_variant, s->plan_T, iterations);
}

}  // namespace sayal
