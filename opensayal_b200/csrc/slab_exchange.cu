// slab_exchange.cu — ghost-row exchange between y-slabs over NVLink peer memory (no reference equivalent:
// OpenSayal is single-GPU).
//
// One process per GPU.  Each sim owns ONE device block that neighbours may write: a flag word and a
// double-buffered receive area per side.  The block's CUDA IPC handle is handed to the two neighbour processes
// once (sayal_slab_ipc_export / sayal_slab_ipc_connect; slabs of one process link with sayal_slab_connect_local).
// After that an exchange is ONE kernel on the sim's own stream and nothing else — no host synchronisation, no
// collective library, capturable in the step's CUDA graph:
//
//   push   reads this slab's `halo` edge rows of the selected fields and stores them straight into the
//          neighbour's receive area (posted NVLink writes), fences, and the last CTA to finish publishes the
//          exchange's sequence number in the neighbour's flag word (st.release.sys);
//   wait   spins (ld.acquire.sys on a LOCAL word) until the neighbour's push with the same sequence number has
//          landed, then copies the receive area into the ghost rows.
//
// Sequence numbers live in device memory and are advanced by the kernels themselves, so a captured graph can be
// replayed.  The receive area is double-buffered by sequence parity; a buffer is reused two exchanges later,
// and by then this rank has consumed the neighbour's next push, which the neighbour issued (stream order) after
// it had unpacked the earlier one — so no acknowledgement is needed.  A spin that lasts longer than ~2 s gives
// up and raises the sim's `link_error` word instead of hanging the GPU.
#include <cstdio>

#include "sayal_internal.h"

namespace sayal {

namespace {

constexpr int XTHREADS = 1024;
constexpr long long kSpinLimitNs = 2000000000ll;

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

struct Fields3 {
  float* f[3];
  int rows[3];  // edge rows of each field that travel: u, v to the full halo depth; smoke only as deep as the
                // advection gathers (advect_margin + 2) — the projection never reads it
  int n;
};

// `rows[f]` rows of every selected field <-> packed buffer [field][row][W].  The rows next to the slab edge on
// `side` are the ones that travel: on the low side the range starts at `edge`, on the high side it ends there.
// Four independent 16-byte loads are in flight per thread before the first store: the copy is latency bound
// (local L2 read, then a posted NVLink write), not bandwidth bound.
template <bool UNPACK>
__device__ __forceinline__ void copy_rows(const Grid& g, const Fields3& fs, int edge, bool ends_at_edge, float* buf, int block,
                                          int nblocks) {
  const int W = g.W;
  size_t base = 0;
  for (int f = 0; f < fs.n; f++) {
    const int nrows = fs.rows[f];
    const int row0 = ends_at_edge ? edge - nrows : edge;
    const size_t per = (size_t)nrows * W;
    if ((W & 3) == 0) {  // 16-byte path: pitch and W are multiples of 4, rows are 16-byte aligned on both sides
      const int w4 = W >> 2;
      const size_t per4 = (size_t)nrows * w4, step = (size_t)nblocks * XTHREADS;
      float4* b4 = reinterpret_cast<float4*>(buf + base);
      float* field = fs.f[f] + (size_t)row0 * g.pitch;
      auto cell = [&](size_t t) -> float4* {
        int row = (int)(t / w4), c = (int)(t - (size_t)row * w4);
        return reinterpret_cast<float4*>(field + (size_t)row * g.pitch) + c;
      };
      size_t t = (size_t)block * XTHREADS + threadIdx.x;
      for (; t + 3 * step < per4; t += 4 * step) {
        float4 v0, v1, v2, v3;
        if (UNPACK) {  // peer-written: read through L2, never a stale L1 line
          v0 = __ldcg(b4 + t); v1 = __ldcg(b4 + t + step); v2 = __ldcg(b4 + t + 2 * step); v3 = __ldcg(b4 + t + 3 * step);
          *cell(t) = v0; *cell(t + step) = v1; *cell(t + 2 * step) = v2; *cell(t + 3 * step) = v3;
        } else {
          v0 = *cell(t); v1 = *cell(t + step); v2 = *cell(t + 2 * step); v3 = *cell(t + 3 * step);
          b4[t] = v0; b4[t + step] = v1; b4[t + 2 * step] = v2; b4[t + 3 * step] = v3;
        }
      }
      for (; t < per4; t += step) {
        if (UNPACK) *cell(t) = __ldcg(b4 + t);
        else b4[t] = *cell(t);
      }
    } else {
      for (size_t t = (size_t)block * XTHREADS + threadIdx.x; t < per; t += (size_t)nblocks * XTHREADS) {
        int row = (int)(t / W), c = (int)(t - (size_t)row * W);
        float* p = fs.f[f] + (size_t)(row0 + row) * g.pitch + c;
        if (UNPACK) *p = __ldcg(buf + base + t);
        else buf[base + t] = *p;
      }
    }
    base += per;
  }
}

// One exchange, one kernel.  grid (blocks, 2): blockIdx.y = side (0 = low memory rows, 1 = high memory rows).
//   push: every CTA stores its share of this slab's edge rows into the neighbour's receive area; one thread per
//         CTA then fences at system scope (the CTA barrier orders the other threads' stores before it) and takes a
//         ticket; the last CTA publishes the sequence number in the neighbour's flag word.
//   wait: one thread per CTA spins on the LOCAL flag word until the neighbour's push of the same exchange has
//         landed, then the CTA copies its share of the receive area into the ghost rows.
__global__ void __launch_bounds__(XTHREADS) slab_exchange_kernel(Grid g, int halo, Fields3 fs, SlabLinkDev d) {
  const int side = blockIdx.y;
  if (!d.peer_recv[side]) return;
  const unsigned seq = d.send_seq[side];  // == recv_seq: advanced below by the last CTA, after every CTA has read it
  float* dst = d.peer_recv[side] + (size_t)(seq & 1u) * d.stage_elems;
  copy_rows<false>(g, fs, side == 0 ? g.own_lo : g.own_hi, side == 1, dst, blockIdx.x, gridDim.x);
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();
    unsigned t = atomicAdd(&d.ticket[side], 1u);
    if (t == gridDim.x - 1) {
      d.ticket[side] = 0;
      __threadfence_system();
      st_release_sys(d.peer_flag[side], seq + 1);
    }
    const long long t0 = now_ns();
    while ((int)(ld_acquire_sys(d.my_flag + side) - (seq + 1)) < 0) {
      if (now_ns() - t0 > kSpinLimitNs) {  // neighbour gone: do not hang the GPU
        atomicExch(d.link_error, 1);
        break;
      }
    }
  }
  __syncthreads();
  const float* src = d.my_recv[side] + (size_t)(seq & 1u) * d.stage_elems;
  copy_rows<true>(g, fs, side == 0 ? g.own_lo : g.own_hi, side == 0, const_cast<float*>(src), blockIdx.x, gridDim.x);
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned t = atomicAdd(&d.ticket[2 + side], 1u);
    if (t == gridDim.x - 1) {
      d.ticket[2 + side] = 0;
      d.send_seq[side] = seq + 1;
    }
  }
}

size_t stage_elems(const Sim* s) { return (size_t)s->slab_halo * s->g.W * 3; }
size_t block_bytes(const Sim* s) { return 256 + 4 * stage_elems(s) * sizeof(float); }  // flags + 2 sides x 2 parities

}  // namespace

// Allocate the neighbour-writable block and the private counters of a slab sim.
int slab_link_alloc(Sim* s) {
  if (s->link_block) return SAYAL_OK;
  if (s->slab_halo <= 0) return set_error(SAYAL_EINVAL, "slab link: the sim was created without ghost rows");
  size_t bytes = block_bytes(s);
  cudaError_t e = cudaMalloc(&s->link_block, bytes);
  if (e != cudaSuccess) return set_error(SAYAL_ENOMEM, cudaGetErrorString(e));
  cudaMemsetAsync(s->link_block, 0, bytes, s->stream);
  e = cudaMalloc(&s->link_counters, 16 * sizeof(unsigned));
  if (e != cudaSuccess) return set_error(SAYAL_ENOMEM, cudaGetErrorString(e));
  cudaMemsetAsync(s->link_counters, 0, 16 * sizeof(unsigned), s->stream);
  e = cudaStreamSynchronize(s->stream);
  if (e != cudaSuccess) return set_error(SAYAL_ECUDA, cudaGetErrorString(e));
  SlabLinkDev& d = s->link;
  d.stage_elems = stage_elems(s);
  d.my_flag = reinterpret_cast<unsigned*>(s->link_block);
  float* recv = reinterpret_cast<float*>(reinterpret_cast<char*>(s->link_block) + 256);
  d.my_recv[0] = recv;
  d.my_recv[1] = recv + 2 * d.stage_elems;
  d.send_seq = s->link_counters;
  d.recv_seq = s->link_counters + 2;
  d.ticket = s->link_counters + 4;
  d.link_error = reinterpret_cast<int*>(s->link_counters + 8);
  d.peer_recv[0] = d.peer_recv[1] = nullptr;
  d.peer_flag[0] = d.peer_flag[1] = nullptr;
  return SAYAL_OK;
}

// `peer_block` is the neighbour's link block as addressable from this device.  My side-0 neighbour receives my
// rows on ITS side 1, and vice versa.
int slab_link_connect(Sim* s, int side, void* peer_block, size_t peer_stage_elems) {
  if (side != 0 && side != 1) return set_error(SAYAL_EINVAL, "slab link: side must be 0 or 1");
  int r = slab_link_alloc(s);
  if (r != SAYAL_OK) return r;
  if (peer_stage_elems != s->link.stage_elems) return set_error(SAYAL_EINVAL, "slab link: neighbours disagree on halo x width");
  const int their_side = 1 - side;
  s->link.peer_flag[side] = reinterpret_cast<unsigned*>(peer_block) + their_side;
  s->link.peer_recv[side] = reinterpret_cast<float*>(reinterpret_cast<char*>(peer_block) + 256) + (size_t)their_side * 2 * peer_stage_elems;
  return SAYAL_OK;
}

size_t slab_link_stage_elems(const Sim* s) { return stage_elems(s); }

int launch_slab_exchange(Sim* s, int field_mask) { return launch_slab_exchange_on(s, field_mask, s->stream); }

int launch_slab_exchange_on(Sim* s, int field_mask, cudaStream_t stream) {
  SlabLinkDev& d = s->link;
  if (!s->link_block || (!d.peer_recv[0] && !d.peer_recv[1])) return SAYAL_OK;  // no neighbours: nothing to do
  Fields3 fs;
  fs.n = 0;
  const int smoke_rows = s->slab_halo < s->advect_margin + 2 ? s->slab_halo : s->advect_margin + 2;
  if (field_mask & 1) { fs.rows[fs.n] = s->slab_halo; fs.f[fs.n++] = s->u; }
  if (field_mask & 2) { fs.rows[fs.n] = s->slab_halo; fs.f[fs.n++] = s->v; }
  if (field_mask & 4) { fs.rows[fs.n] = smoke_rows; fs.f[fs.n++] = s->smoke; }
  if (fs.n == 0) return set_error(SAYAL_EINVAL, "slab exchange: field_mask selects nothing");
  size_t items = 0;
  for (int k = 0; k < fs.n; k++) items += (size_t)fs.rows[k] * s->g.W / 4;
  for (int k = fs.n; k < 3; k++) { fs.f[k] = fs.f[0]; fs.rows[k] = 0; }
  int blocks = (int)((items + XTHREADS - 1) / XTHREADS);
  if (blocks < 1) blocks = 1;
  // at most 8 fat CTAs per side: the exchange shares the GPU with the interior tiles it overlaps with, and the
  // waiting CTAs of one slab must never fill the device, or a second slab on the same device (tests) could not
  // run the kernels the wait is waiting for
  if (blocks > 8) blocks = 8;
  slab_exchange_kernel<<<dim3(blocks, 2), XTHREADS, 0, stream>>>(s->g, s->slab_halo, fs, d);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(SAYAL_ECUDA, cudaGetErrorString(e));
  s->launches++;
  return SAYAL_OK;
}

}  // namespace sayal
