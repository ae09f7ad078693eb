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
// up, raises the sim's sticky `link_error` word (device memory, mirrored once into mapped host memory) and skips the unpack instead of hanging the
// GPU or consuming a stale buffer; sayal_sync / sayal_run / sayal_get_field then return SAYAL_ELINK.
//
// The same block carries two more neighbour-to-neighbour messages: the pass flags of projection passes that push
// their edge rows themselves (projection_pack.cu) and the chain reduction of the pressure range.
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "sayal_internal.h"

namespace sayal {

namespace {

constexpr int XTHREADS = 1024;
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
// A wait that gives up (neighbour gone, or its host never enqueued the matching step) raises the sticky link_error
// word and does NOT unpack: the ghost rows keep their last valid contents, later exchanges of this sim return at
// once (no further two-second spins), and the host reports SAYAL_ELINK at its next synchronisation point.
__global__ void __launch_bounds__(XTHREADS) slab_exchange_kernel(Grid g, int halo, Fields3 fs, SlabLinkDev d) {
  const int side = blockIdx.y;
  if (!d.peer_recv[side]) return;
  __shared__ int s_ok;
  const unsigned seq = d.send_seq[side];  // == recv_seq: advanced below by the last CTA, after every CTA has read it
  const bool broken = *reinterpret_cast<volatile int*>(d.link_error) != 0;
  if (!broken) {
    float* dst = d.peer_recv[side] + (size_t)(seq & 1u) * d.stage_elems;
    copy_rows<false>(g, fs, side == 0 ? g.own_lo : g.own_hi, side == 1, dst, blockIdx.x, gridDim.x);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int ok = !broken;
    if (ok) {
      __threadfence_system();
      unsigned t = atomicAdd(&d.ticket[side], 1u);
      if (t == gridDim.x - 1) {
        d.ticket[side] = 0;
        __threadfence_system();
        st_release_sys(d.peer_words[side] + LW_XFLAG + (1 - side), seq + 1);
      }
      const long long t0 = now_ns();
      while ((int)(ld_acquire_sys(d.my_words + LW_XFLAG + side) - (seq + 1)) < 0) {
        if (now_ns() - t0 > kSpinLimitNs || *reinterpret_cast<volatile int*>(d.link_error) != 0) {
          raise_link_error(d.link_error, LINK_TIMEOUT);  // neighbour gone: do not hang the GPU
          ok = 0;
          break;
        }
      }
    }
    s_ok = ok;
  }
  __syncthreads();
  if (s_ok) {
    const float* src = d.my_recv[side] + (size_t)(seq & 1u) * d.stage_elems;
    copy_rows<true>(g, fs, side == 0 ? g.own_lo : g.own_hi, side == 0, const_cast<float*>(src), blockIdx.x, gridDim.x);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned t = atomicAdd(&d.ticket[2 + side], 1u);
    if (t == gridDim.x - 1) {
      d.ticket[2 + side] = 0;
      d.send_seq[side] = seq + 1;
    }
  }
}

// Global pressure range of linked slabs (Fluid::min_pressure / max_pressure are ONE pair for the frame,
// fluid.cu:778-787, graphics_handler.cu:288-289): the local ordered-int pairs travel down the chain (side 0 -> 1),
// each slab folding in its own, and the result travels back up.  One thread; 2 (N-1) hops of a few microseconds on
// the aux stream, off the step's critical path.  range[0..1] = local (pressure_range_kernel), range[2..3] = global.
__global__ void slab_range_reduce_kernel(SlabLinkDev d, int32_t* range) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const unsigned seq = *d.range_seq + 1;
  int mn = range[0], mx = range[1];
  bool ok = *reinterpret_cast<volatile int*>(d.link_error) == 0;
  if (ok && d.peer_words[0]) {  // partial result of the slabs above me
    ok = spin_until(d.my_words + LW_RANGE_DOWN, seq, d.link_error);
    if (ok) {
      mn = min(mn, (int)__ldcg(d.my_words + LW_RANGE_DOWN + 1));
      mx = max(mx, (int)__ldcg(d.my_words + LW_RANGE_DOWN + 2));
    }
  }
  if (ok && d.peer_words[1]) {
    unsigned* w = d.peer_words[1] + LW_RANGE_DOWN;
    w[1] = (unsigned)mn;
    w[2] = (unsigned)mx;
    __threadfence_system();
    st_release_sys(w, seq);
    ok = spin_until(d.my_words + LW_RANGE_UP, seq, d.link_error);  // the global result coming back
    if (ok) {
      mn = (int)__ldcg(d.my_words + LW_RANGE_UP + 1);
      mx = (int)__ldcg(d.my_words + LW_RANGE_UP + 2);
    }
  }
  if (ok && d.peer_words[0]) {
    unsigned* w = d.peer_words[0] + LW_RANGE_UP;
    w[1] = (unsigned)mn;
    w[2] = (unsigned)mx;
    __threadfence_system();
    st_release_sys(w, seq);
  }
  range[2] = mn;
  range[3] = mx;
  *d.range_seq = seq;
}

// After the last projection pass of a step in push mode (projection_pack.cu): the neighbours' pushes of THEIR last
// pass must have landed in my ghost rows before the advection reads them.  Thread `side` waits for that side's pass
// flag and cross-checks the neighbour's plan signature (iterations and passes per step: both sides must split the
// projection the same way, or pass k of one would consume pass k of the other at a different iteration count);
// then the step counter advances.  Pass flags count (step << 10) + passes published.
__global__ void slab_push_wait_kernel(SlabLinkDev d, int passes, int signature) {
  const int side = threadIdx.x;
  if (side < 2 && d.peer_words[side] && *reinterpret_cast<volatile int*>(d.link_error) == 0) {
    const unsigned target = (*d.step_seq << 10) + (unsigned)passes;
    if (spin_until(d.my_words + LW_PFLAG + side, target, d.link_error)) {
      if ((int)__ldcg(d.my_words + LW_PITER + side) != signature) raise_link_error(d.link_error, LINK_PLAN_MISMATCH);
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) *d.step_seq = *d.step_seq + 1;
}

size_t stage_elems(const Sim* s) { return (size_t)s->slab_halo * s->g.W * 3; }
size_t block_bytes(const Sim* s) { return LW_WORDS * sizeof(unsigned) + 4 * stage_elems(s) * sizeof(float); }  // control words + 2 sides x 2 parities

}  // namespace

// Allocate the neighbour-writable block and the private counters of a slab sim.
int slab_link_alloc(Sim* s) {
  if (s->link_block) return SAYAL_OK;
  if (s->slab_halo <= 0) return set_error(SAYAL_EINVAL, "slab link: the sim was created without ghost rows");
  size_t bytes = block_bytes(s);
  cudaError_t e = cudaMalloc(&s->link_block, bytes);
  if (e != cudaSuccess) return set_error(SAYAL_ENOMEM, cudaGetErrorString(e));
  cudaMemsetAsync(s->link_block, 0, bytes, s->stream);
  e = cudaMalloc(&s->link_counters, 32 * sizeof(unsigned));
  if (e != cudaSuccess) return set_error(SAYAL_ENOMEM, cudaGetErrorString(e));
  cudaMemsetAsync(s->link_counters, 0, 32 * sizeof(unsigned), s->stream);
  e = cudaHostAlloc(reinterpret_cast<void**>(&s->h_link_error), sizeof(int), cudaHostAllocMapped);
  if (e != cudaSuccess) return set_error(SAYAL_ENOMEM, cudaGetErrorString(e));
  *s->h_link_error = LINK_OK;
  int* d_err = nullptr;
  e = cudaHostGetDevicePointer(reinterpret_cast<void**>(&d_err), s->h_link_error, 0);
  if (e != cudaSuccess) return set_error(SAYAL_ECUDA, cudaGetErrorString(e));
  e = cudaStreamSynchronize(s->stream);
  if (e != cudaSuccess) return set_error(SAYAL_ECUDA, cudaGetErrorString(e));
  SlabLinkDev& d = s->link;
  d.stage_elems = stage_elems(s);
  d.my_words = reinterpret_cast<unsigned*>(s->link_block);
  float* recv = reinterpret_cast<float*>(reinterpret_cast<char*>(s->link_block) + LW_WORDS * sizeof(unsigned));
  d.my_recv[0] = recv;
  d.my_recv[1] = recv + 2 * d.stage_elems;
  d.send_seq = s->link_counters;
  d.recv_seq = s->link_counters + 2;
  d.ticket = s->link_counters + 4;
  d.range_seq = s->link_counters + 8;
  d.step_seq = s->link_counters + 9;
  d.push_ticket = s->link_counters + 10;  // [2]
  // device word + the address of its host mirror (raise_link_error)
  d.link_error = reinterpret_cast<int*>(s->link_counters + 12);
  e = cudaMemcpy(s->link_counters + 14, &d_err, sizeof d_err, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) return set_error(SAYAL_ECUDA, cudaGetErrorString(e));
  d.peer_recv[0] = d.peer_recv[1] = nullptr;
  d.peer_words[0] = d.peer_words[1] = nullptr;
  return SAYAL_OK;
}

// What a neighbour needs to know about this slab (traded once; between processes it travels as bytes).
int slab_link_export(Sim* s, LinkInfo* out) {
  int r = slab_link_alloc(s);
  if (r != SAYAL_OK) return r;
  std::memset(out, 0, sizeof *out);
  out->stage_elems = (int64_t)stage_elems(s);
  out->vel_stride = (int64_t)s->vel_stride;
  out->W = s->g.W;
  out->pitch = s->g.pitch;
  out->local_rows = s->g.local_rows;
  out->own_lo = s->g.own_lo;
  out->own_hi = s->g.own_hi;
  out->halo = s->slab_halo;
  out->advect_margin = s->advect_margin;
  out->parity = s->parity;
  out->row_base = s->g.row_base;
  out->global_height = s->g.H;
  out->abi = SAYAL_ABI_VERSION;
  return SAYAL_OK;
}

// `peer_block` / `peer_vel` are the neighbour's link block and velocity block as addressable from this device.
// My side-0 neighbour receives my rows on ITS side 1, and vice versa.
int slab_link_connect_info(Sim* s, int side, const LinkInfo* info, void* peer_block, void* peer_vel) {
  if (side != 0 && side != 1) return set_error(SAYAL_EINVAL, "slab link: side must be 0 or 1");
  int r = slab_link_alloc(s);
  if (r != SAYAL_OK) return r;
  if (info->abi != SAYAL_ABI_VERSION) return set_error(SAYAL_EINVAL, "slab link: neighbour runs another ABI version");
  if ((size_t)info->stage_elems != s->link.stage_elems || info->halo != s->slab_halo || info->W != s->g.W || info->pitch != s->g.pitch ||
      info->global_height != s->g.H)
    return set_error(SAYAL_EINVAL, "slab link: neighbours disagree on halo x width");
  // the neighbour must hold the rows next to mine: its owned rows end where mine begin (side 0) or begin where mine end
  const int my_first = s->g.row_base + s->g.own_lo, my_end = s->g.row_base + s->g.own_hi;
  const int their_first = info->row_base + info->own_lo, their_end = info->row_base + info->own_hi;
  if (side == 0 ? their_end != my_first : their_first != my_end)
    return set_error(SAYAL_EINVAL, "slab link: the neighbour's rows do not adjoin this slab on that side");
  if (s->g.own_hi - s->g.own_lo < s->slab_halo || info->own_hi - info->own_lo < info->halo)
    return set_error(SAYAL_EINVAL, "slab link: a linked slab must own at least `halo` rows (its edge rows are what travels)");
  if ((info->parity & 1) != (s->parity & 1))
    return set_error(SAYAL_EINVAL, "slab link: neighbours must be linked in the same buffer phase (link before stepping)");
  const int their_side = 1 - side;
  s->link.peer_words[side] = reinterpret_cast<unsigned*>(peer_block);
  s->link.peer_recv[side] = reinterpret_cast<float*>(reinterpret_cast<char*>(peer_block) + LW_WORDS * sizeof(unsigned)) +
                            (size_t)their_side * 2 * (size_t)info->stage_elems;
  for (int k = 0; k < 4; k++)
    s->peer_vel[side][k] = peer_vel ? reinterpret_cast<float*>(reinterpret_cast<char*>(peer_vel) + (size_t)k * info->vel_stride) : nullptr;
  // my rows [own_lo, own_lo + halo) are the neighbour's ghost rows [own_hi, own_hi + halo) (side 0), and my rows
  // [own_hi - halo, own_hi) its ghost rows [own_lo - halo, own_lo) (side 1)
  s->peer_ghost_row0[side] = side == 0 ? info->own_hi : info->own_lo - info->halo;
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
  static const int max_blocks = getenv("SAYAL_XCHG_BLOCKS") ? atoi(getenv("SAYAL_XCHG_BLOCKS")) : 8;  // experiments
  if (blocks > max_blocks) blocks = max_blocks;
  slab_exchange_kernel<<<dim3(blocks, 2), XTHREADS, 0, stream>>>(s->g, s->slab_halo, fs, d);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(SAYAL_ECUDA, cudaGetErrorString(e));
  s->launches++;
  return SAYAL_OK;
}

int launch_slab_range_reduce(Sim* s, cudaStream_t stream) {
  slab_range_reduce_kernel<<<1, 32, 0, stream>>>(s->link, s->d_range);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(SAYAL_ECUDA, cudaGetErrorString(e));
  s->launches++;
  return SAYAL_OK;
}

int launch_slab_push_wait(Sim* s, int passes, int signature) {
  slab_push_wait_kernel<<<1, 32, 0, s->stream>>>(s->link, passes, signature);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(SAYAL_ECUDA, cudaGetErrorString(e));
  s->launches++;
  return SAYAL_OK;
}

// Load this file's kernels now: CUDA loads a kernel lazily at its first launch, and that load can wait for the device
// to drain — which never happens while a linked slab on the same device spins for rows this thread has yet to enqueue.
int preload_slab() {
  cudaFuncAttributes fa;
  cudaError_t e = cudaSuccess;
  if (e == cudaSuccess) e = cudaFuncGetAttributes(&fa, slab_exchange_kernel);
  if (e == cudaSuccess) e = cudaFuncGetAttributes(&fa, slab_range_reduce_kernel);
  if (e == cudaSuccess) e = cudaFuncGetAttributes(&fa, slab_push_wait_kernel);
  return e == cudaSuccess ? SAYAL_OK : set_error(SAYAL_ECUDA, cudaGetErrorString(e));
}

}  // namespace sayal
