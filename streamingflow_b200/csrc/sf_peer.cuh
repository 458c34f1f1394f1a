// sf_peer.cuh -- the row-sharded rollout's exchange steps as kernels over NVLink peer memory (no NCCL call on the per-event path).
//
// Every rank owns one "arena" (cudaMalloc, exported to the other ranks of the node with cudaIpc*): arrival counters, receive
// buffers for the halo rows of its two neighbours (two copies, used alternately) and one slot per rank for the small vectors of
// an all-reduce (two copies).  The exchange of an event is then
//   halo_push_kernel   the band's boundary rows of all halo tensors -> STORED straight into the neighbours' receive buffers,
//                      system-scope fence, then the neighbours' arrival counters are released with the launch's sequence number;
//   halo_pull_kernel   waits (ld.acquire.sys) until both neighbours' rows of this sequence number have arrived, copies them into
//                      the local halos;
//   peer_allreduce_kernel (one block)   stores the rank's [n] partial sums into slot[rank] on EVERY rank, releases one counter per
//                      rank, waits for the world's contributions and adds the slots in rank order (the same order on every rank:
//                      all ranks get bit-identical sums, as an NCCL all-reduce would give).
// Sequence numbers live in device memory and are advanced by the kernels themselves, so the launches carry no per-call argument
// and a whole rollout (stages + exchanges) replays as ONE CUDA graph.  Two copies of every buffer are enough: a rank can run at
// most one exchange ahead of a neighbour, because its next push needs the rows that neighbour sends only after it has consumed
// the previous ones (stream order: pull(e) -> stages(e+1) -> push(e+1)); same argument for the all-reduce slots.
// Nothing here can spin forever: a wait gives up after timeout_ns, latches an error code in *err (checked by the host after the
// rollout) and every later wait returns at once.
#pragma once
#include <cstdint>

namespace sf {

constexpr int PEER_MAX = 8;                   // ranks of one NVSwitch node

__device__ __forceinline__ unsigned ld_acquire_sys_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys_u32(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// until the counter a peer GPU releases has reached `want` (wrap-safe); false (and *err latched) on time-out or earlier error
__device__ __forceinline__ bool peer_wait(const unsigned* flag, unsigned want, unsigned long long timeout_ns, int* err, int code) {
  if (*reinterpret_cast<volatile int*>(err) != 0) return false;
  const unsigned long long t0 = globaltimer_ns();
  while ((int)(ld_acquire_sys_u32(flag) - want) < 0) {
    __nanosleep(32);
    if (globaltimer_ns() - t0 > timeout_ns) {
      atomicCAS(err, 0, code);
      return false;
    }
  }
  return true;
}

struct PeerHalo {
  HaloCopy h;                  // h.flat[r]: copy 0 of the flat buffer of direction r (push: in the NEIGHBOUR's arena; pull: local)
  long long parity_stride;     // bytes between the two copies
  unsigned* flag[2];           // push: the neighbours' arrival counters; pull: the local ones
  unsigned* seq;               // local: [0] launches completed (sequence number), [1] block ticket
  int* err;
  unsigned long long timeout_ns;
  unsigned long long* trace;   // optional [64][4] ring of %globaltimer stamps per launch: start, wait done, end (profiling; NULL = off)
};

// unit u (16 bytes) of direction r's flat buffer -> its address in the tensors (32-bit index arithmetic: a flat buffer is < 64 GB)
__device__ __forceinline__ char* halo_element(const HaloCopy& h, unsigned u, int r) {
  int t = 0;
  for (; t < h.n_tensors - 1; ++t) {
    const unsigned sz = (unsigned)h.B * (unsigned)h.nrows * (unsigned)(h.row_bytes[t] >> 4);
    if (u < sz) break;
    u -= sz;
  }
  const unsigned chunk = (unsigned)h.nrows * (unsigned)(h.row_bytes[t] >> 4);
  const unsigned b = u / chunk, inner = u - b * chunk;
  return h.base[t] + (long long)b * h.batch_stride[t] + (long long)h.row0[r] * h.row_bytes[t] + ((long long)inner << 4);
}

// last block of the grid: advance the sequence number, reset the ticket; returns true in that block's thread 0
__device__ __forceinline__ bool peer_grid_done(unsigned* seq, unsigned value) {
  __syncthreads();
  if (threadIdx.x != 0) return false;
  __threadfence();
  const unsigned t = atomicAdd(seq + 1, 1u);
  if (t != gridDim.x - 1) return false;
  __threadfence();
  seq[1] = 0;
  seq[0] = value;
  return true;
}

__device__ __forceinline__ unsigned halo_units(const HaloCopy& h) {
  unsigned n = 0;
  for (int t = 0; t < h.n_tensors; ++t) n += (unsigned)h.B * (unsigned)h.nrows * (unsigned)(h.row_bytes[t] >> 4);
  return n;
}

__global__ void __launch_bounds__(256) halo_push_kernel(const PeerHalo a) {
  const unsigned seq = *reinterpret_cast<volatile unsigned*>(a.seq) + 1;       // every block reads it before the last one bumps it
  const HaloCopy& h = a.h;
  const unsigned units = halo_units(h);
  const long long par = (long long)(seq & 1u) * a.parity_stride;
  if (a.trace && blockIdx.x == 0 && threadIdx.x == 0) a.trace[(seq & 63u) * 4] = globaltimer_ns();
  for (int r = 0; r < 2; ++r) {
    if (h.flat[r] == nullptr) continue;
    uint4* dst = reinterpret_cast<uint4*>(h.flat[r] + par);                    // the neighbour's HBM
    for (unsigned u = blockIdx.x * blockDim.x + threadIdx.x; u < units; u += gridDim.x * blockDim.x)
      dst[u] = *reinterpret_cast<const uint4*>(halo_element(h, u, r));
  }
  // ONE system-scope fence per block, by the thread that takes the ticket, after the block barrier (the barrier orders the other
  // threads' stores before it; the fence is cumulative).  A fence per thread would serialise ~30 MEMBAR.SYS per SM, each waiting
  // for NVLink acknowledgements.
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();
    const unsigned t = atomicAdd(a.seq + 1, 1u);
    if (t == gridDim.x - 1) {
      __threadfence_system();
      a.seq[1] = 0;
      a.seq[0] = seq;
      for (int r = 0; r < 2; ++r)
        if (h.flat[r] != nullptr) st_release_sys_u32(a.flag[r], seq);
      if (a.trace) a.trace[(seq & 63u) * 4 + 2] = globaltimer_ns();
    }
  }
}

__global__ void __launch_bounds__(256) halo_pull_kernel(const PeerHalo a) {
  const unsigned seq = *reinterpret_cast<volatile unsigned*>(a.seq) + 1;
  const HaloCopy& h = a.h;
  __shared__ int ok;
  if (threadIdx.x == 0) {
    if (a.trace && blockIdx.x == 0) a.trace[(seq & 63u) * 4] = globaltimer_ns();
    bool good = true;
    for (int r = 0; r < 2; ++r)
      if (h.flat[r] != nullptr) good = peer_wait(a.flag[r], seq, a.timeout_ns, a.err, 1 + r) && good;
    ok = good;
    if (a.trace && blockIdx.x == 0) a.trace[(seq & 63u) * 4 + 1] = globaltimer_ns();
  }
  __syncthreads();
  if (ok) {
    const unsigned units = halo_units(h);
    const long long par = (long long)(seq & 1u) * a.parity_stride;
    for (int r = 0; r < 2; ++r) {
      if (h.flat[r] == nullptr) continue;
      const uint4* src = reinterpret_cast<const uint4*>(h.flat[r] + par);
      for (unsigned u = blockIdx.x * blockDim.x + threadIdx.x; u < units; u += gridDim.x * blockDim.x)
        *reinterpret_cast<uint4*>(halo_element(h, u, r)) = __ldcg(src + u);      // written by a peer: read at L2, never from a stale L1 line
    }
  }
  if (peer_grid_done(a.seq, seq) && a.trace) a.trace[(seq & 63u) * 4 + 2] = globaltimer_ns();
}

struct PeerReduce {
  float* data;                 // local [n]: in = this rank's partial sums, out = the sum over all ranks
  int n, n_max, rank, world;
  float* slots[PEER_MAX];      // slot arena of every rank: [2 copies][world][n_max]
  unsigned* flags[PEER_MAX];   // arrival counters of every rank: [world]
  unsigned* seq;               // local: [0] launches completed
  int* err;
  unsigned long long timeout_ns;
  unsigned long long* trace;   // optional [64][4] ring of %globaltimer stamps per launch (profiling; NULL = off)
};

__global__ void __launch_bounds__(256) peer_allreduce_kernel(const PeerReduce a) {
  const unsigned seq = *reinterpret_cast<volatile unsigned*>(a.seq) + 1;
  const size_t par = (size_t)(seq & 1u) * a.world * a.n_max;
  if (a.trace && threadIdx.x == 0) a.trace[(seq & 63u) * 4] = globaltimer_ns();
  for (int r = 0; r < a.world; ++r) {
    float* dst = a.slots[r] + par + (size_t)a.rank * a.n_max;
    for (int i = threadIdx.x; i < a.n; i += blockDim.x) dst[i] = a.data[i];
  }
  __shared__ int ok;
  if (threadIdx.x == 0) ok = 1;
  __syncthreads();
  if (threadIdx.x < a.world) {
    __threadfence_system();
    st_release_sys_u32(a.flags[threadIdx.x] + a.rank, seq);
    if (!peer_wait(a.flags[a.rank] + threadIdx.x, seq, a.timeout_ns, a.err, 3)) ok = 0;
  }
  __syncthreads();
  if (a.trace && threadIdx.x == 0) a.trace[(seq & 63u) * 4 + 1] = globaltimer_ns();
  if (ok) {
    const float* mine = a.slots[a.rank] + par;
    for (int i = threadIdx.x; i < a.n; i += blockDim.x) {
      float s = 0.0f;
      for (int r = 0; r < a.world; ++r) s += __ldcg(mine + (size_t)r * a.n_max + i);
      a.data[i] = s;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    a.seq[0] = seq;
    if (a.trace) a.trace[(seq & 63u) * 4 + 2] = globaltimer_ns();
  }
}

}  // namespace sf
