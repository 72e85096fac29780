// Exchange steps of the bin-sharded training step over NVLink peer memory (one process per GPU of one node).
//
// With the per-bin solves sharded over the ranks a step has three small exchanges on its critical
// path: the all-gather of y (K x G complex64, 1.5 MB), the sum of dL/dhy (G x tn float32, 568 KB) and the sum of the
// flat parameter-gradient bucket (< 1 MB). They are latency bound; a library all-reduce costs ~15-30 us each. Here every
// rank owns one SYMMETRIC buffer (same layout on every rank, every rank holds the device pointers of all of them: P2P
// mappings over NVLink / NVSwitch set up by the caller) and an exchange is
//
//   push:  every rank stores its contribution into EVERY rank's buffer, in the slot reserved for it (posted writes: no
//          round trip), then raises its flag in every rank's buffer (st.release.sys of a sequence number);
//   wait:  a rank spins on the flags in its OWN memory until every rank's sequence number arrived (ld.acquire.sys), then
//          copies (all-gather) or adds the slots in rank order (all-reduce: deterministic and identical on every rank).
//
// No collective library call, no host synchronisation: both kernels are ordinary stream-ordered launches and sit inside
// the captured CUDA graph of the step. The data region of a channel is double buffered by the parity of its sequence
// number: a rank can be at most one exchange ahead of the slowest one (it cannot complete exchange t + 1 before everybody
// pushed t + 1, i.e. finished reading t), so parity t + 2 == t is free again by then. A lost peer aborts the launch after
// ~30 s (trap) instead of hanging the device.
#include "common.cuh"

namespace dgfdn {
namespace {

constexpr int kMaxPeers = 8;
constexpr int kMaxChannels = 8;
constexpr long long kSpinLimit = 60000000000LL;  // clock64 ticks (~30 s: ranks may be seconds apart around set-up / capture)

struct PeerParams {
  int world, rank, channel;
  unsigned char* peer[kMaxPeers];  // base of every rank's symmetric buffer as mapped in this process
  int64_t flags_off;               // flags[channel][source rank] (uint32), at the same offset in every buffer
  int64_t data_off;                // two regions (parity) of `region_bytes`, slot of source rank r at r * slot_stride
  int64_t region_bytes, slot_stride;
  unsigned int* seq;               // [kMaxChannels] local: sequence number of the last completed push per channel
  unsigned int* done;              // [kMaxChannels] local: blocks of the running push that finished their part
};

__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned int ld_volatile(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// src (nbytes, multiple of 16) -> slot `rank` of the current parity in every rank's buffer; the last block raises the flags.
__global__ void __launch_bounds__(256) peer_push_kernel(PeerParams p, const uint4* __restrict__ src, int64_t n16) {
  const unsigned int seq = ld_volatile(p.seq + p.channel) + 1u;  // (bumped by the last block below, after everybody read it)
  const int64_t region = p.data_off + (int64_t)(seq & 1u) * p.region_bytes + (int64_t)p.rank * p.slot_stride;
  const int64_t total = n16 * p.world;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int dst = (int)(i / n16);
    const int64_t j = i - (int64_t)dst * n16;
    reinterpret_cast<uint4*>(p.peer[dst] + region)[j] = src[j];
  }
  __threadfence_system();  // this thread's stores are visible system wide before the ticket below
  __syncthreads();
  __shared__ unsigned int ticket;
  if (threadIdx.x == 0) ticket = atomicAdd(p.done + p.channel, 1u);
  __syncthreads();
  if (ticket != gridDim.x - 1) return;
  // last block: every block's stores are fenced; publish
  __threadfence();
  if (threadIdx.x < p.world) {
    unsigned int* flag = reinterpret_cast<unsigned int*>(p.peer[threadIdx.x] + p.flags_off) + p.channel * kMaxPeers + p.rank;
    st_release_sys(flag, seq);
  }
  if (threadIdx.x == 0) {
    p.done[p.channel] = 0u;
    p.seq[p.channel] = seq;
  }
}

// Block-level wait for the current sequence number of the channel from every rank (flags in this rank's own buffer).
__device__ __forceinline__ unsigned int wait_all(const PeerParams& p) {
  const unsigned int seq = ld_volatile(p.seq + p.channel);
  if (threadIdx.x < p.world) {
    const unsigned int* flag =
        reinterpret_cast<const unsigned int*>(p.peer[p.rank] + p.flags_off) + p.channel * kMaxPeers + threadIdx.x;
    const long long t0 = clock64();
    while ((int)(ld_acquire_sys(flag) - seq) < 0) {
      if (clock64() - t0 > kSpinLimit) __trap();  // a lost peer must abort the launch, never hang the device
    }
  }
  __syncthreads();
  return seq;
}

// all-gather: out (world * n16 uint4) <- the world slots of the current parity, in rank order
__global__ void __launch_bounds__(256) peer_gather_kernel(PeerParams p, uint4* __restrict__ out, int64_t n16) {
  const unsigned int seq = wait_all(p);
  const unsigned char* region = p.peer[p.rank] + p.data_off + (int64_t)(seq & 1u) * p.region_bytes;
  const int64_t total = n16 * p.world;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / n16);
    const int64_t j = i - (int64_t)r * n16;
    out[i] = __ldcg(reinterpret_cast<const uint4*>(region + (int64_t)r * p.slot_stride) + j);  // L2: peers wrote it
  }
}

// all-reduce (SUM, float32): out[i] = slot_0[i] + slot_1[i] + ... in rank order
__global__ void __launch_bounds__(256) peer_reduce_kernel(PeerParams p, float4* __restrict__ out, int64_t n16) {
  const unsigned int seq = wait_all(p);
  const unsigned char* region = p.peer[p.rank] + p.data_off + (int64_t)(seq & 1u) * p.region_bytes;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (int64_t)gridDim.x * blockDim.x) {
    float4 acc = __ldcg(reinterpret_cast<const float4*>(region) + i);  // L2: peers wrote it
    for (int r = 1; r < p.world; ++r) {
      const float4 v = __ldcg(reinterpret_cast<const float4*>(region + (int64_t)r * p.slot_stride) + i);
      acc.x += v.x, acc.y += v.y, acc.z += v.z, acc.w += v.w;
    }
    out[i] = acc;
  }
}

int fill(PeerParams& p, int world, int rank, int channel, const void* const* peer_ptrs, int64_t flags_off, int64_t data_off,
         int64_t region_bytes, int64_t slot_stride, void* state, int64_t nbytes) {
  DGFDN_CHECK(world >= 1 && world <= kMaxPeers && rank >= 0 && rank < world, "peer: world %d / rank %d out of range", world, rank);
  DGFDN_CHECK(channel >= 0 && channel < kMaxChannels, "peer: channel %d out of range", channel);
  DGFDN_CHECK(peer_ptrs && state, "peer: null pointer");
  DGFDN_CHECK(nbytes > 0 && nbytes % 16 == 0 && slot_stride % 16 == 0 && slot_stride >= nbytes && data_off % 16 == 0 &&
                  region_bytes % 16 == 0 && region_bytes >= slot_stride * world,
              "peer: sizes must be multiples of 16 bytes, slot_stride >= nbytes, region >= world slots");
  DGFDN_CHECK(flags_off % 4 == 0 && flags_off + (int64_t)kMaxChannels * kMaxPeers * 4 <= data_off, "peer: flags overlap the data");
  p.world = world;
  p.rank = rank;
  p.channel = channel;
  for (int i = 0; i < world; ++i) {
    DGFDN_CHECK(peer_ptrs[i] != nullptr, "peer: null buffer pointer of rank %d", i);
    p.peer[i] = static_cast<unsigned char*>(const_cast<void*>(peer_ptrs[i]));
  }
  p.flags_off = flags_off;
  p.data_off = data_off;
  p.region_bytes = region_bytes;
  p.slot_stride = slot_stride;
  p.seq = static_cast<unsigned int*>(state);
  p.done = p.seq + kMaxChannels;
  return 0;
}

int grid_for(int64_t n16) {
  const int64_t want = (n16 + 255) / 256;
  const int64_t cap = 2 * (int64_t)sm_count();
  return (int)(want < 1 ? 1 : (want < cap ? want : cap));
}

}  // namespace
}  // namespace dgfdn

using namespace dgfdn;

extern "C" int64_t dgfdn_peer_state_bytes(void) { return 2 * kMaxChannels * (int64_t)sizeof(unsigned int); }
extern "C" int64_t dgfdn_peer_flags_bytes(void) { return (int64_t)kMaxChannels * kMaxPeers * sizeof(unsigned int); }

extern "C" int dgfdn_peer_push(int world, int rank, int channel, const void* const* peer_ptrs, int64_t flags_off,
                               int64_t data_off, int64_t region_bytes, int64_t slot_stride, void* state, const void* src,
                               int64_t nbytes, void* stream) {
  PeerParams p{};
  if (fill(p, world, rank, channel, peer_ptrs, flags_off, data_off, region_bytes, slot_stride, state, nbytes)) return 1;
  DGFDN_CHECK(src && (reinterpret_cast<uintptr_t>(src) & 15u) == 0, "peer_push: src must be 16-byte aligned");
  const int64_t n16 = nbytes / 16;
  peer_push_kernel<<<grid_for(n16 * world), 256, 0, static_cast<cudaStream_t>(stream)>>>(p, static_cast<const uint4*>(src), n16);
  DGFDN_LAUNCH_CHECK();
  return 0;
}

extern "C" int dgfdn_peer_gather(int world, int rank, int channel, const void* const* peer_ptrs, int64_t flags_off,
                                 int64_t data_off, int64_t region_bytes, int64_t slot_stride, void* state, void* out,
                                 int64_t nbytes, void* stream) {
  PeerParams p{};
  if (fill(p, world, rank, channel, peer_ptrs, flags_off, data_off, region_bytes, slot_stride, state, nbytes)) return 1;
  DGFDN_CHECK(out && (reinterpret_cast<uintptr_t>(out) & 15u) == 0, "peer_gather: out must be 16-byte aligned");
  const int64_t n16 = nbytes / 16;
  peer_gather_kernel<<<grid_for(n16 * world), 256, 0, static_cast<cudaStream_t>(stream)>>>(p, static_cast<uint4*>(out), n16);
  DGFDN_LAUNCH_CHECK();
  return 0;
}

extern "C" int dgfdn_peer_reduce(int world, int rank, int channel, const void* const* peer_ptrs, int64_t flags_off,
                                 int64_t data_off, int64_t region_bytes, int64_t slot_stride, void* state, void* out,
                                 int64_t nbytes, void* stream) {
  PeerParams p{};
  if (fill(p, world, rank, channel, peer_ptrs, flags_off, data_off, region_bytes, slot_stride, state, nbytes)) return 1;
  DGFDN_CHECK(out && (reinterpret_cast<uintptr_t>(out) & 15u) == 0, "peer_reduce: out must be 16-byte aligned");
  const int64_t n16 = nbytes / 16;
  peer_reduce_kernel<<<grid_for(n16), 256, 0, static_cast<cudaStream_t>(stream)>>>(p, static_cast<float4*>(out), n16);
  DGFDN_LAUNCH_CHECK();
  return 0;
}
