// gvv_collective.cuh -- one-shot all-reduce of the shared-parameter gradients over peer memory (NVLink 5 / NVSwitch).
//
// The reference is single-GPU; this is the only exchange step of the sharded path: the gradients of
// parameters that several ranks share (SH, colours, texture; positions when the cameras of one batch element are
// split over ranks) are summed across ranks once per step.  The message is small (config 2: 27*C + 3*N floats =
// 420 KB), so the collective is LATENCY-bound: a ring/tree with several launches and hops costs more than moving
// the bytes.  One-shot instead: every rank keeps its gradients in a symmetric buffer that all peers have mapped;
// after a signal-pad barrier each rank reads the W copies and adds them up itself -- one launch, one NVLink round
// trip -- or, with NVLS, reads the SUM from the switch (multimem.ld_reduce on the multicast mapping: the switch
// adds the W replicas, one load instead of W).
//
// The device code below runs as a few CTAs INSIDE the backward's last kernel (normal_term_kernel): the vertex-
// normal term only touches vertex_pos_grad, while SH and colour gradients are final when pixel_grad_kernel ends, so
// the exchange overlaps that kernel's math instead of following it (and the call stays one linear, CUDA-graph-
// capturable chain).  Ranges that include vertex_pos_grad are reduced by the same code in a launch of its own.
//
// Memory model: the previous kernel's writes to this rank's buffer are performed device-wide at the kernel
// boundary; the barrier's release / acquire pair at system scope orders them before every peer's loads, which are
// system-scope (never served from a stale L1 line).  Sums run in rank order 0..W-1 on every rank: bit-identical
// results everywhere.  A slot may be rewritten two barriers later (callers alternate two slots), see sharding.py.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace gvv {

constexpr int kMaxWorld = 16;

struct ARParams {
  const float* const* peers;   // DEVICE array [world]: base of every rank's symmetric buffer (own included)
  uint32_t* const* pads;       // DEVICE array [world]: signal pads (uint32 words, all zero between barriers)
  const float* mc;             // multicast (NVLS) mapping of the buffer, or nullptr
  float* result;               // local output [count]
  long long offset, count;     // range in floats inside every buffer; offset is a multiple of 4
  int rank, world, mode;       // mode 0 = peer loads, 1 = NVLS multimem.ld_reduce
  int blocks, channelBase;     // CTAs taking part (0 = no collective); signal word = (channelBase + block) * world + peer
};

__device__ __forceinline__ void ar_put(uint32_t* addr) {     // 0 -> 1, release at system scope
  unsigned old;
  do { asm volatile("atom.global.release.sys.cas.b32 %0, [%1], 0, 1;" : "=r"(old) : "l"(addr) : "memory"); } while (old != 0u);
}
__device__ __forceinline__ void ar_wait(uint32_t* addr) {    // 1 -> 0, acquire at system scope
  unsigned old;
  do { asm volatile("atom.global.acquire.sys.cas.b32 %0, [%1], 1, 0;" : "=r"(old) : "l"(addr) : "memory"); } while (old != 1u);
}
__device__ __forceinline__ float4 ar_ld_sys(const float* p) {
  float4 v;
  asm volatile("ld.global.relaxed.sys.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float ar_ld_sys1(const float* p) {
  float v;
  asm volatile("ld.global.relaxed.sys.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 ar_ld_reduce(const float* mc) {   // the switch adds the replicas (NVLS)
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(mc) : "memory");
  return v;
}

// One CTA's share of the all-reduce; every thread of the CTA calls it (it contains a block barrier).
__device__ __forceinline__ void allreduce_block(const ARParams& a, int block) {
  const int tid = threadIdx.x, nth = blockDim.x;
  if (tid < a.world) {
    const size_t word = (size_t)(a.channelBase + block) * a.world;
    __threadfence_system();
    ar_put(a.pads[tid] + word + a.rank);        // tell peer `tid` that this rank's gradients are complete
    ar_wait(a.pads[a.rank] + word + tid);       // wait until peer `tid` says the same, and reset the word
  }
  __syncthreads();
  const long long n4 = a.count >> 2;
  for (long long i = (long long)block * nth + tid; i < n4; i += (long long)a.blocks * nth) {
    const long long off = a.offset + 4 * i;
    float4 s;
    if (a.mode == 1) {
      s = ar_ld_reduce(a.mc + off);
    } else {
      float4 v[kMaxWorld];
#pragma unroll
      for (int p = 0; p < kMaxWorld; ++p)
        if (p < a.world) v[p] = ar_ld_sys(a.peers[p] + off);      // all W loads in flight, then a fixed-order sum
      s = v[0];
#pragma unroll
      for (int p = 1; p < kMaxWorld; ++p)
        if (p < a.world) { s.x += v[p].x; s.y += v[p].y; s.z += v[p].z; s.w += v[p].w; }
    }
    *reinterpret_cast<float4*>(a.result + 4 * i) = s;
  }
  if (block == 0 && tid < (int)(a.count & 3)) {                    // tail of a count that is not a multiple of 4
    const long long i = (n4 << 2) + tid;
    float s = 0.f;
    for (int p = 0; p < a.world; ++p) s += ar_ld_sys1(a.peers[p] + a.offset + i);
    a.result[i] = s;
  }
}

}  // namespace gvv
