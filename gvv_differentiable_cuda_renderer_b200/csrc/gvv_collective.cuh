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

constexpr int kMaxWorld = 8;    // one NVSwitch domain of this path: the 8 GPUs of a box

struct ARParams {
  const float* const* peers;   // DEVICE array [world]: base of every rank's symmetric buffer (own included)
  uint32_t* const* pads;       // DEVICE array [world]: signal pads (uint32 words, all zero between barriers)
  const float* mc;             // multicast (NVLS) mapping of the buffer, or nullptr
  float* result;               // local output [count]
  long long offset, count;     // range in floats inside every buffer; offset is a multiple of 4
  int rank, world, mode;       // mode 0 = peer loads, 1 = NVLS multimem.ld_reduce
  int blocks, channelBase;     // CTAs taking part (0 = no collective); signal word = (channelBase + block) * world + peer
  int epochBase;               // word index (own pad) of the per-CTA barrier counters: epochBase + block
};

// Barrier signals are monotonic counters: arriving at barrier number e, a CTA adds 1 to its word in every peer's pad
// (fire-and-forget red, release at system scope: no NVLink round trip to wait for) and spins on its own pad until the
// word of every peer has reached e.  Nothing is ever reset, so a rank that runs ahead cannot be confused with a
// previous barrier; the counter e lives in the CTA's own pad and only that CTA touches it.
__device__ __forceinline__ void ar_signal(uint32_t* addr) {
  asm volatile("red.release.sys.global.add.u32 [%0], 1;" :: "l"(addr) : "memory");
}
__device__ __forceinline__ unsigned ar_poll(const uint32_t* addr) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(addr) : "memory");
  return v;
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
  uint32_t* epochWord = a.pads[a.rank] + a.epochBase + block;
  unsigned epoch = 0;
  if (tid < a.world) {
    const size_t word = (size_t)(a.channelBase + block) * a.world;
    epoch = *reinterpret_cast<volatile uint32_t*>(epochWord) + 1u;      // this CTA's barrier number (written back below)
    __threadfence_system();
    ar_signal(a.pads[tid] + word + a.rank);                              // tell peer `tid`: this rank's gradients are complete
    const uint32_t* mine = a.pads[a.rank] + word + tid;
    while ((int)(ar_poll(mine) - epoch) < 0) { }                         // until peer `tid` has said the same for this barrier
  }
  __syncthreads();
  if (tid == 0) *reinterpret_cast<volatile uint32_t*>(epochWord) = epoch;
  // Remote loads cost a full NVLink round trip (microseconds) each, so the exchange is sized to be ONE round trip deep:
  // every thread issues all the loads of two float4 columns (2 x W peer loads, or 2 multimem loads) before it adds anything,
  // and the caller picks enough CTAs that the range is covered in about one such pass (sharding.SymmetricGradBuffer).
  const long long n4 = a.count >> 2;
  const long long stride = (long long)a.blocks * nth;
  for (long long i0 = (long long)block * nth + tid; i0 < n4; i0 += 2 * stride) {
    const long long i1 = i0 + stride;
    const bool two = i1 < n4;
    const long long off0 = a.offset + 4 * i0, off1 = a.offset + 4 * (two ? i1 : i0);
    float4 s0, s1;
    if (a.mode == 1) {
      s0 = ar_ld_reduce(a.mc + off0);
      s1 = ar_ld_reduce(a.mc + off1);
    } else {
      float4 v0[kMaxWorld], v1[kMaxWorld];
#pragma unroll
      for (int p = 0; p < kMaxWorld; ++p)
        if (p < a.world) { v0[p] = ar_ld_sys(a.peers[p] + off0); v1[p] = ar_ld_sys(a.peers[p] + off1); }
      s0 = v0[0]; s1 = v1[0];
#pragma unroll
      for (int p = 1; p < kMaxWorld; ++p)                            // fixed rank order: the same bits on every rank
        if (p < a.world) {
          s0.x += v0[p].x; s0.y += v0[p].y; s0.z += v0[p].z; s0.w += v0[p].w;
          s1.x += v1[p].x; s1.y += v1[p].y; s1.z += v1[p].z; s1.w += v1[p].w;
        }
    }
    *reinterpret_cast<float4*>(a.result + 4 * i0) = s0;
    if (two) *reinterpret_cast<float4*>(a.result + 4 * i1) = s1;
  }
  if (block == 0 && tid < (int)(a.count & 3)) {                    // tail of a count that is not a multiple of 4
    const long long i = (n4 << 2) + tid;
    float s = 0.f;
    for (int p = 0; p < a.world; ++p) s += ar_ld_sys1(a.peers[p] + a.offset + i);
    a.result[i] = s;
  }
}

}  // namespace gvv
