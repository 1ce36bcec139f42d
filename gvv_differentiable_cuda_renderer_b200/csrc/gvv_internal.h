// gvv_internal.h -- handle layout and launcher declarations shared by the .cu files.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "gvv_exact.cuh"
#include "gvv_collective.cuh"
#include "../../include/gvv_b200.h"

namespace gvv {

// A triangle whose bbox overlaps more than kMaxSmallTiles tiles goes to the per-view "big" list,
// which every tile of that view scans; everything else is appended to per-tile bins.  This bounds
// the bin pool by F*kMaxSmallTiles entries per view without ever reading a count back to the host.
constexpr int kMaxSmallTiles = 16;
constexpr int kMaxTiles = 40960;       // bin_scan_kernel keeps a view's tile histogram in shared memory (160 KB): e.g. 6400 x 6400 pixels at 32 x 32

struct Scratch {
  int capViews = 0, capBatch = 0;
  CamRec* cams = nullptr;        // [V]
  float4* proj = nullptr;        // [V*N]  (x/z, y/z, z, 0)
  float4* vscaled = nullptr;     // [B*N]  vertex / 1000 (div.rn), w = 0
  float4* fnorm4 = nullptr;      // [B*F]  face normal cross(v1-v0, v2-v0)
  float4* vnorm4 = nullptr;      // [B*N]  unnormalised vertex normal
  float4* vcol4 = nullptr;       // [B*N]  vertex colour
  int* tileCount = nullptr;      // [V*nT] self-cleaning (the raster kernel zeroes its own entry)
  int* tileCursor = nullptr;     // [V*nT] near triangles filled into the tile's bin (from its front); self-cleaning
  int* tileCursorFar = nullptr;  // [V*nT] far triangles filled (from the back of the bin); self-cleaning
  int* tileMinK = nullptr;       // [V*nT] min / max depth-key lower bound of the tile's triangles; reset by bin_scan_kernel
  int* tileMaxK = nullptr;
  int* tileThr = nullptr;        // [V*nT] near/far threshold of the tile
  int* tileOffset = nullptr;     // [V*nT]
  int* tileOrder = nullptr;      // [V*(nT+nT/2)] raster work items (tile strips) of a view, heaviest first; -1 = unused slot
  int* tileDone = nullptr;       // [V*nT] strips of a split tile that have finished; self-cleaning
  int* bigCount = nullptr;       // [V]    zeroed by camera_kernel of the next call
  int* bigList = nullptr;        // [V*F]
  int* bins = nullptr;           // [V*F*kMaxSmallTiles]
  float* gnorm = nullptr;        // [B*N*4] backward: dL/d(unnormalised vertex normal), float4-strided
  float4* bpos4 = nullptr;       // [B*N]  backward: repacked vertex_pos (raw)
  float4* bcol4 = nullptr;       // [B*N]  backward: repacked vertex_color
  float4* bnor4 = nullptr;       // [V*N]  backward: repacked vertex_normal input
  int* tileCounter = nullptr;    // [1]    backward: work counter of the persistent pixel-gradient kernel, set by prep_kernel
  unsigned long long* ctaTrace = nullptr;   // [V*nT*4] debug (option cta_trace): globaltimer start, end, bin size, SM id per raster CTA
};

// Optional per-kernel timing with CUDA events on the launching stream (bench.py roofline leg).
enum KernelSlot { K_CAMERA = 0, K_VERTEX, K_BIN_COUNT, K_BIN_SCAN, K_BIN_FILL, K_RASTER, K_ZERO, K_PIXEL_GRAD, K_NORMAL_TERM, K_NORMAL_MAP, K_ALLREDUCE, K_NUM_SLOTS };

struct KernelTimer {
  bool enabled = false;
  static constexpr int kCap = 4096;
  cudaEvent_t* ev = nullptr;   // 2*kCap events, created on first enable
  int* slot = nullptr;
  int used = 0;
  inline void begin(int s, cudaStream_t st) {
    if (enabled && used < kCap) { slot[used] = s; cudaEventRecord(ev[2 * used], st); }
  }
  inline void end(cudaStream_t st) {
    if (enabled && used < kCap) { cudaEventRecord(ev[2 * used + 1], st); ++used; }
  }
};

}  // namespace gvv

struct gvv_renderer {
  int device = 0;
  int F = 0, N = 0, C = 0, W = 0, H = 0;
  int albedo = 0, shading = 0, imgFilter = 1, texFilter = 1, computeNormalMap = 0;
  int tile = 32, tilesX = 0, tilesY = 0, nT = 0;
  const float* targetDu = nullptr; const float* targetDv = nullptr;   // caller-owned precomputed target-image gradient (gvv_set_target_gradient)
  int captured = 0;           // the handle has been used while its stream was being captured: the scratch pointers are baked into CUDA graphs
  int chain = 1;              // launch the kernels of a call as a programmatic dependent-launch chain
  int resolvePrefetch = 0;    // raster: L1 prefetch sweep of the resolve stage's vertex gathers (measured slower: 0.272 -> 0.282 ms)
  int sharedBatchGrads = 0;   // backward: vertex_color / texture / sh_coeff gradients are summed over the batch into [1, ...] outputs (parameters shared across the batch)
  int texBilinear = 0;        // non-default: bilinear texture fetch + weighted 4-texel gradient scatter (the variants the reference has commented out)
  int spreadEmpty = 0;        // raster: interleave the (HBM-bound) empty tiles with the (ALU-bound) non-empty ones
  int heavyMode = 1;          // 0 = never, 1 = only where such a bin would be the critical path of the launch (decided on the GPU), 2 = always
  int ctaSlots = 592;         // resident 256-thread raster CTAs of the device (4 per SM), set at create
  int heavySlots = 32;        // raster: heavy candidates per view = the first heavySlots items of its work list
  int heavyThr = 768;         // raster: bins of >= heavyThr triangles are candidates for the 1024-thread launch (one SM per tile)
  int splitUnit = 0;          // raster: a bin of >= splitUnit (2x, 4x) triangles is cut into 2 (4, 8) strips with a CTA each; 0 = never (measured slower: every strip re-scans the bin)
  int ctaTrace = 0;           // debug: record per-CTA start/end times of the raster kernel
  int spanZ = 2;              // raster: trim every row span to the pixels whose current winner the triangle could still beat (1 = both passes, 2 = far pass only)
  int hizMin = 64;            // raster: bins shorter than this are rasterised in one pass
  int hiz = 1;                // raster: two-pass hierarchical z (skips triangles behind the whole tile)
  int interleave = 1;         // raster: batch j takes bin entries j, j+nBatches, ... instead of a contiguous chunk
  int ctaThreads = 256;       // raster: threads per tile CTA (256 | 128)
  int batchDiv = 8;           // raster: a bin of n triangles is cut into batches of ceil(n / batchDiv) (<= 32) triangles
  int bwdPersistent = 0;      // (measured slower: 0.272 vs 0.218 ms) backward: persistent pixel-gradient kernel with a TMA face-tile ring (whole-tile images); 0 = one tile per CTA
  int bulkOut = 0;            // raster: write the tile's outputs through shared memory + TMA bulk copies (measured slower: 0.268 -> 0.289 ms)
  int rayCache = 0;           // raster: 1 = per-pixel ray cache in shared memory (3 CTAs/SM), 0 = recompute (4 CTAs/SM, measured faster)
  float cullMargin = 0.0625f; // px (fixed part of the margin); < 0 disables the conservative screen-space pre-test
  bool hasTexcoords = false;
  int4* faces4 = nullptr;     // [F] (v0,v1,v2,0)
  float* texcoords = nullptr; // [F*6]
  int* vfOffsets = nullptr;   // [N+1]  vertex -> incident faces (ascending face id), CSR
  int* vfList = nullptr;      // [vfOffsets[N]]
  // UV-space (face, a, b, c) table for compute_normal_map, built lazily per texture size
  float4* texelTable = nullptr; int tableH = 0, tableW = 0;
  gvv::ARParams ar = {};     // one-shot all-reduce of shared-parameter gradients at the end of gvv_backward (gvv_set_allreduce); blocks == 0: none
  int arAfter = 0;            // 1: the reduced range includes vertex_pos_grad -> own launch after the last kernel instead of CTAs inside it
  gvv::Scratch s;
  int64_t launches = 0;
  gvv::KernelTimer timer;
};

namespace gvv {

struct FwdArgs {
  int B, C, N, F, W, H, texH, texW, albedo, shading;
  int tile, tilesX, tilesY, nT, rayCache, batchDiv, ctaThreads, interleave, hiz, hizMin, spanZ, splitUnit, heavyThr, heavyMode, heavySlots, ctaSlots, spreadEmpty, texBilinear, resolvePrefetch, chain, bulkOut;
  float cullMargin;
  const float *vertex_pos, *vertex_color, *texture, *sh_coeff, *extrinsics, *intrinsics;
  const float* texcoords;
  const int4* faces4;
  const int *vfOffsets, *vfList;
  float* bary; int32_t* face; float* render; float* vertex_normal;
  Scratch s;
};

struct BwdArgs {
  int B, C, N, F, W, H, texH, texW, albedo, shading, imgFilter, texBilinear, chain, sharedBatch, bwdPersistent, ctaSlots;
  const float *render_grad, *target_grad, *vertex_pos, *vertex_color, *texture, *sh_coeff, *target_image,
      *vertex_normal, *bary, *extrinsics, *intrinsics, *texcoords, *target_du, *target_dv;
  const int32_t* face;
  const int4* faces4;
  const int *vfOffsets, *vfList;
  float *vpos_grad, *vcol_grad, *tex_grad, *sh_grad;
  ARParams ar; int arAfter;
  Scratch s;
};

// Programmatic dependent launch (sm_90+): every kernel of a call is launched with the programmatic-stream-
// serialization attribute and starts with chain_wait() (= wait until the preceding kernel of the stream has
// completed and flushed) followed by chain_trigger() (= the next kernel may be scheduled as soon as all CTAs of
// this one are resident or gone).  Semantics are those of plain stream order; what overlaps is the launch latency
// and CTA ramp-up of a kernel with the tail of its predecessor -- a few microseconds per boundary, which is what
// the small kernels of this pipeline cost.
__device__ __forceinline__ void chain_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void chain_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline cudaError_t launch_chained(bool chained, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = chained ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// cudaFuncSetAttribute is per DEVICE: the opt-in for more than 48 KB of dynamic shared memory is repeated on every
// device a handle lives on (one flag word per kernel, one bit per device; a process normally drives one GPU).
inline bool first_use_on_device(unsigned long long* flags) {
  int dev = 0;
  cudaGetDevice(&dev);
  const unsigned long long bit = 1ull << (dev & 63);
  if (*flags & bit) return false;
  *flags |= bit;
  return true;
}

// Each returns the number of kernels launched, or -1 after a launch error.
int launch_forward(const FwdArgs& a, cudaStream_t st, KernelTimer* tm);
int launch_normal_map(const FwdArgs& a, const float4* texelTable, float* normal_map, cudaStream_t st);
int launch_backward(const BwdArgs& a, cudaStream_t st, KernelTimer* tm);
int launch_debug_eval(const Scratch& s, const int4* faces4, int N, int C, int W, int H, int n, const int* dq, int* dkey, float* dab, cudaStream_t st);
int launch_build_texel_table(const float* texcoords, int F, int texH, int texW, float4* table, cudaStream_t st);

}  // namespace gvv
