// gvv_forward.cu -- forward pass of the rasteriser for sm_100a.
//
// Replaces the reference's eight per-batch-element launches (CUDABasedRasterization.cu:449-473)
// by six launches that cover ALL batch elements and cameras at once (a seventh, the heavy-tile instance of
// the raster kernel, when a call has at most two views), chained by programmatic dependent launch:
//
//   face_normal_kernel per (b, tri)    cross(v1-v0, v2-v0), once per batch element          (ref :122-141)
//   vertex_kernel      per (view, vtx) /1000 pre-scale, vertex normal via CSR (high-valence vertices summed by
//                                      the whole warp), exact projection, colour repack       (ref :98-174)
//   bin_count_kernel   per (view, tri) bbox (ref :184-208) -> tile histogram + per-tile range of the triangles'
//                                      depth-key lower bounds / big-triangle list
//   bin_scan_kernel    per view        exclusive scan, near/far threshold per tile, raster work list (heaviest
//                                      tile first), camera records E^-1, (K*E)^-1, ray origin  (ref :23-67)
//   bin_fill_kernel    per (view, tri) triangle ids into per-tile bins, near ones from the front, far from the back
//   raster_kernel      per (view, tile) 64-bit (depth|id) z-tile + winner barycentrics in shared memory resolved
//                                      with atomicMin semantics (one 128-bit CAS), then resolve + shade + write
//                                      of the outputs of the tile (ref :215-408, both raster passes and the
//                                      28 B/px clear of initializeDevice :74-91 fused away)
//
// No global z-buffer exists: the z-tile lives in shared memory, so HBM sees only the compulsory
// output stores (24 B/px) plus the (L2-resident) mesh reads.  camera_kernel is only used by the normal-map path.
#include <algorithm>
#include "gvv_internal.h"

namespace gvv {

#define FULL_MASK 0xffffffffu

// ------------------------------------------------------------------------------------------------
// camera_kernel
// ------------------------------------------------------------------------------------------------
__global__ void camera_kernel(const float* __restrict__ extr, const float* __restrict__ intr,
                              CamRec* __restrict__ cams, int* __restrict__ bigCount, int V) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= V) return;
  fill_camrec(extr, intr, cams, v);
  if (bigCount) bigCount[v] = 0;
}

// ------------------------------------------------------------------------------------------------
// vertex_kernel: one thread per (batch element, vertex)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ F3 ld3(const float* __restrict__ p, int i) {
  return mk3(__ldg(p + 3 * i), __ldg(p + 3 * i + 1), __ldg(p + 3 * i + 2));
}

// face normals cross(v1-v0, v2-v0), world space, once per batch element (ref :122-141 computes them C times)
__global__ void __launch_bounds__(256)
face_normal_kernel(const float* __restrict__ vertex_pos, const int4* __restrict__ faces4, float4* __restrict__ fnorm4, int N, int F) {
  chain_wait(); chain_trigger();
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (f >= F) return;
  const float* pos = vertex_pos + (size_t)b * N * 3;
  const int4 fc = __ldg(faces4 + f);
  const F3 a = ld3(pos, fc.x), bb = ld3(pos, fc.y), cc = ld3(pos, fc.z);
  const F3 fn = cross3x(sub3(bb, a), sub3(cc, a));
  fnorm4[(size_t)b * F + f] = make_float4(fn.x, fn.y, fn.z, 0.f);
}

// one thread per (view, vertex): the C cameras of a batch element run in parallel (the CSR gather of
// the normal is repeated per camera -- six 16-byte loads -- which is cheaper than serialising the
// C projections with their IEEE divides in one thread)
__global__ void __launch_bounds__(128)
vertex_kernel(const float* __restrict__ vertex_pos, const float* __restrict__ vertex_color,
              const float4* __restrict__ fnorm4, const int* __restrict__ vfOffsets, const int* __restrict__ vfList,
              const float* __restrict__ extr, const float* __restrict__ intr, int* __restrict__ bigCount,
              float4* __restrict__ proj, float4* __restrict__ vscaled,
              float4* __restrict__ vnorm4, float4* __restrict__ vcol4, float* __restrict__ vertex_normal_out,
              int N, int F, int C) {
  chain_wait(); chain_trigger();
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  const int view = blockIdx.y, b = view / C, c = view - b * C;
  if (blockIdx.x == 0 && threadIdx.x == 0) bigCount[view] = 0;   // consumed by bin_count_kernel (next launch)
  const bool valid = n < N;
  const int lane = threadIdx.x & 31;
  const float* pos = vertex_pos + (size_t)b * N * 3;
  // vertex normal = sum of incident face normals in ascending face order (ref :148-174); the
  // reference leaves vertices without faces uninitialised, we define them as 0.
  F3 nrm = mk3(0.f, 0.f, 0.f);
  int beg = 0, end = 0;
  if (valid) { beg = __ldg(vfOffsets + n); end = __ldg(vfOffsets + n + 1); }
  const float4* fnb = fnorm4 + (size_t)b * F;
  const bool fan = end - beg > 16;
  if (!fan) {
    // four incident faces at a time, so that their (dependent) loads are in flight together; summed in order
    for (int base = beg; base < end; base += 4) {
      int f[4];
      float4 fn[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) f[k] = __ldg(vfList + min(base + k, end - 1));
#pragma unroll
      for (int k = 0; k < 4; ++k) fn[k] = __ldg(fnb + f[k]);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (base + k < end) {
          if (base + k == beg) nrm = mk3(fn[k].x, fn[k].y, fn[k].z);
          else nrm = mk3(__fadd_rn(nrm.x, fn[k].x), __fadd_rn(nrm.y, fn[k].y), __fadd_rn(nrm.z, fn[k].z));
        }
      }
    }
  }
  // High-valence vertices (the poles of a UV sphere have degree ~segments) would serialise hundreds of
  // dependent loads in ONE thread and become the tail of the launch (measured: SMs busy 26 % of its
  // duration): the warp gathers 32 of their face normals at a time and adds them up in order via shuffles.
  unsigned fans = __ballot_sync(FULL_MASK, fan);
  while (fans) {
    const int L = __ffs(fans) - 1;
    fans &= fans - 1;
    const int hb = __shfl_sync(FULL_MASK, beg, L), he = __shfl_sync(FULL_MASK, end, L);
    F3 acc = mk3(0.f, 0.f, 0.f);
    for (int g0 = hb; g0 < he; g0 += 32 * 4) {          // four chunks of 32 incident faces in flight at a time
      float4 fn[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int i = g0 + 32 * c + lane;
        fn[c] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (i < he) fn[c] = __ldg(fnb + __ldg(vfList + i));
      }
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int c0 = g0 + 32 * c;
        const int cnt = min(32, he - c0);
        for (int k = 0; k < cnt; ++k) {
          const float fx = __shfl_sync(FULL_MASK, fn[c].x, k), fy = __shfl_sync(FULL_MASK, fn[c].y, k), fz = __shfl_sync(FULL_MASK, fn[c].z, k);
          if (c0 + k == hb) acc = mk3(fx, fy, fz);
          else acc = mk3(__fadd_rn(acc.x, fx), __fadd_rn(acc.y, fy), __fadd_rn(acc.z, fz));
        }
      }
    }
    if (lane == L) nrm = acc;
  }
  if (!valid) return;
  const F3 p = ld3(pos, n);
  float* vn = vertex_normal_out + ((size_t)view * N + n) * 3;
  vn[0] = nrm.x; vn[1] = nrm.y; vn[2] = nrm.z;
  proj[(size_t)view * N + n] = project_exact(intr + view * 9, extr + view * 12, p.x, p.y, p.z);
  if (c == 0) {
    vscaled[(size_t)b * N + n] = make_float4(__fdiv_rn(p.x, 1000.f), __fdiv_rn(p.y, 1000.f), __fdiv_rn(p.z, 1000.f), 0.f);
    vnorm4[(size_t)b * N + n] = make_float4(nrm.x, nrm.y, nrm.z, 0.f);
    if (vertex_color) {
      const F3 col = ld3(vertex_color + (size_t)b * N * 3, n);
      vcol4[(size_t)b * N + n] = make_float4(col.x, col.y, col.z, 0.f);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// binning
// ------------------------------------------------------------------------------------------------
// Lower bound of every depth key a triangle can produce.  z = 1/s, s = a/z0 + b/z1 + c/z2 with a,b,c in
// [-0.001, 1.001] and a+b+c = 1: the largest s puts weight 1.001 on the smallest depth and -0.001 on another,
// s <= 1.001/zmin, and with zmax <= 2 zmin the smallest s stays positive (>= (0.501 - 0.002)/zmin), so
// z >= zmin/1.001 = 0.999001 zmin.  The bound uses 0.9985 zmin (fp32 rounding of the five operations of
// depth_key_exact is ~1e-6 relative) and one key unit of slack; INT_MIN (= no bound) when the ratio test fails.
// Used only to SKIP work that provably cannot win the depth test.
__device__ __forceinline__ int key_lower_bound(float z0, float z1, float z2) {
  const float zmin = fminf(z0, fminf(z1, z2)), zmax = fmaxf(z0, fmaxf(z1, z2));
  return (zmin > 0.f && zmax <= 2.f * zmin) ? __float2int_rd(zmin * 9985.f) - 1 : (int)0x80000000;
}

struct TileRange { int tx0, ty0, tx1, ty1, n, klb; };

__device__ __forceinline__ TileRange tile_range(const int4* __restrict__ faces4, const float4* __restrict__ projv,
                                                int f, int W, int H, int tileShift) {
  const int4 fc = __ldg(faces4 + f);
  const float4 p0 = __ldg(projv + fc.x), p1 = __ldg(projv + fc.y), p2 = __ldg(projv + fc.z);
  const int4 bb = bbox_exact(p0, p1, p2, W, H);
  TileRange r;
  r.n = 0;
  r.klb = key_lower_bound(p0.z, p1.z, p2.z);
  if (bb.x > bb.z || bb.y > bb.w) { r.tx0 = r.ty0 = 0; r.tx1 = r.ty1 = -1; return r; }
  r.tx0 = bb.x >> tileShift; r.tx1 = bb.z >> tileShift;
  r.ty0 = bb.y >> tileShift; r.ty1 = bb.w >> tileShift;
  r.n = (r.tx1 - r.tx0 + 1) * (r.ty1 - r.ty0 + 1);
  return r;
}

// Each block bins kBinFacesPerThread * 256 consecutive triangles of one view, so that clearing and
// flushing the per-block tile histogram (nT entries) is amortised over 1024 triangles.
constexpr int kBinFacesPerThread = 4;

// The binning also prepares the two depth passes of the raster kernel ("hierarchical z"): per tile the range
// [min, max] of the depth-key lower bounds of its triangles (bin_count_kernel), the threshold in the middle of
// it (bin_scan_kernel), and bins filled with the NEAR triangles (lower bound <= threshold) from the front and
// the FAR ones from the back (bin_fill_kernel), so that the raster warps set every triangle up exactly once.
constexpr int kNoBound = (int)0x80000000;   // key_lower_bound of a triangle without a usable bound: always "near"

// Shared-memory histograms cover a WINDOW of the tile grid: the band of tile rows [ty0, ty1] that the block's
// kBinFacesPerThread * 256 consecutive triangles touch (the whole grid when it has at most kBinWindow tiles).  Meshes
// are stored spatially coherent, so the band is short; clearing and flushing it costs O(band) instead of O(nT) per
// block -- at 3840x2160 (8160 tiles) the three full-grid sweeps were 70 % of the kernel and the 98 / 130 KB of shared
// memory they needed left one block per SM.  A block whose band does not fit falls back to global atomics.
constexpr int kBinWindow = 2560;   // tiles per window: 30 KB (count) / 40 KB (fill) of shared memory per block

struct BinWindow { int t0, n; bool local; };   // first tile of the band, tiles in it, fits in shared memory

// r[k].n > 0 marks the triangles that go to the bins; ty0 / ty1 of those define the band (block-wide min / max)
template <bool WINDOWED>
__device__ __forceinline__ BinWindow bin_window(const TileRange* r, int tilesX, int tilesY, int nT, int* wb) {
  BinWindow w;
  if (!WINDOWED) { w.t0 = 0; w.n = nT; w.local = true; return w; }      // nT <= kBinWindow: the window is the whole grid
  if (threadIdx.x == 0) { wb[0] = 0x7fffffff; wb[1] = -1; }
  __syncthreads();
  int lo = 0x7fffffff, hi = -1;
#pragma unroll
  for (int k = 0; k < kBinFacesPerThread; ++k)
    if (r[k].n > 0 && r[k].n <= kMaxSmallTiles) { lo = min(lo, r[k].ty0); hi = max(hi, r[k].ty1); }
  lo = __reduce_min_sync(FULL_MASK, lo); hi = __reduce_max_sync(FULL_MASK, hi);
  if ((threadIdx.x & 31) == 0 && hi >= 0) { atomicMin(&wb[0], lo); atomicMax(&wb[1], hi); }
  __syncthreads();
  lo = wb[0]; hi = wb[1];
  (void)tilesY;
  w.t0 = hi >= 0 ? lo * tilesX : 0;
  w.n = hi >= 0 ? (hi - lo + 1) * tilesX : 0;
  w.local = w.n <= kBinWindow;
  return w;
}

template <bool WINDOWED>
__global__ void __launch_bounds__(256)
bin_count_kernel(const int4* __restrict__ faces4, const float4* __restrict__ proj, int* __restrict__ tileCount,
                 int* __restrict__ tileMinK, int* __restrict__ tileMaxK,
                 int* __restrict__ bigCount, int* __restrict__ bigList, int F, int N, int W, int H, int tileShift,
                 int tilesX, int tilesY, int nT) {
  chain_wait(); chain_trigger();
  extern __shared__ int hist[];   // hist[cap], mn[cap], mx[cap], cap = min(nT, kBinWindow)
  __shared__ int wb[2];
  const int cap = min(nT, kBinWindow);
  int* mn = hist + cap;
  int* mx = hist + 2 * cap;
  const int view = blockIdx.y;
  TileRange r[kBinFacesPerThread];
  if (WINDOWED) {      // the band must be known before the first add: all tile ranges first
#pragma unroll
    for (int k = 0; k < kBinFacesPerThread; ++k) {
      const int f = (blockIdx.x * kBinFacesPerThread + k) * blockDim.x + threadIdx.x;
      r[k].n = 0; r[k].tx0 = r[k].ty0 = 0; r[k].tx1 = r[k].ty1 = -1; r[k].klb = kNoBound;
      if (f < F) r[k] = tile_range(faces4, proj + (size_t)view * N, f, W, H, tileShift);
    }
  }
  const BinWindow w = bin_window<WINDOWED>(r, tilesX, tilesY, nT, wb);
  int* gCount = tileCount + (size_t)view * nT;
  int* gMin = tileMinK + (size_t)view * nT;
  int* gMax = tileMaxK + (size_t)view * nT;
  if (w.local) {
    for (int i = threadIdx.x; i < w.n; i += blockDim.x) { hist[i] = 0; mn[i] = 0x7fffffff; mx[i] = kNoBound; }
    __syncthreads();
  }
#pragma unroll
  for (int k = 0; k < kBinFacesPerThread; ++k) {
    const int f = (blockIdx.x * kBinFacesPerThread + k) * blockDim.x + threadIdx.x;
    if (f >= F) continue;
    const TileRange q = WINDOWED ? r[k] : tile_range(faces4, proj + (size_t)view * N, f, W, H, tileShift);
    if (q.n > kMaxSmallTiles) {
      const int slot = atomicAdd(bigCount + view, 1);
      bigList[(size_t)view * F + slot] = f;
    } else if (q.n > 0) {
      for (int ty = q.ty0; ty <= q.ty1; ++ty)
        for (int tx = q.tx0; tx <= q.tx1; ++tx) {
          const int t = ty * tilesX + tx;
          if (w.local) {
            atomicAdd(&hist[t - w.t0], 1);
            if (q.klb != kNoBound) { atomicMin(&mn[t - w.t0], q.klb); atomicMax(&mx[t - w.t0], q.klb); }
          } else {
            atomicAdd(gCount + t, 1);
            if (q.klb != kNoBound) { atomicMin(gMin + t, q.klb); atomicMax(gMax + t, q.klb); }
          }
        }
    }
  }
  if (w.local) {
    __syncthreads();
    for (int i = threadIdx.x; i < w.n; i += blockDim.x) {
      const int c = hist[i];
      if (c) {
        atomicAdd(gCount + w.t0 + i, c);
        if (mx[i] != kNoBound) { atomicMin(gMin + w.t0 + i, mn[i]); atomicMax(gMax + w.t0 + i, mx[i]); }
      }
    }
  }
}

// exclusive scan of one view's tile histogram (one block per view) + the raster work list of the view.
// A work item is a horizontal STRIP of a tile: item = tx | ty << 12 | strip << 24 | log2(K) << 27 | heavy << 29
// (tile coordinates, so that the raster CTA needs no integer division), the tile being cut
// into K = 1, 2, 4 or 8 strips of TS/K rows, each rasterised by its own CTA (every strip scans the whole
// bin of the tile and keeps the rows that fall into it).  Heavy bins are split so that the longest CTA
// of the launch -- the critical path: a pole/silhouette tile holds ~15x the average bin -- shrinks; the
// items are ordered heaviest first (counting sort by log2 of bin size / K: longest-processing-time-first).
// The list has nItems = nT + nT/2 slots per view; unused slots hold -1.
__device__ __forceinline__ int strip_log2(int cnt, int unit, int maxLog) {
  int l = 0;
  while (l < maxLog && cnt >= (unit << l)) ++l;      // cnt >= unit -> 2 strips, >= 2 unit -> 4, >= 4 unit -> 8
  return l;
}

__global__ void __launch_bounds__(1024) bin_scan_kernel(const int* __restrict__ tileCount, int* __restrict__ tileOffset,
                                                        int* __restrict__ tileOrder, int* __restrict__ tileMinK, int* __restrict__ tileMaxK, int* __restrict__ tileThr,
                                                        int nT, int tilesX, int nItems, int splitUnit, int maxLog,
                                                        int heavyThr, int heavySlots, int heavyLoad, int spreadEmpty,
                                                        const float* __restrict__ extr, const float* __restrict__ intr,
                                                        CamRec* __restrict__ cams, int numViews) {
  chain_wait(); chain_trigger();
  // blocks beyond the views: the camera records (E^-1, (KE)^-1, ray origin) of all views, one thread each, beside
  // the scans instead of after them -- the reference spends a <<<1,1>>> launch with a serial loop on this
  if ((int)blockIdx.x >= numViews) {
    const int v = ((int)blockIdx.x - numViews) * blockDim.x + threadIdx.x;
    if (v < numViews) fill_camrec(extr, intr, cams, v);
    return;
  }
  extern __shared__ int cnt[];   // the view's tile histogram, read four times below
  __shared__ int warpSum[32];
  __shared__ int carry;
  __shared__ int bucketStart[33], bucketFill[33];
  __shared__ int extraItems;
  const int view = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) carry = 0;
  for (int i = threadIdx.x; i < nT; i += blockDim.x) cnt[i] = tileCount[(size_t)view * nT + i];
  __syncthreads();
  for (int base = 0; base < nT; base += blockDim.x) {
    const int i = base + threadIdx.x;
    const int v = (i < nT) ? cnt[i] : 0;
    int x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(FULL_MASK, x, o); if (lane >= o) x += y; }
    if (lane == 31) warpSum[warp] = x;
    __syncthreads();
    if (warp == 0) {
      int w = warpSum[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(FULL_MASK, w, o); if (lane >= o) w += y; }
      warpSum[lane] = w;
    }
    __syncthreads();
    const int prefix = carry + (warp ? warpSum[warp - 1] : 0) + x - v;
    if (i < nT) tileOffset[(size_t)view * nT + i] = prefix;
    __syncthreads();
    if (threadIdx.x == 0) carry += warpSum[31];
    __syncthreads();
  }
  // near/far threshold of every tile = middle of the range of its depth-key lower bounds; the range is reset
  // for the next call (self-cleaning scratch)
  for (int i = threadIdx.x; i < nT; i += blockDim.x) {
    const size_t ti = (size_t)view * nT + i;
    const int lo = tileMinK[ti], hi = tileMaxK[ti];
    tileThr[ti] = hi > lo ? lo + ((hi - lo) >> 1) : 0x7fffffff;
    tileMinK[ti] = 0x7fffffff; tileMaxK[ti] = kNoBound;
  }
  // strips per tile: double the split unit until the extra items fit into the nItems - nT spare slots
  int unit = splitUnit;
  for (int attempt = 0; attempt < 8; ++attempt) {
    if (threadIdx.x == 0) extraItems = 0;
    __syncthreads();
    int extra = 0;
    for (int i = threadIdx.x; i < nT; i += blockDim.x)
      extra += (1 << strip_log2(cnt[i], unit, maxLog)) - 1;
    if (extra) atomicAdd(&extraItems, extra);
    __syncthreads();
    const int total = extraItems;
    __syncthreads();
    if (total <= nItems - nT) break;
    unit = attempt < 6 ? unit * 2 : 0x7fffffff;      // last resort: no split at all
  }
  // counting sort of the items by floor(log2(bin size / K)) (33 buckets, 32 = heaviest ... 0 = empty)
  if (threadIdx.x < 33) { bucketStart[threadIdx.x] = 0; bucketFill[threadIdx.x] = 0; }
  __syncthreads();
  for (int i = threadIdx.x; i < nT; i += blockDim.x) {
    const int c = cnt[i];
    const int l = strip_log2(c, unit, maxLog), w = c >> l;
    atomicAdd(&bucketStart[w > 0 ? 32 - __clz(w) : 0], 1 << l);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int run = 0;
    for (int k = 32; k >= 0; --k) { const int c = bucketStart[k]; bucketStart[k] = run; run += c; }
    extraItems = run;                                  // = number of items of this view
  }
  __syncthreads();
  int* order = tileOrder + (size_t)view * nItems;
  for (int i = threadIdx.x; i < nT; i += blockDim.x) {
    const int c = cnt[i];
    const int l = strip_log2(c, unit, maxLog), w = c >> l;
    const int k = w > 0 ? 32 - __clz(w) : 0;
    const int pos0 = bucketStart[k] + atomicAdd(&bucketFill[k], 1 << l);
    // Empty tiles (bucket 0; pure 24 B/px background stores, HBM-bound) are spread evenly between the non-empty
    // ones (ALU-bound) instead of forming the tail of the launch: non-empty i moves to i + floor(i n0/n1),
    // empty j fills the gaps -- a bijection on [0, n0 + n1) that keeps the heaviest-first order of the non-empty items.
    const int n1 = bucketStart[0], n0 = extraItems - n1;     // non-empty / empty items of this view
    for (int sidx = 0; sidx < (1 << l); ++sidx) {
      int pos = pos0 + sidx;
      if (spreadEmpty && n0 > 0 && n1 > 0) {
        if (k > 0) pos += (int)(((long long)pos * n0) / n1);
        else { const int j = pos - n1; pos = j + min(n1, (int)((((long long)(j + 1)) * n1 + n0 - 1) / n0)); }
      }
      // bit 29: the item is rasterised by the 1024-thread launch (one whole SM per tile) instead of a 256-thread CTA
      // ... when it would otherwise be the critical path of the launch: its bin is more than 1.4x the average
      // load of a 256-thread CTA slot (heavyLoad = slots / views, carry = bin entries of this view; heavyLoad < 0: always)
      const bool critical = heavyLoad < 0 || 5ll * w * heavyLoad > 7ll * carry;
      const int heavy = (heavyThr > 0 && w >= heavyThr && pos < heavySlots && critical) ? (1 << 29) : 0;
      order[pos] = (i % tilesX) | ((i / tilesX) << 12) | (sidx << 24) | (l << 27) | heavy;
    }
  }
  for (int i = extraItems + threadIdx.x; i < nItems; i += blockDim.x) order[i] = -1;
}

template <bool WINDOWED>
__global__ void __launch_bounds__(256)
bin_fill_kernel(const int4* __restrict__ faces4, const float4* __restrict__ proj, const int* __restrict__ tileOffset,
                const int* __restrict__ tileCount, const int* __restrict__ tileThr,
                int* __restrict__ tileCursor, int* __restrict__ tileCursorFar, int* __restrict__ bins, int F, int N, int W, int H, int tileShift,
                int tilesX, int tilesY, int nT) {
  chain_wait(); chain_trigger();
  extern __shared__ int sm[];   // histN[cap], histF[cap], baseN[cap], baseF[cap] over the block's window (see bin_window), cap = min(nT, kBinWindow)
  __shared__ int wb[2];
  const int cap = min(nT, kBinWindow);
  int* histN = sm;
  int* histF = sm + cap;
  int* baseN = sm + 2 * cap;
  int* baseF = sm + 3 * cap;
  const int view = blockIdx.y;
  TileRange r[kBinFacesPerThread];
  int fid[kBinFacesPerThread];
#pragma unroll
  for (int k = 0; k < kBinFacesPerThread; ++k) {
    fid[k] = (blockIdx.x * kBinFacesPerThread + k) * blockDim.x + threadIdx.x;
    r[k].n = 0; r[k].tx0 = r[k].ty0 = 0; r[k].tx1 = r[k].ty1 = -1; r[k].klb = kNoBound;
    if (fid[k] < F) r[k] = tile_range(faces4, proj + (size_t)view * N, fid[k], W, H, tileShift);
    if (r[k].n > kMaxSmallTiles) r[k].n = 0;      // big triangles live in the big list (bin_count_kernel)
  }
  const BinWindow w = bin_window<WINDOWED>(r, tilesX, tilesY, nT, wb);
  int* viewBins = bins + (size_t)view * F * kMaxSmallTiles;
  const int* off = tileOffset + (size_t)view * nT;
  const int* cnt = tileCount + (size_t)view * nT;
  const int* thr = tileThr + (size_t)view * nT;
  int* curN = tileCursor + (size_t)view * nT;
  int* curF = tileCursorFar + (size_t)view * nT;
  // a triangle is FAR in a tile when its depth-key lower bound is beyond the tile's threshold
  auto is_far = [&](int klb, int t) { return klb != kNoBound && klb > __ldg(thr + t); };
  if (w.local) {
    for (int i = threadIdx.x; i < w.n; i += blockDim.x) { histN[i] = 0; histF[i] = 0; }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kBinFacesPerThread; ++k)
      if (r[k].n > 0)
        for (int ty = r[k].ty0; ty <= r[k].ty1; ++ty)
          for (int tx = r[k].tx0; tx <= r[k].tx1; ++tx) {
            const int t = ty * tilesX + tx;
            atomicAdd(is_far(r[k].klb, t) ? &histF[t - w.t0] : &histN[t - w.t0], 1);
          }
    __syncthreads();
    // one global reservation per tile and side this block touches: near entries grow from the front of the
    // tile's segment, far entries from its back
    for (int i = threadIdx.x; i < w.n; i += blockDim.x) {
      const int cN = histN[i], cF = histF[i], t = w.t0 + i;
      if (cN) { baseN[i] = off[t] + atomicAdd(curN + t, cN); histN[i] = 0; }
      if (cF) { baseF[i] = off[t] + cnt[t] - atomicAdd(curF + t, cF) - cF; histF[i] = 0; }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kBinFacesPerThread; ++k)
      if (r[k].n > 0)
        for (int ty = r[k].ty0; ty <= r[k].ty1; ++ty)
          for (int tx = r[k].tx0; tx <= r[k].tx1; ++tx) {
            const int t = ty * tilesX + tx, i = t - w.t0;
            if (is_far(r[k].klb, t)) viewBins[baseF[i] + atomicAdd(&histF[i], 1)] = fid[k];
            else viewBins[baseN[i] + atomicAdd(&histN[i], 1)] = fid[k];
          }
  } else {
#pragma unroll
    for (int k = 0; k < kBinFacesPerThread; ++k)
      if (r[k].n > 0)
        for (int ty = r[k].ty0; ty <= r[k].ty1; ++ty)
          for (int tx = r[k].tx0; tx <= r[k].tx1; ++tx) {
            const int t = ty * tilesX + tx;
            if (is_far(r[k].klb, t)) viewBins[off[t] + cnt[t] - 1 - atomicAdd(curF + t, 1)] = fid[k];
            else viewBins[off[t] + atomicAdd(curN + t, 1)] = fid[k];
          }
  }
}

// ------------------------------------------------------------------------------------------------
// raster_kernel
// ------------------------------------------------------------------------------------------------
// Per-(triangle, tile) records staged in shared memory.
// TriRec: everything the exact hit test needs (five float4).
struct __align__(16) TriRec {
  float v0x, v0y, v0z, v1x;
  float v1y, v1z, v2x, v2y;
  float v2z, Nx, Ny, Nz;
  float num, den, z0, z1;
  float z2; int face; int pad0; int pad1;
};
static_assert(sizeof(TriRec) == 80, "TriRec is 80 bytes");
// EdgeRec: fragment decode + the conservative screen-space pre-test (three float4).
struct __align__(16) EdgeRec {
  float A0, B0, C0, A1;
  float B1, C1, A2, B2;
  float C2, R0, R1, R2;                      // R_i = -1/A_i (0 when the edge is treated as horizontal)
  int start; int geom; int pad0; int pad1;   // start = first ROW of this triangle in the chunk's row list;
};                                           // geom = x0 | y0 << 8 | w << 16 (tile-local)
static_assert(sizeof(EdgeRec) == 64, "EdgeRec is 64 bytes");

// Conservative screen-space reject.  The reference tests EVERY pixel of the bbox with the exact
// 3-D test; a pair that fails it has no effect at all, so pairs that provably fail may be skipped.
// Under the pinhole map the 3-D triangle projects exactly onto the 2-D triangle of the projected
// vertices, and the reference's inside test admits perspective barycentrics >= -0.001, i.e. screen
// distances up to 0.001 * (zmax/zmin) * height outside an edge.  We keep every pixel centre within
//     m = margin + 0.002 * (zmax/zmin) * (longest edge)       (2x that bound + `margin` for rounding;
//                                                              margin = 1/16 px by default, while the fp32
//                                                              error of the pixel position is ~1e-3 px)
// of the triangle and send it to the exact test; only pixels farther out are dropped.  Triangles
// whose projection is unreliable (a vertex at/behind the camera plane, depth ratio > 2, non-finite
// coordinates) are not culled at all; thin triangles (|area| < 1 px^2, orientation ambiguous) use a
// two-sided band around their longest edge.  e_i(lx,ly) = A_i*lx + B_i*ly + C_i >= 0 keeps the pixel.
// Row form of the pre-test: on row ly, e_i >= 0 <=> x >= t_i (A_i > 0) or x <= t_i (A_i < 0) with
// t_i = (B_i*ly + C_i) * R_i, R_i = -1/A_i.  An edge with |A_i| * 32 below 1e-4 of its own slack is
// treated as horizontal (A_i := 0, its 32*|A_i| variation across the tile is added to C_i).
__device__ __forceinline__ void edge_finish(EdgeRec& e) {
  float* A[3] = {&e.A0, &e.A1, &e.A2};
  float* B[3] = {&e.B0, &e.B1, &e.B2};
  float* C[3] = {&e.C0, &e.C1, &e.C2};
  float* R[3] = {&e.R0, &e.R1, &e.R2};
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const float a = *A[i];
    if (fabsf(a) * 64.f <= 1.0e-3f * fabsf(*B[i]) || a == 0.f) {
      *C[i] += 64.f * fabsf(a);
      *A[i] = 0.f;
      *R[i] = 0.f;
    } else {
      *R[i] = -1.f / a;
    }
  }
}

__device__ __forceinline__ void edge_setup(float4 p0, float4 p1, float4 p2, float ox, float oy, float margin, EdgeRec& e) {
  const float x0 = p0.x - ox, y0 = p0.y - oy, x1 = p1.x - ox, y1 = p1.y - oy, x2 = p2.x - ox, y2 = p2.y - oy;
  const float zmin = fminf(p0.z, fminf(p1.z, p2.z)), zmax = fmaxf(p0.z, fmaxf(p1.z, p2.z));
  e.A0 = e.B0 = e.A1 = e.B1 = e.A2 = e.B2 = 0.f;
  e.R0 = e.R1 = e.R2 = 0.f;
  e.C0 = e.C1 = e.C2 = 1.f;                                   // default: keep everything
  if (!(margin >= 0.f) || !(zmin > 1.0e-4f) || !(zmax <= 2.f * zmin)) return;   // margin < 0: culling off
  const float ax = x1 - x0, ay = y1 - y0, bx = x2 - x1, by = y2 - y1, cx = x0 - x2, cy = y0 - y2;
  const float area2 = ax * (y2 - y0) - ay * (x2 - x0);
  const float la = sqrtf(ax * ax + ay * ay), lb = sqrtf(bx * bx + by * by), lc = sqrtf(cx * cx + cy * cy);
  const float lmax = fmaxf(la, fmaxf(lb, lc));
  const float m = margin + 0.002f * (zmax / zmin) * lmax;
  if (!(lmax < 1.0e6f) || !(fabsf(area2) < 1.0e12f)) return;   // also rejects NaN / inf
  // E(x,y) = dx*(y - ya) - dy*(x - xa) for the edge a->b: A = -dy, B = dx, C = dy*xa - dx*ya; pixel centre = (lx+.5, ly+.5)
  if (fabsf(area2) >= 1.f) {
    const float s = area2 > 0.f ? 1.f : -1.f;
    e.A0 = -s * ay; e.B0 = s * ax; e.C0 = s * (ay * x0 - ax * y0) + 0.5f * (e.A0 + e.B0) + m * la;
    e.A1 = -s * by; e.B1 = s * bx; e.C1 = s * (by * x1 - bx * y1) + 0.5f * (e.A1 + e.B1) + m * lb;
    e.A2 = -s * cy; e.B2 = s * cx; e.C2 = s * (cy * x2 - cx * y2) + 0.5f * (e.A2 + e.B2) + m * lc;
  } else {
    // thin: |E_long| <= |area2| inside the triangle, so keep |E_long| <= |area2| + m*len
    float dx = ax, dy = ay, xa = x0, ya = y0, len = la;
    if (lb >= la && lb >= lc) { dx = bx; dy = by; xa = x1; ya = y1; len = lb; }
    else if (lc >= la && lc >= lb) { dx = cx; dy = cy; xa = x2; ya = y2; len = lc; }
    const float A = -dy, B = dx, C = dy * xa - dx * ya + 0.5f * (A + B), T = fabsf(area2) + m * len;
    e.A0 = A; e.B0 = B; e.C0 = C + T;
    e.A1 = -A; e.B1 = -B; e.C1 = -C + T;
  }
  edge_finish(e);
}

struct RasterParams {
  const int4* faces4; const float4* proj; const float4* vscaled; const float4* vnorm4; const float4* vcol4;
  const CamRec* cams;
  int* tileCount; int* tileCursor; int* tileCursorFar; int* tileDone; const int* tileOffset; const int* tileOrder; const int* bigCount; const int* bigList; const int* bins;
  const float* texture; const float* texcoords; const float* sh_coeff;
  float* bary; int32_t* face; float* render;
  unsigned long long* ctaTrace;
  int C, N, F, W, H, texH, texW, albedo, shading, tilesX, nT, nItems, V, batchDiv, interleave, hiz, hizMin, spanZ, role, pdl, grid2d, texBilinear, resolvePrefetch, bulkOut;
  float cullMargin, invC;
};

__device__ __forceinline__ TriSetup load_setup(const RasterParams& p, int b, int view, int4 fc, F3 ros,
                                               float& z0, float& z1, float& z2, int4* bb) {
  const float4* vs = p.vscaled + (size_t)b * p.N;
  const float4* pj = p.proj + (size_t)view * p.N;
  const float4 s0 = __ldg(vs + fc.x), s1 = __ldg(vs + fc.y), s2 = __ldg(vs + fc.z);
  const float4 p0 = __ldg(pj + fc.x), p1 = __ldg(pj + fc.y), p2 = __ldg(pj + fc.z);
  z0 = p0.z; z1 = p1.z; z2 = p2.z;
  if (bb) *bb = bbox_exact(p0, p1, p2, p.W, p.H);
  return tri_setup_exact(mk3(s0.x, s0.y, s0.z), mk3(s1.x, s1.y, s1.z), mk3(s2.x, s2.y, s2.z), ros);
}

// spherical-harmonics irradiance * albedo: getShading (RendererUtil.h:135-173)
__device__ __forceinline__ float sh_eval(const float* __restrict__ sh, F3 n) {
  float s = sh[0];
  s = __fmaf_rn(n.y, sh[1], s);
  s = __fmaf_rn(n.z, sh[2], s);
  s = __fmaf_rn(n.x, sh[3], s);
  s = __fmaf_rn(__fmul_rn(n.x, n.y), sh[4], s);
  s = __fmaf_rn(__fmul_rn(n.z, n.y), sh[5], s);
  s = __fmaf_rn(__fmaf_rn(__fmul_rn(n.z, n.z), 3.f, -1.f), sh[6], s);
  s = __fmaf_rn(__fmul_rn(n.x, n.z), sh[7], s);
  s = __fmaf_rn(__fmaf_rn(n.x, n.x, -__fmul_rn(n.y, n.y)), sh[8], s);
  return s;
}

// z-tile entry: the 64-bit (depth|id) key plus the winner's barycentrics, updated together by ONE
// 128-bit shared-memory compare-and-swap (ATOMS.CAS.128, sm_90+), so that the resolve stage reads
// (a,b) instead of re-running triangle setup and the exact test for every covered pixel.
struct __align__(16) ZEntry { unsigned long long key; float a, b; };

__device__ __forceinline__ ZEntry cas128_shared(ZEntry* addr, ZEntry cmp, ZEntry val) {
  ZEntry old;
  unsigned long long oab;
  const unsigned sa = (unsigned)__cvta_generic_to_shared(addr);
  const unsigned long long cab = ((unsigned long long)__float_as_uint(cmp.b) << 32) | __float_as_uint(cmp.a);
  const unsigned long long vab = ((unsigned long long)__float_as_uint(val.b) << 32) | __float_as_uint(val.a);
  asm volatile("{\n\t.reg .b128 c, v, o;\n\tmov.b128 c, {%3, %4};\n\tmov.b128 v, {%5, %6};\n\t"
               "atom.shared.cas.b128 o, [%2], c, v;\n\tmov.b128 {%0, %1}, o;\n\t}"
               : "=l"(old.key), "=l"(oab) : "r"(sa), "l"(cmp.key), "l"(cab), "l"(val.key), "l"(vab) : "memory");
  old.a = __uint_as_float((unsigned)oab);
  old.b = __uint_as_float((unsigned)(oab >> 32));
  return old;
}

// Shared-memory plan of raster_kernel (dynamic, carved by hand):
//   zt[TS*TS] ZEntry | rayx,rayy,rayz[TS*TS] f32 (RC only) | per warp: rec[kBatch], erec[kBatch], startArr[60], qStart[96], qInfo[64]
constexpr int kBatch = 24;   // triangles a warp sets up at a time
constexpr int kQStart = 96;  // span queue: <= 31 spans left over + 32 new ones, + 33 sentinels for the 32-wide search
constexpr int kQInfo = 64;
constexpr int kWarpSmemBytes = kBatch * (int)sizeof(TriRec) + kBatch * (int)sizeof(EdgeRec) + (60 + kQStart + kQInfo) * 4;
// RC = per-pixel ray cache in shared memory (3 CTAs/SM) or rays recomputed per use (4 CTAs/SM, <= 64 registers)
// NTH = threads per tile CTA (256 or 128)
template <int TS, bool RC, int NTH>
constexpr int raster_smem_bytes() { return TS * TS * (16 + (RC ? 12 : 0)) + (NTH / 32) * kWarpSmemBytes; }

template <int TS, bool RC, int NTH>
__global__ void __launch_bounds__(NTH, (RC ? 3 : 4) * (256 / NTH))
raster_kernel(const RasterParams p) {
  constexpr int NPIX = TS * TS;
  constexpr int ZRAY = NPIX * (16 + (RC ? 12 : 0));
  extern __shared__ __align__(16) unsigned char smemRaw[];
  ZEntry* zt = reinterpret_cast<ZEntry*>(smemRaw);
  float* rayx = reinterpret_cast<float*>(smemRaw + NPIX * 16);
  float* rayy = rayx + NPIX;
  float* rayz = rayy + NPIX;
  __shared__ float shc[27];
  __shared__ CamRec cam;
  __shared__ int nextBatch;
  __shared__ unsigned sZmax;
  __shared__ int sPlan[2];      // batch size and batch count of the current pass: two integer divisions by run-time values, done by ONE thread

  // Grid (V, items): x = view runs fastest, so the heaviest work items of every view are scheduled first
  // (1-D fallback with a division when the work list is longer than gridDim.y allows).
  // item = tx | ty << 12 | strip << 24 | log2(K) << 27 | heavy << 29: this CTA owns rows [rowLo, rowLo + rowN) of tile (tx, ty)
  const int view = p.grid2d ? (int)blockIdx.x : (int)(blockIdx.x % p.V);
  const int rank = p.grid2d ? (int)blockIdx.y : (int)(blockIdx.x / p.V);
  // heavy launch (role 1): wait for the binning, then let the 256-thread launch start beside this one.  The
  // 256-thread launch waits for its predecessor itself unless that is the heavy launch (p.pdl), whose CTAs
  // only trigger after their own wait -- so the binning is complete either way.
  if (p.role == 1 || !p.pdl) chain_wait();
  chain_trigger();
  const int item = p.tileOrder[(size_t)view * p.nItems + rank];
  // The 256-thread launch may FINISH before the heavy one; whatever follows in the stream only waits for this
  // launch, so its last CTA (scheduled last) does not leave before the heavy launch has completed and flushed.
  if (p.role == 0 && p.pdl && blockIdx.x == gridDim.x - 1 && blockIdx.y == gridDim.y - 1) asm volatile("griddepcontrol.wait;" ::: "memory");
  if (item < 0) return;                                // spare slot of the work list
  if (((item >> 29) & 1) != p.role) return;            // heavy items belong to the 1024-thread launch (role 1), the rest to role 0
  const int tileX = item & 0xfff, tileY = (item >> 12) & 0xfff, tile = tileY * p.tilesX + tileX;
  const int stripLog = (item >> 27) & 3, rowN = TS >> stripLog, rowLo = ((item >> 24) & 7) * rowN;
  const int qLo = rowLo * TS, qHi = (rowLo + rowN) * TS;
  const int b = (int)(((float)view + 0.5f) * p.invC);   // = view / C without the integer division (exact below 2^22 views)
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tileX0 = tileX * TS, tileY0 = tileY * TS;
  const size_t tidx = (size_t)view * p.nT + tile;
  const int cntSmall = p.tileCount[tidx];
  const int cntBig = p.bigCount[view];
  const size_t pixBase = (size_t)view * p.W * p.H;
  if (p.ctaTrace && tid == 0) {
    unsigned long long t; unsigned smid;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    const size_t lin = (size_t)rank * p.V + view;
    p.ctaTrace[4 * lin] = t; p.ctaTrace[4 * lin + 1] = t;
    p.ctaTrace[4 * lin + 2] = (unsigned long long)(cntSmall + cntBig); p.ctaTrace[4 * lin + 3] = smid;
  }

  if (cntSmall == 0 && cntBig == 0) {
    // empty tile: background only (face -1, bary 0, render (0,1,0): initializeDevice :80-89).
    // Full-width tiles of 16-byte aligned images are written with 128-bit stores (6 per thread instead of 20).
    const bool vec = ((p.W & 3) == 0) && tileX0 + TS <= p.W &&
                     ((reinterpret_cast<uintptr_t>(p.face) | reinterpret_cast<uintptr_t>(p.bary) | reinterpret_cast<uintptr_t>(p.render)) & 15) == 0;
    if (vec) {
      const int rowsValid = min(TS, p.H - tileY0);
      const size_t rowBase = pixBase + (size_t)tileY0 * p.W + tileX0;
      for (int i = tid; i < rowsValid * (TS / 4); i += NTH) {
        const int r = i / (TS / 4), c = i % (TS / 4);
        __stcs(reinterpret_cast<int4*>(p.face + rowBase + (size_t)r * p.W) + c, make_int4(-1, -1, -1, -1));
      }
      for (int i = tid; i < rowsValid * (TS / 2); i += NTH) {
        const int r = i / (TS / 2), c = i % (TS / 2);
        __stcs(reinterpret_cast<float4*>(p.bary + 2 * (rowBase + (size_t)r * p.W)) + c, make_float4(0.f, 0.f, 0.f, 0.f));
      }
      for (int i = tid; i < rowsValid * (3 * TS / 4); i += NTH) {
        const int r = i / (3 * TS / 4), c = i % (3 * TS / 4), m = c % 3;   // (0,1,0,0) (1,0,0,1) (0,0,1,0)
        __stcs(reinterpret_cast<float4*>(p.render + 3 * (rowBase + (size_t)r * p.W)) + c,
               make_float4(m == 1 ? 1.f : 0.f, m == 0 ? 1.f : 0.f, m == 2 ? 1.f : 0.f, m == 1 ? 1.f : 0.f));
      }
      return;
    }
    for (int q = tid; q < NPIX; q += NTH) {
      const int x = tileX0 + (q % TS), y = tileY0 + (q / TS);
      if (x < p.W && y < p.H) {
        const size_t pix = pixBase + (size_t)y * p.W + x;
        __stcs(p.face + pix, -1);
        __stcs(reinterpret_cast<float2*>(p.bary) + pix, make_float2(0.f, 0.f));
        __stcs(p.render + 3 * pix + 0, 0.f); __stcs(p.render + 3 * pix + 1, 1.f); __stcs(p.render + 3 * pix + 2, 0.f);
      }
    }
    return;
  }

  if (tid < 64) reinterpret_cast<float*>(&cam)[tid] = reinterpret_cast<const float*>(p.cams + view)[tid];
  if (tid >= 64 && tid < 64 + 27) shc[tid - 64] = p.sh_coeff[(size_t)view * 27 + (tid - 64)];
  const int nNear = p.tileCursor[tidx];
  const bool twoPass = p.hiz && cntSmall + cntBig >= p.hizMin && nNear < cntSmall;
  if (tid == 96) {
    nextBatch = 0; sZmax = 0u;
    const int nPass0 = (twoPass ? nNear : cntSmall) + cntBig;
    const int G0 = min(kBatch, max(1, (nPass0 + p.batchDiv - 1) / p.batchDiv));
    sPlan[0] = G0; sPlan[1] = (nPass0 + G0 - 1) / G0;
  }
  __syncthreads();
  const F3 ros = mk3(cam.ros[0], cam.ros[1], cam.ros[2]);

  // z-tile clear + per-pixel ray cache (the ray depends on pixel and camera only)
  for (int q = qLo + tid; q < qHi; q += NTH) {
    { ZEntry e; e.key = kEmptyKey; e.a = 0.f; e.b = 0.f; zt[q] = e; }
    const int x = tileX0 + (q % TS), y = tileY0 + (q / TS);
    if (RC) {
      const F3 rd = ray_dir_exact(cam.Pinv, cam.ro, __fadd_rn((float)x, 0.5f), __fadd_rn((float)y, 0.5f));
      rayx[q] = rd.x; rayy[q] = rd.y; rayz[q] = rd.z;
    }
  }
  __syncthreads();
  auto ray_of = [&](int q) {
    if (RC) return mk3(rayx[q], rayy[q], rayz[q]);
    return ray_dir_exact(cam.Pinv, cam.ro, __fadd_rn((float)(tileX0 + (q % TS)), 0.5f), __fadd_rn((float)(tileY0 + (q / TS)), 0.5f));
  };

  // ---- rasterise.  Every warp works on its own: it grabs a batch of up to 32 triangles of the
  // tile's bin (then of the view's big-triangle list), sets them up into its private records, and
  // rasterises them; no block-wide barrier until all batches are done. ----
  unsigned char* wbase = smemRaw + ZRAY + warp * kWarpSmemBytes;
  TriRec* rec = reinterpret_cast<TriRec*>(wbase);
  EdgeRec* erec = reinterpret_cast<EdgeRec*>(wbase + kBatch * sizeof(TriRec));
  int* startArr = reinterpret_cast<int*>(wbase + kBatch * (sizeof(TriRec) + sizeof(EdgeRec)));
  int* mySpanStart = startArr + 60;
  int* mySpanInfo = mySpanStart + kQStart;

  // exact test + atomicMin into the z-tile for one (triangle k of the batch, tile pixel q) pair
  auto exact_pair = [&](int k, int q) {
    const float4* rp = reinterpret_cast<const float4*>(&rec[k]);
    const float4 r0 = rp[0], r1 = rp[1], r2 = rp[2], r3 = rp[3];
    const int4 r4 = reinterpret_cast<const int4*>(rp)[4];
    TriSetup ts;
    ts.v0 = mk3(r0.x, r0.y, r0.z); ts.v1 = mk3(r0.w, r1.x, r1.y); ts.v2 = mk3(r1.z, r1.w, r2.x);
    ts.N = mk3(r2.y, r2.z, r2.w); ts.num = r3.x; ts.den = r3.y;
    float a, bq, c;
    // conservative early-z: r4.z is a lower bound of every depth key this triangle can produce
    // (0.9985 * min vertex depth, only for triangles with zmax <= 2 zmin); if even that is behind the
    // pixel's current winner the pair cannot win (ties have equal keys and are never skipped)
    ZEntry cur = zt[q];
    const bool behind = (unsigned)(r4.z ^ 0x80000000) > (unsigned)(cur.key >> 32);
    if (p.spanZ && behind) return;
    const F3 rd = ray_of(q);
    if (hit_exact(ts, ros, rd, a, bq, c, behind)) {
      const int depth = depth_key_exact(a, bq, c, r3.z, r3.w, __int_as_float(r4.x));
      ZEntry nv;
      nv.key = pack_key(depth, r4.y); nv.a = a; nv.b = bq;
      while (nv.key < cur.key) {              // atomicMin on the 64-bit key; (a,b) ride along in the same 128-bit CAS
        const ZEntry old = cas128_shared(&zt[q], cur, nv);
        if (old.key == cur.key && __float_as_uint(old.a) == __float_as_uint(cur.a) && __float_as_uint(old.b) == __float_as_uint(cur.b)) break;
        cur = old;
      }
    }
  };

  const int* smallList = p.bins + (size_t)view * p.F * kMaxSmallTiles + p.tileOffset[tidx];
  const int* bigList = p.bigList + (size_t)view * p.F;
  const float4* vs = p.vscaled + (size_t)b * p.N;
  const float4* pj = p.proj + (size_t)view * p.N;

  // ---- hierarchical z.  The bin arrives partitioned (bin_fill_kernel): first the nNear triangles whose
  // depth-key lower bound lies in the nearer half of the bin's range, then the far ones.  It is rasterised
  // in two passes, near (+ the view's big triangles) then far, and a far triangle whose lower bound is behind
  // EVERY pixel of the (by then fully covered) z-tile is dropped before any of its rows is touched.  Keys only
  // ever decrease, so such a triangle can win no pixel: the result is bit-identical (hiz = 0 and short bins
  // rasterise everything in one pass). ----
  for (int pass = 0; pass < 2; ++pass) {
  unsigned zmaxBits = 0xffffffffu;                    // pass 1: farthest current winner of the tile (0xffffffff if a pixel is still empty)
  if (pass == 1) {
    if (!twoPass) break;
    __syncthreads();                                  // pass 0 complete
    unsigned m = 0u;
    const unsigned* zhi = reinterpret_cast<const unsigned*>(zt) + 1;     // high word of every key
    if (tileX0 + TS <= p.W && tileY0 + TS <= p.H) {                       // tile inside the image: no per-pixel bounds test
      for (int q = qLo + tid; q < qHi; q += NTH) m = max(m, zhi[4 * q]);
    } else {
      for (int q = qLo + tid; q < qHi; q += NTH) {
        const int x = tileX0 + (q % TS), y = tileY0 + (q / TS);
        if (x < p.W && y < p.H) m = max(m, zhi[4 * q]);
      }
    }
    m = __reduce_max_sync(FULL_MASK, m);
    if (lane == 0) atomicMax(&sZmax, m);
    if (tid == 0) {
      nextBatch = 0;
      const int nPass1 = cntSmall - nNear;
      const int G1 = min(kBatch, max(1, (nPass1 + p.batchDiv - 1) / p.batchDiv));
      sPlan[0] = G1; sPlan[1] = (nPass1 + G1 - 1) / G1;      // (every thread read the pass-0 plan before the barrier above)
    }
    __syncthreads();
    zmaxBits = sZmax;
  }
  const int passLo = pass == 0 ? 0 : nNear;                                     // first bin entry of this pass
  const int nSmall = pass == 0 ? (twoPass ? nNear : cntSmall) : cntSmall - nNear;   // bin entries of this pass
  const int nPass = nSmall + (pass == 0 ? cntBig : 0);                           // + the big list in pass 0
  const int G = sPlan[0];          // triangles per batch = min(kBatch, ceil(nPass / batch_div)): a short bin is spread over the warps
  const int nBatches = sPlan[1];   // ceil(nPass / G)
  for (;;) {
    // batch j takes the entries j, j + nBatches, j + 2 nBatches, ... (interleave = 1, the default):
    // every warp gets a sample of the whole bin instead of one spatially coherent chunk, which evens
    // out the batches and shortens the wait at the barriers
    int base = 0;
    if (lane == 0) base = atomicAdd(&nextBatch, p.interleave ? 1 : G);
    base = __shfl_sync(FULL_MASK, base, 0);
    if (base >= (p.interleave ? nBatches : nPass)) break;
    const int i = p.interleave ? base + lane * nBatches : base + lane;
    int n = 0;
    TriRec mine;
    EdgeRec em;
    if (lane < G && i < nPass) {
      const int f = __ldg(i < nSmall ? smallList + passLo + i : bigList + (i - nSmall));
      const int4 fc = __ldg(p.faces4 + f);
      const float4 s0 = __ldg(vs + fc.x), s1 = __ldg(vs + fc.y), s2 = __ldg(vs + fc.z);
      const float4 p0 = __ldg(pj + fc.x), p1 = __ldg(pj + fc.y), p2 = __ldg(pj + fc.z);
      const int4 bb = bbox_exact(p0, p1, p2, p.W, p.H);
      const int cx0 = max(bb.x, tileX0), cx1 = min(bb.z, tileX0 + TS - 1);
      const int cy0 = max(bb.y, tileY0 + rowLo), cy1 = min(bb.w, tileY0 + rowLo + rowN - 1);
      const int w = cx1 - cx0 + 1, h = cy1 - cy0 + 1;
      const int klb = key_lower_bound(p0.z, p1.z, p2.z);
      bool take = w > 0 && h > 0;
      if (pass == 1 && (unsigned)(klb ^ 0x80000000) > zmaxBits) take = false;   // behind the whole tile
      if (take) {
        n = h;                 // work items are ROWS of the clipped bbox
        const TriSetup ts = tri_setup_exact(mk3(s0.x, s0.y, s0.z), mk3(s1.x, s1.y, s1.z), mk3(s2.x, s2.y, s2.z), ros);
        mine.v0x = ts.v0.x; mine.v0y = ts.v0.y; mine.v0z = ts.v0.z;
        mine.v1x = ts.v1.x; mine.v1y = ts.v1.y; mine.v1z = ts.v1.z;
        mine.v2x = ts.v2.x; mine.v2y = ts.v2.y; mine.v2z = ts.v2.z;
        mine.Nx = ts.N.x; mine.Ny = ts.N.y; mine.Nz = ts.N.z;
        mine.num = ts.num; mine.den = ts.den; mine.z0 = p0.z; mine.z1 = p1.z; mine.z2 = p2.z;
        mine.face = f; mine.pad1 = 0;
        mine.pad0 = klb;       // also drives the per-pixel early-z in exact_pair
        edge_setup(p0, p1, p2, (float)tileX0, (float)tileY0, p.cullMargin, em);
        em.geom = (cx0 - tileX0) | ((cy0 - tileY0) << 8) | (w << 16);
        em.pad0 = 0; em.pad1 = 0;
      }
    }
    // warp scan of the row counts: compacted rank + first row of every non-empty triangle
    int rows = n;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(FULL_MASK, rows, o); if (lane >= o) rows += y; }
    const int nrows = __shfl_sync(FULL_MASK, rows, 31);
    const unsigned nzt = __ballot_sync(FULL_MASK, n > 0);
    const int ntri = __popc(nzt);
    if (n > 0) {
      const int rank = __popc(nzt & ((1u << lane) - 1u));
      em.start = rows - n;
      rec[rank] = mine;
      erec[rank] = em;
      startArr[rank] = em.start;
    }
    startArr[ntri + lane] = 0x7fffffff;
    if (lane == 0) startArr[ntri + 32] = 0x7fffffff;
    __syncwarp();
    int K0 = 0;
    int qn = 0, qtot = 0;                                // span queue: spans / pixels waiting for the exact test
    for (int fb = 0; fb < nrows; fb += 32) {
      // (1) one bbox row per lane: which triangle, which row
      const int s = startArr[K0 + 1 + lane];
      const unsigned bits = (s < fb + 32) ? (1u << (s - fb)) : 0u;
      const unsigned mask = __reduce_or_sync(FULL_MASK, bits);
      const int k = K0 + __popc(mask & ((2u << lane) - 1u));
      K0 += __popc(mask);
      const int fr = fb + lane;
      // (2) conservative x-span of the row: pixels outside it provably fail the exact test
      int cnt = 0, info = 0;
      if (fr < nrows) {
        const float4* ep = reinterpret_cast<const float4*>(&erec[k]);
        const float4 e0 = ep[0], e1 = ep[1], e2 = ep[2];
        const int4 e3 = reinterpret_cast<const int4*>(ep)[3];
        const int x0 = e3.y & 0xff, ly = ((e3.y >> 8) & 0xff) + (fr - e3.x), w = e3.y >> 16;
        const float fy = (float)ly;
        const float s0 = fmaf(e0.y, fy, e0.z), s1 = fmaf(e1.x, fy, e1.y), s2 = fmaf(e1.w, fy, e2.x);
        // A_i > 0: x >= t_i ; A_i < 0: x <= t_i ; A_i == 0: row kept iff s_i >= 0
        float xlo = -1.0e30f, xhi = 1.0e30f;
        bool ok = true;
        { const float t = s0 * e2.y; if (e0.x > 0.f) xlo = fmaxf(xlo, t); else if (e0.x < 0.f) xhi = fminf(xhi, t); else ok = ok && (s0 >= 0.f); }
        { const float t = s1 * e2.z; if (e0.w > 0.f) xlo = fmaxf(xlo, t); else if (e0.w < 0.f) xhi = fminf(xhi, t); else ok = ok && (s1 >= 0.f); }
        { const float t = s2 * e2.w; if (e1.z > 0.f) xlo = fmaxf(xlo, t); else if (e1.z < 0.f) xhi = fminf(xhi, t); else ok = ok && (s2 >= 0.f); }
        // 1e-3 px of slack for the rounding of t_i (|t| <= ~64 wherever it matters), then clamp to the bbox row
        int xl = max(x0, (int)ceilf(fmaxf(xlo - 1.0e-3f, -1.0f)));
        int xr = min(x0 + w - 1, (int)floorf(fminf(xhi + 1.0e-3f, 64.0f)));
        if (p.spanZ && (p.spanZ == 1 || pass == 1) && ok && xr >= xl) {
          // span-level early z: trim the span to the pixels whose current winner is NOT provably in
          // front of everything this triangle can produce (depth-key lower bound, see key_lower_bound).
          // Keys only decrease, so a stale read keeps a pixel that could have been dropped, never the
          // other way round; dropped pairs cannot win the depth test => bit-identical result.
          const unsigned kb = (unsigned)(rec[k].pad0 ^ 0x80000000);
          const unsigned* zhi = reinterpret_cast<const unsigned*>(zt + ly * TS) + 1;
          int first = 0x7fffffff, last = -1;
          for (int x = xl; x <= xr; ++x)
            if (!(kb > zhi[4 * x])) { first = min(first, x); last = x; }
          xl = first; xr = last;
        }
        cnt = ok ? max(0, xr - xl + 1) : 0;
        info = (k << 10) | (ly << 5) | xl;
      }
      // (3) append the non-empty spans of these 32 rows to the warp's span queue.  The queue is
      // drained in FULL groups of 32 pixels; what is left (< 32 pixels) waits for the spans of the
      // next 32 rows, so the exact test runs with all lanes busy even when the depth culling above has
      // thinned the spans out.  The queue is flushed completely at the end of the batch (rec[] is reused).
      int incl = cnt;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(FULL_MASK, incl, o); if (lane >= o) incl += y; }
      const int total = __shfl_sync(FULL_MASK, incl, 31);
      const unsigned nz = __ballot_sync(FULL_MASK, cnt > 0);
      if (cnt > 0) {
        const int r = qn + __popc(nz & ((1u << lane) - 1u));
        mySpanStart[r] = qtot + incl - cnt;
        mySpanInfo[r] = info;
      }
      qn += __popc(nz);
      qtot += total;
      mySpanStart[qn + lane] = 0x7fffffff;
      if (lane == 0) mySpanStart[qn + 32] = 0x7fffffff;
      __syncwarp();
      const bool lastRows = fb + 32 >= nrows;
      const int limit = lastRows ? qtot : (qtot & ~31);
      // (4) expand the queued spans 32 pixels at a time
      int S0 = 0;
      for (int t0 = 0; t0 < limit; t0 += 32) {
        const int ss = mySpanStart[S0 + 1 + lane];
        const unsigned sb = (ss < t0 + 32) ? (1u << (ss - t0)) : 0u;
        const unsigned sm = __reduce_or_sync(FULL_MASK, sb);
        const int si = S0 + __popc(sm & ((2u << lane) - 1u));
        S0 += __popc(sm);
        const int t = t0 + lane;
        if (t < limit) {
          const int inf = mySpanInfo[si];
          exact_pair(inf >> 10, ((inf >> 5) & 31) * TS + (inf & 31) + (t - mySpanStart[si]));
        }
      }
      __syncwarp();
      // (5) keep the unconsumed tail: span S0 holds pixel `limit` (partly consumed), the later ones are untouched
      if (limit == qtot) { qn = 0; qtot = 0; }
      else if (limit > 0) {
        if (mySpanStart[S0 + 1] == limit) ++S0;           // span S0 ended exactly at `limit`: nothing of it is left
        const int e0 = S0 + lane, e1 = S0 + lane + 32;
        int st0 = 0, in0 = 0, st1 = 0, in1 = 0;
        if (e0 < qn) { st0 = mySpanStart[e0]; in0 = mySpanInfo[e0]; }
        if (e1 < qn) { st1 = mySpanStart[e1]; in1 = mySpanInfo[e1]; }
        __syncwarp();
        if (e0 < qn) {
          const int used = max(limit - st0, 0);          // > 0 only for the first kept span
          mySpanStart[lane] = max(st0 - limit, 0);
          mySpanInfo[lane] = in0 + used;                  // xl advances by the consumed pixels (stays < 32)
        }
        if (e1 < qn) { mySpanStart[lane + 32] = st1 - limit; mySpanInfo[lane + 32] = in1; }
        qn -= S0; qtot -= limit;
        __syncwarp();
        mySpanStart[qn + lane] = 0x7fffffff;
        if (lane == 0) mySpanStart[qn + 32] = 0x7fffffff;
        __syncwarp();
      }
    }
  }
  }   // pass
  __syncthreads();

  // ---- resolve + shade + write (ref pass 2, :286-403) ----
  const float4* vn = p.vnorm4 + (size_t)b * p.N;
  const float4* vc = p.vcol4 + (size_t)b * p.N;
  const bool doShade = (p.shading == GVV_SHADING_SHADED && p.albedo != GVV_ALBEDO_NORMAL) || p.albedo == GVV_ALBEDO_LIGHTING;
  const bool needNormal = doShade || p.albedo == GVV_ALBEDO_NORMAL;
  if (p.resolvePrefetch) {
    // The resolve loop below walks a thread's pixels one after the other, each with a two-level dependent
    // gather (face -> three vertex normals and colours).  A first sweep pulls those lines into L1 for all of the
    // thread's pixels at once (prefetches need no destination registers), so the loop itself hits L1.
    for (int q = qLo + tid; q < qHi; q += NTH) {
      const unsigned long long key = zt[q].key;
      if (key != kEmptyKey) {
        const int4 fc = __ldg(p.faces4 + (int)(unsigned)(key & 0xffffffffull));
        asm volatile("prefetch.global.L1 [%0];" :: "l"(vn + fc.x));
        asm volatile("prefetch.global.L1 [%0];" :: "l"(vn + fc.y));
        asm volatile("prefetch.global.L1 [%0];" :: "l"(vn + fc.z));
        if (p.albedo == GVV_ALBEDO_VERTEX_COLOR) {
          asm volatile("prefetch.global.L1 [%0];" :: "l"(vc + fc.x));
          asm volatile("prefetch.global.L1 [%0];" :: "l"(vc + fc.y));
          asm volatile("prefetch.global.L1 [%0];" :: "l"(vc + fc.z));
        }
      }
    }
  }
  // Option bulk_out (off: measured slower, raster 0.268 -> 0.289 ms -- every CTA ends waiting for its copies to drain, and
  // the scalar stores it replaces were never the bound).  Output tile through shared memory + bulk async copies (TMA, cp.async.bulk): the 24 B/px of a tile that lies
  // inside a 16-byte aligned image are staged in the (by now dead) per-warp batch buffers and leave as 3 x TS
  // row copies of 128 / 256 / 384 contiguous bytes instead of five scalar stores per pixel, three of them at a
  // 12-byte stride.  Border tiles, odd widths and strips keep the direct stores.
  constexpr bool kStageFits = NPIX * 24 <= (NTH / 32) * kWarpSmemBytes;
  const bool bulkOut = kStageFits && p.bulkOut && stripLog == 0 && ((p.W & 3) == 0) && tileX0 + TS <= p.W && tileY0 + TS <= p.H &&
                       ((reinterpret_cast<uintptr_t>(p.face) | reinterpret_cast<uintptr_t>(p.bary) | reinterpret_cast<uintptr_t>(p.render)) & 15) == 0;
  int* sFace = reinterpret_cast<int*>(smemRaw + ZRAY);
  float2* sBary = reinterpret_cast<float2*>(smemRaw + ZRAY + NPIX * 4);
  float* sRender = reinterpret_cast<float*>(smemRaw + ZRAY + NPIX * 12);
  for (int q = qLo + tid; q < qHi; q += NTH) {
    const int x = tileX0 + (q % TS), y = tileY0 + (q / TS);
    if (x >= p.W || y >= p.H) continue;
    const size_t pix = pixBase + (size_t)y * p.W + x;
    const ZEntry ze = zt[q];
    const unsigned long long key = ze.key;
    int faceId = -1;
    float a = 0.f, bq = 0.f, cr = 0.f, cg = 1.f, cb = 0.f;
    if (key != kEmptyKey) {
      faceId = (int)(unsigned)(key & 0xffffffffull);
      const int4 fc = __ldg(p.faces4 + faceId);
      a = ze.a; bq = ze.b;                               // the winner's barycentrics, stored with its key
      const float c = __fsub_rn(__fsub_rn(1.f, a), bq);  // c = 1 - a - b (RendererUtil.h:125)
      // the pixel normal only reaches the output through the SH shading or the `normal` albedo: shadeless
      // vertexColor / textured / foregroundMask renders skip the three gathers, the ray, the sqrt and the divides
      F3 nr = mk3(0.f, 0.f, 1.f);
      if (needNormal) {
        const F3 rd = ray_of(q);
        const float4 n0 = __ldg(vn + fc.x), n1 = __ldg(vn + fc.y), n2 = __ldg(vn + fc.z);
        nr = mk3(interp3(a, bq, c, n0.x, n1.x, n2.x), interp3(a, bq, c, n0.y, n1.y, n2.y), interp3(a, bq, c, n0.z, n1.z, n2.z));
        const float len = __fsqrt_rn(dot3x(nr, nr));
        nr = mk3(__fdiv_rn(nr.x, len), __fdiv_rn(nr.y, len), __fdiv_rn(nr.z, len));
        if (dot3x(nr, rd) > 0.f) nr = mk3(-nr.x, -nr.y, -nr.z);
      }
      if (p.albedo == GVV_ALBEDO_TEXTURED) {
        const float* tc = p.texcoords + (size_t)faceId * 6;
        const float u = interp3(a, bq, c, __ldg(tc + 0), __ldg(tc + 2), __ldg(tc + 4));
        const float v = interp3(a, bq, c, __fsub_rn(1.f, __ldg(tc + 1)), __fsub_rn(1.f, __ldg(tc + 3)), __fsub_rn(1.f, __ldg(tc + 5)));
        const float fu = fminf(fmaxf(__fmul_rn(u, (float)p.texW), 0.f), (float)(p.texW - 1));
        const float fv = fminf(fmaxf(__fmul_rn(v, (float)p.texH), 0.f), (float)(p.texH - 1));
        // nearest texel (LU,LV); the bilinear mix is commented out in the reference (:372-373)
        const int iu = __float2int_rz(__fadd_rn((float)__float2int_rz(__fadd_rn(fu, -0.5f)), 0.5f));
        const int iv = __float2int_rz(__fadd_rn((float)__float2int_rz(__fadd_rn(fv, -0.5f)), 0.5f));
        const float* tb = p.texture + (size_t)b * p.texH * p.texW * 3;
        const float* tx = tb + ((size_t)iv * p.texW + iu) * 3;
        cr = __ldg(tx); cg = __ldg(tx + 1); cb = __ldg(tx + 2);
        if (p.texBilinear) {
          // non-default variant: the bilinear mix the reference has commented out (:365-372), weights as written there
          const float LU = __fadd_rn((float)__float2int_rz(__fadd_rn(fu, -0.5f)), 0.5f), HU = __fadd_rn((float)__float2int_rz(__fadd_rn(fu, -0.5f)), 1.5f);
          const float LV = __fadd_rn((float)__float2int_rz(__fadd_rn(fv, -0.5f)), 0.5f), HV = __fadd_rn((float)__float2int_rz(__fadd_rn(fv, -0.5f)), 1.5f);
          const int hu = min(__float2int_rz(HU), p.texW - 1), hv = min(__float2int_rz(HV), p.texH - 1);
          const float wLULV = (fv - LV) * (fu - LU), wLUHV = (HV - fv) * (fu - LU), wHULV = (fv - LV) * (HU - fu), wHUHV = (HV - fv) * (HU - fu);
          const float* tLUHV = tb + ((size_t)hv * p.texW + iu) * 3;
          const float* tHULV = tb + ((size_t)iv * p.texW + hu) * 3;
          const float* tHUHV = tb + ((size_t)hv * p.texW + hu) * 3;
          cr = wLULV * cr + wHULV * __ldg(tHULV) + wLUHV * __ldg(tLUHV) + wHUHV * __ldg(tHUHV);
          cg = wLULV * cg + wHULV * __ldg(tHULV + 1) + wLUHV * __ldg(tLUHV + 1) + wHUHV * __ldg(tHUHV + 1);
          cb = wLULV * cb + wHULV * __ldg(tHULV + 2) + wLUHV * __ldg(tLUHV + 2) + wHUHV * __ldg(tHUHV + 2);
        }
      } else if (p.albedo == GVV_ALBEDO_VERTEX_COLOR) {
        const float4 c0 = __ldg(vc + fc.x), c1 = __ldg(vc + fc.y), c2 = __ldg(vc + fc.z);
        cr = interp3(a, bq, c, c0.x, c1.x, c2.x);
        cg = interp3(a, bq, c, c0.y, c1.y, c2.y);
        cb = interp3(a, bq, c, c0.z, c1.z, c2.z);
      } else if (p.albedo == GVV_ALBEDO_NORMAL) {
        cr = __fmul_rn(__fadd_rn(nr.x, 1.f), 0.5f);
        cg = __fmul_rn(__fadd_rn(nr.y, 1.f), 0.5f);
        cb = __fmul_rn(__fadd_rn(nr.z, 1.f), 0.5f);
      } else {
        cr = cg = cb = 1.f;   // lighting / foregroundMask
      }
      if (doShade) {
        cr = __fmul_rn(cr, sh_eval(shc, nr));
        cg = __fmul_rn(cg, sh_eval(shc + 9, nr));
        cb = __fmul_rn(cb, sh_eval(shc + 18, nr));
      }
    }
    if (bulkOut) {
      sFace[q] = faceId;
      sBary[q] = make_float2(a, bq);
      sRender[3 * q + 0] = cr; sRender[3 * q + 1] = cg; sRender[3 * q + 2] = cb;
    } else {
      // streaming (evict-first) stores: 24 B/px of outputs that nothing re-reads soon must not displace the mesh
      // arrays and bins from L2 / the stores' own lines from each other (raster 0.268 -> 0.263 ms with the background path)
      __stcs(p.face + pix, faceId);
      __stcs(reinterpret_cast<float2*>(p.bary) + pix, make_float2(a, bq));
      __stcs(p.render + 3 * pix + 0, cr); __stcs(p.render + 3 * pix + 1, cg); __stcs(p.render + 3 * pix + 2, cb);
    }
  }
  if (bulkOut) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the generic-proxy writes above become visible to the copy engine
    __syncthreads();
    if (tid < TS) {                                                // one row of the tile per thread: three bulk copies
      const size_t rowPix = pixBase + (size_t)(tileY0 + tid) * p.W + tileX0;
      const unsigned sF = (unsigned)__cvta_generic_to_shared(sFace + tid * TS);
      const unsigned sB = (unsigned)__cvta_generic_to_shared(sBary + tid * TS);
      const unsigned sR = (unsigned)__cvta_generic_to_shared(sRender + tid * TS * 3);
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" :: "l"(p.face + rowPix), "r"(sF), "n"(TS * 4) : "memory");
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" :: "l"(p.bary + 2 * rowPix), "r"(sB), "n"(TS * 8) : "memory");
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" :: "l"(p.render + 3 * rowPix), "r"(sR), "n"(TS * 12) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // shared memory must stay allocated until it has been read
    }
  }
  // self-cleaning scratch: the last strip of the tile to finish resets the tile's counters (every strip
  // read them before it got here)
  if (tid == 0) {
    if (stripLog == 0 || atomicAdd(p.tileDone + tidx, 1) == (1 << stripLog) - 1) { p.tileCount[tidx] = 0; p.tileCursor[tidx] = 0; p.tileCursorFar[tidx] = 0; p.tileDone[tidx] = 0; }
  }
  if (p.ctaTrace) {
    __syncthreads();
    if (tid == 0) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); p.ctaTrace[4 * ((size_t)rank * p.V + view) + 1] = t; }
  }
}

// ------------------------------------------------------------------------------------------------
// debug: one exact (pixel, triangle) evaluation per thread (same device functions as the raster)
// ------------------------------------------------------------------------------------------------
__global__ void debug_eval_kernel(const CamRec* __restrict__ cams, const float4* __restrict__ vscaled, const float4* __restrict__ proj,
                                  const int4* __restrict__ faces4, int N, int C, int n, const int* __restrict__ q,
                                  int* __restrict__ key, float* __restrict__ ab) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int view = q[4 * i], x = q[4 * i + 1], y = q[4 * i + 2], f = q[4 * i + 3];
  const CamRec* cam = cams + view;
  const int b = view / C;
  const int4 fc = faces4[f];
  const float4 s0 = vscaled[(size_t)b * N + fc.x], s1 = vscaled[(size_t)b * N + fc.y], s2 = vscaled[(size_t)b * N + fc.z];
  const float4 p0 = proj[(size_t)view * N + fc.x], p1 = proj[(size_t)view * N + fc.y], p2 = proj[(size_t)view * N + fc.z];
  const F3 ros = mk3(cam->ros[0], cam->ros[1], cam->ros[2]);
  const TriSetup ts = tri_setup_exact(mk3(s0.x, s0.y, s0.z), mk3(s1.x, s1.y, s1.z), mk3(s2.x, s2.y, s2.z), ros);
  const F3 rd = ray_dir_exact(cam->Pinv, cam->ro, __fadd_rn((float)x, 0.5f), __fadd_rn((float)y, 0.5f));
  float a = 0.f, bq = 0.f, c = 0.f;
  const bool hit = hit_exact(ts, ros, rd, a, bq, c);
  key[i] = hit ? depth_key_exact(a, bq, c, p0.z, p1.z, p2.z) : (int)0x80000000;
  ab[2 * i] = hit ? a : -1.f; ab[2 * i + 1] = hit ? bq : -1.f;
}

int launch_debug_eval(const Scratch& s, const int4* faces4, int N, int C, int W, int H, int n, const int* dq, int* dkey, float* dab, cudaStream_t st) {
  (void)W; (void)H;
  debug_eval_kernel<<<(n + 127) / 128, 128, 0, st>>>(s.cams, s.vscaled, s.proj, faces4, N, C, n, dq, dkey, dab);
  return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

// ------------------------------------------------------------------------------------------------
// launcher
// ------------------------------------------------------------------------------------------------
static inline bool launch_ok() { return cudaGetLastError() == cudaSuccess; }

int launch_camera(const float* extr, const float* intr, CamRec* cams, int* bigCount, int V, cudaStream_t st) {
  camera_kernel<<<(V + 63) / 64, 64, 0, st>>>(extr, intr, cams, bigCount, V);
  return 1;
}

int launch_vertex(const FwdArgs& a, cudaStream_t st) {
  if (a.F > 0) launch_chained(a.chain, face_normal_kernel, dim3((a.F + 255) / 256, a.B), dim3(256), 0, st, a.vertex_pos, a.faces4, a.s.fnorm4, a.N, a.F);
  launch_chained(a.chain, vertex_kernel, dim3((a.N + 127) / 128, a.B * a.C), dim3(128), 0, st, a.vertex_pos, a.vertex_color, a.s.fnorm4, a.vfOffsets, a.vfList,
                 a.extrinsics, a.intrinsics, a.s.bigCount, a.s.proj, a.s.vscaled, a.s.vnorm4, a.s.vcol4, a.vertex_normal, a.N, a.F, a.C);
  return a.F > 0 ? 2 : 1;
}

int launch_forward(const FwdArgs& a, cudaStream_t st, KernelTimer* tm) {
  const int V = a.B * a.C;
  int launches = 0;
  tm->begin(K_VERTEX, st);
  launches += launch_vertex(a, st);   // face normals + per-vertex work
  tm->end(st);
  const int tileShift = a.tile == 32 ? 5 : 4;
  const dim3 gridF((a.F + 256 * kBinFacesPerThread - 1) / (256 * kBinFacesPerThread), V);
  tm->begin(K_BIN_COUNT, st);
  if (a.nT <= kBinWindow)
    launch_chained(a.chain, bin_count_kernel<false>, gridF, dim3(256), 3 * a.nT * sizeof(int), st, a.faces4, a.s.proj, a.s.tileCount, a.s.tileMinK, a.s.tileMaxK,
                   a.s.bigCount, a.s.bigList, a.F, a.N, a.W, a.H, tileShift, a.tilesX, a.tilesY, a.nT);
  else
    launch_chained(a.chain, bin_count_kernel<true>, gridF, dim3(256), 3 * kBinWindow * sizeof(int), st, a.faces4, a.s.proj, a.s.tileCount, a.s.tileMinK, a.s.tileMaxK,
                   a.s.bigCount, a.s.bigList, a.F, a.N, a.W, a.H, tileShift, a.tilesX, a.tilesY, a.nT);
  tm->end(st);
  ++launches;
  tm->begin(K_BIN_SCAN, st);
  const int nItems = a.splitUnit > 0 ? a.nT + a.nT / 2 : a.nT;   // work-list slots per view (strips need spare ones); Scratch::tileOrder holds nT + nT/2
  // Heavy bins (>= heavyThr triangles among the kHeavySlots heaviest items of a view) get a 1024-thread CTA, i.e.
  // a whole SM, from a second launch: a 256-thread CTA shares its SM with three others and would make the
  // tile the critical path of the launch (1M triangles at 3840x2160, one view: 1.53 ms -> 0.59 ms).  Which bins
  // qualify is decided on the GPU (bin_scan_kernel).  A CTA that needs a whole SM must find one empty and its
  // per-triangle cost is ~1.5x that of a 256-thread CTA (32 warps on one z-tile), so with many views the launch
  // only costs (8 views, 70k triangles: 0.298 -> 0.307 ms).  The heavy launch therefore holds at most SMs/4 CTAs,
  // split evenly over the views, and is dropped when that leaves fewer than 16 per view (more than two views on
  // a B200): a single tile can only be the critical path of a launch that has very few views -- the work of the
  // other tiles grows with the number of views, the longest tile does not.  heavy_mode 2 skips this gate.
  const int smCount = max(1, a.ctaSlots / 4);
  const int kHeavySlots = min(min(nItems, a.heavySlots), max(1, (smCount / 4) / max(V, 1)));
  const bool fewViews = (smCount / 4) / max(V, 1) >= 16 || a.heavyMode == 2;
  const bool useHeavy = a.heavyMode > 0 && a.heavyThr > 0 && a.tile == 32 && !a.rayCache && a.ctaThreads == 256 && fewViews;
  const int maxLog = a.splitUnit > 0 ? (a.tile == 32 ? 3 : 2) : 0;
  static unsigned long long scanAttr = 0;
  if (first_use_on_device(&scanAttr)) cudaFuncSetAttribute(bin_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
  if (a.nT > kMaxTiles) return -1;   // rejected by gvv_create / gvv_set_option("tile") already
  launch_chained(a.chain, bin_scan_kernel, dim3(V + (V + 1023) / 1024), dim3(1024), a.nT * sizeof(int), st, a.s.tileCount, a.s.tileOffset, a.s.tileOrder, a.s.tileMinK, a.s.tileMaxK, a.s.tileThr,
                 a.nT, a.tilesX, nItems, a.splitUnit > 0 ? a.splitUnit : 1, maxLog,
                 useHeavy ? a.heavyThr : 0, kHeavySlots, a.heavyMode == 2 ? -1 : max(1, a.ctaSlots / V), a.spreadEmpty, a.extrinsics, a.intrinsics, a.s.cams, V);
  tm->end(st);
  ++launches;
  tm->begin(K_BIN_FILL, st);
  if (a.nT <= kBinWindow)
    launch_chained(a.chain, bin_fill_kernel<false>, gridF, dim3(256), 4 * a.nT * sizeof(int), st, a.faces4, a.s.proj, a.s.tileOffset, a.s.tileCount, a.s.tileThr,
                   a.s.tileCursor, a.s.tileCursorFar, a.s.bins, a.F, a.N, a.W, a.H, tileShift, a.tilesX, a.tilesY, a.nT);
  else
    launch_chained(a.chain, bin_fill_kernel<true>, gridF, dim3(256), 4 * kBinWindow * sizeof(int), st, a.faces4, a.s.proj, a.s.tileOffset, a.s.tileCount, a.s.tileThr,
                   a.s.tileCursor, a.s.tileCursorFar, a.s.bins, a.F, a.N, a.W, a.H, tileShift, a.tilesX, a.tilesY, a.nT);
  tm->end(st);
  ++launches;
  RasterParams p;
  p.faces4 = a.faces4; p.proj = a.s.proj; p.vscaled = a.s.vscaled; p.vnorm4 = a.s.vnorm4; p.vcol4 = a.s.vcol4;
  p.cams = a.s.cams; p.tileCount = a.s.tileCount; p.tileCursor = a.s.tileCursor; p.tileCursorFar = a.s.tileCursorFar; p.tileDone = a.s.tileDone; p.nItems = nItems; p.tileOffset = a.s.tileOffset; p.tileOrder = a.s.tileOrder; p.V = V;
  p.bigCount = a.s.bigCount; p.bigList = a.s.bigList; p.bins = a.s.bins;
  p.texture = a.texture; p.texcoords = a.texcoords; p.sh_coeff = a.sh_coeff;
  p.bary = a.bary; p.face = a.face; p.render = a.render; p.ctaTrace = a.s.ctaTrace;
  p.C = a.C; p.N = a.N; p.F = a.F; p.W = a.W; p.H = a.H; p.texH = a.texH; p.texW = a.texW;
  p.albedo = a.albedo; p.shading = a.shading; p.tilesX = a.tilesX; p.nT = a.nT; p.cullMargin = a.cullMargin; p.invC = 1.f / (float)a.C; p.batchDiv = a.batchDiv; p.interleave = a.interleave; p.hiz = a.hiz; p.hizMin = a.hizMin; p.spanZ = a.spanZ; p.texBilinear = a.texBilinear; p.resolvePrefetch = a.resolvePrefetch; p.bulkOut = a.bulkOut;
  p.grid2d = nItems <= 65535 ? 1 : 0;
  const dim3 gridT = p.grid2d ? dim3((unsigned)V, (unsigned)nItems) : dim3((unsigned)nItems * (unsigned)V);
  tm->begin(K_RASTER, st);
  static unsigned long long attrSet = 0;
  if (first_use_on_device(&attrSet)) {   // > 48 KB of dynamic shared memory needs the opt-in (once per device and device function)
#define GVV_RASTER_ATTR(TS, RC, NTH) cudaFuncSetAttribute(raster_kernel<TS, RC, NTH>, cudaFuncAttributeMaxDynamicSharedMemorySize, raster_smem_bytes<TS, RC, NTH>())
    GVV_RASTER_ATTR(16, true, 256); GVV_RASTER_ATTR(32, true, 256); GVV_RASTER_ATTR(16, false, 256); GVV_RASTER_ATTR(32, false, 256);
    GVV_RASTER_ATTR(16, false, 128); GVV_RASTER_ATTR(32, false, 128); GVV_RASTER_ATTR(32, false, 1024);
#undef GVV_RASTER_ATTR
  }
  // The heavy launch goes FIRST, so that its one-SM CTAs are placed while the SMs are empty (behind the small CTAs
  // they would starve until the very end -- measured).  The 256-thread launch follows on the SAME stream as a
  // programmatic dependent launch: the heavy CTAs signal griddepcontrol.launch_dependents as their first
  // instruction, so the small CTAs start as soon as every heavy CTA is resident (or gone) and fill the other SMs.
  // The two launches write disjoint tiles and both only read what bin_fill_kernel finished before the heavy
  // launch began, so the dependent launch needs no griddepcontrol.wait at its top.  (A side stream with fork / join
  // events did the same; one stream keeps the call capturable as a linear chain and costs nothing when no bin is heavy.)
  if (useHeavy) {
    RasterParams ph = p;
    ph.role = 1; ph.pdl = 0;
    launch_chained(a.chain, raster_kernel<32, false, 1024>, (p.grid2d ? dim3((unsigned)V, (unsigned)kHeavySlots) : dim3((unsigned)kHeavySlots * (unsigned)V)), dim3(1024), raster_smem_bytes<32, false, 1024>(), st, ph);
    ++launches;
  }
  p.role = 0; p.pdl = useHeavy ? 1 : 0;
  {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = gridT; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = (useHeavy || a.chain) ? 1 : 0;
#define GVV_RASTER_LAUNCH(TS, RC, NTH) do { cfg.blockDim = dim3(NTH); cfg.dynamicSmemBytes = raster_smem_bytes<TS, RC, NTH>(); \
                                            cudaLaunchKernelEx(&cfg, raster_kernel<TS, RC, NTH>, p); } while (0)
    if (a.tile == 16) {
      if (a.rayCache) GVV_RASTER_LAUNCH(16, true, 256);
      else if (a.ctaThreads == 128) GVV_RASTER_LAUNCH(16, false, 128);
      else GVV_RASTER_LAUNCH(16, false, 256);
    } else {
      if (a.rayCache) GVV_RASTER_LAUNCH(32, true, 256);
      else if (a.ctaThreads == 128) GVV_RASTER_LAUNCH(32, false, 128);
      else GVV_RASTER_LAUNCH(32, false, 256);
    }
#undef GVV_RASTER_LAUNCH
  }
  tm->end(st);
  ++launches;
  return launch_ok() ? launches : -1;
}

}  // namespace gvv
