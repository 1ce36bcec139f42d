// gvv_api.cu -- C-ABI shim (include/gvv_b200.h): handle life cycle, argument validation, scratch.
//
// Mirrors the reference's op boundary (CudaRenderer.cpp / CudaRendererGrad.cpp) minus TensorFlow:
// the per-batch host loop of Compute() (CudaRenderer.cpp:309-328) is gone -- one call enqueues
// the kernels for all B*C views on the caller's stream and returns without synchronising.
#include <cstdio>
#include <cstring>
#include <cstdarg>
#include <cstdlib>
#include <string>
#include <vector>
#include <new>
#include "gvv_internal.h"

using namespace gvv;

static thread_local char g_err[512] = "";

static int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

#define CK(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess) return fail(GVV_ECUDA, "%s failed: %s", #call, cudaGetErrorString(e_)); \
  } while (0)

extern "C" const char* gvv_last_error(void) { return g_err; }

template <typename T>
static cudaError_t dmalloc(T** p, size_t count) {
  *p = nullptr;
  if (count == 0) count = 1;
  return cudaMalloc((void**)p, count * sizeof(T));
}

static void free_scratch(Scratch& s) {
  cudaFree(s.cams); cudaFree(s.proj); cudaFree(s.vscaled); cudaFree(s.fnorm4); cudaFree(s.vnorm4); cudaFree(s.vcol4);
  cudaFree(s.tileCount); cudaFree(s.tileCursor); cudaFree(s.tileCursorFar); cudaFree(s.tileMinK); cudaFree(s.tileMaxK); cudaFree(s.tileThr); cudaFree(s.tileOffset); cudaFree(s.tileOrder); cudaFree(s.tileDone); cudaFree(s.bigCount);
  cudaFree(s.bigList); cudaFree(s.bins); cudaFree(s.gnorm); cudaFree(s.bpos4); cudaFree(s.bcol4); cudaFree(s.bnor4); cudaFree(s.tileCounter); cudaFree(s.ctaTrace);
  s = Scratch();
}

static void set_tile(gvv_renderer* h, int tile) {
  h->tile = tile;
  h->tilesX = (h->W + tile - 1) / tile;
  h->tilesY = (h->H + tile - 1) / tile;
  h->nT = h->tilesX * h->tilesY;
}

// Scratch is sized for the largest batch seen (or reserved with gvv_reserve); growing it is the only time a call
// allocates.  Growth frees and re-allocates every scratch buffer, so it must not happen while anything can still
// use the old ones:
//   - the whole device is synchronised first (work of this handle on ANY stream, not only the caller's);
//   - it is refused while `st` is being captured (allocation is illegal there) and once the handle has been used
//     under stream capture -- the captured graphs have the old scratch pointers baked in, and replaying them
//     after a re-allocation would read and write freed memory.  Reserve the largest batch before capturing.
static int ensure_scratch(gvv_renderer* h, int B, cudaStream_t st) {
  Scratch& s = h->s;
  const int V = B * h->C;
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(st, &cap) != cudaSuccess) { cudaGetLastError(); cap = cudaStreamCaptureStatusNone; }
  const bool capturing = cap != cudaStreamCaptureStatusNone;
  if (V <= s.capViews && B <= s.capBatch) { if (capturing) h->captured = 1; return GVV_OK; }
  if (capturing)
    return fail(GVV_EINVAL, "scratch for batch %d is not allocated and the stream is being captured: call gvv_reserve (or run one eager call) first", B);
  if (h->captured)
    return fail(GVV_EINVAL, "batch %d exceeds the scratch (batch %d) that CUDA graphs captured on this handle have baked in: "
                            "gvv_reserve the largest batch before capturing, or use another handle", B, s.capBatch);
  CK(cudaDeviceSynchronize());
  free_scratch(s);
  const size_t N = h->N, F = h->F, nT = h->nT;
  cudaError_t e = cudaSuccess;
  auto acc = [&](cudaError_t r) { if (e == cudaSuccess) e = r; };
  acc(dmalloc(&s.cams, (size_t)V));
  acc(dmalloc(&s.proj, (size_t)V * N));
  acc(dmalloc(&s.vscaled, (size_t)B * N));
  acc(dmalloc(&s.fnorm4, (size_t)B * F));
  acc(dmalloc(&s.vnorm4, (size_t)B * N));
  acc(dmalloc(&s.vcol4, (size_t)B * N));
  acc(dmalloc(&s.tileCount, (size_t)V * nT));
  acc(dmalloc(&s.tileCursor, (size_t)V * nT));
  acc(dmalloc(&s.tileCursorFar, (size_t)V * nT));
  acc(dmalloc(&s.tileMinK, (size_t)V * nT));
  acc(dmalloc(&s.tileMaxK, (size_t)V * nT));
  acc(dmalloc(&s.tileThr, (size_t)V * nT));
  acc(dmalloc(&s.tileOffset, (size_t)V * nT));
  acc(dmalloc(&s.tileOrder, (size_t)V * (nT + nT / 2)));
  acc(dmalloc(&s.tileDone, (size_t)V * nT));
  acc(dmalloc(&s.bigCount, (size_t)V));
  acc(dmalloc(&s.bigList, (size_t)V * F));
  acc(dmalloc(&s.bins, (size_t)V * F * kMaxSmallTiles));
  acc(dmalloc(&s.gnorm, (size_t)B * N * 4));
  acc(dmalloc(&s.bpos4, (size_t)B * N));
  acc(dmalloc(&s.bcol4, (size_t)B * N));
  acc(dmalloc(&s.bnor4, (size_t)V * N));
  acc(dmalloc(&s.tileCounter, (size_t)4));
  if (h->ctaTrace) acc(dmalloc(&s.ctaTrace, (size_t)V * (nT + nT / 2) * 4));
  if (e != cudaSuccess) {
    free_scratch(s);
    return fail(GVV_ENOMEM, "scratch allocation for %d views failed: %s", V, cudaGetErrorString(e));
  }
  CK(cudaMemsetAsync(s.tileCount, 0, (size_t)V * nT * sizeof(int), st));
  CK(cudaMemsetAsync(s.tileCursor, 0, (size_t)V * nT * sizeof(int), st));
  CK(cudaMemsetAsync(s.tileCursorFar, 0, (size_t)V * nT * sizeof(int), st));
  CK(cudaMemsetAsync(s.tileMinK, 0x7f, (size_t)V * nT * sizeof(int), st));   // any value >= every key: bin_scan_kernel resets to INT_MAX
  CK(cudaMemsetAsync(s.tileMaxK, 0x80, (size_t)V * nT * sizeof(int), st));   // 0x80808080 < every key
  CK(cudaMemsetAsync(s.tileDone, 0, (size_t)V * nT * sizeof(int), st));
  CK(cudaMemsetAsync(s.bigCount, 0, (size_t)V * sizeof(int), st));
  s.capViews = V;
  s.capBatch = B;
  return GVV_OK;
}

extern "C" int gvv_create(const gvv_desc* d, gvv_handle* out) {
  if (!d || !out) return fail(GVV_EINVAL, "gvv_create: null argument");
  *out = nullptr;
  // attribute checks of CudaRenderer.cpp:47-76 (errors instead of prints / OP_REQUIRES)
  if (d->num_vertices <= 0) return fail(GVV_EINVAL, "number_of_vertices not set!");
  if (d->num_cameras <= 0) return fail(GVV_EINVAL, "number_of_cameras not set!");
  if (d->width <= 0) return fail(GVV_EINVAL, "render_resolution_u not set!");
  if (d->height <= 0) return fail(GVV_EINVAL, "render_resolution_v not set!");
  if (d->width > 65535 || d->height > 65535) return fail(GVV_EINVAL, "render resolution above 65535 is not supported");
  // the per-view tile histogram of bin_scan_kernel lives in shared memory: at most kMaxTiles tiles of 32 x 32 pixels
  if ((long long)((d->width + 31) / 32) * ((d->height + 31) / 32) > kMaxTiles)
    return fail(GVV_EINVAL, "render resolution %d x %d needs more than %d tiles of 32 x 32 pixels: not supported", d->width, d->height, kMaxTiles);
  if (d->albedo_mode < 0 || d->albedo_mode > 4) return fail(GVV_EINVAL, "INVALID ALBEDO MODE");
  if (d->shading_mode < 0 || d->shading_mode > 1) return fail(GVV_EINVAL, "INVALID SHADING MODE");
  if (d->num_faces < 0 || (d->num_faces > 0 && !d->faces)) return fail(GVV_EINVAL, "faces missing");
  if (d->albedo_mode == GVV_ALBEDO_TEXTURED && d->num_faces > 0 && !d->texcoords)
    return fail(GVV_EINVAL, "textured albedo needs texture_coordinates");
  for (int i = 0; i < d->num_faces * 3; ++i)
    if (d->faces[i] < 0 || d->faces[i] >= d->num_vertices)
      return fail(GVV_EINVAL, "face %d references vertex %d outside [0,%d)", i / 3, d->faces[i], d->num_vertices);
  CK(cudaSetDevice(d->device));

  gvv_renderer* h = new (std::nothrow) gvv_renderer();
  if (!h) return fail(GVV_ENOMEM, "out of host memory");
  h->device = d->device;
  h->F = d->num_faces; h->N = d->num_vertices; h->C = d->num_cameras; h->W = d->width; h->H = d->height;
  h->albedo = d->albedo_mode;
  h->shading = d->shading_mode;
  if (h->albedo == GVV_ALBEDO_FOREGROUND_MASK) h->shading = GVV_SHADING_SHADELESS;   // CudaRenderer.cpp:72-76
  h->imgFilter = d->image_filter_size;
  h->texFilter = d->texture_filter_size;
  h->computeNormalMap = d->compute_normal_map ? 1 : 0;
  set_tile(h, 32);
  { int sms = 0; if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, d->device) == cudaSuccess && sms > 0) h->ctaSlots = 4 * sms; }

  // topology: faces padded to int4; vertex -> incident faces CSR by counting sort, O(N+F)
  // (reference: O(N*F) double loop, CUDABasedRasterization.cpp:125-154; same ascending face order;
  //  a face that repeats a vertex is listed once for it, as there).
  const int F = h->F, N = h->N;
  std::vector<int4> f4((size_t)F);
  std::vector<int> offs((size_t)N + 1, 0);
  auto distinct = [&](int f, int k) {
    const int* v = d->faces + 3 * f;
    for (int j = 0; j < k; ++j) if (v[j] == v[k]) return false;
    return true;
  };
  for (int f = 0; f < F; ++f) {
    f4[f] = make_int4(d->faces[3 * f], d->faces[3 * f + 1], d->faces[3 * f + 2], 0);
    for (int k = 0; k < 3; ++k) if (distinct(f, k)) offs[d->faces[3 * f + k] + 1]++;
  }
  for (int n = 0; n < N; ++n) offs[n + 1] += offs[n];
  std::vector<int> list((size_t)offs[N]);
  std::vector<int> cur(offs.begin(), offs.end() - 1);
  for (int f = 0; f < F; ++f)
    for (int k = 0; k < 3; ++k) if (distinct(f, k)) list[cur[d->faces[3 * f + k]]++] = f;

  cudaError_t e = cudaSuccess;
  auto acc = [&](cudaError_t r) { if (e == cudaSuccess) e = r; };
  acc(dmalloc(&h->faces4, (size_t)F));
  acc(dmalloc(&h->vfOffsets, (size_t)N + 1));
  acc(dmalloc(&h->vfList, list.size()));
  if (e == cudaSuccess && F) acc(cudaMemcpy(h->faces4, f4.data(), sizeof(int4) * F, cudaMemcpyHostToDevice));
  if (e == cudaSuccess) acc(cudaMemcpy(h->vfOffsets, offs.data(), sizeof(int) * (N + 1), cudaMemcpyHostToDevice));
  if (e == cudaSuccess && !list.empty()) acc(cudaMemcpy(h->vfList, list.data(), sizeof(int) * list.size(), cudaMemcpyHostToDevice));
  if (d->texcoords && F) {
    h->hasTexcoords = true;
    acc(dmalloc(&h->texcoords, (size_t)F * 6));
    if (e == cudaSuccess) acc(cudaMemcpy(h->texcoords, d->texcoords, sizeof(float) * 6 * F, cudaMemcpyHostToDevice));
  }
  if (e != cudaSuccess) {
    gvv_destroy(h);
    return fail(GVV_ECUDA, "topology upload failed: %s", cudaGetErrorString(e));
  }
  // GVV_OPTIONS="key=value,key=value": gvv_set_option knobs applied to every new handle (tuning / A-B runs
  // through wrappers that create their handles internally); unknown keys are an error
  if (const char* env = getenv("GVV_OPTIONS")) {
    std::string spec(env);
    size_t pos = 0;
    while (pos < spec.size()) {
      size_t end = spec.find(',', pos);
      if (end == std::string::npos) end = spec.size();
      const std::string kv = spec.substr(pos, end - pos);
      const size_t eq = kv.find('=');
      if (eq != std::string::npos && eq > 0) {
        const int rc = gvv_set_option(h, kv.substr(0, eq).c_str(), atoi(kv.c_str() + eq + 1));
        if (rc != GVV_OK) { gvv_destroy(h); return fail(rc, "GVV_OPTIONS: bad entry '%s'", kv.c_str()); }
      }
      pos = end + 1;
    }
  }
  *out = h;
  return GVV_OK;
}

extern "C" int gvv_reserve(gvv_handle h, int32_t max_batch, void* stream) {
  if (!h) return fail(GVV_EINVAL, "gvv_reserve: null handle");
  if (max_batch <= 0) return fail(GVV_EINVAL, "gvv_reserve: batch must be positive");
  CK(cudaSetDevice(h->device));
  return ensure_scratch(h, max_batch, (cudaStream_t)stream);
}

extern "C" int gvv_destroy(gvv_handle h) {
  if (!h) return GVV_OK;
  cudaSetDevice(h->device);
  cudaDeviceSynchronize();
  free_scratch(h->s);
  if (h->timer.ev) { for (int i = 0; i < 2 * KernelTimer::kCap; ++i) cudaEventDestroy(h->timer.ev[i]); delete[] h->timer.ev; delete[] h->timer.slot; }
  cudaFree(h->faces4); cudaFree(h->texcoords); cudaFree(h->vfOffsets); cudaFree(h->vfList); cudaFree(h->texelTable);
  delete h;
  return GVV_OK;
}

extern "C" int gvv_set_option(gvv_handle h, const char* key, int32_t value) {
  if (!h || !key) return fail(GVV_EINVAL, "gvv_set_option: null argument");
  if (!strcmp(key, "tile")) {
    if (value != 16 && value != 32) return fail(GVV_EINVAL, "tile must be 16 or 32");
    if ((long long)((h->W + value - 1) / value) * ((h->H + value - 1) / value) > kMaxTiles)
      return fail(GVV_EINVAL, "tile %d gives more than %d tiles at %d x %d: not supported", value, kMaxTiles, h->W, h->H);
    if (value != h->tile && h->captured) return fail(GVV_EINVAL, "tile cannot change after the handle was used under stream capture");
    if (value != h->tile) {
      cudaSetDevice(h->device);
      cudaDeviceSynchronize();
      free_scratch(h->s);
      set_tile(h, value);
    }
    return GVV_OK;
  }
  if (!strcmp(key, "cull_margin_milli")) {   // fixed part of the pre-test margin in 1/1000 px; < 0 = no culling
    h->cullMargin = value < 0 ? -1.f : (float)value / 1000.f;
    return GVV_OK;
  }
  if (!strcmp(key, "batch_div")) { if (value < 1) return fail(GVV_EINVAL, "batch_div must be >= 1"); h->batchDiv = value; return GVV_OK; }
  if (!strcmp(key, "cta_threads")) { if (value != 128 && value != 256) return fail(GVV_EINVAL, "cta_threads must be 128 or 256"); h->ctaThreads = value; return GVV_OK; }
  if (!strcmp(key, "span_z")) { if (value < 0 || value > 2) return fail(GVV_EINVAL, "span_z must be 0 (off), 1 (both passes) or 2 (far pass only)"); h->spanZ = value; return GVV_OK; }
  if (!strcmp(key, "cta_trace")) { if (h->captured) return fail(GVV_EINVAL, "cta_trace cannot change after the handle was used under stream capture"); cudaSetDevice(h->device); cudaDeviceSynchronize(); free_scratch(h->s); h->ctaTrace = value ? 1 : 0; return GVV_OK; }   // scratch is re-allocated by the next call
  if (!strcmp(key, "shared_batch_grads")) { h->sharedBatchGrads = value ? 1 : 0; return GVV_OK; }
  if (!strcmp(key, "texture_bilinear")) { h->texBilinear = value ? 1 : 0; return GVV_OK; }
  if (!strcmp(key, "bwd_persistent")) { h->bwdPersistent = value ? 1 : 0; return GVV_OK; }
  if (!strcmp(key, "bulk_out")) { h->bulkOut = value ? 1 : 0; return GVV_OK; }
  if (!strcmp(key, "chain")) { h->chain = value ? 1 : 0; return GVV_OK; }
  if (!strcmp(key, "resolve_prefetch")) { h->resolvePrefetch = value ? 1 : 0; return GVV_OK; }
  if (!strcmp(key, "spread_empty")) { h->spreadEmpty = value ? 1 : 0; return GVV_OK; }
  if (!strcmp(key, "heavy_mode")) { if (value < 0 || value > 2) return fail(GVV_EINVAL, "heavy_mode must be 0 (never), 1 (adaptive) or 2 (always)"); h->heavyMode = value; return GVV_OK; }
  if (!strcmp(key, "heavy_slots")) { if (value < 1 || value > 1024) return fail(GVV_EINVAL, "heavy_slots must be in [1,1024]"); h->heavySlots = value; return GVV_OK; }
  if (!strcmp(key, "heavy_thr")) { if (value < 0) return fail(GVV_EINVAL, "heavy_thr must be >= 0"); h->heavyThr = value; return GVV_OK; }
  if (!strcmp(key, "split_unit")) { if (value < 0) return fail(GVV_EINVAL, "split_unit must be >= 0"); h->splitUnit = value; return GVV_OK; }
  if (!strcmp(key, "hiz_min")) { if (value < 0) return fail(GVV_EINVAL, "hiz_min must be >= 0"); h->hizMin = value; return GVV_OK; }
  if (!strcmp(key, "hiz")) { h->hiz = value ? 1 : 0; return GVV_OK; }
  if (!strcmp(key, "interleave")) { h->interleave = value ? 1 : 0; return GVV_OK; }
  if (!strcmp(key, "ray_cache")) { h->rayCache = value ? 1 : 0; return GVV_OK; }
  if (!strcmp(key, "time_kernels")) {
    KernelTimer& t = h->timer;
    cudaSetDevice(h->device);
    if (value && !t.ev) {
      t.ev = new (std::nothrow) cudaEvent_t[2 * KernelTimer::kCap];
      t.slot = new (std::nothrow) int[KernelTimer::kCap];
      if (!t.ev || !t.slot) return fail(GVV_ENOMEM, "out of host memory");
      for (int i = 0; i < 2 * KernelTimer::kCap; ++i)
        if (cudaEventCreate(&t.ev[i]) != cudaSuccess) return fail(GVV_ECUDA, "cudaEventCreate failed");
    }
    t.used = 0;
    t.enabled = value != 0;
    return GVV_OK;
  }
  return fail(GVV_EINVAL, "unknown option '%s'", key);
}

extern "C" int64_t gvv_launch_count(gvv_handle h) { return h ? h->launches : 0; }

extern "C" int gvv_forward(gvv_handle h, int32_t B, int32_t texH, int32_t texW,
                           const float* vertex_pos, const float* vertex_color, const float* texture,
                           const float* sh_coeff, const float* target_image,
                           const float* extrinsics, const float* intrinsics,
                           float* bary, int32_t* face, float* render, float* vertex_normal,
                           float* target_image_out, float* normal_map, void* stream) {
  if (!h) return fail(GVV_EINVAL, "gvv_forward: null handle");
  if (B <= 0) return fail(GVV_EINVAL, "gvv_forward: batch must be positive");
  if (!vertex_pos || !sh_coeff || !extrinsics || !intrinsics || !vertex_normal)
    return fail(GVV_EINVAL, "gvv_forward: null input/output pointer");
  if (!h->computeNormalMap && (!bary || !face || !render)) return fail(GVV_EINVAL, "gvv_forward: null output pointer");
  if (h->albedo == GVV_ALBEDO_VERTEX_COLOR && !vertex_color) return fail(GVV_EINVAL, "gvv_forward: vertex_color is null");
  if (h->albedo == GVV_ALBEDO_TEXTURED && (!texture || texH <= 0 || texW <= 0)) return fail(GVV_EINVAL, "gvv_forward: texture missing");
  cudaStream_t st = (cudaStream_t)stream;
  CK(cudaSetDevice(h->device));
  const int rc = ensure_scratch(h, B, st);
  if (rc) return rc;

  FwdArgs a;
  a.B = B; a.C = h->C; a.N = h->N; a.F = h->F; a.W = h->W; a.H = h->H; a.texH = texH; a.texW = texW;
  a.albedo = h->albedo; a.shading = h->shading;
  a.tile = h->tile; a.tilesX = h->tilesX; a.tilesY = h->tilesY; a.nT = h->nT; a.cullMargin = h->cullMargin; a.rayCache = h->rayCache; a.batchDiv = h->batchDiv; a.ctaThreads = h->ctaThreads; a.interleave = h->interleave; a.hiz = h->hiz; a.hizMin = h->hizMin; a.spanZ = h->spanZ; a.splitUnit = h->splitUnit; a.heavyThr = h->heavyThr; a.heavySlots = h->heavySlots; a.heavyMode = h->heavyMode; a.ctaSlots = h->ctaSlots; a.spreadEmpty = h->spreadEmpty; a.texBilinear = h->texBilinear; a.resolvePrefetch = h->resolvePrefetch; a.bulkOut = h->bulkOut; a.chain = h->chain && !h->timer.enabled;   // per-kernel timing wants plain stream order
  a.vertex_pos = vertex_pos; a.vertex_color = vertex_color; a.texture = texture; a.sh_coeff = sh_coeff;
  a.extrinsics = extrinsics; a.intrinsics = intrinsics; a.texcoords = h->texcoords;
  a.faces4 = h->faces4; a.vfOffsets = h->vfOffsets; a.vfList = h->vfList;
  a.bary = bary; a.face = face; a.render = render; a.vertex_normal = vertex_normal;
  a.s = h->s;

  // out4 is a copy of in4 (CudaRenderer.cpp:286); non-blocking here
  if (target_image && target_image_out && target_image != target_image_out)
    CK(cudaMemcpyAsync(target_image_out, target_image, sizeof(float) * 3 * (size_t)B * h->C * h->W * h->H,
                       cudaMemcpyDeviceToDevice, st));

  int n;
  if (h->computeNormalMap) {
    // compute_normal_map replaces rasterisation (CUDABasedRasterization.cu:463-466)
    if (!normal_map || texH <= 0 || texW <= 0) return fail(GVV_EINVAL, "gvv_forward: normal_map output / texture size missing");
    if (!h->hasTexcoords) return fail(GVV_EINVAL, "gvv_forward: compute_normal_map needs texture_coordinates");
    if (!h->texelTable || h->tableH != texH || h->tableW != texW) {
      CK(cudaStreamSynchronize(st));
      cudaFree(h->texelTable); h->texelTable = nullptr;
      CK(dmalloc(&h->texelTable, (size_t)texH * texW));
      const int k = launch_build_texel_table(h->texcoords, h->F, texH, texW, h->texelTable, st);
      if (k < 0) return fail(GVV_ECUDA, "texel table build failed: %s", cudaGetErrorString(cudaGetLastError()));
      h->launches += k;
      h->tableH = texH; h->tableW = texW;
    }
    n = launch_normal_map(a, h->texelTable, normal_map, st);
  } else {
    n = launch_forward(a, st, &h->timer);
  }
  if (n < 0) return fail(GVV_ECUDA, "forward launch failed: %s", cudaGetErrorString(cudaGetLastError()));
  h->launches += n;
  return GVV_OK;
}

extern "C" int gvv_backward(gvv_handle h, int32_t B, int32_t texH, int32_t texW,
                            const float* render_grad,
                            const float* vertex_pos, const float* vertex_color, const float* texture,
                            const float* sh_coeff, const float* target_image, const float* vertex_normal,
                            const float* bary, const int32_t* face, const float* target_grad,
                            const float* extrinsics, const float* intrinsics,
                            float* vpos_grad, float* vcol_grad, float* tex_grad, float* sh_grad, void* stream) {
  if (!h) return fail(GVV_EINVAL, "gvv_backward: null handle");
  if (B <= 0) return fail(GVV_EINVAL, "gvv_backward: batch must be positive");
  // the Python gradient only calls the op for these modes (CudaRenderer.py:175); the kernel
  // itself prints "Unsupported color mode" otherwise (CUDABasedRasterizationGrad.cu:391-394)
  if (h->albedo == GVV_ALBEDO_NORMAL || h->albedo == GVV_ALBEDO_LIGHTING)
    return fail(GVV_EUNSUPPORTED, "Unsupported color mode in renderer gradient!");
  if (!render_grad || !vertex_pos || !sh_coeff || !vertex_normal || !bary || !face || !extrinsics || !intrinsics ||
      !vpos_grad || !vcol_grad || !sh_grad)
    return fail(GVV_EINVAL, "gvv_backward: null input/output pointer");
  if (h->albedo != GVV_ALBEDO_FOREGROUND_MASK && !vertex_color) return fail(GVV_EINVAL, "gvv_backward: vertex_color is null");
  if (h->albedo == GVV_ALBEDO_TEXTURED && (!texture || !tex_grad || texH <= 0 || texW <= 0))
    return fail(GVV_EINVAL, "gvv_backward: texture / texture_grad missing");
  if (target_grad && !target_image) return fail(GVV_EINVAL, "gvv_backward: target_buffer_grad given without target_image");
  cudaStream_t st = (cudaStream_t)stream;
  CK(cudaSetDevice(h->device));
  const int rc = ensure_scratch(h, B, st);
  if (rc) return rc;

  BwdArgs a;
  a.B = B; a.C = h->C; a.N = h->N; a.F = h->F; a.W = h->W; a.H = h->H; a.texH = texH; a.texW = texW;
  a.albedo = h->albedo; a.shading = h->shading; a.imgFilter = h->imgFilter; a.texBilinear = h->texBilinear; a.chain = h->chain && !h->timer.enabled; a.sharedBatch = h->sharedBatchGrads; a.bwdPersistent = h->bwdPersistent; a.ctaSlots = h->ctaSlots; a.target_du = h->targetDu; a.target_dv = h->targetDv;
  a.render_grad = render_grad; a.target_grad = target_grad; a.vertex_pos = vertex_pos; a.vertex_color = vertex_color;
  a.texture = texture; a.sh_coeff = sh_coeff; a.target_image = target_image; a.vertex_normal = vertex_normal;
  a.bary = bary; a.extrinsics = extrinsics; a.intrinsics = intrinsics; a.texcoords = h->texcoords;
  a.face = face; a.faces4 = h->faces4; a.vfOffsets = h->vfOffsets; a.vfList = h->vfList;
  a.vpos_grad = vpos_grad; a.vcol_grad = vcol_grad; a.tex_grad = tex_grad; a.sh_grad = sh_grad;
  a.ar = h->ar; a.arAfter = h->arAfter;
  a.s = h->s;
  const int n = launch_backward(a, st, &h->timer);
  if (n < 0) return fail(GVV_ECUDA, "backward launch failed: %s", cudaGetErrorString(cudaGetLastError()));
  h->launches += n;
  return GVV_OK;
}

extern "C" int gvv_set_allreduce(gvv_handle h, const gvv_allreduce_desc* d) {
  if (!h) return fail(GVV_EINVAL, "gvv_set_allreduce: null handle");
  if (!d) { h->ar = ARParams{}; h->arAfter = 0; return GVV_OK; }
  if (d->world < 1 || d->world > kMaxWorld || d->rank < 0 || d->rank >= d->world) return fail(GVV_EINVAL, "gvv_set_allreduce: bad rank / world (at most %d ranks)", kMaxWorld);
  if (!d->peer_buffers || !d->signal_pads || !d->result) return fail(GVV_EINVAL, "gvv_set_allreduce: null pointer");
  if (d->count_floats <= 0 || d->offset_floats < 0 || (d->offset_floats & 3)) return fail(GVV_EINVAL, "gvv_set_allreduce: the range must be non-empty and start at a multiple of 4 floats");
  if (reinterpret_cast<uintptr_t>(d->result) & 15) return fail(GVV_EINVAL, "gvv_set_allreduce: result must be 16-byte aligned");
  if (d->mode != 0 && d->mode != 1) return fail(GVV_EINVAL, "gvv_set_allreduce: mode must be 0 (peer loads) or 1 (NVLS multimem.ld_reduce)");
  if (d->mode == 1 && (!d->multicast_ptr || (d->count_floats & 3))) return fail(GVV_EINVAL, "gvv_set_allreduce: NVLS mode needs a multicast mapping and a count that is a multiple of 4");
  if (d->channels < 1 || d->channels > 64 || d->first_channel < 0) return fail(GVV_EINVAL, "gvv_set_allreduce: channels must be in [1,64]");
  if (d->epoch_word < (d->first_channel + d->channels) * d->world) return fail(GVV_EINVAL, "gvv_set_allreduce: epoch_word overlaps the signal words");
  ARParams a;
  a.peers = reinterpret_cast<const float* const*>(d->peer_buffers);
  a.pads = reinterpret_cast<uint32_t* const*>(d->signal_pads);
  a.mc = reinterpret_cast<const float*>(d->multicast_ptr);
  a.result = d->result; a.offset = d->offset_floats; a.count = d->count_floats;
  a.rank = d->rank; a.world = d->world; a.mode = d->mode; a.blocks = d->channels; a.channelBase = d->first_channel; a.epochBase = d->epoch_word;
  h->ar = a; h->arAfter = d->after_backward ? 1 : 0;
  return GVV_OK;
}

extern "C" int gvv_set_target_gradient(gvv_handle h, const float* d_du, const float* d_dv) {
  if (!h) return fail(GVV_EINVAL, "gvv_set_target_gradient: null handle");
  if ((d_du == nullptr) != (d_dv == nullptr)) return fail(GVV_EINVAL, "gvv_set_target_gradient: pass both gradients or neither");
  h->targetDu = d_du; h->targetDv = d_dv;
  return GVV_OK;
}

extern "C" int64_t gvv_debug_copy(gvv_handle h, int32_t which, void* dst, int64_t capacity, void* stream) {
  if (!h || !dst) return -1;
  cudaStream_t st = (cudaStream_t)stream;
  cudaSetDevice(h->device);
  const void* src = nullptr;
  int64_t bytes = 0;
  if (which == 0) { src = h->s.cams; bytes = (int64_t)h->s.capViews * sizeof(CamRec); }
  else if (which == 1) { src = h->s.proj; bytes = (int64_t)h->s.capViews * h->N * sizeof(float4); }
  else if (which == 2) { src = h->s.ctaTrace; bytes = src ? (int64_t)h->s.capViews * (h->nT + h->nT / 2) * 4 * sizeof(unsigned long long) : 0; }
  else return -1;
  if (!src) return 0;
  const int64_t n = bytes < capacity ? bytes : capacity;
  if (cudaMemcpyAsync(dst, src, (size_t)n, cudaMemcpyDeviceToHost, st) != cudaSuccess) return -1;
  if (cudaStreamSynchronize(st) != cudaSuccess) return -1;
  return bytes;
}

extern "C" int gvv_debug_eval(gvv_handle h, int32_t n, const int32_t* queries, int32_t* out_key, float* out_ab, void* stream) {
  if (!h || n < 0 || (n > 0 && (!queries || !out_key || !out_ab))) return fail(GVV_EINVAL, "gvv_debug_eval: bad argument");
  if (n == 0) return GVV_OK;
  if (!h->s.cams) return fail(GVV_EINVAL, "gvv_debug_eval: no forward call yet");
  for (int i = 0; i < n; ++i) {
    const int32_t* q = queries + 4 * i;
    if (q[0] < 0 || q[0] >= h->s.capViews || q[1] < 0 || q[1] >= h->W || q[2] < 0 || q[2] >= h->H || q[3] < 0 || q[3] >= h->F)
      return fail(GVV_EINVAL, "gvv_debug_eval: query %d out of range", i);
  }
  cudaStream_t st = (cudaStream_t)stream;
  CK(cudaSetDevice(h->device));
  int *dq = nullptr, *dk = nullptr; float* da = nullptr;
  CK(cudaMalloc((void**)&dq, sizeof(int) * 4 * (size_t)n));
  CK(cudaMalloc((void**)&dk, sizeof(int) * (size_t)n));
  CK(cudaMalloc((void**)&da, sizeof(float) * 2 * (size_t)n));
  CK(cudaMemcpyAsync(dq, queries, sizeof(int) * 4 * (size_t)n, cudaMemcpyHostToDevice, st));
  const int k = launch_debug_eval(h->s, h->faces4, h->N, h->C, h->W, h->H, n, dq, dk, da, st);
  if (k > 0) h->launches += k;
  CK(cudaMemcpyAsync(out_key, dk, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(out_ab, da, sizeof(float) * 2 * (size_t)n, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  cudaFree(dq); cudaFree(dk); cudaFree(da);
  return k > 0 ? GVV_OK : fail(GVV_ECUDA, "debug_eval launch failed");
}

static const char* kKernelNames[K_NUM_SLOTS] = {"camera_kernel", "vertex_kernel", "bin_count_kernel", "bin_scan_kernel",
                                                "bin_fill_kernel", "raster_kernel", "prep_kernel", "pixel_grad_kernel",
                                                "normal_term_kernel", "normal_map_kernel", "allreduce_kernel"};

extern "C" int32_t gvv_kernel_count(void) { return K_NUM_SLOTS; }
extern "C" const char* gvv_kernel_name(int32_t i) { return (i >= 0 && i < K_NUM_SLOTS) ? kKernelNames[i] : ""; }

extern "C" int gvv_kernel_times(gvv_handle h, double* total_ms, int64_t* launches) {
  if (!h || !total_ms || !launches) return fail(GVV_EINVAL, "gvv_kernel_times: null argument");
  for (int i = 0; i < K_NUM_SLOTS; ++i) { total_ms[i] = 0.0; launches[i] = 0; }
  KernelTimer& t = h->timer;
  CK(cudaSetDevice(h->device));
  for (int r = 0; r < t.used; ++r) {
    CK(cudaEventSynchronize(t.ev[2 * r + 1]));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, t.ev[2 * r], t.ev[2 * r + 1]));
    total_ms[t.slot[r]] += ms;
    launches[t.slot[r]] += 1;
  }
  t.used = 0;
  return GVV_OK;
}
