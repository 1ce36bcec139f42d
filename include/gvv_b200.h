/*
 * gvv_b200.h -- C ABI of the B200-native differentiable rasteriser.
 *
 * Drop-in boundary for the one hot path of ayushtewari/GVV-Differentiable-CUDA-Renderer:
 * the TensorFlow custom ops `CudaRendererGpu` / `CudaRendererGradGpu`.  Every entry point
 * below replaces one piece of the reference's op boundary (file:line are relative to the
 * reference checkout):
 *
 *   gvv_create    <- CudaRenderer::CudaRenderer(OpKernelConstruction*)       cpp/src/TensorflowOperators/CudaRenderer/CudaRenderer.cpp:37-155
 *                    + CUDABasedRasterization ctor (topology upload, CSR)    cpp/src/Renderer/CUDABasedRasterization.cpp:11-106,125-154
 *                    + CudaRendererGrad ctor / CUDABasedRasterizationGrad    cpp/src/TensorflowOperators/CudaRenderer/CudaRendererGrad.cpp:43-101
 *   gvv_forward   <- CudaRenderer::Compute (per-batch host loop)             CudaRenderer.cpp:298-335, op signature :5-33
 *                    -> renderBuffersGPU                                     cpp/src/Renderer/CUDABasedRasterization.cu:449-473
 *   gvv_backward  <- CudaRendererGrad::Compute                               CudaRendererGrad.cpp:252-292, op signature :6-39
 *                    -> renderBuffersGradGPU                                 cpp/src/Renderer/CUDABasedRasterizationGrad.cu:624-635
 *   gvv_destroy   <- ~CUDABasedRasterization / ~CUDABasedRasterizationGrad   CUDABasedRasterization.cpp:110-121
 *   gvv_last_error<- replaces cutilSafeCall -> exit(-1)                      cpp/thirdParty/Shared/cutil/inc/cutil_inline_runtime.h:278-285
 *
 * Conventions
 *   - plain C, no torch / TF types; all tensors are dense row-major fp32 (int32 for faces
 *     and the face buffer) exactly as the reference op lays them out;
 *   - every pointer passed to gvv_forward / gvv_backward is a DEVICE pointer on the device
 *     the handle was created for; the caller owns all of them;
 *   - all work is enqueued on `stream` (a cudaStream_t passed as void*; NULL = legacy default
 *     stream); the calls never synchronise the host and never allocate once the scratch for
 *     the largest batch seen so far exists;
 *   - a handle is not re-entrant (same as one TF kernel instance) and is used from ONE stream at a time: its scratch
 *     (projected vertices, bins, self-cleaning tile counters) is shared by all calls, so calls on different streams
 *     must be ordered by the caller (events) -- concurrent streams need one handle each;
 *   - scratch grows (device-wide synchronisation, free + allocate) when a call's batch exceeds every batch seen
 *     before.  gvv_reserve sizes it up front.  Growth is refused (GVV_EINVAL) while the stream is being captured
 *     and after the handle has been used under capture: CUDA graphs bake the scratch pointers in;
 *   - return value 0 = success; otherwise a GVV_E* code and gvv_last_error() (thread-local)
 *     describes it.  The library never calls exit() and never throws across the ABI.
 */
#ifndef GVV_B200_H
#define GVV_B200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* albedo_mode / shading_mode: same order as the reference enums
 * (cpp/src/Renderer/CUDABasedRasterizationInput.h:25-35). */
enum gvv_albedo_mode {
  GVV_ALBEDO_VERTEX_COLOR    = 0,  /* "vertexColor"    */
  GVV_ALBEDO_TEXTURED        = 1,  /* "textured"       */
  GVV_ALBEDO_NORMAL          = 2,  /* "normal"         */
  GVV_ALBEDO_LIGHTING        = 3,  /* "lighting"       */
  GVV_ALBEDO_FOREGROUND_MASK = 4   /* "foregroundMask" (forces shadeless, CudaRenderer.cpp:72-76) */
};
enum gvv_shading_mode {
  GVV_SHADING_SHADED    = 0,       /* "shaded"    */
  GVV_SHADING_SHADELESS = 1        /* "shadeless" */
};

enum gvv_status {
  GVV_OK = 0,
  GVV_EINVAL = 1,     /* bad argument (the reference's OP_REQUIRES / silent early returns) */
  GVV_ECUDA = 2,      /* a CUDA runtime call failed */
  GVV_ENOMEM = 3,     /* scratch allocation failed */
  GVV_EUNSUPPORTED = 4
};

/* Attributes of the op (CudaRenderer.cpp:23-33).  `faces` / `texcoords` are HOST pointers,
 * copied at create time. */
typedef struct gvv_desc {
  const int32_t* faces;          /* [num_faces*3] vertex indices                        (attr faces)               */
  int32_t        num_faces;
  const float*   texcoords;      /* [num_faces*3*2] per-corner (u,v); may be NULL       (attr texture_coordinates) */
  int32_t        num_vertices;   /*                                                      (attr number_of_vertices)  */
  int32_t        num_cameras;    /*                                                      (attr number_of_cameras)   */
  int32_t        width;          /* render_resolution_u                                                             */
  int32_t        height;         /* render_resolution_v                                                             */
  int32_t        albedo_mode;    /* enum gvv_albedo_mode                                                            */
  int32_t        shading_mode;   /* enum gvv_shading_mode                                                           */
  int32_t        image_filter_size;    /* half-width of the target-image gradient filter (backward, B8)             */
  int32_t        texture_filter_size;  /* accepted and ignored, as in the reference's live code                     */
  int32_t        compute_normal_map;   /* 1: UV-space normal map INSTEAD of rasterisation (CUDABasedRasterization.cu:463-466) */
  int32_t        device;         /* CUDA device ordinal                                                             */
} gvv_desc;

typedef struct gvv_renderer* gvv_handle;

int gvv_create(const gvv_desc* desc, gvv_handle* out);
int gvv_destroy(gvv_handle h);

/* Allocates the scratch for calls of up to `max_batch` batch elements now (replaces the scratch cudaMallocs of the
 * reference constructors, CUDABasedRasterization.cpp:26-101, which size theirs for ONE batch element because the op
 * loops over the batch on the host).  Call it before capturing gvv_forward / gvv_backward into a CUDA graph with a
 * batch larger than any eager call has used.  Synchronises the device when it has to (re)allocate. */
int gvv_reserve(gvv_handle h, int32_t max_batch, void* stream);

/* Forward: inputs in0..in6 and outputs out0..out5 of CudaRendererGpu (CudaRenderer.cpp:5-21).
 *   vertex_pos   [B,N,3]   vertex_color [B,N,3]   texture [B,texH,texW,3]   sh_coeff [B,C,27]
 *   target_image [B,C,H,W,3] (may be NULL together with target_image_out)
 *   extrinsics   [B,C*12]  intrinsics   [B,C*9]
 *   barycentric_buffer [B,C,H,W,2]   face_buffer int32 [B,C,H,W]   render_buffer [B,C,H,W,3]
 *   vertex_normal [B,C,N,3]   target_image_out [B,C,H,W,3] (copy of target_image; skipped when
 *   it aliases target_image or is NULL)   normal_map [B,texH,texW,3] (written only when
 *   compute_normal_map; may be NULL otherwise) */
int gvv_forward(gvv_handle h, int32_t batch, int32_t tex_h, int32_t tex_w,
                const float* vertex_pos, const float* vertex_color, const float* texture,
                const float* sh_coeff, const float* target_image,
                const float* extrinsics, const float* intrinsics,
                float* barycentric_buffer, int32_t* face_buffer, float* render_buffer,
                float* vertex_normal, float* target_image_out, float* normal_map,
                void* stream);

/* Backward: inputs and outputs of CudaRendererGradGpu (CudaRendererGrad.cpp:6-28).
 *   render_buffer_grad [B,C,H,W,3]; target_buffer_grad [B,C,H,W,3] or NULL (= all zeros: the
 *   model-to-data term B8 is skipped); the rest as in gvv_forward.
 *   Outputs (overwritten, not accumulated): vertex_pos_grad [B,N,3], vertex_color_grad [B,N,3],
 *   texture_grad [B,texH,texW,3], sh_coeff_grad [B,C,27]. */
int gvv_backward(gvv_handle h, int32_t batch, int32_t tex_h, int32_t tex_w,
                 const float* render_buffer_grad,
                 const float* vertex_pos, const float* vertex_color, const float* texture,
                 const float* sh_coeff, const float* target_image, const float* vertex_normal,
                 const float* barycentric_buffer, const int32_t* face_buffer,
                 const float* target_buffer_grad,
                 const float* extrinsics, const float* intrinsics,
                 float* vertex_pos_grad, float* vertex_color_grad, float* texture_grad,
                 float* sh_coeff_grad,
                 void* stream);

const char* gvv_last_error(void);

/* ---- multi-GPU: one-shot all-reduce of shared-parameter gradients over peer memory ----------------------------
 * The reference has no multi-GPU path (python/utils/CheckGPU.py:51-52 masks all but one GPU).  Views are
 * sharded over one process per GPU; the only exchange is the sum of gradients of parameters several ranks share.
 * With a descriptor set, gvv_backward ends by reducing the float range [offset_floats, offset_floats + count_floats)
 * of every rank's SYMMETRIC buffer (each rank passes gradient output pointers inside its own buffer to gvv_backward)
 * into `result` on every rank:  result[i] = sum over ranks r of peer_buffers[r][offset_floats + i], rank order 0..W-1.
 *   peer_buffers   DEVICE array [world] of the base addresses, as mapped in THIS process, of every rank's buffer
 *   signal_pads    DEVICE array [world] of uint32 signal pads, zero-initialised and from then on touched only by this
 *                  library: words (first_channel + c) * world + r, c < channels, are monotonic barrier counters, words
 *                  epoch_word + c hold each CTA's barrier number (epoch_word >= (first_channel + channels) * world)
 *   multicast_ptr  NVLS multicast mapping of the buffer (0 if none); mode 1 reads the sum from the switch
 *   after_backward 0: the range holds only gradients that are final after the per-pixel kernel (SH, colours, texture):
 *                  the exchange runs as `channels` CTAs INSIDE the backward's last kernel, overlapped with its math;
 *                  1: the range includes vertex_pos_grad: a kernel of its own follows the backward
 * The caller alternates between two ranges (slots) from step to step: a slot is rewritten two barriers after it was
 * read.  NULL removes the descriptor.  Buffers, pads and the pointer arrays stay owned by the caller. */
typedef struct gvv_allreduce_desc {
  const void* peer_buffers;
  const void* signal_pads;
  uint64_t    multicast_ptr;
  int32_t     rank, world;
  int64_t     offset_floats, count_floats;
  float*      result;
  int32_t     mode;            /* 0 = peer loads over NVLink (ld.global.sys), 1 = NVLS multimem.ld_reduce */
  int32_t     channels;        /* CTAs taking part, 1..64 */
  int32_t     first_channel;
  int32_t     epoch_word;
  int32_t     after_backward;
} gvv_allreduce_desc;
int gvv_set_allreduce(gvv_handle h, const gvv_allreduce_desc* desc);

/* ---- loss-side helpers next to the op; no handle needed ----------------- */

/* smoothImage (python/utils/GaussianSmoothingGpu.py:12-37): depthwise Gaussian over `images` dense
 * [height,width,3] fp32 images with zero "SAME" padding, as two separable passes.  taps: HOST
 * float[2*half_size+1], the normalised 1-D kernel (the reference's 2-D kernel is outer(vals,vals)/sum);
 * in/tmp/out: DEVICE buffers of images*height*width*3 floats (tmp is scratch; out may alias in, not tmp).
 * Cross-correlation like tf.nn.depthwise_conv2d, so the adjoint is the same call with reversed taps. */
int gvv_gaussian_smooth(int32_t device, int64_t images, int32_t height, int32_t width, int32_t half_size,
                        const float* taps, const float* in, float* tmp, float* out, void* stream);

/* imageGradient (cpp/src/Utils/RendererUtil.h:566-620) of `images` dense [height,width,3] images:
 * d_du, d_dv (DEVICE, same shape) = dI/du and dI/dv, zero within filter_size+1 pixels of the border.
 * The backward evaluates exactly this per covered pixel on every call when target_buffer_grad is
 * given; for a constant target, compute it once and pass it with gvv_set_target_gradient. */
int gvv_image_gradient(int32_t device, int64_t images, int32_t height, int32_t width, int32_t filter_size,
                       const float* image, float* d_du, float* d_dv, void* stream);

/* Caller-owned precomputed target-image gradient [B,C,H,W,3] x 2 for the model-to-data term of
 * gvv_backward (must match the handle's image_filter_size and the target passed to gvv_backward);
 * NULL, NULL = recompute per pixel like the reference (default). */
int gvv_set_target_gradient(gvv_handle h, const float* d_du, const float* d_dv);

/* ---- diagnostics (not part of the reference boundary) ------------------------------------ */

/* Number of kernels the library launched on behalf of this handle since creation. */
int64_t gvv_launch_count(gvv_handle h);

/* Copies an internal buffer of the LAST forward call to host memory (synchronises `stream`).
 * which: 0 = per-view camera records (64 floats each: K[9] E[12] Einv[16] Pinv[16] ro[3] ros[3] pad),
 *        1 = projected vertices float4[V*N] (x/z, y/z, z, 0).
 * Returns the number of bytes the buffer holds (copies min(bytes, capacity)). */
int64_t gvv_debug_copy(gvv_handle h, int32_t which, void* host_dst, int64_t capacity, void* stream);

/* Re-evaluates the exact (pixel, triangle) test on the geometry of the LAST forward call, for
 * tie analysis in the parity tests.  queries: HOST int32[n*4] = (view, x, y, face);
 * out_key: HOST int32[n] depth key as the reference's atomicMin sees it, INT32_MIN when the pair
 * fails the inside test; out_ab: HOST float[n*2] barycentrics (a,b).  Synchronises `stream`. */
int gvv_debug_eval(gvv_handle h, int32_t n, const int32_t* queries, int32_t* out_key, float* out_ab, void* stream);

/* Runtime knobs.  Behaviour: "shared_batch_grads" (1: vertex_color_grad, texture_grad and sh_coeff_grad are accumulated
 * over the batch into outputs of batch extent ONE -- [1,N,3], [1,texH,texW,3], [1,C,27] -- for parameters the caller
 * shares across the batch, e.g. one template colour set / texture / illumination for all batch elements; the reference
 * leaves that sum to the framework's autodiff of a broadcast; vertex_pos_grad stays per batch element; default 0),
 * "texture_bilinear" (1: the bilinear texture fetch and the weighted 4-texel
 * texture-gradient scatter the reference has commented out, CUDABasedRasterization.cu:365-372,
 * CUDABasedRasterizationGrad.cu:361-378; default 0 = reference behaviour: nearest texel, unweighted add).
 * Scheduling / culling (results are bit-identical for every setting, see tests; defaults = measured best on a
 * B200): "tile" (16|32 rasteriser tile edge), "cull_margin_milli" (fixed part, in 1/1000 pixel, of the margin of
 * the conservative screen-space pre-test that decides which bbox pixels get the exact test; default 62 = 1/16 px;
 * negative = test every bbox pixel exactly, like the reference), "hiz" (two depth passes per tile), "hiz_min" (bins shorter than this: one pass), "span_z"
 * (0|1|2 span-level early z: off / both passes / far pass), "batch_div", "cta_threads" (128|256), "interleave",
 * "ray_cache", "heavy_mode" (0|1|2 1024-thread CTAs for critical-path tiles: never / decided on the GPU for calls
 * of at most two views / always), "heavy_thr", "heavy_slots", "split_unit", "spread_empty", "resolve_prefetch",
 * "chain" (programmatic dependent launch of the kernels of a call).  Diagnostics: "time_kernels" (1: record a
 * CUDA-event pair around every kernel on the launching stream, 0: off; either resets the log), "cta_trace".
 * The environment variable GVV_OPTIONS="key=value,..." applies knobs to every handle at creation.
 * Returns 0 on success. */
int gvv_set_option(gvv_handle h, const char* key, int32_t value);

/* Per-kernel device time accumulated since "time_kernels" was enabled (synchronises): total_ms and
 * launches are arrays of gvv_kernel_count() entries, named by gvv_kernel_name(i).  Clears the log. */
int32_t gvv_kernel_count(void);
const char* gvv_kernel_name(int32_t i);
int gvv_kernel_times(gvv_handle h, double* total_ms, int64_t* launches);

/* Atomic-throughput micro-benchmark used for the roofline denominators:
 * kind 0: red.global.min.u64 over `n_addr` 64-bit words, kind 1: red.global.add.f32 over
 * `n_addr` floats; `n_ops` operations with pseudo-random addresses, one per thread.
 * Returns operations per second through *ops_per_s (device-timed with CUDA events). */
int gvv_bench_atomics(int32_t device, int32_t kind, int64_t n_addr, int64_t n_ops, int32_t iters,
                      double* ops_per_s);

#ifdef __cplusplus
}
#endif
#endif /* GVV_B200_H */
