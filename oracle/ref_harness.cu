// ref_harness.cu -- TEST INFRASTRUCTURE ONLY (Oracle 1: "the reference's CUDA rebuilt").
//
// A TensorFlow-free driver for the UNMODIFIED reference renderer core.  It is compiled together
// with the reference's own sources, read in place from /root/reference (never copied into this
// repository), by oracle/build_ref.sh into oracle/_ref/libgvv_ref.so.  It replays exactly the
// pointer arithmetic of the TF op boundary:
//     forward   CudaRenderer::Compute        cpp/src/TensorflowOperators/CudaRenderer/CudaRenderer.cpp:298-335
//     backward  CudaRendererGrad::Compute    cpp/src/TensorflowOperators/CudaRenderer/CudaRendererGrad.cpp:252-292
// and the foregroundMask -> shadeless rule of the op constructors (CudaRenderer.cpp:72-76,
// CudaRendererGrad.cpp:78-82).  Only tests/, __graft_entry__.smoke() and bench.py's reference arm
// may load the resulting library; the product (libgvv_b200.so) never does.
//
// Caveats inherited from the reference: CUDA errors exit(-1) (cutilSafeCall); the destructors
// cudaFree caller-owned camera pointers, so handles are intentionally never destroyed.
#define private public   // the intermediate buffers (inverse matrices) have no getters
#include "Renderer/CUDABasedRasterization.h"
#include "Renderer/CUDABasedRasterizationGrad.h"
#undef private
#include <string>
#include <vector>

struct RefHandle {
  CUDABasedRasterization* fwd;
  CUDABasedRasterizationGrad* bwd;
  int N, C, W, H;
};

extern "C" void* gvvref_create(const int* faces, int F, const float* texcoords, int N, int C, int W, int H,
                               const char* albedo, const char* shading, int imageFilter, int textureFilter,
                               int computeNormal, int withBackward) {
  std::vector<int> f(faces, faces + 3 * (size_t)F);
  std::vector<float> t(texcoords, texcoords + 6 * (size_t)F);
  std::string a(albedo), s(shading);
  if (a == "foregroundMask") s = "shadeless";
  RefHandle* h = new RefHandle();
  h->N = N; h->C = C; h->W = W; h->H = H;
  h->fwd = new CUDABasedRasterization(f, t, N, C, W, H, a, s, computeNormal != 0);
  h->bwd = withBackward ? new CUDABasedRasterizationGrad(f, t, N, C, W, H, a, s, imageFilter, textureFilter) : nullptr;
  return h;
}

// All pointers are device pointers laid out as the op's tensors.  depth_copy (optional, int32
// [B,C,H,W]) and cam_copy (optional, float [B,C,32]: Einv[16] then Pinv[16]) receive the
// reference's intermediate buffers after each batch element, for stage-by-stage bit comparison.
extern "C" int gvvref_forward(void* hv, int B, int texH, int texW,
                              const float* vpos, const float* vcol, const float* tex, const float* sh,
                              const float* target, const float* extr, const float* intr,
                              float* bary, int* face, float* render, float* vnormal, float* target_out, float* normal_map,
                              int* depth_copy, float* cam_copy, float* proj_copy, int* bbox_copy) {
  RefHandle* h = (RefHandle*)hv;
  CUDABasedRasterization* r = h->fwd;
  const int N = h->N, C = h->C, W = h->W, H = h->H;
  if (target && target_out && target != target_out)
    cudaMemcpy(target_out, target, sizeof(float) * (size_t)B * C * H * W * 3, cudaMemcpyDeviceToDevice);
  r->setTextureWidth(texW);
  r->setTextureHeight(texH);
  for (int b = 0; b < B; b++) {
    r->set_D_vertices((float3*)vpos + (size_t)b * N);
    r->set_D_vertexColors((float3*)vcol + (size_t)b * N);
    r->set_D_textureMap(tex + (size_t)b * texH * texW * 3);
    r->set_D_shCoeff(sh + (size_t)b * C * 27);
    r->set_D_extrinsics(extr + (size_t)b * C * 12);
    r->set_D_intrinsics(intr + (size_t)b * C * 9);
    r->set_D_barycentricCoordinatesBuffer(bary + (size_t)b * C * H * W * 2);
    r->set_D_faceIDBuffer(face + (size_t)b * C * H * W);
    r->set_D_renderBuffer(render + (size_t)b * C * H * W * 3);
    r->set_D_vertexNormal((float3*)vnormal + (size_t)b * C * N);
    r->set_D_normalMap((float3*)normal_map + (size_t)b * texW * texH);
    r->renderBuffers();
    if (depth_copy)
      cudaMemcpy(depth_copy + (size_t)b * C * H * W, r->get_D_depthBuffer(), sizeof(int) * (size_t)C * H * W, cudaMemcpyDeviceToDevice);
    if (cam_copy)
      for (int c = 0; c < C; c++) {
        cudaMemcpy(cam_copy + ((size_t)b * C + c) * 32, r->input.d_inverseExtrinsics + 4 * c, 64, cudaMemcpyDeviceToDevice);
        cudaMemcpy(cam_copy + ((size_t)b * C + c) * 32 + 16, r->input.d_inverseProjection + 4 * c, 64, cudaMemcpyDeviceToDevice);
      }
    if (proj_copy)
      cudaMemcpy(proj_copy + (size_t)b * C * N * 3, r->get_D_projectedVertices(), sizeof(float) * 3 * (size_t)C * N, cudaMemcpyDeviceToDevice);
    if (bbox_copy)
      cudaMemcpy(bbox_copy + (size_t)b * C * r->getNumberOfFaces() * 4, r->get_D_BBoxes(), sizeof(int) * 4 * (size_t)C * r->getNumberOfFaces(), cudaMemcpyDeviceToDevice);
  }
  return (int)cudaGetLastError();
}

extern "C" int gvvref_backward(void* hv, int B, int texH, int texW,
                               const float* render_grad, const float* vpos, const float* vcol, const float* tex,
                               const float* sh, const float* target, const float* vnormal, const float* bary,
                               const int* face, const float* target_grad, const float* extr, const float* intr,
                               float* vpos_grad, float* vcol_grad, float* tex_grad, float* sh_grad) {
  RefHandle* h = (RefHandle*)hv;
  CUDABasedRasterizationGrad* g = h->bwd;
  if (!g) return -1;
  const int N = h->N, C = h->C, W = h->W, H = h->H;
  for (int b = 0; b < B; b++) {
    g->setTextureWidth(texW);
    g->setTextureHeight(texH);
    g->set_D_RenderBufferGrad((float3*)render_grad + (size_t)b * C * H * W);
    g->set_D_TargetBufferGrad((float3*)target_grad + (size_t)b * C * H * W);
    g->set_D_vertices((float3*)vpos + (size_t)b * N);
    g->set_D_vertexColors((float3*)vcol + (size_t)b * N);
    g->set_D_textureMap(tex + (size_t)b * texH * texW * 3);
    g->set_D_shCoeff(sh + (size_t)b * C * 27);
    g->set_D_vertexNormal((float3*)vnormal + (size_t)b * C * N);
    g->set_D_barycentricCoordinatesBuffer((float2*)bary + (size_t)b * C * H * W);
    g->set_D_faceIDBuffer((int*)face + (size_t)b * C * H * W);
    g->set_D_targetImage(target + (size_t)b * C * H * W * 3);
    g->set_D_extrinsics(extr + (size_t)b * C * 12);
    g->set_D_intrinsics(intr + (size_t)b * C * 9);
    g->set_D_vertexPosGrad((float3*)vpos_grad + (size_t)b * N);
    g->set_D_vertexColorGrad((float3*)vcol_grad + (size_t)b * N);
    g->set_D_textureGrad((float3*)tex_grad + (size_t)b * texW * texH);
    g->set_D_shCoeffGrad(sh_grad + (size_t)b * C * 27);
    g->renderBuffersGrad();
  }
  return (int)cudaGetLastError();
}
