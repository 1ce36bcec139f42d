// gvv_normalmap.cu -- UV-space normal map (compute_normal_map), SURVEY.md 8(f) row 2.
//
// Reference: the first renderBuffers() call rasterises every triangle in UV space ON THE HOST
// (CUDABasedRasterization.cpp:237-298, OpenMP, racy where UV triangles overlap) into a per-texel
// (face, a, b, c) table, then renderNormalMapDevice (CUDABasedRasterization.cu:415-445) shades it
// INSTEAD of rasterising (:463-466).  Here the table is built on the GPU, once per texture size:
//   texel_face_kernel   one thread per triangle: same +-2 texel bbox, same ray (texel centre, z=1,
//                       direction -z) / triangle test with the /1000 pre-scale; atomicMax(face id)
//                       so overlaps resolve to the highest face id (what the reference's loop gives
//                       when run serially)
//   texel_bary_kernel   one thread per texel: recompute (a,b,c) of the winning face
//   normal_map_kernel   one thread per texel: interpolate camera-0 vertex normals, normalise if
//                       non-zero, map to [0,1]
#include "gvv_internal.h"

namespace gvv {

// The reference builds the table ON THE HOST (x86-64, SSE scalar fp32, no FMA contraction): every operation below
// is therefore a separately rounded IEEE single operation in the reference's source order, pinned with __f*_rn
// intrinsics so that nvcc cannot fuse them -- the (a, b, c) of a texel are then bit-identical to the host's.
struct H3 { float x, y, z; };
__device__ __forceinline__ H3 h3(float x, float y, float z) { H3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ H3 hsub(H3 a, H3 b) { return h3(__fsub_rn(a.x, b.x), __fsub_rn(a.y, b.y), __fsub_rn(a.z, b.z)); }
__device__ __forceinline__ H3 hdiv(H3 a, float s) { return h3(__fdiv_rn(a.x, s), __fdiv_rn(a.y, s), __fdiv_rn(a.z, s)); }
__device__ __forceinline__ float hdot(H3 a, H3 b) {          // cutil_math.h:1119-1122, left to right
  return __fadd_rn(__fadd_rn(__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y)), __fmul_rn(a.z, b.z));
}
__device__ __forceinline__ H3 hcross(H3 a, H3 b) {           // cutil_math.h:1295-1298
  return h3(__fsub_rn(__fmul_rn(a.y, b.z), __fmul_rn(a.z, b.y)), __fsub_rn(__fmul_rn(a.z, b.x), __fmul_rn(a.x, b.z)),
            __fsub_rn(__fmul_rn(a.x, b.y), __fmul_rn(a.y, b.x)));
}

// rayTriangleIntersectHost (CUDABasedRasterization.cpp:156-235) for the fixed ray of a texel
__device__ __forceinline__ bool texel_hit(float px, float py, H3 v0, H3 v1, H3 v2, float& a, float& b) {
  v0 = hdiv(v0, 1000.f); v1 = hdiv(v1, 1000.f); v2 = hdiv(v2, 1000.f);
  const H3 orig = hdiv(h3(px, py, 1.f), 1000.f);
  const H3 dir = h3(0.f, 0.f, -1.f);
  const H3 N = hcross(hsub(v1, v0), hsub(v2, v0));
  const float nd = hdot(dir, N);
  if (fabsf(nd) < 0.0000001f) return false;
  const float t = __fdiv_rn(__fsub_rn(hdot(v0, N), hdot(orig, N)), nd);
  if (t < 0.f) return false;
  const H3 P = h3(__fadd_rn(orig.x, __fmul_rn(t, dir.x)), __fadd_rn(orig.y, __fmul_rn(t, dir.y)), __fadd_rn(orig.z, __fmul_rn(t, dir.z)));
  if (hdot(N, hcross(hsub(v1, v0), hsub(P, v0))) < 0.f) return false;
  a = hdot(N, hcross(hsub(v2, v1), hsub(P, v1)));
  if (a < 0.f) return false;
  b = hdot(N, hcross(hsub(v0, v2), hsub(P, v2)));
  if (b < 0.f) return false;
  const float den = hdot(N, N);
  a = __fdiv_rn(a, den); b = __fdiv_rn(b, den);
  return true;
}

__device__ __forceinline__ void texel_tri(const float* __restrict__ tc, int f, int texW, int texH, H3& t0, H3& t1, H3& t2) {
  const float* t = tc + (size_t)f * 6;
  const float w = (float)texW, h = (float)texH;
  t0 = h3(__fmul_rn(w, t[0]), __fmul_rn(h, __fsub_rn(1.f, t[1])), 0.f);
  t1 = h3(__fmul_rn(w, t[2]), __fmul_rn(h, __fsub_rn(1.f, t[3])), 0.f);
  t2 = h3(__fmul_rn(w, t[4]), __fmul_rn(h, __fsub_rn(1.f, t[5])), 0.f);
}

__global__ void texel_clear_kernel(int* __restrict__ faceTable, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) faceTable[i] = -1;
}

__global__ void texel_face_kernel(const float* __restrict__ tc, int F, int texH, int texW, int* __restrict__ faceTable) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= F) return;
  H3 t0, t1, t2;
  texel_tri(tc, f, texW, texH, t0, t1, t2);
  const int xMin = (int)fmaxf(fminf(t0.x, fminf(t1.x, t2.x)) - 2.f, 0.f);
  const int xMax = (int)fminf(fmaxf(t0.x, fmaxf(t1.x, t2.x)) + 2.f, (float)texW);
  const int yMin = (int)fmaxf(fminf(t0.y, fminf(t1.y, t2.y)) - 2.f, 0.f);
  const int yMax = (int)fminf(fmaxf(t0.y, fmaxf(t1.y, t2.y)) + 2.f, (float)texH);
  for (int x = xMin; x < xMax; ++x)
    for (int y = yMin; y < yMax; ++y) {
      float a, b;
      if (texel_hit(x + 0.5f, y + 0.5f, t0, t1, t2, a, b)) atomicMax(faceTable + (size_t)y * texW + x, f);
    }
}

__global__ void texel_bary_kernel(const float* __restrict__ tc, int texH, int texW, const int* __restrict__ faceTable,
                                  float4* __restrict__ table) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= texH * texW) return;
  const int f = faceTable[i];
  float4 out = make_float4(0.f, 0.f, 0.f, 0.f);   // the reference's initial value (:250)
  if (f >= 0) {
    H3 t0, t1, t2;
    texel_tri(tc, f, texW, texH, t0, t1, t2);
    float a = 0.f, b = 0.f;
    texel_hit((i % texW) + 0.5f, (i / texW) + 0.5f, t0, t1, t2, a, b);
    out = make_float4((float)f, a, b, __fsub_rn(__fsub_rn(1.f, a), b));
  }
  table[i] = out;
}

// renderNormalMapDevice (CUDABasedRasterization.cu:415-445); vnorm4 = per-batch vertex normals
__global__ void normal_map_kernel(const float4* __restrict__ table, const int4* __restrict__ faces4, const float4* __restrict__ vnorm4,
                                  float* __restrict__ normal_map, int texels, int N) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (i >= texels) return;
  const float4 info = __ldg(table + i);
  const int4 fc = __ldg(faces4 + (int)info.x);
  const float4* vn = vnorm4 + (size_t)b * N;
  const float4 n0 = __ldg(vn + fc.x), n1 = __ldg(vn + fc.y), n2 = __ldg(vn + fc.z);
  // contraction as in the compiled reference kernel (SASS of renderNormalMapDevice, nvcc 12.9 sm_100a):
  // fma(c, n2, fma(a, n0, b * n1)), dot = fma(z,z, fma(x,x, y*y)), IEEE sqrt and divides
  float nx = interp3(info.y, info.z, info.w, n0.x, n1.x, n2.x);
  float ny = interp3(info.y, info.z, info.w, n0.y, n1.y, n2.y);
  float nz = interp3(info.y, info.z, info.w, n0.z, n1.z, n2.z);
  const float len = __fsqrt_rn(dot3x(mk3(nx, ny, nz), mk3(nx, ny, nz)));
  if (len != 0.f) { nx = __fdiv_rn(nx, len); ny = __fdiv_rn(ny, len); nz = __fdiv_rn(nz, len); }
  float* o = normal_map + ((size_t)b * texels + i) * 3;
  o[0] = __fmul_rn(__fadd_rn(nx, 1.f), 0.5f); o[1] = __fmul_rn(__fadd_rn(ny, 1.f), 0.5f); o[2] = __fmul_rn(__fadd_rn(nz, 1.f), 0.5f);
}

int launch_build_texel_table(const float* texcoords, int F, int texH, int texW, float4* table, cudaStream_t st) {
  int* faceTable = nullptr;
  const int n = texH * texW;
  if (cudaMalloc((void**)&faceTable, sizeof(int) * (size_t)n) != cudaSuccess) return -1;
  texel_clear_kernel<<<(n + 255) / 256, 256, 0, st>>>(faceTable, n);
  if (F > 0) texel_face_kernel<<<(F + 127) / 128, 128, 0, st>>>(texcoords, F, texH, texW, faceTable);
  texel_bary_kernel<<<(n + 255) / 256, 256, 0, st>>>(texcoords, texH, texW, faceTable, table);
  const bool ok = cudaGetLastError() == cudaSuccess && cudaStreamSynchronize(st) == cudaSuccess;
  cudaFree(faceTable);
  return ok ? 3 : -1;
}

int launch_camera(const float* extr, const float* intr, CamRec* cams, int* bigCount, int V, cudaStream_t st);
int launch_vertex(const FwdArgs& a, cudaStream_t st);

int launch_normal_map(const FwdArgs& a, const float4* texelTable, float* normal_map, cudaStream_t st) {
  // the reference still runs the camera / projection / normal kernels in this mode (:451-461)
  int launches = launch_camera(a.extrinsics, a.intrinsics, a.s.cams, a.s.bigCount, a.B * a.C, st);
  launches += launch_vertex(a, st);
  const int texels = a.texH * a.texW;
  normal_map_kernel<<<dim3((texels + 255) / 256, a.B), 256, 0, st>>>(texelTable, a.faces4, a.s.vnorm4, normal_map, texels, a.N);
  ++launches;
  return cudaGetLastError() == cudaSuccess ? launches : -1;
}

}  // namespace gvv
