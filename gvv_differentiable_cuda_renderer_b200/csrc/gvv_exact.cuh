// gvv_exact.cuh -- the bit-exact fp32 arithmetic of the visibility path.
//
// Visibility in the reference is decided by (int)(z*10000) (CUDABasedRasterization.cu:247-250).
// Neighbouring triangles meet at shared edges where their depths differ only by rounding noise,
// so the face buffer is bit-exact against the reference only if every rounding on the path
//   camera inverse -> ray -> ray/triangle intersection -> barycentrics -> depth key
// is the same.  The reference is compiled with nvcc's default -fmad=true; which a*b+c become
// FMAs was read off its PTX/SASS (nvcc 12.9, sm_100a) and is pinned here with explicit
// round-to-nearest intrinsics, which the compiler may neither fuse nor re-associate:
//
//   dot(a,b)   = fma(a.z,b.z, fma(a.x,b.x, a.y*b.y))               cutil_math.h:1123-1126
//   cross(a,b) = ( fma(a.y,b.z, -(a.z*b.y)), fma(a.z,b.x, -(a.x*b.z)), fma(a.x,b.y, -(a.y*b.x)) )
//                                                                   cutil_math.h:1295-1298
//   x / y      = IEEE div.rn ; 1.f / x = rcp.rn ; normalize = v * rsqrt.approx(dot(v,v))
//
// Tiling changes WHO evaluates a (pixel, triangle) pair, never WHAT is evaluated.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace gvv {

struct F3 { float x, y, z; };

__device__ __forceinline__ F3 mk3(float x, float y, float z) { F3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ F3 sub3(F3 a, F3 b) { return mk3(__fsub_rn(a.x, b.x), __fsub_rn(a.y, b.y), __fsub_rn(a.z, b.z)); }

__device__ __forceinline__ float dot3x(F3 a, F3 b) {
  return __fmaf_rn(a.z, b.z, __fmaf_rn(a.x, b.x, __fmul_rn(a.y, b.y)));
}
__device__ __forceinline__ F3 cross3x(F3 a, F3 b) {
  return mk3(__fmaf_rn(a.y, b.z, -__fmul_rn(a.z, b.y)),
             __fmaf_rn(a.z, b.x, -__fmul_rn(a.x, b.z)),
             __fmaf_rn(a.x, b.y, -__fmul_rn(a.y, b.x)));
}

// Per-view camera record (64 floats).  Filled by camera_kernel.
struct CamRec {
  float K[9];      // intrinsics, row-major 3x3
  float E[12];     // extrinsics, row-major 3x4
  float Einv[16];  // inverse(E4)
  float Pinv[16];  // inverse(K4*E4)
  float ro[3];     // ray origin  = Einv[:,3] / Einv[3,3]      (CameraUtil.h:253-255)
  float ros[3];    // ro / 1000                                 (RendererUtil.h:32)
  float pad[5];
};
static_assert(sizeof(CamRec) == 64 * 4, "CamRec must be 64 floats");

// Reference: initializeCamerasDevice (CUDABasedRasterization.cu:23-67) -> float4x4::getInverse
// (float4x4.h:160-285).  The statement list is generated from the compiled reference, see
// tools/sass2intrinsics.py.
__device__ __forceinline__ void camera_inverse_exact(const float* __restrict__ K, const float* __restrict__ E,
                                                     float* __restrict__ Einv, float* __restrict__ Pinv) {
#include "cam_inverse_exact.inc"
}

// One view's camera record: initializeCamerasDevice (CUDABasedRasterization.cu:23-67) + the ray origin.
__device__ __forceinline__ void fill_camrec(const float* __restrict__ extr, const float* __restrict__ intr, CamRec* __restrict__ cams, int v) {
  float K[9], E[12], Einv[16], Pinv[16];
#pragma unroll
  for (int i = 0; i < 9; ++i) K[i] = intr[v * 9 + i];
#pragma unroll
  for (int i = 0; i < 12; ++i) E[i] = extr[v * 12 + i];
  camera_inverse_exact(K, E, Einv, Pinv);
  CamRec& r = cams[v];
#pragma unroll
  for (int i = 0; i < 9; ++i) r.K[i] = K[i];
#pragma unroll
  for (int i = 0; i < 12; ++i) r.E[i] = E[i];
#pragma unroll
  for (int i = 0; i < 16; ++i) { r.Einv[i] = Einv[i]; r.Pinv[i] = Pinv[i]; }
  // o = 4th column of E^-1, o /= o.w (CameraUtil.h:253-255); ros = o / 1000 (RendererUtil.h:32)
  const float ox = __fdiv_rn(Einv[3], Einv[15]);
  const float oy = __fdiv_rn(Einv[7], Einv[15]);
  const float oz = __fdiv_rn(Einv[11], Einv[15]);
  r.ro[0] = ox; r.ro[1] = oy; r.ro[2] = oz;
  r.ros[0] = __fdiv_rn(ox, 1000.f); r.ros[1] = __fdiv_rn(oy, 1000.f); r.ros[2] = __fdiv_rn(oz, 1000.f);
#pragma unroll
  for (int i = 0; i < 5; ++i) r.pad[i] = 0.f;   // the record is copied whole into shared memory by its readers
}

// Reference: getRayCuda2 + backprojectPixelCuda (CameraUtil.h:223-236,251-258).
// px,py = pixel centre (u+0.5, v+0.5).  Returns the normalised world-space direction.
__device__ __forceinline__ F3 ray_dir_exact(const float* __restrict__ Pinv, const float* __restrict__ ro, float px, float py) {
  const float tx = __fmul_rn(px, 1000.f);
  const float ty = __fmul_rn(py, 1000.f);
  // dot(row, (tx,ty,1000,1)) = row.w + fma(row.z,1000, fma(tx,row.x, ty*row.y))
  const float wx = __fadd_rn(Pinv[3],  __fmaf_rn(Pinv[2],  1000.f, __fmaf_rn(tx, Pinv[0], __fmul_rn(ty, Pinv[1]))));
  const float wy = __fadd_rn(Pinv[7],  __fmaf_rn(Pinv[6],  1000.f, __fmaf_rn(tx, Pinv[4], __fmul_rn(ty, Pinv[5]))));
  const float wz = __fadd_rn(Pinv[11], __fmaf_rn(Pinv[10], 1000.f, __fmaf_rn(tx, Pinv[8], __fmul_rn(ty, Pinv[9]))));
  const float dx = __fsub_rn(wx, ro[0]);
  const float dy = __fsub_rn(wy, ro[1]);
  const float dz = __fsub_rn(wz, ro[2]);
  const float l2 = __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
  const float inv = rsqrtf(l2);                       // rsqrt.approx.f32, as cutil_math.h:1186
  return mk3(__fmul_rn(dx, inv), __fmul_rn(dy, inv), __fmul_rn(inv, dz));
}

// Triangle setup: everything of rayTriangleIntersect (RendererUtil.h:26-101) that does not depend
// on the pixel.  v0s,v1s,v2s are the vertices already divided by 1000 (div.rn), ros = ro/1000.
struct TriSetup {
  F3 v0, v1, v2;   // scaled vertices
  F3 N;            // cross(v1-v0, v2-v0)
  float num;       // dot(v0,N) - dot(ros,N)
  float den;       // dot(N,N)
};

__device__ __forceinline__ TriSetup tri_setup_exact(F3 v0s, F3 v1s, F3 v2s, F3 ros) {
  TriSetup t;
  t.v0 = v0s; t.v1 = v1s; t.v2 = v2s;
  const F3 e01 = sub3(v1s, v0s);
  const F3 e02 = sub3(v2s, v0s);
  t.N = cross3x(e01, e02);
  t.num = __fsub_rn(dot3x(v0s, t.N), dot3x(ros, t.N));
  t.den = dot3x(t.N, t.N);
  return t;
}

// Pixel part of rayTriangleIntersect + uv2barycentric (RendererUtil.h:46-127) + the inside test of
// CUDABasedRasterization.cu:243.  Returns true iff the reference would call atomicMin for this pair;
// a,b,c are then the reference's barycentrics (bit-exact).
// skip_rest: the caller already knows this pair cannot win the depth test (conservative early-z, see
// raster_kernel); the remaining divides are then not spent.  It never changes a result that is used.
__device__ __forceinline__ bool hit_exact(const TriSetup& t, F3 ros, F3 rd, float& a, float& b, float& c, bool skip_rest = false) {
  const float nd = dot3x(rd, t.N);
  if (fabsf(nd) < 0.0000001f) return false;
  const float tt = __fdiv_rn(t.num, nd);
  if (tt < 0.f) return false;
  const F3 P = mk3(__fmaf_rn(rd.x, tt, ros.x), __fmaf_rn(rd.y, tt, ros.y), __fmaf_rn(rd.z, tt, ros.z));
  // the three edge functions are evaluated together (independent chains, more ILP); the reference
  // rejects on them one after the other -- same values, same decision
  const float e0n = dot3x(t.N, cross3x(sub3(t.v1, t.v0), sub3(P, t.v0)));
  const float an = dot3x(t.N, cross3x(sub3(t.v2, t.v1), sub3(P, t.v1)));
  const float bn = dot3x(t.N, cross3x(sub3(t.v0, t.v2), sub3(P, t.v2)));
  if (e0n < 0.f || an < 0.f || bn < 0.f) return false;
  if (skip_rest) return false;
  a = __fdiv_rn(an, t.den);
  b = __fdiv_rn(bn, t.den);
  c = __fsub_rn(__fsub_rn(1.f, a), b);
  // !(x >= -0.001f) also rejects NaN, like the reference's comparison chain
  return (a >= -0.001f) && (b >= -0.001f) && (c >= -0.001f) && (a <= 1.001f) && (b <= 1.001f) && (c <= 1.001f);
}

// Depth key: CUDABasedRasterization.cu:247-250.  z0,z1,z2 = projected depths of the three vertices.
__device__ __forceinline__ int depth_key_exact(float a, float b, float c, float z0, float z1, float z2) {
  const float s = __fadd_rn(__fdiv_rn(c, z2), __fadd_rn(__fdiv_rn(a, z0), __fdiv_rn(b, z1)));
  const float z = __fmul_rn(__frcp_rn(s), 10000.f);
  return __float2int_rz(z);          // cvt.rzi.s32.f32 (saturating, NaN -> 0)
}

// 64-bit z-buffer word: depth in the high half (sign bit flipped so unsigned order == the
// reference's signed atomicMin order), triangle id in the low half => ties resolve to the
// smallest triangle id, deterministically.
__device__ __forceinline__ unsigned long long pack_key(int depth, int face) {
  return ((unsigned long long)((unsigned)depth ^ 0x80000000u) << 32) | (unsigned)face;
}
constexpr unsigned long long kEmptyKey = 0xFFFFFFFFFFFFFFFFull;

// Vertex projection: getCamSpacePoint + projectPointFloat3 (CameraUtil.h:174-186,141-170).
__device__ __forceinline__ float4 project_exact(const float* __restrict__ K, const float* __restrict__ E, float vx, float vy, float vz) {
  const float cx = __fadd_rn(E[3],  __fmaf_rn(vz, E[2],  __fmaf_rn(vx, E[0], __fmul_rn(vy, E[1]))));
  const float cy = __fadd_rn(E[7],  __fmaf_rn(vz, E[6],  __fmaf_rn(vx, E[4], __fmul_rn(vy, E[5]))));
  const float cz = __fadd_rn(E[11], __fmaf_rn(vz, E[10], __fmaf_rn(vx, E[8], __fmul_rn(vy, E[9]))));
  const float x = __fmaf_rn(cz, K[2], __fmaf_rn(cx, K[0], __fmul_rn(cy, K[1])));
  const float y = __fmaf_rn(cz, K[5], __fmaf_rn(cx, K[3], __fmul_rn(cy, K[4])));
  float z       = __fmaf_rn(cz, K[8], __fmaf_rn(cx, K[6], __fmul_rn(cy, K[7])));
  z = (z > 0.0000001f) ? z : 0.00001f;
  return make_float4(__fdiv_rn(x, z), __fdiv_rn(y, z), z, 0.f);
}

// Screen bbox of a triangle: projectFacesDevice (CUDABasedRasterization.cu:202-206); inclusive.
__device__ __forceinline__ int4 bbox_exact(float4 p0, float4 p1, float4 p2, int W, int H) {
  int4 bb;
  bb.x = __float2int_rz(fmaxf(__fadd_rn(fminf(p0.x, fminf(p1.x, p2.x)), -0.5f), 0.f));
  bb.y = __float2int_rz(fmaxf(__fadd_rn(fminf(p0.y, fminf(p1.y, p2.y)), -0.5f), 0.f));
  bb.z = __float2int_rz(fminf(__fadd_rn(fmaxf(p0.x, fmaxf(p1.x, p2.x)), 0.5f), (float)(W - 1)));
  bb.w = __float2int_rz(fminf(__fadd_rn(fmaxf(p0.y, fmaxf(p1.y, p2.y)), 0.5f), (float)(H - 1)));
  return bb;
}

// a*q0 + b*q1 + c*q2 as the reference's pass 2 contracts it: fma(c,q2, fma(a,q0, b*q1)).
__device__ __forceinline__ float interp3(float a, float b, float c, float q0, float q1, float q2) {
  return __fmaf_rn(c, q2, __fmaf_rn(a, q0, __fmul_rn(b, q1)));
}

}  // namespace gvv
