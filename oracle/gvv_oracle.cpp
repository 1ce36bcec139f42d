// gvv_oracle.cpp -- TEST INFRASTRUCTURE ONLY (Oracle 2).
//
// Multi-threaded C++17/OpenMP CPU restatement of the reference's differentiable rasteriser,
// written from the reference sources as the *specification*; it is the checker for
// tests/, __graft_entry__.smoke() and the `cpu_baseline` leg of bench.py and is never linked
// into, imported by or called from the product (libgvv_b200.so / the Python package).
//
// Pinning: the reference ships no golden vectors.  This oracle is pinned against
// outputs of the reference's own CUDA core (Oracle 1, oracle/_ref) captured on a B200 and
// committed under tests/golden/ (see tests/golden/README.md and tools/make_golden.py).
//
// It cannot be bit-identical to a GPU run: device rsqrtf is an approximation and nvcc contracts
// a*b+c into FMAs.  Visibility is therefore resolved deterministically (min over a packed
// (depth, face id) key) and every pixel also reports the runner-up depth and a tie flag so that a
// test can tell a genuine mismatch from a rounding-level near-tie.
//
// Each function cites the reference lines it restates (paths relative to the reference checkout).
#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

// GVVO_FP64 build (oracle/libgvv_oracle64.so): the SAME source with every `float` -- arithmetic AND the I/O arrays --
// turned into `double` after the standard headers have been parsed.  It evaluates the reference's formulas at the
// same inputs without fp32 rounding and is what the gradient protocol uses to tell a genuine difference from the
// reference's own fp32 noise (the position gradient's dJBCDVerpos chain, RendererUtil.h:670-861, subtracts terms
// ~1e5 times larger than their difference when triangles are millimetres wide and the camera metres away).
#ifdef GVVO_FP64
#define float double
#endif

namespace {

struct V3 { float x, y, z; };
inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator*(V3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline V3 operator*(float s, V3 a) { return {a.x * s, a.y * s, a.z * s}; }
inline V3 operator/(V3 a, float s) { return {a.x / s, a.y / s, a.z / s}; }
inline float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }                       // cutil_math.h:1123
inline V3 cross(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }  // :1295
inline V3 normalize(V3 v) { return v * (1.0f / std::sqrt(dot(v, v))); }                         // :1184 (rsqrtf on device)
inline V3 ld3(const float* p, long i) { return {p[3 * i], p[3 * i + 1], p[3 * i + 2]}; }

enum Albedo { VertexColor = 0, Textured = 1, Normal = 2, Lighting = 3, ForegroundMask = 4 };   // CUDABasedRasterizationInput.h:25-28
enum Shading { Shaded = 0, Shadeless = 1 };                                                      // :32-35

struct Camera {
  float K[9], E[12];
  float Einv[16], Pinv[16];   // inverse(E4), inverse(K4*E4)
  V3 o;                       // ray origin
};

// float4x4::getInverse (cpp/src/Utils/float4x4.h:160-285): adjugate / determinant.
void inverse4(const float* m, float* out) {
  auto minor3 = [&](int r, int c) {
    int rr[3], cc[3], k = 0;
    for (int i = 0; i < 4; ++i) if (i != r) rr[k++] = i;
    k = 0;
    for (int i = 0; i < 4; ++i) if (i != c) cc[k++] = i;
    auto a = [&](int i, int j) { return m[4 * rr[i] + cc[j]]; };
    return a(0, 0) * (a(1, 1) * a(2, 2) - a(1, 2) * a(2, 1)) - a(0, 1) * (a(1, 0) * a(2, 2) - a(1, 2) * a(2, 0)) +
           a(0, 2) * (a(1, 0) * a(2, 1) - a(1, 1) * a(2, 0));
  };
  float cof[16];
  for (int r = 0; r < 4; ++r)
    for (int c = 0; c < 4; ++c) cof[4 * r + c] = (((r + c) & 1) ? -1.f : 1.f) * minor3(r, c);
  const float det = m[0] * cof[0] + m[1] * cof[1] + m[2] * cof[2] + m[3] * cof[3];
  const float idet = 1.0f / det;
  for (int r = 0; r < 4; ++r)
    for (int c = 0; c < 4; ++c) out[4 * r + c] = cof[4 * c + r] * idet;
}

// initializeCamerasDevice (cpp/src/Renderer/CUDABasedRasterization.cu:23-67)
Camera make_camera(const float* K, const float* E) {
  Camera c;
  std::memcpy(c.K, K, sizeof(c.K));
  std::memcpy(c.E, E, sizeof(c.E));
  float K4[16] = {K[0], K[1], K[2], 0, K[3], K[4], K[5], 0, K[6], K[7], K[8], 0, 0, 0, 0, 1};
  float E4[16] = {E[0], E[1], E[2], E[3], E[4], E[5], E[6], E[7], E[8], E[9], E[10], E[11], 0, 0, 0, 1};
  float KE[16];
  for (int r = 0; r < 4; ++r)
    for (int col = 0; col < 4; ++col) {
      float s = 0.f;
      for (int k = 0; k < 4; ++k) s += K4[4 * r + k] * E4[4 * k + col];   // float4x4.h:78-101
      KE[4 * r + col] = s;
    }
  inverse4(E4, c.Einv);
  inverse4(KE, c.Pinv);
  const float w = c.Einv[15];
  c.o = {c.Einv[3] / w, c.Einv[7] / w, c.Einv[11] / w};                    // CameraUtil.h:253-255
  return c;
}

// getRayCuda2 + backprojectPixelCuda (cpp/src/Utils/CameraUtil.h:223-236,251-258)
V3 ray_dir(const Camera& c, float px, float py) {
  const float t[4] = {px * 1000.f, py * 1000.f, 1000.f, 1.f};
  V3 w;
  w.x = c.Pinv[0] * t[0] + c.Pinv[1] * t[1] + c.Pinv[2] * t[2] + c.Pinv[3] * t[3];
  w.y = c.Pinv[4] * t[0] + c.Pinv[5] * t[1] + c.Pinv[6] * t[2] + c.Pinv[7] * t[3];
  w.z = c.Pinv[8] * t[0] + c.Pinv[9] * t[1] + c.Pinv[10] * t[2] + c.Pinv[11] * t[3];
  return normalize(w - c.o);
}

// getCamSpacePoint + projectPointFloat3 (CameraUtil.h:174-186,141-170)
V3 project(const Camera& c, V3 v) {
  const float cx = c.E[0] * v.x + c.E[1] * v.y + c.E[2] * v.z + c.E[3];
  const float cy = c.E[4] * v.x + c.E[5] * v.y + c.E[6] * v.z + c.E[7];
  const float cz = c.E[8] * v.x + c.E[9] * v.y + c.E[10] * v.z + c.E[11];
  float x = cx * c.K[0] + cy * c.K[1] + cz * c.K[2];
  float y = cx * c.K[3] + cy * c.K[4] + cz * c.K[5];
  float z = cx * c.K[6] + cy * c.K[7] + cz * c.K[8];
  if (!(z > 0.0000001f)) z = 0.00001f;
  return {x / z, y / z, z};
}

// rayTriangleIntersect + uv2barycentric (cpp/src/Utils/RendererUtil.h:26-128)
V3 uv2barycentric(const Camera& c, float px, float py, V3 v0, V3 v1, V3 v2) {
  const V3 miss = {-1.f, -1.f, -1.f};
  const V3 dir = ray_dir(c, px, py);
  v0 = v0 / 1000.f; v1 = v1 / 1000.f; v2 = v2 / 1000.f;
  const V3 orig = c.o / 1000.f;
  const V3 N = cross(v1 - v0, v2 - v0);
  const float nd = dot(dir, N);
  if (std::fabs(nd) < 0.0000001f) return miss;
  const float t = (dot(v0, N) - dot(orig, N)) / nd;
  if (t < 0) return miss;
  const V3 P = orig + t * dir;
  if (dot(N, cross(v1 - v0, P - v0)) < 0) return miss;
  float a = dot(N, cross(v2 - v1, P - v1));
  if (a < 0) return miss;
  float b = dot(N, cross(v0 - v2, P - v2));
  if (b < 0) return miss;
  const float den = dot(N, N);
  a /= den; b /= den;
  return {a, b, 1.f - a - b};
}

inline bool inside(V3 abc) {   // CUDABasedRasterization.cu:243
  return (abc.x >= -0.001f) && (abc.y >= -0.001f) && (abc.z >= -0.001f) && (abc.x <= 1.001f) && (abc.y <= 1.001f) && (abc.z <= 1.001f);
}

inline int f2i_rz(float z) {   // cvt.rzi.s32.f32: truncation, saturating, NaN -> 0
  if (std::isnan(z)) return 0;
  if (z >= 2147483648.f) return INT_MAX;
  if (z <= -2147483648.f) return INT_MIN;
  return (int)z;
}

// getShading / getIllum (RendererUtil.h:135-214)
V3 illum(V3 n, const float* sh) {
  float L[3];
  for (int ch = 0; ch < 3; ++ch) {
    const float* s = sh + 9 * ch;
    float v = s[0];
    v += s[1] * n.y; v += s[2] * n.z; v += s[3] * n.x; v += s[4] * (n.x * n.y); v += s[5] * (n.z * n.y);
    v += s[6] * (3.f * n.z * n.z - 1.f); v += s[7] * (n.x * n.z); v += s[8] * (n.x * n.x - n.y * n.y);
    L[ch] = v;
  }
  return {L[0], L[1], L[2]};
}

struct Mesh {
  int N, F;
  const int* faces;
  const float* tc;
  std::vector<int> vfOff, vfList;   // getVertexFaces (CUDABasedRasterization.cpp:125-154), ascending face order
  Mesh(const int* f, int F_, const float* t, int N_) : N(N_), F(F_), faces(f), tc(t), vfOff(N_ + 1, 0) {
    auto distinct = [&](int fi, int k) { for (int j = 0; j < k; ++j) if (f[3 * fi + j] == f[3 * fi + k]) return false; return true; };
    for (int i = 0; i < F; ++i) for (int k = 0; k < 3; ++k) if (distinct(i, k)) vfOff[f[3 * i + k] + 1]++;
    for (int n = 0; n < N; ++n) vfOff[n + 1] += vfOff[n];
    vfList.resize(vfOff[N]);
    std::vector<int> cur(vfOff.begin(), vfOff.end() - 1);
    for (int i = 0; i < F; ++i) for (int k = 0; k < 3; ++k) if (distinct(i, k)) vfList[cur[f[3 * i + k]]++] = i;
  }
};

// Non-default variant (gvvo_set_texture_bilinear): the bilinear texture fetch (CUDABasedRasterization.cu:365-372)
// and the four weighted gradient adds (CUDABasedRasterizationGrad.cu:361-378) that the reference has commented out.
static int g_texBilinear = 0;

struct TexSample { float u, v; int lu, lv, hu, hv; float LU, LV, HU, HV; };
// texture coordinate -> texel (CUDABasedRasterization.cu:326-345, CUDABasedRasterizationGrad.cu:250-289)
TexSample tex_coord(const float* tc, int face, V3 abc, int texW, int texH) {
  const float* t = tc + 6 * (long)face;
  float u = t[0] * abc.x + t[2] * abc.y + t[4] * abc.z;
  float v = (1.f - t[1]) * abc.x + (1.f - t[3]) * abc.y + (1.f - t[5]) * abc.z;
  u *= texW; v *= texH;
  u = std::fmin(std::fmax(u, 0.f), (float)(texW - 1));
  v = std::fmin(std::fmax(v, 0.f), (float)(texH - 1));
  TexSample s;
  s.u = u; s.v = v;
  s.LU = (float)(int)(u - 0.5f) + 0.5f; s.HU = (float)(int)(u - 0.5f) + 1.5f;
  s.LV = (float)(int)(v - 0.5f) + 0.5f; s.HV = (float)(int)(v - 0.5f) + 1.5f;
  s.lu = (int)s.LU; s.hu = std::min((int)s.HU, texW - 1); s.lv = (int)s.LV; s.hv = std::min((int)s.HV, texH - 1);   // clamps only act on 1-texel-wide textures
  return s;
}

}  // namespace

extern "C" {

void gvvo_set_texture_bilinear(int on) { g_texBilinear = on ? 1 : 0; }

// Forward of the op for all batch elements (CudaRenderer.cpp:298-335 -> renderBuffersGPU,
// CUDABasedRasterization.cu:449-473).  Extra outputs (may be null): best_depth / second_depth
// int32 [B,C,H,W] (INT_MAX where absent) and tie uint8 [B,C,H,W] (1 if >= 2 triangles share the
// winning depth).  Returns the number of (triangle, pixel) pairs that passed the inside test.
long long gvvo_forward(const int* faces, int F, const float* texcoords, int N, int C, int W, int H, int albedo, int shading,
                       int B, int texH, int texW, const float* vertex_pos, const float* vertex_color, const float* texture,
                       const float* sh_coeff, const float* extrinsics, const float* intrinsics,
                       float* bary, int* face_buf, float* render, float* vertex_normal,
                       int* best_depth, int* second_depth, unsigned char* tie, int nthreads) {
  if (albedo == ForegroundMask) shading = Shadeless;   // CudaRenderer.cpp:72-76
  Mesh mesh(faces, F, texcoords, N);
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
  long long fragments = 0;
  const long P = (long)W * H;
  for (int b = 0; b < B; ++b) {
    const float* pos = vertex_pos + (long)b * N * 3;
    // renderFaceNormalDevice / renderVertexNormalDevice (:122-174): unnormalised, ascending faces
    std::vector<V3> fn(F), vn(N);
#pragma omp parallel for schedule(static)
    for (int f = 0; f < F; ++f) {
      const V3 v0 = ld3(pos, faces[3 * f]), v1 = ld3(pos, faces[3 * f + 1]), v2 = ld3(pos, faces[3 * f + 2]);
      fn[f] = cross(v1 - v0, v2 - v0);
    }
#pragma omp parallel for schedule(static)
    for (int n = 0; n < N; ++n) {
      V3 s = {0.f, 0.f, 0.f};   // the reference leaves isolated vertices uninitialised; defined as 0 here
      for (int i = mesh.vfOff[n]; i < mesh.vfOff[n + 1]; ++i) s = (i == mesh.vfOff[n]) ? fn[mesh.vfList[i]] : s + fn[mesh.vfList[i]];
      vn[n] = s;
      for (int c = 0; c < C; ++c) {
        float* o = vertex_normal + (((long)b * C + c) * N + n) * 3;
        o[0] = s.x; o[1] = s.y; o[2] = s.z;
      }
    }
    for (int c = 0; c < C; ++c) {
      const long view = (long)b * C + c;
      const Camera cam = make_camera(intrinsics + view * 9, extrinsics + view * 12);
      std::vector<V3> pv(N);
#pragma omp parallel for schedule(static)
      for (int n = 0; n < N; ++n) pv[n] = project(cam, ld3(pos, n));                       // projectVerticesDevice :98-115
      std::vector<int> bb(4 * (size_t)F);
#pragma omp parallel for schedule(static)
      for (int f = 0; f < F; ++f) {                                                         // projectFacesDevice :184-208
        const V3 a = pv[faces[3 * f]], bq = pv[faces[3 * f + 1]], cq = pv[faces[3 * f + 2]];
        bb[4 * f + 0] = f2i_rz(std::fmax(std::fmin(a.x, std::fmin(bq.x, cq.x)) - 0.5f, 0.f));
        bb[4 * f + 1] = f2i_rz(std::fmax(std::fmin(a.y, std::fmin(bq.y, cq.y)) - 0.5f, 0.f));
        bb[4 * f + 2] = f2i_rz(std::fmin(std::fmax(a.x, std::fmax(bq.x, cq.x)) + 0.5f, (float)(W - 1)));
        bb[4 * f + 3] = f2i_rz(std::fmin(std::fmax(a.y, std::fmax(bq.y, cq.y)) + 0.5f, (float)(H - 1)));
      }
      // depth pass + buffer pass (:215-408) over row bands; per pixel min of (depth, face id)
      const int bands = std::max(1, std::min(H, 64));
      long long frag = 0;
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : frag)
      for (int band = 0; band < bands; ++band) {
        const int y0 = (int)((long)H * band / bands), y1 = (int)((long)H * (band + 1) / bands) - 1;
        const int rows = y1 - y0 + 1;
        if (rows <= 0) continue;
        std::vector<int> bestD((size_t)rows * W, INT_MAX), secondD((size_t)rows * W, INT_MAX), bestF((size_t)rows * W, -1);
        std::vector<unsigned char> tieF((size_t)rows * W, 0), has((size_t)rows * W, 0);
        std::vector<V3> bestABC((size_t)rows * W, V3{0.f, 0.f, 0.f});
        for (int f = 0; f < F; ++f) {
          const int* q = &bb[4 * f];
          const int ya = std::max(q[1], y0), yb = std::min(q[3], y1);
          if (ya > yb || q[0] > q[2]) continue;
          const int i0 = faces[3 * f], i1 = faces[3 * f + 1], i2 = faces[3 * f + 2];
          const V3 v0 = ld3(pos, i0), v1 = ld3(pos, i1), v2 = ld3(pos, i2);
          for (int u = q[0]; u <= q[2]; ++u)
            for (int v = ya; v <= yb; ++v) {
              const V3 abc = uv2barycentric(cam, u + 0.5f, v + 0.5f, v0, v1, v2);
              if (!inside(abc)) continue;
              float z = 1.f / (abc.x / pv[i0].z + abc.y / pv[i1].z + abc.z / pv[i2].z);   // :247
              z *= 10000.f;
              const int d = f2i_rz(z);
              ++frag;
              const size_t k = (size_t)(v - y0) * W + u;
              if (!has[k] || d < bestD[k] || (d == bestD[k] && f < bestF[k])) {
                if (has[k]) { secondD[k] = bestD[k]; tieF[k] = (d == bestD[k]) ? 1 : 0; }
                bestD[k] = d; bestF[k] = f; bestABC[k] = abc; has[k] = 1;
              } else {
                if (d == bestD[k]) tieF[k] = 1;
                if (d < secondD[k]) secondD[k] = d;
              }
            }
        }
        for (int v = y0; v <= y1; ++v)
          for (int u = 0; u < W; ++u) {
            const size_t k = (size_t)(v - y0) * W + u;
            const long pix = view * P + (long)v * W + u;
            float r[3] = {0.f, 1.f, 0.f}, ab[2] = {0.f, 0.f};                              // initializeDevice :80-89
            const int f = bestF[k];
            if (f >= 0) {
              const V3 abc = bestABC[k];
              ab[0] = abc.x; ab[1] = abc.y;
              const int i0 = faces[3 * f], i1 = faces[3 * f + 1], i2 = faces[3 * f + 2];
              V3 n = vn[i0] * abc.x + vn[i1] * abc.y + vn[i2] * abc.z;                     // :308-319
              n = n / std::sqrt(dot(n, n));
              const V3 d = ray_dir(cam, u + 0.5f, v + 0.5f);
              if (dot(n, d) > 0.f) n = {-n.x, -n.y, -n.z};
              V3 col = {0.f, 0.f, 0.f};
              if (albedo == Textured) {                                                    // :324-374 (nearest texel)
                const TexSample s = tex_coord(texcoords, f, abc, texW, texH);
                const float* tb = texture + (long)b * texH * texW * 3;
                col = ld3(tb, (long)texW * s.lv + s.lu);
                if (g_texBilinear) {
                  const float wLULV = (s.v - s.LV) * (s.u - s.LU), wLUHV = (s.HV - s.v) * (s.u - s.LU);
                  const float wHULV = (s.v - s.LV) * (s.HU - s.u), wHUHV = (s.HV - s.v) * (s.HU - s.u);
                  col = wLULV * col + wHULV * ld3(tb, (long)texW * s.lv + s.hu) + wLUHV * ld3(tb, (long)texW * s.hv + s.lu) +
                        wHUHV * ld3(tb, (long)texW * s.hv + s.hu);
                }
              } else if (albedo == VertexColor) {                                          // :375-381
                const float* vc = vertex_color + (long)b * N * 3;
                col = ld3(vc, i0) * abc.x + ld3(vc, i1) * abc.y + ld3(vc, i2) * abc.z;
              } else if (albedo == Normal) {
                col = {(1.f + n.x) / 2.f, (1.f + n.y) / 2.f, (1.f + n.z) / 2.f};
              } else {
                col = {1.f, 1.f, 1.f};
              }
              if ((shading == Shaded && albedo != Normal) || albedo == Lighting) {          // :396-399
                const V3 L = illum(n, sh_coeff + view * 27);
                col = {col.x * L.x, col.y * L.y, col.z * L.z};
              }
              r[0] = col.x; r[1] = col.y; r[2] = col.z;
            }
            face_buf[pix] = f;
            bary[2 * pix] = ab[0]; bary[2 * pix + 1] = ab[1];
            render[3 * pix] = r[0]; render[3 * pix + 1] = r[1]; render[3 * pix + 2] = r[2];
            if (best_depth) best_depth[pix] = bestD[k];
            if (second_depth) second_depth[pix] = secondD[k];
            if (tie) tie[pix] = tieF[k];
          }
      }
      fragments += frag;
    }
  }
  return fragments;
}

// Backward of the op (CudaRendererGrad.cpp:252-292 -> renderBuffersGradDevice,
// CUDABasedRasterizationGrad.cu:114-617), literal per-pixel form including the loop over every
// face incident to the pixel's three vertices (:559-615).  Accumulation in double (order-free
// reference for the atomics-ordered GPU results).  target_grad may be null (= zeros).
int gvvo_backward(const int* faces, int F, const float* texcoords, int N, int C, int W, int H, int albedo, int shading,
                  int imageFilter, int B, int texH, int texW,
                  const float* render_grad, const float* target_grad, const float* vertex_pos, const float* vertex_color,
                  const float* texture, const float* sh_coeff, const float* target_image, const float* vertex_normal,
                  const float* bary, const int* face_buf, const float* extrinsics, const float* intrinsics,
                  float* vpos_grad, float* vcol_grad, float* tex_grad, float* sh_grad, int nthreads) {
  if (albedo == ForegroundMask) shading = Shadeless;   // CudaRendererGrad.cpp:78-82
  if (albedo == Normal || albedo == Lighting) return 1; // "Unsupported color mode" (:391-394)
  Mesh mesh(faces, F, texcoords, N);
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
  const int T = omp_get_max_threads();
#else
  const int T = 1;
#endif
  const long P = (long)W * H;
  const size_t texN = tex_grad ? (size_t)texH * texW * 3 : 0;
  for (int b = 0; b < B; ++b) {
    const float* pos = vertex_pos + (long)b * N * 3;
    const float* vc = vertex_color ? vertex_color + (long)b * N * 3 : nullptr;
    const float* tex = texture ? texture + (long)b * texH * texW * 3 : nullptr;
    std::vector<std::vector<double>> gp(T, std::vector<double>((size_t)N * 3, 0.0)), gc(T, std::vector<double>((size_t)N * 3, 0.0));
    std::vector<std::vector<double>> gt(T, std::vector<double>(albedo == Textured ? texN : 0, 0.0));
    std::vector<std::vector<double>> gs(T, std::vector<double>((size_t)C * 27, 0.0));
    for (int c = 0; c < C; ++c) {
      const long view = (long)b * C + c;
      const Camera cam = make_camera(intrinsics + view * 9, extrinsics + view * 12);
      const float* sh = sh_coeff + view * 27;
      const float* vnorm = vertex_normal + view * N * 3;
#pragma omp parallel for schedule(dynamic, 4)
      for (int idh = 0; idh < H; ++idh) {
#ifdef _OPENMP
        const int t = omp_get_thread_num();
#else
        const int t = 0;
#endif
        double* GP = gp[t].data(); double* GC = gc[t].data(); double* GS = gs[t].data() + c * 27;
        for (int idw = 0; idw < W; ++idw) {
          const long pix = view * P + (long)idh * W + idw;
          const int idf = face_buf[pix];
          if (idf == -1) continue;                                                           // :194-197
          const V3 o = cam.o, d = ray_dir(cam, idw + 0.5f, idh + 0.5f);                      // :205-208
          const V3 bcc = {bary[2 * pix], bary[2 * pix + 1], 1.f - bary[2 * pix] - bary[2 * pix + 1]};
          const int id[3] = {faces[3 * idf], faces[3 * idf + 1], faces[3 * idf + 2]};
          const V3 p0 = ld3(pos, id[0]), p1 = ld3(pos, id[1]), p2 = ld3(pos, id[2]);
          const V3 n0 = ld3(vnorm, id[0]), n1 = ld3(vnorm, id[1]), n2 = ld3(vnorm, id[2]);
          const V3 frag = bcc.x * p0 + bcc.y * p1 + bcc.z * p2;
          const V3 nUn = bcc.x * n0 + bcc.y * n1 + bcc.z * n2;                               // :231-233
          const float len = std::sqrt(nUn.x * nUn.x + nUn.y * nUn.y + nUn.z * nUn.z);
          V3 n = nUn / len;
          bool flipped = false;
          if (dot(n, d) > 0.f) { n = {-n.x, -n.y, -n.z}; flipped = true; }
          const V3 light = illum(n, sh);                                                     // :262
          const float jcoal[3] = {shading == Shaded ? light.x : 1.f, shading == Shaded ? light.y : 1.f, shading == Shaded ? light.z : 1.f};
          V3 alb = {0.f, 0.f, 0.f};                                                          // :274-319
          TexSample ts{};
          if (albedo == VertexColor) {
            alb = bcc.x * ld3(vc, id[0]) + bcc.y * ld3(vc, id[1]) + bcc.z * ld3(vc, id[2]);
          } else if (albedo == Textured) {
            ts = tex_coord(texcoords, idf, bcc, texW, texH);
            const V3 cLULV = ld3(tex, (long)texW * ts.lv + ts.lu), cLUHV = ld3(tex, (long)texW * ts.hv + ts.lu);
            const V3 cHULV = ld3(tex, (long)texW * ts.lv + ts.hu), cHUHV = ld3(tex, (long)texW * ts.hv + ts.hu);
            alb = (ts.v - ts.LV) * (((ts.u - ts.LU) * cLULV) + ((ts.HU - ts.u) * cHULV)) +
                  (ts.HV - ts.v) * (((ts.u - ts.LU) * cLUHV) + ((ts.HU - ts.u) * cHUHV));    // :311-312
          }
          const V3 g = ld3(render_grad, pix);
          const float gl[3] = {g.x * jcoal[0], g.y * jcoal[1], g.z * jcoal[2]};
          const float bc[3] = {bcc.x, bcc.y, bcc.z};
          if (albedo == VertexColor) {                                                       // :334-342
            for (int i = 0; i < 3; ++i) for (int ch = 0; ch < 3; ++ch) GC[3 * (size_t)id[i] + ch] += (double)(gl[ch] * bc[i]);
          } else if (albedo == Textured && !flipped) {                                       // :343-385
            if (g_texBilinear) {
              const float w4[4] = {(ts.v - ts.LV) * (ts.u - ts.LU), (ts.HV - ts.v) * (ts.u - ts.LU), (ts.v - ts.LV) * (ts.HU - ts.u), (ts.HV - ts.v) * (ts.HU - ts.u)};
              const size_t t4[4] = {(size_t)texW * ts.lv + ts.lu, (size_t)texW * ts.hv + ts.lu, (size_t)texW * ts.lv + ts.hu, (size_t)texW * ts.hv + ts.hu};
              for (int k = 0; k < 4; ++k) for (int ch = 0; ch < 3; ++ch) gt[t][t4[k] * 3 + ch] += (double)(gl[ch] * w4[k]);
            } else {
              double* GT = gt[t].data() + ((size_t)texW * ts.lv + ts.lu) * 3;
              GT[0] += gl[0]; GT[1] += gl[1]; GT[2] += gl[2];
            }
          }
          const float gA[3] = {g.x * alb.x, g.y * alb.y, g.z * alb.z};                       // GVCB * JCoLi
          if (shading == Shaded) {                                                           // :402-436
            const float Y[9] = {1.f, n.y, n.z, n.x, n.x * n.y, n.z * n.y, 3 * n.z * n.z - 1, n.x * n.z, (n.x * n.x) - (n.y * n.y)};
            for (int ch = 0; ch < 3; ++ch) for (int k = 0; k < 9; ++k) GS[9 * ch + k] += (double)(gA[ch] * Y[k]);
          }
          // ---- position gradient, "data to model" (:458-525) ----
          float q[3] = {0.f, 0.f, 0.f};
          if (shading == Shaded) {
            float JLiNo[3][3];                                                               // RendererUtil.h:371-391
            for (int i = 0; i < 3; ++i) {
              const float* s = sh + 9 * i;
              JLiNo[i][0] = s[3] + (s[4] * n.y) + (s[7] * n.z) + (s[8] * 2 * n.x);
              JLiNo[i][1] = s[1] + (s[4] * n.x) + (s[5] * n.z) + (s[8] * -2.f * n.y);
              JLiNo[i][2] = s[2] + (s[5] * n.y) + (s[6] * 6 * n.z) + (s[7] * n.x);
            }
            float u3[3];
            for (int j = 0; j < 3; ++j) u3[j] = gA[0] * JLiNo[0][j] + gA[1] * JLiNo[1][j] + gA[2] * JLiNo[2][j];
            const float l2 = len * len, l3 = l2 * len;                                       // getJNoNu :398-415
            const float un[3] = {nUn.x, nUn.y, nUn.z};
            float J[3][3];
            for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) J[i][j] = (i == j) ? (l2 - un[i] * un[i]) / l3 : -(un[i] * un[j]) / l3;
            for (int j = 0; j < 3; ++j) q[j] = u3[0] * J[0][j] + u3[1] * J[1][j] + u3[2] * J[2][j];
            // JNoBc (:650-663) then JBcVp = dJBCDVerpos (:670-861), forward mode over the 9 coordinates
            const V3 qv = {q[0], q[1], q[2]};
            const float r[3] = {dot(qv, n0), dot(qv, n1), dot(qv, n2)};
            const V3 N_ = cross(p1 - p0, p2 - p0);
            const float denom = dot(N_, N_), NdR = dot(d, N_);
            if (!(std::fabs(dot(normalize(d), normalize(N_))) < 0.001f || std::fabs(denom * denom) < 0.001f)) {
              const float tt = (dot(p0, N_) - dot(o, N_)) / NdR;
              const V3 Pp = o + tt * d;
              const V3 E[2] = {p2 - p1, p0 - p2}, vp[2] = {Pp - p1, Pp - p2};
              const V3 Cc[2] = {cross(E[0], vp[0]), cross(E[1], vp[1])};
              const V3 verts[3] = {p0, p1, p2};
              for (int var = 0; var < 9; ++var) {
                const int wv = var / 3, ax = var % 3;
                V3 e = {ax == 0 ? 1.f : 0.f, ax == 1 ? 1.f : 0.f, ax == 2 ? 1.f : 0.f};
                // dN/d(var): N = (v1-v0) x (v2-v0)
                V3 dN;
                if (wv == 0) dN = cross(V3{-e.x, -e.y, -e.z}, verts[2] - verts[0]) + cross(verts[1] - verts[0], V3{-e.x, -e.y, -e.z});
                else if (wv == 1) dN = cross(e, verts[2] - verts[0]);
                else dN = cross(verts[1] - verts[0], e);
                const V3 dV0 = (wv == 0) ? e : V3{0.f, 0.f, 0.f};
                const float dt = (-1.f / (NdR * NdR)) * dot(dN, d) * dot(p0 - o, N_) + (1.f / NdR) * (dot(dV0, N_) + dot(p0 - o, dN));
                float dJ[2];
                for (int abc = 0; abc < 2; ++abc) {
                  // dE: edge1 = v2 - v1, edge2 = v0 - v2 ; d(vp) = dt*dir - dV[abc+1]
                  V3 dE = {0.f, 0.f, 0.f};
                  if (abc == 0) { if (wv == 1) dE = V3{-e.x, -e.y, -e.z}; else if (wv == 2) dE = e; }
                  else { if (wv == 0) dE = e; else if (wv == 2) dE = V3{-e.x, -e.y, -e.z}; }
                  const V3 dVk = (wv == abc + 1) ? e : V3{0.f, 0.f, 0.f};
                  const float tmp = dot(N_, Cc[abc]);
                  const float dtmp = dot(dN, Cc[abc]) + dot(N_, cross(dE, vp[abc]) + cross(E[abc], dt * d - dVk));
                  dJ[abc] = dtmp / denom + tmp * (-1.f / (denom * denom)) * (2.f * dot(N_, dN));
                }
                const float dJ2 = -dJ[0] - dJ[1];
                GP[3 * (size_t)id[wv] + ax] += (double)(r[0] * dJ[0] + r[1] * dJ[1] + r[2] * dJ2);
              }
            }
          }
          // ---- "model to data" (:531-555) ----
          if (target_grad) {
            const int fs = imageFilter;
            V3 dIu = {0.f, 0.f, 0.f}, dIv = {0.f, 0.f, 0.f};                                 // imageGradient RendererUtil.h:566-620
            if (idw >= fs + 1 && idh >= fs + 1 && idw < W - (fs + 1) && idh < H - (fs + 1)) {
              const float* img = target_image + view * P * 3;
              float nf = 0.f;
              for (int y = -fs; y <= fs; ++y)
                for (int x = -fs; x <= fs; ++x) {
                  const V3 I = ld3(img, (long)(idh + y) * W + (idw + x));
                  const float den = (float)(x * x + y * y);
                  float Gu = 0.f, Gv = 0.f;
                  if (den != 0.f) { Gu = (float)x / den; Gv = (float)y / den; }
                  dIu = dIu + I * Gu; dIv = dIv + I * Gv;
                  nf += std::fabs(Gu);
                }
              dIu = dIu / nf; dIv = dIv / nf;
            }
            const V3 gtv = ld3(target_grad, pix);
            const float w0 = dot(gtv, dIu), w1 = dot(gtv, dIv);
            float M[3][4];                                                                   // getJProjection :275-332
            for (int r_ = 0; r_ < 3; ++r_) for (int c4 = 0; c4 < 4; ++c4)
              M[r_][c4] = cam.K[3 * r_] * cam.E[c4] + cam.K[3 * r_ + 1] * cam.E[4 + c4] + cam.K[3 * r_ + 2] * cam.E[8 + c4];
            const float Px = M[0][0] * frag.x + M[0][1] * frag.y + M[0][2] * frag.z + M[0][3];
            const float Py = M[1][0] * frag.x + M[1][1] * frag.y + M[1][2] * frag.z + M[1][3];
            const float Pz = M[2][0] * frag.x + M[2][1] * frag.y + M[2][2] * frag.z + M[2][3];
            if (std::fabs(Pz) > 0.0001f) {
              for (int j = 0; j < 3; ++j) {
                const float dpx = (1.f / Pz) * M[0][j] + (-Px / (Pz * Pz)) * M[2][j];
                const float dpy = (1.f / Pz) * M[1][j] + (-Py / (Pz * Pz)) * M[2][j];
                const float w2 = w0 * dpx + w1 * dpy;
                for (int i = 0; i < 3; ++i) GP[3 * (size_t)id[i] + j] += (double)(bc[i] * w2);
              }
            }
          }
          // ---- vertex-normal term: every face incident to the pixel's vertices (:559-615) ----
          if (shading == Shaded) {
            const V3 qv = {q[0], q[1], q[2]};
            for (int i = 0; i < 3; ++i) {
              const V3 qi = bc[i] * qv;
              for (int j = mesh.vfOff[id[i]]; j < mesh.vfOff[id[i] + 1]; ++j) {
                const int f2 = mesh.vfList[j];
                const int a = faces[3 * f2], b2 = faces[3 * f2 + 1], c2 = faces[3 * f2 + 2];
                const V3 vi = ld3(pos, a), vj = ld3(pos, b2), vk = ld3(pos, c2);
                // row-vector * J with J from getJ_vi / getJ_vj / getJ_vk (RendererUtil.h:422-539):
                // q*J_vi = (vj-vi) x q - (vk-vi) x q ; q*J_vj = (vk-vi) x q ; q*J_vk = q x (vj-vi)
                const V3 gi = cross(vj - vi, qi) - cross(vk - vi, qi), gj = cross(vk - vi, qi), gk = cross(qi, vj - vi);
                GP[3 * (size_t)a] += gi.x; GP[3 * (size_t)a + 1] += gi.y; GP[3 * (size_t)a + 2] += gi.z;
                GP[3 * (size_t)b2] += gj.x; GP[3 * (size_t)b2 + 1] += gj.y; GP[3 * (size_t)b2 + 2] += gj.z;
                GP[3 * (size_t)c2] += gk.x; GP[3 * (size_t)c2 + 1] += gk.y; GP[3 * (size_t)c2 + 2] += gk.z;
              }
            }
          }
        }
      }
    }
    for (size_t i = 0; i < (size_t)N * 3; ++i) {
      double sp = 0, sc = 0;
      for (int t = 0; t < T; ++t) { sp += gp[t][i]; sc += gc[t][i]; }
      vpos_grad[(size_t)b * N * 3 + i] = (float)sp;
      vcol_grad[(size_t)b * N * 3 + i] = (float)sc;
    }
    if (tex_grad)
      for (size_t i = 0; i < texN; ++i) {
        double s = 0;
        if (albedo == Textured) for (int t = 0; t < T; ++t) s += gt[t][i];
        tex_grad[(size_t)b * texN + i] = (float)s;
      }
    for (int i = 0; i < C * 27; ++i) {
      double s = 0;
      for (int t = 0; t < T; ++t) s += gs[t][i];
      sh_grad[(size_t)b * C * 27 + i] = (float)s;
    }
  }
  return 0;
}

// UV-space normal map (compute_normal_map): host rasterisation of every triangle in texture space
// (CUDABasedRasterization.cpp:237-298, rayTriangleIntersectHost :156-235) + renderNormalMapDevice
// (CUDABasedRasterization.cu:415-445).  Faces are visited in ascending order and later faces
// overwrite earlier ones (the reference's OpenMP loop is racy where UV triangles overlap; this is
// its serial order).  Also writes vertex_normal [B,C,N,3] like the forward.  covered (may be null):
// uint8 [texH,texW], 1 where some face covers the texel; tie: 1 where more than one face does.
int gvvo_normal_map(const int* faces, int F, const float* texcoords, int N, int C, int B, int texH, int texW,
                    const float* vertex_pos, float* vertex_normal, float* normal_map, unsigned char* covered, unsigned char* tie) {
  Mesh mesh(faces, F, texcoords, N);
  std::vector<float> table((size_t)texH * texW * 4, 0.f);   // (face, a, b, c), initial (0,0,0,0) (:250)
  std::vector<int> hits((size_t)texH * texW, 0);
  for (int f = 0; f < F; ++f) {
    const float* t = texcoords + 6 * (long)f;
    const V3 t0 = {texW * t[0], texH * (1.f - t[1]), 0.f}, t1 = {texW * t[2], texH * (1.f - t[3]), 0.f}, t2 = {texW * t[4], texH * (1.f - t[5]), 0.f};
    const int xMin = (int)std::fmax(std::fmin(t0.x, std::fmin(t1.x, t2.x)) - 2, 0), xMax = (int)std::fmin(std::fmax(t0.x, std::fmax(t1.x, t2.x)) + 2, texW);
    const int yMin = (int)std::fmax(std::fmin(t0.y, std::fmin(t1.y, t2.y)) - 2, 0), yMax = (int)std::fmin(std::fmax(t0.y, std::fmax(t1.y, t2.y)) + 2, texH);
    for (int x = xMin; x < xMax; ++x)
      for (int y = yMin; y < yMax; ++y) {
        const V3 d = {0.f, 0.f, -1.f};
        const V3 v0 = t0 / 1000.f, v1 = t1 / 1000.f, v2 = t2 / 1000.f, orig = V3{x + 0.5f, y + 0.5f, 1.f} / 1000.f;
        const V3 Nn = cross(v1 - v0, v2 - v0);
        const float nd = dot(d, Nn);
        if (std::fabs(nd) < 0.0000001f) continue;
        const float tt = (dot(v0, Nn) - dot(orig, Nn)) / nd;
        if (tt < 0) continue;
        const V3 P = orig + tt * d;
        if (dot(Nn, cross(v1 - v0, P - v0)) < 0) continue;
        float a = dot(Nn, cross(v2 - v1, P - v1));
        if (a < 0) continue;
        float b = dot(Nn, cross(v0 - v2, P - v2));
        if (b < 0) continue;
        const float den = dot(Nn, Nn);
        a /= den; b /= den;
        float* o = &table[((size_t)y * texW + x) * 4];
        o[0] = (float)f; o[1] = a; o[2] = b; o[3] = 1.f - a - b;
        hits[(size_t)y * texW + x]++;
      }
  }
  for (int b = 0; b < B; ++b) {
    const float* pos = vertex_pos + (long)b * N * 3;
    std::vector<V3> vn(N);
    for (int n = 0; n < N; ++n) {
      V3 s = {0.f, 0.f, 0.f};
      for (int i = mesh.vfOff[n]; i < mesh.vfOff[n + 1]; ++i) {
        const int f = mesh.vfList[i];
        const V3 v0 = ld3(pos, faces[3 * f]), v1 = ld3(pos, faces[3 * f + 1]), v2 = ld3(pos, faces[3 * f + 2]);
        const V3 fn = cross(v1 - v0, v2 - v0);
        s = (i == mesh.vfOff[n]) ? fn : s + fn;
      }
      vn[n] = s;
      for (int c = 0; c < C; ++c) { float* o = vertex_normal + (((long)b * C + c) * N + n) * 3; o[0] = s.x; o[1] = s.y; o[2] = s.z; }
    }
    for (long i = 0; i < (long)texH * texW; ++i) {
      const float* info = &table[i * 4];
      const int idf = (int)info[0];
      V3 n = vn[faces[3 * idf]] * info[1] + vn[faces[3 * idf + 1]] * info[2] + vn[faces[3 * idf + 2]] * info[3];
      const float len = std::sqrt(dot(n, n));
      if (len != 0.f) n = n / len;
      float* o = normal_map + ((long)b * texH * texW + i) * 3;
      o[0] = (n.x + 1.f) / 2.f; o[1] = (n.y + 1.f) / 2.f; o[2] = (n.z + 1.f) / 2.f;
    }
  }
  if (covered) for (size_t i = 0; i < hits.size(); ++i) covered[i] = hits[i] > 0;
  if (tie) for (size_t i = 0; i < hits.size(); ++i) tie[i] = hits[i] > 1;
  return 0;
}

int gvvo_max_threads() {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

}  // extern "C"
