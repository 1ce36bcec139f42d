// gvv_backward.cu -- backward pass of the rasteriser for sm_100a.
//
// Reference: renderBuffersGradGPU (CUDABasedRasterizationGrad.cu:624-635): one thread per pixel,
// 157 registers, ~214 scalar float atomics per covered pixel, 27 of them on the same 27 SH
// addresses for every pixel of a camera.  Here, for all B*C views in one go:
//
//   prep_kernel          camera records (ref :17-61), zero fill of all four gradient outputs + the vertex-normal
//                        gradient (ref :68-107), repack of the caller's 12-byte vertex arrays to float4
//   pixel_grad_kernel    one warp per 32-pixel scanline segment (segment_grad).  The 27 per-vertex values of a
//                        pixel (9 colour, 9 position, 9 vertex-normal gradient) are transposed
//                        through shared memory so that lane j sums value j over each run of
//                        pixels that see the same triangle: ONE warp-wide atomic per run instead
//                        of 27 per pixel.  SH gradients are reduced warp -> block -> 27 atomics.
//                        (pixel_grad_persistent_kernel: the same segments driven by persistent CTAs with a
//                        TMA ring of face tiles -- option bwd_persistent, measured slower, off.)
//   normal_term_kernel   mesh-space pass for the vertex-normal -> position term (ref :559-615):
//                        the reference loops over every face incident to the pixel's three
//                        vertices (27*deg atomics per pixel); that sum is linear in the per-pixel
//                        factor q, so we scatter Gn[v_i] += bcc_i*q per pixel and apply the
//                        cross-product Jacobians (RendererUtil.h:422-539) once per TRIANGLE here
//                        (9 atomics per triangle that received any gradient) -- same mathematics.
#include <algorithm>
#include "gvv_internal.h"

namespace gvv {

#define FULL_MASK 0xffffffffu

struct V3 { float x, y, z; };
__device__ __forceinline__ V3 v3(float x, float y, float z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ V3 operator+(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ V3 operator-(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ V3 operator*(float s, V3 a) { return v3(s * a.x, s * a.y, s * a.z); }
__device__ __forceinline__ float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ V3 cross(V3 a, V3 b) { return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
__device__ __forceinline__ V3 ldv3(const float* __restrict__ p, size_t i) { return v3(__ldg(p + 3 * i), __ldg(p + 3 * i + 1), __ldg(p + 3 * i + 2)); }

// ------------------------------------------------------------------------------------------------
struct ZeroArgs { float* p[5]; long long n[5]; };

// ------------------------------------------------------------------------------------------------
struct PixelParams {
  const float *render_grad, *target_grad, *vertex_pos, *vertex_color, *texture, *sh_coeff, *target_image,
      *vertex_normal, *bary, *texcoords, *target_du, *target_dv;
  const int32_t* face;
  const int4* faces4;
  const CamRec* cams;
  const float4 *pos4, *col4, *nor4;
  float *vpos_grad, *vcol_grad, *tex_grad, *sh_grad, *gnorm;
  int C, N, W, H, texH, texW, albedo, shading, imgFilter, texBilinear;
  int sharedBatch;   // 1: colour / texture / SH gradients of every batch element accumulate into ONE slice (parameters shared across the batch)
  float invC;
};

constexpr int kVals = 27;
constexpr int kShRows = 12;  // rows 27..38 of the warp buffer: g*albedo (3) and the SH basis (9) of every pixel, for the SH gradient
constexpr int kIdRows = 3;   // rows 39..41: the three vertex ids of every pixel's triangle (int bits), read at the end of a run
constexpr int kRow = 36;    // 32 pixels + 4 pad: a quarter-warp's float4 reads of 8 different rows hit 32 different banks
constexpr int kWarpBufFloats = (kVals + kShRows + kIdRows) * kRow;

// Tolerance-level arithmetic of the backward: reciprocal-multiply instead of IEEE divides
// (gradients are compared to rel-L2 1e-4; the visibility-critical ray uses the exact functions).
__device__ __forceinline__ float rcpf(float x) { return __frcp_rn(x); }

// d(alpha*a + beta*b)/d(v0,v1,v2) for the barycentrics of the ray/plane hit: the product [alpha beta gamma] *
// dJBCDVerpos (RendererUtil.h:670-861) with the third row = -(row0+row1) folded into alpha, beta.  The reference
// builds the 3x9 Jacobian column by column; this is the same derivative in reverse mode.  Early-outs as :682-685.
__device__ __forceinline__ void bary_vjp_fast(V3 o, V3 d, V3 v0, V3 v1, V3 v2, float alpha, float beta, V3& g0, V3& g1, V3& g2) {
  g0 = g1 = g2 = v3(0.f, 0.f, 0.f);
  const V3 e01 = v1 - v0, e02 = v2 - v0;
  const V3 N = cross(e01, e02);
  const float D = dot(N, N);
  const float nd = dot(d, N);
  // |dot(normalize(d), normalize(N))| < 0.001  <=>  nd^2 < 1e-6 * |d|^2 * D   (RendererUtil.h:682)
  if (nd * nd < 1.0e-6f * dot(d, d) * D || fabsf(D * D) < 0.001f) return;
  const float iD = rcpf(D), ind = rcpf(nd);
  const V3 w = v0 - o;
  const float t = dot(w, N) * ind;
  const V3 P = o + t * d;
  const V3 E1 = v2 - v1, p1 = P - v1, C1 = cross(E1, p1);
  const V3 E2 = v0 - v2, p2 = P - v2, C2 = cross(E2, p2);
  const float A = dot(N, C1), Bn = dot(N, C2);
  const float Ab = alpha * iD, Bb = beta * iD;
  const float Db = -(alpha * A + beta * Bn) * iD * iD;
  V3 Nb = Ab * C1 + Bb * C2 + (2.f * Db) * N;
  const V3 C1b = Ab * N, C2b = Bb * N;
  const V3 E1b = cross(p1, C1b), p1b = cross(C1b, E1);
  const V3 E2b = cross(p2, C2b), p2b = cross(C2b, E2);
  const float tb = dot(p1b + p2b, d);
  const float mb = tb * ind;
  Nb = Nb + mb * w + (-mb * t) * d;
  const V3 e01b = cross(e02, Nb), e02b = cross(Nb, e01);
  g0 = E2b + mb * N - e01b - e02b;
  g1 = e01b - p1b - E1b;
  g2 = e02b - p2b + E1b - E2b;
}

__device__ __forceinline__ V3 ld4(const float4* __restrict__ p, size_t i) { const float4 v = __ldg(p + i); return v3(v.x, v.y, v.z); }

// mbarrier + bulk async copy (TMA) helpers for the face-tile ring of pixel_grad_persistent_kernel
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, int bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, int parity) {
  const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
  unsigned done = 0;
  while (!done)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(a), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_load(void* smemDst, const void* gsrc, int bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :: "r"((unsigned)__cvta_generic_to_shared(smemDst)), "l"(gsrc), "r"(bytes), "r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}

// grid (W/32, H/32, V), 256 threads: the CTA owns a 32x32 pixel tile, warp w the 32-pixel scanline segments
// y = 32*by + 8*s + w of its four 8-row slabs s.  All four face ids of a thread are loaded up front, so an empty
// tile (about 40 % of them at 50 % coverage) costs ONE memory round trip and one barrier; the warps then run
// their segments without any block barrier, and the SH gradient is reduced once per tile.
constexpr int kSlabs = 4;

// SHADED is a template parameter: the shaded instance is the full chain; the shadeless one drops, at compile time,
// everything its gradients do not depend on (SH basis and light, the albedo VALUE and its texel / colour gathers,
// the shading-normal position term and its buffers), which the register allocator could not do behind a runtime flag.
// One 32-pixel scanline segment of one warp: per-pixel chain, run-aggregated scatter, SH partial sums (shsum: lane j < 27
// holds the running sum of SH gradient (ch,k) = (j / 9, j % 9)).  Shared by the one-tile-per-CTA and the persistent kernel.
template <bool SHADED, int ALBEDO>
__device__ __forceinline__ void segment_grad(const PixelParams& p, const CamRec& cam, const float4* __restrict__ shc4, const float4* __restrict__ camv,
                                             float* __restrict__ mybuf, const int lane, const int view, const int b, const int x, const int y,
                                             const size_t pix, const int face, float& shsum) {
  constexpr bool shaded = SHADED;
  constexpr int albedo = ALBEDO;       // vertexColor | textured | foregroundMask (the other modes have no gradient)
  float* mine = mybuf + lane;   // value j of this lane's pixel lives at mine[j * kRow]
  const bool covered = face >= 0;
  const unsigned cv = __ballot_sync(FULL_MASK, covered);
  if (cv == 0) return;
  float gA[3] = {0.f, 0.f, 0.f};
  float Y[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) Y[k] = 0.f;

  {
    if (covered) {
      // ---- per-pixel setup (CUDABasedRasterizationGrad.cu:205-240) ----
      const float4 r0 = camv[0], r1 = camv[1], r2 = camv[2], ro4 = camv[3];
      const float Pv[12] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w, r2.x, r2.y, r2.z, r2.w};
      const float rov[3] = {ro4.x, ro4.y, ro4.z};
      const F3 rdx = ray_dir_exact(Pv, rov, (float)x + 0.5f, (float)y + 0.5f);
      const V3 d = v3(rdx.x, rdx.y, rdx.z), o = v3(ro4.x, ro4.y, ro4.z);
      const int4 fc = __ldg(p.faces4 + face);
      const float2 ab = __ldcs(reinterpret_cast<const float2*>(p.bary) + pix);      // image-sized streams are read once: evict-first, so that they do not push the mesh gathers out of the L1 (pixel_grad 0.218 -> 0.216 ms)
      const float bc[3] = {ab.x, ab.y, 1.f - ab.x - ab.y};
      mine[(kVals + kShRows + 0) * kRow] = __int_as_float(fc.x);   // vertex ids for the run-end atomics of the scatter stage
      mine[(kVals + kShRows + 1) * kRow] = __int_as_float(fc.y);
      mine[(kVals + kShRows + 2) * kRow] = __int_as_float(fc.z);
      const float4* pos = p.pos4 + (size_t)b * p.N;
      const float4* nor = p.nor4 + (size_t)view * p.N;
      // (conditions that start with `shaded ||` are compile-time true in the shaded instance)
      V3 p0 = v3(0.f, 0.f, 0.f), p1 = p0, p2 = p0, n0 = p0, n1 = p0, n2 = p0;
      if (shaded || p.target_grad) { p0 = ld4(pos, fc.x); p1 = ld4(pos, fc.y); p2 = ld4(pos, fc.z); }   // positions: shading-normal and model-to-data terms
      const bool needNormal = shaded || albedo == GVV_ALBEDO_TEXTURED;                                   // pixel normal: shading, and the flipped-normal rule of the texture gradient
      if (needNormal) { n0 = ld4(nor, fc.x); n1 = ld4(nor, fc.y); n2 = ld4(nor, fc.z); }
      const V3 nUn = needNormal ? bc[0] * n0 + bc[1] * n1 + bc[2] * n2 : v3(0.f, 0.f, 1.f);
      const float len2 = dot(nUn, nUn);
      const float ilen = rsqrtf(len2);
      V3 n = ilen * nUn;
      const bool flipped = needNormal && dot(n, d) > 0.f;
      if (flipped) n = v3(-n.x, -n.y, -n.z);

      // SH basis (getIllum / getJLiGm, RendererUtil.h:179-214,351-364)
      Y[0] = 1.f; Y[1] = n.y; Y[2] = n.z; Y[3] = n.x; Y[4] = n.x * n.y; Y[5] = n.z * n.y;
      Y[6] = 3.f * n.z * n.z - 1.f; Y[7] = n.x * n.z; Y[8] = n.x * n.x - n.y * n.y;
      float light[3];
      V3 jl[3];      // JLiNo row of every channel (RendererUtil.h:371-391), from the same three 128-bit reads as the light
#pragma unroll
      for (int ch = 0; ch < 3; ++ch) {
        const float4 a4 = shc4[3 * ch], b4 = shc4[3 * ch + 1], c4 = shc4[3 * ch + 2];
        const float sc[9] = {a4.x, a4.y, a4.z, a4.w, b4.x, b4.y, b4.z, b4.w, c4.x};
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < 9; ++k) s += sc[k] * Y[k];
        light[ch] = s;
        if (shaded)
          jl[ch] = v3(sc[3] + sc[4] * n.y + sc[7] * n.z + sc[8] * 2.f * n.x,
                      sc[1] + sc[4] * n.x + sc[5] * n.z + sc[8] * -2.f * n.y,
                      sc[2] + sc[5] * n.y + sc[6] * 6.f * n.z + sc[7] * n.x);
      }
      const float3 g = make_float3(__ldcs(p.render_grad + 3 * pix), __ldcs(p.render_grad + 3 * pix + 1), __ldcs(p.render_grad + 3 * pix + 2));
      const float gl[3] = {shaded ? g.x * light[0] : g.x, shaded ? g.y * light[1] : g.y, shaded ? g.z * light[2] : g.z};

      // ---- albedo (:242-319) and its gradients (:327-395) ----
      float alb[3] = {0.f, 0.f, 0.f};
      if (albedo == GVV_ALBEDO_VERTEX_COLOR) {
        if (shaded) {      // the albedo VALUE only feeds the SH and shading-normal gradients
          const float4* col = p.col4 + (size_t)b * p.N;
          const V3 c0 = ld4(col, fc.x), c1 = ld4(col, fc.y), c2 = ld4(col, fc.z);
          const V3 al = bc[0] * c0 + bc[1] * c1 + bc[2] * c2;
          alb[0] = al.x; alb[1] = al.y; alb[2] = al.z;
        }
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
          for (int ch = 0; ch < 3; ++ch) mine[(i * 3 + ch) * kRow] = gl[ch] * bc[i];
      } else if (albedo == GVV_ALBEDO_TEXTURED) {
        const float* tc = p.texcoords + (size_t)face * 6;
        float u = (__ldg(tc + 0) * bc[0] + __ldg(tc + 2) * bc[1] + __ldg(tc + 4) * bc[2]) * p.texW;
        float v = ((1.f - __ldg(tc + 1)) * bc[0] + (1.f - __ldg(tc + 3)) * bc[1] + (1.f - __ldg(tc + 5)) * bc[2]) * p.texH;
        u = fminf(fmaxf(u, 0.f), (float)(p.texW - 1));
        v = fminf(fmaxf(v, 0.f), (float)(p.texH - 1));
        const float LU = (float)(int)(u - 0.5f) + 0.5f, HU = (float)(int)(u - 0.5f) + 1.5f;
        const float LV = (float)(int)(v - 0.5f) + 0.5f, HV = (float)(int)(v - 0.5f) + 1.5f;
        const float* tex = p.texture + (size_t)b * p.texH * p.texW * 3;
        const int lu = (int)LU, hu = min((int)HU, p.texW - 1), lv = (int)LV, hv = min((int)HV, p.texH - 1);   // hu, hv only clamp for 1-texel-wide textures
        // bilinear mix exactly as written in :311-312 (the forward uses the nearest texel); only shaded modes use the value
        if (shaded)
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
          const float cLULV = __ldg(tex + 3 * ((size_t)p.texW * lv + lu) + ch), cLUHV = __ldg(tex + 3 * ((size_t)p.texW * hv + lu) + ch);
          const float cHULV = __ldg(tex + 3 * ((size_t)p.texW * lv + hu) + ch), cHUHV = __ldg(tex + 3 * ((size_t)p.texW * hv + hu) + ch);
          alb[ch] = (v - LV) * ((u - LU) * cLULV + (HU - u) * cHULV) + (HV - v) * ((u - LU) * cLUHV + (HU - u) * cHUHV);
        }
        if (!flipped) {
          float* tgb = p.tex_grad + (size_t)(p.sharedBatch ? 0 : b) * p.texH * p.texW * 3;
          if (p.texBilinear) {
            // non-default variant: the four weighted adds the reference has commented out (:361-378), weights as written there
            const float wLULV = (v - LV) * (u - LU), wLUHV = (HV - v) * (u - LU), wHULV = (v - LV) * (HU - u), wHUHV = (HV - v) * (HU - u);
            float* t0 = tgb + ((size_t)p.texW * lv + lu) * 3; float* t1 = tgb + ((size_t)p.texW * hv + lu) * 3;
            float* t2 = tgb + ((size_t)p.texW * lv + hu) * 3; float* t3 = tgb + ((size_t)p.texW * hv + hu) * 3;
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) {
              atomicAdd(t0 + ch, gl[ch] * wLULV); atomicAdd(t1 + ch, gl[ch] * wLUHV);
              atomicAdd(t2 + ch, gl[ch] * wHULV); atomicAdd(t3 + ch, gl[ch] * wHUHV);
            }
          } else {          // unweighted add to texel (LV,LU) (:382-384)
            float* tg = tgb + ((size_t)p.texW * lv + lu) * 3;
            atomicAdd(tg + 0, gl[0]); atomicAdd(tg + 1, gl[1]); atomicAdd(tg + 2, gl[2]);
          }
        }
      }
      gA[0] = g.x * alb[0]; gA[1] = g.y * alb[1]; gA[2] = g.z * alb[2];
      if (shaded) {   // parked in the warp buffer now, so that Y and gA do not stay live across the scatter stage
        mine[(kVals + 0) * kRow] = gA[0]; mine[(kVals + 1) * kRow] = gA[1]; mine[(kVals + 2) * kRow] = gA[2];
#pragma unroll
        for (int k = 0; k < 9; ++k) mine[(kVals + 3 + k) * kRow] = Y[k];
      }

      float pos9[9];
#pragma unroll
      for (int j = 0; j < 9; ++j) pos9[j] = 0.f;
      if (shaded) {
        // ---- position gradient through the shading normal (:458-525) ----
        V3 u3 = v3(0.f, 0.f, 0.f);   // (g*albedo) * JLiNo  (RendererUtil.h:371-391)
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) u3 = u3 + gA[ch] * jl[ch];
        // * JNoNu (:398-415): (len^2 I - nUn nUn^T) / len^3 = (u - (u.n^)n^) / len, UNflipped normal n^ = nUn/len
        const V3 nh = ilen * nUn;
        const float un = dot(u3, nh);
        const V3 q = ilen * (u3 - un * nh);
        // * JNoBc * JBcVp: direct dependence of the barycentrics on the triangle's own vertices
        const float r0 = dot(q, n0), r1 = dot(q, n1), r2 = dot(q, n2);
        V3 g0, g1, g2;
        bary_vjp_fast(o, d, p0, p1, p2, r0 - r2, r1 - r2, g0, g1, g2);
        pos9[0] = g0.x; pos9[1] = g0.y; pos9[2] = g0.z;
        pos9[3] = g1.x; pos9[4] = g1.y; pos9[5] = g1.z;
        pos9[6] = g2.x; pos9[7] = g2.y; pos9[8] = g2.z;
        // vertex-normal gradient, finished in normal_term_kernel
#pragma unroll
        for (int i = 0; i < 3; ++i) { mine[(18 + i * 3) * kRow] = bc[i] * q.x; mine[(19 + i * 3) * kRow] = bc[i] * q.y; mine[(20 + i * 3) * kRow] = bc[i] * q.z; }
      }

      // ---- model-to-data term (:531-555) ----
      if (p.target_grad) {
        const int fs = p.imgFilter;
        V3 dIu = v3(0.f, 0.f, 0.f), dIv = v3(0.f, 0.f, 0.f);
        if (p.target_du) {            // precomputed once per target (gvv_image_gradient + gvv_set_target_gradient)
          dIu = ldv3(p.target_du, pix); dIv = ldv3(p.target_dv, pix);
        } else if (x >= fs + 1 && y >= fs + 1 && x < p.W - (fs + 1) && y < p.H - (fs + 1)) {   // imageGradient, RendererUtil.h:566-620
          const float* img = p.target_image + (size_t)view * p.W * p.H * 3;
          float norm = 0.f;
          for (int yy = -fs; yy <= fs; ++yy)
            for (int xx = -fs; xx <= fs; ++xx) {
              const V3 I = ldv3(img, (size_t)(y + yy) * p.W + (x + xx));
              const float den = (float)(xx * xx + yy * yy);
              float Gu = 0.f, Gv = 0.f;
              if (den != 0.f) { Gu = (float)xx / den; Gv = (float)yy / den; }
              dIu = dIu + Gu * I; dIv = dIv + Gv * I;
              norm += fabsf(Gu);
            }
          const float inorm = 1.f / norm;
          dIu = inorm * dIu; dIv = inorm * dIv;
        }
        const V3 gt = ldv3(p.target_grad, pix);
        const float w0 = dot(gt, dIu), w1 = dot(gt, dIv);
        // getJProjection (RendererUtil.h:275-332): d(K*E*[p;1] after divide)/dp at the fragment
        const V3 fp = bc[0] * p0 + bc[1] * p1 + bc[2] * p2;
        float M[3][4];   // K * E
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
          for (int c4 = 0; c4 < 4; ++c4) M[r][c4] = cam.K[3 * r] * cam.E[c4] + cam.K[3 * r + 1] * cam.E[4 + c4] + cam.K[3 * r + 2] * cam.E[8 + c4];
        const float Px = M[0][0] * fp.x + M[0][1] * fp.y + M[0][2] * fp.z + M[0][3];
        const float Py = M[1][0] * fp.x + M[1][1] * fp.y + M[1][2] * fp.z + M[1][3];
        const float Pz = M[2][0] * fp.x + M[2][1] * fp.y + M[2][2] * fp.z + M[2][3];
        if (fabsf(Pz) > 0.0001f) {
          const float iz = 1.f / Pz, kx = -Px * iz * iz, ky = -Py * iz * iz;
          float w2[3];
#pragma unroll
          for (int j = 0; j < 3; ++j) w2[j] = w0 * (iz * M[0][j] + kx * M[2][j]) + w1 * (iz * M[1][j] + ky * M[2][j]);
#pragma unroll
          for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) pos9[i * 3 + j] += bc[i] * w2[j];
        }
      }
#pragma unroll
      for (int j = 0; j < 9; ++j) mine[(9 + j) * kRow] = pos9[j];
    } else if (cv != FULL_MASK) {
      // pixels that see no triangle contribute zeros, so that the scatter stage can add every pixel of the row
#pragma unroll
      for (int j = 0; j < kVals; ++j) mine[j * kRow] = 0.f;
    }

    // ---- run-aggregated scatter: lane j owns value j ----
    const int prevFace = __shfl_up_sync(FULL_MASK, face, 1);
    const unsigned head = __ballot_sync(FULL_MASK, covered && (lane == 0 || prevFace != face));
    const unsigned cont = (cv & ~head) >> 1;         // bit l: lane l+1 continues lane l's run
    const unsigned endm = cv & ~cont;
    __syncwarp();
    const int arr = lane / 9, vi = (lane % 9) / 3, comp = lane % 3;
    const bool active = lane < kVals && ((arr == 0 && albedo == GVV_ALBEDO_VERTEX_COLOR) ||
                                         (arr == 1 && (shaded || p.target_grad)) || (arr == 2 && shaded));
    float* base = arr == 0 ? p.vcol_grad : (arr == 1 ? p.vpos_grad : p.gnorm);
    const int vstride = arr == 2 ? 4 : 3;          // gnorm is float4-strided for aligned gathers in normal_term_kernel
    base += (size_t)((arr == 0 && p.sharedBatch) ? 0 : b) * p.N * vstride + comp;
    // Lane j walks row j (the 32 pixels' value j) with a segmented running sum: flushed with ONE
    // warp-wide atomic at the end of every run of equal face id and reset to zero there (pixels outside
    // a run hold zeros).  endm is warp-uniform, so the 32 steps are unrolled with uniform branches.
    const float4* row = reinterpret_cast<const float4*>(mybuf + (lane < kVals ? lane : 0) * kRow);
    const float* idrow = mybuf + (kVals + kShRows + vi) * kRow;
    float acc = 0.f;
#pragma unroll     // (rolling this loop to shrink the code was measured slower: 0.252 -> 0.257 ms)
    for (int g = 0; g < 8; ++g) {
      const float4 v4 = row[g];
      const float vv[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int l = 4 * g + u;
        acc += vv[u];
        if ((endm >> l) & 1u) {
          if (active && acc != 0.f) {
            const int vid = __float_as_int(idrow[l]);        // vertex vi of the triangle pixel l sees (one shared-memory read)
            atomicAdd(base + (unsigned)(vid * 3 + (arr == 2 ? vid : 0)), acc);   // 32-bit index arithmetic (N * 4 < 2^32): the 64-bit form cost nine instructions per flush
          }
          acc = 0.f;
        }
      }
    }
    __syncwarp();
  }

  // ---- SH gradient: 12 values per pixel (g*albedo, Y) stored value-major; lane j = (ch,k) sums
  // gA[ch]*Y[k] over the 32 pixels with float4 reads (uncovered pixels store zeros) ----
  if (shaded) {
    if (!covered) {   // zeros where not covered
#pragma unroll
      for (int k = 0; k < kShRows; ++k) mine[(kVals + k) * kRow] = 0.f;
    }
    __syncwarp();
    if (lane < kVals) {
      const float4* pa = reinterpret_cast<const float4*>(mybuf + (kVals + lane / 9) * kRow);
      const float4* py = reinterpret_cast<const float4*>(mybuf + (kVals + 3 + lane % 9) * kRow);
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        const float4 a4 = pa[g], y4 = py[g];
        shsum = fmaf(a4.x, y4.x, fmaf(a4.y, y4.y, fmaf(a4.z, y4.z, fmaf(a4.w, y4.w, shsum))));
      }
    }
    __syncwarp();
  }
}

// The per-view constants of both kernels are laid out for 128-bit broadcast reads: the kernel was bound by the L1 data
// pipe (l1tex__data_pipe_lsu_wavefronts 79 % of peak, profiles/r02_ncu_summaries.md), where a scalar LDS costs a
// wavefront like a 128-bit one: 9 instead of 51 reads per pixel for the SH coefficients, 4 instead of 15 for the ray.
//   shc4[9]: channel ch has its coefficients 0..8 in shc4[3 ch .. 3 ch + 2] (three pad words); camv[4]: rows 0..2 of (K E)^-1, ray origin
__device__ __forceinline__ void stage_view_constants(const PixelParams& p, int view, int t, CamRec* cam, float4* shc4, float4* camv) {
  // t = thread index within a group of >= 144 threads
  if (t < 64) { if (p.target_grad) reinterpret_cast<float*>(cam)[t] = __ldg(reinterpret_cast<const float*>(p.cams + view) + t); }   // K, E: model-to-data term only
  else if (t < 64 + 36) { const int i = t - 64, ch = i / 12, k = i - 12 * ch; reinterpret_cast<float*>(shc4)[i] = k < 9 ? __ldg(p.sh_coeff + (size_t)view * 27 + ch * 9 + k) : 0.f; }
  else if (t >= 128 && t < 128 + 16) {
    const int i = t - 128;
    const CamRec* cr = p.cams + view;
    reinterpret_cast<float*>(camv)[i] = i < 12 ? __ldg(cr->Pinv + i) : (i < 15 ? __ldg(cr->ro + (i - 12)) : 0.f);
  }
}

// One 32x32 tile per CTA (any image size).  The shaded instances run at 80 registers, 3 CTAs/SM: at 64 registers
// (4 CTAs/SM) the per-pixel chain spilled ~90 bytes in its hot path -- local-memory traffic through the same L1 data
// pipe that bounds the kernel (vertexColor + shaded: 0.236 -> 0.214 ms).  The shadeless instances fit 64 registers
// without spilling and keep 4 CTAs/SM (textured + shadeless, 32 views: 0.29 ms against 0.33 ms at 3 CTAs/SM).
template <bool SHADED, int ALBEDO>
constexpr int pixel_grad_ctas_per_sm() { return SHADED ? 3 : 4; }      // (textured + shaded, 32 views: 0.617 ms at 3, 0.635 ms at 4)

template <bool SHADED, int ALBEDO>
__global__ void __launch_bounds__(256, (pixel_grad_ctas_per_sm<SHADED, ALBEDO>()))
pixel_grad_kernel(const PixelParams p) {
  chain_wait(); chain_trigger();
  extern __shared__ __align__(16) float buf_dyn[];   // per warp: (kVals + kShRows + kIdRows) rows, value-major, kRow floats per value (32 pixels + pad)
  __shared__ float shPart[8][kVals];
  __shared__ CamRec cam;
  __shared__ float4 shc4[9];
  __shared__ float4 camv[4];

  const int view = blockIdx.z, b = (int)(((float)view + 0.5f) * p.invC);   // = view / C without the integer division (exact below 2^22 views)
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int x = blockIdx.x * 32 + lane;
  const size_t viewBase = (size_t)view * p.W * p.H;
  bool any = false;
#pragma unroll
  for (int s = 0; s < kSlabs; ++s) {
    const int ys = blockIdx.y * 32 + s * 8 + warp;
    const int f = (x < p.W && ys < p.H) ? __ldg(p.face + viewBase + (size_t)ys * p.W + x) : -1;
    any = any || f >= 0;
  }
  // constant staging overlaps the latency of the face loads; ONE barrier publishes it and
  // tells whether anything is visible in this 32x32 tile
  stage_view_constants(p, view, tid, &cam, shc4, camv);
  if (__syncthreads_or(any) == 0) return;

  float* mybuf = buf_dyn + warp * kWarpBufFloats;
  float shsum = 0.f;            // lane j < 27: SH gradient (ch,k) summed over this warp's segments
#pragma unroll 1
  for (int slab = 0; slab < kSlabs; ++slab) {
    const int y = blockIdx.y * 32 + slab * 8 + warp;
    const size_t pix = viewBase + (size_t)y * p.W + x;
    const int face = (x < p.W && y < p.H) ? __ldg(p.face + pix) : -1;      // second read of the line: L1/L2 hit
    segment_grad<SHADED, ALBEDO>(p, cam, shc4, camv, mybuf, lane, view, b, x, y, pix, face, shsum);
  }

  // warp -> block -> 27 atomics per tile
  if (lane < kVals) shPart[warp][lane] = shsum;
  __syncthreads();
  if (SHADED && tid < kVals) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += shPart[w][tid];
    if (s != 0.f) atomicAdd(p.sh_grad + (size_t)(p.sharedBatch ? view - b * p.C : view) * 27 + tid, s);
  }
}

// Persistent variant for images made of whole tiles (W, H multiples of 32, 16-byte aligned face buffer): one CTA per
// resident slot, each pulling 32x32 tiles (view-major order) from a global counter.  The 4 KB face tile -- the first
// hop of the face -> triangle -> vertices chain, and all an empty tile (40 % of them at 50 % coverage) ever needs --
// arrives by bulk async copies (TMA, one 128-byte row per copy) into a ring of kFaceRing shared-memory buffers, each
// with its mbarrier, issued kFaceRing tiles ahead (a slot is refilled as soon as its face ids sit in registers): the
// load latency and the per-CTA prologue (launch, constants) that the one-tile-per-CTA kernel pays 1024 times per view
// are off the critical path.  Per-view constants are double-buffered (a warp may enter the next view's first tile
// while another still finishes the previous one), SH partial sums stay in registers until the view changes.
constexpr int kFaceRing = 4;

template <bool SHADED, int ALBEDO>
__global__ void __launch_bounds__(256, 3)
pixel_grad_persistent_kernel(const PixelParams p, int* __restrict__ tileCounter, int tilesX, int tilesPerView, int totalTiles) {
  chain_wait(); chain_trigger();
  extern __shared__ __align__(16) float buf_dyn[];
  __shared__ CamRec cam[2];
  __shared__ float4 shc4[2][9];
  __shared__ float4 camv[2][4];
  __shared__ __align__(8) unsigned long long mbar[kFaceRing];
  __shared__ int ringTile[kFaceRing];
  int* sFace = reinterpret_cast<int*>(buf_dyn + 8 * kWarpBufFloats);      // kFaceRing tiles of 1024 face ids

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float* mybuf = buf_dyn + warp * kWarpBufFloats;
  // tile t of the ring slot s: 32 rows of 128 bytes, one bulk copy per lane of warp 0
  auto issue = [&](int t, int s) {
    const int view = t / tilesPerView, r = t - view * tilesPerView, ty = r / tilesX, tx = r - ty * tilesX;
    const int32_t* src = p.face + (size_t)view * p.W * p.H + (size_t)(ty * 32 + lane) * p.W + tx * 32;
    mbar_expect_tx(&mbar[s], 128);
    bulk_load(sFace + s * 1024 + lane * 32, src, 128, &mbar[s]);
  };
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < kFaceRing; ++s) mbar_init(&mbar[s], 32);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    ringTile[0] = blockIdx.x;                                   // the first tile is static, the rest comes from the counter
#pragma unroll
    for (int s = 1; s < kFaceRing; ++s) ringTile[s] = atomicAdd(tileCounter, 1);
  }
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int s = 0; s < kFaceRing; ++s) if (ringTile[s] < totalTiles) issue(ringTile[s], s);
  }
  int curView = -1, b = 0, cs = 0;      // cs: constant slot, toggled at every view change (consecutive tiles of a CTA may be several views apart)
  float shsum = 0.f;
  auto flush_sh = [&]() {
    if (SHADED && lane < kVals && shsum != 0.f) atomicAdd(p.sh_grad + (size_t)(p.sharedBatch ? curView - b * p.C : curView) * 27 + lane, shsum);
    shsum = 0.f;
  };
#pragma unroll 1
  for (int it = 0;; ++it) {
    const int s = it % kFaceRing;
    const int tile = ringTile[s];
    if (tile >= totalTiles) break;                              // the counter only grows: every later slot is past the end too
    const int view = tile / tilesPerView, r = tile - view * tilesPerView, ty = r / tilesX, tx = r - ty * tilesX;
    if (view != curView) {
      if (curView >= 0) flush_sh();
      curView = view; b = (int)(((float)view + 0.5f) * p.invC); cs ^= 1;
      stage_view_constants(p, view, tid, &cam[cs], shc4[cs], camv[cs]);   // every thread sees the change at the same tile; published by the barrier below
    }
    mbar_wait(&mbar[s], (it / kFaceRing) & 1);
    const int* tf = sFace + s * 1024;
    int f4[kSlabs];
    bool any = false;
#pragma unroll
    for (int sl = 0; sl < kSlabs; ++sl) { f4[sl] = tf[(sl * 8 + warp) * 32 + lane]; any = any || f4[sl] >= 0; }
    const int busy = __syncthreads_or(any);     // every thread holds its face ids (and the tile id) in registers: the slot is refilled right away
    if (warp == 0) {
      int tn = 0;
      if (lane == 0) { tn = atomicAdd(tileCounter, 1); ringTile[s] = tn; }      // read again kFaceRing iterations (barriers) later
      tn = __shfl_sync(FULL_MASK, tn, 0);
      if (tn < totalTiles) issue(tn, s);
    }
    if (!busy) continue;
    const int x = tx * 32 + lane;
    const size_t viewBase = (size_t)view * p.W * p.H;
#pragma unroll 1
    for (int sl = 0; sl < kSlabs; ++sl) {
      const int y = ty * 32 + sl * 8 + warp;
      const int face = sl == 0 ? f4[0] : (sl == 1 ? f4[1] : (sl == 2 ? f4[2] : f4[3]));
      segment_grad<SHADED, ALBEDO>(p, cam[cs], shc4[cs], camv[cs], mybuf, lane, view, b, x, y, viewBase + (size_t)y * p.W + x, face, shsum);
    }
  }
  if (curView >= 0) flush_sh();
}

// Repack of the caller's 12-byte AoS vertex arrays into aligned float4 (one 16-B gather instead of
// three 4-B ones per vertex in pixel_grad_kernel) + zero fill of all gradient outputs.
struct PrepArgs {
  ZeroArgs z;
  const float *vertex_pos, *vertex_color, *vertex_normal;
  float4 *pos4, *col4, *nor4;
  long long nBN, nVN;
  const float *extr, *intr;
  CamRec* cams;
  int V;
  int* tileCounter; int counterInit;   // work counter of pixel_grad_persistent_kernel: the first gridDim.x tiles are static
};

__global__ void prep_kernel(PrepArgs a) {
  chain_wait(); chain_trigger();
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long t0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t0 < a.V) fill_camrec(a.extr, a.intr, a.cams, (int)t0);   // camera records for pixel_grad_kernel (ref :17-61)
  if (t0 == 0 && a.tileCounter) *a.tileCounter = a.counterInit;
  for (long long i = t0; i < a.nBN; i += stride) {
    a.pos4[i] = make_float4(__ldg(a.vertex_pos + 3 * i), __ldg(a.vertex_pos + 3 * i + 1), __ldg(a.vertex_pos + 3 * i + 2), 0.f);
    if (a.vertex_color) a.col4[i] = make_float4(__ldg(a.vertex_color + 3 * i), __ldg(a.vertex_color + 3 * i + 1), __ldg(a.vertex_color + 3 * i + 2), 0.f);
  }
  for (long long i = t0; i < a.nVN; i += stride)
    a.nor4[i] = make_float4(__ldg(a.vertex_normal + 3 * i), __ldg(a.vertex_normal + 3 * i + 1), __ldg(a.vertex_normal + 3 * i + 2), 0.f);
#pragma unroll
  for (int r = 0; r < 5; ++r) {
    float* p = a.z.p[r];
    if (!p) continue;
    const long long n = a.z.n[r];
    if ((reinterpret_cast<uintptr_t>(p) & 15) == 0) {
      float4* p4 = reinterpret_cast<float4*>(p);
      const long long n4 = n >> 2;
      for (long long i = t0; i < n4; i += stride) p4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      for (long long i = (n4 << 2) + t0; i < n; i += stride) p[i] = 0.f;
    } else {
      for (long long i = t0; i < n; i += stride) p[i] = 0.f;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// One thread per (batch element, triangle): S = sum of Gn over the triangle's distinct vertices, then
// the three cross-product Jacobians (getJ_vi/vj/vk, RendererUtil.h:422-539) scattered to its vertices.
// Triangles none of whose vertices received a normal gradient (about 2/3 of them: back faces,
// occluded parts) leave after three 16-byte loads.
// The first ar.blocks CTAs (of batch element 0) do not touch triangles: they run the one-shot all-reduce of the
// shared-parameter gradients over peer memory (gvv_collective.cuh) while the others compute -- SH and colour gradients
// are final once pixel_grad_kernel has ended, this kernel only adds to vertex_pos_grad.
__global__ void __launch_bounds__(256)
normal_term_kernel(const float4* __restrict__ pos4, const float4* __restrict__ gnorm4, const int4* __restrict__ faces4,
                   float* __restrict__ vpos_grad, int N, int F, const ARParams ar) {
  chain_wait(); chain_trigger();
  int bx = blockIdx.x;
  if (blockIdx.y == 0) {
    if (bx < ar.blocks) { allreduce_block(ar, bx); return; }
    bx -= ar.blocks;
  }
  const int f = bx * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (f >= F) return;
  const int4 fc = __ldg(faces4 + f);
  const float4* gn = gnorm4 + (size_t)b * N;
  V3 S = ld4(gn, fc.x);
  if (fc.y != fc.x) S = S + ld4(gn, fc.y);
  if (fc.z != fc.x && fc.z != fc.y) S = S + ld4(gn, fc.z);
  if (S.x == 0.f && S.y == 0.f && S.z == 0.f) return;
  const float4* pos = pos4 + (size_t)b * N;
  const V3 pi = ld4(pos, fc.x), pj = ld4(pos, fc.y), pk = ld4(pos, fc.z);
  const V3 e1 = pj - pi, e2 = pk - pi;
  // S * J_vi = e1 x S - e2 x S ; S * J_vj = e2 x S ; S * J_vk = S x e1
  const V3 gj = cross(e2, S), gk = cross(S, e1);
  const V3 gi = v3(-gj.x - gk.x, -gj.y - gk.y, -gj.z - gk.z);
  float* out = vpos_grad + (size_t)b * N * 3;
  atomicAdd(out + 3 * (size_t)fc.x, gi.x); atomicAdd(out + 3 * (size_t)fc.x + 1, gi.y); atomicAdd(out + 3 * (size_t)fc.x + 2, gi.z);
  atomicAdd(out + 3 * (size_t)fc.y, gj.x); atomicAdd(out + 3 * (size_t)fc.y + 1, gj.y); atomicAdd(out + 3 * (size_t)fc.y + 2, gj.z);
  atomicAdd(out + 3 * (size_t)fc.z, gk.x); atomicAdd(out + 3 * (size_t)fc.z + 1, gk.y); atomicAdd(out + 3 * (size_t)fc.z + 2, gk.z);
}

// the same all-reduce as a launch of its own: ranges that include vertex_pos_grad (cameras of one batch element split
// over ranks) are only final after normal_term_kernel
__global__ void __launch_bounds__(256) allreduce_kernel(const ARParams ar) {
  chain_wait(); chain_trigger();
  allreduce_block(ar, blockIdx.x);
}

// ------------------------------------------------------------------------------------------------
int launch_camera(const float* extr, const float* intr, CamRec* cams, int* bigCount, int V, cudaStream_t st);

int launch_backward(const BwdArgs& a, cudaStream_t st, KernelTimer* tm) {
  const int V = a.B * a.C;
  int launches = 0;
  ZeroArgs z;
  const long long nv = (long long)a.B * a.N * 3;
  z.p[0] = a.vpos_grad; z.n[0] = nv;
  const long long Bs = a.sharedBatch ? 1 : a.B;        // batch extent of the colour / texture / SH gradients
  z.p[1] = a.vcol_grad; z.n[1] = Bs * a.N * 3;
  z.p[2] = a.s.gnorm;   z.n[2] = (long long)a.B * a.N * 4;
  z.p[3] = a.sh_grad;   z.n[3] = Bs * a.C * 27;
  z.p[4] = a.tex_grad;  z.n[4] = a.tex_grad ? Bs * a.texH * a.texW * 3 : 0;
  PrepArgs pa;
  pa.z = z;
  pa.vertex_pos = a.vertex_pos; pa.vertex_color = a.vertex_color; pa.vertex_normal = a.vertex_normal;
  pa.pos4 = a.s.bpos4; pa.col4 = a.s.bcol4; pa.nor4 = a.s.bnor4;
  pa.nBN = (long long)a.B * a.N; pa.nVN = (long long)V * a.N;
  pa.extr = a.extrinsics; pa.intr = a.intrinsics; pa.cams = a.s.cams; pa.V = V;
  // persistent kernel: whole-tile images only (the bulk copies need 16-byte aligned 128-byte rows)
  const bool persistent = a.bwdPersistent && (a.W % 32) == 0 && (a.H % 32) == 0 && (reinterpret_cast<uintptr_t>(a.face) & 15) == 0;
  const int tilesX = a.W / 32, tilesPerView = tilesX * (a.H / 32), totalTiles = tilesPerView * V;
  const int pgCtas = std::min(totalTiles, std::max(1, a.ctaSlots / 4 * 3));     // 3 resident CTAs per SM
  pa.tileCounter = a.s.tileCounter; pa.counterInit = pgCtas;
  tm->begin(K_ZERO, st);
  launch_chained(a.chain, prep_kernel, dim3(148 * 8), dim3(256), 0, st, pa);
  tm->end(st);
  ++launches;
  PixelParams p;
  p.render_grad = a.render_grad; p.target_grad = a.target_grad; p.vertex_pos = a.vertex_pos; p.vertex_color = a.vertex_color;
  p.texture = a.texture; p.sh_coeff = a.sh_coeff; p.target_image = a.target_image; p.vertex_normal = a.vertex_normal;
  p.bary = a.bary; p.texcoords = a.texcoords; p.target_du = a.target_du; p.target_dv = a.target_dv; p.face = a.face; p.faces4 = a.faces4; p.cams = a.s.cams;
  p.pos4 = a.s.bpos4; p.col4 = a.s.bcol4; p.nor4 = a.s.bnor4;
  p.vpos_grad = a.vpos_grad; p.vcol_grad = a.vcol_grad; p.tex_grad = a.tex_grad; p.sh_grad = a.sh_grad; p.gnorm = a.s.gnorm;
  p.C = a.C; p.N = a.N; p.W = a.W; p.H = a.H; p.texH = a.texH; p.texW = a.texW;
  p.albedo = a.albedo; p.shading = a.shading; p.imgFilter = a.imgFilter; p.texBilinear = a.texBilinear; p.sharedBatch = a.sharedBatch; p.invC = 1.f / (float)a.C;
  tm->begin(K_PIXEL_GRAD, st);
  constexpr int kPixelSmem = 8 * kWarpBufFloats * (int)sizeof(float);
  const dim3 pgGrid((a.W + 31) / 32, (a.H + 31) / 32, V);
#define GVV_PG(S, A) do { static unsigned long long attr = 0, attrP = 0; \
    if (persistent) { \
      constexpr int smemP = kPixelSmem + kFaceRing * 4096; \
      if (first_use_on_device(&attrP)) cudaFuncSetAttribute(pixel_grad_persistent_kernel<S, A>, cudaFuncAttributeMaxDynamicSharedMemorySize, smemP); \
      launch_chained(a.chain, pixel_grad_persistent_kernel<S, A>, dim3(pgCtas), dim3(256), smemP, st, p, a.s.tileCounter, tilesX, tilesPerView, totalTiles); \
    } else { \
      if (first_use_on_device(&attr)) cudaFuncSetAttribute(pixel_grad_kernel<S, A>, cudaFuncAttributeMaxDynamicSharedMemorySize, kPixelSmem); \
      launch_chained(a.chain, pixel_grad_kernel<S, A>, pgGrid, dim3(256), kPixelSmem, st, p); } } while (0)
  const bool sh = a.shading == GVV_SHADING_SHADED;
  if (a.albedo == GVV_ALBEDO_VERTEX_COLOR) { if (sh) GVV_PG(true, GVV_ALBEDO_VERTEX_COLOR); else GVV_PG(false, GVV_ALBEDO_VERTEX_COLOR); }
  else if (a.albedo == GVV_ALBEDO_TEXTURED) { if (sh) GVV_PG(true, GVV_ALBEDO_TEXTURED); else GVV_PG(false, GVV_ALBEDO_TEXTURED); }
  else { if (sh) GVV_PG(true, GVV_ALBEDO_FOREGROUND_MASK); else GVV_PG(false, GVV_ALBEDO_FOREGROUND_MASK); }   // foregroundMask (and any mode without an albedo gradient)
#undef GVV_PG
  tm->end(st);
  ++launches;
  const bool fusedAR = a.ar.blocks > 0 && !a.arAfter;
  if (a.shading == GVV_SHADING_SHADED || fusedAR) {
    ARParams ar = a.ar;
    if (!fusedAR) ar.blocks = 0;
    const int Fn = a.shading == GVV_SHADING_SHADED ? a.F : 0;         // shadeless: the kernel only carries the collective
    tm->begin(K_NORMAL_TERM, st);
    launch_chained(a.chain, normal_term_kernel, dim3((Fn + 255) / 256 + ar.blocks, Fn ? a.B : 1), dim3(256), 0, st, a.s.bpos4, reinterpret_cast<const float4*>(a.s.gnorm), a.faces4,
                   a.vpos_grad, a.N, Fn, ar);
    tm->end(st);
    ++launches;
  }
  if (a.ar.blocks > 0 && a.arAfter) {
    tm->begin(K_ALLREDUCE, st);
    launch_chained(a.chain, allreduce_kernel, dim3(a.ar.blocks), dim3(256), 0, st, a.ar);
    tm->end(st);
    ++launches;
  }
  return cudaGetLastError() == cudaSuccess ? launches : -1;
}

}  // namespace gvv
