// sm_100a kernels of the CRF inference path.  Compiled with -fmad=false: every float/double
// operation below rounds separately unless it is an explicit __fmaf_rn, which is what the
// canonical arithmetic requires.
//
// No tensor cores on purpose: nothing on this path is a dense contraction that tolerates reduced
// precision (the Gabor bank must stay f32-exact in raster order; the forests are integer gathers).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/crf_b200.h"
#include "device_forest.h"

namespace crf {

// Per-face work descriptor, filled on the host (src/FaceForest.cpp:199-204 size arithmetic).
struct FaceDesc {
  unsigned long long img_off;  // byte offset of the BGR frame inside the device image buffer
  unsigned long long img_step; // bytes per frame row
  int bx, by, bw, bh;          // face box inside the frame
  int W, H;                    // size after cv::resize (W in {124, 125})
  float scale;                 // face_size / bw (f32)
  int pad;
  double scale_x, scale_y;     // 1 / ((double)W / bw), 1 / ((double)H / bh)   (cv::resize)
};

enum { CNT_HP_TESTS = 0, CNT_FFD_TESTS, CNT_HP_TRAV, CNT_FFD_TRAV, CNT_VOTES, CNT_VOTE_PASSES, CNT_NUM };

__device__ __forceinline__ int border101(int p, int len) {  // BORDER_REFLECT_101
  if (len == 1) return 0;
  while (p < 0 || p >= len) p = p < 0 ? -p : 2 * len - 2 - p;
  return p;
}

// ---------------------------------------------------------------------------------------------
// a2 + a3: cvtColor(BGR2GRAY) on the ROI and cv::resize(INTER_LINEAR) to W x H
// (src/FaceForest.cpp:196-204).  One thread per destination pixel.
// scaled: [face][Hcap][128] u8.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int gray_at(const uint8_t* __restrict__ roi, unsigned long long step, int y, int x) {
  const uint8_t* p = roi + (size_t)y * step + (size_t)x * 3;
  return (p[2] * 9798 + p[1] * 19235 + p[0] * 3735 + 16384) >> 15;
}

__global__ void __launch_bounds__(128) k_gray_resize(const FaceDesc* __restrict__ fd, const uint8_t* __restrict__ imgs,
                                                     uint8_t* __restrict__ scaled, size_t scaled_face_stride) {
  const FaceDesc d = fd[blockIdx.y];
  const int dy = blockIdx.x, dx = threadIdx.x;
  if (dy >= d.H) return;
  uint8_t* out = scaled + blockIdx.y * scaled_face_stride + (size_t)dy * 128;
  if (dx >= d.W) { out[dx] = 0; return; }
  const uint8_t* roi = imgs + d.img_off + (size_t)d.by * d.img_step + (size_t)d.bx * 3;
  float fx = (float)((dx + 0.5) * d.scale_x - 0.5);
  int sx = (int)floorf(fx);
  fx -= sx;
  if (sx < 0) { fx = 0; sx = 0; }
  if (sx >= d.bw - 1) { fx = 0; sx = d.bw - 1; }
  const int a0 = __float2int_rn((1.f - fx) * 2048.f), a1 = __float2int_rn(fx * 2048.f);
  float fy = (float)((dy + 0.5) * d.scale_y - 0.5);
  int sy = (int)floorf(fy);
  fy -= sy;
  const int b0 = __float2int_rn((1.f - fy) * 2048.f), b1 = __float2int_rn(fy * 2048.f);
  const int sy0 = min(max(sy, 0), d.bh - 1), sy1 = min(max(sy + 1, 0), d.bh - 1);
  const int sx1 = min(sx + 1, d.bw - 1);
  const int row0 = gray_at(roi, d.img_step, sy0, sx) * a0 + gray_at(roi, d.img_step, sy0, sx1) * a1;
  const int row1 = gray_at(roi, d.img_step, sy1, sx) * a0 + gray_at(roi, d.img_step, sy1, sx1) * a1;
  const int v = (((b0 * (row0 >> 4)) >> 16) + ((b1 * (row1 >> 4)) >> 16) + 2) >> 2;
  out[dx] = (uint8_t)min(max(v, 0), 255);
}

// ---------------------------------------------------------------------------------------------
// cv::integral (FeatureChannelFactory.hpp:51 etc.) of one W x H plane whose pixels
// come from `pix(r, c)`.  128 threads.  Column sums run down the rows in registers (no
// communication), then each 32-row band is scanned horizontally by warps and written with
// fully coalesced rows.  Exact u32 sums (the reference's f32 integrals hold the same integers < 2^24), stored as
// stack_t (device_forest.h).  out: (H+1) rows x kRowStride elements.  u8out (optional): dense H x W copy of the
// 8-bit plane; out32 (optional): the full 32-bit integral, same pitch (stage API / parity tests).
// ---------------------------------------------------------------------------------------------
// NT = threads of the CTA (a multiple of 128): the first 128 own the columns, every warp takes part in the row scans.
template <int NT, class PixFn>
__device__ __forceinline__ void integral_plane_in(uint32_t (*band)[kRowStride] /* shared memory, 32 rows */, PixFn pix, int W, int H, stack_t* __restrict__ out,
                                                  uint8_t* __restrict__ u8out, uint32_t* __restrict__ out32) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool col = tid < kRowStride;
  if (col) {
    out[tid] = 0;  // row 0
    if (out32) out32[tid] = 0;
  }
  uint32_t run = 0;
  for (int r0 = 0; r0 < H; r0 += 32) {
    const int nr = min(32, H - r0);
    if (col)
    for (int r = 0; r < nr; r++) {
      uint32_t p = 0;
      if (tid < W) {
        p = pix(r0 + r, tid);
        if (u8out) u8out[(size_t)(r0 + r) * W + tid] = (uint8_t)p;
      }
      run += p;
      band[r][tid] = run;
    }
    __syncthreads();
    for (int r = warp; r < nr; r += NT / 32) {
      uint4 v = *reinterpret_cast<uint4*>(&band[r][lane * 4]);
      v.y += v.x; v.z += v.y; v.w += v.z;
      uint32_t incl = v.w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        uint32_t n = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += n;
      }
      const uint32_t excl = incl - v.w;
      v.x += excl; v.y += excl; v.z += excl; v.w += excl;
      *reinterpret_cast<uint4*>(&band[r][lane * 4]) = v;
    }
    __syncthreads();
    for (int i = tid; i < nr * kRowStride; i += NT) {  // I[y+1][x+1] = sum; column 0 stays zero
      const int r = i / kRowStride, c = i - r * kRowStride;
      const uint32_t v = c == 0 ? 0u : band[r][c - 1];
      out[(size_t)(r0 + r + 1) * kRowStride + c] = (stack_t)v;   // truncates only in the 16-bit layout (device_forest.h)
      if (out32) out32[(size_t)(r0 + r + 1) * kRowStride + c] = v;
    }
    __syncthreads();
  }
}
template <int NT = 128, class PixFn>
__device__ __forceinline__ void integral_plane(PixFn pix, int W, int H, stack_t* __restrict__ out, uint8_t* __restrict__ u8out, uint32_t* __restrict__ out32) {
  __shared__ __align__(16) uint32_t band[32][kRowStride];
  integral_plane_in<NT>(band, pix, W, H, out, u8out, out32);
}

// a5 + a7 (+ a7b): FC_GRAY, FC_SOBEL (d/dy then d/dx, 8U-saturated), FC_MIN_MAX
// (include/FeatureChannelFactory.hpp:46-57, :120-165).  grid = (nwhich, faces), 128 threads.
// which: 0 gray, 1 Sobel dy, 2 Sobel dx, 3 erode, 4 dilate, 5 equalizeHist (FC_NORM, :58-70), 6 Canny (FC_CANNY, :166-179).
// plane[k] = output plane index of entry k.
struct PlainPlanes { int which[8]; int plane[8]; };

__global__ void __launch_bounds__(128) k_plain_channels(const FaceDesc* __restrict__ fd, const uint8_t* __restrict__ scaled, size_t scaled_face_stride,
                                                        stack_t* __restrict__ stacks, size_t stack_face_stride, size_t plane_stride,
                                                        uint8_t* __restrict__ u8planes, size_t u8_face_stride, uint32_t* __restrict__ dbg32, PlainPlanes pp) {
  const FaceDesc d = fd[blockIdx.y];
  const int W = d.W, H = d.H;
  const uint8_t* __restrict__ g = scaled + blockIdx.y * scaled_face_stride;
  const int which = pp.which[blockIdx.x], plane = pp.plane[blockIdx.x];
  stack_t* out = stacks + blockIdx.y * stack_face_stride + (size_t)plane * plane_stride;
  uint8_t* u8o = u8planes ? u8planes + blockIdx.y * u8_face_stride + (size_t)plane * W * H : nullptr;
  uint32_t* o32 = dbg32 ? dbg32 + blockIdx.y * stack_face_stride + (size_t)plane * plane_stride : nullptr;
  auto G = [&](int y, int x) -> int { return g[(size_t)y * 128 + x]; };
  __shared__ int s_hist[256];
  __shared__ uint8_t s_lut[256];
  extern __shared__ __align__(16) uint8_t s_canny[];   // which == 6 only: magnitudes u16 [(H+2)(W+2)] + edge map u8 [(H+2)(W+2)]
  const int P = W + 2;
  uint16_t* s_mag = reinterpret_cast<uint16_t*>(s_canny);
  uint8_t* s_map = s_canny + (size_t)(H + 2) * P * 2;
  if (which == 6) {
    // cv::Canny(img, out, -1, 5) (include/FeatureChannelFactory.hpp:166-179), aperture 3, L1 gradient; OpenCV's canny.cpp
    // restated: 16-bit Sobel with BORDER_REPLICATE, |dx| + |dy|, directional non-maximum suppression, then hysteresis as a
    // monotone relaxation (candidate next to an edge becomes an edge) iterated to its fixed point, which is the flood fill's.
    auto GR = [&](int y, int x) -> int { return g[(size_t)min(max(y, 0), H - 1) * 128 + min(max(x, 0), W - 1)]; };
    auto grad = [&](int y, int x, int& gx, int& gy) {
      const int a = GR(y - 1, x - 1), b = GR(y - 1, x), c2 = GR(y - 1, x + 1), d2 = GR(y, x - 1), f = GR(y, x + 1), g2 = GR(y + 1, x - 1), h = GR(y + 1, x), i2 = GR(y + 1, x + 1);
      gx = (c2 + 2 * f + i2) - (a + 2 * d2 + g2);
      gy = (g2 + 2 * h + i2) - (a + 2 * b + c2);
    };
    for (int i = threadIdx.x; i < (H + 2) * P; i += 128) { s_mag[i] = 0; s_map[i] = 1; }
    __syncthreads();
    for (int i = threadIdx.x; i < W * H; i += 128) {
      const int y = i / W, x = i - y * W;
      int gx, gy; grad(y, x, gx, gy);
      s_mag[(y + 1) * P + x + 1] = (uint16_t)(abs(gx) + abs(gy));
    }
    __syncthreads();
    for (int i = threadIdx.x; i < W * H; i += 128) {
      const int y = i / W, x = i - y * W, c = (y + 1) * P + x + 1;
      int xs, ys; grad(y, x, xs, ys);
      const int m = s_mag[c];
      const int ax = abs(xs), ay = abs(ys) << 15, tg22x = ax * 13573;   // TG22 = (int)(tan(22.5 deg) * 2^15 + 0.5)
      bool ok;
      if (ay < tg22x) ok = m > s_mag[c - 1] && m >= s_mag[c + 1];
      else if (ay > tg22x + (ax << 16)) ok = m > s_mag[c - P] && m >= s_mag[c + P];
      else { const int sg = (xs ^ ys) < 0 ? -1 : 1; ok = m > s_mag[c - P - sg] && m > s_mag[c + P + sg]; }
      // low threshold -1: every magnitude (>= 0) passes `m > low`
      if (ok) s_map[c] = m > 5 ? 2 : 0;
    }
    __syncthreads();
    for (;;) {
      int changed = 0;
      for (int i = threadIdx.x; i < W * H; i += 128) {
        const int y = i / W, x = i - y * W, c = (y + 1) * P + x + 1;
        if (s_map[c] != 0) continue;
        const bool near = s_map[c - P - 1] == 2 || s_map[c - P] == 2 || s_map[c - P + 1] == 2 || s_map[c - 1] == 2 || s_map[c + 1] == 2 ||
                          s_map[c + P - 1] == 2 || s_map[c + P] == 2 || s_map[c + P + 1] == 2;
        if (near) { s_map[c] = 2; changed = 1; }
      }
      if (!__syncthreads_or(changed)) break;
    }
  }
  if (which == 5) {
    // cv::equalizeHist: histogram in shared memory, the 256-entry LUT by one thread (sequential cumulative sum, f32 scale)
    for (int i = threadIdx.x; i < 256; i += 128) s_hist[i] = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < W * H; i += 128) atomicAdd(&s_hist[G(i / W, i % W)], 1);
    __syncthreads();
    if (threadIdx.x == 0) {
      int i = 0;
      while (i < 256 && !s_hist[i]) ++i;
      const int total = W * H;
      if (i < 256 && s_hist[i] == total) {
        for (int k = 0; k < 256; k++) s_lut[k] = (uint8_t)i;
      } else {
        for (int k = 0; k <= i && k < 256; k++) s_lut[k] = 0;
        const float scale = __fdiv_rn(255.f, (float)(total - s_hist[i]));
        int sum = 0;
        for (++i; i < 256; ++i) {
          sum += s_hist[i];
          s_lut[i] = (uint8_t)min(max(__float2int_rn(__fmul_rn((float)sum, scale)), 0), 255);
        }
      }
    }
    __syncthreads();
  }
  // one pixel function for every kind (a single instance of the integral's shared-memory band)
  integral_plane([&](int r, int c) -> uint32_t {
    switch (which) {
      case 0: return (uint32_t)G(r, c);
      case 1: case 2: {
        const int ym = border101(r - 1, H), yp = border101(r + 1, H), xm = border101(c - 1, W), xp = border101(c + 1, W);
        int v;
        if (which == 2) v = (G(ym, xp) + 2 * G(r, xp) + G(yp, xp)) - (G(ym, xm) + 2 * G(r, xm) + G(yp, xm));   // d/dx
        else v = (G(yp, xm) + 2 * G(yp, c) + G(yp, xp)) - (G(ym, xm) + 2 * G(ym, c) + G(ym, xp));              // d/dy
        return (uint32_t)min(max(v, 0), 255);
      }
      case 5: return s_lut[G(r, c)];
      case 6: return s_map[(r + 1) * P + c + 1] == 2 ? 255u : 0u;
      default: {
        int lo = 255, hi = 0;
        for (int j = -1; j <= 1; j++)
          for (int i = -1; i <= 1; i++) {
            const int yy = r + j, xx = c + i;
            if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
            const int v = G(yy, xx);
            lo = min(lo, v); hi = max(hi, v);
          }
        return (uint32_t)(which == 4 ? hi : lo);
      }
    }
  }, W, H, out, u8o, o32);
}

// Integral of caller-supplied dense u8 planes [C][H][W] (stage API: synthetic channels).
__global__ void __launch_bounds__(128) k_integral_from_u8(const uint8_t* __restrict__ planes, int W, int H,
                                                          stack_t* __restrict__ stacks, size_t stack_face_stride, size_t plane_stride) {
  const uint8_t* __restrict__ p = planes + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * W * H;
  stack_t* out = stacks + blockIdx.y * stack_face_stride + (size_t)blockIdx.x * plane_stride;
  integral_plane([&](int r, int c) -> uint32_t { return p[(size_t)r * W + c]; }, W, H, out, nullptr, nullptr);
}

// ---------------------------------------------------------------------------------------------
// a6, first half: the 35 complex Gabor responses (include/FeatureChannelFactory.hpp:253-270).
// cv::filter2D = correlation, anchor centre, REFLECT_101, u8 -> f32; canonical accumulation =
// raster order over the kernel; product and sum rounded separately (bit-identical to cv2 for the 7x7 kernels; the
// larger kernels, where cv2 itself switches to a DFT path, go through k_gabor_sep).  Then
// magnitude = sqrt(im*im + re*re).  One CTA = one 16-row band of one (face, orientation) at scale
// NU (kernel size K); each thread owns 1 x 4 output strips, the band and its halo live in shared
// memory as f32, coefficients are broadcast LDS.64.  Per-plane min/max via integer atomics
// (magnitudes are >= 0, so float order == unsigned order).
// mag: [face][35][Hcap][128] f32.  minmax: [face][35][2] u32 (initialised to {0x7f800000, 0}).
// grid = (bands, 7, faces), 256 threads.
// ---------------------------------------------------------------------------------------------
template <int K, int BAND_ = 16>
struct GaborGeom {
  static constexpr int R = K / 2;
  static constexpr int NF4 = (K + 3 + 3) / 4;      // float4 loads covering 4 + K - 1 pixels
  static constexpr int PITCH = 124 + 4 * NF4;      // floats per tile row
  static constexpr int BAND = BAND_;
  static constexpr int TH = BAND + K - 1;
};

template <int K>
__global__ void __launch_bounds__(256, 2) k_gabor_mag(const FaceDesc* __restrict__ fd, const uint8_t* __restrict__ scaled, size_t scaled_face_stride,
                                                      const float2* __restrict__ coef /* this scale: [7][K*K] (re, im) raster order */, int nu,
                                                      float* __restrict__ mag, size_t mag_face_stride, size_t mag_plane_stride,
                                                      uint32_t* __restrict__ minmax) {
  using G = GaborGeom<K>;
  constexpr int KP = K + 1;            // coefficient row padded to an even tap count: two taps per LDS.128
  constexpr bool FUSED = false;        // 7x7 only: product and sum rounded separately == cv2's filter2D bit for bit (larger kernels: k_gabor_sep)
  __shared__ __align__(16) float tile[G::TH][G::PITCH];
  __shared__ __align__(16) float2 cf[K][KP];
  const FaceDesc d = fd[blockIdx.z];
  const int W = d.W, H = d.H;
  const int r0 = blockIdx.x * G::BAND;
  if (r0 >= H) return;
  const int tid = threadIdx.x;
  const uint8_t* __restrict__ g = scaled + blockIdx.z * scaled_face_stride;
  for (int i = tid; i < G::TH * G::PITCH; i += 256) {
    const int ty = i / G::PITCH, tx = i - ty * G::PITCH;
    const int sy = border101(r0 + ty - G::R, H), sx = border101(tx - G::R, W);
    tile[ty][tx] = (float)g[(size_t)sy * 128 + sx];
  }
  // each thread: two 1 x 4 strips, rows rA and rA + 8 of the band, sharing every coefficient load
  const int x0 = (tid & 31) * 4;
  const int rA = tid >> 5;
  // the 7 orientations on one tile (the tile load was 20 % of the kernel's instructions when every orientation had its own CTA), or the
  // one blockIdx.y names: small batches spread the orientations over CTAs for latency
  const int mu_begin = gridDim.y == 7 ? blockIdx.y : 0, mu_end = gridDim.y == 7 ? mu_begin + 1 : 7;
#pragma unroll 1
  for (int mu = mu_begin; mu < mu_end; mu++) {
  const int plane = nu * 7 + mu;
  __syncthreads();   // the previous orientation is done with cf (and, the first time, the tile is complete after the next barrier)
  for (int i = tid; i < K * KP; i += 256) {
    const int jj = i / KP, ii = i - jj * KP;
    cf[jj][ii] = ii < K ? coef[mu * K * K + jj * K + ii] : make_float2(0.f, 0.f);
  }
  __syncthreads();
  float re[2][4], im[2][4];
#pragma unroll
  for (int o = 0; o < 4; o++) { re[0][o] = re[1][o] = 0.f; im[0][o] = im[1][o] = 0.f; }
#pragma unroll 1
  for (int j = 0; j < K; j++) {
    float px[2][4 * G::NF4];
#pragma unroll
    for (int h = 0; h < 2; h++) {
      const float4* row = reinterpret_cast<const float4*>(&tile[rA + 8 * h + j][x0]);
#pragma unroll
      for (int q = 0; q < G::NF4; q++) {
        const float4 v = row[q];
        px[h][4 * q] = v.x; px[h][4 * q + 1] = v.y; px[h][4 * q + 2] = v.z; px[h][4 * q + 3] = v.w;
      }
    }
    const float4* crow = reinterpret_cast<const float4*>(&cf[j][0]);
#pragma unroll
    for (int i2 = 0; i2 < KP / 2; i2++) {
      const float4 c2 = crow[i2];
#pragma unroll
      for (int e = 0; e < 2; e++) {
        const int i = 2 * i2 + e;
        if (i < K) {
          const float cr = e ? c2.z : c2.x, ci = e ? c2.w : c2.y;
#pragma unroll
          for (int h = 0; h < 2; h++)
#pragma unroll
            for (int o = 0; o < 4; o++) {
              if (FUSED) {
                re[h][o] = __fmaf_rn(px[h][o + i], cr, re[h][o]);
                im[h][o] = __fmaf_rn(px[h][o + i], ci, im[h][o]);
              } else {
                re[h][o] = __fadd_rn(re[h][o], __fmul_rn(px[h][o + i], cr));
                im[h][o] = __fadd_rn(im[h][o], __fmul_rn(px[h][o + i], ci));
              }
            }
        }
      }
    }
  }
  float vmin = __int_as_float(0x7f800000), vmax = 0.f;
  float* mplane = mag + blockIdx.z * mag_face_stride + (size_t)plane * mag_plane_stride;
#pragma unroll
  for (int h = 0; h < 2; h++) {
    const int r = r0 + rA + 8 * h;
    if (r < H) {
      float m[4];
#pragma unroll
      for (int o = 0; o < 4; o++) {
        m[o] = __fsqrt_rn(__fadd_rn(__fmul_rn(im[h][o], im[h][o]), __fmul_rn(re[h][o], re[h][o])));
        if (x0 + o < W) { vmin = fminf(vmin, m[o]); vmax = fmaxf(vmax, m[o]); }
      }
      *reinterpret_cast<float4*>(&mplane[(size_t)r * 128 + x0]) = make_float4(m[0], m[1], m[2], m[3]);
    }
  }
  uint32_t umin = __float_as_uint(vmin), umax = __float_as_uint(vmax);
  umin = __reduce_min_sync(0xffffffffu, umin);
  umax = __reduce_max_sync(0xffffffffu, umax);
  if ((tid & 31) == 0) {
    uint32_t* mm = minmax + ((size_t)blockIdx.z * 35 + plane) * 2;
    atomicMin(&mm[0], umin);
    atomicMax(&mm[1], umax);
  }
  }   // orientations
}

// a6 for the 9x9 .. 25x25 kernels in SEPARABLE form (the canonical arithmetic for these sizes):
//   t1(x,y) e^{i(ax x + ay y)} - dc t1(x,y) = [g(x) e^{i ax x}] [g(y) e^{i ay y}] - dc g(x) g(y)
// i.e. a complex row pass and a complex column pass per orientation plus one real Gaussian pair per scale: 6K + 2K/7
// instead of K^2 multiply-adds per pixel and orientation.  Order (all __fmaf_rn, taps ascending):
//   row pass on the REFLECT_101-padded rows: Rre, Rim (and Gr with g1);  column pass: re += Rre*hy_re; re += -Rim*hy_im;
//   im += Rre*hy_im; im += Rim*hy_re; G += Gr*g1;  then re = fmaf(-dc, G, re).
// One CTA = a 16-row band of one face, looping over the 7 orientations; the band + halo and the row-pass results
// live in (dynamic) shared memory.  coef (per scale): float2 hx[7][K], float2 hy[7][K], float g1[K], float dc.
// grid = (bands, faces), 256 threads.
template <int K, int BAND = 16>
struct GaborSepSmem {
  using G = GaborGeom<K, BAND>;
  static constexpr int NCOEF = 7 * K * 2 * 2 + K + 1;   // floats
  static constexpr size_t bytes = sizeof(float) * ((size_t)G::TH * G::PITCH + 2 * (size_t)G::TH * 128 + ((NCOEF + 3) & ~3));
};

// BAND rows per CTA (16 or 32), 16 * BAND threads: the row pass runs on BAND + K - 1 rows, so a taller band repeats fewer halo rows
// (K = 25: 2.5x -> 1.75x the row-pass work per output row).
template <int K, int BAND = 16>
__global__ void __launch_bounds__(16 * BAND, BAND == 16 ? 3 : 2) k_gabor_sep(const FaceDesc* __restrict__ fd, const uint8_t* __restrict__ scaled, size_t scaled_face_stride,
                                                      const float* __restrict__ coef, int nu, float* __restrict__ mag, size_t mag_face_stride,
                                                      size_t mag_plane_stride, uint32_t* __restrict__ minmax) {
  using G = GaborGeom<K, BAND>;
  constexpr int R = K / 2, TH = G::TH, PITCH = G::PITCH, HT = TH / 2, NT = 16 * BAND;   // TH = BAND + K - 1 is even
  static_assert(TH % 2 == 0, "row pass pairs rows r and r + TH/2");
  extern __shared__ __align__(16) float s_gs[];
  float (*tile)[PITCH] = reinterpret_cast<float (*)[PITCH]>(s_gs);
  float (*rre)[128] = reinterpret_cast<float (*)[128]>(s_gs + TH * PITCH);
  float (*rim)[128] = rre + TH;
  float* cf = s_gs + TH * PITCH + 2 * TH * 128;
  const float2* hx = reinterpret_cast<const float2*>(cf);      // [7][K]
  const float2* hy = hx + 7 * K;                               // [7][K]
  const float* g1 = cf + 7 * K * 4;                            // [K]
  const FaceDesc d = fd[blockIdx.y];
  const int W = d.W, H = d.H;
  const int r0 = blockIdx.x * G::BAND;
  if (r0 >= H) return;
  const int tid = threadIdx.x;
  const uint8_t* __restrict__ g = scaled + blockIdx.y * scaled_face_stride;
  for (int i = tid; i < GaborSepSmem<K, BAND>::NCOEF; i += NT) cf[i] = coef[i];
  for (int i = tid; i < TH * PITCH; i += NT) {
    const int ty = i / PITCH, tx = i - ty * PITCH;
    const int sy = border101(r0 + ty - R, H), sx = border101(tx - R, W);
    tile[ty][tx] = (float)g[(size_t)sy * 128 + sx];
  }
  __syncthreads();
  const float dc = cf[7 * K * 4 + K];
  // Shared-memory traffic is what bounds this kernel, so every coefficient fetched serves two rows: the row pass works on
  // rows (r, r + TH/2) together, the column pass on two vertically adjacent outputs (which also share the row-pass
  // values they read).  Column-pass thread: outputs (2q, 2q + 1) x 4 columns, q = warp.
  const int x0 = (tid & 31) * 4, q2 = (tid >> 5) * 2;

  // ---- Gaussian: row pass into rre, column pass into registers
  for (int it = tid; it < HT * 32; it += NT) {
    const int r = it >> 5, xs = (it & 31) * 4;
    float px[2][4 * G::NF4];
#pragma unroll
    for (int h = 0; h < 2; h++) {
      const float4* row = reinterpret_cast<const float4*>(&tile[r + h * HT][xs]);
#pragma unroll
      for (int q = 0; q < G::NF4; q++) { const float4 v = row[q]; px[h][4 * q] = v.x; px[h][4 * q + 1] = v.y; px[h][4 * q + 2] = v.z; px[h][4 * q + 3] = v.w; }
    }
    float a[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll
    for (int i = 0; i < K; i++) {
      const float c = g1[i];
#pragma unroll
      for (int h = 0; h < 2; h++)
#pragma unroll
        for (int o = 0; o < 4; o++) a[h][o] = __fmaf_rn(px[h][o + i], c, a[h][o]);
    }
#pragma unroll
    for (int h = 0; h < 2; h++) *reinterpret_cast<float4*>(&rre[r + h * HT][xs]) = make_float4(a[h][0], a[h][1], a[h][2], a[h][3]);
  }
  __syncthreads();
  float Gs[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
  {
    float cprev = 0.f;
#pragma unroll 2
    for (int t = 0; t <= K; t++) {   // row-pass row q2 + t feeds output q2 (tap t) and output q2 + 1 (tap t - 1)
      const float4 v = *reinterpret_cast<const float4*>(&rre[q2 + t][x0]);
      const float c = t < K ? g1[t] : 0.f;
      if (t < K) { Gs[0][0] = __fmaf_rn(v.x, c, Gs[0][0]); Gs[0][1] = __fmaf_rn(v.y, c, Gs[0][1]); Gs[0][2] = __fmaf_rn(v.z, c, Gs[0][2]); Gs[0][3] = __fmaf_rn(v.w, c, Gs[0][3]); }
      if (t > 0) { Gs[1][0] = __fmaf_rn(v.x, cprev, Gs[1][0]); Gs[1][1] = __fmaf_rn(v.y, cprev, Gs[1][1]); Gs[1][2] = __fmaf_rn(v.z, cprev, Gs[1][2]); Gs[1][3] = __fmaf_rn(v.w, cprev, Gs[1][3]); }
      cprev = c;
    }
  }
  __syncthreads();
  // ---- the 7 orientations (all of them, or the one blockIdx.z names: small batches spread the orientations over CTAs for latency,
  //      each recomputing the Gaussian pair above)
  const int mu_begin = gridDim.z == 7 ? blockIdx.z : 0, mu_end = gridDim.z == 7 ? mu_begin + 1 : 7;
#pragma unroll 1
  for (int mu = mu_begin; mu < mu_end; mu++) {
    const float2* hxm = hx + mu * K;
    const float2* hym = hy + mu * K;
#pragma unroll 1
    for (int it = tid; it < HT * 32; it += NT) {
      const int r = it >> 5, xs = (it & 31) * 4;
      float px[2][4 * G::NF4];
#pragma unroll
      for (int h = 0; h < 2; h++) {
        const float4* row = reinterpret_cast<const float4*>(&tile[r + h * HT][xs]);
#pragma unroll
        for (int q = 0; q < G::NF4; q++) { const float4 v = row[q]; px[h][4 * q] = v.x; px[h][4 * q + 1] = v.y; px[h][4 * q + 2] = v.z; px[h][4 * q + 3] = v.w; }
      }
      float a[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}}, b[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll
      for (int i = 0; i < K; i++) {
        const float2 c = hxm[i];
#pragma unroll
        for (int h = 0; h < 2; h++)
#pragma unroll
          for (int o = 0; o < 4; o++) { a[h][o] = __fmaf_rn(px[h][o + i], c.x, a[h][o]); b[h][o] = __fmaf_rn(px[h][o + i], c.y, b[h][o]); }
      }
#pragma unroll
      for (int h = 0; h < 2; h++) {
        *reinterpret_cast<float4*>(&rre[r + h * HT][xs]) = make_float4(a[h][0], a[h][1], a[h][2], a[h][3]);
        *reinterpret_cast<float4*>(&rim[r + h * HT][xs]) = make_float4(b[h][0], b[h][1], b[h][2], b[h][3]);
      }
    }
    __syncthreads();
    float re[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}}, im[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
    {
      float2 cprev = make_float2(0.f, 0.f);
#pragma unroll 2
      for (int t = 0; t <= K; t++) {
        const float4 a = *reinterpret_cast<const float4*>(&rre[q2 + t][x0]);
        const float4 b = *reinterpret_cast<const float4*>(&rim[q2 + t][x0]);
        const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
        const float2 c = t < K ? hym[t] : make_float2(0.f, 0.f);
        if (t < K) {
#pragma unroll
          for (int o = 0; o < 4; o++) {
            re[0][o] = __fmaf_rn(av[o], c.x, re[0][o]); re[0][o] = __fmaf_rn(-bv[o], c.y, re[0][o]);
            im[0][o] = __fmaf_rn(av[o], c.y, im[0][o]); im[0][o] = __fmaf_rn(bv[o], c.x, im[0][o]);
          }
        }
        if (t > 0) {
#pragma unroll
          for (int o = 0; o < 4; o++) {
            re[1][o] = __fmaf_rn(av[o], cprev.x, re[1][o]); re[1][o] = __fmaf_rn(-bv[o], cprev.y, re[1][o]);
            im[1][o] = __fmaf_rn(av[o], cprev.y, im[1][o]); im[1][o] = __fmaf_rn(bv[o], cprev.x, im[1][o]);
          }
        }
        cprev = c;
      }
    }
    float vmin = __int_as_float(0x7f800000), vmax = 0.f;
    float* mplane = mag + blockIdx.y * mag_face_stride + (size_t)(nu * 7 + mu) * mag_plane_stride;
#pragma unroll
    for (int h = 0; h < 2; h++) {
      const int r = r0 + q2 + h;
      if (r < H) {
        float m[4];
#pragma unroll
        for (int o = 0; o < 4; o++) {
          const float rr = __fmaf_rn(-dc, Gs[h][o], re[h][o]);
          m[o] = __fsqrt_rn(__fadd_rn(__fmul_rn(im[h][o], im[h][o]), __fmul_rn(rr, rr)));
          if (x0 + o < W) { vmin = fminf(vmin, m[o]); vmax = fmaxf(vmax, m[o]); }
        }
        *reinterpret_cast<float4*>(&mplane[(size_t)r * 128 + x0]) = make_float4(m[0], m[1], m[2], m[3]);
      }
    }
    const uint32_t umin = __reduce_min_sync(0xffffffffu, __float_as_uint(vmin));
    const uint32_t umax = __reduce_max_sync(0xffffffffu, __float_as_uint(vmax));
    if ((tid & 31) == 0) {
      uint32_t* mm = minmax + ((size_t)blockIdx.y * 35 + nu * 7 + mu) * 2;
      atomicMin(&mm[0], umin);
      atomicMax(&mm[1], umax);
    }
    __syncthreads();   // rre / rim are rewritten by the next orientation
  }
}

// a6, second half: cv::normalize(NORM_MINMAX, 0, 1) as one single-rounded FMA, convertTo(8U, x255)
// with round-half-even, then cv::integral (FeatureChannelFactory.hpp:271-283).
// grid = (35, faces), 128 threads.  Gabor planes are 1..35 of the stack.
__global__ void __launch_bounds__(128) k_gabor_quant_integral(const FaceDesc* __restrict__ fd, const float* __restrict__ mag, size_t mag_face_stride,
                                                              size_t mag_plane_stride, const uint32_t* __restrict__ minmax,
                                                              stack_t* __restrict__ stacks, size_t stack_face_stride, size_t plane_stride, int first_plane,
                                                              uint8_t* __restrict__ u8planes, size_t u8_face_stride, uint32_t* __restrict__ dbg32) {
  const FaceDesc d = fd[blockIdx.y];
  const int gp = blockIdx.x;
  const float* __restrict__ m = mag + blockIdx.y * mag_face_stride + (size_t)gp * mag_plane_stride;
  const uint32_t* mm = minmax + ((size_t)blockIdx.y * 35 + gp) * 2;
  const double smin = (double)__uint_as_float(mm[0]), smax = (double)__uint_as_float(mm[1]);
  const double dscale = (smax - smin) > 2.220446049250313e-16 ? 1. / (smax - smin) : 0.;
  const double dshift = 0.0 - smin * dscale;
  const float a = (float)dscale, b = (float)dshift;
  const int plane = first_plane + gp;
  stack_t* out = stacks + blockIdx.y * stack_face_stride + (size_t)plane * plane_stride;
  uint8_t* u8o = u8planes ? u8planes + blockIdx.y * u8_face_stride + (size_t)plane * d.W * d.H : nullptr;
  uint32_t* o32 = dbg32 ? dbg32 + blockIdx.y * stack_face_stride + (size_t)plane * plane_stride : nullptr;
  integral_plane([&](int r, int c) -> uint32_t {
    const float v = __fmaf_rn(m[(size_t)r * 128 + c], a, b);
    const int iv = __float2int_rn(__fmul_rn(v, 255.f));
    return (uint32_t)min(max(iv, 0), 255);
  }, d.W, d.H, out, u8o, o32);
}

// ---------------------------------------------------------------------------------------------
// a6, whole: one scale of the Gabor bank for one face per work item, fused with normalize / convertTo / integral
// (FeatureChannelFactory.hpp:253-283).  Same arithmetic, bit for bit, as k_gabor_sep / k_gabor_mag<7> + k_gabor_quant_integral
// (every output is the same sequence of fmaf over the same operands), re-organised around three facts:
//   * the row pass of a padded row does not depend on the band that needs it: bands of 16 output rows are walked top to bottom and
//     only the 16 new rows are filtered, the K - 1 rows a band shares with the previous one are kept (moved to the top of the ring).
//     Row-pass work drops from 8 x (16 + K - 1) to H + K - 1 rows per plane (-27 % of all multiply-adds of the separable scales);
//   * a magnitude plane is consumed by its own CTA as soon as its minimum and maximum are known, so it never needs to reach DRAM:
//     it is written to a per-CTA scratch slot (2 planes per resident CTA, ~57 MB for the whole GPU, L2-resident and overwritten
//     plane after plane) and read back for the quantisation, instead of an 18 GB round trip through a per-face scratch;
//   * min / max are CTA-local: no atomics, no initialisation kernel.
// Persistent CTAs fetch (face) items from a global counter.  grid = min(items, SMs x 3), 256 threads.
// scratch: [gridDim.x][2][Hcap][128] f32 (plane 0: Gaussian term of the scale, plane 1: magnitudes of the current orientation).
// ---------------------------------------------------------------------------------------------
struct GaborFusedArgs {
  const FaceDesc* fd; int nfaces;
  const uint8_t* scaled; size_t scaled_face_stride;
  float* scratch; size_t scratch_plane_stride;   // Hcap * 128
  stack_t* stacks; size_t stack_face_stride, plane_stride; int first_plane;
  uint8_t* u8planes; size_t u8_face_stride; uint32_t* dbg32;
  int* counter;
};

// min / max of the CTA's valid magnitudes -> the scale and shift of cv::normalize(NORM_MINMAX, 0, 1) (as k_gabor_quant_integral)
__device__ __forceinline__ void gabor_minmax_to_affine(float vmin, float vmax, float& a, float& b) {
  __shared__ uint32_t s_mn[8], s_mx[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t umin = __reduce_min_sync(0xffffffffu, __float_as_uint(vmin)), umax = __reduce_max_sync(0xffffffffu, __float_as_uint(vmax));
  __syncthreads();   // the previous plane's readers are done with s_mn / s_mx
  if (lane == 0) { s_mn[warp] = umin; s_mx[warp] = umax; }
  __syncthreads();
  uint32_t mn = s_mn[0], mx = s_mx[0];
#pragma unroll
  for (int j = 1; j < 8; j++) { mn = min(mn, s_mn[j]); mx = max(mx, s_mx[j]); }
  const double smin = (double)__uint_as_float(mn), smax = (double)__uint_as_float(mx);
  const double dscale = (smax - smin) > 2.220446049250313e-16 ? 1. / (smax - smin) : 0.;
  const double dshift = 0.0 - smin * dscale;
  a = (float)dscale; b = (float)dshift;
}

// cv::normalize + convertTo(8U, x255) + cv::integral of a magnitude plane held in the CTA's scratch slot, by all 256 threads:
// a 32-row band is quantised cooperatively (independent float4 loads: the scratch sits in L2, and a per-column loop over dependent
// loads would serialise on its latency), then 128 threads run the column sums out of shared memory, the warps scan the rows, and
// the band is written with coalesced rows.  band: 16 KB of shared memory the caller does not need meanwhile (its filter ring).
__device__ __forceinline__ void gabor_quantise_plane(const GaborFusedArgs& g, int f, const FaceDesc& d, int gp, const float* __restrict__ mplane, float a, float b,
                                                     uint32_t (*band)[kRowStride]) {
  const int plane = g.first_plane + gp;
  stack_t* __restrict__ out = g.stacks + f * g.stack_face_stride + (size_t)plane * g.plane_stride;
  uint8_t* __restrict__ u8o = g.u8planes ? g.u8planes + f * g.u8_face_stride + (size_t)plane * d.W * d.H : nullptr;
  uint32_t* __restrict__ o32 = g.dbg32 ? g.dbg32 + f * g.stack_face_stride + (size_t)plane * g.plane_stride : nullptr;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, W = d.W, H = d.H;
  if (tid < kRowStride) { out[tid] = 0; if (o32) o32[tid] = 0; }   // row 0
  uint32_t run = 0;
  for (int r0 = 0; r0 < H; r0 += 32) {
    const int nr = min(32, H - r0);
    for (int i = tid; i < nr * 32; i += 256) {
      const int r = i >> 5, c = (i & 31) * 4;
      const float4 m = *reinterpret_cast<const float4*>(&mplane[(size_t)(r0 + r) * 128 + c]);   // this CTA's own stores: coherent load
      const float mv[4] = {m.x, m.y, m.z, m.w};
      uint32_t q[4];
#pragma unroll
      for (int o = 0; o < 4; o++) {
        const int iv = __float2int_rn(__fmul_rn(__fmaf_rn(mv[o], a, b), 255.f));
        q[o] = c + o < W ? (uint32_t)min(max(iv, 0), 255) : 0u;
        if (u8o && c + o < W) u8o[(size_t)(r0 + r) * W + c + o] = (uint8_t)q[o];
      }
      *reinterpret_cast<uint4*>(&band[r][c]) = make_uint4(q[0], q[1], q[2], q[3]);
    }
    __syncthreads();
    if (tid < kRowStride) {
#pragma unroll 8
      for (int r = 0; r < nr; r++) { run += band[r][tid]; band[r][tid] = run; }
    }
    __syncthreads();
    for (int r = warp; r < nr; r += 8) {
      uint4 v = *reinterpret_cast<uint4*>(&band[r][lane * 4]);
      v.y += v.x; v.z += v.y; v.w += v.z;
      uint32_t incl = v.w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t n = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += n;
      }
      const uint32_t excl = incl - v.w;
      v.x += excl; v.y += excl; v.z += excl; v.w += excl;
      *reinterpret_cast<uint4*>(&band[r][lane * 4]) = v;
    }
    __syncthreads();
    for (int i = tid; i < nr * kRowStride; i += 256) {   // I[y+1][x+1] = sum; column 0 stays zero
      const int r = i / kRowStride, c = i - r * kRowStride;
      const uint32_t v = c == 0 ? 0u : band[r][c - 1];
      out[(size_t)(r0 + r + 1) * kRowStride + c] = (stack_t)v;
      if (o32) o32[(size_t)(r0 + r + 1) * kRowStride + c] = v;
    }
    __syncthreads();
  }
}

// a6, second half for the banded kernels (k_gabor_sep / k_gabor_mag<7>): the per-plane minimum and maximum they reduced with atomics
// give the affine map of cv::normalize; quantisation and integral as above.  Replaces k_gabor_quant_integral, whose one-column-per-
// thread loop over dependent magnitude loads ran at the DRAM latency (4.2 ms for 143 360 planes; this one is bandwidth-bound).
// grid = (35, faces), 256 threads.
__global__ void __launch_bounds__(256) k_gabor_quant_band(GaborFusedArgs g, const float* __restrict__ mag, size_t mag_face_stride, size_t mag_plane_stride,
                                                          const uint32_t* __restrict__ minmax) {
  __shared__ __align__(16) uint32_t s_band[32][kRowStride];
  const int f = blockIdx.y, gp = blockIdx.x;
  const FaceDesc d = g.fd[f];
  const uint32_t* mm = minmax + ((size_t)f * 35 + gp) * 2;
  const double smin = (double)__uint_as_float(mm[0]), smax = (double)__uint_as_float(mm[1]);
  const double dscale = (smax - smin) > 2.220446049250313e-16 ? 1. / (smax - smin) : 0.;
  const double dshift = 0.0 - smin * dscale;
  gabor_quantise_plane(g, f, d, gp, mag + f * mag_face_stride + (size_t)gp * mag_plane_stride, (float)dscale, (float)dshift, s_band);
}

template <int K>
__global__ void __launch_bounds__(256, 3) k_gabor_fused(GaborFusedArgs g, const float* __restrict__ coef, int nu) {
  using G = GaborGeom<K>;
  constexpr int R = K / 2, TH = G::TH, PITCH = G::PITCH, NEW = G::BAND;   // TH = 16 + K - 1 is even
  extern __shared__ __align__(16) float s_gs[];
  float (*tile)[PITCH] = reinterpret_cast<float (*)[PITCH]>(s_gs);
  float (*rre)[128] = reinterpret_cast<float (*)[128]>(s_gs + TH * PITCH);
  float (*rim)[128] = rre + TH;
  float* cf = s_gs + TH * PITCH + 2 * TH * 128;
  const float2* hx = reinterpret_cast<const float2*>(cf);      // [7][K]
  const float2* hy = hx + 7 * K;                               // [7][K]
  const float* g1 = cf + 7 * K * 4;                            // [K]
  __shared__ int s_item;
  const int tid = threadIdx.x;
  for (int i = tid; i < GaborSepSmem<K>::NCOEF; i += 256) cf[i] = coef[i];
  __syncthreads();
  const float dc = cf[7 * K * 4 + K];
  const int x0 = (tid & 31) * 4, q2 = (tid >> 5) * 2;
  float* gplane = g.scratch + (size_t)blockIdx.x * 2 * g.scratch_plane_stride;
  float* mplane = gplane + g.scratch_plane_stride;

  for (;;) {
    if (tid == 0) s_item = atomicAdd(g.counter, 1);
    __syncthreads();
    const int f = s_item;
    if (f >= g.nfaces) break;
    const FaceDesc d = g.fd[f];
    const int W = d.W, H = d.H;
    const uint8_t* __restrict__ gray = g.scaled + f * g.scaled_face_stride;
    const int nbands = (H + NEW - 1) / NEW;

    // rows [first, first + n) of the ring <- REFLECT_101-padded gray rows starting at padded row p0 (as float, with the halo columns)
    auto load_tile = [&](int first, int n, int p0) {
      for (int i = tid; i < n * PITCH; i += 256) {
        const int ty = i / PITCH, tx = i - ty * PITCH;
        tile[first + ty][tx] = (float)gray[(size_t)border101(p0 + ty, H) * 128 + border101(tx - R, W)];
      }
    };
    // ================= the Gaussian term of this scale: G plane (shared by the 7 orientations)
    for (int band = 0; band < nbands; band++) {
      const int r0 = band * NEW;
      const int first = band == 0 ? 0 : K - 1, nrows = band == 0 ? TH : NEW;
      if (band > 0) {
        // moving rows down by NEW inside one array: sources [NEW, TH) and destinations [0, K - 1) overlap when K - 1 > NEW, so the
        // move goes through registers with a barrier between all loads and all stores
        float4 keep[((K - 1) * 32 + 255) / 256];
#pragma unroll
        for (int j = 0; j < ((K - 1) * 32 + 255) / 256; j++) { const int i = tid + j * 256; if (i < (K - 1) * 32) keep[j] = *reinterpret_cast<const float4*>(&rre[NEW + (i >> 5)][(i & 31) * 4]); }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < ((K - 1) * 32 + 255) / 256; j++) { const int i = tid + j * 256; if (i < (K - 1) * 32) *reinterpret_cast<float4*>(&rre[i >> 5][(i & 31) * 4]) = keep[j]; }
      }
      load_tile(first, nrows, r0 - R + first);
      __syncthreads();
      // row pass: pairs of rows (r, r + nrows / 2) share every coefficient load
      for (int it = tid; it < (nrows / 2) * 32; it += 256) {
        const int r = first + (it >> 5), xs = (it & 31) * 4, hstep = nrows / 2;
        float px[2][4 * G::NF4];
#pragma unroll
        for (int h = 0; h < 2; h++) {
          const float4* row = reinterpret_cast<const float4*>(&tile[r + h * hstep][xs]);
#pragma unroll
          for (int q = 0; q < G::NF4; q++) { const float4 v = row[q]; px[h][4 * q] = v.x; px[h][4 * q + 1] = v.y; px[h][4 * q + 2] = v.z; px[h][4 * q + 3] = v.w; }
        }
        float a[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll
        for (int i = 0; i < K; i++) {
          const float c = g1[i];
#pragma unroll
          for (int h = 0; h < 2; h++)
#pragma unroll
            for (int o = 0; o < 4; o++) a[h][o] = __fmaf_rn(px[h][o + i], c, a[h][o]);
        }
#pragma unroll
        for (int h = 0; h < 2; h++) *reinterpret_cast<float4*>(&rre[r + h * hstep][xs]) = make_float4(a[h][0], a[h][1], a[h][2], a[h][3]);
      }
      __syncthreads();
      float Gs[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
      {
        float cprev = 0.f;
#pragma unroll 2
        for (int t = 0; t <= K; t++) {
          const float4 v = *reinterpret_cast<const float4*>(&rre[q2 + t][x0]);
          const float c = t < K ? g1[t] : 0.f;
          if (t < K) { Gs[0][0] = __fmaf_rn(v.x, c, Gs[0][0]); Gs[0][1] = __fmaf_rn(v.y, c, Gs[0][1]); Gs[0][2] = __fmaf_rn(v.z, c, Gs[0][2]); Gs[0][3] = __fmaf_rn(v.w, c, Gs[0][3]); }
          if (t > 0) { Gs[1][0] = __fmaf_rn(v.x, cprev, Gs[1][0]); Gs[1][1] = __fmaf_rn(v.y, cprev, Gs[1][1]); Gs[1][2] = __fmaf_rn(v.z, cprev, Gs[1][2]); Gs[1][3] = __fmaf_rn(v.w, cprev, Gs[1][3]); }
          cprev = c;
        }
      }
#pragma unroll
      for (int h = 0; h < 2; h++) {
        const int r = r0 + q2 + h;
        if (r < H) *reinterpret_cast<float4*>(&gplane[(size_t)r * 128 + x0]) = make_float4(Gs[h][0], Gs[h][1], Gs[h][2], Gs[h][3]);
      }
      __syncthreads();   // the column pass is done with the ring before the next band moves it
    }

    // ================= the 7 orientations
#pragma unroll 1
    for (int mu = 0; mu < 7; mu++) {
      const float2* hxm = hx + mu * K;
      const float2* hym = hy + mu * K;
      float vmin = __int_as_float(0x7f800000), vmax = 0.f;
#pragma unroll 1
      for (int band = 0; band < nbands; band++) {
        const int r0 = band * NEW;
        const int first = band == 0 ? 0 : K - 1, nrows = band == 0 ? TH : NEW;
        if (band > 0) {
          float4 ka[((K - 1) * 32 + 255) / 256], kb[((K - 1) * 32 + 255) / 256];
#pragma unroll
          for (int j = 0; j < ((K - 1) * 32 + 255) / 256; j++) {
            const int i = tid + j * 256;
            if (i < (K - 1) * 32) { ka[j] = *reinterpret_cast<const float4*>(&rre[NEW + (i >> 5)][(i & 31) * 4]); kb[j] = *reinterpret_cast<const float4*>(&rim[NEW + (i >> 5)][(i & 31) * 4]); }
          }
          __syncthreads();
#pragma unroll
          for (int j = 0; j < ((K - 1) * 32 + 255) / 256; j++) {
            const int i = tid + j * 256;
            if (i < (K - 1) * 32) { *reinterpret_cast<float4*>(&rre[i >> 5][(i & 31) * 4]) = ka[j]; *reinterpret_cast<float4*>(&rim[i >> 5][(i & 31) * 4]) = kb[j]; }
          }
        }
        load_tile(first, nrows, r0 - R + first);
        __syncthreads();
#pragma unroll 1
        for (int it = tid; it < (nrows / 2) * 32; it += 256) {
          const int r = first + (it >> 5), xs = (it & 31) * 4, hstep = nrows / 2;
          float px[2][4 * G::NF4];
#pragma unroll
          for (int h = 0; h < 2; h++) {
            const float4* row = reinterpret_cast<const float4*>(&tile[r + h * hstep][xs]);
#pragma unroll
            for (int q = 0; q < G::NF4; q++) { const float4 v = row[q]; px[h][4 * q] = v.x; px[h][4 * q + 1] = v.y; px[h][4 * q + 2] = v.z; px[h][4 * q + 3] = v.w; }
          }
          float a[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}}, b[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll
          for (int i = 0; i < K; i++) {
            const float2 c = hxm[i];
#pragma unroll
            for (int h = 0; h < 2; h++)
#pragma unroll
              for (int o = 0; o < 4; o++) { a[h][o] = __fmaf_rn(px[h][o + i], c.x, a[h][o]); b[h][o] = __fmaf_rn(px[h][o + i], c.y, b[h][o]); }
          }
#pragma unroll
          for (int h = 0; h < 2; h++) {
            *reinterpret_cast<float4*>(&rre[r + h * hstep][xs]) = make_float4(a[h][0], a[h][1], a[h][2], a[h][3]);
            *reinterpret_cast<float4*>(&rim[r + h * hstep][xs]) = make_float4(b[h][0], b[h][1], b[h][2], b[h][3]);
          }
        }
        __syncthreads();
        float re[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}}, im[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
        {
          float2 cprev = make_float2(0.f, 0.f);
#pragma unroll 2
          for (int t = 0; t <= K; t++) {
            const float4 a = *reinterpret_cast<const float4*>(&rre[q2 + t][x0]);
            const float4 b = *reinterpret_cast<const float4*>(&rim[q2 + t][x0]);
            const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
            const float2 c = t < K ? hym[t] : make_float2(0.f, 0.f);
            if (t < K) {
#pragma unroll
              for (int o = 0; o < 4; o++) {
                re[0][o] = __fmaf_rn(av[o], c.x, re[0][o]); re[0][o] = __fmaf_rn(-bv[o], c.y, re[0][o]);
                im[0][o] = __fmaf_rn(av[o], c.y, im[0][o]); im[0][o] = __fmaf_rn(bv[o], c.x, im[0][o]);
              }
            }
            if (t > 0) {
#pragma unroll
              for (int o = 0; o < 4; o++) {
                re[1][o] = __fmaf_rn(av[o], cprev.x, re[1][o]); re[1][o] = __fmaf_rn(-bv[o], cprev.y, re[1][o]);
                im[1][o] = __fmaf_rn(av[o], cprev.y, im[1][o]); im[1][o] = __fmaf_rn(bv[o], cprev.x, im[1][o]);
              }
            }
            cprev = c;
          }
        }
#pragma unroll
        for (int h = 0; h < 2; h++) {
          const int r = r0 + q2 + h;
          if (r < H) {
            const float4 gq = *reinterpret_cast<const float4*>(&gplane[(size_t)r * 128 + x0]);   // this CTA's own earlier stores
            const float gv[4] = {gq.x, gq.y, gq.z, gq.w};
            float m[4];
#pragma unroll
            for (int o = 0; o < 4; o++) {
              const float rr = __fmaf_rn(-dc, gv[o], re[h][o]);
              m[o] = __fsqrt_rn(__fadd_rn(__fmul_rn(im[h][o], im[h][o]), __fmul_rn(rr, rr)));
              if (x0 + o < W) { vmin = fminf(vmin, m[o]); vmax = fmaxf(vmax, m[o]); }
            }
            *reinterpret_cast<float4*>(&mplane[(size_t)r * 128 + x0]) = make_float4(m[0], m[1], m[2], m[3]);
          }
        }
        __syncthreads();   // ring free for the next band; the magnitudes of this band are visible to the CTA
      }
      float a, b;
      gabor_minmax_to_affine(vmin, vmax, a, b);
      static_assert(((size_t)TH * PITCH + 2 * (size_t)TH * 128) * sizeof(float) >= 32 * kRowStride * sizeof(uint32_t), "tile + ring (not the coefficients behind them) are lent to the integral scan");
      gabor_quantise_plane(g, f, d, nu * 7 + mu, mplane, a, b, reinterpret_cast<uint32_t (*)[kRowStride]>(s_gs));
    }
  }
}

// The 7 x 7 scale the same way: the direct raster sum of k_gabor_mag<7> (== cv2's filter2D), magnitudes to the CTA's scratch
// slot, CTA-local min / max, quantise + integral.  One item = one face; 256 threads, 16-row bands.
__global__ void __launch_bounds__(256, 2) k_gabor_fused7(GaborFusedArgs g, const float2* __restrict__ coef /* [7][49] (re, im) raster order */) {
  constexpr int K = 7, KP = K + 1;
  using G = GaborGeom<K>;
  __shared__ __align__(16) float tile[G::TH][G::PITCH];
  __shared__ __align__(16) float2 cf[K][KP];
  __shared__ __align__(16) uint32_t s_band[32][kRowStride];
  __shared__ int s_item;
  const int tid = threadIdx.x;
  float* mplane = g.scratch + ((size_t)blockIdx.x * 2 + 1) * g.scratch_plane_stride;
  const int x0 = (tid & 31) * 4, rA = tid >> 5;
  for (;;) {
    if (tid == 0) s_item = atomicAdd(g.counter, 1);
    __syncthreads();
    const int f = s_item;
    if (f >= g.nfaces) break;
    const FaceDesc d = g.fd[f];
    const int W = d.W, H = d.H;
    const uint8_t* __restrict__ gray = g.scaled + f * g.scaled_face_stride;
#pragma unroll 1
    for (int mu = 0; mu < 7; mu++) {
      __syncthreads();   // cf / tile of the previous plane are no longer read
      for (int i = tid; i < K * KP; i += 256) {
        const int jj = i / KP, ii = i - jj * KP;
        cf[jj][ii] = ii < K ? coef[mu * K * K + jj * K + ii] : make_float2(0.f, 0.f);
      }
      float vmin = __int_as_float(0x7f800000), vmax = 0.f;
#pragma unroll 1
      for (int r0 = 0; r0 < H; r0 += G::BAND) {
        __syncthreads();
        for (int i = tid; i < G::TH * G::PITCH; i += 256) {
          const int ty = i / G::PITCH, tx = i - ty * G::PITCH;
          tile[ty][tx] = (float)gray[(size_t)border101(r0 + ty - G::R, H) * 128 + border101(tx - G::R, W)];
        }
        __syncthreads();
        float re[2][4], im[2][4];
#pragma unroll
        for (int o = 0; o < 4; o++) { re[0][o] = re[1][o] = 0.f; im[0][o] = im[1][o] = 0.f; }
#pragma unroll 1
        for (int j = 0; j < K; j++) {
          float px[2][4 * G::NF4];
#pragma unroll
          for (int h = 0; h < 2; h++) {
            const float4* row = reinterpret_cast<const float4*>(&tile[rA + 8 * h + j][x0]);
#pragma unroll
            for (int q = 0; q < G::NF4; q++) { const float4 v = row[q]; px[h][4 * q] = v.x; px[h][4 * q + 1] = v.y; px[h][4 * q + 2] = v.z; px[h][4 * q + 3] = v.w; }
          }
          const float4* crow = reinterpret_cast<const float4*>(&cf[j][0]);
#pragma unroll
          for (int i2 = 0; i2 < KP / 2; i2++) {
            const float4 c2 = crow[i2];
#pragma unroll
            for (int e = 0; e < 2; e++) {
              const int i = 2 * i2 + e;
              if (i < K) {
                const float cr = e ? c2.z : c2.x, ci = e ? c2.w : c2.y;
#pragma unroll
                for (int h = 0; h < 2; h++)
#pragma unroll
                  for (int o = 0; o < 4; o++) {   // product and sum rounded separately == cv2's filter2D bit for bit
                    re[h][o] = __fadd_rn(re[h][o], __fmul_rn(px[h][o + i], cr));
                    im[h][o] = __fadd_rn(im[h][o], __fmul_rn(px[h][o + i], ci));
                  }
              }
            }
          }
        }
#pragma unroll
        for (int h = 0; h < 2; h++) {
          const int r = r0 + rA + 8 * h;
          if (r < H) {
            float m[4];
#pragma unroll
            for (int o = 0; o < 4; o++) {
              m[o] = __fsqrt_rn(__fadd_rn(__fmul_rn(im[h][o], im[h][o]), __fmul_rn(re[h][o], re[h][o])));
              if (x0 + o < W) { vmin = fminf(vmin, m[o]); vmax = fmaxf(vmax, m[o]); }
            }
            *reinterpret_cast<float4*>(&mplane[(size_t)r * 128 + x0]) = make_float4(m[0], m[1], m[2], m[3]);
          }
        }
      }
      float a, b;
      gabor_minmax_to_affine(vmin, vmax, a, b);   // its barriers also publish the magnitudes to the CTA
      gabor_quantise_plane(g, f, d, mu, mplane, a, b, s_band);
    }
  }
}

// Number of faces whose record asks for the wide re-run (flags bits 1-2): lets the device-resident entry point fetch 4 bytes instead of
// every record when, as always with real inputs, no face overflowed the batched capacities.
__global__ void k_count_wide(const crf_face_t* __restrict__ faces, int n, int* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const bool wide = i < n && (faces[i].flags & 6) != 0;
  const unsigned b = __ballot_sync(0xffffffffu, wide);
  if ((threadIdx.x & 31) == 0 && b) atomicAdd(out, __popc(b));
}

__global__ void k_init_minmax(uint32_t* mm, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) mm[i] = (i & 1) ? 0u : 0x7f800000u;
}

// ---------------------------------------------------------------------------------------------
// a9 + a10: Forest<S>::evaluateMT over the dense patch grid (HOT LOOPS A and B; src/face_utils.cpp:
// 198-216, :256-274; include/Tree.hpp:174-191; src/ImageSample.cpp:48-63).
// One CTA = one row of 32 x-adjacent patches of one face; its warps take the trees in turn, a lane
// is one patch.  Lanes of a warp sit on the same node near the root, so the slot fetch is a broadcast
// and each of the 8 corner loads is one coalesced 128-byte row segment; deeper down the lanes
// diverge and the loads degrade into sector gathers served by L1/L2 (the stacks of the faces in
// flight are L2-resident).
// leaf_out: [face][patch (x outer, y inner)][tree] = forest-global leaf index.
// grid = (tiles, faces); tile -> (iy, ixb).  NW warps per CTA.
// ---------------------------------------------------------------------------------------------
struct TraverseArgs {
  const FaceDesc* fd;
  const stack_t* stacks;
  size_t stack_face_stride, plane_stride;
  const DevSlot* slots;
  const DevSlot16* slots16;      // compact form of the same slots (k_traverse16)
  const DevSlotW* slotsw;        // window form of the same slots (k_traverse_win)
  cudaTextureObject_t slotsw_tex;  // the same array as a linear uint4 texture (two texels per record)
  cudaTextureObject_t slots_tex;   // `slots` as a linear uint4 texture (k_traverse with MODE bit 3)
  cudaTextureObject_t slotsn_tex;  // DevSlotN records (internal nodes only) as a linear uint4 texture (k_traverse_win2)
  const int32_t* nroot_of_slot;    // DevSlot root index -> DevSlotN index of the same root (or ~leaf)
  const int32_t* roots;        // shared tree list (head pose) or nullptr
  const int32_t* face_roots;   // [face][kMaxList] composed lists (FFD) or nullptr
  const int32_t* face_ntrees;  // [face] or nullptr
  int ntrees;                  // used when face_ntrees == nullptr
  int stride;
  int32_t* leaf_out;
  size_t leaf_face_stride;
  const float* leaf_value;       // if set, write bits(leaf_value[leaf]) instead of the leaf index (head-pose path: expected label per leaf)
  unsigned long long* counters;  // nullptr = no counting
  int cnt_tests, cnt_trav;
};

__device__ __forceinline__ uint32_t ldg_corner(const stack_t* p, bool no_alloc) {
  if (no_alloc) {
    if (kStack16) { uint16_t v; asm volatile("ld.global.nc.L1::no_allocate.u16 %0, [%1];" : "=h"(v) : "l"(p)); return v; }
    uint32_t v; asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p)); return v;
  }
  return __ldg(p);
}

// Sum of one rectangle from the modulo-2^16 integral: `ns` strips, each exact (device_forest.h).
template <bool NA>
__device__ __forceinline__ uint32_t rect_sum_strips(const stack_t* __restrict__ p, uint32_t a, uint32_t w, uint32_t ns, uint32_t hs, uint32_t hl) {
  uint32_t L = ldg_corner(p + a, NA), R = ldg_corner(p + a + w, NA);
  uint32_t s = 0;
  for (uint32_t k = 1; k < ns; k++) {
    a += hs;
    const uint32_t L2 = ldg_corner(p + a, NA), R2 = ldg_corner(p + a + w, NA);
    s += (R2 - L2 - R + L) & kSumMask;
    L = L2; R = R2;
  }
  a += hl;
  const uint32_t L2 = ldg_corner(p + a, NA), R2 = ldg_corner(p + a + w, NA);
  return s + ((R2 - L2 - R + L) & kSumMask);
}

// LW = lanes along x (32: one row of 32 x-adjacent patches; 8: an 8 x 4 block of patches).
// MODE bit 0: corner loads bypass L1 allocation; bit 1: one 256-bit load per node record instead of two 128-bit;
// bit 3: the node record comes through the texture pipe (two uint4 texel fetches).
template <int NW, bool COUNT, int LW, int MODE>
__global__ void __launch_bounds__(NW * 32) k_traverse(TraverseArgs a) {
  extern __shared__ int32_t s_leaf[];  // [32][nt]
  constexpr int LH = 32 / LW;
  const int f = blockIdx.y;
  const FaceDesc d = a.fd[f];
  const int nx = (d.W - kPatch + a.stride - 1) / a.stride, ny = (d.H - kPatch + a.stride - 1) / a.stride;
  const int nxb = (nx + LW - 1) / LW, nyb = (ny + LH - 1) / LH;
  const int tile = blockIdx.x;
  if (nx <= 0 || ny <= 0 || tile >= nxb * nyb) return;
  const int iyb = tile / nxb, ixb = tile - iyb * nxb;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int ix = ixb * LW + (lane % LW), iy = iyb * LH + (lane / LW);
  const bool active = ix < nx && iy < ny;
  const int nt = a.face_ntrees ? a.face_ntrees[f] : a.ntrees;
  const int32_t* roots = a.face_roots ? a.face_roots + (size_t)f * kMaxList : a.roots;
  const stack_t* __restrict__ origin = a.stacks + f * a.stack_face_stride + (size_t)((active ? iy : 0) * a.stride) * kRowStride + (active ? ix : 0) * a.stride;
  const DevSlot* __restrict__ slots = a.slots;
  constexpr bool NA = (MODE & 1) != 0;
  unsigned tests = 0;
  for (int t = warp; t < nt; t += NW) {
    int cur = roots[t];
    int leaf = -1;
    if (active) {
      for (;;) {
        uint4 q0, q1;
        if (MODE & 8) {   // records through the texture pipe: leaves the LSU data pipe to the corner gathers
          q0 = tex1Dfetch<uint4>(a.slots_tex, 2 * cur); q1 = tex1Dfetch<uint4>(a.slots_tex, 2 * cur + 1);
        } else if (MODE & 2) {
          asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                       : "=r"(q0.x), "=r"(q0.y), "=r"(q0.z), "=r"(q0.w), "=r"(q1.x), "=r"(q1.y), "=r"(q1.z), "=r"(q1.w) : "l"(slots + cur));
        } else {
          q0 = __ldg(reinterpret_cast<const uint4*>(slots + cur));
          q1 = __ldg(reinterpret_cast<const uint4*>(slots + cur) + 1);
        }
        // q0.x = a1 | w1<<16 | ns1<<24 ; q0.y = hs1 | hl1<<16 ; q0.z, q0.w = the same for rect2
        // q1.x = ch | leaf<<8 | thr<<16 ; q1.y = m1 ; q1.z = m2 ; q1.w = child
        if ((q1.x >> 8) & 0xff) { leaf = a.leaf_value ? __float_as_int(__ldg(a.leaf_value + q1.w)) : (int)q1.w; break; }
        const stack_t* __restrict__ p = origin + (size_t)(q1.x & 0xff) * a.plane_stride;
        const uint32_t a1 = q0.x & 0xffff, w1 = (q0.x >> 16) & 0xff, ns1 = q0.x >> 24, a2 = q0.z & 0xffff, w2 = (q0.z >> 16) & 0xff, ns2 = q0.z >> 24;
        uint32_t s1, s2;
        if (!kStack16 || (ns1 | ns2) == 1) {  // both rectangles in one strip (always, with 32-bit planes): 8 independent loads
          const uint32_t e1 = a1 + (q0.y >> 16), e2 = a2 + (q0.w >> 16);
          const uint32_t A1 = ldg_corner(p + a1, NA), B1 = ldg_corner(p + a1 + w1, NA), C1 = ldg_corner(p + e1, NA), D1 = ldg_corner(p + e1 + w1, NA);
          const uint32_t A2 = ldg_corner(p + a2, NA), B2 = ldg_corner(p + a2 + w2, NA), C2 = ldg_corner(p + e2, NA), D2 = ldg_corner(p + e2 + w2, NA);
          s1 = (D1 - B1 - C1 + A1) & kSumMask;
          s2 = (D2 - B2 - C2 + A2) & kSumMask;
        } else {
          s1 = rect_sum_strips<NA>(p, a1, w1, ns1, q0.y & 0xffff, q0.y >> 16);
          s2 = rect_sum_strips<NA>(p, a2, w2, ns2, q0.w & 0xffff, q0.w >> 16);
        }
        const int m1 = (int)__umulhi(s1 << 1, q1.y), m2 = (int)__umulhi(s2 << 1, q1.z);
        const int thr = (int)(short)(q1.x >> 16);
        cur = (int)q1.w + ((m1 - m2) > thr ? 1 : 0);  // go left iff mean1 - mean2 <= threshold
        if (COUNT) tests++;
      }
    }
    s_leaf[lane * nt + t] = leaf;
  }
  __syncthreads();
  // rows of nt contiguous ints per patch, written at the reference's index (x outer, y inner)
  int32_t* out = a.leaf_out + f * a.leaf_face_stride;
  int n_active = 0;
  for (int i = threadIdx.x; i < 32 * nt; i += NW * 32) {
    const int l = i / nt, t = i - l * nt;
    const int lx = ixb * LW + (l % LW), ly = iyb * LH + (l / LW);
    if (lx < nx && ly < ny) out[((size_t)lx * ny + ly) * nt + t] = s_leaf[i];
  }
  if (COUNT) {
    tests = __reduce_add_sync(0xffffffffu, tests);
    if (lane == 0 && tests) atomicAdd(&a.counters[a.cnt_tests], (unsigned long long)tests);
    n_active = __popc(__ballot_sync(0xffffffffu, active));
    if (threadIdx.x == 0) atomicAdd(&a.counters[a.cnt_trav], (unsigned long long)n_active * nt);
  }
}

// Forest<S>::evaluateMT for an explicit list of patch origins (the per-sample level of the reference interface:
// include/Forest.hpp:81-90 is called once per patch).  One thread per (patch, tree), global gathers on the wide records.
// patches: (x, y) of the patch's top-left corner in the scaled face; leaf_out: [patch][tree].
__global__ void __launch_bounds__(128) k_traverse_patches(TraverseArgs a, const int2* __restrict__ patches, int npatches) {
  const int nt = a.face_ntrees ? a.face_ntrees[0] : a.ntrees;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= npatches * nt) return;
  const int p = idx / nt, t = idx - p * nt;
  const int32_t* roots = a.face_roots ? a.face_roots : a.roots;
  const stack_t* __restrict__ origin = a.stacks + (size_t)patches[p].y * kRowStride + patches[p].x;
  int cur = roots[t];
  for (;;) {
    const uint4 q0 = __ldg(reinterpret_cast<const uint4*>(a.slots + cur)), q1 = __ldg(reinterpret_cast<const uint4*>(a.slots + cur) + 1);
    if ((q1.x >> 8) & 0xff) { a.leaf_out[idx] = (int)q1.w; return; }
    const stack_t* __restrict__ pl = origin + (size_t)(q1.x & 0xff) * a.plane_stride;
    const uint32_t a1 = q0.x & 0xffff, w1 = (q0.x >> 16) & 0xff, a2 = q0.z & 0xffff, w2 = (q0.z >> 16) & 0xff;
    const uint32_t s1 = rect_sum_strips<false>(pl, a1, w1, q0.x >> 24, q0.y & 0xffff, q0.y >> 16);
    const uint32_t s2 = rect_sum_strips<false>(pl, a2, w2, q0.z >> 24, q0.w & 0xffff, q0.w >> 16);
    const int m1 = (int)__umulhi(s1 << 1, q1.y), m2 = (int)__umulhi(s2 << 1, q1.z);
    cur = (int)q1.w + ((m1 - m2) > (int)(short)(q1.x >> 16) ? 1 : 0);   // go left iff mean1 - mean2 <= threshold
  }
}

// ImageSample::evalTest(SimplePatchFeature, Rect) (src/ImageSample.cpp:30-64) for a list of tests on one face's integral stack:
// tests[i] = {channel, x1, y1, w1, h1, x2, y2, w2, h2, patch_x, patch_y}; out[i] = mean(rect1) - mean(rect2) with the
// reference's truncating means.
__global__ void k_eval_tests(const stack_t* __restrict__ stack, size_t plane_stride, const int* __restrict__ tests, int n, int* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int* t = tests + (size_t)i * 11;
  const stack_t* __restrict__ pl = stack + (size_t)t[0] * plane_stride;
  int mean[2];
#pragma unroll
  for (int k = 0; k < 2; k++) {
    const int x = t[9] + t[1 + 4 * k], y = t[10] + t[2 + 4 * k], w = t[3 + 4 * k], h = t[4 + 4 * k];
    const uint32_t A = pl[(size_t)y * kRowStride + x], B = pl[(size_t)y * kRowStride + x + w], C = pl[(size_t)(y + h) * kRowStride + x], D = pl[(size_t)(y + h) * kRowStride + x + w];
    mean[k] = (int)((D - B - C + A) / (uint32_t)(w * h));   // == (int)(sum / float(w * h)) of the reference for sums < 2^24
  }
  out[i] = mean[0] - mean[1];
}

// The same test on the 8-bit planes themselves: the !m_use_integral branch of ImageSample::evalTest (src/ImageSample.cpp:40-47), cv::sum over
// the two rectangles.  planes: [C][H][W] u8, dense.  cv::sum is exact (double), the quotient by float(area) truncates to the same integer as the
// integral branch (sums < 2^24), so the integer division below is both branches' result; what differs is where the sum comes from.
__global__ void k_eval_tests_sum(const uint8_t* __restrict__ planes, int W, int H, const int* __restrict__ tests, int n, int* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int* t = tests + (size_t)i * 11;
  const uint8_t* __restrict__ pl = planes + (size_t)t[0] * W * H;
  int mean[2];
#pragma unroll
  for (int k = 0; k < 2; k++) {
    const int x = t[9] + t[1 + 4 * k], y = t[10] + t[2 + 4 * k], w = t[3 + 4 * k], h = t[4 + 4 * k];
    uint32_t s = 0;
    for (int r = 0; r < h; r++)
      for (int c = 0; c < w; c++) s += pl[(size_t)(y + r) * W + x + c];
    mean[k] = (int)(s / (uint32_t)(w * h));
  }
  out[i] = mean[0] - mean[1];
}

// Exact floor(s / area) for s <= 255 * area < 2^18 without a stored reciprocal: float estimate, then a +-1 fix-up.
__device__ __forceinline__ int mean_exact(uint32_t s, uint32_t area) {
  int q = (int)(__uint2float_rn(s) * __frcp_rn(__uint2float_rn(area)));
  const int r = (int)s - q * (int)area;
  q += (r >= (int)area) ? 1 : 0;
  q -= (r < 0) ? 1 : 0;
  return q;
}

// k_traverse on the compact 16-byte slots: half the bytes per node fetch (the 256-bit fetch of the wide slot returns
// 1 KB per warp to the register file, as much as the eight corner loads together).
template <int NW, bool COUNT>
__global__ void __launch_bounds__(NW * 32) k_traverse16(TraverseArgs a) {
  extern __shared__ int32_t s_leaf[];  // [32][nt]
  const int f = blockIdx.y;
  const FaceDesc d = a.fd[f];
  const int nx = (d.W - kPatch + a.stride - 1) / a.stride, ny = (d.H - kPatch + a.stride - 1) / a.stride;
  const int nxb = (nx + 31) >> 5;
  const int tile = blockIdx.x;
  if (nx <= 0 || ny <= 0 || tile >= nxb * ny) return;
  const int iy = tile / nxb, ixb = tile - iy * nxb;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int ix = ixb * 32 + lane;
  const bool active = ix < nx;
  const int nt = a.face_ntrees ? a.face_ntrees[f] : a.ntrees;
  const int32_t* roots = a.face_roots ? a.face_roots + (size_t)f * kMaxList : a.roots;
  const stack_t* __restrict__ origin = a.stacks + f * a.stack_face_stride + (size_t)(iy * a.stride) * kRowStride + (active ? ix : 0) * a.stride;
  const DevSlot16* __restrict__ slots = a.slots16;
  unsigned tests = 0;
  for (int t = warp; t < nt; t += NW) {
    int cur = roots[t];
    int leaf = -1;
    if (active) {
      for (;;) {
        const uint4 q = __ldg(reinterpret_cast<const uint4*>(slots + cur));   // r1, r2, child, areas
        if ((q.x >> 26) & 1u) { leaf = a.leaf_value ? __float_as_int(__ldg(a.leaf_value + q.z)) : (int)q.z; break; }
        const stack_t* __restrict__ p = origin + (size_t)((q.x >> 20) & 0x3f) * a.plane_stride;
        const uint32_t a1 = ((q.x >> 5) & 31) * kRowStride + (q.x & 31), w1 = (q.x >> 10) & 31, e1 = a1 + ((q.x >> 15) & 31) * kRowStride;
        const uint32_t a2 = ((q.y >> 5) & 31) * kRowStride + (q.y & 31), w2 = (q.y >> 10) & 31, e2 = a2 + ((q.y >> 15) & 31) * kRowStride;
        const uint32_t A1 = __ldg(p + a1), B1 = __ldg(p + a1 + w1), C1 = __ldg(p + e1), D1 = __ldg(p + e1 + w1);
        const uint32_t A2 = __ldg(p + a2), B2 = __ldg(p + a2 + w2), C2 = __ldg(p + e2), D2 = __ldg(p + e2 + w2);
        const int m1 = mean_exact(D1 - B1 - C1 + A1, q.w & 0xffff), m2 = mean_exact(D2 - B2 - C2 + A2, q.w >> 16);
        const int thr = (int)((q.y >> 20) & 0x3ff) - 256;
        cur = (int)q.z + ((m1 - m2) > thr ? 1 : 0);  // go left iff mean1 - mean2 <= threshold
        if (COUNT) tests++;
      }
    }
    s_leaf[lane * nt + t] = leaf;
  }
  __syncthreads();
  const int npatch_tile = min(32, nx - ixb * 32);
  int32_t* out = a.leaf_out + f * a.leaf_face_stride;
  for (int i = threadIdx.x; i < npatch_tile * nt; i += NW * 32) {
    const int l = i / nt, t = i - l * nt;
    out[((size_t)(ixb * 32 + l) * ny + iy) * nt + t] = s_leaf[i];
  }
  if (COUNT) {
    tests = __reduce_add_sync(0xffffffffu, tests);
    if (lane == 0 && tests) atomicAdd(&a.counters[a.cnt_tests], (unsigned long long)tests);
    if (threadIdx.x == 0) atomicAdd(&a.counters[a.cnt_trav], (unsigned long long)npatch_tile * nt);
  }
}


// ---------------------------------------------------------------------------------------------
// Dense (stride 1) traversal with the corner gathers served from SHARED memory.
// The global-gather kernels above are bound by the L1 data pipe and the L2 (16.5 G sector misses per
// 4096 faces, 5.5 kB/clk of L2 traffic, 131-137 instructions per node visit).  Here a persistent CTA
// owns a column of 8 x 8-patch tiles of one face and keeps, for ALL planes, the 38 x 38 samples the
// tile can touch (device_forest.h: DevSlotW) in 231 040 bytes of shared memory, as a ring over rows:
// the first tile of a column loads 38 rows, each further tile only the 8 new ones (coalesced 160-byte
// row pieces, 16-byte vectors).  Warp = tree, lane = an 8 x 4 block of patches, two walks per lane
// (rows ly and ly + 4) in one instruction stream; row pitch 40 words == 8 (mod 32) makes the 32 lanes
// of a converged warp hit 32 different banks.  Only the 32-byte node records still come from L1/L2.
// Used when stride == 1, the model's rectangles end at <= kWinExtent, and the batch fills the GPU.
// grid = persistent CTAs; item = (face, tile column); NW warps.
// ---------------------------------------------------------------------------------------------
constexpr int kWinSmemBytes = kWinPlaneBytes * 38;   // sized for the 38 planes of the shipped feature set
constexpr int kWinMaxPlanes = 38;
#ifndef CRF_WIN_PREFETCH
#define CRF_WIN_PREFETCH 1
#endif
constexpr bool WIN_PREFETCH = CRF_WIN_PREFETCH != 0;

// Predicated shared load: a walk parked on its leaf issues no request (its lane-private address would only add bank
// conflicts to the live lanes of the warp).
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr, uint32_t live) {
  uint32_t v;   // left undefined for a parked walk: its node test is never used
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p ld.shared.u32 %0, [%1];\n\t}" : "=r"(v) : "r"(addr), "r"(live));
  return v;
}

__device__ __forceinline__ void ldg_slotw(const DevSlotW* p, uint4& q0, uint4& q1) {
  asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(q0.x), "=r"(q0.y), "=r"(q0.z), "=r"(q0.w), "=r"(q1.x), "=r"(q1.y), "=r"(q1.z), "=r"(q1.w) : "l"(p));
}

// One node test of a walk.  col = shared address of the lane's patch column, row = byte offset of the lane's patch
// row inside the ring.  q0 = px1, px2, yh1, yh2; q1 = m1, m2, child, tw.  Returns the next slot.
__device__ __forceinline__ int win_step(uint32_t col, uint32_t row, const uint4 q0, const uint4 q1, uint32_t live) {
  constexpr uint32_t kRing = kWinPlaneBytes;
  uint32_t ra1 = row + (q0.z & 0xffffu); ra1 = min(ra1, ra1 - kRing);
  uint32_t rc1 = ra1 + (q0.z >> 16);     rc1 = min(rc1, rc1 - kRing);
  uint32_t ra2 = row + (q0.w & 0xffffu); ra2 = min(ra2, ra2 - kRing);
  uint32_t rc2 = ra2 + (q0.w >> 16);     rc2 = min(rc2, rc2 - kRing);
  const uint32_t ca1 = col + q0.x, cb1 = ca1 + ((q1.w >> 16) & 0xffu);
  const uint32_t ca2 = col + q0.y, cb2 = ca2 + ((q1.w >> 24) & 0x7fu);
  const uint32_t A1 = lds_u32(ca1 + ra1, live), B1 = lds_u32(cb1 + ra1, live), C1 = lds_u32(ca1 + rc1, live), D1 = lds_u32(cb1 + rc1, live);
  const uint32_t A2 = lds_u32(ca2 + ra2, live), B2 = lds_u32(cb2 + ra2, live), C2 = lds_u32(ca2 + rc2, live), D2 = lds_u32(cb2 + rc2, live);
  const int m1 = (int)__umulhi((D1 - B1 - C1 + A1) << 1, q1.x), m2 = (int)__umulhi((D2 - B2 - C2 + A2) << 1, q1.y);
  const int thr = (int)(short)(q1.w & 0xffffu);
  return (int)q1.z + ((m1 - m2) > thr ? 1 : 0);   // go left iff mean1 - mean2 <= threshold
}

// PAIRX (two walks per lane only): the lane's two patches are HORIZONTAL neighbours instead of rows ly / ly + 4.  Neighbours one
// pixel apart share most of their path (~12 of ~15 nodes), so the second walk's record fetch is skipped while both sit
// on the same node (the record is copied between registers): fewer requests on the texture pipe and to L2.  Lane (i = lane & 3,
// y = lane >> 2) takes patches (2i + s, y) and (2i + 1 - s, y), s = y >= 4: with the 40-word row pitch (== 8 mod 32) each walk's 32
// lanes then still fall into 32 different banks when the warp is converged.
template <int NW, int WALKS, bool COUNT, bool TEX = false, bool PAIRX = false>
__global__ void __launch_bounds__(NW * 32, 1) k_traverse_win(TraverseArgs a, int nitems, int ncols, int nplanes) {
  static_assert(WALKS == 1 || WALKS == 2, "one or two walks per lane");
  static_assert(!PAIRX || WALKS == 2, "PAIRX pairs the two walks of a lane");
  // record fetch: 256-bit global load, or (TEX, the default) two 128-bit texel fetches that return through the texture pipe and
  // leave the LSU data pipe, the kernel's limiter, to the shared-memory gathers (-5 % time; CRF_WIN_TEX=0 switches it off)
  auto fetch = [&](int slot, uint4& q0, uint4& q1) {
    if (TEX) { q0 = tex1Dfetch<uint4>(a.slotsw_tex, 2 * slot); q1 = tex1Dfetch<uint4>(a.slotsw_tex, 2 * slot + 1); }
    else ldg_slotw(a.slotsw + slot, q0, q1);
  };
  extern __shared__ __align__(16) uint8_t s_win[];
  const uint32_t s_base = (uint32_t)__cvta_generic_to_shared(s_win);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int lx = lane & 7, ly = lane >> 3;
  unsigned tests = 0;
  unsigned long long trav = 0;
  for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
    const int f = item / ncols, ixb = item - f * ncols;
    const FaceDesc d = a.fd[f];
    const int nx = d.W - kPatch, ny = d.H - kPatch;   // stride 1
    const int x0 = ixb * kWinTile;
    if (nx <= 0 || ny <= 0 || x0 >= nx) continue;
    const int nrows = d.H + 1;
    const stack_t* __restrict__ src = a.stacks + f * a.stack_face_stride + x0;
    const int nt = a.face_ntrees ? a.face_ntrees[f] : a.ntrees;
    const int32_t* roots = a.face_roots ? a.face_roots + (size_t)f * kMaxList : a.roots;
    int32_t* out = a.leaf_out + f * a.leaf_face_stride;
    const bool vx = x0 + lx < nx;
    const int ntile_y = (ny + kWinTile - 1) / kWinTile;
    for (int iyb = 0; iyb < ntile_y; iyb++) {
      const int y0 = iyb * kWinTile;
      // rows of this tile's window that are not in the ring yet
      const int rbeg = iyb == 0 ? 0 : y0 + kWinExtent, rend = min(y0 + kWinRows, nrows);
      const int nvec = max(rend - rbeg, 0) * nplanes * (kWinCols / 4);
      __syncthreads();   // every walk of the previous tile is done with the rows about to be replaced
#pragma unroll 4
      for (int i = threadIdx.x; i < nvec; i += NW * 32) {
        const int c = i % (kWinCols / 4), pr = i / (kWinCols / 4);
        const int p = pr % nplanes, r = rbeg + pr / nplanes;
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(src + (size_t)p * a.plane_stride + (size_t)r * kRowStride) + c);
        const uint32_t dst = s_base + p * kWinPlaneBytes + (r % kWinRows) * kWinRowBytes + c * 16;
        asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" :: "r"(dst), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
      }
      __syncthreads();
      if (WIN_PREFETCH) {
        // the stacks of a launch (2.4 MB per face) do not stay in L2: pull the rows of the next tile step (or the first window
        // of the next item) towards L2 while this tile's walks run, one 32-byte sector per request
        const stack_t* psrc = src;
        int pbeg = y0 + kWinTile + kWinExtent, pend = min(y0 + kWinTile + kWinRows, nrows);
        if (iyb + 1 >= ntile_y) {
          const int nitem = item + gridDim.x;
          pbeg = pend = 0;
          if (nitem < nitems) {
            const int nf = nitem / ncols;
            psrc = a.stacks + nf * a.stack_face_stride + (nitem - nf * ncols) * kWinTile;
            pend = min(kWinRows, a.fd[nf].H + 1);
          }
        }
        const int nsec = max(pend - pbeg, 0) * nplanes * (kWinCols / 8);
        for (int i = threadIdx.x; i < nsec; i += NW * 32) {
          const int c = i % (kWinCols / 8), pr = i / (kWinCols / 8);
          const int p = pr % nplanes, r = pbeg + pr / nplanes;
          asm volatile("prefetch.global.L2 [%0];" :: "l"(psrc + (size_t)p * a.plane_stride + (size_t)r * kRowStride + c * 8));
        }
      }
      const int r0 = y0 % kWinRows;
      const uint32_t col = s_base + lx * 4;
      if (WALKS == 2 && PAIRX) {
        const int py = lane >> 2, sft = py >> 2, xa = 2 * (lane & 3) + sft, xb = 2 * (lane & 3) + 1 - sft;
        int ra = r0 + py;
        ra -= ra >= kWinRows ? kWinRows : 0;
        const uint32_t rowA = ra * kWinRowBytes;
        const uint32_t colA = s_base + xa * 4, colB = s_base + xb * 4;
        const bool vy = y0 + py < ny, vA = vy && x0 + xa < nx, vB = vy && x0 + xb < nx;
        for (int t = warp; t < nt; t += NW) {
          uint4 a0, a1, b0, b1;
          fetch(roots[t], a0, a1);
          b0 = a0; b1 = a1;
          for (;;) {
            const bool la = (int)a1.w < 0, lb = (int)b1.w < 0;
            if (la && lb) break;
            const int ca = win_step(colA, rowA, a0, a1, la ? 0u : 1u), cb = win_step(colB, rowA, b0, b1, lb ? 0u : 1u);
            if (COUNT) tests += (vA && !la ? 1 : 0) + (vB && !lb ? 1 : 0);
            // both walks moved to the same node: one fetch serves both.  The second walk's own fetch (lanes that split) is issued
            // BEFORE the register copy, which has to wait for the first fetch to land
            const bool same = !la && !lb && cb == ca;
            if (!la) fetch(ca, a0, a1);
            if (!lb && !same) fetch(cb, b0, b1);
            if (same) { b0 = a0; b1 = a1; }
          }
          int va = (int)a1.y, vb = (int)b1.y;
          if (a.leaf_value) { va = __float_as_int(__ldg(a.leaf_value + va)); vb = __float_as_int(__ldg(a.leaf_value + vb)); }
          if (vA) out[((size_t)(x0 + xa) * ny + (y0 + py)) * nt + t] = va;
          if (vB) out[((size_t)(x0 + xb) * ny + (y0 + py)) * nt + t] = vb;
        }
      } else if (WALKS == 2) {
        int ra = r0 + ly, rb = r0 + ly + 4;
        ra -= ra >= kWinRows ? kWinRows : 0;
        rb -= rb >= kWinRows ? kWinRows : 0;
        const uint32_t rowA = ra * kWinRowBytes, rowB = rb * kWinRowBytes;
        const bool vA = vx && y0 + ly < ny, vB = vx && y0 + ly + 4 < ny;
        for (int t = warp; t < nt; t += NW) {
          uint4 a0, a1, b0, b1;
          fetch(roots[t], a0, a1);
          b0 = a0; b1 = a1;
          for (;;) {
            const bool la = (int)a1.w < 0, lb = (int)b1.w < 0;
            if (la && lb) break;
            const int ca = win_step(col, rowA, a0, a1, la ? 0u : 1u), cb = win_step(col, rowB, b0, b1, lb ? 0u : 1u);
            if (COUNT) tests += (vA && !la ? 1 : 0) + (vB && !lb ? 1 : 0);
            if (!la) fetch(ca, a0, a1);
            if (!lb) fetch(cb, b0, b1);
          }
          int va = (int)a1.y, vb = (int)b1.y;
          if (a.leaf_value) { va = __float_as_int(__ldg(a.leaf_value + va)); vb = __float_as_int(__ldg(a.leaf_value + vb)); }
          if (vA) out[((size_t)(x0 + lx) * ny + (y0 + ly)) * nt + t] = va;
          if (vB) out[((size_t)(x0 + lx) * ny + (y0 + ly + 4)) * nt + t] = vb;
        }
      } else {
        // one walk per lane: a task is (tree, upper / lower half of the tile), so twice as many independent warps
        for (int k = warp; k < 2 * nt; k += NW) {
          const int t = k >> 1, py = ly + 4 * (k & 1);
          int ra = r0 + py;
          ra -= ra >= kWinRows ? kWinRows : 0;
          const uint32_t rowA = ra * kWinRowBytes;
          const bool vA = vx && y0 + py < ny;
          uint4 a0, a1;
          fetch(roots[t], a0, a1);
          while ((int)a1.w >= 0) {
            const int ca = win_step(col, rowA, a0, a1, 1u);
            if (COUNT) tests += vA ? 1 : 0;
            fetch(ca, a0, a1);
          }
          int va = (int)a1.y;
          if (a.leaf_value) va = __float_as_int(__ldg(a.leaf_value + va));
          if (vA) out[((size_t)(x0 + lx) * ny + (y0 + py)) * nt + t] = va;
        }
      }
      if (COUNT && threadIdx.x == 0) trav += (unsigned long long)min(kWinTile, nx - x0) * min(kWinTile, ny - y0) * nt;
    }
  }
  if (COUNT) {
    tests = __reduce_add_sync(0xffffffffu, tests);
    if (lane == 0 && tests) atomicAdd(&a.counters[a.cnt_tests], (unsigned long long)tests);
    if (threadIdx.x == 0 && trav) atomicAdd(&a.counters[a.cnt_trav], trav);
  }
}

// ---------------------------------------------------------------------------------------------
// k_traverse_win2: the same window, ring and item order as k_traverse_win, on the internal-nodes-only records (DevSlotN,
// device_forest.h).  What changes is the walk loop:
//  * a walk's position is one int: >= 0 a record index, < 0 ~leaf.  Whether the next node is a leaf is known from the parent's
//    record, so no leaf record is ever fetched (one fetch less per walk, half the record array in L1/L2) ...
//  * ... and the loop condition no longer depends on a fetch in flight.  With two walks per lane the loop body is
//    [test A, fetch A', test B, fetch B']: A's fetch has the ~60 instructions of B's test to land in, and vice versa, where
//    k_traverse_win issued both fetches at the end of the body and then waited for both (long-scoreboard stalls, 39-45 % of its
//    warp samples).
//  * lanes whose patch lies outside the grid start parked (they used to walk on whatever the window held).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int win2_step(uint32_t col, uint32_t row, const uint4 q0, const uint4 q1, uint32_t live) {
  constexpr uint32_t kRing = kWinPlaneBytes;
  uint32_t ra1 = row + (q0.z & 0xffffu); ra1 = min(ra1, ra1 - kRing);
  uint32_t rc1 = ra1 + (q0.z >> 16);     rc1 = min(rc1, rc1 - kRing);
  uint32_t ra2 = row + (q0.w & 0xffffu); ra2 = min(ra2, ra2 - kRing);
  uint32_t rc2 = ra2 + (q0.w >> 16);     rc2 = min(rc2, rc2 - kRing);
  const uint32_t ca1 = col + (q0.x & 0x3ffffu), cb1 = ca1 + (q0.x >> 18);
  const uint32_t ca2 = col + (q0.y & 0x3ffffu), cb2 = ca2 + (q0.y >> 18);
  const uint32_t A1 = lds_u32(ca1 + ra1, live), B1 = lds_u32(cb1 + ra1, live), C1 = lds_u32(ca1 + rc1, live), D1 = lds_u32(cb1 + rc1, live);
  const uint32_t A2 = lds_u32(ca2 + ra2, live), B2 = lds_u32(cb2 + ra2, live), C2 = lds_u32(ca2 + rc2, live), D2 = lds_u32(cb2 + rc2, live);
  const int m1 = (int)__umulhi((D1 - B1 - C1 + A1) << 1, q1.x), m2 = (int)__umulhi((D2 - B2 - C2 + A2) << 1, q1.y);
  const int thr = (int)q1.z >> kSlotNChildBits;
  const int left = ((int)q1.z << (32 - kSlotNChildBits)) >> (32 - kSlotNChildBits);
  return (m1 - m2) > thr ? (int)q1.w : left;   // go left iff mean1 - mean2 <= threshold
}

// Record fetch of a walk: two texel fetches, UNCONDITIONAL (a parked walk re-reads record 0, a broadcast hit) and volatile, so that
// (a) ptxas counts the fetches in flight on its scoreboard and waits for exactly the older walk's pair while the younger one's stays
// in flight, and (b) the order [gathers of A, fetch A', gathers of B, fetch B'] survives instruction scheduling.
__device__ __forceinline__ void fetch_slotn(cudaTextureObject_t tex, int slot, uint4& q0, uint4& q1) {
  const int s2 = 2 * max(slot, 0);
  asm volatile("tex.1d.v4.u32.s32 {%0,%1,%2,%3}, [%4, {%5}];" : "=r"(q0.x), "=r"(q0.y), "=r"(q0.z), "=r"(q0.w) : "l"(tex), "r"(s2));
  asm volatile("tex.1d.v4.u32.s32 {%0,%1,%2,%3}, [%4, {%5}];" : "=r"(q1.x), "=r"(q1.y), "=r"(q1.z), "=r"(q1.w) : "l"(tex), "r"(s2 + 1));
}

template <int NW, int WALKS, bool COUNT>
__global__ void __launch_bounds__(NW * 32, 1) k_traverse_win2(TraverseArgs a, int nitems, int ncols, int nplanes) {
  static_assert(WALKS == 1 || WALKS == 2, "one or two walks per lane");
  auto fetch = [&](int slot, uint4& q0, uint4& q1) { fetch_slotn(a.slotsn_tex, slot, q0, q1); };
  extern __shared__ __align__(16) uint8_t s_win[];
  const uint32_t s_base = (uint32_t)__cvta_generic_to_shared(s_win);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int lx = lane & 7, ly = lane >> 3;
  unsigned tests = 0;
  unsigned long long trav = 0;
  for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
    const int f = item / ncols, ixb = item - f * ncols;
    const FaceDesc d = a.fd[f];
    const int nx = d.W - kPatch, ny = d.H - kPatch;   // stride 1
    const int x0 = ixb * kWinTile;
    if (nx <= 0 || ny <= 0 || x0 >= nx) continue;
    const int nrows = d.H + 1;
    const stack_t* __restrict__ src = a.stacks + f * a.stack_face_stride + x0;
    const int nt = a.face_ntrees ? a.face_ntrees[f] : a.ntrees;
    const int32_t* roots = a.face_roots ? a.face_roots + (size_t)f * kMaxList : a.roots;
    int32_t* out = a.leaf_out + f * a.leaf_face_stride;
    const bool vx = x0 + lx < nx;
    const int ntile_y = (ny + kWinTile - 1) / kWinTile;
    // the warp's first tree is the same for every tile of the item: its root index (two dependent global loads) is looked up once here,
    // not at the start of every tile, where all warps would wait for it together
    const int task0 = WALKS == 2 ? warp : (warp >> 1);
    const int root0 = task0 < nt ? __ldg(a.nroot_of_slot + roots[task0]) : -1;
    // ... and so is that tree's root RECORD: kept in registers for the item, so the first test of every tile starts without a fetch
    uint4 rr0, rr1;
    fetch(root0, rr0, rr1);
    for (int iyb = 0; iyb < ntile_y; iyb++) {
      const int y0 = iyb * kWinTile;
      const int rbeg = iyb == 0 ? 0 : y0 + kWinExtent, rend = min(y0 + kWinRows, nrows);
      const int npiece = max(rend - rbeg, 0) * nplanes;   // 160-byte row pieces: (row, plane) pairs, plane fastest
      __syncthreads();   // every walk of the previous tile is done with the rows about to be replaced
      // Ring fill with cp.async (LDGSTS): ten threads per row piece (one 16-byte vector each), groups striding over the pieces with
      // incremental (plane, row) arithmetic — the first version of this loop (two run-time integer divisions per vector, LDG -> STS through
      // registers, four loads in flight per thread) took 14-17 % of the kernel (measured by running it twice per tile step).
      {
        constexpr int kVecPerPiece = kWinCols / 4, kGroups = (NW * 32) / kVecPerPiece;
        const int grp = threadIdx.x / kVecPerPiece, cvec = threadIdx.x - grp * kVecPerPiece;
        if (grp < kGroups) {
          int p = grp % nplanes, r = rbeg + grp / nplanes;
          const int dp = kGroups % nplanes, dr = kGroups / nplanes;
          for (int piece = grp; piece < npiece; piece += kGroups) {
            const stack_t* g = src + (size_t)p * a.plane_stride + (size_t)r * kRowStride + cvec * 4;
            const uint32_t dst = s_base + p * kWinPlaneBytes + (r % kWinRows) * kWinRowBytes + cvec * 16;
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(dst), "l"(g) : "memory");
            p += dp; r += dr;
            if (p >= nplanes) { p -= nplanes; r++; }
          }
        }
        asm volatile("cp.async.wait_all;" ::: "memory");
      }
      __syncthreads();
      if (WIN_PREFETCH) {
        // The stacks of a launch (2.4 MB per face) do not stay in L2: pull the rows of the next tile step (or the first window of the
        // next item) towards L2 while this tile's walks run.  The warps share the planes: one bulk prefetch (UBLKPF, warp-uniform operands) per plane and
        // group of up to 8 FULL rows (4 KB contiguous): the other tile columns of the face, walked by the neighbouring CTAs at the same
        // time, want the rest of those rows anyway.  (A prefetch per 32-byte sector of the 160-byte row pieces, as k_traverse_win does,
        // costs 5 % of the kernel's instructions.)
        const stack_t* psrc = a.stacks + f * a.stack_face_stride;
        int pbeg = y0 + kWinTile + kWinExtent, pend = min(y0 + kWinTile + kWinRows, nrows);
        if (iyb + 1 >= ntile_y) {
          const int nitem = item + gridDim.x;
          pbeg = pend = 0;
          if (nitem < nitems) {
            const int nf = nitem / ncols;
            psrc = a.stacks + nf * a.stack_face_stride;
            pend = min(kWinRows, a.fd[nf].H + 1);
          }
        }
        for (int rb = pbeg; rb < pend; rb += 8) {
          const uint32_t bytes = (uint32_t)(min(pend - rb, 8) * kRowStride * sizeof(stack_t));
          for (int p = warp; p < nplanes; p += NW)
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" :: "l"(psrc + (size_t)p * a.plane_stride + (size_t)rb * kRowStride), "r"(bytes));
        }
      }
      const int r0 = y0 % kWinRows;
      const uint32_t col = s_base + lx * 4;
      if (WALKS == 2) {
        int ra = r0 + ly, rb = r0 + ly + 4;
        ra -= ra >= kWinRows ? kWinRows : 0;
        rb -= rb >= kWinRows ? kWinRows : 0;
        const uint32_t rowA = ra * kWinRowBytes, rowB = rb * kWinRowBytes;
        const bool vA = vx && y0 + ly < ny, vB = vx && y0 + ly + 4 < ny;
        for (int t = warp; t < nt; t += NW) {
          const int root = t == warp ? root0 : __ldg(a.nroot_of_slot + roots[t]);
          int ca = vA ? root : -1, cb = vB ? root : -1;   // lanes outside the grid start parked; their result is never stored
          uint4 a0, a1, b0, b1;
          if (t == warp) { a0 = rr0; a1 = rr1; } else fetch(ca, a0, a1);
          for (;;) {
            // Steady state: test A_k, fetch A_k+1, test B_k, fetch B_k+1, ... — one walk's fetch is in flight while the other walk is
            // tested.  The body starts at "fetch B" so that the fetches in flight are in the same order on every path into a test
            // (ptxas waits by counting them).  The warp-uniform branches skip a walk whose 32 lanes are all parked, and they are also
            // what keeps ptxas from merging the halves (it would sink A's fetch below B's gathers and lose the overlap).
            fetch(cb, b0, b1);
            const bool la = ca >= 0;
            if (__any_sync(0xffffffffu, la)) {
              const int na = win2_step(col, rowA, a0, a1, la ? 1u : 0u);
              ca = la ? na : ca;
              if (COUNT) tests += la ? 1 : 0;
            }
            fetch(ca, a0, a1);
            const bool lb = cb >= 0;
            if (__any_sync(0xffffffffu, lb)) {
              const int nb = win2_step(col, rowB, b0, b1, lb ? 1u : 0u);
              cb = lb ? nb : cb;
              if (COUNT) tests += lb ? 1 : 0;
            }
            if (!__any_sync(0xffffffffu, (ca & cb) >= 0)) break;   // no walk of the warp sits on a record any more
          }
          int va = ~ca, vb = ~cb;
          if (a.leaf_value) { if (vA) va = __float_as_int(__ldg(a.leaf_value + va)); if (vB) vb = __float_as_int(__ldg(a.leaf_value + vb)); }
          if (vA) out[((size_t)(x0 + lx) * ny + (y0 + ly)) * nt + t] = va;
          if (vB) out[((size_t)(x0 + lx) * ny + (y0 + ly + 4)) * nt + t] = vb;
        }
      } else {
        for (int k = warp; k < 2 * nt; k += NW) {
          const int t = k >> 1, py = ly + 4 * (k & 1);
          int ra = r0 + py;
          ra -= ra >= kWinRows ? kWinRows : 0;
          const uint32_t rowA = ra * kWinRowBytes;
          const bool vA = vx && y0 + py < ny;
          int ca = vA ? (k < NW ? root0 : __ldg(a.nroot_of_slot + roots[t])) : -1;
          uint4 a0, a1;
          while (__any_sync(0xffffffffu, ca >= 0)) {
            const bool la = ca >= 0;
            fetch(ca, a0, a1);
            const int n = win2_step(col, rowA, a0, a1, la ? 1u : 0u);
            ca = la ? n : ca;
            if (COUNT) tests += la ? 1 : 0;
          }
          int va = ~ca;
          if (a.leaf_value && vA) va = __float_as_int(__ldg(a.leaf_value + va));
          if (vA) out[((size_t)(x0 + lx) * ny + (y0 + py)) * nt + t] = va;
        }
      }
      if (COUNT && threadIdx.x == 0) trav += (unsigned long long)min(kWinTile, nx - x0) * min(kWinTile, ny - y0) * nt;
    }
  }
  if (COUNT) {
    tests = __reduce_add_sync(0xffffffffu, tests);
    if (lane == 0 && tests) atomicAdd(&a.counters[a.cnt_tests], (unsigned long long)tests);
    if (threadIdx.x == 0 && trav) atomicAdd(&a.counters[a.cnt_trav], trav);
  }
}

// Measured alternatives that did NOT pay on B200 (tools/traverse_variants.py; kept out of the source): two lanes per
// patch loading the left / right rectangle columns in one instruction and meeting through a shuffle (halves the cache
// lines touched per load at diverged levels: 1.4x slower at stride 1, with or without 2-4 trees in flight per lane),
// L1 no-allocate corner loads (1.2x slower), 8x4 lane blocks at stride 1 (1.2x slower), 16-bit integral planes with
// strip-split rectangles (1.7x slower, device_forest.h).

// ---------------------------------------------------------------------------------------------
// a8 (reduce) + a11: head-pose mean/variance in the reference's sequential f32 order
// (src/face_utils.cpp:219-241), areaUnderCurve and forest composition (src/face_utils.cpp:304-323,
// src/FaceForest.cpp:215-250).  One warp per face: lanes fetch 32 leaves,
// then every lane folds them in index order (same instruction stream, so no divergence).
// ---------------------------------------------------------------------------------------------
struct ComposeTables {
  const double* xs;        // concatenated Riemann abscissae of the 5 bins (x += 0.01 in double from (double)poseT[j])
  int bin_begin[6];
  const int32_t* jungle_roots;
  int forest_base[CRF_NUM_POSE_FORESTS];
  int forest_ntrees[CRF_NUM_POSE_FORESTS];
  int ntrees_cfg;
  int list_cap;            // trees a face may evaluate in this launch (<= kMaxList)
};

__device__ __forceinline__ void compose_face(float headpose, float variance, const ComposeTables& ct, int lane,
                                             crf_face_t* face, int32_t* list, int32_t* ntrees_out) {
  // areaUnderCurve(x1, x2, mean, std): lanes evaluate exp() of their abscissae, lane 0 folds in order.
  const double mean = (double)headpose, sd = sqrt((double)variance);
  __shared__ double s_e[10][32];   // one slot per warp of the callers (at most 10 warps: k_hp_reduce_compose)
  double* e = s_e[(threadIdx.x >> 5) % 10];
  float area[CRF_NUM_POSE_FORESTS];
  for (int j = 0; j < CRF_NUM_POSE_FORESTS; j++) {
    double sum = 0;
    for (int b = ct.bin_begin[j]; b < ct.bin_begin[j + 1]; b += 32) {
      const int i = b + lane;
      double v = 0;
      if (i < ct.bin_begin[j + 1]) {
        const double t = (ct.xs[i] - mean) / sd;
        v = exp(-0.5 * (t * t)) * 0.01;
      }
      e[lane] = v;
      __syncwarp();
      const int n = min(32, ct.bin_begin[j + 1] - b);
      for (int k = 0; k < n; k++) sum += e[k];
      __syncwarp();
    }
    area[j] = (float)(sum * 1.0 / (sd * sqrt(2 * 3.14159265358979323846)));
  }
  if (lane != 0) return;
  float max_area = 0;
  int dominant = 0;
  for (int j = 0; j < CRF_NUM_POSE_FORESTS; j++)
    if (max_area < area[j]) { max_area = area[j]; dominant = j; }
  int n = 0, flags = 0;
  for (int i = 0; i < CRF_NUM_POSE_FORESTS; i++) {
    const float prod = area[i] * ct.ntrees_cfg;
    const double fl = floor((double)prod);
    int cnt_req = (fl != fl || fl < -2147483648.0 || fl > 2147483647.0) ? INT_MIN : (int)fl;  // x86 cvttsd2si semantics
    if (cnt_req > ct.forest_ntrees[i]) { cnt_req = ct.forest_ntrees[i]; flags |= 1; }
    int cnt = 0;
    for (int j = 0; j < cnt_req; j++) {
      if (n < ct.list_cap) { list[n++] = ct.jungle_roots[ct.forest_base[i] + j]; cnt++; }
      else flags |= 2;  // more trees than this launch holds: the engine re-runs the face with a wider list
    }
    face->tree_counts[i] = cnt;
  }
  for (int i = n; i < ct.ntrees_cfg; i++) {
    if (i >= ct.forest_ntrees[dominant]) { flags |= 1; break; }
    if (n < ct.list_cap) list[n++] = ct.jungle_roots[ct.forest_base[dominant] + i];
    else flags |= 2;
  }
  face->dominant = dominant;
  face->flags = flags;
  *ntrees_out = n;
}

// Sequential folds are latency-bound chains (one dependent FADD every 4 cycles), so 32 of them share a warp:
// lane j of warp 0 folds chain j from shared-memory tiles that the other 8 warps fill one tile ahead
// (double buffer).  Tiles are laid out [element][chain] with a pitch of 33 words: conflict-free for the
// producers (lanes = elements) and for the fold (lanes = chains).
constexpr int kFoldChains = 32;
constexpr int kFoldThreads = 288;  // warp 0 folds, warps 1..8 produce (4 chains each)

constexpr int kHpTile = 256;
constexpr int kHpPitch = kHpTile + 4;   // chain-major tiles: 260 == 4 (mod 32), so the 8 lanes of a quarter warp reading 16 B each hit 32 different banks
constexpr size_t kHpSmem = (size_t)2 * kFoldChains * kHpPitch * sizeof(float);  // dynamic shared memory of k_hp_reduce_compose

__global__ void __launch_bounds__(kFoldThreads) k_hp_reduce_compose(const FaceDesc* __restrict__ fd, int nfaces, const float* __restrict__ leaf_m, size_t leaf_face_stride,
                                                                    int ntrees, int stride, ComposeTables ct, int do_compose,
                                                                    crf_face_t* __restrict__ faces, int32_t* __restrict__ face_roots, int32_t* __restrict__ face_ntrees) {
  extern __shared__ __align__(16) float s_hp_dyn[];
  typedef float HpTile[kFoldChains][kHpPitch];   // [chain][element]: producers write coalesced rows, the fold lane reads its row by float4
  HpTile* s_m = reinterpret_cast<HpTile*>(s_hp_dyn);  // [2]
  __shared__ int s_n[kFoldChains], s_cnt[kFoldChains];
  __shared__ float s_mean[kFoldChains], s_var[kFoldChains];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int f0 = blockIdx.x * kFoldChains;
  if (threadIdx.x < kFoldChains) {
    const int f = f0 + threadIdx.x;
    int n = 0;
    if (f < nfaces) {
      const FaceDesc d = fd[f];
      const int nx = (d.W - kPatch + stride - 1) / stride, ny = (d.H - kPatch + stride - 1) / stride;
      n = max(nx, 0) * max(ny, 0) * ntrees;
    }
    s_n[threadIdx.x] = n;
    s_cnt[threadIdx.x] = 0;
  }
  __syncthreads();
  int maxn = 0;
  for (int j = 0; j < kFoldChains; j++) maxn = max(maxn, s_n[j]);
  const int ntiles = (maxn + kHpTile - 1) / kHpTile;
  // The producers do everything of the reference's loop body that is not the order-dependent sum (src/face_utils.cpp:222-231): the
  // fg > min_foreground_probability test is folded into the table at load time (-1 = skip), a skipped vote is stored as +0 (adding +0 is
  // exact: the sums never hold -0), and the votes are counted as integers (the reference's float count of 1.0s is exact below 2^24).  The
  // single fold warp is left with load, add, multiply, add per element (a second fold warp for the sum of squares was measured: no gain, the
  // dependent add chain of the sum is the critical path).  This fold is the latency of a single-face call at stride 1.
  int valid_count[4] = {0, 0, 0, 0};
  // A producer keeps the NEXT tile's values in registers across the barrier (loads issued one iteration before they are stored), so the
  // global-load latency overlaps the fold of a whole tile: with a single live chain the fold is otherwise faster than the load.
  float pv[4][kHpTile / 32];
  auto load_tile = [&](int tile) {
#pragma unroll
    for (int jj = 0; jj < 4; jj++) {
      const int j = (warp - 1) * 4 + jj;
      const int n = s_n[j];
      const float* __restrict__ ms = leaf_m + (size_t)(f0 + j) * leaf_face_stride;
#pragma unroll
      for (int h = 0; h < kHpTile / 32; h++) {
        const int k = tile * kHpTile + h * 32 + lane;
        pv[jj][h] = k < n ? ms[k] : -1.f;
      }
    }
  };
  auto store_tile = [&](int tile) {
    float(*buf)[kHpPitch] = s_m[tile & 1];
#pragma unroll
    for (int jj = 0; jj < 4; jj++) {
      const int j = (warp - 1) * 4 + jj;
      if (tile * kHpTile >= s_n[j]) continue;
#pragma unroll
      for (int h = 0; h < kHpTile / 32; h++) {
        const bool valid = !(pv[jj][h] < 0.f);
        valid_count[jj] += __popc(__ballot_sync(0xffffffffu, valid));
        buf[j][h * 32 + lane] = valid ? pv[jj][h] : 0.f;
      }
    }
  };
  float sum = 0, sum_sq = 0;
  if (warp > 0 && ntiles > 0) { load_tile(0); store_tile(0); if (ntiles > 1) load_tile(1); }
  __syncthreads();
  for (int tile = 0; tile < ntiles; tile++) {
    if (warp == 0) {
      const float4* row = reinterpret_cast<const float4*>(s_m[tile & 1][lane]);
      // the producers stop refreshing the row of a chain that has ended, so its lane adds +0 from here on
      const bool live = tile * kHpTile < s_n[lane];
#pragma unroll 8
      for (int k4 = 0; k4 < kHpTile / 4; k4++) {
        float4 v = row[k4];
        if (!live) v = make_float4(0.f, 0.f, 0.f, 0.f);
        sum += v.x; sum_sq += v.x * v.x;
        sum += v.y; sum_sq += v.y * v.y;
        sum += v.z; sum_sq += v.z * v.z;
        sum += v.w; sum_sq += v.w * v.w;
      }
    } else if (tile + 1 < ntiles) {
      store_tile(tile + 1);
      if (tile + 2 < ntiles) load_tile(tile + 2);
    }
    __syncthreads();
  }
  if (warp > 0 && lane == 0) {
#pragma unroll
    for (int jj = 0; jj < 4; jj++) if (valid_count[jj]) atomicAdd(&s_cnt[(warp - 1) * 4 + jj], valid_count[jj]);
  }
  __syncthreads();
  if (warp == 0) {
    const float cnt = (float)s_cnt[lane];
    float mean = sum / cnt;
    float var = (sum_sq / cnt) - (mean * mean);
    mean -= 2;
    var *= 0.05f;  // NORM_HEADPOSE_VARIANCE_FACTOR (include/Constants.hpp:67)
    s_mean[lane] = mean; s_var[lane] = var;
    const int f = f0 + lane;
    if (f < nfaces) {
      const FaceDesc d = fd[f];
      crf_face_t* face = faces + f;
      face->headpose = mean;
      face->variance = var;
      face->scaled_w = d.W; face->scaled_h = d.H; face->scale = d.scale;
    }
  }
  __syncthreads();
  if (do_compose)
    for (int j = warp; j < kFoldChains && f0 + j < nfaces; j += kFoldThreads / 32)
      compose_face(s_mean[j], s_var[j], ct, lane, faces + f0 + j, face_roots + (size_t)(f0 + j) * kMaxList, face_ntrees + f0 + j);
}

// areaUnderCurve(x1, x2, mean, std) (src/face_utils.cpp:304-323) for arbitrary bounds: the reference's serial Riemann sum
// (x += 0.01 in double from (double)x1) on one thread.
__global__ void k_area_under_curve(float x1, float x2, double mean, double sd, float* out) {
  double sum = 0;
  const double step = 0.01;
  for (double x = x1; x < x2; x += step) {
    const double t = (x - mean) / sd;
    sum += exp(-0.5 * (t * t)) * step;
  }
  *out = (float)(sum * 1.0 / (sd * sqrt(2 * 3.14159265358979323846)));
}

// Composition alone (stage API).
__global__ void k_compose_only(float headpose, float variance, ComposeTables ct, crf_face_t* face, int32_t* list, int32_t* ntrees) {
  compose_face(headpose, variance, ct, threadIdx.x & 31, face, list, ntrees);
}

// Composition of many (headpose, variance) pairs (stage API: knife-edge sweeps of floor(area * ntrees)).  One warp per pair;
// 288 threads = the 9 warps compose_face's shared scratch is sized for.
__global__ void __launch_bounds__(288) k_compose_batch(const float* __restrict__ headpose, const float* __restrict__ variance, int n, ComposeTables ct,
                                                       crf_face_t* __restrict__ faces, int32_t* __restrict__ lists, int32_t* __restrict__ ntrees) {
  const int i = blockIdx.x * 9 + (threadIdx.x >> 5);
  if (i >= n) return;
  compose_face(headpose[i], variance[i], ct, threadIdx.x & 31, faces + i, lists + (size_t)i * kMaxList, ntrees + i);
}

// ---------------------------------------------------------------------------------------------
// a12 (emission): ordered vote lists (src/face_utils.cpp:277-301).  Votes of part i keep
// the reference's order (leaf index = patch-major, tree-minor) because MeanShift sums them in f32.
// One CTA per face walks the leaves in chunks of 256 and compacts per part with warp ballots.
// votes: [face][face_vote_cap] {x, y, weight}; the list of part p starts at vote_base[face][p] (exclusive scan of the
// per-part counts).  face_vote_cap is a budget (a few votes per leaf); a face that exceeds it is flagged (bit 2)
// and re-run by the engine with the worst-case capacity.
// ---------------------------------------------------------------------------------------------
struct __align__(8) DevVote { short x, y; float w; };

// The leaves of a face are split into kVoteSegs contiguous warp-segments.  Pass 1 counts the votes of every
// (segment, part); pass 2 derives each segment's base by summing the earlier segments and compacts with warp
// ballots only (no block barrier): a warp owns a contiguous run of leaves, so order is preserved.
constexpr int kVoteSegs = 128;

struct VoteArgs {
  const FaceDesc* fd;
  const int32_t* leaf_ids; size_t leaf_face_stride;
  const int32_t* face_ntrees; int stride;
  const uint16_t* mp_mask; const DevMpLeaf* mp_leaf;
  DevVote* votes; size_t vote_cap;   // votes per face
  int32_t* seg_counts;   // [face][kVoteSegs][kParts]
  int32_t* vote_counts;  // [face][kParts] (zeroed before pass 1)
  int32_t* vote_base;    // [face][kParts] start of each part's list inside the face's region
  crf_face_t* faces;
};

__device__ __forceinline__ void vote_segment(const VoteArgs& a, int f, int seg, int& n, int& k0, int& k1, int& nt, int& ny) {
  const FaceDesc d = a.fd[f];
  nt = a.face_ntrees[f];
  const int nx = (d.W - kPatch + a.stride - 1) / a.stride;
  ny = (d.H - kPatch + a.stride - 1) / a.stride;
  n = max(nx, 0) * max(ny, 0) * nt;
  const int per = (((n + kVoteSegs - 1) / kVoteSegs) + 31) & ~31;
  k0 = min(n, seg * per);
  k1 = min(n, k0 + per);
}

// grid = (kVoteSegs / 8, faces), 256 threads: one warp per segment.
__global__ void __launch_bounds__(256) k_votes_count(VoteArgs a) {
  const int f = blockIdx.y, lane = threadIdx.x & 31, seg = blockIdx.x * 8 + (threadIdx.x >> 5);
  int n, k0, k1, nt, ny;
  vote_segment(a, f, seg, n, k0, k1, nt, ny);
  const int32_t* __restrict__ ids = a.leaf_ids + f * a.leaf_face_stride;
  // Per-lane counts of the ten parts as 6-bit fields of two words (a lane sees at most ceil(n / kVoteSegs / 32) + 1 <= 63 leaves of a
  // segment for every grid this library launches: checked below), summed over the warp once at the end.  A ballot + POPC per part and step
  // (POPC issues at a quarter of the ALU rate) made this kernel math-pipe-bound: 2.4 ms for a 2.9 GB read.
  // spread5(x): bit i of x -> bit 6 i (the cross terms of the multiply land between the fields and are masked off).
  auto spread5 = [](unsigned x) { return (x * 0x108421u) & 0x01041041u; };
  unsigned lo = 0, hi = 0;
  const bool swar = (k1 - k0 + 31) / 32 <= 63;
  int cnt = 0;  // fallback (segments longer than 63 x 32 leaves): lane p < kParts accumulates part p from ballots
  for (int kb = k0; kb < k1; kb += 128) {   // four steps per trip: the id -> mask gathers of all four are in flight together
    int id[4];
    unsigned mask[4];
#pragma unroll
    for (int u = 0; u < 4; u++) { const int k = kb + u * 32 + lane; id[u] = k < k1 ? ids[k] : -1; }
#pragma unroll
    for (int u = 0; u < 4; u++) mask[u] = id[u] >= 0 ? (unsigned)__ldg(a.mp_mask + id[u]) : 0u;
    if (swar) {
#pragma unroll
      for (int u = 0; u < 4; u++) { lo += spread5(mask[u] & 31u); hi += spread5(mask[u] >> 5); }
    } else {
#pragma unroll
      for (int p = 0; p < kParts; p++) {
        int c = 0;
#pragma unroll
        for (int u = 0; u < 4; u++) c += __popc(__ballot_sync(0xffffffffu, (mask[u] >> p) & 1u));
        if (lane == p) cnt += c;
      }
    }
  }
  if (swar) {
#pragma unroll
    for (int p = 0; p < kParts; p++) {
      const int c = (int)__reduce_add_sync(0xffffffffu, ((p < 5 ? lo : hi) >> (6 * (p % 5))) & 63u);
      if (lane == p) cnt = c;
    }
  }
  if (lane < kParts) {
    a.seg_counts[((size_t)f * kVoteSegs + seg) * kParts + lane] = cnt;
    if (cnt) atomicAdd(&a.vote_counts[f * kParts + lane], cnt);
  }
}

// One thread per face: exclusive scan of the 10 per-part counts; capacity check.
__global__ void k_votes_offsets(VoteArgs a, int nfaces) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= nfaces) return;
  long long run = 0;
  for (int p = 0; p < kParts; p++) { a.vote_base[f * kParts + p] = (int32_t)run; run += a.vote_counts[f * kParts + p]; }
  if (run > (long long)a.vote_cap) {   // over budget: emit nothing, flag the face for the wide re-run
    for (int p = 0; p < kParts; p++) { a.vote_base[f * kParts + p] = 0; a.vote_counts[f * kParts + p] = -1; }
    a.faces[f].flags |= 4;
  }
}

__global__ void __launch_bounds__(256) k_votes_emit(VoteArgs a) {
  const int f = blockIdx.y, lane = threadIdx.x & 31, seg = blockIdx.x * 8 + (threadIdx.x >> 5);
  int n, k0, k1, nt, ny;
  vote_segment(a, f, seg, n, k0, k1, nt, ny);
  if (k0 >= k1 || a.vote_counts[f * kParts] < 0) return;
  // base of this segment for every part: sum over the earlier segments (lanes stride over them)
  int base[kParts];
  const int32_t* sc = a.seg_counts + (size_t)f * kVoteSegs * kParts;
#pragma unroll
  for (int p = 0; p < kParts; p++) {
    int v = 0;
    for (int s = lane; s < seg; s += 32) v += sc[s * kParts + p];
    base[p] = __reduce_add_sync(0xffffffffu, v) + a.vote_base[f * kParts + p];
  }
  const int32_t* __restrict__ ids = a.leaf_ids + f * a.leaf_face_stride;
  DevVote* __restrict__ fv = a.votes + (size_t)f * a.vote_cap;
  // k / nt and patch / ny by multiplication: umulhi(k, 0xffffffff / d + 1) is exact for every k < 2^32 / d (k < 2^23 with
  // d = nt <= 128, patch < 2^16 with d = ny <= 490; checked exhaustively in test_index_division_by_multiplication)
  const unsigned m_nt = 0xffffffffu / (unsigned)nt + 1u, m_ny = 0xffffffffu / (unsigned)max(ny, 1) + 1u;
  int leaf_n = 0;
  unsigned mask_n = 0;
  if (k0 + lane < k1) { leaf_n = ids[k0 + lane]; mask_n = __ldg(a.mp_mask + leaf_n); }
  // Ranks of a step's votes inside their part lists WITHOUT a ballot + POPC per part (POPC issues at a quarter of the ALU rate and there
  // were twenty per step): every lane spreads its 10-bit mask into ten 6-bit fields of two words (spread5, as in k_votes_count), an
  // inclusive scan over the lanes (5 shuffle steps on the two words; a field holds at most 32) gives all ten ranks at once, and lane
  // 31's words are the step's totals.
  auto spread5 = [](unsigned x) { return (x * 0x108421u) & 0x01041041u; };
  for (int k = k0 + lane; k - lane < k1; k += 32) {
    const int leaf = leaf_n;
    const unsigned mask = mask_n;
    leaf_n = 0; mask_n = 0;
    if (k + 32 < k1) { leaf_n = ids[k + 32]; mask_n = __ldg(a.mp_mask + leaf_n); }   // next step's gathers overlap this step's emission
    const unsigned own_lo = spread5(mask & 31u), own_hi = spread5(mask >> 5);
    unsigned lo = own_lo, hi = own_hi;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned nl = __shfl_up_sync(0xffffffffu, lo, o), nh = __shfl_up_sync(0xffffffffu, hi, o);
      if (lane >= o) { lo += nl; hi += nh; }
    }
    const unsigned tot_lo = __shfl_sync(0xffffffffu, lo, 31), tot_hi = __shfl_sync(0xffffffffu, hi, 31);
    if (mask) {
      const unsigned ex_lo = lo - own_lo, ex_hi = hi - own_hi;   // votes of the lanes before this one, per part
      const int patch = nt == 1 ? k : (int)__umulhi((unsigned)k, m_nt);
      const int ix = ny == 1 ? patch : (int)__umulhi((unsigned)patch, m_ny), iy = patch - ix * ny;
      const int cx = ix * a.stride + kHalfPatch, cy = iy * a.stride + kHalfPatch;  // patch centre (src/face_utils.cpp:281-282)
      union { uint4 q[3]; DevMpLeaf L; } u;
      const uint4* lp = reinterpret_cast<const uint4*>(a.mp_leaf + leaf);
      u.q[0] = __ldg(lp); u.q[1] = __ldg(lp + 1); u.q[2] = __ldg(lp + 2);
      const DevMpLeaf& L = u.L;
#pragma unroll
      for (int p = 0; p < kParts; p++) {
        if ((mask >> p) & 1u) {
          DevVote v;
          v.x = (short)(L.off[p][0] + cx);
          v.y = (short)(L.off[p][1] + cy);
          v.w = L.weight;
          fv[base[p] + (int)(((p < 5 ? ex_lo : ex_hi) >> (6 * (p % 5))) & 63u)] = v;
        }
      }
    }
#pragma unroll
    for (int p = 0; p < kParts; p++) base[p] += (int)(((p < 5 ? tot_lo : tot_hi) >> (6 * (p % 5))) & 63u);
  }
}

// ---------------------------------------------------------------------------------------------
// a14: MeanShift::shift (include/MeanShift.hpp:52-135) + the final rescale of
// FaceForest::analyzeFace (src/FaceForest.cpp:256-257).  The f32 sums are order-dependent, so each
// (face, part) chain is folded sequentially — 32 chains per CTA, lane j of warp 0 folds chain j while
// warps 1..8 evaluate the kernel weights of the next tile of votes in parallel (see the fold note above).
// ---------------------------------------------------------------------------------------------
struct MeanShiftOpt { int kernel; int max_iterations; float stopping; };

// expf as glibc computes it (sysdeps/ieee754/flt-32/e_expf.c: 2^(k/32) table + cubic in double, one final
// rounding): bit-identical to the host's expf on every input tried (4e7 in [-60, 0]), which a correctly
// rounded exp() is not.  tab[i] = bits(2^(i/32)) - (i << 47), built on the host with exp2().
struct ExpfTable { unsigned long long tab[32]; };
__constant__ ExpfTable c_expf;

__device__ __forceinline__ float expf_glibc(float x, const unsigned long long* __restrict__ tab) {
  const double N = 32.0;
  const double InvLn2N = 0x1.71547652b82fep+0 * N;
  const double C0 = 0x1.c6af84b912394p-5 / N / N / N, C1 = 0x1.ebfce50fac4f3p-3 / N / N, C2 = 0x1.62e42ff0c52d6p-1 / N;
  const double xd = (double)x;
  if (!(xd > -103.0)) return xd != xd ? x : 0.f;   // underflow (arguments here are -distance/lambda <= 0)
  if (xd > 88.0) return __int_as_float(0x7f800000);
  const double z = InvLn2N * xd;
  const double kd = rint(z);
  const long long ki = (long long)kd;
  const double r = z - kd;
  const double s = __longlong_as_double((long long)(tab[ki & 31] + ((unsigned long long)ki << 47)));
  const double zz = fma(C0, r, C1);
  const double r2 = r * r;
  double y = fma(C2, r, 1.0);
  y = fma(zz, r2, y);
  return (float)(y * s);
}

__device__ __forceinline__ void vote_terms(const DevVote q, bool first_pass, float mx, float my, float lamda, const unsigned long long* __restrict__ tab,
                                           float& w, float& wx, float& wy) {
  if (first_pass) {
    w = q.w;
  } else {
    const float dx = mx - (float)q.x, dy = my - (float)q.y;
    const float dist = (float)sqrt((double)dx * dx + (double)dy * dy);  // cv::norm(Point2f) accumulates in double
    w = q.w * expf_glibc(-dist / lamda, tab);
  }
  wx = q.x * w;
  wy = q.y * w;
}

// Dense CTAs: 32 chains, tiles of 64 votes.  Sparse CTAs (small batches): at most 8 chains and tiles of 256 votes, so a
// chain's pass takes a quarter of the barrier-separated steps.
template <bool SPARSE> struct MsGeom {
  static constexpr int TILE = SPARSE ? 256 : 64;
  static constexpr int PITCH = SPARSE ? 9 : 33;     // odd pitches: conflict-free for producers (lanes = votes) and the fold (lanes = chains)
  static constexpr int MAX_CPC = SPARSE ? 8 : 32;
  static constexpr size_t smem = (size_t)2 * 3 * TILE * PITCH * sizeof(float);   // dynamic shared memory of k_meanshift
};

template <int MINB, bool SPARSE>
__global__ void __launch_bounds__(kFoldThreads, MINB) k_meanshift(const FaceDesc* __restrict__ fd, int nchains, const DevVote* __restrict__ votes, size_t vote_cap,
                                                               const int32_t* __restrict__ vote_counts, const int32_t* __restrict__ vote_base, int cpc /* chains per CTA, <= 32 */, MeanShiftOpt o, crf_face_t* __restrict__ faces,
                                                            unsigned long long* counters) {
  extern __shared__ __align__(16) float s_dyn[];
  constexpr int kMsTile = MsGeom<SPARSE>::TILE;
  typedef float Tile[kMsTile][MsGeom<SPARSE>::PITCH];
  Tile* s_w = reinterpret_cast<Tile*>(s_dyn);   // [2]
  Tile* s_x = s_w + 2;
  Tile* s_y = s_w + 4;
  __shared__ int s_n[kFoldChains], s_active[kFoldChains];
  __shared__ const DevVote* s_list[kFoldChains];
  __shared__ float s_mx[kFoldChains], s_my[kFoldChains];
  __shared__ unsigned long long s_tab[32];
  __shared__ int s_maxn, s_nact, s_act[kFoldChains];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c0 = blockIdx.x * cpc;
  if (threadIdx.x < kFoldChains) {
    const int c = c0 + threadIdx.x;
    const bool mine = threadIdx.x < cpc && c < nchains;
    const int cnt = mine ? vote_counts[c] : 0;
    s_n[threadIdx.x] = max(cnt, 0);
    s_active[threadIdx.x] = mine && cnt >= 0;   // cnt < 0: face over the vote budget, left to the wide re-run
    s_list[threadIdx.x] = mine ? votes + (size_t)(c / kParts) * vote_cap + vote_base[c] : votes;
    s_mx[threadIdx.x] = 0.f; s_my[threadIdx.x] = 0.f;
    s_tab[threadIdx.x] = c_expf.tab[threadIdx.x];
  }
  __syncthreads();
  const float lamda = (float)o.kernel;
  int it = 0;                 // warp 0, lane j: iterations of chain j
  float mx = 0.f, my = 0.f;   // warp 0, lane j: current mean of chain j
  for (int pass = 0; pass <= o.max_iterations; pass++) {
    if (warp == 0) {
      int m = s_active[lane] ? max(s_n[lane], 1) : 0;   // an empty chain still runs its (empty) passes
      m = __reduce_max_sync(0xffffffffu, m);
      // compact list of the chains that still have votes to weigh: few of them (small batches, late passes) are spread
      // over all producer warps instead of leaving most of them idle
      const unsigned live = __ballot_sync(0xffffffffu, s_active[lane] && s_n[lane] > 0);
      if (s_active[lane] && s_n[lane] > 0) s_act[__popc(live & ((1u << lane) - 1u))] = lane;
      if (lane == 0) { s_maxn = m; s_nact = __popc(live); }
    }
    __syncthreads();
    const int maxn = s_maxn;
    if (maxn == 0) break;
    const int nact = s_nact;
    const int ntiles = (maxn + kMsTile - 1) / kMsTile;
    auto produce = [&](int tile) {
      const int b = tile & 1;
      if (SPARSE) {   // few live chains per CTA: items = (live chain, 32-vote group), round-robin over the 8 producer warps
        constexpr int G = kMsTile / 32;
#pragma unroll 2
        for (int it = warp - 1; it < nact * G; it += 8) {
          const int j = s_act[it / G], h = it % G;
          const int k = tile * kMsTile + h * 32 + lane;
          if (k < s_n[j]) {
            float w, wx, wy;
            vote_terms(s_list[j][k], pass == 0, s_mx[j], s_my[j], lamda, s_tab, w, wx, wy);
            s_w[b][h * 32 + lane][j] = w; s_x[b][h * 32 + lane][j] = wx; s_y[b][h * 32 + lane][j] = wy;
          }
        }
      } else {
      DevVote q[4][kMsTile / 32];
      bool ok[4][kMsTile / 32];
#pragma unroll
      for (int jj = 0; jj < 4; jj++) {
        const int j = (warp - 1) * 4 + jj;
        const int n = s_active[j] ? s_n[j] : 0;
#pragma unroll
        for (int h = 0; h < kMsTile / 32; h++) {
          const int k = tile * kMsTile + h * 32 + lane;
          ok[jj][h] = k < n;
          if (ok[jj][h]) q[jj][h] = s_list[j][k];
        }
      }
#pragma unroll
      for (int jj = 0; jj < 4; jj++) {
        const int j = (warp - 1) * 4 + jj;
        const float cmx = s_mx[j], cmy = s_my[j];
#pragma unroll
        for (int h = 0; h < kMsTile / 32; h++) {
          if (!ok[jj][h]) continue;
          float w, wx, wy;
          vote_terms(q[jj][h], pass == 0, cmx, cmy, lamda, s_tab, w, wx, wy);
          s_w[b][h * 32 + lane][j] = w; s_x[b][h * 32 + lane][j] = wx; s_y[b][h * 32 + lane][j] = wy;
        }
      }
      }
    };
    float sw = 0.f, sx = 0.f, sy = 0.f;
    if (warp > 0) produce(0);
    __syncthreads();
    for (int tile = 0; tile < ntiles; tile++) {
      if (warp == 0) {
        const int b = tile & 1;
        const int kmax = s_active[lane] ? min(kMsTile, s_n[lane] - tile * kMsTile) : 0;
        const int jl = lane < MsGeom<SPARSE>::MAX_CPC ? lane : 0;   // lanes beyond the CTA's chains fold nothing (kmax = 0)
#pragma unroll 8
        for (int k = 0; k < kMsTile; k++) {   // branch-free: adding +0 is exact (the sums never hold -0)
          const bool valid = k < kmax;
          sx += valid ? s_x[b][k][jl] : 0.f;
          sy += valid ? s_y[b][k][jl] : 0.f;
          sw += valid ? s_w[b][k][jl] : 0.f;
        }
      } else if (tile + 1 < ntiles) {
        produce(tile + 1);
      }
      __syncthreads();
    }
    if (warp == 0 && s_active[lane]) {
      if (sw > 0) { sx /= sw; sy /= sw; }
      if (pass == 0) {
        mx = sx; my = sy;
        if (o.max_iterations <= 0) s_active[lane] = 0;
      } else {
        const float ex = sx - mx, ey = sy - my;
        const bool conv = sqrt((double)ex * ex + (double)ey * ey) < o.stopping;
        mx = sx; my = sy;
        it++;
        if (conv || it >= o.max_iterations) s_active[lane] = 0;
      }
      s_mx[lane] = mx; s_my[lane] = my;
    }
    __syncthreads();
  }
  if (warp == 0 && lane < cpc && c0 + lane < nchains) {
    const int c = c0 + lane, f = c / kParts, p = c - f * kParts;
    const int rx = __float2int_rn(mx), ry = __float2int_rn(my);  // Point_<int> = Point_<float>: cvRound
    crf_face_t* face = faces + f;
    face->ffd_f[p][0] = mx; face->ffd_f[p][1] = my;
    face->ffd_scaled[p][0] = rx; face->ffd_scaled[p][1] = ry;
    const float inv = 1.0f / fd[f].scale;  // Point_<int> *= float: saturate_cast<int>(x * b)
    face->ffd[p][0] = __float2int_rn(rx * inv);
    face->ffd[p][1] = __float2int_rn(ry * inv);
    face->ms_iters[p] = it;
    face->n_votes[p] = s_n[lane];
    if (counters) {
      atomicAdd(&counters[CNT_VOTES], (unsigned long long)s_n[lane]);
      atomicAdd(&counters[CNT_VOTE_PASSES], (unsigned long long)s_n[lane] * (1 + it));
    }
  }
}


// ---------------------------------------------------------------------------------------------
// a14, tolerance mode (crf_options_t::ms_mode = CRF_MS_FAST, the default): the same MeanShift::shift iteration
// (include/MeanShift.hpp:52-135) with the three sums of a pass reduced as a fixed tree instead of the reference's
// sequential f32 fold, f32 distance and exp2-based kernel weight.  One CTA per (face, part) chain: its vote list
// (8 B per vote) is streamed once per pass with coalesced loads and stays L2-resident between passes; per-thread
// partial sums -> warp butterfly -> fixed-order sum over the warps, so the result is deterministic (independent of
// scheduling) though not bit-identical to the sequential order.  The north star's landmark tolerance is 0.5 px; the
// observed distance to the exact mode is <= 0.024 px on 3543 campaign faces (profiles/r2w_parity_campaign_fast.json).  Head pose, composition and the
// vote lists are untouched by the mode.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float ex2_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float sqrt_approx(float x) { float y; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

template <int BLOCK>
__global__ void __launch_bounds__(BLOCK) k_meanshift_fast(const FaceDesc* __restrict__ fd, int nchains, const DevVote* __restrict__ votes, size_t vote_cap,
                                                          const int32_t* __restrict__ vote_counts, const int32_t* __restrict__ vote_base, MeanShiftOpt o,
                                                          crf_face_t* __restrict__ faces, unsigned long long* counters) {
  constexpr int NWARP = BLOCK / 32;
  __shared__ float s_part[2][3][NWARP];
  const int c = blockIdx.x;
  if (c >= nchains) return;
  const int f = c / kParts, p = c - f * kParts;
  const int cnt = vote_counts[c];
  if (cnt < 0) return;   // face over the vote budget: left to the wide re-run
  const int2* __restrict__ v = reinterpret_cast<const int2*>(votes + (size_t)f * vote_cap + vote_base[c]);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float nk = -1.4426950408889634f / (float)o.kernel;   // exp(-d / lamda) = 2^(d * nk)
  float mx = 0.f, my = 0.f;
  int it = 0;
  for (int pass = 0; pass <= o.max_iterations; pass++) {
    float sw = 0.f, sx = 0.f, sy = 0.f;
#pragma unroll 4
    for (int k = tid; k < cnt; k += BLOCK) {
      const int2 q = __ldg(v + k);
      // int16 -> float through the mantissa (1.5 * 2^23 + i is exact for |i| < 2^22): an integer add and a float subtract instead of I2F, which
      // issues on the quarter-rate XU pipe — with two I2F, a reciprocal square root and an exp2 per vote and pass this kernel was XU-bound
      const float x = __int_as_float(0x4B400000 + (int)(short)(q.x & 0xffff)) - 12582912.f, y = __int_as_float(0x4B400000 + (q.x >> 16)) - 12582912.f;
      float w = __int_as_float(q.y);
      if (pass > 0) {
        const float dx = mx - x, dy = my - y;
        w *= ex2_approx(sqrt_approx(__fmaf_rn(dx, dx, dy * dy)) * nk);
      }
      sw += w;
      sx = __fmaf_rn(x, w, sx);
      sy = __fmaf_rn(y, w, sy);
    }
#pragma unroll
    for (int ofs = 16; ofs > 0; ofs >>= 1) {
      sw += __shfl_xor_sync(0xffffffffu, sw, ofs);
      sx += __shfl_xor_sync(0xffffffffu, sx, ofs);
      sy += __shfl_xor_sync(0xffffffffu, sy, ofs);
    }
    float (*part)[NWARP] = s_part[pass & 1];
    if (lane == 0) { part[0][warp] = sw; part[1][warp] = sx; part[2][warp] = sy; }
    __syncthreads();   // double-buffered: the next pass writes the other buffer, so one barrier per pass suffices
    sw = sx = sy = 0.f;
#pragma unroll
    for (int j = 0; j < NWARP; j++) { sw += part[0][j]; sx += part[1][j]; sy += part[2][j]; }   // same order in every thread
    if (sw > 0.f) { sx /= sw; sy /= sw; }
    if (pass == 0) {
      mx = sx; my = sy;
    } else {
      const float ex = sx - mx, ey = sy - my;
      const bool conv = sqrt((double)ex * ex + (double)ey * ey) < o.stopping;
      mx = sx; my = sy;
      it++;
      if (conv) break;
    }
  }
  if (tid == 0) {
    const int rx = __float2int_rn(mx), ry = __float2int_rn(my);
    crf_face_t* face = faces + f;
    face->ffd_f[p][0] = mx; face->ffd_f[p][1] = my;
    face->ffd_scaled[p][0] = rx; face->ffd_scaled[p][1] = ry;
    const float inv = 1.0f / fd[f].scale;
    face->ffd[p][0] = __float2int_rn(rx * inv);
    face->ffd[p][1] = __float2int_rn(ry * inv);
    face->ms_iters[p] = it;
    face->n_votes[p] = cnt;
    if (counters) {
      atomicAdd(&counters[CNT_VOTES], (unsigned long long)cnt);
      atomicAdd(&counters[CNT_VOTE_PASSES], (unsigned long long)cnt * (1 + it));
    }
  }
}

}  // namespace crf
