// Per-frame image preprocessing of the reference's dataset layer on the device: what the reference does
// with OpenCV on the host before every frame reaches the UNet / the prior --
//   ImageSample._process_image  (awesome/dataset/image_sample.py:212-221): (image*255) -> uint8 -> GaussianBlur 5x5 ->
//                               /255 (float32) -> optional BGR order
//   ImageSample.create_edge_map (image_sample.py:260-275): uint8 -> GaussianBlur 3x3 -> RGB2GRAY -> Sobel x / y (16S) ->
//                               convertScaleAbs -> addWeighted(.5, .5) -> /255 (float64) -> GaussianBlur 5x5 -> float32
// with OpenCV's arithmetic restated exactly (fixed binomial kernels for sigma = 0, one round-half-up at the end of the
// 8-bit filters, BORDER_REFLECT_101 at every stage, 15-bit RGB2GRAY coefficients, cvRound half-to-even in addWeighted,
// rows-first symmetric summation of the float64 filter): bit-identical to cv2 (tests/test_gpu_image.py against outputs of
// OpenCV itself, tests/golden/make_image_golden.py).  Byte / integer work, HBM bound: 12 B read + 12 B (4 B) written per pixel;
// one fused kernel per function, every intermediate image lives in shared memory.
#include <stdint.h>

#include "awb_internal.cuh"

namespace awb {

__device__ __forceinline__ int reflect101(int i, int n) {
  if (n == 1) return 0;
  if (i < 0) i = -i;                                        // the common cases: at most one reflection per side
  if (i >= n) i = 2 * n - 2 - i;
  while (i < 0 || i >= n) i = i < 0 ? -i : 2 * n - 2 - i;   // tiny images / positions of a partial tile far outside
  return i;
}
// BORDER = false: the tile and its halo lie inside the image, coordinates need no mapping
template <bool BORDER>
__device__ __forceinline__ int rmap(int i, int n) { return BORDER ? reflect101(i, n) : i; }
__device__ __forceinline__ int to_u8(float v) {            // (image * 255).astype(np.uint8)
  // truncation with saturation to [0, 255] in one conversion (NaN -> 0), as the numpy cast does for in-range values;
  // out-of-range inputs clamp (the reference's frames are in [0, 1])
  uint32_t q;
  asm("cvt.rzi.u8.f32 %0, %1;" : "=r"(q) : "f"(v * 255.0f));
  return (int)q;
}

// ---- _process_image: tile 64 x 16 outputs (4 per thread), 68 x 20 uint8 inputs per channel in shared memory, separable
constexpr int PT_W = 64, PT_H = 16;
template <bool BORDER>
__device__ __forceinline__ void image_process_tile(const float* __restrict__ img, float* __restrict__ out, int H, int W, int bgr,
                                                   unsigned char (*s)[PT_H + 4][PT_W + 4], unsigned short (*hs)[PT_H + 4][PT_W]) {
  const int x0 = blockIdx.x * PT_W, y0 = blockIdx.y * PT_H;
  const int64_t hw = (int64_t)H * W;
  for (int i = threadIdx.x; i < 3 * (PT_H + 4) * (PT_W + 4); i += 256) {
    const int c = i / ((PT_H + 4) * (PT_W + 4)), r = i % ((PT_H + 4) * (PT_W + 4));
    const int yy = rmap<BORDER>(y0 - 2 + r / (PT_W + 4), H), xx = rmap<BORDER>(x0 - 2 + r % (PT_W + 4), W);
    s[c][r / (PT_W + 4)][r % (PT_W + 4)] = (unsigned char)to_u8(img[c * hw + (int64_t)yy * W + xx]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 3 * (PT_H + 4) * PT_W; i += 256) {          // rows: [1 4 6 4 1], exact in 16 bit
    const int c = i / ((PT_H + 4) * PT_W), r = i % ((PT_H + 4) * PT_W), ry = r / PT_W, rx = r % PT_W;
    const unsigned char* p = &s[c][ry][rx];
    hs[c][ry][rx] = (unsigned short)(p[0] + 4 * p[1] + 6 * p[2] + 4 * p[3] + p[4]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 3 * PT_H * PT_W; i += 256) {                // columns + the one rounding, half up
    const int c = i / (PT_H * PT_W), r = i % (PT_H * PT_W), ry = r / PT_W, rx = r % PT_W;
    const int x = x0 + rx, y = y0 + ry;
    if (x >= W || y >= H) continue;
    const int acc = hs[c][ry][rx] + 4 * hs[c][ry + 1][rx] + 6 * hs[c][ry + 2][rx] + 4 * hs[c][ry + 3][rx] + hs[c][ry + 4][rx];
    out[(bgr ? 2 - c : c) * hw + (int64_t)y * W + x] = (float)((acc + 128) >> 8) / 255.0f;
  }
}

__global__ void __launch_bounds__(256) k_image_process(const float* __restrict__ img, float* __restrict__ out, int H, int W,
                                                       int do_blur, int bgr) {
  __shared__ unsigned char s[3][PT_H + 4][PT_W + 4];
  __shared__ unsigned short hs[3][PT_H + 4][PT_W];
  const int x0 = blockIdx.x * PT_W, y0 = blockIdx.y * PT_H;
  const int64_t hw = (int64_t)H * W;
  img += (int64_t)blockIdx.z * 3 * hw;                       // frame of the batch
  out += (int64_t)blockIdx.z * 3 * hw;
  if (!do_blur) {
    for (int i = threadIdx.x; i < 3 * PT_H * PT_W; i += 256) {
      const int c = i / (PT_H * PT_W), r = i % (PT_H * PT_W), x = x0 + r % PT_W, y = y0 + r / PT_W;
      if (x < W && y < H) out[(bgr ? 2 - c : c) * hw + (int64_t)y * W + x] = img[c * hw + (int64_t)y * W + x];
    }
    return;
  }
  const bool interior = x0 >= 2 && y0 >= 2 && x0 + PT_W + 2 <= W && y0 + PT_H + 2 <= H;
  if (interior) image_process_tile<false>(img, out, H, W, bgr, s, hs);
  else image_process_tile<true>(img, out, H, W, bgr, s, hs);
}

// ---- create_edge_map: tile 32 x 16 outputs; gray (halo 3), gradient magnitude (halo 2) and the row-filtered float64
// image live in shared memory.  Out-of-image coordinates are never stored: every read maps through reflect101 first,
// exactly as OpenCV extends each intermediate image at its border.
constexpr int ET_W = 32, ET_H = 16;
struct EdgeSmem {
  unsigned char src[3][ET_H + 8][ET_W + 8];
  unsigned char gray[ET_H + 6][ET_W + 6];
  unsigned char grad[ET_H + 4][ET_W + 4];
  double rows[ET_H + 4][ET_W];
};
template <bool BORDER>
__device__ __forceinline__ void edge_tile(const float* __restrict__ img, float* __restrict__ out, int H, int W, EdgeSmem& m) {
  const int x0 = blockIdx.x * ET_W, y0 = blockIdx.y * ET_H;
  const int64_t hw = (int64_t)H * W;
  auto inside = [&](int y, int x) { return !BORDER || (y >= 0 && y < H && x >= 0 && x < W); };
  // stage 0: (image * 255) -> uint8 on the region [x0-4, x0+ET_W+4) x [y0-4, y0+ET_H+4) (in-image part), read once
  for (int i = threadIdx.x; i < 3 * (ET_H + 8) * (ET_W + 8); i += 256) {
    const int c = i / ((ET_H + 8) * (ET_W + 8)), r = i % ((ET_H + 8) * (ET_W + 8));
    const int y = y0 - 4 + r / (ET_W + 8), x = x0 - 4 + r % (ET_W + 8);
    if (!inside(y, x)) continue;
    m.src[c][r / (ET_W + 8)][r % (ET_W + 8)] = (unsigned char)to_u8(img[c * hw + (int64_t)y * W + x]);
  }
  __syncthreads();
  auto S = [&](int c, int y, int x) { return (int)m.src[c][rmap<BORDER>(y, H) - (y0 - 4)][rmap<BORDER>(x, W) - (x0 - 4)]; };
  // stage 1: GaussianBlur 3x3 -> RGB2GRAY on the region [x0-3, x0+ET_W+3) x [y0-3, y0+ET_H+3) (in-image part)
  for (int i = threadIdx.x; i < (ET_H + 6) * (ET_W + 6); i += 256) {
    const int ry = i / (ET_W + 6), rx = i % (ET_W + 6);
    const int y = y0 - 3 + ry, x = x0 - 3 + rx;
    if (!inside(y, x)) continue;
    int ch[3];
#pragma unroll
    for (int c = 0; c < 3; c++) {
      int acc = 0;
#pragma unroll
      for (int dy = -1; dy <= 1; dy++) {
        int row = 0;
#pragma unroll
        for (int dx = -1; dx <= 1; dx++) row += (dx == 0 ? 2 : 1) * S(c, y + dy, x + dx);
        acc += (dy == 0 ? 2 : 1) * row;
      }
      ch[c] = (acc + 8) >> 4;
    }
    m.gray[ry][rx] = (unsigned char)((ch[0] * 9798 + ch[1] * 19235 + ch[2] * 3735 + 16384) >> 15);
  }
  __syncthreads();
  auto G = [&](int y, int x) { return (int)m.gray[rmap<BORDER>(y, H) - (y0 - 3)][rmap<BORDER>(x, W) - (x0 - 3)]; };
  // stage 2: Sobel x / y (16S), |.| saturated to 8 bit, addWeighted(.5, .5) with round-half-even
  for (int i = threadIdx.x; i < (ET_H + 4) * (ET_W + 4); i += 256) {
    const int ry = i / (ET_W + 4), rx = i % (ET_W + 4);
    const int y = y0 - 2 + ry, x = x0 - 2 + rx;
    if (!inside(y, x)) continue;
    const int a = G(y - 1, x - 1), b = G(y - 1, x), c = G(y - 1, x + 1);
    const int d = G(y, x - 1), f = G(y, x + 1);
    const int g = G(y + 1, x - 1), h = G(y + 1, x), k = G(y + 1, x + 1);
    int gx = (c + 2 * f + k) - (a + 2 * d + g), gy = (g + 2 * h + k) - (a + 2 * b + c);
    gx = gx < 0 ? -gx : gx; gy = gy < 0 ? -gy : gy;
    const int s = (gx > 255 ? 255 : gx) + (gy > 255 ? 255 : gy);
    m.grad[ry][rx] = (unsigned char)(s / 2 + ((s & 1) & ((s / 2) & 1)));
  }
  __syncthreads();
  auto Q = [&](int y, int x) { return (double)m.grad[rmap<BORDER>(y, H) - (y0 - 2)][rmap<BORDER>(x, W) - (x0 - 2)] / 255.0; };
  // stage 3: float64 GaussianBlur 5x5, rows first, symmetric summation order (k0*x0 + k1*(x-1 + x+1) + k2*(x-2 + x+2))
  const double k0 = 0.375, k1 = 0.25, k2 = 0.0625;
  for (int i = threadIdx.x; i < (ET_H + 4) * ET_W; i += 256) {
    const int ry = i / ET_W, rx = i % ET_W;
    const int y = y0 - 2 + ry, x = x0 + rx;
    if (BORDER && (y < 0 || y >= H || x >= W)) continue;
    m.rows[ry][rx] = k0 * Q(y, x) + k1 * (Q(y, x - 1) + Q(y, x + 1)) + k2 * (Q(y, x - 2) + Q(y, x + 2));
  }
  __syncthreads();
  auto R = [&](int y, int rx) { return m.rows[rmap<BORDER>(y, H) - (y0 - 2)][rx]; };
  for (int i = threadIdx.x; i < ET_H * ET_W; i += 256) {
    const int ry = i / ET_W, rx = i % ET_W;
    const int y = y0 + ry, x = x0 + rx;
    if (BORDER && (y >= H || x >= W)) continue;
    out[(int64_t)y * W + x] = (float)(k0 * R(y, rx) + k1 * (R(y - 1, rx) + R(y + 1, rx)) + k2 * (R(y - 2, rx) + R(y + 2, rx)));
  }
}

__global__ void __launch_bounds__(256) k_image_edge_map(const float* __restrict__ img, float* __restrict__ out, int H, int W) {
  __shared__ EdgeSmem m;
  const int x0 = blockIdx.x * ET_W, y0 = blockIdx.y * ET_H;
  const int64_t hw = (int64_t)H * W;
  img += (int64_t)blockIdx.z * 3 * hw;                       // frame of the batch
  out += (int64_t)blockIdx.z * hw;
  const bool interior = x0 >= 4 && y0 >= 4 && x0 + ET_W + 4 <= W && y0 + ET_H + 4 <= H;
  if (interior) edge_tile<false>(img, out, H, W, m);
  else edge_tile<true>(img, out, H, W, m);
}

// ======================================================================= streaming versions (W % 4 == 0, W >= 128, H >= 8)
// The tiled kernels above spend their time on index arithmetic and byte-wide shared-memory traffic (21 % / 7 % of the HBM
// roofline).  These walk the image top to bottom instead: a lane owns 4 neighbouring pixels (one LDG.128 per channel and
// row, packed to one 32-bit word of 4 uint8), the horizontal taps come from the neighbouring lanes' words (2 shuffles per
// row and stage), the vertical taps from a register ring of the last 3 / 5 filtered rows -- no shared memory apart from
// the 256-entry division tables, no barrier in the row loop.  A warp covers a window of 128 columns, of which the outer
// lanes are the horizontal halo (4 pixels per side: the sum of the stages' radii), i.e. 120 output columns.
// Borders: every stage of the OpenCV chain extends ITS input by BORDER_REFLECT_101.  All kernels of the chain are symmetric
// (or antisymmetric with an absolute value behind them), so an intermediate image of the reflect-extended source is itself
// reflect-symmetric: reflecting once, at the source fetch (virtual rows / columns outside the image read pixel reflect101(.)),
// reproduces the per-stage extension exactly -- the row loop has no border cases at all.
constexpr int ST_COLS = 120;      // output columns per warp
// reflect101 without the loop: one reflection per side is enough for -n < i < 2n - 1 (the streaming kernels run on images of
// at least 8 x 128 pixels, where their halo rows and the columns of the last window stay inside that range)
__device__ __forceinline__ int reflect1(int i, int n) {
  i = i < 0 ? -i : i;
  return i >= n ? 2 * n - 2 - i : i;
}
__device__ __forceinline__ float4 load_raw4(const float* __restrict__ row, int xv, int W) {
  if (xv >= 0 && xv + 3 < W) return __ldg(reinterpret_cast<const float4*>(row + xv));
  return make_float4(row[reflect1(xv, W)], row[reflect1(xv + 1, W)], row[reflect1(xv + 2, W)], row[reflect1(xv + 3, W)]);
}
__device__ __forceinline__ uint32_t pack_u8x4(float4 v) {
  return (uint32_t)to_u8(v.x) | ((uint32_t)to_u8(v.y) << 8) | ((uint32_t)to_u8(v.z) << 16) | ((uint32_t)to_u8(v.w) << 24);
}
// the 8 pixels x-2 .. x+5 around a lane's word from its neighbours' words
__device__ __forceinline__ void neighbours(uint32_t w, int* p) {
  const uint32_t l = __shfl_up_sync(0xffffffffu, w, 1), r = __shfl_down_sync(0xffffffffu, w, 1);
  p[0] = (l >> 16) & 255; p[1] = l >> 24;
  p[2] = w & 255; p[3] = (w >> 8) & 255; p[4] = (w >> 16) & 255; p[5] = w >> 24;
  p[6] = r & 255; p[7] = (r >> 8) & 255;
}

// the 4 byte windows (x-1, x, x+1, x+2) of a lane's 4 pixels, for 3-tap filters as one dp4a each (4th coefficient 0)
__device__ __forceinline__ void windows3(uint32_t w, uint32_t* win) {
  const uint32_t l = __shfl_up_sync(0xffffffffu, w, 1), r = __shfl_down_sync(0xffffffffu, w, 1);
  win[0] = __byte_perm(l, w, 0x6543);
  win[1] = w;
  win[2] = __byte_perm(w, r, 0x4321);
  win[3] = __byte_perm(w, r, 0x5432);
}
__device__ __forceinline__ int dp4a_us(uint32_t a, uint32_t b_signed, int c) {      // unsigned bytes x signed bytes + c
  int d;
  asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b_signed), "r"(c));
  return d;
}

// _process_image: grid (windows, bands / 4, frames * 3); a warp = one band of R rows of one channel
__global__ void __launch_bounds__(128) k_image_process_stream(const float* __restrict__ img, float* __restrict__ out, int H, int W,
                                                              int R, int bgr) {
  __shared__ float lut[256];
  for (int i = threadIdx.x; i < 256; i += 128) lut[i] = (float)i / 255.0f;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int yb = (blockIdx.y * 4 + warp) * R;
  if (yb >= H) return;
  const int f = blockIdx.z / 3, c = blockIdx.z % 3;
  const int64_t hw = (int64_t)H * W;
  const float* src = img + ((int64_t)f * 3 + c) * hw;
  float* dst = out + ((int64_t)f * 3 + (bgr ? 2 - c : c)) * hw;
  const int xv = (int)blockIdx.x * ST_COLS - 4 + lane * 4;          // first of this lane's 4 (virtual) columns
  const bool store = lane >= 1 && lane <= 30 && xv < W;
  const int y_end = yb + R < H ? yb + R : H;
  int ring[4][4];                                                   // horizontally filtered rows v-4 .. v-1
#pragma unroll
  for (int k = 0; k < 4; k++) { ring[k][0] = ring[k][1] = ring[k][2] = ring[k][3] = 0; }
  // rows v+1 .. v+3 are in flight (raw, converted only when their turn comes) while row v is filtered
  constexpr int DEPTH = 3;
  float4 raw[DEPTH];
#pragma unroll
  for (int k = 0; k < DEPTH; k++) raw[k] = load_raw4(src + (int64_t)reflect1(yb - 2 + k, H) * W, xv, W);
#pragma unroll 4      // the ring of 4 rows turns into register renaming
  for (int v = yb - 2; v < y_end + 2; v++) {
    const uint32_t w = pack_u8x4(raw[0]);
#pragma unroll
    for (int k = 0; k + 1 < DEPTH; k++) raw[k] = raw[k + 1];
    raw[DEPTH - 1] = load_raw4(src + (int64_t)reflect1(v + DEPTH, H) * W, xv, W);
    // [1 4 6 4 1], exact: four taps as one dp4a on the byte window (x-2 .. x+1) of the pixel, the fifth added
    int h[4];
    {
      const uint32_t l = __shfl_up_sync(0xffffffffu, w, 1), r = __shfl_down_sync(0xffffffffu, w, 1);
      h[0] = (int)__dp4a(__byte_perm(l, w, 0x5432), 0x04060401u, (w >> 16) & 255u);
      h[1] = (int)__dp4a(__byte_perm(l, w, 0x6543), 0x04060401u, w >> 24);
      h[2] = (int)__dp4a(w, 0x04060401u, r & 255u);
      h[3] = (int)__dp4a(__byte_perm(w, r, 0x4321), 0x04060401u, (r >> 8) & 255u);
    }
    const int y = v - 2;
    if (y >= yb && store) {
      float o[4];
#pragma unroll
      for (int i = 0; i < 4; i++) {
        const int acc = ring[0][i] + 4 * ring[1][i] + 6 * ring[2][i] + 4 * ring[3][i] + h[i];
        o[i] = lut[(acc + 128) >> 8];                               // the one rounding (half up), then / 255
      }
      *reinterpret_cast<float4*>(dst + (int64_t)y * W + xv) = make_float4(o[0], o[1], o[2], o[3]);
    }
#pragma unroll
    for (int i = 0; i < 4; i++) { ring[0][i] = ring[1][i]; ring[1][i] = ring[2][i]; ring[2][i] = ring[3][i]; ring[3][i] = h[i]; }
  }
}

// create_edge_map: grid (windows, bands / 4, frames); a warp = one band of R output rows
__global__ void __launch_bounds__(128, 4) k_image_edge_stream(const float* __restrict__ img, float* __restrict__ out, int H, int W, int R) {
  __shared__ double lutd[256];
  for (int i = threadIdx.x; i < 256; i += 128) lutd[i] = (double)i / 255.0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int yb = (blockIdx.y * 4 + warp) * R;
  if (yb >= H) return;
  const int64_t hw = (int64_t)H * W;
  const float* src = img + (int64_t)blockIdx.z * 3 * hw;
  float* dst = out + (int64_t)blockIdx.z * hw;
  const int xv = (int)blockIdx.x * ST_COLS - 4 + lane * 4;
  const bool store = lane >= 1 && lane <= 30 && xv < W;
  const int y_end = yb + R < H ? yb + R : H;
  const double k0 = 0.375, k1 = 0.25, k2 = 0.0625;
  // rings, newest last.  A: [1 2 1]-filtered source rows (3 channels), G: gray rows as 3-pixel windows,
  // D: row-filtered float64 gradient rows
  int ra[2][3][4];
  uint32_t rg[2][4];
  double rd[4][4];
#pragma unroll
  for (int k = 0; k < 2; k++) {
#pragma unroll
    for (int i = 0; i < 4; i++) rg[k][i] = 0u;
#pragma unroll
    for (int cc = 0; cc < 3; cc++) { ra[k][cc][0] = ra[k][cc][1] = ra[k][cc][2] = ra[k][cc][3] = 0; }
  }
#pragma unroll
  for (int k = 0; k < 4; k++) { rd[k][0] = rd[k][1] = rd[k][2] = rd[k][3] = 0.0; }
  constexpr int DEPTH = 2;                              // source rows in flight (raw), 3 channels each
  float4 raw[DEPTH][3];
#pragma unroll
  for (int k = 0; k < DEPTH; k++) {
    const int64_t ro = (int64_t)reflect1(yb - 4 + k, H) * W;
#pragma unroll
    for (int cc = 0; cc < 3; cc++) raw[k][cc] = load_raw4(src + cc * hw + ro, xv, W);
  }
#pragma unroll 2      // the rings of 2 rows turn into register renaming (4 iterations would spill under 128 registers)
  for (int v = yb - 4; v < y_end + 4; v++) {            // source row v -> blurred / gray row v-1 -> gradient row v-2 -> output row v-4
    uint32_t w[3];
#pragma unroll
    for (int cc = 0; cc < 3; cc++) w[cc] = pack_u8x4(raw[0][cc]);
    {
      const int64_t ro = (int64_t)reflect1(v + DEPTH, H) * W;
#pragma unroll
      for (int cc = 0; cc < 3; cc++) {
#pragma unroll
        for (int k = 0; k + 1 < DEPTH; k++) raw[k][cc] = raw[k + 1][cc];
        raw[DEPTH - 1][cc] = load_raw4(src + cc * hw + ro, xv, W);
      }
    }
    // stage 1: GaussianBlur 3x3 per channel (rows of [1 2 1] now -- one dp4a per pixel --, columns from the ring), RGB2GRAY
    int ha[3][4];
#pragma unroll
    for (int cc = 0; cc < 3; cc++) {
      uint32_t win[4];
      windows3(w[cc], win);
#pragma unroll
      for (int i = 0; i < 4; i++) ha[cc][i] = (int)__dp4a(win[i], 0x00010201u, 0u);
    }
    uint32_t gw = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
      int ch[3];
#pragma unroll
      for (int cc = 0; cc < 3; cc++) ch[cc] = (ra[0][cc][i] + 2 * ra[1][cc][i] + ha[cc][i] + 8) >> 4;
      gw |= (uint32_t)((ch[0] * 9798 + ch[1] * 19235 + ch[2] * 3735 + 16384) >> 15) << (8 * i);
    }
#pragma unroll
    for (int cc = 0; cc < 3; cc++) {
#pragma unroll
      for (int i = 0; i < 4; i++) { ra[0][cc][i] = ra[1][cc][i]; ra[1][cc][i] = ha[cc][i]; }
    }
    // gray row v-1 as windows (x-1, x, x+1) per pixel
    uint32_t g[4];
    windows3(gw, g);
    // stage 2: Sobel on gray rows v-3, v-2, v-1 -> gradient row v-2; five dp4a per pixel:
    //   gx = top . (-1, 0, 1) + mid . (-2, 0, 2) + bottom . (-1, 0, 1),  gy = bottom . (1, 2, 1) - top . (1, 2, 1)
    uint32_t dw = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
      int gx = dp4a_us(rg[0][i], 0x000100FFu, dp4a_us(rg[1][i], 0x000200FEu, dp4a_us(g[i], 0x000100FFu, 0)));
      int gy = dp4a_us(g[i], 0x00010201u, dp4a_us(rg[0][i], 0x00FFFEFFu, 0));
      gx = gx < 0 ? -gx : gx; gy = gy < 0 ? -gy : gy;
      const int sm = (gx > 255 ? 255 : gx) + (gy > 255 ? 255 : gy);
      dw |= (uint32_t)(sm / 2 + ((sm & 1) & ((sm / 2) & 1))) << (8 * i);      // addWeighted(.5, .5): round half to even
    }
#pragma unroll
    for (int i = 0; i < 4; i++) { rg[0][i] = rg[1][i]; rg[1][i] = g[i]; }
    // stage 3: float64 GaussianBlur 5x5 of gradient / 255: row v-2 filtered now, columns from the ring -> output row v-4
    double hd[4];
    {
      int p[8];
      neighbours(dw, p);
      double q[8];
#pragma unroll
      for (int i = 0; i < 8; i++) q[i] = lutd[p[i]];
#pragma unroll
      for (int i = 0; i < 4; i++) hd[i] = k0 * q[i + 2] + k1 * (q[i + 1] + q[i + 3]) + k2 * (q[i] + q[i + 4]);
    }
    const int y = v - 4;
    if (y >= yb && store) {
      float o[4];
#pragma unroll
      for (int i = 0; i < 4; i++) o[i] = (float)(k0 * rd[2][i] + k1 * (rd[1][i] + rd[3][i]) + k2 * (rd[0][i] + hd[i]));
      *reinterpret_cast<float4*>(dst + (int64_t)y * W + xv) = make_float4(o[0], o[1], o[2], o[3]);
    }
#pragma unroll
    for (int i = 0; i < 4; i++) { rd[0][i] = rd[1][i]; rd[1][i] = rd[2][i]; rd[2][i] = rd[3][i]; rd[3][i] = hd[i]; }
  }
}

// rows per warp: long bands amortise the 4 / 8 halo rows, short ones keep a small batch spread over the SMs
static int stream_band_rows(int H, int W, int n_frames, int planes, int min_warps) {
  const int64_t windows = (W + ST_COLS - 1) / ST_COLS;
  int R = 128;
  while (R > 8 && windows * ((H + R - 1) / R) * n_frames * planes < min_warps) R >>= 1;
  return R;
}

}  // namespace awb

using namespace awb;

extern "C" {

int awb_image_process(const float* image, float* out, int32_t n_frames, int32_t H, int32_t W, int32_t do_blur, int32_t bgr,
                      void* stream) {
  if (!image || !out || image == out) { set_error("image / out must be distinct non-null device pointers"); return AWB_ERR_INVALID; }
  if (H < 1 || W < 1 || n_frames < 1 || n_frames > 65535) { set_error("bad batch %d x %dx%d", n_frames, H, W); return AWB_ERR_INVALID; }
  cudaStream_t st = (cudaStream_t)stream;
  if (do_blur && W % 4 == 0 && W >= 128 && H >= 8 && ((uintptr_t)image & 15) == 0 && ((uintptr_t)out & 15) == 0 && n_frames * 3 <= 65535) {
    const int R = stream_band_rows(H, W, n_frames, 3, 148 * 48);
    dim3 sgrid((W + ST_COLS - 1) / ST_COLS, ((H + R - 1) / R + 3) / 4, n_frames * 3);
    AWB_LAUNCH(PK_MISC, st, k_image_process_stream<<<sgrid, 128, 0, st>>>(image, out, H, W, R, bgr));
    AWB_CUDA(cudaGetLastError());
    return AWB_OK;
  }
  dim3 grid((W + PT_W - 1) / PT_W, (H + PT_H - 1) / PT_H, n_frames);
  AWB_LAUNCH(PK_MISC, st, k_image_process<<<grid, 256, 0, st>>>(image, out, H, W, do_blur, bgr));
  AWB_CUDA(cudaGetLastError());
  return AWB_OK;
}

int awb_image_edge_map(const float* image, float* out, int32_t n_frames, int32_t H, int32_t W, void* stream) {
  if (!image || !out) { set_error("null argument"); return AWB_ERR_INVALID; }
  if (H < 1 || W < 1 || n_frames < 1 || n_frames > 65535) { set_error("bad batch %d x %dx%d", n_frames, H, W); return AWB_ERR_INVALID; }
  cudaStream_t st = (cudaStream_t)stream;
  if (W % 4 == 0 && W >= 128 && H >= 8 && ((uintptr_t)image & 15) == 0 && ((uintptr_t)out & 15) == 0) {
    const int R = stream_band_rows(H, W, n_frames, 1, 148 * 24);
    dim3 sgrid((W + ST_COLS - 1) / ST_COLS, ((H + R - 1) / R + 3) / 4, n_frames);
    AWB_LAUNCH(PK_MISC, st, k_image_edge_stream<<<sgrid, 128, 0, st>>>(image, out, H, W, R));
    AWB_CUDA(cudaGetLastError());
    return AWB_OK;
  }
  dim3 grid((W + ET_W - 1) / ET_W, (H + ET_H - 1) / ET_H, n_frames);
  AWB_LAUNCH(PK_MISC, st, k_image_edge_map<<<grid, 256, 0, st>>>(image, out, H, W));
  AWB_CUDA(cudaGetLastError());
  return AWB_OK;
}

}  // extern "C"
