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
  int q = __float2int_rz(v * 255.0f);
  return q < 0 ? 0 : (q > 255 ? 255 : q);
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

}  // namespace awb

using namespace awb;

extern "C" {

int awb_image_process(const float* image, float* out, int32_t n_frames, int32_t H, int32_t W, int32_t do_blur, int32_t bgr,
                      void* stream) {
  if (!image || !out || image == out) { set_error("image / out must be distinct non-null device pointers"); return AWB_ERR_INVALID; }
  if (H < 1 || W < 1 || n_frames < 1 || n_frames > 65535) { set_error("bad batch %d x %dx%d", n_frames, H, W); return AWB_ERR_INVALID; }
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid((W + PT_W - 1) / PT_W, (H + PT_H - 1) / PT_H, n_frames);
  AWB_LAUNCH(PK_MISC, st, k_image_process<<<grid, 256, 0, st>>>(image, out, H, W, do_blur, bgr));
  AWB_CUDA(cudaGetLastError());
  return AWB_OK;
}

int awb_image_edge_map(const float* image, float* out, int32_t n_frames, int32_t H, int32_t W, void* stream) {
  if (!image || !out) { set_error("null argument"); return AWB_ERR_INVALID; }
  if (H < 1 || W < 1 || n_frames < 1 || n_frames > 65535) { set_error("bad batch %d x %dx%d", n_frames, H, W); return AWB_ERR_INVALID; }
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid((W + ET_W - 1) / ET_W, (H + ET_H - 1) / ET_H, n_frames);
  AWB_LAUNCH(PK_MISC, st, k_image_edge_map<<<grid, 256, 0, st>>>(image, out, H, W));
  AWB_CUDA(cudaGetLastError());
  return AWB_OK;
}

}  // extern "C"
