// Star-shape prior: the network ``myNet`` of notebooks/icml_teaser_code/star_shaped/star.ipynb cell 2
//   x <- x + offset;  r = ||x||;  u = x / (0.01 + r);  a = relu(W0 u + b0);
//   b = relu(W1 a + b1 + W1_r r + b1_r);  y = r * (W2 a + b2 + W2_r b + b2_r) - 1
// and the body of its training loop (cell 3): sigmoid + MSE on a sampled point set, backward, Adam, clamp of
// W2_r.weight.  The point sets are small (1000 samples per step in the notebook), so the step is latency bound:
// ONE fused kernel does forward + loss + backward for its slice of the points with one thread per hidden unit,
// W1 and its gradient accumulator resident in shared memory ((h+1)-padded rows: row and column walks are both
// bank-conflict free), cross-thread sums by warp shuffles; per-CTA partials are reduced in a fixed order by the
// shared optimizer kernel (no atomics).
#include <math.h>

#include "awb_internal.cuh"

namespace awb {

struct StarP {
  const float* params; int64_t P; int h;
  const float* x;        // [n][2]
  const float* target;   // [n] (fit)
  int64_t n, chunk;
  awb_loss_spec loss;
  float* logits;         // [n] or null
  float* part;           // [S][P] (fit)
  float* lossp;          // [S]
};

// sum of up to 3 values over the whole block (blockDim.x = 160 = 5 warps); result in every thread
__device__ __forceinline__ void block_sum3(float& a, float& b, float& c, float (*red)[3]) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, off);
    b += __shfl_xor_sync(0xffffffffu, b, off);
    c += __shfl_xor_sync(0xffffffffu, c, off);
  }
  const int w = threadIdx.x >> 5;
  __syncthreads();                       // previous readers of red are done
  if ((threadIdx.x & 31) == 0) { red[w][0] = a; red[w][1] = b; red[w][2] = c; }
  __syncthreads();
  a = b = c = 0.f;
  for (int i = 0; i < (int)(blockDim.x >> 5); i++) { a += red[i][0]; b += red[i][1]; c += red[i][2]; }   // fixed order
}

template <bool FIT>
__global__ void __launch_bounds__(160) k_star(StarP p) {
  extern __shared__ float sm[];
  const int h = p.h, ld = h + 1;
  float* sW1 = sm;                       // [h][h+1]
  float* sG = sm + h * ld;               // [h][h+1] gradient accumulator (FIT)
  float* sa = sG + (FIT ? h * ld : 0);   // [h]
  float* sdb = sa + h;                   // [h]
  __shared__ float red[8][3];
  const int j = threadIdx.x;
  const bool act = j < h;
  const float* par = p.params;
  const int o_W0 = 2, o_b0 = 2 + 2 * h, o_W1 = 2 + 3 * h, o_b1 = o_W1 + h * h, o_W2 = o_b1 + h, o_b2 = o_W2 + h,
            o_W1r = o_b2 + 1, o_b1r = o_W1r + h, o_W2r = o_b1r + h, o_b2r = o_W2r + h;
  for (int i = threadIdx.x; i < h * h; i += blockDim.x) { sW1[(i / h) * ld + (i % h)] = par[o_W1 + i]; }
  if (FIT) for (int i = threadIdx.x; i < h * ld; i += blockDim.x) sG[i] = 0.f;
  const float off0 = par[0], off1 = par[1];
  float w0x = 0.f, w0y = 0.f, b0 = 0.f, w1r = 0.f, bb1 = 0.f, w2 = 0.f, w2r = 0.f;
  if (act) {
    w0x = par[o_W0 + 2 * j]; w0y = par[o_W0 + 2 * j + 1]; b0 = par[o_b0 + j];
    w1r = par[o_W1r + j]; bb1 = par[o_b1 + j] + par[o_b1r + j];
    w2 = par[o_W2 + j]; w2r = par[o_W2r + j];
  }
  const float bias2 = par[o_b2] + par[o_b2r];
  float g_w0x = 0.f, g_w0y = 0.f, g_b0 = 0.f, g_w1r = 0.f, g_b1 = 0.f, g_w2 = 0.f, g_w2r = 0.f;
  float g_b2 = 0.f, g_ox = 0.f, g_oy = 0.f, lacc = 0.f;
  __syncthreads();
  const int64_t r0 = (int64_t)blockIdx.x * p.chunk, r1 = r0 + p.chunk < p.n ? r0 + p.chunk : p.n;
  for (int64_t pt = r0; pt < r1; pt++) {
    const float px = p.x[2 * pt] + off0, py = p.x[2 * pt + 1] + off1;
    const float r = sqrtf(px * px + py * py), q = 0.01f + r;
    const float ux = px / q, uy = py / q;
    const float apre = act ? fmaf(w0x, ux, fmaf(w0y, uy, b0)) : 0.f;
    const float a = fmaxf(apre, 0.f);
    if (act) sa[j] = a;
    __syncthreads();
    float bpre = 0.f;
    if (act) {
      bpre = fmaf(w1r, r, bb1);
      const float* wr = sW1 + j * ld;
#pragma unroll 5
      for (int k = 0; k < h; k++) bpre = fmaf(wr[k], sa[k], bpre);
    }
    const float b = fmaxf(bpre, 0.f);
    float s = act ? fmaf(w2, a, w2r * b) : 0.f, z1 = 0.f, z2 = 0.f;
    block_sum3(s, z1, z2, red);
    s += bias2;
    const float y = fmaf(r, s, -1.f);
    if (p.logits && j == 0) p.logits[pt] = y;
    if (!FIT) continue;
    const float t = p.target[pt];
    const bool fg = p.loss.cls_rule == AWB_CLS_UNARY_LT_HALF ? (t < 0.5f) : (t != 1.0f);
    const float coef = fg ? p.loss.coef_fg : p.loss.coef_bg;
    const float sg = 1.f / (1.f + expf(-y));
    float l, dl;
    if (p.loss.kind == AWB_LOSS_SE_SIGMOID) { float d = t - sg; l = d * d; dl = -2.f * d * sg * (1.f - sg); }
    else { l = fmaxf(y, 0.f) - y * t + log1pf(expf(-fabsf(y))); dl = sg - t; }
    const float dy = coef * dl, dyr = dy * r;
    float db = 0.f;
    if (act) {
      g_w2 = fmaf(dyr, a, g_w2); g_w2r = fmaf(dyr, b, g_w2r);
      db = bpre > 0.f ? dyr * w2r : 0.f;
      g_b1 += db; g_w1r = fmaf(db, r, g_w1r);
      sdb[j] = db;
      float* gr = sG + j * ld;
#pragma unroll 5
      for (int k = 0; k < h; k++) gr[k] = fmaf(db, sa[k], gr[k]);
    }
    __syncthreads();
    float dapre = 0.f;
    if (act) {
      float da = dyr * w2;
#pragma unroll 5
      for (int i = 0; i < h; i++) da = fmaf(sdb[i], sW1[i * ld + j], da);
      dapre = apre > 0.f ? da : 0.f;
      g_w0x = fmaf(dapre, ux, g_w0x); g_w0y = fmaf(dapre, uy, g_w0y); g_b0 += dapre;
    }
    float dux = dapre * w0x, duy = dapre * w0y, drb = db * w1r;
    block_sum3(dux, duy, drb, red);
    if (j == 0) {
      lacc = fmaf(coef, l, lacc);
      g_b2 += dyr;
      if (r > 0.f) {
        const float dr = fmaf(dy, s, drb) - (dux * px + duy * py) / (q * q);
        g_ox += dux / q + dr * (px / r);
        g_oy += duy / q + dr * (py / r);
      }
    }
    __syncthreads();     // sa / sdb are rewritten by the next point
  }
  if (!FIT) return;
  __syncthreads();
  float* out = p.part + (int64_t)blockIdx.x * p.P;
  if (act) {
    out[o_W0 + 2 * j] = g_w0x; out[o_W0 + 2 * j + 1] = g_w0y; out[o_b0 + j] = g_b0;
    out[o_b1 + j] = g_b1; out[o_b1r + j] = g_b1; out[o_W1r + j] = g_w1r;
    out[o_W2 + j] = g_w2; out[o_W2r + j] = g_w2r;
  }
  if (j == 0) { out[0] = g_ox; out[1] = g_oy; out[o_b2] = g_b2; out[o_b2r] = g_b2; p.lossp[blockIdx.x] = lacc; }
  for (int i = threadIdx.x; i < h * h; i += blockDim.x) out[o_W1 + i] = sG[(i / h) * ld + (i % h)];
}

int star_n_ctas(int64_t n) {
  int64_t c = (n + 3) / 4;
  return (int)(c < 1 ? 1 : (c > kMaxSplits ? kMaxSplits : c));
}

int star_run(const awb_prior* h, const float* params, const float* x, const float* target, int64_t n,
             const awb_loss_spec* loss, float* logits, float* part, float* lossp, bool fit, int* n_ctas, cudaStream_t st) {
  const int hh = h->lay.h;
  if (hh > 160) { set_error("star prior supports n_hidden <= 160, got %d", hh); return AWB_ERR_UNSUPPORTED; }
  StarP p = {};
  p.params = params; p.P = h->lay.P; p.h = hh; p.x = x; p.target = target; p.n = n;
  const int S = fit ? star_n_ctas(n) : (int)((n + 15) / 16 < 1184 ? (n + 15) / 16 : 1184);
  p.chunk = (n + S - 1) / S;
  if (loss) p.loss = *loss;
  p.logits = logits; p.part = part; p.lossp = lossp;
  const size_t smem = sizeof(float) * ((size_t)(fit ? 2 : 1) * hh * (hh + 1) + 2 * hh);
  if (fit) {
    AWB_CUDA(cudaFuncSetAttribute(k_star<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    AWB_LAUNCH(PK_MISC, st, k_star<true><<<S, 160, smem, st>>>(p));
  } else {
    AWB_CUDA(cudaFuncSetAttribute(k_star<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    AWB_LAUNCH(PK_MISC, st, k_star<false><<<S, 160, smem, st>>>(p));
  }
  AWB_CUDA(cudaGetLastError());
  if (n_ctas) *n_ctas = S;
  return AWB_OK;
}

}  // namespace awb
