// RealNVP path-connectedness flow: grouped 1x1 conv on the coordinates
// (path_connected_net.py:65,82), MinMax to [-1,1] (min_max.py:9-19, norm_net.py:17-27), F x
// { MaskedAffineFlow(b, t, s); ActNorm } with s,t = MLP([C,m,C], LeakyReLU(0), tanh) (normflows
// 1.7.3 published semantics; built at net_factory.py:101-113), inverse
// MinMax.  K = C = 2..3 is too thin for tensor cores: CUDA-core FMAs, everything per pixel stays in
// registers, weights are staged once per CTA in shared memory.
//
// forward : one thread per pixel (no cross-pixel reduction).  Writes the deformed coordinates
//           X[n] = (x, y, t, 1) consumed by the ICNN and, when training, the input of every
//           coupling (F*C floats per pixel) so that the backward pass needs no forward sweep.
// backward: one lane per hidden unit (m <= 32), 32/gs pixels per warp: hidden-unit gradients
//           accumulate lane-locally into a private shared-memory buffer per lane group (no
//           atomics), outputs use log2(gs) warp shuffles.  Buffers are reduced in a fixed order.
#include <math.h>

#include "awb_internal.cuh"

namespace awb {

struct FlowP {
  GridDev g;
  FlowConsts fc;
  const float* params;   // arena, object 0
  int64_t P, off_flow, P_flow, per_flow, off_lin;
  int C, F, m, tanh_out, use_linear;
  int64_t N;
  float* X;              // [O][N][4]
  float* zin;            // [O][N][F*C] or null
  float* deformed;       // [O][N][C] or null
  const float* dX;       // backward: [O][N][4]
  float* fpart;          // backward: [S][O][PF]
  int64_t chunk; int O;
};

__device__ __forceinline__ float mm_fwd(float v, float vmin, float vmax, float nmin, float nmax) {
  return (v - vmin) / (vmax - vmin) * (nmax - nmin) + nmin;
}

template <int C>
__global__ void __launch_bounds__(256) k_flow_fwd(FlowP p) {
  extern __shared__ float sp[];   // [P_flow + 2C]
  const int o = blockIdx.y;
  const float* par = p.params + (int64_t)o * p.P + p.off_flow;
  const int PF = (int)p.P_flow + 2 * C;
  for (int i = threadIdx.x; i < PF; i += blockDim.x) sp[i] = par[i];
  __syncthreads();
  int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= p.N) return;
  const float* lin = sp + p.P_flow;
  float z[C];
#pragma unroll
  for (int c = 0; c < C; c++) {
    float x = coord(p.g, n, c);
    if (p.use_linear) x = x * lin[c] + lin[C + c];
    z[c] = mm_fwd(x, p.fc.nmin[c], p.fc.nmax[c], p.fc.new_min, p.fc.new_max);
  }
  const int m = p.m;
  const int half = 2 * m * C + m + C;
  float* zin = p.zin ? p.zin + ((int64_t)o * p.N + n) * (p.F * C) : nullptr;
  for (int f = 0; f < p.F; f++) {
    const float* w = sp + (int64_t)f * p.per_flow;
    if (zin) {
#pragma unroll
      for (int c = 0; c < C; c++) zin[f * C + c] = z[c];
    }
    float zm[C];
    bool b[C];
#pragma unroll
    for (int c = 0; c < C; c++) { b[c] = p.fc.masks[f * C + c] != 0; zm[c] = b[c] ? z[c] : 0.f; }
    float so[C], to[C];
#pragma unroll
    for (int c = 0; c < C; c++) { so[c] = w[m * C + m + C * m + c]; to[c] = w[half + m * C + m + C * m + c]; }
    for (int k = 0; k < m; k++) {
      float ps = w[m * C + k], pt = w[half + m * C + k];
#pragma unroll
      for (int c = 0; c < C; c++) { ps = fmaf(w[k * C + c], zm[c], ps); pt = fmaf(w[half + k * C + c], zm[c], pt); }
      float hs = fmaxf(ps, 0.f), ht = fmaxf(pt, 0.f);
#pragma unroll
      for (int c = 0; c < C; c++) {
        so[c] = fmaf(w[m * C + m + c * m + k], hs, so[c]);
        to[c] = fmaf(w[half + m * C + m + c * m + k], ht, to[c]);
      }
    }
    const float* an = w + 2 * half;
#pragma unroll
    for (int c = 0; c < C; c++) {
      float s = p.tanh_out ? tanhf(so[c]) : so[c];
      float t = p.tanh_out ? tanhf(to[c]) : to[c];
      if (!isfinite(s)) s = NAN;
      if (!isfinite(t)) t = NAN;
      float zc = b[c] ? z[c] : fmaf(z[c], expf(s), t);
      z[c] = fmaf(zc, expf(an[c]), an[C + c]);      // ActNorm
    }
  }
  float xd[3] = {0.f, 0.f, 0.f};
#pragma unroll
  for (int c = 0; c < C; c++) xd[c] = mm_fwd(z[c], p.fc.new_min, p.fc.new_max, p.fc.nmin[c], p.fc.nmax[c]);
  *reinterpret_cast<float4*>(p.X + ((int64_t)o * p.N + n) * 4) = make_float4(xd[0], xd[1], xd[2], 1.f);
  if (p.deformed) {
#pragma unroll
    for (int c = 0; c < C; c++) p.deformed[((int64_t)o * p.N + n) * C + c] = xd[c];
  }
}

// ---------------------------------------------------------------- inverse (PathConnectedNet.inverse, path_connected_net.py:107-122)
// x (a get_deformation output) -> MinMax -> flows in reverse order, each inverted (ActNorm: (z - t) exp(-s);
// coupling: zm + (1 - b)(z - t(zm)) exp(-s(zm)), normflows MaskedAffineFlow.inverse) -> inverse MinMax -> inverse 1x1 conv.
template <int C>
__global__ void __launch_bounds__(256) k_flow_inv(FlowP p, float* out) {
  extern __shared__ float sp[];   // [P_flow + 2C]
  const int o = blockIdx.y;
  const float* par = p.params + (int64_t)o * p.P + p.off_flow;
  const int PF = (int)p.P_flow + 2 * C;
  for (int i = threadIdx.x; i < PF; i += blockDim.x) sp[i] = par[i];
  __syncthreads();
  int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= p.N) return;
  const float* lin = sp + p.P_flow;
  float z[C];
#pragma unroll
  for (int c = 0; c < C; c++) z[c] = mm_fwd(coord(p.g, n, c), p.fc.nmin[c], p.fc.nmax[c], p.fc.new_min, p.fc.new_max);
  const int m = p.m;
  const int half = 2 * m * C + m + C;
  for (int f = p.F - 1; f >= 0; f--) {
    const float* w = sp + (int64_t)f * p.per_flow;
    const float* an = w + 2 * half;
    float zm[C];
    bool b[C];
#pragma unroll
    for (int c = 0; c < C; c++) {
      z[c] = (z[c] - an[C + c]) * expf(-an[c]);          // ActNorm inverse
      b[c] = p.fc.masks[f * C + c] != 0; zm[c] = b[c] ? z[c] : 0.f;
    }
    float so[C], to[C];
#pragma unroll
    for (int c = 0; c < C; c++) { so[c] = w[m * C + m + C * m + c]; to[c] = w[half + m * C + m + C * m + c]; }
    for (int k = 0; k < m; k++) {
      float ps = w[m * C + k], pt = w[half + m * C + k];
#pragma unroll
      for (int c = 0; c < C; c++) { ps = fmaf(w[k * C + c], zm[c], ps); pt = fmaf(w[half + k * C + c], zm[c], pt); }
      float hs = fmaxf(ps, 0.f), ht = fmaxf(pt, 0.f);
#pragma unroll
      for (int c = 0; c < C; c++) {
        so[c] = fmaf(w[m * C + m + c * m + k], hs, so[c]);
        to[c] = fmaf(w[half + m * C + m + c * m + k], ht, to[c]);
      }
    }
#pragma unroll
    for (int c = 0; c < C; c++) {
      float s_ = p.tanh_out ? tanhf(so[c]) : so[c];
      float t_ = p.tanh_out ? tanhf(to[c]) : to[c];
      if (!isfinite(s_)) s_ = NAN;
      if (!isfinite(t_)) t_ = NAN;
      if (!b[c]) z[c] = (z[c] - t_) * expf(-s_);
    }
  }
#pragma unroll
  for (int c = 0; c < C; c++) {
    float x = mm_fwd(z[c], p.fc.new_min, p.fc.new_max, p.fc.nmin[c], p.fc.nmax[c]);
    out[((int64_t)o * p.N + n) * C + c] = (1.0f / lin[c]) * (x - lin[C + c]);
  }
}

// ---------------------------------------------------------------- backward
template <int C>
__global__ void k_flow_bwd(FlowP p, int gs, int nbuf) {
  extern __shared__ float sm[];                     // [PF] weights + [nbuf][PF] gradient buffers
  const int PF = (int)p.P_flow + 2 * C;
  const int o = blockIdx.y, s = blockIdx.x;
  float* sp = sm;
  float* gbuf = sm + PF;
  const float* par = p.params + (int64_t)o * p.P + p.off_flow;
  for (int i = threadIdx.x; i < PF; i += blockDim.x) sp[i] = par[i];
  for (int i = threadIdx.x; i < nbuf * PF; i += blockDim.x) gbuf[i] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int ppw = 32 / gs;                 // pixels per warp iteration
  const int grp = lane / gs, k = lane % gs;
  const bool act = k < p.m;
  const int nw = blockDim.x >> 5;
  float* g = gbuf + (int64_t)(warp * ppw + grp) * PF;
  const int m = p.m, half = 2 * m * C + m + C;
  const int64_t r0 = (int64_t)s * p.chunk, r1 = r0 + p.chunk < p.N ? r0 + p.chunk : p.N;
  const int64_t span = (r1 > r0 ? r1 - r0 : 0);
  const int64_t iters = (span + (int64_t)nw * ppw - 1) / ((int64_t)nw * ppw);
  for (int64_t it = 0; it < iters; it++) {
    int64_t n = r0 + (it * nw + warp) * ppw + grp;
    bool live = n < r1;                      // whole warp keeps iterating (shuffles stay converged)
    int64_t nn = live ? n : r0;
    float dz[C];
    const float* dx = p.dX + ((int64_t)o * p.N + nn) * 4;
#pragma unroll
    for (int c = 0; c < C; c++)
      dz[c] = live ? dx[c] * ((p.fc.nmax[c] - p.fc.nmin[c]) / (p.fc.new_max - p.fc.new_min)) : 0.f;
    const float* zin = p.zin + ((int64_t)o * p.N + nn) * (p.F * C);
    for (int f = p.F - 1; f >= 0; f--) {
      const float* w = sp + (int64_t)f * p.per_flow;
      float* gw = g + (int64_t)f * p.per_flow;
      float z[C], zm[C];
      bool b[C];
#pragma unroll
      for (int c = 0; c < C; c++) { z[c] = zin[f * C + c]; b[c] = p.fc.masks[f * C + c] != 0; zm[c] = b[c] ? z[c] : 0.f; }
      float ps = 0.f, pt = 0.f, hs = 0.f, ht = 0.f;
      if (act) {
        ps = w[m * C + k]; pt = w[half + m * C + k];
#pragma unroll
        for (int c = 0; c < C; c++) { ps = fmaf(w[k * C + c], zm[c], ps); pt = fmaf(w[half + k * C + c], zm[c], pt); }
        hs = fmaxf(ps, 0.f); ht = fmaxf(pt, 0.f);
      }
      float so[C], to[C];
#pragma unroll
      for (int c = 0; c < C; c++) {
        so[c] = (act && !b[c]) ? w[m * C + m + c * m + k] * hs : 0.f;
        to[c] = (act && !b[c]) ? w[half + m * C + m + c * m + k] * ht : 0.f;
      }
      for (int off = gs >> 1; off > 0; off >>= 1) {
#pragma unroll
        for (int c = 0; c < C; c++) {
          so[c] += __shfl_xor_sync(0xffffffffu, so[c], off);
          to[c] += __shfl_xor_sync(0xffffffffu, to[c], off);
        }
      }
      const float* an = w + 2 * half;
      float dsr[C], dtr[C], dzin[C];
#pragma unroll
      for (int c = 0; c < C; c++) {
        float ea = expf(an[c]);
        float dzp = dz[c] * ea;
        if (b[c]) {
          // masked component passes through the coupling
          if (k == 0 && live) { gw[2 * half + c] += dz[c] * z[c] * ea; gw[2 * half + C + c] += dz[c]; }
          dsr[c] = 0.f; dtr[c] = 0.f; dzin[c] = dzp;
        } else {
          float sr = so[c] + w[m * C + m + C * m + c], tr = to[c] + w[half + m * C + m + C * m + c];
          float sv = p.tanh_out ? tanhf(sr) : sr, tv = p.tanh_out ? tanhf(tr) : tr;
          float e = expf(sv);
          float zp = fmaf(z[c], e, tv);
          if (k == 0 && live) { gw[2 * half + c] += dz[c] * zp * ea; gw[2 * half + C + c] += dz[c]; }
          float ds = dzp * z[c] * e, dt = dzp;
          dsr[c] = p.tanh_out ? ds * (1.f - sv * sv) : ds;
          dtr[c] = p.tanh_out ? dt * (1.f - tv * tv) : dt;
          dzin[c] = dzp * e;
          if (k == 0 && live) { gw[m * C + m + C * m + c] += dsr[c]; gw[half + m * C + m + C * m + c] += dtr[c]; }
        }
      }
      float dps = 0.f, dpt = 0.f;
      if (act) {
        float dhs = 0.f, dht = 0.f;
#pragma unroll
        for (int c = 0; c < C; c++) {
          if (!b[c]) {
            dhs = fmaf(dsr[c], w[m * C + m + c * m + k], dhs);
            dht = fmaf(dtr[c], w[half + m * C + m + c * m + k], dht);
            if (live) { gw[m * C + m + c * m + k] += dsr[c] * hs; gw[half + m * C + m + c * m + k] += dtr[c] * ht; }
          }
        }
        dps = ps > 0.f ? dhs : 0.f;
        dpt = pt > 0.f ? dht : 0.f;
        if (live) {
          gw[m * C + k] += dps; gw[half + m * C + k] += dpt;
#pragma unroll
          for (int c = 0; c < C; c++) if (b[c]) { gw[k * C + c] += dps * zm[c]; gw[half + k * C + c] += dpt * zm[c]; }
        }
      }
      float dzm[C];
#pragma unroll
      for (int c = 0; c < C; c++) dzm[c] = (act && b[c]) ? fmaf(dps, w[k * C + c], dpt * w[half + k * C + c]) : 0.f;
      for (int off = gs >> 1; off > 0; off >>= 1) {
#pragma unroll
        for (int c = 0; c < C; c++) dzm[c] += __shfl_xor_sync(0xffffffffu, dzm[c], off);
      }
#pragma unroll
      for (int c = 0; c < C; c++) dz[c] = dzin[c] + (b[c] ? dzm[c] : 0.f);
    }
    if (k == 0 && live && p.use_linear) {
#pragma unroll
      for (int c = 0; c < C; c++) {
        float dxc = dz[c] * ((p.fc.new_max - p.fc.new_min) / (p.fc.nmax[c] - p.fc.nmin[c]));
        g[p.P_flow + c] += dxc * coord(p.g, n, c);     // linear.weight
        g[p.P_flow + C + c] += dxc;                    // linear.bias
      }
    }
  }
  __syncthreads();
  float* out = p.fpart + ((int64_t)s * p.O + o) * PF;
  for (int i = threadIdx.x; i < PF; i += blockDim.x) {
    float a = 0.f;
    for (int bb = 0; bb < nbuf; bb++) a += gbuf[(int64_t)bb * PF + i];
    out[i] = a;
  }
}

// ---------------------------------------------------------------- ActNorm data-dependent init
// z state lives in X[n] (pre-ActNorm output of the previous coupling).  One pass per flow.
template <int C>
__global__ void __launch_bounds__(256) k_flow_init_pass(FlowP p, int f, double* stats) {
  extern __shared__ float sp[];
  const float* par = p.params + p.off_flow;
  const int PF = (int)p.P_flow + 2 * C;
  for (int i = threadIdx.x; i < PF; i += blockDim.x) sp[i] = par[i];
  __syncthreads();
  int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int m = p.m, half = 2 * m * C + m + C;
  float zc[C];
  bool live = n < p.N;
#pragma unroll
  for (int c = 0; c < C; c++) zc[c] = 0.f;
  if (live) {
    float z[C];
    const float* lin = sp + p.P_flow;
    if (f == 0) {
#pragma unroll
      for (int c = 0; c < C; c++)
        z[c] = mm_fwd(coord(p.g, n, c) * lin[c] + lin[C + c], p.fc.nmin[c], p.fc.nmax[c], p.fc.new_min, p.fc.new_max);
    } else {
      const float* an = sp + (int64_t)(f - 1) * p.per_flow + 2 * half;   // ActNorm f-1, already initialised
      float4 v = *reinterpret_cast<const float4*>(p.X + n * 4);
      float zz[3] = {v.x, v.y, v.z};
#pragma unroll
      for (int c = 0; c < C; c++) z[c] = fmaf(zz[c], expf(an[c]), an[C + c]);
    }
    const float* w = sp + (int64_t)f * p.per_flow;
    float zm[C], so[C], to[C];
    bool b[C];
#pragma unroll
    for (int c = 0; c < C; c++) {
      b[c] = p.fc.masks[f * C + c] != 0; zm[c] = b[c] ? z[c] : 0.f;
      so[c] = w[m * C + m + C * m + c]; to[c] = w[half + m * C + m + C * m + c];
    }
    for (int k = 0; k < m; k++) {
      float ps = w[m * C + k], pt = w[half + m * C + k];
#pragma unroll
      for (int c = 0; c < C; c++) { ps = fmaf(w[k * C + c], zm[c], ps); pt = fmaf(w[half + k * C + c], zm[c], pt); }
      float hs = fmaxf(ps, 0.f), ht = fmaxf(pt, 0.f);
#pragma unroll
      for (int c = 0; c < C; c++) {
        so[c] = fmaf(w[m * C + m + c * m + k], hs, so[c]);
        to[c] = fmaf(w[half + m * C + m + c * m + k], ht, to[c]);
      }
    }
#pragma unroll
    for (int c = 0; c < C; c++) {
      float s = p.tanh_out ? tanhf(so[c]) : so[c], t = p.tanh_out ? tanhf(to[c]) : to[c];
      zc[c] = b[c] ? z[c] : fmaf(z[c], expf(s), t);
    }
    *reinterpret_cast<float4*>(p.X + n * 4) = make_float4(zc[0], zc[1], C > 2 ? zc[2] : 0.f, 1.f);
  }
  // block sums of z and z^2 in double, then one atomic per block and statistic
  __shared__ double red[2 * 3][8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int c = 0; c < C; c++) {
    double a = live ? (double)zc[c] : 0.0, q = a * a;
    for (int off = 16; off > 0; off >>= 1) {
      a += __shfl_xor_sync(0xffffffffu, a, off);
      q += __shfl_xor_sync(0xffffffffu, q, off);
    }
    if (lane == 0) { red[c][warp] = a; red[3 + c][warp] = q; }
  }
  __syncthreads();
  if (threadIdx.x < 2 * C) {
    int c = threadIdx.x % C, which = threadIdx.x / C;
    double a = 0.0;
    for (int i = 0; i < 8; i++) a += red[which * 3 + c][i];
    atomicAdd(&stats[which * C + c], a);
  }
}

template <int C>
__global__ void k_flow_init_set(FlowP p, int f, const double* stats, float* params) {
  // s = -log(std + 1e-6) (unbiased std), t = -mean * exp(s)   (normflows ActNorm.forward, first call)
  int c = threadIdx.x;
  if (c >= C) return;
  const int m = p.m, half = 2 * m * C + m + C;
  double n = (double)p.N;
  double mean = stats[c] / n;
  double var = (stats[C + c] - stats[c] * stats[c] / n) / (n - 1.0);
  if (var < 0.0) var = 0.0;
  float sd = (float)sqrt(var);
  float s = -logf(sd + 1e-6f);
  float t = -(float)mean * expf(s);
  float* an = params + p.off_flow + (int64_t)f * p.per_flow + 2 * half;
  an[c] = s;
  an[C + c] = t;
}

// ---------------------------------------------------------------- learn_flow_identity loss
// loss = mean over N*C of (grid - flow_net(grid))^2  (SE("mean"), path_connected_net.py:172,223)
template <int C>
__global__ void __launch_bounds__(256) k_identity_loss(FlowP p, float* dX, float* lossp) {
  const int s = blockIdx.x, o = blockIdx.y;
  const int64_t r0 = (int64_t)s * p.chunk, r1 = r0 + p.chunk < p.N ? r0 + p.chunk : p.N;
  const float inv = 1.0f / ((float)p.N * (float)C);
  float acc = 0.f;
  for (int64_t n = r0 + threadIdx.x; n < r1; n += 256) {
    float4 v = *reinterpret_cast<const float4*>(p.X + ((int64_t)o * p.N + n) * 4);
    float xd[3] = {v.x, v.y, v.z};
    float d[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int c = 0; c < C; c++) {
      float e = coord(p.g, n, c) - xd[c];
      acc = fmaf(e, e, acc);
      d[c] = -2.f * e * inv;
    }
    *reinterpret_cast<float4*>(dX + ((int64_t)o * p.N + n) * 4) = make_float4(d[0], d[1], d[2], 0.f);
  }
  __shared__ float red[8];
  for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f;
    for (int i = 0; i < 8; i++) a += red[i];
    lossp[s * p.O + o] = a * inv;
  }
}

// ======================================================================= launchers
static FlowP make_p(const awb_prior* h, const float* params, const awb_grid_spec* g, const Workspace& ws) {
  const Layout& L = h->lay;
  FlowP p = {};
  p.g.mode = g->mode; p.g.B = g->B; p.g.H = g->H; p.g.W = g->W; p.g.C = L.C;
  p.g.t0 = g->t0; p.g.t_step = g->t_step; p.g.grid = g->grid;
  p.fc = h->fc;
  p.params = params; p.P = L.P; p.off_flow = L.off_flow; p.P_flow = L.P_flow; p.per_flow = L.per_flow;
  p.off_lin = L.off_lin;
  p.C = L.C; p.F = L.F; p.m = L.m; p.tanh_out = h->desc.flow_tanh; p.use_linear = 1;
  p.N = (int64_t)g->B * g->H * g->W;
  p.X = ws.X; p.zin = nullptr; p.deformed = nullptr; p.dX = ws.dX; p.fpart = ws.fpart;
  p.chunk = split_chunk(p.N); p.O = h->desc.n_objects;
  return p;
}

int flow_forward(const awb_prior* h, const float* params, const awb_grid_spec* g, const Workspace& ws,
                 float* deformed, cudaStream_t st, bool use_linear) {
  FlowP p = make_p(h, params, g, ws);
  p.use_linear = use_linear ? 1 : 0;
  p.zin = ws.flowz;
  p.deformed = deformed;
  const int PF = (int)(h->lay.P_flow + 2 * h->lay.C);
  size_t smem = sizeof(float) * PF;
  dim3 grid((unsigned)((p.N + 255) / 256), h->desc.n_objects);
  if (h->lay.C == 2) {
    AWB_CUDA(cudaFuncSetAttribute(k_flow_fwd<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    AWB_LAUNCH(PK_FLOW_FWD, st, k_flow_fwd<2><<<grid, 256, smem, st>>>(p));
  } else {
    AWB_CUDA(cudaFuncSetAttribute(k_flow_fwd<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    AWB_LAUNCH(PK_FLOW_FWD, st, k_flow_fwd<3><<<grid, 256, smem, st>>>(p));
  }
  AWB_CUDA(cudaGetLastError());
  return AWB_OK;
}

int flow_inverse(const awb_prior* h, const float* params, const awb_grid_spec* g, float* out, cudaStream_t st) {
  Workspace none = {};
  FlowP p = make_p(h, params, g, none);
  const int PF = (int)(h->lay.P_flow + 2 * h->lay.C);
  size_t smem = sizeof(float) * PF;
  dim3 grid((unsigned)((p.N + 255) / 256), h->desc.n_objects);
  if (h->lay.C == 2) {
    AWB_CUDA(cudaFuncSetAttribute(k_flow_inv<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    AWB_LAUNCH(PK_FLOW_FWD, st, k_flow_inv<2><<<grid, 256, smem, st>>>(p, out));
  } else {
    AWB_CUDA(cudaFuncSetAttribute(k_flow_inv<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    AWB_LAUNCH(PK_FLOW_FWD, st, k_flow_inv<3><<<grid, 256, smem, st>>>(p, out));
  }
  AWB_CUDA(cudaGetLastError());
  return AWB_OK;
}

int flow_backward(const awb_prior* h, const float* params, const awb_grid_spec* g, const Workspace& ws,
                  cudaStream_t st, bool use_linear) {
  FlowP p = make_p(h, params, g, ws);
  p.use_linear = use_linear ? 1 : 0;
  p.zin = ws.flowz;
  if (!p.zin) { set_error("flow backward needs a training workspace"); return AWB_ERR_WORKSPACE; }
  const int PF = (int)(h->lay.P_flow + 2 * h->lay.C);
  int gs = 8;
  while (gs < h->lay.m) gs <<= 1;
  const int ppw = 32 / gs;
  // as many warps as the gradient buffers allow within ~200 KB of shared memory
  int nw = 8;
  while (nw > 1 && sizeof(float) * PF * (1 + (size_t)nw * ppw) > 200 * 1024) nw--;
  size_t smem = sizeof(float) * PF * (1 + (size_t)nw * ppw);
  if (smem > 220 * 1024) { set_error("flow too large for the shared-memory gradient buffers (%zu bytes)", smem); return AWB_ERR_UNSUPPORTED; }
  dim3 grid(n_splits(p.N), h->desc.n_objects);
  if (h->lay.C == 2) {
    AWB_CUDA(cudaFuncSetAttribute(k_flow_bwd<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    AWB_LAUNCH(PK_FLOW_BWD, st, k_flow_bwd<2><<<grid, 32 * nw, smem, st>>>(p, gs, nw * ppw));
  } else {
    AWB_CUDA(cudaFuncSetAttribute(k_flow_bwd<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    AWB_LAUNCH(PK_FLOW_BWD, st, k_flow_bwd<3><<<grid, 32 * nw, smem, st>>>(p, gs, nw * ppw));
  }
  AWB_CUDA(cudaGetLastError());
  return AWB_OK;
}

int flow_identity_loss(const awb_prior* h, const awb_grid_spec* g, const Workspace& ws, cudaStream_t st) {
  FlowP p = make_p(h, nullptr, g, ws);
  dim3 grid(n_splits(p.N), h->desc.n_objects);
  if (h->lay.C == 2) AWB_LAUNCH(PK_OUT_LOSS, st, k_identity_loss<2><<<grid, 256, 0, st>>>(p, ws.dX, ws.lossp));
  else AWB_LAUNCH(PK_OUT_LOSS, st, k_identity_loss<3><<<grid, 256, 0, st>>>(p, ws.dX, ws.lossp));
  AWB_CUDA(cudaGetLastError());
  return AWB_OK;
}

int flow_actnorm_init(const awb_prior* h, float* params, const awb_grid_spec* g, const Workspace& ws,
                      cudaStream_t st) {
  if (h->desc.n_objects != 1) { set_error("ActNorm init is per object: call it on single-object handles"); return AWB_ERR_UNSUPPORTED; }
  FlowP p = make_p(h, params, g, ws);
  const int C = h->lay.C;
  const int PF = (int)(h->lay.P_flow + 2 * C);
  size_t smem = sizeof(float) * PF;
  double* stats = (double*)ws.lossp;            // 2*C doubles of scratch (lossp is 256-byte aligned, >= 8 floats)
  if (C == 2) AWB_CUDA(cudaFuncSetAttribute(k_flow_init_pass<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  else AWB_CUDA(cudaFuncSetAttribute(k_flow_init_pass<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  unsigned blocks = (unsigned)((p.N + 255) / 256);
  for (int f = 0; f < h->lay.F; f++) {
    AWB_CUDA(cudaMemsetAsync(stats, 0, sizeof(double) * 6, st));
    if (C == 2) {
      AWB_LAUNCH(PK_MISC, st, k_flow_init_pass<2><<<blocks, 256, smem, st>>>(p, f, stats));
      AWB_LAUNCH(PK_MISC, st, k_flow_init_set<2><<<1, 32, 0, st>>>(p, f, stats, params));
    } else {
      AWB_LAUNCH(PK_MISC, st, k_flow_init_pass<3><<<blocks, 256, smem, st>>>(p, f, stats));
      AWB_LAUNCH(PK_MISC, st, k_flow_init_set<3><<<1, 32, 0, st>>>(p, f, stats, params));
    }
  }
  AWB_CUDA(cudaGetLastError());
  return AWB_OK;
}

}  // namespace awb
