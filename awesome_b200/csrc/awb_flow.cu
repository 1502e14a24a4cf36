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
// backward: (A) one thread per pixel propagates the coordinate gradient through the flows and records the per-pixel
//           factors of every weight gradient; (B) one lane per hidden unit and flow sums them over pixels in
//           registers.  No atomics, no per-pixel shuffles; partials are reduced in a fixed order.
#include <math.h>

#include <type_traits>

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

// ---- k-packed weight staging for the per-pixel kernels.  The MLP loops read, per hidden unit k, the 4C+2 scalars
// (w1s[C], b1s, w2s[C], w1t[C], b1t, w2t[C]); in state_dict order they sit in 6 different arrays (6 broadcast LDS per
// unit).  Staged as one padded record per (flow, unit) they are 3 (C=2) / 4 (C=3) LDS.128 -- the loops are
// shared-memory-issue bound, not FMA bound.  Per flow: [m records | b2s[C] | b2t[C] | an_s[C] | an_t[C]].
template <int C> struct FlowPack {
  static constexpr int RK = (4 * C + 2 + 3) / 4 * 4;       // 12 (C=2), 16 (C=3)
  __host__ __device__ static int flow_stride(int m) { return m * RK + 4 * C; }
};
// Flows whose mask pattern is one of the compiled-in ones (binary codes 1 .. 2^C - 2) get a COMPACT record in the first 8
// floats: only what the pattern uses -- [w1s of the masked comps | b1s | w2s of the transformed comps | w1t masked | b1t |
// w2t transformed] = 2C + 2 <= 8 floats = 2 LDS.128 per hidden unit instead of 3 (C = 2) / 4 (C = 3).
template <int C>
__device__ __forceinline__ bool flow_pattern_compiled(int mb) { return mb >= 1 && mb <= (1 << C) - 2; }
template <int C>
__device__ void stage_flow_packed(const float* __restrict__ par, int F, int m, int64_t per_flow, float* sp, const FlowConsts& fc) {
  constexpr int RK = FlowPack<C>::RK;
  const int half = 2 * m * C + m + C, FS = FlowPack<C>::flow_stride(m);
  for (int t = threadIdx.x; t < F * m; t += blockDim.x) {
    const int f = t / m, k = t - f * m;
    const float* w = par + (int64_t)f * per_flow;
    float* r = sp + f * FS + k * RK;
    int mb = 0;
#pragma unroll
    for (int c = 0; c < C; c++) mb |= fc.masks[f * C + c] != 0 ? 1 << c : 0;
    if (flow_pattern_compiled<C>(mb)) {
      int pos = 0;
      for (int c = 0; c < C; c++) if ((mb >> c) & 1) r[pos++] = w[k * C + c];                              // s.W1[k][c], masked c
      r[pos++] = w[m * C + k];                                                                         // s.b1[k]
      for (int c = 0; c < C; c++) if (!((mb >> c) & 1)) r[pos++] = w[m * C + m + c * m + k];               // s.W2[c][k], transformed c
      for (int c = 0; c < C; c++) if ((mb >> c) & 1) r[pos++] = w[half + k * C + c];                       // t.W1[k][c]
      r[pos++] = w[half + m * C + k];                                                                  // t.b1[k]
      for (int c = 0; c < C; c++) if (!((mb >> c) & 1)) r[pos++] = w[half + m * C + m + c * m + k];        // t.W2[c][k]
      continue;
    }
#pragma unroll
    for (int c = 0; c < C; c++) {
      r[c] = w[k * C + c];                               // s.W1[k][c]
      r[C + 1 + c] = w[m * C + m + c * m + k];           // s.W2[c][k]
      r[2 * C + 1 + c] = w[half + k * C + c];            // t.W1[k][c]
      r[3 * C + 2 + c] = w[half + m * C + m + c * m + k];   // t.W2[c][k]
    }
    r[C] = w[m * C + k];                                 // s.b1[k]
    r[3 * C + 1] = w[half + m * C + k];                  // t.b1[k]
  }
  for (int t = threadIdx.x; t < F * 4 * C; t += blockDim.x) {
    const int f = t / (4 * C), j = t - f * 4 * C, q = j / C, c = j - q * C;
    const float* w = par + (int64_t)f * per_flow;
    const float v = q == 0 ? w[m * C + m + C * m + c] : q == 1 ? w[half + m * C + m + C * m + c] : w[2 * half + (q - 2) * C + c];
    sp[f * FS + m * RK + j] = v;
  }
}


// ---- coupling MLPs with the mask pattern of the flow known at compile time.  MB: bit c set = component c is masked
// (passes through the coupling and feeds s / t); MB < 0 = pattern only known at run time.  With the pattern fixed, the
// multiplications by a masked-out zero and the outputs nobody reads disappear (half of the FMAs for the alternating
// masks of net_factory.py:86-99); the surviving arithmetic is unchanged, so results are bit-identical.
template <int C, int MB>
__device__ __forceinline__ bool masked_c(int c, const bool* b) { return MB < 0 ? b[c] : ((MB >> c) & 1) != 0; }

// s / t pre-outputs of one coupling: so[c], to[c] for the transformed components (the others are left untouched)
// record of hidden unit k: compact (8 floats) for a compiled-in pattern, the full RK floats otherwise
template <int C, int MB>
struct FlowRec {
  static constexpr int NM = MB < 0 ? 0 : ((MB & 1) + ((MB >> 1) & 1) + ((MB >> 2) & 1));     // masked components
  static constexpr int NU = C - NM;
  static constexpr int N4 = MB < 0 ? FlowPack<C>::RK / 4 : 2;
  float v[4 * N4];
  __device__ __forceinline__ void load(const float* __restrict__ rec) {
#pragma unroll
    for (int q4 = 0; q4 < N4; q4++) *reinterpret_cast<float4*>(&v[4 * q4]) = reinterpret_cast<const float4*>(rec)[q4];
  }
  // rank of component c among the masked / the transformed ones (constant-folded after unrolling)
  static __device__ __forceinline__ int rm(int c) { return __popc(MB & ((1 << c) - 1)); }
  static __device__ __forceinline__ int ru(int c) { return c - rm(c); }
  __device__ __forceinline__ float w1s(int c) const { return MB < 0 ? v[c] : v[rm(c)]; }
  __device__ __forceinline__ float b1s() const { return MB < 0 ? v[C] : v[NM]; }
  __device__ __forceinline__ float w2s(int c) const { return MB < 0 ? v[C + 1 + c] : v[NM + 1 + ru(c)]; }
  __device__ __forceinline__ float w1t(int c) const { return MB < 0 ? v[2 * C + 1 + c] : v[NM + 1 + NU + rm(c)]; }
  __device__ __forceinline__ float b1t() const { return MB < 0 ? v[3 * C + 1] : v[2 * NM + 1 + NU]; }
  __device__ __forceinline__ float w2t(int c) const { return MB < 0 ? v[3 * C + 2 + c] : v[2 * NM + 2 + NU + ru(c)]; }
};

// P pixels per thread share every weight record: the per-pixel kernels are bound by the shared-memory loads of the
// records (2 LDS.128 per ~6 FMAs with one pixel), not by the FMAs.
template <int C, int MB, int P>
__device__ __forceinline__ void coupling_mlp_fwd(const float* __restrict__ wf, int m, const float (*z)[C], const bool* b,
                                                 float (*so)[C], float (*to)[C]) {
  constexpr int RK = FlowPack<C>::RK;
#pragma unroll 4
  for (int k = 0; k < m; k++) {
    FlowRec<C, MB> r;
    r.load(wf + k * RK);
#pragma unroll
    for (int q = 0; q < P; q++) {
      float ps = r.b1s(), pt = r.b1t();
#pragma unroll
      for (int c = 0; c < C; c++)
        if (masked_c<C, MB>(c, b)) { ps = fmaf(r.w1s(c), z[q][c], ps); pt = fmaf(r.w1t(c), z[q][c], pt); }
      const float hs = fmaxf(ps, 0.f), ht = fmaxf(pt, 0.f);
#pragma unroll
      for (int c = 0; c < C; c++)
        if (MB < 0 || !masked_c<C, MB>(c, b)) {
          so[q][c] = fmaf(r.w2s(c), hs, so[q][c]);
          to[q][c] = fmaf(r.w2t(c), ht, to[q][c]);
        }
    }
  }
}

// gradient reaching the masked inputs through the two MLPs: dzm[c] for masked c, from dsr / dtr of the transformed ones
template <int C, int MB>
__device__ __forceinline__ void coupling_mlp_bwd(const float* __restrict__ wf, int m, const float* z, const bool* b,
                                                 const float* dsr, const float* dtr, float* dzm) {
  constexpr int RK = FlowPack<C>::RK;
#pragma unroll 4
  for (int k = 0; k < m; k++) {
    FlowRec<C, MB> r;
    r.load(wf + k * RK);
    float ps = r.b1s(), pt = r.b1t();
#pragma unroll
    for (int c = 0; c < C; c++)
      if (masked_c<C, MB>(c, b)) { ps = fmaf(r.w1s(c), z[c], ps); pt = fmaf(r.w1t(c), z[c], pt); }
    float dps = 0.f, dpt = 0.f;
#pragma unroll
    for (int c = 0; c < C; c++)
      if (MB < 0 || !masked_c<C, MB>(c, b)) {
        dps = fmaf(dsr[c], r.w2s(c), dps);
        dpt = fmaf(dtr[c], r.w2t(c), dpt);
      }
    dps = ps > 0.f ? dps : 0.f;
    dpt = pt > 0.f ? dpt : 0.f;
#pragma unroll
    for (int c = 0; c < C; c++)
      if (MB < 0 || masked_c<C, MB>(c, b)) dzm[c] = fmaf(dps, r.w1s(c), fmaf(dpt, r.w1t(c), dzm[c]));
  }
}

// warp-uniform dispatch on the flow's mask pattern: every pattern net_factory.py:86-99 produces (the binary codes
// 1 .. 2^C - 2) is compiled in for C = 2 and C = 3; anything else runs the generic path
#define AWB_FLOW_DISPATCH(C_, mb_, CALL)                                  \
  do {                                                                    \
    if ((mb_) == 1) { CALL(1); }                                          \
    else if ((mb_) == 2) { CALL(2); }                                     \
    else if ((C_) == 3 && (mb_) == 3) { CALL(3); }                        \
    else if ((C_) == 3 && (mb_) == 4) { CALL(4); }                        \
    else if ((C_) == 3 && (mb_) == 5) { CALL(5); }                        \
    else if ((C_) == 3 && (mb_) == 6) { CALL(6); }                        \
    else { CALL(-1); }                                                    \
  } while (0)

constexpr int FLOW_FWD_P = 2;      // pixels per thread of k_flow_fwd
template <int C>
__global__ void __launch_bounds__(256) k_flow_fwd(FlowP p) {
  extern __shared__ __align__(16) float sp[];   // k-packed flow weights + [2C] linear
  constexpr int RK = FlowPack<C>::RK;
  constexpr int P = FLOW_FWD_P;
  const int o = blockIdx.y;
  const float* par = p.params + (int64_t)o * p.P + p.off_flow;
  const int m = p.m, FS = FlowPack<C>::flow_stride(m);
  stage_flow_packed<C>(par, p.F, m, p.per_flow, sp, p.fc);
  float* lin = sp + p.F * FS;
  if (threadIdx.x < 2 * C) lin[threadIdx.x] = par[p.P_flow + threadIdx.x];
  __syncthreads();
  // pixel q of this thread: consecutive threads take consecutive pixels within each of the block's P row groups
  int64_t n[P];
  bool ok[P];
  float z[P][C];
#pragma unroll
  for (int q = 0; q < P; q++) {
    const int64_t nq = ((int64_t)blockIdx.x * P + q) * blockDim.x + threadIdx.x;
    ok[q] = nq < p.N;
    n[q] = ok[q] ? nq : p.N - 1;          // padding threads recompute the last pixel and store nothing
#pragma unroll
    for (int c = 0; c < C; c++) {
      float x = coord(p.g, n[q], c);
      if (p.use_linear) x = x * lin[c] + lin[C + c];
      z[q][c] = mm_fwd(x, p.fc.nmin[c], p.fc.nmax[c], p.fc.new_min, p.fc.new_max);
    }
  }
  if (!ok[0]) return;
  for (int f = 0; f < p.F; f++) {
    const float* wf = sp + f * FS;
    const float* tail = wf + m * RK;          // b2s[C] | b2t[C] | an_s[C] | an_t[C]
    bool b[C];
    int mb = 0;
#pragma unroll
    for (int c = 0; c < C; c++) { b[c] = p.fc.masks[f * C + c] != 0; mb |= b[c] ? 1 << c : 0; }
    float so[P][C], to[P][C];
#pragma unroll
    for (int q = 0; q < P; q++) {
      if (p.zin && ok[q]) {
        float* zin = p.zin + ((int64_t)o * p.N + n[q]) * (p.F * C);
        if (C == 2) *reinterpret_cast<float2*>(zin + f * C) = make_float2(z[q][0], z[q][1]);
        else {
#pragma unroll
          for (int c = 0; c < C; c++) zin[f * C + c] = z[q][c];
        }
      }
#pragma unroll
      for (int c = 0; c < C; c++) { so[q][c] = tail[c]; to[q][c] = tail[C + c]; }
    }
#define AWB_CALL(MB) coupling_mlp_fwd<C, MB, P>(wf, m, z, b, so, to)
    AWB_FLOW_DISPATCH(C, mb, AWB_CALL);
#undef AWB_CALL
#pragma unroll
    for (int q = 0; q < P; q++) {
#pragma unroll
      for (int c = 0; c < C; c++) {
        float zc = z[q][c];
        if (!b[c]) {
          float s_ = p.tanh_out ? tanhf(so[q][c]) : so[q][c];
          float t_ = p.tanh_out ? tanhf(to[q][c]) : to[q][c];
          if (!isfinite(s_)) s_ = NAN;
          if (!isfinite(t_)) t_ = NAN;
          zc = fmaf(z[q][c], expf(s_), t_);
        }
        z[q][c] = fmaf(zc, expf(tail[2 * C + c]), tail[3 * C + c]);      // ActNorm
      }
    }
  }
#pragma unroll
  for (int q = 0; q < P; q++) {
    if (!ok[q]) continue;
    float xd[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int c = 0; c < C; c++) xd[c] = mm_fwd(z[q][c], p.fc.new_min, p.fc.new_max, p.fc.nmin[c], p.fc.nmax[c]);
    *reinterpret_cast<float4*>(p.X + ((int64_t)o * p.N + n[q]) * 4) = make_float4(xd[0], xd[1], xd[2], 1.f);
    if (p.deformed) {
#pragma unroll
      for (int c = 0; c < C; c++) p.deformed[((int64_t)o * p.N + n[q]) * C + c] = xd[c];
    }
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
// Two kernels, neither with a per-pixel cross-lane exchange:
//  (A) k_flow_bwd_px: one thread per pixel walks the flows in reverse (recomputing each coupling from its stored
//      input), propagates d loss / d z, and records per flow the 4C per-pixel factors every parameter gradient of
//      that flow is linear in:  dsr, dtr (gradients at the pre-tanh outputs of the s / t nets) and the ActNorm terms.
//      It also owns the 1x1-conv gradients (block reduction at the end).
//  (B) k_flow_wgrad: grid = (pixel range, flow, object); lane k of every warp owns hidden unit k of that flow's two
//      MLPs and accumulates its weight gradients in registers over the pixels of the range (broadcast loads of the
//      per-pixel record, no shuffles, no atomics); the 8 warps are combined in a fixed order through shared memory.
template <int C>
__global__ void __launch_bounds__(1024) k_flow_bwd_px(FlowP p, float* __restrict__ rec) {
  extern __shared__ __align__(16) float sp[];   // k-packed flow weights
  constexpr int RK = FlowPack<C>::RK;
  const int o = blockIdx.y, s = blockIdx.x;
  const float* par = p.params + (int64_t)o * p.P + p.off_flow;
  const int PF = (int)p.P_flow + 2 * C;
  const int m = p.m, FS = FlowPack<C>::flow_stride(m);
  stage_flow_packed<C>(par, p.F, m, p.per_flow, sp, p.fc);
  __syncthreads();
  const int64_t r0 = (int64_t)s * p.chunk, r1 = r0 + p.chunk < p.N ? r0 + p.chunk : p.N;
  float glw[C], glb[C];
#pragma unroll
  for (int c = 0; c < C; c++) { glw[c] = 0.f; glb[c] = 0.f; }
  for (int64_t n = r0 + threadIdx.x; n < r1; n += blockDim.x) {
    float dz[C];
    const float* dx = p.dX + ((int64_t)o * p.N + n) * 4;
#pragma unroll
    for (int c = 0; c < C; c++) dz[c] = dx[c] * ((p.fc.nmax[c] - p.fc.nmin[c]) / (p.fc.new_max - p.fc.new_min));
    const float* zin = p.zin + ((int64_t)o * p.N + n) * (p.F * C);
    float* rn = rec + ((int64_t)o * p.N + n) * (p.F * 4 * C);
    for (int f = p.F - 1; f >= 0; f--) {
      const float* wf = sp + f * FS;
      const float* tail = wf + m * RK;
      float z[C], zm[C];
      bool b[C];
#pragma unroll
      for (int c = 0; c < C; c++) b[c] = p.fc.masks[f * C + c] != 0;
      if (C == 2) { const float2 zz = *reinterpret_cast<const float2*>(zin + f * C); z[0] = zz.x; z[1] = zz.y; }
      else {
#pragma unroll
        for (int c = 0; c < C; c++) z[c] = zin[f * C + c];
      }
#pragma unroll
      for (int c = 0; c < C; c++) zm[c] = b[c] ? z[c] : 0.f;
      float so[C], to[C];
      int mb = 0;
#pragma unroll
      for (int c = 0; c < C; c++) { so[c] = tail[c]; to[c] = tail[C + c]; mb |= b[c] ? 1 << c : 0; }
#define AWB_CALL(MB) coupling_mlp_fwd<C, MB, 1>(wf, m, reinterpret_cast<const float(*)[C]>(zm), b, reinterpret_cast<float(*)[C]>(so), reinterpret_cast<float(*)[C]>(to))
      AWB_FLOW_DISPATCH(C, mb, AWB_CALL);
#undef AWB_CALL
      float dsr[C], dtr[C], dzin[C], rv[4 * C];
#pragma unroll
      for (int c = 0; c < C; c++) {
        const float ea = expf(tail[2 * C + c]);
        const float dzp = dz[c] * ea;
        float das;
        if (b[c]) {
          das = dz[c] * z[c] * ea;            // masked component passes through the coupling
          dsr[c] = 0.f; dtr[c] = 0.f; dzin[c] = dzp;
        } else {
          const float sv = p.tanh_out ? tanhf(so[c]) : so[c], tv = p.tanh_out ? tanhf(to[c]) : to[c];
          const float e = expf(sv);
          das = dz[c] * fmaf(z[c], e, tv) * ea;
          const float ds = dzp * z[c] * e;
          dsr[c] = p.tanh_out ? ds * (1.f - sv * sv) : ds;
          dtr[c] = p.tanh_out ? dzp * (1.f - tv * tv) : dzp;
          dzin[c] = dzp * e;
        }
        rv[c] = dsr[c];
        rv[C + c] = dtr[c];
        rv[2 * C + c] = das;                  // d ActNorm.s
        rv[3 * C + c] = dz[c];                // d ActNorm.t
      }
      // the record of this flow is 4C contiguous floats, 16-byte aligned: full-sector vector stores
#pragma unroll
      for (int q4 = 0; q4 < C; q4++)
        reinterpret_cast<float4*>(rn + f * 4 * C)[q4] = make_float4(rv[4 * q4], rv[4 * q4 + 1], rv[4 * q4 + 2], rv[4 * q4 + 3]);
      // gradient reaching the masked inputs through the two MLPs
      float dzm[C];
#pragma unroll
      for (int c = 0; c < C; c++) dzm[c] = 0.f;
#define AWB_CALL(MB) coupling_mlp_bwd<C, MB>(wf, m, zm, b, dsr, dtr, dzm)
      AWB_FLOW_DISPATCH(C, mb, AWB_CALL);
#undef AWB_CALL
#pragma unroll
      for (int c = 0; c < C; c++) dz[c] = dzin[c] + (b[c] ? dzm[c] : 0.f);
    }
    if (p.use_linear) {
#pragma unroll
      for (int c = 0; c < C; c++) {
        const float dxc = dz[c] * ((p.fc.new_max - p.fc.new_min) / (p.fc.nmax[c] - p.fc.nmin[c]));
        glw[c] = fmaf(dxc, coord(p.g, n, c), glw[c]);     // linear.weight
        glb[c] += dxc;                                     // linear.bias
      }
    }
  }
  // 1x1-conv gradients: fixed-order block reduction
  __shared__ float red[32][2 * C];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int c = 0; c < C; c++) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      glw[c] += __shfl_xor_sync(0xffffffffu, glw[c], off);
      glb[c] += __shfl_xor_sync(0xffffffffu, glb[c], off);
    }
    if (lane == 0) { red[warp][c] = glw[c]; red[warp][C + c] = glb[c]; }
  }
  __syncthreads();
  if (threadIdx.x < 2 * C) {
    float a = 0.f;
    for (int i = 0; i < (int)(blockDim.x >> 5); i++) a += red[i][threadIdx.x];
    p.fpart[((int64_t)s * p.O + o) * PF + p.P_flow + threadIdx.x] = p.use_linear ? a : 0.f;
  }
}

template <int C>
__global__ void __launch_bounds__(256) k_flow_wgrad(FlowP p, const float* __restrict__ rec) {
  constexpr int NA = 2 * (2 * C + 1);      // per-lane accumulators: (W1[k][C], b1[k], W2[C][k]) x {s, t}
  __shared__ float red[8][NA + 1][32];
  __shared__ float red4[8][4 * C];
  const int s = blockIdx.x, f = blockIdx.y, o = blockIdx.z;
  const int m = p.m, half = 2 * m * C + m + C;
  const float* w = p.params + (int64_t)o * p.P + p.off_flow + (int64_t)f * p.per_flow;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool act = lane < m;
  const int k = act ? lane : 0;
  float w1s[C], w1t[C], w2s[C], w2t[C];
  bool b[C];
#pragma unroll
  for (int c = 0; c < C; c++) {
    w1s[c] = w[k * C + c]; w1t[c] = w[half + k * C + c];
    w2s[c] = w[m * C + m + c * m + k]; w2t[c] = w[half + m * C + m + c * m + k];
    b[c] = p.fc.masks[f * C + c] != 0;
  }
  const float b1s = w[m * C + k], b1t = w[half + m * C + k];
  float acc[NA];
#pragma unroll
  for (int i = 0; i < NA; i++) acc[i] = 0.f;
  float a4 = 0.f;                          // lanes < 4C: output biases of s / t and the ActNorm pair
  const int64_t r0 = (int64_t)s * p.chunk, r1 = r0 + p.chunk < p.N ? r0 + p.chunk : p.N;
  const int FC = p.F * C;
  // running pointers (8 pixels per step of the warp); the per-pixel record of this flow is 4C contiguous floats,
  // 16-byte aligned: float4 loads; the coupling input pair is 8-byte aligned for C = 2
  const float* zp = p.zin + ((int64_t)o * p.N + r0 + warp) * FC + f * C;
  const float* rp = rec + ((int64_t)o * p.N + r0 + warp) * (4 * FC) + f * 4 * C;
  const int64_t zstep = (int64_t)8 * FC, rstep = (int64_t)8 * 4 * FC;
  const int n_it = (int)((r1 - r0 - warp + 7) / 8);
  // the mask pattern is a property of the flow (blockIdx.y): the whole block runs one instantiation of the pixel loop
  auto pixel_loop = [&](auto mb_c) {
    constexpr int MB = decltype(mb_c)::value;
#pragma unroll 4
    for (int it = 0; it < (n_it > 0 ? n_it : 0); it++, zp += zstep, rp += rstep) {
      float zm[C], dsr[C], dtr[C];
      if (C == 2) {
        const float2 zz = *reinterpret_cast<const float2*>(zp);
        const float4 r4 = *reinterpret_cast<const float4*>(rp);
        zm[0] = zz.x; zm[1] = zz.y;
        dsr[0] = r4.x; dsr[1] = r4.y; dtr[0] = r4.z; dtr[1] = r4.w;
      } else {
#pragma unroll
        for (int c = 0; c < C; c++) { zm[c] = zp[c]; dsr[c] = rp[c]; dtr[c] = rp[C + c]; }
      }
      if (lane < 4 * C) a4 += rp[lane];
      float ps = b1s, pt = b1t, dps = 0.f, dpt = 0.f;
#pragma unroll
      for (int c = 0; c < C; c++) {
        if (masked_c<C, MB>(c, b)) { ps = fmaf(w1s[c], zm[c], ps); pt = fmaf(w1t[c], zm[c], pt); }
        if (MB < 0 || !masked_c<C, MB>(c, b)) { dps = fmaf(dsr[c], w2s[c], dps); dpt = fmaf(dtr[c], w2t[c], dpt); }
      }
      const float hs = fmaxf(ps, 0.f), ht = fmaxf(pt, 0.f);
      dps = ps > 0.f ? dps : 0.f;
      dpt = pt > 0.f ? dpt : 0.f;
#pragma unroll
      for (int c = 0; c < C; c++) {
        if (masked_c<C, MB>(c, b)) {
          acc[c] = fmaf(dps, zm[c], acc[c]);                             // d s.W1[k][c]
          acc[2 * C + 1 + c] = fmaf(dpt, zm[c], acc[2 * C + 1 + c]);     // d t.W1[k][c]
        }
        if (MB < 0 || !masked_c<C, MB>(c, b)) {
          acc[C + 1 + c] = fmaf(dsr[c], hs, acc[C + 1 + c]);             // d s.W2[c][k]
          acc[3 * C + 2 + c] = fmaf(dtr[c], ht, acc[3 * C + 2 + c]);     // d t.W2[c][k]
        }
      }
      acc[C] += dps;                                                     // d s.b1[k]
      acc[3 * C + 1] += dpt;                                             // d t.b1[k]
    }
  };
  {
    int mb = 0;
#pragma unroll
    for (int c = 0; c < C; c++) mb |= b[c] ? 1 << c : 0;
#define AWB_CALL(MB) pixel_loop(std::integral_constant<int, MB>{})
    AWB_FLOW_DISPATCH(C, mb, AWB_CALL);
#undef AWB_CALL
  }
#pragma unroll
  for (int i = 0; i < NA; i++) red[warp][i][lane] = acc[i];
  if (lane < 4 * C) red4[warp][lane] = a4;
  __syncthreads();
  const int PF = (int)p.P_flow + 2 * C;
  float* out = p.fpart + ((int64_t)s * p.O + o) * PF + (int64_t)f * p.per_flow;
  for (int t = threadIdx.x; t < NA * 32; t += blockDim.x) {
    const int i = t >> 5, kk = t & 31;
    if (kk >= m) continue;
    float a = 0.f;
    for (int ww = 0; ww < 8; ww++) a += red[ww][i][kk];
    const int net = i / (2 * C + 1), j = i % (2 * C + 1);
    int idx;
    if (j < C) idx = kk * C + j;                                   // W1[k][c]
    else if (j == C) idx = m * C + kk;                             // b1[k]
    else idx = m * C + m + (j - C - 1) * m + kk;                   // W2[c][k]
    out[net * half + idx] = a;
  }
  if (threadIdx.x < 4 * C) {
    float a = 0.f;
    for (int ww = 0; ww < 8; ww++) a += red4[ww][threadIdx.x];
    const int q = threadIdx.x / C, c = threadIdx.x % C;
    if (q == 0) out[m * C + m + C * m + c] = a;                    // s.b2[c]
    else if (q == 1) out[half + m * C + m + C * m + c] = a;        // t.b2[c]
    else if (q == 2) out[2 * half + c] = a;                        // ActNorm.s[c]
    else out[2 * half + C + c] = a;                                // ActNorm.t[c]
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
  size_t smem = sizeof(float) * ((size_t)h->lay.F * (h->lay.C == 2 ? FlowPack<2>::flow_stride(h->lay.m) : FlowPack<3>::flow_stride(h->lay.m)) + 2 * h->lay.C);
  dim3 grid((unsigned)((p.N + 256 * FLOW_FWD_P - 1) / (256 * FLOW_FWD_P)), h->desc.n_objects);
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
  if (!p.zin || !ws.flowg) { set_error("flow backward needs a training workspace"); return AWB_ERR_WORKSPACE; }
  if (h->lay.m > 32) { set_error("flow MLP width must be <= 32"); return AWB_ERR_UNSUPPORTED; }
  const size_t smem = sizeof(float) * ((size_t)h->lay.F * (h->lay.C == 2 ? FlowPack<2>::flow_stride(h->lay.m) : FlowPack<3>::flow_stride(h->lay.m)));
  const int S = n_splits(p.N);
  dim3 gridA(S, h->desc.n_objects), gridB(S, h->lay.F, h->desc.n_objects);
  if (h->lay.C == 2) {
    AWB_CUDA(cudaFuncSetAttribute(k_flow_bwd_px<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    AWB_LAUNCH(PK_FLOW_BWD, st, k_flow_bwd_px<2><<<gridA, 1024, smem, st>>>(p, ws.flowg));
    AWB_LAUNCH(PK_FLOW_BWD, st, k_flow_wgrad<2><<<gridB, 256, 0, st>>>(p, ws.flowg));
  } else {
    AWB_CUDA(cudaFuncSetAttribute(k_flow_bwd_px<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    AWB_LAUNCH(PK_FLOW_BWD, st, k_flow_bwd_px<3><<<gridA, 1024, smem, st>>>(p, ws.flowg));
    AWB_LAUNCH(PK_FLOW_BWD, st, k_flow_wgrad<3><<<gridB, 256, 0, st>>>(p, ws.flowg));
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
