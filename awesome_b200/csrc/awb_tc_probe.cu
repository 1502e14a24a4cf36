// UMMA operand-layout probe: one CTA computes D[128 x N] (+)= A * B with caller-supplied descriptor
// fields, so that tests can pin the shared-memory layout / descriptor conventions of awb_tc.cuh
// (K-major and MN-major operands, M-window offsets, narrow N) against a plain matmul before the
// fused kernels rely on them.
#include "awb_internal.cuh"
#include "awb_tc.cuh"

namespace awb {

struct ProbeP {
  const uint8_t* a_bytes; int a_size;
  const uint8_t* b_bytes; int b_size;
  float* D;                  // [128][N] row-major
  int N, K;
  int a_mn, b_mn;            // 1: MN-major operand
  uint32_t a_off, a_lbo, a_sbo, a_kstep;
  uint32_t b_off, b_lbo, b_sbo, b_kstep;
};

__global__ void __launch_bounds__(128) k_umma_probe(ProbeP p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  uint8_t* sa = smem;
  uint8_t* sb = smem + ((p.a_size + 1023) / 1024) * 1024;
  for (int i = threadIdx.x * 16; i < p.a_size; i += 128 * 16)
    *reinterpret_cast<uint4*>(sa + i) = *reinterpret_cast<const uint4*>(p.a_bytes + i);
  for (int i = threadIdx.x * 16; i < p.b_size; i += 128 * 16)
    *reinterpret_cast<uint4*>(sb + i) = *reinterpret_cast<const uint4*>(p.b_bytes + i);
  tc::fence_async_smem();
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tc::tmem_alloc<256>(&tmem_base);
  if (threadIdx.x == 0) { tc::mbar_init(&bar, 1); tc::mbar_fence_init(); }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tbase = tmem_base;
  if (threadIdx.x == 0) {
    const uint32_t idesc = tc::make_idesc(128, p.N, p.a_mn, p.b_mn);
    const uint32_t a0 = tc::smem_u32(sa) + p.a_off, b0 = tc::smem_u32(sb) + p.b_off;
    for (int k = 0; k < p.K / 16; k++) {
      uint64_t ad = tc::make_desc(a0 + k * p.a_kstep, p.a_lbo, p.a_sbo);
      uint64_t bd = tc::make_desc(b0 + k * p.b_kstep, p.b_lbo, p.b_sbo);
      tc::umma_f16(tbase, ad, bd, idesc, k > 0 ? 1u : 0u);
    }
    tc::umma_commit(&bar);
  }
  tc::mbar_wait(&bar, 0);
  tc::fence_after_sync();
  const int row = threadIdx.x;              // TMEM lane == accumulator row
  for (int c0 = 0; c0 < p.N; c0 += 8) {
    float v[8];
    tc::tmem_ld8(tbase + ((uint32_t)(warp * 32) << 16) + c0, v);
    tc::tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 8; j++) p.D[row * p.N + c0 + j] = v[j];
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc<256>(tbase);
}

}  // namespace awb

using namespace awb;

extern "C" int awb_debug_umma_probe(const void* a_bytes, int32_t a_size, const void* b_bytes, int32_t b_size, float* D,
                                    int32_t N, int32_t K, int32_t a_mn, int32_t b_mn, const uint32_t* a_desc4,
                                    const uint32_t* b_desc4, void* stream) {
  if (!a_bytes || !b_bytes || !D || !a_desc4 || !b_desc4) { set_error("null argument"); return AWB_ERR_INVALID; }
  if (N < 16 || N > 256 || N % 16 || K < 16 || K % 16 || a_size % 16 || b_size % 16) { set_error("bad probe shape"); return AWB_ERR_INVALID; }
  ProbeP p;
  p.a_bytes = (const uint8_t*)a_bytes; p.a_size = a_size; p.b_bytes = (const uint8_t*)b_bytes; p.b_size = b_size;
  p.D = D; p.N = N; p.K = K; p.a_mn = a_mn; p.b_mn = b_mn;
  p.a_off = a_desc4[0]; p.a_lbo = a_desc4[1]; p.a_sbo = a_desc4[2]; p.a_kstep = a_desc4[3];
  p.b_off = b_desc4[0]; p.b_lbo = b_desc4[1]; p.b_sbo = b_desc4[2]; p.b_kstep = b_desc4[3];
  size_t smem = ((a_size + 1023) / 1024) * 1024 + ((b_size + 1023) / 1024) * 1024 + 1024;
  AWB_CUDA(cudaFuncSetAttribute(k_umma_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  AWB_LAUNCH(PK_MISC, (cudaStream_t)stream, k_umma_probe<<<1, 128, smem, (cudaStream_t)stream>>>(p));
  AWB_CUDA(cudaGetLastError());
  return AWB_OK;
}

extern "C" int awb_debug_tc_trace_read(unsigned long long* host, int32_t max_ctas) {
  if (!host || max_ctas < 1) { set_error("bad argument"); return AWB_ERR_INVALID; }
  return tc_trace_read(host, max_ctas);
}
