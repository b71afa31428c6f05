// Measured int8 tensor-core peak of the device: a shared-memory-resident tcgen05.mma kind::i8 loop.
//
// BASELINE.json's first metric is "int8 TOPS & %-of-peak on q8gemm"; MEASURED_PEAKS.json (driver-written) carries HBM and
// bf16 numbers only, so the denominator is measured here ("own tcgen05 kind::i8 smem-resident
// microbenchmark").  One CTA per SM issues back-to-back UMMAs (M = 128, N = 256, K = 32 per instruction, u8 x s8 -> s32)
// on operands that never leave shared memory, round-robin over two TMEM accumulator stages so that consecutive
// instructions do not serialise on one accumulator; nothing is loaded, stored or requantised.  ops = 2 * M * N * K per
// instruction, the reference's own counter (bench/q8gemm.cc:108).  This is the ceiling a q8gemm kernel on this chip can
// approach, at the clocks the power limit allows while the tensor pipe is saturated.
#include <cuda_runtime.h>
#include <stdint.h>

#include "sm100_ptx.cuh"

namespace q8 {
namespace {

constexpr int kPeakN = 256;
constexpr int kPeakKSteps = 8;  // K = 256 bytes of operand per accumulator pass (A: 32 KB, B: 64 KB of smem)

template <int CG>
__global__ void __launch_bounds__(128, 1) q8_peak_kernel(int iters) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ uint64_t done_bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_bytes = 128 * 32 * kPeakKSteps, b_bytes = kPeakN * 32 * kPeakKSteps;
  // operands: any bytes will do (the tensor pipe's speed does not depend on the data); zero them for determinism
  for (uint32_t o = threadIdx.x * 16; o < a_bytes + b_bytes; o += blockDim.x * 16)
    asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(smem_base + o), "r"(0x01010101u) : "memory");
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&done_bar), 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc<512>(smem_u32(&tmem_slot));
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = __shfl_sync(0xffffffffu, tmem_slot, 0);
  if (warp == 0) {
    const uint32_t idesc = umma_idesc_i8(128, kPeakN, false, true);
    const uint64_t a_tmpl = umma_desc_kmajor_noswizzle(0, 128 * 16, 128);
    const uint64_t b_tmpl = umma_desc_kmajor_noswizzle(0, kPeakN * 16, 128);
    const uint32_t a0 = smem_base, b0 = smem_base + a_bytes;
    if (elect_one()) {
      for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int ks = 0; ks < kPeakKSteps; ks++) {
          const uint64_t ad = a_tmpl + ((a0 + (uint32_t) ks * 2 * 128 * 16) >> 4);
          const uint64_t bd = b_tmpl + ((b0 + (uint32_t) ks * 2 * kPeakN * 16) >> 4);
          umma_i8(tmem + (uint32_t) ((ks & 1) * kPeakN), ad, bd, idesc, 1u);
        }
      }
      umma_commit(smem_u32(&done_bar));
    }
    __syncwarp();
    mbar_wait(smem_u32(&done_bar), 0);
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after_sync();
    tmem_dealloc<512>(tmem);
  }
}

}  // namespace

// -> tera-ops/s (2 * MACs), timed with CUDA events on `stream` over `reps` launches after one warm-up launch
cudaError_t measure_int8_peak(int num_sms, int iters, int reps, cudaStream_t stream, double* tops, double* ms_out) {
  auto kern = q8_peak_kernel<1>;
  const int smem = 128 * 32 * kPeakKSteps + kPeakN * 32 * kPeakKSteps + 2048;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return e;
  cudaEvent_t e0, e1;
  if ((e = cudaEventCreate(&e0)) != cudaSuccess) return e;
  if ((e = cudaEventCreate(&e1)) != cudaSuccess) return e;
  kern<<<num_sms, 128, smem, stream>>>(iters);
  cudaEventRecord(e0, stream);
  for (int r = 0; r < reps; r++) kern<<<num_sms, 128, smem, stream>>>(iters);
  cudaEventRecord(e1, stream);
  e = cudaEventSynchronize(e1);
  if (e == cudaSuccess) e = cudaGetLastError();
  float ms = 0.f;
  if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, e0, e1);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  if (e != cudaSuccess) return e;
  const double ops = 2.0 * 128 * kPeakN * 32 * kPeakKSteps * (double) iters * num_sms * reps;
  *tops = ops / (ms * 1e-3) / 1e12;
  *ms_out = ms / reps;
  return cudaSuccess;
}

}  // namespace q8
