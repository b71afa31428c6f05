// Probe: does a tcgen05 K-major SWIZZLE_32B A descriptor accept start addresses that are multiples of 32 bytes but not of the
// 256-byte swizzle atom (tap offsets of a depthwise tile), and does it need the descriptor's base-offset field?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I qnnpack_b200/csrc -o sw32_probe sw32_probe.cu
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <cuda_runtime.h>
#include "sm100_ptx.cuh"
using namespace q8;

constexpr int kPix = 192;  // pixels (32-byte rows) in the A buffer

__device__ __forceinline__ uint64_t desc_sw32(uint32_t addr, uint32_t sbo, uint32_t base_off) {
  uint64_t d = 0;
  d |= (uint64_t) ((addr >> 4) & 0x3FFF);
  d |= (uint64_t) 1 << 16;
  d |= (uint64_t) ((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t) 1 << 46;
  d |= (uint64_t) (base_off & 7) << 49;
  d |= (uint64_t) 6 << 61;  // SWIZZLE_32B
  return d;
}

__global__ void __launch_bounds__(128) probe(int32_t* out, int shift, int use_base_off, int row_shift_bytes) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_base = base, b_base = base + 16384;
  const int tid = threadIdx.x, warp = tid >> 5;
  // A: pixel q, byte k -> value; stored with the 32-byte swizzle of its absolute address (16-byte chunk ^= address bit 7)
  for (int i = tid; i < kPix * 32; i += 128) {
    const int q = i >> 5, k = i & 31;
    const uint32_t row = a_base + (uint32_t) q * 32;
    const uint32_t chunk = (uint32_t) (k >> 4) ^ ((row >> 7) & 1u);
    const uint32_t addr = row + chunk * 16 + (k & 15);
    const uint8_t v = (uint8_t) ((q * 7 + k * 3) & 0x7f);
    asm volatile("st.shared.u8 [%0], %1;" ::"r"(addr), "r"((uint32_t) v));
  }
  // B: identity, no-swizzle K-major [2 K chunks][32 rows][16 B]
  for (int i = tid; i < 2 * 32 * 16; i += 128) {
    const int c = i / 512, n = (i / 16) % 32, kk = i % 16;
    const uint8_t v = (c * 16 + kk == n) ? 1 : 0;
    asm volatile("st.shared.u8 [%0], %1;" ::"r"(b_base + (uint32_t) i), "r"((uint32_t) v));
  }
  if (tid == 0) {
    mbar_init(smem_u32(&bar), 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc<32>(smem_u32(&tmem_base_s));
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = tmem_base_s;
  if (tid == 0) {
    const uint32_t start = a_base + (uint32_t) shift * 32 + (uint32_t) row_shift_bytes;
    const uint64_t ad = desc_sw32(start, 256, use_base_off ? (start >> 7) & 7 : 0);
    const uint64_t bd = umma_desc_kmajor_noswizzle(b_base, 32 * 16, 128);
    umma_i8(tmem_base, ad, bd, umma_idesc_i8(128, 32, false, false), 0);
    umma_commit(smem_u32(&bar));
  }
  mbar_wait(smem_u32(&bar), 0);
  tc_fence_after_sync();
  int32_t v[32];
  tmem_ld32(tmem_base + ((uint32_t) (warp * 32) << 16), v);
  tmem_ld_wait();
  for (int n = 0; n < 32; n++) out[tid * 32 + n] = v[n];
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc<32>(tmem_base);
}

int main() {
  int32_t* d;
  cudaMalloc(&d, 128 * 32 * 4);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 40000);
  static int32_t h[128 * 32];
  for (int use_bo = 0; use_bo < 2; use_bo++)
    for (int shift = 0; shift < 10; shift++) {
      probe<<<1, 128, 40000>>>(d, shift, use_bo, 0);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("base_off %d shift %d: CUDA error %s\n", use_bo, shift, cudaGetErrorString(e)); return 1; }
      cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
      int bad = 0, swapped = 0;
      for (int m = 0; m < 128; m++)
        for (int n = 0; n < 32; n++) {
          const int q = m + shift;
          if (h[m * 32 + n] != ((q * 7 + n * 3) & 0x7f)) bad++;
          if (h[m * 32 + n] == ((q * 7 + (n ^ 16) * 3) & 0x7f)) swapped++;
        }
      printf("base_offset %s  start = base + %d px: %4d of 4096 wrong (%d equal the chunk-swapped value)\n", use_bo ? "set  " : "zero ", shift, bad, swapped);
    }
  return 0;
}
