// Throughput of the requantisation instruction mix on one SM's pipes, as a function of resident warps.
// Each thread requantises 16 register-resident values per iteration with the library's "U" form; no memory traffic.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o requant_pipes requant_pipes.cu && ./requant_pipes
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(1024) k(uint32_t* out, uint32_t m2, uint64_t k2, int sh, int sm, int iters, uint32_t seed) {
  uint32_t v[16];
#pragma unroll
  for (int i = 0; i < 16; i++) v[i] = seed * (threadIdx.x + 1) + i * 977u;
  uint32_t accum = 0;
  for (int it = 0; it < iters; it++) {
    uint32_t o[4];
#pragma unroll
    for (int t = 0; t < 4; t++) {
      int32_t y[4];
#pragma unroll
      for (int i = 0; i < 4; i++) {
        uint32_t nu = v[4 * t + i] + accum;                      // (keeps the chain alive across iterations: 1 ALU op)
        if (MODE == 0) {                                         // full U form, balanced shift
          uint32_t hi = (uint32_t) (((uint64_t) nu * m2 + k2) >> 32);
          int32_t tt = (int32_t) (hi + (nu >> 31));
          y[i] = (i & 1) ? (tt >> sh) : __mulhi(tt, sm);
        } else if (MODE == 1) {                                  // only the 64-bit multiply-add high word
          y[i] = (int32_t) (uint32_t) (((uint64_t) nu * m2 + k2) >> 32);
        } else if (MODE == 2) {                                  // only ALU ops: LEA.HI + SHF
          int32_t tt = (int32_t) (nu + (nu >> 31));
          y[i] = tt >> sh;
        } else {                                                 // U form, all shifts on the ALU
          uint32_t hi = (uint32_t) (((uint64_t) nu * m2 + k2) >> 32);
          int32_t tt = (int32_t) (hi + (nu >> 31));
          y[i] = tt >> sh;
        }
      }
      uint32_t r;
      asm("{ .reg .b32 t; cvt.pack.sat.u8.s32.b32 t, %4, %3, 0; cvt.pack.sat.u8.s32.b32 %0, %2, %1, t; }"
          : "=r"(r) : "r"(y[0]), "r"(y[1]), "r"(y[2]), "r"(y[3]));
      o[t] = r;
    }
    accum += o[0] ^ o[1] ^ o[2] ^ o[3];
  }
  if (accum == 0x12345678u) out[threadIdx.x] = accum;
}

template <int MODE>
void run(const char* name, int sms) {
  uint32_t* d;
  cudaMalloc(&d, 4096);
  const int iters = 20000;
  for (int warps : {4, 8, 16, 24, 32}) {
    cudaEvent_t a, b;
    cudaEventCreate(&a), cudaEventCreate(&b);
    k<MODE><<<sms, warps * 32>>>(d, 0x9abcdef1u, 0x123456789abcull, 9, 1 << 23, 100, 3);
    cudaEventRecord(a);
    k<MODE><<<sms, warps * 32>>>(d, 0x9abcdef1u, 0x123456789abcull, 9, 1 << 23, iters, 3);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    int clk_khz;
    cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    const double values = (double) sms * warps * 32 * 16 * iters;
    printf("%-34s warps/SM %2d: %7.3f ms  %6.2f values/clk/SM (at %d MHz)\n", name, warps, ms,
           values / (ms * 1e-3) / sms / (clk_khz * 1e3), clk_khz / 1000);
  }
  cudaFree(d);
}

int main() {
  int sms;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  run<0>("U form (balanced shift) + pack", sms);
  run<3>("U form (ALU shift) + pack", sms);
  run<1>("IMAD.HI.U32 only + pack", sms);
  run<2>("LEA.HI + SHF only + pack", sms);
  return 0;
}
