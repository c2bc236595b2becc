// Instruction-rate probe for the node test of the wide traversal: how many results per clock per SM do
// I2F (byte -> float), PRMT, FFMA and FMNMX deliver on this GPU?  Build: nvcc -arch=sm_100a -O3.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256) probe(uint32_t* out, uint32_t seed, int iters) {
  uint32_t w0 = seed + threadIdx.x, w1 = w0 * 3u, w2 = w0 * 5u, w3 = w0 * 7u;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f, a4 = 0.f, a5 = 0.f, a6 = 0.f, a7 = 0.f;
  const float k = __uint_as_float(seed | 0x3f000000u);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (MODE == 0) {         // I2F from a byte lane + FADD to keep it alive
        a0 += (float)(w0 & 0xffu); a1 += (float)((w0 >> 8) & 0xffu); a2 += (float)((w0 >> 16) & 0xffu); a3 += (float)(w0 >> 24);
        a4 += (float)(w1 & 0xffu); a5 += (float)((w1 >> 8) & 0xffu); a6 += (float)((w1 >> 16) & 0xffu); a7 += (float)(w1 >> 24);
      } else if (MODE == 1) {  // PRMT into a float's mantissa + FADD
        a0 += __uint_as_float(__byte_perm(w0, 0x3f800000u, 0x7604)); a1 += __uint_as_float(__byte_perm(w0, 0x3f800000u, 0x7614));
        a2 += __uint_as_float(__byte_perm(w0, 0x3f800000u, 0x7624)); a3 += __uint_as_float(__byte_perm(w0, 0x3f800000u, 0x7634));
        a4 += __uint_as_float(__byte_perm(w1, 0x3f800000u, 0x7604)); a5 += __uint_as_float(__byte_perm(w1, 0x3f800000u, 0x7614));
        a6 += __uint_as_float(__byte_perm(w1, 0x3f800000u, 0x7624)); a7 += __uint_as_float(__byte_perm(w1, 0x3f800000u, 0x7634));
      } else if (MODE == 2) {  // FADD only (baseline of the two above)
        a0 += k; a1 += k; a2 += k; a3 += k; a4 += k; a5 += k; a6 += k; a7 += k;
      } else if (MODE == 3) {  // FFMA
        a0 = fmaf(a0, k, k); a1 = fmaf(a1, k, k); a2 = fmaf(a2, k, k); a3 = fmaf(a3, k, k);
        a4 = fmaf(a4, k, k); a5 = fmaf(a5, k, k); a6 = fmaf(a6, k, k); a7 = fmaf(a7, k, k);
      } else if (MODE == 4) {  // FMNMX
        a0 = fmaxf(a0, k + a1); a2 = fminf(a2, a3); a4 = fmaxf(a4, a5); a6 = fminf(a6, a7);
        a1 = fmaxf(a1, a0); a3 = fminf(a3, a2); a5 = fmaxf(a5, a4); a7 = fminf(a7, a6);
      }
      w0 = w0 * 1664525u + 1013904223u; w1 ^= w0;
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = __float_as_uint(a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7) ^ w2 ^ w3;
}

template <int MODE>
double run(const char* name, uint32_t* out, int sms, double clockGHz) {
  const int iters = 4096, blocks = sms * 8;
  probe<MODE><<<blocks, 256>>>(out, 12345u, 16);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  probe<MODE><<<blocks, 256>>>(out, 12345u, iters);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  double groups = (double)blocks * 256 * iters * 4;  // groups of 8 probed ops (+ 2 integer ops of the LCG)
  double perClkSm = groups * 8 / (ms * 1e-3) / (clockGHz * 1e9) / sms;
  printf("%-28s %8.3f ms   %7.1f probed ops / clk / SM (at %.3f GHz)\n", name, ms, perClkSm, clockGHz);
  return perClkSm;
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  int clk = 0;
  cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  double ghz = clk * 1e-6;
  uint32_t* out;
  cudaMalloc(&out, (size_t)p.multiProcessorCount * 8 * 256 * 4);
  printf("%s, %d SMs\n", p.name, p.multiProcessorCount);
  run<2>("FADD x8", out, p.multiProcessorCount, ghz);
  run<0>("I2F.U8 + FADD x8", out, p.multiProcessorCount, ghz);
  run<1>("PRMT + FADD x8", out, p.multiProcessorCount, ghz);
  run<3>("FFMA x8", out, p.multiProcessorCount, ghz);
  run<4>("FMNMX x8", out, p.multiProcessorCount, ghz);
  return 0;
}
