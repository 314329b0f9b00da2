// Microbenchmark: the SeedExtension key chain (key, r = split(key): two threefry2x32 blocks per step, one chain per
// lane) with the additions / rotations on the ALU pipe (IADD3, SHF, LOP3) or moved to the FMA pipe (IMAD, IMAD.WIDE).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tf_chain_bench tf_chain_bench.cu && ./tf_chain_bench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__constant__ uint32_t c_one = 1u;
__constant__ uint32_t c_pow[32];

template <int V> struct Ops;
// V bit0: additions as IMAD; bits 1..: number of rounds (of 20) whose rotation is an IMAD.WIDE: 0, 5, 10, 20
template <int V>
__device__ __forceinline__ void tf(uint32_t k0, uint32_t k1, uint32_t x0, uint32_t x1, uint32_t &o0, uint32_t &o1, const uint32_t one, const uint32_t *pw) {
  constexpr bool IM = V & 1;
  constexpr int NR = V >> 1;
  const uint32_t ks2 = k0 ^ k1 ^ 0x1BD11BDAu;
  auto add = [&](uint32_t a, uint32_t b) { return IM ? b * one + a : a + b; };
  int rd = 0;
  auto round = [&](int r) {
    x0 = add(x0, x1);
    // which rounds rotate on the FMA pipe: spread evenly
    const bool fm = NR == 20 || (NR == 10 && (rd & 1)) || (NR == 5 && (rd & 3) == 1) || (NR == 7 && (rd % 3) == 1);
    if (fm) {
      const unsigned long long p = (unsigned long long)x1 * (unsigned long long)pw[r];
      x1 = ((uint32_t)p | (uint32_t)(p >> 32)) ^ x0;
    } else {
      x1 = __funnelshift_l(x1, x1, r) ^ x0;
    }
    ++rd;
  };
  x0 = add(x0, k0); x1 = add(x1, k1);
  round(13); round(15); round(26); round(6);
  x0 = add(x0, k1); x1 = add(x1, ks2 + 1u);
  round(17); round(29); round(16); round(24);
  x0 = add(x0, ks2); x1 = add(x1, k0 + 2u);
  round(13); round(15); round(26); round(6);
  x0 = add(x0, k0); x1 = add(x1, k1 + 3u);
  round(17); round(29); round(16); round(24);
  x0 = add(x0, k1); x1 = add(x1, ks2 + 4u);
  round(13); round(15); round(26); round(6);
  x0 = add(x0, ks2); x1 = add(x1, k0 + 5u);
  o0 = x0; o1 = x1;
}

template <int V>
__global__ void __launch_bounds__(128) chain_kernel(uint32_t *out, int steps) {
  __shared__ uint32_t pw[32];
  if (threadIdx.x < 32) pw[threadIdx.x] = c_pow[threadIdx.x];
  __syncthreads();
  uint32_t pr[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) pr[i] = pw[i];  // only the 8 used rotation counts survive as registers
  const uint32_t one = c_one;
  uint32_t k0 = blockIdx.x * blockDim.x + threadIdx.x, k1 = 42u, acc = 0;
  for (int s = 0; s < steps; ++s) {
    uint32_t n0, r0, n1, r1;
    tf<V>(k0, k1, 0u, 2u, n0, r0, one, pr);
    tf<V>(k0, k1, 1u, 3u, n1, r1, one, pr);
    acc ^= r0 + r1;
    k0 = n0; k1 = n1;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc ^ k0 ^ k1;
}

template <int V>
void run(const char *name, uint32_t *out, uint32_t *ref) {
  const int steps = 2000;
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  const int cfg[][2] = {{1, 32}, {148, 32}, {148, 128}, {148 * 2, 128}, {148 * 4, 128}, {148 * 8, 128}};  // ctas, threads
  printf("%-28s", name);
  for (auto &c : cfg) {
    chain_kernel<V><<<c[0], c[1]>>>(out, 200);
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    chain_kernel<V><<<c[0], c[1]>>>(out, steps);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    const double warps_per_smsp = (double)c[0] * c[1] / 32 / (c[0] >= 148 ? 148.0 : 1.0) / 4;
    // cycles per chain step per warp at 1.965 GHz, and per SMSP (throughput view)
    const double cyc = ms * 1e-3 * 1.965e9 / steps;
    printf(" | %5.2f w/smsp: %6.1f cyc/step, %6.1f cyc/step/warp-slot", warps_per_smsp, cyc, cyc / (warps_per_smsp < 1 ? 1 : warps_per_smsp));
  }
  uint32_t h[32];
  chain_kernel<V><<<1, 32>>>(out, 50);
  cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
  if (ref[0] == 0 && ref[1] == 0) { for (int i = 0; i < 32; ++i) ref[i] = h[i]; }
  bool same = true;
  for (int i = 0; i < 32; ++i) same &= ref[i] == h[i];
  printf(" | %s\n", same ? "same" : "DIFFERENT");
}

int main() {
  uint32_t pw[32];
  for (int i = 0; i < 32; ++i) pw[i] = 1u << i;
  cudaMemcpyToSymbol(c_pow, pw, sizeof(pw));
  uint32_t *out;
  cudaMalloc(&out, 148 * 8 * 128 * 4);
  uint32_t ref[32] = {0};
  run<0>("plain (IADD3/SHF/LOP3)", out, ref);
  run<1>("IMAD adds", out, ref);
  run<(5 << 1)>("rot IMAD.WIDE 5/20", out, ref);
  run<(7 << 1)>("rot IMAD.WIDE 7/20", out, ref);
  run<(10 << 1)>("rot IMAD.WIDE 10/20", out, ref);
  run<(20 << 1)>("rot IMAD.WIDE 20/20", out, ref);
  run<(5 << 1) | 1>("IMAD adds + rot 5/20", out, ref);
  run<(10 << 1) | 1>("IMAD adds + rot 10/20", out, ref);
  run<(20 << 1) | 1>("IMAD adds + rot 20/20", out, ref);
  return 0;
}
