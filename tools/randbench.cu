// randbench.cu — measures the B200 random-access ceiling the FM-index stages are bounded by:
// independent dependent-chain gathers of 32-byte (one sector) or 64-byte granules over tables of
// several sizes.  Each thread follows a pseudo-random pointer chain (like successive rank queries).
// Output: one JSON line per (table size, granule, chains) with sectors/s and GB/s.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

template <int GRAN>  // bytes per access: 32 or 64
__global__ void chase(const uint4* __restrict__ tab, uint64_t n_gran, int steps, uint64_t* out) {
  uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  uint64_t x = t * 0x9E3779B97F4A7C15ull + 0x1234567;
  uint64_t acc = 0;
  for (int s = 0; s < steps; ++s) {
    x ^= x >> 29; x *= 0xBF58476D1CE4E5B9ull; x ^= x >> 32;
    uint64_t g = x % n_gran;
    const uint4* p = tab + g * (GRAN / 16);
    uint4 a = __ldg(p), b = __ldg(p + 1);
    uint64_t v = a.x ^ a.w ^ b.y ^ b.z;
    if (GRAN == 64) { uint4 c = __ldg(p + 2), d = __ldg(p + 3); v ^= c.x ^ d.w; }
    acc += v;
    x += v;  // make the next address depend on the loaded data
  }
  if (acc == 0xdeadbeef) out[0] = acc;
}

int main(int argc, char** argv) {
  // optional: L2 fetch granularity hint (bytes) as argv[1]
  if (argc > 1) {
    size_t g = atoi(argv[1]);
    cudaError_t e = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, g);
    size_t got = 0;
    cudaDeviceGetLimit(&got, cudaLimitMaxL2FetchGranularity);
    printf("{\"l2_fetch_granularity_requested\":%zu,\"got\":%zu,\"err\":\"%s\"}\n", g, got, cudaGetErrorString(e));
  } else {
    size_t got = 0;
    cudaDeviceGetLimit(&got, cudaLimitMaxL2FetchGranularity);
    printf("{\"l2_fetch_granularity_default\":%zu}\n", got);
  }
  size_t sizes_mb[] = {64, 2048, 8192};
  int steps = 64;
  for (size_t smb : sizes_mb) {
    size_t bytes = smb << 20;
    uint4* tab; uint64_t* out;
    if (cudaMalloc(&tab, bytes) != cudaSuccess) { printf("{\"error\":\"alloc %zu MB\"}\n", smb); continue; }
    cudaMalloc(&out, 8);
    cudaMemset(tab, 1, bytes);
    for (int gran : {32, 64}) {
      for (int blocks_per_sm : {8}) {
        int threads = 256, grid = 148 * blocks_per_sm * 8;  // 8 waves
        uint64_t n_gran = bytes / gran;
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        for (int rep = 0; rep < 3; ++rep) {
          cudaEventRecord(e0);
          if (gran == 32) chase<32><<<grid, threads>>>(tab, n_gran, steps, out);
          else chase<64><<<grid, threads>>>(tab, n_gran, steps, out);
          cudaEventRecord(e1); cudaEventSynchronize(e1);
        }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double acc = (double)grid * threads * steps;
        printf("{\"table_mb\":%zu,\"granule\":%d,\"threads\":%d,\"ms\":%.3f,\"accesses_per_s\":%.4g,\"GBps\":%.1f}\n",
               smb, gran, grid * threads, ms, acc / (ms * 1e-3), acc * gran / (ms * 1e-3) / 1e9);
      }
    }
    cudaFree(tab); cudaFree(out);
  }
  return 0;
}
