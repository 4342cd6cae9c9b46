// Flag ping-pong between two GPUs over peer memory: what one signal/wait round trip costs, (a) inside one
// long-running kernel per GPU, (b) with the library's structure: a signal kernel and a wait kernel per round.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/flag_latency tools/flag_latency.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <chrono>
#include <thread>

__device__ __forceinline__ void st_rel(uint32_t* p, uint32_t v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t ld_acq(const uint32_t* p) { uint32_t v; asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }

__global__ void pingpong(uint32_t* mine, uint32_t* peer, int rounds, int first, long long* cycles) {
  long long t0 = clock64();
  for (int i = 1; i <= rounds; ++i) {
    if (first) { st_rel(peer, i); while (ld_acq(mine) < (uint32_t)i) {} }
    else { while (ld_acq(mine) < (uint32_t)i) {} st_rel(peer, i); }
  }
  *cycles = clock64() - t0;
}
__global__ void signal_k(uint32_t* peer, uint32_t* mine_self, uint32_t e) { __threadfence_system(); st_rel(peer, e); st_rel(mine_self, e); }
__global__ void wait_k(const uint32_t* a, const uint32_t* b, uint32_t e, int sleep) {
  while (ld_acq(a) < e || ld_acq(b) < e) { if (sleep) __nanosleep(200); }
}

int main() {
  int n = 0; cudaGetDeviceCount(&n);
  if (n < 2) { printf("needs 2 GPUs\n"); return 0; }
  uint32_t* f[2]; long long* cyc[2];
  for (int d = 0; d < 2; ++d) { cudaSetDevice(d); cudaDeviceEnablePeerAccess(1 - d, 0); cudaMalloc(&f[d], 256); cudaMemset(f[d], 0, 256); cudaMalloc(&cyc[d], 8); }
  cudaStream_t st[2];
  for (int d = 0; d < 2; ++d) { cudaSetDevice(d); cudaStreamCreateWithFlags(&st[d], cudaStreamNonBlocking); cudaDeviceSynchronize(); }
  const int rounds = 1000;
  for (int d = 0; d < 2; ++d) { cudaSetDevice(d); pingpong<<<1, 1, 0, st[d]>>>(f[d], f[1 - d], rounds, d == 0, cyc[d]); }
  for (int d = 0; d < 2; ++d) { cudaSetDevice(d); cudaStreamSynchronize(st[d]); }
  long long c = 0; cudaSetDevice(0); cudaMemcpy(&c, cyc[0], 8, cudaMemcpyDeviceToHost);
  int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  printf("in-kernel ping-pong: %.2f us per round trip (%d rounds)\n", (double)c / rounds / (khz / 1000.0), rounds);
  // (b) kernel pairs, the two GPUs driven by two host threads
  for (int sleep = 0; sleep < 2; ++sleep) {
    for (int d = 0; d < 2; ++d) { cudaSetDevice(d); cudaMemset(f[d], 0, 256); cudaDeviceSynchronize(); }
    double ms[2] = {0, 0};
    auto run = [&](int d) {
      cudaSetDevice(d);
      cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
      const int R = 200;
      float tot = 0;
      for (int i = 1; i <= R; ++i) {
        cudaEventRecord(a, st[d]);
        signal_k<<<1, 1, 0, st[d]>>>(f[1 - d] + 8 * d, f[d] + 8 * d, (uint32_t)i);   // slot d at the peer and at home
        wait_k<<<1, 1, 0, st[d]>>>(f[d], f[d] + 8, (uint32_t)i, sleep);
        cudaEventRecord(b, st[d]);
        cudaStreamSynchronize(st[d]);
        float t; cudaEventElapsedTime(&t, a, b); tot += t;
      }
      ms[d] = tot / R;
    };
    std::thread t1(run, 1); run(0); t1.join();
    printf("signal kernel + wait kernel per round (nanosleep %d): gpu0 %.4f ms, gpu1 %.4f ms\n", sleep, ms[0], ms[1]);
  }
  return 0;
}
