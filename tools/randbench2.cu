// randbench2.cu — which load flavour gives the highest random 32-byte-sector rate on B200?
// Each thread follows a dependent pseudo-random chain over an 8 GB table, one 32-B granule per step.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

template <int MODE>
__device__ __forceinline__ uint64_t load32(const uint4* p) {
  uint4 a, b;
  if (MODE == 0) { a = __ldg(p); b = __ldg(p + 1); }                       // ld.global.nc x2
  else if (MODE == 1) { a = p[0]; b = p[1]; }                               // ld.global x2
  else if (MODE == 2) {                                                     // ld.global.cg (L2 only)
    asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w) : "l"(p));
    asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w) : "l"(p + 1));
  } else if (MODE == 3) {                                                   // nc + L1::no_allocate
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w) : "l"(p));
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w) : "l"(p + 1));
  } else if (MODE == 4) {                                                   // one 256-bit load
    uint32_t r0, r1, r2, r3, r4, r5, r6, r7;
    asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3), "=r"(r4), "=r"(r5), "=r"(r6), "=r"(r7) : "l"(p));
    a = make_uint4(r0, r1, r2, r3); b = make_uint4(r4, r5, r6, r7);
  } else if (MODE == 5) {                                                   // 16 bytes only (half a sector)
    a = __ldg(p); b = make_uint4(0, 0, 0, 0);
  } else {                                                                  // 8 bytes only
    uint2 v = __ldg(reinterpret_cast<const uint2*>(p)); a = make_uint4(v.x, v.y, 0, 0); b = make_uint4(0, 0, 0, 0);
  }
  return (uint64_t)(a.x ^ a.w ^ b.y ^ b.z);
}

template <int MODE>
__global__ void chase(const uint4* __restrict__ tab, uint64_t n_gran, int steps, uint64_t* out) {
  uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  uint64_t x = t * 0x9E3779B97F4A7C15ull + 0x1234567;
  uint64_t acc = 0;
  for (int s = 0; s < steps; ++s) {
    x ^= x >> 29; x *= 0xBF58476D1CE4E5B9ull; x ^= x >> 32;
    uint64_t g = x % n_gran;
    uint64_t v = load32<MODE>(tab + g * 2);
    acc += v; x += v;
  }
  if (acc == 0xdeadbeef) out[0] = acc;
}

template <int MODE>
void run(const char* name, const uint4* tab, uint64_t n_gran, uint64_t* out) {
  int threads = 256, grid = 148 * 8 * 8, steps = 64;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float ms = 0;
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0);
    chase<MODE><<<grid, threads>>>(tab, n_gran, steps, out);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
  }
  cudaError_t e = cudaGetLastError();
  double acc = (double)grid * threads * steps;
  printf("{\"mode\":\"%s\",\"ms\":%.3f,\"accesses_per_s\":%.4g,\"err\":\"%s\"}\n", name, ms, acc / (ms * 1e-3), cudaGetErrorString(e));
}

int main() {
  size_t bytes = (size_t)8 << 30;
  uint4* tab; uint64_t* out;
  cudaMalloc(&tab, bytes); cudaMalloc(&out, 8); cudaMemset(tab, 1, bytes);
  uint64_t n_gran = bytes / 32;
  run<0>("ldg_nc_2x16", tab, n_gran, out);
  run<1>("ld_2x16", tab, n_gran, out);
  run<2>("ld_cg_2x16", tab, n_gran, out);
  run<3>("nc_noalloc_2x16", tab, n_gran, out);
  run<4>("nc_v8_32B", tab, n_gran, out);
  run<5>("ldg_16B_only", tab, n_gran, out);
  run<6>("ldg_8B_only", tab, n_gran, out);
  return 0;
}
