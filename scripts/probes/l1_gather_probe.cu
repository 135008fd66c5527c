// Micro-benchmark: what does one warp-wide gather cost on the L1 data pipe of a B200 SM, as a function of the access
// width and of how the 32 addresses fall into 128-byte lines / 32-byte sectors?  All data is L1-resident (64 KiB
// table), every block runs the same pattern, loads are independent (8 in flight per thread).  Prints SM cycles per
// warp-level load instruction at full SM load -- the "data wavefronts" a load of that shape costs.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o l1_gather_probe l1_gather_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int W> struct Vec;
template <> struct Vec<4>  { float v[1]; };
template <> struct Vec<8>  { float v[2]; };
template <> struct Vec<16> { float v[4]; };
template <> struct Vec<32> { float v[8]; };

template <int W> __device__ __forceinline__ float load(const char *p);
template <> __device__ __forceinline__ float load<4>(const char *p) { float a; asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(a) : "l"(p)); return a; }
template <> __device__ __forceinline__ float load<8>(const char *p) { float a, b; asm volatile("ld.global.nc.v2.f32 {%0,%1}, [%2];" : "=f"(a), "=f"(b) : "l"(p)); return a + b; }
template <> __device__ __forceinline__ float load<16>(const char *p) { float a, b, c, d; asm volatile("ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(a), "=f"(b), "=f"(c), "=f"(d) : "l"(p)); return a + b + c + d; }
template <> __device__ __forceinline__ float load<32>(const char *p) {
  float a, b, c, d, e, f, g, h;
  asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=f"(a), "=f"(b), "=f"(c), "=f"(d), "=f"(e), "=f"(f), "=f"(g), "=f"(h) : "l"(p));
  return a + b + c + d + e + f + g + h;
}

__device__ __forceinline__ int pattern_offset(int pat, int t) {   // byte offset of lane t inside a 4 KiB window (32-byte granules)
  switch (pat) {
    case 0: return t * 32;                               // coalesced: 8 lines x 4 sectors
    case 1: return t * 128;                              // 32 lines, 1 sector each
    case 2: return (t / 2) * 128 + (t % 2) * 32;         // 16 lines, 2 sectors each
    case 3: return (t / 2) * 128 + (t % 2) * 64;         // 16 lines, 2 sectors each, other banks
    case 4: return (t / 2) * 32;                         // 16 distinct sectors, pairs of lanes share one
    case 5: return ((t * 5) % 32) * 32;                  // coalesced set, shuffled lane order
    case 6: return (t % 8) * 128 + (t / 8) * 32;         // 8 lines x 4 sectors, strided lane order
    case 7: return 0;                                    // broadcast
    case 8: return (t / 4) * 128 + (t % 4) * 32 + 2048;  // coalesced again (control)
    case 9: return (t / 8) * 512 + (t % 8) * 32;         // 4 groups of 8 lanes, each 2 lines (sorted-column like)
    default: return t * 160 % 4096 / 32 * 32;            // 32 distinct sectors in ~27 lines
  }
}

template <int W>
__global__ void probe(const char *tab, int pat, int iters, int step, float *out) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const char *p = tab + pattern_offset(pat, lane) + (warp & 3) * 4096;
  float acc = 0.f;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int k = 0; k < 8; k++) acc += load<W>(p + (((i * 8 + k) * step) & 0xC000));   // step is a run-time value: nothing can be hoisted
  }
  if (acc == 12345.678f) out[0] = acc;
}

int main() {
  char *tab; float *out;
  cudaMalloc(&tab, 1 << 18); cudaMemset(tab, 0, 1 << 18); cudaMalloc(&out, 4);
  cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  const int iters = 2000, blocks = pr.multiProcessorCount * 2, threads = 512;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  printf("# %s, %d SMs, nominal %d kHz; cycles are at the nominal clock (relative numbers matter)\n", pr.name, pr.multiProcessorCount, clk);
  printf("%-8s", "pattern");
  for (int w : {4, 8, 16, 32}) printf("  W=%-2d cyc/ld", w);
  printf("\n");
  for (int pat = 0; pat <= 10; pat++) {
    printf("%-8d", pat);
    for (int w : {4, 8, 16, 32}) {
      auto run = [&](int it) {
        if (w == 4) probe<4><<<blocks, threads>>>(tab, pat, it, 16384, out);
        else if (w == 8) probe<8><<<blocks, threads>>>(tab, pat, it, 16384, out);
        else if (w == 16) probe<16><<<blocks, threads>>>(tab, pat, it, 16384, out);
        else probe<32><<<blocks, threads>>>(tab, pat, it, 16384, out);
      };
      run(50); cudaDeviceSynchronize();
      cudaEventRecord(e0); run(iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      const double loads_per_sm = (double)iters * 8 * (threads / 32) * 2;   // warp-level loads per SM
      printf("  %10.2f", ms * 1e-3 * clk * 1e3 / loads_per_sm);
    }
    printf("\n");
  }
  return 0;
}
