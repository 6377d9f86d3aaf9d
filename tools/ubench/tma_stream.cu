// Micro-benchmark: how fast can ONE SM stream global memory into shared memory with bulk async copies (the decode kernel's
// producer pattern), as a function of the bytes in flight per SM, the number of SMs streaming, and whether the data sits in L2?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_stream tools/ubench/tma_stream.cu && ./tma_stream
// Each CTA (1 per SM, 1 thread issues) walks its own region with a ring of NS stages of STAGE bytes: wait(full[s]) -> re-issue.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done) asm volatile("{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\tselp.u32 %0, 1, 0, P1;\n\t}\n" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

__global__ void __launch_bounds__(32, 1) stream_kernel(const uint8_t* base, size_t region, int stage_bytes, int ns, int iters, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t full[16];
  if (threadIdx.x != 0) return;
  for (int i = 0; i < ns; ++i) mbar_init(smem_u32(&full[i]), 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  const uint8_t* src = base + (size_t)blockIdx.x * region;
  const int per_region = (int)(region / stage_bytes);
  const long long t0 = clock64();
  for (int i = 0; i < ns && i < iters; ++i) {
    mbar_expect_tx(smem_u32(&full[i]), stage_bytes);
    bulk_g2s(smem_u32(smem + (size_t)i * stage_bytes), src + (size_t)(i % per_region) * stage_bytes, stage_bytes, smem_u32(&full[i]));
  }
  int slot = 0; uint32_t phase = 0;
  for (int i = 0; i < iters; ++i) {
    mbar_wait(smem_u32(&full[slot]), phase);
    const int nxt = i + ns;
    if (nxt < iters) {
      mbar_expect_tx(smem_u32(&full[slot]), stage_bytes);
      bulk_g2s(smem_u32(smem + (size_t)slot * stage_bytes), src + (size_t)(nxt % per_region) * stage_bytes, stage_bytes, smem_u32(&full[slot]));
    }
    if (++slot == ns) { slot = 0; phase ^= 1; }
  }
  cycles[blockIdx.x] = clock64() - t0;
}

int main() {
  int dev = 0; cudaSetDevice(dev);
  cudaDeviceProp p; cudaGetDeviceProperties(&p, dev);
  const size_t total = (size_t)4 << 30;                      // 4 GiB >> L2
  uint8_t* buf; cudaMalloc(&buf, total); cudaMemset(buf, 1, total);
  long long* cyc; cudaMallocManaged(&cyc, 256 * sizeof(long long));
  cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  printf("%s, %d SMs, clock %d MHz\n", p.name, p.multiProcessorCount, p.clockRate / 1000);
  printf("%-10s %5s %4s %9s | %10s %12s %10s\n", "source", "SMs", "NS", "stage KB", "GB/s", "B/clk/SM", "us");
  const int sm_list[] = {1, 16, 64, 104, 148};
  const int ns_list[] = {1, 2, 4, 6};
  for (int l2 = 0; l2 < 2; ++l2)
    for (int sms : sm_list)
      for (int ns : ns_list) {
        const int stage = 32768;
        const size_t region = l2 ? (size_t)256 << 10 : total / 148 / stage * stage;      // L2 case: 256 KB per SM, re-read
        const int iters = l2 ? 2048 : 1024;                                              // 64 / 32 MB per SM
        cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
        float best = 1e9f; long long cmax = 0;
        for (int rep = 0; rep < 3; ++rep) {
          cudaEventRecord(a);
          stream_kernel<<<sms, 32, ns * stage>>>(buf, region, stage, ns, iters, cyc);
          cudaEventRecord(b); cudaEventSynchronize(b);
          float ms; cudaEventElapsedTime(&ms, a, b);
          if (ms < best) { best = ms; cmax = 0; for (int i = 0; i < sms; ++i) cmax = cyc[i] > cmax ? cyc[i] : cmax; }
        }
        if (cudaGetLastError() != cudaSuccess) { printf("launch failed\n"); return 1; }
        const double bytes = (double)sms * iters * stage;
        printf("%-10s %5d %4d %9d | %10.1f %12.2f %10.1f\n", l2 ? "L2" : "HBM", sms, ns, stage / 1024, bytes / best / 1e6, (double)iters * stage / cmax, best * 1e3);
      }
  return 0;
}
