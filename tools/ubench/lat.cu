// Developer micro-benchmarks (B200): dependent-issue latencies of the instructions the fused decode kernel chains.
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdio.h>
#include <stdint.h>
#define N 256
__device__ __forceinline__ void mma16816(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__global__ void k(long long* out, float* sink, int nwarps_active) {
  __shared__ __align__(16) uint8_t sm[8192];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 2048; i += blockDim.x) ((uint32_t*)sm)[i] = i * 2654435761u >> 20;
  __syncthreads();
  if (warp >= nwarps_active) return;
  uint32_t a[4] = {0x3f803f80u, 0x3f803f80u, 0x3f803f80u, 0x3f803f80u}; uint32_t b0 = 0x3f803f80u, b1 = 0x3f803f80u;
  float c[4] = {0, 0, 0, 0}, d[4] = {0,0,0,0};
  long long t0, t1;
  // 1. dependent HMMA chain
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) mma16816(c, a, b0, b1);
  t1 = clock64();
  if (threadIdx.x == 0) out[0] = t1 - t0;
  // 2. two independent HMMA chains
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) { mma16816(c, a, b0, b1); mma16816(d, a, b0, b1); }
  t1 = clock64();
  if (threadIdx.x == 0) out[1] = t1 - t0;
  // 3. dependent shuffle chain
  float v = c[0] + lane;
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) v += __shfl_xor_sync(0xffffffffu, v, 4);
  t1 = clock64();
  if (threadIdx.x == 0) out[2] = t1 - t0;
  // 4. dependent exp2f chain
  float e = v * 1e-30f;
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) e = exp2f(e - 1.0f);
  t1 = clock64();
  if (threadIdx.x == 0) out[3] = t1 - t0;
  // 5. dependent ldmatrix chain (address depends on the loaded value)
  uint32_t addr = (uint32_t)__cvta_generic_to_shared(sm) + (lane & 7) * 64;
  uint32_t r[4] = {0,0,0,0};
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr + ((r[0] & 1) << 4)));
  }
  t1 = clock64();
  if (threadIdx.x == 0) out[4] = t1 - t0;
  // 6. dependent LDS.32 chain
  uint32_t idx = lane;
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) idx = ((volatile uint32_t*)sm)[idx & 2047];
  t1 = clock64();
  if (threadIdx.x == 0) out[5] = t1 - t0;
  // 7. dependent FFMA chain
  float f = e;
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) f = fmaf(f, 1.0001f, 0.5f);
  t1 = clock64();
  if (threadIdx.x == 0) out[6] = t1 - t0;
  // 8. bf16 convert + pack chain
  float g2 = f;
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) { __nv_bfloat16 h = __float2bfloat16_rn(g2); g2 = g2 - __bfloat162float(h) + 1.0f; }
  t1 = clock64();
  if (threadIdx.x == 0) out[7] = t1 - t0;
  // 9. clock64 back-to-back overhead
  t0 = clock64();
  long long acc = 0;
#pragma unroll 16
  for (int i = 0; i < N; ++i) acc += clock64();
  t1 = clock64();
  if (threadIdx.x == 0) out[8] = t1 - t0;
  // 10. HMMA -> FADD -> pack -> HMMA dependent round (the attention chain shape)
  t0 = clock64();
#pragma unroll 8
  for (int i = 0; i < N; ++i) { mma16816(c, a, b0, b1); float s = c[0] + c[1]; __nv_bfloat162 hh = __floats2bfloat162_rn(s, s); b0 = *reinterpret_cast<uint32_t*>(&hh) & 0x3f803f80u; }
  t1 = clock64();
  if (threadIdx.x == 0) out[9] = t1 - t0;
  sink[threadIdx.x] = c[0] + c[1] + d[0] + v + e + r[0] + idx + f + g2 + (float)acc + b0;
}
// mbarrier try_wait on an already-completed phase, and a named barrier among 256 threads
__global__ void k2(long long* out) {
  __shared__ uint64_t bar;
  uint32_t b = (uint32_t)__cvta_generic_to_shared(&bar);
  if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b)); }
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < 64; ++i) {
    if (threadIdx.x == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(b) : "memory");
    uint32_t done = 0;
    while (!done) asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0,1,0,p;\n}" : "=r"(done) : "r"(b), "r"(i & 1) : "memory");
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) out[16] = t1 - t0;
  t0 = clock64();
  for (int i = 0; i < 64; ++i) asm volatile("bar.sync 1, 256;" ::: "memory");
  t1 = clock64();
  if (threadIdx.x == 0) out[17] = t1 - t0;
}
int main() {
  long long* out; float* sink; cudaMalloc(&out, 256); cudaMalloc(&sink, 4096); cudaMemset(out, 0, 256);
  const char* names[] = {"HMMA dep chain", "2 indep HMMA chains (per pair)", "SHFL+FADD dep", "exp2f(x-1) dep", "ldmatrix.x4 dep", "LDS.32 dep", "FFMA dep", "bf16 cvt round trip dep", "clock64+add", "HMMA->FADD->pack->HMMA"};
  for (int nw = 1; nw <= 8; nw *= 8) {
    k<<<1, 256>>>(out, sink, nw); cudaDeviceSynchronize();
    long long h[32]; cudaMemcpy(h, out, 256, cudaMemcpyDeviceToHost);
    printf("active warps %d (%s)\n", nw, cudaGetErrorString(cudaGetLastError()));
    for (int i = 0; i < 10; ++i) printf("  %-34s %7.1f cyc/iter\n", names[i], (double)h[i] / N);
  }
  k2<<<1, 256>>>(out); cudaDeviceSynchronize();
  long long h[32]; cudaMemcpy(h, out, 256, cudaMemcpyDeviceToHost);
  printf("mbarrier arrive + try_wait (256 thr): %.1f cyc/iter; bar.sync 1,256: %.1f cyc/iter\n", h[16] / 64.0, h[17] / 64.0);
  return 0;
}
