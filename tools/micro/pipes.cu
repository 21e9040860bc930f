// Developer microbenchmark: issue rate of the instruction classes the fused layer's SIMT roles use,
// per SM sub-partition, at 1 / 2 / 4 warps per sub-partition (independent chains, 8 per thread).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu && ./pipes
#include <cstdio>
#include <cstdint>
#include <cuda_fp16.h>

constexpr int kIters = 2048, kChains = 8;

template <int OP>
__global__ void __launch_bounds__(1024, 1) bench(uint32_t *out, long long *cycles, uint32_t seed) {
  extern __shared__ uint32_t sm[];
  uint32_t a[kChains];
  unsigned long long w[kChains];
#pragma unroll
  for (int i = 0; i < kChains; ++i) { a[i] = seed + i * 7 + threadIdx.x; w[i] = a[i] * 0x100000001ull; }
  uint32_t b = seed * 3 + 1, c = 0x3c003c00u;
  unsigned long long wb = 0x3f8000003f800000ull;
  const uint32_t saddr = uint32_t(__cvta_generic_to_shared(sm)) + (threadIdx.x & 1023) * 8;
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < kIters; ++it) {
#pragma unroll
    for (int i = 0; i < kChains; ++i) {
      if (OP == 0) asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(c), "r"(b));
      if (OP == 1) asm volatile("fma.rn.relu.f16x2 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(c), "r"(b));
      if (OP == 2) asm volatile("add.rn.f16x2 %0, %0, %1;" : "+r"(a[i]) : "r"(b));
      if (OP == 3) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(w[i]) : "l"(wb));
      if (OP == 4) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(w[i]) : "l"(wb));
      if (OP == 5) asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+r"(a[i]) : "r"(b));
      if (OP == 6) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b), "r"(c));
      if (OP == 7) asm volatile("cvt.rn.f16x2.f32 %0, %0, %1;" : "+r"(a[i]) : "r"(b));
      if (OP == 8) asm volatile("selp.b32 %0, %0, %1, p;" : "+r"(a[i]) : "r"(b));   // unused
      if (OP == 9) { uint32_t x, y; asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(x), "=r"(y) : "r"(saddr + i * 8192)); a[i] ^= x + y; }
      if (OP == 10) asm volatile("st.shared.v2.u32 [%0], {%1, %1};" :: "r"(saddr + i * 8192), "r"(a[i]) : "memory");
      if (OP == 11) a[i] = __shfl_sync(0xffffffffu, a[i], (i + it) & 31);
      if (OP == 12) asm volatile("add.rn.f32 %0, %0, %1;" : "+r"(a[i]) : "r"(b));
      if (OP == 20) { if (i & 1) asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(c), "r"(b)); else asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b), "r"(c)); }
      if (OP == 21) { if (i & 1) asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(c), "r"(b)); else asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+r"(a[i]) : "r"(b)); }
      if (OP == 22) { if (i & 1) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b), "r"(c)); else asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+r"(a[i]) : "r"(b)); }
      if (OP == 23) { if (i & 1) asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(c), "r"(b)); else asm volatile("max.f16x2 %0, %0, %1;" : "+r"(a[i]) : "r"(b)); }
      if (OP == 24) asm volatile("max.f16x2 %0, %0, %1;" : "+r"(a[i]) : "r"(b));
      if (OP == 25) { if ((i & 3) == 0) { uint32_t x, y; asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(x), "=r"(y) : "r"(saddr + i * 8192)); a[i] ^= x + y; } else if (i & 1) asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(c), "r"(b)); else asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b), "r"(c)); }
      if (OP == 26) { if (i & 1) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(w[i]) : "l"(wb)); else asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b), "r"(c)); }
      if (OP == 27) { if (i & 1) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(w[i]) : "l"(wb)); else asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(c), "r"(b)); }
      if (OP == 28) asm volatile("add.s32 %0, %0, %1;" : "+r"(a[i]) : "r"(b));
      if (OP == 29) asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c));
    }
  }
  const long long t1 = clock64();
  uint32_t r = 0;
#pragma unroll
  for (int i = 0; i < kChains; ++i) r ^= a[i] ^ uint32_t(w[i]) ^ uint32_t(w[i] >> 32);
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

template <int OP>
void run(const char *name) {
  uint32_t *out; long long *cyc, h;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
  cudaFuncSetAttribute(bench<OP>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536 + 8192);
  printf("%-16s", name);
  for (int warps : {4, 8, 16, 32}) {
    bench<OP><<<148, warps * 32, 65536 + 8192>>>(out, cyc, 1);
    cudaDeviceSynchronize();
    bench<OP><<<148, warps * 32, 65536 + 8192>>>(out, cyc, 2);
    cudaDeviceSynchronize();
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    // cycles per warp instruction per sub-partition
    printf("  %dw/smsp: %.2f cyc/instr", warps / 4, double(h) / (double(kIters) * kChains * (warps / 4)));
  }
  printf("  (%s)\n", cudaGetErrorString(cudaGetLastError()));
  cudaFree(out); cudaFree(cyc);
}

int main() {
  run<0>("HFMA2"); run<1>("HFMA2.RELU"); run<2>("HADD2"); run<3>("FADD2"); run<4>("FFMA2");
  run<5>("FFMA"); run<12>("FADD"); run<6>("LOP3"); run<7>("F2FP.pack"); run<9>("LDS.64"); run<10>("STS.64");
  run<11>("SHFL");
  run<24>("HMNMX2"); run<28>("IADD"); run<29>("IMAD");
  run<20>("HFMA2+LOP3"); run<21>("HFMA2+FFMA"); run<22>("LOP3+FFMA"); run<23>("HFMA2+HMNMX2"); run<25>("HFMA2+LOP3+LDS");
  run<26>("FADD2+LOP3"); run<27>("FADD2+HFMA2");
  return 0;
}
