// Arithmetic shared, operation for operation, by the two fused GINE layer kernels
// (gfx_fused8.cu: banded producers; gfx_fused6.cu: producers that walk the CSR arrays), so that
// both produce the same bits for the same row whatever kernel, tile position or batch it is in.
#pragma once
#include <stdint.h>

#include "gfx_common.cuh"
#include "gfx_umma.cuh"

namespace gfx {
namespace lmath {

using namespace ptx;

constexpr int kKbBytes = 128 * 128;           // one K block of a 128-row tile: [128 x 64] fp16

// per-column vectors of the layer as kernel parameters: constant-bank operands, no loads
struct Consts {
  float b1[kMlpHidden];
  float b2[kHidden], g[kHidden], be[kHidden];
};

// ---- packed arithmetic -------------------------------------------------------------------
__device__ __forceinline__ uint32_t relu_pack2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
// fp16x2: relu(x + t), a + b, a * b + c  (each rounded to nearest even, like torch's half ops)
__device__ __forceinline__ uint32_t h2_relu_add(uint32_t x, uint32_t t) {
  uint32_t r;
  asm("fma.rn.relu.f16x2 %0, %1, %2, %3;" : "=r"(r) : "r"(x), "r"(0x3c003c00u), "r"(t));
  return r;
}
__device__ __forceinline__ uint32_t h2_add(uint32_t a, uint32_t b) {
  uint32_t r;
  asm("add.rn.f16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  return r;
}
__device__ __forceinline__ uint32_t h2_fma(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t r;
  asm("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
  return r;
}
// fp32x2 (FADD2 / FFMA2 on sm_100): the two lanes are independent IEEE operations
__device__ __forceinline__ float2 f2_add(float2 a, float2 b) {
  float2 r;
  asm("{\n.reg .b64 ra, rb, rc;\nmov.b64 ra, {%2, %3};\nmov.b64 rb, {%4, %5};\n"
      "add.rn.f32x2 rc, ra, rb;\nmov.b64 {%0, %1}, rc;\n}"
      : "=f"(r.x), "=f"(r.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return r;
}
__device__ __forceinline__ float2 f2_fma(float2 a, float2 b, float2 c) {
  float2 r;
  asm("{\n.reg .b64 ra, rb, rc, rd;\nmov.b64 ra, {%2, %3};\nmov.b64 rb, {%4, %5};\n"
      "mov.b64 rc, {%6, %7};\nfma.rn.f32x2 rd, ra, rb, rc;\nmov.b64 {%0, %1}, rd;\n}"
      : "=f"(r.x), "=f"(r.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
  return r;
}

// ---- shared memory ---------------------------------------------------------------------------
__device__ __forceinline__ uint4 lds128(uint32_t saddr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "r"(saddr));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t saddr, const uint4 &v) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "r"(v.x), "r"(v.y), "r"(v.z),
               "r"(v.w)
               : "memory");
}
__device__ __forceinline__ uint2 lds64(uint32_t saddr) {
  uint2 v;
  asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(saddr));
  return v;
}
__device__ __forceinline__ void sts64(uint32_t saddr, const uint2 &v) {
  asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(saddr), "r"(v.x), "r"(v.y) : "memory");
}
__device__ __forceinline__ void sts32(uint32_t saddr, uint32_t v) {
  asm volatile("st.shared.u32 [%0], %1;" ::"r"(saddr), "r"(v) : "memory");
}
// 8 bytes global -> shared without a register in between (LDGSTS); complete for the issuing
// thread after cp_async_wait_all()
__device__ __forceinline__ void cp_async8(uint32_t saddr, const void *gptr) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(saddr), "l"(gptr) : "memory");
}
__device__ __forceinline__ void cp_async16_plain(uint32_t saddr, const void *gptr) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(saddr), "l"(gptr) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ float2 lds_f2(uint32_t saddr) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(saddr));
  return v;
}
__device__ __forceinline__ void sts_f2(uint32_t saddr, float2 v) {
  asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(saddr), "f"(v.x), "f"(v.y) : "memory");
}
__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}
// 8 bytes of a row: from the resident h tile when the row lies in this tile, else from global
// memory (L2); one predicated instruction of each kind writing the same registers
__device__ __forceinline__ uint2 ld_tile_or_global8(uint32_t in_tile, uint32_t saddr, const uint2 *gptr) {
  uint2 v;
  asm volatile(
      "{\n"
      ".reg .pred q;\n"
      "setp.ne.b32 q, %2, 0;\n"
      "@q ld.shared.v2.u32 {%0, %1}, [%3];\n"
      "@!q ld.global.nc.v2.u32 {%0, %1}, [%4];\n"
      "}\n"
      : "=r"(v.x), "=r"(v.y)
      : "r"(in_tile), "r"(saddr), "l"(gptr));
  return v;
}
// byte offset of 16-byte chunk `c8` (0..7) of row `r` inside one swizzled K block
__device__ __forceinline__ uint32_t sw_off(int r, int c8) {
  return uint32_t(r) * 128u + (uint32_t((c8 ^ r) & 7) << 4);
}

// ---- epilogue B arithmetic (shared, operation for operation, with gfx_fused6.cu) -----------------
// Row r, columns [64 half, 64 half + 64), u = D2 accumulator (fp32):
//   t_c = u_c + b2_c
//   S1 = (sum of t over the even columns, sum over the odd columns), S2 the same for t*t via fma,
//        both in ascending column order;   partial = (S1.x + S1.y, S2.x + S2.y)
//   total = partial[half 0] + partial[half 1];  mean = total.x / 128;
//   var = max(total.y / 128 - mean^2, 0);  rstd = rsqrtf(var + 1e-5);  nm = -mean * rstd
//   y_c = fma(fma(t_c, rstd, nm), g_c, be_c);  out = half(y pair) + h pair   (fp16 add)
// `half` as a compile-time constant (HALF >= 0: every per-column vector element is an immediate
// constant-bank operand) or as a run-time value (HALF < 0: indexed constant loads, small code)
template <int HALF>
__device__ __forceinline__ float2 epi_b_partial_any(const Consts &c, uint32_t tcol, int half_rt) {
  const int half = HALF >= 0 ? HALF : half_rt;
  float2 s1 = make_float2(0.f, 0.f), s2 = make_float2(0.f, 0.f);
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    float u[16];
    tmem_ld16(tcol + 16 * q, u);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int col = half * 64 + 16 * q + 2 * j;
      const float2 t = f2_add(make_float2(u[2 * j], u[2 * j + 1]), make_float2(c.b2[col], c.b2[col + 1]));
      s1 = f2_add(s1, t);
      s2 = f2_fma(t, t, s2);
    }
  }
  return make_float2(s1.x + s1.y, s2.x + s2.y);
}
template <int HALF>
__device__ __forceinline__ float2 epi_b_partial(const Consts &c, uint32_t tcol) {
  return epi_b_partial_any<HALF>(c, tcol, 0);
}
__device__ __forceinline__ float2 epi_b_partial_rt(const Consts &c, uint32_t tcol, int half) {
  return epi_b_partial_any<-1>(c, tcol, half);
}

template <int HALF>
__device__ __forceinline__ void epi_b_normalise_any(const Consts &c, uint32_t tcol, uint32_t hrow, int r,
                                                    float rstd, float nm, int half_rt) {
  const int half = HALF >= 0 ? HALF : half_rt;
  const float2 rs = make_float2(rstd, rstd), nmv = make_float2(nm, nm);
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    float u[16];
    tmem_ld16(tcol + 16 * q, u);
    tmem_ld_wait();
#pragma unroll
    for (int gi = 0; gi < 2; ++gi) {
      const int c16 = half * 8 + q * 2 + gi;                  // 16-byte chunk of the 256-byte row
      const uint32_t cell = hrow + uint32_t(c16 >> 3) * kKbBytes + sw_off(r, c16 & 7);
      const uint4 raw = lds128(cell);
      const uint32_t hin[4] = {raw.x, raw.y, raw.z, raw.w};
      uint32_t o[4];
#pragma unroll
      for (int w = 0; w < 4; ++w) {
        const int j = gi * 8 + 2 * w, col = half * 64 + 16 * q + j;
        const float2 t = f2_add(make_float2(u[j], u[j + 1]), make_float2(c.b2[col], c.b2[col + 1]));
        const float2 y = f2_fma(f2_fma(t, rs, nmv), make_float2(c.g[col], c.g[col + 1]),
                                make_float2(c.be[col], c.be[col + 1]));
        o[w] = h2_add(pack2(y.x, y.y), hin[w]);
      }
      sts128(cell, make_uint4(o[0], o[1], o[2], o[3]));
    }
  }
}

template <int HALF>
__device__ __forceinline__ void epi_b_normalise(const Consts &c, uint32_t tcol, uint32_t hrow, int r,
                                                float rstd, float nm) {
  epi_b_normalise_any<HALF>(c, tcol, hrow, r, rstd, nm, 0);
}
__device__ __forceinline__ void epi_b_normalise_rt(const Consts &c, uint32_t tcol, uint32_t hrow, int r,
                                                   float rstd, float nm, int half) {
  epi_b_normalise_any<-1>(c, tcol, hrow, r, rstd, nm, half);
}

// ---- single pass over TMEM: the thread keeps its 64 columns in registers -------------------------
// (TMEM reads run at ~64 B per cycle and SM: the accumulators of a tile are worth reading once.)
// Same operations in the same order as epi_b_partial / epi_b_normalise above.
template <int HALF>
__device__ __forceinline__ void epi_b_load(const Consts &c, uint32_t tcol, float (&t)[64]) {
  tmem_ld32(tcol, t);
  tmem_ld32(tcol + 32, t + 32);
  tmem_ld_wait();
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    const int col = HALF * 64 + 2 * j;
    const float2 v = f2_add(make_float2(t[2 * j], t[2 * j + 1]), make_float2(c.b2[col], c.b2[col + 1]));
    t[2 * j] = v.x;
    t[2 * j + 1] = v.y;
  }
}
__device__ __forceinline__ float2 epi_b_stats(const float (&t)[64]) {
  float2 s1 = make_float2(0.f, 0.f), s2 = make_float2(0.f, 0.f);
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    const float2 v = make_float2(t[2 * j], t[2 * j + 1]);
    s1 = f2_add(s1, v);
    s2 = f2_fma(v, v, s2);
  }
  return make_float2(s1.x + s1.y, s2.x + s2.y);
}
template <int HALF>
__device__ __forceinline__ void epi_b_store(const Consts &c, const float (&t)[64], uint32_t hrow, int r,
                                            float rstd, float nm) {
  const float2 rs = make_float2(rstd, rstd), nmv = make_float2(nm, nm);
#pragma unroll
  for (int k = 0; k < 8; ++k) {                               // 16-byte chunks of this half row
    const int c16 = HALF * 8 + k;
    const uint32_t cell = hrow + uint32_t(c16 >> 3) * kKbBytes + sw_off(r, c16 & 7);
    const uint4 raw = lds128(cell);
    const uint32_t hin[4] = {raw.x, raw.y, raw.z, raw.w};
    uint32_t o[4];
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      const int j = 8 * k + 2 * w, col = HALF * 64 + j;
      const float2 y = f2_fma(f2_fma(make_float2(t[j], t[j + 1]), rs, nmv),
                              make_float2(c.g[col], c.g[col + 1]), make_float2(c.be[col], c.be[col + 1]));
      o[w] = h2_add(pack2(y.x, y.y), hin[w]);
    }
    sts128(cell, make_uint4(o[0], o[1], o[2], o[3]));
  }
}

}  // namespace lmath
}  // namespace gfx
