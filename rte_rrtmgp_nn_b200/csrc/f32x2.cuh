// Packed fp32x2 arithmetic for sm_100a: add / mul / fma on a 64-bit register pair are ONE instruction for two lanes
// (FADD2 / FMUL2 / FFMA2), each lane IEEE round-to-nearest like the scalar instruction.  The RTE solvers are bound by
// instruction issue, not by HBM, so every thread carries two g-points in an `f2` and halves its arithmetic
// instruction count; MUFU (ex2 / rcp / rsqrt), comparisons and selects stay per component.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace rrnn {

struct f2 {
  unsigned long long v;
};

__device__ __forceinline__ f2 mk2(float x, float y) {
  f2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(x), "f"(y));
  return r;
}
__device__ __forceinline__ f2 splat2(float x) { return mk2(x, x); }
__device__ __forceinline__ void unpack2(f2 a, float& x, float& y) { asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(a.v)); }
__device__ __forceinline__ float lo2(f2 a) { float x, y; unpack2(a, x, y); return x; }
__device__ __forceinline__ float hi2(f2 a) { float x, y; unpack2(a, x, y); return y; }
__device__ __forceinline__ f2 from_float2(float2 a) { return mk2(a.x, a.y); }

__device__ __forceinline__ f2 operator+(f2 a, f2 b) { f2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
__device__ __forceinline__ f2 operator-(f2 a, f2 b) { f2 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
__device__ __forceinline__ f2 operator*(f2 a, f2 b) { f2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) { f2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v)); return r; }
// -a: written on the components; ptxas folds it into the negate modifier of the consuming FADD2 / FFMA2
__device__ __forceinline__ f2 neg2(f2 a) { float x, y; unpack2(a, x, y); return mk2(-x, -y); }
// c - a*b
__device__ __forceinline__ f2 fnma2(f2 a, f2 b, f2 c) { return fma2(neg2(a), b, c); }
__device__ __forceinline__ float hsum2(f2 a) { float x, y; unpack2(a, x, y); return x + y; }

// componentwise helpers (no packed form exists)
template <typename F>
__device__ __forceinline__ f2 map2(f2 a, F f) { float x, y; unpack2(a, x, y); return mk2(f(x), f(y)); }
__device__ __forceinline__ f2 max2(f2 a, f2 b) { float ax, ay, bx, by; unpack2(a, ax, ay); unpack2(b, bx, by); return mk2(fmaxf(ax, bx), fmaxf(ay, by)); }
__device__ __forceinline__ f2 min2(f2 a, f2 b) { float ax, ay, bx, by; unpack2(a, ax, ay); unpack2(b, bx, by); return mk2(fminf(ax, bx), fminf(ay, by)); }

__device__ __forceinline__ float mufu_ex2(float x) { float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float mufu_rcp(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float mufu_rsqrt(float x) { float r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }

// exp(y) for y <= 0 (any finite y works).  2^(y*log2e) with the rounding error of the product folded back in:
// t = RN(y*L2E), r = y*L2E - t (exact, one FMA), exp(y) = ex2(t) * (1 + r*ln2).  ex2.approx is accurate to ~2^-22.5
// over its whole range (it splits its argument into integer and fraction exactly), which is the error of the result:
// the same 1-2 ulp class as CUDA's expf, in 5 packed instructions + 2 MUFU per pair instead of ~20.
// FAST: the bare ex2(t) (the product's rounding error |y| * 2^-24 stays in the exponent).
template <bool FAST>
__device__ __forceinline__ f2 exp2x(f2 y) {
  const f2 L2E = splat2(1.4426950408889634f);
  const f2 t = y * L2E;
  float tx, ty;
  unpack2(t, tx, ty);
  const f2 e0 = mk2(mufu_ex2(tx), mufu_ex2(ty));
  if (FAST) return e0;
  const f2 r = fma2(y, L2E, mk2(-tx, -ty));  // y*L2E - t
  return fma2(e0, r * splat2(0.6931471805599453f), e0);
}
// 2^t per component, the bare MUFU
__device__ __forceinline__ f2 ex2_raw(f2 t) { float x, y; unpack2(t, x, y); return mk2(mufu_ex2(x), mufu_ex2(y)); }
// 1/x: MUFU seed (+ one Newton step unless FAST): within ~1 ulp, branch-free
template <bool FAST>
__device__ __forceinline__ f2 rcp2(f2 x) {
  float a, b;
  unpack2(x, a, b);
  const f2 r = mk2(mufu_rcp(a), mufu_rcp(b));
  if (FAST) return r;
  return fma2(fnma2(x, r, splat2(1.0f)), r, r);
}
// a/b: q = a*r with the raw MUFU reciprocal (relative error e ~ 2^-23), refined with one residual step: the corrected quotient
// carries e^2 plus the rounding of its own FMA, i.e. it is within ~1 ulp whether or not r was refined first -- so r is not
template <bool FAST>
__device__ __forceinline__ f2 div2(f2 a, f2 b) {
  const f2 r = rcp2<true>(b);
  const f2 q = a * r;
  if (FAST) return q;
  return fma2(fnma2(b, q, a), r, q);
}
// sqrt(x), x > 0: rsqrt seed + one Newton step
template <bool FAST>
__device__ __forceinline__ f2 sqrt2(f2 x) {
  float a, b;
  unpack2(x, a, b);
  const f2 y = mk2(mufu_rsqrt(a), mufu_rsqrt(b));
  const f2 s = x * y;
  if (FAST) return s;
  return fma2(fnma2(s, s, x), y * splat2(0.5f), s);
}

}  // namespace rrnn
