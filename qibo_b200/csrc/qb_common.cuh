// qibo_b200: shared host/device helpers (bit-insertion index math, complex arithmetic).
// Everything here is QB_HD so that tests/emul can compile the *same* index math for the CPU.
#pragma once
#include <stdint.h>
#include <math.h>

#if defined(__CUDACC__)
#define QB_HD __host__ __device__ __forceinline__
#define QB_D __device__ __forceinline__
#include <cuda_runtime.h>
#else
#define QB_HD inline
#define QB_D inline
struct double2 { double x, y; };
struct float2 { float x, y; };
#endif

namespace qb {

// ---- complex helpers on float2 / double2 ---------------------------------------------------------
template <typename R> struct cplx_of;
template <> struct cplx_of<float> { using type = float2; };
template <> struct cplx_of<double> { using type = double2; };
template <typename C> struct real_of;
template <> struct real_of<float2> { using type = float; };
template <> struct real_of<double2> { using type = double; };

template <typename C> QB_HD C cmake(typename real_of<C>::type re, typename real_of<C>::type im) {
  C c; c.x = re; c.y = im; return c;
}
template <typename C> QB_HD C cmul(C a, C b) {
  C c; c.x = a.x * b.x - a.y * b.y; c.y = a.x * b.y + a.y * b.x; return c;
}
template <typename C> QB_HD C cadd(C a, C b) { C c; c.x = a.x + b.x; c.y = a.y + b.y; return c; }
// acc += a * b
template <typename C> QB_HD void cfma(C& acc, C a, C b) {
  acc.x += a.x * b.x; acc.x -= a.y * b.y;
  acc.y += a.x * b.y; acc.y += a.y * b.x;
}
template <typename C> QB_HD typename real_of<C>::type cnorm2(C a) { return a.x * a.x + a.y * a.y; }

// explicit fused multiply-add (DFMA / FFMA); std::fma on the host
QB_HD double qfma(double a, double b, double c) { return fma(a, b, c); }
QB_HD float qfma(float a, float b, float c) { return fmaf(a, b, c); }
// v *= ph, written so that the results land in v's own registers (two temporaries, no register moves)
template <typename C> QB_HD void cmul_inplace(C& v, const C ph) {
  const typename real_of<C>::type t = v.y * ph.y, u = v.x * ph.y;
  v.x = qfma(v.x, ph.x, -t);
  v.y = qfma(v.y, ph.x, u);
}

// ---- packed FP32 pairs (Blackwell FFMA2 / FMUL2 / FADD2: one instruction for both components of a complex64) ------
// complex64 sweeps are FP32-issue bound (a 50-gate sweep of the variational ansatz needs ~200 FP32 instructions per
// amplitude against ~91 issue slots per amplitude at HBM speed), so real-coefficient updates use the packed forms.
#if defined(__CUDA_ARCH__)
QB_D float2 f2_fma(float2 a, float2 b, float2 c) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(*reinterpret_cast<unsigned long long*>(&a)), "l"(*reinterpret_cast<unsigned long long*>(&b)),
      "l"(*reinterpret_cast<unsigned long long*>(&c)));
  return *reinterpret_cast<float2*>(&d);
}
QB_D float2 f2_mul(float2 a, float2 b) {
  unsigned long long d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(*reinterpret_cast<unsigned long long*>(&a)), "l"(*reinterpret_cast<unsigned long long*>(&b)));
  return *reinterpret_cast<float2*>(&d);
}
QB_D float2 f2_add(float2 a, float2 b) {
  unsigned long long d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(*reinterpret_cast<unsigned long long*>(&a)), "l"(*reinterpret_cast<unsigned long long*>(&b)));
  return *reinterpret_cast<float2*>(&d);
}
#else
QB_HD float2 f2_fma(float2 a, float2 b, float2 c) { float2 d; d.x = fmaf(a.x, b.x, c.x); d.y = fmaf(a.y, b.y, c.y); return d; }
QB_HD float2 f2_mul(float2 a, float2 b) { float2 d; d.x = a.x * b.x; d.y = a.y * b.y; return d; }
QB_HD float2 f2_add(float2 a, float2 b) { float2 d; d.x = a.x + b.x; d.y = a.y + b.y; return d; }
#endif
// both components of v times / plus real coefficients: packed for complex64, two scalar ops for complex128
QB_HD float2 creal_fma(float k, float2 v, float2 c) { float2 kk; kk.x = k; kk.y = k; return f2_fma(kk, v, c); }
QB_HD float2 creal_mul(float k, float2 v) { float2 kk; kk.x = k; kk.y = k; return f2_mul(kk, v); }
QB_HD float2 creal_add(float2 a, float2 b) { return f2_add(a, b); }
QB_HD double2 creal_fma(double k, double2 v, double2 c) { double2 d; d.x = fma(k, v.x, c.x); d.y = fma(k, v.y, c.y); return d; }
QB_HD double2 creal_mul(double k, double2 v) { double2 d; d.x = k * v.x; d.y = k * v.y; return d; }
QB_HD double2 creal_add(double2 a, double2 b) { double2 d; d.x = a.x + b.x; d.y = a.y + b.y; return d; }

// v = -v when `bit` (0/1) is set: one XOR per component on the sign bits (integer pipe; the FP pipes stay free for the gates)
#if defined(__CUDA_ARCH__)
QB_D void flip_sign(float2& v, uint32_t bit) {
  const uint32_t m = bit << 31;
  v.x = __uint_as_float(__float_as_uint(v.x) ^ m);
  v.y = __uint_as_float(__float_as_uint(v.y) ^ m);
}
QB_D void flip_sign(double2& v, uint32_t bit) {
  const int m = (int)(bit << 31);
  v.x = __hiloint2double(__double2hiint(v.x) ^ m, __double2loint(v.x));
  v.y = __hiloint2double(__double2hiint(v.y) ^ m, __double2loint(v.y));
}
QB_D uint32_t popc32(uint32_t x) { return (uint32_t)__popc(x); }
#else
QB_HD void flip_sign(float2& v, uint32_t bit) { if (bit) { v.x = -v.x; v.y = -v.y; } }
QB_HD void flip_sign(double2& v, uint32_t bit) { if (bit) { v.x = -v.x; v.y = -v.y; } }
QB_HD uint32_t popc32(uint32_t x) { return (uint32_t)__builtin_popcount(x); }
#endif

// ---- bit utilities --------------------------------------------------------------------------------
// insert a zero bit at position p (bits >= p move up by one)
QB_HD uint64_t insert_zero(uint64_t x, int p) {
  uint64_t lo = x & ((uint64_t(1) << p) - 1);
  return ((x >> p) << (p + 1)) | lo;
}
// software pdep: deposit the low bits of x into the set bits of mask (ascending)
QB_HD uint64_t deposit(uint64_t x, uint64_t mask) {
  uint64_t r = 0;
  int k = 0;
  while (mask) {
    uint64_t low = mask & (~mask + 1);
    if ((x >> k) & 1) r |= low;
    mask ^= low;
    ++k;
  }
  return r;
}
// software pext
QB_HD uint64_t extract(uint64_t x, uint64_t mask) {
  uint64_t r = 0;
  int k = 0;
  while (mask) {
    uint64_t low = mask & (~mask + 1);
    if (x & low) r |= uint64_t(1) << k;
    mask ^= low;
    ++k;
  }
  return r;
}

// Sorted list of bit positions to insert zeros at (targets and controls of a gate).
struct InsertList {
  int32_t n;        // number of positions
  uint8_t pos[48];  // ascending
};
QB_HD uint64_t expand(uint64_t g, const InsertList& il) {
  for (int i = 0; i < il.n; ++i) g = insert_zero(g, il.pos[i]);
  return g;
}

}  // namespace qb
