// qibo_b200 K1: one gate per HBM sweep, bit-insertion indexing, in place.
//
// Replaces Backend.apply_gate / _apply_gate_controlled_by (abstract.py:2322-2361, 3176-3197): instead of
// transpose -> matmul -> inverse transpose (four full-state copies) each thread owns one group of 2^K
// amplitudes -- the ones that differ only in the K target bits, with every control bit set -- loads them
// with 128-bit accesses, multiplies by the 2^K x 2^K matrix held in the kernel-parameter constant bank
// (warp-uniform broadcast reads) and stores them back.  Algorithmic traffic: 2 * B * 2^(n - c).
#pragma once
#include "qb_common.cuh"

namespace qb {

constexpr int K1_THREADS = 256;

// streaming 128-bit / 64-bit accesses: every amplitude is touched once per sweep, keep L1 out of it
template <typename C> QB_D C ld_stream(const C* p);
template <> QB_D double2 ld_stream<double2>(const double2* p) {
  double2 v;
  asm volatile("ld.global.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
  return v;
}
template <> QB_D float2 ld_stream<float2>(const float2* p) {
  float2 v;
  asm volatile("ld.global.L1::no_allocate.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
  return v;
}
template <typename C> QB_D void st_stream(C* p, C v);
template <> QB_D void st_stream<double2>(double2* p, double2 v) {
  asm volatile("st.global.L1::no_allocate.v2.f64 [%0], {%1,%2};" ::"l"(p), "d"(v.x), "d"(v.y) : "memory");
}
template <> QB_D void st_stream<float2>(float2* p, float2 v) {
  asm volatile("st.global.L1::no_allocate.v2.f32 [%0], {%1,%2};" ::"l"(p), "f"(v.x), "f"(v.y) : "memory");
}

template <typename C, int K> struct DenseParams {
  uint64_t ngroups;          // 2^(n - K - c)
  uint64_t cmask;            // control bits, all set in every touched index
  uint64_t off[1 << K];      // offset of matrix index j (targets[0] = MSB)
  InsertList ins;            // sorted target+control positions
  C m[(1 << K) * (1 << K)];  // row-major, already cast to the state's precision
};

// One group per thread, UNROLL groups in flight per thread.
template <typename C, int K, int UNROLL>
__global__ void __launch_bounds__(K1_THREADS) k1_dense(C* __restrict__ state, const __grid_constant__ DenseParams<C, K> p) {
  constexpr int D = 1 << K;
  uint64_t g0 = (uint64_t(blockIdx.x) * K1_THREADS) * UNROLL + threadIdx.x;
  C v[UNROLL][D];
  uint64_t base[UNROLL];
#pragma unroll
  for (int u = 0; u < UNROLL; ++u) {
    uint64_t g = g0 + uint64_t(u) * K1_THREADS;
    base[u] = expand(g, p.ins) | p.cmask;
    if (g < p.ngroups) {
#pragma unroll
      for (int j = 0; j < D; ++j) v[u][j] = ld_stream(state + (base[u] | p.off[j]));
    }
  }
#pragma unroll
  for (int u = 0; u < UNROLL; ++u) {
    uint64_t g = g0 + uint64_t(u) * K1_THREADS;
    if (g < p.ngroups) {
#pragma unroll
      for (int i = 0; i < D; ++i) {
        C acc = cmul(p.m[i * D], v[u][0]);
#pragma unroll
        for (int j = 1; j < D; ++j) cfma(acc, p.m[i * D + j], v[u][j]);
        st_stream(state + (base[u] | p.off[i]), acc);
      }
    }
  }
}

template <typename C, int K> struct DiagParams {
  uint64_t ngroups;
  uint64_t cmask;
  uint64_t off[1 << K];
  InsertList ins;
  C d[1 << K];
};

template <typename C, int K, int UNROLL>
__global__ void __launch_bounds__(K1_THREADS) k1_diag(C* __restrict__ state, const __grid_constant__ DiagParams<C, K> p) {
  constexpr int D = 1 << K;
  uint64_t g0 = (uint64_t(blockIdx.x) * K1_THREADS) * UNROLL + threadIdx.x;
  C v[UNROLL][D];
  uint64_t base[UNROLL];
#pragma unroll
  for (int u = 0; u < UNROLL; ++u) {
    uint64_t g = g0 + uint64_t(u) * K1_THREADS;
    base[u] = expand(g, p.ins) | p.cmask;
    if (g < p.ngroups) {
#pragma unroll
      for (int j = 0; j < D; ++j) v[u][j] = ld_stream(state + (base[u] | p.off[j]));
    }
  }
#pragma unroll
  for (int u = 0; u < UNROLL; ++u) {
    uint64_t g = g0 + uint64_t(u) * K1_THREADS;
    if (g < p.ngroups) {
#pragma unroll
      for (int j = 0; j < D; ++j) st_stream(state + (base[u] | p.off[j]), cmul(p.d[j], v[u][j]));
    }
  }
}

// PHASE: scalar on the all-controls-set slice (K = 0).  SWAP: exchange |01> and |10>.
template <typename C> struct SliceParams {
  uint64_t ngroups;
  uint64_t cmask;
  uint64_t off01, off10;  // SWAP only
  InsertList ins;
  C phase;
};

template <typename C, int UNROLL>
__global__ void __launch_bounds__(K1_THREADS) k1_phase(C* __restrict__ state, const __grid_constant__ SliceParams<C> p) {
  uint64_t g0 = (uint64_t(blockIdx.x) * K1_THREADS) * UNROLL + threadIdx.x;
  C v[UNROLL];
  uint64_t idx[UNROLL];
#pragma unroll
  for (int u = 0; u < UNROLL; ++u) {
    uint64_t g = g0 + uint64_t(u) * K1_THREADS;
    idx[u] = expand(g, p.ins) | p.cmask;
    if (g < p.ngroups) v[u] = ld_stream(state + idx[u]);
  }
#pragma unroll
  for (int u = 0; u < UNROLL; ++u) {
    uint64_t g = g0 + uint64_t(u) * K1_THREADS;
    if (g < p.ngroups) st_stream(state + idx[u], cmul(p.phase, v[u]));
  }
}

template <typename C, int UNROLL>
__global__ void __launch_bounds__(K1_THREADS) k1_swap(C* __restrict__ state, const __grid_constant__ SliceParams<C> p) {
  uint64_t g0 = (uint64_t(blockIdx.x) * K1_THREADS) * UNROLL + threadIdx.x;
  C a[UNROLL], b[UNROLL];
  uint64_t base[UNROLL];
#pragma unroll
  for (int u = 0; u < UNROLL; ++u) {
    uint64_t g = g0 + uint64_t(u) * K1_THREADS;
    base[u] = expand(g, p.ins) | p.cmask;
    if (g < p.ngroups) {
      a[u] = ld_stream(state + (base[u] | p.off01));
      b[u] = ld_stream(state + (base[u] | p.off10));
    }
  }
#pragma unroll
  for (int u = 0; u < UNROLL; ++u) {
    uint64_t g = g0 + uint64_t(u) * K1_THREADS;
    if (g < p.ngroups) {
      st_stream(state + (base[u] | p.off01), b[u]);
      st_stream(state + (base[u] | p.off10), a[u]);
    }
  }
}

// ---- K6: state construction / conversion -------------------------------------------------------
template <typename C>
__global__ void __launch_bounds__(256) k6_fill(C* __restrict__ state, uint64_t count, C value, uint64_t one_index, int set_one) {
  uint64_t stride = uint64_t(gridDim.x) * blockDim.x;
  for (uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < count; i += stride) {
    C v = value;
    if (set_one && i == one_index) v = cmake<C>(1, 0);
    st_stream(state + i, v);
  }
}

template <typename CD, typename CS>
__global__ void __launch_bounds__(256) k6_cast(CD* __restrict__ dst, const CS* __restrict__ src, uint64_t count) {
  uint64_t stride = uint64_t(gridDim.x) * blockDim.x;
  for (uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < count; i += stride) {
    CS s = ld_stream(src + i);
    CD d;
    d.x = (typename real_of<CD>::type)s.x;
    d.y = (typename real_of<CD>::type)s.y;
    st_stream(dst + i, d);
  }
}

}  // namespace qb
