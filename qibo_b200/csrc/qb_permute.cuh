// qibo_b200 K8: out-of-place qubit permutation in ONE HBM sweep (runs of SWAP gates, e.g. the bit reversal
// that ends a QFT: models/qft.py:55-57).
//
// dst[P(i)] = src[i], where bit pi[b] of P(i) is bit b of i.  A tile is spanned (in source index space) by
// the LA lowest source bits -- contiguous, coalesced reads -- and by the source bits that land on the LB
// lowest destination bits -- contiguous, coalesced writes; the transposition happens in shared memory.  Any
// permutation therefore costs 2 * B * 2^n bytes, like one gate sweep, instead of one sweep per ~3 swaps.
#pragma once
#include "qb_common.cuh"

namespace qb {

constexpr int PERM_THREADS = 256;
constexpr int PERM_MAX_TILE_BITS = 12;
constexpr int PERM_MAX_ITERS = (1 << PERM_MAX_TILE_BITS) / PERM_THREADS;
// one element of padding per 64: the store phase reads the tile transposed (a bit reversal walks it with a stride of
// 2^6 elements = every lane on the same banks)
QB_HD uint32_t perm_pad(uint32_t e) { return e + (e >> 6); }

struct PermParams {
  int n;               // qubits
  int tbits;           // tile bits |S|
  int lb;              // destination-contiguous bits (low tbits of the dst-ordered tile index f that are dst bits 0..lb-1)
  int la;              // source-contiguous bits
  uint64_t smask;      // S: source bits of the tile (contains bits 0..la-1)
  uint64_t dmask;      // pi(S): destination bits of the tile (contains bits 0..lb-1)
  uint64_t ntiles;
  uint8_t pi[48];      // destination bit of source bit b
  uint8_t emap[16];    // bit k of the dst-ordered tile index f  ->  bit of the src-ordered tile index e
};

// destination index of a source index whose tile bits are zero
QB_HD uint64_t permute_base(uint64_t src_base, const PermParams& p) {
  uint64_t d = 0;
  for (int b = 0; b < p.n; ++b)
    if ((src_base >> b) & 1) d |= uint64_t(1) << p.pi[b];
  return d;
}
QB_HD uint32_t perm_e_of_f(uint32_t f, const PermParams& p) {
  uint32_t e = 0;
  for (int k = 0; k < p.tbits; ++k)
    if ((f >> k) & 1) e |= 1u << p.emap[k];
  return e;
}

// host: fill the tile description for permutation pi (destination bit of every source bit)
inline void perm_setup(int n, int la_cfg, int lb_cfg, const int* pi, PermParams& p) {
  p.n = n;
  for (int b = 0; b < n; ++b) p.pi[b] = (uint8_t)pi[b];
  int la = la_cfg < n ? la_cfg : n, lb = lb_cfg < n ? lb_cfg : n;
  uint64_t s = (uint64_t(1) << la) - 1;
  for (int b = 0; b < n; ++b)
    if (pi[b] < lb) s |= uint64_t(1) << b;
  p.smask = s;
  p.dmask = 0;
  for (int b = 0; b < n; ++b)
    if ((s >> b) & 1) p.dmask |= uint64_t(1) << pi[b];
  p.tbits = __builtin_popcountll(s);
  p.la = la;
  p.lb = lb;
  p.ntiles = uint64_t(1) << (n - p.tbits);
  // dst-ordered tile index f: bit k <-> k-th lowest bit of dmask; src-ordered e: bit r <-> r-th lowest bit of smask
  int k = 0;
  for (int d = 0; d < n; ++d) {
    if (!((p.dmask >> d) & 1)) continue;
    int srcbit = 0;
    for (int b = 0; b < n; ++b)
      if (pi[b] == d) srcbit = b;
    int r = __builtin_popcountll(s & ((uint64_t(1) << srcbit) - 1));
    p.emap[k++] = (uint8_t)r;
  }
}

// split a mask into its 8 lowest set bits (indexed by the thread id) and the rest (walked per iteration)
QB_HD void split_mask8(uint64_t mask, uint64_t& lo, uint64_t& hi) {
  lo = 0;
  uint64_t m = mask;
  for (int i = 0; i < 8 && m; ++i) {
    uint64_t low = m & (~m + 1);
    lo |= low;
    m ^= low;
  }
  hi = m;
}

#if defined(__CUDACC__)
template <typename C>
__global__ void __launch_bounds__(PERM_THREADS) k8_permute(const C* __restrict__ src, C* __restrict__ dst, const __grid_constant__ PermParams p) {
  extern __shared__ __align__(16) unsigned char perm_smem[];
  C* tile = reinterpret_cast<C*>(perm_smem);
  __shared__ uint32_t e_of_k[PERM_MAX_ITERS];  // source-ordered tile index contributed by the iteration counter (store phase)
  const uint32_t tsize = 1u << p.tbits;
  const uint32_t iters = tsize > PERM_THREADS ? tsize / PERM_THREADS : 1;
  if (threadIdx.x < iters) e_of_k[threadIdx.x] = perm_e_of_f((threadIdx.x * PERM_THREADS) & (tsize - 1), p);
  __syncthreads();
  const uint64_t other = ~p.smask & ((uint64_t(1) << p.n) - 1);
  // per-thread constants: the thread id supplies the 8 low bits of the src-ordered index e (load phase) and of
  // the dst-ordered index f (store phase); the remaining tile bits advance with a masked increment
  uint64_t s_lo, s_hi, d_lo, d_hi;
  split_mask8(p.smask, s_lo, s_hi);
  split_mask8(p.dmask, d_lo, d_hi);
  const uint64_t soff_lo = deposit(threadIdx.x, s_lo), doff_lo = deposit(threadIdx.x, d_lo);
  const uint32_t e_lo_of_f = perm_e_of_f(threadIdx.x & (tsize - 1), p);
  const bool active = threadIdx.x < tsize;
  for (uint64_t t = blockIdx.x; t < p.ntiles; t += gridDim.x) {
    const uint64_t sbase = deposit(t, other);
    const uint64_t dbase = permute_base(sbase, p);
    if (active) {
      // batches of 8 independent loads per thread: with one load in flight per thread (the rolled loop) a CTA keeps
      // only 4 KiB in flight and the sweep ran at half of the HBM rate
      uint64_t hi = 0;
      uint32_t k = 0;
      for (; k + 8 <= iters; k += 8) {
        C v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          v[u] = ld_stream(src + (sbase | soff_lo | hi));
          hi = ((hi | ~s_hi) + 1) & s_hi;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) tile[perm_pad(threadIdx.x + (k + u) * PERM_THREADS)] = v[u];
      }
      for (; k < iters; ++k) {
        tile[perm_pad(threadIdx.x + k * PERM_THREADS)] = ld_stream(src + (sbase | soff_lo | hi));
        hi = ((hi | ~s_hi) + 1) & s_hi;
      }
    }
    __syncthreads();
    if (active) {
      uint64_t hi = 0;
      for (uint32_t k = 0; k < iters; ++k) {
        const uint32_t e = e_lo_of_f | e_of_k[k];
        st_stream(dst + (dbase | doff_lo | hi), tile[perm_pad(e)]);
        hi = ((hi | ~d_hi) + 1) & d_hi;
      }
    }
    __syncthreads();
  }
}

inline int launch_permute(cudaStream_t stream, int sm_count, const void* src, void* dst, int dtype, const PermParams& p) {
  const size_t esize = dtype == QB_C128 ? 16 : 8;
  const size_t smem = esize * (size_t)perm_pad(1u << p.tbits);
  // CTAs per SM: as many tiles as fit the shared memory (3 x 65 KiB complex128 tiles; more when the tile is smaller)
  int per_sm = (int)((size_t)220 * 1024 / (smem + 1024));
  if (per_sm < 1) per_sm = 1;
  if (per_sm > 8) per_sm = 8;
  per_sm = env_int("QB_PERM_CTAS_PER_SM", per_sm);
  uint64_t cap = (uint64_t)sm_count * per_sm;
  unsigned grid = (unsigned)(p.ntiles < cap ? p.ntiles : cap);
  if (dtype == QB_C128) {
    cudaFuncSetAttribute(k8_permute<double2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k8_permute<double2><<<grid, PERM_THREADS, smem, stream>>>((const double2*)src, (double2*)dst, p);
  } else {
    cudaFuncSetAttribute(k8_permute<float2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k8_permute<float2><<<grid, PERM_THREADS, smem, stream>>>((const float2*)src, (float2*)dst, p);
  }
  return cudaPeekAtLastError() == cudaSuccess ? QB_OK : QB_ERR_CUDA;
}
#endif

}  // namespace qb
