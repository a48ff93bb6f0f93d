// qibo_b200 K3 (|amp|^2 + marginals), K4 (CDF scan + inverse-CDF search), K5 (collapse).
#pragma once
#include "qb_common.cuh"
#include "qb_gate_kernels.cuh"

namespace qb {

QB_D double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// deterministic block sum (fixed tree); result valid in thread 0
template <int THREADS> QB_D double block_sum(double v, double* smem /* THREADS/32 */) {
  v = warp_sum(v);
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) smem[w] = v;
  __syncthreads();
  double r = 0.0;
  if (w == 0) {
    r = lane < THREADS / 32 ? smem[lane] : 0.0;
    r = warp_sum(r);
  }
  __syncthreads();
  return r;
}

// ---------------------------------------------------------------------------------------------
// K3.  Replaces calculate_probabilities (abstract.py:2734-2758) + _order_probabilities (:3371-3381).
// A warp owns one value of the measured high bits (`hm`) and a chunk of the unmeasured high bits; its 32
// lanes span the 5 lowest state bits, so every load is one coalesced 256/512-byte row.  Unmeasured low
// bits are folded with warp shuffles; unmeasured high bits are accumulated sequentially (fixed order ->
// deterministic).  When the measured-high space is too small to fill the GPU the unmeasured range is
// split into `nsplit` chunks whose partial marginals are summed, in order, by k3_finish.
// ---------------------------------------------------------------------------------------------
struct ProbParams {
  int n;               // state qubits
  int nlow;            // min(n, 5): bits spanned by the lanes
  uint32_t low_unmeasured;   // mask over the low bits that are NOT measured
  uint64_t mhigh_mask;       // measured bits >= nlow   (state bit positions)
  uint64_t uhigh_mask;       // unmeasured bits >= nlow
  int n_mhigh, n_uhigh;
  int log_split;             // unmeasured-high range split into 2^log_split chunks
  uint64_t nbins;            // 2^m
  int8_t outbit[48];         // state bit position -> output index bit (or -1)
};

QB_D uint64_t out_index(uint64_t idx, const ProbParams& p) {
  uint64_t b = 0;
  for (int pos = 0; pos < p.n; ++pos) {
    int ob = p.outbit[pos];
    if (ob >= 0) b |= ((idx >> pos) & 1ull) << ob;
  }
  return b;
}

template <typename C, typename R>
__global__ void __launch_bounds__(256) k3_probs(const C* __restrict__ state, R* __restrict__ out, double* __restrict__ partial,
                                                const __grid_constant__ ProbParams p) {
  const int lane = threadIdx.x & 31;
  const uint64_t warp = (uint64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const uint64_t ntasks = uint64_t(1) << (p.n_mhigh + p.log_split);
  if (warp >= ntasks) return;
  const uint64_t split = warp & ((uint64_t(1) << p.log_split) - 1);
  const uint64_t hm = warp >> p.log_split;
  const uint64_t base = deposit(hm, p.mhigh_mask);
  const int per_bits = p.n_uhigh - p.log_split;
  const uint64_t per = uint64_t(1) << per_bits;
  const bool valid = lane < (1 << p.nlow);
  double acc = 0.0;
  // walk the unmeasured high bits with a masked increment: d -> ((d | ~mask) + 1) & mask  (no pdep per row)
  const uint64_t um = p.uhigh_mask;
  uint64_t d = deposit(split << per_bits, um);
  uint64_t it = 0;
  for (; it + 4 <= per; it += 4) {
    C v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      v[u] = ld_stream(state + (base | d | lane));
      d = ((d | ~um) + 1) & um;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) acc += (double)cnorm2(v[u]);
  }
  for (; it < per; ++it) {
    if (valid) acc += (double)cnorm2(ld_stream(state + (base | d | lane)));
    d = ((d | ~um) + 1) & um;
  }
  // fold unmeasured low bits
  for (int b = 0; b < 5; ++b)
    if ((p.low_unmeasured >> b) & 1) acc += __shfl_xor_sync(0xffffffffu, acc, 1 << b);
  if (valid && (lane & p.low_unmeasured) == 0) {
    uint64_t bin = out_index(base | lane, p);
    if (p.log_split == 0) out[bin] = (R)acc;
    else partial[split * p.nbins + bin] = acc;
  }
}

// all qubits measured in ascending order: out[i] = |state[i]|^2, two amplitudes per thread and iteration
template <typename C, typename R>
__global__ void __launch_bounds__(256) k3_probs_full(const C* __restrict__ state, R* __restrict__ out, uint64_t count) {
  const uint64_t stride = uint64_t(gridDim.x) * blockDim.x * 2;
  for (uint64_t i = (uint64_t(blockIdx.x) * blockDim.x + threadIdx.x) * 2; i < count; i += stride) {
    const C a = ld_stream(state + i);
    if (i + 1 < count) {
      const C b = ld_stream(state + i + 1);
      out[i] = (R)cnorm2(a);
      out[i + 1] = (R)cnorm2(b);
    } else {
      out[i] = (R)cnorm2(a);
    }
  }
}

template <typename R>
__global__ void __launch_bounds__(256) k3_finish(const double* __restrict__ partial, R* __restrict__ out, uint64_t nbins, int nsplit) {
  uint64_t stride = uint64_t(gridDim.x) * blockDim.x;
  for (uint64_t b = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; b < nbins; b += stride) {
    double s = 0.0;
    for (int k = 0; k < nsplit; ++k) s += partial[uint64_t(k) * nbins + b];
    out[b] = (R)s;
  }
}

// ---------------------------------------------------------------------------------------------
// K4.  CDF + search.  np.random.choice == cumsum (strictly sequential float64) ; cdf /= cdf[-1] ;
// searchsorted(cdf, u, side="right")   (parity hazard 1).
// ---------------------------------------------------------------------------------------------
constexpr int SCAN_CHUNK = 4096;

// mode 0: numpy-exact.  One CTA; chunks are staged through shared memory with coalesced loads, one
// thread carries the running sum in the exact left-to-right order NumPy uses.
template <typename R>
__global__ void __launch_bounds__(256) k4_scan_exact(const R* __restrict__ probs, double* __restrict__ cdf, uint64_t nbins) {
  __shared__ double buf[SCAN_CHUNK];
  __shared__ double carry;
  if (threadIdx.x == 0) carry = 0.0;
  for (uint64_t c0 = 0; c0 < nbins; c0 += SCAN_CHUNK) {
    int len = (int)min((uint64_t)SCAN_CHUNK, nbins - c0);
    for (int i = threadIdx.x; i < len; i += blockDim.x) buf[i] = (double)probs[c0 + i];
    __syncthreads();
    if (threadIdx.x == 0) {
      double run = carry;
      for (int i = 0; i < len; ++i) {
        run += buf[i];
        buf[i] = run;
      }
      carry = run;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < len; i += blockDim.x) cdf[c0 + i] = buf[i];
    __syncthreads();
  }
}

// mode 1: three-phase parallel scan.  Phase A and C run the *same* per-block scan so that the total used
// for normalisation equals the last unnormalised CDF entry bit-for-bit (cdf[-1] / cdf[-1] == 1.0).
constexpr int PSCAN_THREADS = 256;
constexpr int PSCAN_PER_THREAD = 16;
constexpr int PSCAN_BLOCK = PSCAN_THREADS * PSCAN_PER_THREAD;

template <typename R, bool WRITE>
__global__ void __launch_bounds__(PSCAN_THREADS) k4_scan_block(const R* __restrict__ probs, double* __restrict__ cdf,
                                                                double* __restrict__ block_tot,
                                                                const double* __restrict__ block_off, uint64_t nbins) {
  __shared__ double wsum[PSCAN_THREADS / 32];
  const uint64_t b0 = uint64_t(blockIdx.x) * PSCAN_BLOCK;
  const uint64_t t0 = b0 + uint64_t(threadIdx.x) * PSCAN_PER_THREAD;
  double v[PSCAN_PER_THREAD];
  double run = 0.0;
#pragma unroll
  for (int i = 0; i < PSCAN_PER_THREAD; ++i) {
    uint64_t k = t0 + i;
    double x = k < nbins ? (double)probs[k] : 0.0;
    run += x;
    v[i] = run;
  }
  // exclusive scan of the per-thread totals: warp shuffle scan + scan of warp totals
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  double inc = run;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    double t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) wsum[w] = inc;
  __syncthreads();
  double woff = 0.0;
  for (int k = 0; k < w; ++k) woff += wsum[k];
  double prev = __shfl_up_sync(0xffffffffu, inc, 1);
  double excl = woff + (lane == 0 ? 0.0 : prev);
  if (WRITE) {
    double off = block_off[blockIdx.x] + excl;
#pragma unroll
    for (int i = 0; i < PSCAN_PER_THREAD; ++i) {
      uint64_t k = t0 + i;
      if (k < nbins) cdf[k] = off + v[i];
    }
  } else if (threadIdx.x == PSCAN_THREADS - 1) {
    block_tot[blockIdx.x] = excl + run;
  }
}

// exclusive scan of the block totals by one thread (nblocks = nbins / 4096 <= 2^18 at 2^30 bins)
__global__ void k4_scan_totals(const double* __restrict__ block_tot, double* __restrict__ block_off, uint64_t nblocks) {
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    double run = 0.0;
    for (uint64_t i = 0; i < nblocks; ++i) {
      block_off[i] = run;
      run += block_tot[i];
    }
  }
}

__global__ void __launch_bounds__(256) k4_normalize(double* __restrict__ cdf, uint64_t nbins) {
  const double last = cdf[nbins - 1];
  __syncthreads();
  uint64_t stride = uint64_t(gridDim.x) * blockDim.x;
  // the last element is written by the thread that owns it, after everybody has read it: two kernels
  for (uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; i + 1 < nbins; i += stride) cdf[i] = cdf[i] / last;
}
__global__ void k4_normalize_last(double* __restrict__ cdf, uint64_t nbins) {
  if (blockIdx.x == 0 && threadIdx.x == 0) cdf[nbins - 1] = cdf[nbins - 1] / cdf[nbins - 1];
}

// idx = #{k : cdf[k] <= u}  (searchsorted side="right")
__global__ void __launch_bounds__(256) k4_search(const double* __restrict__ cdf, uint64_t nbins, const double* __restrict__ u,
                                                 uint64_t nshots, long long* __restrict__ out) {
  uint64_t stride = uint64_t(gridDim.x) * blockDim.x;
  for (uint64_t s = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; s < nshots; s += stride) {
    const double x = u[s];
    uint64_t lo = 0, hi = nbins;  // first index with cdf[idx] > x
    while (lo < hi) {
      uint64_t mid = lo + ((hi - lo) >> 1);
      if (__ldg(cdf + mid) <= x) lo = mid + 1;
      else hi = mid;
    }
    out[s] = (long long)lo;
  }
}

// ---------------------------------------------------------------------------------------------
// K5.  collapse_state (abstract.py:2424-2440 -> _collapse_statevector :3279-3304): keep the slice where the
// measured bits equal the outcome, divide it by sqrt(sum |.|^2), zero the rest.  Also the plain norm.
// ---------------------------------------------------------------------------------------------
constexpr int RED_THREADS = 256;

template <typename C>
__global__ void __launch_bounds__(RED_THREADS) k5_slice_norm2(const C* __restrict__ state, uint64_t ngroups, InsertList ins,
                                                              uint64_t val, double* __restrict__ partial) {
  __shared__ double sm[RED_THREADS / 32];
  double acc = 0.0;
  uint64_t stride = uint64_t(gridDim.x) * RED_THREADS;
  for (uint64_t g = uint64_t(blockIdx.x) * RED_THREADS + threadIdx.x; g < ngroups; g += stride)
    acc += (double)cnorm2(ld_stream(state + (expand(g, ins) | val)));
  double r = block_sum<RED_THREADS>(acc, sm);
  if (threadIdx.x == 0) partial[blockIdx.x] = r;
}

__global__ void k5_sum_partials(const double* __restrict__ partial, int n, double* __restrict__ out) {
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    double s = 0.0;
    for (int i = 0; i < n; ++i) s += partial[i];
    out[0] = s;
  }
}

template <typename C>
__global__ void __launch_bounds__(256) k5_project(C* __restrict__ state, uint64_t count, uint64_t mask, uint64_t val,
                                                  const double* __restrict__ norm2, int normalize) {
  typedef typename real_of<C>::type R;
  const R nrm = normalize ? (R)sqrt(norm2[0]) : (R)1;
  uint64_t stride = uint64_t(gridDim.x) * blockDim.x;
  for (uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < count; i += stride) {
    if ((i & mask) == val) {
      if (normalize) {
        C v = ld_stream(state + i);
        v.x = v.x / nrm;
        v.y = v.y / nrm;
        st_stream(state + i, v);
      }
    } else {
      st_stream(state + i, cmake<C>(0, 0));
    }
  }
}

// ---- X1: density matrices (rho as a flat 2^n x 2^n row-major array) --------------------------------------------------
// calculate_probabilities(density_matrix=True) (abstract.py:2741-2749): abs(sum of the diagonal over the unmeasured
// qubits), bins in the caller's qubit order.  One warp per bin; density matrices are small (n <= ~15).
struct DmProbParams {
  int n, m;
  uint64_t umask;       // unmeasured bit positions of the row index
  int n_u;
  uint8_t pos[48];      // bit position of measured qubit i (caller order: qubit 0 of the list = MSB of the bin)
};
template <typename C, typename R>
__global__ void __launch_bounds__(256) k3_probs_dm(const C* __restrict__ rho, R* __restrict__ out, const __grid_constant__ DmProbParams p) {
  const int lane = threadIdx.x & 31;
  const uint64_t bin = (uint64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  if (bin >> p.m) return;
  uint64_t base = 0;
  for (int i = 0; i < p.m; ++i) base |= ((bin >> (p.m - 1 - i)) & 1ull) << p.pos[i];
  const uint64_t diag = (uint64_t(1) << p.n) + 1;
  double re = 0.0, im = 0.0;
  for (uint64_t u = lane; u < (uint64_t(1) << p.n_u); u += 32) {
    const C v = rho[(base | deposit(u, p.umask)) * diag];
    re += (double)v.x;
    im += (double)v.y;
  }
  re = warp_sum(re);
  im = warp_sum(im);
  if (lane == 0) out[bin] = (R)hypot(re, im);
}

// _collapse_density_matrix (abstract.py:3249-3277): keep the block whose row AND column bits equal the outcome, divide by
// its (complex) trace, zero everything else.
template <typename C>
__global__ void __launch_bounds__(RED_THREADS) k5_diag_slice_sum(const C* __restrict__ rho, int n, uint64_t ngroups, InsertList ins, uint64_t val,
                                                                 double* __restrict__ partial_re, double* __restrict__ partial_im) {
  __shared__ double sm[RED_THREADS / 32];
  double re = 0.0, im = 0.0;
  const uint64_t diag = (uint64_t(1) << n) + 1;
  const uint64_t stride = uint64_t(gridDim.x) * RED_THREADS;
  for (uint64_t g = uint64_t(blockIdx.x) * RED_THREADS + threadIdx.x; g < ngroups; g += stride) {
    const C v = rho[(expand(g, ins) | val) * diag];
    re += (double)v.x;
    im += (double)v.y;
  }
  const double r = block_sum<RED_THREADS>(re, sm);
  const double i = block_sum<RED_THREADS>(im, sm);
  if (threadIdx.x == 0) {
    partial_re[blockIdx.x] = r;
    partial_im[blockIdx.x] = i;
  }
}
template <typename C>
__global__ void __launch_bounds__(256) k5_project_dm(C* __restrict__ rho, uint64_t count, uint64_t mask, uint64_t val, const double* __restrict__ trace,
                                                     int normalize) {
  typedef typename real_of<C>::type R;
  const double tr = trace[0], ti = trace[1];
  const double den = tr * tr + ti * ti;
  uint64_t stride = uint64_t(gridDim.x) * blockDim.x;
  for (uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < count; i += stride) {
    if ((i & mask) == val) {
      if (normalize) {
        const C v = rho[i];
        const double x = (double)v.x, y = (double)v.y;
        rho[i] = cmake<C>((R)((x * tr + y * ti) / den), (R)((y * tr - x * ti) / den));
      }
    } else {
      rho[i] = cmake<C>(0, 0);
    }
  }
}

// ---- K9: <psi| P |psi> for a Pauli string, and <a|b> ------------------------------------------------------------------
// P|x> = i^{nY} (-1)^{popcount(x & zmask)} |x ^ xmask>  (xmask: bits with X or Y, zmask: bits with Z or Y), so
// <psi|P|psi> = i^{nY} sum_x (-1)^{popcount(x & zmask)} conj(psi[x ^ xmask]) psi[x]: one read pass over the state (the
// partner read hits the same lines when xmask only flips low bits, another resident line otherwise), no copy of the
// state and no gate application.  Deterministic two-stage sums (block partials in a fixed order).
template <typename C>
__global__ void __launch_bounds__(RED_THREADS) k9_pauli_expval(const C* __restrict__ state, uint64_t count, uint64_t xmask, uint64_t zmask,
                                                               double* __restrict__ partial_re, double* __restrict__ partial_im) {
  __shared__ double sm[RED_THREADS / 32];
  double re = 0.0, im = 0.0;
  const uint64_t stride = uint64_t(gridDim.x) * RED_THREADS;
  for (uint64_t x = uint64_t(blockIdx.x) * RED_THREADS + threadIdx.x; x < count; x += stride) {
    const C a = ld_stream(state + x);
    const C b = xmask ? __ldg(state + (x ^ xmask)) : a;
    // conj(b) * a
    double pr = (double)b.x * (double)a.x + (double)b.y * (double)a.y;
    double pi = (double)b.x * (double)a.y - (double)b.y * (double)a.x;
    if (__popcll(x & zmask) & 1) {
      pr = -pr;
      pi = -pi;
    }
    re += pr;
    im += pi;
  }
  const double r = block_sum<RED_THREADS>(re, sm);
  __syncthreads();
  const double i = block_sum<RED_THREADS>(im, sm);
  if (threadIdx.x == 0) {
    partial_re[blockIdx.x] = r;
    partial_im[blockIdx.x] = i;
  }
}

template <typename C>
__global__ void __launch_bounds__(RED_THREADS) k9_vdot(const C* __restrict__ a, const C* __restrict__ b, uint64_t count,
                                                       double* __restrict__ partial_re, double* __restrict__ partial_im) {
  __shared__ double sm[RED_THREADS / 32];
  double re = 0.0, im = 0.0;
  const uint64_t stride = uint64_t(gridDim.x) * RED_THREADS;
  for (uint64_t x = uint64_t(blockIdx.x) * RED_THREADS + threadIdx.x; x < count; x += stride) {
    const C u = ld_stream(a + x), v = ld_stream(b + x);
    re += (double)u.x * (double)v.x + (double)u.y * (double)v.y;  // conj(u) * v
    im += (double)u.x * (double)v.y - (double)u.y * (double)v.x;
  }
  const double r = block_sum<RED_THREADS>(re, sm);
  __syncthreads();
  const double i = block_sum<RED_THREADS>(im, sm);
  if (threadIdx.x == 0) {
    partial_re[blockIdx.x] = r;
    partial_im[blockIdx.x] = i;
  }
}

__global__ void k9_sum_partials2(const double* __restrict__ pre, const double* __restrict__ pim, int n, double* __restrict__ out) {
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    double s = 0.0, t = 0.0;
    for (int i = 0; i < n; ++i) {
      s += pre[i];
      t += pim[i];
    }
    out[0] = s;
    out[1] = t;
  }
}

}  // namespace qb
