// qibo_b200 K2: shared-memory passes of the sweep kernel.  Pure index math + complex arithmetic, QB_HD so
// that tests/emul compiles the very same code for the CPU (nct = 1) to check the planner and the passes
// without a GPU.  The CUDA kernel that calls them (with nct = 512 compute threads) is in qb_sweep.cuh.
#pragma once
#include "qb_common.cuh"
#include "qb_planner.hpp"

namespace qb {

QB_HD uint32_t insert_zero32(uint32_t x, int p) {
  uint32_t lo = x & ((1u << p) - 1);
  return ((x >> p) << (p + 1)) | lo;
}
// index of group g after inserting zeros at the tile-local bits of `mask`
QB_HD uint32_t expand_mask(uint32_t g, uint32_t mask) {
  while (mask) {
    uint32_t low = mask & (~mask + 1);
    uint32_t lo = g & (low - 1);
    g = ((g & ~(low - 1)) << 1) | lo;
    mask ^= low;
  }
  return g;
}

// ---- shared-memory passes (QB_HD so tests/emul can run them on the CPU) -------------------------------
// `ctid`/`nct`: index / number of cooperating threads.  The emulator calls them with nct = 1.
template <typename C, int K>
QB_HD void pass_dense(C* __restrict__ tile, const DevOp& op, const C* __restrict__ m, int T, uint32_t ctid, uint32_t nct) {
  constexpr int D = 1 << K;
  uint32_t off[D];
#pragma unroll
  for (int j = 0; j < D; ++j) {
    uint32_t o = 0;
#pragma unroll
    for (int i = 0; i < K; ++i)
      if ((j >> (K - 1 - i)) & 1) o |= 1u << op.tbit[i];
    off[j] = o;
  }
  const uint32_t ngroups = 1u << (T - (int)op.nins);
  const uint32_t ins_mask = op.ins_mask, cmask = op.tl_cmask;
  for (uint32_t g = ctid; g < ngroups; g += nct) {
    const uint32_t t0 = expand_mask(g, ins_mask) | cmask;
    C v[D];
#pragma unroll
    for (int j = 0; j < D; ++j) v[j] = tile[t0 | off[j]];
#pragma unroll
    for (int i = 0; i < D; ++i) {
      C acc = cmul(m[i * D], v[0]);
#pragma unroll
      for (int j = 1; j < D; ++j) cfma(acc, m[i * D + j], v[j]);
      tile[t0 | off[i]] = acc;
    }
  }
}

template <typename C> QB_HD void pass_swap(C* tile, const DevOp& op, int T, uint32_t ctid, uint32_t nct) {
  const uint32_t o01 = 1u << op.tbit[1], o10 = 1u << op.tbit[0];
  const uint32_t ngroups = 1u << (T - (int)op.nins);
  const uint32_t ins_mask = op.ins_mask, cmask = op.tl_cmask;
  for (uint32_t g = ctid; g < ngroups; g += nct) {
    const uint32_t t0 = expand_mask(g, ins_mask) | cmask;
    C a = tile[t0 | o01], b = tile[t0 | o10];
    tile[t0 | o01] = b;
    tile[t0 | o10] = a;
  }
}

template <typename C>
QB_HD void pass_fan(C* tile, const DevOp& op, const char* blob, C scal, int T, uint32_t ctid, uint32_t nct) {
  const C* tab0 = (const C*)(blob + op.payload);
  const C* tab1 = tab0 + (1u << op.chunk_len[0]);
  const uint32_t lo0 = op.chunk_lo[0], mk0 = (1u << op.chunk_len[0]) - 1;
  const uint32_t lo1 = op.chunk_lo[1], mk1 = (1u << op.chunk_len[1]) - 1;
  const uint32_t cm = op.tl_cmask;
#if defined(__CUDA_ARCH__)
  const int nc = __popc(cm);
#else
  const int nc = __builtin_popcount(cm);
#endif
  const uint32_t ngroups = 1u << (T - nc);
  const int nch = (int)op.n_chunks;
  for (uint32_t g = ctid; g < ngroups; g += nct) {
    const uint32_t t = expand_mask(g, cm) | cm;
    C f = scal;
    if (nch > 0) f = cmul(f, tab0[(t >> lo0) & mk0]);
    if (nch > 1) f = cmul(f, tab1[(t >> lo1) & mk1]);
    tile[t] = cmul(tile[t], f);
  }
}

template <typename C>
QB_HD void pass_diagk(C* tile, const DevOp& op, const char* blob, uint32_t aux, int T, uint32_t ctid, uint32_t nct) {
  const C* tab = (const C*)(blob + op.payload);
  const uint32_t cm = op.tl_cmask;
#if defined(__CUDA_ARCH__)
  const int nc = __popc(cm);
#else
  const int nc = __builtin_popcount(cm);
#endif
  const int k = (int)op.k;
  const uint32_t ngroups = 1u << (T - nc);
  for (uint32_t g = ctid; g < ngroups; g += nct) {
    const uint32_t t = expand_mask(g, cm) | cm;
    uint32_t idx = aux;
    for (int i = 0; i < k; ++i)
      if (op.tbit[i] != 0xFF) idx |= ((t >> op.tbit[i]) & 1u) << (k - 1 - i);
    tile[t] = cmul(tile[t], tab[idx]);
  }
}

// k = 5, 6: 2^(k-3) threads share one group, 8 output rows each; inputs are re-read from shared memory.
// Split in a read/accumulate half and a write half with a barrier between (the caller provides it).
template <typename C> struct BigAcc { C acc[8]; uint32_t t0; bool valid; uint32_t sub; };

QB_HD uint32_t big_offset(const DevOp& op, int k, int j) {
  uint32_t o = 0;
  for (int i = 0; i < k; ++i)
    if ((j >> (k - 1 - i)) & 1) o |= 1u << op.tbit[i];
  return o;
}

template <typename C>
QB_HD void big_read(const C* tile, const DevOp& op, const C* m, int T, uint32_t task, BigAcc<C>& a) {
  const int k = (int)op.k, D = 1 << k, lt = k - 3;
  const uint32_t ntasks = (1u << (T - (int)op.nins)) << lt;
  a.valid = task < ntasks;
  if (!a.valid) return;
  a.sub = task & ((1u << lt) - 1);
  a.t0 = expand_mask(task >> lt, op.ins_mask) | op.tl_cmask;
#pragma unroll
  for (int r = 0; r < 8; ++r) a.acc[r] = cmake<C>(0, 0);
  for (int j = 0; j < D; ++j) {
    const C x = tile[a.t0 | big_offset(op, k, j)];
#pragma unroll
    for (int r = 0; r < 8; ++r) cfma(a.acc[r], m[(a.sub * 8 + r) * D + j], x);
  }
}
template <typename C> QB_HD void big_write(C* tile, const DevOp& op, const BigAcc<C>& a) {
  if (!a.valid) return;
  const int k = (int)op.k;
#pragma unroll
  for (int r = 0; r < 8; ++r) tile[a.t0 | big_offset(op, k, (int)(a.sub * 8 + r))] = a.acc[r];
}

// per-tile, per-op set-up: is the op active on this tile (controls outside the tile), the fan's factor from
// the bits outside the tile, the DIAGK table-index part from the bits outside the tile
template <typename C>
QB_HD void op_prephase(const DevOp& op, const char* blob, uint64_t base, uint32_t& flag, C& scal, uint32_t& aux) {
  flag = (base & op.ext_cmask) == op.ext_cmask ? 1u : 0u;
  aux = 0;
  scal = cmake<C>(1, 0);
  if (op.type == OP_FAN) {
    typedef typename real_of<C>::type R;
    C s = cmake<C>((R)op.scalar[0], (R)op.scalar[1]);
    const C* tab = (const C*)(blob + op.payload);
    for (uint32_t c = 0; c < op.n_chunks; ++c) tab += (1u << op.chunk_len[c]);
    for (uint32_t e = 0; e < op.n_ext; ++e) {
      const uint64_t mask = op.ext_mask[e];
      s = cmul(s, tab[extract(base, mask)]);
      int nb = 0;
      for (uint64_t mm = mask; mm; mm &= mm - 1) ++nb;
      tab += (1u << nb);
    }
    scal = s;
  } else if (op.type == OP_DIAGK) {
    const int k = (int)op.k;
    for (int i = 0; i < k; ++i)
      if (op.tbit[i] == 0xFF && (base & op.ext_mask[i])) aux |= 1u << (k - 1 - i);
  }
}

}  // namespace qb
