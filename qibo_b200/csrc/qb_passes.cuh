// qibo_b200 K2: shared-memory passes of the sweep kernel.  Pure index math + complex arithmetic, QB_HD so
// that tests/emul compiles the very same code for the CPU to check the planner and the passes without a
// GPU.  The CUDA kernel that calls them (nct = 256 compute threads) is in qb_sweep.cuh.
//
// REGTILE pass: R tile-local bits are "register bits".  A thread owns groups of 2^R amplitudes that
// differ only in those bits: it loads a group once (LDS.128), runs the pass's whole micro-op list on it
// in registers, and stores it once -- so several gates share one trip through shared memory.  Consecutive
// threads own consecutive values of the remaining low tile bits, i.e. bank-conflict-free 16-byte accesses
// as long as the register bits are not among the lowest three (complex128) / four (complex64) bits.
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
QB_HD uint32_t deposit32(uint32_t x, uint32_t mask) {
  uint32_t r = 0;
  int k = 0;
  while (mask) {
    uint32_t low = mask & (~mask + 1);
    if ((x >> k) & 1) r |= low;
    mask ^= low;
    ++k;
  }
  return r;
}

// ---- micro-ops on a register tile v[2^R] -----------------------------------------------------------------
// Every micro-op has a straight-line fast path selected by uniform branches on (register bit, control mask),
// with all register indices and control tests resolved at compile time: the hot loop is then pure FP64/FP32
// math.  (A first version tested `(j & creg) == creg` at run time per pair: ncu showed 15 % BRA + 12 % ISETP
// + 17 % IMAD.MOV and the FP64 pipe at 18 %.)
template <typename C, int R, int I, bool REAL, uint32_t CREG>
QB_HD void mu_dense1_static(C* v, const C m00, const C m01, const C m10, const C m11) {
  typedef typename real_of<C>::type Re;
  constexpr int D = 1 << R;
#pragma unroll
  for (int j0 = 0; j0 < D; ++j0) {
    if ((j0 & (1 << I)) || (j0 & CREG) != CREG) continue;  // compile-time
    const int j1 = j0 | (1 << I);
    // cross terms first (temporaries), then one FMA per output INTO the register that holds its own input:
    // the results need no register moves (ncu: the naive form spent one IMAD.MOV per FP64 pair)
    C& a = v[j0];
    C& b = v[j1];
    if (REAL) {
      const Re r00 = m00.x, r01 = m01.x, r10 = m10.x, r11 = m11.x;
      const Re tx = r01 * b.x, ty = r01 * b.y, ux = r10 * a.x, uy = r10 * a.y;
      a.x = qfma(r00, a.x, tx);
      a.y = qfma(r00, a.y, ty);
      b.x = qfma(r11, b.x, ux);
      b.y = qfma(r11, b.y, uy);
    } else {
      // x = m00 a + m01 b ; y = m10 a + m11 b
      Re xr = m01.x * b.x, xi = m01.x * b.y, yr = m10.x * a.x, yi = m10.x * a.y;
      xr = qfma(-m01.y, b.y, xr);
      xi = qfma(m01.y, b.x, xi);
      yr = qfma(-m10.y, a.y, yr);
      yi = qfma(m10.y, a.x, yi);
      xr = qfma(-m00.y, a.y, xr);
      xi = qfma(m00.y, a.x, xi);
      yr = qfma(-m11.y, b.y, yr);
      yi = qfma(m11.y, b.x, yi);
      a.x = qfma(m00.x, a.x, xr);
      a.y = qfma(m00.x, a.y, xi);
      b.x = qfma(m11.x, b.x, yr);
      b.y = qfma(m11.x, b.y, yi);
    }
  }
}

// run-time control mask (rare: a controlled gate whose control is itself a register bit): branch-free selects
template <typename C, int R, int I> QB_HD void mu_dense1_masked(C* v, const C m00, const C m01, const C m10, const C m11, uint32_t creg) {
  constexpr int D = 1 << R;
#pragma unroll
  for (int j0 = 0; j0 < D; ++j0) {
    if (j0 & (1 << I)) continue;
    const int j1 = j0 | (1 << I);
    const bool on = (uint32_t(j0) & creg) == creg;
    const C a = v[j0], b = v[j1];
    C x = cmul(m00, a);
    cfma(x, m01, b);
    C y = cmul(m10, a);
    cfma(y, m11, b);
    v[j0] = on ? x : a;
    v[j1] = on ? y : b;
  }
}

template <typename C, int R, int I> QB_HD void mu_dense1_bit(C* v, const C* m, uint32_t creg, bool real) {
  const C m00 = m[0], m01 = m[1], m10 = m[2], m11 = m[3];
  if (creg == 0) {
    if (real) mu_dense1_static<C, R, I, true, 0>(v, m00, m01, m10, m11);
    else mu_dense1_static<C, R, I, false, 0>(v, m00, m01, m10, m11);
  } else {
    mu_dense1_masked<C, R, I>(v, m00, m01, m10, m11, creg);
  }
}

template <typename C, int R> QB_HD void mu_dense1(C* v, const C* m, uint32_t rb0, uint32_t creg, bool real) {
  switch (rb0) {
    case 0: mu_dense1_bit<C, R, 0>(v, m, creg, real); break;
    case 1: if constexpr (R > 1) mu_dense1_bit<C, R, 1>(v, m, creg, real); break;
    case 2: if constexpr (R > 2) mu_dense1_bit<C, R, 2>(v, m, creg, real); break;
    case 3: if constexpr (R > 3) mu_dense1_bit<C, R, 3>(v, m, creg, real); break;
    default: break;
  }
}

// register bits I1 (MSB of the 4x4 matrix index) and I2
template <typename C, int R, int I1, int I2> QB_HD void mu_dense2_bits(C* v, const C* m, uint32_t creg) {
  constexpr int D = 1 << R;
#pragma unroll
  for (int j = 0; j < D; ++j) {
    if (j & ((1 << I1) | (1 << I2))) continue;
    const bool on = (uint32_t(j) & creg) == creg;
    const int idx[4] = {j, j | (1 << I2), j | (1 << I1), j | (1 << I1) | (1 << I2)};
    const C in0 = v[idx[0]], in1 = v[idx[1]], in2 = v[idx[2]], in3 = v[idx[3]];
    const C in[4] = {in0, in1, in2, in3};
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      C acc = cmul(m[r * 4], in0);
      cfma(acc, m[r * 4 + 1], in1);
      cfma(acc, m[r * 4 + 2], in2);
      cfma(acc, m[r * 4 + 3], in3);
      v[idx[r]] = on ? acc : in[r];
    }
  }
}

template <typename C, int R, int I1, int I2> QB_HD void mu_swap_bits(C* v, uint32_t creg) {
  constexpr int D = 1 << R;
#pragma unroll
  for (int j = 0; j < D; ++j) {
    if (j & ((1 << I1) | (1 << I2))) continue;
    const bool on = (uint32_t(j) & creg) == creg;
    const C a = v[j | (1 << I1)], b = v[j | (1 << I2)];
    v[j | (1 << I1)] = on ? b : a;
    v[j | (1 << I2)] = on ? a : b;
  }
}

// dispatch on an (ordered) pair of register bits; F is 0 for dense2, 1 for swap
template <typename C, int R, int F> QB_HD void mu_pair(C* v, const C* m, uint32_t rb0, uint32_t rb1, uint32_t creg) {
  const uint32_t code = rb0 * 4 + rb1;
#define QB_PAIR(A, B)                                                   \
  case (A * 4 + B):                                                     \
    if constexpr ((A) < R && (B) < R) {                                 \
      if constexpr (F == 0) mu_dense2_bits<C, R, A, B>(v, m, creg);     \
      else mu_swap_bits<C, R, A, B>(v, creg);                           \
    }                                                                   \
    break;
  switch (code) {
    QB_PAIR(0, 1) QB_PAIR(0, 2) QB_PAIR(0, 3) QB_PAIR(1, 0) QB_PAIR(1, 2) QB_PAIR(1, 3)
    QB_PAIR(2, 0) QB_PAIR(2, 1) QB_PAIR(2, 3) QB_PAIR(3, 0) QB_PAIR(3, 1) QB_PAIR(3, 2)
    default: break;
  }
#undef QB_PAIR
}

// fan: phase(t) = ext_factor * TA[g & mask] * TB[g >> la] * G[j]  on the amplitudes whose controls are set.
// The control mask over the register index is resolved at compile time (2^R-way uniform dispatch).
template <typename C, int R, uint32_t CREG> QB_HD void mu_fan_static(C* v, const C p0, const C* gt) {
  constexpr int D = 1 << R;
#pragma unroll
  for (int j = 0; j < D; ++j) {
    if ((uint32_t(j) & CREG) != CREG) continue;  // compile-time
    cmul_inplace(v[j], cmul(p0, gt[j]));
  }
}
template <typename C, int R> QB_HD void mu_fan_dispatch(C* v, const C p0, const C* gt, uint32_t creg) {
#define QB_FAN(CR) \
  case CR:         \
    if constexpr ((CR) < (1 << R)) mu_fan_static<C, R, CR>(v, p0, gt); \
    break;
  switch (creg) {  // one jump table instead of a compare chain
    QB_FAN(0) QB_FAN(1) QB_FAN(2) QB_FAN(3) QB_FAN(4) QB_FAN(5) QB_FAN(6) QB_FAN(7)
    QB_FAN(8) QB_FAN(9) QB_FAN(10) QB_FAN(11) QB_FAN(12) QB_FAN(13) QB_FAN(14) QB_FAN(15)
    default: break;
  }
#undef QB_FAN
}
template <typename C, int R>
QB_HD void mu_fan(C* v, const C* ta, uint32_t la, int gbits, C ext_factor, uint32_t creg, uint32_t g) {
  const C* tb = ta + (1u << la);
  const C* gt = tb + (1u << (gbits - (int)la));
  C p0 = cmul(ext_factor, ta[g & ((1u << la) - 1)]);
  p0 = cmul(p0, tb[g >> la]);
  mu_fan_dispatch<C, R>(v, p0, gt, creg);
}

// lone controlled phase: v[j] *= ph on the register indices whose control bits are set
template <typename C, int R, uint32_t CREG> QB_HD void mu_phase_static(C* v, const C ph) {
  constexpr int D = 1 << R;
#pragma unroll
  for (int j = 0; j < D; ++j) {
    if ((uint32_t(j) & CREG) != CREG) continue;  // compile-time
    cmul_inplace(v[j], ph);
  }
}
template <typename C, int R> QB_HD void mu_phase(C* v, const C ph, uint32_t creg) {
#define QB_PH(CR) \
  case CR:        \
    if constexpr ((CR) < (1 << R)) mu_phase_static<C, R, CR>(v, ph); \
    break;
  switch (creg) {
    QB_PH(0) QB_PH(1) QB_PH(2) QB_PH(3) QB_PH(4) QB_PH(5) QB_PH(6) QB_PH(7)
    QB_PH(8) QB_PH(9) QB_PH(10) QB_PH(11) QB_PH(12) QB_PH(13) QB_PH(14) QB_PH(15)
    default: break;
  }
#undef QB_PH
}

template <typename C, int R> QB_HD void mu_diagk(C* v, const MicroOp& mo, const char* blob, uint32_t aux, uint32_t t0) {
  constexpr int D = 1 << R;
  const C* tab = reinterpret_cast<const C*>(blob + mo.payload);
  const int k = (int)mo.k;
  uint32_t base = aux;
  for (int i = 0; i < k; ++i)
    if (mo.tbit[i] != 0xFF) base |= ((t0 >> mo.tbit[i]) & 1u) << (k - 1 - i);
  const uint32_t creg = mo.creg;
#pragma unroll
  for (int j = 0; j < D; ++j) {
    uint32_t idx = base;
    for (int i = 0; i < k; ++i)
      if (mo.rsel[i] != 0xFF) idx |= ((uint32_t(j) >> mo.rsel[i]) & 1u) << (k - 1 - i);
    const C ph = cmul(v[j], tab[idx]);
    v[j] = (uint32_t(j) & creg) == creg ? ph : v[j];
  }
}

// One REGTILE pass over the tile.  Each thread keeps GPT groups in registers at once, so that the decode of a
// micro-op (one 16-byte header load + its inline matrix) is paid once per GPT * 2^R amplitudes.
struct MicroHot { uint32_t w0, creg, cthr, active; };  // first 16 bytes of MicroOp

template <typename C, int R, int GPT>
QB_HD void pass_regtile(C* tile, const char* blob, const PassHeader& ph, int T, uint32_t ctid, uint32_t nct) {
  constexpr int D = 1 << R;
  const uint32_t rmask = ph.rmask;
  uint32_t off[D];
#pragma unroll
  for (int j = 0; j < D; ++j) off[j] = ph.off[j];
  const MicroOp* mops = reinterpret_cast<const MicroOp*>(blob + ph.offset);
  const int nmicro = (int)ph.nmicro;
  const int gbits = T - R;
  const uint32_t ngroups = 1u << gbits;  // the caller guarantees ngroups % (nct * GPT) == 0 or ngroups < nct with GPT == 1
  for (uint32_t gbase = ctid; gbase < ngroups; gbase += nct * GPT) {
    C v[GPT][D];
    uint32_t g[GPT], t0[GPT];
#pragma unroll
    for (int u = 0; u < GPT; ++u) {
      g[u] = gbase + (uint32_t)u * nct;
      t0[u] = expand_mask(g[u], rmask);
#pragma unroll
      for (int j = 0; j < D; ++j) v[u][j] = tile[t0[u] | off[j]];
    }
    for (int mi = 0; mi < nmicro; ++mi) {
      const MicroOp& mo = mops[mi];
      const MicroHot hot = *reinterpret_cast<const MicroHot*>(&mo);
      if (!hot.active) continue;
      const uint32_t type = hot.w0 & 0xFF, rb0 = (hot.w0 >> 8) & 0xFF, rb1 = (hot.w0 >> 16) & 0xFF, flags = hot.w0 >> 24;
      bool run[GPT];
#pragma unroll
      for (int u = 0; u < GPT; ++u) run[u] = (t0[u] & hot.cthr) == hot.cthr;
      switch (type) {
        case MU_DENSE1: {
          const C* mi_ = reinterpret_cast<const C*>(mo.inl);
          const C m[4] = {mi_[0], mi_[1], mi_[2], mi_[3]};
#pragma unroll
          for (int u = 0; u < GPT; ++u)
            if (run[u]) mu_dense1<C, R>(v[u], m, rb0, hot.creg, (flags & MU_REAL) != 0);
        } break;
        case MU_DENSE2: {
          const C* m = reinterpret_cast<const C*>(blob + mo.payload);
#pragma unroll
          for (int u = 0; u < GPT; ++u)
            if (run[u]) mu_pair<C, R, 0>(v[u], m, rb0, rb1, hot.creg);
        } break;
        case MU_SWAP: {
#pragma unroll
          for (int u = 0; u < GPT; ++u)
            if (run[u]) mu_pair<C, R, 1>(v[u], (const C*)nullptr, rb0, rb1, hot.creg);
        } break;
        case MU_FAN: {
          const C* ta = reinterpret_cast<const C*>(blob + mo.payload);
          const C ext = *reinterpret_cast<const C*>(mo.inl);
          const uint32_t la = mo.la;
#pragma unroll
          for (int u = 0; u < GPT; ++u)
            if (run[u]) mu_fan<C, R>(v[u], ta, la, gbits, ext, hot.creg, g[u]);
        } break;
        case MU_DIAGK: {
#pragma unroll
          for (int u = 0; u < GPT; ++u)
            if (run[u]) mu_diagk<C, R>(v[u], mo, blob, mo.aux, t0[u]);
        } break;
        case MU_PHASE: {
          const C ph = *reinterpret_cast<const C*>(mo.inl);
#pragma unroll
          for (int u = 0; u < GPT; ++u)
            if (run[u]) mu_phase<C, R>(v[u], ph, hot.creg);
        } break;
        default: break;
      }
    }
#pragma unroll
    for (int u = 0; u < GPT; ++u) {
#pragma unroll
      for (int j = 0; j < D; ++j) tile[t0[u] | off[j]] = v[u][j];
    }
  }
}

// picks the groups-per-thread variant: two groups in flight when the tile has enough of them
template <typename C, int R> QB_HD void run_regtile(C* tile, const char* blob, const PassHeader& ph, int T, uint32_t ctid, uint32_t nct) {
#ifndef QB_MAX_GPT
#define QB_MAX_GPT 2
#endif
  const uint32_t ngroups = 1u << (T - R);
  // two groups in flight only while the register tile stays <= 64 data registers per thread
  if constexpr (QB_MAX_GPT >= 2 && (sizeof(C) << R) <= 128) {
    if (ngroups >= 2 * nct) {
      pass_regtile<C, R, 2>(tile, blob, ph, T, ctid, nct);
      return;
    }
  }
  pass_regtile<C, R, 1>(tile, blob, ph, T, ctid, nct);
}

// ---- BIG pass: one dense gate on k = 3..6 tile-local targets ------------------------------------------------
// 2^(k-3) threads share one group, 8 output rows each; inputs are re-read from shared memory.  Split in a
// read/accumulate half and a write half with a barrier between (the caller provides it).
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

// ---- per-tile set-up of one micro-op: written INTO the shared-memory copy of the op -------------------------
template <typename C> QB_HD void micro_prephase(MicroOp& mo, const char* blob, uint64_t base, int T) {
  const int R = (int)mo.R, gbits = T - R;
  mo.active = (base & mo.ext_cmask) == mo.ext_cmask ? 1u : 0u;
  if (mo.type == MU_FAN) {
    typedef typename real_of<C>::type Re;
    C s = cmake<C>((Re)mo.scalar[0], (Re)mo.scalar[1]);
    const C* tab = reinterpret_cast<const C*>(blob + mo.payload);
    tab += (1u << mo.la) + (1u << (gbits - (int)mo.la)) + (1u << R);
    for (uint32_t e = 0; e < mo.n_ext; ++e) {
      const uint64_t mask = mo.ext_mask[e];
      s = cmul(s, tab[extract(base, mask)]);
      int nb = 0;
      for (uint64_t mm = mask; mm; mm &= mm - 1) ++nb;
      tab += (1u << nb);
    }
    *reinterpret_cast<C*>(mo.inl) = s;
  } else if (mo.type == MU_DIAGK) {
    const int k = (int)mo.k;
    uint32_t aux = 0;
    for (int i = 0; i < k; ++i)
      if (mo.ext_mask[i] && (base & mo.ext_mask[i])) aux |= 1u << (k - 1 - i);
    mo.aux = aux;
  }
}

}  // namespace qb
