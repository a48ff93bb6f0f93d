// qibo_b200 K2: shared-memory passes of the sweep kernel.  Pure index math + complex arithmetic, QB_HD so
// that tests/emul compiles the very same code for the CPU to check the planner and the passes without a
// GPU.  The CUDA kernel that calls them is in qb_sweep.cuh.
//
// REGTILE pass: R tile-local bits are "register bits".  A thread owns groups of 2^R amplitudes that
// differ only in those bits: it loads a group once (LDS.128), runs the pass's whole micro-op list on it
// in registers, and stores it once -- so several gates share one trip through shared memory.  Consecutive
// threads own consecutive values of the remaining low tile bits, i.e. bank-conflict-free 16-byte accesses
// as long as the register bits are not among the lowest three (complex128) / four (complex64) bits.
//
// Every micro-op carries ONE handler code chosen by the planner (qb_planner.hpp, MicroHandler): the kernel
// decodes an op with a single 16-byte shared-memory load and a single jump table; each handler is straight-
// line code with its register indices (and, for the common cases, its control test) fixed at compile time.
// ncu history (profiles/): a first version tested `(j & creg) == creg` per pair at run time and selected
// results with FSEL -- twice the FP64 work and 2.5 non-FP64 instructions per FP64 one; a second one nested
// four switches per op and grew to 26 k SASS instructions (instruction-cache misses, 9 % of the issue slots
// in a software pdep).  This version keeps the per-op overhead at ~20 instructions.
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

// Shared-memory layout of a tile: linear, or the TMA 128-byte swizzle (the 16-byte chunk index inside a 128-byte row
// is XORed with the row index mod 8), which spreads amplitudes that differ only in bits >= 3 over the banks -- without
// it a pass whose register bits are the lowest tile bits is an 8-way bank conflict.  `on` is 7 (swizzled) or 0.
// swz is linear over XOR, so swz(t0 | off) = swz(t0) ^ swz(off) for disjoint t0 / off.
template <typename C> QB_HD uint32_t swz(uint32_t x, uint32_t on) {
  return sizeof(C) == 16 ? x ^ ((x >> 3) & on) : x ^ (((x >> 4) & on) << 1);
}

// bit that changes between the k-th and the (k+1)-th Gray code: the number of trailing ones of k
constexpr QB_HD int gray_flip(int k) {
  int b = 0;
  while (k & 1) {
    k >>= 1;
    ++b;
  }
  return b;
}

template <typename C> QB_HD C slot_ext(const TileSlot& s) { return *reinterpret_cast<const C*>(s.ext); }

// ---- micro-op bodies on a register tile v[2^R] --------------------------------------------------------------
// MODE: 0 = no register-bit control (static), 1 = run-time control mask `creg` (uniform branch per pair)
template <typename C, int R, int I, int MODE> QB_HD void mu_addsub(C* v, uint32_t creg) {
  constexpr int D = 1 << R;
#pragma unroll
  for (int j0 = 0; j0 < D; ++j0) {
    if (j0 & (1 << I)) continue;
    if (MODE == 1 && (uint32_t(j0) & creg) != creg) continue;
    // a' = a + b in place, then b' = a - b as a' - 2b: one FMA into b's own register.  (The textbook form needs a
    // temporary, which costs one register move per add in the SASS because every handler must leave v[] in place.)
    C& a = v[j0];
    C& b = v[j0 | (1 << I)];
    typedef typename real_of<C>::type Re;
    a = creal_add(a, b);
    b = creal_fma((Re)-2, b, a);
  }
}

template <typename C, int R, int I, int MODE> QB_HD void mu_real1(C* v, const C* m, uint32_t creg) {
  typedef typename real_of<C>::type Re;
  constexpr int D = 1 << R;
  const Re r00 = m[0].x, r01 = m[1].x, r10 = m[2].x, r11 = m[3].x;
#pragma unroll
  for (int j0 = 0; j0 < D; ++j0) {
    if (j0 & (1 << I)) continue;
    if (MODE == 1 && (uint32_t(j0) & creg) != creg) continue;
    // cross terms first (temporaries), then one FMA per output INTO the register that holds its own input
    C& a = v[j0];
    C& b = v[j0 | (1 << I)];
    const C t = creal_mul(r01, b), u = creal_mul(r10, a);
    a = creal_fma(r00, a, t);
    b = creal_fma(r11, b, u);
  }
}

template <typename C, int R, int I, int MODE> QB_HD void mu_cplx1(C* v, const C* m, uint32_t creg) {
  typedef typename real_of<C>::type Re;
  constexpr int D = 1 << R;
  const C m00 = m[0], m01 = m[1], m10 = m[2], m11 = m[3];
#pragma unroll
  for (int j0 = 0; j0 < D; ++j0) {
    if (j0 & (1 << I)) continue;
    if (MODE == 1 && (uint32_t(j0) & creg) != creg) continue;
    C& a = v[j0];
    C& b = v[j0 | (1 << I)];
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

// pair exchange (X / CNOT / TOFFOLI target): no arithmetic
template <typename C, int R, int I> QB_HD void mu_xpair(C* v, uint32_t creg) {
  constexpr int D = 1 << R;
#pragma unroll
  for (int j0 = 0; j0 < D; ++j0) {
    if (j0 & (1 << I)) continue;
    if ((uint32_t(j0) & creg) != creg) continue;
    const C t = v[j0];
    v[j0] = v[j0 | (1 << I)];
    v[j0 | (1 << I)] = t;
  }
}

// register bits HI (MSB of the 4x4 matrix index) > LO; the planner reorders the matrix when needed
template <typename C, int R, int HI, int LO> QB_HD void mu_dense2(C* v, const C* m, uint32_t creg) {
  constexpr int D = 1 << R;
#pragma unroll
  for (int j = 0; j < D; ++j) {
    if (j & ((1 << HI) | (1 << LO))) continue;
    if ((uint32_t(j) & creg) != creg) continue;
    const int idx[4] = {j, j | (1 << LO), j | (1 << HI), j | (1 << HI) | (1 << LO)};
    const C in0 = v[idx[0]], in1 = v[idx[1]], in2 = v[idx[2]], in3 = v[idx[3]];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      C acc = cmul(m[r * 4], in0);
      cfma(acc, m[r * 4 + 1], in1);
      cfma(acc, m[r * 4 + 2], in2);
      cfma(acc, m[r * 4 + 3], in3);
      v[idx[r]] = acc;
    }
  }
}

template <typename C, int R, int HI, int LO> QB_HD void mu_swap(C* v, uint32_t creg) {
  constexpr int D = 1 << R;
#pragma unroll
  for (int j = 0; j < D; ++j) {
    if (j & ((1 << HI) | (1 << LO))) continue;
    if ((uint32_t(j) & creg) != creg) continue;
    const C t = v[j | (1 << HI)];
    v[j | (1 << HI)] = v[j | (1 << LO)];
    v[j | (1 << LO)] = t;
  }
}

// fan: phase(t) = ext_factor * TA[g & mask] * TB[g >> la] * G[j]  on the amplitudes whose controls are set.
// CB: 0..R-1 = the only register-bit control is bit CB (static) AND the fan has no factor on register bits above CB
// (the planner checks it; true for every QFT stage, whose fan only reaches the qubits after the control): the 2^(R-1)
// controlled amplitudes then share 2^CB distinct phases -- 47 instead of 64 complex products per four-stage pass;
// R = no register-bit control; R+1 = run-time mask
template <typename C, int R, int CB> QB_HD void mu_fan(C* v, const C p0, const C* gt, uint32_t creg) {
  constexpr int D = 1 << R;
  if constexpr (CB < R) {
    constexpr int ND = 1 << CB;
    C ph[ND];
#pragma unroll
    for (int jl = 0; jl < ND; ++jl) ph[jl] = cmul(p0, gt[jl | ND]);
#pragma unroll
    for (int j = 0; j < D; ++j) {
      if (!((j >> CB) & 1)) continue;
      cmul_inplace(v[j], ph[j & (ND - 1)]);
    }
  } else {
#pragma unroll
    for (int j = 0; j < D; ++j) {
      if (CB == R + 1 && (uint32_t(j) & creg) != creg) continue;
      cmul_inplace(v[j], cmul(p0, gt[j]));
    }
  }
}
// lone controlled phase: v[j] *= ph on the register indices whose control bits are set
template <typename C, int R, int CB> QB_HD void mu_phase(C* v, const C ph, uint32_t creg) {
  constexpr int D = 1 << R;
  if (ph.y == 0) {  // CZ, Z...: a real factor is half the multiplies (one packed instruction for complex64)
#pragma unroll
    for (int j = 0; j < D; ++j) {
      if (CB < R && !((j >> CB) & 1)) continue;
      if (CB == R + 1 && (uint32_t(j) & creg) != creg) continue;
      v[j] = creal_mul(ph.x, v[j]);
    }
    return;
  }
#pragma unroll
  for (int j = 0; j < D; ++j) {
    if (CB < R && !((j >> CB) & 1)) continue;
    if (CB == R + 1 && (uint32_t(j) & creg) != creg) continue;
    cmul_inplace(v[j], ph);
  }
}

// lone phase controlled by the two register bits HI > LO: the 2^(R-2) register indices with both bits set
template <typename C, int R, int HI, int LO> QB_HD void mu_phase2(C* v, const C ph) {
  constexpr int D = 1 << R;
  constexpr int M = (1 << HI) | (1 << LO);
  if (ph.y == 0) {
#pragma unroll
    for (int j = 0; j < D; ++j)
      if ((j & M) == M) v[j] = creal_mul(ph.x, v[j]);
    return;
  }
#pragma unroll
  for (int j = 0; j < D; ++j)
    if ((j & M) == M) cmul_inplace(v[j], ph);
}

template <typename C, int R> QB_HD void mu_diagk(C* v, const MicroOp& mo, const char* blob, uint32_t aux, uint32_t t0) {
  constexpr int D = 1 << R;
  const C* tab = reinterpret_cast<const C*>(blob + mo.payload);
  const int k = (int)mo.k;
  uint32_t base = aux;
  for (int i = 0; i < k; ++i)
    if (mo.tbit[i] != 0xFF) base |= ((t0 >> mo.tbit[i]) & 1u) << (k - 1 - i);
  const uint32_t creg = mo.creg;
  uint32_t rsel_bits[R > 0 ? R : 1];  // table-index bit contributed by register bit i (0 when it is not a target)
#pragma unroll
  for (int i = 0; i < R; ++i) rsel_bits[i] = 0;
  for (int i = 0; i < k; ++i)
    if (mo.rsel[i] != 0xFF) {
#pragma unroll
      for (int r = 0; r < R; ++r)
        if (mo.rsel[i] == r) rsel_bits[r] = 1u << (k - 1 - i);
    }
#pragma unroll
  for (int j = 0; j < D; ++j) {
    if ((uint32_t(j) & creg) != creg) continue;
    uint32_t idx = base;
#pragma unroll
    for (int r = 0; r < R; ++r)
      if ((j >> r) & 1) idx |= rsel_bits[r];
    cmul_inplace(v[j], tab[idx]);
  }
}

// a set of +-1 diagonal gates (MH_SIGNS, qb_planner.hpp): sign of register index j = bit j of the mask built here
template <typename C, int R> QB_HD void mu_signs(C* v, uint32_t t0, uint32_t g16, uint32_t zmask, const uint16_t* sp, uint32_t aux) {
  constexpr int D = 1 << R;
  uint32_t m = g16;
#pragma unroll
  for (int r = 0; r < R; ++r) {
    uint32_t mr = 0;
#pragma unroll
    for (int j = 0; j < D; ++j)
      if ((j >> r) & 1) mr |= 1u << j;
    if ((popc32(t0 & sp[r]) ^ (aux >> r)) & 1u) m ^= mr;
  }
  uint32_t c = popc32(t0 & (zmask & 0xFFFFu)) ^ (aux >> 4) ^ (zmask >> 31);
  const uint32_t np = sp[4];
  for (uint32_t k = 0; k < np; ++k) {
    const uint32_t pm = sp[5 + k];
    c ^= (t0 & pm) == pm ? 1u : 0u;
  }
  if (c & 1u) m = ~m;
#pragma unroll
  for (int j = 0; j < D; ++j) flip_sign(v[j], (m >> j) & 1u);
}

struct alignas(16) MicroHot { uint32_t w0, creg, cthr, payload; };  // first 16 bytes of MicroOp

// one fused stage: 2x2 gate on register bit I ((a+b, a-b) or a real matrix), then the fan controlled by bit I
template <typename C, int R, int GPT, int I>
QB_HD void stage_op(C (&v)[GPT][1 << R], const uint32_t* g, const bool* valid, const MicroOp& mo, const char* blob, const TileSlot* ts, int gbits) {
  const MicroHot hot = *reinterpret_cast<const MicroHot*>(&mo);
  const uint32_t handler = hot.w0 & 0xFF, slot = hot.w0 >> 16;
  const bool has_fan = handler >= MH_STAGE_A;  // else a bare (a+b, a-b) / real 2x2 gate (the last H of a QFT has no fan)
  if (has_fan && !ts[slot].active) return;
  if (handler >= MH_STAGE_R || (handler >= MH_REAL1 && handler < MH_REAL1 + 4)) {
    const C* inl = reinterpret_cast<const C*>(mo.inl);
    const C m[4] = {inl[0], inl[1], inl[2], inl[3]};
#pragma unroll
    for (int u = 0; u < GPT; ++u)
      if (valid[u]) mu_real1<C, R, I, 0>(v[u], m, 0u);
  } else {
#pragma unroll
    for (int u = 0; u < GPT; ++u)
      if (valid[u]) mu_addsub<C, R, I, 0>(v[u], 0u);
  }
  if (!has_fan) return;
  const C* ta = reinterpret_cast<const C*>(blob + hot.payload);
  const uint32_t la = mo.la;
  const C* tb = ta + (1u << la);
  const C* gt = tb + (1u << (gbits - (int)la));
  const C ext = slot_ext<C>(ts[slot]);
#pragma unroll
  for (int u = 0; u < GPT; ++u)
    if (valid[u]) {
      C ph0 = cmul(ext, ta[g[u] & ((1u << la) - 1)]);
      ph0 = cmul(ph0, tb[g[u] >> la]);
      mu_fan<C, R, I>(v[u], ph0, gt, 0u);
    }
}
template <typename C, int R, int GPT, int I>
QB_HD void stage_chain(C (&v)[GPT][1 << R], const uint32_t* g, const bool* valid, const MicroOp* mops, int& mi, uint32_t mask, const char* blob,
                       const TileSlot* ts, int gbits) {
  if constexpr (I >= 0) {
    if ((mask >> I) & 1) stage_op<C, R, GPT, I>(v, g, valid, mops[mi++], blob, ts, gbits);
    stage_chain<C, R, GPT, I - 1>(v, g, valid, mops, mi, mask, blob, ts, gbits);
  }
}

// PASS_FULL_STAGE: R add/sub stages with their fans on register bits R-1 .. 0 as one basic block (qb_planner.hpp)
struct alignas(16) MicroTab { uint16_t la, R; uint32_t tb_off, n_ext, gt_off; };  // second 16 bytes of MicroOp
template <typename C, int R, int GPT, int I, uint32_t SMASK, uint32_t RMASK>
QB_HD void stage_full(C (&v)[GPT][1 << R], const uint32_t* g, const MicroOp* mops, const char* blob, const TileSlot* ts) {
  if constexpr (I >= 0) {
    if constexpr ((SMASK >> I) & 1u) {
      constexpr int INDEX = __builtin_popcount(SMASK >> (I + 1));  // ops are stored by descending register bit
      const MicroOp& mo = mops[INDEX];
      const MicroHot hot = *reinterpret_cast<const MicroHot*>(&mo);
      const MicroTab tab = *reinterpret_cast<const MicroTab*>(reinterpret_cast<const char*>(&mo) + 16);
      const C ext = slot_ext<C>(ts[hot.w0 >> 16]);
      const C* ta = reinterpret_cast<const C*>(blob + hot.payload);
      const C* tb = reinterpret_cast<const C*>(blob + tab.tb_off);
      const C* gt = reinterpret_cast<const C*>(blob + tab.gt_off);
      const uint32_t la = tab.la, ma = (1u << la) - 1u;
      if constexpr ((RMASK >> I) & 1u) {
        const C* inl = reinterpret_cast<const C*>(mo.inl);
        const C m[4] = {inl[0], inl[1], inl[2], inl[3]};
#pragma unroll
        for (int u = 0; u < GPT; ++u) mu_real1<C, R, I, 0>(v[u], m, 0u);
      } else {
#pragma unroll
        for (int u = 0; u < GPT; ++u) mu_addsub<C, R, I, 0>(v[u], 0u);
      }
#pragma unroll
      for (int u = 0; u < GPT; ++u) {
        C ph0 = cmul(ext, ta[g[u] & ma]);
        ph0 = cmul(ph0, tb[g[u] >> la]);
        mu_fan<C, R, I>(v[u], ph0, gt, 0u);
      }
    }
    stage_full<C, R, GPT, I - 1, SMASK, RMASK>(v, g, mops, blob, ts);
  }
}

template <typename C> struct alignas(2 * sizeof(C)) Pair { C a, b; };  // two adjacent amplitudes: one 16-byte access for complex64

// One REGTILE pass over the tile.  Each thread keeps GPT groups in registers at once, so that the decode of a
// micro-op is paid once per GPT * 2^R amplitudes.  `ts` = this team's per-tile slot states.
// A PASS_PERMUTED_STORE pass (the last pass of a permuting sweep) writes its groups to `out` (the kernel: the same buffer,
// after `barrier()` -- every thread of the team has loaded its groups by then; the CPU emulation: a second buffer) at the
// DESTINATION-layout addresses: dtab gives the destination-local index of each group base, dpos the destination
// positions of the register bits.
struct NoBarrier { QB_HD void operator()() const {} };
template <typename C, int R, int GPT, bool SO = false, typename Barrier = NoBarrier>
QB_HD void run_pass(C* tile, const char* blob, const TileSlot* ts, const PassHeader& ph, int T, uint32_t swz_on, uint32_t ctid, uint32_t nct,
                    C* out = nullptr, Barrier barrier = Barrier()) {
  constexpr int D = 1 << R;
  const int gbits = T - R;
  uint32_t stride[R > 0 ? R : 1];  // physical (swizzled) offset of register bit i (swz is linear over XOR)
#pragma unroll
  for (int i = 0; i < R; ++i) stride[i] = swz<C>(1u << ph.pos[i], swz_on);
  const uint16_t* gtab = reinterpret_cast<const uint16_t*>(blob + ph.gtab);
  C v[GPT][D];
  uint32_t g[GPT], t0[GPT], p0[GPT];  // group index, logical tile index of the group base, its physical index
  bool valid[GPT];
#pragma unroll
  for (int u = 0; u < GPT; ++u) {
    g[u] = gtab[ctid + (uint32_t)u * nct];
    valid[u] = g[u] != 0xFFFFu;
    uint32_t t = g[u];
#pragma unroll
    for (int i = 0; i < R; ++i) t = insert_zero32(t, ph.pos[i]);
    t0[u] = t;
    p0[u] = swz<C>(t, swz_on);
  }
  // PASS_PAIRED_GROUPS (complex64, two groups per thread that differ in tile bit 0): one 16-byte access per pair
  bool paired = false;
  if constexpr (GPT == 2 && sizeof(C) == 8) paired = (ph.flags & PASS_PAIRED_GROUPS) != 0 && !(ph.flags & PASS_PERMUTED_STORE);
  if constexpr (GPT == 2 && sizeof(C) == 8) {
    if (paired) {
      uint32_t o = p0[0];  // even: group 2c, register bits zero; the partner amplitude of group 2c + 1 is the next element
#pragma unroll
      for (int k = 0; k < D; ++k) {
        const Pair<C> w = *reinterpret_cast<const Pair<C>*>(&tile[o]);
        v[0][k ^ (k >> 1)] = w.a;
        v[1][k ^ (k >> 1)] = w.b;
        if (k + 1 < D) o ^= stride[gray_flip(k)];
      }
    }
  }
#pragma unroll
  for (int u = 0; u < GPT; ++u) {
    if (valid[u] && !paired) {
      // Gray-code walk over the 2^R register indices: consecutive addresses differ by ONE stride (ncu, round 2: the
      // address arithmetic of the tile loads / stores was 20 % of a layered sweep's instructions)
      uint32_t o = p0[u];
#pragma unroll
      for (int k = 0; k < D; ++k) {
        v[u][k ^ (k >> 1)] = tile[o];
        if (k + 1 < D) o ^= stride[gray_flip(k)];
      }
    }
  }
  const MicroOp* mops = reinterpret_cast<const MicroOp*>(blob + ph.offset);
  const int nmicro = (int)ph.nmicro;
  if ((ph.flags & PASS_FULL_STAGE) && R == 4) {
    switch (ph.stage_mask | (ph.flags & 0xF0u)) {
      case 0x0F: stage_full<C, R, GPT, R - 1, 0xFu, 0x0u>(v, g, mops, blob, ts); break;
      case 0x1F: stage_full<C, R, GPT, R - 1, 0xFu, 0x1u>(v, g, mops, blob, ts); break;
      case 0x07: stage_full<C, R, GPT, R - 1, 0x7u, 0x0u>(v, g, mops, blob, ts); break;
      default: stage_full<C, R, GPT, R - 1, 0x7u, 0x1u>(v, g, mops, blob, ts); break;
    }
  } else if (ph.stage_mask) {
    // a pass made of fused stage ops on descending register bits (every QFT pass): straight-line code, no per-op
    // jump sequence (ncu: the compare tree nvcc makes of the handler switch costs ~5 dependent branches per op)
    int mi = 0;
    stage_chain<C, R, GPT, R - 1>(v, g, valid, mops, mi, ph.stage_mask, blob, ts, gbits);
  } else if constexpr (!SO) {  // (a stage-only kernel is launched for sweeps made of stage passes alone)
  for (int mi = 0; mi < nmicro; ++mi) {
    const MicroOp& mo = mops[mi];
    const MicroHot hot = *reinterpret_cast<const MicroHot*>(&mo);
    const uint32_t handler = hot.w0 & 0xFF, slot = hot.w0 >> 16;
    if (slot != MU_NO_SLOT && !ts[slot].active) continue;
    bool run[GPT];
#pragma unroll
    for (int u = 0; u < GPT; ++u) run[u] = valid[u] && (t0[u] & hot.cthr) == hot.cthr;
    const C* inl = reinterpret_cast<const C*>(mo.inl);
#define QB_EACH(...)                                \
  {                                                 \
    _Pragma("unroll") for (int u = 0; u < GPT; ++u) \
      if (run[u]) { __VA_ARGS__; }                  \
  }
#define QB_CASE_BIT(BASE, ...)                                                        \
  case BASE + 0: { constexpr int I = 0; if constexpr (I < R) { __VA_ARGS__ } } break; \
  case BASE + 1: { constexpr int I = 1; if constexpr (I < R) { __VA_ARGS__ } } break; \
  case BASE + 2: { constexpr int I = 2; if constexpr (I < R) { __VA_ARGS__ } } break; \
  case BASE + 3: { constexpr int I = 3; if constexpr (I < R) { __VA_ARGS__ } } break;
#define QB_CASE_PAIR(BASE, HI_, LO_, K, ...) \
  case BASE + K: { constexpr int HI = HI_, LO = LO_; if constexpr (HI < R) { __VA_ARGS__ } } break;
#define QB_CASE_PAIRS(BASE, ...)                                                                                          \
  QB_CASE_PAIR(BASE, 1, 0, 0, __VA_ARGS__) QB_CASE_PAIR(BASE, 2, 0, 1, __VA_ARGS__) QB_CASE_PAIR(BASE, 2, 1, 2, __VA_ARGS__) \
  QB_CASE_PAIR(BASE, 3, 0, 3, __VA_ARGS__) QB_CASE_PAIR(BASE, 3, 1, 4, __VA_ARGS__) QB_CASE_PAIR(BASE, 3, 2, 5, __VA_ARGS__)
#define QB_FAN_BODY(CB)                                                              \
  {                                                                                  \
    const C* ta = reinterpret_cast<const C*>(blob + hot.payload);                    \
    const uint32_t la = mo.la;                                                       \
    const C* tb = ta + (1u << la);                                                   \
    const C* gt = tb + (1u << (gbits - (int)la));                                    \
    const C ext = slot_ext<C>(ts[slot]);                                             \
    QB_EACH({                                                                        \
      C ph0 = cmul(ext, ta[g[u] & ((1u << la) - 1)]);                                \
      ph0 = cmul(ph0, tb[g[u] >> la]);                                               \
      mu_fan<C, R, CB>(v[u], ph0, gt, hot.creg);                                     \
    })                                                                               \
  }
    switch (handler) {
      QB_CASE_BIT(MH_STAGE_A, { QB_EACH((mu_addsub<C, R, I, 0>(v[u], 0u))) QB_FAN_BODY(I) })
      QB_CASE_BIT(MH_STAGE_R, { const C m[4] = {inl[0], inl[1], inl[2], inl[3]}; QB_EACH((mu_real1<C, R, I, 0>(v[u], m, 0u))) QB_FAN_BODY(I) })
      QB_CASE_BIT(MH_ADDSUB, QB_EACH((mu_addsub<C, R, I, 0>(v[u], 0u))))
      QB_CASE_BIT(MH_REAL1, { const C m[4] = {inl[0], inl[1], inl[2], inl[3]}; QB_EACH((mu_real1<C, R, I, 0>(v[u], m, 0u))) })
      QB_CASE_BIT(MH_CPLX1, { const C m[4] = {inl[0], inl[1], inl[2], inl[3]}; QB_EACH((mu_cplx1<C, R, I, 0>(v[u], m, 0u))) })
      QB_CASE_BIT(MH_CPLX1_M, { const C m[4] = {inl[0], inl[1], inl[2], inl[3]}; QB_EACH((mu_cplx1<C, R, I, 1>(v[u], m, hot.creg))) })
      QB_CASE_BIT(MH_XPAIR, QB_EACH((mu_xpair<C, R, I>(v[u], hot.creg))))
      QB_CASE_PAIRS(MH_DENSE2, { const C* m = reinterpret_cast<const C*>(blob + hot.payload); QB_EACH((mu_dense2<C, R, HI, LO>(v[u], m, hot.creg))) })
      QB_CASE_PAIRS(MH_SWAP, QB_EACH((mu_swap<C, R, HI, LO>(v[u], hot.creg))))
      QB_CASE_BIT(MH_FAN_C, QB_FAN_BODY(I))
      case MH_FAN_NC: QB_FAN_BODY(R) break;
      case MH_FAN_M: QB_FAN_BODY(R + 1) break;
      QB_CASE_BIT(MH_PHASE_C, { const C ph = inl[0]; QB_EACH((mu_phase<C, R, I>(v[u], ph, 0u))) })
      case MH_PHASE_NC: { const C ph = inl[0]; QB_EACH((mu_phase<C, R, R>(v[u], ph, 0u))) } break;
      case MH_PHASE_M: { const C ph = inl[0]; QB_EACH((mu_phase<C, R, R + 1>(v[u], ph, hot.creg))) } break;
      QB_CASE_PAIRS(MH_PHASE_C2, { const C ph = inl[0]; QB_EACH((mu_phase2<C, R, HI, LO>(v[u], ph))) })
      case MH_DIAGK: QB_EACH((mu_diagk<C, R>(v[u], mo, blob, ts[slot].aux, t0[u]))) break;
      case MH_SIGNS: {  // (creg / cthr carry the sign table and the thread-bit mask: `run` does not apply)
        const uint16_t* sp = reinterpret_cast<const uint16_t*>(mo.inl);
        const uint32_t aux = slot != MU_NO_SLOT ? ts[slot].aux : 0u;
#pragma unroll
        for (int u = 0; u < GPT; ++u)
          if (valid[u]) mu_signs<C, R>(v[u], t0[u], hot.creg, hot.cthr, sp, aux);
      } break;
      case MH_REAL_LAYER: {
        const C* m = reinterpret_cast<const C*>(blob + hot.payload);
        const uint32_t present = mo.k;
        if (present & 1u) QB_EACH((mu_real1<C, R, 0, 0>(v[u], m, 0u)))
        if constexpr (R > 1) { if (present & 2u) QB_EACH((mu_real1<C, R, 1, 0>(v[u], m + 4, 0u))) }
        if constexpr (R > 2) { if (present & 4u) QB_EACH((mu_real1<C, R, 2, 0>(v[u], m + 8, 0u))) }
        if constexpr (R > 3) { if (present & 8u) QB_EACH((mu_real1<C, R, 3, 0>(v[u], m + 12, 0u))) }
      } break;
      default: break;
    }
#undef QB_FAN_BODY
#undef QB_CASE_PAIRS
#undef QB_CASE_PAIR
#undef QB_CASE_BIT
#undef QB_EACH
  }
  }
  if (ph.flags & PASS_PERMUTED_STORE) {
    const uint32_t dswz_on = (ph.flags & PASS_PERMUTED_DSWZ) ? 7u : 0u;
    const uint16_t* dtab = reinterpret_cast<const uint16_t*>(blob + ph.dtab);
#pragma unroll
    for (int i = 0; i < R; ++i) stride[i] = swz<C>(1u << ph.dpos[i], dswz_on);
#pragma unroll
    for (int u = 0; u < GPT; ++u) p0[u] = swz<C>(dtab[ctid + (uint32_t)u * nct], dswz_on);
    barrier();
    if (out) tile = out;
  }
  if constexpr (GPT == 2 && sizeof(C) == 8) {
    if (paired) {
      uint32_t o = p0[0];
#pragma unroll
      for (int k = 0; k < D; ++k) {
        Pair<C> w;
        w.a = v[0][k ^ (k >> 1)];
        w.b = v[1][k ^ (k >> 1)];
        *reinterpret_cast<Pair<C>*>(&tile[o]) = w;
        if (k + 1 < D) o ^= stride[gray_flip(k)];
      }
      return;
    }
  }
#pragma unroll
  for (int u = 0; u < GPT; ++u) {
    if (valid[u]) {
      uint32_t o = p0[u];
#pragma unroll
      for (int k = 0; k < D; ++k) {
        tile[o] = v[u][k ^ (k >> 1)];
        if (k + 1 < D) o ^= stride[gray_flip(k)];
      }
    }
  }
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
QB_HD void big_read(const C* tile, const DevOp& op, const C* m, int T, uint32_t swz_on, uint32_t task, BigAcc<C>& a) {
  const int k = (int)op.k, D = 1 << k, lt = k - 3;
  const uint32_t ntasks = (1u << (T - (int)op.nins)) << lt;
  a.valid = task < ntasks;
  if (!a.valid) return;
  a.sub = task & ((1u << lt) - 1);
  a.t0 = expand_mask(task >> lt, op.ins_mask) | op.tl_cmask;
#pragma unroll
  for (int r = 0; r < 8; ++r) a.acc[r] = cmake<C>(0, 0);
  for (int j = 0; j < D; ++j) {
    const C x = tile[swz<C>(a.t0 | big_offset(op, k, j), swz_on)];
#pragma unroll
    for (int r = 0; r < 8; ++r) cfma(a.acc[r], m[(a.sub * 8 + r) * D + j], x);
  }
}
template <typename C> QB_HD void big_write(C* tile, const DevOp& op, uint32_t swz_on, const BigAcc<C>& a) {
  if (!a.valid) return;
  const int k = (int)op.k;
#pragma unroll
  for (int r = 0; r < 8; ++r) tile[swz<C>(a.t0 | big_offset(op, k, (int)(a.sub * 8 + r)), swz_on)] = a.acc[r];
}

// ---- per-tile set-up of one slot (written into the calling team's TileSlot) ---------------------------------
template <typename C> QB_HD void micro_prephase(const MicroOp& mo, const char* blob, uint64_t base, int T, TileSlot& ts) {
  const int R = (int)mo.R, gbits = T - R;
  ts.active = (base & mo.ext_cmask) == mo.ext_cmask ? 1u : 0u;
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
    *reinterpret_cast<C*>(ts.ext) = s;
  } else if (mo.type == MU_SIGNS) {
    // parities of the bits outside the tile: bit r = over the partners of register bit r, bit 4 = over the lone terms
    uint32_t aux = 0;
    for (int e = 0; e < 5; ++e) {
      uint64_t x = base & mo.ext_mask[e];
      x ^= x >> 32;
      aux |= (popc32((uint32_t)x) & 1u) << e;
    }
    ts.aux = aux;
  } else if (mo.type == MU_DIAGK) {
    const int k = (int)mo.k;
    uint32_t aux = 0;
    for (int i = 0; i < k; ++i)
      if (mo.ext_mask[i] && (base & mo.ext_mask[i])) aux |= 1u << (k - 1 - i);
    ts.aux = aux;
  }
}

}  // namespace qb
