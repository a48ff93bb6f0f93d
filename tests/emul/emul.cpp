// TEST INFRASTRUCTURE ONLY.  CPU emulation of the sweep kernel's data path: it compiles the PRODUCT's own
// planner (qb_planner.hpp), canonicaliser (qb_canon.hpp) and shared-memory passes (qb_passes.cuh, nct = 1)
// with g++ and walks tiles exactly as qb_sweep.cuh does (gather runs -> per-op prephase -> passes ->
// scatter).  It lets the CPU test-suite check index math, fan tables and program serialisation without a
// GPU.  It is never loaded by qibo_b200/ -- the product has no CPU path.
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../qibo_b200/csrc/qb_passes.cuh"
#include "../../qibo_b200/csrc/qb_permute.cuh"
#include "../../qibo_b200/csrc/qb_families.hpp"

using namespace qb;

constexpr uint32_t EMUL_THREADS = 256;  // the kernel's compute-thread count: same group -> thread mapping

// SO: the sweep would be launched on the stage-only kernel instantiation (sweep_kernel<C, true>): REGTILE passes run
// through run_pass<..., true>, which has no handler switch, and other pass kinds are not executed at all -- so a sweep
// the planner flags stage_only by mistake gives a wrong state here, as it would on the GPU.
template <typename C, bool SO>
static void run_sweep(C* state, const char* blob, C* dst = nullptr) {
  const SweepHeader& hdr = *reinterpret_cast<const SweepHeader*>(blob);
  const int T = (int)hdr.T, L = (int)hdr.L;
  const uint32_t nruns = 1u << (T - L), run = 1u << L;
  const PassHeader* passes = reinterpret_cast<const PassHeader*>(blob + hdr.passes_offset);
  const uint32_t* slot_table = reinterpret_cast<const uint32_t*>(blob + hdr.slots_offset);
  std::vector<C> tile(size_t(1) << T), tile2(size_t(1) << T);  // tile2: where a permuting sweep's last pass writes (the
                                                              // kernel writes in place behind a team barrier)
  std::vector<TileSlot> ts(hdr.nslots + 1);
  const uint32_t swz_on = hdr.swizzle ? 7u : 0u;
  const int GPT = sizeof(C) == 16 ? 1 : 2;  // both group counts the kernel variants use are exercised
  for (uint64_t t = 0; t < hdr.ntiles; ++t) {
    const uint64_t base = deposit(t, hdr.other_mask);
    for (uint32_t r = 0; r < nruns; ++r) {
      const uint64_t off = deposit(uint64_t(r) << L, hdr.tile_mask);
      for (uint32_t e = 0; e < run; ++e) tile[swz<C>((r << L) + e, swz_on)] = state[base + off + e];  // TMA swizzle when planned
    }
    for (uint32_t sl = 0; sl < hdr.nslots; ++sl) {
      const uint32_t so = slot_table[sl];
      if (so & 0x80000000u) {
        const DevOp& bop = *reinterpret_cast<const DevOp*>(blob + (so & 0x7fffffffu));
        ts[sl].active = (base & bop.ext_cmask) == bop.ext_cmask ? 1u : 0u;
      } else {
        micro_prephase<C>(*reinterpret_cast<const MicroOp*>(blob + so), blob, base, T, ts[sl]);
      }
    }
    auto regtile = [&](const PassHeader& ph, uint32_t ctid) {
      if (GPT == 1) {
        switch (ph.R) {
          case 1: run_pass<C, 1, 1, SO>(tile.data(), blob, ts.data(), ph, T, swz_on, ctid, EMUL_THREADS, tile2.data()); break;
          case 2: run_pass<C, 2, 1, SO>(tile.data(), blob, ts.data(), ph, T, swz_on, ctid, EMUL_THREADS, tile2.data()); break;
          case 3: run_pass<C, 3, 1, SO>(tile.data(), blob, ts.data(), ph, T, swz_on, ctid, EMUL_THREADS, tile2.data()); break;
          default: run_pass<C, 4, 1, SO>(tile.data(), blob, ts.data(), ph, T, swz_on, ctid, EMUL_THREADS, tile2.data()); break;
        }
      } else {
        switch (ph.R) {
          case 1: run_pass<C, 1, 2, SO>(tile.data(), blob, ts.data(), ph, T, swz_on, ctid, EMUL_THREADS, tile2.data()); break;
          case 2: run_pass<C, 2, 2, SO>(tile.data(), blob, ts.data(), ph, T, swz_on, ctid, EMUL_THREADS, tile2.data()); break;
          case 3: run_pass<C, 3, 2, SO>(tile.data(), blob, ts.data(), ph, T, swz_on, ctid, EMUL_THREADS, tile2.data()); break;
          default: run_pass<C, 4, 2, SO>(tile.data(), blob, ts.data(), ph, T, swz_on, ctid, EMUL_THREADS, tile2.data()); break;
        }
      }
    };
    if (hdr.warp_private) {
      // the kernel separates these passes by __syncwarp only: a warp may run all of them before another warp starts.
      // Emulate exactly that order, so that a pass that needed another warp's output would give a wrong result here.
      for (uint32_t w = 0; w < EMUL_THREADS / 32; ++w)
        for (uint32_t pi = 0; pi < hdr.npasses; ++pi)
          for (uint32_t lane = 0; lane < 32; ++lane) regtile(passes[pi], w * 32 + lane);
    } else
    for (uint32_t pi = 0; pi < hdr.npasses; ++pi) {
      const PassHeader& ph = passes[pi];
      if (ph.kind == PASS_REGTILE) {
        for (uint32_t ctid = 0; ctid < EMUL_THREADS; ++ctid) regtile(ph, ctid);
      } else if (!SO) {
        const DevOp& op = *reinterpret_cast<const DevOp*>(blob + ph.offset);
        if (op.slot != MU_NO_SLOT && !ts[op.slot].active) continue;
        const C* payload = reinterpret_cast<const C*>(blob + op.payload);  // the emulator keeps the whole blob in one buffer
        const uint32_t ntasks = (1u << (T - (int)op.nins)) << (op.k - 3);
        std::vector<BigAcc<C>> accs(ntasks);
        for (uint32_t task = 0; task < ntasks; ++task) big_read<C>(tile.data(), op, payload, T, swz_on, task, accs[task]);
        for (uint32_t task = 0; task < ntasks; ++task) big_write<C>(tile.data(), op, swz_on, accs[task]);
      }
    }
    if (hdr.permuted) {
      // the storer: destination tile base = every bit of the source base moved to its destination; the tile leaves in
      // the destination layout (tensor-map order = ascending destination bits, swizzled when planned so)
      uint64_t dbase = 0;
      for (int b = 0; b < 64; ++b)
        if ((hdr.other_mask >> b) & 1) dbase |= ((base >> b) & uint64_t(1)) << hdr.dst_bit[b];
      const uint32_t dswz_on = hdr.dswizzle ? 7u : 0u;
      for (uint32_t d = 0; d < (1u << T); ++d) dst[dbase + deposit(d, hdr.dtile_mask)] = tile2[swz<C>(d, dswz_on)];
      continue;
    }
    for (uint32_t r = 0; r < nruns; ++r) {
      const uint64_t off = deposit(uint64_t(r) << L, hdr.tile_mask);
      for (uint32_t e = 0; e < run; ++e) state[base + off + e] = tile[swz<C>((r << L) + e, swz_on)];
    }
  }
}

static std::string g_err;

extern "C" const char* emul_last_error() { return g_err.c_str(); }

extern "C" int emul_apply_program(void* state, int nqubits, int dtype, const qb_op* ops, int nops, int flags,
                                  qb_program_stats* stats) {
  std::vector<CanonOp> canon;
  for (int i = 0; i < nops; ++i) {
    CanonOp c;
    if (!canonicalize(nqubits, ops[i].data, ops[i].is_diagonal != 0, ops[i].ntargets, ops[i].targets, ops[i].ncontrols,
                      ops[i].controls, c, g_err))
      return QB_ERR_INVALID;
    canon.push_back(c);
  }
  Plan plan;
  if (!plan_program(nqubits, dtype, canon, (flags & QB_PROGRAM_NO_FUSE) != 0, plan, g_err)) return QB_ERR_UNSUPPORTED;
  if (stats) fill_stats(plan, nqubits, dtype, nops, stats);
  for (auto& sd : plan.sweeps) {
    char* blob = plan.blob.data() + sd.blob_offset;
    const bool so = sd.stage_only != 0 && !env_int("QB_NO_STAGE_KERNEL", 0);  // as launch_sweep decides
    if (dtype == QB_C128) so ? run_sweep<double2, true>((double2*)state, blob) : run_sweep<double2, false>((double2*)state, blob);
    else so ? run_sweep<float2, true>((float2*)state, blob) : run_sweep<float2, false>((float2*)state, blob);
  }
  return QB_OK;
}

static void run_plan(void* state, int dtype, Plan& plan, void* dst = nullptr) {
  for (auto& sd : plan.sweeps) {
    char* blob = plan.blob.data() + sd.blob_offset;
    const bool so = sd.stage_only != 0 && !env_int("QB_NO_STAGE_KERNEL", 0);
    void* d = sd.permuted ? dst : nullptr;
    if (dtype == QB_C128) so ? run_sweep<double2, true>((double2*)state, blob, (double2*)d) : run_sweep<double2, false>((double2*)state, blob, (double2*)d);
    else so ? run_sweep<float2, true>((float2*)state, blob, (float2*)d) : run_sweep<float2, false>((float2*)state, blob, (float2*)d);
  }
}

// qb_apply_program_permuted in miniature: the ops, then dst[.. qubit dest_of_qubit[q] ..] = state[.. qubit q ..].  *fused = 1
// when the permutation rode on the last sweep (else the plain plan ran and the permutation is done here index by index).
extern "C" int emul_apply_program_permuted(void* state, void* dst, int nqubits, int dtype, const qb_op* ops, int nops,
                                           const int* dest_of_qubit, int flags, qb_program_stats* stats, int* fused, int replay) {
  std::vector<CanonOp> canon;
  for (int i = 0; i < nops; ++i) {
    CanonOp c;
    if (!canonicalize(nqubits, ops[i].data, ops[i].is_diagonal != 0, ops[i].ntargets, ops[i].targets, ops[i].ncontrols,
                      ops[i].controls, c, g_err))
      return QB_ERR_INVALID;
    canon.push_back(c);
  }
  PermSpec perm;
  for (int q = 0; q < nqubits; ++q) perm.pi[nqubits - 1 - q] = nqubits - 1 - dest_of_qubit[q];
  Plan plan;
  const bool no_fuse = (flags & QB_PROGRAM_NO_FUSE) != 0;
  if (!plan_program(nqubits, dtype, canon, no_fuse, plan, g_err, nullptr, &perm)) return QB_ERR_UNSUPPORTED;
  if (replay) {  // emit the same program again on the kept schedule (qb_program_set_params with unchanged values)
    Plan again;
    if (!plan_program(nqubits, dtype, canon, no_fuse, again, g_err, &plan, &perm)) return QB_ERR_UNSUPPORTED;
    if (again.perm_fused != plan.perm_fused || again.sweeps.size() != plan.sweeps.size()) {
      g_err = "replay changed the plan";
      return QB_ERR_UNSUPPORTED;
    }
    plan = again;
  }
  if (stats) fill_stats(plan, nqubits, dtype, nops, stats);
  *fused = plan.perm_fused;
  run_plan(state, dtype, plan, dst);
  if (!plan.perm_fused) {
    const size_t esz = dtype == QB_C128 ? 16 : 8;
    for (uint64_t x = 0; x < (uint64_t(1) << nqubits); ++x) {
      uint64_t y = 0;
      for (int b = 0; b < nqubits; ++b) y |= ((x >> b) & uint64_t(1)) << perm.pi[b];
      memcpy((char*)dst + y * esz, (const char*)state + x * esz, esz);
    }
  }
  return QB_OK;
}

// qb_program_set_params in miniature: plan `ops_old`, then emit `ops_new` (same gate structure, new numbers) on the OLD
// schedule and run it; *replayed = 0 when the structure differed and the new ops were planned from scratch.
extern "C" int emul_apply_program_replay(void* state, int nqubits, int dtype, const qb_op* ops_old, const qb_op* ops_new, int nops,
                                         int flags, int* replayed) {
  std::vector<CanonOp> a, b;
  for (int i = 0; i < nops; ++i) {
    CanonOp c, d;
    if (!canonicalize(nqubits, ops_old[i].data, ops_old[i].is_diagonal != 0, ops_old[i].ntargets, ops_old[i].targets,
                      ops_old[i].ncontrols, ops_old[i].controls, c, g_err) ||
        !canonicalize(nqubits, ops_new[i].data, ops_new[i].is_diagonal != 0, ops_new[i].ntargets, ops_new[i].targets,
                      ops_new[i].ncontrols, ops_new[i].controls, d, g_err))
      return QB_ERR_INVALID;
    a.push_back(c);
    b.push_back(d);
  }
  const bool no_fuse = (flags & QB_PROGRAM_NO_FUSE) != 0;
  Plan old_plan, fresh;
  if (!plan_program(nqubits, dtype, a, no_fuse, old_plan, g_err)) return QB_ERR_UNSUPPORTED;
  *replayed = 1;
  if (!plan_program(nqubits, dtype, b, no_fuse, fresh, g_err, &old_plan)) {
    *replayed = 0;
    if (!plan_program(nqubits, dtype, b, no_fuse, fresh, g_err)) return QB_ERR_UNSUPPORTED;
  } else if (fresh.sweeps.size() != old_plan.sweeps.size()) {
    g_err = "replay changed the number of sweeps";
    return QB_ERR_UNSUPPORTED;
  }
  run_plan(state, dtype, fresh);
  return QB_OK;
}

extern "C" int emul_family_matrix(int family, const double* theta, int ntargets, int is_diagonal, double* out, int* count) {
  std::vector<double> m;
  if (!family_matrix(family, theta, ntargets, is_diagonal != 0, m)) return QB_ERR_INVALID;
  for (size_t i = 0; i < m.size(); ++i) out[i] = m[i];
  *count = (int)m.size();
  return QB_OK;
}

// canonical form of one gate, for structural tests: kind, #targets, #controls
extern "C" int emul_canon_kind(int nqubits, const double* data, int is_diag, int nt, const int* targets, int nc,
                               const int* controls, int* kind, int* ntargets, int* ncontrols) {
  CanonOp c;
  if (!canonicalize(nqubits, data, is_diag != 0, nt, targets, nc, controls, c, g_err)) return QB_ERR_INVALID;
  *kind = c.kind;
  *ntargets = (int)c.tpos.size();
  *ncontrols = (int)c.cpos.size();
  return QB_OK;
}

// K8 data path: the kernel's per-thread loops (thread id = 8 low tile bits, masked increment for the rest)
template <typename C> static void run_permute(const C* src, C* dst, const PermParams& p) {
  const uint32_t tsize = 1u << p.tbits;
  const uint32_t iters = tsize > PERM_THREADS ? tsize / PERM_THREADS : 1;
  const uint64_t other = ~p.smask & ((uint64_t(1) << p.n) - 1);
  uint64_t s_lo, s_hi, d_lo, d_hi;
  split_mask8(p.smask, s_lo, s_hi);
  split_mask8(p.dmask, d_lo, d_hi);
  std::vector<C> tile(tsize);
  for (uint64_t t = 0; t < p.ntiles; ++t) {
    const uint64_t sbase = deposit(t, other), dbase = permute_base(sbase, p);
    for (uint32_t tid = 0; tid < (uint32_t)PERM_THREADS && tid < tsize; ++tid) {
      uint64_t hi = 0;
      for (uint32_t k = 0; k < iters; ++k) {
        tile[tid + k * PERM_THREADS] = src[sbase | deposit(tid, s_lo) | hi];
        hi = ((hi | ~s_hi) + 1) & s_hi;
      }
    }
    for (uint32_t tid = 0; tid < (uint32_t)PERM_THREADS && tid < tsize; ++tid) {
      uint64_t hi = 0;
      const uint32_t e_lo = perm_e_of_f(tid & (tsize - 1), p);
      for (uint32_t k = 0; k < iters; ++k) {
        const uint32_t e = e_lo | perm_e_of_f((k * PERM_THREADS) & (tsize - 1), p);
        dst[dbase | deposit(tid, d_lo) | hi] = tile[e];
        hi = ((hi | ~d_hi) + 1) & d_hi;
      }
    }
  }
}

extern "C" int emul_permute_qubits(const void* src, void* dst, int nqubits, int dtype, const int* dest_of_qubit) {
  int pi[64];
  for (int q = 0; q < nqubits; ++q) pi[nqubits - 1 - q] = nqubits - 1 - dest_of_qubit[q];
  PermParams p;
  memset(&p, 0, sizeof(p));
  perm_setup(nqubits, 6, 6, pi, p);
  if (dtype == QB_C128) run_permute<double2>((const double2*)src, (double2*)dst, p);
  else run_permute<float2>((const float2*)src, (float2*)dst, p);
  return QB_OK;
}
