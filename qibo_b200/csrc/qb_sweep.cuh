// qibo_b200 K2: the multi-gate sweep kernel (sm_100a).
//
// One persistent CTA per SM.  Warp 0 (LOADER) and warp 1 (STORER) move tiles between HBM and shared
// memory with bulk-async copies (cp.async.bulk, the TMA engine; SASS UBLKCP) -- one copy per contiguous
// run of the tile -- tracked by mbarriers (loads: complete_tx; stores: bulk groups).  Warps 2..9 are
// COMPUTE warps: they apply every gate of the sweep to the resident tile in shared memory.  Three 64 KiB tiles rotate
// through the states LOADING -> COMPUTING -> STORING, so HBM reads, shared-memory math and HBM writes of
// three consecutive tiles overlap.  HBM traffic per sweep: 2 * B * 2^n regardless of how many gates ride.
//
// The device program (SweepHeader + DevOp[] + payloads, qb_planner.hpp) is copied to shared memory once.
#pragma once
#include <cuda.h>  // CUtensorMap (types only: the encoder is fetched through cudaGetDriverEntryPoint, libcuda is not linked)
#include <cuda_runtime.h>

#include "qb_common.cuh"
#include "qb_gate_kernels.cuh"
#include "qb_planner.hpp"
#include "qb_passes.cuh"

namespace qb {

constexpr int SW_NBUF = 3;
constexpr int SW_TILE_BYTES = 1 << SWEEP_TILE_BYTES_LOG2;
// TEAMS of 8 compute warps: each team owns one resident tile at a time (one group of 2^4 amplitudes per thread for
// complex128, two for complex64), so with two teams two tiles are in their compute phase while the third buffer is
// being stored / reloaded, and one team's barriers, shared-memory bursts and set-up phases hide behind the other's
// math.  Per dtype (measured on B200, scripts/variant_bench.py): complex64 runs two teams; complex128 needs 64 data
// registers per thread and is faster with one team and no spills (168 registers) than with two at 112.
#ifndef QB_TEAMS128
#define QB_TEAMS128 1
#endif
#ifndef QB_TEAMS64
#define QB_TEAMS64 2
#endif
// Register reallocation (setmaxnreg, sm_90a+): the copy warps live in a warpgroup of their own (warps 2-3 of it idle)
// and hand their registers to the compute warpgroups right after start-up; 0 disables it.  Budget: the compute
// threads can only take what the copy warpgroup released, (launch_regs - 24) * 128 registers.
#ifndef QB_REGS128
#define QB_REGS128 0
#endif
#ifndef QB_REGS64
#define QB_REGS64 112
#endif
#ifndef QB_TEAM_THREADS
#define QB_TEAM_THREADS 256
#endif
constexpr int SW_TEAM_THREADS = QB_TEAM_THREADS;
constexpr int SW_MAX_TEAMS = 2;
// SO ("stage only"): a second instantiation for sweeps whose passes are all straight-line stage passes (every QFT
// sweep).  Without the 54-handler switch the compute code fits 112 registers, so complex128 can run two teams there.
#ifndef QB_TEAMS128_SO
#define QB_TEAMS128_SO 2
#endif
#ifndef QB_REGS128_SO
#define QB_REGS128_SO 112
#endif
template <typename C, bool SO> struct SweepCfg;
template <> struct SweepCfg<double2, false> {
  static constexpr int TEAMS = QB_TEAMS128, COMPUTE_REGS = QB_REGS128, RB = QB_R128;
};
template <> struct SweepCfg<double2, true> {
  static constexpr int TEAMS = QB_TEAMS128_SO, COMPUTE_REGS = QB_REGS128_SO, RB = QB_R128;
};
template <bool SO> struct SweepCfg<float2, SO> {
  static constexpr int TEAMS = QB_TEAMS64, COMPUTE_REGS = QB_REGS64, RB = QB_R64;
};
template <typename C, bool SO> constexpr int sw_copy_threads() { return SweepCfg<C, SO>::COMPUTE_REGS ? 128 : 64; }  // warp 0: loader, warp 1: storer
template <typename C, bool SO> constexpr int sw_threads() { return SweepCfg<C, SO>::TEAMS * SW_TEAM_THREADS + sw_copy_threads<C, SO>(); }
constexpr int SW_MAX_RUNS = 256;
constexpr int SW_FIXED_BYTES = SW_NBUF * SW_TILE_BYTES + SW_MAX_RUNS * 8 + SW_NBUF * SWEEP_MAX_SLOTS * (int)sizeof(TileSlot) + 128;
constexpr int SW_BLOB_REGION = ((226 * 1024 - SW_FIXED_BYTES) / 16) * 16;
constexpr int SW_SMEM_BYTES = SW_FIXED_BYTES + SW_BLOB_REGION;
static_assert(SW_SMEM_BYTES <= 227 * 1024, "sweep kernel shared memory exceeds 227 KB");
static_assert(SW_BLOB_REGION >= SWEEP_BLOB_MAX + 1024, "shared-memory program region is smaller than the planner's limit");

// Tile <-> HBM as ONE tensor-map copy per tile (cp.async.bulk.tensor, SASS UTMALDG / UTMASTG) when the tile's bits
// form <= 5 alternating runs of tile / non-tile bits (QFT and layered circuits do); else one bulk copy per
// contiguous run.  ncu: the per-run form costs ~20 issue slots per 512-byte copy on the copy warps -- 18 % of all
// instructions of a QFT sweep -- and caps a sweep at the copy-issue rate when a single lane issues them.
struct TmaDesc {
  CUtensorMap map;     // rank 5, 8-byte elements, dim d = d-th run of state bits (a complex128 amplitude = 2 elements)
  int32_t shift[5];    // coordinate of dim d for a tile at state index `base`: (base >> shift[d]) & mask[d]
  uint32_t mask[5];    // 0 for dims spanned by the tile
  int32_t enabled;
  int32_t pad[5];
};
static_assert(sizeof(TmaDesc) % 64 == 0, "CUtensorMap must stay 64-byte aligned inside the kernel parameter");

// ---- PTX wrappers -------------------------------------------------------------------------------------
QB_D uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
QB_D void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
QB_D void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
QB_D void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
QB_D bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// bounded wait: a protocol bug traps (error returned to the host) instead of hanging the GPU
QB_D void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 20000000000LL) __trap();
  }
}
// for the copy warps: they wait for whole tile periods, so let the hardware suspend the warp (time hint) instead of
// spinning through issue slots of the scheduler they share with compute warps
QB_D bool mbar_try_wait_hint(uint64_t* bar, uint32_t parity, uint32_t ns) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(ns)
      : "memory");
  return ok != 0;
}
QB_D void mbar_wait_sleep(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  // ncu (round 2, profiles/r2a_*): with the suspend-time hint alone the copy warps still executed 12 % of all the
  // instructions of a sweep -- NANOSLEEP.SYNCS wakes on every mbarrier event of the CTA, and the wake-up loop (try_wait,
  // clock64, compare) competed with the compute warps of its scheduler.  A plain timed sleep between polls costs at most
  // a fraction of a microsecond of latency per tile (a tile period is ~10-20 us and three buffers are in flight).
  long long t0 = 0;
  for (uint32_t spins = 0;; ++spins) {
    __nanosleep(400);
    if (mbar_try_wait_hint(bar, parity, 2000u)) return;
    if ((spins & 0xFFFu) == 0) {
      const long long now = clock64();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 20000000000LL) __trap();
    }
  }
}
// state index of tile t = deposit(t, other_mask); stepping t by a constant is a masked add (carries ripple through the
// bits outside the mask), so the software pdep runs once per role instead of once per tile per warp
QB_D uint64_t masked_add(uint64_t x, uint64_t d, uint64_t mask) { return ((x | ~mask) + d) & mask; }
QB_D void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
QB_D void bulk_s2g(void* gdst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(smem_src)), "r"(bytes)
               : "memory");
}
QB_D void tma_load_5d(void* smem_dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, int c4, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4), "r"(smem_u32(bar))
               : "memory");
}
QB_D void tma_store_5d(const CUtensorMap* map, int c0, int c1, int c2, int c3, int c4, const void* smem_src) {
  asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.tile.bulk_group [%0, {%1, %2, %3, %4, %5}], [%6];" ::"l"(
                   reinterpret_cast<uint64_t>(map)),
               "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4), "r"(smem_u32(smem_src))
               : "memory");
}
QB_D void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> QB_D void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
QB_D void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
QB_D void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
QB_D void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// literal barrier ids: with a register operand ptxas reserves all 16 hardware barriers
QB_D void team_bar(int team) {
  if (team == 0) asm volatile("bar.sync 1, %0;" ::"n"(SW_TEAM_THREADS) : "memory");
  else asm volatile("bar.sync 2, %0;" ::"n"(SW_TEAM_THREADS) : "memory");
}

// ---- the kernel -----------------------------------------------------------------------------------------
template <typename C, bool SO>
__global__ void __launch_bounds__(sw_threads<C, SO>(), 1) sweep_kernel(C* __restrict__ state, const char* __restrict__ prog, const __grid_constant__ TmaDesc tma, const __grid_constant__ TmaDesc tma_dst) {
  extern __shared__ __align__(1024) unsigned char smem[];
  C* tiles = reinterpret_cast<C*>(smem);
  char* blob = reinterpret_cast<char*>(smem + SW_NBUF * SW_TILE_BYTES);
  uint64_t* run_off = reinterpret_cast<uint64_t*>(blob + SW_BLOB_REGION);
  TileSlot* tslots = reinterpret_cast<TileSlot*>(run_off + SW_MAX_RUNS);
  constexpr int SW_TEAMS = SweepCfg<C, SO>::TEAMS, SW_COPY_THREADS = sw_copy_threads<C, SO>(), SW_THREADS = sw_threads<C, SO>();
  // One `full` barrier per (buffer, team) pair: tile i lands in buffer i % NBUF and signals full[i % NFULL].  With one
  // barrier per buffer a team would revisit it every 2 * NBUF tiles with the SAME parity while the phase in between
  // belongs to the other team: if that load is still in flight when a fast team comes round (one-stage sweeps: the
  // compute phase is shorter than the jitter of a 64 KiB load), a parity wait is satisfied by the stale phase and both
  // teams end up in one buffer.  Here consecutive phases of a barrier are all waited for by the same team.
  constexpr int SW_NFULL = SW_NBUF * SW_TEAMS;
  static_assert((2 * SW_NBUF + SW_NBUF * SW_MAX_TEAMS) * 8 <= 128, "mbarrier area");
  uint64_t* full = reinterpret_cast<uint64_t*>(tslots + SW_NBUF * SWEEP_MAX_SLOTS);
  uint64_t* done = full + SW_NBUF * SW_MAX_TEAMS;
  uint64_t* freeb = done + SW_NBUF;

  const int tid = threadIdx.x;
  {
    const uint32_t nbytes = reinterpret_cast<const SweepHeader*>(prog)->blob_bytes;
    for (uint32_t i = tid * 16; i < nbytes; i += SW_THREADS * 16)
      *reinterpret_cast<int4*>(blob + i) = __ldg(reinterpret_cast<const int4*>(prog + i));
  }
  __syncthreads();
  const SweepHeader& hdr = *reinterpret_cast<const SweepHeader*>(blob);
  const int T = (int)hdr.T, L = (int)hdr.L;
  const uint32_t nruns = 1u << (T - L);
  const uint32_t run_bytes = (uint32_t)sizeof(C) << L;
  const uint32_t tile_bytes = (uint32_t)sizeof(C) << T;
  const uint32_t tile_elems = 1u << T;
  for (uint32_t r = tid; r < nruns; r += SW_THREADS) run_off[r] = deposit(uint64_t(r) << L, hdr.tile_mask) * sizeof(C);
  if (tid == 0) {
    for (int b = 0; b < SW_NFULL; ++b) mbar_init(&full[b], 2);  // the tile's bytes (expect_tx) + its slot states (loader)
    for (int b = 0; b < SW_NBUF; ++b) {
      mbar_init(&done[b], 1);
      mbar_init(&freeb[b], 1);
    }
    fence_mbar_init();
    fence_proxy_async();
  }
  __syncthreads();

  const uint32_t* slot_table = reinterpret_cast<const uint32_t*>(blob + hdr.slots_offset);
  const int nslots = (int)hdr.nslots;
  const uint64_t ntiles = hdr.ntiles;
  const uint64_t my_n = blockIdx.x < ntiles ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  if constexpr (SweepCfg<C, SO>::COMPUTE_REGS != 0) {
    if (tid < SW_COPY_THREADS) asm volatile("setmaxnreg.dec.sync.aligned.u32 24;");
    else asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(SweepCfg<C, SO>::COMPUTE_REGS));
  }

  if (tid < 32) {
    // ================= loader warp: HBM -> shared =================
    const int lane = tid;
    uint64_t base = deposit(blockIdx.x, hdr.other_mask);
    const uint64_t dstep = deposit(gridDim.x, hdr.other_mask);
    for (uint64_t i = 0; i < my_n; ++i, base = masked_add(base, dstep, hdr.other_mask)) {
      const int b = (int)(i % SW_NBUF);
      uint64_t* fullb = &full[i % SW_NFULL];
      // buffer b was last used by tile i-3: wait until the storer has drained it
      if (i >= SW_NBUF) mbar_wait_sleep(&freeb[b], (uint32_t)(((i / SW_NBUF) - 1) & 1));
      // The copy is issued FIRST: the buffer has just come free and every cycle until the load is under way is a bubble
      // in the buffer's load -> compute -> store cycle (three buffers per SM: the sweep's throughput).  The per-tile set-up
      // of the ops that depend on bits outside the tile (control predicates, fan factors) -- done by the otherwise idle
      // lanes of this warp, off the compute warps' critical path -- then runs while the bytes are in flight.  `full` counts
      // two arrivals: the copy's expect_tx and, after the set-up, the release of the slot states (-> the compute
      // threads' acquire when they wait on `full`).
      char* sbase = reinterpret_cast<char*>(tiles + (size_t)b * tile_elems);
      const bool synth = tma.pad[1] != 0;
      if (synth) {
        // the input is |0...0> and has not been written anywhere (QB_PROGRAM_INPUT_ZERO): the tile is made here instead of
        // being read -- the first sweep of an execution is write-only and the 2^n-amplitude fill before it disappears
        int4* zt = reinterpret_cast<int4*>(sbase);
        for (uint32_t k = lane; k < tile_bytes / 16; k += 32) zt[k] = make_int4(0, 0, 0, 0);
        if (lane == 0 && base == 0) {
          C one;
          one.x = 1;
          one.y = 0;
          tiles[(size_t)b * tile_elems] = one;  // tile-local index 0 sits at offset 0 in either layout
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(fullb);
      } else if (tma.enabled) {
        if (lane == 0) {
          mbar_expect_tx(fullb, tile_bytes);
          tma_load_5d(sbase, &tma.map, (int)((base >> tma.shift[0]) & tma.mask[0]), (int)((base >> tma.shift[1]) & tma.mask[1]),
                      (int)((base >> tma.shift[2]) & tma.mask[2]), (int)((base >> tma.shift[3]) & tma.mask[3]),
                      (int)((base >> tma.shift[4]) & tma.mask[4]), fullb);
        }
      } else {
        if (lane == 0) mbar_expect_tx(fullb, tile_bytes);
        __syncwarp();
        const char* gbase = reinterpret_cast<const char*>(state + base);
        for (uint32_t r = lane; r < nruns; r += 32) bulk_g2s(sbase + r * run_bytes, gbase + run_off[r], run_bytes, fullb);
      }
      {
        TileSlot* ts = tslots + b * SWEEP_MAX_SLOTS;
        for (int sl = lane; sl < nslots; sl += 32) {
          const uint32_t so = slot_table[sl];
          if (so & 0x80000000u) {
            const DevOp& bop = *reinterpret_cast<const DevOp*>(blob + (so & 0x7fffffffu));
            ts[sl].active = (base & bop.ext_cmask) == bop.ext_cmask ? 1u : 0u;
          } else {
            micro_prephase<C>(*reinterpret_cast<const MicroOp*>(blob + so), blob, base, T, ts[sl]);
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(fullb);
      }
    }
  } else if (tid < 64) {
    // ================= storer warp: shared -> HBM =================
    const int lane = tid - 32;
    uint64_t base = deposit(blockIdx.x, hdr.other_mask);
    const uint64_t dstep = deposit(gridDim.x, hdr.other_mask);
    for (uint64_t j = 0; j < my_n; ++j, base = masked_add(base, dstep, hdr.other_mask)) {
      const int b = (int)(j % SW_NBUF);
      mbar_wait_sleep(&done[b], (uint32_t)((j / SW_NBUF) & 1));
      const char* sbase = reinterpret_cast<const char*>(tiles + (size_t)b * tile_elems);
      if (hdr.permuted) {
        // permuting sweep: the compute team left the tile in the DESTINATION layout; it goes out through the destination
        // tensor map (another buffer) at the tile base with every bit moved to its destination
        uint64_t contrib = 0;
        for (int bit = lane; bit < 64; bit += 32)
          if ((hdr.other_mask >> bit) & 1) contrib |= ((base >> bit) & uint64_t(1)) << hdr.dst_bit[bit];
        const uint64_t dbase = (uint64_t)__reduce_or_sync(0xffffffffu, (uint32_t)contrib) |
                               ((uint64_t)__reduce_or_sync(0xffffffffu, (uint32_t)(contrib >> 32)) << 32);
        if (lane == 0) {
          tma_store_5d(&tma_dst.map, (int)((dbase >> tma_dst.shift[0]) & tma_dst.mask[0]), (int)((dbase >> tma_dst.shift[1]) & tma_dst.mask[1]),
                       (int)((dbase >> tma_dst.shift[2]) & tma_dst.mask[2]), (int)((dbase >> tma_dst.shift[3]) & tma_dst.mask[3]),
                       (int)((dbase >> tma_dst.shift[4]) & tma_dst.mask[4]), sbase);
          bulk_commit();
          bulk_wait_read<0>();
          mbar_arrive(&freeb[b]);
        }
      } else if (tma.enabled) {
        if (lane == 0) {
          tma_store_5d(&tma.map, (int)((base >> tma.shift[0]) & tma.mask[0]), (int)((base >> tma.shift[1]) & tma.mask[1]),
                       (int)((base >> tma.shift[2]) & tma.mask[2]), (int)((base >> tma.shift[3]) & tma.mask[3]),
                       (int)((base >> tma.shift[4]) & tma.mask[4]), sbase);
          bulk_commit();
          bulk_wait_read<0>();
          mbar_arrive(&freeb[b]);
        }
      } else {
        char* gbase = reinterpret_cast<char*>(state + base);
        for (uint32_t r = lane; r < nruns; r += 32) bulk_s2g(gbase + run_off[r], sbase + r * run_bytes, run_bytes);
        bulk_commit();
        bulk_wait_read<0>();  // shared memory of this tile has been read out: the loader may refill it
        __syncwarp();
        if (lane == 0) mbar_arrive(&freeb[b]);
      }
    }
    bulk_wait_all();
  } else if (tid >= SW_COPY_THREADS) {
    // ================= compute teams =================
    const int team = (tid - SW_COPY_THREADS) / SW_TEAM_THREADS;
    const uint32_t ctid = (uint32_t)(tid - SW_COPY_THREADS) % SW_TEAM_THREADS;
    const PassHeader* passes = reinterpret_cast<const PassHeader*>(blob + hdr.passes_offset);
    const int npasses = (int)hdr.npasses;
    const uint32_t swz_on = hdr.swizzle ? 7u : 0u;
    const bool warp_private = hdr.warp_private != 0;
    static_assert(SW_TEAM_THREADS == SWEEP_TEAM_THREADS, "group tables are laid out for SWEEP_TEAM_THREADS");
    constexpr int RB = SweepCfg<C, SO>::RB;                                     // register bits of a REGTILE pass
    constexpr int TB = SWEEP_TILE_BYTES_LOG2 - (sizeof(C) == 16 ? 4 : 3);  // tile bits of a full tile
    constexpr int GPT = ((1 << (TB - RB)) + SW_TEAM_THREADS - 1) / SW_TEAM_THREADS;  // groups per thread
    for (uint64_t i = team; i < my_n; i += SW_TEAMS) {
      const int b = (int)(i % SW_NBUF);
      C* tile = tiles + (size_t)b * tile_elems;
      const TileSlot* ts = tslots + b * SWEEP_MAX_SLOTS;
      mbar_wait(&full[i % SW_NFULL], (uint32_t)((i / SW_NFULL) & 1));  // tile data (async proxy) + slot states (loader warp) are visible
      // (tma.pad[0]: profiling knob QB_SWEEP_SKIP_COMPUTE -- the tile passes through untouched, which times the data
      // movement of a sweep's tile shape alone)
      for (int pi = 0; pi < (tma.pad[0] ? 0 : npasses); ++pi) {
        const PassHeader& ph = passes[pi];
        if (pi) {
          if (warp_private) __syncwarp();  // the pass only re-reads what this warp wrote
          else team_bar(team);
        }
        if (ph.kind == PASS_REGTILE) {
          // R < RB only occurs for n < 4, which qb_apply_program routes to the K1 kernels
          run_pass<C, RB, GPT, SO>(tile, blob, ts, ph, T, swz_on, ctid, SW_TEAM_THREADS, (C*)nullptr, [team]() { team_bar(team); });
        } else if constexpr (!SO) {
          const DevOp& op = *reinterpret_cast<const DevOp*>(blob + ph.offset);
          if (op.slot == MU_NO_SLOT || ts[op.slot].active) {
            const C* payload = op.payload_global ? reinterpret_cast<const C*>(prog + op.payload) : reinterpret_cast<const C*>(blob + op.payload);
            const uint32_t ntasks = (1u << (T - (int)op.nins)) << (op.k - 3);
            for (uint32_t t = 0; t < ntasks; t += SW_TEAM_THREADS) {
              BigAcc<C> a;
              big_read<C>(tile, op, payload, T, swz_on, t + ctid, a);
              team_bar(team);
              big_write<C>(tile, op, swz_on, a);
              if (t + SW_TEAM_THREADS < ntasks) team_bar(team);
            }
          }
        }
      }
      fence_proxy_async();
      team_bar(team);
      if (ctid == 0) mbar_arrive(&done[b]);
    }
  }
}

inline int sweep_configure(const cudaDeviceProp& prop) {
  if ((int)prop.sharedMemPerBlockOptin < SW_SMEM_BYTES) return QB_ERR_UNSUPPORTED;
  if (cudaFuncSetAttribute(sweep_kernel<double2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SW_SMEM_BYTES) != cudaSuccess) return QB_ERR_CUDA;
  if (cudaFuncSetAttribute(sweep_kernel<double2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SW_SMEM_BYTES) != cudaSuccess) return QB_ERR_CUDA;
  if (cudaFuncSetAttribute(sweep_kernel<float2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SW_SMEM_BYTES) != cudaSuccess) return QB_ERR_CUDA;
  if (cudaFuncSetAttribute(sweep_kernel<float2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SW_SMEM_BYTES) != cudaSuccess) return QB_ERR_CUDA;
  return QB_OK;
}

// ---- host: tensor map of one sweep ---------------------------------------------------------------------------
typedef CUresult (*qb_encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                       const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                       CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline qb_encode_tiled_fn tma_encoder() {
  static qb_encode_tiled_fn fn = []() -> qb_encode_tiled_fn {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess || qres != cudaDriverEntryPointSuccess)
      return nullptr;
    return reinterpret_cast<qb_encode_tiled_fn>(p);
  }();
  return fn;
}

// One tensor dimension per run of tile bits plus one for all the other bits (tile_segments, qb_planner.hpp).  Returns
// false (per-run copies) when the tile needs more than 5 dimensions or the driver entry point is missing.
inline bool tma_describe(void* state, int nqubits, int dtype, uint64_t tile_mask, bool swizzle, TmaDesc& d) {
  memset(&d, 0, sizeof(d));
  if (env_int("QB_NO_TMA", 0)) return false;
  qb_encode_tiled_fn enc = tma_encoder();
  if (!enc) return false;
  const int epa = dtype == QB_C128 ? 2 : 1;  // 8-byte elements per amplitude
  std::vector<TileSeg> segs = tile_segments(nqubits, dtype, tile_mask, swizzle);
  if (segs.empty() || !segs[0].tile || segs.size() > 5) return false;
  cuuint64_t gdim[5], gstride[4];
  cuuint32_t box[5], estr[5];
  const cuuint64_t amp_bytes = 8 * (cuuint64_t)epa;
  for (int i = 0; i < 5; ++i) {
    estr[i] = 1;
    if (i < (int)segs.size()) {
      if (segs[i].len > 32) return false;
      gdim[i] = (cuuint64_t(1) << segs[i].len) * (i == 0 ? epa : 1);
      if (gdim[i] > (cuuint64_t(1) << 32)) return false;
      box[i] = segs[i].tile ? (cuuint32_t)gdim[i] : 1;
      d.shift[i] = segs[i].start;
      d.mask[i] = segs[i].tile ? 0u : (segs[i].len >= 32 ? 0xFFFFFFFFu : (uint32_t)((uint64_t(1) << segs[i].len) - 1));
      // stride of dimension i = the byte offset of its first bit (dimensions are ordered by ascending first bit)
      if (i > 0) {
        const cuuint64_t bytes = amp_bytes << segs[i].start;
        if (bytes >= (cuuint64_t(1) << 40)) return false;
        gstride[i - 1] = bytes;
      }
    } else {
      gdim[i] = 1;
      box[i] = 1;
      d.shift[i] = 0;
      d.mask[i] = 0;
      // unused trailing dimensions: extent 1, any legal stride (a multiple of the one below)
      if (i > 0) gstride[i - 1] = i >= 2 ? gstride[i - 2] : amp_bytes * (gdim[0] / epa);
    }
  }
  if (enc(&d.map, CU_TENSOR_MAP_DATA_TYPE_UINT64, 5, state, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
          swizzle ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
          CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return false;
  d.enabled = 1;
  return true;
}

// `dst`: the buffer a PERMUTING sweep (sd.permuted) writes to -- another buffer of the state's size; ignored otherwise.
// `input_zero`: the state is |0...0> and `state` holds nothing yet -- the loader makes the tiles instead of reading them.
inline int launch_sweep(cudaStream_t stream, int sm_count, void* state, int nqubits, int dtype, const SweepDesc& sd,
                        const char* prog_dev, void* dst = nullptr, bool input_zero = false) {
  uint64_t grid = sd.ntiles < (uint64_t)sm_count ? sd.ntiles : (uint64_t)sm_count;
  TmaDesc tma, tma_dst;
  if (!tma_describe(state, nqubits, dtype, sd.tile_mask, sd.swizzle != 0, tma) && sd.swizzle) return QB_ERR_UNSUPPORTED;  // planned for a swizzled tile
  if (sd.permuted) {
    if (!dst || dst == state) return QB_ERR_INVALID;
    if (!tma_describe(dst, nqubits, dtype, sd.dmask, sd.dswizzle != 0, tma_dst)) return QB_ERR_UNSUPPORTED;
  } else {
    tma_dst = tma;
  }
  const bool so = sd.stage_only != 0 && !env_int("QB_NO_STAGE_KERNEL", 0);
  tma.pad[0] = env_int("QB_SWEEP_SKIP_COMPUTE", 0);
  tma.pad[1] = input_zero ? 1 : 0;
  if (dtype == QB_C128) {
    if (so) sweep_kernel<double2, true><<<(unsigned)grid, sw_threads<double2, true>(), SW_SMEM_BYTES, stream>>>((double2*)state, prog_dev + sd.blob_offset, tma, tma_dst);
    else sweep_kernel<double2, false><<<(unsigned)grid, sw_threads<double2, false>(), SW_SMEM_BYTES, stream>>>((double2*)state, prog_dev + sd.blob_offset, tma, tma_dst);
  } else {
    if (so) sweep_kernel<float2, true><<<(unsigned)grid, sw_threads<float2, true>(), SW_SMEM_BYTES, stream>>>((float2*)state, prog_dev + sd.blob_offset, tma, tma_dst);
    else sweep_kernel<float2, false><<<(unsigned)grid, sw_threads<float2, false>(), SW_SMEM_BYTES, stream>>>((float2*)state, prog_dev + sd.blob_offset, tma, tma_dst);
  }
  return cudaPeekAtLastError() == cudaSuccess ? QB_OK : QB_ERR_CUDA;
}

// resources of the compiled kernel, for launch-failure messages
inline std::string sweep_resources(int dtype) {
  cudaFuncAttributes a;
  cudaError_t e = dtype == QB_C128 ? cudaFuncGetAttributes(&a, sweep_kernel<double2, false>) : cudaFuncGetAttributes(&a, sweep_kernel<float2, false>);
  if (e != cudaSuccess) return "(no attributes)";
  return "threads=" + std::to_string(dtype == QB_C128 ? sw_threads<double2, false>() : sw_threads<float2, false>()) + " regs=" + std::to_string(a.numRegs) + " maxThreadsPerBlock=" + std::to_string(a.maxThreadsPerBlock) +
         " static_smem=" + std::to_string(a.sharedSizeBytes) + " dyn_smem=" + std::to_string(SW_SMEM_BYTES) +
         " max_dyn_smem=" + std::to_string(a.maxDynamicSharedSizeBytes) + " local=" + std::to_string(a.localSizeBytes);
}

// ---- K7 helper: gather / scatter the half of the shard with bit `pos` == `bit` ----------------------------
template <typename C>
__global__ void __launch_bounds__(256) k7_half_copy(C* __restrict__ state, C* __restrict__ staging, uint64_t half, int pos, int bit,
                                                    int unpack) {
  const uint64_t stride = uint64_t(gridDim.x) * blockDim.x;
  for (uint64_t g = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; g < half; g += stride) {
    const uint64_t idx = insert_zero(g, pos) | (uint64_t(bit) << pos);
    if (unpack) st_stream(state + idx, ld_stream(staging + g));
    else st_stream(staging + g, ld_stream(state + idx));
  }
}

// ---- K7: pairwise half-shard swap through NVLink peer memory ------------------------------------------------
// mine[a] <-> peer[b] for every index i of this rank's share; 4 independent pairs in flight per thread (remote
// latency is ~2 us: bytes in flight, not arithmetic, set the rate).
template <typename C, int U>
__global__ void __launch_bounds__(256) k7_swap_half_p2p(C* __restrict__ mine, C* __restrict__ peer, int pos, int mybit, uint64_t begin,
                                                        uint64_t end) {
  const uint64_t stride = uint64_t(gridDim.x) * blockDim.x;
  const uint64_t ma = uint64_t(1 - mybit) << pos, mb = uint64_t(mybit) << pos;
  for (uint64_t i0 = begin + uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; i0 < end; i0 += stride * U) {
    C x[U], y[U];
    uint64_t base[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const uint64_t i = i0 + uint64_t(u) * stride;
      base[u] = insert_zero(i, pos);
      if (i < end) {
        x[u] = ld_stream(mine + (base[u] | ma));
        y[u] = ld_stream(peer + (base[u] | mb));
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const uint64_t i = i0 + uint64_t(u) * stride;
      if (i < end) {
        st_stream(mine + (base[u] | ma), y[u]);
        st_stream(peer + (base[u] | mb), x[u]);
      }
    }
  }
}

inline int launch_swap_half_p2p(cudaStream_t stream, int sm_count, void* state, void* peer, int nqubits, int dtype, int pos, int mybit,
                                int part, int nparts) {
  const uint64_t half = uint64_t(1) << (nqubits - 1);
  const uint64_t begin = half / nparts * part, end = part == nparts - 1 ? half : half / nparts * (part + 1);
  const int grid = sm_count * env_int("QB_P2P_BLOCKS_PER_SM", 8);
  const int unroll = env_int("QB_P2P_UNROLL", 4);
  if (dtype == QB_C128) {
    if (unroll >= 8) k7_swap_half_p2p<double2, 8><<<grid, 256, 0, stream>>>((double2*)state, (double2*)peer, pos, mybit, begin, end);
    else if (unroll >= 4) k7_swap_half_p2p<double2, 4><<<grid, 256, 0, stream>>>((double2*)state, (double2*)peer, pos, mybit, begin, end);
    else k7_swap_half_p2p<double2, 2><<<grid, 256, 0, stream>>>((double2*)state, (double2*)peer, pos, mybit, begin, end);
  } else {
    if (unroll >= 8) k7_swap_half_p2p<float2, 8><<<grid, 256, 0, stream>>>((float2*)state, (float2*)peer, pos, mybit, begin, end);
    else k7_swap_half_p2p<float2, 4><<<grid, 256, 0, stream>>>((float2*)state, (float2*)peer, pos, mybit, begin, end);
  }
  return cudaPeekAtLastError() == cudaSuccess ? QB_OK : QB_ERR_CUDA;
}

// ---- K7b: several global<->local exchanges at once = an all-to-all of contiguous chunks ------------------------------
// k exchanges on the k leading local bits move chunk t of rank r to chunk t' of rank r' (an involution): ONE kernel swaps
// this rank's share of every pair through the peers' mapped shards -- (2^k - 1) / 2^k of a shard crosses NVLink instead
// of k / 2.  Block b serves peer b % npeers, so all pairs progress together.
struct A2ATable {
  void* peer[8];
  uint64_t my_off[8], peer_off[8], begin[8], end[8];  // amplitudes; [begin, end) = this rank's share of the pair's chunk
  int npeers;
  int pad;
};
template <typename C, int U>
__global__ void __launch_bounds__(256) k7_alltoall_p2p(C* __restrict__ mine, const __grid_constant__ A2ATable tab) {
  const int p = blockIdx.x % tab.npeers;
  const uint64_t stride = uint64_t(gridDim.x / tab.npeers) * blockDim.x;
  C* __restrict__ a = mine + tab.my_off[p];
  C* __restrict__ b = reinterpret_cast<C*>(tab.peer[p]) + tab.peer_off[p];
  const uint64_t end = tab.end[p];
  for (uint64_t i0 = tab.begin[p] + uint64_t(blockIdx.x / tab.npeers) * blockDim.x + threadIdx.x; i0 < end; i0 += stride * U) {
    C x[U], y[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const uint64_t i = i0 + uint64_t(u) * stride;
      if (i < end) {
        x[u] = ld_stream(a + i);
        y[u] = ld_stream(b + i);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const uint64_t i = i0 + uint64_t(u) * stride;
      if (i < end) {
        st_stream(a + i, y[u]);
        st_stream(b + i, x[u]);
      }
    }
  }
}

// Out-of-place variant: every chunk is read locally and WRITTEN into the destination rank's second buffer (or this
// rank's own for the chunk that stays) -- remote stores only, no remote loads.  The shards then flip buffers.
template <typename C, int U>
__global__ void __launch_bounds__(256) k7_alltoall_push(const C* __restrict__ mine, const __grid_constant__ A2ATable tab) {
  const int p = blockIdx.x % tab.npeers;
  const uint64_t stride = uint64_t(gridDim.x / tab.npeers) * blockDim.x;
  const C* __restrict__ a = mine + tab.my_off[p];
  C* __restrict__ b = reinterpret_cast<C*>(tab.peer[p]) + tab.peer_off[p];
  const uint64_t end = tab.end[p];
  for (uint64_t i0 = tab.begin[p] + uint64_t(blockIdx.x / tab.npeers) * blockDim.x + threadIdx.x; i0 < end; i0 += stride * U) {
    C x[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const uint64_t i = i0 + uint64_t(u) * stride;
      if (i < end) x[u] = ld_stream(a + i);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const uint64_t i = i0 + uint64_t(u) * stride;
      if (i < end) st_stream(b + i, x[u]);
    }
  }
}

inline int launch_alltoall_push(cudaStream_t stream, int sm_count, const void* state, int dtype, const A2ATable& tab) {
  const int per_peer = (sm_count * env_int("QB_P2P_BLOCKS_PER_SM", 8) + tab.npeers - 1) / tab.npeers;
  const int grid = per_peer * tab.npeers;
  if (dtype == QB_C128) k7_alltoall_push<double2, 8><<<grid, 256, 0, stream>>>((const double2*)state, tab);
  else k7_alltoall_push<float2, 8><<<grid, 256, 0, stream>>>((const float2*)state, tab);
  return cudaPeekAtLastError() == cudaSuccess ? QB_OK : QB_ERR_CUDA;
}

inline int launch_alltoall_p2p(cudaStream_t stream, int sm_count, void* state, int dtype, const A2ATable& tab) {
  const int per_peer = (sm_count * env_int("QB_P2P_BLOCKS_PER_SM", 8) + tab.npeers - 1) / tab.npeers;
  const int grid = per_peer * tab.npeers;
  if (dtype == QB_C128) k7_alltoall_p2p<double2, 4><<<grid, 256, 0, stream>>>((double2*)state, tab);
  else k7_alltoall_p2p<float2, 4><<<grid, 256, 0, stream>>>((float2*)state, tab);
  return cudaPeekAtLastError() == cudaSuccess ? QB_OK : QB_ERR_CUDA;
}

inline int launch_half_copy(cudaStream_t stream, int sm_count, void* state, void* staging, int nqubits, int dtype, int pos, int bit,
                            int unpack) {
  const uint64_t half = uint64_t(1) << (nqubits - 1);
  uint64_t blocks = (half + 255) / 256;
  uint64_t cap = (uint64_t)sm_count * 32;
  int grid = (int)(blocks < cap ? blocks : cap);
  if (dtype == QB_C128)
    k7_half_copy<double2><<<grid, 256, 0, stream>>>((double2*)state, (double2*)staging, half, pos, bit, unpack);
  else
    k7_half_copy<float2><<<grid, 256, 0, stream>>>((float2*)state, (float2*)staging, half, pos, bit, unpack);
  return cudaPeekAtLastError() == cudaSuccess ? QB_OK : QB_ERR_CUDA;
}
}  // namespace qb
