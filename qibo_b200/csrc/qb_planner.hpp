// qibo_b200: host-side sweep planner ("several gates per HBM sweep").
//
// The reference applies the gate queue one gate at a time (abstract.py:3321-3322), four full-state copies
// per gate.  Here the queue is packed into SWEEPS: one sweep = one pass of the state through shared
// memory in tiles of 2^T amplitudes.  A tile spans the L lowest state bits (so every contiguous run is
// >= 512 B and is moved by one bulk-async copy) plus up to T-L arbitrary higher bits.  Any gate whose
// TARGET bits all lie inside the tile can be applied while the tile is resident; CONTROL bits may lie
// anywhere (a control outside the tile predicates the whole tile) and DIAGONAL gates never constrain the
// tile at all (bits outside the tile are constant per tile, so they fold into a per-tile scalar).
//
// Device program of one sweep (the "blob"): SweepHeader, DevOp[nops], then 16-byte aligned payloads
// (matrices / tables in the state's precision).
#pragma once
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <cstring>
#include <deque>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/qibo_b200.h"
#include "qb_canon.hpp"

namespace qb {

// ---- device program layout (shared with qb_passes.cuh / qb_sweep.cuh) --------------------------------
// A sweep is a list of PASSES over the shared-memory tile.  A REGTILE pass fixes R tile-local "register
// bits": every thread owns groups of 2^R amplitudes that differ only in those bits, loads them once,
// applies the pass's whole list of MICRO-OPS in registers and stores them once.  A BIG pass applies one
// dense gate on 3..6 targets (threads share a group, inputs re-read from shared memory).
enum MicroType { MU_DENSE1 = 1, MU_DENSE2 = 2, MU_SWAP = 3, MU_FAN = 4, MU_DIAGK = 5, MU_PHASE = 6, MU_SIGNS = 7 };
enum PassKind { PASS_REGTILE = 1, PASS_BIG = 2 };

// One handler code per micro-op: (operation, register bit(s), control mode) flattened so that the kernel needs a
// single jump table.  "+I" = register-bit index 0..3; "+P" = register-bit pair (hi,lo): (1,0) (2,0) (2,1) (3,0) (3,1) (3,2).
enum MicroHandler {
  MH_NOP = 0,
  MH_ADDSUB = 1,    // +I  (a+b, a-b): H-like gate whose scalar has been moved into another gate of the sweep
  MH_REAL1 = 5,     // +I  real 2x2, no register-bit control
  MH_CPLX1 = 9,     // +I  complex 2x2, no register-bit control
  MH_CPLX1_M = 13,  // +I  complex 2x2, run-time register-bit control mask
  MH_XPAIR = 17,    // +I  pair exchange (X, CNOT, TOFFOLI...), run-time register-bit control mask
  MH_DENSE2 = 21,   // +P  4x4 (matrix index MSB = hi), run-time mask
  MH_SWAP = 27,     // +P  run-time mask
  MH_FAN_C = 33,    // +I  fan whose only register-bit control is bit I
  MH_FAN_NC = 37,   //     fan without register-bit control
  MH_FAN_M = 38,    //     fan, run-time mask
  MH_PHASE_C = 39,  // +I
  MH_PHASE_NC = 43,
  MH_PHASE_M = 44,
  MH_DIAGK = 45,
  MH_STAGE_A = 46,  // +I  (a+b, a-b) on bit I, then a fan whose only control is bit I (a QFT stage in one dispatch)
  MH_STAGE_R = 50,  // +I  real 2x2 on bit I, then that fan
  MH_REAL_LAYER = 54,  //  up to R uncontrolled real 2x2 gates on distinct register bits (an RY layer) in one dispatch:
                       //  `k` = mask of the register bits present, payload = their matrices (4 entries each, .x used)
  MH_PHASE_C2 = 55,    // +P  lone phase whose two register-bit controls are the pair (hi, lo): CZ / CU1 with both qubits
                       //  among the register bits -- compile-time masks (ncu, round 2: the run-time mask form MH_PHASE_M
                       //  tests every register index, ~3 instructions per amplitude instead of 1/4)
  MH_SIGNS = 61,       //     a SET of +-1 diagonal gates (CZ, Z and products of them) merged into one dispatch: the sign of
                       //     register index j is bit j of  G16 ^ sum_r a_r MASK_r ^ c  with a_r / c parities of the thread's
                       //     tile bits (and, through the slot, of the bits outside the tile) -- round 2: a lone CZ costs as
                       //     much as an RY in this kernel (all of it dispatch), the ansatz has as many CZs as RYs
  MH_COUNT = 62,
};

constexpr int SWEEP_MAX_SLOTS = 48;        // ops per sweep that need per-tile set-up (controls / factors from outside the tile)
constexpr int SWEEP_BLOB_MAX = 26 * 1024;  // program bytes resident in shared memory next to the tiles
constexpr int SWEEP_TILE_BYTES_LOG2 = 16;  // 64 KiB tiles, three in flight per SM
constexpr uint32_t MU_NO_SLOT = 0xFFFF;
#ifndef QB_TEAM_THREADS
#define QB_TEAM_THREADS 256
#endif
constexpr int SWEEP_TEAM_THREADS = QB_TEAM_THREADS;  // compute threads that share one tile (8 warps); the group tables are laid out for it
constexpr int FAN_EXT_CHUNK = 4;           // a fan's bits outside the tile are folded through tables of 2^4 entries (<= 8 tables: 32 bits)
constexpr size_t BIG_PAYLOAD_SMEM_MAX = 16 * 1024 + 64;  // larger dense matrices (6 targets) stay in global memory

// per-tile, per-team state of a slot (the program itself stays constant, so two teams of compute warps can work on
// two tiles at once)
struct TileSlot {
  double ext[2];         // FAN: factor from the bits outside the tile, in the state's precision (a C)
  uint32_t active;       // controls outside the tile are all 1
  uint32_t aux;          // DIAGK: table-index part from the bits outside the tile
  uint32_t pad[2];
};
static_assert(sizeof(TileSlot) == 32, "TileSlot layout");

struct alignas(16) MicroOp {
  // ---- hot header: one 16-byte shared-memory load decodes the op ------------------------------------
  uint8_t handler;       // MicroHandler
  uint8_t type;          // MicroType
  uint16_t slot;         // TileSlot index, or MU_NO_SLOT: always active, nothing depends on the tile
  uint32_t creg;         // controls that are register bits (mask over the register index j)
  uint32_t cthr;         // controls inside the tile but outside R (tile-local mask, tested on the group base)
  uint32_t payload;      // byte offset in the blob: DENSE2 matrix | FAN: TA, TB, G, ext tables | DIAGK: table
  // ---- second 16 bytes ---------------------------------------------------------------------------------
  uint16_t la;           // FAN: TA is indexed by the low `la` bits of the group index, TB by the rest
  uint16_t R;            // register bits of the pass this op belongs to (table layout)
  uint32_t k;            // DIAGK: number of target bits | FAN: byte offset of TB in the blob (filled with the payload offset)
  uint32_t n_ext;        // FAN: number of ext tables
  uint32_t gt_off;       // FAN: byte offset of G in the blob
  // ---- inline data -------------------------------------------------------------------------------------
  double inl[8];         // DENSE1: the 2x2 matrix in the state's precision (4 x C); PHASE: the phase (a C)
  // ---- cold part (per-tile set-up, DIAGK) ------------------------------------------------------------------
  uint64_t ext_cmask;    // controls outside the tile (state bit positions)
  uint8_t tbit[8];       // DIAGK: tile-local bit of target i if it is outside R, else 0xFF
  uint8_t rsel[8];       // DIAGK: register-bit index of target i if it is in R, else 0xFF
  uint64_t ext_mask[8];  // FAN: state-bit mask of ext table e; DIAGK: state-bit mask of target i when outside the tile
  double scalar[2];      // FAN: global factor
  uint64_t pad1;
};
static_assert(sizeof(MicroOp) % 16 == 0, "MicroOp must stay 16-byte aligned");

struct DevOp {           // BIG pass: one dense gate on k = 3..6 tile-local targets
  uint32_t k;
  uint32_t nins;         // tile-local bits to insert (targets + tile-local controls)
  uint32_t ins_mask;
  uint32_t tl_cmask;     // controls inside the tile
  uint64_t ext_cmask;    // controls outside the tile
  uint32_t payload;      // matrix
  uint32_t slot;
  uint8_t tbit[8];       // tile-local bit of target i (tbit[0] = MSB of the matrix index)
  uint32_t payload_global;  // 1: the matrix is too large for shared memory and is read from the global copy of the blob
  uint32_t pad2;
};
static_assert(sizeof(DevOp) % 16 == 0, "DevOp must stay 16-byte aligned");

struct PassHeader {
  uint32_t kind;
  uint32_t rmask;        // REGTILE: tile-local mask of the R register bits
  uint16_t nmicro;
  uint16_t R;            // REGTILE: number of register bits of this pass
  uint32_t offset;       // byte offset of MicroOp[0] (REGTILE) or of the DevOp (BIG)
  uint16_t off[14];      // REGTILE: tile-local offset of register index j (deposit of j into rmask), j < 14 (host-side debugging aid)
  uint32_t flags;        // PASS_FULL_STAGE: see below
  uint8_t pos[8];        // REGTILE: tile-local positions of the register bits, ascending
  uint32_t gtab;         // REGTILE: byte offset of uint16[SWEEP_TEAM_THREADS * groups-per-thread]: the group (index over the
                         // non-register tile bits, ascending) that thread `ctid` handles as its u-th, or 0xFFFF for none
  uint32_t stage_mask;   // REGTILE: non-zero when the pass is nothing but fused stage ops (MH_STAGE_*) on strictly descending
                         // register bits -- the shape of every QFT pass; bit I set = a stage on register bit I.  The kernel
                         // then runs them as straight-line code, without the per-op jump sequence
  // PASS_PERMUTED_STORE (the last pass of a permuting sweep): the groups go back to shared memory in the destination layout
  uint8_t dpos[8];       // tile-local DESTINATION positions of the register bits (same order as pos[])
  uint32_t dtab;         // byte offset of uint16[thread slots]: destination-local index of the slot's group base
  uint32_t pad1;
};
static_assert(sizeof(PassHeader) % 16 == 0, "PassHeader must stay 16-byte aligned");
// PASS_FULL_STAGE: the pass is exactly R fused stages of the add/sub kind (MH_STAGE_A) on register bits R-1 .. 0, none with
// a control outside the tile, on a full tile (every thread owns valid groups): the kernel runs them as ONE basic block --
// no handler decode, no activity test, no validity test per stage -- so that the next stage's table loads can be
// scheduled under the current stage's arithmetic (round 2: the QFT sweeps are issue-bound at ~13 instructions per
// amplitude and stage, 5.2 of them FP64).  Every pass of a QFT but the one holding its last Hadamard qualifies.
constexpr uint32_t PASS_PERMUTED_STORE = 2u;  // the pass writes its groups in the destination layout of a permuting sweep
constexpr uint32_t PASS_PERMUTED_DSWZ = 4u;   // ... which is swizzled
constexpr uint32_t PASS_PAIRED_GROUPS = 8u;   // complex64: the thread's two groups differ in tile bit 0 alone, so one 16-byte
                                              // shared-memory access moves an amplitude of each (cost model, round 2: a
                                              // complex64 pass ran at twice the shared-memory floor on 8-byte accesses)
constexpr uint32_t PASS_FULL_STAGE = 1u;  // flags bits 4-7: which of the stages use the real-matrix form (MH_STAGE_R)
// (stage mask | real mask << 4) combinations the kernel instantiates: four stages, or the three lowest (the second pass of a
// 7-stage sweep), each with or without a real-matrix stage on bit 0
inline bool pass_full_stage_supported(uint32_t key) { return key == 0x0F || key == 0x1F || key == 0x07 || key == 0x17; }

struct SweepHeader {
  uint32_t npasses;
  uint32_t T;           // log2(amplitudes per tile)
  uint32_t L;           // log2(amplitudes per contiguous run)
  uint32_t blob_bytes;
  uint64_t tile_mask;   // state bits spanned by the tile
  uint64_t other_mask;  // remaining state bits (enumerated by the tile index)
  uint64_t ntiles;
  uint32_t passes_offset;  // byte offset of PassHeader[0]
  uint32_t nslots;
  uint32_t R;              // register bits of the REGTILE passes
  uint32_t slots_offset;   // uint32[nslots]: byte offset of the slot's MicroOp, or of its DevOp with bit 31 set (ops that need per-tile set-up)
  uint32_t swizzle;        // 1: the tile sits in shared memory in the TMA 128-byte swizzle (qb_passes.cuh swz); needs the tensor-map copy
  uint32_t warp_private;   // 1: every warp owns the same 1/8 of the tile in every pass (no REGTILE pass mixes across three fixed
                           // tile bits): passes are separated by __syncwarp instead of a team barrier, so warps drift apart and
                           // one warp's shared-memory bursts overlap another's FP math
  // ---- permuting sweep (a trailing qubit permutation -- the QFT's bit reversal -- rides on the last sweep) -----------
  uint32_t permuted;       // 1: the last pass leaves the tile in the DESTINATION layout and the storer writes it through
                           // the destination tensor map at the permuted tile base (out of place)
  uint32_t dswizzle;       // the destination layout is swizzled
  uint64_t dtile_mask;     // destination state bits spanned by the (permuted) tile
  uint8_t dst_bit[64];     // destination bit of every source state bit (bits outside the tile: the storer's tile base)
};
static_assert(sizeof(SweepHeader) % 16 == 0, "SweepHeader must stay 16-byte aligned");

// ---- plan ------------------------------------------------------------------------------------------
struct PlanOp {
  int kind = 0;                 // CK_DENSE / CK_SWAP / CK_DIAG (general table) / CK_PHASE (fan)
  std::vector<int> tpos;        // DENSE/SWAP/DIAG targets (tpos[0] = MSB)
  std::vector<int> cpos;        // control bit positions
  std::vector<cd> data;         // DENSE matrix / DIAG table
  std::map<int, std::pair<cd, cd>> fan;  // FAN: bit position -> (factor if bit == 0, factor if bit == 1)
  cd scalar = cd(1.0, 0.0);
  std::vector<int> src;         // indices of the original ops merged into this one
  int special = 0;              // 1: apply as (a+b, a-b); the scalar of the gate rides in another gate of the sweep
  // CK_SIGNS (made inside emit_regtile_pass from runs of +-1 diagonal gates): x -> (-1)^(neg + sum over terms of x_a [x_b])
  std::vector<std::pair<int, int>> sign_terms;  // (bit a, bit b) or (bit a, -1)
  int sign_neg = 0;
};
constexpr int CK_SIGNS = 100;
constexpr int SIGNS_MAX_PAIRS = 27;  // thread-level pairs one MH_SIGNS op can carry inline

struct SweepDesc {
  int permuted = 0;          // the sweep writes its tiles, permuted, into ANOTHER buffer (launch_sweep needs `dst`)
  int dswizzle = 0;
  uint64_t dmask = 0;        // destination tile mask
  int swizzle = 0;
  int stage_only = 0;  // every pass is a straight-line stage pass: the lean two-team kernel instantiation may run it
  int T = 0, L = 0;
  uint64_t tile_mask = 0;
  size_t blob_offset = 0, blob_bytes = 0;
  int npasses = 0, ndiag = 0;
  uint64_t ntiles = 0;
};

// A qubit permutation to apply after the ops: pi[b] = destination bit of source state bit b.
struct PermSpec {
  int pi[64];
};
// its tile geometry when the permutation rides on a sweep: the `ls` lowest source bits (contiguous reads), the source bits
// that land on the `ld` lowest destination bits (contiguous writes) and, when those leave room in the tile, the bits the
// last gates of the program mix
struct PermGeom {
  uint64_t S = 0, D = 0;    // source / destination tile masks
  int sigma[16];            // tile-local source bit -> tile-local destination bit
  int dswizzle = 0;
};

struct Plan {
  std::vector<SweepDesc> sweeps;
  std::vector<char> blob;
  std::vector<int> sweep_of_op;
  int npasses = 0, ndiag = 0;
  // the schedule itself, kept so that the same gate structure with NEW matrices (Circuit.set_parameters,
  // models/circuit.py:788-857) is re-emitted without scheduling again: the merged ops each sweep took, in order, and
  // the mixing / diagonal bit sets the schedule's legality rests on
  std::vector<std::vector<uint32_t>> sweep_ops;
  std::vector<uint64_t> xsets, dsets;
  int perm_fused = 0;  // the trailing permutation the caller asked for rides on the last sweep (its tile: pgeom)
  PermGeom pgeom;
};

inline int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return v && *v ? atoi(v) : dflt;
}

inline int tile_bits_for(int dtype) { return SWEEP_TILE_BYTES_LOG2 - (dtype == QB_C128 ? 4 : 3); }

// ---- step 1: merge consecutive diagonal gates into fans -----------------------------------------------
inline void merge_ops(const std::vector<CanonOp>& ops, bool no_fuse, std::vector<PlanOp>& out) {
  out.clear();
  bool open = false;  // out.back() is an open fan that may still absorb phases
  auto fan_set = [](const PlanOp& f) {
    std::vector<int> s(f.cpos);
    for (auto& kv : f.fan) s.push_back(kv.first);
    std::sort(s.begin(), s.end());
    return s;
  };
  for (size_t i = 0; i < ops.size(); ++i) {
    const CanonOp& o = ops[i];
    if (o.kind == CK_NOOP) continue;
    if (o.kind == CK_DENSE || o.kind == CK_SWAP || (o.kind == CK_DIAG && o.tpos.size() >= 2)) {
      PlanOp p;
      p.kind = o.kind;
      p.tpos = o.tpos;
      p.cpos = o.cpos;
      p.data = o.data;
      p.src.push_back((int)i);
      out.push_back(std::move(p));
      open = false;
      continue;
    }
    // PHASE (scalar on all-controls slice) or 1-bit DIAG: a fan entry
    std::vector<int> P(o.cpos);  // sorted
    int bit = -1;
    std::pair<cd, cd> entry;
    if (o.kind == CK_DIAG) {  // k == 1
      bit = o.tpos[0];
      entry = {o.data[0], o.data[1]};
    } else {
      entry = {cd(1.0, 0.0), o.data[0]};
    }
    bool merged = false;
    if (open && !no_fuse) {
      PlanOp& f = out.back();
      if (bit >= 0) {  // fixed target bit: controls must equal the fan's
        if (P == f.cpos) {
          auto it = f.fan.find(bit);
          if (it == f.fan.end()) f.fan[bit] = entry;
          else it->second = {it->second.first * entry.first, it->second.second * entry.second};
          merged = true;
        }
      } else {
        // pure phase on the set P: any member may play the target
        if (!f.cpos.empty() && P == f.cpos) {
          // exactly the fan's controls (a CU1 whose other qubit is a global qubit set to 1 on this rank): a scalar on
          // the slice the fan acts on
          f.scalar *= entry.second;
          merged = true;
        } else if (P.size() == f.cpos.size() + 1 && std::includes(P.begin(), P.end(), f.cpos.begin(), f.cpos.end())) {
          int b = -1;
          for (int x : P)
            if (!std::binary_search(f.cpos.begin(), f.cpos.end(), x)) b = x;
          auto it = f.fan.find(b);
          if (it == f.fan.end()) f.fan[b] = entry;
          else it->second.second *= entry.second;
          merged = true;
        } else if (f.fan.size() == 1 && f.fan.begin()->second.first == cd(1.0, 0.0) && f.scalar == cd(1.0, 0.0)) {
          // re-root a single-entry fan so that the shared bits become the controls (a pure phase on its bit set: only
          // then is it symmetric in its bits -- a scalar on the control slice breaks that)
          std::vector<int> P0 = fan_set(f), common;
          std::set_intersection(P.begin(), P.end(), P0.begin(), P0.end(), std::back_inserter(common));
          if (!P.empty() && P.size() + 1 == P0.size() && common.size() == P.size()) {
            // P is the fan's set minus one bit b0: control on P, and b0 selects between the new phase alone and the
            // product (the tail of a QFT on a rank whose global control qubits are 1)
            cd ph0 = f.fan.begin()->second.second;
            int b0 = -1;
            for (int x : P0)
              if (!std::binary_search(P.begin(), P.end(), x)) b0 = x;
            f.cpos = P;
            f.fan.clear();
            f.fan[b0] = {entry.second, entry.second * ph0};
            merged = true;
          } else if (P.size() == P0.size() && common.size() + 1 == P0.size()) {
            cd ph0 = f.fan.begin()->second.second;
            int b0 = -1, b1 = -1;
            for (int x : P0)
              if (!std::binary_search(common.begin(), common.end(), x)) b0 = x;
            for (int x : P)
              if (!std::binary_search(common.begin(), common.end(), x)) b1 = x;
            f.cpos = common;
            f.fan.clear();
            f.fan[b0] = {cd(1.0, 0.0), ph0};
            f.fan[b1] = entry;
            merged = true;
          }
        }
      }
      if (merged) f.src.push_back((int)i);
    }
    if (!merged) {
      PlanOp p;
      p.kind = CK_PHASE;
      if (bit >= 0) {
        p.cpos = P;
        p.fan[bit] = entry;
      } else if (P.empty()) {
        p.scalar = entry.second;  // global phase (0 controls, 0 targets)
      } else {
        int b = P.back();  // highest bit plays the target
        P.pop_back();
        p.cpos = P;
        p.fan[b] = entry;
      }
      p.src.push_back((int)i);
      out.push_back(std::move(p));
      open = true;
    }
  }
}

// ---- step 2 + 3: pack into sweeps, split sweeps into passes, serialise ------------------------------
template <typename C> inline C to_dev(cd v);
struct f2 { float x, y; };
struct d2 { double x, y; };
template <> inline f2 to_dev<f2>(cd v) { return f2{(float)v.real(), (float)v.imag()}; }
template <> inline d2 to_dev<d2>(cd v) { return d2{v.real(), v.imag()}; }

inline size_t align16(size_t x) { return (x + 15) & ~size_t(15); }

// register bits per REGTILE pass, per dtype: the kernel (qb_sweep.cuh) instantiates exactly these
#ifndef QB_R128
#define QB_R128 4
#endif
#ifndef QB_R64
#define QB_R64 4
#endif
inline int regtile_bits_for(int dtype) { return dtype == QB_C128 ? QB_R128 : QB_R64; }

// upper bound of the blob bytes a PlanOp needs (independent of the tile)
inline size_t blob_estimate(const PlanOp& p, int csize, int T, int R) {
  switch (p.kind) {
    case CK_DENSE: {
      size_t pay = align16(p.data.size() * csize);
      if (pay > BIG_PAYLOAD_SMEM_MAX) pay = 0;  // kept in global memory
      return (p.tpos.size() <= 2 ? sizeof(MicroOp) : sizeof(DevOp) + sizeof(PassHeader)) + pay;
    }
    case CK_SWAP: return sizeof(MicroOp);
    case CK_DIAG: return sizeof(MicroOp) + align16(p.data.size() * csize);
    default: {  // fan: TA + TB + G + ext tables of <= FAN_EXT_CHUNK bits
      int gb = T - R > 0 ? T - R : 0;
      int la = gb < 5 ? gb : 5;
      size_t local = (size_t(1) << la) + (size_t(1) << (gb - la)) + (size_t(1) << R);
      size_t next = (p.fan.size() + FAN_EXT_CHUNK - 1) / FAN_EXT_CHUNK;
      return sizeof(MicroOp) + align16((local + next * (size_t(1) << FAN_EXT_CHUNK)) * csize);
    }
  }
}

template <typename C> struct SweepBuilder {
  int T = 0, R = 0;
  uint64_t tile_mask = 0;
  std::vector<int> local_of_pos;  // state bit position -> tile-local bit
  std::vector<PassHeader> passes;
  std::vector<std::vector<MicroOp>> micro;   // per pass (empty for BIG)
  std::vector<DevOp> big;                    // per pass (valid for BIG)
  std::vector<std::vector<C>> payloads;      // one per op
  std::vector<std::pair<int, int>> payload_owner;  // payload -> (pass, micro index or -1)
  std::vector<std::pair<int, int>> slots;    // per-tile set-up slot -> (pass, micro index or -1)
  std::vector<std::vector<uint16_t>> gtabs;  // distinct group tables
  std::map<uint32_t, int> gtab_of_rmask;
  std::vector<int> gtab_of_pass;             // per pass: index into gtabs, -1 for BIG
  std::vector<int> dtab_of_pass;             // per pass: index into gtabs of the destination-index table (permuted store), or -1
  uint32_t split_mask = 0;                   // warp-private sweeps: the three tile bits that select the warp
};

inline bool is_real_matrix(const std::vector<cd>& m) {
  for (auto& v : m)
    if (v.imag() != 0.0) return false;
  return true;
}

// s * [[1, 1], [1, -1]] on one uncontrolled target (H): candidates for the add/sub form
inline bool is_hadamard_like(const PlanOp& p) {
  if (p.kind != CK_DENSE || p.tpos.size() != 1 || !p.cpos.empty() || p.data.size() != 4) return false;
  const cd s = p.data[0];
  return s != cd(0.0, 0.0) && p.data[1] == s && p.data[2] == s && p.data[3] == -s;
}

inline int pair_index(int hi, int lo) { return hi * (hi - 1) / 2 + lo; }

// ---- thread -> group mapping of a REGTILE pass ---------------------------------------------------------------
inline uint32_t deposit_u32(uint32_t x, uint32_t mask) {
  uint32_t r = 0;
  for (int k = 0; mask; ++k) {
    uint32_t low = mask & (~mask + 1);
    if ((x >> k) & 1) r |= low;
    mask ^= low;
  }
  return r;
}
inline uint32_t extract_u32(uint32_t x, uint32_t mask) {
  uint32_t r = 0;
  for (int k = 0; mask; ++k) {
    uint32_t low = mask & (~mask + 1);
    if (x & low) r |= 1u << k;
    mask ^= low;
  }
  return r;
}
inline uint32_t swz_host(uint32_t x, int csize, bool swizzle) {
  if (!swizzle) return x;
  return csize == 16 ? x ^ ((x >> 3) & 7u) : x ^ (((x >> 4) & 7u) << 1);
}
// the five lowest tile bits that are neither register bits nor warp-select bits: one per lane bit
inline uint32_t lane_bits_of(int T, uint32_t rmask, uint32_t smask) {
  uint32_t lanes = 0;
  for (int lb = 0; lb < T && __builtin_popcount(lanes) < 5; ++lb)
    if (!(((rmask | smask) >> lb) & 1)) lanes |= 1u << lb;
  return lanes;
}
// shared-memory wavefronts of one warp-wide tile access relative to the conflict-free count (1.0 = no bank conflict)
inline double conflict_cost(int T, int csize, bool swizzle, uint32_t rmask, uint32_t smask) {
  const uint32_t lanes = lane_bits_of(T, rmask, smask);
  const int per_phase = csize == 16 ? 8 : 16;  // lanes served by one 128-byte wavefront
  const int unit_mask = csize == 16 ? 7 : 15;  // 16-byte / 8-byte slots of a 128-byte row
  double total = 0;
  for (int ph = 0; ph < 32 / per_phase; ++ph) {
    int count[16] = {0};
    int worst = 0;
    for (int l = 0; l < per_phase; ++l) {
      const uint32_t t = deposit_u32((uint32_t)(ph * per_phase + l), lanes);
      const int slot = (int)(swz_host(t, csize, swizzle) & unit_mask);
      worst = std::max(worst, ++count[slot]);
    }
    total += worst;
  }
  return total / (32 / per_phase);
}
// register bits of a pass: the demanded ones, padded to R with the highest tile bits outside `smask`
inline uint32_t pad_register_bits(int T, int R, uint32_t need, uint32_t smask) {
  uint32_t rmask = need;
  for (int lb = T - 1; lb >= 0 && __builtin_popcount(rmask) < R; --lb)
    if (!(((rmask | smask) >> lb) & 1)) rmask |= 1u << lb;
  return rmask;
}
// complex64, full tile, two groups per thread, tile bit 0 not a register bit: the pass can pair its groups (PASS_PAIRED_GROUPS)
inline bool paired_groups(int T, int R, int csize, uint32_t rmask, uint32_t smask) {
  if (csize != 8 || smask != 0 || (rmask & 1u) || env_int("QB_NO_PAIRED_GROUPS", 0)) return false;
  return (1u << (T - R)) == 2u * (uint32_t)SWEEP_TEAM_THREADS && T == SWEEP_TILE_BYTES_LOG2 - 3 && R == regtile_bits_for(QB_C64);
}
// uint16 table: thread slot (u * SWEEP_TEAM_THREADS + ctid) -> group index (over the non-register bits, ascending)
inline std::vector<uint16_t> make_group_table(int T, int R, int csize, uint32_t rmask, uint32_t smask) {
  const int gbits = T - R;
  const uint32_t ngroups = 1u << gbits;
  // the kernel reads a fixed number of groups per thread for each dtype (a full tile has 2^8 / 2^9 groups)
  const uint32_t full_groups = 1u << (SWEEP_TILE_BYTES_LOG2 - (csize == 16 ? 4 : 3) - regtile_bits_for(csize == 16 ? QB_C128 : QB_C64));
  const uint32_t gpt = std::max<uint32_t>((ngroups + SWEEP_TEAM_THREADS - 1) / SWEEP_TEAM_THREADS, (full_groups + SWEEP_TEAM_THREADS - 1) / SWEEP_TEAM_THREADS);
  std::vector<uint16_t> tab((size_t)gpt * SWEEP_TEAM_THREADS, 0xFFFF);
  const uint32_t all = (1u << T) - 1;
  if (smask == 0) {
    if (paired_groups(T, R, csize, rmask, smask)) {
      // thread c owns groups 2c and 2c + 1: group bit 0 is tile bit 0, the two amplitudes sit in one 16-byte chunk
      for (uint32_t c = 0; c < (uint32_t)SWEEP_TEAM_THREADS; ++c) {
        tab[c] = (uint16_t)(2 * c);
        tab[(size_t)SWEEP_TEAM_THREADS + c] = (uint16_t)(2 * c + 1);
      }
      return tab;
    }
    for (uint32_t g = 0; g < ngroups; ++g) tab[g] = (uint16_t)g;
    return tab;
  }
  const uint32_t lanes = lane_bits_of(T, rmask, smask);
  const uint32_t rest = all & ~(rmask | smask | lanes);  // walked by the group-per-thread index u
  for (uint32_t u = 0; u < gpt; ++u)
    for (uint32_t tid = 0; tid < (uint32_t)SWEEP_TEAM_THREADS; ++tid) {
      const uint32_t t = deposit_u32(tid & 31u, lanes) | deposit_u32(tid >> 5, smask) | deposit_u32(u, rest);
      tab[(size_t)u * SWEEP_TEAM_THREADS + tid] = (uint16_t)extract_u32(t, all & ~rmask);
    }
  return tab;
}
// Three tile bits that no pass needs as a register bit, chosen for the fewest bank conflicts; 0 = none (team barriers).
inline uint32_t choose_split(int T, int R, int csize, bool swizzle, const std::vector<uint32_t>& needs, bool has_big) {
  // Measured on B200 (QFT(30) complex128, gpurun_out/r3i_*): 9.5 ms per stage sweep with warp-private sub-tiles against
  // 8.9 ms with team barriers -- the barrier stalls it removes are smaller than the bank conflicts its lane mapping
  // adds -- so it stays opt-in.
  if (has_big || !env_int("QB_WARP_PRIVATE", 0)) return 0;
  if (T - R - 3 < 5 || (1u << (T - R)) % SWEEP_TEAM_THREADS) return 0;  // every lane must own a group
  uint32_t used = 0;
  for (uint32_t nd : needs) used |= nd;
  std::vector<int> free_bits;
  for (int lb = 0; lb < T; ++lb)
    if (!((used >> lb) & 1)) free_bits.push_back(lb);
  if (free_bits.size() < 3) return 0;
  double best = 1e30;
  uint32_t best_mask = 0;
  for (size_t a = 0; a < free_bits.size(); ++a)
    for (size_t b = a + 1; b < free_bits.size(); ++b)
      for (size_t c = b + 1; c < free_bits.size(); ++c) {
        const uint32_t sm = (1u << free_bits[a]) | (1u << free_bits[b]) | (1u << free_bits[c]);
        double cost = 0;
        bool ok = true;
        for (uint32_t nd : needs) {
          const uint32_t rm = pad_register_bits(T, R, nd, sm);
          if (__builtin_popcount(rm) != R) { ok = false; break; }
          cost += conflict_cost(T, csize, swizzle, rm, sm);
        }
        if (ok && cost < best - 1e-9) {
          best = cost;
          best_mask = sm;
        }
      }
  return best_mask;
}

// ---- permuting sweep: geometry and the thread -> group tables of its last pass ---------------------------------------
// 16-byte chunk (bank group) of a tile-local index inside its 128-byte row: what must differ between the eight lanes of a
// quarter-warp for a conflict-free LDS.128 / STS.128 (complex64: LDS.64, sixteen lanes, two amplitudes per chunk)
inline uint32_t chunk_of(uint32_t x, int csize, bool swizzle) {
  const uint32_t p = swz_host(x, csize, swizzle);
  return csize == 16 ? (p & 7u) : ((p >> 1) & 7u);
}
inline int rank3(uint32_t a, uint32_t b, uint32_t c) {  // rank over GF(2) of three 3-bit vectors
  uint32_t v[3] = {a & 7u, b & 7u, c & 7u};
  int r = 0;
  for (int bit = 2; bit >= 0; --bit) {
    int piv = -1;
    for (int i = r; i < 3; ++i)
      if ((v[i] >> bit) & 1u) { piv = i; break; }
    if (piv < 0) continue;
    std::swap(v[r], v[piv]);
    for (int i = 0; i < 3; ++i)
      if (i != r && ((v[i] >> bit) & 1u)) v[i] ^= v[r];
    ++r;
  }
  return r;
}

// Thread slot -> group of the last pass of a permuting sweep.  Lane bits 0..2 (the eight lanes of a quarter-warp) must
// reach eight different 16-byte chunks BOTH when the pass loads its groups (source layout) and when it stores them
// (destination layout): lane bit k flips a set V_k of one or two free tile bits, chosen so that the chunk images of
// V_0, V_1, V_2 are linearly independent on both sides (a bit reversal maps the low source bits to high destination
// bits: the lanes then walk a diagonal).  Returns the number of conflict ways left (1 = none).
inline int make_permuted_tables_uncached(int T, int R, int csize, uint32_t rmask, const int* sigma, bool sswz, bool dswz, uint32_t nthreads,
                                         std::vector<uint16_t>& gtab, std::vector<uint16_t>& dtab);
// (the search below walks ~40 k lane assignments: a program planned again -- every step of a loop that sends host matrices
// -- finds its tables here)
inline int make_permuted_tables(int T, int R, int csize, uint32_t rmask, const int* sigma, bool sswz, bool dswz, uint32_t nthreads,
                                std::vector<uint16_t>& gtab, std::vector<uint16_t>& dtab) {
  struct Entry { std::vector<uint16_t> g, d; int ways; };
  static std::mutex mu;
  static std::map<std::vector<int>, Entry> cache;
  std::vector<int> key = {T, R, csize, (int)rmask, (int)sswz, (int)dswz, (int)nthreads};
  for (int b = 0; b < T; ++b) key.push_back(sigma[b]);
  {
    std::lock_guard<std::mutex> lk(mu);
    auto it = cache.find(key);
    if (it != cache.end()) {
      gtab = it->second.g;
      dtab = it->second.d;
      return it->second.ways;
    }
  }
  const int ways = make_permuted_tables_uncached(T, R, csize, rmask, sigma, sswz, dswz, nthreads, gtab, dtab);
  std::lock_guard<std::mutex> lk(mu);
  if (cache.size() > 256) cache.clear();
  cache[key] = Entry{gtab, dtab, ways};
  return ways;
}
inline int make_permuted_tables_uncached(int T, int R, int csize, uint32_t rmask, const int* sigma, bool sswz, bool dswz, uint32_t nthreads,
                                         std::vector<uint16_t>& gtab, std::vector<uint16_t>& dtab) {
  std::vector<int> F;  // free (non-register) tile-local source positions, ascending
  for (int b = 0; b < T; ++b)
    if (!((rmask >> b) & 1u)) F.push_back(b);
  const int nf = (int)F.size();
  std::vector<uint32_t> cs(T), cd(T);
  for (int b = 0; b < T; ++b) {
    cs[b] = chunk_of(1u << b, csize, sswz);
    cd[b] = chunk_of(1u << sigma[b], csize, dswz);
  }
  // candidate sets: single free positions and pairs
  struct Cand { int a, b; uint32_t A, B; };
  std::vector<Cand> cands;
  for (int i = 0; i < nf; ++i) cands.push_back({F[i], -1, cs[F[i]], cd[F[i]]});
  for (int i = 0; i < nf; ++i)
    for (int j = 0; j < nf; ++j)
      if (i != j) cands.push_back({F[i], F[j], cs[F[i]] ^ cs[F[j]], cd[F[i]] ^ cd[F[j]]});
  int best[3] = {-1, -1, -1}, best_rank = -1;
  long best_cost = 0;
  for (size_t x = 0; x < cands.size(); ++x)
    for (size_t y = x + 1; y < cands.size(); ++y) {
      const Cand &cx = cands[x], &cy = cands[y];
      if (cx.a == cy.a || cx.a == cy.b || (cx.b >= 0 && (cx.b == cy.a || cx.b == cy.b))) continue;
      for (size_t z = y + 1; z < cands.size(); ++z) {
        const Cand& cz = cands[z];
        const int used[4] = {cx.a, cx.b, cy.a, cy.b};
        bool clash = false;
        for (int u : used)
          if (u >= 0 && (u == cz.a || u == cz.b)) clash = true;
        if (clash) continue;
        const int rk = rank3(cx.A, cy.A, cz.A) + rank3(cx.B, cy.B, cz.B);
        const long cost = (cx.b >= 0) + (cy.b >= 0) + (cz.b >= 0);  // prefer plain lanes when they are enough
        if (rk > best_rank || (rk == best_rank && cost < best_cost)) {
          best_rank = rk;
          best_cost = cost;
          best[0] = (int)x;
          best[1] = (int)y;
          best[2] = (int)z;
        }
        if (best_rank == 6 && best_cost == 0) break;
      }
    }
  if (best[0] < 0) return 8;
  // slot bit -> set of free positions it flips
  const int gbits = nf;
  std::vector<uint32_t> flips(gbits, 0);
  uint32_t taken = 0;
  std::vector<int> second;  // second members: each gets an independent slot bit of its own
  for (int k = 0; k < 3; ++k) {
    const Cand& c = cands[best[k]];
    flips[k] = 1u << c.a;
    taken |= 1u << c.a;
    if (c.b >= 0) {
      flips[k] |= 1u << c.b;
      taken |= 1u << c.b;
      second.push_back(c.b);
    }
  }
  int sb = 3;
  for (int b : F)
    if (!((taken >> b) & 1u)) flips[sb++] = 1u << b;
  for (int b : second) flips[sb++] = 1u << b;
  const uint32_t nslots = 1u << gbits;
  const uint32_t full_groups = 1u << (SWEEP_TILE_BYTES_LOG2 - (csize == 16 ? 4 : 3) - regtile_bits_for(csize == 16 ? QB_C128 : QB_C64));
  const uint32_t gpt = std::max<uint32_t>((nslots + nthreads - 1) / nthreads, (full_groups + nthreads - 1) / nthreads);
  gtab.assign((size_t)gpt * nthreads, 0xFFFF);
  dtab.assign((size_t)gpt * nthreads, 0);
  const uint32_t all = (1u << T) - 1;
  for (uint32_t slot = 0; slot < nslots; ++slot) {
    uint32_t t = 0;
    for (int k = 0; k < gbits; ++k)
      if ((slot >> k) & 1u) t ^= flips[k];
    uint32_t y = 0;
    for (int b = 0; b < T; ++b)
      if ((t >> b) & 1u) y |= 1u << sigma[b];
    gtab[slot] = (uint16_t)extract_u32(t, all & ~rmask);
    dtab[slot] = (uint16_t)y;
  }
  return best_rank == 6 ? 1 : 2;
}

// Emits the micro-ops of one REGTILE pass given its final register-bit mask.
template <typename C>
inline bool emit_regtile_pass(SweepBuilder<C>& sb, const std::vector<const PlanOp*>& ops, uint32_t rmask, std::string& err,
                              const PermGeom* pg = nullptr, bool sswz = false) {
  const int T = sb.T, R = __builtin_popcount(rmask);
  std::vector<int> rbit_of_local(32, -1), gbit_of_local(32, -1);
  {
    int r = 0, g = 0;
    for (int lb = 0; lb < T; ++lb) {
      if ((rmask >> lb) & 1) rbit_of_local[lb] = r++;
      else gbit_of_local[lb] = g++;
    }
  }
  const int gb = T - R;
  const int la = gb < 5 ? gb : 5;
  PassHeader ph;
  memset(&ph, 0, sizeof(ph));
  ph.kind = PASS_REGTILE;
  ph.rmask = rmask;
  ph.R = (uint16_t)R;
  for (int j = 0; j < (1 << R) && j < 14; ++j) {
    uint32_t o = 0;
    int kbit = 0;
    for (int lb = 0; lb < T; ++lb)
      if ((rmask >> lb) & 1) {
        if ((j >> kbit) & 1) o |= 1u << lb;
        ++kbit;
      }
    ph.off[j] = (uint16_t)o;
  }
  {
    int kbit = 0;
    for (int lb = 0; lb < T; ++lb)
      if ((rmask >> lb) & 1) ph.pos[kbit++] = (uint8_t)lb;
  }
  std::vector<MicroOp> mops;
  const int pass_index = (int)sb.passes.size();
  const PlanOp* prev_op = nullptr;
  // ---- runs of +-1 diagonal gates (CZ, Z, CZ fans) -> one CK_SIGNS op each.  Diagonal gates commute with each other and
  // with every gate that mixes none of their bits, so a sign gate may wait (move later) until a gate mixes one of the
  // bits it touches; everything collected until then is applied by ONE micro-op.
  std::deque<PlanOp> synth;
  std::vector<const PlanOp*> merged_ops;
  // OPT-IN (QB_SIGNS=1): measured on B200 (profiles/r2o_*), the 32-qubit ansatz runs 1.88 s with it against 1.75 s without --
  // the sixteen sign flips per register group cost more issue slots than the two or three lone-phase dispatches they replace
  // (a lone CZ flips four); kept for circuits with long runs of +-1 gates per pass.
  if (env_int("QB_SIGNS", 0) && R == 4) {
    auto loc = [&](int pos) {  // 0 = register bit, 1 = thread (tile, non-register) bit, 2 = outside the tile
      if (!((sb.tile_mask >> pos) & 1)) return 2;
      return rbit_of_local[sb.local_of_pos[pos]] >= 0 ? 0 : 1;
    };
    auto is_pm1 = [](cd v) { return v == cd(1.0, 0.0) || v == cd(-1.0, 0.0); };
    // terms of a +-1 phase op, false when it is not one (or has a term MH_SIGNS cannot evaluate)
    auto sign_terms_of = [&](const PlanOp& p, std::vector<std::pair<int, int>>& terms, int& neg) {
      terms.clear();
      neg = 0;
      if (p.kind != CK_PHASE || p.cpos.size() > 1 || !is_pm1(p.scalar)) return false;
      for (auto& kv : p.fan)
        if (!is_pm1(kv.second.first) || !is_pm1(kv.second.second)) return false;
      const int a = p.cpos.empty() ? -1 : p.cpos[0];
      int toggles_a = p.scalar.real() < 0 ? 1 : 0;  // factors on the whole slice x_a = 1 (or on everything when a < 0)
      for (auto& kv : p.fan) {
        const bool f0 = kv.second.first.real() < 0, f1 = kv.second.second.real() < 0;
        if (f0) ++toggles_a;
        if (f0 != f1) terms.push_back(a < 0 ? std::make_pair(kv.first, -1) : std::make_pair(a, kv.first));
      }
      if (toggles_a & 1) {
        if (a < 0) neg = 1;
        else terms.push_back({a, -1});
      }
      for (auto& t : terms) {
        if (t.second < 0) continue;
        const int la_ = loc(t.first), lb_ = loc(t.second);
        if ((la_ == 2 && lb_ != 0) || (lb_ == 2 && la_ != 0)) return false;  // thread x outside, outside x outside
      }
      return true;
    };
    std::vector<const PlanOp*> members;
    std::vector<std::pair<int, int>> acc_terms;
    int acc_neg = 0, acc_thread_pairs = 0;
    uint64_t support = 0;
    auto flush = [&]() {
      if (members.size() >= 2) {
        PlanOp sp;
        sp.kind = CK_SIGNS;
        sp.sign_terms = acc_terms;
        sp.sign_neg = acc_neg;
        for (const PlanOp* m_ : members) sp.src.insert(sp.src.end(), m_->src.begin(), m_->src.end());
        synth.push_back(std::move(sp));
        merged_ops.push_back(&synth.back());
      } else {
        for (const PlanOp* m_ : members) merged_ops.push_back(m_);
      }
      members.clear();
      acc_terms.clear();
      acc_neg = 0;
      acc_thread_pairs = 0;
      support = 0;
    };
    std::vector<std::pair<int, int>> terms;
    for (const PlanOp* pp : ops) {
      int neg = 0;
      if (sign_terms_of(*pp, terms, neg)) {
        int tp = 0;
        for (auto& t : terms)
          if (t.second >= 0 && loc(t.first) == 1 && loc(t.second) == 1) ++tp;
        if (acc_thread_pairs + tp > SIGNS_MAX_PAIRS) flush();
        members.push_back(pp);
        acc_terms.insert(acc_terms.end(), terms.begin(), terms.end());
        acc_neg ^= neg;
        acc_thread_pairs += tp;
        for (auto& t : terms) support |= (uint64_t(1) << t.first) | (t.second >= 0 ? uint64_t(1) << t.second : 0);
        continue;
      }
      uint64_t x = 0;
      if (pp->kind == CK_DENSE || pp->kind == CK_SWAP)
        for (int t : pp->tpos) x |= uint64_t(1) << t;
      if (x & support) flush();
      merged_ops.push_back(pp);
    }
    flush();
  } else {
    merged_ops = ops;
  }
  for (const PlanOp* pp : merged_ops) {
    if (pp->kind == CK_SIGNS) {
      // G16: sign table over the register index; P[r]: thread bits whose value toggles the sign of register bit r's half;
      // zmask: thread bits that toggle everything; pairs: thread-bit pairs; E[r] / E[4]: the same from outside the tile
      MicroOp m;
      memset(&m, 0, sizeof(m));
      memset(m.tbit, 0xFF, sizeof(m.tbit));
      memset(m.rsel, 0xFF, sizeof(m.rsel));
      m.R = (uint16_t)R;
      m.slot = (uint16_t)MU_NO_SLOT;
      m.type = MU_SIGNS;
      m.handler = MH_SIGNS;
      uint32_t g16 = 0, zmask = 0;
      uint16_t P[4] = {0, 0, 0, 0};
      std::vector<uint16_t> pairs;
      auto regmask = [&](int r) {
        uint32_t mk = 0;
        for (int j = 0; j < (1 << R); ++j)
          if ((j >> r) & 1) mk |= 1u << j;
        return mk;
      };
      auto in_tile = [&](int pos) { return ((sb.tile_mask >> pos) & 1) != 0; };
      auto rb_of = [&](int pos) { return in_tile(pos) ? rbit_of_local[sb.local_of_pos[pos]] : -1; };  // -1: not a register bit
      for (auto t : pp->sign_terms) {
        int a = t.first, b = t.second;
        if (b < 0) {
          if (!in_tile(a)) m.ext_mask[4] ^= uint64_t(1) << a;
          else if (rb_of(a) >= 0) g16 ^= regmask(rb_of(a));
          else zmask ^= 1u << sb.local_of_pos[a];
          continue;
        }
        if (rb_of(a) < 0 && rb_of(b) >= 0) std::swap(a, b);  // a: the register bit when there is one
        if (rb_of(a) >= 0) {
          const int r = rb_of(a);
          if (!in_tile(b)) m.ext_mask[r] ^= uint64_t(1) << b;
          else if (rb_of(b) >= 0) g16 ^= regmask(r) & regmask(rb_of(b));
          else P[r] ^= (uint16_t)(1u << sb.local_of_pos[b]);
        } else {
          const uint16_t pm = (uint16_t)((1u << sb.local_of_pos[a]) | (1u << sb.local_of_pos[b]));
          auto it = std::find(pairs.begin(), pairs.end(), pm);
          if (it != pairs.end()) pairs.erase(it);  // the same pair twice cancels
          else pairs.push_back(pm);
        }
      }
      if ((int)pairs.size() > SIGNS_MAX_PAIRS) { err = "internal: too many thread-level pairs in a sign set"; return false; }
      m.creg = g16;
      m.cthr = zmask | (pp->sign_neg ? 0x80000000u : 0u);
      uint16_t* inl16 = reinterpret_cast<uint16_t*>(m.inl);
      for (int r = 0; r < 4; ++r) inl16[r] = P[r];
      inl16[4] = (uint16_t)pairs.size();
      for (size_t k = 0; k < pairs.size(); ++k) inl16[5 + k] = pairs[k];
      bool needs_slot = false;
      for (int e = 0; e < 5; ++e)
        if (m.ext_mask[e]) needs_slot = true;
      if (needs_slot) {
        m.slot = (uint16_t)sb.slots.size();
        sb.slots.push_back({pass_index, (int)mops.size()});
      }
      sb.payload_owner.push_back({pass_index, (int)mops.size()});
      sb.payloads.push_back({});
      mops.push_back(m);
      prev_op = pp;
      continue;
    }
    // A two-qubit controlled phase right after an uncontrolled one-qubit gate on one of its qubits (the tail of a QFT:
    // H(q) CU1(q, q+1)) is re-rooted as a one-entry fan controlled by that qubit, so that the pair fuses into a stage op
    // and the pass stays a straight-line stage pass.  (A pure phase is symmetric in its qubits.)
    PlanOp rerooted;
    bool force_fan = false;
    if (pp->kind == CK_PHASE && pp->fan.size() == 1 && pp->fan.begin()->second.first == cd(1.0, 0.0) && pp->scalar == cd(1.0, 0.0) &&
        pp->cpos.size() == 1 && prev_op &&
        prev_op->kind == CK_DENSE && prev_op->tpos.size() == 1 && prev_op->cpos.empty()) {
      const int d = prev_op->tpos[0], a = pp->cpos[0], b = pp->fan.begin()->first;
      if (d == a || d == b) {
        const int other = d == a ? b : a;
        const bool in_tile = ((sb.tile_mask >> d) & 1) && ((sb.tile_mask >> other) & 1);
        if (in_tile && rbit_of_local[sb.local_of_pos[d]] >= 0 && rbit_of_local[sb.local_of_pos[other]] >= 0 &&
            rbit_of_local[sb.local_of_pos[other]] < rbit_of_local[sb.local_of_pos[d]]) {
          rerooted = *pp;
          rerooted.cpos = {d};
          const cd phase = pp->fan.begin()->second.second;
          rerooted.fan.clear();
          rerooted.fan[other] = {cd(1.0, 0.0), phase};
          force_fan = true;
        }
      }
    }
    // A lone phase on the qubit of the one-qubit gate right before it (H(q) U1(q): the last stage of a QFT on a rank whose
    // global control qubits are 1) becomes a fan without factors controlled by that qubit, for the same reason.
    if (!force_fan && pp->kind == CK_PHASE && pp->fan.size() == 1 && pp->fan.begin()->second.first == cd(1.0, 0.0) &&
        pp->scalar == cd(1.0, 0.0) && pp->cpos.empty() &&
        prev_op && prev_op->kind == CK_DENSE && prev_op->tpos.size() == 1 && prev_op->cpos.empty() &&
        prev_op->tpos[0] == pp->fan.begin()->first && ((sb.tile_mask >> prev_op->tpos[0]) & 1) &&
        rbit_of_local[sb.local_of_pos[prev_op->tpos[0]]] >= 0) {
      rerooted = *pp;
      rerooted.cpos = {prev_op->tpos[0]};
      rerooted.scalar = pp->scalar * pp->fan.begin()->second.second;
      rerooted.fan.clear();
      force_fan = true;
    }
    prev_op = pp;
    const PlanOp& p = force_fan ? rerooted : *pp;
    MicroOp m;
    memset(&m, 0, sizeof(m));
    memset(m.tbit, 0xFF, sizeof(m.tbit));
    memset(m.rsel, 0xFF, sizeof(m.rsel));
    m.R = (uint16_t)R;
    m.slot = (uint16_t)MU_NO_SLOT;
    for (int c : p.cpos) {
      if (!((sb.tile_mask >> c) & 1)) m.ext_cmask |= uint64_t(1) << c;
      else if (rbit_of_local[sb.local_of_pos[c]] >= 0) m.creg |= 1u << rbit_of_local[sb.local_of_pos[c]];
      else m.cthr |= 1u << sb.local_of_pos[c];
    }
    bool needs_slot = false;
    std::vector<C> payload;
    if (p.kind == CK_DENSE || p.kind == CK_SWAP) {
      int k = (int)p.tpos.size();
      int rb[2] = {0, 0};
      for (int i = 0; i < k; ++i) {
        int lb = sb.local_of_pos[p.tpos[i]];
        if (lb < 0 || rbit_of_local[lb] < 0) { err = "internal: dense target is not a register bit"; return false; }
        rb[i] = rbit_of_local[lb];
      }
      if (p.kind == CK_SWAP) {
        m.type = MU_SWAP;
        m.handler = (uint8_t)(MH_SWAP + pair_index(std::max(rb[0], rb[1]), std::min(rb[0], rb[1])));
      } else if (k == 1) {
        m.type = MU_DENSE1;
        C* inl = reinterpret_cast<C*>(m.inl);
        for (int e = 0; e < 4; ++e) inl[e] = to_dev<C>(p.data[e]);
        const cd zero(0.0, 0.0), one(1.0, 0.0);
        if (p.special == 1) m.handler = (uint8_t)(MH_ADDSUB + rb[0]);
        else if (p.data[0] == zero && p.data[3] == zero && p.data[1] == one && p.data[2] == one) m.handler = (uint8_t)(MH_XPAIR + rb[0]);
        else if (m.creg != 0) m.handler = (uint8_t)(MH_CPLX1_M + rb[0]);
        else m.handler = (uint8_t)((is_real_matrix(p.data) ? MH_REAL1 : MH_CPLX1) + rb[0]);
      } else {
        m.type = MU_DENSE2;
        const bool flip = rb[0] < rb[1];  // the handler wants the matrix-index MSB on the higher register bit
        auto sw = [](int x) { return ((x & 1) << 1) | (x >> 1); };
        payload.resize(16);
        for (int r = 0; r < 4; ++r)
          for (int c = 0; c < 4; ++c) payload[(flip ? sw(r) : r) * 4 + (flip ? sw(c) : c)] = to_dev<C>(p.data[r * 4 + c]);
        m.handler = (uint8_t)(MH_DENSE2 + pair_index(std::max(rb[0], rb[1]), std::min(rb[0], rb[1])));
      }
    } else if (p.kind == CK_DIAG) {
      m.type = MU_DIAGK;
      m.handler = MH_DIAGK;
      needs_slot = true;
      int k = (int)p.tpos.size();
      m.k = k;
      for (int i = 0; i < k; ++i) {
        int pos = p.tpos[i];
        if (!((sb.tile_mask >> pos) & 1)) m.ext_mask[i] = uint64_t(1) << pos;
        else if (rbit_of_local[sb.local_of_pos[pos]] >= 0) m.rsel[i] = (uint8_t)rbit_of_local[sb.local_of_pos[pos]];
        else m.tbit[i] = (uint8_t)sb.local_of_pos[pos];
      }
      for (auto& v : p.data) payload.push_back(to_dev<C>(v));
    } else if (!force_fan && p.fan.size() == 1 && p.fan.begin()->second.first == cd(1.0, 0.0) && p.scalar == cd(1.0, 0.0)) {
      // a lone controlled phase (CZ, CU1, Z, T...): a factor on the slice where all of its bits are 1.  (PlanOp semantics:
      // phase(x) = [controls set] * scalar * prod_b f_b(x_b) -- with a non-unit scalar the slice where the fan bit is 0
      // gets a factor too, which only the general fan below applies)
      m.type = MU_PHASE;
      int pos = p.fan.begin()->first;
      if (!((sb.tile_mask >> pos) & 1)) m.ext_cmask |= uint64_t(1) << pos;
      else if (rbit_of_local[sb.local_of_pos[pos]] >= 0) m.creg |= 1u << rbit_of_local[sb.local_of_pos[pos]];
      else m.cthr |= 1u << sb.local_of_pos[pos];
      *reinterpret_cast<C*>(m.inl) = to_dev<C>(p.fan.begin()->second.second);
      if (m.creg == 0) m.handler = MH_PHASE_NC;
      else if (__builtin_popcount(m.creg) == 1) m.handler = (uint8_t)(MH_PHASE_C + __builtin_ctz(m.creg));
      else if (__builtin_popcount(m.creg) == 2 && !env_int("QB_NO_PHASE_C2", 0))
        m.handler = (uint8_t)(MH_PHASE_C2 + pair_index(31 - __builtin_clz(m.creg), __builtin_ctz(m.creg)));
      else m.handler = MH_PHASE_M;
    } else {  // fan
      m.type = MU_FAN;
      needs_slot = true;
      // MH_FAN_C + I shares one phase among the register indices that differ only above bit I: it needs a fan without
      // factors on those register bits
      bool above_free = true;
      if (__builtin_popcount(m.creg) == 1) {
        const int cb = __builtin_ctz(m.creg);
        for (auto& kv : p.fan) {
          if (!((sb.tile_mask >> kv.first) & 1)) continue;
          const int rb = rbit_of_local[sb.local_of_pos[kv.first]];
          if (rb > cb && (kv.second.first != cd(1.0, 0.0) || kv.second.second != cd(1.0, 0.0))) above_free = false;
        }
      }
      if (m.creg == 0) m.handler = MH_FAN_NC;
      else if (__builtin_popcount(m.creg) == 1 && above_free) m.handler = (uint8_t)(MH_FAN_C + __builtin_ctz(m.creg));
      else m.handler = MH_FAN_M;
      m.scalar[0] = p.scalar.real();
      m.scalar[1] = p.scalar.imag();
      m.la = (uint16_t)la;
      auto factor = [&](int pos, int bitval) { auto& pr = p.fan.at(pos); return bitval ? pr.second : pr.first; };
      // TA / TB over the group-index bits, G over the register bits
      for (int part = 0; part < 3; ++part) {
        int nb = part == 0 ? la : part == 1 ? gb - la : R;
        for (int v = 0; v < (1 << nb); ++v) {
          cd f(1.0, 0.0);
          for (auto& kv : p.fan) {
            if (!((sb.tile_mask >> kv.first) & 1)) continue;
            int lb = sb.local_of_pos[kv.first];
            int idx;
            if (part == 2) {
              idx = rbit_of_local[lb];
              if (idx < 0) continue;
            } else {
              int g = gbit_of_local[lb];
              if (g < 0) continue;
              if (part == 0 ? g >= la : g < la) continue;
              idx = part == 0 ? g : g - la;
            }
            f *= factor(kv.first, (v >> idx) & 1);
          }
          payload.push_back(to_dev<C>(f));
        }
      }
      std::vector<int> ext;
      for (auto& kv : p.fan)
        if (!((sb.tile_mask >> kv.first) & 1)) ext.push_back(kv.first);
      for (size_t s = 0; s < ext.size(); s += FAN_EXT_CHUNK) {
        size_t e = std::min(ext.size(), s + FAN_EXT_CHUNK);
        if (m.n_ext >= 8) { err = "internal: too many ext tables"; return false; }
        uint64_t mask = 0;
        for (size_t i = s; i < e; ++i) mask |= uint64_t(1) << ext[i];
        m.ext_mask[m.n_ext++] = mask;
        int nb = (int)(e - s);
        for (int v = 0; v < (1 << nb); ++v) {  // ext[] ascending == extract() order
          cd f(1.0, 0.0);
          for (int i = 0; i < nb; ++i) f *= factor(ext[s + i], (v >> i) & 1);
          payload.push_back(to_dev<C>(f));
        }
      }
    }
    if (m.ext_cmask != 0) needs_slot = true;
    if (needs_slot) {
      m.slot = (uint16_t)sb.slots.size();
      sb.slots.push_back({pass_index, (int)mops.size()});
    }
    sb.payload_owner.push_back({pass_index, (int)mops.size()});
    sb.payloads.push_back(std::move(payload));
    mops.push_back(m);
  }
  // ---- fuse "1-qubit gate on register bit I" + "fan controlled by bit I alone" into one stage op (one dispatch)
  if (!env_int("QB_NO_STAGE", 0)) {
    std::vector<MicroOp> fused;
    std::vector<int> new_index(mops.size(), -1);
    for (size_t q = 0; q < mops.size(); ++q) {
      const MicroOp& a = mops[q];
      const bool dense_ok = (a.handler >= MH_ADDSUB && a.handler < MH_ADDSUB + 4) || (a.handler >= MH_REAL1 && a.handler < MH_REAL1 + 4);
      if (dense_ok && a.cthr == 0 && a.slot == MU_NO_SLOT && q + 1 < mops.size()) {
        const int I = a.handler >= MH_REAL1 ? a.handler - MH_REAL1 : a.handler - MH_ADDSUB;
        const MicroOp& f = mops[q + 1];
        if (f.handler == MH_FAN_C + I && f.cthr == 0 && f.ext_cmask == 0) {
          MicroOp m = f;  // fan fields (tables, slot, la, ext masks) + the matrix of the dense gate
          memcpy(m.inl, a.inl, sizeof(m.inl));
          m.handler = (uint8_t)((a.handler >= MH_REAL1 ? MH_STAGE_R : MH_STAGE_A) + I);
          new_index[q] = new_index[q + 1] = (int)fused.size();
          fused.push_back(m);
          ++q;
          continue;
        }
      }
      new_index[q] = (int)fused.size();
      fused.push_back(a);
    }
    if (fused.size() != mops.size()) {
      // (by owner, not by position: a micro-op may own several payload entries once ops have been fused)
      for (auto& po : sb.payload_owner)
        if (po.first == pass_index && po.second >= 0) po.second = new_index[po.second];
      for (auto& so : sb.slots)
        if (so.first == pass_index && so.second >= 0) so.second = new_index[so.second];
      // a fused pair owns two payload entries (the dense gate's is empty): keep the fan's payload as the op's
      mops.swap(fused);
    }
  }
  // ---- fuse runs of uncontrolled real one-qubit gates on distinct register bits (they commute) into layer ops
  if (!env_int("QB_NO_LAYER", 0)) {
    std::vector<MicroOp> fused;
    std::vector<int> new_index(mops.size(), -1);
    std::vector<std::vector<C>> extra_payloads;  // payload of each new layer op
    std::vector<int> extra_owner;
    size_t q = 0;
    while (q < mops.size()) {
      auto is_real1 = [&](const MicroOp& m) { return m.handler >= MH_REAL1 && m.handler < MH_REAL1 + 4 && m.cthr == 0 && m.slot == MU_NO_SLOT; };
      if (!is_real1(mops[q])) {
        new_index[q] = (int)fused.size();
        fused.push_back(mops[q]);
        ++q;
        continue;
      }
      uint32_t mask = 0;
      size_t e = q;
      while (e < mops.size() && is_real1(mops[e]) && !((mask >> (mops[e].handler - MH_REAL1)) & 1)) {
        mask |= 1u << (mops[e].handler - MH_REAL1);
        ++e;
      }
      if (e - q < 2) {
        new_index[q] = (int)fused.size();
        fused.push_back(mops[q]);
        ++q;
        continue;
      }
      MicroOp m = mops[q];
      m.handler = MH_REAL_LAYER;
      m.k = mask;
      std::vector<C> pay(4 * R, to_dev<C>(cd(0.0, 0.0)));
      for (size_t t = q; t < e; ++t) {
        const int I = mops[t].handler - MH_REAL1;
        const C* inl = reinterpret_cast<const C*>(mops[t].inl);
        for (int x = 0; x < 4; ++x) pay[4 * I + x] = inl[x];
        new_index[t] = (int)fused.size();
      }
      extra_payloads.push_back(std::move(pay));
      extra_owner.push_back((int)fused.size());
      fused.push_back(m);
      q = e;
    }
    if (fused.size() != mops.size()) {
      // by owner: after the stage fusion above the pass no longer has one payload entry per micro-op, so positions
      // counted back from the end of the payload list would be off by the number of fused pairs
      for (auto& po : sb.payload_owner)
        if (po.first == pass_index && po.second >= 0) po.second = new_index[po.second];
      for (auto& so : sb.slots)
        if (so.first == pass_index && so.second >= 0) so.second = new_index[so.second];
      for (size_t x = 0; x < extra_payloads.size(); ++x) {  // the members' own payload entries are empty (inline matrices)
        sb.payload_owner.push_back({pass_index, extra_owner[x]});
        sb.payloads.push_back(std::move(extra_payloads[x]));
      }
      mops.swap(fused);
    }
  }
  ph.nmicro = (uint16_t)mops.size();
  if (!mops.empty() && !env_int("QB_NO_STAGE_PASS", 0)) {
    uint32_t mask = 0;
    int prev = 99;
    bool ok = true;
    bool any_stage = false;
    for (auto& m : mops) {
      int I = -1;
      if (m.handler >= MH_STAGE_A && m.handler < MH_STAGE_R + 4) {
        I = m.handler >= MH_STAGE_R ? m.handler - MH_STAGE_R : m.handler - MH_STAGE_A;
        any_stage = true;
      } else if (m.handler >= MH_ADDSUB && m.handler < MH_REAL1 + 4 && m.slot == MU_NO_SLOT) {
        I = (m.handler - MH_ADDSUB) & 3;  // a bare (a+b, a-b) / real 2x2 gate rides along (the last H of a QFT has no fan)
      }
      if (I < 0 || I >= prev || m.cthr != 0) { ok = false; break; }
      prev = I;
      mask |= 1u << I;
    }
    ph.stage_mask = ok && any_stage ? mask : 0u;
    const int Rfull = regtile_bits_for(sizeof(C) == 16 ? QB_C128 : QB_C64);
    if (ph.stage_mask && R == Rfull && Rfull == 4 && T == tile_bits_for(sizeof(C) == 16 ? QB_C128 : QB_C64) && !env_int("QB_NO_FULL_STAGE", 0)) {
      // which stages multiply by a real 2x2 matrix (MH_STAGE_R: the Hadamard that carries the sweep's scalars) instead of
      // the add/sub form; every op must be a fused stage with a slot and without controls outside the tile
      uint32_t real_mask = 0;
      bool full = true;
      for (auto& m : mops) {
        if (m.handler >= MH_STAGE_A && m.handler < MH_STAGE_A + 4) {
        } else if (m.handler >= MH_STAGE_R && m.handler < MH_STAGE_R + 4) {
          real_mask |= 1u << (m.handler - MH_STAGE_R);
        } else {
          full = false;
        }
        if (m.ext_cmask != 0 || m.slot == MU_NO_SLOT) full = false;
      }
      const uint32_t key = ph.stage_mask | (real_mask << 4);
      if (full && pass_full_stage_supported(key)) ph.flags |= PASS_FULL_STAGE | (real_mask << 4);
    }
  }
  if (pg) {
    // the last pass of a permuting sweep: its own thread -> group table (conflict-free on both layouts) and the
    // destination index of every slot's group
    std::vector<uint16_t> gt, dt;
    make_permuted_tables(T, R, (int)sizeof(C), rmask, pg->sigma, sswz, pg->dswizzle != 0, SWEEP_TEAM_THREADS, gt, dt);
    ph.flags |= PASS_PERMUTED_STORE | (pg->dswizzle ? PASS_PERMUTED_DSWZ : 0u);
    for (int i = 0; i < R; ++i) ph.dpos[i] = (uint8_t)pg->sigma[ph.pos[i]];
    sb.gtab_of_pass.push_back((int)sb.gtabs.size());
    sb.gtabs.push_back(std::move(gt));
    sb.dtab_of_pass.resize(sb.passes.size() + 1, -1);
    sb.dtab_of_pass[sb.passes.size()] = (int)sb.gtabs.size();
    sb.gtabs.push_back(std::move(dt));
  } else {
    if (paired_groups(T, R, (int)sizeof(C), rmask, sb.split_mask)) ph.flags |= PASS_PAIRED_GROUPS;
    auto it = sb.gtab_of_rmask.find(rmask);
    if (it == sb.gtab_of_rmask.end()) {
      it = sb.gtab_of_rmask.emplace(rmask, (int)sb.gtabs.size()).first;
      sb.gtabs.push_back(make_group_table(T, R, (int)sizeof(C), rmask, sb.split_mask));
    }
    sb.gtab_of_pass.push_back(it->second);
  }
  sb.passes.push_back(ph);
  sb.micro.push_back(std::move(mops));
  sb.big.push_back(DevOp());
  return true;
}

template <typename C> inline bool emit_big_pass(SweepBuilder<C>& sb, const PlanOp& p, std::string& err) {
  DevOp d;
  memset(&d, 0, sizeof(d));
  memset(d.tbit, 0xFF, sizeof(d.tbit));
  int k = (int)p.tpos.size();
  d.k = k;
  std::vector<int> ins;
  for (int i = 0; i < k; ++i) {
    if (!((sb.tile_mask >> p.tpos[i]) & 1)) { err = "internal: dense target outside tile"; return false; }
    d.tbit[i] = (uint8_t)sb.local_of_pos[p.tpos[i]];
    ins.push_back(sb.local_of_pos[p.tpos[i]]);
  }
  for (int c : p.cpos) {
    if ((sb.tile_mask >> c) & 1) {
      d.tl_cmask |= 1u << sb.local_of_pos[c];
      ins.push_back(sb.local_of_pos[c]);
    } else {
      d.ext_cmask |= uint64_t(1) << c;
    }
  }
  d.nins = (uint32_t)ins.size();
  for (int b : ins) d.ins_mask |= 1u << b;
  std::vector<C> payload;
  for (auto& v : p.data) payload.push_back(to_dev<C>(v));
  d.payload_global = payload.size() * sizeof(C) > BIG_PAYLOAD_SMEM_MAX ? 1u : 0u;
  d.slot = MU_NO_SLOT;
  if (d.ext_cmask != 0) {
    d.slot = (uint32_t)sb.slots.size();
    sb.slots.push_back({(int)sb.passes.size(), -1});
  }
  PassHeader ph;
  memset(&ph, 0, sizeof(ph));
  ph.kind = PASS_BIG;
  sb.payload_owner.push_back({(int)sb.passes.size(), -1});
  sb.payloads.push_back(std::move(payload));
  sb.gtab_of_pass.push_back(-1);
  sb.passes.push_back(ph);
  sb.micro.push_back({});
  sb.big.push_back(d);
  return true;
}

template <typename C> inline void finish_blob(SweepBuilder<C>& sb, SweepHeader& hdr, std::vector<char>& out, SweepDesc& sd) {
  size_t off = sizeof(SweepHeader);
  hdr.passes_offset = (uint32_t)off;
  off += align16(sb.passes.size() * sizeof(PassHeader));
  for (size_t p = 0; p < sb.passes.size(); ++p) {
    sb.passes[p].offset = (uint32_t)off;
    off += sb.passes[p].kind == PASS_BIG ? sizeof(DevOp) : sb.micro[p].size() * sizeof(MicroOp);
  }
  hdr.slots_offset = (uint32_t)off;
  off += align16(sb.slots.size() * sizeof(uint32_t));
  for (size_t s = 0; s < sb.payloads.size(); ++s) {
    auto own = sb.payload_owner[s];
    if (own.second < 0 && sb.big[own.first].payload_global) continue;
    if (sb.payloads[s].empty()) continue;  // (a fused stage op owns two entries; the dense gate's is empty)
    if (own.second < 0) sb.big[own.first].payload = (uint32_t)off;
    else sb.micro[own.first][own.second].payload = (uint32_t)off;
    off += align16(sb.payloads[s].size() * sizeof(C));
  }
  for (size_t p = 0; p < sb.passes.size(); ++p)
    for (auto& m : sb.micro[p])
      if (m.type == MU_FAN) {  // TA | TB | G follow each other in the fan's payload
        const int gb = sb.T - (int)m.R;
        m.k = m.payload + (uint32_t)((size_t(1) << m.la) * sizeof(C));
        m.gt_off = m.k + (uint32_t)((size_t(1) << (gb - (int)m.la)) * sizeof(C));
      }
  std::vector<uint32_t> gtab_off(sb.gtabs.size(), 0);
  for (size_t g = 0; g < sb.gtabs.size(); ++g) {
    gtab_off[g] = (uint32_t)off;
    off += align16(sb.gtabs[g].size() * sizeof(uint16_t));
  }
  for (size_t p = 0; p < sb.passes.size(); ++p) {
    if (sb.gtab_of_pass[p] >= 0) sb.passes[p].gtab = gtab_off[sb.gtab_of_pass[p]];
    if (p < sb.dtab_of_pass.size() && sb.dtab_of_pass[p] >= 0) sb.passes[p].dtab = gtab_off[sb.dtab_of_pass[p]];
  }
  hdr.npasses = (uint32_t)sb.passes.size();
  hdr.nslots = (uint32_t)sb.slots.size();
  hdr.blob_bytes = (uint32_t)off;  // the part the kernel copies to shared memory
  for (size_t s = 0; s < sb.payloads.size(); ++s) {  // large matrices follow; read through the global pointer
    auto own = sb.payload_owner[s];
    if (own.second < 0 && sb.big[own.first].payload_global) {
      sb.big[own.first].payload = (uint32_t)off;
      off += align16(sb.payloads[s].size() * sizeof(C));
    }
  }
  size_t start = align16(out.size());
  out.resize(start + off, 0);
  char* base = out.data() + start;
  memcpy(base, &hdr, sizeof(hdr));
  memcpy(base + hdr.passes_offset, sb.passes.data(), sb.passes.size() * sizeof(PassHeader));
  for (size_t p = 0; p < sb.passes.size(); ++p) {
    if (sb.passes[p].kind == PASS_BIG) memcpy(base + sb.passes[p].offset, &sb.big[p], sizeof(DevOp));
    else if (!sb.micro[p].empty()) memcpy(base + sb.passes[p].offset, sb.micro[p].data(), sb.micro[p].size() * sizeof(MicroOp));
  }
  for (size_t s = 0; s < sb.slots.size(); ++s) {
    auto own = sb.slots[s];
    uint32_t so = own.second < 0 ? (sb.passes[own.first].offset | 0x80000000u)
                                 : sb.passes[own.first].offset + (uint32_t)(own.second * sizeof(MicroOp));
    memcpy(base + hdr.slots_offset + s * sizeof(uint32_t), &so, sizeof(uint32_t));
  }
  for (size_t g = 0; g < sb.gtabs.size(); ++g) memcpy(base + gtab_off[g], sb.gtabs[g].data(), sb.gtabs[g].size() * sizeof(uint16_t));
  for (size_t s = 0; s < sb.payloads.size(); ++s) {
    auto own = sb.payload_owner[s];
    uint32_t po = own.second < 0 ? sb.big[own.first].payload : sb.micro[own.first][own.second].payload;
    if (!sb.payloads[s].empty()) memcpy(base + po, sb.payloads[s].data(), sb.payloads[s].size() * sizeof(C));
  }
  sd.blob_offset = start;
  sd.blob_bytes = off;
}

// Dimensions of the tensor map that moves a tile with ONE TMA copy (qb_sweep.cuh).  A dimension is (first bit, length):
// every run of consecutive tile bits is one (cut so that a box edge stays <= 256 8-byte elements; with the 128-byte
// swizzle the innermost one is exactly one 128-byte row), and ALL the bits outside the tile share a single "fold"
// dimension: its stride is the lowest non-tile bit and its extent everything above, so the coordinate base >> start
// addresses any tile (the tile bits of a tile base are zero; a TMA address is just base + sum coord_d * stride_d, the
// dimensions need not be disjoint).  Scattered tiles (random circuits, the wrap-around CZ of the ansatz) thus cost one
// dimension per run of TILE bits plus one, instead of one per run of either kind.  More than 5: no tensor map.
struct TileSeg { int start, len; bool tile; };
inline std::vector<TileSeg> tile_segments(int nqubits, int dtype, uint64_t tile_mask, bool swizzle) {
  const int row_bits = dtype == QB_C128 ? 3 : 4;  // amplitudes per 128-byte row
  std::vector<TileSeg> segs;
  bool folded = false;
  for (int b = 0; b < nqubits;) {
    const bool t = (tile_mask >> b) & 1;
    if (!t) {
      if (!folded) {
        segs.push_back({b, nqubits - b, false});
        folded = true;
      }
      while (b < nqubits && !((tile_mask >> b) & 1)) ++b;
      continue;
    }
    int e = b;
    const int cap = segs.empty() ? (swizzle ? row_bits : row_bits + 4) : 8;
    while (e < nqubits && ((tile_mask >> e) & 1) && e - b < cap) ++e;
    segs.push_back({b, e - b, true});
    b = e;
  }
  return segs;
}
inline bool tile_swizzle_ok(int nqubits, int dtype, uint64_t tile_mask) {
  if (env_int("QB_NO_TMA", 0) || env_int("QB_NO_SWIZZLE", 0)) return false;
  const int row_bits = dtype == QB_C128 ? 3 : 4;
  if ((int)__builtin_popcountll(tile_mask) < row_bits + 3) return false;  // at least eight full rows
  if ((tile_mask & ((uint64_t(1) << row_bits) - 1)) != ((uint64_t(1) << row_bits) - 1)) return false;
  return tile_segments(nqubits, dtype, tile_mask, true).size() <= 5;
}

// Tile of a permuting sweep for permutation `perm` on n bits: the `ls` lowest source bits (contiguous reads), the source
// bits that land on the `ld` lowest destination bits (contiguous writes), then the bits of `extra` in order (what the last
// gates of the program mix) while there is room, then more low bits of either side.  False when a side cannot get a
// swizzled tensor map.
inline bool perm_geometry(int n, int dtype, int T, const PermSpec& perm, int ls, int ld, const std::vector<int>& extra, PermGeom& g) {
  if (n <= T || ls + ld > T) return false;
  int inv[64];
  uint64_t seen = 0;
  for (int b = 0; b < n; ++b) {
    if (perm.pi[b] < 0 || perm.pi[b] >= n || ((seen >> perm.pi[b]) & 1)) return false;
    seen |= uint64_t(1) << perm.pi[b];
    inv[perm.pi[b]] = b;
  }
  uint64_t S = (uint64_t(1) << ls) - 1;
  for (int d = 0; d < ld; ++d) S |= uint64_t(1) << inv[d];
  for (int b : extra)
    if ((int)__builtin_popcountll(S) < T) S |= uint64_t(1) << b;
  for (int k = 0; (int)__builtin_popcountll(S) < T && k < n; ++k) {  // fill: next source bit, next destination bit, ...
    S |= uint64_t(1) << k;
    if ((int)__builtin_popcountll(S) < T) S |= uint64_t(1) << inv[k];
  }
  if ((int)__builtin_popcountll(S) != T) return false;
  uint64_t D = 0;
  for (int b = 0; b < n; ++b)
    if ((S >> b) & 1) D |= uint64_t(1) << perm.pi[b];
  if (!tile_swizzle_ok(n, dtype, S) || !tile_swizzle_ok(n, dtype, D)) return false;
  g.S = S;
  g.D = D;
  g.dswizzle = 1;
  int i = 0;
  for (int b = 0; b < n; ++b)
    if ((S >> b) & 1) g.sigma[i++] = (int)__builtin_popcountll(D & ((uint64_t(1) << perm.pi[b]) - 1));
  return true;
}

template <typename C>
inline bool build_plan(int n, int dtype, const std::vector<PlanOp>& pops, bool no_fuse, Plan& plan, std::string& err,
                       const Plan* replay = nullptr, const PermSpec* perm = nullptr) {
  const int Tfull = tile_bits_for(dtype);
  const int T = n < Tfull ? n : Tfull;
  // low bits every tile spans.  complex128 up to 30 qubits: 4 (256-byte rows; the tensor-map copy moves 128-byte
  // swizzle rows anyway), which leaves 8 free high bits = two full passes of four stages per QFT sweep (measured:
  // QFT(30) 33.0 -> 31.5 ms).  Beyond 32 GiB a tile of 256 short rows touches 256 distant pages and loses more than
  // the ninth pass costs (QFT(32): 172.7 ms with 4 low bits, 164.5 ms with 5).
  int Lcfg = env_int("QB_SWEEP_LOW_BITS", dtype == QB_C128 ? (n <= 30 ? 4 : 5) : 6);
  if (Lcfg > T) Lcfg = T;
  if (Lcfg < 1) Lcfg = 1;
  if (T - Lcfg > 8) Lcfg = T - 8;
  int R = regtile_bits_for(dtype);
  if (R > T) R = T;
  if (R < 1) R = 1;
  const int free_high = T - Lcfg;
  const int max_ops = no_fuse ? 1 : env_int("QB_SWEEP_MAX_OPS", 160);
  const uint64_t all = (uint64_t(1) << n) - 1;
  const uint64_t lowmask = (uint64_t(1) << Lcfg) - 1;
  const int csize = (int)sizeof(C);

  // ---- commutation bookkeeping: X = bits an op mixes (dense / swap targets), D = bits it only reads or multiplies
  // (controls, diagonal targets, fan bits).  Two ops commute iff neither one's X meets the other's X or D.
  const size_t N = pops.size();
  std::vector<uint64_t> xset(N, 0), dset(N, 0);
  for (size_t q = 0; q < N; ++q) {
    const PlanOp& p = pops[q];
    for (int c : p.cpos) dset[q] |= uint64_t(1) << c;
    if (p.kind == CK_DENSE || p.kind == CK_SWAP)
      for (int t : p.tpos) xset[q] |= uint64_t(1) << t;
    else if (p.kind == CK_DIAG)
      for (int t : p.tpos) dset[q] |= uint64_t(1) << t;
    else
      for (auto& kv : p.fan) dset[q] |= uint64_t(1) << kv.first;
  }
  if (replay && (replay->xsets != xset || replay->dsets != dset)) {
    err = "the gate structure changed";
    return false;
  }
  plan.xsets = xset;
  plan.dsets = dset;
  const bool reorder = !no_fuse && !env_int("QB_NO_REORDER", 0);
  const size_t window = (size_t)env_int("QB_REORDER_WINDOW", 4096);
  std::vector<char> done(N, 0);
  size_t ndone = 0, first = 0;
  // ---- a trailing permutation rides on the LAST sweep: the maximal suffix of ops that mix only bits of that sweep's tile
  // is held back for it (possibly no op at all: the sweep is then the permutation alone).  Candidate tiles trade row
  // length for room: f free tile bits next to 2^ls-amplitude source rows and 2^ld-amplitude destination rows take the f
  // bits the last gates mix; the candidate that leaves the fewest sweeps in front wins, longer rows first.
  PermGeom pgeom;
  bool fuse_perm = false;
  std::vector<size_t> held;
  if (perm != nullptr && replay) {
    if (replay->perm_fused && !replay->sweep_ops.empty()) {
      fuse_perm = true;
      pgeom = replay->pgeom;
      for (uint32_t q : replay->sweep_ops.back()) {
        if (q >= N) { err = "replay: bad op index"; return false; }
        held.push_back(q);
      }
    }
  } else if (perm != nullptr && !no_fuse && !env_int("QB_NO_FUSE_PERM", 0)) {
    auto hold = [&](const PermGeom& pg, std::vector<size_t>& out) {
      out.clear();
      size_t est = sizeof(SweepHeader) + 2 * sizeof(PassHeader) + 4096;
      int slot_est = 0;
      for (size_t q = N; q > 0; --q) {
        const PlanOp& p = pops[q - 1];
        if (xset[q - 1] & ~pg.S) break;
        est += blob_estimate(p, csize, T, R) + sizeof(PassHeader);
        slot_est += (p.kind == CK_PHASE || p.kind == CK_DIAG || !p.cpos.empty()) ? 1 : 0;
        if ((int)out.size() >= max_ops / 2 || est > (size_t)SWEEP_BLOB_MAX || slot_est > SWEEP_MAX_SLOTS) break;
        out.push_back(q - 1);
      }
      std::reverse(out.begin(), out.end());
      // leading diagonal ops belong to the gate in front of them (the fan of a QFT stage whose Hadamard stays behind)
      size_t lead = 0;
      while (lead < out.size() && xset[out[lead]] == 0) ++lead;
      out.erase(out.begin(), out.begin() + (long)lead);
    };
    int best_est = 1 << 30;
    const int fmax = env_int("QB_PERM_MAX_FREE", 4);
    for (int f = 0; f <= fmax && f <= T - 8; ++f) {
      const int ld = (T - f) / 2, ls = T - f - ld;
      // the f bits the last mixing gates need beyond the rows
      PermGeom base;
      if (!perm_geometry(n, dtype, T, *perm, ls, ld, {}, base)) continue;
      std::vector<int> extra;
      if (f > 0) {
        uint64_t have = (uint64_t(1) << ls) - 1;
        for (int b = 0; b < n; ++b)
          if (perm->pi[b] < ld) have |= uint64_t(1) << b;
        for (size_t q = N; q > 0 && (int)extra.size() < f; --q) {
          uint64_t miss = xset[q - 1] & ~have;
          if ((int)__builtin_popcountll(miss) + (int)extra.size() > f) break;
          for (int b = 0; miss; ++b)
            if ((miss >> b) & 1) {
              extra.push_back(b);
              have |= uint64_t(1) << b;
              miss &= ~(uint64_t(1) << b);
            }
        }
      }
      PermGeom cand;
      if (!perm_geometry(n, dtype, T, *perm, ls, ld, extra, cand)) continue;
      std::vector<size_t> h;
      hold(cand, h);
      // sweeps left in front: the high bits the other gates mix, free_high per sweep
      std::vector<char> is_held(N, 0);
      for (size_t q : h) is_held[q] = 1;
      uint64_t mixed = 0;
      for (size_t q = 0; q < N; ++q)
        if (!is_held[q]) mixed |= xset[q];
      const int hb = (int)__builtin_popcountll(mixed & ~lowmask);
      const int est = mixed == 0 ? 0 : std::max(1, (hb + std::max(free_high, 1) - 1) / std::max(free_high, 1));
      if (est < best_est) {
        best_est = est;
        pgeom = cand;
        held = h;
        fuse_perm = true;
      }
    }
  }
  if (fuse_perm) {
    for (size_t h : held) done[h] = 2;  // not available to the sweeps before
    ndone += held.size();
  }
  // ops a sweep may take: the blob estimate of the scan below cannot know the passes (one group table per distinct
  // register set, 1 KiB each for complex64), so a sweep whose program does not fit after all is planned again with
  // half as many ops
  enum { SW_OK = 0, SW_RETRY = 1, SW_FAIL = 2 };
  // one sweep from the ops `chosen` (program order) on the tile lowmask | high; `pg`: the permuting sweep
  auto emit_sweep = [&](const std::vector<size_t>& chosen, uint64_t high, const PermGeom* pg) -> int {
    // ---- complete the tile with the lowest unused bits (longest contiguous runs)
    uint64_t tile_mask = pg ? pg->S : (lowmask | high);
    for (int b = 0; b < n && __builtin_popcountll(tile_mask) < T; ++b) tile_mask |= uint64_t(1) << b;
    if (pg && tile_mask != pg->S) { err = "internal: permuting sweep tile mismatch"; return SW_FAIL; }
    SweepBuilder<C> sb;
    sb.T = T;
    sb.R = R;
    sb.tile_mask = tile_mask;
    sb.local_of_pos.assign(64, -1);
    {
      int lb = 0;
      for (int b = 0; b < n; ++b)
        if ((tile_mask >> b) & 1) sb.local_of_pos[b] = lb++;
    }
    int L = 0;
    while (L < n && ((tile_mask >> L) & 1)) ++L;
    SweepHeader hdr;
    memset(&hdr, 0, sizeof(hdr));
    hdr.T = T;
    hdr.L = L;
    hdr.R = R;
    hdr.tile_mask = tile_mask;
    hdr.other_mask = all & ~tile_mask;
    hdr.ntiles = uint64_t(1) << (n - T);
    hdr.swizzle = tile_swizzle_ok(n, dtype, tile_mask) ? 1u : 0u;
    SweepDesc sd;
    sd.swizzle = (int)hdr.swizzle;
    sd.T = T;
    sd.L = L;
    sd.tile_mask = tile_mask;
    sd.ntiles = hdr.ntiles;
    if (pg) {
      hdr.permuted = 1;
      hdr.dswizzle = pg->dswizzle ? 1u : 0u;
      hdr.dtile_mask = pg->D;
      for (int b = 0; b < n; ++b) hdr.dst_bit[b] = (uint8_t)perm->pi[b];
      sd.permuted = 1;
      sd.dswizzle = pg->dswizzle;
      sd.dmask = pg->D;
    }

    // ---- H-like gates: all but the last one of the sweep run as (a+b, a-b); the last one carries the product of
    // their scalars (a scalar commutes with everything).  Halves the FP work of those gates.
    std::vector<PlanOp> sops;
    sops.reserve(chosen.size());
    for (size_t q : chosen) sops.push_back(pops[q]);
    if (!no_fuse && !env_int("QB_NO_ADDSUB", 0)) {
      std::vector<size_t> cand;
      for (size_t q = 0; q < sops.size(); ++q)
        if (is_hadamard_like(sops[q])) cand.push_back(q);
      if (cand.size() >= 2) {
        cd prod(1.0, 0.0);
        for (size_t q : cand) prod *= sops[q].data[0];
        for (size_t c = 0; c + 1 < cand.size(); ++c) sops[cand[c]].special = 1;
        PlanOp& last = sops[cand.back()];
        last.data = {prod, prod, prod, -prod};
      }
    }

    // ---- split the sweep's ops into passes: a REGTILE pass holds ops whose dense targets fit R register bits
    std::vector<const PlanOp*> cur;
    uint32_t cur_r = 0;  // tile-local mask of the register bits demanded so far
    struct PassPlan { std::vector<const PlanOp*> ops; uint32_t need; const PlanOp* big; };
    std::vector<PassPlan> pplans;
    auto close_pass = [&]() -> bool {
      if (cur.empty()) return true;
      pplans.push_back({cur, cur_r, nullptr});
      cur.clear();
      cur_r = 0;
      return true;
    };
    // list scheduling again, one level down: a gate joins the open pass when it commutes with every gate of the sweep
    // that stays behind and the pass still has a register bit for it (a pass follows the light cone of its <= R qubits)
    {
      const size_t M = sops.size();
      std::vector<uint64_t> sx(M, 0), sd_(M, 0);
      for (size_t q = 0; q < M; ++q) {
        const PlanOp& p = sops[q];
        for (int s : p.src) plan.sweep_of_op[s] = (int)plan.sweeps.size();
        for (int c : p.cpos) sd_[q] |= uint64_t(1) << c;
        if (p.kind == CK_DENSE || p.kind == CK_SWAP)
          for (int t : p.tpos) sx[q] |= uint64_t(1) << t;
        else if (p.kind == CK_DIAG)
          for (int t : p.tpos) sd_[q] |= uint64_t(1) << t;
        else
          for (auto& kv : p.fan) sd_[q] |= uint64_t(1) << kv.first;
        if (p.kind == CK_PHASE || p.kind == CK_DIAG) ++sd.ndiag;
      }
      // balanced passes: when every tile bit is mixed by exactly one gate of the sweep (QFT-like: the passes partition the
      // stage bits), spread the bits evenly over the minimal number of passes -- 3 + 3 instead of 4 + 2 stages.  A pass whose
      // register bits are the FOUR lowest tile bits leaves only two conflict-free lane bits below the swizzle range
      // (ncu, permuting sweep of QFT(31): 32 % of the shared-memory wavefronts were bank conflicts with 4 + 2)
      int reg_cap = R;
      if (!no_fuse && !env_int("QB_NO_BALANCED_PASSES", 0)) {
        uint64_t mixed = 0;
        size_t nmix = 0;
        for (size_t q = 0; q < M; ++q)
          if (sx[q]) {
            mixed |= sx[q];
            ++nmix;
          }
        const int nbits = (int)__builtin_popcountll(mixed);
        if (nbits > R && nmix == (size_t)nbits) {
          const int npass = (nbits + R - 1) / R;
          reg_cap = (nbits + npass - 1) / npass;
        }
      }
      std::vector<char> placed(M, 0);
      size_t nplaced = 0, head = 0;
      while (nplaced < M) {
        while (head < M && placed[head]) ++head;
        uint64_t bx = 0, bd = 0;
        bool took_big = false;
        for (size_t q = head; q < M; ++q) {
          if (placed[q]) continue;
          const PlanOp& p = sops[q];
          bool ok = !((sx[q] & (bx | bd)) || (sd_[q] & bx));
          if (ok && p.kind == CK_DENSE && p.tpos.size() > 2) {
            if (cur.empty()) {  // a dense block on 3..6 targets is a pass of its own
              pplans.push_back({{}, 0u, &p});
              ++sd.npasses;
              placed[q] = 1;
              ++nplaced;
              took_big = true;
              break;
            }
            ok = false;
          } else if (ok) {
            uint32_t need = 0;
            if (p.kind == CK_DENSE || p.kind == CK_SWAP)
              for (int t : p.tpos) need |= 1u << sb.local_of_pos[t];
            if (__builtin_popcount(need) > R) { err = "internal: gate needs more register bits than a pass has"; return SW_FAIL; }
            if (__builtin_popcount(cur_r | need) > std::max(reg_cap, (int)__builtin_popcount(need))) {
              ok = false;
            } else {
              cur_r |= need;
              cur.push_back(&p);
              placed[q] = 1;
              ++nplaced;
            }
          }
          if (!ok) {
            if (!reorder) break;
            bx |= sx[q];
            bd |= sd_[q];
          }
        }
        if (took_big) continue;
        if (cur.empty()) { err = "internal: the pass scheduler made no progress"; return SW_FAIL; }
        ++sd.npasses;
        if (!close_pass()) return SW_FAIL;
      }
    }
    if (!close_pass()) return SW_FAIL;
    // ---- warp-private sub-tiles when three tile bits stay out of every pass's register set, then emit the passes
    {
      std::vector<uint32_t> needs;
      bool has_big = false;
      for (auto& pp : pplans) {
        if (pp.big) has_big = true;
        else needs.push_back(pp.need);
      }
      // (a permuting sweep's last pass stores across warps: team barriers)
      sb.split_mask = (no_fuse || pg) ? 0u : choose_split(T, R, csize, hdr.swizzle != 0, needs, has_big);
      hdr.warp_private = sb.split_mask ? 1u : 0u;
      // the LAST pass of a permuting sweep writes its groups in the destination layout: it must be a REGTILE pass (an
      // empty one when the sweep ends with a dense block, or has no gate at all)
      if (pg && (pplans.empty() || pplans.back().big)) {
        pplans.push_back({{}, 0u, nullptr});
        ++sd.npasses;
      }
      for (size_t pi_ = 0; pi_ < pplans.size(); ++pi_) {
        auto& pp = pplans[pi_];
        if (pp.big) {
          if (!emit_big_pass<C>(sb, *pp.big, err)) return SW_FAIL;
        } else {
          // pad the register set to R bits, preferring high tile-local bits (keeps lanes on the low bits)
          const bool last_permuted = pg && pi_ + 1 == pplans.size();
          if (!emit_regtile_pass<C>(sb, pp.ops, pad_register_bits(T, R, pp.need, sb.split_mask), err, last_permuted ? pg : nullptr, hdr.swizzle != 0))
            return SW_FAIL;
        }
      }
    }
    // (an empty pass -- the permuted store alone -- runs on either kernel instantiation)
    sd.stage_only = sb.passes.empty() ? 0 : 1;
    for (auto& ph_ : sb.passes)
      if (ph_.kind != PASS_REGTILE || (ph_.stage_mask == 0 && ph_.nmicro != 0)) sd.stage_only = 0;
    const size_t blob_before = plan.blob.size();
    const bool too_many_slots = (int)sb.slots.size() > SWEEP_MAX_SLOTS;
    if (!too_many_slots) finish_blob<C>(sb, hdr, plan.blob, sd);
    if (too_many_slots || hdr.blob_bytes > (size_t)SWEEP_BLOB_MAX + 1024) {
      if (replay) { err = "replay: sweep program no longer fits"; return SW_FAIL; }
      if (chosen.size() <= 1) {
        err = too_many_slots ? "internal: too many per-tile slots in one sweep" : "internal: sweep program too large";
        return SW_FAIL;
      }
      plan.blob.resize(blob_before);  // plan this sweep again with fewer ops
      return SW_RETRY;
    }
    plan.npasses += sd.npasses;
    plan.ndiag += sd.ndiag;
    if (env_int("QB_PLAN_DEBUG", 0))
      fprintf(stderr, "[qb plan] sweep %zu: ops %zu passes %d tile %#llx L %d swizzle %u warp_private %u split %#x blob %u B slots %zu stage_only %d\n", plan.sweeps.size(),
              sops.size(), sd.npasses, (unsigned long long)tile_mask, L, hdr.swizzle, hdr.warp_private, sb.split_mask, hdr.blob_bytes, sb.slots.size(), sd.stage_only);
    if (env_int("QB_PLAN_DEBUG", 0) > 1)
      for (size_t p_ = 0; p_ < sb.passes.size(); ++p_) {
        fprintf(stderr, "[qb plan]   pass %zu kind %u rmask %#x stage_mask %#x flags %u handlers:", p_, sb.passes[p_].kind, sb.passes[p_].rmask, sb.passes[p_].stage_mask, sb.passes[p_].flags);
        for (auto& m_ : sb.micro[p_]) fprintf(stderr, " %d", (int)m_.handler);
        fprintf(stderr, "\n");
      }
    plan.sweeps.push_back(sd);
    plan.sweep_ops.emplace_back(chosen.begin(), chosen.end());
    return SW_OK;
  };
  int cur_max_ops = max_ops;
  while (ndone < N) {
    while (first < N && done[first]) ++first;
    // ---- list scheduling: walk the remaining ops in program order; an op joins this sweep when it commutes with
    // every earlier op that stays behind and its dense targets fit the tile.  (Layered circuits: a sweep follows
    // the light cone of its tile's qubits through many layers instead of stopping at the first gate outside.)
    // `allowed`: the high tile bits a candidate tile may use (~0 = whatever the first ready gates ask for)
    uint64_t high = 0;
    std::vector<size_t> chosen;
    bool fatal = false;
    auto scan = [&](uint64_t allowed, std::vector<size_t>& out, uint64_t& out_high, int& out_dense) {
      uint64_t hi = 0, blocked_x = 0, blocked_d = 0;
      size_t est = sizeof(SweepHeader) + 64;
      int slot_est = 0;
      out.clear();
      out_dense = 0;
      for (size_t q = first; q < N && q < first + window; ++q) {
        if (done[q]) continue;
        const PlanOp& p = pops[q];
        bool ok = !((xset[q] & (blocked_x | blocked_d)) || (dset[q] & blocked_x));
        if (ok) {
          const uint64_t nh = hi | (xset[q] & ~lowmask);
          const size_t add = blob_estimate(p, csize, T, R) + sizeof(PassHeader);
          const int slot_add = (p.kind == CK_PHASE || p.kind == CK_DIAG || !p.cpos.empty()) ? 1 : 0;  // may need per-tile set-up
          if (__builtin_popcountll(nh) > free_high || (nh & ~allowed)) {
            if (allowed == ~uint64_t(0) && out.empty() && q == first) { err = "gate has more target qubits outside the low bits than a tile can hold"; fatal = true; return; }
            ok = false;
          } else if (!out.empty() && (est + add > (size_t)SWEEP_BLOB_MAX || (int)out.size() >= cur_max_ops || slot_est + slot_add > SWEEP_MAX_SLOTS)) {
            ok = false;
            if ((int)out.size() >= cur_max_ops) break;
          } else if (est + add > (size_t)SWEEP_BLOB_MAX) {
            err = "single gate does not fit the sweep program buffer";
            fatal = true;
            return;
          } else {
            hi = nh;
            est += add;
            slot_est += slot_add;
            out.push_back(q);
            if (xset[q]) ++out_dense;
          }
        }
        if (!ok) {
          if (!reorder) break;
          blocked_x |= xset[q];
          blocked_d |= dset[q];
          if ((blocked_x & all) == all) break;
        }
      }
      out_high = hi;
    };
    int best_dense = 0;
    if (replay) {
      // the schedule of the program this one re-parametrises: same ops per sweep, no scan
      const size_t si = plan.sweeps.size();
      if (si >= replay->sweep_ops.size()) { err = "replay: sweep count mismatch"; return false; }
      for (uint32_t q : replay->sweep_ops[si]) {
        if (q >= N || done[q]) { err = "replay: bad op index"; return false; }
        chosen.push_back(q);
        high |= xset[q] & ~lowmask;
      }
    } else {
      scan(~uint64_t(0), chosen, high, best_dense);
    }
    if (fatal) return false;
    if (!replay && reorder && free_high > 0 && n > T) {
      // candidate tiles: the low bits plus a window of contiguous higher bits (widest light cone); keep the one that
      // takes the most mixing gates
      const int step = N <= 5000 ? 1 : 3;
      std::vector<size_t> cand;
      for (int s0 = Lcfg; s0 + free_high <= n; s0 += step) {
        const uint64_t allowed = ((uint64_t(1) << free_high) - 1) << s0;
        uint64_t chigh = 0;
        int cdense = 0;
        scan(allowed, cand, chigh, cdense);
        if (fatal) return false;
        if (cdense > best_dense) {
          best_dense = cdense;
          chosen = cand;
          high = chigh;
        }
      }
    }
    if (chosen.empty()) { err = "internal: the sweep scheduler made no progress"; return false; }
    const int rc = emit_sweep(chosen, high, nullptr);
    if (rc == SW_FAIL) return false;
    if (rc == SW_RETRY) {
      cur_max_ops = (int)chosen.size() / 2;
      continue;
    }
    for (size_t q : chosen) done[q] = 1;
    ndone += chosen.size();
    cur_max_ops = max_ops;
  }
  if (fuse_perm) {
    const int rc = emit_sweep(held, 0, &pgeom);
    if (rc != SW_OK) {
      if (rc == SW_RETRY) err = "the ops held back for the permuting sweep do not fit one sweep program";
      return false;  // (the caller plans again without the permutation)
    }
    plan.perm_fused = 1;
    plan.pgeom = pgeom;
  }
  return true;
}

// `replay`: the plan of a program with the same gate structure (targets, controls, which gates are diagonal): its
// schedule is reused and only the passes are emitted again with the new numbers; false + err when the structure differs
// (e.g. a rotation angle became 0 and the gate an identity) -- the caller then plans from scratch.
// `perm`: a qubit permutation to apply after the ops.  plan.perm_fused tells whether it rides on the last sweep (which then
// writes OUT OF PLACE, launch_sweep's `dst`); otherwise the plan is the plain one and the caller runs K8 after it.
inline bool plan_program(int n, int dtype, const std::vector<CanonOp>& ops, bool no_fuse, Plan& plan, std::string& err,
                         const Plan* replay = nullptr, const PermSpec* perm = nullptr) {
  std::vector<PlanOp> pops;
  merge_ops(ops, no_fuse, pops);
  for (int attempt = 0; attempt < 2; ++attempt) {
    plan = Plan();
    plan.sweep_of_op.assign(ops.size(), -1);
    const PermSpec* pm = attempt == 0 ? perm : nullptr;
    if (replay && (replay->perm_fused != 0) != (pm != nullptr)) {
      if (replay->perm_fused) { err = "replay: the program was planned with a fused permutation"; return false; }
      pm = nullptr;
    }
    const bool ok = dtype == QB_C128 ? build_plan<d2>(n, dtype, pops, no_fuse, plan, err, replay, pm)
                                     : build_plan<f2>(n, dtype, pops, no_fuse, plan, err, replay, pm);
    if (ok) return true;
    if (pm == nullptr || replay) return false;  // (a failed fusion: plan again without the permutation)
  }
  return false;
}

inline void fill_stats(const Plan& plan, int n, int dtype, int nops, qb_program_stats* st) {
  memset(st, 0, sizeof(*st));
  st->nops = nops;
  st->nsweeps = (int)plan.sweeps.size();
  st->ndense_passes = plan.npasses;
  st->ndiag_ops = plan.ndiag;
  st->bytes_moved = (double)plan.sweeps.size() * 2.0 * (dtype == QB_C128 ? 16.0 : 8.0) * (double)(uint64_t(1) << n);
  for (auto& sd : plan.sweeps) st->nstage_sweeps += sd.stage_only ? 1 : 0;
  st->perm_fused = plan.perm_fused;
}

}  // namespace qb
