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
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/qibo_b200.h"
#include "qb_canon.hpp"

namespace qb {

// ---- device program layout (shared with qb_sweep.cuh) ---------------------------------------------
enum DevOpType { OP_DENSE = 1, OP_SWAP = 2, OP_FAN = 3, OP_DIAGK = 4, OP_DENSE_BIG = 5 };

constexpr int SWEEP_MAX_OPS = 48;
constexpr int SWEEP_BLOB_MAX = 30 * 1024;  // program bytes resident in shared memory next to the tiles
constexpr int SWEEP_TILE_BYTES_LOG2 = 16;  // 64 KiB tiles, three in flight per SM

struct DevOp {
  uint32_t type;
  uint32_t k;          // number of target bits
  uint32_t nins;       // DENSE/SWAP: tile-local bits to insert (targets + tile-local controls), sorted in ins[]
  uint32_t tl_cmask;   // controls inside the tile (tile-local bit mask)
  uint64_t ext_cmask;  // controls outside the tile (state bit positions)
  uint32_t payload;    // byte offset inside the blob of the matrix / tables
  uint32_t n_ext;      // FAN: number of ext tables
  uint8_t tbit[8];     // tile-local bit of target i (tbit[0] = MSB of the matrix index); 0xFF = outside the tile (DIAGK)
  uint8_t ins[16];
  uint8_t chunk_lo[4];   // FAN: tile-local chunk c covers bits [chunk_lo[c], chunk_lo[c] + chunk_len[c])
  uint8_t chunk_len[4];
  uint32_t n_chunks;
  uint32_t ins_mask;     // DENSE/SWAP: the same insert positions as a tile-local bit mask
  uint64_t ext_mask[6];  // FAN: state-bit mask of ext table e; DIAGK: single-bit mask of ext target i
  double scalar[2];      // FAN: global factor
  uint64_t pad2;
};
static_assert(sizeof(DevOp) % 16 == 0, "DevOp must stay 16-byte aligned");

struct SweepHeader {
  uint32_t nops;
  uint32_t T;           // log2(amplitudes per tile)
  uint32_t L;           // log2(amplitudes per contiguous run)
  uint32_t blob_bytes;
  uint64_t tile_mask;   // state bits spanned by the tile
  uint64_t other_mask;  // remaining state bits (enumerated by the tile index)
  uint64_t ntiles;
  uint32_t ops_offset;  // byte offset of DevOp[0]
  uint32_t pad[5];
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
};

struct SweepDesc {
  int T = 0, L = 0;
  uint64_t tile_mask = 0;
  size_t blob_offset = 0, blob_bytes = 0;
  int npasses = 0, ndiag = 0;
  uint64_t ntiles = 0;
};

struct Plan {
  std::vector<SweepDesc> sweeps;
  std::vector<char> blob;
  std::vector<int> sweep_of_op;
  int npasses = 0, ndiag = 0;
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
        if (P.size() == f.cpos.size() + 1 && std::includes(P.begin(), P.end(), f.cpos.begin(), f.cpos.end())) {
          int b = -1;
          for (int x : P)
            if (!std::binary_search(f.cpos.begin(), f.cpos.end(), x)) b = x;
          auto it = f.fan.find(b);
          if (it == f.fan.end()) f.fan[b] = entry;
          else it->second.second *= entry.second;
          merged = true;
        } else if (f.fan.size() == 1 && f.fan.begin()->second.first == cd(1.0, 0.0)) {
          // re-root a single-entry fan so that the shared bits become the controls
          std::vector<int> P0 = fan_set(f), common;
          std::set_intersection(P.begin(), P.end(), P0.begin(), P0.end(), std::back_inserter(common));
          if (P.size() == P0.size() && common.size() + 1 == P0.size()) {
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

// ---- step 2 + 3: pack into sweeps and serialise --------------------------------------------------------
template <typename C> inline C to_dev(cd v);
struct f2 { float x, y; };
struct d2 { double x, y; };
template <> inline f2 to_dev<f2>(cd v) { return f2{(float)v.real(), (float)v.imag()}; }
template <> inline d2 to_dev<d2>(cd v) { return d2{v.real(), v.imag()}; }

inline size_t align16(size_t x) { return (x + 15) & ~size_t(15); }

template <typename C> struct BlobBuilder {
  std::vector<char> bytes;
  std::vector<DevOp> ops;
  std::vector<std::vector<C>> payloads;
  size_t payload_bytes = 0;
  size_t size_with(size_t extra_ops, size_t extra_payload) const {
    return sizeof(SweepHeader) + (ops.size() + extra_ops) * sizeof(DevOp) + payload_bytes + extra_payload;
  }
};

// bytes of payload a PlanOp needs for a given tile (upper bound, independent of the tile)
inline size_t payload_estimate(const PlanOp& p, int csize, int T) {
  switch (p.kind) {
    case CK_DENSE: return align16((size_t)p.data.size() * csize);
    case CK_SWAP: return 0;
    case CK_DIAG: return align16((size_t)p.data.size() * csize);
    default: {  // fan: two local chunk tables + ext tables of <= 5 bits
      size_t local = ((size_t(1) << ((T + 1) / 2)) + (size_t(1) << (T / 2))) * csize;
      size_t next = (p.fan.size() + 4) / 5;
      return align16(local + next * 32 * csize);
    }
  }
}

template <typename C>
inline bool emit_op(const PlanOp& p, uint64_t tile_mask, int T, const std::vector<int>& local_of_pos, BlobBuilder<C>& bb,
                    std::string& err) {
  DevOp d;
  memset(&d, 0, sizeof(d));
  memset(d.tbit, 0xFF, sizeof(d.tbit));
  uint64_t cm = 0;
  for (int c : p.cpos) cm |= uint64_t(1) << c;
  d.ext_cmask = cm & ~tile_mask;
  for (int c : p.cpos)
    if ((tile_mask >> c) & 1) d.tl_cmask |= 1u << local_of_pos[c];
  std::vector<C> payload;
  if (p.kind == CK_DENSE || p.kind == CK_SWAP) {
    int k = (int)p.tpos.size();
    d.k = k;
    d.type = p.kind == CK_SWAP ? OP_SWAP : (k <= 2 ? OP_DENSE : OP_DENSE_BIG);
    std::vector<int> ins;
    for (int i = 0; i < k; ++i) {
      if (!((tile_mask >> p.tpos[i]) & 1)) { err = "internal: dense target outside tile"; return false; }
      d.tbit[i] = (uint8_t)local_of_pos[p.tpos[i]];
      ins.push_back(local_of_pos[p.tpos[i]]);
    }
    for (int c : p.cpos)
      if ((tile_mask >> c) & 1) ins.push_back(local_of_pos[c]);
    std::sort(ins.begin(), ins.end());
    if (ins.size() > 16) { err = "too many tile-local controls"; return false; }
    d.nins = (uint32_t)ins.size();
    for (size_t i = 0; i < ins.size(); ++i) {
      d.ins[i] = (uint8_t)ins[i];
      d.ins_mask |= 1u << ins[i];
    }
    if (p.kind == CK_DENSE)
      for (auto& v : p.data) payload.push_back(to_dev<C>(v));
  } else if (p.kind == CK_DIAG) {
    int k = (int)p.tpos.size();
    d.type = OP_DIAGK;
    d.k = k;
    for (int i = 0; i < k; ++i) {
      if ((tile_mask >> p.tpos[i]) & 1) d.tbit[i] = (uint8_t)local_of_pos[p.tpos[i]];
      else d.ext_mask[i] = uint64_t(1) << p.tpos[i];
    }
    for (auto& v : p.data) payload.push_back(to_dev<C>(v));
  } else {  // fan
    d.type = OP_FAN;
    d.scalar[0] = p.scalar.real();
    d.scalar[1] = p.scalar.imag();
    // local chunks: two contiguous ranges of tile-local bits
    int len0 = (T + 1) / 2, len1 = T - len0;
    int los[2] = {0, len0}, lens[2] = {len0, len1};
    d.n_chunks = 0;
    for (int c = 0; c < 2; ++c) {
      if (lens[c] == 0) continue;
      bool any = false;
      for (auto& kv : p.fan)
        if (((tile_mask >> kv.first) & 1) && local_of_pos[kv.first] >= los[c] && local_of_pos[kv.first] < los[c] + lens[c]) any = true;
      if (!any) continue;
      int ci = d.n_chunks++;
      d.chunk_lo[ci] = (uint8_t)los[c];
      d.chunk_len[ci] = (uint8_t)lens[c];
      for (int v = 0; v < (1 << lens[c]); ++v) {
        cd f(1.0, 0.0);
        for (auto& kv : p.fan) {
          if (!((tile_mask >> kv.first) & 1)) continue;
          int lb = local_of_pos[kv.first];
          if (lb < los[c] || lb >= los[c] + lens[c]) continue;
          f *= ((v >> (lb - los[c])) & 1) ? kv.second.second : kv.second.first;
        }
        payload.push_back(to_dev<C>(f));
      }
    }
    // ext tables: groups of <= 5 outside bits
    std::vector<int> ext;
    for (auto& kv : p.fan)
      if (!((tile_mask >> kv.first) & 1)) ext.push_back(kv.first);
    d.n_ext = 0;
    for (size_t s = 0; s < ext.size(); s += 5) {
      size_t e = std::min(ext.size(), s + 5);
      if (d.n_ext >= 6) { err = "internal: too many ext tables"; return false; }
      uint64_t mask = 0;
      for (size_t i = s; i < e; ++i) mask |= uint64_t(1) << ext[i];
      d.ext_mask[d.n_ext++] = mask;
      int nb = (int)(e - s);
      for (int v = 0; v < (1 << nb); ++v) {  // ext[] ascending == extract() order
        cd f(1.0, 0.0);
        for (int i = 0; i < nb; ++i) {
          auto& pr = p.fan.at(ext[s + i]);
          f *= ((v >> i) & 1) ? pr.second : pr.first;
        }
        payload.push_back(to_dev<C>(f));
      }
    }
  }
  bb.ops.push_back(d);
  bb.payload_bytes += align16(payload.size() * sizeof(C));
  bb.payloads.push_back(std::move(payload));
  return true;
}

template <typename C> inline void finish_blob(BlobBuilder<C>& bb, SweepHeader& hdr, std::vector<char>& out, SweepDesc& sd) {
  size_t off = sizeof(SweepHeader);
  hdr.ops_offset = (uint32_t)off;
  off += bb.ops.size() * sizeof(DevOp);
  for (size_t i = 0; i < bb.ops.size(); ++i) {
    bb.ops[i].payload = (uint32_t)off;
    off += align16(bb.payloads[i].size() * sizeof(C));
  }
  hdr.nops = (uint32_t)bb.ops.size();
  hdr.blob_bytes = (uint32_t)off;
  size_t start = align16(out.size());
  out.resize(start + off, 0);
  char* base = out.data() + start;
  memcpy(base, &hdr, sizeof(hdr));
  memcpy(base + hdr.ops_offset, bb.ops.data(), bb.ops.size() * sizeof(DevOp));
  for (size_t i = 0; i < bb.ops.size(); ++i)
    if (!bb.payloads[i].empty()) memcpy(base + bb.ops[i].payload, bb.payloads[i].data(), bb.payloads[i].size() * sizeof(C));
  sd.blob_offset = start;
  sd.blob_bytes = off;
}

template <typename C>
inline bool build_plan(int n, int dtype, const std::vector<PlanOp>& pops, bool no_fuse, Plan& plan, std::string& err) {
  const int Tfull = tile_bits_for(dtype);
  const int T = n < Tfull ? n : Tfull;
  int Lcfg = env_int("QB_SWEEP_LOW_BITS", dtype == QB_C128 ? 5 : 6);
  if (Lcfg > T) Lcfg = T;
  if (Lcfg < 1) Lcfg = 1;
  const int free_high = T - Lcfg;
  const int max_passes = no_fuse ? 1 : env_int("QB_SWEEP_MAX_PASSES", 6);
  const uint64_t all = (uint64_t(1) << n) - 1;
  const uint64_t lowmask = (uint64_t(1) << Lcfg) - 1;
  const int csize = (int)sizeof(C);

  size_t i = 0;
  while (i < pops.size()) {
    // ---- greedy: take ops while their dense targets fit the tile
    uint64_t high = 0;
    size_t j = i;
    int passes = 0;
    size_t est = sizeof(SweepHeader);
    while (j < pops.size()) {
      const PlanOp& p = pops[j];
      uint64_t need = 0;
      if (p.kind == CK_DENSE || p.kind == CK_SWAP)
        for (int t : p.tpos) need |= uint64_t(1) << t;
      need &= ~lowmask;
      uint64_t nh = high | need;
      if (__builtin_popcountll(nh) > free_high) {
        if (j == i) { err = "gate has more target qubits outside the low bits than a tile can hold"; return false; }
        break;
      }
      size_t add = sizeof(DevOp) + payload_estimate(p, csize, T);
      if (j > i && (est + add > (size_t)SWEEP_BLOB_MAX || passes + 1 > max_passes || (int)(j - i) >= SWEEP_MAX_OPS)) break;
      if (est + add > (size_t)SWEEP_BLOB_MAX) { err = "single gate does not fit the sweep program buffer"; return false; }
      high = nh;
      est += add;
      ++passes;
      ++j;
    }
    // ---- complete the tile with the lowest unused bits (longest contiguous runs)
    uint64_t tile_mask = lowmask | high;
    for (int b = 0; b < n && __builtin_popcountll(tile_mask) < T; ++b) tile_mask |= uint64_t(1) << b;
    std::vector<int> local_of_pos(64, -1);
    {
      int lb = 0;
      for (int b = 0; b < n; ++b)
        if ((tile_mask >> b) & 1) local_of_pos[b] = lb++;
    }
    int L = 0;
    while (L < n && ((tile_mask >> L) & 1)) ++L;
    SweepHeader hdr;
    memset(&hdr, 0, sizeof(hdr));
    hdr.T = T;
    hdr.L = L;
    hdr.tile_mask = tile_mask;
    hdr.other_mask = all & ~tile_mask;
    hdr.ntiles = uint64_t(1) << (n - T);
    BlobBuilder<C> bb;
    SweepDesc sd;
    sd.T = T;
    sd.L = L;
    sd.tile_mask = tile_mask;
    sd.ntiles = hdr.ntiles;
    for (size_t q = i; q < j; ++q) {
      if (!emit_op<C>(pops[q], tile_mask, T, local_of_pos, bb, err)) return false;
      if (pops[q].kind == CK_DENSE || pops[q].kind == CK_SWAP) ++sd.npasses;
      else ++sd.ndiag;
      for (int s : pops[q].src) plan.sweep_of_op[s] = (int)plan.sweeps.size();
    }
    finish_blob<C>(bb, hdr, plan.blob, sd);
    if (sd.blob_bytes > (size_t)SWEEP_BLOB_MAX + 2048) { err = "internal: sweep program too large"; return false; }
    plan.npasses += sd.npasses;
    plan.ndiag += sd.ndiag;
    plan.sweeps.push_back(sd);
    i = j;
  }
  return true;
}

inline bool plan_program(int n, int dtype, const std::vector<CanonOp>& ops, bool no_fuse, Plan& plan, std::string& err) {
  plan = Plan();
  plan.sweep_of_op.assign(ops.size(), -1);
  std::vector<PlanOp> pops;
  merge_ops(ops, no_fuse, pops);
  if (dtype == QB_C128) return build_plan<d2>(n, dtype, pops, no_fuse, plan, err);
  return build_plan<f2>(n, dtype, pops, no_fuse, plan, err);
}

inline void fill_stats(const Plan& plan, int n, int dtype, int nops, qb_program_stats* st) {
  memset(st, 0, sizeof(*st));
  st->nops = nops;
  st->nsweeps = (int)plan.sweeps.size();
  st->ndense_passes = plan.npasses;
  st->ndiag_ops = plan.ndiag;
  st->bytes_moved = (double)plan.sweeps.size() * 2.0 * (dtype == QB_C128 ? 16.0 : 8.0) * (double)(uint64_t(1) << n);
}

}  // namespace qb
