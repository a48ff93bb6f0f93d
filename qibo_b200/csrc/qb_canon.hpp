// qibo_b200: host-side gate canonicalisation.
//
// The reference applies named controlled gates (CNOT, CZ, CU1, TOFFOLI...) as their FULL matrix over
// gate.qubits (abstract.py:2339-2355), and `controlled_by` gates through the control
// slice path (abstract.py:3176-3197).  Both are "controls + small target matrix".  This pass recovers
// that structure exactly (entries are compared with 0 and 1 exactly -- gate tables are built from
// literals), so that kernels touch only the amplitudes that can change:
//   DENSE  : 2^k x 2^k matrix on k target bits, c control bits
//   DIAG   : 2^k diagonal on k target bits, c control bits
//   PHASE  : a scalar on the slice where all control bits are 1 (CZ, CU1, Z, S, T, U1, CCZ...)
//   SWAP   : exchange of the 01 / 10 amplitudes of two bits, c control bits
//   NOOP   : identity
#pragma once
#include <algorithm>
#include <complex>
#include <string>
#include <vector>

namespace qb {

enum CanonKind { CK_NOOP = 0, CK_DENSE = 1, CK_DIAG = 2, CK_PHASE = 3, CK_SWAP = 4 };

typedef std::complex<double> cd;

struct CanonOp {
  int kind = CK_NOOP;
  std::vector<int> tpos;   // target bit positions; tpos[0] is the MSB of the matrix index
  std::vector<int> cpos;   // control bit positions (all must be 1), ascending
  std::vector<cd> data;    // DENSE: 2^k*2^k row-major; DIAG: 2^k; PHASE: 1 entry
  uint64_t cmask() const {
    uint64_t m = 0;
    for (int p : cpos) m |= uint64_t(1) << p;
    return m;
  }
  uint64_t tmask() const {
    uint64_t m = 0;
    for (int p : tpos) m |= uint64_t(1) << p;
    return m;
  }
};

// Is matrix-index bit `b` (of a dim x dim row-major matrix) a control?
inline bool is_control_bit(const std::vector<cd>& m, int dim, int b) {
  const cd one(1.0, 0.0), zero(0.0, 0.0);
  for (int r = 0; r < dim; ++r)
    for (int c = 0; c < dim; ++c) {
      if (((r >> b) & 1) && ((c >> b) & 1)) continue;
      if (m[(size_t)r * dim + c] != (r == c ? one : zero)) return false;
    }
  return true;
}

inline std::vector<cd> reduce_on_bit(const std::vector<cd>& m, int dim, int b) {
  int nd = dim / 2;
  std::vector<cd> out((size_t)nd * nd);
  auto widen = [b](int x) { return ((x >> b) << (b + 1)) | (1 << b) | (x & ((1 << b) - 1)); };
  for (int r = 0; r < nd; ++r)
    for (int c = 0; c < nd; ++c) out[(size_t)r * nd + c] = m[(size_t)widen(r) * dim + widen(c)];
  return out;
}

// nqubits: qubits of the buffer; targets/controls: Qibo ids (0 = MSB).
inline bool canonicalize(int nqubits, const double* data, bool is_diagonal, int nt, const int* targets, int nc,
                         const int* controls, CanonOp& out, std::string& err) {
  if (nt < 0 || nc < 0 || nt + nc > nqubits) { err = "too many qubits in gate"; return false; }
  uint64_t seen = 0;
  for (int i = 0; i < nt + nc; ++i) {
    int q = i < nt ? targets[i] : controls[i - nt];
    if (q < 0 || q >= nqubits) { err = "qubit index out of range"; return false; }
    if ((seen >> q) & 1) { err = "repeated qubit in gate"; return false; }
    seen |= uint64_t(1) << q;
  }
  out = CanonOp();
  for (int i = 0; i < nc; ++i) out.cpos.push_back(nqubits - 1 - controls[i]);
  std::vector<int> tpos(nt);
  for (int i = 0; i < nt; ++i) tpos[i] = nqubits - 1 - targets[i];
  int dim = 1 << nt;
  const cd* src = reinterpret_cast<const cd*>(data);
  const cd one(1.0, 0.0), zero(0.0, 0.0);

  bool diag = is_diagonal;
  std::vector<cd> m;
  if (is_diagonal) {
    m.assign(src, src + dim);
  } else {
    m.assign(src, src + (size_t)dim * dim);
    diag = true;
    for (int r = 0; r < dim && diag; ++r)
      for (int c = 0; c < dim; ++c)
        if (r != c && m[(size_t)r * dim + c] != zero) { diag = false; break; }
    if (diag) {
      std::vector<cd> d(dim);
      for (int r = 0; r < dim; ++r) d[r] = m[(size_t)r * dim + r];
      m.swap(d);
    }
  }

  if (diag) {
    // control bits of a diagonal: entries with that bit == 0 are all exactly 1
    bool changed = true;
    while (changed && dim > 1) {
      changed = false;
      int k = (int)tpos.size();
      for (int i = 0; i < k; ++i) {
        int b = k - 1 - i;
        bool ctrl = true;
        for (int r = 0; r < dim; ++r)
          if (!((r >> b) & 1) && m[r] != one) { ctrl = false; break; }
        if (!ctrl) continue;
        std::vector<cd> d(dim / 2);
        for (int r = 0; r < dim / 2; ++r) d[r] = m[((r >> b) << (b + 1)) | (1 << b) | (r & ((1 << b) - 1))];
        m.swap(d);
        out.cpos.push_back(tpos[i]);
        tpos.erase(tpos.begin() + i);
        dim /= 2;
        changed = true;
        break;
      }
    }
    std::sort(out.cpos.begin(), out.cpos.end());
    if (dim == 1) {
      if (m[0] == one) { out.kind = CK_NOOP; out.cpos.clear(); return true; }
      out.kind = CK_PHASE; out.data = m; return true;
    }
    bool all_one = true;
    for (auto& v : m) if (v != one) { all_one = false; break; }
    if (all_one) { out.kind = CK_NOOP; out.cpos.clear(); return true; }
    out.kind = CK_DIAG; out.tpos = tpos; out.data = m;
    return true;
  }

  // dense: peel control bits
  bool changed = true;
  while (changed && dim > 2) {
    changed = false;
    int k = (int)tpos.size();
    for (int i = 0; i < k; ++i) {
      int b = k - 1 - i;
      if (!is_control_bit(m, dim, b)) continue;
      m = reduce_on_bit(m, dim, b);
      out.cpos.push_back(tpos[i]);
      tpos.erase(tpos.begin() + i);
      dim /= 2;
      changed = true;
      break;
    }
  }
  std::sort(out.cpos.begin(), out.cpos.end());
  if (dim == 4) {
    static const double sw[16] = {1, 0, 0, 0, 0, 0, 1, 0, 0, 1, 0, 0, 0, 0, 0, 1};
    bool is_swap = true;
    for (int i = 0; i < 16; ++i)
      if (m[i] != cd(sw[i], 0.0)) { is_swap = false; break; }
    if (is_swap) { out.kind = CK_SWAP; out.tpos = tpos; return true; }
  }
  out.kind = CK_DENSE; out.tpos = tpos; out.data = m;
  return true;
}

}  // namespace qb
