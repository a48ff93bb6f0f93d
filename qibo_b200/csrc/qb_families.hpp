// qibo_b200: gate families whose matrix the library evaluates from angles (qb_program_set_params), so that a variational
// loop sends angles instead of matrices.  Formulas as backends/npmatrices.py (RX :79, RY :84, RZ :89, U1, CU1 :230,
// CRX / CRY / CRZ), in double.  Host code, shared with tests/emul.
#pragma once
#include <cmath>
#include <vector>

#include "../../include/qibo_b200.h"
#include "qb_canon.hpp"

namespace qb {

inline bool family_matrix(int family, const double* th, int ntargets, bool is_diagonal, std::vector<double>& out) {
  const double c = cos(0.5 * th[0]), s = sin(0.5 * th[0]);
  auto put = [&](std::vector<cd> m) {
    out.resize(2 * m.size());
    for (size_t i = 0; i < m.size(); ++i) {
      out[2 * i] = m[i].real();
      out[2 * i + 1] = m[i].imag();
    }
  };
  const cd I(0.0, 1.0), one(1.0, 0.0), zero(0.0, 0.0);
  const cd e1 = cd(cos(th[0]), sin(th[0]));      // exp(i theta)
  const cd eh = cd(c, s), ehc = cd(c, -s);        // exp(+-i theta / 2)
  std::vector<cd> blk;
  switch (family) {
    case QB_GATE_RX: blk = {cd(c, 0), -I * s, -I * s, cd(c, 0)}; break;
    case QB_GATE_RY: blk = {cd(c, 0), cd(-s, 0), cd(s, 0), cd(c, 0)}; break;
    case QB_GATE_RZ: blk = {ehc, zero, zero, eh}; break;
    case QB_GATE_U1: blk = {one, zero, zero, e1}; break;
    case QB_GATE_CRX: blk = {cd(c, 0), -I * s, -I * s, cd(c, 0)}; break;
    case QB_GATE_CRY: blk = {cd(c, 0), cd(-s, 0), cd(s, 0), cd(c, 0)}; break;
    case QB_GATE_CRZ: blk = {ehc, zero, zero, eh}; break;
    case QB_GATE_CU1: blk = {one, zero, zero, e1}; break;
    default: return false;
  }
  const bool controlled = family >= QB_GATE_CRX;
  if (ntargets != (controlled ? 2 : 1)) return false;
  if (is_diagonal) {
    if (blk[1] != zero || blk[2] != zero) return false;
    if (controlled) put({one, one, blk[0], blk[3]});
    else put({blk[0], blk[3]});
  } else if (controlled) {
    put({one, zero, zero, zero, zero, one, zero, zero, zero, zero, blk[0], blk[1], zero, zero, blk[2], blk[3]});
  } else {
    put(blk);
  }
  return true;
}


}  // namespace qb
