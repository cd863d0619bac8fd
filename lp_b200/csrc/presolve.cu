// presolve.cu -- opt-in presolve of a slack-form LP (host only, O(mn), runs once per problem).
//
// The reference has none: it is the crate's own TODO (/root/reference/CONTRIBUTING.md:8, "presolving: removing
// redundant constraints, scaling").  SURVEY.md 8(f)4.  Nothing here is on the solve path unless the caller asks for it
// (lp_b200.presolve / lpb_presolve_*), so parity with the reference is untouched.  Steps, in this order:
//   1. empty rows: a_i = 0 with b_i = 0 is dropped; with b_i != 0 the LP is infeasible (reported, nothing solved);
//   2. duplicate rows: rows that are positive or negative multiples of an earlier row (hashed on their normalised
//      pattern, then compared entry by entry) are dropped when their right-hand sides agree, and make the LP
//      infeasible when they do not -- exactly the "linearly dependent constraints" the reference's NumericalProblem
//      message warns about (error.rs:12-14);
//   3. scaling: `passes` rounds of geometric row / column equilibration (each row, then each column, divided by
//      sqrt(min |a| * max |a|) over its non-zeros, rounded to a power of two so the scaling itself is exact).
// The presolved problem is  min (C c)' y  st  (R A C) y = R b,  y >= 0  with x = C y;  restore_x undoes the scaling.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <new>
#include <unordered_map>
#include <vector>

#include "common.cuh"

struct lpb_presolve {
  int64_t m = 0, n = 0, n_slack = 0;  // of the presolved problem (columns are never removed)
  std::vector<double> A, b, c;        // row-major m x n
  std::vector<double> col_scale;      // x = col_scale * y
  std::vector<int64_t> kept_rows;     // original index of every kept row
  int64_t dropped_empty = 0, dropped_duplicate = 0;
  int status = LPB_OK;                // LPB_OK or LPB_ERR_INFEASIBLE
};

namespace {

double pow2_round(double v) {  // nearest power of two: scaling by it is exact in binary floating point
  if (!(v > 0.0) || !std::isfinite(v)) return 1.0;
  return std::exp2(std::round(std::log2(v)));
}

}  // namespace

extern "C" {

int lpb_presolve_create(lpb_presolve** out, int64_t m, int64_t n, const double* A, int64_t lda, const double* b,
                        const double* c, int64_t n_slack, int scale_passes) {
  if (!out || m <= 0 || n <= 0 || !A || !b || !c || lda < n || n_slack < 0 || n_slack > n || scale_passes < 0) {
    lpb::set_last_error("presolve: bad argument");
    return LPB_ERR_BAD_ARGUMENT;
  }
  lpb_presolve* p = new (std::nothrow) lpb_presolve();
  if (!p) return LPB_ERR_BAD_ARGUMENT;
  p->n = n;
  p->n_slack = n_slack;
  // ---- 1 + 2: which rows stay
  std::unordered_map<uint64_t, std::vector<int64_t>> buckets;  // hash of the normalised row -> kept rows with it
  std::vector<double> norm((size_t)n);
  for (int64_t i = 0; i < m && p->status == LPB_OK; ++i) {
    const double* row = A + i * lda;
    int64_t first = -1;
    double big = 0.0;
    for (int64_t j = 0; j < n; ++j) {
      if (row[j] != 0.0 && first < 0) first = j;
      big = std::max(big, std::fabs(row[j]));
    }
    if (first < 0) {  // empty row
      if (b[i] != 0.0) p->status = LPB_ERR_INFEASIBLE;
      ++p->dropped_empty;
      continue;
    }
    // normalise: leading non-zero = +1, so multiples (of either sign) of one row share a key
    const double s = 1.0 / row[first];
    uint64_t h = 1469598103934665603ull;
    for (int64_t j = 0; j < n; ++j) {
      norm[(size_t)j] = row[j] * s;
      if (row[j] != 0.0) {
        // hash the pattern and ~40 leading bits of the normalised value (equality is checked exactly below)
        uint64_t bits;
        const double q = norm[(size_t)j];
        std::memcpy(&bits, &q, sizeof(bits));
        h = (h ^ (uint64_t)j) * 1099511628211ull;
        h = (h ^ (bits >> 24)) * 1099511628211ull;
      }
    }
    bool duplicate = false;
    auto it = buckets.find(h);
    if (it != buckets.end()) {
      for (int64_t k : it->second) {
        const double* other = A + k * lda;
        int64_t fk = 0;
        while (other[fk] == 0.0) ++fk;
        if (fk != first) continue;
        const double sk = 1.0 / other[fk];
        bool same = true;
        for (int64_t j = 0; j < n && same; ++j) {
          const double u = norm[(size_t)j], v = other[j] * sk;
          same = std::fabs(u - v) <= 1e-12 * std::max(1.0, std::max(std::fabs(u), std::fabs(v)));
        }
        if (!same) continue;
        duplicate = true;
        const double bi = b[i] * s, bk = b[k] * sk;
        if (std::fabs(bi - bk) > 1e-9 * std::max(1.0, std::max(std::fabs(bi), std::fabs(bk)))) p->status = LPB_ERR_INFEASIBLE;
        break;
      }
    }
    if (duplicate) {
      ++p->dropped_duplicate;
      continue;
    }
    buckets[h].push_back(i);
    p->kept_rows.push_back(i);
  }
  if (p->status == LPB_OK && p->kept_rows.empty()) p->status = LPB_ERR_UNCONSTRAINED;
  p->m = (int64_t)p->kept_rows.size();
  p->A.assign((size_t)(p->m * n), 0.0);
  p->b.assign((size_t)p->m, 0.0);
  p->c.assign(c, c + n);
  p->col_scale.assign((size_t)n, 1.0);
  for (int64_t r = 0; r < p->m; ++r) {
    std::memcpy(&p->A[(size_t)(r * n)], A + p->kept_rows[(size_t)r] * lda, sizeof(double) * (size_t)n);
    p->b[(size_t)r] = b[p->kept_rows[(size_t)r]];
  }
  // ---- 3: geometric equilibration, powers of two
  for (int pass = 0; pass < scale_passes && p->status == LPB_OK; ++pass) {
    for (int64_t r = 0; r < p->m; ++r) {
      double lo = INFINITY, hi = 0.0;
      double* row = &p->A[(size_t)(r * n)];
      for (int64_t j = 0; j < n; ++j)
        if (row[j] != 0.0) {
          lo = std::min(lo, std::fabs(row[j]));
          hi = std::max(hi, std::fabs(row[j]));
        }
      const double f = 1.0 / pow2_round(std::sqrt(lo * hi));
      for (int64_t j = 0; j < n; ++j) row[j] *= f;
      p->b[(size_t)r] *= f;
    }
    std::vector<double> lo((size_t)n, INFINITY), hi((size_t)n, 0.0);
    for (int64_t r = 0; r < p->m; ++r) {
      const double* row = &p->A[(size_t)(r * n)];
      for (int64_t j = 0; j < n; ++j)
        if (row[j] != 0.0) {
          lo[(size_t)j] = std::min(lo[(size_t)j], std::fabs(row[j]));
          hi[(size_t)j] = std::max(hi[(size_t)j], std::fabs(row[j]));
        }
    }
    for (int64_t j = 0; j < n; ++j) {
      if (hi[(size_t)j] == 0.0) continue;  // an all-zero column keeps its scale
      const double f = 1.0 / pow2_round(std::sqrt(lo[(size_t)j] * hi[(size_t)j]));
      for (int64_t r = 0; r < p->m; ++r) p->A[(size_t)(r * n + j)] *= f;
      p->c[(size_t)j] *= f;
      p->col_scale[(size_t)j] *= f;
    }
  }
  *out = p;
  return LPB_OK;
}

int lpb_presolve_info(const lpb_presolve* p, int64_t* m_out, int64_t* n_out, int64_t* n_slack_out,
                      int64_t* dropped_empty, int64_t* dropped_duplicate, int* status) {
  if (!p) return LPB_ERR_BAD_ARGUMENT;
  if (m_out) *m_out = p->m;
  if (n_out) *n_out = p->n;
  if (n_slack_out) *n_slack_out = p->n_slack;
  if (dropped_empty) *dropped_empty = p->dropped_empty;
  if (dropped_duplicate) *dropped_duplicate = p->dropped_duplicate;
  if (status) *status = p->status;
  return LPB_OK;
}

int lpb_presolve_get(const lpb_presolve* p, double* A_out, int64_t lda_out, double* b_out, double* c_out) {
  if (!p || !A_out || !b_out || !c_out || lda_out < p->n) return LPB_ERR_BAD_ARGUMENT;
  for (int64_t r = 0; r < p->m; ++r)
    std::memcpy(A_out + r * lda_out, &p->A[(size_t)(r * p->n)], sizeof(double) * (size_t)p->n);
  std::memcpy(b_out, p->b.data(), sizeof(double) * (size_t)p->m);
  std::memcpy(c_out, p->c.data(), sizeof(double) * (size_t)p->n);
  return LPB_OK;
}

int lpb_presolve_restore_x(const lpb_presolve* p, const double* y, double* x) {
  if (!p || !y || !x) return LPB_ERR_BAD_ARGUMENT;
  for (int64_t j = 0; j < p->n; ++j) x[j] = p->col_scale[(size_t)j] * y[j];
  return LPB_OK;
}

int lpb_presolve_destroy(lpb_presolve* p) {
  delete p;
  return LPB_OK;
}

}  // extern "C"
