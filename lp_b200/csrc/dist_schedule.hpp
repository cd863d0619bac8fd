// dist_schedule.hpp -- host/device-agnostic pieces of the distributed (panel-broadcast) Cholesky:
// the tile list of the block columns one rank owns, and the per-rank order of operations.
// Kept free of CUDA so tests/cpp/test_dist_schedule.cpp can check both on the CPU for any world size.
#pragma once
#include <cmath>

#if defined(__CUDACC__)
#define LPB_SCHED_HD __host__ __device__
#else
#define LPB_SCHED_HD
#endif

namespace lpb {

// Tile list of the trailing update when the block columns are dealt round-robin to `G` ranks: this rank
// owns the trailing block columns tj = f + l G (l = 0, 1, ...), column l has n - tj tiles (ti = tj .. n-1),
// columns are walked one after the other.   before(l) = tiles in columns 0 .. l-1.
struct OwnedCols {
  int G, f, n;  // modulus, first owned trailing column, trailing tile rows
  LPB_SCHED_HD int before(int l) const { return l * (n - f) - G * (l * (l - 1) / 2); }
  LPB_SCHED_HD int count() const {
    if (f >= n) return 0;
    const int nown = (n - f + G - 1) / G;
    return before(nown);
  }
  LPB_SCHED_HD void decode(int t, int* ti, int* tj) const {
    const double a = 0.5 * G, b = (n - f) + 0.5 * G;
    const double disc = b * b - 4.0 * a * t;
    int l = static_cast<int>((b - sqrt(disc > 0.0 ? disc : 0.0)) / (2.0 * a));
    if (l < 0) l = 0;
    while (l > 0 && before(l) > t) --l;
    while (before(l + 1) <= t) ++l;
    *tj = f + l * G;
    *ti = *tj + (t - before(l));
  }
  // first owned trailing column for a trailing matrix that starts at global block column tile0
  static LPB_SCHED_HD OwnedCols make(int G, int rank, int tile0, int ntr) {
    OwnedCols o;
    o.G = G;
    o.f = ((rank - tile0) % G + G) % G;
    o.n = ntr;
    return o;
  }
};

// Order of operations of rank `me` of `G` for a matrix of T block columns (block column k belongs to
// rank k mod G).  Ops supplies, each returning an lpb status (0 = ok):
//   factor_panel(k)            potf2 + TRSM of block column k, pack for the broadcast     (owner only)
//   broadcast(k, owner)        collective: panel k (and inv(L_kk)) from its owner to everyone
//   store_panel(k)             write the received panel into the local copy of M          (non-owners)
//   update_column(p, col)      C[:, col] -= P_p P_p[col]^T, rows >= col                   (look-ahead)
//   update_owned(p, tile0)     the same for every OWNED block column >= tile0
// The rank that owns panel k+1 brings that column up to date right after panel k arrives and defers the rest
// of its share of update k until its own panel k+1 is on the wire: its factor chain overlaps the other
// ranks' updates.  Invariant: when factor_panel(k) runs, every tile of column k has seen panels 0 .. k-1.
template <class Ops>
int potrf_dist_schedule(int T, int G, int me, Ops& ops) {
  int pending = -1;  // panel whose update of the owned columns >= pending + 2 this rank still owes
  for (int k = 0; k < T; ++k) {
    const int owner = k % G;
    int rc;
    if (owner == me && (rc = ops.factor_panel(k)) != 0) return rc;
    if ((rc = ops.broadcast(k, owner)) != 0) return rc;
    if (owner != me && (rc = ops.store_panel(k)) != 0) return rc;
    if (pending >= 0) {
      if ((rc = ops.update_owned(pending, pending + 2)) != 0) return rc;
      pending = -1;
    }
    if (k + 1 < T) {
      if ((k + 1) % G == me) {
        if ((rc = ops.update_column(k, k + 1)) != 0) return rc;
        pending = k;
      } else if ((rc = ops.update_owned(k, k + 1)) != 0) {
        return rc;
      }
    }
  }
  return 0;
}

// ------------------------------------------------------------------------------------------------ schedule v2
// The same factorisation with a shorter serial chain per panel.  What sits between "panel k finished on its
// owner" and "panel k+1 finished on ITS owner" in the schedule above is
//     pack -> broadcast (whole panel) -> unpack -> column update -> potf2 -> TRSM ,
// here it is
//     broadcast of ONE 128 x 128 block -> diagonal-tile update -> potf2 -> TRSM :
//  * the panel is produced straight into a packed send buffer (potf2 / TRSM write M and the buffer), and the
//    updates read it from there (TMA map over the buffer), so pack and unpack leave the chain (non-owners copy
//    the panel into their M after the updates that read it);
//  * the panel travels as TWO broadcasts: first the rows of block k+1 -- all the next owner needs to bring its
//    DIAGONAL tile up to date and start potf2(k+1) -- then the rest.  potf2(k+1) runs on a side stream while the
//    rest of the panel arrives and the column below the diagonal tile is updated on the main stream.
// Buffers: panel p lives in slot p & 1.  Every tile still sees the same updates in the same order as in the
// single-GPU loop, so the factor stays bit-identical.
// Ops (all return an lpb status, 0 = ok; `side` = 1 selects the side stream):
//   potf2(k, side)  trsm(k)  bcast_small(k, owner)  bcast_large(k, owner)
//   record(ev, side)  wait(ev, side)            ev = kEvSmall + (k & 1)  or  kEvPotf2 + (k & 1)
//   update_diag(p, col, side)  update_col_below(p, col)  update_owned(p, tile0)  unpack(k)
enum { kEvSmall = 0, kEvPotf2 = 2, kNumDistEvents = 4 };

template <class Ops>
int potrf_dist_schedule2(int T, int G, int me, Ops& ops) {
#define LPB_S2(call)            \
  do {                          \
    const int rc__ = (call);    \
    if (rc__ != 0) return rc__; \
  } while (0)
  int pending = -1;  // panel whose update of the owned columns >= pending + 2 (and whose unpack) this rank still owes
  for (int k = 0; k < T; ++k) {
    const int owner = k % G;
    if (owner == me) {
      if (k == 0)
        LPB_S2(ops.potf2(0, 0));
      else
        LPB_S2(ops.wait(kEvPotf2 + (k & 1), 0));  // potf2(k) ran on the side stream during step k - 1
      LPB_S2(ops.trsm(k));
    }
    if (k + 1 < T) {
      LPB_S2(ops.bcast_small(k, owner));
      LPB_S2(ops.record(kEvSmall + (k & 1), 0));
    }
    LPB_S2(ops.bcast_large(k, owner));
    if (pending >= 0) {
      LPB_S2(ops.update_owned(pending, pending + 2));
      if (pending % G != me) LPB_S2(ops.unpack(pending));
      pending = -1;
    }
    if (k + 1 < T && (k + 1) % G == me) {
      LPB_S2(ops.wait(kEvSmall + (k & 1), 1));
      LPB_S2(ops.update_diag(k, k + 1, 1));
      LPB_S2(ops.potf2(k + 1, 1));
      LPB_S2(ops.record(kEvPotf2 + ((k + 1) & 1), 1));
      LPB_S2(ops.update_col_below(k, k + 1));
      pending = k;
    } else {
      if (k + 1 < T) LPB_S2(ops.update_owned(k, k + 1));
      if (owner != me) LPB_S2(ops.unpack(k));
    }
  }
  if (pending >= 0 && pending % G != me) LPB_S2(ops.unpack(pending));  // cannot happen (pending < T - 1), kept for safety
#undef LPB_S2
  return 0;
}

}  // namespace lpb
