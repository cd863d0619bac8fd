// CPU check of lp_b200/csrc/dist_schedule.hpp for any world size: (1) OwnedCols::decode enumerates exactly
// the tiles of the owned block columns, each once; (2) potrf_dist_schedule, simulated on all ranks at once,
// applies every panel to every tile exactly once, in order, before the tile's own column is factored, and
// never reads a panel a rank has not received.
#include <cstdio>
#include <cstdlib>
#include <set>
#include <utility>
#include <vector>

#include "../../lp_b200/csrc/dist_schedule.hpp"

using lpb::OwnedCols;

static int fails = 0;
#define CHECK(cond, ...)                 \
  do {                                   \
    if (!(cond)) {                       \
      if (fails < 20) {                  \
        std::printf("FAIL %s: ", #cond); \
        std::printf(__VA_ARGS__);        \
        std::printf("\n");               \
      }                                  \
      ++fails;                           \
    }                                    \
  } while (0)

static void test_decode() {
  for (int G : {1, 2, 3, 4, 5, 8, 16})
    for (int ntr = 0; ntr <= 70; ++ntr)
      for (int tile0 = 0; tile0 < 2 * G + 1; ++tile0)
        for (int rank = 0; rank < G; ++rank) {
          const OwnedCols o = OwnedCols::make(G, rank, tile0, ntr);
          std::set<std::pair<int, int>> want, got;
          for (int tj = 0; tj < ntr; ++tj)
            if ((tile0 + tj) % G == rank)
              for (int ti = tj; ti < ntr; ++ti) want.insert({ti, tj});
          CHECK(o.count() == (int)want.size(), "count G=%d ntr=%d tile0=%d rank=%d: %d vs %zu", G, ntr, tile0, rank,
                o.count(), want.size());
          for (int t = 0; t < o.count(); ++t) {
            int ti, tj;
            o.decode(t, &ti, &tj);
            CHECK(got.insert({ti, tj}).second, "duplicate tile (%d,%d) G=%d ntr=%d", ti, tj, G, ntr);
          }
          CHECK(got == want, "tile set G=%d ntr=%d tile0=%d rank=%d", G, ntr, tile0, rank);
        }
  // the sizes the benchmarks run: C5 on 8 GPUs (256 block columns)
  const OwnedCols big = OwnedCols::make(8, 3, 1, 255);
  long long sum = 0;
  for (int t = 0; t < big.count(); ++t) {
    int ti, tj;
    big.decode(t, &ti, &tj);
    CHECK(tj <= ti && ti < 255 && (1 + tj) % 8 == 3, "big decode t=%d -> (%d,%d)", t, ti, tj);
    sum += ti - tj;
  }
  CHECK(sum > 0, "big decode sum");
}

// ---- schedule simulation
enum OpKind { FACTOR, BCAST, STORE, UPD_COL, UPD_OWNED };
struct Op {
  OpKind kind;
  int a, b;
};
struct Recorder {
  std::vector<Op> ops;
  int factor_panel(int k) { ops.push_back({FACTOR, k, 0}); return 0; }
  int broadcast(int k, int owner) { ops.push_back({BCAST, k, owner}); return 0; }
  int store_panel(int k) { ops.push_back({STORE, k, 0}); return 0; }
  int update_column(int p, int col) { ops.push_back({UPD_COL, p, col}); return 0; }
  int update_owned(int p, int tile0) { ops.push_back({UPD_OWNED, p, tile0}); return 0; }
};

static void test_schedule(int T, int G) {
  std::vector<Recorder> rec(G);
  for (int r = 0; r < G; ++r) lpb::potrf_dist_schedule(T, G, r, rec[r]);
  // per rank: ver[i][j] = panels applied to tile (i, j) of the local copy; have[k] = panel k present locally
  std::vector<std::vector<std::vector<int>>> ver(G, std::vector<std::vector<int>>(T, std::vector<int>(T, 0)));
  std::vector<std::vector<char>> have(G, std::vector<char>(T, 0)), packed(G, std::vector<char>(T, 0));
  std::vector<size_t> pc(G, 0);
  auto update = [&](int r, int p, int col) {
    CHECK(have[r][p], "rank %d updates with panel %d it does not hold (T=%d G=%d)", r, p, T, G);
    for (int i = col; i < T; ++i) {
      CHECK(ver[r][i][col] == p, "rank %d tile (%d,%d): panel %d applied out of order (seen %d)", r, i, col, p,
            ver[r][i][col]);
      ver[r][i][col]++;
    }
  };
  for (;;) {
    bool progress = false, all_done = true;
    for (int r = 0; r < G; ++r) {
      while (pc[r] < rec[r].ops.size() && rec[r].ops[pc[r]].kind != BCAST) {
        const Op op = rec[r].ops[pc[r]++];
        progress = true;
        if (op.kind == FACTOR) {
          CHECK(op.a % G == r, "rank %d factors panel %d", r, op.a);
          for (int i = op.a; i < T; ++i)
            CHECK(ver[r][i][op.a] == op.a, "panel %d factored with tile (%d,%d) at %d updates (T=%d G=%d)", op.a, i,
                  op.a, ver[r][i][op.a], T, G);
          have[r][op.a] = 1;
          packed[r][op.a] = 1;
        } else if (op.kind == STORE) {
          CHECK(packed[r][op.a], "rank %d stores panel %d before receiving it", r, op.a);
          have[r][op.a] = 1;
        } else if (op.kind == UPD_COL) {
          CHECK(op.b % G == r, "look-ahead column %d on rank %d", op.b, r);
          update(r, op.a, op.b);
        } else {  // UPD_OWNED
          for (int j = op.b; j < T; ++j)
            if (j % G == r) update(r, op.a, j);
        }
      }
      if (pc[r] < rec[r].ops.size()) all_done = false;
    }
    if (all_done) break;
    // every unfinished rank sits at a broadcast: it completes when all of them are at the SAME one
    int k = -1;
    bool same = true;
    for (int r = 0; r < G; ++r) {
      if (pc[r] >= rec[r].ops.size()) { same = false; break; }
      const Op& op = rec[r].ops[pc[r]];
      if (k < 0) k = op.a;
      if (op.a != k) same = false;
    }
    CHECK(same, "ranks disagree on the next broadcast (T=%d G=%d)", T, G);
    if (!same) return;
    const int owner = rec[0].ops[pc[0]].b;
    CHECK(packed[owner][k], "panel %d broadcast before its owner %d factored it", k, owner);
    for (int r = 0; r < G; ++r) {
      packed[r][k] = 1;
      pc[r]++;
    }
    progress = true;
    if (!progress) { CHECK(false, "deadlock T=%d G=%d", T, G); return; }
  }
  for (int r = 0; r < G; ++r)
    for (int k = 0; k < T; ++k) {
      CHECK(have[r][k], "rank %d never got panel %d (T=%d G=%d)", r, k, T, G);
      if (k % G == r)
        for (int i = k; i < T; ++i) CHECK(ver[r][i][k] == k, "owned tile (%d,%d) saw %d panels", i, k, ver[r][i][k]);
    }
}

int main() {
  test_decode();
  for (int G : {2, 3, 4, 8})
    for (int T : {1, 2, 3, 4, 7, 8, 9, 16, 17, 33}) test_schedule(T, G);
  if (fails) {
    std::printf("%d check(s) failed\n", fails);
    return 1;
  }
  std::printf("dist_schedule ok\n");
  return 0;
}
