// dmma_gemm.cu -- K1: lower(M) = A diag(d) A^T, and the Cholesky trailing update C -= P P^T,
// on the Blackwell FP64 tensor path (DMMA) fed by TMA.
//
// Replaces /root/reference/src/solvers/interior_point/newton_equations.rs:54-57
// (`A.dot(&(Dinv[:,None] * A.t()))`: a full 2m^2n GEMM plus an n x m temporary) with a
// triangle-only m(m+1)n SYRK whose D-scaling is fused into the operand fragments.
//
// Design (sm_100a):
//  * tcgen05.mma has no f64 kind; FP64 tensor math on Blackwell is the warp-level
//    mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4).  Both operands of M[i,j] = sum_k A[i,k] d[k] A[j,k]
//    are K-contiguous rows of row-major A, which is exactly the `.row.col` form.
//  * One persistent CTA per SM walks the lower-triangular 128x128 tile list.  Lane 0 of warp 0
//    streams 128x16 operand boxes (128-byte rows, SWIZZLE_128B) plus the 16 d[k] values of each
//    K-block with cp.async.bulk.tensor into a 5-stage mbarrier ring, 3 K-blocks ahead of the math
//    (a ninth, dedicated producer warp does not fit: the register file is 4 x 16K per sub-partition
//    and each DMMA warp needs ~200 registers); 8 warps hold a 64x32 accumulator slab each
//    (64 doubles / thread) and issue 64 DMMAs per K-block.
//  * Fragments are read with conflict-free ld.shared.v2.f64: thread (g,t) of a warp fetches the
//    16-byte chunk (2t+P)^g of row g, so a quarter-warp covers all 8 chunks of the 128-byte
//    swizzle atom; the .x halves feed one DMMA k-group {4t+2P}, the .y halves the next {4t+2P+1}
//    (any k permutation is legal as long as A- and B-fragments agree).
//  * d[k] is applied to the B-fragment in registers (8 DMUL per 64 DMMA).
#include <cuda.h>

#include "dist_schedule.hpp"
#include "kernels.hpp"

namespace lpb {

namespace {

constexpr int BM = 128, BN = 128, BK = 16;
constexpr int kStages = 5;
constexpr int kLead = 3;  // TMA runs this many K-blocks ahead of the DMMA warps
// K1 only: the accumulators are folded into C every `fblocks` K-blocks (LaunchCtx::syrk_flush_blocks, default 32 =
// 512 columns) instead of summing the whole K extent in one register chain.  A chain of n/4 DMMA steps rounds relative
// to the growing partial sum: 33 ulp rms (117 max) on the same-sign diagonal sums of M at n = 24576 -- a single cuBLAS
// DGEMM call measures the same -- against 3.1 ulp blocked (cuBLAS called per 1024 columns: 3.6; tools/time_syrk.py).
// That, not the factorisation or the solves, is what made the unrefined Newton solve lose d_tau's denominator late in
// the iteration (tools/diag_bisect_host.py: M formed on the host -> every device factor / solve is fine; M formed on
// the device in one chain, by this kernel or by cuBLAS -> none is).  The flush is staggered over the row groups of a
// slab and over the two warps of a sub-partition and its reads are asynchronous (see the main loop): +1.0 % at C3.
// kFlushBlocks is only the fallback period of callers that pass none.
constexpr int kFlushBlocks = 128;
constexpr int kConsumerWarps = 8;
constexpr int kThreads = kConsumerWarps * 32;
constexpr uint32_t kTileBytes = BM * BK * 8;  // 16 KB
constexpr uint32_t kDBytes = BK * 8;          // 128 B
constexpr uint32_t kSmemA = 0;
constexpr uint32_t kSmemB = kSmemA + kStages * kTileBytes;
constexpr uint32_t kSmemD = kSmemB + kStages * kTileBytes;
constexpr uint32_t kSmemBar = kSmemD + kStages * kDBytes;
constexpr uint32_t kSmemTotal = kSmemBar + 2 * kStages * 8;
constexpr uint32_t kSmemAlloc = kSmemTotal + 1024;  // slack for 1024-byte alignment

// Shared-memory plan per kernel variant.  VAR bit 3 (MODE_UPDATE only, "C prefetch"): the 128 x 128 tile of C that
// seeds the accumulators is fetched by TMA into shared memory (8 boxes of 128 x 16, same swizzled layout as an
// operand stage) while the PREVIOUS tile is still in its main loop, so its HBM latency no longer sits at the head
// of every tile (K is only 128 there: 25 % of a tile's time was the exposed load + store of C).  The ring shrinks to
// 3 stages / a lead of 2 K-blocks to make room: 96 KB ring + 128 KB C tile.
template <int MODE, int VAR>
struct Plan {
  static constexpr bool kCpf = MODE == 1 && (VAR & 8) != 0;
  static constexpr int kNS = kCpf ? 3 : kStages;
  static constexpr int kLd = kCpf ? 2 : kLead;
  static constexpr uint32_t kA = 0;
  static constexpr uint32_t kB = kA + kNS * kTileBytes;
  static constexpr uint32_t kC = kB + kNS * kTileBytes;                  // 1024-byte aligned (multiples of 16 KB)
  static constexpr uint32_t kD = kC + (kCpf ? 8 * kTileBytes : 0);
  static constexpr uint32_t kBar = kD + kNS * kDBytes;
  static constexpr uint32_t kF = (kBar + (2 * kNS + 2) * 8 + 15u) & ~15u;  // MODE_SYRK: staging of old C for the flushes
  static constexpr uint32_t kTotal = kF + (MODE == 0 && (VAR & 16) == 0 ? kThreads * 64 : 0);
  static constexpr uint32_t kAlloc = kTotal + 1024;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// Release of a ring stage by a consumer warp.  The stage is overwritten by TMA as soon as the eighth
// arrive lands, so every ld.shared of the iteration must have COMPLETED, not merely been issued,
// before it: ptxas places SYNCS.ARRIVE directly behind the last LDS without waiting on its scoreboard,
// and under store back-pressure (the epilogues of the other warps at a tile boundary) an issued LDS
// was observed to read the stage after its refill (bit-wrong rows in ~1e-6 of the tiles, found by
// factoring the same matrix twice).  `dep` is a word of the last fragment loaded in the iteration
// (shared loads of a warp complete in order) and `zero` a run-time zero the compiler cannot fold: the barrier address now depends on
// the loaded data, which forces the scoreboard wait.  The fragment loads are ld.volatile so that the last
// one in program order is also the last one issued.
__device__ __forceinline__ void mbar_arrive_after_loads(uint32_t bar, uint32_t dep, uint32_t zero) {
  asm volatile(
      "{\n\t.reg .b32 t;\n\t"
      "mad.lo.u32 t, %1, %2, %0;\n\t"
      "mbarrier.arrive.shared::cta.b64 _, [t];\n\t}" ::"r"(bar),
      "r"(dep), "r"(zero)
      : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok;
}
// Bounded wait: a lost TMA / barrier bug traps after ~4 s instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 8000000000ll) __trap();
  }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
          dst),
      "l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1), "r"(bar)
      : "memory");
}
__device__ __forceinline__ double2 lds_v2(uint32_t addr) {
  double2 v;
  // volatile: ptxas keeps these loads in program order among themselves (see mbar_arrive_after_loads)
  asm volatile("ld.volatile.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
  return v;
}
__device__ __forceinline__ void cp_async_16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
      : "+d"(c0), "+d"(c1)
      : "d"(a), "d"(b));
}

// linear lower-triangle index -> (ti, tj), tj <= ti
__device__ __forceinline__ void tri_decode(int L, int* ti, int* tj) {
  int i = static_cast<int>((sqrt(8.0 * static_cast<double>(L) + 1.0) - 1.0) * 0.5);
  while (i * (i + 1) / 2 > L) --i;
  while ((i + 1) * (i + 2) / 2 <= L) ++i;
  *ti = i;
  *tj = L - i * (i + 1) / 2;
}

// The same lower-triangular tile list walked in BANDS of kBand tile rows, column by column inside a band.
// The ~148 tiles in flight then cover ~12 row blocks x ~12 column blocks of A instead of 1 x 148, so a wave
// streams (12 + 12) operand blocks from HBM rather than (1 + 148): DRAM traffic of the SYRK drops ~4x
// (profiles/ncu_syrk_C3_r01.txt: 137 GB per launch for 3.2 GB of A with the row-major order).  Band
// boundaries coincide with row boundaries of the row-major numbering, so band b = (row of L) / kBand.
constexpr int kBand = 12;  // ~ sqrt(148)
__device__ __forceinline__ void band_decode(int L, int ntr, int* ti, int* tj) {
  int row, col;
  tri_decode(L, &row, &col);
  const int r0 = (row / kBand) * kBand;
  const int r1 = r0 + kBand < ntr ? r0 + kBand : ntr;
  const int h = r1 - r0;
  int off = L - r0 * (r0 + 1) / 2;
  const int rect = (r0 + 1) * h;  // columns 0 .. r0 hold h tiles each
  if (off < rect) {
    *tj = off / h;
    *ti = r0 + off % h;
    return;
  }
  off -= rect;
  int c = r0 + 1, hh = h - 1;     // columns r0+1 .. r1-1 hold h-1, h-2, ... tiles
  while (off >= hh) {
    off -= hh;
    ++c;
    --hh;
  }
  *tj = c;
  *ti = c + off;
}

// Producer cursor: walks this CTA's (tile, k-block) sequence `kLead` iterations ahead of the
// consumers.  Lives in lane 0 of warp 0 (the register file is split 4 x 16K per SM sub-partition,
// so a ninth warp would not fit next to eight 200-register DMMA warps).
struct LoadCursor {
  int L, kb, row_i, row_j, stage;
  int kb_lo, kb_n;  // K-block range of the work item (a whole tile: 0, nkb)
  uint32_t phase;
};

// MODE_SYRK, tail of the persistent schedule: when the tile count leaves a short last wave (C2: 528 tiles on 148 SMs =
// 3.57 waves, the fourth one 57 % full but a whole tile-time long), the last `ntiles - n_full` tiles are cut into
// `nsplit` K-ranges each.  Every part is an item of its own with a private 128 x 128 slot in `ws`
// (syrk_tail_fixup_kernel adds the parts in order afterwards: deterministic, no atomics), so the last wave's work
// spreads over all SMs.  nsplit = 1: every item is a whole tile.
struct TailSplit {
  int n_full, nsplit;
  double* ws;
};

// MODE_SYRK   : C = A diag(d) A^T (d optional), lower-triangular tile list, K = [k_begin, k_begin + 16 nkb).
// MODE_UPDATE : C -= P P^T (Cholesky trailing update): accumulators start from C, the B fragment is
//               negated, so the epilogue is store-only.
// MODE_TRSM   : C[:, col_origin + 0..127] = A B^T with B = inv(L_kk) (128 x 128, its own tensor map):
//               the panel TRSM X L_kk^T = P expressed as a GEMM; rectangular tile list (ntr x 1),
//               in place (a CTA reads all K-blocks of its own rows before it stores them).
// Tile lists (`shape`): SHAPE_TRI the lower triangle row by row; SHAPE_COL its first block column only
// (ntr x 1: the look-ahead update of the next panel's column); SHAPE_OWNED the block columns this rank
// owns in the distributed factorisation (OwnedCols).
enum { MODE_SYRK = 0, MODE_UPDATE = 1, MODE_TRSM = 2 };
enum { SHAPE_TRI = 0, SHAPE_COL = 1, SHAPE_OWNED = 2, SHAPE_COLB = 3 };  // COLB: column tile0 WITHOUT its diagonal tile

// VAR (experiments on the trailing update): bit 0 = accumulate from zero and read-modify-write C in
// the epilogue instead of initialising the accumulators from C; bit 1 = CTA barrier at every tile end;
// bit 2 = legacy stage release (plain arrive right behind the last LDS: reproduces the race, control only).
template <int MODE, bool SCALE, int VAR = 0>
__global__ void __maxnreg__(255)
syrk_dmma_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 double* __restrict__ C, int64_t ldc, int m_total, int tile0, int ntr, int k_begin, int nkb,
                 int col_origin, int shape, OwnedCols own, int p_row0, TailSplit ts) {
  // p_row0 >= 0 (MODE_UPDATE only): the panel P is read from a PACKED buffer (tmA maps it: 128 columns, row
  // (r - p_row0) of the panel at buffer row (r - p_row0), except that the first two 128-row blocks are swapped:
  // the rows of block p_row0 / 128 + 1 come first, see k_potrf_dist2) instead of from columns k_begin.. of C.
  using P = Plan<MODE, VAR>;
  constexpr bool FLUSH = (VAR & 16) == 0;  // VAR bit 4: one register chain over the whole K extent (the round-1 kernel)
  constexpr int NS = P::kNS, LEAD = P::kLd;
  constexpr bool CPF = P::kCpf;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sA = smem_base + P::kA, sB = smem_base + P::kB, sD = smem_base + P::kD, sC = smem_base + P::kC;
  const uint32_t bar_full = smem_base + P::kBar, bar_empty = bar_full + NS * 8;
  const uint32_t bar_cfull = bar_empty + NS * 8, bar_cempty = bar_cfull + 8;
  const uint32_t sF = smem_base + P::kF + threadIdx.x * 16u;  // this thread's staging slots: sF + ni * (kThreads * 16)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < NS; ++s) {
      mbar_init(bar_full + s * 8, 1);
      mbar_init(bar_empty + s * 8, kConsumerWarps);
    }
    mbar_init(bar_cfull, 1);
    mbar_init(bar_cempty, kConsumerWarps);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  // VAR bit 5: tail tiles may be cut along K (TailSplit).  A separate instantiation: the item / destination indirection
  // costs the plain kernel 0.8 % at C3 (193.7 against 192.3 ms), where no launch ever splits.
  constexpr bool SPLIT = MODE == MODE_SYRK && FLUSH && (VAR & 32) != 0;
  const int ntiles_whole = (MODE == MODE_TRSM || shape == SHAPE_COL) ? ntr
                     : (shape == SHAPE_COLB ? ntr - 1 : (shape == SHAPE_OWNED ? own.count() : ntr * (ntr + 1) / 2));
  // work items: whole tiles, then (SPLIT only) the K-parts of the tail tiles
  const int ntiles = SPLIT ? ts.n_full + (ntiles_whole - ts.n_full) * ts.nsplit : ntiles_whole;
  auto item = [&](int I, int* L, int* part, int* kb_lo, int* kb_n) {
    if (!SPLIT || I < ts.n_full) {
      *L = I;
      *part = -1;
      *kb_lo = 0;
      *kb_n = nkb;
    } else {
      const int q = I - ts.n_full;
      *L = ts.n_full + q / ts.nsplit;
      *part = q % ts.nsplit;
      *kb_lo = (int)((int64_t)nkb * *part / ts.nsplit);
      *kb_n = (int)((int64_t)nkb * (*part + 1) / ts.nsplit) - *kb_lo;
    }
  };
  const bool is_producer = threadIdx.x == 0;
  constexpr uint32_t kBytes = 2 * kTileBytes + (SCALE ? kDBytes : 0);

  auto decode = [&](int L, int* ti, int* tj) {
    if (MODE == MODE_TRSM || shape == SHAPE_COL) {
      *ti = L;
      *tj = 0;
    } else if (shape == SHAPE_COLB) {
      *ti = L + 1;
      *tj = 0;
    } else if (shape == SHAPE_OWNED) {
      own.decode(L, ti, tj);
    } else if (MODE == MODE_SYRK) {
      band_decode(L, ntr, ti, tj);
    } else {
      tri_decode(L, ti, tj);
    }
  };

  LoadCursor pc;
  pc.L = blockIdx.x;
  pc.kb = 0;
  pc.stage = 0;
  pc.phase = 0;
  pc.row_i = pc.row_j = 0;
  pc.kb_lo = 0;
  pc.kb_n = nkb;
  auto cursor_tile = [&]() {
    if (pc.L < ntiles) {
      int ti, tj, Lt, part;
      item(pc.L, &Lt, &part, &pc.kb_lo, &pc.kb_n);
      decode(Lt, &ti, &tj);
      pc.row_i = (tile0 + ti) * BM;
      pc.row_j = MODE == MODE_TRSM ? 0 : (tile0 + tj) * BN;
    }
  };
  // issue the TMA loads of the cursor's iteration (if any is left) and advance it
  auto produce_one = [&]() {
    if (pc.L >= ntiles) return;
    mbar_wait(bar_empty + pc.stage * 8, pc.phase ^ 1u);
    const uint32_t full = bar_full + pc.stage * 8;
    mbar_expect_tx(full, kBytes);
    const int k = k_begin + (pc.kb_lo + pc.kb) * BK;
    int ra = pc.row_i, rb = pc.row_j;
    if (MODE == MODE_UPDATE && p_row0 >= 0) {  // packed panel: block (p_row0 / 128 + 1) sits at buffer row 0
      ra -= p_row0;
      rb -= p_row0;
      if (ra == BM) ra = 0;
      if (rb == BN) rb = 0;
    }
    tma_load_2d(sA + pc.stage * kTileBytes, &tmA, k, ra, full);
    if (MODE == MODE_TRSM)
      tma_load_2d(sB + pc.stage * kTileBytes, &tmB, (pc.kb_lo + pc.kb) * BK, 0, full);
    else
      tma_load_2d(sB + pc.stage * kTileBytes, &tmA, k, rb, full);
    if (SCALE) tma_load_2d(sD + pc.stage * kDBytes, &tmB, k, 0, full);
    if (++pc.stage == NS) {
      pc.stage = 0;
      pc.phase ^= 1u;
    }
    if (++pc.kb == pc.kb_n) {
      pc.kb = 0;
      pc.L += gridDim.x;
      cursor_tile();
    }
  };
  // C prefetch (CPF): the tile of C for this CTA's n-th tile travels through sC; tmB maps the matrix C lives in.
  int c_next = blockIdx.x;  // tile whose C is fetched next (producer only)
  int c_issued = 0;         // number of C tiles requested so far
  auto produce_c = [&]() {
    if (!CPF || c_next >= ntiles) return;
    if (c_issued > 0) mbar_wait(bar_cempty, static_cast<uint32_t>((c_issued - 1) & 1));  // all 8 warps seeded from sC
    int ti, tj;
    decode(c_next, &ti, &tj);
    const int r0 = (tile0 + ti) * BM, c0 = (tile0 + tj) * BN;
    mbar_expect_tx(bar_cfull, 8 * kTileBytes);
#pragma unroll
    for (int b = 0; b < 8; ++b) tma_load_2d(sC + b * kTileBytes, &tmB, c0 + b * BK, r0, bar_cfull);
    c_next += gridDim.x;
    ++c_issued;
  };
  if (is_producer) {
    cursor_tile();
    produce_c();
    for (int i = 0; i < LEAD; ++i) produce_one();
  }

  // -------------------------------------------------------------- DMMA consumers (all 8 warps)
  int stage = 0;
  uint32_t phase = 0;
  const int wm = warp >> 2, wn = warp & 3;  // 2 x 4 warps -> 64 x 32 slab each
  const int g = lane >> 2, t = lane & 3;
  const uint32_t offA = static_cast<uint32_t>(wm * 64 + g) * 128u;
  const uint32_t offB = static_cast<uint32_t>(wn * 32 + g) * 128u;
  const uint32_t sw[2] = {static_cast<uint32_t>(((2 * t + 0) ^ g) << 4), static_cast<uint32_t>(((2 * t + 1) ^ g) << 4)};
  const uint32_t dof[2] = {static_cast<uint32_t>((2 * t + 0) << 4), static_cast<uint32_t>((2 * t + 1) << 4)};

  // MODE_SYRK: `col_origin` carries the flush period in K-blocks (a power of two >= 32, kFlushBlocks by default)
  const int fblocks = MODE == MODE_SYRK && col_origin >= 32 ? col_origin : kFlushBlocks;
  const int fstride = fblocks >> 3;  // the 8 row groups of a warp's slab flush this many K-blocks apart
  const int fshift = 31 - __clz(fstride);
  // warps w and w + 4 share an SM sub-partition: the second four flush half a stride later, so a sub-partition's DMMA
  // pipe always has one warp issuing while the other folds a row group (both at once idled it ~0.6 us per flush)
  const int fwarp = (threadIdx.x >> 7) * (fstride >> 1);
  const uint32_t rt_zero = static_cast<uint32_t>(static_cast<uint64_t>(ldc) >> 62);  // 0 at run time, opaque to the compiler
  int tile_n = 0;  // tiles this CTA has started (parity of the C-prefetch barriers)
  for (int L = blockIdx.x; L < ntiles; L += gridDim.x, ++tile_n) {
    int ti, tj, Lt, part, kb_lo, kbn;
    item(L, &Lt, &part, &kb_lo, &kbn);
    decode(Lt, &ti, &tj);
    // destination of this item: the tile of C, or (a K-part of a tail tile) its private slot in ts.ws, addressed with
    // the same global row / column indices
    double* Cd = C;
    int64_t ldd = ldc;
    if (SPLIT && part >= 0) {
      ldd = BN;
      Cd = ts.ws + (static_cast<int64_t>(Lt - ts.n_full) * ts.nsplit + part) * (BM * BN) -
           (static_cast<int64_t>(tile0 + ti) * BM * BN + static_cast<int64_t>(tile0 + tj) * BN);
    }
    const int row0 = (tile0 + ti) * BM + wm * 64 + g;
    const int col0 = (MODE == MODE_TRSM ? col_origin : (tile0 + tj) * BN) + wn * 32 + 2 * t;
    const int col_limit = MODE == MODE_TRSM ? col_origin + BN : m_total;
    double acc[8][4][2];
    if (CPF) {
      // seed the accumulators from the prefetched tile: box = 16-column group, 128-byte rows, chunk ^ (row & 7)
      mbar_wait(bar_cfull, static_cast<uint32_t>(tile_n & 1));
      uint32_t dep = 0;
#pragma unroll
      for (int mi = 0; mi < 8; ++mi) {
        const uint32_t rbase = sC + static_cast<uint32_t>(wm * 64 + mi * 8 + g) * 128u;
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) {
          const uint32_t box = static_cast<uint32_t>(wn * 2 + (ni >> 1)) * kTileBytes;
          const uint32_t chunk = static_cast<uint32_t>((((ni & 1) * 4 + t) ^ g) << 4);
          const double2 v = lds_v2(rbase + box + chunk);
          acc[mi][ni][0] = v.x;
          acc[mi][ni][1] = v.y;
          if (mi == 7 && ni == 3) dep = static_cast<uint32_t>(__double2hiint(v.y));
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive_after_loads(bar_cempty, dep, rt_zero);  // sC may be refilled once all 8 warps are here
    } else if (MODE == MODE_UPDATE && !(VAR & 1)) {
      // start from C: the loads overlap the wait for the first operand stages
#pragma unroll
      for (int mi = 0; mi < 8; ++mi) {
        const int r = row0 + mi * 8;
        const double* crow = C + static_cast<int64_t>(r) * ldc;
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) {
          const int c = col0 + ni * 8;
          double2 v = make_double2(0.0, 0.0);
          if (r < m_total && c < col_limit) v = *reinterpret_cast<const double2*>(crow + c);
          acc[mi][ni][0] = v.x;
          acc[mi][ni][1] = v.y;
        }
      }
    } else {
#pragma unroll
      for (int mi = 0; mi < 8; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
    }

    for (int kb = 0; kb < kbn; ++kb) {
      // keep the ring kLead iterations ahead: the stage being refilled was released kStages - kLead
      // iterations ago, so this wait does not normally stall
      if (is_producer) {
        if (CPF && kb == 0) produce_c();  // next tile's C: lands during this tile's main loop
        produce_one();
      }
      mbar_wait(bar_full + stage * 8, phase);
      const uint32_t a_base = sA + stage * kTileBytes + offA;
      const uint32_t b_base = sB + stage * kTileBytes + offB;
      const uint32_t d_base = sD + stage * kDBytes;
      uint32_t dep = 0;
#pragma unroll
      for (int P = 0; P < 2; ++P) {
        double2 b[4];
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) b[ni] = lds_v2(b_base + ni * 1024 + sw[P]);
        if (SCALE) {
          const double2 d = lds_v2(d_base + dof[P]);
#pragma unroll
          for (int ni = 0; ni < 4; ++ni) {
            b[ni].x *= d.x;
            b[ni].y *= d.y;
          }
        }
        if (MODE == MODE_UPDATE) {
#pragma unroll
          for (int ni = 0; ni < 4; ++ni) {
            b[ni].x = -b[ni].x;
            b[ni].y = -b[ni].y;
          }
        }
        double2 a[8];
#pragma unroll
        for (int mi = 0; mi < 8; ++mi) a[mi] = lds_v2(a_base + mi * 1024 + sw[P]);
        // the last ld.shared of the iteration in program order (shared loads of a warp complete in order)
        if (P == 1) dep = static_cast<uint32_t>(__double2hiint(a[7].y));
#pragma unroll
        for (int mi = 0; mi < 8; ++mi)
#pragma unroll
          for (int ni = 0; ni < 4; ++ni) dmma884(acc[mi][ni][0], acc[mi][ni][1], a[mi].x, b[ni].x);
#pragma unroll
        for (int mi = 0; mi < 8; ++mi)
#pragma unroll
          for (int ni = 0; ni < 4; ++ni) dmma884(acc[mi][ni][0], acc[mi][ni][1], a[mi].y, b[ni].y);
      }
      __syncwarp();
      if (lane == 0) {
        if (VAR & 4)
          mbar_arrive(bar_empty + stage * 8);
        else
          mbar_arrive_after_loads(bar_empty + stage * 8, dep, rt_zero);
      }
      if (++stage == NS) {
        stage = 0;
        phase ^= 1u;
      }
      if (MODE == MODE_SYRK && FLUSH) {
        // Blocked accumulation, staggered and asynchronous: row group mi of the slab (8 doubles per thread) is folded
        // into C at K-blocks = fstride (mi + 1) - 1 - fwarp (mod fblocks).  Two K-blocks (~4 us) earlier each thread
        // requests its 4 x 16 bytes of old C with cp.async into a staging slot in shared memory -- no register and no
        // scoreboard is tied up while 128 DMMAs run (a plain prefetching load was waited for at the next branch
        // merge: 1.5 us per flush, all eight warps at once) -- and the flush itself is 4 LDS + 8 DADD + 4 STG.
        // Each element of the tile belongs to exactly one thread, so the read-modify-write needs no
        // synchronisation; the first flush of a tile is a plain store (C is not read before it is written).
        // Addresses are rebuilt from row0 behind an opaque move so that the compiler does not keep eight row
        // pointers alive across the main loop (they cost the fragment double-buffering its registers).
        const int ph = (kb + fwarp) & (fblocks - 1);
        const int sub = ph & (fstride - 1);
        const uint32_t grp_bit = 1u << (ph >> fshift);  // a bit test per row group keeps acc[] statically indexed
        if (sub == fstride - 3 && kb + 2 >= fblocks) {
          int rbase = row0;
          asm volatile("" : "+r"(rbase));
#pragma unroll
          for (int mi = 0; mi < 8; ++mi)
            if ((grp_bit >> mi) & 1u) {
              const int r = rbase + mi * 8;
              const double* crow = Cd + static_cast<int64_t>(r) * ldd;
#pragma unroll
              for (int ni = 0; ni < 4; ++ni) {
                const int cc = col0 + ni * 8;
                if (r < m_total && cc < col_limit) cp_async_16(sF + ni * (kThreads * 16), crow + cc);
              }
            }
          cp_async_commit();
        } else if (sub == fstride - 1 && kb + 1 < kbn) {
          const bool first = kb < fblocks;
          int rbase = row0;
          asm volatile("" : "+r"(rbase));
          if (!first) cp_async_wait_all();
#pragma unroll
          for (int mi = 0; mi < 8; ++mi)
            if ((grp_bit >> mi) & 1u) {
              const int r = rbase + mi * 8;
              double* crow = Cd + static_cast<int64_t>(r) * ldd;
#pragma unroll
              for (int ni = 0; ni < 4; ++ni) {
                const int cc = col0 + ni * 8;
                double2 v = make_double2(acc[mi][ni][0], acc[mi][ni][1]);
                if (!first) {
                  const double2 o = lds_v2(sF + ni * (kThreads * 16));
                  v.x += o.x;
                  v.y += o.y;
                }
                if (r < m_total && cc < col_limit) *reinterpret_cast<double2*>(crow + cc) = v;
                acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
              }
            }
        }
      }
    }

    // ------------------------------------------------------------ epilogue: store-only, each warp its own slab
#pragma unroll
    for (int mi = 0; mi < 8; ++mi) {
      const int r = row0 + mi * 8;
      if (r < m_total) {
        double* crow = Cd + static_cast<int64_t>(r) * ldd;
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) {
          const int c = col0 + ni * 8;
          if (c < col_limit) {
            double2 v = make_double2(acc[mi][ni][0], acc[mi][ni][1]);
            // MODE_SYRK: row group mi has been folded into C before iff its first flush point lies inside the tile
            if ((MODE == MODE_UPDATE && (VAR & 1)) ||
                (MODE == MODE_SYRK && FLUSH && fstride * (mi + 1) - fwarp < kbn)) {
              const double2 o = *reinterpret_cast<const double2*>(crow + c);
              v.x += o.x;
              v.y += o.y;
            }
            *reinterpret_cast<double2*>(crow + c) = v;
          }
        }
      }
    }
    if (VAR & 2) __syncthreads();
  }
}

// Sum of the K-parts of the tail tiles (TailSplit), in part order, into C: one CTA per tail tile.
__global__ void __launch_bounds__(256)
syrk_tail_fixup_kernel(double* __restrict__ C, int64_t ldc, int m_total, int ntr, TailSplit ts) {
  int ti, tj;
  band_decode(ts.n_full + blockIdx.x, ntr, &ti, &tj);
  const double* w = ts.ws + static_cast<int64_t>(blockIdx.x) * ts.nsplit * (BM * BN);
  for (int e = threadIdx.x; e < BM * BN / 2; e += 256) {
    const int r = ti * BM + (e >> 6), c = tj * BN + 2 * (e & 63);
    if (r >= m_total || c >= m_total) continue;  // m_total and ldc are even: a pair is in or out as a whole
    double2 sum = make_double2(0.0, 0.0);
    for (int p = 0; p < ts.nsplit; ++p) {
      const double2 v = *reinterpret_cast<const double2*>(w + static_cast<int64_t>(p) * (BM * BN) + 2 * e);
      sum.x += v.x;
      sum.y += v.y;
    }
    *reinterpret_cast<double2*>(C + static_cast<int64_t>(r) * ldc + c) = sum;
  }
}

// ------------------------------------------------------------------ host: tensor maps
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int get_encode_fn(EncodeTiledFn* out) {
  static std::mutex mu;
  static EncodeTiledFn fn = nullptr;
  std::lock_guard<std::mutex> g(mu);
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    LPB_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
    if (q != cudaDriverEntryPointSuccess || !p) {
      set_last_error("cuTensorMapEncodeTiled not available from the driver");
      return LPB_ERR_CUDA;
    }
    fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  *out = fn;
  return LPB_OK;
}

int make_tmap(CUtensorMap* tm, const double* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows,
              uint32_t box_cols, bool swizzle) {
  EncodeTiledFn fn;
  LPB_TRY(get_encode_fn(&fn));
  const cuuint64_t dims[2] = {cols, rows};
  const cuuint64_t strides[1] = {ld * sizeof(double)};
  const cuuint32_t box[2] = {box_cols, box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE,
                        swizzle ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error("cuTensorMapEncodeTiled failed (%d): base=%p rows=%llu cols=%llu ld=%llu", (int)r, (const void*)base,
                   (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)ld);
    return LPB_ERR_CUDA;
  }
  return LPB_OK;
}

template <int MODE, bool SCALE, int VAR = 0>
int launch_dmma(LaunchCtx& lc, const CUtensorMap& tmA, const CUtensorMap& tmB, double* C, int64_t ldc, int m_total,
                int tile0, int ntr, int k_begin, int nkb, int col_origin, int shape = SHAPE_TRI,
                OwnedCols own = OwnedCols{1, 0, 0}, int p_row0 = -1, TailSplit ts = TailSplit{0, 1, nullptr}) {
  static PerDeviceOnce once;  // one per template instantiation
  auto kern = syrk_dmma_kernel<MODE, SCALE, VAR>;
  LPB_TRY(once.run([&](int) -> int {
    LPB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Plan<MODE, VAR>::kAlloc));
    return LPB_OK;
  }));
  const int ntiles = (MODE == MODE_TRSM || shape == SHAPE_COL) ? ntr
                     : (shape == SHAPE_COLB ? ntr - 1 : (shape == SHAPE_OWNED ? own.count() : ntr * (ntr + 1) / 2));
  if (ntiles <= 0) return LPB_OK;
  if (ts.nsplit <= 1) ts = TailSplit{ntiles, 1, nullptr};  // every item is a whole tile
  int grid = ntiles < kNumSMs ? ntiles : kNumSMs;
  if (MODE == MODE_UPDATE && lc.update_grid_cap > 0 && grid > lc.update_grid_cap) grid = lc.update_grid_cap;
  cudaStream_t st = lc.launch_on_side ? lc.side_stream : lc.stream;
  kern<<<grid, kThreads, Plan<MODE, VAR>::kAlloc, st>>>(tmA, tmB, C, ldc, m_total, tile0, ntr, k_begin, nkb, col_origin, shape, own,
                                           p_row0, ts);
  lc.launches++;
  LPB_CUDA(cudaGetLastError());
  return LPB_OK;
}

}  // namespace

int k_syrk_dmma(LaunchCtx& lc, int64_t m, int64_t n, const double* A, int64_t lda, const double* d, double* Cmat,
                int64_t ldc) {
  if ((lda & 1) || (ldc & 1) || (reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(Cmat) & 15) ||
      (d && (reinterpret_cast<uintptr_t>(d) & 15))) {
    set_last_error("syrk_dmma: A, d, M must be 16-byte aligned with even leading dimensions");
    return LPB_ERR_BAD_ARGUMENT;
  }
  if (m <= 0 || n <= 0) return LPB_OK;
  CUtensorMap tmA, tmD;
  LPB_TRY(make_tmap(&tmA, A, (uint64_t)m, (uint64_t)n, (uint64_t)lda, BM, BK, true));
  if (d)
    LPB_TRY(make_tmap(&tmD, d, 1, (uint64_t)n, (uint64_t)round_up(n, 2), 1, BK, false));
  else
    tmD = tmA;
  const int ntr = (int)ceil_div(m, BM);
  const int nkb = (int)ceil_div(n, BK);
  if (lc.syrk_chain) {  // option "syrk_chain" = 1: no blocked accumulation (accuracy / timing comparison)
    if (d) return launch_dmma<MODE_SYRK, true, 16>(lc, tmA, tmD, Cmat, ldc, (int)m, 0, ntr, 0, nkb, 0);
    return launch_dmma<MODE_SYRK, false, 16>(lc, tmA, tmD, Cmat, ldc, (int)m, 0, ntr, 0, nkb, 0);
  }
  const int fb = lc.syrk_flush_blocks;  // option "syrk_flush_blocks": K-blocks between two flushes (power of two >= 32)
  // Tail split (option "syrk_tail_split", default on): worth it when the last wave is at most 3/4 full.  The number of
  // parts minimises the length of the tail, ceil(rem * S / SMs) / S tile-times, plus a little for the extra traffic.
  TailSplit ts{0, 1, nullptr};
  const int ntiles = ntr * (ntr + 1) / 2;
  const int rem = ntiles % kNumSMs;
  if (lc.syrk_tail_split && ntiles > kNumSMs && rem > 0 && 4 * rem <= 3 * kNumSMs && nkb >= 64 && !(m & 1)) {
    int best_s = 1;
    double best = 1.0;
    for (int sp = 2; sp <= 8; ++sp) {
      const double len = (double)ceil_div((int64_t)rem * sp, (int64_t)kNumSMs) / sp + 0.01 * sp;
      if (len < best - 1e-9) {
        best = len;
        best_s = sp;
      }
    }
    if (best_s > 1) {
      const int64_t need = (int64_t)rem * best_s * BM * BN;
      if (lc.syrk_ws_cap < need) {
        if (lc.syrk_ws) LPB_CUDA(cudaFree(lc.syrk_ws));
        lc.syrk_ws = nullptr;
        lc.syrk_ws_cap = 0;
        void* p = nullptr;
        LPB_CUDA(cudaMalloc(&p, sizeof(double) * (size_t)need));
        lc.syrk_ws = static_cast<double*>(p);
        lc.syrk_ws_cap = need;
      }
      ts = TailSplit{ntiles - rem, best_s, lc.syrk_ws};
    }
  }
  if (ts.nsplit > 1) {
    if (d)
      LPB_TRY((launch_dmma<MODE_SYRK, true, 32>(lc, tmA, tmD, Cmat, ldc, (int)m, 0, ntr, 0, nkb, fb, SHAPE_TRI, OwnedCols{1, 0, 0}, -1, ts)));
    else
      LPB_TRY((launch_dmma<MODE_SYRK, false, 32>(lc, tmA, tmD, Cmat, ldc, (int)m, 0, ntr, 0, nkb, fb, SHAPE_TRI, OwnedCols{1, 0, 0}, -1, ts)));
  } else if (d) {
    LPB_TRY((launch_dmma<MODE_SYRK, true>(lc, tmA, tmD, Cmat, ldc, (int)m, 0, ntr, 0, nkb, fb)));
  } else {
    LPB_TRY((launch_dmma<MODE_SYRK, false>(lc, tmA, tmD, Cmat, ldc, (int)m, 0, ntr, 0, nkb, fb)));
  }
  if (ts.nsplit > 1) {
    syrk_tail_fixup_kernel<<<rem, 256, 0, lc.stream>>>(Cmat, ldc, (int)m, ntr, ts);
    lc.launches++;
    LPB_CUDA(cudaGetLastError());
  }
  return LPB_OK;
}

int k_trailing_update_dmma(LaunchCtx& lc, int64_t m, double* Mat, int64_t ldm, int64_t k0, int64_t kb) {
  const int64_t row0 = k0 + kb;
  if (row0 >= m) return LPB_OK;
  if ((row0 % BM) || (kb % BK) || (ldm & 1) || (reinterpret_cast<uintptr_t>(Mat) & 15)) {
    set_last_error("trailing_update_dmma: panel origin must be a multiple of %d, width a multiple of %d", BM, BK);
    return LPB_ERR_BAD_ARGUMENT;
  }
  CUtensorMap tm;
  LPB_TRY(make_tmap(&tm, Mat, (uint64_t)m, (uint64_t)m, (uint64_t)ldm, BM, BK, true));
  const int ntr = (int)ceil_div(m - row0, BM);
  if (lc.update_impl == 2)
    return launch_dmma<MODE_UPDATE, false, 1>(lc, tm, tm, Mat, ldm, (int)m, (int)(row0 / BM), ntr, (int)k0,
                                              (int)(kb / BK), 0);
  if (lc.update_impl == 3)
    return launch_dmma<MODE_UPDATE, false, 2>(lc, tm, tm, Mat, ldm, (int)m, (int)(row0 / BM), ntr, (int)k0,
                                              (int)(kb / BK), 0);
  if (lc.update_impl == 4)
    return launch_dmma<MODE_UPDATE, false, 4>(lc, tm, tm, Mat, ldm, (int)m, (int)(row0 / BM), ntr, (int)k0,
                                              (int)(kb / BK), 0);
  if (lc.update_impl == 5)
    return launch_dmma<MODE_UPDATE, false, 6>(lc, tm, tm, Mat, ldm, (int)m, (int)(row0 / BM), ntr, (int)k0,
                                              (int)(kb / BK), 0);
  return launch_dmma<MODE_UPDATE, false>(lc, tm, tm, Mat, ldm, (int)m, (int)(row0 / BM), ntr, (int)k0, (int)(kb / BK),
                                         0);
}

// Trailing update with panel [k0, k0 + kb) restricted to a tile list: rows / columns from tile `tile0` on;
// single_col: only block column tile0 (look-ahead); otherwise the block columns J >= tile0 with
// J % own_mod == own_rem (own_mod == 1: all of them).
int k_trailing_update_part(LaunchCtx& lc, int64_t m, double* Mat, int64_t ldm, int64_t k0, int64_t kb, int tile0,
                           int single_col, int own_mod, int own_rem, const double* packed, int64_t packed_rows) {
  int ntr = (int)ceil_div(m, BM) - tile0;
  if (ntr <= 0) return LPB_OK;
  if ((kb % BK) || (ldm & 1) || (reinterpret_cast<uintptr_t>(Mat) & 15) || (int64_t)tile0 * BM < k0 + kb) {
    set_last_error("trailing_update_part: bad panel / tile origin");
    return LPB_ERR_BAD_ARGUMENT;
  }
  CUtensorMap tm, tmc;
  int k_begin = (int)k0, p_row0 = -1;
  LPB_TRY(make_tmap(&tmc, Mat, (uint64_t)m, (uint64_t)m, (uint64_t)ldm, BM, BK, true));  // C tiles (VAR bit 3)
  if (packed) {  // the panel comes from a packed buffer (128 doubles per row), see syrk_dmma_kernel::p_row0
    LPB_TRY(make_tmap(&tm, packed, (uint64_t)packed_rows, (uint64_t)BN, (uint64_t)BN, BM, BK, true));
    k_begin = 0;
    p_row0 = (int)k0;
  } else {
    tm = tmc;
  }
  // single_col: 1 = the whole block column tile0, 2 = only its diagonal tile, 3 = the column without its diagonal tile
  int shape = SHAPE_TRI;
  if (single_col == 1) shape = SHAPE_COL;
  if (single_col == 2) {
    shape = SHAPE_COL;
    ntr = 1;
  }
  if (single_col == 3) shape = SHAPE_COLB;
  if (!single_col && own_mod > 1) shape = SHAPE_OWNED;
  const OwnedCols own = shape == SHAPE_OWNED ? OwnedCols::make(own_mod, own_rem, tile0, ntr) : OwnedCols{1, 0, 0};
  if (lc.update_impl == 2)  // products accumulated from zero, C read-modify-written once per panel (see VAR)
    return launch_dmma<MODE_UPDATE, false, 1>(lc, tm, tm, Mat, ldm, (int)m, tile0, ntr, k_begin, (int)(kb / BK), 0, shape,
                                              own, p_row0);
  if (lc.update_impl == 6)  // C tile prefetched through shared memory (VAR bit 3): measured NOT faster -- C3 potrf
    // 53.6 ms against 52.6 ms, C2 2.84 against 2.79 (profiles/README.md): the head-of-tile load already overlapped the
    // wait for the first operand stage, and the ring shrinks from 5 to 3 stages to make room.  Kept as an option.
    return launch_dmma<MODE_UPDATE, false, 8>(lc, tm, tmc, Mat, ldm, (int)m, tile0, ntr, k_begin, (int)(kb / BK), 0,
                                              shape, own, p_row0);
  return launch_dmma<MODE_UPDATE, false>(lc, tm, tm, Mat, ldm, (int)m, tile0, ntr, k_begin, (int)(kb / BK), 0, shape, own,
                                         p_row0);
}

// Panel TRSM as a GEMM: Mat[row0.., k0..k0+128) <- Mat[row0.., k0..k0+128) * Linv^T, row0 = k0 + 128,
// Linv = inv(L_kk) stored dense 128 x 128 (ld 128, upper triangle zero).
int k_trsm_dmma(LaunchCtx& lc, int64_t m, double* Mat, int64_t ldm, int64_t k0, const double* Linv) {
  const int64_t row0 = k0 + BN;
  if (row0 >= m) return LPB_OK;
  if ((row0 % BM) || (ldm & 1) || (reinterpret_cast<uintptr_t>(Mat) & 15) || (reinterpret_cast<uintptr_t>(Linv) & 15)) {
    set_last_error("trsm_dmma: panel origin must be a multiple of %d", BM);
    return LPB_ERR_BAD_ARGUMENT;
  }
  CUtensorMap tm, tl;
  LPB_TRY(make_tmap(&tm, Mat, (uint64_t)m, (uint64_t)m, (uint64_t)ldm, BM, BK, true));
  LPB_TRY(make_tmap(&tl, Linv, (uint64_t)BN, (uint64_t)BN, (uint64_t)BN, BN, BK, true));
  const int ntr = (int)ceil_div(m - row0, BM);
  return launch_dmma<MODE_TRSM, false>(lc, tm, tl, Mat, ldm, (int)m, (int)(row0 / BM), ntr, (int)k0, BN / BK, (int)k0);
}

// ------------------------------------------------------------------ DMMA issue peak (measurement only)
__global__ void __launch_bounds__(256)
dmma_peak_kernel(double* out, int iters, double a0, double b0) {
  double c[32][2];
#pragma unroll
  for (int i = 0; i < 32; ++i) c[i][0] = c[i][1] = threadIdx.x * 1e-3 + i;
  const double a = a0 + threadIdx.x * 1e-9, b = b0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 32; ++i) dmma884(c[i][0], c[i][1], a, b);
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 32; ++i) s += c[i][0] + c[i][1];
  if (s == 12345.678) out[0] = s;  // keeps the loop alive; never true in practice
}

int k_dmma_peak(LaunchCtx& lc, double seconds, double* tflops) {
  if (seconds < 0.05) seconds = 0.05;
  if (seconds > 2.0) seconds = 2.0;
  const int iters = 4000;
  const double flop_per_launch = 2.0 * 256 * 32 * iters * 8.0 * kNumSMs;  // m8n8k4 = 256 MAC per warp instruction
  cudaEvent_t e0, e1;
  LPB_CUDA(cudaEventCreate(&e0));
  LPB_CUDA(cudaEventCreate(&e1));
  auto run = [&](int launches, float* ms) -> int {
    LPB_CUDA(cudaEventRecord(e0, lc.stream));
    for (int i = 0; i < launches; ++i) dmma_peak_kernel<<<kNumSMs, 256, 0, lc.stream>>>(lc.red_partials, iters, 1.0000001, 1e-9);
    LPB_CUDA(cudaEventRecord(e1, lc.stream));
    LPB_CUDA(cudaEventSynchronize(e1));
    LPB_CUDA(cudaEventElapsedTime(ms, e0, e1));
    lc.launches += launches;
    return LPB_OK;
  };
  float ms = 0.f;
  int rc = run(4, &ms);  // warm-up + calibration (~2 ms per launch)
  if (rc == LPB_OK) {
    int launches = (int)(seconds * 1e3 / (ms / 4.0 > 1e-3 ? ms / 4.0 : 1e-3));
    if (launches < 8) launches = 8;
    if (launches > 4000) launches = 4000;
    rc = run(launches, &ms);
    if (rc == LPB_OK && tflops) *tflops = flop_per_launch * launches / (ms * 1e-3) * 1e-12;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  return rc;
}

// ------------------------------------------------------------------ plain DFMA reference kernel
// Used by the parity tests to bisect the DMMA/TMA kernel; never selected by default.
template <bool ACCUM>
__global__ void __launch_bounds__(256)
syrk_simple_kernel(int m_total, int row_origin, int k_begin, int k_end, const double* __restrict__ A, int64_t lda,
                   const double* __restrict__ d, double* __restrict__ C, int64_t ldc) {
  __shared__ double sa[16][65], sb[16][65];
  const int bi = blockIdx.y, bj = blockIdx.x;
  if (bj > bi) return;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4, tid = threadIdx.x;
  double acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
  for (int k0 = k_begin; k0 < k_end; k0 += 16) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int idx = tid + e * 256, r = idx >> 4, k = idx & 15;
      const int gk = k0 + k, gi = row_origin + bi * 64 + r, gj = row_origin + bj * 64 + r;
      const bool kin = gk < k_end;
      sa[k][r] = (kin && gi < m_total) ? A[(int64_t)gi * lda + gk] : 0.0;
      sb[k][r] = (kin && gj < m_total) ? A[(int64_t)gj * lda + gk] * (d ? d[gk] : 1.0) : 0.0;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      double a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = sa[k][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = sb[k][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] += a[i] * b[j];
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = row_origin + bi * 64 + ty * 4 + i;
    if (r >= m_total) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = row_origin + bj * 64 + tx * 4 + j;
      if (c >= m_total) continue;
      double* p = C + (int64_t)r * ldc + c;
      *p = ACCUM ? (*p - acc[i][j]) : acc[i][j];
    }
  }
}

int k_syrk_simple(LaunchCtx& lc, int64_t m, int64_t n, const double* A, int64_t lda, const double* d, double* Cmat,
                  int64_t ldc) {
  if (m <= 0) return LPB_OK;
  const int nt = (int)ceil_div(m, 64);
  syrk_simple_kernel<false><<<dim3(nt, nt), 256, 0, lc.stream>>>((int)m, 0, 0, (int)n, A, lda, d, Cmat, ldc);
  lc.launches++;
  LPB_CUDA(cudaGetLastError());
  return LPB_OK;
}

int k_trailing_update_simple(LaunchCtx& lc, int64_t m, double* Mat, int64_t ldm, int64_t k0, int64_t kb) {
  const int64_t row0 = k0 + kb;
  if (row0 >= m) return LPB_OK;
  const int nt = (int)ceil_div(m - row0, 64);
  syrk_simple_kernel<true><<<dim3(nt, nt), 256, 0, lc.stream>>>((int)m, (int)row0, (int)k0, (int)(k0 + kb), Mat, ldm,
                                                                 nullptr, Mat, ldm);
  lc.launches++;
  LPB_CUDA(cudaGetLastError());
  return LPB_OK;
}

}  // namespace lpb
