// common.cuh -- shared definitions for liblpb200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <mutex>

#include "../../include/lpb200.h"

namespace lpb {

void set_last_error(const char* fmt, ...);
const char* get_last_error();

#define LPB_CUDA(call)                                                                         \
  do {                                                                                         \
    cudaError_t e__ = (call);                                                                  \
    if (e__ != cudaSuccess) {                                                                  \
      ::lpb::set_last_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
      return LPB_ERR_CUDA;                                                                     \
    }                                                                                          \
  } while (0)

#define LPB_TRY(call)              \
  do {                             \
    int rc__ = (call);             \
    if (rc__ != LPB_OK) return rc__; \
  } while (0)

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs

// One-time per-DEVICE setup (cudaFuncSetAttribute and occupancy queries apply to the current device only, and
// several host threads may own contexts on several devices of one process): run(f) calls f(device) the first
// time it is reached on each device, under a lock.
constexpr int kMaxDevices = 64;
struct PerDeviceOnce {
  std::mutex mu;
  bool done[kMaxDevices] = {};
  template <class F>
  int run(F&& f) {
    int dev = 0;
    LPB_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= kMaxDevices) return f(dev);
    std::lock_guard<std::mutex> g(mu);
    if (done[dev]) return LPB_OK;
    const int rc = f(dev);
    if (rc == LPB_OK) done[dev] = true;
    return rc;
  }
};

inline int64_t round_up(int64_t v, int64_t a) { return (v + a - 1) / a * a; }
inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ------------------------------------------------------------------ reductions
// Deterministic two-stage reductions: each block writes one partial per value into
// partials[val * kMaxRedBlocks + block]; `finalize_reduce` folds them in block order.
constexpr int kMaxRedVals = 8;
constexpr int kMaxRedBlocks = 1024;
enum RedOp : int { kRedSum = 0, kRedMin = 1 };

// ------------------------------------------------------------------ per-context launch state
struct LaunchCtx {
  cudaStream_t stream = nullptr;
  int64_t launches = 0;       // kernels launched by this library on this context
  double* red_partials = nullptr;  // kMaxRedVals * kMaxRedBlocks
  double* red_out = nullptr;       // device, kMaxRedVals + 1: the last word is the context's fault word
  double* red_host = nullptr;      // pinned, kMaxRedVals + 1
  unsigned long long* fault_dev = nullptr;  // = red_out + kMaxRedVals: raised by a kernel that gave up waiting
  double* gemv_partials = nullptr; // gemv_t row-chunk partials
  int64_t gemv_partials_cap = 0;   // in doubles
  double* chol_ws = nullptr;       // inv(L_kk) blocks of the last k_potrf + solve scratch (cholesky.cu)
  int64_t chol_ws_cap = 0;         // in doubles
  int64_t ws_m = -1;               // order of the matrix the hand-off words of chol_ws are laid out for
  int64_t linv_valid_m = -1;       // order of the matrix whose block inverses chol_ws holds
  const double* linv_mat = nullptr; // ... and its address
  bool linv_full = false;          // chol_ws holds the full 128 x 128 inverses (else only their 16 x 16 diagonal blocks)
  int solve_epoch = 0;             // flag value of the next pipelined solve (cholesky.cu)
  int solve_impl = 0;              // 0 = solve_ll_kernel, 1 = one launch per block step, 2 = substitution, 3 = flag-based pipelined
  int sync_each_launch = 0;        // debug: stream-synchronise after every launch of k_potrf
  int syrk_flush_blocks = 32;      // K1: K-blocks (16 columns each) summed in registers between two flushes into C
  int syrk_tail_split = 1;         // K1: cut the tiles of a short last wave into K-parts (dmma_gemm.cu: TailSplit)
  double* syrk_ws = nullptr;       // ... their private 128 x 128 slots
  int64_t syrk_ws_cap = 0;
  int syrk_chain = 0;              // 1 = K1 sums the whole K extent in one register chain (no blocked accumulation)
  int potf2_impl = 0;              // 1 = textbook potf2 (sqrt, divisions, no inverses): needs trsm_impl 1; solves substitute
  int trsm_impl = 0, update_impl = 0;  // bisecting knobs of k_potrf: 1 = plain DFMA kernel for that step
  int solve_grid_cap = 0;          // > 0: cap the pipelined solve's grid (tests: several block rows per CTA)
  // column-sharded contexts: the distributed factorisation (k_potrf_dist) broadcasts panels over NCCL
  void* nccl_comm = nullptr;       // ncclComm_t of this process, or null
  int rank = 0, world = 1;
  int potrf_dist = 2;              // 0 = replicated on every rank, 1 = one broadcast per panel, 2 = two broadcasts +
                                   // side-stream potf2 + packed panels (dist_schedule.hpp, schedule v2)
  double* panel_buf = nullptr;     // packed panel + inverted diagonal block, the broadcast payload
  int64_t panel_buf_cap = 0;       // in doubles
  // single-GPU look-ahead of k_potrf: potf2 of panel k+1 runs on `side_stream` beside the trailing update of panel k
  int potrf_lookahead = 1;         // 0 = strictly sequential panels on `stream`
  cudaStream_t side_stream = nullptr;
  cudaEvent_t ev_col[2] = {nullptr, nullptr}, ev_pan[2] = {nullptr, nullptr};
  bool launch_on_side = false;     // the next trailing-update launch goes to side_stream (k_potrf_dist2)
  cudaEvent_t ev_dist[4] = {nullptr, nullptr, nullptr, nullptr};  // kEvSmall / kEvPotf2 of dist_schedule.hpp
  double* panel_slot[2] = {nullptr, nullptr};  // packed panels of the two-broadcast distributed factorisation
  int64_t panel_slot_cap = 0;      // doubles per slot
  // Peer-memory panel hand-off (k_potrf_dist2, optional): ONE allocation per rank = [flags][ring of 2 x world panel
  // slots], exported with cudaIpcGetMemHandle and mapped by every other rank of the box.  The owner of panel k writes
  // the 128 rows the next owner needs straight into every peer's slot k % ring over NVLink and raises the peer's flag;
  // see peer_push_kernel in cholesky.cu.
  static constexpr int kMaxPeers = 8;
  double* ring_base = nullptr;     // this rank's allocation (cudaMalloc)
  int64_t ring_slot_doubles = 0;   // capacity of one slot
  int ring_slots = 0;              // 2 x world
  double* peer_base[kMaxPeers] = {};  // the same allocation of rank g, mapped here (own rank: ring_base)
  bool peer_mapped = false;        // k_peer_import succeeded on this rank
  bool peer_ready = false;         // ... and the hand-off is switched on (option "peer_panels"): bcast_small goes through peer memory
  uint32_t dist_epoch = 0;         // factorisations done through k_potrf_dist2 (upper half of the flag values)
  int update_grid_cap = 0;         // > 0: CTAs of the trailing update (leaves SMs free for the side stream)
  int* info_dev = nullptr;         // potrf info flag
  int* info_host = nullptr;        // pinned
};

}  // namespace lpb
