// spectral_reduce.cuh - data structures shared by the spectral reduction
// (LBL spectra -> per-g-point averages) host code and kernels.
#pragma once
#include "common.cuh"
#include "trans_core.cuh"

// Table form of 1 - exp(-k D tau) in reduce2_kernel (trans_core.cuh): 2^ECCKD_TAB_LOGN entries replicated
// 2^ECCKD_TAB_LOGR times, minimax q of degree ECCKD_TAB_QDEG.  Measured on the B200 (19 x 53 x 2^23,
// kernel ms): 64 x 16 copies, Taylor degree 5 (10 FP64 instructions per value) 17.25; 256 x 8, Taylor
// degree 4 (9) 16.10; 128 x 16, minimax q3 (9) 16.42; 1024 x 2, minimax q2 (8) 16.17; 2048 x 1, minimax
// q2 (8) 15.71 - the consumer warps are bound by the FP64 pipe while they evaluate, so what pays is
// fewer FP64 instructions, not fewer shared-memory wavefronts.
#ifndef ECCKD_TAB_LOGN
#define ECCKD_TAB_LOGN 11
#define ECCKD_TAB_LOGR 0
#define ECCKD_TAB_QDEG 2
#endif

namespace ecckd {

constexpr int SLOTS = 16;          // plan slots per lane per round
constexpr int CH_MAX = 480;        // longest chunk (wavenumbers); chunk bytes stay a multiple of 16
constexpr int STAGE_BYTES = CH_MAX * 8;
constexpr int STAGE3_BYTES = STAGE_BYTES + 16;  // room for the 8-byte shift of a row that is not 16-byte aligned
constexpr int RPW = 2;             // rows per consumer warp (they share the lane's offsets and weights)
constexpr int NSETS = 2;           // stage sets per consumer warp: chunk visit k uses set k & 1
constexpr int NSW = 4;             // ring slots of (weight tile, plan round) per CTA
constexpr int NPW = 2;             // producer (weight) warps per CTA
constexpr int NW_MAX = 10;         // consumer warps per CTA of reduce_kernel: 12 warps x 168 registers fill the register file
// reduce2_kernel.  Measured on the B200 (19 x 53 x 2^23): 11 consumer warps + 2 producers (rows 2,2,1 /
// 2,2,1 / 2,2,1 / 2,2+producers over the four schedulers instead of 6 / 5 / 4 / 4) is SLOWER, 19.6 ms
// against 18.6 ms for 10 + 2 built with the same 416-thread bound: the round time is the latency of a
// two-row warp, not the issue load of its scheduler, and the 152-register cap makes ptxas serialise
// the polynomial chains (17.1 ms with 12 warps at 168 registers).
constexpr int NW_MAX2 = 10;
constexpr int GROUP_ROWS = NW_MAX2 * RPW;
// lane meta word of the plan: bits 0-11 g point (0xFFF = lane unused), 12-16 valid slots (0..16),
// 17-21 "lane + 2^s belongs to the same g point in this round" for shuffle step s, 22 head lane of its g point
constexpr unsigned NO_G = 0xFFFu;
constexpr int MAX_NG = 4094;
// Windowed plans (reduce2_kernel): the g points of one round span fewer than WIN consecutive
// indices, and PlanRound::win holds the base of the round's window [win, win + WIN).
constexpr int WIN = 8;

// One round of the plan (32 lanes x 16 slots): per lane 16 byte offsets into the chunk (padding slots
// repeat slot 0: neutral for min/max; the producers give them zero weight) and
// the lane meta word; per round the number of shuffle steps, its chunk and first/last-of-chunk flags.
// 1168 bytes, copied to shared memory with one bulk copy.
struct PlanRound {
  uint16_t off[SLOTS][32];  // byte offset of (slot, lane) into the chunk
  uint32_t meta[32];
  uint32_t steps;
  uint32_t chunk;
  uint32_t flags;     // bit 0: first round of its chunk, bit 1: last round of its chunk
  uint32_t win;       // windowed plans: first g point of the round's window (unused otherwise)
};
static_assert(sizeof(PlanRound) == 1168, "PlanRound layout");
constexpr int WPERM_BYTES = SLOTS * 32 * 8;  // weights of a round in (slot, lane) order, zero for padding slots
constexpr int RING_SLOT_BYTES = WPERM_BYTES + (int)sizeof(PlanRound);

// internal per-row method ids (public enum resolved per layer for hybrid)
enum { M_LINEAR = 0, M_TRANS = 1, M_SQRT = 2, M_LOG = 3, M_SUMW = 4 };
// kernel "f kind": which transform the streaming kernel applies
enum { F_LINEAR = 0, F_TRANS = 1, F_SQRT = 2, F_LOG = 3, F_HYBRID = 4, F_NONE = 5 };
enum { W_MATRIX = 0, W_VECTOR = 1, W_PLANCK = 2 };

// One (spectrum, layer) row of the streaming reduction.
struct RowDesc {
  const double* od;   // row base (nwav doubles)
  int32_t method;     // M_*
  int32_t wrow;       // weight row (layer index / temperature index)
  int32_t row;        // index of this row in the partial-sum arrays
  int32_t pad_;
};

// Where and how one row's result is written by the finalize kernel.
struct OutDesc {
  double* out;        // &molar_abs[layer][0]
  double* out_min;    // may be NULL
  double* out_max;    // may be NULL
  int32_t* warn;      // int32[ng] bit flags, may be NULL
  double factor;      // (g*0.001*M_air)/vmr_ref, or <= 0 for "raw optical depth"
  double dp;          // p_hl[l+1]-p_hl[l]
  int32_t method;     // M_*
  int32_t pad_;
};

// Per-(row, g) partial results, the only data that crosses GPUs when the
// wavenumber axis is sharded: sums[row][g] = {sum f*w, sum w, count(tau<=0)}
// (allreduce SUM) and minmax[row][g] = {min tau, -max tau} (allreduce MIN).
constexpr int NSUM = 3;
constexpr int NMINMAX = 2;

// Up to GROUP_ROWS rows that share one weight row.  Row i of a group is handled by consumer warp
// i % nw as its row i / nw.
struct GroupDesc {
  const double* od[GROUP_ROWS];
  int32_t row[GROUP_ROWS];     // index in the partial-sum arrays
  uint8_t is_log[GROUP_ROWS];  // row uses logarithmic averaging (hybrid method)
  uint8_t bulk_ok[GROUP_ROWS]; // row base is 16-byte aligned: bulk async copies allowed
  int32_t nrows;
  int32_t wrow;
};

struct ReduceParams {
  const GroupDesc* groups;
  int ngroups;
  int nw;                  // consumer warps per CTA
  int ncta;                // grid size P
  int maxg;                // partial slots per CTA ((window, group) pairs a CTA can touch)
  int ch;                  // chunk length of the plan
  int nwin;                // wavenumber windows: items are ordered (window, group, chunk) so that the
  int cw;                  //   per-wavenumber data of a window (cw chunks) stay in L2 across groups
  int64_t nchunks;
  int64_t nwav;
  int ng;
  const uint32_t* round_off;  // [nchunks+1]
  const PlanRound* rounds;    // [nrounds]
  const double* weight;       // MATRIX: [wrow][nwav] (stride w_stride); VECTOR: [nwav]
  int64_t w_stride;
  int weight_bulk_ok;
  const double* temperature;  // PLANCK: device [n weight rows], (h/k)/T
  const double* temperature2; // PLANCK, reduce2_kernel: (h/k)(100 c)/T
  const double* pref;         // PLANCK, reduce2_kernel: dnu c1 (nu c)^3 per wavenumber (planck_prefactor_kernel)
  const double* wn;
  const double* dwn;
  TransConst tc;              // constants of 1 - exp(-kd tau), kd = LW_DIFFUSIVITY*OD_SCALING*k
  TransConst256 tc256;        // the same for the 256-entry shared-memory table of the consumer warps
  TransConst64 tc64;          // the same for the replicated 64-entry table of reduce2_kernel
  TransConstT tct;            // the same for the general table form (trans_core.cuh)
  const double* exp2tab;      // 2^(j/N), N entries (ecckd_ctx::d_exp2tab)
  int max_rows;               // reduce2_kernel: rows of the largest group (data stages are packed by row)
  int float_row;              // reduce2_kernel: groups of 2 nw - 1 rows pass their last row between the last two warps
  int s1_warp;                // the consumer warp that accumulates the weight sums of a group
  int nsw;                    // reduce2_kernel: ring slots
  int stage_bytes;            // reduce2_kernel: bytes of one row-chunk stage (chunk length x 8 + 16)
  int balanced;               // the plan is a balanced one: lane meta bits 12-27 are a slot mask (not a count)
  int debug;                  // timing experiments only (ECCKD_DEBUG): 1 skip the row evaluation, 4 skip only the slot loop
  double* partial;            // [ncta][maxg][nw*RPW][ng][4]
  int32_t* partial_z;         // [ncta][maxg][nw*RPW][ng]
};

struct CombineParams {
  const GroupDesc* groups;
  int ngroups;
  int nw;
  int rpw;         // rows per consumer warp of the kernel that wrote the pieces
  int ncta;
  int maxg;
  int nwin;
  int cw;
  int64_t nchunks;
  int ng;
  int shared_s1;   // weight sums are accumulated once per group (by the last consumer warp)
  int float_row;   // reduce2_kernel's floating row: its second partial sits in row 1 of the last warp
  int s1_warp;     // the consumer warp whose row 0 holds the weight sums
  const double* partial;
  const int32_t* partial_z;
  double* sums;    // [nrows][ng][NSUM]
  double* minmax;  // [nrows][ng][NMINMAX]
};

struct FinalizeParams {
  const OutDesc* outs;  // [nrows]
  int nrows;
  int ng;
  const double* sums;
  const double* minmax;
  const long long* count;  // [ng] wavenumbers per g point (global)
  double kd;
};

}  // namespace ecckd

struct ecckd_gpoint_plan {
  ecckd_ctx* ctx = nullptr;
  int64_t nwav = 0;
  int64_t nchunks = 0;
  int64_t nrounds = 0;
  int ng = 0;
  int ch = 0;                            // chunk length chosen for this g-point map
  bool windowed = false;                 // rounds obey the WIN window constraint (reduce2/3_kernel can run)
  bool balanced = false;                 // bank-conflict-free cell placement (reduce2/3_kernel ONLY: lanes of a g
                                         // point are not adjacent, the lane meta word carries a slot mask)
  uint64_t env_sig = 0;                  // plan_env_signature() at creation (the host calls' plan cache)
  int32_t* d_g_copy = nullptr;           // balanced plans keep the g-point map to build `legacy` on demand
  ecckd_gpoint_plan* legacy = nullptr;   // free-layout plan for reduce_kernel (kinds with per-row weight sums)
  uint32_t* d_round_off = nullptr;       // [nchunks+1]
  ecckd::PlanRound* d_rounds = nullptr;  // [nrounds]
  unsigned long long* d_count = nullptr; // [ng]
  std::vector<int64_t> h_count;
  double* d_scratch = nullptr;           // per-(row, g) partial results of the single-GPU path
  size_t scratch_bytes = 0;
};
