// spectral_reduce.cuh - data structures shared by the spectral reduction
// (LBL spectra -> per-g-point averages) host code and kernels.
#pragma once
#include "common.cuh"

namespace ecckd {

constexpr int CH = 224;            // wavenumbers per chunk (1792 B per row: a multiple of 16 B)
constexpr int SLOTS_PER_LANE = 8;  // plan slots per lane per round
constexpr int NS = 3;              // data stages per consumer warp
constexpr int NSW = 3;             // weight stages per CTA
constexpr int NPW = 2;             // producer (weight) warps per CTA
constexpr int NW_MAX = 19;         // consumer warps (= rows) per CTA: 23 warps x 88 registers fill the register file
// lane meta word of the plan: bits 0-11 g point (0xFFF = lane unused), bits 12-15 valid slots (1..8)
constexpr unsigned NO_G = 0xFFFu;
constexpr int MAX_NG = 4094;
constexpr int WT_STRIDE = CH + 2;  // weight tile: CH weights + a zero entry that padding slots point at

// One round of the plan: for each lane, 8 byte offsets into the row stage (padding repeats
// slot 0), 8 byte offsets into the weight tile (padding -> the zero entry), the lane meta word
// and the number of shuffle steps the segmented reduction needs.  1104 bytes, copied to shared
// memory with one bulk copy.
struct PlanRound {
  uint4 tau_off[32];
  uint4 w_off[32];
  uint16_t meta[32];
  uint32_t steps;
  uint32_t pad_[3];
};
static_assert(sizeof(PlanRound) == 1104, "PlanRound layout");

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

// Up to NW_MAX rows that share one weight row: one consumer warp each.
struct GroupDesc {
  const double* od[NW_MAX];
  int32_t row[NW_MAX];     // index in the partial-sum arrays
  uint8_t is_log[NW_MAX];  // row uses logarithmic averaging (hybrid method)
  uint8_t bulk_ok[NW_MAX]; // row base is 16-byte aligned: bulk async copies allowed
  int32_t nrows;
  int32_t wrow;
};

struct ReduceParams {
  const GroupDesc* groups;
  int ngroups;
  int nw;                  // consumer warps per CTA
  int ncta;                // grid size P
  int maxg;                // partial slots per CTA (groups a CTA can touch)
  int64_t nchunks;
  int64_t nwav;
  int ng;
  const uint32_t* round_off;  // [nchunks+1]
  const PlanRound* rounds;    // [nrounds]
  const double* weight;       // MATRIX: [wrow][nwav] (stride w_stride); VECTOR: [nwav]
  int64_t w_stride;
  int weight_bulk_ok;
  const double* temperature;  // PLANCK: device [n weight rows]
  const double* wn;
  const double* dwn;
  double kd;                  // transmission exponent scale (LW_DIFFUSIVITY*OD_SCALING*k)
  double* partial;            // [ncta][maxg][nw][ng][4]
  int32_t* partial_z;         // [ncta][maxg][nw][ng]
};

struct CombineParams {
  const GroupDesc* groups;
  int ngroups;
  int nw;
  int ncta;
  int maxg;
  int64_t nchunks;
  int ng;
  int shared_s1;   // weight sums are accumulated by the first row of each group only
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
  uint32_t* d_round_off = nullptr;       // [nchunks+1]
  ecckd::PlanRound* d_rounds = nullptr;  // [nrounds]
  unsigned long long* d_count = nullptr; // [ng]
  std::vector<int64_t> h_count;
  double* d_scratch = nullptr;           // per-(row, g) partial results of the single-GPU path
  size_t scratch_bytes = 0;
};
