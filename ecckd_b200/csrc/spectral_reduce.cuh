// spectral_reduce.cuh - data structures shared by the spectral reduction
// (LBL spectra -> per-g-point averages) host code and kernels.
#pragma once
#include "common.cuh"

namespace ecckd {

constexpr int CHUNK = 64;          // wavenumbers staged per pipeline stage
constexpr int PITCH = CHUNK + 2;   // doubles per staged row (16 B pad: at most 2-way bank conflicts)
constexpr int NSTAGE = 2;
constexpr unsigned SLOT_END = 0xFFFFu;
constexpr int MAX_NG = 1023;       // slot = (g << 6) | local index, 0xFFFF reserved

// internal per-row method ids (public enum resolved per layer for hybrid)
enum { M_LINEAR = 0, M_TRANS = 1, M_SQRT = 2, M_LOG = 3, M_SUMW = 4 };
// kernel "f kind": which transform the streaming kernel applies
enum { F_LINEAR = 0, F_TRANS = 1, F_SQRT = 2, F_LOG = 3, F_HYBRID = 4, F_NONE = 5 };
enum { W_TILE_MATRIX = 0, W_VECTOR = 1, W_TILE_PLANCK = 2 };

// One (spectrum, layer) row of the streaming reduction.
struct RowDesc {
  const double* od;   // row base (nwav doubles); NULL = unused slot
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

// 32 rows handled by one warp (row = lane).
struct GroupDesc {
  const double* od[32];  // NULL = inactive lane
  int32_t wrows[32];     // distinct weight rows used by this group
  int32_t n_wrows;
  int32_t pad_[3];
  uint8_t wslot[32];     // lane -> index into wrows
  uint8_t is_log[32];    // lane uses logarithmic averaging (hybrid)
};

struct ReduceParams {
  const GroupDesc* groups;
  int ngroups;
  int nspans;
  int ntasks;
  const uint16_t* plan;  // [nchunks][CHUNK]
  int64_t nwav;
  int64_t nchunks;
  int ng;
  const double* weight;  // MATRIX: [wrow][nwav] (stride w_stride); VECTOR: [nwav]
  int64_t w_stride;
  const double* temperature;  // PLANCK: device [n weight rows]
  const double* wn;
  const double* dwn;
  double kd;        // transmission exponent scale (LW_DIFFUSIVITY*OD_SCALING*k)
  double* partial;  // [ntasks][ng][4][32]
  int32_t* zcount;  // [ngroups*32][ng] count of tau <= 0 (log modes)
  int maxw;         // weight rows staged per warp stage
  int smem_per_warp;
};

struct CombineParams {
  const RowDesc* rows;  // indexed by row slot = group*32 + lane; od == NULL for inactive
  int nslots;
  int nspans;
  int ng;
  const double* partial;
  const int32_t* zcount;
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
  int ng = 0;
  uint16_t* d_slots = nullptr;           // [nchunks][CHUNK]
  unsigned long long* d_count = nullptr; // [ng]
  std::vector<int64_t> h_count;
  double* d_scratch = nullptr;           // per-(row, g) partial results of the single-GPU path
  size_t scratch_bytes = 0;
};
