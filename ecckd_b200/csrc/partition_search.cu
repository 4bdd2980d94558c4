// Partition search of find_g_points: the control flow of the reference's Equipartition
// class (src/ecckd/equipartition.cpp:160-805, equipartition.h:64-209) as a host driver over a BATCHED interval
// error evaluator - one call of ecckd_interval_error_batch_*_dev per calc_error_all (equipartition.h:97-113)
// instead of one calc_error per OpenMP thread, and the candidate bounds of a line search evaluated
// speculatively, several per batch (the sequence of candidates does not depend on the errors).
//
// The iterates are defined by the reference's update formulas, which are restated here one to one (the
// search is discrete and non-smooth: another formula gives another partition); given an evaluator that
// returns the same errors, bounds, errors, number of intervals and status are bitwise those of the
// reference (tests/test_partition_search.py drives both with the same function).  Host code only.
#include <algorithm>
#include <cmath>
#include <limits>
#include <vector>

#include "common.cuh"

using namespace ecckd;

struct ecckd_partition_search {
  ecckd_partition_options opt;
  ecckd_interval_error_fn fn = nullptr;
  void* user = nullptr;
  bool errors_valid = false;  // the caller's error[] belongs to the caller's bounds[] (errors_up_to_date)
  int rc = ECCKD_OK;          // first evaluator failure
  int64_t batches = 0, intervals = 0, speculative_unused = 0;

  // ---- evaluation
  bool eval(int n, const double* b1, const double* b2, double* err) {
    if (rc != ECCKD_OK) return false;
    rc = fn(user, n, b1, b2, err);
    ++batches;
    intervals += n;
    return rc == ECCKD_OK;
  }
  bool eval_all(int ni, const double* bounds, double* err) { return eval(ni, bounds, bounds + 1, err); }
  bool eval_one(double b1, double b2, double* err) { return eval(1, &b1, &b2, err); }

  // (max - min) / mean, or the fractional standard deviation (equipartition.cpp:132-158)
  double cost(int ni, const double* e) const {
    double mean = 0.0, lo = std::numeric_limits<double>::infinity(), hi = -std::numeric_limits<double>::infinity();
    for (int i = 0; i < ni; ++i) {
      mean += e[i];
      if (e[i] < lo) lo = e[i];
      if (e[i] > hi) hi = e[i];
    }
    mean /= ni;
    if (opt.minimize_frac_range) return (hi - lo) / mean;
    double chi2 = 0.0;
    for (int i = 0; i < ni; ++i) chi2 += (e[i] - mean) * (e[i] - mean);
    return std::sqrt(chi2 / ni) / mean;
  }

  int line_search(int ni, double* bounds, double* newbounds, double* error);
  int pair(double* bounds, double* error);
  int search_n(int ni, double* bounds_out, double* error);
  int search_e(double target, double bound0, double boundn, std::vector<double>& bounds, std::vector<double>& error);
  double bound_below(double target, double bound0, double bound2, double test, double* error_test);
  double bound_above(double target, double bound1, double boundn, double test, double* error_test);
};

namespace {
inline void halve_towards(int ni, double* a, const double* b) {  // a <- 0.5 a + 0.5 b, element by element
  for (int i = 0; i <= ni; ++i) a[i] = 0.5 * a[i] + 0.5 * b[i];
}
}  // namespace

// equipartition.cpp:160-196.  Candidate k is the proposal moved k times halfway back to the current bounds;
// the candidates of one batch are formed before any of their errors is known.
int ecckd_partition_search::line_search(int ni, double* bounds, double* newbounds, double* error) {
  if (!errors_valid) {
    if (!eval_all(ni, bounds, error)) return ECCKD_EP_FAILURE;
    errors_valid = true;
  }
  const double start = cost(ni, error);
  int remaining = opt.line_search_max_iterations;
  const int spec = std::max(1, opt.speculative_line_search);
  std::vector<double> cand, b1, b2, err;
  halve_towards(ni, newbounds, bounds);
  while (remaining > 0) {
    const int k = std::min(spec, remaining);
    cand.assign((size_t)k * (ni + 1), 0.0);
    for (int c = 0; c < k; ++c) {
      double* dst = &cand[(size_t)c * (ni + 1)];
      std::copy(newbounds, newbounds + ni + 1, dst);
      if (c + 1 < k) halve_towards(ni, newbounds, bounds);
    }
    b1.resize((size_t)k * ni);
    b2.resize((size_t)k * ni);
    err.resize((size_t)k * ni);
    for (int c = 0; c < k; ++c)
      for (int i = 0; i < ni; ++i) {
        b1[(size_t)c * ni + i] = cand[(size_t)c * (ni + 1) + i];
        b2[(size_t)c * ni + i] = cand[(size_t)c * (ni + 1) + i + 1];
      }
    if (!eval(k * ni, b1.data(), b2.data(), err.data())) return ECCKD_EP_FAILURE;
    errors_valid = false;
    for (int c = 0; c < k; ++c) {
      const double* e = &err[(size_t)c * ni];
      std::copy(e, e + ni, error);
      if (cost(ni, e) < start) {
        std::copy(&cand[(size_t)c * (ni + 1)], &cand[(size_t)c * (ni + 1)] + ni + 1, bounds);
        errors_valid = true;
        speculative_unused += (int64_t)(k - 1 - c) * ni;
        return ECCKD_EP_SUCCESS;
      }
    }
    remaining -= k;
    halve_towards(ni, newbounds, bounds);
  }
  return ECCKD_EP_NO_PROGRESS;
}

// Two intervals, one interior bound: bracket it from both sides, then regula falsi with a bisection
// fallback (equipartition.cpp:200-345).
int ecckd_partition_search::pair(double* bounds, double* error) {
  if (!errors_valid) {
    if (!eval_all(2, bounds, error)) return ECCKD_EP_FAILURE;
    errors_valid = true;
  }
  double left = bounds[0], right = bounds[2];
  double ediff_left = 0.0, ediff_right = 0.0;
  double frac = 0.5 * std::fabs(error[1] - error[0]) / (error[0] + error[1]);
  const double tol = opt.partition_tolerance;
  const double frac_orig = frac;
  double nb[3] = {bounds[0], bounds[1], bounds[2]};
  double ne[2] = {error[0], error[1]};
  int remaining = opt.partition_max_iterations;
  if (error[0] > error[1]) {  // the bound is too far right: that is a right limit; look for a left one
    right = bounds[1];
    ediff_right = error[1] - error[0];
    while (remaining) {
      nb[1] = (-ediff_right * nb[0] + (ne[0] + ediff_right) * nb[1]) / ne[0];
      if (!eval_all(2, nb, ne)) return ECCKD_EP_FAILURE;
      if (ne[0] < ne[1]) {
        left = nb[1];
        ediff_left = ne[1] - ne[0];
        break;
      }
      ediff_right = ne[1] - ne[0];
      --remaining;
    }
  } else {
    left = bounds[1];
    ediff_left = error[1] - error[0];
    while (remaining) {
      nb[1] = (ediff_left * nb[2] + (ne[1] - ediff_left) * nb[1]) / ne[1];
      if (!eval_all(2, nb, ne)) return ECCKD_EP_FAILURE;
      if (ne[0] > ne[1]) {
        right = nb[1];
        ediff_right = ne[1] - ne[0];
        break;
      }
      ediff_left = ne[1] - ne[0];
      --remaining;
    }
  }
  bool stalled = false;
  double prev_frac = frac;
  while (remaining) {
    nb[1] = stalled ? 0.5 * (right + left) : (ediff_left * right - ediff_right * left) / (ediff_left - ediff_right);
    if (!eval_all(2, nb, ne)) return ECCKD_EP_FAILURE;
    const double ediff = ne[1] - ne[0];
    frac = 0.5 * std::fabs(ediff) / (ne[0] + ne[1]);
    if (frac < tol && frac < frac_orig) {
      bounds[1] = nb[1];
      error[0] = ne[0];
      error[1] = ne[1];
      errors_valid = true;
      return ECCKD_EP_SUCCESS;
    } else if (frac == prev_frac) {
      if (stalled) break;
      stalled = true;
    }
    if (ediff < 0) {
      ediff_right = ediff;
      right = nb[1];
    } else {
      ediff_left = ediff;
      left = nb[1];
    }
    prev_frac = frac;
    --remaining;
  }
  if (frac < frac_orig) {  // not converged, but better than it was: keep it
    bounds[1] = nb[1];
    error[0] = ne[0];
    error[1] = ne[1];
    errors_valid = true;
    if (right - left < opt.resolution) return ECCKD_EP_RESOLUTION_LIMIT_REACHED;
    if (!remaining) return ECCKD_EP_MAX_ITERATIONS_REACHED;
    return ECCKD_EP_SUCCESS;
  }
  return ECCKD_EP_NO_PROGRESS;
}

// equipartition.cpp:353-562: move every interior bound to where the cumulative error reaches its share, accept
// the move (or a fraction of it) if the spread of the errors shrinks; when no fraction does, sweep through the
// neighbouring pairs.
int ecckd_partition_search::search_n(int ni, double* bounds_out, double* error) {
  if (ni == 2) return pair(bounds_out, error);
  int status = ECCKD_EP_SUCCESS;
  int shuffles = opt.partition_max_iterations / 2;
  for (int i = 0; i < ni; ++i)
    if (bounds_out[i + 1] <= bounds_out[i]) return ECCKD_EP_INPUT_ERROR;
  std::vector<double> bounds(bounds_out, bounds_out + ni + 1), cum(ni + 1), nb(ni + 1);
  int remaining = opt.partition_max_iterations;
  while (remaining > 0) {
    if (!errors_valid) {
      if (!eval_all(ni, bounds.data(), error)) return ECCKD_EP_FAILURE;
      errors_valid = true;
    }
    if (cost(ni, error) < opt.partition_tolerance) break;
    cum[0] = 0.0;
    for (int i = 0; i < ni; ++i) cum[i + 1] = cum[i] + error[i];
    const double share = cum[ni] / ni;
    nb[0] = bounds[0];
    nb[ni] = bounds[ni];
    int io = 0;
    for (int in = 1; in < ni; ++in) {
      const double target = share * in;
      while (cum[io + 1] < target) ++io;
      if (opt.cubic_interpolation) {  // Hermite cubic through the interval with centred end gradients
        const double u = (target - cum[io]) / (cum[io + 1] - cum[io]);
        const double u2 = u * u, u3 = u2 * u;
        const double grad = (bounds[io + 1] - bounds[io]) / (cum[io + 1] - cum[io]);
        const double g0 = io == 0 ? grad : (bounds[io + 1] - bounds[io - 1]) / (cum[io + 1] - cum[io - 1]);
        const double g1 = io == ni - 1 ? grad : (bounds[io + 2] - bounds[io]) / (cum[io + 2] - cum[io]);
        nb[in] = (2.0 * u3 - 3.0 * u2 + 1) * bounds[io] + (u3 - 2.0 * u2 + u) * g0 + (-2.0 * u3 + 3.0 * u2) * bounds[io + 1] +
                 (u3 - u2) * g1;
      } else {
        nb[in] = ((cum[io + 1] - target) * bounds[io] + (target - cum[io]) * bounds[io + 1]) / (cum[io + 1] - cum[io]);
      }
    }
    if (opt.resolution > 0.0) {
      bool moved = false;
      for (int i = 1; i < ni && !moved; ++i) moved = std::fabs(nb[i] - bounds[i]) > opt.resolution;
      if (!moved) {
        std::copy(bounds.begin(), bounds.end(), bounds_out);
        return ECCKD_EP_RESOLUTION_LIMIT_REACHED;
      }
    }
    const int ls = line_search(ni, bounds.data(), nb.data(), error);
    if (ls == ECCKD_EP_FAILURE) return ls;
    if (ls != ECCKD_EP_SUCCESS) {
      status = ECCKD_EP_FAILED_TO_CONVERGE;
      int stuck = 0;
      if (ni > 2 && shuffles > 0) {
        auto sweep = [&](int from, int to, int step) {
          for (int i = from; step > 0 ? i <= to : i >= to; i += step) {
            const int s = pair(&bounds[i], &error[i]);
            if (s == ECCKD_EP_NO_PROGRESS) ++stuck;
          }
        };
        if (shuffles % 2) {  // low -> high -> low
          sweep(0, ni - 2, +1);
          sweep(ni - 3, 0, -1);
        } else {  // high -> low -> high
          sweep(ni - 2, 0, -1);
          sweep(1, ni - 2, +1);
        }
        if (rc != ECCKD_OK) return ECCKD_EP_FAILURE;
        --shuffles;
        if (cost(ni, error) < opt.partition_tolerance) {
          status = ECCKD_EP_SUCCESS;
          break;
        } else if (stuck >= ni * 2 - 3) {
          status = ECCKD_EP_FAILED_TO_CONVERGE;
        } else {
          status = ECCKD_EP_SUCCESS;  // the spread shrank a little: try the general step again
        }
      }
      if (status != ECCKD_EP_SUCCESS) break;
    }
    --remaining;
  }
  std::copy(bounds.begin(), bounds.end(), bounds_out);
  if (remaining == 0) status = ECCKD_EP_MAX_ITERATIONS_REACHED;
  errors_valid = false;
  return status;
}

// equipartition.cpp:626-712: the largest bound1 < bound2 whose interval [bound1, bound2] has an error within
// (1 - tolerance, 1] of the target; secant / damped steps between a too-large and a too-small trial.
double ecckd_partition_search::bound_below(double target, double bound0, double bound2, double test, double* error_test_io) {
  const double emax = target, emin = target * (1.0 - opt.next_bound_error_tolerance);
  double low = bound0, high = bound2, e_low = -1.0, e_high = 0.0, e_test;
  int remaining = opt.next_bound_max_iterations;
  if (*error_test_io < 0.0) {
    if (!eval_one(test, bound2, &e_test)) return test;
  } else {
    e_test = *error_test_io;
  }
  while (remaining > 0 && (e_test > emax || e_test < emin)) {
    if (e_test > target) {
      low = test;
      e_low = e_test;
    } else {
      high = test;
      e_high = e_test;
    }
    if (low == high) break;
    if (e_low > 0.0) {
      test = ((target - e_high) * low + (e_low - target) * high) / (e_low - e_high);
      if (e_high == 0.0) {
        test = 0.5 * (test + high);
      } else if (e_test < emin && e_low > 2.0 * emax) {
        test = 0.75 * test + 0.25 * low;
      }
    } else {
      test = std::max(low, high - 0.5 * target * (bound2 - high) / e_high);
    }
    if (!eval_one(test, bound2, &e_test)) return test;
    --remaining;
  }
  *error_test_io = e_test;
  return test;
}

// equipartition.cpp:714-805, the mirror image from the lower end.
double ecckd_partition_search::bound_above(double target, double bound1, double boundn, double test, double* error_test_io) {
  const double emax = target, emin = target * (1.0 - opt.next_bound_error_tolerance);
  double low = bound1, high = boundn, e_low = 0.0, e_high = -1.0, e_test;
  int remaining = opt.next_bound_max_iterations;
  if (*error_test_io < 0.0) {
    if (!eval_one(bound1, test, &e_test)) return test;
  } else {
    e_test = *error_test_io;
  }
  while (remaining > 0 && (e_test > emax || e_test < emin)) {
    if (e_test > target) {
      high = test;
      e_high = e_test;
    } else {
      low = test;
      e_low = e_test;
    }
    if (low == high) break;
    if (e_high > 0.0) {
      test = ((target - e_low) * high + (e_high - target) * low) / (e_high - e_low);
      if (e_low == 0.0) {
        test = 0.5 * (test + low);
      } else if (e_test < emin && e_low > 2.0 * emax) {
        test = 0.75 * test + 0.25 * high;
      }
    } else {
      test = std::max(high, high - 0.5 * target * (low - bound1) / e_low);
    }
    if (!eval_one(bound1, test, &e_test)) return test;
    --remaining;
  }
  *error_test_io = e_test;
  return test;
}

// equipartition.cpp:569-623: how many intervals of at most `target` error the range needs - the top one
// first, then upwards from the bottom - followed by the equal-error repartition of that many.
int ecckd_partition_search::search_e(double target, double bound0, double boundn, std::vector<double>& bounds,
                                     std::vector<double>& error) {
  if (boundn <= bound0) return ECCKD_EP_INPUT_ERROR;
  double upper_error = -1.0;
  const double upper = bound_below(target, bound0, boundn, 0.05 * bound0 + 0.95 * boundn, &upper_error);
  if (rc != ECCKD_OK) return ECCKD_EP_FAILURE;
  if (upper == bound0) {
    bounds.assign({bound0, boundn});
    error.assign(1, upper_error);
    return ECCKD_EP_SUCCESS;
  }
  bounds.assign(1, bound0);
  error.clear();
  size_t i = 0;
  while (bounds[i] < upper) {
    double e = -1.0;
    const double b = bound_above(target, bounds[i], upper, 0.25 * bounds[i] + 0.75 * upper, &e);
    if (rc != ECCKD_OK) return ECCKD_EP_FAILURE;
    error.push_back(e);
    bounds.push_back(b);
    ++i;
  }
  error.push_back(upper_error);
  bounds.push_back(boundn);
  errors_valid = true;
  return search_n((int)error.size(), bounds.data(), error.data());
}

// ------------------------------------------------------------------ C ABI
ECCKD_API void ecckd_partition_options_default(ecckd_partition_options* o) {
  if (!o) return;
  o->partition_tolerance = 0.05;         // equipartition.h:187-192
  o->next_bound_error_tolerance = 0.05;
  o->resolution = 0.0;
  o->partition_max_iterations = 20;
  o->line_search_max_iterations = 10;
  o->next_bound_max_iterations = 20;
  o->minimize_frac_range = 1;
  o->cubic_interpolation = 0;
  o->speculative_line_search = 4;
}

ECCKD_API int ecckd_partition_search_create(const ecckd_partition_options* options, ecckd_interval_error_fn fn, void* user,
                                            ecckd_partition_search** out) {
  ECCKD_REQUIRE(fn && out, ECCKD_PARAMETER_ERROR, "ecckd_partition_search_create: NULL argument");
  ecckd_partition_search* ps = new ecckd_partition_search;
  if (options) ps->opt = *options; else ecckd_partition_options_default(&ps->opt);
  ECCKD_REQUIRE(ps->opt.partition_max_iterations >= 0 && ps->opt.line_search_max_iterations >= 0 &&
                    ps->opt.next_bound_max_iterations >= 0,
                ECCKD_PARAMETER_ERROR, "ecckd_partition_search_create: negative iteration limit");
  ps->fn = fn;
  ps->user = user;
  *out = ps;
  return ECCKD_OK;
}

ECCKD_API int ecckd_partition_search_destroy(ecckd_partition_search* ps) {
  delete ps;
  return ECCKD_OK;
}

ECCKD_API int ecckd_partition_search_n(ecckd_partition_search* ps, int ni, double* bounds, double* error,
                                       int errors_given, int* ep_status) {
  ECCKD_REQUIRE(ps && bounds && error && ep_status && ni >= 1, ECCKD_PARAMETER_ERROR,
                "ecckd_partition_search_n: bad argument");
  ps->rc = ECCKD_OK;
  ps->errors_valid = errors_given != 0;
  if (ni == 1) {  // nothing to move
    if (!ps->errors_valid && !ps->eval_all(1, bounds, error)) return ps->rc;
    *ep_status = ECCKD_EP_SUCCESS;
    return ECCKD_OK;
  }
  *ep_status = ps->search_n(ni, bounds, error);
  return ps->rc;
}

ECCKD_API int ecckd_partition_search_e(ecckd_partition_search* ps, double target_error, double bound0, double boundn,
                                       int capacity, int* ni, double* bounds, double* error, int* ep_status) {
  ECCKD_REQUIRE(ps && ni && bounds && error && ep_status && capacity >= 1, ECCKD_PARAMETER_ERROR,
                "ecckd_partition_search_e: bad argument");
  ps->rc = ECCKD_OK;
  ps->errors_valid = false;
  std::vector<double> b, e;
  *ep_status = ps->search_e(target_error, bound0, boundn, b, e);
  if (ps->rc != ECCKD_OK) return ps->rc;
  *ni = (int)e.size();
  if (*ep_status == ECCKD_EP_INPUT_ERROR) {
    *ni = 0;
    return ECCKD_OK;
  }
  ECCKD_REQUIRE((int)e.size() <= capacity, ECCKD_PARAMETER_ERROR,
                "ecckd_partition_search_e: %d intervals needed, room for %d", (int)e.size(), capacity);
  std::copy(b.begin(), b.end(), bounds);
  std::copy(e.begin(), e.end(), error);
  return ECCKD_OK;
}

// find_g_points.cpp:1229-1256 (one band without sub-bands) and :1396-1400: equipartition_e on [0, 1]; when the
// number of intervals falls outside [min_g_points, max_g_points] the band is re-partitioned into that many from
// the bounds sqrt(i / ng); rank1 / rank2 of every g point are the snapped indices of its bounds (band-relative).
ECCKD_API int ecckd_partition_search_band(ecckd_partition_search* ps, int64_t npoints, double target_error,
                                          int min_g_points, int max_g_points, int capacity, int* ng, double* bounds,
                                          double* error, int64_t* rank1, int64_t* rank2, int* ep_status) {
  ECCKD_REQUIRE(ps && ng && bounds && error && ep_status && capacity >= 1 && npoints >= 1 && min_g_points >= 1 &&
                    max_g_points >= min_g_points && max_g_points <= capacity,
                ECCKD_PARAMETER_ERROR, "ecckd_partition_search_band: bad argument");
  ECCKD_TRY(ecckd_partition_search_e(ps, target_error, 0.0, 1.0, capacity, ng, bounds, error, ep_status));
  const int want = *ng < min_g_points ? min_g_points : (*ng > max_g_points ? max_g_points : *ng);
  if (want != *ng) {
    *ng = want;
    for (int i = 0; i <= want; ++i) bounds[i] = std::sqrt(static_cast<double>(i) / static_cast<double>(want));
    ECCKD_TRY(ecckd_partition_search_n(ps, want, bounds, error, 0, ep_status));
  }
  for (int i = 0; i < *ng; ++i)  // find_g_points.cpp:1390-1393
    ECCKD_REQUIRE(bounds[i + 1] - bounds[i] > 0.0, ECCKD_PARAMETER_ERROR, "Bounds are not monotonically increasing");
  if (rank1 && rank2)
    for (int i = 0; i < *ng; ++i) {  // lower_index / upper_index, find_g_points.cpp:282-287
      rank1[i] = static_cast<int64_t>(std::ceil(bounds[i] * (npoints - 1)));
      rank2[i] = static_cast<int64_t>(std::floor(bounds[i + 1] * (npoints - 1)));
    }
  return ECCKD_OK;
}

ECCKD_API int ecckd_partition_search_stats(ecckd_partition_search* ps, int64_t* batches, int64_t* intervals,
                                           int64_t* speculative_unused) {
  ECCKD_REQUIRE(ps, ECCKD_PARAMETER_ERROR, "ecckd_partition_search_stats: NULL argument");
  if (batches) *batches = ps->batches;
  if (intervals) *intervals = ps->intervals;
  if (speculative_unused) *speculative_unused = ps->speculative_unused;
  return ECCKD_OK;
}

// ---- the evaluators of find_g_points: CkdEquipartition::calc_error (find_g_points.cpp:291-405) over a batch
static int eval_bounds(ecckd_ctx* ctx, int64_t n, int nint, const double* b1, const double* b2, std::vector<int64_t>& i1,
                       std::vector<int64_t>& i2, double** d_err) {
  i1.resize(nint);
  i2.resize(nint);
  ECCKD_TRY(ecckd_interval_bounds_to_indices(n, nint, b1, b2, i1.data(), i2.data()));
  void* p = nullptr;
  ECCKD_TRY(ws_reserve(ctx, 6, (size_t)nint * sizeof(double), &p));
  *d_err = static_cast<double*>(p);
  return ECCKD_OK;
}

ECCKD_API int ecckd_interval_error_lw_eval(void* user, int nint, const double* bound1, const double* bound2, double* error) {
  const ecckd_interval_error_lw_args* a = static_cast<const ecckd_interval_error_lw_args*>(user);
  ECCKD_REQUIRE(a && a->ctx && nint >= 1 && bound1 && bound2 && error, ECCKD_PARAMETER_ERROR,
                "ecckd_interval_error_lw_eval: bad argument");
  ECCKD_DEVICE(a->ctx);
  std::vector<int64_t> i1, i2;
  double* d_err = nullptr;
  ECCKD_TRY(eval_bounds(a->ctx, a->n, nint, bound1, bound2, i1, i2, &d_err));
  ECCKD_TRY(ecckd_interval_error_batch_lw_dev(a->ctx, a->averaging_method, a->nlay, a->n, a->stride, a->d_pressure_hl,
                                              a->d_planck_hl, a->d_surf_emissivity, a->d_surf_planck, a->d_bg_optical_depth,
                                              a->d_metric, a->d_flux_dn_surf, a->d_flux_up_toa, a->d_hr, a->flux_weight,
                                              a->d_layer_weight, nint, i1.data(), i2.data(), d_err, nullptr));
  return ecckd_memcpy_d2h(a->ctx, error, d_err, (size_t)nint * sizeof(double));
}

ECCKD_API int ecckd_interval_error_sw_eval(void* user, int nint, const double* bound1, const double* bound2, double* error) {
  const ecckd_interval_error_sw_args* a = static_cast<const ecckd_interval_error_sw_args*>(user);
  ECCKD_REQUIRE(a && a->ctx && nint >= 1 && bound1 && bound2 && error, ECCKD_PARAMETER_ERROR,
                "ecckd_interval_error_sw_eval: bad argument");
  ECCKD_DEVICE(a->ctx);
  std::vector<int64_t> i1, i2;
  double* d_err = nullptr;
  ECCKD_TRY(eval_bounds(a->ctx, a->n, nint, bound1, bound2, i1, i2, &d_err));
  ECCKD_TRY(ecckd_interval_error_batch_sw_dev(
      a->ctx, a->averaging_method, a->nlay, a->n, a->stride, a->cos_sza, a->d_pressure_hl, a->d_ssi, a->albedo,
      a->d_bg_optical_depth, a->d_metric, a->d_flux_dn_surf, a->d_flux_up_toa, a->d_hr, a->flux_weight, a->d_layer_weight,
      a->d_flux_dn_surf_low, a->d_flux_up_toa_low, a->d_hr_low, a->d_flux_dn_surf_high, a->d_flux_up_toa_high, a->d_hr_high,
      a->min_scaling, a->max_scaling, nint, i1.data(), i2.data(), d_err, nullptr));
  return ecckd_memcpy_d2h(a->ctx, error, d_err, (size_t)nint * sizeof(double));
}
