// ref_tools_fgp_glue.cpp - C ABI over the internals of the reference's find_g_points.cpp.
//
// TEST INFRASTRUCTURE ONLY (see oracle/ecckd_oracle.cpp header).  No algorithm is restated here.
// find_g_points.cpp defines fit_optical_depth_{lw,sw,sw_total_trans} and the class
// CkdEquipartition (the `calc_error` of the partition search, :206-427) inside the translation
// unit of the tool's main, with no header; this file therefore INCLUDES that source, where it lies
// under /root/reference/src/ecckd, and wraps
//   * the three fit functions (:54-204),
//   * CkdEquipartition::init_lw / init_sw / init_sw_extras + calc_error (:209-405),
//   * the reference's own partition search Equipartition::equipartition_e / _n
//     (equipartition.cpp:201-805) driven by that calc_error - the oracle of the batched GPU search.
// (CkdModel is wrapped in ref_tools_glue.cpp: ckd_model.h and single_gas_data.h each define a
// different SingleGasData and cannot meet in one translation unit.)
// (its `main` keeps its name here - only a function called main may reach its end without a return -
// and is renamed to ref_main_find_g_points in the object file by oracle/build_ref_tools.sh)
#include "find_g_points.cpp"

#include <cstdint>

#define REF_API extern "C" __attribute__((visibility("default")))

namespace {
Vector vec(const double* p, int64_t n) {
  if (!p) return Vector();
  return Vector(const_cast<double*>(p), dimensions((int)n));
}
Matrix mat(const double* p, int64_t n0, int64_t n1) {
  if (!p) return Matrix();
  return Matrix(const_cast<double*>(p), dimensions((int)n0, (int)n1));
}
const char* FIT_METHODS[] = {"linear", "transmission", "transmission-2", "square-root", "logarithmic",
                             "total-transmission"};
template <typename F>
int guarded(F f) {
  try {
    f();
    return 0;
  } catch (int code) {
    return code;
  } catch (...) {
    return -1;
  }
}
}  // namespace

// method names are passed as strings (the reference compares strings)
REF_API int ref_fit_optical_depth_lw(const char* method, int nlay, int64_t n, int64_t i1, int64_t i2,
                                     const double* planck_hl, const double* metric, double* out) {
  return guarded([&]() {
    Vector r = fit_optical_depth_lw(method, (int)i1, (int)i2, mat(planck_hl, nlay + 1, n), mat(metric, nlay, n));
    for (int k = 0; k < nlay; ++k) out[k] = r(k);
  });
}
REF_API int ref_fit_optical_depth_sw(const char* method, int nlay, int64_t n, int64_t i1, int64_t i2, const double* ssi,
                                     const double* metric, double* out) {
  return guarded([&]() {
    Vector r = fit_optical_depth_sw(method, (int)i1, (int)i2, vec(ssi, n), mat(metric, nlay, n));
    for (int k = 0; k < nlay; ++k) out[k] = r(k);
  });
}
REF_API int ref_fit_optical_depth_sw_total_trans(int nlay, int64_t n, int64_t i1, int64_t i2, const double* ssi,
                                                 const double* bg_od, const double* od, double* out) {
  return guarded([&]() {
    Vector r = fit_optical_depth_sw_total_trans((int)i1, (int)i2, vec(ssi, n), mat(bg_od, nlay, n), mat(od, nlay, n));
    for (int k = 0; k < nlay; ++k) out[k] = r(k);
  });
}

// One band of the partition search, longwave (find_g_points.cpp:1158-1166): all arrays are in rank
// order, [.][n]; the band is the index range [ibegin, iend].
struct RefBandLw {
  const char* method;
  double flux_weight;
  int nlay;
  int64_t n, ibegin, iend;
  const double *layer_weight, *pressure_hl, *surf_emissivity, *surf_planck, *flux_dn_surf, *flux_up_toa;
  const double *planck_hl, *bg_od, *metric, *hr;
};
static void init_lw(CkdEquipartition& eq, const RefBandLw& b) {
  eq.init_lw(b.method, b.flux_weight, vec(b.layer_weight, b.nlay), vec(b.pressure_hl, b.nlay + 1),
             vec(b.surf_emissivity, b.n), vec(b.surf_planck, b.n), vec(b.flux_dn_surf, b.n), vec(b.flux_up_toa, b.n),
             mat(b.planck_hl, b.nlay + 1, b.n), mat(b.bg_od, b.nlay, b.n), mat(b.metric, b.nlay, b.n),
             mat(b.hr, b.nlay, b.n), (int)b.ibegin, (int)b.iend);
  eq.set_verbose(false);
}
REF_API int ref_calc_error_lw(const RefBandLw* b, int nint, const double* bound1, const double* bound2, double* err) {
  return guarded([&]() {
    CkdEquipartition eq;
    init_lw(eq, *b);
    for (int k = 0; k < nint; ++k) err[k] = eq.calc_error(bound1[k], bound2[k]);
  });
}
// equipartition_e (find_g_points.cpp:1230): returns the status; *ng, bounds[ng+1], error[ng] (capacity max_ng)
REF_API int ref_equipartition_e_lw(const RefBandLw* b, double tolerance, int max_iterations, double tolerance_tolerance,
                                   int parallel, int max_ng, int* ng, double* bounds, double* error, int* status) {
  return guarded([&]() {
    CkdEquipartition eq;
    init_lw(eq, *b);
    eq.set_parallel(parallel != 0);
    eq.set_partition_max_iterations(max_iterations);
    eq.set_partition_tolerance(tolerance_tolerance);
    std::vector<ep_real> bnd, e;
    int n = 10;
    const EpStatus st = eq.equipartition_e(tolerance, 0.0, 1.0, n, bnd, e);
    *status = (int)st;
    *ng = n;
    if (n > max_ng) throw 1;
    for (int k = 0; k <= n; ++k) bounds[k] = bnd[k];
    for (int k = 0; k < n; ++k) error[k] = e[k];
  });
}
// equipartition_n from the reference's initial bounds sqrt(i / ng) (find_g_points.cpp:1239-1242)
REF_API int ref_equipartition_n_lw(const RefBandLw* b, int max_iterations, double tolerance_tolerance, int parallel,
                                   int ng, double* bounds, double* error, int* status) {
  return guarded([&]() {
    CkdEquipartition eq;
    init_lw(eq, *b);
    eq.set_parallel(parallel != 0);
    eq.set_partition_max_iterations(max_iterations);
    eq.set_partition_tolerance(tolerance_tolerance);
    *status = (int)eq.equipartition_n(ng, bounds, error);
  });
}

struct RefBandSw {
  const char* method;
  double flux_weight, cos_sza, albedo;
  int nlay;
  int64_t n, ibegin, iend;
  const double *layer_weight, *pressure_hl, *ssi, *flux_dn_surf, *flux_up_toa, *bg_od, *metric, *hr;
  // total-transmission extras (NULL otherwise), find_g_points.cpp:1175-1180
  const double *flux_dn_surf_low, *flux_up_toa_low, *flux_dn_surf_high, *flux_up_toa_high, *hr_low, *hr_high;
  double min_scaling, max_scaling;
};
static void init_sw(CkdEquipartition& eq, const RefBandSw& b) {
  eq.init_sw(b.method, b.flux_weight, vec(b.layer_weight, b.nlay), b.cos_sza, vec(b.pressure_hl, b.nlay + 1),
             vec(b.ssi, b.n), b.albedo, vec(b.flux_dn_surf, b.n), vec(b.flux_up_toa, b.n), mat(b.bg_od, b.nlay, b.n),
             mat(b.metric, b.nlay, b.n), mat(b.hr, b.nlay, b.n), (int)b.ibegin, (int)b.iend);
  if (b.flux_dn_surf_low)
    eq.init_sw_extras(vec(b.flux_dn_surf_low, b.n), vec(b.flux_up_toa_low, b.n), vec(b.flux_dn_surf_high, b.n),
                      vec(b.flux_up_toa_high, b.n), b.min_scaling, b.max_scaling, mat(b.hr_low, b.nlay, b.n),
                      mat(b.hr_high, b.nlay, b.n), (int)b.ibegin, (int)b.iend);
  eq.set_verbose(false);
}
REF_API int ref_calc_error_sw(const RefBandSw* b, int nint, const double* bound1, const double* bound2, double* err) {
  return guarded([&]() {
    CkdEquipartition eq;
    init_sw(eq, *b);
    for (int k = 0; k < nint; ++k) err[k] = eq.calc_error(bound1[k], bound2[k]);
  });
}


// The reference's Equipartition over an arbitrary interval-error function (a callback): the oracle of the
// host driver ecckd_partition_search_* (ecckd_b200/csrc/partition_search.cu) on synthetic error functions.
typedef double (*ref_interval_error_fn)(void* user, double bound1, double bound2);
namespace {
class CallbackEquipartition : public Equipartition {
public:
  CallbackEquipartition(ref_interval_error_fn f, void* u) : fn(f), user(u) {}
  virtual ep_real calc_error(ep_real bound1, ep_real bound2) {
    ++calls;
    return fn(user, bound1, bound2);
  }
  ref_interval_error_fn fn;
  void* user;
  long calls = 0;
};
struct RefEpOptions {  // the setters of equipartition.h:120-175; next_bound_* have no setter (defaults stay)
  double partition_tolerance, resolution;
  int partition_max_iterations, line_search_max_iterations, minimize_frac_range, cubic_interpolation;
};
void configure(Equipartition& eq, const RefEpOptions* o) {
  eq.set_verbose(false);
  eq.set_parallel(false);
  if (!o) return;
  eq.set_partition_tolerance(o->partition_tolerance);
  eq.set_resolution(o->resolution);
  eq.set_partition_max_iterations(o->partition_max_iterations);
  eq.set_line_search_max_iterations(o->line_search_max_iterations);
  eq.set_minimize_frac_range(o->minimize_frac_range != 0);
  eq.set_cubic_interpolation(o->cubic_interpolation != 0);
}
}  // namespace
REF_API int ref_equipartition_n_callback(ref_interval_error_fn fn, void* user, const RefEpOptions* opt, int ni,
                                         double* bounds, double* error, int* status, long* calls) {
  return guarded([&]() {
    CallbackEquipartition eq(fn, user);
    configure(eq, opt);
    *status = (int)eq.equipartition_n(ni, bounds, error);
    if (calls) *calls = eq.calls;
  });
}
REF_API int ref_equipartition_e_callback(ref_interval_error_fn fn, void* user, const RefEpOptions* opt, double target_error,
                                         double bound0, double boundn, int max_ni, int* ni, double* bounds, double* error,
                                         int* status, long* calls) {
  return guarded([&]() {
    CallbackEquipartition eq(fn, user);
    configure(eq, opt);
    std::vector<ep_real> bnd, e;
    int n = 0;
    *status = (int)eq.equipartition_e(target_error, bound0, boundn, n, bnd, e);
    *ni = n;
    if (calls) *calls = eq.calls;
    if ((int)e.size() > max_ni) throw 1;
    for (size_t k = 0; k < bnd.size(); ++k) bounds[k] = bnd[k];
    for (size_t k = 0; k < e.size(); ++k) error[k] = e[k];
  });
}
