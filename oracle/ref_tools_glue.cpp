// ref_tools_glue.cpp - C ABI over the reference's own tool mains.
//
// TEST INFRASTRUCTURE ONLY (see oracle/ecckd_oracle.cpp header).  No algorithm is restated here:
// oracle/build_ref_tools.sh compiles run_ckd.cpp, create_look_up_table.cpp, reorder_spectrum.cpp and
// find_g_points.cpp where they lie under /root/reference/src/ecckd with `main` renamed, and this
// file calls them with an argv assembled by the test harness, turning the reference's
// `throw <ESA exit code>` (Logging.h:115-117) into a return value.  It also wraps
// CkdModel<false>::calc_optical_depth / calc_planck_function / calc_rayleigh_optical_depth
// (ckd_model.cpp:923-1145, ckd_model.h:242-252) on a model read from a mock ckd-definition file.
#include <cstdint>
#include <exception>
#include <string>

#include <adept_arrays.h>

#include "ckd_model.h"

using namespace adept;

#define REF_API extern "C" __attribute__((visibility("default")))

// (the tools' `main` symbols, renamed in their object files: unmangled names)
extern "C" {
int ref_main_run_ckd(int argc, const char* argv[]);
int ref_main_create_look_up_table(int argc, const char* argv[]);
int ref_main_reorder_spectrum(int argc, const char* argv[]);
int ref_main_find_g_points(int argc, const char* argv[]);
}

static std::string g_what;

template <typename F>
static int guarded(F f, int argc, const char** argv) {
  g_what.clear();
  try {
    return f(argc, argv);
  } catch (int code) {
    return code;
  } catch (const std::exception& e) {
    g_what = e.what();
    return -1;
  } catch (...) {
    g_what = "unknown exception";
    return -2;
  }
}

REF_API int ref_tool_run(const char* tool, int argc, const char** argv) {
  const std::string t(tool);
  if (t == "run_ckd") return guarded(ref_main_run_ckd, argc, argv);
  if (t == "create_look_up_table") return guarded(ref_main_create_look_up_table, argc, argv);
  if (t == "reorder_spectrum") return guarded(ref_main_reorder_spectrum, argc, argv);
  if (t == "find_g_points") return guarded(ref_main_find_g_points, argc, argv);
  g_what = "no such tool";
  return -3;
}
REF_API const char* ref_tool_last_exception() { return g_what.c_str(); }
REF_API const char* ref_tools_build_info() {
  return "reference tools compiled in place: run_ckd.cpp create_look_up_table.cpp reorder_spectrum.cpp "
         "find_g_points.cpp (+ ckd_model.cpp read_spectrum.cpp read_merged_spectrum.cpp write_order.cpp "
         "equipartition.cpp, the hot-path units, Logging.cpp Error.cpp file_manager.cpp DataFileEngineCfg.cpp "
         "readconfig.c) against oracle/adept_shim and oracle/ref_mock";
}

namespace {
Vector vec(const double* p, int64_t n) {
  if (!p) return Vector();
  return Vector(const_cast<double*>(p), dimensions((int)n));
}
Matrix mat(const double* p, int64_t n0, int64_t n1) {
  if (!p) return Matrix();
  return Matrix(const_cast<double*>(p), dimensions((int)n0, (int)n1));
}
template <typename F>
int guarded_call(F f) {
  try {
    f();
    return 0;
  } catch (int code) {
    return code;
  } catch (const std::exception& e) {
    g_what = e.what();
    return -1;
  } catch (...) {
    return -2;
  }
}
}  // namespace

// ---------------------------------------------------------------- CkdModel on a mock ckd-definition file
REF_API int ref_ckd_calc_optical_depth(const char* file, int igas, int ncol, int nlay, const double* pressure_hl,
                                       const double* temperature_fl, const double* vmr_fl, int* ng_out, double* od,
                                       int64_t capacity) {
  return guarded_call([&]() {
    CkdModel<false> model(file);
    Array3 r = model.calc_optical_depth(igas, mat(pressure_hl, ncol, nlay + 1), mat(temperature_fl, ncol, nlay),
                                        mat(vmr_fl, ncol, nlay));
    *ng_out = r.size(2);
    if ((int64_t)r.size() > capacity) throw 1;
    for (long k = 0; k < (long)r.size(); ++k) od[k] = r.flat(k);
  });
}
REF_API int ref_ckd_calc_planck_function(const char* file, int64_t n, const double* temperature, double* out,
                                         int64_t capacity) {
  return guarded_call([&]() {
    CkdModel<false> model(file);
    Matrix r = model.calc_planck_function(vec(temperature, n));
    if ((int64_t)r.size() > capacity) throw 1;
    for (long k = 0; k < (long)r.size(); ++k) out[k] = r.flat(k);
  });
}
REF_API int ref_ckd_calc_rayleigh_optical_depth(const char* file, int ncol, int nlay, const double* pressure_hl,
                                                double* out, int64_t capacity) {
  return guarded_call([&]() {
    CkdModel<false> model(file);
    Array3 r = model.calc_rayleigh_optical_depth(mat(pressure_hl, ncol, nlay + 1));
    if ((int64_t)r.size() > capacity) throw 1;
    for (long k = 0; k < (long)r.size(); ++k) out[k] = r.flat(k);
  });
}
