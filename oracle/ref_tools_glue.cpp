// ref_tools_glue.cpp - C ABI over the reference's own tool mains.
//
// TEST INFRASTRUCTURE ONLY (see oracle/ecckd_oracle.cpp header).  No algorithm is restated here:
// oracle/build_ref_tools.sh compiles run_ckd.cpp, create_look_up_table.cpp, reorder_spectrum.cpp and
// find_g_points.cpp where they lie under /root/reference/src/ecckd with `main` renamed, and this
// file calls them with an argv assembled by the test harness, turning the reference's
// `throw <ESA exit code>` (Logging.h:115-117) into a return value.
#include <exception>
#include <string>

#define REF_API extern "C" __attribute__((visibility("default")))

int ref_main_run_ckd(int argc, const char* argv[]);
int ref_main_create_look_up_table(int argc, const char* argv[]);
int ref_main_reorder_spectrum(int argc, const char* argv[]);
int ref_main_find_g_points(int argc, const char* argv[]);

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
