// mock_files.cpp - in-memory stand-in for the reference's NetCDF file layer.
//
// TEST INFRASTRUCTURE ONLY (see oracle/ecckd_oracle.cpp header).
//
// The reference tools (run_ckd, create_look_up_table, reorder_spectrum, find_g_points) read and
// write NetCDF-4 files through DataFile / OutputDataFile (src/include/DataFile.h,
// OutputDataFile.h, src/tools/DataFile.cpp, DataFileEngineNetcdf.cpp, OutputDataFile.cpp).
// NetCDF is not installed here, so oracle/build_ref_tools.sh compiles the reference's tool sources
// unmodified, where they lie, against the Adept API shim (oracle/adept_shim) and THIS file, which
// defines
//   * DataFile::open / open_absolute / init / close  (what src/tools/DataFile.cpp defines): names
//     ending in .cfg go to the reference's own DataFileEngineCfg (compiled in place, with its
//     readconfig.c), everything else to DataFileEngineMem below;
//   * DataFileEngineMem: the read semantics of DataFileEngineNetcdf (leading dimensions sliced
//     with the optional j, i indices; attributes addressed by scope) over a registry of named
//     in-memory "files";
//   * every out-of-line member of OutputDataFile, writing into the same registry;
//   * a small C API (ref_mock_*) with which the test harness fills the input "files" and reads the
//     output "files" back.
// Nothing of the reference's algorithms is restated here: this file moves arrays in and out.
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include <adept_arrays.h>

#include "DataFile.h"
#include "DataFileEngineCfg.h"
#include "Error.h"
#include "OutputDataFile.h"
#include "file_manager.h"

using namespace adept;

std::string OUTPUT_TIME_DIMENSION = "time";
std::string OUTPUT_RANGE_DIMENSION = OUTPUT_RANGE_DIMENSION_NAME;

namespace {

struct MemVar {
  std::vector<std::string> dimnames;
  std::vector<int> dims;
  std::vector<double> data;                     // row-major
  int type = DOUBLE;                            // OutputDataFileVariableType
  std::map<std::string, std::string> sattr;     // string attributes
  std::map<std::string, std::vector<double> > nattr;  // numeric attributes
  long count() const {
    long n = 1;
    for (size_t d = 0; d < dims.size(); ++d) n *= dims[d];
    return n;
  }
};

struct MemFile {
  std::map<std::string, int> dimlen;
  std::map<std::string, MemVar> vars;
  std::map<std::string, std::string> gsattr;
  std::map<std::string, std::vector<double> > gnattr;
};

std::map<std::string, MemFile>& registry() {
  static std::map<std::string, MemFile> r;
  return r;
}

// "files" are addressed by base name: the tools prepend search directories
std::string key_of(const std::string& filename) {
  const size_t p = filename.find_last_of('/');
  return p == std::string::npos ? filename : filename.substr(p + 1);
}

// ---------------------------------------------------------------- reader
class DataFileEngineMem : public DataFileEngine {
 public:
  DataFileEngineMem(const std::string& filename, bool throw_exceptions) : throw_(throw_exceptions) {
    filename_ = filename;
    std::map<std::string, MemFile>::iterator it = registry().find(key_of(filename));
    if (it == registry().end()) {
      ERROR << "mock file \"" << filename << "\" has not been provided";
      THROW(MISSING_MANDATORY_FILE);
    }
    file_ = &it->second;
  }

  virtual IntVector size(const std::string& varname) const {
    const MemVar* v = var(varname);
    if (!v) return IntVector();
    if (v->dims.empty()) return IntVector();
    IntVector s((int)v->dims.size());
    for (size_t d = 0; d < v->dims.size(); ++d) s((int)d) = v->dims[d];
    return s;
  }
  virtual IntVector size(const std::string& scope, const std::string& varname) const {
    IntVector s;
    std::vector<double> num;
    std::string str;
    if (attr(scope, varname, &num, &str)) {
      s.resize(1);
      s(0) = num.empty() ? (int)str.size() : (int)num.size();
    }
    return s;
  }
  virtual bool exist(const std::string& varname) const {
    return file_->vars.count(varname) > 0 || file_->gsattr.count(varname) > 0 || file_->gnattr.count(varname) > 0;
  }
  virtual bool exist(const std::string& scope, const std::string& varname) const {
    return attr(scope, varname, 0, 0);
  }
  virtual bool get_names(std::list<std::string>& names, const std::string& scope = std::string()) const {
    names.clear();
    if (scope.empty() || scope == DATA_FILE_GLOBAL_SCOPE) {
      for (std::map<std::string, MemVar>::const_iterator it = file_->vars.begin(); it != file_->vars.end(); ++it)
        names.push_back(it->first);
    }
    return true;
  }

  // scalars: element (j, i) of a variable with at most two dimensions (DataFileEngineNetcdf.cpp:153-172)
  virtual bool read(Real& x, const std::string& varname, int j = -1, int i = -1) const {
    double v;
    if (!element(varname, j, i, &v)) return false;
    x = v;
    return true;
  }
  virtual bool read(int& x, const std::string& varname, int j = -1, int i = -1) const {
    double v;
    if (!element(varname, j, i, &v)) return false;
    x = (int)v;
    return true;
  }
  virtual bool read(bool& x, const std::string& varname, int j = -1, int i = -1) const {
    double v;
    if (!element(varname, j, i, &v)) return false;
    x = v != 0.0;
    return true;
  }
  virtual bool read(Real& x, const std::string& scope, const std::string& varname, int = -1, int = -1) const {
    std::vector<double> num;
    if (!attr(scope, varname, &num, 0) || num.size() != 1) return fail(scope + "." + varname);
    x = num[0];
    return true;
  }
  virtual bool read(int& x, const std::string& scope, const std::string& varname, int = -1, int = -1) const {
    std::vector<double> num;
    if (!attr(scope, varname, &num, 0) || num.size() != 1) return fail(scope + "." + varname);
    x = (int)num[0];
    return true;
  }
  virtual bool read(bool& x, const std::string& scope, const std::string& varname, int = -1, int = -1) const {
    std::vector<double> num;
    if (!attr(scope, varname, &num, 0) || num.size() != 1) return fail(scope + "." + varname);
    x = num[0] != 0.0;
    return true;
  }
  // strings: a global attribute, or an attribute of `scope`
  virtual bool read(std::string& s, const std::string& varname, int = -1) const {
    std::map<std::string, std::string>::const_iterator it = file_->gsattr.find(varname);
    if (it == file_->gsattr.end()) return fail(varname);
    s = it->second;
    return true;
  }
  virtual bool read(std::string& s, const std::string& scope, const std::string& varname, int = -1) const {
    std::string str;
    std::vector<double> num;
    if (!attr(scope, varname, &num, &str) || !num.empty()) return fail(scope + "." + varname);
    s = str;
    return true;
  }
  virtual bool read(std::string&) const { return false; }

  // arrays: the trailing Rank dimensions of the variable, the leading ones fixed by j then i
  virtual bool read_slice(Vector& v, const std::string& varname, int i = -1) const {
    // a "column" of a matrix variable: element i of every row
    const MemVar* mv = var(varname);
    if (!mv || mv->dims.size() != 2 || i < 0 || i >= mv->dims[1]) return fail(varname);
    v.resize(mv->dims[0]);
    for (int r = 0; r < mv->dims[0]; ++r) v(r) = mv->data[(size_t)r * mv->dims[1] + i];
    return true;
  }
  virtual bool read(Vector& v, const std::string& varname, int j = -1, int i = -1) const {
    const double* p;
    int n[4];
    if (!slice(varname, 1, j, i, &p, n)) return false;
    v.resize(n[0]);
    for (int k = 0; k < n[0]; ++k) v(k) = p[k];
    return true;
  }
  virtual bool read(IntVector& v, const std::string& varname, int j = -1, int i = -1) const {
    const double* p;
    int n[4];
    if (!slice(varname, 1, j, i, &p, n)) return false;
    v.resize(n[0]);
    for (int k = 0; k < n[0]; ++k) v(k) = (int)p[k];
    return true;
  }
  virtual bool read(Matrix& M, const std::string& varname, int j = -1, int i = -1) const {
    const MemVar* mv = var(varname);
    if (mv && mv->dims.size() == 1 && j < 0) {  // a vector read as a matrix is n x 1 (DataFileEngineNetcdf.cpp:607-622)
      M.resize(mv->dims[0], 1);
      for (int k = 0; k < mv->dims[0]; ++k) M(k, 0) = mv->data[k];
      return true;
    }
    const double* p;
    int n[4];
    if (!slice(varname, 2, j, i, &p, n)) return false;
    M.resize(n[0], n[1]);
    for (int a = 0; a < n[0]; ++a)
      for (int b = 0; b < n[1]; ++b) M(a, b) = p[(size_t)a * n[1] + b];
    return true;
  }
  virtual bool read(Array3D& A, const std::string& varname, int j = -1, int i = -1) const {
    const double* p;
    int n[4];
    if (!slice(varname, 3, j, i, &p, n)) return false;
    A.resize(n[0], n[1], n[2]);
    for (int a = 0; a < n[0]; ++a)
      for (int b = 0; b < n[1]; ++b)
        for (int c = 0; c < n[2]; ++c) A(a, b, c) = p[((size_t)a * n[1] + b) * n[2] + c];
    return true;
  }
  virtual bool read(Array4D& A, const std::string& varname, int j = -1, int i = -1) const {
    const double* p;
    int n[4];
    if (!slice(varname, 4, j, i, &p, n)) return false;
    A.resize(n[0], n[1], n[2], n[3]);
    for (int a = 0; a < n[0]; ++a)
      for (int b = 0; b < n[1]; ++b)
        for (int c = 0; c < n[2]; ++c)
          for (int d = 0; d < n[3]; ++d) A(a, b, c, d) = p[(((size_t)a * n[1] + b) * n[2] + c) * n[3] + d];
    return true;
  }
  virtual bool read(Vector& v, const std::string& scope, const std::string& varname) const {
    std::vector<double> num;
    if (!attr(scope, varname, &num, 0) || num.empty()) return fail(scope + "." + varname);
    v.resize((int)num.size());
    for (size_t k = 0; k < num.size(); ++k) v((int)k) = num[k];
    return true;
  }
  virtual bool read(IntVector& v, const std::string& scope, const std::string& varname) const {
    std::vector<double> num;
    if (!attr(scope, varname, &num, 0) || num.empty()) return fail(scope + "." + varname);
    v.resize((int)num.size());
    for (size_t k = 0; k < num.size(); ++k) v((int)k) = (int)num[k];
    return true;
  }
  virtual bool read(Matrix&, const std::string& scope, const std::string& varname) const {
    return fail(scope + "." + varname);
  }
  virtual bool read_missing_value(Real&, const std::string&) const { return false; }
  virtual bool copy_attributes(const std::string&, OutputDataFile*, const std::string& = std::string()) const {
    return false;
  }

 private:
  bool fail(const std::string& what) const {
    if (throw_) {
      ERROR << "mock file " << filename_ << ": cannot read \"" << what << "\"";
      THROW(PRODUCT_FORMAT_ERROR);
    }
    return false;
  }
  const MemVar* var(const std::string& name) const {
    std::map<std::string, MemVar>::const_iterator it = file_->vars.find(name);
    if (it == file_->vars.end()) {
      fail(name);
      return 0;
    }
    return &it->second;
  }
  bool attr(const std::string& scope, const std::string& name, std::vector<double>* num, std::string* str) const {
    const std::map<std::string, std::string>* sa = &file_->gsattr;
    const std::map<std::string, std::vector<double> >* na = &file_->gnattr;
    if (scope != DATA_FILE_GLOBAL_SCOPE) {
      std::map<std::string, MemVar>::const_iterator it = file_->vars.find(scope);
      if (it == file_->vars.end()) return false;
      sa = &it->second.sattr;
      na = &it->second.nattr;
    }
    std::map<std::string, std::string>::const_iterator is = sa->find(name);
    if (is != sa->end()) {
      if (str) *str = is->second;
      if (num) num->clear();
      return true;
    }
    std::map<std::string, std::vector<double> >::const_iterator in = na->find(name);
    if (in != na->end()) {
      if (num) *num = in->second;
      return true;
    }
    return false;
  }
  bool element(const std::string& name, int j, int i, double* out) const {
    const MemVar* v = var(name);
    if (!v) return false;
    if (v->dims.size() > 2) return fail(name);
    if (j < 0) j = 0;
    if (i < 0) i = 0;
    const int n0 = v->dims.size() > 0 ? v->dims[0] : 1, n1 = v->dims.size() > 1 ? v->dims[1] : 1;
    if (j >= n0 || i >= n1) return fail(name);
    *out = v->data[(size_t)j * n1 + i];
    return true;
  }
  // the trailing `rank` dimensions, leading ones fixed by j, i (DataFileEngineNetcdf.cpp:408-457, 555-625)
  bool slice(const std::string& name, int rank, int j, int i, const double** p, int* n) const {
    const MemVar* v = var(name);
    if (!v) return false;
    const int nd = (int)v->dims.size();
    const int lead = nd - rank;
    if (lead < 0 || lead > 2 || (lead >= 1 && j < 0) || (lead == 2 && i < 0)) return fail(name);
    size_t off = 0, inner = 1;
    for (int d = lead; d < nd; ++d) inner *= (size_t)v->dims[d];
    if (lead >= 1) {
      if (j >= v->dims[0]) return fail(name);
      size_t stride = inner;
      if (lead == 2) {
        if (i >= v->dims[1]) return fail(name);
        off += (size_t)i * stride;
        stride *= (size_t)v->dims[1];
      }
      off += (size_t)j * stride;
    }
    for (int d = 0; d < rank; ++d) n[d] = v->dims[lead + d];
    *p = v->data.data() + off;
    return true;
  }
  MemFile* file_;
  bool throw_;
};

MemFile* out_file(const std::string& filename) { return &registry()[key_of(filename)]; }

}  // namespace

// ---------------------------------------------------------------- DataFile (src/tools/DataFile.cpp)
void DataFile::open(const std::string& filename) {
  const size_t dot = filename.find_last_of('.');
  if (dot != std::string::npos && filename.substr(dot + 1) == "cfg") {
    const std::string abs = find_file(filename);
    if (abs.empty()) {
      ERROR << "Failed to find " << filename << " in " << search_directories();
      THROW(MISSING_MANDATORY_FILE);
    }
    open_absolute(abs);
  } else {
    open_absolute(filename);  // mock files live in the registry, not on the search path
  }
}

void DataFile::close() {
  if (engine_) {
    delete engine_;
    engine_ = 0;
  }
}

void DataFile::open_absolute(const std::string& filename) {
  close();
  const size_t dot = filename.find_last_of('.');
  if (dot == std::string::npos) {
    ERROR << "Cannot determine file type for \"" << filename << "\"";
    THROW(PRODUCT_FORMAT_ERROR);
  }
  if (filename.substr(dot + 1) == "cfg") engine_ = new DataFileEngineCfg(filename, throw_exceptions_);
  else engine_ = new DataFileEngineMem(filename, throw_exceptions_);
}

bool DataFile::init(int argc, const char** argv) {
  close();
  engine_ = new DataFileEngineCfg(argc, argv, throw_exceptions_);
  return true;
}

// ---------------------------------------------------------------- OutputDataFile (src/tools/OutputDataFile.cpp)
void OutputDataFile::open(const std::string& filename, OutputDataFileMode mode, DataFileType type) {
  open_absolute(filename, mode, type);
}
void OutputDataFile::open_absolute(const std::string& filename, OutputDataFileMode mode, DataFileType) {
  filename_ = filename;
  type_ = NETCDF;
  ncid_ = 1;
  if (mode == OUTPUT_MODE_CLOBBER) registry()[key_of(filename)] = MemFile();
  else out_file(filename);
}
void OutputDataFile::close() { ncid_ = -1; }
void OutputDataFile::set_options(DataFile&) {}
void OutputDataFile::define_dimension(const std::string& dimname, int length) {
  out_file(filename_)->dimlen[dimname] = length;
}
void OutputDataFile::define_variable(const std::string& varname, OutputDataFileVariableType type,
                                     const std::string& dim1name, const std::string& dim2name,
                                     const std::string& dim3name, const std::string& dim4name) {
  MemFile* f = out_file(filename_);
  MemVar v;
  v.type = (int)type;
  const std::string* names[4] = {&dim1name, &dim2name, &dim3name, &dim4name};
  for (int d = 0; d < 4; ++d) {
    if (names[d]->empty()) break;
    std::map<std::string, int>::const_iterator it = f->dimlen.find(*names[d]);
    if (it == f->dimlen.end()) {
      ERROR << "mock file " << filename_ << ": dimension \"" << *names[d] << "\" of \"" << varname << "\" not defined";
      THROW(PRODUCT_FORMAT_ERROR);
    }
    v.dimnames.push_back(*names[d]);
    v.dims.push_back(it->second);
  }
  v.data.assign((size_t)v.count(), 0.0);
  f->vars[varname] = v;
}
void OutputDataFile::deflate_variable(const std::string&) {}
bool OutputDataFile::write_missing_value(Real missing_value, const std::string& varname) {
  return write(missing_value, varname, "missing_value");
}
bool OutputDataFile::append_title() { return true; }
bool OutputDataFile::append_convention() { return true; }
bool OutputDataFile::append_history(int argc, const char* argv[]) {
  std::string h;
  for (int k = 0; k < argc; ++k) h += std::string(k ? " " : "") + argv[k];
  out_file(filename_)->gsattr["history"] = h;
  return true;
}
void OutputDataFile::end_define_mode() {}
bool OutputDataFile::write_comment(const std::string& comment) {
  out_file(filename_)->gsattr["comment"] = comment;
  return true;
}

namespace {
// FLOAT / INT / SHORT variables hold what the real file would hold after the type conversion
double stored(const MemVar& v, double x) {
  switch (v.type) {
    case FLOAT: return (double)(float)x;
    case INT: return (double)(int)x;
    case SHORT: return (double)(short)x;
    case BYTE: return (double)(signed char)x;
    default: return x;
  }
}
MemVar* out_var(const std::string& filename, const std::string& varname) {
  MemFile* f = out_file(filename);
  std::map<std::string, MemVar>::iterator it = f->vars.find(varname);
  if (it == f->vars.end()) {
    ERROR << "mock file " << filename << ": variable \"" << varname << "\" not defined";
    THROW(PRODUCT_FORMAT_ERROR);
  }
  return &it->second;
}
// destination of a rank-`rank` write with the leading dimensions fixed by j, i
double* out_slice(MemVar* v, int rank, int j, int i, long expect, const std::string& name) {
  const int nd = (int)v->dims.size();
  const int lead = nd - rank;
  long inner = 1;
  for (int d = std::max(lead, 0); d < nd; ++d) inner *= v->dims[d];
  if (lead < 0 || lead > 2 || (lead >= 1 && j < 0) || (lead == 2 && i < 0) || inner != expect) {
    ERROR << "mock file: shape mismatch writing \"" << name << "\"";
    THROW(PRODUCT_FORMAT_ERROR);
  }
  long off = 0, stride = inner;
  if (lead == 2) {
    off += (long)i * stride;
    stride *= v->dims[1];
  }
  if (lead >= 1) off += (long)j * stride;
  return v->data.data() + off;
}
}  // namespace

// scalars: an element of a variable; with a scope, a numeric attribute
bool OutputDataFile::write(Real x, const std::string& varname, int j, int i) {
  MemFile* f = out_file(filename_);
  if (f->vars.find(varname) == f->vars.end()) {  // no such variable: a numeric global attribute
    f->gnattr[varname] = std::vector<double>(1, x);
    return true;
  }
  MemVar* v = out_var(filename_, varname);
  const int rank = 0;
  double* p = out_slice(v, rank, j, i, 1, varname);
  *p = stored(*v, x);
  return true;
}
bool OutputDataFile::write(Real x, const std::string& scope, const std::string& varname, int, int) {
  if (scope == DATA_FILE_GLOBAL_SCOPE) out_file(filename_)->gnattr[varname] = std::vector<double>(1, x);
  else out_var(filename_, scope)->nattr[varname] = std::vector<double>(1, x);
  return true;
}
bool OutputDataFile::write(int x, const std::string& varname, int j, int i) { return write((Real)x, varname, j, i); }
bool OutputDataFile::write(int x, const std::string& scope, const std::string& varname, int j, int i) {
  return write((Real)x, scope, varname, j, i);
}
bool OutputDataFile::write(const std::string& s, const std::string& varname, int) {
  out_file(filename_)->gsattr[varname] = s;  // OutputDataFile.cpp: nc_put_att_text(root, NC_GLOBAL, varname)
  return true;
}
bool OutputDataFile::write(const std::string& s, const std::string& scope, const std::string& varname, int) {
  if (scope == DATA_FILE_GLOBAL_SCOPE) out_file(filename_)->gsattr[varname] = s;
  else out_var(filename_, scope)->sattr[varname] = s;
  return true;
}
bool OutputDataFile::write(const Vector& x, const std::string& varname, int j, int i) {
  MemVar* v = out_var(filename_, varname);
  double* p = out_slice(v, 1, j, i, x.size(), varname);
  for (int k = 0; k < x.size(); ++k) p[k] = stored(*v, x(k));
  return true;
}
bool OutputDataFile::write(const Vector& x, const std::string& scope, const std::string& varname) {
  std::vector<double> a((size_t)x.size());
  for (int k = 0; k < x.size(); ++k) a[(size_t)k] = x(k);
  if (scope == DATA_FILE_GLOBAL_SCOPE) out_file(filename_)->gnattr[varname] = a;
  else out_var(filename_, scope)->nattr[varname] = a;
  return true;
}
bool OutputDataFile::write(const IntVector& x, const std::string& varname, int j, int i) {
  MemVar* v = out_var(filename_, varname);
  double* p = out_slice(v, 1, j, i, x.size(), varname);
  for (int k = 0; k < x.size(); ++k) p[k] = stored(*v, (double)x(k));
  return true;
}
bool OutputDataFile::write(const IntVector& x, const std::string& scope, const std::string& varname) {
  std::vector<double> a((size_t)x.size());
  for (int k = 0; k < x.size(); ++k) a[(size_t)k] = x(k);
  if (scope == DATA_FILE_GLOBAL_SCOPE) out_file(filename_)->gnattr[varname] = a;
  else out_var(filename_, scope)->nattr[varname] = a;
  return true;
}
bool OutputDataFile::write(const Matrix& M, const std::string& varname, int j, int i) {
  MemVar* v = out_var(filename_, varname);
  double* p = out_slice(v, 2, j, i, (long)M.size(0) * M.size(1), varname);
  for (int a = 0; a < M.size(0); ++a)
    for (int b = 0; b < M.size(1); ++b) p[(size_t)a * M.size(1) + b] = stored(*v, M(a, b));
  return true;
}
bool OutputDataFile::write(const Matrix&, const std::string&, const std::string&) { return false; }
bool OutputDataFile::write(const Array3D& A, const std::string& varname, int j, int i) {
  MemVar* v = out_var(filename_, varname);
  double* p = out_slice(v, 3, j, i, (long)A.size(0) * A.size(1) * A.size(2), varname);
  for (int a = 0; a < A.size(0); ++a)
    for (int b = 0; b < A.size(1); ++b)
      for (int c = 0; c < A.size(2); ++c) p[((size_t)a * A.size(1) + b) * A.size(2) + c] = stored(*v, A(a, b, c));
  return true;
}
bool OutputDataFile::exist(const std::string& varname) { return out_file(filename_)->vars.count(varname) > 0; }
bool OutputDataFile::exist(const std::string& scope, const std::string& varname) {
  MemFile* f = out_file(filename_);
  if (scope == DATA_FILE_GLOBAL_SCOPE) return f->gsattr.count(varname) > 0 || f->gnattr.count(varname) > 0;
  std::map<std::string, MemVar>::const_iterator it = f->vars.find(scope);
  return it != f->vars.end() && (it->second.sattr.count(varname) > 0 || it->second.nattr.count(varname) > 0);
}
void OutputDataFile::set_chunking(const IntVector&, const std::string&) {}

// ---------------------------------------------------------------- harness API
#define MOCK_API extern "C" __attribute__((visibility("default")))

MOCK_API void ref_mock_reset() { registry().clear(); }
MOCK_API void ref_mock_put(const char* file, const char* var, int ndim, const int* dims, const double* data) {
  MemVar v;
  for (int d = 0; d < ndim; ++d) v.dims.push_back(dims[d]);
  v.data.assign(data, data + v.count());
  registry()[file].vars[var] = v;
}
MOCK_API void ref_mock_put_string(const char* file, const char* scope, const char* name, const char* value) {
  MemFile& f = registry()[file];
  if (!scope || !*scope) f.gsattr[name] = value;
  else f.vars[scope].sattr[name] = value;
}
MOCK_API void ref_mock_put_number(const char* file, const char* scope, const char* name, double value) {
  MemFile& f = registry()[file];
  if (!scope || !*scope) f.gnattr[name] = std::vector<double>(1, value);
  else f.vars[scope].nattr[name] = std::vector<double>(1, value);
}
// dims of a variable: returns ndim (or -1), writes up to 4 extents
MOCK_API int ref_mock_dims(const char* file, const char* var, int* dims) {
  std::map<std::string, MemFile>::const_iterator f = registry().find(file);
  if (f == registry().end()) return -1;
  std::map<std::string, MemVar>::const_iterator v = f->second.vars.find(var);
  if (v == f->second.vars.end()) return -1;
  for (size_t d = 0; d < v->second.dims.size() && d < 4; ++d) dims[d] = v->second.dims[d];
  return (int)v->second.dims.size();
}
MOCK_API int ref_mock_get(const char* file, const char* var, double* out, long capacity) {
  std::map<std::string, MemFile>::const_iterator f = registry().find(file);
  if (f == registry().end()) return -1;
  std::map<std::string, MemVar>::const_iterator v = f->second.vars.find(var);
  if (v == f->second.vars.end() || (long)v->second.data.size() > capacity) return -1;
  std::memcpy(out, v->second.data.data(), v->second.data.size() * sizeof(double));
  return 0;
}
MOCK_API int ref_mock_get_string(const char* file, const char* scope, const char* name, char* out, int capacity) {
  std::map<std::string, MemFile>::const_iterator f = registry().find(file);
  if (f == registry().end()) return -1;
  const std::map<std::string, std::string>* sa = &f->second.gsattr;
  if (scope && *scope) {
    std::map<std::string, MemVar>::const_iterator v = f->second.vars.find(scope);
    if (v == f->second.vars.end()) return -1;
    sa = &v->second.sattr;
  }
  std::map<std::string, std::string>::const_iterator it = sa->find(name);
  if (it == sa->end() || (int)it->second.size() + 1 > capacity) return -1;
  std::strcpy(out, it->second.c_str());
  return 0;
}
// names of the variables of a file, '\n'-separated
MOCK_API int ref_mock_list(const char* file, char* out, int capacity) {
  std::map<std::string, MemFile>::const_iterator f = registry().find(file);
  if (f == registry().end()) return -1;
  std::string s;
  for (std::map<std::string, MemVar>::const_iterator v = f->second.vars.begin(); v != f->second.vars.end(); ++v)
    s += v->first + "\n";
  if ((int)s.size() + 1 > capacity) return -1;
  std::strcpy(out, s.c_str());
  return 0;
}
