// adept_arrays.h - minimal, eagerly evaluated stand-in for the Adept 2.1 array API.
//
// TEST INFRASTRUCTURE ONLY (see oracle/ecckd_oracle.cpp header).
//
// Purpose: let the reference's own hot-path translation units
// (average_optical_depth.cpp, planck_function.cpp, radiative_transfer_{lw,sw}.cpp,
// heating_rate.h, single_gas_data.cpp, calc_cost_function_{lw,sw}.cpp,
// rayleigh_scattering.h) compile UNMODIFIED, where they lie under /root/reference,
// without the Adept library (not installed, not installable offline).  It is written
// from Adept's documented semantics (User Guide ch. 3), not from
// Adept's sources:
//   * Array<Rank,T,IsActive>: row-major, 0-based, reference-counted storage;
//     copy construction makes a shallow alias; `a = b` copies elements into
//     existing storage (resizing only when `a` is empty); `a >>= b` re-links.
//   * `__` selects a whole dimension, `range(a,b)` is inclusive, `end` is the
//     last index of the dimension it is used in; integer-vector arguments gather
//     (and scatter, when assigned to).
//   * `sum/minval/maxval(x)` reduce everything, `(x, d)` reduce dimension d;
//     `find(mask)` returns ascending indices; `x.where(m) = either_or(a,b)`.
// Differences that matter for parity and are therefore stated: every expression is
// evaluated eagerly into a temporary (no expression templates), reductions run
// sequentially in ascending index order, `exp/log/pow` are libm's (the reference is
// built with -DADEPT_FAST_EXPONENTIAL), and "active" (automatic differentiation)
// arrays are plain arrays: only the forward values of the reference code are
// reproduced, never its derivatives.
#ifndef ECCKD_ADEPT_SHIM_ARRAYS_H
#define ECCKD_ADEPT_SHIM_ARRAYS_H 1

#include <algorithm>
#include <cmath>
#include <initializer_list>
#include <cstddef>
#include <iostream>
#include <memory>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <vector>

namespace adept {

typedef double Real;
typedef int Index;
typedef double aReal;  // forward values only

struct size_mismatch : public std::runtime_error {
  explicit size_mismatch(const std::string& s) : std::runtime_error(s) {}
};
struct index_out_of_bounds : public std::runtime_error {
  explicit index_out_of_bounds(const std::string& s) : std::runtime_error(s) {}
};

// ---------------------------------------------------------------- index helpers
struct AllIndex {};
static const AllIndex __ = AllIndex();

struct EndIndex {
  int offset;
  int resolve(int n) const { return n - 1 + offset; }
};
static const EndIndex end = {0};
inline EndIndex operator-(EndIndex e, int i) { EndIndex r = {e.offset - i}; return r; }
inline EndIndex operator+(EndIndex e, int i) { EndIndex r = {e.offset + i}; return r; }

struct RangeIndex {
  bool b_end, e_end;
  int b, e;
  int first(int n) const { return b_end ? n - 1 + b : b; }
  int last(int n) const { return e_end ? n - 1 + e : e; }
};
inline RangeIndex range(int a, int b) { RangeIndex r = {false, false, a, b}; return r; }
inline RangeIndex range(int a, EndIndex b) { RangeIndex r = {false, true, a, b.offset}; return r; }
inline RangeIndex range(EndIndex a, EndIndex b) { RangeIndex r = {true, true, a.offset, b.offset}; return r; }
inline RangeIndex range(EndIndex a, int b) { RangeIndex r = {true, false, a.offset, b}; return r; }

// ExpressionSize<Rank>: what `dimensions(n0, n1, ...)`, `expression_size(...)` and `a.dimensions()` return
template <int Rank>
struct ExpressionSize {
  int dim[Rank > 0 ? Rank : 1];
  int operator[](int d) const { return dim[d]; }
  int& operator[](int d) { return dim[d]; }
};
inline ExpressionSize<1> dimensions(int n0) { ExpressionSize<1> e; e.dim[0] = n0; return e; }
inline ExpressionSize<2> dimensions(int n0, int n1) { ExpressionSize<2> e; e.dim[0] = n0; e.dim[1] = n1; return e; }
inline ExpressionSize<3> dimensions(int n0, int n1, int n2) {
  ExpressionSize<3> e; e.dim[0] = n0; e.dim[1] = n1; e.dim[2] = n2; return e;
}
inline ExpressionSize<4> dimensions(int n0, int n1, int n2, int n3) {
  ExpressionSize<4> e; e.dim[0] = n0; e.dim[1] = n1; e.dim[2] = n2; e.dim[3] = n3; return e;
}
inline ExpressionSize<1> expression_size(int n0) { return dimensions(n0); }
inline ExpressionSize<2> expression_size(int n0, int n1) { return dimensions(n0, n1); }
template <int Rank>
std::ostream& operator<<(std::ostream& os, const ExpressionSize<Rank>& e) {
  os << "(";
  for (int d = 0; d < Rank; ++d) os << (d ? "," : "") << e.dim[d];
  return os << ")";
}

// stride(a, b, step): indices a, a+step, ..., <= b
struct StrideIndex { int b, e, step; };
inline StrideIndex stride(int a, int b, int step) { StrideIndex s = {a, b, step}; return s; }

template <int Rank, typename T, bool IsActive = false> class Array;
template <int Rank, typename T, bool IsActive> class ScatterRef;
template <int Rank, typename T, bool IsActive> class WhereRef;

template <typename A, typename B>
struct EitherOr {
  const A& a;
  const B& b;
};
template <typename A, typename B>
EitherOr<A, B> either_or(const A& a, const B& b) {
  EitherOr<A, B> r = {a, b};
  return r;
}

namespace shim {
template <typename X> struct is_array : std::false_type {};
template <int R, typename T, bool A> struct is_array<Array<R, T, A> > : std::true_type {};
template <int R, typename T, bool A> struct is_array<ScatterRef<R, T, A> > : std::true_type {};
template <typename S> struct is_scalar : std::is_arithmetic<S> {};

// one index argument of operator(): an integer, `end±k`, `__`, a range, or an index vector
struct Sel {
  enum Kind { SCALAR, ALL, RANGE, LIST } kind;
  int i;
  bool i_end;
  RangeIndex r;
  const int* list;
  int nlist;
  std::shared_ptr<std::vector<int> > keep;  // owns a contiguous copy of a strided index vector
};
inline Sel sel(int i) { Sel s; s.kind = Sel::SCALAR; s.i = i; s.i_end = false; s.list = 0; s.nlist = 0; return s; }
inline Sel sel(EndIndex e) { Sel s; s.kind = Sel::SCALAR; s.i = e.offset; s.i_end = true; s.list = 0; s.nlist = 0; return s; }
inline Sel sel(AllIndex) { Sel s; s.kind = Sel::ALL; s.i = 0; s.i_end = false; s.list = 0; s.nlist = 0; return s; }
inline Sel sel(RangeIndex r) { Sel s; s.kind = Sel::RANGE; s.i = 0; s.i_end = false; s.r = r; s.list = 0; s.nlist = 0; return s; }
template <bool A> Sel sel(const Array<1, int, A>& v);

template <typename X> struct sel_rank { static const int value = 1; };  // kept dimensions
template <> struct sel_rank<int> { static const int value = 0; };
template <> struct sel_rank<EndIndex> { static const int value = 0; };
template <typename X> struct sel_is_list : std::false_type {};
template <bool A> struct sel_is_list<Array<1, int, A> > : std::true_type {};
}  // namespace shim

// ---------------------------------------------------------------- Array
template <int Rank, typename T, bool IsActive>
class Array {
 public:
  typedef T value_type;
  static const int rank = Rank;

  Array() : data_(0) { for (int d = 0; d < Rank; ++d) dim_[d] = stride_[d] = 0; }
  Array(const Array& o) : store_(o.store_), data_(o.data_) { copy_shape(o); }  // shallow alias
  template <bool B>
  Array(const Array<Rank, T, B>& o) : store_(o.storage()), data_(const_cast<T*>(o.data())) {
    for (int d = 0; d < Rank; ++d) { dim_[d] = o.size(d); stride_[d] = o.stride(d); }
  }
  explicit Array(int n0) : data_(0) { int n[1] = {n0}; static_assert(Rank == 1, "rank"); alloc(n); }
  Array(int n0, int n1) : data_(0) { int n[2] = {n0, n1}; static_assert(Rank == 2, "rank"); alloc(n); }
  Array(int n0, int n1, int n2) : data_(0) { int n[3] = {n0, n1, n2}; static_assert(Rank == 3, "rank"); alloc(n); }
  Array(int n0, int n1, int n2, int n3) : data_(0) {
    int n[4] = {n0, n1, n2, n3}; static_assert(Rank == 4, "rank"); alloc(n);
  }
  explicit Array(const ExpressionSize<Rank>& e) : data_(0) { alloc(e.dim); }
  // Adept: an array that points at caller-owned, contiguous row-major memory
  Array(T* p, const ExpressionSize<Rank>& e) : data_(p) {
    long st = 1;
    for (int d = Rank - 1; d >= 0; --d) { dim_[d] = e.dim[d]; stride_[d] = st; st *= e.dim[d]; }
  }
  // element-type conversion (Adept evaluates `const Vector& = intVector` into a temporary)
  template <typename U, bool B, typename std::enable_if<!std::is_same<U, T>::value && std::is_arithmetic<U>::value &&
                                                         std::is_arithmetic<T>::value, int>::type = 0>
  Array(const Array<Rank, U, B>& o) : data_(0) {
    for (int d = 0; d < Rank; ++d) dim_[d] = stride_[d] = 0;
    if (!o.empty()) {
      int n[Rank];
      for (int d = 0; d < Rank; ++d) n[d] = o.size(d);
      alloc(n);
      const long tot = size();
      for (long k = 0; k < tot; ++k) data_[k] = (T)o.flat(k);
    }
  }
  // `Vector v = {a, b, ...}`
  Array(std::initializer_list<T> il) : data_(0) {
    static_assert(Rank == 1, "rank");
    int n[1] = {(int)il.size()};
    alloc(n);
    int k = 0;
    for (typename std::initializer_list<T>::const_iterator it = il.begin(); it != il.end(); ++it) data_[k++] = *it;
  }
  // strided view of a vector
  Array operator()(const StrideIndex& si) const {
    static_assert(Rank == 1, "rank");
    if (si.step <= 0 || si.e < si.b) throw index_out_of_bounds("adept shim: stride");
    const int cnt = (si.e - si.b) / si.step + 1;
    bounds(si.b, 0);
    bounds(si.b + (cnt - 1) * si.step, 0);
    int dims[1] = {cnt};
    long str[1] = {stride_[0] * si.step};
    Array r = wrap(data_ + stride_[0] * si.b, dims, str);
    r.adopt(store_);
    return r;
  }
  // `Vector v = range(a, b)`
  Array(const RangeIndex& r) : data_(0) {
    static_assert(Rank == 1, "rank");
    const int a = r.first(0), b = r.last(0);
    int n[1] = {b >= a ? b - a + 1 : 0};
    alloc(n);
    for (int k = 0; k < n[0]; ++k) data_[k] = (T)(a + k);
  }
  // wrap caller-owned memory (used by the glue; not part of Adept's API surface used by the reference)
  static Array wrap(T* p, const int* dims, const long* strides) {
    Array a;
    a.data_ = p;
    for (int d = 0; d < Rank; ++d) { a.dim_[d] = dims[d]; a.stride_[d] = strides[d]; }
    return a;
  }

  // ---- shape
  int size() const { long n = 1; for (int d = 0; d < Rank; ++d) n *= dim_[d]; return (int)n; }
  int size(int d) const { return dim_[d]; }
  int dimension(int d) const { return dim_[d]; }
  long stride(int d) const { return stride_[d]; }
  bool empty() const { return size() == 0; }
  T* data() { return data_; }
  const T* data() const { return data_; }
  const T* const_data() const { return data_; }
  long offset(int d) const { return stride_[d]; }  // Adept's name for the stride of dimension d
  const std::shared_ptr<T>& storage() const { return store_; }
  void clear() { store_.reset(); data_ = 0; for (int d = 0; d < Rank; ++d) dim_[d] = stride_[d] = 0; }
  void resize(int n0) { int n[1] = {n0}; static_assert(Rank == 1, "rank"); alloc(n); }
  void resize(int n0, int n1) { int n[2] = {n0, n1}; static_assert(Rank == 2, "rank"); alloc(n); }
  void resize(int n0, int n1, int n2) { int n[3] = {n0, n1, n2}; static_assert(Rank == 3, "rank"); alloc(n); }
  void resize(int n0, int n1, int n2, int n3) { int n[4] = {n0, n1, n2, n3}; static_assert(Rank == 4, "rank"); alloc(n); }
  void resize(const ExpressionSize<Rank>& e) { alloc(e.dim); }
  ExpressionSize<Rank> dimensions() const { ExpressionSize<Rank> e; for (int d = 0; d < Rank; ++d) e.dim[d] = dim_[d]; return e; }
  Array<Rank, T, false> inactive_link() const { return Array<Rank, T, false>(*this); }
  // a contiguous vector seen as a higher-rank array sharing its storage
  template <int R2>
  Array<R2, T, IsActive> reshape(const ExpressionSize<R2>& e) const {
    static_assert(Rank == 1, "reshape of a vector");
    long tot = 1;
    for (int d = 0; d < R2; ++d) tot *= e.dim[d];
    if (tot != size() || (size() > 1 && stride_[0] != 1)) throw size_mismatch("adept shim: reshape");
    long str[R2];
    long st = 1;
    for (int d = R2 - 1; d >= 0; --d) { str[d] = st; st *= e.dim[d]; }
    Array<R2, T, IsActive> r = Array<R2, T, IsActive>::wrap(data_, e.dim, str);
    r.adopt(store_);
    return r;
  }
  void resize_like(const int* n) { alloc(n); }
  Array soft_link() const { return *this; }
  Array& operator>>=(const Array& o) { store_ = o.store_; data_ = o.data_; copy_shape(o); return *this; }

  // ---- flat element access in row-major logical order
  T& flat(long k) { return data_[offset_of(k)]; }
  const T& flat(long k) const { return data_[offset_of(k)]; }

  // ---- assignment: copies elements (Adept: `=` never re-links a non-empty array)
  Array& operator=(const Array& o) { return assign(o); }
  template <typename U, bool B>
  Array& operator=(const Array<Rank, U, B>& o) { return assign(o); }
  template <typename S, typename std::enable_if<shim::is_scalar<S>::value, int>::type = 0>
  Array& operator=(S s) {
    const long n = size();
    for (long k = 0; k < n; ++k) flat(k) = (T)s;
    return *this;
  }
  Array& operator=(const RangeIndex& r) { return assign(Array<1, int, false>(r)); }
#define ECCKD_SHIM_COMPOUND(OP)                                                          \
  template <typename U, bool B>                                                          \
  Array& operator OP(const Array<Rank, U, B>& o) {                                       \
    check_same(o);                                                                       \
    Array<Rank, U, B> tmp = o.overlaps(*this) ? o.copy() : o;                            \
    const long n = size();                                                               \
    for (long k = 0; k < n; ++k) flat(k) OP tmp.flat(k);                                 \
    return *this;                                                                        \
  }                                                                                      \
  template <typename S, typename std::enable_if<shim::is_scalar<S>::value, int>::type = 0> \
  Array& operator OP(S s) {                                                              \
    const long n = size();                                                               \
    for (long k = 0; k < n; ++k) flat(k) OP s;                                           \
    return *this;                                                                        \
  }
  ECCKD_SHIM_COMPOUND(+=)
  ECCKD_SHIM_COMPOUND(-=)
  ECCKD_SHIM_COMPOUND(*=)
  ECCKD_SHIM_COMPOUND(/=)
#undef ECCKD_SHIM_COMPOUND

  Array copy() const {
    Array r;
    if (empty()) return r;
    r.alloc(dim_);
    const long n = size();
    for (long k = 0; k < n; ++k) r.data_[k] = flat(k);
    return r;
  }
  template <typename U, bool B>
  bool overlaps(const Array<Rank, U, B>& o) const {
    return (const void*)store_.get() != 0 && (const void*)store_.get() == (const void*)o.storage().get();
  }

  // ---- indexing.  Rank 1
  template <typename I0>
  typename std::enable_if<shim::sel_rank<I0>::value == 0 && Rank == 1, T&>::type operator()(const I0& i0) {
    return data_[stride_[0] * resolve(shim::sel(i0), 0)];
  }
  template <typename I0>
  typename std::enable_if<shim::sel_rank<I0>::value == 0 && Rank == 1, const T&>::type operator()(const I0& i0) const {
    return data_[stride_[0] * resolve(shim::sel(i0), 0)];
  }
  template <typename I0>
  typename std::enable_if<shim::sel_rank<I0>::value == 1 && !shim::sel_is_list<I0>::value && Rank == 1, Array>::type
  operator()(const I0& i0) const {
    shim::Sel s[1] = {shim::sel(i0)};
    return view<1>(s);
  }
  template <bool B>
  ScatterRef<1, T, IsActive> operator()(const Array<1, int, B>& idx) const {
    static_assert(Rank == 1, "rank");
    shim::Sel s[1] = {shim::sel(idx)};
    return gather<1>(s);
  }
  // `v[i]`: element of a vector / slice of the first dimension
  template <int R = Rank>
  typename std::enable_if<R == 1, T&>::type operator[](int i) { return (*this)(i); }
  template <int R = Rank>
  typename std::enable_if<R == 1, const T&>::type operator[](int i) const { return (*this)(i); }
  template <int R = Rank>
  typename std::enable_if<(R > 1), Array<Rank - 1, T, IsActive> >::type operator[](int i) const {
    Array<Rank - 1, T, IsActive> r;
    int dims[Rank];
    long str[Rank];
    for (int d = 1; d < Rank; ++d) { dims[d - 1] = dim_[d]; str[d - 1] = stride_[d]; }
    bounds(i, 0);
    r = Array<Rank - 1, T, IsActive>::wrap(data_ + stride_[0] * i, dims, str);
    r.adopt(store_);
    return r;
  }

  // Rank 2
  template <typename I0, typename I1>
  typename std::enable_if<shim::sel_rank<I0>::value + shim::sel_rank<I1>::value == 0 && Rank == 2, T&>::type
  operator()(const I0& i0, const I1& i1) {
    return data_[stride_[0] * resolve(shim::sel(i0), 0) + stride_[1] * resolve(shim::sel(i1), 1)];
  }
  template <typename I0, typename I1>
  typename std::enable_if<shim::sel_rank<I0>::value + shim::sel_rank<I1>::value == 0 && Rank == 2, const T&>::type
  operator()(const I0& i0, const I1& i1) const {
    return data_[stride_[0] * resolve(shim::sel(i0), 0) + stride_[1] * resolve(shim::sel(i1), 1)];
  }
  template <typename I0, typename I1>
  typename std::enable_if<(shim::sel_rank<I0>::value + shim::sel_rank<I1>::value > 0) &&
                              !shim::sel_is_list<I0>::value && !shim::sel_is_list<I1>::value && Rank == 2,
                          Array<shim::sel_rank<I0>::value + shim::sel_rank<I1>::value, T, IsActive> >::type
  operator()(const I0& i0, const I1& i1) const {
    shim::Sel s[2] = {shim::sel(i0), shim::sel(i1)};
    return view<shim::sel_rank<I0>::value + shim::sel_rank<I1>::value>(s);
  }
  template <typename I0, typename I1>
  typename std::enable_if<(shim::sel_is_list<I0>::value || shim::sel_is_list<I1>::value) && Rank == 2,
                          ScatterRef<shim::sel_rank<I0>::value + shim::sel_rank<I1>::value, T, IsActive> >::type
  operator()(const I0& i0, const I1& i1) const {
    shim::Sel s[2] = {shim::sel(i0), shim::sel(i1)};
    return gather<shim::sel_rank<I0>::value + shim::sel_rank<I1>::value>(s);
  }

  // Rank 3
  template <typename I0, typename I1, typename I2>
  typename std::enable_if<shim::sel_rank<I0>::value + shim::sel_rank<I1>::value + shim::sel_rank<I2>::value == 0 &&
                              Rank == 3, T&>::type
  operator()(const I0& i0, const I1& i1, const I2& i2) {
    return data_[stride_[0] * resolve(shim::sel(i0), 0) + stride_[1] * resolve(shim::sel(i1), 1) +
                 stride_[2] * resolve(shim::sel(i2), 2)];
  }
  template <typename I0, typename I1, typename I2>
  typename std::enable_if<shim::sel_rank<I0>::value + shim::sel_rank<I1>::value + shim::sel_rank<I2>::value == 0 &&
                              Rank == 3, const T&>::type
  operator()(const I0& i0, const I1& i1, const I2& i2) const {
    return data_[stride_[0] * resolve(shim::sel(i0), 0) + stride_[1] * resolve(shim::sel(i1), 1) +
                 stride_[2] * resolve(shim::sel(i2), 2)];
  }
  template <typename I0, typename I1, typename I2>
  typename std::enable_if<(shim::sel_rank<I0>::value + shim::sel_rank<I1>::value + shim::sel_rank<I2>::value > 0) &&
                              !shim::sel_is_list<I0>::value && !shim::sel_is_list<I1>::value &&
                              !shim::sel_is_list<I2>::value && Rank == 3,
                          Array<shim::sel_rank<I0>::value + shim::sel_rank<I1>::value + shim::sel_rank<I2>::value, T,
                                IsActive> >::type
  operator()(const I0& i0, const I1& i1, const I2& i2) const {
    shim::Sel s[3] = {shim::sel(i0), shim::sel(i1), shim::sel(i2)};
    return view<shim::sel_rank<I0>::value + shim::sel_rank<I1>::value + shim::sel_rank<I2>::value>(s);
  }

  // Rank 4
  template <typename I0, typename I1, typename I2, typename I3>
  typename std::enable_if<shim::sel_rank<I0>::value + shim::sel_rank<I1>::value + shim::sel_rank<I2>::value +
                                  shim::sel_rank<I3>::value == 0 && Rank == 4, T&>::type
  operator()(const I0& i0, const I1& i1, const I2& i2, const I3& i3) {
    return data_[stride_[0] * resolve(shim::sel(i0), 0) + stride_[1] * resolve(shim::sel(i1), 1) +
                 stride_[2] * resolve(shim::sel(i2), 2) + stride_[3] * resolve(shim::sel(i3), 3)];
  }
  template <typename I0, typename I1, typename I2, typename I3>
  typename std::enable_if<shim::sel_rank<I0>::value + shim::sel_rank<I1>::value + shim::sel_rank<I2>::value +
                                  shim::sel_rank<I3>::value == 0 && Rank == 4, const T&>::type
  operator()(const I0& i0, const I1& i1, const I2& i2, const I3& i3) const {
    return data_[stride_[0] * resolve(shim::sel(i0), 0) + stride_[1] * resolve(shim::sel(i1), 1) +
                 stride_[2] * resolve(shim::sel(i2), 2) + stride_[3] * resolve(shim::sel(i3), 3)];
  }
  template <typename I0, typename I1, typename I2, typename I3>
  typename std::enable_if<(shim::sel_rank<I0>::value + shim::sel_rank<I1>::value + shim::sel_rank<I2>::value +
                           shim::sel_rank<I3>::value > 0) &&
                              !shim::sel_is_list<I0>::value && !shim::sel_is_list<I1>::value &&
                              !shim::sel_is_list<I2>::value && !shim::sel_is_list<I3>::value && Rank == 4,
                          Array<shim::sel_rank<I0>::value + shim::sel_rank<I1>::value + shim::sel_rank<I2>::value +
                                    shim::sel_rank<I3>::value, T, IsActive> >::type
  operator()(const I0& i0, const I1& i1, const I2& i2, const I3& i3) const {
    shim::Sel s[4] = {shim::sel(i0), shim::sel(i1), shim::sel(i2), shim::sel(i3)};
    return view<shim::sel_rank<I0>::value + shim::sel_rank<I1>::value + shim::sel_rank<I2>::value +
                shim::sel_rank<I3>::value>(s);
  }

  // ---- where
  template <bool B>
  WhereRef<Rank, T, IsActive> where(const Array<Rank, bool, B>& mask) {
    check_same(mask);
    return WhereRef<Rank, T, IsActive>(*this, mask);
  }

  void adopt(const std::shared_ptr<T>& s) { store_ = s; }

  template <typename U, bool B>
  void check_same(const Array<Rank, U, B>& o) const {
    for (int d = 0; d < Rank; ++d)
      if (dim_[d] != o.size(d))
        throw size_mismatch("adept shim: array size mismatch in expression");
  }

 private:
  template <int, typename, bool> friend class Array;

  void copy_shape(const Array& o) { for (int d = 0; d < Rank; ++d) { dim_[d] = o.dim_[d]; stride_[d] = o.stride_[d]; } }
  void alloc(const int* n) {
    long tot = 1;
    for (int d = 0; d < Rank; ++d) { if (n[d] < 0) throw size_mismatch("adept shim: negative dimension"); tot *= n[d]; }
    // a zero-sized array keeps its other extents (e.g. x(__,empty_index) is nlay x 0)
    if (tot == 0) { store_.reset(); data_ = 0; }
    else { store_ = std::shared_ptr<T>(new T[(size_t)tot](), std::default_delete<T[]>()); data_ = store_.get(); }
    long s = 1;
    for (int d = Rank - 1; d >= 0; --d) { dim_[d] = n[d]; stride_[d] = s; s *= n[d]; }
  }
  long offset_of(long k) const {
    long off = 0;
    for (int d = Rank - 1; d >= 0; --d) { off += (k % dim_[d]) * stride_[d]; k /= dim_[d]; }
    return off;
  }
  void bounds(int i, int d) const {
    if (i < 0 || i >= dim_[d]) throw index_out_of_bounds("adept shim: index out of bounds");
  }
  int resolve(const shim::Sel& s, int d) const {
    const int i = s.i_end ? dim_[d] - 1 + s.i : s.i;
    bounds(i, d);
    return i;
  }
  template <typename U, bool B>
  Array& assign(const Array<Rank, U, B>& o) {
    if (empty()) {
      if (o.empty()) return *this;
      int n[Rank];
      for (int d = 0; d < Rank; ++d) n[d] = o.size(d);
      alloc(n);
    } else {
      check_same(o);
    }
    const long n = size();
    if (o.overlaps(*this)) {
      Array<Rank, U, B> tmp = o.copy();
      for (long k = 0; k < n; ++k) flat(k) = (T)tmp.flat(k);
    } else {
      for (long k = 0; k < n; ++k) flat(k) = (T)o.flat(k);
    }
    return *this;
  }
  // strided view (no index lists)
  template <int R2>
  Array<R2, T, IsActive> view(const shim::Sel* s) const {
    int dims[R2 > 0 ? R2 : 1];
    long str[R2 > 0 ? R2 : 1];
    T* p = data_;
    int k = 0;
    for (int d = 0; d < Rank; ++d) {
      if (s[d].kind == shim::Sel::SCALAR) {
        p += stride_[d] * resolve(s[d], d);
      } else if (s[d].kind == shim::Sel::ALL) {
        dims[k] = dim_[d]; str[k] = stride_[d]; ++k;
      } else {
        const int a = s[d].r.first(dim_[d]), b = s[d].r.last(dim_[d]);
        if (b < a - 1) throw index_out_of_bounds("adept shim: reversed range");
        if (b >= a) { bounds(a, d); bounds(b, d); }
        p += stride_[d] * a;
        dims[k] = b - a + 1; str[k] = stride_[d]; ++k;
      }
    }
    Array<R2, T, IsActive> r = Array<R2, T, IsActive>::wrap(p, dims, str);
    r.adopt(store_);
    return r;
  }
  // gather through index lists -> a fresh array that remembers where its elements came from
  template <int R2>
  ScatterRef<R2, T, IsActive> gather(const shim::Sel* s) const;

  std::shared_ptr<T> store_;
  T* data_;
  int dim_[Rank];
  long stride_[Rank];
};

typedef Array<1, Real, false> Vector;
typedef Array<2, Real, false> Matrix;
typedef Array<3, Real, false> Array3D;
typedef Array<4, Real, false> Array4D;
typedef Array<1, Real, true> aVector;
typedef Array<2, Real, true> aMatrix;
typedef Array<3, Real, true> aArray3D;
typedef Array<1, int, false> intVector;
typedef Array<2, int, false> intMatrix;
typedef Array<1, int, false> IntVector;  // Adept spells the integer arrays both ways
typedef Array<2, int, false> IntMatrix;
typedef Array<3, Real, false> Array3;
typedef Array<3, Real, true> aArray3;
typedef Array<4, Real, true> aArray4D;
typedef Array<3, int, false> intArray3D;
typedef Array<2, Real, false> SymmMatrix;  // stored full (only the prior-covariance code of CkdModel uses it)
typedef Array<1, bool, false> boolVector;
typedef Array<2, bool, false> boolMatrix;

namespace shim {
template <bool A>
Sel sel(const Array<1, int, A>& v) {
  Sel s;
  s.kind = Sel::LIST; s.i = 0; s.i_end = false;
  s.nlist = v.size();
  if (v.size() > 0 && v.stride(0) != 1) {
    s.keep = std::make_shared<std::vector<int> >((size_t)v.size());
    for (int k = 0; k < v.size(); ++k) (*s.keep)[k] = v(k);
    s.list = s.keep->data();
  } else {
    s.list = v.data();
  }
  return s;
}
}  // namespace shim

// Result of indexing with an integer vector: the gathered values, plus the addresses they came
// from so that `x(index) = y`, `x(index) *= s` write through (Adept's IndexedArray lvalue).
template <int Rank, typename T, bool IsActive>
class ScatterRef : public Array<Rank, T, IsActive> {
 public:
  typedef Array<Rank, T, IsActive> Base;
  ScatterRef() {}
  std::vector<T*> target;
  template <typename U, bool B>
  ScatterRef& operator=(const Array<Rank, U, B>& o) {
    this->check_same(o);
    const long n = this->size();
    Array<Rank, U, B> tmp = o.copy();
    for (long k = 0; k < n; ++k) { *target[k] = (T)tmp.flat(k); this->flat(k) = (T)tmp.flat(k); }
    return *this;
  }
  ScatterRef& operator=(const ScatterRef& o) { return operator=(static_cast<const Base&>(o)); }
  ScatterRef& operator=(const RangeIndex& r) { return operator=(Array<1, int, false>(r)); }
  template <typename S, typename std::enable_if<shim::is_scalar<S>::value, int>::type = 0>
  ScatterRef& operator=(S s) {
    for (size_t k = 0; k < target.size(); ++k) { *target[k] = (T)s; this->flat((long)k) = (T)s; }
    return *this;
  }
#define ECCKD_SHIM_SCATTER(OP)                                                           \
  template <typename S, typename std::enable_if<shim::is_scalar<S>::value, int>::type = 0> \
  ScatterRef& operator OP(S s) {                                                         \
    for (size_t k = 0; k < target.size(); ++k) { *target[k] OP s; this->flat((long)k) = *target[k]; } \
    return *this;                                                                        \
  }                                                                                      \
  template <typename U, bool B>                                                          \
  ScatterRef& operator OP(const Array<Rank, U, B>& o) {                                  \
    this->check_same(o);                                                                 \
    Array<Rank, U, B> tmp = o.copy();                                                    \
    for (size_t k = 0; k < target.size(); ++k) { *target[k] OP tmp.flat((long)k); this->flat((long)k) = *target[k]; } \
    return *this;                                                                        \
  }
  ECCKD_SHIM_SCATTER(+=)
  ECCKD_SHIM_SCATTER(-=)
  ECCKD_SHIM_SCATTER(*=)
  ECCKD_SHIM_SCATTER(/=)
#undef ECCKD_SHIM_SCATTER
};

template <int Rank, typename T, bool IsActive>
template <int R2>
ScatterRef<R2, T, IsActive> Array<Rank, T, IsActive>::gather(const shim::Sel* s) const {
  ScatterRef<R2, T, IsActive> r;
  int dims[R2 > 0 ? R2 : 1];
  int first[Rank], count[Rank];
  int k = 0;
  bool none = false;
  for (int d = 0; d < Rank; ++d) {
    switch (s[d].kind) {
      case shim::Sel::SCALAR: first[d] = resolve(s[d], d); count[d] = 1; break;
      case shim::Sel::ALL: first[d] = 0; count[d] = dim_[d]; dims[k++] = dim_[d]; break;
      case shim::Sel::RANGE: {
        const int a = s[d].r.first(dim_[d]), b = s[d].r.last(dim_[d]);
        if (b >= a) { bounds(a, d); bounds(b, d); }
        first[d] = a; count[d] = b - a + 1; dims[k++] = count[d];
        break;
      }
      case shim::Sel::LIST:
        first[d] = 0; count[d] = s[d].nlist; dims[k++] = count[d];
        for (int j = 0; j < s[d].nlist; ++j) bounds(s[d].list[j], d);
        break;
    }
    if (count[d] <= 0) none = true;
  }
  // an empty selection keeps its other extents so that `sum(x(__,empty),1)` has the right length
  for (int d = 0; d < R2; ++d) if (dims[d] < 0) dims[d] = 0;
  r.resize_like(dims);
  if (none) return r;
  long n = r.size();
  r.target.resize((size_t)n);
  int idx[Rank];
  for (long f = 0; f < n; ++f) {
    long rem = f;
    for (int d = Rank - 1; d >= 0; --d) { idx[d] = (int)(rem % count[d]); rem /= count[d]; }
    long off = 0;
    for (int d = 0; d < Rank; ++d) {
      const int i = s[d].kind == shim::Sel::LIST ? s[d].list[idx[d]] : first[d] + idx[d];
      off += stride_[d] * i;
    }
    r.target[(size_t)f] = data_ + off;
    r.flat(f) = data_[off];
  }
  return r;
}

template <int Rank, typename T, bool IsActive>
class WhereRef {
 public:
  template <bool B>
  WhereRef(Array<Rank, T, IsActive>& a, const Array<Rank, bool, B>& m) : a_(a), m_(m) {}
  template <typename S, typename std::enable_if<shim::is_scalar<S>::value, int>::type = 0>
  void operator=(S s) {
    const long n = a_.size();
    for (long k = 0; k < n; ++k) if (m_.flat(k)) a_.flat(k) = (T)s;
  }
  template <typename U, bool B>
  void operator=(const Array<Rank, U, B>& o) {
    a_.check_same(o);
    Array<Rank, U, B> tmp = o.overlaps(a_) ? o.copy() : o;
    const long n = a_.size();
    for (long k = 0; k < n; ++k) if (m_.flat(k)) a_.flat(k) = (T)tmp.flat(k);
  }
  template <typename A, typename B>
  void operator=(const EitherOr<A, B>& e) {
    const long n = a_.size();
    for (long k = 0; k < n; ++k) a_.flat(k) = m_.flat(k) ? (T)pick(e.a, k) : (T)pick(e.b, k);
  }

 private:
  template <typename S>
  static typename std::enable_if<shim::is_scalar<S>::value, S>::type pick(const S& s, long) { return s; }
  template <typename U, bool B>
  U pick(const Array<Rank, U, B>& x, long k) { a_.check_same(x); return x.flat(k); }
  Array<Rank, T, IsActive> a_;
  Array<Rank, bool, false> m_;
};

// ---------------------------------------------------------------- element-wise operators
namespace shim {
template <typename A, typename B> struct promote { typedef typename std::common_type<A, B>::type type; };
template <int R, typename T, bool A>
Array<R, T, A> like(const Array<R, T, A>&, const int* dims) {
  Array<R, T, A> r;
  r.resize_like(dims);
  return r;
}
template <int R, typename U, bool A>
void dims_of(const Array<R, U, A>& x, int* d) { for (int i = 0; i < R; ++i) d[i] = x.size(i); }
}  // namespace shim

#define ECCKD_SHIM_BINARY(OPNAME, EXPR, RESULT_T)                                                    \
  template <int R, typename T, bool A, typename U, bool B>                                           \
  Array<R, RESULT_T(T, U), (A || B)> OPNAME(const Array<R, T, A>& x, const Array<R, U, B>& y) {       \
    x.check_same(y);                                                                                 \
    Array<R, RESULT_T(T, U), (A || B)> r;                                                            \
    if (x.empty()) return r;                                                                         \
    int d[R]; shim::dims_of(x, d); r.resize_like(d);                                                 \
    const long n = x.size();                                                                         \
    for (long k = 0; k < n; ++k) { const T a = x.flat(k); const U b = y.flat(k); r.flat(k) = (EXPR); } \
    return r;                                                                                        \
  }                                                                                                  \
  template <int R, typename T, bool A, typename S,                                                   \
            typename std::enable_if<shim::is_scalar<S>::value, int>::type = 0>                       \
  Array<R, RESULT_T(T, S), A> OPNAME(const Array<R, T, A>& x, S b) {                                  \
    Array<R, RESULT_T(T, S), A> r;                                                                   \
    if (x.empty()) return r;                                                                         \
    int d[R]; shim::dims_of(x, d); r.resize_like(d);                                                 \
    const long n = x.size();                                                                         \
    for (long k = 0; k < n; ++k) { const T a = x.flat(k); r.flat(k) = (EXPR); }                      \
    return r;                                                                                        \
  }                                                                                                  \
  template <int R, typename T, bool A, typename S,                                                   \
            typename std::enable_if<shim::is_scalar<S>::value, int>::type = 0>                       \
  Array<R, RESULT_T(S, T), A> OPNAME(S a, const Array<R, T, A>& y) {                                  \
    Array<R, RESULT_T(S, T), A> r;                                                                   \
    if (y.empty()) return r;                                                                         \
    int d[R]; shim::dims_of(y, d); r.resize_like(d);                                                 \
    const long n = y.size();                                                                         \
    for (long k = 0; k < n; ++k) { const T b = y.flat(k); r.flat(k) = (EXPR); }                      \
    return r;                                                                                        \
  }
#define ECCKD_SHIM_ARITH_T(T, U) typename shim::promote<T, U>::type
#define ECCKD_SHIM_BOOL_T(T, U) bool
ECCKD_SHIM_BINARY(operator+, a + b, ECCKD_SHIM_ARITH_T)
ECCKD_SHIM_BINARY(operator-, a - b, ECCKD_SHIM_ARITH_T)
ECCKD_SHIM_BINARY(operator*, a * b, ECCKD_SHIM_ARITH_T)
ECCKD_SHIM_BINARY(operator/, a / b, ECCKD_SHIM_ARITH_T)
ECCKD_SHIM_BINARY(operator==, a == b, ECCKD_SHIM_BOOL_T)
ECCKD_SHIM_BINARY(operator!=, a != b, ECCKD_SHIM_BOOL_T)
ECCKD_SHIM_BINARY(operator<, a < b, ECCKD_SHIM_BOOL_T)
ECCKD_SHIM_BINARY(operator<=, a <= b, ECCKD_SHIM_BOOL_T)
ECCKD_SHIM_BINARY(operator>, a > b, ECCKD_SHIM_BOOL_T)
ECCKD_SHIM_BINARY(operator>=, a >= b, ECCKD_SHIM_BOOL_T)
ECCKD_SHIM_BINARY(operator&&, a && b, ECCKD_SHIM_BOOL_T)
ECCKD_SHIM_BINARY(operator||, a || b, ECCKD_SHIM_BOOL_T)
// Adept's binary min/max/pow on arrays (std::min/std::max on two scalars stay the std ones)
ECCKD_SHIM_BINARY(min, (b < a ? (typename shim::promote<decltype(a), decltype(b)>::type)b : (typename shim::promote<decltype(a), decltype(b)>::type)a), ECCKD_SHIM_ARITH_T)
ECCKD_SHIM_BINARY(max, (a < b ? (typename shim::promote<decltype(a), decltype(b)>::type)b : (typename shim::promote<decltype(a), decltype(b)>::type)a), ECCKD_SHIM_ARITH_T)
ECCKD_SHIM_BINARY(pow, std::pow((double)a, (double)b), ECCKD_SHIM_ARITH_T)
#undef ECCKD_SHIM_BINARY

#define ECCKD_SHIM_UNARY(NAME, EXPR)                                    \
  template <int R, typename T, bool A>                                  \
  Array<R, T, A> NAME(const Array<R, T, A>& x) {                        \
    Array<R, T, A> r;                                                   \
    if (x.empty()) return r;                                            \
    int d[R]; shim::dims_of(x, d); r.resize_like(d);                    \
    const long n = x.size();                                            \
    for (long k = 0; k < n; ++k) { const T a = x.flat(k); r.flat(k) = (EXPR); } \
    return r;                                                           \
  }
ECCKD_SHIM_UNARY(operator-, -a)
ECCKD_SHIM_UNARY(exp, std::exp(a))
ECCKD_SHIM_UNARY(log, std::log(a))
ECCKD_SHIM_UNARY(sqrt, std::sqrt(a))
ECCKD_SHIM_UNARY(abs, std::abs(a))
ECCKD_SHIM_UNARY(fabs, std::fabs(a))
#undef ECCKD_SHIM_UNARY
template <int R, bool A>
Array<R, bool, A> operator!(const Array<R, bool, A>& x) {
  Array<R, bool, A> r;
  if (x.empty()) return r;
  int d[R]; shim::dims_of(x, d); r.resize_like(d);
  for (long k = 0; k < (long)x.size(); ++k) r.flat(k) = !x.flat(k);
  return r;
}
// scalar forms used next to the array ones after `using namespace adept`
using std::abs;
using std::exp;
using std::fabs;
using std::log;
using std::pow;
using std::sqrt;

// ---------------------------------------------------------------- reductions
template <int R, typename T, bool A>
T sum(const Array<R, T, A>& x) {
  T s = 0;
  const long n = x.size();
  for (long k = 0; k < n; ++k) s += x.flat(k);
  return s;
}
template <int R, typename T, bool A>
T maxval(const Array<R, T, A>& x) {
  const long n = x.size();
  if (n == 0) throw size_mismatch("adept shim: maxval of an empty array");
  T s = x.flat(0);
  for (long k = 1; k < n; ++k) if (s < x.flat(k)) s = x.flat(k);
  return s;
}
template <int R, typename T, bool A>
T minval(const Array<R, T, A>& x) {
  const long n = x.size();
  if (n == 0) throw size_mismatch("adept shim: minval of an empty array");
  T s = x.flat(0);
  for (long k = 1; k < n; ++k) if (x.flat(k) < s) s = x.flat(k);
  return s;
}
template <int R, bool A>
bool any(const Array<R, bool, A>& x) {
  const long n = x.size();
  for (long k = 0; k < n; ++k) if (x.flat(k)) return true;
  return false;
}
template <int R, bool A>
bool all(const Array<R, bool, A>& x) {
  const long n = x.size();
  for (long k = 0; k < n; ++k) if (!x.flat(k)) return false;
  return true;
}
template <int R, bool A>
int count(const Array<R, bool, A>& x) {
  int c = 0;
  for (long k = 0; k < (long)x.size(); ++k) c += x.flat(k) ? 1 : 0;
  return c;
}

namespace shim {
// reduce dimension `dim` of a rank-2 array; OP(acc, value, first)
template <typename T, bool A, typename F>
Array<1, T, A> reduce2(const Array<2, T, A>& x, int dim, F f, bool need_data) {
  Array<1, T, A> r;
  const int n0 = x.size(0), n1 = x.size(1);
  const int nout = dim == 0 ? n1 : n0, nred = dim == 0 ? n0 : n1;
  if (nout == 0) return r;
  r.resize(nout);
  for (int o = 0; o < nout; ++o) {
    if (nred == 0) {
      if (need_data) throw size_mismatch("adept shim: minval/maxval over an empty dimension");
      r(o) = 0;
      continue;
    }
    T acc = dim == 0 ? x(0, o) : x(o, 0);
    for (int k = 1; k < nred; ++k) acc = f(acc, dim == 0 ? x(k, o) : x(o, k));
    r(o) = acc;
  }
  return r;
}
}  // namespace shim

template <typename T, bool A>
Array<1, T, A> sum(const Array<2, T, A>& x, int dim) {
  return shim::reduce2(x, dim, [](T a, T b) { return a + b; }, false);
}
template <typename T, bool A>
Array<1, T, A> maxval(const Array<2, T, A>& x, int dim) {
  return shim::reduce2(x, dim, [](T a, T b) { return a < b ? b : a; }, true);
}
template <typename T, bool A>
Array<1, T, A> minval(const Array<2, T, A>& x, int dim) {
  return shim::reduce2(x, dim, [](T a, T b) { return b < a ? b : a; }, true);
}

template <bool A>
intVector find(const Array<1, bool, A>& mask) {
  const int n = mask.size();
  int c = 0;
  for (int k = 0; k < n; ++k) c += mask(k) ? 1 : 0;
  intVector r;
  if (c == 0) return r;
  r.resize(c);
  c = 0;
  for (int k = 0; k < n; ++k) if (mask(k)) r(c++) = k;
  return r;
}

template <int Dim, typename T, bool A>
Array<2, T, A> spread(const Array<1, T, A>& v, int n) {
  static_assert(Dim == 0 || Dim == 1, "spread dimension");
  Array<2, T, A> r;
  if (v.empty() || n <= 0) return r;
  if (Dim == 1) {
    r.resize(v.size(), n);
    for (int i = 0; i < v.size(); ++i) for (int j = 0; j < n; ++j) r(i, j) = v(i);
  } else {
    r.resize(n, v.size());
    for (int j = 0; j < n; ++j) for (int i = 0; i < v.size(); ++i) r(j, i) = v(i);
  }
  return r;
}

// spread of an integer range (Adept treats range(a,b) as an integer vector expression)
template <int Dim>
Array<2, int, false> spread(const RangeIndex& r, int n) { return spread<Dim>(Array<1, int, false>(r), n); }
// spread of a matrix to rank 3: the new dimension of length n is inserted at position Dim
template <int Dim, typename T, bool A>
Array<3, T, A> spread(const Array<2, T, A>& m, int n) {
  static_assert(Dim >= 0 && Dim <= 2, "spread dimension");
  Array<3, T, A> r;
  if (m.empty() || n <= 0) return r;
  const int n0 = m.size(0), n1 = m.size(1);
  if (Dim == 0) r.resize(n, n0, n1);
  else if (Dim == 1) r.resize(n0, n, n1);
  else r.resize(n0, n1, n);
  for (int k = 0; k < n; ++k)
    for (int i = 0; i < n0; ++i)
      for (int j = 0; j < n1; ++j) {
        if (Dim == 0) r(k, i, j) = m(i, j);
        else if (Dim == 1) r(i, k, j) = m(i, j);
        else r(i, j, k) = m(i, j);
      }
  return r;
}
// arithmetic on integer ranges: `dwav * range(a, b)`
template <typename S, typename std::enable_if<shim::is_scalar<S>::value, int>::type = 0>
Array<1, typename shim::promote<S, int>::type, false> operator*(S a, const RangeIndex& r) {
  return a * Array<1, int, false>(r);
}

template <int R, typename T, bool A>
T mean(const Array<R, T, A>& x) { return sum(x) / (T)x.size(); }
template <int R, typename T, bool A>
T product(const Array<R, T, A>& x) {
  T s = 1;
  for (long k = 0; k < (long)x.size(); ++k) s *= x.flat(k);
  return s;
}
template <typename T, bool A, typename U, bool B>
typename shim::promote<T, U>::type dot_product(const Array<1, T, A>& x, const Array<1, U, B>& y) {
  x.check_same(y);
  typename shim::promote<T, U>::type s = 0;
  for (int k = 0; k < x.size(); ++k) s += x(k) * y(k);
  return s;
}
// spread of a rank-3 array along a new leading dimension
template <int Dim, typename T, bool A>
Array<4, T, A> spread(const Array<3, T, A>& m, int n) {
  static_assert(Dim == 0, "spread dimension");
  Array<4, T, A> r;
  if (m.empty() || n <= 0) return r;
  r.resize(n, m.size(0), m.size(1), m.size(2));
  for (int k = 0; k < n; ++k)
    for (int i = 0; i < m.size(0); ++i)
      for (int j = 0; j < m.size(1); ++j)
        for (int q = 0; q < m.size(2); ++q) r(k, i, j, q) = m(i, j, q);
  return r;
}

// sum of a rank-3 array over one dimension
template <typename T, bool A>
Array<2, T, A> sum(const Array<3, T, A>& x, int dim) {
  const int n0 = x.size(0), n1 = x.size(1), n2 = x.size(2);
  Array<2, T, A> r;
  if (dim == 0) r.resize(n1, n2);
  else if (dim == 1) r.resize(n0, n2);
  else r.resize(n0, n1);
  if (r.empty()) return r;
  r = (T)0;
  for (int i = 0; i < n0; ++i)
    for (int j = 0; j < n1; ++j)
      for (int k = 0; k < n2; ++k) {
        if (dim == 0) r(j, k) += x(i, j, k);
        else if (dim == 1) r(i, k) += x(i, j, k);
        else r(i, j) += x(i, j, k);
      }
  return r;
}

// Piecewise-linear interpolation interp(x, y, xi): x ascending; outside [x0, xn] the end values are
// held (the reference only interpolates inside the range; the policy outside is not pinned).
namespace shim {
inline void interp_locate(const Array<1, Real, false>& x, Real xi, int& i0, Real& w1) {
  const int n = x.size();
  if (n == 1 || xi <= x(0)) { i0 = 0; w1 = 0.0; return; }
  if (xi >= x(n - 1)) { i0 = n - 2; w1 = 1.0; return; }
  int lo = 0, hi = n - 1;
  while (hi - lo > 1) { const int mid = (lo + hi) / 2; if (x(mid) <= xi) lo = mid; else hi = mid; }
  i0 = lo;
  w1 = (xi - x(lo)) / (x(lo + 1) - x(lo));
}
}  // namespace shim
template <bool A, bool B, bool C>
Array<1, Real, B> interp(const Array<1, Real, A>& x, const Array<1, Real, B>& y, const Array<1, Real, C>& xi) {
  if (x.size() != y.size() || x.size() < 1) throw size_mismatch("adept shim: interp");
  Array<1, Real, false> xs(x);
  Array<1, Real, B> r;
  if (xi.empty()) return r;
  r.resize(xi.size());
  for (int k = 0; k < xi.size(); ++k) {
    int i0; Real w1;
    shim::interp_locate(xs, xi(k), i0, w1);
    r(k) = x.size() == 1 ? y(0) : (1.0 - w1) * y(i0) + w1 * y(i0 + 1);
  }
  return r;
}
// y(nx, m): every column interpolated
template <bool A, bool B, bool C>
Array<2, Real, B> interp(const Array<1, Real, A>& x, const Array<2, Real, B>& y, const Array<1, Real, C>& xi) {
  if (x.size() != y.size(0) || x.size() < 1) throw size_mismatch("adept shim: interp");
  Array<1, Real, false> xs(x);
  Array<2, Real, B> r;
  if (xi.empty() || y.size(1) == 0) return r;
  r.resize(xi.size(), y.size(1));
  for (int k = 0; k < xi.size(); ++k) {
    int i0; Real w1;
    shim::interp_locate(xs, xi(k), i0, w1);
    for (int j = 0; j < y.size(1); ++j) r(k, j) = x.size() == 1 ? y(0, j) : (1.0 - w1) * y(i0, j) + w1 * y(i0 + 1, j);
  }
  return r;
}

// Matrix multiplication `A ** B` = A * (*B): unary * tags the right operand
template <int R, typename T, bool A>
struct MatMulRhs { Array<R, T, A> x; };
template <int R, typename T, bool A>
MatMulRhs<R, T, A> operator*(const Array<R, T, A>& x) { MatMulRhs<R, T, A> m; m.x = x.copy(); return m; }
template <typename T, bool A, typename U, bool B>
Array<1, typename shim::promote<T, U>::type, (A || B)> operator*(const Array<2, T, A>& m, const MatMulRhs<1, U, B>& v) {
  if (m.size(1) != v.x.size()) throw size_mismatch("adept shim: matrix-vector product");
  Array<1, typename shim::promote<T, U>::type, (A || B)> r;
  if (m.size(0) == 0) return r;
  r.resize(m.size(0));
  for (int i = 0; i < m.size(0); ++i) {
    typename shim::promote<T, U>::type s = 0;
    for (int j = 0; j < m.size(1); ++j) s += m(i, j) * v.x(j);
    r(i) = s;
  }
  return r;
}
template <typename T, bool A, typename U, bool B>
Array<2, typename shim::promote<T, U>::type, (A || B)> operator*(const Array<2, T, A>& m, const MatMulRhs<2, U, B>& n) {
  if (m.size(1) != n.x.size(0)) throw size_mismatch("adept shim: matrix-matrix product");
  Array<2, typename shim::promote<T, U>::type, (A || B)> r;
  if (m.size(0) == 0 || n.x.size(1) == 0) return r;
  r.resize(m.size(0), n.x.size(1));
  for (int i = 0; i < m.size(0); ++i)
    for (int k = 0; k < n.x.size(1); ++k) {
      typename shim::promote<T, U>::type s = 0;
      for (int j = 0; j < m.size(1); ++j) s += m(i, j) * n.x(j, k);
      r(i, k) = s;
    }
  return r;
}
// inverse of a (symmetric, well-conditioned) matrix by Gauss-Jordan with partial pivoting; only the
// prior-covariance code of CkdModel calls it (outside the hot path, compiled for completeness)
template <bool A>
Array<2, Real, A> inv(const Array<2, Real, A>& m) {
  const int n = m.size(0);
  if (m.size(1) != n) throw size_mismatch("adept shim: inv of a non-square matrix");
  Array<2, Real, A> a = m.copy(), r;
  if (n == 0) return r;
  r.resize(n, n);
  r = 0.0;
  for (int i = 0; i < n; ++i) r(i, i) = 1.0;
  for (int c = 0; c < n; ++c) {
    int piv = c;
    for (int i = c + 1; i < n; ++i) if (std::fabs(a(i, c)) > std::fabs(a(piv, c))) piv = i;
    if (a(piv, c) == 0.0) throw std::runtime_error("adept shim: singular matrix in inv");
    if (piv != c)
      for (int j = 0; j < n; ++j) { std::swap(a(c, j), a(piv, j)); std::swap(r(c, j), r(piv, j)); }
    const Real d = 1.0 / a(c, c);
    for (int j = 0; j < n; ++j) { a(c, j) *= d; r(c, j) *= d; }
    for (int i = 0; i < n; ++i) {
      if (i == c) continue;
      const Real f = a(i, c);
      if (f == 0.0) continue;
      for (int j = 0; j < n; ++j) { a(i, j) -= f * a(c, j); r(i, j) -= f * r(c, j); }
    }
  }
  return r;
}

template <int R, typename T, bool A>
Array<R, T, A> eval(const Array<R, T, A>& x) { return x.copy(); }

template <int R, typename T, bool A>
std::ostream& operator<<(std::ostream& os, const Array<R, T, A>& x) {
  os << "{";
  for (long k = 0; k < (long)x.size(); ++k) os << (k ? ", " : "") << x.flat(k);
  return os << "}";
}

enum ArrayPrintStyle { PRINT_STYLE_PLAIN, PRINT_STYLE_CSV, PRINT_STYLE_CURLY, PRINT_STYLE_MATLAB };
inline void set_array_print_style(ArrayPrintStyle) {}
// value(): the passive value of an active scalar / array
inline Real value(Real x) { return x; }
template <int R, typename T, bool A>
Array<R, T, false> value(const Array<R, T, A>& x) { return Array<R, T, false>(x); }

// Stack stand-in so that code mentioning it compiles; records nothing.
class Stack {
 public:
  void new_recording() {}
  void pause_recording() {}
  void continue_recording() {}
};

}  // namespace adept

#endif
