/*
 * oracle/shim/RcppArmadillo.h -- TEST INFRASTRUCTURE ONLY.
 *
 * A throw-away stand-in for the subset of Armadillo / Rcpp / R-nmath that the reference's hot-path
 * sources (src/LNA_functional.cpp, basic_util.cpp, SIR_LNA.cpp, SIRS_period.cpp) touch, so that those
 * sources can be compiled VERBATIM, where they lie under /root/reference, into oracle/_ref/ (recipe:
 * oracle/Makefile target `ref`).  R, Rcpp, RcppArmadillo and LAPACK are absent from the image; nothing of
 * theirs is copied here -- this file only restates the documented behaviour of the ~40 features used:
 *
 *   arma::mat / vec / ivec / cube (column-major, 16-element local buffer like arma's mat_prealloc,
 *   zero-filled on construction), element access with bounds checks, submat/subvec/row/col/cols/slice views,
 *   the `M << a << b << endr` element injector, + - * / % on dense operands, t(), min(), has_nan(),
 *   exp/log/sqrt, eye/ones/zeros/diagmat, chol (upper, throws std::runtime_error on a non-positive pivot --
 *   LAPACK dpotrf's failure condition), det, inv, solve (partial-pivot LU as dgesv), norm, sum/accu, randn;
 *   Rcpp::List (named/positional slots holding double/int/bool/mat/cube/List), as<T>, XPtr<T>, Rcout;
 *   R::runif.
 *
 * The random draws are third-party in the reference (arma::randn over R's RNG, R::runif); here they read
 * PRE-DRAWN streams installed by the harness (shim_rng), in the call order of the reference code -- which is
 * exactly how the oracle and the CUDA path take them.
 */
#ifndef LNA_SHIM_RCPPARMADILLO_H_
#define LNA_SHIM_RCPPARMADILLO_H_

#include <cmath>
#include <math.h>
#include <cstring>
#include <iostream>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

// ------------------------------------------------------------------------------------------------
// pre-drawn random streams
// ------------------------------------------------------------------------------------------------
struct shim_rng_state {
    const double *normals = nullptr, *uniforms = nullptr;
    long n_norm = 0, n_unif = 0;      // stream lengths
    long used_norm = 0, used_unif = 0;  // consumed so far
};
inline shim_rng_state &shim_rng() {
    static thread_local shim_rng_state s;
    return s;
}
inline double shim_next_normal() {
    shim_rng_state &s = shim_rng();
    if (s.used_norm >= s.n_norm) throw std::out_of_range("shim: normal stream exhausted");
    return s.normals[s.used_norm++];
}
inline double shim_next_uniform() {
    shim_rng_state &s = shim_rng();
    if (s.used_unif >= s.n_unif) throw std::out_of_range("shim: uniform stream exhausted");
    return s.uniforms[s.used_unif++];
}

namespace arma {

typedef unsigned int uword;
typedef int sword;
using std::cerr;
using std::cout;
using std::endl;
using std::size_t;

struct endr_t {};
static const endr_t endr = endr_t();

template <class T>
class SubView;
template <class T>
class MatT;

[[noreturn]] inline void shim_bounds(const char *what) { throw std::logic_error(std::string("index out of bounds: ") + what); }

template <class T>
class MatT {
  public:
    uword n_rows = 0, n_cols = 0, n_elem = 0;

    MatT() {}
    MatT(long r, long c) { init(r, c); }
    MatT(const MatT &o) {
        init_nofill(o.n_rows, o.n_cols);
        if (n_elem) std::memcpy(mem_, o.mem_, sizeof(T) * n_elem);
    }
    MatT(MatT &&o) noexcept { steal(o); }
    MatT(const SubView<T> &v);
    ~MatT() { release(); }

    MatT &operator=(const MatT &o) {
        if (this != &o) {
            resize_nofill(o.n_rows, o.n_cols);
            if (n_elem) std::memcpy(mem_, o.mem_, sizeof(T) * n_elem);
        }
        return *this;
    }
    MatT &operator=(MatT &&o) noexcept {
        if (this != &o) {
            release();
            steal(o);
        }
        return *this;
    }
    MatT &operator=(const SubView<T> &v) { return *this = MatT(v); }

    T *memptr() { return mem_; }
    const T *memptr() const { return mem_; }

    T &operator[](uword i) { return mem_[i]; }
    const T &operator[](uword i) const { return mem_[i]; }
    T &operator()(long i) {
        if (i < 0 || (uword)i >= n_elem) shim_bounds("Mat::operator()");
        return mem_[i];
    }
    const T &operator()(long i) const {
        if (i < 0 || (uword)i >= n_elem) shim_bounds("Mat::operator()");
        return mem_[i];
    }
    T &operator()(long r, long c) {
        if (r < 0 || c < 0 || (uword)r >= n_rows || (uword)c >= n_cols) shim_bounds("Mat::operator()");
        return mem_[r + (size_t)c * n_rows];
    }
    const T &operator()(long r, long c) const {
        if (r < 0 || c < 0 || (uword)r >= n_rows || (uword)c >= n_cols) shim_bounds("Mat::operator()");
        return mem_[r + (size_t)c * n_rows];
    }
    T &at(uword r, uword c) { return mem_[r + (size_t)c * n_rows]; }
    const T &at(uword r, uword c) const { return mem_[r + (size_t)c * n_rows]; }

    uword size() const { return n_elem; }
    MatT &zeros() {
        for (uword i = 0; i < n_elem; ++i) mem_[i] = T(0);
        return *this;
    }
    MatT &zeros(long r, long c) {
        resize_nofill(r, c);
        return zeros();
    }
    MatT &ones() {
        for (uword i = 0; i < n_elem; ++i) mem_[i] = T(1);
        return *this;
    }
    MatT &fill(T v) {
        for (uword i = 0; i < n_elem; ++i) mem_[i] = v;
        return *this;
    }
    void set_size(long r, long c) { resize_nofill(r, c); }
    MatT &operator+=(T s) {
        for (uword i = 0; i < n_elem; ++i) mem_[i] += s;
        return *this;
    }
    MatT &operator-=(T s) {
        for (uword i = 0; i < n_elem; ++i) mem_[i] -= s;
        return *this;
    }
    MatT &operator*=(T s) {
        for (uword i = 0; i < n_elem; ++i) mem_[i] *= s;
        return *this;
    }
    MatT &operator/=(T s) {
        for (uword i = 0; i < n_elem; ++i) mem_[i] /= s;
        return *this;
    }
    MatT &operator+=(const MatT &o) {
        if (o.n_rows != n_rows || o.n_cols != n_cols) throw std::logic_error("addition: incompatible matrix dimensions");
        for (uword i = 0; i < n_elem; ++i) mem_[i] += o.mem_[i];
        return *this;
    }
    MatT &operator-=(const MatT &o) {
        if (o.n_rows != n_rows || o.n_cols != n_cols) throw std::logic_error("subtraction: incompatible matrix dimensions");
        for (uword i = 0; i < n_elem; ++i) mem_[i] -= o.mem_[i];
        return *this;
    }

    bool has_nan() const {
        for (uword i = 0; i < n_elem; ++i)
            if (mem_[i] != mem_[i]) return true;
        return false;
    }
    T min() const {
        if (!n_elem) throw std::logic_error("min(): object has no elements");
        T m = mem_[0];
        for (uword i = 1; i < n_elem; ++i)
            if (mem_[i] < m) m = mem_[i];
        return m;
    }
    T max() const {
        if (!n_elem) throw std::logic_error("max(): object has no elements");
        T m = mem_[0];
        for (uword i = 1; i < n_elem; ++i)
            if (mem_[i] > m) m = mem_[i];
        return m;
    }
    MatT t() const {
        MatT r;
        r.init_nofill(n_cols, n_rows);
        for (uword j = 0; j < n_cols; ++j)
            for (uword i = 0; i < n_rows; ++i) r.mem_[j + (size_t)i * n_cols] = mem_[i + (size_t)j * n_rows];
        return r;
    }

    // views (non-const: assignable; const: a copy)
    SubView<T> submat(long r1, long c1, long r2, long c2) { return SubView<T>(*this, r1, c1, r2, c2); }
    MatT submat(long r1, long c1, long r2, long c2) const { return MatT(SubView<T>(const_cast<MatT &>(*this), r1, c1, r2, c2)); }
    SubView<T> row(long r) { return SubView<T>(*this, r, 0, r, (long)n_cols - 1); }
    MatT row(long r) const { return submat(r, 0, r, (long)n_cols - 1); }
    SubView<T> col(long c) { return SubView<T>(*this, 0, c, (long)n_rows - 1, c); }
    MatT col(long c) const { return submat(0, c, (long)n_rows - 1, c); }
    SubView<T> cols(long c1, long c2) { return SubView<T>(*this, 0, c1, (long)n_rows - 1, c2); }
    MatT cols(long c1, long c2) const { return submat(0, c1, (long)n_rows - 1, c2); }
    SubView<T> rows(long r1, long r2) { return SubView<T>(*this, r1, 0, r2, (long)n_cols - 1); }
    MatT rows(long r1, long r2) const { return submat(r1, 0, r2, (long)n_cols - 1); }
    SubView<T> subvec(long a, long b) {
        return n_cols == 1 ? SubView<T>(*this, a, 0, b, 0) : SubView<T>(*this, 0, a, 0, b);
    }
    MatT subvec(long a, long b) const { return MatT(const_cast<MatT &>(*this).subvec(a, b)); }

    // used by the shim itself
    void init(long r, long c) {
        init_nofill(r, c);
        for (uword i = 0; i < n_elem; ++i) mem_[i] = T(0);
    }
    void init_nofill(long r, long c) {
        if (r < 0 || c < 0) throw std::logic_error("Mat: negative size");
        n_rows = (uword)r;
        n_cols = (uword)c;
        n_elem = n_rows * n_cols;
        mem_ = n_elem <= kLocal ? local_ : new T[n_elem];
    }
    void resize_nofill(long r, long c) {
        if ((uword)r == n_rows && (uword)c == n_cols && mem_) return;
        if ((size_t)r * (size_t)c == n_elem && mem_) {
            n_rows = (uword)r;
            n_cols = (uword)c;
            return;
        }
        release();
        init_nofill(r, c);
    }

  protected:
    static const uword kLocal = 16;
    T *mem_ = local_;
    T local_[kLocal];

    void release() {
        if (mem_ != local_ && mem_) delete[] mem_;
        mem_ = local_;
        n_rows = n_cols = n_elem = 0;
    }
    void steal(MatT &o) {
        n_rows = o.n_rows;
        n_cols = o.n_cols;
        n_elem = o.n_elem;
        if (o.mem_ == o.local_) {
            mem_ = local_;
            if (n_elem) std::memcpy(local_, o.local_, sizeof(T) * n_elem);
        } else {
            mem_ = o.mem_;
            o.mem_ = o.local_;
        }
        o.n_rows = o.n_cols = o.n_elem = 0;
    }
};

template <class T>
class SubView {
  public:
    MatT<T> &m;
    uword r1, c1, n_rows, n_cols, n_elem;
    SubView(MatT<T> &mm, long a, long b, long c, long d) : m(mm) {
        if (a < 0 || b < 0 || c < a - 1 || d < b - 1 || c >= (long)mm.n_rows || d >= (long)mm.n_cols)
            shim_bounds("submat / subvec / row / col");
        r1 = (uword)a;
        c1 = (uword)b;
        n_rows = (uword)(c - a + 1);
        n_cols = (uword)(d - b + 1);
        n_elem = n_rows * n_cols;
    }
    SubView(const SubView &) = default;
    T &at(uword i, uword j) const { return m.at(r1 + i, c1 + j); }

    template <class Src>
    void assign(const Src &s) {
        if (s.n_rows != n_rows || s.n_cols != n_cols) throw std::logic_error("copy into submatrix: incompatible matrix dimensions");
        for (uword j = 0; j < n_cols; ++j)
            for (uword i = 0; i < n_rows; ++i) at(i, j) = s.at(i, j);
    }
    SubView &operator=(const MatT<T> &s) {
        assign(s);
        return *this;
    }
    SubView &operator=(const SubView &s) {
        if (&s.m == &m) {  // possible overlap: go through a copy
            MatT<T> tmp(s);
            assign(tmp);
        } else {
            assign(s);
        }
        return *this;
    }
    T min() const { return MatT<T>(*this).min(); }
    T max() const { return MatT<T>(*this).max(); }
    bool has_nan() const { return MatT<T>(*this).has_nan(); }
    MatT<T> t() const { return MatT<T>(*this).t(); }
    T &operator()(long i) const {
        if (i < 0 || (uword)i >= n_elem) shim_bounds("subview::operator()");
        return n_rows == 1 ? at(0, (uword)i) : at((uword)i % n_rows, (uword)i / n_rows);
    }
    T &operator()(long i, long j) const {
        if (i < 0 || j < 0 || (uword)i >= n_rows || (uword)j >= n_cols) shim_bounds("subview::operator()");
        return at((uword)i, (uword)j);
    }
};

template <class T>
MatT<T>::MatT(const SubView<T> &v) {
    init_nofill(v.n_rows, v.n_cols);
    for (uword j = 0; j < n_cols; ++j)
        for (uword i = 0; i < n_rows; ++i) mem_[i + (size_t)j * n_rows] = v.at(i, j);
}

template <class T>
class ColT : public MatT<T> {
  public:
    ColT() { this->n_cols = 1; }
    explicit ColT(long n) : MatT<T>(n, 1) {}
    ColT(long r, long c) : MatT<T>(r, c) { check(); }
    ColT(const MatT<T> &o) : MatT<T>(o) { check(); }
    ColT(MatT<T> &&o) : MatT<T>(std::move(o)) { check(); }
    ColT(const SubView<T> &v) : MatT<T>(v) { check(); }
    ColT(const ColT &o) : MatT<T>(o) {}
    ColT(ColT &&o) noexcept : MatT<T>(std::move(o)) {}
    ColT &operator=(const ColT &o) {
        MatT<T>::operator=(o);
        return *this;
    }
    ColT &operator=(ColT &&o) noexcept {
        MatT<T>::operator=(std::move(o));
        return *this;
    }
    ColT &operator=(const MatT<T> &o) {
        MatT<T>::operator=(o);
        check();
        return *this;
    }
    ColT &operator=(MatT<T> &&o) {
        MatT<T>::operator=(std::move(o));
        check();
        return *this;
    }
    ColT &operator=(const SubView<T> &v) {
        MatT<T>::operator=(v);
        check();
        return *this;
    }
    ColT &operator=(T v) {  // arma: scalar assignment gives a 1-element vector
        this->resize_nofill(1, 1);
        (*this)[0] = v;
        return *this;
    }
    ColT &zeros(long n) {
        MatT<T>::zeros(n, 1);
        return *this;
    }
    ColT &zeros() {
        MatT<T>::zeros();
        return *this;
    }

  private:
    void check() {
        if (this->n_cols != 1) {
            if (this->n_elem == 0) {
                this->n_cols = 1;
                this->n_rows = 0;
            } else {
                throw std::logic_error("Col: incompatible matrix dimensions (not a column vector)");
            }
        }
    }
};

typedef MatT<double> mat;
typedef ColT<double> vec;
typedef ColT<double> colvec;
typedef MatT<sword> imat;
typedef ColT<sword> ivec;

// ---- cube -----------------------------------------------------------------------------------
class cube;
class SliceRef {
  public:
    cube &c;
    uword s;
    SliceRef(cube &cc, uword ss) : c(cc), s(ss) {}
    SliceRef &operator=(const mat &m);
    operator mat() const;
};

class cube {
  public:
    uword n_rows = 0, n_cols = 0, n_slices = 0, n_elem = 0, n_elem_slice = 0;
    std::vector<double> mem;
    cube() {}
    cube(long r, long c, long s) {
        if (r < 0 || c < 0 || s < 0) throw std::logic_error("Cube: negative size");
        n_rows = (uword)r;
        n_cols = (uword)c;
        n_slices = (uword)s;
        n_elem_slice = n_rows * n_cols;
        n_elem = n_elem_slice * n_slices;
        mem.assign(n_elem, 0.0);
    }
    double *memptr() { return mem.data(); }
    const double *memptr() const { return mem.data(); }
    double &operator()(long r, long c, long s) {
        if (r < 0 || c < 0 || s < 0 || (uword)r >= n_rows || (uword)c >= n_cols || (uword)s >= n_slices)
            shim_bounds("Cube::operator()");
        return mem[r + (size_t)c * n_rows + (size_t)s * n_elem_slice];
    }
    SliceRef slice(long s) {
        if (s < 0 || (uword)s >= n_slices) shim_bounds("Cube::slice()");
        return SliceRef(*this, (uword)s);
    }
    mat slice(long s) const {
        if (s < 0 || (uword)s >= n_slices) shim_bounds("Cube::slice()");
        mat m;
        m.init_nofill(n_rows, n_cols);
        if (n_elem_slice) std::memcpy(m.memptr(), mem.data() + (size_t)s * n_elem_slice, sizeof(double) * n_elem_slice);
        return m;
    }
    bool has_nan() const {
        for (double v : mem)
            if (v != v) return true;
        return false;
    }
};
inline SliceRef &SliceRef::operator=(const mat &m) {
    if (m.n_rows != c.n_rows || m.n_cols != c.n_cols) throw std::logic_error("copy into slice: incompatible dimensions");
    if (c.n_elem_slice) std::memcpy(c.mem.data() + (size_t)s * c.n_elem_slice, m.memptr(), sizeof(double) * c.n_elem_slice);
    return *this;
}
inline SliceRef::operator mat() const { return static_cast<const cube &>(c).slice(s); }

// ---- element injector: M << a << b << endr << c << d << endr ----------------------------------
class MatInjector {
  public:
    MatInjector(mat &m, double first) : m_(m) {
        rows_.emplace_back();
        rows_.back().push_back(first);
    }
    MatInjector(const MatInjector &) = delete;
    MatInjector &operator<<(double v) {
        rows_.back().push_back(v);
        return *this;
    }
    MatInjector &operator<<(const endr_t &) {
        rows_.emplace_back();
        return *this;
    }
    ~MatInjector() {
        size_t nr = rows_.size();
        if (nr && rows_.back().empty()) --nr;
        size_t nc = 0;
        for (size_t i = 0; i < nr; ++i) nc = rows_[i].size() > nc ? rows_[i].size() : nc;
        m_.zeros((long)nr, (long)nc);
        for (size_t i = 0; i < nr; ++i)
            for (size_t j = 0; j < rows_[i].size(); ++j) m_.at((uword)i, (uword)j) = rows_[i][j];
    }

  private:
    mat &m_;
    std::vector<std::vector<double>> rows_;
};
// guaranteed copy elision (C++17): the injector lives until the end of the full expression
inline MatInjector operator<<(mat &m, double v) { return MatInjector(m, v); }

// ---- dense arithmetic -------------------------------------------------------------------------
inline void shim_same(const mat &a, const mat &b, const char *op) {
    if (a.n_rows != b.n_rows || a.n_cols != b.n_cols) throw std::logic_error(std::string(op) + ": incompatible matrix dimensions");
}
#define SHIM_EW(NAME, OP)                                                         \
    inline mat NAME(const mat &a, const mat &b) {                                 \
        shim_same(a, b, #OP);                                                     \
        mat r;                                                                    \
        r.init_nofill(a.n_rows, a.n_cols);                                        \
        for (uword i = 0; i < a.n_elem; ++i) r[i] = a[i] OP b[i];                 \
        return r;                                                                 \
    }
SHIM_EW(operator+, +)
SHIM_EW(operator-, -)
SHIM_EW(operator%, *)
#undef SHIM_EW
#define SHIM_SC(NAME, EXPR_MS, EXPR_SM)                     \
    inline mat NAME(const mat &a, double s) {               \
        mat r;                                              \
        r.init_nofill(a.n_rows, a.n_cols);                  \
        for (uword i = 0; i < a.n_elem; ++i) r[i] = EXPR_MS; \
        return r;                                           \
    }                                                       \
    inline mat NAME(double s, const mat &a) {               \
        mat r;                                              \
        r.init_nofill(a.n_rows, a.n_cols);                  \
        for (uword i = 0; i < a.n_elem; ++i) r[i] = EXPR_SM; \
        return r;                                           \
    }
SHIM_SC(operator+, a[i] + s, s + a[i])
SHIM_SC(operator-, a[i] - s, s - a[i])
SHIM_SC(operator/, a[i] / s, s / a[i])
#undef SHIM_SC
inline mat operator*(const mat &a, double s) {
    mat r;
    r.init_nofill(a.n_rows, a.n_cols);
    for (uword i = 0; i < a.n_elem; ++i) r[i] = a[i] * s;
    return r;
}
inline mat operator*(double s, const mat &a) { return a * s; }
inline mat operator-(const mat &a) {
    mat r;
    r.init_nofill(a.n_rows, a.n_cols);
    for (uword i = 0; i < a.n_elem; ++i) r[i] = -a[i];
    return r;
}
inline mat operator*(const mat &a, const mat &b) {
    if (a.n_cols != b.n_rows) {
        if (a.n_elem == 1) return b * a[0];
        if (b.n_elem == 1) return a * b[0];
        throw std::logic_error("matrix multiplication: incompatible matrix dimensions");
    }
    mat r;
    r.init_nofill(a.n_rows, b.n_cols);
    for (uword j = 0; j < b.n_cols; ++j)
        for (uword i = 0; i < a.n_rows; ++i) {
            double acc = 0.0;
            for (uword k = 0; k < a.n_cols; ++k) acc += a.at(i, k) * b.at(k, j);
            r.at(i, j) = acc;
        }
    return r;
}

#define SHIM_MAP(NAME, FN)                                        \
    inline mat NAME(const mat &a) {                               \
        mat r;                                                    \
        r.init_nofill(a.n_rows, a.n_cols);                        \
        for (uword i = 0; i < a.n_elem; ++i) r[i] = FN(a[i]);     \
        return r;                                                 \
    }
SHIM_MAP(exp, std::exp)
SHIM_MAP(log, std::log)
SHIM_MAP(sqrt, std::sqrt)
SHIM_MAP(abs, std::fabs)
#undef SHIM_MAP

// two interleaved accumulators, the order arma's accu()/sum() uses on contiguous memory
inline double accu(const mat &a) {
    double s1 = 0.0, s2 = 0.0;
    uword i = 0;
    for (; i + 1 < a.n_elem; i += 2) {
        s1 += a[i];
        s2 += a[i + 1];
    }
    if (i < a.n_elem) s1 += a[i];
    return s1 + s2;
}
inline double sum(const vec &a) { return accu(a); }
inline double norm(const mat &a, int p = 2) {
    double s = 0.0;
    if (p == 1) {
        for (uword i = 0; i < a.n_elem; ++i) s += std::fabs(a[i]);
        return s;
    }
    for (uword i = 0; i < a.n_elem; ++i) s += a[i] * a[i];
    return std::sqrt(s);
}

inline mat eye(long r, long c) {
    mat m(r, c);
    for (uword i = 0; i < m.n_rows && i < m.n_cols; ++i) m.at(i, i) = 1.0;
    return m;
}
inline vec ones(long n) {
    vec v(n);
    v.ones();
    return v;
}
inline mat ones(long r, long c) {
    mat m(r, c);
    m.ones();
    return m;
}
inline vec zeros(long n) { return vec(n); }
inline mat zeros(long r, long c) { return mat(r, c); }
inline mat diagmat(const mat &a) {
    if (a.n_rows == 1 || a.n_cols == 1) {
        mat m((long)a.n_elem, (long)a.n_elem);
        for (uword i = 0; i < a.n_elem; ++i) m.at(i, i) = a[i];
        return m;
    }
    mat m((long)a.n_rows, (long)a.n_cols);
    for (uword i = 0; i < a.n_rows && i < a.n_cols; ++i) m.at(i, i) = a.at(i, i);
    return m;
}

// upper-triangular R with A = R^T R (column sweep of LAPACK dpotrf 'U'); a non-positive or NaN pivot is a failure
inline mat chol(const mat &A) {
    if (A.n_rows != A.n_cols) throw std::logic_error("chol(): given matrix must be square sized");
    const uword n = A.n_rows;
    mat R((long)n, (long)n);
    for (uword j = 0; j < n; ++j) {
        double d = A.at(j, j);
        for (uword k = 0; k < j; ++k) d -= R.at(k, j) * R.at(k, j);
        if (!(d > 0.0)) throw std::runtime_error("chol(): decomposition failed");
        const double rjj = std::sqrt(d);
        R.at(j, j) = rjj;
        for (uword i = j + 1; i < n; ++i) {
            double s = A.at(j, i);
            for (uword k = 0; k < j; ++k) s -= R.at(k, j) * R.at(k, i);
            R.at(j, i) = s / rjj;
        }
    }
    return R;
}

// partial-pivot LU (dgetrf); returns the sign of the permutation, 0 when exactly singular
inline int shim_lu(mat &a, std::vector<uword> &piv) {
    const uword n = a.n_rows;
    int sgn = 1;
    piv.resize(n);
    for (uword k = 0; k < n; ++k) {
        uword p = k;
        double best = std::fabs(a.at(k, k));
        for (uword i = k + 1; i < n; ++i)
            if (std::fabs(a.at(i, k)) > best) {
                best = std::fabs(a.at(i, k));
                p = i;
            }
        piv[k] = p;
        if (best == 0.0) return 0;
        if (p != k) {
            sgn = -sgn;
            for (uword j = 0; j < n; ++j) std::swap(a.at(k, j), a.at(p, j));
        }
        for (uword i = k + 1; i < n; ++i) {
            a.at(i, k) /= a.at(k, k);
            for (uword j = k + 1; j < n; ++j) a.at(i, j) -= a.at(i, k) * a.at(k, j);
        }
    }
    return sgn;
}
inline double det(const mat &A) {
    if (A.n_rows != A.n_cols) throw std::logic_error("det(): given matrix must be square sized");
    const uword n = A.n_rows;
    if (n == 0) return 1.0;
    if (n == 1) return A[0];
    if (n == 2) return A.at(0, 0) * A.at(1, 1) - A.at(1, 0) * A.at(0, 1);  // arma's tiny-matrix path
    if (n == 3) {
        const double *m = A.memptr();
        return m[0] * (m[4] * m[8] - m[5] * m[7]) - m[1] * (m[3] * m[8] - m[5] * m[6]) + m[2] * (m[3] * m[7] - m[4] * m[6]);
    }
    mat a = A;
    std::vector<uword> piv;
    const int sgn = shim_lu(a, piv);
    if (!sgn) return 0.0;
    double d = sgn;
    for (uword i = 0; i < n; ++i) d *= a.at(i, i);
    return d;
}
inline mat solve(const mat &A, const mat &B) {
    if (A.n_rows != A.n_cols) throw std::logic_error("solve(): only square systems in this shim");
    if (A.n_rows != B.n_rows) throw std::logic_error("solve(): number of rows in A and B must be the same");
    const uword n = A.n_rows;
    mat a = A, x = B;
    std::vector<uword> piv;
    if (!shim_lu(a, piv)) throw std::runtime_error("solve(): solution not found");
    for (uword c = 0; c < x.n_cols; ++c) {
        for (uword k = 0; k < n; ++k)
            if (piv[k] != k) std::swap(x.at(k, c), x.at(piv[k], c));
        for (uword i = 1; i < n; ++i)
            for (uword k = 0; k < i; ++k) x.at(i, c) -= a.at(i, k) * x.at(k, c);
        for (long i = (long)n - 1; i >= 0; --i) {
            for (uword k = (uword)i + 1; k < n; ++k) x.at((uword)i, c) -= a.at((uword)i, k) * x.at(k, c);
            x.at((uword)i, c) /= a.at((uword)i, (uword)i);
        }
    }
    return x;
}
inline mat inv(const mat &A) { return solve(A, eye((long)A.n_rows, (long)A.n_cols)); }

// standard normals from the installed pre-drawn stream, in memory (column-major) order
inline mat randn(long r, long c) {
    mat m;
    m.init_nofill(r, c);
    for (uword i = 0; i < m.n_elem; ++i) m[i] = shim_next_normal();
    return m;
}
inline vec randn(long n) { return vec(randn(n, 1)); }

inline std::ostream &operator<<(std::ostream &os, const mat &m) {
    for (uword i = 0; i < m.n_rows; ++i) {
        for (uword j = 0; j < m.n_cols; ++j) os << (j ? " " : "") << m.at(i, j);
        os << "\n";
    }
    return os;
}

}  // namespace arma

// ------------------------------------------------------------------------------------------------
// Rcpp
// ------------------------------------------------------------------------------------------------
struct SEXPREC;
typedef SEXPREC *SEXP;
static SEXP const R_NilValue = nullptr;

namespace Rcpp {

class List;
struct Value;

class List {
  public:
    struct Slot {
        std::string name;
        std::shared_ptr<const Value> v;
    };
    class Proxy {
      public:
        Proxy(List &l, size_t i) : l_(l), i_(i) {}
        Proxy &operator=(const arma::mat &m);
        Proxy &operator=(const arma::cube &c);
        Proxy &operator=(const List &l);
        Proxy &operator=(double d);
        Proxy &operator=(int i);
        Proxy &operator=(bool b);
        const Value &get() const;

      private:
        List &l_;
        size_t i_;
    };
    List() {}
    explicit List(int n) : slots_((size_t)n) {}
    int size() const { return (int)slots_.size(); }
    int length() const { return size(); }
    Proxy operator[](int i) {
        if (i < 0 || (size_t)i >= slots_.size()) throw std::out_of_range("List: index out of bounds");
        return Proxy(*this, (size_t)i);
    }
    Proxy operator[](const char *name) { return Proxy(*this, find_or_add(name)); }
    Proxy operator[](const std::string &name) { return Proxy(*this, find_or_add(name)); }
    bool containsElementNamed(const char *name) const {
        for (const Slot &s : slots_)
            if (s.name == name) return true;
        return false;
    }
    std::vector<Slot> slots_;

  private:
    size_t find_or_add(const std::string &name) {
        for (size_t i = 0; i < slots_.size(); ++i)
            if (slots_[i].name == name) return i;
        slots_.push_back(Slot{name, nullptr});
        return slots_.size() - 1;
    }
};

struct Value {
    enum Kind { NUM, INT, BOOL, MAT, CUBE, LIST } kind = NUM;
    double num = 0.0;
    arma::mat m;
    arma::cube c;
    List l;
};

inline List::Proxy &List::Proxy::operator=(const arma::mat &m) {
    auto v = std::make_shared<Value>();
    v->kind = Value::MAT;
    v->m = m;
    l_.slots_[i_].v = v;
    return *this;
}
inline List::Proxy &List::Proxy::operator=(const arma::cube &c) {
    auto v = std::make_shared<Value>();
    v->kind = Value::CUBE;
    v->c = c;
    l_.slots_[i_].v = v;
    return *this;
}
inline List::Proxy &List::Proxy::operator=(const List &l) {
    auto v = std::make_shared<Value>();
    v->kind = Value::LIST;
    v->l = l;
    l_.slots_[i_].v = v;
    return *this;
}
inline List::Proxy &List::Proxy::operator=(double d) {
    auto v = std::make_shared<Value>();
    v->kind = Value::NUM;
    v->num = d;
    l_.slots_[i_].v = v;
    return *this;
}
inline List::Proxy &List::Proxy::operator=(int i) {
    auto v = std::make_shared<Value>();
    v->kind = Value::INT;
    v->num = i;
    l_.slots_[i_].v = v;
    return *this;
}
inline List::Proxy &List::Proxy::operator=(bool b) {
    auto v = std::make_shared<Value>();
    v->kind = Value::BOOL;
    v->num = b ? 1.0 : 0.0;
    l_.slots_[i_].v = v;
    return *this;
}
inline const Value &List::Proxy::get() const {
    const std::shared_ptr<const Value> &p = l_.slots_[i_].v;
    if (!p) throw std::runtime_error("List: NULL element");
    return *p;
}

template <class T>
struct AsImpl;
template <>
struct AsImpl<double> {
    static double get(const Value &v) {
        if (v.kind == Value::MAT && v.m.n_elem == 1) return v.m[0];
        if (v.kind > Value::BOOL) throw std::runtime_error("as<double>: not a scalar");
        return v.num;
    }
};
template <>
struct AsImpl<int> {
    static int get(const Value &v) { return (int)AsImpl<double>::get(v); }
};
template <>
struct AsImpl<bool> {
    static bool get(const Value &v) { return AsImpl<double>::get(v) != 0.0; }
};
template <>
struct AsImpl<arma::mat> {
    static arma::mat get(const Value &v) {
        if (v.kind == Value::MAT) return v.m;
        if (v.kind <= Value::BOOL) {
            arma::mat m(1, 1);
            m[0] = v.num;
            return m;
        }
        throw std::runtime_error("as<mat>: not a matrix");
    }
};
template <>
struct AsImpl<arma::vec> {
    static arma::vec get(const Value &v) {
        arma::mat m = AsImpl<arma::mat>::get(v);
        if (m.n_cols != 1) {  // R vectors carry no dim: flatten
            arma::vec r((long)m.n_elem);
            for (arma::uword i = 0; i < m.n_elem; ++i) r[i] = m[i];
            return r;
        }
        return arma::vec(std::move(m));
    }
};
template <>
struct AsImpl<arma::cube> {
    static arma::cube get(const Value &v) {
        if (v.kind != Value::CUBE) throw std::runtime_error("as<cube>: not a cube");
        return v.c;
    }
};
template <>
struct AsImpl<List> {
    static List get(const Value &v) {
        if (v.kind != Value::LIST) throw std::runtime_error("as<List>: not a list");
        return v.l;
    }
};
template <class T>
T as(const List::Proxy &p) {
    return AsImpl<T>::get(p.get());
}

template <class T>
class XPtr {
  public:
    explicit XPtr(T *p) : p_(p) {}
    explicit XPtr(SEXP) {}
    T &operator*() const {
        if (!p_) throw std::runtime_error("external pointer is not valid");
        return *p_;
    }
    T *operator->() const { return &**this; }

  private:
    std::shared_ptr<T> p_;
};

// console output of the reference's diagnostics is dropped (no iostream formatting at all)
struct NullStream {
    template <class T>
    NullStream &operator<<(const T &) {
        return *this;
    }
    NullStream &operator<<(std::ostream &(*)(std::ostream &)) { return *this; }
};
inline NullStream &shim_rcout() {
    static NullStream s;
    return s;
}
#define Rcout shim_rcout()

// Rcpp::stop(msg) throws Rcpp::exception, which END_RCPP turns into an R error
struct exception : std::runtime_error {
    explicit exception(const std::string &m) : std::runtime_error(m) {}
};
[[noreturn]] inline void stop(const std::string &m) { throw exception(m); }

// Rcpp modules (RCPP_MODULE body of src/generalLNA.cpp): registration is a no-op here, the class is used directly
template <class T>
struct class_ {
    explicit class_(const char *) {}
    template <class... A> class_ &constructor() { return *this; }
    template <class M> class_ &field(const char *, M) { return *this; }
    template <class G> class_ &property(const char *, G) { return *this; }
    template <class G, class S> class_ &property(const char *, G, S) { return *this; }
    template <class M> class_ &method(const char *, M) { return *this; }
};

}  // namespace Rcpp

namespace R {
// R nmath runif(a, b) = a + (b - a) * unif_rand(); the uniform comes from the installed stream
inline double runif(double a, double b) {
    const double u = shim_next_uniform();
    if (a == b) return a;
    return a + (b - a) * u;
}
}  // namespace R

#endif
