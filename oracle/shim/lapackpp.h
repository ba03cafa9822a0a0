// Stand-in for LAPACK++ 2.5.4 (lapackpp.h), which the reference downloads at
// build time (/root/reference/CMakeLists.txt:86-96) and which is NOT present in
// this image.  TEST INFRASTRUCTURE ONLY: it exists so that the reference's own
// aku/*.cc sources compile, unmodified and from where they lie, into the
// oracle binaries under oracle/_ref/.  Nothing in the product links it.
//
// Scope: the symbol surface aku/ uses.  Dense column-major
// double storage with shared-handle views, naive loops instead of BLAS/LAPACK.
// The diagonal-Gaussian scoring path (aku/Distributions.cc:1041-1062) executes
// none of this arithmetic -- it only uses operator() for element access -- so
// that path is faithful to a real LapackPP build up to libm.
#ifndef ORACLE_SHIM_LAPACKPP_H
#define ORACLE_SHIM_LAPACKPP_H

#include <assert.h>
#include <math.h>
#include <stdlib.h>
#include <stdio.h>
#include <iostream>
#include <memory>
#include <vector>
#include <complex>

struct LaIndex {
  int lo, hi;
  LaIndex() : lo(0), hi(-1) {}
  LaIndex(int a, int b) : lo(a), hi(b) {}
  int start() const { return lo; }
  int end() const { return hi; }
};

class LaGenMatDouble {
public:
  typedef std::shared_ptr<std::vector<double> > Store;
  Store st;
  int off, nr, nc, ld, rinc;  // element (i,j) at off + i*rinc + j*ld

  LaGenMatDouble() : st(new std::vector<double>()), off(0), nr(0), nc(0), ld(0), rinc(1) {}
  LaGenMatDouble(int r, int c) { alloc(r, c); }
  LaGenMatDouble(const LaGenMatDouble &o) { init_from(o); }
  virtual ~LaGenMatDouble() {}

  void alloc(int r, int c) {
    st.reset(new std::vector<double>((size_t)(r > 0 ? r : 0) * (c > 0 ? c : 0), 0.0));
    off = 0; nr = r; nc = c; ld = r; rinc = 1; view = false;
  }
  void init_from(const LaGenMatDouble &o) {
    if (o.view) { st = o.st; off = o.off; nr = o.nr; nc = o.nc; ld = o.ld; rinc = o.rinc; view = false; }
    else { alloc(o.nr, o.nc); assign_elems(o); }
  }
  void assign_elems(const LaGenMatDouble &o) {
    for (int j = 0; j < nc; j++) for (int i = 0; i < nr; i++) (*this)(i, j) = o(i, j);
  }
  int rows() const { return nr; }
  int cols() const { return nc; }
  int size(int d) const { return d == 0 ? nr : nc; }
  int inc(int d) const { return d == 0 ? rinc : 1; }
  int gdim(int d) const { return d == 0 ? ld : nc; }
  double *addr() const { return st->data() + off; }
  double &operator()(int i, int j) { return (*st)[off + (size_t)i * rinc + (size_t)j * ld]; }
  const double &operator()(int i, int j) const { return (*st)[off + (size_t)i * rinc + (size_t)j * ld]; }
  LaGenMatDouble operator()(const LaIndex &ri, const LaIndex &ci) const {
    LaGenMatDouble v; v.st = st; v.off = off + ri.lo * rinc + ci.lo * ld;
    v.nr = ri.hi - ri.lo + 1; v.nc = ci.hi - ci.lo + 1; v.ld = ld; v.rinc = rinc; v.view = true; return v;
  }
  LaGenMatDouble row(int k) const { return (*this)(LaIndex(k, k), LaIndex(0, nc - 1)); }
  LaGenMatDouble col(int k) const { return (*this)(LaIndex(0, nr - 1), LaIndex(k, k)); }
  LaGenMatDouble &resize(int r, int c) { alloc(r, c); return *this; }
  LaGenMatDouble &resize(const LaGenMatDouble &o) { alloc(o.nr, o.nc); return *this; }
  LaGenMatDouble &operator=(double s) {
    for (int j = 0; j < nc; j++) for (int i = 0; i < nr; i++) (*this)(i, j) = s; return *this;
  }
  LaGenMatDouble &operator=(const LaGenMatDouble &o) { return copy(o); }
  LaGenMatDouble &copy(const LaGenMatDouble &o) {
    if (this == &o) return *this;
    if (nr != o.nr || nc != o.nc) alloc(o.nr, o.nc);
    // copy through a temporary in case of aliasing views
    std::vector<double> tmp((size_t)nr * nc);
    for (int j = 0; j < nc; j++) for (int i = 0; i < nr; i++) tmp[(size_t)j * nr + i] = o(i, j);
    for (int j = 0; j < nc; j++) for (int i = 0; i < nr; i++) (*this)(i, j) = tmp[(size_t)j * nr + i];
    return *this;
  }
  LaGenMatDouble copy() const { LaGenMatDouble r(nr, nc); r.assign_elems(*this); return r; }
  LaGenMatDouble &ref(const LaGenMatDouble &o) {
    st = o.st; off = o.off; nr = o.nr; nc = o.nc; ld = o.ld; rinc = o.rinc; return *this;
  }
  LaGenMatDouble &inject(const LaGenMatDouble &o) {
    assert(nr == o.nr && nc == o.nc); assign_elems(o); return *this;
  }
  LaGenMatDouble &shallow_assign() { view = true; return *this; }
  LaGenMatDouble operator+(const LaGenMatDouble &o) const {
    LaGenMatDouble r(nr, nc);
    for (int j = 0; j < nc; j++) for (int i = 0; i < nr; i++) r(i, j) = (*this)(i, j) + o(i, j);
    return r;
  }
  double trace() const { double t = 0; for (int i = 0; i < nr && i < nc; i++) t += (*this)(i, i); return t; }
  void scale(double s) { for (int j = 0; j < nc; j++) for (int i = 0; i < nr; i++) (*this)(i, j) *= s; }
  static LaGenMatDouble zeros(int r, int c = 0) { return LaGenMatDouble(r, c ? c : r); }
  static LaGenMatDouble eye(int r, int c = 0) {
    LaGenMatDouble m(r, c ? c : r); for (int i = 0; i < m.nr && i < m.nc; i++) m(i, i) = 1; return m;
  }
  static LaGenMatDouble ones(int r, int c = 0) { LaGenMatDouble m(r, c ? c : r); m = 1.0; return m; }
protected:
  bool view = false;
};

class LaVectorDouble : public LaGenMatDouble {
public:
  LaVectorDouble() : LaGenMatDouble() {}
  LaVectorDouble(int n) : LaGenMatDouble(n, 1) {}
  LaVectorDouble(int r, int c) : LaGenMatDouble(r, c) {}
  LaVectorDouble(const LaGenMatDouble &o) : LaGenMatDouble(o) {}
  LaVectorDouble(const LaVectorDouble &o) : LaGenMatDouble(o) {}
  int size() const { return nr * nc; }
  using LaGenMatDouble::size;
  int inc() const { return nc == 1 ? rinc : ld; }
  double &operator()(int i) { return nc == 1 ? LaGenMatDouble::operator()(i, 0) : LaGenMatDouble::operator()(0, i); }
  const double &operator()(int i) const { return nc == 1 ? LaGenMatDouble::operator()(i, 0) : LaGenMatDouble::operator()(0, i); }
  using LaGenMatDouble::operator();
  LaVectorDouble &resize(int n) { alloc(n, 1); return *this; }
  LaVectorDouble &resize(int r, int c) { alloc(r, c); return *this; }
  LaVectorDouble &operator=(double s) { LaGenMatDouble::operator=(s); return *this; }
  LaVectorDouble &operator=(const LaGenMatDouble &o) { LaGenMatDouble::copy(o); return *this; }
  LaVectorDouble &operator=(const LaVectorDouble &o) { LaGenMatDouble::copy(o); return *this; }
  LaVectorDouble &copy(const LaGenMatDouble &o) { LaGenMatDouble::copy(o); return *this; }
  LaVectorDouble &ref(const LaGenMatDouble &o) { LaGenMatDouble::ref(o); return *this; }
  LaVectorDouble &inject(const LaGenMatDouble &o) { LaGenMatDouble::inject(o); return *this; }
};

class LaSymmMatDouble {
public:
  LaGenMatDouble m;
  LaSymmMatDouble() {}
  LaSymmMatDouble(int r, int c) : m(r, c) {}
  LaSymmMatDouble &resize(int r, int c) { m.resize(r, c); return *this; }
  LaSymmMatDouble &operator=(double s) { m = s; return *this; }
  int rows() const { return m.rows(); }
  int cols() const { return m.cols(); }
  double &operator()(int i, int j) { return i >= j ? m(i, j) : m(j, i); }
  const double &operator()(int i, int j) const { return i >= j ? m(i, j) : m(j, i); }
  int size(int d) const { return m.size(d); }
  operator LaGenMatDouble() const {
    LaGenMatDouble r(m.rows(), m.cols());
    for (int j = 0; j < m.cols(); j++) for (int i = 0; i < m.rows(); i++) r(i, j) = (*this)(i, j);
    return r;
  }
};

class LaVectorLongInt {
public:
  std::vector<long> v;
  LaVectorLongInt() {}
  LaVectorLongInt(int n) : v(n, 0) {}
  LaVectorLongInt(int r, int c) : v((size_t)r * c, 0) {}
  int size() const { return (int)v.size(); }
  void resize(int r, int c = 1) { v.assign((size_t)r * c, 0); }
  long &operator()(int i) { return v[i]; }
  const long &operator()(int i) const { return v[i]; }
};

struct LaComplex {
  double r, i;
  LaComplex() : r(0), i(0) {}
  LaComplex(double re) : r(re), i(0) {}
  LaComplex(double re, double im) : r(re), i(im) {}
};
class LaGenMatComplex {
public:
  int nr, nc; std::vector<LaComplex> d;
  LaGenMatComplex() : nr(0), nc(0) {}
  LaGenMatComplex(int r, int c) : nr(r), nc(c), d((size_t)r * c) {}
  LaGenMatComplex(const LaGenMatDouble &o) : nr(o.rows()), nc(o.cols()), d((size_t)o.rows() * o.cols()) {
    for (int j = 0; j < nc; j++) for (int i = 0; i < nr; i++) d[(size_t)j * nr + i] = o(i, j);
  }
  int rows() const { return nr; }
  int cols() const { return nc; }
  int size(int k) const { return k == 0 ? nr : nc; }
  LaComplex &operator()(int i, int j) { return d[(size_t)j * nr + i]; }
  const LaComplex &operator()(int i, int j) const { return d[(size_t)j * nr + i]; }
  void resize(int r, int c) { nr = r; nc = c; d.assign((size_t)r * c, LaComplex()); }
};
class LaVectorComplex : public LaGenMatComplex {
public:
  LaVectorComplex() {}
  LaVectorComplex(int n) : LaGenMatComplex(n, 1) {}
  int size() const { return nr * nc; }
  using LaGenMatComplex::size;
  LaComplex &operator()(int i) { return d[i]; }
  const LaComplex &operator()(int i) const { return d[i]; }
  void resize(int n, int c = 1) { LaGenMatComplex::resize(n, c); }
};

inline std::ostream &operator<<(std::ostream &os, const LaGenMatDouble &m) {
  for (int i = 0; i < m.rows(); i++) { for (int j = 0; j < m.cols(); j++) os << m(i, j) << " "; os << "\n"; }
  return os;
}

// ---- BLAS-like helpers (naive loops) ---------------------------------------
inline void Blas_Scale(double a, LaGenMatDouble &x) { x.scale(a); }
inline void Blas_Add_Mult(LaVectorDouble &y, double a, const LaVectorDouble &x) {
  for (int i = 0; i < y.size(); i++) y(i) += a * x(i);
}
inline void Blas_Mult(LaVectorDouble &y, double a, const LaVectorDouble &x) {
  for (int i = 0; i < y.size(); i++) y(i) = a * x(i);
}
inline double Blas_Dot_Prod(const LaVectorDouble &x, const LaVectorDouble &y) {
  double s = 0; for (int i = 0; i < x.size(); i++) s += x(i) * y(i); return s;
}
inline double Blas_Norm2(const LaVectorDouble &x) { return sqrt(Blas_Dot_Prod(x, x)); }
inline void Blas_Mat_Vec_Mult(const LaGenMatDouble &A, const LaVectorDouble &x, LaVectorDouble &y,
                              double alpha = 1.0, double beta = 0.0) {
  std::vector<double> t(A.rows());
  for (int i = 0; i < A.rows(); i++) { double s = 0; for (int j = 0; j < A.cols(); j++) s += A(i, j) * x(j); t[i] = s; }
  for (int i = 0; i < A.rows(); i++) y(i) = alpha * t[i] + (beta == 0.0 ? 0.0 : beta * y(i));
}
inline void Blas_Mat_Trans_Vec_Mult(const LaGenMatDouble &A, const LaVectorDouble &x, LaVectorDouble &y,
                                    double alpha = 1.0, double beta = 0.0) {
  std::vector<double> t(A.cols());
  for (int j = 0; j < A.cols(); j++) { double s = 0; for (int i = 0; i < A.rows(); i++) s += A(i, j) * x(i); t[j] = s; }
  for (int j = 0; j < A.cols(); j++) y(j) = alpha * t[j] + (beta == 0.0 ? 0.0 : beta * y(j));
}
inline void Blas_R1_Update(LaGenMatDouble &A, const LaVectorDouble &x, const LaVectorDouble &y, double alpha = 1.0) {
  for (int j = 0; j < A.cols(); j++) for (int i = 0; i < A.rows(); i++) A(i, j) += alpha * x(i) * y(j);
}
inline void Blas_R1_Update(LaSymmMatDouble &A, const LaVectorDouble &x, double alpha = 1.0) {
  for (int j = 0; j < A.size(1); j++) for (int i = j; i < A.size(0); i++) A(i, j) += alpha * x(i) * x(j);
}
inline void shim_gemm(const LaGenMatDouble &A, bool ta, const LaGenMatDouble &B, bool tb, LaGenMatDouble &C,
                      double alpha, double beta) {
  int m = ta ? A.cols() : A.rows(), k = ta ? A.rows() : A.cols(), n = tb ? B.rows() : B.cols();
  std::vector<double> t((size_t)m * n);
  for (int j = 0; j < n; j++) for (int i = 0; i < m; i++) {
    double s = 0;
    for (int p = 0; p < k; p++) s += (ta ? A(p, i) : A(i, p)) * (tb ? B(j, p) : B(p, j));
    t[(size_t)j * m + i] = s;
  }
  if (C.rows() != m || C.cols() != n) { assert(beta == 0.0); C.resize(m, n); }
  for (int j = 0; j < n; j++) for (int i = 0; i < m; i++)
    C(i, j) = alpha * t[(size_t)j * m + i] + (beta == 0.0 ? 0.0 : beta * C(i, j));
}
inline void Blas_Mat_Mat_Mult(const LaGenMatDouble &A, const LaGenMatDouble &B, LaGenMatDouble &C,
                              double alpha = 1.0, double beta = 0.0) { shim_gemm(A, false, B, false, C, alpha, beta); }
inline void Blas_Mat_Mat_Mult(const LaGenMatDouble &A, const LaGenMatDouble &B, LaGenMatDouble &C,
                              bool ta, bool tb, double alpha = 1.0, double beta = 0.0) { shim_gemm(A, ta, B, tb, C, alpha, beta); }
inline void Blas_Mat_Mat_Trans_Mult(const LaGenMatDouble &A, const LaGenMatDouble &B, LaGenMatDouble &C,
                                    double alpha = 1.0, double beta = 0.0) { shim_gemm(A, false, B, true, C, alpha, beta); }
inline void Blas_Mat_Trans_Mat_Mult(const LaGenMatDouble &A, const LaGenMatDouble &B, LaGenMatDouble &C,
                                    double alpha = 1.0, double beta = 0.0) { shim_gemm(A, true, B, false, C, alpha, beta); }
// symmetric rank-k update (blas3pp.h): C := alpha*A*A' + beta*C (or A'*A when right_transposed is false)
inline void Blas_R1_Update(LaSymmMatDouble &C, const LaGenMatDouble &A, double alpha = 1.0, double beta = 1.0,
                           bool right_transposed = true) {
  int n = C.size(0), k = right_transposed ? A.cols() : A.rows();
  for (int j = 0; j < n; j++) for (int i = j; i < n; i++) {
    double s = 0;
    for (int p = 0; p < k; p++) s += right_transposed ? A(i, p) * A(j, p) : A(p, i) * A(p, j);
    C(i, j) = alpha * s + beta * C(i, j);
  }
}
// rank-k update: C := alpha*A*A' + beta*C
inline void Blas_R1_Update(LaGenMatDouble &C, const LaGenMatDouble &A, double alpha = 1.0, double beta = 1.0) {
  shim_gemm(A, false, A, true, C, alpha, beta);
}
// added by vendor/lapackpp-2.5.4.ics.patch: A += alpha*B
inline void Blas_Add_Mat_Mult(LaGenMatDouble &A, double alpha, const LaGenMatDouble &B) {
  for (int j = 0; j < A.cols(); j++) for (int i = 0; i < A.rows(); i++) A(i, j) += alpha * B(i, j);
}

// ---- LAPACK-like helpers -----------------------------------------------------
// LU with partial pivoting, in place; pivots 1-based like LAPACK dgetrf.
inline void LUFactorizeIP(LaGenMatDouble &A, LaVectorLongInt &piv) {
  int n = A.rows();
  for (int k = 0; k < n; k++) {
    int p = k; double best = fabs(A(k, k));
    for (int i = k + 1; i < n; i++) if (fabs(A(i, k)) > best) { best = fabs(A(i, k)); p = i; }
    piv(k) = p + 1;
    if (p != k) for (int j = 0; j < n; j++) { double t = A(k, j); A(k, j) = A(p, j); A(p, j) = t; }
    if (A(k, k) != 0.0)
      for (int i = k + 1; i < n; i++) {
        A(i, k) /= A(k, k);
        for (int j = k + 1; j < n; j++) A(i, j) -= A(i, k) * A(k, j);
      }
  }
}
inline void LaLUInverseIP(LaGenMatDouble &A, LaVectorLongInt &piv) {
  int n = A.rows();
  LaGenMatDouble inv(n, n);
  for (int c = 0; c < n; c++) {
    std::vector<double> b(n, 0.0); b[c] = 1.0;
    for (int k = 0; k < n; k++) { int p = (int)piv(k) - 1; if (p != k) { double t = b[k]; b[k] = b[p]; b[p] = t; } }
    for (int i = 0; i < n; i++) for (int j = 0; j < i; j++) b[i] -= A(i, j) * b[j];
    for (int i = n - 1; i >= 0; i--) { for (int j = i + 1; j < n; j++) b[i] -= A(i, j) * b[j]; b[i] /= A(i, i); }
    for (int i = 0; i < n; i++) inv(i, c) = b[i];
  }
  A.copy(inv);
}
inline void LaLUInverseIP(LaGenMatDouble &A, LaVectorLongInt &piv, LaVectorDouble &) { LaLUInverseIP(A, piv); }
// Symmetric eigen-decomposition by cyclic Jacobi; eigenvalues ascending, vectors in columns of A.
inline void LaEigSolveSymmetricVecIP(LaGenMatDouble &A, LaVectorDouble &w) {
  int n = A.rows();
  LaGenMatDouble V = LaGenMatDouble::eye(n);
  for (int i = 0; i < n; i++) for (int j = 0; j < i; j++) A(j, i) = A(i, j);  // uses lower triangle
  for (int sweep = 0; sweep < 100; sweep++) {
    double offd = 0; for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) if (i != j) offd += A(i, j) * A(i, j);
    if (offd < 1e-300) break;
    for (int p = 0; p < n; p++) for (int q = p + 1; q < n; q++) {
      if (A(p, q) == 0.0) continue;
      double th = (A(q, q) - A(p, p)) / (2 * A(p, q));
      double t = (th >= 0 ? 1.0 : -1.0) / (fabs(th) + sqrt(th * th + 1));
      double c = 1 / sqrt(t * t + 1), s = t * c;
      for (int k = 0; k < n; k++) { double a = A(k, p), b = A(k, q); A(k, p) = c * a - s * b; A(k, q) = s * a + c * b; }
      for (int k = 0; k < n; k++) { double a = A(p, k), b = A(q, k); A(p, k) = c * a - s * b; A(q, k) = s * a + c * b; }
      for (int k = 0; k < n; k++) { double a = V(k, p), b = V(k, q); V(k, p) = c * a - s * b; V(k, q) = s * a + c * b; }
    }
  }
  if (w.size() != n) w.resize(n, 1);
  std::vector<int> ord(n); for (int i = 0; i < n; i++) ord[i] = i;
  for (int i = 0; i < n; i++) for (int j = i + 1; j < n; j++) if (A(ord[j], ord[j]) < A(ord[i], ord[i])) { int t = ord[i]; ord[i] = ord[j]; ord[j] = t; }
  LaGenMatDouble R(n, n);
  for (int c = 0; c < n; c++) { w(c) = A(ord[c], ord[c]); for (int k = 0; k < n; k++) R(k, c) = V(k, ord[c]); }
  A.copy(R);
}
inline void LaEigSolve(const LaGenMatComplex &, LaVectorComplex &, LaGenMatComplex &) {
  fprintf(stderr, "oracle shim: LaEigSolve (general complex eigenproblem) is training-only and not provided\n");
  abort();
}

#endif
