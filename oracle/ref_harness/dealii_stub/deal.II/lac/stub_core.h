/*
 * Minimal stand-in for the handful of deal.II linear-algebra types that the reference's
 * augmented_lagrangian_preconditioner.h touches.  TEST INFRASTRUCTURE (see oracle/ref_prec.py):
 * written from deal.II's documented public semantics, not from its sources, and just large enough
 * that the UNMODIFIED reference header compiles from where it lies (/root/reference) so that its
 * five vmult()s can be executed as the pinning oracle for the block-preconditioner algebra.
 *
 * Semantics kept (deal.II 9.6 documentation of LinearOperator / PackagedOperation):
 *   Vector<Number>           dense vector; `v = s` sets every entry; reinit(n) / reinit(other)
 *   BlockVector<Number>      blocks of Vector; reinit(n_blocks); block(i)
 *   LinearOperator<R,D>      std::function members vmult / vmult_add / Tvmult / Tvmult_add and
 *                            reinit_range_vector / reinit_domain_vector
 *   s * op, op * op          scaled operator, composition (intermediate sized by the inner
 *                            operator's reinit_range_vector)
 *   op * v, op * pkg         PackagedOperation: lazily evaluated, converts to Range through a
 *                            freshly reinit-ed temporary (so `x = op * f(x)` never aliases)
 *   v +/- pkg, pkg +/- pkg, v +/- v, s * pkg
 *   transpose_operator(op)   swaps vmult <-> Tvmult
 */
#ifndef FDAL_DEALII_STUB_CORE_H
#define FDAL_DEALII_STUB_CORE_H
#include <cassert>
#include <cstddef>
#include <functional>
#include <stdexcept>
#include <string>
#include <vector>

namespace dealii {

/* exceptions: deal.II's Assert / AssertThrow throw the exception object they are handed */
struct ExcMessage : std::runtime_error {
  explicit ExcMessage(const std::string &m) : std::runtime_error(m) {}
};
inline ExcMessage ExcDimensionMismatch(const std::size_t a, const std::size_t b) {
  return ExcMessage("dimension mismatch: " + std::to_string(a) + " != " + std::to_string(b));
}
inline ExcMessage ExcNotInitialized() { return ExcMessage("not initialized"); }
#define Assert(cond, exc) \
  do {                    \
    if (!(cond)) throw exc; \
  } while (0)
#define AssertThrow(cond, exc) Assert(cond, exc)

/* SolverControl family: only what an adapter reads (limits) and throws (NoConvergence) */
class SolverControl {
 public:
  class NoConvergence : public std::exception {
   public:
    NoConvergence(const unsigned int step, const double residual) : last_step(step), last_residual(residual) {}
    const char *what() const noexcept override { return "SolverControl::NoConvergence"; }
    const unsigned int last_step;
    const double last_residual;
  };
  explicit SolverControl(const unsigned int n = 100, const double tol = 1.e-10) : maxsteps(n), tol_(tol) {}
  virtual ~SolverControl() = default;
  unsigned int max_steps() const { return maxsteps; }
  double tolerance() const { return tol_; }

 private:
  unsigned int maxsteps;
  double tol_;
};
class ReductionControl : public SolverControl {
 public:
  explicit ReductionControl(const unsigned int n = 100, const double tol = 1.e-10, const double red = 1.e-2)
      : SolverControl(n, tol), reduce(red) {}
  double reduction() const { return reduce; }

 private:
  double reduce;
};
class IterationNumberControl : public SolverControl {
 public:
  explicit IterationNumberControl(const unsigned int n = 100, const double tol = 1.e-12) : SolverControl(n, tol) {}
};

template <typename Range>
class PackagedOperation;

template <typename Number>
class Vector {
 public:
  using value_type = Number;
  using size_type = std::size_t;
  Vector() = default;
  explicit Vector(const size_type n) : data(n, Number(0)) {}
  size_type size() const { return data.size(); }
  void reinit(const size_type n, const bool omit_zeroing = false) {
    data.resize(n);
    if (!omit_zeroing) std::fill(data.begin(), data.end(), Number(0));
  }
  template <typename N2>
  void reinit(const Vector<N2> &other, const bool omit_zeroing = false) {
    reinit(other.size(), omit_zeroing);
  }
  Vector &operator=(const Number s) {
    std::fill(data.begin(), data.end(), s);
    return *this;
  }
  Vector &operator=(const Vector &) = default;
  Vector(const Vector &) = default;
  Number &operator[](const size_type i) { return data[i]; }
  const Number &operator[](const size_type i) const { return data[i]; }
  Number &operator()(const size_type i) { return data[i]; }
  const Number &operator()(const size_type i) const { return data[i]; }
  Number *begin() { return data.data(); }
  const Number *begin() const { return data.data(); }
  Number *end() { return data.data() + data.size(); }
  const Number *end() const { return data.data() + data.size(); }
  Vector &operator+=(const Vector &o) {
    assert(o.size() == size());
    for (size_type i = 0; i < size(); ++i) data[i] += o.data[i];
    return *this;
  }
  Vector &operator-=(const Vector &o) {
    assert(o.size() == size());
    for (size_type i = 0; i < size(); ++i) data[i] -= o.data[i];
    return *this;
  }
  Vector &operator*=(const Number s) {
    for (auto &x : data) x *= s;
    return *this;
  }
  void add(const Number a, const Vector &o) {
    assert(o.size() == size());
    for (size_type i = 0; i < size(); ++i) data[i] += a * o.data[i];
  }

 private:
  std::vector<Number> data;
};

template <typename Number>
class BlockVector {
 public:
  using value_type = Number;
  using BlockType = Vector<Number>;
  using size_type = std::size_t;
  BlockVector() = default;
  explicit BlockVector(const unsigned int n_blocks, const size_type block_size = 0) { reinit(n_blocks, block_size); }
  explicit BlockVector(const std::vector<size_type> &sizes) {
    blocks.resize(sizes.size());
    for (std::size_t b = 0; b < sizes.size(); ++b) blocks[b].reinit(sizes[b]);
  }
  void reinit(const unsigned int n_blocks, const size_type block_size = 0, const bool omit_zeroing = false) {
    blocks.resize(n_blocks);
    for (auto &b : blocks) b.reinit(block_size, omit_zeroing);
  }
  void reinit(const BlockVector &other, const bool omit_zeroing = false) {
    blocks.resize(other.n_blocks());
    for (unsigned int b = 0; b < other.n_blocks(); ++b) blocks[b].reinit(other.block(b), omit_zeroing);
  }
  void collect_sizes() {}
  unsigned int n_blocks() const { return static_cast<unsigned int>(blocks.size()); }
  size_type size() const {
    size_type s = 0;
    for (const auto &b : blocks) s += b.size();
    return s;
  }
  BlockType &block(const unsigned int i) { return blocks.at(i); }
  const BlockType &block(const unsigned int i) const { return blocks.at(i); }
  BlockVector &operator=(const Number s) {
    for (auto &b : blocks) b = s;
    return *this;
  }
  BlockVector &operator=(const BlockVector &) = default;
  BlockVector(const BlockVector &) = default;
  BlockVector &operator+=(const BlockVector &o) {
    for (unsigned int b = 0; b < n_blocks(); ++b) blocks[b] += o.blocks[b];
    return *this;
  }
  BlockVector &operator-=(const BlockVector &o) {
    for (unsigned int b = 0; b < n_blocks(); ++b) blocks[b] -= o.blocks[b];
    return *this;
  }
  BlockVector &operator*=(const Number s) {
    for (auto &b : blocks) b *= s;
    return *this;
  }

 private:
  std::vector<BlockType> blocks;
};

/* SparseMatrix<double>: CSR with deal.II's accessor-style row iterators (column(), value()) */
template <typename Number>
class SparseMatrix {
 public:
  using size_type = std::size_t;
  struct Accessor {
    const SparseMatrix *A;
    size_type k;
    size_type column() const { return A->colnums[k]; }
    Number value() const { return A->val[k]; }
  };
  class const_iterator {
   public:
    const_iterator(const SparseMatrix *A, size_type k) : acc{A, k} {}
    const Accessor *operator->() const { return &acc; }
    const Accessor &operator*() const { return acc; }
    const_iterator &operator++() {
      ++acc.k;
      return *this;
    }
    bool operator!=(const const_iterator &o) const { return acc.k != o.acc.k; }
    bool operator==(const const_iterator &o) const { return acc.k == o.acc.k; }

   private:
    Accessor acc;
  };
  SparseMatrix() = default;
  SparseMatrix(size_type rows, size_type cols, std::vector<size_type> rowstart_, std::vector<unsigned int> colnums_,
               std::vector<Number> val_)
      : n_rows(rows), n_cols(cols), rowstart(std::move(rowstart_)), colnums(std::move(colnums_)), val(std::move(val_)) {}
  size_type m() const { return n_rows; }
  size_type n() const { return n_cols; }
  size_type n_nonzero_elements() const { return val.size(); }
  const_iterator begin(const size_type r) const { return const_iterator(this, rowstart[r]); }
  const_iterator end(const size_type r) const { return const_iterator(this, rowstart[r + 1]); }
  void vmult(Vector<Number> &y, const Vector<Number> &x) const {
    for (size_type r = 0; r < n_rows; ++r) {
      Number s = 0;
      for (size_type k = rowstart[r]; k < rowstart[r + 1]; ++k) s += val[k] * x[colnums[k]];
      y[r] = s;
    }
  }

 private:
  size_type n_rows = 0, n_cols = 0;
  std::vector<size_type> rowstart;
  std::vector<unsigned int> colnums;
  std::vector<Number> val;
};

/* ------------------------------------------------------------------ LinearOperator */
template <typename Range, typename Domain = Range>
class LinearOperator {
 public:
  std::function<void(Range &, const Domain &)> vmult, vmult_add;
  std::function<void(Domain &, const Range &)> Tvmult, Tvmult_add;
  std::function<void(Range &, bool)> reinit_range_vector;
  std::function<void(Domain &, bool)> reinit_domain_vector;
  bool is_null_operator = false;

  LinearOperator() {
    vmult = vmult_add = [](Range &, const Domain &) { throw std::runtime_error("uninitialised LinearOperator::vmult"); };
    Tvmult = Tvmult_add = [](Domain &, const Range &) {
      throw std::runtime_error("uninitialised LinearOperator::Tvmult");
    };
    reinit_range_vector = [](Range &, bool) { throw std::runtime_error("uninitialised reinit_range_vector"); };
    reinit_domain_vector = [](Domain &, bool) { throw std::runtime_error("uninitialised reinit_domain_vector"); };
  }
};

template <typename Range, typename Domain>
LinearOperator<Domain, Range> transpose_operator(const LinearOperator<Range, Domain> &op) {
  LinearOperator<Domain, Range> t;
  t.vmult = op.Tvmult;
  t.vmult_add = op.Tvmult_add;
  t.Tvmult = op.vmult;
  t.Tvmult_add = op.vmult_add;
  t.reinit_range_vector = op.reinit_domain_vector;
  t.reinit_domain_vector = op.reinit_range_vector;
  return t;
}

/* s * op */
template <typename Range, typename Domain>
LinearOperator<Range, Domain> operator*(const typename Range::value_type s, const LinearOperator<Range, Domain> &op) {
  LinearOperator<Range, Domain> r = op;
  r.vmult = [s, op](Range &y, const Domain &x) {
    op.vmult(y, x);
    y *= s;
  };
  r.vmult_add = [s, op](Range &y, const Domain &x) {
    Range t;
    op.reinit_range_vector(t, true);
    op.vmult(t, x);
    t *= s;
    y += t;
  };
  r.Tvmult = [s, op](Domain &y, const Range &x) {
    op.Tvmult(y, x);
    y *= s;
  };
  r.Tvmult_add = [s, op](Domain &y, const Range &x) {
    Domain t;
    op.reinit_domain_vector(t, true);
    op.Tvmult(t, x);
    t *= s;
    y += t;
  };
  return r;
}
template <typename Range, typename Domain>
LinearOperator<Range, Domain> operator*(const LinearOperator<Range, Domain> &op, const typename Range::value_type s) {
  return s * op;
}

/* first * second : Domain -> Intermediate -> Range */
template <typename Range, typename Intermediate, typename Domain>
LinearOperator<Range, Domain> operator*(const LinearOperator<Range, Intermediate> &first,
                                        const LinearOperator<Intermediate, Domain> &second) {
  LinearOperator<Range, Domain> r;
  r.reinit_range_vector = first.reinit_range_vector;
  r.reinit_domain_vector = second.reinit_domain_vector;
  r.vmult = [first, second](Range &y, const Domain &x) {
    Intermediate t;
    second.reinit_range_vector(t, true);
    second.vmult(t, x);
    first.vmult(y, t);
  };
  r.vmult_add = [first, second](Range &y, const Domain &x) {
    Intermediate t;
    second.reinit_range_vector(t, true);
    second.vmult(t, x);
    first.vmult_add(y, t);
  };
  r.Tvmult = [first, second](Domain &y, const Range &x) {
    Intermediate t;
    first.reinit_domain_vector(t, true);
    first.Tvmult(t, x);
    second.Tvmult(y, t);
  };
  r.Tvmult_add = [first, second](Domain &y, const Range &x) {
    Intermediate t;
    first.reinit_domain_vector(t, true);
    first.Tvmult(t, x);
    second.Tvmult_add(y, t);
  };
  return r;
}

/* ------------------------------------------------------------------ PackagedOperation */
template <typename Range>
class PackagedOperation {
 public:
  std::function<void(Range &)> apply, apply_add;
  std::function<void(Range &, bool)> reinit_vector;

  PackagedOperation() = default;
  /* a plain vector as a (lazy) operation; the vector must outlive the expression */
  PackagedOperation(const Range &u) {
    const Range *p = &u;
    apply = [p](Range &v) { v = *p; };
    apply_add = [p](Range &v) { v += *p; };
    reinit_vector = [p](Range &v, bool omit) { v.reinit(*p, omit); };
  }
  operator Range() const {
    Range r;
    reinit_vector(r, /*omit_zeroing=*/true);
    apply(r);
    return r;
  }
};

template <typename Range>
PackagedOperation<Range> operator+(const PackagedOperation<Range> &a, const PackagedOperation<Range> &b) {
  PackagedOperation<Range> r;
  r.reinit_vector = a.reinit_vector;
  r.apply = [a, b](Range &v) {
    a.apply(v);
    b.apply_add(v);
  };
  r.apply_add = [a, b](Range &v) {
    a.apply_add(v);
    b.apply_add(v);
  };
  return r;
}
template <typename Range>
PackagedOperation<Range> operator*(const typename Range::value_type s, const PackagedOperation<Range> &a) {
  PackagedOperation<Range> r;
  r.reinit_vector = a.reinit_vector;
  r.apply = [s, a](Range &v) {
    a.apply(v);
    v *= s;
  };
  r.apply_add = [s, a](Range &v) {
    Range t;
    a.reinit_vector(t, true);
    a.apply(t);
    t *= s;
    v += t;
  };
  return r;
}
template <typename Range>
PackagedOperation<Range> operator-(const PackagedOperation<Range> &a, const PackagedOperation<Range> &b) {
  return a + typename Range::value_type(-1) * b;
}
template <typename Range>
PackagedOperation<Range> operator-(const PackagedOperation<Range> &a) {
  return typename Range::value_type(-1) * a;
}

/* vector (+|-) packaged, packaged (+|-) vector, vector (+|-) vector -- Range restricted to the two
 * vector types so these do not swallow arithmetic on unrelated types */
#define FDAL_STUB_VECTOR_OPS(VEC)                                                                               \
  template <typename N>                                                                                         \
  PackagedOperation<VEC<N>> operator+(const VEC<N> &u, const PackagedOperation<VEC<N>> &p) {                    \
    return PackagedOperation<VEC<N>>(u) + p;                                                                    \
  }                                                                                                             \
  template <typename N>                                                                                         \
  PackagedOperation<VEC<N>> operator-(const VEC<N> &u, const PackagedOperation<VEC<N>> &p) {                    \
    return PackagedOperation<VEC<N>>(u) - p;                                                                    \
  }                                                                                                             \
  template <typename N>                                                                                         \
  PackagedOperation<VEC<N>> operator+(const PackagedOperation<VEC<N>> &p, const VEC<N> &u) {                    \
    return p + PackagedOperation<VEC<N>>(u);                                                                    \
  }                                                                                                             \
  template <typename N>                                                                                         \
  PackagedOperation<VEC<N>> operator-(const PackagedOperation<VEC<N>> &p, const VEC<N> &u) {                    \
    return p - PackagedOperation<VEC<N>>(u);                                                                    \
  }                                                                                                             \
  template <typename N>                                                                                         \
  PackagedOperation<VEC<N>> operator+(const VEC<N> &u, const VEC<N> &w) {                                       \
    return PackagedOperation<VEC<N>>(u) + PackagedOperation<VEC<N>>(w);                                         \
  }                                                                                                             \
  template <typename N>                                                                                         \
  PackagedOperation<VEC<N>> operator-(const VEC<N> &u, const VEC<N> &w) {                                       \
    return PackagedOperation<VEC<N>>(u) - PackagedOperation<VEC<N>>(w);                                         \
  }                                                                                                             \
  template <typename N>                                                                                         \
  PackagedOperation<VEC<N>> operator*(const N s, const VEC<N> &u) {                                             \
    return s * PackagedOperation<VEC<N>>(u);                                                                    \
  }
FDAL_STUB_VECTOR_OPS(Vector)
FDAL_STUB_VECTOR_OPS(BlockVector)
#undef FDAL_STUB_VECTOR_OPS

/* op * u */
template <typename Range, typename Domain>
PackagedOperation<Range> operator*(const LinearOperator<Range, Domain> &op, const Domain &u) {
  PackagedOperation<Range> r;
  const Domain *p = &u;
  r.reinit_vector = op.reinit_range_vector;
  r.apply = [op, p](Range &v) { op.vmult(v, *p); };
  r.apply_add = [op, p](Range &v) { op.vmult_add(v, *p); };
  return r;
}
/* op * packaged */
template <typename Range, typename Domain>
PackagedOperation<Range> operator*(const LinearOperator<Range, Domain> &op, const PackagedOperation<Domain> &q) {
  PackagedOperation<Range> r;
  r.reinit_vector = op.reinit_range_vector;
  r.apply = [op, q](Range &v) {
    Domain t;
    q.reinit_vector(t, true);
    q.apply(t);
    op.vmult(v, t);
  };
  r.apply_add = [op, q](Range &v) {
    Domain t;
    q.reinit_vector(t, true);
    q.apply(t);
    op.vmult_add(v, t);
  };
  return r;
}

}  // namespace dealii
#endif
