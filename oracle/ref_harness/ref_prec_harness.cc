/*
 * Pinning harness — TEST INFRASTRUCTURE, never on the product path.
 *
 * Compiles the UNMODIFIED reference header /root/reference/augmented_lagrangian_preconditioner.h
 * (from where it lies; nothing is copied into this repository) against the deal.II stand-in
 * types of dealii_stub/, and exposes the vmult() of its five preconditioner classes through one
 * C entry point.  Every LinearOperator the classes receive is backed by a caller-supplied C
 * callback, so a test can feed the oracle's (or the CUDA library's) operator applications in and
 * compare the block algebra the REFERENCE CODE ITSELF performs with fdalo_apply_prec /
 * fdal_apply_prec.  Built by ../Makefile into ../_ref/libref_prec.so (git-ignored).
 */
#include <augmented_lagrangian_preconditioner.h>

#include <cstdint>
#include <cstring>
#include <exception>

extern "C" {
/* operator ids handed to the callback */
enum {
  REF_OP_AUG_INV = 0,       /* Aug_inv / A11_inv : n0 -> n0                               */
  REF_OP_A22_INV = 1,       /* A22_inv           : n1 -> n1                               */
  REF_OP_AUG_INV_BLOCK = 2, /* LinearOperator<BlockVector> Aug_inv : (n0+n1) -> (n0+n1)   */
  REF_OP_C = 3,             /* C  : n0 -> n_lambda                                        */
  REF_OP_CT = 4,            /* Ct : n_lambda -> n0                                        */
  REF_OP_BT = 5,            /* Bt : n1 -> n0                                              */
  REF_OP_INVW = 6,          /* invW : n_lambda -> n_lambda                                */
  REF_OP_MP_INV = 7,        /* Mp_inv : n1 -> n1                                          */
  REF_OP_M = 8              /* M (elliptic): n_lambda -> n1                               */
};
/* kinds: the values of FDAL_KIND_* in include/fdal.h */
enum { REF_KIND_LAPLACE = 0, REF_KIND_STOKES = 1, REF_KIND_STOKES_DIAG = 2, REF_KIND_ELLIPTIC_IDEAL = 3,
       REF_KIND_ELLIPTIC_MODIFIED = 4 };
typedef void (*ref_op_fn)(void *user, int op, const double *x, int64_t nx, double *y, int64_t ny);
}

namespace {
using Vec = dealii::Vector<double>;
using BVec = dealii::BlockVector<double>;

struct Callbacks {
  ref_op_fn fn;
  void *user;
};

dealii::LinearOperator<Vec> make_op(const Callbacks cb, const int op, const int t_op, const std::size_t n_range,
                                    const std::size_t n_domain) {
  dealii::LinearOperator<Vec> L;
  L.reinit_range_vector = [n_range](Vec &v, bool omit) { v.reinit(n_range, omit); };
  L.reinit_domain_vector = [n_domain](Vec &v, bool omit) { v.reinit(n_domain, omit); };
  L.vmult = [cb, op, n_range](Vec &y, const Vec &x) {
    if (y.size() != n_range) y.reinit(n_range, true);
    cb.fn(cb.user, op, x.begin(), static_cast<int64_t>(x.size()), y.begin(), static_cast<int64_t>(y.size()));
  };
  L.vmult_add = [cb, op, n_range](Vec &y, const Vec &x) {
    Vec t(n_range);
    cb.fn(cb.user, op, x.begin(), static_cast<int64_t>(x.size()), t.begin(), static_cast<int64_t>(t.size()));
    y += t;
  };
  if (t_op >= 0) {
    L.Tvmult = [cb, t_op, n_domain](Vec &y, const Vec &x) {
      if (y.size() != n_domain) y.reinit(n_domain, true);
      cb.fn(cb.user, t_op, x.begin(), static_cast<int64_t>(x.size()), y.begin(), static_cast<int64_t>(y.size()));
    };
    L.Tvmult_add = [cb, t_op, n_domain](Vec &y, const Vec &x) {
      Vec t(n_domain);
      cb.fn(cb.user, t_op, x.begin(), static_cast<int64_t>(x.size()), t.begin(), static_cast<int64_t>(t.size()));
      y += t;
    };
  }
  return L;
}

dealii::LinearOperator<BVec> make_block_aug_inv(const Callbacks cb, const std::size_t n0, const std::size_t n1) {
  dealii::LinearOperator<BVec> L;
  auto reinit = [n0, n1](BVec &v, bool omit) {
    v.reinit(2);
    v.block(0).reinit(n0, omit);
    v.block(1).reinit(n1, omit);
  };
  L.reinit_range_vector = reinit;
  L.reinit_domain_vector = reinit;
  L.vmult = [cb, n0, n1](BVec &y, const BVec &x) {
    std::vector<double> in(n0 + n1), out(n0 + n1);
    std::memcpy(in.data(), x.block(0).begin(), n0 * sizeof(double));
    std::memcpy(in.data() + n0, x.block(1).begin(), n1 * sizeof(double));
    cb.fn(cb.user, REF_OP_AUG_INV_BLOCK, in.data(), static_cast<int64_t>(n0 + n1), out.data(),
          static_cast<int64_t>(n0 + n1));
    std::memcpy(y.block(0).begin(), out.data(), n0 * sizeof(double));
    std::memcpy(y.block(1).begin(), out.data() + n0, n1 * sizeof(double));
  };
  return L;
}

void load(BVec &v, const double *src) {
  for (unsigned int b = 0; b < v.n_blocks(); ++b) {
    std::memcpy(v.block(b).begin(), src, v.block(b).size() * sizeof(double));
    src += v.block(b).size();
  }
}
void store(const BVec &v, double *dst) {
  for (unsigned int b = 0; b < v.n_blocks(); ++b) {
    std::memcpy(dst, v.block(b).begin(), v.block(b).size() * sizeof(double));
    dst += v.block(b).size();
  }
}
}  // namespace

extern "C" {

const char *ref_prec_source(void) { return REF_HEADER_PATH; }

/* v = P.vmult(u) with P the reference class for `kind`; sizes = block sizes of u (2 for the
 * Laplace kind, 3 otherwise).  Returns 0, or 1 when the reference code threw. */
int ref_prec_vmult(int kind, double gamma, double gamma_grad_div, const int64_t *sizes, ref_op_fn fn, void *user,
                   const double *u_in, double *v_out) {
  try {
    const Callbacks cb{fn, user};
    const bool two = kind == REF_KIND_LAPLACE;
    const std::size_t n0 = sizes[0], n1 = sizes[1], nl = two ? sizes[1] : sizes[2];
    std::vector<std::size_t> bs = two ? std::vector<std::size_t>{n0, nl} : std::vector<std::size_t>{n0, n1, nl};
    BVec u(bs), v(bs);
    load(u, u_in);
    const auto Aug_inv = make_op(cb, REF_OP_AUG_INV, REF_OP_AUG_INV, n0, n0);
    const auto A22_inv = make_op(cb, REF_OP_A22_INV, REF_OP_A22_INV, n1, n1);
    const auto C = make_op(cb, REF_OP_C, REF_OP_CT, nl, n0);
    const auto Ct = make_op(cb, REF_OP_CT, REF_OP_C, n0, nl);
    const auto Bt = make_op(cb, REF_OP_BT, -1, n0, n1);
    const auto invW = make_op(cb, REF_OP_INVW, REF_OP_INVW, nl, nl);
    const auto Mp_inv = make_op(cb, REF_OP_MP_INV, REF_OP_MP_INV, n1, n1);
    const auto M = make_op(cb, REF_OP_M, -1, n1, nl);
    switch (kind) {
      case REF_KIND_LAPLACE: {
        const BlockPreconditionerAugmentedLagrangian P(Aug_inv, C, Ct, invW, gamma);
        P.vmult(v, u);
      } break;
      case REF_KIND_STOKES: {
        const BlockPreconditionerAugmentedLagrangianStokes P(Aug_inv, Bt, Ct, invW, Mp_inv, gamma, gamma_grad_div);
        P.vmult(v, u);
      } break;
      case REF_KIND_STOKES_DIAG: {
        const BlockPreconditionerAugmentedLagrangianDiagonal P(Aug_inv, invW, Mp_inv, gamma, gamma_grad_div);
        P.vmult(v, u);
      } break;
      case REF_KIND_ELLIPTIC_IDEAL: {
        const EllipticInterfacePreconditioners::BlockTriangularALPreconditioner P(make_block_aug_inv(cb, n0, n1), C, M,
                                                                                  invW, gamma);
        P.vmult(v, u);
      } break;
      case REF_KIND_ELLIPTIC_MODIFIED: {
        const EllipticInterfacePreconditioners::BlockTriangularALPreconditionerModified P(C, M, invW, gamma, Aug_inv,
                                                                                         A22_inv);
        P.vmult(v, u);
      } break;
      default:
        return 2;
    }
    store(v, v_out);
    return 0;
  } catch (const std::exception &) {
    return 1;
  }
}
}
