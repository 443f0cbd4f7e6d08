/*
 * TEST INFRASTRUCTURE: stand-in for the few Trilinos (Epetra / ML) declarations that
 * include/fdal_dealii.h::export_amg touches, written from the documented public interface of
 * ml_MultiLevelPreconditioner.h / ml_struct.h / ml_operator.h / Epetra_CrsMatrix.h (Trilinos >= 14.4) —
 * Trilinos itself is not installed in this image.  Purpose: export_amg goes through a compiler and
 * runs against a hierarchy laid out the way ML lays it out ("increasing" level numbering: level 0 =
 * finest; Pmat[l+1] maps level l+1 -> l; Rmat[l] maps level l -> l+1; Amat[l].lambda_max = the
 * eigenvalue estimate the Chebyshev smoother uses).  NOT a substitute for a build against real
 * Trilinos: types and member names are the real ones, behaviour is the minimum the adapter needs.
 */
#ifndef FDAL_TRILINOS_STUB_ML_H
#define FDAL_TRILINOS_STUB_ML_H
#include <cstdint>
#include <vector>

class Epetra_Operator {
public:
  virtual ~Epetra_Operator() = default;
};

/* row-wise view access of a (serial) Epetra_CrsMatrix: local == global indices on one rank */
class Epetra_CrsMatrix {
public:
  Epetra_CrsMatrix(int n_rows, int n_cols, std::vector<int> rp, std::vector<int> ci, std::vector<double> v)
    : n_rows_(n_rows), n_cols_(n_cols), rp_(std::move(rp)), ci_(std::move(ci)), v_(std::move(v)) {}
  int NumMyRows() const { return n_rows_; }
  int NumGlobalCols() const { return n_cols_; }
  int GCID(int local_col) const { return local_col; }
  int ExtractMyRowView(int row, int &n_entries, double *&values, int *&indices) const {
    n_entries = rp_[row + 1] - rp_[row];
    values = const_cast<double *>(v_.data()) + rp_[row];
    indices = const_cast<int *>(ci_.data()) + rp_[row];
    return 0;
  }

private:
  int n_rows_, n_cols_;
  std::vector<int> rp_, ci_;
  std::vector<double> v_;
};

/* ml_operator.h: only the members the adapter reads; `data` holds the stand-in's CSR */
struct ML_Operator {
  double lambda_max = 0.0;
  double lambda_min = 0.0;
  int invec_leng = 0, outvec_leng = 0;
  void *data = nullptr;
};
/* ml_struct.h */
struct ML {
  int ML_num_actual_levels = 0;
  int ML_num_levels = 0;
  ML_Operator *Amat = nullptr;
  ML_Operator *Pmat = nullptr;
  ML_Operator *Rmat = nullptr;
};

namespace ML_Epetra {
class MultiLevelPreconditioner : public Epetra_Operator {
public:
  explicit MultiLevelPreconditioner(const ML *ml) : ml_(ml) {}
  const ML *GetML(const int /*WhichML*/ = -1) const { return ml_; }

private:
  const ML *ml_;
};
}  // namespace ML_Epetra
#endif
