/* TEST INFRASTRUCTURE: stand-in for ml_epetra_utils.h (see ml_MultiLevelPreconditioner.h beside it) */
#ifndef FDAL_TRILINOS_STUB_ML_EPETRA_UTILS_H
#define FDAL_TRILINOS_STUB_ML_EPETRA_UTILS_H
#include "ml_MultiLevelPreconditioner.h"

struct MlStubCsr {  // what ML_Operator::data points to in the stand-in
  int n_rows, n_cols;
  std::vector<int> rp, ci;
  std::vector<double> v;
};
/* ML's signature: int ML_Operator2EpetraCrsMatrix(ML_Operator *Amat, Epetra_CrsMatrix *&CrsMatrix,
 *                                                 int &MaxNumNonzeros, bool CheckNonzeroRow, double &CPUTime,
 *                                                 bool verbose = false);  the caller owns the result */
inline int ML_Operator2EpetraCrsMatrix(ML_Operator *Amat, Epetra_CrsMatrix *&CrsMatrix, int &MaxNumNonzeros,
                                       bool /*CheckNonzeroRow*/, double &CPUTime, bool /*verbose*/ = false) {
  const MlStubCsr *m = static_cast<const MlStubCsr *>(Amat->data);
  CrsMatrix = new Epetra_CrsMatrix(m->n_rows, m->n_cols, m->rp, m->ci, m->v);
  MaxNumNonzeros = 0;
  for (int r = 0; r < m->n_rows; ++r) MaxNumNonzeros = MaxNumNonzeros > m->rp[r + 1] - m->rp[r] ? MaxNumNonzeros : m->rp[r + 1] - m->rp[r];
  CPUTime = 0.0;
  return 0;
}
#endif
