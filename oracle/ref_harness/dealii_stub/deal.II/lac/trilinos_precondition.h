/* stand-in header, see ../lac/stub_core.h (test infrastructure).  With -DFDAL_STUB_TRILINOS the
 * TrilinosWrappers::PreconditionAMG stand-in below exists too: it only carries the Epetra_Operator
 * (ML_Epetra::MultiLevelPreconditioner) that deal.II's class exposes through trilinos_operator(). */
#include <deal.II/lac/stub_core.h>
#ifdef FDAL_STUB_TRILINOS
#ifndef FDAL_STUB_PRECONDITION_AMG
#define FDAL_STUB_PRECONDITION_AMG
#include <ml_MultiLevelPreconditioner.h>

#include <memory>
namespace dealii {
namespace TrilinosWrappers {
class PreconditionAMG {
public:
  explicit PreconditionAMG(std::shared_ptr<Epetra_Operator> op) : preconditioner(std::move(op)) {}
  Epetra_Operator &trilinos_operator() const { return *preconditioner; }

private:
  std::shared_ptr<Epetra_Operator> preconditioner;
};
}  // namespace TrilinosWrappers
}  // namespace dealii
#endif
#endif
