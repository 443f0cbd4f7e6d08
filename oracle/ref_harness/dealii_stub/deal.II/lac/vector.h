/* stand-in header, see stub_core.h (test infrastructure) */
#include <deal.II/lac/stub_core.h>
