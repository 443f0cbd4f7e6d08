/* stand-in header, see ../lac/stub_core.h (test infrastructure) */
#include <deal.II/lac/stub_core.h>
