/* stand-in header, see stub_core.h (test infrastructure) */
#include "stub_core.h"
