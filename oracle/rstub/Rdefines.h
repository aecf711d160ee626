/* oracle/rstub/Rdefines.h -- TEST INFRASTRUCTURE (see R.h). */
#ifndef RSTUB_RDEFINES_H
#define RSTUB_RDEFINES_H
#include "Rinternals.h"
#endif
