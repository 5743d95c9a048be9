/* TEST INFRASTRUCTURE: stand-in for R.h (see Rinternals.h in this directory). */
#ifndef RSTUB_R_H
#define RSTUB_R_H
#include <stdlib.h>
#include <stdio.h>
#include <math.h>
#endif
