/* SYNTAX-CHECK STUB ONLY - see Rinternals.h in this directory. */
#ifndef COCONS_STUB_R_H
#define COCONS_STUB_R_H
#include <stdlib.h>
#include <string.h>
#endif
