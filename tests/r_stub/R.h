/* TEST INFRASTRUCTURE — see Rinternals.h in this directory. */
#ifndef R_STUB_R_H
#define R_STUB_R_H
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#endif
