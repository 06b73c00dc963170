/* TEST INFRASTRUCTURE: see Rinternals.h in this directory. */
#ifndef BARTINT_R_MOCK_R_H
#define BARTINT_R_MOCK_R_H
#include <stdlib.h>
#endif
