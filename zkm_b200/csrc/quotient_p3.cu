// Part 3 of the quotient kernel instantiations (see quotient.cu).
#define ZKM_QPART 3
#include "quotient.cu"
