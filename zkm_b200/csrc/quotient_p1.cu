// Part 1 of the quotient kernel instantiations (see quotient.cu).
#define ZKM_QPART 1
#include "quotient.cu"
