// Part 2 of the quotient kernel instantiations (see quotient.cu).
#define ZKM_QPART 2
#include "quotient.cu"
