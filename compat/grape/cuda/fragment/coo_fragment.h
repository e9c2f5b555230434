// Drop-in replacement for the reference's grape/cuda/fragment/coo_fragment.h: same include path,
// same public names, implemented on the B200 engine (see b200_compat.h).
#ifndef GRAPE_B200_COMPAT_FRAGMENT_COO_FRAGMENT_H
#define GRAPE_B200_COMPAT_FRAGMENT_COO_FRAGMENT_H
#include "grape/cuda/b200_compat.h"
#endif
