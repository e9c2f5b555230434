// Drop-in replacement for the reference's grape/cuda/utils/sorted_search.h: same include path,
// same public names, implemented on the B200 engine (see b200_compat.h).
#ifndef GRAPE_B200_COMPAT_UTILS_SORTED_SEARCH_H
#define GRAPE_B200_COMPAT_UTILS_SORTED_SEARCH_H
#include "grape/cuda/b200_compat.h"
#endif
