// Drop-in replacement for the reference's grape/cuda/parallel/parallel_engine.h: same include path,
// same public names, implemented on the B200 engine (see b200_compat.h).
#ifndef GRAPE_B200_COMPAT_PARALLEL_PARALLEL_ENGINE_H
#define GRAPE_B200_COMPAT_PARALLEL_PARALLEL_ENGINE_H
#include "grape/cuda/b200_compat.h"
#endif
