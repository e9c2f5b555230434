// Drop-in replacement for the reference's grape/cuda/app/batch_shuffle_app_base.h: same include path,
// same public names, implemented on the B200 engine (see b200_compat.h).
#ifndef GRAPE_B200_COMPAT_APP_BATCH_SHUFFLE_APP_BASE_H
#define GRAPE_B200_COMPAT_APP_BATCH_SHUFFLE_APP_BASE_H
#include "grape/cuda/b200_compat.h"
#endif
