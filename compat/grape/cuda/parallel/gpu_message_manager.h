// Drop-in replacement for the reference's grape/cuda/parallel/gpu_message_manager.h: same include path,
// same public names, implemented on the B200 engine (see b200_compat.h).
#ifndef GRAPE_B200_COMPAT_PARALLEL_GPU_MESSAGE_MANAGER_H
#define GRAPE_B200_COMPAT_PARALLEL_GPU_MESSAGE_MANAGER_H
#include "grape/cuda/b200_compat.h"
#endif
