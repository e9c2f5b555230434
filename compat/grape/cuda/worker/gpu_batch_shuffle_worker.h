// Drop-in replacement for the reference's grape/cuda/worker/gpu_batch_shuffle_worker.h: same include path,
// same public names, implemented on the B200 engine (see b200_compat.h).
#ifndef GRAPE_B200_COMPAT_WORKER_GPU_BATCH_SHUFFLE_WORKER_H
#define GRAPE_B200_COMPAT_WORKER_GPU_BATCH_SHUFFLE_WORKER_H
#include "grape/cuda/b200_compat.h"
#endif
