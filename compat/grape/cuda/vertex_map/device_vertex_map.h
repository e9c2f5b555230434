// Drop-in replacement for the reference's grape/cuda/vertex_map/device_vertex_map.h: same include path,
// same public names, implemented on the B200 engine (see b200_compat.h).
#ifndef GRAPE_B200_COMPAT_VERTEX_MAP_DEVICE_VERTEX_MAP_H
#define GRAPE_B200_COMPAT_VERTEX_MAP_DEVICE_VERTEX_MAP_H
#include "grape/cuda/b200_compat.h"
#endif
