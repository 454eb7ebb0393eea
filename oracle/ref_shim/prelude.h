/* force-included ahead of every reference translation unit: the headers MSVC pulled in transitively
 * (tex.cpp / prog.cpp use std::out_of_range, rasterizer.cpp uses fabs/FLT_MAX/memset). */
#pragma once
#ifdef __cplusplus
#include <stdexcept>
#include <new>
#include "windows.h"   /* MSVC declares _aligned_malloc in <stdlib.h>/<malloc.h> (vbo.cpp:14, udm.cpp:30) */
#endif
#include <string.h>
#include <math.h>
#include <float.h>
#include <stdint.h>
#include <stdlib.h>
