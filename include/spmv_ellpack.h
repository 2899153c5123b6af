/* spmv_ellpack.h -- drop-in shim: same include name as the reference's include/spmv_ellpack.h; the declarations live
 * in b200/api.h (host API) and b200_kernels.h (thin kernel C ABI). */
#pragma once
#include "b200/api.h"
