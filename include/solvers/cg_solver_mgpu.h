/* solvers/cg_solver_mgpu.h -- drop-in shim: same include name as the reference's include/solvers/cg_solver_mgpu.h; the
 * declarations live in b200/api.h. */
#pragma once
#include "../b200/api.h"
