/* solvers/cg_metrics.h -- drop-in shim: same include name as the reference's include/solvers/cg_metrics.h; the
 * declarations live in b200/api.h. */
#pragma once
#include "../b200/api.h"
