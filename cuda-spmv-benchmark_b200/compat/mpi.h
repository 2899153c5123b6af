/*
 * mpi.h -- single-process stand-in for the handful of MPI calls the reference's multi-GPU main makes
 * (src/main/cg_solver_mgpu_stencil.cu:23-60,195: Init, Comm_rank, Comm_size, Abort, Finalize).
 *
 * The reference runs one MPI rank per GPU.  Here ONE process drives all the GPUs (peer-memory halos and
 * scalar exchanges inside libspmv_b200.so, no MPI on the data path), so the unmodified reference main
 * compiles against this header and runs as "rank 0": cg_solve_mgpu_partitioned() then spreads the row
 * bands over the GPUs itself.  MPI_Comm_size reports that GPU count (B200_GPUS, else every visible
 * device) so that the reference's own export_cg_mgpu_json(..., world_size) call records it.
 *
 * Put this directory on the include path INSTEAD of a real MPI installation:
 *   nvcc -I include -I include/solvers -I cuda-spmv-benchmark_b200/compat \
 *        <reference>/src/main/cg_solver_mgpu_stencil.cu -L... -lspmv_b200
 * and launch through cuda-spmv-benchmark_b200/scripts/mpirun (maps `-np P` to B200_GPUS=P).
 */
#ifndef B200_COMPAT_MPI_H
#define B200_COMPAT_MPI_H

#include <stdlib.h>

#include <cuda_runtime_api.h>

typedef int MPI_Comm;
#define MPI_COMM_WORLD 0
#define MPI_SUCCESS 0

static inline int b200_compat_world(void) {
    const char* e = getenv("B200_GPUS");
    int n = e ? atoi(e) : 0;
    if (n < 1 && cudaGetDeviceCount(&n) != cudaSuccess) n = 1;
    return n < 1 ? 1 : n;
}
static inline int MPI_Init(int* argc, char*** argv) { (void)argc; (void)argv; return MPI_SUCCESS; }
static inline int MPI_Finalize(void) { return MPI_SUCCESS; }
static inline int MPI_Comm_rank(MPI_Comm c, int* rank) { (void)c; *rank = 0; return MPI_SUCCESS; }
static inline int MPI_Comm_size(MPI_Comm c, int* size) { (void)c; *size = b200_compat_world(); return MPI_SUCCESS; }
static inline int MPI_Barrier(MPI_Comm c) { (void)c; return MPI_SUCCESS; }
static inline int MPI_Abort(MPI_Comm c, int code) { (void)c; exit(code); return MPI_SUCCESS; }

#endif /* B200_COMPAT_MPI_H */
