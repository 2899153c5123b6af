// cli_common.h -- argument helpers shared by the command-line tools.
#pragma once
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "b200/api.h"

extern "C" MatrixData b200_synthetic_stencil(int grid_size);
extern "C" int b200_mgpu_init_single_process(int world, const int* devices, int max_grid);
extern "C" int b200_mgpu_world(void);
extern "C" int b200_load_matrix_market_device(const char* filename, MatrixData* meta, void** d_entries_out);
extern "C" int b200_operator_init_device_coo(SpmvOperator* op, const MatrixData* meta, const void* d_entries);
extern "C" void b200_free_device(void* d_ptr);
extern "C" int b200_pcg_set_preconditioner(int kind);

struct CliArgs {
    std::string matrix;                 // .mtx path ("" with --grid)
    std::vector<std::string> modes;     // --mode=a,b
    std::string json, csv;
    int grid = 0;                       // --grid=n : synthetic stencil, generated on the device
    int gpus = 0;                       // --gpus=P
    double tol = 1e-6;
    int maxiter = 1000;
    bool timers = false, host = false;
    bool jacobi = false;                // --precond=jacobi | --precond=block-jacobi (extension: preconditioned CG)
    bool device_ingest = false;         // --device-ingest: parse the .mtx and build the CSR on the GPU (extension)
    int runs = 10;
};

inline bool starts(const char* s, const char* p) { return strncmp(s, p, strlen(p)) == 0; }

inline CliArgs parse_cli(int argc, char** argv) {
    CliArgs a;
    for (int i = 1; i < argc; i++) {
        const char* s = argv[i];
        if (starts(s, "--mode=")) {
            std::string m = s + 7;
            size_t pos = 0;
            while (pos <= m.size() && a.modes.size() < 10) {  // at most 10 modes, like the reference
                size_t c = m.find(',', pos);
                if (c == std::string::npos) c = m.size();
                if (c > pos) a.modes.push_back(m.substr(pos, c - pos));
                pos = c + 1;
            }
        } else if (starts(s, "--json=")) a.json = s + 7;
        else if (starts(s, "--csv=")) a.csv = s + 6;
        else if (starts(s, "--grid=")) a.grid = atoi(s + 7);
        else if (starts(s, "--gpus=")) a.gpus = atoi(s + 7);
        else if (starts(s, "--tol=")) a.tol = atof(s + 6);
        else if (starts(s, "--maxiter=")) a.maxiter = atoi(s + 10);
        else if (starts(s, "--runs=")) a.runs = atoi(s + 7);
        else if (!strcmp(s, "--timers")) a.timers = true;
        else if (!strcmp(s, "--host")) a.host = true;
        else if (!strcmp(s, "--precond=jacobi")) { a.jacobi = true; b200_pcg_set_preconditioner(1); }
        else if (!strcmp(s, "--precond=block-jacobi")) { a.jacobi = true; b200_pcg_set_preconditioner(2); }
        else if (!strcmp(s, "--device-ingest")) a.device_ingest = true;
        else if (s[0] != '-' && a.matrix.empty()) a.matrix = s;
    }
    return a;
}

// "<base>_<name><ext>" (or "<file>_<name>.json" without an extension) -- reference main.cu:200-210
inline std::string per_mode_name(const std::string& file, const std::string& name, const char* default_ext) {
    size_t dot = file.rfind('.');
    size_t slash = file.rfind('/');
    if (dot != std::string::npos && (slash == std::string::npos || dot > slash))
        return file.substr(0, dot) + "_" + name + file.substr(dot);
    return file + "_" + name + default_ext;
}

// --device-ingest: the file's entry lines are parsed on the GPU (b200_load_matrix_market_device) and stay
// there; *d_entries receives the device COO array, mat->entries stays NULL (no host Entry[] / CSR at all)
inline int load_or_generate(const CliArgs& a, MatrixData* mat, void** d_entries = nullptr) {
    if (d_entries) *d_entries = nullptr;
    if (a.device_ingest && a.grid <= 0) {
        if (!d_entries) { fprintf(stderr, "--device-ingest is not available for this tool\n"); return 1; }
        if (b200_load_matrix_market_device(a.matrix.c_str(), mat, d_entries) != 0) {
            fprintf(stderr, "Failed to load matrix %s on the device\n", a.matrix.c_str());
            return 1;
        }
        printf("Matrix parsed on the device: %d rows, %d nonzeros (COO resident in HBM, no host copy)\n", mat->rows, mat->nnz);
        return 0;
    }
    if (a.grid > 0) {
        *mat = b200_synthetic_stencil(a.grid);
        printf("Synthetic 5-point stencil %dx%d generated on the device (no .mtx): %d rows, %d nonzeros\n", a.grid,
               a.grid, mat->rows, mat->nnz);
        return 0;
    }
    if (load_matrix_market(a.matrix.c_str(), mat) != 0) {
        fprintf(stderr, "Failed to load matrix %s\n", a.matrix.c_str());
        return 1;
    }
    return 0;
}

// operator set-up for either ingest path; device COO: CSR-based operators only (COO -> CSR on the GPU)
inline int init_operator(SpmvOperator* op, MatrixData* mat, const void* d_entries) {
    if (!d_entries) return op->init(mat);
    if (strcmp(op->name, "cusparse-csr") != 0 && strcmp(op->name, "stencil5-csr") != 0) {
        fprintf(stderr, "--device-ingest supports the cusparse-csr and stencil5-csr operators (got '%s')\n", op->name);
        return 1;
    }
    return b200_operator_init_device_coo(op, mat, d_entries);
}
