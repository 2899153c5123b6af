// ingest.cpp -- Matrix Market file -> COO entries in DEVICE memory.
// The host only reads the raw bytes and the few header lines (reference src/io/io.cu:117-134);
// every "i j value" line is parsed by the GPU (b200_parse_mtx_entries).  Literals outside the exact
// fast path are re-read here with strtod, and files that are not "three tokens per line" fall back
// to the host reader followed by one upload -- results are identical either way.
#include <string>
#include <vector>

#include "host_common.h"

extern "C" int b200_load_matrix_market_device(const char* filename, MatrixData* meta, void** d_entries_out) {
    if (!filename || !meta || !d_entries_out) return 1;
    *d_entries_out = nullptr;
    meta->entries = nullptr; meta->rows = meta->cols = meta->nnz = 0; meta->grid_size = -1;
    FILE* f = fopen(filename, "rb");
    if (!f) { fprintf(stderr, "Error opening file\n"); return 1; }
    fseek(f, 0, SEEK_END);
    const long long size = ftell(f);
    fseek(f, 0, SEEK_SET);
    char* host = nullptr;
    if (cudaHostAlloc((void**)&host, (size_t)size + 1, cudaHostAllocDefault) != cudaSuccess) { fclose(f); return 3; }
    const size_t got = fread(host, 1, (size_t)size, f);
    fclose(f);
    host[got] = 0;
    // header: '%' lines (type qualifier, optional "% STENCIL_GRID_SIZE n"), then "rows cols nnz"
    long long pos = 0;
    bool symmetric = false, have_size = false;
    while (pos < (long long)got) {
        long long eol = pos;
        while (eol < (long long)got && host[eol] != '\n') eol++;
        std::string line(host + pos, host + eol);
        pos = eol + 1;
        if (!line.empty() && line[0] == '%') {
            if (line.find("symmetric") != std::string::npos) symmetric = true;
            if (line.find("STENCIL_GRID_SIZE") != std::string::npos) sscanf(line.c_str(), "%% STENCIL_GRID_SIZE %d", &meta->grid_size);
            continue;
        }
        if (sscanf(line.c_str(), "%d %d %d", &meta->rows, &meta->cols, &meta->nnz) == 3) have_size = true;
        break;
    }
    auto host_fallback = [&]() -> int {
        cudaFreeHost(host);
        MatrixData m;
        if (load_matrix_market(filename, &m) != 0) return 4;
        *meta = m;
        meta->entries = nullptr;
        void* d = nullptr;
        B200_CUDA(cudaMalloc(&d, (size_t)(m.nnz > 0 ? m.nnz : 1) * sizeof(Entry)));
        B200_CUDA(cudaMemcpy(d, m.entries, (size_t)m.nnz * sizeof(Entry), cudaMemcpyHostToDevice));
        free(m.entries);
        *d_entries_out = d;
        return 0;
    };
    if (!have_size || meta->nnz < 0) { cudaFreeHost(host); fprintf(stderr, "Error reading matrix size line\n"); return 2; }
    if (symmetric) return host_fallback();  // mirrored entries: host expansion, then upload
    const long long text_bytes = (long long)got - pos;
    void* d_text = nullptr;
    void* d_entries = nullptr;
    B200_CUDA(cudaMalloc(&d_text, (size_t)(text_bytes > 0 ? text_bytes : 1)));
    B200_CUDA(cudaMalloc(&d_entries, (size_t)(meta->nnz > 0 ? meta->nnz : 1) * sizeof(Entry)));
    B200_CUDA(cudaMemcpy(d_text, host + pos, (size_t)(text_bytes > 0 ? text_bytes : 0), cudaMemcpyHostToDevice));
    const int cap = 1 << 16;
    std::vector<long long> pairs((size_t)2 * cap);
    long long lines = 0;
    int inexact = 0, malformed = 0;
    int rc = b200_parse_mtx_entries(d_text, text_bytes, meta->nnz, d_entries, &lines, &inexact, pairs.data(), cap,
                                    &malformed, 0);
    cudaFree(d_text);
    if (rc != 0 || malformed || lines < meta->nnz || inexact > cap) {
        cudaFree(d_entries);
        if (rc == 0 && lines < meta->nnz && !malformed) {  // truncated file: same verdict as the host reader
            cudaFreeHost(host);
            fprintf(stderr, "Error reading matrix entry %lld (expected 3 items)\n", lines);
            return 4;
        }
        return host_fallback();
    }
    for (int k = 0; k < inexact; k++) {  // literals the exact fast path does not cover: strtod, like fscanf
        const long long entry = pairs[2 * k], off = pairs[2 * k + 1];
        const double v = strtod(host + pos + off, nullptr);
        B200_CUDA(cudaMemcpy(static_cast<char*>(d_entries) + entry * sizeof(Entry) + offsetof(Entry, value), &v,
                             sizeof(double), cudaMemcpyHostToDevice));
    }
    cudaFreeHost(host);
    *d_entries_out = d_entries;
    return 0;
}

extern "C" void b200_free_device(void* d_ptr) { cudaFree(d_ptr); }

// plain device->host copy of a library-owned buffer (bindings without a CUDA runtime of their own)
extern "C" int b200_copy_to_host(void* h_dst, const void* d_src, size_t bytes) {
    B200_CUDA(cudaMemcpy(h_dst, d_src, bytes, cudaMemcpyDeviceToHost));
    return 0;
}
