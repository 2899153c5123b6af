// ingest.cpp -- Matrix Market file -> COO entries in DEVICE memory.
// The host only reads the raw bytes and the few header lines (reference src/io/io.cu:117-134);
// every "i j value" line is parsed by the GPU (b200_parse_mtx_entries).  Literals outside the exact
// fast path are re-read here with strtod, and files that are not "three tokens per line" fall back
// to the host reader followed by one upload -- results are identical either way.
#include <string>
#include <vector>

#include "host_common.h"

namespace {
// owners: every early return of the loader releases what it holds
struct PinnedText {
    char* p = nullptr;
    ~PinnedText() { if (p) cudaFreeHost(p); }
};
struct DeviceMem {
    void* p = nullptr;
    ~DeviceMem() { if (p) cudaFree(p); }
    void* release() { void* q = p; p = nullptr; return q; }
};

int upload_host_entries(const char* filename, MatrixData* meta, void** d_entries_out) {
    MatrixData m;
    if (load_matrix_market(filename, &m) != 0) return 4;
    *meta = m;
    meta->entries = nullptr;
    DeviceMem d;
    cudaError_t e = cudaMalloc(&d.p, (size_t)(m.nnz > 0 ? m.nnz : 1) * sizeof(Entry));
    if (e == cudaSuccess) e = cudaMemcpy(d.p, m.entries, (size_t)m.nnz * sizeof(Entry), cudaMemcpyHostToDevice);
    free(m.entries);
    if (e != cudaSuccess) { fprintf(stderr, "[b200] CUDA error: %s (%s:%d)\n", cudaGetErrorString(e), __FILE__, __LINE__); return 2; }
    *d_entries_out = d.release();
    return 0;
}
}  // namespace

extern "C" int b200_load_matrix_market_device(const char* filename, MatrixData* meta, void** d_entries_out) {
    if (!filename || !meta || !d_entries_out) return 1;
    *d_entries_out = nullptr;
    meta->entries = nullptr; meta->rows = meta->cols = meta->nnz = 0; meta->grid_size = -1;
    FILE* f = fopen(filename, "rb");
    if (!f) { fprintf(stderr, "Error opening file\n"); return 1; }
    fseek(f, 0, SEEK_END);
    const long long size = ftell(f);
    fseek(f, 0, SEEK_SET);
    PinnedText text;
    if (cudaHostAlloc((void**)&text.p, (size_t)size + 1, cudaHostAllocDefault) != cudaSuccess) { fclose(f); return 3; }
    char* host = text.p;
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
    if (!have_size || meta->nnz < 0) { fprintf(stderr, "Error reading matrix size line\n"); return 2; }
    if (symmetric) return upload_host_entries(filename, meta, d_entries_out);  // mirrored entries: host expansion, then upload
    const long long text_bytes = (long long)got - pos;
    DeviceMem d_text, d_entries;
    B200_CUDA(cudaMalloc(&d_text.p, (size_t)(text_bytes > 0 ? text_bytes : 1)));
    B200_CUDA(cudaMalloc(&d_entries.p, (size_t)(meta->nnz > 0 ? meta->nnz : 1) * sizeof(Entry)));
    B200_CUDA(cudaMemcpy(d_text.p, host + pos, (size_t)(text_bytes > 0 ? text_bytes : 0), cudaMemcpyHostToDevice));
    const int cap = 1 << 16;
    std::vector<long long> pairs((size_t)2 * cap);
    long long lines = 0;
    int inexact = 0, malformed = 0;
    int rc = b200_parse_mtx_entries(d_text.p, text_bytes, meta->nnz, d_entries.p, &lines, &inexact, pairs.data(), cap,
                                    &malformed, 0);
    if (rc != 0 || malformed || lines < meta->nnz || inexact > cap) {
        if (rc == 0 && lines < meta->nnz && !malformed) {  // truncated file: same verdict as the host reader
            fprintf(stderr, "Error reading matrix entry %lld (expected 3 items)\n", lines);
            return 4;
        }
        return upload_host_entries(filename, meta, d_entries_out);
    }
    // literals the exact fast path does not cover: strtod on the host (like fscanf), then ONE staged
    // upload + scatter launch instead of one blocking 8-byte copy per literal
    for (int k = 0; k < inexact; k++) {
        const double v = strtod(host + pos + pairs[2 * k + 1], nullptr);
        memcpy(&pairs[2 * k + 1], &v, sizeof v);  // (entry index, byte offset) -> (entry index, value bits)
    }
    B200_K(b200_patch_entry_values(d_entries.p, pairs.data(), inexact, 0));
    *d_entries_out = d_entries.release();
    return 0;
}

extern "C" void b200_free_device(void* d_ptr) { cudaFree(d_ptr); }

// plain device->host copy of a library-owned buffer (bindings without a CUDA runtime of their own)
extern "C" int b200_copy_to_host(void* h_dst, const void* d_src, size_t bytes) {
    B200_CUDA(cudaMemcpy(h_dst, d_src, bytes, cudaMemcpyDeviceToHost));
    return 0;
}
