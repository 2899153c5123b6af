// host_alloc.cpp -- pinned host buffers placed on the NUMA node the GPU hangs off.
//
// The solver entry points take HOST vectors (reference cg_solver.h:59-77): at 20k x 20k a solve moves
// 9.6 GB over PCIe, at 8 GPUs every rank moves 1.2 GB at the same time.  Pinned memory that sits on the
// other socket crosses the inter-socket link on every transfer.  b200_host_alloc_near() asks the kernel
// for pages on the device's own node (set_mempolicy(MPOL_PREFERRED) around the allocation + first touch,
// no libnuma needed) and pins them; when the platform gives no NUMA information it degrades to a plain
// pinned allocation.  b200_host_node_of_device() / b200_host_free() complete the set.
#include <errno.h>
#include <sys/syscall.h>
#include <unistd.h>

#include <string>

#include "host_common.h"

namespace {
constexpr int kMpolDefault = 0, kMpolPreferred = 1;

long set_mempolicy_raw(int mode, const unsigned long* mask, unsigned long maxnode) {
#ifdef SYS_set_mempolicy
    return syscall(SYS_set_mempolicy, mode, mask, maxnode);
#else
    (void)mode; (void)mask; (void)maxnode;
    errno = ENOSYS;
    return -1;
#endif
}
}  // namespace

// NUMA node of a CUDA device (-1: unknown / single node)
extern "C" int b200_host_node_of_device(int device) {
    char bus[32] = "";
    if (cudaDeviceGetPCIBusId(bus, sizeof bus, device) != cudaSuccess) { cudaGetLastError(); return -1; }
    for (char* c = bus; *c; c++) *c = (char)tolower(*c);
    const std::string path = std::string("/sys/bus/pci/devices/") + bus + "/numa_node";
    FILE* f = fopen(path.c_str(), "r");
    if (!f) return -1;
    int node = -1;
    if (fscanf(f, "%d", &node) != 1) node = -1;
    fclose(f);
    return node;
}

// bytes of pinned (page-locked, portable) host memory, first-touched on the NUMA node of `device`.
// *node_out (optional) = the node that was requested, -1 if placement was not possible.
extern "C" int b200_host_alloc_near(int device, size_t bytes, void** out, int* node_out) {
    if (!out || bytes == 0) return 1;
    *out = nullptr;
    int node = b200_host_node_of_device(device);
    bool bound = false;
    if (node >= 0 && node < 1024) {
        unsigned long mask[16] = {0};
        mask[node / (8 * sizeof(unsigned long))] |= 1ul << (node % (8 * sizeof(unsigned long)));
        bound = set_mempolicy_raw(kMpolPreferred, mask, 8 * sizeof mask) == 0;
    }
    void* p = nullptr;
    cudaError_t e = cudaHostAlloc(&p, bytes, cudaHostAllocPortable);
    if (e == cudaSuccess) {
        // cudaHostAlloc faults the pages in while it pins them -- under the policy set above
        volatile char* c = static_cast<volatile char*>(p);
        for (size_t i = 0; i < bytes; i += 4096) c[i] = 0;
    }
    if (bound) set_mempolicy_raw(kMpolDefault, nullptr, 0);
    if (e != cudaSuccess) {
        cudaGetLastError();
        fprintf(stderr, "[b200] cudaHostAlloc(%zu) failed: %s\n", bytes, cudaGetErrorString(e));
        return 2;
    }
    if (node_out) *node_out = bound ? node : -1;
    *out = p;
    return 0;
}

extern "C" void b200_host_free(void* p) {
    if (p) cudaFreeHost(p);
}
