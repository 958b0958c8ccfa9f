// Experiment: a torch pluggable allocator whose device memory is COMPRESSIBLE (cuMemCreate with CU_MEM_ALLOCATION_COMP_GENERIC), to
// measure what the L2's inline compression does for the step kernel's traffic (ncu: "bytes sent to the L2 Compression unit ... 0 %
// compressed" - ordinary allocations are not compressible). Build: g++ -O2 -shared -fPIC tools/comp_alloc.cpp -I/usr/local/cuda/include
// -L/usr/local/cuda/lib64/stubs -lcuda -o tools/libcomp_alloc.so ; use: tools/comp_probe.py
#include <cuda.h>
#include <stdio.h>
#include <stdlib.h>
#include <map>
#include <mutex>
struct Rec { CUmemGenericAllocationHandle h; size_t size; };
static std::map<CUdeviceptr, Rec> g_live;
static std::mutex g_mu;
static int g_comp = -1, g_report = 0;
extern "C" void *comp_malloc(ssize_t size, int device, void *stream) {
    if (g_comp < 0) g_comp = getenv("HEXB_COMP") ? atoi(getenv("HEXB_COMP")) : 1;
    CUmemAllocationProp prop = {};
    prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
    prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    prop.location.id = device;
    prop.allocFlags.compressionType = g_comp ? CU_MEM_ALLOCATION_COMP_GENERIC : CU_MEM_ALLOCATION_COMP_NONE;
    size_t gran = 0;
    if (cuMemGetAllocationGranularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED) != CUDA_SUCCESS || !gran) gran = 2u << 20;
    size_t sz = ((size_t)(size > 0 ? size : 1) + gran - 1) / gran * gran;
    CUmemGenericAllocationHandle h;
    if (cuMemCreate(&h, sz, &prop, 0) != CUDA_SUCCESS) { fprintf(stderr, "comp_alloc: cuMemCreate(%zu) failed\n", sz); return nullptr; }
    if (!g_report) {
        CUmemAllocationProp got = {};
        cuMemGetAllocationPropertiesFromHandle(&got, h);
        fprintf(stderr, "comp_alloc: requested compression %d, got compressionType %d, granularity %zu\n", g_comp, (int)got.allocFlags.compressionType, gran);
        g_report = 1;
    }
    CUdeviceptr p = 0;
    if (cuMemAddressReserve(&p, sz, 0, 0, 0) != CUDA_SUCCESS) { cuMemRelease(h); return nullptr; }
    if (cuMemMap(p, sz, 0, h, 0) != CUDA_SUCCESS) { cuMemAddressFree(p, sz); cuMemRelease(h); return nullptr; }
    CUmemAccessDesc acc = {};
    acc.location = prop.location;
    acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
    if (cuMemSetAccess(p, sz, &acc, 1) != CUDA_SUCCESS) { cuMemUnmap(p, sz); cuMemAddressFree(p, sz); cuMemRelease(h); return nullptr; }
    std::lock_guard<std::mutex> g(g_mu);
    g_live[p] = Rec{h, sz};
    return (void *)p;
}
extern "C" void comp_free(void *ptr, ssize_t size, int device, void *stream) {
    std::lock_guard<std::mutex> g(g_mu);
    auto it = g_live.find((CUdeviceptr)ptr);
    if (it == g_live.end()) return;
    cuMemUnmap(it->first, it->second.size);
    cuMemAddressFree(it->first, it->second.size);
    cuMemRelease(it->second.h);
    g_live.erase(it);
}
