// common.cuh — shared helpers for libuvip_orb.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <mutex>
#include <string>
#include <nvtx3/nvToolsExt.h>
#include "../../include/uvip_orb.h"

namespace uvip {

void set_last_error(const char* fmt, ...);

#define UVIP_CUDA(call)                                                                         \
    do {                                                                                        \
        cudaError_t e__ = (call);                                                               \
        if (e__ != cudaSuccess) {                                                               \
            uvip::set_last_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
            return UVIP_ERR_CUDA;                                                               \
        }                                                                                       \
    } while (0)

#define UVIP_CHECK_ARG(cond)                                                                    \
    do {                                                                                        \
        if (!(cond)) {                                                                          \
            uvip::set_last_error("%s:%d bad argument: %s", __FILE__, __LINE__, #cond);          \
            return UVIP_ERR_ARG;                                                                \
        }                                                                                       \
    } while (0)

// NVTX range around the host-side enqueue of one pipeline stage (header-only NVTX v3: a no-op unless a tool is attached).  ncu can
// then select a stage by name: `ncu --nvtx --nvtx-include "uvip_extract_group/uvip_fast/" ...` (tools/ncu_stage.sh).
struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
};

static inline int div_up(int a, int b) { return (a + b - 1) / b; }
static inline size_t align_up(size_t a, size_t b) { return (a + b - 1) / b * b; }

// RAII guard: make `device` current for the duration of a C-ABI call, restore afterwards
struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int device) {
        if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; return; }
        if (prev != device && cudaSetDevice(device) != cudaSuccess) ok = false;
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

// growable device scratch buffer owned by a handle
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    int reserve(size_t bytes) {
        if (bytes <= cap) return UVIP_OK;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = align_up(bytes, 256);
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) { set_last_error("cudaMalloc(%zu) -> %s", want, cudaGetErrorString(e)); return UVIP_ERR_CUDA; }
        cap = want;
        return UVIP_OK;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <typename T> T* as() const { return reinterpret_cast<T*>(p); }
};

#ifdef __CUDACC__
// ---- PTX helpers: mbarrier + 1-D bulk async copy (TMA engine, SASS UBLKCP) ----
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ int atom_add_shared(int* p, int v) {
    int old;
    asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(smem_u32(p)), "r"(v) : "memory");
    return old;
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "LAB_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra LAB_DONE_%=;\n"
        "bra LAB_WAIT_%=;\n"
        "LAB_DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// the same wait for barriers that may stay closed for microseconds: the hardware parks the thread for up to `ns` per attempt instead
// of spinning in the issue slots the other warps need
__device__ __forceinline__ void mbar_wait_parked(uint64_t* bar, uint32_t parity, uint32_t ns) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "LAB_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
        "@p bra LAB_DONE_%=;\n"
        "bra LAB_WAIT_%=;\n"
        "LAB_DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity), "r"(ns) : "memory");
}
// orders earlier generic-proxy accesses to shared memory before later async-proxy (TMA) accesses to the same bytes
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// 3-D tiled TMA load (SASS UTMALDG): box of the tensor map at element coordinates (x, y, z) -> shared memory
__device__ __forceinline__ void tma_load_3d(void* dst_smem, const void* tmap, int x, int y, int z, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(smem_u32(dst_smem)), "l"(tmap), "r"(x), "r"(y), "r"(z), "r"(smem_u32(bar)) : "memory");
}
#endif

}  // namespace uvip
