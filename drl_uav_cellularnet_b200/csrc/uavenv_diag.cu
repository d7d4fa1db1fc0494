// Diagnostics library (include/uavenv_diag.h): zero-fill kernels that isolate the store mechanisms of the step kernel's
// observation stream.  Not part of the product C-ABI (no reference counterpart); used by profiles/write_ceiling.py,
// ring_sweep.py and env_pattern_sweep.py, whose measurements are in profiles/r1/NOTES.md.
#include "../../include/uavenv_diag.h"

#include <cuda_runtime.h>
#include <stdint.h>

#include "env_kernels.cuh"

namespace uavk {

constexpr int ZERO_TILE_BYTES = 16384;             // fill_kernel

// ---------------------------------------------------------------------------------------------------------
// Diagnostic: a pure zero-fill of `bytes` bytes with the two store mechanisms the step kernel can use, so the
// write-only HBM ceiling of the box can be measured next to the step kernel (bench.py --write-ceiling).
//   mode 0: st.global.cs.v4 from all threads;  mode 1: cp.async.bulk shared->global of a zero tile (UBLKCP)
__global__ void __launch_bounds__(CTA_THREADS) fill_kernel(char *dst, unsigned long long bytes, unsigned long long per_cta,
                                                           int mode) {
    __shared__ __align__(128) float zero_tile[ZERO_TILE_BYTES / 4];
    const unsigned long long lo = (unsigned long long)blockIdx.x * per_cta;
    if (lo >= bytes) return;
    const unsigned long long n = min(per_cta, bytes - lo);
    if (mode == 0) {
        float4 *d4 = reinterpret_cast<float4 *>(dst + lo);
        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        for (unsigned long long i = threadIdx.x; i < n / 16; i += CTA_THREADS) __stcs(d4 + i, z);
    } else {
        float4 *z4 = reinterpret_cast<float4 *>(zero_tile);
        for (int i = threadIdx.x; i < ZERO_TILE_BYTES / 16; i += CTA_THREADS) z4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        fence_proxy_async_smem();
        __syncthreads();
        if (threadIdx.x < 32) {
            for (unsigned long long off = (unsigned long long)threadIdx.x * ZERO_TILE_BYTES; off < n;
                 off += 32ull * ZERO_TILE_BYTES)
                bulk_store(dst + lo + off, zero_tile, (uint32_t)min((unsigned long long)ZERO_TILE_BYTES, n - off));
            bulk_commit();
            bulk_wait_all();
        }
    }
}

// mode 2: the store-warp pattern of env_kernel in isolation -- persistent CTAs, one warp, a ring of `ring` tiles,
// every tile re-armed (wait for its previous copy to be read, optional proxy fence) before the next bulk copy.
// CTA j streams chunks j, j + gridDim.x, ... of per_cta bytes.  flags bit0: fence.proxy.async per tile;
// bit1: all lanes of the warp issue (tile t by lane t % 32) instead of lane rb; bit2: rewrite a few floats of the
// tile before every copy (clear + atomic add, as the step kernel does); bit3: a short dependent chain per chunk.
__global__ void __launch_bounds__(128) fill_ring_kernel(char *dst, unsigned long long bytes, unsigned long long per_cta,
                                                        int ring, int tile_bytes, int flags) {
    extern __shared__ __align__(128) unsigned char dyn_smem[];
    float4 *z4 = reinterpret_cast<float4 *>(dyn_smem);
    for (int i = threadIdx.x; i < ring * tile_bytes / 16; i += blockDim.x) z4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    fence_proxy_async_smem();
    __syncthreads();
    if (threadIdx.x >= 32) return;
    const int lane = threadIdx.x;
    const unsigned long long n_chunks = (bytes + per_cta - 1) / per_cta;
    int g = 0;
    for (unsigned long long ch = blockIdx.x; ch < n_chunks; ch += gridDim.x) {
        const unsigned long long lo = ch * per_cta, n = min(per_cta, bytes - lo);
        const int nT = (int)((n + tile_bytes - 1) / tile_bytes);
        for (int t = 0; t < nT; t++, g++) {
            const int rb = g % ring;
            const int issuer = (flags & 2) ? (g % 32) : rb;
            if (g >= ring) {
                // the lane that issued the previous copy from this tile waits for its read
                const int prev_issuer = (flags & 2) ? ((g - ring) % 32) : rb;
                if (lane == prev_issuer) bulk_wait_read_all();
                __syncwarp();
            }
            if (flags & 4) {
                // what the step kernel's store warp does per tile: clear the previous counts, add the new ones
                float *buf = reinterpret_cast<float *>(dyn_smem + (size_t)rb * tile_bytes);
                const int nf = tile_bytes / 4;
                if (lane < 4 && g >= ring) buf[(unsigned)(lane * 977 + (g - ring) * 131) % (unsigned)nf] = 0.f;
                __syncwarp();
                if (lane < 4) atomicAdd(&buf[(unsigned)(lane * 977 + g * 131) % (unsigned)nf], 1.f);
            }
            if (flags & 1) fence_proxy_async_smem();
            __syncwarp();
            if (flags & 8) {
                // emulate the per-env sort of the store warp: a dependent chain of shared-memory round trips
                if (t == 0) {
                    int v = lane;
                    for (int r = 0; r < 12; r++) { v = __shfl_xor_sync(0xffffffffu, v, 1) + 1; }
                    if (v == -1) bulk_commit();
                }
            }
            if (lane == issuer) {
                const unsigned long long off = (unsigned long long)t * tile_bytes;
                bulk_store(dst + lo + off, dyn_smem + (size_t)rb * tile_bytes, (uint32_t)min((unsigned long long)tile_bytes, n - off));
                bulk_commit();
            }
        }
    }
    bulk_wait_read_all();
}

// mode 3: the "zeros now, counts one chunk later" pattern -- persistent CTAs, one warp; per chunk the lanes issue the
// chunk's tiles from ONE constant zero tile, commit, wait for the PREVIOUS chunk's group and (flags bit0) add
// n_red float REDs into the previous chunk.  flags bit1: wait for the chunk's own group instead (no overlap).
__global__ void __launch_bounds__(128) fill_env_kernel(char *dst, unsigned long long bytes, unsigned long long per_cta,
                                                       int tile_bytes, int flags, int n_red) {
    extern __shared__ __align__(128) unsigned char dyn_smem[];
    float4 *z4 = reinterpret_cast<float4 *>(dyn_smem);
    for (int i = threadIdx.x; i < tile_bytes / 16; i += blockDim.x) z4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    fence_proxy_async_smem();
    __syncthreads();
    if (threadIdx.x >= 32) return;
    const int lane = threadIdx.x;
    const unsigned long long n_chunks = (bytes + per_cta - 1) / per_cta;
    float *prev = nullptr;
    unsigned prev_n = 0;
    for (unsigned long long ch = blockIdx.x; ch < n_chunks; ch += gridDim.x) {
        const unsigned long long lo = ch * per_cta, n = min(per_cta, bytes - lo);
        for (unsigned long long off = (unsigned long long)lane * tile_bytes; off < n; off += 32ull * tile_bytes)
            bulk_store(dst + lo + off, dyn_smem, (uint32_t)min((unsigned long long)tile_bytes, n - off));
        bulk_commit();
        if (flags & 2) {
            bulk_wait_all();
            __syncwarp();
            if (flags & 1)
                for (int q = lane; q < n_red; q += 32)
                    atomicAdd(reinterpret_cast<float *>(dst + lo) + (unsigned)(q * 1031 + 17) % (unsigned)(n / 4), 1.f);
        } else {
            if (prev) {
                bulk_wait_prev();
                __syncwarp();
                if (flags & 1)
                    for (int q = lane; q < n_red; q += 32) atomicAdd(prev + (unsigned)(q * 1031 + 17) % prev_n, 1.f);
            }
            prev = reinterpret_cast<float *>(dst + lo);
            prev_n = (unsigned)(n / 4);
        }
    }
    bulk_wait_all();
}

}  // namespace uavk

using namespace uavk;

extern "C" {

int uavenv_diag_fill(void *dst_dev, int64_t bytes, int64_t bytes_per_cta, int32_t mode, void *stream) {
    if (!dst_dev || bytes < 16 || (bytes & 15) || bytes_per_cta < 16 || (bytes_per_cta & 15) || (mode != 0 && mode != 1) ||
        ((uintptr_t)dst_dev & 15))
        return UAVENV_DIAG_EINVAL;
    const int64_t grid = (bytes + bytes_per_cta - 1) / bytes_per_cta;
    if (grid > 0x7fffffffLL) return UAVENV_DIAG_EINVAL;
    fill_kernel<<<(unsigned)grid, CTA_THREADS, 0, (cudaStream_t)stream>>>((char *)dst_dev, (unsigned long long)bytes,
                                                                          (unsigned long long)bytes_per_cta, mode);
    return cudaGetLastError() == cudaSuccess ? UAVENV_DIAG_OK : UAVENV_DIAG_ECUDA;
}

int uavenv_diag_fill_ring(void *dst_dev, int64_t bytes, int64_t bytes_per_chunk, int32_t grid, int32_t ring,
                          int32_t tile_bytes, int32_t flags, void *stream) {
    if (!dst_dev || bytes < 16 || (bytes & 15) || bytes_per_chunk < 16 || (bytes_per_chunk & 15) || grid < 1 || ring < 1 ||
        ring > 32 || tile_bytes < 128 || (tile_bytes & 127) || ((uintptr_t)dst_dev & 15))
        return UAVENV_DIAG_EINVAL;
    const size_t dyn = (size_t)ring * tile_bytes;
    if (dyn > 200 * 1024) return UAVENV_DIAG_EINVAL;
    if (cudaFuncSetAttribute((const void *)fill_ring_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn) != cudaSuccess)
        return UAVENV_DIAG_ECUDA;
    fill_ring_kernel<<<grid, 128, dyn, (cudaStream_t)stream>>>((char *)dst_dev, (unsigned long long)bytes,
                                                               (unsigned long long)bytes_per_chunk, ring, tile_bytes, flags);
    return cudaGetLastError() == cudaSuccess ? UAVENV_DIAG_OK : UAVENV_DIAG_ECUDA;
}

int uavenv_diag_fill_env(void *dst_dev, int64_t bytes, int64_t bytes_per_chunk, int32_t grid, int32_t tile_bytes,
                         int32_t flags, int32_t n_red, void *stream) {
    if (!dst_dev || bytes < 16 || (bytes & 15) || bytes_per_chunk < 16 || (bytes_per_chunk & 15) || grid < 1 ||
        tile_bytes < 128 || (tile_bytes & 127) || tile_bytes > 200 * 1024 || n_red < 0 || ((uintptr_t)dst_dev & 15))
        return UAVENV_DIAG_EINVAL;
    if (cudaFuncSetAttribute((const void *)fill_env_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tile_bytes) != cudaSuccess)
        return UAVENV_DIAG_ECUDA;
    fill_env_kernel<<<grid, 128, tile_bytes, (cudaStream_t)stream>>>((char *)dst_dev, (unsigned long long)bytes,
                                                                     (unsigned long long)bytes_per_chunk, tile_bytes, flags, n_red);
    return cudaGetLastError() == cudaSuccess ? UAVENV_DIAG_OK : UAVENV_DIAG_ECUDA;
}

}  // extern "C"
