// C-ABI of the actor-critic MLP's sparse first layer and optimiser step (include/uavnet.h), sm_100a.
//
// Reference semantics restated (file:line into the reference repo):
//   main.py:147,151   tf.layers.dense(self.s, 200, tf.nn.relu6)  on a count vector with ~44 non-zeros of 50 000
//                     -> sparse_fwd_kernel (gather-sum of weight rows) / sparse_bwd_kernel (scatter-add of the gradient)
//   main.py:300-301   tf.train.RMSPropOptimizer(1e-4)            -> rmsprop_kernel (TF1 defaults: decay .9, eps 1e-10)
#include "../../include/uavnet.h"

#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include "philox.cuh"
#include "tc_gemm.cuh"
#include <stdlib.h>

namespace uavk {

constexpr int NET_THREADS = 256;

// (-DUAVNET_BOUNDS_CHECK builds test every index against n_rows and raise bit 1 of the device error word that
// uavnet_gemm_check reads; the production build trusts the env kernel's obs_idx.  The gather-side backward and its counting
// sort ignore out-of-range indices in every build.)
// One thread per (sample m, float4 column c4): consecutive threads read consecutive 16-byte pieces of the same weight
// row (coalesced, rows come from L2 -- the 80 MB first-layer matrix is L2-resident on B200), the row indices are
// warp-broadcast loads.  K is unrolled by 8 so that 8 independent row reads are in flight per thread.  (A warp-per-sample
// variant with the indices broadcast by shuffles has fewer load instructions per byte but a quarter of the threads in
// flight: 39.9 us against 35.4 us for 8192 x 44 rows x 1.6 kB -- the kernel wants the parallelism.)
__global__ void __launch_bounds__(NET_THREADS) sparse_fwd_kernel(const int32_t *__restrict__ idx, long long M, int K,
                                                                      const float4 *__restrict__ W4, const float4 *__restrict__ b4,
                                                                      int H4, float4 *__restrict__ out4, int relu6, int n_rows,
                                                                      unsigned int *err) {
    const long long total = M * H4;
    for (long long w = (long long)blockIdx.x * NET_THREADS + threadIdx.x; w < total; w += (long long)gridDim.x * NET_THREADS) {
        const long long m = w / H4;
        const int c4 = (int)(w - m * H4);
        const int32_t *row = idx + m * K;
        float4 acc = b4 ? __ldg(b4 + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
        int k = 0;
        for (; k + 8 <= K; k += 8) {
            int r[8];
#pragma unroll
            for (int j = 0; j < 8; j++) r[j] = __ldg(row + k + j);
#ifdef UAVNET_BOUNDS_CHECK
#pragma unroll
            for (int j = 0; j < 8; j++) if ((unsigned)r[j] >= (unsigned)n_rows) { if (err) atomicOr(err, 2u); r[j] = 0; }
#endif
            float4 v[8];
#pragma unroll
            for (int j = 0; j < 8; j++) v[j] = __ldg(W4 + (size_t)r[j] * H4 + c4);
#pragma unroll
            for (int j = 0; j < 8; j++) { acc.x += v[j].x; acc.y += v[j].y; acc.z += v[j].z; acc.w += v[j].w; }
        }
        for (; k < K; k++) {
            int rk = __ldg(row + k);
#ifdef UAVNET_BOUNDS_CHECK
            if ((unsigned)rk >= (unsigned)n_rows) { if (err) atomicOr(err, 2u); rk = 0; }
#endif
            const float4 v = __ldg(W4 + (size_t)rk * H4 + c4);
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        if (relu6) {
            acc.x = fminf(fmaxf(acc.x, 0.f), 6.f); acc.y = fminf(fmaxf(acc.y, 0.f), 6.f);
            acc.z = fminf(fmaxf(acc.z, 0.f), 6.f); acc.w = fminf(fmaxf(acc.w, 0.f), 6.f);
        }
        out4[w] = acc;
    }
}

// dW[idx[m,k], :] += dpre[m, :] with 16-byte vector REDs (red.global.add.v4.f32).  Pieces whose four gradients are all
// zero (units outside relu6's linear range) are skipped.
__global__ void __launch_bounds__(NET_THREADS) sparse_bwd_kernel(const int32_t *__restrict__ idx, long long M, int K,
                                                                 const float4 *__restrict__ dpre4, int H4, float4 *dW4) {
    const long long total = M * H4;
    for (long long w = (long long)blockIdx.x * NET_THREADS + threadIdx.x; w < total; w += (long long)gridDim.x * NET_THREADS) {
        const long long m = w / H4;
        const int c4 = (int)(w - m * H4);
        const float4 g = __ldg(dpre4 + w);
        if (g.x == 0.f && g.y == 0.f && g.z == 0.f && g.w == 0.f) continue;
        const int32_t *row = idx + m * K;
        for (int k = 0; k < K; k++) atomicAdd(dW4 + (size_t)__ldg(row + k) * H4 + c4, g);
    }
}

// ---------------------------------------------------------------------------------------------------------
// First-layer weight gradient, gather side: dW[r, :] += sum over the (sample, slot) pairs with idx[m, k] = r of dpre[m, :].
// The scatter version above sends 360 M 16-byte REDs through the L2 atomic units (0.9 ms for a rollout batch of 81 920 x 44).
// Here the pairs are first bucketed by row (a counting sort of the 3.6 M row indices: histogram, exclusive scan, scatter of
// the sample numbers), then every row sums its samples' gradient vectors with plain loads -- the same access pattern as
// sparse_fwd_kernel with the roles of weights and activations swapped -- and is written once.  The caller runs the sum per
// column half so that the gathered operand (dpre[:, half], 65 MB for the rollout batch) stays L2-resident.
__global__ void __launch_bounds__(NET_THREADS) row_hist_kernel(const int32_t *__restrict__ idx, long long n, int n_rows,
                                                               int32_t *__restrict__ count) {
    for (long long i = (long long)blockIdx.x * NET_THREADS + threadIdx.x; i < n; i += (long long)gridDim.x * NET_THREADS) {
        const int r = __ldg(idx + i);
        if ((unsigned)r < (unsigned)n_rows) atomicAdd(count + r, 1);
    }
}

// exclusive scan of count[0..n_rows) into cursor[0..n_rows) (the fill kernel's write positions); one CTA of 1024 threads walks
// the array in tiles of 4096 (one coalesced int4 per thread): thread-local scan, warp scan by shuffles, the 32 warp totals
// through shared memory, a running carry between tiles.  n_rows is padded to a multiple of 4 by the workspace layout.
__global__ void __launch_bounds__(1024) row_scan_kernel(const int32_t *__restrict__ count, int n_rows, int32_t *__restrict__ cursor) {
    __shared__ int32_t wtot[32];
    __shared__ int32_t carry_sh;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) carry_sh = 0;
    __syncthreads();
    const int n4 = (n_rows + 3) >> 2;
    for (int base = 0; base < n4; base += 1024) {
        const int i4 = base + tid;
        int4 c = make_int4(0, 0, 0, 0);
        if (i4 < n4) c = reinterpret_cast<const int4 *>(count)[i4];
        if (4 * i4 + 1 >= n_rows) c.y = 0;
        if (4 * i4 + 2 >= n_rows) c.z = 0;
        if (4 * i4 + 3 >= n_rows) c.w = 0;
        if (4 * i4 >= n_rows) c.x = 0;
        const int32_t mine = c.x + c.y + c.z + c.w;
        int32_t inc = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int32_t up = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += up;
        }
        if (lane == 31) wtot[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            int32_t w = wtot[lane], winc = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int32_t up = __shfl_up_sync(0xffffffffu, winc, o);
                if (lane >= o) winc += up;
            }
            wtot[lane] = winc - w;                          // exclusive warp offsets
        }
        __syncthreads();
        const int32_t carry = carry_sh;
        int32_t ex = carry + wtot[warp] + inc - mine;       // exclusive prefix of this thread's first element
        if (i4 < n4) {
            int4 o;
            o.x = ex; o.y = ex + c.x; o.z = o.y + c.y; o.w = o.z + c.z;
            reinterpret_cast<int4 *>(cursor)[i4] = o;
        }
        __syncthreads();
        if (tid == 1023) carry_sh = ex + mine;
        __syncthreads();
    }
}

// bucket fill: position = cursor[row]++; the sorted arrays hold the sample number and the row of every (sample, slot) pair
__global__ void __launch_bounds__(NET_THREADS) row_fill_kernel(const int32_t *__restrict__ idx, long long n, int K, int n_rows,
                                                               int32_t *__restrict__ cursor, int32_t *__restrict__ sample_of,
                                                               int32_t *__restrict__ row_of) {
    for (long long i = (long long)blockIdx.x * NET_THREADS + threadIdx.x; i < n; i += (long long)gridDim.x * NET_THREADS) {
        const int r = __ldg(idx + i);
        if ((unsigned)r < (unsigned)n_rows) {
            const int pos = atomicAdd(cursor + r, 1);
            sample_of[pos] = (int32_t)(i / K);
            row_of[pos] = r;
        }
    }
}

// Work item = (slice of SLICE consecutive entries of the row-sorted pair list, float4 column of [c0_4, c0_4 + nc4)): the
// thread adds up the gradient vectors of its entries and flushes the sum with one 16-byte RED whenever the row changes and at
// the end of the slice.  Fixed-size slices keep the load balanced whatever the row histogram looks like -- in the aggregation
// phases of the mobility model a few hundred cells hold most UEs, and a thread-per-row sum then serialises on them (measured:
// 0.58 ms with spread-out UEs, 2.2 ms right after an aggregation phase) -- and there are ~60x fewer REDs than in the scatter.
constexpr int BWD_SLICE = 64;
__global__ void __launch_bounds__(NET_THREADS) sparse_bwd_gather_kernel(const int32_t *__restrict__ sample_of,
                                                                        const int32_t *__restrict__ row_of, const int32_t *total_dev,
                                                                        const float4 *__restrict__ dpre4, int H4, int c0_4, int nc4,
                                                                        float4 *dW4) {
    const int total = *total_dev;                           // pairs that landed in a bucket (= cursor of the last row after the fill)
    const long long n_slices = (total + BWD_SLICE - 1) / BWD_SLICE, items = n_slices * nc4;
    for (long long w = (long long)blockIdx.x * NET_THREADS + threadIdx.x; w < items; w += (long long)gridDim.x * NET_THREADS) {
        const long long sl = w / nc4;
        const int c4 = c0_4 + (int)(w - sl * nc4);
        const int lo = (int)(sl * BWD_SLICE), hi = min(lo + BWD_SLICE, total);
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        int cur = __ldg(row_of + lo);
        for (int j = lo; j < hi; j += 8) {
            int r[8], m[8];
            float4 v[8];
#pragma unroll
            for (int q = 0; q < 8; q++) {
                const int jj = min(j + q, hi - 1);
                r[q] = __ldg(row_of + jj); m[q] = __ldg(sample_of + jj);
            }
#pragma unroll
            for (int q = 0; q < 8; q++) v[q] = __ldg(dpre4 + (size_t)m[q] * H4 + c4);
#pragma unroll
            for (int q = 0; q < 8; q++) {
                if (j + q < hi) {
                    if (r[q] != cur) {
                        atomicAdd(dW4 + (size_t)cur * H4 + c4, acc);
                        acc = make_float4(0.f, 0.f, 0.f, 0.f);
                        cur = r[q];
                    }
                    acc.x += v[q].x; acc.y += v[q].y; acc.z += v[q].z; acc.w += v[q].w;
                }
            }
        }
        atomicAdd(dW4 + (size_t)cur * H4 + c4, acc);
    }
}

__global__ void __launch_bounds__(NET_THREADS) rmsprop_kernel(float *__restrict__ p, float *__restrict__ g, float *__restrict__ ms,
                                                              long long n, float lr, float decay, float eps, float gs, int zero_grad) {
    const long long n4 = n >> 2;
    float4 *p4 = reinterpret_cast<float4 *>(p), *g4 = reinterpret_cast<float4 *>(g), *m4 = reinterpret_cast<float4 *>(ms);
    const float od = 1.f - decay;
    for (long long i = (long long)blockIdx.x * NET_THREADS + threadIdx.x; i < n4; i += (long long)gridDim.x * NET_THREADS) {
        float4 gv = g4[i], mv = m4[i], pv = p4[i];
        gv.x *= gs; gv.y *= gs; gv.z *= gs; gv.w *= gs;
        mv.x = decay * mv.x + od * gv.x * gv.x; mv.y = decay * mv.y + od * gv.y * gv.y;
        mv.z = decay * mv.z + od * gv.z * gv.z; mv.w = decay * mv.w + od * gv.w * gv.w;
        pv.x -= lr * gv.x / sqrtf(mv.x + eps); pv.y -= lr * gv.y / sqrtf(mv.y + eps);
        pv.z -= lr * gv.z / sqrtf(mv.z + eps); pv.w -= lr * gv.w / sqrtf(mv.w + eps);
        m4[i] = mv; p4[i] = pv;
        if (zero_grad) g4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    // tail (n not a multiple of 4)
    const long long t = (n4 << 2) + (long long)blockIdx.x * NET_THREADS + threadIdx.x;
    if (t < n) {
        const float gv = g[t] * gs;
        const float mv = decay * ms[t] + od * gv * gv;
        ms[t] = mv;
        p[t] -= lr * gv / sqrtf(mv + eps);
        if (zero_grad) g[t] = 0.f;
    }
}

// Gradient of the actor loss w.r.t. the logits, one warp per sample (main.py:68-76 through the softmax):
//   L = -(1/M) sum_i [ log(p_i,a_i + 1e-5) * td_i + beta * H_i ],   H_i = -sum_j p_ij log(p_ij + 1e-5)
//   g_ij = dL/dp_ij = (beta/M) (log(p_ij + 1e-5) + p_ij / (p_ij + 1e-5)) - [j = a_i] td_i / (M (p_i,a_i + 1e-5))
//   dz_ij = p_ij (g_ij - sum_k p_ik g_ik)
// One read of the probabilities (the second pass hits L1), one write of dz, and the sample's loss term.
template <int PL>   // PL = ceil(A / 32) register slots per lane; 0 = any A, probabilities re-read for the second pass
__global__ void __launch_bounds__(NET_THREADS) actor_head_bwd_kernel(const float *__restrict__ prob, const long long *__restrict__ a_his,
                                                                     const float *__restrict__ td, long long M, int A, float beta,
                                                                     float inv_m, float *__restrict__ dz, long long ldz,
                                                                     float *__restrict__ loss_row) {
    const int lane = threadIdx.x & 31;
    const long long warp0 = ((long long)blockIdx.x * NET_THREADS + threadIdx.x) >> 5;
    const long long n_warps = ((long long)gridDim.x * NET_THREADS) >> 5;
    const float bm = beta * inv_m;
    for (long long m = warp0; m < M; m += n_warps) {
        const float *pr = prob + m * A;
        const int a = (int)a_his[m];
        const float t = td[m];
        float pa, ga;
        float ent = 0.f, dot = 0.f;                       // sum p*lp (= -H), sum p*g
        float *dr = dz + m * ldz;
        if constexpr (PL > 0) {
            float pv[PL], gv[PL];                         // the row lives in registers between the two passes
#pragma unroll
            for (int k = 0; k < PL; k++) {                // all loads of the row in flight before anything is used
                const int j = lane + 32 * k;
                pv[k] = j < A ? __ldg(pr + j) : 0.f;
            }
            float cand = 0.f;                             // p[a] out of the registers: no second dependent load
#pragma unroll
            for (int k = 0; k < PL; k++) cand = (k == (a >> 5)) ? pv[k] : cand;
            pa = __shfl_sync(0xffffffffu, cand, a & 31);
            ga = -t * inv_m / (pa + 1e-5f);
#pragma unroll
            for (int k = 0; k < PL; k++) {
                const int j = lane + 32 * k;
                const float lp = __logf(pv[k] + 1e-5f);
                gv[k] = bm * (lp + __fdividef(pv[k], pv[k] + 1e-5f)) + (j == a ? ga : 0.f);
                ent += pv[k] * lp;
                dot += pv[k] * gv[k];
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                ent += __shfl_xor_sync(0xffffffffu, ent, o);
                dot += __shfl_xor_sync(0xffffffffu, dot, o);
            }
#pragma unroll
            for (int k = 0; k < PL; k++) {
                const int j = lane + 32 * k;
                if (j < A) dr[j] = pv[k] * (gv[k] - dot);
            }
        } else {
            pa = pr[a];
            ga = -t * inv_m / (pa + 1e-5f);
            for (int j = lane; j < A; j += 32) {
                const float p = pr[j];
                const float lp = __logf(p + 1e-5f);
                ent += p * lp;
                dot += p * (bm * (lp + __fdividef(p, p + 1e-5f)) + (j == a ? ga : 0.f));
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                ent += __shfl_xor_sync(0xffffffffu, ent, o);
                dot += __shfl_xor_sync(0xffffffffu, dot, o);
            }
            for (int j = lane; j < A; j += 32) {
                const float p = pr[j];
                const float lp = __logf(p + 1e-5f);
                dr[j] = p * ((bm * (lp + __fdividef(p, p + 1e-5f)) + (j == a ? ga : 0.f)) - dot);
            }
        }
        if (lane == 0 && loss_row) loss_row[m] = -(__logf(pa + 1e-5f) * t - beta * ent);
    }
}

// The gradient push fused with the optimiser over NVLink peer memory (main.py:85-86,159-163 on `world` GPUs):
// rank r owns elements [lo, hi) of the flat buffers.  For each of them it sums the gradient over ALL ranks with peer
// loads (reduce-scatter), applies the RMSProp step once, and stores the new parameter -- and a zero gradient -- into
// rank's parameter buffer with peer stores (all-gather).  One pass over 1/world of the parameters per GPU, no staging
// buffer, no separate optimiser kernel.  The caller orders it between two tiny stream-ordered collectives and clears
// its own gradient buffer afterwards (the peers read the other slices of it during their launches).
struct PeerPtrs {
    float *g[UAVNET_MAX_PEERS];
    float *p[UAVNET_MAX_PEERS];
};

__global__ void __launch_bounds__(NET_THREADS) p2p_rmsprop_kernel(const __grid_constant__ PeerPtrs pp, float *__restrict__ ms,
                                                                  long long lo4, long long hi4, int rank, int world, float lr,
                                                                  float decay, float eps, float gs) {
    const float od = 1.f - decay;
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
    for (long long i = lo4 + (long long)blockIdx.x * NET_THREADS + threadIdx.x; i < hi4; i += (long long)gridDim.x * NET_THREADS) {
        float4 acc = zero;
        // fixed summation order (rank 0, 1, ...): every rank computes bit-identical parameters
        for (int r = 0; r < world; r++) {
            const float4 v = reinterpret_cast<const float4 *>(pp.g[r])[i];
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        acc.x *= gs; acc.y *= gs; acc.z *= gs; acc.w *= gs;
        float4 mv = reinterpret_cast<float4 *>(ms)[i], pv = reinterpret_cast<const float4 *>(pp.p[rank])[i];
        mv.x = decay * mv.x + od * acc.x * acc.x; mv.y = decay * mv.y + od * acc.y * acc.y;
        mv.z = decay * mv.z + od * acc.z * acc.z; mv.w = decay * mv.w + od * acc.w * acc.w;
        pv.x -= lr * acc.x / sqrtf(mv.x + eps); pv.y -= lr * acc.y / sqrtf(mv.y + eps);
        pv.z -= lr * acc.z / sqrtf(mv.z + eps); pv.w -= lr * acc.w / sqrtf(mv.w + eps);
        reinterpret_cast<float4 *>(ms)[i] = mv;
        for (int j = 0; j < world; j++) {
            const int r = (rank + j) % world;                 // start with the local copy, then walk the peers
            reinterpret_cast<float4 *>(pp.p[r])[i] = pv;
        }
        reinterpret_cast<float4 *>(pp.g[rank])[i] = zero;     // own slice of the own gradient buffer; the rest of it is
                                                              // still being read by the peers: the caller clears it later
    }
}

// ---------------------------------------------------------------------------------------------------------
// The same push without any host-side or NCCL ordering: the ranks synchronise through flag words in peer memory.
//   flags (uint32, one 256-byte IPC buffer per rank):  [0..7] ready[r]  [8..15] done[r]  [16] block counter (push kernel)
//                                                      [17] epoch of the last completed push  [18] block counter (finish)
//                                                      [19] sticky: a wait gave up (peer missing)
// Push number e = flags[17] + 1 on every rank (all ranks run the same number of pushes; the count lives in device
// memory so that a CUDA graph replays it).  Protocol per rank:
//   push kernel    block 0 writes ready[rank] = e into EVERY rank's flags (all kernels before it in the stream, i.e. the
//                  whole backward pass, have completed);  every block waits until ready[r] >= e for all r (all ranks'
//                  gradients are complete), then does its share of reduce-scatter + RMSProp + all-gather;  the last
//                  block to finish writes done[rank] = e into every rank's flags.
//   finish kernel  waits until done[r] >= e for all r (every rank's slice has reached this rank's parameters, nobody
//                  reads this rank's gradients any more), zeroes the slices of the own gradient buffer that the push
//                  kernel did not, and publishes flags[17] = e.
// A rank never waits for something that depends on its own later progress, so the protocol cannot deadlock as long as
// every rank launches both kernels; every wait is bounded all the same (about 2 s, then flags[19] is raised and the
// kernel carries on): a missing peer must not hang the GPU.
struct PeerPtrs2 {
    float *g[UAVNET_MAX_PEERS];
    float *p[UAVNET_MAX_PEERS];
    unsigned int *f[UAVNET_MAX_PEERS];
};
enum { PF_READY = 0, PF_DONE = 8, PF_CNT_PUSH = 16, PF_EPOCH = 17, PF_CNT_FIN = 18, PF_TIMEOUT = 19 };

__device__ __forceinline__ unsigned int ld_flag(const unsigned int *p) {
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_flag(unsigned int *p, unsigned int v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// all flags[base + r] >= e, r < world; bounded
__device__ __forceinline__ void wait_flags(unsigned int *mine, int base, int world, unsigned int e) {
    const long long t0 = clock64();
    for (int r = 0; r < world; r++) {
        while ((int)(ld_flag(mine + base + r) - e) < 0) {
            __nanosleep(64);
            if (clock64() - t0 > 4000000000LL) { mine[PF_TIMEOUT] = 1u; return; }
        }
    }
}

// What a push covers: up to two regions of the flat buffers, each a sub-matrix (all rows, a column range) of a row-major
// matrix or a plain range -- in float4 units.  The elements are numbered region a first, then region b; rank r owns the
// r-th of `world` equal slices of that numbering.  (The learner pushes the actor half of the first layer's gradient as soon
// as it is complete, and the critic half together with all the small layers at the end.)
struct PushRegion { long long base4, n4; int row_w4, col0_4, sub_w4; };
struct PushRegions { PushRegion a, b; };
__device__ __forceinline__ long long push_flat4(const PushRegions &R, long long j) {
    const bool in_a = j < R.a.n4;
    const PushRegion &r = in_a ? R.a : R.b;
    const long long jj = in_a ? j : j - R.a.n4;
    if (r.sub_w4 == r.row_w4) return r.base4 + jj;
    const long long row = jj / r.sub_w4;
    return r.base4 + row * r.row_w4 + r.col0_4 + (jj - row * r.sub_w4);
}

template <int W>   // W = world size known at compile time (all peer loads of an element in flight at once), 0 = any
__global__ void __launch_bounds__(NET_THREADS) p2p_push_kernel(const __grid_constant__ PeerPtrs2 pp, float *__restrict__ ms,
                                                               const __grid_constant__ PushRegions R, long long lo4,
                                                               long long hi4, int rank, int world_rt, float lr, float decay,
                                                               float eps, float gs, int dbg) {
    const int world = W > 0 ? W : world_rt;
    unsigned int *mine = pp.f[rank];
    __shared__ unsigned int e_sh;
    if (threadIdx.x == 0) {
        const unsigned int e = ld_flag(mine + PF_EPOCH) + 1u;
        if (blockIdx.x == 0) {
            __threadfence_system();
            for (int r = 0; r < world; r++) st_flag(pp.f[(rank + r) % world] + PF_READY + rank, e);
        }
        wait_flags(mine, PF_READY, world, e);
        e_sh = e;
    }
    __syncthreads();
    const float od = 1.f - decay;
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
    for (long long jn = lo4 + (long long)blockIdx.x * NET_THREADS + threadIdx.x; jn < hi4; jn += (long long)gridDim.x * NET_THREADS) {
        const long long i = push_flat4(R, jn);
        float4 acc = zero;
        if constexpr (W > 0) {
            // an NVLink peer load takes microseconds: all W of them are requested before the first is used
            float4 v[W];
#pragma unroll
            for (int r = 0; r < W; r++) v[r] = __ldcs(reinterpret_cast<const float4 *>(pp.g[(dbg & 2) ? rank : r]) + i);
#pragma unroll
            for (int r = 0; r < W; r++) { acc.x += v[r].x; acc.y += v[r].y; acc.z += v[r].z; acc.w += v[r].w; }   // fixed order
        } else {
            for (int r = 0; r < world; r++) {              // fixed summation order: bit-identical parameters on every rank
                const float4 v = reinterpret_cast<const float4 *>(pp.g[r])[i];
                acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
            }
        }
        acc.x *= gs; acc.y *= gs; acc.z *= gs; acc.w *= gs;
        float4 mv = reinterpret_cast<float4 *>(ms)[i], pv = reinterpret_cast<const float4 *>(pp.p[rank])[i];
        mv.x = decay * mv.x + od * acc.x * acc.x; mv.y = decay * mv.y + od * acc.y * acc.y;
        mv.z = decay * mv.z + od * acc.z * acc.z; mv.w = decay * mv.w + od * acc.w * acc.w;
        pv.x -= lr * acc.x / sqrtf(mv.x + eps); pv.y -= lr * acc.y / sqrtf(mv.y + eps);
        pv.z -= lr * acc.z / sqrtf(mv.z + eps); pv.w -= lr * acc.w / sqrtf(mv.w + eps);
        reinterpret_cast<float4 *>(ms)[i] = mv;
        for (int j = 0; j < world; j++) reinterpret_cast<float4 *>(pp.p[(dbg & 1) ? rank : (rank + j) % world])[i] = pv;
        reinterpret_cast<float4 *>(pp.g[rank])[i] = zero;
    }
    __threadfence_system();                                 // this thread's peer stores are performed
    __syncthreads();
    if (threadIdx.x == 0) {
        if (atomicAdd(mine + PF_CNT_PUSH, 1u) == gridDim.x - 1) {
            mine[PF_CNT_PUSH] = 0u;
            __threadfence_system();
            for (int r = 0; r < world; r++) st_flag(pp.f[(rank + r) % world] + PF_DONE + rank, e_sh);
        }
    }
}

// waits for every rank's push kernel, zeroes the pushed elements of the own gradient buffer outside the own slice [lo4, hi4)
// (the push kernel zeroed that one) and publishes the epoch
__global__ void __launch_bounds__(NET_THREADS) p2p_finish_kernel(unsigned int *mine, float4 *__restrict__ g4,
                                                                 const __grid_constant__ PushRegions R, long long lo4, long long hi4,
                                                                 int world) {
    __shared__ unsigned int e_sh;
    if (threadIdx.x == 0) {
        const unsigned int e = ld_flag(mine + PF_EPOCH) + 1u;
        wait_flags(mine, PF_DONE, world, e);
        e_sh = e;
    }
    __syncthreads();
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
    const long long n = R.a.n4 + R.b.n4;
    for (long long j = (long long)blockIdx.x * NET_THREADS + threadIdx.x; j < n; j += (long long)gridDim.x * NET_THREADS)
        if (j < lo4 || j >= hi4) g4[push_flat4(R, j)] = zero;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(mine + PF_CNT_FIN, 1u) == gridDim.x - 1) {
            mine[PF_CNT_FIN] = 0u;
            st_flag(mine + PF_EPOCH, e_sh);
        }
    }
}

// softmax over the logits + np.random.choice(N_A, p=a_prob) (main.py:149,165-169), one warp per sample: the row is
// read once into registers (PL > 0: PL = ceil(A / 32) slots per lane), exponentiated once, the probabilities are written
// for the update, and the action is drawn by inverse CDF with one Philox uniform keyed by (seed, global row, call
// counter): the first j whose cumulative probability exceeds u.  Cumulative sums run over chunks of 32 consecutive
// actions (warp-inclusive scan), so the pick does not depend on PL.
template <int PL>   // 0 = any A (the row is re-read from L1/L2 for every pass)
__global__ void __launch_bounds__(NET_THREADS) softmax_sample_kernel(const float *__restrict__ logits, long long M, int A,
                                                                     uint32_t k0, uint32_t k1, uint32_t row0,
                                                                     const uint32_t *__restrict__ counter_dev, uint32_t counter_add,
                                                                     float *__restrict__ prob, long long *__restrict__ action) {
    const int lane = threadIdx.x & 31;
    const uint32_t seq = (counter_dev ? *counter_dev : 0u) + counter_add;
    const long long warp0 = ((long long)blockIdx.x * NET_THREADS + threadIdx.x) >> 5;
    const long long n_warps = ((long long)gridDim.x * NET_THREADS) >> 5;
    for (long long m = warp0; m < M; m += n_warps) {
        const float *z = logits + m * A;
        float zv[PL > 0 ? PL : 1];
        float mx = -3.0e38f;
        if constexpr (PL > 0) {
#pragma unroll
            for (int k = 0; k < PL; k++) {
                const int j = lane + 32 * k;
                zv[k] = j < A ? __ldg(z + j) : -3.0e38f;
                mx = fmaxf(mx, zv[k]);
            }
        } else {
            for (int j = lane; j < A; j += 32) mx = fmaxf(mx, z[j]);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        float sum = 0.f;
        if constexpr (PL > 0) {
#pragma unroll
            for (int k = 0; k < PL; k++) {
                zv[k] = (lane + 32 * k < A) ? expf(zv[k] - mx) : 0.f;      // exponentiated once, kept in registers
                sum += zv[k];
            }
        } else {
            for (int j = lane; j < A; j += 32) sum += expf(z[j] - mx);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        const float inv = 1.f / sum;
        double ua, ub;
        philox_uniform2(k0, k1, row0 + (uint32_t)m, 0u, seq, DOM_SAMPLE, ua, ub);
        const float target = (float)ua;                   // in [0, 1)
        int pick = -1;
        if constexpr (PL > 0) {
            // probabilities out (coalesced), then a two-level inverse-CDF search over chunks of 32 consecutive actions: the
            // chunk totals of all PL chunks by one butterfly (every lane ends up with all of them), a serial walk over the
            // totals to the chunk the uniform falls into, and ONE warp-inclusive scan inside that chunk -- instead of a
            // scan + ballot per chunk (the kernel was bound by those: 520 -> ~300 instructions per row)
#pragma unroll
            for (int k = 0; k < PL; k++) {
                const int j = lane + 32 * k;
                if (j < A && prob) prob[m * A + j] = zv[k] * inv;
            }
            float T[PL];
#pragma unroll
            for (int k = 0; k < PL; k++) T[k] = zv[k];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
                for (int k = 0; k < PL; k++) T[k] += __shfl_xor_sync(0xffffffffu, T[k], o);
            }
            const float thr = target * sum;                // compare unnormalised sums with u * sum
            float carry = 0.f, carry_at = 0.f, val = 0.f;
            int kst = -1;
#pragma unroll
            for (int k = 0; k < PL; k++) {
                if (kst < 0 && carry + T[k] > thr) { kst = k; carry_at = carry; val = zv[k]; }
                carry += T[k];
            }
            if (kst >= 0) {
                float inc = val;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const float up = __shfl_up_sync(0xffffffffu, inc, o);
                    if (lane >= o) inc += up;
                }
                const int j = 32 * kst + lane;
                const unsigned hit = __ballot_sync(0xffffffffu, j < A && carry_at + inc > thr);
                // (the chunk total and the scan add the same 32 numbers in different orders: if rounding left no lane
                // above the threshold, the chunk's last action is the pick)
                pick = hit ? 32 * kst + __ffs(hit) - 1 : min(32 * kst + 31, A - 1);
            }
        } else {
            // any A: chunks of 32 consecutive actions, a warp-inclusive scan per chunk, first crossing wins
            float carry = 0.f;
            for (int base = 0; base < A; base += 32) {
                const int j = base + lane;
                const float p = j < A ? expf(z[j] - mx) * inv : 0.f;
                if (j < A && prob) prob[m * A + j] = p;
                float inc = p;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const float up = __shfl_up_sync(0xffffffffu, inc, o);
                    if (lane >= o) inc += up;
                }
                const unsigned hit = __ballot_sync(0xffffffffu, j < A && carry + inc > target);
                if (pick < 0 && hit) pick = base + __ffs(hit) - 1;
                carry += __shfl_sync(0xffffffffu, inc, 31);
            }
        }
        if (pick < 0) pick = A - 1;                       // rounding: the cumulative sum ended just below u
        if (lane == 0 && action) action[m] = pick;
    }
}

// out[m, j] = dv[m] * w[j] * (0 < h[m, j] < 6): the critic's value-head data gradient through relu6 (a rank-1 product)
__global__ void __launch_bounds__(NET_THREADS) rank1_mask_kernel(const float *__restrict__ dv, const float4 *__restrict__ w4,
                                                                 const float4 *__restrict__ h4, long long M, int H4, float4 *__restrict__ out4) {
    const long long total = M * H4;
    for (long long i = (long long)blockIdx.x * NET_THREADS + threadIdx.x; i < total; i += (long long)gridDim.x * NET_THREADS) {
        const long long m = i / H4;
        const int c = (int)(i - m * H4);
        const float d = __ldg(dv + m);
        const float4 w = __ldg(w4 + c), h = __ldg(h4 + i);
        float4 o;
        o.x = (h.x > 0.f && h.x < 6.f) ? d * w.x : 0.f; o.y = (h.y > 0.f && h.y < 6.f) ? d * w.y : 0.f;
        o.z = (h.z > 0.f && h.z < 6.f) ? d * w.z : 0.f; o.w = (h.w > 0.f && h.w < 6.f) ? d * w.w : 0.f;
        out4[i] = o;
    }
}

// Critic loss pieces of one update in one launch (main.py:64-66,80): td = v_target - v, dv = d(mean td^2)/dv = -2 td / M,
// and the two loss scalars as sums of per-block partial sums: loss[0] += sum(a_loss_row) / M (rows from
// actor_head_bwd_kernel of the PREVIOUS call are not used: a_rows may be null), loss[1] += sum(td^2) / M.  `loss` is
// zeroed by the caller (2 floats).
__global__ void __launch_bounds__(NET_THREADS) critic_td_kernel(const float *__restrict__ v_target, const float *__restrict__ v,
                                                                long long M, float inv_m, float *__restrict__ td,
                                                                float *__restrict__ dv, float *__restrict__ loss) {
    __shared__ float part[NET_THREADS / 32];
    float acc = 0.f;
    for (long long i = (long long)blockIdx.x * NET_THREADS + threadIdx.x; i < M; i += (long long)gridDim.x * NET_THREADS) {
        const float t = v_target[i] - v[i];
        td[i] = t;
        dv[i] = -2.f * inv_m * t;
        acc += t * t;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float tot = 0.f;
        for (int w = 0; w < NET_THREADS / 32; w++) tot += part[w];
        atomicAdd(loss + 1, tot * inv_m);
    }
}

// loss[0] += sum(rows) / M  (the actor loss from the per-sample terms uavnet_actor_head_bwd wrote)
__global__ void __launch_bounds__(NET_THREADS) mean_rows_kernel(const float *__restrict__ rows, long long M, float inv_m,
                                                                float *__restrict__ out) {
    __shared__ float part[NET_THREADS / 32];
    float acc = 0.f;
    for (long long i = (long long)blockIdx.x * NET_THREADS + threadIdx.x; i < M; i += (long long)gridDim.x * NET_THREADS) acc += rows[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float tot = 0.f;
        for (int w = 0; w < NET_THREADS / 32; w++) tot += part[w];
        atomicAdd(out, tot * inv_m);
    }
}

// one rollout step's bookkeeping (main.py:199-211: ep_r += r; buffer_r.append(r)) in one launch
__global__ void __launch_bounds__(NET_THREADS) rollout_record_kernel(const double *__restrict__ r, const uint8_t *__restrict__ done,
                                                                     long long E, float *__restrict__ r_out,
                                                                     uint8_t *__restrict__ done_out, double *__restrict__ ep_return,
                                                                     double *__restrict__ ep_finished) {
    const long long e = (long long)blockIdx.x * NET_THREADS + threadIdx.x;
    if (e >= E) return;
    const double x = r[e];
    const uint8_t d = done[e];
    r_out[e] = (float)x;
    done_out[e] = d;
    if (ep_return) {
        const double tot = ep_return[e] + x;                     // ep_r += r (main.py:199)
        if (d) {                                                  // episode over: GLOBAL_RUNNING_R gets ep_r, ep_r = 0 (:188,246-252)
            if (ep_finished) ep_finished[e] = tot;
            ep_return[e] = 0.0;
        } else ep_return[e] = tot;
    }
}

// n-step value targets of the worker loop (main.py:217-227), one thread per env walking its T rewards backwards
__global__ void __launch_bounds__(NET_THREADS) nstep_targets_kernel(const float *__restrict__ r, const uint8_t *__restrict__ done,
                                                                    const float *__restrict__ v_boot, int T, long long E, float gamma,
                                                                    float *__restrict__ out) {
    const long long e = (long long)blockIdx.x * NET_THREADS + threadIdx.x;
    if (e >= E) return;
    float v = v_boot[e];
    for (int t = T - 1; t >= 0; t--) {
        v = r[t * E + e] + gamma * (done[t * E + e] ? 0.f : v);
        out[t * E + e] = v;
    }
}

// ---- per-device host state.  One process usually drives one GPU, but nothing here assumes it: every entry point
// switches to the device that owns its first pointer argument (like uavenv's use_device), and the SM count, the gemm
// error word and the function attributes are kept per device ordinal.
constexpr int MAX_DEVICES = 64;
struct DeviceState {
    int n_sm;                     // cudaDevAttrMultiProcessorCount (0 = not queried yet)
    unsigned int *gemm_err;       // device word, sticky: a CTA gave up on an mbarrier (protocol error)
    bool gemm_attr_done[18];
};
DeviceState g_dev[MAX_DEVICES];

// make the device that owns `p` current; returns its ordinal (or the current device if p is not a device pointer)
int use_device_of(const void *p, void *stream) {
    int cur = 0;
    if (cudaGetDevice(&cur) != cudaSuccess) { cudaGetLastError(); return 0; }
    static int n_dev = -1;
    if (n_dev < 0 && cudaGetDeviceCount(&n_dev) != cudaSuccess) { cudaGetLastError(); n_dev = 1; }
    // a stream that is being captured already lives on the right device; pointer queries are kept out of captures
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    if (n_dev > 1 && cudaStreamIsCapturing((cudaStream_t)stream, &cap) != cudaSuccess) { cudaGetLastError(); cap = cudaStreamCaptureStatusNone; }
    if (p && n_dev > 1 && cap == cudaStreamCaptureStatusNone) {
        cudaPointerAttributes at;
        if (cudaPointerGetAttributes(&at, p) == cudaSuccess) {
            if ((at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged) && at.device != cur && at.device >= 0) {
                if (cudaSetDevice(at.device) == cudaSuccess) cur = at.device;
            }
        } else cudaGetLastError();
    }
    return cur < MAX_DEVICES ? cur : 0;
}

int sm_count(int dev) {
    DeviceState &d = g_dev[dev];
    if (d.n_sm <= 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) { cudaGetLastError(); n = 148; }
        d.n_sm = n;
    }
    return d.n_sm;
}

// the device error word (bit 0: a uavnet_gemm CTA gave up on a barrier; bit 1: bounds-check build, index outside the table)
unsigned int *ensure_err_word(int dev) {
    DeviceState &d = g_dev[dev];
    if (!d.gemm_err) {
        if (cudaMalloc(&d.gemm_err, sizeof(unsigned int)) != cudaSuccess) { cudaGetLastError(); d.gemm_err = nullptr; return nullptr; }
        cudaMemset(d.gemm_err, 0, sizeof(unsigned int));
    }
    return d.gemm_err;
}

int grid_for(long long items, int dev) {
    long long g = (items + NET_THREADS - 1) / NET_THREADS;
    const long long cap = (long long)sm_count(dev) * 8 * 4;   // a few waves of 8 CTAs per SM; the kernels are grid-stride
    if (g > cap) g = cap;
    return (int)(g < 1 ? 1 : g);
}

bool aligned16(const void *p) { return ((uintptr_t)p & 15) == 0; }

}  // namespace uavk
using namespace uavk;

extern "C" {

int uavnet_sparse_fwd(const int32_t *idx, int64_t M, int32_t K, int64_t n_rows, const float *W, const float *b, int32_t H,
                      float *out, int32_t relu6, void *stream) {
    if (!idx || !W || !out || M < 1 || K < 1 || n_rows < 1 || H < 4 || (H & 3) || !aligned16(W) || !aligned16(out) ||
        (b && !aligned16(b)))
        return UAVNET_EINVAL;
    const int dev = use_device_of(idx, stream);
    sparse_fwd_kernel<<<grid_for(M * (H / 4), dev), NET_THREADS, 0, (cudaStream_t)stream>>>(
        idx, M, K, (const float4 *)W, (const float4 *)b, H / 4, (float4 *)out, relu6, (int)n_rows,
#ifdef UAVNET_BOUNDS_CHECK
        ensure_err_word(dev)
#else
        nullptr
#endif
    );
    return cudaGetLastError() == cudaSuccess ? UAVNET_OK : UAVNET_ECUDA;
}

int uavnet_sparse_bwd(const int32_t *idx, int64_t M, int32_t K, int64_t n_rows, const float *dpre, int32_t H, float *dW,
                      void *stream) {
    if (!idx || !dpre || !dW || M < 1 || K < 1 || n_rows < 1 || H < 4 || (H & 3) || !aligned16(dpre) || !aligned16(dW))
        return UAVNET_EINVAL;
    const int dev = use_device_of(idx, stream);
    sparse_bwd_kernel<<<grid_for(M * (H / 4), dev), NET_THREADS, 0, (cudaStream_t)stream>>>(idx, M, K, (const float4 *)dpre,
                                                                                    H / 4, (float4 *)dW);
    return cudaGetLastError() == cudaSuccess ? UAVNET_OK : UAVNET_ECUDA;
}

int64_t uavnet_sparse_bwd_gather_workspace(int64_t M, int32_t K, int64_t n_rows) {
    if (M < 1 || K < 1 || n_rows < 1) return 0;
    // count[n_rows] | cursor[n_rows] | sample_of[M * K] | row_of[M * K], int32 each, 16-byte aligned pieces
    auto pad = [](int64_t n) { return (n + 3) / 4 * 4; };
    return 4 * (2 * pad(n_rows) + 2 * pad(M * K));
}

// step 1 of the gather-side gradient: bucket the (sample, slot) pairs by row.  Depends on the indices only, so a caller can
// run it on a side stream as soon as the rollout's indices exist, under the backward products of the dense layers.
int uavnet_sparse_bwd_gather_prepare(const int32_t *idx, int64_t M, int32_t K, int64_t n_rows, void *workspace, void *stream) {
    if (!idx || !workspace || M < 1 || K < 1 || n_rows < 1 || n_rows > 0x7fffffffLL || M * K > 0x7fffffffLL || !aligned16(workspace))
        return UAVNET_EINVAL;
    const int dev = use_device_of(idx, stream);
    cudaStream_t st = (cudaStream_t)stream;
    auto pad = [](int64_t n) { return (n + 3) / 4 * 4; };
    int32_t *count = (int32_t *)workspace, *cursor = count + pad(n_rows), *sample_of = cursor + pad(n_rows), *row_of = sample_of + pad(M * K);
    if (cudaMemsetAsync(count, 0, (size_t)pad(n_rows) * 4, st) != cudaSuccess) { cudaGetLastError(); return UAVNET_ECUDA; }
    const long long n = M * K;
    row_hist_kernel<<<grid_for(n, dev), NET_THREADS, 0, st>>>(idx, n, (int)n_rows, count);
    row_scan_kernel<<<1, 1024, 0, st>>>(count, (int)n_rows, cursor);
    row_fill_kernel<<<grid_for(n, dev), NET_THREADS, 0, st>>>(idx, n, K, (int)n_rows, cursor, sample_of, row_of);
    return cudaGetLastError() == cudaSuccess ? UAVNET_OK : UAVNET_ECUDA;
}

// step 2: the slice-wise sums over the sorted list (after the fill cursor[r] = end of row r's bucket: the last row's cursor
// is the number of bucketed pairs)
int uavnet_sparse_bwd_gather_apply(int64_t M, int32_t K, int64_t n_rows, const float *dpre, int32_t H, float *dW, void *workspace,
                                   int32_t col_passes, void *stream) {
    if (!dpre || !dW || !workspace || M < 1 || K < 1 || n_rows < 1 || n_rows > 0x7fffffffLL || M * K > 0x7fffffffLL || H < 4 || (H & 3) ||
        !aligned16(dpre) || !aligned16(dW) || !aligned16(workspace) || col_passes < 1 || (H / 4) % col_passes != 0)
        return UAVNET_EINVAL;
    const int dev = use_device_of(dpre, stream);
    cudaStream_t st = (cudaStream_t)stream;
    auto pad = [](int64_t n) { return (n + 3) / 4 * 4; };
    int32_t *count = (int32_t *)workspace, *cursor = count + pad(n_rows), *sample_of = cursor + pad(n_rows), *row_of = sample_of + pad(M * K);
    const long long n = M * K;
    const int H4 = H / 4, nc4 = H4 / col_passes;
    const long long items = ((n + BWD_SLICE - 1) / BWD_SLICE) * nc4;
    for (int p = 0; p < col_passes; p++)
        sparse_bwd_gather_kernel<<<grid_for(items, dev), NET_THREADS, 0, st>>>(sample_of, row_of, cursor + (n_rows - 1), (const float4 *)dpre,
                                                                          H4, p * nc4, nc4, (float4 *)dW);
    return cudaGetLastError() == cudaSuccess ? UAVNET_OK : UAVNET_ECUDA;
}

// the sums for the columns [col0, col0 + n_cols) only (a caller that wants to do something between the column passes)
int uavnet_sparse_bwd_gather_apply_cols(int64_t M, int32_t K, int64_t n_rows, const float *dpre, int32_t H, int32_t col0, int32_t n_cols,
                                        float *dW, void *workspace, void *stream) {
    if (!dpre || !dW || !workspace || M < 1 || K < 1 || n_rows < 1 || n_rows > 0x7fffffffLL || M * K > 0x7fffffffLL || H < 4 || (H & 3) ||
        !aligned16(dpre) || !aligned16(dW) || !aligned16(workspace) || col0 < 0 || (col0 & 3) || n_cols < 4 || (n_cols & 3) || col0 + n_cols > H)
        return UAVNET_EINVAL;
    const int dev = use_device_of(dpre, stream);
    auto pad = [](int64_t n) { return (n + 3) / 4 * 4; };
    int32_t *count = (int32_t *)workspace, *cursor = count + pad(n_rows), *sample_of = cursor + pad(n_rows), *row_of = sample_of + pad(M * K);
    const long long n = M * K;
    const int nc4 = n_cols / 4;
    const long long items = ((n + BWD_SLICE - 1) / BWD_SLICE) * nc4;
    sparse_bwd_gather_kernel<<<grid_for(items, dev), NET_THREADS, 0, (cudaStream_t)stream>>>(sample_of, row_of, cursor + (n_rows - 1),
                                                                                        (const float4 *)dpre, H / 4, col0 / 4, nc4, (float4 *)dW);
    return cudaGetLastError() == cudaSuccess ? UAVNET_OK : UAVNET_ECUDA;
}

int uavnet_sparse_bwd_gather(const int32_t *idx, int64_t M, int32_t K, int64_t n_rows, const float *dpre, int32_t H, float *dW,
                             void *workspace, int32_t col_passes, void *stream) {
    if (!idx || !dpre || !dW || !workspace || H < 4 || (H & 3) || col_passes < 1 || (H / 4) % col_passes != 0) return UAVNET_EINVAL;
    const int rc = uavnet_sparse_bwd_gather_prepare(idx, M, K, n_rows, workspace, stream);
    return rc ? rc : uavnet_sparse_bwd_gather_apply(M, K, n_rows, dpre, H, dW, workspace, col_passes, stream);
}

int uavnet_actor_head_bwd(const float *prob, const int64_t *a_his, const float *td, int64_t M, int32_t A, float beta,
                          float *dz, int64_t ldz, float *loss_row, void *stream) {
    if (!prob || !a_his || !td || !dz || M < 1 || A < 1 || ldz < A) return UAVNET_EINVAL;
    const int dev = use_device_of(prob, stream);
    const int grid = grid_for(M * 32, dev);
    const float inv_m = 1.0f / (float)M;
    const long long *ah = (const long long *)a_his;
    cudaStream_t st = (cudaStream_t)stream;
    if (A <= 256) actor_head_bwd_kernel<8><<<grid, NET_THREADS, 0, st>>>(prob, ah, td, M, A, beta, inv_m, dz, ldz, loss_row);
    else if (A <= 640) actor_head_bwd_kernel<20><<<grid, NET_THREADS, 0, st>>>(prob, ah, td, M, A, beta, inv_m, dz, ldz, loss_row);
    else if (A <= 1024) actor_head_bwd_kernel<32><<<grid, NET_THREADS, 0, st>>>(prob, ah, td, M, A, beta, inv_m, dz, ldz, loss_row);
    else actor_head_bwd_kernel<0><<<grid, NET_THREADS, 0, st>>>(prob, ah, td, M, A, beta, inv_m, dz, ldz, loss_row);
    return cudaGetLastError() == cudaSuccess ? UAVNET_OK : UAVNET_ECUDA;
}

int uavnet_p2p_alloc(int64_t bytes, void **dev_ptr, uint8_t *handle64) {
    if (bytes < 16 || !dev_ptr || !handle64) return UAVNET_EINVAL;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    void *p = nullptr;
    if (cudaMalloc(&p, (size_t)bytes) != cudaSuccess) { cudaGetLastError(); return UAVNET_ECUDA; }
    cudaIpcMemHandle_t hdl;
    if (cudaMemset(p, 0, (size_t)bytes) != cudaSuccess || cudaIpcGetMemHandle(&hdl, p) != cudaSuccess) {
        cudaGetLastError();
        cudaFree(p);
        return UAVNET_ECUDA;
    }
    memcpy(handle64, &hdl, 64);
    *dev_ptr = p;
    return UAVNET_OK;
}

int uavnet_p2p_open(const uint8_t *handle64, void **dev_ptr) {
    if (!handle64 || !dev_ptr) return UAVNET_EINVAL;
    cudaIpcMemHandle_t hdl;
    memcpy(&hdl, handle64, 64);
    if (cudaIpcOpenMemHandle(dev_ptr, hdl, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); return UAVNET_ECUDA; }
    return UAVNET_OK;
}

int uavnet_p2p_close(void *dev_ptr) {
    if (!dev_ptr) return UAVNET_EINVAL;
    return cudaIpcCloseMemHandle(dev_ptr) == cudaSuccess ? UAVNET_OK : UAVNET_ECUDA;
}

int uavnet_p2p_free(void *dev_ptr) {
    if (!dev_ptr) return UAVNET_EINVAL;
    return cudaFree(dev_ptr) == cudaSuccess ? UAVNET_OK : UAVNET_ECUDA;
}

int uavnet_p2p_rmsprop(float *const *grads, float *const *params, float *ms_local, int64_t n, int32_t rank, int32_t world,
                       float lr, float decay, float eps, void *stream) {
    if (!grads || !params || !ms_local || n < 4 || (n & 3) || world < 1 || world > UAVNET_MAX_PEERS || rank < 0 || rank >= world)
        return UAVNET_EINVAL;
    PeerPtrs pp;
    memset(&pp, 0, sizeof(pp));
    for (int r = 0; r < world; r++) {
        if (!grads[r] || !params[r] || !aligned16(grads[r]) || !aligned16(params[r])) return UAVNET_EINVAL;
        pp.g[r] = grads[r];
        pp.p[r] = params[r];
    }
    const long long n4 = n >> 2, per4 = (n4 + world - 1) / world;
    const long long lo4 = (long long)rank * per4, hi4 = lo4 + per4 < n4 ? lo4 + per4 : n4;
    if (lo4 >= hi4) return UAVNET_OK;
    const int dev = use_device_of(ms_local, stream);
    p2p_rmsprop_kernel<<<grid_for(hi4 - lo4, dev), NET_THREADS, 0, (cudaStream_t)stream>>>(pp, ms_local, lo4, hi4, rank, world, lr,
                                                                                     decay, eps, 1.0f / (float)world);
    return cudaGetLastError() == cudaSuccess ? UAVNET_OK : UAVNET_ECUDA;
}

static int p2p_push_regions(float *const *grads, float *const *params, uint32_t *const *flags, float *ms_local, const PushRegions &R,
                            int32_t rank, int32_t world, float lr, float decay, float eps, void *stream) {
    PeerPtrs2 pp;
    memset(&pp, 0, sizeof(pp));
    for (int r = 0; r < world; r++) {
        if (!grads[r] || !params[r] || !flags[r] || !aligned16(grads[r]) || !aligned16(params[r])) return UAVNET_EINVAL;
        pp.g[r] = grads[r]; pp.p[r] = params[r]; pp.f[r] = flags[r];
    }
    const long long n4 = R.a.n4 + R.b.n4, per4 = (n4 + world - 1) / world;
    const long long lo4 = (long long)rank * per4 < n4 ? (long long)rank * per4 : n4, hi = lo4 + per4 < n4 ? lo4 + per4 : n4;
    const int dev = use_device_of(ms_local, stream);
    const long long mine4 = hi > lo4 ? hi - lo4 : 1;
    // Two blocks per SM: measured on 8 B200s (80.8 MB, profiles/r2/NOTES.md) 64 / 148 / 296 / 592 / 1184 / 2368 blocks take
    // 0.336 / 0.321 / 0.316 / 0.360 / 0.387 / 0.401 ms -- NVLink likes few fat streams of peer loads and stores better than
    // many thin ones.  (Every block also spins on the flags first, so the grid has to stay within a wave anyway.)
    int grid = grid_for(mine4, dev);
    const int one_wave = sm_count(dev) * 8;
    if (grid > 2 * sm_count(dev)) grid = 2 * sm_count(dev);
    const float gs = 1.0f / (float)world;
    cudaStream_t st = (cudaStream_t)stream;
    // timing experiments only (results are wrong when set): bit 0 = no peer stores, bit 1 = no peer loads
    static int dbg = -1;
    if (dbg < 0) { const char *ev = getenv("UAVNET_P2P_DBG"); dbg = ev ? atoi(ev) : 0; }
    if (const char *ev = getenv("UAVNET_P2P_GRID")) { const int gmax = atoi(ev); if (gmax > 0) grid = gmax; }
    switch (world) {
        case 1: p2p_push_kernel<1><<<grid, NET_THREADS, 0, st>>>(pp, ms_local, R, lo4, hi, rank, world, lr, decay, eps, gs, dbg); break;
        case 2: p2p_push_kernel<2><<<grid, NET_THREADS, 0, st>>>(pp, ms_local, R, lo4, hi, rank, world, lr, decay, eps, gs, dbg); break;
        case 4: p2p_push_kernel<4><<<grid, NET_THREADS, 0, st>>>(pp, ms_local, R, lo4, hi, rank, world, lr, decay, eps, gs, dbg); break;
        case 8: p2p_push_kernel<8><<<grid, NET_THREADS, 0, st>>>(pp, ms_local, R, lo4, hi, rank, world, lr, decay, eps, gs, dbg); break;
        default: p2p_push_kernel<0><<<grid, NET_THREADS, 0, st>>>(pp, ms_local, R, lo4, hi, rank, world, lr, decay, eps, gs, dbg); break;
    }
    if (cudaGetLastError() != cudaSuccess) return UAVNET_ECUDA;
    int grid2 = grid_for(n4, dev);
    if (grid2 > one_wave) grid2 = one_wave;
    p2p_finish_kernel<<<grid2, NET_THREADS, 0, st>>>(flags[rank], (float4 *)grads[rank], R, lo4, hi, world);
    return cudaGetLastError() == cudaSuccess ? UAVNET_OK : UAVNET_ECUDA;
}

int uavnet_p2p_push(float *const *grads, float *const *params, uint32_t *const *flags, float *ms_local, int64_t n, int32_t rank,
                    int32_t world, float lr, float decay, float eps, void *stream) {
    if (!grads || !params || !flags || !ms_local || n < 4 || (n & 3) || world < 1 || world > UAVNET_MAX_PEERS || rank < 0 || rank >= world)
        return UAVNET_EINVAL;
    PushRegions R;
    memset(&R, 0, sizeof(R));
    R.a.base4 = 0; R.a.n4 = n >> 2; R.a.row_w4 = R.a.sub_w4 = 1;
    R.b.row_w4 = R.b.sub_w4 = 1;
    return p2p_push_regions(grads, params, flags, ms_local, R, rank, world, lr, decay, eps, stream);
}

int uavnet_p2p_push_part(float *const *grads, float *const *params, uint32_t *const *flags, float *ms_local, const uavnet_push_part *a,
                         const uavnet_push_part *b, int32_t rank, int32_t world, float lr, float decay, float eps, void *stream) {
    if (!grads || !params || !flags || !ms_local || !a || world < 1 || world > UAVNET_MAX_PEERS || rank < 0 || rank >= world) return UAVNET_EINVAL;
    PushRegions R;
    memset(&R, 0, sizeof(R));
    const uavnet_push_part *src[2] = {a, b};
    PushRegion *dst[2] = {&R.a, &R.b};
    for (int k = 0; k < 2; k++) {
        dst[k]->row_w4 = dst[k]->sub_w4 = 1;
        const uavnet_push_part *q = src[k];
        if (!q) continue;
        if (q->offset < 0 || (q->offset & 3) || q->rows < 1 || q->row_width < 4 || (q->row_width & 3) || q->col0 < 0 || (q->col0 & 3) ||
            q->n_cols < 4 || (q->n_cols & 3) || q->col0 + q->n_cols > q->row_width)
            return UAVNET_EINVAL;
        dst[k]->base4 = q->offset >> 2; dst[k]->row_w4 = q->row_width >> 2; dst[k]->col0_4 = q->col0 >> 2; dst[k]->sub_w4 = q->n_cols >> 2;
        dst[k]->n4 = (long long)q->rows * (q->n_cols >> 2);
    }
    return p2p_push_regions(grads, params, flags, ms_local, R, rank, world, lr, decay, eps, stream);
}

int uavnet_p2p_push_status(const uint32_t *flags_own, uint32_t *epoch_out, uint32_t *timeout_out) {
    if (!flags_own) return UAVNET_EINVAL;
    uint32_t h[20];
    if (cudaMemcpy(h, flags_own, sizeof(h), cudaMemcpyDeviceToHost) != cudaSuccess) { cudaGetLastError(); return UAVNET_ECUDA; }
    if (epoch_out) *epoch_out = h[PF_EPOCH];
    if (timeout_out) *timeout_out = h[PF_TIMEOUT];
    return UAVNET_OK;
}

int uavnet_softmax_sample(const float *logits, int64_t M, int32_t A, uint64_t seed, uint32_t row_offset,
                          const uint32_t *counter_dev, uint32_t counter_add, float *prob, int64_t *action, void *stream) {
    if (!logits || M < 1 || A < 1 || (!prob && !action)) return UAVNET_EINVAL;
    const int dev = use_device_of(logits, stream);
    const uint32_t s0 = (uint32_t)seed, s1 = (uint32_t)(seed >> 32);
    const int grid = grid_for(M * 32, dev);
    cudaStream_t st = (cudaStream_t)stream;
    long long *act = (long long *)action;
    if (A <= 256) softmax_sample_kernel<8><<<grid, NET_THREADS, 0, st>>>(logits, M, A, s0, s1, row_offset, counter_dev, counter_add, prob, act);
    else if (A <= 640) softmax_sample_kernel<20><<<grid, NET_THREADS, 0, st>>>(logits, M, A, s0, s1, row_offset, counter_dev, counter_add, prob, act);
    else softmax_sample_kernel<0><<<grid, NET_THREADS, 0, st>>>(logits, M, A, s0, s1, row_offset, counter_dev, counter_add, prob, act);
    return cudaGetLastError() == cudaSuccess ? UAVNET_OK : UAVNET_ECUDA;
}

int uavnet_rank1_mask(const float *dv, const float *w, const float *h, int64_t M, int32_t H, float *out, void *stream) {
    if (!dv || !w || !h || !out || M < 1 || H < 4 || (H & 3) || !aligned16(w) || !aligned16(h) || !aligned16(out)) return UAVNET_EINVAL;
    const int dev = use_device_of(dv, stream);
    rank1_mask_kernel<<<grid_for(M * (H / 4), dev), NET_THREADS, 0, (cudaStream_t)stream>>>(dv, (const float4 *)w, (const float4 *)h, M,
                                                                                     H / 4, (float4 *)out);
    return cudaGetLastError() == cudaSuccess ? UAVNET_OK : UAVNET_ECUDA;
}

int uavnet_critic_td(const float *v_target, const float *v, int64_t M, float *td, float *dv, float *loss2, void *stream) {
    if (!v_target || !v || !td || !dv || !loss2 || M < 1) return UAVNET_EINVAL;
    const int dev = use_device_of(v, stream);
    cudaStream_t st = (cudaStream_t)stream;
    if (cudaMemsetAsync(loss2, 0, 2 * sizeof(float), st) != cudaSuccess) { cudaGetLastError(); return UAVNET_ECUDA; }
    int grid = grid_for(M, dev);
    if (grid > 2 * sm_count(dev)) grid = 2 * sm_count(dev);
    critic_td_kernel<<<grid, NET_THREADS, 0, st>>>(v_target, v, M, 1.0f / (float)M, td, dv, loss2);
    return cudaGetLastError() == cudaSuccess ? UAVNET_OK : UAVNET_ECUDA;
}

int uavnet_mean_rows(const float *rows, int64_t M, float *out_accum, void *stream) {
    if (!rows || !out_accum || M < 1) return UAVNET_EINVAL;
    const int dev = use_device_of(rows, stream);
    int grid = grid_for(M, dev);
    if (grid > 2 * sm_count(dev)) grid = 2 * sm_count(dev);
    mean_rows_kernel<<<grid, NET_THREADS, 0, (cudaStream_t)stream>>>(rows, M, 1.0f / (float)M, out_accum);
    return cudaGetLastError() == cudaSuccess ? UAVNET_OK : UAVNET_ECUDA;
}

int uavnet_rollout_record(const double *reward, const uint8_t *done, int64_t E, float *reward_out, uint8_t *done_out,
                          double *ep_return, double *ep_finished, void *stream) {
    if (!reward || !done || !reward_out || !done_out || E < 1) return UAVNET_EINVAL;
    use_device_of(reward, stream);
    rollout_record_kernel<<<(unsigned)((E + NET_THREADS - 1) / NET_THREADS), NET_THREADS, 0, (cudaStream_t)stream>>>(
        reward, done, E, reward_out, done_out, ep_return, ep_finished);
    return cudaGetLastError() == cudaSuccess ? UAVNET_OK : UAVNET_ECUDA;
}

int uavnet_nstep_targets(const float *rewards, const uint8_t *dones, const float *v_boot, int32_t T, int64_t E, float gamma,
                         float *out, void *stream) {
    if (!rewards || !dones || !v_boot || !out || T < 1 || E < 1) return UAVNET_EINVAL;
    use_device_of(rewards, stream);
    nstep_targets_kernel<<<(unsigned)((E + NET_THREADS - 1) / NET_THREADS), NET_THREADS, 0, (cudaStream_t)stream>>>(
        rewards, dones, v_boot, T, E, gamma, out);
    return cudaGetLastError() == cudaSuccess ? UAVNET_OK : UAVNET_ECUDA;
}

int uavnet_rmsprop(float *param, float *grad, float *ms, int64_t n, float lr, float decay, float eps, float grad_scale,
                   int32_t zero_grad, void *stream) {
    if (!param || !grad || !ms || n < 1 || !aligned16(param) || !aligned16(grad) || !aligned16(ms)) return UAVNET_EINVAL;
    const int dev = use_device_of(param, stream);
    rmsprop_kernel<<<grid_for((n + 3) / 4, dev), NET_THREADS, 0, (cudaStream_t)stream>>>(param, grad, ms, n, lr, decay, eps,
                                                                                   grad_scale, zero_grad);
    return cudaGetLastError() == cudaSuccess ? UAVNET_OK : UAVNET_ECUDA;
}

// ---- dense layers on the tensor cores (tc_gemm.cuh) ----
typedef CUresult (*tmap_encode_t)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// 2-D map of a row-major float32 matrix [rows, k] (leading dimension ld elements): boxes of 32 k x box_rows rows,
// 128-byte swizzle, out-of-range elements read as zero (partial k chunks and partial row tiles need no special case)
static bool make_tmap(tmap_encode_t encode, CUtensorMap *map, const float *base, long long k, long long rows, long long ld, int box_rows) {
    if (!encode || k < 1 || rows < 1 || box_rows < 1 || box_rows > 256 || (ld * 4) % 16 != 0) return false;
    const cuuint64_t gdim[2] = {(cuuint64_t)k, (cuuint64_t)rows};
    const cuuint64_t gstride[1] = {(cuuint64_t)ld * 4};
    const cuuint32_t box[2] = {(cuuint32_t)tc::KC, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    return encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(base), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}


int uavnet_gemm(const uavnet_gemm_desc *d, void *stream) {
    if (!d || !d->B || d->M < 0 || d->N < 1 || d->K < 1 || d->N > 65536) return UAVNET_EINVAL;
    if (d->M > 0 && !d->A) return UAVNET_EINVAL;
    if (!d->D && !d->colsum && !d->dot_out) return UAVNET_EINVAL;
    if (d->M == 0 && !d->colsum) return UAVNET_EINVAL;
    if ((d->D && d->ldd < d->N) || (d->mask_src && d->ld_mask < d->N)) return UAVNET_EINVAL;
    if (d->a_trans ? d->lda < d->M : d->lda < d->K) { if (d->M > 0) return UAVNET_EINVAL; }
    if (d->b_trans ? d->ldb < d->K : d->ldb < d->N) return UAVNET_EINVAL;
    if (d->split_k < 0 || (d->split_k > 1 && !d->accumulate)) return UAVNET_EINVAL;
    if (d->colsum && !d->accumulate) return UAVNET_EINVAL;                  // the ones row accumulates like the rest
    if (d->accumulate && (d->bias || d->relu6 || d->dot_out)) return UAVNET_EINVAL;
    if (d->dot_out && (!d->dot_w || d->N > 256)) return UAVNET_EINVAL;
    if (d->out_colsum && (d->accumulate || !d->D)) return UAVNET_EINVAL;
    if (d->precision != UAVNET_GEMM_TF32 && d->precision != UAVNET_GEMM_3XTF32) return UAVNET_EINVAL;
    const bool p3 = d->precision == UAVNET_GEMM_3XTF32;
    const int dev = use_device_of(d->B, stream);
    DeviceState &ds = g_dev[dev];
    tc::GemmArgs g;
    memset(&g, 0, sizeof(g));
    g.A = d->A; g.lda = d->lda; g.a_trans = d->a_trans ? 1 : 0;
    g.B = d->B; g.ldb = d->ldb; g.b_trans = d->b_trans ? 1 : 0;
    g.D = d->D; g.ldd = d->ldd; g.M = d->M; g.N = d->N; g.K = d->K;
    g.bias = d->bias; g.relu6 = d->relu6; g.mask_src = d->mask_src; g.ld_mask = d->ld_mask; g.accumulate = d->accumulate;
    g.colsum = d->colsum; g.out_colsum = d->out_colsum; g.dot_w = d->dot_w; g.dot_b = d->dot_b; g.dot_out = d->dot_out;
    static int bn_max_env = -1;                                              // tuning experiments only
    if (bn_max_env < 0) { const char *e = getenv("UAVNET_GEMM_BN_MAX"); bn_max_env = e ? atoi(e) : 0; }
    const long long rows = d->M + (d->colsum ? 1 : 0);
    const long long tiles_m = (rows + tc::BM - 1) / tc::BM;
    // N tile: up to 256 columns (one tile for N = 200, three of 208 for N = 625) -- except for products with many row tiles
    // (the update's 81 920-row data gradients): there 112-column tiles measured 15 % faster on the 625-deep product (121 vs
    // 142 us: three pipeline stages instead of two inside the same shared-memory budget, twice the CTAs to overlap
    // epilogues with main loops).  The value-head epilogue needs the whole row in one tile.
    int bn_max = 256;
    if (tiles_m >= 2 * sm_count(dev) && !d->dot_out && !d->accumulate && d->N <= 256) bn_max = 112;
    if (bn_max_env >= 16 && bn_max_env <= 256) bn_max = bn_max_env / 16 * 16;
    int tiles_n = (d->N + bn_max - 1) / bn_max;
    g.tiles_n = tiles_n;
    g.BN = (((d->N + tiles_n - 1) / tiles_n) + 15) / 16 * 16;
    g.tmem_cols = 32;
    while (g.tmem_cols < g.BN) g.tmem_cols <<= 1;
    const int stage_bytes = (tc::A_TILE_BYTES + g.BN * tc::KC * 4) * (p3 ? 2 : 1);
    static int budget_kb = -1;                                               // tuning experiments only
    if (budget_kb < 0) { const char *e = getenv("UAVNET_GEMM_BUDGET_KB"); budget_kb = e ? atoi(e) : 0; }
    int budget = p3 ? tc::SMEM_BUDGET_3X : tc::SMEM_BUDGET_1X;
    if (budget_kb >= 64 && budget_kb <= 220) budget = budget_kb * 1024;
    g.stages = budget / stage_bytes;
    if (g.stages > tc::MAX_STAGES) g.stages = tc::MAX_STAGES;
    if (g.stages < 2) return UAVNET_EINVAL;
    int smem = g.stages * stage_bytes;
    if (smem < tc::EPI_BYTES) smem = tc::EPI_BYTES;
    const long long chunks = (d->K + tc::KC - 1) / tc::KC;
    long long split = d->split_k;
    if (!d->accumulate) split = 1;
    else if (split == 0) {                          // fill the GPU about twice
        split = (2 * sm_count(dev)) / (tiles_m * tiles_n);
        if (split < 1) split = 1;
    }
    if (split > chunks) split = chunks;
    const long long cps = (chunks + split - 1) / split;
    g.k_per_split = cps * tc::KC;
    g.split_k = (int)((chunks + cps - 1) / cps);
    const long long grid = tiles_m * tiles_n * g.split_k;
    if (grid < 1 || grid > 0x7fffffffLL) return UAVNET_EINVAL;
    g.a_vec = d->A && aligned16(d->A) && (d->lda % 4 == 0);
    g.b_vec = aligned16(d->B) && (d->ldb % 4 == 0);
    static int dbg = -1;
    if (dbg < 0) { const char *e = getenv("UAVNET_GEMM_DBG"); dbg = e ? atoi(e) : 0; }
    g.dbg = dbg;
    g.err = ensure_err_word(dev);
    if (!g.err) return UAVNET_ECUDA;
    // staging mode per operand: TMA boxes for row-major, 16-byte aligned operands (one MMA per k-step only: the hi/lo
    // split of 3xTF32 needs the data in registers), threads otherwise
    static int no_tma = -1;
    if (no_tma < 0) { const char *e = getenv("UAVNET_GEMM_NO_TMA"); no_tma = (e && atoi(e)) ? 1 : 0; }
    static tmap_encode_t encode = nullptr;
    if (!encode && !no_tma) {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn) {
            cudaGetLastError();
            no_tma = 1;
        } else {
            encode = (tmap_encode_t)fn;
        }
    }
    int am = g.a_trans ? tc::OP_TRANSPOSED : tc::OP_ROWMAJOR, bm = g.b_trans ? tc::OP_ROWMAJOR : tc::OP_TRANSPOSED;
    if (!p3 && !no_tma) {
        if (am == tc::OP_ROWMAJOR && g.a_vec && !d->colsum && d->M > 0 &&
            make_tmap(encode, &g.tmap_a, d->A, d->K, d->M, d->lda, tc::BM)) am = tc::OP_TMA;
        if (bm == tc::OP_ROWMAJOR && g.b_vec && make_tmap(encode, &g.tmap_b, d->B, d->K, d->N, d->ldb, g.BN)) bm = tc::OP_TMA;
    }
    typedef void (*kern_t)(const tc::GemmArgs);
#define UAVK_G(P, A_, B_) tc::gemm_kernel<P, A_, B_>
    static const kern_t kerns[18] = {
        UAVK_G(false, 0, 0), UAVK_G(false, 0, 1), UAVK_G(false, 0, 2), UAVK_G(false, 1, 0), UAVK_G(false, 1, 1), UAVK_G(false, 1, 2),
        UAVK_G(false, 2, 0), UAVK_G(false, 2, 1), UAVK_G(false, 2, 2),
        UAVK_G(true, 0, 0), UAVK_G(true, 0, 1), nullptr, UAVK_G(true, 1, 0), UAVK_G(true, 1, 1), nullptr, nullptr, nullptr, nullptr};
#undef UAVK_G
    const int which = (p3 ? 9 : 0) + am * 3 + bm;
    if (!kerns[which]) return UAVNET_EINVAL;
    if (!ds.gemm_attr_done[which]) {
        if (cudaFuncSetAttribute(kerns[which], cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024) != cudaSuccess) {
            cudaGetLastError();
            return UAVNET_ECUDA;
        }
        ds.gemm_attr_done[which] = true;
    }
    kerns[which]<<<(unsigned)grid, tc::NTHR, smem, (cudaStream_t)stream>>>(g);
    return cudaGetLastError() == cudaSuccess ? UAVNET_OK : UAVNET_ECUDA;
}

int uavnet_gemm_check(void) {
    // any device this process launched uavnet_gemm on
    int cur = 0, bad = 0;
    cudaGetDevice(&cur);
    for (int dv = 0; dv < MAX_DEVICES; dv++) {
        if (!g_dev[dv].gemm_err) continue;
        unsigned int v = 0;
        if (cudaSetDevice(dv) != cudaSuccess || cudaMemcpy(&v, g_dev[dv].gemm_err, sizeof(v), cudaMemcpyDeviceToHost) != cudaSuccess) {
            cudaGetLastError();
            cudaSetDevice(cur);
            return UAVNET_ECUDA;
        }
        bad |= (int)v;
    }
    cudaSetDevice(cur);
    return bad;
}

}  // extern "C"
