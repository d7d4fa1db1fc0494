// The dense layers of the actor-critic MLP (main.py:148-149,152-153: 200x200, 200x625, 200x1 and their gradients) on
// the 5th-generation tensor cores: D[M,N] (+)= op(A) . op(B) with tcgen05.mma kind::tf32, accumulators in TMEM, and a
// fused epilogue (bias, relu6, relu6' mask, row dot = the critic's value head, column sums = bias gradients, split-K
// accumulation with float REDs).  sm_100a only.
//
// Shape of the problem: K is 200 or 625 for the forward / data-gradient GEMMs and 81 920 (the rollout batch) for the
// weight gradients; N is 200, 625 or 1.  Operand staging is chosen per operand (OP_* below):
//   * row-major, 16-byte aligned sources (activations, dz, weights kept as [N,K]) arrive by TMA: 2-D tensor maps, boxes
//     of 32 x 128 / 32 x BN floats with the 128-byte swizzle, one thread issues the copies and the MMAs;
//   * everything TMA cannot take (625-float rows, 4-byte aligned slices, the transposed reads X^T . dY of the weight
//     gradients, the hi/lo split of 3xTF32) is staged by the 256 threads: global -> registers (next chunk in flight) ->
//     shared memory in the canonical no-swizzle K-major UMMA layout (8 x 16-byte core matrices; transposed sources are
//     transposed on the way in), fence.proxy.async, then one elected thread issues the MMAs of the chunk.
// Either way the MMAs of a chunk are committed to the stage's mbarrier (tcgen05.commit), which frees the stage.
//
// fp32 accuracy: kind::tf32 reads 10 mantissa bits of every operand.  PREC3X stages every tile twice -- hi = the top 19
// bits, lo = a - hi -- and issues hi.hi + hi.lo + lo.hi into the same accumulator (3xTF32: ~2e-7 relative, what the
// parity tests of the learner need); the plain mode is one MMA per k-step.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace uavk {
namespace tc {

constexpr int BM = 128;            // UMMA M (one accumulator row per TMEM lane)
constexpr int KC = 32;             // reduction elements per stage = 4 MMAs of K = 8
constexpr int NTHR = 256;
constexpr int MAX_STAGES = 4;
constexpr int A_TILE_BYTES = BM * KC * 4;
constexpr int SMEM_BUDGET_3X = 200 * 1024;       // one CTA per SM
constexpr int SMEM_BUDGET_1X = 100 * 1024;       // two CTAs per SM: one CTA's epilogue overlaps the other's main loop
constexpr int EPI_BYTES = (NTHR / 32) * 32 * 33 * 4;       // per-warp 32 x 33 transposition buffers

struct GemmArgs {
    const float *A; long long lda; int a_trans;      // a_trans = 0: A[M,K] row-major; 1: stored [K,M] row-major
    const float *B; long long ldb; int b_trans;      // b_trans = 0: B[K,N] row-major; 1: stored [N,K] row-major
    float *D; long long ldd;
    long long M; int N; long long K;
    int BN, tiles_n, stages, tmem_cols;
    int split_k; long long k_per_split;
    const float *bias; int relu6;
    const float *mask_src; long long ld_mask;
    int accumulate;
    float *colsum;
    float *out_colsum;                                // += column sums of the stored D (after bias / relu6 / mask)
    const float *dot_w; const float *dot_b; float *dot_out;
    int a_vec, b_vec;                                 // 16-byte loads legal for the operand
    unsigned int *err;
    int dbg;                                          // UAVNET_GEMM_DBG, 0 in production; bit 1: MMA-rate probe (no operand loads)
    alignas(64) CUtensorMap tmap_a;                   // operand mode TMA: [rows, K] float32, box 32 x 128 (A) / 32 x BN (B),
    alignas(64) CUtensorMap tmap_b;                   // 128-byte swizzle
};

// operand staging modes
constexpr int OP_ROWMAJOR = 0;     // threads, source [rows, k]
constexpr int OP_TRANSPOSED = 1;   // threads, source [k, rows], transposed on the way into shared memory
constexpr int OP_TMA = 2;          // cp.async.bulk.tensor boxes of a [rows, k] source (16-byte aligned, ld % 4 == 0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}

__device__ __forceinline__ uint32_t mbar_try(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok;
}

// Bounded wait: a protocol error must not hang the GPU.  After ~2 s the CTA gives up (sticky), the error word is set
// and the host entry point's caller finds it with uavnet_gemm_check.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, volatile int *dead, unsigned int *err) {
    if (mbar_try(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try(bar, parity)) {
        __nanosleep(40);
        if (*dead) return;
        if (clock64() - t0 > 4000000000LL) {
            *dead = 1;
            if (err) atomicOr(err, 1u);
            return;
        }
    }
}

// Shared-memory matrix descriptor, no swizzle (cute::UMMA::SmemDescriptor): start address, leading-dimension byte
// offset (between the core matrices along K), stride byte offset (between the core matrices along M/N), version 1.
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}

// K-major tile as TMA writes it with CU_TENSOR_MAP_SWIZZLE_128B: 128-byte rows, 16-byte pieces XOR-ed with (row & 7);
// 8-row groups 1024 bytes apart (SBO), layout type 2, the leading offset is not used (1, as CUTLASS encodes it).
// A k-step of 8 tf32 advances the start address by 32 bytes inside the row.
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t saddr) {
    uint64_t d = (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, int c0, int c1, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar)
                 : "memory");
}

// cute::UMMA::InstrDescriptor for kind::tf32: D = F32 (bits 4-5 = 1), A/B format TF32 (bits 7-9, 10-12 = 2), a_major bit
// 15, b_major bit 16 (1 = MN-major), N >> 3 at bits 17-22, M >> 4 at bits 24-28.
__device__ __forceinline__ uint32_t instr_desc(int a_mn, int b_mn, int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(n >> 3) << 17) |
           ((uint32_t)(BM >> 4) << 24);
}

__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

__device__ __forceinline__ void mma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float *v) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int j = 0; j < 32; j++) v[j] = __uint_as_float(r[j]);
}

// ---- operand staging -------------------------------------------------------------------------------------------------
// A tile of R rows (R = 128 for A, BN for B; the M/N index of the GEMM) x KC reduction elements is R * 8 float4 items;
// thread t takes items t, t + 256, ...  Shared memory always holds the K-major no-swizzle layout: byte offset of (r, k) =
//   (r >> 3) * 1024 + (k >> 2) * 128 + (r & 7) * 16 + (k & 3) * 4          (LBO = 128, SBO = 1024)
// (kind::tf32 with MN-major descriptors returned zeros on B200 in every variant tried, profiles/gemm_debug.py; the
// transposition therefore happens on the way into shared memory.)
//
// Row-major source [rows, k] (k contiguous): a float4 is 4 consecutive k of one row = one 16-byte piece of a core
// matrix; every quarter-warp writes 128 contiguous bytes and reads 64 contiguous bytes per global row.
// Transposed source [k, rows] (rows contiguous): a float4 is 4 consecutive rows of one k.  A warp item covers 16 k x
// 8 rows (one 32-byte sector per k, fully used); the four components go out as scalar stores in an order rotated by
// lane >> 3, which makes the 32 lanes of every store hit 32 different banks.
// In both mappings a thread's item j is its item 0 moved down by 32 rows (4096 bytes of shared memory).
template <bool TRANS>
struct Stager {
    // hot path (full chunks of vector-loadable operands): one predicated LDG.128 per item from base + k0 * kstride +
    // j * jstride, one STS.128 (or four rotated scalar STS) at immediate offsets
    const float *base;     // (row0, kk) of the source
    long long jstride;     // elements between a thread's items (32 rows)
    long long kstride;     // elements per unit of k
    uint32_t valid;        // bit j: item j is a plain in-bounds vector load
    uint32_t special;      // bit j: item j exists but needs the careful path (partial vector, appended row of ones)
    uint32_t so[4];        // shared-memory byte offsets of item 0: [0] for row-major sources, one per rotated component else
    int rot;
    // careful path
    const float *src;
    long long ld, row0, rows_valid, ones_row;
    int kk, n_items, vec_ok;

    __device__ __forceinline__ void init(const float *s, long long ld_, int vec, int R, long long r0, long long rv, long long ones, int tid) {
        src = s; ld = ld_; vec_ok = vec; rows_valid = rv; ones_row = ones;
        const int l = tid & 31, w = tid >> 5;
        n_items = (R * 8 - tid + NTHR - 1) / NTHR;
        if (n_items < 0) n_items = 0;
        int rlo = 0;
        if (!TRANS) {
            const int r_in = l & 7, q = ((w & 1) << 2) | (l >> 3), g = w >> 1;
            row0 = r0 + g * 8 + r_in;
            kk = q * 4;
            rot = 0;
            so[0] = so[1] = so[2] = so[3] = (uint32_t)(g * 1024 + q * 128 + r_in * 16);
            base = s + row0 * ld_ + kk;
            jstride = 32 * ld_;
            kstride = 1;
        } else {
            const int t = l >> 3, oct = w >> 1;
            kk = ((w & 1) << 4) | (t << 2) | (l & 3);
            rlo = ((l >> 2) & 1) << 2;
            rot = t;
            row0 = r0 + oct * 8 + rlo;
            const uint32_t o = (uint32_t)(oct * 1024 + (kk >> 2) * 128 + (kk & 3) * 4);
#pragma unroll
            for (int i = 0; i < 4; i++) so[i] = o + (uint32_t)((rlo + ((i + rot) & 3)) * 16);
            base = s + kk * ld_ + row0;
            jstride = 32;
            kstride = ld_;
        }
        valid = special = 0;
        for (int j = 0; j < n_items; j++) {
            const long long row = row0 + 32 * j;
            const bool full = TRANS ? row + 3 < rv : row < rv;
            const bool part = TRANS ? (row < rv || (ones >= row && ones < row + 4)) : row == ones;
            if (full && vec) valid |= 1u << j;
            else if (full || part) special |= 1u << j;
        }
    }
};

// the rare cases: partial vectors, unaligned operands, the appended row of ones -- kept out of line (code size)
template <bool TRANS>
__device__ __noinline__ float4 load_item_careful(const float *src, long long ld, long long row, long long rows_valid,
                                                 long long ones_row, long long k, long long k_end) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (!TRANS) {
        if (row < rows_valid) {
            const float *p = src + row * ld;
            if (k < k_end) v.x = __ldg(p + k);
            if (k + 1 < k_end) v.y = __ldg(p + k + 1);
            if (k + 2 < k_end) v.z = __ldg(p + k + 2);
            if (k + 3 < k_end) v.w = __ldg(p + k + 3);
        } else if (row == ones_row) {
            v.x = k < k_end ? 1.f : 0.f; v.y = k + 1 < k_end ? 1.f : 0.f;
            v.z = k + 2 < k_end ? 1.f : 0.f; v.w = k + 3 < k_end ? 1.f : 0.f;
        }
    } else if (k < k_end) {
        const float *p = src + k * ld;
        if (row < rows_valid) v.x = __ldg(p + row);
        if (row + 1 < rows_valid) v.y = __ldg(p + row + 1);
        if (row + 2 < rows_valid) v.z = __ldg(p + row + 2);
        if (row + 3 < rows_valid) v.w = __ldg(p + row + 3);
        if (ones_row >= row && ones_row < row + 4) {
            const int d = (int)(ones_row - row);
            if (d == 0) v.x = 1.f; else if (d == 1) v.y = 1.f; else if (d == 2) v.z = 1.f; else v.w = 1.f;
        }
    }
    return v;
}

template <bool TRANS, int MAXI>
__device__ __forceinline__ void load_tile(float4 (&reg)[MAXI], const Stager<TRANS> &st, long long k0, long long k_end) {
    // in a partial last chunk an item is either wholly inside the reduction range (plain vector load), wholly outside
    // (zeros), or -- row-major sources with K % 4 != 0 only -- straddles its end (careful path)
    const long long k = k0 + st.kk;
    const bool k_in = TRANS ? k < k_end : k + 3 < k_end;
    const bool k_part = !TRANS && k < k_end && !k_in;
    const float *p = st.base + k0 * st.kstride;
#pragma unroll
    for (int j = 0; j < MAXI; j++) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        const uint32_t bit = 1u << j;
        if ((st.valid & bit) && k_in) v = __ldg(reinterpret_cast<const float4 *>(p));
        else if (((st.special & bit) && (k_in || k_part)) || ((st.valid & bit) && k_part))
            v = load_item_careful<TRANS>(st.src, st.ld, st.row0 + 32 * j, st.rows_valid, st.ones_row, k, k_end);
        reg[j] = v;
        p += st.jstride;
    }
}

__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }

template <bool TRANS, int MAXI, bool PREC3X>
__device__ __forceinline__ void store_tile(const float4 (&reg)[MAXI], const Stager<TRANS> &st, uint8_t *tile_hi, uint8_t *tile_lo) {
#pragma unroll
    for (int j = 0; j < MAXI; j++) {
        if (j < st.n_items) {
            float4 v = reg[j];
            if (!TRANS) {
                uint8_t *d = tile_hi + st.so[0] + j * 4096;
                if (PREC3X) {
                    const float4 h = make_float4(tf32_hi(v.x), tf32_hi(v.y), tf32_hi(v.z), tf32_hi(v.w));
                    *reinterpret_cast<float4 *>(d) = h;
                    *reinterpret_cast<float4 *>(tile_lo + st.so[0] + j * 4096) = make_float4(v.x - h.x, v.y - h.y, v.z - h.z, v.w - h.w);
                } else {
                    *reinterpret_cast<float4 *>(d) = v;
                }
            } else {
                // rotate the components by rot (two conditional stages), then four scalar stores at fixed offsets
                if (st.rot & 1) { const float t0 = v.x; v.x = v.y; v.y = v.z; v.z = v.w; v.w = t0; }
                if (st.rot & 2) { const float t0 = v.x, t1 = v.y; v.x = v.z; v.y = v.w; v.z = t0; v.w = t1; }
                const float c[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    if (PREC3X) {
                        const float h = tf32_hi(c[i]);
                        *reinterpret_cast<float *>(tile_hi + st.so[i] + j * 4096) = h;
                        *reinterpret_cast<float *>(tile_lo + st.so[i] + j * 4096) = c[i] - h;
                    } else {
                        *reinterpret_cast<float *>(tile_hi + st.so[i] + j * 4096) = c[i];
                    }
                }
            }
        }
    }
}

// PREC3X: 3xTF32;  AM / BM_: staging mode of the A / B tile (OP_*)
template <bool PREC3X, int AM, int BM_>
__global__ void __launch_bounds__(NTHR, PREC3X ? 1 : 2) gemm_kernel(const __grid_constant__ GemmArgs g) {
    static_assert(!PREC3X || (AM != OP_TMA && BM_ != OP_TMA), "the hi/lo split needs the operands in registers");
    constexpr bool ANY_TMA = AM == OP_TMA || BM_ == OP_TMA;
    constexpr bool ANY_THREADS = AM != OP_TMA || BM_ != OP_TMA;
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bars[MAX_STAGES + 1];        // stage free (MMAs have read it); [MAX_STAGES]: all done
    __shared__ __align__(8) uint64_t full[MAX_STAGES];            // TMA bytes of the stage have landed
    __shared__ uint32_t tmem_slot;
    __shared__ int dead;
    __shared__ float dot_part[2][BM];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int split = blockIdx.x % g.split_k;
    const int t = blockIdx.x / g.split_k;
    const int tn = t % g.tiles_n;
    const long long m0 = (long long)(t / g.tiles_n) * BM;
    const int n0 = tn * g.BN;
    int bn_eff = ((g.N - n0 + 15) >> 4) << 4;
    if (bn_eff > g.BN) bn_eff = g.BN;
    const long long k_begin = (long long)split * g.k_per_split;
    long long k_end = k_begin + g.k_per_split;
    if (k_end > g.K) k_end = g.K;
    const int nchunks = (int)((k_end - k_begin + KC - 1) / KC);
    if (nchunks <= 0) return;                                   // uniform over the CTA, nothing allocated yet

    const int b_tile_bytes = g.BN * KC * 4;
    const int stage_bytes = (A_TILE_BYTES + b_tile_bytes) * (PREC3X ? 2 : 1);
    const uint32_t tx_bytes = (AM == OP_TMA ? A_TILE_BYTES : 0) + (BM_ == OP_TMA ? b_tile_bytes : 0);

    if (tid == 0) {
        for (int s = 0; s <= MAX_STAGES; s++) mbar_init(smem_u32(&bars[s]), 1);
        for (int s = 0; s < MAX_STAGES; s++) mbar_init(smem_u32(&full[s]), 1);
        dead = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(g.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    const uint32_t idesc = instr_desc(0, 0, bn_eff);

    Stager<AM == OP_TRANSPOSED> sa;
    Stager<BM_ == OP_TRANSPOSED> sb;
    if (AM != OP_TMA) sa.init(g.A, g.lda, g.a_vec, BM, m0, g.M, g.colsum ? g.M : -1, tid);   // a row of ones appended to A: bias gradient
    if (BM_ != OP_TMA) sb.init(g.B, g.ldb, g.b_vec, g.BN, n0, g.N, -1, tid);

    // TMA operands: thread 0 keeps stages - 1 chunks in flight (issue for chunk c = the box at k = k_begin + 32 c)
    auto tma_issue = [&](int c) {
        const int s = c % g.stages;
        const uint32_t fb = smem_u32(&full[s]);
        const uint32_t dst = smem_u32(smem + (size_t)s * stage_bytes);
        const int kc = (int)(k_begin + (long long)c * KC);
        mbar_expect_tx(fb, tx_bytes);
        if (AM == OP_TMA) tma_load_2d(dst, &g.tmap_a, kc, (int)m0, fb);
        if (BM_ == OP_TMA) tma_load_2d(dst + A_TILE_BYTES, &g.tmap_b, kc, n0, fb);
    };
    const bool probe_mma = ANY_TMA && (g.dbg & 2);            // tuning probe: MMAs on whatever is in shared memory, no loads
    if (ANY_TMA && tid == 0 && !probe_mma) {
        for (int c = 0; c < g.stages - 1 && c < nchunks; c++) tma_issue(c);
    }

    float4 ra[4], rb[8];
    for (int i = ANY_THREADS ? -1 : 0; i < nchunks; i++) {
        if (!ANY_THREADS && tid != 0) break;                      // both operands by TMA: one thread drives the pipeline
        uint8_t *st = nullptr;
        if (i >= 0) {
            const int s = i % g.stages, u = i / g.stages;
            st = smem + (size_t)s * stage_bytes;
            if (ANY_THREADS) {
                if (u > 0) mbar_wait(smem_u32(&bars[s]), (uint32_t)((u - 1) & 1), &dead, g.err);
                if (AM != OP_TMA) store_tile<AM == OP_TRANSPOSED, 4, PREC3X>(ra, sa, st, st + A_TILE_BYTES + b_tile_bytes);
                if (BM_ != OP_TMA) store_tile<BM_ == OP_TRANSPOSED, 8, PREC3X>(rb, sb, st + A_TILE_BYTES, st + 2 * A_TILE_BYTES + b_tile_bytes);
            }
        }
        if (ANY_THREADS && i + 1 < nchunks) {                    // the next chunk's loads fly during the barrier and the MMAs
            const long long k0 = k_begin + (long long)(i + 1) * KC;
            if (AM != OP_TMA) load_tile<AM == OP_TRANSPOSED, 4>(ra, sa, k0, k_end);
            if (BM_ != OP_TMA) load_tile<BM_ == OP_TRANSPOSED, 8>(rb, sb, k0, k_end);
        }
        if (i >= 0) {
            if (ANY_THREADS) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncthreads();
            }
            if (tid == 0) {
                if (ANY_TMA && !probe_mma) mbar_wait(smem_u32(&full[i % g.stages]), (uint32_t)((i / g.stages) & 1), &dead, g.err);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                // thread-staged tiles are K-major without swizzle: 256 bytes per k-step of 8, LBO 128, SBO 1024;
                // TMA tiles carry the 128-byte swizzle: 32 bytes per k-step inside the row
                const uint32_t a_hi = smem_u32(st), b_hi = a_hi + A_TILE_BYTES;
                const uint32_t a_lo = b_hi + b_tile_bytes, b_lo = a_lo + A_TILE_BYTES;
#pragma unroll
                for (int j = 0; j < KC / 8; j++) {
                    const uint64_t da = AM == OP_TMA ? smem_desc_sw128(a_hi + j * 32) : smem_desc(a_hi + j * 256, 128u, 1024u);
                    const uint64_t db = BM_ == OP_TMA ? smem_desc_sw128(b_hi + j * 32) : smem_desc(b_hi + j * 256, 128u, 1024u);
                    if (PREC3X) {
                        const uint64_t dal = smem_desc(a_lo + j * 256, 128u, 1024u), dbl = smem_desc(b_lo + j * 256, 128u, 1024u);
                        mma_tf32(tmem, dal, db, idesc, (i > 0 || j > 0) ? 1u : 0u);
                        mma_tf32(tmem, da, dbl, idesc, 1u);
                        mma_tf32(tmem, da, db, idesc, 1u);
                    } else {
                        mma_tf32(tmem, da, db, idesc, (i > 0 || j > 0) ? 1u : 0u);
                    }
                }
                mma_commit(smem_u32(&bars[i % g.stages]));       // the stage is free once these MMAs have read it
                if (i == nchunks - 1) mma_commit(smem_u32(&bars[MAX_STAGES]));
                if (ANY_TMA && !probe_mma) {
                    const int c = i + g.stages - 1;              // refill the stage chunk i - 1 used
                    if (c < nchunks) {
                        if (c >= g.stages) mbar_wait(smem_u32(&bars[c % g.stages]), (uint32_t)(((c / g.stages) - 1) & 1), &dead, g.err);
                        tma_issue(c);
                    }
                }
            }
        }
    }
    if (tid == 0) mbar_wait(smem_u32(&bars[MAX_STAGES]), 0u, &dead, g.err);   // every MMA has written the accumulator
    __syncthreads();                                               // (the other threads sleep in the hardware barrier)
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    // ---- epilogue: TMEM -> registers (thread = row) -> per-warp transposition buffer -> coalesced global rows ----
    // bias / relu6 / mask are applied in the second phase (lane = column: one bias register per lane, the mask read
    // coalesced); only the value head needs the finished activations in the first phase (thread = row).
    const int lane_base = (warp & 3) * 32, half = warp >> 2;
    float *buf = reinterpret_cast<float *>(smem) + warp * (32 * 33);
    const long long row_base = m0 + lane_base;
    const float *bias = g.bias, *dot_w = g.dot_w, *mask_src = g.mask_src;
    float *D = g.D;
    const bool relu6 = g.relu6 != 0, accumulate = g.accumulate != 0;
    float *out_colsum = g.out_colsum;
    const long long ldd = g.ldd, ld_mask = g.ld_mask;
    long long left = g.M - row_base;                              // rows of this warp that exist in D
    const int nrows = left >= 32 ? 32 : (left > 0 ? (int)left : 0);
    const bool ones_here = g.colsum && left >= 0 && left < 32;    // the appended row of ones lands in this warp's rows
    float dot = 0.f;
    for (int c0 = half * 32; c0 < bn_eff; c0 += 64) {
        const int col = n0 + c0 + lane;
        const bool col_ok = col < g.N && c0 + lane < bn_eff;
        // the relu6' mask of the chunk's 32 rows is requested before the accumulators are touched: 32 loads in flight per
        // lane while tensor memory is read and transposed (8 in flight made the epilogue a latency chain)
        float m[32];
        const bool mask_all = mask_src && col_ok && nrows == 32;
        if (mask_all) {
            const float *mp = mask_src + row_base * ld_mask + col;
#pragma unroll
            for (int q = 0; q < 32; q++) m[q] = __ldg(mp + (long long)q * ld_mask);
        }
        float v[32];
        tmem_ld32(tmem + ((uint32_t)lane_base << 16) + (uint32_t)c0, v);
        if (dot_w) {                                               // critic: relu6(. + bias) . w3 per row
#pragma unroll
            for (int j = 0; j < 32; j++) {
                const int cj = n0 + c0 + j;
                if (cj < g.N) {
                    float x = v[j] + (bias ? __ldg(bias + cj) : 0.f);
                    if (relu6) x = fminf(fmaxf(x, 0.f), 6.f);
                    dot += x * __ldg(dot_w + cj);
                }
            }
        }
#pragma unroll
        for (int j = 0; j < 32; j++) buf[lane * 33 + j] = v[j];
        __syncwarp();
        if (col_ok) {
            if (D && nrows > 0) {
                const float bcol = bias ? __ldg(bias + col) : 0.f;
                float *dp = D + row_base * ldd + col;
                const float *bp = buf + lane;
                float csum = 0.f;
                if (nrows == 32) {
#pragma unroll
                    for (int q = 0; q < 32; q++) {
                        float y = bp[q * 33] + bcol;
                        if (relu6) y = fminf(fmaxf(y, 0.f), 6.f);
                        if (mask_all) y = (m[q] > 0.f && m[q] < 6.f) ? y : 0.f;
                        csum += y;
                        if (accumulate) atomicAdd(dp, y); else *dp = y;
                        dp += ldd;
                    }
                } else {
                    const float *mp = mask_src ? mask_src + row_base * ld_mask + col : nullptr;
                    for (int rr = 0; rr < nrows; rr++) {
                        float y = bp[rr * 33] + bcol;
                        if (relu6) y = fminf(fmaxf(y, 0.f), 6.f);
                        if (mp) { const float mm = __ldg(mp); mp += ld_mask; y = (mm > 0.f && mm < 6.f) ? y : 0.f; }
                        csum += y;
                        if (accumulate) atomicAdd(dp, y); else *dp = y;
                        dp += ldd;
                    }
                }
                if (out_colsum) atomicAdd(out_colsum + col, csum);
            }
            if (ones_here) atomicAdd(g.colsum + col, buf[(int)left * 33 + lane]);
        }
        __syncwarp();
    }
    if (g.dot_out) {
        dot_part[half][lane_base + lane] = dot;
        __syncthreads();
        if (tid < BM && m0 + tid < g.M) g.dot_out[m0 + tid] = dot_part[0][tid] + dot_part[1][tid] + (g.dot_b ? __ldg(g.dot_b) : 0.f);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(g.tmem_cols) : "memory");
    }
}

}  // namespace tc
}  // namespace uavk
