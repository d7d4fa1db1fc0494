// Counter-based RNG for the batched MobiEnvironment step: Philox4x32-10 (Salmon et al., SC'11).
//
// The reference draws from numpy's global MT19937 stream (ue_mobility.py:6,408; channel.py:240), which is
// unseeded, sequential and shared between threads.  Here every draw is a pure function of
//   key     = (seed_lo, seed_hi)
//   counter = (global env id, index inside the env, sequence number, domain)
// so a result never depends on how environments are sharded over CTAs / handles / GPUs.
// oracle/mobi_oracle.c implements the identical scheme on the CPU (test infrastructure).
#pragma once
#include <stdint.h>

namespace uavk {

enum PhiloxDomain : uint32_t {
    DOM_INIT_XY = 1,   // idx = UE:    a -> x0,  b -> y0                       (ue_mobility.py:434-435)
    DOM_INIT_TH = 2,   // idx = UE:    a -> theta0                             (ue_mobility.py:437)
    DOM_INIT_GXY = 3,  // idx = group: a -> g_x, b -> g_y                      (ue_mobility.py:442-443)
    DOM_INIT_GFV = 4,  // idx = group: a -> g_fl, b -> g_v                     (ue_mobility.py:444-445)
    DOM_INIT_GTH = 5,  // idx = group: a -> g_theta                            (ue_mobility.py:446)
    DOM_THETA = 6,     // idx = UE, seq = tick: a -> theta redraw              (ue_mobility.py:508)
    DOM_GRP_TF = 7,    // idx = group, seq = tick: a -> theta, b -> g_fl       (ue_mobility.py:516,519)
    DOM_GRP_V = 8,     // idx = group, seq = tick: a -> g_v                    (ue_mobility.py:520)
    DOM_FADING = 9,    // idx = UE*ceil(nBS/4) + b/4, seq = channel pass: 4 normals (channel.py:240)
    DOM_ACTION = 10,   // idx = BS, seq = step: a -> digit (synthetic actions for benchmarks)
    DOM_AREA = 11,     // idx = cell*ceil(nBS/4) + b/4, seq = coverage-map call: 4 normals (channel.py:426,429)
    DOM_SAMPLE = 12    // env = sample row, idx = 0, seq = call counter: a -> action draw (np.random.choice, main.py:167)
};

struct Philox4 {
    uint32_t w[4];
};

__device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                                 uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; r++) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    Philox4 o;
    o.w[0] = c0; o.w[1] = c1; o.w[2] = c2; o.w[3] = c3;
    return o;
}

// 53-bit uniform in [0,1) from two words (same construction as numpy's random_sample)
__device__ __forceinline__ double u53(uint32_t hi, uint32_t lo) {
    return ((double)(hi >> 5) * 67108864.0 + (double)(lo >> 6)) * (1.0 / 9007199254740992.0);
}

__device__ __forceinline__ void philox_uniform2(uint32_t k0, uint32_t k1, uint32_t env, uint32_t idx, uint32_t seq,
                                                uint32_t dom, double &a, double &b) {
    const Philox4 p = philox4x32_10(env, idx, seq, dom, k0, k1);
    a = u53(p.w[0], p.w[1]);
    b = u53(p.w[2], p.w[3]);
}

// Four N(0,1): (w0,w1) and (w2,w3) each feed one Box-Muller pair, u = (w + 0.5) * 2^-32 in (0,1).
__device__ __forceinline__ void normal4_f64(const Philox4 &p, double z[4]) {
#pragma unroll
    for (int k = 0; k < 2; k++) {
        const double u1 = ((double)p.w[2 * k] + 0.5) * (1.0 / 4294967296.0);
        const double u2 = ((double)p.w[2 * k + 1] + 0.5) * (1.0 / 4294967296.0);
        const double r = sqrt(-2.0 * log(u1));
        double s, c;
        sincos(6.283185307179586 * u2, &s, &c);
        z[2 * k] = r * c;
        z[2 * k + 1] = r * s;
    }
}

// The bare special-function unit: MUFU.LG2 / EX2 / SQRT without the denormal-range guards that __log2f / exp2f / sqrtf
// wrap around them (3-8 instructions and a branch per call: a sixth of the dense channel pass, profiles/r2/NOTES.md).
// Every argument on the fp32 paths is a normal float (squared distances >= 1, powers >= the noise floor 8e-16 W,
// uniforms >= 2^-33), so the guards never fire.
__device__ __forceinline__ float mufu_lg2(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float mufu_ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float mufu_sqrt(float x) { float y; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// fp32 / MUFU version of the same draw (lg2, sqrt, sin, cos).  Differs from normal4_f64 by ~1e-6 absolute.
__device__ __forceinline__ void normal4_f32(const Philox4 &p, float z[4]) {
#pragma unroll
    for (int k = 0; k < 2; k++) {
        // (w + 0.5) * 2^-32; the fp32 rounding of w can reach 2^32, giving u1 = 1 -> r = 0 (harmless)
        const float u1 = fmaf((float)p.w[2 * k], 2.3283064365386963e-10f, 1.1641532182693481e-10f);
        const float u2 = fmaf((float)p.w[2 * k + 1], 2.3283064365386963e-10f, 1.1641532182693481e-10f);
        const float r = mufu_sqrt(-1.3862943611198906f * mufu_lg2(u1));   // -2 ln u1 = -2 ln2 * log2 u1
        float s, c;
        __sincosf(6.2831853071795865f * (u2 - 0.5f), &s, &c);     // angle in (-pi, pi); the half-turn shift
        z[2 * k] = -r * c;                                         // is undone by the sign flips
        z[2 * k + 1] = -r * s;
    }
}

}  // namespace uavk
