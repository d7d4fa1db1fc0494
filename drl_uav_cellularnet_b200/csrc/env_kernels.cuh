// Fused MobiEnvironment step / reset / constructor-pass kernels for sm_100a.
//
// One CTA owns one environment for the whole call: UE mobility tick -> BS move -> UE x BS channel pass ->
// best server / time-to-trigger handover / new-outage count -> reward -> observation.  Nothing is exchanged
// between environments (reference: one env object per worker, main.py:173), so the grid is E CTAs.  The dense
// observation -- 99 % of the bytes a step moves -- is streamed by the TMA engine (cp.async.bulk copies of one zeroed
// shared-memory tile) while the warps compute; DESIGN.md section 3 explains the shape and profiles/r1/NOTES.md the
// measurements behind it.
//
// Reference semantics restated here (file:line into the reference repo):
//   mobile_env.py:150-194 / 196-233  step / step_test          -> env_kernel<.., MODE_STEP>
//   mobile_env.py:115-148            reset                     -> MODE_RESET
//   mobile_env.py:100 + channel.py:92-93,110  LTEChannel ctor  -> MODE_CTOR
//   ue_mobility.py:409-523           reference_point_group     -> mob_group_load / mob_ue_move / mob_group_finish
//   ue_mobility.py:191-271,310-336   BS_move, Decimal_to_Base_N-> the BS warp of env_kernel / bs_move_warp
//   channel.py:220-269               gain / SINR               -> ue_channel_pass (thread = UE), ue_channel_quad (lane = 4 BSs)
//   channel.py:138-176,216           UpdateDroneNet            -> ho_decide (handover word)
//   channel.py:387-409, ue_mobility.py:173-188  association / BS grid map -> issue_zero_stream + obs_add
//   channel.py:411-433               GetSinrInArea             -> coverage_kernel
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "philox.cuh"

namespace uavk {

constexpr int MAX_BS = 32;
constexpr int MAX_GROUPS = 32;
constexpr int CTA_THREADS = 256;
constexpr int GUARD_LIST = 64;

enum { MODE_STEP = 0, MODE_RESET = 1, MODE_CTOR = 2 };
enum { MOB_GROUP = 0, MOB_TRACE = 1 };
enum { FADE_PHILOX = 0, FADE_INJECTED = 1, FADE_NONE = 2 };
enum { OBS_NONE = 0, OBS_F32 = 1, OBS_F32_INCREMENTAL = 3 };
enum { ERR_ACTION = 1u, ERR_TRACE = 2u, ERR_CLAMP = 4u, ERR_BOUNDS = 8u };

// Observation count update.  -DUAVENV_BOUNDS_CHECK builds verify every index against the env's observation and raise
// the sticky ERR_BOUNDS flag instead of writing (compute-sanitizer is closed on the measurement pool; this is the
// bounds check of our own that profiles/all_paths_case.py runs).
__device__ __forceinline__ void obs_add(float *obs_env, long long lin, float v, long long n_cells, uint32_t *err_flags) {
#ifdef UAVENV_BOUNDS_CHECK
    if (lin < 0 || lin >= n_cells) { atomicOr(err_flags, ERR_BOUNDS); return; }
#endif
    atomicAdd(obs_env + lin, v);
}

enum { CTR_TICK = 0, CTR_EPOCH = 1, CTR_STEP = 2, CTR_AGG = 3, CTR_DEAGG = 4, CTR_STRIDE = 8 };

// handover word, one per UE (channel.py:75-81,92-93): current_BS, the <=3 rows of bestBS_buf, its depth, and
// whether the UE was in outage after the previous pass (membership in self.ue_out, channel.py:116,171-174)
__device__ __forceinline__ uint32_t ho_pack(int cur, int f0, int f1, int f2, int depth, int outp) {
    return (uint32_t)cur | ((uint32_t)f0 << 5) | ((uint32_t)f1 << 10) | ((uint32_t)f2 << 15) |
           ((uint32_t)depth << 20) | ((uint32_t)outp << 22);
}

// FP32_GUARDED: the UE's decision is pending (set only between the UE loop and the guard phase of one launch)
constexpr uint32_t HO_PENDING = 1u << 28;

// Everything that is fixed for the lifetime of a handle.  Passed by value as a __grid_constant__ parameter.
struct DevCfg {
    int E, nBS, nUE, G, nG;
    int mobility, fading, obs_mode;
    int max_step, n_act, bs_step, lock_r2;
    int deagg_len, agg_len;
    uint32_t k0, k1;        // Philox key = seed
    uint32_t env_offset;    // global id of env 0 of this handle
    // float64 constants (parity path; reference operation order)
    double grid_width, P, N, pl_a, pl_b, pl_dis, ant_gain, eq_loss, sh_mean, sh_sd, ho_thr, out_thr;
    double v_min, v_max, aggr, max_xy, fl_max;
    // float32 constants (fast path; log-domain form)
    float f_q_scale;        // grid_width^2
    float f_q_min;          // pl_dis^2: loss applies when q > f_q_min
    float f_g0;             // ant_gain - eq_loss
    float f_loss_a;         // pl_a
    float f_loss_k;         // (pl_b / 2) * log10(2):   pl_b*log10(sqrt(q)) = f_loss_k * log2(q)
    float f_exp_k;          // log2(10) / 10:           10^(g/10) = 2^(g * f_exp_k)
    float f_log2P;          // log2(P)
    float f_Pdb;            // 10 log10(P)
    float f_db_k;           // 10 log10(2):             10 log10(x) = f_db_k * log2(x)
    float f_N, f_sh_mean, f_sh_sd;
    // folded constants of the 4-BSs-per-lane mapping (ue_channel_quad): squared distance in cells
    float f_d2_min;         // (pl_dis / grid_width)^2: path loss applies when d2 > f_d2_min
    float f_c1;             // ant_gain - eq_loss - pl_a - f_loss_k log2(grid_width^2) + 10 log10(P)
    float f_c0;             // ant_gain - eq_loss + 10 log10(P)                (no path loss)
    // FP32_GUARDED: a UE whose fp32 SINR row is within guard_db (dB) of a decision boundary -- top-2 gap (argmax),
    // best - current - ho_thr (handover), serving SINR - out_thr (outage) -- is re-evaluated in float64 with the
    // reference operation order, so every decision equals the FP64_PARITY kernel's.  0 = off (FP32_FAST).
    float guard_db;
    unsigned long long *guard_hits;   // [1] UEs re-evaluated so far (diagnostic)
    // persistent state (device)
    double2 *xy;            // [E,nUE] float UE positions (x, y): one 16-byte load (ue_mobility.py:434-435)
    double *th_u;           // [E,nUE] last theta uniform, injected-mobility runs only
    double *grp;            // [E,6,nG] g_x g_y g_fl g_v g_cos g_sin           (ue_mobility.py:442-448)
    int32_t *ctr;           // [E,8]   tick, epoch, step_n, aggregating, deaggregating
    int16_t *bs_xy;         // [E,nBS,2]
    int16_t *ue_cell;       // [E,nUE,2] integer UE cells                      (mobile_env.py:155)
    uint32_t *ho;           // [E,nUE] handover words
    const int16_t *init_bs; // [nBS,2]
    const uint8_t *ue_group;// [nUE]   g_ref                                    (ue_mobility.py:423-426)
    const int32_t *trace;   // [T,(E|1),nUE,2] or null
    int64_t trace_T;
    int trace_per_env;
    uint32_t *err_flags;    // [1] sticky
};

struct CallArgs {
    int mode;
    int inject_mob;              // mob_u supplied
    const int64_t *action;       // [E]
    const uint8_t *digits;       // [E,nBS]
    const double *fading;        // [E,nUE,nBS]
    const double *mob_u;         // [E,nUE+3nG]
    const uint8_t *env_mask;     // [E]
    float *obs;                  // [E,nBS+1,G,G]
    double *reward, *mean_sinr;  // [E]
    int32_t *n_out, *n_ho, *n_blocked, *step_n;
    uint8_t *done;
    uint8_t *serving;            // [E,nUE]
    void *serving_sinr;          // [E,nUE] f32 / f64
    void *sinr_all;              // [E,nUE,nBS] f32 / f64
    float *fading_used;          // [E,nUE,nBS]
    int16_t *ue_xy, *bs_xy_out;  // [E,nUE,2], [E,nBS,2]
    uint8_t *bs_digits;          // [E,nBS]
    int32_t *obs_idx;            // [E,nUE+nBS] flat indices of the observation's non-zero cells
    int tile_bytes;              // bytes of the zeroed shared-memory tile the TMA warp streams from (0: no TMA path)
    int cells_off;               // byte offset of the per-UE observation indices in dynamic shared memory (-1: re-read HBM)
    int copies_per_turn;         // fp32, more than 4 BSs: bulk copies of the zero stream the TMA warp issues per chunk it draws
};

struct EnvShared {
    double gx[MAX_GROUPS], gy[MAX_GROUPS], gfl[MAX_GROUPS], gv[MAX_GROUPS], gcos[MAX_GROUPS], gsin[MAX_GROUPS];
    int refl[4][MAX_GROUPS];
    int bsx[MAX_BS], bsy[MAX_BS];
    int digit[MAX_BS];
    double red_sinr[CTA_THREADS / 32];
    int red_out[CTA_THREADS / 32], red_ho[CTA_THREADS / 32];
    int ok, blocked;
    // FP32_GUARDED, more than 4 BSs: UEs waiting for their float64 re-evaluation by a whole warp (lane = BS), and the
    // serving SINR of those UEs as an order-independent fixed-point sum (2^-32 dB units)
    int next_chunk;               // fp32, more than 4 BSs: the next chunk of 32 UEs a warp will take
    int guard_n;
    int guard_ue[GUARD_LIST];
    long long guard_sum;
};

// U(MIN,MAX,.) = rand*(MAX-MIN)+MIN (ue_mobility.py:408), no fused multiply-add
__device__ __forceinline__ double U_(double lo, double hi, double r) { return __dadd_rn(__dmul_rn(r, hi - lo), lo); }

constexpr double TWO_PI = 6.283185307179586;

// ---------------------------------------------------------------------------------------------------------
// L2 eviction-priority hints.  Every step streams hundreds of MB of write-once observation through L2 (evict-first,
// below) while the environments' own state (a few MB: positions, handover words, cells) is read and written every
// step: it is tagged evict-last so that it stays L2-resident under the stream instead of being re-fetched from HBM
// behind the stream's write queue.
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
#ifndef UAVENV_NO_KEEP_HINT
__device__ __forceinline__ double ldk(const double *a, uint64_t pol) {
    double v;
    asm volatile("ld.global.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(a), "l"(pol));
    return v;
}
__device__ __forceinline__ uint32_t ldk(const uint32_t *a, uint64_t pol) {
    uint32_t v;
    asm volatile("ld.global.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(v) : "l"(a), "l"(pol));
    return v;
}
__device__ __forceinline__ void stk(double *a, double v, uint64_t pol) {
    asm volatile("st.global.L2::cache_hint.f64 [%0], %1, %2;" ::"l"(a), "d"(v), "l"(pol) : "memory");
}
__device__ __forceinline__ void stk(uint32_t *a, uint32_t v, uint64_t pol) {
    asm volatile("st.global.L2::cache_hint.u32 [%0], %1, %2;" ::"l"(a), "r"(v), "l"(pol) : "memory");
}
__device__ __forceinline__ double2 ldk(const double2 *a, uint64_t pol) {
    double2 v;
    asm volatile("ld.global.L2::cache_hint.v2.f64 {%0, %1}, [%2], %3;" : "=d"(v.x), "=d"(v.y) : "l"(a), "l"(pol));
    return v;
}
__device__ __forceinline__ void stk(double2 *a, double2 v, uint64_t pol) {
    asm volatile("st.global.L2::cache_hint.v2.f64 [%0], {%1, %2}, %3;" ::"l"(a), "d"(v.x), "d"(v.y), "l"(pol) : "memory");
}
__device__ __forceinline__ void stk(float *a, float v, uint64_t pol) {
    asm volatile("st.global.L2::cache_hint.f32 [%0], %1, %2;" ::"l"(a), "f"(v), "l"(pol) : "memory");
}
__device__ __forceinline__ void stk(int32_t *a, int32_t v, uint64_t pol) {
    asm volatile("st.global.L2::cache_hint.s32 [%0], %1, %2;" ::"l"(a), "r"(v), "l"(pol) : "memory");
}
__device__ __forceinline__ void stk(uint8_t *a, uint8_t v, uint64_t pol) {
    asm volatile("st.global.L2::cache_hint.u8 [%0], %1, %2;" ::"l"(a), "r"((uint32_t)v), "l"(pol) : "memory");
}
#else
__device__ __forceinline__ void stk(float *a, float v, uint64_t) { *a = v; }
__device__ __forceinline__ void stk(int32_t *a, int32_t v, uint64_t) { *a = v; }
__device__ __forceinline__ void stk(uint8_t *a, uint8_t v, uint64_t) { *a = v; }
__device__ __forceinline__ double2 ldk(const double2 *a, uint64_t) { return *a; }
__device__ __forceinline__ void stk(double2 *a, double2 v, uint64_t) { *a = v; }
__device__ __forceinline__ double ldk(const double *a, uint64_t) { return *a; }
__device__ __forceinline__ uint32_t ldk(const uint32_t *a, uint64_t) { return *a; }
__device__ __forceinline__ void stk(double *a, double v, uint64_t) { *a = v; }
__device__ __forceinline__ void stk(uint32_t *a, uint32_t v, uint64_t) { *a = v; }
#endif
__device__ __forceinline__ short2 ldk_cell(const int16_t *cells, size_t i, uint64_t pol) {
    const uint32_t w = ldk(reinterpret_cast<const uint32_t *>(cells) + i, pol);
    return make_short2((short)(w & 0xffffu), (short)(w >> 16));
}
__device__ __forceinline__ void stk_cell(int16_t *cells, size_t i, short2 c, uint64_t pol) {
    stk(reinterpret_cast<uint32_t *>(cells) + i, (uint32_t)(uint16_t)c.x | ((uint32_t)(uint16_t)c.y << 16), pol);
}

// cp.async (LDGSTS) with an L2 eviction-priority hint: one lane's piece of the NEXT chunk's state goes from HBM / L2 to
// shared memory while the warp works on the current chunk (the fp32 mapping for more than 4 BSs, env_kernel)
__device__ __forceinline__ void cp_async16(void *sdst, const void *gsrc, uint64_t pol) {
    asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(sdst)),
                 "l"(__cvta_generic_to_global(gsrc)), "l"(pol) : "memory");
}
__device__ __forceinline__ void cp_async4(void *sdst, const void *gsrc, uint64_t pol) {
    asm volatile("cp.async.ca.shared.global.L2::cache_hint [%0], [%1], 4, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(sdst)),
                 "l"(__cvta_generic_to_global(gsrc)), "l"(pol) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// ---------------------------------------------------------------------------------------------------------
// One tick of the reference_point_group generator (ue_mobility.py:453-523), split in three so that the step
// kernel can fuse the per-UE part with the channel pass of the same thread:
//   mob_group_load   (nG lanes of one warp)  group centres advance, state -> shared memory      :458-459
//   mob_ue_move      (one thread per UE)     random walk + group drift + aggregation + walls    :455-456,461-505
//   mob_group_finish (one warp)              wall flips, flight length, arrivals, state -> HBM  :493-521
// A CTA barrier separates each part from the next.
struct GroupRow { double gx, gy, gfl, gv, gcos, gsin; };     // one group's state as stored (c.grp[e, :, g])

__device__ __forceinline__ GroupRow group_row_load(const DevCfg &c, int e, int g) {
    const int nG = c.nG;
    const double *grp = c.grp + (size_t)e * 6 * nG;
    GroupRow r;
    r.gx = grp[0 * nG + g]; r.gy = grp[1 * nG + g]; r.gfl = grp[2 * nG + g];
    r.gv = grp[3 * nG + g]; r.gcos = grp[4 * nG + g]; r.gsin = grp[5 * nG + g];
    return r;
}

__device__ __forceinline__ void mob_group_load(EnvShared &s, const GroupRow &r, int g) {
    s.gx[g] = __dadd_rn(r.gx, __dmul_rn(r.gv, r.gcos));               // :458
    s.gy[g] = __dadd_rn(r.gy, __dmul_rn(r.gv, r.gsin));               // :459
    s.gfl[g] = r.gfl;
    s.gv[g] = r.gv; s.gcos[g] = r.gcos; s.gsin[g] = r.gsin;
    s.refl[0][g] = 0; s.refl[1][g] = 0; s.refl[2][g] = 0; s.refl[3][g] = 0;
}

// inj: the uniforms the reference generator would draw this tick, in its order (theta[nUE], then for the k
// arrived groups theta[k], fl[k], v[k]); null = Philox.  Returns the UE's integer cell.
// x, y (and thu, the injected direction uniform) are the UE's stored state, loaded by the caller.
__device__ __forceinline__ short2 mob_ue_move(const DevCfg &c, EnvShared &s, int e, uint32_t genv, int tick,
                                              bool aggregating, const double *inj, int u, int g, double x, double y,
                                              double thu, uint64_t keep) {
    const size_t i = (size_t)e * c.nUE + u;
    // direction drawn at the end of the previous tick (:508-510) or at init (:437-439)
    double tu;
    if (inj) tu = thu;
    else {
        double b_;
        if (tick == 0) philox_uniform2(c.k0, c.k1, genv, (uint32_t)u, 0u, DOM_INIT_TH, tu, b_);
        else philox_uniform2(c.k0, c.k1, genv, (uint32_t)u, (uint32_t)(tick - 1), DOM_THETA, tu, b_);
    }
    double sn, cs;
    sincos(U_(0.0, TWO_PI, tu), &sn, &cs);
    x = __dadd_rn(x, cs);                                               // :455 (velocity 1.0, :436)
    y = __dadd_rn(y, sn);                                               // :456
    const double gvx = __dmul_rn(s.gv[g], s.gcos[g]), gvy = __dmul_rn(s.gv[g], s.gsin[g]);
    if (aggregating) {                                                  // :461-470
        double sc, cc;
        sincos(atan2(s.gy[g] - y, s.gx[g] - x), &sc, &cc);
        x = __dadd_rn(__dadd_rn(x, gvx), __dmul_rn(c.aggr, cc));
        y = __dadd_rn(__dadd_rn(y, gvy), __dmul_rn(c.aggr, sc));
    } else {                                                            // :475-484
        x = __dadd_rn(x, gvx);
        y = __dadd_rn(y, gvy);
    }
    // reflecting walls, reference order (:490-505); a group flips once per wall if any member reflects
    if (x < 0.0) { x = -x; s.refl[0][g] = 1; }
    if (x > c.max_xy) { x = __dadd_rn(__dmul_rn(2.0, c.max_xy), -x); s.refl[1][g] = 1; }
    if (y < 0.0) { y = -y; s.refl[2][g] = 1; }
    if (y > c.max_xy) { y = __dadd_rn(__dmul_rn(2.0, c.max_xy), -y); s.refl[3][g] = 1; }
    stk(c.xy + i, make_double2(x, y), keep);
    if (inj) stk(c.th_u + i, inj[u], keep);
    // np.concatenate(...).astype(int): truncation toward zero (mobile_env.py:154-155).  A UE exactly on the far
    // wall would index cell G (IndexError in the reference, ue_mobility.py:186): clamp + flag.
    int cx = (int)x, cy = (int)y;
    if (cx >= c.G || cy >= c.G) {
        cx = min(cx, c.G - 1); cy = min(cy, c.G - 1);
        atomicOr(c.err_flags, ERR_CLAMP);
    }
    return make_short2((short)cx, (short)cy);
}

// the phase counters (:471-473, :485-487); CTA-uniform register copies
__device__ __forceinline__ void mob_phase_advance(const DevCfg &c, int &agg, int &deagg) {
    if (agg != 0) { agg -= 1; if (agg == 0) deagg = c.deagg_len; }
    else { deagg -= 1; if (deagg == 0) agg = c.agg_len; }
}

// all 32 lanes of one warp; lane = group
__device__ __forceinline__ void mob_group_finish(const DevCfg &c, EnvShared &s, int e, uint32_t genv, int tick,
                                                 const double *inj, int lane) {
    const int nG = c.nG, nUE = c.nUE, g = lane;
    double *grp = c.grp + (size_t)e * 6 * nG;
    const bool live = g < nG;
    double gc = 0, gs = 0, gfl = 0, gv = 0;
    if (live) {
        gc = s.gcos[g]; gs = s.gsin[g];
        if (s.refl[0][g]) gc = -gc;
        if (s.refl[1][g]) gc = -gc;
        if (s.refl[2][g]) gs = -gs;
        if (s.refl[3][g]) gs = -gs;
        gv = s.gv[g];
        gfl = __dadd_rn(s.gfl[g], -gv);                                // :513
    }
    const bool arrived = live && gv > 0.0 && gfl <= 0.0;               // :514
    const unsigned am = __ballot_sync(0xffffffffu, arrived);
    if (arrived) {                                                      // :515-521
        double ut, uf, uv;
        if (inj) {
            const int k = __popc(am), r = __popc(am & ((1u << g) - 1u));
            ut = inj[nUE + r]; uf = inj[nUE + k + r]; uv = inj[nUE + 2 * k + r];
        } else {
            double b_;
            philox_uniform2(c.k0, c.k1, genv, (uint32_t)g, (uint32_t)tick, DOM_GRP_TF, ut, uf);
            philox_uniform2(c.k0, c.k1, genv, (uint32_t)g, (uint32_t)tick, DOM_GRP_V, uv, b_);
        }
        sincos(U_(0.0, TWO_PI, ut), &gs, &gc);
        gfl = U_(0.0, c.fl_max, uf);
        gv = U_(c.v_min, c.v_max, uv);
    }
    if (live) {
        grp[0 * nG + g] = s.gx[g]; grp[1 * nG + g] = s.gy[g];
        grp[2 * nG + g] = gfl; grp[3 * nG + g] = gv;
        grp[4 * nG + g] = gc; grp[5 * nG + g] = gs;
    }
}

// ---------------------------------------------------------------------------------------------------------
// Mobility init (ue_mobility.py:434-448) + warm-up ticks (mobile_env.py:77-79) + the tick that yields the
// constructor's UE positions (mobile_env.py:93-97).  Philox only.
__global__ void __launch_bounds__(CTA_THREADS) mob_init_kernel(const __grid_constant__ DevCfg c, int warmup,
                                                               int agg0, int deagg0) {
    __shared__ EnvShared s;
    const int e = blockIdx.x, tid = threadIdx.x;
    const uint32_t genv = c.env_offset + (uint32_t)e;
    for (int u = tid; u < c.nUE; u += blockDim.x) {
        double a, b;
        philox_uniform2(c.k0, c.k1, genv, (uint32_t)u, 0u, DOM_INIT_XY, a, b);
        c.xy[(size_t)e * c.nUE + u] = make_double2(U_(0.0, c.max_xy, a), U_(0.0, c.max_xy, b));
    }
    if (tid < c.nG) {
        double *grp = c.grp + (size_t)e * 6 * c.nG;
        double a, b, sn, cs;
        philox_uniform2(c.k0, c.k1, genv, (uint32_t)tid, 0u, DOM_INIT_GXY, a, b);
        grp[0 * c.nG + tid] = U_(0.0, c.max_xy, a);
        grp[1 * c.nG + tid] = U_(0.0, c.max_xy, b);                    // MAX_X (sic), ue_mobility.py:443
        philox_uniform2(c.k0, c.k1, genv, (uint32_t)tid, 0u, DOM_INIT_GFV, a, b);
        grp[2 * c.nG + tid] = U_(0.0, c.fl_max, a);
        grp[3 * c.nG + tid] = U_(c.v_min, c.v_max, b);
        philox_uniform2(c.k0, c.k1, genv, (uint32_t)tid, 0u, DOM_INIT_GTH, a, b);
        sincos(U_(0.0, TWO_PI, a), &sn, &cs);
        grp[4 * c.nG + tid] = cs;
        grp[5 * c.nG + tid] = sn;
    }
    __syncthreads();
    int agg = agg0, deagg = deagg0;
    const uint64_t keep = l2_policy_evict_last();
    for (int t = 0; t <= warmup; t++) {
        if (tid < c.nG) mob_group_load(s, group_row_load(c, e, tid), tid);
        __syncthreads();
        for (int u = tid; u < c.nUE; u += blockDim.x) {
            const size_t i = (size_t)e * c.nUE + u;
            const double2 p = c.xy[i];
            const short2 cell = mob_ue_move(c, s, e, genv, t, agg != 0, nullptr, u, c.ue_group[u], p.x, p.y, 0.0, keep);
            if (t == warmup) reinterpret_cast<short2 *>(c.ue_cell)[(size_t)e * c.nUE + u] = cell;
        }
        mob_phase_advance(c, agg, deagg);
        __syncthreads();
        if (tid < 32) mob_group_finish(c, s, e, genv, t, nullptr, tid);
        __syncthreads();
    }
    if (tid == 0) {
        int32_t *ctr = c.ctr + (size_t)e * CTR_STRIDE;
        ctr[CTR_TICK] = warmup + 1; ctr[CTR_EPOCH] = 0; ctr[CTR_STEP] = 0;
        ctr[CTR_AGG] = agg; ctr[CTR_DEAGG] = deagg;
    }
}

// ---------------------------------------------------------------------------------------------------------
// BS_move (ue_mobility.py:191-271) by one warp, lane i = BS i.  Sequential over BSs; the lock test compares
// BS i's PRE-move cell with the others' CURRENT cells (ue_mobility.py:256-263) and blocks the move if any
// is within lock radius.  Returns the number of blocked BSs (all lanes).
__device__ __forceinline__ int bs_move_warp(const DevCfg &c, int &bx, int &by, int digit, int lane) {
    const int s1 = c.bs_step, s2 = 2 * c.bs_step, G = c.G;
    int px = bx, py = by;
    switch (digit) {                                                   // :221-253 with [xMin,xMax,yMin,yMax]=[1,G,1,G]
        case 0: if (bx + s1 < G) px = bx + s1; break;
        case 1: if (bx - s1 > 1) px = bx - s1; break;
        case 2: if (by + s1 < G) py = by + s1; break;
        case 3: if (by - s1 > 1) py = by - s1; break;
        case 5: if (bx + s2 < G) px = bx + s2; break;
        case 6: if (bx - s2 > 1) px = bx - s2; break;
        case 7: if (by + s2 < G) py = by + s2; break;
        case 8: if (by - s2 > 1) py = by - s2; break;
        default: break;                                                // 4 = stay
    }
    int blocked = 0;
    for (int i = 0; i < c.nBS; i++) {
        const int xi = __shfl_sync(0xffffffffu, bx, i), yi = __shfl_sync(0xffffffffu, by, i);
        const int dx = bx - xi, dy = by - yi;
        const bool col = lane < c.nBS && lane != i && (dx * dx + dy * dy <= c.lock_r2);
        const unsigned m = __ballot_sync(0xffffffffu, col);
        if (m) blocked++;
        else if (lane == i) { bx = px; by = py; }                      // :265-266
    }
    return blocked;
}

// ---------------------------------------------------------------------------------------------------------
// UpdateDroneNet's decisions for one UE (channel.py:145-176) from this pass's best server (index, SINR) and the
// SINR of the UE's current cell.  Decisions are taken in float64 on the (exactly converted) SINR values in both
// precisions, so they are a pure function of the SINR matrix the pass produced.  Updates the handover word and
// returns the serving-cell SINR the reference stores in current_BS_sinr (PRE-handover cell, channel.py:145-146).
template <typename T>
__device__ __forceinline__ T ho_decide(const DevCfg &c, int mode, int best, T bestS, T curS, uint32_t &word,
                                       int &new_out, int &did_ho) {
    const double out_thr = c.out_thr, ho_thr = c.ho_thr;
    if (mode != MODE_STEP) {
        // LTEChannel ctor / reset (channel.py:92-93,110,113-116): associate to the best server, FIFO = 1 row
        word = ho_pack(best, best, 0, 0, 1, (double)bestS <= out_thr ? 1 : 0);
        new_out = 0; did_ho = 0;
        return bestS;
    }
    int cur = word & 31, f0 = (word >> 5) & 31, f1 = (word >> 10) & 31, f2 = (word >> 15) & 31;
    int depth = (word >> 20) & 3;
    const int outp = (word >> 22) & 1;
    bool remain;                                                       // channel.py:148-155
    if (depth == 1) { f1 = best; depth = 2; remain = (f1 == f0); }
    else if (depth == 2) { f2 = best; depth = 3; remain = (f1 == f0) && (f2 == f0); }
    else { f0 = f1; f1 = f2; f2 = best; remain = (f1 == f0) && (f2 == f0); }
    const bool need = remain && (cur != best) && ((double)bestS - (double)curS > ho_thr);   // :156-159
    if (need) cur = best;                                              // :162-167
    did_ho = need ? 1 : 0;
    const int o = (double)curS <= out_thr ? 1 : 0;                     // :170
    new_out = (o && !outp) ? 1 : 0;                                    // :171-174
    word = ho_pack(cur, f0, f1, f2, depth, o);
    return curS;
}

// ---------------------------------------------------------------------------------------------------------
// float64 building blocks in the reference's operation order (no fma contraction), shared by the FP64_PARITY kernels
// and the FP32_GUARDED re-evaluation so that both produce the same bits.
//   received power P * 10^((gain - loss - fading - eq_loss) / 10) of one (UE, BS) pair (channel.py:220-247,264)
__device__ __forceinline__ double pair_power_f64(const DevCfg &c, int cx, int cy, int bx, int by, double fade) {
    const double ax = __dadd_rn(__dmul_rn((double)cx, c.grid_width), -__dmul_rn((double)bx, c.grid_width));
    const double ay = __dadd_rn(__dmul_rn((double)cy, c.grid_width), -__dmul_rn((double)by, c.grid_width));
    const double d = sqrt(__dadd_rn(__dmul_rn(ax, ax), __dmul_rn(ay, ay)));                 // :220-226
    double loss = 0.0;
    if (d > c.pl_dis) loss = __dadd_rn(c.pl_a, __dmul_rn(c.pl_b, log10(d)));                // :230-235
    const double gdb = __dadd_rn(__dadd_rn(__dadd_rn(c.ant_gain, -loss), -fade), -c.eq_loss);  // :245
    return __dmul_rn(c.P, pow(10.0, gdb / 10.0));                                           // :246, :264
}
//   10 log10(p / (N + interference)) (channel.py:266-268)
__device__ __forceinline__ double sinr_db_f64(const DevCfg &c, double p, double interf) {
    return __dmul_rn(10.0, log10(p / __dadd_rn(c.N, interf)));
}

// FP32_GUARDED re-evaluation of one UE's row by a whole warp, lane = BS, in the FP64_PARITY arithmetic: interference sums
// in index order for nBS <= 8, prefix + suffix beyond (what env_kernel<NB, true> does for the same nBS).  A lone thread
// walking the row's float64 log10 / pow chains would hold its CTA for tens of microseconds.  Rare (about 5e-4 of the
// UE-steps at the default guard) and kept out of line, off the fp32 kernels' register budget.  All 32 lanes call; every
// lane returns the UE's (best server, its SINR, SINR of the current cell).
__device__ __noinline__ void ue_row_f64_warp(const DevCfg &c, const double *fading_row, const int *bsx, const int *bsy,
                                             uint32_t genv, int u, int cx, int cy, uint32_t epoch, int cur, int *best_out,
                                             double *bestS_out, double *curS_out) {
    const int lane = threadIdx.x & 31, nBS = c.nBS, cpu = (nBS + 3) >> 2, b = lane;
    double p = 0.0;
    if (b < nBS) {
        double fade = 0.0;
        if (c.fading == FADE_INJECTED) fade = fading_row[b];
        else if (c.fading == FADE_PHILOX) {
            double z[4];
            normal4_f64(philox4x32_10(genv, (uint32_t)(u * cpu + (b >> 2)), epoch, DOM_FADING, c.k0, c.k1), z);
            const int k = b & 3;
            const double zk = k == 0 ? z[0] : (k == 1 ? z[1] : (k == 2 ? z[2] : z[3]));
            fade = __dadd_rn(c.sh_mean, __dmul_rn(c.sh_sd, zk));
        }
        p = pair_power_f64(c, cx, cy, bsx[b], bsy[b], fade);
    }
    double interf = 0.0;
    if (nBS > 8) {
        double pre = 0.0, suf = 0.0;
        for (int j = 0; j < nBS; j++) { const double pj = __shfl_sync(0xffffffffu, p, j); if (j < b) pre = __dadd_rn(pre, pj); }
        for (int j = nBS - 1; j >= 0; j--) { const double pj = __shfl_sync(0xffffffffu, p, j); if (j > b) suf = __dadd_rn(suf, pj); }
        interf = __dadd_rn(pre, suf);
    } else {
        for (int j = 0; j < nBS; j++) { const double pj = __shfl_sync(0xffffffffu, p, j); if (j != b) interf = __dadd_rn(interf, pj); }
    }
    const double S = b < nBS ? sinr_db_f64(c, p, interf) : -1.0e300;
    int best = b;
    double bestS = S;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const double oS = __shfl_xor_sync(0xffffffffu, bestS, o);
        const int oB = __shfl_xor_sync(0xffffffffu, best, o);
        if (oS > bestS || (oS == bestS && oB < best)) { bestS = oS; best = oB; }
    }
    *best_out = best; *bestS_out = bestS;
    *curS_out = __shfl_sync(0xffffffffu, S, cur);
    if (lane == 0) atomicAdd(c.guard_hits, 1ull);
}

// FP32_GUARDED: true if a decision ho_decide would take from these fp32 values could differ from the one taken on the
// float64 values, given |fp32 - float64| < guard_db / 2 per SINR: the argmax (top-2 gap), and per mode the outage
// compare and the handover-threshold compare (channel.py:141,156-159,170; reset / ctor: :92-93,116).
__device__ __forceinline__ bool guard_needed(const DevCfg &c, int mode, int best, int cur, float bestS, float second,
                                             float curS) {
    const float g = c.guard_db;
    bool need = (bestS - second) < g;
    if (mode == MODE_STEP) {
        need |= fabsf(curS - (float)c.out_thr) < g;
        need |= (best != cur) && fabsf((bestS - curS) - (float)c.ho_thr) < g;
    } else {
        need |= fabsf(bestS - (float)c.out_thr) < g;
    }
    return need;
}

// ---------------------------------------------------------------------------------------------------------
template <bool F64> struct Real { using T = float; };
template <> struct Real<true> { using T = double; };

// The channel pass of one UE against all BSs + its handover-word update.
// NB: compile-time bound on nBS (register arrays).  Returns the serving-cell SINR (pre-handover cell,
// channel.py:145-146) and updates `word`; flags receive new-outage / handover events.
template <int NB, bool F64, bool DIAG, bool GUARD>
__device__ __forceinline__ typename Real<F64>::T ue_channel_pass(const DevCfg &c, const CallArgs &a,
                                                                 const EnvShared &s, int e, uint32_t genv, int u,
                                                                 int cx, int cy, uint32_t epoch, int mode,
                                                                 uint32_t &word, int &new_out, int &did_ho, bool &pending) {
    using T = typename Real<F64>::T;
    const int nBS = c.nBS;
    const size_t pair0 = ((size_t)e * c.nUE + u) * nBS;
    T fade[NB];
    // ---- shadow fading, dB (channel.py:240) ----
    if (c.fading == FADE_INJECTED) {
#pragma unroll
        for (int b = 0; b < NB; b++) fade[b] = b < nBS ? (T)a.fading[pair0 + b] : (T)0;
    } else if (c.fading == FADE_PHILOX) {
        const int cpu = (nBS + 3) >> 2;
#pragma unroll
        for (int q = 0; q < (NB + 3) / 4; q++) {
            if (4 * q < nBS) {
                const Philox4 p = philox4x32_10(genv, (uint32_t)(u * cpu + q), epoch, DOM_FADING, c.k0, c.k1);
                T z[4];
                if constexpr (F64) normal4_f64(p, z); else normal4_f32(p, z);
#pragma unroll
                for (int k = 0; k < 4; k++)
                    if (4 * q + k < NB) {
                        if constexpr (F64) fade[4 * q + k] = __dadd_rn(c.sh_mean, __dmul_rn(c.sh_sd, z[k]));
                        else fade[4 * q + k] = fmaf(c.f_sh_sd, z[k], c.f_sh_mean);
                    }
            } else {
#pragma unroll
                for (int k = 0; k < 4; k++) if (4 * q + k < NB) fade[4 * q + k] = (T)0;
            }
        }
    } else {
#pragma unroll
        for (int b = 0; b < NB; b++) fade[b] = (T)0;
    }
    if constexpr (DIAG) {
        if (a.fading_used) {
#pragma unroll
            for (int b = 0; b < NB; b++) if (b < nBS) a.fading_used[pair0 + b] = (float)fade[b];
        }
    }

    T S[NB];   // SINR in dB per BS
    if constexpr (F64) {
        // reference operation order in float64 (channel.py:220-247, 259-269); no fma contraction
        double p[NB];
#pragma unroll
        for (int b = 0; b < NB; b++) {
            p[b] = 0.0;
            if (b < nBS) {
                p[b] = pair_power_f64(c, cx, cy, s.bsx[b], s.bsy[b], fade[b]);
            }
        }
        double pre[NB], suf[NB];
        if constexpr (NB > 8) {
            // exclude-self interference as prefix + suffix sums (never total - own: SURVEY H4)
            double acc = 0.0;
#pragma unroll
            for (int b = 0; b < NB; b++) { pre[b] = acc; acc = __dadd_rn(acc, p[b]); }
            acc = 0.0;
#pragma unroll
            for (int b = NB - 1; b >= 0; b--) { suf[b] = acc; acc = __dadd_rn(acc, p[b]); }
        }
#pragma unroll
        for (int b = 0; b < NB; b++) {
            double interf;
            if constexpr (NB > 8) interf = __dadd_rn(pre[b], suf[b]);
            else {
                interf = 0.0;                                          // np.sum over the other BSs in index order (:265)
#pragma unroll
                for (int j = 0; j < NB; j++) if (j != b && j < nBS) interf = __dadd_rn(interf, p[j]);
            }
            S[b] = b < nBS ? sinr_db_f64(c, p[b], interf) : -1.0e300;
        }
    } else {
        // fp32 log-domain form: gdb[b] = g0 - (a + k log2 q) - fade;  p = 2^(gdb*ek + log2 P);
        // S[b] = gdb[b] + Pdb - dbk * log2(N + sum_{j != b} p[j])
        float gdb[NB], p[NB];
#pragma unroll
        for (int b = 0; b < NB; b++) {
            gdb[b] = 0.f; p[b] = 0.f;
            if (b < nBS) {
                const int dx = cx - s.bsx[b], dy = cy - s.bsy[b];
                const float q = c.f_q_scale * (float)(dx * dx + dy * dy);
                const float loss = q > c.f_q_min ? fmaf(c.f_loss_k, mufu_lg2(q), c.f_loss_a) : 0.f;
                gdb[b] = c.f_g0 - loss - fade[b];
                p[b] = mufu_ex2(fmaf(gdb[b], c.f_exp_k, c.f_log2P));
            }
        }
        float pre[NB], suf[NB];
        if constexpr (NB > 4) {
            float acc = 0.f;
#pragma unroll
            for (int b = 0; b < NB; b++) { pre[b] = acc; acc += p[b]; }
            acc = 0.f;
#pragma unroll
            for (int b = NB - 1; b >= 0; b--) { suf[b] = acc; acc += p[b]; }
        }
#pragma unroll
        for (int b = 0; b < NB; b++) {
            float interf;
            if constexpr (NB > 4) interf = pre[b] + suf[b];
            else {
                interf = 0.f;
#pragma unroll
                for (int j = 0; j < NB; j++) if (j != b) interf += p[j];
            }
            S[b] = b < nBS ? (gdb[b] + c.f_Pdb) - c.f_db_k * mufu_lg2(c.f_N + interf) : -3.0e38f;
        }
    }
    if constexpr (DIAG) {
        if (a.sinr_all) {
#pragma unroll
            for (int b = 0; b < NB; b++) if (b < nBS) reinterpret_cast<T *>(a.sinr_all)[pair0 + b] = S[b];
        }
    }

    // ---- best server: first maximum (np.argmax / np.max, channel.py:141-142) ----
    int best = 0;
    T bestS = S[0];
#pragma unroll
    for (int b = 1; b < NB; b++) if (S[b] > bestS) { bestS = S[b]; best = b; }
    const int cur = word & 31;
    T curS = S[0];
#pragma unroll
    for (int b = 1; b < NB; b++) if (b == cur) curS = S[b];            // serving SINR of the PRE-handover cell
    if constexpr (!F64 && GUARD) {
        // FP32_GUARDED: is any decision of this UE within guard_db of its boundary?  Then no decision is taken here: the
        // caller queues the UE for the float64 re-evaluation after the UE loop (env_kernel, "guard phase").
        float second = -3.0e38f;
#pragma unroll
        for (int b = 0; b < NB; b++) if (b != best) second = fmaxf(second, S[b]);
        if (guard_needed(c, mode, best, cur, bestS, second, curS)) { pending = true; return curS; }
    }
    return ho_decide<T>(c, mode, best, bestS, curS, word, new_out, did_ho);
}

// ---------------------------------------------------------------------------------------------------------
// fp32 channel pass for more than 4 BSs: NB/4 adjacent lanes share one UE, each lane owns 4 BSs (one Philox call =
// its 4 fading normals), so a warp runs 32/(NB/4) UEs at once with small register arrays.  The exclude-self
// interference sum is (the other three of my four) + (the quad sums of the other lanes, gathered with shuffles) --
// never total - own; the best server is a shuffle argmax with lowest-index tie break.  All lanes of a
// group end up with the same (best server, its SINR, SINR of the current cell); the handover / outage decisions are
// taken by the caller, once per UE.  `u` must be clamped to a valid UE on every lane (shuffles need the whole warp).
template <int NB, bool DIAG, bool FULL, bool GUARD>
__device__ __forceinline__ float ue_channel_quad(const DevCfg &c, const CallArgs &a, const int (&bx4)[4], const int (&by4)[4], int e,
                                                 uint32_t genv, int u, int cx, int cy, uint32_t epoch, uint32_t word,
                                                 int &best_out, float &bestS_out, float &second_out) {
    constexpr int LPU = NB / 4;                                        // lanes per UE
    static_assert(NB == 8 || NB == 16 || NB == 32, "quad mapping");
    const int lane = threadIdx.x & 31, q = lane & (LPU - 1), gbase = lane & ~(LPU - 1);
    // FULL: nBS == NB, every lane owns four real BSs and no validity test is compiled in
    const int nBS = FULL ? NB : c.nBS, cpu = FULL ? LPU : (c.nBS + 3) >> 2, b0 = 4 * q;
    const size_t pair0 = ((size_t)e * c.nUE + u) * nBS;
    float fade[4] = {0.f, 0.f, 0.f, 0.f};
    if (c.fading == FADE_PHILOX) {
        if (FULL || q < cpu) {
            const Philox4 p = philox4x32_10(genv, (uint32_t)(u * cpu + q), epoch, DOM_FADING, c.k0, c.k1);
            float z[4];
            normal4_f32(p, z);
#pragma unroll
            for (int k = 0; k < 4; k++) fade[k] = fmaf(c.f_sh_sd, z[k], c.f_sh_mean);
        }
    } else if (c.fading == FADE_INJECTED) {
#pragma unroll
        for (int k = 0; k < 4; k++) if (FULL || b0 + k < nBS) fade[k] = (float)a.fading[pair0 + b0 + k];
    }
    // log-domain form with the constants folded on the host (DevCfg): with d2 the squared distance in cells,
    //   gP = 10 log10(P g) = f_c1 - f_loss_k log2(d2) - fade   (f_c0 - fade where the reference applies no path loss),
    //   p  = P g = 2^(gP f_exp_k),   S = gP - f_db_k log2(N + interference)
    float gP[4], p[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        gP[k] = 0.f; p[k] = 0.f;
        const int b = b0 + k;
        if (FULL || b < nBS) {
            const int dx = cx - bx4[k], dy = cy - by4[k];
            const float d2 = (float)(dx * dx + dy * dy);
            const float base = d2 > c.f_d2_min ? fmaf(-c.f_loss_k, mufu_lg2(d2), c.f_c1) : c.f_c0;
            gP[k] = base - fade[k];
            p[k] = mufu_ex2(gP[k] * c.f_exp_k);
        }
    }
    const float s01 = p[0] + p[1], s23 = p[2] + p[3], quad = s01 + s23;
    // quad sums of the OTHER lanes of the group: an exclusive butterfly (log2(LPU) shuffles) -- at every level a lane
    // adds its partner's partial sum to `others` and to its own partial sum; only sums of positive terms, never a difference
    float others = 0.f, part = quad;
#pragma unroll
    for (int o = 1; o < LPU; o <<= 1) {
        const float t = __shfl_xor_sync(0xffffffffu, part, o);
        others += t;
        part += t;
    }
    // exclude-self sums inside the quad: (others + the other three)
    const float i0 = others + (p[1] + s23), i1 = others + (p[0] + s23);
    const float i2 = others + (s01 + p[3]), i3 = others + (s01 + p[2]);
    float S0 = fmaf(-c.f_db_k, mufu_lg2(c.f_N + i0), gP[0]), S1 = fmaf(-c.f_db_k, mufu_lg2(c.f_N + i1), gP[1]);
    float S2 = fmaf(-c.f_db_k, mufu_lg2(c.f_N + i2), gP[2]), S3 = fmaf(-c.f_db_k, mufu_lg2(c.f_N + i3), gP[3]);
    if (!FULL) {
        if (b0 + 0 >= nBS) S0 = -3.0e38f;
        if (b0 + 1 >= nBS) S1 = -3.0e38f;
        if (b0 + 2 >= nBS) S2 = -3.0e38f;
        if (b0 + 3 >= nBS) S3 = -3.0e38f;
    }
    if constexpr (DIAG) {
        const float Sv[4] = {S0, S1, S2, S3};
#pragma unroll
        for (int k = 0; k < 4; k++)
            if (FULL || b0 + k < nBS) {
                if (a.sinr_all) reinterpret_cast<float *>(a.sinr_all)[pair0 + b0 + k] = Sv[k];
                if (a.fading_used) a.fading_used[pair0 + b0 + k] = fade[k];
            }
    }
    // best server = the FIRST maximum (np.argmax, channel.py:141): the group's maximum by an fmax butterfly, then the
    // lowest BS index that attains it by a min butterfly (cheaper than carrying (value, index) pairs with tie tests)
    const float hi01 = fmaxf(S0, S1), hi23 = fmaxf(S2, S3), mine = fmaxf(hi01, hi23);
    float bestS = mine;
#pragma unroll
    for (int o = 1; o < LPU; o <<= 1) bestS = fmaxf(bestS, __shfl_xor_sync(0xffffffffu, bestS, o));
    int best = S0 == bestS ? b0 : (S1 == bestS ? b0 + 1 : (S2 == bestS ? b0 + 2 : (S3 == bestS ? b0 + 3 : 255)));
#pragma unroll
    for (int o = 1; o < LPU; o <<= 1) best = min(best, __shfl_xor_sync(0xffffffffu, best, o));
    float second = -3.0e38f;
    if constexpr (GUARD) {
        // FP32_GUARDED: the runner-up (largest SINR of any OTHER BS): its gap to the maximum says whether the argmax can
        // be trusted.  The lane that holds the best server offers the second largest of its four (equal to the largest
        // if two of them tie), every other lane its largest.
        const float mine2 = fmaxf(fminf(hi01, hi23), fmaxf(fminf(S0, S1), fminf(S2, S3)));
        second = (best >> 2) == q ? mine2 : mine;
#pragma unroll
        for (int o = 1; o < LPU; o <<= 1) second = fmaxf(second, __shfl_xor_sync(0xffffffffu, second, o));
    }
    // SINR of the UE's current (pre-handover) cell lives on lane cur/4 of the group
    const int cur = word & 31, ks = cur & 3;
    const float mineS = ks == 0 ? S0 : (ks == 1 ? S1 : (ks == 2 ? S2 : S3));
    const float curS = __shfl_sync(0xffffffffu, mineS, gbase | ((cur >> 2) & (LPU - 1)));
    best_out = best;
    bestS_out = bestS;
    second_out = second;
    return curS;                                                       // the decisions (ho_decide) are the caller's
}

// ---------------------------------------------------------------------------------------------------------
// Observation, float32 [nBS+1, G, G]: plane 0 = BS counts (GetGridMap, ue_mobility.py:173-188), plane 1+b =
// UEs served by b after handover (GetCurrentAssociationMap, channel.py:387-409), indexed [plane, x, y].
//
// The dense observation is >99 % zeros (<= nBS + nUE non-zero cells of (nBS+1) G^2) and is 99 % of the bytes a
// step moves.  The zeros are streamed by the TMA engine: right after the action has been validated one
// warp issues cp.async.bulk shared->global copies (SASS: UBLKCP) of ONE constant zeroed shared-memory tile, the
// copies drain while all warps run mobility and the channel pass, and once the bulk group has completed the env's
// few counts are added with float REDs that hit the still L2-resident lines.  Tile size and CTAs per SM follow the
// measurements in profiles/r1/NOTES.md: large copies (64 KB) amortise the TMA's per-copy cost, and few resident
// CTAs keep the in-flight observations inside L2.
#ifndef UAVENV_TILE_BYTES
#define UAVENV_TILE_BYTES 73728
#endif
constexpr int TILE_BYTES = UAVENV_TILE_BYTES;      // largest bulk copy / zero tile (the launch plan may shrink it)
static_assert((TILE_BYTES % 128) == 0, "tile");

__device__ __forceinline__ void bulk_store(void *gdst, const void *ssrc, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(__cvta_generic_to_global(gdst)),
                 "r"((uint32_t)__cvta_generic_to_shared(ssrc)), "r"(bytes)
                 : "memory");
}
// same copy with an L2 eviction-priority hint (createpolicy): the observation is a write-once stream
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void bulk_store_hint(void *gdst, const void *ssrc, uint32_t bytes, uint64_t policy) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(__cvta_generic_to_global(gdst)),
                 "r"((uint32_t)__cvta_generic_to_shared(ssrc)), "r"(bytes), "l"(policy)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// all but the most recent bulk group of this thread have completed (their writes are visible)
__device__ __forceinline__ void bulk_wait_prev() { asm volatile("cp.async.bulk.wait_group 1;" ::: "memory"); }
// the source tiles of all bulk copies this thread committed have been read (they may be rewritten)
__device__ __forceinline__ void bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// fallback for observations whose per-env size / base is not 16-byte aligned (odd G): plain streaming stores
__device__ __forceinline__ void obs_zero_fill_lsu(float *obs_env, int n_cells) {
    for (int i = threadIdx.x; i < n_cells; i += blockDim.x) __stcs(obs_env + i, 0.f);
}

// The env's observation as ceil(total / tile_bytes) bulk copies of the zero tile: copy k is issued (and committed as
// its own bulk group) by lane k % 32 of the calling warp.
__device__ __forceinline__ void issue_zero_stream(float *obs_env, const float *zero_tile, uint32_t tile_bytes, uint32_t total,
                                                  int lane, uint32_t *err_flags) {
    char *dst = reinterpret_cast<char *>(obs_env);
#ifdef UAVENV_BOUNDS_CHECK
    if (tile_bytes == 0 || (tile_bytes & 15) || (total & 15) || (reinterpret_cast<uintptr_t>(dst) & 15))
        atomicOr(err_flags, ERR_BOUNDS);
#endif
    // The stream itself carries no eviction hint: with the state tagged evict-last (above) a plain stream measured
    // 1.4 % faster than an evict-first one -- its lines then survive in L2 until their REDs arrive (profiles/r1/NOTES.md).
#ifdef UAVENV_STREAM_EVICT_FIRST
    const uint64_t pol = l2_policy_evict_first();
    for (uint32_t off = (uint32_t)lane * tile_bytes; off < total; off += 32u * tile_bytes)
        bulk_store_hint(dst + off, zero_tile, min(tile_bytes, total - off), pol);
#else
    for (uint32_t off = (uint32_t)lane * tile_bytes; off < total; off += 32u * tile_bytes)
        bulk_store(dst + off, zero_tile, min(tile_bytes, total - off));
#endif
    bulk_commit();
}

// CTAs per SM the fp32 kernels are compiled for (register budget); the launch plan (uavenv.cu: plan_kernel) keeps
// (resident CTAs) x (bytes of one env's observation) well inside the 126 MB L2 so that the REDs hit.
#ifndef UAVENV_MINB
#define UAVENV_MINB 3
#endif
#ifndef UAVENV_NT_SMALL
#define UAVENV_NT_SMALL 64
#endif
// CTA size for handles without a dense observation stream and <= 64 UEs per env.  That step is a latency chain per UE
// (Philox -> Box-Muller -> lg2/ex2 -> SINR -> handover), so it wants resident warps, not registers: measured on B200 for
// 8192 envs (bench.py --obs none): 128 threads x 6 CTAs/SM 72.9 us, x 8 59.4, x 10 55.4, x 16 49.0; 64 threads x 16
// CTAs/SM (64 registers, no spills) 45.8 us, x 24 45.8, x 32 46.5.
constexpr int NT_SMALL = UAVENV_NT_SMALL;
#ifndef UAVENV_MINB_WIDE
#define UAVENV_MINB_WIDE 3
#endif
#ifndef UAVENV_MINB_SMALL
#define UAVENV_MINB_SMALL 16
#endif
#ifndef UAVENV_QUAD_UNROLL
#define UAVENV_QUAD_UNROLL 1      /* UE groups of the per-(UE, 4 BS) loop in flight per lane (tuning: 2 needs UAVENV_MINB_WIDE=2) */
#endif
constexpr int min_blocks(int nb, bool f64, int nt) { return f64 ? 1 : (nb > 8 ? UAVENV_MINB_WIDE : (nt == NT_SMALL ? UAVENV_MINB_SMALL : UAVENV_MINB)); }

// fp32 mapping for more than 4 BSs: per warp, double-buffered landing area of the next chunk's state (cp.async)
template <int NW> struct ChunkPrefetch {
    double2 xy[NW][2][32];
    uint32_t cell[NW][2][32], word[NW][2][32];
    uint8_t grp[NW][2][32];      // the UEs' group ids (ue_group is padded to whole chunks): lanes 0-7 fetch 4 bytes each
};

// The guard phase of FP32_GUARDED (see env_kernel): kept out of line so that its live ranges and the float64 call inside it
// stay off the step kernel's register budget -- the kernel runs at the 80-register cap of 3 CTAs per SM, and with the phase
// inlined two values were spilled for the whole kernel.  All threads of the CTA call; warp w takes entries w, w + NW, ...
template <int NB, int NT>
__device__ __noinline__ void guard_phase(const DevCfg &c, const CallArgs &a, EnvShared &s, int e, uint32_t genv, uint32_t epoch, int mode,
                                         bool incremental, float *obs_env, int n_cells, int32_t *lin_arr) {
    constexpr int NW = NT / 32;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nUE = c.nUE, nBS = c.nBS, G = c.G;
    const uint64_t keep = l2_policy_evict_last();
    const int n_g = s.guard_n;
    const bool listed = n_g <= GUARD_LIST;                               // else: every warp scans its share for the marks
    for (int k = warp; k < (listed ? n_g : nUE); k += NW) {
        const int u = listed ? s.guard_ue[k] : k;
        const size_t i = (size_t)e * nUE + u;
        uint32_t word = ldk(c.ho + i, keep);
        if (!(word & HO_PENDING)) continue;
        word &= ~HO_PENDING;
        const short2 cell = ldk_cell(c.ue_cell, i, keep);
        double bS, cS;
        int bb, new_out, did_ho;
        ue_row_f64_warp(c, c.fading == FADE_INJECTED ? a.fading + i * nBS : nullptr, s.bsx, s.bsy, genv, u, cell.x, cell.y, epoch,
                        (int)(word & 31), &bb, &bS, &cS);
        const double srvS = ho_decide<double>(c, mode, bb, bS, cS, word, new_out, did_ho);
        if (lane == 0) {
            stk(c.ho + i, word, keep);
            // order-independent accumulation: the list order is not deterministic, the results must be
            atomicAdd(reinterpret_cast<unsigned long long *>(&s.guard_sum), (unsigned long long)__double2ll_rn(srvS * 4294967296.0));
            if (new_out) atomicAdd(&s.red_out[0], 1);
            if (did_ho) atomicAdd(&s.red_ho[0], 1);
            const int srv = word & 31;
            const int lin = ((1 + srv) * G + cell.x) * G + cell.y;
            if (a.serving) stk(a.serving + i, (uint8_t)srv, keep);
            if (a.serving_sinr) stk(reinterpret_cast<float *>(a.serving_sinr) + i, (float)srvS, keep);
            if (a.obs_idx) stk(a.obs_idx + (size_t)e * (nUE + nBS) + u, lin, keep);
            if (incremental) obs_add(obs_env, (long long)lin, 1.f, n_cells, c.err_flags);
            if (lin_arr) lin_arr[u] = lin;
        }
    }
}

// One CTA per environment.  Warp roles (every warp also takes part in the per-UE loop):
//   last warp: bulk copies of the zero tile;  last-1: group state load / finish;  last-2: action + BS_move
template <int NB, bool F64, int NT, bool DIAG, bool GUARD>
__global__ void __launch_bounds__(NT, min_blocks(NB, F64, NT))
env_kernel(const __grid_constant__ DevCfg c, const __grid_constant__ CallArgs a) {
    using T = typename Real<F64>::T;
    constexpr int NW = NT / 32;
    // two-warp CTAs (no observation stream to issue): the group state is loaded by warp 1 while warp 0 decodes the action
    // and moves the BSs
    constexpr int WARP_TMA = NW - 1, WARP_GRP = NW == 2 ? 1 : (NW >= 2 ? NW - 2 : 0), WARP_BS = NW >= 3 ? NW - 3 : 0;
    static_assert(NW >= 1 && NW <= CTA_THREADS / 32, "CTA size");
    __shared__ EnvShared s;
    extern __shared__ __align__(128) unsigned char dyn_smem[];
    float *zero_tile = reinterpret_cast<float *>(dyn_smem);
    const uint32_t tile_bytes = (uint32_t)a.tile_bytes;   // 0: no TMA path for this launch
    const int e = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int mode = a.mode;
    if (a.env_mask && !a.env_mask[e]) return;
    const uint32_t genv = c.env_offset + (uint32_t)e;
    const int nBS = c.nBS, nUE = c.nUE, G = c.G;
    int32_t *ctr = c.ctr + (size_t)e * CTR_STRIDE;
    int tick = ctr[CTR_TICK], epoch = ctr[CTR_EPOCH], step_n = ctr[CTR_STEP];
    int agg = ctr[CTR_AGG], deagg = ctr[CTR_DEAGG];

    const int n_cells = (nBS + 1) * G * G;
    float *obs_env = (a.obs && c.obs_mode != OBS_NONE) ? a.obs + (size_t)e * n_cells : nullptr;
    const bool incremental = c.obs_mode == OBS_F32_INCREMENTAL && obs_env && mode == MODE_STEP;
    const bool full_obs = obs_env && mode != MODE_CTOR && !incremental;
    const bool bulk_ok = full_obs && tile_bytes != 0;     // host: whole float4s, aligned buffer, tile fits
    const bool group_tick = c.mobility == MOB_GROUP && mode != MODE_CTOR;   // mobile_env.py:152-155 / :122-127
    const double *inj = (group_tick && a.inject_mob) ? a.mob_u + (size_t)e * (nUE + 3 * c.nG) : nullptr;

    // ---- phase 0 (no barrier yet): zero stream (one warp) | action + BS move (one warp) | group state (one warp) ----
    if (bulk_ok && warp == WARP_TMA && !(NB > 4 && !F64)) {
        // The observation's zeros start streaming before anything has been read from HBM: the warp zeroes the tile,
        // publishes it to the async proxy and issues the env's bulk copies.  (An env whose action turns out to be
        // invalid keeps its state but its observation is zeroed.)
        float4 *z4 = reinterpret_cast<float4 *>(zero_tile);
        for (int i = lane; i < (int)(tile_bytes / 16); i += 32) z4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        fence_proxy_async_smem();          // generic-proxy writes of the tile -> visible to the async proxy
        __syncwarp();
        issue_zero_stream(obs_env, zero_tile, tile_bytes, (uint32_t)n_cells * 4u, lane, c.err_flags);
    }
    if (warp == WARP_BS) {
        // validate + decode the action (Decimal_to_Base_N, ue_mobility.py:310-336: MSB first, digit 0 <-> BS 0)
        int ok = 1, digit = 4;
        if (mode == MODE_STEP) {
            if (a.digits) {
                if (lane < nBS) { digit = a.digits[(size_t)e * nBS + lane]; ok = digit < c.n_act; }
            } else if (a.action) {
                long long cur = a.action[e];
                if (cur < 0) { ok = 0; cur = 0; }
                // lane b takes digit b = (action / n_act^(nBS-1-b)) mod n_act; more digits than BSs is an error
                // (the reference fails at ue_mobility.py:334)
                int mine = 0;
                long long rem = cur;
                if (cur <= 0x7fffffffLL) {
                    // the usual case (5^nBS fits 31 bits up to 13 BSs): 32-bit divisions -- the 64-bit ones are ~70-instruction
                    // subroutines, and this serial decode sits in front of barrier 1 (a fifth of the small kernel's lifetime)
                    unsigned int r32 = (unsigned int)cur;
                    const unsigned int n_act = (unsigned int)c.n_act;
                    for (int pos = nBS - 1; pos >= 0; pos--) {
                        const unsigned int q = r32 / n_act;
                        if (pos == lane) mine = (int)(r32 - q * n_act);
                        r32 = q;
                    }
                    rem = r32;
                } else {
                    for (int pos = nBS - 1; pos >= 0; pos--) {
                        const int d = (int)(rem % c.n_act);
                        rem /= c.n_act;
                        if (pos == lane) mine = d;
                    }
                }
                if (rem != 0) ok = 0;
                if (lane < nBS) digit = mine;
            } else ok = 0;
        }
        ok = __all_sync(0xffffffffu, ok);
        if (!ok && lane == 0) atomicOr(c.err_flags, ERR_ACTION);
        if (c.mobility == MOB_TRACE) {
            const long long row = mode == MODE_STEP ? step_n : 0;      // mobile_env.py:203 / :87,130
            if (!c.trace || row >= c.trace_T) {
                ok = 0;
                if (lane == 0) atomicOr(c.err_flags, ERR_TRACE);
            }
        }
        if (ok) {
            // BS movement (mobile_env.py:157; reset: back to the initial layout, :119)
            int bx = 0, by = 0, blocked = 0;
            if (lane < nBS) {
                if (mode == MODE_RESET) { bx = c.init_bs[2 * lane]; by = c.init_bs[2 * lane + 1]; }
                else {
                    const short2 b2 = reinterpret_cast<const short2 *>(c.bs_xy)[(size_t)e * nBS + lane];
                    bx = b2.x; by = b2.y;
                }
            }
            const int ox = bx, oy = by;
            if (mode == MODE_STEP) blocked = bs_move_warp(c, bx, by, digit, lane);
            if (lane < nBS) {
                s.bsx[lane] = bx; s.bsy[lane] = by;
                const short2 nb = make_short2((short)bx, (short)by);
                if (mode != MODE_CTOR) reinterpret_cast<short2 *>(c.bs_xy)[(size_t)e * nBS + lane] = nb;
                if (a.bs_xy_out) reinterpret_cast<short2 *>(a.bs_xy_out)[(size_t)e * nBS + lane] = nb;
                if (a.bs_digits && mode == MODE_STEP) a.bs_digits[(size_t)e * nBS + lane] = (uint8_t)digit;
                if (a.obs_idx) a.obs_idx[(size_t)e * (nUE + nBS) + nUE + lane] = bx * G + by;
                if (incremental && (ox != bx || oy != by)) {
                    obs_add(obs_env, (long long)((size_t)ox * G + oy), -1.f, n_cells, c.err_flags);
                    obs_add(obs_env, (long long)((size_t)bx * G + by), 1.f, n_cells, c.err_flags);
                }
            }
            if (lane == 0) s.blocked = blocked;
        }
        if (lane == 0) { s.ok = ok; s.guard_n = 0; s.guard_sum = 0; s.next_chunk = NW; }
    }
    if (warp == WARP_GRP && group_tick && lane < c.nG) mob_group_load(s, group_row_load(c, e, lane), lane);
    // the state of this thread's first UE is requested before the barrier: its HBM / L2 latency overlaps the BS warp's
    const uint64_t keep = l2_policy_evict_last();                      // the env's own state stays in L2 under the stream
    short2 cell_0 = make_short2(0, 0);
    uint32_t word_0 = 0u;
    double2 p_0 = make_double2(0.0, 0.0);
    double thu_0 = 0.0;
    constexpr bool CHUNKED = NB > 4 && !F64;                           // the warp-pipelined mapping below
    __shared__ __align__(16) unsigned char pf_raw[CHUNKED ? sizeof(ChunkPrefetch<NW>) : 16];
    ChunkPrefetch<NW> &pf = *reinterpret_cast<ChunkPrefetch<NW> *>(pf_raw);
    const bool need_word = mode == MODE_STEP || incremental;
    const int n_chunks = (nUE + 31) >> 5;
    auto prefetch = [&](int ch, int buf) {
        if constexpr (CHUNKED) {
            const int u = (ch << 5) + lane;
            if (ch < n_chunks && u < nUE) {
                const size_t i = (size_t)e * nUE + u;
                cp_async4(&pf.cell[warp][buf][lane], reinterpret_cast<const uint32_t *>(c.ue_cell) + i, keep);
                if (need_word) cp_async4(&pf.word[warp][buf][lane], c.ho + i, keep);
                if (group_tick) cp_async16(&pf.xy[warp][buf][lane], c.xy + i, keep);
            }
            if (ch < n_chunks && group_tick && lane < 8) cp_async4(&pf.grp[warp][buf][4 * lane], c.ue_group + (ch << 5) + 4 * lane, keep);
            cp_async_commit();
        }
    };
    if constexpr (CHUNKED) {
        prefetch(warp, 0);                                             // chunk `warp` is this warp's first
        if (bulk_ok) {                                                 // the zero tile, by everyone; barrier 1 publishes it
            float4 *z4 = reinterpret_cast<float4 *>(zero_tile);
            for (int i = tid; i < (int)(tile_bytes / 16); i += NT) z4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            fence_proxy_async_smem();          // generic-proxy writes of the tile -> visible to the async proxy
        }
    }
    if (tid < nUE && !CHUNKED) {
        const size_t i0 = (size_t)e * nUE + tid;
        cell_0 = ldk_cell(c.ue_cell, i0, keep);
        word_0 = ldk(c.ho + i0, keep);
        if (group_tick) { p_0 = ldk(c.xy + i0, keep); if (inj) thu_0 = ldk(c.th_u + i0, keep); }
    }
    __syncthreads();                                                   // barrier 1
    if (!s.ok) {                                                       // the env's state is left untouched
        if constexpr (CHUNKED) {
            cp_async_wait<0>();
            if (bulk_ok && warp == WARP_TMA)                           // (its observation is zeroed, like on the other path)
                issue_zero_stream(obs_env, zero_tile, tile_bytes, (uint32_t)n_cells * 4u, lane, c.err_flags);
        }
        if (bulk_ok && warp == WARP_TMA) bulk_wait_read_all();         // the tile must outlive the copies' reads
        return;
    }
    if (full_obs && !bulk_ok) obs_zero_fill_lsu(obs_env, n_cells);     // fallback: odd sizes / unaligned buffer

    // ---- per-UE: movement + channel pass + handover word, one thread per UE ----
    const bool aggregating = agg != 0;
    const int32_t *tr = nullptr;
    if (c.mobility == MOB_TRACE) {
        const long long row = mode == MODE_STEP ? step_n : 0;
        tr = c.trace + ((size_t)row * (c.trace_per_env ? c.E : 1) + (c.trace_per_env ? e : 0)) * nUE * 2;
    }
    double sum_sinr = 0.0;
    int cnt_out = 0, cnt_ho = 0;
    if constexpr (!F64 && NB > 4) {
        // ---- more than 4 BSs, fp32: every WARP runs the whole pipeline for chunks of 32 UEs that it draws from a shared
        // counter -- (A) lane = UE: movement -> cell;  (B) NB/4 lanes = one UE, lane = 4 BSs: channel pass with shuffle
        // reductions (ue_channel_quad), 32 / (NB/4) UEs at a time;  (C) lane = UE again: handover / outage decisions and the
        // per-UE outputs, coalesced.  Between the passes only the warp synchronises (a 512-byte staging area per warp),
        // so the warps of a CTA drift apart and the float64 movement of one overlaps the fp32 / MUFU channel pass of
        // another.  (The first version ran the passes CTA-wide with a barrier between them: 28 % of all warp cycles were
        // barrier stalls, profiles/r2/NOTES.md.)  Which warp takes which chunk is not deterministic; every per-env result
        // is: sums are accumulated in fixed point.
        __shared__ int4 wstage_all[NW * 32];
        int4 *wst = wstage_all + warp * 32;
        __shared__ float wsecond_all[GUARD ? NW * 32 : 1];             // FP32_GUARDED: runner-up SINR per staged UE
        float *wsec = wsecond_all + (GUARD ? warp * 32 : 0);
        // flat observation index of every UE for the count REDs after barrier 2 (if the env's UEs fit; else HBM is re-read)
        int32_t *lin_arr = a.cells_off >= 0 ? reinterpret_cast<int32_t *>(dyn_smem + a.cells_off) : nullptr;
        // The zero stream (the tile was zeroed by all threads before barrier 1) is issued by the TMA warp only now, after
        // barrier 1: every bulk-copy instruction waits until the SM's TMA queue accepts it, and with tens of copies per env
        // and three CTAs per SM that adds up to 40 % of the issuing warp's life (ncu: 5 % of all stall samples sit behind
        // the UBLKCP loop) -- nobody may wait for it at a barrier.  Spreading the copies over all warps, a few per chunk
        // drawn, was measured and is worse (343 vs 308 us: every warp then pays the issue latency inside its pipeline);
        // starting the stream later in the CTA's life so that the count REDs find the zeros still in L2 is worse too
        // (340 / 350 / 382 us for a start after 0 / 70 / 100 % of the chunks): profiles/r2/NOTES.md.
        // The TMA warp therefore drips the copies out, `copies_per_turn` each time it draws a chunk, sized so that the
        // stream is complete one turn before the warp's expected last chunk: the queue is nearly empty whenever it issues.
        const uint32_t obs_bytes = (uint32_t)n_cells * 4u;
        const int n_copies = (bulk_ok && warp == WARP_TMA) ? (int)((obs_bytes + tile_bytes - 1) / tile_bytes) : 0;
        int copies_done = 0;
        auto issue_copies = [&](int n) {
            const int k = copies_done + lane;
            if (lane < n && k < n_copies) {
                const uint32_t off = (uint32_t)k * tile_bytes;
                bulk_store(reinterpret_cast<char *>(obs_env) + off, zero_tile, min(tile_bytes, obs_bytes - off));
                bulk_commit();
            }
            copies_done += n;
        };
        constexpr int LPU = NB / 4, UPW = 32 / LPU;
        const int q = lane & (LPU - 1);
        const bool full_bs = nBS == NB;
        // this lane's four BS cells stay in registers for the whole call
        int bx4[4], by4[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int b = min(4 * q + k, nBS - 1);
            bx4[k] = s.bsx[b]; by4[k] = s.bsy[b];
        }
        long long acc_fix = 0;                                         // serving SINR, 2^-32 dB units
        // The state of the warp's NEXT chunk (float position, cell, handover word: 24 bytes per UE) is fetched with
        // cp.async into a double-buffered 768-byte area per warp while the current chunk is processed: under the
        // observation stream an HBM read takes microseconds, and a quarter of all warp cycles waited for these loads.
        auto grab = [&]() {
            int ch = 0;
            if (lane == 0) ch = atomicAdd(&s.next_chunk, 1);
            return __shfl_sync(0xffffffffu, ch, 0);
        };
        int ch = warp, buf = 0;                                        // first chunk: static, requested before barrier 1
        for (;;) {
            if (copies_done < n_copies) issue_copies(a.copies_per_turn);
            if (ch >= n_chunks) break;
            const int ch_next = grab();
            prefetch(ch_next, buf ^ 1);
            cp_async_wait<1>();                                        // this chunk's state has arrived (own copies) ...
            __syncwarp();                                              // ... and the group ids, fetched by lanes 0-7
            const int u0 = ch << 5, uA = u0 + lane;
            const bool liveA = uA < nUE;
            const size_t iA = (size_t)e * nUE + (liveA ? uA : 0);
            // ---- (A) lane = UE: movement
            {
                short2 cell = make_short2(0, 0);
                uint32_t word0 = 0u;
                if (liveA) {
                    const uint32_t cw = pf.cell[warp][buf][lane];
                    cell = make_short2((short)(cw & 0xffffu), (short)(cw >> 16));
                    if (need_word) word0 = pf.word[warp][buf][lane];
                    if (incremental)       // the cell of the previous step leaves its association plane
                        obs_add(obs_env, (long long)(((size_t)(1 + (word0 & 31)) * G + cell.x) * G + cell.y), -1.f, n_cells, c.err_flags);
                    if (group_tick) {
                        const double2 p = pf.xy[warp][buf][lane];
                        const double thu = inj ? ldk(c.th_u + iA, keep) : 0.0;
                        cell = mob_ue_move(c, s, e, genv, tick, aggregating, inj, uA, pf.grp[warp][buf][lane], p.x, p.y, thu, keep);
                    } else if (tr) {
                        const int2 xy = reinterpret_cast<const int2 *>(tr)[uA];
                        cell = make_short2((short)xy.x, (short)xy.y);
                    }
                    if (mode != MODE_CTOR || c.mobility == MOB_TRACE) stk_cell(c.ue_cell, iA, cell, keep);
                    if (a.ue_xy) reinterpret_cast<short2 *>(a.ue_xy)[iA] = cell;
                }
                wst[lane] = make_int4((int)((uint32_t)(uint16_t)cell.x | ((uint32_t)(uint16_t)cell.y << 16)), (int)word0, 0, 0);
            }
            __syncwarp();
            // ---- (B) NB/4 lanes = one UE, lane = 4 BSs
            const int last_slot = min(31, nUE - 1 - u0);               // shuffles need every lane on a valid UE
            constexpr int QUAD_UNROLL = UAVENV_QUAD_UNROLL;
#pragma unroll QUAD_UNROLL
            for (int j = 0; j < LPU; j++) {
                const int slot_raw = j * UPW + lane / LPU;
                const bool live = slot_raw <= last_slot;
                const int slot = live ? slot_raw : last_slot;
                const int2 sv = *reinterpret_cast<const int2 *>(wst + slot);
                const int cx = sv.x & 0xffff, cy = (int)((uint32_t)sv.x >> 16);
                uint32_t word = mode == MODE_STEP ? (uint32_t)sv.y : 0u;
                int best;
                float bestS, second;
                const float curS = full_bs
                    ? ue_channel_quad<NB, DIAG, true, GUARD>(c, a, bx4, by4, e, genv, u0 + slot, cx, cy, (uint32_t)epoch, word, best, bestS, second)
                    : ue_channel_quad<NB, DIAG, false, GUARD>(c, a, bx4, by4, e, genv, u0 + slot, cx, cy, (uint32_t)epoch, word, best, bestS, second);
                if (live && q == 0) {
                    // (best server, its SINR, current-cell SINR, FP32_GUARDED: the runner-up) for the decision pass
                    reinterpret_cast<int2 *>(wst + slot)[0].y = (int)(word | ((uint32_t)best << 23));
                    reinterpret_cast<int2 *>(wst + slot)[1] = make_int2(__float_as_int(bestS), __float_as_int(curS));
                    if (GUARD) wsec[slot] = second;
                }
            }
            __syncwarp();
            // ---- (C) lane = UE: handover / outage decisions once per UE, per-UE outputs coalesced
            if (liveA) {
                const int4 sv = wst[lane];
                const int cx = sv.x & 0xffff, cy = (int)((uint32_t)sv.x >> 16);
                uint32_t word = (uint32_t)sv.y & 0x7fffffu;
                const int best = ((uint32_t)sv.y >> 23) & 31;
                // the three margin tests once per UE here, not on every lane of the UE's group in pass B
                if (GUARD && guard_needed(c, mode, best, (int)(word & 31), __int_as_float(sv.z), wsec[lane], __int_as_float(sv.w))) {
                    // FP32_GUARDED: decision deferred to the guard phase; the handover word keeps its pre-step value + a mark
                    const int slot = atomicAdd(&s.guard_n, 1);
                    if (slot < GUARD_LIST) s.guard_ue[slot] = uA;
                    stk(c.ho + iA, word | HO_PENDING, keep);
                } else {
                    int new_out, did_ho;
                    const float srvS = ho_decide<float>(c, mode, best, __int_as_float(sv.z), __int_as_float(sv.w), word, new_out, did_ho);
                    acc_fix += __double2ll_rn((double)srvS * 4294967296.0);
                    cnt_out += new_out;
                    cnt_ho += did_ho;
                    const int srv = word & 31;
                    const int lin = ((1 + srv) * G + cx) * G + cy;
                    stk(c.ho + iA, word, keep);
                    if (lin_arr) lin_arr[uA] = lin;
                    if (a.serving) a.serving[iA] = (uint8_t)srv;
                    if (a.serving_sinr) reinterpret_cast<float *>(a.serving_sinr)[iA] = srvS;
                    if (a.obs_idx) a.obs_idx[(size_t)e * (nUE + nBS) + uA] = lin;
                    if (incremental) obs_add(obs_env, (long long)lin, 1.f, n_cells, c.err_flags);
                }
            }
            __syncwarp();                                              // the staging area is reused by the next chunk
            ch = ch_next;
            buf ^= 1;
        }
        cp_async_wait<0>();
        while (copies_done < n_copies) issue_copies(32);               // few chunks per warp: the rest of the stream
        sum_sinr = (double)acc_fix * (1.0 / 4294967296.0);             // multiples of 2^-32 below 2^20: every later sum is exact
    } else {
    for (int u = tid; u < nUE; u += NT) {
        const size_t i = (size_t)e * nUE + u;
        const bool first = u == tid;                                   // requested before barrier 1
        short2 cell = first ? cell_0 : ldk_cell(c.ue_cell, i, keep);
        uint32_t word = first ? word_0 : ldk(c.ho + i, keep);
        if (incremental) {
            // the cell of the previous step leaves its association plane
            obs_add(obs_env, (long long)(((size_t)(1 + (word & 31)) * G + cell.x) * G + cell.y), -1.f, n_cells, c.err_flags);
        }
        if (group_tick) {
            const double2 p = first ? p_0 : ldk(c.xy + i, keep);
            const double thu = inj ? (first ? thu_0 : ldk(c.th_u + i, keep)) : 0.0;
            cell = mob_ue_move(c, s, e, genv, tick, aggregating, inj, u, c.ue_group[u], p.x, p.y, thu, keep);
        } else if (tr) {
            const int2 xy = reinterpret_cast<const int2 *>(tr)[u];
            cell = make_short2((short)xy.x, (short)xy.y);
        }
        if (mode != MODE_STEP) word = 0u;
        int new_out = 0, did_ho = 0;
        bool pending = false;
        const T curS = ue_channel_pass<NB, F64, DIAG, GUARD>(c, a, s, e, genv, u, cell.x, cell.y, (uint32_t)epoch, mode, word,
                                                      new_out, did_ho, pending);
        if (mode != MODE_CTOR || c.mobility == MOB_TRACE) stk_cell(c.ue_cell, i, cell, keep);
        if (a.ue_xy) stk_cell(a.ue_xy, i, cell, keep);
        if (GUARD && pending) {
            // FP32_GUARDED: decision deferred to the guard phase; the handover word keeps its pre-step value + a mark
            const int slot = atomicAdd(&s.guard_n, 1);
            if (slot < GUARD_LIST) s.guard_ue[slot] = u;
            stk(c.ho + i, word | HO_PENDING, keep);
            continue;
        }
        stk(c.ho + i, word, keep);
        sum_sinr += (double)curS;
        cnt_out += new_out;
        cnt_ho += did_ho;
        const int srv = word & 31;
        // per-call outputs live at the same addresses every step: kept in L2 like the state
        if (a.serving) stk(a.serving + i, (uint8_t)srv, keep);
        if (a.serving_sinr) stk(reinterpret_cast<T *>(a.serving_sinr) + i, curS, keep);
        if (a.obs_idx) stk(a.obs_idx + (size_t)e * (nUE + nBS) + u, ((1 + srv) * G + cell.x) * G + cell.y, keep);
        if (incremental) obs_add(obs_env, (long long)(((size_t)(1 + srv) * G + cell.x) * G + cell.y), 1.f, n_cells, c.err_flags);
    }
    }
    if (group_tick) { mob_phase_advance(c, agg, deagg); tick++; }

    // ---- per-env reductions in a fixed order (warp tree, then warps in index order) ----
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sum_sinr += __shfl_down_sync(0xffffffffu, sum_sinr, o);
        cnt_out += __shfl_down_sync(0xffffffffu, cnt_out, o);
        cnt_ho += __shfl_down_sync(0xffffffffu, cnt_ho, o);
    }
    if (lane == 0) { s.red_sinr[warp] = sum_sinr; s.red_out[warp] = cnt_out; s.red_ho[warp] = cnt_ho; }
#ifndef UAVENV_EXP_NO_WAIT
    if (bulk_ok && warp == WARP_TMA) bulk_wait_all();                  // the zeros have landed
#endif
    __syncthreads();                                                   // barrier 2

    if constexpr (GUARD) {
        // ---- guard phase (FP32_GUARDED): the UEs whose fp32 row was within guard_db of a decision
        // boundary are re-evaluated in float64 now that the UE loop's registers are dead (the call costs the hot loop
        // nothing), one warp per UE with lane = BS: a lone thread walking the row's float64 log10 / pow chains would hold
        // its CTA for tens of microseconds.  CTA-uniform branch, taken by about 2 % of the CTAs at the default guard.
        if (s.guard_n > 0) {
            guard_phase<NB, NT>(c, a, s, e, genv, (uint32_t)epoch, mode, incremental, obs_env, n_cells,
                                (NB > 4 && a.cells_off >= 0) ? reinterpret_cast<int32_t *>(dyn_smem + a.cells_off) : nullptr);
            __syncthreads();
        }
    }
    if (warp == WARP_GRP && group_tick) mob_group_finish(c, s, e, genv, tick - 1, inj, lane);
    {
        // the dense observation gets its non-zero cells (UEs on the plane of their post-handover serving BS, BSs on
        // plane 0); thread = UE, flat indices from shared memory when the env's UEs fit
        const int32_t *lin_arr = (!F64 && NB > 4 && a.cells_off >= 0) ? reinterpret_cast<const int32_t *>(dyn_smem + a.cells_off) : nullptr;
        if (full_obs) {
            for (int u = tid; u < nUE; u += NT) {
                long long lin;
                if (lin_arr) lin = lin_arr[u];
                else {
                    const size_t i = (size_t)e * nUE + u;
                    const short2 cell = ldk_cell(c.ue_cell, i, keep);
                    const int srv = ldk(c.ho + i, keep) & 31;
                    lin = (long long)(((size_t)(1 + srv) * G + cell.x) * G + cell.y);
                }
#ifndef UAVENV_EXP_NO_RED
                obs_add(obs_env, lin, 1.f, n_cells, c.err_flags);
#endif
            }
        }
        if (full_obs && tid < nBS) obs_add(obs_env, (long long)((size_t)s.bsx[tid] * G + s.bsy[tid]), 1.f, n_cells, c.err_flags);
    }
    if (tid == 0) {
        double tot = 0.0;
        int no = 0, nh = 0;
        for (int w = 0; w < NW; w++) { tot += s.red_sinr[w]; no += s.red_out[w]; nh += s.red_ho[w]; }
        if constexpr (!F64 && GUARD) tot += (double)s.guard_sum * (1.0 / 4294967296.0);   // FP32_GUARDED re-evaluated UEs
        const double mean = tot / (double)nUE;                         // channel.py:216
        if (mode == MODE_STEP) {
            step_n += 1;                                               // mobile_env.py:179
            const double r0 = mean / 20.0, r1 = -1.0 * no / nUE;       // :165,167
            const double r = __dadd_rn(r0, r1);
            if (a.reward) a.reward[e] = r > -1.0 ? r : -1.0;           // :189
            if (a.n_out) a.n_out[e] = no;
            if (a.n_ho) a.n_ho[e] = nh;
            if (a.n_blocked) a.n_blocked[e] = s.blocked;
        } else if (mode == MODE_RESET) {
            step_n = 0;                                                // :146
        }
        if (a.mean_sinr) a.mean_sinr[e] = mean;
        if (a.done) a.done[e] = step_n >= c.max_step ? 1 : 0;          // :186-187
        if (a.step_n) a.step_n[e] = step_n;
        ctr[CTR_TICK] = tick; ctr[CTR_EPOCH] = epoch + 1; ctr[CTR_STEP] = step_n;
        ctr[CTR_AGG] = agg; ctr[CTR_DEAGG] = deagg;
    }
}

// ---------------------------------------------------------------------------------------------------------
// Coverage map, LTEChannel.GetSinrInArea (channel.py:411-433): one thread per grid cell.  For every cell (x, y) in
// [1, G)^2 the downlink SINR from the NEAREST BS (first minimum, :420-422) against all the others, a fresh fading draw
// per gain (:426,429); row / column 0 stay 0.  Fading: `injected` = the draws in the reference's call order
// ([E, (G-1)^2, nBS]: per cell the interferers in ascending index order, then the serving BS); else Philox keyed by
// (env, cell, BS, call number), or none.
template <bool F64>
__global__ void __launch_bounds__(CTA_THREADS) coverage_kernel(const __grid_constant__ DevCfg c, const int16_t *__restrict__ bs_xy,
                                                               const double *__restrict__ injected, uint32_t seq, void *out) {
    using T = typename Real<F64>::T;
    __shared__ int sbx[MAX_BS], sby[MAX_BS];
    const int e = blockIdx.y, G = c.G, nBS = c.nBS;
    if (threadIdx.x < nBS) {
        sbx[threadIdx.x] = bs_xy[((size_t)e * nBS + threadIdx.x) * 2];
        sby[threadIdx.x] = bs_xy[((size_t)e * nBS + threadIdx.x) * 2 + 1];
    }
    __syncthreads();
    const int id = blockIdx.x * CTA_THREADS + threadIdx.x;
    if (id >= G * G) return;
    const int x = id / G, y = id - x * G;
    T *o = reinterpret_cast<T *>(out) + (size_t)e * G * G + id;
    if (x == 0 || y == 0) { *o = (T)0; return; }                       // dl_sinr = np.zeros(...); range(xMin, xMax) (:413,416)
    // nearest BS: dist = gridWidth * sqrt(d2) is monotone in the integer d2, so the first minimum of d2 is np.argmin's
    int srv = 0, best = 0x7fffffff;
    for (int b = 0; b < nBS; b++) {
        const int dx = x - sbx[b], dy = y - sby[b], d2 = dx * dx + dy * dy;
        if (d2 < best) { best = d2; srv = b; }
    }
    const int cell = (x - 1) * (G - 1) + (y - 1), cpu = (nBS + 3) >> 2;
    const uint32_t genv = c.env_offset + (uint32_t)e;
    const double *inj = injected ? injected + ((size_t)e * (G - 1) * (G - 1) + cell) * nBS : nullptr;
    T p_interf = (T)0, g_srv = (T)0;
    T z[4] = {(T)0, (T)0, (T)0, (T)0};
    int k = 0;                                                          // position in the reference's draw order
    for (int b = 0; b < nBS; b++) {
        T fade = (T)0;
        if (inj) fade = (T)(b == srv ? inj[nBS - 1] : inj[k++]);
        else if (c.fading != FADE_NONE) {
            if ((b & 3) == 0) {
                const Philox4 ph = philox4x32_10(genv, (uint32_t)(cell * cpu + (b >> 2)), seq, DOM_AREA, c.k0, c.k1);
                if constexpr (F64) normal4_f64(ph, z); else normal4_f32(ph, z);
            }
            if constexpr (F64) fade = __dadd_rn(c.sh_mean, __dmul_rn(c.sh_sd, z[b & 3]));
            else fade = fmaf(c.f_sh_sd, z[b & 3], c.f_sh_mean);
        }
        const int dx = x - sbx[b], dy = y - sby[b];
        T gain;
        if constexpr (F64) {
            const double ax = __dadd_rn(__dmul_rn((double)x, c.grid_width), -__dmul_rn((double)sbx[b], c.grid_width));
            const double ay = __dadd_rn(__dmul_rn((double)y, c.grid_width), -__dmul_rn((double)sby[b], c.grid_width));
            const double d = sqrt(__dadd_rn(__dmul_rn(ax, ax), __dmul_rn(ay, ay)));
            double loss = 0.0;
            if (d > c.pl_dis) loss = __dadd_rn(c.pl_a, __dmul_rn(c.pl_b, log10(d)));
            gain = pow(10.0, __dadd_rn(__dadd_rn(__dadd_rn(c.ant_gain, -loss), -fade), -c.eq_loss) / 10.0);
            if (b == srv) g_srv = gain; else p_interf = __dadd_rn(p_interf, __dmul_rn(c.P, gain));
        } else {
            const float qd = c.f_q_scale * (float)(dx * dx + dy * dy);
            const float loss = qd > c.f_q_min ? fmaf(c.f_loss_k, mufu_lg2(qd), c.f_loss_a) : 0.f;
            gain = mufu_ex2(((c.f_g0 - loss) - fade) * c.f_exp_k);
            if (b == srv) g_srv = gain; else p_interf = fmaf((float)c.P, gain, p_interf);
        }
    }
    if constexpr (F64) *o = __dmul_rn(10.0, log10(__dmul_rn(c.P, g_srv) / __dadd_rn(c.N, p_interf)));
    else *o = c.f_db_k * (mufu_lg2((float)c.P * g_srv) - mufu_lg2(c.f_N + p_interf));
}

}  // namespace uavk
