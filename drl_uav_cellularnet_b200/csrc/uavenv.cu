// C-ABI host side of libuavenv (include/uavenv.h): handle management, configuration, kernel dispatch.
// No torch types here: device pointers are plain pointers, the stream is a cudaStream_t passed as void*.
#include "../../include/uavenv.h"

#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <new>

#include "env_kernels.cuh"

using namespace uavk;

namespace {

enum { F_XY = 0, F_THU, F_GRP, F_CTR, F_BS, F_CELL, F_HO, F_COUNT };

struct Field {
    void **dev;
    int64_t bytes;
};

}  // namespace

#define UAVENV_ALIAS_CACHE 32

struct uavenv {
    uavenv_cfg cfg;
    DevCfg d;
    int device;
    char err[512];
    int64_t launches;
    // device allocations
    void *xy, *th_u, *grp, *ctr, *bs_xy, *ue_cell, *ho, *init_bs, *ue_group, *trace, *err_flags;
    // step_host staging (device)
    void *h_action, *h_reward, *h_mean, *h_nout, *h_done, *h_idx;
    /* step_host: host buffer -> device alias (NULL = pageable) */
    const void *alias_host[UAVENV_ALIAS_CACHE];
    void *alias_dev[UAVENV_ALIAS_CACHE];
    int n_alias;
    Field fields[F_COUNT];
    bool ctor_done;
    // launch plan of the persistent step kernel
    void *kernel, *kernel_diag;
    int64_t area_calls;          // coverage-map call counter (Philox sequence number)
    int threads, grid, tile_bytes, ctas_per_sm, cells_off;
    size_t dyn_smem;
    bool tiles_ok;
};

namespace {

int fail(uavenv_t *h, int code, const char *fmt, const char *detail = "") {
    if (h) snprintf(h->err, sizeof(h->err), fmt, detail);
    return code;
}

#define CU(h, call)                                                                  \
    do {                                                                             \
        cudaError_t e_ = (call);                                                     \
        if (e_ != cudaSuccess) return fail((h), UAVENV_ECUDA, #call ": %s", cudaGetErrorString(e_)); \
    } while (0)

int64_t pad8(int64_t b) { return (b + 7) & ~(int64_t)7; }

int use_device(uavenv_t *h) {
    int cur = -1;
    CU(h, cudaGetDevice(&cur));
    if (cur != h->device) CU(h, cudaSetDevice(h->device));
    return UAVENV_OK;
}

typedef void (*env_kernel_fn)(const DevCfg, const CallArgs);

template <bool F64, int NT, bool DIAG, bool GUARD>
env_kernel_fn pick_nb(int nBS) {
    if (nBS <= 4) return env_kernel<4, F64, NT, DIAG, GUARD>;
    if (nBS <= 8) return env_kernel<8, F64, NT, DIAG, GUARD>;
    if (nBS <= 16) return env_kernel<16, F64, NT, DIAG, GUARD>;
    return env_kernel<32, F64, NT, DIAG, GUARD>;
}

/* Launch plan of the step kernel (one CTA per env): the zero tile the TMA warp streams from and the CTAs per SM.
 * Measured on B200 (profiles/r1/NOTES.md): 64 KB bulk copies amortise the TMA's per-copy cost (7.0 TB/s against
 * 6.3 TB/s with 16 KB copies), and the observations of all resident CTAs must stay inside the 126 MB L2 or the
 * count REDs that follow the zeros miss it (-10 %).  The tile is the dynamic shared memory of the kernel, so its
 * size also sets the residency: 64 KB -> 3 CTAs per SM -> 444 x 200 KB = 89 MB in flight at the reference sizes. */
int plan_kernel(uavenv_t *h) {
    const bool f64 = h->cfg.precision == UAVENV_PREC_FP64_PARITY;
    /* Without a dense observation to stream (obs NONE / INCREMENTAL: the policy reads obs_idx) the step is pure
     * latency-bound arithmetic: small envs then run in NT_SMALL-thread CTAs, many more resident per SM. */
    const bool small = !f64 && h->cfg.obs_mode != UAVENV_OBS_F32 && h->d.nUE <= 64;
    h->threads = small ? NT_SMALL : CTA_THREADS;
    /* two builds of every kernel: the lean one, and one that can also write the full SINR matrix / the applied fading
     * (uavenv_out.sinr_all / fading_used; selected on the first call that passes either pointer) */
    const bool guard = h->cfg.precision == UAVENV_PREC_FP32_GUARDED;
    const int nb = h->d.nBS;
    if (small && guard) { h->kernel = (void *)pick_nb<false, NT_SMALL, false, true>(nb); h->kernel_diag = (void *)pick_nb<false, NT_SMALL, true, true>(nb); }
    else if (small) { h->kernel = (void *)pick_nb<false, NT_SMALL, false, false>(nb); h->kernel_diag = (void *)pick_nb<false, NT_SMALL, true, false>(nb); }
    else if (f64) { h->kernel = (void *)pick_nb<true, CTA_THREADS, false, false>(nb); h->kernel_diag = (void *)pick_nb<true, CTA_THREADS, true, false>(nb); }
    else if (guard) { h->kernel = (void *)pick_nb<false, CTA_THREADS, false, true>(nb); h->kernel_diag = (void *)pick_nb<false, CTA_THREADS, true, true>(nb); }
    else { h->kernel = (void *)pick_nb<false, CTA_THREADS, false, false>(nb); h->kernel_diag = (void *)pick_nb<false, CTA_THREADS, true, false>(nb); }
    const int64_t n_cells = (int64_t)(h->d.nBS + 1) * h->d.G * h->d.G;
    int dev_smem = 0, n_sm = 0;
    CU(h, cudaDeviceGetAttribute(&dev_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, h->device));
    CU(h, cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, h->device));
    cudaFuncAttributes fa;
    CU(h, cudaFuncGetAttributes(&fa, (const void *)h->kernel));
    /* the TMA path needs whole float4s per env and 32-bit byte offsets */
    h->tiles_ok = h->cfg.obs_mode == UAVENV_OBS_F32 && (n_cells & 3) == 0 && n_cells * 4 < 0x7fffffffLL;
    /* fp32 kernels with more than 4 BSs keep every UE's flat observation index in shared memory for the count REDs
     * that follow the zero stream (4 bytes per UE); if the env's UEs do not fit, the cells are re-read from HBM */
    int64_t cells_bytes = (!f64 && h->d.nBS > 4) ? (((int64_t)h->d.nUE * 4 + 127) & ~(int64_t)127) : 0;
    if (cells_bytes > 49152) cells_bytes = 0;
    /* Shared-memory budget per CTA for three resident CTAs per SM (what the fp32 kernels are compiled for and what
     * measured best: NOTES.md); the zero tile takes what the staging area leaves, in equal copies of <= TILE_BYTES. */
    int sm_smem = 0;
    CU(h, cudaDeviceGetAttribute(&sm_smem, cudaDevAttrMaxSharedMemoryPerMultiprocessor, h->device));
    const int64_t fixed = (int64_t)fa.sharedSizeBytes + 1024;               /* static + per-CTA reservation */
    /* resident CTAs per SM the kernel is compiled for (env_kernels.cuh: min_blocks) */
    const int target_ctas = (!f64 && h->d.nBS > 8) ? UAVENV_MINB_WIDE : 3;
    int64_t budget = f64 ? dev_smem - fixed : sm_smem / target_ctas - fixed;
    if (budget > dev_smem - fixed) budget = dev_smem - fixed;
    int64_t tile = TILE_BYTES;
    if (const char *ev = getenv("UAVENV_TILE_BYTES")) { const long v = atol(ev); if (v >= 128 && v % 128 == 0) tile = v; }
    if (tile > budget - cells_bytes) tile = (budget - cells_bytes) / 128 * 128;
    if (tile < 16384) {                                                     /* no room at 3 CTAs/SM: fewer, larger CTAs */
        tile = TILE_BYTES;
        if (tile > dev_smem - fixed - cells_bytes) tile = (dev_smem - fixed - cells_bytes) / 128 * 128;
    }
    {   /* equal copies: ceil(total / tile) of them, each rounded up to 128 B (reference sizes: 3 x 66 688 B) */
        const int64_t total = n_cells * 4, n_ops = (total + tile - 1) / tile;
        tile = ((total + n_ops - 1) / n_ops + 127) / 128 * 128;
    }
    h->tile_bytes = h->tiles_ok ? (int)tile : 0;
    h->cells_off = cells_bytes ? h->tile_bytes : -1;
    h->dyn_smem = (size_t)h->tile_bytes + (h->cells_off >= 0 ? (size_t)cells_bytes : 0);
    if (h->dyn_smem) {
        CU(h, cudaFuncSetAttribute((const void *)h->kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->dyn_smem));
        CU(h, cudaFuncSetAttribute((const void *)h->kernel_diag, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->dyn_smem));
    }
    int per_sm = 0;
    CU(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (const void *)h->kernel, h->threads, h->dyn_smem));
    if (per_sm < 1) return fail(h, UAVENV_ECUDA, "step kernel does not fit on an SM%s");
    h->ctas_per_sm = per_sm;
    h->grid = h->d.E;
    (void)n_sm;
    return UAVENV_OK;
}

int run_env(uavenv_t *h, int mode, const uavenv_in *in, const uavenv_out *out, void *stream) {
    if (!h) return UAVENV_EINVAL;
    int rc = use_device(h);
    if (rc) return rc;
    CallArgs a;
    memset(&a, 0, sizeof(a));
    a.mode = mode;
    if (in) {
        a.action = in->action; a.digits = in->digits; a.fading = in->fading; a.mob_u = in->mob_uniforms;
        a.env_mask = mode == MODE_RESET ? in->env_mask : nullptr;
        a.inject_mob = in->mob_uniforms != nullptr;
    }
    if (out) {
        a.obs = (float *)out->obs; a.reward = out->reward; a.mean_sinr = out->mean_sinr; a.n_out = out->n_out;
        a.n_ho = out->n_ho; a.n_blocked = out->n_blocked; a.done = out->done; a.step_n = out->step_n;
        a.serving = out->serving; a.serving_sinr = out->serving_sinr; a.sinr_all = out->sinr_all;
        a.fading_used = out->fading_used; a.ue_xy = out->ue_xy; a.bs_xy_out = out->bs_xy; a.bs_digits = out->bs_digits;
        a.obs_idx = out->obs_idx;
    }
    if (mode == MODE_STEP && !a.action && !a.digits) return fail(h, UAVENV_EACTION, "step needs in->action or in->digits%s");
    if (h->cfg.fading == UAVENV_FADE_INJECTED && !a.fading)
        return fail(h, UAVENV_EINVAL, "fading mode is INJECTED but in->fading is NULL%s");
    if (h->cfg.mobility == UAVENV_MOB_TRACE && !h->trace) return fail(h, UAVENV_ETRACE, "trace mode but no trace set%s");
    if (h->cfg.obs_mode == UAVENV_OBS_NONE) a.obs = nullptr;
    h->d.trace = (const int32_t *)h->trace;
    /* the TMA warp streams the observation's zeros when the plan allows it and the buffer is float4-aligned */
    a.tile_bytes = (h->tiles_ok && a.obs && mode != MODE_CTOR && ((uintptr_t)a.obs & 15) == 0) ? h->tile_bytes : 0;
    a.cells_off = h->cells_off;
    {   /* chunked mapping: the TMA warp issues this many bulk copies of the zero stream per chunk it draws, so that the
         * stream is out one turn before its expected last chunk (UAVENV_STREAM_TURNS overrides the number of turns) */
        static int turns_env = -1;
        if (turns_env < 0) { const char *ev = getenv("UAVENV_STREAM_TURNS"); turns_env = ev ? atoi(ev) : 0; }
        const int64_t n_chunks = (h->d.nUE + 31) / 32, per_warp = n_chunks / (h->threads / 32);
        int64_t turns = turns_env > 0 ? turns_env : per_warp - 1;
        if (turns < 1) turns = 1;
        const int64_t total = (int64_t)(h->d.nBS + 1) * h->d.G * h->d.G * 4;
        const int64_t n_copies = a.tile_bytes ? (total + a.tile_bytes - 1) / a.tile_bytes : 0;
        const int64_t cpt = (n_copies + turns - 1) / turns;
        a.copies_per_turn = (int)(cpt < 1 ? 1 : (cpt > 32 ? 32 : cpt));
    }
    const env_kernel_fn k = (env_kernel_fn)((a.sinr_all || a.fading_used) ? h->kernel_diag : h->kernel);
    k<<<h->grid, h->threads, h->dyn_smem, (cudaStream_t)stream>>>(h->d, a);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(h, UAVENV_ECUDA, "env_kernel launch: %s", cudaGetErrorString(e));
    h->launches++;
    if (mode == MODE_CTOR) h->ctor_done = true;
    return UAVENV_OK;
}

}  // namespace

extern "C" {

const char *uavenv_version(void) { return "uavenv-b200 0.1 (sm_100a)"; }

int uavenv_cfg_default(uavenv_cfg *c, int32_t n_envs, int32_t n_bs, int32_t n_ue, int32_t grid_n) {
    if (!c) return UAVENV_EINVAL;
    memset(c, 0, sizeof(*c));
    c->n_envs = n_envs; c->n_bs = n_bs; c->n_ue = n_ue; c->grid_n = grid_n;
    c->mobility = UAVENV_MOB_GROUP; c->fading = UAVENV_FADE_PHILOX; c->precision = UAVENV_PREC_FP32_FAST;
    c->obs_mode = UAVENV_OBS_F32;
    c->seed = 0; c->env_offset = 0; c->device = 0;
    c->max_step = 2000;        /* mobile_env.py:18 */
    c->n_act = 5;              /* mobile_env.py:21 */
    c->bs_step = 2;            /* mobile_env.py:32 */
    c->min_bs_dist = 2;        /* mobile_env.py:28 */
    c->warmup_ticks = 200;     /* mobile_env.py:77-79 */
    /* groups: the reference hard-codes [10,10,10,10] (mobile_env.py:76); other sizes: near-even split over
     * 4 groups (n_bs <= 4) or min(32, n_bs) groups */
    int ng = n_bs <= 4 ? 4 : (n_bs < UAVENV_MAX_GROUPS ? n_bs : UAVENV_MAX_GROUPS);
    if (ng > n_ue) ng = n_ue > 0 ? n_ue : 1;
    c->n_groups = ng;
    for (int g = 0; g < ng; g++) c->group_sizes[g] = n_ue / ng + (g < n_ue % ng ? 1 : 0);
    c->has_init_bs = 0;
    c->aggregating0 = 200; c->deaggregating0 = 100;      /* ue_mobility.py:450-451 */
    c->deaggregating_len = 100; c->aggregating_len = 10; /* ue_mobility.py:473,487 */
    c->grid_width = 5;         /* channel.py:21 */
    c->p_bs_dbm = 20; c->noise_dbm = -121;               /* channel.py:36,40 */
    c->pl_a = 38; c->pl_b = 30; c->pl_dis = 0;           /* channel.py:46-48 */
    c->ant_gain = 2; c->eq_loss = 0;                     /* channel.py:50,52 */
    c->shadow_mean = 0; c->shadow_sd = 2;                /* channel.py:54-55 */
    c->ho_thresh_db = 1;       /* channel.py:82 */
    c->out_thresh_db = 0;      /* channel.py:7 */
    c->v_min = 0; c->v_max = 1; c->aggregation = 0.8;    /* mobile_env.py:76 */
    c->guard_db = 1e-3;
    return UAVENV_OK;
}

/* initial BS layout: the reference's four corners-of-quadrants for nBS = 4 (mobile_env.py:49-50); otherwise
 * the first nBS points of a ceil(sqrt(nBS))^2 lattice (documented extension) */
static void default_bs_layout(const uavenv_cfg *c, int16_t *xy) {
    const int G = c->grid_n, n = c->n_bs;
    if (n == 4) {
        const int lo = (int)(G / 4.0), hi = (int)(G * 3 / 4.0);
        const int xs[4] = {lo, lo, hi, hi}, ys[4] = {lo, hi, lo, hi};
        for (int b = 0; b < 4; b++) { xy[2 * b] = (int16_t)xs[b]; xy[2 * b + 1] = (int16_t)ys[b]; }
        return;
    }
    int side = 1;
    while (side * side < n) side++;
    for (int b = 0; b < n; b++) {
        int x = (b / side + 1) * G / (side + 1), y = (b % side + 1) * G / (side + 1);
        if (x < 2) x = 2;
        if (y < 2) y = 2;
        xy[2 * b] = (int16_t)x; xy[2 * b + 1] = (int16_t)y;
    }
}

int uavenv_create(const uavenv_cfg *cfg, uavenv_t **out) {
    if (!cfg || !out) return UAVENV_EINVAL;
    *out = nullptr;
    uavenv_t *h = new (std::nothrow) uavenv_t;
    if (!h) return UAVENV_ENOMEM;
    memset(h, 0, sizeof(*h));
    *out = h;   /* returned even on failure so that uavenv_last_error() can be read; caller destroys it */
    h->cfg = *cfg;
    h->device = cfg->device;
    const int E = cfg->n_envs, nBS = cfg->n_bs, nUE = cfg->n_ue, G = cfg->grid_n, nG = cfg->n_groups;
    if (E < 1 || nBS < 1 || nBS > UAVENV_MAX_BS || nUE < 1 || G < 4 || G > 32767)
        return fail(h, UAVENV_EINVAL, "bad sizes (need n_envs>=1, 1<=n_bs<=32, n_ue>=1, 4<=grid_n<=32767)%s");
    if ((int64_t)(nBS + 1) * G * G >= 0x7fffffffLL)
        return fail(h, UAVENV_EINVAL, "(n_bs+1)*grid_n^2 must stay below 2^31 (observation cells are indexed with int32)%s");
    if (cfg->n_act < 2 || cfg->n_act > 9) return fail(h, UAVENV_EINVAL, "n_act must be in [2,9]%s");
    if (cfg->mobility != UAVENV_MOB_GROUP && cfg->mobility != UAVENV_MOB_TRACE)
        return fail(h, UAVENV_EINVAL, "mobility model not defined%s");   /* sys.exit at mobile_env.py:91 */
    if (cfg->fading < 0 || cfg->fading > 2 || (cfg->precision < 0 || cfg->precision > 2) ||
        (cfg->obs_mode != UAVENV_OBS_NONE && cfg->obs_mode != UAVENV_OBS_F32 && cfg->obs_mode != UAVENV_OBS_F32_INCREMENTAL))
        return fail(h, UAVENV_EINVAL, "bad fading / precision / obs_mode%s");
    if (cfg->precision == UAVENV_PREC_FP32_GUARDED && !(cfg->guard_db > 0 && cfg->guard_db <= 1))
        return fail(h, UAVENV_EINVAL, "guard_db must be in (0, 1] dB%s");
    if (nG < 1 || nG > UAVENV_MAX_GROUPS) return fail(h, UAVENV_EINVAL, "n_groups must be in [1,32]%s");
    int64_t tot = 0;
    for (int g = 0; g < nG; g++) {
        if (cfg->group_sizes[g] < 0) return fail(h, UAVENV_EINVAL, "negative group size%s");
        tot += cfg->group_sizes[g];
    }
    if (tot != nUE) return fail(h, UAVENV_EINVAL, "group_sizes must sum to n_ue%s");
    if (cfg->env_offset < 0 || cfg->env_offset + E > 0xffffffffLL) return fail(h, UAVENV_EINVAL, "env_offset out of range%s");

    CU(h, cudaSetDevice(h->device));
    DevCfg &d = h->d;
    d.E = E; d.nBS = nBS; d.nUE = nUE; d.G = G; d.nG = nG;
    d.mobility = cfg->mobility; d.fading = cfg->fading; d.obs_mode = cfg->obs_mode;
    d.max_step = cfg->max_step; d.n_act = cfg->n_act; d.bs_step = cfg->bs_step;
    const int lock = cfg->min_bs_dist + cfg->bs_step;                  /* mobile_env.py:157 */
    d.lock_r2 = lock * lock;
    d.deagg_len = cfg->deaggregating_len; d.agg_len = cfg->aggregating_len;
    d.k0 = (uint32_t)cfg->seed; d.k1 = (uint32_t)(cfg->seed >> 32);
    d.env_offset = (uint32_t)cfg->env_offset;
    d.grid_width = cfg->grid_width;
    d.P = pow(10.0, cfg->p_bs_dbm / 10.0) * 1e-3;                      /* channel.py:58 */
    d.N = pow(10.0, cfg->noise_dbm / 10.0) * 1e-3;                     /* channel.py:59 */
    d.pl_a = cfg->pl_a; d.pl_b = cfg->pl_b; d.pl_dis = cfg->pl_dis;
    d.ant_gain = cfg->ant_gain; d.eq_loss = cfg->eq_loss;
    d.sh_mean = cfg->shadow_mean; d.sh_sd = cfg->shadow_sd;
    d.ho_thr = cfg->ho_thresh_db; d.out_thr = cfg->out_thresh_db;
    d.v_min = cfg->v_min; d.v_max = cfg->v_max; d.aggr = cfg->aggregation;
    d.max_xy = (double)G; d.fl_max = (double)G;                        /* dimensions=(G,G); FL_MAX = max(dimensions) */
    d.f_q_scale = (float)(cfg->grid_width * cfg->grid_width);
    d.f_q_min = (float)(cfg->pl_dis * cfg->pl_dis);
    d.f_g0 = (float)(cfg->ant_gain - cfg->eq_loss);
    d.f_loss_a = (float)cfg->pl_a;
    d.f_loss_k = (float)(cfg->pl_b * 0.5 * log10(2.0));
    d.f_exp_k = (float)(log2(10.0) / 10.0);
    d.f_log2P = (float)log2(d.P);
    d.f_Pdb = (float)(10.0 * log10(d.P));
    d.f_db_k = (float)(10.0 * log10(2.0));
    d.f_N = (float)d.N; d.f_sh_mean = (float)cfg->shadow_mean; d.f_sh_sd = (float)cfg->shadow_sd;
    {
        const double q_scale = cfg->grid_width * cfg->grid_width, loss_k = cfg->pl_b * 0.5 * log10(2.0), pdb = 10.0 * log10(d.P);
        d.f_d2_min = (float)(cfg->pl_dis * cfg->pl_dis / q_scale);
        d.f_c1 = (float)(cfg->ant_gain - cfg->eq_loss - cfg->pl_a - loss_k * log2(q_scale) + pdb);
        d.f_c0 = (float)(cfg->ant_gain - cfg->eq_loss + pdb);
    }
    d.guard_db = cfg->precision == UAVENV_PREC_FP32_GUARDED ? (float)cfg->guard_db : 0.f;

    const int64_t nu = (int64_t)E * nUE;
    Field f[F_COUNT] = {
        {&h->xy, nu * 16}, {&h->th_u, nu * 8}, {&h->grp, (int64_t)E * 6 * nG * 8},
        {&h->ctr, (int64_t)E * CTR_STRIDE * 4}, {&h->bs_xy, (int64_t)E * nBS * 4}, {&h->ue_cell, nu * 4}, {&h->ho, nu * 4}};
    for (int i = 0; i < F_COUNT; i++) {
        h->fields[i] = f[i];
        CU(h, cudaMalloc(f[i].dev, (size_t)f[i].bytes));
        CU(h, cudaMemset(*f[i].dev, 0, (size_t)f[i].bytes));
    }
    CU(h, cudaMalloc(&h->init_bs, nBS * 4));
    CU(h, cudaMalloc(&h->ue_group, (size_t)(nUE + 31) / 32 * 32));       /* whole chunks of 32: fetched 4 bytes per lane */
    CU(h, cudaMemset(h->ue_group, 0, (size_t)(nUE + 31) / 32 * 32));
    CU(h, cudaMalloc(&h->err_flags, 16));                 /* [0] sticky flags, [8..16) FP32_GUARDED re-evaluation count */
    CU(h, cudaMemset(h->err_flags, 0, 16));
    CU(h, cudaMalloc(&h->h_action, (size_t)E * 8));
    CU(h, cudaMalloc(&h->h_reward, (size_t)E * 8));
    CU(h, cudaMalloc(&h->h_mean, (size_t)E * 8));
    CU(h, cudaMalloc(&h->h_nout, (size_t)E * 4));
    CU(h, cudaMalloc(&h->h_done, (size_t)E));

    int16_t bs[UAVENV_MAX_BS * 2];
    if (cfg->has_init_bs) {
        for (int b = 0; b < nBS; b++) {
            const int x = cfg->init_bs_xy[2 * b], y = cfg->init_bs_xy[2 * b + 1];
            if (x < 0 || x >= G || y < 0 || y >= G) return fail(h, UAVENV_EINVAL, "init_bs_xy outside the grid%s");
            bs[2 * b] = (int16_t)x; bs[2 * b + 1] = (int16_t)y;
        }
    } else default_bs_layout(cfg, bs);
    for (int b = 0; b < nBS; b++) { h->cfg.init_bs_xy[2 * b] = bs[2 * b]; h->cfg.init_bs_xy[2 * b + 1] = bs[2 * b + 1]; }
    h->cfg.has_init_bs = 1;
    CU(h, cudaMemcpy(h->init_bs, bs, nBS * 4, cudaMemcpyHostToDevice));
    /* every env starts on the initial layout (mobile_env.py:58) */
    {
        int16_t *all = (int16_t *)malloc((size_t)E * nBS * 4);
        if (!all) return fail(h, UAVENV_ENOMEM, "host malloc%s");
        for (int e = 0; e < E; e++) memcpy(all + (size_t)e * nBS * 2, bs, nBS * 4);
        cudaError_t ce = cudaMemcpy(h->bs_xy, all, (size_t)E * nBS * 4, cudaMemcpyHostToDevice);
        free(all);
        CU(h, ce);
    }
    {
        uint8_t *gr = (uint8_t *)malloc(nUE);
        if (!gr) return fail(h, UAVENV_ENOMEM, "host malloc%s");
        int u = 0;
        for (int g = 0; g < nG; g++) for (int k = 0; k < cfg->group_sizes[g]; k++) gr[u++] = (uint8_t)g;
        cudaError_t ce = cudaMemcpy(h->ue_group, gr, nUE, cudaMemcpyHostToDevice);
        free(gr);
        CU(h, ce);
    }
    d.xy = (double2 *)h->xy; d.th_u = (double *)h->th_u; d.grp = (double *)h->grp;
    d.ctr = (int32_t *)h->ctr; d.bs_xy = (int16_t *)h->bs_xy; d.ue_cell = (int16_t *)h->ue_cell; d.ho = (uint32_t *)h->ho;
    d.init_bs = (const int16_t *)h->init_bs; d.ue_group = (const uint8_t *)h->ue_group;
    d.trace = nullptr; d.trace_T = 0; d.trace_per_env = 0;
    d.err_flags = (uint32_t *)h->err_flags;
    d.guard_hits = (unsigned long long *)((char *)h->err_flags + 8);

    {
        int rc = plan_kernel(h);
        if (rc) return rc;
    }
    if (cfg->mobility == UAVENV_MOB_GROUP) {
        if (cfg->warmup_ticks >= 0) {
            /* mobility init + warm-up + the constructor's positions (mobile_env.py:76-79,93-97) */
            mob_init_kernel<<<E, CTA_THREADS>>>(d, cfg->warmup_ticks, cfg->aggregating0, cfg->deaggregating0);
            CU(h, cudaGetLastError());
            h->launches++;
        }
        /* warmup_ticks < 0: the caller loads the mobility state with uavenv_set_state */
        if (cfg->fading != UAVENV_FADE_INJECTED && cfg->warmup_ticks >= 0) {
            int rc = run_env(h, MODE_CTOR, nullptr, nullptr, nullptr);  /* LTEChannel ctor pass, channel.py:92-93,110 */
            if (rc) return rc;
        }
    }
    CU(h, cudaDeviceSynchronize());
    return UAVENV_OK;
}

void uavenv_destroy(uavenv_t *h) {
    if (!h) return;
    int cur = -1;
    cudaGetDevice(&cur);
    if (cur != h->device) cudaSetDevice(h->device);
    void *p[] = {h->xy, h->th_u, h->grp, h->ctr, h->bs_xy, h->ue_cell, h->ho, h->init_bs, h->ue_group,
                 h->trace, h->err_flags, h->h_action, h->h_reward, h->h_mean, h->h_nout, h->h_done, h->h_idx};
    for (void *q : p) if (q) cudaFree(q);
    delete h;
}

int uavenv_set_trace(uavenv_t *h, const int32_t *xy_host, int64_t T, int32_t per_env) {
    if (!h || !xy_host || T < 1) return fail(h, UAVENV_EINVAL, "set_trace: bad arguments%s");
    int rc = use_device(h);
    if (rc) return rc;
    const int64_t n = T * (per_env ? h->d.E : 1) * h->d.nUE * 2;
    for (int64_t i = 0; i < n; i++)
        if (xy_host[i] < 0 || xy_host[i] >= h->d.G) return fail(h, UAVENV_ETRACE, "trace cell outside the grid%s");
    CU(h, cudaDeviceSynchronize());
    if (h->trace) { CU(h, cudaFree(h->trace)); h->trace = nullptr; }
    CU(h, cudaMalloc(&h->trace, (size_t)n * 4));
    CU(h, cudaMemcpy(h->trace, xy_host, (size_t)n * 4, cudaMemcpyHostToDevice));
    h->d.trace = (const int32_t *)h->trace; h->d.trace_T = T; h->d.trace_per_env = per_env ? 1 : 0;
    return UAVENV_OK;
}

int uavenv_ctor_pass(uavenv_t *h, const uavenv_in *in, const uavenv_out *out, void *stream) {
    return run_env(h, MODE_CTOR, in, out, stream);
}
int uavenv_reset(uavenv_t *h, const uavenv_in *in, const uavenv_out *out, void *stream) {
    return run_env(h, MODE_RESET, in, out, stream);
}
int uavenv_step(uavenv_t *h, const uavenv_in *in, const uavenv_out *out, void *stream) {
    return run_env(h, MODE_STEP, in, out, stream);
}

/* device alias of a pinned (page-locked, mapped) host buffer, or NULL for pageable memory.  The answer is cached per
 * handle (a hot loop passes the same few buffers every step and cudaPointerGetAttributes costs about a microsecond);
 * a caller that unregisters or frees a pinned buffer and reuses its address as pageable memory must make a new handle. */
static void *pinned_alias(uavenv_t *h, const void *p) {
    if (!p) return nullptr;
    for (int i = 0; i < h->n_alias; i++)
        if (h->alias_host[i] == p) return h->alias_dev[i];
    cudaPointerAttributes at;
    void *dev = nullptr;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) cudaGetLastError();
    else if (at.type == cudaMemoryTypeHost) dev = at.devicePointer;
    if (h->n_alias < UAVENV_ALIAS_CACHE) {
        h->alias_host[h->n_alias] = p;
        h->alias_dev[h->n_alias++] = dev;
    }
    return dev;
}

static int step_host_impl(uavenv_t *h, const int64_t *action_host, void *obs_dev, double *reward_host, uint8_t *done_host,
                          double *mean_sinr_host, int32_t *n_out_host, int32_t *obs_idx_host, void *stream) {
    if (!h || !action_host) return fail(h, UAVENV_EINVAL, "step_host: action_host is NULL%s");
    int rc = use_device(h);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t E = (size_t)h->d.E;
    /* Pinned (mapped) host buffers are used in place: the BS warp of every CTA reads its env's action over PCIe
     * (the observation stream does not wait for it) and the kernel writes the results straight into the caller's
     * buffers with posted PCIe writes -- one launch and one synchronise per step, no copies on the stream.
     * Pageable buffers go through device staging and cudaMemcpyAsync.  UAVENV_HOST_ACTION_COPY=1 forces the staged
     * copy for the actions (tuning). */
    static const bool force_copy = getenv("UAVENV_HOST_ACTION_COPY") != nullptr;
    const void *act = force_copy ? nullptr : pinned_alias(h, action_host);
    if (!act) {
        CU(h, cudaMemcpyAsync(h->h_action, action_host, E * 8, cudaMemcpyHostToDevice, st));
        act = h->h_action;
    }
    uavenv_in in;
    memset(&in, 0, sizeof(in));
    in.action = (const int64_t *)act;
    void *rw = pinned_alias(h, reward_host), *dn = pinned_alias(h, done_host), *ms = pinned_alias(h, mean_sinr_host),
         *no = pinned_alias(h, n_out_host);
    uavenv_out out;
    memset(&out, 0, sizeof(out));
    out.obs = obs_dev;
    out.reward = reward_host ? (double *)(rw ? rw : h->h_reward) : nullptr;
    out.done = done_host ? (uint8_t *)(dn ? dn : h->h_done) : nullptr;
    out.mean_sinr = mean_sinr_host ? (double *)(ms ? ms : h->h_mean) : nullptr;
    out.n_out = n_out_host ? (int32_t *)(no ? no : h->h_nout) : nullptr;
    const size_t idx_bytes = E * (size_t)(h->d.nUE + h->d.nBS) * 4;
    if (obs_idx_host) {
        /* the sparse state goes through device staging + one copy: 4 (nUE + nBS) bytes per env */
        if (!h->h_idx) CU(h, cudaMalloc(&h->h_idx, idx_bytes));
        out.obs_idx = (int32_t *)h->h_idx;
    }
    rc = run_env(h, MODE_STEP, &in, &out, stream);
    if (rc) return rc;
    if (reward_host && !rw) CU(h, cudaMemcpyAsync(reward_host, h->h_reward, E * 8, cudaMemcpyDeviceToHost, st));
    if (done_host && !dn) CU(h, cudaMemcpyAsync(done_host, h->h_done, E, cudaMemcpyDeviceToHost, st));
    if (mean_sinr_host && !ms) CU(h, cudaMemcpyAsync(mean_sinr_host, h->h_mean, E * 8, cudaMemcpyDeviceToHost, st));
    if (n_out_host && !no) CU(h, cudaMemcpyAsync(n_out_host, h->h_nout, E * 4, cudaMemcpyDeviceToHost, st));
    if (obs_idx_host) CU(h, cudaMemcpyAsync(obs_idx_host, h->h_idx, idx_bytes, cudaMemcpyDeviceToHost, st));
    CU(h, cudaStreamSynchronize(st));
    return UAVENV_OK;
}

int uavenv_step_host(uavenv_t *h, const int64_t *action_host, void *obs_dev, double *reward_host, uint8_t *done_host,
                     double *mean_sinr_host, int32_t *n_out_host, void *stream) {
    return step_host_impl(h, action_host, obs_dev, reward_host, done_host, mean_sinr_host, n_out_host, nullptr, stream);
}

int uavenv_step_host_state(uavenv_t *h, const int64_t *action_host, void *obs_dev, double *reward_host, uint8_t *done_host,
                           double *mean_sinr_host, int32_t *n_out_host, int32_t *obs_idx_host, void *stream) {
    return step_host_impl(h, action_host, obs_dev, reward_host, done_host, mean_sinr_host, n_out_host, obs_idx_host, stream);
}

int uavenv_coverage_map(uavenv_t *h, const int16_t *bs_xy_dev, const double *fading_dev, void *out_dev, void *stream) {
    if (!h || !out_dev) return fail(h, UAVENV_EINVAL, "coverage_map: out_dev is NULL%s");
    int rc = use_device(h);
    if (rc) return rc;
    const int G = h->d.G;
    const dim3 grid((G * G + CTA_THREADS - 1) / CTA_THREADS, h->d.E);
    const int16_t *bs = bs_xy_dev ? bs_xy_dev : (const int16_t *)h->bs_xy;
    const uint32_t seq = (uint32_t)h->area_calls++;
    if (h->cfg.precision == UAVENV_PREC_FP64_PARITY)
        coverage_kernel<true><<<grid, CTA_THREADS, 0, (cudaStream_t)stream>>>(h->d, bs, fading_dev, seq, out_dev);
    else
        coverage_kernel<false><<<grid, CTA_THREADS, 0, (cudaStream_t)stream>>>(h->d, bs, fading_dev, seq, out_dev);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(h, UAVENV_ECUDA, "coverage_kernel launch: %s", cudaGetErrorString(e));
    h->launches++;
    return UAVENV_OK;
}

int64_t uavenv_state_bytes(const uavenv_t *h) {
    if (!h) return 0;
    int64_t t = 0;
    for (int i = 0; i < F_COUNT; i++) t += pad8(h->fields[i].bytes);
    return t;
}

int uavenv_state_field(const uavenv_t *h, int32_t field, int64_t *offset, int64_t *bytes) {
    if (!h || field < 0 || field >= F_COUNT) return UAVENV_EINVAL;
    int64_t t = 0;
    for (int i = 0; i < field; i++) t += pad8(h->fields[i].bytes);
    if (offset) *offset = t;
    if (bytes) *bytes = h->fields[field].bytes;
    return UAVENV_OK;
}

int uavenv_get_state(uavenv_t *h, void *host_buf, int64_t bytes) {
    if (!h || !host_buf || bytes < uavenv_state_bytes(h)) return fail(h, UAVENV_EINVAL, "get_state: buffer too small%s");
    int rc = use_device(h);
    if (rc) return rc;
    CU(h, cudaDeviceSynchronize());
    int64_t t = 0;
    for (int i = 0; i < F_COUNT; i++) {
        CU(h, cudaMemcpy((char *)host_buf + t, *h->fields[i].dev, (size_t)h->fields[i].bytes, cudaMemcpyDeviceToHost));
        t += pad8(h->fields[i].bytes);
    }
    return UAVENV_OK;
}

int uavenv_set_state(uavenv_t *h, const void *host_buf, int64_t bytes) {
    if (!h || !host_buf || bytes < uavenv_state_bytes(h)) return fail(h, UAVENV_EINVAL, "set_state: buffer too small%s");
    int rc = use_device(h);
    if (rc) return rc;
    CU(h, cudaDeviceSynchronize());
    int64_t t = 0;
    for (int i = 0; i < F_COUNT; i++) {
        CU(h, cudaMemcpy(*h->fields[i].dev, (const char *)host_buf + t, (size_t)h->fields[i].bytes, cudaMemcpyHostToDevice));
        t += pad8(h->fields[i].bytes);
    }
    return UAVENV_OK;
}

int uavenv_check(uavenv_t *h, uint32_t *flags_out, void *stream) {
    if (!h) return UAVENV_EINVAL;
    int rc = use_device(h);
    if (rc) return rc;
    uint32_t f = 0;
    cudaStream_t st = (cudaStream_t)stream;
    CU(h, cudaMemcpyAsync(&f, h->err_flags, 4, cudaMemcpyDeviceToHost, st));
    CU(h, cudaMemsetAsync(h->err_flags, 0, 4, st));
    CU(h, cudaStreamSynchronize(st));
    if (flags_out) *flags_out = f;
    if (f & 1u) return fail(h, UAVENV_EACTION, "an action was outside [0, n_act^n_bs) (or a digit >= n_act); those envs were not stepped%s");
    if (f & 8u) return fail(h, UAVENV_ECUDA, "bounds-check build: an observation index fell outside the env's observation%s");
    if (f & 2u) return fail(h, UAVENV_ETRACE, "trace exhausted (step_n past the end of the trace); those envs were not stepped%s");
    return UAVENV_OK;
}

int uavenv_guard_hits(uavenv_t *h, int64_t *hits_out, void *stream) {
    if (!h || !hits_out) return UAVENV_EINVAL;
    int rc = use_device(h);
    if (rc) return rc;
    unsigned long long v = 0;
    cudaStream_t st = (cudaStream_t)stream;
    CU(h, cudaMemcpyAsync(&v, (char *)h->err_flags + 8, 8, cudaMemcpyDeviceToHost, st));
    CU(h, cudaStreamSynchronize(st));
    *hits_out = (int64_t)v;
    return UAVENV_OK;
}

int uavenv_launch_plan(const uavenv_t *h, int32_t *grid, int32_t *threads, int32_t *tile_bytes, int32_t *ctas_per_sm) {
    if (!h) return UAVENV_EINVAL;
    if (grid) *grid = h->grid;
    if (threads) *threads = h->threads;
    if (tile_bytes) *tile_bytes = h->tile_bytes;
    if (ctas_per_sm) *ctas_per_sm = h->ctas_per_sm;
    return UAVENV_OK;
}

const uavenv_cfg *uavenv_get_cfg(const uavenv_t *h) { return h ? &h->cfg : nullptr; }
const char *uavenv_last_error(const uavenv_t *h) { return h ? h->err : "null handle"; }
int64_t uavenv_launch_count(const uavenv_t *h) { return h ? h->launches : 0; }

}  // extern "C"
