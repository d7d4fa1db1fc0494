/* uavenv -- C-ABI of the B200-native batched MobiEnvironment step path.
 *
 * The reference has no plugin/FFI layer: its boundary is the Python class MobiEnvironment
 * (mobile_env.py:35-233) and the LTEChannel / ue_mobility functions it calls.  This header is the
 * boundary a maintainer binds instead (ctypes stub in INTEGRATION.md); every entry point cites the
 * reference interface it replaces (file:line relative to /root/reference).  Plain C types only:
 * no torch types, device pointers are raw `void*`/typed pointers, the stream is a cudaStream_t
 * passed as `void*`.
 *
 * Conventions
 *   - every entry point returns 0 on success or a negative UAVENV_E* code; it never aborts
 *     (reference: sys.exit / assert / IndexError, mobile_env.py:84,91,203).  uavenv_last_error()
 *     returns the message of the last failure on that handle.
 *   - a handle is single-owner and not thread-safe (reference: one env per worker thread, main.py:173).
 *   - all work is enqueued on the caller's stream; no entry point except create/destroy/get/set_state,
 *     uavenv_step_host and uavenv_check synchronises.
 *   - E environments are stepped per call.  Environment e of this handle is GLOBAL environment
 *     cfg.env_offset + e: every random draw is keyed by (seed, global env id, sequence number, lane),
 *     so results do not depend on how environments are sharded over handles / GPUs.
 */
#ifndef UAVENV_H
#define UAVENV_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define UAVENV_MAX_BS 32      /* one warp lane per BS */
#define UAVENV_MAX_GROUPS 32
#define UAVENV_HO_DEPTH 3     /* hoBufDepth, channel.py:81 */

enum { UAVENV_OK = 0, UAVENV_EINVAL = -1, UAVENV_ECUDA = -2, UAVENV_ETRACE = -3, UAVENV_ENOMEM = -4,
       UAVENV_EACTION = -5 };

enum { UAVENV_MOB_GROUP = 0,      /* reference_point_group, mobile_env.py:74-79 */
       UAVENV_MOB_TRACE = 1 };    /* "read_trace", mobile_env.py:82-88,202-203 */
enum { UAVENV_FADE_PHILOX = 0,    /* N(mean,sd) per (UE,BS) pair per pass from Philox4x32-10 (replaces np.random.normal, channel.py:240) */
       UAVENV_FADE_INJECTED = 1,  /* caller supplies float64 fading[E,nUE,nBS] for every pass (parity replay) */
       UAVENV_FADE_NONE = 2 };
enum { UAVENV_PREC_FP32_FAST = 0, /* fp32 + MUFU intrinsics, log-domain SINR */
       UAVENV_PREC_FP64_PARITY = 1, /* float64, reference operation order; decision-exact replay */
       UAVENV_PREC_FP32_GUARDED = 2 }; /* the fp32 pass; a UE whose row is within cfg.guard_db of a decision boundary (top-2
                                          gap of the argmax, handover threshold, outage threshold: channel.py:141,156-159,170)
                                          is re-evaluated in float64 inside the same kernel, so serving BS / handover /
                                          outage decisions equal FP64_PARITY's bit for bit; SINR / reward stay fp32-accurate
                                          (1e-3 dB / 1e-5 relative, BASELINE.json north_star) */
enum { UAVENV_OBS_NONE = 0,
       UAVENV_OBS_F32 = 1,        /* float32 [E, nBS+1, G, G], fully rewritten every step (mobile_env.py:107,169-170,194) */
       UAVENV_OBS_F32_INCREMENTAL = 3 }; /* float32; the caller keeps the SAME buffer between calls and never writes it:
                                            reset rewrites it in full, step only touches the cells that changed */

/* reference: module constants mobile_env.py:17-32, channel.py:21-82, ue_mobility.py:436,450-451,473,487 */
typedef struct uavenv_cfg {
    int32_t n_envs, n_bs, n_ue, grid_n;
    int32_t mobility, fading, precision, obs_mode;
    uint64_t seed;
    int64_t env_offset;          /* global id of env 0 (sharding) */
    int32_t device;              /* CUDA device ordinal */
    int32_t max_step;            /* MAXSTEP 2000 */
    int32_t n_act;               /* N_ACT 5 (up to 9: actions 5-8 are the long moves of BS_move, ue_mobility.py:239-253) */
    int32_t bs_step;             /* BS_STEP 2 */
    int32_t min_bs_dist;         /* MIN_BS_DIST 2; lock radius = min_bs_dist + bs_step (mobile_env.py:157) */
    int32_t warmup_ticks;        /* 200 (mobile_env.py:77-79); one more tick gives the initial positions (:93-97) */
    int32_t n_groups;            /* 4 */
    int32_t group_sizes[UAVENV_MAX_GROUPS]; /* [10,10,10,10] (mobile_env.py:76) */
    int32_t has_init_bs;         /* 0: reference layout for nBS=4 (mobile_env.py:49-50), lattice otherwise */
    int32_t init_bs_xy[UAVENV_MAX_BS * 2];
    int32_t aggregating0, deaggregating0, deaggregating_len, aggregating_len; /* 200,100,100,10 */
    double grid_width;           /* 5 m per cell */
    double p_bs_dbm, noise_dbm;  /* 20, -121 */
    double pl_a, pl_b, pl_dis;   /* 38, 30, 0 */
    double ant_gain, eq_loss;    /* 2, 0 */
    double shadow_mean, shadow_sd; /* 0, 2 */
    double ho_thresh_db;         /* 1 */
    double out_thresh_db;        /* OUT_THRESH 0 */
    double v_min, v_max;         /* group velocity (0,1) */
    double aggregation;          /* 0.8 */
    double guard_db;             /* FP32_GUARDED: width of the re-evaluation band in dB.  Default 1e-3: a decision compares two
                                    SINRs, so the band must exceed twice the largest |fp32 - float64| SINR difference; measured
                                    2.6e-5 dB at the reference sizes, 1.0e-4 dB at 32 BS x 2048 UE (profiles/r2/parity_stats.json) */
} uavenv_cfg;

/* Optional per-call inputs; device pointers; NULL = not supplied. */
typedef struct uavenv_in {
    const int64_t *action;       /* [E] joint action, MSB-first base-n_act digits (Decimal_to_Base_N, ue_mobility.py:310-336) */
    const uint8_t *digits;       /* [E,nBS] per-BS digits; takes precedence (needed when n_act^nBS overflows int64, mobile_env.py:104) */
    const double *fading;        /* [E,nUE,nBS] float64, UAVENV_FADE_INJECTED only */
    const double *mob_uniforms;  /* [E, nUE+3*nG] float64: the uniforms one reference generator tick would draw, in its order
                                    (theta[nUE], then theta[k] fl[k] v[k] for the k arrived groups; ue_mobility.py:508-521) */
    const uint8_t *env_mask;     /* [E] reset only: 1 = reset this env, 0 = leave untouched; NULL = all */
} uavenv_in;

/* Outputs; device pointers; any may be NULL.  Valid until the next call on the handle. */
typedef struct uavenv_out {
    void *obs;                   /* per cfg.obs_mode: [E,nBS+1,G,G], plane 0 = BS counts, plane 1+b = UEs served by b, [plane,x,y] */
    double *reward;              /* [E] max(meanSINR/20 - nOut/nUE, -1)            (mobile_env.py:163-167,189) */
    double *mean_sinr;           /* [E] mean serving SINR in dB                    (channel.py:216) */
    int32_t *n_out;              /* [E] NEW outages this step                      (channel.py:170-174) */
    int32_t *n_ho;               /* [E] handovers applied this step                (channel.py:162-167) */
    int32_t *n_blocked;          /* [E] BS moves blocked by the lock test          ("COLLIDED", ue_mobility.py:267-268) */
    uint8_t *done;               /* [E] step_n >= max_step                         (mobile_env.py:186-187) */
    int32_t *step_n;             /* [E]                                            (mobile_env.py:179) */
    uint8_t *serving;            /* [E,nUE] current_BS after handover              (channel.py:167) */
    void *serving_sinr;          /* [E,nUE] current_BS_sinr (pre-handover cell): float32 (fast) / float64 (parity) (channel.py:145-146) */
    void *sinr_all;              /* [E,nUE,nBS] full SINR matrix of this pass, float32 / float64 (diagnostic; channel.py:140) */
    float *fading_used;          /* [E,nUE,nBS] the fading the pass applied, dB (diagnostic: lets the oracle audit fast mode) */
    int16_t *ue_xy;              /* [E,nUE,2] UE cells                             (mobile_env.py:155) */
    int16_t *bs_xy;              /* [E,nBS,2] BS cells after the move              (mobile_env.py:157) */
    uint8_t *bs_digits;          /* [E,nBS] decoded per-BS actions (act_all)       (mobile_env.py:209) */
    int32_t *obs_idx;            /* [E,nUE+nBS] the observation in sparse form: flat index (plane*G + x)*G + y of every count of
                                    `obs` (UE u -> plane 1+serving[u]; BS b -> plane 0), duplicates = counts > 1.  What the
                                    policy network's first layer consumes (main.py:147-153 on ~44 non-zeros of 50 000) */
} uavenv_out;

typedef struct uavenv uavenv_t;

/* Fill cfg with the reference defaults for the given sizes (groups: n_ue split evenly over 4 groups for the
 * reference sizes, else over min(32, n_bs) groups). */
int uavenv_cfg_default(uavenv_cfg *cfg, int32_t n_envs, int32_t n_bs, int32_t n_ue, int32_t grid_n);

/* MobiEnvironment.__init__ (mobile_env.py:37-108) for E envs: BS layout, mobility init + warm-up ticks, and the
 * LTEChannel constructor's best-server pass (channel.py:92-93,110).  In trace / injected modes the constructor
 * pass is deferred: call uavenv_set_trace and then uavenv_ctor_pass. */
int uavenv_create(const uavenv_cfg *cfg, uavenv_t **out);
void uavenv_destroy(uavenv_t *h);

/* np.load(test_mobi_file_name) (mobile_env.py:85): xy is HOST int32 [T, (per_env ? E : 1), nUE, 2]; copied to the device. */
int uavenv_set_trace(uavenv_t *h, const int32_t *xy_host, int64_t T, int32_t per_env);

/* The LTEChannel constructor pass on its own (channel.py:92-93,110), for trace / injected-fading handles. */
int uavenv_ctor_pass(uavenv_t *h, const uavenv_in *in, const uavenv_out *out, void *stream);

/* MobiEnvironment.reset (mobile_env.py:115-148): BS back to the initial layout, one mobility tick (or trace[0]),
 * channel.reset (channel.py:113-116), step_n = 0. */
int uavenv_reset(uavenv_t *h, const uavenv_in *in, const uavenv_out *out, void *stream);

/* MobiEnvironment.step / step_test (mobile_env.py:150-194 / 196-233). */
int uavenv_step(uavenv_t *h, const uavenv_in *in, const uavenv_out *out, void *stream);

/* Host-buffer convenience: copies action_host [E] (pinned or pageable) to the device, steps, copies reward/done/
 * mean_sinr/n_out (any may be NULL) back to HOST buffers and synchronises the stream.  obs_dev is a DEVICE pointer
 * (or NULL) as in uavenv_out.obs. */
int uavenv_step_host(uavenv_t *h, const int64_t *action_host, void *obs_dev, double *reward_host, uint8_t *done_host,
                     double *mean_sinr_host, int32_t *n_out_host, void *stream);

/* The same, and the step's STATE is returned to the host too, in sparse form: obs_idx_host int32 [E, nUE + nBS] (see
 * uavenv_out.obs_idx) -- what a host-side policy consumes of gym's `state` (mobile_env.py:194) without moving the dense
 * 4 (nBS+1) G^2 bytes per env over PCIe. */
int uavenv_step_host_state(uavenv_t *h, const int64_t *action_host, void *obs_dev, double *reward_host, uint8_t *done_host,
                           double *mean_sinr_host, int32_t *n_out_host, int32_t *obs_idx_host, void *stream);

/* LTEChannel.GetSinrInArea (channel.py:411-433; main_test.py:89): per env the coverage map out[e, x, y] = downlink SINR (dB)
 * of cell (x, y) from its nearest BS, row / column 0 zero; out is float32 (fast) / float64 (parity) [E, G, G].
 * bs_xy_dev: int16 [E,nBS,2] BS cells, NULL = the envs' current BS cells.  fading_dev: float64 [E, (G-1)^2, nBS] draws in
 * the reference's call order (per cell: interferers in ascending index order, then the serving BS), NULL = Philox draws
 * keyed by (seed, env, cell, BS, number of this call) or none, per cfg.fading. */
int uavenv_coverage_map(uavenv_t *h, const int16_t *bs_xy_dev, const double *fading_dev, void *out_dev, void *stream);

/* State blob (copy.deepcopy(env), gradient.py:15; checkpoint).  Host buffer; layout in DESIGN.md. Synchronises. */
int64_t uavenv_state_bytes(const uavenv_t *h);
/* byte offset / size inside the blob of field 0..6: xy f64[E,nUE,2] (float UE positions), theta_u f64[E,nUE],
 * group f64[E,6,nG] (g_x g_y g_fl g_v g_cos g_sin), counters i32[E,8] (tick, epoch, step_n, aggregating,
 * deaggregating), bs_xy i16[E,nBS,2], ue_cell i16[E,nUE,2], ho_word u32[E,nUE] */
int uavenv_state_field(const uavenv_t *h, int32_t field, int64_t *offset, int64_t *bytes);
int uavenv_get_state(uavenv_t *h, void *host_buf, int64_t bytes);
int uavenv_set_state(uavenv_t *h, const void *host_buf, int64_t bytes);

/* Sticky device-side error flags since the last check (bit0 action out of range, bit1 trace exhausted,
 * bit2 UE cell clamped from G to G-1).  Synchronises the stream. */
int uavenv_check(uavenv_t *h, uint32_t *flags_out, void *stream);

/* FP32_GUARDED diagnostic: UEs re-evaluated in float64 since the handle was created.  Synchronises the stream. */
int uavenv_guard_hits(uavenv_t *h, int64_t *hits_out, void *stream);

/* Diagnostic: the launch plan of the step kernel -- CTAs (one per env), threads per CTA, bytes of the zeroed
 * shared-memory tile the TMA warp streams the observation from (0: plain stores + atomics fallback for odd sizes)
 * and resident CTAs per SM. */
int uavenv_launch_plan(const uavenv_t *h, int32_t *grid, int32_t *threads, int32_t *tile_bytes, int32_t *ctas_per_sm);

const uavenv_cfg *uavenv_get_cfg(const uavenv_t *h);
const char *uavenv_last_error(const uavenv_t *h);
/* kernels launched by this handle so far (bench.py's gpu_launches) */
int64_t uavenv_launch_count(const uavenv_t *h);
const char *uavenv_version(void);

#ifdef __cplusplus
}
#endif
#endif /* UAVENV_H */
