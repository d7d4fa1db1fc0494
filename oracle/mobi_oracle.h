/* TEST INFRASTRUCTURE -- the parity oracle, never the product.
 *
 * Plain-C, float64, single-environment restatement of the reference's
 * MobiEnvironment step path.  Every function cites the reference file:line it
 * restates (paths relative to /root/reference).  Only tests/, bench.py's
 * cpu_baseline / --impl reference leg and __graft_entry__.smoke() may link or
 * call this; nothing under drl_uav_cellularnet_b200/ does.
 *
 * Pinning: tests/golden/ *.npz were produced by the UNMODIFIED reference
 * (oracle/make_golden.py, through oracle/ref_loader.py) and
 * tests/test_oracle_golden.py replays them through this file.  The reference
 * itself ships no tests / golden vectors, so those
 * self-generated fixtures are the pin.
 */
#ifndef MOBI_ORACLE_H
#define MOBI_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_MAX_BS 32
#define ORC_HO_DEPTH 3

typedef struct {
    int32_t n_bs, n_ue, grid_n;
    int32_t n_groups;
    int32_t max_step;      /* MAXSTEP, mobile_env.py:18 */
    int32_t n_act;         /* N_ACT,   mobile_env.py:21 */
    int32_t bs_step;       /* BS_STEP, mobile_env.py:32 */
    int32_t min_bs_dist;   /* MIN_BS_DIST, mobile_env.py:28 (lock radius = min_bs_dist + bs_step, :157) */
    double grid_width;     /* channel.py:21 */
    double p_bs_dbm;       /* channel.py:36 */
    double noise_dbm;      /* channel.py:40 */
    double pl_a, pl_b, pl_dis; /* channel.py:46-48 */
    double ant_gain;       /* channel.py:50 */
    double eq_loss;        /* channel.py:52 */
    double shadow_mean, shadow_sd; /* channel.py:54-55 */
    double ho_thresh_db;   /* channel.py:82 */
    double out_thresh_db;  /* channel.py:7 */
    double v_min, v_max;   /* mobile_env.py:76 velocity=(0,1) */
    double aggregation;    /* mobile_env.py:76 */
    int32_t aggregating0, deaggregating0;      /* ue_mobility.py:450-451 */
    int32_t deaggregating_len, aggregating_len; /* ue_mobility.py:473,487 */
} orc_cfg;

void orc_cfg_default(orc_cfg *c, int n_bs, int n_ue, int grid_n, int n_groups);

/* numpy add.reduce order for a contiguous float64 vector (pairwise_sum). */
double orc_np_sum(const double *a, int64_t n);

/* ---- Philox4x32-10 (Salmon et al., SC'11); the repo's counter scheme ---- */
void orc_philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                    uint32_t k0, uint32_t k1, uint32_t out[4]);
void orc_philox_uniform2(uint64_t seed, uint32_t env, uint32_t idx, uint32_t seq,
                         uint32_t domain, double *a, double *b);
void orc_philox_normal4(uint64_t seed, uint32_t env, uint32_t idx, uint32_t seq, uint32_t domain, double z[4]);
/* the fading matrix (nUE,nBS) channel pass `epoch` of env `env_id` draws in Philox mode */
void orc_philox_fading(const orc_cfg *c, uint64_t seed, uint32_t env_id, uint32_t epoch, double *out);

/* ---- a4: reference_point_group, ue_mobility.py:409-523 ---- */
typedef struct orc_mob orc_mob;
orc_mob *orc_mob_create(const orc_cfg *c, const int32_t *group_sizes);
void orc_mob_destroy(orc_mob *m);
/* consume the init draws in reference order (ue_mobility.py:434-448):
 * x[N] y[N] theta[N] g_x[nG] g_y[nG] g_fl[nG] g_v[nG] g_theta[nG]; returns #consumed (3N+5nG) */
int64_t orc_mob_init(orc_mob *m, const double *uniforms);
/* one generator tick (ue_mobility.py:453-523); uniforms in reference order:
 * theta[N], then (only if k>0 groups arrived) theta[k] fl[k] v[k]; returns #consumed.
 * xy_out: float positions (N,2) as yielded (ue_mobility.py:523). */
int64_t orc_mob_tick(orc_mob *m, const double *uniforms, double *xy_out);
/* same tick drawing from Philox with the repo's counter scheme (tick index = seq) */
void orc_mob_init_philox(orc_mob *m, uint64_t seed, uint32_t env);
void orc_mob_tick_philox(orc_mob *m, uint64_t seed, uint32_t env, uint32_t tick, double *xy_out);
/* raw state access for tests: layout x[N] y[N] cos[N] sin[N] g_x g_y g_fl g_v g_cos g_sin [nG each], aggregating, deaggregating */
int64_t orc_mob_state_len(const orc_mob *m);
void orc_mob_get_state(const orc_mob *m, double *out);
void orc_mob_set_state(orc_mob *m, const double *in);

/* ---- a6/a7: Decimal_to_Base_N ue_mobility.py:310-336, BS_move ue_mobility.py:191-271 ---- */
int orc_action_digits(int64_t action, int base, int n_digits, int32_t *digits_out);
/* loc: (nBS,2) int64 x,y, updated in place; returns number of BS whose move was blocked by the lock test */
int orc_bs_move(const orc_cfg *c, int64_t *loc, const int32_t *digits);

/* ---- a8/a9: GetChannelGainAll channel.py:249-257, GetDLSinrAllDb channel.py:259-269 ---- */
void orc_sinr_all(const orc_cfg *c, const int64_t *ue_xy, const int64_t *bs_xy,
                  const double *fading /* (nUE,nBS) or NULL = 0 */, double *sinr_db /* (nUE,nBS) */);

/* ---- a10/a11: LTEChannel state machine, channel.py:92-93,110,113-116,138-176,216 ---- */
typedef struct {
    int32_t n_ue, n_bs;
    int32_t fifo_depth;               /* rows currently in bestBS_buf */
    int64_t *cur;                     /* current_BS (nUE) */
    double *cur_sinr;                 /* current_BS_sinr (nUE) */
    int64_t *fifo;                    /* bestBS_buf (3,nUE), row 0 oldest */
    uint8_t *out_prev;                /* membership of self.ue_out (nUE) */
} orc_chan;
orc_chan *orc_chan_create(int n_ue, int n_bs);
void orc_chan_destroy(orc_chan *ch);
/* channel.reset / ctor: best-server association from one SINR pass */
void orc_chan_reset(const orc_cfg *c, orc_chan *ch, const double *sinr_db);
/* UpdateDroneNet decisions from one SINR pass; outputs mean SINR, new-outage count, handover count */
void orc_chan_update(const orc_cfg *c, orc_chan *ch, const double *sinr_db,
                     double *mean_sinr, int32_t *n_out, int32_t *n_ho);

/* ---- a12: GetGridMap ue_mobility.py:173-188 + GetCurrentAssociationMap channel.py:387-409 ---- */
/* state: (nBS+1, G, G) float64, zeroed and refilled */
void orc_build_state(const orc_cfg *c, const int64_t *ue_xy, const int64_t *bs_xy,
                     const int64_t *cur, double *state);

/* ---- a1/a2/a3: the env shell mobile_env.py:37-233 ---- */
enum { ORC_MOB_GROUP = 0, ORC_MOB_TRACE = 1 };
enum { ORC_FADE_PHILOX = 0, ORC_FADE_INJECTED = 1, ORC_FADE_NONE = 2 };

typedef struct orc_env orc_env;
/* init_bs_xy (nBS,2) or NULL for the reference 4-BS layout (mobile_env.py:49-50) */
orc_env *orc_env_create(const orc_cfg *c, const int32_t *group_sizes, const int32_t *init_bs_xy,
                        int mobility, int fading, uint64_t seed, uint32_t env_id,
                        int warmup_ticks /* 200 = reference; <0 = caller loads mobility state */);
/* the LTEChannel constructor pass (mobile_env.py:100; channel.py:92-93,110); call after set_trace/set_state.
 * trace mode first takes ueLoc = trace[0] (mobile_env.py:87). */
int orc_env_ctor_channel(orc_env *e, const double *fading);
/* overwrite the UE cells from float positions (reference-state replay, mobile_env.py:94-97) */
void orc_env_set_ue_from_float(orc_env *e, const double *xy);
void orc_env_destroy(orc_env *e);
/* trace: (T, nUE, 2) int32, borrowed (caller keeps it alive) */
void orc_env_set_trace(orc_env *e, const int32_t *trace, int64_t T);
/* reset(): mobile_env.py:115-148.  fading: injected (nUE,nBS) or NULL.  mob_uniforms: injected draws for
 * the one mobility tick or NULL (Philox).  state_out may be NULL. */
int orc_env_reset(orc_env *e, const double *fading, const double *mob_uniforms, double *state_out);
typedef struct {
    double reward, mean_sinr;
    double r_dissect[2];
    int32_t n_out, n_ho, done, step_n, n_blocked;
} orc_step_out;
/* step()/step_test(): mobile_env.py:150-233 (identical arithmetic; UE source follows the env's mobility mode) */
int orc_env_step(orc_env *e, const int32_t *digits, const double *fading, const double *mob_uniforms,
                 double *state_out, orc_step_out *out);
/* live views for tests */
const int64_t *orc_env_ue_xy(const orc_env *e);
const int64_t *orc_env_bs_xy(const orc_env *e);
const orc_chan *orc_env_chan(const orc_env *e);
orc_mob *orc_env_mob(orc_env *e);
const double *orc_env_last_sinr(const orc_env *e); /* (nUE,nBS) of the last pass */
int32_t orc_env_n_clamped(const orc_env *e);
int32_t orc_env_step_n(const orc_env *e);
void orc_env_set_step_n(orc_env *e, int32_t s);

/* Bounded CPU throughput sample for bench.py: n_envs independent default envs, n_steps each, round-robin
 * on this thread; returns elapsed seconds.  Dense float64 state is rebuilt every step (as the reference does). */
double orc_bench_run(const orc_cfg *c, const int32_t *group_sizes, const int32_t *init_bs_xy,
                     int n_envs, int n_steps, uint64_t seed, uint32_t env_id0, double *checksum);

/* ---- GetSinrInArea, channel.py:411-433 (coverage map; see mobi_oracle.c) ---- */
void orc_sinr_in_area(const orc_cfg *c, const int64_t *bs_xy, const double *fading, const double *by_bs, double *out);
void orc_philox_area_fading(const orc_cfg *c, uint64_t seed, uint32_t env_id, uint32_t seq, double *by_bs);

/* config-5 sweep helpers (see mobi_oracle.c) */
int orc_replay_run(const orc_cfg *c, const int32_t *trace, int64_t T, uint64_t seed, uint32_t env_id,
                   const int64_t *actions, int n_steps, int32_t *n_out, int32_t *n_ho, double *reward,
                   int64_t *serving_hash);
void orc_make_trace(const orc_cfg *c, uint64_t seed, uint32_t env_id, int64_t T, int32_t *trace_out);
int orc_group_run(const orc_cfg *c, uint64_t seed, uint32_t env_id, const int64_t *actions, int n_steps,
                  int32_t *n_out, int32_t *n_ho, double *reward, int64_t *serving_hash, int64_t *cell_hash);

#ifdef __cplusplus
}
#endif
#endif
