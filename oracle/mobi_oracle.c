/* TEST INFRASTRUCTURE -- the parity oracle, never the product.  See mobi_oracle.h.
 *
 * Float64 restatement of the reference's MobiEnvironment step path, one env at a
 * time, written to be bit-faithful to the reference's arithmetic (same operation
 * order, numpy's pairwise summation, libm log10/pow/sqrt as CPython calls them).
 * Compile with -ffp-contract=off so no FMA contraction changes the rounding.
 *
 * Citations are file:line into /root/reference.
 */
#include "mobi_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

/* ------------------------------------------------------------------------- */
void orc_cfg_default(orc_cfg *c, int n_bs, int n_ue, int grid_n, int n_groups) {
    memset(c, 0, sizeof(*c));
    c->n_bs = n_bs;
    c->n_ue = n_ue;
    c->grid_n = grid_n;
    c->n_groups = n_groups;
    c->max_step = 2000;     /* mobile_env.py:18 */
    c->n_act = 5;           /* mobile_env.py:21 */
    c->bs_step = 2;         /* mobile_env.py:32 */
    c->min_bs_dist = 2;     /* mobile_env.py:28 */
    c->grid_width = 5;      /* channel.py:21 */
    c->p_bs_dbm = 20;       /* channel.py:36 */
    c->noise_dbm = -121;    /* channel.py:40 */
    c->pl_a = 38;           /* channel.py:46 */
    c->pl_b = 30;           /* channel.py:47 */
    c->pl_dis = 0;          /* channel.py:48 */
    c->ant_gain = 2;        /* channel.py:50 */
    c->eq_loss = 0;         /* channel.py:52 */
    c->shadow_mean = 0;     /* channel.py:54 */
    c->shadow_sd = 2;       /* channel.py:55 */
    c->ho_thresh_db = 1;    /* channel.py:82 */
    c->out_thresh_db = 0;   /* channel.py:7 */
    c->v_min = 0;           /* mobile_env.py:76 */
    c->v_max = 1;
    c->aggregation = 0.8;   /* mobile_env.py:76 */
    c->aggregating0 = 200;  /* ue_mobility.py:450 */
    c->deaggregating0 = 100;/* ue_mobility.py:451 */
    c->deaggregating_len = 100; /* ue_mobility.py:473 */
    c->aggregating_len = 10;    /* ue_mobility.py:487 */
}

/* numpy's DOUBLE_pairwise_sum (the loop np.sum / np.mean run on a contiguous
 * float64 vector): n<8 sequential, n<=128 eight interleaved accumulators,
 * otherwise split in halves (left half rounded down to a multiple of 8).
 * Used where the reference calls np.sum (channel.py:265) and np.mean (channel.py:216). */
double orc_np_sum(const double *a, int64_t n) {
    if (n < 8) {
        double res = 0.0;
        for (int64_t i = 0; i < n; i++) res += a[i];
        return res;
    } else if (n <= 128) {
        double r[8];
        int64_t i;
        for (int j = 0; j < 8; j++) r[j] = a[j];
        for (i = 8; i < n - (n % 8); i += 8)
            for (int j = 0; j < 8; j++) r[j] += a[i + j];
        double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; i++) res += a[i];
        return res;
    } else {
        int64_t n2 = n / 2;
        n2 -= n2 % 8;
        return orc_np_sum(a, n2) + orc_np_sum(a + n2, n - n2);
    }
}

/* ------------------------------------------------------------------------- */
/* Philox4x32-10.  Not part of the reference (it uses numpy's global MT19937,
 * channel.py:240, ue_mobility.py:6,408); it is how BOTH this oracle and the CUDA
 * path produce "the same inputs" in synthetic mode.  Counter scheme: DESIGN.md. */
static inline void philox_round(uint32_t c[4], uint32_t k0, uint32_t k1) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
    uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
    uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
    uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}

void orc_philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                    uint32_t k0, uint32_t k1, uint32_t out[4]) {
    uint32_t c[4] = {c0, c1, c2, c3};
    for (int r = 0; r < 10; r++) {
        philox_round(c, k0, k1);
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    memcpy(out, c, sizeof(c));
}

static inline double u53(uint32_t hi, uint32_t lo) {
    return ((double)(hi >> 5) * 67108864.0 + (double)(lo >> 6)) * (1.0 / 9007199254740992.0);
}

void orc_philox_uniform2(uint64_t seed, uint32_t env, uint32_t idx, uint32_t seq,
                         uint32_t domain, double *a, double *b) {
    uint32_t w[4];
    orc_philox4x32(env, idx, seq, domain, (uint32_t)seed, (uint32_t)(seed >> 32), w);
    *a = u53(w[0], w[1]);
    *b = u53(w[2], w[3]);
}

/* four N(0,1) from one Philox call: (w0,w1) and (w2,w3) each feed one Box-Muller pair,
 * u = (w + 0.5) * 2^-32 in (0,1) */
void orc_philox_normal4(uint64_t seed, uint32_t env, uint32_t idx, uint32_t seq, uint32_t domain, double z[4]) {
    uint32_t w[4];
    orc_philox4x32(env, idx, seq, domain, (uint32_t)seed, (uint32_t)(seed >> 32), w);
    for (int k = 0; k < 2; k++) {
        double u1 = ((double)w[2 * k] + 0.5) * (1.0 / 4294967296.0);
        double u2 = ((double)w[2 * k + 1] + 0.5) * (1.0 / 4294967296.0);
        double r = sqrt(-2.0 * log(u1));
        z[2 * k] = r * cos(6.283185307179586 * u2);
        z[2 * k + 1] = r * sin(6.283185307179586 * u2);
    }
}

enum {
    DOM_INIT_XY = 1, DOM_INIT_TH = 2, DOM_INIT_GXY = 3, DOM_INIT_GFV = 4, DOM_INIT_GTH = 5,
    DOM_THETA = 6, DOM_GRP_TF = 7, DOM_GRP_V = 8, DOM_FADING = 9, DOM_ACTION = 10
};

#define TWO_PI 6.283185307179586 /* 2*np.pi */

/* ------------------------------------------------------------------------- */
/* a4: reference_point_group generator state (ue_mobility.py:415-451) */
struct orc_mob {
    int n, ng;
    int *g_ref;            /* ue_mobility.py:423-426 */
    int *g_start, *g_size; /* groups, ue_mobility.py:417-421 */
    double *x, *y, *cost, *sint;
    double *g_x, *g_y, *g_fl, *g_v, *g_cos, *g_sin;
    int aggregating, deaggregating;
    double max_x, max_y, fl_max, v_min, v_max, aggr;
    int deagg_len, agg_len;
    unsigned char *flag;   /* scratch: per-group "has a reflecting member" */
};

orc_mob *orc_mob_create(const orc_cfg *c, const int32_t *group_sizes) {
    orc_mob *m = (orc_mob *)calloc(1, sizeof(orc_mob));
    m->n = c->n_ue;
    m->ng = c->n_groups;
    m->g_ref = (int *)calloc(m->n, sizeof(int));
    m->g_start = (int *)calloc(m->ng, sizeof(int));
    m->g_size = (int *)calloc(m->ng, sizeof(int));
    int prev = 0;
    for (int g = 0; g < m->ng; g++) {
        m->g_start[g] = prev;
        m->g_size[g] = group_sizes[g];
        for (int k = 0; k < group_sizes[g] && prev + k < m->n; k++) m->g_ref[prev + k] = g;
        prev += group_sizes[g];
    }
    m->x = (double *)calloc(4 * (size_t)m->n, sizeof(double));
    m->y = m->x + m->n;
    m->cost = m->y + m->n;
    m->sint = m->cost + m->n;
    m->g_x = (double *)calloc(6 * (size_t)m->ng, sizeof(double));
    m->g_y = m->g_x + m->ng;
    m->g_fl = m->g_y + m->ng;
    m->g_v = m->g_fl + m->ng;
    m->g_cos = m->g_v + m->ng;
    m->g_sin = m->g_cos + m->ng;
    m->flag = (unsigned char *)calloc(m->ng, 1);
    m->aggregating = c->aggregating0;
    m->deaggregating = c->deaggregating0;
    m->max_x = c->grid_n;  /* dimensions=(grid_n, grid_n), mobile_env.py:76 */
    m->max_y = c->grid_n;
    m->fl_max = c->grid_n; /* FL_MAX = max(dimensions), ue_mobility.py:428 */
    m->v_min = c->v_min;
    m->v_max = c->v_max;
    m->aggr = c->aggregation;
    m->deagg_len = c->deaggregating_len;
    m->agg_len = c->aggregating_len;
    return m;
}

void orc_mob_destroy(orc_mob *m) {
    if (!m) return;
    free(m->g_ref); free(m->g_start); free(m->g_size);
    free(m->x); free(m->g_x); free(m->flag);
    free(m);
}

/* U(MIN,MAX,.) = rand*(MAX-MIN)+MIN, ue_mobility.py:408 */
static inline double U(double lo, double hi, double r) { return r * (hi - lo) + lo; }

int64_t orc_mob_init(orc_mob *m, const double *u) {
    int n = m->n, ng = m->ng;
    int64_t k = 0;
    for (int i = 0; i < n; i++) m->x[i] = U(0, m->max_x, u[k++]);           /* :434 */
    for (int i = 0; i < n; i++) m->y[i] = U(0, m->max_y, u[k++]);           /* :435 */
    for (int i = 0; i < n; i++) {                                            /* :437-439 */
        double th = U(0, TWO_PI, u[k++]);
        m->cost[i] = cos(th);
        m->sint[i] = sin(th);
    }
    for (int g = 0; g < ng; g++) m->g_x[g] = U(0, m->max_x, u[k++]);        /* :442 */
    for (int g = 0; g < ng; g++) m->g_y[g] = U(0, m->max_x, u[k++]);        /* :443 (MAX_X, sic) */
    for (int g = 0; g < ng; g++) m->g_fl[g] = U(0, m->fl_max, u[k++]);      /* :444 */
    for (int g = 0; g < ng; g++) m->g_v[g] = U(m->v_min, m->v_max, u[k++]); /* :445 */
    for (int g = 0; g < ng; g++) {                                           /* :446-448 */
        double th = U(0, TWO_PI, u[k++]);
        m->g_cos[g] = cos(th);
        m->g_sin[g] = sin(th);
    }
    return k;
}

/* everything of one tick that happens before the random redraws (ue_mobility.py:455-505) */
static void mob_advance(orc_mob *m) {
    int n = m->n, ng = m->ng;
    const size_t ngz = ng > 0 ? (size_t)ng : 0;
    for (int i = 0; i < n; i++) {           /* :455-456, velocity = 1. (:436) */
        m->x[i] = m->x[i] + 1.0 * m->cost[i];
        m->y[i] = m->y[i] + 1.0 * m->sint[i];
    }
    for (int g = 0; g < ng; g++) {          /* :458-459 */
        m->g_x[g] = m->g_x[g] + m->g_v[g] * m->g_cos[g];
        m->g_y[g] = m->g_y[g] + m->g_v[g] * m->g_sin[g];
    }
    if (m->aggregating) {                   /* :461-473 */
        for (int g = 0; g < ng; g++)
            for (int i = m->g_start[g]; i < m->g_start[g] + m->g_size[g] && i < n; i++) {
                double xg = m->x[i], yg = m->y[i];
                double c_theta = atan2(m->g_y[g] - yg, m->g_x[g] - xg);
                m->x[i] = xg + m->g_v[g] * m->g_cos[g] + m->aggr * cos(c_theta);
                m->y[i] = yg + m->g_v[g] * m->g_sin[g] + m->aggr * sin(c_theta);
            }
        m->aggregating -= 1;
        if (m->aggregating == 0) m->deaggregating = m->deagg_len;
    } else {                                /* :475-487 */
        for (int g = 0; g < ng; g++)
            for (int i = m->g_start[g]; i < m->g_start[g] + m->g_size[g] && i < n; i++) {
                m->x[i] = m->x[i] + m->g_v[g] * m->g_cos[g];
                m->y[i] = m->y[i] + m->g_v[g] * m->g_sin[g];
            }
        m->deaggregating -= 1;
        if (m->deaggregating == 0) m->aggregating = m->agg_len;
    }
    /* reflections, in the reference's order; each group with >=1 reflecting member flips once (:490-505) */
    memset(m->flag, 0, ngz);
    for (int i = 0; i < n; i++) if (m->x[i] < 0) { m->x[i] = -m->x[i]; m->flag[m->g_ref[i]] = 1; }
    for (int g = 0; g < ng; g++) if (m->flag[g]) m->g_cos[g] = -m->g_cos[g];
    memset(m->flag, 0, ngz);
    for (int i = 0; i < n; i++) if (m->x[i] > m->max_x) { m->x[i] = 2 * m->max_x - m->x[i]; m->flag[m->g_ref[i]] = 1; }
    for (int g = 0; g < ng; g++) if (m->flag[g]) m->g_cos[g] = -m->g_cos[g];
    memset(m->flag, 0, ngz);
    for (int i = 0; i < n; i++) if (m->y[i] < 0) { m->y[i] = -m->y[i]; m->flag[m->g_ref[i]] = 1; }
    for (int g = 0; g < ng; g++) if (m->flag[g]) m->g_sin[g] = -m->g_sin[g];
    memset(m->flag, 0, ngz);
    for (int i = 0; i < n; i++) if (m->y[i] > m->max_y) { m->y[i] = 2 * m->max_y - m->y[i]; m->flag[m->g_ref[i]] = 1; }
    for (int g = 0; g < ng; g++) if (m->flag[g]) m->g_sin[g] = -m->g_sin[g];
}

static void mob_yield(const orc_mob *m, double *xy_out) {
    if (!xy_out) return;
    for (int i = 0; i < m->n; i++) { xy_out[2 * i] = m->x[i]; xy_out[2 * i + 1] = m->y[i]; } /* :523 */
}

int64_t orc_mob_tick(orc_mob *m, const double *u, double *xy_out) {
    int n = m->n, ng = m->ng;
    int64_t k = 0;
    mob_advance(m);
    for (int i = 0; i < n; i++) {           /* :508-510 */
        double th = U(0, TWO_PI, u[k++]);
        m->cost[i] = cos(th);
        m->sint[i] = sin(th);
    }
    int arrived[256], na = 0;               /* :513-521 */
    for (int g = 0; g < ng; g++) {
        m->g_fl[g] = m->g_fl[g] - m->g_v[g];
        if (m->g_v[g] > 0. && m->g_fl[g] <= 0. && na < 256) arrived[na++] = g;
    }
    if (na > 0) {
        for (int j = 0; j < na; j++) {
            double th = U(0, TWO_PI, u[k + j]);
            m->g_cos[arrived[j]] = cos(th);
            m->g_sin[arrived[j]] = sin(th);
        }
        k += na;
        for (int j = 0; j < na; j++) m->g_fl[arrived[j]] = U(0, m->fl_max, u[k + j]);
        k += na;
        for (int j = 0; j < na; j++) m->g_v[arrived[j]] = U(m->v_min, m->v_max, u[k + j]);
        k += na;
    }
    mob_yield(m, xy_out);
    return k;
}

void orc_mob_init_philox(orc_mob *m, uint64_t seed, uint32_t env) {
    double a, b;
    for (int i = 0; i < m->n; i++) {
        orc_philox_uniform2(seed, env, (uint32_t)i, 0, DOM_INIT_XY, &a, &b);
        m->x[i] = U(0, m->max_x, a);
        m->y[i] = U(0, m->max_y, b);
        orc_philox_uniform2(seed, env, (uint32_t)i, 0, DOM_INIT_TH, &a, &b);
        double th = U(0, TWO_PI, a);
        m->cost[i] = cos(th);
        m->sint[i] = sin(th);
    }
    for (int g = 0; g < m->ng; g++) {
        orc_philox_uniform2(seed, env, (uint32_t)g, 0, DOM_INIT_GXY, &a, &b);
        m->g_x[g] = U(0, m->max_x, a);
        m->g_y[g] = U(0, m->max_x, b);
        orc_philox_uniform2(seed, env, (uint32_t)g, 0, DOM_INIT_GFV, &a, &b);
        m->g_fl[g] = U(0, m->fl_max, a);
        m->g_v[g] = U(m->v_min, m->v_max, b);
        orc_philox_uniform2(seed, env, (uint32_t)g, 0, DOM_INIT_GTH, &a, &b);
        double th = U(0, TWO_PI, a);
        m->g_cos[g] = cos(th);
        m->g_sin[g] = sin(th);
    }
}

void orc_mob_tick_philox(orc_mob *m, uint64_t seed, uint32_t env, uint32_t tick, double *xy_out) {
    double a, b;
    mob_advance(m);
    for (int i = 0; i < m->n; i++) {
        orc_philox_uniform2(seed, env, (uint32_t)i, tick, DOM_THETA, &a, &b);
        double th = U(0, TWO_PI, a);
        m->cost[i] = cos(th);
        m->sint[i] = sin(th);
    }
    for (int g = 0; g < m->ng; g++) {
        m->g_fl[g] = m->g_fl[g] - m->g_v[g];
        if (m->g_v[g] > 0. && m->g_fl[g] <= 0.) {
            orc_philox_uniform2(seed, env, (uint32_t)g, tick, DOM_GRP_TF, &a, &b);
            double th = U(0, TWO_PI, a);
            m->g_cos[g] = cos(th);
            m->g_sin[g] = sin(th);
            m->g_fl[g] = U(0, m->fl_max, b);
            orc_philox_uniform2(seed, env, (uint32_t)g, tick, DOM_GRP_V, &a, &b);
            m->g_v[g] = U(m->v_min, m->v_max, a);
        }
    }
    mob_yield(m, xy_out);
}

int64_t orc_mob_state_len(const orc_mob *m) { return 4 * (int64_t)m->n + 6 * (int64_t)m->ng + 2; }

void orc_mob_get_state(const orc_mob *m, double *out) {
    memcpy(out, m->x, 4 * (size_t)m->n * sizeof(double));
    memcpy(out + 4 * m->n, m->g_x, 6 * (size_t)m->ng * sizeof(double));
    out[4 * m->n + 6 * m->ng] = m->aggregating;
    out[4 * m->n + 6 * m->ng + 1] = m->deaggregating;
}

void orc_mob_set_state(orc_mob *m, const double *in) {
    memcpy(m->x, in, 4 * (size_t)m->n * sizeof(double));
    memcpy(m->g_x, in + 4 * m->n, 6 * (size_t)m->ng * sizeof(double));
    m->aggregating = (int)in[4 * m->n + 6 * m->ng];
    m->deaggregating = (int)in[4 * m->n + 6 * m->ng + 1];
}

/* ------------------------------------------------------------------------- */
/* a6: Decimal_to_Base_N, ue_mobility.py:310-336 -- most significant digit first (digit 0 <-> BS 0) */
int orc_action_digits(int64_t action, int base, int n_digits, int32_t *digits) {
    if (!(1 < base && base < 37)) return -1;      /* :323-324 */
    if (action < 0) return -2;
    for (int i = 0; i < n_digits; i++) digits[i] = 0;
    int64_t cur = action;
    int pos = n_digits - 1;
    while (cur) {                                  /* :327-330 */
        if (pos < 0) return -3;                    /* more digits than BSs: the reference fails at :334 */
        digits[pos--] = (int32_t)(cur % base);
        cur = cur / base;
    }
    return 0;
}

/* a7: BS_move, ue_mobility.py:191-271.  Sequential, in place; the lock test (:256-263)
 * compares BS i's PRE-move cell with the others' current cells (already moved for j<i). */
int orc_bs_move(const orc_cfg *c, int64_t *loc, const int32_t *digits) {
    const int64_t xMin = 1, xMax = c->grid_n, yMin = 1, yMax = c->grid_n; /* mobile_env.py:45 */
    const int64_t s = c->bs_step, sl = 2 * (int64_t)c->bs_step;           /* :211 */
    const int64_t lock = c->min_bs_dist + c->bs_step;                      /* mobile_env.py:157 */
    int blocked = 0;
    for (int i = 0; i < c->n_bs; i++) {
        int64_t x = loc[2 * i], y = loc[2 * i + 1];
        switch (digits[i]) {                       /* :221-253 */
        case 0: if (x + s < xMax) x = x + s; break;
        case 1: if (x - s > xMin) x = x - s; break;
        case 2: if (y + s < yMax) y = y + s; break;
        case 3: if (y - s > yMin) y = y - s; break;
        case 5: if (x + sl < xMax) x = x + sl; break;
        case 6: if (x - sl > xMin) x = x - sl; break;
        case 7: if (y + sl < yMax) y = y + sl; break;
        case 8: if (y - sl > yMin) y = y - sl; break;
        default: break;                            /* 4 = stay */
        }
        int collision = 0;
        for (int j = 0; j < c->n_bs; j++) {        /* :258-263 ; z equal for all BS (mobile_env.py:58) */
            if (i == j) continue;
            int64_t dx = loc[2 * i] - loc[2 * j], dy = loc[2 * i + 1] - loc[2 * j + 1];
            double dist = sqrt((double)(dx * dx + dy * dy));
            if (dist <= (double)lock) collision = 1;
        }
        if (!collision) { loc[2 * i] = x; loc[2 * i + 1] = y; } /* :265-266 */
        else blocked++;
    }
    return blocked;
}

/* ------------------------------------------------------------------------- */
/* a8 + a9.  GetDistance channel.py:220-226, GetPassLoss :230-235, GetChannelGain :237-247,
 * GetDLSinrAllDb :259-269 (interferers = every other BS, :85-90). */
void orc_sinr_all(const orc_cfg *c, const int64_t *ue_xy, const int64_t *bs_xy,
                  const double *fading, double *sinr_db) {
    const int nu = c->n_ue, nb = c->n_bs;
    const double P = pow(10.0, c->p_bs_dbm / 10.0) * 1e-3;     /* channel.py:58 */
    const double N = pow(10.0, c->noise_dbm / 10.0) * 1e-3;    /* channel.py:59 */
    double gain[ORC_MAX_BS], tmp[ORC_MAX_BS];
    for (int u = 0; u < nu; u++) {
        for (int b = 0; b < nb; b++) {
            double ax = (double)ue_xy[2 * u] * c->grid_width - (double)bs_xy[2 * b] * c->grid_width;
            double ay = (double)ue_xy[2 * u + 1] * c->grid_width - (double)bs_xy[2 * b + 1] * c->grid_width;
            double d = sqrt(ax * ax + ay * ay);                /* np.linalg.norm of the 2-D difference */
            double loss = 0;
            if (d > c->pl_dis) loss = c->pl_a + c->pl_b * log10(d);
            double f = fading ? fading[u * nb + b] : 0.0;
            double gdb = c->ant_gain - loss - f - c->eq_loss;  /* :245 */
            gain[b] = pow(10.0, gdb / 10.0);                   /* :246 */
        }
        for (int b = 0; b < nb; b++) {
            int k = 0;
            for (int j = 0; j < nb; j++) if (j != b) tmp[k++] = P * gain[j];
            double p_interf = orc_np_sum(tmp, k);              /* :265 */
            double sinr = P * gain[b] / (N + p_interf);        /* :266 */
            sinr_db[u * nb + b] = 10 * log10(sinr);            /* :268 */
        }
    }
}

/* ------------------------------------------------------------------------- */
orc_chan *orc_chan_create(int n_ue, int n_bs) {
    orc_chan *ch = (orc_chan *)calloc(1, sizeof(orc_chan));
    ch->n_ue = n_ue;
    ch->n_bs = n_bs;
    ch->cur = (int64_t *)calloc(n_ue, sizeof(int64_t));
    ch->cur_sinr = (double *)calloc(n_ue, sizeof(double));
    ch->fifo = (int64_t *)calloc((size_t)ORC_HO_DEPTH * n_ue, sizeof(int64_t));
    ch->out_prev = (uint8_t *)calloc(n_ue, 1);
    return ch;
}

void orc_chan_destroy(orc_chan *ch) {
    if (!ch) return;
    free(ch->cur); free(ch->cur_sinr); free(ch->fifo); free(ch->out_prev);
    free(ch);
}

/* np.argmax / np.max over axis 1: first index wins ties (channel.py:122-123,141-142) */
static inline void best_of(const double *row, int nb, int64_t *arg, double *val) {
    int a = 0;
    double v = row[0];
    for (int b = 1; b < nb; b++) if (row[b] > v) { v = row[b]; a = b; }
    *arg = a; *val = v;
}

/* ctor channel.py:92-93,110 and reset channel.py:113-116 */
void orc_chan_reset(const orc_cfg *c, orc_chan *ch, const double *sinr_db) {
    for (int u = 0; u < ch->n_ue; u++) {
        best_of(sinr_db + (size_t)u * ch->n_bs, ch->n_bs, &ch->cur[u], &ch->cur_sinr[u]);
        ch->fifo[u] = ch->cur[u];
        ch->out_prev[u] = ch->cur_sinr[u] <= c->out_thresh_db;
    }
    ch->fifo_depth = 1;
}

/* UpdateDroneNet channel.py:138-176,216 */
void orc_chan_update(const orc_cfg *c, orc_chan *ch, const double *sinr_db,
                     double *mean_sinr, int32_t *n_out, int32_t *n_ho) {
    const int nu = ch->n_ue, nb = ch->n_bs;
    int64_t *best = (int64_t *)malloc(nu * sizeof(int64_t));
    double *bestv = (double *)malloc(nu * sizeof(double));
    for (int u = 0; u < nu; u++) {
        best_of(sinr_db + (size_t)u * nb, nb, &best[u], &bestv[u]);   /* :141-142 */
        ch->cur_sinr[u] = sinr_db[(size_t)u * nb + ch->cur[u]];       /* :145-146, PRE-handover cell */
    }
    if (ch->fifo_depth < ORC_HO_DEPTH) {                              /* :148-149 */
        memcpy(ch->fifo + (size_t)ch->fifo_depth * nu, best, nu * sizeof(int64_t));
        ch->fifo_depth++;
    } else {                                                          /* :150-153 */
        memmove(ch->fifo, ch->fifo + nu, (size_t)(ORC_HO_DEPTH - 1) * nu * sizeof(int64_t));
        memcpy(ch->fifo + (size_t)(ORC_HO_DEPTH - 1) * nu, best, nu * sizeof(int64_t));
    }
    const int64_t *last = ch->fifo + (size_t)(ch->fifo_depth - 1) * nu;
    int ho = 0;
    for (int u = 0; u < nu; u++) {
        int remain = 1;                                               /* :155 */
        for (int r = 1; r < ch->fifo_depth; r++) if (ch->fifo[(size_t)r * nu + u] != ch->fifo[u]) remain = 0;
        int changed = ch->cur[u] != last[u];                          /* :156 */
        int need = remain && changed && (bestv[u] - ch->cur_sinr[u] > c->ho_thresh_db); /* :158-159 */
        if (need) { ch->cur[u] = last[u]; ho++; }                     /* :162-167 */
    }
    int nout = 0;                                                     /* :170-174 */
    for (int u = 0; u < nu; u++) {
        uint8_t o = ch->cur_sinr[u] <= c->out_thresh_db;
        if (o && !ch->out_prev[u]) nout++;
        ch->out_prev[u] = o;
    }
    *mean_sinr = orc_np_sum(ch->cur_sinr, nu) / (double)nu;           /* np.mean, :216 */
    *n_out = nout;
    *n_ho = ho;
    free(best); free(bestv);
}

/* ------------------------------------------------------------------------- */
/* a12: state[0] = GetGridMap(bsLoc) (ue_mobility.py:173-188; mobile_env.py:160,169),
 *      state[1+b] = association map with the post-handover current_BS (channel.py:401-406) */
void orc_build_state(const orc_cfg *c, const int64_t *ue_xy, const int64_t *bs_xy,
                     const int64_t *cur, double *state) {
    const int64_t G = c->grid_n;
    memset(state, 0, sizeof(double) * (size_t)(c->n_bs + 1) * G * G);
    for (int b = 0; b < c->n_bs; b++) state[bs_xy[2 * b] * G + bs_xy[2 * b + 1]] += 1;
    for (int u = 0; u < c->n_ue; u++)
        state[(1 + cur[u]) * G * G + ue_xy[2 * u] * G + ue_xy[2 * u + 1]] += 1;
}

/* ------------------------------------------------------------------------- */
struct orc_env {
    orc_cfg c;
    int mobility, fading_mode;
    uint64_t seed;
    uint32_t env_id;
    uint32_t tick;       /* mobility ticks done (Philox sequence number) */
    uint32_t epoch;      /* channel passes done (Philox sequence number) */
    int32_t step_n;
    int32_t n_clamped;   /* UE cells clamped from G to G-1 (reference would IndexError, ue_mobility.py:186) */
    orc_mob *mob;
    orc_chan *ch;
    int64_t *ue_xy, *bs_xy, *init_bs_xy;
    double *xy_f, *sinr, *fade_buf;
    const int32_t *trace;
    int64_t trace_T;
};

static void env_cells_from_float(orc_env *e) {
    /* np.concatenate((positions, z), axis=1).astype(int): truncation toward zero (mobile_env.py:96-97,154-155) */
    for (int u = 0; u < e->c.n_ue; u++)
        for (int k = 0; k < 2; k++) {
            int64_t v = (int64_t)e->xy_f[2 * u + k];
            if (v >= e->c.grid_n) { v = e->c.grid_n - 1; e->n_clamped++; }
            e->ue_xy[2 * u + k] = v;
        }
}

static int env_move_ues(orc_env *e, const double *mob_uniforms, int trace_row) {
    if (e->mobility == ORC_MOB_GROUP) {
        if (mob_uniforms) orc_mob_tick(e->mob, mob_uniforms, e->xy_f);
        else orc_mob_tick_philox(e->mob, e->seed, e->env_id, e->tick, e->xy_f);
        e->tick++;
        env_cells_from_float(e);
    } else {
        if (!e->trace || trace_row >= e->trace_T) return -1;   /* IndexError at mobile_env.py:203 */
        for (int u = 0; u < e->c.n_ue; u++) {
            e->ue_xy[2 * u] = e->trace[((size_t)trace_row * e->c.n_ue + u) * 2];
            e->ue_xy[2 * u + 1] = e->trace[((size_t)trace_row * e->c.n_ue + u) * 2 + 1];
        }
    }
    return 0;
}

void orc_philox_fading(const orc_cfg *c, uint64_t seed, uint32_t env_id, uint32_t epoch, double *out) {
    const int nu = c->n_ue, nb = c->n_bs, cpu = (nb + 3) / 4;
    for (int u = 0; u < nu; u++)
        for (int q = 0; q < cpu; q++) {
            double z[4];
            orc_philox_normal4(seed, env_id, (uint32_t)(u * cpu + q), epoch, DOM_FADING, z);
            for (int k = 0; k < 4 && 4 * q + k < nb; k++)
                out[u * nb + 4 * q + k] = c->shadow_mean + c->shadow_sd * z[k];
        }
}

static const double *env_fading(orc_env *e, const double *injected) {
    if (e->fading_mode == ORC_FADE_NONE) return NULL;
    if (e->fading_mode == ORC_FADE_INJECTED) return injected;
    /* Philox: N(mean, sd) per pair (replaces np.random.normal, channel.py:240).  One Philox4x32-10 call yields
     * four normals (two Box-Muller pairs from 32-bit uniforms): BS b of UE u takes normal (b & 3) of call
     * idx = u * ceil(nBS/4) + (b >> 2), sequence number = channel-pass epoch. */
    orc_philox_fading(&e->c, e->seed, e->env_id, e->epoch, e->fade_buf);
    return e->fade_buf;
}

orc_env *orc_env_create(const orc_cfg *c, const int32_t *group_sizes, const int32_t *init_bs_xy,
                        int mobility, int fading, uint64_t seed, uint32_t env_id, int warmup_ticks) {
    orc_env *e = (orc_env *)calloc(1, sizeof(orc_env));
    e->c = *c;
    e->mobility = mobility;
    e->fading_mode = fading;
    e->seed = seed;
    e->env_id = env_id;
    const int nu = c->n_ue, nb = c->n_bs;
    e->ue_xy = (int64_t *)calloc(2 * (size_t)nu, sizeof(int64_t));
    e->bs_xy = (int64_t *)calloc(2 * (size_t)nb, sizeof(int64_t));
    e->init_bs_xy = (int64_t *)calloc(2 * (size_t)nb, sizeof(int64_t));
    e->xy_f = (double *)calloc(2 * (size_t)nu, sizeof(double));
    e->sinr = (double *)calloc((size_t)nu * nb, sizeof(double));
    e->fade_buf = (double *)calloc((size_t)nu * nb, sizeof(double));
    e->ch = orc_chan_create(nu, nb);
    if (init_bs_xy) {
        for (int i = 0; i < 2 * nb; i++) e->init_bs_xy[i] = init_bs_xy[i];
    } else {
        /* mobile_env.py:49-50: (G/4,G/4) (G/4,3G/4) (3G/4,G/4) (3G/4,3G/4) with xMax = grid_n */
        const int G = c->grid_n;
        const int64_t xs[4] = {(int64_t)(G / 4.0), (int64_t)(G / 4.0), (int64_t)(G * 3 / 4.0), (int64_t)(G * 3 / 4.0)};
        const int64_t ys[4] = {(int64_t)(G / 4.0), (int64_t)(G * 3 / 4.0), (int64_t)(G / 4.0), (int64_t)(G * 3 / 4.0)};
        for (int b = 0; b < nb && b < 4; b++) { e->init_bs_xy[2 * b] = xs[b]; e->init_bs_xy[2 * b + 1] = ys[b]; }
    }
    memcpy(e->bs_xy, e->init_bs_xy, 2 * (size_t)nb * sizeof(int64_t));
    if (mobility == ORC_MOB_GROUP) {
        e->mob = orc_mob_create(c, group_sizes);
        if (warmup_ticks >= 0) {
            /* mobile_env.py:76-79 (200 warm-up ticks) and :93-97 (one more for the initial positions) */
            orc_mob_init_philox(e->mob, seed, env_id);
            for (int i = 0; i < warmup_ticks; i++) { orc_mob_tick_philox(e->mob, seed, env_id, e->tick, NULL); e->tick++; }
            env_move_ues(e, NULL, 0);
        }
        /* warmup_ticks < 0: the caller loads mobility state with orc_mob_set_state (reference-state replay) */
    }
    return e;
}

/* the LTEChannel constructor pass (mobile_env.py:100; channel.py:92-93,110); separate so that trace-mode
 * tests can install the trace first (mobile_env.py:85-87 uses trace[0]) */
static void env_ctor_channel(orc_env *e, const double *fading) {
    orc_sinr_all(&e->c, e->ue_xy, e->bs_xy, env_fading(e, fading), e->sinr);
    e->epoch++;
    orc_chan_reset(&e->c, e->ch, e->sinr);
}

void orc_env_destroy(orc_env *e) {
    if (!e) return;
    orc_mob_destroy(e->mob);
    orc_chan_destroy(e->ch);
    free(e->ue_xy); free(e->bs_xy); free(e->init_bs_xy);
    free(e->xy_f); free(e->sinr); free(e->fade_buf);
    free(e);
}

void orc_env_set_trace(orc_env *e, const int32_t *trace, int64_t T) {
    e->trace = trace;
    e->trace_T = T;
}

/* exported: run the constructor's channel pass (after set_trace / set_state) */
int orc_env_ctor_channel(orc_env *e, const double *fading) {
    if (e->mobility == ORC_MOB_TRACE) {
        if (env_move_ues(e, NULL, 0)) return -1;  /* ueLoc = trace[0], mobile_env.py:87 */
    }
    env_ctor_channel(e, fading);
    return 0;
}

/* exported: overwrite the UE cells from float positions (reference-state replay, mobile_env.py:94-97) */
void orc_env_set_ue_from_float(orc_env *e, const double *xy) {
    memcpy(e->xy_f, xy, 2 * (size_t)e->c.n_ue * sizeof(double));
    env_cells_from_float(e);
}

int orc_env_reset(orc_env *e, const double *fading, const double *mob_uniforms, double *state_out) {
    memcpy(e->bs_xy, e->init_bs_xy, 2 * (size_t)e->c.n_bs * sizeof(int64_t));  /* mobile_env.py:119 */
    if (env_move_ues(e, mob_uniforms, 0)) return -1;                            /* :122-131 */
    orc_sinr_all(&e->c, e->ue_xy, e->bs_xy, env_fading(e, fading), e->sinr);   /* :136 -> channel.py:113-116 */
    e->epoch++;
    orc_chan_reset(&e->c, e->ch, e->sinr);
    if (state_out) orc_build_state(&e->c, e->ue_xy, e->bs_xy, e->ch->cur, state_out); /* :138-141 */
    e->step_n = 0;                                                              /* :146 */
    return 0;
}

int orc_env_step(orc_env *e, const int32_t *digits, const double *fading, const double *mob_uniforms,
                 double *state_out, orc_step_out *out) {
    if (env_move_ues(e, mob_uniforms, e->step_n)) return -1;            /* mobile_env.py:152-155 / :202-208 */
    out->n_blocked = orc_bs_move(&e->c, e->bs_xy, digits);              /* :157 / :209 */
    orc_sinr_all(&e->c, e->ue_xy, e->bs_xy, env_fading(e, fading), e->sinr); /* :158 -> channel.py:139-140 */
    e->epoch++;
    orc_chan_update(&e->c, e->ch, e->sinr, &out->mean_sinr, &out->n_out, &out->n_ho);
    if (state_out) orc_build_state(&e->c, e->ue_xy, e->bs_xy, e->ch->cur, state_out); /* :160,169-170 */
    out->r_dissect[0] = out->mean_sinr / 20;                            /* :165 */
    out->r_dissect[1] = -1.0 * out->n_out / e->c.n_ue;                  /* :167 */
    e->step_n += 1;                                                     /* :179 */
    out->done = e->step_n >= e->c.max_step;                             /* :186-187 */
    double r = 0 + out->r_dissect[0];                                   /* sum(r_dissect) starts from int 0 */
    r = r + out->r_dissect[1];
    out->reward = r > -1 ? r : -1;                                      /* max(sum, -1), :189 */
    out->step_n = e->step_n;
    return 0;
}

const int64_t *orc_env_ue_xy(const orc_env *e) { return e->ue_xy; }
const int64_t *orc_env_bs_xy(const orc_env *e) { return e->bs_xy; }
const orc_chan *orc_env_chan(const orc_env *e) { return e->ch; }
orc_mob *orc_env_mob(orc_env *e) { return e->mob; }
const double *orc_env_last_sinr(const orc_env *e) { return e->sinr; }
int32_t orc_env_n_clamped(const orc_env *e) { return e->n_clamped; }
int32_t orc_env_step_n(const orc_env *e) { return e->step_n; }
void orc_env_set_step_n(orc_env *e, int32_t s) { e->step_n = s; }

/* ------------------------------------------------------------------------- */
double orc_bench_run(const orc_cfg *c, const int32_t *group_sizes, const int32_t *init_bs_xy,
                     int n_envs, int n_steps, uint64_t seed, uint32_t env_id0, double *checksum) {
    orc_env **envs = (orc_env **)calloc(n_envs, sizeof(orc_env *));
    const size_t ns = (size_t)(c->n_bs + 1) * c->grid_n * c->grid_n;
    double *state = (double *)malloc(ns * sizeof(double));
    double *copy = (double *)malloc(ns * sizeof(double));
    int32_t digits[ORC_MAX_BS];
    for (int i = 0; i < n_envs; i++) {
        envs[i] = orc_env_create(c, group_sizes, init_bs_xy, ORC_MOB_GROUP, ORC_FADE_PHILOX, seed, env_id0 + i, 200);
        orc_env_ctor_channel(envs[i], NULL);
        orc_env_reset(envs[i], NULL, NULL, state);
    }
    struct timespec t0, t1;
    double acc = 0;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (int s = 0; s < n_steps; s++)
        for (int i = 0; i < n_envs; i++) {
            orc_step_out o;
            for (int b = 0; b < c->n_bs; b++) {
                double a, bb;
                orc_philox_uniform2(seed, env_id0 + i, (uint32_t)b, (uint32_t)s, DOM_ACTION, &a, &bb);
                digits[b] = (int32_t)(a * c->n_act);
            }
            orc_env_step(envs[i], digits, NULL, NULL, state, &o);
            memcpy(copy, state, ns * sizeof(double));  /* np.array(self.state), mobile_env.py:194 */
            acc += o.reward + copy[(size_t)(s % ns)];
            if (o.done) orc_env_reset(envs[i], NULL, NULL, state);
        }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    for (int i = 0; i < n_envs; i++) orc_env_destroy(envs[i]);
    free(envs); free(state); free(copy);
    if (checksum) *checksum = acc;
    return (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
}

/* ------------------------------------------------------------------------- */
/* Trace-replay sweep (BASELINE config 5): one read_trace env, `n_steps` step_test calls with the given joint actions
 * and Philox fading (env id = env_id), recording per step the new-outage count, the handover count, the reward and
 * an exact integer hash of current_BS (sum_u (u+1)*(cur[u]+1)).  MAXSTEP/done is ignored like main_test.py:70-103
 * ignores it.  Returns 0, or the step index + 1 at which the trace ran out. */
int orc_replay_run(const orc_cfg *c, const int32_t *trace, int64_t T, uint64_t seed, uint32_t env_id,
                   const int64_t *actions, int n_steps, int32_t *n_out, int32_t *n_ho, double *reward,
                   int64_t *serving_hash) {
    int32_t gs[64];
    int32_t digits[ORC_MAX_BS];
    for (int g = 0; g < c->n_groups; g++) gs[g] = c->n_ue / c->n_groups + (g < c->n_ue % c->n_groups ? 1 : 0);
    orc_env *e = orc_env_create(c, gs, NULL, ORC_MOB_TRACE, ORC_FADE_PHILOX, seed, env_id, 0);
    orc_env_set_trace(e, trace, T);
    int rc = orc_env_ctor_channel(e, NULL);
    if (!rc) rc = orc_env_reset(e, NULL, NULL, NULL);
    for (int s = 0; s < n_steps && !rc; s++) {
        orc_step_out o;
        orc_action_digits(actions[s], c->n_act, c->n_bs, digits);
        if (orc_env_step(e, digits, NULL, NULL, NULL, &o)) { rc = s + 1; break; }
        n_out[s] = o.n_out; n_ho[s] = o.n_ho; reward[s] = o.reward;
        const orc_chan *ch = orc_env_chan(e);
        int64_t h = 0;
        for (int u = 0; u < c->n_ue; u++) h += (int64_t)(u + 1) * (ch->cur[u] + 1);
        serving_hash[s] = h;
    }
    orc_env_destroy(e);
    return rc;
}

/* A (T, nUE, 2) int32 cell trace from the oracle's own reference_point_group port (Philox draws of env `env_id`,
 * 200 warm-up ticks like mobile_env.py:77-79) -- how the README says ue_trace_10k.npy was made (README.md:31-32). */
void orc_make_trace(const orc_cfg *c, uint64_t seed, uint32_t env_id, int64_t T, int32_t *trace_out) {
    int32_t gs[64];
    for (int g = 0; g < c->n_groups; g++) gs[g] = c->n_ue / c->n_groups + (g < c->n_ue % c->n_groups ? 1 : 0);
    orc_mob *m = orc_mob_create(c, gs);
    double *xy = (double *)malloc(sizeof(double) * 2 * c->n_ue);
    orc_mob_init_philox(m, seed, env_id);
    uint32_t tick = 0;
    for (int i = 0; i < 200; i++) orc_mob_tick_philox(m, seed, env_id, tick++, NULL);
    for (int64_t t = 0; t < T; t++) {
        orc_mob_tick_philox(m, seed, env_id, tick++, xy);
        for (int u = 0; u < c->n_ue; u++) {
            int x = (int)xy[2 * u], y = (int)xy[2 * u + 1];
            if (x >= c->grid_n) x = c->grid_n - 1;
            if (y >= c->grid_n) y = c->grid_n - 1;
            trace_out[(t * c->n_ue + u) * 2] = x;
            trace_out[(t * c->n_ue + u) * 2 + 1] = y;
        }
    }
    free(xy);
    orc_mob_destroy(m);
}

/* ------------------------------------------------------------------------- */
/* GetSinrInArea, channel.py:411-433: per cell (x, y) in [xMin, xMax) x [yMin, yMax) = [1, G)^2 (row / column 0 stay 0)
 * the downlink SINR from the NEAREST BS (np.argmin: first minimum, :420-422), interference from all the other BSs in
 * ascending index order with a fresh fading draw per gain (:425-427, GetChannelGain :237-247), then the serving gain
 * with one more draw (:429).  P_interf is a plain sequential Python sum.
 *   fading != NULL: the draws in the reference's order, (G-1)^2 * nBS values (x outer, y inner; per cell: interferers
 *                   ascending, then the serving BS)
 *   fading == NULL and by_bs != NULL: by_bs[cell * nBS + b] is the draw used for BS b at that cell (cell = (x-1)*(G-1)+(y-1))
 *   both NULL: no fading. */
void orc_sinr_in_area(const orc_cfg *c, const int64_t *bs_xy, const double *fading, const double *by_bs, double *out) {
    const int G = c->grid_n, nb = c->n_bs;
    const double P = pow(10.0, c->p_bs_dbm / 10.0) * 1e-3, N = pow(10.0, c->noise_dbm / 10.0) * 1e-3;
    memset(out, 0, sizeof(double) * (size_t)G * G);
    size_t k = 0;
    for (int x = 1; x < G; x++)
        for (int y = 1; y < G; y++) {
            double dist[ORC_MAX_BS];
            int bs_id = 0;
            for (int b = 0; b < nb; b++) {
                double ax = (double)x * c->grid_width - (double)bs_xy[2 * b] * c->grid_width;
                double ay = (double)y * c->grid_width - (double)bs_xy[2 * b + 1] * c->grid_width;
                dist[b] = sqrt(ax * ax + ay * ay);
                if (dist[b] < dist[bs_id]) bs_id = b;
            }
            const size_t cell = (size_t)(x - 1) * (G - 1) + (y - 1);
            double p_interf = 0, g_srv = 0;
            for (int pass = 0; pass < 2; pass++)
                for (int b = 0; b < nb; b++) {
                    if ((pass == 0) == (b == bs_id)) continue;          /* pass 0: interferers, pass 1: serving */
                    double loss = 0;
                    if (dist[b] > c->pl_dis) loss = c->pl_a + c->pl_b * log10(dist[b]);
                    double f = fading ? fading[k++] : (by_bs ? by_bs[cell * nb + b] : 0.0);
                    double gain = pow(10.0, (c->ant_gain - loss - f - c->eq_loss) / 10.0);
                    if (pass == 0) p_interf += P * gain; else g_srv = gain;
                }
            out[(size_t)x * G + y] = 10 * log10(P * g_srv / (N + p_interf));
        }
}

/* the draws the repo's Philox scheme gives GetSinrInArea call number `seq` of env `env_id`: by_bs[cell * nBS + b] */
void orc_philox_area_fading(const orc_cfg *c, uint64_t seed, uint32_t env_id, uint32_t seq, double *by_bs) {
    const int G = c->grid_n, nb = c->n_bs, cpu = (nb + 3) / 4;
    for (int cell = 0; cell < (G - 1) * (G - 1); cell++)
        for (int q = 0; q < cpu; q++) {
            double z[4];
            orc_philox_normal4(seed, env_id, (uint32_t)(cell * cpu + q), seq, 11 /* DOM_AREA */, z);
            for (int k = 0; k < 4 && 4 * q + k < nb; k++) by_bs[(size_t)cell * nb + 4 * q + k] = c->shadow_mean + c->shadow_sd * z[k];
        }
}

/* ------------------------------------------------------------------------- */
/* Group-mode run of one env (the training path, env.step): reset, then n_steps steps with the given joint actions,
 * Philox mobility + fading of env `env_id`, reset again whenever done (main.py:188-190).  Records per step the
 * new-outage count, handover count, reward and the serving-BS hash, and a hash of the UE cells
 * (sum_u (u+1) * (x_u * G + y_u + 1)). */
int orc_group_run(const orc_cfg *c, uint64_t seed, uint32_t env_id, const int64_t *actions, int n_steps,
                  int32_t *n_out, int32_t *n_ho, double *reward, int64_t *serving_hash, int64_t *cell_hash) {
    int32_t gs[64];
    int32_t digits[ORC_MAX_BS];
    for (int g = 0; g < c->n_groups; g++) gs[g] = c->n_ue / c->n_groups + (g < c->n_ue % c->n_groups ? 1 : 0);
    orc_env *e = orc_env_create(c, gs, NULL, ORC_MOB_GROUP, ORC_FADE_PHILOX, seed, env_id, 200);
    orc_env_ctor_channel(e, NULL);
    orc_env_reset(e, NULL, NULL, NULL);
    for (int s = 0; s < n_steps; s++) {
        orc_step_out o;
        orc_action_digits(actions[s], c->n_act, c->n_bs, digits);
        orc_env_step(e, digits, NULL, NULL, NULL, &o);
        n_out[s] = o.n_out; n_ho[s] = o.n_ho; reward[s] = o.reward;
        const orc_chan *ch = orc_env_chan(e);
        const int64_t *xy = orc_env_ue_xy(e);
        int64_t h = 0, hc = 0;
        for (int u = 0; u < c->n_ue; u++) {
            h += (int64_t)(u + 1) * (ch->cur[u] + 1);
            hc += (int64_t)(u + 1) * (xy[2 * u] * c->grid_n + xy[2 * u + 1] + 1);
        }
        serving_hash[s] = h;
        cell_hash[s] = hc;
        if (o.done) orc_env_reset(e, NULL, NULL, NULL);
    }
    orc_env_destroy(e);
    return 0;
}
