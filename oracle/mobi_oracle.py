"""TEST INFRASTRUCTURE -- ctypes binding of the C parity oracle (oracle/mobi_oracle.c).

Only tests/, bench.py's cpu_baseline / ``--impl reference`` leg and
``__graft_entry__.smoke()`` may import this module.  The product package
(``drl_uav_cellularnet_b200``) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libmobi_oracle.so")

MOB_GROUP, MOB_TRACE = 0, 1
FADE_PHILOX, FADE_INJECTED, FADE_NONE = 0, 1, 2


class OrcCfg(C.Structure):
    _fields_ = [
        ("n_bs", C.c_int32), ("n_ue", C.c_int32), ("grid_n", C.c_int32), ("n_groups", C.c_int32),
        ("max_step", C.c_int32), ("n_act", C.c_int32), ("bs_step", C.c_int32), ("min_bs_dist", C.c_int32),
        ("grid_width", C.c_double), ("p_bs_dbm", C.c_double), ("noise_dbm", C.c_double),
        ("pl_a", C.c_double), ("pl_b", C.c_double), ("pl_dis", C.c_double),
        ("ant_gain", C.c_double), ("eq_loss", C.c_double),
        ("shadow_mean", C.c_double), ("shadow_sd", C.c_double),
        ("ho_thresh_db", C.c_double), ("out_thresh_db", C.c_double),
        ("v_min", C.c_double), ("v_max", C.c_double), ("aggregation", C.c_double),
        ("aggregating0", C.c_int32), ("deaggregating0", C.c_int32),
        ("deaggregating_len", C.c_int32), ("aggregating_len", C.c_int32),
    ]


class OrcChan(C.Structure):
    _fields_ = [
        ("n_ue", C.c_int32), ("n_bs", C.c_int32), ("fifo_depth", C.c_int32),
        ("cur", C.POINTER(C.c_int64)), ("cur_sinr", C.POINTER(C.c_double)),
        ("fifo", C.POINTER(C.c_int64)), ("out_prev", C.POINTER(C.c_uint8)),
    ]


class OrcStepOut(C.Structure):
    _fields_ = [
        ("reward", C.c_double), ("mean_sinr", C.c_double), ("r_dissect", C.c_double * 2),
        ("n_out", C.c_int32), ("n_ho", C.c_int32), ("done", C.c_int32), ("step_n", C.c_int32),
        ("n_blocked", C.c_int32),
    ]


def build(force: bool = False) -> str:
    """Compile oracle/mobi_oracle.c -> oracle/_build/libmobi_oracle.so (gcc, a second or two)."""
    src = os.path.join(_HERE, "mobi_oracle.c")
    hdr = os.path.join(_HERE, "mobi_oracle.h")
    if (not force and os.path.isfile(_SO)
            and os.path.getmtime(_SO) >= max(os.path.getmtime(src), os.path.getmtime(hdr))):
        return _SO
    subprocess.check_call(["make", "-B", "-C", _HERE], stdout=subprocess.DEVNULL)
    return _SO


_lib = None


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


def _ip32(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32)) if a is not None else None


def _ip64(a):
    return a.ctypes.data_as(C.POINTER(C.c_int64)) if a is not None else None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    L = C.CDLL(build())
    P = C.POINTER
    vp = C.c_void_p
    L.orc_cfg_default.argtypes = [P(OrcCfg), C.c_int, C.c_int, C.c_int, C.c_int]
    L.orc_np_sum.argtypes = [P(C.c_double), C.c_int64]
    L.orc_np_sum.restype = C.c_double
    L.orc_philox4x32.argtypes = [C.c_uint32] * 6 + [P(C.c_uint32)]
    L.orc_philox_uniform2.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                                      P(C.c_double), P(C.c_double)]
    L.orc_mob_create.argtypes = [P(OrcCfg), P(C.c_int32)]
    L.orc_mob_create.restype = vp
    L.orc_mob_destroy.argtypes = [vp]
    L.orc_mob_init.argtypes = [vp, P(C.c_double)]
    L.orc_mob_init.restype = C.c_int64
    L.orc_mob_tick.argtypes = [vp, P(C.c_double), P(C.c_double)]
    L.orc_mob_tick.restype = C.c_int64
    L.orc_mob_init_philox.argtypes = [vp, C.c_uint64, C.c_uint32]
    L.orc_mob_tick_philox.argtypes = [vp, C.c_uint64, C.c_uint32, C.c_uint32, P(C.c_double)]
    L.orc_mob_state_len.argtypes = [vp]
    L.orc_mob_state_len.restype = C.c_int64
    L.orc_mob_get_state.argtypes = [vp, P(C.c_double)]
    L.orc_mob_set_state.argtypes = [vp, P(C.c_double)]
    L.orc_action_digits.argtypes = [C.c_int64, C.c_int, C.c_int, P(C.c_int32)]
    L.orc_bs_move.argtypes = [P(OrcCfg), P(C.c_int64), P(C.c_int32)]
    L.orc_sinr_all.argtypes = [P(OrcCfg), P(C.c_int64), P(C.c_int64), P(C.c_double), P(C.c_double)]
    L.orc_chan_create.argtypes = [C.c_int, C.c_int]
    L.orc_chan_create.restype = P(OrcChan)
    L.orc_chan_destroy.argtypes = [P(OrcChan)]
    L.orc_chan_reset.argtypes = [P(OrcCfg), P(OrcChan), P(C.c_double)]
    L.orc_chan_update.argtypes = [P(OrcCfg), P(OrcChan), P(C.c_double), P(C.c_double), P(C.c_int32), P(C.c_int32)]
    L.orc_build_state.argtypes = [P(OrcCfg), P(C.c_int64), P(C.c_int64), P(C.c_int64), P(C.c_double)]
    L.orc_env_create.argtypes = [P(OrcCfg), P(C.c_int32), P(C.c_int32), C.c_int, C.c_int, C.c_uint64, C.c_uint32, C.c_int]
    L.orc_env_create.restype = vp
    L.orc_env_ctor_channel.argtypes = [vp, P(C.c_double)]
    L.orc_env_set_ue_from_float.argtypes = [vp, P(C.c_double)]
    L.orc_env_destroy.argtypes = [vp]
    L.orc_env_set_trace.argtypes = [vp, P(C.c_int32), C.c_int64]
    L.orc_env_reset.argtypes = [vp, P(C.c_double), P(C.c_double), P(C.c_double)]
    L.orc_env_step.argtypes = [vp, P(C.c_int32), P(C.c_double), P(C.c_double), P(C.c_double), P(OrcStepOut)]
    L.orc_env_ue_xy.argtypes = [vp]
    L.orc_env_ue_xy.restype = P(C.c_int64)
    L.orc_env_bs_xy.argtypes = [vp]
    L.orc_env_bs_xy.restype = P(C.c_int64)
    L.orc_env_chan.argtypes = [vp]
    L.orc_env_chan.restype = P(OrcChan)
    L.orc_env_mob.argtypes = [vp]
    L.orc_env_mob.restype = vp
    L.orc_env_last_sinr.argtypes = [vp]
    L.orc_env_last_sinr.restype = P(C.c_double)
    L.orc_env_n_clamped.argtypes = [vp]
    L.orc_env_n_clamped.restype = C.c_int32
    L.orc_env_step_n.argtypes = [vp]
    L.orc_env_step_n.restype = C.c_int32
    L.orc_env_set_step_n.argtypes = [vp, C.c_int32]
    L.orc_bench_run.argtypes = [P(OrcCfg), P(C.c_int32), P(C.c_int32), C.c_int, C.c_int, C.c_uint64, C.c_uint32,
                                P(C.c_double)]
    L.orc_bench_run.restype = C.c_double
    L.orc_replay_run.argtypes = [P(OrcCfg), P(C.c_int32), C.c_int64, C.c_uint64, C.c_uint32, P(C.c_int64), C.c_int,
                                 P(C.c_int32), P(C.c_int32), P(C.c_double), P(C.c_int64)]
    L.orc_replay_run.restype = C.c_int
    L.orc_make_trace.argtypes = [P(OrcCfg), C.c_uint64, C.c_uint32, C.c_int64, P(C.c_int32)]
    L.orc_make_trace.restype = None
    L.orc_group_run.argtypes = [P(OrcCfg), C.c_uint64, C.c_uint32, P(C.c_int64), C.c_int, P(C.c_int32), P(C.c_int32),
                                P(C.c_double), P(C.c_int64), P(C.c_int64)]
    L.orc_group_run.restype = C.c_int
    L.orc_sinr_in_area.argtypes = [P(OrcCfg), P(C.c_int64), P(C.c_double), P(C.c_double), P(C.c_double)]
    L.orc_sinr_in_area.restype = None
    L.orc_philox_area_fading.argtypes = [P(OrcCfg), C.c_uint64, C.c_uint32, C.c_uint32, P(C.c_double)]
    L.orc_philox_area_fading.restype = None
    _lib = L
    return L


def default_cfg(n_bs=4, n_ue=40, grid_n=100, n_groups=4, **over) -> OrcCfg:
    c = OrcCfg()
    lib().orc_cfg_default(C.byref(c), n_bs, n_ue, grid_n, n_groups)
    for k, v in over.items():
        setattr(c, k, v)
    return c


def np_sum(a) -> float:
    a = np.ascontiguousarray(a, dtype=np.float64)
    return lib().orc_np_sum(_dp(a), a.size)


def philox4x32(ctr, key):
    out = (C.c_uint32 * 4)()
    lib().orc_philox4x32(*[int(x) for x in ctr], int(key[0]), int(key[1]), out)
    return [int(x) for x in out]


def philox_uniform2(seed, env, idx, seq, domain):
    a, b = C.c_double(), C.c_double()
    lib().orc_philox_uniform2(seed, env, idx, seq, domain, C.byref(a), C.byref(b))
    return a.value, b.value


def action_digits(action: int, base: int = 5, n: int = 4):
    d = np.zeros(n, dtype=np.int32)
    rc = lib().orc_action_digits(int(action), base, n, _ip32(d))
    if rc:
        raise ValueError("orc_action_digits rc=%d" % rc)
    return d


def bs_move(cfg: OrcCfg, loc, digits):
    loc = np.ascontiguousarray(loc, dtype=np.int64).copy()
    digits = np.ascontiguousarray(digits, dtype=np.int32)
    blocked = lib().orc_bs_move(C.byref(cfg), _ip64(loc), _ip32(digits))
    return loc, blocked


def sinr_all(cfg: OrcCfg, ue_xy, bs_xy, fading=None):
    ue_xy = np.ascontiguousarray(ue_xy, dtype=np.int64)
    bs_xy = np.ascontiguousarray(bs_xy, dtype=np.int64)
    out = np.empty((cfg.n_ue, cfg.n_bs), dtype=np.float64)
    f = None if fading is None else np.ascontiguousarray(fading, dtype=np.float64)
    lib().orc_sinr_all(C.byref(cfg), _ip64(ue_xy), _ip64(bs_xy), _dp(f), _dp(out))
    return out


class Mobility:
    """reference_point_group restatement driven by explicit uniforms or Philox (ue_mobility.py:409-523)."""

    def __init__(self, cfg: OrcCfg, group_sizes):
        self.cfg = cfg
        self.gs = np.ascontiguousarray(group_sizes, dtype=np.int32)
        self.h = lib().orc_mob_create(C.byref(cfg), _ip32(self.gs))

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_mob_destroy(self.h)
            self.h = None

    def init(self, uniforms) -> int:
        u = np.ascontiguousarray(uniforms, dtype=np.float64)
        return lib().orc_mob_init(self.h, _dp(u))

    def tick(self, uniforms):
        u = np.ascontiguousarray(uniforms, dtype=np.float64)
        xy = np.empty((self.cfg.n_ue, 2), dtype=np.float64)
        used = lib().orc_mob_tick(self.h, _dp(u), _dp(xy))
        return xy, used

    def init_philox(self, seed, env):
        lib().orc_mob_init_philox(self.h, seed, env)

    def tick_philox(self, seed, env, tick):
        xy = np.empty((self.cfg.n_ue, 2), dtype=np.float64)
        lib().orc_mob_tick_philox(self.h, seed, env, tick, _dp(xy))
        return xy

    def get_state(self):
        out = np.empty(lib().orc_mob_state_len(self.h), dtype=np.float64)
        lib().orc_mob_get_state(self.h, _dp(out))
        return out

    def set_state(self, s):
        s = np.ascontiguousarray(s, dtype=np.float64)
        assert s.size == lib().orc_mob_state_len(self.h)
        lib().orc_mob_set_state(self.h, _dp(s))


class OracleEnv:
    """Single-env restatement of MobiEnvironment (mobile_env.py:35-233)."""

    def __init__(self, cfg: OrcCfg | None = None, group_sizes=None, init_bs_xy=None, mobility=MOB_GROUP,
                 fading=FADE_PHILOX, seed=0, env_id=0, warmup_ticks=200, trace=None, ctor_fading=None,
                 mob_state=None, ue_float=None):
        self.cfg = cfg if cfg is not None else default_cfg()
        c = self.cfg
        if group_sizes is None:
            assert c.n_ue % c.n_groups == 0
            group_sizes = [c.n_ue // c.n_groups] * c.n_groups
        self.gs = np.ascontiguousarray(group_sizes, dtype=np.int32)
        self.ibs = None if init_bs_xy is None else np.ascontiguousarray(init_bs_xy, dtype=np.int32)
        self.mobility, self.fading = mobility, fading
        if mob_state is not None:
            warmup_ticks = -1
        self.h = lib().orc_env_create(C.byref(c), _ip32(self.gs), _ip32(self.ibs), mobility, fading, seed, env_id,
                                      warmup_ticks)
        self.trace = None
        if trace is not None:
            self.trace = np.ascontiguousarray(np.asarray(trace)[:, :, :2], dtype=np.int32)
            lib().orc_env_set_trace(self.h, _ip32(self.trace), self.trace.shape[0])
        if mob_state is not None:
            m = np.ascontiguousarray(mob_state, dtype=np.float64)
            lib().orc_mob_set_state(lib().orc_env_mob(self.h), _dp(m))
            xy = np.ascontiguousarray(ue_float, dtype=np.float64)
            lib().orc_env_set_ue_from_float(self.h, _dp(xy))
        f = None if ctor_fading is None else np.ascontiguousarray(ctor_fading, dtype=np.float64)
        rc = lib().orc_env_ctor_channel(self.h, _dp(f))
        if rc:
            raise IndexError("trace missing/empty")
        self.state = np.zeros((c.n_bs + 1, c.grid_n, c.grid_n), dtype=np.float64)  # mobile_env.py:107

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_env_destroy(self.h)
            self.h = None

    def reset(self, fading=None, mob_uniforms=None):
        f = None if fading is None else np.ascontiguousarray(fading, dtype=np.float64)
        u = None if mob_uniforms is None else np.ascontiguousarray(mob_uniforms, dtype=np.float64)
        rc = lib().orc_env_reset(self.h, _dp(f), _dp(u), _dp(self.state))
        if rc:
            raise IndexError("trace exhausted")
        return np.array(self.state)

    def step(self, action, fading=None, mob_uniforms=None, want_state=True):
        c = self.cfg
        if np.ndim(action) == 1 and np.size(action) == c.n_bs and c.n_bs != 1:
            digits = np.ascontiguousarray(action, dtype=np.int32)
        else:
            digits = action_digits(int(np.asarray(action).reshape(-1)[0]), c.n_act, c.n_bs)
        f = None if fading is None else np.ascontiguousarray(fading, dtype=np.float64)
        u = None if mob_uniforms is None else np.ascontiguousarray(mob_uniforms, dtype=np.float64)
        out = OrcStepOut()
        rc = lib().orc_env_step(self.h, _ip32(digits), _dp(f), _dp(u), _dp(self.state) if want_state else None,
                                C.byref(out))
        if rc:
            raise IndexError("trace exhausted")
        self.last = out
        info = dict(r_dissect=[out.r_dissect[0], out.r_dissect[1]], step_n=out.step_n, n_out=out.n_out,
                    n_ho=out.n_ho, mean_sinr=out.mean_sinr, n_blocked=out.n_blocked, digits=digits)
        return (np.array(self.state) if want_state else None), out.reward, bool(out.done), info

    # live views (copies)
    @property
    def ue_xy(self):
        return np.ctypeslib.as_array(lib().orc_env_ue_xy(self.h), shape=(self.cfg.n_ue, 2)).copy()

    @property
    def bs_xy(self):
        return np.ctypeslib.as_array(lib().orc_env_bs_xy(self.h), shape=(self.cfg.n_bs, 2)).copy()

    @property
    def current_BS(self):
        ch = lib().orc_env_chan(self.h).contents
        return np.ctypeslib.as_array(ch.cur, shape=(self.cfg.n_ue,)).copy()

    @property
    def current_BS_sinr(self):
        ch = lib().orc_env_chan(self.h).contents
        return np.ctypeslib.as_array(ch.cur_sinr, shape=(self.cfg.n_ue,)).copy()

    @property
    def fifo(self):
        ch = lib().orc_env_chan(self.h).contents
        return np.ctypeslib.as_array(ch.fifo, shape=(3, self.cfg.n_ue))[: ch.fifo_depth].copy()

    @property
    def out_prev(self):
        ch = lib().orc_env_chan(self.h).contents
        return np.ctypeslib.as_array(ch.out_prev, shape=(self.cfg.n_ue,)).copy()

    @property
    def last_sinr(self):
        return np.ctypeslib.as_array(lib().orc_env_last_sinr(self.h), shape=(self.cfg.n_ue, self.cfg.n_bs)).copy()

    @property
    def step_n(self):
        return lib().orc_env_step_n(self.h)

    @step_n.setter
    def step_n(self, v):
        lib().orc_env_set_step_n(self.h, int(v))

    @property
    def n_clamped(self):
        return lib().orc_env_n_clamped(self.h)

    def mob_state(self):
        m = lib().orc_env_mob(self.h)
        out = np.empty(lib().orc_mob_state_len(m), dtype=np.float64)
        lib().orc_mob_get_state(m, _dp(out))
        return out


def bench_run(cfg: OrcCfg, n_envs: int, n_steps: int, seed: int = 0, env_id0: int = 0, group_sizes=None,
              init_bs_xy=None):
    """Time n_envs x n_steps oracle env-steps on the calling thread; returns (seconds, checksum)."""
    if group_sizes is None:
        group_sizes = [cfg.n_ue // cfg.n_groups] * cfg.n_groups
    gs = np.ascontiguousarray(group_sizes, dtype=np.int32)
    ibs = None if init_bs_xy is None else np.ascontiguousarray(init_bs_xy, dtype=np.int32)
    chk = C.c_double()
    t = lib().orc_bench_run(C.byref(cfg), _ip32(gs), _ip32(ibs), n_envs, n_steps, seed, env_id0, C.byref(chk))
    return t, chk.value


def make_trace(cfg: OrcCfg, seed: int, env_id: int, T: int):
    """(T, nUE, 2) int32 cells from the oracle's reference_point_group port (how ue_trace_10k.npy was made, README.md:31-32)."""
    out = np.empty((T, cfg.n_ue, 2), dtype=np.int32)
    lib().orc_make_trace(C.byref(cfg), seed, env_id, T, _ip32(out))
    return out


def replay_run(cfg: OrcCfg, trace, seed: int, env_id: int, actions):
    """One read_trace env, len(actions) step_test calls with Philox fading: (n_out, n_ho, reward, serving_hash) per step."""
    tr = np.ascontiguousarray(trace, dtype=np.int32)
    act = np.ascontiguousarray(actions, dtype=np.int64)
    n = len(act)
    n_out, n_ho = np.empty(n, dtype=np.int32), np.empty(n, dtype=np.int32)
    rew, hsh = np.empty(n, dtype=np.float64), np.empty(n, dtype=np.int64)
    rc = lib().orc_replay_run(C.byref(cfg), _ip32(tr), tr.shape[0], seed, env_id, _ip64(act), n, _ip32(n_out), _ip32(n_ho),
                              _dp(rew), _ip64(hsh))
    if rc:
        raise IndexError("trace exhausted at step %d" % (rc - 1))
    return n_out, n_ho, rew, hsh


def sinr_in_area(cfg: OrcCfg, bs_xy, fading=None, by_bs=None):
    """GetSinrInArea (channel.py:411-433) -> (G, G) float64.  fading: the draws in the reference's order; by_bs: draws per
    (cell, BS); neither: no fading."""
    bs = np.ascontiguousarray(np.asarray(bs_xy)[:, :2], dtype=np.int64)
    f = None if fading is None else np.ascontiguousarray(fading, dtype=np.float64)
    g = None if by_bs is None else np.ascontiguousarray(by_bs, dtype=np.float64)
    out = np.empty((cfg.grid_n, cfg.grid_n), dtype=np.float64)
    lib().orc_sinr_in_area(C.byref(cfg), _ip64(bs), _dp(f), _dp(g), _dp(out))
    return out


def philox_area_fading(cfg: OrcCfg, seed: int, env_id: int, seq: int):
    out = np.empty(((cfg.grid_n - 1) ** 2, cfg.n_bs), dtype=np.float64)
    lib().orc_philox_area_fading(C.byref(cfg), seed, env_id, seq, _dp(out))
    return out


def group_run(cfg: OrcCfg, seed: int, env_id: int, actions):
    """One group-mode env (Philox), reset + len(actions) steps (+ reset on done):
    (n_out, n_ho, reward, serving_hash, cell_hash) per step."""
    act = np.ascontiguousarray(actions, dtype=np.int64)
    n = len(act)
    n_out, n_ho = np.empty(n, dtype=np.int32), np.empty(n, dtype=np.int32)
    rew, hsh, hc = np.empty(n, dtype=np.float64), np.empty(n, dtype=np.int64), np.empty(n, dtype=np.int64)
    lib().orc_group_run(C.byref(cfg), seed, env_id, _ip64(act), n, _ip32(n_out), _ip32(n_ho), _dp(rew), _ip64(hsh), _ip64(hc))
    return n_out, n_ho, rew, hsh, hc
