"""Host-side mirror of the reference environment interface over the CUDA library (include/uavenv.h).

``MobiEnvironment`` keeps the reference's constructor and gym-style contract
(mobile_env.py:37,115,150,196: ``reset() -> state``, ``step(a)``/``step_test(a) -> (state, reward, done, info)``)
for ONE environment and returns host numpy objects exactly like the reference does.
``BatchedMobiEnvironment`` steps E environments per call and returns torch CUDA tensors.

PyTorch is used for device memory and streams only; all arithmetic happens in the sm_100a kernels of
``csrc/env_kernels.cuh``.  There is no CPU path: without a CUDA device construction raises.
"""
from __future__ import annotations

import ctypes as C
from collections import namedtuple
from typing import Optional, Sequence

import numpy as np
import torch

from . import _native as N

# the reference's module constants (mobile_env.py:17-32)
MAXSTEP = 2000
N_ACT = 5
H_BS = 10
MAX_UE_PER_GRID = 1

StepInfo = namedtuple("info_tup", ["r_dissect", "step_n", "ue_loc", "bs_loc", "outage_fraction", "bs_actions"])

_MOBILITY = {"group": N.MOB_GROUP, "read_trace": N.MOB_TRACE}
_FADING = {"philox": N.FADE_PHILOX, "injected": N.FADE_INJECTED, "none": N.FADE_NONE}
_PRECISION = {"fp32": N.PREC_FP32_FAST, "fp64": N.PREC_FP64_PARITY, "fp32_guarded": N.PREC_FP32_GUARDED}
_OBS = {"none": N.OBS_NONE, "f32": N.OBS_F32, "f32_incremental": N.OBS_F32_INCREMENTAL}
_STATE_FIELDS = [("xy", np.float64), ("theta_u", np.float64), ("group", np.float64),
                 ("counters", np.int32), ("bs_xy", np.int16), ("ue_cell", np.int16), ("ho_word", np.uint32)]


def shard_range(n_total: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous env range [lo, hi) of `rank` when n_total envs are sharded over `world` ranks (main.py:173:
    envs are independent, so sharding needs no exchange).  The first n_total % world ranks get one extra."""
    if not (0 <= rank < world):
        raise ValueError("rank %d outside world %d" % (rank, world))
    base, extra = divmod(n_total, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


class BatchedMobiEnvironment:
    """E independent MobiEnvironments advanced by one fused kernel per call.

    Positional arguments follow the reference constructor (mobile_env.py:37) after ``n_envs``.
    Environment e of this object is GLOBAL environment ``env_offset + e``: random draws are keyed by
    (seed, global env id, sequence number, lane), so results do not depend on sharding.
    """

    def __init__(self, n_envs: int, nBS: int = 4, nUE: int = 40, grid_n: int = 100, mobility_model: str = "group",
                 test_mobi_file_name: str = "", *, trace=None, trace_per_env: bool = False, fading: str = "philox",
                 precision: str = "fp32", obs: str = "f32", seed: int = 0, env_offset: int = 0,
                 device: Optional[int] = None, group_sizes: Optional[Sequence[int]] = None,
                 init_bs_xy=None, diagnostics: bool = False, **overrides):
        if mobility_model not in _MOBILITY:
            raise ValueError("mobility model not defined")                 # sys.exit at mobile_env.py:91
        # what __deepcopy__ needs to build a twin handle (copy.deepcopy(env), gradient.py:15)
        self._ctor = dict(n_envs=n_envs, nBS=nBS, nUE=nUE, grid_n=grid_n, mobility_model=mobility_model, fading=fading,
                          precision=precision, obs=obs, seed=seed, env_offset=env_offset, device=device,
                          group_sizes=None if group_sizes is None else list(group_sizes),
                          init_bs_xy=None if init_bs_xy is None else np.array(init_bs_xy), diagnostics=diagnostics,
                          trace_per_env=trace_per_env, **overrides)
        self._trace_np = None
        if not torch.cuda.is_available():
            raise RuntimeError("drl_uav_cellularnet_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self._lib = N.lib()
        self._h = C.c_void_p()
        dev = torch.cuda.current_device() if device is None else int(device)
        self.device = torch.device("cuda", dev)
        cfg = N.Cfg()
        self._lib.uavenv_cfg_default(C.byref(cfg), n_envs, nBS, nUE, grid_n)
        cfg.mobility = _MOBILITY[mobility_model]
        cfg.fading = _FADING[fading]
        cfg.precision = _PRECISION[precision]
        cfg.obs_mode = _OBS[obs]
        cfg.seed, cfg.env_offset, cfg.device = int(seed), int(env_offset), dev
        if group_sizes is not None:
            cfg.n_groups = len(group_sizes)
            for i, g in enumerate(group_sizes):
                cfg.group_sizes[i] = int(g)
        if init_bs_xy is not None:
            b = np.asarray(init_bs_xy, dtype=np.int64).reshape(-1, 2)
            if b.shape[0] != nBS:
                raise ValueError("init_bs_xy must be (nBS, 2)")
            cfg.has_init_bs = 1
            for i, v in enumerate(b.reshape(-1)):
                cfg.init_bs_xy[i] = int(v)
        for k, v in overrides.items():
            if not hasattr(cfg, k):
                raise TypeError("unknown configuration field %r" % k)
            setattr(cfg, k, v)
        rc = self._lib.uavenv_create(C.byref(cfg), C.byref(self._h))
        if rc:
            msg = self._lib.uavenv_last_error(self._h).decode() if self._h else "allocation failed"
            self.close()
            raise (ValueError if rc == N.EINVAL else RuntimeError)("uavenv_create: " + msg)
        self.cfg = self._lib.uavenv_get_cfg(self._h).contents
        self.n_envs, self.nBS, self.nUE, self.grid_n = n_envs, nBS, nUE, grid_n
        self.env_offset = int(env_offset)
        self.mobility_model, self.fading, self.precision, self.obs_mode = mobility_model, fading, precision, obs
        self.action_space_dim = N_ACT ** nBS if cfg.n_act == N_ACT else cfg.n_act ** nBS     # mobile_env.py:104
        self.observation_space_dim = grid_n * grid_n * (nBS + 1) * MAX_UE_PER_GRID           # mobile_env.py:105

        E, d = n_envs, self.device
        ft = torch.float64 if precision == "fp64" else torch.float32
        z = lambda shape, dt: torch.zeros(shape, dtype=dt, device=d)  # noqa: E731
        # state is zeros until the first reset() (mobile_env.py:107)
        self.obs = z((E, nBS + 1, grid_n, grid_n), torch.float32) if obs != "none" else None
        self.reward, self.mean_sinr = z((E,), torch.float64), z((E,), torch.float64)
        self.n_out, self.n_ho, self.n_blocked = z((E,), torch.int32), z((E,), torch.int32), z((E,), torch.int32)
        self.step_n, self.done_u8 = z((E,), torch.int32), z((E,), torch.uint8)
        self.serving, self.serving_sinr = z((E, nUE), torch.uint8), z((E, nUE), ft)
        self.ue_xy, self.bs_xy = z((E, nUE, 2), torch.int16), z((E, nBS, 2), torch.int16)
        self.bs_digits = z((E, nBS), torch.uint8)
        self.obs_idx = z((E, nUE + nBS), torch.int32)       # sparse form of the observation (valid after reset())
        self._own_obs_idx = self.obs_idx
        self.sinr_all = z((E, nUE, nBS), ft) if diagnostics else None
        self.fading_used = z((E, nUE, nBS), torch.float32) if diagnostics else None
        self._out = N.Out(obs=_ptr(self.obs), reward=_ptr(self.reward), mean_sinr=_ptr(self.mean_sinr),
                          n_out=_ptr(self.n_out), n_ho=_ptr(self.n_ho), n_blocked=_ptr(self.n_blocked),
                          done=_ptr(self.done_u8), step_n=_ptr(self.step_n), serving=_ptr(self.serving),
                          serving_sinr=_ptr(self.serving_sinr), sinr_all=_ptr(self.sinr_all),
                          fading_used=_ptr(self.fading_used), ue_xy=_ptr(self.ue_xy), bs_xy=_ptr(self.bs_xy),
                          bs_digits=_ptr(self.bs_digits), obs_idx=_ptr(self.obs_idx))
        self._keep = []          # tensors that must outlive the asynchronous call that reads them
        self._host_args = {}     # step_host: cached ctypes pointers per host-buffer set
        self._ctor_done = mobility_model == "group" and fading != "injected"
        if mobility_model == "read_trace":
            if trace is None:
                assert test_mobi_file_name                                    # mobile_env.py:84
                trace = np.load(test_mobi_file_name)                          # mobile_env.py:85
            self.set_trace(trace, per_env=trace_per_env)
            if fading != "injected":
                self.ctor_pass()

    # -- lifetime -----------------------------------------------------------------------------------------
    _OUT_TENSORS = ("obs", "reward", "mean_sinr", "n_out", "n_ho", "n_blocked", "step_n", "done_u8", "serving", "serving_sinr",
                    "ue_xy", "bs_xy", "bs_digits", "_own_obs_idx", "sinr_all", "fading_used")

    def __deepcopy__(self, memo):
        """copy.deepcopy(env) (gradient.py:15: a virtual env for the look-ahead step): a second handle with the same
        configuration, the same state blob (uavenv_get_state / uavenv_set_state), the same trace and copies of the
        last outputs.  Draws are keyed by (seed, env id, pass counter), and the counters are part of the state, so the
        twin's next step sees exactly the fading / movement the original's next step will see (the reference's twin
        shares the global numpy stream instead and sees the draws that come next in it)."""
        kw = dict(self._ctor)
        kw.pop("trace_per_env")
        if kw["mobility_model"] == "group":
            kw["warmup_ticks"] = -1                     # no warm-up: the state is copied below
        kw["device"] = self.device.index
        twin = BatchedMobiEnvironment(kw.pop("n_envs"), kw.pop("nBS"), kw.pop("nUE"), kw.pop("grid_n"), kw.pop("mobility_model"),
                                      trace=self._trace_np, trace_per_env=getattr(self, "_trace_per_env", False),
                                      **{k: v for k, v in kw.items()}) if self._trace_np is not None or kw["mobility_model"] == "group" \
            else None
        if twin is None:
            raise RuntimeError("deepcopy of a read_trace env needs its trace")
        torch.cuda.synchronize(self.device)
        twin.set_state(self.get_state())
        for name in self._OUT_TENSORS:
            src, dst = getattr(self, name), getattr(twin, name)
            if src is not None and dst is not None:
                dst.copy_(src)
        twin._ctor_done = self._ctor_done
        memo[id(self)] = twin
        return twin

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._lib.uavenv_destroy(self._h)
        self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _raise(self, rc, what):
        msg = self._lib.uavenv_last_error(self._h).decode()
        exc = {N.EINVAL: ValueError, N.EACTION: ValueError, N.ETRACE: IndexError}.get(rc, RuntimeError)
        raise exc("%s: %s" % (what, msg))

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    # -- inputs -------------------------------------------------------------------------------------------
    def _dev(self, a, dtype, shape, what):
        if a is None:
            return None
        t = a if isinstance(a, torch.Tensor) else torch.as_tensor(np.ascontiguousarray(a))
        t = t.to(device=self.device, dtype=dtype).contiguous()
        if tuple(t.shape) != tuple(shape):
            raise ValueError("%s must have shape %s, got %s" % (what, tuple(shape), tuple(t.shape)))
        self._keep.append(t)
        return t

    def _make_in(self, action=None, fading=None, mob_uniforms=None, env_mask=None):
        self._keep = []
        E, nBS, nUE = self.n_envs, self.nBS, self.nUE
        i = N.In()
        if action is not None:
            if isinstance(action, torch.Tensor) and action.dim() == 2 or \
                    (not isinstance(action, torch.Tensor) and np.ndim(action) == 2):
                i.digits = _ptr(self._dev(action, torch.uint8, (E, nBS), "digits"))
            else:
                if not isinstance(action, torch.Tensor):
                    action = np.asarray(action, dtype=np.int64).reshape(-1)       # int, 0-d, shape-(1,)
                i.action = _ptr(self._dev(action.reshape(-1), torch.int64, (E,), "action"))
        if fading is not None:
            i.fading = _ptr(self._dev(fading, torch.float64, (E, nUE, nBS), "fading"))
        if mob_uniforms is not None:
            i.mob_uniforms = _ptr(self._dev(mob_uniforms, torch.float64, (E, nUE + 3 * self.cfg.n_groups), "mob_uniforms"))
        if env_mask is not None:
            i.env_mask = _ptr(self._dev(env_mask, torch.uint8, (E,), "env_mask"))
        return i

    def set_trace(self, trace, per_env: bool = False):
        """np.load(test_mobi_file_name) (mobile_env.py:85): (T, nUE, 2|3) ints, or (T, E, nUE, 2|3) per env."""
        tr = np.asarray(trace)
        want = 4 if per_env else 3
        if tr.ndim != want or tr.shape[-2] != self.nUE or tr.shape[-1] < 2 or (per_env and tr.shape[1] != self.n_envs):
            raise ValueError("trace must be (T%s, nUE, 2|3)" % (", E" if per_env else ""))
        tr = np.ascontiguousarray(tr[..., :2], dtype=np.int32)    # distance and maps use x,y only (channel.py:221-222,397-398)
        rc = self._lib.uavenv_set_trace(self._h, C.c_void_p(tr.ctypes.data), tr.shape[0], 1 if per_env else 0)
        if rc:
            self._raise(rc, "set_trace")
        self.trace_len = tr.shape[0]
        self._trace_np, self._trace_per_env = tr, bool(per_env)

    def bind_obs_idx(self, buf: Optional[torch.Tensor] = None):
        """Redirect the sparse observation (uavenv_out.obs_idx) to a caller-owned int32 [E, nUE + nBS] buffer, e.g. the
        next slot of a rollout storage, so that the states need no copy; None = back to the env's own buffer.  The
        buffer is written by every later reset / step (a masked reset only touches the masked envs)."""
        if buf is None:
            buf = self._own_obs_idx
        if not (buf.is_cuda and buf.dtype == torch.int32 and buf.is_contiguous() and tuple(buf.shape) == tuple(self._own_obs_idx.shape)):
            raise ValueError("obs_idx buffer must be a contiguous int32 CUDA tensor of shape %s" % (tuple(self._own_obs_idx.shape),))
        self.obs_idx = buf
        self._out.obs_idx = buf.data_ptr()

    # -- the three passes ---------------------------------------------------------------------------------
    def ctor_pass(self, fading=None):
        """LTEChannel constructor pass (channel.py:92-93,110); only needed for injected fading / late traces."""
        i = self._make_in(fading=fading)
        rc = self._lib.uavenv_ctor_pass(self._h, C.byref(i), C.byref(self._out), self._stream())
        if rc:
            self._raise(rc, "ctor_pass")
        self._ctor_done = True

    def reset(self, env_mask=None, fading=None, mob_uniforms=None):
        """MobiEnvironment.reset (mobile_env.py:115-148) for all envs or those with env_mask != 0."""
        i = self._make_in(fading=fading, mob_uniforms=mob_uniforms, env_mask=env_mask)
        rc = self._lib.uavenv_reset(self._h, C.byref(i), C.byref(self._out), self._stream())
        if rc:
            self._raise(rc, "reset")
        return self.obs

    def step(self, action, fading=None, mob_uniforms=None):
        """MobiEnvironment.step / step_test (mobile_env.py:150-194 / 196-233).

        action: int64 [E] joint actions (MSB-first base-N_ACT digits) or uint8 [E, nBS] per-BS digits.
        Returns (obs, reward, done, info) as device tensors, valid until the next call."""
        i = self._make_in(action=action, fading=fading, mob_uniforms=mob_uniforms)
        rc = self._lib.uavenv_step(self._h, C.byref(i), C.byref(self._out), self._stream())
        if rc:
            self._raise(rc, "step")
        info = {"mean_sinr": self.mean_sinr, "n_out": self.n_out, "n_ho": self.n_ho, "n_blocked": self.n_blocked,
                "step_n": self.step_n, "serving": self.serving, "serving_sinr": self.serving_sinr,
                "ue_xy": self.ue_xy, "bs_xy": self.bs_xy, "bs_digits": self.bs_digits, "obs_idx": self.obs_idx}
        return self.obs, self.reward, self.done_u8.view(torch.bool), info

    step_test = step

    def step_host(self, action_host: torch.Tensor, reward_host: torch.Tensor, done_host: Optional[torch.Tensor] = None,
                  mean_sinr_host: Optional[torch.Tensor] = None, n_out_host: Optional[torch.Tensor] = None,
                  obs_idx_host: Optional[torch.Tensor] = None):
        """uavenv_step_host: int64 [E] actions in (pinned) HOST memory in, reward/done/... in HOST memory out;
        the copies and the stream synchronisation are inside the call.  The dense observation stays on the device;
        obs_idx_host (int32 [E, nUE + nBS], uavenv_step_host_state) also returns the state to the host in sparse form."""
        bufs = (action_host, reward_host, done_host, mean_sinr_host, n_out_host, obs_idx_host)
        key = tuple(map(id, bufs))            # the entry keeps the tensors alive, so an id cannot be reused while it is cached
        ent = self._host_args.get(key)
        if ent is None:                       # validation and the ctypes argument objects: once per buffer set (hot loop)
            for t in bufs[:5]:
                if t is not None and (t.is_cuda or not t.is_contiguous() or t.numel() != self.n_envs):
                    raise ValueError("step_host takes contiguous host tensors of n_envs elements")
            if obs_idx_host is not None and (obs_idx_host.is_cuda or not obs_idx_host.is_contiguous() or obs_idx_host.dtype != torch.int32
                                             or obs_idx_host.numel() != self.n_envs * (self.nUE + self.nBS)):
                raise ValueError("obs_idx_host must be a contiguous int32 host tensor of n_envs * (nUE + nBS) elements")
            if len(self._host_args) > 64:
                self._host_args.clear()
            args = (_ptr(action_host), _ptr(self.obs), _ptr(reward_host), _ptr(done_host), _ptr(mean_sinr_host), _ptr(n_out_host))
            fn = self._lib.uavenv_step_host if obs_idx_host is None else self._lib.uavenv_step_host_state
            if obs_idx_host is not None:
                args = args + (_ptr(obs_idx_host),)
            ent = self._host_args[key] = (fn, args, bufs, tuple(None if t is None else t.data_ptr() for t in bufs))
        fn, args, _, ptrs = ent
        if action_host.data_ptr() != ptrs[0]:     # a cached tensor object was resized / re-pointed: start over
            del self._host_args[key]
            return self.step_host(action_host, reward_host, done_host, mean_sinr_host, n_out_host, obs_idx_host)
        rc = fn(self._h, *args, self._stream())
        if rc:
            self._raise(rc, "step_host")
        return self.obs

    def coverage_map(self, bs_xy=None, fading=None) -> torch.Tensor:
        """LTEChannel.GetSinrInArea (channel.py:411-433) for every env -> [E, G, G] (float32 / float64 per precision).
        bs_xy: [E, nBS, 2|3] BS cells (default: the current ones); fading: float64 [E, (G-1)^2, nBS] draws in the
        reference's call order (default: Philox / none per the env's fading mode)."""
        E, G = self.n_envs, self.grid_n
        self._keep = []
        b = None
        if bs_xy is not None:
            bb = bs_xy if isinstance(bs_xy, torch.Tensor) else torch.as_tensor(np.ascontiguousarray(bs_xy))
            b = self._dev(bb[..., :2], torch.int16, (E, self.nBS, 2), "bs_xy")
        f = self._dev(fading, torch.float64, (E, (G - 1) * (G - 1), self.nBS), "fading") if fading is not None else None
        out = torch.empty((E, G, G), dtype=torch.float64 if self.precision == "fp64" else torch.float32, device=self.device)
        rc = self._lib.uavenv_coverage_map(self._h, _ptr(b), _ptr(f), _ptr(out), self._stream())
        if rc:
            self._raise(rc, "coverage_map")
        return out

    def check(self) -> int:
        """Sticky device-side error flags since the last check; raises like the reference would
        (ValueError for a bad action, IndexError past the trace end, mobile_env.py:203). Synchronises."""
        f = C.c_uint32(0)
        rc = self._lib.uavenv_check(self._h, C.byref(f), self._stream())
        if rc:
            self._raise(rc, "check")
        return int(f.value)

    @property
    def guard_hits(self) -> int:
        """precision="fp32_guarded": UEs re-evaluated in float64 so far (uavenv_guard_hits). Synchronises."""
        v = C.c_int64(0)
        rc = self._lib.uavenv_guard_hits(self._h, C.byref(v), self._stream())
        if rc:
            self._raise(rc, "guard_hits")
        return int(v.value)

    @property
    def launch_plan(self) -> dict:
        """grid / threads of the step kernel, bytes of its TMA zero tile (0 = fallback path), resident CTAs per SM"""
        g, t, z, o = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32()
        self._lib.uavenv_launch_plan(self._h, C.byref(g), C.byref(t), C.byref(z), C.byref(o))
        return {"grid": g.value, "threads": t.value, "tile_bytes": z.value, "ctas_per_sm": o.value}

    @property
    def launch_count(self) -> int:
        return int(self._lib.uavenv_launch_count(self._h))

    # -- state blob (copy.deepcopy(env) in gradient.py:15; checkpoint) -------------------------------------
    def _field_shapes(self):
        E, nBS, nUE, nG = self.n_envs, self.nBS, self.nUE, self.cfg.n_groups
        return [(E, nUE, 2), (E, nUE), (E, 6, nG), (E, 8), (E, nBS, 2), (E, nUE, 2), (E, nUE)]

    def get_state(self) -> dict:
        n = self._lib.uavenv_state_bytes(self._h)
        buf = np.empty(n, dtype=np.uint8)
        rc = self._lib.uavenv_get_state(self._h, C.c_void_p(buf.ctypes.data), n)
        if rc:
            self._raise(rc, "get_state")
        out = {}
        for k, ((name, dt), shape) in enumerate(zip(_STATE_FIELDS, self._field_shapes())):
            off, nb = C.c_int64(), C.c_int64()
            self._lib.uavenv_state_field(self._h, k, C.byref(off), C.byref(nb))
            out[name] = buf[off.value:off.value + nb.value].view(dt).reshape(shape).copy()
        return out

    def set_state(self, state: dict):
        n = self._lib.uavenv_state_bytes(self._h)
        buf = np.zeros(n, dtype=np.uint8)
        for k, ((name, dt), shape) in enumerate(zip(_STATE_FIELDS, self._field_shapes())):
            off, nb = C.c_int64(), C.c_int64()
            self._lib.uavenv_state_field(self._h, k, C.byref(off), C.byref(nb))
            a = np.ascontiguousarray(state[name], dtype=dt).reshape(shape)
            buf[off.value:off.value + nb.value] = a.view(np.uint8).reshape(-1)
        rc = self._lib.uavenv_set_state(self._h, C.c_void_p(buf.ctypes.data), n)
        if rc:
            self._raise(rc, "set_state")


class _ChannelView:
    """What callers read from env.channel (main_test.py:78; gradient.py:20,64)."""

    def __init__(self, env: "MobiEnvironment"):
        self._env = env

    @property
    def current_BS_sinr(self):
        return self._env._b.serving_sinr[0].double().cpu().numpy()

    @property
    def current_BS(self):
        return self._env._b.serving[0].cpu().numpy().astype(np.int64)

    def GetSinrInArea(self, bsLoc, fading=None):
        """channel.py:411-433 (called every 500 evaluation steps, main_test.py:89) -> (G, G) float64"""
        b = np.asarray(bsLoc)[None, :, :2]
        f = None if fading is None else np.asarray(fading, dtype=np.float64).reshape(1, -1, self._env.nBS)
        return self._env._b.coverage_map(b, f)[0].double().cpu().numpy()


class MobiEnvironment:
    """Single-environment drop-in for the reference class (mobile_env.py:35-233): same constructor, same
    ``reset/step/step_test`` returns (host numpy float64 state copy, float reward, bool done, info).

    Differences, all documented in DESIGN.md: random draws come from Philox instead of numpy's global stream;
    ``step`` in read_trace mode moves UEs like ``step_test`` (the reference crashes there, mobile_env.py:67,152);
    returned arrays are snapshots, not aliases.  Runs the float64 parity kernels by default."""

    def __init__(self, nBS, nUE, grid_n=200, mobility_model="group", test_mobi_file_name="", *, precision="fp64",
                 **kw):
        self.nBS, self.nUE, self.grid_n, self.bs_h = nBS, nUE, grid_n, H_BS
        self.mobility_model = mobility_model
        self._b = BatchedMobiEnvironment(1, nBS, nUE, grid_n, mobility_model, test_mobi_file_name,
                                         precision=precision, **kw)
        self.action_space_dim = self._b.action_space_dim
        self.observation_space_dim = self._b.observation_space_dim
        self.state = np.zeros((nBS + 1, grid_n, grid_n))                           # mobile_env.py:107
        self.step_n = 0
        self.channel = _ChannelView(self)

    def __deepcopy__(self, memo):
        """virtual_env = deepcopy(actual_env) (gradient.py:15)"""
        import copy
        twin = MobiEnvironment.__new__(MobiEnvironment)
        twin.nBS, twin.nUE, twin.grid_n, twin.bs_h = self.nBS, self.nUE, self.grid_n, self.bs_h
        twin.mobility_model = self.mobility_model
        twin._b = copy.deepcopy(self._b, memo)
        twin.action_space_dim, twin.observation_space_dim = self.action_space_dim, self.observation_space_dim
        twin.state = np.array(self.state)
        twin.step_n = self.step_n
        twin.channel = _ChannelView(twin)
        memo[id(self)] = twin
        return twin

    def SetBsH(self, h):                                                           # mobile_env.py:111
        self.bs_h = h

    @property
    def bsLoc(self):
        xy = self._b.get_state()["bs_xy"][0].astype(np.int64)
        return np.concatenate([xy, np.full((self.nBS, 1), self.bs_h, dtype=np.int64)], axis=1)

    @property
    def ueLoc(self):
        return self._b.get_state()["ue_cell"][0].astype(np.int64)

    def _sync_state(self):
        self.state = self._b.obs[0].double().cpu().numpy()
        return np.array(self.state)                                                # fresh copy, mobile_env.py:194

    def reset(self, fading=None, mob_uniforms=None):
        self._b.reset(fading=None if fading is None else np.asarray(fading)[None],
                      mob_uniforms=None if mob_uniforms is None else np.asarray(mob_uniforms)[None])
        self._b.check()
        self.step_n = 0
        return self._sync_state()

    def _step(self, action, fading, mob_uniforms):
        a = np.asarray(action)
        if a.ndim == 1 and a.size == self.nBS and self.nBS != 1:
            act = a.astype(np.uint8)[None]                                         # per-BS digit vector
        else:
            v = int(a.reshape(-1)[0])                                              # int, 0-d or shape-(1,) array
            if not (0 <= v < self.action_space_dim):
                raise ValueError("action %d outside [0, %d)" % (v, self.action_space_dim))
            act = np.array([v], dtype=np.int64)
        b = self._b
        b.step(act, fading=None if fading is None else np.asarray(fading)[None],
               mob_uniforms=None if mob_uniforms is None else np.asarray(mob_uniforms)[None])
        b.check()
        state = self._sync_state()
        mean_sinr, n_out = float(b.mean_sinr[0]), int(b.n_out[0])
        r_dissect = [mean_sinr / 20, -1.0 * n_out / self.nUE]                      # mobile_env.py:165,167
        self.step_n = int(b.step_n[0])
        return state, float(b.reward[0]), bool(b.done_u8[0]), r_dissect, n_out

    def step(self, action, ifrender=False, fading=None, mob_uniforms=None):
        state, reward, done, r_dissect, _ = self._step(action, fading, mob_uniforms)
        return state, reward, done, [r_dissect, self.step_n]                       # mobile_env.py:193-194

    def step_test(self, action, ifrender=False, fading=None, mob_uniforms=None):
        state, reward, done, r_dissect, n_out = self._step(action, fading, mob_uniforms)
        b = self._b
        bs = np.concatenate([b.bs_xy[0].cpu().numpy().astype(np.int64),
                             np.full((self.nBS, 1), self.bs_h, dtype=np.int64)], axis=1)
        info = StepInfo(r_dissect, self.step_n, b.ue_xy[0].cpu().numpy().astype(np.int64), bs,
                        (1.0 * n_out) / self.nUE, b.bs_digits[0].cpu().numpy().astype(np.float64))
        return state, reward, done, info                                           # mobile_env.py:231-233
