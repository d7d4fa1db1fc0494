"""CPU tests (no GPU, no compute calls) of the drop-in boundary: the C-ABI library builds for sm_100a, loads, exports
every symbol include/*.h declares, its configuration struct matches the ctypes mirror and carries the reference's
defaults, argument validation works, and the product package has no path that reaches the oracle."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def native():
    from drl_uav_cellularnet_b200 import build
    build.build()                                   # nvcc cross-compiles without a GPU
    from drl_uav_cellularnet_b200 import _native
    _native.lib()
    return _native


def _declared(header):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(uav(?:env|net)_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported(native):
    L = native.lib()
    declared = _declared("uavenv.h") + _declared("uavnet.h")
    assert len(declared) >= 24
    for name in declared:
        assert hasattr(L, name), name
    assert sorted(set(native.SYMBOLS)) == sorted(set(declared))     # the ctypes mirror binds exactly the header
    assert b"sm_100a" in L.uavenv_version()


def test_cfg_struct_and_reference_defaults(native):
    L = native.lib()
    cfg = native.Cfg()
    assert L.uavenv_cfg_default(C.byref(cfg), 4096, 4, 40, 100) == native.OK
    # module constants of the reference (mobile_env.py:17-32, channel.py:21-82, ue_mobility.py:450-451,473,487)
    assert (cfg.n_envs, cfg.n_bs, cfg.n_ue, cfg.grid_n) == (4096, 4, 40, 100)
    assert (cfg.max_step, cfg.n_act, cfg.bs_step, cfg.min_bs_dist, cfg.warmup_ticks) == (2000, 5, 2, 2, 200)
    assert cfg.n_groups == 4 and list(cfg.group_sizes[:4]) == [10, 10, 10, 10]
    assert (cfg.aggregating0, cfg.deaggregating0, cfg.deaggregating_len, cfg.aggregating_len) == (200, 100, 100, 10)
    assert (cfg.grid_width, cfg.p_bs_dbm, cfg.noise_dbm) == (5.0, 20.0, -121.0)
    assert (cfg.pl_a, cfg.pl_b, cfg.pl_dis, cfg.ant_gain, cfg.eq_loss) == (38.0, 30.0, 0.0, 2.0, 0.0)
    assert (cfg.shadow_mean, cfg.shadow_sd, cfg.ho_thresh_db, cfg.out_thresh_db) == (0.0, 2.0, 1.0, 0.0)
    assert (cfg.v_min, cfg.v_max, cfg.aggregation) == (0.0, 1.0, 0.8)
    # the last field written by the C side is where ctypes expects it: struct layouts agree
    assert C.sizeof(native.Cfg) % 8 == 0
    L.uavenv_cfg_default(C.byref(cfg), 1024, 32, 2048, 100)
    assert cfg.n_groups == 32 and sum(cfg.group_sizes[:32]) == 2048
    assert L.uavenv_cfg_default(None, 1, 4, 40, 100) == native.EINVAL


def test_entry_points_reject_bad_arguments_without_a_gpu(native):
    L = native.lib()
    assert L.uavenv_state_bytes(None) == 0
    assert L.uavenv_launch_count(None) == 0
    assert L.uavenv_last_error(None) == b"null handle"
    assert L.uavenv_get_cfg(None) in (None, 0) or not L.uavenv_get_cfg(None)
    assert L.uavenv_step(None, None, None, None) == native.EINVAL
    assert L.uavenv_reset(None, None, None, None) == native.EINVAL
    assert L.uavenv_check(None, None, None) == native.EINVAL
    assert L.uavenv_launch_plan(None, None, None, None, None) == native.EINVAL
    assert native.diag_lib().uavenv_diag_fill(None, 1024, 1024, 0, None) == -1      # diagnostics library (uavenv_diag.h)
    assert L.uavnet_sparse_fwd(None, 1, 1, 1, None, None, 4, None, 1, None) == -1
    assert L.uavnet_sparse_bwd(None, 1, 1, 1, None, 4, None, None) == -1
    assert L.uavnet_sparse_bwd_gather(None, 1, 1, 1, None, 4, None, None, 1, None) == -1
    assert L.uavnet_sparse_bwd_gather_prepare(None, 1, 1, 1, None, None) == -1
    assert L.uavnet_sparse_bwd_gather_apply(1, 1, 1, None, 4, None, None, 1, None) == -1
    assert L.uavnet_sparse_bwd_gather_apply_cols(1, 1, 1, None, 4, 0, 4, None, None, None) == -1
    assert L.uavnet_p2p_push_part(None, None, None, None, None, None, 0, 1, 1e-4, 0.9, 1e-10, None) == -1
    assert L.uavnet_sparse_bwd_gather_workspace(81920, 44, 50000) >= 4 * (2 * 81920 * 44 + 2 * 50000)
    assert L.uavnet_rmsprop(None, None, None, 4, 1e-4, 0.9, 1e-10, 1.0, 1, None) == -1
    assert L.uavnet_actor_head_bwd(None, None, None, 1, 625, 0.001, None, 625, None, None) == -1
    assert L.uavnet_gemm(None, None) == -1
    assert L.uavnet_rank1_mask(None, None, None, 1, 4, None, None) == -1
    assert L.uavnet_nstep_targets(None, None, None, 1, 1, 0.9, None, None) == -1
    assert L.uavnet_rollout_record(None, None, 1, None, None, None, None, None) == -1
    assert L.uavnet_critic_td(None, None, 1, None, None, None, None) == -1
    assert L.uavnet_mean_rows(None, 1, None, None) == -1
    d = native.GemmDesc()
    assert L.uavnet_gemm(C.byref(d), None) == -1                 # no operands
    assert L.uavnet_gemm_check() == 0                            # nothing launched: no device access
    h = C.c_void_p()
    assert L.uavenv_create(None, C.byref(h)) == native.EINVAL


def test_python_mirror_keeps_the_reference_interface():
    import inspect
    from drl_uav_cellularnet_b200 import BatchedMobiEnvironment, MobiEnvironment, StepInfo, shard_range
    sig = inspect.signature(MobiEnvironment.__init__)
    assert list(sig.parameters)[:6] == ["self", "nBS", "nUE", "grid_n", "mobility_model", "test_mobi_file_name"]
    assert sig.parameters["grid_n"].default == 200 and sig.parameters["mobility_model"].default == "group"  # mobile_env.py:37
    for m in ("reset", "step", "step_test", "SetBsH"):
        assert callable(getattr(MobiEnvironment, m))
    assert StepInfo._fields == ("r_dissect", "step_n", "ue_loc", "bs_loc", "outage_fraction", "bs_actions")  # mobile_env.py:231
    assert shard_range(10, 1, 3) == (4, 7)
    import torch
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError):           # no CPU fallback: construction fails loudly without a GPU
            BatchedMobiEnvironment(2)
        with pytest.raises(ValueError):             # sys.exit("mobility model not defined"), mobile_env.py:91
            BatchedMobiEnvironment(2, mobility_model="random_waypoint")


def test_product_never_touches_the_oracle():
    """Only tests/, bench.py and __graft_entry__.smoke() may use oracle/ (it is the checker, not the product)."""
    pkg = os.path.join(ROOT, "drl_uav_cellularnet_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "mobi_oracle" not in src.replace("oracle/mobi_oracle.c implements the identical scheme", ""), f
                assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), f


def test_n_step_targets_and_net_layout_on_cpu():
    """Host logic of the learner that needs no GPU: the flat parameter layout adds up to the reference's
    20 206 626 parameters and the n-step target recursion matches main.py:223-227."""
    import torch
    from drl_uav_cellularnet_b200.a3c import GAMMA, HIDDEN, n_step_targets
    n_s, n_a, H = 50000, 625, HIDDEN
    actor = n_s * H + H + H * H + H + H * n_a + n_a
    critic = n_s * H + H + H * H + H + H * 1 + 1
    assert actor == 10166025 and critic == 10040601 and actor + critic == 20206626
    r = torch.tensor([[1.0], [2.0], [3.0]], dtype=torch.float64)
    done = torch.tensor([[False], [False], [False]])
    v = n_step_targets(r, done, torch.tensor([10.0], dtype=torch.float64))
    want = []
    v_s_ = 10.0
    for rr in [3.0, 2.0, 1.0]:
        v_s_ = rr + GAMMA * v_s_
        want.append(v_s_)
    assert np.allclose(v[:, 0].numpy(), want[::-1])


def test_dense_wrapper_rejects_what_the_kernel_cannot_take():
    """dense.gemm validates on the host before any device work (CPU tensors, wrong dtypes, non-contiguous rows)."""
    import pytest
    import torch
    from drl_uav_cellularnet_b200 import dense
    a = torch.zeros((4, 4))
    with pytest.raises(ValueError):
        dense.gemm(a, a)                                            # not on a CUDA device: there is no CPU path
    assert set(dense.PRECISIONS) == {"tf32", "fp32", "3xtf32"}


def test_n_step_targets_host_path_matches_the_worker_loop():
    """n_step_targets on CPU float64 (the torch path): v = r + gamma v, cut at episode ends (main.py:217-227)"""
    import numpy as np
    import torch
    from drl_uav_cellularnet_b200.a3c import GAMMA, n_step_targets
    r = torch.tensor([[1.0, 2.0], [3.0, 4.0], [5.0, 6.0]], dtype=torch.float64)
    done = torch.tensor([[False, False], [True, False], [False, False]])
    vb = torch.tensor([10.0, 20.0], dtype=torch.float64)
    got = n_step_targets(r, done, vb).numpy()
    e0 = [1 + GAMMA * 3.0, 3.0, 5 + GAMMA * 10.0]
    e1 = [2 + GAMMA * (4 + GAMMA * (6 + GAMMA * 20.0)), 4 + GAMMA * (6 + GAMMA * 20.0), 6 + GAMMA * 20.0]
    assert np.allclose(got[:, 0], e0) and np.allclose(got[:, 1], e1)
