"""Evaluation driver on the B200 path -- host-side mirror of the reference's ``main_test.py``.

``load_ac_net``  = ``Load_AC_Net`` (main_test.py:11-26): actor weights from ``train/<exp>/Global_A_PARA.npz['arr_0']``.
``run_test``     = ``Run_Test`` (main_test.py:46-114): replay a UE trace for MAX_STEP+1 ``step_test`` calls with the
                   greedy action ``argmax(a_prob)`` (main_test.py:68), record reward / SINR / outage / locations /
                   actions / inference time per step, the coverage map every 500 steps (main_test.py:85-89), and write
                   the same nine ``.npy`` files (main_test.py:105-113).
Differences: rows are snapshots (the reference appends aliases of live arrays, so its ``sinr`` and
``bs_location`` files repeat the final state); ``done`` is ignored exactly as the reference ignores it.
"""
from __future__ import annotations

import os
import time
from typing import Optional

import numpy as np
import torch

from .a3c import ACNet
from .env import BatchedMobiEnvironment

MAX_STEP = 2000            # main_test.py:48
FILES = ("reward", "decomposed_reward", "sinr", "time", "outage_fraction", "ue_location", "bs_location", "action",
         "sinr_area")


def load_ac_net(npz_path: str, n_s: int, n_a: int, device) -> ACNet:
    """main_test.py:11-26"""
    net = ACNet(n_s, n_a, device)
    net.load_actor_npz(npz_path)
    return net


def run_test(net: ACNet, trace, out_dir: Optional[str] = None, n_bs: int = 4, n_ue: int = 40, grid_n: int = 100,
             max_step: int = MAX_STEP, precision: str = "fp64", fading: str = "philox", seed: int = 0, device=None) -> dict:
    """main_test.py:46-114 for one env.  Returns the nine arrays (and writes ``<out_dir>/<name>.npy`` if out_dir)."""
    env = BatchedMobiEnvironment(1, n_bs, n_ue, grid_n, "read_trace", trace=trace, precision=precision, fading=fading,
                                 obs="none", seed=seed, device=device)
    env.reset()                                                           # main_test.py:54
    buf = {k: [] for k in FILES}
    for step in range(max_step + 1):                                      # `while step <= MAX_STEP`, main_test.py:70
        t0 = time.time()
        action = net.greedy_action(env.obs_idx)                           # tf.argmax(a_prob, 1), main_test.py:68,73
        torch.cuda.synchronize(env.device)
        buf["time"].append(time.time() - t0)                              # main_test.py:72-74
        _, r, _, info = env.step(action)                                  # step_test, main_test.py:75
        env.check()
        mean_sinr, n_out = float(info["mean_sinr"][0]), int(info["n_out"][0])
        buf["reward"].append(float(r[0]))
        buf["sinr"].append(info["serving_sinr"][0].double().cpu().numpy())
        buf["decomposed_reward"].append([mean_sinr / 20, -1.0 * n_out / n_ue])     # info.r_dissect, mobile_env.py:165,167
        buf["outage_fraction"].append((1.0 * n_out) / n_ue)                          # mobile_env.py:231
        buf["ue_location"].append(info["ue_xy"][0].cpu().numpy().astype(np.int64))
        bs = info["bs_xy"][0].cpu().numpy().astype(np.int64)
        buf["bs_location"].append(np.concatenate([bs, np.full((n_bs, 1), 10, dtype=np.int64)], axis=1))   # H_BS = 10
        buf["action"].append(info["bs_digits"][0].cpu().numpy().astype(np.float64))
        if step % 500 == 0 or step == max_step:                           # main_test.py:85-89
            buf["sinr_area"].append(env.coverage_map()[0].double().cpu().numpy())
    out = {k: np.asarray(v) for k, v in buf.items()}
    if out_dir:
        os.makedirs(out_dir, exist_ok=True)
        for k, v in out.items():
            np.save(os.path.join(out_dir, k), v)                          # main_test.py:105-113
    return out
