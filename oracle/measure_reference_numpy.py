#!/usr/bin/env python
"""TEST INFRASTRUCTURE -- times the UNMODIFIED reference's own numpy step (mobile_env.py:150-194) in the build container.

    python oracle/measure_reference_numpy.py            -> profiles/reference_numpy_step.json

BASELINE.md section 3: `MobiEnvironment(4, 40, 100).step(int(a))` with group mobility and seeded uniform random actions,
imported through oracle/ref_loader.py (in-memory py2->py3 patch, the files under /root/reference are not edited),
`time.perf_counter` around 1000 steps after 50 warm-up steps, best of 3: (i) one core, (ii) one independent OS process per
host core.  /root/reference does not exist on the GPU box, so this measurement cannot be part of a bench.py run there:
the JSON is committed and bench.py echoes it as `cpu_baseline_reference`, labelled with where it was taken.
"""
from __future__ import annotations

import json
import multiprocessing as mp
import os
import platform
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_STEPS, N_WARM, N_REP = 1000, 50, 3


def _one(seed: int) -> float:
    """env-steps/s of one process (best of N_REP)"""
    import numpy as np
    from oracle import ref_loader
    np.random.seed(seed)
    env = ref_loader.make_reference_env(4, 40, 100, "group")
    with ref_loader.quiet_stdout():
        env.reset()
        acts = np.random.RandomState(seed).randint(0, 625, size=N_STEPS + N_WARM)
        for a in acts[:N_WARM]:
            env.step(int(a))
        best = 0.0
        for _ in range(N_REP):
            t0 = time.perf_counter()
            for a in acts[N_WARM:]:
                env.step(int(a))
            best = max(best, N_STEPS / (time.perf_counter() - t0))
    return best


def _cpu_model() -> str:
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return platform.processor() or "unknown"


def main():
    import numpy as np
    nproc = os.cpu_count() or 1
    single = _one(0)
    with mp.get_context("fork").Pool(nproc) as pool:
        per = pool.map(_one, range(1, nproc + 1))
    out = {
        "what": "the reference's own MobiEnvironment(4, 40, 100, 'group').step(int(a)) (mobile_env.py:150-194), unmodified "
                "arithmetic via oracle/ref_loader.py, perf_counter over %d steps after %d warm-up, best of %d" % (N_STEPS, N_WARM, N_REP),
        "where": "build container (no GPU); /root/reference is absent on the GPU box",
        "unit": "env-steps/s", "nproc": nproc, "cpu_model": _cpu_model(),
        "python": platform.python_version(), "numpy": np.__version__,
        "single_core": {"value": single, "cores": 1, "ms_per_step": 1e3 / single},
        "process_per_core": {"value": float(sum(per)), "cores": nproc, "per_process": [float(v) for v in per]},
        "ue_steps_per_s_single_core": single * 40,
    }
    dst = os.path.join(ROOT, "profiles", "reference_numpy_step.json")
    with open(dst, "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
