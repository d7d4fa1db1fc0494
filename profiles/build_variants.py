#!/usr/bin/env python
"""Builds tuning variants of libuavenv.so (compile-time knobs) into drl_uav_cellularnet_b200/variants/ for A/B runs on
the GPU box:  UAVENV_SO=<variant.so> python bench.py ...   Usage: python profiles/build_variants.py name:DEF=V,DEF=V ..."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from drl_uav_cellularnet_b200 import build as b  # noqa: E402

for spec in sys.argv[1:]:
    name, _, defs = spec.partition(":")
    out = os.path.join(b.PKG, "variants", name + ".so")
    b.build(force=True, defines=[d for d in defs.split(",") if d], out=out)
    print(out)
