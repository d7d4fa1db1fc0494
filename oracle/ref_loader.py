"""TEST INFRASTRUCTURE -- not product code.

In-memory loader for the *unmodified* Python-2 reference modules
(``/root/reference/{sinr_visualisation,channel,ue_mobility,mobile_env}.py``).

The reference is Python 2.7 + matplotlib/IPython/pylab.  None of those exist
in this image, so the files cannot be imported as they are.  This loader reads
each file as text, applies the 7-line py2->py3 patch listed in SURVEY.md
section 8(c) *in memory* (the files under /root/reference are never edited),
registers stub modules for the plotting imports and ``exec``s the sources into
fresh module objects.  Nothing is copied into this repository.

It only works where ``/root/reference`` exists (the build container).  It is
used by ``oracle/make_golden.py`` to generate ``tests/golden/*.npz`` and by the
CPU tests that pin ``oracle/mobi_oracle.c`` against the live reference.  It
must never be imported by anything under ``drl_uav_cellularnet_b200/``.

Patches (reference file:line they touch):
  * ``print "..."`` statements   mobile_env.py:55,83,129          -> print(...)
  * ``xrange``                   channel.py:145, sinr_visualisation.py:115 -> range
  * ``dtype=np.int``             ue_mobility.py:423               -> dtype=int
  * ``[range(nBS) ...]``         channel.py:85 (py3 range has no .remove) -> list(range())
"""
from __future__ import annotations

import io
import os
import re
import sys
import types
import contextlib

REFERENCE_DIR = os.environ.get("UAVENV_REFERENCE_DIR", "/root/reference")

_MODULE_ORDER = ("sinr_visualisation", "channel", "ue_mobility", "mobile_env")


def reference_available() -> bool:
    return all(os.path.isfile(os.path.join(REFERENCE_DIR, m + ".py")) for m in _MODULE_ORDER)


class _Anything:
    """Stub object: any attribute access / call returns another stub."""

    def __getattr__(self, name):
        return _Anything()

    def __call__(self, *a, **k):
        return _Anything()

    def update(self, *a, **k):
        return None


def _stub_module(name: str) -> types.ModuleType:
    m = types.ModuleType(name)
    m.__dict__["__getattr__"] = lambda attr: _Anything()
    return m


def _install_stubs() -> None:
    if "matplotlib" not in sys.modules:
        mpl = _stub_module("matplotlib")
        mpl.rcParams = _Anything()
        for sub in ("cm", "ticker", "pyplot", "animation", "colors"):
            sm = _stub_module("matplotlib." + sub)
            # `from matplotlib.pyplot import *` leaks the name `matplotlib`
            # into sinr_visualisation -> channel -> mobile_env (mobile_env.py:15)
            sm.matplotlib = mpl
            sm.__all__ = ["matplotlib"]
            sys.modules["matplotlib." + sub] = sm
            setattr(mpl, sub, sm)
        sys.modules["matplotlib"] = mpl
    if "IPython" not in sys.modules:
        ip = _stub_module("IPython")
        ipd = _stub_module("IPython.display")
        ip.display = ipd
        sys.modules["IPython"] = ip
        sys.modules["IPython.display"] = ipd
    if "pylab" not in sys.modules:
        sys.modules["pylab"] = _stub_module("pylab")


_PRINT_RE = re.compile(r"^(\s*)print (.+)$", re.M)


def _patch_source(name: str, src: str) -> str:
    src = src.replace("\t", "        ")
    src = _PRINT_RE.sub(lambda m: "%sprint(%s)" % (m.group(1), m.group(2)), src)
    src = src.replace("xrange", "range")
    src = src.replace("dtype=np.int)", "dtype=int)")
    if name == "channel":
        src = src.replace(
            "self.interfDL = [range(self.nBS) for bs in range(self.nBS)]",
            "self.interfDL = [list(range(self.nBS)) for bs in range(self.nBS)]",
        )
    return src


_loaded: dict | None = None


def load_reference(quiet: bool = True) -> dict:
    """Return {"mobile_env": module, "channel": module, "ue_mobility": module}."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not reference_available():
        raise FileNotFoundError("reference sources not found under %s" % REFERENCE_DIR)
    _install_stubs()
    saved = {m: sys.modules.get(m) for m in _MODULE_ORDER}
    mods = {}
    try:
        for name in _MODULE_ORDER:
            path = os.path.join(REFERENCE_DIR, name + ".py")
            with open(path, "r") as f:
                src = _patch_source(name, f.read())
            mod = types.ModuleType(name)
            mod.__file__ = path
            sys.modules[name] = mod  # `from channel import *` in mobile_env.py:8
            code = compile(src, path, "exec")
            sink = io.StringIO()
            with contextlib.redirect_stdout(sink) if quiet else contextlib.nullcontext():
                exec(code, mod.__dict__)
            mods[name] = mod
    finally:
        # keep the names private to this loader: the product package has a
        # module called mobile_env too and must never resolve to the reference
        for m, old in saved.items():
            if old is None:
                sys.modules.pop(m, None)
            else:
                sys.modules[m] = old
    _loaded = mods
    return mods


@contextlib.contextmanager
def quiet_stdout():
    """The reference prints on construct/reset/BS lock (mobile_env.py:55,129; ue_mobility.py:268)."""
    sink = io.StringIO()
    with contextlib.redirect_stdout(sink):
        yield


def make_reference_env(nBS=4, nUE=40, grid_n=100, mobility_model="group", trace_file=""):
    mods = load_reference()
    with quiet_stdout():
        return mods["mobile_env"].MobiEnvironment(nBS, nUE, grid_n, mobility_model, trace_file)
