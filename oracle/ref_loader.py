"""TEST INFRASTRUCTURE ONLY - never imported by the product package.

Loads the *unmodified* reference (describealign.py) from its read-only location with
empty stand-ins for the optional third-party modules it imports at top level but never
uses on the alignment hot path.  Used only by
tools/make_golden.py in the authoring container to produce the fixtures under
tests/golden/; the reference is not present on the GPU box and nothing in tests/,
bench.py or smoke() may depend on it at run time.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

REFERENCE_CANDIDATES = ("/root/reference/describealign.py",)
# a pip --target install of the reference (the only form of it that travels to the GPU box).  The
# reference's own packaging does not build from its checkout (setuptools: "multiple top-level packages
# discovered in a flat-layout"), so this directory normally does not exist; bench.py --impl reference
# probes it and nothing else.
INSTALLED_CANDIDATES = (os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref", "describealign.py"),)


def reference_available() -> bool:
    return any(os.path.isfile(p) for p in REFERENCE_CANDIDATES)


def load():
    """The installed reference (baseline/_ref) or None; never reads /root/reference."""
    path = next((p for p in INSTALLED_CANDIDATES if os.path.isfile(p)), None)
    return None if path is None else load_reference(path)


def load_reference(path=None):
    if path is None:
        path = next((p for p in REFERENCE_CANDIDATES if os.path.isfile(p)), None)
    if path is None:
        raise FileNotFoundError("reference describealign.py not found")
    for name in ("matplotlib", "matplotlib.pyplot", "ffmpeg", "static_ffmpeg",
                 "static_ffmpeg.run", "natsort"):
        if name not in sys.modules:
            try:
                importlib.import_module(name)
            except Exception:
                mod = types.ModuleType(name)
                if name == "ffmpeg":
                    mod.Error = Exception
                sys.modules[name] = mod
    if "matplotlib" in sys.modules and not hasattr(sys.modules["matplotlib"], "pyplot"):
        sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    spec = importlib.util.spec_from_file_location("describealign_reference", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod
