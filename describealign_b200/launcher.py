"""Entry point that keeps the reference's CLI / GUI and swaps in the B200 hot path.

    python -m describealign_b200.launcher <the reference's own arguments>

imports the user's installed `describealign` module, replaces its four hot-path functions
(describealign.py:545, 557, 575, 595) - and `replace_aligned_segments` (:229), the resynthesis of
`--stretch_audio` - with the ones from this package and then calls its
unchanged `command_line_interface()` (describealign.py:1773).  Because the replacement is
done at import time of this module, it also survives the `spawn` start method the GUI worker
uses on some platforms as long as the worker's target imports describealign_b200.launcher
(SURVEY.md section 8b); CUDA itself is only initialised inside the process that first calls a
replaced function.
"""
from __future__ import annotations

import importlib
import sys

_PATCHED = ("get_energy", "get_zero_crossings", "get_freq_bands", "align")
_PATCHED_IF_PRESENT = {"replace_aligned_segments": "stretch"}     # --stretch_audio resynthesis (describealign.py:229)


def patch(module=None, log10: str | None = None):
    """Replace the hot-path functions of `module` (default: import describealign).

    log10: "portable" / "native" / "auto" (describealign_b200.native_log): which log10 the feature kernel
    follows.  main() uses "auto" - next to the stock reference the results should be those of the numpy
    installed on this host; the mode is switched on first use of the GPU, not here."""
    from . import api
    if log10 is not None:
        api.request_log10_mode(log10)
    if module is None:
        module = importlib.import_module("describealign")
    for name in _PATCHED:
        if not hasattr(module, name):
            raise AttributeError(f"{module.__name__} has no function {name!r} to replace")
        setattr(module, "_reference_" + name, getattr(module, name))
        setattr(module, name, getattr(api, name))
    for name, where in _PATCHED_IF_PRESENT.items():
        if hasattr(module, name):
            setattr(module, "_reference_" + name, getattr(module, name))
            setattr(module, name, getattr(importlib.import_module("." + where, __package__), name))
    return module


def main(argv=None):
    module = patch(log10="auto")
    if argv is not None:
        sys.argv = [sys.argv[0]] + list(argv)
    return module.command_line_interface()


if __name__ == "__main__":
    main()
