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


def _streaming_parse_audio(module):
    """parse_audio_from_file (describealign.py:149-157) over the decode hand-off: ffmpeg's pipe goes straight to HBM.
    Mono (the default mode) returns the DevicePcm itself - combine() only hands it to the three feature functions
    and deletes it (:1097-1117); stereo (--stretch_audio) needs the samples on the host afterwards, so the float16
    (channels, samples) array is returned and the features computed on the device are remembered for it."""
    from . import api, decode

    def parse_audio_from_file(media_file, num_channels=2):
        ffmpeg = module.get_ffmpeg() if hasattr(module, "get_ffmpeg") else "ffmpeg"
        pcm = decode.parse_audio_from_file(media_file, num_channels, ffmpeg=ffmpeg)
        pcm.wait()
        if num_channels == 1:
            return pcm
        arr = pcm.to_host()
        api.remember_features(arr, pcm.features())
        pcm.close()
        return arr
    return parse_audio_from_file


def patch(module=None, log10: str | None = None, decode: bool = False):
    """Replace the hot-path functions of `module` (default: import describealign).

    decode: also replace `parse_audio_from_file` by the streaming decode hand-off (describealign_b200.decode);
    off by default - it needs the ffmpeg binary the reference would use and has only been exercised with stand-in
    decoders (no ffmpeg in the build image).

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
    if decode and hasattr(module, "parse_audio_from_file"):
        setattr(module, "_reference_parse_audio_from_file", module.parse_audio_from_file)
        module.parse_audio_from_file = _streaming_parse_audio(module)
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
