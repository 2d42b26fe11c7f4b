"""Decode hand-off: the decoder's s16le pipe goes straight to HBM.

Replaces the tail of the reference's `parse_audio_from_file` (describealign.py:149-157): there ffmpeg writes the
whole track to a pipe, Python collects it in one bytes object, `np.frombuffer(...).astype(np.float16)` converts it
on the host, and only then do the feature functions start.  Here

* `ffmpeg_pcm_command(media_file, num_channels)` is the very command line the reference's ffmpeg-python call
  compiles to (same filter, mapping, sample rate and format);
* `open_stream(fileobj, num_channels)` hands the read end of the pipe to a reader thread inside the library
  (`dab_pcm_reader_*`, csrc/pcm_reader.cu) that moves 8 MiB page-locked chunks to the device while the decoder is
  still producing, and returns a `DevicePcm` at once;
* a `DevicePcm` stands where the reference's float16 `(channels, samples)` array stood: `api.get_energy`,
  `get_zero_crossings`, `get_freq_bands` accept it (the int16 -> float16 conversion of :156 happens inside the
  feature kernel), and `api.align_streams` / `pipeline` run whole pairs from two streams;
* `pipeline(pairs)` starts the decoders of pair k + 1 before it aligns pair k, so decoding overlaps the GPU work
  (the reference's loop, describealign.py:1077, does one after the other).

In `--stretch_audio` mode the reference needs the samples on the host afterwards (describealign.py:1142-1150);
`DevicePcm.to_host()` returns the reference's float16 `(channels, samples)` array for that case.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

from . import _cabi, api

AUDIO_SAMPLE_RATE = 44100      # describealign.py:101


def ffmpeg_pcm_command(media_file: str, num_channels: int = 2, ffmpeg: str = "ffmpeg") -> list[str]:
    """Argument list of the reference's decode call (describealign.py:152-154):
    ffmpeg.input(media_file).output('-', format='s16le', acodec='pcm_s16le', af='aresample=async=1:first_pts=0',
    map='0:a:0', ac=num_channels, ar=44100, loglevel='error'), as ffmpeg-python compiles it (output options in
    sorted keyword order, `format` spelled -f)."""
    return [ffmpeg, "-i", str(media_file), "-f", "s16le", "-ac", str(int(num_channels)), "-acodec", "pcm_s16le",
            "-af", "aresample=async=1:first_pts=0", "-ar", str(AUDIO_SAMPLE_RATE), "-loglevel", "error",
            "-map", "0:a:0", "-"]


class DevicePcm:
    """One decoded track on its way into (or already in) HBM: int16 interleaved, as ffmpeg wrote it.

    Array-like enough for the reference's call pattern (describealign.py:1098-1125): `.shape` is the
    `(channels, samples)` of the float16 array it replaces, `.dtype` float16; the feature functions of
    describealign_b200.api take it directly."""

    dtype = np.dtype(np.float16)
    ndim = 2

    def __init__(self, reader, channels: int, process=None, what: str = ""):
        self._lib = _cabi.load()
        self._reader = reader
        self.channels = int(channels)
        self._process = process
        self._what = what
        self._ptr = None
        self._bytes = None
        self._features = None

    # -- completion -------------------------------------------------------------------------
    def wait(self):
        """Block until the decoder closed its end of the pipe and every chunk is on the device (the interpreter
        lock is released meanwhile).  Raises the reference's errors for a failed decode."""
        if self._ptr is not None:
            return self
        ptr, nbytes = ctypes.c_void_p(), ctypes.c_int64()
        rc = self._lib.dab_pcm_reader_wait(self._reader, ctypes.byref(ptr), ctypes.byref(nbytes))
        err = b""
        if self._process is not None:
            err = self._process.stderr.read() if self._process.stderr is not None else b""
            self._process.wait()
        if self._process is not None and len(err) > 0:
            # the reference's own report of a failed decode (describealign.py:125-133)
            print("  ERROR: ffmpeg failed to " + self._what)
            print("FFmpeg error:")
            print(err.decode("utf-8", "replace"))
            raise ChildProcessError("FFmpeg error.")
        if rc != 0:
            raise _cabi.DabError(f"pcm reader failed ({rc}): {self._lib.dab_last_error(api.context().handle).decode()}")
        if nbytes.value % (2 * self.channels) != 0:
            # np.frombuffer / reshape((-1, num_channels)) of describealign.py:156 fail on a torn stream
            raise ValueError("decoded stream is not a whole number of %d-channel int16 samples" % self.channels)
        self._ptr, self._bytes = ptr.value or 0, nbytes.value
        return self

    @property
    def progress_bytes(self) -> int:
        return int(self._lib.dab_pcm_reader_progress(self._reader)) if self._reader else int(self._bytes or 0)

    @property
    def samples(self) -> int:
        self.wait()
        return self._bytes // (2 * self.channels)

    @property
    def shape(self):
        return (self.channels, self.samples)

    def device(self):
        """(device pointer, samples per channel, channels) for AlignJob.load_pcm_device."""
        self.wait()
        return (self._ptr, self.samples, self.channels)

    # -- what the reference does with the array -----------------------------------------------
    def features(self):
        """The five feature vectors, computed once from the device-resident samples."""
        if self._features is None:
            pair = api.acquire_pair()
            try:
                pair.set_pcm_device(_cabi.VIDEO, *self.device())
                self._features = pair.get_features(_cabi.VIDEO)
            finally:
                api.release_pair(pair)
        return self._features

    def to_host(self) -> np.ndarray:
        """The reference's float16 (channels, samples) array (describealign.py:156), for --stretch_audio."""
        ptr, samples, ch = self.device()
        out = np.empty(samples * ch, dtype=np.int16)
        rc = self._lib.dab_pcm_reader_copy_to_host(self._reader, out.ctypes.data, out.nbytes)
        if rc != 0:
            raise _cabi.DabError(f"dab_pcm_reader_copy_to_host failed ({rc})")
        return out.astype(np.float16).reshape((-1, ch)).T

    def close(self):
        """Free the device copy (after the features of this track have been computed)."""
        if self._reader is not None:
            self._lib.dab_pcm_reader_close(self._reader)
            self._reader = None
            self._ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def open_stream(source, num_channels: int = 2, expected_seconds: float | None = None, process=None, what: str = "") -> DevicePcm:
    """Start moving an s16le stream to the device.  source: a file object with fileno() (the stdout of a decoder
    process, an open file) or a file descriptor.  Returns immediately."""
    fd = source if isinstance(source, int) else source.fileno()
    expected = int(expected_seconds * AUDIO_SAMPLE_RATE) * 2 * num_channels if expected_seconds else 0
    lib = _cabi.load()
    h = ctypes.c_void_p()
    ctx = api.context()
    ctx.check(lib.dab_pcm_reader_open(ctx.handle, int(fd), int(expected), ctypes.byref(h)))
    pcm = DevicePcm(h, num_channels, process=process, what=what)
    pcm._source = source      # keep the pipe open for as long as the reader needs it
    return pcm


def parse_audio_from_file(media_file, num_channels: int = 2, ffmpeg: str = "ffmpeg", command=None) -> DevicePcm:
    """Drop-in for the reference's parse_audio_from_file (describealign.py:149-157) that returns a DevicePcm instead of
    a host array.  command: override of the decoder's argument list (tests run without ffmpeg)."""
    cmd = list(command) if command is not None else ffmpeg_pcm_command(media_file, num_channels, ffmpeg)
    proc = subprocess.Popen(cmd, stdin=subprocess.DEVNULL, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    return open_stream(proc.stdout, num_channels, process=proc, what=f"parse audio from input file: {media_file}")


def pipeline(pairs, num_channels: int = 1, ffmpeg: str = "ffmpeg", commands=None):
    """Align a sequence of (video_file, audio_desc_file) pairs; the decoders of pair k + 1 run while pair k is on
    the GPU.  Yields the reference's align() tuple per pair, in order.  commands: optional list of
    (video decoder argv, description decoder argv) replacing the ffmpeg command lines."""
    pairs = list(pairs)

    def start(k):
        v, a = pairs[k]
        cv, ca = commands[k] if commands is not None else (None, None)
        return (parse_audio_from_file(v, num_channels, ffmpeg, cv), parse_audio_from_file(a, num_channels, ffmpeg, ca))

    nxt = start(0) if pairs else None
    for k in range(len(pairs)):
        cur = nxt
        nxt = start(k + 1) if k + 1 < len(pairs) else None
        try:
            yield api.align_streams(cur[0], cur[1])
        finally:
            cur[0].close()
            cur[1].close()
