"""Decode hand-off (describealign.py:149-157): s16le from a decoder's pipe straight to the device.
ffmpeg is not in the image, so a small Python process plays the decoder: it writes the PCM in irregular pieces."""
import os
import sys
import tempfile

import numpy as np
import pytest


def test_ffmpeg_command_is_the_references():
    """The reference's ffmpeg-python call (describealign.py:152-154) compiles to this argument list: global -loglevel
    is an output kwarg there, so it stays among the output options, which ffmpeg-python emits in sorted order."""
    from describealign_b200 import decode
    cmd = decode.ffmpeg_pcm_command("in put.mkv", 2)
    assert cmd[:3] == ["ffmpeg", "-i", "in put.mkv"] and cmd[-1] == "-"
    opts = dict(zip(cmd[3:-1:2], cmd[4:-1:2]))
    assert opts == {"-f": "s16le", "-acodec": "pcm_s16le", "-af": "aresample=async=1:first_pts=0", "-map": "0:a:0",
                    "-ac": "2", "-ar": "44100", "-loglevel": "error"}
    assert decode.ffmpeg_pcm_command("x", 1)[cmd.index("-ac") + 1] == "1"


_DECODER = r"""
import sys, time
data = open(sys.argv[1], 'rb').read()
pieces = [1, 3, 4097, 65536, 1 << 20, 9 << 20, 333, 7]     # torn sample boundaries, pieces larger than a chunk
k = at = 0
out = sys.stdout.buffer
while at < len(data):
    n = pieces[k % len(pieces)]; k += 1
    out.write(data[at:at + n]); at += n
    if k % 5 == 0:
        out.flush(); time.sleep(0.002)
"""


def _raw_file(tmp, pcm):
    path = os.path.join(tmp, "track.raw")
    np.ascontiguousarray(pcm).tofile(path)
    return path


@pytest.mark.gpu
@pytest.mark.parametrize("ch,seconds,seed", [(1, 21.7, 401), (2, 33.3, 402)])
def test_streamed_track_gives_the_features_of_the_host_array(gpu_ctx, ch, seconds, seed):
    from describealign_b200 import api, decode, synth
    v, _ = synth.make_pair(seconds, 2.0, seed=seed, ch=ch)
    want = api.track_features(synth.as_reference_input(v))
    with tempfile.TemporaryDirectory() as tmp:
        raw = _raw_file(tmp, v)
        pcm = decode.parse_audio_from_file("track", ch, command=[sys.executable, "-c", _DECODER, raw])
        assert pcm.shape == (ch, v.shape[0]) and pcm.dtype == np.float16
        got = [api.get_energy(pcm), api.get_zero_crossings(pcm)] + api.get_freq_bands(pcm)
        for g, w in zip(got, want):
            assert g.dtype == w.dtype and np.array_equal(g, w)
        # the host copy --stretch_audio needs is the reference's float16 (channels, samples) array
        assert np.array_equal(pcm.to_host(), synth.as_reference_input(v))
        pcm.close()


@pytest.mark.gpu
def test_pipeline_overlaps_decoding_and_matches_align_pcm(gpu_ctx):
    from describealign_b200 import api, decode, synth
    pairs = [synth.make_pair(75.0, 6.0, skips=[(30.0, 2.0)], seed=7), synth.make_pair(80.0, 4.0, skips=[(41.0, -1.5)], seed=8)]
    want = [api.align_pcm(v, a) for v, a in pairs]
    with tempfile.TemporaryDirectory() as tmp:
        files, cmds = [], []
        for k, (v, a) in enumerate(pairs):
            fv, fa = os.path.join(tmp, f"v{k}.raw"), os.path.join(tmp, f"a{k}.raw")
            v.tofile(fv); a.tofile(fa)
            files.append((fv, fa))
            cmds.append(([sys.executable, "-c", _DECODER, fv], [sys.executable, "-c", _DECODER, fa]))
        got = list(decode.pipeline(files, num_channels=1, commands=cmds))
    for g, w in zip(got, want):
        assert np.array_equal(g[0], w[0]) and np.array_equal(g[1], w[1]) and g[2] == w[2]
        assert np.array_equal(g[3], w[3]) and g[4] == w[4]


@pytest.mark.gpu
def test_failed_decoder_raises_like_the_reference(gpu_ctx, capsys):
    from describealign_b200 import decode
    bad = [sys.executable, "-c", "import sys; sys.stderr.write('no such stream'); sys.exit(1)"]
    pcm = decode.parse_audio_from_file("missing.mkv", 1, command=bad)
    with pytest.raises(ChildProcessError, match="FFmpeg error."):
        pcm.wait()
    assert "ERROR: ffmpeg failed to parse audio from input file: missing.mkv" in capsys.readouterr().out
    torn = [sys.executable, "-c", "import sys; sys.stdout.buffer.write(b'abc')"]
    pcm = decode.parse_audio_from_file("torn", 1, command=torn)
    with pytest.raises(ValueError):
        pcm.wait()
