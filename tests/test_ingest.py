"""Batched ingestion: pkb_wav_* / pkb_wavlist_* against pk_16kpcm_read
(src/pcm_reader.cc:45-220) and the .scp conventions of src/main.cc:34-46. Host code, no GPU."""

import os
import struct

import numpy as np
import pytest

import pocketkaldi_b200 as pkb
from pocketkaldi_b200 import formats
from pocketkaldi_b200.synth import synth_pcm


def write_wav(path, samples, bits, rate=16000, channels=1, fmt=1, patch=None):
    """Canonical 44-byte header; `patch` = {byte offset: bytes} applied afterwards."""
    dt = {8: "<i1", 16: "<i2", 32: "<i4"}[bits]
    data = np.ascontiguousarray(samples, dtype=dt).tobytes()
    raw = bytearray(b"RIFF" + struct.pack("<i", 36 + len(data)) + b"WAVE" + b"fmt " +
                    struct.pack("<ihhiihh", 16, fmt, channels, rate, rate * bits // 8, bits // 8, bits) +
                    b"data" + struct.pack("<i", len(data)) + data)
    for off, b in (patch or {}).items():
        raw[off:off + len(b)] = b
    with open(path, "wb") as fd:
        fd.write(bytes(raw))


@pytest.fixture(scope="module", autouse=True)
def _lib():
    pkb.build_library()


def test_repo_wavs_equal_reference_reader(tmp_path, golden, reference):
    for name in ("hello", "cat"):
        p = str(tmp_path / (name + ".wav"))
        formats.write_wav16(p, golden[name + "_pcm"])
        f = pkb.read_wav(p, np.float32)
        i = pkb.read_wav(p, np.int16)
        assert np.array_equal(i, golden[name + "_pcm"])
        assert np.array_equal(f, golden[name + "_pcm"].astype(np.float32))
        if reference is not None:
            assert np.array_equal(f, reference.read_wav(p))


@pytest.mark.parametrize("bits", [8, 16, 32])
def test_sample_widths_match_reference(tmp_path, reference, bits):
    rng = np.random.default_rng(bits)
    lim = {8: 127, 16: 32767, 32: 2 ** 31 - 1}[bits]
    x = rng.integers(-lim - 1, lim + 1, size=1000, dtype=np.int64)
    p = str(tmp_path / "w.wav")
    write_wav(p, x, bits)
    f = pkb.read_wav(p, np.float32)
    assert np.array_equal(f, x.astype(np.float32))      # unscaled, like the reference
    if reference is not None:
        assert np.array_equal(f, reference.read_wav(p))
    if bits == 32:
        with pytest.raises(pkb.PkbError) as e:
            pkb.read_wav(p, np.int16)
        assert e.value.code == 5
    else:
        assert np.array_equal(pkb.read_wav(p, np.int16), x.astype(np.int16))


@pytest.mark.parametrize("patch,msg", [
    ({0: b"RIFX"}, "chunk_name == 'RIFF' expected"),
    ({4: struct.pack("<i", 7)}, "chunk_size == "),
    ({8: b"WAVX"}, "Format == 'WAVE' expected"),
    ({12: b"fmtx"}, "subchunk1 == 'fmt ' expected"),
    ({16: struct.pack("<i", 18)}, "subchunk1_size == 16 expected, but 18 found"),
    ({20: struct.pack("<h", 3)}, "audio_format == 1 (PCM) expected, but 3 found"),
    ({22: struct.pack("<h", 2)}, "num_channels == 1 (mono) expected, but 2 found"),
    ({24: struct.pack("<i", 8000)}, "sample_rate == 16000 expected, but 8000 found"),
    ({28: struct.pack("<i", 1)}, "bytes_rate == 32000 expected, but 1 found"),
    ({32: struct.pack("<h", 4)}, "block_align == 2 expected, but 4 found"),
    ({36: b"LIST"}, "subchunk2 == 'data' expected"),
    ({40: struct.pack("<i", 10)}, "subchunk2_size == "),
])
def test_header_checks_and_messages(tmp_path, reference, patch, msg):
    p = str(tmp_path / "bad.wav")
    write_wav(p, np.arange(100), 16, patch=patch)
    with pytest.raises(pkb.PkbError) as e:
        pkb.read_wav(p)
    assert e.value.code == 3 and msg in str(e.value) and p in str(e.value)   # Status::Corruption
    if reference is not None:
        with pytest.raises(RuntimeError):
            reference.read_wav(p)


def test_unsupported_width_and_missing_file(tmp_path):
    p = str(tmp_path / "w24.wav")
    raw = bytearray(b"RIFF" + struct.pack("<i", 36 + 6) + b"WAVE" + b"fmt " +
                    struct.pack("<ihhiihh", 16, 1, 1, 16000, 48000, 3, 24) + b"data" +
                    struct.pack("<i", 6) + bytes(6))
    open(p, "wb").write(bytes(raw))
    with pytest.raises(pkb.PkbError) as e:
        pkb.read_wav(p)
    assert e.value.code == 3 and "bits_per_sample == 8, 16 or 32 expected, but 24 found" in str(e.value)
    with pytest.raises(pkb.PkbError) as e:
        pkb.read_wav(str(tmp_path / "nope.wav"))
    assert e.value.code == 2 and "unable to open" in str(e.value)            # Status::IOError


def test_scp_list_reads_back_to_back(tmp_path):
    lens = [1600, 0, 399, 16000, 2560, 401]
    pcm = [synth_pcm(7, [u], n)[0] for u, n in enumerate(lens)]
    paths = []
    for u, x in enumerate(pcm):
        p = str(tmp_path / ("u%d.wav" % u))
        write_wav(p, x, 8 if u == 2 else 16)
        paths.append(p)
    pcm[2] = pcm[2].astype(np.int8).astype(np.int16)
    scp = str(tmp_path / "all.scp")
    with open(scp, "w") as fd:
        fd.write("\r\n".join(paths))           # CRLF endings, no newline after the last line
    for src in (scp, paths):
        wl = pkb.WavList(src)
        assert len(wl) == len(lens) and [wl.path(i) for i in range(len(wl))] == paths
        assert wl.num_samples.tolist() == lens
        for threads in (1, 4):
            got = wl.read_i16(n_threads=threads)
            assert np.array_equal(got, np.concatenate(pcm))
        part = wl.read_i16(first=2, count=3)
        assert np.array_equal(part, np.concatenate(pcm[2:5]))
        assert wl.read_i16(first=1, count=1).size == 0
        with pytest.raises(pkb.PkbError):
            wl.read_i16(first=4, count=5)
        wl.close()


def test_scp_errors(tmp_path):
    with pytest.raises(pkb.PkbError) as e:
        pkb.WavList(str(tmp_path / "missing.scp"))
    assert e.value.code == 2
    good = str(tmp_path / "g.wav")
    write_wav(good, np.arange(10), 16)
    bad = str(tmp_path / "b.wav")
    write_wav(bad, np.arange(10), 16, patch={24: struct.pack("<i", 44100)})
    scp = str(tmp_path / "x.scp")
    open(scp, "w").write(good + "\n" + bad + "\n")
    with pytest.raises(pkb.PkbError) as e:
        pkb.WavList(scp)
    assert e.value.code == 3 and "sample_rate == 16000 expected, but 44100 found" in str(e.value)
    # a 32-bit file is accepted by the list (its header is valid) but not by the int16 reader
    w32 = str(tmp_path / "w32.wav")
    write_wav(w32, np.arange(10), 32)
    wl = pkb.WavList([good, w32])
    with pytest.raises(pkb.PkbError) as e:
        wl.read_i16(n_threads=2)
    assert e.value.code == 5 and "w32.wav" in str(e.value)
    # a file that shrank after the list was opened
    write_wav(good, np.arange(4), 16)
    wl2 = pkb.WavList([good])
    write_wav(good, np.arange(3), 16)
    with pytest.raises(pkb.PkbError):
        wl2.read_i16()
    assert len(pkb.WavList([])) == 0
