"""Host-side input side of the path: RIFF/WAVE decode into float32 in [-1, 1) the way the reference's
loaders see a file (dataset/dataload_nsvae.py:L181-183: ``librosa.load(path, sr=None)`` = native rate, mono, float32,
integer PCM divided by 2^(bits-1)), and the writer for enhanced output.  numpy only - no librosa / soundfile here."""
import struct

import numpy as np


def read_wav(path):
    """-> (float32 mono samples, sample rate).  PCM 8/16/24/32-bit and IEEE float32/64; multi-channel files are
    averaged to mono like librosa's default."""
    with open(path, "rb") as f:
        data = f.read()
    if len(data) < 12 or data[:4] != b"RIFF" or data[8:12] != b"WAVE":
        raise ValueError("%s is not a RIFF/WAVE file" % path)
    pos, fmt, body = 12, None, None
    while pos + 8 <= len(data):
        cid, size = data[pos:pos + 4], struct.unpack("<I", data[pos + 4:pos + 8])[0]
        chunk = data[pos + 8:pos + 8 + size]
        if cid == b"fmt ":
            tag, ch, fs, _, _, bits = struct.unpack("<HHIIHH", chunk[:16])
            if tag == 0xFFFE and len(chunk) >= 26:                     # WAVE_FORMAT_EXTENSIBLE: sub-format GUID
                tag = struct.unpack("<H", chunk[24:26])[0]
            fmt = (tag, ch, fs, bits)
        elif cid == b"data":
            body = chunk
        pos += 8 + size + (size & 1)
    if fmt is None or body is None:
        raise ValueError("%s: missing fmt or data chunk" % path)
    tag, ch, fs, bits = fmt
    if tag == 1:
        if bits == 8:
            x = (np.frombuffer(body, dtype=np.uint8).astype(np.float32) - 128.0) / 128.0
        elif bits == 16:
            x = np.frombuffer(body[:len(body) // 2 * 2], dtype="<i2").astype(np.float32) / 32768.0
        elif bits == 24:
            b = np.frombuffer(body[:len(body) // 3 * 3], dtype=np.uint8).reshape(-1, 3).astype(np.int32)
            v = b[:, 0] | (b[:, 1] << 8) | (b[:, 2] << 16)
            x = (v - ((v & 0x800000) << 1)).astype(np.float32) / 8388608.0
        elif bits == 32:
            x = (np.frombuffer(body[:len(body) // 4 * 4], dtype="<i4").astype(np.float64) / 2147483648.0).astype(np.float32)
        else:
            raise ValueError("%s: %d-bit PCM is not supported" % (path, bits))
    elif tag == 3:
        x = np.frombuffer(body, dtype="<f4" if bits == 32 else "<f8").astype(np.float32)
    else:
        raise ValueError("%s: WAVE format tag %d is not supported" % (path, tag))
    if ch > 1:
        x = x[:len(x) // ch * ch].reshape(-1, ch).mean(axis=1).astype(np.float32)
    return x, fs


def write_wav(path, x, fs=16000):
    """float32 samples -> 16-bit PCM (clipped), the format of the corpora the reference reads."""
    pcm = np.clip(np.round(np.asarray(x, dtype=np.float64) * 32768.0), -32768, 32767).astype("<i2").tobytes()
    hdr = b"RIFF" + struct.pack("<I", 36 + len(pcm)) + b"WAVE" + b"fmt " + struct.pack("<IHHIIHH", 16, 1, 1, fs, fs * 2, 2, 16)
    with open(path, "wb") as f:
        f.write(hdr + b"data" + struct.pack("<I", len(pcm)) + pcm)


def resample(x, fs_in, fs_out):
    """Polyphase rational resampling (scipy.signal.resample_poly: Kaiser-windowed FIR, up / down by the reduced ratio)
    - the host-side step of dataset/cal_mean_std.py:L52-55 (``librosa.resample``).  librosa's default kernel (soxr_hq) is a
    different low-pass design, so outputs agree to the filters' pass-band accuracy, not bit for bit."""
    if fs_in == fs_out:
        return np.asarray(x, dtype=np.float32)
    from math import gcd
    from scipy.signal import resample_poly
    g = gcd(int(fs_in), int(fs_out))
    return resample_poly(np.asarray(x, dtype=np.float64), int(fs_out) // g, int(fs_in) // g).astype(np.float32)


def load_utterances(paths, fs=16000, allow_resample=False):
    """[float32 CPU tensors] for ragged.enhance_ragged.  The reference's loaders never resample
    (dataset/dataload_nsvae.py:L183: ``sr=None``; its models are trained at 16 kHz), so by default a file at another
    rate is an error; allow_resample=True converts it with ``resample``."""
    import torch
    out = []
    for p in paths:
        x, r = read_wav(p)
        if r != fs:
            if not allow_resample:
                raise ValueError("%s is sampled at %d Hz, the network expects %d Hz" % (p, r, fs))
            x = resample(x, r, fs)
        out.append(torch.from_numpy(x.copy()))
    return out
