"""Batched WAV ingest: a list of files -> ONE host buffer of int16 samples (pinned when a GPU is present) + offsets.

The reference reads one file per call (``utils/tools.py:45-47`` -> ``scipy.io.wavfile.read``, looped in
``GMM_UBM.py:36-46``) and hands a Python list of arrays to the feature loop.  Here the RIFF headers of all files are
parsed first (a few hundred bytes each), one staging buffer of the total size is allocated, and worker threads
``readinto`` the sample payloads straight into their slices of it -- no per-file array, no concatenation; the front-end
then needs a single host-to-device copy.  Files keep their first channel (what ``MFCC_DTW.py:141-144`` does with stereo
input).  Only 16-bit PCM is decoded here (the format of the reference's data set and of every caller on the path);
anything else raises ``ValueError`` and :func:`speech_signal_processing_b200.ubm.load_data` falls back to scipy.
"""
from __future__ import annotations

import os
import struct
from concurrent.futures import ThreadPoolExecutor
from dataclasses import dataclass

import numpy as np


@dataclass
class WavInfo:
    path: str
    rate: int
    channels: int
    data_offset: int   # byte offset of the sample payload in the file
    n_frames: int      # samples per channel


@dataclass
class PcmBatch:
    """``pcm``: 1-D int16 (a pinned ``torch`` tensor on a machine with a GPU, else a numpy array), utterance i =
    ``pcm[sample_offsets[i] : sample_offsets[i + 1]]``."""
    pcm: object
    sample_offsets: np.ndarray
    rates: np.ndarray
    paths: list

    def __len__(self):
        return len(self.paths)

    def host_array(self) -> np.ndarray:
        return self.pcm if isinstance(self.pcm, np.ndarray) else self.pcm.numpy()

    def utterances(self) -> list:
        """The utterances as numpy views of the staging buffer (what ``load_data`` returns as ``x``)."""
        h, o = self.host_array(), self.sample_offsets
        return [h[o[i] : o[i + 1]] for i in range(len(self.paths))]


def parse_header(path: str) -> WavInfo:
    """Walk the RIFF chunks of ``path`` up to the ``data`` chunk (scipy.io.wavfile conventions: chunks are word aligned, a
    data size of 0 or beyond the end of the file means "the rest of the file")."""
    size = os.path.getsize(path)
    with open(path, "rb") as f:
        head = f.read(12)
        if len(head) < 12 or head[:4] != b"RIFF" or head[8:12] != b"WAVE":
            raise ValueError(f"{path}: not a little-endian RIFF/WAVE file")
        fmt = None
        while True:
            hdr = f.read(8)
            if len(hdr) < 8:
                raise ValueError(f"{path}: no data chunk")
            cid, csize = hdr[:4], struct.unpack("<I", hdr[4:])[0]
            if cid == b"fmt ":
                body = f.read(csize + (csize & 1))
                if csize < 16:
                    raise ValueError(f"{path}: short fmt chunk")
                tag, channels, rate, _, align, bits = struct.unpack("<HHIIHH", body[:16])
                if tag == 0xFFFE and csize >= 26:           # WAVE_FORMAT_EXTENSIBLE: the real tag leads the sub-format GUID
                    tag = struct.unpack("<H", body[24:26])[0]
                fmt = (tag, channels, rate, align, bits)
            elif cid == b"data":
                if fmt is None:
                    raise ValueError(f"{path}: data chunk before fmt chunk")
                tag, channels, rate, align, bits = fmt
                if tag != 1 or bits != 16 or channels < 1 or align != 2 * channels:
                    raise ValueError(f"{path}: only 16-bit PCM is decoded here (format tag {tag}, {bits} bits, {channels} channels)")
                off = f.tell()
                nbytes = csize if 0 < csize <= size - off else size - off
                return WavInfo(path, rate, channels, off, nbytes // align)
            else:
                f.seek(csize + (csize & 1), os.SEEK_CUR)


def _fill(info: WavInfo, dst: np.ndarray) -> None:
    """Read the payload of one file into its slice ``dst`` (int16, len == n_frames) of the staging buffer."""
    if info.n_frames == 0:
        return
    with open(info.path, "rb", buffering=0) as f:
        f.seek(info.data_offset)
        if info.channels == 1:
            view = memoryview(dst).cast("B")
            got = 0
            while got < len(view):                      # readinto may return short counts on some file systems
                n = f.readinto(view[got:])
                if not n:
                    raise ValueError(f"{info.path}: truncated data chunk")
                got += n
        else:
            raw = np.empty(info.n_frames * info.channels, dtype="<i2")
            view = memoryview(raw).cast("B")
            got = 0
            while got < len(view):
                n = f.readinto(view[got:])
                if not n:
                    raise ValueError(f"{info.path}: truncated data chunk")
                got += n
            dst[:] = raw.reshape(info.n_frames, info.channels)[:, 0]


def read_wav_batch(paths, threads: int | None = None, pin: bool | None = None) -> PcmBatch:
    """Decode ``paths`` (16-bit PCM WAV files) into one :class:`PcmBatch`.  ``pin``: allocate the staging buffer as pinned
    host memory (default: when CUDA is available), so that the front-end's upload is one asynchronous DMA."""
    paths = [os.fspath(p) for p in paths]
    workers = threads or min(32, 2 * (os.cpu_count() or 4))
    with ThreadPoolExecutor(max_workers=max(1, workers)) as pool:
        infos = list(pool.map(parse_header, paths))
        lens = np.array([i.n_frames for i in infos], dtype=np.int64)
        offs = np.zeros(len(infos) + 1, dtype=np.int64)
        np.cumsum(lens, out=offs[1:])
        total = int(offs[-1])
        if pin is None:
            try:
                import torch

                pin = torch.cuda.is_available()
            except ImportError:
                pin = False
        if pin:
            import torch

            pcm = torch.empty(total, dtype=torch.int16, pin_memory=True)
            host = pcm.numpy()
        else:
            pcm = host = np.empty(total, dtype=np.int16)
        if host.dtype.byteorder == ">":   # (never on the platforms this runs on; the files are little-endian)
            raise ValueError("big-endian host")
        list(pool.map(lambda io: _fill(io[0], host[io[1] : io[2]]), zip(infos, offs[:-1], offs[1:])))
    return PcmBatch(pcm, offs, np.array([i.rate for i in infos], dtype=np.int32), paths)
