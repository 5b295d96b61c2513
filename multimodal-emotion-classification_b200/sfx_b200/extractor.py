"""Batched, device-resident speech feature extraction (host side of libsfx_b200).

``SpeechFeatureExtractor`` hands PyTorch CUDA tensors (device pointers + the current stream) to the C ABI
of include/sfx.h.  PyTorch is used for device memory, streams and torch.distributed only; all arithmetic
runs in csrc/sfx_kernels.cu.  There is no CPU path: constructing an extractor without a CUDA device raises.

Output layout per clip (reference preprocessing/audio_preprocessing.py:45-46):
    [mfcc_0..n_mfcc-1 | chroma C, C#, ..., B | zcr, spectral_centroid_Hz, spectral_rolloff_Hz, rms]
"""
from __future__ import annotations

import ctypes as C
import threading

import numpy as np
import torch

from . import _lib, tables

N_CHROMA = 12
N_SPECTRAL = 4

_EXTRACTORS: dict = {}
_LOCK = threading.Lock()


class NoCudaDeviceError(RuntimeError):
    """Raised instead of falling back to a CPU implementation."""


class SpeechFeatureExtractor:
    """One instance per (device, sample rate): owns the uploaded tables and one growable workspace per CUDA stream it has
    been called on (the C ABI is re-entrant across streams and threads given distinct workspaces; a workspace holds the
    clip queue and all scratch of a launch, so two launches in flight must never share one)."""

    def __init__(self, device=None, sr: int = 22050):
        self.lib = _lib.load()
        if not torch.cuda.is_available() or self.lib.sfx_device_count() <= 0:
            raise NoCudaDeviceError("sfx_b200 needs a CUDA (sm_100a) device; there is no CPU fallback")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if self.device.type != "cuda":
            raise NoCudaDeviceError(f"sfx_b200 runs on CUDA devices only, got {self.device}")
        self.index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.sr = int(sr)
        self._tables = _lib.make_tables_struct(tables.build_tables(self.sr))
        with torch.cuda.device(self.index):
            _lib.check(self.lib.sfx_init_tables(self.index, C.byref(self._tables)))
        self._ws = {}                  # cuda_stream handle -> uint8 workspace tensor
        self._ws_lock = threading.Lock()
        self.launches = 0

    # ------------------------------------------------------------------ workspace
    def _workspace(self, max_samples: int, batch: int, stream: torch.cuda.Stream) -> torch.Tensor:
        """The workspace of `stream`, grown to what a batch of `batch` clips of `max_samples` samples needs (re-queried
        every call: the requirement depends on the pipeline mode and on the table sets uploaded so far).  Kernels on one
        stream run in order, so reusing the stream's workspace for its next launch is safe; a buffer that is replaced is
        handed back to the caching allocator only after `record_stream`, i.e. once the stream's queued work is done."""
        nbytes = self.lib.sfx_workspace_bytes_batch(self.index, int(max_samples), int(batch))
        if nbytes == 0:
            raise _lib.SfxError(-1, self.lib.sfx_last_error().decode())
        key = stream.cuda_stream
        with self._ws_lock:
            ws = self._ws.get(key)
            if ws is None or nbytes > ws.numel():
                if ws is not None:
                    ws.record_stream(stream)
                with torch.cuda.stream(stream):
                    ws = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
                self._ws[key] = ws
        return ws

    # ------------------------------------------------------------------ device path
    def extract(self, waves: torch.Tensor, lengths: torch.Tensor | None = None, n_samples: int | None = None,
                n_mfcc: int = 40, out: torch.Tensor | None = None, debug: dict | None = None) -> torch.Tensor:
        """waves: cuda float32 [B, L] (row-contiguous); lengths: cuda int32 [B] or None; n_samples: common clip
        length when lengths is None (default L; smaller = the pad/trim of load_audio).  Returns cuda float32
        [B, n_mfcc+16].  Stream-ordered on torch's current stream; no host sync."""
        if waves.device != self.device or waves.dtype != torch.float32 or waves.dim() != 2:
            raise ValueError("waves must be a 2-D float32 tensor on the extractor's device")
        if waves.stride(1) != 1:
            raise ValueError("waves rows must be contiguous")
        B, L = waves.shape
        n_default = L if n_samples is None else int(n_samples)
        if lengths is not None:
            if lengths.device != self.device or lengths.dtype != torch.int32 or lengths.shape != (B,):
                raise ValueError("lengths must be int32 [B] on the extractor's device")
            lengths = lengths.contiguous()
            max_samples = L
        else:
            if not 1 <= n_default <= L:
                raise ValueError("n_samples must be in [1, L]")
            max_samples = n_default
        width = n_mfcc + N_CHROMA + N_SPECTRAL
        if out is None:
            out = torch.empty((B, width), dtype=torch.float32, device=self.device)
        elif out.shape != (B, width) or out.dtype != torch.float32 or out.device != self.device or out.stride(1) != 1:
            raise ValueError("out must be float32 [B, n_mfcc+16] on the extractor's device")
        if B == 0:
            return out
        cur = torch.cuda.current_stream(self.device)
        ws = self._workspace(max_samples, B, cur)
        stream = cur.cuda_stream
        args = (self.index, self.sr, waves.data_ptr(), waves.stride(0), lengths.data_ptr() if lengths is not None else None,
                n_default, max_samples, B, n_mfcc, out.data_ptr(), out.stride(0), ws.data_ptr(), ws.numel(),
                C.c_void_p(stream))
        with torch.cuda.device(self.index):
            if debug is None:
                _lib.check(self.lib.sfx_extract(*args))
            else:
                T = 1 + max_samples // tables.HOP
                debug["P"] = torch.zeros((B, T, tables.P_STRIDE), dtype=torch.float32, device=self.device)
                debug["logmel"] = torch.zeros((B, T, tables.N_MELS), dtype=torch.float32, device=self.device)
                debug["frame_feat"] = torch.zeros((B, T, 4), dtype=torch.float32, device=self.device)
                debug["clip_info"] = torch.zeros((B, 8), dtype=torch.float32, device=self.device)
                d = _lib.DebugOut(P=debug["P"].data_ptr(), logmel=debug["logmel"].data_ptr(),
                                  frame_feat=debug["frame_feat"].data_ptr(), clip_info=debug["clip_info"].data_ptr(),
                                  T_dbg=T)
                _lib.check(self.lib.sfx_extract_debug(*args, C.byref(d)))
        self.launches += self.lib.sfx_launches_per_extract()
        return out

    # ------------------------------------------------------------------ host path (the reference-facing one)
    def extract_host(self, waves: np.ndarray, lengths: np.ndarray | None = None, n_samples: int | None = None,
                     n_mfcc: int = 40, out: np.ndarray | None = None, chunk_clips: int = 0) -> np.ndarray:
        """waves: host float32 [B, L] (numpy, or anything exposing a C-contiguous-row buffer; pinned memory is
        copied without staging), or host int16 [B, L] = 16-bit PCM as stored in a WAV file at this extractor's sample
        rate (converted on the device as soundfile does, x / 32768; half the bytes over PCIe); returns host float32
        [B, n_mfcc+16].  H2D, kernel and D2H are chunk-pipelined inside the C library; the call returns when the rows
        are in ``out``."""
        waves = np.asarray(waves)
        if waves.dtype not in (np.float32, np.int16) or waves.ndim != 2 or waves.strides[1] != waves.itemsize:
            raise ValueError("waves must be a 2-D float32 (or int16 PCM) array with contiguous rows")
        B, L = waves.shape
        n_default = L if n_samples is None else int(n_samples)
        width = n_mfcc + N_CHROMA + N_SPECTRAL
        if out is None:
            out = np.empty((B, width), dtype=np.float32)
        if out.shape != (B, width) or out.dtype != np.float32 or out.strides[1] != 4:
            raise ValueError("out must be float32 [B, n_mfcc+16]")
        lp = None
        if lengths is not None:
            lengths = np.ascontiguousarray(lengths, dtype=np.int32)
            if lengths.shape != (B,):
                raise ValueError("lengths must have shape [B]")
            lp = lengths.ctypes.data
        if B == 0:
            return out
        fn = self.lib.sfx_extract_host if waves.dtype == np.float32 else self.lib.sfx_extract_host_pcm16
        with torch.cuda.device(self.index):
            rc = fn(self.index, self.sr, waves.ctypes.data, waves.strides[0] // waves.itemsize, lp, n_default,
                    B, n_mfcc, out.ctypes.data, out.strides[0] // 4, int(chunk_clips))
        # SFX_ERR_BAD_CLIP after the rows were delivered = some clip held a non-finite sample: its row is NaN, which is
        # what the callers look at (preprocessing.audio_preprocessing raises ParameterError for it, like librosa)
        # (a bad *length* is rejected before any work with the same code and does raise)
        if not (rc == _lib.ERR_BAD_CLIP and b"non-finite" in self.lib.sfx_last_error()):
            _lib.check(rc)
        self.launches += self.lib.sfx_launches_per_extract()      # summed over the call's chunks by the library
        return out

    def set_pipeline(self, mode) -> None:
        """Process-wide pipeline choice: 'auto' | 'fused' | 'split' | 'stream' (or 0..3), see include/sfx.h."""
        _lib.check(self.lib.sfx_set_pipeline(_lib.PIPELINES.get(mode, mode)))

    def preprocess_pcm16(self, pcm: np.ndarray, frames: np.ndarray | None, native_sr: int, channels: int = 1,
                         duration: float = 3, n_mfcc: int = 40, out: np.ndarray | None = None,
                         chunk_clips: int = 0) -> np.ndarray:
        """load_audio + feature extraction for raw 16-bit PCM (scope row f3): ``pcm`` is host int16 [B, >= frames*channels]
        (interleaved when stereo) at ``native_sr``; ``frames`` [B] are the frames each file holds (None: the row length).
        Per clip: first round(native_sr*duration) frames, x/32768, channel mean, scipy-equivalent polyphase resampling
        to this extractor's rate on the device (bit-identical to the host path of load_audio), zero pad / trim to
        sr*duration, features.  Returns host float32 [B, n_mfcc+16]."""
        from . import resample
        pcm = np.asarray(pcm)
        if pcm.dtype != np.int16 or pcm.ndim != 2 or pcm.strides[1] != 2:
            raise ValueError("pcm must be a 2-D int16 array with contiguous rows")
        B, L = pcm.shape
        if channels not in (1, 2):
            raise ValueError("channels must be 1 or 2")
        limit = int(round(native_sr * duration))
        if frames is None:
            frames = np.full(B, L // channels, dtype=np.int64)
        frames = np.minimum(np.asarray(frames, dtype=np.int64), limit).astype(np.int32)
        if frames.shape != (B,) or (frames * channels > L).any():
            raise ValueError("frames must have shape [B] and fit the rows")
        width = n_mfcc + N_CHROMA + N_SPECTRAL
        if out is None:
            out = np.empty((B, width), dtype=np.float32)
        if out.shape != (B, width) or out.dtype != np.float32 or out.strides[1] != 4:
            raise ValueError("out must be float32 [B, n_mfcc+16]")
        if B == 0:
            return out
        flt = resample.resample_filter(native_sr, self.sr)
        rs = _lib.ResamplerHost(up=flt["up"], down=flt["down"], n_taps=len(flt["taps"]), n_pre_remove=flt["n_pre_remove"],
                                taps=flt["taps"].ctypes.data)
        n_target = int(self.sr * duration)
        with torch.cuda.device(self.index):
            rc = self.lib.sfx_preprocess_host_pcm16(self.index, self.sr, rs, pcm.ctypes.data, pcm.strides[0] // 2, channels,
                                                    frames.ctypes.data, 0, n_target, B, n_mfcc, out.ctypes.data,
                                                    out.strides[0] // 4, int(chunk_clips))
        if rc:
            raise _lib.SfxError(rc, self.lib.sfx_frontend_last_error().decode())
        return out


def get_extractor(device=None, sr: int = 22050) -> SpeechFeatureExtractor:
    """Process-wide cache: one extractor per (device index, sr)."""
    if not torch.cuda.is_available():
        raise NoCudaDeviceError("sfx_b200 needs a CUDA (sm_100a) device; there is no CPU fallback")
    idx = None if device is None else torch.device(device).index      # "cuda" without an index = the current device
    if idx is None:
        idx = torch.cuda.current_device()
    key = (idx, int(sr))
    with _LOCK:
        ex = _EXTRACTORS.get(key)
        if ex is None:
            ex = _EXTRACTORS[key] = SpeechFeatureExtractor(torch.device("cuda", idx), sr)
    return ex


def extract_features_batch(waveforms, lengths=None, sr: int = 22050, n_mfcc: int = 40, n_samples=None):
    """Batched entry point: cuda float32 [B, L] (+ int32 lengths) -> cuda float32 [B, n_mfcc+16];
    host numpy float32 [B, L] -> host numpy [B, n_mfcc+16] through the chunk-pipelined host path."""
    if isinstance(waveforms, torch.Tensor) and waveforms.is_cuda:
        ex = get_extractor(waveforms.device, sr)
        return ex.extract(waveforms, lengths, n_samples=n_samples, n_mfcc=n_mfcc)
    if isinstance(waveforms, torch.Tensor):
        waveforms = waveforms.numpy()
    ex = get_extractor(None, sr)
    return ex.extract_host(np.asarray(waveforms, dtype=np.float32), lengths, n_samples=n_samples, n_mfcc=n_mfcc)
