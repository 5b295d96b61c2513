"""Device-resident StandardScaler + speech DNN forward (scope row f1) -- host side of csrc/sfx_dnn.cu.

Mirrors what the reference's inference/speech_inference.py does per clip (scaler.transform -> model.predict -> arg-max,
and the layers[-3] 64-d tap of extract_features) for a whole batch of feature rows that never leave the GPU.
Weights come as a dict of numpy arrays (the layout of oracle/speech_dnn.random_model; a converter from a Keras .h5
needs h5py/TF, which are not available here): kernel{i} [in,out], bias{i}, gamma/beta/mean/var{i}, scaler_mean/scale.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib


class SpeechDNN:
    def __init__(self, weights: dict, device=None, bn_eps: float = 1e-3):
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            from .extractor import NoCudaDeviceError
            raise NoCudaDeviceError("sfx_b200 needs a CUDA (sm_100a) device; there is no CPU fallback")
        dev = torch.device("cuda") if device is None else torch.device(device)
        if dev.type != "cuda":
            from .extractor import NoCudaDeviceError
            raise NoCudaDeviceError(f"sfx_b200 runs on CUDA devices only, got {dev}")
        self.index = dev.index if dev.index is not None else torch.cuda.current_device()   # "cuda" = the current device
        self.device = torch.device("cuda", self.index)
        widths = [int(w) for w in weights["widths"]]
        n = len(widths) - 1
        self.widths = widths
        keep = []

        def arr(name, dtype=np.float32):
            a = np.ascontiguousarray(weights[name], dtype=dtype)
            keep.append(a)
            return a.ctypes.data

        def ptr_array(names):
            vals = (C.c_void_p * len(names))(*[arr(nm) for nm in names])
            keep.append(vals)
            return C.cast(vals, C.c_void_p)

        dims = np.array(widths, dtype=np.int32)
        keep.append(dims)
        has_bn = all(f"gamma{i}" in weights for i in range(n - 1))
        has_scaler = "scaler_mean" in weights and "scaler_scale" in weights
        h = _lib.DnnHost(
            n_layers=n, dims=dims.ctypes.data,
            kernel=ptr_array([f"kernel{i}" for i in range(n)]), bias=ptr_array([f"bias{i}" for i in range(n)]),
            bn_gamma=ptr_array([f"gamma{i}" for i in range(n - 1)]) if has_bn else None,
            bn_beta=ptr_array([f"beta{i}" for i in range(n - 1)]) if has_bn else None,
            bn_mean=ptr_array([f"mean{i}" for i in range(n - 1)]) if has_bn else None,
            bn_var=ptr_array([f"var{i}" for i in range(n - 1)]) if has_bn else None,
            bn_eps=float(bn_eps),
            scaler_mean=arr("scaler_mean", np.float64) if has_scaler else None,
            scaler_scale=arr("scaler_scale", np.float64) if has_scaler else None)
        self._handle = C.c_void_p()
        with torch.cuda.device(self.index):
            rc = self.lib.sfx_dnn_create(self.index, C.byref(h), C.byref(self._handle))
        if rc < 0:
            raise _lib.SfxError(rc, self.lib.sfx_dnn_last_error().decode())
        self._ws = None
        self.launches = 0

    def __del__(self):
        try:
            if getattr(self, "_handle", None):
                self.lib.sfx_dnn_destroy(self._handle)
        except Exception:
            pass

    def forward(self, feats: torch.Tensor):
        """feats: cuda float32 [B, widths[0]] (e.g. SpeechFeatureExtractor.extract output) -> (probs [B, n_out], tap [B, 64])."""
        if feats.device != self.device or feats.dtype != torch.float32 or feats.dim() != 2 or feats.shape[1] != self.widths[0]:
            raise ValueError(f"feats must be cuda float32 [B, {self.widths[0]}] on {self.device}")
        if feats.stride(1) != 1:
            feats = feats.contiguous()
        B = feats.shape[0]
        probs = torch.empty((B, self.widths[-1]), dtype=torch.float32, device=self.device)
        tap = torch.empty((B, self.widths[-2]), dtype=torch.float32, device=self.device)
        if B == 0:
            return probs, tap
        need = self.lib.sfx_dnn_workspace_bytes(self._handle, B)
        if self._ws is None or self._ws.numel() < need:
            self._ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        with torch.cuda.device(self.index):
            rc = self.lib.sfx_dnn_forward(self._handle, feats.data_ptr(), feats.stride(0), B, probs.data_ptr(), probs.stride(0),
                                          tap.data_ptr(), tap.stride(0), self._ws.data_ptr(), self._ws.numel(),
                                          C.c_void_p(stream))
        if rc < 0:
            raise _lib.SfxError(rc, self.lib.sfx_dnn_last_error().decode())
        self.launches += self.lib.sfx_dnn_launches_per_forward(self._handle)
        return probs, tap
