"""DeviceArray: the tensor type the backend hands back to Qibo ("interop").

It owns a CUDA-resident ``torch.Tensor`` (``.tensor`` -- the north-star "state buffer also exposed as a
torch tensor") and presents the duck-type Qibo's callers rely on: a NumPy ``dtype``, ``shape``, ``len``,
``tolist``, indexing, ``__array__``.  Anything that is not one of the overridden hot-path methods sees it
degrade to a host ``numpy.ndarray`` (a device->host copy), which is interop for the reference's inherited
NumPy code paths, not a compute fallback: the hot path itself only ever runs in the CUDA library.
"""

import numpy as np
import torch
from numpy.lib.mixins import NDArrayOperatorsMixin

_NP2TORCH = {
    np.dtype("complex128"): torch.complex128,
    np.dtype("complex64"): torch.complex64,
    np.dtype("float64"): torch.float64,
    np.dtype("float32"): torch.float32,
    np.dtype("int64"): torch.int64,
}
_TORCH2NP = {v: k for k, v in _NP2TORCH.items()}


def torch_dtype(dtype):
    return _NP2TORCH[np.dtype(dtype)]


import os as _os

STREAM_D2H_MIN_BYTES = int(_os.environ.get("QB_STREAM_D2H_MIN_BYTES", 1 << 28))  # below: one plain copy
STREAM_D2H_CHUNK_BYTES = int(_os.environ.get("QB_STREAM_D2H_CHUNK_BYTES", 1 << 27))
_staging = {}


def _stream_to_host(t: torch.Tensor) -> torch.Tensor:
    """Chunked, double-buffered device->host copy of a contiguous CUDA tensor into a new (pageable) CPU tensor."""
    flat = t.reshape(-1).view(torch.uint8)
    out = torch.empty(flat.numel(), dtype=torch.uint8)
    key = (t.device.index, STREAM_D2H_CHUNK_BYTES)
    if key not in _staging:
        _staging[key] = ([torch.empty(STREAM_D2H_CHUNK_BYTES, dtype=torch.uint8, pin_memory=True) for _ in range(2)],
                         torch.cuda.Stream(device=t.device))
    bufs, side = _staging[key]
    side.wait_stream(torch.cuda.current_stream(t.device))  # the state must be complete before it is read
    total = flat.numel()
    starts = list(range(0, total, STREAM_D2H_CHUNK_BYTES))
    events = [None, None]

    def issue(i):
        a = starts[i]
        b = min(total, a + STREAM_D2H_CHUNK_BYTES)
        with torch.cuda.stream(side):
            bufs[i % 2][: b - a].copy_(flat[a:b], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(side)
        events[i % 2] = ev

    issue(0)
    for i, a in enumerate(starts):
        b = min(total, a + STREAM_D2H_CHUNK_BYTES)
        events[i % 2].synchronize()
        if i + 1 < len(starts):
            issue(i + 1)  # (the other staging buffer: its previous contents were consumed one iteration ago)
        out[a:b].copy_(bufs[i % 2][: b - a])
    return out.view(t.dtype).reshape(t.shape)


class DeviceArray(NDArrayOperatorsMixin):
    __array_priority__ = 1000

    def __init__(self, tensor: torch.Tensor):
        if not tensor.is_cuda:
            raise ValueError("DeviceArray wraps CUDA tensors only")
        self.tensor = tensor

    # ---- metadata -------------------------------------------------------------------------
    @property
    def shape(self):
        return tuple(self.tensor.shape)

    @property
    def dtype(self):
        return _TORCH2NP[self.tensor.dtype]

    @property
    def ndim(self):
        return self.tensor.dim()

    @property
    def size(self):
        return self.tensor.numel()

    @property
    def nbytes(self):
        return self.tensor.numel() * self.tensor.element_size()

    def __len__(self):
        return self.tensor.shape[0]

    def data_ptr(self):
        return self.tensor.data_ptr()

    # ---- host interop ---------------------------------------------------------------------
    def numpy(self):
        """Host copy (QuantumState.state(numpy=True) / dump, result.py:69-88, 116-163).  Large states stream through two
        pinned staging buffers: the device->host DMA of chunk i + 1 overlaps the (multi-threaded) host copy of chunk i
        into the result, instead of one pageable cudaMemcpy of the whole state."""
        t = self.tensor.detach()
        if t.numel() * t.element_size() < STREAM_D2H_MIN_BYTES or not t.is_contiguous():
            return t.cpu().numpy()
        return _stream_to_host(t).numpy()

    def __array__(self, dtype=None, copy=None):
        out = self.numpy()
        return out if dtype is None else out.astype(dtype, copy=False)

    def __array_ufunc__(self, ufunc, method, *inputs, **kwargs):
        inputs = tuple(np.asarray(x) if isinstance(x, DeviceArray) else x for x in inputs)
        if "out" in kwargs:
            kwargs["out"] = tuple(np.asarray(x) if isinstance(x, DeviceArray) else x for x in kwargs["out"])
        return getattr(ufunc, method)(*inputs, **kwargs)

    def __array_function__(self, func, types, args, kwargs):
        def conv(x):
            if isinstance(x, DeviceArray):
                return np.asarray(x)
            if isinstance(x, (list, tuple)):
                return type(x)(conv(y) for y in x)
            return x

        return func(*conv(args), **{k: conv(v) for k, v in kwargs.items()})

    def tolist(self):
        return self.numpy().tolist()

    def __getitem__(self, idx):
        if isinstance(idx, DeviceArray):
            idx = idx.tensor
        out = self.tensor[idx]
        if out.dim() == 0:
            return out.cpu().numpy()[()]
        return out.cpu().numpy()

    def __iter__(self):
        return iter(self.numpy())

    def __complex__(self):
        return complex(self.numpy())

    def __float__(self):
        return float(self.numpy())

    def astype(self, dtype, copy=True):
        return self.numpy().astype(dtype, copy=False)

    def copy(self):
        return DeviceArray(self.tensor.clone())

    def reshape(self, *shape):
        if len(shape) == 1 and isinstance(shape[0], (tuple, list)):
            shape = tuple(shape[0])
        return DeviceArray(self.tensor.reshape(shape))

    def ravel(self):
        return DeviceArray(self.tensor.reshape(-1))

    flatten = ravel

    def conj(self):
        return self.numpy().conj()

    @property
    def real(self):
        return self.numpy().real

    @property
    def imag(self):
        return self.numpy().imag

    @property
    def T(self):
        return self.numpy().T

    def __dlpack__(self, *args, **kwargs):
        return self.tensor.__dlpack__(*args, **kwargs)

    def __dlpack_device__(self):
        return self.tensor.__dlpack_device__()

    def __repr__(self):
        return f"DeviceArray(shape={self.shape}, dtype={self.dtype}, device={self.tensor.device})"
