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
        return self.tensor.detach().cpu().numpy()

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
