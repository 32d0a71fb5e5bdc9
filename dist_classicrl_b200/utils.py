"""Mixed-radix coding of MultiDiscrete vectors (drop-in for the reference's ``utils.py:12-139``).

Same functions and argument order as the reference.  The *batch* functions (``encode_multi_discretes``,
``decode_to_multi_discretes``) run on the GPU (``qe_radix_encode`` / ``qe_radix_decode``): NumPy arrays are copied
to the device and back, CUDA tensors stay there.  The scalar functions (one vector, ``dims`` multiply-adds) and
``compute_radix`` (``dims`` entries, set-up time) are plain integer arithmetic on the host.
"""

from __future__ import annotations

import ctypes as C
from typing import Any

import numpy as np

from dist_classicrl_b200 import capi


def compute_radix(nvec):
    """``radix[d] = prod(nvec[d+1:])`` as int32 (UTL:12-29)."""
    nvec = np.asarray(nvec)
    radix = np.empty_like(nvec)
    radix[1:] = nvec[::-1][:-1]
    radix[0] = 1
    return np.cumprod(radix, dtype=np.int32)[::-1]


def encode_multi_discrete(multidiscrete_vector, radix) -> int:
    """One vector -> its index (UTL:32-48)."""
    return int(np.dot(np.asarray(multidiscrete_vector), np.asarray(radix)))


def decode_to_multi_discrete(nvec, index: int, radix):
    """One index -> its vector (UTL:72-92)."""
    return (index // np.asarray(radix)) % np.asarray(nvec)


def _one_row(a, what: str) -> np.ndarray:
    """The single [dims] row a broadcastable radix / nvec argument stands for."""
    a = np.asarray(a)
    if a.ndim == 2:
        if a.shape[0] > 1 and not (a == a[0]).all():
            raise ValueError(f"{what}: per-row values are not supported (every row must use the same {what})")
        a = a[0]
    if a.ndim != 1:
        raise ValueError(f"{what} must have shape [dims] (or [n, dims] with identical rows)")
    return np.ascontiguousarray(a, dtype=np.int64)


def _is_cuda_tensor(x) -> bool:
    return type(x).__module__.startswith("torch") and getattr(x, "is_cuda", False)


def _device():
    import torch

    if not torch.cuda.is_available():
        raise RuntimeError("the batch radix functions run on the GPU (there is no CPU fallback)")
    return torch


def encode_multi_discretes(multidiscrete_vectors, radixes):
    """``[n, dims]`` vectors -> ``int64 [n]`` indices, ``sum(vectors * radixes, axis=1)`` (UTL:51-69)."""
    torch = _device()
    radix = _one_row(radixes, "radixes")
    on_device = _is_cuda_tensor(multidiscrete_vectors)
    if on_device:
        v = multidiscrete_vectors.to(torch.int32).contiguous()
    else:
        host = np.asarray(multidiscrete_vectors)
        if host.ndim != 2:
            raise ValueError("multidiscrete_vectors must have shape [n, dims]")
        v = torch.from_numpy(np.ascontiguousarray(host, dtype=np.int32)).cuda()
    if v.ndim != 2 or v.shape[1] != radix.shape[0]:
        raise ValueError(f"operands could not be broadcast together: vectors {tuple(v.shape)} vs radixes {radix.shape}")
    out = torch.empty(v.shape[0], dtype=torch.int64, device=v.device)
    with torch.cuda.device(v.device):
        capi.check(capi.lib().qe_radix_encode(v.data_ptr(), radix.ctypes.data_as(C.c_void_p), radix.shape[0], out.data_ptr(), v.shape[0],
                                             C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    return out if on_device else out.cpu().numpy()


def decode_to_multi_discretes(nvecs, indices, radixes):
    """``[n]`` (or ``[n, 1]``) indices -> ``int32 [n, dims]`` vectors, ``(indices // radixes) % nvecs`` (UTL:95-115)."""
    torch = _device()
    radix, nvec = _one_row(radixes, "radixes"), _one_row(nvecs, "nvecs")
    if radix.shape != nvec.shape:
        raise ValueError("nvecs and radixes must have the same length")
    if (radix == 0).any() or (nvec == 0).any():
        raise ZeroDivisionError("integer division or modulo by zero")
    on_device = _is_cuda_tensor(indices)
    if on_device:
        idx = indices.reshape(-1).to(torch.int64).contiguous()
    else:
        idx = torch.from_numpy(np.ascontiguousarray(np.asarray(indices).reshape(-1), dtype=np.int64)).cuda()
    out = torch.empty((idx.shape[0], radix.shape[0]), dtype=torch.int32, device=idx.device)
    with torch.cuda.device(idx.device):
        capi.check(capi.lib().qe_radix_decode(idx.data_ptr(), nvec.ctypes.data_as(C.c_void_p), radix.ctypes.data_as(C.c_void_p), radix.shape[0],
                                             out.data_ptr(), idx.shape[0], C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    return out if on_device else out.cpu().numpy()


def _make_dummy_vec_env(n_envs: int, env_class: type, env_kwargs: dict[str, Any]):
    """``n_envs`` instances of ``env_class`` behind a :class:`DummyVecWrapper` (UTL:118-139)."""
    from dist_classicrl_b200.wrappers.dummy_vec_wrapper import DummyVecWrapper

    return DummyVecWrapper([env_class(**env_kwargs) for _ in range(n_envs)])
