"""Device plumbing shared by the host-side wrappers: PyTorch owns memory and streams."""
from __future__ import annotations

import numpy as np
import torch

from . import _lib

_workspaces: dict[tuple[int, str, int], torch.Tensor] = {}
_checked_devices: set[int] = set()


def require_device(device=None) -> torch.device:
    """Resolve `device` to a CUDA device that can run the kernels; raise otherwise (no CPU fallback)."""
    if not torch.cuda.is_available():
        raise RuntimeError(
            "hippomm_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback. "
            "torch.cuda.is_available() is False."
        )
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    if dev.type != "cuda":
        raise RuntimeError(f"hippomm_b200 runs on CUDA devices only, got {dev}")
    if dev.index is None:
        dev = torch.device("cuda", torch.cuda.current_device())
    if dev.index not in _checked_devices:
        with torch.cuda.device(dev):
            _lib.check(_lib.load().hippo_device_check())
        _checked_devices.add(dev.index)
    return dev


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def workspace(nbytes: int, device: torch.device, tag: str = "default") -> torch.Tensor:
    """Grow-only scratch buffer per (device, tag, current stream): calls on different streams never share scratch
    (the C ABI is reentrant across streams only under that condition); torch's caching allocator returns
    >= 512 B aligned blocks and recycles a replaced buffer in stream order."""
    with torch.cuda.device(device):
        key = (device.index, tag, torch.cuda.current_stream().cuda_stream)
    buf = _workspaces.get(key)
    nbytes = max(int(nbytes), 256)
    if buf is None or buf.numel() < nbytes:
        if buf is not None:
            del _workspaces[key]
            del buf
        buf = torch.empty(nbytes, dtype=torch.uint8, device=device)
        _workspaces[key] = buf
    return buf


def release_workspaces() -> None:
    _workspaces.clear()


def ptr(t: torch.Tensor | None) -> int | None:
    return None if t is None else t.data_ptr()


def to_device(a, device: torch.device, dtype: torch.dtype | None = None) -> torch.Tensor:
    """Host array / tensor -> contiguous device tensor (through pinned memory for large NumPy inputs)."""
    if isinstance(a, torch.Tensor):
        t = a.detach()
        if dtype is not None and t.dtype != dtype:
            t = t.to(dtype)
        return t.to(device, non_blocking=True).contiguous()
    arr = np.ascontiguousarray(a)
    if not arr.flags.writeable:          # read-only arrays are only ever read here; torch just cannot express that
        import warnings

        with warnings.catch_warnings():
            warnings.simplefilter("ignore", UserWarning)
            t = torch.from_numpy(arr)
    else:
        t = torch.from_numpy(arr)
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    if t.numel() * t.element_size() >= (1 << 20):
        t = t.pin_memory()
    return t.to(device, non_blocking=True)
