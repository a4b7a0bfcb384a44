"""Host-side helpers shared by the API mirrors."""

from __future__ import annotations

import warnings

import torch

from . import _lib

_warned_cpu = False

GRID_KINDS = {"catmull_rom": 0, "bspline": 1}


def resolve_device(image: torch.Tensor, device) -> torch.device:
    """Reference semantics: ``device=None`` means ``image.device``.  There is no CPU path
    (north star: no CPU fallback): CPU requests run on the current CUDA device and results
    stay on CUDA."""
    global _warned_cpu
    _lib.require_cuda()
    dev = torch.device(device) if device is not None else image.device
    if dev.type != "cuda":
        if not _warned_cpu:
            warnings.warn(
                "torch_motion_correction_b200 has no CPU path: running on the current CUDA device; "
                "results are returned as CUDA tensors.",
                stacklevel=3,
            )
            _warned_cpu = True
        dev = torch.device("cuda", torch.cuda.current_device())
    elif dev.index is None:
        dev = torch.device("cuda", torch.cuda.current_device())
    return dev


def as_f32(t: torch.Tensor, device: torch.device) -> torch.Tensor:
    return t.detach().to(device=device, dtype=torch.float32).contiguous()


def grid_kind(grid_type: str) -> int:
    try:
        return GRID_KINDS[grid_type]
    except KeyError:
        raise ValueError(f"Invalid grid type: {grid_type}. Must be 'catmull_rom' or 'bspline'.") from None


import collections
import threading

_device_cache: "collections.OrderedDict" = collections.OrderedDict()
_device_cache_lock = threading.Lock()
_DEVICE_CACHE_ENTRIES = 256


def cached_device_tensor(key, build, device: torch.device) -> torch.Tensor:
    """Small planning tensors (job lists, schedules, patch centres) are pure functions of the call
    geometry: build them once per device and geometry instead of paying a host-blocking pageable
    H2D copy in the middle of every call.  The tensors must be treated as read-only."""
    full_key = (device.type, device.index, key)
    with _device_cache_lock:
        hit = _device_cache.get(full_key)
        if hit is not None:
            _device_cache.move_to_end(full_key)
            return hit
    hit = build().to(device)
    with _device_cache_lock:
        _device_cache[full_key] = hit
        while len(_device_cache) > _DEVICE_CACHE_ENTRIES:  # least recently used first; tensors in use stay alive
            _device_cache.popitem(last=False)
    return hit


_staging_local = threading.local()  # one ring of pinned buffers per host thread and device
_STAGING_SLOTS = 4


def to_device_async(host_tensor: torch.Tensor, device: torch.device) -> torch.Tensor:
    """Small H2D upload that neither blocks the host nor queues behind a large H2D copy on the copy engine (the next
    movie of a pipelined run): the data is staged in a ring of pinned buffers and copied by a kernel reading the host
    mapping (tmc_upload_pinned).  A slot is reused only after the upload that read it has completed."""
    src = host_tensor.contiguous()
    nbytes = src.numel() * src.element_size()
    if nbytes % 4 != 0 or nbytes == 0:
        staged = torch.empty(src.shape, dtype=src.dtype, pin_memory=True)
        staged.copy_(src)
        return staged.to(device, non_blocking=True)
    key = (device.type, device.index)
    rings = getattr(_staging_local, "rings", None)
    if rings is None:
        rings = _staging_local.rings = {}
    ring = rings.setdefault(key, {"next": 0, "slots": [None] * _STAGING_SLOTS})
    i = ring["next"]
    ring["next"] = (i + 1) % _STAGING_SLOTS
    slot = ring["slots"][i]
    if slot is None or slot[0].numel() < nbytes:
        if slot is not None:
            slot[1].synchronize()
        slot = [torch.empty((max(nbytes, 1 << 16),), dtype=torch.uint8, pin_memory=True), torch.cuda.Event()]
        ring["slots"][i] = slot
    else:
        slot[1].synchronize()  # the previous upload from this slot has run
    buf, event = slot
    buf[:nbytes].view(src.dtype).copy_(src.reshape(-1))
    out = torch.empty(src.shape, dtype=src.dtype, device=device)
    with torch.cuda.device(device):
        _lib.call("tmc_upload_pinned", buf.data_ptr(), out.data_ptr(), nbytes, _lib.stream_ptr(device))
        event.record(torch.cuda.current_stream(device))
    return out
