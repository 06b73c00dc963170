"""Host-side placement of a page-locked point matrix on a multi-socket box (Linux; used by tools/numa_probe.py).

One process drives all GPUs (bartint_init), so one host matrix feeds 8 copy engines at once.  Pages that live on the
socket a GPU hangs off are read by its DMA engine without crossing the inter-socket link; `numa_local_pinned_matrix`
allocates an N_total x d column-major fp64 matrix whose row shard g (the rows GPU g receives) is bound to GPU g's NUMA
node before first touch, then page-locks it with cudaHostRegister.  Everything degrades to an ordinary pinned
allocation when the topology is not visible (sysfs numa_node = -1, mbind refused)."""
from __future__ import annotations

import ctypes
import mmap
import os

import numpy as np

_SYS_MBIND = 237                # x86_64
_MPOL_BIND, _MPOL_INTERLEAVE = 2, 3
_libc = ctypes.CDLL(None, use_errno=True)
_keep = []                      # (mmap object, registered) of every matrix handed out: unregistered at release


def numa_nodes() -> list[int]:
    try:
        return sorted(int(n[4:]) for n in os.listdir("/sys/devices/system/node") if n.startswith("node") and n[4:].isdigit())
    except OSError:
        return []


def gpu_numa_node(device: int) -> int:
    """NUMA node of a CUDA device from sysfs (-1 when unknown)."""
    import torch
    try:
        p = torch.cuda.get_device_properties(device)
        bus = f"{getattr(p, 'pci_domain_id', 0):04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bus}/numa_node") as fh:
            return int(fh.read().strip())
    except Exception:
        return -1


def _mbind(addr: int, length: int, mode: int, nodes: list[int]) -> bool:
    page = mmap.PAGESIZE
    lo = addr // page * page
    hi = (addr + length + page - 1) // page * page
    mask = ctypes.c_ulong(0)
    for n in nodes:
        mask.value |= 1 << n
    rc = _libc.syscall(_SYS_MBIND, ctypes.c_void_p(lo), ctypes.c_ulong(hi - lo), ctypes.c_int(mode), ctypes.byref(mask),
                       ctypes.c_ulong(65), ctypes.c_uint(0))
    return rc == 0


def numa_local_pinned_matrix(n_rows_per_shard: int, d: int, shard_nodes: list[int], policy: str = "local"):
    """(matrix [n_rows_per_shard * len(shard_nodes), d] column-major float64 page-locked, info dict).
    policy: "local" (shard g on shard_nodes[g]), "interleave" (pages round-robin over all nodes), "default"."""
    import torch
    G = len(shard_nodes)
    n_total = n_rows_per_shard * G
    nbytes = n_total * d * 8
    mm = mmap.mmap(-1, nbytes + mmap.PAGESIZE)
    base = ctypes.addressof(ctypes.c_char.from_buffer(mm))
    info = {"policy": policy, "bound": False}
    nodes = numa_nodes()
    if policy == "interleave" and len(nodes) > 1:
        info["bound"] = _mbind(base, nbytes, _MPOL_INTERLEAVE, nodes)
    elif policy == "local" and len(nodes) > 1 and all(n >= 0 for n in shard_nodes):
        ok = True
        for j in range(d):
            for g in range(G):
                off = (j * n_total + g * n_rows_per_shard) * 8
                ok &= _mbind(base + off, n_rows_per_shard * 8, _MPOL_BIND, [shard_nodes[g]])
        info["bound"] = bool(ok)
    X = np.frombuffer(mm, dtype=np.float64, count=n_total * d).reshape((n_total, d), order="F")
    X[:] = 0.0                                            # first touch: the pages are placed now
    info["registered"] = False
    if torch.cuda.is_available():
        info["registered"] = int(torch.cuda.cudart().cudaHostRegister(base, nbytes, 0)) == 0
    _keep.append((mm, base, info["registered"]))
    return X, info


def release_all() -> None:
    import torch
    while _keep:
        mm, base, reg = _keep.pop()
        if reg:
            torch.cuda.cudart().cudaHostUnregister(base)
        try:
            mm.close()
        except BufferError:
            pass                                          # numpy views still alive: the mapping goes with them
