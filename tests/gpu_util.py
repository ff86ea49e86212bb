"""Helpers shared by the -m gpu parity tests: numpy <-> torch CUDA plumbing (PyTorch is only the allocator)."""
import numpy as np

DT = {8: np.uint8, 16: np.uint16, 32: np.uint32, 64: np.uint64}
_SIGNED = {8: np.uint8, 16: np.int16, 32: np.int32, 64: np.int64}


def to_dev(a: np.ndarray):
    import torch

    tb = a.dtype.itemsize * 8
    return torch.from_numpy(np.ascontiguousarray(a).view(_SIGNED[tb])).cuda()


def dev_empty(n: int, tbits: int):
    import torch

    tdt = {8: torch.uint8, 16: torch.int16, 32: torch.int32, 64: torch.int64}[tbits]
    return torch.empty(n, dtype=tdt, device="cuda")


def to_host(t, tbits: int) -> np.ndarray:
    return t.cpu().numpy().view(DT[tbits])


def rand_bytes(rng, n_bytes: int, tbits: int) -> np.ndarray:
    return rng.integers(0, 256, size=n_bytes, dtype=np.uint8).view(DT[tbits])


def mask(w: int) -> int:
    return (1 << w) - 1
