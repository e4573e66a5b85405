"""pycuda.driver.mem_alloc / memcpy_htod over torch CUDA tensors."""
import numpy as np
import torch


def mem_alloc(nbytes):
    return torch.empty(int(nbytes), dtype=torch.uint8, device="cuda")


def memcpy_htod(dst, src):
    a = np.ascontiguousarray(src).view(np.uint8).reshape(-1)
    dst[:a.size].copy_(torch.from_numpy(a), non_blocking=True)
