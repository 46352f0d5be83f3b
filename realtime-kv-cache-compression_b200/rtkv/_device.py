"""Device placement helper: the reference accepts tensors on any device; rtkv computes
on the GPU only.  Host tensors are staged to the current CUDA device and results are
returned on the caller's device.  Without a GPU every call raises."""
import torch

from .ops import require_cuda


def to_cuda(t):
    require_cuda()
    if t is None or t.is_cuda:
        return t
    return t.to("cuda", non_blocking=False)


def like_input(result, ref_device):
    if result is None or result.device == ref_device:
        return result
    return result.to(ref_device)
