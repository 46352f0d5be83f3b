"""Where does a host-buffer layer call spend its time?  Times, at the C2 layer shape, each PCIe leg of
rtkv_compress_layer_host in isolation: contiguous and 2-D (256-byte rows) uploads, downloads, both
directions at once, and the prompt-mass kernel reading the pinned attention tensor in place.
Run on the GPU box: python tools/pcie_probe.py
"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "realtime-kv-cache-compression_b200"))
import torch
from cuda.bindings import runtime as rt

from rtkv import ops
from rtkv._lib import lib

B, H, S, P, n = 8, 32, 4096, 128, 4096
dev = torch.device("cuda:0")
torch.cuda.init()


def timed(fn, reps=3):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


kv_h = torch.empty(B, S, n, dtype=torch.float16).pin_memory()
kv_d = torch.empty(B, S, n, dtype=torch.float16, device=dev)
out_h = torch.empty(B, S, n, dtype=torch.float16).pin_memory()
nb = kv_h.numel() * 2
ms = timed(lambda: kv_d.copy_(kv_h, non_blocking=True))
print(f"H2D contiguous {nb/1e6:.0f} MB: {ms:.2f} ms  {nb/ms/1e6:.1f} GB/s")
ms = timed(lambda: out_h.copy_(kv_d, non_blocking=True))
print(f"D2H contiguous {nb/1e6:.0f} MB: {ms:.2f} ms  {nb/ms/1e6:.1f} GB/s")
s2 = torch.cuda.Stream()


def both():
    s2.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s2):
        out_h.copy_(kv_d, non_blocking=True)
    kv_d.copy_(kv_h, non_blocking=True)
    torch.cuda.current_stream().wait_stream(s2)


ms = timed(both)
print(f"H2D+D2H together 2x{nb/1e6:.0f} MB: {ms:.2f} ms  {2*nb/ms/1e6:.1f} GB/s total")

# the attention slice: rows of P*2 bytes out of rows of S*2 bytes.  Only a [B,H,S,Wa] window of the
# (virtual) [B,H,S,S] tensor is allocated, with the same row pitch trick: pitch = Wa elements.
for Wa in (4096, 512):
    # full-width tensor would be 8.6 GB of pinned memory; use one batch row's worth of heads at full pitch
    Bh = B * H if Wa < 4096 else 32
    a_h = torch.empty(Bh, S, Wa, dtype=torch.float16).pin_memory()
    a_d = torch.empty(Bh, S, P, dtype=torch.float16, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    rows = Bh * S

    def up2d():
        (err,) = rt.cudaMemcpy2DAsync(a_d.data_ptr(), P * 2, a_h.data_ptr(), Wa * 2, P * 2, rows,
                                      rt.cudaMemcpyKind.cudaMemcpyHostToDevice, st)
        assert err == rt.cudaError_t.cudaSuccess, err

    ms = timed(up2d)
    sl = rows * P * 2
    print(f"H2D 2-D {rows} rows x {P*2} B (pitch {Wa*2} B) {sl/1e6:.0f} MB: {ms:.2f} ms  {sl/ms/1e6:.1f} GB/s")

    # the prompt-mass kernel reading the pinned tensor in place (UVA zero-copy)
    Bz = Bh // H
    mass = torch.empty(Bz, S, dtype=torch.float32, device=dev)

    def zero_copy():
        rc = lib.rtkv_prompt_mass(ctypes.c_void_p(a_h.data_ptr()), 1, Bz, H, S, H * S * Wa, S * Wa, Wa, None, P,
                                  ctypes.c_void_p(mass.data_ptr()), ctypes.c_void_p(st))
        assert rc == 0, ops.last_error()

    try:
        ms = timed(zero_copy)
        print(f"K1a on pinned host memory, same slice: {ms:.2f} ms  {sl/ms/1e6:.1f} GB/s")
    except Exception as e:  # noqa: BLE001
        print("zero-copy failed:", e)
    del a_h, a_d
