"""Host -> device copy bandwidth of 1 GiB from (a) torch pinned memory, (b) cudaHostAlloc write-combined memory.
usage: python tools/h2d_probe.py"""
import ctypes
import torch

n = 1 << 30
rt = ctypes.CDLL("libcudart.so.12")
dst = torch.empty(n, dtype=torch.uint8, device="cuda")


def bench(ptr, label, reps=10):
    st = torch.cuda.current_stream()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(2):
        rt.cudaMemcpyAsync(ctypes.c_void_p(dst.data_ptr()), ctypes.c_void_p(ptr), ctypes.c_size_t(n), 1, ctypes.c_void_p(st.cuda_stream))
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        rt.cudaMemcpyAsync(ctypes.c_void_p(dst.data_ptr()), ctypes.c_void_p(ptr), ctypes.c_size_t(n), 1, ctypes.c_void_p(st.cuda_stream))
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"{label:40s} {ms:7.2f} ms per GiB  {n / ms / 1e6:6.1f} GB/s")


a = torch.empty(n, dtype=torch.uint8).pin_memory()
bench(a.data_ptr(), "torch pin_memory (cudaHostAlloc default)")
for flags, name in ((4, "cudaHostAllocWriteCombined"), (1, "cudaHostAllocPortable"), (4 | 1, "WriteCombined | Portable")):
    p = ctypes.c_void_p()
    rc = rt.cudaHostAlloc(ctypes.byref(p), ctypes.c_size_t(n), flags)
    assert rc == 0, rc
    ctypes.memset(p, 1, n)
    bench(p.value, name)
    rt.cudaFreeHost(p)
