"""Measure the FP64 / complex128 GEMM denominators (cuBLAS through torch.matmul) and the HBM copy rate.

Writes one JSON object to stdout; committed under profiles/ as the FP64 roofline denominator
(MEASURED_PEAKS.json only carries bf16).
"""
import json
import time
import torch


def _time(fn, reps=10):
    fn(); torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def _sustained(fn, seconds=3.0):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 0
    t0 = time.time()
    e0.record()
    while time.time() - t0 < seconds:
        for _ in range(5):
            fn()
        n += 5
        torch.cuda.synchronize()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


out = {"gpu": torch.cuda.get_device_name(0), "torch": torch.__version__}
for n in (4096, 8192):
    a = torch.randn(n, n, dtype=torch.float64, device="cuda"); b = torch.randn(n, n, dtype=torch.float64, device="cuda")
    ms = _time(lambda: a @ b)
    out[f"dgemm_{n}_tflops_burst"] = 2 * n ** 3 / ms * 1e-9
    if n == 8192:
        ms = _sustained(lambda: a @ b)
        out[f"dgemm_{n}_tflops_sustained"] = 2 * n ** 3 / ms * 1e-9
    del a, b
for n in (4096,):
    a = torch.randn(n, n, dtype=torch.complex128, device="cuda"); b = torch.randn(n, n, dtype=torch.complex128, device="cuda")
    ms = _time(lambda: a @ b)
    out[f"zgemm_{n}_tflops_burst"] = 8 * n ** 3 / ms * 1e-9
    ms = _sustained(lambda: a @ b)
    out[f"zgemm_{n}_tflops_sustained"] = 8 * n ** 3 / ms * 1e-9
    del a, b
# skinny shapes typical of sector GEMMs
for (m, k, nn) in ((1853, 163, 902), (4240, 1089, 5329), (1024, 319, 1024), (926, 408, 1334)):
    a = torch.randn(m, k, dtype=torch.float64, device="cuda"); b = torch.randn(k, nn, dtype=torch.float64, device="cuda")
    ms = _time(lambda: a @ b)
    out[f"dgemm_{m}x{k}x{nn}_tflops"] = 2 * m * k * nn / ms * 1e-9
x = torch.empty(1 << 28, dtype=torch.float64, device="cuda"); y = torch.empty_like(x)
ms = _time(lambda: y.copy_(x))
out["hbm_copy_gbs"] = 2 * x.numel() * 8 / ms * 1e-6
print(json.dumps(out))
