"""cuBLAS DGEMM throughput on this box (library reference point for the FP64 roofline).
Measurement tool only -- torch is not part of the product path."""
import json, torch
n = 8192
a = torch.randn(n, n, dtype=torch.float64, device="cuda")
b = torch.randn(n, n, dtype=torch.float64, device="cuda")
for _ in range(3):
    c = a @ b
torch.cuda.synchronize()
best = 1e9
for _ in range(8):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); c = a @ b; e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
print(json.dumps({"cublas_dgemm_tflops": 2 * n ** 3 / (best * 1e-3) / 1e12, "n": n, "ms": best}))
