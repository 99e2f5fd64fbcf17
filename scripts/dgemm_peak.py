"""Measure the FP64 GEMM-class peak of the box (cuBLAS dgemm via torch) -> roofline denominator."""
import json, sys, torch
n = 8192
a = torch.randn(n, n, dtype=torch.float64, device="cuda"); b = torch.randn(n, n, dtype=torch.float64, device="cuda")
for _ in range(3): c = a @ b
torch.cuda.synchronize()
best = 0
for _ in range(5):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); c = a @ b; e1.record(); torch.cuda.synchronize()
    best = max(best, 2 * n**3 / (e0.elapsed_time(e1) * 1e-3) * 1e-12)
# sustained
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): c = a @ b
e1.record(); torch.cuda.synchronize()
sus = 20 * 2 * n**3 / (e0.elapsed_time(e1) * 1e-3) * 1e-12
print(json.dumps({"fp64_dgemm_tflops_burst": best, "fp64_dgemm_tflops_sustained": sus, "n": n}))
