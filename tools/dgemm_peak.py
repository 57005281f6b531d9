"""cuBLAS DGEMM peak (roofline denominator for the fp64 path), same protocol as MEASURED_PEAKS.json:
best of 10 (burst) and back-to-back for ~4 s (sustained).  Prints one JSON line."""
import json, time, torch

def main():
    n = 8192
    a = torch.randn(n, n, dtype=torch.float64, device="cuda")
    b = torch.randn(n, n, dtype=torch.float64, device="cuda")
    c = torch.empty_like(a)
    for _ in range(3):
        torch.matmul(a, b, out=c)
    torch.cuda.synchronize()
    fl = 2.0 * n ** 3
    best = 1e30
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); torch.matmul(a, b, out=c); e1.record(); e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = max(1, int(4000.0 / best))
    e0.record()
    for _ in range(reps):
        torch.matmul(a, b, out=c)
    e1.record(); e1.synchronize()
    sus = e0.elapsed_time(e1) / reps
    print(json.dumps({"dgemm_n": n, "fp64_tflops": fl / best * 1e-9, "fp64_tflops_sustained": fl / sus * 1e-9,
                      "gpu_name": torch.cuda.get_device_name(0), "how": "torch.matmul fp64 8192^3 (cuBLAS), best of 10 and back-to-back ~4 s"}))

if __name__ == "__main__":
    main()
