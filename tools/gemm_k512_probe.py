"""What the library GEMM gives to K1's contraction shape: torch.matmul (cuBLAS) bf16, [rows x D] @ [D x M] with the
product written to HBM, D = 512 / 768 / 1024 -- against MEASURED_PEAKS.json's 8192^3 figure the roofline fraction is
quoted on.  A short contraction dimension pays one accumulator drain per D flops per element; 8192^3 pays it once per
8192.  K1 reads its accumulators for the same reason (every score is compared with the row's threshold) but writes
no product.  Prints one JSON line per D."""
import json, torch

M = 1_000_000
for D, rows in ((512, 8192), (768, 8192), (1024, 8192)):
    a = torch.randn(rows, D, device="cuda", dtype=torch.bfloat16)
    b = torch.randn(M, D, device="cuda", dtype=torch.bfloat16)
    out = torch.empty(rows, M, device="cuda", dtype=torch.bfloat16)
    for _ in range(3):
        torch.matmul(a, b.T, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 20  # ~ 1 s back to back: the sustained, power-capped regime
    e0.record()
    for _ in range(reps):
        torch.matmul(a, b.T, out=out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(json.dumps({"D": D, "rows": rows, "M": M, "ms": round(ms, 3), "tflops": round(2.0 * rows * M * D / ms / 1e9, 1),
                      "product_bytes_written": rows * M * 2}), flush=True)
    del a, b, out
