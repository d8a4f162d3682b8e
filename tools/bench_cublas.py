"""cuBLAS (torch.matmul, bf16) on the U-ViT GEMM shapes: the library reference point for tools/bench_gemm.py.
Plain GEMM only (no LayerNorm / GELU / residual / statistics epilogue)."""
import torch
dev = torch.device("cuda:0")
M, D = 257 * 128, 512
for name, N, K in (("qkv", 3 * D, D), ("proj", D, D), ("fc1", 4 * D, D), ("fc2", D, 4 * D), ("skip", D, 2 * D), ("8192^3", 8192, 8192)):
    Mm = 8192 if name == "8192^3" else M
    a = [torch.randn(Mm, K, device=dev).bfloat16() for _ in range(3)]
    w = (torch.randn(N, K, device=dev) * 0.05).bfloat16()
    out = [torch.empty(Mm, N, device=dev, dtype=torch.bfloat16) for _ in range(3)]
    for i in range(3):
        torch.matmul(a[i], w.t(), out=out[i])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 100 if name != "8192^3" else 20
    e0.record()
    for i in range(n):
        torch.matmul(a[i % 3], w.t(), out=out[i % 3])
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / n
    print(f"cuBLAS {name:7s} M={Mm} N={N:5d} K={K:5d}: {us:8.1f} us {2 * Mm * N * K / us / 1e6:7.1f} TFLOP/s", flush=True)
