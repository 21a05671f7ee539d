"""Micro-benchmarks of the individual kernels (CUDA events, L2 flushed between iterations).
Usage (GPU box): python tools/bench_kernels.py [gemm|attn|conv|all]"""
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from unidepth_b200 import ops

dev = torch.device("cuda:0")
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2]


def bench_gemm():
    for (M, N, K) in [(12888, 3072, 1024), (12888, 1024, 1024), (12888, 4096, 1024), (12888, 1024, 4096),
                      (12880, 512, 1024), (12880, 2048, 512), (8192, 8192, 8192)]:
        a = torch.randn(M, K, device=dev).half()
        w = torch.randn(N, K, device=dev).half()
        bias = torch.randn(N, device=dev)
        out = torch.empty(M, N, device=dev, dtype=torch.float16)
        ms = timeit(lambda: ops.gemm(a, w, bias=bias, out=out))
        print(f"gemm {M}x{N}x{K}: {ms:.3f} ms  {2*M*N*K/ms/1e9:.1f} TFLOP/s", flush=True)
        ms = timeit(lambda: torch.matmul(a, w.t()))
        print(f"   torch.matmul: {ms:.3f} ms  {2*M*N*K/ms/1e9:.1f} TFLOP/s", flush=True)
        if N <= 1024:
            x = torch.randn(M, N, device=dev)
            gamma = torch.rand(N, device=dev)
            ms = timeit(lambda: ops.gemm(a, w, bias=bias, gamma=gamma, resid=x, out=x))
            print(f"   + gamma, f32 residual in place: {ms:.3f} ms  {2*M*N*K/ms/1e9:.1f} TFLOP/s", flush=True)
        if N == 4096:
            ms = timeit(lambda: ops.gemm(a, w, bias=bias, act=ops.ACT_GELU, out=out))
            print(f"   + GELU: {ms:.3f} ms  {2*M*N*K/ms/1e9:.1f} TFLOP/s", flush=True)


def bench_ln():
    import torch.nn.functional as F
    for (rows, dim) in [(12888, 1024), (12880, 512)]:
        x = torch.randn(rows, dim, device=dev)
        w, b = torch.randn(dim, device=dev), torch.randn(dim, device=dev)
        out = torch.empty(rows, dim, device=dev, dtype=torch.float16)
        ms = timeit(lambda: ops.layernorm(x, w, b, 1e-6, out=out))
        gb = rows * dim * 6 / 1e9
        print(f"layernorm {rows}x{dim} f32->f16: {ms * 1000:.1f} us  {gb / ms * 1000:.0f} GB/s", flush=True)
        ms = timeit(lambda: F.layer_norm(x, (dim,), w, b, 1e-6))
        print(f"   torch layer_norm f32->f32: {ms * 1000:.1f} us  {rows * dim * 8 / 1e9 / ms * 1000:.0f} GB/s", flush=True)
        ms = timeit(lambda: out.copy_(x))
        print(f"   torch cast copy f32->f16: {ms * 1000:.1f} us  {gb / ms * 1000:.0f} GB/s", flush=True)


def bench_attn():
    for (B, H, S) in [(8, 16, 1611), (8, 8, 1610), (4, 16, 3129)]:
        D = H * 64
        qkv = torch.randn(B * S, 3 * D, device=dev).half()
        out = torch.empty(B * S, D, device=dev, dtype=torch.float16)
        ms = timeit(lambda: ops.attention(qkv, qkv, qkv, out, B=B, heads=H, seq_q=S, seq_k=S, head_dim=64, k_col0=D, v_col0=2 * D))
        fl = 4 * B * H * S * S * 64
        print(f"attn B{B} H{H} S{S}: {ms:.3f} ms  {fl/ms/1e9:.1f} TFLOP/s", flush=True)
        q, k, v = qkv.view(B, S, 3, H, 64).permute(2, 0, 3, 1, 4)
        ms = timeit(lambda: torch.nn.functional.scaled_dot_product_attention(q, k, v))
        print(f"   torch sdpa: {ms:.3f} ms  {fl/ms/1e9:.1f} TFLOP/s", flush=True)


def bench_conv():
    for (B, H, W, C, N) in [(8, 35, 46, 512, 512), (8, 70, 92, 512, 512), (8, 140, 184, 256, 256), (8, 280, 368, 128, 64)]:
        x = torch.randn(B, H, W, C, device=dev).half()
        w = torch.randn(N, 9 * C, device=dev).half()
        bias = torch.randn(N, device=dev)
        out = torch.empty(B, H, W, N, device=dev, dtype=torch.float16)
        ms = timeit(lambda: ops.conv3x3(x, w, bias=bias, out=out))
        fl = 2 * B * H * W * N * 9 * C
        print(f"conv3x3 {B}x{H}x{W}x{C}->{N}: {ms:.3f} ms  {fl/ms/1e9:.1f} TFLOP/s", flush=True)
        if N == C:
            lat = torch.randn(B, H, W, N, device=dev)
            act = torch.empty(B, H, W, N, device=dev, dtype=torch.float16)
            gamma = torch.rand(N, device=dev)
            ms = timeit(lambda: ops.conv3x3(x, w, bias=bias, gamma=gamma, resid=lat, out=lat, out2=act))
            print(f"   + gamma, f32 residual in place, f16 leaky copy: {ms:.3f} ms  {fl/ms/1e9:.1f} TFLOP/s", flush=True)


def bench_halo():
    import torch.nn.functional as F
    for (B, H, W, C, N) in [(8, 280, 368, 128, 64), (8, 490, 644, 64, 32)]:
        x = torch.randn(B, H + 2, W + 2, C, device=dev).half()
        w = torch.randn(N, 9 * C, device=dev).half()
        bias = torch.randn(N, device=dev)
        fl = 2 * B * H * W * N * 9 * C
        if N == 64:
            ms = timeit(lambda: ops.conv3x3_halo(x, w, bias=bias))
        else:
            hw = torch.randn(32, device=dev)
            ms = timeit(lambda: ops.conv3x3_halo(x, w, bias=bias, act=ops.ACT_LEAKY, head_w=hw))
        print(f"conv3x3 halo {B}x{H}x{W}x{C}->{N}: {ms:.3f} ms  {fl/ms/1e9:.1f} TFLOP/s", flush=True)


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    if what in ("gemm", "all"):
        bench_gemm()
    if what in ("attn", "all"):
        bench_attn()
    if what in ("ln", "all"):
        bench_ln()
    if what in ("conv", "all"):
        bench_conv()
        bench_halo()
