"""GPU-box diagnostics: per-op timing (CUDA events) for the shapes on the hot path. Writes gpurun_out/diag.json."""
import json, sys, time
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from interactvlm_b200.ops import Context

ctx = Context(0)
out = {}

def timeit(fn, iters=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters

def gemm_case(name, M, N, K, **kw):
    a = torch.randn(M, K, device="cuda").bfloat16(); w = (torch.randn(N, K, device="cuda") * K ** -0.5).bfloat16()
    try:
        o = ctx.gemm(a, w, **kw)
        ref = a.float() @ w.float().t()
        err = ((o.float() - ref).norm() / ref.norm()).item()
        ms = timeit(lambda: ctx.gemm(a, w, out=o, **kw))
        ms_t = timeit(lambda: torch.matmul(a, w.t()))
        out[name] = dict(M=M, N=N, K=K, err=err, ms=ms, tflops=2 * M * N * K / ms / 1e9, cublas_ms=ms_t,
                         cublas_tflops=2 * M * N * K / ms_t / 1e9, gbs=(M * K + N * K + M * N) * 2 / ms / 1e6)
    except Exception as ex:
        out[name] = dict(error=repr(ex))
    print(name, out[name], flush=True)

gemm_case("sam_qkv", 16384, 3840, 1280)
gemm_case("sam_proj", 16384, 1280, 1280)
gemm_case("sam_mlp1", 16384, 5120, 1280)
gemm_case("sam_mlp2", 16384, 1280, 5120)
gemm_case("sam_qkv_b8", 131072, 3840, 1280)
gemm_case("llama_qkv_prefill", 2640, 15360, 5120)
gemm_case("llama_gateup_prefill", 2640, 27648, 5120)
gemm_case("llama_down_prefill", 2640, 5120, 13824)
gemm_case("llama_qkv_decode", 8, 15360, 5120)
gemm_case("llama_o_decode", 8, 5120, 5120)
gemm_case("llama_gateup_decode", 8, 27648, 5120)
gemm_case("llama_down_decode", 8, 5120, 13824)
gemm_case("llama_down_decode_splitk", 8, 5120, 13824, k_splits=8, out_dtype=torch.float32)
gemm_case("llama_o_decode_splitk", 8, 5120, 5120, k_splits=4, out_dtype=torch.float32)
gemm_case("big_square", 8192, 8192, 8192)

def attn_case(name, B, H, S, D, causal=False, rel=None):
    qkv = torch.randn(B, S, 3, H, D, device="cuda").bfloat16()
    q, k, v = qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2]
    try:
        kw = {}
        if rel:
            Hq, Wq = rel
            kw = dict(rel_h=torch.randn(B, H, S, Hq, device="cuda"), rel_w=torch.randn(B, H, S, Wq, device="cuda"), kh=Hq, kw=Wq)
        ms = timeit(lambda: ctx.attention(q, k, v, D ** -0.5, causal=causal, **kw))
        fl = 4 * B * H * S * S * D * (0.5 if causal else 1)
        out[name] = dict(ms=ms, tflops=fl / ms / 1e9)
    except Exception as ex:
        out[name] = dict(error=repr(ex))
    print(name, out[name], flush=True)

attn_case("sam_global_v4", 4, 16, 4096, 80, rel=(64, 64))
attn_case("sam_window_v4", 100, 16, 196, 80, rel=(14, 14))
attn_case("clip_b8", 8, 16, 257, 64)
attn_case("llama_prefill_b8", 8, 40, 330, 128, causal=True)

x = torch.randn(16384, 1280, device="cuda").bfloat16(); g = torch.ones(1280, device="cuda").bfloat16(); b = torch.zeros(1280, device="cuda").bfloat16()
ms = timeit(lambda: ctx.layernorm(x, g, b, 1e-6))
out["layernorm_16384x1280"] = dict(ms=ms, gbs=2 * x.numel() * 2 / ms / 1e6)
print("layernorm", out["layernorm_16384x1280"])
Path(ROOT / "gpurun_out").mkdir(exist_ok=True)
json.dump(out, open(ROOT / "gpurun_out" / "diag.json", "w"), indent=1)
