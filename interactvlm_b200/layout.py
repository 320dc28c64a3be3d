"""Weight-row orders the fused decode kernels need (csrc/decode_stream.cu), applied ONCE at load time so that every path
(prefill GEMM + rope / silu kernels with their `paired` / `interleaved` flags, fused decode steps) reads the same single copy.

  paired (q and k projections, per head): rows come in blocks of 16 = the 8 features 8b..8b+7 of the head followed by their
      rotary partners 8b+hd/2..8b+7+hd/2 -- both halves of a RoPE rotation land in one 16-row tile of the streaming kernel;
  interleaved (gate / up projections): blocks of 16 = gate rows 8t..8t+7 followed by up rows 8t..8t+7.
"""
from __future__ import annotations

import torch


def paired_index(hd: int) -> torch.Tensor:
    """idx[c] = natural feature (inside a head) stored at paired position c."""
    assert hd % 16 == 0
    half = hd // 2
    j = torch.arange(half)
    idx = torch.empty(hd, dtype=torch.long)
    idx[(j // 8) * 16 + j % 8] = j
    idx[(j // 8) * 16 + 8 + j % 8] = j + half
    return idx


def pair_rows(w: torch.Tensor, heads: int, hd: int) -> torch.Tensor:
    """[heads*hd, K] natural row order -> paired row order."""
    idx = paired_index(hd).to(w.device)
    perm = (torch.arange(heads, device=w.device)[:, None] * hd + idx[None]).reshape(-1)
    return w[perm].contiguous()


def unpair_cols(x: torch.Tensor, heads: int, hd: int) -> torch.Tensor:
    """[..., heads*hd] with paired columns -> natural columns (inverse of pair_rows on the output features)."""
    idx = paired_index(hd).to(x.device)
    out = torch.empty_like(x)
    cols = (torch.arange(heads, device=x.device)[:, None] * hd + idx[None]).reshape(-1)
    out[..., cols] = x
    return out


def interleave_gate_up(gate: torch.Tensor, up: torch.Tensor) -> torch.Tensor:
    """gate, up [F, K] -> [2F, K] in blocks of 8 gate rows / 8 up rows."""
    F, K = gate.shape
    assert F % 8 == 0 and up.shape == gate.shape
    return torch.stack([gate.reshape(F // 8, 8, K), up.reshape(F // 8, 8, K)], 1).reshape(2 * F, K).contiguous()


def split_interleaved(gu: torch.Tensor):
    """[..., 2F] interleaved columns -> (gate [..., F], up [..., F])."""
    F2 = gu.shape[-1]
    t = gu.reshape(*gu.shape[:-1], F2 // 16, 2, 8)
    return t[..., 0, :].reshape(*gu.shape[:-1], F2 // 2), t[..., 1, :].reshape(*gu.shape[:-1], F2 // 2)
