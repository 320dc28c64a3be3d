"""Model configuration of the hot path.

Field names follow the reference's HF config attributes (model/InteractVLM.py:146-199 reads them from
`config.json`; scripts/run_train.sh:58-101 sets them for the released checkpoints): `seg_token_idx`,
`img_emb_len`, `multiview_channels`, `multiview_cam_cond`, `cam_encoder_type`, `token_type`,
`hC_sam_view_type`, `hC_loss_weight`, `oC_loss_weight`, `out_dim`, `mm_vision_select_layer` ...
The sub-model dimensions are the ones the reference hard-codes or inherits:
  LLaMA-2-13B (LISA-13B-llama2-v1 config.json), CLIP ViT-L/14 (openai/clip-vit-large-patch14),
  SAM ViT-H (model/segment_anything/build_sam.py:15-23, 52-107).
"""
from __future__ import annotations

from dataclasses import asdict, dataclass, field


@dataclass
class IVLMConfig:
    # ---- LLaMA decoder (HF LlamaConfig names)
    hidden_size: int = 5120
    intermediate_size: int = 13824
    num_hidden_layers: int = 40
    num_attention_heads: int = 40
    vocab_size: int = 32004
    rms_norm_eps: float = 1e-5
    rope_theta: float = 10000.0
    max_position_embeddings: int = 4096
    bos_token_id: int = 1
    eos_token_id: int = 2
    pad_token_id: int = 0
    # ---- CLIP vision tower (HF CLIPVisionConfig names, prefixed)
    clip_hidden_size: int = 1024
    clip_intermediate_size: int = 4096
    clip_num_hidden_layers: int = 24
    clip_num_attention_heads: int = 16
    clip_image_size: int = 224
    clip_patch_size: int = 14
    clip_layer_norm_eps: float = 1e-5
    mm_vision_select_layer: int = -2       # clip_encoder.py:13, hidden_states[-2]
    mm_use_im_start_end: bool = True
    # ---- SAM (build_sam.py)
    sam_embed_dim: int = 1280
    sam_depth: int = 32
    sam_num_heads: int = 16
    sam_global_attn_indexes: tuple = (7, 15, 23, 31)
    sam_window_size: int = 14
    sam_img_size: int = 1024
    sam_patch_size: int = 16
    sam_out_chans: int = 256
    sam_dec_depth: int = 2
    sam_dec_heads: int = 8
    sam_dec_mlp_dim: int = 2048
    sam_num_multimask_outputs: int = 3
    # ---- InteractVLM attributes (InteractVLM.py:146-199)
    out_dim: int = 256
    seg_token_idx: int = 32000
    hseg_token_idx: int | None = None
    oseg_token_idx: int | None = None
    im_start_token_idx: int = 32001
    im_end_token_idx: int = 32002
    img_emb_len: int = 255
    token_type: str = "Gen"
    multiview_channels: int = 4
    multiview_cam_cond: bool = True
    cam_encoder_type: str = "vi_v1"
    hC_sam_view_type: str = "4MV-Z_Vitru"
    oC_sam_view_type: str | None = "4MV-Z_HM"
    hC_loss_weight: float = 3.0
    oC_loss_weight: float = 0.0
    hC_question_type: str = "simple"
    oC_question_type: str = "simple"
    use_fusion: bool = False
    use_uncertainty: bool = False
    train_mask_decoder: bool = True
    vision_tower: str = "openai/clip-vit-large-patch14"
    extra: dict = field(default_factory=dict)

    # ---- derived
    @property
    def head_dim(self) -> int:
        return self.hidden_size // self.num_attention_heads

    @property
    def clip_tokens(self) -> int:
        return (self.clip_image_size // self.clip_patch_size) ** 2 + 1

    @property
    def clip_layers_used(self) -> int:
        """hidden_states[select_layer] of HF CLIP = output of encoder layer (L + select_layer + 1)."""
        return self.clip_num_hidden_layers + self.mm_vision_select_layer + 1

    @property
    def sam_grid(self) -> int:
        return self.sam_img_size // self.sam_patch_size

    def to_dict(self) -> dict:
        d = asdict(self)
        d["sam_global_attn_indexes"] = list(self.sam_global_attn_indexes)
        return d

    @classmethod
    def from_dict(cls, d: dict) -> "IVLMConfig":
        known = {f for f in cls.__dataclass_fields__}
        kw = {k: v for k, v in d.items() if k in known}
        if "sam_global_attn_indexes" in kw:
            kw["sam_global_attn_indexes"] = tuple(kw["sam_global_attn_indexes"])
        cfg = cls(**kw)
        cfg.extra = {k: v for k, v in d.items() if k not in known}
        return cfg

    @classmethod
    def full(cls) -> "IVLMConfig":
        """interactvlm-3d-hcontact-damon: LLaMA-2-13B + CLIP-L/14 + SAM ViT-H."""
        return cls()

    @classmethod
    def tiny(cls) -> "IVLMConfig":
        """Every structural feature of the full model (window + global SAM blocks with the 64->70 window pad,
        head_dim 80 / 64 / 128, 256 CLIP patches, 4 views) at sizes a CPU oracle finishes in seconds."""
        return cls(hidden_size=256, intermediate_size=512, num_hidden_layers=2, num_attention_heads=2, vocab_size=328,
                   clip_hidden_size=128, clip_intermediate_size=256, clip_num_hidden_layers=3,
                   clip_num_attention_heads=2, sam_embed_dim=160, sam_depth=2, sam_num_heads=2,
                   sam_global_attn_indexes=(1,), seg_token_idx=320, im_start_token_idx=321, im_end_token_idx=322,
                   max_position_embeddings=1024)
