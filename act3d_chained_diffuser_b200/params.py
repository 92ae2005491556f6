"""Parameter containers that reproduce the reference's state_dict key names.

The fused kernels consume packed copies of these parameters (packing.py); the modules here only
own the nn.Parameters (so checkpoints load with strict=True, optimisers / DDP see ordinary
parameters) and initialise them like the reference does.
"""
import math

import torch
from torch import nn


class AttnProj(nn.Module):
    """in_proj_weight (3E,E) / in_proj_bias (3E) / out_proj.{weight,bias}
    (keys of MultiheadCustomAttention, multihead_custom_attention.py:51-69; init :80-92)."""

    def __init__(self, embed_dim, num_heads):
        super().__init__()
        self.embed_dim, self.num_heads = embed_dim, num_heads
        self.in_proj_weight = nn.Parameter(torch.empty(3 * embed_dim, embed_dim))
        self.in_proj_bias = nn.Parameter(torch.zeros(3 * embed_dim))
        self.out_proj = nn.Linear(embed_dim, embed_dim)
        nn.init.xavier_uniform_(self.in_proj_weight)
        nn.init.zeros_(self.out_proj.bias)


class XAttnLayerParams(nn.Module):
    """keys: multihead_attn.*, norm.*   (RelativeCrossAttentionLayer, layers.py:293-298)"""

    def __init__(self, embed_dim, num_heads):
        super().__init__()
        self.multihead_attn = AttnProj(embed_dim, num_heads)
        self.norm = nn.LayerNorm(embed_dim)


class FfwParams(nn.Module):
    """keys: linear1.*, linear2.*, norm.*   (FeedforwardLayer, layers.py:313-326: xavier on matrices)"""

    def __init__(self, embed_dim, hidden_dim):
        super().__init__()
        self.linear1 = nn.Linear(embed_dim, hidden_dim)
        self.linear2 = nn.Linear(hidden_dim, embed_dim)
        self.norm = nn.LayerNorm(embed_dim)
        nn.init.xavier_uniform_(self.linear1.weight)
        nn.init.xavier_uniform_(self.linear2.weight)


class XAttnStackParams(nn.Module):
    """keys: attn_layers.l.*, ffw_layers.l.*   (RelativeCrossAttentionModule, layers.py:335-343)"""

    def __init__(self, embed_dim, num_heads, num_layers):
        super().__init__()
        self.embed_dim, self.num_heads, self.num_layers = embed_dim, num_heads, num_layers
        self.attn_layers = nn.ModuleList(XAttnLayerParams(embed_dim, num_heads) for _ in range(num_layers))
        self.ffw_layers = nn.ModuleList(FfwParams(embed_dim, embed_dim) for _ in range(num_layers))


class AdaLNParams(nn.Module):
    """keys: modulation.1.{weight,bias}, zero-initialised (AdaLN, layers.py:273-280)"""

    def __init__(self, dim):
        super().__init__()
        lin = nn.Linear(dim, 2 * dim)
        nn.init.zeros_(lin.weight)
        nn.init.zeros_(lin.bias)
        self.modulation = nn.Sequential(nn.SiLU(), lin)


class ParallelLayerParams(nn.Module):
    """One ParallelAttentionLayer restricted to the seq1 branches the planner builds
    (layers.py:10-100): cross_12 / norm_12 [/ adaln_12], sa1 / norm_1 [/ adaln_1],
    ffn_12.{0,3} / norm_122 [/ adaln_ff1].  dropout modules hold no parameters."""

    def __init__(self, d_model, n_heads, self_attention, use_adaln):
        super().__init__()
        if self_attention:
            if use_adaln:
                self.adaln_1 = AdaLNParams(d_model)
            self.sa1 = AttnProj(d_model, n_heads)
            self.norm_1 = nn.LayerNorm(d_model)
        if use_adaln:
            self.adaln_12 = AdaLNParams(d_model)
        self.cross_12 = AttnProj(d_model, n_heads)
        self.norm_12 = nn.LayerNorm(d_model)
        if use_adaln:
            self.adaln_ff1 = AdaLNParams(d_model)
        self.ffn_12 = nn.Sequential(nn.Linear(d_model, 4 * d_model), nn.ReLU(), nn.Dropout(0.1),
                                    nn.Linear(4 * d_model, d_model), nn.Dropout(0.1))
        self.norm_122 = nn.LayerNorm(d_model)


class ParallelStackParams(nn.Module):
    """keys: layers.l.*   (ParallelAttention, layers.py:221-250)"""

    def __init__(self, num_layers, d_model, n_heads, self_attention=False, rotary_pe=False, use_adaln=False,
                 apply_ffn=True):
        super().__init__()
        self.num_layers, self.d_model, self.n_heads = num_layers, d_model, n_heads
        self.self_attention, self.rotary_pe, self.use_adaln, self.apply_ffn = self_attention, rotary_pe, use_adaln, apply_ffn
        self.layers = nn.ModuleList(ParallelLayerParams(d_model, n_heads, self_attention, use_adaln)
                                    for _ in range(num_layers))


def mlp(in_dim, hidden, out_dim, dropout=None):
    mods = [nn.Linear(in_dim, hidden), nn.ReLU()]
    if dropout is not None:
        mods.append(nn.Dropout(dropout))
    mods.append(nn.Linear(hidden, out_dim))
    return nn.Sequential(*mods)
