"""Oracle: attention building blocks (test infrastructure only).

Functional restatement, driven by a flat ``state_dict`` and a key prefix, of
  * MultiheadCustomAttention          model/utils/multihead_custom_attention.py:157-462
  * RelativeCrossAttentionLayer/Module, FeedforwardLayer
                                      model/utils/layers.py:293-351
  * ParallelAttentionLayer / ParallelAttention / AdaLN
                                      model/utils/layers.py:7-290
Only the branches the shipped configurations execute are restated (SURVEY.md
App. A "unused-but-present code").
"""
import torch
import torch.nn.functional as F

from .rope import rotate_pairs


def mha_rotary(sd, prefix, num_heads, query, key, value,
               q_rope=None, k_rope=None, key_padding_mask=None,
               return_weights=False):
    """Sequence-first multi-head attention with optional rotary embedding.

    query (Lq, B, E); key/value (Lk, B, E); *_rope (B, L, E, 2) tables or None;
    key_padding_mask (B, Lk) bool, True = ignore.  Returns (Lq, B, E).

    Reference: multi_head_attention_forward,
    multihead_custom_attention.py:242-303 (projections; all three value-equality
    branches use the slices W[0:E], W[E:2E], W[2E:3E]), :325 (scale before
    rotation), :348-353 (rotary on the full E vector, before the head split),
    :355-359 (head split), :391-415 (scores, mask, softmax, PV), :451-452.
    """
    w = sd[prefix + "in_proj_weight"]
    b = sd[prefix + "in_proj_bias"]
    e = query.shape[-1]
    hd = e // num_heads
    lq, bsz, _ = query.shape
    lk = key.shape[0]

    q = F.linear(query, w[:e], b[:e]) * (float(hd) ** -0.5)
    k = F.linear(key, w[e:2 * e], b[e:2 * e])
    v = F.linear(value, w[2 * e:], b[2 * e:])

    if q_rope is not None:
        q = rotate_pairs(q.transpose(0, 1), q_rope[..., 0], q_rope[..., 1]).transpose(0, 1)
        k = rotate_pairs(k.transpose(0, 1), k_rope[..., 0], k_rope[..., 1]).transpose(0, 1)

    q = q.contiguous().view(lq, bsz * num_heads, hd).transpose(0, 1)
    k = k.contiguous().view(lk, bsz * num_heads, hd).transpose(0, 1)
    v = v.contiguous().view(lk, bsz * num_heads, hd).transpose(0, 1)

    scores = torch.bmm(q, k.transpose(1, 2))              # (B*H, Lq, Lk)
    if key_padding_mask is not None:
        scores = scores.view(bsz, num_heads, lq, lk).masked_fill(
            key_padding_mask[:, None, None, :], float("-inf")).view(bsz * num_heads, lq, lk)
    probs = F.softmax(scores, dim=-1)
    ctx = torch.bmm(probs, v)                              # (B*H, Lq, hd)
    ctx = ctx.transpose(0, 1).contiguous().view(lq, bsz, e)
    out = F.linear(ctx, sd[prefix + "out_proj.weight"], sd[prefix + "out_proj.bias"])
    if return_weights:
        return out, probs.view(bsz, num_heads, lq, lk)
    return out


def _layer_norm(sd, prefix, x):
    return F.layer_norm(x, (x.shape[-1],), sd[prefix + "weight"], sd[prefix + "bias"], 1e-5)


def relative_cross_attn_stack(sd, prefix, num_heads, num_layers, query, value,
                              query_rope=None, value_rope=None):
    """Act3D's attention stack: per layer  x <- LN(x + MHA(x, ctx, ctx));
    x <- LN(x + W2 relu(W1 x)).  Returns the list of per-layer outputs.

    Reference: RelativeCrossAttentionModule.forward layers.py:345-351,
    RelativeCrossAttentionLayer.forward :300-310 (rotary only when the query has
    a position), FeedforwardLayer.forward :328-332.  Dropout is 0.
    """
    outs = []
    x = query
    for l in range(num_layers):
        ap = f"{prefix}attn_layers.{l}."
        fp = f"{prefix}ffw_layers.{l}."
        use_rope = query_rope is not None
        att = mha_rotary(sd, ap + "multihead_attn.", num_heads, x, value, value,
                         q_rope=query_rope if use_rope else None,
                         k_rope=value_rope if use_rope else None)
        x = _layer_norm(sd, ap + "norm.", x + att)
        hid = F.relu(F.linear(x, sd[fp + "linear1.weight"], sd[fp + "linear1.bias"]))
        x = _layer_norm(sd, fp + "norm.", x + F.linear(hid, sd[fp + "linear2.weight"], sd[fp + "linear2.bias"]))
        outs.append(x)
    return outs


def ada_ln(sd, prefix, x, t_emb):
    """x (B, N, C), t_emb (B, C):  x * (1 + scale) + shift,
    [scale | shift] = Linear(SiLU(t_emb)).   Reference: AdaLN, layers.py:273-290."""
    mod = F.linear(F.silu(t_emb), sd[prefix + "modulation.1.weight"], sd[prefix + "modulation.1.bias"])
    scale, shift = mod.chunk(2, dim=-1)
    return x * (1 + scale.unsqueeze(1)) + shift.unsqueeze(1)


def parallel_attention_stack(sd, prefix, num_heads, num_layers, seq1, seq1_mask, seq2,
                             seq1_rope=None, seq2_rope=None, seq1_sem_pos=None,
                             ada_signal=None, self_attention=False, rotary=False,
                             use_adaln=False, apply_ffn=True):
    """ChainedDiffuser's post-norm layer stack acting on seq1 only
    (cross_attention1 [+ self_attention1] + FFN-1; seq2 is read-only context).

    seq1 (B, S1, C), seq2 (B, S2, C) batch-first.  Reference:
    ParallelAttentionLayer.forward layers.py:115-218 with pre_norm=False,
    cross_attention2=self_attention2=False; ParallelAttention.forward :252-270.
    Eval mode (all dropouts inactive).
    """
    x = seq1
    for l in range(num_layers):
        lp = f"{prefix}layers.{l}."
        # ---- cross attention seq1 -> seq2 (layers.py:124-147)
        q1 = x
        if not rotary:
            pass                                  # seq1_pos/seq2_pos are None at every call site
        if seq1_sem_pos is not None:
            q1 = q1 + seq1_sem_pos
        qq = ada_ln(sd, lp + "adaln_12.", q1, ada_signal) if (use_adaln and ada_signal is not None) else q1
        att = mha_rotary(sd, lp + "cross_12.", num_heads,
                         qq.transpose(0, 1), seq2.transpose(0, 1), seq2.transpose(0, 1),
                         q_rope=seq1_rope if rotary else None,
                         k_rope=seq2_rope if rotary else None).transpose(0, 1)
        x = _layer_norm(sd, lp + "norm_12.", x + att)
        # ---- self attention on seq1 (layers.py:165-182)
        if self_attention:
            qk = x if seq1_sem_pos is None else x + seq1_sem_pos
            vv = x
            if use_adaln and ada_signal is not None:
                qk = ada_ln(sd, lp + "adaln_1.", qk, ada_signal)
                vv = ada_ln(sd, lp + "adaln_1.", vv, ada_signal)
            att = mha_rotary(sd, lp + "sa1.", num_heads,
                             qk.transpose(0, 1), qk.transpose(0, 1), vv.transpose(0, 1),
                             q_rope=seq1_rope if rotary else None,
                             k_rope=seq1_rope if rotary else None,
                             key_padding_mask=seq1_mask).transpose(0, 1)
            x = _layer_norm(sd, lp + "norm_1.", x + att)
        # ---- FFN-1 (layers.py:205-209)
        if apply_ffn:
            y = ada_ln(sd, lp + "adaln_ff1.", x, ada_signal) if (use_adaln and ada_signal is not None) else x
            hid = F.relu(F.linear(y, sd[lp + "ffn_12.0.weight"], sd[lp + "ffn_12.0.bias"]))
            y2 = F.linear(hid, sd[lp + "ffn_12.3.weight"], sd[lp + "ffn_12.3.bias"])
            x = _layer_norm(sd, lp + "norm_122.", y + y2)
    return x
