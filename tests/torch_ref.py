"""Independent float64 torch restatement of the Magpie graph, used only to cross-check the oracle.

Deliberately formulated differently from oracle/magpie_oracle.c: whole-sequence (uncached) decoder
with an explicit causal mask, torch.nn.functional convs (grouped conv_transpose1d, dilated conv1d),
exact-tanh GELU without the f16 table.  Follows the PyTorch-side semantics the reference author
validated against (docs/MAGPIE_ARCHITECTURE.md, docs/CODEC_ARCHITECTURE.md, scripts/dump_*.py).
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F

DT = torch.float64


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DT)


def load_tensors(path):
    """GGUF -> {name: float64 torch tensor in PyTorch shape}. F32/F16/Q8_0 supported."""
    from gguf import GGUFReader, GGMLQuantizationType
    import gguf.quants as q
    r = GGUFReader(path)
    out = {}
    for t in r.tensors:
        shape = [int(x) for x in reversed(t.shape)]
        if t.tensor_type == GGMLQuantizationType.F32:
            a = np.array(t.data, dtype=np.float32)
        elif t.tensor_type == GGMLQuantizationType.F16:
            a = np.array(t.data).astype(np.float32)
        else:
            a = q.dequantize(np.array(t.data), t.tensor_type)
        out[t.name] = _t(a.reshape(shape))
    return out


def layer_norm(x, w, eps):
    return F.layer_norm(x, (x.shape[-1],), w, None, eps)


def gelu(x):
    return F.gelu(x, approximate="tanh")


def mha(q, k, v, heads, mask=None):
    """q [Tq, H*dh], k/v [Tk, H*dh] -> [Tq, H*dh]; mask [Tq, Tk] additive."""
    Tq, Tk = q.shape[0], k.shape[0]
    dh = q.shape[1] // heads
    qh = q.view(Tq, heads, dh).transpose(0, 1)
    kh = k.view(Tk, heads, dh).transpose(0, 1)
    vh = v.view(Tk, heads, dh).transpose(0, 1)
    s = qh @ kh.transpose(1, 2) / math.sqrt(dh)
    if mask is not None:
        s = s + mask
    p = torch.softmax(s, dim=-1)
    return (p @ vh).transpose(0, 1).reshape(Tq, heads * dh)


def causal_mask(T):
    m = torch.full((T, T), float("-inf"), dtype=DT)
    return torch.triu(m, diagonal=1)


def conv_ffn(x, wp, wo, k):
    """x [T, d]; wp (f, d, k); wo (d, f, k); causal conv, GELU in between, no bias."""
    xt = x.t().unsqueeze(0)
    h = F.conv1d(F.pad(xt, (k - 1, 0)), wp)
    h = gelu(h)
    y = F.conv1d(F.pad(h, (k - 1, 0)), wo)
    return y.squeeze(0).t()


def encode_text(W, hp, tokens):
    tok = torch.as_tensor(np.asarray(tokens), dtype=torch.long)
    E = len(tok)
    x = W["text_embedding.weight"][tok] + W["encoder.position_embeddings.weight"][:E]
    mask = causal_mask(E)
    d = hp["d_model"]
    for l in range(hp["enc_layers"]):
        p = f"encoder.layers.{l}."
        n = layer_norm(x, W[p + "norm_self.weight"], hp["eps"])
        qkv = n @ W[p + "self_attention.qkv_net.weight"].t()
        a = mha(qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:], hp["enc_heads"], mask)
        x = x + a @ W[p + "self_attention.o_net.weight"].t()
        n = layer_norm(x, W[p + "norm_pos_ff.weight"], hp["eps"])
        x = x + conv_ffn(n, W[p + "pos_ff.proj.conv.weight"], W[p + "pos_ff.o_net.conv.weight"], hp["enc_kernel"])
    return layer_norm(x, W["encoder.norm_out.weight"], hp["eps"])


def decoder_full(W, hp, x_in, enc_out):
    """Uncached decoder over the whole sequence x_in [T, d] (pos-emb already added)."""
    T = x_in.shape[0]
    d = hp["d_model"]
    dxa = hp["dec_xa_heads"] * hp["dec_xa_d_head"]
    mask = causal_mask(T)
    x = x_in
    for l in range(hp["dec_layers"]):
        p = f"decoder.layers.{l}."
        n = layer_norm(x, W[p + "norm_self.weight"], hp["eps"])
        qkv = n @ W[p + "self_attention.qkv_net.weight"].t()
        a = mha(qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:], hp["dec_sa_heads"], mask)
        x = x + a @ W[p + "self_attention.o_net.weight"].t()
        nq = layer_norm(x, W[p + "norm_xattn_query.weight"], hp["eps"])
        mem = layer_norm(enc_out, W[p + "norm_xattn_memory.weight"], hp["eps"])
        kv = mem @ W[p + "cross_attention.kv_net.weight"].t()
        q = nq @ W[p + "cross_attention.q_net.weight"].t()
        a = mha(q, kv[:, :dxa], kv[:, dxa:], hp["dec_xa_heads"], None)
        x = x + a @ W[p + "cross_attention.o_net.weight"].t()
        n = layer_norm(x, W[p + "norm_pos_ff.weight"], hp["eps"])
        x = x + conv_ffn(n, W[p + "pos_ff.proj.conv.weight"], W[p + "pos_ff.o_net.conv.weight"], hp["dec_kernel"])
    return layer_norm(x, W["decoder.norm_out.weight"], hp["eps"])


def audio_embedding(W, codes):
    return sum(W[f"audio_embeddings.{cb}.weight"][int(codes[cb])] for cb in range(8)) / 8.0


def decoder_teacher_forced(W, hp, enc_out, speaker, frames):
    """frames: [n][8] codes fed at positions C, C+1, ... (first is usually the BOS frame).
    Returns hidden [n][d] for those positions."""
    C, d = hp["context_frames"], hp["d_model"]
    ctx = W["baked_context_embedding.weight"][speaker].view(C, d)
    embs = torch.stack([audio_embedding(W, f) for f in frames])
    x = torch.cat([ctx, embs], 0)
    x = x + W["decoder.position_embeddings.weight"][: x.shape[0]]
    return decoder_full(W, hp, x, enc_out)[C:]


def lt_logits(W, hp, hidden, fed_codes):
    """Teacher-forced local transformer: returns raw (unmasked) logits [8][V]."""
    L = hp["lt_dim"]
    win, bin_ = W["local_transformer_in_projection.weight"], W["local_transformer_in_projection.bias"]
    seq = [hidden @ win.t() + bin_]
    for cb in range(7):
        e = W[f"audio_embeddings.{cb}.weight"][int(fed_codes[cb])]
        seq.append(e @ win.t() + bin_)
    x = torch.stack(seq) + W["local_transformer.position_embeddings.weight"][:8]
    p = "local_transformer.layers.0."
    n = layer_norm(x, W[p + "norm_self.weight"], hp["eps"])
    qkv = n @ W[p + "self_attention.qkv_net.weight"].t()
    a = mha(qkv[:, :L], qkv[:, L:2 * L], qkv[:, 2 * L:], hp["lt_heads"], causal_mask(8))
    x = x + a @ W[p + "self_attention.o_net.weight"].t()
    n = layer_norm(x, W[p + "norm_pos_ff.weight"], hp["eps"])
    x = x + conv_ffn(n, W[p + "pos_ff.proj.conv.weight"], W[p + "pos_ff.o_net.conv.weight"], 1)
    out = []
    for cb in range(8):
        out.append(x[cb] @ W[f"local_transformer_out_projections.{cb}.weight"].t()
                   + W[f"local_transformer_out_projections.{cb}.bias"])
    return torch.stack(out)


# ------------------------------------------------------------------------------------------------
# codec
# ------------------------------------------------------------------------------------------------

def fsq_dequant(codes):
    """codes [8][T] -> [32][T], the documented formula (docs/CODEC_ARCHITECTURE.md:88-101)."""
    codes = torch.as_tensor(np.asarray(codes), dtype=torch.long)
    base = torch.tensor([1, 8, 56, 336])
    lev = torch.tensor([8, 7, 6, 6])
    nonneg = (codes[:, None, :] // base[None, :, None]) % lev[None, :, None]
    half = (lev // 2)[None, :, None]
    v = (nonneg - half).to(torch.float32) / half.to(torch.float32)
    return v.reshape(-1, codes.shape[1])


def half_snake(x, alpha):
    """x [C][T]; alpha (1, n, 1)"""
    n = alpha.numel()
    a = alpha.reshape(n, 1)
    xs, xl = x[:n], x[n:]
    return torch.cat([xs + torch.sin(a * xs) ** 2 / a, F.leaky_relu(xl, 0.01)], 0)


def causal_conv(x, w, b, dil=1):
    k = w.shape[-1]
    return F.conv1d(F.pad(x.unsqueeze(0), ((k - 1) * dil, 0)), w, b, dilation=dil).squeeze(0)


def conv_transpose(x, w, b, stride):
    cout = w.shape[0] // 2
    y = F.conv_transpose1d(x.unsqueeze(0), w, b, stride=stride, groups=cout).squeeze(0)
    return y[:, : x.shape[1] * stride]


def codec_decode(W, codes, up_rates=(8, 8, 4, 2, 2)):
    x = fsq_dequant(codes).to(DT)
    x = causal_conv(x, W["dec.pre.weight"], W["dec.pre.bias"])
    for i, s in enumerate(up_rates):
        x = half_snake(x, W[f"dec.act.{i}.activation.snake_act.alpha"])
        x = conv_transpose(x, W[f"dec.up.{i}.c.weight"], W[f"dec.up.{i}.c.bias"], s)
        acc = None
        for j in range(3):
            o = x
            for k, dil in enumerate((1, 3, 5)):
                p = f"dec.rl.{i}.rb.{j}.rb.{k}."
                h = half_snake(o, W[p + "in_act.alpha"])
                h = causal_conv(h, W[p + "in_conv.weight"], W[p + "in_conv.bias"], dil)
                h = half_snake(h, W[p + "sk_act.alpha"])
                h = causal_conv(h, W[p + "sk_conv.weight"], W[p + "sk_conv.bias"], 1)
                o = o + h
            acc = o if acc is None else acc + o
        x = acc / 3.0
    x = half_snake(x, W["dec.post_act.alpha"])
    x = causal_conv(x, W["dec.post.weight"], W["dec.post.bias"])
    return torch.tanh(x).reshape(-1)
