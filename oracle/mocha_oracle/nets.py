"""NumPy restatement of the reference networks on the hot path (fp32 throughout, like torch CPU).

Weights come in as a dict keyed exactly like the reference state_dicts (numpy arrays)."""
from __future__ import annotations

import numpy as np
from scipy.special import erf

F32 = np.float32


def leaky_relu(x, slope=0.2):
    return np.where(x > 0, x, F32(slope) * x).astype(F32)


def gelu(x):
    # nn.GELU() default = exact erf form (net/transformer.py:28)
    return (F32(0.5) * x * (F32(1.0) + erf(x / np.sqrt(F32(2.0))))).astype(F32)


def conv1x1(x, w, b):
    """nn.Conv2d with a (1,1) kernel on [B,C,T,V] (model.py:44,78; net/blocks.py:49-56)."""
    y = np.einsum("oc,bctv->botv", w[:, :, 0, 0], x, optimize=True)
    return (y + b[None, :, None, None]).astype(F32)


def temporal_conv_reflect(x, w, b):
    """nn.Conv2d (k,1), padding ((k-1)//2, 0), padding_mode='reflect' (net/blocks.py:112-118)."""
    k = w.shape[2]
    pad = (k - 1) // 2
    T = x.shape[2]
    xp = np.pad(x, ((0, 0), (0, 0), (pad, pad), (0, 0)), mode="reflect")
    y = np.zeros((x.shape[0], w.shape[0], T, x.shape[3]), dtype=F32)
    for i in range(k):
        y += np.einsum("oc,bctv->botv", w[:, :, i, 0], xp[:, :, i:i + T], optimize=True)
    return (y + b[None, :, None, None]).astype(F32)


def stgcn_block(sd, prefix, x, A):
    """STGCN_Block.forward with norm='none', activation='lrelu' (net/blocks.py:124-134) and
    SpatialConv.forward (net/blocks.py:57-66)."""
    x = leaky_relu(x)
    y = conv1x1(x, sd[prefix + ".gcn.conv.weight"], sd[prefix + ".gcn.conv.bias"])
    B, KC, T, V = y.shape
    K = A.shape[0]
    y = y.reshape(B, K, KC // K, T, V)
    y = np.einsum("nkctv,kvw->nctw", y, A, optimize=True).astype(F32)
    return temporal_conv_reflect(y, sd[prefix + ".tcn.weight"], sd[prefix + ".tcn.bias"])


def mot_embedding(sd, X, tp=4):
    """Generator.mot_embedding (model.py:42-50): [B,T,V,15] -> [B, (T/tp)*6, D]."""
    x = np.transpose(X, (0, 3, 1, 2)).astype(F32)                       # b t v c -> b c t v
    x = conv1x1(x, sd["mot_embedding.1.weight"], sd["mot_embedding.1.bias"])
    x = stgcn_block(sd, "mot_embedding.2.blk", x, sd["mot_embedding.2.A_j"])
    x = np.einsum("nctv,vw->nctw", x, sd["mot_embedding.3.weight"], optimize=True).astype(F32)  # graph.py:463-465
    B, Cc, T, P = x.shape
    x = x.reshape(B, Cc, T // tp, tp, P).mean(axis=3).astype(F32)       # AvgPool2d((tp,1)) model.py:47
    x = stgcn_block(sd, "mot_embedding.5.blk", x, sd["mot_embedding.5.A_b"])
    return np.transpose(x, (0, 2, 3, 1)).reshape(B, -1, Cc).astype(F32)  # b c t v -> b (t v) c


def mean_variance_norm(x, eps=1e-5):
    """net/transformer.py:13-20 on (B, C, n): unbiased std, eps added to std."""
    B, Cc = x.shape[0], x.shape[1]
    v = x.reshape(B, Cc, -1)
    mean = v.mean(axis=-1, keepdims=True, dtype=F32)
    std = v.std(axis=-1, keepdims=True, ddof=1, dtype=F32)
    return ((v - mean) / (std + F32(eps))).reshape(x.shape).astype(F32)


def _instance_norm_tokens(x, eps=1e-5):
    """InstanceNorm1d wrapped in Rearrange('b s c -> b c s') (net/transformer.py:49-52,116-121)."""
    return np.transpose(mean_variance_norm(np.transpose(x, (0, 2, 1)), eps), (0, 2, 1))


def softmax(x):
    m = x.max(axis=-1, keepdims=True)
    e = np.exp(x - m)
    return (e / e.sum(axis=-1, keepdims=True)).astype(F32)


def attention(sd, prefix, src, tar, heads, adain):
    """Attention.forward (net/transformer.py:63-76); to_q/to_k see instance-normed inputs when adain."""
    tar = src if tar is None else tar
    q_in = _instance_norm_tokens(src) if adain else src
    k_in = _instance_norm_tokens(tar) if adain else tar
    q = q_in @ sd[prefix + ".to_q.1.weight"].T
    k = k_in @ sd[prefix + ".to_k.1.weight"].T
    v = tar @ sd[prefix + ".to_v.weight"].T
    B, n, inner = q.shape
    dh = inner // heads

    def split(t):
        return t.reshape(B, t.shape[1], heads, dh).transpose(0, 2, 1, 3)

    q, k, v = split(q), split(k), split(v)
    dots = (q @ k.transpose(0, 1, 3, 2)) * F32(dh ** -0.5)
    out = softmax(dots) @ v
    out = out.transpose(0, 2, 1, 3).reshape(B, n, inner)
    return (out @ sd[prefix + ".to_out.0.weight"].T + sd[prefix + ".to_out.0.bias"]).astype(F32)


def feed_forward(sd, prefix, x):
    """FeedForward (net/transformer.py:23-34), dropout = identity in eval."""
    h = gelu(x @ sd[prefix + ".net.0.weight"].T + sd[prefix + ".net.0.bias"])
    return (h @ sd[prefix + ".net.3.weight"].T + sd[prefix + ".net.3.bias"]).astype(F32)


def adain(sd, prefix, x, style):
    """AdaIN.forward (net/transformer.py:107-113)."""
    s = style.mean(axis=1, dtype=F32)                                   # AdaptiveAvgPool1d(1)
    h = leaky_relu(s @ sd[prefix + ".style.2.weight"].T + sd[prefix + ".style.2.bias"])
    gb = h @ sd[prefix + ".style.4.weight"].T + sd[prefix + ".style.4.bias"]
    Cc = x.shape[2]
    gamma, beta = gb[:, None, :Cc], gb[:, None, Cc:]
    return ((F32(1.0) + gamma) * _instance_norm_tokens(x) + beta).astype(F32)


def transformer(sd, prefix, x, sty, depth, heads, use_adain):
    """Transformer.forward (net/transformer.py:90-95)."""
    for l in range(depth):
        p = f"{prefix}.layers.{l}"
        if use_adain and sty is not None:
            x = adain(sd, p + ".0", x, sty)
        x = attention(sd, p + ".1", x, sty, heads, use_adain) + x
        x = feed_forward(sd, p + ".2", x) + x
    return x.astype(F32)


def encoder(sd, tokens, depth=2, heads=4):
    return transformer(sd, "encoder", tokens, None, depth, heads, False)


def decoder(sd, src, cha, depth=2, heads=4):
    return transformer(sd, "decoder", src, cha, depth, heads, True)


def to_mot(sd, tokens, tp=4, nbody=6):
    """Generator.to_mot (model.py:71-80): [B, n_tok, D] -> [B, T, V, 15]."""
    B, n, Cc = tokens.shape
    x = tokens.reshape(B, n // nbody, nbody, Cc).transpose(0, 3, 1, 2).astype(F32)  # b (t v) c -> b c t v
    x = stgcn_block(sd, "to_mot.1.blk", x, sd["to_mot.1.A_b"])
    x = np.repeat(x, tp, axis=2)                                          # nearest, scale (tp,1)  model.py:165-174
    x = np.einsum("nctv,vw->nctw", x, sd["to_mot.3.weight"], optimize=True).astype(F32)  # graph.py:606-608
    x = stgcn_block(sd, "to_mot.4.blk", x, sd["to_mot.4.A_j"])
    x = leaky_relu(x)
    x = conv1x1(x, sd["to_mot.6.weight"], sd["to_mot.6.bias"])
    return np.transpose(x, (0, 2, 3, 1)).astype(F32)


def generator_forward(sd, src_X, cha_X, extract_feature=False):
    """Generator.forward (model.py:82-106)."""
    st = mot_embedding(sd, src_X) + sd["pos_emb"][:, :90]
    ct = mot_embedding(sd, cha_X) + sd["pos_emb"][:, :90]
    se, ce = encoder(sd, st), encoder(sd, ct)
    if extract_feature:
        sc = np.transpose(mean_variance_norm(np.transpose(se, (0, 2, 1))), (0, 2, 1))
        cc = np.transpose(mean_variance_norm(np.transpose(ce, (0, 2, 1))), (0, 2, 1))
        return se, ce, sc, cc
    return to_mot(sd, decoder(sd, se, ce))


# ---------------------------------------------------------------------------------------------------
# CVAE (model_CVAE.py); torch.nn.TransformerEncoderLayer / DecoderLayer defaults: post-LN,
# layer_norm_eps=1e-5, batch_first=True, activation=relu, dropout identity in eval
# ---------------------------------------------------------------------------------------------------
def layer_norm(x, g, b, eps=1e-5):
    mean = x.mean(axis=-1, keepdims=True, dtype=F32)
    var = ((x - mean) ** 2).mean(axis=-1, keepdims=True, dtype=F32)
    return ((x - mean) / np.sqrt(var + F32(eps)) * g + b).astype(F32)


def multihead_attention(sd, prefix, q_in, kv_in, heads):
    """nn.MultiheadAttention forward with packed in_proj, no masks."""
    W, bias = sd[prefix + ".in_proj_weight"], sd[prefix + ".in_proj_bias"]
    D = q_in.shape[-1]
    q = q_in @ W[:D].T + bias[:D]
    k = kv_in @ W[D:2 * D].T + bias[D:2 * D]
    v = kv_in @ W[2 * D:].T + bias[2 * D:]
    B, nq, _ = q.shape
    dh = D // heads

    def split(t):
        return t.reshape(B, t.shape[1], heads, dh).transpose(0, 2, 1, 3)

    q, k, v = split(q), split(k), split(v)
    att = softmax((q * F32(dh ** -0.5)) @ k.transpose(0, 1, 3, 2)) @ v
    att = att.transpose(0, 2, 1, 3).reshape(B, nq, D)
    return (att @ sd[prefix + ".out_proj.weight"].T + sd[prefix + ".out_proj.bias"]).astype(F32)


def encoder_layer(sd, p, x, heads):
    x = layer_norm(x + multihead_attention(sd, p + ".self_attn", x, x, heads), sd[p + ".norm1.weight"], sd[p + ".norm1.bias"])
    h = np.maximum(x @ sd[p + ".linear1.weight"].T + sd[p + ".linear1.bias"], 0).astype(F32)
    h = h @ sd[p + ".linear2.weight"].T + sd[p + ".linear2.bias"]
    return layer_norm(x + h, sd[p + ".norm2.weight"], sd[p + ".norm2.bias"])


def decoder_layer(sd, p, x, mem, heads):
    x = layer_norm(x + multihead_attention(sd, p + ".self_attn", x, x, heads), sd[p + ".norm1.weight"], sd[p + ".norm1.bias"])
    x = layer_norm(x + multihead_attention(sd, p + ".multihead_attn", x, mem, heads), sd[p + ".norm2.weight"], sd[p + ".norm2.bias"])
    h = np.maximum(x @ sd[p + ".linear1.weight"].T + sd[p + ".linear1.bias"], 0).astype(F32)
    h = h @ sd[p + ".linear2.weight"].T + sd[p + ".linear2.bias"]
    return layer_norm(x + h, sd[p + ".norm3.weight"], sd[p + ".norm3.bias"])


def cvae_prior(sd, c, depth=2, heads=4):
    """PriorNet.encode (model_CVAE.py:70-79)."""
    B = c.shape[0]
    mu_t = np.repeat(sd["prior_net.mu_token"], B, axis=0)
    lv_t = np.repeat(sd["prior_net.logvar_token"], B, axis=0)
    x = np.concatenate([mu_t, lv_t, c], axis=1).astype(F32)
    x = x + sd["prior_net.pos_encoder.pe"][:, :x.shape[1]]
    for l in range(depth):
        x = encoder_layer(sd, f"prior_net.encoder.layers.{l}", x, heads)
    return x[:, 0], x[:, 1]


def cvae_encode(sd, x, c, depth=2, heads=4):
    """Encoder.encode, the posterior (model_CVAE.py:116-126): the prior's network over [mu, logvar, c, x], own weights."""
    B = c.shape[0]
    mu_t = np.repeat(sd["encoder.mu_token"], B, axis=0)
    lv_t = np.repeat(sd["encoder.logvar_token"], B, axis=0)
    t = np.concatenate([mu_t, lv_t, c, x], axis=1).astype(F32)
    t = t + sd["encoder.pos_encoder.pe"][:, :t.shape[1]]
    for l in range(depth):
        t = encoder_layer(sd, f"encoder.encoder.layers.{l}", t, heads)
    return t[:, 0], t[:, 1]


def cvae_forward(sd, x, c, eps, out_seq=90, depth=2, heads=4):
    """CVAE.forward (model_CVAE.py:37-42) with the posterior's noise given: out, (mu_po, logvar_po), (mu_pr, logvar_pr)."""
    mu_po, lv_po = cvae_encode(sd, x, c, depth, heads)
    z = (mu_po + eps * np.exp(F32(0.5) * lv_po)).astype(F32)
    return cvae_decode(sd, z, c, out_seq, depth, heads), (mu_po, lv_po), cvae_prior(sd, c, depth, heads)


def cvae_decode(sd, z, c, out_seq=90, depth=2, heads=4):
    """Decoder.forward (model_CVAE.py:159-165)."""
    B = c.shape[0]
    mem = np.concatenate([z[:, None, :], c], axis=1).astype(F32)
    x = np.zeros((B, out_seq, c.shape[2]), dtype=F32) + sd["decoder.pos_encoder.pe"][:, :out_seq]
    for l in range(depth):
        x = decoder_layer(sd, f"decoder.decoder.layers.{l}", x, mem, heads)
    return x.astype(F32)


def cvae_sample(sd, c, eps=None, out_seq=90, depth=2, heads=4):
    """CVAE.sample (model_CVAE.py:44-46); eps=None <=> deterministic=True (reparameterize :81-87)."""
    mu, logvar = cvae_prior(sd, c, depth, heads)
    z = mu if eps is None else (mu + eps * np.exp(F32(0.5) * logvar)).astype(F32)
    return cvae_decode(sd, z, c, out_seq, depth, heads), mu, logvar
