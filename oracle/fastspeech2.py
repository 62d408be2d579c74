"""Oracle: FastSpeech2 forward / inference (torch-CPU fp32 restatement).  TEST INFRASTRUCTURE ONLY.

Follows, in the reference (/root/reference/parakeet):
  models/fastspeech2/fastspeech2.py        _forward :377-466, forward :289-375, inference :468-558,
                                           FastSpeech2Inference :662-671, FastSpeech2Loss :701-812
  modules/fastspeech2_transformer/encoder.py        Encoder.forward :171-192
  modules/fastspeech2_transformer/encoder_layer.py  EncoderLayer.forward :64-115 (normalize_before=True, no concat)
  modules/fastspeech2_transformer/attention.py      MultiHeadedAttention :51-156
  modules/fastspeech2_transformer/multi_layer_conv.py MultiLayeredConv1d.forward :62-77
  modules/fastspeech2_transformer/embedding.py      PositionalEncoding.extend_pe :46-62, ScaledPositionalEncoding :111-126
  modules/fastspeech2_predictor/duration_predictor.py  _forward :85-103
  modules/fastspeech2_predictor/variance_predictor.py  forward :77-104
  modules/fastspeech2_predictor/length_regulator.py    expand :46-66, forward :68-89
  modules/tacotron2/decoder.py              Postnet.forward :182-198
  modules/nets_utils.py                     make_pad_mask :54-93, make_non_pad_mask :96-125
  modules/masked_fill.py                    :28-37
  modules/layer_norm.py                     LayerNorm(dim=1) :47-63

Paddle semantics restated (see oracle/README.md): Linear.weight is [in, out]; Conv1D weight [out, in, k]; LayerNorm and
BatchNorm eps = 1e-5; BatchNorm buffers `_mean` / `_variance`; Embedding(padding_idx=0) outputs zeros for id 0;
paddle.round rounds half away from zero; dropout is identity (eval / p=0 parity runs).
Parameters: flat dict keyed by the reference's state-dict names.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

LJSPEECH_MODEL_CFG = dict(  # examples/fastspeech2/ljspeech/conf/default.yaml:33-75
    adim=384, aheads=2, elayers=4, eunits=1536, dlayers=4, dunits=1536,
    positionwise_layer_type="conv1d", positionwise_conv_kernel_size=3,
    duration_predictor_layers=2, duration_predictor_chans=256, duration_predictor_kernel_size=3,
    postnet_layers=5, postnet_filts=5, postnet_chans=256, use_scaled_pos_enc=True,
    encoder_normalize_before=True, decoder_normalize_before=True, reduction_factor=1,
    init_enc_alpha=1.0, init_dec_alpha=1.0,
    pitch_predictor_layers=5, pitch_predictor_chans=256, pitch_predictor_kernel_size=5, pitch_embed_kernel_size=1,
    energy_predictor_layers=2, energy_predictor_chans=256, energy_predictor_kernel_size=3, energy_embed_kernel_size=1,
)


# ----------------------------------------------------------------------------------------
# Dropout (paddle.nn.Dropout, upscale_in_train).  The reference draws its masks from Paddle's generator, which cannot be
# reproduced; what CAN be checked is the arithmetic around a GIVEN mask.  The CUDA path derives its masks from
# Philox4x32-10 keyed by (seed, step, site, element index) - include/parakeet_b200.h: pk_dropout - and this class restates
# that generator in numpy so that the oracle applies the very same masks at the reference's dropout sites.
# ----------------------------------------------------------------------------------------
def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10 (Salmon et al. 2011): uint32 arrays (or scalars) -> four uint32 arrays."""
    M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85
    c0, c1, c2, c3 = (np.asarray(v, dtype=np.uint64) & 0xFFFFFFFF for v in (c0, c1, c2, c3))
    k0, k1 = np.uint64(k0 & 0xFFFFFFFF), np.uint64(k1 & 0xFFFFFFFF)
    mask = np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0, p1 = np.uint64(M0) * c0, np.uint64(M1) * c2
        hi0, lo0, hi1, lo1 = p0 >> np.uint64(32), p0 & mask, p1 >> np.uint64(32), p1 & mask
        c0, c1, c2, c3 = (hi1 ^ c1 ^ k0) & mask, lo1, (hi0 ^ c3 ^ k1) & mask, lo0
        k0, k1 = (k0 + np.uint64(W0)) & mask, (k1 + np.uint64(W1)) & mask
    return c0, c1, c2, c3


def dropout_site(stack, layer, kind):
    """Site numbering shared with parakeet_b200/training/fs2_step.py.  stack: 0 encoder, 1 decoder, 2 pitch, 3 energy,
    4 duration predictor, 5 postnet; kind: 0 positional encoding, 1 attention probabilities, 2 attention sub-layer output,
    3 feed-forward hidden, 4 feed-forward sub-layer output, 5 predictor layer, 6 postnet layer."""
    return stack * 1000 + layer * 10 + kind


class PhiloxDropout:
    """drop(site, x, p) with x laid out exactly as the CUDA path lays the tensor out (the element index is the linear index of
    that contiguous tensor)."""

    def __init__(self, seed, step):
        self.seed, self.step = int(seed), int(step)

    def keep_mask(self, site, n, p):
        blocks = (n + 3) // 4
        idx = np.arange(blocks, dtype=np.uint64)
        w = philox4x32_10(idx & np.uint64(0xFFFFFFFF), idx >> np.uint64(32), np.full(blocks, site, dtype=np.uint64),
                          np.full(blocks, self.step, dtype=np.uint64), self.seed & 0xFFFFFFFF, (self.seed >> 32) & 0xFFFFFFFF)
        r = np.stack(w, axis=1).reshape(-1)[:n]
        t = p * 4294967296.0
        thresh = 0xFFFFFFFF if t >= 4294967295.0 else int(t)
        return torch.from_numpy((r >= np.uint64(thresh)).astype(np.float32))

    def __call__(self, site, x, p):
        if p <= 0:
            return x
        keep = self.keep_mask(site, x.numel(), p).reshape(x.shape)
        return x * keep * np.float32(1.0 / (1.0 - np.float32(p)))


def _drop(dropout, site, x, p):
    return x if dropout is None else dropout(site, x, p)


# the reference's constructor defaults (fastspeech2.py:85-113); the shipped yamls set 0.2 for the six transformer rates, 0.5 for the
# pitch / energy predictors and 0.0 for the two embed dropouts (examples/fastspeech2/*/conf/default.yaml:56-74)
DROPOUT_DEFAULTS = dict(transformer_enc_dropout_rate=0.1, transformer_enc_positional_dropout_rate=0.1, transformer_enc_attn_dropout_rate=0.1,
                        transformer_dec_dropout_rate=0.1, transformer_dec_positional_dropout_rate=0.1, transformer_dec_attn_dropout_rate=0.1,
                        duration_predictor_dropout_rate=0.1, pitch_predictor_dropout=0.5, energy_predictor_dropout=0.5,
                        postnet_dropout_rate=0.5, pitch_embed_dropout=0.5, energy_embed_dropout=0.5)
YAML_DROPOUT = dict(DROPOUT_DEFAULTS, transformer_enc_dropout_rate=0.2, transformer_enc_positional_dropout_rate=0.2,
                    transformer_enc_attn_dropout_rate=0.2, transformer_dec_dropout_rate=0.2, transformer_dec_positional_dropout_rate=0.2,
                    transformer_dec_attn_dropout_rate=0.2, pitch_embed_dropout=0.0, energy_embed_dropout=0.0)


def paddle_round(x):
    """paddle.round: half away from zero (C round), unlike torch.round (half to even)."""
    return torch.sign(x) * torch.floor(torch.abs(x) + 0.5)


def make_pad_mask(lengths, maxlen=None):
    """nets_utils.make_pad_mask :54-93 -> bool (B, maxlen), True on padding."""
    lengths = [int(v) for v in (lengths.tolist() if torch.is_tensor(lengths) else lengths)]
    maxlen = int(max(lengths)) if maxlen is None else maxlen
    seq = torch.arange(0, maxlen, dtype=torch.int64).unsqueeze(0).expand(len(lengths), maxlen)
    return seq >= torch.tensor(lengths, dtype=torch.int64).unsqueeze(-1)


def make_non_pad_mask(lengths, maxlen=None):
    return ~make_pad_mask(lengths, maxlen)


def masked_fill(xs, mask, value):
    """masked_fill.py:28-37."""
    return torch.where(mask.expand_as(xs) if mask.shape != xs.shape else mask, torch.full_like(xs, value), xs)


def positional_encoding(length, d_model):
    """embedding.py:46-62 (fp32 arithmetic, like paddle.arange(dtype=float32))."""
    position = torch.arange(0, length, dtype=torch.float32).unsqueeze(1)
    div_term = torch.exp(torch.arange(0, d_model, 2, dtype=torch.float32) * -(math.log(10000.0) / d_model))
    pe = torch.zeros(length, d_model)
    pe[:, 0::2] = torch.sin(position * div_term)
    pe[:, 1::2] = torch.cos(position * div_term)
    return pe.unsqueeze(0)


def linear(p, name, x):
    return x @ p[name + ".weight"] + p[name + ".bias"]  # Paddle Linear: weight [in, out]


def layer_norm(p, name, x):
    return F.layer_norm(x, (x.shape[-1],), p[name + ".weight"], p[name + ".bias"], eps=1e-5)


def attention(p, pre, x, mask, n_head, dropout=None, site=0, rate=0.0):
    """MultiHeadedAttention.forward (attention.py:133-156); query = key = value = x; mask (B,1,T) bool or None.
    Dropout on the attention probabilities (:124) is indexed in the CUDA layout (B*H, T, ceil64(T))."""
    B, T, A = x.shape
    dk = A // n_head
    q = linear(p, pre + "linear_q", x).reshape(B, T, n_head, dk).transpose(1, 2)
    k = linear(p, pre + "linear_k", x).reshape(B, T, n_head, dk).transpose(1, 2)
    v = linear(p, pre + "linear_v", x).reshape(B, T, n_head, dk).transpose(1, 2)
    scores = torch.matmul(q, k.transpose(-2, -1)) / math.sqrt(dk)
    if mask is not None:
        m = ~mask.unsqueeze(1)  # (B,1,1,T) True where padded key
        scores = masked_fill(scores, m, float(np.finfo(np.float32).min))
        attn = masked_fill(torch.softmax(scores, dim=-1), m, 0.0)
    else:
        attn = torch.softmax(scores, dim=-1)
    if dropout is not None and rate > 0:
        Tp = (T + 63) // 64 * 64
        padded = F.pad(attn, (0, Tp - T)).reshape(B * n_head, T, Tp)
        attn = dropout(site, padded, rate).reshape(B, n_head, T, Tp)[..., :T]
    ctx = torch.matmul(attn, v).transpose(1, 2).reshape(B, T, A)
    return linear(p, pre + "linear_out", ctx)


def conv1d_cl(p, name, x, bias=True):
    """Conv1D on a (B, T, C) tensor with 'same' padding (transpose - conv - transpose of the reference)."""
    w = p[name + ".weight"]
    k = w.shape[-1]
    return F.conv1d(x.transpose(1, 2), w, p[name + ".bias"] if bias else None, padding=(k - 1) // 2).transpose(1, 2)


def multi_layered_conv1d(p, pre, x, dropout=None, site=0, rate=0.0):
    """multi_layer_conv.py:62-77 (dropout between relu(w_1) and w_2, :76)."""
    return conv1d_cl(p, pre + "w_2", _drop(dropout, site, torch.relu(conv1d_cl(p, pre + "w_1", x)), rate))


def encoder_layer(p, pre, x, mask, n_head, dropout=None, stack=0, layer=0, rates=(0.0, 0.0)):
    """EncoderLayer.forward (encoder_layer.py:64-115), pre-LN, concat_after=False, cache=None.  rates = (dropout_rate of the
    layer: both sub-layer outputs :101,108 and the feed-forward hidden, attention_dropout_rate)."""
    r_layer, r_attn = rates
    a = attention(p, pre + "self_attn.", layer_norm(p, pre + "norm1", x), mask, n_head, dropout, dropout_site(stack, layer, 1), r_attn)
    x = x + _drop(dropout, dropout_site(stack, layer, 2), a, r_layer)
    f = multi_layered_conv1d(p, pre + "feed_forward.", layer_norm(p, pre + "norm2", x), dropout, dropout_site(stack, layer, 3), r_layer)
    x = x + _drop(dropout, dropout_site(stack, layer, 4), f, r_layer)
    return x


def encoder(p, pre, xs, masks, n_layers, n_head, embed=True, dropout=None, stack=0, rates=(0.0, 0.0, 0.0)):
    """Encoder.forward (encoder.py:171-192).  embed=True: Embedding(padding_idx=0)+ScaledPE; False: ScaledPE only.
    rates = (dropout_rate, positional_dropout_rate, attention_dropout_rate) as the reference's constructor names them."""
    if embed:
        w = p[pre + "embed.0.weight"]
        e = w[xs]
        e = torch.where((xs == 0).unsqueeze(-1), torch.zeros_like(e), e)  # padding_idx=0 -> zeros
        alpha = p[pre + "embed.1.alpha"]
        xs = e
    else:
        alpha = p[pre + "embed.0.alpha"]
    xs = xs + alpha * positional_encoding(xs.shape[1], xs.shape[2])  # ScaledPositionalEncoding.forward :111-126
    xs = _drop(dropout, dropout_site(stack, 0, 0), xs, rates[1])      # :126
    for i in range(n_layers):
        xs = encoder_layer(p, f"{pre}encoders.{i}.", xs, masks, n_head, dropout, stack, i, (rates[0], rates[2]))
    return layer_norm(p, pre + "after_norm", xs)


def predictor_stack(p, pre, xs, n_layers, dropout=None, stack=0, rate=0.0):
    """[Conv1D -> ReLU -> LayerNorm(channel) -> Dropout] x n (duration_predictor.py:69-83, variance_predictor.py:59-76)."""
    for i in range(n_layers):
        xs = torch.relu(conv1d_cl(p, f"{pre}conv.{i}.0", xs))
        xs = layer_norm(p, f"{pre}conv.{i}.2", xs)  # LayerNorm(n_chans, dim=1) == LN over channels in (B,T,C) view
        xs = _drop(dropout, dropout_site(stack, i, 5), xs, rate)
    return xs


def variance_predictor(p, pre, xs, x_masks, n_layers, dropout=None, stack=0, rate=0.0):
    """VariancePredictor.forward (variance_predictor.py:77-104); x_masks (B,T,1) True on padding."""
    xs = linear(p, pre + "linear", predictor_stack(p, pre, xs, n_layers, dropout, stack, rate))
    if x_masks is not None:
        xs = masked_fill(xs, x_masks, 0.0)
    return xs


def duration_predictor(p, pre, xs, x_masks, n_layers, is_inference, offset=1.0, dropout=None, rate=0.0):
    """DurationPredictor._forward (duration_predictor.py:85-103)."""
    xs = linear(p, pre + "linear", predictor_stack(p, pre, xs, n_layers, dropout, 4, rate)).squeeze(-1)
    if is_inference:
        xs = torch.clip(paddle_round(torch.exp(xs) - offset), min=0)
    if x_masks is not None:
        xs = masked_fill(xs, x_masks, 0.0)
    return xs


def length_regulator_expand(encodings, durations):
    """LengthRegulator.expand (length_regulator.py:46-66): numpy 0/1 matrix (float64) cast to x dtype, then matmul."""
    batch_size, t_enc = durations.shape
    durations = durations.numpy()
    slens = np.sum(durations, -1)
    t_dec = int(np.max(slens))
    M = np.zeros([batch_size, t_dec, t_enc])
    for i in range(batch_size):
        k = 0
        for j in range(t_enc):
            d = int(durations[i, j])
            if d >= 1:
                M[i, k:k + d, j] = 1
            k += d
    M = torch.tensor(M, dtype=encodings.dtype)
    return torch.matmul(M, encodings)


def length_regulator(xs, ds, alpha=1.0):
    """LengthRegulator.forward (length_regulator.py:68-89)."""
    if alpha != 1.0:
        assert alpha > 0
        ds = paddle_round(ds.to(torch.float32) * alpha)
    return length_regulator_expand(xs, ds.to(torch.int64))


def postnet(p, xs, n_layers, train_bn=False, new_stats=None, dropout=None, rate=0.0):
    """Postnet.forward (tacotron2/decoder.py:182-198) on (B, odim, T); dropout identity.  BatchNorm1D in eval mode uses the
    running statistics; with train_bn=True (model.train()) it uses the batch statistics (biased variance) and the updated
    running statistics (paddle momentum 0.9: running = 0.9 * running + 0.1 * batch, biased variance) go to `new_stats`."""
    for i in range(n_layers):
        w = p[f"postnet.postnet.{i}.0.weight"]
        xs = F.conv1d(xs, w, None, padding=(w.shape[-1] - 1) // 2)
        pre = f"postnet.postnet.{i}.1."
        if train_bn:
            mean = xs.mean(dim=(0, 2))
            var = xs.var(dim=(0, 2), unbiased=False)
            if new_stats is not None:
                new_stats[pre + "_mean"] = (0.9 * p[pre + "_mean"] + 0.1 * mean).detach()
                new_stats[pre + "_variance"] = (0.9 * p[pre + "_variance"] + 0.1 * var).detach()
            xs = (xs - mean[None, :, None]) / torch.sqrt(var[None, :, None] + 1e-5) * p[pre + "weight"][None, :, None] \
                + p[pre + "bias"][None, :, None]
            if i != n_layers - 1:
                xs = torch.tanh(xs)
            if dropout is not None and rate > 0:       # decoder.py:144-180: Dropout closes every postnet layer; CUDA layout (B, T, C)
                xs = dropout(dropout_site(5, i, 6), xs.transpose(1, 2).contiguous(), rate).transpose(1, 2)
            continue
        xs = F.batch_norm(xs, p[pre + "_mean"], p[pre + "_variance"], p[pre + "weight"], p[pre + "bias"], False, 0.0, 1e-5)
        if i != n_layers - 1:
            xs = torch.tanh(xs)
    return xs


def embedding_lookup(table, ids, padding_idx=0):
    """paddle nn.Embedding(padding_idx): rows of the table, zeros where ids == padding_idx."""
    e = table[ids]
    return torch.where((ids == padding_idx).unsqueeze(-1), torch.zeros_like(e), e)


def integrate_with_embed(p, name, itype, hs, embs):
    """_integrate_with_spk_embed / _integrate_with_tone_embed (fastspeech2.py:560-616).  F.normalize is paddle's default
    (p=2, axis=1, epsilon=1e-12): axis 1 is the feature axis of (B, D) speaker embeddings and of the (T, D) tone embeddings of
    `inference`, but the TIME axis of the (B, T, D) tone embeddings of the batched `forward` - restated as it is."""
    n = F.normalize(embs, p=2, dim=1, eps=1e-12)
    if name == "spk_projection":
        n = n.unsqueeze(1)                                                    # (B, 1, D)
    if itype == "add":
        return hs + linear(p, name, n)
    return linear(p, name, torch.cat([hs, n.expand(hs.shape[0] if n.dim() == 3 else -1, hs.shape[1], -1)], dim=-1))


def fs2_forward(p, cfg, xs, ilens, olens=None, ds=None, ps=None, es=None, is_inference=False, alpha=1.0,
                return_intermediates=False, train_bn=False, new_stats=None, stop_gradient_from_pitch_predictor=False,
                stop_gradient_from_energy_predictor=False, dropout=None, rates=None, spk_id=None, spembs=None, tone_id=None,
                tone_per_utterance=False):
    """FastSpeech2._forward (fastspeech2.py:377-466), single speaker, no tones.
    dropout: None (eval / p = 0) or a PhiloxDropout; rates: dict with the reference's constructor keywords (DROPOUT_DEFAULTS).

    xs (B,Tmax) int64; ilens (B,); training: olens (B,), ds (B,Tmax) int64, ps/es (B,Tmax,1).
    Returns before_outs, after_outs, d_outs, p_outs, e_outs.
    """
    cfg = {**LJSPEECH_MODEL_CFG, **(cfg or {})}
    nh = cfg["aheads"]
    x_masks = make_non_pad_mask(ilens, xs.shape[1]).unsqueeze(-2)          # _source_mask :618-641
    R = {**DROPOUT_DEFAULTS, **(rates or {})}
    hs = encoder(p, "encoder.", xs, x_masks, cfg["elayers"], nh, embed=True, dropout=dropout, stack=0,
                 rates=(R["transformer_enc_dropout_rate"], R["transformer_enc_positional_dropout_rate"], R["transformer_enc_attn_dropout_rate"]))
    if "spk_projection.weight" in p and (spembs is not None or spk_id is not None):                     # fastspeech2.py:395-401
        emb = spembs if spembs is not None else embedding_lookup(p["spk_embedding_table.weight"], spk_id)
        hs = integrate_with_embed(p, "spk_projection", cfg.get("spk_embed_integration_type", "add"), hs, emb)
    if "tone_projection.weight" in p and tone_id is not None:                                            # :403-407
        emb = embedding_lookup(p["tone_embedding_table.weight"], tone_id)
        if tone_per_utterance:       # `inference` passes (T,) ids -> (T, D) embeddings: axis 1 is the feature axis there
            emb = torch.stack([F.normalize(e, p=2, dim=1, eps=1e-12) for e in emb])
            ty = cfg.get("tone_embed_integration_type", "add")
            hs = hs + linear(p, "tone_projection", emb) if ty == "add" else linear(p, "tone_projection", torch.cat([hs, emb], dim=-1))
        else:
            hs = integrate_with_embed(p, "tone_projection", cfg.get("tone_embed_integration_type", "add"), hs, emb)
    d_masks = make_pad_mask(ilens, xs.shape[1])
    # fastspeech2.py:412-419 (stop_gradient_from_*_predictor -> hs.detach())
    p_outs = variance_predictor(p, "pitch_predictor.", hs.detach() if stop_gradient_from_pitch_predictor else hs,
                                d_masks.unsqueeze(-1), cfg["pitch_predictor_layers"], dropout, 2, R["pitch_predictor_dropout"])
    e_outs = variance_predictor(p, "energy_predictor.", hs.detach() if stop_gradient_from_energy_predictor else hs,
                                d_masks.unsqueeze(-1), cfg["energy_predictor_layers"], dropout, 3, R["energy_predictor_dropout"])
    if is_inference:
        d_outs = duration_predictor(p, "duration_predictor.", hs, d_masks, cfg["duration_predictor_layers"], True)
        p_embs = conv1d_cl(p, "pitch_embed.0", p_outs)
        e_embs = conv1d_cl(p, "energy_embed.0", e_outs)
        hs = hs + e_embs + p_embs
        hs_lr = length_regulator(hs, d_outs, alpha)
    else:
        d_outs = duration_predictor(p, "duration_predictor.", hs, d_masks, cfg["duration_predictor_layers"], False, dropout=dropout,
                                    rate=R["duration_predictor_dropout_rate"])
        p_embs = _drop(dropout, dropout_site(6, 0, 7), conv1d_cl(p, "pitch_embed.0", ps), R["pitch_embed_dropout"])     # :220-247
        e_embs = _drop(dropout, dropout_site(6, 0, 8), conv1d_cl(p, "energy_embed.0", es), R["energy_embed_dropout"])
        hs = hs + e_embs + p_embs
        hs_lr = length_regulator(hs, ds)
    if olens is not None and not is_inference:
        h_masks = make_non_pad_mask(olens, hs_lr.shape[1]).unsqueeze(-2)
    else:
        h_masks = None
    zs = encoder(p, "decoder.", hs_lr, h_masks, cfg["dlayers"], nh, embed=False, dropout=dropout, stack=1,
                 rates=(R["transformer_dec_dropout_rate"], R["transformer_dec_positional_dropout_rate"], R["transformer_dec_attn_dropout_rate"]))
    before_outs = linear(p, "feat_out", zs).reshape(zs.shape[0], -1, p["feat_out.bias"].shape[0])
    if cfg["postnet_layers"] == 0:
        after_outs = before_outs
    else:
        after_outs = before_outs + postnet(p, before_outs.transpose(1, 2), cfg["postnet_layers"], train_bn, new_stats, dropout,
                                           R["postnet_dropout_rate"]).transpose(1, 2)
    if return_intermediates:
        return before_outs, after_outs, d_outs, p_outs, e_outs, dict(hs=hs, hs_lr=hs_lr, zs=zs)
    return before_outs, after_outs, d_outs, p_outs, e_outs


def fs2_inference(p, cfg, text, alpha=1.0):
    """FastSpeech2.inference (fastspeech2.py:468-558), no teacher forcing: (T,) ids -> (L, odim)."""
    xs = text.to(torch.int64).unsqueeze(0)
    ilens = torch.tensor([xs.shape[1]], dtype=torch.int64)
    _, outs, *_ = fs2_forward(p, cfg, xs, ilens, is_inference=True, alpha=alpha)
    return outs[0]


def fs2_inference_denorm(p, cfg, text, mu, sigma):
    """FastSpeech2Inference.forward (fastspeech2.py:662-671): inference then ZScore.inverse (normalizer.py:30-33)."""
    return fs2_inference(p, cfg, text) * sigma + mu


def fs2_loss(after_outs, before_outs, d_outs, p_outs, e_outs, ys, ds, ps, es, ilens, olens):
    """FastSpeech2Loss.forward (fastspeech2.py:701-812), use_masking=True, use_weighted_masking=False."""
    out_masks = make_non_pad_mask(olens, ys.shape[1]).unsqueeze(-1)
    before_m = before_outs.masked_select(out_masks.expand_as(before_outs))
    after_m = after_outs.masked_select(out_masks.expand_as(after_outs))
    ys_m = ys.masked_select(out_masks.expand_as(ys))
    dmask = make_non_pad_mask(ilens, ds.shape[1])
    d_m = d_outs.masked_select(dmask)
    ds_m = ds.masked_select(dmask)
    pmask = dmask.unsqueeze(-1)
    p_m, e_m = p_outs.masked_select(pmask), e_outs.masked_select(pmask)
    ps_m, es_m = ps.masked_select(pmask), es.masked_select(pmask)
    l1 = F.l1_loss(before_m, ys_m) + F.l1_loss(after_m, ys_m)
    dur = F.mse_loss(d_m, torch.log(ds_m.to(torch.float32) + 1.0))  # DurationPredictorLoss :140-184
    return l1, dur, F.mse_loss(p_m, ps_m), F.mse_loss(e_m, es_m)


# ----------------------------------------------------------------------------------------
# Synthetic, seeded parameters and inputs
# ----------------------------------------------------------------------------------------
def synth_params(seed=1, idim=80, odim=80, cfg=None, target_dur=(2.0, 12.0)):
    cfg = {**LJSPEECH_MODEL_CFG, **(cfg or {})}
    g = torch.Generator().manual_seed(seed)
    p = {}
    A = cfg["adim"]

    def xavier(*shape, fan_in, fan_out):
        bound = math.sqrt(6.0 / (fan_in + fan_out))
        return (torch.rand(*shape, generator=g) * 2 - 1) * bound

    def lin(name, i, o, bias_scale=0.02):
        p[name + ".weight"] = xavier(i, o, fan_in=i, fan_out=o)
        p[name + ".bias"] = torch.randn(o, generator=g) * bias_scale  # perturbed (not zero) so bias bugs show

    def conv(name, o, i, k, bias=True):
        p[name + ".weight"] = xavier(o, i, k, fan_in=i * k, fan_out=o * k)
        if bias:
            p[name + ".bias"] = torch.randn(o, generator=g) * 0.02

    def ln(name, c):
        p[name + ".weight"] = 1 + 0.1 * torch.randn(c, generator=g)
        p[name + ".bias"] = 0.1 * torch.randn(c, generator=g)

    def enc(pre, layers, units, k):
        for i in range(layers):
            q = f"{pre}encoders.{i}."
            for nm in ("linear_q", "linear_k", "linear_v", "linear_out"):
                lin(q + "self_attn." + nm, A, A)
            conv(q + "feed_forward.w_1", units, A, k)
            conv(q + "feed_forward.w_2", A, units, k)
            ln(q + "norm1", A)
            ln(q + "norm2", A)
        ln(pre + "after_norm", A)

    emb = torch.randn(idim, A, generator=g) * 0.3
    p["encoder.embed.0.weight"] = emb  # row 0 deliberately non-zero: padding_idx must still yield zeros
    p["encoder.embed.1.alpha"] = torch.tensor([cfg["init_enc_alpha"]])
    enc("encoder.", cfg["elayers"], cfg["eunits"], cfg["positionwise_conv_kernel_size"])
    p["decoder.embed.0.alpha"] = torch.tensor([cfg["init_dec_alpha"]])
    enc("decoder.", cfg["dlayers"], cfg["dunits"], cfg["positionwise_conv_kernel_size"])

    def pred(pre, layers, chans, k):
        for i in range(layers):
            conv(f"{pre}conv.{i}.0", chans, A if i == 0 else chans, k)
            ln(f"{pre}conv.{i}.2", chans)
        lin(pre + "linear", chans, 1)

    pred("duration_predictor.", cfg["duration_predictor_layers"], cfg["duration_predictor_chans"],
         cfg["duration_predictor_kernel_size"])
    pred("pitch_predictor.", cfg["pitch_predictor_layers"], cfg["pitch_predictor_chans"], cfg["pitch_predictor_kernel_size"])
    pred("energy_predictor.", cfg["energy_predictor_layers"], cfg["energy_predictor_chans"],
         cfg["energy_predictor_kernel_size"])
    # durations land in ~[2, 12]: exp(out) - 1 with out ~ N(log 7, 0.35)
    lo, hi = target_dur
    p["duration_predictor.linear.weight"] *= 0.6
    p["duration_predictor.linear.bias"] = torch.tensor([math.log((lo + hi) / 2 + 1.0)])
    conv("pitch_embed.0", A, 1, cfg["pitch_embed_kernel_size"])
    conv("energy_embed.0", A, 1, cfg["energy_embed_kernel_size"])
    lin("feat_out", A, odim)
    n_post, ch, kf = cfg["postnet_layers"], cfg["postnet_chans"], cfg["postnet_filts"]
    for i in range(n_post):
        ic = odim if i == 0 else ch
        oc = odim if i == n_post - 1 else ch
        conv(f"postnet.postnet.{i}.0", oc, ic, kf, bias=False)
        q = f"postnet.postnet.{i}.1."
        p[q + "weight"] = 1 + 0.1 * torch.randn(oc, generator=g)
        p[q + "bias"] = 0.1 * torch.randn(oc, generator=g)
        p[q + "_mean"] = 0.1 * torch.randn(oc, generator=g)
        p[q + "_variance"] = 1 + 0.1 * torch.rand(oc, generator=g)
    return p


def add_speaker_tone_params(p, seed, adim=384, num_speakers=6, spk_embed_dim=256, spk_type="concat", num_tones=7, tone_embed_dim=32,
                            tone_type="add"):
    """Speaker / tone tables and projections in the reference's key names and Paddle layouts (Linear [in, out])."""
    g = torch.Generator().manual_seed(seed + 7000)
    q = dict(p)
    q["spk_embedding_table.weight"] = torch.randn(num_speakers, spk_embed_dim, generator=g)
    i = spk_embed_dim if spk_type == "add" else adim + spk_embed_dim
    q["spk_projection.weight"] = (torch.rand(i, adim, generator=g) * 2 - 1) * math.sqrt(6.0 / (i + adim))
    q["spk_projection.bias"] = torch.randn(adim, generator=g) * 0.02
    q["tone_embedding_table.weight"] = torch.randn(num_tones, tone_embed_dim, generator=g)
    i = tone_embed_dim if tone_type == "add" else adim + tone_embed_dim
    q["tone_projection.weight"] = (torch.rand(i, adim, generator=g) * 2 - 1) * math.sqrt(6.0 / (i + adim))
    q["tone_projection.bias"] = torch.randn(adim, generator=g) * 0.02
    return q


def synth_text(seed, lengths, idim=80):
    """ids ~ U{1..V-2}, zero-padded to max length (data/batch.py:170-189 batch_sequences)."""
    g = torch.Generator().manual_seed(seed + 2000)
    B, Tmax = len(lengths), max(lengths)
    xs = torch.zeros(B, Tmax, dtype=torch.int64)
    for b, n in enumerate(lengths):
        xs[b, :n] = torch.randint(1, idim - 1, (n,), generator=g)
    return xs, torch.tensor(lengths, dtype=torch.int64)


def synth_train_batch(seed, lengths, odim=80, idim=80, dur_range=(2, 12)):
    """cfg5-style teacher-forced batch: durations U{2..12}, speech N(0,1) zero-padded, pitch/energy N(0,1)."""
    g = torch.Generator().manual_seed(seed + 3000)
    xs, ilens = synth_text(seed, lengths, idim)
    B, Tmax = xs.shape
    ds = torch.zeros(B, Tmax, dtype=torch.int64)
    for b, n in enumerate(lengths):
        ds[b, :n] = torch.randint(dur_range[0], dur_range[1] + 1, (n,), generator=g)
    olens = ds.sum(1)
    Lmax = int(olens.max())
    ys = torch.zeros(B, Lmax, odim)
    ps = torch.zeros(B, Tmax, 1)
    es = torch.zeros(B, Tmax, 1)
    for b, n in enumerate(lengths):
        ys[b, :int(olens[b])] = torch.randn(int(olens[b]), odim, generator=g)
        ps[b, :n] = torch.randn(n, 1, generator=g)
        es[b, :n] = torch.randn(n, 1, generator=g)
    return dict(text=xs, text_lengths=ilens, speech=ys, speech_lengths=olens, durations=ds, pitch=ps, energy=es)


# ----------------------------------------------------------------------------------------
# Training step restatement (FastSpeech2Updater.update_core, fastspeech2_updater.py:51-99): forward in train mode with
# dropout off, FastSpeech2Loss, torch autograd for loss.backward(), paddle.optimizer.Adam semantics.
# ----------------------------------------------------------------------------------------
BUFFER_SUFFIXES = ("_mean", "_variance")


def train_step_grads(p, cfg, batch, stop_gradient_from_pitch_predictor=True, stop_gradient_from_energy_predictor=False,
                     dropout=None, rates=None):
    """-> (losses dict, grads dict keyed like p (trainable tensors only), new BN running stats).
    dropout: None (all rates 0, the round-1 behaviour) or PhiloxDropout(seed, step) with `rates` (YAML_DROPOUT ...)."""
    q = {k: (v.clone().requires_grad_(True) if not k.endswith(BUFFER_SUFFIXES) else v.clone()) for k, v in p.items()}
    new_stats = {}
    out = fs2_forward(q, cfg, batch["text"], batch["text_lengths"], batch["speech_lengths"], batch["durations"], batch["pitch"],
                      batch["energy"], train_bn=True, new_stats=new_stats,
                      stop_gradient_from_pitch_predictor=stop_gradient_from_pitch_predictor,
                      stop_gradient_from_energy_predictor=stop_gradient_from_energy_predictor, dropout=dropout, rates=rates)
    l1, dur, pitch, energy = fs2_loss(out[1], out[0], out[2], out[3], out[4], batch["speech"], batch["durations"], batch["pitch"],
                                      batch["energy"], batch["text_lengths"], batch["speech_lengths"])
    loss = l1 + dur + pitch + energy                                           # fastspeech2_updater.py:83
    loss.backward()
    grads = {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in q.items() if not k.endswith(BUFFER_SUFFIXES)}
    losses = {k: float(v.detach()) for k, v in dict(l1_loss=l1, duration_loss=dur, pitch_loss=pitch, energy_loss=energy, loss=loss).items()}
    return losses, grads, new_stats


def adam_step(p, grads, state, lr=1e-3, beta1=0.9, beta2=0.999, eps=1e-8):
    """paddle.optimizer.Adam (training/optimizer.py:17-46): lr_t = lr*sqrt(1-b2^t)/(1-b1^t); p -= lr_t*m/(sqrt(v)+eps*sqrt(1-b2^t))."""
    state["t"] = state.get("t", 0) + 1
    t = state["t"]
    c1, c2 = 1 - beta1 ** t, math.sqrt(1 - beta2 ** t)
    out = dict(p)
    for k, g in grads.items():
        m = state.setdefault("m." + k, torch.zeros_like(g))
        v = state.setdefault("v." + k, torch.zeros_like(g))
        m.mul_(beta1).add_(g, alpha=1 - beta1)
        v.mul_(beta2).addcmul_(g, g, value=1 - beta2)
        out[k] = p[k] - (lr * c2 / c1) * m / (v.sqrt() + eps * c2)
    return out
