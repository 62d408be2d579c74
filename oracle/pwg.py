"""Oracle: Parallel WaveGAN generator forward (torch-CPU fp32 restatement).

Follows parakeet/models/parallel_wavegan/parallel_wavegan.py of the reference:
  Stretch2D.forward          :48-63     nearest interpolate (integer scale = repeat)
  UpsampleNet.forward        :119-138   [stretch, Conv2D(1,1,(1,2s+1),pad (0,s), no bias)] x scales
  ConvInUpsampleNet.forward  :201-216   Conv1D(aux,aux,k=2w+1,no pad,no bias) then UpsampleNet
  ResidualBlock.forward      :284-315
  PWGGenerator.forward       :445-472
  PWGGenerator.inference     :498-520   (noise drawn by the caller here: RNG streams cannot match)
  apply/remove_weight_norm   :474-496   paddle weight_norm(dim=0): w = g * v / ||v||, g is 1-D [out]

Parameters are a flat dict with the reference's state-dict key names
(SURVEY.md 8b / Appendix A), conv weights [out, in, k] like Paddle.
TEST INFRASTRUCTURE ONLY - see oracle/__init__.py.
"""
import math

import torch
import torch.nn.functional as F

DEFAULT_GENERATOR_PARAMS = dict(  # examples/GANVocoder/parallelwave_gan/baker/conf/default.yaml:23-45
    in_channels=1, out_channels=1, kernel_size=3, layers=30, stacks=3,
    residual_channels=64, gate_channels=128, skip_channels=64, aux_channels=80,
    aux_context_window=2, dropout=0.0, bias=True, use_weight_norm=True,
    use_causal_conv=False, upsample_scales=[4, 5, 3, 5], interpolate_mode="nearest",
    freq_axis_kernel_size=1, nonlinear_activation=None, nonlinear_activation_params={})


def fold_weight_norm(params):
    """remove_weight_norm (parallel_wavegan.py:485-496): weight = g * v / ||v||_2 over all dims but 0."""
    out = {}
    for k, v in params.items():
        if k.endswith("weight_g"):
            continue
        if k.endswith("weight_v"):
            g = params[k[:-1] + "g"]
            norm = v.reshape(v.shape[0], -1).norm(dim=1)
            shape = [-1] + [1] * (v.dim() - 1)
            out[k[:-2]] = v * (g / norm).reshape(shape)
        else:
            out[k] = v
    return out


def upsample_net(params, c, scales, prefix="upsample_net.upsample."):
    """UpsampleNet.forward (parallel_wavegan.py:119-138), non-causal, no activation."""
    c = c.unsqueeze(1)  # (B,1,F,T)
    for i, s in enumerate(scales):
        c = F.interpolate(c, scale_factor=(1, s), mode="nearest")  # Stretch2D :61-62
        w = params[f"{prefix}up_layers.{2 * i + 1}.weight"]       # [1,1,1,2s+1]
        c = F.conv2d(c, w, padding=(0, s))
    return c.squeeze(1)


def conv_in_upsample_net(params, c, scales):
    """ConvInUpsampleNet.forward (parallel_wavegan.py:201-216), non-causal."""
    c = F.conv1d(c, params["upsample_net.conv_in.weight"])  # k = 2w+1, no padding, no bias
    return upsample_net(params, c, scales)


def residual_block(params, prefix, x, c, dilation):
    """ResidualBlock.forward (parallel_wavegan.py:284-315), dropout 0, non-causal."""
    x_input = x
    w = params[prefix + "conv.weight"]
    k = w.shape[-1]
    x = F.conv1d(x, w, params.get(prefix + "conv.bias"), padding=(k - 1) // 2 * dilation, dilation=dilation)
    if c is not None:
        x = x + F.conv1d(c, params[prefix + "conv1x1_aux.weight"])
    a, b = torch.chunk(x, 2, dim=1)
    x = torch.tanh(a) * torch.sigmoid(b)
    skip = F.conv1d(x, params[prefix + "conv1x1_skip.weight"], params.get(prefix + "conv1x1_skip.bias"))
    res = (F.conv1d(x, params[prefix + "conv1x1_out.weight"], params.get(prefix + "conv1x1_out.bias")) + x_input) \
        * math.sqrt(0.5)
    return res, skip


def generator_forward(params, x, c, cfg=None, return_intermediates=False):
    """PWGGenerator.forward (parallel_wavegan.py:445-472).

    x: (B, 1, T) noise; c: (B, aux, T' + 2*aux_context_window); returns (B, 1, T).
    """
    cfg = {**DEFAULT_GENERATOR_PARAMS, **(cfg or {})}
    assert not cfg["use_causal_conv"], "causal branch is out of scope (reference bug at :305)"
    assert cfg["nonlinear_activation"] is None
    layers, stacks = cfg["layers"], cfg["stacks"]
    lps = layers // stacks
    c_up = conv_in_upsample_net(params, c, cfg["upsample_scales"])
    assert c_up.shape[-1] == x.shape[-1], (c_up.shape, x.shape)
    x = F.conv1d(x, params["first_conv.weight"], params["first_conv.bias"])
    skips = 0
    inter = []
    for i in range(layers):
        x, s = residual_block(params, f"conv_layers.{i}.", x, c_up, 2 ** (i % lps))
        skips = skips + s
        if return_intermediates:
            inter.append(x)
    skips = skips * math.sqrt(1.0 / layers)
    y = F.relu(skips)
    y = F.conv1d(y, params["last_conv_layers.1.weight"], params["last_conv_layers.1.bias"])
    y = F.relu(y)
    y = F.conv1d(y, params["last_conv_layers.3.weight"], params["last_conv_layers.3.bias"])
    if return_intermediates:
        return y, dict(c_up=c_up, x_layers=inter, skips=skips)
    return y


def generator_inference(params, c, noise, cfg=None):
    """PWGGenerator.inference (parallel_wavegan.py:498-520) with caller-supplied noise.

    c: (T', aux) normalised log-mel; noise: (1, 1, T'*hop); returns (T, out_channels).
    """
    cfg = {**DEFAULT_GENERATOR_PARAMS, **(cfg or {})}
    w = cfg["aux_context_window"]
    c = c.transpose(0, 1).unsqueeze(0)
    c = F.pad(c, (w, w), mode="replicate")  # nn.Pad1D(w, mode='replicate') :518
    out = generator_forward(params, noise, c, cfg)
    return out.squeeze(0).transpose(0, 1)


def pwg_inference(params, logmel, mu, sigma, noise, cfg=None):
    """PWGInference.forward (parallel_wavegan.py:766-775): ZScore then generator.inference."""
    return generator_inference(params, (logmel - mu) / sigma, noise, cfg)


# ----------------------------------------------------------------------------------------
# Synthetic, seeded parameters: U(-1/sqrt(fan_in), 1/sqrt(fan_in)), split g/v.
# ----------------------------------------------------------------------------------------
def synth_params(seed=2, cfg=None, weight_norm=False):
    cfg = {**DEFAULT_GENERATOR_PARAMS, **(cfg or {})}
    g = torch.Generator().manual_seed(seed)
    p = {}

    def conv(name, out_c, in_c, *k, bias=True, scale=1.0):
        fan_in = in_c * math.prod(k)
        bound = scale / math.sqrt(fan_in)
        p[name + ".weight"] = (torch.rand(out_c, in_c, *k, generator=g) * 2 - 1) * bound
        if bias:
            p[name + ".bias"] = (torch.rand(out_c, generator=g) * 2 - 1) * bound

    R, G, S, A = cfg["residual_channels"], cfg["gate_channels"], cfg["skip_channels"], cfg["aux_channels"]
    conv("first_conv", R, cfg["in_channels"], 1)
    conv("upsample_net.conv_in", A, A, 2 * cfg["aux_context_window"] + 1, bias=False, scale=1.7)
    for i, s in enumerate(cfg["upsample_scales"]):
        # positive-ish smoothing FIR like a trained upsampler (keeps c at O(1) scale)
        w = torch.rand(1, 1, 1, 2 * s + 1, generator=g)
        p[f"upsample_net.upsample.up_layers.{2 * i + 1}.weight"] = w / w.sum()
    for i in range(cfg["layers"]):
        pre = f"conv_layers.{i}."
        conv(pre + "conv", G, R, cfg["kernel_size"], bias=cfg["bias"], scale=1.7)
        conv(pre + "conv1x1_aux", G, A, 1, bias=False, scale=1.7)
        conv(pre + "conv1x1_out", R, G // 2, 1, bias=cfg["bias"], scale=3.0)
        conv(pre + "conv1x1_skip", S, G // 2, 1, bias=cfg["bias"], scale=3.0)
    conv("last_conv_layers.1", S, S, 1)
    conv("last_conv_layers.3", cfg["out_channels"], S, 1)
    if weight_norm:
        q = {}
        for k, v in p.items():
            if k.endswith(".weight"):
                norm = v.reshape(v.shape[0], -1).norm(dim=1)
                gg = norm * (0.5 + torch.rand(v.shape[0], generator=g))  # g != ||v|| so folding is exercised
                q[k + "_g"] = gg                                       # 1-D [out] (tests/unit/test_pwg.py:131-132)
                q[k + "_v"] = v
            else:
                q[k] = v
        return q
    return p


def synth_inputs(seed=2, batch=32, mel_frames=400, cfg=None):
    """cfg2: mel N(0,1) (B, 80, T'+2w) (replicate-padded), noise N(0,1) (B,1,T'*hop)."""
    cfg = {**DEFAULT_GENERATOR_PARAMS, **(cfg or {})}
    g = torch.Generator().manual_seed(seed + 1000)
    w = cfg["aux_context_window"]
    hop = math.prod(cfg["upsample_scales"])
    mel = torch.randn(batch, cfg["aux_channels"], mel_frames, generator=g)
    c = F.pad(mel, (w, w), mode="replicate")
    x = torch.randn(batch, cfg["in_channels"], mel_frames * hop, generator=g)
    return x, c


# ----------------------------------------------------------------------------------------
# Discriminator and the GAN training step (SURVEY.md 8f.1 - the "next" row after the generator forward).
# ----------------------------------------------------------------------------------------
DEFAULT_DISCRIMINATOR_PARAMS = dict(  # examples/GANVocoder/parallelwave_gan/baker/conf/default.yaml:50-60
    in_channels=1, out_channels=1, kernel_size=3, layers=10, conv_channels=64, dilation_factor=1, bias=True,
    nonlinear_activation="LeakyReLU", nonlinear_activation_params={"negative_slope": 0.2}, use_weight_norm=True)


def discriminator_dilations(cfg=None):
    """Dilation of conv i (parallel_wavegan.py:571-576): 1 for i = 0, else i (dilation_factor 1) or factor**i; last conv 1."""
    cfg = {**DEFAULT_DISCRIMINATOR_PARAMS, **(cfg or {})}
    f = cfg["dilation_factor"]
    return [1 if i == 0 else (i if f == 1 else f ** i) for i in range(cfg["layers"] - 1)] + [1]


def discriminator_forward(params, x, cfg=None):
    """PWGDiscriminator.forward (parallel_wavegan.py:554-614): (N, 1, T) audio -> (N, 1, T) logits.
    layers-1 x [Conv1D(k, dilation d_i, 'same' zero padding) + LeakyReLU(0.2)] + Conv1D(k); weights folded (w = g v/|v|)."""
    cfg = {**DEFAULT_DISCRIMINATOR_PARAMS, **(cfg or {})}
    k = cfg["kernel_size"]
    slope = cfg["nonlinear_activation_params"]["negative_slope"]
    dil = discriminator_dilations(cfg)
    n = cfg["layers"]
    for i in range(n):
        pre = f"conv_layers.{2 * i}."                                  # Sequential index: conv at 2i, activation at 2i+1
        x = F.conv1d(x, params[pre + "weight"], params.get(pre + "bias"), padding=(k - 1) // 2 * dil[i], dilation=dil[i])
        if i < n - 1:
            x = F.leaky_relu(x, slope)
    return x


def synth_discriminator_params(seed=12, cfg=None):
    cfg = {**DEFAULT_DISCRIMINATOR_PARAMS, **(cfg or {})}
    g = torch.Generator().manual_seed(seed)
    p = {}
    cin = cfg["in_channels"]
    for i in range(cfg["layers"]):
        cout = cfg["out_channels"] if i == cfg["layers"] - 1 else cfg["conv_channels"]
        bound = 1.7 / math.sqrt(cin * cfg["kernel_size"])
        p[f"conv_layers.{2 * i}.weight"] = (torch.rand(cout, cin, cfg["kernel_size"], generator=g) * 2 - 1) * bound
        p[f"conv_layers.{2 * i}.bias"] = (torch.rand(cout, generator=g) * 2 - 1) * bound
        cin = cfg["conv_channels"]
    return p


def gan_step_losses(gen_params, dis_params, noise, mel, wav, adversarial=True, lambda_adv=4.0, gen_cfg=None, dis_cfg=None):
    """PWGUpdater.update_core (parallel_wavegan_updater.py:76-153) as two differentiable scalars.

    generator loss  = sc + mag (MultiResolutionSTFTLoss of G(noise, mel) vs wav) [+ lambda_adv * MSE(D(G(.)), 1)]
    discriminator loss = MSE(D(wav), 1) + MSE(D(G(.).detach()), 0)            (only once the discriminator trains)
    wav, noise: (B, 1, T); mel: (B, 80, T' + 2w).  Returns a dict of tensors (call .backward() on the two losses)."""
    from . import stft as ostft
    wav_ = generator_forward(gen_params, noise, mel, gen_cfg)
    sc, mag = ostft.multi_resolution_stft_loss(wav_.squeeze(1), wav.squeeze(1))
    out = dict(wav_=wav_, spectral_convergence_loss=sc, log_stft_magnitude_loss=mag)
    gen_loss = sc + mag
    if adversarial:
        p_ = discriminator_forward(dis_params, wav_, dis_cfg)
        adv = F.mse_loss(p_, torch.ones_like(p_))
        out["adversarial_loss"] = adv
        gen_loss = gen_loss + lambda_adv * adv
        p = discriminator_forward(dis_params, wav, dis_cfg)
        pf = discriminator_forward(dis_params, wav_.detach(), dis_cfg)
        out["real_loss"] = F.mse_loss(p, torch.ones_like(p))
        out["fake_loss"] = F.mse_loss(pf, torch.zeros_like(pf))
        out["discriminator_loss"] = out["real_loss"] + out["fake_loss"]
    out["generator_loss"] = gen_loss
    return out
