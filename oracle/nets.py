"""Oracle: score-network forwards (torch CPU fp32, functional).  TEST INFRASTRUCTURE ONLY.

Functional restatements operating directly on a reference ``state_dict`` (same key
names), so real reference weights can be fed to both sides:
  * ``unet_forward``  -- ``UNetModel.forward`` (``dlpm/models/unet.py:463-492``) with ResBlock
    ``_forward`` (:182-195, use_scale_shift_norm=True as built at ``dlpm_experiment.py:41-56``),
    AttentionBlock/QKVAttention (:198-250), Up/Downsample (:48-102), ``timestep_embedding``
    (``nn.py:103-121``), GroupNorm32 with min(32, C) groups (``unet.py:141,153,212,433``).
  * ``mlp_forward``   -- ``MLPModel.forward`` (``dlpm/models/Model.py:148-211``) for the only
    configuration that runs (no_a=True, learnable time embedding) with
    ``DiffusionBlockConditioned.forward`` (``DiffusionBlocks.py:125-136``).
Pinned against the imported reference modules by ``tests/test_oracle_golden.py``.
"""
import math

import torch
import torch.nn.functional as F


def timestep_embedding(timesteps, dim, max_period=10000):
    """nn.py:103-121."""
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(0, half, dtype=torch.float32) / half)
    args = timesteps[:, None].float() * freqs[None]
    emb = torch.cat([torch.cos(args), torch.sin(args)], dim=-1)
    if dim % 2:
        emb = torch.cat([emb, torch.zeros_like(emb[:, :1])], dim=-1)
    return emb


def _silu(x):
    return x * torch.sigmoid(x)


def _gn(x, sd, prefix, C):
    return F.group_norm(x.float(), min(32, C), sd[prefix + ".weight"], sd[prefix + ".bias"], eps=1e-5)


def _resblock(sd, p, x, emb):
    """unet.py:182-195 (scale-shift norm)."""
    cin = x.shape[1]
    h = F.conv2d(_silu(_gn(x, sd, p + ".in_layers.0", cin)), sd[p + ".in_layers.2.weight"],
                 sd[p + ".in_layers.2.bias"], padding=1)
    cout = h.shape[1]
    emb_out = F.linear(_silu(emb), sd[p + ".emb_layers.1.weight"], sd[p + ".emb_layers.1.bias"])
    scale, shift = torch.chunk(emb_out[..., None, None], 2, dim=1)
    h = _gn(h, sd, p + ".out_layers.0", cout) * (1 + scale) + shift
    h = F.conv2d(_silu(h), sd[p + ".out_layers.3.weight"], sd[p + ".out_layers.3.bias"], padding=1)
    if p + ".skip_connection.weight" in sd:
        w = sd[p + ".skip_connection.weight"]
        x = F.conv2d(x, w, sd[p + ".skip_connection.bias"], padding=w.shape[-1] // 2)
    return x + h


def _attention(sd, p, x, num_heads):
    """unet.py:220-250."""
    b, c, *spatial = x.shape
    x = x.reshape(b, c, -1)
    xn = F.group_norm(x.float(), min(32, c), sd[p + ".norm.weight"], sd[p + ".norm.bias"], eps=1e-5)
    qkv = F.conv1d(xn, sd[p + ".qkv.weight"], sd[p + ".qkv.bias"])
    qkv = qkv.reshape(b * num_heads, -1, qkv.shape[2])
    ch = qkv.shape[1] // 3
    q, k, v = torch.split(qkv, ch, dim=1)
    scale = 1 / math.sqrt(math.sqrt(ch))
    w = torch.einsum("bct,bcs->bts", q * scale, k * scale)
    w = torch.softmax(w.float(), dim=-1)
    h = torch.einsum("bts,bcs->bct", w, v).reshape(b, -1, x.shape[-1])
    h = F.conv1d(h, sd[p + ".proj_out.weight"], sd[p + ".proj_out.bias"])
    return (x + h).reshape(b, c, *spatial)


def unet_block_plan(model_channels, channel_mult, num_res_blocks, attention_resolutions):
    """Mirror of the constructor loop unet.py:343-430: which sub-modules every block holds.

    Returns (input_blocks, output_blocks): lists of lists of ('conv'|'res'|'attn'|'down'|'up')."""
    inp = [["conv"]]
    ds = 1
    for level, _ in enumerate(channel_mult):
        for _ in range(num_res_blocks):
            layers = ["res"]
            if ds in attention_resolutions:
                layers.append("attn")
            inp.append(layers)
        if level != len(channel_mult) - 1:
            inp.append(["down"])
            ds *= 2
    out = []
    for level in list(range(len(channel_mult)))[::-1]:
        for i in range(num_res_blocks + 1):
            layers = ["res"]
            if ds in attention_resolutions:
                layers.append("attn")
            if level and i == num_res_blocks:
                layers.append("up")
                ds //= 2
            out.append(layers)
    return inp, out


def unet_forward(sd, cfg, x, timesteps):
    """unet.py:463-492.  cfg: dict(model_channels, channel_mult, num_res_blocks,
    attention_resolutions, num_heads)."""
    mc = cfg["model_channels"]
    nh = cfg["num_heads"]
    inp, out = unet_block_plan(mc, cfg["channel_mult"], cfg["num_res_blocks"], cfg["attention_resolutions"])
    emb = timestep_embedding(timesteps, mc)
    emb = F.linear(emb, sd["time_embed.0.weight"], sd["time_embed.0.bias"])
    emb = F.linear(_silu(emb), sd["time_embed.2.weight"], sd["time_embed.2.bias"])

    def run(prefix, layers, h):
        for j, kind in enumerate(layers):
            p = "%s.%d" % (prefix, j)
            if kind == "conv":
                h = F.conv2d(h, sd[p + ".weight"], sd[p + ".bias"], padding=1)
            elif kind == "res":
                h = _resblock(sd, p, h, emb)
            elif kind == "attn":
                h = _attention(sd, p, h, nh)
            elif kind == "down":
                h = F.conv2d(h, sd[p + ".op.weight"], sd[p + ".op.bias"], stride=2, padding=1)
            elif kind == "up":
                h = F.interpolate(h, scale_factor=2, mode="nearest")
                h = F.conv2d(h, sd[p + ".conv.weight"], sd[p + ".conv.bias"], padding=1)
        return h

    hs = []
    h = x.float()
    for i, layers in enumerate(inp):
        h = run("input_blocks.%d" % i, layers, h)
        hs.append(h)
    h = run("middle_block", ["res", "attn", "res"], h)
    for i, layers in enumerate(out):
        h = run("output_blocks.%d" % i, layers, torch.cat([h, hs.pop()], dim=1))
    h = _silu(_gn(h, sd, "out.0", h.shape[1]))
    return F.conv2d(h, sd["out.2.weight"], sd["out.2.bias"], padding=1)


def _mlp_block(sd, p, x, t_emb):
    """DiffusionBlocks.py:125-136 (time=True, a=False, skip, LayerNorm)."""
    n = x.shape[-1]
    skip = x
    h = F.layer_norm(F.linear(x, sd[p + ".mlp_1.1.weight"], sd[p + ".mlp_1.1.bias"]), [n],
                     sd[p + ".mlp_1.2.weight"], sd[p + ".mlp_1.2.bias"])
    h = _silu(h)
    h = h + _silu(F.linear(t_emb, sd[p + ".t_proj.1.weight"], sd[p + ".t_proj.1.bias"]))
    h = F.layer_norm(F.linear(h, sd[p + ".mlp_2.1.weight"], sd[p + ".mlp_2.1.bias"]), [n],
                     sd[p + ".mlp_2.2.weight"], sd[p + ".mlp_2.2.bias"])
    return _silu(h + skip)


def mlp_forward(sd, nblocks, x, timestep):
    """Model.py:148-211.  x: (B, 1, nfeatures); timestep: (B,)."""
    t = timestep.unsqueeze(1).unsqueeze(2).to(torch.float32)
    t = _silu(F.linear(t, sd["time_mlp.0.weight"], sd["time_mlp.0.bias"]))
    t = _silu(F.linear(t, sd["time_mlp.2.weight"], sd["time_mlp.2.bias"]))
    n = sd["linear_in.weight"].shape[0]
    val = F.linear(x, sd["linear_in.weight"], sd["linear_in.bias"])
    val = _silu(F.layer_norm(val, [n], sd["group_norm_in.weight"], sd["group_norm_in.bias"]))
    for i in range(nblocks):
        val = _mlp_block(sd, "midblocks.%d" % i, val, t)
    val = _mlp_block(sd, "outblocks_mean.0", val, t)
    return F.linear(val, sd["outblocks_mean.1.weight"], sd["outblocks_mean.1.bias"])
