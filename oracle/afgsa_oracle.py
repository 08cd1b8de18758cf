"""CPU oracle: functional restatement of the reference AFGSA generator hot path.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Every function restates the
algorithm of the cited reference file:line (paths relative to the reference
checkout) with plain torch CPU ops; autograd through these functions is the
backward oracle.  Pinned against the real reference modules by
``tests/golden/make_golden.py`` (run in the build container where the reference
is mounted); the resulting fixtures live in ``tests/golden/``.

All tensors are NCHW like the reference.  ``sd`` is a state dict with the
reference's parameter names (pht/models/afgsa/model.py:606-715).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

LEAKY_SLOPE = 0.2  # pht/models/afgsa/model.py:67,78


def _act(x: torch.Tensor, kind: str | None) -> torch.Tensor:
    """pht/models/afgsa/model.py:64-83 (relu / leakyrelu(0.2) / none)."""
    if kind is None:
        return x
    if kind == "relu":
        return torch.relu(x)
    if kind == "leakyrelu":
        return torch.where(x > 0, x, x * LEAKY_SLOPE)
    raise ValueError(kind)


def conv_block(x, w, b, act, padding_mode="zeros"):
    """conv_block, pht/models/afgsa/model.py:99-125: 'same' conv + activation.

    torch's Conv2d with padding_mode != zeros pads explicitly then runs a valid
    convolution; we restate exactly that.
    """
    k = w.shape[-1]
    p = (k - 1) // 2
    if p > 0:
        if padding_mode == "zeros":
            x = F.pad(x, (p, p, p, p))
        else:
            x = F.pad(x, (p, p, p, p), mode=padding_mode)
    return _act(F.conv2d(x, w, b), act)


def encoder_noisy(x, sd, padding_mode):
    """pht/models/afgsa/model.py:606-623, 719-722."""
    n1 = conv_block(x, sd["conv1.0.weight"], sd["conv1.0.bias"], "relu")
    n3 = conv_block(x, sd["conv3.0.weight"], sd["conv3.0.bias"], "relu", padding_mode)
    n5 = conv_block(x, sd["conv5.0.weight"], sd["conv5.0.bias"], "relu", padding_mode)
    return conv_block(torch.cat([n1, n3, n5], 1), sd["conv_map.0.weight"], sd["conv_map.0.bias"], "relu")


def encoder_aux(aux, sd, padding_mode):
    """pht/models/afgsa/model.py:625-658, 724-728."""
    a1 = conv_block(aux, sd["conv_a1.0.weight"], sd["conv_a1.0.bias"], "relu")
    a3 = conv_block(aux, sd["conv_a3.0.weight"], sd["conv_a3.0.bias"], "leakyrelu", padding_mode)
    a5 = conv_block(aux, sd["conv_a5.0.weight"], sd["conv_a5.0.bias"], "leakyrelu", padding_mode)
    a = conv_block(torch.cat([a1, a3, a5], 1), sd["conv_aenc1.0.weight"], sd["conv_aenc1.0.bias"], "leakyrelu")
    return conv_block(a, sd["conv_aenc2.0.weight"], sd["conv_aenc2.0.bias"], "leakyrelu")


def attention_core(q, k, v, rel_h, rel_w, block=8, halo=3, heads=4):
    """Block-local attention, pht/models/afgsa/model.py:474-516.

    q is already scaled by head_ch**-0.5.  Keys/values outside the image are
    ZERO (F.unfold padding, :480,:484); the relative position embedding is added
    after that padding (:496-497), so padded keys are NOT masked.  The curve
    permutation (:477,:506) cancels and is omitted.
    Returns O with the reference's channel order c = head*d + j.
    """
    B, C, H, W = q.shape
    d = C // heads
    win = block + 2 * halo
    nby, nbx = H // block, W // block
    kp = F.pad(k, (halo, halo, halo, halo))
    vp = F.pad(v, (halo, halo, halo, halo))
    # [B, C, nby, nbx, win(r), win(c)]
    kw = kp.unfold(2, win, block).unfold(3, win, block)
    vw = vp.unfold(2, win, block).unfold(3, win, block)
    kw = kw.reshape(B, heads, d, nby, nbx, win, win)
    vw = vw.reshape(B, heads, d, nby, nbx, win, win)
    rh = rel_h.reshape(win, 1, d // 2).expand(win, win, d // 2)
    rw = rel_w.reshape(1, win, d // 2).expand(win, win, d // 2)
    rel = torch.cat([rh, rw], dim=-1).permute(2, 0, 1)  # [d, r, c]
    kw = kw + rel[None, None, :, None, None]
    qb = q.reshape(B, heads, d, nby, block, nbx, block)
    sim = torch.einsum("bhdyixj,bhdyxrc->bhyxijrc", qb, kw)
    shp = sim.shape
    attn = torch.softmax(sim.reshape(*shp[:6], win * win), dim=-1).reshape(shp)
    out = torch.einsum("bhyxijrc,bhdyxrc->bhdyixj", attn, vw)
    return out.reshape(B, C, H, W)


def film(x, cond, sd, prefix):
    """FiLM.forward with use_spatial=True (the only way AFGSA builds it, model.py:443-449), film.py:36-45:
    gamma, beta = chunk(conv1x1(relu(conv1x1(cond))), 2, dim=1);  x' = gamma * x + beta."""
    h = conv_block(cond, sd[prefix + "affine.0.weight"], sd[prefix + "affine.0.bias"], "relu")
    gb = conv_block(h, sd[prefix + "affine.2.weight"], sd[prefix + "affine.2.bias"], None)
    gamma, beta = torch.chunk(gb, 2, dim=1)
    return gamma * x + beta


def afgsa(noisy, aux, sd, prefix, block=8, halo=3, heads=4):
    """AFGSA.forward, pht/models/afgsa/model.py:456-516 (conv_map variant, or the FiLM variant :458-460 when the
    state dict holds ``film.affine.*`` -- ``alpha`` is registered by the reference but unused in its forward)."""
    B, C, H, W = noisy.shape
    assert H % block == 0 and W % block == 0  # model.py:469-471
    d = C // heads
    if prefix + "film.affine.0.weight" in sd:
        n_aux = film(noisy, aux, sd, prefix + "film.")
    else:
        n_aux = conv_block(torch.cat([noisy, aux], 1), sd[prefix + "conv_map.0.weight"],
                           sd[prefix + "conv_map.0.bias"], "relu")
    q = F.conv2d(n_aux, sd[prefix + "q_conv.weight"]) * d ** -0.5
    k = F.conv2d(n_aux, sd[prefix + "k_conv.weight"])
    v = F.conv2d(noisy, sd[prefix + "v_conv.weight"])
    return attention_core(q, k, v, sd[prefix + "rel_h"], sd[prefix + "rel_w"], block, halo, heads)


def transformer_block(x0, a, sd, prefix, padding_mode, block=8, halo=3, heads=4):
    """TransformerBlock.forward, pht/models/afgsa/model.py:571-582."""
    x1 = x0 + afgsa(x0, a, sd, prefix + "attention.", block, halo, heads)
    h = conv_block(x1, sd[prefix + "feed_forward.0.0.weight"], sd[prefix + "feed_forward.0.0.bias"], "relu", padding_mode)
    h = conv_block(h, sd[prefix + "feed_forward.1.0.weight"], sd[prefix + "feed_forward.1.0.bias"], "relu", padding_mode)
    return x1 + h


def afgsa_net_forward(x, aux, sd, padding_mode="replicate", num_sa=5, block=8, halo=3, heads=4):
    """AFGSANet.forward, pht/models/afgsa/model.py:717-733."""
    out = encoder_noisy(x, sd, padding_mode)
    a = encoder_aux(aux, sd, padding_mode)
    for i in range(num_sa):
        out = transformer_block(out, a, sd, f"transformer_blocks.{i}.", padding_mode, block, halo, heads)
    out = conv_block(out, sd["decoder.0.0.weight"], sd["decoder.0.0.bias"], "relu", padding_mode)
    out = conv_block(out, sd["decoder.1.0.weight"], sd["decoder.1.0.bias"], "relu", padding_mode)
    out = conv_block(out, sd["decoder.2.0.weight"], sd["decoder.2.0.bias"], None, "zeros")
    return out + x  # model.py:732


def l1_loss(output, target):
    """L1ReconstructionLoss, pht/models/losses.py:175-184 (mean |a-b|)."""
    return (output - target).abs().mean()


def preprocess_batch(noisy_nhwc, gt_nhwc, aux_nhwc):
    """Per-batch preprocessing, pht/models/base_trainer.py:373-383 with
    preprocess_normal / preprocess_specular (preprocessing.py:19-22, 34-38).
    NHWC fp32 in -> NCHW fp32 out (noisy, gt, aux)."""
    aux = aux_nhwc.clone()
    n = torch.nan_to_num(aux[..., :3])
    aux[..., :3] = torch.clamp((n + 1.0) * 0.5, 0.0, 1.0)
    noisy = torch.log(noisy_nhwc + 1)
    gt = torch.log(gt_nhwc + 1)
    perm = (0, 3, 1, 2)
    return noisy.permute(perm).contiguous(), gt.permute(perm).contiguous(), aux.permute(perm).contiguous()


def adam_step(p, g, m, v, step, lr, beta1=0.9, beta2=0.999, eps=1e-8):
    """torch.optim.Adam single-tensor update as used by base_trainer.py:182-187
    (no weight decay, no amsgrad).  In-place on p, m, v; ``step`` is 1-based."""
    m.mul_(beta1).add_(g, alpha=1 - beta1)
    v.mul_(beta2).addcmul_(g, g, value=1 - beta2)
    bc1 = 1 - beta1 ** step
    bc2 = 1 - beta2 ** step
    denom = (v.sqrt() / (bc2 ** 0.5)).add_(eps)
    p.addcdiv_(m, denom, value=-(lr / bc1))


def g_only_train_step(x, aux, gt, sd, padding_mode="replicate", num_sa=5):
    """One generator-only step (G fwd, L1, backward): the part of
    base_trainer.py:388-457 that runs on hand-written kernels.  Returns
    (output, loss, grads dict)."""
    params = {k: v.detach().clone().requires_grad_(True) for k, v in sd.items() if v.dtype.is_floating_point}
    out = afgsa_net_forward(x, aux, params, padding_mode, num_sa=num_sa)
    loss = l1_loss(out, gt)
    loss.backward()
    # (parameters the forward never touches, i.e. FiLM's ``alpha``, have no gradient: reported as zeros)
    return out.detach(), loss.detach(), {k: (p.grad if p.grad is not None else torch.zeros_like(p)) for k, p in params.items()}


def reference_init_state_dict(seed: int = 990819, input_channels: int = 3, aux_channels: int = 7, base_ch: int = 256,
                              num_sa: int = 5, block: int = 8, halo: int = 3, heads: int = 4) -> dict:
    """The reference's random initialisation of AFGSANet (non-FiLM) under ``torch.manual_seed(seed)``, built from plain
    ``torch.nn`` modules created in the reference's registration order (pht/models/afgsa/model.py:606-715; per AFGSA
    layer: rel_h, rel_w ``randn`` :430-437, conv_map :449, q/k/v :450-452, then ``reset_parameters`` :518-524 =
    kaiming_normal(fan_out, relu) on q/k/v and normal(0, 1) on rel_h / rel_w).  Same generator draws in the same order
    => the same numbers as the reference module (pinned by tests/test_oracle_cpu.py against golden_meta.json's
    ``param_checksums``).  Used by bench.py's CPU arm so that it never imports the product package."""
    from torch import nn
    from torch.nn import init
    torch.manual_seed(seed)
    sd = {}

    def conv(name, cin, cout, k, bias=True):
        m = nn.Conv2d(cin, cout, kernel_size=k, padding=(k - 1) // 2, bias=bias)
        sd[name + ".weight"] = m.weight.detach().clone()
        if bias:
            sd[name + ".bias"] = m.bias.detach().clone()
        return m

    for nm, cin, k in (("conv1", input_channels, 1), ("conv3", input_channels, 3), ("conv5", input_channels, 5)):
        conv(nm + ".0", cin, 256, k)
    conv("conv_map.0", 768, base_ch, 1)
    for nm, k in (("conv_a1", 1), ("conv_a3", 3), ("conv_a5", 5)):
        conv(nm + ".0", aux_channels, 256, k)
    conv("conv_aenc1.0", 768, base_ch, 1)
    conv("conv_aenc2.0", base_ch, base_ch, 1)
    win, hd = block + 2 * halo, base_ch // heads
    for i in range(num_sa):
        pre = f"transformer_blocks.{i}."
        rel_h = torch.randn(1, win, 1, hd // 2)
        rel_w = torch.randn(1, 1, win, hd // 2)
        conv(pre + "attention.conv_map.0", 2 * base_ch, base_ch, 1)
        qkv = [nn.Conv2d(base_ch, base_ch, kernel_size=1, bias=False) for _ in range(3)]
        for m in qkv:
            init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
        init.normal_(rel_h, 0, 1)
        init.normal_(rel_w, 0, 1)
        sd[pre + "attention.rel_h"], sd[pre + "attention.rel_w"] = rel_h, rel_w
        for nm, m in zip("qkv", qkv):
            sd[pre + f"attention.{nm}_conv.weight"] = m.weight.detach().clone()
        conv(pre + "feed_forward.0.0", base_ch, base_ch, 3)
        conv(pre + "feed_forward.1.0", base_ch, base_ch, 3)
    conv("decoder.0.0", base_ch, base_ch, 3)
    conv("decoder.1.0", base_ch, base_ch, 3)
    conv("decoder.2.0", base_ch, 3, 3)
    return sd
