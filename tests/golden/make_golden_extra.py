"""Golden fixtures for the components outside the headline Path-B linear rollout, generated from the
UNMODIFIED reference (/root/reference; build container only). Run:

    python tests/golden/make_golden_extra.py            # writes extra_golden.npz

* DLinear predictors -- the class definitions live inside train scripts that import pytorch_lightning /
  wandb / omegaconf (absent here), so ``ref_script_classes`` extracts the ``ClassDef`` nodes with ``ast`` and
  executes exactly that source (no edits) in a namespace holding ``torch`` / ``nn`` / ``F``.
* NLayerDiscriminator (pipeline/models/autoencoderkl/losses/model.py), PosAwareAE_TF
  (pipeline/models/ae_64x8x8_lin.py), AE_ViT_2048 (pipeline/models/ae_vit.py) are imported directly.

Inputs and weights are regenerated from seeds at test time (``weatherforecastingtoolkit_b200.synthetic``);
only reference OUTPUTS are stored.
"""
import ast
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REFERENCE = "/root/reference"
sys.dont_write_bytecode = True
sys.path.insert(0, REFERENCE)
sys.path.insert(0, ROOT)

from weatherforecastingtoolkit_b200 import synthetic as S  # noqa: E402


def ref_script_classes(rel_path, names):
    """Execute the named top-level class definitions of a reference script, unmodified."""
    import torch.nn as nn
    import torch.nn.functional as F
    src = open(os.path.join(REFERENCE, rel_path)).read()
    tree = ast.parse(src)
    ns = {"torch": torch, "nn": nn, "F": F}
    for node in tree.body:
        if isinstance(node, ast.ClassDef) and node.name in names:
            exec(compile(ast.Module(body=[node], type_ignores=[]), rel_path, "exec"), ns)
    missing = [n for n in names if n not in ns]
    assert not missing, missing
    return ns


DLINEAR_SCRIPTS = {
    "shared": "experiments/v1_experiments/pretrained_ae_dlinear_sevir/train.py",
    "individual": "experiments/v1_experiments/pretrained_ae_dlinear_ind/train.py",
    "indc_indp": "experiments/v1_experiments/pretrained_ae_dlinear_indc_indp/train.py",
}


def ref_dlinear(variant, cfg, params):
    """Reference DLinear module of `variant` with the synthetic parameters loaded."""
    ns = ref_script_classes(DLINEAR_SCRIPTS[variant], ["moving_avg", "series_decomp", "DLinear"])
    m = ns["DLinear"](cfg)
    ws, bs, wt, bt = params
    with torch.no_grad():
        if cfg.individual:
            for i in range(cfg.enc_in):
                m.Linear_Seasonal[i].weight.copy_(ws[i])
                m.Linear_Seasonal[i].bias.copy_(bs[i])
                m.Linear_Trend[i].weight.copy_(wt[i])
                m.Linear_Trend[i].bias.copy_(bt[i])
        else:
            m.Linear_Seasonal.weight.copy_(ws)
            m.Linear_Seasonal.bias.copy_(bs)
            m.Linear_Trend.weight.copy_(wt)
            m.Linear_Trend.bias.copy_(bt)
    return m.eval()


def ref_dlinear_step(m, lat, variant, in_frames=13):
    """validation_step algebra around the predictor (pretrained_ae_dlinear_sevir/train.py:179-192;
    indc_indp: train.py:179-192 with the (t c) reshape)."""
    b, t, c, h, w = lat.shape
    inp, tgt = lat[:, :in_frames], lat[:, in_frames:]
    inp_t = inp[:, -1].unsqueeze(1)
    inp = inp - inp_t
    tgt = tgt - inp_t
    if variant == "indc_indp":
        pred = m(inp.reshape(b, in_frames * c, h * w))
    else:
        pred = m(inp.reshape(b, in_frames, c * h * w))
    pred = pred.reshape(b, t - in_frames, c, h, w)
    loss = torch.nn.functional.mse_loss(pred, tgt)
    return pred + inp_t, tgt + inp_t, loss


def gen_dlinear(out):
    for variant in ("shared", "individual", "indc_indp"):
        cfg, params, lat = S.make_dlinear_case(variant)
        m = ref_dlinear(variant, cfg, params)
        with torch.no_grad():
            pred, tgt, loss = ref_dlinear_step(m, lat, variant)
        out[f"dlinear_{variant}_pred"] = pred.numpy()
        out[f"dlinear_{variant}_tgt"] = tgt.numpy()
        out[f"dlinear_{variant}_loss"] = np.array(loss.item())


def disc_inputs(n, hw, seed):
    u8 = S.make_vil_sequences(n, hw, hw, 2, seed=seed)
    x = ((1 / 255) * u8.float()).permute(0, 3, 1, 2).contiguous()
    return x[:, :1].contiguous(), x[:, 1:2].contiguous()


def gen_discriminator(out):
    from pipeline.models.autoencoderkl.losses.model import NLayerDiscriminator
    from pipeline.models.autoencoderkl.losses import model as RM  # noqa: F401  (contperceptual pulls lpips: not imported)
    sd = S.make_discriminator_state_dict()
    m = NLayerDiscriminator(input_nc=1, ndf=64, n_layers=3).eval()
    m.load_state_dict(sd, strict=True)
    with torch.no_grad():
        for tag, n, hw, seed in (("disc64", 2, 64, 41), ("disc384", 1, 384, 42)):
            real, fake = disc_inputs(n, hw, seed)
            lr, lf = m(real), m(fake)
            out[f"{tag}_logits_real"] = lr.numpy()
            out[f"{tag}_logits_fake"] = lf.numpy()
            # hinge_d_loss (losses/contperceptual.py:19-23; that module cannot be imported here: it pulls lpips.py,
            # which writes into the read-only tree) -- three reference lines restated:
            d_loss = 0.5 * (torch.mean(torch.nn.functional.relu(1. - lr)) + torch.mean(torch.nn.functional.relu(1. + lf)))
            out[f"{tag}_hinge"] = np.array(d_loss.item())


POSAWARE_GAIN = 1.1   # keeps activations O(0.2-0.5) through the ~110 layers and the sigmoid unsaturated


def posaware_inputs(n=2, seed=51):
    u8 = S.make_vil_sequences(n, 128, 128, 1, seed=seed)
    return ((1 / 255) * u8.float()).permute(0, 3, 1, 2).contiguous()


def posaware_state_dict():
    from weatherforecastingtoolkit_b200.models.ae_64x8x8_lin import PosAwareAE_TF as Mine
    return S.fill_state_dict(Mine(), "posaware", 0, gain=POSAWARE_GAIN)


def gen_posaware(out):
    from pipeline.models.ae_64x8x8_lin import PosAwareAE_TF
    sd = posaware_state_dict()
    m = PosAwareAE_TF().eval()
    m.load_state_dict(sd, strict=True)
    x = posaware_inputs()
    with torch.no_grad():
        y, z = m(x)
    out["posaware_latent"] = z.numpy()
    out["posaware_recon"] = y.numpy()


def vit_inputs(n=2, seed=61):
    u8 = S.make_vil_sequences(n, 128, 128, 1, seed=seed)
    return ((1 / 255) * u8.float()).permute(0, 3, 1, 2).contiguous()


def vit_state_dict():
    from weatherforecastingtoolkit_b200.models.ae_vit import AE_ViT_2048 as Mine
    return S.fill_state_dict(Mine(), "vit", 0, gain=1.0)


def gen_vit(out):
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):      # the reference module builds a model and prints at import
        from pipeline.models.ae_vit import AE_ViT_2048
    sd = vit_state_dict()
    m = AE_ViT_2048().eval()
    m.load_state_dict(sd, strict=True)
    x = vit_inputs()
    with torch.no_grad():
        y, lat = m(x)
        # token-sequence latent [B, 64, 512] (BASELINE config 4): the encoder stack's output, ae_vit.py:141-145
        z = m.patch_embed(x).flatten(2).transpose(1, 2) + m.pos_embed
        z = m.encoder(z)
    out["vit_recon"] = y.numpy()
    out["vit_latent"] = lat.numpy()
    out["vit_tokens"] = z.numpy()


CONVMODEL_SCRIPT = "experiments/v1_experiments/pretrained_ae_convae_sevir/train.py"


def convmodel_case(seed=0, b=2, t=3):
    """(state_dict, latents [b, t, 4, 48, 48]): kaiming-normal weights like ConvModel.init_weights (train.py:127-131),
    random (non-trivial) LayerNorm affine parameters and biases."""
    from weatherforecastingtoolkit_b200.predictors import ConvModel as Mine
    torch.manual_seed(1234 + seed)
    m = Mine()
    sd = {}
    for k, v in m.state_dict().items():
        g = torch.Generator().manual_seed(S._name_seed("convmodel." + k, seed))
        leaf = k.rsplit(".", 1)[-1]
        is_ln = v.ndim == 3
        if leaf == "weight" and not is_ln:
            fan_in = int(np.prod(v.shape[1:]))
            sd[k] = torch.randn(v.shape, generator=g) * float(np.sqrt(2.0 / (1 + 0.01 ** 2) / fan_in))
        elif leaf == "weight":
            sd[k] = 1.0 + 0.1 * torch.randn(v.shape, generator=g)
        else:
            sd[k] = 0.05 * torch.randn(v.shape, generator=g)
    g = torch.Generator().manual_seed(S._name_seed("convmodel.x", seed))
    x = torch.randn((b, t, 4, 48, 48), generator=g)
    return sd, x


def gen_convmodel(out):
    ns = ref_script_classes(CONVMODEL_SCRIPT, ["ConvEncoder", "ConvDecoder", "ConvModel"])
    sd, x = convmodel_case()
    m = ns["ConvModel"](latent_dim=512).eval()
    m.load_state_dict(sd, strict=True)
    with torch.no_grad():
        z, rec = m(x)
        loss = torch.nn.HuberLoss()(rec, x)
    out["convmodel_z"] = z.numpy()
    out["convmodel_recon"] = rec.numpy()
    out["convmodel_huber"] = np.array(loss.item())


CONVATTN_SCRIPT = "experiments/v1_experiments/pretrained_ae_convattn_ae_sevir/train.py"


def convattn_case(seed=0, b=3, layers=4, latent_dim=512):
    """(state_dict, latents [b, 4, 48, 48]): kaiming-normal matrices like ConvAttnModel.init_weights (train.py:121-129),
    N(0,1) embeddings / queries like the constructor, random (non-trivial) norm affine parameters and biases."""
    from weatherforecastingtoolkit_b200.predictors import ConvAttnModel as Mine
    torch.manual_seed(4321 + seed)
    m = Mine(num_tf_layers=layers, latent_dim=latent_dim)
    sd = {}
    for k, v in m.state_dict().items():
        g = torch.Generator().manual_seed(S._name_seed("convattn." + k, seed))
        leaf = k.rsplit(".", 1)[-1]
        if v.ndim >= 2 and leaf in ("weight", "in_proj_weight"):
            fan_in = int(np.prod(v.shape[1:]))
            sd[k] = torch.randn(v.shape, generator=g) * float(np.sqrt(2.0 / fan_in))
        elif v.ndim == 3:                                   # pos embeddings, queries
            sd[k] = torch.randn(v.shape, generator=g)
        elif leaf == "weight":                              # LayerNorm / GroupNorm scale
            sd[k] = 1.0 + 0.1 * torch.randn(v.shape, generator=g)
        else:
            sd[k] = 0.05 * torch.randn(v.shape, generator=g)
    g = torch.Generator().manual_seed(S._name_seed("convattn.x", seed))
    x = torch.randn((b, 4, 48, 48), generator=g)
    return sd, x


def gen_convattn(out):
    ns = ref_script_classes(CONVATTN_SCRIPT, ["ConvAttnModel"])
    sd, x = convattn_case()
    m = ns["ConvAttnModel"]().eval()
    m.load_state_dict(sd, strict=True)
    with torch.no_grad():
        z, rec = m(x)
        loss = torch.nn.HuberLoss()(rec, x)
    out["convattn_z"] = z.numpy()
    out["convattn_recon"] = rec.numpy()
    out["convattn_huber"] = np.array(loss.item())


def main():
    torch.set_num_threads(os.cpu_count())
    out = {}
    gen_dlinear(out)
    for fn in EXTRA_GENERATORS:
        fn(out)
    np.savez_compressed(os.path.join(HERE, "extra_golden.npz"), **{k: np.asarray(v, dtype=np.float32) for k, v in out.items()})
    print("wrote extra_golden.npz:", {k: v.shape for k, v in out.items()})


EXTRA_GENERATORS = [gen_discriminator, gen_posaware, gen_vit, gen_convmodel, gen_convattn]

if __name__ == "__main__":
    main()
