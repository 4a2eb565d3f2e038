"""Diagnostic: dependence of AutoencoderKL results on how frames are cut into calls."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from weatherforecastingtoolkit_b200.models.autoencoderkl import AutoencoderKL
from weatherforecastingtoolkit_b200.synthetic import PATHB_AKL_CONFIG, make_akl_state_dict, make_vil_sequences

def rel(a, b):
    a, b = a.double(), b.double()
    return ((a - b).norm() / b.norm()).item()

dev = "cuda:0"
cfg = PATHB_AKL_CONFIG
m = AutoencoderKL(**cfg); m.load_state_dict(make_akl_state_dict(cfg, seed=0, affine_jitter=0.1)); m = m.to(dev)
hw = int(sys.argv[1]) if len(sys.argv) > 1 else 384
u8 = make_vil_sequences(1, hw, hw, 25, seed=5).to(dev)
x = (u8.float() / 255).permute(0, 3, 1, 2).reshape(25, 1, hw, hw).contiguous()
def enc(x, k):
    return torch.cat([m.encode(x[i:i + k]).mode() for i in range(0, x.shape[0], k)])
def dec(z, k):
    return torch.cat([m.decode(z[i:i + k]) for i in range(0, z.shape[0], k)])
z25 = enc(x, 25); z25b = enc(x, 25); z8 = enc(x, 8); z1 = enc(x, 1); z37 = enc(torch.cat([x, x[:12]]), 37)[:25]
print("encode: run-to-run", rel(z25b, z25), " 8 vs 25", rel(z8, z25), " 1 vs 25", rel(z1, z25), " 37 vs 25", rel(z37, z25))
y25 = dec(z25, 25); y25b = dec(z25, 25); y8 = dec(z25, 8); y1 = dec(z25, 1)
print("decode: run-to-run", rel(y25b, y25), " 8 vs 25", rel(y8, y25), " 1 vs 25", rel(y1, y25))
# per-frame diffs for k=8 vs 25
d = [(rel(y8[i:i+1], y25[i:i+1])) for i in range(25)]
print("decode per-frame 8 vs 25:", ["%.1e" % v for v in d])
d = [(rel(z8[i:i+1], z25[i:i+1])) for i in range(25)]
print("encode per-frame 8 vs 25:", ["%.1e" % v for v in d])

# ---- which variant is closer to the fp32 oracle? (frame 0 only; CPU oracle)
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import akl_oracle as O
sd = make_akl_state_dict(cfg, seed=0, affine_jitter=0.1)
torch.set_num_threads(os.cpu_count())
with torch.no_grad():
    mom = O.akl_encode_moments(x[:1].cpu(), sd, cfg)
    zo = mom[:, :cfg["latent_channels"]]
    yo = O.akl_decode(z25[:1].cpu(), sd, cfg)
print("encode frame0 vs oracle: n=25 %.3e  n=8 %.3e  n=1 %.3e" % (rel(z25[:1].cpu(), zo), rel(z8[:1].cpu(), zo), rel(z1[:1].cpu(), zo)))
print("decode frame0 vs oracle: n=25 %.3e  n=8 %.3e  n=1 %.3e" % (rel(y25[:1].cpu(), yo), rel(y8[:1].cpu(), yo), rel(y1[:1].cpu(), yo)))
