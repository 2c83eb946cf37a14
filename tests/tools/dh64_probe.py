import sys, os
sys.path.insert(0, ""+os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))+"")
import numpy as np, torch
from mla_b200 import m3ae
def relf(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30)
g = np.load(""+os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))+"/tests/golden/m3ae_dh64.npz")
for mode in ("fp16", "tf32", "permod"):
    m3ae.FUSED_BLOCK = mode != "permod"
    m3ae.BLOCK_BACKWARD = mode if mode != "permod" else "fp16"
    enc = m3ae.MaskedMultimodalAutoencoder(64, dict(model_type=None, emb_dim=128, depth=1, num_heads=2))
    enc.load_state_dict({k[6:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("state/")}, strict=True)
    enc = enc.cuda()
    gen = torch.Generator().manual_seed(17)
    text = torch.randint(0, 64, (3, 70), generator=gen)
    pm = (torch.arange(70)[None, :] >= torch.tensor([70, 33, 1])[:, None]).long()
    image = torch.randn(3, 9, 768, generator=gen)
    t = enc.forward_representation(None, text.cuda(), pm.cuda())
    v = enc.forward_representation(image.cuda(), None, None)
    wt, wv = torch.randn(t.shape, generator=gen), torch.randn(v.shape, generator=gen)
    ((t * wt.cuda()).sum() + (v * wv.cuda()).sum()).backward()
    params = dict(enc.named_parameters())
    out = [mode, "fwd %.1e %.1e" % (relf(t.detach().cpu(), g["rep_text"]), relf(v.detach().cpu(), g["rep_image"]))]
    for k in g.files:
        if k.startswith("grad/"):
            out.append("%s %.1e" % (k[5:].split(".")[-2] + "." + k[5:].split(".")[-1] if "." in k[5:] else k[5:], relf(params[k[5:]].grad.cpu(), g[k])))
    w = params["encoder.blocks.0.attention.qkv_linear.weight"].grad.cpu().numpy(); r = g["grad/encoder.blocks.0.attention.qkv_linear.weight"]
    out.append("q/k/v parts %.1e %.1e %.1e | norms %.1e %.1e %.1e" % (relf(w[:128], r[:128]), relf(w[128:256], r[128:256]), relf(w[256:], r[256:]), np.linalg.norm(r[:128]), np.linalg.norm(r[128:256]), np.linalg.norm(r[256:])))
    print(" | ".join(out), flush=True)
