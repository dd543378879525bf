import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from nf_distillation_b200.models.maf import create_maf_model
from oracle import maf_oracle as MO

def rel(a, b):
    b = b.to(a.device)
    return ((a - b).abs().max() / (b.abs().max() + 1e-12)).item()

for D, H, K, B in ((6, 512, 5, 300), (63, 512, 3, 257)):
    torch.manual_seed(D)
    m = create_maf_model(dict(image_shape=[D], hidden_channels=H, K=K))
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    x = torch.randn(B, D) * 1.5 + 0.5
    o_outs, o_nll = MO.maf_forward(sd, D, K, x)
    m = m.cuda()
    with torch.no_grad():
        outs, nll, _ = m(x.cuda(), None)
    print(f"D={D}: nll rel {rel(nll, o_nll):.3e}", [f"{rel(a, b):.2e}" for a, b in zip(outs, o_outs)])
    with torch.no_grad():
        xb = m(z=outs[-1], reverse=True)[-1]
    print(f"   inverse roundtrip rel {rel(xb, x):.3e}  oracle-inverse rel {rel(xb, MO.maf_inverse(sd, D, K, o_outs[-1])):.3e}")
    # gradients
    xs = x.clone().requires_grad_(True)
    sdg = {k: (v.clone().requires_grad_(True) if v.dtype.is_floating_point and k != 'prior_h' else v) for k, v in sd.items()}
    _, n2 = MO.maf_forward(sdg, D, K, xs)
    n2.mean().backward()
    xg = x.cuda().requires_grad_(True)
    outs, nll, _ = m(xg, None)
    nll.mean().backward()
    print(f"   dx rel {rel(xg.grad, xs.grad):.3e}")
    worst = sorted(((rel(p.grad, sdg[n].grad), n) for n, p in m.named_parameters()), reverse=True)[:5]
    print("   worst param grads", [(f"{e:.2e}", n) for e, n in worst])
    cos = sorted((torch.nn.functional.cosine_similarity(p.grad.flatten().cpu(), sdg[n].grad.flatten(), dim=0).item(), n) for n, p in m.named_parameters())[:4]
    print("   lowest cosine similarity", [(f"{c:.5f}", n) for c, n in cos])
