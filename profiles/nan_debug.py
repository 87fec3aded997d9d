import os, sys, numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import neural_svd_b200 as N
from conftest import build_problem
from oracle import nsvd_oracle as O
engine = sys.argv[1]; B = int(sys.argv[2]); steps = int(sys.argv[3])
cfg = O.PathConfig.hydrogen(sequential=True)
N.set_engine(engine)
method, operator, importance, gt = build_problem(cfg, 0, "cuda")
opt = N.FusedRMSpropEMA(method.parameters(), lr=1e-4, alpha=0.999, eps=1e-10, ema_decay=0.995, num_iters=1500)
for it in range(steps):
    x = N.sample_gaussian(B, cfg.sampling_scale, seed=7, offset=it * B)
    opt.zero_grad()
    loss, aux = method.compute_loss_operator(operator, x, importance=importance)
    loss.backward()
    gn = {n: float(p.grad.norm()) for n, p in method.named_parameters() if p.grad is not None}
    bad = not np.isfinite(float(loss.detach())) or not all(np.isfinite(v) for v in gn.values())
    if it % 20 == 0 or bad:
        r = x.norm(dim=1)
        print(f"[{engine}] it {it} loss {float(loss.detach()):.2f} |f|max {float(aux['f'].abs().max()):.3g} |Tf|max {float(aux['Tf'].abs().max()):.3g} "
              f"rmin {float(r.min()):.3g} rmax {float(r.max()):.3g} gW0 {gn['model.base.ws.0']:.3g} gW3 {gn['model.base.ws.3']:.3g} "
              f"|W0| {float(method.model.base.ws[0].norm()):.4g} |W3| {float(method.model.base.ws[3].norm()):.4g}", flush=True)
    if bad:
        f, Tf = aux["f"], aux["Tf"]
        print("nonfinite f:", int((~torch.isfinite(f)).sum()), "Tf:", int((~torch.isfinite(Tf)).sum()))
        idx = (~torch.isfinite(Tf)).nonzero()[:5]
        for i in idx:
            print("  row", int(i[0]), "col", int(i[1]), "x", x[i[0]].tolist(), "r", float(x[i[0]].norm()), "f", float(f[i[0], i[1]]))
        for n, p in method.named_parameters():
            print("  ", n, "param finite", bool(torch.isfinite(p).all()), "grad finite", p.grad is None or bool(torch.isfinite(p.grad).all()))
        break
    opt.step()
