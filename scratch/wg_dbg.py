import sys, os, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "diamond-ppo_b200"))
from diamond import _native as N
ctx = N.get_context(0)
torch.set_printoptions(linewidth=200, precision=3, sci_mode=False)
M, N1, N2 = 1024, 128, 64
g = torch.Generator(device="cuda").manual_seed(1)
Dm = torch.randn(M, N1, device="cuda", generator=g); Hm = torch.randn(M, N2, device="cuda", generator=g)
ref = Dm.double().T @ Hm.double()
dW = ctx.tc_wgrad(Dm, Hm)
print("max |dW|", dW.abs().max().item(), "max|ref|", ref.abs().max().item())
print("ratio sample", (dW[:4, :8].double() / ref[:4, :8]))
print("dW[:4,:8]", dW[:4, :8]); print("ref[:4,:8]", ref[:4, :8])
# structured: single row nonzero
Dm2 = torch.zeros(M, N1, device="cuda"); Dm2[5, :] = torch.arange(N1, device="cuda").float() + 1
Hm2 = torch.zeros(M, N2, device="cuda"); Hm2[5, :] = torch.arange(N2, device="cuda").float() + 1
dW2 = ctx.tc_wgrad(Dm2, Hm2)
ref2 = Dm2.T @ Hm2
print("one-row err", (dW2 - ref2).abs().max().item())
print(dW2[:6, :10]); print(dW2[30:36, :10]); print(dW2[:3, 30:40])
