import sys, os, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "diamond-ppo_b200"))
from diamond import _native as N
ctx = N.get_context(0)
g = torch.Generator(device="cuda").manual_seed(0)
M = 65536
A = torch.randn(M, 256, device="cuda", generator=g); W = torch.randn(256, 256, device="cuda", generator=g) / 16
b = torch.randn(256, device="cuda", generator=g)
A5 = torch.randn(M, 512, device="cuda", generator=g); W5 = torch.randn(512, 256, device="cuda", generator=g) / 22
Hact = torch.tanh(torch.randn(M, 256, device="cuda", generator=g))
for _ in range(2):
    ctx.tc_linear(1, A, W, False, bias=b, variant=3)
    ctx.tc_linear(2, A5, W5, True, Hact=Hact, colsum=True, variant=3)
torch.cuda.synchronize()
print("ok")
