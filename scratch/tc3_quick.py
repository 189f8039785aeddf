import sys, os, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "diamond-ppo_b200"))
from diamond import _native as N
ctx = N.get_context(0)
g = torch.Generator(device="cuda").manual_seed(0)
for (M, Nn, K) in [(1024, 256, 64), (4096, 256, 256), (5000, 512, 256), (65536, 256, 256)]:
    A = torch.randn(M, K, device="cuda", generator=g); W = torch.randn(Nn, K, device="cuda", generator=g) / K ** 0.5
    b = torch.randn(Nn, device="cuda", generator=g)
    C, _ = ctx.tc_linear(1, A, W, False, bias=b, variant=3)
    torch.cuda.synchronize()
    ref = torch.tanh(A.double() @ W.double().T + b.double())
    print(M, Nn, K, "err", (C.double() - ref).abs().max().item(), flush=True)
