import sys, os, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "diamond-ppo_b200"))
from diamond import _native as N
ctx = N.get_context(0)
torch.backends.cuda.matmul.allow_tf32 = False
def t(f, n=20):
    f(); torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n): f()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / n * 1e3
for (M, Nn, K) in [(4096, 256, 64), (4096, 256, 256), (4096, 512, 256), (65536, 256, 64), (65536, 256, 256), (65536, 512, 256)]:
    g = torch.Generator(device="cuda").manual_seed(M + Nn + K)
    A = torch.randn(M, K, device="cuda", generator=g); W = torch.randn(Nn, K, device="cuda", generator=g) / K ** 0.5
    b = torch.randn(Nn, device="cuda", generator=g)
    ref = torch.tanh(A.double() @ W.double().T + b.double())
    e32 = (torch.tanh(A @ W.T + b).double() - ref).abs().max().item()
    out = [f"fwd M={M} N={Nn} K={K}: torch-fp32 {e32:.2e}"]
    for v in (3, 2):
        C, _ = ctx.tc_linear(1, A, W, False, bias=b, variant=v)
        out.append(f"v{v} err {(C.double() - ref).abs().max().item():.2e} {t(lambda: ctx.tc_linear(1, A, W, False, bias=b, variant=v)):.1f}us")
    print(" | ".join(out), flush=True)
for (M, Nn, K) in [(4096, 256, 512), (65536, 256, 512), (65536, 256, 256)]:
    g = torch.Generator(device="cuda").manual_seed(M + Nn + K + 1)
    A = torch.randn(M, K, device="cuda", generator=g); W = torch.randn(K, Nn, device="cuda", generator=g) / K ** 0.5
    Hact = torch.tanh(torch.randn(M, Nn, device="cuda", generator=g))
    ref = (A.double() @ W.double()) * (1.0 - Hact.double() ** 2)
    e32 = (((A @ W) * (1 - Hact ** 2)).double() - ref).abs().max().item()
    out = [f"dgrad M={M} N={Nn} K={K}: torch-fp32 {e32:.2e}"]
    for v in (3, 2):
        C, cs = ctx.tc_linear(2, A, W, True, Hact=Hact, colsum=True, variant=v)
        ce = (cs.double().sum(0) - ref.sum(0)).abs().max().item()
        out.append(f"v{v} err {(C.double() - ref).abs().max().item():.2e} colsum {ce:.2e} {t(lambda: ctx.tc_linear(2, A, W, True, Hact=Hact, colsum=True, variant=v)):.1f}us")
    print(" | ".join(out), flush=True)
for (M, N1, N2) in [(4096, 256, 256), (4096, 512, 256), (4096, 256, 64), (4100, 128, 128), (65536, 512, 256), (65536, 256, 256), (65536, 256, 64)]:
    g = torch.Generator(device="cuda").manual_seed(M + N1 + N2)
    Dm = torch.randn(M, N1, device="cuda", generator=g); Hm = torch.randn(M, N2, device="cuda", generator=g)
    ref = Dm.double().T @ Hm.double()
    e32 = ((Dm.T @ Hm).double() - ref).abs().max().item() / ref.abs().max().item()
    dW = ctx.tc_wgrad(Dm, Hm)
    err = ((dW.double() - ref).abs().max() / ref.abs().max()).item()
    print(f"wgrad M={M} N1={N1} N2={N2}: torch-fp32 {e32:.2e} | tc err {err:.2e} {t(lambda: ctx.tc_wgrad(Dm, Hm)):.1f}us (incl. reduce)", flush=True)
