import sys, os
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "diamond-ppo_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import ppo_oracle as O
from diamond import _native as N
from diamond.flat import FlatMlp
from test_kernels_gpu import rand_params, dev, nerr
ctx = N.get_context(0)

# 1. stress the flaky SIMT forward
D, H, A, B = 4, 64, 2, 1024
rng = np.random.default_rng(D * 7 + H)
p = rand_params(rng, O.DISCRETE_PARAM_NAMES, D, H, A, False)
fm = FlatMlp(D, H, A, False)
flat = fm.pack(p, device="cuda")
obs = rng.standard_normal((B, D)).astype(np.float32)
ref_out, ref_v = O.mlp_forward(p, torch.as_tensor(obs), False)
bad = 0
for it in range(300):
    ws = torch.empty(ctx.mlp_workspace_bytes(fm.desc, 300, False) // 4, device="cuda")
    out = torch.empty(B, A, device="cuda"); val = torch.empty(B, device="cuda")
    ctx.mlp_forward(fm.desc, flat, dev(obs), B, 3, out, val, ws)
    e = nerr(out.cpu().numpy(), ref_out.numpy())
    if e > 1e-5:
        bad += 1
        d = np.abs(out.cpu().numpy() - ref_out.numpy())
        print("iter", it, "err", e, "rows", np.where(d.max(1) > 1e-6)[0][:10], flush=True)
print("stress: bad", bad, "of 300", flush=True)

# 2. the failing tensor-core config
D, H, A, rows = 16, 512, 5, 1024
rng = np.random.default_rng(H + rows)
p = rand_params(rng, O.DISCRETE_PARAM_NAMES, D, H, A, False)
fm = FlatMlp(D, H, A, False)
flat = fm.pack(p, device="cuda")
obs = rng.standard_normal((rows, D)).astype(np.float32)
ref_out, ref_v = O.mlp_forward(p, torch.as_tensor(obs), False)
ws = torch.empty(ctx.mlp_workspace_bytes(fm.desc, rows, False) // 4 + 512, device="cuda")
for tc in (1, 0):
    ctx.set_option("tensor_cores", tc)
    out = torch.empty(rows, A, device="cuda"); val = torch.empty(rows, device="cuda")
    dobs = dev(obs)
    ctx.mlp_forward(fm.desc, flat, dobs, rows, 3, out, val, ws)
    torch.cuda.synchronize()
    d = np.abs(out.cpu().numpy() - ref_out.numpy()); dv = np.abs(val.cpu().numpy() - ref_v.numpy())
    print("tc", tc, "out nerr", nerr(out.cpu().numpy(), ref_out.numpy()), "val nerr", nerr(val.cpu().numpy(), ref_v.numpy()),
          "max|ref|", np.abs(ref_out.numpy()).max(), np.abs(ref_v.numpy()).max(), "bad rows", np.where(d.max(1) > 1e-4)[0][:10], flush=True)
