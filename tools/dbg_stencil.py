import subprocess, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = r'''
import sys, numpy as np, torch
sys.path.insert(0, %r)
import pyxu_b200.operator as pxo
dt, n1, n2, k1, k2, c1, c2, dense, batch = sys.argv[1:]
dt = np.float32 if dt == "f32" else np.float64
n1, n2, k1, k2, c1, c2, dense, batch = map(int, (n1, n2, k1, k2, c1, c2, dense, batch))
rng = np.random.default_rng(0)
kern = rng.standard_normal((k1, k2)).astype(dt) if dense else [rng.standard_normal(k1).astype(dt), rng.standard_normal(k2).astype(dt)]
a = pxo.Stencil(arg_shape=(n1, n2), kernel=kern, center=(c1, c2)); b = pxo.Stencil(arg_shape=(n1, n2), kernel=kern, center=(c1, c2)); b._tiled_ok = False
x = torch.randn(batch, n1 * n2, device="cuda", dtype=torch.float32 if dt == np.float32 else torch.float64)
ya, yb = a.apply(x), b.apply(x)
torch.cuda.synchronize()
print("ok", a._tiled_ok, float((ya - yb).norm() / yb.norm()))
''' % ROOT
for args in ["f32 517 1028 9 9 4 4 0 3", "f32 517 1028 9 9 4 3 0 3", "f64 517 1028 9 9 4 4 0 3", "f64 517 1028 9 9 4 3 0 3", "f64 333 260 5 5 1 3 1 3", "f64 333 260 5 5 1 2 1 3",
             "f32 8192 8192 9 9 4 4 0 1", "f32 2048 2048 9 9 4 4 0 1", "f32 4096 4096 9 9 4 4 0 1", "f32 1024 1024 5 5 2 2 1 64"]:
    r = subprocess.run([sys.executable, "-c", code, *args.split()], capture_output=True, text=True)
    print(args, "->", (r.stdout.strip() or r.stderr.strip().splitlines()[-1])[:150], flush=True)
