#!/bin/bash
# strip K1 kernel: ragged / padded shapes first, each under a tight timeout (a hang costs a minute, not the lease)
set -u
cat > /tmp/k1shapes.py <<'P'
import sys, numpy as np, torch
sys.path.insert(0, ".")
from stheno_b200 import ops
from oracle import gp_oracle as O
rng = np.random.default_rng(0)
bad = 0
for n, d in [(3, 1), (5, 1), (1, 1), (63, 1), (64, 1), (65, 1), (130, 1), (3, 2), (7, 3), (100, 3), (129, 8), (1000, 1), (1030, 2), (600, 5)]:
    x = rng.standard_normal((n, d))
    xg = torch.as_tensor(x, device="cuda")[None, None]
    for kind in ("eq", "matern52"):
        flat = ops.FlatKernel([(1.3, [(kind, 0)])], 1)
        ref = 1.3 * O.kernel_matrix((kind,), x) + 0.25 * np.eye(n)
        K = ops.kernel_matrix(flat, xg, noise_scalar=0.25)[0].cpu().numpy()
        ch = ops.chol_from_kernel(flat, xg, noise_scalar=0.25, jitter=0.0)   # LOWER | SAME | PAD_IDENTITY
        torch.cuda.synchronize()
        W = ch.W[0, : ch.n_pad].cpu().numpy()
        Lref = np.linalg.cholesky(ref)
        e1 = np.abs(K - ref).max(); e2 = np.abs(np.tril(W[:n, :n]) - Lref).max()
        padok = np.allclose(np.tril(W)[n:, :], np.eye(ch.n_pad)[n:, :])
        if e1 > 1e-12 or e2 > 1e-10 or not padok: bad += 1; print("BAD", n, d, kind, e1, e2, padok)
    # rectangular with zero padding (kernel rows for the solves)
    xs = rng.standard_normal((n + 3, d)); xsg = torch.as_tensor(xs, device="cuda")[None, None]
    out = ops.kernel_rows_padded(flat, xsg, xg, ch)[0].cpu().numpy()
    ref = 1.3 * O.kernel_matrix((kind,), xs, x)
    if np.abs(out[: n + 3, :n] - ref).max() > 1e-12 or np.abs(out[n + 3 :, :]).max() > 0 or np.abs(out[:, n:]).max() > 0: bad += 1; print("BAD rows", n, d)
print("shapes done, bad =", bad)
P
timeout 100 python /tmp/k1shapes.py || echo "SHAPES FAILED/TIMED OUT rc=$?"
echo "== K1 A/B"; timeout 200 python tools/time_k1.py 2>&1 | python -c "
import sys, json
for line in sys.stdin:
    try: d = json.loads(line)
    except Exception: print(line[:300]); continue
    print({k: (round(v, 4) if isinstance(v, float) else v) for k, v in d.items() if 'ms' in k or 'variant' in k or 'GBs' in k or 'err' in k})
"
GPK_K1_ONE_TILE=1 timeout 100 python - <<'P'
import sys, torch
sys.path.insert(0, ".")
from stheno_b200 import ops
g = torch.Generator(device="cuda").manual_seed(1); n = 16384
x = torch.randn(1, 1, n, 8, device="cuda", dtype=torch.float64, generator=g) / 2.0
flat = ops.FlatKernel([(1.3, [("eq", 0)])], 1); W = torch.empty(1, n, n, device="cuda", dtype=torch.float64)
run = lambda: ops._km_launch(flat, x, x, n, n, 8, ops.KM_LOWER | ops.KM_SAME | ops.KM_PAD_IDENTITY, 0.1, None, 1e-12, W, n, n * n, 1)
for _ in range(3): run()
torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): run()
e1.record(); torch.cuda.synchronize(); print("one-tile kernel eq lower ms", e0.elapsed_time(e1) / 10)
P
echo "== tests (tight timeouts)"
timeout 200 python -m pytest tests/test_gpu_primitives.py tests/test_model.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -3
timeout 300 python -m pytest tests/test_configs.py tests/test_sparse_streamed.py tests/test_emulation.py tests/test_input_maps.py tests/test_derivatives.py tests/test_advice_r1.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -3
