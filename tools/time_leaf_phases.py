"""Phase clocks of one recursive leaf Cholesky (gpk_debug_leaf_phase_clock): where the 33 us go."""
import ctypes, sys, torch
sys.path.insert(0, ".")
from stheno_b200 import _lib, ops
lib = _lib.load()
buf = torch.zeros(16, dtype=torch.int64, device="cuda")
W0 = (torch.eye(128, device="cuda", dtype=torch.float64) * 4 + 0.01).reshape(1, 128, 128).contiguous()
logdet = torch.zeros(1, device="cuda", dtype=torch.float64); info = torch.zeros(1, device="cuda", dtype=torch.int32)
fn = ops._fn("gpk_potrf", torch.float64)
for rep in range(3):
    W = W0.clone()
    lib.gpk_debug_leaf_phase_clock(ctypes.c_void_p(buf.data_ptr()))
    fn(ops._ptr(W), 128, 128 * 128, 128, 0, ops._ptr(logdet), ops._ptr(info), 1, ops._stream())
    torch.cuda.synchronize()
    lib.gpk_debug_leaf_phase_clock(ctypes.c_void_p(0))
    t = buf.cpu().tolist()
    names = ["load"] + [f"b{b}:{p}" for b in range(4) for p in ("S0", "S1", "S2")][:10] + ["store"]
    d = [t[i + 1] - t[i] for i in range(len(names))]
    print("cycles:", dict(zip(names, d)), "total", t[len(names)] - t[0])
