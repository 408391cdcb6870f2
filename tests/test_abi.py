"""The C-ABI library builds, loads and exports every symbol ``include/gpk.h`` declares (no compute: no GPU needed)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "gpk.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gpk_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree():
    from stheno_b200 import _lib

    decl = declared_symbols()
    assert len(decl) >= 25
    assert sorted(_lib.EXPORTED) == decl


def test_library_loads_and_exports_everything():
    from stheno_b200 import _lib

    lib = _lib.load()
    for name in declared_symbols():
        assert hasattr(lib, name), name
    assert lib.gpk_version() == 100
    assert lib.gpk_round_up(1) == 128 and lib.gpk_round_up(128) == 128 and lib.gpk_round_up(16385) == 16512


def test_no_cpu_fallback():
    import numpy as np
    import torch

    import stheno_b200 as S

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    f = S.GP(S.EQ())
    with pytest.raises(RuntimeError, match="CUDA"):
        f(np.linspace(0, 1, 5), 0.1).logpdf(np.zeros(5))
    from stheno_b200 import ops

    with pytest.raises(RuntimeError, match="CUDA"):
        ops.kernel_matrix(ops.FlatKernel([(1.0, [("eq", 0)])], 1), torch.zeros(1, 1, 4, 1, dtype=torch.float64))


def test_flat_kernel_descriptor():
    import stheno_b200 as S

    k = 2.0 * S.EQ().stretch(1.5) + S.Matern32() * S.Linear().stretch(0.5) + 0.1 * S.Delta()
    flat, scales = k._flat()
    assert len(flat.terms) == 3 and flat.n_groups == 3
    d = flat.desc()
    assert d.n_terms == 3 and list(d.term_begin[:4]) == [0, 1, 3, 4]
    assert [d.fac_kind[i] for i in range(4)] == [0, 2, 4, 5]
    assert abs(d.coef[0] - 2.0) < 1e-15 and abs(d.coef[2] - 0.1) < 1e-15
    # algebraic simplifications of the reference's `algebra` package that the model layer relies on
    assert isinstance(0 * S.EQ(), S.ZeroKernel) and (S.EQ() + S.ZeroKernel()).render() == "EQ()"
    assert (S.EQ() * (S.EQ() + S.Matern12())).flat_terms() is not None
    assert len(((S.EQ() + S.Matern12()) * (S.EQ() + S.Linear())).flat_terms()) == 4


def test_emulated_gemm_tile_order_is_a_bijection():
    """The banded, column-major tile order of the emulation GEMM (``oz_tile`` in csrc/gemm_oz.cu), evaluated on the HOST
    through ``gpk_debug_oz_tile`` -- the very function the kernel runs: every tile of a launch exactly once, inside the
    lower triangle in lower mode, for square / tall / ragged shapes and band heights 1..16 (no GPU needed; a wrong order
    would corrupt or hang a launch on the device)."""
    import ctypes

    from stheno_b200 import _lib

    lib = _lib.load()
    tm, tn = ctypes.c_int32(), ctypes.c_int32()
    shapes = [(1, 2), (2, 4), (3, 2), (5, 10), (16, 32), (17, 34), (33, 8), (40, 80), (64, 16), (114, 228), (120, 8)]
    for lower in (0, 1):
        for tiles_m, tiles_n in shapes:
            if lower and tiles_n > 2 * tiles_m:
                continue
            for band in (1, 3, 16):
                total = lib.gpk_debug_oz_tile(lower, tiles_m, tiles_n, band, -1, None, None)
                tri = min(tiles_m, tiles_n // 2) if lower else 0
                assert total == (tri * (tri + 1) + (tiles_m - tri) * tiles_n if lower else tiles_m * tiles_n)
                seen = set()
                for t in range(total):
                    lib.gpk_debug_oz_tile(lower, tiles_m, tiles_n, band, t, ctypes.byref(tm), ctypes.byref(tn))
                    a, b = tm.value, tn.value
                    assert 0 <= a < tiles_m and 0 <= b < tiles_n, (lower, tiles_m, tiles_n, band, t, a, b)
                    assert not lower or b < min(tiles_n, 2 * (a + 1)), (lower, tiles_m, tiles_n, band, t, a, b)
                    seen.add((a, b))
                assert len(seen) == total, (lower, tiles_m, tiles_n, band)
