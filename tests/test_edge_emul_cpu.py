"""The order-free restatement the edge-detector kernels use (litiv_b200/csrc/edge_px.cuh: per-level maps instead of one shared map,
suppression as a pure function of the maps, hysteresis as relaxation sweeps) against the sequential oracle (oracle/lvo_edge_lbsp.hpp),
on the CPU: tests/edge_emul.cpp compiles the kernels' own per-pixel bodies with g++ and drives them with plain loops. The launch code
(grids, indexing, the shared-memory flood) is covered by tests/test_gpu_edge.py."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from litiv_b200.synth import SynthSequence

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def emul(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("edge_emul") / "edge_emul.so")
    cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-I", cuda_inc, "-o", so, os.path.join(HERE, "edge_emul.cpp")])
    L = C.CDLL(so)
    L.emul_create.restype = C.c_void_p
    L.emul_create.argtypes = [C.c_int, C.c_double]
    L.emul_destroy.argtypes = [C.c_void_p]
    L.emul_apply_threshold.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_double]
    L.emul_apply.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
    L.emul_gradient_map.argtypes = [C.c_void_p, C.c_void_p]
    return L


def _frame(seq, t, ch):
    f = np.ascontiguousarray(seq.frame(t))
    if ch == 1 and f.ndim == 3:
        return np.ascontiguousarray(f[..., 0])
    if ch in (2, 4):   # the reference instantiates the detector for 1 to 4 channels (EdgeDetectorLBSP.cpp:144-160)
        extra = (f[..., :1].astype(np.int32) * 3 + f[..., 1:2] * 5 + 17 * t) % 256
        return np.ascontiguousarray(np.concatenate([f, extra.astype(np.uint8)], axis=2)[..., :ch] if ch == 4 else f[..., :2])
    return f


@pytest.mark.parametrize("size", [(96, 72), (97, 73), (96, 73), (97, 72), (43, 41)])
@pytest.mark.parametrize("ch", [1, 2, 3, 4])
@pytest.mark.parametrize("levels", [1, 2, 3])
def test_kernel_bodies_equal_the_sequential_oracle(oracle, emul, size, ch, levels):
    w, h = size
    seq = SynthSequence(w, h, 1 if ch == 1 else 3, seed=w + h + ch)
    o, e = oracle.EdgeDetectorLBSPOracle(levels=levels), emul.emul_create(levels, 0.5)
    try:
        for t, thr in [(3, 0.5), (5, 0.25), (7, 0.75), (9, 0.0), (11, 0.9), (12, -1.0)]:   # one object, a sequence of calls (the maps persist)
            f = _frame(seq, t, ch)
            want = o.apply_threshold(f, thr)
            got, grad = np.empty((h, w), np.uint8), np.empty((h, w, 4), np.uint8)
            emul.emul_apply_threshold(e, f.ctypes.data, w, h, ch, got.ctypes.data, thr)
            emul.emul_gradient_map(e, grad.ctypes.data)
            assert np.array_equal(grad, o.gradient_map(f.shape)), (t, thr)
            assert np.array_equal(got, want), (t, thr, int((got != want).sum()))
        f = _frame(seq, 14, ch)
        got = np.empty((h, w), np.uint8)
        emul.emul_apply(e, f.ctypes.data, w, h, ch, got.ctypes.data)
        want = o.apply(f)
        assert np.array_equal(got, want) and want.any()
    finally:
        emul.emul_destroy(e)


def test_blocky_noise_with_many_edges(oracle, emul):
    rng = np.random.default_rng(5)
    for (w, h), ch, levels in [((64, 49), 3, 3), ((65, 48), 1, 2), ((51, 51), 3, 1)]:
        o, e = oracle.EdgeDetectorLBSPOracle(levels=levels), emul.emul_create(levels, 0.5)
        for thr in (0.5, 0.2, 0.05, 0.8, 0.0):
            base = rng.integers(0, 256, (h // 4 + 2, w // 4 + 2, ch), dtype=np.uint8)
            f = np.kron(base, np.ones((4, 4, 1), np.uint8))[:h, :w]
            f = np.ascontiguousarray((f.astype(int) + rng.integers(-8, 9, f.shape)).clip(0, 255).astype(np.uint8))
            f = np.ascontiguousarray(f[..., 0]) if ch == 1 else f
            want, got = o.apply_threshold(f, thr), np.empty((h, w), np.uint8)
            emul.emul_apply_threshold(e, f.ctypes.data, w, h, ch, got.ctypes.data, thr)
            assert np.array_equal(got, want) and want.any()
            # gradient rows H-2, H-1 lie in the LBSP border (magnitude 0, never a maximum): mask rows H-4, H-3 are always "no edge", which
            # walls the two rows the suppression loop never writes off from every seed -- they stay empty whatever the call history
            assert not want[-4:].any()
        emul.emul_destroy(e)


def test_random_small_images_levels_and_hysteresis_factors(oracle, emul):
    """edge cases of the domain: the smallest legal sizes per level count (5, 9, 17, 33 px), four levels, hysteresis factors near 0 and 1,
    every threshold of the 0..16 scale, random content (many short edges, seeds next to image borders)"""
    rng = np.random.default_rng(11)
    cases = [((5, 5), 1), ((6, 5), 1), ((9, 9), 2), ((10, 9), 2), ((17, 17), 3), ((18, 19), 3), ((33, 35), 4), ((40, 33), 4), ((31, 47), 2)]
    for (w, h), levels in cases:
        for ch in (1, 3):
            for hyst in (0.5, 0.05, 0.95):
                o, e = oracle.EdgeDetectorLBSPOracle(levels=levels, hyst_low_factor=hyst), emul.emul_create(levels, hyst)
                for k in range(17):
                    f = rng.integers(0, 256, (h, w, ch), dtype=np.uint8)
                    f[h // 3:, w // 2:] //= 4                                   # one large-scale structure besides the noise
                    f = np.ascontiguousarray(f[..., 0]) if ch == 1 else np.ascontiguousarray(f)
                    want, got = o.apply_threshold(f, k / 16.0), np.empty((h, w), np.uint8)
                    emul.emul_apply_threshold(e, f.ctypes.data, w, h, ch, got.ctypes.data, k / 16.0)
                    assert np.array_equal(got, want), (w, h, levels, ch, hyst, k)
                emul.emul_destroy(e)
    with pytest.raises(RuntimeError):
        oracle.EdgeDetectorLBSPOracle(levels=3).apply_threshold(np.zeros((16, 16), np.uint8))   # 16 -> 8 -> 4 < the 5x5 patch
