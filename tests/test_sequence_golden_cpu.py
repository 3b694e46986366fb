"""The oracle against the committed sequence fixtures (tests/golden/sequence_golden.npz, made by tests/golden/make_sequence_golden.py):
every algorithm's snapshot-mode masks, final model state and the edge-detector / LBSP-gradient outputs on integer-generated sequences.
The reference holds no golden vector for these paths; the fixtures guard OUR parity anchor against drift."""
import importlib.util
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_oracle_reproduces_the_committed_sequence_fixtures(oracle):
    spec = importlib.util.spec_from_file_location("make_sequence_golden", os.path.join(GOLDEN, "make_sequence_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    g = np.load(os.path.join(GOLDEN, "sequence_golden.npz"))
    res = mod.run_cases()
    assert len(res) == 11 and {k.split("__")[0] for k in g.files} == set(res)
    for name, (masks_sha, last, state_sha) in res.items():
        assert np.array_equal(last, g[name + "__last_mask"]), f"{name}: last mask differs in {(last != g[name + '__last_mask']).sum()} px"
        assert masks_sha == str(g[name + "__masks_sha256"]), f"{name}: mask sequence digest"
        assert state_sha == str(g[name + "__state_sha256"]), f"{name}: final state digest"
        assert (last > 0).any(), name


def test_fixture_inputs_are_integer_generated_and_stable():
    spec = importlib.util.spec_from_file_location("make_sequence_golden", os.path.join(GOLDEN, "make_sequence_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    a, b = mod.frames(3, 103), mod.frames(3, 103)
    assert all(np.array_equal(x, y) for x, y in zip(a, b)) and a[0].dtype == np.uint8 and a[0].shape == (72, 96, 3)
    assert mod.sha(*a) == "%s" % mod.sha(*b)


def test_method_name_protocol_reproduces_the_fixtures_with_the_oracle(oracle):
    """run_mask_cases() is what the GPU test drives with the CUDA classes; here the same code runs over the oracle classes and must land
    on the committed digests (so that a GPU failure of that test can only come from the CUDA path, not from the protocol)"""
    spec = importlib.util.spec_from_file_location("make_sequence_golden", os.path.join(GOLDEN, "make_sequence_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    g = np.load(os.path.join(GOLDEN, "sequence_golden.npz"))
    res = mod.run_mask_cases(mod.oracle_factories())
    assert len(res) == 11
    for name, (masks_sha, last, extra) in res.items():
        assert masks_sha == str(g[name + "__masks_sha256"]) and np.array_equal(last, g[name + "__last_mask"]), name
        if extra is not None:
            assert extra == str(g[name + "__state_sha256"]), name
