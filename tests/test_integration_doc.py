"""INTEGRATION.md section B (the ctypes stub a maintainer would add to the reference) must agree with
include/bgm_b200.h: every call site has the header's arity and argtypes (CPU), and the stub runs
and reproduces `bayesgm_b200.CausalBGM.get_log_posterior` (GPU)."""
import ast
import os
import re

import numpy as np
import pytest

from bayesgm_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def stub_source():
    md = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    blocks = re.findall(r"```python\n(.*?)```", md, flags=re.S)
    stub = [b for b in blocks if "C.CDLL" in b]
    assert len(stub) == 1, "expected exactly one ctypes stub in INTEGRATION.md"
    return stub[0]


def test_stub_calls_have_header_arity():
    tree = ast.parse(stub_source())
    seen = {}
    for node in ast.walk(tree):
        if isinstance(node, ast.Call) and isinstance(node.func, ast.Attribute) \
                and isinstance(node.func.value, ast.Name) and node.func.value.id == "_lib" \
                and node.func.attr.startswith("bgm_"):
            seen.setdefault(node.func.attr, []).append(len(node.args))
    assert {"bgm_causal_create", "bgm_causal_logpost", "bgm_causal_project", "bgm_causal_info"} <= set(seen)
    for name, counts in seen.items():
        want = len(_lib.SYMBOLS[name][1])
        assert all(c == want for c in counts), "%s called with %s args, header takes %d" % (name, counts, want)
    # argtypes lists written in the stub have the header's length too
    for node in ast.walk(tree):
        if isinstance(node, ast.Assign) and isinstance(node.targets[0], ast.Attribute) \
                and node.targets[0].attr == "argtypes":
            name = node.targets[0].value.attr
            assert isinstance(node.value, ast.List)
            n = 0
            for e in node.value.elts:          # `[X] * k` does not occur; plain lists only
                n += 1
            assert n == len(_lib.SYMBOLS[name][1]), "%s argtypes: %d entries, header takes %d" % (
                name, n, len(_lib.SYMBOLS[name][1]))


@pytest.mark.gpu
@pytest.mark.parametrize("v_dim", [200, 40])     # projected (v_dim > H + 8) and direct last layer
def test_stub_runs_and_matches_product(v_dim):
    from helpers import causal_params, causal_nets, causal_data, product_model
    src = stub_source().replace('"libbgm_b200.so"', repr(_lib.LIB_PATH))
    ns = {}
    exec(compile(src, "INTEGRATION.md", "exec"), ns)
    params = causal_params(v_dim, [1, 1, 1, 2])
    nets = causal_nets(params)

    class FakeKeras(object):                       # what the stub reads from a Keras net
        def __init__(self, layers):
            self.all_layers = list(layers)
            self._w = [a for W, b in layers for a in (W, b)]

        def get_weights(self):
            return self._w

    ref = ns["CausalBGM"]()
    ref.params = params
    ref.g_net, ref.f_net, ref.h_net = FakeKeras(nets['g']), FakeKeras(nets['f']), FakeKeras(nets['h'])
    x, y, v = causal_data(300, v_dim)
    z = np.random.RandomState(1).standard_normal((300, 5)).astype(np.float32)
    got = ref.get_log_posterior(x, y, v, z)
    got2 = ref.get_log_posterior(x, y, v, z)       # cached handle
    want = product_model(params, nets).get_log_posterior(x, y, v, z)
    np.testing.assert_array_equal(got, want)
    np.testing.assert_array_equal(got2, want)
    ref.invalidate_weights()
