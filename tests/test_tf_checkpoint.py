"""bayesgm_b200/tf_checkpoint.py against bundles written by tests/tf_bundle_writer.py (an independent
restatement of the container format; NOT TensorFlow output -- see the reader's STATUS note)."""
import os

import numpy as np
import pytest

from bayesgm_b200 import tf_checkpoint as tfc
from tf_bundle_writer import write_bundle
from helpers import causal_params

SUF = '/.ATTRIBUTES/VARIABLE_VALUE'


def _reference_names(model):
    """Names an object-based checkpoint gives the variables of BaseFullyConnectedNet (all_layers[i] =
    [Dense, BatchNormalization], networks/base.py:18-36) plus Adam slots and the save counter."""
    rs = np.random.RandomState(0)
    tensors, want = {}, {}
    for root in ('g_net', 'e_net', 'f_net', 'h_net'):
        net = getattr(model, root)
        ws = []
        for i, (W, b) in enumerate(net.layers):
            W2 = rs.standard_normal(W.shape).astype(np.float32)
            b2 = rs.standard_normal(b.shape).astype(np.float32)
            tensors['%s/all_layers/%d/0/kernel%s' % (root, i, SUF)] = W2
            tensors['%s/all_layers/%d/0/bias%s' % (root, i, SUF)] = b2
            tensors['%s/all_layers/%d/0/kernel/.OPTIMIZER_SLOT/g_optimizer/m%s' % (root, i, SUF)] = np.zeros_like(W2)
            ws += [W2, b2]
        want[root] = ws
    tensors['save_counter' + SUF] = np.array(3, np.int64)
    tensors['g_optimizer/iter' + SUF] = np.array(1200, np.int64)
    return tensors, want


@pytest.mark.parametrize("compress,per_block", [(False, 5), (True, 3), (False, 1000)])
def test_bundle_round_trip_and_model_import(tmp_path, compress, per_block):
    from bayesgm_b200 import CausalBGM
    params = causal_params(30, [1, 1, 1, 2], g_units=(64, 64, 64, 64, 64, 64, 64, 64, 64, 64, 64, 64))   # > 10 layers: 10 sorts after 2
    m = CausalBGM(params=params)
    tensors, want = _reference_names(m)
    prefix = str(tmp_path / 'ckpt-3')
    write_bundle(prefix, tensors, per_block=per_block, compress=compress)
    got = tfc.read_checkpoint(prefix)
    assert set(got) == set(tensors)                       # the string tensor (object graph) is skipped
    for k in tensors:
        assert got[k].dtype == tensors[k].dtype and np.array_equal(got[k], tensors[k]), k
    done = tfc.load_tf_checkpoint(m, prefix)
    assert done == ['g_net', 'e_net', 'f_net', 'h_net']
    for root in want:
        for a, b in zip(getattr(m, root).get_weights(), want[root]):
            assert np.array_equal(a, b)
    # CheckpointManager state file / directory form
    with open(str(tmp_path / 'checkpoint'), 'w') as f:
        f.write('model_checkpoint_path: "ckpt-3"\nall_model_checkpoint_paths: "ckpt-3"\n')
    assert tfc.latest_checkpoint(str(tmp_path)) == prefix
    assert set(tfc.read_checkpoint(str(tmp_path))) == set(tensors)


def test_import_fails_loudly_on_a_shape_mismatch_or_bad_file(tmp_path):
    from bayesgm_b200 import CausalBGM
    m = CausalBGM(params=causal_params(30, [1, 1, 1, 2]))
    tensors, _ = _reference_names(m)
    tensors['g_net/all_layers/0/0/kernel' + SUF] = np.zeros((7, 64), np.float32)
    prefix = str(tmp_path / 'ckpt-1')
    write_bundle(prefix, tensors)
    with pytest.raises(ValueError, match="g_net"):
        tfc.load_tf_checkpoint(m, prefix)
    with open(str(tmp_path / 'junk.index'), 'wb') as f:
        f.write(b'\0' * 100)
    with pytest.raises(ValueError, match="magic"):
        tfc.read_checkpoint(str(tmp_path / 'junk'))
    with pytest.raises(NotImplementedError):
        tfc.load_tf_checkpoint(CausalBGM(params=causal_params(30, [1, 1, 1, 2], use_bnn=True)), prefix)


def test_snappy_copy_elements():
    # "abc" literal, then a 1-byte-offset copy of 6 bytes from 3 back (overlapping), then a 2-byte-offset copy
    src = bytes([12]) + bytes([2 << 2]) + b'abc' + bytes([((6 - 4) << 2) | 1, 3]) + bytes([((3 - 1) << 2) | 2, 9, 0])
    assert tfc._snappy_decompress(src) == b'abcabcabcabc'
    with pytest.raises(ValueError):
        tfc._snappy_decompress(bytes([5, 0 << 2]) + b'a' + bytes([1, 9]))
