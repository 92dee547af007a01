"""Reader for TensorFlow checkpoints (`tf.train.Checkpoint` / `CheckpointManager`, the format the
reference saves with `save_model: True`, causalbgm/base.py:112-127, :529) without TensorFlow.

A checkpoint `<prefix>` is a "tensor bundle": `<prefix>.index`, a LevelDB-format sorted string table
mapping tensor names to `BundleEntryProto` records (dtype, shape, shard, offset, size), and
`<prefix>.data-0000k-of-0000n` shards holding the raw little-endian tensor bytes.  Object-based
checkpoints name a variable by its attribute path from the root object,
`g_net/all_layers/0/0/kernel/.ATTRIBUTES/VARIABLE_VALUE` for the `[Dense, BatchNormalization]` pairs of
`BaseFullyConnectedNet.all_layers` (networks/base.py:18-36).

`read_checkpoint(prefix)` returns {name: ndarray} for every numeric tensor and depends on the container
format only.  `load_tf_checkpoint(model, prefix)` maps those names onto a model's deterministic nets:
under each net's root attribute the Dense kernels / biases are ordered by the integers in their path and
their shapes must chain input_dim -> units -> output_dim (anything else raises, listing the names found,
so the mapping can be done by hand from `read_checkpoint`).

STATUS: written from the published format (tensorflow/core/util/tensor_bundle, leveldb table_format.md,
snappy format_description.txt).  TensorFlow is not installable in this environment, so the reader has
only been exercised against bundles produced by the independent writer in tests/tf_bundle_writer.py,
never against a file written by TensorFlow itself.
"""
import glob
import os
import re
import struct

import numpy as np

_MAGIC = 0xdb4775248b80fb57
_DTYPES = {1: np.float32, 2: np.float64, 3: np.int32, 4: np.uint8, 5: np.int16, 6: np.int8, 9: np.int64,
           10: np.bool_, 17: np.uint16, 19: np.float16, 22: np.uint32, 23: np.uint64}


def _varint(buf, pos):
    out, shift = 0, 0
    while True:
        b = buf[pos]
        pos += 1
        out |= (b & 0x7F) << shift
        if not b & 0x80:
            return out, pos
        shift += 7


def _snappy_decompress(src):
    n, pos = _varint(src, 0)
    out = bytearray()
    while pos < len(src):
        tag = src[pos]
        pos += 1
        kind = tag & 3
        if kind == 0:                                   # literal
            ln = tag >> 2
            if ln >= 60:
                nb = ln - 59
                ln = int.from_bytes(src[pos:pos + nb], 'little')
                pos += nb
            ln += 1
            out += src[pos:pos + ln]
            pos += ln
            continue
        if kind == 1:
            ln = ((tag >> 2) & 7) + 4
            off = ((tag >> 5) << 8) | src[pos]
            pos += 1
        elif kind == 2:
            ln = (tag >> 2) + 1
            off = int.from_bytes(src[pos:pos + 2], 'little')
            pos += 2
        else:
            ln = (tag >> 2) + 1
            off = int.from_bytes(src[pos:pos + 4], 'little')
            pos += 4
        if off == 0 or off > len(out):
            raise ValueError("corrupt snappy block")
        for _ in range(ln):                             # copies may overlap their own output
            out.append(out[-off])
    if len(out) != n:
        raise ValueError("corrupt snappy block (length)")
    return bytes(out)


def _block(data, offset, size):
    raw = data[offset:offset + size]
    ctype = data[offset + size]                         # 1-byte type + 4-byte crc follow the block
    if ctype == 1:
        raw = _snappy_decompress(raw)
    elif ctype != 0:
        raise ValueError("unknown block compression %d" % ctype)
    n_restarts = struct.unpack_from('<I', raw, len(raw) - 4)[0]
    end = len(raw) - 4 - 4 * n_restarts
    pos, key, out = 0, b'', []
    while pos < end:
        shared, pos = _varint(raw, pos)
        unshared, pos = _varint(raw, pos)
        vlen, pos = _varint(raw, pos)
        key = key[:shared] + raw[pos:pos + unshared]
        pos += unshared
        out.append((key, raw[pos:pos + vlen]))
        pos += vlen
    return out


def _table(data):
    """All (key, value) pairs of a LevelDB-format table, in key order."""
    if len(data) < 48 or struct.unpack_from('<Q', data, len(data) - 8)[0] != _MAGIC:
        raise ValueError("not a TensorFlow checkpoint index (bad table magic)")
    footer = data[-48:]
    _, p = _varint(footer, 0)                           # metaindex handle
    _, p = _varint(footer, p)
    ioff, p = _varint(footer, p)
    isz, p = _varint(footer, p)
    out = []
    for _, handle in _block(data, ioff, isz):
        off, q = _varint(handle, 0)
        sz, q = _varint(handle, q)
        out += _block(data, off, sz)
    return out


def _proto_fields(buf):
    pos, out = 0, []
    while pos < len(buf):
        tag, pos = _varint(buf, pos)
        num, wt = tag >> 3, tag & 7
        if wt == 0:
            v, pos = _varint(buf, pos)
        elif wt == 1:
            v = buf[pos:pos + 8]
            pos += 8
        elif wt == 2:
            ln, pos = _varint(buf, pos)
            v = buf[pos:pos + ln]
            pos += ln
        elif wt == 5:
            v = buf[pos:pos + 4]
            pos += 4
        else:
            raise ValueError("unsupported protobuf wire type %d" % wt)
        out.append((num, wt, v))
    return out


def _entry(buf):
    """BundleEntryProto: dtype = 1, shape = 2, shard_id = 3, offset = 4, size = 5, crc32c = 6, slices = 7."""
    e = dict(dtype=0, shape=[], shard=0, offset=0, size=0, sliced=False)
    for num, wt, v in _proto_fields(buf):
        if num == 1:
            e['dtype'] = v
        elif num == 2:
            for n2, _, v2 in _proto_fields(v):           # TensorShapeProto: repeated Dim dim = 2 {int64 size = 1}
                if n2 == 2:
                    size = 0
                    for n3, _, v3 in _proto_fields(v2):
                        if n3 == 1:
                            size = v3
                    e['shape'].append(size)
        elif num == 3:
            e['shard'] = v
        elif num == 4:
            e['offset'] = v
        elif num == 5:
            e['size'] = v
        elif num == 7:
            e['sliced'] = True
    return e


def latest_checkpoint(directory):
    """`tf.train.latest_checkpoint`: the prefix named by the `checkpoint` state file, else the highest ckpt-N."""
    state = os.path.join(directory, 'checkpoint')
    if os.path.exists(state):
        m = re.search(r'^model_checkpoint_path:\s*"(.*)"', open(state).read(), flags=re.M)
        if m:
            p = m.group(1)
            return p if os.path.isabs(p) else os.path.join(directory, p)
    found = glob.glob(os.path.join(directory, '*.index'))
    if not found:
        return None
    num = lambda f: int((re.findall(r'(\d+)\.index$', f) or ['-1'])[0])
    return max(found, key=num)[:-len('.index')]


def read_checkpoint(prefix):
    """{tensor name: ndarray} of every numeric, unsliced tensor of the bundle `<prefix>.index` + data shards."""
    if os.path.isdir(prefix):
        prefix = latest_checkpoint(prefix)
    with open(prefix + '.index', 'rb') as f:
        pairs = _table(f.read())
    num_shards = 1
    out, shards = {}, {}
    for key, val in pairs:
        if key == b'':                                   # BundleHeaderProto: num_shards = 1, endianness = 2
            for num, _, v in _proto_fields(val):
                if num == 1:
                    num_shards = v
                if num == 2 and v != 0:
                    raise ValueError("big-endian checkpoints are not supported")
            continue
        e = _entry(val)
        dt = _DTYPES.get(e['dtype'])
        if dt is None or e['sliced']:
            continue                                     # strings (the object graph), variants, partitioned variables
        if e['shard'] not in shards:
            with open('%s.data-%05d-of-%05d' % (prefix, e['shard'], num_shards), 'rb') as f:
                shards[e['shard']] = f.read()
        raw = shards[e['shard']][e['offset']:e['offset'] + e['size']]
        a = np.frombuffer(raw, dtype=dt)
        if a.size != int(np.prod(e['shape'], dtype=np.int64)):
            raise ValueError("tensor %r: %d bytes do not match shape %r" % (key, e['size'], e['shape']))
        out[key.decode('utf-8', 'replace')] = a.reshape(e['shape']).copy()
    return out


_SUFFIX = '/.ATTRIBUTES/VARIABLE_VALUE'


def _net_variables(tensors, root):
    """[(path ints, leaf name, array)] of the model variables under `<root>/`, optimizer slots excluded."""
    out = []
    for name, a in tensors.items():
        if not name.startswith(root + '/') or not name.endswith(_SUFFIX) or '.OPTIMIZER_SLOT' in name:
            continue
        parts = name[len(root) + 1:-len(_SUFFIX)].split('/')
        ints = tuple(int(x) for p in parts for x in re.findall(r'\d+', p))
        out.append((ints, parts[-1], a, name))
    return sorted(out, key=lambda t: (t[0], t[1]))


def _dense_stack(tensors, root, dims):
    found = _net_variables(tensors, root)
    kernels = [t for t in found if t[1] == 'kernel']
    biases = [t for t in found if t[1] == 'bias']
    want = list(zip(dims[:-1], dims[1:]))
    names = [t[3] for t in found]
    if len(kernels) != len(want) or len(biases) != len(want):
        raise ValueError("%s: expected %d Dense layers, checkpoint has kernels %r" % (root, len(want), names))
    out = []
    for (k, b, (fi, fo)) in zip(kernels, biases, want):
        if k[2].shape != (fi, fo) or b[2].shape != (fo,):
            raise ValueError("%s: layer shapes %r / %r do not match (%d, %d); checkpoint names: %r"
                             % (root, k[2].shape, b[2].shape, fi, fo, names))
        out += [k[2].astype(np.float32), b[2].astype(np.float32)]
    return out


def load_tf_checkpoint(model, prefix):
    """Loads g / e / f / h (and dz, prior_net where the model has them) of a deterministic-net CausalBGM-family
    model from a reference checkpoint.  Returns the list of net roots that were loaded."""
    if getattr(model, '_bnn', False):
        raise NotImplementedError("bayesgm_b200: checkpoint import covers the deterministic nets (use_bnn=False); "
                                  "read_checkpoint() gives the raw tensors of a Bayesian model")
    tensors = read_checkpoint(prefix)
    loaded = {}
    for root in ('g_net', 'e_net', 'f_net', 'h_net'):
        net = getattr(model, root)
        loaded[root[0]] = _dense_stack(tensors, root, net.dims)
    model.set_weights(**loaded)
    done = ['g_net', 'e_net', 'f_net', 'h_net']
    if hasattr(model, 'prior_net'):
        model.prior_net.set_weights(_dense_stack(tensors, 'prior_net', model.prior_net.dims))
        done.append('prior_net')
    dz = getattr(model, 'dz_net', None)
    if dz is not None and any(n.startswith('dz_net/') for n in tensors):
        try:
            model.set_weights(dz=_disc_stack(tensors, 'dz_net', dz))
            done.append('dz_net')
        except ValueError:
            pass                                         # the discriminator is not needed for the posterior path
    return done


def _disc_stack(tensors, root, dz):
    """Discriminator (networks/base.py:338-385): [Dense, BatchNormalization] pairs then a Dense(1)."""
    found = _net_variables(tensors, root)
    want = dz.trainable_list()
    by_leaf = {}
    for t in found:
        by_leaf.setdefault(t[1], []).append(t[2])
    order = []
    kernels, biases = by_leaf.get('kernel', []), by_leaf.get('bias', [])
    gammas, betas = by_leaf.get('gamma', []), by_leaf.get('beta', [])
    n_hidden = len(gammas)
    if len(kernels) != n_hidden + 1:
        raise ValueError("%s: unexpected layer structure" % root)
    for i in range(n_hidden):
        order += [kernels[i], biases[i], gammas[i], betas[i]]
    order += [kernels[-1], biases[-1]]
    if len(order) != len(want) or any(a.shape != b.shape for a, b in zip(order, want)):
        raise ValueError("%s: shapes do not match the model" % root)
    return [a.astype(np.float32) for a in order]
