"""Test-only writer of TensorFlow "tensor bundle" checkpoints, written independently of the reader in
bayesgm_b200/tf_checkpoint.py from the same published format descriptions (leveldb table_format.md,
tensorflow/core/protobuf/tensor_bundle.proto).  Not produced by TensorFlow: see the STATUS note of the reader."""
import struct

import numpy as np

_DT = {np.dtype('float32'): 1, np.dtype('float64'): 2, np.dtype('int32'): 3, np.dtype('int64'): 9}


def varint(v):
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def field(num, wt, payload):
    return varint((num << 3) | wt) + payload


def entry_proto(arr, shard, offset):
    shape = b''.join(field(2, 2, varint(len(d)) + d) for d in (field(1, 0, varint(int(s))) for s in arr.shape))
    return (field(1, 0, varint(_DT[arr.dtype])) + field(2, 2, varint(len(shape)) + shape) +
            (field(3, 0, varint(shard)) if shard else b'') + field(4, 0, varint(offset)) +
            field(5, 0, varint(arr.nbytes)) + field(6, 5, struct.pack('<I', 0)))


def snappy_literals(raw):
    """A valid snappy stream made of literal elements only."""
    out = bytearray(varint(len(raw)))
    for i in range(0, len(raw), 60):
        chunk = raw[i:i + 60]
        out.append((len(chunk) - 1) << 2)
        out += chunk
    return bytes(out)


def block(pairs, restart_interval=16, compress=False):
    body, restarts, prev = bytearray(), [], b''
    for i, (k, v) in enumerate(pairs):
        shared = 0
        if i % restart_interval == 0:
            restarts.append(len(body))
        else:
            while shared < min(len(prev), len(k)) and prev[shared] == k[shared]:
                shared += 1
        body += varint(shared) + varint(len(k) - shared) + varint(len(v)) + k[shared:] + v
        prev = k
    if not restarts:
        restarts = [0]
    body += b''.join(struct.pack('<I', r) for r in restarts) + struct.pack('<I', len(restarts))
    if compress:
        return snappy_literals(bytes(body)), 1
    return bytes(body), 0


def write_bundle(prefix, tensors, per_block=5, compress=False, extra_string_entry=True):
    """tensors: {name: ndarray}.  Writes <prefix>.index and <prefix>.data-00000-of-00001."""
    data = bytearray()
    pairs = [(b'', field(1, 0, varint(1)) + field(3, 2, varint(2) + field(1, 0, varint(1))))]   # header: 1 shard, version
    items = {k.encode(): v for k, v in tensors.items()}
    if extra_string_entry:      # the object graph of a real checkpoint is a DT_STRING tensor: readers must skip it
        items[b'_CHECKPOINTABLE_OBJECT_GRAPH'] = None
    for name in sorted(items):
        arr = items[name]
        if arr is None:
            blob = b'\x05hello'
            pairs.append((name, field(1, 0, varint(7)) + field(2, 2, varint(0)) + field(4, 0, varint(len(data))) +
                          field(5, 0, varint(len(blob)))))
            data += blob
            continue
        arr = np.asarray(arr)            # (ascontiguousarray would turn a scalar into shape (1,))
        pairs.append((name, entry_proto(arr, 0, len(data))))
        data += arr.tobytes()
    out = bytearray()
    index_pairs = []
    for i in range(0, len(pairs), per_block):
        chunk = pairs[i:i + per_block]
        raw, ctype = block(chunk, compress=compress)
        index_pairs.append((chunk[-1][0], varint(len(out)) + varint(len(raw))))
        out += raw + bytes([ctype]) + b'\0\0\0\0'
    meta, _ = block([])
    meta_handle = varint(len(out)) + varint(len(meta))
    out += meta + b'\0' + b'\0\0\0\0'
    idx, _ = block(index_pairs, restart_interval=1)
    idx_handle = varint(len(out)) + varint(len(idx))
    out += idx + b'\0' + b'\0\0\0\0'
    footer = meta_handle + idx_handle
    footer += b'\0' * (40 - len(footer)) + struct.pack('<Q', 0xdb4775248b80fb57)
    out += footer
    with open(prefix + '.index', 'wb') as f:
        f.write(bytes(out))
    with open(prefix + '.data-00000-of-00001', 'wb') as f:
        f.write(bytes(data))
