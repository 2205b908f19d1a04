"""TensorFlow tensor-bundle checkpoints without TensorFlow (tf_raft_b200/checkpoint.py).

TensorFlow is not available here, so these tests pin (a) the format description against itself (writer -> reader,
checksums, table framing), (b) published check values of the primitives (crc32c, masking, varints), and (c) the
key mapping on the reference's attribute tree."""
import os
import struct

import numpy as np
import pytest

from oracle import weights
from tf_raft_b200 import checkpoint as ck


def test_crc32c_known_answers():
    # RFC 3720 appendix B.4 check values for CRC-32C
    assert ck.crc32c(b'123456789') == 0xE3069283
    assert ck.crc32c(bytes(32)) == 0x8A9136AA
    assert ck.crc32c(bytes([0xFF] * 32)) == 0x62A8AB43
    assert ck.crc32c(bytes(range(32))) == 0x46DD794E
    # LevelDB's mask is a rotate + add of 0xa282ead8; the CRC of an empty string is 0
    assert ck._mask(0) == 0xa282ead8


def test_varint_round_trip():
    for v in (0, 1, 127, 128, 300, 2 ** 31 - 1, 2 ** 40 + 5):
        enc = ck._put_varint(v)
        got, pos = ck._get_varint(enc + b'\xff', 0)
        assert got == v and pos == len(enc)
    assert ck._put_varint(300) == b'\xac\x02'           # protobuf documentation example


def test_write_read_round_trip(tmp_path):
    rng = np.random.default_rng(0)
    tensors = {
        'fnet/conv1/kernel/.ATTRIBUTES/VARIABLE_VALUE': rng.standard_normal((7, 7, 3, 64)).astype(np.float32),
        'fnet/conv1/bias/.ATTRIBUTES/VARIABLE_VALUE': rng.standard_normal(64).astype(np.float32),
        'save_counter/.ATTRIBUTES/VARIABLE_VALUE': np.array(3, dtype=np.int64),
        'misc/half': rng.standard_normal((2, 3)).astype(np.float16),
        'misc/flags': np.array([True, False, True]),
        'misc/empty': np.zeros((0, 4), dtype=np.float32),
    }
    prefix = str(tmp_path / 'ckpt' / 'model')
    ck.write_tf_checkpoint(prefix, tensors)
    assert os.path.exists(prefix + '.index') and os.path.exists(prefix + '.data-00000-of-00001')
    got = ck.read_tf_checkpoint(prefix, verify=True)
    assert set(got) == set(tensors)
    for k, v in tensors.items():
        assert got[k].dtype == v.dtype and got[k].shape == v.shape and np.array_equal(got[k], v), k
    # table framing: 48-byte footer ending in the LevelDB magic
    raw = open(prefix + '.index', 'rb').read()
    assert struct.unpack('<Q', raw[-8:])[0] == 0xdb4775248b80fb57


def test_corruption_is_detected(tmp_path):
    prefix = str(tmp_path / 'model')
    ck.write_tf_checkpoint(prefix, {'a/.ATTRIBUTES/VARIABLE_VALUE': np.arange(16, dtype=np.float32)})
    data = bytearray(open(prefix + '.data-00000-of-00001', 'rb').read())
    data[5] ^= 0x40
    open(prefix + '.data-00000-of-00001', 'wb').write(bytes(data))
    ck.read_tf_checkpoint(prefix)                          # tensor checksums are opt-in ...
    with pytest.raises(ValueError, match='checksum'):
        ck.read_tf_checkpoint(prefix, verify=True)         # ... and catch the flipped bit
    idx = bytearray(open(prefix + '.index', 'rb').read())
    idx[3] ^= 0x01
    open(prefix + '.index', 'wb').write(bytes(idx))
    with pytest.raises(ValueError):
        ck.read_tf_checkpoint(prefix)                      # index blocks are always verified
    with pytest.raises(ValueError, match='magic'):
        open(prefix + '.index', 'wb').write(b'\0' * 64)
        ck.read_tf_checkpoint(prefix)


def _to_tf_key(name, with_model_prefix):
    """Inverse of tf_key_to_param for the reference's attribute tree: numeric path components are children of a keras
    Sequential, which object-graph checkpoints call layer_with_weights-N (N counts weight-owning children only)."""
    parts, out = name.split('.'), []
    for i, p in enumerate(parts):
        if p.isdigit():
            n = int(p)
            if parts[i - 1] == 'mask':
                n = {0: 0, 2: 1}[n]                        # [Conv2D, ReLU, Conv2D]: the ReLU owns no weights
            out.append(f'layer_with_weights-{n}')
        else:
            out.append(p)
    return ('model/' if with_model_prefix else '') + '/'.join(out) + '/.ATTRIBUTES/VARIABLE_VALUE'


@pytest.mark.parametrize('variant', ['raft', 'small'])
@pytest.mark.parametrize('with_model_prefix', [False, True])
def test_key_mapping_covers_the_reference_parameter_tree(tmp_path, variant, with_model_prefix):
    params = weights.init_params(variant, 7)
    tensors = {_to_tf_key(k, with_model_prefix): v for k, v in params.items()}
    assert len(tensors) == len(params)
    # things a real training checkpoint also holds and the loader must ignore
    tensors['optimizer/iter/.ATTRIBUTES/VARIABLE_VALUE'] = np.array(10, dtype=np.int64)
    k0 = _to_tf_key('fnet.conv1.kernel', with_model_prefix)
    tensors[k0[:-len('/.ATTRIBUTES/VARIABLE_VALUE')] + '/.OPTIMIZER_SLOT/optimizer/m/.ATTRIBUTES/VARIABLE_VALUE'] = \
        np.zeros((7, 7, 3, params['fnet.conv1.kernel'].shape[-1]), dtype=np.float32)
    tensors['save_counter/.ATTRIBUTES/VARIABLE_VALUE'] = np.array(1, dtype=np.int64)
    prefix = str(tmp_path / 'model')
    ck.write_tf_checkpoint(prefix, tensors)
    got = ck.load_tf_checkpoint(prefix)
    assert set(got) == set(params)
    for k, v in params.items():
        assert np.array_equal(got[k], np.asarray(v)), k
    # expect=: one error that names what is missing / unexpected / mis-shaped, instead of a KeyError later
    assert set(ck.load_tf_checkpoint(prefix, expect=params)) == set(params)
    broken = dict(params)
    gone = broken.pop('fnet.conv1.bias')
    broken['fnet.conv1.kernel'] = np.zeros((7, 7, 3, 1), dtype=np.float32)
    broken['cnet.not_in_the_checkpoint'] = gone
    with pytest.raises(ValueError, match='missing.*cnet.not_in_the_checkpoint.*unexpected.*fnet.conv1.bias.*shapes.*fnet.conv1.kernel'):
        ck.load_tf_checkpoint(prefix, expect=broken)


def test_key_mapping_examples():
    f = ck.tf_key_to_param
    s = '/.ATTRIBUTES/VARIABLE_VALUE'
    assert f('fnet/conv1/kernel' + s) == 'fnet.conv1.kernel'
    assert f('model/cnet/layer2/layer_with_weights-0/downsample/layer_with_weights-1/moving_variance' + s) == \
        'cnet.layer2.0.downsample.1.moving_variance'
    assert f('update_block/mask/layer_with_weights-1/kernel' + s) == 'update_block.mask.2.kernel'
    assert f('update_block/gru/convz1/bias' + s) == 'update_block.gru.convz1.bias'
    assert f('_CHECKPOINTABLE_OBJECT_GRAPH') is None
    assert f('optimizer/beta_1' + s) is None
    assert f('fnet/conv1/kernel/.OPTIMIZER_SLOT/optimizer/m' + s) is None


# ---- hand-assembled index bytes (format description only; none of checkpoint.py's writer helpers) -----------------------
def _crc32c_bitwise(data):
    """Independent CRC-32C: bit-at-a-time over the reflected Castagnoli polynomial (no table)."""
    crc = 0xFFFFFFFF
    for byte in data:
        crc ^= byte
        for _ in range(8):
            crc = (crc >> 1) ^ (0x82F63B78 if crc & 1 else 0)
    return crc ^ 0xFFFFFFFF


def _trailer(block):
    """LevelDB block trailer: compression type 0 + masked crc32c of (block || type), little-endian."""
    crc = _crc32c_bitwise(block + b'\x00')
    masked = (((crc >> 15) | (crc << 17)) + 0xa282ead8) & 0xFFFFFFFF
    return b'\x00' + masked.to_bytes(4, 'little')


def test_reader_on_hand_assembled_table(tmp_path):
    """An `.index` typed out byte by byte from the LevelDB table / tensor-bundle description: two data blocks, prefix
    compression inside a block (shared > 0), a block with TWO restart points, a multi-byte varint offset, a string entry
    the reader must skip, and the 48-byte footer.  Pins the reader against the format rather than against the writer."""
    f32 = np.array([1.0, -2.5, 3.25], dtype='<f4')
    i64 = np.array([[7, -1]], dtype='<i8')
    pad = bytes(200)                                        # pushes the second tensor to offset 212 (two-byte varint)
    data = f32.tobytes() + pad + i64.tobytes()
    tensor_crc = lambda raw: ((((_crc32c_bitwise(raw) >> 15) | (_crc32c_bitwise(raw) << 17)) + 0xa282ead8) & 0xFFFFFFFF)

    # BundleHeaderProto {num_shards: 1, version {producer: 1}}
    header = bytes([0x08, 0x01, 0x1a, 0x02, 0x08, 0x01])
    # BundleEntryProto {dtype: DT_FLOAT(1), shape {dim {size: 3}}, size: 12, crc32c: fixed32}   (offset 0 omitted)
    e_f32 = bytes([0x08, 0x01, 0x12, 0x04, 0x12, 0x02, 0x08, 0x03, 0x28, 0x0c, 0x35]) + tensor_crc(f32.tobytes()).to_bytes(4, 'little')
    # {dtype: DT_INT64(9), shape {dim {size: 1} dim {size: 2}}, offset: 212 = d4 01, size: 16, crc32c}
    e_i64 = bytes([0x08, 0x09, 0x12, 0x08, 0x12, 0x02, 0x08, 0x01, 0x12, 0x02, 0x08, 0x02, 0x20, 0xd4, 0x01, 0x28, 0x10, 0x35]) \
        + tensor_crc(i64.tobytes()).to_bytes(4, 'little')
    # {dtype: DT_STRING(7), shape {}, size: 5}: not a numeric tensor, skipped by the reader
    e_str = bytes([0x08, 0x07, 0x12, 0x00, 0x28, 0x05])

    def entry(shared, suffix, value):
        return bytes([shared, len(suffix), len(value)]) + suffix + value

    # data block 0: keys "", "ab/w", "ab/x" (shares "ab/" with its predecessor); one restart at offset 0
    b0 = entry(0, b'', header) + entry(0, b'ab/w', e_f32) + entry(3, b'x', e_str)
    b0 += (0).to_bytes(4, 'little') + (1).to_bytes(4, 'little')
    # data block 1: keys "cd/a", "cd/b" with a restart point at EACH (second entry therefore stores its key in full)
    first = entry(0, b'cd/a', e_str)
    b1 = first + entry(0, b'cd/b', e_i64)
    b1 += (0).to_bytes(4, 'little') + len(first).to_bytes(4, 'little') + (2).to_bytes(4, 'little')
    meta = (0).to_bytes(4, 'little') + (1).to_bytes(4, 'little')                 # empty metaindex block
    off0 = 0
    off1 = len(b0) + 5
    off_meta = off1 + len(b1) + 5
    off_index = off_meta + len(meta) + 5
    assert max(off1, off_meta, off_index, len(b0), len(b1)) < 128 * 2            # handles below: one- or two-byte varints

    def varint(v):
        return bytes([v]) if v < 128 else bytes([(v & 0x7f) | 0x80, v >> 7])

    # index block: separator key (>= last key of the block) -> BlockHandle (offset, size), restart interval 1
    i0 = entry(0, b'ab/x', varint(off0) + varint(len(b0)))
    i1 = entry(0, b'cd/b', varint(off1) + varint(len(b1)))
    index = i0 + i1 + (0).to_bytes(4, 'little') + len(i0).to_bytes(4, 'little') + (2).to_bytes(4, 'little')
    footer = varint(off_meta) + varint(len(meta)) + varint(off_index) + varint(len(index))
    footer += bytes(40 - len(footer)) + bytes([0x57, 0xfb, 0x80, 0x8b, 0x24, 0x75, 0x47, 0xdb])
    table = b0 + _trailer(b0) + b1 + _trailer(b1) + meta + _trailer(meta) + index + _trailer(index) + footer

    prefix = str(tmp_path / 'hand')
    open(prefix + '.index', 'wb').write(table)
    open(prefix + '.data-00000-of-00001', 'wb').write(data)
    got = ck.read_tf_checkpoint(prefix, verify=True)
    assert set(got) == {'ab/w', 'cd/b'}
    assert got['ab/w'].dtype == np.float32 and np.array_equal(got['ab/w'], f32)
    assert got['cd/b'].dtype == np.int64 and got['cd/b'].shape == (1, 2) and np.array_equal(got['cd/b'], i64)

    # a flipped bit inside a data block must trip the block trailer CRC; one in the tensor bytes the per-tensor CRC
    bad = bytearray(table)
    bad[off1 + 3] ^= 0x01
    open(prefix + '.index', 'wb').write(bytes(bad))
    with pytest.raises(ValueError, match='checksum'):
        ck.read_tf_checkpoint(prefix)
    open(prefix + '.index', 'wb').write(table)
    bad = bytearray(data)
    bad[1] ^= 0x10
    open(prefix + '.data-00000-of-00001', 'wb').write(bytes(bad))
    with pytest.raises(ValueError, match='checksum'):
        ck.read_tf_checkpoint(prefix, verify=True)


def test_writer_output_matches_hand_assembled_bytes(tmp_path):
    """The writer, on one small tensor, must produce exactly the bytes the format description gives."""
    arr = np.array([1.0, 2.0], dtype=np.float32)
    prefix = str(tmp_path / 'w')
    ck.write_tf_checkpoint(prefix, {'k': arr})
    raw = arr.astype('<f4').tobytes()
    crc = _crc32c_bitwise(raw)
    masked = (((crc >> 15) | (crc << 17)) + 0xa282ead8) & 0xFFFFFFFF
    header = bytes([0x08, 0x01, 0x1a, 0x02, 0x08, 0x01])
    e = bytes([0x08, 0x01, 0x12, 0x04, 0x12, 0x02, 0x08, 0x02, 0x28, 0x08, 0x35]) + masked.to_bytes(4, 'little')
    b0 = bytes([0, 0, len(header)]) + header + bytes([0, 1, len(e)]) + b'k' + e + (0).to_bytes(4, 'little') + (1).to_bytes(4, 'little')
    meta = (0).to_bytes(4, 'little') + (1).to_bytes(4, 'little')
    off_meta = len(b0) + 5
    off_index = off_meta + len(meta) + 5
    handle = bytes([0, len(b0)])
    index = bytes([0, 1, len(handle)]) + b'k' + handle + (0).to_bytes(4, 'little') + (1).to_bytes(4, 'little')
    footer = bytes([off_meta, len(meta), off_index, len(index)])
    footer += bytes(40 - len(footer)) + bytes([0x57, 0xfb, 0x80, 0x8b, 0x24, 0x75, 0x47, 0xdb])
    want = b0 + _trailer(b0) + meta + _trailer(meta) + index + _trailer(index) + footer
    assert open(prefix + '.index', 'rb').read() == want
    assert open(prefix + '.data-00000-of-00001', 'rb').read() == raw
