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
