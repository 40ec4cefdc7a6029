"""TensorFlow V2 checkpoint bundle: writer/reader round trip of the restated format and the mapping onto the
flat variable store.  PARITY unpinned against TensorFlow itself (not installable here) -- see tf_checkpoint.py."""
import numpy as np
import pytest

from weaksuppointcloudseg_b200 import tf_checkpoint as tfc


def _graph_like_vars(rng):
    v = {}
    for i in range(1, 8):                                        # > 32 entries -> several data blocks, shared key prefixes
        cin, cout = 6 * i, 16 * i
        s = 'adj_conv%d' % i
        v[s + '/weights'] = rng.standard_normal((1, 1, cin, cout)).astype(np.float32)
        v[s + '/biases'] = rng.standard_normal(cout).astype(np.float32)
        for n in ('beta', 'gamma', 'pop_mean', 'pop_var'):
            v[s + '/bn/' + n] = rng.standard_normal(cout).astype(np.float32)
        v[s + '/weights/Adam'] = rng.standard_normal((1, 1, cin, cout)).astype(np.float32)
        v[s + '/weights/Adam_1'] = rng.random((1, 1, cin, cout)).astype(np.float32)
    v['Variable'] = np.asarray(1234, np.int64)                   # global step: a scalar
    v['beta1_power'] = np.asarray(0.9 ** 1234, np.float32)
    v['seg/conv3/weights'] = rng.standard_normal((1, 1, 256, 13)).astype(np.float32)
    return v


def test_bundle_round_trip(tmp_path):
    rng = np.random.default_rng(0)
    v = _graph_like_vars(rng)
    prefix = str(tmp_path / 'Checkpoint_epoch-best')
    tfc.write(prefix, v, entries_per_block=7)
    assert tfc.exists(prefix)
    idx = tfc.index(prefix)
    assert sorted(idx) == sorted(v) and idx['Variable']['shape'] == () and idx['adj_conv2/weights']['shape'] == (1, 1, 12, 32)
    back = tfc.read(prefix)
    assert sorted(back) == sorted(v)
    for k in v:
        assert back[k].dtype == v[k].dtype and back[k].shape == v[k].shape and np.array_equal(back[k], v[k]), k
    some = tfc.read(prefix, names={'Variable', 'seg/conv3/weights'})
    assert sorted(some) == ['Variable', 'seg/conv3/weights']


def test_bundle_rejects_foreign_and_damaged_files(tmp_path):
    prefix = str(tmp_path / 'ck')
    open(prefix + '.index', 'wb').write(b'\0' * 100)
    open(prefix + '.data-00000-of-00001', 'wb').write(b'')
    with pytest.raises(ValueError):
        tfc.read(prefix)
    tfc.write(prefix, {'a': np.arange(10, dtype=np.float32)})
    open(prefix + '.data-00000-of-00001', 'wb').write(b'\0' * 8)   # truncated shard
    with pytest.raises(ValueError):
        tfc.read(prefix)


def test_mapping_onto_the_flat_store():
    rng = np.random.default_rng(1)
    v = _graph_like_vars(rng)
    trainable = ['adj_conv1/weights', 'adj_conv1/biases', 'adj_conv1/bn/beta', 'adj_conv1/bn/gamma']
    state = ['adj_conv1/bn/pop_mean', 'adj_conv1/bn/pop_var']
    shapes = {'adj_conv1/weights': (6, 16), 'adj_conv1/biases': (16,), 'adj_conv1/bn/beta': (16,),
              'adj_conv1/bn/gamma': (16,), 'adj_conv1/bn/pop_mean': (16,), 'adj_conv1/bn/pop_var': (16,)}
    with pytest.raises(KeyError, match="optimiser state"):        # some variables of this checkpoint carry no Adam slots:
        tfc.to_store_blob(v, trainable, state, shapes)            # a silent restart of Adam / the schedules is refused ...
    with pytest.warns(UserWarning, match="optimiser state"):      # ... unless the caller asks for the weights only
        blob = tfc.to_store_blob(v, trainable, state, shapes, allow_missing_optimizer_state=True)
    assert blob['adj_conv1/weights'].shape == (6, 16)             # (1,1,Cin,Cout) kernel -> (Cin,Cout)
    assert np.array_equal(blob['adj_conv1/weights'], v['adj_conv1/weights'][0, 0]) and int(blob['Variable']) == 1234
    n = 6 * 16 + 3 * 16
    assert blob['__adam_m'].shape == (n,) and np.array_equal(blob['__adam_m'][:96], v['adj_conv1/weights/Adam'].reshape(-1))
    assert not blob['__adam_m'][96:].any()                        # variables without slots in the checkpoint start at zero
    with pytest.raises(ValueError):
        tfc.to_store_blob(v, trainable, state, dict(shapes, **{'adj_conv1/biases': (15,)}), allow_missing_optimizer_state=True)
    with pytest.raises(KeyError):
        tfc.to_store_blob(v, trainable + ['nope/weights'], state, dict(shapes, **{'nope/weights': (1,)}),
                          allow_missing_optimizer_state=True)


def test_restore_checkpoint_reads_a_tf_bundle(tmp_path):
    """S3DIS_Trainer.RestoreCheckPoint on `<prefix>.index/.data-*` (what the reference's Saver leaves behind): every variable
    of the S3DIS graph, the global step and the Adam slots land in the flat store (run on a CPU-resident store)."""
    import types
    import torch
    from weaksuppointcloudseg_b200.S3DIS_DGCNN_trainer import S3DIS_Trainer, xavier_params
    from weaksuppointcloudseg_b200.engine_s3dis import LAYERS
    from weaksuppointcloudseg_b200.runtime import VariableStore

    rng = np.random.default_rng(5)
    src = VariableStore(xavier_params(LAYERS, 21), 'cpu')
    src.state.copy_(torch.from_numpy(rng.random(src.state.numel()).astype(np.float32)))
    tf_vars = {}
    for k, a in src.export().items():
        tf_vars[k] = a.reshape((1, 1) + a.shape) if k.endswith('/weights') and a.ndim == 2 else a     # conv2d kernels
    for k in src.trainable_names:
        tf_vars[k + '/Adam'] = rng.standard_normal(tf_vars[k].shape).astype(np.float32)
        tf_vars[k + '/Adam_1'] = rng.random(tf_vars[k].shape).astype(np.float32)
    tf_vars['Variable'] = np.asarray(777, np.int64)
    tf_vars['beta1_power'] = np.asarray(0.5, np.float32)
    prefix = str(tmp_path / 'Checkpoint_epoch-best')
    tfc.write(prefix, tf_vars)

    tr = S3DIS_Trainer(5, device='cpu', seed=0)
    tr.engine = types.SimpleNamespace(vs=VariableStore(xavier_params(LAYERS, 99), 'cpu'))
    assert not torch.equal(tr.engine.vs.theta, src.theta)
    tr.RestoreCheckPoint(prefix)
    dst = tr.engine.vs
    assert torch.equal(dst.theta, src.theta) and torch.equal(dst.state, src.state) and dst.step == 777 == tr.batch
    for k in src.trainable_names:
        o, shape = dst._toffs[k]
        n = int(np.prod(shape))
        assert np.array_equal(dst.adam_m[o:o + n].numpy(), tf_vars[k + '/Adam'].reshape(-1)), k
        assert np.array_equal(dst.adam_v[o:o + n].numpy(), tf_vars[k + '/Adam_1'].reshape(-1)), k


def test_bundle_round_trip_property(tmp_path):
    """Randomised variable sets: names with shared prefixes across block boundaries, scalars, several dtypes."""
    from hypothesis import given, settings, strategies as st, HealthCheck

    dtypes = [np.float32, np.float64, np.int32, np.int64, np.uint8, np.float16]

    @settings(max_examples=30, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])
    @given(st.integers(0, 2 ** 31 - 1), st.integers(1, 60), st.integers(1, 20))
    def run(seed, n_vars, per_block):
        rng = np.random.default_rng(seed)
        v = {}
        for i in range(n_vars):
            scope = 'layer%d' % rng.integers(0, 6) + '/' * int(rng.integers(0, 2)) + 'bn' * int(rng.integers(0, 2))
            name = '%s/v%d' % (scope, i)
            rank = int(rng.integers(0, 4))
            shape = tuple(int(x) for x in rng.integers(1, 6, rank))
            dt = np.dtype(dtypes[int(rng.integers(0, len(dtypes)))])
            v[name] = (rng.standard_normal(shape) * 10).astype(dt)
        prefix = str(tmp_path / ('c%d' % (seed % 5)))
        tfc.write(prefix, v, entries_per_block=per_block)
        back = tfc.read(prefix)
        assert sorted(back) == sorted(v)
        for k in v:
            assert back[k].dtype == v[k].dtype and back[k].shape == v[k].shape and np.array_equal(back[k], v[k]), k

    run()
