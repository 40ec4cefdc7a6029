"""Loaders: the same call scripts that tests/golden/make_dataio_golden.py ran on the REFERENCE's
DataIO_S3DIS / DataIO_ShapeNet classes are run on ours, on the same fabricated files and numpy seeds; every returned
array must be identical.  Plus the HDF5 subset round trip (`_h5`: parity unpinned against libhdf5, see its header)."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))

import dataio_fab as fab  # noqa: E402
from weaksuppointcloudseg_b200 import _h5  # noqa: E402
from weaksuppointcloudseg_b200.DataIO_S3DIS import S3DIS_IO, S3DIS_Test  # noqa: E402
from weaksuppointcloudseg_b200.DataIO_ShapeNet import ShapeNetIO  # noqa: E402


@pytest.fixture(scope="module")
def golden():
    return np.load(os.path.join(HERE, "golden", "ref_dataio.npz"))


def _same(golden, prefix, got):
    want = {k[len(prefix):]: golden[k] for k in golden.files if k.startswith(prefix)}
    assert sorted(want) == sorted(got), (sorted(set(want) ^ set(got)))
    for k, w in want.items():
        g = np.asarray(got[k])
        assert g.shape == w.shape, (k, g.shape, w.shape)
        if w.dtype.kind == 'f':
            np.testing.assert_allclose(g, w, rtol=0, atol=1e-12, err_msg=k)
        else:
            assert np.array_equal(g, w), k


def test_h5_round_trip_dtypes_and_ragged_chunks(tmp_path):
    rng = np.random.default_rng(0)
    d = {'data': rng.standard_normal((150, 64, 9)).astype(np.float32),
         'label': rng.integers(0, 13, (150, 64)).astype(np.uint8),
         'pid': rng.integers(-5, 50, (150, 64)).astype(np.int64),
         'w': rng.standard_normal(7),
         'one': np.arange(3, dtype=np.int32).reshape(3, 1)}
    p = str(tmp_path / 'a.h5')
    _h5.write(p, d, chunk_rows=64)                               # 150 rows -> 2 full chunks + a ragged one
    back = _h5.read(p)
    assert sorted(back) == sorted(d)
    for k in d:
        assert back[k].dtype == d[k].dtype and np.array_equal(back[k], d[k]), k
    with _h5.File(p) as f:                                       # the h5py idiom of the reference's loaders
        assert f['data'][:].shape == (150, 64, 9) and f['label'].dtype == np.uint8


def test_h5_rejects_what_it_does_not_restate(tmp_path):
    p = str(tmp_path / 'bad.h5')
    open(p, 'wb').write(b'not hdf5 at all' * 10)
    with pytest.raises(ValueError):
        _h5.read(p)
    blob = bytearray(open(_mk(tmp_path), 'rb').read())
    blob[8] = 2                                                  # superblock v2 (libver='latest')
    open(p, 'wb').write(bytes(blob))
    with pytest.raises(NotImplementedError):
        _h5.read(p)


def _mk(tmp_path):
    p = str(tmp_path / 'ok.h5')
    _h5.write(p, {'x': np.arange(10, dtype=np.float32)})
    return p


def test_s3dis_io_matches_reference_loader(tmp_path, golden):
    got = fab.trace_s3dis_io(S3DIS_IO, fab.make_s3dis(str(tmp_path / 's3dis')))
    _same(golden, 's3dis_io/', got)
    assert got['train0.0.3'].shape == (4, 13) and int(got['train0.steps']) == 3   # 11 train blocks = 4 + 4 + 3


def test_s3dis_room_to_blocks_matches_reference(tmp_path, golden):
    root = fab.make_s3dis_room(str(tmp_path / 'rooms'))
    t = S3DIS_Test('area5', NUM_POINT=128, data_path=root)
    assert [os.path.basename(p) for p in t.ROOM_PATH_LIST] == ['Area_5_office_1.npy', 'Area_5_office_2.txt']
    got = fab.trace_s3dis_test(t)
    _same(golden, 's3dis_test/', got)
    d = got['room0.data']
    assert d.shape[1:] == (128, 9) and d[..., 6:9].max() <= 1.0 and abs(d[..., 0]).max() <= 0.5 + 1e-9


def test_shapenet_io_matches_reference_loader(tmp_path, golden):
    got = fab.trace_shapenet(ShapeNetIO, fab.make_shapenet(str(tmp_path / 'shapenet')))
    _same(golden, 'shapenet/', got)
    assert int(got['NUM_PART_CATS']) == 8 and int(got['te.steps']) == 2


def test_missing_files_are_loud(tmp_path):
    with pytest.raises(FileNotFoundError):
        S3DIS_IO(str(tmp_path / 'nowhere'))
    with pytest.raises(FileNotFoundError):
        ShapeNetIO(str(tmp_path / 'nowhere'))


def test_samp_index_mat_layouts(tmp_path):
    """SampIndex_m-*.mat: dense matrix for m > 0, (1, n) object array of (1, n_i) rows for m == 0 (the layouts of the
    reference's Dataset/*/Preprocess fixtures), unpacked as train_S3DIS.py:92-101 / train_ShapeNet.py:91-96 do."""
    import scipy.io as scio
    from weaksuppointcloudseg_b200 import DataIO_S3DIS, DataIO_ShapeNet
    from weaksuppointcloudseg_b200.S3DIS_DGCNN_trainer import S3DIS_Trainer
    rng = np.random.default_rng(2)
    dense = np.stack([rng.choice(64, 4, replace=False) for _ in range(9)]).astype(np.int64)
    scio.savemat(str(tmp_path / 'SampIndex_m-0.010.mat'), {'pts_idx_list': dense})
    got = DataIO_S3DIS.LoadSampIndex(str(tmp_path / 'SampIndex_m-0.010.mat'), 0.01)
    assert np.array_equal(got, dense)
    ragged = np.empty((1, 9), dtype=object)
    rows = [rng.choice(64, int(n), replace=False).astype(np.int64) for n in rng.integers(1, 6, 9)]
    for i, r in enumerate(rows):
        ragged[0, i] = r[None, :]
    scio.savemat(str(tmp_path / 'SampIndex_m-0.000.mat'), {'pts_idx_list': ragged})
    got0 = DataIO_S3DIS.LoadSampIndex(str(tmp_path / 'SampIndex_m-0.000.mat'), 0)
    assert len(got0) == 9 and all(np.array_equal(a, b) for a, b in zip(got0, rows))
    mask = S3DIS_Trainer._mask_from_idx(got0, np.array([3, 8]), 2, 64)          # what the epoch loop builds from it
    assert mask.sum() == len(rows[3]) + len(rows[8]) and mask[1, rows[8]].all()
    f, d, p = DataIO_ShapeNet.LoadSampIndex(str(tmp_path / 'SampIndex_m-0.010.mat'))
    assert f.shape == (9,) and not f.any() and np.array_equal(d, np.arange(9)) and np.array_equal(p, dense)


def test_h5_multidimensional_chunks_shuffle_and_two_level_btree(tmp_path):
    """The shapes h5py's auto-chunking produces for the reference's files: chunks over every axis with ragged edges, the
    byte-shuffle filter in front of deflate, and more chunks than one B-tree node holds (a level-1 node over leaves)."""
    rng = np.random.default_rng(1)
    d = {'data': rng.standard_normal((37, 130, 9)).astype(np.float32),
         'label': rng.integers(0, 13, (37, 130)).astype(np.uint8),
         'pid': rng.integers(0, 50, (300,)).astype(np.int32)}
    p = str(tmp_path / 'b.h5')
    _h5.write(p, d, chunk_shape={3: (8, 50, 4), 2: (5, 64), 1: (7,)}, shuffle=True, node_entries=6)
    back = _h5.read(p)                                           # data: 5 x 3 x 3 = 45 chunks -> 8 leaves under one node
    for k in d:
        assert back[k].dtype == d[k].dtype and np.array_equal(back[k], d[k]), k


def test_h5_round_trip_property(tmp_path):
    """Randomised shapes, dtypes and chunk shapes (ragged edges, more chunks than a node holds) read back bit-identically."""
    from hypothesis import given, settings, strategies as st, HealthCheck

    dtypes = [np.float32, np.float64, np.uint8, np.int8, np.int16, np.int32, np.int64, np.uint16]

    @settings(max_examples=40, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])
    @given(st.integers(0, 2 ** 31 - 1), st.integers(1, 3), st.integers(0, len(dtypes) - 1), st.booleans(), st.integers(2, 9))
    def run(seed, rank, dt_i, shuffle, node_entries):
        rng = np.random.default_rng(seed)
        shape = tuple(int(x) for x in rng.integers(1, 23, rank))
        chunk = tuple(int(rng.integers(1, s + 3)) for s in shape)
        dt = np.dtype(dtypes[dt_i])
        a = (rng.standard_normal(shape) * 100).astype(dt) if dt.kind == 'f' else \
            rng.integers(np.iinfo(dt).min, np.iinfo(dt).max, shape, dtype=dt, endpoint=True)
        p = str(tmp_path / ('p%d.h5' % (seed % 7)))
        _h5.write(p, {'x': a, 'second/name'.replace('/', '_'): a[:1]}, chunk_shape={rank: chunk}, shuffle=shuffle,
                  node_entries=node_entries)
        back = _h5.read(p)
        assert back['x'].dtype == dt and back['x'].shape == shape and np.array_equal(back['x'], a)
        assert np.array_equal(back['second_name'], a[:1])

    run()


@pytest.mark.parametrize("stride,random_sample", [(1.0, False), (0.5, False), (1.0, True), (0.7, True)])
def test_room2blocks_binning_equals_whole_room_masks(stride, random_sample):
    """The binned block membership against the reference's formulation (a mask over every point of the room per block,
    DataIO_S3DIS.py:398-404), for overlapping strides and randomly placed blocks (negative corners included)."""
    rng = np.random.default_rng(12)
    n = 6000
    data = np.concatenate([rng.random((n, 3)) * np.array([3.3, 2.4, 3.0]), rng.random((n, 3))], 1)
    data[:50, 0] = np.round(data[:50, 0])                          # points exactly on block borders (inclusive both sides)
    label = rng.integers(0, 13, n).astype(np.uint8)
    t = S3DIS_Test.__new__(S3DIS_Test)
    np.random.seed(3)
    got_d, got_l = t.room2blocks(data, label, 64, block_size=1.0, stride=stride, random_sample=random_sample, sample_num=None)

    # the reference's formulation, restated: same corner lists, same random draws
    np.random.seed(3)
    limit = data.max(0)[0:3]
    if not random_sample:
        nx = int(np.ceil((limit[0] - 1.0) / stride)) + 1
        ny = int(np.ceil((limit[1] - 1.0) / stride)) + 1
        corners = [(i * stride, j * stride) for i in range(nx) for j in range(ny)]
    else:
        nb = int(np.ceil(limit[0])) * int(np.ceil(limit[1]))
        corners = [(np.random.uniform(-1.0, limit[0]), np.random.uniform(-1.0, limit[1])) for _ in range(nb)]
    want_d, want_l = [], []
    for xb, yb in corners:
        cond = (data[:, 0] <= xb + 1.0) & (data[:, 0] >= xb) & (data[:, 1] <= yb + 1.0) & (data[:, 1] >= yb)
        if cond.sum() < 100:
            continue
        idx = np.flatnonzero(cond)
        pick = idx[S3DIS_Test._sample_indices(idx.size, 64)]
        want_d.append(data[pick])
        want_l.append(label[pick])
    assert len(want_d) > 3 and got_d.shape == (len(want_d), 64, 6)
    assert np.array_equal(got_d, np.stack(want_d)) and np.array_equal(got_l, np.stack(want_l))
