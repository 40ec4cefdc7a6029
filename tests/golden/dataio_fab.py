"""Fabricates tiny on-disk S3DIS / ShapeNet datasets in the reference's formats (seeded), shared by
tests/test_dataio_cpu.py and tests/golden/make_dataio_golden.py so both see the same bytes."""
import json
import os

import numpy as np

from weaksuppointcloudseg_b200 import _h5


def make_s3dis(root, n_point=64):
    rng = np.random.default_rng(77)
    os.makedirs(root, exist_ok=True)
    rooms = []
    files = []
    for fi, nblk in enumerate((7, 6)):
        data = rng.random((nblk, n_point, 9)).astype(np.float32)
        label = rng.integers(0, 13, (nblk, n_point)).astype(np.uint8)
        label[0, :] = 3                                          # a single-class block
        name = 'ply_data_all_%d.h5' % fi
        _h5.write(os.path.join(root, name), {'data': data, 'label': label}, chunk_rows=4)
        files.append('indoor3d_sem_seg_hdf5_data/' + name)
        for b in range(nblk):
            rooms.append('Area_%d_office_%d' % ((fi * 7 + b) % 6 + 1, b))
    with open(os.path.join(root, 'all_files.txt'), 'w') as fh:
        fh.write('\n'.join(files) + '\n')
    with open(os.path.join(root, 'room_filelist.txt'), 'w') as fh:
        fh.write('\n'.join(rooms) + '\n')
    return root


def make_s3dis_room(root, area='area5'):
    """A 2.6 m x 1.7 m room of 1500 points (x y z r g b label), one sparse corner (< 100 points -> dropped column)."""
    rng = np.random.default_rng(78)
    os.makedirs(os.path.join(root, 'meta'), exist_ok=True)
    n = 1500
    xyz = rng.random((n, 3)) * np.array([2.6, 1.7, 3.0])
    far = (xyz[:, 0] > 2.0) & (xyz[:, 1] > 1.0)
    keep = ~far | (rng.random(n) < 0.15)
    xyz = xyz[keep]
    xyz -= xyz.min(0)
    rgb = rng.integers(0, 256, (xyz.shape[0], 3)).astype(np.float64)
    lab = rng.integers(0, 13, (xyz.shape[0], 1)).astype(np.float64)
    room = np.concatenate([xyz, rgb, lab], 1)
    np.save(os.path.join(root, 'Area_5_office_1.npy'), room)
    np.savetxt(os.path.join(root, 'Area_5_office_2.txt'), room[:900], fmt='%.6f')
    with open(os.path.join(root, 'meta', '%s_data_label.txt' % area), 'w') as fh:
        fh.write('Area_5_office_1.npy\nArea_5_office_2.txt\n')
    return root


SHAPENET_CATS = [('Airplane', '02691156', 4), ('Bag', '02773838', 2), ('Cap', '02954340', 2)]


def make_shapenet(root, n_point=48):
    rng = np.random.default_rng(79)
    h5 = os.path.join(root, 'hdf5_data')
    os.makedirs(h5, exist_ok=True)
    oid2cpid, cpid2oid = [], {}
    for _name, cid, nparts in SHAPENET_CATS:
        for p in range(1, nparts + 1):
            cpid2oid['%s_%d' % (cid, p)] = len(oid2cpid)
            oid2cpid.append([cid, p])
    json.dump(oid2cpid, open(os.path.join(h5, 'overallid_to_catid_partid.json'), 'w'))
    json.dump(cpid2oid, open(os.path.join(h5, 'catid_partid_to_overallid.json'), 'w'))
    json.dump([[0.1 * i, 0.5, 0.5] for i in range(len(oid2cpid))], open(os.path.join(h5, 'part_color_mapping.json'), 'w'))
    with open(os.path.join(h5, 'all_object_categories.txt'), 'w') as fh:
        fh.write(''.join('%s\t%s\n' % (n, c) for n, c, _ in SHAPENET_CATS))
    starts = np.cumsum([0] + [c[2] for c in SHAPENET_CATS])

    def one_file(name, n):
        label = rng.integers(0, len(SHAPENET_CATS), (n, 1)).astype(np.uint8)
        data = rng.standard_normal((n, n_point, 3)).astype(np.float32)
        pid = np.stack([rng.integers(starts[c], starts[c + 1], n_point) for c in label[:, 0]]).astype(np.uint8)
        _h5.write(os.path.join(h5, name), {'data': data, 'label': label, 'pid': pid}, chunk_rows=3)

    for lst, names, sizes in (('train_hdf5_file_list.txt', ('ply_data_train0.h5', 'ply_data_train1.h5'), (6, 5)),
                              ('val_hdf5_file_list.txt', ('ply_data_val0.h5',), (8,))):
        for nm, sz in zip(names, sizes):
            one_file(nm, sz)
        with open(os.path.join(h5, lst), 'w') as fh:
            fh.write('\n'.join(names) + '\n')
    # test shapes: <cat>/points/<id>.pts, <cat>/expert_verified/points_label/<id>.seg
    lines = []
    for si, (_name, cid, nparts) in enumerate(SHAPENET_CATS[:2]):
        n = 30 + 7 * si
        pd = os.path.join(root, 'PartAnnotation', cid, 'points')
        sd = os.path.join(root, 'PartAnnotation', cid, 'expert_verified', 'points_label')
        os.makedirs(pd, exist_ok=True)
        os.makedirs(sd, exist_ok=True)
        np.savetxt(os.path.join(pd, 's%d.pts' % si), rng.standard_normal((n, 3)) * 0.3 + 0.1, fmt='%.5f')
        np.savetxt(os.path.join(sd, 's%d.seg' % si), rng.integers(1, nparts + 1, n), fmt='%d')
        lines.append('%s/points/s%d.pts %s/expert_verified/points_label/s%d.seg %s' % (cid, si, cid, si, cid))
    with open(os.path.join(root, 'testing_ply_file_list.txt'), 'w') as fh:
        fh.write('\n'.join(lines) + '\n')
    return root


# ---- call scripts: the same sequence of loader calls is run on the reference's classes (golden generator) and on ours ----

def _put(out, key, tup):
    for j, v in enumerate(tup):
        if v is not None and not isinstance(v, (bool, str)):
            out['%s.%d' % (key, j)] = np.asarray(v)


def trace_s3dis_io(cls, root):
    out = {}
    ld = cls(root, 13, batchsize=4, NUM_POINT=64)
    ld.LoadS3DIS_AllData()
    ld.CreateDataSplit(5)
    out['train_idx'], out['test_idx'] = np.asarray(ld.train_data_idxs), np.asarray(ld.test_data_idxs)
    np.random.seed(5)
    for epoch in range(2):
        ld.Shuffle_TrainSet()
        for step in range(20):
            o = ld.NextBatch_TrainSet() if epoch == 0 else ld.NextBatch_TrainSet_v1()
            if not o[0]:
                break
            _put(out, 'train%d.%d' % (epoch, step), o)
        out['train%d.steps' % epoch] = np.asarray(step)
    ld.ResetLoader_TestSet()
    for step in range(20):
        o = ld.NextBatch_TestSet_v1(batchsize=2)
        if not o[0]:
            break
        _put(out, 'test.%d' % step, o)
    out['test.steps'] = np.asarray(step)
    for step in range(20):
        o = ld.NextBatch_TrainValSet()
        if not o[0]:
            break
        _put(out, 'all.%d' % step, o)
    out['all.steps'] = np.asarray(step)
    return out


def trace_s3dis_test(obj):
    """`obj`: an S3DIS_Test whose ROOM_PATH_LIST is set."""
    out = {}
    np.random.seed(6)
    obj.ResetTestRoom()
    r = 0
    while True:
        data, label, path = obj.LoadNextTestRoomData_v1()
        if data is None:
            break
        out['room%d.data' % r], out['room%d.label' % r] = np.asarray(data), np.asarray(label)
        r += 1
    out['rooms'] = np.asarray(r)
    return out


def trace_shapenet(cls, root):
    out = {}
    ld = cls(root, batchsize=4)
    ld.LoadTrainValFiles()
    ld.LoadTestFiles()
    out['NUM_PART_CATS'] = np.asarray(ld.NUM_PART_CATS)
    out['oids_bag'] = np.asarray(ld.object2setofoid['02773838'])
    np.random.seed(7)
    for epoch in range(2):
        ld.Shuffle_TrainSet()
        for step in range(20):
            o = ld.NextBatch_TrainSet()
            if not o[0]:
                break
            _put(out, 'train%d.%d' % (epoch, step), o)
        out['train%d.steps' % epoch] = np.asarray(step)
    for rep in range(2):
        for step in range(20):
            o = ld.NextBatch_ValSet()
            if not o[0]:
                break
            _put(out, 'val%d.%d' % (rep, step), o)
        out['val%d.steps' % rep] = np.asarray(step)
    for step in range(5):
        o = ld.NextSamp_TestSet()
        if not o[0]:
            break
        _put(out, 'te.%d' % step, o)
    out['te.steps'] = np.asarray(step)
    return out
