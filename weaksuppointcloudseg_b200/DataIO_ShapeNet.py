"""ShapeNet part-segmentation loader with the reference's interface (ShapeNet/DataIO_ShapeNet.py:9-344): the h5 train /
val sets (`data` (n,2048,3), `label` (n,1), `pid` (n,2048)) and the per-shape `.pts` / `.seg` test files.

Host-side only; feeds `ShapeNet_Trainer.TrainOneEpoch_Full / EvalOneEpoch_Full / Test`.  The 8-tuple every `Next*` call
returns is the contract of SURVEY App. F: (ok, data, label, seg, weak_seg_onehot, mb_size, file_idx, data_idx).  Internals
differ from the reference: one cursor routine serves the train and val sets, the class-presence indicator is vectorised, h5
files go through `_h5`.
"""
import json
import os

import numpy as np

from . import _h5
from .DataIO_S3DIS import _lines, weak_onehot

_NONE8 = (False, None, None, None, None, None, None, None)


class ShapeNetIO:

    def __init__(self, BASE_DIR='/vision01/pointnet/part_seg/', batchsize=24):
        self.BASE_DIR = BASE_DIR
        self.h5_base_path = os.path.join(BASE_DIR, 'hdf5_data')
        self.ply_data_dir = os.path.join(BASE_DIR, 'PartAnnotation')
        self.batchsize = batchsize

        def meta(name):
            return os.path.join(self.h5_base_path, name)

        self.color_map_file = meta('part_color_mapping.json')
        with open(self.color_map_file) as fh:
            self.color_map = json.load(fh)
        self.all_obj_cats_file = meta('all_object_categories.txt')
        pairs = [ln.split() for ln in _lines(self.all_obj_cats_file) if ln]
        self.all_obj_cats = [(p[0], p[1]) for p in pairs]
        self.objnames = [p[0] for p in pairs]                    # 'Airplane'
        self.objcats = [p[1] for p in pairs]                     # '02691156'
        self.on2oid = {c: i for i, c in enumerate(self.objcats)}
        with open(meta('overallid_to_catid_partid.json')) as fh:
            oid2cpid = json.load(fh)
        with open(meta('catid_partid_to_overallid.json')) as fh:
            self.cpid2oid = json.load(fh)
        self.NUM_CATEGORIES = 16
        self.NUM_PART_CATS = len(oid2cpid)
        self.object2setofoid = {}
        for oid, (objid, _pid) in enumerate(oid2cpid):
            self.object2setofoid.setdefault(objid, []).append(oid)

    # ---- train / val ---------------------------------------------------------------------------------------------------
    def _load_split(self, list_file):
        files = _lines(list_file)
        data, labels, seg, idx, n = [], [], [], [], 0
        for name in files:
            d, l, s, m, i = self.loadDataFile_with_seg(os.path.join(self.h5_base_path, name))
            data.append(d)
            labels.append(l)
            seg.append(s)
            idx.append(i + n)
            n += m
        return files, np.concatenate(data), np.concatenate(labels), np.concatenate(seg), np.concatenate(idx), n

    def LoadTrainValFiles(self):
        self.TRAINING_FILE_LIST = os.path.join(self.h5_base_path, 'train_hdf5_file_list.txt')
        self.VAL_FILE_LIST = os.path.join(self.h5_base_path, 'val_hdf5_file_list.txt')
        (self.train_file_list, self.train_data, self.train_labels, self.train_seg, self.train_data_idx,
         self.num_train) = self._load_split(self.TRAINING_FILE_LIST)
        (self.val_file_list, self.val_data, self.val_labels, self.val_seg, self.val_data_idx,
         self.num_val) = self._load_split(self.VAL_FILE_LIST)
        self.num_train_file = len(self.train_file_list)
        self.num_test_file = len(self.val_file_list)
        self.train_file_idx = np.arange(self.num_train_file)
        self.val_file_idx = np.arange(self.num_test_file)
        self.ResetLoader_TrainSet()
        self.ResetLoader_ValSet()

    def _next(self, split):
        """One cursor step (:163-186): full batches while more than a batch remains, then the tail, which also arms the
        end-of-set flag; the call after that resets the cursor and returns the all-None tuple."""
        ptr, end, n = split + '_samp_ptr', split + '_end_dataset', getattr(self, 'num_' + split)
        p = getattr(self, ptr)
        if getattr(self, end) or p >= n:
            setattr(self, ptr, 0)
            setattr(self, end, False)
            return _NONE8
        stop = min(p + self.batchsize, n)
        if p + self.batchsize >= n:
            setattr(self, end, True)
        setattr(self, ptr, stop)
        data_idx = getattr(self, split + '_data_idx')[p:stop].copy()
        data = getattr(self, split + '_data')[data_idx]
        label = getattr(self, split + '_labels')[data_idx]
        seg = getattr(self, split + '_seg')[data_idx]
        return (True, data, label, seg, weak_onehot(seg, self.NUM_PART_CATS), stop - p, np.zeros_like(data_idx), data_idx)

    def NextBatch_TrainSet(self, shuffle_flag=False):
        return self._next('train')

    def NextBatch_ValSet(self):
        return self._next('val')

    def Shuffle_TrainSet(self):
        np.random.shuffle(self.train_data_idx)

    def ResetLoader_TrainSet(self):
        self.train_samp_ptr = 0
        self.train_end_dataset = False

    def ResetLoader_ValSet(self):
        self.val_samp_ptr = 0
        self.val_end_dataset = False

    # ---- test ------------------------------------------------------------------------------------------------------------
    def LoadTestFiles(self):
        """`testing_ply_file_list.txt`: `<pts file> <seg file> <category id>` per line (:126-146)."""
        self.TEST_FILE_LIST = os.path.join(self.BASE_DIR, 'testing_ply_file_list.txt')
        rows = [ln.split() for ln in _lines(self.TEST_FILE_LIST) if ln]
        self.test_pts_files = [r[0] for r in rows]
        self.test_seg_files = [r[1] for r in rows]
        self.test_labels = [r[2] for r in rows]
        self.test_samp_num = len(rows)
        self.test_file_idx = np.arange(self.test_samp_num)
        self.ResetLoader_TestSet()

    def ResetLoader_TestSet(self):
        self.te_samp_ptr = 0

    def NextSamp_TestSet(self):
        """One shape with all its points, unit-sphere normalised: data (1,n,3), label [[cat]], seg (1,n) overall part ids."""
        if self.te_samp_ptr >= self.test_samp_num:
            self.ResetLoader_TestSet()
            return _NONE8
        i = self.te_samp_ptr
        cat = self.on2oid[self.test_labels[i]]
        pts, seg = self.load_pts_seg_files(os.path.join(self.ply_data_dir, self.test_pts_files[i]),
                                           os.path.join(self.ply_data_dir, self.test_seg_files[i]), self.objcats[cat])
        seg = seg[np.newaxis]
        self.te_samp_ptr += 1
        return (True, self.pc_normalize(pts)[np.newaxis], np.array([[cat]]), seg, weak_onehot(seg, self.NUM_PART_CATS), 1, 0, i)

    # ---- file formats --------------------------------------------------------------------------------------------------
    def loadDataFile_with_seg(self, filename):
        return self.load_h5_data_label_seg(filename)

    def load_h5_data_label_seg(self, h5_filename):
        d = _h5.read(h5_filename)
        n = d['data'].shape[0]
        return d['data'], d['label'], d['pid'], n, np.arange(n)

    def load_pts_seg_files(self, pts_file, seg_file, catid):
        """`.pts`: one `x y z` per line; `.seg`: one per-category part id per line, mapped to the overall part id."""
        pts = np.loadtxt(pts_file, dtype=np.float32, ndmin=2)
        part_ids = np.loadtxt(seg_file, dtype=np.int64, ndmin=1)
        lut = {}
        seg = np.array([lut.setdefault(x, self.cpid2oid[catid + '_' + str(x)]) for x in part_ids.tolist()])
        return pts, seg

    def pc_normalize(self, pc):
        pc = pc - np.mean(pc, axis=0)
        return pc / np.max(np.sqrt(np.sum(pc ** 2, axis=1)))


def LoadSampIndex(save_filepath):
    """`Dataset/ShapeNet/Preprocess/SampIndex_m-<m>.mat` as train_ShapeNet.py:91-96 uses it: the (n_shapes, n_labelled)
    matrix of labelled point indices, with file index 0 and data index i for row i (the single concatenated training array
    of ShapeNetIO.LoadTrainValFiles).  -> (file_idx_list, data_idx_list, pts_idx_list)"""
    import scipy.io as scio
    pts_idx_list = scio.loadmat(save_filepath)['pts_idx_list']
    return np.zeros(shape=[pts_idx_list.shape[0]]), np.arange(0, pts_idx_list.shape[0]), pts_idx_list
