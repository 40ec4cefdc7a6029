"""S3DIS loaders with the reference's interface (S3DIS/DataIO_S3DIS.py): `S3DIS_IO` (the pre-blocked h5 training set,
:6-254) and `S3DIS_Test` (room -> 1 m x 1 m blocks of NUM_POINT points with the 9-channel normalisation, :257-432).

Host-side only (numpy): these feed `S3DIS_Trainer.TrainOneEpoch_Full / EvalOneEpoch_Full / Test`, which move each mini-batch
to the device through pinned staging buffers.  Differences from the reference are internal: the per-cloud python loops that
build `weak_seg_onehot` and the per-block normalisation are vectorised, h5 files are read through `_h5` (h5py when present,
else the restated subset), and a missing file raises instead of printing.  The numpy random stream is consumed by exactly
the calls the reference makes (`np.random.shuffle` of the train indices, one `np.random.choice` per block), so a seeded run
visits the same samples.
"""
import os

import numpy as np

from . import _h5


def _lines(path):
    with open(path) as fh:
        return [ln.rstrip() for ln in fh]


def weak_onehot(seg, num_parts):
    """(B, num_parts) indicator of the classes present in each cloud (DataIO_S3DIS.py:117-120)."""
    seg = np.asarray(seg)
    B = seg.shape[0]
    out = np.zeros([B, num_parts])
    if B:
        out[np.arange(B)[:, None], seg.reshape(B, -1).astype(np.int64)] = 1
    return out


class S3DIS_IO:
    CategoryName = None

    def __init__(self, h5filepath='./', numParts=13, batchsize=24, NUM_POINT=4096):
        self.data_base_path = h5filepath
        self.numParts = numParts
        self.ALL_FILES = _lines(os.path.join(h5filepath, 'all_files.txt'))
        self.room_filelist = _lines(os.path.join(h5filepath, 'room_filelist.txt'))
        self.batchsize = batchsize
        self.NUM_POINT = NUM_POINT
        self.NUM_PART_CATS = 13
        self.NUM_CATEGORIES = self.NUM_PART_CATS

    def load_h5(self, h5_filename):
        d = _h5.read(h5_filename)
        return d['data'], d['label']

    def loadDataFile(self, filename):
        return self.load_h5(filename)

    def LoadS3DIS_AllData(self):
        """Concatenate every file of all_files.txt; the list holds `<dir>/<file>` and only the file name is used (:53)."""
        data, label = [], []
        for name in self.ALL_FILES:
            d, l = self.loadDataFile(os.path.join(self.data_base_path, name.split('/')[1]))
            data.append(d)
            label.append(l)
        self.data_batches = np.concatenate(data, 0)
        self.label_batches = np.concatenate(label, 0)
        if len(self.room_filelist) != self.data_batches.shape[0]:
            raise ValueError("room_filelist.txt names %d blocks, the h5 files hold %d"
                             % (len(self.room_filelist), self.data_batches.shape[0]))

    def CreateDataSplit(self, test_area):
        """Blocks whose room name contains `Area_<test_area>` are the test split (:60-83)."""
        tag = 'Area_' + str(test_area)
        is_test = np.array([tag in r for r in self.room_filelist], dtype=bool)
        self.all_data_idxs = np.arange(len(self.room_filelist))
        self.train_data_idxs = self.all_data_idxs[~is_test]
        self.test_data_idxs = self.all_data_idxs[is_test]
        self.train_samp_ptr = self.test_samp_ptr = self.all_samp_ptr = 0

    def ResetLoader_TrainSet(self):
        self.train_samp_ptr = 0
        self.shuffled_train_data_idxs = self.train_data_idxs.copy()

    def Shuffle_TrainSet(self):
        self.ResetLoader_TrainSet()
        np.random.shuffle(self.shuffled_train_data_idxs)

    def ResetLoader_TestSet(self):
        self.test_samp_ptr = 0

    # One cursor step shared by every NextBatch_* flavour: a full batch, then the short tail, then None (:94-107).
    def _advance(self, ptr_name, idxs, batchsize):
        ptr, n = getattr(self, ptr_name), idxs.shape[0]
        if ptr >= n:
            return None
        sel = idxs[ptr:min(ptr + batchsize, n)]
        # a full batch advances by its own size, the tail by the loader's batch size (as the reference does)
        setattr(self, ptr_name, ptr + (batchsize if ptr + batchsize < n else self.batchsize))
        return sel

    def _collect(self, data_idx):
        data = self.data_batches[data_idx].copy()
        seg = self.label_batches[data_idx].copy()
        return data, seg, weak_onehot(seg, self.numParts), data_idx.shape[0]

    def NextBatch_TrainSet(self):
        """-> (ok, data (B,N,9), seg (B,N), weak_seg_onehot (B,13), mb_size).  data: 0:3 xyz (xy block-centred), 3:6 rgb/255,
        6:9 xyz normalised by the room extent."""
        sel = self._advance('train_samp_ptr', self.shuffled_train_data_idxs, self.batchsize)
        if sel is None:
            return False, None, None, None, None
        return (True,) + self._collect(sel)

    def NextBatch_TrainSet_v1(self):
        sel = self._advance('train_samp_ptr', self.shuffled_train_data_idxs, self.batchsize)
        if sel is None:
            return False, None, None, None, None, None
        return (True,) + self._collect(sel) + (sel,)

    def NextBatch_TrainValSet(self):
        sel = self._advance('all_samp_ptr', self.all_data_idxs, self.batchsize)
        if sel is None:
            return False, None, None, None, None, None
        return (True,) + self._collect(sel) + (sel,)

    def NextBatch_TestSet(self, batchsize=None):
        sel = self._advance('test_samp_ptr', self.test_data_idxs, self.batchsize if batchsize is None else batchsize)
        if sel is None:
            return False, None, None, None, None
        return (True,) + self._collect(sel)

    def NextBatch_TestSet_v1(self, batchsize=None):
        sel = self._advance('test_samp_ptr', self.test_data_idxs, self.batchsize if batchsize is None else batchsize)
        if sel is None:
            return False, None, None, None, None, None
        return (True,) + self._collect(sel) + (sel,)


class S3DIS_Test:
    """Whole-room test loader: each room file (`.npy` / `.txt`, rows `x y z r g b label`, min corner at the origin) is cut into
    a fixed grid of 1 m x 1 m columns; columns with fewer than 100 points are dropped; each is resampled to NUM_POINT points."""

    def __init__(self, te_area, NUM_POINT=4096, data_path=None):
        self.te_area = te_area
        self.NUM_POINT = NUM_POINT
        # The reference joins '<repo>/Dataset/S3DIS/' with '/meta/<area>_data_label.txt' (:264-266); os.path.join drops the
        # first part when the second is absolute, so it only works from '/'.  Here `data_path` names the dataset directory.
        root = data_path if data_path is not None else os.path.join(os.getcwd(), 'Dataset', 'S3DIS')
        listing = os.path.join(root, 'meta', '{}_data_label.txt'.format(te_area))
        self.ROOM_PATH_LIST = [os.path.join(root, ln) for ln in _lines(listing)]
        self.ResetTestRoom()

    def ResetTestRoom(self):
        self.te_room_ptr = 0

    def LoadNextTestRoomData(self):
        data, label, _ = self.LoadNextTestRoomData_v1()
        return data, label

    def LoadNextTestRoomData_v1(self):
        if self.te_room_ptr >= len(self.ROOM_PATH_LIST):
            return None, None, None
        room_path = self.ROOM_PATH_LIST[self.te_room_ptr]
        data, label = self.room2blocks_wrapper_normalized(room_path, self.NUM_POINT)
        self.te_room_ptr += 1
        return data, label, room_path

    def room2blocks_wrapper_normalized(self, data_label_filename, num_point, block_size=1.0, stride=1.0,
                                       random_sample=False, sample_num=None, sample_aug=1):
        if data_label_filename.endswith('txt'):
            data_label = np.loadtxt(data_label_filename)
        elif data_label_filename.endswith('npy'):
            data_label = np.load(data_label_filename)
        else:
            raise ValueError('unknown room file type: ' + data_label_filename)
        return self.room2blocks_plus_normalized(data_label, num_point, block_size, stride, random_sample, sample_num,
                                                sample_aug)

    def room2blocks_plus_normalized(self, data_label, num_point, block_size, stride, random_sample, sample_num, sample_aug):
        """-> (K, num_point, 9) blocks and (K, num_point) uint8 labels: 0:3 xyz with xy centred on the block, 3:6 rgb/255,
        6:9 xyz / room maximum (:323-349).  Unlike the reference the caller's array is not modified in place."""
        data = np.array(data_label[:, 0:6], dtype=np.float64)
        data[:, 3:6] /= 255.0
        label = data_label[:, -1].astype(np.uint8)
        room_max = data[:, 0:3].max(0)
        blocks, labels = self.room2blocks(data, label, num_point, block_size, stride, random_sample, sample_num, sample_aug)
        out = np.empty((blocks.shape[0], num_point, 9))
        out[:, :, 6:9] = blocks[:, :, 0:3] / room_max
        out[:, :, 0:6] = blocks
        out[:, :, 0:2] -= blocks[:, :, 0:2].min(1, keepdims=True) + block_size / 2
        return out, labels

    def room2blocks(self, data, label, num_point, block_size=1.0, stride=1.0, random_sample=False, sample_num=None,
                    sample_aug=1):
        """data (n, 6) xyz in metres + rgb in [0,1], label (n,) -> (K, num_point, 6), (K, num_point) (:353-420)."""
        assert stride <= block_size
        limit = data.max(0)[0:3]
        if not random_sample:
            nx = int(np.ceil((limit[0] - block_size) / stride)) + 1
            ny = int(np.ceil((limit[1] - block_size) / stride)) + 1
            xbeg = np.repeat(np.arange(nx) * stride, ny)
            ybeg = np.tile(np.arange(ny) * stride, nx)
        else:
            nx = int(np.ceil(limit[0] / block_size))
            ny = int(np.ceil(limit[1] / block_size))
            if sample_num is None:
                sample_num = nx * ny * sample_aug
            corners = [(np.random.uniform(-block_size, limit[0]), np.random.uniform(-block_size, limit[1]))
                       for _ in range(sample_num)]
            xbeg = np.array([c[0] for c in corners])
            ybeg = np.array([c[1] for c in corners])
        x, y = data[:, 0], data[:, 1]
        block_data, block_label = [], []
        # Points are binned once on the stride grid; a block then tests only the bins it can touch instead of the whole
        # room (the reference masks all points per block).  Same inclusive bounds, same ascending point order inside a block.
        bx = np.floor(x / stride).astype(np.int64)
        by = np.floor(y / stride).astype(np.int64)
        ox, oy = int(bx.min()), int(by.min())
        ny_bins = int(by.max()) - oy + 1
        key = (bx - ox) * ny_bins + (by - oy)
        order = np.argsort(key, kind='stable')
        starts = np.searchsorted(key[order], np.arange((int(bx.max()) - ox + 1) * ny_bins + 1))
        reach = int(np.ceil(block_size / stride))
        for xb, yb in zip(xbeg, ybeg):
            i0, j0 = int(np.floor(xb / stride)) - ox, int(np.floor(yb / stride)) - oy
            cand = []
            for i in range(max(i0 - 1, 0), min(i0 + reach + 1, int(bx.max()) - ox) + 1):   # one bin of slack for rounding
                lo_j, hi_j = max(j0 - 1, 0), min(j0 + reach + 1, ny_bins - 1)
                if lo_j <= hi_j:
                    cand.append(order[starts[i * ny_bins + lo_j]:starts[i * ny_bins + hi_j + 1]])
            cand = np.concatenate(cand) if cand else np.zeros(0, np.int64)
            cx, cy = x[cand], y[cand]
            inside = np.sort(cand[(cx <= xb + block_size) & (cx >= xb) & (cy <= yb + block_size) & (cy >= yb)])
            if inside.size < 100:
                continue
            pick = inside[self._sample_indices(inside.size, num_point)]
            block_data.append(data[pick])
            block_label.append(label[pick])
        return np.stack(block_data, 0), np.stack(block_label, 0)

    @staticmethod
    def _sample_indices(n, num_sample):
        """Indices that keep / subsample with replacement / pad by random duplicates (:427-441)."""
        if n == num_sample:
            return np.arange(n)
        if n > num_sample:
            return np.random.choice(n, num_sample)
        return np.concatenate([np.arange(n), np.random.choice(n, num_sample - n)])

    def sample_data(self, data, num_sample):
        idx = self._sample_indices(data.shape[0], num_sample)
        return data[idx, ...], idx

    def sample_data_label(self, data, label, num_sample):
        new_data, idx = self.sample_data(data, num_sample)
        return new_data, label[idx]


def LoadSampIndex(save_filepath, m=None):
    """The labelled-point lists `Dataset/S3DIS/Preprocess/SampIndex_m-<m>.mat` exactly as train_S3DIS.py:92-101 unpacks them:
    a dense (n_blocks, n_labelled) int matrix for m > 0; for m == 0 (one point per class present in the block) the .mat
    holds a (1, n_blocks) object array of (1, n_i) rows, returned as a list of 1-D arrays.  Indexable by the loader's
    `data_idx`, which is what `TrainOneEpoch(_Full)` does."""
    import scipy.io as scio
    pts = scio.loadmat(save_filepath)['pts_idx_list']
    if pts.dtype == object or (m is not None and m == 0):
        return [np.asarray(pts[0, b_i][0]).reshape(-1) for b_i in range(pts.shape[1])]
    return pts
