"""Runs the REFERENCE'S OWN loaders (S3DIS/DataIO_S3DIS.py, ShapeNet/DataIO_ShapeNet.py, imported unmodified from
/root/reference) on the fabricated datasets of dataio_fab.py and stores what they return:
    python tests/golden/make_dataio_golden.py  ->  tests/golden/ref_dataio.npz
h5py is absent from this image, so `import h5py` resolves to a two-line shim over weaksuppointcloudseg_b200._h5 (the
fixture therefore pins the loaders' cursor / split / sampling / normalisation logic, not libhdf5).  Only this generator
reads /root/reference."""
import os
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("WSPC_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
for sub in ("S3DIS", "ShapeNet"):
    sys.path.insert(1, os.path.join(REF, sub))

from weaksuppointcloudseg_b200 import _h5  # noqa: E402

sys.modules['h5py'] = types.SimpleNamespace(File=_h5.File)

import DataIO_S3DIS as ref_s3dis  # noqa: E402
import DataIO_ShapeNet as ref_shapenet  # noqa: E402
import dataio_fab as fab  # noqa: E402

assert ref_s3dis.__file__.startswith(REF) and ref_shapenet.__file__.startswith(REF)

out = {}
with tempfile.TemporaryDirectory() as tmp:
    for k, v in fab.trace_s3dis_io(ref_s3dis.S3DIS_IO, fab.make_s3dis(os.path.join(tmp, 's3dis'))).items():
        out['s3dis_io/' + k] = v
    room_root = fab.make_s3dis_room(os.path.join(tmp, 'rooms'))
    t = ref_s3dis.S3DIS_Test.__new__(ref_s3dis.S3DIS_Test)      # its __init__ builds an unusable absolute path (:264-266)
    t.te_area, t.NUM_POINT = 'area5', 128
    t.ROOM_PATH_LIST = [os.path.join(room_root, n) for n in ('Area_5_office_1.npy', 'Area_5_office_2.txt')]
    for k, v in fab.trace_s3dis_test(t).items():
        out['s3dis_test/' + k] = v
    for k, v in fab.trace_shapenet(ref_shapenet.ShapeNetIO, fab.make_shapenet(os.path.join(tmp, 'shapenet'))).items():
        out['shapenet/' + k] = v
np.savez_compressed(os.path.join(HERE, 'ref_dataio.npz'), **out)
print('wrote', len(out), 'arrays')
