import sys, os
sys.path.insert(0, "/root/repo")
sys.path.insert(0, "/root/repo/tools")
import torch
from weaksuppointcloudseg_b200 import _lib as L
import time_knn as tk
for path in (0, 1):
    L.lib().wspc_set_knn_path(path)
    print("knn path", path)
    tk.bench(128, 4096, 3, 20, 0, coff=6)
    tk.bench(128, 4096, 6, 10, 1, coff=0)
