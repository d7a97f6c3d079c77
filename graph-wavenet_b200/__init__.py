"""gwnet_b200 -- B200-native forward/backward hot path of Graph WaveNet (sklin93/Graph-WaveNet).

The directory name carries a hyphen, so import it either as the reference does -- put this
directory on ``sys.path`` and ``from model import *`` (engine.py:2) -- or as the package
``graph_wavenet_b200`` through ``__graft_entry__.load_package()``.

Modules
  native   ctypes binding of include/gwnet_b200.h (CUDA library only; no CPU fallback)
  runtime  marshalling for the whole-network plan
  model    drop-in for the reference ``model.py``: nconv, linear, gcn, gwnet
  engine   mirror of ``engine.trainer`` train/eval
"""
__version__ = "0.1.0"
