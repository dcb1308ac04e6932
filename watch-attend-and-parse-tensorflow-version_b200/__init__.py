"""B200-native Watch-Attend-and-Parse hot path (sm_100a CUDA behind a C ABI).

The directory name follows the reference repository; import it as `wap_b200` (see wap_b200.py at the
repository root) or via importlib.import_module("watch-attend-and-parse-tensorflow-version_b200").
"""
from . import _lib  # noqa: F401
from .data import PackedBatch, Prefetcher, dataIterator, load_dict, pack_batch, prepare_data  # noqa: F401


def __getattr__(name):          # engine needs torch + the built extension: import lazily, fail loudly
    if name in ("engine", "WAPEngine", "ParamStore", "Watcher_train", "Attender", "Parser", "WAP"):
        import importlib
        eng = importlib.import_module(__name__ + ".engine")
        return eng if name == "engine" else getattr(eng, name)
    raise AttributeError(name)
