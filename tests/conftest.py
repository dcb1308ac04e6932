import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


class ParityReport:
    """Collects every engine-vs-oracle comparison so that one GPU run yields the full picture
    (written to gpurun_out/parity_report.json) even when an assertion fails early."""

    def __init__(self):
        self.rows = []

    def add(self, test, name, err, tol, extra=None):
        self.rows.append(dict(test=test, name=name, err=float(err), tol=float(tol), ok=bool(err <= tol), extra=extra))

    def dump(self):
        out = os.path.join(ROOT, "gpurun_out")
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, "parity_report.json"), "w") as fh:
            json.dump(self.rows, fh, indent=1)


_REPORT = ParityReport()


@pytest.fixture(scope="session")
def report():
    yield _REPORT
    _REPORT.dump()
