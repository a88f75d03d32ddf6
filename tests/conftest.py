import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'oracle'), os.path.join(ROOT, 'tests')):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')
    config.addinivalue_line('markers', 'reference: needs the read-only reference mount (/root/reference); skipped elsewhere')
    config.addinivalue_line('markers', 'slow: long soak run (minutes of GPU time); deselect with -m "gpu and not slow"')


def pytest_collection_modifyitems(config, items):
    import ref_harness

    if not ref_harness.available():
        skip = pytest.mark.skip(reason='reference mount not present (expected on the GPU box)')
        for item in items:
            if 'reference' in item.keywords:
                item.add_marker(skip)
