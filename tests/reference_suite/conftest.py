"""Everything below this directory is the reference's own suite run against the device-backed
`oxli` module: mark it `gpu` from the outside so the files stay byte-identical."""
import os

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def pytest_collection_modifyitems(config, items):
    for item in items:
        if str(item.fspath).startswith(HERE + os.sep):
            item.add_marker(pytest.mark.gpu)
