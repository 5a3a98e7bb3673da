import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


@pytest.fixture(scope="session")
def lena_jpg() -> bytes:
    return (ROOT / "tests" / "golden" / "lena.jpg").read_bytes()


@pytest.fixture(scope="session")
def decoder():
    import libkpeg_b200 as K

    dec = K.Decoder(device=0)
    yield dec
    dec.close()
