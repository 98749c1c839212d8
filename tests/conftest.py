import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def built_lib():
    """libsilent_b200.so, (re)built in-tree when nvcc is available and the sources are newer."""
    from pysilent_b200 import build, _lib
    try:
        build.build_library()
    except Exception as exc:   # no nvcc on this machine: the prebuilt .so that travelled with the repo must exist
        if not os.path.exists(_lib.LIB_PATH):
            pytest.fail("libsilent_b200.so missing and cannot be built: %s" % exc)
    return _lib.lib()


@pytest.fixture(scope="session")
def c_oracle():
    from oracle import c_oracle as co
    co.build()
    return co


@pytest.fixture(scope="session")
def goldens():
    return {name: np.load(os.path.join(GOLDEN, name + ".npz")) for name in ("generators", "pyramid", "stack")}


@pytest.fixture(scope="session")
def default_filters():
    import pysilent_b200.constant_convolutions as cc
    return dict(rgc=cc.midget_rgc(2), rgby=cc.rgby_3(2), stripe=cc.rgb_2d_stripe_tensors(),
                blur=cc.blur_tensor(2, lengths=7), end=cc.rgb_2d_end_tensors())


def synthetic_frame(config, index, h, w, c=3):
    """BASELINE/SURVEY 8(d): uint8 uniform noise from RandomState(1000 * config + frame_index)."""
    return np.random.RandomState(1000 * config + index).randint(0, 256, size=(h, w, c)).astype(np.uint8)


def structured_frame(seed, h, w):
    """Parity-only frame with edges, a flat patch (regulator gain >> 1) and a black patch (0 * inf = NaN)."""
    rs = np.random.RandomState(seed)
    img = rs.randint(0, 256, size=(h, w, 3)).astype(np.uint8)
    yy, xx = np.mgrid[0:h, 0:w]
    smooth = (96 + 80 * np.sin(xx / 9.0) * np.cos(yy / 7.0)).astype(np.uint8)
    img[: h // 3] = smooth[: h // 3, :, None]
    img[(2 * h) // 5: (3 * h) // 5, (2 * w) // 5: (3 * w) // 5] = 0
    img[(7 * h) // 10: (9 * h) // 10, w // 10: (4 * w) // 10] = 40
    return img
