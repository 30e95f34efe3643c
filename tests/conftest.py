# -*- coding: utf-8 -*-
import os
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))


def pytest_configure(config):
    config.addinivalue_line(
        "markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _build_native():
    """ Make sure the native libraries exist (cheap no-op when up to date). """
    from fractalshades_b200 import build
    build.build_orbit()
    import oracle_lib
    oracle_lib.build()
    yield


@pytest.fixture(autouse=True)
def _reference_point_is_image_centre():
    """ The fixtures were generated with the image centre as reference point
    (settings.no_newton = True on both sides); cases and
    tests of the nucleus search switch it off themselves. """
    from fractalshades_b200 import settings
    old = settings.no_newton
    settings.no_newton = True
    yield
    settings.no_newton = old
