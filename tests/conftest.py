import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle.rq_oracle import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def product():
    """The product library through the C ABI.  Built by __graft_entry__.build(); missing = hard error."""
    import cases
    return cases.rt.RTCore()


@pytest.fixture(scope="session", params=["gpu_builder=lbvh", "gpu_builder=ploc", "gpu_builder=sah"])
def gpu_device(product, request):
    """Every GPU test runs against all three binary-tree front ends of the builder (radix tree / PLOC / binned-SAH treelets)."""
    return product.new_device(request.param)


@pytest.fixture(scope="session")
def reflib():
    """The real reference library, when it was built in this checkout (oracle/_ref)."""
    import cases
    from oracle.rq_oracle import REF_LIB
    if not os.path.exists(REF_LIB):
        pytest.skip("oracle/_ref/libembree3_ref.so not built here")
    return cases.rt.RTCore(REF_LIB)
