import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


# A kernel that faults (e.g. a new tcgen05 kernel on its first hardware run) leaves a STICKY CUDA error: the
# tests after it fail fast, the summary is printed -- and then torch's teardown of the broken context can abort the
# interpreter, replacing pytest's exit status with SIGABRT.  If (and only if) the context is broken at the very end, leave
# with pytest's own status without running those destructors.  A healthy run never takes this path.
def pytest_sessionfinish(session, exitstatus):
    session.config._apla_exitstatus = int(exitstatus)


def pytest_unconfigure(config):
    try:
        import torch
        if not (torch.cuda.is_available() and torch.cuda.is_initialized()):
            return
        torch.cuda.synchronize()
    except Exception as e:                                  # noqa: BLE001
        sys.stdout.write(f"\n[conftest] CUDA context unusable at exit ({type(e).__name__}); leaving with pytest's "
                         f"status {getattr(config, '_apla_exitstatus', 1)}\n")
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(getattr(config, "_apla_exitstatus", 1))
