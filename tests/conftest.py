import os
import sys
import time

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


SESSION_T0 = time.time()


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


# ---- host-memory watchdog ------------------------------------------------------------------------------------------------------
# A test that runs away with host memory must kill ITSELF, not the machine: a box that goes down under the test suite looks like a
# lost GPU to whoever drives it.  A daemon thread polls this process's resident set and aborts it (exit code 86, message on
# stderr) once it passes half of the machine's RAM or comes within 16 GiB of all of it.  (Round 2 lost a test box to the oracle's
# saxpy-form masked product on RMAT-24 -- ~1e11 unmasked entries; the oracle now uses the dot form, and this is the backstop.)
def _start_memory_watchdog():
    import threading
    import time as _time
    try:
        import psutil
    except Exception:
        return
    total = psutil.virtual_memory().total
    limit = min(total // 2, max(total - (16 << 30), total // 4))
    me = psutil.Process()

    def watch():
        while True:
            try:
                if me.memory_info().rss > limit:
                    sys.stderr.write(f"\n[conftest] resident set passed {limit >> 30} GiB: aborting the test process to protect the host\n")
                    sys.stderr.flush()
                    os._exit(86)
            except Exception:
                return
            _time.sleep(0.25)

    threading.Thread(target=watch, daemon=True, name="rss-watchdog").start()


_start_memory_watchdog()
