"""The pure-C11 consumer of include/hj.h (tests/c/abi_smoke.c): compiled with -std=c11 -pedantic -Werror
against the header and linked against libhj_b200.so, no Python or C++ on its path.  Without a GPU it
must report HJ_ERR_NO_DEVICE and exit 77; on a GPU it drives device -> buffers -> hj_execute_graph
(kernel + reduce + scan + compress passes, then the captured-graph relaunch path) -> to_host and checks
every value itself — the call sequence of bindings/rust/backend_cuda.rs."""
import importlib
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
hj = importlib.import_module("hephaestus-jit_b200")


def _exe():
    import __graft_entry__ as g
    return g.build_abi_smoke()


def test_header_is_valid_c11_and_library_links():
    exe = _exe()
    assert os.path.exists(exe)
    if hj.device_count() == 0:
        out = subprocess.run([exe], capture_output=True, text=True, timeout=60)
        assert out.returncode == 77, out.stdout + out.stderr
        assert "no CPU fallback" in out.stdout


@pytest.mark.gpu
def test_c_consumer_drives_the_backend():
    out = subprocess.run([_exe()], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "ok:" in out.stdout and "how = 2" in out.stdout, out.stdout
