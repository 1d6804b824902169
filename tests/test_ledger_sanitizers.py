"""the C++ ledger runtime under ASan+UBSan and under TSan.  A stand-alone driver
(csrc/ledger/ledger_selftest.cpp) runs 4 full protocol rounds with 20 concurrent client threads and
3 reader threads; any sanitizer report makes it exit non-zero."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "bflc_demo_b200", "csrc")


@pytest.mark.parametrize("san", ["address,undefined", "thread"])
def test_ledger_selftest_under_sanitizer(san, tmp_path):
    if shutil.which("g++") is None:
        pytest.skip("no g++")
    exe = str(tmp_path / "ledger_selftest")
    build = subprocess.run(
        ["g++", "-std=c++17", "-O1", "-g", f"-fsanitize={san}", "-fno-omit-frame-pointer",
         "-I" + os.path.join(SRC, "include"), os.path.join(SRC, "ledger", "ledger.cpp"),
         os.path.join(SRC, "ledger", "ledger_selftest.cpp"), "-lpthread", "-o", exe],
        capture_output=True, text=True, timeout=600)
    if build.returncode != 0 and "sanitize" in build.stderr and "cannot find" in build.stderr:
        pytest.skip("sanitizer runtime not installed")
    assert build.returncode == 0, build.stderr[-2000:]
    env = dict(os.environ, TSAN_OPTIONS="halt_on_error=1", ASAN_OPTIONS="detect_leaks=1",
               UBSAN_OPTIONS="halt_on_error=1:print_stacktrace=1")
    run = subprocess.run([exe], capture_output=True, text=True, timeout=600, env=env)
    assert run.returncode == 0, (run.stdout + run.stderr)[-3000:]
    assert "ledger_selftest OK" in run.stdout
