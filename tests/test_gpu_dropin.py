"""GPU: the C++ drop-in.  tests/dropin_driver.cpp -- a caller shaped like mapOptmization.cpp -- is compiled against
include/Scancontext.h + libscgpu.so and its output is compared BYTE FOR BYTE with the same source compiled against the
reference's own Scancontext.h/.cpp (oracle/_ref/dropin_ref, built by oracle/Makefile where /root/reference exists):
loop ids, yaw bits, every "[Loop found]/[Not loop]" line the class prints, descriptors, keys, pairwise distances."""
import os
import subprocess

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("devices", [None, "0,0", "0,0,0,0,0,0,0,0"])
def test_same_caller_same_output(tmp_path, devices):
    ref_bin = os.path.join(ROOT, "oracle", "_ref", "dropin_ref")
    if not os.path.exists(ref_bin):
        pytest.skip("oracle/_ref/dropin_ref not built (needs /root/reference at build time)")
    from sc_lego_loam_b200 import build
    exe = str(tmp_path / "dropin_scgpu")
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    pkg = os.path.dirname(build.LIB_SCGPU)
    build.build_scangen()
    cmd = [cxx, "-std=c++14", "-O2", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "oracle", "stubs"),
           os.path.join(ROOT, "tests", "dropin_driver.cpp"), "-o", exe, build.LIB_SCGPU, build.LIB_SCANGEN, f"-Wl,-rpath,{pkg}"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    n = "150"
    want = subprocess.run([ref_bin, n], capture_output=True, text=True, timeout=300)
    env = dict(os.environ)
    if devices:      # the database sharded over a device list behind the same class (here: shards sharing the one GPU)
        env["SCGPU_DEVICES"] = devices
    got = subprocess.run([exe, n], capture_output=True, text=True, timeout=300, env=env)
    assert want.returncode == 0 and got.returncode == 0, got.stderr
    wl, gl = want.stdout.splitlines(), got.stdout.splitlines()
    assert len(wl) == len(gl)
    for i, (a, b) in enumerate(zip(wl, gl)):
        assert a == b, f"line {i}: reference {a!r} != scgpu {b!r}"
    assert any(l.startswith("[Loop found]") for l in wl) and any(l.startswith("[Not loop]") for l in wl)
