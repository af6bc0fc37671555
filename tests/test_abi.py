"""CPU: the C-ABI library builds for sm_100a, loads, and exports every symbol include/scgpu.h declares.
No compute call is made here (there is no GPU in the build container and no CPU fallback in the library)."""
import ctypes
import os
import re
import subprocess

import pytest

from conftest import ROOT


def _declared():
    text = open(os.path.join(ROOT, "include", "scgpu.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(?:int|const char\*)\s+(scgpu_\w+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib_path():
    from sc_lego_loam_b200 import build
    build.build_scgpu()
    return build.LIB_SCGPU


def test_header_declares_reference_surface():
    names = _declared()
    for fn in ("scgpu_make_sc", "scgpu_ringkey", "scgpu_sectorkey", "scgpu_fast_align", "scgpu_dist_direct",
               "scgpu_distance", "scgpu_append_scan", "scgpu_detect"):
        assert fn in names


def test_library_exports_every_declared_symbol(lib_path):
    lib = ctypes.CDLL(lib_path)
    missing = [n for n in _declared() if not hasattr(lib, n)]
    assert not missing, missing


def test_python_binding_covers_header(lib_path):
    from sc_lego_loam_b200 import scgpu
    assert sorted(scgpu.EXPORTED_SYMBOLS) == _declared()
    scgpu.load_library()


def test_library_is_sm100a_only(lib_path):
    out = subprocess.run(["cuobjdump", "-lelf", lib_path], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, out


def test_no_device_means_error_not_fallback(lib_path):
    """Without a CUDA device scgpu_create must fail with SCGPU_E_NODEVICE (-3): there is no CPU path."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from sc_lego_loam_b200 import scgpu
    with pytest.raises(scgpu.ScgpuError) as e:
        scgpu.SCManager()
    assert "-3" in str(e.value) or "no CUDA device" in str(e.value)


def test_drop_in_header_compiles_against_stub_eigen_pcl(lib_path, tmp_path):
    """include/Scancontext.h (the C++ drop-in) compiles in a translation unit that uses it exactly the way
    mapOptmization.cpp does (member object, the two API calls) and links against libscgpu.so."""
    src = tmp_path / "caller.cpp"
    src.write_text('''
#include "Scancontext.h"
struct Node { SCManager scManager; };
int main() {
  Node n;
  pcl::PointCloud<SCPointType> cloud;
  if (false) {
    n.scManager.makeAndSaveScancontextAndKeys(cloud);
    auto r = n.scManager.detectLoopClosureID();
    (void)r.first; (void)r.second;
    Eigen::MatrixXd sc = n.scManager.makeScancontext(cloud);
    Eigen::MatrixXd rk = n.scManager.makeRingkeyFromScancontext(sc);
    Eigen::MatrixXd sk = n.scManager.makeSectorkeyFromScancontext(sc);
    std::pair<double,int> d = n.scManager.distanceBtnScanContext(sc, sc);
    (void)n.scManager.fastAlignUsingVkey(sk, sk); (void)n.scManager.distDirectSC(sc, sc); (void)d; (void)rk;
    n.scManager.setDownsampleLeaf(0.5f);   // extension: the caller's pcl::VoxelGrid moved in front of the build
  }
  return n.scManager.PC_NUM_RING == 20 && n.scManager.NUM_CANDIDATES_FROM_TREE == 10 ? 0 : 1;
}
''')
    exe = tmp_path / "caller"
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    cmd = [cxx, "-std=c++14", "-DSCGPU_LAZY_HANDLE", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "oracle", "stubs"),
           str(src), "-o", str(exe), lib_path, f"-Wl,-rpath,{os.path.dirname(lib_path)}"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert subprocess.run([str(exe)]).returncode == 0


def test_product_never_touches_the_oracle():
    """The oracle is test infrastructure: nothing in the product package or the public headers imports, links or
    executes anything under oracle/ (there is no CPU fallback to route through)."""
    bad = []
    for base in ("sc-lego-loam_b200", "sc_lego_loam_b200", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if not f.endswith((".py", ".cu", ".cuh", ".h", ".c", ".cpp")):
                    continue
                text = open(os.path.join(dirpath, f), errors="replace").read()
                if re.search(r"^\s*(from|import)\s+oracle\b", text, re.M) or re.search(r'#include\s+["<][^">]*oracle', text) \
                        or "libscoracle" in text or "libscref" in text or "libvoxoracle" in text:
                    bad.append(os.path.join(base, f))
    assert not bad, bad
