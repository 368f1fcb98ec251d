"""CPU tests of the C++ host mirror that do not need a GPU: the MCout sink, and the loud failure
of MCPar::run / VLFunc when no CUDA device exists (no CPU fallback)."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "mcpar_b200", "host")


def test_mcout_mirror(tmp_path):
    exe = str(tmp_path / "mcout_check")
    subprocess.check_call(["/usr/bin/g++", "-std=c++11", "-O1", "-I", HOST, "-o", exe,
                           os.path.join(ROOT, "tests", "host_mcout_check.cc"), os.path.join(HOST, "mcout.cc")])
    assert subprocess.check_output([exe]).decode().strip() == "ok"


def test_driver_fails_loudly_without_gpu(tmp_path, mcgpu_lib):
    from mcpar_b200 import engine
    if engine.device_count() > 0:
        import pytest
        pytest.skip("a GPU is present")
    exe = os.path.join(ROOT, "mcpar_b200", "bin", "mcpar-rosen1")
    if not os.path.exists(exe):
        from mcpar_b200 import build
        build.build_host()
    r = subprocess.run([exe, "5"], cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode != 0 and "no usable CUDA device" in r.stderr
    assert r.stdout.splitlines()[0] == "nsamp = 5"          # the banner precedes engine creation, as in the reference


def test_ngpu_shape_rule_is_checked_before_any_device_work(tmp_path, mcgpu_lib):
    """MCPar::ngpu needs mpisiz % ngpu == 0 and 32-chain blocks per engine; the rule is checked on the
    host, so the refusal is the same with or without a GPU."""
    exe = os.path.join(ROOT, "mcpar_b200", "bin", "mcpar-rosen1")
    if not os.path.exists(exe):
        from mcpar_b200 import build
        build.build_host()
    r = subprocess.run([exe, "5", "--ranks=3", "--ngpu=2"], cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode == 2 and "ngpu" in r.stderr
    r = subprocess.run([exe, "5", "--ranks=4", "--ngpu=2"], cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode == 2 and "multiple of 32" in r.stderr        # 2 ranks x 4 chains per engine
