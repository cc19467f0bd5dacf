"""The second device class behind DeviceManager::Create: CompressDevice<Class_ZLIB>, the
host-zlib software path as a backend of its own (bitar/device_zlib.h, libbitar_host_zlib.so) -- and the guarantee that the
CUDA class can never reach it: it is a separate library that nothing on the CUDA path links or loads."""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ZDIR = os.path.join(ROOT, "bitar_b200", "host_zlib")


def test_zlib_class_round_trip_and_contract():
    subprocess.check_call(["make", "-C", ZDIR, "-s"])
    r = subprocess.run([os.path.join(ZDIR, "zlib_class_test")], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "zlib class: OK" in r.stdout


def _needed(path):
    out = subprocess.run(["readelf", "-d", path], capture_output=True, text=True).stdout
    return re.findall(r"\(NEEDED\)\s+Shared library: \[([^\]]+)\]", out)


def test_cuda_path_never_links_the_zlib_class():
    subprocess.check_call(["make", "-C", ZDIR, "-s"])
    assert not any("bitar_cuda" in n or "bitar_host.so" in n for n in _needed(os.path.join(ZDIR, "libbitar_host_zlib.so")))
    for lib in (os.path.join(ROOT, "bitar_b200", "csrc", "libbitar_cuda.so"), os.path.join(ROOT, "bitar_b200", "host", "libbitar_host.so")):
        if os.path.exists(lib):
            needed = _needed(lib)
            assert not any("host_zlib" in n for n in needed), (lib, needed)
            assert not any(n.startswith("libz.") for n in needed), (lib, needed)     # no zlib anywhere on the CUDA path
    # the Python package and the bench's product arm never name the library either
    for dirpath, _, files in os.walk(os.path.join(ROOT, "bitar_b200")):
        if os.path.basename(dirpath) == "host_zlib":
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cc")) and "device_zlib" not in f:
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "host_zlib" not in text and "device_zlib" not in text, os.path.join(dirpath, f)
