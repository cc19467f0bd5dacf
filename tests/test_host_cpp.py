"""The C++ host facade (bitar_b200/host): bitar's headers for Class_CUDA over the C-ABI, built against the
Arrow C++ of this image, and bitar_demo, the reference's demo_app flow (apps/demo_app.cc:487-693:
sync Compress / Decompress / memcmp / Recycle on device 0, then CompressAsync / DecompressAsync over every
queue pair with per-part memcmp)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "bitar_b200", "host")
DEMO = os.path.join(HOST, "bitar_demo")


def _build():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "bitar_b200", "csrc"), "-s"])
    subprocess.check_call(["make", "-C", HOST, "-s"])


def test_host_facade_builds_and_fails_loudly_without_a_device():
    _build()
    syms = subprocess.check_output(["nm", "-DC", os.path.join(HOST, "libbitar_host.so")], text=True)
    for name in ("bitar::CompressDevice<", "::Compress(", "::Decompress(", "::Recycle(", "::Initialize(",
                 "bitar::CompressDriver<", "::ListAvailableDeviceIds()", "::GetDevices(", "bitar::GetMemoryPool(",
                 "bitar::CudaConfiguration::to_c()", "bitar::internal::DistributeWorkers("):
        assert name in syms, name
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present: covered by the gpu test")
    r = subprocess.run([DEMO, "--bytes", "100000"], capture_output=True, text=True)
    assert r.returncode != 0 and "No compress device is available" in r.stderr      # src/driver.cc:184-187: no CPU fallback


@pytest.mark.gpu
@pytest.mark.parametrize("args", [["--bytes", str(48 << 20)], ["--bytes", str(20 << 20), "--device", "--seg", "65536"],
                                  ["--bytes", str(3 << 20), "--seg", "4096", "--qps", "3"],
                                  ["--bytes", str(40 << 20), "--sgl", "4"], ["--bytes", str(24 << 20), "--device", "--sgl", "8", "--seg", "32768"]])
def test_demo_app_flow(args):
    _build()
    r = subprocess.run([DEMO] + args, capture_output=True, text=True, timeout=600)
    print(r.stdout[-3000:], r.stderr[-2000:])
    assert r.returncode == 0 and "PASSED" in r.stdout and "MISMATCH" not in r.stdout
    if "--device" not in args and "--sgl" not in args:   # Decompress() stages heap-resident compressed buffers itself
        assert "pageable compressed input: OK" in r.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("fmt", ["parquet", "feather", "raw"])
def test_demo_app_content_mode(tmp_path, fmt):
    """The reference's content mode (apps/demo_app.cc:144-247): a Parquet / Feather file becomes Arrow IPC stream
    bytes in the pinned-host pool, is compressed and decompressed on the GPU, and must still deserialise."""
    import numpy as np
    import pyarrow as pa
    import pyarrow.feather as feather
    import pyarrow.parquet as pq
    _build()
    rng = np.random.default_rng(7)
    n = 400_000
    table = pa.table({
        "l_orderkey": np.cumsum(rng.integers(0, 4, n)).astype(np.int64),
        "l_returnflag": pa.array(rng.choice(["A", "N", "R"], n)).dictionary_encode(),
        "l_extendedprice": rng.integers(90000, 10500000, n) / 100.0,
        "l_comment": pa.array([f"row {i % 977} of the synthetic lineitem table" for i in range(n)]),
    })
    path = str(tmp_path / f"lineitem.{fmt}")
    if fmt == "parquet":
        pq.write_table(table, path)
    elif fmt == "feather":
        feather.write_feather(table, path, compression="uncompressed")
    else:
        with open(path, "wb") as f:
            f.write(table.column("l_orderkey").chunk(0).buffers()[1].to_pybytes())
    r = subprocess.run([DEMO, "--file", path, "--mode", "sync"], capture_output=True, text=True, timeout=600)
    print(r.stdout[-3000:], r.stderr[-2000:])
    assert r.returncode == 0 and "PASSED" in r.stdout and "MISMATCH" not in r.stdout
    if fmt != "raw":
        assert "deserialised table: OK" in r.stdout and f"{n} rows x 4 columns" in r.stdout


def test_framing_helpers_against_zlib():
    """bitar/framing.h (zlib / gzip wrappers around a chunk's raw stream, and unframing) against zlib itself."""
    subprocess.check_call(["make", "-C", HOST, "-s", "tests/framing_test"])
    r = subprocess.run([os.path.join(HOST, "tests", "framing_test")], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "framing: OK" in r.stdout, r.stderr


@pytest.mark.gpu
def test_device_memory_pool():
    """GetMemoryPool(CudaDevice): Allocate / Reallocate (device-side copy) / Free, statistics, tracker, and
    AllocateDeviceBuffer() on top of it (/root/reference/src/memory_pool.cc:125-188,321-350)."""
    _build()
    r = subprocess.run([os.path.join(HOST, "tests", "pool_test")], capture_output=True, text=True, timeout=300)
    print(r.stdout, r.stderr)
    assert r.returncode == 0 and "device pool: OK" in r.stdout


@pytest.mark.gpu
def test_demo_app_async_over_all_devices_in_one_process():
    """bitar's own multi-device model: ONE process, every visible device, CompressAsync / DecompressAsync over every
    (device, queue pair) (/root/reference/apps/demo_app.cc:556-607, src/driver.cc:100-158).  Needs two GPUs."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("one GPU visible: the single-process multi-device flow needs two")
    _build()
    r = subprocess.run([DEMO, "--bytes", str(512 << 20), "--mode", "async", "--qps", "2"], capture_output=True, text=True, timeout=900)
    print(r.stdout[-4000:], r.stderr[-2000:])
    out_dir = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out_dir, exist_ok=True)
    with open(os.path.join(out_dir, "demo_async_all_devices.txt"), "w") as f:
        f.write(r.stdout)
    assert r.returncode == 0 and "PASSED" in r.stdout and "MISMATCH" not in r.stdout
    assert f"{torch.cuda.device_count()} device(s)" in r.stdout
