"""CPU-side checks (no GPU): the C-ABI library loads and exports every declared symbol, and the host
instantiation of the bit-exact atan2f restatement equals the host libm (the reference's atan2f)."""
import ctypes
import re
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(pkg):
    lib = pkg.load_library()
    header = open(os.path.join(os.path.dirname(pkg.HERE), "include", "lmot.h")).read()
    declared = sorted(set(re.findall(r"\b(lmot_[a-z0-9_]+)\s*\(", header)))
    assert declared, "no declarations parsed"
    for name in declared:
        assert hasattr(lib, name), f"liblmot.so does not export {name}"
    assert set(pkg.ABI_SYMBOLS) == set(declared)


def test_default_params_match_reference_constants(pkg):
    p = pkg.default_params()
    assert (p.r_min, p.r_max) == (np.float32(3.4), 120.0)
    assert (p.t_hmin, p.t_hmax, p.t_hdiff, p.h_sensor) == (-2.0, np.float32(-0.4), np.float32(0.4), 2.0)
    assert p.ground_tolerance == 0.25 and p.roi_m == 50.0
    assert (p.ram_points, p.l_slope_dist, p.l_num_points, p.min_cluster_points) == (80, 1, 5, 30)
    assert p.rule_filter == pkg.RULE_INTENDED


def test_atan2f_restatement_equals_host_libm(pkg):
    libm = ctypes.CDLL("libm.so.6")
    libm.atan2f.restype = ctypes.c_float
    libm.atan2f.argtypes = [ctypes.c_float, ctypes.c_float]
    rng = np.random.default_rng(0)
    y = rng.uniform(-120, 120, 200000).astype(np.float32)
    x = rng.uniform(-120, 120, 200000).astype(np.float32)
    x[::7] *= np.float32(1e-3); y[::11] *= np.float32(1e-4); x[::1013] = 0; y[::1019] = 0; x[::2027] = 1.0
    got = pkg.selftest_atan2f(y, x)
    want = np.array([libm.atan2f(float(a), float(b)) for a, b in zip(y[:20000], x[:20000])], np.float32)
    assert np.array_equal(got[:20000].view(np.uint32), want.view(np.uint32))
    # the rest against numpy's float32 arctan2 (which calls the same libm atan2f on this platform)
    np_at = np.arctan2(y, x)
    if np.array_equal(np_at[:20000].view(np.uint32), want.view(np.uint32)):
        assert np.array_equal(got.view(np.uint32), np_at.view(np.uint32))


def test_no_cuda_device_is_a_loud_failure(pkg):
    """Without a GPU lmot_create must fail (LMOT_ERR_CUDA), never fall back to a CPU path."""
    import torch
    if torch.cuda.is_available():
        return
    try:
        pkg.Lmot(device=0)
    except pkg.LmotError as e:
        assert e.status == pkg.ERR_CUDA
    else:
        raise AssertionError("lmot_create succeeded without a CUDA device")


def test_ros_codecs_cpu():
    """ros/include/lmot_ros_codec.hpp (PointCloud2 layout / zero-copy test, trackbox pack + unpack, box edges) against plain structs."""
    import subprocess
    exe = os.path.join(ROOT, "tests", "cpp", "ros_codec_test")
    subprocess.run(["g++", "-std=c++14", "-O1", "-Wall", os.path.join(ROOT, "tests", "cpp", "ros_codec_test.cpp"), "-o", exe], check=True)
    out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout
    assert "ros codec ok" in out


def test_ros_node_shells_compile():
    """The three node shells (ros/src) need ROS to link; here they are type-checked against API-shaped stand-ins of the ROS headers
    (tests/cpp/ros_stubs) and the real include/lmot.h + ros/include/lmot_ros_codec.hpp."""
    import subprocess
    for node in ("ground", "cluster", "tracking"):
        r = subprocess.run(["g++", "-std=c++14", "-fsyntax-only", "-w", "-I", os.path.join(ROOT, "tests", "cpp", "ros_stubs"), "-I",
                            os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "ros", "include"),
                            os.path.join(ROOT, "ros", "src", node + "_node.cpp")], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr


def test_frame_loop_driver_builds_and_fails_loudly_without_a_gpu(pkg, tmp_path):
    """host/frame_loop.cpp (the C++ submit / collect loop bench.py times `e2e` with) builds against include/lmot.h and the in-tree
    library; without a CUDA device it must stop with an error (pinned allocation / lmot_create fail), never fall back."""
    import subprocess
    import torch
    src = os.path.join(pkg.HERE, "host", "frame_loop.cpp")
    exe = str(tmp_path / "frame_loop")
    subprocess.run(["g++", "-std=c++14", "-O1", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), src, "-o", exe, "-L", pkg.HERE, "-llmot",
                    "-Wl,-rpath," + pkg.HERE], check=True)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 2 and "usage" in r.stderr
    if not torch.cuda.is_available():
        f = tmp_path / "frames.bin"
        np.zeros((2, 64, 4), np.float32).tofile(f)
        r = subprocess.run([exe, str(f), "2", "64", "1", "1"], capture_output=True, text=True)
        assert r.returncode != 0 and r.stdout.strip() == ""


def test_pinned_alloc_without_a_gpu_returns_null(pkg):
    import torch
    if torch.cuda.is_available():
        return
    lib = pkg.load_library()
    lib.lmot_pinned_alloc.restype = ctypes.c_void_p
    lib.lmot_pinned_alloc.argtypes = [ctypes.c_size_t]
    assert not lib.lmot_pinned_alloc(1 << 20)
    lib.lmot_pinned_free.argtypes = [ctypes.c_void_p]
    lib.lmot_pinned_free(None)


def test_shared_tracker_library_exports_every_declared_symbol(pkg):
    """include/lmot_shared.h <-> liblmot_shared.so (C++ over liblmot.so + NCCL); no compute without GPUs."""
    lib = pkg.load_shared_library()
    header = open(os.path.join(os.path.dirname(pkg.HERE), "include", "lmot_shared.h")).read()
    declared = sorted(set(re.findall(r"\b(lmot_shared_[a-z0-9_]+)\s*\(", header)))
    assert declared and set(declared) == set(pkg.SHARED_ABI_SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), name
