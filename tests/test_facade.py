"""The C++ facade (reference class names over the C-ABI): API surface on CPU, behaviour on the GPU."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "raytracing-in-windows-console_b200", "host")


def build_facade():
    import rtc_b200
    rtc_b200.load_library()                                  # builds librtc_b200.so if needed
    subprocess.check_call(["make", "-s", "-C", HOST])


def test_facade_exposes_reference_api():
    """Same class / method names as the reference: Engine3D, Scene3D, Camera3D,
    RayTracingManager, PrintMachine -- checked on the built archive, no GPU needed."""
    build_facade()
    syms = subprocess.run(["nm", "-C", "--defined-only", os.path.join(HOST, "libconsole_rt_facade.a")],
                          capture_output=True, text=True, check=True).stdout
    for want in [
        "Engine3D::Start()", "Engine3D::Run()", "Engine3D::CleanUp()", "Engine3D::Render(long double)",
        "Scene3D::Init()", "Scene3D::CleanUp()", "Scene3D::GetObjects()",
        "Scene3D::CreateSphere(float, MyMath::Vector3 const&, MyMath::Vector3 const&)",
        "Scene3D::CreatePlane(MyMath::Vector3 const&, MyMath::Vector3 const&, MyMath::Vector3 const&, float, float)",
        "Camera3D::Init()", "Camera3D::Update()", "Camera3D::GetInverseVMatrix() const", "Camera3D::Move(long double)",
        "Camera3D::AddRot(long double, short, short, short)", "Camera3D::SetPos(float, float, float)",
        "RayTracingManager::Update(RayTracingCPUToGPUData const&, DeviceObjectArray<Object3D*> const&, double)",
        "RayTracingManager::SetRenderingMode(RenderingMode)", "RayTracingManager::SetPipelined(bool)", "RayTracingManager::Flush()",
        "PrintMachine::Start(unsigned long, unsigned long)", "PrintMachine::SetDataInBackBuffer(char const*, unsigned long)",
        "PrintMachine::GetBackBuffer()", "PrintMachine::GetMaxSize()", "PrintMachine::Print()", "PrintMachine::TerminateThread()",
        "RayTracing::RayTrace(dim3 const&, dim3 const&, Object3D**, unsigned int, RayTracingCPUToGPUData const*, char*, RenderingMode)",
    ]:
        assert want in syms, f"facade lacks {want}"


@pytest.mark.gpu
def test_facade_frames_match_reference(golden, tmp_path):
    """Engine3D-style start-up + RayTracingManager::Update per mode == the reference's bytes."""
    build_facade()
    subprocess.check_call([os.path.join(HOST, "facade_test"), str(tmp_path)])
    for mode in range(6):
        got = np.fromfile(tmp_path / f"default_240x64_m{mode}.bin", np.uint8)
        assert np.array_equal(got, golden[f"default_240x64_m{mode}_stream"]), f"mode {mode}"
    # facade defaults: pipelined sink (rtc_submit / rtc_collect behind RayTracingManager::Update) + per-tile culling
    subprocess.check_call([os.path.join(HOST, "facade_test"), str(tmp_path), "engine", "4"])
    got = np.fromfile(tmp_path / "engine_240x64_m3.bin", np.uint8)
    assert np.array_equal(got, golden["default_240x64_m3_stream"])
    # the reference's synchronous hand-over, brute force: same bytes
    subprocess.check_call([os.path.join(HOST, "facade_test"), str(tmp_path), "sync", "3"])
    got = np.fromfile(tmp_path / "sync_240x64_m3.bin", np.uint8)
    assert np.array_equal(got, golden["default_240x64_m3_stream"])


@pytest.mark.gpu
@pytest.mark.parametrize("gather", ["host", "p2p"])
def test_facade_on_two_gpus(golden, tmp_path, gather):
    """RTC_GPUS=2: the reference-facing call RayTracingManager::Update renders through the multi-GPU frame driver
    (rtc_mgpu_*, two row bands) and lands the reference's bytes in PrintMachine.  Uses two real GPUs when the box has
    them, two contexts on GPU 0 otherwise."""
    import torch
    build_facade()
    two = torch.cuda.device_count() >= 2
    env = dict(os.environ, RTC_GPUS="2", RTC_DEVICES="0,1" if two else "0,0", RTC_GATHER=gather)
    subprocess.check_call([os.path.join(HOST, "facade_test"), str(tmp_path), "engine", "5"], env=env)
    got = np.fromfile(tmp_path / "engine_240x64_m3.bin", np.uint8)
    assert np.array_equal(got, golden["default_240x64_m3_stream"])
    subprocess.check_call([os.path.join(HOST, "facade_test"), str(tmp_path), "sync", "2"], env=env)
    got = np.fromfile(tmp_path / "sync_240x64_m3.bin", np.uint8)
    assert np.array_equal(got, golden["default_240x64_m3_stream"])


@pytest.mark.gpu
def test_facade_inner_seam_raw_cells(oracle, tmp_path):
    """RayTracing::RayTrace (reference RayTracing.h:31-38): the raw 20*x*y-byte cell buffer, byte for byte what the
    reference's kernels leave in m_deviceResultArray (oracle.trace_raw restates them; pinned in test_oracle_vs_reference)."""
    from rtc_b200 import scenes
    build_facade()
    subprocess.check_call([os.path.join(HOST, "facade_test"), str(tmp_path), "raw"])
    p = scenes.config_camera("config1_240x64")
    objs = scenes.default_scene()                                       # RayTracing::RayTrace runs no physics step
    for mode in range(6):
        got = np.fromfile(tmp_path / f"raw_240x64_m{mode}.bin", np.uint8)
        want = oracle.trace_raw(objs, p, mode)
        assert got.size == 20 * 240 * 64
        assert np.array_equal(got, want), f"raw cell buffer differs in mode {mode}"


@pytest.mark.gpu
def test_facade_print_thread(golden, tmp_path):
    """The reference's print thread (PrintMachine.cpp:257-306) into a file: every printed frame is ESC[H + the stream
    handed to SetDataInBackBuffer + ESC[m + the two FPS lines."""
    build_facade()
    subprocess.check_call([os.path.join(HOST, "facade_test"), str(tmp_path), "print", "6"])
    data = (tmp_path / "printed.bin").read_bytes()
    frame = golden["default_240x64_m3_stream"].tobytes()
    assert data.startswith(b"\x1b[H" + frame + b"\x1b[mRendering FPS: ")
    parts = data.split(b"\x1b[H")
    assert parts[0] == b"" and 1 <= len(parts) - 1 <= 6
    for part in parts[1:]:
        assert part.startswith(frame + b"\x1b[mRendering FPS: ") and part.rstrip().split(b"\n")[-1].startswith(b"Printing FPS: ")


def test_facade_fails_loudly_without_gpu(tmp_path):
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    build_facade()
    r = subprocess.run([os.path.join(HOST, "facade_test"), str(tmp_path)], capture_output=True, text=True)
    assert r.returncode != 0 and "no usable CUDA device" in (r.stderr + r.stdout)
