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
        "PrintMachine::GetBackBuffer()", "PrintMachine::GetMaxSize()", "PrintMachine::Print()",
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
    subprocess.check_call([os.path.join(HOST, "facade_test"), str(tmp_path), "engine", "3"])
    got = np.fromfile(tmp_path / "engine_240x64_m3.bin", np.uint8)
    assert np.array_equal(got, golden["default_240x64_m3_stream"])
    # pipelined sink (rtc_submit / rtc_collect behind RayTracingManager::Update) + per-tile culling: same bytes
    subprocess.check_call([os.path.join(HOST, "facade_test"), str(tmp_path), "pipelined", "4"])
    got = np.fromfile(tmp_path / "pipelined_240x64_m3.bin", np.uint8)
    assert np.array_equal(got, golden["default_240x64_m3_stream"])


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
