"""CPU: the parity tests of the CUDA path, run on the kernel SOURCES compiled for the host.

tests/host_emu builds variantcalling_b200/csrc/kernels.cu (K1 field parse, K2 feature assembly, K3
inference) and capi.cu with g++ over a stand-in cuda_runtime.h -- one emulated thread per CTA, device
memory on the heap -- into tests/host_emu/_build/libugvc_emu.so.  This test points a child pytest at
that library (UGVC_LIB_PATH, the developer override of variantcalling_b200/lib.py) and runs the
`gpu`-marked parity files there, so the driver's CPU check already compares the code the GPU will
execute with the oracle: features bit-identical, FILTER identical, CLI output line for line, the
inputs the reference raises on, --treat_multiallelics, the fuzzed records.

Test infrastructure only: the package never loads the emulated library, and
tests/test_capi_cpu.py::test_no_cuda_device_fails_loudly keeps asserting that the product library
refuses to run without a CUDA device.  Not emulated (their tests stay GPU-only): K0's cooperative
tile scan (lines are indexed by a host loop with the same outputs and error codes), the device
synthetic generator and the concordance kernels (CUB pipelines)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU_DIR = os.path.join(ROOT, "tests", "host_emu")
EMU_LIB = os.path.join(EMU_DIR, "_build", "libugvc_emu.so")

FILES = ["tests/test_gpu_parity.py", "tests/test_gpu_edges.py", "tests/test_gpu_fuzz.py", "tests/test_gpu_cnv.py",
         "tests/test_gpu_cli.py", "tests/test_gpu_multiallelics.py"]
NEED_THE_DEVICE_GENERATOR = ["tests/test_gpu_edges.py::test_device_generator_text_parity",
                             "tests/test_gpu_parity.py::test_full_size_properties_batching_invariance"]


def test_kernel_sources_pass_the_parity_suite_on_the_host_emulation():
    build = subprocess.run(["make", "-C", EMU_DIR], capture_output=True, text=True, timeout=900)
    assert build.returncode == 0, build.stdout[-2000:] + build.stderr[-4000:]
    env = dict(os.environ, UGVC_LIB_PATH=EMU_LIB)
    cmd = [sys.executable, "-m", "pytest", "-q", "-m", "gpu", "-p", "no:cacheprovider", *FILES]
    for t in NEED_THE_DEVICE_GENERATOR:
        cmd += ["--deselect", t]
    run = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=3000)
    tail = run.stdout[-3000:] + run.stderr[-2000:]
    assert run.returncode == 0, tail
    assert " passed" in run.stdout and "failed" not in run.stdout.splitlines()[-1], tail
