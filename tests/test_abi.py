"""CPU-side checks of the drop-in boundary: both libraries export every symbol include/sailor_pt.h declares, the
header and the ctypes mirror agree, and the product library refuses to compute without a CUDA device."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import ROOT
from sailor_b200.capi import ERR_ARG, ERR_NO_DEVICE, SYMBOLS, Library, Params, SailorPtParams, SailorPtStats
import scenes


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "sailor_pt.h")).read()
    return sorted(set(re.findall(r"SAILOR_PT_API\s+[\w\s\*]+?\b(SailorPt_\w+)\s*\(", text)))


def test_header_and_binding_list_the_same_symbols():
    assert _header_symbols() == sorted(SYMBOLS)


def test_product_library_builds_and_exports_every_symbol():
    from sailor_b200 import build as product_build
    lib = product_build.build()
    out = subprocess.run(["nm", "-D", "--defined-only", lib], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r"\bT (SailorPt_\w+)", out))
    assert set(SYMBOLS) <= exported, sorted(set(SYMBOLS) - exported)
    sass = subprocess.run(["cuobjdump", "-lelf", lib], capture_output=True, text=True).stdout
    assert "sm_100a" in sass, "the product library must carry an sm_100a image"


def test_oracle_exports_every_symbol(oracle):
    for s in SYMBOLS:
        getattr(oracle.lib, s)
    assert oracle.backend() == "reference-cpu"


def test_struct_layouts_match_the_header():
    # 3 pointers + 5 u32 + 3 f32 + u32 + (pad) u64 + 4 u32
    assert C.sizeof(SailorPtParams) == 96
    assert SailorPtParams.seed.offset == 64 and SailorPtParams.rowBegin.offset == 72
    assert C.sizeof(SailorPtStats) == 176 and SailorPtStats.devicesUsed.offset == 168


def test_product_fails_loudly_without_a_cuda_device(scene_dir):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    import sailor_b200
    from sailor_b200 import build as product_build
    product_build.build()
    L = sailor_b200.library()
    assert L.backend() == "cuda sm_100a"
    h = C.c_void_p()
    rc = L.lib.SailorPt_SceneLoad(scenes.ensure(scene_dir, "cube").encode(), C.byref(h))
    assert rc == ERR_NO_DEVICE and not h.value
    assert b"no CPU path" in L.lib.SailorPt_LastError()
    lin = np.zeros((4, 4, 3), np.float32)
    out = np.zeros((4, 4, 3), np.uint8)
    assert L.lib.SailorPt_OutputStage(4, 4, lin.ctypes.data_as(C.POINTER(C.c_float)), out.ctypes.data_as(C.POINTER(C.c_uint8))) == ERR_NO_DEVICE
    assert sailor_b200.PathTracer().Run(Params(path_to_model=scenes.ensure(scene_dir, "cube"), height=8)) == ERR_NO_DEVICE


def test_argument_errors(emu):
    assert emu.lib.SailorPt_SceneLoad(None, None) == ERR_ARG
    assert emu.lib.SailorPt_GetStats(None) == ERR_ARG


@pytest.mark.parametrize("which", ["oracle", "emu"])
def test_parse_command_line_args_matches_the_reference(which, request):
    lib = request.getfixturevalue(which)
    for samples, msaa, s in ((1, 1, 1), (3, 3, 1), (16, 4, 4), (32, 4, 8), (33, 8, 4), (256, 8, 32), (1024, 8, 128), (4096, 8, 512)):
        p = lib.parse_command_line_args(Params(), ["exe", "--in", "a b.glb", "--out", "o.png", "--height", "77", "--samples", str(samples),
                                                   "--bounces", "6", "--camera", "cam", "--ambient", "80ff00"])
        assert (p.m_pathToModel, p.m_output, p.m_camera) == ("a b.glb", "o.png", "cam")
        assert (p.m_height, p.m_maxBounces, p.m_msaa, p.m_numSamples) == (77, 6, msaa, s)
        assert tuple(np.float32(v) for v in p.m_ambient) == (np.float32(128) / np.float32(255), np.float32(1), np.float32(0))
        q = Params.from_samples(samples)
        assert (q.m_msaa, q.m_numSamples) == (msaa, s)
    p = lib.parse_command_line_args(Params(), ["exe", "--in", '"my', 'scene.glb"'])
    assert p.m_pathToModel == "my scene.glb"


def test_trim_memory_is_a_no_op_for_the_checkers(emu, oracle):
    emu.trim_memory(); oracle.trim_memory()


def test_cpp_host_builds_and_fails_loudly_without_a_cuda_device(scene_dir):
    """sailor_b200/sailor_pt (csrc/cli_main.cpp): the C++ caller the reference lacks, linked against the C-ABI only."""
    import subprocess
    import torch
    from sailor_b200 import build as product_build
    product_build.build()
    exe = product_build.build_cli()
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 1 and "usage:" in r.stderr and "cuda sm_100a" in r.stderr
    if not torch.cuda.is_available():
        r = subprocess.run([exe, "--in", scenes.ensure(scene_dir, "cube"), "--out", os.path.join(scene_dir, "cli.png")], capture_output=True, text=True)
        assert r.returncode == -ERR_NO_DEVICE and "no CPU path" in r.stderr


def test_header_is_plain_c_and_a_c_program_links_against_the_library(tmp_path):
    """The boundary is a C ABI: include/sailor_pt.h compiles as C99, and a C program (no C++ runtime of its own) links and calls it."""
    from sailor_b200 import build as product_build
    lib = product_build.build()
    src = tmp_path / "host.c"
    src.write_text('#include "sailor_pt.h"\n#include <stdio.h>\n#include <string.h>\n'
                   'int main(void) { SailorPtParams p; memset(&p, 0, sizeof p); const char* a[] = {"exe", "--samples", "64", "--bounces", "3"};\n'
                   '  if (SailorPt_ParseCommandLineArgs(&p, a, 5) != SAILOR_PT_OK) return 2;\n'
                   '  printf("%s %u %u %u\\n", SailorPt_Backend(), p.msaa, p.numSamples, p.maxBounces); return SailorPt_Run(0) == SAILOR_PT_ERR_ARG ? 0 : 3; }\n')
    exe = tmp_path / "host"
    inc = os.path.join(ROOT, "include")
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", inc, str(src), "-o", str(exe), "-L", os.path.dirname(lib), "-lsailor_pt_cuda",
                        "-Wl,-rpath," + os.path.dirname(lib)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.split() == ["cuda", "sm_100a", "8", "8", "3"], (r.returncode, r.stdout, r.stderr)
