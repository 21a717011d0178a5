"""CPU checks of the drop-in boundary: libhj_b200.so loads without a GPU, exports every symbol
include/hj.h declares, every symbol has a ctypes signature in the host mirror, and the compute entry
points fail loudly (no CPU fallback) when no CUDA device is present."""
import ctypes
import os
import importlib

import numpy as np
import pytest

hj = importlib.import_module("hephaestus-jit_b200")
L = importlib.import_module("hephaestus-jit_b200._lib")


def test_library_exports_every_declared_symbol():
    declared = L.declared_symbols()
    assert len(declared) > 100
    missing = [s for s in declared if not hasattr(L.lib, s)]
    assert not missing, f"libhj_b200.so does not export {missing}"
    unsigned = [s for s in declared if s not in L._SIGS]
    assert not unsigned, f"no ctypes signature for {unsigned}"


def test_abi_version_and_device_count():
    assert L.lib.hj_abi_version() >= 1
    assert L.lib.hj_device_count() >= 0


def test_no_cpu_fallback_without_a_device():
    if hj.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(hj.HjError) as e:
        hj.Device.cuda(0)
    assert e.value.status in (L.ERR_NO_DEVICE, L.ERR_CUDA)
    # null handles are rejected, not dereferenced
    assert L.lib.hj_reduce(None, 2, 7, 10, None, None) != 0
    assert L.lib.hj_prefix_sum(None, 7, 10, 1, None, None, None) != 0
    assert L.lib.hj_execute_graph(None, None, 0, None, None, 0, None) != 0
    assert L.last_error() != ""


def test_ir_codegen_and_cubin_work_without_a_gpu():
    irm = importlib.import_module("hephaestus-jit_b200.ir")
    ir = irm.c2_chain_ir().build()
    src = irm.codegen(ir)
    assert "hj_kernel_vec" in src and "fmaf" in src and "sinf" in src and "exp2f" in src
    cubin = irm.compile_cubin(ir)
    assert cubin[:4] == b"\x7fELF"
    assert irm.ir_hash(ir) == irm.ir_hash(irm.c2_chain_ir().build())  # stable content hash


def test_vector_entry_is_cut_in_two_only_for_read_modify_write_through_computed_indices():
    """codegen.cpp: split_point — a slot that is gathered AND scattered through computed indices makes the vector
    entry run the part of the element body in front of the first side effect for all elements of a thread first
    (every variable becomes `rN[e_]`); every other kernel keeps the single loop."""
    import ir_cases   # tests/ is on sys.path (pytest rootdir conftest)
    irm = importlib.import_module("hephaestus-jit_b200.ir")
    cut = {fn().name for fn in ir_cases.ALL_CASES if "[e_]" in irm.codegen(fn().builder.build())}
    assert cut == {"rmw_through_index_list", "rmw_then_loop"}
    passes, _ = irm.wavefront_step_passes(1 << 16)
    mask_update, a_update = (irm.codegen(p["ir"].build()) for p in passes[1:])
    assert "[e_]" not in mask_update      # `a` is only read there: __ldg loads move freely already
    assert "[e_]" in a_update
    body = a_update[a_update.index("hj_kernel_vec"):]
    assert body.index("((const f32*)b0)[") < body.index("((f32*)b0)[")   # all gathers, then all stores
    assert body.count("_Pragma(\"unroll\") for (int k = 0;") == 2



def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver times beside the GPU arm): one JSON line on stdout, the
    own arm's metric / unit / config, `cpu_baseline` describing the run and an `e2e` that moved no bytes."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    proc = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                          capture_output=True, text=True, timeout=300, cwd=root)
    assert proc.returncode == 0, proc.stderr[-2000:]
    lines = [ln for ln in proc.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, "stdout must carry the JSON line only"
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "GB/s" and d["higher_is_better"] is True and d["n_gpus"] == 1
    assert d["value"] > 0 and d["cpu_baseline"]["value"] == d["value"] == d["e2e"]["value"]
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["sample"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"] and d["gpu_launches"] == 0
