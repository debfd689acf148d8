"""The scan kernel's speed rests on properties of the GENERATED code that a harmless-looking source change can
silently lose (ptxas drops the uniform-register query operand as soon as a kernel contains a CALL, spills, or an
awkward unroll factor -- measured: 16.2 -> 20.1 ms).  This test reads the SASS of the built library (no GPU needed)
and checks them for the three shapes the engine selects by itself."""
import os
import re
import shutil
import subprocess

import pytest

from spotify_recommender_b200 import build

# scan_kernel<S, THREADS, CTAs/SM, DEFER, STAGE, DYN>: the small-batch (TMA-staged) and large-batch (dynamic) shapes
AUTO_SHAPES = {"small-batch S8xT256x2-dyn-tma": "ILi8ELi256ELi2ELb1ELi1ELb1E", "mid-batch S8xT256x2-dyn": "ILi8ELi256ELi2ELb1ELi0ELb1E",
               "large-batch S8xT512x1-dyn": "ILi8ELi512ELi1ELb1ELi0ELb1E"}


@pytest.fixture(scope="module")
def sass():
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(exe):
        pytest.skip("cuobjdump not available")
    build.build_engine()
    text = subprocess.run([exe, "-sass", build.ENGINE_SO], capture_output=True, text=True, check=True).stdout
    funcs = {}
    for chunk in text.split("Function : ")[1:]:
        name, _, body = chunk.partition("\n")
        funcs[name.strip()] = body
    return funcs


@pytest.mark.parametrize("label", sorted(AUTO_SHAPES))
def test_hot_loop_keeps_its_uniform_register_operands(sass, label):
    body = next((b for n, b in sass.items() if "scan_kernel" in n and AUTO_SHAPES[label] in n), None)
    assert body is not None, f"{label}: kernel not found in the library"
    ffma2 = [l for l in body.splitlines() if re.search(r"\bFFMA2\b", l)]
    with_ur = [l for l in ffma2 if re.search(r"\bUR\d+\.F32\b", l)]
    # 16 queries x 48 FFMA2 in the unrolled hot loop, all with the query value as a uniform scalar operand
    assert len(with_ur) >= 768, f"{label}: only {len(with_ur)} of {len(ffma2)} FFMA2 take the query from a uniform register"
    assert not re.search(r"\bCALL\b", body), f"{label}: a CALL appeared in the kernel (ptxas then drops the uniform operands)"
    assert not re.search(r"\b(STL|LDL)\b", body), f"{label}: local-memory traffic (register spills) in the kernel"
    assert re.search(r"\bLDCU", body), f"{label}: no uniform constant loads"


def test_small_batch_shape_uses_the_bulk_copy_engine(sass):
    for label in ("small-batch S8xT256x2-dyn-tma",):
        body = next(b for n, b in sass.items() if "scan_kernel" in n and AUTO_SHAPES[label] in n)
        assert re.search(r"\bUBLKCP\b", body) and re.search(r"\bSYNCS\b", body)  # cp.async.bulk + mbarrier
