"""CPU-side checks of the C ABI: the library loads without a GPU, exports exactly the symbols the
header declares, and fails loudly (no CPU fallback) when no device is present."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "sbte_b200.h")


def _declared():
    txt = open(HEADER).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    txt = "\n".join(l for l in txt.splitlines() if not l.lstrip().startswith("#"))  # drop the macro definition
    names = re.findall(r"SBTE_API[^;(]*?(\w+)\s*\(", txt)
    return sorted(set(names))


def test_header_declares_reference_link_interface():
    d = _declared()
    for name in ["initialize_coll", "dealloc_coll", "ComputeQ", "ComputeQ_maxPreserve", "fft3D",
                 "initialize_conservation", "initialize_conservation_fast", "conserveAllMoments",
                 "dealloc_conservation", "initialize_transport", "advectOne", "advectTwo", "dealloc_trans"]:
        assert name in d, name


def test_library_exports_every_declared_symbol():
    from spectralbte_b200 import _lib, build
    build.build()
    L = _lib.load()  # raises AttributeError for a declared-but-missing symbol
    declared = _declared()
    assert sorted(_lib.PROTOTYPES) == declared
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = sorted(line.split()[-1] for line in out.splitlines() if " T " in line)
    assert exported == declared  # nothing else leaks out of the library
    for name in declared:
        assert ctypes.cast(getattr(L, name), ctypes.c_void_p).value


def test_species_struct_layout_matches_reference():
    """sbte_species must mirror `species` (/root/reference/src/species.h:11-26): 8+8+8+8*8+8+8+80."""
    from oracle import oracle as orc
    assert ctypes.sizeof(orc._Species) == 3 * 8 + 8 * 8 + 2 * 8 + 80
    assert orc._Species.mass.offset == 32 and orc._Species.name.offset == 104
    hdr = open(HEADER).read()
    body = hdr[hdr.index("typedef struct sbte_species"):hdr.index("} sbte_species;")]
    order = re.findall(r"\b(id|num_levels|lev_id|Rgas|mass|mm|d_ref|T_ref|mu_ref|omega|E0|Ei|gi|name)\b", body)
    assert order == ["id", "num_levels", "lev_id", "Rgas", "mass", "mm", "d_ref", "T_ref", "mu_ref", "omega",
                     "E0", "Ei", "gi", "name"]


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from spectralbte_b200 import Collisions
    from spectralbte_b200._lib import SbteError
    with pytest.raises(SbteError, match="no CUDA device"):
        Collisions(8, 5.0)


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under spectralbte_b200/ may reference it."""
    pkg = os.path.join(ROOT, "spectralbte_b200")
    for base, _, files in os.walk(pkg):
        if "_obj" in base:
            continue
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".c")):
                txt = open(os.path.join(base, fn)).read()
                assert "oracle" not in txt.lower().replace("test oracle", ""), os.path.join(base, fn)
