"""The C-ABI library loads and exports every symbol include/qnb.h declares (no compute without a GPU)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "qnb.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(qnb_[a-z_0-9]+)\s*\(", text)))


def test_header_symbols_exported():
    from q6_b200 import engine
    lib = engine.load_library()
    names = _declared()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), f"libqnb.so does not export {n}"


def test_struct_layout_matches_header():
    """ctypes image of qnb_system has the header's field order (names parsed from the header)."""
    from q6_b200.system import qnb_system
    text = open(os.path.join(ROOT, "include", "qnb.h")).read()
    start = text.index("typedef struct qnb_system {") + len("typedef struct qnb_system {")
    body = text[start:text.index("} qnb_system;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = []
    for decl in body.split(";"):
        decl = decl.strip()
        m = re.match(r"(?:const\s+)?(int32_t|double)\s*(.*)$", decl, flags=re.S)
        if not m:
            continue
        for nm in m.group(2).split(","):
            nm = nm.strip().lstrip("*").strip()
            nm = re.sub(r"\[.*\]", "", nm)
            if nm:
                fields.append(nm)
    assert fields == [f[0] for f in qnb_system._fields_]


def test_no_gpu_fails_loudly():
    """Without a CUDA device construction raises; there is no CPU fallback."""
    from q6_b200 import engine, synth
    lib = engine.load_library()
    if lib.qnb_device_count() > 0:
        pytest.skip("a GPU is present")
    q = synth.solvated_sphere(10.0, 5.0, 4, 1, 3)
    with pytest.raises(engine.QnbError):
        engine.Qnb(q)


def test_product_does_not_import_oracle():
    """The shipped package never references oracle/ (only tests, smoke() and bench's CPU legs may)."""
    pkg = os.path.join(ROOT, "q6_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h", ".cpp")):
                src = open(os.path.join(dp, f), errors="replace").read()
                assert "qoracle" not in src and "pyoracle" not in src and "oracle/" not in src, f


def test_fortran_interface_module_matches_header():
    """integration/qnb_mod.f90 cannot be compiled here (no Fortran compiler), so at least its text is held against the
    header: type(qnb_system) lists the C struct's fields in the same order with matching kinds, and every C prototype of
    include/qnb.h that the Fortran host calls has a bind(c) interface of that name."""
    from q6_b200.system import qnb_system
    import ctypes as C
    f90 = open(os.path.join(ROOT, "integration", "qnb_mod.f90")).read()
    body = f90[f90.index("type, bind(c) :: qnb_system"):f90.index("end type qnb_system")]
    fields = []
    for line in body.splitlines()[1:]:
        line = line.split("!")[0].strip()
        if "::" not in line:
            continue
        kind, names = line.split("::")
        kind = kind.strip().replace(" ", "")
        for nm in names.split(","):
            nm = re.sub(r"\(.*\)", "", nm).strip()
            if nm:
                fields.append((nm, kind))
    want = []
    for nm, t in qnb_system._fields_:
        kind = {C.c_int32: "integer(c_int32_t)", C.c_double: "real(c_double)"}.get(t)
        if kind is None:
            kind = "real(c_double)" if getattr(t, "_type_", None) is C.c_double and hasattr(t, "_length_") else "type(c_ptr)"
        want.append((nm, kind))
    assert fields == want
    # measurement hooks and debug exports are not part of the Fortran host's surface
    host_calls = [n for n in _declared() if not n.startswith(("qnb_bench_", "qnb_last_timing", "qnb_last_copy_bytes",
                                                              "qnb_launch_count"))]
    assert len(host_calls) >= 18
    for n in host_calls:
        assert re.search(rf"bind\(c,\s*name='{n}'\)", f90), f"no Fortran interface for {n}"


@pytest.mark.skipif(not os.path.isdir("/root/reference/src"), reason="reference tree not mounted")
def test_integration_patch_applies_to_the_reference(tmp_path):
    """integration/q6_qnb.patch (the #if defined(USE_QNB) branches of make_pair_lists, pot_energy_nonbonds, nonbond_qq,
    nonbond_qqp, the two calls in qdyn.f90, the three MC_volume hooks in md.f90 and the removal of gather_nonbond / the
    master's sum for MPI runs) applies cleanly to the reference sources, and every procedure it calls
    exists in integration/qnb_glue.f90."""
    import shutil
    import subprocess
    src = tmp_path / "src"
    src.mkdir()
    for f in ("md.f90", "nonbondene.f90", "potene.f90", "qdyn.f90"):
        shutil.copy(os.path.join("/root/reference/src", f), src / f)
    patch = os.path.join(ROOT, "integration", "q6_qnb.patch")
    out = subprocess.run(["patch", "-p1", "--dry-run", "-i", patch], cwd=tmp_path, capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    glue = open(os.path.join(ROOT, "integration", "qnb_glue.f90")).read().lower()
    called = set(re.findall(r"^\+\s*call (qnb_\w+)", open(patch).read(), flags=re.M))
    assert called == {"qnb_glue_make_pair_lists", "qnb_glue_add_qq", "qnb_glue_nonbond", "qnb_setup", "qnb_shutdown",
                      "qnb_glue_save_lists", "qnb_glue_update_box", "qnb_glue_restore_lists"}
    for name in called:
        assert f"subroutine {name}" in glue, name
