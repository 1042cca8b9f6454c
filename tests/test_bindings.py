"""The JVM-side binding (bindings/): there is no JDK in this image, so the shim cannot run; what can be checked is checked --
the C shim compiles warning-free against a declarations-only jni.h and links against the product ABI, its exports and Native.scala's
@native methods name each other one to one, and the record offsets HitDecoder.scala reads are the ones of the C struct."""
import ctypes as C
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
B = os.path.join(ROOT, "bindings")


def test_jni_shim_compiles_and_matches_native_scala(tmp_path):
    obj = tmp_path / "jni.o"
    subprocess.check_call(["gcc", "-std=c11", "-Wall", "-Wextra", "-Werror", "-O1", "-fPIC", "-c", "-I" + os.path.join(B, "jni", "stub"), "-I" + os.path.join(ROOT, "include"),
                           os.path.join(B, "jni", "calitas_b200_jni.c"), "-o", str(obj)])
    syms = subprocess.check_output(["nm", str(obj)], text=True)
    exported = {m.group(1) for m in re.finditer(r" T Java_com_editasmedicine_aligner_b200_Native_(\w+)", syms)}
    undefined = {m.group(1) for m in re.finditer(r" U (calitas_\w+)", syms)}
    scala = open(os.path.join(B, "scala", "com", "editasmedicine", "aligner", "b200", "Native.scala")).read()
    declared = set(re.findall(r"@native def (\w+)\(", scala))
    assert exported == declared and len(declared) == 9
    header = open(os.path.join(ROOT, "include", "calitas_b200.h")).read()
    assert undefined and all(re.search(r"\b%s\(" % u, header) for u in undefined)         # every ABI call the shim makes is declared in the header


def test_hit_decoder_offsets_match_the_struct():
    from calitas_b200 import _capi
    H = _capi.Hit
    assert C.sizeof(H) == 72
    scala = open(os.path.join(B, "scala", "com", "editasmedicine", "aligner", "b200", "HitDecoder.scala")).read()
    assert "RecordBytes = 72" in scala
    want = {"guide_idx": 0, "pam_idx": 4, "contig_idx": 8, "task_idx": 12, "start_offset": 16, "end_offset": 20, "guide_start_offset": 24,
            "guide_end_offset": 28, "score": 32, "strand": 36, "n_ops": 37, "gap_bases": 38, "edits": 39, "ops": 40}
    for name, off in want.items():
        assert getattr(H, name).offset == off, name
    for off in (4, 8, 12, 16, 20, 24, 28, 32, 36, 37):
        assert ("o + %d" % off) in scala
