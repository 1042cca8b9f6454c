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
    assert exported == declared and len(declared) == 11
    header = open(os.path.join(ROOT, "include", "calitas_b200.h")).read()
    assert undefined and all(re.search(r"\b%s\(" % u, header) for u in undefined)         # every ABI call the shim makes is declared in the header


def test_hit_decoder_offsets_match_the_struct():
    from calitas_b200 import _capi
    H, W = _capi.Hit, _capi.HitWide
    assert C.sizeof(H) == 32 and C.sizeof(W) == 64
    scala = open(os.path.join(B, "scala", "com", "editasmedicine", "aligner", "b200", "HitDecoder.scala")).read()
    assert "HeaderBytes = 20" in scala
    want = {"start_offset": 0, "task_idx": 4, "score": 8, "where": 12, "shape": 16, "ops": 20}
    for name, off in want.items():
        assert getattr(H, name).offset == off and getattr(W, name).offset == off, name
    for off in (4, 8, 12, 16):
        assert ("o + %d" % off) in scala
    header = open(os.path.join(ROOT, "include", "calitas_b200.h")).read()
    # the bit fields the Scala decoder unpacks are the header's
    for sc, hd in (("where & 0x1FFF", "h->where & 0x1FFFu"), ("(where >>> 13) & 0x3FFFF", "(h->where >> 13) & 0x3FFFFu"), ("(where >>> 31)", "(h->where >> 31)"),
                   ("shape & 0xFF", "h->shape & 0xFFu"), ("(shape >>> 8) & 0xFF", "(h->shape >> 8) & 0xFFu"), ("(shape >>> 16) & 0x3F", "(h->shape >> 16) & 0x3Fu"),
                   ("(shape >>> 22) & 0x3F", "(h->shape >> 22) & 0x3Fu"), ("(shape >>> 28) - 1", "(h->shape >> 28) - 1")):
        assert sc in scala and hd in header, (sc, hd)


def test_decode_encode_round_trip():
    import numpy as np
    from calitas_b200 import _capi
    rng = np.random.default_rng(5)
    n = 200
    w = np.zeros((n, 16), dtype=np.uint32)
    w[:, 0] = rng.integers(0, 1 << 30, n); w[:, 1] = rng.integers(0, 1 << 30, n); w[:, 2] = rng.integers(-5000, 5000, n).astype(np.int32).view(np.uint32)
    w[:, 3] = rng.integers(0, 8192, n) | (rng.integers(0, 1 << 18, n) << 13) | (rng.integers(0, 2, n) << 31)
    nops = rng.integers(1, 129, n)
    w[:, 4] = nops | (rng.integers(0, 129, n) << 8) | (rng.integers(0, 64, n) << 16) | (rng.integers(0, 64, n) << 22) | (rng.integers(0, 10, n) << 28)
    for i in range(n):
        for k in range(int(nops[i])):
            w[i, 5 + (k >> 4)] |= np.uint32(int(rng.integers(0, 4)) << ((k & 15) * 2))
    rec = _capi.decode_hits(w)
    assert np.array_equal(_capi.encode_hits(rec), w)
    for i in (0, 7, 100):
        ops = [(int(w[i, 5 + (k >> 4)]) >> ((k & 15) * 2)) & 3 for k in range(int(nops[i]))]
        assert rec["gap_bases"][i] == sum(o >= 2 for o in ops) and rec["edits"][i] == sum(o != 0 for o in ops)
