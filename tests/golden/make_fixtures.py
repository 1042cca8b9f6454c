"""Extracts the fixture genome of SequentialGuideAlignerTest.scala:12-44 into sga_test_ref.fa.
Runs only where /root/reference exists (the build container); the output is committed."""
import os, re
SRC = "/root/reference/calitas/src/test/scala/com/editasmedicine/aligner/SequentialGuideAlignerTest.scala"
if __name__ == "__main__":
    src = open(SRC).read()
    chr1 = re.findall(r'\.add\("([ACGT]{100})"\)', src)
    assert len(chr1) == 24
    chr2 = ["GATACaaCTCGTACTGTCAGT", "GATACGTCTCGTACTGTCAtT"]
    with open(os.path.join(os.path.dirname(__file__), "sga_test_ref.fa"), "w") as f:
        f.write(">chr1\n")
        for l in chr1:
            f.write(l + "\n")
        f.write(">chr2\n" + "".join(chr2) + "\n")
