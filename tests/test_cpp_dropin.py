"""The C++ drop-in headers (include/pislam/*.h) driven by the reference's own caller loop (examples/demo.cpp,
after reference demo/demo.cpp:77-101) on the reference's fixture; output pinned by the FNV hashes of
SURVEY.md section 8c (1754 keypoints).  Needs a B200."""
import os
import re
import subprocess
import tempfile

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def test_demo_loop_on_fixture(fixture_stack):
    from pislam_b200 import api
    api.load_library()
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "examples")], stdout=subprocess.DEVNULL)
    with tempfile.TemporaryDirectory() as tmp:
        raw = os.path.join(tmp, "stack.raw")
        fixture_stack.tofile(raw)
        out = subprocess.run([os.path.join(ROOT, "examples", "demo"), raw], stdout=subprocess.PIPE, text=True,
                             check=True, timeout=120).stdout
    lines = out.strip().splitlines()
    assert len(lines) == 3, out
    m = re.search(r"harrisScoreSobel: (\d+) of (\d+)", lines[1])
    assert m and m.group(1) == m.group(2) and int(m.group(2)) >= 8, lines[1]
    for ln in (lines[0], lines[2]):
        m = re.search(r"(\d+) features, keypoints ([0-9a-f]{16}), descriptors ([0-9a-f]{16})", ln)
        assert m, ln
        assert (m.group(1), m.group(2), m.group(3)) == ("1754", "1ab0b92100c9c317", "af0ceaba8d844310"), ln
