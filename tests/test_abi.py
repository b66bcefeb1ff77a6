"""CPU-only checks of the drop-in boundary: the shared library loads without a GPU and exports exactly the
entry points include/pislam_b200.h declares; the C++ template wrappers in include/pislam/ have the
reference's signatures (reference Fast.h:54-56,166-168,196-198, Harris.h:80-82, Orb.h:80-82,310,396-398,
Brief.h:637-639, Util.h:27-45); there is no CPU fallback (creating a context without a GPU fails)."""
import os
import re
import subprocess
import tempfile

import numpy as np
import pytest

from pislam_b200 import api, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "pislam_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pislam_b200_\w+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = api.load_library()
    names = declared_symbols()
    assert len(names) >= 28
    for n in names:
        assert hasattr(lib, n), "libpislam_b200.so does not export " + n
    assert sorted(api._SIGNATURES) == names, "api._SIGNATURES and include/pislam_b200.h disagree"
    out = subprocess.run(["nm", "-D", "--defined-only", build.LIB_PATH], stdout=subprocess.PIPE, text=True).stdout
    exported = sorted(set(re.findall(r"\b(pislam_b200_\w+)", out)))
    assert exported == names, "exported but undeclared (or the reverse): %s" % (set(exported) ^ set(names))


def test_library_is_sm100a_native():
    out = subprocess.run(["cuobjdump", "-lelf", build.LIB_PATH], stdout=subprocess.PIPE, text=True).stdout
    assert "sm_100a" in out
    sass = subprocess.run(["cuobjdump", "-sass", build.LIB_PATH], stdout=subprocess.PIPE, text=True).stdout
    assert "UTMALDG" in sass, "the fused tile kernel must stage its tile with TMA"
    assert "VABSDIFF4" in sass and "IDP.4A" in sass


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(api.PislamError) as e:
        api.Context(0)
    assert e.value.status == api.PISLAM_ERR_NO_DEVICE


def test_record_layout_and_host_helpers():
    lay = api.record_layout(2048, 8)
    assert lay["kp"] == 128 and lay["desc"] == 128 + 4 * 2048 and lay["level"] == lay["desc"] + 32 * 2048
    assert lay["bytes"] % 16 == 0 and lay["bytes"] >= lay["level"] + 2048
    assert api.record_layout(7, 0)["desc"] == api.record_layout(7, 0)["level"]
    e = api.encodeFast(137, 291, 16)
    assert (api.decodeFastScore(e), api.decodeFastX(e), api.decodeFastY(e)) == (137, 291, 16)
    assert api.rencodeFastScore(7, e) == (7 << 24) | (e & 0xFFFFFF)


def test_device_brief_table_matches_reference(golden):
    assert np.array_equal(api.brief_table(), golden["brief_table"])


SIGNATURE_TU = r"""
#include <cstdint>
#include <vector>
#include "pislam/Fast.h"
#include "pislam/Harris.h"
#include "pislam/Orb.h"
#include "pislam/Brief.h"
#include "pislam/Util.h"
#include "pislam/Gaussian.h"
#include "pislam/Bilinear.h"
#include "pislam/Match.h"  // our addition (SURVEY.md 8f N2): must at least instantiate
std::vector<pislam::Match> (*p_match)(const std::vector<uint32_t>&, const std::vector<uint32_t>&, int,
                                     const std::vector<uint32_t>*, const std::vector<uint8_t>*,
                                     const std::vector<uint32_t>*, const std::vector<uint8_t>*) = &pislam::matchHamming<8>;
// exact function types of the reference's templates (vstep 640, border 16, words 8)
typedef uint8_t Row[640];
void (*p_detect)(const int, const int, Row*, Row*, int) = &pislam::fastDetect<640, 16>;
void (*p_score)(int, int, Row*, int32_t, Row*) = &pislam::fastScoreHarris<640, 16>;
std::vector<uint32_t> (*p_extract)(const int, const int, Row*, std::vector<uint32_t>&) = &pislam::fastExtract<640, 16>;
std::vector<uint32_t> (*p_extract_b)(const int, const int, Row*, std::vector<uint32_t>&) = &pislam::fastExtract<640, 16, 4, 3>;
uint8_t (*p_harris)(Row*, int, int, int32_t) = &pislam::harrisScoreSobel<640>;
std::vector<int32_t> (*p_centroids)(Row*, const std::vector<uint32_t>&) = &pislam::orbCentroids<640>;
std::vector<uint8_t> (*p_atan2)(const std::vector<int32_t>&) = &pislam::atan2;
void (*p_orb)(Row*, const std::vector<uint32_t>&, std::vector<uint32_t>&) = &pislam::orbCompute<640, 8>;
void (*p_brief)(Row*, int, int, int, uint32_t*) = &pislam::briefDescribe<640, 8>;
void (*p_gauss)(const int, const int, Row*, Row*) = &pislam::gaussian5x5<640>;     // Gaussian.h:48-50
void (*p_b78)(const int, const int, Row*, Row*) = &pislam::bilinear7_8<640>;       // Bilinear.h:42-44
void (*p_b1316)(const int, const int, Row*, Row*) = &pislam::bilinear13_16<640>;   // Bilinear.h:165-167
uint32_t (*p_enc)(uint32_t, uint32_t, uint32_t) = &pislam::encodeFast;
uint32_t (*p_renc)(uint32_t, uint32_t) = &pislam::rencodeFastScore;
uint32_t (*p_dx)(uint32_t) = &pislam::decodeFastX;
uint32_t (*p_dy)(uint32_t) = &pislam::decodeFastY;
uint32_t (*p_ds)(uint32_t) = &pislam::decodeFastScore;
int main() {
  uint32_t e = pislam::encodeFast(137, 291, 16);
  return (pislam::decodeFastX(e) == 291 && pislam::decodeFastY(e) == 16 && pislam::decodeFastScore(e) == 137 &&
          pislam::rencodeFastScore(7, e) == ((7u << 24) | (e & 0xffffffu))) ? 0 : 1;
}
"""


def test_cpp_wrappers_have_reference_signatures():
    with tempfile.TemporaryDirectory() as tmp:
        src = os.path.join(tmp, "sig.cpp")
        exe = os.path.join(tmp, "sig")
        open(src, "w").write(SIGNATURE_TU)
        api.load_library()
        subprocess.check_call(["g++", "-std=c++11", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), src, "-o", exe,
                               "-L", build.LIB_DIR, "-lpislam_b200", "-Wl,-rpath," + build.LIB_DIR])
        assert subprocess.call([exe]) == 0  # the Util.h helpers need no GPU
