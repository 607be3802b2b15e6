"""The generated Java drop-in classes (java/gen_java.py) expose every public method of the reference classes with the
same parameter types.  The Java sources cannot be compiled here (no JDK); this is the by-inspection check, automated.
Needs the reference checkout (skipped on the GPU box)."""
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/src/main/java/org/jtransforms"
SIG = re.compile(r"public\s+(?:final\s+)?(?:void|int|long|double|float)\s+(\w+)\s*\(([^)]*)\)", re.S)


def signatures(path):
    src = open(path).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    out = set()
    for name, args in SIG.findall(src):
        if name == "run":
            continue
        types = []
        for a in args.split(","):
            a = a.strip()
            if not a:
                continue
            a = re.sub(r"\bfinal\s+", "", a)
            types.append(a.rsplit(None, 1)[0].replace(" ", ""))
        out.add((name, tuple(types)))
    return out


def classes():
    for p in ("Double", "Float"):
        for r in (1, 2, 3):
            yield "fft", "%sFFT_%dD" % (p, r)
            for k in ("DCT", "DST", "DHT"):
                yield k.lower(), "%s%s_%dD" % (p, k, r)
    yield "fft", "RealFFTUtils_2D"
    yield "fft", "RealFFTUtils_3D"


def test_common_utils_user_surface():
    """utils/CommonUtils: the user-facing statics (thresholds, power-of-two helpers, getReminder: reference lines 67-332)
    exist with the same signatures; the Ooura routines below them are the hot path itself and are not mirrored"""
    mine = signatures(os.path.join(ROOT, "java", "org", "jtransforms", "utils", "CommonUtils.java"))
    if not os.path.isdir(REF):
        pytest.skip("reference checkout not present")
    src = open(os.path.join(REF, "utils", "CommonUtils.java")).read().split("public static void makeipt")[0]
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    ref = set()
    for m in re.finditer(r"public\s+static\s+(?:void|int|long|boolean)\s+(\w+)\s*\(([^)]*)\)", src):
        types = tuple(re.sub(r"\bfinal\s+", "", a.strip()).rsplit(None, 1)[0].replace(" ", "") for a in m.group(2).split(",") if a.strip())
        ref.add((m.group(1), types))
    got = set()
    for m in re.finditer(r"public\s+static\s+(?:void|int|long|boolean)\s+(\w+)\s*\(([^)]*)\)", open(os.path.join(ROOT, "java", "org", "jtransforms", "utils", "CommonUtils.java")).read()):
        types = tuple(a.strip().rsplit(None, 1)[0].replace(" ", "") for a in m.group(2).split(",") if a.strip())
        got.add((m.group(1), types))
    assert not sorted(ref - got), sorted(ref - got)


def test_generator_is_up_to_date(tmp_path):
    before = {}
    for pkg, cls in classes():
        before[cls] = open(os.path.join(ROOT, "java", "org", "jtransforms", pkg, cls + ".java")).read()
    subprocess.run([sys.executable, os.path.join(ROOT, "java", "gen_java.py")], check=True, capture_output=True)
    for pkg, cls in classes():
        assert open(os.path.join(ROOT, "java", "org", "jtransforms", pkg, cls + ".java")).read() == before[cls], cls


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not present")
@pytest.mark.parametrize("pkg,cls", list(classes()))
def test_public_surface_matches_reference(pkg, cls):
    ref = signatures(os.path.join(REF, pkg, cls + ".java"))
    mine = signatures(os.path.join(ROOT, "java", "org", "jtransforms", pkg, cls + ".java"))
    # protected/internal helpers of the reference are not part of the surface; only its public methods are listed
    missing = sorted(s for s in ref if s not in mine)
    assert not missing, "%s lacks %s" % (cls, missing)
    src = open(os.path.join(ROOT, "java", "org", "jtransforms", pkg, cls + ".java")).read()
    assert "org.visnow.jlargearrays" in src and "pl.edu.icm" not in src
