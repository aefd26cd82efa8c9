"""The C++ drop-in header (cpprcoder_b200/include/cpprcoder_b200.h) driven the way the
reference's harness drives cpprcoder.h (tests/cpp/harness.cpp)."""
import subprocess
import tarfile
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
PKG = ROOT / "cpprcoder_b200"


@pytest.fixture(scope="module")
def harness(built, tmp_path_factory):
    built.build_native()
    out = tmp_path_factory.mktemp("harness") / "harness"
    cmd = ["g++", "-O2", "-std=c++17", "-Wall", "-Wextra", str(ROOT / "tests" / "cpp" / "harness.cpp"), "-o", str(out),
           f"-L{PKG}", "-lb2rc", f"-Wl,-rpath,{PKG}"]
    subprocess.check_call(cmd)
    return out


def test_header_compiles_and_links(harness):
    assert harness.exists()


def test_without_a_device_calls_fail_loudly(harness):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a device is present")
    r = subprocess.run([str(harness), "--selftest"], capture_output=True, text=True)
    assert r.returncode != 0, "no CUDA device: the drop-in must report failure, not fall back to a CPU coder"


@pytest.mark.gpu
def test_selftests_on_gpu(harness):
    r = subprocess.run([str(harness), "--selftest"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "test_rangecoder ok" in r.stdout and "test_adaptive_chunked ok" in r.stdout


@pytest.mark.gpu
def test_canterbury_rows_on_gpu(harness, tmp_path, golden):
    with tarfile.open(ROOT / "tests" / "golden" / "cantrbry.tar.bz2", "r:bz2") as tf:
        tf.extractall(tmp_path, filter="data")
    files = sorted(str(p) for p in (tmp_path / "cantrbry").iterdir())
    r = subprocess.run([str(harness)] + files, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    rows = [line.split("|") for line in r.stdout.splitlines() if line.startswith("|")]
    assert len(rows) == 2 * len(files)
    # ratio column = original / container bytes: block framing costs a little against the whole-file README ratio
    for row in rows:
        name = Path(row[1]).name
        whole = golden["canterbury"][name]["bytes"] / golden["canterbury"][name]["whole"]["static"]["size"]
        assert 0.5 * whole < float(row[2]) < 1.6 * whole


@pytest.mark.gpu
def test_rans_rows_on_gpu(harness, tmp_path):
    # cppans::rANS through the drop-in header, both variants, the reference harness's call sequence
    with tarfile.open(ROOT / "tests" / "golden" / "cantrbry.tar.bz2", "r:bz2") as tf:
        tf.extractall(tmp_path, filter="data")
    files = sorted(str(p) for p in (tmp_path / "cantrbry").iterdir())
    r = subprocess.run([str(harness), "--ans"] + files, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    rows = [line.split("|") for line in r.stdout.splitlines() if line.startswith("|")]
    assert len(rows) == 2 * len(files)
    assert all(float(row[2]) > 0.9 for row in rows)


@pytest.mark.gpu
def test_blksort_rows_on_gpu(harness, tmp_path, golden):
    # blksort::BlkSort through the drop-in header: alone (ratio = size / encodeBound) and in front of the static coder
    with tarfile.open(ROOT / "tests" / "golden" / "cantrbry.tar.bz2", "r:bz2") as tf:
        tf.extractall(tmp_path, filter="data")
    files = sorted(str(p) for p in (tmp_path / "cantrbry").iterdir())
    r = subprocess.run([str(harness), "--blk"] + files, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    rows = [line.split("|") for line in r.stdout.splitlines() if line.startswith("|")]
    assert len(rows) == 2 * len(files)
    for alone, coded in zip(rows[0::2], rows[1::2]):
        n = golden["canterbury"][Path(alone[1]).name]["bytes"]
        assert abs(float(alone[2]) - n / ((n >> 15) * 32770 + n % 32768)) < 1e-5
        assert float(coded[2]) > 0.9


@pytest.mark.gpu
def test_file_cli_round_trip(tmp_path):
    import subprocess
    import sys
    with tarfile.open(ROOT / "tests" / "golden" / "cantrbry.tar.bz2", "r:bz2") as tf:
        tf.extractall(tmp_path, filter="data")
    src = tmp_path / "cantrbry" / "lcet10.txt"
    for flag in ([], ["--adaptive"], ["--coder", "rans"], ["--coder", "rans-word"]):
        enc, dec = tmp_path / "x.b2rc", tmp_path / "x.out"
        subprocess.check_call([sys.executable, "-m", "cpprcoder_b200", "encode", *flag, str(src), str(enc)], cwd=ROOT)
        subprocess.check_call([sys.executable, "-m", "cpprcoder_b200", "decode", str(enc), str(dec)], cwd=ROOT)
        assert dec.read_bytes() == src.read_bytes()
    r = subprocess.run([sys.executable, "-m", "cpprcoder_b200", "rows", str(src)], cwd=ROOT, capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.count("|") == 20  # four coders, five bars per row
    bs, back = tmp_path / "x.bs", tmp_path / "x.back"
    subprocess.check_call([sys.executable, "-m", "cpprcoder_b200", "blksort", str(src), str(bs)], cwd=ROOT)
    subprocess.check_call([sys.executable, "-m", "cpprcoder_b200", "unblksort", str(bs), str(back)], cwd=ROOT)
    assert back.read_bytes() == src.read_bytes() and bs.stat().st_size == src.stat().st_size + 2 * (src.stat().st_size >> 15)
    r = subprocess.run([sys.executable, "-m", "cpprcoder_b200", "rows", "--blk", str(src)], cwd=ROOT, capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.count("|") == 10


# ---- the reference's own harness (test/main.cpp), unchanged, against the drop-in headers ----------
REF_MAIN = ROOT / "oracle" / "_ref" / "ref_main_dropin"


def test_reference_harness_compiles_unchanged_against_the_dropin(built):
    """oracle/Makefile `dropin`: a copy of /root/reference/test/main.cpp in a temporary directory, one-line
    cpprcoder.h / blksort.h shims above it, -DUSE_RC -DUSE_ADAPTIVE (USE_BLKSORT is the file's own default).
    Only where the reference tree is mounted; the GPU box runs the prebuilt binary."""
    if not Path("/root/reference/test/main.cpp").exists():
        pytest.skip("reference tree not mounted here")
    built.build_native()
    if REF_MAIN.exists():
        REF_MAIN.unlink()
    subprocess.check_call(["make", "-C", str(ROOT / "oracle"), "dropin"], stdout=subprocess.DEVNULL)
    assert REF_MAIN.exists()


@pytest.mark.gpu
def test_reference_harness_runs_on_the_gpu(tmp_path, golden):
    """run_rangecoder / run_adaptive / run_blksort of the reference's main() (test/main.cpp:254-363, :791-841)
    over the Canterbury corpus: a row per file and coder, the harness's own byte compare silent, and the
    ratio column equal to input size / container size of the same call made through the C ABI."""
    if not REF_MAIN.exists():
        pytest.skip("oracle/_ref/ref_main_dropin was not built (needs the reference tree at build time)")
    import numpy as np
    from cpprcoder_b200 import api
    with tarfile.open(ROOT / "tests" / "golden" / "cantrbry.tar.bz2", "r:bz2") as tf:
        tf.extractall(tmp_path, filter="data")
    (tmp_path / "test").mkdir()
    r = subprocess.run([str(REF_MAIN)], cwd=tmp_path / "test", capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "!=" not in r.stdout, r.stdout[-2000:]          # the harness prints "[i] a != b" for every wrong byte
    sections, cur = {}, None
    for line in r.stdout.splitlines():
        if line.startswith("|"):
            sections[cur].append(line.split("|"))
        elif line and not line.startswith("-"):
            cur = line.strip()
            sections.setdefault(cur, [])
    assert len(sections["Range Coder"]) == 11 and len(sections["Adaptive Range Coder"]) == 11
    ctx = api.Context(0)
    try:
        for title, mode in (("Range Coder", 0), ("Adaptive Range Coder", 1)):
            for row in sections[title]:
                data = np.fromfile(tmp_path / "cantrbry" / Path(row[1]).name, dtype=np.uint8)
                made = ctx.encode(mode, data, 65536).size
                assert abs(float(row[2]) - data.size / made) < 2e-6, (title, row)
    finally:
        ctx.close()
    assert len(sections["BLKSORT"]) == 11, list(sections)
