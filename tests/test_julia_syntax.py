"""The Julia package cannot be executed in this image (no Julia).  Beyond the ccall-signature / struct-mirror checks of
tests/test_abi.py, this keeps the sources structurally sound: every block opener has its `end`, brackets balance."""
import glob
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import julia_block_check  # noqa: E402


def test_julia_sources_are_block_balanced():
    files = sorted(glob.glob(os.path.join(ROOT, "julia", "*.jl")))
    assert len(files) >= 2
    for f in files:
        assert julia_block_check.check(f) == [], f


def test_block_check_catches_a_missing_end(tmp_path):
    src = open(os.path.join(ROOT, "julia", "codegen.jl")).read()
    i = src.rfind("\nend")
    bad = tmp_path / "bad.jl"
    bad.write_text(src[:i] + src[i + 4:])
    assert julia_block_check.check(str(bad)) != []
