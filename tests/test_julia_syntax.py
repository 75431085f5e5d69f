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


def test_julia_sources_lex_cleanly():
    """Pygments' Julia lexer finds no token it cannot classify (unterminated strings, stray characters, bad escapes) in the
    package sources.  Regex and substitution literals (r"...", s"...": raw strings whose backslash sequences that lexer
    does not accept) are neutralised first."""
    import re
    from pygments.lexers import JuliaLexer
    from pygments.token import Token
    for f in sorted(glob.glob(os.path.join(ROOT, "julia", "*.jl"))):
        src = open(f, encoding="utf-8").read()
        src = re.sub(r'\b[rs]"([^"\\]|\\.)*"', '"raw"', src)
        bad = [(src.count("\n", 0, i) + 1, v) for i, t, v in JuliaLexer().get_tokens_unprocessed(src) if t in Token.Error]
        assert bad == [], (f, bad[:5])
