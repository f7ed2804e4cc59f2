"""The R .Call shim cannot be built here (no R headers), so it is at least syntax- and
type-checked against the real C ABI header with a stand-in for R's declarations."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_rshim_compiles_against_the_abi_header():
    r = subprocess.run(['gcc', '-fsyntax-only', '-Wall', '-Werror', '-I', os.path.join(ROOT, 'tests', 'rstub'),
                        '-I', os.path.join(ROOT, 'include'), os.path.join(ROOT, 'r', 'src', 'rshim.c')],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_header_is_plain_c():
    src = '#include "soundgen_b200.h"\nint main(void){ sgb_batch_desc d; (void)d; return sgb_version() ? 0 : 1; }\n'
    r = subprocess.run(['gcc', '-std=c99', '-fsyntax-only', '-Wall', '-Werror', '-I', os.path.join(ROOT, 'include'),
                        '-x', 'c', '-'], input=src, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
