"""The REFERENCE's own test scripts, unmodified, on top of `clockpunch_b200.install_as_punchclock()`.

The reference's tests are top-level scripts that print (almost nothing is asserted), so "the
reference's tests pass on the swapped package" is checked the only way they allow: every script that exercises a
mirrored module (tests/golden/make_golden_ref_scripts.py: SCRIPTS) must run to its end, and every NUMBER it prints
must equal what the same script prints on the all-reference stack (tests/golden/ref_scripts/*.txt, recorded by that
generator with every random source seeded identically), to the round-off of the mirrored arithmetic.  The text
between the numbers must be identical too (after class paths and object addresses are normalised).

CPU tests, reference tree required (skipped on the GPU box): the mirrors run on the TEST-ONLY doubles of
tests/_oracle_engine.py / tests/_oracle_lib.py; what the CUDA library computes is compared with the same oracle by
the `-m gpu` tests.
"""
import os
import re
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import make_golden_ref_scripts as mk  # noqa: E402

needs_ref = pytest.mark.skipif(not os.path.isfile(os.path.join(mk.REF, "punchclock", "environment", "env.py")),
                               reason="reference tree not available on this box")
NUM = re.compile(r"(?<![\w.])[-+]?(?:\d+\.\d*|\.\d+|\d+)(?:[eE][-+]?\d+)?(?![\w.])|\bnan\b|[-+]?\binf\b")


# Wording of error messages is not part of the contract (the exception TYPE is: the scripts catch ValueError and
# print it): lines of these messages are dropped from both sides before the comparison.
MESSAGE_LINES = {
    "tests/environment/test_env_parameters.py":
        r"^(Argument\(s\) was/were provided|\s+but values were also provided|\s+Resolve by setting|\s+agent_params\[|\s+to None or setting)",
}


def normalise(text: str, script: str) -> list:
    text = re.sub(r"0x[0-9a-f]+", "0xX", text)
    text = text.replace("clockpunch_b200.", "punchclock.")
    # a filter outside the hot path is an instance of the reference's own class, loaded under a private module name
    text = re.sub(r"punchclock\._reference_modules\.([a-z]+)_([a-z0-9_]+)\.", r"punchclock.\1.\2.", text)
    lines = text.splitlines()
    if script in MESSAGE_LINES:
        lines = [ln for ln in lines if not re.match(MESSAGE_LINES[script], ln)]
    return lines


def split(text: str):
    """(skeleton, numbers): the text with every number replaced by '#', and the numbers."""
    nums = [float(m.group(0)) for m in NUM.finditer(text)]
    skel = NUM.sub("#", text)
    skel = re.sub(r"[ \t]+", " ", skel)                     # column padding of array reprs depends on the digits
    skel = re.sub(r"#\.(?=[\s,\]\)])", "#", skel)
    return skel, np.array(nums)


@needs_ref
@pytest.mark.parametrize("script", mk.SCRIPTS)
def test_reference_script_prints_the_same_numbers_on_the_swapped_package(script):
    want_text = open(os.path.join(ROOT, "tests", "golden", "ref_scripts", mk.name_of(script))).read()
    rc, got_text, err = mk.run("seam", script)
    assert rc == 0, err[-4000:]
    want_lines, got_lines = normalise(want_text, script), normalise(got_text, script)
    assert len(got_lines) >= len(want_lines) > 0
    # (a script the stand-ins stop part-way on the all-reference stack is compared up to that point)
    got_lines = got_lines[: len(want_lines)]
    checked = 0
    for ln, (w, g) in enumerate(zip(want_lines, got_lines), 1):
        ws, wn = split(w)
        gs, gn = split(g)
        assert ws == gs, f"{script} line {ln}:\n  reference: {w}\n  swapped:   {g}"
        assert len(wn) == len(gn)
        if len(wn):
            np.testing.assert_allclose(gn, wn, rtol=2e-6, atol=1e-7, equal_nan=True,
                                       err_msg=f"{script} line {ln}:\n  reference: {w}\n  swapped:   {g}")
            checked += len(wn)
    assert checked > 0
