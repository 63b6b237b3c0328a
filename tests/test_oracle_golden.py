"""The oracle (oracle/smallz4_oracle.c) against digests of the unmodified reference's output.

tests/golden/golden.json was produced by tests/golden/make_golden.py from oracle/_ref
(the reference compiled from /root/reference).  This pins the oracle without needing the
reference at run time.
"""
import json
import os

import pytest

from golden_util import case_dict, case_id, case_input, digest
from oracle_lib import oracle_compress, oracle_decompress

with open(os.path.join(os.path.dirname(__file__), "golden", "golden.json")) as _f:
    CASES = json.load(_f)["cases"]

# level -9 of a few MiB costs the scalar oracle tens of seconds; keep the default CPU suite short
FAST = [c for c in CASES if c["size"] <= 400_000 or c["level"] <= 3]
SLOW = [c for c in CASES if c not in FAST]


def _check(c):
    data = case_input(c)
    d = case_dict(c)
    frame, stats = oracle_compress(data, c["level"], c["legacy"], d)
    assert len(frame) == c["frame_size"]
    assert digest(frame) == c["sha256"]
    if "frame_hex" in c:
        assert frame.hex() == c["frame_hex"]
    # the decoder restatement (smallz4cat) must give the input back.  Dictionary frames are
    # excluded: the reference's -D output does not decode with its own smallz4cat (DESIGN.md
    # "Q-dict", shown by test_reference_dictionary_frames_do_not_round_trip below)
    if not c["dict"] and not (c["legacy"] and c["level"] == 0):
        assert oracle_decompress(frame, len(data), d) == data
    assert stats["oob_first_reads"] == 0      # no case depends on the reference's out-of-bounds read


@pytest.mark.parametrize("c", FAST, ids=case_id)
def test_oracle_matches_reference_digest(c):
    _check(c)


@pytest.mark.slow
@pytest.mark.parametrize("c", SLOW, ids=case_id)
def test_oracle_matches_reference_digest_slow(c):
    _check(c)


def test_reference_dictionary_frames_do_not_round_trip():
    """Documents reference behaviour Q-dict: with -D the reference writes its chain ring with the
    block-relative index (smallz4.h:656) but reads it with the absolute one (smallz4.h:190), which a
    dictionary shifts by 65535.  The frames it produces are reproduced bit-exactly (digests above)
    and, exactly like the reference's own, they do not decode back to the input."""
    c = next(c for c in CASES if c["dict"] and c["dict"][2] == 65536 and c["level"] == 9 and c["kind"] == "text")
    data, d = case_input(c), case_dict(c)
    frame, _ = oracle_compress(data, c["level"], c["legacy"], d)
    assert digest(frame) == c["sha256"]
    try:
        back = oracle_decompress(frame, len(data) + 65536, d)
    except AssertionError:
        back = None
    assert back != data
